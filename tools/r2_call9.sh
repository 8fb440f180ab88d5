#!/bin/bash
# final validation of HEAD: full GPU suite, smoke(), driver-style bench pair
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2k_pytest.log 2>&1
tail -4 gpurun_out/r2k_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2k_smoke.log 2>&1
tail -2 gpurun_out/r2k_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -c 400 gpurun_out/r2k_bench.err
