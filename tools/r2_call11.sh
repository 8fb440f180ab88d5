#!/bin/bash
# last call of the round: full GPU suite on HEAD, cfg3 line (micro-batch-wise upload in e2e), fp16-operand line
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2m_pytest.log 2>&1
tail -4 gpurun_out/r2m_pytest.log
python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_cfg3.json 2> gpurun_out/r2m_bench_cfg3.err
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline --operands fp16 > gpurun_out/r2m_bench_fp16.json 2> gpurun_out/r2m_bench_fp16.err
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2m_bench_bf16.json 2> gpurun_out/r2m_bench_bf16.err
