#!/bin/bash
# 1-GPU: validate epilogue micro-opts + CUDA graphs + TC similarity in the full suite; bench (full extras); ncu of one layer
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2e_pytest.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 300 gpurun_out/r2e_bench.err
PLIP_BENCH_MB=512 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2e_bench_mb512.json 2> gpurun_out/r2e_bench_mb512.err
python bench.py --config cfg5 --tiles 125000 --steps 2 --quick --no-cpu-baseline > gpurun_out/r2e_bench_cfg5_125k.json 2> gpurun_out/r2e_bench_cfg5_125k.err
PLIP_GRAPH_MAX=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-context > gpurun_out/r2e_bench_nograph.json 2> gpurun_out/r2e_bench_nograph.err
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 73 -c 5 -o gpurun_out/r2e_vision_layer \
    python tools/profile_step.py vision 2 > gpurun_out/r2e_ncu_vision.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 71 -c 5 -o gpurun_out/r2e_text_layer \
    python tools/profile_step.py text 2 > gpurun_out/r2e_ncu_text.log 2>&1
