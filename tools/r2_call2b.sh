#!/bin/bash
# quick 1-GPU validation of the packed-fp32 epilogues + fp16 operand mode before the 8-GPU call
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2d_pytest.log 2>&1
tail -6 gpurun_out/r2d_pytest.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2d_bench_bf16.json 2> gpurun_out/r2d_bench_bf16.err
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline --operands fp16 > gpurun_out/r2d_bench_fp16.json 2> gpurun_out/r2d_bench_fp16.err
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2d_bench_bf16_b.json 2> gpurun_out/r2d_bench_bf16_b.err
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 73 -c 5 -o gpurun_out/r2d_vision_layer \
    python tools/profile_step.py vision 2 > gpurun_out/r2d_ncu_vision.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 71 -c 5 -o gpurun_out/r2d_text_layer \
    python tools/profile_step.py text 2 > gpurun_out/r2d_ncu_text.log 2>&1
