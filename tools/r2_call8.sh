#!/bin/bash
# A/B: residual prefetch also for the long-K residual GEMM (fc2, N tile 256: 5 -> 3 operand stages) — -DPLIP_RPF_ALL build
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err
cp plip_b200/libplip_b200.so /tmp/lib_default.so
( time PLIP_EXTRA_NVCC_FLAGS=-DPLIP_RPF_ALL python -m plip_b200.build ) > gpurun_out/r2j_build_all.log 2>&1
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -k "gemm or layernorm_folded or hidden_states or golden" > gpurun_out/r2j_pytest_all.log 2>&1
tail -3 gpurun_out/r2j_pytest_all.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2j_bench_all.json 2> gpurun_out/r2j_bench_all.err
cp /tmp/lib_default.so plip_b200/libplip_b200.so
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2j_bench_default2.json 2> gpurun_out/r2j_bench_default2.err
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 73 -c 5 -o gpurun_out/r2j_vision_layer \
    python tools/profile_step.py vision 2 > gpurun_out/r2j_ncu_vision.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 71 -c 5 -o gpurun_out/r2j_text_layer \
    python tools/profile_step.py text 2 > gpurun_out/r2j_ncu_text.log 2>&1
