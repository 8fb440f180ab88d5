#!/bin/bash
# A-B-A: both TMEM halves of a 64-column epilogue block in flight (default) vs one wait per half (-DPLIP_EPI_SERIAL_LD)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fp16.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1
tail -3 gpurun_out/r2l_pytest.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2l_bench_a1.json 2> gpurun_out/r2l_bench_a1.err
cp plip_b200/libplip_b200.so /tmp/lib_default.so
PLIP_EXTRA_NVCC_FLAGS=-DPLIP_EPI_SERIAL_LD python -m plip_b200.build > gpurun_out/r2l_build.log 2>&1
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2l_bench_b.json 2> gpurun_out/r2l_bench_b.err
cp /tmp/lib_default.so plip_b200/libplip_b200.so
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2l_bench_a2.json 2> gpurun_out/r2l_bench_a2.err
