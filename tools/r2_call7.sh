#!/bin/bash
# residual-prefetch (cp.async double buffer in the out_proj epilogue): full parity suite, then A/B against a -DPLIP_NO_RPF build
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2i_pytest.log 2>&1
tail -4 gpurun_out/r2i_pytest.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2i_bench_rpf.json 2> gpurun_out/r2i_bench_rpf.err
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel<2, 192, 2|gemm_kernel<2, 128, 2' -s 14 -c 1 -o gpurun_out/r2i_vision_outproj_rpf \
    python tools/profile_step.py vision 2 > gpurun_out/r2i_ncu_v.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel<2, 192, 2|gemm_kernel<2, 128, 2' -s 14 -c 1 -o gpurun_out/r2i_text_outproj_rpf \
    python tools/profile_step.py text 2 > gpurun_out/r2i_ncu_t.log 2>&1
cp plip_b200/libplip_b200.so /tmp/lib_rpf.so
( time PLIP_EXTRA_NVCC_FLAGS=-DPLIP_NO_RPF python -m plip_b200.build ) > gpurun_out/r2i_build_norpf.log 2>&1
PLIP_EXTRA_NVCC_FLAGS=-DPLIP_NO_RPF python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2i_bench_norpf.json 2> gpurun_out/r2i_bench_norpf.err
cp /tmp/lib_rpf.so plip_b200/libplip_b200.so
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2i_bench_rpf2.json 2> gpurun_out/r2i_bench_rpf2.err
