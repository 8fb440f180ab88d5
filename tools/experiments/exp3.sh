export PYTHONPATH=.
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -o gpurun_out/gemm_cg2 python tools/gpu_gemm_check.py 2 256 0 51200 2304 768 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -o gpurun_out/gemm_cg1 python tools/gpu_gemm_check.py 1 256 0 51200 2304 768 2>&1 | tail -2
