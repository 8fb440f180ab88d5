export PYTHONPATH=.
for d in 0 1 2; do echo "== st_mode=$d"; for c in "2 256 0 51200 2304 768" "2 256 1 51200 3072 768"; do PLIP_GEMM_ST=$d timeout 120 python tools/gpu_gemm_check.py $c 2>&1 | tail -1 | sed -e 's/"ref_max": [0-9.]*, //' -e 's/"max_abs_err": [0-9.e-]*, //'; done; done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
