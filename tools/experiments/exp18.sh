export PYTHONPATH=.
for c in "2 256 7 51200 2304 768" "2 256 0 51200 2304 768" "2 256 1 51200 3072 768" "2 256 2 51200 768 3072" "2 192 2 51200 768 768" "2 256 1 78848 2048 512"; do timeout 120 python tools/gpu_gemm_check.py $c 2>&1 | tail -1 | sed -e 's/"ref_max": [0-9.]*, //' -e 's/"max_abs_err": [0-9.e-]*, //'; done
timeout 300 python tools/gpu_bringup.py perf 2>&1 | tail -4
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
