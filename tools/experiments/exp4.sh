export PYTHONPATH=.
python tools/gpu_gemm_check.py 2>&1 | grep -E "rc=|TIMEOUT" | sed -e 's/"ref_max": [0-9.]*, //' 
python tools/gpu_bringup.py 2>&1
