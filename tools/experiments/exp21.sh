export PYTHONPATH=.
for d in 0 1; do echo "== tma_store=$d"; for c in "2 256 0 51200 2304 768" "2 256 1 51200 3072 768" "2 256 0 1000 768 768" "2 256 1 78848 2048 512"; do PLIP_GEMM_TMA_STORE=$d timeout 120 python tools/gpu_gemm_check.py $c 2>&1 | tail -1 | sed -e 's/"ref_max": [0-9.]*, //'; done; done
PLIP_GEMM_TMA_STORE=1 timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for d in 0 1; do PLIP_GEMM_TMA_STORE=$d timeout 300 python tools/gpu_bringup.py perf 2>&1 | tail -4; done
