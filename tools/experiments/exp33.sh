export PYTHONPATH=.
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for d in 0 1 0 1; do echo "== xb_tma=$d"; PLIP_GEMM_XB_TMA=$d timeout 300 python tools/gpu_bringup.py perf 2>&1 | grep -E "bf16\", \"micro_batch\": 1024|perf_text"; done
