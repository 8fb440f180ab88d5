export PYTHONPATH=.
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -4
for pf in 0 1; do for shape in "2 192 2 51200 768 768" "2 256 2 51200 768 3072" "2 128 2 78848 512 512" "2 256 2 78848 512 2048"; do
  echo "== pf_next=$pf $shape: $(PLIP_GEMM_PF_NEXT=$pf timeout 120 python tools/gpu_gemm_check.py $shape 2>&1 | tail -1)"; done; done
for pf in 0 1 0 1; do echo "== pf_next=$pf"; PLIP_GEMM_PF_NEXT=$pf timeout 300 python tools/gpu_bringup.py perf 2>&1 | grep -E "bf16\", \"micro_batch\": 1024|perf_text"; done
