export PYTHONPATH=.
for g in 0 74 72 70 64 37; do
  echo "== groups=$g"; PLIP_DEBUG=1 PLIP_GEMM_GROUPS=$g python tools/gpu_gemm_check.py 2 256 0 51200 2304 768 2>&1 | tail -2
done
echo "== cg1 groups 148/140/128"; for g in 148 140 128; do PLIP_GEMM_GROUPS=$g python tools/gpu_gemm_check.py 1 256 0 51200 2304 768 2>&1 | tail -1; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
