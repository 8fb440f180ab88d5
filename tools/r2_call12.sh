#!/bin/bash
# experimental two-pair clusters with a multicast W tile (PLIP_GEMM_QUAD=1): parity under a watchdog, then A-B-A
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( PLIP_GEMM_QUAD=1 PLIP_DEBUG=1 timeout 240 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "golden or hidden_states or clip_forward" ) > gpurun_out/r2n_pytest_quad.log 2>&1
echo "quad pytest rc=$?" > gpurun_out/r2n_rc.txt
tail -15 gpurun_out/r2n_pytest_quad.log
if grep -q "rc=0" gpurun_out/r2n_rc.txt; then
  ( PLIP_GEMM_QUAD=1 timeout 300 python -m pytest tests/test_gpu_parity_sizes.py tests/test_gpu_kernels.py -m gpu -x -q ) > gpurun_out/r2n_pytest_quad2.log 2>&1
  echo "quad pytest2 rc=$?" >> gpurun_out/r2n_rc.txt
  python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2n_bench_a1.json 2> gpurun_out/r2n_bench_a1.err
  PLIP_GEMM_QUAD=1 timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2n_bench_quad.json 2> gpurun_out/r2n_bench_quad.err
  python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2n_bench_a2.json 2> gpurun_out/r2n_bench_a2.err
fi
cat gpurun_out/r2n_rc.txt
