#!/bin/bash
# fp32-output epilogues (out_proj / fc2 / patch / similarity): rolling TMEM prefetch + early release.  Parity, then A-B-A.
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_parity_sizes.py -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2r_rc.txt
tail -5 gpurun_out/r2r_pytest.log
for v in a1:0 b:1 a2:0; do
  n=${v%%:*}; f=${v##*:}
  PLIP_GEMM_F32_SERIAL=$f python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2r_bench_$n.json 2> gpurun_out/r2r_bench_$n.err
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
d=[json.loads(x) for x in open(f'gpurun_out/r2r_bench_{n}.json') if x.startswith('{')][0]
ks={k['kernel']:k['us'] for k in d['extra']['kernels_in_step']}
print(n, round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in ks.items() if 'out_proj' in k or 'fc2' in k or 'patch' in k})
PY
done
cat gpurun_out/r2r_rc.txt
