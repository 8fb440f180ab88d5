#!/bin/bash
# software-pipelined 16-bit epilogue (rolling TMEM prefetch + early accumulator release): parity, probe, step
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fp16.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2q_rc.txt
tail -5 gpurun_out/r2q_pytest.log
PROBE_VARIANTS=full,no_stores timeout 300 python tools/epilogue_probe.py gpurun_out/r2q_epilogue_probe.json > gpurun_out/r2q_epilogue_probe.log 2>&1
cat gpurun_out/r2q_epilogue_probe.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -3 gpurun_out/r2q_bench.err
python - <<'PY'
import json
d=[json.loads(x) for x in open('gpurun_out/r2q_bench.json') if x.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['clocks'])
for k in d['extra']['kernels_in_step']:
    print(k['kernel'], round(k['us'],1), round(k.get('frac_of_roofline') or 0,3))
PY
cat gpurun_out/r2q_rc.txt
