#!/bin/bash
# last-layer pruning option: parity + A/B inside bench.py (extra.last_layer_pruning_opt_in); cluster-of-4 occupancy print
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_abi.py -m gpu -q > gpurun_out/r2o_pytest_model.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2o_rc.txt
tail -25 gpurun_out/r2o_pytest_model.log
PLIP_GEMM_QUAD=1 PLIP_DEBUG=1 timeout 120 python - > gpurun_out/r2o_quad_groups.log 2>&1 <<'PY'
import torch
from plip_b200 import synthetic as synth
from plip_b200.engine import Engine
e = Engine(synth.make_state_dict(), max_micro_batch=64)
e.encode_images(synth.pixel_values(8).cuda()); e.encode_text(synth.token_ids(8)[0].cuda()); torch.cuda.synchronize()
PY
grep "gemm<" gpurun_out/r2o_quad_groups.log | sort | uniq -c
python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
tail -3 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=[json.loads(x) for x in open('gpurun_out/r2o_bench.json') if x.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['extra'].get('last_layer_pruning_opt_in'))
PY
cat gpurun_out/r2o_rc.txt
