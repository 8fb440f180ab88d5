#!/bin/bash
# gpurun call 2: attention v2 parity + A/B, fc2 N-tile experiment, ncu captures of one layer (vision + text)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.log 2>&1
tail -4 gpurun_out/r2b_pytest.log
for cfg in "v2:" "v1:PLIP_ATT_V1=1" "v2c3:PLIP_ATT_CTAS=3" "fc2bn192:PLIP_GEMM_BN_FC2=192"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err
done
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 73 -c 5 -o gpurun_out/r2b_vision_layer \
    python tools/profile_step.py vision 2 > gpurun_out/r2b_ncu_vision.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 71 -c 5 -o gpurun_out/r2b_text_layer \
    python tools/profile_step.py text 2 > gpurun_out/r2b_ncu_text.log 2>&1
PLIP_ATT_V1=1 ncu --set full --clock-control none --import-source on -k regex:'attention_kernel' -s 14 -c 1 -o gpurun_out/r2b_vision_attn_v1 \
    python tools/profile_step.py vision 2 > gpurun_out/r2b_ncu_v1.log 2>&1
ls -la gpurun_out | grep r2b
