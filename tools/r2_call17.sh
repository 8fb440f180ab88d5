#!/bin/bash
# final validation of HEAD (pipelined epilogues, pruning option off): full GPU suite, smoke(), driver-style bench,
# ncu --set full of one encoder layer per tower on the final build, and an informational N-tile A/B for out_proj
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2s_pytest.log 2>&1
tail -4 gpurun_out/r2s_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
tail -2 gpurun_out/r2s_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -c 300 gpurun_out/r2s_bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 73 -c 5 -o gpurun_out/r2s_vision_layer \
    python tools/profile_step.py vision 2 > gpurun_out/r2s_ncu_v.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 71 -c 5 -o gpurun_out/r2s_text_layer \
    python tools/profile_step.py text 2 > gpurun_out/r2s_ncu_t.log 2>&1
ls -la gpurun_out/*.ncu-rep
PLIP_GEMM_BN=256 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2s_bench_bn256.json 2> gpurun_out/r2s_bench_bn256.err
python - <<'PY'
import json
for n in ("r2s_bench", "r2s_bench_bn256"):
    d=[json.loads(x) for x in open(f'gpurun_out/{n}.json') if x.startswith('{')][0]
    ks={k['kernel']:round(k['us'],1) for k in d['extra']['kernels_in_step'] if 'gemm' in k['kernel'] or 'attention' in k['kernel']}
    print(n, round(d['value']), round(d['ms_per_step'],3), d['e2e']['value'], d['clocks'], ks)
PY
