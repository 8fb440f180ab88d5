#!/bin/bash
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2g_pytest.log 2>&1
tail -8 gpurun_out/r2g_pytest.log
python tools/topk_probe.py > gpurun_out/r2g_topk_tc.log 2>&1
PLIP_SIM_SIMT=1 python tools/topk_probe.py > gpurun_out/r2g_topk_simt.log 2>&1
python tools/topk_probe.py 2000 40000 > gpurun_out/r2g_topk_tc_small.log 2>&1
cat gpurun_out/r2g_topk_*.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2g_topk_launches.csv python tools/topk_probe.py > /dev/null 2>&1
