#!/bin/bash
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python tools/epilogue_probe.py gpurun_out/r2p_epilogue_probe.json > gpurun_out/r2p_epilogue_probe.log 2>&1
cat gpurun_out/r2p_epilogue_probe.log
