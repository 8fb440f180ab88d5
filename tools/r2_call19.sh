#!/bin/bash
# opt-in direct patch embedding (PLIP_PATCH_DIRECT=1: 4-D tensor map over the bf16 pixels, no im2col matrix):
# parity of the changed patch-GEMM instances (default and direct), then timing of both
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2u_rc.txt
tail -12 gpurun_out/r2u_pytest.log
timeout 120 python tools/patch_direct_probe.py > gpurun_out/r2u_probe_default.json 2> gpurun_out/r2u_probe_default.err
PLIP_PATCH_DIRECT=1 timeout 120 python tools/patch_direct_probe.py > gpurun_out/r2u_probe_direct.json 2> gpurun_out/r2u_probe_direct.err
cat gpurun_out/r2u_probe_default.json gpurun_out/r2u_probe_direct.json
tail -3 gpurun_out/r2u_probe_direct.err
cat gpurun_out/r2u_rc.txt
