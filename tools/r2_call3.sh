#!/bin/bash
# gpurun --gpus 8 call: 2-GPU NCCL parity test, cfg4 / cfg5 strong scaling (N = 1, 2, 4, 8), pairs at N = 8
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/r2c_smi.txt
# the tensor-core similarity path is new: validate it on one GPU first, fall back to the fp32 SIMT kernel if it fails
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -k "similarity or clip_forward or evaluation" > gpurun_out/r2c_pytest_sim.log 2>&1
if [ $? -ne 0 ]; then export PLIP_SIM_SIMT=1; echo "TC similarity FAILED: using SIMT" >> gpurun_out/r2c_rc.txt; fi
tail -3 gpurun_out/r2c_pytest_sim.log
( time python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2c_pytest_multi.log 2>&1
tail -3 gpurun_out/r2c_pytest_multi.log
run() {  # N configs steps
  local n=$1 cfgs=$2 extra=$3 port=$((29500 + $1))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --config $cfgs --quick --no-cpu-baseline $extra > gpurun_out/r2c_bench_${cfgs//,/_}_n$n.jsonl 2> gpurun_out/r2c_bench_${cfgs//,/_}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfgs --quick --no-cpu-baseline $extra \
      > gpurun_out/r2c_bench_${cfgs//,/_}_n$n.jsonl 2> gpurun_out/r2c_bench_${cfgs//,/_}_n$n.err
  fi
  echo "N=$n $cfgs rc=$?" >> gpurun_out/r2c_rc.txt
}
run 8 pairs,cfg4,cfg5 ""
run 1 cfg4,cfg5 ""
run 4 cfg4,cfg5 ""
run 2 cfg4,cfg5 ""
cat gpurun_out/r2c_rc.txt
ls -la gpurun_out | grep r2c
