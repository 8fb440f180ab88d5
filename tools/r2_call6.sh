#!/bin/bash
# gpurun --gpus 8, final code: weak-scaling step + cfg4 / cfg5 at N = 8, 4, 2, 1 and the 2-rank NCCL parity test
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2h_pytest_multi.log 2>&1
tail -3 gpurun_out/r2h_pytest_multi.log
run() {
  local n=$1 cfgs=$2 port=$((29600 + $1))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --config $cfgs --quick --no-cpu-baseline > gpurun_out/r2h_bench_n$n.jsonl 2> gpurun_out/r2h_bench_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfgs --quick --no-cpu-baseline \
      > gpurun_out/r2h_bench_n$n.jsonl 2> gpurun_out/r2h_bench_n$n.err
  fi
  echo "N=$n $cfgs rc=$?" >> gpurun_out/r2h_rc.txt
}
run 8 pairs,cfg4,cfg5
run 1 pairs,cfg4,cfg5
run 4 pairs,cfg4,cfg5
run 2 pairs,cfg4,cfg5
cat gpurun_out/r2h_rc.txt
