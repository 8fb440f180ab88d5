export PYTHONPATH=.
nvidia-smi -L
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/gpu_bringup.py perf 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_2gpu_r1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2
