export PYTHONPATH=.
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-context 2>&1 | tail -2 | tee gpurun_out/bench_2gpu_r1c.json | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-context --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
