export PYTHONPATH=.
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1g.json | cut -c1-150
