export PYTHONPATH=.
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1e.json | cut -c1-200
