export PYTHONPATH=.
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 2>&1 | tail -16
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1f.json | cut -c1-150
