export PYTHONPATH=.
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_r1a.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_r1a.json
