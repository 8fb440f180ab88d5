export PYTHONPATH=.
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1c.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_r1c.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 134 -c 67 --csv --log-file gpurun_out/launches_vision_r1f.csv python tools/profile_step.py vision 3 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 132 -c 66 --csv --log-file gpurun_out/launches_text_r1f.csv python tools/profile_step.py text 3 2>&1 | tail -1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 404 -c 272 --csv --log-file gpurun_out/launches_bench_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
