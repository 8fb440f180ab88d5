export PYTHONPATH=.
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/gpu_bringup.py vision text perf 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 134 -c 67 --csv --log-file gpurun_out/launches_vision_r1c.csv python tools/profile_step.py vision 3 2>&1 | tail -1
