export PYTHONPATH=.
timeout 900 python tools/pdl_probe.py 2>&1
timeout 300 python tools/gpu_bringup.py perf 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 90 --csv --log-file gpurun_out/launches_vision_r1b.csv python tools/profile_step.py vision 3 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 176 -c 88 --csv --log-file gpurun_out/launches_text_r1b.csv python tools/profile_step.py text 3 2>&1 | tail -1
