export PYTHONPATH=.
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -4
timeout 300 python tools/resize_probe.py gpurun_out/resize_probe_v3.json 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_crop -c 1 -f -o gpurun_out/r1_resize_full_v3 python tools/resize_probe.py 2>&1 | tail -2
