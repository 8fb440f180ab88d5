export PYTHONPATH=.
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -15
timeout 300 python tools/resize_probe.py gpurun_out/resize_probe.json 2>&1 | tail -8
