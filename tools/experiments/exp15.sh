export PYTHONPATH=.
timeout 300 python tools/two_stream_probe.py 2>&1 | tail -6
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
