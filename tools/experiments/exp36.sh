export PYTHONPATH=.
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -15
timeout 300 python tools/resize_probe.py gpurun_out/resize_probe_b512.json 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_resize.py -q -m gpu -k "mid_word or matches_pil or rejects" 2>&1 | tail -8
