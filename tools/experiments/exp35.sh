export PYTHONPATH=.
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -15
for kb in 72 96 48 128; do echo "== PLIP_RESIZE_SMEM_KB=$kb"; PLIP_RESIZE_SMEM_KB=$kb timeout 300 python tools/resize_probe.py gpurun_out/resize_probe_$kb.json 2>&1 | tail -5; done  # env knob removed afterwards (rule is now per image)
