export PYTHONPATH=.
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err; head -c 1500 gpurun_out/bench_final.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_crop -c 1 -f -o gpurun_out/r1_resize_full python tools/resize_probe.py 2>&1 | tail -3
