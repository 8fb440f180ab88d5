export PYTHONPATH=.
for c in "2 256 1 51200 3072 768" "2 256 0 51200 2304 768" "2 256 2 51200 768 3072" "2 192 2 51200 768 768" "2 256 2 51200 768 768" "2 256 4 51200 768 768"; do timeout 120 python tools/gpu_gemm_check.py $c 2>&1 | tail -1 | sed -e 's/"ref_max": [0-9.]*, //'; done
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/gpu_bringup.py perf 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 134 -c 67 --csv --log-file gpurun_out/launches_vision_r1d.csv python tools/profile_step.py vision 3 2>&1 | tail -1
