export PYTHONPATH=.
for c in "2 256 2 51200 768 3072" "1 256 2 51200 768 3072" "2 256 2 51200 768 768" "2 256 0 51200 2304 768" "2 256 1 51200 3072 768" "2 256 3 50176 768 3072"; do python tools/gpu_gemm_check.py $c 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 90 --csv --log-file gpurun_out/launches_vision_r1a.csv python tools/profile_step.py vision 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 174 -c 87 --csv --log-file gpurun_out/launches_text_r1a.csv python tools/profile_step.py text 3 2>&1 | tail -1
python tools/gpu_bringup.py perf 2>&1 | tail -6
