export PYTHONPATH=.
for pf in 0 1; do echo "== prefetch_a=$pf"; for c in "2 256 0 51200 2304 768" "2 256 1 51200 3072 768" "2 256 2 51200 768 3072" "2 192 2 51200 768 3072" "2 192 2 51200 768 768" "2 256 0 78848 1536 512" "2 256 1 78848 2048 512" "2 256 2 78848 512 2048" "2 256 2 78848 512 512" "2 128 2 78848 512 512"; do PLIP_GEMM_PREFETCH_A=$pf timeout 120 python tools/gpu_gemm_check.py $c 2>&1 | tail -1 | sed -e 's/"ref_max": [0-9.]*, //' -e 's/"max_abs_err": [0-9.e-]*, //'; done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 134 -c 67 --csv --log-file gpurun_out/launches_vision_r1e.csv python tools/profile_step.py vision 3 2>&1 | tail -1
for pf in 0 1; do PLIP_GEMM_PREFETCH_A=$pf timeout 300 python tools/gpu_bringup.py perf 2>&1 | tail -4; done
