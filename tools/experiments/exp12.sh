export PYTHONPATH=.
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 132 -c 66 --csv --log-file gpurun_out/launches_text_r1d.csv python tools/profile_step.py text 3 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 14 -c 1 -o gpurun_out/attn_vision_r1 python tools/profile_step.py vision 2 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 70 -c 4 -o gpurun_out/gemm_layer_r1 python tools/profile_step.py vision 2 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1b.json
