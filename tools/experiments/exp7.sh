export PYTHONPATH=.
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())"
for c in "2 256 1 51200 3072 768" "2 256 0 51200 2304 768"; do python tools/gpu_gemm_check.py $c 2>&1 | tail -1; done
python tools/gpu_bringup.py attn_vision attn_text attn_text_mask vision text perf 2>&1
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
