export PYTHONPATH=.
for c in "2 192 2 51200 768 3072" "2 256 2 51200 768 3072" "2 128 2 51200 768 3072" "2 192 2 51200 768 768" "2 256 2 51200 768 768" "2 128 2 51200 768 768" "2 192 4 1000 768 768" "2 256 1 51200 3072 768" "2 256 0 51200 2304 768" "2 192 0 51200 2304 768"; do python tools/gpu_gemm_check.py $c 2>&1 | tail -1 | sed -e 's/"ref_max": [0-9.]*, //'; done
python tools/gpu_bringup.py attn_vision attn_text attn_text_mask vision text perf 2>&1
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
