export PYTHONPATH=.
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 120 --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm_epilogues or layernorm_folded or attention or im2col or layernorm or similarity" 2>&1 | tail -15
echo "== model"
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 120 --error-exitcode 3 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "image_embeddings or text_embeddings or microbatching" 2>&1 | tail -12
