#!/bin/bash
# final 1-GPU evidence run: compute-sanitizer over the new kernels, driver-style bench lines (reference arm first), cfg3 line,
# launch list of the bench command
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -x -q \
    -k "attention or similarity or gemm or layernorm or im2col" ) > gpurun_out/r2f_sanitizer_kernels.log 2>&1
echo "sanitizer kernels rc=$?" >> gpurun_out/r2f_rc.txt
( time compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_model.py -m gpu -x -q \
    -k "graph or clip_forward or image_embeddings" ) > gpurun_out/r2f_sanitizer_model.log 2>&1
echo "sanitizer model rc=$?" >> gpurun_out/r2f_rc.txt
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 560 --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
TOPK_DUP=1 python tools/topk_probe.py > gpurun_out/r2f_topk_dup.log 2>&1
python bench.py --config cfg5 --tiles 125000 --steps 2 --quick --no-cpu-baseline > gpurun_out/r2f_bench_cfg5_125k.json 2> gpurun_out/r2f_bench_cfg5_125k.err
cat gpurun_out/r2f_rc.txt gpurun_out/r2f_topk_dup.log
