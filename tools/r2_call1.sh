#!/bin/bash
# gpurun call 1 of round 2: new parity tests, bench contract, in-step profile, ncu launch list + captures, PDL stress
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err
# launch list of the bench command itself (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file gpurun_out/r2a_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
# full captures: one vision layer's 4 GEMMs + attention (vision), then text attention + text GEMMs
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 73 -c 5 -o gpurun_out/r2a_vision_layer \
    python tools/profile_step.py vision 2 > gpurun_out/r2a_ncu_vision.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attention_kernel' -s 71 -c 5 -o gpurun_out/r2a_text_layer \
    python tools/profile_step.py text 2 > gpurun_out/r2a_ncu_text.log 2>&1
# PDL stress (bounded): 200 full steps with PDL on, then off
for pdl in 1 0; do
  ( time PLIP_PDL=$pdl timeout 150 python bench.py --steps 200 --warmup 5 --quick --no-cpu-baseline ) > gpurun_out/r2a_pdl$pdl.json 2> gpurun_out/r2a_pdl$pdl.err
  echo "pdl=$pdl rc=$?" >> gpurun_out/r2a_pdl_rc.txt
done
ls -la gpurun_out | head -40
