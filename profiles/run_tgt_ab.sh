#!/bin/bash
# A/B of the target prefetch slots (MBIK_TGT_PREFETCH=0/1, same library): humanoid22 lockstep kernel, device-resident launches,
# interleaved so that box clocks hit both arms alike; then the abi-feature + parity tests with the slots on.
mkdir -p gpurun_out
out=gpurun_out/r2_exp_tgt_prefetch.log
: > $out
for rep in 1 2 3; do
  for v in 0 1; do
    echo "== MBIK_TGT_PREFETCH=$v rep $rep" >> $out
    MBIK_TGT_PREFETCH=$v timeout 300 python profiles/run_variants.py --child --rig humanoid22 --poses 1048576 --launches 6 >> $out 2>&1
  done
done
for v in 0 1; do
  echo "== MBIK_TGT_PREFETCH=$v poses 303104" >> $out
  MBIK_TGT_PREFETCH=$v timeout 300 python profiles/run_variants.py --child --rig humanoid22 --poses 303104 --launches 6 >> $out 2>&1
  echo "== MBIK_TGT_PREFETCH=$v poses 4096 (SP off via MBIK_SP=0 not set: whatever launch_solve picks)" >> $out
  MBIK_TGT_PREFETCH=$v timeout 300 python profiles/run_variants.py --child --rig humanoid22 --poses 4096 --launches 20 >> $out 2>&1
done
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_exp_tgt_gpu_tests.log 2>&1
tail -3 gpurun_out/r2_exp_tgt_gpu_tests.log >> $out
cat $out
