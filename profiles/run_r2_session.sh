#!/bin/bash
# One GPU session of round 2: bench line, ncu launch list of the same command, one `ncu --set full` capture per rig.
# usage: bash profiles/run_r2_session.sh <tag>
set -u
tag=${1:-r2}
out=gpurun_out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; tail -c 600 $out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_reference_arm.json 2> $out/${tag}_reference_arm.err
echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
echo "launch list rc=$?"
for spec in ${NCU_SPECS:-humanoid22:303104 chain64:65536 quad80:65536}; do
  rig=${spec%%:*}; poses=${spec##*:}
  ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -f -o $out/${tag}_${rig} \
      python profiles/run_kernel.py --rig $rig --poses $poses --launches 3 > $out/${tag}_ncu_${rig}.log 2>&1
  echo "ncu $rig rc=$?"
done
ls -la $out | grep ${tag}
