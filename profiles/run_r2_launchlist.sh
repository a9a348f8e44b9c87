#!/bin/bash
# Last tree of round 2: ncu launch list of the bench command (all launches up to 2000) and one `ncu --set full` capture of the
# headline kernel.  usage: bash profiles/run_r2_launchlist.sh <tag>
set -u
tag=${1:-r2_h}
out=gpurun_out
mkdir -p $out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_plain_bench.json 2> $out/${tag}_plain_bench.err
echo "plain run rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
echo "launch list rc=$?"
python profiles/run_kernel.py --rig humanoid22 --poses 303104 --launches 3 > $out/${tag}_plain_kernel.log 2>&1
echo "plain kernel run rc=$?"
ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -f -o $out/${tag}_humanoid22 \
    python profiles/run_kernel.py --rig humanoid22 --poses 303104 --launches 3 > $out/${tag}_ncu_humanoid22.log 2>&1
echo "ncu humanoid22 rc=$?"
ls -la $out | grep ${tag}
