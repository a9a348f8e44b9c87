#!/bin/bash
# The driver's scaling run, reproduced: bench.py under torchrun at N = 2, 4, 8 on one box (N = 1 comes from run_r2_session.sh).
# usage: bash profiles/run_r2_scale.sh <tag> [Ns...]
tag=${1:-r2}; shift
Ns=${@:-2 4 8}
for N in $Ns; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
  echo "N=$N rc=$?"; tail -c 300 gpurun_out/${tag}_bench_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_${N}gpu.json"))
    print("  value %.1f M  e2e %.1f M  strong %.1f M  stream(dl) %s  single-process %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["strong_scaling_1M_batch"]["value"]/1e6,
          d["stream_e2e"].get("value_with_pose_download"), (d.get("single_process_multi_gpu_e2e") or {}).get("value")))
except Exception as e:
    print("  no line:", e)
PY
done
