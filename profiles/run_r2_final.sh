#!/bin/bash
# Final round-2 session on one GPU: GPU tests, smoke, the driver's bench command and reference arm.
set -u
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/r2_final_gpu_tests.log 2>&1; tail -2 $out/r2_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2_final_smoke.log 2>&1; tail -1 $out/r2_final_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/r2_final_reference_arm.json 2> $out/r2_final_reference_arm.err; echo "reference arm rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/r2_final_bench.json 2> $out/r2_final_bench.err; echo "bench rc=$?"
