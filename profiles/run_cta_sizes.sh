#!/bin/bash
# ms per launch of the thread-per-pose kernel for forced CTA sizes (MBIK_THREADS), large rigs: does a smaller resident
# batch (state closer to the L2 capacity) pay?   usage: bash profiles/run_cta_sizes.sh > log
for rig in chain64 quad80; do
  for poses in 65536 75776; do
    for th in 512 448 384 128; do
      MBIK_THREADS=$th python profiles/run_kernel.py --rig $rig --poses $poses --launches 3 | tail -1 | sed "s/^/$rig poses=$poses threads=$th  /"
    done
  done
done
