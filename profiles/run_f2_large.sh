#!/bin/bash
# packed FP32x2 composites in the large-rig translation units, now that their big batches run the streamed-walk
# instantiation (no spills there): A/B against the stock library, bit-identity checked by run_variants.py
mkdir -p gpurun_out
out=gpurun_out/r2_exp_f2_large_rigs.log
: > $out
for spec in "chain64 65536" "chain64 75776" "quad80 65536" "chain150 75776" "chain200 75776" "big_tree240 75776" "big_tree120 75776"; do
  set -- $spec
  python profiles/run_variants.py --rig $1 --poses $2 --launches 4 "${@:3}" f2all f2mat f2nodiv 2>&1 | grep -v "^{" >> $out
done
cat $out
