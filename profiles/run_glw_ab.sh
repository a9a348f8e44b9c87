#!/bin/bash
# streamed-walk instantiation, A/B of library builds on the chain rigs (ms per 75776-pose launch, best of launches 2..)
for lib in stock "$@"; do
  if [ "$lib" = stock ]; then unset MBIK_LIB; else export MBIK_LIB=many_bone_ik_b200/_variants/libmbik_$lib.so; fi
  for p in 65536 75776; do python profiles/run_kernel.py --rig chain64 --poses $p --launches 5 | sort -t: -k2 -n | awk -v l=$lib -v p=$p 'NR>0{print l, "chain64", p, $0}' | sort -k6 -n | head -1; done
  for r in chain150 chain200; do python profiles/run_large_rig.py --rig $r --launches 5 | awk -v l=$lib '{print l, $0}' | sort -k5 -n | head -1; done
done
