#!/bin/bash
# chain64 / quad80 / big rigs: ms per 75776-pose launch of the thread-per-pose kernel (for A/B of a rebuilt library)
for rig in chain64 quad80; do python profiles/run_kernel.py --rig $rig --poses 75776 --launches 3 | tail -1 | sed "s/^/$rig /"; done
