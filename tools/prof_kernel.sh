#!/bin/bash
# usage: tools/prof_kernel.sh <tag> <kernel regex> [pairs] -- one ncu --set full capture of the 2nd launch of a kernel
tag=$1; rx=$2; pairs=${3:-8000000}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:$rx -s 1 -c 1 -f -o gpurun_out/prof_$tag \
    python tools/profile_map.py --pairs $pairs --passes 2 > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log
