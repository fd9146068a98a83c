#!/bin/bash
# usage: tools/prof_map.sh <tag> [pairs]   -- plain timing, then one ncu --set full capture of map_reads_kernel
tag=$1; pairs=${2:-8000000}
mkdir -p gpurun_out
python tools/profile_map.py --pairs $pairs --passes 3 > gpurun_out/plain_$tag.log 2>&1 || { tail -5 gpurun_out/plain_$tag.log; exit 1; }
grep pass gpurun_out/plain_$tag.log
ncu --set full --import-source on --clock-control none -k regex:map_reads_kernel -s 1 -c 1 -f -o gpurun_out/prof_map_$tag \
    python tools/profile_map.py --pairs $pairs --passes 2 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
