#!/bin/bash
# usage: tools/launches.sh <tag> [pairs] -- per-kernel durations of one profile_map run (ncu, times are cold/serialised)
tag=$1; pairs=${2:-8000000}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pack_reads|map_reads|tally_units' -c 9 --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_map.py --pairs $pairs --passes 3 > /dev/null 2>&1
python3 - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$tag.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-3:]: print(r[4][:40], float(r[-1])/1e6,'ms')
PY
