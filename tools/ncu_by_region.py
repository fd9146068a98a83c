#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by code region of map_kernel.cuh.
usage: ncu_by_region.py <sass csv> <nvdisasm -g -c output> <kernel substring> [reads]"""
import csv, re, sys
from collections import defaultdict
csv_path, dis_path, kname = sys.argv[1:4]
reads = float(sys.argv[4]) if len(sys.argv) > 4 else 1
rows = list(csv.reader(open(csv_path))); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}; insts = rows[2:]
lines = []; cur = None; active = False
for ln in open(dis_path):
    if ln.startswith('.text.'):
        active = kname in ln; cur = None; continue
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): lines.append(cur)
src = open('/root/repo/seekmer_b200/csrc/map_kernel.cuh').read().split('\n')
MARKS = [('struct ReadView', 'ReadView'), ('struct List', 'List get/set'), ('void lane_load', 'lane_load'),
         ('void lane_store', 'lane_store'), ('uint32_t contig_window', 'windows'), ('int32_t list_get', 'list slow paths'),
         ('void map_contig(', 'map_contig'), ('int filter_long', 'list slow paths'), ('bool filter_on_contig', 'filter'),
         ('bool intersect', 'intersect'), ('uint32_t hash_bucket', 'hash/probe wrappers'), ('int sift4_edge', 'sift4_edge'),
         ('map_reads_kernel(const DevIndex', 'prologue'), ('// ---- vote', 'vote+claim'), ('Lane<ITEMS> L;', 'step setup'),
         ('if (phase == P_LOAD)', 'P_LOAD'), ('} else if (phase == P_LOOKUP)', 'P_LOOKUP'), ('} else if (phase == P_SCAN)', 'P_SCAN'),
         ('} else if (phase == P_MAP', 'P_MAP/FILTER'), ('} else if (phase == P_WALK)', 'P_WALK'), ('} else {  // P_TALLY', 'P_TALLY'),
         ('// ---- transitions', 'transitions'), ('// ---- publish', 'publish')]
marks = [(1, 'pre')]
for pat, name in MARKS:
    for i, l in enumerate(src):
        if pat in l:
            marks.append((i + 1, name)); break
marks.sort()
def cat(loc):
    if not loc: return '?'
    f, l = loc
    if f != 'map_kernel.cuh': return f
    name = 'pre'
    for a, n in marks:
        if l >= a: name = n
    return name
agg = defaultdict(lambda: [0, 0, 0, 0]); tot = [0, 0, 0]
for loc, r in zip(lines, insts):
    w = int(float(r[col['Instructions Executed']] or 0)); t = int(float(r[col['Thread Instructions Executed']] or 0)); s = int(float(r[col['# Samples']] or 0))
    a = agg[cat(loc)]; a[0] += w; a[1] += t; a[2] += s; a[3] += 1; tot[0] += w; tot[1] += t; tot[2] += s
print('static %d instr; warp-inst/read %.1f  thread-inst/read %.0f  avg threads %.1f' % (len(lines), tot[0] / reads, tot[1] / reads, tot[1] / max(tot[0], 1)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-28s static %5d  warp%% %5.1f  winst/read %6.1f avgthr %5.1f  samp%% %5.1f' % (k, a[3], 100 * a[0] / tot[0], a[0] / reads, a[1] / max(a[0], 1), 100 * a[2] / tot[2]))
