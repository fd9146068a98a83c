#!/usr/bin/env python3
"""Aggregate an ncu SASS-level source page by CUDA source line, using nvdisasm -g line info.

usage: ncu_by_line.py <sass_csv from `ncu --page source --csv`> <nvdisasm -g -c output> <kernel substring>
"""
import csv
import re
import sys
from collections import defaultdict

csv_path, dis_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
insts = rows[2:]
# line info per instruction, in order
lines = []
cur = None
active = False
for ln in open(dis_path):
    if ln.startswith('.text.'):
        active = kname in ln
        cur = None
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
if len(lines) != len(insts):
    print('warning: %d sass lines vs %d profiled instructions' % (len(lines), len(insts)))
agg = defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0]
for (loc, r) in zip(lines, insts):
    w = int(float(r[col['Instructions Executed']] or 0))
    t = int(float(r[col['Thread Instructions Executed']] or 0))
    s = int(float(r[col['# Samples']] or 0))
    lsb = int(float(r[col['stall_long_sb']] or 0))
    noi = int(float(r[col['stall_no_inst']] or 0))
    a = agg[loc]
    a[0] += w; a[1] += t; a[2] += s; a[3] += lsb; a[4] += noi
    tot[0] += w; tot[1] += t; tot[2] += s
print('total warp-inst %.3e thread-inst %.3e avg threads %.2f samples %d' % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
src = {}
print('%-18s %8s %7s %7s %7s %7s  %s' % ('location', 'warp%', 'avgthr', 'samp%', 'longsb%', 'noinst%', 'source'))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ''
    if loc:
        try:
            if loc[0] not in src:
                import glob
                cand = glob.glob('/root/repo/seekmer_b200/csrc/' + loc[0])
                src[loc[0]] = open(cand[0]).read().split('\n') if cand else []
            text = src[loc[0]][loc[1] - 1].strip()[:90]
        except Exception:
            pass
    print('%-18s %7.2f%% %7.2f %6.2f%% %6.2f%% %6.2f%%  %s' % (
        '%s:%d' % loc if loc else '?', 100.0 * a[0] / tot[0], a[1] / max(a[0], 1), 100.0 * a[2] / max(tot[2], 1),
        100.0 * a[3] / max(tot[2], 1), 100.0 * a[4] / max(tot[2], 1), text))
