#!/usr/bin/env python3
"""Time map_reads_kernel variants in ONE process on the GPU box: the human-scale workload is
built once, then every variant (a library file + environment settings) maps the same reads.

    python tools/sweep_map.py --pairs 8000000 --out gpurun_out/sweep.json \
        'base||' 'persist||SKM_L2_PERSIST=1' 'stage2|seekmer_b200/variants/lib_stage2.so|SKM_ROWS=28'

A spec is  label|library path (empty = the product library)|ENV=VAL,ENV=VAL.  Every variant's
exported class table is hashed; all hashes must agree (the variants differ in speed only)."""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from seekmer_b200 import _lib  # noqa: E402


def table_digest(t):
    h = hashlib.sha256()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        h.update(numpy.ascontiguousarray(t[k]).tobytes())
    h.update(str((t['unaligned'], t['aligned'])).encode())
    return h.hexdigest()[:16]


def use_library(path):
    _lib._lib = None
    _lib.LIB_PATH = _lib.pathlib.Path(path) if path else _lib.HERE / 'libseekmer_b200.so'
    return _lib.load()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('specs', nargs='+')
    ap.add_argument('--pairs', type=int, default=8_000_000)
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--passes', type=int, default=3)
    ap.add_argument('--out', default='gpurun_out/sweep.json')
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    use_library('')
    built, sim, lengths = bench.make_workload(a, 0, 1, dev)
    d_bases = torch.empty(a.pairs * 2 * bench.READ_LEN, dtype=torch.uint8, device=dev)
    bench.synth_reads(sim, 0, a.pairs, d_bases, 0)
    torch.cuda.synchronize()
    results = []
    digest0 = None
    for spec in a.specs:
        label, path, env = (spec.split('|') + ['', ''])[:3]
        envs = dict(kv.split('=', 1) for kv in env.split(',') if kv)
        saved = {k: os.environ.get(k) for k in envs}
        os.environ.update(envs)
        row = {'label': label, 'lib': path, 'env': envs}
        try:
            use_library(os.path.join(ROOT, path) if path else '')
            index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
            cap = int(os.environ.get('SWEEP_CLASS_CAP_LOG2', '23'))  # dictionary slots = 2 x class capacity
            mp = _lib.DeviceMapper(index, class_capacity=1 << cap, id_capacity=1 << min(27, cap + 4))
            times = []
            for p in range(a.passes):
                mp.reset()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                mp.map_batch(d_bases, None, a.pairs, True, fixed_len=bench.READ_LEN)
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3
                k = mp.kernel_ms()
                k['wall_ms'] = wall
                times.append(k)
            best = min(times, key=lambda k: k['map_reads_kernel'])
            # a diagnostics build (-DSKM_STATS=1): one more pass, counted
            st = numpy.zeros(32, dtype='u8')
            L = _lib.load()
            if L.skm_debug_map_stats(None, 1) == 0:
                mp.reset()
                mp.map_batch(d_bases, None, a.pairs, True, fixed_len=bench.READ_LEN)
                torch.cuda.synchronize()
                L.skm_debug_map_stats(st.ctypes.data, 1)
                names = ['load', 'scan', 'lookup', 'contig', 'walk', 'tally']
                row['stats'] = {n: {'iters': int(st[4 * i]), 'fill': round(float(st[4 * i + 1]) / max(1, int(st[4 * i])), 2),
                                    'cycles_per_iter': round(float(st[4 * i + 3]) / max(1, int(st[4 * i])), 1),
                                    'steps_per_read': round(float(st[4 * i + 1]) / (2 * a.pairs), 3)}
                                for i, n in enumerate(names)}
                row['stats']['idle_polls'] = int(st[24])
                row['stats']['iters_per_read'] = round(float(st[0:24:4].sum()) / (2 * a.pairs), 4)
                row['stats']['fill'] = round(float(st[1:24:4].sum()) / max(1.0, float(st[0:24:4].sum())), 2)
            dg = table_digest(mp.export())
            if digest0 is None:
                digest0 = dg
            row.update(map_ms=round(best['map_reads_kernel'], 3), pack_ms=round(best['pack_reads_kernel'], 3),
                       tally_ms=round(best['tally_units_kernel'], 3),
                       all_map_ms=[round(k['map_reads_kernel'], 3) for k in times],
                       mpairs_per_s_kernel=round(a.pairs / best['map_reads_kernel'] / 1e3, 1),
                       digest=dg, same_result=dg == digest0, sizes=mp.sizes())
            mp.close()
            index.close()
        except Exception as exc:  # a variant that fails must not cost the rest of the sweep
            row['error'] = repr(exc)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        print(json.dumps(row), flush=True)
        results.append(row)
    os.makedirs(os.path.dirname(os.path.join(ROOT, a.out)) or '.', exist_ok=True)
    json.dump({'pairs': a.pairs, 'results': results}, open(os.path.join(ROOT, a.out), 'w'), indent=1)


if __name__ == '__main__':
    main()
