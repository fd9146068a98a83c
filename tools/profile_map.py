#!/usr/bin/env python3
"""Minimal driver for profiling the mapping kernel: human-scale index + N pairs, a few passes."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from seekmer_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=4_000_000)
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--passes', type=int, default=3)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    built, sim, lengths = bench.make_workload(a, 0, 1, dev)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    mp = _lib.DeviceMapper(index, class_capacity=1 << 23, id_capacity=1 << 27)
    d_bases = torch.empty(a.pairs * 2 * bench.READ_LEN, dtype=torch.uint8, device=dev)
    bench.synth_reads(sim, 0, a.pairs, d_bases, 0)
    torch.cuda.synchronize()
    for p in range(a.passes):
        mp.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mp.map_batch(d_bases, None, a.pairs, True, fixed_len=bench.READ_LEN)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('pass %d: %.3f ms, %.1f M pairs/s' % (p, ms, a.pairs / ms / 1e3), flush=True)
    print(mp.sizes())


if __name__ == '__main__':
    main()
