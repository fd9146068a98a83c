#!/usr/bin/env python3
"""Main EM on the benchmark's own class structure (human-scale index, N pairs mapped on the GPU,
plan made from the mapper device to device): wall and device time per run; the target of
`ncu -k regex:em_loop_kernel`."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from seekmer_b200 import _lib, mapper  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=30_000_000)
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--runs', type=int, default=3)
    ap.add_argument('--bootstraps', type=int, default=0)
    a = ap.parse_args()
    a.config = 'c2'
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    built, sim, lengths = bench.make_workload(a, 0, 1, dev)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    mp = _lib.DeviceMapper(index, class_capacity=1 << 22, id_capacity=1 << 26)
    d_bases = torch.empty(a.pairs * 2 * bench.READ_LEN, dtype=torch.uint8, device=dev)
    bench.synth_reads(sim, 0, a.pairs, d_bases, 0)
    mp.map_batch(d_bases, None, a.pairs, True, fixed_len=bench.READ_LEN)
    table = mp.export()
    t0 = time.perf_counter()
    plan = _lib.EmPlan.from_mapper(mp, lengths.shape[0])
    print('plan from mapper: %.2f ms (C=%d nnz=%d)' % ((time.perf_counter() - t0) * 1e3, plan.n_classes, plan.nnz))

    class FakeIndex:
        transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
    FakeIndex.transcripts['length'] = lengths
    mr = mapper.MapResult(FakeIndex)
    mr.fragment_length_counts = table['fld'].astype('i8')
    eff = mr.effective_lengths
    x = numpy.ones(lengths.shape[0]) / eff
    x /= x.sum()
    deg = numpy.bincount(table['key_ids'], minlength=lengths.shape[0])
    print('transcript degree: mean %.1f max %d, rows > 256: %d' % (deg.mean(), deg.max(), int((deg > 256).sum())))
    d_eff, d_x = torch.from_numpy(eff).to(dev), torch.from_numpy(x).to(dev)
    d_out = torch.zeros_like(d_x)
    d_it = torch.zeros(1, dtype=torch.int32, device=dev)
    for r in range(a.runs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        _lib.check(_lib.load().skm_em_plan_run(plan._h, None, _lib._ptr(d_eff), _lib._ptr(d_x), 1, 0, _lib._ptr(d_out),
                                               _lib._ptr(d_it), 1, _lib.current_stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        it = int(d_it.item())
        print('main EM run %d: device %.3f ms, wall %.3f ms, %d iterations, %.1f us / iteration'
              % (r, e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, it, e0.elapsed_time(e1) * 1e3 / max(it, 1)))
    if a.bootstraps:
        main = d_out.cpu().numpy()
        for r in range(2):
            t0 = time.perf_counter()
            out, its = plan.bootstrap(eff, main / main.sum(), a.bootstraps, 1234)
            print('%d bootstraps: %.1f ms (mean %.1f iterations)' % (a.bootstraps, (time.perf_counter() - t0) * 1e3, its.mean()))


if __name__ == '__main__':
    main()
