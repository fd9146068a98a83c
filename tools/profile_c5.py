#!/usr/bin/env python3
"""BASELINE.json configs[4] on one GPU: single-end 75 bp reads at 2 % substitution error, 16 samples, each mapped
into its own class table and quantified (main EM).  With --check, sample 0 is compared unit by unit with the
C oracle."""
import argparse, os, sys, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seekmer_b200 import _lib, dist as sdist, infer, mapper

L75 = 75


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=16)
    ap.add_argument('--reads', type=int, default=4_000_000, help='reads per sample')
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--check', type=int, default=200_000, help='reads of sample 0 checked against the oracle')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    built, sim, lengths = bench.make_workload(a, 0, 1, dev)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    mp = _lib.DeviceMapper(index, class_capacity=1 << 22, id_capacity=1 << 26)
    lib = _lib.load()
    d = torch.empty(a.samples, a.reads * L75, dtype=torch.uint8, device=dev)
    for s in range(a.samples):
        _lib.check(lib.skm_synth_reads(
            _lib._ptr(sim['codes']), _lib._ptr(sim['offsets']), sim['n_tx'], _lib._ptr(sim['cum']), sim['total'],
            L75, 200, 30, int(round(0.02 * 65536)), int(round(0.001 * 65536)), 1, 100 + s, 0,
            0, a.reads, d[s].data_ptr(), 0, _lib.current_stream_ptr()))
    torch.cuda.synchronize()

    class FakeIndex:
        transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
    FakeIndex.transcripts['length'] = lengths

    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tables = []
        for s in range(a.samples):
            mp.reset()
            mp.map_batch(d[s], None, a.reads, False, fixed_len=L75)
            tables.append(sdist.table_to_host(mp.export_torch()))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tpms = []
        for tab in tables:
            mr = mapper.MapResult(FakeIndex)
            mr.fragment_length_counts = tab['fld'].astype('i8')
            tpms.append(infer.quantify(mapper.summarize_table(tab, mr)))
        t2 = time.perf_counter()
        n = a.samples * a.reads
        print('pass %d: %d samples x %d SE75 reads: mapping %.1f ms = %.1f M reads/s (kernels %s); %d EMs %.1f ms; '
              'aligned %.1f %%' % (rep, a.samples, a.reads, (t1 - t0) * 1e3, n / (t1 - t0) / 1e6,
                                   {k: round(v, 2) for k, v in mp.kernel_ms().items()}, a.samples, (t2 - t1) * 1e3,
                                   100.0 * sum(t['aligned'] for t in tables) / n), flush=True)
    if a.check:
        from oracle import oracle as orc
        oidx = orc.OracleIndex(*built.numpy_arrays())
        k = min(a.check, a.reads)
        hb = d[0, :k * L75].cpu().numpy()
        offs = numpy.arange(k + 1, dtype='i8') * L75
        aligned, h, cnt, length, fld = orc.map_batch_mt(oidx, hb, offs, False, os.cpu_count() or 1)
        mp.reset()
        g_cls, g_len = mp.map_batch(d[0, :k * L75], None, k, False, fixed_len=L75, per_read=True)
        chk = mp.export()
        g_cls, g_len = g_cls.cpu().numpy().astype('i8'), g_len.cpu().numpy()
        o_key = numpy.where(cnt > 0, h, numpy.uint64(0)).astype('u8')
        g_key = numpy.where(g_cls >= 0, g_cls + 1, 0).astype('u8')
        pairs = numpy.unique(numpy.stack([g_key, o_key]), axis=1)
        ok = (numpy.unique(pairs[0]).size == pairs.shape[1] and numpy.unique(pairs[1]).size == pairs.shape[1])
        print('oracle check on %d reads: fld %s, aligned %s, lengths %s, classes %s' %
              (k, bool((chk['fld'] == fld).all()), chk['aligned'] == aligned, bool((g_len == length).all()), bool(ok)))


if __name__ == '__main__':
    main()
