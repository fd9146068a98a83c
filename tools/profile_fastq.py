#!/usr/bin/env python3
"""End-to-end rate from FASTQ files: raw-text GPU path vs the line-by-line feeder protocol."""
import argparse, os, sys, tempfile, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seekmer_b200 import common, mapper


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=4_000_000)
    ap.add_argument('--slow-pairs', type=int, default=200_000)
    ap.add_argument('--transcripts', type=int, default=200_000)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    built, sim, lengths = bench.make_workload(a, 0, 1, dev)
    arrays = built.numpy_arrays()
    tr = numpy.zeros(lengths.shape[0], dtype=[('transcript_id', 'S12'), ('gene_id', 'S12'), ('length', 'f8')])
    tr['length'] = lengths
    index = common.KMerIndex(*arrays, tr, None)
    d = torch.empty(a.pairs * 2 * bench.READ_LEN, dtype=torch.uint8, device=dev)
    bench.synth_reads(sim, 0, a.pairs, d, 0)
    reads = d.cpu().numpy().reshape(a.pairs, 2, bench.READ_LEN)
    tmp = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    paths = [os.path.join(tmp, 'r_%d.fastq' % m) for m in (1, 2)]
    t = time.perf_counter()
    for m, p in enumerate(paths):
        L = bench.READ_LEN
        rec = numpy.empty((a.pairs, 2 * L + 24), dtype='u1')
        hdr = numpy.frombuffer(b'@SIM.%012d/1\n', dtype='u1')
        rec[:, :6] = numpy.frombuffer(b'@SIM.0', dtype='u1')
        idx = numpy.arange(a.pairs)
        for k in range(11):
            rec[:, 16 - k] = 48 + (idx // 10 ** k) % 10
        rec[:, 5] = ord('.')
        rec[:, 17] = ord('/'); rec[:, 18] = ord('1') + m; rec[:, 19] = 10
        rec[:, 20:20 + L] = reads[:, m, :]
        rec[:, 20 + L] = 10; rec[:, 21 + L] = ord('+'); rec[:, 22 + L] = 10
        rec[:, 23 + L:23 + 2 * L] = ord('I'); rec[:, 23 + 2 * L] = 10
        rec.tofile(p)
    print('wrote 2 x %.2f GB FASTQ in %.1f s' % (os.path.getsize(paths[0]) / 1e9, time.perf_counter() - t), flush=True)
    for rep in range(2):
        t = time.perf_counter()
        res = mapper.map_reads(index, common.feed_pair_ended_reads(*paths))
        dt = time.perf_counter() - t
        n = sum(res.counter.values())
        print('raw-text GPU path: %d pairs in %.2f s = %.2f M pairs/s (%.2f GB/s of FASTQ)' %
              (n, dt, n / dt / 1e6, 2 * os.path.getsize(paths[0]) / dt / 1e9), flush=True)
    # the same, stage by stage
    from seekmer_b200 import _lib
    dm = _lib.DeviceMapper(index.device_index(0), 1 << 23, 1 << 27)
    for rep in range(2):
        dm.reset()
        src = common.feed_pair_ended_reads(*paths)
        t = time.perf_counter()
        first = 0
        t_read = t_map = 0.0
        t0 = time.perf_counter()
        for b1, n1, b2, n2, eof in src.text_chunks(mapper.FASTQ_CHUNK_BYTES):
            t1 = time.perf_counter()
            units, c1, c2 = dm.map_fastq(b1, n1, b2, n2, first_unit=first)
            t2 = time.perf_counter()
            src.consumed(c1, c2)
            first += units
            t_read += t1 - t0
            t_map += t2 - t1
            t0 = time.perf_counter()
        t3 = time.perf_counter()
        table = dm.export()
        t4 = time.perf_counter()
        cl = mapper._class_tuples(table)
        t5 = time.perf_counter()
        print('stages: read+carry %.2f s, skm_map_fastq %.2f s, export %.2f s, python tuples %.2f s; mapping loop alone %.2f M pairs/s'
              % (t_read, t_map, t4 - t3, t5 - t4, first / (t3 - t) / 1e6), flush=True)
    dm.close()
    # the feeder protocol (what the reference's mapper consumes), on a prefix
    small = [os.path.join(tmp, 's_%d.fastq' % m) for m in (1, 2)]
    for p, s in zip(paths, small):
        with open(p, 'rb') as f, open(s, 'wb') as g:
            g.write(f.read((2 * bench.READ_LEN + 24) * a.slow_pairs))
    t = time.perf_counter()
    res2 = mapper.map_reads(index, iter(common.feed_pair_ended_reads(*small)))
    dt = time.perf_counter() - t
    print('line-by-line feeder protocol: %d pairs in %.2f s = %.2f M pairs/s' % (a.slow_pairs, dt, a.slow_pairs / dt / 1e6))
    for p in paths + small:
        os.remove(p)
    os.rmdir(tmp)


if __name__ == '__main__':
    main()
