#!/usr/bin/env python3
"""Stage timings of the single-cell `impute` workflow at a non-toy size.

    python tools/profile_impute.py --transcripts 20000 --cells 64 --pairs 50000

Synthetic transcriptome + index built on the device (as bench.py does), CELLS cells drawn from
four expression programmes, reads synthesised on the device.  Reports, per stage, the time of
`impute.impute_cells` pieces, the batched second round against (a) the same cells quantified
one `skm_em` call at a time (the reference's loop structure, `impute.py:110-115`, on the GPU)
and (b) the CPU oracle on a bounded sample of cells, with which the results are compared
(rel 1e-6).  One JSON line on stdout.  Test infrastructure: the oracle is the checker here.
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

READ_LEN, FRAG_MEAN, FRAG_SD = 100, 250, 30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--transcripts', type=int, default=20000)
    ap.add_argument('--cells', type=int, default=64)
    ap.add_argument('--pairs', type=int, default=50000)
    ap.add_argument('--power', type=int, default=16)
    ap.add_argument('--oracle-cells', type=int, default=2)
    ap.add_argument('--skip-baselines', action='store_true',
                    help='only the stages impute_cells runs (no dense / per-cell / oracle comparisons)')
    ap.add_argument('--fastq', action='store_true',
                    help='also write every cell as FASTQ files and time `impute.run` end to end')
    args = ap.parse_args()

    import torch
    from seekmer_b200 import _lib, common, impute, index_build, infer, mapper, synth
    from oracle import oracle as orc

    device = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    tx = synth.make_transcriptome(args.transcripts, seed=1, mean_exons=11)
    codes = torch.from_numpy(tx.codes).to(device)
    offsets = torch.from_numpy(tx.offsets).to(device)
    built = index_build.build_index(codes, offsets, device=device)
    lengths = (offsets[1:] - offsets[:-1]).cpu().numpy()
    n_tx = lengths.shape[0]
    genes = [b'G%06d' % (i // 3) for i in range(n_tx)]
    index = common.KMerIndex(None, None, None, None, built.transcripts_table(gene_ids=genes), None)
    dev_index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, n_tx)
    index._device[0] = dev_index

    programmes = [synth.make_expression(n_tx, seed=3 + p) for p in range(4)]
    fastq_dir = None
    fastq_paths = []
    if args.fastq:
        import tempfile
        fastq_dir = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)

    def write_fastq(cell, reads):
        # fixed-width records written as one byte matrix: '@c<cell>.<pair>/<mate>', bases, '+', qualities
        n, length = reads.shape[0], READ_LEN
        idx = numpy.arange(n)
        for mate in (0, 1):
            rec = numpy.empty((n, 2 * length + 16), dtype='u1')
            rec[:, 0] = ord('@')
            for k in range(8):
                rec[:, 8 - k] = 48 + (idx // 10 ** k) % 10
            rec[:, 9] = ord('/')
            rec[:, 10] = ord('1') + mate
            rec[:, 11] = 10
            rec[:, 12:12 + length] = reads[:, mate, :]
            rec[:, 12 + length] = 10
            rec[:, 13 + length] = ord('+')
            rec[:, 14 + length] = 10
            rec[:, 15 + length:15 + 2 * length] = ord('I')
            rec[:, 15 + 2 * length] = 10
            path = os.path.join(fastq_dir, 'cell%04d_%d.fastq' % (cell, mate + 1))
            rec.tofile(path)
            fastq_paths.append(path)
    d_bases = torch.empty(args.pairs * 2 * READ_LEN, dtype=torch.uint8, device=device)
    L = _lib.load()

    def map_cell(cell):
        rng = numpy.random.Generator(numpy.random.PCG64(100 + cell))
        expr = programmes[cell % 4] * rng.uniform(0.8, 1.25, size=n_tx)
        w = expr * numpy.maximum(lengths - FRAG_MEAN + 1, 1)
        w = w / w.sum()
        cum = numpy.cumsum(numpy.floor(w * float(1 << 40)).astype('u8')).astype('u8')
        d_cum = torch.from_numpy(cum.view('i8')).to(device)
        _lib.check(L.skm_synth_reads(
            _lib._ptr(codes), _lib._ptr(offsets), n_tx, _lib._ptr(d_cum), int(cum[-1]), READ_LEN, FRAG_MEAN,
            FRAG_SD, int(round(0.01 * 65536)), int(round(0.001 * 65536)), 1, 1000 + cell, 1, 0, args.pairs,
            d_bases.data_ptr(), 0, _lib.current_stream_ptr()))
        if fastq_dir is not None:
            write_fastq(cell, d_bases.cpu().numpy().reshape(args.pairs, 2, READ_LEN))
        mp.reset()   # one device mapper for all cells, as mapper.map_multiple_samples does
        mp.map_batch(d_bases, None, args.pairs, True, first_unit=0, fixed_len=READ_LEN)
        table = mp.export()
        result = mapper.MapResult(index)
        result.update_counts(mapper._class_tuples(table))
        if table['unaligned']:
            result.counter[()] += table['unaligned']
        result._table = table
        result.merge_fragment_lengths(table['fld'])
        return result

    t = {}
    t0 = time.time()
    mp = _lib.DeviceMapper(dev_index, 0, 0)
    results = [map_cell(c) for c in range(args.cells)]
    mp.close()
    torch.cuda.synchronize()
    t['map_cells_s'] = time.time() - t0   # read synthesis + mapping + export + host tuples

    t0 = time.time()
    impute._merge_fragment_lengths(results)
    summarized = [r.summarize() for r in results]
    t['summarize_s'] = time.time() - t0
    infer.quantify_samples(summarized[:2])  # warm the EM scratch cache
    t0 = time.time()
    base = infer.quantify_samples(summarized)
    t['first_round_s'] = time.time() - t0   # what impute_cells runs: all cells in one skm_em_samples call
    t0 = time.time()
    base_loop = numpy.asarray([infer.quantify(r) for r in summarized])
    t['first_round_one_call_per_cell_s'] = time.time() - t0
    first_round_identical = bool((base == base_loop).all())
    t0 = time.time()
    weight = impute._calculate_cell_weights(index, base, None)
    t['weights_s'] = time.time() - t0
    powered = weight ** args.power
    impute._quantify_weighted(summarized[:2], powered[:2, :2] + numpy.eye(2))  # warm the EM scratch cache
    t0 = time.time()
    grouped = impute._quantify_weighted(summarized, powered)
    t['second_round_grouped_s'] = time.time() - t0   # what impute_cells runs
    if args.skip_baselines:
        print(json.dumps({'workload': '%d cells x %d pairs, %d transcripts' % (args.cells, args.pairs, n_tx),
                          'stages': {k_: round(v, 4) for k_, v in t.items()},
                          'first_round_samples_call_bit_identical_to_loop': first_round_identical}))
        return
    t0 = time.time()
    impute._blend_mapping_results(summarized, powered)
    t['blend_dense_s'] = time.time() - t0

    t0 = time.time()
    tpm = impute._quantify_blended(summarized)
    t['second_round_dense_batched_s'] = time.time() - t0
    grouped_close = bool(numpy.allclose(grouped, tpm, rtol=1e-6, atol=0))
    t0 = time.time()
    serial = numpy.asarray([infer.quantify(r) for r in summarized])
    t['second_round_one_call_per_cell_s'] = time.time() - t0
    serial_close = bool(numpy.allclose(serial, tpm, rtol=1e-6, atol=0))

    k = min(args.oracle_cells, args.cells)
    t0 = time.time()
    want = [orc.quantify(summarized[i].effective_lengths, summarized[i].class_map, summarized[i].class_count)
            for i in range(k)]
    t['oracle_s_per_cell'] = (time.time() - t0) / max(k, 1)
    oracle_close = all(bool(numpy.allclose(tpm[i], want[i], rtol=1e-6, atol=0)) for i in range(k))

    end_to_end = None
    if fastq_dir is not None:
        import pathlib
        import shutil
        host_index = common.KMerIndex(*built.numpy_arrays(), index.transcripts, None)
        index_path = pathlib.Path(fastq_dir) / 'index.npz'
        host_index.save(index_path)
        out_dir = pathlib.Path(fastq_dir) / 'out'
        t0 = time.time()
        impute.run(index_path, out_dir, [pathlib.Path(p) for p in fastq_paths], job_count=1,
                   single_ended=False, debug=False, power=args.power)
        seconds = time.time() - t0
        rows = [l.rstrip('\n').split(',') for l in open(str(out_dir / 'tpm.csv'))]
        table = numpy.asarray([[float(v) for v in r[1:]] for r in rows[1:]]).T
        end_to_end = {'impute_run_s': round(seconds, 3), 'cells_per_s': round(args.cells / seconds, 2),
                      'fastq_gb': round(sum(os.path.getsize(p) for p in fastq_paths) / 1e9, 3),
                      'tpm_csv_equals_in_memory_1e-9': bool(numpy.allclose(table, grouped, rtol=1e-9, atol=0))}
        shutil.rmtree(fastq_dir, ignore_errors=True)

    shared = summarized[0].class_map
    line = {
        'workload': '%d cells x %d 2x%d pairs, %d transcripts, power %d' %
                    (args.cells, args.pairs, READ_LEN, n_tx, args.power),
        'classes_per_cell_mean': float(numpy.mean([r._table['counts'].shape[0] for r in results])),
        'blended_classes': int(summarized[0].class_count.size), 'blended_nnz': int(shared.shape[1]),
        'weights_kept_per_cell_mean': float((weight != 0).sum(axis=1).mean()),
        'stages': {k_: round(v, 4) for k_, v in t.items()},
        'from_fastq_files': end_to_end,
        'support_groups': len(impute._support_groups(powered)),
        'second_round_speedup_vs_per_cell_calls':
            round(t['second_round_one_call_per_cell_s'] / t['second_round_grouped_s'], 2),
        'second_round_speedup_vs_oracle_cpu':
            round(t['oracle_s_per_cell'] * args.cells / t['second_round_grouped_s'], 1),
        'parity': {'first_round_samples_call_bit_identical_to_loop': first_round_identical,
                   'grouped_equals_dense_batched_1e-6': grouped_close,
                   'batched_equals_per_cell_calls_1e-6': serial_close,
                   'batched_equals_oracle_1e-6_on_%d_cells' % k: oracle_close},
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
