#!/usr/bin/env python3
"""Parity at benchmark scale on an index the REFERENCE's own builder produced, against the
REFERENCE's own compiled mapper ("parity gaps by config").

    python tools/ref_index_parity.py --transcripts 200000 --pairs 2000000 --out gpurun_out/ref_index_parity.json

1. synthetic transcriptome of --transcripts isoforms (the benchmark's generator and seed);
2. `_index_builder.ContigAssembler().assemble` of the unmodified reference (oracle/_ref, compiled
   by oracle/build_ref.py) builds kmers / contigs / sequences / targets on the host;
3. --pairs simulated 2x150 pairs (1 % substitutions) are mapped by the CUDA library (device index
   made from those four arrays) and by the reference's `_mapper.ReadMapper` on all host cores;
4. the class dictionaries (ordered id tuple -> count), the unaligned count and the fragment length
   histogram must be equal; a slice is also mapped with ONE reference thread and compared in
   Counter insertion order (first-seen order).
Test infrastructure: the only place outside tests/ and bench.py that runs oracle/_ref, as the checker."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from seekmer_b200 import _lib, synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402


def table_dict(table):
    off, ids, cnt = table['key_offsets'].tolist(), table['key_ids'].tolist(), table['counts'].tolist()
    return {tuple(ids[off[i]:off[i + 1]]): cnt[i] for i in range(len(cnt))}


def feeder_batches(raw, first, n, read_len):
    bsz = 65536  # common.BUFFER_SIZE
    out = []
    for s in range(first, first + n, bsz):
        m = min(bsz, first + n - s)
        reads = [raw[(2 * s + i) * read_len:(2 * s + i + 1) * read_len] for i in range(2 * m)]
        out.append((m, [b''] * m, reads))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--pairs', type=int, default=2_000_000)
    ap.add_argument('--ordered-pairs', type=int, default=200_000)
    ap.add_argument('--read-len', type=int, default=bench.READ_LEN)
    ap.add_argument('--frag-mean', type=int, default=bench.FRAG_MEAN)
    ap.add_argument('--em', action='store_true', help='also: abundances of the GPU EM against the stock infer.quantify')
    ap.add_argument('--out', default='gpurun_out/ref_index_parity.json')
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    line = {'transcripts': a.transcripts, 'pairs': a.pairs, 'read_len': a.read_len}
    L = a.read_len
    t0 = time.time()
    tx = synth.make_transcriptome(a.transcripts, seed=bench.SEED_TX, mean_exons=11)
    line['cdna_mb'] = round(tx.codes.shape[0] / 1e6, 1)
    t1 = time.time()
    rk, rc, rs, rt = rh.ref_build_index(tx.sequences())
    rk, rc, rs, rt = (numpy.asarray(x) for x in (rk, rc, rs, rt))
    t2 = time.time()
    line.update(transcriptome_s=round(t1 - t0, 1), reference_builder_s=round(t2 - t1, 1),
                index={'table_slots': int(rk.shape[0]), 'n_contigs': int(rc.shape[0]), 'n_bases': int(rs.shape[0]),
                       'n_targets': int(rt.shape[0])})
    print('[ref-index] %.1f Mb cDNA, reference ContigAssembler: %.1f s, %d contigs' %
          (line['cdna_mb'], t2 - t1, rc.shape[0]), flush=True)

    # simulated reads, generated on the device by the benchmark's generator
    lengths = (tx.offsets[1:] - tx.offsets[:-1])
    expr = synth.make_expression(lengths.shape[0], seed=bench.SEED_EXPR)
    w = expr * numpy.maximum(lengths - a.frag_mean + 1, 1)
    w = w / w.sum()
    cum = numpy.cumsum(numpy.floor(w * float(1 << 40)).astype('u8')).astype('u8')
    sim = dict(codes=torch.from_numpy(tx.codes).to(dev), offsets=torch.from_numpy(tx.offsets).to(dev),
               cum=torch.from_numpy(cum.view('i8')).to(dev), total=int(cum[-1]), n_tx=lengths.shape[0])
    d_bases = torch.empty(a.pairs * 2 * L, dtype=torch.uint8, device=dev)
    bench.synth_reads(sim, 0, a.pairs, d_bases, 0, read_len=L, frag_mean=a.frag_mean)
    torch.cuda.synchronize()

    # CUDA path
    index = _lib.DeviceIndex(rk, rc, rs, rt, tx.n_transcripts)
    mp = _lib.DeviceMapper(index, class_capacity=1 << 22, id_capacity=1 << 26)
    mp.map_batch(d_bases, None, a.pairs, True, fixed_len=L)
    torch.cuda.synchronize()
    line['gpu_kernel_ms'] = {k: round(v, 3) for k, v in mp.kernel_ms().items()}
    table = mp.export()
    gpu = table_dict(table)
    gpu_em = None
    if a.em:
        from seekmer_b200 import infer
        from oracle import oracle as orc
        eff = orc.effective_lengths(numpy.asarray(table['fld']), lengths.astype('f8'))
        plan = _lib.EmPlan.from_mapper(mp, tx.n_transcripts)
        x0 = numpy.ones(eff.size) / eff
        x0 /= x0.sum()
        out, its = plan.run(eff, x0)
        gpu_em = (infer._finish(out[0]), int(its[0]), eff)
        plan.close()
    # ... and the first --ordered-pairs of them alone, for the first-seen order
    mp.reset()
    mp.map_batch(d_bases[:a.ordered_pairs * 2 * L], None, a.ordered_pairs, True, fixed_len=L)
    torch.cuda.synchronize()
    head = mp.export()
    mp.close()
    index.close()

    # the reference: compiled _mapper.ReadMapper on its own index object
    raw = d_bases.cpu().numpy().tobytes()
    ridx = rh.ref_index_from_arrays(rk, rc, rs, rt)
    cores = os.cpu_count() or 1
    t3 = time.time()
    res = rh.ref_map_threads(ridx, feeder_batches(raw, 0, a.pairs, L), cores)
    t4 = time.time()
    want = dict(res.counter)
    want_unaligned = want.pop((), 0)
    line['reference_mapper'] = {'cores': cores, 'seconds': round(t4 - t3, 2),
                                'pairs_per_s': round(a.pairs / (t4 - t3), 1)}
    one = rh.ref_map(ridx, feeder_batches(raw, 0, a.ordered_pairs, L))
    order = [k for k in one.counter.keys() if k]
    off, ids = head['key_offsets'].tolist(), head['key_ids'].tolist()
    gpu_order = [tuple(ids[off[i]:off[i + 1]]) for i in range(len(off) - 1)]

    fld = numpy.asarray(res.fragment_length_counts)
    gfld = numpy.asarray(table['fld'])
    n = min(fld.shape[0], gfld.shape[0])
    checks = {
        'class_dictionary_equal': gpu == want,
        'n_classes': [len(gpu), len(want)],
        'unaligned': [int(table['unaligned']), int(want_unaligned)],
        'aligned': [int(table['aligned']), int(sum(want.values()))],
        'fld_equal': bool((fld[:n] == gfld[:n]).all() and fld[n:].sum() == 0 and gfld[n:].sum() == 0),
        'first_seen_order_equal': gpu_order == order,
        'ordered_classes': [len(gpu_order), len(order)],
        'ordered_counts_equal': [int(c) for c in head['counts'].tolist()] == [one.counter[k] for k in order],
    }
    if gpu_em is not None and rh.have_python_reference():
        # the reference's own result objects and its own infer.quantify on ITS class table
        import importlib
        ref_mapper = importlib.import_module('seekmer.mapper')
        ref_infer = importlib.import_module('seekmer.infer')
        keys = [k for k in res.counter.keys() if k]
        class_map = numpy.asarray([[c, t] for c, k in enumerate(keys) for t in k], dtype='i8').T
        class_count = numpy.asarray([res.counter[k] for k in keys], dtype='f8')
        tpm_gpu, its_gpu, eff = gpu_em
        summ = ref_mapper.SummarizedResult(int(class_count.sum()), int(want_unaligned), a.pairs, class_map,
                                           class_count, None, eff)
        t5 = time.time()
        tpm_ref = ref_infer.quantify(summ)
        line['reference_quantify_s'] = round(time.time() - t5, 2)
        from oracle import oracle as orc
        # iteration count: the numpy restatement of infer.em on the SAME class order as the device's
        # (first-seen order; the 16-thread reference Counter has the same classes in another order)
        _, its_ref = orc.quantify(eff, orc.class_map_from_csr(table['key_offsets'], table['key_ids']),
                                  numpy.asarray(table['counts'], dtype='f8'), return_iters=True)
        checks['tpm_rel_1e-6'] = bool(numpy.allclose(tpm_gpu, tpm_ref, rtol=1e-6, atol=0))
        checks['em_iterations'] = [its_gpu, int(its_ref)]
    line['checks'] = checks
    line['ok'] = bool(checks['class_dictionary_equal'] and checks['fld_equal'] and checks['first_seen_order_equal']
                      and checks['ordered_counts_equal'] and checks['unaligned'][0] == checks['unaligned'][1]
                      and checks['aligned'][0] == checks['aligned'][1]
                      and checks.get('tpm_rel_1e-6', True)
                      and len(set(checks.get('em_iterations', [0]))) == 1)
    os.makedirs(os.path.dirname(os.path.join(ROOT, a.out)) or '.', exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, a.out), 'w'), indent=1)
    print(json.dumps(line), flush=True)
    return 0 if line['ok'] else 1


if __name__ == '__main__':
    sys.exit(main())
