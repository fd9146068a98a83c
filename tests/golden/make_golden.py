#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden.py

Everything stored is an output of unmodified reference code:
  * index arrays           seekmer._index_builder.ContigAssembler.assemble   (compiled .pyx)
  * per-unit id tuples/FLD  seekmer._mapper.ReadMapper                        (compiled .pyx)
  * Counter / summarize     seekmer.mapper.MapResult                          (reference .py)
  * eff. lengths, em, tpm   seekmer.mapper / seekmer.infer                    (reference .py)
Inputs are either the reference's own test data (chr21 subset + its 21 read pairs) or
seeded synthetic data that the tests regenerate from the seed (seekmer_b200/synth.py).
"""
import bz2
import pathlib
import sys
import warnings

import numpy

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_harness as rh  # noqa: E402
from seekmer_b200 import synth  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent
REF_DATA = pathlib.Path('/root/reference/seekmer/test/data')


def csr(tuples):
    ptr = numpy.zeros(len(tuples) + 1, dtype='i8')
    numpy.cumsum([len(t) for t in tuples], out=ptr[1:])
    ids = numpy.asarray([x for t in tuples for x in t], dtype='i4')
    return ptr, ids


def transcripts_table(ids, seqs):
    n = max(len(i) for i in ids)
    tab = numpy.zeros(len(ids), dtype=[('transcript_id', 'S%d' % n), ('gene_id', 'S%d' % n),
                                       ('length', 'f8')])
    tab['transcript_id'] = ids
    tab['gene_id'] = ids
    tab['length'] = [len(s) for s in seqs]
    return tab


def run_reference(pkg, arrays, transcripts, batches):
    """reference mapper (1 thread) + reference MapResult/summarize/quantify."""
    from seekmer import mapper as ref_mapper, infer as ref_infer
    kmers, contigs, sequences, targets = arrays
    index = pkg._common.KMerIndex(kmers, contigs, sequences, targets, transcripts, None)
    # per-read tuples through the harness collector
    per = rh.ref_map(index, batches, keep_per_read=True)
    # the reference's own result model
    mr = ref_mapper.map_reads(index, iter(batches), job_count=1, debug=True)
    assert dict(mr.counter) == dict(per.counter)
    assert (mr.fragment_length_counts == per.fragment_length_counts).all()
    summ = mr.summarize()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        # em with iteration count: replay the reference loop around its own em() pieces
        tl = summ.effective_lengths.astype('f8')
        x = numpy.ones(tl.size) / tl
        x /= x.sum()
        x_em = ref_infer.em(x.copy(), tl, summ.class_map, summ.class_count)
        tpm = ref_infer.quantify(summ)
    return per, mr, summ, x_em, tpm


def em_iterations(x, l, class_map, class_count):
    """Count iterations of infer.em by replaying its update (restated; only the count is used,
    and it is cross-checked against the reference's returned x)."""
    from oracle import oracle as orc
    xo, it = orc.em(x.copy(), l, class_map, class_count, return_iters=True)
    return xo, it


def golden_chr21(pkg):
    ids, seqs = [], []
    name, cur = None, []
    for line in bz2.open(str(REF_DATA / 'human.cdna.21.fa.bz2'), 'rb'):
        if line[:1] == b'>':
            if name is not None:
                ids.append(name)
                seqs.append(b''.join(cur))
            name, cur = line[1:].strip().split()[0].split(b'.')[0], []
        else:
            cur.append(line.strip())
    ids.append(name)
    seqs.append(b''.join(cur))
    # transcripts touched by the 21 fixture pairs on the full index and neighbours
    keep = sorted(set(range(860, 872)) | set(range(930, 940)) | set(range(1199, 1204))
                  | set(range(86, 91)) | set(range(265, 275)))
    ids = [ids[i] for i in keep]
    seqs = [seqs[i] for i in keep]

    def fastq(p):
        lines = open(str(p), 'rb').read().split(b'\n')
        return ([lines[i].strip()[1:] for i in range(0, len(lines) - 1, 4)],
                [lines[i + 1].strip() for i in range(0, len(lines) - 1, 4)])
    n1, r1 = fastq(REF_DATA / '20_1.fastq')
    _, r2 = fastq(REF_DATA / '20_2.fastq')
    reads = []
    for a, b in zip(r1, r2):
        reads += [a, b]
    arrays = rh.ref_build_index(seqs)
    tab = transcripts_table(ids, seqs)
    per, mr, summ, x_em, tpm = run_reference(pkg, arrays, tab, [(len(r1), n1, reads)])
    _, iters = em_iterations(numpy.ones(len(seqs)) / summ.effective_lengths
                             / (1.0 / summ.effective_lengths).sum(),
                             summ.effective_lengths.astype('f8'), summ.class_map, summ.class_count)
    ptr, tids = csr(per.per_read)
    numpy.savez_compressed(
        str(OUT / 'chr21_subset.npz'),
        kmers=arrays[0], contigs=numpy.asarray(arrays[1]), sequences=arrays[2], targets=arrays[3],
        transcripts=tab, reads=numpy.asarray(reads), unit_ptr=ptr, unit_ids=tids,
        fld=per.fragment_length_counts, class_map=summ.class_map, class_count=summ.class_count,
        eff_lengths=summ.effective_lengths, em_x=x_em, em_iters=numpy.asarray(iters), tpm=tpm,
        aligned=numpy.asarray(summ.aligned), unaligned=numpy.asarray(summ.unaligned),
        harmonic_mean=numpy.asarray(float(mr.harmonic_mean_fragment_length)))
    print('chr21_subset: %d transcripts, %d slots, %d contigs, classes %d, unaligned %d, em iters %d'
          % (len(seqs), arrays[0].shape[0], arrays[1].shape[0], summ.class_count.size,
             summ.unaligned, iters))


ADVERSARIAL_NOTE = 'see tests/adversarial.py'


def golden_synth(pkg):
    sys.path.insert(0, str(ROOT / 'tests'))
    import adversarial
    tx = synth.make_transcriptome(60, seed=7)
    seqs = tx.sequences()
    arrays = rh.ref_build_index(seqs)
    tab = transcripts_table(tx.ids(), seqs)
    expr = synth.make_expression(tx.n_transcripts, seed=3)
    out = dict(kmers=arrays[0], contigs=numpy.asarray(arrays[1]), sequences=arrays[2],
               targets=arrays[3], transcripts=tab)
    cases = {
        'pe100': dict(read_length=100, frag_mean=250, frag_sd=30, sub_rate=0.01, paired=True, seed=10),
        'pe150': dict(read_length=150, frag_mean=350, frag_sd=50, sub_rate=0.01, paired=True, seed=11),
        'se75': dict(read_length=75, frag_mean=250, frag_sd=30, sub_rate=0.02, paired=False, seed=12),
    }
    n_units = 3000
    for name, kw in cases.items():
        sim = synth.ReadSimulator(tx, expr, **kw)
        batches = list(sim.batches(0, n_units, batch=1024))
        per, mr, summ, x_em, tpm = run_reference(pkg, arrays, tab, batches)
        ptr, tids = csr(per.per_read)
        tl = summ.effective_lengths.astype('f8')
        x0 = numpy.ones(tl.size) / tl
        x0 /= x0.sum()
        _, iters = em_iterations(x0, tl, summ.class_map, summ.class_count)
        out.update({name + '_unit_ptr': ptr, name + '_unit_ids': tids,
                    name + '_fld': per.fragment_length_counts, name + '_class_map': summ.class_map,
                    name + '_class_count': summ.class_count, name + '_eff_lengths': summ.effective_lengths,
                    name + '_em_x': x_em, name + '_em_iters': numpy.asarray(iters), name + '_tpm': tpm})
        print('%s: classes %d aligned %d unaligned %d iters %d' % (name, summ.class_count.size,
                                                                  summ.aligned, summ.unaligned, iters))
    # adversarial reads (explicit)
    for paired in (True, False):
        reads = adversarial.make_reads(tx, paired)
        n = len(reads) // 2 if paired else len(reads)
        names = [b'a%d' % i for i in range(n)]
        per = rh.ref_map(pkg._common.KMerIndex(*arrays, tab, None), [(n, names, reads)], keep_per_read=True)
        ptr, tids = csr(per.per_read)
        key = 'adv_pe' if paired else 'adv_se'
        out.update({key + '_unit_ptr': ptr, key + '_unit_ids': tids, key + '_fld': per.fragment_length_counts})
        print(key, n, 'units; unaligned', per.counter.get((), 0))
    numpy.savez_compressed(str(OUT / 'synthetic_small.npz'), **out)


if __name__ == '__main__':
    pkg = rh.load_ref()
    golden_chr21(pkg)
    golden_synth(pkg)
