"""The sort-based index builder (seekmer_b200/index_build.py) produces a valid index in the
reference's array layout, with the reference's contig partition."""
import numpy
import pytest
import torch

from seekmer_b200 import index_build, synth


@pytest.fixture(scope='module')
def built():
    tx = synth.make_transcriptome(300, seed=9)
    return tx, index_build.build_index(tx.codes, tx.offsets, device='cpu')


def contig_set(contigs, sequences):
    raw = numpy.asarray(sequences).tobytes()
    out = set()
    for o, l in zip(contigs['offset'].tolist(), contigs['length'].tolist()):
        q = raw[o:o + l]
        out.add(min(q, synth.reverse_complement_ascii(q)))
    return out


def test_structural_invariants(built):
    # the reference's own index checks (test/test_index_builder.py:46-75)
    tx, b = built
    k, c, s, t = b.numpy_arrays()
    n_slots = k.shape[0]
    assert n_slots & (n_slots - 1) == 0 and b.stats['n_kmers'] <= 0.8 * n_slots
    occ = k[k['kmer'] != numpy.uint64(0xFFFFFFFFFFFFFFFF)]
    assert occ.shape[0] == b.stats['n_kmers']
    assert occ['entry'].min() >= 0 and occ['entry'].max() == c.shape[0] - 1
    assert (occ['offset'] >= 0).all() and (occ['offset'] + 25 <= c['length'][occ['entry']]).all()
    assert c['offset'][0] == 0 and (c['offset'][1:] == numpy.cumsum(c['length'])[:-1]).all()
    assert c['offset'][-1] + c['length'][-1] == s.shape[0]
    assert set(numpy.unique(s.view('u1'))) <= set(b'ACGT')
    assert c['target_offset'][0] == 0 and (c['target_count'] >= 1).all()
    assert c['target_offset'][-1] + c['target_count'][-1] == t.shape[0]
    ent = numpy.where(t['entry'] < 0, ~t['entry'], t['entry'])
    assert ent.min() >= 0 and ent.max() < tx.n_transcripts
    # per-contig lists sorted by signed (entry, offset) (`_index_builder.pyx:540`)
    contig_of = numpy.repeat(numpy.arange(c.shape[0]), c['target_count'])
    key = numpy.stack([contig_of, t['entry'].astype('i8'), t['offset'].astype('i8')], axis=1)
    assert (numpy.lexsort((key[:, 2], key[:, 1], key[:, 0])) == numpy.arange(t.shape[0])).all()
    # number of k-mers: every contig of length L holds L-24 of them
    assert (c['length'] - 24).sum() == b.stats['n_kmers']


def test_every_transcript_kmer_resolves_to_its_contig_sequence(built, orc):
    tx, b = built
    k, c, s, t = b.numpy_arrays()
    idx = orc.OracleIndex(k, c, s, t)
    raw = s.tobytes()
    rng = numpy.random.Generator(numpy.random.PCG64(2))
    seqs = tx.sequences()
    for _ in range(3000):
        ti = int(rng.integers(0, len(seqs)))
        p = int(rng.integers(0, len(seqs[ti]) - 24))
        kmer = seqs[ti][p:p + 25]
        e, o = idx.map_kmer(orc.encode(kmer))
        assert o >= 0
        ci = e if e >= 0 else ~e
        got = raw[c['offset'][ci] + o:c['offset'][ci] + o + 25]
        assert got == (kmer if e >= 0 else synth.reverse_complement_ascii(kmer))
        # first/last k-mer fields
        assert c['first_kmer'][ci] == orc.encode(raw[c['offset'][ci]:c['offset'][ci] + 25])
        end = c['offset'][ci] + c['length'][ci]
        assert c['last_kmer'][ci] == orc.encode(raw[end - 25:end])
        # the transcript is in the contig's target list with the right strand
        lst = t['entry'][c['target_offset'][ci]:c['target_offset'][ci] + c['target_count'][ci]]
        assert (ti if e >= 0 else ~ti) in lst.tolist()


def test_same_partition_and_same_mapping_as_the_reference_index(built, orc, ref):
    tx, b = built
    k, c, s, t = b.numpy_arrays()
    rk, rc, rs, rt = ref.ref_build_index(tx.sequences())
    mine, theirs = contig_set(c, s), contig_set(numpy.asarray(rc), rs)
    # the reference assembler mishandles the k-mer that lands in hash slot 0 (slot 0 doubles as
    # "no link"): at most that contig may differ
    assert len(theirs - mine) <= 3 and len(mine - theirs) <= 2
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), 100, 250, 30, seed=4)
    n = 20000
    bases, _ = sim.generate(0, n)
    a = orc.map_batch(orc.OracleIndex(k, c, s, t), bases, sim.offsets(n), True)
    r = orc.map_batch(orc.OracleIndex(rk, rc, rs, rt), bases, sim.offsets(n), True)
    ta, tr = a.tuples(), r.tuples()
    differ = sum(1 for x, y in zip(ta, tr) if x != y)
    assert differ <= n // 500, differ
    assert numpy.abs(a.fld - r.fld).sum() <= n // 250


def test_reference_hash_matches_oracle(orc):
    rng = numpy.random.Generator(numpy.random.PCG64(3))
    ks = rng.integers(0, 1 << 50, size=500, dtype='i8')
    h = index_build.reference_hash(torch.from_numpy(ks)).numpy().view('u8')
    for kk, hh in zip(ks.tolist(), h.tolist()):
        assert orc.kmer_hash(kk) == hh
    assert (index_build._revcomp(torch.from_numpy(ks)).numpy()
            == numpy.asarray([orc.reverse_complement(x) for x in ks.tolist()], dtype='i8')).all()


@pytest.mark.gpu
def test_gpu_build_equals_cpu_build(built, orc):
    tx, b = built
    g = index_build.build_index(tx.codes, tx.offsets, device='cuda')
    assert g.stats == b.stats
    k, c, s, t = b.numpy_arrays()
    gk, gc, gs, gt = g.numpy_arrays()
    assert (gc == c).all() and (gs == s).all() and (gt == t).all()
    # slot placement depends on insertion order; contents and reachability must agree
    occ, gocc = k[k['kmer'] != numpy.uint64(2**64 - 1)], gk[gk['kmer'] != numpy.uint64(2**64 - 1)]
    assert (numpy.sort(occ, order=['kmer']) == numpy.sort(gocc, order=['kmer'])).all()
    idx = orc.OracleIndex(gk, gc, gs, gt)
    for row in gocc[::37]:
        assert idx.map_kmer(int(row['kmer'])) == (int(row['entry']), int(row['offset']))
