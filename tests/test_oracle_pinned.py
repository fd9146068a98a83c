"""Pin the CPU oracle (oracle/seekmer_oracle.c + oracle/oracle.py) against the reference.

(a) always: against the committed golden vectors (outputs of the unmodified reference,
    tests/golden/make_golden.py);
(b) when oracle/_ref is present: live against the compiled reference on fresh inputs.
"""
import numpy
import pytest

import adversarial
from conftest import N_GOLDEN_UNITS, SYNTH_CASES
from seekmer_b200 import synth


def test_chr21_fixture_known_answer(orc, golden_chr21):
    g = golden_chr21
    idx = orc.OracleIndex(*g.index_arrays())
    reads = [bytes(r) for r in g['reads']]
    bases, offs = orc.pack_reads(reads)
    out = orc.map_batch(idx, bases, offs, paired=True)
    assert out.tuples() == g.tuples('')
    assert (out.fld == g['fld']).all()
    # the reference's only result-pinning assertion (test/test_mapper.py:76)
    assert sum(1 for t in out.tuples() if not t) == 0 == int(g['unaligned'])
    cls_ptr, cls_ids, cls_count, una = orc.tally(out.ptr, out.ids)
    cm = orc.class_map_from_csr(cls_ptr, cls_ids)
    assert (cm == g['class_map']).all()
    assert (cls_count == g['class_count']).all()
    eff = orc.effective_lengths(out.fld, g['transcripts']['length'])
    assert (eff == g['eff_lengths']).all()
    assert orc.harmonic_mean_fragment_length(out.fld) == pytest.approx(float(g['harmonic_mean']), rel=1e-15)
    x0 = numpy.ones(eff.size) / eff
    x0 /= x0.sum()
    x, iters = orc.em(x0, eff, cm, cls_count.astype('f8'), return_iters=True)
    assert (x == g['em_x']).all()
    assert iters == int(g['em_iters'])
    tpm = orc.quantify(eff, cm, cls_count)
    assert (tpm == g['tpm']).all()


@pytest.mark.parametrize('case', sorted(SYNTH_CASES))
def test_synthetic_golden(orc, golden_synth, small_tx, case):
    g = golden_synth
    idx = orc.OracleIndex(*g.index_arrays())
    kw = SYNTH_CASES[case]
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    bases, _ = sim.generate(0, N_GOLDEN_UNITS)
    out = orc.map_batch(idx, bases, sim.offsets(N_GOLDEN_UNITS), kw['paired'])
    assert out.tuples() == g.tuples(case + '_')
    assert (out.fld == g[case + '_fld']).all()
    cls_ptr, cls_ids, cls_count, una = orc.tally(out.ptr, out.ids)
    cm = orc.class_map_from_csr(cls_ptr, cls_ids)
    assert (cm == g[case + '_class_map']).all()
    assert (cls_count == g[case + '_class_count']).all()
    eff = orc.effective_lengths(out.fld, g['transcripts']['length'])
    assert (eff == g[case + '_eff_lengths']).all()
    x0 = numpy.ones(eff.size) / eff
    x0 /= x0.sum()
    x, iters = orc.em(x0, eff, cm, cls_count.astype('f8'), return_iters=True)
    assert (x == g[case + '_em_x']).all()
    assert iters == int(g[case + '_em_iters'])
    assert (orc.quantify(eff, cm, cls_count) == g[case + '_tpm']).all()


@pytest.mark.parametrize('paired', [True, False])
def test_adversarial_golden(orc, golden_synth, small_tx, paired):
    g = golden_synth
    idx = orc.OracleIndex(*g.index_arrays())
    reads = adversarial.make_reads(small_tx, paired)
    bases, offs = orc.pack_reads(reads)
    out = orc.map_batch(idx, bases, offs, paired)
    key = 'adv_pe_' if paired else 'adv_se_'
    assert out.tuples() == g.tuples(key)
    assert (out.fld == g[key + 'fld']).all()


def test_primitives_against_table_invariant(orc, golden_chr21):
    """Every occupied slot is reachable from hash(min(k, rc)) & mask without crossing an empty
    slot and map_kmer returns its stored position on both strands."""
    g = golden_chr21
    idx = orc.OracleIndex(*g.index_arrays())
    kmers = g['kmers']
    occ = numpy.nonzero(kmers['kmer'] != numpy.uint64(0xFFFFFFFFFFFFFFFF))[0]
    rng = numpy.random.Generator(numpy.random.PCG64(0))
    for i in rng.choice(occ, size=2000, replace=False):
        k = int(kmers['kmer'][i])
        assert idx.map_kmer(k) == (int(kmers['entry'][i]), int(kmers['offset'][i]))
        rc = orc.reverse_complement(k)
        assert orc.reverse_complement(rc) == k
        assert idx.map_kmer(rc) == (~int(kmers['entry'][i]), int(kmers['offset'][i]))
    assert idx.map_kmer(0x123456789ABC) == (0, -1)
    assert orc.encode(b'ACGTACGTACGTACGTACGTACGTAnnnn') == int('0123' * 6 + '0', 4)


def test_live_against_compiled_reference(orc, ref, medium):
    """Fresh inputs, not in the golden files: per-read tuples, FLD and dict vs oracle/_ref."""
    tx, arrays = medium
    ridx = ref.ref_index_from_arrays(*arrays)
    oidx = orc.OracleIndex(*arrays)
    expr = synth.make_expression(tx.n_transcripts)
    for L, mu, sd, paired, sub in [(100, 250, 30, True, 0.01), (150, 350, 50, True, 0.02),
                                   (75, 250, 30, False, 0.02), (36, 250, 30, True, 0.0)]:
        sim = synth.ReadSimulator(tx, expr, L, mu, sd, sub_rate=sub, paired=paired, seed=77)
        n = 6000
        res = ref.ref_map(ridx, list(sim.batches(0, n, batch=2048)), keep_per_read=True)
        bases, _ = sim.generate(0, n)
        out = orc.map_batch(oidx, bases, sim.offsets(n), paired)
        assert out.tuples() == res.per_read
        assert (out.fld == res.fragment_length_counts).all()
        assert orc.tally_dict(out.ptr, out.ids) == dict(res.counter)
    for paired in (True, False):
        reads = adversarial.make_reads(tx, paired)
        n = len(reads) // 2 if paired else len(reads)
        res = ref.ref_map(ridx, [(n, [b'x'] * n, reads)], keep_per_read=True)
        bases, offs = orc.pack_reads(reads)
        out = orc.map_batch(oidx, bases, offs, paired)
        assert out.tuples() == res.per_read
        assert (out.fld == res.fragment_length_counts).all()


def test_threaded_reference_is_j_invariant(ref, medium):
    tx, arrays = medium
    ridx = ref.ref_index_from_arrays(*arrays)
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), 100, 250, 30, seed=5)
    batches = list(sim.batches(0, 8000, batch=1000))
    a = ref.ref_map(ridx, batches)
    b = ref.ref_map_threads(ridx, batches, 4)
    assert dict(a.counter) == dict(b.counter)
    assert (a.fragment_length_counts == b.fragment_length_counts).all()


def test_bootstrap_resampler_distribution(orc):
    counts = numpy.asarray([0, 5, 100, 1, 0, 894, 3000], dtype='i8')
    out = orc.bootstrap_counts(counts, 64, seed=1234)
    n = counts.sum()
    assert (out.sum(axis=1) == n).all()
    assert (out[:, counts == 0] == 0).all()
    p = counts / n
    mean = out.mean(axis=0)
    sd = numpy.sqrt(n * p * (1 - p) / 64) + 1e-9
    assert (numpy.abs(mean - n * p) < 5 * sd + 1e-9).all()
    again = orc.bootstrap_counts(counts, 2, seed=1234)
    assert (again == out[:2]).all()
