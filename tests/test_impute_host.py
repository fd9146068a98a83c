"""Host side of the `impute` workflow against the reference's own outputs
(tests/golden/impute_small.npz, made by tests/golden/make_golden_impute.py).  No GPU needed:
cell weights, blending and the gene table are host logic; the batched EM is in
tests/test_gpu_impute.py."""
import numpy
import pytest

from conftest import GOLDEN, Golden
from seekmer_b200 import impute, mapper


@pytest.fixture(scope='module')
def gi():
    return Golden(GOLDEN / 'impute_small.npz')


class _Index:
    def __init__(self, transcripts):
        self.transcripts = transcripts


def _index(gi, golden_synth):
    tab = golden_synth['transcripts'].copy()
    tab['gene_id'] = gi['gene_id']
    return _Index(tab)


def _first_round_results(gi):
    """SummarizedResult objects as the reference had them before blending."""
    ptr, nnz = gi['class_ptr'], numpy.concatenate([[0], numpy.cumsum(gi['class_nnz'])])
    out = []
    for c in range(len(ptr) - 1):
        counts = gi['class_count'][ptr[c]:ptr[c + 1]].copy()
        cmap = gi['class_map'][:, nnz[c]:nnz[c + 1]].copy()
        out.append(mapper.SummarizedResult(int(counts.sum()), 0, int(counts.sum()), cmap, counts,
                                           gi['fld'], gi['eff_lengths']))
    return out


def test_cell_weights_match_reference(gi, golden_synth, tmp_path):
    idx = _index(gi, golden_synth)
    numpy.random.seed(1)                       # the seed the fixture was generated under
    w = impute._calculate_cell_weights(idx, gi['base'], tmp_path)
    assert (w == gi['weight']).all()
    assert (tmp_path / 'initial_gene_table.csv').read_bytes() == gi['gene_table_csv'].tobytes()
    assert (tmp_path / 'weight.csv').exists()
    # the deterministic split agrees on this (clearly bimodal) fixture, whatever the RNG state
    numpy.random.seed(99)
    exact = impute._calculate_cell_weights(idx, gi['base'], None, clustering='exact')
    assert (exact == gi['weight']).all()
    with pytest.raises(ValueError):
        impute._calculate_cell_weights(idx, gi['base'], None, clustering='nope')


def test_blend_matches_reference(gi):
    cells = _first_round_results(gi)
    impute._blend_mapping_results(cells, gi['weight'] ** int(gi['power']))
    assert all(c.class_map is cells[0].class_map for c in cells)
    assert (cells[0].class_map == gi['blended_map']).all()
    assert numpy.allclose(cells[0].class_count, gi['blended_count_first'], rtol=1e-13, atol=0)
    assert numpy.allclose(cells[-1].class_count, gi['blended_count_last'], rtol=1e-13, atol=0)
    # class mass of a cell is preserved up to the weights: row i sums to total_i * sum_j w_ij
    totals = [gi['class_count'][gi['class_ptr'][c]:gi['class_ptr'][c + 1]].sum() for c in range(len(cells))]
    w = gi['weight'] ** int(gi['power'])
    for i, c in enumerate(cells):
        assert c.class_count.sum() == pytest.approx(totals[i] * w[i].sum(), rel=1e-12)


def test_merge_fragment_lengths_shares_one_array():
    class R:
        def __init__(self, k):
            self.fragment_length_counts = numpy.zeros(mapper.MAX_FRAGMENT_LENGTH, dtype='i8')
            self.fragment_length_counts[k] = k
    rs = [R(3), R(5), R(5)]
    impute._merge_fragment_lengths(rs)
    assert all(r.fragment_length_counts is rs[0].fragment_length_counts for r in rs)
    assert rs[0].fragment_length_counts[3] == 3 and rs[0].fragment_length_counts[5] == 10
    assert rs[0].fragment_length_counts.sum() == 13


def test_two_means_is_the_optimal_split():
    rng = numpy.random.default_rng(5)
    for _ in range(50):
        v = rng.normal(size=rng.integers(2, 12))
        lo, hi = impute._two_means(v)
        s = numpy.sort(v)
        best = min(((s[:k] - s[:k].mean()) ** 2).sum() + ((s[k:] - s[k:].mean()) ** 2).sum()
                   for k in range(1, len(s)))
        assign_hi = numpy.abs(v - hi) < numpy.abs(v - lo)
        assert 0 < assign_hi.sum() < len(v)
        cost = ((v[assign_hi] - hi) ** 2).sum() + ((v[~assign_hi] - lo) ** 2).sum()
        assert cost == pytest.approx(best, rel=1e-9, abs=1e-12)
    with pytest.raises(ValueError):
        impute._two_means([0.5])


def test_gene_matrix_truncates_like_the_reference(gi, golden_synth):
    idx = _index(gi, golden_synth)
    m, names = impute._gene_matrix(idx, gi['base'])
    assert m.dtype == numpy.dtype('i8') and b'' not in set(names.tolist())
    genes = numpy.unique(idx.transcripts['gene_id'])
    assert m.shape == (gi['base'].shape[0], len(genes) - 1)
    g0 = idx.transcripts['gene_id'] == names[0]
    assert (m[:, 0] == numpy.trunc(gi['base'][:, g0].sum(axis=1))).all()


def test_cli_lists_impute():
    import argparse
    sub = argparse.ArgumentParser().add_subparsers(dest='subcommand')
    impute.add_subcommand_parser(sub)
    ns = sub.choices['impute'].parse_args(['i.npz', 'out', 'a.fq', 'b.fq', '-s', '-p', '4'])
    assert ns.single_ended and ns.power == 4 and ns.job_count == 1 and len(ns.fastq_paths) == 2


def test_oracle_em_on_blended_counts_is_pinned(gi, orc):
    """The oracle's EM on the reference's blended (fractional, mostly zero) class counts
    reproduces the reference's second-round TPM: the GPU test may then use it as the checker."""
    for counts, want in ((gi['blended_count_first'], gi['tpm'][0]), (gi['blended_count_last'], gi['tpm'][-1])):
        got = orc.quantify(gi['eff_lengths'], gi['blended_map'], counts)
        assert numpy.allclose(got, want, rtol=1e-12, atol=0)


def test_gene_sums_equal_the_per_gene_masked_sums():
    """Bit-for-bit the reference's `base_matrix[:, gene_mask].sum(axis=1)` (`impute.py:200-202`),
    including genes large enough for numpy's pairwise summation."""
    rng = numpy.random.default_rng(11)
    sizes = numpy.concatenate([rng.integers(1, 7, 40), [8, 9, 33, 130, 200]])
    gene_of = rng.permutation(numpy.repeat(numpy.arange(len(sizes)), sizes))
    tab = numpy.zeros(len(gene_of), dtype=[('transcript_id', 'S8'), ('gene_id', 'S8'), ('length', 'f8')])
    tab['gene_id'] = [b'g%04d' % g if g else b'' for g in gene_of]
    base = rng.lognormal(3, 3, size=(7, len(gene_of)))
    got, names = impute._gene_matrix(_Index(tab), base)
    genes, inverse = numpy.unique(tab['gene_id'], return_inverse=True)
    want = numpy.zeros((7, len(genes)), dtype='i8')
    for g in range(len(genes)):
        want[:, g] = base[:, inverse == g].sum(axis=1)
    assert (got == want[:, genes != b'']).all() and (names == genes[genes != b'']).all()
    exact = numpy.zeros((7, len(genes)))
    for g in range(len(genes)):
        exact[:, g] = base[:, inverse == g].sum(axis=1)
    sums = numpy.zeros_like(exact)   # the float sums themselves, before truncation
    order = numpy.argsort(inverse, kind='stable')
    pos = 0
    for g, k in enumerate(numpy.bincount(inverse)):
        sums[:, g] = base[:, order[pos:pos + k]][:, None, :].sum(axis=2)[:, 0]
        pos += k
    assert (sums == exact).all()


def test_prune_and_support_groups(gi):
    cells = _first_round_results(gi)
    w = gi['weight'] ** int(gi['power'])
    impute._blend_mapping_results(cells, w)
    groups = impute._support_groups(w)
    assert sorted(map(tuple, groups)) == [(0, 2, 4, 6), (1, 3, 5, 7)]
    counts = numpy.stack([cells[i].class_count for i in groups[0]])
    cmap, pruned = impute._prune_classes(cells[0].class_map, counts)
    ptr = gi['class_ptr']
    kept = sum(int(ptr[c + 1] - ptr[c]) for c in groups[0])
    assert pruned.shape == (4, kept) and (pruned != 0).any(axis=0).all()
    assert cmap[0].max() == kept - 1 and (numpy.diff(cmap[0]) >= 0).all()
    # the surviving classes keep their transcripts, in order
    full = cells[0].class_map
    active = (counts != 0).any(axis=0)
    assert (cmap[1] == full[1][active[full[0]]]).all()
    assert pruned.sum() == pytest.approx(counts.sum(), rel=1e-15)
    # nothing to drop / everything zero: returned untouched
    same_map, same = impute._prune_classes(full, numpy.ones((2, counts.shape[1])))
    assert same_map is full and same.shape == (2, counts.shape[1])


def test_blended_group_equals_blend_then_prune(gi):
    w = gi['weight'] ** int(gi['power'])
    dense = _first_round_results(gi)
    impute._blend_mapping_results(dense, w)
    cells = _first_round_results(gi)
    for group in impute._support_groups(w):
        class_map, counts = impute._blended_group(cells, w, group)
        want_map, want = impute._prune_classes(dense[0].class_map,
                                               numpy.stack([dense[i].class_count for i in group]))
        assert (class_map == want_map).all() and class_map.dtype == want_map.dtype
        assert (counts == want).all()
    # the inputs are left as they were (the reference's blend renumbers them in place)
    fresh = _first_round_results(gi)
    assert all((a.class_map == b.class_map).all() for a, b in zip(cells, fresh))


def test_cells_without_support_are_reported_as_zeros(gi):
    cells = _first_round_results(gi)[:2]
    out = impute._quantify_weighted(cells, numpy.zeros((2, 2)))   # no device work to do
    assert out.shape == (2, 60) and (out == 0).all()


def test_quantify_samples_chunks_keep_their_places(gi, orc, monkeypatch):
    """Host bookkeeping of `infer.quantify_samples` (which samples go into which device call,
    where their rows land, empty samples), with the device call replaced by the oracle's EM."""
    from seekmer_b200 import infer
    calls = []

    def fake_call(samples, n_tx, device=0):
        calls.append(len(samples))
        xs, its = [], []
        for r in samples:
            eff = r.effective_lengths.astype('f8')
            x0 = numpy.ones(n_tx) / eff
            x0 /= x0.sum()
            x, it = orc.em(x0, eff, r.class_map, r.class_count, return_iters=True)
            xs.append(x)
            its.append(it)
        return numpy.stack(xs), numpy.asarray(its, dtype='i4')

    monkeypatch.setattr(infer, '_em_samples_device', fake_call)
    monkeypatch.setattr(infer, '_SAMPLE_ROWS_PER_CALL', 3 * 60)          # three samples per call
    cells = _first_round_results(gi)
    empty = mapper.SummarizedResult(0, 2, 2, numpy.asarray([]).T, numpy.zeros(0), gi['fld'], gi['eff_lengths'])
    samples = [empty] + cells[:4] + [empty] + cells[4:]
    got, iters = infer.quantify_samples(samples, return_iters=True)
    assert calls == [3, 3, 2]
    live = [i for i in range(len(samples)) if i not in (0, 5)]
    assert numpy.allclose(got[live], gi['base'], rtol=1e-12, atol=0)
    assert (got[[0, 5]] == 0).all() and (iters[[0, 5]] == 0).all() and (iters[live] > 0).all()
    assert infer.quantify_samples([]).shape == (0, 0)
