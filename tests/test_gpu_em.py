"""GPU parity tests for EM / bootstrap / effective lengths vs the numpy oracle and the golden
vectors of the reference's infer.em / infer.quantify.  Tolerance: rel 1e-6 (north_star), with
equal iteration counts; integer work (resampling) bit-exact."""
import numpy
import pytest

from conftest import SYNTH_CASES
from seekmer_b200 import _lib, infer, mapper

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def rel_close(a, b, rtol=RTOL):
    a, b = numpy.asarray(a), numpy.asarray(b)
    scale = numpy.maximum(numpy.abs(b), 1e-300)
    return bool((numpy.abs(a - b) <= rtol * scale + 1e-300).all())


class FakeIndex:
    def __init__(self, lengths):
        self.transcripts = numpy.zeros(len(lengths), dtype=[('transcript_id', 'S8'), ('length', 'f8')])
        self.transcripts['length'] = lengths


def summarized(g, prefix, lengths):
    mr = mapper.MapResult(FakeIndex(lengths))
    mr.fragment_length_counts = g[prefix + 'fld'].astype('i8')
    return mapper.SummarizedResult(
        aligned=int(g[prefix + 'class_count'].sum()), unaligned=0, total=int(g[prefix + 'class_count'].sum()),
        class_map=g[prefix + 'class_map'], class_count=g[prefix + 'class_count'],
        fragment_length_frequencies=mr.fragment_length_counts, effective_lengths=mr.effective_lengths)


@pytest.mark.parametrize('case', ['chr21'] + sorted(SYNTH_CASES))
def test_em_and_quantify_golden(golden_chr21, golden_synth, case):
    g, prefix = (golden_chr21, '') if case == 'chr21' else (golden_synth, case + '_')
    lengths = g['transcripts']['length']
    res = summarized(g, prefix, lengths)
    # effective lengths: same accumulation order => bit-exact
    assert (res.effective_lengths == g[prefix + 'eff_lengths']).all()
    eff = res.effective_lengths
    x0 = numpy.ones(eff.size) / eff
    x0 /= x0.sum()
    x, iters = infer.em(x0, eff, res.class_map, res.class_count, return_iters=True)
    assert iters == int(g[prefix + 'em_iters'])
    assert rel_close(x, g[prefix + 'em_x'])
    tpm = infer.quantify(res)
    assert rel_close(tpm, g[prefix + 'tpm'])
    assert ((tpm == 0) == (g[prefix + 'tpm'] == 0)).all()


def synthetic_structure(T, C, seed):
    rng = numpy.random.Generator(numpy.random.PCG64(seed))
    sizes = numpy.minimum(rng.geometric(0.3, size=C), 6)
    fam = rng.integers(0, T // 8, size=C)
    rows = numpy.repeat(numpy.arange(C), sizes)
    cols = (numpy.repeat(fam * 8, sizes) + rng.integers(0, 8, size=rows.size)) % T
    # a few very promiscuous transcripts and one huge class
    cols[rng.integers(0, rows.size, size=rows.size // 50)] = 3
    class_map = numpy.stack([rows, cols]).astype('i8')
    counts = rng.multinomial(3000000, rng.dirichlet(numpy.ones(C) * 0.3)).astype('f8')
    eff = rng.uniform(200, 4000, size=T)
    return class_map, counts, eff


def test_em_larger_structure_vs_oracle(orc):
    class_map, counts, eff = synthetic_structure(20000, 100000, 3)
    x0 = numpy.ones(eff.size) / eff
    x0 /= x0.sum()
    want, want_iters = orc.em(x0.copy(), eff, class_map, counts, return_iters=True)
    got, iters = infer.em(x0, eff, class_map, counts, return_iters=True)
    assert iters == want_iters
    assert rel_close(got, want)


def test_em_zero_count_classes_and_dead_transcripts(orc):
    class_map, counts, eff = synthetic_structure(4000, 20000, 4)
    counts[::3] = 0  # bootstrap-like zero classes: inner = inf / nan semantics
    x0 = numpy.ones(eff.size) / eff
    x0[::7] = 0
    x0 /= x0.sum()
    want, want_iters = orc.em(x0.copy(), eff, class_map, counts, return_iters=True)
    got, iters = infer.em(x0, eff, class_map, counts, return_iters=True)
    assert iters == want_iters
    assert rel_close(got, want)
    assert ((got == 0) == (want == 0)).all()


def test_multinomial_bit_exact_and_batched_bootstrap(orc):
    class_map, counts, eff = synthetic_structure(3000, 12000, 5)
    counts = numpy.floor(counts / 10)
    R, seed = 37, 0xDEADBEEFCAFE
    # the per-draw resampler is bit-identical to its numpy restatement
    want = orc.bootstrap_counts(counts, R, seed)
    got = infer._resample(counts, R, seed, method=_lib.RESAMPLE_DRAWS)
    assert (got == want).all()
    assert (got.sum(axis=1) == counts.sum()).all()
    # replicate ids are global: a shard starting at replicate 20 reproduces rows 20..
    assert (infer._resample(counts, 5, seed, first_replicate=20, method=_lib.RESAMPLE_DRAWS) == want[20:25]).all()

    # the bootstraps resample with the O(classes) split tree; the parity gate of BASELINE.md: the
    # reference EM (numpy restatement) fed the GPU-resampled counts agrees to 1e-6, equal iterations
    tree = infer._resample(counts, R, seed)
    assert (tree.sum(axis=1) == counts.sum()).all() and (tree[:, counts == 0] == 0).all()
    assert (infer._resample(counts, 5, seed, first_replicate=20) == tree[20:25]).all()
    res = mapper.SummarizedResult(int(counts.sum()), 0, int(counts.sum()), class_map, counts, None, eff)
    main = orc.quantify(eff, class_map, counts)
    outs, iters = infer.quantify_bootstraps(res, main, R, seed=seed, return_iters=True)
    for r in range(R):
        w, wi = orc.quantify(eff, class_map, tree[r].astype('f8'), x0=main, return_iters=True)
        assert iters[r] == wi, r
        assert rel_close(outs[r], w), r
    # a plan that owns its counts (what the mapper hands over) gives the same replicates
    ptr, tx = infer._csr_from_class_map(class_map, counts.shape[0])
    plan = _lib.EmPlan.from_csr(ptr, tx, eff.shape[0], counts=counts.astype('i8'))
    x = main / main.sum()
    o2, i2 = plan.bootstrap(eff, x, 7, seed, first_replicate=3)
    assert (i2 == iters[3:10]).all() and all((o2[k] == outs[3 + k]).all() for k in range(7))
    m2, mi = plan.run(eff, numpy.ones(eff.shape[0]) / eff / (1.0 / eff).sum())
    wm, wmi = orc.em(numpy.ones(eff.shape[0]) / eff / (1.0 / eff).sum(), eff, class_map, counts, return_iters=True)
    assert int(mi[0]) == wmi and rel_close(m2[0], wm)
    plan.close()


def test_tree_resampler_matches_its_host_build_and_the_multinomial(tmp_path):
    """The device split tree against the host build of the same header (tests/binomial_host.cpp):
    identical streams, so the counts agree except where the device's and the host's log/exp
    differ in the last place at an accept/reject boundary; and the marginals are binomial."""
    import ctypes
    import subprocess
    from conftest import ROOT
    so = tmp_path / 'libbinomial_host.so'
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', str(ROOT / 'tests' / 'binomial_host.cpp'), '-o',
                    str(so)], check=True)
    host = ctypes.CDLL(str(so))
    host.multinomial_tree_host.restype = None
    host.multinomial_tree_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64,
                                           ctypes.c_void_p]
    rng = numpy.random.Generator(numpy.random.PCG64(4))
    counts = numpy.floor(numpy.exp(rng.normal(2.5, 2.2, size=5000))).astype('i8')
    counts[rng.random(5000) < 0.15] = 0
    counts[17] = 3_000_000
    n, seed, R = int(counts.sum()), 0x1234ABCD5678, 400
    got = infer._resample(counts, R, seed)
    assert (got.sum(axis=1) == n).all() and (got[:, counts == 0] == 0).all()
    want = numpy.zeros((8, counts.shape[0]), dtype='i8')
    for r in range(8):
        host.multinomial_tree_host(counts.ctypes.data, counts.shape[0], seed, r, want[r].ctypes.data)
    assert (got[:8] != want).mean() < 1e-4
    p = counts / n
    live = counts > 0
    z = (got.mean(axis=0)[live] - n * p[live]) / numpy.sqrt(n * p[live] * (1 - p[live]) / R)
    assert abs(z).max() < 5.5 and abs(z.mean()) < 0.2 and 0.85 < z.std() < 1.15
    big = counts >= 50
    ratio = got.var(axis=0, ddof=1)[big] / (n * p[big] * (1 - p[big]))
    assert 0.93 < ratio.mean() < 1.07
    # one class, and a single replicate of many classes
    assert (infer._resample(numpy.asarray([41]), 3, 1) == 41).all()
    assert infer._resample(counts, 1, 9).sum() == n


def test_effective_lengths_bit_exact(orc):
    rng = numpy.random.Generator(numpy.random.PCG64(6))
    fld = numpy.zeros(2000, dtype='i8')
    fld[25] = 17
    fld[150:600] = rng.integers(0, 5000, size=450)
    fld[1999] = 3
    lengths = numpy.concatenate([rng.integers(25, 20000, size=5000), [1, 24, 25, 1999, 2000, 2001]]).astype('f8')
    mr = mapper.MapResult(FakeIndex(lengths))
    mr.fragment_length_counts = fld
    assert (mr.effective_lengths == orc.effective_lengths(fld, lengths)).all()


def test_em_samples_bit_identical_to_one_call_per_sample(golden_synth):
    """`skm_em_samples`: samples with different class structures, counts, effective lengths and
    iteration counts in one set of launches == `infer.quantify` sample by sample (bit for bit),
    and == the reference's first-round results of the impute fixture (1e-6)."""
    from conftest import GOLDEN, Golden
    from test_impute_host import _first_round_results
    gi = Golden(GOLDEN / 'impute_small.npz')
    g = golden_synth
    cells = _first_round_results(gi)                                   # 8 structures, shared lengths
    bulk = [summarized(g, case + '_', g['transcripts']['length']) for case in sorted(SYNTH_CASES)]
    empty = mapper.SummarizedResult(0, 5, 5, numpy.asarray([]).T, numpy.zeros(0), g['pe100_fld'],
                                    bulk[0].effective_lengths)
    samples = cells[:3] + [empty] + bulk + cells[3:]
    got, iters = infer.quantify_samples(samples, return_iters=True)
    assert got.shape == (len(samples), 60)
    for i, r in enumerate(samples):
        want = infer.quantify(r)
        assert (got[i] == want).all(), i
        if r.class_map.size:
            eff = r.effective_lengths.astype('f8')
            x0 = numpy.ones(eff.size) / eff
            x0 /= x0.sum()
            assert iters[i] == infer.em(x0, eff, r.class_map, r.class_count, return_iters=True)[1]
    assert len(set(iters.tolist())) > 3          # the samples do stop at different iterations
    assert (got[3] == 0).all() and iters[3] == 0
    for k, case in enumerate(sorted(SYNTH_CASES)):
        assert iters[4 + k] == int(g[case + '_em_iters'])
        assert rel_close(got[4 + k], g[case + '_tpm'])
    base = numpy.concatenate([got[:3], got[7:]])
    assert rel_close(base, gi['base'])
    # error behaviour: a transcript index outside the sample's range is refused
    bad = mapper.SummarizedResult(10, 0, 10, numpy.asarray([[0, 0], [1, 60]]), numpy.asarray([10.0]),
                                  g['pe100_fld'], bulk[0].effective_lengths)
    with pytest.raises(_lib.SeekmerCudaError, match='out of range'):
        infer.quantify_samples([cells[0], bad])


def test_plans_of_many_samples_run_in_one_set_of_launches(golden_synth):
    """`skm_em_plans_run` (`EmPlan.run_many`): samples whose class structures sit on the device as
    plans iterate side by side, bit-identical to one `plan.run` each; `quantify_samples` takes
    that route for results that carry a plan and stays bit-identical to `quantify`."""
    g = golden_synth
    lengths = g['transcripts']['length']
    samples = [summarized(g, case + '_', lengths) for case in sorted(SYNTH_CASES)]
    plans, effs, x0s = [], [], []
    for r in samples:
        counts = numpy.ascontiguousarray(r.class_count, dtype='i8')
        ptr, tx = infer._csr_from_class_map(r.class_map, counts.shape[0])
        plans.append(_lib.EmPlan.from_csr(ptr, tx, lengths.shape[0], counts=counts))
        eff = r.effective_lengths.astype('f8')
        x0 = numpy.ones(eff.size) / eff
        x0 /= x0.sum()
        effs.append(eff)
        x0s.append(x0)
    xs, its = _lib.EmPlan.run_many(plans, numpy.stack(effs), numpy.stack(x0s))
    for k, plan in enumerate(plans):
        one, it = plan.run(effs[k], x0s[k])
        assert (xs[k] == one[0]).all() and its[k] == it[0], k
        assert its[k] == int(g[sorted(SYNTH_CASES)[k] + '_em_iters'])
    assert len(set(its.tolist())) > 1
    # through the host API: results with a plan attached + one without
    for r, plan in zip(samples[:-1], plans[:-1]):
        r.plan = plan
    got, iters = infer.quantify_samples(samples, return_iters=True)
    for k, r in enumerate(samples):
        assert (got[k] == infer.quantify(r)).all(), k
        assert iters[k] == its[k]
    # large samples run one after the other on their plans (same results)
    saved = infer._BATCH_CLASSES_PER_SAMPLE
    infer._BATCH_CLASSES_PER_SAMPLE = 0
    try:
        serial, serial_iters = infer.quantify_samples(samples, return_iters=True)
    finally:
        infer._BATCH_CLASSES_PER_SAMPLE = saved
    assert (serial == got).all() and (serial_iters == iters).all()
    # error behaviour: a plan without counts of its own is refused
    r = samples[0]
    ptr, tx = infer._csr_from_class_map(r.class_map, r.class_count.shape[0])
    bare = _lib.EmPlan.from_csr(ptr, tx, lengths.shape[0])
    with pytest.raises(_lib.SeekmerCudaError, match='own its class counts'):
        _lib.EmPlan.run_many([plans[0], bare], numpy.stack(effs[:2]), numpy.stack(x0s[:2]))
    for plan in plans + [bare]:
        plan.close()


def test_scratch_cache_can_be_released_and_refilled(golden_synth):
    """`skm_release_cache` gives the idle scratch blocks back to the driver (cudaMalloc blocks, or - in a
    process with peer GPUs - blocks from the virtual-memory API that only the owning device maps); the next
    call simply allocates again and gives the same results."""
    import ctypes
    g = golden_synth
    lengths = g['transcripts']['length']
    r = summarized(g, sorted(SYNTH_CASES)[0] + '_', lengths)
    counts = numpy.ascontiguousarray(r.class_count, dtype='i8')
    ptr, tx = infer._csr_from_class_map(r.class_map, counts.shape[0])
    eff = r.effective_lengths.astype('f8')
    x0 = numpy.ones(eff.size) / eff
    x0 /= x0.sum()
    plan = _lib.EmPlan.from_csr(ptr, tx, lengths.shape[0], counts=counts)
    # 300 replicates of 60 transcripts are small; the draws / tree scratch of 1 500 000 padded
    # classes x replicates is not: blocks of their own, which the cache keeps when the call ends
    big_counts = numpy.ones(1_500_000, dtype='i8')
    first = infer._resample(big_counts, 24, 7)
    before, _ = plan.bootstrap(eff, x0, 8, 99)
    freed = ctypes.c_int64(0)
    try:
        for peers in (False, True, False):  # cudaMalloc blocks, VMM blocks (a process with peer GPUs), back
            _lib.uses_peer_gpus(peers)
            _lib.check(_lib.load().skm_release_cache(0, ctypes.byref(freed)))
            assert freed.value > 0
            again = infer._resample(big_counts, 24, 7)
            after, _ = plan.bootstrap(eff, x0, 8, 99)
            assert (first == again).all() and (before == after).all()
    finally:
        _lib.uses_peer_gpus(False)
    _lib.check(_lib.load().skm_release_cache(0, ctypes.byref(freed)))
    plan.close()
