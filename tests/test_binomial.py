"""CPU goodness-of-fit tests of the exact binomial sampler (seekmer_b200/csrc/binomial.cuh, host
build) and of the binary split tree that the O(classes) bootstrap resampler is made of."""
import ctypes
import subprocess

import numpy
import pytest
import scipy.stats

from conftest import ROOT


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    out = tmp_path_factory.mktemp('binomial') / 'libbinomial_host.so'
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', str(ROOT / 'tests' / 'binomial_host.cpp'),
                    '-o', str(out)], check=True)
    lib = ctypes.CDLL(str(out))
    lib.binomial_batch.restype = None
    lib.binomial_batch.argtypes = [ctypes.c_int64, ctypes.c_double, ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p]
    lib.multinomial_tree_host.restype = None
    lib.multinomial_tree_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64,
                                          ctypes.c_void_p]
    return lib


def draws(host, n, p, seed, count):
    out = numpy.zeros(count, dtype='i8')
    host.binomial_batch(n, p, seed, count, out.ctypes.data)
    return out


CASES = [(1, 0.3), (7, 0.5), (40, 0.1), (40, 0.9), (200, 0.14), (200, 0.15), (200, 0.16), (1000, 0.03), (1000, 0.5),
         (12345, 0.77), (10 ** 6, 0.25), (3 * 10 ** 7, 1e-6), (3 * 10 ** 7, 0.5), (3 * 10 ** 7, 0.999), (10 ** 9, 0.37)]


@pytest.mark.parametrize('n,p', CASES)
def test_binomial_draws_fit_the_distribution(host, n, p):
    """Chi-square of 200 000 draws against scipy.stats.binom over equal-probability-ish bins, plus
    range, mean and variance.  Both algorithms (inversion below n min(p, 1-p) = 30, BTPE above)
    and the p > 0.5 flip are covered by the cases."""
    count = 200_000
    x = draws(host, n, p, 0xC0FFEE + n, count)
    assert x.min() >= 0 and x.max() <= n
    dist = scipy.stats.binom(n, p)
    # bin edges at quantiles of the true distribution; merge to expected counts >= 50
    qs = numpy.unique(dist.ppf(numpy.linspace(0, 1, 41)[1:-1]).astype('i8'))
    edges = numpy.concatenate([[-1], qs, [n]])
    cdf = dist.cdf(edges)
    cdf[0] = 0.0
    expected = numpy.diff(cdf) * count
    observed = numpy.histogram(x, bins=edges + 0.5)[0]
    keep = expected >= 50
    if keep.sum() >= 2:
        obs = numpy.append(observed[keep], observed[~keep].sum())
        exp = numpy.append(expected[keep], expected[~keep].sum())
        if exp[-1] == 0:
            obs, exp = obs[:-1], exp[:-1]
        stat, pval = scipy.stats.chisquare(obs, exp * obs.sum() / exp.sum())
        assert pval > 1e-4, (n, p, stat, pval)
    mean, var = n * p, n * p * (1 - p)
    assert abs(x.mean() - mean) < 6 * numpy.sqrt(var / count) + 1e-9
    if var > 0:
        assert abs(x.var() - var) < 0.05 * var + 1e-9


def test_degenerate_probabilities(host):
    assert (draws(host, 100, 0.0, 1, 100) == 0).all()
    assert (draws(host, 100, 1.0, 1, 100) == 100).all()
    assert (draws(host, 0, 0.5, 1, 100) == 0).all()


def test_split_tree_is_a_multinomial(host):
    """Replicates of the tree resampler: every replicate sums to n, zero-count classes stay zero,
    and the marginal of each class over many replicates has the binomial mean and variance
    (z-scores over classes behave like standard normals)."""
    rng = numpy.random.Generator(numpy.random.PCG64(11))
    counts = numpy.floor(numpy.exp(rng.normal(2.0, 2.0, size=777))).astype('i8')
    counts[rng.random(777) < 0.2] = 0
    counts[5] = 400_000  # one dominant class
    n = int(counts.sum())
    R = 600
    out = numpy.zeros((R, counts.shape[0]), dtype='i8')
    for r in range(R):
        host.multinomial_tree_host(counts.ctypes.data, counts.shape[0], 0xABCDEF12345, r, out[r].ctypes.data)
    assert (out.sum(axis=1) == n).all()
    assert (out[:, counts == 0] == 0).all()
    p = counts / n
    live = counts > 0
    z_mean = (out.mean(axis=0)[live] - n * p[live]) / numpy.sqrt(n * p[live] * (1 - p[live]) / R)
    assert abs(z_mean).max() < 5.0 and abs(z_mean.mean()) < 0.25 and 0.8 < z_mean.std() < 1.2
    big = counts >= 50
    ratio = out.var(axis=0, ddof=1)[big] / (n * p[big] * (1 - p[big]))
    assert 0.9 < ratio.mean() < 1.1
    # two classes are negatively correlated like a multinomial: cov = -n p_i p_j
    i, j = 5, int(numpy.argsort(counts)[-2])
    cov = numpy.cov(out[:, i], out[:, j])[0, 1]
    want = -n * p[i] * p[j]
    assert abs(cov - want) < 0.35 * abs(want)
    # replicates are functions of (seed, replicate): a second pass reproduces them
    again = numpy.zeros_like(out[3])
    host.multinomial_tree_host(counts.ctypes.data, counts.shape[0], 0xABCDEF12345, 3, again.ctypes.data)
    assert (again == out[3]).all()
