#!/usr/bin/env python3
"""EM / bootstrap timing on the SURVEY §8(d) micro-benchmark structure (T=2e5, C=1e6, nnz~3e6)."""
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from seekmer_b200 import infer, mapper  # noqa: E402


def structure(T=200000, C=1000000, seed=1, reads=30000000):
    rng = numpy.random.Generator(numpy.random.PCG64(seed))
    sizes = numpy.minimum(rng.geometric(0.3, size=C), 6)
    fam = rng.integers(0, T // 8, size=C)
    rows = numpy.repeat(numpy.arange(C), sizes)
    cols = (numpy.repeat(fam * 8, sizes) + rng.integers(0, 8, size=rows.size)) % T
    class_map = numpy.stack([rows, cols]).astype('i8')
    counts = rng.multinomial(reads, rng.dirichlet(numpy.ones(C) * 0.3)).astype('f8')
    eff = rng.uniform(200, 4000, size=T)
    return class_map, counts, eff


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    class_map, counts, eff = structure()
    print('T=%d C=%d nnz=%d reads=%d' % (eff.size, counts.size, class_map.shape[1], counts.sum()))
    x0 = numpy.ones(eff.size) / eff
    x0 /= x0.sum()
    for rep in range(2):
        t = time.perf_counter()
        x, iters = infer.em(x0, eff, class_map, counts, return_iters=True)
        print('main EM: %.1f ms, %d iterations' % ((time.perf_counter() - t) * 1e3, iters))
    main_tpm = infer._finish(x.copy())
    for rep in range(2):
        t = time.perf_counter()
        res = infer._resample(counts, R, 1234)
        t1 = time.perf_counter()
        print('resample x%d: %.1f ms' % (R, (t1 - t) * 1e3))
    summ = mapper.SummarizedResult(int(counts.sum()), 0, int(counts.sum()), class_map, counts, None, eff)
    for rep in range(2):
        t = time.perf_counter()
        out, its = infer.quantify_bootstraps(summ, main_tpm, R, seed=1234, return_iters=True)
        print('quantify_bootstraps x%d: %.1f ms, mean iters %.1f max %d' % (R, (time.perf_counter() - t) * 1e3,
                                                                              its.mean(), its.max()))


if __name__ == '__main__':
    main()
