#!/usr/bin/env python3
"""Cold and warm timing of quantify_bootstraps on the §8(d) micro-benchmark structure."""
import os, sys, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekmer_b200 import infer, mapper
from tools.profile_em import structure
class_map, counts, eff = structure()
x0 = numpy.ones(eff.size) / eff; x0 /= x0.sum()
t = time.perf_counter(); x, it = infer.em(x0, eff, class_map, counts, return_iters=True); print('main EM cold %.1f ms' % ((time.perf_counter() - t) * 1e3))
tpm = infer._finish(x.copy())
summ = mapper.SummarizedResult(int(counts.sum()), 0, int(counts.sum()), class_map, counts, None, eff)
for rep in range(2):
    t = time.perf_counter()
    out, its = infer.quantify_bootstraps(summ, tpm, 100, seed=1234, return_iters=True)
    print('quantify_bootstraps x100: %.1f ms' % ((time.perf_counter() - t) * 1e3), flush=True)
