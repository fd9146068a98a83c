#!/usr/bin/env python3
"""Randomised differential run of the EM entry points vs the numpy oracle (rel 1e-6, equal
iteration counts): `skm_em` with R replicates on one structure and `skm_em_samples` with a
different structure per sample.

    python tools/fuzz_em.py --seconds 20 --seed 5

Test infrastructure (the oracle is the checker).  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def close(a, b):
    a, b = numpy.asarray(a), numpy.asarray(b)
    return bool((numpy.abs(a - b) <= 1e-6 * numpy.maximum(numpy.abs(b), 1e-300) + 1e-300).all())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=20.0)
    ap.add_argument('--seed', type=int, default=5)
    args = ap.parse_args()
    from oracle import oracle as orc
    from seekmer_b200 import infer, mapper
    from test_gpu_em import synthetic_structure
    rng = numpy.random.Generator(numpy.random.PCG64(args.seed))
    t_end = time.time() + args.seconds
    runs = {'em_replicates': 0, 'em_samples': 0}
    failures = []
    while time.time() < t_end:
        T = int(rng.integers(64, 6000))
        # ---- R replicates on one structure (zero counts included, like bootstrap resamples)
        C = int(rng.integers(10, 8000))
        class_map, counts, eff = synthetic_structure(T, C, int(rng.integers(1, 1 << 30)))
        R = int(rng.integers(1, 24))
        p = counts / counts.sum()
        reps = numpy.stack([rng.multinomial(int(rng.integers(100, 2000000)), p).astype('f8') for _ in range(R)])
        x0 = numpy.ones(T) / eff
        x0 /= x0.sum()
        got, iters = infer._em_device(numpy.tile(x0, (R, 1)), eff, class_map, reps)
        runs['em_replicates'] += 1
        for r in range(R):
            want, it = orc.em(x0.copy(), eff, class_map, reps[r], return_iters=True)
            if it != iters[r] or not close(got[r], want):
                failures.append({'kind': 'em_replicates', 'T': T, 'C': C, 'R': R, 'replicate': r,
                                 'iters': [int(iters[r]), int(it)]})
                break
        # ---- samples with their own structures
        P = int(rng.integers(1, 10))
        samples = []
        for _ in range(P):
            Cs = int(rng.integers(5, 4000))
            cm, cnt, eff_s = synthetic_structure(T, Cs, int(rng.integers(1, 1 << 30)))
            cnt = rng.multinomial(int(rng.integers(50, 500000)), cnt / cnt.sum()).astype('f8')
            keep = cnt > 0                       # a summarised sample holds observed classes only
            new_id = numpy.cumsum(keep) - 1
            sel = keep[cm[0]]
            cm = numpy.stack([new_id[cm[0][sel]], cm[1][sel]])
            samples.append(mapper.SummarizedResult(int(cnt.sum()), 0, int(cnt.sum()), cm, cnt[keep], None, eff_s))
        got, iters = infer.quantify_samples(samples, return_iters=True)
        runs['em_samples'] += 1
        for k, smp in enumerate(samples):
            want, it = orc.quantify(smp.effective_lengths, smp.class_map, smp.class_count, return_iters=True)
            if it != iters[k] or not close(got[k], want):
                failures.append({'kind': 'em_samples', 'T': T, 'P': P, 'sample': k,
                                 'iters': [int(iters[k]), int(it)]})
                break
    print(json.dumps({'seed': args.seed, 'seconds': args.seconds, 'runs': runs,
                      'mismatching_runs': len(failures), 'failures': failures[:5]}))


if __name__ == '__main__':
    main()
