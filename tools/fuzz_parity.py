#!/usr/bin/env python3
"""Randomised differential run: CUDA mapper (through the C ABI) vs the CPU oracle.

    python tools/fuzz_parity.py --seconds 45 --seed 3

Random transcriptomes (isoform families of random shape, indexed by the reference assembler in
oracle/_ref), random read lengths / fragment sizes / error and N rates / single or paired, plus
the adversarial read set on every transcriptome.  Compares per-unit ordered id tuples, raw
fragment lengths, the FLD and the class dictionary in first-seen order, bit for bit.  Prints one
JSON line.  Test infrastructure (the oracle is the checker)."""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=45.0)
    ap.add_argument('--seed', type=int, default=3)
    ap.add_argument('--units', type=int, default=20000)
    args = ap.parse_args()
    from oracle import oracle as orc, ref_harness as ref
    from seekmer_b200 import synth
    import adversarial
    import test_gpu_mapper as tg
    orc.build()
    ref.load_ref()
    rng = numpy.random.Generator(numpy.random.PCG64(args.seed))
    t_end = time.time() + args.seconds
    runs, units, failures = 0, 0, []
    while time.time() < t_end:
        n_tx = int(rng.integers(20, 1500))
        kw = dict(seed=int(rng.integers(1, 1 << 30)), mean_exons=int(rng.integers(3, 30)),
                  median_exon=int(rng.integers(40, 300)), min_exon=int(rng.integers(5, 40)),
                  max_isoforms=int(rng.integers(1, 40)), min_length=int(rng.integers(260, 600)))
        tx = synth.make_transcriptome(n_tx, **kw)
        arrays = ref.ref_build_index(tx.sequences())
        expr = synth.make_expression(tx.n_transcripts, seed=int(rng.integers(1, 1000)))
        if not (expr > 0).any():
            continue
        cases = []
        for _ in range(3):
            paired = bool(rng.integers(0, 2))
            L = min(int(rng.choice([25, 26, 31, 33, 36, 50, 75, 100, 101, 150, 250])), int(tx.lengths.min()))
            sim = synth.ReadSimulator(
                tx, expr, L, int(rng.integers(L, max(L + 1, 500))), int(rng.integers(1, 80)),
                sub_rate=float(rng.choice([0.0, 0.002, 0.01, 0.03, 0.08])),
                n_rate=float(rng.choice([0.0, 0.001, 0.02])), random_rate_pct=int(rng.integers(0, 10)),
                paired=paired, seed=int(rng.integers(1, 1 << 30)))
            bases, _ = sim.generate(0, args.units)
            cases.append((dict(L=L, mu=sim.mu, sd=sim.sd, sub=sim.sub_thresh, n=sim.n_thresh, paired=paired,
                               seed=sim.seed), bases, sim.offsets(args.units), args.units, paired))
        for paired in (True, False):
            reads = adversarial.make_reads(tx, paired)
            b, offs = orc.pack_reads(reads)
            cases.append((dict(adversarial=True, paired=paired), b, offs,
                          len(reads) // 2 if paired else len(reads), paired))
        for what, bases, offs, n, paired in cases:
            runs += 1
            units += n
            try:
                tg.check_against_oracle(orc, arrays, bases, offs, n, paired, batches=int(rng.integers(1, 4)))
            except AssertionError as e:
                failures.append({'transcriptome': dict(n=n_tx, **kw), 'reads': what, 'error': str(e)[:300]})
    print(json.dumps({'seed': args.seed, 'seconds': args.seconds, 'runs': runs, 'units': units,
                      'mismatching_runs': len(failures), 'failures': failures[:5]}))


if __name__ == '__main__':
    main()
