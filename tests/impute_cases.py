"""The synthetic single-cell input of tests/golden/impute_small.npz: shared by the generator
(tests/golden/make_golden_impute.py, which ran the reference on it) and the GPU tests (which
regenerate the same reads from the seeds)."""
import numpy

from seekmer_b200 import synth

N_CELLS = 8
PAIRS_PER_CELL = 1500
POWER = 16
SIM = dict(read_length=100, frag_mean=250, frag_sd=30, sub_rate=0.01, paired=True)


def gene_ids(n_transcripts):
    """Genes of three transcripts; every tenth gene has the empty id the reference masks out."""
    ids = []
    for i in range(n_transcripts):
        g = i // 3
        ids.append(b'' if g % 10 == 9 else b'G%05d' % g)
    return ids


def cell_expression(n_transcripts, cell):
    """Two programmes (even / odd cells) with a little per-cell jitter."""
    base = synth.make_expression(n_transcripts, seed=3 if cell % 2 == 0 else 5)
    rng = numpy.random.Generator(numpy.random.PCG64(100 + cell))
    return base * rng.uniform(0.8, 1.25, size=n_transcripts)


def cell_batches(tx, cell, batch=1024):
    sim = synth.ReadSimulator(tx, cell_expression(tx.n_transcripts, cell), seed=200 + cell, **SIM)
    return list(sim.batches(0, PAIRS_PER_CELL, batch=batch))
