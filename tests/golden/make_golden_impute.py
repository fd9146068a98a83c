#!/usr/bin/env python3
"""Golden vectors for the single-cell `impute` workflow, made by RUNNING THE
REFERENCE (`/root/reference/seekmer/impute.py` + `infer.py` + `mapper.py` over the compiled
natives in oracle/_ref).  Build container only:
    python tests/golden/make_golden_impute.py

Input: the 60-transcript synthetic transcriptome of synthetic_small.npz (its index arrays are
reused), grouped into genes of three transcripts (every tenth gene id left empty, which the
reference masks out), and N_CELLS cells of two expression programmes.  The tests regenerate
the reads from the same seeds (tests/impute_cases.py).

Stored, all outputs of unmodified reference code:
  fld            merged fragment length counts         impute._merge_fragment_lengths
  base           first-round TPM, cell x transcript     infer.quantify per cell
  weight         filtered cell-cell weights (power 1)   impute._calculate_cell_weights
  blended_count  second-round class counts, cell 0 and the last cell
  tpm            second-round TPM, cell x transcript    infer.quantify on the blended results
The reference's KMeans is unseeded; it draws from numpy's global RNG, seeded here.
"""
import pathlib
import sys
import tempfile
import warnings

import numpy

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_harness as rh  # noqa: E402
from seekmer_b200 import synth  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent

sys.path.insert(0, str(ROOT / 'tests'))
from impute_cases import N_CELLS, POWER, cell_batches, gene_ids  # noqa: E402


def main():
    pkg = rh.load_ref()
    from seekmer import mapper as ref_mapper, infer as ref_infer, impute as ref_impute
    g = numpy.load(str(OUT / 'synthetic_small.npz'))
    tx = synth.make_transcriptome(60, seed=7)
    tab = g['transcripts'].copy()
    tab['gene_id'] = gene_ids(len(tab))
    index = pkg._common.KMerIndex(g['kmers'], g['contigs'], g['sequences'], g['targets'], tab, None)

    feeders = [iter(cell_batches(tx, c)) for c in range(N_CELLS)]
    results = ref_mapper.map_multiple_samples(index, feeders, job_count=1, debug=True)
    ref_impute._merge_fragment_lengths(results)
    summarized = [r.summarize() for r in results]
    first_counts = [s.class_count.copy() for s in summarized]
    first_maps = [s.class_map.copy() for s in summarized]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        base = numpy.asarray([ref_infer.quantify(s) for s in summarized])
        numpy.random.seed(1)
        with tempfile.TemporaryDirectory() as tmp:
            weight = ref_impute._calculate_cell_weights(index, base, pathlib.Path(tmp))
            gene_table = open(pathlib.Path(tmp) / 'initial_gene_table.csv').read()
        powered = weight ** POWER
        ref_impute._blend_mapping_results(summarized, powered)
        tpm = numpy.asarray([ref_infer.quantify(s) for s in summarized])
    ptr = numpy.zeros(N_CELLS + 1, dtype='i8')
    numpy.cumsum([c.size for c in first_counts], out=ptr[1:])
    numpy.savez_compressed(
        str(OUT / 'impute_small.npz'),
        gene_id=tab['gene_id'], fld=summarized[0].fragment_length_frequencies,
        eff_lengths=summarized[0].effective_lengths,
        class_ptr=ptr, class_count=numpy.concatenate(first_counts),
        class_nnz=numpy.asarray([m.shape[1] for m in first_maps]),
        class_map=numpy.concatenate(first_maps, axis=1),
        base=base, weight=weight, power=numpy.asarray(POWER),
        blended_map=summarized[0].class_map,
        blended_count_first=summarized[0].class_count, blended_count_last=summarized[-1].class_count,
        tpm=tpm, gene_table_csv=numpy.frombuffer(gene_table.encode(), dtype='u1'))
    print('cells %d, classes per cell %s' % (N_CELLS, [c.size for c in first_counts]))
    print('weights kept per row:', (weight != 0).sum(axis=1))
    print(numpy.round(weight, 3))
    print('tpm sums', tpm.sum(axis=1)[:3], 'nonzero', (tpm > 0).sum(axis=1))


if __name__ == '__main__':
    main()
