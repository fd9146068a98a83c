"""`impute` on the GPU: mapping of every cell, first-round EM per cell and
the batched second-round EM, against the reference's outputs in tests/golden/impute_small.npz
(reads regenerated from the seeds in tests/impute_cases.py) and against the oracle."""
import numpy
import pytest

import impute_cases as ic
from conftest import GOLDEN, Golden
from seekmer_b200 import common, impute, mapper
from seekmer_b200.__main__ import main as cli_main
from test_impute_host import _first_round_results

pytestmark = pytest.mark.gpu

TOL = 1e-6  # north_star tolerance for the fp64 EM


@pytest.fixture(scope='module')
def gi():
    return Golden(GOLDEN / 'impute_small.npz')


def _index(gi, golden_synth):
    tab = golden_synth['transcripts'].copy()
    tab['gene_id'] = gi['gene_id']
    return common.KMerIndex(*golden_synth.index_arrays(), tab, None)


def test_impute_cells_matches_reference(gi, golden_synth, small_tx, tmp_path):
    index = _index(gi, golden_synth)
    feeders = [iter(ic.cell_batches(small_tx, c, batch=600)) for c in range(ic.N_CELLS)]
    results = mapper.map_multiple_samples(index, feeders, job_count=3)
    numpy.random.seed(1)   # the reference draws its KMeans initialisation from numpy's global RNG
    tpm, base, weight = impute.impute_cells(index, results, power=ic.POWER, output_path=tmp_path,
                                            return_stages=True)
    assert (results[0].fragment_length_counts == gi['fld']).all()
    assert all(r.fragment_length_counts is results[0].fragment_length_counts for r in results)
    assert numpy.allclose(base, gi['base'], rtol=TOL, atol=0)
    assert ((weight != 0) == (gi['weight'] != 0)).all()
    assert numpy.allclose(weight, gi['weight'], rtol=1e-9, atol=0)
    assert ((tpm == 0) == (gi['tpm'] == 0)).all()
    assert numpy.allclose(tpm, gi['tpm'], rtol=TOL, atol=0)
    assert (tmp_path / 'initial_gene_table.csv').read_bytes() == gi['gene_table_csv'].tobytes()
    index.release_device()


def test_batched_second_round_vs_oracle_and_chunking(gi, orc, monkeypatch):
    cells = _first_round_results(gi)
    impute._blend_mapping_results(cells, gi['weight'] ** ic.POWER)
    whole = impute._quantify_blended(cells)
    for i, cell in enumerate(cells):
        want = orc.quantify(gi['eff_lengths'], cell.class_map, cell.class_count)
        assert numpy.allclose(whole[i], want, rtol=TOL, atol=0)
    assert numpy.allclose(whole, gi['tpm'], rtol=TOL, atol=0)
    # the path impute_cells takes: per support group, unsupported blocks never built
    grouped = impute._quantify_weighted(_first_round_results(gi), gi['weight'] ** ic.POWER)
    assert numpy.allclose(grouped, gi['tpm'], rtol=TOL, atol=0)
    assert numpy.allclose(grouped, whole, rtol=1e-9, atol=0)
    by_group = impute._quantify_blended(cells, groups=impute._support_groups(gi['weight']))
    assert numpy.allclose(by_group, grouped, rtol=1e-12, atol=0)
    # three cells per device call: same answers
    monkeypatch.setattr(impute, '_EM_BATCH_BYTES', 3 * 8 * cells[0].class_count.size)
    assert (impute._quantify_blended(cells) == whole).all()
    # power=None: first round only (`impute.py:116-122`)
    with pytest.raises(ValueError):
        impute._quantify_blended(_first_round_results(gi))


def test_cli_impute_end_to_end(gi, golden_synth, small_tx, tmp_path):
    index = _index(gi, golden_synth)
    index.save(tmp_path / 'index.npz')
    paths = []
    for c in range(4):
        reads = [r for _, _, rs in ic.cell_batches(small_tx, c) for r in rs]
        for mate in (0, 1):
            p = tmp_path / ('cell%d_%d.fastq' % (c, mate + 1))
            with open(str(p), 'wb') as f:
                for i, r in enumerate(reads[mate::2]):
                    f.write(b'@r%d/%d\n' % (i, mate + 1) + r + b'\n+\n' + b'I' * len(r) + b'\n')
            paths.append(str(p))
    out = tmp_path / 'out'
    assert cli_main(['impute', '-p', '16', str(tmp_path / 'index.npz'), str(out)] + paths) == 0
    rows = [l.rstrip('\n').split(',') for l in open(str(out / 'tpm.csv'))]
    assert rows[0] == [''] + paths[0::2]
    assert [r[0] for r in rows[1:]] == [t.decode() for t in index.transcripts['transcript_id']]
    table = numpy.asarray([[float(v) for v in r[1:]] for r in rows[1:]]).T
    # the same four cells through the Python entry point
    feeders = [iter(ic.cell_batches(small_tx, c)) for c in range(4)]
    want = impute.impute_cells(index, mapper.map_multiple_samples(index, feeders), power=16)
    assert numpy.allclose(table, want, rtol=1e-9, atol=0)
    assert table.sum(axis=1) == pytest.approx(numpy.full(4, 1e6), rel=1e-9)
    assert (out / 'weight.csv').exists() and (out / 'initial_gene_table.csv').exists()
    index.release_device()
