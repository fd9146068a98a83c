"""Drop-in API tests on the GPU: the reference-facing Python entry points
(`mapper.map_reads`, `MapResult`, `infer.quantify`, `infer.run`, CLI) produce the reference's
results on the reference's own fixture."""
import json

import numpy
import pytest

from seekmer_b200 import _lib, common, infer, mapper
from seekmer_b200.__main__ import main as cli_main

pytestmark = pytest.mark.gpu


def make_index(g):
    return common.KMerIndex(*g.index_arrays(), g['transcripts'], None)


def write_fastq(path, names, reads):
    with open(str(path), 'wb') as f:
        for n, r in zip(names, reads):
            f.write(b'@' + n + b'\n' + r + b'\n+\n' + b'#' * len(r) + b'\n')


def test_map_reads_matches_reference_counter(golden_chr21):
    g = golden_chr21
    index = make_index(g)
    reads = [bytes(r) for r in g['reads']]
    names = [b'r%d' % i for i in range(21)]
    res = mapper.map_reads(index, iter([(21, names, reads)]), job_count=4)
    want = {}
    for t in g.tuples(''):
        want[t] = want.get(t, 0) + 1
    assert {k: v for k, v in res.counter.items() if v} == want
    assert list(k for k in res.counter if k) == list(dict.fromkeys(g.tuples('')))  # insertion order
    assert (res.fragment_length_counts == g['fld']).all()
    s = res.summarize()
    assert s.unaligned == 0 and s.aligned == 21 and s.total == 21  # test/test_mapper.py:76
    assert (s.class_map == g['class_map']).all()
    assert (s.class_count == g['class_count']).all()
    assert (s.effective_lengths == g['eff_lengths']).all()
    assert res.harmonic_mean_fragment_length == pytest.approx(float(g['harmonic_mean']), rel=1e-14)
    tpm = infer.quantify(s)
    assert numpy.allclose(tpm, g['tpm'], rtol=1e-6, atol=0)
    boots = [infer.quantify(s, x0=tpm, bootstrap=True) for _ in range(2)]  # test_infer.py:29-35
    assert all(b.shape == tpm.shape and numpy.isfinite(b).all() for b in boots)
    index.release_device()


def test_cli_infer_end_to_end(tmp_path, golden_chr21):
    g = golden_chr21
    index = make_index(g)
    index.save(tmp_path / 'index.npz')
    reads = [bytes(r) for r in g['reads']]
    names = [b'read%d/1' % i for i in range(21)]
    write_fastq(tmp_path / 'r_1.fastq', names, reads[0::2])
    write_fastq(tmp_path / 'r_2.fastq', names, reads[1::2])
    out = tmp_path / 'out'
    numpy.random.seed(7)
    assert cli_main(['infer', '-b', '3', '-m', str(tmp_path / 'index.npz'), str(out),
                     str(tmp_path / 'r_1.fastq'), str(tmp_path / 'r_2.fastq')]) == 0
    info = json.load(open(str(out / 'run_info.json')))
    assert info['n_processed'] == 21 and info['n_pseudoaligned'] == 21 and info['n_bootstraps'] == 3
    assert info['n_targets'] == len(g['transcripts'])
    rows = [l.rstrip('\n').split('\t') for l in open(str(out / 'abundance.tsv'))]
    assert rows[0] == ['target_id', 'length', 'eff_length', 'est_count', 'tpm']
    tpm = numpy.asarray([float(r[4]) for r in rows[1:]])
    assert numpy.allclose(tpm, g['tpm'], rtol=1e-5, atol=1e-12)
    est = numpy.asarray([float(r[3]) for r in rows[1:]])
    assert est.sum() == pytest.approx(21, rel=1e-4)
    lines = open(str(out / 'readmap.txt')).read().splitlines()
    assert len(lines) == 21
    ids = g['transcripts']['transcript_id']
    for line, t in zip(lines, g.tuples('')):
        assert line.split('\t')[1:] == [ids[i].decode() for i in t]
    assert (out / 'abundance.h5').exists() or (out / 'abundance.npz').exists()


def test_single_end_and_multiple_samples(golden_synth, small_tx):
    from conftest import SYNTH_CASES
    from seekmer_b200 import synth
    g = golden_synth
    index = make_index(g)
    kw = SYNTH_CASES['se75']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    feeders = [sim.batches(0, 3000, batch=700), sim.batches(0, 1500, batch=512)]
    results = mapper.map_multiple_samples(index, feeders, job_count=2)
    want = {}
    for t in g.tuples('se75_'):
        want[t] = want.get(t, 0) + 1
    assert dict(results[0].counter) == want
    assert (results[0].fragment_length_counts == g['se75_fld']).all()
    assert sum(results[1].counter.values()) == 1500


def test_device_image_round_trip(orc, golden_synth, small_tx, tmp_path):
    """The native GPU-layout index file: save the device image, load it back with
    plain copies (no relayout, no link probes) and map: the same class table, bit for bit."""
    from conftest import SYNTH_CASES
    from seekmer_b200 import synth
    g = golden_synth
    index = common.KMerIndex(*g.index_arrays(), g['transcripts'], None)
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **SYNTH_CASES['pe150'])
    want = mapper.map_reads(index, sim.batches(0, 4000, batch=1000))
    path = tmp_path / 'small.skmidx'
    index.save(path)
    index.release_device()
    again = common.KMerIndex.load(path)
    assert again.kmers is None and (again.transcripts == index.transcripts).all()
    got = mapper.map_reads(again, sim.batches(0, 4000, batch=1000))
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (got._table[k] == want._table[k]).all(), k
    assert dict(got.counter) == dict(want.counter)
    info_a, info_b = again.device_index(0).info(), common.KMerIndex(*g.index_arrays(), g['transcripts'], None).device_index(0).info()
    assert info_a == info_b
    # a truncated file and a foreign layout version are refused by the library as well
    raw = path.read_bytes()
    (tmp_path / 'cut.skmidx').write_bytes(raw[:len(raw) // 2])
    with pytest.raises(_lib.SeekmerCudaError, match='truncated'):
        _lib.DeviceIndex.load(tmp_path / 'cut.skmidx')
    bumped = bytearray(raw)
    bumped[8] = 99
    (tmp_path / 'v99.skmidx').write_bytes(bytes(bumped))
    with pytest.raises(_lib.SeekmerCudaError, match='invalid index version'):
        _lib.DeviceIndex.load(tmp_path / 'v99.skmidx')
    again.release_device()
