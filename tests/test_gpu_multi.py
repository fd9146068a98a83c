"""Multi-GPU behind the drop-in API (`job_count` = GPUs): mapping, sample round-robin and
bootstrap sharding over two devices give exactly what one device gives.  Needs two GPUs
(`gpurun --gpus 2`); skipped on a one-GPU box."""
import numpy
import pytest

from conftest import SYNTH_CASES
from seekmer_b200 import _lib, common, infer, mapper, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def two_gpus():
    if _lib.device_count() < 2:
        pytest.skip('needs two CUDA devices')


def _index(g):
    return common.KMerIndex(*g.index_arrays(), g['transcripts'], None)


def _same(a, b):
    ta, tb = a._table, b._table
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        if k == 'first_unit':
            continue  # global unit indices differ between the paths; the order they induce must not
        assert (ta[k] == tb[k]).all(), k
    assert ta['unaligned'] == tb['unaligned'] and ta['aligned'] == tb['aligned']
    assert dict(a.counter) == dict(b.counter)
    assert list(a.counter) == list(b.counter)  # Counter insertion order = first-seen class order
    assert (a.fragment_length_counts == b.fragment_length_counts).all()


def test_map_reads_over_two_gpus_equals_one(two_gpus, golden_synth, small_tx, tmp_path):
    g = golden_synth
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **SYNTH_CASES['pe100'])
    one = mapper.map_reads(_index(g), sim.batches(0, 6000, batch=500), job_count=1)
    two = mapper.map_reads(_index(g), sim.batches(0, 6000, batch=500), job_count=2)
    _same(one, two)
    s1, s2 = one.summarize(), two.summarize()
    assert (s1.class_map == s2.class_map).all() and (s1.class_count == s2.class_count).all()
    t1, t2 = infer.quantify(s1), infer.quantify(s2)
    assert (t1 == t2).all()
    # FASTQ file groups: three pairs of files dealt to two devices
    paths = []
    for k in range(3):
        bases, _ = sim.generate(k * 2000, 2000)
        reads = bases.reshape(2000, 2, 100)
        for mate in range(2):
            p = tmp_path / ('s%d_%d.fq' % (k, mate + 1))
            with open(p, 'wb') as f:
                for i in range(2000):
                    f.write(b'@r%d\n' % i + reads[i, mate].tobytes() + b'\n+\n' + b'I' * 100 + b'\n')
            paths.append(p)
    f1 = mapper.map_reads(_index(g), common.feed_pair_ended_reads(*paths), job_count=1)
    f2 = mapper.map_reads(_index(g), common.feed_pair_ended_reads(*paths), job_count=2)
    _same(f1, f2)
    _same(f1, one)


def test_samples_and_bootstraps_over_two_gpus(two_gpus, golden_synth, small_tx):
    g = golden_synth
    sims = [synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=20 + k),
                                **dict(SYNTH_CASES['se75'], seed=50 + k)) for k in range(5)]
    one = mapper.map_multiple_samples(_index(g), [s.batches(0, 1500, batch=512) for s in sims], job_count=1)
    two = mapper.map_multiple_samples(_index(g), [s.batches(0, 1500, batch=512) for s in sims], job_count=2)
    for a, b in zip(one, two):
        _same(a, b)
        assert (infer.quantify(a.summarize()) == infer.quantify(b.summarize())).all()
    # samples mapped on device 1 carry a plan on device 1
    assert {r._plan.device for r in two if r._plan is not None} == {0, 1}
    s = one[0].summarize()
    tpm = infer.quantify(s)
    b1, i1 = infer.quantify_bootstraps(s, tpm, 11, seed=99, return_iters=True, devices=[0])
    b2, i2 = infer.quantify_bootstraps(s, tpm, 11, seed=99, return_iters=True, devices=[0, 1])
    assert (i1 == i2).all() and all((x == y).all() for x, y in zip(b1, b2))
