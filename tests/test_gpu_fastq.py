"""Device-side FASTQ ingestion (`skm_map_fastq` behind `common.feed_*_reads`): the
raw-text path gives exactly what the line-by-line feeder protocol gives, which the other GPU
tests pin to the oracle and the golden vectors."""
import gzip

import numpy
import pytest

from conftest import N_GOLDEN_UNITS, SYNTH_CASES
from seekmer_b200 import common, mapper, synth

pytestmark = pytest.mark.gpu


def make_index(g):
    return common.KMerIndex(*g.index_arrays(), g['transcripts'], None)


def reads_of(sim, n, paired):
    bases, _ = sim.generate(0, n)
    offs = sim.offsets(n)
    raw = bases.tobytes()
    per = 2 if paired else 1
    return [raw[offs[i]:offs[i + 1]] for i in range(per * n)]


def write_fastq(path, reads, eol=b'\n', last_newline=True, opener=open):
    with opener(str(path), 'wb') as f:
        for i, r in enumerate(reads):
            rec = b'@read' + str(i).encode() + b' some/description' + eol + r + eol + b'+' + eol + b'I' * len(r)
            f.write(rec + (eol if (last_newline or i + 1 < len(reads)) else b''))


def want_counts(g, case):
    want = {}
    for t in g.tuples(case + '_'):
        want[t] = want.get(t, 0) + 1
    return want


@pytest.mark.parametrize('case', ['pe100', 'se75'])
@pytest.mark.parametrize('variant', ['plain', 'crlf_nofinalnewline', 'gz', 'tiny_chunks'])
def test_fastq_text_path_equals_golden(tmp_path, monkeypatch, golden_synth, small_tx, case, variant):
    g = golden_synth
    kw = SYNTH_CASES[case]
    paired = kw['paired']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    reads = reads_of(sim, N_GOLDEN_UNITS, paired)
    eol = b'\r\n' if variant.startswith('crlf') else b'\n'
    opener, suffix = (gzip.open, '.fastq.gz') if variant == 'gz' else (open, '.fastq')
    last_nl = variant != 'crlf_nofinalnewline'
    if variant == 'tiny_chunks':  # many chunks, records straddling every boundary
        monkeypatch.setattr(mapper, 'FASTQ_CHUNK_BYTES', 70000)
    if paired:
        p1, p2 = tmp_path / ('r_1' + suffix), tmp_path / ('r_2' + suffix)
        write_fastq(p1, reads[0::2], eol, last_nl, opener)
        write_fastq(p2, reads[1::2], eol, last_nl, opener)
        feeder = common.feed_pair_ended_reads(p1, p2)
    else:
        p1 = tmp_path / ('r' + suffix)
        write_fastq(p1, reads, eol, last_nl, opener)
        feeder = common.feed_single_ended_reads(p1)
    assert isinstance(feeder, common.FastqSource)
    index = make_index(g)
    res = mapper.map_reads(index, feeder)
    assert {k: v for k, v in res.counter.items() if v} == want_counts(g, case)
    assert (res.fragment_length_counts == g[case + '_fld']).all()
    # first-seen class order survives chunking
    assert [k for k in res.counter if k] == list(dict.fromkeys(t for t in g.tuples(case + '_') if t))
    index.release_device()


def test_feeder_protocol_still_iterates(tmp_path, golden_synth, small_tx):
    """The returned object still yields the reference's (count, names, reads) batches."""
    kw = SYNTH_CASES['pe100']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    reads = reads_of(sim, 50, True)
    p1, p2 = tmp_path / 'a_1.fastq', tmp_path / 'a_2.fastq'
    write_fastq(p1, reads[0::2])
    write_fastq(p2, reads[1::2])
    batches = list(common.feed_pair_ended_reads(p1, p2))
    assert len(batches) == 1 and batches[0][0] == 50
    assert batches[0][2] == reads
    assert batches[0][1][0] == b'read0 some/description'


def test_partial_trailing_record_and_uneven_mates(tmp_path, golden_synth, small_tx):
    """A last record cut after its sequence line still counts (common.py:137-138); surplus reads
    of the longer mate file are ignored (zip)."""
    g = golden_synth
    kw = SYNTH_CASES['pe100']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    reads = reads_of(sim, 40, True)
    p1, p2 = tmp_path / 'b_1.fastq', tmp_path / 'b_2.fastq'
    write_fastq(p1, reads[0::2])
    write_fastq(p2, reads[1::2][:30])
    with open(str(p2), 'ab') as f:
        f.write(b'@cut\n' + reads[61] + b'\n')  # record 31 of mate 2: header + sequence only
    index = make_index(g)
    fast = mapper.map_reads(index, common.feed_pair_ended_reads(p1, p2))
    slow = mapper.map_reads(index, iter(common.feed_pair_ended_reads(p1, p2)))
    assert sum(fast.counter.values()) == 31
    assert dict(fast.counter) == dict(slow.counter)
    assert (fast.fragment_length_counts == slow.fragment_length_counts).all()
    index.release_device()
