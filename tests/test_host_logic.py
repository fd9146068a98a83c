"""CPU-side tests: C-ABI surface, host mirrors of the reference API, synthetic generators."""
import ctypes
import re

import numpy
import pytest

from conftest import ROOT
from seekmer_b200 import _lib, common, mapper, synth


def test_library_exports_every_declared_symbol():
    header = (ROOT / 'include' / 'seekmer_b200.h').read_text()
    declared = set(re.findall(r'\b(skm_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert b'sm_100a' in _lib.load().skm_version()


def test_no_cpu_fallback_without_device(golden_chr21):
    if _lib.device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(_lib.SeekmerCudaError, match='no CPU fallback'):
        _lib.DeviceIndex(*golden_chr21.index_arrays(), 42)
    idx = common.KMerIndex(*golden_chr21.index_arrays(), golden_chr21['transcripts'], None)
    with pytest.raises(_lib.SeekmerCudaError):
        mapper.map_reads(idx, iter([(1, [b'a'], [b'A' * 50])]))
    mr = mapper.MapResult(idx)
    with pytest.raises(_lib.SeekmerCudaError):
        mr.effective_lengths


def test_product_never_imports_the_oracle():
    for path in (ROOT / 'seekmer_b200').rglob('*'):
        if path.suffix in ('.py', '.cu', '.cuh') and path.is_file():
            text = path.read_text()
            assert 'import oracle' not in text and 'from oracle' not in text, path
            assert 'seekmer_oracle' not in text, path


def write_fastq(path, names, reads):
    with open(str(path), 'wb') as f:
        for n, r in zip(names, reads):
            f.write(b'@' + n + b'\n' + r + b'\n+\n' + b'#' * len(r) + b'\n')


def test_feeders_concatenate_files_into_one_sample(tmp_path):
    # mirrors test/test_mapper.py:20-68
    names = [b'r%d desc' % i for i in range(21)]
    r1 = [b'ACGTN' * 20 for _ in range(21)]
    r2 = [b'acgtn' * 20 for _ in range(21)]
    write_fastq(tmp_path / 'a_1.fastq', names, r1)
    write_fastq(tmp_path / 'a_2.fastq', names, r2)
    batches = list(common.feed_single_ended_reads(tmp_path / 'a_1.fastq'))
    assert len(batches) == 1 and batches[0][0] == 21 and len(batches[0][2]) == 21
    batches = list(common.feed_single_ended_reads(tmp_path / 'a_1.fastq', tmp_path / 'a_1.fastq'))
    assert len(batches) == 1 and batches[0][0] == 42
    n, nm, rd = next(common.feed_pair_ended_reads(tmp_path / 'a_1.fastq', tmp_path / 'a_2.fastq'))
    assert n == 21 and len(rd) == 42 and rd[0] == r1[0] and rd[1] == r2[0] and nm[0] == names[0]
    n, _, rd = next(common.feed_pair_ended_reads(*[tmp_path / 'a_1.fastq', tmp_path / 'a_2.fastq'] * 2))
    assert n == 42 and len(rd) == 84
    with pytest.raises(ValueError):
        list(common.feed_pair_ended_reads(tmp_path / 'a_1.fastq'))


def test_feeder_batches_and_gzip(tmp_path, monkeypatch):
    import gzip
    monkeypatch.setattr(common, 'BUFFER_SIZE', 8)
    names = [b'x%d' % i for i in range(21)]
    reads = [b'ACGT' * 10] * 21
    write_fastq(tmp_path / 's.fastq', names, reads)
    with open(str(tmp_path / 's.fastq'), 'rb') as f, gzip.open(str(tmp_path / 's.fastq.gz'), 'wb') as z:
        z.write(f.read())
    for p in (tmp_path / 's.fastq', tmp_path / 's.fastq.gz'):
        sizes = [b[0] for b in common.feed_single_ended_reads(p)]
        assert sizes == [8, 8, 5]


def test_map_result_model_and_summarize(golden_chr21):
    g = golden_chr21
    idx = common.KMerIndex(*g.index_arrays(), g['transcripts'], None)
    mr = mapper.MapResult(idx)
    tuples = g.tuples('')
    mr.update([b'n'] * len(tuples), tuples + [()])
    assert mr.counter[()] == 1
    mr.merge_fragment_lengths(g['fld'])
    assert mr.harmonic_mean_fragment_length == pytest.approx(float(g['harmonic_mean']), rel=1e-14)
    # summarize without touching the device: patch effective_lengths
    mapper.MapResult.effective_lengths, saved = property(lambda self: g['eff_lengths']), mapper.MapResult.effective_lengths
    try:
        s = mr.summarize()
    finally:
        mapper.MapResult.effective_lengths = saved
    assert (s.class_map == g['class_map']).all() and s.class_map.dtype == numpy.int64
    assert (s.class_count == g['class_count']).all()
    assert (s.aligned, s.unaligned, s.total) == (21, 1, 22)
    assert mr.counter[()] == 1
    mr.clear()
    assert not mr.counter
    assert mapper.MAX_FRAGMENT_LENGTH == 2000


def test_index_npz_roundtrip_and_version_check(tmp_path, golden_chr21):
    g = golden_chr21
    idx = common.KMerIndex(*g.index_arrays(), g['transcripts'], None)
    idx.save(tmp_path / 'i.npz')
    back = common.KMerIndex.load(tmp_path / 'i.npz')
    for name in ('kmers', 'contigs', 'sequences', 'targets', 'transcripts'):
        assert (numpy.asarray(getattr(back, name)) == numpy.asarray(getattr(idx, name))).all()
    z = dict(numpy.load(str(tmp_path / 'i.npz')))
    z['seekmer_version'] = numpy.asarray('1999.0.0')
    numpy.savez(str(tmp_path / 'bad.npz'), **z)
    with pytest.raises(RuntimeError, match='invalid index version'):
        common.KMerIndex.load(tmp_path / 'bad.npz')


def test_synthetic_generators_are_deterministic_and_sliceable():
    tx = synth.make_transcriptome(120, seed=4)
    tx2 = synth.make_transcriptome(120, seed=4)
    assert (tx.codes == tx2.codes).all() and (tx.offsets == tx2.offsets).all()
    assert tx.lengths.min() >= 400 and tx.strand_flipped.any() and not tx.strand_flipped.all()
    expr = synth.make_expression(120)
    sim = synth.ReadSimulator(tx, expr, 100, 250, 30)
    a, truth = sim.generate(0, 500)
    b, _ = sim.generate(200, 100)
    assert (a.reshape(500, 200)[200:300] == b.reshape(100, 200)).all()
    assert set(numpy.unique(a)) <= set(b'ACGTN')
    # error-free, non-random pairs reproduce the transcript
    clean = synth.ReadSimulator(tx, expr, 100, 250, 30, sub_rate=0, n_rate=0, random_rate_pct=0)
    r, t = clean.generate(0, 50)
    r = r.reshape(50, 200)
    for i in range(50):
        s = tx.sequence(int(t['transcript'][i]))
        st, fr = int(t['start'][i]), int(t['fragment'][i])
        m1, m2 = s[st:st + 100], synth.reverse_complement_ascii(s[st + fr - 100:st + fr])
        got = (r[i, :100].tobytes(), r[i, 100:].tobytes())
        assert got == ((m2, m1) if t['swap'][i] else (m1, m2))
    assert 200 < numpy.mean(t['fragment']) < 300


def test_abundance_tsv_is_byte_identical_to_the_pandas_writer(tmp_path):
    """`infer._output_abundance_table` against what the reference writes (`infer.py:219-230`:
    a DataFrame through `to_csv(sep='\\t', index=False, float_format='%g')`), byte for byte,
    with values that stress '%g': tiny, huge, zero, NaN, lengths of a million bases and more."""
    import pandas
    from seekmer_b200 import infer
    rng = numpy.random.Generator(numpy.random.PCG64(3))
    n = 500
    tr = numpy.zeros(n, dtype=[('transcript_id', 'S12'), ('gene_id', 'S12'), ('length', 'f8')])
    tr['transcript_id'] = [('TX%06d' % i).encode() for i in range(n)]
    tr['length'] = rng.integers(30, 200000, size=n)
    tr['length'][:4] = [1000000, 1234567, 25, 99999999]

    class Index:
        transcripts = tr

    class Results:
        effective_lengths = numpy.maximum(tr['length'] - 187.123456, 1.0)

    est = numpy.exp(rng.normal(0, 6, size=n))
    tpm = numpy.exp(rng.normal(0, 6, size=n))
    est[:6] = [0.0, 1e-300, 123456789.0, 0.001, 1e6, numpy.nan]
    tpm[:6] = [0.0, 5e-324, 999999.5, 1e-5, 100000.0, numpy.nan]
    for lengths in (tr['length'], tr['length'].astype('i8')):
        Index.transcripts = tr if lengths.dtype.kind == 'f' else tr.astype(
            [('transcript_id', 'S12'), ('gene_id', 'S12'), ('length', 'i8')])
        infer._output_abundance_table(tmp_path, Index, Results, est, tpm)
        table = pandas.DataFrame()
        table['target_id'] = [i.decode() for i in Index.transcripts['transcript_id']]
        table['length'] = Index.transcripts['length']
        table['eff_length'] = Results.effective_lengths.astype('f4')
        table['est_count'] = est
        table['tpm'] = tpm
        table.to_csv(str(tmp_path / 'pandas.tsv'), sep='\t', index=False, float_format='%g')
        ours = (tmp_path / 'abundance.tsv').read_bytes()
        want = (tmp_path / 'pandas.tsv').read_bytes()
        assert ours == want


def test_device_image_of_a_wrong_version_is_refused(tmp_path):
    """`KMerIndex.load` of a native GPU-layout index file checks the layout version before
    anything is uploaded, like `_common.pyx:303-304` ("invalid index version.")."""
    import io
    buf = io.BytesIO()
    numpy.savez(buf, seekmer_version=numpy.asarray('2019.0.0'), transcripts=numpy.zeros(3), exons=numpy.zeros(0))
    trailer = buf.getvalue()

    def image(version, trailer_bytes=trailer, magic=b'SKMB200\0'):
        head = common._IMAGE_HEADER.pack(magic, version, 25, 4, 16, 128, 8, 1024, 10, 5, 100, 7, 3, 2, 0, 0, 0, 0,
                                         len(trailer_bytes), 0)
        return head + trailer_bytes

    good = tmp_path / 'ok.skmidx'
    good.write_bytes(image(common._IMAGE_VERSION))
    index = common.KMerIndex.load(good)  # host side only: nothing is uploaded before a mapper asks
    assert index.kmers is None and len(index.transcripts) == 3 and index._image_path == good
    bad = tmp_path / 'bad.skmidx'
    bad.write_bytes(image(common._IMAGE_VERSION + 1))
    with pytest.raises(RuntimeError, match='invalid index version'):
        common.KMerIndex.load(bad)
    other = io.BytesIO()
    numpy.savez(other, seekmer_version=numpy.asarray('2018.0.0'), transcripts=numpy.zeros(3), exons=numpy.zeros(0))
    bad.write_bytes(image(common._IMAGE_VERSION, other.getvalue()))
    with pytest.raises(RuntimeError, match='invalid index version'):
        common.KMerIndex.load(bad)
    bad.write_bytes(image(common._IMAGE_VERSION, magic=b'NOTANIDX'))
    with pytest.raises(RuntimeError, match='not a seekmer_b200 device image'):
        common.KMerIndex.load(bad)
