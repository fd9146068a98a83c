"""GPU parity tests for the mapping path: CUDA (through the C ABI) vs the CPU oracle and the
golden vectors.  Bit-exact: per-unit ordered id tuples, raw fragment lengths, FLD, class
counts, first-seen class order."""
import numpy
import pytest

import adversarial
from conftest import N_GOLDEN_UNITS, SYNTH_CASES
from seekmer_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def gpu_map(arrays, bases, offsets, n_units, paired, n_tx=0, batches=1, **kw):
    """Map through the C ABI with host buffers; returns per-unit tuples, lengths, export."""
    ix = _lib.DeviceIndex(*arrays, n_tx)
    mp = _lib.DeviceMapper(ix, **kw)
    per = 2 if paired else 1
    cls, lens = [], []
    bounds = numpy.linspace(0, n_units, batches + 1).astype('i8')
    for b in range(batches):
        u0, u1 = int(bounds[b]), int(bounds[b + 1])
        if u1 == u0:
            continue
        o = offsets[u0 * per:u1 * per + 1]
        c, l = mp.map_batch(bases[o[0]:o[-1]], o - o[0], u1 - u0, paired, first_unit=u0,
                            max_len=int((o[1:] - o[:-1]).max()), per_read=True)
        cls.append(c)
        lens.append(l)
    table = mp.export(with_slots=True)
    off, ids = table['key_offsets'].tolist(), table['key_ids'].tolist()
    by_slot = {s: tuple(ids[off[i]:off[i + 1]]) for i, s in enumerate(table['slots'].tolist())}
    cls = numpy.concatenate(cls)
    tuples = [by_slot[s] if s >= 0 else () for s in cls.tolist()]
    mp.close()
    ix.close()
    return tuples, numpy.concatenate(lens), table


def table_dict(table):
    off, ids, cnt = table['key_offsets'].tolist(), table['key_ids'].tolist(), table['counts'].tolist()
    d = {tuple(ids[off[i]:off[i + 1]]): cnt[i] for i in range(len(cnt))}
    if table['unaligned']:
        d[()] = table['unaligned']
    return d


def check_against_oracle(orc, arrays, bases, offsets, n_units, paired, **kw):
    oidx = orc.OracleIndex(*arrays)
    want = orc.map_batch(oidx, bases, offsets, paired)
    tuples, lens, table = gpu_map(arrays, bases, offsets, n_units, paired, **kw)
    wt = want.tuples()
    bad = [i for i in range(n_units) if tuples[i] != wt[i]]
    assert not bad, 'first mismatch at unit %d: gpu %r oracle %r' % (bad[0], tuples[bad[0]], wt[bad[0]])
    assert (lens == want.length).all()
    assert (table['fld'] == want.fld).all()
    assert table_dict(table) == orc.tally_dict(want.ptr, want.ids)
    # first-seen order == Counter insertion order at job_count=1
    cls_ptr, cls_ids, cls_count, una = orc.tally(want.ptr, want.ids)
    assert (table['key_offsets'] == cls_ptr).all()
    assert (table['key_ids'] == cls_ids).all()
    assert (table['counts'] == cls_count).all()
    assert table['unaligned'] == una
    assert table['aligned'] == int(cls_count.sum())
    return table


def test_map_kmers_parity(orc, golden_chr21):
    arrays = golden_chr21.index_arrays()
    oidx = orc.OracleIndex(*arrays)
    ix = _lib.DeviceIndex(*arrays, 42)
    info = ix.info()
    kmers = arrays[0]
    occ = kmers['kmer'][kmers['kmer'] != numpy.uint64(0xFFFFFFFFFFFFFFFF)]
    assert info['n_kmers'] == occ.shape[0]
    assert info['table_slots'] >= 4 * occ.shape[0]
    rng = numpy.random.Generator(numpy.random.PCG64(1))
    rc = numpy.asarray([orc.reverse_complement(int(k)) for k in occ[:4000]], dtype='u8')
    miss = rng.integers(0, 1 << 50, size=4000, dtype='u8')
    q = numpy.concatenate([occ, rc, miss])
    e, o = ix.map_kmers(q)
    # every stored k-mer maps to its stored position, forward
    slot_of = {int(k): i for i, k in enumerate(kmers['kmer'])}
    for j in rng.choice(q.shape[0], size=6000, replace=False):
        assert (int(e[j]), int(o[j])) == oidx.map_kmer(int(q[j]))
    pos = numpy.asarray([slot_of[int(k)] for k in occ])
    assert (e[:occ.shape[0]] == kmers['entry'][pos]).all()
    assert (o[:occ.shape[0]] == kmers['offset'][pos]).all()
    ix.close()


def test_chr21_fixture(orc, golden_chr21):
    g = golden_chr21
    reads = [bytes(r) for r in g['reads']]
    bases, offs = orc.pack_reads(reads)
    tuples, lens, table = gpu_map(g.index_arrays(), bases, offs, 21, True)
    assert tuples == g.tuples('')
    assert (table['fld'] == g['fld']).all()
    assert table['unaligned'] == 0  # test/test_mapper.py:76
    assert (orc.class_map_from_csr(table['key_offsets'], table['key_ids']) == g['class_map']).all()
    assert (table['counts'] == g['class_count']).all()


@pytest.mark.parametrize('case', sorted(SYNTH_CASES))
def test_synthetic_golden(orc, golden_synth, small_tx, case):
    g = golden_synth
    kw = SYNTH_CASES[case]
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    bases, _ = sim.generate(0, N_GOLDEN_UNITS)
    offs = sim.offsets(N_GOLDEN_UNITS)
    tuples, lens, table = gpu_map(g.index_arrays(), bases, offs, N_GOLDEN_UNITS, kw['paired'], batches=3)
    assert tuples == g.tuples(case + '_')
    assert (table['fld'] == g[case + '_fld']).all()
    assert (orc.class_map_from_csr(table['key_offsets'], table['key_ids']) == g[case + '_class_map']).all()
    assert (table['counts'] == g[case + '_class_count']).all()


@pytest.mark.parametrize('paired', [True, False])
def test_adversarial_ragged(orc, golden_synth, small_tx, paired):
    g = golden_synth
    reads = adversarial.make_reads(small_tx, paired)
    bases, offs = orc.pack_reads(reads)
    n = len(reads) // 2 if paired else len(reads)
    tuples, lens, table = gpu_map(g.index_arrays(), bases, offs, n, paired)
    key = 'adv_pe_' if paired else 'adv_se_'
    assert tuples == g.tuples(key)
    assert (table['fld'] == g[key + 'fld']).all()


@pytest.mark.parametrize('L,mu,sd,paired,sub', [(100, 250, 30, True, 0.01), (150, 350, 50, True, 0.01),
                                                 (75, 250, 30, False, 0.02), (25, 250, 30, True, 0.0),
                                                 (251, 400, 60, True, 0.03)])
def test_medium_vs_oracle(orc, medium, L, mu, sd, paired, sub):
    tx, arrays = medium
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), L, mu, sd, sub_rate=sub,
                              paired=paired, seed=21)
    n = 40000
    bases, _ = sim.generate(0, n)
    check_against_oracle(orc, arrays, bases, sim.offsets(n), n, paired, batches=2)


def test_arbitrary_bytes_vs_oracle(orc, medium):
    """Every byte value can turn up in a read: anything but ACGTacgt encodes as 'A' and anything
    but upper-case ACGT is a wildcard for the edge checks (_kmer.pxd:253-273, _mapper.pyx:500-501)."""
    tx, arrays = medium
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), 100, 250, 30, sub_rate=0.005, seed=31)
    n = 30000
    bases, _ = sim.generate(0, n)
    bases = bases.copy()
    rng = numpy.random.Generator(numpy.random.PCG64(17))
    hit = rng.random(bases.shape[0]) < 0.01
    bases[hit] = rng.integers(0, 256, size=int(hit.sum()), dtype='u1')
    check_against_oracle(orc, arrays, bases, sim.offsets(n), n, True, batches=2)


def test_long_target_lists_spill_to_arena(orc, ref):
    """Contigs shared by more than LIST_CAP=16 transcripts exercise the arena path."""
    tx = synth.make_transcriptome(400, seed=5, max_isoforms=40, mean_exons=8)
    arrays = ref.ref_build_index(tx.sequences())
    assert int(numpy.asarray(arrays[1])['target_count'].max()) > 16
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), 100, 250, 30, seed=3)
    n = 20000
    bases, _ = sim.generate(0, n)
    check_against_oracle(orc, arrays, bases, sim.offsets(n), n, True)


def test_device_buffers_match_host_buffers(golden_synth, small_tx):
    import torch
    g = golden_synth
    kw = SYNTH_CASES['pe100']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    n = 3000
    bases, _ = sim.generate(0, n)
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    a = _lib.DeviceMapper(ix)
    b = _lib.DeviceMapper(ix)
    ca, la = a.map_batch(bases, None, n, True, fixed_len=100, per_read=True)
    d_bases = torch.from_numpy(bases).cuda()
    cb, lb = b.map_batch(d_bases, None, n, True, fixed_len=100, per_read=True)
    torch.cuda.synchronize()
    ta, tb = a.export(), b.export()
    assert (la == lb.cpu().numpy()).all()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (ta[k] == tb[k]).all()
    assert ta['unaligned'] == tb['unaligned']
    # reset clears everything
    a.reset()
    s = a.sizes()
    assert s['n_classes'] == 0 and s['aligned'] == 0 and s['unaligned'] == 0
    assert a.export()['fld'].sum() == 0


def test_capacity_errors_are_loud(golden_synth, small_tx):
    g = golden_synth
    kw = SYNTH_CASES['pe100']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    bases, _ = sim.generate(0, 3000)
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    mp = _lib.DeviceMapper(ix, class_capacity=8, id_capacity=16)
    with pytest.raises(_lib.SeekmerCudaError, match='capacity'):
        mp.map_batch(bases, None, 3000, True, fixed_len=100)


def test_reads_shorter_than_k_are_unaligned_units(orc, golden_synth, small_tx):
    """Trimmed FASTQ holds reads shorter than k=25 (and empty ones).  Undefined in the reference;
    defined here (oracle and CUDA alike): the unit is unaligned with span length 0, the run goes
    on, and the reads are counted."""
    g = golden_synth
    arrays = g.index_arrays()
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **SYNTH_CASES['pe100'])
    n = 2000
    bases, _ = sim.generate(0, n)
    reads = [bases[i * 100:(i + 1) * 100].tobytes() for i in range(2 * n)]
    rng = numpy.random.Generator(numpy.random.PCG64(5))
    cut = rng.choice(2 * n, size=300, replace=False)
    for j, i in enumerate(cut):
        reads[i] = reads[i][:int(rng.integers(0, 25))] if j % 3 else b''
    for paired in (True, False):
        flat = numpy.frombuffer(b''.join(reads), dtype='u1')
        offs = numpy.zeros(2 * n + 1, dtype='i8')
        numpy.cumsum([len(r) for r in reads], out=offs[1:])
        units = n if paired else 2 * n
        check_against_oracle(orc, arrays, flat, offs, units, paired)
        ix = _lib.DeviceIndex(*arrays, 60)
        mp = _lib.DeviceMapper(ix)
        mp.map_batch(flat, offs, units, paired)
        short = numpy.asarray([len(r) < 25 for r in reads])
        void = int((short[0::2] | short[1::2]).sum()) if paired else int(short.sum())
        assert mp.sizes()['short_units'] == void
    # a batch of nothing but short reads
    ix = _lib.DeviceIndex(*arrays, 60)
    mp = _lib.DeviceMapper(ix)
    mp.map_batch(numpy.frombuffer(b'ACGT' * 10, dtype='u1'), numpy.asarray([0, 10, 20, 30, 40], dtype='i8'), 2, True)
    t = mp.export()
    assert t['unaligned'] == 2 and t['aligned'] == 0 and t['fld'].sum() == 0 and t['short_units'] == 2


def test_merge_equals_single_mapper(golden_synth, small_tx):
    """Two shards mapped separately then merged == one mapper over everything (multi-GPU merge)."""
    g = golden_synth
    kw = SYNTH_CASES['pe150']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    n = 3000
    bases, _ = sim.generate(0, n)
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    whole = _lib.DeviceMapper(ix)
    whole.map_batch(bases, None, n, True, fixed_len=150)
    a, b = _lib.DeviceMapper(ix), _lib.DeviceMapper(ix)
    half = n // 2
    a.map_batch(bases[:half * 300], None, half, True, first_unit=0, fixed_len=150)
    b.map_batch(bases[half * 300:], None, n - half, True, first_unit=half, fixed_len=150)
    tb = b.export()
    a.merge(tb['key_offsets'], tb['key_ids'], tb['counts'], tb['first_unit'], tb['fld'], tb['unaligned'])
    ta, tw = a.export(), whole.export()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (ta[k] == tw[k]).all(), k
    assert ta['unaligned'] == tw['unaligned'] and ta['aligned'] == tw['aligned']


def test_merge_on_device_matches_single_mapper(golden_synth, small_tx):
    """The multi-GPU exchange path (`dist.merge_mappers`): a peer's raw export, shipped as torch
    tensors, merged into the local dictionary on the device == one mapper over everything."""
    g = golden_synth
    kw = SYNTH_CASES['pe100']
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **kw)
    n = 3000
    bases, _ = sim.generate(0, n)
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    whole = _lib.DeviceMapper(ix)
    whole.map_batch(bases, None, n, True, fixed_len=100)
    parts = [_lib.DeviceMapper(ix) for _ in range(3)]
    bounds = [0, 900, 2100, n]
    for p, lo, hi in zip(parts, bounds[:-1], bounds[1:]):
        p.map_batch(bases[lo * 200:hi * 200], None, hi - lo, True, first_unit=lo, fixed_len=100)
    for peer in parts[1:]:
        t = peer.export_raw_torch()
        parts[0].merge_device(t['key_offsets'], t['key_ids'], t['counts'], t['first_unit'], t['fld'], t['unaligned'])
    ta, tw = parts[0].export(), whole.export()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (ta[k] == tw[k]).all(), k
    assert ta['unaligned'] == tw['unaligned'] and ta['aligned'] == tw['aligned']
    tt = parts[0].export_torch()
    assert (tt['counts'].cpu().numpy() == tw['counts']).all()
    # the same through the packed all-gather layout, all peers in one launch
    import torch
    parts = [_lib.DeviceMapper(ix) for _ in range(3)]
    for p, lo, hi in zip(parts, bounds[:-1], bounds[1:]):
        p.map_batch(bases[lo * 200:hi * 200], None, hi - lo, True, first_unit=lo, fixed_len=100)
    packed = [p.pack_raw_torch() for p in parts]
    cap = max(int(t.shape[0]) for t in packed)
    gathered = torch.zeros(3 * cap, dtype=torch.int64, device='cuda')
    for r, t in enumerate(packed):
        gathered[r * cap:r * cap + t.shape[0]] = t
    parts[1].merge_packed(gathered, cap, 3, 1)
    tb = parts[1].export()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (tb[k] == tw[k]).all(), k
    assert tb['unaligned'] == tw['unaligned'] and tb['aligned'] == tw['aligned']


def test_size_independent_properties_on_a_large_batch(medium):
    """2 M pairs generated on the device (too many for the CPU oracle in a test): every unit lands
    in exactly one class or in `unaligned`; the FLD holds exactly the units with a positive span;
    mapping the same batch again doubles every count and adds no class; two shards merged equal
    the whole."""
    import torch
    tx, arrays = medium
    n = 2_000_000
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), 100, 250, 30, seed=77)
    codes = torch.from_numpy(tx.codes).cuda()
    offs = torch.from_numpy(tx.offsets).cuda()
    cum = torch.from_numpy(sim.cum_weights.view('i8')).cuda()
    d = torch.empty(n * 200, dtype=torch.uint8, device='cuda')
    L = _lib.load()
    _lib.check(L.skm_synth_reads(_lib._ptr(codes), _lib._ptr(offs), tx.n_transcripts, _lib._ptr(cum), sim.total_weight,
                                 sim.L, sim.mu, sim.sd, sim.sub_thresh, sim.n_thresh, sim.random_pct, sim.seed, 1, 0, n,
                                 _lib._ptr(d), 0, _lib.current_stream_ptr()))
    ix = _lib.DeviceIndex(*arrays, tx.n_transcripts)
    whole = _lib.DeviceMapper(ix)
    cls, length = whole.map_batch(d, None, n, True, fixed_len=100, per_read=True)
    t1 = whole.export()
    assert int(t1['counts'].sum()) == t1['aligned'] and t1['aligned'] + t1['unaligned'] == n
    assert int((cls >= 0).sum()) == t1['aligned']
    assert int(t1['fld'].sum()) == int((length > 0).sum())
    assert t1['aligned'] > 0.9 * n and len(t1['counts']) > 100
    whole.map_batch(d, None, n, True, first_unit=n, fixed_len=100)
    t2 = whole.export()
    assert (t2['key_ids'] == t1['key_ids']).all() and (t2['counts'] == 2 * t1['counts']).all()
    assert (t2['fld'] == 2 * t1['fld']).all() and (t2['first_unit'] == t1['first_unit']).all()
    a, b = _lib.DeviceMapper(ix), _lib.DeviceMapper(ix)
    h = n // 3
    a.map_batch(d[:h * 200], None, h, True, first_unit=0, fixed_len=100)
    b.map_batch(d[h * 200:], None, n - h, True, first_unit=h, fixed_len=100)
    tb = b.export_raw_torch()
    a.merge_device(tb['key_offsets'], tb['key_ids'], tb['counts'], tb['first_unit'], tb['fld'], tb['unaligned'])
    ta = a.export()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (ta[k] == t1[k]).all(), k


def test_synth_reads_device_twin(small_tx):
    import ctypes
    import torch
    expr = synth.make_expression(small_tx.n_transcripts, seed=3)
    for kw in (SYNTH_CASES['pe150'], SYNTH_CASES['se75']):
        sim = synth.ReadSimulator(small_tx, expr, **kw)
        n, first = 2500, 123456789012
        want, _ = sim.generate(first, n)
        codes = torch.from_numpy(small_tx.codes).cuda()
        offs = torch.from_numpy(small_tx.offsets).cuda()
        cum = torch.from_numpy(sim.cum_weights.view('i8')).cuda()
        out = torch.empty(want.shape[0], dtype=torch.uint8, device='cuda')
        L = _lib.load()
        _lib.check(L.skm_synth_reads(_lib._ptr(codes), _lib._ptr(offs), small_tx.n_transcripts, _lib._ptr(cum),
                                     sim.total_weight, sim.L, sim.mu, sim.sd, sim.sub_thresh, sim.n_thresh,
                                     sim.random_pct, sim.seed, int(sim.paired), first, n, _lib._ptr(out), 0,
                                     _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        assert (out.cpu().numpy() == want).all()


@pytest.mark.parametrize('ragged', [False, True])
def test_pipelined_host_chunks_match_single_pass(orc, medium, ragged):
    """Host-buffer calls are split into chunks whose H2D copies overlap the kernels; many small
    chunks must give exactly what one device-resident pass gives (fixed and ragged reads)."""
    import torch
    tx, arrays = medium
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts), 100, 250, 30, seed=8)
    n = 300000
    bases, _ = sim.generate(0, n)
    offsets = None
    if ragged:  # drop the last 0..9 bases of every read
        rng = numpy.random.Generator(numpy.random.PCG64(5))
        keep = 100 - rng.integers(0, 10, size=2 * n)
        mask = (numpy.arange(100)[None, :] < keep[:, None]).reshape(-1)
        bases = numpy.ascontiguousarray(bases[mask])
        offsets = numpy.zeros(2 * n + 1, dtype='i8')
        numpy.cumsum(keep, out=offsets[1:])
    ix = _lib.DeviceIndex(*arrays, tx.n_transcripts)
    host, dev = _lib.DeviceMapper(ix), _lib.DeviceMapper(ix)
    hc, hl = host.map_batch(bases, offsets, n, True, fixed_len=0 if ragged else 100, max_len=100, per_read=True)
    d_off = None if offsets is None else torch.from_numpy(offsets).cuda()
    dc, dl = dev.map_batch(torch.from_numpy(bases).cuda(), d_off, n, True, fixed_len=0 if ragged else 100,
                           max_len=100, per_read=True)
    torch.cuda.synchronize()
    th, td = host.export(), dev.export()
    assert (hl == dl.cpu().numpy()).all()
    for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'):
        assert (th[k] == td[k]).all(), k
    assert th['unaligned'] == td['unaligned'] and th['aligned'] == td['aligned']


@pytest.fixture(scope='module')
def long_tx(ref):
    """Transcripts of 8-18 kb (60 exons) indexed by the reference assembler."""
    tx = synth.make_transcriptome(40, seed=13, mean_exons=60, median_exon=200, max_isoforms=6)
    return tx, ref.ref_build_index(tx.sequences())


@pytest.mark.parametrize('paired', [True, False])
def test_maximum_read_length(orc, long_tx, paired):
    """Reads up to the 4096-base limit of the shared-memory staging (128 code words per read:
    the kernel variant with the fewest item rows), ragged, 1 % substitutions: walks over ~35
    contigs per read."""
    tx, arrays = long_tx
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts, seed=3), 4096, 6000, 600,
                              sub_rate=0.01, seed=43)
    n = 400
    raw = sim.generate(0, n)[0].tobytes()
    rng = numpy.random.Generator(numpy.random.PCG64(44))
    reads = []
    for i in range(2 * n):
        length = 4096 if i < 8 else int(rng.integers(1000, 4097))
        reads.append(raw[i * 4096:i * 4096 + length])
    bases, offs = orc.pack_reads(reads)
    units = n if paired else 2 * n
    table = check_against_oracle(orc, arrays, bases, offs, units, paired, batches=2)
    assert table['aligned'] > units // 2


def test_empty_oversized_and_unmappable_inputs(golden_synth):
    from seekmer_b200 import common, infer, mapper
    g = golden_synth
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    mp = _lib.DeviceMapper(ix)
    # an empty batch is a no-op and an empty dictionary exports cleanly
    mp.map_batch(numpy.zeros(0, dtype='u1'), numpy.zeros(1, dtype='i8'), 0, True)
    table = mp.export()
    assert table['counts'].shape[0] == 0 and table['key_ids'].shape[0] == 0
    assert table['aligned'] == 0 and table['unaligned'] == 0 and int(table['fld'].sum()) == 0
    # one base more than the limit is refused, not truncated
    with pytest.raises(_lib.SeekmerCudaError, match='4096'):
        mp.map_batch(numpy.full(4097, ord('A'), dtype='u1'), numpy.asarray([0, 4097], dtype='i8'), 1, False)
    # reads that hit nothing: all N, and a homopolymer run the index does not contain
    junk = [b'N' * 100, b'N' * 100, b'A' * 80, b'C' * 80] * 50
    bases = numpy.frombuffer(b''.join(junk), dtype='u1')
    offs = numpy.concatenate([[0], numpy.cumsum([len(r) for r in junk])]).astype('i8')
    cls, lens = mp.map_batch(bases, offs, 100, True, max_len=100, per_read=True)
    table = mp.export()
    assert (cls == -1).all() and table['unaligned'] == 100 and table['aligned'] == 0
    assert table['counts'].shape[0] == 0 and int(table['fld'].sum()) == 0
    mp.close()
    ix.close()
    # the same through the reference-facing API: nothing mapped -> zero abundances
    index = common.KMerIndex(*g.index_arrays(), g['transcripts'], None)
    for feeder in ([], [(100, [b'r%d' % i for i in range(100)], junk)]):
        res = mapper.map_reads(index, iter(feeder))
        s = res.summarize()
        assert s.aligned == 0 and s.class_map.size == 0 and s.class_count.size == 0
        assert s.unaligned == (100 if feeder else 0) and s.total == s.unaligned
        tpm = infer.quantify(s)                                   # `infer.py:104-105`
        assert tpm.shape == (60,) and (tpm == 0).all()
        assert infer.quantify_bootstraps(s, tpm, 2)[1].shape == (60,)
    index.release_device()


def test_id_pool_holds_stored_ids_only(golden_synth, small_tx):
    """Hundreds of small batches with the default capacities: the id-pool cursor must equal the ids
    stored (a per-warp chunk reservation once leaked ~1 M ids per launch and exhausted the
    default pool on long runs)."""
    g = golden_synth
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **SYNTH_CASES['pe100'])
    n = 2048
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    mp = _lib.DeviceMapper(ix)
    for b in range(300):
        bases, _ = sim.generate(b * n, n)
        mp.map_batch(bases, None, n, True, first_unit=b * n, fixed_len=100)
    s = mp.sizes()
    assert s['n_classes'] > 0 and s['pool_cursor'] == s['n_ids'], s
    t = mp.export()
    assert int(t['key_offsets'][-1]) == s['n_ids']


def test_dictionary_key_collisions_are_refused(golden_synth, small_tx, monkeypatch):
    """Classes are found by a 128-bit hash of their id tuple, and a unit that finds its key present
    compares its ids with the stored tuple.  With deliberately weak keys (only the tuple length)
    different tuples collide: the library must refuse them (SKM_ERR_COLLISION) instead of
    counting them as one class."""
    g = golden_synth
    sim = synth.ReadSimulator(small_tx, synth.make_expression(small_tx.n_transcripts, seed=3), **SYNTH_CASES['pe100'])
    bases, _ = sim.generate(0, 4000)
    ix = _lib.DeviceIndex(*g.index_arrays(), 60)
    monkeypatch.setenv('SKM_TEST_WEAK_KEYS', '1')
    weak = _lib.DeviceMapper(ix)
    monkeypatch.delenv('SKM_TEST_WEAK_KEYS')
    with pytest.raises(_lib.SeekmerCudaError, match='share a 128-bit dictionary key'):
        # ids of a class inserted by a launch become comparable when the launch ends: the second
        # batch at the latest meets them
        weak.map_batch(bases[:2000 * 200], None, 2000, True, fixed_len=100)
        weak.map_batch(bases[2000 * 200:], None, 2000, True, first_unit=2000, fixed_len=100)
    # the same reads with the real keys: fine, and more than one class per tuple length
    ok = _lib.DeviceMapper(ix)
    ok.map_batch(bases, None, 4000, True, fixed_len=100)
    t = ok.export()
    lens = t['key_offsets'][1:] - t['key_offsets'][:-1]
    assert numpy.unique(lens).size < lens.size


def test_reference_built_index_against_the_compiled_reference_mapper(ref, tmp_path):
    """tools/ref_index_parity.py at a small scale: index from the reference's ContigAssembler, reads
    mapped by the CUDA library and by the reference's compiled ReadMapper - class dictionary,
    unaligned count, FLD and first-seen order equal, abundances of the device EM within 1e-6 of the
    stock `infer.quantify` with equal iteration counts.  (BASELINE configs[0] in full - 2 000
    transcripts, 1 M 2x100 pairs - is kept as profiles/r02c_c1_parity.json; benchmark scale - 200 000
    transcripts, 2 M 2x150 pairs - as profiles/r02c_ref_index_parity.json.)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / 'parity.json'
    subprocess.run([sys.executable, os.path.join(root, 'tools', 'ref_index_parity.py'), '--transcripts', '2000',
                    '--pairs', '50000', '--ordered-pairs', '20000', '--read-len', '100', '--frag-mean', '250', '--em',
                    '--out', str(out)], check=True, cwd=root)
    line = json.load(open(out))
    assert line['ok'], line['checks']
