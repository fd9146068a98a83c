"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY).

* mapping: ctypes binding of ``oracle/seekmer_oracle.c`` (plain-C restatement of
  `_mapper.pyx` / `_common.pyx` / `_kmer.pxd`);
* EM / quantify / summarize: numpy restatement of `infer.py:88-168` and
  `mapper.py:77-141`.

Parity status: PINNED against the compiled reference (``oracle/_ref``) and the
golden vectors under ``tests/golden/`` — see ``tests/test_oracle_pinned.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  ``seekmer_b200`` never does.
"""
import ctypes
import os
import pathlib
import subprocess

import numpy

HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = HERE / '_build' / 'libseekmer_oracle.so'
SRC_PATH = HERE / 'seekmer_oracle.c'

K = 25
MAX_FRAGMENT_LENGTH = 2000

SLOT_DTYPE = numpy.dtype([('kmer', '<u8'), ('entry', '<i4'), ('offset', '<i4')])
CONTIG_DTYPE = numpy.dtype([('offset', '<i8'), ('length', '<i8'), ('first_kmer', '<u8'),
                            ('last_kmer', '<u8'), ('target_offset', '<i8'),
                            ('target_count', '<i8')])
TARGET_DTYPE = numpy.dtype([('entry', '<i4'), ('offset', '<i4')])


class _Index(ctypes.Structure):
    _fields_ = [('kmers', ctypes.c_void_p), ('n_slots', ctypes.c_int64),
                ('contigs', ctypes.c_void_p), ('n_contigs', ctypes.c_int64),
                ('sequences', ctypes.c_void_p), ('n_bases', ctypes.c_int64),
                ('targets', ctypes.c_void_p), ('n_targets', ctypes.c_int64)]


class Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in (
        'map_kmer_calls', 'slots', 'map_contig_calls', 'map_contig_items', 'filter_calls',
        'filter_items', 'windows', 'tail_kmers', 'contig_reads')]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """gcc -O3 -fopenmp the C restatement into oracle/_build/ (a few seconds)."""
    if (not force and LIB_PATH.exists()
            and LIB_PATH.stat().st_mtime >= SRC_PATH.stat().st_mtime):
        return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    tmp = LIB_PATH.with_suffix('.tmp%d.so' % os.getpid())
    subprocess.run(['gcc', '-O3', '-march=x86-64-v2', '-fPIC', '-shared', '-fopenmp', '-std=gnu11',
                    str(SRC_PATH), '-o', str(tmp)], check=True)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(LIB_PATH))
        L.skmo_encode.restype = ctypes.c_uint64
        L.skmo_encode.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.skmo_reverse_complement.restype = ctypes.c_uint64
        L.skmo_reverse_complement.argtypes = [ctypes.c_uint64]
        L.skmo_hash.restype = ctypes.c_uint64
        L.skmo_hash.argtypes = [ctypes.c_uint64]
        L.skmo_map_kmer.restype = ctypes.c_uint64  # 8-byte struct returned in rax
        L.skmo_map_kmer.argtypes = [ctypes.POINTER(_Index), ctypes.c_uint64]
        L.skmo_sift4_align_left.restype = ctypes.c_int
        L.skmo_sift4_align_left.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                            ctypes.c_int, ctypes.c_int]
        L.skmo_sift4_align_right.restype = ctypes.c_int
        L.skmo_sift4_align_right.argtypes = L.skmo_sift4_align_left.argtypes
        L.skmo_map_batch.restype = ctypes.c_int64
        L.skmo_map_batch.argtypes = [ctypes.POINTER(_Index), ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]
        L.skmo_map_batch_mt.restype = ctypes.c_int64
        L.skmo_map_batch_mt.argtypes = [ctypes.POINTER(_Index), ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]
        L.skmo_tally.restype = ctypes.c_int64
        L.skmo_tally.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] + [ctypes.c_void_p] * 4
        L.skmo_effective_lengths.restype = None
        L.skmo_effective_lengths.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_void_p]
        L.skmo_table_insert.restype = ctypes.c_int64
        L.skmo_table_insert.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleIndex:
    """Borrowed view of the four hot-path index arrays (SURVEY §8(a) I1–I4)."""

    def __init__(self, kmers, contigs, sequences, targets):
        self.kmers = numpy.ascontiguousarray(numpy.asarray(kmers).view(SLOT_DTYPE))
        c = numpy.asarray(contigs)
        self.contigs = numpy.ascontiguousarray(c.view(CONTIG_DTYPE) if c.dtype.itemsize == 48 else c)
        self.sequences = numpy.ascontiguousarray(numpy.asarray(sequences).view('u1'))
        self.targets = numpy.ascontiguousarray(numpy.asarray(targets).view(TARGET_DTYPE))
        assert self.kmers.shape[0] & (self.kmers.shape[0] - 1) == 0
        self.c = _Index(_ptr(self.kmers), self.kmers.shape[0], _ptr(self.contigs),
                        self.contigs.shape[0], _ptr(self.sequences), self.sequences.shape[0],
                        _ptr(self.targets), self.targets.shape[0])

    def map_kmer(self, kmer):
        v = lib().skmo_map_kmer(ctypes.byref(self.c), ctypes.c_uint64(int(kmer)))
        entry = v & 0xFFFFFFFF
        offset = (v >> 32) & 0xFFFFFFFF
        if entry >= 1 << 31:
            entry -= 1 << 32
        if offset >= 1 << 31:
            offset -= 1 << 32
        return entry, offset


def encode(seq, offset=0):
    return lib().skmo_encode(seq, offset)


def reverse_complement(kmer):
    return lib().skmo_reverse_complement(ctypes.c_uint64(int(kmer)))


def kmer_hash(kmer):
    return lib().skmo_hash(ctypes.c_uint64(int(kmer)))


def pack_reads(reads):
    """list[bytes] -> (uint8 bases, int64 offsets[n+1]); every read must have len >= 25."""
    lens = numpy.fromiter((len(r) for r in reads), dtype='i8', count=len(reads))
    offsets = numpy.zeros(len(reads) + 1, dtype='i8')
    numpy.cumsum(lens, out=offsets[1:])
    bases = numpy.frombuffer(b''.join(reads), dtype='u1')
    return bases, offsets


class MapOutput:
    __slots__ = ('ptr', 'ids', 'length', 'fld', 'counters')

    def tuples(self):
        p = self.ptr
        ids = self.ids.tolist()
        return [tuple(ids[p[i]:p[i + 1]]) for i in range(len(p) - 1)]


def map_batch(index, bases, offsets, paired, counters=False):
    """Single-threaded map of one batch; per-unit ordered id tuples + lengths + FLD."""
    bases = numpy.ascontiguousarray(bases, dtype='u1')
    offsets = numpy.ascontiguousarray(offsets, dtype='i8')
    n_reads = offsets.shape[0] - 1
    n_units = n_reads // 2 if paired else n_reads
    out = MapOutput()
    out.ptr = numpy.zeros(n_units + 1, dtype='i8')
    out.length = numpy.zeros(n_units, dtype='i4')
    out.fld = numpy.zeros(MAX_FRAGMENT_LENGTH, dtype='i8')
    cnt = Counters() if counters else None
    cap = max(1024, n_units * 8)
    while True:
        ids = numpy.zeros(cap, dtype='i4')
        out.fld[:] = 0
        r = lib().skmo_map_batch(ctypes.byref(index.c), _ptr(bases), _ptr(offsets), n_units,
                                 int(bool(paired)), _ptr(out.ptr), _ptr(ids), cap,
                                 _ptr(out.length), _ptr(out.fld),
                                 ctypes.byref(cnt) if cnt is not None else None)
        if r >= 0:
            out.ids = ids[:r].copy()
            break
        cap = -r
        if cnt is not None:
            cnt = Counters()
    out.counters = cnt.as_dict() if cnt is not None else None
    return out


def map_batch_mt(index, bases, offsets, paired, n_threads):
    """Multi-threaded timing driver (CPU baseline). Returns (aligned, hash, count, length, fld)."""
    bases = numpy.ascontiguousarray(bases, dtype='u1')
    offsets = numpy.ascontiguousarray(offsets, dtype='i8')
    n_reads = offsets.shape[0] - 1
    n_units = n_reads // 2 if paired else n_reads
    h = numpy.zeros(n_units, dtype='u8')
    cnt = numpy.zeros(n_units, dtype='i4')
    length = numpy.zeros(n_units, dtype='i4')
    fld = numpy.zeros(MAX_FRAGMENT_LENGTH, dtype='i8')
    aligned = lib().skmo_map_batch_mt(ctypes.byref(index.c), _ptr(bases), _ptr(offsets), n_units,
                                      int(bool(paired)), int(n_threads), _ptr(h), _ptr(cnt),
                                      _ptr(length), _ptr(fld))
    return aligned, h, cnt, length, fld


def tally(ptr, ids):
    """Counter semantics of MapResult.update at job_count=1 (`mapper.py:60-75`).

    Returns (cls_ptr, cls_ids, cls_count, unaligned) with classes in first-insertion order.
    """
    ptr = numpy.ascontiguousarray(ptr, dtype='i8')
    ids = numpy.ascontiguousarray(ids, dtype='i4')
    n_units = ptr.shape[0] - 1
    cls_ptr = numpy.zeros(n_units + 1, dtype='i8')
    cls_ids = numpy.zeros(max(1, ids.shape[0]), dtype='i4')
    cls_count = numpy.zeros(max(1, n_units), dtype='i8')
    una = ctypes.c_int64(0)
    n = lib().skmo_tally(_ptr(ptr), _ptr(ids), n_units, _ptr(cls_ptr), _ptr(cls_ids),
                         _ptr(cls_count), ctypes.byref(una))
    cls_ptr = cls_ptr[:n + 1].copy()
    return cls_ptr, cls_ids[:cls_ptr[-1]].copy(), cls_count[:n].copy(), int(una.value)


def tally_dict(ptr, ids):
    cls_ptr, cls_ids, cls_count, una = tally(ptr, ids)
    lst = cls_ids.tolist()
    d = {tuple(lst[cls_ptr[i]:cls_ptr[i + 1]]): int(cls_count[i]) for i in range(len(cls_count))}
    if una:
        d[()] = una
    return d


def effective_lengths(fld, lengths):
    """`MapResult.effective_lengths` (`mapper.py:134-141`)."""
    fld = numpy.ascontiguousarray(fld, dtype='i8')
    lengths = numpy.ascontiguousarray(lengths, dtype='f8')
    out = numpy.zeros(lengths.shape[0], dtype='f8')
    with numpy.errstate(all='ignore'):
        lib().skmo_effective_lengths(_ptr(fld), _ptr(lengths), lengths.shape[0], _ptr(out))
    return out


def harmonic_mean_fragment_length(fld):
    """`MapResult.harmonic_mean_fragment_length` (`mapper.py:117-132`)."""
    fld = numpy.asarray(fld)
    numerator = fld.sum()
    if numerator == 0:
        return 0
    denominator = (fld[1:].astype('f8') / numpy.arange(1, MAX_FRAGMENT_LENGTH)).sum()
    return numerator / denominator


def class_map_from_csr(cls_ptr, cls_ids):
    """`MapResult.summarize` class_map (2, nnz) int64 (`mapper.py:85-93`)."""
    cls_ptr = numpy.asarray(cls_ptr, dtype='i8')
    sizes = cls_ptr[1:] - cls_ptr[:-1]
    rows = numpy.repeat(numpy.arange(sizes.shape[0], dtype='i8'), sizes)
    return numpy.stack([rows, numpy.asarray(cls_ids, dtype='i8')])


def em(x, l, class_map, class_count, return_iters=False):
    """`infer.em` (`infer.py:133-168`), restated; also reports the iteration count."""
    n = class_count.sum()
    iters = 1
    with numpy.errstate(all='ignore'):
        old_x = x
        x = x[class_map[1]]
        class_inner = numpy.bincount(class_map[0], weights=x,
                                     minlength=class_count.size) / class_count
        x = numpy.bincount(class_map[1], weights=x / class_inner[class_map[0]],
                           minlength=l.size) / l / n
        x[x != x] = 0
        while True:
            sel = x > 1e-8
            if not sel.any():
                # the reference raises ValueError here (max of empty, infer.py:160);
                # the restatement reports "converged" so callers can define behaviour
                break
            if not ((numpy.absolute(x - old_x) / x)[sel].max() > 0.01):
                break
            old_x = x
            x = x[class_map[1]]
            class_inner = numpy.bincount(class_map[0], weights=x,
                                         minlength=class_count.size) / class_count
            x = numpy.bincount(class_map[1], weights=x / class_inner[class_map[0]],
                               minlength=l.size) / l / n
            x[x != x] = 0
            iters += 1
    return (x, iters) if return_iters else x


def quantify(eff_lengths, class_map, class_count, x0=None, return_iters=False):
    """`infer.quantify` without the resampling step (`infer.py:88-130`).

    Bootstrap resampling is done by the caller (`class_count` = resampled counts) because
    the reference uses an unseeded global RNG (`infer.py:108-111`).
    """
    transcript_length = numpy.asarray(eff_lengths, dtype='f8')
    if class_map.size == 0:
        z = numpy.zeros(transcript_length.size, dtype='f8')
        return (z, 0) if return_iters else z
    class_count = numpy.asarray(class_count, dtype='f8')
    if x0 is None:
        x = numpy.ones(transcript_length.size, dtype='f8') / transcript_length
    else:
        x = numpy.array(x0, dtype='f8', copy=True)
    x /= x.sum()
    x, iters = em(x, transcript_length, class_map, class_count, return_iters=True)
    with numpy.errstate(all='ignore'):
        x /= x.sum() / 1000000
        x[x < 0.001] = 0
        x /= x.sum() / 1000000
    return (x, iters) if return_iters else x


def est_counts(tpm, lengths, aligned):
    """`infer._infer_est_counts` (`infer.py:233-252`)."""
    e = numpy.asarray(tpm, dtype='f8') * numpy.asarray(lengths, dtype='f8')
    with numpy.errstate(all='ignore'):
        e *= aligned / e.sum()
    return e


# ---- counter-based RNG shared by the synthetic-data generators and the bootstrap ----
# Philox4x32-10 (Salmon et al., SC'11), restated in numpy; the CUDA side restates it too and
# tests compare the two bit-for-bit.
_PHILOX_M0 = numpy.uint64(0xD2511F53)
_PHILOX_M1 = numpy.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = numpy.asarray(c0, dtype='u8') & 0xFFFFFFFF
    c1 = numpy.asarray(c1, dtype='u8') & 0xFFFFFFFF
    c2 = numpy.asarray(c2, dtype='u8') & 0xFFFFFFFF
    c3 = numpy.asarray(c3, dtype='u8') & 0xFFFFFFFF
    c0, c1, c2, c3 = numpy.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & 0xFFFFFFFF, lo1, (hi0 ^ c3 ^ k1) & 0xFFFFFFFF, lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype('u4'), c1.astype('u4'), c2.astype('u4'), c3.astype('u4'))


def bootstrap_counts(class_count, n_replicates, seed):
    """Resample reads with replacement: counts ~ Multinomial(n, class_count / n).

    Same distribution as `scipy.stats.multinomial(n, p).rvs()` at `infer.py:108-111`, realised
    as n i.i.d. categorical draws with integer arithmetic only: draws 2j and 2j+1 of replicate r
    take Philox(counter=(j_lo, j_hi, r, 0), key=(seed_lo, seed_hi)) -> 64-bit words
    w = (x1<<32)|x0 and (x3<<32)|x2, u = (w * n) >> 64 (uniform in [0, n)),
    class = upper_bound(cumsum(count), u).
    Bit-exact with the device resampler for the same (seed, replicate).
    """
    cc = numpy.asarray(class_count)
    counts_i = cc.astype('i8')
    assert (counts_i == cc).all(), 'class counts must be integral'
    cum = numpy.cumsum(counts_i).astype('u8')
    n = int(cum[-1]) if cum.size else 0
    out = numpy.zeros((n_replicates, counts_i.shape[0]), dtype='i8')
    if n == 0:
        return out
    j = numpy.arange((n + 1) // 2, dtype='u8')
    for r in range(n_replicates):
        x0, x1, x2, x3 = philox4x32(j & 0xFFFFFFFF, j >> 32, r, 0, seed & 0xFFFFFFFF, seed >> 32)
        w = numpy.empty(2 * j.shape[0], dtype='u8')
        w[0::2] = (x1.astype('u8') << numpy.uint64(32)) | x0.astype('u8')
        w[1::2] = (x3.astype('u8') << numpy.uint64(32)) | x2.astype('u8')
        w = w[:n]
        # (w * n) >> 64 with 64-bit pieces
        n_lo, n_hi = n & 0xFFFFFFFF, n >> 32
        w_lo, w_hi = w & 0xFFFFFFFF, w >> 32
        ll = w_lo * numpy.uint64(n_lo)
        lh = w_lo * numpy.uint64(n_hi)
        hl = w_hi * numpy.uint64(n_lo)
        hh = w_hi * numpy.uint64(n_hi)
        mid = (ll >> 32) + (lh & 0xFFFFFFFF) + (hl & 0xFFFFFFFF)
        u = hh + (lh >> 32) + (hl >> 32) + (mid >> 32)
        cls = numpy.searchsorted(cum, u, side='right')
        out[r] = numpy.bincount(cls, minlength=counts_i.shape[0])
    return out
