"""ctypes binding of ``libseekmer_b200.so`` (the C ABI in ``include/seekmer_b200.h``).

There is no CPU fallback: if the CUDA library is missing or no device is visible, the
compute entry points raise.  Nothing here imports ``oracle/``.
"""
import ctypes
import os
import pathlib

import numpy

HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get('SEEKMER_B200_LIB', HERE / 'libseekmer_b200.so'))

K = 25
MAX_FRAGMENT_LENGTH = 2000
RESAMPLE_DRAWS, RESAMPLE_TREE = 0, 1  # skm_multinomial methods (include/seekmer_b200.h)

SLOT_DTYPE = numpy.dtype([('kmer', '<u8'), ('entry', '<i4'), ('offset', '<i4')])
CONTIG_DTYPE = numpy.dtype([('offset', '<i8'), ('length', '<i8'), ('first_kmer', '<u8'),
                            ('last_kmer', '<u8'), ('target_offset', '<i8'),
                            ('target_count', '<i8')])
TARGET_DTYPE = numpy.dtype([('entry', '<i4'), ('offset', '<i4')])

EXPORTS = (
    'skm_last_error', 'skm_device_count', 'skm_version', 'skm_index_create', 'skm_index_destroy',
    'skm_index_info', 'skm_map_kmers', 'skm_mapper_create', 'skm_mapper_destroy',
    'skm_mapper_reset', 'skm_map_batch', 'skm_map_fastq', 'skm_mapper_kernel_ms', 'skm_debug_map_stats', 'skm_classes_size', 'skm_classes_export',
    'skm_classes_merge', 'skm_classes_merge_packed', 'skm_release_cache', 'skm_effective_lengths', 'skm_em', 'skm_em_samples', 'skm_multinomial', 'skm_em_bootstrap', 'skm_synth_reads',
    'skm_build_kmer_table', 'skm_index_save', 'skm_index_load', 'skm_em_plan_create', 'skm_em_plan_from_mapper', 'skm_em_plan_info',
    'skm_em_plan_destroy', 'skm_em_plan_run', 'skm_em_plan_bootstrap', 'skm_em_plans_run', 'skm_scratch_local_only',
)


class SeekmerCudaError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SeekmerCudaError(
            'seekmer_b200: %s is missing - build it with `python -m seekmer_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    L = ctypes.CDLL(str(LIB_PATH))
    vp, i64, i32, u64, ci = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64,
                             ctypes.c_int)
    L.skm_last_error.restype = ctypes.c_char_p
    L.skm_last_error.argtypes = []
    L.skm_version.restype = ctypes.c_char_p
    L.skm_version.argtypes = []
    L.skm_device_count.restype = ci
    L.skm_device_count.argtypes = []
    L.skm_index_create.restype = ci
    L.skm_index_create.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i64, ci, ci, vp,
                                   ctypes.POINTER(vp)]
    L.skm_index_destroy.restype = None
    L.skm_index_destroy.argtypes = [vp]
    L.skm_index_save.restype = ci
    L.skm_index_save.argtypes = [vp, ctypes.c_char_p, vp, i64, vp]
    L.skm_index_load.restype = ci
    L.skm_index_load.argtypes = [ctypes.c_char_p, ci, vp, ctypes.POINTER(vp), vp, vp]
    L.skm_index_info.restype = ci
    L.skm_index_info.argtypes = [vp, vp]
    L.skm_map_kmers.restype = ci
    L.skm_map_kmers.argtypes = [vp, vp, i64, vp, vp, ci, vp]
    L.skm_mapper_create.restype = ci
    L.skm_mapper_create.argtypes = [vp, i64, i64, ctypes.POINTER(vp)]
    L.skm_mapper_destroy.restype = None
    L.skm_mapper_destroy.argtypes = [vp]
    L.skm_mapper_reset.restype = ci
    L.skm_mapper_reset.argtypes = [vp, vp]
    L.skm_map_batch.restype = ci
    L.skm_map_batch.argtypes = [vp, vp, vp, i32, i32, i64, ci, i64, ci, vp, vp, vp]
    L.skm_map_fastq.restype = ci
    L.skm_map_fastq.argtypes = [vp, vp, i64, vp, i64, i64, ci, vp, vp, vp, vp, vp, vp]
    L.skm_debug_map_stats.restype = ci
    L.skm_debug_map_stats.argtypes = [vp, ci]
    L.skm_mapper_kernel_ms.restype = ci
    L.skm_mapper_kernel_ms.argtypes = [vp, vp]
    L.skm_classes_size.restype = ci
    L.skm_classes_size.argtypes = [vp, vp, vp]
    L.skm_classes_export.restype = ci
    L.skm_classes_export.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, vp]
    L.skm_classes_merge.restype = ci
    L.skm_classes_merge.argtypes = [vp, vp, vp, vp, vp, i64, vp, i64, ci, vp]
    L.skm_classes_merge_packed.restype = ci
    L.skm_classes_merge_packed.argtypes = [vp, vp, i64, ci, ci, vp]
    L.skm_release_cache.restype = ci
    L.skm_release_cache.argtypes = [ci, vp]
    L.skm_scratch_local_only.restype = ci
    L.skm_scratch_local_only.argtypes = [ci]
    L.skm_effective_lengths.restype = ci
    L.skm_effective_lengths.argtypes = [vp, vp, i64, vp, ci, ci, vp]
    L.skm_em.restype = ci
    L.skm_em.argtypes = [vp, vp, i64, i64, vp, vp, i64, vp, i64, i64, vp, vp, ci, ci, vp]
    L.skm_em_samples.restype = ci
    L.skm_em_samples.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, ci, ci, vp]
    L.skm_em_bootstrap.restype = ci
    L.skm_em_bootstrap.argtypes = [vp, vp, i64, i64, vp, vp, i64, vp, i64, i64, u64, ci, i64, ci, vp, vp, ci, ci, vp]
    L.skm_em_plan_create.restype = ci
    L.skm_em_plan_create.argtypes = [vp, vp, i64, i64, i64, vp, ci, ci, vp, ctypes.POINTER(vp)]
    L.skm_em_plan_from_mapper.restype = ci
    L.skm_em_plan_from_mapper.argtypes = [vp, i64, vp, ctypes.POINTER(vp)]
    L.skm_em_plan_info.restype = ci
    L.skm_em_plan_info.argtypes = [vp, vp]
    L.skm_em_plan_destroy.restype = None
    L.skm_em_plan_destroy.argtypes = [vp]
    L.skm_em_plan_run.restype = ci
    L.skm_em_plan_run.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, ci, vp]
    L.skm_em_plans_run.restype = ci
    L.skm_em_plans_run.argtypes = [vp, i64, vp, vp, i64, vp, vp, ci, vp]
    L.skm_em_plan_bootstrap.restype = ci
    L.skm_em_plan_bootstrap.argtypes = [vp, vp, vp, vp, i64, i64, u64, ci, i64, ci, vp, vp, ci, vp]
    L.skm_multinomial.restype = ci
    L.skm_multinomial.argtypes = [vp, i64, i64, i64, u64, ci, vp, ci, ci, vp]
    L.skm_synth_reads.restype = ci
    L.skm_synth_reads.argtypes = [vp, vp, i64, vp, u64, i32, i32, i32, i32, i32, i32, u64, ci,
                                  i64, i64, vp, ci, vp]
    L.skm_build_kmer_table.restype = ci
    L.skm_build_kmer_table.argtypes = [vp, vp, vp, i64, vp, i64, ci, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = load().skm_last_error()
        raise SeekmerCudaError('%s (status %d)' % (msg.decode() if msg else 'unknown error', rc))


def device_count():
    return load().skm_device_count()


def require_device():
    if device_count() < 1:
        raise SeekmerCudaError('seekmer_b200: no CUDA device visible; there is no CPU fallback')


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _is_torch(x):
    return type(x).__module__.startswith('torch')


def _ptr(x):
    """void* of a numpy array, a torch tensor (host or device) or None."""
    if x is None:
        return None
    if _is_torch(x):
        return ctypes.c_void_p(x.data_ptr())
    return _np_ptr(x)


def current_stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DeviceIndex:
    """Handle to the HBM-resident, re-laid-out index (``skm_index``)."""

    def __init__(self, kmers, contigs, sequences, targets, n_transcripts, device=0, stream=None):
        L = load()
        require_device()
        on_device = _is_torch(kmers)
        if on_device:
            if not (kmers.is_cuda and contigs.is_cuda and sequences.is_cuda and targets.is_cuda):
                raise ValueError('torch index arrays must all live on the CUDA device')
            device = kmers.device.index
            n_slots = kmers.numel() * kmers.element_size() // 16
            n_contigs = contigs.numel() * contigs.element_size() // 48
            n_bases = sequences.numel() * sequences.element_size()
            n_targets = targets.numel() * targets.element_size() // 8
            if stream is None:
                stream = current_stream_ptr(kmers.device)
        else:
            kmers = _as_struct(kmers, SLOT_DTYPE)
            contigs = _as_struct(contigs, CONTIG_DTYPE)
            sequences = numpy.ascontiguousarray(numpy.asarray(sequences).view('u1'))
            targets = _as_struct(targets, TARGET_DTYPE)
            n_slots, n_contigs = kmers.shape[0], contigs.shape[0]
            n_bases, n_targets = sequences.shape[0], targets.shape[0]
        self._keep = (kmers, contigs, sequences, targets)
        handle = ctypes.c_void_p()
        check(L.skm_index_create(_ptr(kmers), n_slots, _ptr(contigs), n_contigs, _ptr(sequences),
                                 n_bases, _ptr(targets), n_targets, int(n_transcripts),
                                 int(device), int(on_device), stream, ctypes.byref(handle)))
        self._h = handle
        self._keep = None
        self.device = int(device)

    @classmethod
    def load(cls, path, device=0, stream=None):
        """From a device image file (`skm_index_save`): plain copies, no relayout."""
        require_device()
        self = cls.__new__(cls)
        handle = ctypes.c_void_p()
        where = numpy.zeros(2, dtype='i8')
        check(load().skm_index_load(str(path).encode(), int(device), stream, ctypes.byref(handle),
                                    ctypes.c_void_p(where.ctypes.data), ctypes.c_void_p(where.ctypes.data + 8)))
        self._h = handle
        self._keep = None
        self.device = int(device)
        self.trailer = (int(where[0]), int(where[1]))
        return self

    def save(self, path, trailer=b'', stream=None):
        """Write the device image (+ `trailer` bytes after it)."""
        buf = ctypes.create_string_buffer(trailer, len(trailer)) if trailer else None
        check(load().skm_index_save(self._h, str(path).encode(), buf, len(trailer), stream))

    def info(self):
        a = numpy.zeros(8, dtype='i8')
        check(load().skm_index_info(self._h, _np_ptr(a)))
        keys = ('n_kmers', 'table_slots', 'max_target_count', 'device_bytes', 'n_contigs',
                'n_targets', 'n_transcripts', 'device')
        return dict(zip(keys, a.tolist()))

    def map_kmers(self, kmers):
        """`KMerIndex.map_kmer` for a vector of encoded k-mers -> (entry, offset) int32 arrays."""
        kmers = numpy.ascontiguousarray(kmers, dtype='u8')
        e = numpy.zeros(kmers.shape[0], dtype='i4')
        o = numpy.zeros(kmers.shape[0], dtype='i4')
        check(load().skm_map_kmers(self._h, _np_ptr(kmers), kmers.shape[0], _np_ptr(e), _np_ptr(o),
                                   0, None))
        return e, o

    def close(self):
        if getattr(self, '_h', None):
            load().skm_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_struct(a, dtype):
    a = numpy.asarray(a)
    if a.dtype != dtype:
        if a.dtype.itemsize != dtype.itemsize:
            raise ValueError('index array has item size %d, expected %d'
                             % (a.dtype.itemsize, dtype.itemsize))
        a = numpy.ascontiguousarray(a).view(dtype)
    return numpy.ascontiguousarray(a)


class DeviceMapper:
    """Handle to a device class dictionary + FLD (``skm_mapper``)."""

    def __init__(self, index, class_capacity=0, id_capacity=0):
        self.index = index
        handle = ctypes.c_void_p()
        check(load().skm_mapper_create(index._h, int(class_capacity), int(id_capacity),
                                       ctypes.byref(handle)))
        self._h = handle

    def reset(self, stream=None):
        check(load().skm_mapper_reset(self._h, stream))

    def map_batch(self, bases, offsets, n_units, paired, first_unit=0, fixed_len=0, max_len=0,
                  per_read=False, stream=None):
        """Map one batch.  `bases`/`offsets` are numpy (host) or torch CUDA tensors."""
        on_device = _is_torch(bases) and bases.is_cuda
        out_class = out_length = None
        if per_read:
            if on_device:
                import torch
                out_class = torch.empty(n_units, dtype=torch.int32, device=bases.device)
                out_length = torch.empty(n_units, dtype=torch.int32, device=bases.device)
            else:
                out_class = numpy.empty(n_units, dtype='i4')
                out_length = numpy.empty(n_units, dtype='i4')
        if on_device and stream is None:
            stream = current_stream_ptr(bases.device)
        if not on_device:
            if _is_torch(bases):  # pinned host tensor
                pass
            else:
                bases = numpy.ascontiguousarray(bases, dtype='u1')
                if offsets is not None:
                    offsets = numpy.ascontiguousarray(offsets, dtype='i8')
        check(load().skm_map_batch(self._h, _ptr(bases), _ptr(offsets), int(fixed_len),
                                   int(max_len), int(n_units), int(bool(paired)), int(first_unit),
                                   int(on_device), _ptr(out_class), _ptr(out_length), stream))
        return out_class, out_length

    def map_fastq(self, text1, n1, text2=None, n2=0, first_unit=0, stream=None):
        """Map the whole records of raw FASTQ text (uint8 numpy arrays, host).  Returns
        (n_units, consumed1, consumed2)."""
        out = numpy.zeros(3, dtype='i8')
        base = out.ctypes.data
        check(load().skm_map_fastq(self._h, _ptr(text1), int(n1), _ptr(text2), int(n2), int(first_unit), 0,
                                   ctypes.c_void_p(base), ctypes.c_void_p(base + 8) if text2 is not None else None,
                                   ctypes.c_void_p(base + 16), None, None, stream))
        return int(out[2]), int(out[0]), int(out[1])

    def kernel_ms(self):
        """Device durations (ms) of pack / map / tally kernels of the last mapped chunk."""
        a = numpy.zeros(3, dtype='f8')
        check(load().skm_mapper_kernel_ms(self._h, _np_ptr(a)))
        return dict(zip(('pack_reads_kernel', 'map_reads_kernel', 'tally_units_kernel'), a.tolist()))

    def sizes(self, stream=None):
        a = numpy.zeros(8, dtype='i8')
        check(load().skm_classes_size(self._h, _np_ptr(a), stream))
        keys = ('n_classes', 'n_ids', 'unaligned', 'aligned', 'capacity', 'status', 'short_units', 'pool_cursor')
        return dict(zip(keys, a.tolist()))

    def export(self, with_slots=False, stream=None):
        """Host copy of the dictionary, sorted by first-seen unit (the reference's Counter order
        at job_count=1).  Returns dict(key_offsets, key_ids, counts, first_unit, fld, unaligned,
        aligned[, slots])."""
        sz = self.sizes(stream)
        n, n_ids = sz['n_classes'], sz['n_ids']
        off = numpy.zeros(n + 1, dtype='i8')
        ids = numpy.zeros(max(n_ids, 1), dtype='i4')
        counts = numpy.zeros(max(n, 1), dtype='i8')
        first = numpy.zeros(max(n, 1), dtype='i8')
        slots = numpy.zeros(max(n, 1), dtype='i4')
        fld = numpy.zeros(MAX_FRAGMENT_LENGTH, dtype='i8')
        check(load().skm_classes_export(self._h, _np_ptr(off), _np_ptr(ids), _np_ptr(counts),
                                        _np_ptr(first), _np_ptr(slots), _np_ptr(fld), 0, stream))
        counts, first, slots, ids = counts[:n], first[:n], slots[:n], ids[:n_ids]
        order = numpy.argsort(first, kind='stable')
        lens = off[1:] - off[:-1]
        new_off = numpy.zeros(n + 1, dtype='i8')
        numpy.cumsum(lens[order], out=new_off[1:])
        if n:
            gather = (numpy.repeat(off[:-1][order] - new_off[:-1], lens[order])
                      + numpy.arange(n_ids, dtype='i8'))
            ids = ids[gather]
        out = dict(key_offsets=new_off, key_ids=ids, counts=counts[order], first_unit=first[order],
                   fld=fld, unaligned=sz['unaligned'], aligned=sz['aligned'], short_units=sz['short_units'])
        if with_slots:
            out['slots'] = slots[order]
        return out

    def export_torch(self, stream=None):
        """Dictionary as torch tensors on the mapper's GPU, sorted by first-seen unit — the table
        shape `seekmer_b200.dist.merge_class_tables` exchanges (no host round trip)."""
        import torch
        dev = torch.device('cuda', self.index.device)
        if stream is None:
            stream = current_stream_ptr(dev)
        sz = self.sizes(stream)
        n, n_ids = sz['n_classes'], sz['n_ids']
        off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ids = torch.zeros(n_ids, dtype=torch.int32, device=dev)
        counts = torch.zeros(n, dtype=torch.int64, device=dev)
        first = torch.zeros(n, dtype=torch.int64, device=dev)
        fld = torch.zeros(MAX_FRAGMENT_LENGTH, dtype=torch.int64, device=dev)
        check(load().skm_classes_export(self._h, _ptr(off), _ptr(ids) if n_ids else None,
                                        _ptr(counts) if n else None, _ptr(first) if n else None,
                                        None, _ptr(fld), 1, stream))
        order = torch.argsort(first, stable=True)
        lens = (off[1:] - off[:-1])[order]
        new_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        new_off[1:] = torch.cumsum(lens, 0)
        src = (torch.repeat_interleave(off[:-1][order] - new_off[:-1], lens)
               + torch.arange(n_ids, device=dev))
        return dict(key_offsets=new_off, key_ids=ids[src], counts=counts[order], first_unit=first[order],
                    fld=fld, scalars=torch.tensor([sz['unaligned'], sz['aligned']], dtype=torch.int64,
                                                  device=dev))

    def export_raw_torch(self, stream=None):
        """Dictionary as torch tensors on the mapper's GPU in table order (unsorted): the cheap
        form for shipping to other ranks."""
        import torch
        dev = torch.device('cuda', self.index.device)
        if stream is None:
            stream = current_stream_ptr(dev)
        sz = self.sizes(stream)
        n, n_ids = sz['n_classes'], sz['n_ids']
        off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ids = torch.zeros(max(n_ids, 1), dtype=torch.int32, device=dev)
        counts = torch.zeros(max(n, 1), dtype=torch.int64, device=dev)
        first = torch.zeros(max(n, 1), dtype=torch.int64, device=dev)
        fld = torch.zeros(MAX_FRAGMENT_LENGTH, dtype=torch.int64, device=dev)
        check(load().skm_classes_export(self._h, _ptr(off), _ptr(ids), _ptr(counts), _ptr(first), None,
                                        _ptr(fld), 1, stream))
        return dict(key_offsets=off, key_ids=ids[:n_ids], counts=counts[:n], first_unit=first[:n], fld=fld,
                    unaligned=sz['unaligned'], aligned=sz['aligned'])

    def merge_device(self, key_offsets, key_ids, counts, first_unit, fld, unaligned, stream=None):
        """Add another rank's exported classes (torch tensors on this GPU) to the dictionary."""
        if stream is None:
            stream = current_stream_ptr(key_ids.device)
        n = int(counts.shape[0])
        check(load().skm_classes_merge(self._h, _ptr(key_offsets), _ptr(key_ids) if key_ids.numel() else None,
                                       _ptr(counts) if n else None, _ptr(first_unit) if n else None, n,
                                       _ptr(fld), int(unaligned), 1, stream))

    def pack_raw_torch(self, stream=None):
        """The raw export as ONE int64 tensor on the GPU, in the layout `skm_classes_merge_packed`
        reads: [n_classes, n_ids, unaligned | fld | key_offsets | counts | first_unit | key_ids]."""
        import torch
        t = self.export_raw_torch(stream)
        dev = t['fld'].device
        n_ids = int(t['key_ids'].shape[0])
        ids64 = torch.zeros((n_ids + 1) // 2, dtype=torch.int64, device=dev)
        ids64.view(torch.int32)[:n_ids] = t['key_ids']
        head = torch.tensor([int(t['counts'].shape[0]), n_ids, t['unaligned']], dtype=torch.int64, device=dev)
        return torch.cat([head, t['fld'], t['key_offsets'], t['counts'], t['first_unit'], ids64])

    def merge_packed(self, gathered, words_per_rank, world, rank, stream=None):
        """Insert every other rank's packed export (`pack_raw_torch`, all-gathered) at once."""
        if stream is None:
            stream = current_stream_ptr(gathered.device)
        check(load().skm_classes_merge_packed(self._h, _ptr(gathered), int(words_per_rank), int(world), int(rank),
                                              stream))

    def merge(self, key_offsets, key_ids, counts, first_unit, fld=None, unaligned=0, stream=None):
        key_offsets = numpy.ascontiguousarray(key_offsets, dtype='i8')
        key_ids = numpy.ascontiguousarray(key_ids, dtype='i4')
        counts = numpy.ascontiguousarray(counts, dtype='i8')
        first_unit = numpy.ascontiguousarray(first_unit, dtype='i8')
        if fld is not None:
            fld = numpy.ascontiguousarray(fld, dtype='i8')
        check(load().skm_classes_merge(self._h, _np_ptr(key_offsets), _np_ptr(key_ids),
                                       _np_ptr(counts), _np_ptr(first_unit), counts.shape[0],
                                       _ptr(fld), int(unaligned), 0, stream))

    def close(self):
        if getattr(self, '_h', None):
            load().skm_mapper_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PEERS = {'on': False}


def uses_peer_gpus(on=True):
    """Tell the library that this process has (or will have) peer access to other GPUs switched on
    - NCCL, or one host thread per device with peer copies: EM scratch then comes from blocks only
    the owning device maps (`skm_scratch_local_only`), because `cudaMalloc` under peer access
    costs 100+ ms per GB.  Called by the multi-GPU entry points; idempotent."""
    if _PEERS['on'] != bool(on):
        load().skm_scratch_local_only(1 if on else 0)
        _PEERS['on'] = bool(on)


def _note_process_group():
    """NCCL process group with more than one rank -> `uses_peer_gpus()` (checked when EM scratch is
    about to be taken: the group may be created long after this module was loaded)."""
    if _PEERS['on']:
        return
    import sys
    dist = sys.modules.get('torch.distributed')
    try:
        if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            uses_peer_gpus()
    except Exception:
        pass


class EmPlan:
    """Handle to a device-resident EM class structure (``skm_em_plan``): CSR by class + CSC by
    transcript, built once; the main EM and the bootstrap replicates run on it."""

    def __init__(self, handle):
        self._h = handle
        info = numpy.zeros(5, dtype='i8')
        check(load().skm_em_plan_info(self._h, _np_ptr(info)))
        self.n_classes, self.nnz, self.n_transcripts, self.device = (int(v) for v in info[:4])
        self.owns_counts = bool(info[4])

    @classmethod
    def from_csr(cls, class_ptr, class_tx, n_transcripts, counts=None, device=0, stream=None):
        """From host CSR arrays (int64[C + 1], int32[nnz]); integer `counts` make bootstraps possible
        without passing them again."""
        require_device()
        _note_process_group()
        class_ptr = numpy.ascontiguousarray(class_ptr, dtype='i8')
        class_tx = numpy.ascontiguousarray(class_tx, dtype='i4')
        if counts is not None:
            counts = numpy.ascontiguousarray(counts, dtype='i8')
        handle = ctypes.c_void_p()
        check(load().skm_em_plan_create(_np_ptr(class_ptr), _np_ptr(class_tx), class_ptr.shape[0] - 1,
                                        class_tx.shape[0], int(n_transcripts), _ptr(counts), 0, int(device),
                                        stream, ctypes.byref(handle)))
        return cls(handle)

    @classmethod
    def from_mapper(cls, mapper, n_transcripts=0, stream=None):
        """From a `DeviceMapper`'s dictionary, device to device, classes in first-seen order."""
        _note_process_group()
        handle = ctypes.c_void_p()
        check(load().skm_em_plan_from_mapper(mapper._h, int(n_transcripts), stream, ctypes.byref(handle)))
        return cls(handle)

    def run(self, eff_len, x0, counts=None, max_iters=0, stream=None):
        """`infer.em` for the rows of x0 (R, T); counts (R, C) fp64 or None = the plan's own (R = 1).
        Returns (x (R, T), iterations (R,))."""
        x0 = numpy.ascontiguousarray(numpy.atleast_2d(x0), dtype='f8')
        eff_len = numpy.ascontiguousarray(eff_len, dtype='f8')
        if counts is not None:
            counts = numpy.ascontiguousarray(numpy.atleast_2d(counts), dtype='f8')
        out = numpy.zeros_like(x0)
        iters = numpy.zeros(x0.shape[0], dtype='i4')
        check(load().skm_em_plan_run(self._h, _ptr(counts), _np_ptr(eff_len), _np_ptr(x0), x0.shape[0],
                                     int(max_iters), _np_ptr(out), _np_ptr(iters), 0, stream))
        return out, iters

    @staticmethod
    def run_many(plans, eff_lens, x0s, max_iters=0, stream=None):
        """`[p.run(l, x) for p, l, x in zip(plans, eff_lens, x0s)]` in ONE set of launches
        (`skm_em_plans_run`): plans of different samples on one device, each with its own counts.
        eff_lens, x0s: (P, T).  Returns (x (P, T), iterations (P,)), bit-identical to the loop."""
        x0s = numpy.ascontiguousarray(numpy.atleast_2d(x0s), dtype='f8')
        eff_lens = numpy.ascontiguousarray(numpy.atleast_2d(eff_lens), dtype='f8')
        if x0s.shape != eff_lens.shape or x0s.shape[0] != len(plans):
            raise ValueError('run_many: one row of lengths and of first guesses per plan')
        handles = (ctypes.c_void_p * len(plans))(*[p._h for p in plans])
        out = numpy.zeros_like(x0s)
        iters = numpy.zeros(len(plans), dtype='i4')
        check(load().skm_em_plans_run(ctypes.cast(handles, ctypes.c_void_p), len(plans), _np_ptr(eff_lens),
                                      _np_ptr(x0s), int(max_iters), _np_ptr(out), _np_ptr(iters), 0, stream))
        return out, iters

    def bootstrap(self, eff_len, x0, n_replicates, seed, first_replicate=0, counts=None, tpm=True, max_iters=0,
                  method=RESAMPLE_TREE, stream=None):
        """Resample + EM (+ TPM step) for replicates [first_replicate, first_replicate + n)."""
        x0 = numpy.ascontiguousarray(x0, dtype='f8')
        eff_len = numpy.ascontiguousarray(eff_len, dtype='f8')
        if counts is not None:
            counts = numpy.ascontiguousarray(counts, dtype='i8')
        out = numpy.zeros((n_replicates, x0.shape[0]), dtype='f8')
        iters = numpy.zeros(n_replicates, dtype='i4')
        check(load().skm_em_plan_bootstrap(self._h, _ptr(counts), _np_ptr(eff_len), _np_ptr(x0), int(n_replicates),
                                           int(first_replicate), int(seed) & (2 ** 64 - 1), int(method),
                                           int(max_iters), int(bool(tpm)), _np_ptr(out), _np_ptr(iters), 0, stream))
        return out, iters

    def close(self):
        if getattr(self, '_h', None):
            load().skm_em_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
