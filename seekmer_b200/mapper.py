"""Mapping orchestration and result model — the `seekmer.mapper` / `seekmer._mapper` surface.

Mirrors (same names, argument meaning and error behaviour):
  MAX_FRAGMENT_LENGTH, ReadMapper   `_mapper.pyx:18-20,31-105`
  MapResult, SummarizedResult       `mapper.py:18-145`
  map_reads                         `mapper.py:148-193`
  map_multiple_samples              `mapper.py:196-234`

What changes underneath: `ReadMapper.__call__` hands every feeder batch to the CUDA mapper
(`skm_map_batch`), which keeps the class dictionary and the fragment-length histogram on the
device; the `collections.Counter` of the reference is filled once per `__call__` from the
exported dictionary instead of once per read.  Only `-m/--save-readmap` (a `readmap` file)
needs per-read tuples and takes the per-read output path.  There is no CPU mapping fallback.
"""
import collections
import threading

import numpy

from . import _lib, common
from ._log import Logger

__all__ = ('MAX_FRAGMENT_LENGTH', 'MapResult', 'ReadMapper', 'SummarizedResult', 'map_reads',
           'map_multiple_samples')

_LOG = Logger(__name__)

MAX_FRAGMENT_LENGTH = _lib.MAX_FRAGMENT_LENGTH

EPS = numpy.finfo('f4').eps


class SummarizedResult:
    __slots__ = ['aligned', 'unaligned', 'total', 'class_map', 'class_count',
                 'fragment_length_frequencies', 'effective_lengths']

    def __init__(self, aligned, unaligned, total, class_map, class_count,
                 fragment_length_frequencies, effective_lengths):
        self.aligned = aligned
        self.unaligned = unaligned
        self.total = total
        self.class_map = class_map
        self.class_count = class_count
        self.fragment_length_frequencies = fragment_length_frequencies
        self.effective_lengths = effective_lengths


class MapResult:
    """A mapping result collection with a lock (`mapper.py:40-145`)."""

    def __init__(self, index, readmap=None):
        self.lock = threading.Lock()
        self.counter = collections.Counter()
        self.index = index
        self.readmap = readmap
        self.fragment_length_counts = numpy.zeros(MAX_FRAGMENT_LENGTH, dtype='i8')
        self._table = None  # exported device dictionary, valid while the counter mirrors it

    def update(self, read_names, iterable):
        """Add per-read mapping results (list of ordered id tuples)."""
        self._table = None
        self.counter.update(iterable)
        if self.readmap is not None:
            for read_name, targets in zip(read_names, iterable):
                ids = self.index.transcripts[targets,]['transcript_id']
                print(read_name.decode(), *[id_.decode() for id_ in ids], sep='\t',
                      file=self.readmap)

    def update_counts(self, classes):
        """Add already tallied classes: iterable of (ordered id tuple, count) in first-seen
        order — what the device dictionary exports."""
        self._table = None
        counter = self.counter
        for key, count in classes:
            counter[key] += count

    def summarize(self):
        if self._table is not None:
            return summarize_table(self._table, self)
        unaligned = self.counter.pop((), 0)
        n = len(self.counter)
        class_count = numpy.fromiter(self.counter.values(), dtype='f8', count=n)
        sizes = numpy.fromiter((len(k) for k in self.counter), dtype='i8', count=n)
        nnz = int(sizes.sum())
        flat = numpy.fromiter((t for k in self.counter for t in k), dtype='i8', count=nnz)
        self.counter[()] = unaligned
        if nnz:
            class_map = numpy.stack([numpy.repeat(numpy.arange(n, dtype='i8'), sizes), flat])
        else:
            class_map = numpy.asarray([]).T  # `numpy.asarray([]).T` of the reference: size 0
        aligned = class_count.sum()
        return SummarizedResult(
            aligned=int(aligned),
            unaligned=int(unaligned),
            total=int(aligned + unaligned),
            class_map=class_map,
            class_count=class_count,
            fragment_length_frequencies=self.fragment_length_counts,
            effective_lengths=self.effective_lengths,
        )

    def merge_fragment_lengths(self, fragment_length_counts):
        self.fragment_length_counts += fragment_length_counts

    @property
    def harmonic_mean_fragment_length(self):
        assert self.fragment_length_counts[0] == 0
        numerator = self.fragment_length_counts.sum()
        if numerator == 0:
            return 0
        denominator = (self.fragment_length_counts[1:].astype('f8')
                       / numpy.arange(1, MAX_FRAGMENT_LENGTH)).sum()
        return numerator / denominator

    @property
    def effective_lengths(self):
        """fp64 on the device, same accumulation order as `mapper.py:134-141`."""
        length = numpy.ascontiguousarray(self.index.transcripts['length'], dtype='f8')
        out = numpy.zeros(length.shape, dtype='f8')
        if length.size == 0:
            return out
        fld = numpy.ascontiguousarray(self.fragment_length_counts, dtype='i8')
        _lib.require_device()
        _lib.check(_lib.load().skm_effective_lengths(
            _lib._np_ptr(fld), _lib._np_ptr(length), length.shape[0], _lib._np_ptr(out), 0,
            _device_of(self.index), None))
        return out

    def clear(self):
        self._table = None
        self.counter.clear()


def summarize_table(table, map_result):
    """`MapResult.summarize` (`mapper.py:77-104`) straight from an exported class table
    (classes already in first-seen order): no per-class Python objects."""
    off = numpy.asarray(table['key_offsets'], dtype='i8')
    sizes = off[1:] - off[:-1]
    n = sizes.shape[0]
    class_count = numpy.asarray(table['counts'], dtype='f8')
    if int(off[-1]):
        class_map = numpy.stack([numpy.repeat(numpy.arange(n, dtype='i8'), sizes),
                                 numpy.asarray(table['key_ids'], dtype='i8')])
    else:
        class_map = numpy.asarray([]).T
    aligned = class_count.sum()
    unaligned = int(table['unaligned'])
    return SummarizedResult(
        aligned=int(aligned), unaligned=unaligned, total=int(aligned + unaligned), class_map=class_map,
        class_count=class_count, fragment_length_frequencies=map_result.fragment_length_counts,
        effective_lengths=map_result.effective_lengths)


FASTQ_CHUNK_BYTES = 128 << 20  # raw text handed to the GPU per file and call


def _device_of(index):
    return getattr(index, 'default_device', 0)


def _pack_batch(reads):
    """list[bytes] -> (uint8 bases, int64 offsets or None, fixed_len, max_len)."""
    n = len(reads)
    lens = numpy.fromiter(map(len, reads), dtype='i8', count=n)
    bases = numpy.frombuffer(b''.join(reads), dtype='u1')
    lo, hi = (int(lens.min()), int(lens.max())) if n else (0, 0)
    if lo == hi and lo > 0:
        return bases, None, lo, hi
    offsets = numpy.zeros(n + 1, dtype='i8')
    numpy.cumsum(lens, out=offsets[1:])
    return bases, offsets, 0, hi


class ReadMapper:
    """A read mapper bound to one index and one result collection (`_mapper.pyx:31-105`).

    One instance drives one CUDA mapper handle (class dictionary + FLD on the device); like
    the reference's, an instance is used by one thread at a time.
    """

    def __init__(self, index, map_result, device=None, class_capacity=0, id_capacity=0,
                 device_mapper=None):
        self.index = index
        self.map_result = map_result
        self.device = _device_of(index) if device is None else device
        self._class_capacity = class_capacity
        self._id_capacity = id_capacity
        # a caller-owned handle to run on (reset, used, left open): creating the class
        # dictionary costs far more than mapping a small sample (`map_multiple_samples`)
        self._shared = device_mapper
        self.fragment_length_counts = numpy.zeros(MAX_FRAGMENT_LENGTH, dtype='i8')

    def __call__(self, reads_iterator):
        """Run the mapping loop over `(read_count, read_names, reads)` batches."""
        if self._shared is not None:
            mapper = self._shared
            mapper.reset()
        else:
            dev_index = self.index.device_index(self.device)
            mapper = _lib.DeviceMapper(dev_index, self._class_capacity, self._id_capacity)
        want_reads = self.map_result.readmap is not None
        try:
            first_unit = 0
            if isinstance(reads_iterator, common.FastqSource) and not want_reads:
                # nobody needs the read names: the raw FASTQ text goes to the GPU and is parsed
                # there (skm_map_fastq) instead of line by line in Python
                for chunk in reads_iterator.text_chunks(FASTQ_CHUNK_BYTES):
                    if reads_iterator.paired:
                        b1, n1, b2, n2, eof = chunk
                        units, c1, c2 = mapper.map_fastq(b1, n1, b2, n2, first_unit=first_unit)
                        reads_iterator.consumed(c1, c2)
                        if eof and (c1 < n1 or c2 < n2):
                            _LOG.debug('Mate files differ in length; surplus reads ignored (zip semantics).')
                    else:
                        b1, n1, eof = chunk
                        units, c1, _ = mapper.map_fastq(b1, n1, first_unit=first_unit)
                        reads_iterator.consumed(c1)
                    first_unit += units
                    _LOG.debug('Mapped {} reads.', units)
                reads_iterator = ()
            for read_count, read_names, reads in reads_iterator:
                single_ended = read_count == len(reads)  # `_mapper.pyx:75`
                bases, offsets, fixed_len, max_len = _pack_batch(reads)
                out_class, _ = mapper.map_batch(bases, offsets, read_count, not single_ended,
                                                first_unit=first_unit, fixed_len=fixed_len,
                                                max_len=max_len, per_read=want_reads)
                if want_reads:
                    table = mapper.export(with_slots=True)
                    lookup = _tuples_by_slot(table)
                    ids = [lookup[s] if s >= 0 else () for s in out_class.tolist()]
                    with self.map_result.lock:
                        _write_readmap(self.map_result, read_names, ids)
                first_unit += read_count
                _LOG.debug('Mapped {} reads.', read_count)
            table = mapper.export()
        finally:
            if self._shared is None:
                mapper.close()
        if table.get('short_reads'):
            # undefined in the reference (`_kmer.pxd:46-68` reads past the end of such a read)
            _LOG.warn('{} reads are shorter than k={}: their units were counted as unaligned.',
                      table['short_reads'], _lib.K)
        classes = _class_tuples(table)
        with self.map_result.lock:
            fresh = not self.map_result.counter
            self.map_result.update_counts(classes)
            if table['unaligned']:
                self.map_result.counter[()] += table['unaligned']
            if fresh:
                self.map_result._table = table
        self.fragment_length_counts += table['fld']
        with self.map_result.lock:
            self.map_result.merge_fragment_lengths(self.fragment_length_counts)


def _class_tuples(table):
    off = table['key_offsets'].tolist()
    ids = table['key_ids'].tolist()
    counts = table['counts'].tolist()
    return [(tuple(ids[off[i]:off[i + 1]]), counts[i]) for i in range(len(counts))]


def _tuples_by_slot(table):
    off = table['key_offsets'].tolist()
    ids = table['key_ids'].tolist()
    return {s: tuple(ids[off[i]:off[i + 1]]) for i, s in enumerate(table['slots'].tolist())}


def _write_readmap(map_result, read_names, ids):
    transcripts = map_result.index.transcripts
    for read_name, targets in zip(read_names, ids):
        names = transcripts[targets,]['transcript_id']
        print(read_name.decode(), *[n.decode() for n in names], sep='\t', file=map_result.readmap)


def map_reads(index, read_feeder, job_count=1, readmap=None, debug=False):
    """Map reads (`mapper.py:148-193`).

    `job_count` was the reference's number of host mapper threads; the GPU mapper needs one
    (a single `ReadMapper` saturates the device), so it is accepted and ignored.  `debug` keeps
    its only observable meaning (run in the calling thread) trivially.
    """
    map_result = MapResult(index, readmap)
    try:
        ReadMapper(index, map_result)(read_feeder)
    finally:
        if readmap is not None:
            readmap.close()
    return map_result


def map_multiple_samples(index, read_feeders, job_count=1, debug=False):
    """One `MapResult` per sample (`mapper.py:196-234`).  The samples go through one device
    mapper, reset in between: a cell of a single-cell run maps in under a millisecond, while
    allocating a class dictionary takes tens."""
    map_results = []
    shared = None
    try:
        for read_feeder in read_feeders:
            result = MapResult(index)
            map_results.append(result)
            if shared is None:
                shared = _lib.DeviceMapper(index.device_index(_device_of(index)), 0, 0)
            ReadMapper(index, result, device_mapper=shared)(read_feeder)
    finally:
        if shared is not None:
            shared.close()
    return map_results
