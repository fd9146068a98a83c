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
    """`mapper.py:18-37`, plus `plan`: the same class structure already resident on a GPU
    (`_lib.EmPlan`, classes in `class_map` order, integer counts included) when the result comes
    straight from the device mapper; `infer.quantify` then skips the host round trip."""
    __slots__ = ['aligned', 'unaligned', 'total', 'class_map', 'class_count',
                 'fragment_length_frequencies', 'effective_lengths', 'plan']

    def __init__(self, aligned, unaligned, total, class_map, class_count,
                 fragment_length_frequencies, effective_lengths, plan=None):
        self.aligned = aligned
        self.unaligned = unaligned
        self.total = total
        self.class_map = class_map
        self.class_count = class_count
        self.fragment_length_frequencies = fragment_length_frequencies
        self.effective_lengths = effective_lengths
        self.plan = plan


class MapResult:
    """A mapping result collection with a lock (`mapper.py:40-145`)."""

    def __init__(self, index, readmap=None):
        self.lock = threading.Lock()
        self._counter = collections.Counter()
        self.index = index
        self.readmap = readmap
        self.fragment_length_counts = numpy.zeros(MAX_FRAGMENT_LENGTH, dtype='i8')
        self._table = None    # exported device dictionary, valid while the counter mirrors it
        self._plan = None     # the same classes as an EM plan on the device (`_lib.EmPlan`)
        self._pending = False  # the table has not been poured into the Counter yet

    @property
    def counter(self):
        """`collections.Counter`: ordered id tuple -> count, `()` = unaligned (`mapper.py:54`).  A
        result that came from the device dictionary fills it on first access: 10^6 Python tuples
        are only built when somebody asks for them."""
        if self._pending:
            self._pending = False
            counter = self._counter
            for key, count in _class_tuples(self._table):
                counter[key] += count
            if self._table['unaligned']:
                counter[()] += self._table['unaligned']
        return self._counter

    @counter.setter
    def counter(self, value):
        self._invalidate()
        self._counter = value

    def _adopt_table(self, table, plan):
        self._table = table
        self._plan = plan
        self._pending = True

    def _invalidate(self):
        if self._pending:
            self.counter  # noqa: B018 - pour the table in before it stops being the truth
        self._table = None
        if self._plan is not None:
            self._plan.close()
            self._plan = None

    def update(self, read_names, iterable):
        """Add per-read mapping results (list of ordered id tuples)."""
        self._invalidate()
        self.counter.update(iterable)
        if self.readmap is not None:
            for read_name, targets in zip(read_names, iterable):
                ids = self.index.transcripts[targets,]['transcript_id']
                print(read_name.decode(), *[id_.decode() for id_ in ids], sep='\t',
                      file=self.readmap)

    def update_counts(self, classes):
        """Add already tallied classes: iterable of (ordered id tuple, count) in first-seen
        order — what the device dictionary exports."""
        self._invalidate()
        counter = self.counter
        for key, count in classes:
            counter[key] += count

    def summarize(self):
        if self._table is not None:
            return summarize_table(self._table, self, self._plan)
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
        self._pending = False
        self._invalidate()
        self._counter.clear()


def summarize_table(table, map_result, plan=None):
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
        effective_lengths=map_result.effective_lengths, plan=plan)


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
        try:
            _feed(mapper, reads_iterator, 0, self.map_result)
            table, plan = _finish_mapper(mapper, self.index)
        finally:
            if self._shared is None:
                mapper.close()
        _deliver(self, table, plan)


def _feed(mapper, reads_iterator, first_unit, map_result):
    """Everything `reads_iterator` yields goes through `mapper`; returns the units mapped."""
    want_reads = map_result.readmap is not None
    start = first_unit
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
        return first_unit - start
    for read_count, read_names, reads in reads_iterator:
        first_unit += _map_one_batch(mapper, read_count, read_names, reads, first_unit, map_result)
    return first_unit - start


def _map_one_batch(mapper, read_count, read_names, reads, first_unit, map_result):
    want_reads = map_result.readmap is not None
    single_ended = read_count == len(reads)  # `_mapper.pyx:75`
    bases, offsets, fixed_len, max_len = _pack_batch(reads)
    out_class, _ = mapper.map_batch(bases, offsets, read_count, not single_ended,
                                    first_unit=first_unit, fixed_len=fixed_len,
                                    max_len=max_len, per_read=want_reads)
    if want_reads:
        table = mapper.export(with_slots=True)
        lookup = _tuples_by_slot(table)
        ids = [lookup[s] if s >= 0 else () for s in out_class.tolist()]
        with map_result.lock:
            _write_readmap(map_result, read_names, ids)
    _LOG.debug('Mapped {} reads.', read_count)
    return read_count


def _finish_mapper(mapper, index):
    """Host copy of the dictionary (for `MapResult.counter` / `summarize`) and, while the
    dictionary still sits in HBM, the EM's class structure made from it device to device."""
    table = mapper.export()
    plan = None
    n_tx = len(index.transcripts) if getattr(index, 'transcripts', None) is not None else 0
    if table['counts'].shape[0] and n_tx:
        plan = _lib.EmPlan.from_mapper(mapper, n_tx)
    if table.get('short_units'):
        # undefined in the reference (`_kmer.pxd:46-68` reads past the end of such a read)
        _LOG.warn('{} reads or pairs hold a read shorter than k={}: they were counted as unaligned.',
                  table['short_units'], _lib.K)
    return table, plan


def _deliver(read_mapper, table, plan):
    """What the end of `ReadMapper.__call__` does in the reference (`_mapper.pyx:100-105`): counts
    into the shared `MapResult` under its lock, then the fragment lengths."""
    map_result = read_mapper.map_result
    with map_result.lock:
        fresh = map_result._table is None and not map_result._counter
        if fresh:
            map_result._adopt_table(table, plan)  # the Counter is filled when somebody looks at it
        else:
            map_result.update_counts(_class_tuples(table))
            if table['unaligned']:
                map_result.counter[()] += table['unaligned']
            if plan is not None:
                plan.close()
    read_mapper.fragment_length_counts += table['fld']
    with map_result.lock:
        map_result.merge_fragment_lengths(read_mapper.fragment_length_counts)


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


def _devices_for(job_count):
    """`-j/--jobs` was the reference's host thread count (`infer.py:341-343`); here it is the
    number of GPUs of this box to spread the work over."""
    try:
        n = int(job_count)
    except (TypeError, ValueError):
        n = 1
    return list(range(max(1, min(n, _lib.device_count()))))


def map_reads(index, read_feeder, job_count=1, readmap=None, debug=False):
    """Map reads (`mapper.py:148-193`).

    `job_count` (the reference's mapper threads) is the number of GPUs to use: with N > 1 the
    feeder's work is spread over N devices (whole FASTQ file groups, or feeder batches pulled
    from the shared iterator as the reference's threads do), every device fills its own class
    dictionary, and the dictionaries are merged on the first device (`skm_classes_merge`,
    first-seen order kept by global unit indices).  A readmap (`-m`) needs the reads in file
    order and `debug` means "in the calling thread" (`mapper.py:171-172`): both run on one GPU.
    """
    map_result = MapResult(index, readmap)
    devices = _devices_for(job_count)
    try:
        if len(devices) == 1 or readmap is not None or debug:
            ReadMapper(index, map_result)(read_feeder)
        else:
            _map_reads_on_devices(index, read_feeder, devices, map_result)
    finally:
        if readmap is not None:
            readmap.close()
    return map_result


_GROUP_SHIFT = 40  # unit indices of FASTQ file group g start at g << 40: first-seen order = file order


def _map_reads_on_devices(index, read_feeder, devices, map_result):
    """One host thread and one device mapper per GPU (`mapper.py:174-189` with devices in place
    of threads)."""
    import torch
    _lib.uses_peer_gpus()  # the dictionaries are merged by peer copies: EM scratch must stay unmapped there
    work, shared_iter = None, None
    if isinstance(read_feeder, common.FastqSource):
        size = 2 if read_feeder.paired else 1
        groups = [read_feeder.paths[i:i + size] for i in range(0, len(read_feeder.paths), size)]
        work = [(g << _GROUP_SHIFT, common.FastqSource(paths, read_feeder.paired)) for g, paths in enumerate(groups)]
        devices = devices[:max(1, len(work))]
    else:
        shared_iter = iter(read_feeder)
    pull = threading.Lock()
    state = {'first_unit': 0, 'next': 0}
    mappers = [_lib.DeviceMapper(index.device_index(d), 0, 0) for d in devices]
    errors = []

    def run(mapper):
        try:
            while True:
                with pull:  # the reference's threads pull from the shared feeder the same way
                    if work is not None:
                        if state['next'] >= len(work):
                            return
                        first, source = work[state['next']]
                        state['next'] += 1
                        batch = None
                    else:
                        batch = next(shared_iter, None)
                        if batch is None:
                            return
                        first = state['first_unit']
                        state['first_unit'] += batch[0]
                if batch is None:
                    _feed(mapper, source, first, map_result)
                else:
                    _map_one_batch(mapper, batch[0], batch[1], batch[2], first, map_result)
        except BaseException as exc:  # noqa: BLE001 - re-raised in the calling thread
            errors.append(exc)

    try:
        threads = [threading.Thread(target=run, args=(m,)) for m in mappers]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        # every other device's dictionary -> the first device, over NVLink, merged there
        dev0 = torch.device('cuda', devices[0])
        for m in mappers[1:]:
            t = m.export_raw_torch()
            with torch.cuda.device(dev0):
                moved = {k: (v.to(dev0) if hasattr(v, 'to') else v) for k, v in t.items()}
                torch.cuda.current_stream(dev0).synchronize()
                mappers[0].merge_device(moved['key_offsets'], moved['key_ids'], moved['counts'], moved['first_unit'],
                                        moved['fld'], moved['unaligned'])
        table, plan = _finish_mapper(mappers[0], index)
    finally:
        for m in mappers:
            m.close()
    rm = ReadMapper(index, map_result, device=devices[0])
    _deliver(rm, table, plan)


def map_multiple_samples(index, read_feeders, job_count=1, debug=False):
    """One `MapResult` per sample (`mapper.py:196-234`).  Samples are dealt round-robin to
    `job_count` GPUs (one host thread and one device mapper each, reset between samples: a cell
    of a single-cell run maps in under a millisecond, allocating a class dictionary takes tens)."""
    read_feeders = list(read_feeders)
    map_results = [MapResult(index) for _ in read_feeders]
    devices = _devices_for(1 if debug else job_count)[:max(1, len(read_feeders))]
    errors = []

    def run(device, mine):
        shared = None
        try:
            for i in mine:
                if shared is None:
                    shared = _lib.DeviceMapper(index.device_index(device), 0, 0)
                ReadMapper(index, map_results[i], device=device, device_mapper=shared)(read_feeders[i])
        except BaseException as exc:  # noqa: BLE001 - re-raised in the calling thread
            errors.append(exc)
        finally:
            if shared is not None:
                shared.close()

    shares = [list(range(k, len(read_feeders), len(devices))) for k in range(len(devices))]
    if len(devices) == 1:
        run(devices[0], shares[0])
    else:
        _lib.uses_peer_gpus()
        for d in devices:  # upload the index replicas one after the other (not thread-safe per object)
            index.device_index(d)
        threads = [threading.Thread(target=run, args=(d, s)) for d, s in zip(devices, shares)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
    if errors:
        raise errors[0]
    return map_results
