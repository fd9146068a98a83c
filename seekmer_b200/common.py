"""Host-side index object and FASTQ feeders — the `seekmer.common` surface on the infer path.

Mirrors (same names, arguments and error behaviour):
  KMerIndex                 `_common.pyx:19-48,268-313` + `_common.pxd:57-66`
  feed_single_ended_reads   `common.py:126-158`
  feed_pair_ended_reads     `common.py:161-197`
  decompress_and_open       `common.py:23-75`
  read_fasta                `common.py:78-105`

`KMerIndex` keeps the six public ndarray attributes of the reference object (the input
contract produced by `seekmer index`) and lazily uploads / re-lays them out into HBM the first
time a mapper needs them (`device_index`).  The lookups of `_common.pyx:54-266` themselves live
in `csrc/kmer.cuh` + `csrc/mapper.cu`.
"""
import bz2
import contextlib
import gzip
import io
import lzma
import os
import pathlib
import struct
import subprocess

import numpy

from . import _lib
from ._log import Logger

__all__ = ('BUFFER_SIZE', 'KMerIndex', 'decompress_and_open', 'read_fasta', 'iterate_by_group',
           'feed_single_ended_reads', 'feed_pair_ended_reads')

BUFFER_SIZE = 65536

_LOG = Logger(__name__)

_INDEX_VERSION = '2019.0.0'
IMAGE_SUFFIX = '.skmidx'  # the native GPU-layout index file (device image + transcript tables)
_IMAGE_VERSION = 1
# csrc/index.cu ImageHeader: magic, version, 5 layout constants, 7 counts, 4 sizes/offsets, trailer bytes, check
_IMAGE_HEADER = struct.Struct('<8s6I12qQ')


def _lib_default_device(index):
    return getattr(index, 'default_device', 0)
_EXTERNAL = {'.gz': ('zcat', gzip), '.bz2': ('bzcat', bz2), '.xz': ('xzcat', lzma),
             '.lzma': ('xzcat', lzma)}


class KMerIndex:
    """The Seekmer index: six numpy arrays plus a lazily created device image."""

    def __init__(self, kmers, contigs, sequences, targets, transcripts, exons):
        self.kmers = kmers
        self.contigs = contigs
        self.sequences = sequences
        self.targets = targets
        self.transcripts = transcripts
        self.exons = exons
        self._device = {}
        self._image_path = None  # a device image file (".skmidx") this index was loaded from

    # -- device image ---------------------------------------------------------------
    def device_index(self, device=0):
        """HBM-resident re-laid-out index for `device` (created once, then cached)."""
        dev = self._device.get(device)
        if dev is None:
            n_tx = len(self.transcripts) if self.transcripts is not None else 0
            if self.kmers is None and self._image_path is not None:
                dev = _lib.DeviceIndex.load(self._image_path, device=device)  # plain copies, no relayout
            else:
                dev = _lib.DeviceIndex(self.kmers, self.contigs, self.sequences, self.targets, n_tx,
                                       device=device)
            self._device[device] = dev
        return dev

    def release_device(self):
        for dev in self._device.values():
            dev.close()
        self._device = {}

    # -- persistence ----------------------------------------------------------------
    def save(self, path):
        """HDF5 when PyTables is importable (`_common.pyx:268-285` layout); `.npz` otherwise or
        when the suffix is `.npz`."""
        path = pathlib.Path(path)
        if path.suffix == IMAGE_SUFFIX:
            self._save_image(path)
            return
        if path.suffix != '.npz':
            try:
                import tables
            except ImportError:
                tables = None
            if tables is not None and hasattr(tables, '__version__'):
                filters = tables.Filters(complib='blosc', complevel=9, fletcher32=True)
                with tables.open_file(str(path), 'w', filters=filters) as f:
                    f.root._v_attrs['seekmer_version'] = _INDEX_VERSION
                    f.create_table('/', 'kmers', obj=self.kmers)
                    f.create_table('/', 'contigs', obj=self.contigs)
                    f.create_array('/', 'sequences', obj=self.sequences)
                    f.create_table('/', 'targets', obj=self.targets)
                    f.create_table('/', 'transcripts', obj=self.transcripts)
                    f.create_table('/', 'exons', obj=self.exons)
                _LOG.info('Saved index to "{}"', path)
                return
            raise RuntimeError('PyTables (HDF5) is not available; save to a ".npz" path instead')
        exons = self.exons if self.exons is not None else numpy.zeros(0, dtype='i4')
        with open(str(path), 'wb') as f:
            numpy.savez(f, seekmer_version=numpy.asarray(_INDEX_VERSION),
                        kmers=numpy.asarray(self.kmers), contigs=numpy.asarray(self.contigs),
                        sequences=numpy.asarray(self.sequences), targets=numpy.asarray(self.targets),
                        transcripts=numpy.asarray(self.transcripts), exons=numpy.asarray(exons))
        _LOG.info('Saved index to "{}"', path)

    def _save_image(self, path):
        """The native GPU-layout index file: the device image as it lies in HBM
        (`skm_index_save`) + the host-side tables `infer` needs (transcripts, exons) as a trailer.
        `KMerIndex.load` of such a file uploads with plain copies: no relayout, no link probes."""
        buf = io.BytesIO()
        exons = self.exons if self.exons is not None else numpy.zeros(0, dtype='i4')
        numpy.savez(buf, seekmer_version=numpy.asarray(_INDEX_VERSION),
                    transcripts=numpy.asarray(self.transcripts), exons=numpy.asarray(exons))
        self.device_index(_lib_default_device(self)).save(path, buf.getvalue())
        _LOG.info('Saved index to "{}"', path)

    @classmethod
    def _load_image(cls, path):
        with open(str(path), 'rb') as f:
            raw = f.read(_IMAGE_HEADER.size)
            if len(raw) < _IMAGE_HEADER.size or raw[:8] != b'SKMB200\0':
                raise RuntimeError('not a seekmer_b200 device image: %s' % path)
            fields = _IMAGE_HEADER.unpack(raw)
            if fields[1] != _IMAGE_VERSION:
                raise RuntimeError('invalid index version.')
            table_bytes, hot_bytes, trailer_bytes = fields[14], fields[15], fields[18]
            f.seek(_IMAGE_HEADER.size + table_bytes + hot_bytes)
            trailer = f.read(trailer_bytes)
        with numpy.load(io.BytesIO(trailer), allow_pickle=False) as z:
            if str(z['seekmer_version']) != _INDEX_VERSION:
                raise RuntimeError('invalid index version.')
            transcripts, exons = z['transcripts'], z['exons']
        index = cls(None, None, None, None, transcripts, exons)
        index._image_path = path
        _LOG.info('Loaded index from "{}"', path)
        return index

    @classmethod
    def load(cls, path):
        path = pathlib.Path(path)
        if path.suffix == IMAGE_SUFFIX:
            return cls._load_image(path)
        if path.suffix == '.npz':
            with numpy.load(str(path), allow_pickle=False) as z:
                if str(z['seekmer_version']) != _INDEX_VERSION:
                    raise RuntimeError('invalid index version.')
                parts = [z[k] for k in ('kmers', 'contigs', 'sequences', 'targets', 'transcripts',
                                        'exons')]
            _LOG.info('Loaded index from "{}"', path)
            return cls(*parts)
        try:
            import tables
            if not hasattr(tables, '__version__'):  # a stand-in module without HDF5 behind it
                raise ImportError('not PyTables')
        except ImportError as exc:
            raise RuntimeError('reading an HDF5 index needs PyTables, which is not installed; '
                               'convert the index to ".npz" where it is') from exc
        with tables.open_file(str(path), 'r') as f:
            if f.root._v_attrs['seekmer_version'] != _INDEX_VERSION:
                raise RuntimeError('invalid index version.')
            parts = [f.get_node('/' + k).read() for k in ('kmers', 'contigs', 'sequences',
                                                          'targets', 'transcripts', 'exons')]
        _LOG.info('Loaded index from "{}"', path)
        return cls(*parts)


@contextlib.contextmanager
def decompress_and_open(path):
    """Binary read handle; `.gz/.bz2/.xz/.lzma` go through the external decompressor when it
    can be started, else through the stdlib module."""
    path = pathlib.Path(path)
    tool = _EXTERNAL.get(path.suffix)
    if tool is None:
        with path.open('rb') as f:
            yield f
        return
    command, module = tool
    try:
        process = subprocess.Popen([command, str(path)], stdout=subprocess.PIPE)
    except OSError:
        _LOG.warn('Unable to call {}, falling back to the {} module.', command, module.__name__)
        with module.open(str(path), 'rb') as raw, io.BufferedReader(raw) as f:
            yield f
        return
    with process:
        yield process.stdout


def read_fasta(path):
    name, chunks = None, []
    with decompress_and_open(path) as f:
        for line in f:
            if line[:1] == b'>':
                if name is not None:
                    yield name, b''.join(chunks)
                name, chunks = line[1:].strip(), []
            else:
                chunks.append(line.strip())
    if name is not None:
        yield name, b''.join(chunks)


def iterate_by_group(iterator, group_size):
    return zip(*([iter(iterator)] * group_size))


def _fastq_records(handle):
    """(name, sequence) per 4-line FASTQ record; a trailing partial record yields what it has,
    like the reference's line-index logic."""
    name = None
    for i, line in enumerate(handle):
        phase = i & 3
        if phase == 0:
            name = line.strip()[1:]
        elif phase == 1:
            yield name, line.strip()


class FastqSource:
    """What `feed_single_ended_reads` / `feed_pair_ended_reads` return: iterating it yields the
    reference's `(count, names, reads)` batches (`common.py:126-197`), and it remembers the
    file paths, so that `ReadMapper` can instead hand the raw FASTQ text to the GPU
    (`skm_map_fastq`) when nobody needs the read names."""

    def __init__(self, paths, paired):
        self.paths = [pathlib.Path(p) for p in paths]
        self.paired = paired
        if paired and len(self.paths) % 2 != 0:
            raise ValueError('cannot process odd numbers of pair-ended files')

    def __iter__(self):
        return (_iterate_pair_ended if self.paired else _iterate_single_ended)(*self.paths)

    def __next__(self):  # the reference's feeders are generators: next(feeder) must work
        if getattr(self, '_gen', None) is None:
            self._gen = iter(self)
        return next(self._gen)

    def text_chunks(self, chunk_bytes):
        """Yield `(buf1, n1, buf2, n2, eof)`: raw text of the file (pair) in reusable uint8
        buffers, each call handing over `chunk_bytes` more per file.  The consumer reports how
        much it used through `.consumed(c1, c2)` before asking for the next chunk; the rest is
        kept in front of the new data."""
        files = iterate_by_group(self.paths, 2) if self.paired else [(p,) for p in self.paths]
        for group in files:
            with contextlib.ExitStack() as stack:
                handles = [stack.enter_context(decompress_and_open(p)) for p in group]
                carries = [_Carry(chunk_bytes, p.stat().st_size if p.suffix not in _EXTERNAL else None)
                           for p in group]
                self._carries = carries
                while True:
                    if len(carries) > 1 and carries[0].size_hint is None:
                        list(_read_pool().map(lambda ch: ch[0].fill(ch[1]), zip(carries, handles)))  # two pipes
                    else:
                        for c, h in zip(carries, handles):
                            c.fill(h)
                    eof = all(c.eof for c in carries)
                    if eof:
                        for c in carries:
                            c.finish()
                    yield tuple(x for c in carries for x in (c.buf, c.n)) + (eof,)
                    if eof:
                        break
                for c in carries:
                    _release_bytes(c.buf)
                    c.buf = None

    def consumed(self, *used):
        for c, u in zip(self._carries, used):
            c.drop(u)


class _Carry:
    """A growing text window over one FASTQ stream: unread tail + freshly read bytes."""

    def __init__(self, chunk_bytes, size_hint=None):
        self.size_hint = size_hint
        if size_hint is not None:  # a regular file: no need for more than it holds
            chunk_bytes = max(min(chunk_bytes, size_hint + 16), 1 << 16)
        self.chunk = chunk_bytes
        self.buf = _pinned_bytes(2 * chunk_bytes + 64)
        self.n = 0
        self.eof = False

    def fill(self, handle):
        if self.eof:
            return
        if self.n + self.chunk > self.buf.shape[0]:  # a record longer than a chunk: grow
            grown = _pinned_bytes(2 * (self.n + self.chunk) + 64)
            grown[:self.n] = self.buf[:self.n]
            _release_bytes(self.buf)
            self.buf = grown
        want = self.chunk
        view = memoryview(self.buf)[self.n:self.n + want]
        got = self._fill_regular_file(handle, view) if self.size_hint is not None else None
        if got is None:
            got = 0
            while got < want:
                k = handle.readinto(view[got:])
                if not k:
                    self.eof = True
                    break
                got += k
        elif got < want:
            self.eof = True
        self.n += got

    def _fill_regular_file(self, handle, view):
        """A plain file is read in slices by several threads (os.preadv releases the GIL): one
        Python thread copies ~1.5 GB/s out of the page cache, the PCIe link takes 50."""
        try:
            fd = handle.fileno()
            pos = handle.tell()
        except (OSError, AttributeError, io.UnsupportedOperation):
            return None
        want = min(len(view), max(self.size_hint - pos, 0))
        if want <= 0:
            return 0
        n_slices = max(1, min(_READ_THREADS, want >> 22))
        step = (want + n_slices - 1) // n_slices

        def read_slice(i):
            lo, hi = i * step, min((i + 1) * step, want)
            done = lo
            while done < hi:
                k = os.preadv(fd, [view[done:hi]], pos + done)
                if k <= 0:
                    break
                done += k
            return done - lo

        if n_slices == 1:
            got = read_slice(0)
        else:
            counts = list(_read_pool().map(read_slice, range(n_slices)))
            got = 0
            for i, k in enumerate(counts):  # a short slice ends the valid prefix
                got += k
                if k < min((i + 1) * step, want) - i * step:
                    break
        handle.seek(pos + got)
        return got

    def finish(self):
        """End of stream: terminate the last line, and complete a trailing partial record that
        has its sequence line (the reference's line-index logic yields it, `common.py:137-138`)."""
        if self.n and self.buf[self.n - 1] != 10:
            self.buf[self.n] = 10
            self.n += 1
        lines = int(numpy.count_nonzero(self.buf[:self.n] == 10))
        extra = (-lines) % 4
        if extra in (1, 2):  # 3 or 2 lines of the last record are there: header + sequence (+ '+')
            self.buf[self.n:self.n + extra] = 10
            self.n += extra
        elif extra == 3:  # a lone header line: no read
            nl = numpy.flatnonzero(self.buf[:self.n - 1] == 10)
            self.n = int(nl[-1]) + 1 if nl.size else 0

    def drop(self, used):
        rest = self.n - used
        if used and rest:
            self.buf[:rest] = self.buf[used:self.n].copy()
        self.n = rest


_READ_THREADS = 8
_POOL = None


def _read_pool():
    global _POOL
    if _POOL is None:
        import concurrent.futures
        _POOL = concurrent.futures.ThreadPoolExecutor(max_workers=_READ_THREADS, thread_name_prefix='skm-read')
    return _POOL


def _pinned_bytes(n):
    """uint8 host buffer, page-locked when torch can provide it (faster H2D copies).  Page-locking
    costs ~0.4 ms per MB, so buffers are handed back (`_release_bytes`) and reused."""
    for i, (a, _) in enumerate(_FREE_BUFFERS):
        if n <= a.shape[0] <= 2 * n + (1 << 20):
            return _FREE_BUFFERS.pop(i)[0]
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            a = t.numpy()
            _OWNERS[a.ctypes.data] = t
            return a
    except Exception:
        pass
    return numpy.empty(n, dtype='u1')


def _release_bytes(a):
    if len(_FREE_BUFFERS) < 8:
        _FREE_BUFFERS.append((a, _OWNERS.get(a.ctypes.data)))
    else:
        _OWNERS.pop(a.ctypes.data, None)


_FREE_BUFFERS = []
_OWNERS = {}


def feed_single_ended_reads(*paths):
    """Single-ended FASTQ file(s) as one sample (`common.py:126-158`)."""
    return FastqSource(paths, paired=False)


def feed_pair_ended_reads(*paths):
    """Pair-ended FASTQ file pairs as one sample, mates interleaved (`common.py:161-197`)."""
    return FastqSource(paths, paired=True)


def _iterate_single_ended(*paths):
    """Yield `(count, names, reads)` batches of up to BUFFER_SIZE reads; all files form one
    sample."""
    names, reads = [], []
    for path in paths:
        with decompress_and_open(path) as f:
            for name, seq in _fastq_records(f):
                names.append(name)
                reads.append(seq)
                if len(names) >= BUFFER_SIZE:
                    yield len(names), names, reads
                    names, reads = [], []
    if reads:
        yield len(names), names, reads
    _LOG.debug('Finished reading sequence file(s)')


def _iterate_pair_ended(*paths):
    """Yield `(pair_count, names, reads)` with mates interleaved (2i, 2i+1)."""
    if len(paths) % 2 != 0:
        raise ValueError('cannot process odd numbers of pair-ended files')
    names, reads = [], []
    for path1, path2 in iterate_by_group(paths, 2):
        with decompress_and_open(path1) as f1, decompress_and_open(path2) as f2:
            for (name, seq1), (_, seq2) in zip(_fastq_records(f1), _fastq_records(f2)):
                names.append(name)
                reads.append(seq1)
                reads.append(seq2)
                if len(names) >= BUFFER_SIZE:
                    yield len(names), names, reads
                    names, reads = [], []
    if reads:
        yield len(names), names, reads
    _LOG.debug('Finished reading sequence file(s)')
