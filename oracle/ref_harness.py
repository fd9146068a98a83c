"""Load the compiled reference (``oracle/_ref``) as the parity oracle.

TEST INFRASTRUCTURE ONLY.  Nothing in ``seekmer_b200/`` imports this module.

``load_ref()`` returns the ``seekmer`` package whose native modules are the
unmodified reference ``.pyx`` files compiled by ``oracle/build_ref.py``.  When
``/root/reference`` is present (this container) the reference's pure-Python
modules (``mapper``, ``infer``, ``index_builder``, ``common``) are importable too,
straight from where they lie; on the GPU box only the natives exist and callers
must drive ``_mapper.ReadMapper`` / ``_index_builder.ContigAssembler`` directly
(see ``RefMapResult`` below, a 20-line stand-in for the result collector the
native mapper calls back into: `_mapper.pyx:100-105`).
"""
import collections
import importlib
import pathlib
import sys
import threading
import warnings

import numpy

HERE = pathlib.Path(__file__).resolve().parent
REF_PY = pathlib.Path('/root/reference/seekmer')

_pkg = None


def available():
    from . import build_ref
    return build_ref.built()


def have_python_reference():
    return (REF_PY / 'mapper.py').exists() or (HERE / '_ref' / 'seekmer' / 'mapper.py').exists()


def load_ref():
    """Import and return the reference ``seekmer`` package (natives from _ref)."""
    global _pkg
    if _pkg is not None:
        return _pkg
    for p in (str(HERE / 'shims'), str(HERE / '_ref')):
        if p not in sys.path:
            sys.path.insert(0, p)
    if 'seekmer' in sys.modules:  # pragma: no cover
        raise RuntimeError('a different `seekmer` is already imported')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pkg = importlib.import_module('seekmer')
        if (REF_PY / 'mapper.py').exists():
            pkg.__path__.append(str(REF_PY))  # everything else of the reference, where it lies
        importlib.import_module('seekmer._common')
        importlib.import_module('seekmer._mapper')
        importlib.import_module('seekmer._index_builder')
    _pkg = pkg
    return pkg


class RefMapResult:
    """What `_mapper.ReadMapper` needs from its collector (`mapper.py:40-75,106-115`)."""

    def __init__(self):
        self.lock = threading.Lock()
        self.counter = collections.Counter()
        self.per_read = []          # list of tuples in read order (debug/parity)
        self.keep_per_read = False
        self.fragment_length_counts = numpy.zeros(2000, dtype='i8')

    def update(self, read_names, iterable):
        self.counter.update(iterable)
        if self.keep_per_read:
            self.per_read.extend(iterable)

    def merge_fragment_lengths(self, arr):
        self.fragment_length_counts += arr


def ref_index_from_arrays(kmers, contigs, sequences, targets, transcripts=None,
                          exons=None):
    pkg = load_ref()
    return pkg._common.KMerIndex(kmers, contigs, sequences, targets, transcripts,
                                 exons)


def ref_build_index(sequences):
    """Run the reference ContigAssembler on a list of transcript byte strings.

    Returns the four hot-path arrays (kmers, contigs, sequences, targets)
    (`_index_builder.pyx:105-150`).
    """
    pkg = load_ref()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        asm = pkg._index_builder.ContigAssembler()
        return asm.assemble(list(sequences))


def ref_map(index, batches, keep_per_read=False):
    """Map pre-built feeder batches with the native reference mapper, 1 thread."""
    pkg = load_ref()
    res = RefMapResult()
    res.keep_per_read = keep_per_read
    pkg._mapper.ReadMapper(index, res)(iter(batches))
    return res


def ref_map_threads(index, batches, job_count):
    """`mapper.map_reads` threading model (`mapper.py:174-189`) on the natives."""
    import queue
    pkg = load_ref()
    res = RefMapResult()
    q = queue.Queue(job_count * 2)
    threads = []
    for _ in range(job_count):
        t = threading.Thread(target=pkg._mapper.ReadMapper(index, res),
                             args=(iter(q.get, None),))
        threads.append(t)
        t.start()
    for b in batches:
        q.put(b)
    for _ in range(job_count):
        q.put(None)
    for t in threads:
        t.join()
    return res
