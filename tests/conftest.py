import pathlib
import sys

import numpy
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _has_gpu():
    try:
        from seekmer_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


class Golden:
    def __init__(self, path):
        z = numpy.load(str(path), allow_pickle=False)
        self.z = {k: z[k] for k in z.files}

    def __getitem__(self, k):
        return self.z[k]

    def index_arrays(self):
        return self.z['kmers'], self.z['contigs'], self.z['sequences'], self.z['targets']

    def tuples(self, prefix):
        ptr = self.z[prefix + 'unit_ptr'].tolist()
        ids = self.z[prefix + 'unit_ids'].tolist()
        return [tuple(ids[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


@pytest.fixture(scope='session')
def golden_chr21():
    return Golden(GOLDEN / 'chr21_subset.npz')


@pytest.fixture(scope='session')
def golden_synth():
    return Golden(GOLDEN / 'synthetic_small.npz')


@pytest.fixture(scope='session')
def small_tx():
    from seekmer_b200 import synth
    return synth.make_transcriptome(60, seed=7)


SYNTH_CASES = {
    'pe100': dict(read_length=100, frag_mean=250, frag_sd=30, sub_rate=0.01, paired=True, seed=10),
    'pe150': dict(read_length=150, frag_mean=350, frag_sd=50, sub_rate=0.01, paired=True, seed=11),
    'se75': dict(read_length=75, frag_mean=250, frag_sd=30, sub_rate=0.02, paired=False, seed=12),
}
N_GOLDEN_UNITS = 3000


def ref_available():
    from oracle import build_ref
    return build_ref.built()


@pytest.fixture(scope='session')
def ref():
    """The compiled reference natives, when oracle/_ref has been built."""
    if not ref_available():
        pytest.skip('oracle/_ref not built (needs /root/reference; run oracle/build_ref.py)')
    from oracle import ref_harness
    ref_harness.load_ref()
    return ref_harness


@pytest.fixture(scope='session')
def medium(ref):
    """2 000-transcript synthetic transcriptome (config C1 scale) indexed by the reference."""
    from seekmer_b200 import synth
    tx = synth.make_transcriptome(2000, seed=1)
    arrays = ref.ref_build_index(tx.sequences())
    return tx, arrays
