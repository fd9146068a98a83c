"""Stand-in for PyTables (absent here; no HDF5 library in the image).

TEST INFRASTRUCTURE ONLY. The reference imports `tables` at module scope
(`_common.pyx:8`, `infer.py:12`); nothing on the mapping/EM path calls it.
"""


class Filters:
    def __init__(self, *args, **kwargs):
        pass


def open_file(*args, **kwargs):
    raise RuntimeError('PyTables/HDF5 is not available in this environment')
