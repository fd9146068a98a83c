"""CPU unit test of the unified, register-window SIFT4 (seekmer_b200/csrc/sift4.cuh) against
the oracle's literal restatements of sift4_align_left / sift4_align_right."""
import ctypes
import subprocess

import numpy
import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    out = tmp_path_factory.mktemp('sift4') / 'libsift4_host.so'
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', str(ROOT / 'tests' / 'sift4_host.cpp'),
                    '-o', str(out)], check=True)
    lib = ctypes.CDLL(str(out))
    lib.sift4_window.restype = ctypes.c_int
    lib.sift4_window.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.reverse_pairs_c.restype = ctypes.c_uint
    lib.reverse_pairs_c.argtypes = [ctypes.c_uint, ctypes.c_int]
    lib.reverse_bits_c.restype = ctypes.c_uint
    lib.reverse_bits_c.argtypes = [ctypes.c_uint, ctypes.c_int]
    return lib


def mutate(rng, s, kind):
    a = bytearray(s)
    if kind == 0:
        return bytes(a)
    n_edits = int(rng.integers(1, 4))
    for _ in range(n_edits):
        p = int(rng.integers(0, len(a)))
        op = int(rng.integers(0, 5))
        if op == 0:
            a[p] = b'ACGT'[int(rng.integers(0, 4))]
        elif op == 1 and len(a) > 30:
            del a[p]
            a.append(b'ACGT'[int(rng.integers(0, 4))])
        elif op == 2:
            a.insert(p, b'ACGT'[int(rng.integers(0, 4))])
            a.pop()
        elif op == 3:
            a[p] = ord('N')
        else:
            a[p] = a[p] | 0x20  # lower case: wildcard for the check, same code for k-mers
    return bytes(a)


def test_unified_sift4_equals_reference_routines(host, orc):
    L = orc.lib()
    rng = numpy.random.Generator(numpy.random.PCG64(11))
    n_cases = 0
    seen = set()
    for trial in range(60000):
        alphabet = b'ACGT' if trial % 3 else b'AC'  # low-complexity cases hit the misaligned paths
        length = int(rng.integers(25, 60))
        read = bytes(rng.choice(list(alphabet), size=length).astype('u1'))
        left = int(trial & 1)
        if left:
            offset = int(rng.integers(0, length - 25 + 1))
        else:
            offset = int(rng.integers(17, length - 8 + 1))
        # contig windows are upper-case ACGT only: substitute, never mask
        ref = bytearray(read[offset:offset + 8])
        for _ in range(int(rng.integers(0, 3))):
            ref[int(rng.integers(0, 8))] = b'ACGT'[int(rng.integers(0, 4))]
        ref = bytes(ref)
        query = mutate(rng, read, int(rng.integers(0, 3)))[:length]
        if left:
            want = L.skmo_sift4_align_left(ref, 8, query, length, offset)
        else:
            want = L.skmo_sift4_align_right(ref, 8, query, length, offset)
        got = host.sift4_window(ref, query, length, offset, left)
        assert got == want, (trial, ref, query, offset, left, got, want)
        seen.add(want)
        n_cases += 1
    assert n_cases > 50000
    # the interesting outcomes all occurred: clean, shifted both ways, invalid
    assert 0 in seen and 0x7FFF in seen and any(0 < s < 9 for s in seen) and any(s < 0 for s in seen)


def test_bit_helpers(host):
    rng = numpy.random.Generator(numpy.random.PCG64(12))
    for _ in range(2000):
        n = int(rng.integers(1, 16))
        fields = rng.integers(0, 4, size=n)
        x = sum(int(f) << (2 * i) for i, f in enumerate(fields))
        y = sum(int(f) << (2 * (n - 1 - i)) for i, f in enumerate(fields))
        assert host.reverse_pairs_c(x, n) == y
        m = int(rng.integers(1, 32))
        bits = rng.integers(0, 2, size=m)
        x = sum(int(b) << i for i, b in enumerate(bits))
        y = sum(int(b) << (m - 1 - i) for i, b in enumerate(bits))
        assert host.reverse_bits_c(x, m) == y
