#!/usr/bin/env python3
"""Compile the UNMODIFIED reference Cython modules into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline); never imported by the
product package.  Sources are read where they lie under ``/root/reference``;
only generated C and the built ``.so`` files land in ``oracle/_ref/`` (which is
git-ignored, so no reference source ever enters history).

Recipe follows SURVEY.md §8(c) / Appendix A:
* ``legacy_implicit_noexcept=True`` restores the Cython-0.28 ``nogil`` semantics
  the reference was written for (results identical, 6x faster single-threaded).
* ``-O3``; numpy include dir.

Only the three ``.pyx`` native modules are built.  The reference's pure-Python
modules (``mapper.py``, ``infer.py`` ...) cannot travel to the GPU box; they
are imported straight from ``/root/reference`` in this container when golden
vectors are (re)generated (``tests/golden/make_golden.py``).
"""
import os
import pathlib
import subprocess
import sys
import sysconfig

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path(os.environ.get('SEEKMER_REFERENCE', '/root/reference'))
OUT = HERE / '_ref'
MODULES = ('_common', '_mapper', '_index_builder')


def ref_available():
    return (REF / 'seekmer' / '_mapper.pyx').exists()


def built():
    suffix = sysconfig.get_config_var('EXT_SUFFIX')
    return all((OUT / 'seekmer' / (m + suffix)).exists() for m in MODULES)


def build(force=False, verbose=False):
    if not ref_available():
        return built()
    if built() and not force:
        return True
    import numpy
    suffix = sysconfig.get_config_var('EXT_SUFFIX')
    (OUT / 'build').mkdir(parents=True, exist_ok=True)
    (OUT / 'seekmer').mkdir(parents=True, exist_ok=True)
    # package marker written by us (the reference's own is not copied)
    (OUT / 'seekmer' / '__init__.py').write_text(
        '"""Compiled reference natives only (built by oracle/build_ref.py)."""\n')
    py_inc = sysconfig.get_paths()['include']
    for mod in MODULES:
        c_file = OUT / 'build' / (mod + '.c')
        cmd = [sys.executable, '-m', 'cython', '-3',
               '-X', 'legacy_implicit_noexcept=True',
               '-I', str(REF), str(REF / 'seekmer' / (mod + '.pyx')),
               '-o', str(c_file)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('cython failed for ' + mod)
        so = OUT / 'seekmer' / (mod + suffix)
        cmd = ['gcc', '-O3', '-fPIC', '-shared', '-w',
               '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION',
               '-I', py_inc, '-I', numpy.get_include(),
               str(c_file), '-o', str(so)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('gcc failed for ' + mod)
        if verbose:
            print('built', so)
    return True


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv, verbose=True)
    print('oracle/_ref', 'ready' if ok else 'unavailable (no reference sources, no prebuilt files)')
