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

The three ``.pyx`` native modules are compiled.  The reference's pure-Python modules of the
path (``mapper.py``, ``infer.py``, ``common.py``) are PLACED next to them, unmodified, so that the
CPU baseline can call the stock ``mapper.map_reads`` / ``infer.quantify`` on the GPU box, where
``/root/reference`` does not exist.  ``oracle/_ref/`` is git-ignored: no reference source enters
history; the directory only travels with the gpurun snapshot like any other build product.
"""
import os
import pathlib
import shutil
import subprocess
import sys
import sysconfig

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path(os.environ.get('SEEKMER_REFERENCE', '/root/reference'))
OUT = HERE / '_ref'
MODULES = ('_common', '_mapper', '_index_builder')
PURE_PYTHON = ('mapper.py', 'infer.py', 'common.py')


def ref_available():
    return (REF / 'seekmer' / '_mapper.pyx').exists()


def built():
    suffix = sysconfig.get_config_var('EXT_SUFFIX')
    return all((OUT / 'seekmer' / (m + suffix)).exists() for m in MODULES)


def python_modules_placed():
    return all((OUT / 'seekmer' / name).exists() for name in PURE_PYTHON)


def place_python_modules():
    """Unmodified copies of the reference's pure-Python modules into the git-ignored build dir."""
    if not ref_available():
        return python_modules_placed()
    (OUT / 'seekmer').mkdir(parents=True, exist_ok=True)
    for name in PURE_PYTHON:
        shutil.copyfile(str(REF / 'seekmer' / name), str(OUT / 'seekmer' / name))
    return True


def build(force=False, verbose=False):
    if not ref_available():
        return built()
    place_python_modules()
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
