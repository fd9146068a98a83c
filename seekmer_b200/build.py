"""Compile the CUDA library in-tree: csrc/*.cu -> seekmer_b200/libseekmer_b200.so (sm_100a)."""
import os
import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
LIB = HERE / 'libseekmer_b200.so'
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def sources():
    return sorted(CSRC.glob('*.cu'))


def stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + [HERE.parent / 'include' / 'seekmer_b200.h']
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), output=None):
    if output is None and not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    (HERE / '_obj').mkdir(exist_ok=True)
    for src in sources():
        obj = HERE / '_obj' / (src.stem + ('' if output is None else '.' + pathlib.Path(output).stem) + '.o')
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ['-Xptxas', '-v', '-c', str(src), '-o', str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed for %s' % src.name)
        if verbose:
            sys.stderr.write(out)
    target = LIB if output is None else pathlib.Path(output)
    tmp = target.with_suffix('.tmp%d.so' % os.getpid())
    subprocess.run([nvcc, '-shared', '-o', str(tmp)] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'],
                   check=True)
    os.replace(tmp, target)
    return target


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
