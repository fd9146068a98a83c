#!/usr/bin/env python3
"""Build compile-time variants of the CUDA library for a parameter sweep on the GPU box.

    python tools/build_variants.py name1:-DSKM_X=1,-DSKM_Y=2 name2:...

Each variant becomes seekmer_b200/variants/lib_<name>.so (git-ignored, travels with gpurun).
ptxas register / spill figures of map_reads_kernel<24> are printed per variant."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seekmer_b200 import build  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, 'seekmer_b200', 'variants')
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(':')
        flags = [f for f in flags.split(',') if f]
        target = os.path.join(out_dir, 'lib_%s.so' % name)
        nvcc = os.environ.get('NVCC', 'nvcc')
        # ptxas figures of the map kernel instantiations
        cmd = [nvcc] + build.NVCC_FLAGS + flags + ['-Xptxas', '-v', '-c', str(build.CSRC / 'mapper.cu'), '-o',
                                                  '/tmp/_variant_probe.o']
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        regs = {}
        cur = None
        for line in out.splitlines():
            m = re.search(r"Compiling entry function '_ZN3skm16map_reads_kernelILi(\d+)E", line)
            if m:
                cur = int(m.group(1))
            elif 'Compiling entry function' in line:
                cur = None
            if cur is not None:
                m = re.search(r'(\d+) bytes spill stores, (\d+) bytes spill loads', line)
                if m:
                    regs.setdefault(cur, {})['spill'] = (int(m.group(1)), int(m.group(2)))
                m = re.search(r'Used (\d+) registers', line)
                if m:
                    regs.setdefault(cur, {})['regs'] = int(m.group(1))
        if 'error' in out:
            print(out)
            raise SystemExit(1)
        build.build(force=True, extra_flags=flags, output=target)
        print(name, flags, {k: regs[k] for k in sorted(regs) if k in (20, 24, 28, 32)}, flush=True)


if __name__ == '__main__':
    main()
