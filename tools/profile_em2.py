#!/usr/bin/env python3
"""Where does EM time go?  host prep vs skm_em with host buffers vs skm_em with device buffers."""
import os, sys, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seekmer_b200 import infer, _lib
from tools.profile_em import structure

def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    class_map, counts, eff = structure()
    T, C = eff.size, counts.size
    x0 = numpy.ones(T) / eff; x0 /= x0.sum()
    t = time.perf_counter(); ptr, tx = infer._csr_from_class_map(class_map, C); print('csr prep %.1f ms' % ((time.perf_counter() - t) * 1e3))
    L = _lib.load()
    dev = torch.device('cuda', 0)
    d_ptr = torch.from_numpy(ptr).to(dev); d_tx = torch.from_numpy(tx).to(dev); d_eff = torch.from_numpy(eff).to(dev)
    for reps in (1, R):
        cnt = numpy.tile(counts, (reps, 1)).astype('f8')
        if reps > 1:
            cnt = infer._resample(counts, reps, 1234).astype('f8')
        xs = numpy.tile(x0, (reps, 1))
        out = numpy.zeros_like(xs); iters = numpy.zeros(reps, dtype='i4')
        for rep in range(2):
            t = time.perf_counter()
            _lib.check(L.skm_em(_lib._np_ptr(ptr), _lib._np_ptr(tx), C, tx.shape[0], _lib._np_ptr(cnt), _lib._np_ptr(eff), T,
                                _lib._np_ptr(xs), reps, 0, _lib._np_ptr(out), _lib._np_ptr(iters), 0, 0, None))
            print('R=%d host buffers: %.1f ms (iters max %d mean %.1f)' % (reps, (time.perf_counter() - t) * 1e3, iters.max(), iters.mean()))
        d_cnt = torch.from_numpy(cnt).to(dev); d_x = torch.from_numpy(xs).to(dev)
        d_out = torch.zeros_like(d_x); d_it = torch.zeros(reps, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for rep in range(2):
            t = time.perf_counter()
            _lib.check(L.skm_em(d_ptr.data_ptr(), d_tx.data_ptr(), C, tx.shape[0], d_cnt.data_ptr(), d_eff.data_ptr(), T,
                                d_x.data_ptr(), reps, 0, d_out.data_ptr(), d_it.data_ptr(), 1, 0, None))
            torch.cuda.synchronize()
            print('R=%d device buffers: %.1f ms' % (reps, (time.perf_counter() - t) * 1e3))

if __name__ == '__main__':
    main()
