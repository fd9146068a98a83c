"""world_size-2 gloo tests (CPU) of the multi-GPU class-table merge, plus a world_size-1
no-op check.  The same `merge_class_tables` code runs on CUDA tensors under NCCL."""
import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, SYNTH_CASES


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _shard_tables(case, n_shards):
    """Per-shard and whole-sample class tables from the CPU oracle on the golden index."""
    from oracle import oracle as orc
    from seekmer_b200 import synth
    z = numpy.load(str(GOLDEN / 'synthetic_small.npz'))
    idx = orc.OracleIndex(z['kmers'], z['contigs'], z['sequences'], z['targets'])
    tx = synth.make_transcriptome(60, seed=7)
    kw = SYNTH_CASES[case]
    sim = synth.ReadSimulator(tx, synth.make_expression(tx.n_transcripts, seed=3), **kw)
    n = 3000

    def table(first, count):
        bases, _ = sim.generate(first, count)
        out = orc.map_batch(idx, bases, sim.offsets(count), kw['paired'])
        cls_ptr, cls_ids, cls_count, una = orc.tally(out.ptr, out.ids)
        # first-seen global unit of every class, in tally (first-seen) order
        seen, first_unit = set(), []
        for u, t in enumerate(out.tuples()):
            if t and t not in seen:
                seen.add(t)
                first_unit.append(first + u)
        return dict(key_offsets=cls_ptr, key_ids=cls_ids, counts=cls_count,
                    first_unit=numpy.asarray(first_unit, dtype='i8'), fld=out.fld, unaligned=una,
                    aligned=int(cls_count.sum()))
    bounds = numpy.linspace(0, n, n_shards + 1).astype(int)
    return [table(int(bounds[i]), int(bounds[i + 1] - bounds[i])) for i in range(n_shards)], table(0, n)


def _worker(rank, world, port, case, result_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from seekmer_b200 import dist as sdist
        shards, whole = _shard_tables(case, world)
        merged = sdist.table_to_host(sdist.merge_class_tables(sdist.table_from_host(shards[rank])))
        ok = all((merged[k] == whole[k]).all() for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld'))
        ok = ok and merged['unaligned'] == whole['unaligned'] and merged['aligned'] == whole['aligned']
        # a process group of more than one rank: the library is told to keep its EM scratch out of
        # the peers' address spaces (`_lib._note_process_group` -> skm_scratch_local_only)
        from seekmer_b200 import _lib
        ok = ok and _lib.load().skm_scratch_local_only(0) == 0   # default: cudaMalloc blocks
        _lib._note_process_group()
        ok = ok and _lib._PEERS['on'] and _lib.load().skm_scratch_local_only(1) == 1
        numpy.save(result_path % rank, numpy.asarray([int(ok), merged['counts'].shape[0]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case', ['pe100', 'se75'])
def test_two_rank_merge_equals_single_pass(tmp_path, case):
    port = _free_port()
    pattern = str(tmp_path / 'r%d.npy')
    mp.spawn(_worker, args=(2, port, case, pattern), nprocs=2, join=True)
    for r in range(2):
        ok, n = numpy.load(pattern % r).tolist()
        assert ok == 1 and n > 100


def test_reconcile_orders_by_first_seen_and_detects_duplicates():
    from seekmer_b200 import dist as sdist
    # rows: (5,), (1,2), (2,1), (5,), (1,2)  — order-sensitive, duplicates across "ranks"
    off = torch.tensor([0, 1, 3, 5, 6, 8])
    ids = torch.tensor([5, 1, 2, 2, 1, 5, 1, 2], dtype=torch.int32)
    first = torch.tensor([40, 7, 9, 3, 100])
    g_off, g_ids, g_first, gi = sdist.reconcile(off, ids, first, None)
    assert g_first.tolist() == [3, 7, 9]
    assert g_off.tolist() == [0, 1, 3, 5] and g_ids.tolist() == [5, 1, 2, 2, 1]
    assert gi.tolist() == [0, 1, 2, 0, 1]
    t = dict(key_offsets=off, key_ids=ids, counts=torch.ones(5, dtype=torch.int64), first_unit=first,
             fld=torch.zeros(2000, dtype=torch.int64), scalars=torch.zeros(2, dtype=torch.int64))
    assert sdist.merge_class_tables(t) is t  # not initialised => single process no-op
    from seekmer_b200 import _lib
    _lib._note_process_group()                # ... and no peers: the scratch policy stays at its default
    assert not _lib._PEERS['on'] and _lib.load().skm_scratch_local_only(0) == 0
