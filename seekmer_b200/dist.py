"""Multi-GPU merge of per-rank class tables.

Reads shard across ranks with the index replicated; every rank ends a mapping pass with a
local class dictionary (ordered id tuple -> count, first-seen global unit), an FLD histogram
and an unaligned counter.  The exchange step is:

  1. all-gather of the local key pools (+ first-seen unit per key) — tens of MB;
  2. every rank builds the identical global key set, ordered by the smallest first-seen
     global unit index (the reference's Counter insertion order at job_count=1), and remaps
     its local counts into a dense int64[C_global] vector;
  3. ONE all-reduce(sum, int64) over [counts | FLD[2000] | unaligned, aligned].

`torch.distributed` is plumbing only (NCCL over NVLink on the GPUs, gloo in the CPU tests).
Everything between the collectives is device-agnostic torch tensor code, so the very same
function runs on CUDA tensors under NCCL and on CPU tensors under gloo.
"""
import torch
import torch.distributed as dist

MAX_FRAGMENT_LENGTH = 2000
_I64_MAX = torch.iinfo(torch.int64).max


def _lsr(x, s):
    """Logical shift right of int64 tensors."""
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix(x, c1, c2):
    x = x ^ _lsr(x, 33)
    x = x * c1
    x = x ^ _lsr(x, 29)
    x = x * c2
    x = x ^ _lsr(x, 32)
    return x


def _c(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def row_signatures(key_offsets, key_ids):
    """Two independent order-sensitive 64-bit hashes of every CSR row (wrap-around int64)."""
    lens = key_offsets[1:] - key_offsets[:-1]
    n_rows = lens.shape[0]
    dev = key_ids.device
    if key_ids.numel() == 0:
        z = torch.zeros(n_rows, dtype=torch.int64, device=dev)
        return z, z.clone(), lens
    row = torch.repeat_interleave(torch.arange(n_rows, device=dev), lens)
    pos = torch.arange(key_ids.shape[0], device=dev) - key_offsets[:-1][row]
    v = key_ids.to(torch.int64)
    a = _mix(v * _c(0x9E3779B97F4A7C15) + pos * _c(0xD6E8FEB86659FD93) + 1,
             _c(0xff51afd7ed558ccd), _c(0xc4ceb9fe1a85ec53))
    b = _mix(v * _c(0xC2B2AE3D27D4EB4F) + pos * _c(0x165667B19E3779F9) + 7,
             _c(0xBF58476D1CE4E5B9), _c(0x94D049BB133111EB))
    h1 = torch.zeros(n_rows, dtype=torch.int64, device=dev).index_add_(0, row, a)
    h2 = torch.zeros(n_rows, dtype=torch.int64, device=dev).index_add_(0, row, b)
    return h1, h2 ^ lens, lens


def reconcile(key_offsets, key_ids, first_unit, owner):
    """Union of CSR rows by content.  Returns (g_offsets, g_ids, g_first, global_index_of_row)
    with global classes ordered by their smallest first-seen unit."""
    dev = key_ids.device
    h1, h2, lens = row_signatures(key_offsets, key_ids)
    uniq, inverse = torch.unique(h1, return_inverse=True)
    n_global = uniq.shape[0]
    # collision check on the second signature: all rows of a group must agree
    h2_min = torch.full((n_global,), _I64_MAX, dtype=torch.int64, device=dev).scatter_reduce_(
        0, inverse, h2, reduce='amin', include_self=True)
    if not bool((h2_min[inverse] == h2).all()):
        raise RuntimeError('64-bit signature collision between distinct classes; refusing to merge')
    g_first = torch.full((n_global,), _I64_MAX, dtype=torch.int64, device=dev).scatter_reduce_(
        0, inverse, first_unit, reduce='amin', include_self=True)
    order = torch.argsort(g_first, stable=True)
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(n_global, device=dev)
    global_idx = rank_of[inverse]
    # representative row of each global class = the row that saw it first
    is_rep = first_unit == g_first[inverse]
    rep = torch.zeros(n_global, dtype=torch.int64, device=dev)
    rep[global_idx[is_rep]] = torch.nonzero(is_rep).squeeze(1)
    rep_len = lens[rep]
    g_off = torch.zeros(n_global + 1, dtype=torch.int64, device=dev)
    g_off[1:] = torch.cumsum(rep_len, 0)
    n_ids = int(g_off[-1])
    src = torch.repeat_interleave(key_offsets[:-1][rep] - g_off[:-1], rep_len) + torch.arange(n_ids, device=dev)
    return g_off, key_ids[src], g_first[order], global_idx


def _all_gather_var(t, group):
    """all_gather of 1-D tensors of different lengths (same dtype, same device)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=t.dtype, device=t.device)
    buf[:t.shape[0]] = t
    outs = [torch.zeros(cap, dtype=t.dtype, device=t.device) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def merge_class_tables(table, group=None):
    """Collective.  `table`: dict of torch tensors on one device — key_offsets int64[C+1],
    key_ids int32, counts int64[C], first_unit int64[C], fld int64[2000], scalars
    int64[2] = (unaligned, aligned).  Returns the identical global table on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return table
    rank = dist.get_rank(group)
    lens = table['key_offsets'][1:] - table['key_offsets'][:-1]
    # (1) key pools: lengths, ids and first-seen units packed into two gathers
    meta = torch.stack([lens, table['first_unit']], dim=1).reshape(-1)
    g_meta = _all_gather_var(meta, group)
    g_ids = _all_gather_var(table['key_ids'], group)
    all_lens = torch.cat([m.reshape(-1, 2)[:, 0] for m in g_meta])
    all_first = torch.cat([m.reshape(-1, 2)[:, 1] for m in g_meta])
    all_ids = torch.cat(g_ids)
    rows_per_rank = [m.shape[0] // 2 for m in g_meta]
    all_off = torch.zeros(all_lens.shape[0] + 1, dtype=torch.int64, device=all_lens.device)
    all_off[1:] = torch.cumsum(all_lens, 0)
    # (2) identical global dictionary everywhere + dense local counts
    g_off, g_key_ids, g_first, global_idx = reconcile(all_off, all_ids, all_first, None)
    n_global = g_off.shape[0] - 1
    start = sum(rows_per_rank[:rank])
    mine = global_idx[start:start + rows_per_rank[rank]]
    dense = torch.zeros(n_global + MAX_FRAGMENT_LENGTH + 2, dtype=torch.int64, device=all_lens.device)
    dense.index_add_(0, mine, table['counts'])
    dense[n_global:n_global + MAX_FRAGMENT_LENGTH] = table['fld']
    dense[n_global + MAX_FRAGMENT_LENGTH:] = table['scalars']
    # (3) one all-reduce
    dist.all_reduce(dense, op=dist.ReduceOp.SUM, group=group)
    return dict(key_offsets=g_off, key_ids=g_key_ids, counts=dense[:n_global].clone(), first_unit=g_first,
                fld=dense[n_global:n_global + MAX_FRAGMENT_LENGTH].clone(),
                scalars=dense[n_global + MAX_FRAGMENT_LENGTH:].clone())


_CAP_HINT = {}  # words per rank of the packed exchange buffer, per process group


def packed_words(n_classes, n_ids):
    """int64 words of one rank's packed export (`DeviceMapper.pack_raw_torch`)."""
    return 3 + MAX_FRAGMENT_LENGTH + (n_classes + 1) + 2 * n_classes + (n_ids + 1) // 2


def merge_mappers(mp, group=None, stages=None):
    """Collective, CUDA only: make every rank's device dictionary the global one and return the
    global table (same shape and order as `merge_class_tables`).

    One exchange step: every rank ships its exported dictionary (CSR keys, counts, first-seen
    units, FLD, unaligned count; ~30 MB) in ONE all-gather over NVLink, then inserts the other
    ranks' classes into its own device dictionary with the library's merge kernel
    (`skm_classes_merge_packed`: same 128-bit tuple hash, ids compared on a key hit, counts added
    with atomics, first-seen unit by atomicMin).  All ranks end with the same dictionary,
    exported in first-seen order.  The buffer size per rank is remembered from the previous
    exchange (with headroom); the sizes the ranks actually sent are checked after the gather
    and only an overflow costs a second, larger exchange.

    `stages`, when a dict, receives the device time (ms, CUDA events) of the four stages."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mp.export_torch()
    from . import _lib
    _lib.uses_peer_gpus()  # NCCL has peer access on: EM scratch from blocks only this device maps
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if stages is not None else None

    def mark(k):
        if ev is not None:
            ev[k].record()
    mark(0)
    packed = mp.pack_raw_torch()
    dev = packed.device
    own = int(packed.shape[0])
    mark(1)
    key = id(group) if group is not None else 0
    cap = _CAP_HINT.get(key)
    while True:
        if cap is None:  # first exchange: learn the sizes
            sizes = torch.zeros(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(sizes, torch.tensor([own], dtype=torch.int64, device=dev), group=group)
            cap = int(sizes.max().item())
            cap += cap // 4 + 1024
        mine = torch.zeros(cap, dtype=torch.int64, device=dev)
        mine[:min(own, cap)] = packed[:cap]  # a buffer that does not fit still carries its header
        gathered = torch.zeros(world * cap, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=group)
        heads = gathered.view(world, cap)[:, :2].cpu().tolist()  # [n_classes, n_ids] of every rank
        need = max(packed_words(int(n), int(k)) for n, k in heads)
        if need <= cap:
            break
        cap = need + need // 4 + 1024  # every rank sees the same headers: all of them go round again
    _CAP_HINT[key] = cap
    mark(2)
    mp.merge_packed(gathered, cap, world, rank)
    mark(3)
    table = mp.export_torch()
    mark(4)
    if ev is not None:
        torch.cuda.synchronize()
        for k, name in enumerate(('export_pack_ms', 'all_gather_ms', 'merge_kernel_ms', 'export_ms')):
            stages[name] = ev[k].elapsed_time(ev[k + 1])
    return table


def table_to_host(table):
    """Torch table -> the numpy dict shape of `_lib.DeviceMapper.export()`."""
    sc = table['scalars'].cpu().tolist()
    return dict(key_offsets=table['key_offsets'].cpu().numpy(), key_ids=table['key_ids'].cpu().numpy(),
                counts=table['counts'].cpu().numpy(), first_unit=table['first_unit'].cpu().numpy(),
                fld=table['fld'].cpu().numpy(), unaligned=int(sc[0]), aligned=int(sc[1]))


def table_from_host(table, device='cpu'):
    import numpy
    t = lambda a, dt: torch.from_numpy(numpy.ascontiguousarray(a, dtype=dt)).to(device)  # noqa: E731
    return dict(key_offsets=t(table['key_offsets'], 'i8'), key_ids=t(table['key_ids'], 'i4'),
                counts=t(table['counts'], 'i8'), first_unit=t(table['first_unit'], 'i8'),
                fld=t(table['fld'], 'i8'),
                scalars=torch.tensor([table['unaligned'], table['aligned']], dtype=torch.int64, device=device))
