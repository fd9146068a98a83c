"""Sort-based construction of a Seekmer index (reference array layout) from transcripts.

Scope note: index construction (`seekmer index`, `_index_builder.pyx`) is OUT of the hot-path
scope.  This module exists because the benchmark
configuration needs a human-scale index (~300 Mb of cDNA) on the GPU box within seconds, and
the reference's sequential assembler would take tens of minutes there.  It is workload
preparation, not the measured path, and it is written with device-agnostic torch ops so the
same code runs on the CPU for tests and on the B200 at full size.

What it produces is the reference's *input contract* (SURVEY.md §8(a) I1-I5): a linear-probing
k-mer table placed with the reference's SipHash variant, the 48-byte contig table, the pooled
ASCII contig sequences and per-contig sorted target lists.  Contigs are the maximal
non-branching k-mer paths of the transcript-coloured de Bruijn graph: two consecutive k-mers
stay in one contig iff every occurrence of either is adjacent to the other, and no transcript
starts or ends between them — the same rule the reference applies incrementally
(`_index_builder.pyx:256-307`, unlink at transcript ends `:248-249,271-274`).  Contig numbering
and orientation differ from the reference's (they depend on its hash-table scan order); mapping
results do not depend on either, only on the partition (checked in tests/test_index_build.py).
"""
import numpy
import torch

from . import _lib

K = 25
_MASK50 = (1 << 50) - 1


def _kmers_at_all_positions(codes):
    """Forward and reverse-complement 25-mers starting at every base (int64 tensors of length
    N; entries within 24 bases of the end are garbage and are masked out by the caller)."""
    n = codes.shape[0]
    c = codes.to(torch.int64)
    pad = torch.zeros(K, dtype=torch.int64, device=codes.device)
    c = torch.cat([c, pad])
    fwd = torch.zeros(n, dtype=torch.int64, device=codes.device)
    rev = torch.zeros(n, dtype=torch.int64, device=codes.device)
    for j in range(K):
        b = c[j:j + n]
        fwd |= b << (2 * (K - 1 - j))
        rev |= (3 - b) << (2 * j)
    return fwd, rev


def _scatter_min_max(index, value, size):
    big = torch.iinfo(torch.int64).max
    mn = torch.full((size,), big, dtype=torch.int64, device=index.device)
    mx = torch.full((size,), -1, dtype=torch.int64, device=index.device)
    mn.scatter_reduce_(0, index, value, reduce='amin', include_self=True)
    mx.scatter_reduce_(0, index, value, reduce='amax', include_self=True)
    return mn, mx


def _sip_round(v0, v1, v2, v3):
    def rotl(x, s):
        return (x << s) | ((x >> (64 - s)) & ((1 << s) - 1))
    v0 = v0 + v1
    v2 = v2 + v3
    v1 = rotl(v1, 13) ^ v0
    v3 = rotl(v3, 16) ^ v2
    v0 = rotl(v0, 32)
    v2 = v2 + v1
    v0 = v0 + v3
    v1 = rotl(v1, 17) ^ v2
    v3 = rotl(v3, 21) ^ v0
    v2 = rotl(v2, 32)
    return v0, v1, v2, v3


def reference_hash(m):
    """The reference's SipHash-2-4 variant (`_kmer.pxd:174-231`) on an int64 tensor (two's
    complement wrap-around arithmetic == uint64)."""
    def c(x):
        x &= (1 << 64) - 1
        return x - (1 << 64) if x >= (1 << 63) else x
    v0 = torch.full_like(m, c(5381 ^ 0x736f6d6570736575))
    v1 = torch.full_like(m, c(42 ^ 0x646f72616e646f6d))
    v2 = torch.full_like(m, c(5381 ^ 0x6c7967656e657261))
    v3 = torch.full_like(m, c(42 ^ 0x7465646279746573)) ^ m
    for _ in range(2):
        v0, v1, v2, v3 = _sip_round(v0, v1, v2, v3)
    v0 = v0 ^ m
    v3 = v3 ^ c(8 << 56)
    for _ in range(2):
        v0, v1, v2, v3 = _sip_round(v0, v1, v2, v3)
    v2 = v2 ^ 0xff
    for _ in range(4):
        v0, v1, v2, v3 = _sip_round(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def _table_size(count):
    size = 1024  # `_index_builder.pyx:18`; doubled while count > 0.8 * size (`:194-197`)
    while count > 0.8 * size:
        size <<= 1
    return size


def _insert_table_torch(stored, canon, entry, offset, n_slots):
    """Vectorised linear-probing insert (CPU path for tests; the GPU uses the CUDA kernel)."""
    dev = stored.device
    keys = torch.full((n_slots,), -1, dtype=torch.int64, device=dev)
    ent = torch.full((n_slots,), -1, dtype=torch.int32, device=dev)
    off = torch.full((n_slots,), -1, dtype=torch.int32, device=dev)
    occupied = torch.zeros(n_slots, dtype=torch.bool, device=dev)
    slot = reference_hash(canon) & (n_slots - 1)
    pending = torch.arange(stored.shape[0], device=dev)
    while pending.numel():
        s = slot[pending]
        free = ~occupied[s]
        # among pending items that see a free slot, the smallest index per slot wins
        cand = pending[free]
        cs = s[free]
        winner = torch.full((n_slots,), stored.shape[0], dtype=torch.int64, device=dev)
        winner.scatter_reduce_(0, cs, cand, reduce='amin', include_self=True)
        won = winner[cs] == cand
        w, ws = cand[won], cs[won]
        keys[ws] = stored[w]
        ent[ws] = entry[w].to(torch.int32)
        off[ws] = offset[w].to(torch.int32)
        occupied[ws] = True
        done = torch.zeros(stored.shape[0], dtype=torch.bool, device=dev)
        done[w] = True
        pending = pending[~done[pending]]
        slot[pending] = (slot[pending] + 1) & (n_slots - 1)
    return keys, ent, off


def _insert_table_cuda(stored, entry, offset, n_slots):
    table = torch.empty(n_slots * 2, dtype=torch.int64, device=stored.device)
    L = _lib.load()
    _lib.check(L.skm_build_kmer_table(_lib._ptr(stored), _lib._ptr(entry), _lib._ptr(offset),
                                      stored.shape[0], _lib._ptr(table), n_slots,
                                      stored.device.index or 0,
                                      _lib.current_stream_ptr(stored.device)))
    return table


class BuiltIndex:
    """Reference-layout index arrays as torch tensors on `device` (+ numpy views on demand)."""

    def __init__(self, kmers, contigs, sequences, targets, n_transcripts, lengths, stats):
        self.kmers = kmers            # int64 [n_slots * 2]  (kmer, entry | offset << 32)
        self.contigs = contigs        # int64 [n_contigs * 6]
        self.sequences = sequences    # uint8 [n_bases] ASCII
        self.targets = targets        # int32 [n_targets * 2]
        self.n_transcripts = n_transcripts
        self.lengths = lengths
        self.stats = stats

    def numpy_arrays(self):
        k = self.kmers.cpu().numpy().view(_lib.SLOT_DTYPE)
        c = self.contigs.cpu().numpy().view(_lib.CONTIG_DTYPE)
        s = self.sequences.cpu().numpy().view('S1')
        t = self.targets.cpu().numpy().view(_lib.TARGET_DTYPE)
        return k, c, s, t

    def transcripts_table(self, ids=None, gene_ids=None):
        n = self.n_transcripts
        ids = ids if ids is not None else [b'TX%07d' % i for i in range(n)]
        gene_ids = gene_ids if gene_ids is not None else ids
        w = max(len(i) for i in ids)
        g = max(len(i) for i in gene_ids)
        tab = numpy.zeros(n, dtype=[('transcript_id', 'S%d' % w), ('gene_id', 'S%d' % g),
                                    ('length', 'f8')])
        tab['transcript_id'] = ids
        tab['gene_id'] = gene_ids
        tab['length'] = numpy.asarray(self.lengths, dtype='f8')
        return tab


def build_index(codes, offsets, device=None):
    """codes: uint8 tensor/array (values 0..3) of all transcripts concatenated;
    offsets: int64 [T+1].  Returns BuiltIndex with tensors on `device`."""
    if device is None:
        device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    device = torch.device(device)
    codes = torch.as_tensor(codes).to(device)
    offsets = torch.as_tensor(offsets).to(device=device, dtype=torch.int64)
    n_tx = offsets.shape[0] - 1
    n = codes.shape[0]
    lengths = (offsets[1:] - offsets[:-1])

    # ---- occurrences: every position that starts a k-mer inside one transcript
    tx_of_pos = torch.repeat_interleave(torch.arange(n_tx, device=device), lengths)
    pos = torch.arange(n, device=device)
    valid = pos + K <= offsets[tx_of_pos + 1]
    fwd_all, rev_all = _kmers_at_all_positions(codes)
    occ_pos = pos[valid]
    occ_tx = tx_of_pos[valid]
    fwd = fwd_all[valid]
    rev = rev_all[valid]
    del fwd_all, rev_all, tx_of_pos, pos, valid
    orient = (fwd < rev)                      # True: the transcript reads the canonical k-mer
    canon = torch.where(orient, fwd, rev)
    del rev
    uniq, node = torch.unique(canon, return_inverse=True)
    n_nodes = uniq.shape[0]
    m = occ_pos.shape[0]

    # ---- half-edge events.  side id = 2*node + s, s=1: canonical 3' side, s=0: 5' side.
    o = orient.to(torch.int64)
    out_side = 2 * node + o           # side through which the transcript leaves this k-mer
    in_side = 2 * node + (1 - o)      # side through which it enters
    same_tx = occ_tx[1:] == occ_tx[:-1]
    a = out_side[:-1][same_tx]
    b = in_side[1:][same_tx]
    mn, mx = _scatter_min_max(torch.cat([a, b]), torch.cat([b, a]), 2 * n_nodes)
    first = torch.ones(m, dtype=torch.bool, device=device)
    first[1:] = ~same_tx
    last = torch.ones(m, dtype=torch.bool, device=device)
    last[:-1] = ~same_tx
    term = torch.zeros(2 * n_nodes, dtype=torch.bool, device=device)
    term[in_side[first]] = True
    term[out_side[last]] = True
    link = torch.where((mn == mx) & ~term, mn, torch.full_like(mn, -1))
    del mn, mx, term, a, b
    sides = torch.arange(2 * n_nodes, device=device)
    has = link >= 0
    mutual = torch.zeros_like(has)
    mutual[has] = (link[link[has]] == sides[has]) & (link[has] != sides[has])
    link = torch.where(mutual, link, torch.full_like(link, -1))
    del has, mutual

    # ---- path decomposition by pointer jumping over directed states.
    # state 2*node+e = "entered node through side e", leaves through side 1-e.
    nxt = link[sides ^ 1]
    jump = nxt.clone()
    dist = (nxt >= 0).to(torch.int64)
    end = sides.clone()
    rounds = 0
    while True:
        act = jump >= 0
        if not bool(act.any()):
            break
        j = jump[act]
        end[act] = end[j]
        dist[act] = dist[act] + dist[j] * (jump[j] >= 0).to(torch.int64) + 0
        # dist counts steps: after hopping to j we still need dist[j] more steps
        jump[act] = jump[j]
        rounds += 1
        if rounds > 64:
            raise RuntimeError('cycle in the unitig graph')
    # fix-up: the loop above adds dist[j] only while j is non-terminal; recompute exactly
    del jump
    dist = _exact_distance(nxt)
    end_l, end_r = end[0::2], end[1::2]
    dist_l, dist_r = dist[0::2], dist[1::2]
    forward = end_l < end_r            # contig runs in the canonical direction of this node
    chosen_end = torch.where(forward, end_l, end_r)
    node_offset = torch.where(forward, dist_r, dist_l)
    contig_len_nodes = dist_l + dist_r + 1
    ends_sorted, contig_of_node = torch.unique(chosen_end, return_inverse=True)
    n_contigs = ends_sorted.shape[0]
    del end, dist, end_l, end_r, nxt, link

    # ---- contig table
    stored = torch.where(forward, uniq, _revcomp(uniq))       # contig-forward k-mers
    is_first = node_offset == 0
    is_last = node_offset == contig_len_nodes - 1
    c_len = torch.zeros(n_contigs, dtype=torch.int64, device=device)
    c_len[contig_of_node[is_first]] = contig_len_nodes[is_first] + (K - 1)
    c_first = torch.zeros(n_contigs, dtype=torch.int64, device=device)
    c_first[contig_of_node[is_first]] = stored[is_first]
    c_last = torch.zeros(n_contigs, dtype=torch.int64, device=device)
    c_last[contig_of_node[is_last]] = stored[is_last]
    c_off = torch.cumsum(c_len, 0) - c_len
    n_bases = int(c_len.sum())

    # ---- pooled sequences
    ascii_lut = torch.tensor(list(b'ACGT'), dtype=torch.uint8, device=device)
    seq = torch.zeros(n_bases, dtype=torch.uint8, device=device)
    base_pos = c_off[contig_of_node] + node_offset
    seq[base_pos] = ascii_lut[(stored >> (2 * (K - 1))) & 3]
    lp = base_pos[is_last]
    ls = stored[is_last]
    for j in range(1, K):
        seq[lp + j] = ascii_lut[(ls >> (2 * (K - 1 - j))) & 3]

    # ---- targets: occurrences of each contig's first k-mer (`_index_builder.pyx:502-518`)
    occ_first = is_first[node]
    t_contig = contig_of_node[node[occ_first]]
    same_dir = orient[occ_first] == forward[node[occ_first]]
    t_tx = occ_tx[occ_first]
    t_entry = torch.where(same_dir, t_tx, ~t_tx)
    t_off = occ_pos[occ_first] - offsets[t_tx]
    order = torch.argsort(t_off, stable=True)
    order = order[torch.argsort(t_entry[order], stable=True)]
    order = order[torch.argsort(t_contig[order], stable=True)]
    t_contig, t_entry, t_off = t_contig[order], t_entry[order], t_off[order]
    c_tcount = torch.bincount(t_contig, minlength=n_contigs)
    c_toff = torch.cumsum(c_tcount, 0) - c_tcount
    targets = torch.stack([t_entry.to(torch.int32), t_off.to(torch.int32)], dim=1).contiguous().view(-1)
    contigs = torch.stack([c_off, c_len, c_first, c_last, c_toff, c_tcount], dim=1).contiguous().view(-1)

    # ---- k-mer table in the reference layout
    n_slots = _table_size(n_nodes)
    ent32 = contig_of_node.to(torch.int32)
    off32 = node_offset.to(torch.int32)
    if device.type == 'cuda':
        table = _insert_table_cuda(stored.contiguous(), ent32.contiguous(), off32.contiguous(), n_slots)
    else:
        keys, ent, off = _insert_table_torch(stored, uniq, ent32, off32, n_slots)
        packed = (ent.to(torch.int64) & 0xFFFFFFFF) | (off.to(torch.int64) << 32)
        table = torch.stack([keys, packed], dim=1).contiguous().view(-1)
    stats = dict(n_kmers=int(n_nodes), n_slots=int(n_slots), n_contigs=int(n_contigs),
                 n_bases=n_bases, n_targets=int(t_entry.shape[0]), occurrences=int(m),
                 max_target_count=int(c_tcount.max()), jump_rounds=rounds)
    return BuiltIndex(table, contigs, seq, targets, n_tx, lengths.cpu().numpy(), stats)


def _exact_distance(nxt):
    """Steps from every state to the end of its path, by pointer doubling."""
    jump = nxt.clone()
    dist = (nxt >= 0).to(torch.int64)
    while True:
        act = jump >= 0
        if not bool(act.any()):
            return dist
        j = jump[act]
        dist[act] = dist[act] + dist[j]
        jump[act] = jump[j]


def _revcomp(kmer):
    """Reverse complement of 25-mers held in int64 tensors."""
    x = kmer
    x = ((x >> 2) & 0x3333333333333333) | ((x & 0x3333333333333333) << 2)
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0f) | ((x & 0x0f0f0f0f0f0f0f0f) << 4)
    x = ((x >> 8) & 0x00ff00ff00ff00ff) | ((x & 0x00ff00ff00ff00ff) << 8)
    x = ((x >> 16) & 0x0000ffff0000ffff) | ((x & 0x0000ffff0000ffff) << 16)
    x = ((x >> 32) & 0xffffffff) | (x << 32)
    x = (x >> (64 - 2 * K)) & _MASK50
    return ~x & _MASK50
