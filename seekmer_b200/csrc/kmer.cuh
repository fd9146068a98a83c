// Device primitives: 2-bit k-mers, reverse complement, table hash, probes.
// Semantics follow the reference's _kmer.pxd / _coordinate.pxd / _common.pyx
// (cited per function); the implementation is specific to the GPU layout in
// common.cuh.
#pragma once

#include "common.cuh"

namespace skm {

struct Coord {
    int32_t entry;
    int32_t offset;
};

__device__ __forceinline__ Coord coord_invalid() { return Coord{0, -1}; }  // _coordinate.pxd:13-24

// Read-only global loads that carry an L2 eviction policy (ld.global.nc.L2::cache_hint): the
// index parts a contig walk revisits are kept (evict_last), what streams through once is
// marked evict_first, so that 4 GB of table buckets do not push 118 MB of contig data out.
__device__ __forceinline__ ulonglong2 ld_hint_16(const void *p, uint64_t pol)
{
    ulonglong2 v;
    asm("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol));
    return v;
}
// 32 bytes per lane in one instruction (LDG.E.256, sm_100), streaming: packed reads are read once
struct Quad64 {
    uint64_t a, b, c, d;
};
__device__ __forceinline__ Quad64 ld_cs_32(const void *p)
{
    Quad64 v;
    asm("ld.global.cs.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_hint_8(const void *p, uint64_t pol)
{
    int2 v;
    asm("ld.global.nc.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_hint_4(const void *p, uint64_t pol)
{
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// Reverse complement of a 25-mer held in the low 50 bits (_kmer.pxd:146-171).
// brev reverses all 64 bits; swapping the two bits of every pair restores base
// codes; the k-mer then sits in the top 50 bits.
__host__ __device__ __forceinline__ uint64_t revcomp(uint64_t kmer)
{
#ifdef __CUDA_ARCH__
    uint64_t r = __brevll(kmer);
#else
    uint64_t r = kmer;
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    r = ((r >> 2) & 0x3333333333333333ULL) | ((r & 0x3333333333333333ULL) << 2);
    r = ((r >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((r & 0x0f0f0f0f0f0f0f0fULL) << 4);
    r = ((r >> 8) & 0x00ff00ff00ff00ffULL) | ((r & 0x00ff00ff00ff00ffULL) << 8);
    r = ((r >> 16) & 0x0000ffff0000ffffULL) | ((r & 0x0000ffff0000ffffULL) << 16);
    r = (r >> 32) | (r << 32);
#endif
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    return ~(r >> (64 - 2 * K)) & KMER_MASK;
}

// Reverse complement of 8 bases held in the low 16 bits (_sequence.pxd:53-75 on a
// 2-bit window).
__device__ __forceinline__ uint32_t revcomp8(uint32_t w)
{
    uint32_t r = __brev(w) >> 16;
    r = ((r >> 1) & 0x5555u) | ((r & 0x5555u) << 1);
    return ~r & 0xFFFFu;
}

// Home-slot hash of the device table.  The reference's SipHash variant
// (_kmer.pxd:174-219) only has to place keys; exact-membership lookup gives the same
// answer under any hash (k odd => a k-mer never equals its reverse complement, so at
// most one slot matches), so the re-laid-out table uses a 2-multiply finaliser.
__host__ __device__ __forceinline__ uint64_t table_hash(uint64_t canon)
{
    uint64_t x = canon;
    x ^= x >> 31;
    x *= 0x9E3779B97F4A7C15ULL;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 32;
    return x;
}

// The reference's own hash, needed only to build reference-layout tables on the
// device (index construction support) — _kmer.pxd:174-231.
__host__ __device__ __forceinline__ void sip_half(uint64_t &a, uint64_t &b, uint64_t &c,
                                                   uint64_t &d, int s, int t)
{
    a += b;
    c += d;
    b = ((b << s) | (b >> (64 - s))) ^ a;
    d = ((d << t) | (d >> (64 - t))) ^ c;
    a = (a << 32) | (a >> 32);
}

__host__ __device__ __forceinline__ uint64_t reference_hash(uint64_t m)
{
    uint64_t v0 = 5381ULL ^ 0x736f6d6570736575ULL;
    uint64_t v1 = 42ULL ^ 0x646f72616e646f6dULL;
    uint64_t v2 = 5381ULL ^ 0x6c7967656e657261ULL;
    uint64_t v3 = 42ULL ^ 0x7465646279746573ULL;
    v3 ^= m;
    for (int r = 0; r < 2; ++r) {
        sip_half(v0, v1, v2, v3, 13, 16);
        sip_half(v2, v1, v0, v3, 17, 21);
    }
    v0 ^= m;
    v3 ^= 8ULL << 56;
    for (int r = 0; r < 2; ++r) {
        sip_half(v0, v1, v2, v3, 13, 16);
        sip_half(v2, v1, v0, v3, 17, 21);
    }
    v2 ^= 0xff;  // and no `v0 ^= b` — the reference's deviation (_kmer.pxd:209)
    for (int r = 0; r < 4; ++r) {
        sip_half(v0, v1, v2, v3, 13, 16);
        sip_half(v2, v1, v0, v3, 17, 21);
    }
    return (v0 ^ v1) ^ (v2 ^ v3);
}

// Home bucket of a k-mer in the device table.
__host__ __device__ __forceinline__ uint32_t home_bucket_of(uint64_t canon, uint64_t bucket_mask)
{
    return (uint32_t)((table_hash(canon) >> 8) & bucket_mask);
}

// KMerIndex.map_kmer (_common.pyx:54-97) on the canonical-key table: hit on the
// canonical key; strand of the query relative to the canonical form decides whether
// the stored coordinate is returned as is or reverse-complemented (~entry).
// One iteration reads a whole 64-byte bucket with four 16-byte loads; at load
// <= 0.25 a second bucket is needed ~0.4 % of the time.
__device__ __forceinline__ Coord probe_canonical(const Slot *table, uint64_t bucket_mask, uint64_t canon, bool fwd,
                                                 uint32_t bucket)
{
    uint64_t b = bucket;
    for (;;) {
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(table + BUCKET_SLOTS * b);
        // streaming loads: a bucket is used once, it should not push the read's own lines out of L1
        // (nor the contig data out of L2)
        const ulonglong2 s0 = __ldcs(p), s1 = __ldcs(p + 1), s2 = __ldcs(p + 2), s3 = __ldcs(p + 3);
        uint64_t v = 0;
        bool found = false;
        if (s0.x == canon) { v = s0.y; found = true; }
        if (s1.x == canon) { v = s1.y; found = true; }
        if (s2.x == canon) { v = s2.y; found = true; }
        if (s3.x == canon) { v = s3.y; found = true; }
        if (found) {
            const int32_t entry = (int32_t)(uint32_t)v;
            return Coord{fwd ? entry : ~entry, (int32_t)(uint32_t)(v >> 32)};
        }
        if (s3.x == EMPTY_KEY) return coord_invalid();
        b = (b + 1) & bucket_mask;
    }
}

__device__ __forceinline__ Coord map_kmer(const DevIndex &ix, uint64_t kmer)
{
    const uint64_t rc = revcomp(kmer);
    const bool fwd = kmer < rc;
    const uint64_t canon = fwd ? kmer : rc;
    return probe_canonical(ix.table, ix.bucket_mask, canon, fwd, home_bucket_of(canon, ix.bucket_mask));
}

struct Contig {
    uint64_t first_kmer, last_kmer;
    int64_t seq_offset;
    uint32_t target_offset;
    int32_t target_count;
    int32_t length;
    int32_t t[INLINE_TARGETS];  // first 8 target entries
};

__device__ __forceinline__ Contig load_contig(const DevIndex &ix, int32_t index)
{
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(ix.contigs + index);
    const ulonglong2 a = ld_hint_16(p, ix.pol_hot);
    const ulonglong2 b = ld_hint_16(p + 1, ix.pol_hot);
    const ulonglong2 t0 = ld_hint_16(p + 2, ix.pol_hot);
    const ulonglong2 t1 = ld_hint_16(p + 3, ix.pol_hot);
    Contig c;
    c.t[0] = (int32_t)(uint32_t)t0.x; c.t[1] = (int32_t)(uint32_t)(t0.x >> 32);
    c.t[2] = (int32_t)(uint32_t)t0.y; c.t[3] = (int32_t)(uint32_t)(t0.y >> 32);
    c.t[4] = (int32_t)(uint32_t)t1.x; c.t[5] = (int32_t)(uint32_t)(t1.x >> 32);
    c.t[6] = (int32_t)(uint32_t)t1.y; c.t[7] = (int32_t)(uint32_t)(t1.y >> 32);
    c.first_kmer = a.x & KMER_MASK;
    c.last_kmer = a.y & KMER_MASK;
    c.target_count = (int32_t)((a.x >> 50) | ((a.y >> 50) << 14));
    c.seq_offset = (int64_t)b.x;
    c.target_offset = (uint32_t)b.y;
    c.length = (int32_t)(uint32_t)(b.y >> 32);
    return c;
}

// The coordinate of the k-mer that continues a contig walk across the junction at the anchor's
// contig edge with read base `b` (what _filter_targets_to_left/right look up at
// _mapper.pyx:247-249,309-311), from the contig's links.  Walking right (dir 1) along a forward
// anchor, or left along a reverse one, leaves through the contig's last k-mer; the other two
// cases through its first.  On the reverse strand the query is the reverse complement of a
// stored one: complement the base, bit-negate the entry.
__device__ __forceinline__ Coord contig_link(const DevIndex &ix, Coord anchor, int dir, uint32_t b)
{
    const bool forward = anchor.entry >= 0;
    const ContigRec *rec = ix.contigs + (forward ? anchor.entry : ~anchor.entry);
    const bool via_last = (dir != 0) == forward;
    const int2 *links = via_last ? rec->right_of_last : rec->left_of_first;
    const bool direct = (dir != 0) == via_last;  // stored queries: append to last, prepend to first
    const int2 v = ld_hint_8(links + (direct ? b : 3u - b), ix.pol_hot);
    if (v.y < 0) return Coord{v.x, v.y};  // a miss: the caller does the real lookup
    return Coord{direct ? v.x : ~v.x, v.y};
}

// 8 bases starting at absolute base position p of the packed contig pool, as 16 bits
// (first base in the top two bits).
__device__ __forceinline__ uint32_t seq_window8(const DevIndex &ix, int64_t p)
{
    // positions outside the pool can only come from a corrupt index (undefined behaviour in
    // the reference); clamp so the read stays inside the allocation
    p = p < 0 ? 0 : (p > ix.n_bases ? ix.n_bases : p);
    const int64_t w = p >> 4;
    const int s = (int)(p & 15);
    const uint32_t hi = ld_hint_4(ix.seq2 + w, ix.pol_hot);
    uint32_t lo = 0;
    if (s > 8) lo = ld_hint_4(ix.seq2 + w + 1, ix.pol_hot);
    const uint64_t both = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(both >> (48 - 2 * s)) & 0xFFFFu;
}

}  // namespace skm
