// Device-resident class dictionary: ordered transcript-id tuple -> count, first-seen unit.
// Replaces MapResult.counter (mapper.py:54-75) on the device: open addressing on a 128-bit
// tuple hash (atom.cas.b128), SoA arrays, tuples copied once into an id pool.
#pragma once

#include "common.cuh"

namespace skm {

struct DictDev {
    ulonglong2 *keys;             // 128-bit tuple hash; all-ones = empty
    unsigned long long *counts;
    unsigned long long *first;    // smallest global unit index that produced the class
    uint32_t *pool_off;
    uint32_t *len;
    int32_t *pool;                // transcript ids of every class, tuple order
    uint64_t mask;                // slots - 1
    uint64_t pool_cap;
    unsigned long long *scalars;  // [0]=pool cursor [1]=n_classes [2]=unaligned [3]=aligned [4]=ids stored
    unsigned long long *fld;      // FLD_BINS
    uint32_t *status;
    uint32_t weak_keys;           // test hook (SKM_TEST_WEAK_KEYS): keys carry the tuple length only, so
                                  // different tuples collide and the id comparison below is exercised
};

struct DenseIds {
    const int32_t *p;
    __device__ __forceinline__ int32_t get(int i) const { return p[i]; }
};

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ ulonglong2 cas128(ulonglong2 *addr, ulonglong2 cmp, ulonglong2 val)
{
    ulonglong2 old;
    asm volatile(
        "{\n\t"
        ".reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}\n"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}

// 128-bit identity of an ordered id tuple (length included).  Never all-ones.
template <typename Ids>
__device__ __forceinline__ ulonglong2 tuple_key(const Ids &ids, int n, bool strip_sign)
{
    uint64_t h1 = 0x9E3779B97F4A7C15ULL ^ (uint64_t)n;
    uint64_t h2 = 0xD6E8FEB86659FD93ULL + (uint64_t)n;
    for (int i = 0; i < n; ++i) {
        int32_t e = ids.get(i);
        if (strip_sign && e < 0) e = ~e;  // _get_ids, _mapper.pyx:533-536
        const uint64_t v = (uint64_t)(uint32_t)e;
        h1 = mix64(h1 ^ v) + 0x632BE59BD9B4E019ULL;
        h2 = (h2 ^ (v + 0x9E3779B97F4A7C15ULL + (h2 << 6) + (h2 >> 2))) * 0xBF58476D1CE4E5B9ULL;
        h2 ^= h2 >> 29;
    }
    h2 = mix64(h2);
    if (h1 == EMPTY_KEY) h1 = 0;
    return make_ulonglong2(h1, h2);
}

template <typename Ids>
__device__ __forceinline__ ulonglong2 dict_key(const DictDev &d, const Ids &ids, int n, bool strip_sign)
{
    if (d.weak_keys) return make_ulonglong2((uint64_t)n, (uint64_t)n);
    return tuple_key(ids, n, strip_sign);
}

// A unit that found its key already in the table compares its ids with the stored tuple, so that
// equal keys of different tuples (probability ~ C^2 / 2^129) are refused instead of merged.  `len`
// is written last by the slot's owner (after a fence): 0 means the ids are not visible yet - only
// possible for a class inserted by this very launch - and the comparison is left to the class's
// later units.
template <typename Ids>
__device__ __forceinline__ void dict_verify_hit(const DictDev &d, int64_t s, const Ids &ids, int n, bool strip_sign)
{
    const uint32_t stored = *reinterpret_cast<volatile const uint32_t *>(&d.len[s]);
    if (stored == 0) return;
    bool same = stored == (uint32_t)n;
    if (same) {
        __threadfence();  // ids were written before len
        const volatile int32_t *p = d.pool + *reinterpret_cast<volatile const uint32_t *>(&d.pool_off[s]);
        for (int i = 0; i < n; ++i) {
            int32_t e = ids.get(i);
            if (strip_sign && e < 0) e = ~e;
            same = same && p[i] == e;
        }
    }
    if (!same) atomicOr(d.status, ST_KEY_COLLISION);
}

// Find the key's slot or claim an empty one; returns the slot, or -1 when the table is full.
// `won` tells the caller that it claimed the slot and owes the class its ids in the pool.
__device__ __forceinline__ int64_t dict_find_or_claim(const DictDev &d, ulonglong2 key, bool &won)
{
    won = false;
    uint64_t s = (key.x ^ (key.y >> 17)) & d.mask;
    const ulonglong2 empty = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
    for (uint64_t probes = 0; probes <= d.mask; ++probes) {
        // A slot is written once, by the 128-bit CAS, and never changes.  The 64-bit halves are
        // read individually (each atomic): a matching h1 followed by a matching h2 is this key
        // - the common case, no atomic needed; a foreign h1 moves on; anything else (empty, or
        // h1 seen before h2 became visible) goes through the CAS.
        const uint64_t seen = *reinterpret_cast<volatile const uint64_t *>(&d.keys[s].x);
        if (seen == key.x && *reinterpret_cast<volatile const uint64_t *>(&d.keys[s].y) == key.y) return (int64_t)s;
        if (seen == EMPTY_KEY || seen == key.x) {
            const ulonglong2 old = cas128(d.keys + s, empty, key);
            if (old.x == EMPTY_KEY && old.y == EMPTY_KEY) {
                won = true;
                return (int64_t)s;
            }
            if (old.x == key.x && old.y == key.y) return (int64_t)s;
        }
        s = (s + 1) & d.mask;
    }
    atomicOr(d.status, ST_DICT_FULL);
    return -1;
}

// The winner of a slot copies the tuple into the id pool at `off`; nobody reads it before the
// kernel ends.
template <typename Ids>
__device__ __forceinline__ void dict_store_ids(const DictDev &d, int64_t s, unsigned long long off, const Ids &ids,
                                               int n, bool strip_sign)
{
    if (off + (unsigned long long)n > d.pool_cap) {
        atomicOr(d.status, ST_POOL_FULL);
        d.pool_off[s] = 0;
        d.len[s] = 0;
        return;
    }
    for (int i = 0; i < n; ++i) {
        int32_t e = ids.get(i);
        if (strip_sign && e < 0) e = ~e;
        d.pool[off + i] = e;
    }
    d.pool_off[s] = (uint32_t)off;
    __threadfence();  // dict_verify_hit reads len first
    *reinterpret_cast<volatile uint32_t *>(&d.len[s]) = (uint32_t)n;
}

// Find-or-insert for callers without warp-level aggregation (the merge path).
template <typename Ids>
__device__ int64_t dict_find_or_insert(const DictDev &d, ulonglong2 key, const Ids &ids, int n,
                                       bool strip_sign)
{
    bool won;
    const int64_t s = dict_find_or_claim(d, key, won);
    if (won) {
        dict_store_ids(d, s, atomicAdd(&d.scalars[0], (unsigned long long)n), ids, n, strip_sign);
        atomicAdd(&d.scalars[1], 1ULL);
        atomicAdd(&d.scalars[4], (unsigned long long)n);
    } else if (s >= 0) {
        dict_verify_hit(d, s, ids, n, strip_sign);
    }
    return s;
}

}  // namespace skm
