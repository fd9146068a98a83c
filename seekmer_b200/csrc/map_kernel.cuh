// The mapping kernel: reads -> equivalence classes against the HBM-resident index.
//
// A block is a pool of ITEMS (a unit = read or pair, with the 2-bit codes of the current mate,
// its two target lists and 80 bytes of state) resident in shared memory, and a set of worker
// warps.  The reference's per-read state machine is cut at its memory accesses into PHASES:
//
//   P_LOAD    take the next unit (mate 0) and stage the packed read into shared memory
//   P_LOOKUP  KMerIndex.map_kmer (canonical form, hash, bucket probe) for the one pending k-mer:
//             the first k-mer of a scan, a walk fallback, a junction k-mer without a stored link
//   P_SCAN    _find_first_kmer after its first miss: the next SCAN_WIDTH read positions are
//             hashed and probed together (independent bucket loads in flight), first hit wins
//   P_CONTIG  contig record of a hit: map_contig for the scan hit, _filter_on_contig at a
//             junction (or the reload of the first contig before the right walk)
//   P_WALK    one head of the left/right contig-walk loops, or the final edge check: jump to
//             the contig edge, 8-base SIFT4 check (direction-generic); at a junction the next
//             contig comes from the record's graph link for the next read base
//   P_TALLY   map_read_pair mate intersection, span length; the unit's record goes to
//             tally_units_kernel (FLD, class dictionary)
//
// Item (row, lane) is only ever worked on by lane `lane` of some warp, so all of its shared
// memory ([field][row][lane]) is bank-conflict free.  For each phase and lane a 32-bit mask
// says which rows are waiting for that phase.  A warp iteration: every lane reads its masks,
// the warp votes for the phase most of its lanes can serve, each lane claims one waiting row
// of that phase (atomicAnd), loads the item state (four 16-byte shared loads), runs the phase,
// stores the state and sets the row's bit in the mask of the phase the item needs next
// (atomicOr).  Each heavy piece of code therefore exists once and runs with most lanes
// populated, instead of diverging 32 ways; a phase that still has rows for half the lanes runs
// again without a vote.  The kernel text is kept near the 32 KB the SM's instruction cache
// holds (the dictionary lives in its own kernel, slow paths are out of line): at 50+ KB this
// kernel was bound by instruction fetch.
//
// Reference semantics restated here (paths under /root/reference/seekmer/):
//   map_read            _mapper.pyx:151-193      find_first_kmer   :199-216
//   filter_to_left      _mapper.pyx:222-275      filter_to_right   :281-343
//   intersect (mates)   _mapper.pyx:350-397      map_read_pair     :111-145
//   sift4_align_left    _mapper.pyx:404-445      sift4_align_right :452-493
//   map_contig          _common.pyx:143-179      filter_on_contig  :185-235
//   get_contig_sequence _common.pyx:103-137      get_tail_kmer     :241-266
//   batch driver + FLD  _mapper.pyx:73-101       tuple ids         :528-537
#pragma once

#include "dict.cuh"
#include "kmer.cuh"
#include "sift4.cuh"

namespace skm {

#ifndef SKM_LIST_CAP
#define SKM_LIST_CAP 12
#endif
#ifndef SKM_Q_THREADS
#define SKM_Q_THREADS 768
#endif
constexpr int Q_THREADS = SKM_Q_THREADS;  // worker threads per block (one block per SM)
constexpr int LIST_CAP = SKM_LIST_CAP;    // per-read target list entries kept in shared memory
constexpr int ALIGN_LENGTH = 8;           // _mapper.pyx:22
constexpr int INVALID_SHIFT = SIFT4_INVALID_SHIFT;  // _mapper.pyx:28
#ifndef SKM_SCAN_WIDTH
#define SKM_SCAN_WIDTH 3
#endif
#ifndef SKM_COOP_PROBE
#define SKM_COOP_PROBE 0  // table buckets fetched by lane pairs, one warp instruction per bucket (lookup_warp)
#endif
constexpr int SCAN_WIDTH = SKM_SCAN_WIDTH;  // read positions probed per P_SCAN step
#ifndef SKM_IDLE_NS
#define SKM_IDLE_NS 200
#endif
#ifndef SKM_STICKY_LANES
#define SKM_STICKY_LANES 16
#endif
constexpr int STICKY_LANES = SKM_STICKY_LANES;  // a phase repeats while this many lanes still have rows for it
// how many lanes of the warp have a row waiting for a phase, given every lane's mask of waiting rows
__device__ __forceinline__ unsigned lanes_served(uint32_t m)
{
    return (unsigned)__popc(__ballot_sync(0xffffffffu, m != 0));
}

#ifndef SKM_STATS
#define SKM_STATS 0
#endif
// Diagnostics build (-DSKM_STATS=1, tools/build_variants.py): per phase [iterations, items claimed,
// unused, clock cycles of the iterations], then the idle polls; skm_debug_map_stats.
__device__ unsigned long long g_map_stats[32];

// What the mapper leaves behind for one unit, field-major so that tally_units_kernel reads it
// coalesced: units[0][u] = number of targets, units[1][u] = span length of _mapper.pyx:90,
// units[2 + k][u] = k-th target (signed entries; signs are stripped when the class key is
// formed, :528-537).  Lists longer than REC_IDS live in the arena: rows 2 and 3 then hold the
// low and high half of the arena offset.
constexpr int REC_IDS = 14;
constexpr int REC_ROWS = 2 + REC_IDS;

struct MapArgs {
    const uint64_t *packed;   // [n_reads][words] from pack_reads_kernel
    const int32_t *lens;      // per-read length, or NULL with fixed_len
    int32_t fixed_len;
    int32_t code_words;       // u64 words of 2-bit codes per read (from max read length)
    int32_t wild_words;       // u64 words of wildcard bits per read
    int32_t words;            // code_words + wild_words, rounded up to a multiple of four (32-byte records)
    int32_t paired;
    int64_t n_units;
    int64_t first_unit;
    int32_t *units;           // [REC_ROWS][n_units] mapping results, consumed by tally_units_kernel
    int32_t *arena;           // spill space for target lists longer than LIST_CAP
    uint64_t arena_cap;
    unsigned long long *cursors;  // [0]=work counter [1]=arena cursor
    unsigned long long *short_units;  // units with a read shorter than k so far (reported unaligned)
};

enum : int { P_LOAD = 0, P_SCAN, P_LOOKUP, P_CONTIG, P_WALK, P_TALLY, N_PHASES, P_DEAD = N_PHASES };
// who asked for the pending lookup / filter operation
enum : int {
    C_FIND = 0,  // _find_first_kmer scan (_mapper.pyx:199-216)
    C_LEFT_J,    // left walk junction (:247-251)
    C_LEFT_F,    // left walk fallback (:257-261)
    C_RIGHT_C,   // right walk start: contig of the cached first hit (:283-290)
    C_RIGHT_J    // right walk junction (:309-313)
};
// item flag bits
enum : uint32_t {
    F_CTX = 7u, F_DIR = 8u, F_MATE = 16u, F_ATTEMPT = 32u, F_FORWARD = 64u, F_L_ARENA = 128u,
    F_M1_ARENA = 256u, F_CTG_A0 = 512u, F_WILD = 1024u, F_VOID = 2048u
};
constexpr int STATE_VECS = 5;  // uint4 words of item state
constexpr int CTG_WORDS = 3;   // contig stash: first_kmer, last_kmer, seq_offset

constexpr size_t map_item_bytes(int code_words)
{
    return sizeof(uint64_t) * ((size_t)code_words + CTG_WORDS) + sizeof(int32_t) * 2 * LIST_CAP + 16 * STATE_VECS;
}
constexpr size_t map_fixed_bytes() { return sizeof(uint32_t) * (N_PHASES * 32) + 16; }

// ---- shared-memory views of one item; ITEMS = rows * 32 is a compile-time constant so every
// ---- field access is one ld/st.shared with an immediate offset
template <int ITEMS>
struct ReadView {
    const uint64_t *w;     // &codes[0][item]; word k at w[k * ITEMS]
    const uint64_t *wild;  // global wildcard words of the read, NULL when it has none
    int len;
    int wild_words;

    __device__ __forceinline__ uint64_t word(int k) const { return w[k * ITEMS]; }
    __device__ __forceinline__ uint32_t code(int p) const
    {
        return (uint32_t)(word(p >> 5) >> (62 - 2 * (p & 31))) & 3u;
    }
    // 25-mer starting at base p (_kmer.pxd:46-68)
    __device__ __forceinline__ uint64_t kmer(int p) const
    {
        const int k = p >> 5, s = p & 31;
        uint64_t x = word(k) << (2 * s);
        if (s > 7) x |= word(k + 1) >> (64 - 2 * s);
        return x >> 14;
    }
    // 9 read bases starting at base s >= 0, first base in bits 17:16
    __device__ __forceinline__ uint32_t codes9(int s) const
    {
        const int k = s >> 5, sh = s & 31;
        uint64_t x = word(k) << (2 * sh);
        if (sh > 23) x |= word(k + 1) >> (64 - 2 * sh);  // k + 1 <= code_words: in bounds
        return (uint32_t)(x >> 46);
    }
    // wildcard bits (_match_base, _mapper.pyx:500-501) of 9 read bases from s: bit i <-> base s + i
    __device__ __forceinline__ uint32_t wild9(int s) const
    {
        if (!wild) return 0u;
        const int k = s >> 6, b = s & 63;
        uint64_t x = __ldg(wild + k) >> b;
        if (b > 55 && k + 1 < wild_words) x |= __ldg(wild + k + 1) << (64 - b);
        return (uint32_t)x & 0x1FFu;
    }
};

// A target list: shared memory ([entry][item], element stride ITEMS) or, when longer than
// LIST_CAP, the arena (dense).  One generic pointer and a stride serve both, so element access
// is a single load/store without a branch.
template <int ITEMS>
struct List {
    int32_t *p;   // element 0
    int stride;   // ITEMS (shared memory) or 1 (arena)
    int n;
    __device__ __forceinline__ bool in_arena() const { return stride == 1; }
    __device__ __forceinline__ int32_t get(int i) const { return p[i * stride]; }
    __device__ __forceinline__ void set(int i, int32_t v) { p[i * stride] = v; }
};

struct Span {
    int begin, end;
    Coord anchor;
};

template <int ITEMS>
struct Lane {
    int st;
    // The item's flag word (F_* bits) stays packed in one register; the accessors below test and
    // set single bits, so nothing is unpacked on load or re-packed on store.
    uint32_t flags;
    bool m1_dirty;
    __device__ __forceinline__ int ctx() const { return (int)(flags & F_CTX); }
    __device__ __forceinline__ void set_ctx(int c) { flags = (flags & ~F_CTX) | (uint32_t)c; }
    __device__ __forceinline__ bool bit(uint32_t f) const { return (flags & f) != 0; }
    __device__ __forceinline__ void set_bit(uint32_t f, bool on) { flags = on ? flags | f : flags & ~f; }
    __device__ __forceinline__ int dir() const { return bit(F_DIR) ? 1 : 0; }
    __device__ __forceinline__ void set_dir(int d) { set_bit(F_DIR, d != 0); }
    __device__ __forceinline__ int mate() const { return bit(F_MATE) ? 1 : 0; }
    __device__ __forceinline__ void set_mate(int m) { set_bit(F_MATE, m != 0); }
    __device__ __forceinline__ int attempt() const { return bit(F_ATTEMPT) ? 1 : 0; }
    __device__ __forceinline__ void set_attempt(int v) { set_bit(F_ATTEMPT, v != 0); }
    __device__ __forceinline__ bool forward() const { return bit(F_FORWARD); }
    __device__ __forceinline__ void set_forward(bool b) { set_bit(F_FORWARD, b); }
    __device__ __forceinline__ bool ctg_a0() const { return bit(F_CTG_A0); }
    __device__ __forceinline__ void set_ctg_a0(bool b) { set_bit(F_CTG_A0, b); }
    __device__ __forceinline__ bool has_wild() const { return bit(F_WILD); }
    __device__ __forceinline__ void set_has_wild(bool b) { set_bit(F_WILD, b); }
    // one of the unit's reads is shorter than k: reported unaligned, span length 0
    __device__ __forceinline__ bool void_unit() const { return bit(F_VOID); }
    __device__ __forceinline__ void set_void_unit(bool b) { set_bit(F_VOID, b); }
    long long unit;
    int pos, move, len, clen;
    uint64_t kmer;
    Coord anchor0;
    Span sp;
    List<ITEMS> l;
    // mate 1 results while mate 2 is mapped
    int m1_begin, m1_len;
    Coord m1_anchor;
    List<ITEMS> m1;
};

template <int ITEMS>
struct ItemMem {
    uint4 *state;      // &state[0][item]; vector v at state[v * ITEMS]
    uint64_t *codes;   // &codes[0][item]
    uint64_t *ctg;     // &ctg[0][item]
    int32_t *list0;    // &lists[0][item]; mate 2 uses the second LIST_CAP entries
    int32_t *arena;
    __device__ __forceinline__ List<ITEMS> fresh_list(int mate) const
    {
        return List<ITEMS>{list0 + (mate ? LIST_CAP * ITEMS : 0), ITEMS, 0};
    }
};

__device__ __forceinline__ int sx16(uint32_t v) { return (int)(int16_t)(uint16_t)v; }

template <int ITEMS>
__device__ __forceinline__ void lane_load(Lane<ITEMS> &L, const ItemMem<ITEMS> &I, bool with_m1)
{
    const uint4 v0 = I.state[0], v1 = I.state[ITEMS], v2 = I.state[2 * ITEMS], v3 = I.state[3 * ITEMS];
    L.unit = (long long)v0.x;
    const uint32_t f = v0.y;
    L.flags = f;
    L.m1_dirty = false;
    L.pos = (int)(v0.z & 0xFFFFu);
    L.len = (int)(v0.z >> 16);
    L.move = (int)v0.w;
    L.kmer = (uint64_t)v1.x | ((uint64_t)v1.y << 32);
    L.sp.begin = (int)v1.z;
    L.sp.end = (int)v1.w;
    L.anchor0 = Coord{(int32_t)v2.x, (int32_t)v2.y};
    L.sp.anchor = Coord{(int32_t)v2.z, (int32_t)v2.w};
    L.l = I.fresh_list(L.mate());
    L.l.n = (int)(v3.x & 0xFFFFu);
    if (f & F_L_ARENA) L.l = List<ITEMS>{I.arena + v3.z, 1, L.l.n};
    L.m1 = I.fresh_list(0);
    L.m1.n = (int)(v3.x >> 16);
    if (f & F_M1_ARENA) L.m1 = List<ITEMS>{I.arena + v3.w, 1, L.m1.n};
    L.clen = (int)v3.y;
    if (with_m1) {
        const uint4 v4 = I.state[4 * ITEMS];
        L.m1_begin = sx16(v4.x);
        L.m1_len = (int)(v4.x >> 16);
        L.m1_anchor = Coord{(int32_t)v4.y, (int32_t)v4.z};
    }
}

template <int ITEMS>
__device__ __forceinline__ void lane_store(const Lane<ITEMS> &L, const ItemMem<ITEMS> &I)
{
    uint4 v0, v1, v2, v3;
    v0.x = (uint32_t)L.unit;
    v0.y = (L.flags & ~(F_L_ARENA | F_M1_ARENA)) | (L.l.in_arena() ? F_L_ARENA : 0u) | (L.m1.in_arena() ? F_M1_ARENA : 0u);
    v0.z = ((uint32_t)L.pos & 0xFFFFu) | ((uint32_t)L.len << 16);
    v0.w = (uint32_t)L.move;
    v1.x = (uint32_t)L.kmer;
    v1.y = (uint32_t)(L.kmer >> 32);
    v1.z = (uint32_t)L.sp.begin;
    v1.w = (uint32_t)L.sp.end;
    v2.x = (uint32_t)L.anchor0.entry;
    v2.y = (uint32_t)L.anchor0.offset;
    v2.z = (uint32_t)L.sp.anchor.entry;
    v2.w = (uint32_t)L.sp.anchor.offset;
    v3.x = ((uint32_t)L.l.n & 0xFFFFu) | ((uint32_t)L.m1.n << 16);
    v3.y = (uint32_t)L.clen;
    v3.z = L.l.in_arena() ? (uint32_t)(L.l.p - I.arena) : 0u;
    v3.w = L.m1.in_arena() ? (uint32_t)(L.m1.p - I.arena) : 0u;
    I.state[0] = v0;
    I.state[ITEMS] = v1;
    I.state[2 * ITEMS] = v2;
    I.state[3 * ITEMS] = v3;
    if (L.m1_dirty) {
        uint4 v4;
        v4.x = ((uint32_t)L.m1_begin & 0xFFFFu) | ((uint32_t)L.m1_len << 16);
        v4.y = (uint32_t)L.m1_anchor.entry;
        v4.z = (uint32_t)L.m1_anchor.offset;
        v4.w = 0;
        I.state[4 * ITEMS] = v4;
    }
}

// ---- index-side helpers ------------------------------------------------------------------
// get_contig_sequence(coordinate, +-8) as a 16-bit window
__device__ __forceinline__ uint32_t contig_window(const DevIndex &ix, int64_t seq_offset, Coord a, bool left_edge)
{
    const int64_t p = seq_offset + a.offset;
    // one load site for both strands: the window sits at the k-mer's low end when the edge asked
    // for and the strand agree, at its high end otherwise; the reverse strand complements it
    const bool forward = a.entry >= 0;
    const uint32_t w = seq_window8(ix, left_edge == forward ? p : p + K - ALIGN_LENGTH);
    return forward ? w : revcomp8(w);
}

// 8-base window at a contig EDGE, taken from the record's first/last k-mer instead of the
// sequence pool: inside the walk loops the anchor always sits on the first or last k-mer of
// its contig (offset 0 or length-k, _mapper.pyx:229-236,289-295), and first_kmer/last_kmer
// are the encodings of the contig's first/last 25 bases (_index_builder.pyx:565-567).
__device__ __forceinline__ uint32_t edge_window(uint64_t first_kmer, uint64_t last_kmer, Coord a, bool left_edge)
{
    const uint32_t head = (uint32_t)(first_kmer >> (2 * K - 16)) & 0xFFFFu;  // first 8 bases
    const uint32_t tail = (uint32_t)last_kmer & 0xFFFFu;                      // last 8 bases
    if (a.entry >= 0) return left_edge ? head : tail;
    return revcomp8(left_edge ? tail : head);
}

// map_contig for a contig with more than 8 targets: the full list is read from targets[]; more
// than LIST_CAP entries spill to the arena.  Out of line (rare) with scalar arguments only, so
// nothing of the caller's state is forced into local memory.  Returns the arena offset used,
// -1 when the list went to the caller's shared-memory list, -2 when the arena is exhausted.
__device__ __noinline__ long long map_contig_long(const int32_t *t, int n, int forward, int32_t *p, int stride,
                                                  int32_t *arena, unsigned long long arena_cap,
                                                  unsigned long long *cursor, uint32_t *status)
{
    const int32_t x = forward ? 0 : -1;
    long long off = -1;
    if (n > LIST_CAP) {
        off = (long long)atomicAdd(cursor, (unsigned long long)n);
        if ((unsigned long long)off + (unsigned long long)n > arena_cap) {
            atomicOr(status, ST_ARENA_FULL);
            return -2;
        }
        p = arena + off;
        stride = 1;
    }
#pragma unroll 1
    for (int i = 0; i < n; ++i) p[i * stride] = __ldg(t + (forward ? i : n - 1 - i)) ^ x;
    return off;
}

// entries 8..15 of a contig's target list (zero beyond its end), for lists of 9..16 targets
__device__ __forceinline__ void load_targets_hi(const DevIndex &ix, const Contig &c, int32_t (&t)[INLINE_TARGETS])
{
    const int32_t *src = ix.targets + c.target_offset + INLINE_TARGETS;
#pragma unroll
    for (int j = 0; j < INLINE_TARGETS; ++j)
        t[j] = INLINE_TARGETS + j < c.target_count ? (int32_t)ld_hint_4(src + j, ix.pol_hot) : 0;
}

// map_contig (_common.pyx:143-179) for an already loaded contig record.  Forward: the
// contig's targets in order; reverse: reversed order, every entry bit-negated.
// `sp` is the list's home in shared memory (&lists[0][item] of the mate): the short path writes
// through it, so that the stores are plain STS with immediate offsets (l.p may also point into the
// arena, which makes every access through it a generic one).
template <int ITEMS>
__device__ __forceinline__ void map_contig(const DevIndex &ix, const MapArgs &a, uint32_t *status, const Contig &c,
                                           Coord at, List<ITEMS> &l, int32_t *sp)
{
    const bool forward = at.entry >= 0;
    const int n = c.target_count;
    const int32_t x = forward ? 0 : -1;
    l.n = n;
    if (n <= 2 * INLINE_TARGETS && n <= LIST_CAP) {
        // straight-line: the first 8 entries came with the record, 8 more are fetched if needed
#pragma unroll
        for (int j = 0; j < INLINE_TARGETS; ++j)
            if (j < n) sp[(forward ? j : n - 1 - j) * ITEMS] = c.t[j] ^ x;
        if (n > INLINE_TARGETS) {
            int32_t t[INLINE_TARGETS];
            load_targets_hi(ix, c, t);
#pragma unroll
            for (int j = 0; j < INLINE_TARGETS; ++j)
                if (INLINE_TARGETS + j < n) sp[(forward ? INLINE_TARGETS + j : n - 1 - INLINE_TARGETS - j) * ITEMS] = t[j] ^ x;
        }
        return;
    }
    const long long off = map_contig_long(ix.targets + c.target_offset, n, forward ? 1 : 0, l.p, l.stride, a.arena,
                                          a.arena_cap, a.cursors + 1, status);
    if (off == -2) l.n = 0;
    else if (off >= 0) l = List<ITEMS>{a.arena + off, 1, n};
}

// the sorted merge itself, for contigs with more than 8 targets (list read from targets[]);
// returns the new list length, 0 = no match (list left intact)
__device__ __noinline__ int filter_long(const int32_t *t, int length, int forward, int32_t *p, int stride, int n)
{
    const int32_t x = forward ? 0 : -1;
    int read_index = 0, write_index = 0, track = 0;
    int32_t index_entry = __ldg(t + (forward ? 0 : length - 1)) ^ x;
    int32_t target_entry = p[0];
    while (true) {
        if (target_entry == index_entry) {
            p[write_index * stride] = target_entry;
            read_index += 1;
            write_index += 1;
            track += 1;
            if (read_index == n || track == length) break;
            target_entry = p[read_index * stride];
            index_entry = __ldg(t + (forward ? track : length - 1 - track)) ^ x;
        } else if (target_entry < index_entry) {
            read_index += 1;
            if (read_index == n) break;
            target_entry = p[read_index * stride];
        } else {
            track += 1;
            if (track == length) break;
            index_entry = __ldg(t + (forward ? track : length - 1 - track)) ^ x;
        }
    }
    return write_index;
}

// _filter_on_contig (_common.pyx:185-235): direction-aware sorted-merge intersection of the
// span's list with the contig's list; equal entries pair off one to one; zero matches leave
// the list intact and return false.
template <int ITEMS>
__device__ __forceinline__ bool filter_on_contig(const DevIndex &ix, const Contig &c, Coord at, List<ITEMS> &l,
                                                 int32_t *sp)
{
    if (l.n == 0) return true;
    const bool forward = at.entry >= 0;
    const int length = c.target_count;
    if (length == 0) return false;
    const int32_t x = forward ? 0 : -1;
    int w;
    if (length <= 2 * INLINE_TARGETS && !l.in_arena()) {
        // Both lists are ascending (targets are sorted per contig, _index_builder.pyx:540, and
        // map_contig / this filter keep that order), so the merge keeps the r-th occurrence of a
        // value v in the span's list exactly when the contig's list holds more than r copies of
        // v.  Counting against 8 registers needs no data-dependent control flow.
        // An element whose value differs from its predecessor's (no duplicate: nearly always)
        // only needs to know whether the value occurs at all: 8 compare-and-accumulate
        // instructions against registers, unused slots holding a copy of entry 0.  Duplicates
        // take the exact count.
        int32_t t[INLINE_TARGETS], t2[INLINE_TARGETS];
#pragma unroll
        for (int j = 0; j < INLINE_TARGETS; ++j) t[j] = (j < length ? c.t[j] : c.t[0]) ^ x;
        const bool more = length > INLINE_TARGETS;  // 9..16 targets: 8 more from targets[]
#pragma unroll
        for (int j = 0; j < INLINE_TARGETS; ++j) t2[j] = t[0];
        if (more) {
            load_targets_hi(ix, c, t2);
#pragma unroll
            for (int j = 0; j < INLINE_TARGETS; ++j) t2[j] = (INLINE_TARGETS + j < length ? t2[j] : c.t[0]) ^ x;
        }
        int run = 0;
        int32_t prev = 0;
        const int n = l.n;
        w = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int32_t v = sp[i * ITEMS];
            run = (i > 0 && v == prev) ? run + 1 : 0;
            prev = v;
            bool keep = false;
#pragma unroll
            for (int j = 0; j < INLINE_TARGETS; ++j) keep |= t[j] == v;
            if (more) {
#pragma unroll
                for (int j = 0; j < INLINE_TARGETS; ++j) keep |= t2[j] == v;
            }
            if (run > 0 && keep) {  // the r-th copy of v stays iff the contig lists more than r copies
                int copies = 0;
#pragma unroll
                for (int j = 0; j < INLINE_TARGETS; ++j)
                    copies += (j < length && t[j] == v) + (INLINE_TARGETS + j < length && t2[j] == v);
                keep = run < copies;
            }
            if (keep) {
                sp[w * ITEMS] = v;
                w += 1;
            }
        }
    } else {
        w = filter_long(ix.targets + c.target_offset, length, forward ? 1 : 0, l.p, l.stride, l.n);
    }
    if (w == 0) return false;
    l.n = w;
    return true;
}

// mate intersection (_mapper.pyx:350-397): list 1 ascending vs list 2 descending, negated.
// Returns the new length of list 1 (0 = nothing in common, list left intact).
__device__ __noinline__ int intersect_generic(int32_t *p1, int stride1, int n1, const int32_t *p2, int stride2, int n2)
{
    int cursor1_read = 0, cursor1_write = 0, cursor2 = n2 - 1;
#pragma unroll 1
    while (cursor1_read != n1 && cursor2 != -1) {
        const int32_t entry1 = p1[cursor1_read * stride1];
        const int32_t entry2 = ~p2[cursor2 * stride2];
        if (entry1 == entry2) {
            p1[cursor1_write * stride1] = entry1;
            cursor1_read += 1;
            cursor1_write += 1;
            cursor2 -= 1;
        } else if (entry1 < entry2) {
            cursor1_read += 1;
        } else {
            cursor2 -= 1;
        }
    }
    return cursor1_write;
}

// sp1 / sp2: the shared-memory homes of the two lists (mate 1: &lists[0][item], mate 2: LIST_CAP rows on)
template <int ITEMS>
__device__ __forceinline__ bool intersect(List<ITEMS> &l1, const List<ITEMS> &l2, int32_t *sp1, const int32_t *sp2)
{
    if (l1.n == 0) return true;
    if (l2.n == 0) return false;
    int cursor1_write;
    if (l1.in_arena() || l2.in_arena()) {
        cursor1_write = intersect_generic(l1.p, l1.stride, l1.n, l2.p, l2.stride, l2.n);
    } else {
        int cursor1_read = 0, cursor2 = l2.n - 1;
        cursor1_write = 0;
#pragma unroll 1
        while (cursor1_read != l1.n && cursor2 != -1) {
            const int32_t entry1 = sp1[cursor1_read * ITEMS];
            const int32_t entry2 = ~sp2[cursor2 * ITEMS];
            if (entry1 == entry2) {
                sp1[cursor1_write * ITEMS] = entry1;
                cursor1_read += 1;
                cursor1_write += 1;
                cursor2 -= 1;
            } else if (entry1 < entry2) {
                cursor1_read += 1;
            } else {
                cursor2 -= 1;
            }
        }
    }
    if (cursor1_write == 0) return false;
    l1.n = cursor1_write;
    return true;
}

// ---- state-machine transitions ------------------------------------------------------------
// A phase ends by naming what happens next; the transitions of map_read (:151-193) and of the
// walk fallbacks are applied once, after the phase switch, by every lane together.
enum : int {
    EV_NONE = 0,
    EV_LEFT_FAILED,    // _filter_targets_to_left :250-263: junction lookup or its filter failed
    EV_AFTER_LEFT,     // map_read :174-176
    EV_AFTER_ATTEMPT,  // map_read :177-193
    EV_READ_DONE
};

// One k-mer lookup, out of line so that P_LOOKUP and the four probes of P_SCAN share one copy
// of the code; the result is packed as entry | offset << 32.
__device__ __noinline__ unsigned long long lookup_packed(const Slot *table, uint64_t bucket_mask, uint64_t canon,
                                                         int fwd, uint32_t bucket)
{
    const Coord c = probe_canonical(table, bucket_mask, canon, fwd != 0, bucket);
    return (unsigned long long)(uint32_t)c.entry | ((unsigned long long)(uint32_t)c.offset << 32);
}

struct Probe {
    uint64_t canon;
    uint32_t bucket;
    bool fwd;
};

__device__ __forceinline__ Probe prepare_probe(uint64_t kmer, uint64_t bucket_mask)
{
    const uint64_t rc = revcomp(kmer);
    Probe p;
    p.fwd = kmer < rc;
    p.canon = p.fwd ? kmer : rc;
    p.bucket = home_bucket_of(p.canon, bucket_mask);
    return p;
}

__device__ __forceinline__ Coord run_probe(const DevIndex &ix, const Probe &p)
{
    const unsigned long long v = lookup_packed(ix.table, ix.bucket_mask, p.canon, p.fwd ? 1 : 0, p.bucket);
    return Coord{(int32_t)(uint32_t)v, (int32_t)(uint32_t)(v >> 32)};
}

// The pending k-mers of a whole warp, probed TOGETHER (all 32 lanes call this converged; `active`
// says which lanes have a k-mer).  Why: the table (4.3 GB) is far beyond the 256 MB the TLB
// reaches, and what a random probe costs there is one page walk per lane REQUEST, not bytes
// (tools/rand_access_bench2.cu on B200, 4 GB table: a lane fetching its own 64-byte bucket with
// 4 x LDG.128: 15.9 G buckets/s; the same buckets fetched by lane pairs, one LDG.256 each, so that
// a bucket is touched by ONE warp instruction: 37.7 G buckets/s, the walker's ceiling).  Two rounds
// of 16 owners; lanes 2o and 2o+1 fetch the two halves of owner o's bucket, both rounds' loads are
// issued before the first compare.  Returns the slot's value word (entry | offset << 32) for the
// canonical key, or offset -1; the caller applies the strand.  Same probe sequence and the same
// "last slot empty = bucket not full = stop" rule as probe_canonical (kmer.cuh).
constexpr unsigned long long LOOKUP_MISS = 0xFFFFFFFF00000000ULL;
__device__ __noinline__ unsigned long long lookup_warp(const Slot *table, uint64_t bucket_mask, uint64_t canon,
                                                       uint32_t bucket, int active)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned half = lane & 1u, pair = lane >> 1, own = 2u * (lane & 15u);
    unsigned long long result = LOOKUP_MISS;
    unsigned pending = __ballot_sync(0xffffffffu, active);
    while (pending) {
        Quad64 s[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned src = (unsigned)r * 16u + pair;
            const uint32_t b = __shfl_sync(0xffffffffu, bucket, src);
            s[r] = Quad64{EMPTY_KEY, 0, EMPTY_KEY, 0};
            if ((pending >> src) & 1u) s[r] = ld_cs_32(table + (uint64_t)b * BUCKET_SLOTS + 2u * half);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned src = (unsigned)r * 16u + pair;
            const uint64_t c = __shfl_sync(0xffffffffu, canon, src);
            const bool h0 = s[r].a == c, h1 = s[r].c == c;
            const unsigned hits = __ballot_sync(0xffffffffu, h0 || h1);
            const unsigned open = __ballot_sync(0xffffffffu, s[r].c == EMPTY_KEY);  // odd lanes: the bucket's last slot
            const uint64_t v = h0 ? s[r].b : s[r].d;
            const unsigned got = (hits >> own) & 3u;
            const uint64_t word = __shfl_sync(0xffffffffu, v, own + (got >> 1));
            if ((lane >> 4) == (unsigned)r && active) {
                if (got) {
                    result = word;
                    active = 0;
                } else if ((open >> (own + 1u)) & 1u) {
                    active = 0;
                } else {
                    bucket = (bucket + 1u) & (uint32_t)bucket_mask;
                }
            }
        }
        pending = __ballot_sync(0xffffffffu, active);
    }
    return result;
}

__device__ __forceinline__ Coord run_probe_warp(const DevIndex &ix, const Probe &p, bool active)
{
    const unsigned long long v = lookup_warp(ix.table, ix.bucket_mask, p.canon, p.bucket, active ? 1 : 0);
    const int32_t entry = (int32_t)(uint32_t)v, offset = (int32_t)(uint32_t)(v >> 32);
    if (offset < 0) return coord_invalid();
    return Coord{p.fwd ? entry : ~entry, offset};
}

// A final list longer than a unit record: make sure it lives in the arena, return its offset.
__device__ __noinline__ long long emit_long_list(const int32_t *p, int stride, int n, int32_t *arena,
                                                 unsigned long long arena_cap, unsigned long long *cursor,
                                                 uint32_t *status)
{
    if (stride == 1) return (long long)(p - arena);
    const unsigned long long off = atomicAdd(cursor, (unsigned long long)n);
    if (off + (unsigned long long)n > arena_cap) {
        atomicOr(status, ST_ARENA_FULL);
        return 0;
    }
#pragma unroll 1
    for (int i = 0; i < n; ++i) arena[off + i] = p[i * stride];
    return (long long)off;
}

// sift4_align_left(window, read, qoff) for dir == 0, sift4_align_right for dir == 1 (sift4.cuh)
template <int ITEMS>
__device__ __forceinline__ int sift4_edge(uint32_t ref16, const ReadView<ITEMS> &rv, int qoff, int dir)
{
    int s = qoff - (1 - dir);  // the left routine may look one base left of its window (:421)
    const int pad = s < 0 ? 1 : 0;
    s += pad;
    uint32_t codes = rv.codes9(s) >> (2 * pad);
    uint32_t wild = (rv.wild9(s) << pad) & 0x1FFu;
    if (dir) {
        wild = reverse_bits(wild, 9);
    } else {
        codes = reverse_pairs(codes, 9);
        ref16 = reverse_pairs(ref16, 8);
    }
    return sift4_unified(ref16, codes, wild, 1 - dir, dir ? rv.len - qoff : qoff + 8);
}

template <int ROWS>
__global__ void __launch_bounds__(Q_THREADS, 1)
map_reads_kernel(const DevIndex ix, const MapArgs a, uint32_t *const status)
{
    constexpr int ITEMS = ROWS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sm_state = reinterpret_cast<uint4 *>(smem_raw);                            // [STATE_VECS][ITEMS]
    uint64_t *sm_codes = reinterpret_cast<uint64_t *>(sm_state + STATE_VECS * ITEMS);  // [code_words][ITEMS]
    uint64_t *sm_ctg = sm_codes + (size_t)a.code_words * ITEMS;                       // [CTG_WORDS][ITEMS]
    int32_t *sm_lists = reinterpret_cast<int32_t *>(sm_ctg + CTG_WORDS * ITEMS);      // [2 * LIST_CAP][ITEMS]
    uint32_t *sm_masks = reinterpret_cast<uint32_t *>(sm_lists + 2 * LIST_CAP * ITEMS);  // [N_PHASES][32]
    int *sm_live = reinterpret_cast<int *>(sm_masks + N_PHASES * 32);

    for (int i = threadIdx.x; i < N_PHASES * 32; i += blockDim.x)
        sm_masks[i] = i < 32 ? (ROWS == 32 ? 0xFFFFFFFFu : (1u << ROWS) - 1u) : 0u;  // everything in P_LOAD
    for (int i = threadIdx.x; i < STATE_VECS * ITEMS; i += blockDim.x) sm_state[i] = make_uint4(0, 0, 0, 0);  // mate 0
    if (threadIdx.x == 0) *sm_live = ITEMS;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    volatile uint32_t *vmasks = sm_masks;
    unsigned iter = (unsigned)warp * 5u;
    int phase = -1;

    for (;;) {
        // ---- vote: the phase most lanes have a waiting row for ---------------------------
        // ... unless the phase just run still has rows for most lanes: then it runs again (its
        // code is hot in the instruction cache and the full vote is skipped)
        uint32_t mm = 0;
        bool again = false;
        if (phase >= 0) {
            mm = vmasks[phase * 32 + lane];
            again = lanes_served(mm) >= (unsigned)STICKY_LANES;
        }
        if (!again) {
            uint32_t m[N_PHASES];
            unsigned best = 0;
#pragma unroll
            for (int p = 0; p < N_PHASES; ++p) {
                m[p] = vmasks[p * 32 + lane];
                const unsigned c = lanes_served(m[p]);
                const unsigned cand = c ? (c << 3) | (unsigned)p : 0u;
                best = cand > best ? cand : best;
            }
            if (best == 0) {
                int live = 0;
                if (lane == 0) live = *reinterpret_cast<volatile int *>(sm_live);
                live = __shfl_sync(0xffffffffu, live, 0);
                if (live == 0) break;
#if SKM_STATS
                if (lane == 0) atomicAdd(&g_map_stats[N_PHASES * 4], 1ULL);
#endif
                __nanosleep(SKM_IDLE_NS);
                phase = -1;
                continue;
            }
            phase = (int)(best & 7u);
#pragma unroll
            for (int p = 0; p < N_PHASES; ++p)
                if (p == phase) mm = m[p];
        }
        // ---- claim one waiting row of that phase ---------------------------------------------
        bool mine = false;
        int row = 0;
        if (mm) {
            const unsigned rot = iter & 31u;
            const uint32_t mr = __funnelshift_r(mm, mm, rot);
            row = (int)((unsigned)(__ffs((int)mr) - 1) + rot) & 31;
            const uint32_t old = atomicAnd(&sm_masks[phase * 32 + lane], ~(1u << row));
            mine = (old >> row) & 1u;
        }
        iter += 1;
#if SKM_STATS
        const long long stat_t0 = clock64();
        const unsigned stat_got = __ballot_sync(0xffffffffu, mine);
#endif
        __threadfence_block();
        const int item = row * 32 + lane;
        ItemMem<ITEMS> I;
        I.state = sm_state + item;
        I.codes = sm_codes + item;
        I.ctg = sm_ctg + item;
        I.list0 = sm_lists + item;
        I.arena = a.arena;
        Lane<ITEMS> L;
        L.st = phase;
        L.flags = 0;
        L.unit = 0;
        L.len = 0;
        if (mine) lane_load(L, I, phase == P_TALLY);
        ReadView<ITEMS> rv;
        rv.w = I.codes;
        rv.len = L.len;
        rv.wild_words = a.wild_words;
        rv.wild = nullptr;
        if (L.has_wild()) {
            const long long read_idx = a.paired ? 2 * L.unit + L.mate() : L.unit;
            rv.wild = a.packed + read_idx * (long long)a.words + a.code_words;
        }
        int ev = EV_NONE;
        int want_pos = -1;       // look up the read's k-mer at this position next ...
        bool want = false;       // ... or this explicit k-mer
        uint64_t want_kmer = 0;

        if (phase == P_LOAD) {
            // ---- new units for finished items (mate 0): one global atomic per warp -------------
            const bool need = mine && L.mate() == 0;
            const unsigned nb = __ballot_sync(0xffffffffu, need);
            if (nb) {
                long long base = 0;
                if (lane == 0) base = (long long)atomicAdd(&a.cursors[0], (unsigned long long)__popc(nb));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (need) {
                    L.unit = base + __popc(nb & ((1u << lane) - 1u));
                    if (L.unit >= a.n_units) L.st = P_DEAD;
                }
            }
            if (mine && L.st == P_LOAD) {
                const long long read_idx = a.paired ? 2 * L.unit + L.mate() : L.unit;
                const uint64_t *src = a.packed + read_idx * (long long)a.words;
                uint64_t any_wild = 0;
                for (int k = 0; k < a.words; k += 4) {  // 32-byte records: words is a multiple of four
                    const Quad64 v = ld_cs_32(src + k);  // one LDG.E.256 per four words, read once
                    if (k < a.code_words) I.codes[k * ITEMS] = v.a;
                    else any_wild |= v.a;  // wildcard words; the padding words are zero
                    if (k + 1 < a.code_words) I.codes[(k + 1) * ITEMS] = v.b;
                    else any_wild |= v.b;
                    if (k + 2 < a.code_words) I.codes[(k + 2) * ITEMS] = v.c;
                    else any_wild |= v.c;
                    if (k + 3 < a.code_words) I.codes[(k + 3) * ITEMS] = v.d;
                    else any_wild |= v.d;
                }
                int len = a.lens ? __ldg(a.lens + read_idx) : a.fixed_len;
                const int max_len = a.code_words * 32;
                if (len > max_len) len = max_len;
                L.len = len;
                rv.len = len;
                L.set_has_wild(any_wild != 0);
                L.sp = Span{0, 0, coord_invalid()};
                L.l = I.fresh_list(L.mate());
                L.set_attempt(0);
                L.pos = 0;
                L.set_ctg_a0(false);
                L.set_ctx(C_FIND);
                if (len >= K) {
                    want_pos = 0;
                } else {
                    // A read shorter than k is undefined in the reference (_kmer.pxd:46-68 reads past its
                    // end).  Here its unit is reported unaligned with span length 0, whatever the mate does.
                    atomicOr(status, ST_SHORT_READ);
                    atomicAdd(a.short_units, 1ULL);
                    L.set_void_unit(true);
                    L.l.n = 0;
                    L.st = P_TALLY;
                }
            }
        } else if (phase == P_LOOKUP) {
#if SKM_COOP_PROBE
            Probe pr{0, 0, false};
            if (mine) pr = prepare_probe(L.kmer, ix.bucket_mask);
            const Coord h = run_probe_warp(ix, pr, mine);
            if (mine) {
#else
            if (mine) {
                const Coord h = run_probe(ix, prepare_probe(L.kmer, ix.bucket_mask));
#endif
                L.sp.anchor = h;
                if (h.offset >= 0) {
                    L.st = P_CONTIG;
                } else if (L.ctx() == C_FIND) {
                    // _find_first_kmer keeps rolling (:208-212); an exhausted scan leaves the
                    // targets empty and map_read returns (:170-171, :186-187)
                    L.pos += 1;
                    if (L.pos + K <= L.len) L.st = P_SCAN;
                    else ev = EV_READ_DONE;
                } else if (L.ctx() == C_RIGHT_J) {
                    L.l.n = 0;  // :312-315
                    ev = EV_AFTER_ATTEMPT;
                } else {
                    ev = EV_LEFT_FAILED;
                }
            }
        } else if (phase == P_SCAN) {
#if SKM_COOP_PROBE
            // positions pos .. pos+2 (while they fit), one after the other, first hit wins; L.kmer is
            // the k-mer at pos-1.  Every step is one warp-wide probe (lookup_warp) of the lanes still
            // scanning.
            uint64_t kk = 0;
            int fit = 0, first = SCAN_WIDTH;
            Coord hh = coord_invalid();
            if (mine) {
                kk = L.kmer;
                fit = L.len - K + 1 - L.pos;  // >= 1
            }
#pragma unroll 1
            for (int j = 0; j < SCAN_WIDTH; ++j) {
                const bool go = mine && j < fit && first == SCAN_WIDTH;
                if (!__any_sync(0xffffffffu, go)) break;
                Probe pr{0, 0, false};
                if (go) {
                    kk = ((kk << 2) | rv.code(L.pos + j + K - 1)) & KMER_MASK;
                    pr = prepare_probe(kk, ix.bucket_mask);
                }
                const Coord h = run_probe_warp(ix, pr, go);
                if (go && h.offset >= 0) {
                    first = j;
                    hh = h;
                }
            }
            if (mine) {
#else
            if (mine) {
                // positions pos .. pos+2 (while they fit), one after the other, first hit wins;
                // L.kmer is the k-mer at pos-1.  (One copy of the hash + probe code instead of three:
                // the kernel text sits at the edge of what the instruction caches hold)
                uint64_t kk = L.kmer;
                const int fit = L.len - K + 1 - L.pos;  // >= 1
                int first = SCAN_WIDTH;
                Coord hh = coord_invalid();
#pragma unroll 1
                for (int j = 0; j < SCAN_WIDTH && j < fit; ++j) {
                    kk = ((kk << 2) | rv.code(L.pos + j + K - 1)) & KMER_MASK;
                    const Coord h = run_probe(ix, prepare_probe(kk, ix.bucket_mask));
                    if (h.offset >= 0) {
                        first = j;
                        hh = h;
                        break;
                    }
                }
#endif
                L.sp.anchor = hh;
                L.kmer = kk;
                if (first < SCAN_WIDTH) {
                    L.pos += first;
                    L.st = P_CONTIG;
                } else {
                    L.pos += min(SCAN_WIDTH, fit);
                    if (L.pos + K <= L.len) L.st = P_SCAN;
                    else ev = EV_READ_DONE;
                }
            }
        } else if (phase == P_CONTIG) {
            if (mine) {
                const Coord at = L.ctx() == C_RIGHT_C ? L.anchor0 : L.sp.anchor;
                const Contig c = load_contig(ix, at.entry >= 0 ? at.entry : ~at.entry);
                I.ctg[0] = c.first_kmer;
                I.ctg[ITEMS] = c.last_kmer;
                I.ctg[2 * ITEMS] = (uint64_t)c.seq_offset;
                L.clen = c.length;
                L.set_forward(at.entry >= 0);
                const int to_start = L.forward() ? at.offset : c.length - at.offset - K;
                const int to_end = L.forward() ? c.length - at.offset - K : at.offset;
                if (L.ctx() == C_FIND) {
                    map_contig(ix, a, status, c, at, L.l, I.list0 + (L.mate() ? LIST_CAP * ITEMS : 0));
                    L.sp.begin = L.pos;
                    L.sp.end = L.pos;
                    L.anchor0 = at;
                    L.set_ctg_a0(true);
                    if (L.l.n == 0) {
                        ev = EV_READ_DONE;  // `if is_empty(targets): return span`
                    } else if (L.sp.begin > 0) {
                        L.move = to_start;
                        L.set_dir(0);
                        L.st = P_WALK;
                    } else {
                        ev = EV_AFTER_LEFT;
                    }
                } else {
                    bool ok = true;
                    if (L.ctx() != C_RIGHT_C) {
                        ok = filter_on_contig(ix, c, at, L.l, I.list0 + (L.mate() ? LIST_CAP * ITEMS : 0));
                        L.set_ctg_a0(false);
                    } else {
                        L.sp.anchor = at;  // :283-284 — same k-mer as the scan hit, lookup cached
                        L.set_ctg_a0(true);
                    }
                    if (L.ctx() == C_RIGHT_C || L.ctx() == C_RIGHT_J) {
                        if (ok) {
                            L.move = to_end;
                            L.set_dir(1);
                            L.st = P_WALK;
                        } else {
                            L.l.n = 0;  // :312-315
                            ev = EV_AFTER_ATTEMPT;
                        }
                    } else if (ok) {
                        L.move = to_start;
                        L.set_dir(0);
                        L.st = P_WALK;
                    } else {
                        ev = EV_LEFT_FAILED;
                    }
                }
            }
        } else if (phase == P_WALK) {
            if (mine) {
                // heads of the loops of _filter_targets_to_left (:234-275) and _to_right (:293-343)
                const uint64_t first_kmer = I.ctg[0], last_kmer = I.ctg[ITEMS];
                const int dir = L.dir();
                int rem = dir ? L.len - L.sp.end - K : L.sp.begin;  // bases left towards the read end
                const bool in_loop = rem > L.move;
                const int step = in_loop ? L.move : rem;
                const int delta = L.forward() ? step : -step;
                L.sp.anchor.offset += dir ? delta : -delta;
                uint32_t ref16;
                int qoff;
                if (in_loop) {
                    rem -= L.move;
                    ref16 = edge_window(first_kmer, last_kmer, L.sp.anchor, dir == 0);
                    qoff = dir ? L.len - rem - ALIGN_LENGTH : rem;
                } else {
                    ref16 = contig_window(ix, (int64_t)I.ctg[2 * ITEMS], L.sp.anchor, dir == 0);
                    qoff = dir ? L.len - ALIGN_LENGTH : 0;
                }
                const int shift = sift4_edge(ref16, rv, qoff, dir);
                bool finished = true;  // this direction is over (success or failure)
                if (in_loop) {
                    if (shift == INVALID_SHIFT || shift + 1 + L.move <= 0) {
                        L.l.n = 0;
                    } else {
                        rem -= shift + 1;
                        if (rem < 0) rem = 0;  // :244-246 / :306-308, list intact
                        else finished = false;
                    }
                    if (L.l.n != 0) {
                        if (dir) L.sp.end = L.len - rem - K;
                        else L.sp.begin = rem;
                    } else if (!dir) {
                        L.sp.begin = rem;  // the failed left walk leaves begin where it stopped (:235,241-242)
                    } else {
                        L.sp.end = L.len - rem - K;
                    }
                    if (!finished) {
                        // the junction k-mer is the contig's edge k-mer shifted by one read base
                        // (:247-249, :309-311); where it lives is a link of the contig record
                        const uint32_t base = rv.code(dir ? L.sp.end + K - 1 : L.sp.begin);
                        L.set_ctx(dir ? C_RIGHT_J : C_LEFT_J);
                        const Coord next = contig_link(ix, L.sp.anchor, dir, base);
                        if (next.offset >= 0) {
                            L.sp.anchor = next;
                            L.st = P_CONTIG;
                        } else {  // not a hit (or a degenerate slot): the table lookup decides
                            uint64_t tail = L.sp.anchor.offset == 0 ? first_kmer : last_kmer;  // get_tail_kmer
                            if (L.sp.anchor.entry < 0) tail = revcomp(tail);
                            want = true;
                            want_kmer = dir ? ((tail << 2) | base) & KMER_MASK
                                            : (tail >> 2) | ((uint64_t)base << (2 * K - 2));
                        }
                    }
                } else if (shift == INVALID_SHIFT) {
                    L.l.n = 0;
                }
                if (finished) ev = dir ? EV_AFTER_ATTEMPT : EV_AFTER_LEFT;
            }
        } else {  // P_TALLY: the unit is mapped; leave its record for tally_units_kernel
            if (mine) {
                int length;
                if (L.void_unit()) {
                    length = 0;
                    L.l.n = 0;
                } else if (a.paired) {  // map_read_pair (:127-145): span1 = m1, span2 = (sp, l)
                    int begin1 = L.m1_begin, end1;
                    if (!intersect(L.m1, L.l, I.list0, I.list0 + LIST_CAP * ITEMS)) {
                        L.m1.n = 0;
                        begin1 = 0;
                        end1 = -K;
                    } else if (L.m1_anchor.entry != ~L.sp.anchor.entry) {
                        begin1 = 0;
                        end1 = -K;
                    } else {
                        end1 = L.m1_len - K;
                        int interval = L.sp.anchor.offset - L.m1_anchor.offset;
                        if (L.m1_anchor.entry < 0) interval = -interval;
                        end1 += interval + (L.len - K) - L.sp.begin;
                    }
                    length = end1 - begin1 + K;
                    L.l = L.m1;
                } else {
                    length = L.sp.end - L.sp.begin + K;
                }
                int32_t *rec = a.units + L.unit;
                const int n = L.l.n;
                rec[0] = n;
                rec[a.n_units] = length;
                rec += 2 * a.n_units;
                if (n <= REC_IDS && !L.l.in_arena()) {
                    // the final list of a unit is that of mate 1 for a pair, of the read otherwise:
                    // either way the first LIST_CAP rows of the item's lists
                    const int32_t *sp = I.list0;
#pragma unroll 1
                    for (int i = 0; i < n; ++i) rec[i * a.n_units] = sp[i * ITEMS];
                } else if (n <= REC_IDS) {
#pragma unroll 1
                    for (int i = 0; i < n; ++i) rec[i * a.n_units] = L.l.get(i);
                } else {
                    const long long off = emit_long_list(L.l.p, L.l.stride, n, a.arena, a.arena_cap, a.cursors + 1, status);
                    rec[0] = (int32_t)(uint32_t)off;
                    rec[a.n_units] = (int32_t)(off >> 32);
                }
                // the item is free again: mate 0 of a new unit
                L.set_void_unit(false);
                L.set_mate(0);
                L.l = I.fresh_list(0);
                L.m1 = I.fresh_list(0);
                L.st = P_LOAD;
            }
        }

        // ---- transitions (one copy, all phases) ------------------------------------------------
        if (mine) {
            if (ev == EV_LEFT_FAILED) {
                if (L.ctx() == C_LEFT_J) {
                    if (L.sp.begin < K) {
                        L.sp.begin = 0;
                        ev = EV_AFTER_LEFT;
                    } else {
                        L.sp.begin -= K;
                        L.set_ctx(C_LEFT_F);
                        want_pos = L.sp.begin;
                        ev = EV_NONE;
                    }
                } else {  // C_LEFT_F
                    L.l.n = 0;
                    ev = EV_AFTER_LEFT;
                }
            }
            if (ev == EV_AFTER_LEFT) {
                if (L.l.n != 0 && L.sp.end < L.len - K) {
                    if (L.ctg_a0()) {
                        // the stash still holds the contig of the first hit (no junction was
                        // crossed): _filter_targets_to_right starts there (:283-295) without
                        // another record load
                        L.sp.anchor = L.anchor0;
                        L.set_forward(L.anchor0.entry >= 0);
                        L.move = L.forward() ? L.clen - L.anchor0.offset - K : L.anchor0.offset;
                        L.set_dir(1);
                        L.st = P_WALK;
                    } else {
                        L.set_ctx(C_RIGHT_C);
                        L.st = P_CONTIG;
                    }
                    ev = EV_NONE;
                } else {
                    ev = EV_AFTER_ATTEMPT;
                }
            }
            if (ev == EV_AFTER_ATTEMPT) {
                if (L.l.n != 0 || L.attempt() == 1) {
                    ev = EV_READ_DONE;
                } else {
                    L.set_attempt(1);
                    L.sp.anchor = coord_invalid();
                    L.sp.begin += K;
                    if (L.sp.begin + K > L.len) L.sp.begin = L.len - K;
                    L.sp.end = L.sp.begin;
                    L.pos = L.sp.begin;
                    L.l = I.fresh_list(L.mate());
                    L.set_ctx(C_FIND);
                    want_pos = L.pos;
                    ev = EV_NONE;
                }
            }
            if (ev == EV_READ_DONE) {
                if (a.paired && L.mate() == 0) {
                    L.m1_begin = L.sp.begin;
                    L.m1_anchor = L.sp.anchor;
                    L.m1_len = L.len;
                    L.m1 = L.l;
                    L.m1_dirty = true;
                    L.set_mate(1);
                    L.st = P_LOAD;
                } else {
                    L.st = P_TALLY;
                }
            }
            if (want_pos >= 0) {
                want = true;
                want_kmer = rv.kmer(want_pos);
            }
            if (want) {  // the next single k-mer; P_LOOKUP hashes it, with all its lanes together
                L.kmer = want_kmer;
                L.st = P_LOOKUP;
            }
        }
        __syncwarp();
        if (mine) {
            // ---- publish: state first, then the row's bit in the next phase's mask ----------
            if (L.st == P_DEAD) {
                atomicSub(sm_live, 1);
            } else {
                lane_store(L, I);
                __threadfence_block();
                atomicOr(&sm_masks[L.st * 32 + lane], 1u << row);
            }
        }
        __syncwarp();
#if SKM_STATS
        if (lane == 0) {
            atomicAdd(&g_map_stats[phase * 4 + 0], 1ULL);
            atomicAdd(&g_map_stats[phase * 4 + 1], (unsigned long long)__popc(stat_got));
            atomicAdd(&g_map_stats[phase * 4 + 3], (unsigned long long)(clock64() - stat_t0));
        }
#endif
    }
}

// Pass 3: tally.  One thread per unit reads what the mapper left (coalesced, field-major),
// updates the FLD (_mapper.pyx:90-94, shared-memory histogram flushed once per block), finds or
// inserts the ordered id tuple in the class dictionary (MapResult.update, mapper.py:60-75; one
// count atomic per distinct class per warp) and keeps the smallest unit index per class
// (first-seen order).
struct UnitIds {
    const int32_t *p;  // &units[2][u], or the arena list
    int64_t stride;    // n_units, or 1
    __device__ __forceinline__ int32_t get(int i) const { return __ldg(p + i * stride); }
};

__global__ void __launch_bounds__(256)
tally_units_kernel(const DictDev dict, const int32_t *__restrict__ units, const int32_t *__restrict__ arena,
                   int64_t n_units, int64_t first_unit, int32_t *__restrict__ out_class,
                   int32_t *__restrict__ out_length)
{
    __shared__ uint32_t sm_fld[SKM_MAX_FRAGMENT_LENGTH];
    __shared__ unsigned long long sm_totals[4];  // new classes, ids stored, unaligned, aligned
    __shared__ unsigned sm_warp_ids[8];          // ids of the new classes of each warp (this pass)
    __shared__ unsigned long long sm_pool_base;  // where the block's new classes start in the id pool
    for (int i = threadIdx.x; i < SKM_MAX_FRAGMENT_LENGTH; i += blockDim.x) sm_fld[i] = 0;
    if (threadIdx.x < 4) sm_totals[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // every counter that all units share is aggregated: id-pool space per BLOCK pass (exactly the
    // ids of the block's new classes, so the pool holds nothing but stored ids: one atomic on
    // the pool cursor per pass that found a new class), class / id / unit totals per block
    const int warp = threadIdx.x >> 5;
    unsigned new_classes = 0, new_ids = 0, n_unaligned = 0, n_aligned = 0;  // lane 0 only
    for (int64_t block_base = blockIdx.x * (int64_t)blockDim.x; block_base < n_units;
         block_base += (int64_t)gridDim.x * blockDim.x) {  // same trip count for the whole block
        const int64_t u = block_base + threadIdx.x;
        const bool live = u < n_units;
        long long slot = -1;
        bool won = false;
        int n = 0;
        UnitIds view{units + 2 * n_units + u, n_units};
        if (live) {
            n = __ldg(units + u);
            int length = __ldg(units + n_units + u);
            if (out_length) out_length[u] = length;
            if (length > 0) {
                if (length >= SKM_MAX_FRAGMENT_LENGTH) length = SKM_MAX_FRAGMENT_LENGTH - 1;
                atomicAdd(&sm_fld[length], 1u);
            }
            if (n > 0) {
                if (n > REC_IDS) {
                    const long long off = (long long)(uint32_t)view.get(0) | ((long long)view.get(1) << 32);
                    view = UnitIds{arena + off, 1};
                }
                slot = dict_find_or_claim(dict, dict_key(dict, view, n, true), won);
                if (!won && slot >= 0) dict_verify_hit(dict, slot, view, n, true);
            }
            if (out_class) out_class[u] = (int32_t)slot;
            if (slot >= 0) {
                const unsigned long long g = (unsigned long long)(first_unit + u);
                if (g < *reinterpret_cast<volatile unsigned long long *>(&dict.first[slot]))
                    atomicMin(&dict.first[slot], g);
            }
        }
        // ---- new classes: pool space for all winners of the block at once ----------------------
        if (__syncthreads_or(won)) {
            const unsigned winners = __ballot_sync(0xffffffffu, won);
            int incl = won ? n : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned total = (unsigned)__shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) sm_warp_ids[warp] = total;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long all = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) all += sm_warp_ids[w];
                sm_pool_base = atomicAdd(&dict.scalars[0], all);
            }
            __syncthreads();
            if (won) {
                unsigned long long off = sm_pool_base + (unsigned long long)(incl - n);
                for (int w = 0; w < warp; ++w) off += sm_warp_ids[w];
                dict_store_ids(dict, slot, off, view, n, true);
            }
            new_classes += (unsigned)__popc(winners);
            new_ids += total;
            __syncthreads();  // sm_warp_ids / sm_pool_base are rewritten by the next pass
        }
        // ---- one count atomic per distinct class per warp (mapper.py:60-75) -------------------------
        const unsigned same = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0 && lane == __ffs(same) - 1)
            atomicAdd(&dict.counts[slot], (unsigned long long)__popc(same));
        const unsigned done = __ballot_sync(0xffffffffu, live);
        const unsigned mapped = __ballot_sync(0xffffffffu, live && slot >= 0);
        n_aligned += (unsigned)__popc(mapped);
        n_unaligned += (unsigned)(__popc(done) - __popc(mapped));
    }
    if (lane == 0) {
        if (new_classes) atomicAdd(&sm_totals[0], (unsigned long long)new_classes);
        if (new_ids) atomicAdd(&sm_totals[1], (unsigned long long)new_ids);
        if (n_unaligned) atomicAdd(&sm_totals[2], (unsigned long long)n_unaligned);
        if (n_aligned) atomicAdd(&sm_totals[3], (unsigned long long)n_aligned);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SKM_MAX_FRAGMENT_LENGTH; i += blockDim.x) {
        const uint32_t v = sm_fld[i];
        if (v) atomicAdd(&dict.fld[i], (unsigned long long)v);
    }
    if (threadIdx.x == 0) {
        if (sm_totals[0]) atomicAdd(&dict.scalars[1], sm_totals[0]);
        if (sm_totals[1]) atomicAdd(&dict.scalars[4], sm_totals[1]);
        if (sm_totals[2]) atomicAdd(&dict.scalars[2], sm_totals[2]);
        if (sm_totals[3]) atomicAdd(&dict.scalars[3], sm_totals[3]);
    }
}

}  // namespace skm
