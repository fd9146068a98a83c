// Read -> equivalence-class mapping on the GPU.
//
// Two passes per batch: pack_reads_kernel turns the ASCII reads into 2-bit codes plus one
// wildcard bit per base; map_reads_kernel runs the reference's per-read state machine against
// the HBM-resident index with persistent lanes and warp-voted phases (see the kernel comment),
// intersects the mates and tallies the ordered transcript-id tuple into a device-resident
// class dictionary.  The fragment-length histogram is accumulated in shared memory and
// flushed once per block.
//
// Reference semantics restated here (paths under /root/reference/seekmer/):
//   map_read            _mapper.pyx:151-193      find_first_kmer   :199-216
//   filter_to_left      _mapper.pyx:222-275      filter_to_right   :281-343
//   intersect (mates)   _mapper.pyx:350-397      map_read_pair     :111-145
//   sift4_align_left    _mapper.pyx:404-445      sift4_align_right :452-493
//   map_contig          _common.pyx:143-179      filter_on_contig  :185-235
//   get_contig_sequence _common.pyx:103-137      get_tail_kmer     :241-266
//   batch driver + FLD  _mapper.pyx:73-101       tuple ids         :528-537
#include <algorithm>
#include <cstdlib>

#include "kmer.cuh"
#include "sift4.cuh"

namespace skm {

#ifndef SKM_LIST_CAP
#define SKM_LIST_CAP 16
#endif
#ifndef SKM_Q_THREADS
#define SKM_Q_THREADS 512
#endif
constexpr int Q_THREADS = SKM_Q_THREADS;  // worker threads per block (one block per SM)
constexpr int LIST_CAP = SKM_LIST_CAP;  // per-read target list entries kept in shared memory
constexpr int ALIGN_LENGTH = 8;    // _mapper.pyx:22
constexpr int INVALID_SHIFT = SIFT4_INVALID_SHIFT;  // _mapper.pyx:28
constexpr int FLD_BINS = SKM_MAX_FRAGMENT_LENGTH;

struct DictDev {
    ulonglong2 *keys;             // 128-bit tuple hash; all-ones = empty
    unsigned long long *counts;
    unsigned long long *first;    // smallest global unit index that produced the class
    uint32_t *pool_off;
    uint32_t *len;
    int32_t *pool;                // transcript ids of every class, tuple order
    uint64_t mask;                // slots - 1
    uint64_t pool_cap;
    unsigned long long *scalars;  // [0]=pool cursor [1]=n_classes [2]=unaligned [3]=aligned
    unsigned long long *fld;      // FLD_BINS
    uint32_t *status;
};

struct MapArgs {
    const uint64_t *packed;   // [n_reads][words] from pack_reads_kernel
    const int32_t *lens;      // per-read length, or NULL with fixed_len
    int32_t fixed_len;
    int32_t code_words;       // u64 words of 2-bit codes per read (from max read length)
    int32_t words;            // code_words + wildcard words
    int32_t paired;
    int64_t n_units;
    int64_t first_unit;
    int32_t *out_class;
    int32_t *out_length;
    int32_t *arena;           // spill space for target lists longer than LIST_CAP
    uint64_t arena_cap;
    unsigned long long *cursors;  // [0]=work counter [1]=arena cursor
};

// ---- a read packed in shared memory, item-interleaved [word][row][lane] ---------------
struct ReadView {
    const uint64_t *w;  // &reads[0][row][lane]; word k at w[k * stride]
    int len;
    int code_words;
    int stride;         // items in the block's pool (rows * 32)

    __device__ __forceinline__ uint64_t word(int k) const { return w[k * stride]; }
    __device__ __forceinline__ uint32_t code(int p) const
    {
        return (uint32_t)(word(p >> 5) >> (62 - 2 * (p & 31))) & 3u;
    }
    __device__ __forceinline__ bool wild(int p) const
    {
        return (word(code_words + (p >> 6)) >> (p & 63)) & 1ULL;
    }
    // 25-mer starting at base p (_kmer.pxd:46-68)
    __device__ __forceinline__ uint64_t kmer(int p) const
    {
        const int k = p >> 5, s = p & 31;
        uint64_t x = word(k) << (2 * s);
        if (s > 7) x |= word(k + 1) >> (64 - 2 * s);
        return x >> 14;
    }
    // _match_base (_mapper.pyx:500-501): equal, or the read byte is not one of "ACGT"
    __device__ __forceinline__ bool match(uint32_t ref_code, int p) const
    {
        return wild(p) || code(p) == ref_code;
    }
};

// ---- a target list: shared memory ([entry][row][lane]) or arena (dense) ----------------
// Shared-memory lists are addressed with explicit ld/st.shared (a generic pointer with a
// run-time stride would force generic loads); `sa` == 0 selects the global arena.
struct List {
    uint32_t sa;   // shared-window address of element 0, or 0
    uint32_t sb;   // byte stride between elements in shared memory (pool items * 4)
    int32_t *gp;   // arena pointer when sa == 0
    int n;
    __device__ __forceinline__ int32_t get(int i) const
    {
        if (sa) {
            int32_t v;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(sa + sb * (uint32_t)i));
            return v;
        }
        return gp[i];
    }
    __device__ __forceinline__ void set(int i, int32_t v)
    {
        if (sa) asm volatile("st.shared.s32 [%0], %1;" ::"r"(sa + sb * (uint32_t)i), "r"(v) : "memory");
        else gp[i] = v;
    }
};

__device__ __forceinline__ List shared_list(const int32_t *smem_ptr, int pool_items)
{
    return List{(uint32_t)__cvta_generic_to_shared(smem_ptr), 4u * (uint32_t)pool_items, nullptr, 0};
}

struct Span {
    int begin, end;
    Coord anchor;
};

struct Ctx {
    const DevIndex &ix;
    const MapArgs &a;
    uint32_t *status;
    int pool_items;
};

// get_contig_sequence(coordinate, +-8) as a 16-bit window
__device__ __forceinline__ uint32_t contig_window(const DevIndex &ix, const Contig &c, Coord a,
                                                  bool left_edge)
{
    const int64_t p = c.seq_offset + a.offset;
    if (a.entry >= 0) return seq_window8(ix, left_edge ? p : p + K - ALIGN_LENGTH);
    return revcomp8(seq_window8(ix, left_edge ? p + K - ALIGN_LENGTH : p));
}

__device__ __forceinline__ uint64_t tail_kmer(const Contig &c, Coord a)
{
    const uint64_t kmer = a.offset == 0 ? c.first_kmer : c.last_kmer;
    return a.entry < 0 ? revcomp(kmer) : kmer;
}

// map_contig (_common.pyx:143-179) for an already loaded contig record
__device__ void map_contig(const Ctx &cx, const Contig &c, Coord a, List &l, const int32_t *smem_list)
{
    const bool forward = a.entry >= 0;
    const int n = c.target_count;
    l = shared_list(smem_list, cx.pool_items);
    if (n > LIST_CAP) {
        const unsigned long long off = atomicAdd(&cx.a.cursors[1], (unsigned long long)n);
        if (off + (unsigned long long)n > cx.a.arena_cap) {
            atomicOr(cx.status, ST_ARENA_FULL);
            l.n = 0;
            return;
        }
        l.sa = 0;
        l.gp = cx.a.arena + off;
    }
    const int32_t *t = cx.ix.targets + c.target_offset;
    if (forward) {
        for (int i = 0; i < n; ++i) l.set(i, __ldg(t + i));
    } else {
        for (int i = 0; i < n; ++i) l.set(i, ~__ldg(t + (n - 1 - i)));
    }
    l.n = n;
}

// _filter_on_contig (_common.pyx:185-235): sorted-merge intersection, direction aware
__device__ bool filter_on_contig(const Ctx &cx, const Contig &c, Coord a, List &l)
{
    if (l.n == 0) return true;
    const bool forward = a.entry >= 0;
    const int32_t *t = cx.ix.targets + c.target_offset;
    const int length = c.target_count;
    int read_index = 0, write_index = 0, track = 0;
    if (length == 0) return false;
    int32_t index_entry = forward ? __ldg(t) : ~__ldg(t + length - 1);
    int32_t target_entry = l.get(0);
    while (true) {
        if (target_entry == index_entry) {
            l.set(write_index, target_entry);
            read_index += 1;
            write_index += 1;
            track += 1;
            if (read_index == l.n || track == length) break;
            target_entry = l.get(read_index);
            index_entry = forward ? __ldg(t + track) : ~__ldg(t + length - 1 - track);
        } else if (target_entry < index_entry) {
            read_index += 1;
            if (read_index == l.n) break;
            target_entry = l.get(read_index);
        } else {
            track += 1;
            if (track == length) break;
            index_entry = forward ? __ldg(t + track) : ~__ldg(t + length - 1 - track);
        }
    }
    if (write_index == 0) return false;
    l.n = write_index;
    return true;
}

// 8-base window at a contig EDGE, taken from the record's first/last k-mer instead of the
// sequence pool: inside the walk loops the anchor always sits on the first or last k-mer of
// its contig (offset 0 or length-k, _mapper.pyx:229-236,289-295), and first_kmer/last_kmer
// are the encodings of the contig's first/last 25 bases (_index_builder.pyx:565-567).
__device__ __forceinline__ uint32_t edge_window(const Contig &c, Coord a, bool left_edge)
{
    const uint32_t head = (uint32_t)(c.first_kmer >> (2 * K - 16)) & 0xFFFFu;  // first 8 bases
    const uint32_t tail = (uint32_t)c.last_kmer & 0xFFFFu;                      // last 8 bases
    if (a.entry >= 0) return left_edge ? head : tail;
    return revcomp8(left_edge ? tail : head);
}

// mate intersection (_mapper.pyx:350-397): list 1 ascending vs list 2 descending, negated
__device__ bool intersect(List &l1, const List &l2)
{
    if (l1.n == 0) return true;
    if (l2.n == 0) return false;
    int cursor1_read = 0, cursor1_write = 0, cursor2 = l2.n - 1;
    while (cursor1_read != l1.n && cursor2 != -1) {
        const int32_t entry1 = l1.get(cursor1_read);
        const int32_t entry2 = ~l2.get(cursor2);
        if (entry1 == entry2) {
            l1.set(cursor1_write, entry1);
            cursor1_read += 1;
            cursor1_write += 1;
            cursor2 -= 1;
        } else if (entry1 < entry2) {
            cursor1_read += 1;
        } else {
            cursor2 -= 1;
        }
    }
    if (cursor1_write == 0) return false;
    l1.n = cursor1_write;
    return true;
}

// ---- class dictionary ---------------------------------------------------------------
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

// Find-or-insert; returns the slot, or -1 when the table is full.  The winner of the
// 128-bit CAS copies the tuple into the id pool; nobody reads it before the kernel ends.
template <typename Ids>
__device__ int64_t dict_find_or_insert(const DictDev &d, ulonglong2 key, const Ids &ids, int n,
                                       bool strip_sign)
{
    uint64_t s = (key.x ^ (key.y >> 17)) & d.mask;
    const ulonglong2 empty = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
    for (uint64_t probes = 0; probes <= d.mask; ++probes) {
        // 64-bit halves are individually atomic; only a definite foreign h1 skips the CAS
        const uint64_t seen = *reinterpret_cast<volatile const uint64_t *>(&d.keys[s].x);
        if (seen == EMPTY_KEY || seen == key.x) {
            const ulonglong2 old = cas128(d.keys + s, empty, key);
            if (old.x == EMPTY_KEY && old.y == EMPTY_KEY) {
                const unsigned long long off = atomicAdd(&d.scalars[0], (unsigned long long)n);
                if (off + (unsigned long long)n > d.pool_cap) {
                    atomicOr(d.status, ST_POOL_FULL);
                    d.pool_off[s] = 0;
                    d.len[s] = 0;
                } else {
                    for (int i = 0; i < n; ++i) {
                        int32_t e = ids.get(i);
                        if (strip_sign && e < 0) e = ~e;
                        d.pool[off + i] = e;
                    }
                    d.pool_off[s] = (uint32_t)off;
                    d.len[s] = (uint32_t)n;
                }
                atomicAdd(&d.scalars[1], 1ULL);
                return (int64_t)s;
            }
            if (old.x == key.x && old.y == key.y) return (int64_t)s;
        }
        s = (s + 1) & d.mask;
    }
    atomicOr(d.status, ST_DICT_FULL);
    return -1;
}

// ---- the kernels ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t lut_entry(uint32_t b)
{
    // bits 1:0 = 2-bit code (_kmer.pxd:253-273), bit 2 = "not one of ACGT" (_mapper.pyx:501)
    const uint32_t u = b & 0xDFu;
    const uint32_t code = u == 'T' ? 3u : u == 'G' ? 2u : u == 'C' ? 1u : 0u;
    const bool upper = b == 'A' || b == 'C' || b == 'G' || b == 'T';
    return code | (upper ? 0u : 4u);
}

// Pass 1: ASCII reads -> packed reads (2-bit codes, first base in the top bits of each u64,
// followed by one wildcard bit per base).  Fully convergent streaming kernel; the byte ->
// (code, wildcard) table lives in shared memory.
__global__ void __launch_bounds__(256)
pack_reads_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ offsets,
                  int32_t fixed_len, int32_t code_words, int32_t words, int64_t n_reads,
                  uint64_t *__restrict__ packed, int32_t *__restrict__ lens)
{
    __shared__ uint8_t sm_lut[256];
    sm_lut[threadIdx.x] = (uint8_t)lut_entry(threadIdx.x);
    __syncthreads();
    const int64_t read = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (read >= n_reads) return;
    int64_t off;
    int len;
    if (offsets) {
        off = offsets[read];
        len = (int)(offsets[read + 1] - off);
    } else {
        off = read * (int64_t)fixed_len;
        len = fixed_len;
    }
    if (lens) lens[read] = len;
    const int max_len = code_words * 32;
    if (len > max_len) len = max_len;  // the host sizes code_words from the longest read
    const uint8_t *src = bases + off;
    uint64_t *out = packed + read * (int64_t)words;
    uint64_t acc = 0, wacc = 0;
    int cw = 0, ww = code_words;
    int j = 0;
    auto push = [&](uint32_t byte) {
        const uint32_t v = sm_lut[byte];
        acc = (acc << 2) | (v & 3u);
        wacc |= (uint64_t)(v >> 2) << (j & 63);
        if ((j & 31) == 31) {
            out[cw] = acc;
            cw += 1;
            acc = 0;
            if ((j & 63) == 63) {
                out[ww] = wacc;
                ww += 1;
                wacc = 0;
            }
        }
        j += 1;
    };
    while (j < len && (reinterpret_cast<uintptr_t>(src + j) & 7)) push(__ldg(src + j));
    while (j + 8 <= len) {
        const uint64_t w8 = __ldg(reinterpret_cast<const unsigned long long *>(src + j));
#pragma unroll
        for (int b = 0; b < 8; ++b) push((uint32_t)(w8 >> (8 * b)) & 0xFFu);
    }
    while (j < len) push(__ldg(src + j));
    if (len & 31) out[cw] = acc << (2 * (32 - (len & 31)));
    if (len & 63) out[ww] = wacc;
}

// Pass 2: the mapper.  A block is a pool of ITEMS (a unit = read or pair, with its packed
// read, its two target lists and ~100 bytes of state) resident in shared memory, and a set
// of worker warps.  The reference's per-read state machine is cut at its memory accesses into
// PHASES:
//
//   P_LOAD    take the next unit (mate 0) and stage the packed read into shared memory
//   P_LOOKUP  KMerIndex.map_kmer (hash + probe) for the pending k-mer; misses are resolved
//             here (_find_first_kmer keeps rolling, walk fallbacks)
//   P_LIST    contig record + map_contig / _filter_on_contig for the hit
//   P_WALK    one head of the left/right contig-walk loops, or the final edge check: jump to
//             the contig edge, 8-base SIFT4 check (direction-generic), next junction k-mer
//   P_TALLY   map_read_pair mate intersection, FLD, class dictionary
//
// Item (row, lane) is only ever worked on by lane `lane` of some warp, so all of its shared
// memory ([field][row][lane]) is bank-conflict free.  For each phase and lane a 32-bit mask
// says which rows are waiting for that phase.  A warp iteration: every lane reads its five
// masks, the warp votes for the phase most of its lanes can serve, each lane claims one
// waiting row of that phase (atomicAnd), loads the item state, runs the phase, stores the
// state and sets the row's bit in the mask of the phase the item needs next (atomicOr).
// With rows >> phases nearly every lane finds work in the voted phase, so each heavy piece
// of code runs once, fully populated, instead of diverging 32 ways; the address an item will
// need next is prefetched into L2 when it is queued, so by the time some warp claims it the
// line is usually on its way.
enum : int { P_LOAD = 0, P_LOOKUP, P_LIST, P_WALK, P_TALLY, N_PHASES, P_DEAD };
// who asked for the pending lookup / list operation
enum : int {
    C_FIND = 0,  // _find_first_kmer scan (_mapper.pyx:199-216)
    C_LEFT_J,    // left walk junction (:247-251)
    C_LEFT_F,    // left walk fallback (:257-261)
    C_RIGHT_C,   // right walk start: contig of the cached first hit (:283-290)
    C_RIGHT_J    // right walk junction (:309-313)
};

// item state words in shared memory
enum : int {
    S_UNIT = 0,   // unit index within the launch
    S_FLAGS,      // ctx[2:0] dir[3] mate[4] attempt[5] forward[6] l-in-arena[7] m1-in-arena[8]
    S_POSLEN,     // pos[15:0] read length[31:16]
    S_MOVE,
    S_KMER_LO,
    S_KMER_HI,
    S_SLOT,
    S_A0_ENTRY,
    S_A0_OFFSET,
    S_SPAN,       // begin[15:0] end[31:16] (signed)
    S_AN_ENTRY,
    S_AN_OFFSET,
    S_NS,         // l.n[15:0] m1.n[31:16]
    S_LARENA,     // arena offsets of spilled lists
    S_M1ARENA,
    S_M1SPAN,     // m1_begin[15:0] m1_len[31:16]
    S_M1_ENTRY,
    S_M1_OFFSET,
    S_WORDS
};
constexpr int CTG_WORDS = 3;  // contig stash: first_kmer, last_kmer, seq_offset

struct Pool {
    uint64_t *reads;   // [words][rows][32]
    uint64_t *ctg;     // [CTG_WORDS][rows][32]
    int32_t *lists;    // [2 * LIST_CAP][rows][32]
    uint32_t *state;   // [S_WORDS][rows][32]
    uint32_t *masks;   // [N_PHASES][32]
    uint32_t *fld;     // [FLD_BINS]
    int *live;         // items that may still produce work
    int items;         // rows * 32
};

struct Lane {
    int st, ctx, dir;
    long long unit;
    int mate, attempt, pos, move;
    bool forward;
    uint64_t kmer;
    uint32_t slot;  // home slot of `kmer`, prefetched into L2 when the k-mer was produced
    Coord anchor0;
    Span sp;
    List l;
    // mate 1 results while mate 2 is mapped
    int m1_begin, m1_len;
    Coord m1_anchor;
    List m1;
};

struct LaneMem {
    const DevIndex *ix;
    ReadView rv;
    int32_t *list0, *list1;
    uint64_t *ctg;  // stash of the current contig: word k at ctg[k * items]
    int items;
    int paired;
    int32_t *arena;
};

__device__ __forceinline__ int sx16(uint32_t v) { return (int)(int16_t)(uint16_t)v; }

__device__ __forceinline__ void lane_load(Lane &L, LaneMem &M, const Pool &P, int item)
{
    const uint32_t *s = P.state + item;
    const int n = P.items;
    L.unit = (long long)s[S_UNIT * n];
    const uint32_t f = s[S_FLAGS * n];
    L.ctx = (int)(f & 7u);
    L.dir = (int)((f >> 3) & 1u);
    L.mate = (int)((f >> 4) & 1u);
    L.attempt = (int)((f >> 5) & 1u);
    L.forward = (f >> 6) & 1u;
    const uint32_t pl = s[S_POSLEN * n];
    L.pos = (int)(pl & 0xFFFFu);
    M.rv.len = (int)(pl >> 16);
    L.move = (int)s[S_MOVE * n];
    L.kmer = (uint64_t)s[S_KMER_LO * n] | ((uint64_t)s[S_KMER_HI * n] << 32);
    L.slot = s[S_SLOT * n];
    L.anchor0 = Coord{(int32_t)s[S_A0_ENTRY * n], (int32_t)s[S_A0_OFFSET * n]};
    const uint32_t sp = s[S_SPAN * n];
    L.sp.begin = sx16(sp);
    L.sp.end = sx16(sp >> 16);
    L.sp.anchor = Coord{(int32_t)s[S_AN_ENTRY * n], (int32_t)s[S_AN_OFFSET * n]};
    const uint32_t ns = s[S_NS * n];
    L.l = shared_list(L.mate ? M.list1 : M.list0, n);
    L.l.n = (int)(ns & 0xFFFFu);
    if (f & 128u) {
        L.l.sa = 0;
        L.l.gp = M.arena + s[S_LARENA * n];
    }
    L.m1 = shared_list(M.list0, n);
    L.m1.n = (int)(ns >> 16);
    if (f & 256u) {
        L.m1.sa = 0;
        L.m1.gp = M.arena + s[S_M1ARENA * n];
    }
    const uint32_t ms = s[S_M1SPAN * n];
    L.m1_begin = sx16(ms);
    L.m1_len = (int)(ms >> 16);
    L.m1_anchor = Coord{(int32_t)s[S_M1_ENTRY * n], (int32_t)s[S_M1_OFFSET * n]};
}

__device__ __forceinline__ void lane_store(const Lane &L, const LaneMem &M, const Pool &P, int item)
{
    uint32_t *s = P.state + item;
    const int n = P.items;
    s[S_UNIT * n] = (uint32_t)L.unit;
    s[S_FLAGS * n] = (uint32_t)L.ctx | ((uint32_t)L.dir << 3) | ((uint32_t)L.mate << 4)
                     | ((uint32_t)L.attempt << 5) | (L.forward ? 64u : 0u) | (L.l.sa ? 0u : 128u)
                     | (L.m1.sa ? 0u : 256u);
    s[S_POSLEN * n] = ((uint32_t)L.pos & 0xFFFFu) | ((uint32_t)M.rv.len << 16);
    s[S_MOVE * n] = (uint32_t)L.move;
    s[S_KMER_LO * n] = (uint32_t)L.kmer;
    s[S_KMER_HI * n] = (uint32_t)(L.kmer >> 32);
    s[S_SLOT * n] = L.slot;
    s[S_A0_ENTRY * n] = (uint32_t)L.anchor0.entry;
    s[S_A0_OFFSET * n] = (uint32_t)L.anchor0.offset;
    s[S_SPAN * n] = ((uint32_t)L.sp.begin & 0xFFFFu) | ((uint32_t)L.sp.end << 16);
    s[S_AN_ENTRY * n] = (uint32_t)L.sp.anchor.entry;
    s[S_AN_OFFSET * n] = (uint32_t)L.sp.anchor.offset;
    s[S_NS * n] = ((uint32_t)L.l.n & 0xFFFFu) | ((uint32_t)L.m1.n << 16);
    if (!L.l.sa) s[S_LARENA * n] = (uint32_t)(L.l.gp - M.arena);
    if (!L.m1.sa) s[S_M1ARENA * n] = (uint32_t)(L.m1.gp - M.arena);
    s[S_M1SPAN * n] = ((uint32_t)L.m1_begin & 0xFFFFu) | ((uint32_t)L.m1_len << 16);
    s[S_M1_ENTRY * n] = (uint32_t)L.m1_anchor.entry;
    s[S_M1_OFFSET * n] = (uint32_t)L.m1_anchor.offset;
}

// Record the next k-mer to look up and start pulling its home slot towards L2: the probe
// phase that consumes it runs when some warp claims the item again.
__device__ __forceinline__ void want_kmer(Lane &L, const LaneMem &M, uint64_t kmer)
{
    L.kmer = kmer;
    L.slot = home_slot(*M.ix, kmer);
    prefetch_l2(M.ix->table + L.slot);
}

__device__ __forceinline__ void read_done(Lane &L, const LaneMem &M)
{
    if (M.paired && L.mate == 0) {
        L.m1_begin = L.sp.begin;
        L.m1_anchor = L.sp.anchor;
        L.m1_len = M.rv.len;
        L.m1 = L.l;
        L.mate = 1;
        L.st = P_LOAD;
    } else {
        L.st = P_TALLY;
    }
}

__device__ __forceinline__ void after_attempt(Lane &L, const LaneMem &M)  // map_read :177-193
{
    if (L.l.n != 0 || L.attempt == 1) {
        read_done(L, M);
        return;
    }
    L.attempt = 1;
    L.sp.anchor = coord_invalid();
    L.sp.begin += K;
    if (L.sp.begin + K > M.rv.len) L.sp.begin = M.rv.len - K;
    L.sp.end = L.sp.begin;
    L.pos = L.sp.begin;
    want_kmer(L, M, M.rv.kmer(L.pos));
    L.l = shared_list(L.mate ? M.list1 : M.list0, M.items);
    L.ctx = C_FIND;
    L.st = P_LOOKUP;
}

__device__ __forceinline__ void after_left(Lane &L, const LaneMem &M)  // map_read :174-176
{
    if (L.l.n != 0 && L.sp.end < M.rv.len - K) {
        prefetch_l2(M.ix->contigs + (L.anchor0.entry >= 0 ? L.anchor0.entry : ~L.anchor0.entry));
        L.ctx = C_RIGHT_C;
        L.st = P_LIST;
    } else {
        after_attempt(L, M);
    }
}

// _filter_targets_to_left :250-263 when the junction lookup or its filter failed
__device__ __forceinline__ void left_junction_failed(Lane &L, const LaneMem &M)
{
    if (L.ctx == C_LEFT_J) {
        if (L.sp.begin < K) {
            L.sp.begin = 0;
            after_left(L, M);
        } else {
            L.sp.begin -= K;
            want_kmer(L, M, M.rv.kmer(L.sp.begin));
            L.ctx = C_LEFT_F;
            L.st = P_LOOKUP;
        }
    } else {  // C_LEFT_F
        L.l.n = 0;
        after_left(L, M);
    }
}

// 9 read bases starting at base s >= 0, first base in bits 17:16
__device__ __forceinline__ uint32_t extract_codes9(const ReadView &rv, int s)
{
    const int k = s >> 5, sh = s & 31;
    uint64_t x = rv.word(k) << (2 * sh);
    if (sh > 23) x |= rv.word(k + 1) >> (64 - 2 * sh);  // k + 1 <= code_words: in bounds
    return (uint32_t)(x >> 46);
}

// wildcard bits of 9 read bases starting at base s >= 0: bit i <-> base s + i
__device__ __forceinline__ uint32_t extract_wild9(const ReadView &rv, int s, int total_words)
{
    const int k = rv.code_words + (s >> 6), b = s & 63;
    uint64_t x = rv.word(k) >> b;
    if (b > 55 && k + 1 < total_words) x |= rv.word(k + 1) << (64 - b);
    return (uint32_t)x & 0x1FFu;
}

// sift4_align_left(window, read, qoff) for dir == 0, sift4_align_right for dir == 1 (sift4.cuh)
__device__ __forceinline__ int sift4_edge(uint32_t ref16, const ReadView &rv, int qoff, int dir, int total_words)
{
    int s = qoff - (1 - dir);  // the left routine may look one base left of its window (:421)
    const int pad = s < 0 ? 1 : 0;
    s += pad;
    uint32_t codes = extract_codes9(rv, s) >> (2 * pad);
    uint32_t wild = (extract_wild9(rv, s, total_words) << pad) & 0x1FFu;
    if (dir) {
        wild = reverse_bits(wild, 9);
    } else {
        codes = reverse_pairs(codes, 9);
        ref16 = reverse_pairs(ref16, 8);
    }
    return sift4_unified(ref16, codes, wild, 1 - dir, dir ? rv.len - qoff : qoff + 8);
}

__global__ void __launch_bounds__(Q_THREADS, 1)
map_reads_kernel(const DevIndex ix, const DictDev dict, const MapArgs a, const int rows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pool P;
    P.items = rows * 32;
    P.reads = reinterpret_cast<uint64_t *>(smem_raw);
    P.ctg = P.reads + (size_t)a.words * P.items;
    P.lists = reinterpret_cast<int32_t *>(P.ctg + (size_t)CTG_WORDS * P.items);
    P.state = reinterpret_cast<uint32_t *>(P.lists + (size_t)2 * LIST_CAP * P.items);
    P.masks = P.state + (size_t)S_WORDS * P.items;
    P.fld = P.masks + N_PHASES * 32;
    P.live = reinterpret_cast<int *>(P.fld + FLD_BINS);

    for (int i = threadIdx.x; i < FLD_BINS; i += blockDim.x) P.fld[i] = 0;
    for (int i = threadIdx.x; i < N_PHASES * 32; i += blockDim.x)
        P.masks[i] = i < 32 ? (rows == 32 ? 0xFFFFFFFFu : (1u << rows) - 1u) : 0u;  // all in P_LOAD
    for (int i = threadIdx.x; i < P.items; i += blockDim.x) P.state[S_FLAGS * P.items + i] = 0;  // mate 0
    if (threadIdx.x == 0) *P.live = P.items;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LaneMem M;
    M.ix = &ix;
    M.items = P.items;
    M.paired = a.paired;
    M.arena = a.arena;
    M.rv.code_words = a.code_words;
    M.rv.stride = P.items;
    M.rv.len = 0;
    Ctx cx{ix, a, dict.status, P.items};
    volatile uint32_t *vmasks = P.masks;

    unsigned iter = (unsigned)warp * 5u;

    for (;;) {
        // ---- vote: the phase most lanes have a waiting row for ---------------------------
        uint32_t mm = 0;
        int phase;
        {
            uint32_t m[N_PHASES];
            unsigned best = 0;
#pragma unroll
            for (int p = 0; p < N_PHASES; ++p) {
                m[p] = vmasks[p * 32 + lane];
                const unsigned c = (unsigned)__popc(__ballot_sync(0xffffffffu, m[p] != 0));
                const unsigned cand = c ? (c << 3) | (unsigned)p : 0u;
                best = cand > best ? cand : best;
            }
            if (best == 0) {
                int live = 0;
                if (lane == 0) live = *reinterpret_cast<volatile int *>(P.live);
                live = __shfl_sync(0xffffffffu, live, 0);
                if (live == 0) break;
                __nanosleep(100);
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
            const uint32_t old = atomicAnd(&P.masks[phase * 32 + lane], ~(1u << row));
            mine = (old >> row) & 1u;
        }
        iter += 1;
        __threadfence_block();
        const int item = row * 32 + lane;
        Lane L;
        L.st = phase;
        if (mine) {
            M.rv.w = P.reads + item;
            M.ctg = P.ctg + item;
            M.list0 = P.lists + item;
            M.list1 = M.list0 + (size_t)LIST_CAP * P.items;
            lane_load(L, M, P, item);
        }

        if (phase == P_LOAD) {
            // ---- new units for finished items (mate 0) ---------------------------------------
            const bool need = mine && L.mate == 0;
            const unsigned nb = __ballot_sync(0xffffffffu, need);
            if (nb) {
                // one global atomic per warp: a unit is only taken when an item is ready for it
                long long base = 0;
                if (lane == 0) base = (long long)atomicAdd(&a.cursors[0], (unsigned long long)__popc(nb));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (need) {
                    L.unit = base + __popc(nb & ((1u << lane) - 1u));
                    if (L.unit >= a.n_units) L.st = P_DEAD;
                }
            }
            if (mine && L.st == P_LOAD) {
                const long long read_idx = a.paired ? 2 * L.unit + L.mate : L.unit;
                const uint64_t *src = a.packed + read_idx * (long long)a.words;
                uint64_t *dst = P.reads + item;
                for (int k = 0; k < a.words; ++k) dst[k * P.items] = __ldg(src + k);
                int len = a.lens ? __ldg(a.lens + read_idx) : a.fixed_len;
                const int max_len = a.code_words * 32;
                if (len > max_len) len = max_len;
                M.rv.len = len;
                L.sp = Span{0, 0, coord_invalid()};
                L.l = shared_list(L.mate ? M.list1 : M.list0, P.items);
                L.attempt = 0;
                L.pos = 0;
                if (len >= K) {
                    want_kmer(L, M, M.rv.kmer(0));
                    L.ctx = C_FIND;
                    L.st = P_LOOKUP;
                } else {  // undefined in the reference; reported unaligned and flagged
                    atomicOr(dict.status, ST_SHORT_READ);
                    read_done(L, M);
                }
            }
        } else if (phase == P_LOOKUP) {
            if (mine) {
                const Coord hit = map_kmer_at(ix, L.kmer, L.slot);
                L.sp.anchor = hit;
                if (hit.offset >= 0) {
                    prefetch_l2(ix.contigs + (hit.entry >= 0 ? hit.entry : ~hit.entry));
                    L.st = P_LIST;
                } else if (L.ctx == C_FIND) {
                    // _find_first_kmer keeps rolling (:208-212); an exhausted scan leaves the
                    // targets empty and map_read returns (:170-171, :186-187)
                    L.pos += 1;
                    if (L.pos + K <= M.rv.len)
                        want_kmer(L, M, ((L.kmer << 2) | M.rv.code(L.pos + K - 1)) & KMER_MASK);
                    else
                        read_done(L, M);
                } else if (L.ctx == C_RIGHT_J) {
                    L.l.n = 0;  // :312-315
                    after_attempt(L, M);
                } else {
                    left_junction_failed(L, M);
                }
            }
        } else if (phase == P_LIST) {
            if (mine) {
                const Coord at = L.ctx == C_RIGHT_C ? L.anchor0 : L.sp.anchor;
                const Contig c = load_contig(ix, at.entry >= 0 ? at.entry : ~at.entry);
                M.ctg[0] = c.first_kmer;
                M.ctg[P.items] = c.last_kmer;
                M.ctg[2 * P.items] = (uint64_t)c.seq_offset;
                bool ok = true;
                if (L.ctx == C_FIND) {
                    map_contig(cx, c, at, L.l, L.mate ? M.list1 : M.list0);
                    L.sp.begin = L.pos;
                    L.sp.end = L.pos;
                    L.anchor0 = at;
                    ok = L.l.n != 0;
                } else if (L.ctx != C_RIGHT_C) {
                    ok = filter_on_contig(cx, c, at, L.l);
                } else {
                    L.sp.anchor = at;  // :283-284 — same k-mer as the scan hit, lookup cached
                }
                L.forward = at.entry >= 0;
                const int to_start = L.forward ? at.offset : c.length - at.offset - K;
                const int to_end = L.forward ? c.length - at.offset - K : at.offset;
                if (L.ctx == C_FIND) {
                    if (!ok) {
                        read_done(L, M);  // `if is_empty(targets): return span`
                    } else if (L.sp.begin > 0) {
                        L.move = to_start;
                        L.dir = 0;
                        L.st = P_WALK;
                    } else {
                        after_left(L, M);
                    }
                } else if (L.ctx == C_RIGHT_C || L.ctx == C_RIGHT_J) {
                    if (ok) {
                        L.move = to_end;
                        L.dir = 1;
                        L.st = P_WALK;
                    } else {
                        L.l.n = 0;  // :312-315
                        after_attempt(L, M);
                    }
                } else if (ok) {
                    L.move = to_start;
                    L.dir = 0;
                    L.st = P_WALK;
                } else {
                    left_junction_failed(L, M);
                }
            }
        } else if (phase == P_WALK) {
            if (mine) {
                // heads of the loops of _filter_targets_to_left (:234-275) and _to_right (:293-343)
                Contig c;
                c.first_kmer = M.ctg[0];
                c.last_kmer = M.ctg[P.items];
                c.seq_offset = (int64_t)M.ctg[2 * P.items];
                const int dir = L.dir;
                int rem = dir ? M.rv.len - L.sp.end - K : L.sp.begin;  // bases left towards the read end
                const bool in_loop = rem > L.move;
                const int step = in_loop ? L.move : rem;
                const int delta = L.forward ? step : -step;
                L.sp.anchor.offset += dir ? delta : -delta;
                uint32_t ref16;
                int qoff;
                if (in_loop) {
                    rem -= L.move;
                    ref16 = edge_window(c, L.sp.anchor, dir == 0);
                    qoff = dir ? M.rv.len - rem - ALIGN_LENGTH : rem;
                } else {
                    ref16 = contig_window(ix, c, L.sp.anchor, dir == 0);
                    qoff = dir ? M.rv.len - ALIGN_LENGTH : 0;
                }
                const int shift = sift4_edge(ref16, M.rv, qoff, dir, a.words);
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
                        if (dir) L.sp.end = M.rv.len - rem - K;
                        else L.sp.begin = rem;
                    } else if (!dir) {
                        L.sp.begin = rem;  // the failed left walk leaves begin where it stopped (:235,241-242)
                    } else {
                        L.sp.end = M.rv.len - rem - K;
                    }
                    if (!finished) {
                        const uint64_t tail = tail_kmer(c, L.sp.anchor);
                        if (dir) want_kmer(L, M, ((tail << 2) | M.rv.code(L.sp.end + K - 1)) & KMER_MASK);
                        else want_kmer(L, M, (tail >> 2) | ((uint64_t)M.rv.code(L.sp.begin) << (2 * K - 2)));
                        L.ctx = dir ? C_RIGHT_J : C_LEFT_J;
                        L.st = P_LOOKUP;
                    }
                } else if (shift == INVALID_SHIFT) {
                    L.l.n = 0;
                }
                if (finished) {
                    if (dir) after_attempt(L, M);
                    else after_left(L, M);
                }
            }
        } else {  // P_TALLY
            long long slot = -1;
            if (mine) {
                int length;
                if (a.paired) {  // map_read_pair (:127-145): span1 = m1, span2 = (sp, l)
                    int begin1 = L.m1_begin, end1;
                    if (!intersect(L.m1, L.l)) {
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
                        end1 += interval + (M.rv.len - K) - L.sp.begin;
                    }
                    length = end1 - begin1 + K;
                    L.l = L.m1;
                } else {
                    length = L.sp.end - L.sp.begin + K;
                }
                if (a.out_length) a.out_length[L.unit] = length;
                if (length > 0) {  // _mapper.pyx:90-94
                    if (length >= FLD_BINS) length = FLD_BINS - 1;
                    atomicAdd(&P.fld[length], 1u);
                }
                if (L.l.n > 0) {
                    const ulonglong2 key = tuple_key(L.l, L.l.n, true);
                    slot = dict_find_or_insert(dict, key, L.l, L.l.n, true);
                }
                if (a.out_class) a.out_class[L.unit] = (int32_t)slot;
                if (slot >= 0) {
                    const unsigned long long g = (unsigned long long)(a.first_unit + L.unit);
                    if (g < *reinterpret_cast<volatile unsigned long long *>(&dict.first[slot]))
                        atomicMin(&dict.first[slot], g);
                }
                // the item is free again: mate 0 of a new unit
                L.mate = 0;
                L.l = shared_list(M.list0, P.items);
                L.m1 = shared_list(M.list0, P.items);
                L.st = P_LOAD;
            }
            __syncwarp();
            // one count atomic per distinct class per warp (mapper.py:60-75)
            const unsigned same = __match_any_sync(0xffffffffu, slot);
            if (slot >= 0 && lane == __ffs(same) - 1)
                atomicAdd(&dict.counts[slot], (unsigned long long)__popc(same));
            const unsigned done = __ballot_sync(0xffffffffu, mine);
            const unsigned mapped = __ballot_sync(0xffffffffu, mine && slot >= 0);
            if (lane == 0) {
                const int n_al = __popc(mapped);
                const int n_un = __popc(done) - n_al;
                if (n_un) atomicAdd(&dict.scalars[2], (unsigned long long)n_un);
                if (n_al) atomicAdd(&dict.scalars[3], (unsigned long long)n_al);
            }
        }
        // ---- publish: state first, then the row's bit in the next phase's mask --------------
        if (mine) {
            if (L.st == P_DEAD) {
                atomicSub(P.live, 1);
            } else {
                lane_store(L, M, P, item);
                __threadfence_block();
                atomicOr(&P.masks[L.st * 32 + lane], 1u << row);
            }
        }
        __syncwarp();
    }

    __syncthreads();
    for (int i = threadIdx.x; i < FLD_BINS; i += blockDim.x) {
        const uint32_t v = P.fld[i];
        if (v) atomicAdd(&dict.fld[i], (unsigned long long)v);
    }
}

// ---- export / merge -------------------------------------------------------------------
// Compaction of the dictionary to CSR.  One 64-bit atomic hands out the class index
// (high 24 bits) and the id start (low 40 bits) together, so starts are monotone in the
// class index and key_offsets is a proper CSR row pointer.  Class order is arbitrary;
// callers sort by first_unit for the reference's insertion order.
__global__ void dict_export_kernel(const DictDev d, int64_t slots, unsigned long long *cursor,
                                   int64_t *key_offsets, int32_t *key_ids, int64_t *counts,
                                   int64_t *first_unit, int32_t *slot_ids)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= slots) return;
    if (d.keys[s].x == EMPTY_KEY && d.keys[s].y == EMPTY_KEY) return;
    const uint32_t n = d.len[s];
    const unsigned long long old = atomicAdd(cursor, (1ULL << 40) | (unsigned long long)n);
    const unsigned long long c = old >> 40;
    const unsigned long long p = old & ((1ULL << 40) - 1);
    if (key_offsets) key_offsets[c] = (int64_t)p;
    if (key_ids) {
        const int32_t *src = d.pool + d.pool_off[s];
        for (uint32_t i = 0; i < n; ++i) key_ids[p + i] = src[i];
    }
    if (counts) counts[c] = (int64_t)d.counts[s];
    if (first_unit) first_unit[c] = (int64_t)d.first[s];
    if (slot_ids) slot_ids[c] = (int32_t)s;
}

__global__ void set_i64_kernel(int64_t *dst, int64_t v) { *dst = v; }

__global__ void dict_merge_kernel(const DictDev d, const int64_t *key_offsets, const int32_t *key_ids,
                                  const int64_t *counts, const int64_t *first_unit, int64_t n_classes)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_classes) return;
    const int64_t start = key_offsets[c];
    const int n = (int)(key_offsets[c + 1] - start);
    if (n <= 0) return;
    const DenseIds ids{key_ids + start};
    const ulonglong2 key = tuple_key(ids, n, false);
    const int64_t slot = dict_find_or_insert(d, key, ids, n, false);
    if (slot < 0) return;
    atomicAdd(&d.counts[slot], (unsigned long long)counts[c]);
    atomicMin(&d.first[slot], (unsigned long long)first_unit[c]);
    atomicAdd(&d.scalars[3], (unsigned long long)counts[c]);
}

__global__ void add_i64_kernel(unsigned long long *dst, const int64_t *src, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += (unsigned long long)src[i];
}

}  // namespace skm

using namespace skm;

struct skm_mapper {
    skm_index *index = nullptr;
    int device = 0;
    DictDev d{};
    int64_t slots = 0, pool_cap = 0;
    unsigned long long *cursors = nullptr;  // [0]=work [1]=arena [2..3]=export cursors
    int32_t *arena = nullptr;
    uint64_t arena_cap = 0;
    int sm_count = 148;
    // staging for host-buffer calls
    uint8_t *d_bases[2] = {nullptr, nullptr};  // double-buffered H2D staging
    size_t d_bases_cap[2] = {0, 0};
    int64_t *d_offsets[2] = {nullptr, nullptr};
    size_t d_offsets_cap[2] = {0, 0};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr};
    size_t smem_configured = 0, smem_max = 0;
    int threads = Q_THREADS, rows_limit = 0;  // SKM_THREADS / SKM_ROWS override for experiments
    int32_t *d_out = nullptr;
    size_t d_out_cap = 0;
    uint64_t *d_packed = nullptr;  // pack_reads_kernel output
    size_t d_packed_cap = 0;
    int32_t *d_lens = nullptr;
    size_t d_lens_cap = 0;
};

static int ensure(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4;
    SKM_CUDA(cudaMalloc(p, want));
    *cap = want;
    return 0;
}

SKM_API void skm_mapper_destroy(skm_mapper *m)
{
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(m->device);
    cudaFree(m->d.keys);
    cudaFree(m->d.counts);
    cudaFree(m->d.first);
    cudaFree(m->d.pool_off);
    cudaFree(m->d.len);
    cudaFree(m->d.pool);
    cudaFree(m->d.scalars);
    cudaFree(m->d.fld);
    cudaFree(m->d.status);
    cudaFree(m->cursors);
    cudaFree(m->arena);
    for (int i = 0; i < 2; ++i) {
        cudaFree(m->d_bases[i]);
        cudaFree(m->d_offsets[i]);
        if (m->ev_copy[i]) cudaEventDestroy(m->ev_copy[i]);
        if (m->ev_compute[i]) cudaEventDestroy(m->ev_compute[i]);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    cudaFree(m->d_out);
    cudaFree(m->d_packed);
    cudaFree(m->d_lens);
    cudaSetDevice(prev);
    delete m;
}

SKM_API int skm_mapper_reset(skm_mapper *m, void *stream)
{
    if (!m) return fail(SKM_ERR_INVALID, "skm_mapper_reset: NULL mapper");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA(cudaMemsetAsync(m->d.keys, 0xFF, sizeof(ulonglong2) * (size_t)m->slots, st));
    SKM_CUDA(cudaMemsetAsync(m->d.counts, 0, sizeof(unsigned long long) * (size_t)m->slots, st));
    SKM_CUDA(cudaMemsetAsync(m->d.first, 0xFF, sizeof(unsigned long long) * (size_t)m->slots, st));
    SKM_CUDA(cudaMemsetAsync(m->d.len, 0, sizeof(uint32_t) * (size_t)m->slots, st));
    SKM_CUDA(cudaMemsetAsync(m->d.pool_off, 0, sizeof(uint32_t) * (size_t)m->slots, st));
    SKM_CUDA(cudaMemsetAsync(m->d.scalars, 0, sizeof(unsigned long long) * 4, st));
    SKM_CUDA(cudaMemsetAsync(m->d.fld, 0, sizeof(unsigned long long) * FLD_BINS, st));
    SKM_CUDA(cudaMemsetAsync(m->d.status, 0, sizeof(uint32_t), st));
    return SKM_OK;
}

SKM_API int skm_mapper_create(skm_index *index, int64_t class_capacity, int64_t id_capacity,
                                 skm_mapper **out)
{
    if (!out) return fail(SKM_ERR_INVALID, "skm_mapper_create: out is NULL");
    *out = nullptr;
    if (!index) return fail(SKM_ERR_INVALID, "skm_mapper_create: NULL index");
    SKM_CUDA(cudaSetDevice(index->device));
    // Random 16/32-byte probes dominate the traffic: ask L2 to fetch single 32-byte sectors from
    // DRAM instead of the default 64 bytes (a hint; SKM_L2_FETCH overrides for experiments).
    {
        const char *g = getenv("SKM_L2_FETCH");
        const size_t gran = g ? (size_t)atoi(g) : 32;
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    if (class_capacity <= 0) class_capacity = 1LL << 22;
    int64_t slots = 1024;
    while (slots < 2 * class_capacity) slots <<= 1;
    if (id_capacity <= 0) id_capacity = 8 * class_capacity;
    if (id_capacity >= (1LL << 32)) id_capacity = (1LL << 32) - 1;
    skm_mapper *m = new skm_mapper();
    m->index = index;
    m->device = index->device;
    m->slots = slots;
    m->pool_cap = id_capacity;
    cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device);
    {
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
        m->smem_max = (size_t)optin;
        if (const char *t = getenv("SKM_THREADS")) m->threads = std::max(32, std::min(Q_THREADS, atoi(t) & ~31));
        if (const char *r = getenv("SKM_ROWS")) m->rows_limit = atoi(r);
    }
    // list arena: room for every resident thread to spill its largest possible list a few
    // times over, bounded to 2 GiB
    uint64_t arena = (uint64_t)std::max<int64_t>(index->max_target_count, 64) * 2048ULL * (uint64_t)m->sm_count;
    arena = std::min<uint64_t>(std::max<uint64_t>(arena, 1ULL << 22), 1ULL << 29);
    m->arena_cap = arena;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    };
    A((void **)&m->d.keys, sizeof(ulonglong2) * (size_t)slots);
    A((void **)&m->d.counts, sizeof(unsigned long long) * (size_t)slots);
    A((void **)&m->d.first, sizeof(unsigned long long) * (size_t)slots);
    A((void **)&m->d.pool_off, sizeof(uint32_t) * (size_t)slots);
    A((void **)&m->d.len, sizeof(uint32_t) * (size_t)slots);
    A((void **)&m->d.pool, sizeof(int32_t) * (size_t)id_capacity);
    A((void **)&m->d.scalars, sizeof(unsigned long long) * 4);
    A((void **)&m->d.fld, sizeof(unsigned long long) * FLD_BINS);
    A((void **)&m->d.status, sizeof(uint32_t));
    A((void **)&m->cursors, sizeof(unsigned long long) * 4);
    A((void **)&m->arena, sizeof(int32_t) * (size_t)arena);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&m->ev_copy[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_compute[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        skm_mapper_destroy(m);
        return fail(SKM_ERR_OOM, std::string("skm_mapper_create: ") + cudaGetErrorString(e));
    }
    m->d.mask = (uint64_t)slots - 1;
    m->d.pool_cap = (uint64_t)id_capacity;
    int rc = skm_mapper_reset(m, nullptr);
    if (rc == 0 && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = fail(SKM_ERR_CUDA, "reset failed");
    if (rc != 0) {
        skm_mapper_destroy(m);
        return rc;
    }
    *out = m;
    return SKM_OK;
}

static size_t map_item_bytes(int words)
{
    return sizeof(uint64_t) * ((size_t)words + CTG_WORDS) + sizeof(int32_t) * 2 * LIST_CAP + sizeof(uint32_t) * S_WORDS;
}

static size_t map_smem_bytes(int words, int rows)
{
    return map_item_bytes(words) * 32 * (size_t)rows + sizeof(uint32_t) * (N_PHASES * 32 + FLD_BINS) + 16;
}

static int check_status(skm_mapper *m, cudaStream_t st, const char *who)
{
    uint32_t status = 0;
    SKM_CUDA(cudaMemcpyAsync(&status, m->d.status, sizeof(status), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaStreamSynchronize(st));
    if (status & (ST_ARENA_FULL | ST_DICT_FULL | ST_POOL_FULL)) {
        std::string msg = std::string(who) + ": device capacity exhausted:";
        if (status & ST_ARENA_FULL) msg += " target-list arena";
        if (status & ST_DICT_FULL) msg += " class table (raise class_capacity)";
        if (status & ST_POOL_FULL) msg += " class id pool (raise id_capacity)";
        return fail(SKM_ERR_CAPACITY, msg);
    }
    return SKM_OK;
}

// pack + map of one device-resident chunk on stream `st`
static int launch_chunk(skm_mapper *m, const uint8_t *d_bases, const int64_t *d_offsets, MapArgs a,
                        int64_t n_units, int64_t first_unit, int32_t *d_out_class, int32_t *d_out_length,
                        cudaStream_t st)
{
    const int64_t n_reads = a.paired ? 2 * n_units : n_units;
    int rc = ensure((void **)&m->d_packed, &m->d_packed_cap, sizeof(uint64_t) * (size_t)n_reads * a.words);
    if (rc) return rc;
    int32_t *lens = nullptr;
    if (d_offsets) {
        rc = ensure((void **)&m->d_lens, &m->d_lens_cap, sizeof(int32_t) * (size_t)n_reads);
        if (rc) return rc;
        lens = m->d_lens;
    }
    pack_reads_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(d_bases, d_offsets, a.fixed_len, a.code_words,
                                                                        a.words, n_reads, m->d_packed, lens);
    SKM_CUDA(cudaGetLastError());
    a.packed = m->d_packed;
    a.lens = lens;
    a.n_units = n_units;
    a.first_unit = first_unit;
    a.out_class = d_out_class;
    a.out_length = d_out_length;
    SKM_CUDA(cudaMemsetAsync(m->cursors, 0, sizeof(unsigned long long) * 2, st));
    // one block per SM; as many item rows as shared memory holds (at most 32: one mask bit each)
    int rows = (int)std::min<size_t>(32, (m->smem_max - map_smem_bytes(a.words, 0)) / (map_item_bytes(a.words) * 32));
    if (m->rows_limit > 0) rows = std::min(rows, m->rows_limit);
    if (rows < 1) return fail(SKM_ERR_INVALID, "skm_map_batch: reads too long for shared-memory staging");
    const size_t smem = map_smem_bytes(a.words, rows);
    if (smem > m->smem_configured) {
        SKM_CUDA(cudaFuncSetAttribute(map_reads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        m->smem_configured = smem;
    }
    const int64_t want = (n_units + rows * 32 - 1) / (rows * 32);
    const int grid = (int)std::min<int64_t>(m->sm_count, std::max<int64_t>(want, 1));
    map_reads_kernel<<<grid, m->threads, smem, st>>>(m->index->d, m->d, a, rows);
    SKM_CUDA(cudaGetLastError());
    return SKM_OK;
}

SKM_API int skm_map_batch(skm_mapper *m, const uint8_t *bases, const int64_t *read_offsets,
                          int32_t fixed_read_len, int32_t max_read_len, int64_t n_units, int paired,
                          int64_t first_unit, int buffers_on_device, int32_t *out_class, int32_t *out_length,
                          void *stream)
{
    if (!m || !bases) return fail(SKM_ERR_INVALID, "skm_map_batch: NULL argument");
    if (n_units < 0) return fail(SKM_ERR_INVALID, "skm_map_batch: negative unit count");
    if (n_units == 0) return SKM_OK;
    if (!read_offsets && fixed_read_len < K)
        return fail(SKM_ERR_INVALID, "skm_map_batch: need read_offsets or fixed_read_len >= 25");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int per_unit = paired ? 2 : 1;
    const int64_t n_reads = per_unit * n_units;

    if (!read_offsets) max_read_len = fixed_read_len;
    if (!buffers_on_device && read_offsets) {
        int64_t mx = 0, mn = INT64_MAX;
        for (int64_t i = 0; i < n_reads; ++i) {
            const int64_t L = read_offsets[i + 1] - read_offsets[i];
            mx = std::max(mx, L);
            mn = std::min(mn, L);
        }
        if (mn < K) return fail(SKM_ERR_INVALID, "skm_map_batch: read shorter than k=25 (undefined in the reference)");
        max_read_len = (int32_t)mx;
    }
    if (max_read_len < K) return fail(SKM_ERR_INVALID, "skm_map_batch: max_read_len must be >= 25");
    if (max_read_len > 4096) return fail(SKM_ERR_INVALID, "skm_map_batch: reads longer than 4096 bases are not supported");

    MapArgs a{};
    a.fixed_len = read_offsets ? 0 : fixed_read_len;
    a.code_words = (max_read_len + 31) / 32;
    a.words = a.code_words + (max_read_len + 63) / 64;
    a.paired = paired ? 1 : 0;
    a.arena = m->arena;
    a.arena_cap = m->arena_cap;
    a.cursors = m->cursors;

    if (buffers_on_device)
        return launch_chunk(m, bases, read_offsets, a, n_units, first_unit, out_class, out_length, st);

    // ---- host buffers: double-buffered H2D copies overlapped with pack + map --------------
    int32_t *d_class = nullptr, *d_length = nullptr;
    if (out_class || out_length) {
        int rc = ensure((void **)&m->d_out, &m->d_out_cap, sizeof(int32_t) * 2 * (size_t)n_units);
        if (rc) return rc;
        d_class = out_class ? m->d_out : nullptr;
        d_length = out_length ? m->d_out + n_units : nullptr;
    }
    const int64_t avg_unit_bytes = std::max<int64_t>(
        1, (read_offsets ? read_offsets[n_reads] - read_offsets[0] : n_reads * (int64_t)fixed_read_len) / n_units);
    const int64_t chunk_units = std::max<int64_t>(65536, std::min<int64_t>(n_units, (384LL << 20) / avg_unit_bytes));
    // the caller's stream must have finished with previous work on the staging buffers
    SKM_CUDA(cudaEventRecord(m->ev_compute[0], st));
    SKM_CUDA(cudaEventRecord(m->ev_compute[1], st));
    int slot = 0;
    for (int64_t u0 = 0; u0 < n_units; u0 += chunk_units, slot ^= 1) {
        const int64_t nu = std::min(chunk_units, n_units - u0);
        const int64_t r0 = u0 * per_unit, nr = nu * per_unit;
        const int64_t b0 = read_offsets ? read_offsets[r0] : r0 * (int64_t)fixed_read_len;
        const int64_t b1 = read_offsets ? read_offsets[r0 + nr] : (r0 + nr) * (int64_t)fixed_read_len;
        int rc = ensure((void **)&m->d_bases[slot], &m->d_bases_cap[slot], (size_t)(b1 - b0) + 16);
        if (rc) return rc;
        if (read_offsets) {
            rc = ensure((void **)&m->d_offsets[slot], &m->d_offsets_cap[slot], sizeof(int64_t) * (size_t)(nr + 1));
            if (rc) return rc;
        }
        // copy stream: wait until the kernels that last read this slot are done, then copy
        SKM_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_compute[slot], 0));
        SKM_CUDA(cudaMemcpyAsync(m->d_bases[slot], bases + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, m->copy_stream));
        if (read_offsets)
            SKM_CUDA(cudaMemcpyAsync(m->d_offsets[slot], read_offsets + r0, sizeof(int64_t) * (size_t)(nr + 1),
                                     cudaMemcpyHostToDevice, m->copy_stream));
        SKM_CUDA(cudaEventRecord(m->ev_copy[slot], m->copy_stream));
        // compute stream: wait for the copy, pack + map
        SKM_CUDA(cudaStreamWaitEvent(st, m->ev_copy[slot], 0));
        // the kernels index reads from 0 within the chunk; explicit offsets keep their global
        // byte values, so only then is the base pointer shifted back by b0
        rc = launch_chunk(m, read_offsets ? m->d_bases[slot] - b0 : m->d_bases[slot],
                          read_offsets ? m->d_offsets[slot] : nullptr, a, nu,
                          first_unit + u0, d_class ? d_class + u0 : nullptr, d_length ? d_length + u0 : nullptr, st);
        if (rc) return rc;
        SKM_CUDA(cudaEventRecord(m->ev_compute[slot], st));
    }
    if (out_class)
        SKM_CUDA(cudaMemcpyAsync(out_class, d_class, sizeof(int32_t) * (size_t)n_units, cudaMemcpyDeviceToHost, st));
    if (out_length)
        SKM_CUDA(cudaMemcpyAsync(out_length, d_length, sizeof(int32_t) * (size_t)n_units, cudaMemcpyDeviceToHost, st));
    return check_status(m, st, "skm_map_batch");
}

SKM_API int skm_classes_size(skm_mapper *m, int64_t sizes[6], void *stream)
{
    if (!m || !sizes) return fail(SKM_ERR_INVALID, "skm_classes_size: NULL argument");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long sc[4];
    uint32_t status = 0;
    SKM_CUDA(cudaMemcpyAsync(sc, m->d.scalars, sizeof(sc), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaMemcpyAsync(&status, m->d.status, sizeof(status), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaStreamSynchronize(st));
    sizes[0] = (int64_t)sc[1];
    sizes[1] = (int64_t)sc[0];
    sizes[2] = (int64_t)sc[2];
    sizes[3] = (int64_t)sc[3];
    sizes[4] = m->slots / 2;
    sizes[5] = (int64_t)status;
    return SKM_OK;
}

SKM_API int skm_classes_export(skm_mapper *m, int64_t *key_offsets, int32_t *key_ids,
                                  int64_t *counts, int64_t *first_unit, int32_t *slots,
                                  int64_t *fld, int buffers_on_device, void *stream)
{
    if (!m) return fail(SKM_ERR_INVALID, "skm_classes_export: NULL mapper");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_status(m, st, "skm_classes_export");
    if (rc) return rc;
    int64_t sizes[6];
    rc = skm_classes_size(m, sizes, stream);
    if (rc) return rc;
    const int64_t n_cls = sizes[0], n_ids = sizes[1];
    const bool host = !buffers_on_device;

    int64_t *d_off = key_offsets, *d_counts = counts, *d_first = first_unit;
    int32_t *d_ids = key_ids, *d_slots = slots;
    void *owned[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    if (host) {
        auto A = [&](int k, void **p, bool wanted, size_t bytes) {
            *p = nullptr;
            if (!wanted || e != cudaSuccess) return;
            e = cudaMalloc(&owned[k], std::max<size_t>(bytes, 16));
            *p = owned[k];
        };
        A(0, (void **)&d_off, key_offsets != nullptr, sizeof(int64_t) * (size_t)(n_cls + 1));
        A(1, (void **)&d_ids, key_ids != nullptr, sizeof(int32_t) * (size_t)n_ids);
        A(2, (void **)&d_counts, counts != nullptr, sizeof(int64_t) * (size_t)n_cls);
        A(3, (void **)&d_first, first_unit != nullptr, sizeof(int64_t) * (size_t)n_cls);
        A(4, (void **)&d_slots, slots != nullptr, sizeof(int32_t) * (size_t)n_cls);
    }
    auto cleanup = [&]() {
        for (void *p : owned) cudaFree(p);
    };
    if (e != cudaSuccess) {
        cleanup();
        return fail(SKM_ERR_OOM, std::string("skm_classes_export: ") + cudaGetErrorString(e));
    }
    cudaMemsetAsync(m->cursors + 2, 0, sizeof(unsigned long long), st);
    if (n_cls > 0)
        dict_export_kernel<<<(unsigned)((m->slots + 255) / 256), 256, 0, st>>>(
            m->d, m->slots, m->cursors + 2, d_off, d_ids, d_counts, d_first, d_slots);
    if (d_off) set_i64_kernel<<<1, 1, 0, st>>>(d_off + n_cls, n_ids);
    e = cudaGetLastError();
    if (e == cudaSuccess && host) {
        if (key_offsets) cudaMemcpyAsync(key_offsets, d_off, sizeof(int64_t) * (size_t)(n_cls + 1), cudaMemcpyDeviceToHost, st);
        if (key_ids && n_ids > 0) cudaMemcpyAsync(key_ids, d_ids, sizeof(int32_t) * (size_t)n_ids, cudaMemcpyDeviceToHost, st);
        if (counts && n_cls > 0) cudaMemcpyAsync(counts, d_counts, sizeof(int64_t) * (size_t)n_cls, cudaMemcpyDeviceToHost, st);
        if (first_unit && n_cls > 0) cudaMemcpyAsync(first_unit, d_first, sizeof(int64_t) * (size_t)n_cls, cudaMemcpyDeviceToHost, st);
        if (slots && n_cls > 0) cudaMemcpyAsync(slots, d_slots, sizeof(int32_t) * (size_t)n_cls, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess && fld)
        cudaMemcpyAsync(fld, m->d.fld, sizeof(int64_t) * FLD_BINS,
                        host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) return fail(SKM_ERR_CUDA, std::string("skm_classes_export: ") + cudaGetErrorString(e));
    return SKM_OK;
}

SKM_API int skm_classes_merge(skm_mapper *m, const int64_t *key_offsets, const int32_t *key_ids,
                                 const int64_t *counts, const int64_t *first_unit, int64_t n_classes,
                                 const int64_t *fld, int64_t unaligned, int buffers_on_device,
                                 void *stream)
{
    if (!m) return fail(SKM_ERR_INVALID, "skm_classes_merge: NULL mapper");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (n_classes > 0 && (!key_offsets || !key_ids || !counts || !first_unit))
        return fail(SKM_ERR_INVALID, "skm_classes_merge: NULL class arrays");
    int64_t *d_off = nullptr, *d_cnt = nullptr, *d_first = nullptr, *d_fld = nullptr;
    int32_t *d_ids = nullptr;
    const int64_t *p_off = key_offsets, *p_cnt = counts, *p_first = first_unit, *p_fld = fld;
    const int32_t *p_ids = key_ids;
    cudaError_t e = cudaSuccess;
    if (!buffers_on_device) {
        const int64_t n_ids = n_classes > 0 ? key_offsets[n_classes] : 0;
        auto up = [&](void **d, const void *h, size_t bytes) {
            if (e != cudaSuccess || bytes == 0) return;
            e = cudaMalloc(d, bytes);
            if (e == cudaSuccess) e = cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st);
        };
        if (n_classes > 0) {
            up((void **)&d_off, key_offsets, sizeof(int64_t) * (size_t)(n_classes + 1));
            up((void **)&d_ids, key_ids, sizeof(int32_t) * (size_t)n_ids);
            up((void **)&d_cnt, counts, sizeof(int64_t) * (size_t)n_classes);
            up((void **)&d_first, first_unit, sizeof(int64_t) * (size_t)n_classes);
        }
        if (fld) up((void **)&d_fld, fld, sizeof(int64_t) * FLD_BINS);
        p_off = d_off;
        p_ids = d_ids;
        p_cnt = d_cnt;
        p_first = d_first;
        p_fld = d_fld;
    }
    if (e == cudaSuccess && n_classes > 0) {
        dict_merge_kernel<<<(unsigned)((n_classes + 127) / 128), 128, 0, st>>>(m->d, p_off, p_ids, p_cnt,
                                                                             p_first, n_classes);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && fld) {
        add_i64_kernel<<<(FLD_BINS + 255) / 256, 256, 0, st>>>(m->d.fld, p_fld, FLD_BINS);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && unaligned > 0) {
        int64_t *d_un = nullptr;
        e = cudaMalloc((void **)&d_un, sizeof(int64_t));
        if (e == cudaSuccess) {
            cudaMemcpyAsync(d_un, &unaligned, sizeof(int64_t), cudaMemcpyHostToDevice, st);
            add_i64_kernel<<<1, 32, 0, st>>>(m->d.scalars + 2, d_un, 1);
            cudaStreamSynchronize(st);
            cudaFree(d_un);
        }
    }
    int rc = SKM_OK;
    if (e == cudaSuccess) rc = check_status(m, st, "skm_classes_merge");
    cudaFree(d_off);
    cudaFree(d_ids);
    cudaFree(d_cnt);
    cudaFree(d_first);
    cudaFree(d_fld);
    if (e != cudaSuccess) return fail(SKM_ERR_CUDA, std::string("skm_classes_merge: ") + cudaGetErrorString(e));
    return rc;
}
