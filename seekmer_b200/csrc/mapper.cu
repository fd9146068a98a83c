// Read -> equivalence-class mapping on the GPU: host side of the C ABI (skm_mapper_*,
// skm_map_batch, skm_classes_*) plus the streaming kernels around the mapper proper.
//
// Two passes per batch: pack_reads_kernel turns the ASCII reads into 2-bit codes plus one
// wildcard bit per base; map_reads_kernel (map_kernel.cuh) runs the reference's per-read state
// machine against the HBM-resident index, intersects the mates and tallies the ordered
// transcript-id tuple into a device-resident class dictionary (dict.cuh).  The fragment-length
// histogram is accumulated in shared memory and flushed once per block.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include <cub/cub.cuh>

#include "em_plan.cuh"
#include "map_kernel.cuh"

namespace skm {

constexpr int FLD_BINS = SKM_MAX_FRAGMENT_LENGTH;

// ---- pass 1: ASCII reads -> packed reads ----------------------------------------------------
// Four ASCII bases in one 32-bit word -> their 2-bit codes (first base in bits 7:6) and their
// wildcard bits (bit k <-> base k).  Codes follow _kmer.pxd:253-273 (A=0 C=1 G=2 T=3, case
// folded, every other byte 0); a wildcard is any byte that is not one of upper-case "ACGT"
// (_mapper.pyx:500-501).  Branch-free, 4 bytes at a time.
__device__ __forceinline__ void convert4(uint32_t x, uint32_t &codes8, uint32_t &wild4)
{
    // The low 3 bits of the case-folded byte tell the four letters apart (A 1, C 3, G 7, T 4); two
    // byte permutes use them as indexes into 8-entry tables: the 2-bit code, and what bits 7..3 of
    // the byte must be for it to really be that letter (0x41,0x43,0x47 >> 3 = 8, 0x54 >> 3 = 10).
    const uint32_t u = x & 0xDFDFDFDFu;
    const uint32_t lo3 = u & 0x07070707u;
    const uint32_t t = lo3 | (lo3 >> 4);
    const uint32_t sel = __byte_perm(t, 0u, 0x4420u);             // nibble j = low 3 bits of byte j
    const uint32_t code = __byte_perm(0x01000000u, 0x02000003u, sel);  // [1]=A 0 [3]=C 1 [7]=G 2 [4]=T 3
    const uint32_t want = __byte_perm(0x08FF08FFu, 0x08FFFF0Au, sel);  // bits 7..3 of that letter, FF = none
    const uint32_t z = ((u >> 3) & 0x1F1F1F1Fu) ^ want;                   // zero byte <=> the byte is the letter
    const uint32_t letter = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u;  // 0x80 per letter
    const uint32_t codes = code & ((letter >> 7) | (letter >> 6));
    codes8 = (codes * 0x40100401u) >> 24;       // byte j -> bits 7-2j:6-2j, no carries (disjoint fields)
    const uint32_t wild = ~(letter & ~(x << 2)) & 0x80808080u;  // not an upper-case letter of the four
    wild4 = ((wild >> 7) * 0x10204080u) >> 28;  // byte j -> bit j
}

// One thread per 32 bases = one 64-bit code word and half a wildcard word of the packed record
// [code_words | wild_words | pad to a multiple of four words].  Consecutive threads convert consecutive 32-byte
// pieces of the input, so loads and stores are coalesced whatever the read length; the piece is
// fetched with three aligned 16-byte loads and shifted into place.
__global__ void __launch_bounds__(256)
pack_reads_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ starts,
                  const int64_t *__restrict__ ends, int32_t fixed_len, int32_t code_words, int32_t wild_words,
                  int32_t words, int64_t n_reads, uint64_t *__restrict__ packed, int32_t *__restrict__ lens)
{
    const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const int64_t read = idx <= 0xFFFFFFFFULL ? (int64_t)((uint32_t)idx / (uint32_t)code_words)
                                              : (int64_t)(idx / (uint32_t)code_words);
    const int w = (int)(idx - (uint64_t)read * (uint32_t)code_words);
    if (read >= n_reads) return;
    int64_t off;
    int len;
    if (starts) {  // read i = bases[starts[i], ends[i]); a CSR offsets array is starts = o, ends = o + 1
        off = __ldg(starts + read);
        len = (int)(__ldg(ends + read) - off);
    } else {
        off = read * (int64_t)fixed_len;
        len = fixed_len;
    }
    uint64_t *out = packed + read * (int64_t)words;
    if (w == 0) {
        if (lens) lens[read] = len;
        for (int k = code_words + wild_words; k < words; ++k) out[k] = 0;  // padding words
    }
    const int max_len = code_words * 32;
    if (len > max_len) len = max_len;  // the host sizes code_words from the longest read
    const int nb = min(32, len - 32 * w);  // bases of this piece
    uint64_t codes = 0;
    uint32_t wild = 0;
    if (nb > 0) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(bases) + (uintptr_t)off + 32u * (uintptr_t)w;
        const uint4 *q = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
        const int sh = (int)(addr & 15);
        // three aligned chunks cover the 32 bytes from any alignment; a chunk is read only if it
        // holds bytes of this piece (so nothing outside the caller's allocation granule is touched)
        uint32_t x[13];
        const uint4 q0 = __ldg(q);
        const uint4 q1 = sh + nb > 16 ? __ldg(q + 1) : make_uint4(0, 0, 0, 0);
        const uint4 q2 = sh + nb > 32 ? __ldg(q + 2) : make_uint4(0, 0, 0, 0);
        x[0] = q0.x; x[1] = q0.y; x[2] = q0.z; x[3] = q0.w;
        x[4] = q1.x; x[5] = q1.y; x[6] = q1.z; x[7] = q1.w;
        x[8] = q2.x; x[9] = q2.y; x[10] = q2.z; x[11] = q2.w;
        x[12] = 0;
        if (sh & 8) {
#pragma unroll
            for (int i = 0; i < 11; ++i) x[i] = x[i + 2];
        }
        if (sh & 4) {
#pragma unroll
            for (int i = 0; i < 10; ++i) x[i] = x[i + 1];
        }
        const int bs = 8 * (sh & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t c8, w4;
            convert4(__funnelshift_r(x[i], x[i + 1], bs), c8, w4);
            codes |= (uint64_t)c8 << (56 - 8 * i);
            wild |= w4 << (4 * i);
        }
        if (nb < 32) {  // bases past the end of the read are zero codes, no wildcards
            codes &= ~0ULL << (64 - 2 * nb);
            wild &= (1u << nb) - 1u;
        }
    }
    out[w] = codes;
    uint32_t *wout = reinterpret_cast<uint32_t *>(out + code_words) + w;
    *wout = wild;
    if (w == code_words - 1 && !(w & 1)) wout[1] = 0;  // upper half of the last wildcard word
}

// ---- FASTQ text on the device (replaces the line loop of common.feed_*_reads) -----------------
// A chunk of FASTQ text that starts at a record boundary is parsed where it lies: newline
// positions are compacted (count per 4 KB block, scan, write), line 4r+1 of the chunk is the
// sequence of record r, and the pack kernel reads the bases straight out of the text.
constexpr int NL_BLOCK_BYTES = 4096;
constexpr int NL_THREADS = 256;  // 16 bytes per thread

__device__ __forceinline__ unsigned newline_mask16(const uint8_t *text, int64_t n, int64_t at)
{
    unsigned mask = 0;
    if (at + 16 <= n && (reinterpret_cast<uintptr_t>(text + at) & 15) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + at));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t z = w[k] ^ 0x0A0A0A0Au;  // zero byte <=> newline
            const uint32_t hit = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u;
            mask |= (((hit >> 7) & 1u) | ((hit >> 14) & 2u) | ((hit >> 21) & 4u) | ((hit >> 28) & 8u)) << (4 * k);
        }
    } else {
        for (int k = 0; k < 16; ++k)
            if (at + k < n && text[at + k] == '\n') mask |= 1u << k;
    }
    return mask;
}

__global__ void __launch_bounds__(NL_THREADS)
count_newlines_kernel(const uint8_t *__restrict__ text, int64_t n, unsigned long long *__restrict__ block_counts)
{
    const int64_t at = blockIdx.x * (int64_t)NL_BLOCK_BYTES + threadIdx.x * 16;
    const int c = at < n ? __popc(newline_mask16(text, n, at)) : 0;
    __shared__ int warp_sums[NL_THREADS / 32];
    int v = c;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < NL_THREADS / 32; ++w) t += warp_sums[w];
        block_counts[blockIdx.x] = (unsigned long long)t;
    }
}

// block_base = exclusive scan of block_counts; positions are relative to `text`
__global__ void __launch_bounds__(NL_THREADS)
write_newlines_kernel(const uint8_t *__restrict__ text, int64_t n, const unsigned long long *__restrict__ block_base,
                      int64_t *__restrict__ positions)
{
    const int64_t at = blockIdx.x * (int64_t)NL_BLOCK_BYTES + threadIdx.x * 16;
    const unsigned mask = at < n ? newline_mask16(text, n, at) : 0u;
    const int c = __popc(mask);
    // exclusive prefix of c over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ int warp_sums[NL_THREADS / 32];
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int before = incl - c;
    for (int w = 0; w < warp; ++w) before += warp_sums[w];
    int64_t *out = positions + block_base[blockIdx.x] + before;
    unsigned m = mask;
    while (m) {
        const int k = __ffs((int)m) - 1;
        m &= m - 1;
        *out++ = at + k;
    }
}

// Sequence line of record u in each file: [start, end) offsets into the device text buffer,
// stripped of blanks at both ends like bytes.strip() (common.py:137-138); mates interleaved.
__global__ void fastq_units_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ nl1,
                                   const int64_t *__restrict__ nl2, int64_t base2, int64_t n_units, int paired,
                                   int64_t *__restrict__ starts, int64_t *__restrict__ ends, int *__restrict__ len_range)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_reads = paired ? 2 * n_units : n_units;
    if (i >= n_reads) return;
    const int64_t u = paired ? i >> 1 : i;
    const bool second = paired && (i & 1);
    const int64_t *nl = second ? nl2 : nl1;
    const int64_t base = second ? base2 : 0;
    int64_t b = base + nl[4 * u] + 1, e = base + nl[4 * u + 1];
    while (e > b && text[e - 1] <= ' ') --e;
    while (b < e && text[b] <= ' ') ++b;
    starts[i] = b;
    ends[i] = e;
    const int len = (int)(e - b);
    atomicMax(&len_range[0], len);
    atomicMin(&len_range[1], len);
}

// ---- export / merge -------------------------------------------------------------------
// Compaction of the dictionary to CSR.  One 64-bit atomic hands out the class index
// (high 32 bits) and the id start (low 32 bits: the id pool holds fewer than 2^32 ids) together,
// so starts are monotone in the class index and key_offsets is a proper CSR row pointer.  Class order is arbitrary;
// callers sort by first_unit for the reference's insertion order.
__global__ void __launch_bounds__(256)
dict_export_kernel(const DictDev d, int64_t slots, unsigned long long *cursor, int64_t *key_offsets,
                   int32_t *key_ids, int64_t *counts, int64_t *first_unit, int32_t *slot_ids)
{
    // one atomic per BLOCK on the shared cursor (class index in the high 32 bits, id start in the
    // low 32): a block-wide exclusive scan of (1, n ids) hands every occupied slot its place
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = s < slots && !(d.keys[s].x == EMPTY_KEY && d.keys[s].y == EMPTY_KEY);
    const uint32_t n = live ? d.len[s] : 0u;
    const unsigned long long mine = live ? ((1ULL << 32) | (unsigned long long)n) : 0ULL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ unsigned long long warp_total[8];
    __shared__ unsigned long long block_base;
    if (lane == 31) warp_total[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long total = 0;
        for (int w = 0; w < 8; ++w) {
            const unsigned long long t = warp_total[w];
            warp_total[w] = total;
            total += t;
        }
        block_base = total ? atomicAdd(cursor, total) : 0ULL;
    }
    __syncthreads();
    if (!live) return;
    const unsigned long long old = block_base + warp_total[warp] + (incl - mine);
    const unsigned long long c = old >> 32;
    const unsigned long long p = old & 0xFFFFFFFFULL;
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

// ---- the exported dictionary in first-seen order, on the device (skm_em_plan_from_mapper) -------
__global__ void iota_i32_kernel(int32_t *p, int64_t n)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

// lens[i] = ids of the class that comes i-th in first-seen order (lens[n] = 0 closes the scan)
__global__ void ordered_lens_kernel(const int64_t *__restrict__ off, const int32_t *__restrict__ perm, int64_t n,
                                    int64_t *__restrict__ lens)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    lens[i] = i < n ? off[perm[i] + 1] - off[perm[i]] : 0;
}

__global__ void ordered_gather_kernel(const int64_t *__restrict__ off, const int32_t *__restrict__ ids,
                                      const int64_t *__restrict__ counts, const int32_t *__restrict__ perm, int64_t n,
                                      const int64_t *__restrict__ new_off, int32_t *__restrict__ new_ids,
                                      int64_t *__restrict__ new_counts)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = perm[i];
    const int64_t src = off[c], dst = new_off[i], len = off[c + 1] - src;
    for (int64_t k = 0; k < len; ++k) new_ids[dst + k] = ids[src + k];
    new_counts[i] = counts[c];
}

__global__ void dict_merge_kernel(const DictDev d, const int64_t *key_offsets, const int32_t *key_ids,
                                  const int64_t *counts, const int64_t *first_unit, int64_t n_classes)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long added = 0;
    if (c < n_classes) {
        const int64_t start = key_offsets[c];
        const int n = (int)(key_offsets[c + 1] - start);
        if (n > 0) {
            const DenseIds ids{key_ids + start};
            const ulonglong2 key = dict_key(d, ids, n, false);
            const int64_t slot = dict_find_or_insert(d, key, ids, n, false);
            if (slot >= 0) {
                added = (unsigned long long)counts[c];
                atomicAdd(&d.counts[slot], added);
                atomicMin(&d.first[slot], (unsigned long long)first_unit[c]);
            }
        }
    }
    // the aligned total is one address for everybody: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&d.scalars[3], added);
}

// All peers of a multi-GPU exchange in one launch.  `gathered` holds one packed export per rank
// (blockIdx.y), `cap` int64 words apart: [n_classes, n_ids, unaligned | fld[2000] |
// key_offsets[n+1] | counts[n] | first_unit[n] | key_ids as int32]; the rank's own block is skipped.
__global__ void dict_merge_packed_kernel(const DictDev d, const int64_t *__restrict__ gathered, int64_t cap, int rank)
{
    const int peer = blockIdx.y;
    if (peer == rank) return;
    const int64_t *buf = gathered + (int64_t)peer * cap;
    const int64_t n_classes = buf[0];
    // a peer whose export did not fit its block only sent the header: nothing of it is read (the
    // caller checks the headers before it launches this and exchanges again with larger blocks)
    if (n_classes < 0 || buf[1] < 0 || 3 + SKM_MAX_FRAGMENT_LENGTH + 3 * n_classes + 1 + (buf[1] + 1) / 2 > cap) return;
    const int64_t *fld = buf + 3;
    const int64_t *key_offsets = fld + SKM_MAX_FRAGMENT_LENGTH;
    const int64_t *counts = key_offsets + n_classes + 1;
    const int64_t *first_unit = counts + n_classes;
    const int32_t *key_ids = reinterpret_cast<const int32_t *>(first_unit + n_classes);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < SKM_MAX_FRAGMENT_LENGTH; i += blockDim.x)
            if (fld[i]) atomicAdd(&d.fld[i], (unsigned long long)fld[i]);
        if (threadIdx.x == 0 && buf[2]) atomicAdd(&d.scalars[2], (unsigned long long)buf[2]);
    }
    const int64_t rounds = (n_classes + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
    for (int64_t k = 0; k < rounds; ++k) {  // whole warps stay together for the shuffle below
        const int64_t c = (k * gridDim.x + blockIdx.x) * (int64_t)blockDim.x + threadIdx.x;
        unsigned long long added = 0;
        if (c < n_classes) {
            const int64_t start = key_offsets[c];
            const int n = (int)(key_offsets[c + 1] - start);
            if (n > 0) {
                const DenseIds ids{key_ids + start};
                const int64_t slot = dict_find_or_insert(d, dict_key(d, ids, n, false), ids, n, false);
                if (slot >= 0) {
                    added = (unsigned long long)counts[c];
                    atomicAdd(&d.counts[slot], added);
                    atomicMin(&d.first[slot], (unsigned long long)first_unit[c]);
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
        if ((threadIdx.x & 31) == 0 && added) atomicAdd(&d.scalars[3], added);
    }
}

__global__ void add_value_kernel(unsigned long long *dst, unsigned long long v) { *dst += v; }

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
    size_t smem_max = 0;
    int threads = Q_THREADS, rows_limit = 0;  // SKM_THREADS / SKM_ROWS override for experiments
    // L2 access-policy window over the index's hot block (contigs | seq2 | targets) for the map
    // kernel: hit ratio = persisting carve-out / window size; 0 = no window
    float l2_hit_ratio = 0.f;
    size_t l2_window_bytes = 0;
    int32_t *d_out = nullptr;
    size_t d_out_cap = 0;
    cudaEvent_t ev_kernel[4] = {nullptr, nullptr, nullptr, nullptr};  // around pack | map | tally of the last chunk
    bool timed = false;
    int32_t *d_units = nullptr;     // map_reads_kernel output, tally_units_kernel input
    size_t d_units_cap = 0;
    uint64_t *d_packed = nullptr;  // pack_reads_kernel output
    size_t d_packed_cap = 0;
    int32_t *d_lens = nullptr;
    size_t d_lens_cap = 0;
    // skm_map_fastq
    uint8_t *d_text = nullptr;
    size_t d_text_cap = 0;
    int64_t *d_nl[2] = {nullptr, nullptr};
    size_t d_nl_cap[2] = {0, 0};
    unsigned long long *d_blk = nullptr;  // per-block newline counts / bases
    size_t d_blk_cap = 0;
    int64_t *d_bounds = nullptr;  // starts | ends of the reads of a FASTQ chunk
    size_t d_bounds_cap = 0;
    void *d_scan_tmp = nullptr;
    size_t d_scan_tmp_cap = 0;
    int *d_len_range = nullptr;
    // export / merge staging (host-buffer calls): one block, reused
    void *d_stage = nullptr;
    size_t d_stage_cap = 0;
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
    for (cudaEvent_t e : m->ev_kernel)
        if (e) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    cudaFree(m->d_out);
    cudaFree(m->d_packed);
    cudaFree(m->d_units);
    cudaFree(m->d_text);
    cudaFree(m->d_nl[0]);
    cudaFree(m->d_nl[1]);
    cudaFree(m->d_blk);
    cudaFree(m->d_bounds);
    cudaFree(m->d_scan_tmp);
    cudaFree(m->d_len_range);
    cudaFree(m->d_lens);
    cudaFree(m->d_stage);
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
    SKM_CUDA(cudaMemsetAsync(m->d.scalars, 0, sizeof(unsigned long long) * 8, st));
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
        // SKM_L2_PERSIST = fraction of the device's maximum persisting-L2 carve-out to claim
        // (default 0: the per-load eviction hints of DevIndex::pol_hot do the job)
        double frac = 0.0;
        if (const char *p = getenv("SKM_L2_PERSIST")) frac = atof(p);
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, m->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, m->device);
        if (frac > 0.0 && max_persist > 0 && max_window > 0 && index->hot_bytes > 0) {
            const size_t carve = (size_t)((double)max_persist * std::min(frac, 1.0));
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                m->l2_window_bytes = std::min<size_t>((size_t)index->hot_bytes, (size_t)max_window);
                m->l2_hit_ratio = (float)std::min(1.0, (double)carve / (double)m->l2_window_bytes);
            } else {
                cudaGetLastError();
            }
        }
        if (getenv("SKM_TRACE"))
            fprintf(stderr, "[skm trace] L2: max persisting %d B, max window %d B, hot block %lld B, hit ratio %.3f\n",
                    max_persist, max_window, (long long)index->hot_bytes, m->l2_hit_ratio);
    }
    // list arena (target lists longer than LIST_CAP, a bump allocator that starts over with every
    // launch): room for every resident thread to spill its largest possible list a few times
    // over; launch_chunk grows it with the units of a launch
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
    A((void **)&m->d.scalars, sizeof(unsigned long long) * 8);
    A((void **)&m->d.fld, sizeof(unsigned long long) * FLD_BINS);
    A((void **)&m->d.status, sizeof(uint32_t));
    A((void **)&m->cursors, sizeof(unsigned long long) * 4);
    A((void **)&m->arena, sizeof(int32_t) * (size_t)arena);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&m->ev_copy[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_compute[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&m->ev_kernel[i]);
    if (e != cudaSuccess) {
        skm_mapper_destroy(m);
        return fail(SKM_ERR_OOM, std::string("skm_mapper_create: ") + cudaGetErrorString(e));
    }
    m->d.mask = (uint64_t)slots - 1;
    m->d.pool_cap = (uint64_t)id_capacity;
    m->d.weak_keys = getenv("SKM_TEST_WEAK_KEYS") ? 1u : 0u;
    int rc = skm_mapper_reset(m, nullptr);
    if (rc == 0 && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = fail(SKM_ERR_CUDA, "reset failed");
    if (rc != 0) {
        skm_mapper_destroy(m);
        return rc;
    }
    *out = m;
    return SKM_OK;
}

static size_t map_smem_bytes(int code_words, int rows)
{
    return map_item_bytes(code_words) * 32 * (size_t)rows + map_fixed_bytes();
}

// map_reads_kernel is instantiated for a few pool sizes; the launch takes the largest that fits
typedef void (*map_kernel_fn)(const DevIndex, const MapArgs, uint32_t *);
struct MapVariant {
    int rows;
    map_kernel_fn fn;
};
static const MapVariant MAP_VARIANTS[] = {
    {32, map_reads_kernel<32>}, {28, map_reads_kernel<28>}, {24, map_reads_kernel<24>}, {20, map_reads_kernel<20>},
    {16, map_reads_kernel<16>}, {12, map_reads_kernel<12>}, {8, map_reads_kernel<8>},   {4, map_reads_kernel<4>},
    {2, map_reads_kernel<2>},   {1, map_reads_kernel<1>},
};

static int check_status(skm_mapper *m, cudaStream_t st, const char *who)
{
    uint32_t status = 0;
    SKM_CUDA(cudaMemcpyAsync(&status, m->d.status, sizeof(status), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaStreamSynchronize(st));
    if (status & ST_KEY_COLLISION)
        return fail(SKM_ERR_COLLISION, std::string(who) + ": two different classes share a 128-bit dictionary key; "
                                       "refusing to merge them (results of this mapper are invalid)");
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
static int launch_chunk(skm_mapper *m, const uint8_t *d_bases, const int64_t *d_starts, const int64_t *d_ends, MapArgs a,
                 int64_t n_units, int64_t first_unit, int32_t *d_out_class, int32_t *d_out_length, cudaStream_t st)
{
    if (n_units >= (1LL << 32)) return fail(SKM_ERR_INVALID, "skm_map_batch: more than 2^32 units in one chunk");
    const int64_t n_reads = a.paired ? 2 * n_units : n_units;
    int rc = ensure((void **)&m->d_packed, &m->d_packed_cap, sizeof(uint64_t) * (size_t)n_reads * a.words);
    if (rc) return rc;
    int32_t *lens = nullptr;
    if (d_starts) {
        rc = ensure((void **)&m->d_lens, &m->d_lens_cap, sizeof(int32_t) * (size_t)n_reads);
        if (rc) return rc;
        lens = m->d_lens;
    }
    SKM_CUDA(cudaEventRecord(m->ev_kernel[0], st));
    const int64_t pack_threads = n_reads * a.code_words;
    if (pack_threads >= (1LL << 31) * 256) return fail(SKM_ERR_INVALID, "skm_map_batch: batch too large for one launch");
    pack_reads_kernel<<<(unsigned)((pack_threads + 255) / 256), 256, 0, st>>>(
        d_bases, d_starts, d_ends, a.fixed_len, a.code_words, a.wild_words, a.words, n_reads, m->d_packed, lens);
    SKM_CUDA(cudaGetLastError());
    SKM_CUDA(cudaEventRecord(m->ev_kernel[1], st));
    if (m->index->max_target_count > LIST_CAP) {
        // nothing is given back inside a launch: size the arena by the launch (16 entries per unit,
        // i.e. every read spilling one list of 8 beyond LIST_CAP would still fit)
        const uint64_t want = std::min<uint64_t>((uint64_t)n_units * 16ULL, 1ULL << 30);
        if (want > m->arena_cap) {
            size_t cap_bytes = sizeof(int32_t) * (size_t)m->arena_cap;
            rc = ensure((void **)&m->arena, &cap_bytes, sizeof(int32_t) * (size_t)want);
            if (rc) return rc;
            m->arena_cap = cap_bytes / sizeof(int32_t);
        }
    }
    a.arena = m->arena;
    a.arena_cap = m->arena_cap;
    a.packed = m->d_packed;
    a.lens = lens;
    a.n_units = n_units;
    a.first_unit = first_unit;
    rc = ensure((void **)&m->d_units, &m->d_units_cap, sizeof(int32_t) * REC_ROWS * (size_t)n_units);
    if (rc) return rc;
    a.units = m->d_units;
    SKM_CUDA(cudaMemsetAsync(m->cursors, 0, sizeof(unsigned long long) * 2, st));
    // one block per SM; as many item rows as shared memory holds (at most 32: one mask bit each)
    const MapVariant *var = nullptr;
    for (const MapVariant &v : MAP_VARIANTS) {
        if (m->rows_limit > 0 && v.rows > m->rows_limit) continue;
        if (map_smem_bytes(a.code_words, v.rows) <= m->smem_max) {
            var = &v;
            break;
        }
    }
    if (!var) return fail(SKM_ERR_INVALID, "skm_map_batch: reads too long for shared-memory staging");
    const size_t smem = map_smem_bytes(a.code_words, var->rows);
    SKM_CUDA(cudaFuncSetAttribute(var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (n_units + var->rows * 32 - 1) / (var->rows * 32);
    const int grid = (int)std::min<int64_t>(m->sm_count, std::max<int64_t>(want, 1));
    // the three kernels are bracketed by events on their own stream (skm_mapper_kernel_ms); the
    // work-counter memset above sits between events 1 and 2 with the map kernel
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)m->threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        if (m->l2_hit_ratio > 0.f) {
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[0].val.accessPolicyWindow.base_ptr = m->index->hot;
            attr[0].val.accessPolicyWindow.num_bytes = m->l2_window_bytes;
            attr[0].val.accessPolicyWindow.hitRatio = m->l2_hit_ratio;
            attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
        }
        SKM_CUDA(cudaLaunchKernelEx(&cfg, var->fn, m->index->d, a, m->d.status));
    }
    SKM_CUDA(cudaGetLastError());
    SKM_CUDA(cudaEventRecord(m->ev_kernel[2], st));
    const int tally_blocks = (int)std::min<int64_t>((n_units + 255) / 256, (int64_t)m->sm_count * 8);
    tally_units_kernel<<<tally_blocks, 256, 0, st>>>(m->d, m->d_units, m->arena, n_units, first_unit, d_out_class,
                                                    d_out_length);
    SKM_CUDA(cudaEventRecord(m->ev_kernel[3], st));
    m->timed = true;
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
    if (!read_offsets && fixed_read_len <= 0)
        return fail(SKM_ERR_INVALID, "skm_map_batch: need read_offsets or a positive fixed_read_len");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int per_unit = paired ? 2 : 1;
    const int64_t n_reads = per_unit * n_units;

    if (!read_offsets) max_read_len = fixed_read_len;
    if (!buffers_on_device && read_offsets) {
        int64_t mx = 0;
        for (int64_t i = 0; i < n_reads; ++i) {
            const int64_t L = read_offsets[i + 1] - read_offsets[i];
            if (L < 0) return fail(SKM_ERR_INVALID, "skm_map_batch: read_offsets must be non-decreasing");
            mx = std::max(mx, L);
        }
        if (mx > 4096) return fail(SKM_ERR_INVALID, "skm_map_batch: reads longer than 4096 bases are not supported");
        max_read_len = (int32_t)mx;
    }
    // reads shorter than k are legal input (trimmed FASTQ): their units come out unaligned
    if (max_read_len < K) max_read_len = K;
    if (max_read_len > 4096) return fail(SKM_ERR_INVALID, "skm_map_batch: reads longer than 4096 bases are not supported");

    MapArgs a{};
    a.fixed_len = read_offsets ? 0 : fixed_read_len;
    a.code_words = (max_read_len + 31) / 32;
    a.wild_words = (max_read_len + 63) / 64;
    a.words = (a.code_words + a.wild_words + 3) & ~3;  // 32-byte records: one LDG.E.256 per four words
    a.paired = paired ? 1 : 0;
    a.arena = m->arena;
    a.arena_cap = m->arena_cap;
    a.cursors = m->cursors;
    a.short_units = m->d.scalars + 5;

    if (buffers_on_device)
        return launch_chunk(m, bases, read_offsets, read_offsets ? read_offsets + 1 : nullptr, a, n_units, first_unit,
                            out_class, out_length, st);

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
                          read_offsets ? m->d_offsets[slot] : nullptr,
                          read_offsets ? m->d_offsets[slot] + 1 : nullptr, a, nu,
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

// Newline positions of text[0, n) on the device; returns their number.
static int find_newlines(skm_mapper *m, const uint8_t *d_text, int64_t n, int which, int64_t *count, cudaStream_t st)
{
    *count = 0;
    if (n <= 0) return SKM_OK;
    const int64_t blocks = (n + NL_BLOCK_BYTES - 1) / NL_BLOCK_BYTES;
    int rc = ensure((void **)&m->d_blk, &m->d_blk_cap, sizeof(unsigned long long) * (size_t)(2 * (blocks + 1)));
    if (rc) return rc;
    unsigned long long *cnt = m->d_blk, *base = m->d_blk + blocks + 1;
    SKM_CUDA(cudaMemsetAsync(cnt + blocks, 0, sizeof(unsigned long long), st));
    count_newlines_kernel<<<(unsigned)blocks, NL_THREADS, 0, st>>>(d_text, n, cnt);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, base, (int)(blocks + 1), st);
    rc = ensure(&m->d_scan_tmp, &m->d_scan_tmp_cap, tmp);
    if (rc) return rc;
    SKM_CUDA(cub::DeviceScan::ExclusiveSum(m->d_scan_tmp, tmp, cnt, base, (int)(blocks + 1), st));
    unsigned long long total = 0;
    SKM_CUDA(cudaMemcpyAsync(&total, base + blocks, sizeof(total), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaStreamSynchronize(st));
    *count = (int64_t)total;
    if (total == 0) return SKM_OK;
    rc = ensure((void **)&m->d_nl[which], &m->d_nl_cap[which], sizeof(int64_t) * (size_t)total);
    if (rc) return rc;
    write_newlines_kernel<<<(unsigned)blocks, NL_THREADS, 0, st>>>(d_text, n, base, m->d_nl[which]);
    SKM_CUDA(cudaGetLastError());
    // the scratch (m->d_blk) is reused by the next call on the same stream: ordered
    return SKM_OK;
}

SKM_API int skm_map_fastq(skm_mapper *m, const uint8_t *text1, int64_t n1, const uint8_t *text2, int64_t n2,
                          int64_t first_unit, int buffers_on_device, int64_t *consumed1, int64_t *consumed2,
                          int64_t *n_units_out, int32_t *out_class, int32_t *out_length, void *stream)
{
    if (!m || !text1 || !consumed1 || !n_units_out) return fail(SKM_ERR_INVALID, "skm_map_fastq: NULL argument");
    if (n1 < 0 || n2 < 0 || (text2 && !consumed2)) return fail(SKM_ERR_INVALID, "skm_map_fastq: bad argument");
    if (n1 >= (1LL << 31) || n2 >= (1LL << 31))
        return fail(SKM_ERR_INVALID, "skm_map_fastq: chunks of 2 GiB or more are not supported");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int paired = text2 != nullptr;
    *consumed1 = 0;
    if (consumed2) *consumed2 = 0;
    *n_units_out = 0;
    // ---- the text in one device buffer: file 1 at 0, file 2 at base2 (16-byte aligned) ---------
    const int64_t base2 = (n1 + 15) & ~15LL;
    const uint8_t *d_text = text1;
    if (!buffers_on_device || paired) {
        int rc = ensure((void **)&m->d_text, &m->d_text_cap, (size_t)(base2 + n2 + 64));
        if (rc) return rc;
        const cudaMemcpyKind kind = buffers_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        SKM_CUDA(cudaMemcpyAsync(m->d_text, text1, (size_t)n1, kind, st));
        if (paired && n2 > 0) SKM_CUDA(cudaMemcpyAsync(m->d_text + base2, text2, (size_t)n2, kind, st));
        d_text = m->d_text;
    }
    // ---- lines, records, units --------------------------------------------------------------------
    int64_t lines1 = 0, lines2 = 0;
    int rc = find_newlines(m, d_text, n1, 0, &lines1, st);
    if (rc) return rc;
    if (paired) {
        rc = find_newlines(m, d_text + base2, n2, 1, &lines2, st);
        if (rc) return rc;
    }
    const int64_t n_units = paired ? std::min(lines1, lines2) / 4 : lines1 / 4;
    if (n_units == 0) return SKM_OK;
    const int64_t n_reads = paired ? 2 * n_units : n_units;
    rc = ensure((void **)&m->d_bounds, &m->d_bounds_cap, sizeof(int64_t) * 2 * (size_t)n_reads);
    if (rc) return rc;
    if (!m->d_len_range) SKM_CUDA(cudaMalloc(&m->d_len_range, 2 * sizeof(int)));
    const int init_range[2] = {0, INT32_MAX};
    SKM_CUDA(cudaMemcpyAsync(m->d_len_range, init_range, sizeof(init_range), cudaMemcpyHostToDevice, st));
    int64_t *d_starts = m->d_bounds, *d_ends = m->d_bounds + n_reads;
    fastq_units_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(d_text, m->d_nl[0], m->d_nl[1], base2, n_units,
                                                                         paired, d_starts, d_ends, m->d_len_range);
    SKM_CUDA(cudaGetLastError());
    int len_range[2] = {0, 0};
    int64_t last1 = 0, last2 = 0;
    SKM_CUDA(cudaMemcpyAsync(len_range, m->d_len_range, sizeof(len_range), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaMemcpyAsync(&last1, m->d_nl[0] + (4 * n_units - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (paired)
        SKM_CUDA(cudaMemcpyAsync(&last2, m->d_nl[1] + (4 * n_units - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaStreamSynchronize(st));
    if (len_range[0] > 4096) return fail(SKM_ERR_INVALID, "skm_map_fastq: reads longer than 4096 bases are not supported");
    const int longest = std::max(len_range[0], K);  // shorter reads are mapped as unaligned units
    MapArgs a{};
    a.fixed_len = 0;
    a.code_words = (longest + 31) / 32;
    a.wild_words = (longest + 63) / 64;
    a.words = (a.code_words + a.wild_words + 3) & ~3;
    a.paired = paired;
    a.arena = m->arena;
    a.arena_cap = m->arena_cap;
    a.cursors = m->cursors;
    a.short_units = m->d.scalars + 5;
    int32_t *d_class = out_class, *d_length = out_length;
    if (!buffers_on_device && (out_class || out_length)) {
        rc = ensure((void **)&m->d_out, &m->d_out_cap, sizeof(int32_t) * 2 * (size_t)n_units);
        if (rc) return rc;
        d_class = out_class ? m->d_out : nullptr;
        d_length = out_length ? m->d_out + n_units : nullptr;
    }
    rc = launch_chunk(m, d_text, d_starts, d_ends, a, n_units, first_unit, d_class, d_length, st);
    if (rc) return rc;
    if (!buffers_on_device) {
        if (out_class)
            SKM_CUDA(cudaMemcpyAsync(out_class, d_class, sizeof(int32_t) * (size_t)n_units, cudaMemcpyDeviceToHost, st));
        if (out_length)
            SKM_CUDA(cudaMemcpyAsync(out_length, d_length, sizeof(int32_t) * (size_t)n_units, cudaMemcpyDeviceToHost, st));
    }
    *consumed1 = last1 + 1;
    if (paired) *consumed2 = last2 + 1;
    *n_units_out = n_units;
    return check_status(m, st, "skm_map_fastq");  // synchronises: the text buffers may be reused
}

SKM_API int skm_classes_merge_packed(skm_mapper *m, const int64_t *gathered, int64_t words_per_rank, int world,
                                     int rank, void *stream)
{
    if (!m || !gathered) return fail(SKM_ERR_INVALID, "skm_classes_merge_packed: NULL argument");
    if (world < 1 || rank < 0 || rank >= world || words_per_rank < 3 + SKM_MAX_FRAGMENT_LENGTH)
        return fail(SKM_ERR_INVALID, "skm_classes_merge_packed: bad layout");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (world > 1) {
        const dim3 grid((unsigned)(m->sm_count * 4), (unsigned)world);
        dict_merge_packed_kernel<<<grid, 128, 0, st>>>(m->d, gathered, words_per_rank, rank);
        SKM_CUDA(cudaGetLastError());
    }
    return check_status(m, st, "skm_classes_merge_packed");
}

SKM_API int skm_mapper_kernel_ms(skm_mapper *m, double ms[3])
{
    if (!m || !ms) return fail(SKM_ERR_INVALID, "skm_mapper_kernel_ms: NULL argument");
    if (!m->timed) return fail(SKM_ERR_INVALID, "skm_mapper_kernel_ms: nothing has been mapped yet");
    SKM_CUDA(cudaSetDevice(m->device));
    SKM_CUDA(cudaEventSynchronize(m->ev_kernel[3]));
    for (int i = 0; i < 3; ++i) {
        float t = 0.f;
        SKM_CUDA(cudaEventElapsedTime(&t, m->ev_kernel[i], m->ev_kernel[i + 1]));
        ms[i] = (double)t;
    }
    return SKM_OK;
}

SKM_API int skm_debug_map_stats(uint64_t stats[32], int reset)
{
    if (stats) SKM_CUDA(cudaMemcpyFromSymbol(stats, g_map_stats, sizeof(uint64_t) * 32));
    if (reset) {
        const uint64_t zero[32] = {};
        SKM_CUDA(cudaMemcpyToSymbol(g_map_stats, zero, sizeof(zero)));
    }
    return SKM_STATS ? SKM_OK : fail(SKM_ERR_INVALID, "skm_debug_map_stats: the library was built without -DSKM_STATS=1");
}

SKM_API int skm_classes_size(skm_mapper *m, int64_t sizes[8], void *stream)
{
    if (!m || !sizes) return fail(SKM_ERR_INVALID, "skm_classes_size: NULL argument");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long sc[6];
    uint32_t status = 0;
    SKM_CUDA(cudaMemcpyAsync(sc, m->d.scalars, sizeof(sc), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaMemcpyAsync(&status, m->d.status, sizeof(status), cudaMemcpyDeviceToHost, st));
    SKM_CUDA(cudaStreamSynchronize(st));
    sizes[0] = (int64_t)sc[1];
    sizes[1] = (int64_t)sc[4];  // ids stored (the pool cursor sc[0] also counts unused chunk tails)
    sizes[2] = (int64_t)sc[2];
    sizes[3] = (int64_t)sc[3];
    sizes[4] = m->slots / 2;
    sizes[5] = (int64_t)status;
    sizes[6] = (int64_t)sc[5];  // units with a read shorter than k (they are in `unaligned`)
    sizes[7] = (int64_t)sc[0];  // id-pool cursor (== sizes[1]: the pool holds stored ids only)
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
    int64_t sizes[8];
    rc = skm_classes_size(m, sizes, stream);
    if (rc) return rc;
    const int64_t n_cls = sizes[0], n_ids = sizes[1];
    const bool host = !buffers_on_device;

    int64_t *d_off = key_offsets, *d_counts = counts, *d_first = first_unit;
    int32_t *d_ids = key_ids, *d_slots = slots;
    cudaError_t e = cudaSuccess;
    if (host) {
        // staging for the host copy: one block owned by the mapper, carved into the five arrays
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_off = up(sizeof(int64_t) * (size_t)(n_cls + 1)), b_ids = up(sizeof(int32_t) * (size_t)n_ids);
        const size_t b_cnt = up(sizeof(int64_t) * (size_t)n_cls), b_slot = up(sizeof(int32_t) * (size_t)n_cls);
        rc = ensure(&m->d_stage, &m->d_stage_cap, b_off + b_ids + 2 * b_cnt + b_slot);
        if (rc) return rc;
        char *base = static_cast<char *>(m->d_stage);
        d_off = key_offsets ? reinterpret_cast<int64_t *>(base) : nullptr;
        d_ids = key_ids ? reinterpret_cast<int32_t *>(base + b_off) : nullptr;
        d_counts = counts ? reinterpret_cast<int64_t *>(base + b_off + b_ids) : nullptr;
        d_first = first_unit ? reinterpret_cast<int64_t *>(base + b_off + b_ids + b_cnt) : nullptr;
        d_slots = slots ? reinterpret_cast<int32_t *>(base + b_off + b_ids + 2 * b_cnt) : nullptr;
    }
    auto cleanup = []() {};
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
    const int64_t *p_off = key_offsets, *p_cnt = counts, *p_first = first_unit, *p_fld = fld;
    const int32_t *p_ids = key_ids;
    if (!buffers_on_device) {
        // host arrays go through the mapper's staging block (one allocation, reused)
        const int64_t n_ids = n_classes > 0 ? key_offsets[n_classes] : 0;
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_off = up(sizeof(int64_t) * (size_t)(n_classes + 1)), b_ids = up(sizeof(int32_t) * (size_t)n_ids);
        const size_t b_cnt = up(sizeof(int64_t) * (size_t)n_classes), b_fld = up(sizeof(int64_t) * FLD_BINS);
        int rc = ensure(&m->d_stage, &m->d_stage_cap, b_off + b_ids + 2 * b_cnt + b_fld);
        if (rc) return rc;
        char *base = static_cast<char *>(m->d_stage);
        int64_t *d_off = reinterpret_cast<int64_t *>(base);
        int32_t *d_ids = reinterpret_cast<int32_t *>(base + b_off);
        int64_t *d_cnt = reinterpret_cast<int64_t *>(base + b_off + b_ids);
        int64_t *d_first = reinterpret_cast<int64_t *>(base + b_off + b_ids + b_cnt);
        int64_t *d_fld = reinterpret_cast<int64_t *>(base + b_off + b_ids + 2 * b_cnt);
        if (n_classes > 0) {
            SKM_CUDA(cudaMemcpyAsync(d_off, key_offsets, sizeof(int64_t) * (size_t)(n_classes + 1), cudaMemcpyHostToDevice, st));
            if (n_ids > 0) SKM_CUDA(cudaMemcpyAsync(d_ids, key_ids, sizeof(int32_t) * (size_t)n_ids, cudaMemcpyHostToDevice, st));
            SKM_CUDA(cudaMemcpyAsync(d_cnt, counts, sizeof(int64_t) * (size_t)n_classes, cudaMemcpyHostToDevice, st));
            SKM_CUDA(cudaMemcpyAsync(d_first, first_unit, sizeof(int64_t) * (size_t)n_classes, cudaMemcpyHostToDevice, st));
        }
        if (fld) SKM_CUDA(cudaMemcpyAsync(d_fld, fld, sizeof(int64_t) * FLD_BINS, cudaMemcpyHostToDevice, st));
        p_off = d_off;
        p_ids = d_ids;
        p_cnt = d_cnt;
        p_first = d_first;
        p_fld = d_fld;
    }
    if (n_classes > 0) {
        dict_merge_kernel<<<(unsigned)((n_classes + 127) / 128), 128, 0, st>>>(m->d, p_off, p_ids, p_cnt, p_first, n_classes);
        SKM_CUDA(cudaGetLastError());
    }
    if (fld) {
        add_i64_kernel<<<(FLD_BINS + 255) / 256, 256, 0, st>>>(m->d.fld, p_fld, FLD_BINS);
        SKM_CUDA(cudaGetLastError());
    }
    if (unaligned > 0) {
        add_value_kernel<<<1, 1, 0, st>>>(m->d.scalars + 2, (unsigned long long)unaligned);
        SKM_CUDA(cudaGetLastError());
    }
    return check_status(m, st, "skm_classes_merge");  // synchronises: the staging block may be reused
}

// Replaces the host round trip MapResult.summarize -> quantify (mapper.py:77-104, infer.py:88-130):
// the dictionary becomes the EM's class structure where it lies.  Classes are ordered by their
// first-seen unit (the Counter's insertion order at job_count=1, which is the order of
// class_map / class_count), ids stay in tuple order, counts stay integers.
SKM_API int skm_em_plan_from_mapper(skm_mapper *m, int64_t n_transcripts, void *stream, skm_em_plan **out)
{
    if (!out) return fail(SKM_ERR_INVALID, "skm_em_plan_from_mapper: out is NULL");
    *out = nullptr;
    if (!m) return fail(SKM_ERR_INVALID, "skm_em_plan_from_mapper: NULL mapper");
    if (n_transcripts <= 0) n_transcripts = m->index->n_transcripts;
    if (n_transcripts <= 0 || n_transcripts >= (1LL << 31))
        return fail(SKM_ERR_INVALID, "skm_em_plan_from_mapper: bad transcript count");
    SKM_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_status(m, st, "skm_em_plan_from_mapper");
    if (rc) return rc;
    int64_t sizes[8];
    rc = skm_classes_size(m, sizes, stream);
    if (rc) return rc;
    const int64_t n = sizes[0], n_ids = sizes[1];
    if (n <= 0 || n_ids <= 0) return fail(SKM_ERR_INVALID, "skm_em_plan_from_mapper: the dictionary is empty");
    // raw export (table order) + sort scratch in the mapper's staging block
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_off = up(sizeof(int64_t) * (size_t)(n + 1)), b_ids = up(sizeof(int32_t) * (size_t)n_ids);
    const size_t b_i64 = up(sizeof(int64_t) * (size_t)(n + 1)), b_i32 = up(sizeof(int32_t) * (size_t)n);
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, 64, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (const int64_t *)nullptr, (int64_t *)nullptr, (int)(n + 1), st);
    const size_t b_tmp = up(std::max(tmp_sort, tmp_scan));
    rc = ensure(&m->d_stage, &m->d_stage_cap, b_off + b_ids + 4 * b_i64 + 2 * b_i32 + b_tmp);
    if (rc) return rc;
    char *base = static_cast<char *>(m->d_stage);
    int64_t *r_off = reinterpret_cast<int64_t *>(base);
    int32_t *r_ids = reinterpret_cast<int32_t *>(base + b_off);
    int64_t *r_cnt = reinterpret_cast<int64_t *>(base + b_off + b_ids);
    int64_t *r_first = reinterpret_cast<int64_t *>(base + b_off + b_ids + b_i64);
    int64_t *s_first = reinterpret_cast<int64_t *>(base + b_off + b_ids + 2 * b_i64);
    int64_t *lens = reinterpret_cast<int64_t *>(base + b_off + b_ids + 3 * b_i64);
    int32_t *iota = reinterpret_cast<int32_t *>(base + b_off + b_ids + 4 * b_i64);
    int32_t *perm = reinterpret_cast<int32_t *>(base + b_off + b_ids + 4 * b_i64 + b_i32);
    void *tmp = base + b_off + b_ids + 4 * b_i64 + 2 * b_i32;
    SKM_CUDA(cudaMemsetAsync(m->cursors + 2, 0, sizeof(unsigned long long), st));
    dict_export_kernel<<<(unsigned)((m->slots + 255) / 256), 256, 0, st>>>(m->d, m->slots, m->cursors + 2, r_off, r_ids,
                                                                          r_cnt, r_first, nullptr);
    set_i64_kernel<<<1, 1, 0, st>>>(r_off + n, n_ids);
    iota_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(iota, n);
    SKM_CUDA(cudaGetLastError());
    size_t tb = b_tmp;
    SKM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, reinterpret_cast<const unsigned long long *>(r_first),
                                             reinterpret_cast<unsigned long long *>(s_first), iota, perm, (int)n, 0, 64, st));
    // the plan's own arrays
    int64_t *p_off = nullptr, *p_cnt = nullptr;
    int32_t *p_ids = nullptr;
    cudaError_t e = dev_alloc(m->device, sizeof(int64_t) * (size_t)(n + 1), (void **)&p_off);
    if (e == cudaSuccess) e = dev_alloc(m->device, sizeof(int32_t) * (size_t)n_ids, (void **)&p_ids);
    if (e == cudaSuccess) e = dev_alloc(m->device, sizeof(int64_t) * (size_t)n, (void **)&p_cnt);
    if (e == cudaSuccess) {
        ordered_lens_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(r_off, perm, n, lens);
        tb = b_tmp;
        e = cub::DeviceScan::ExclusiveSum(tmp, tb, lens, p_off, (int)(n + 1), st);
    }
    if (e == cudaSuccess) {
        ordered_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r_off, r_ids, r_cnt, perm, n, p_off, p_ids, p_cnt);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the staging block may be reused after this
    if (e != cudaSuccess) {
        dev_free(m->device, p_off);
        dev_free(m->device, p_ids);
        dev_free(m->device, p_cnt);
        return fail(e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA,
                    std::string("skm_em_plan_from_mapper: ") + cudaGetErrorString(e));
    }
    return em_plan_adopt(m->device, n, n_ids, n_transcripts, p_off, p_ids, p_cnt, st, out);
}
