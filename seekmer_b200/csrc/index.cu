// Index upload and re-layout: KMerIndex arrays (reference layout) -> HBM-resident
// canonical-key hash table in 4-slot buckets + 128-byte contig records (header, 8 inline
// targets, graph links) + 2-bit sequence pool + entry-only target lists.  Replaces
// KMerIndex.__init__/load on the device side (_common.pyx:21-48,287-313) and KMerIndex.map_kmer
// for vectors of k-mers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "em_plan.cuh"
#include "kmer.cuh"

namespace skm {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg)
{
    g_error = msg;
    return code;
}
const char *last_error() { return g_error.c_str(); }

// ---------------------------------------------------------------------------------
__global__ void count_occupied_kernel(const skm_kmer_slot *__restrict__ src, int64_t n,
                                      unsigned long long *__restrict__ count)
{
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        local += __ldg(&src[i].kmer) != EMPTY_KEY;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// Re-insert every occupied reference slot under its canonical key.  The stored
// coordinate is expressed for the canonical orientation: if the contig-forward k-mer
// (what the reference stores) is not the canonical one, the entry is
// bit-negated, so that map_kmer() in kmer.cuh returns exactly what
// _common.pyx:82-87 returns for either query strand.
__global__ void relayout_table_kernel(const skm_kmer_slot *__restrict__ src, int64_t n,
                                      Slot *__restrict__ table, uint64_t mask, int64_t n_contigs,
                                      unsigned int *bad)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(src + i));
        const uint64_t kmer = v.x;
        if (kmer == EMPTY_KEY) continue;
        const uint64_t rc = revcomp(kmer);
        const bool fwd = kmer < rc;
        const uint64_t canon = fwd ? kmer : rc;
        int32_t entry = (int32_t)(uint32_t)v.y;
        const int32_t offset = (int32_t)(uint32_t)(v.y >> 32);
        // a usable position (offset >= 0) must name an existing contig; slots the reference
        // assembler left with a negative offset act as misses and are kept verbatim
        if (offset >= 0 && (entry < 0 || entry >= n_contigs)) atomicOr(bad, 2u);
        if (!fwd) entry = ~entry;
        // first free slot from the home bucket onwards (slots fill front to back per bucket)
        uint64_t s = (uint64_t)home_bucket_of(canon, mask / BUCKET_SLOTS) * BUCKET_SLOTS;
        for (;;) {
            const unsigned long long old = atomicCAS(
                reinterpret_cast<unsigned long long *>(&table[s].key), EMPTY_KEY, canon);
            if (old == EMPTY_KEY) {
                table[s].entry = entry;
                table[s].offset = offset;
                break;
            }
            s = (s + 1) & mask;
        }
    }
}

__global__ void relayout_contigs_kernel(const skm_contig_entry *__restrict__ src, int64_t n,
                                        const skm_target *__restrict__ targets,
                                        ContigRec *__restrict__ dst, int64_t n_bases,
                                        int64_t n_targets, unsigned long long *max_tc,
                                        unsigned int *bad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const skm_contig_entry c = src[i];
    // The reference assembler can emit a degenerate contig (shorter than k, no targets, not
    // referenced by any k-mer: slot 0 doubles as "no link" there), so only
    // memory safety is enforced here.
    const bool broken = c.length < 0 || c.length > 0x7FFFFFFFLL || c.offset < 0 || c.offset + c.length > n_bases
                        || c.target_offset < 0 || c.target_count < 0 || c.target_count >= (1LL << 28)
                        || c.target_offset + c.target_count > n_targets;
    if (broken) atomicOr(bad, 1u);
    ContigRec r;
    for (int j = 0; j < INLINE_TARGETS; ++j)
        r.inline_targets[j] = (!broken && j < c.target_count) ? targets[c.target_offset + j].entry : 0;
    const uint64_t tc = (uint64_t)c.target_count;
    r.w0 = (c.first_kmer & KMER_MASK) | ((tc & 0x3FFF) << 50);
    r.w1 = (c.last_kmer & KMER_MASK) | (((tc >> 14) & 0x3FFF) << 50);
    r.seq_offset = c.offset;
    r.target_offset = (uint32_t)c.target_offset;
    r.length = (uint32_t)c.length;
    dst[i] = r;
    atomicMax(max_tc, (unsigned long long)tc);
}

// Graph links (common.cuh): 8 probes of the finished device table per contig.
__global__ void contig_links_kernel(const DevIndex ix, ContigRec *__restrict__ recs, int64_t n_contigs)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_contigs * 8) return;
    const int64_t c = i >> 3;
    const int k = (int)(i & 7);
    const uint64_t b = (uint64_t)(k & 3);
    const uint64_t first = recs[c].w0 & KMER_MASK, last = recs[c].w1 & KMER_MASK;
    const uint64_t query = k < 4 ? ((last << 2) | b) & KMER_MASK              // append(last_kmer, b), _kmer.pxd:71-86
                                 : (first >> 2) | (b << (2 * K - 2));          // prepend(first_kmer, b), :89-106
    const Coord hit = map_kmer(ix, query);
    int2 *dst = k < 4 ? recs[c].right_of_last : recs[c].left_of_first;
    dst[k & 3] = make_int2(hit.entry, hit.offset);
}

// L2 cache policies for the mapper's loads (DevIndex::pol_hot / pol_stream): kind 0 = normal,
// 1 = evict_last, 2 = evict_first.  A policy is an opaque 64-bit value made on the device.
__global__ void make_policies_kernel(uint64_t *out, int hot_kind, int stream_kind)
{
    for (int i = 0; i < 2; ++i) {
        const int kind = i == 0 ? hot_kind : stream_kind;
        uint64_t p;
        if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
        else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
        else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
        out[i] = p;
    }
}

__device__ __forceinline__ uint32_t code_of(uint8_t b)  // _kmer.pxd:253-273
{
    const uint8_t u = b & 0xDF;
    return u == 'T' ? 3u : u == 'G' ? 2u : u == 'C' ? 1u : 0u;
}

__global__ void pack_sequences_kernel(const uint8_t *__restrict__ src, int64_t n_bases,
                                      uint32_t *__restrict__ dst, int64_t n_words)
{
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t acc = 0;
    const int64_t base = w * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int64_t p = base + j;
        const uint32_t c = p < n_bases ? code_of(__ldg(src + p)) : 0u;
        acc = (acc << 2) | c;
    }
    dst[w] = acc;
}

__global__ void extract_targets_kernel(const skm_target *__restrict__ src, int64_t n,
                                       int32_t *__restrict__ dst)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i].entry;
}

__global__ void map_kmers_kernel(DevIndex ix, const uint64_t *__restrict__ kmers, int64_t n,
                                 int32_t *__restrict__ out_entry, int32_t *__restrict__ out_offset)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const Coord c = map_kmer(ix, kmers[i] & KMER_MASK);
        out_entry[i] = c.entry;
        out_offset[i] = c.offset;
    }
}

// Index-construction support: place contig-forward k-mers into a table in the REFERENCE
// layout (home slot = reference SipHash variant of the canonical k-mer, linear probing,
// _index_builder.pyx:313-342), so that indexes built on the device are byte-compatible
// with what `seekmer index` writes.
__global__ void build_reference_table_kernel(const uint64_t *__restrict__ kmers,
                                             const int32_t *__restrict__ entry,
                                             const int32_t *__restrict__ offset, int64_t n,
                                             skm_kmer_slot *__restrict__ table, uint64_t mask)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t kmer = kmers[i] & KMER_MASK;
        const uint64_t rc = revcomp(kmer);
        uint64_t s = reference_hash(kmer < rc ? kmer : rc) & mask;
        for (;;) {
            const unsigned long long old = atomicCAS(
                reinterpret_cast<unsigned long long *>(&table[s].kmer), EMPTY_KEY, kmer);
            if (old == EMPTY_KEY) {
                table[s].entry = entry[i];
                table[s].offset = offset[i];
                break;
            }
            s = (s + 1) & mask;
        }
    }
}

template <typename T>
static int to_device(const T *src, int64_t n, bool on_device, cudaStream_t st, const T **out,
                     T **owned)
{
    *owned = nullptr;
    if (on_device || n == 0) {
        *out = src;
        return 0;
    }
    T *buf = nullptr;
    SKM_CUDA(cudaMalloc(&buf, sizeof(T) * (size_t)n));
    cudaError_t e = cudaMemcpyAsync(buf, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        cudaFree(buf);
        return fail(SKM_ERR_CUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
    }
    *out = buf;
    *owned = buf;
    return 0;
}

}  // namespace skm

using namespace skm;

// DevIndex::pol_hot / pol_stream for an index whose arrays are in place
static int make_policies(skm_index *ix, cudaStream_t st)
{
    // SKM_L2_HINT = last | normal: eviction priority of the contig-side loads (default: last)
    const char *hint = getenv("SKM_L2_HINT");
    const int hot_kind = (hint && hint[0] == 'n') ? 0 : 1;
    uint64_t *d_pol = nullptr, pol[2] = {0, 0};
    cudaError_t e = cudaMalloc(&d_pol, sizeof(pol));
    if (e == cudaSuccess) {
        skm::make_policies_kernel<<<1, 1, 0, st>>>(d_pol, hot_kind, 2);
        e = cudaMemcpyAsync(pol, d_pol, sizeof(pol), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(d_pol);
    if (e != cudaSuccess)
        return skm::fail(SKM_ERR_CUDA, std::string("skm_index: cache policies: ") + cudaGetErrorString(e));
    ix->d.pol_hot = pol[0];
    ix->d.pol_stream = pol[1];
    return 0;
}

SKM_API const char *skm_last_error(void) { return skm::last_error(); }

SKM_API const char *skm_version(void) { return "seekmer_b200 0.1 (sm_100a)"; }

SKM_API int skm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

SKM_API void skm_index_destroy(skm_index *ix)
{
    if (!ix) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ix->device);
    cudaFree(ix->table);
    cudaFree(ix->hot);  // contigs, seq2 and targets live in this one block
    cudaSetDevice(prev);
    delete ix;
}

SKM_API int skm_index_create(const skm_kmer_slot *kmers, int64_t n_slots,
                                const skm_contig_entry *contigs, int64_t n_contigs,
                                const char *sequences, int64_t n_bases, const skm_target *targets,
                                int64_t n_targets, int64_t n_transcripts, int device,
                                int inputs_on_device, void *stream, skm_index **out)
{
    if (!out) return fail(SKM_ERR_INVALID, "skm_index_create: out is NULL");
    *out = nullptr;
    if (!kmers || !contigs || !sequences || !targets)
        return fail(SKM_ERR_INVALID, "skm_index_create: NULL index array");
    if (n_slots <= 0 || (n_slots & (n_slots - 1)) != 0)
        return fail(SKM_ERR_INVALID, "skm_index_create: k-mer table size must be a power of two");
    if (n_contigs <= 0 || n_bases <= 0 || n_targets <= 0)
        return fail(SKM_ERR_INVALID, "skm_index_create: empty index");
    if (n_targets >= (1LL << 32))
        return fail(SKM_ERR_INVALID, "skm_index_create: more than 2^32 targets");
    if (n_contigs >= (1LL << 31))
        return fail(SKM_ERR_INVALID, "skm_index_create: more than 2^31 contigs");
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_index_create: no such CUDA device (there is no CPU fallback)");
    SKM_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool dev = inputs_on_device != 0;

    skm_index *ix = new skm_index();
    ix->device = device;
    ix->n_contigs = n_contigs;
    ix->n_bases = n_bases;
    ix->n_targets = n_targets;
    ix->n_transcripts = n_transcripts;

    unsigned long long *d_scalars = nullptr;  // [0]=count, [1]=max_tc, [2]=bad
    const skm_kmer_slot *d_kmers = nullptr;
    skm_kmer_slot *own_kmers = nullptr;
    const skm_contig_entry *d_contigs = nullptr;
    skm_contig_entry *own_contigs = nullptr;
    const char *d_seq = nullptr;
    char *own_seq = nullptr;
    const skm_target *d_targets = nullptr;
    skm_target *own_targets = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_scalars);
        cudaFree(own_kmers);
        cudaFree(own_contigs);
        cudaFree(own_seq);
        cudaFree(own_targets);
    };
#define STEP(expr)                 \
    do {                           \
        rc = (expr);               \
        if (rc != 0) {             \
            cleanup();             \
            skm_index_destroy(ix); \
            return rc;             \
        }                          \
    } while (0)
#define STEP_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            cleanup();                                                                    \
            skm_index_destroy(ix);                                                        \
            return fail(_e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA,     \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));              \
        }                                                                                 \
    } while (0)

    STEP_CUDA(cudaMalloc(&d_scalars, 4 * sizeof(unsigned long long)));
    STEP_CUDA(cudaMemsetAsync(d_scalars, 0, 4 * sizeof(unsigned long long), st));

    // -- k-mer table
    STEP(to_device(kmers, n_slots, dev, st, &d_kmers, &own_kmers));
    {
        const int threads = 256;
        const int blocks = (int)std::min<int64_t>((n_slots + threads - 1) / threads, 148 * 16);
        count_occupied_kernel<<<blocks, threads, 0, st>>>(d_kmers, n_slots, d_scalars);
        unsigned long long cnt = 0;
        STEP_CUDA(cudaMemcpyAsync(&cnt, d_scalars, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        STEP_CUDA(cudaStreamSynchronize(st));
        ix->n_kmers = (int64_t)cnt;
        // load factor <= 0.25 with 4-slot buckets: a lookup, hit or miss, is answered by its home
        // bucket (one 64-byte burst) except when 4+ keys share it (~0.4 %); HBM is plentiful
        int64_t slots = 1024;
        while (slots < 4 * ix->n_kmers) slots <<= 1;
        ix->n_slots = slots;
        STEP_CUDA(cudaMalloc(&ix->table, sizeof(Slot) * (size_t)slots));
        STEP_CUDA(cudaMemsetAsync(ix->table, 0xFF, sizeof(Slot) * (size_t)slots, st));
        relayout_table_kernel<<<blocks, threads, 0, st>>>(d_kmers, n_slots, ix->table,
                                                          (uint64_t)slots - 1, n_contigs,
                                                          reinterpret_cast<unsigned int *>(d_scalars + 2));
        STEP_CUDA(cudaGetLastError());
        ix->bytes += (int64_t)sizeof(Slot) * slots;
    }
    // -- everything a contig walk touches (contig records, 2-bit sequences, target lists) sits
    //    in ONE allocation, so that a single L2 access-policy window can keep it resident while
    //    the 4 GB table and the reads stream through (mapper.cu: launch_chunk)
    const int64_t n_seq_words = (n_bases + 15) / 16 + 2;
    {
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_contigs = up(sizeof(ContigRec) * (size_t)n_contigs);
        const size_t b_seq = up(sizeof(uint32_t) * (size_t)n_seq_words);
        const size_t b_targets = up(sizeof(int32_t) * (size_t)n_targets);
        STEP_CUDA(cudaMalloc(&ix->hot, b_contigs + b_seq + b_targets));
        ix->hot_bytes = (int64_t)(b_contigs + b_seq + b_targets);
        ix->bytes += ix->hot_bytes;
        ix->contigs = reinterpret_cast<ContigRec *>(ix->hot);
        ix->seq2 = reinterpret_cast<uint32_t *>(ix->hot + b_contigs);
        ix->targets = reinterpret_cast<int32_t *>(ix->hot + b_contigs + b_seq);
    }
    // -- targets
    STEP(to_device(targets, n_targets, dev, st, &d_targets, &own_targets));
    extract_targets_kernel<<<(unsigned)((n_targets + 255) / 256), 256, 0, st>>>(
        d_targets, n_targets, ix->targets);
    STEP_CUDA(cudaGetLastError());

    // -- contigs
    STEP(to_device(contigs, n_contigs, dev, st, &d_contigs, &own_contigs));
    relayout_contigs_kernel<<<(unsigned)((n_contigs + 255) / 256), 256, 0, st>>>(
        d_contigs, n_contigs, d_targets, ix->contigs, n_bases, n_targets, d_scalars + 1,
        reinterpret_cast<unsigned int *>(d_scalars + 2));
    STEP_CUDA(cudaGetLastError());
    // -- sequences (2-bit, one padding word so window reads may touch word+1)
    STEP(to_device(sequences, n_bases, dev, st, &d_seq, &own_seq));
    {
        const int64_t n_words = n_seq_words;
        pack_sequences_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const uint8_t *>(d_seq), n_bases, ix->seq2, n_words);
        STEP_CUDA(cudaGetLastError());
    }
    unsigned long long scal[4] = {0, 0, 0, 0};
    STEP_CUDA(cudaMemcpyAsync(scal, d_scalars, sizeof(scal), cudaMemcpyDeviceToHost, st));
    STEP_CUDA(cudaStreamSynchronize(st));
    ix->max_target_count = (int64_t)scal[1];
    if (scal[2] != 0) {
        cleanup();
        skm_index_destroy(ix);
        return fail(SKM_ERR_INVALID, (scal[2] & 2u)
                        ? "skm_index_create: a k-mer slot points to a contig that does not exist"
                        : "skm_index_create: contig table is inconsistent with sequences/targets "
                          "(offset, length or target range out of bounds)");
    }
    cleanup();
#undef STEP
#undef STEP_CUDA
    ix->d.table = ix->table;
    ix->d.bucket_mask = (uint64_t)ix->n_slots / BUCKET_SLOTS - 1;
    ix->d.contigs = ix->contigs;
    ix->d.seq2 = ix->seq2;
    ix->d.targets = ix->targets;
    ix->d.n_contigs = n_contigs;
    ix->d.n_bases = n_bases;
    ix->d.n_targets = n_targets;
    {
        const int prc = make_policies(ix, st);
        if (prc) {
            skm_index_destroy(ix);
            return prc;
        }
    }
    contig_links_kernel<<<(unsigned)((n_contigs * 8 + 255) / 256), 256, 0, st>>>(ix->d, ix->contigs, n_contigs);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        skm_index_destroy(ix);
        return fail(SKM_ERR_CUDA, "skm_index_create: contig link kernel failed");
    }
    skm::scratch_warm(ix->device);  // the EM scratch cache takes its first blocks now, not inside the first EM call
    *out = ix;
    return SKM_OK;
}

SKM_API int skm_index_info(const skm_index *ix, int64_t info[8])
{
    if (!ix || !info) return fail(SKM_ERR_INVALID, "skm_index_info: NULL argument");
    info[0] = ix->n_kmers;
    info[1] = ix->n_slots;
    info[2] = ix->max_target_count;
    info[3] = ix->bytes;
    info[4] = ix->n_contigs;
    info[5] = ix->n_targets;
    info[6] = ix->n_transcripts;
    info[7] = ix->device;
    return SKM_OK;
}

SKM_API int skm_map_kmers(const skm_index *ix, const uint64_t *kmers, int64_t n,
                             int32_t *out_entry, int32_t *out_offset, int buffers_on_device,
                             void *stream)
{
    if (!ix || !kmers || !out_entry || !out_offset)
        return fail(SKM_ERR_INVALID, "skm_map_kmers: NULL argument");
    if (n <= 0) return SKM_OK;
    SKM_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t *d_k = kmers;
    int32_t *d_e = out_entry, *d_o = out_offset;
    uint64_t *own_k = nullptr;
    int32_t *own_e = nullptr;
    if (!buffers_on_device) {
        SKM_CUDA(cudaMalloc(&own_k, sizeof(uint64_t) * (size_t)n));
        if (cudaMalloc(&own_e, sizeof(int32_t) * 2 * (size_t)n) != cudaSuccess) {
            cudaFree(own_k);
            return fail(SKM_ERR_OOM, "skm_map_kmers: cudaMalloc failed");
        }
        cudaMemcpyAsync(own_k, kmers, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, st);
        d_k = own_k;
        d_e = own_e;
        d_o = own_e + n;
    }
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 148 * 8);
    map_kmers_kernel<<<blocks, threads, 0, st>>>(ix->d, d_k, n, d_e, d_o);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !buffers_on_device) {
        cudaMemcpyAsync(out_entry, d_e, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_offset, d_o, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
    }
    cudaFree(own_k);
    cudaFree(own_e);
    if (e != cudaSuccess) return fail(SKM_ERR_CUDA, std::string("skm_map_kmers: ") + cudaGetErrorString(e));
    return SKM_OK;
}

SKM_API int skm_build_kmer_table(const uint64_t *kmers, const int32_t *entry, const int32_t *offset,
                                 int64_t n, skm_kmer_slot *table, int64_t n_slots, int device,
                                 void *stream)
{
    if (!kmers || !entry || !offset || !table) return fail(SKM_ERR_INVALID, "skm_build_kmer_table: NULL argument");
    if (n_slots <= 0 || (n_slots & (n_slots - 1)) != 0 || n >= n_slots)
        return fail(SKM_ERR_INVALID, "skm_build_kmer_table: table size must be a power of two larger than n");
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_build_kmer_table: no such CUDA device");
    SKM_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA(cudaMemsetAsync(table, 0xFF, sizeof(skm_kmer_slot) * (size_t)n_slots, st));
    if (n > 0) {
        const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
        build_reference_table_kernel<<<blocks, 256, 0, st>>>(kmers, entry, offset, n, table,
                                                            (uint64_t)n_slots - 1);
        SKM_CUDA(cudaGetLastError());
    }
    return SKM_OK;
}

// ---- the device image on disk -----------------------------------------------------------------
// a native GPU-layout index file.  The HBM-resident form (bucketed canonical-key
// table, contig records with inline targets and graph links, 2-bit sequences, target entries)
// is written as it lies and read back with plain copies: loading skips the relayout kernels
// and the 8 table probes per contig of contig_links_kernel.  A trailer of caller-owned bytes
// rides along (the Python layer keeps the transcript table there).
namespace {

constexpr char IMAGE_MAGIC[8] = {'S', 'K', 'M', 'B', '2', '0', '0', '\0'};
constexpr uint32_t IMAGE_VERSION = 1;

struct ImageHeader {
    char magic[8];
    uint32_t version;
    uint32_t kmer_size, bucket_slots, slot_bytes, contig_bytes, inline_targets;
    int64_t n_slots, n_kmers, n_contigs, n_bases, n_targets, n_transcripts, max_target_count;
    int64_t table_bytes, hot_bytes, off_seq2, off_targets;  // offsets inside the hot block
    int64_t trailer_bytes;
    uint64_t check;  // sum of the fields above, a cheap guard against truncated / foreign files
};

uint64_t header_check(const ImageHeader &h)
{
    const int64_t f[] = {h.version, h.kmer_size, h.bucket_slots, h.slot_bytes, h.contig_bytes, h.inline_targets,
                         h.n_slots, h.n_kmers, h.n_contigs, h.n_bases, h.n_targets, h.n_transcripts,
                         h.max_target_count, h.table_bytes, h.hot_bytes, h.off_seq2, h.off_targets, h.trailer_bytes};
    uint64_t s = 0x9E3779B97F4A7C15ULL;
    for (int64_t v : f) s = (s ^ (uint64_t)v) * 0xBF58476D1CE4E5B9ULL + 0x632BE59BD9B4E019ULL;
    return s;
}

constexpr size_t IO_CHUNK = 64u << 20;

int copy_out(FILE *f, const void *d_src, size_t bytes, void *pinned, cudaStream_t st)
{
    const char *src = static_cast<const char *>(d_src);
    for (size_t done = 0; done < bytes; done += IO_CHUNK) {
        const size_t n = std::min(IO_CHUNK, bytes - done);
        SKM_CUDA(cudaMemcpyAsync(pinned, src + done, n, cudaMemcpyDeviceToHost, st));
        SKM_CUDA(cudaStreamSynchronize(st));
        if (fwrite(pinned, 1, n, f) != n) return fail(SKM_ERR_INVALID, "skm_index_save: short write");
    }
    return 0;
}

int copy_in(FILE *f, void *d_dst, size_t bytes, void *pinned[2], cudaStream_t st)
{
    // file -> pinned buffer k while the copy out of buffer 1-k is in flight
    char *dst = static_cast<char *>(d_dst);
    int k = 0;
    for (size_t done = 0; done < bytes; done += IO_CHUNK, k ^= 1) {
        const size_t n = std::min(IO_CHUNK, bytes - done);
        if (fread(pinned[k], 1, n, f) != n) return fail(SKM_ERR_INVALID, "skm_index_load: the file is truncated");
        SKM_CUDA(cudaStreamSynchronize(st));  // the previous copy (other buffer) is done before it is refilled next round
        SKM_CUDA(cudaMemcpyAsync(dst + done, pinned[k], n, cudaMemcpyHostToDevice, st));
    }
    SKM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace

SKM_API int skm_index_save(const skm_index *ix, const char *path, const void *trailer, int64_t trailer_bytes,
                           void *stream)
{
    if (!ix || !path || trailer_bytes < 0 || (trailer_bytes > 0 && !trailer))
        return fail(SKM_ERR_INVALID, "skm_index_save: bad argument");
    SKM_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = (cudaStream_t)stream;
    ImageHeader h{};
    memcpy(h.magic, IMAGE_MAGIC, 8);
    h.version = IMAGE_VERSION;
    h.kmer_size = K;
    h.bucket_slots = BUCKET_SLOTS;
    h.slot_bytes = sizeof(Slot);
    h.contig_bytes = sizeof(ContigRec);
    h.inline_targets = INLINE_TARGETS;
    h.n_slots = ix->n_slots;
    h.n_kmers = ix->n_kmers;
    h.n_contigs = ix->n_contigs;
    h.n_bases = ix->n_bases;
    h.n_targets = ix->n_targets;
    h.n_transcripts = ix->n_transcripts;
    h.max_target_count = ix->max_target_count;
    h.table_bytes = (int64_t)sizeof(Slot) * ix->n_slots;
    h.hot_bytes = ix->hot_bytes;
    h.off_seq2 = (int64_t)(reinterpret_cast<const unsigned char *>(ix->seq2) - ix->hot);
    h.off_targets = (int64_t)(reinterpret_cast<const unsigned char *>(ix->targets) - ix->hot);
    h.trailer_bytes = trailer_bytes;
    h.check = header_check(h);
    FILE *f = fopen(path, "wb");
    if (!f) return fail(SKM_ERR_INVALID, std::string("skm_index_save: cannot open ") + path);
    void *pinned = nullptr;
    int rc = 0;
    if (cudaMallocHost(&pinned, IO_CHUNK) != cudaSuccess) rc = fail(SKM_ERR_OOM, "skm_index_save: no pinned staging buffer");
    if (!rc && fwrite(&h, sizeof(h), 1, f) != 1) rc = fail(SKM_ERR_INVALID, "skm_index_save: short write");
    if (!rc) rc = copy_out(f, ix->table, (size_t)h.table_bytes, pinned, st);
    if (!rc) rc = copy_out(f, ix->hot, (size_t)h.hot_bytes, pinned, st);
    if (!rc && trailer_bytes > 0 && fwrite(trailer, 1, (size_t)trailer_bytes, f) != (size_t)trailer_bytes)
        rc = fail(SKM_ERR_INVALID, "skm_index_save: short write");
    cudaFreeHost(pinned);
    if (fclose(f) != 0 && !rc) rc = fail(SKM_ERR_INVALID, "skm_index_save: close failed");
    return rc;
}

SKM_API int skm_index_load(const char *path, int device, void *stream, skm_index **out, int64_t *trailer_offset,
                           int64_t *trailer_bytes)
{
    if (!out) return fail(SKM_ERR_INVALID, "skm_index_load: out is NULL");
    *out = nullptr;
    if (!path) return fail(SKM_ERR_INVALID, "skm_index_load: NULL path");
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_index_load: no such CUDA device (there is no CPU fallback)");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(SKM_ERR_INVALID, std::string("skm_index_load: cannot open ") + path);
    ImageHeader h{};
    const bool got = fread(&h, sizeof(h), 1, f) == 1;
    if (!got || memcmp(h.magic, IMAGE_MAGIC, 8) != 0) {
        fclose(f);
        return fail(SKM_ERR_INVALID, "skm_index_load: not a seekmer_b200 device image");
    }
    if (h.version != IMAGE_VERSION || h.kmer_size != (uint32_t)K || h.bucket_slots != (uint32_t)BUCKET_SLOTS
        || h.slot_bytes != sizeof(Slot) || h.contig_bytes != sizeof(ContigRec) || h.inline_targets != (uint32_t)INLINE_TARGETS) {
        fclose(f);
        return fail(SKM_ERR_INVALID, "skm_index_load: invalid index version.");  // _common.pyx:303-304
    }
    if (h.check != header_check(h) || h.n_slots <= 0 || (h.n_slots & (h.n_slots - 1)) != 0 || h.n_contigs <= 0
        || h.table_bytes != (int64_t)sizeof(Slot) * h.n_slots || h.hot_bytes <= 0 || h.off_seq2 < 0 || h.off_targets < h.off_seq2
        || h.off_targets > h.hot_bytes) {
        fclose(f);
        return fail(SKM_ERR_INVALID, "skm_index_load: corrupt header");
    }
    cudaError_t e = cudaSetDevice(device);
    cudaStream_t st = (cudaStream_t)stream;
    skm_index *ix = new skm_index();
    ix->device = device;
    ix->n_slots = h.n_slots;
    ix->n_kmers = h.n_kmers;
    ix->n_contigs = h.n_contigs;
    ix->n_bases = h.n_bases;
    ix->n_targets = h.n_targets;
    ix->n_transcripts = h.n_transcripts;
    ix->max_target_count = h.max_target_count;
    ix->hot_bytes = h.hot_bytes;
    ix->bytes = h.table_bytes + h.hot_bytes;
    void *pinned[2] = {nullptr, nullptr};
    if (e == cudaSuccess) e = cudaMalloc(&ix->table, (size_t)h.table_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&ix->hot, (size_t)h.hot_bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&pinned[0], IO_CHUNK);
    if (e == cudaSuccess) e = cudaMallocHost(&pinned[1], IO_CHUNK);
    int rc = e == cudaSuccess ? 0 : fail(e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA,
                                         std::string("skm_index_load: ") + cudaGetErrorString(e));
    if (!rc) rc = copy_in(f, ix->table, (size_t)h.table_bytes, pinned, st);
    if (!rc) rc = copy_in(f, ix->hot, (size_t)h.hot_bytes, pinned, st);
    cudaFreeHost(pinned[0]);
    cudaFreeHost(pinned[1]);
    fclose(f);
    if (rc) {
        skm_index_destroy(ix);
        return rc;
    }
    ix->contigs = reinterpret_cast<ContigRec *>(ix->hot);
    ix->seq2 = reinterpret_cast<uint32_t *>(ix->hot + h.off_seq2);
    ix->targets = reinterpret_cast<int32_t *>(ix->hot + h.off_targets);
    ix->d.table = ix->table;
    ix->d.bucket_mask = (uint64_t)ix->n_slots / BUCKET_SLOTS - 1;
    ix->d.contigs = ix->contigs;
    ix->d.seq2 = ix->seq2;
    ix->d.targets = ix->targets;
    ix->d.n_contigs = ix->n_contigs;
    ix->d.n_bases = ix->n_bases;
    ix->d.n_targets = ix->n_targets;
    rc = make_policies(ix, st);
    if (rc) {
        skm_index_destroy(ix);
        return rc;
    }
    if (trailer_offset) *trailer_offset = (int64_t)sizeof(ImageHeader) + h.table_bytes + h.hot_bytes;
    if (trailer_bytes) *trailer_bytes = h.trailer_bytes;
    skm::scratch_warm(ix->device);  // the EM scratch cache takes its first blocks now, not inside the first EM call
    *out = ix;
    return SKM_OK;
}
