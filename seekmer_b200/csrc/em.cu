// EM abundance estimation, bootstrap resampling and effective lengths on the GPU (fp64).
//
//   skm_em                 infer.em          infer.py:133-168   (R replicates batched)
//   skm_multinomial        bootstrap resample infer.py:108-111
//   skm_effective_lengths  MapResult.effective_lengths  mapper.py:134-141
//
// Nothing here is a dense contraction: each EM iteration is two segmented reductions over
// the class x transcript incidence structure (CSR by class, CSC by transcript).  Replicates
// are laid out fastest ([row][replicate]) so that a warp reads one row's 32 replicates as one
// coalesced 256-byte access and every replicate's sum runs in the reference's order
// (numpy.bincount accumulates sequentially in nnz order).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>
#include <cooperative_groups.h>
#include <cuda.h>
#include <cub/cub.cuh>

#include "binomial.cuh"
#include "common.cuh"
#include "em_plan.cuh"

namespace skm {

constexpr int EM_BLOCK = 256;

// ---- structure setup -------------------------------------------------------------------
__global__ void expand_rows_kernel(const int64_t *__restrict__ ptr, int64_t n_rows,
                                   int32_t *__restrict__ row_of)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_rows) return;
    for (int64_t j = ptr[c]; j < ptr[c + 1]; ++j) row_of[j] = (int32_t)c;
}

__global__ void iota_kernel(int32_t *p, int64_t n)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

__global__ void histogram_kernel(const int32_t *__restrict__ keys, int64_t n, int64_t n_bins,
                                 unsigned long long *__restrict__ hist, unsigned int *bad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t k = keys[i];
    if (k < 0 || k >= n_bins) {
        atomicOr(bad, 1u);
        return;
    }
    atomicAdd(&hist[k + 1], 1ULL);
}

__global__ void gather_i32_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ idx,
                                  int64_t n, int32_t *__restrict__ dst)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// [R][N] (replicate-major, the ABI layout) <-> [N][R] (replicate-fastest, the kernel layout)
__global__ void transpose_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t rows,
                                 int64_t cols)
{
    // dst[c * rows + r] = src[r * cols + c]
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, c = i % cols;
    dst[c * rows + r] = src[i];
}

__global__ void sum_counts_kernel(const double *__restrict__ counts_cr, int64_t n_classes, int R,
                                  double *__restrict__ n_out)
{
    // one block per replicate; integer-valued doubles => order-independent exact sum
    const int r = blockIdx.x;
    double local = 0.0;
    for (int64_t c = threadIdx.x; c < n_classes; c += blockDim.x) local += counts_cr[c * R + r];
    __shared__ double sm[EM_BLOCK];
    sm[threadIdx.x] = local;
    __syncthreads();
    for (int s = EM_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) n_out[r] = sm[0];
}

// x[t][r] = x0[t] for every replicate
__global__ void broadcast_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n, int R)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n * R) dst[i] = src[i / R];
}

// Class counts [R][C] in the caller's class order (int64 or fp64) -> fp64 [C][R] in the plan's
// class order: row i of the result is the caller's class perm[i].
template <typename T>
__global__ void counts_to_plan_kernel(const T *__restrict__ src, const int32_t *__restrict__ perm,
                                      double *__restrict__ dst, int64_t n_classes, int R)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_classes * R) return;
    const int64_t c = i / R;
    const int r = (int)(i - c * R);
    dst[i] = (double)src[(int64_t)r * n_classes + perm[c]];
}

// dst[i] = src[i] + add: the CSR row pointer of one plan moved to its place behind the entries of
// the plans before it (skm_em_plans_run)
__global__ void offset_i64_kernel(const int64_t *__restrict__ src, int64_t n, int64_t add, int64_t *__restrict__ dst)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}

// ---- class order of a plan: by first transcript, so that neighbouring threads gather from
// ---- neighbouring places (see em_plan_adopt)
__global__ void class_sort_key_kernel(const int64_t *__restrict__ class_ptr, const int32_t *__restrict__ class_tx,
                                      int64_t n_classes, uint32_t *__restrict__ key, int32_t *__restrict__ idx)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_classes) return;
    const int64_t b = class_ptr[c];
    key[c] = class_ptr[c + 1] > b ? (uint32_t)class_tx[b] : 0xFFFFFFFFu;
    idx[c] = (int32_t)c;
}

__global__ void permuted_lens_kernel(const int64_t *__restrict__ class_ptr, const int32_t *__restrict__ perm,
                                     int64_t n_classes, int64_t *__restrict__ lens)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n_classes) return;
    lens[i] = i < n_classes ? class_ptr[perm[i] + 1] - class_ptr[perm[i]] : 0;
}

__global__ void permuted_ids_kernel(const int64_t *__restrict__ class_ptr, const int32_t *__restrict__ class_tx,
                                    const int32_t *__restrict__ perm, int64_t n_classes,
                                    const int64_t *__restrict__ new_ptr, int32_t *__restrict__ new_tx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_classes) return;
    const int64_t src = class_ptr[perm[i]], dst = new_ptr[i], len = class_ptr[perm[i] + 1] - src;
    for (int64_t k = 0; k < len; ++k) new_tx[dst + k] = class_tx[src + k];
}

// TPM post-processing of infer.py:127-129 for every replicate (one block each), x in [R][T]:
// x /= sum(x) / 1e6;  x[x < 0.001] = 0;  x /= sum(x) / 1e6
__device__ double block_sum(double v, double *sm)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double total = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += sm[w];
    return total;
}

__global__ void tpm_finish_kernel(double *__restrict__ x, int64_t n_tx)
{
    __shared__ double sm[32];
    double *row = x + (int64_t)blockIdx.x * n_tx;
    double local = 0.0;
    for (int64_t t = threadIdx.x; t < n_tx; t += blockDim.x) local += row[t];
    const double scale1 = __ddiv_rn(block_sum(local, sm), 1000000.0);
    local = 0.0;
    for (int64_t t = threadIdx.x; t < n_tx; t += blockDim.x) {
        double v = __ddiv_rn(row[t], scale1);
        if (v < 0.001) v = 0.0;  // NaN (an all-zero replicate) compares false and stays NaN, as in numpy
        row[t] = v;
        local += v;
    }
    const double scale2 = __ddiv_rn(block_sum(local, sm), 1000000.0);
    for (int64_t t = threadIdx.x; t < n_tx; t += blockDim.x) row[t] = __ddiv_rn(row[t], scale2);
}

// keep the columns map[0..R_new) of a [rows][R_old] matrix: dst is [rows][R_new]
__global__ void compact_cols_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t rows,
                                    int R_old, int R_new, const int32_t *__restrict__ map)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows * R_new) return;
    const int64_t row = i / R_new;
    const int j = (int)(i - row * R_new);
    dst[i] = src[row * R_old + map[j]];
}

// columns cols[j] of x [rows][R_old] -> rows origs[j] of out [.][rows] (the ABI layout)
__global__ void extract_cols_kernel(const double *__restrict__ src, int64_t rows, int R_old, int n_cols,
                                    const int32_t *__restrict__ cols, const int32_t *__restrict__ origs,
                                    double *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows * n_cols) return;
    const int64_t t = i / n_cols;
    const int j = (int)(i - t * n_cols);
    out[(int64_t)origs[j] * rows + t] = src[t * R_old + cols[j]];
}

struct EmState {
    const int64_t *class_ptr;   // CSR by class
    const int32_t *class_tx;
    const int64_t *tx_ptr;      // CSC by transcript (entries in nnz order)
    const int32_t *tx_class;
    const double *counts;       // [C][R] (R = live replicate columns)
    const double *eff_len;      // [T]
    const double *n;            // [R]
    double *inner;              // [C][R]
    unsigned long long *maxd;   // [R] bit pattern of the max relative change
    int32_t *active;            // [R]
    int32_t *iters;             // [R]
    int32_t *n_active;          // scalar
    int64_t n_classes, n_tx;
    int R;
    // many-samples variant (skm_em_samples): R counts samples, rows belong to one sample each
    const int32_t *class_sample;  // [C] sample of a class
    int64_t tx_per_sample;        // transcript rows are [sample][transcript]
    // transcript rows with more than HEAVY_ROW entries (one replicate / samples kernels): a whole
    // block sums each of them, the 8-lane groups skip them
    const int32_t *heavy_rows;
    int32_t n_heavy;
};

constexpr int HEAVY_ROW = 256;

__global__ void select_heavy_rows_kernel(const int64_t *__restrict__ tx_ptr, int64_t n_rows, int32_t *__restrict__ out,
                                         int32_t *__restrict__ count)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n_rows && tx_ptr[t + 1] - tx_ptr[t] > HEAVY_ROW) out[atomicAdd(count, 1)] = (int32_t)t;
}

// ---- E step: inner_c = (sum_{j in c} x[t_j]) / count_c  (infer.py:155-156,162-163) ----------
// R > 1: one thread per (class, replicate), replicates fastest: consecutive lanes read consecutive
// replicates of the same rows (coalesced), and every lane has work whatever R is (the state is
// compacted to the running replicates as they finish, so R shrinks during a run).
__global__ void em_class_kernel(const EmState s, const double *__restrict__ x)
{
    if (*s.n_active == 0) return;
    const int64_t total = s.n_classes * s.R;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e / s.R;
        const int r = (int)(e - c * s.R);
        if (!s.active[r]) continue;
        double sum = 0.0;
        const int64_t b = s.class_ptr[c], en = s.class_ptr[c + 1];
        for (int64_t j = b; j < en; ++j) sum = __dadd_rn(sum, x[(int64_t)s.class_tx[j] * s.R + r]);
        s.inner[e] = __ddiv_rn(sum, s.counts[e]);
    }
}

__device__ __forceinline__ void note_change(const EmState &s, int r, double xn, double xo)
{
    if (xn > 1e-8) {
        const double d = __ddiv_rn(fabs(__dsub_rn(xn, xo)), xn);
        atomicMax(&s.maxd[r], (unsigned long long)__double_as_longlong(d));
    }
}

// ---- M step: x_t = (sum_{j in t} x_t / inner_{c_j}) / l_t / n, NaN -> 0  (infer.py:157-159) ---
__global__ void em_tx_kernel(const EmState s, const double *__restrict__ x, double *__restrict__ xn_out)
{
    if (*s.n_active == 0) return;
    const int64_t total = s.n_tx * s.R;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / s.R;
        const int r = (int)(e - t * s.R);
        const double xt = x[e];
        if (!s.active[r]) {
            xn_out[e] = xt;
            continue;
        }
        double acc = 0.0;
        const int64_t b = s.tx_ptr[t], en = s.tx_ptr[t + 1];
        for (int64_t j = b; j < en; ++j)
            acc = __dadd_rn(acc, __ddiv_rn(xt, s.inner[(int64_t)s.tx_class[j] * s.R + r]));
        double v = __ddiv_rn(__ddiv_rn(acc, s.eff_len[t]), s.n[r]);
        if (v != v) v = 0.0;
        note_change(s, r, v, xt);
        xn_out[e] = v;
    }
}

// Loop condition of infer.py:160 evaluated per replicate after each update.
__global__ void em_decide_kernel(const EmState s)
{
    if (*s.n_active == 0) return;
    __shared__ int still;
    if (threadIdx.x == 0) still = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < s.R; r += blockDim.x) {
        if (s.active[r]) {
            s.iters[r] += 1;
            const double d = __longlong_as_double((long long)s.maxd[r]);
            if (d > 0.01) atomicAdd(&still, 1);
            else s.active[r] = 0;
            s.maxd[r] = 0ULL;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *s.n_active = still;
}

// ---- the fused iteration loop --------------------------------------------------------------
// One cooperative launch runs up to n_iters EM iterations with TWO grid barriers each: E step,
// barrier, M step (with the convergence reduction), barrier.  Nothing returns to the host
// between iterations; the whole structure (~46 MB at human scale) stays in L2 from the second
// iteration on.  The loop condition of infer.py:160 needs no barrier of its own: every block
// keeps the run / stop flag of every replicate (sample) in shared memory and updates it from the
// same maximum-change words after the M-step barrier, so all blocks decide alike; the two
// max-change buffers alternate, and block 0 clears the one of the previous iteration while
// nobody reads it.  Everything another SM may have written during this launch (x, inner, the
// max-change words) is read past L1 (ld.global.cg); the class structure, counts and lengths are
// read-only and may sit in L1.
namespace cg = cooperative_groups;
#ifndef SKM_EM_LOOP_THREADS
#define SKM_EM_LOOP_THREADS 1024
#endif
#ifndef SKM_EM_LOOP_BLOCKS
#define SKM_EM_LOOP_BLOCKS 2
#endif
constexpr int EM_LOOP_THREADS = SKM_EM_LOOP_THREADS;  // block size of the fused loop
constexpr int EM_LOOP_BLOCKS = SKM_EM_LOOP_BLOCKS;    // blocks per SM it is compiled (and launched) for

struct EmLoop {
    double *xa, *xb;      // ping-pong buffers; iteration 0 reads xa
    int32_t *executed;    // iterations this launch executed (the result sits in xa when even, xb when odd)
    int n_iters;
};

// SAMPLES = false: one replicate of one structure; true: samples laid end to end (skm_em_samples).
// s.maxd holds 2 x s.R words (zero on entry); s.active / s.iters / s.n_active are read at the
// start and written at the end.
template <bool SAMPLES>
__global__ void __launch_bounds__(EM_LOOP_THREADS, EM_LOOP_BLOCKS) em_loop_kernel(const EmState s, const EmLoop lp)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ uint32_t sm_run[];  // bit r: replicate / sample r still iterates
    __shared__ double sm_part[EM_LOOP_THREADS / 32];
    const int words = (s.R + 31) >> 5;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        uint32_t bits = 0;
        for (int b = 0; b < 32 && 32 * w + b < s.R; ++b) bits |= (s.active[32 * w + b] != 0 ? 1u : 0u) << b;
        sm_run[w] = bits;
    }
    __syncthreads();
    auto running = [&](int r) { return (sm_run[r >> 5] >> (r & 31)) & 1u; };
    bool any = false;
    for (int w = threadIdx.x; w < words; w += blockDim.x) any |= sm_run[w] != 0;
    any = __syncthreads_or(any);
    const double *cur = lp.xa;
    double *nxt = lp.xb;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    int it = 0, par = 0;
    for (; it < lp.n_iters && any; ++it, par ^= 1) {
        EmState sp = s;
        sp.maxd = s.maxd + (size_t)par * s.R;  // this iteration's max-change words
        // ---- E step: one thread per class
        for (int64_t c = gtid; c < s.n_classes; c += gsize) {
            if (SAMPLES && !running(__ldg(s.class_sample + c))) continue;
            double sum = 0.0;
            const int64_t b = __ldg(s.class_ptr + c), en = __ldg(s.class_ptr + c + 1);
            for (int64_t j = b; j < en; ++j) sum = __dadd_rn(sum, __ldcg(cur + __ldg(s.class_tx + j)));
            s.inner[c] = __ddiv_rn(sum, __ldg(s.counts + c));
        }
        grid.sync();
        // the words of the previous iteration have been read by every block before it got here
        if (blockIdx.x == 0)
            for (int r = threadIdx.x; r < s.R; r += blockDim.x) s.maxd[(size_t)(par ^ 1) * s.R + r] = 0ULL;
        // ---- M step, rows of up to HEAVY_ROW entries: 8 lanes per row, strided partial sums,
        // fixed-order shuffle tree.  A row whose x is exactly 0 stays 0 whatever its classes hold
        // (0 / inner is 0, or NaN -> 0), so its sum is not formed.
        const int sub = threadIdx.x & 7;
        for (int64_t g0 = (gtid >> 5) * 4; g0 < s.n_tx; g0 += gsize >> 3) {  // warp-uniform bound
            const int64_t g = g0 + ((threadIdx.x & 31) >> 3);
            const bool ok = g < s.n_tx;
            int sample = 0;
            bool live = ok, heavy = false;
            double acc = 0.0, xt = 0.0;
            if (ok) {
                if (SAMPLES) sample = (int)(g / s.tx_per_sample);
                live = running(sample) != 0;
                xt = __ldcg(cur + g);
                if (live) {
                    const int64_t b = __ldg(s.tx_ptr + g), en = __ldg(s.tx_ptr + g + 1);
                    heavy = en - b > HEAVY_ROW;
                    if (!heavy && xt != 0.0)
                        for (int64_t j = b + sub; j < en; j += 8)
                            acc = __dadd_rn(acc, __ddiv_rn(xt, __ldcg(s.inner + __ldg(s.tx_class + j))));
                }
            }
            acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
            acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
            acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
            if (ok && sub == 0 && !heavy) {
                if (!live) {
                    nxt[g] = xt;  // a finished sample is carried through the ping-pong unchanged
                } else {
                    double v = __ddiv_rn(__ddiv_rn(acc, __ldg(s.eff_len + g)), __ldg(s.n + sample));
                    if (v != v) v = 0.0;
                    note_change(sp, sample, v, xt);
                    nxt[g] = v;
                }
            }
        }
        // ---- M step, heavy rows (highly expressed transcripts sit in thousands of classes): one
        // block per row, strided partial sums, fixed-order tree over lanes then over warps
        for (int h = blockIdx.x; h < s.n_heavy; h += gridDim.x) {
            const int64_t g = __ldg(s.heavy_rows + h);
            const int sample = SAMPLES ? (int)(g / s.tx_per_sample) : 0;
            const bool live = running(sample) != 0;
            const double xt = __ldcg(cur + g);
            double acc = 0.0;
            if (live && xt != 0.0) {
                const int64_t b = __ldg(s.tx_ptr + g), en = __ldg(s.tx_ptr + g + 1);
                for (int64_t j = b + threadIdx.x; j < en; j += EM_LOOP_THREADS)
                    acc = __dadd_rn(acc, __ddiv_rn(xt, __ldcg(s.inner + __ldg(s.tx_class + j))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if ((threadIdx.x & 31) == 0) sm_part[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                if (!live) {
                    nxt[g] = xt;
                } else {
                    double total = 0.0;
                    for (int w = 0; w < EM_LOOP_THREADS / 32; ++w) total = __dadd_rn(total, sm_part[w]);
                    double v = __ddiv_rn(__ddiv_rn(total, __ldg(s.eff_len + g)), __ldg(s.n + sample));
                    if (v != v) v = 0.0;
                    note_change(sp, sample, v, xt);
                    nxt[g] = v;
                }
            }
            __syncthreads();
        }
        grid.sync();
        // ---- the loop condition (infer.py:160), in every block alike
        any = false;
        for (int w = threadIdx.x; w < words; w += blockDim.x) {
            uint32_t bits = sm_run[w];
            for (int b = 0; b < 32; ++b) {
                if (!((bits >> b) & 1u)) continue;
                const int r = 32 * w + b;
                if (blockIdx.x == 0) s.iters[r] += 1;
                const double d = __longlong_as_double((long long)__ldcg(sp.maxd + r));
                if (!(d > 0.01)) bits &= ~(1u << b);
            }
            sm_run[w] = bits;
            any |= bits != 0;
        }
        any = __syncthreads_or(any);
        const double *t0 = cur;
        cur = nxt;
        nxt = const_cast<double *>(t0);
    }
    if (blockIdx.x == 0) {
        __shared__ int still;
        if (threadIdx.x == 0) still = 0;
        __syncthreads();
        for (int r = threadIdx.x; r < s.R; r += blockDim.x) {
            const int on = (int)running(r);
            s.active[r] = on;
            if (on) atomicAdd(&still, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            *s.n_active = still;
            *lp.executed = it;
        }
    }
}

// ---- many samples with their own class structures (skm_em_samples) ---------------------------
// The samples' structures are laid end to end (classes, and transcript rows [sample][transcript],
// class_tx already holding the global row), so one launch serves every sample that still runs.
// Per row the arithmetic is that of the R == 1 kernels above, with the sample's own n, stop
// flag and change maximum: results are bit-identical to one skm_em call per sample.
__global__ void global_tx_kernel(const int32_t *__restrict__ class_tx, const int32_t *__restrict__ class_of,
                                 const int32_t *__restrict__ class_sample, int64_t nnz, int64_t tx_per_sample,
                                 int32_t *__restrict__ out, unsigned int *bad)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const int32_t t = class_tx[j];
    if (t < 0 || t >= tx_per_sample) {
        atomicOr(bad, 1u);
        out[j] = 0;
        return;
    }
    out[j] = (int32_t)((int64_t)class_sample[class_of[j]] * tx_per_sample + t);
}

__global__ void sum_counts_samples_kernel(const double *__restrict__ counts, const int64_t *__restrict__ sample_class_ptr,
                                          double *__restrict__ n_out)
{
    // one block per sample, the partition and tree of sum_counts_kernel over the sample's classes
    const int64_t first = sample_class_ptr[blockIdx.x], last = sample_class_ptr[blockIdx.x + 1];
    double local = 0.0;
    for (int64_t c = first + threadIdx.x; c < last; c += blockDim.x) local += counts[c];
    __shared__ double sm[EM_BLOCK];
    sm[threadIdx.x] = local;
    __syncthreads();
    for (int s = EM_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) n_out[blockIdx.x] = sm[0];
}

__global__ void fill_i32_kernel(int32_t *p, int64_t n, int32_t v)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- effective lengths -----------------------------------------------------------------------
__global__ void eff_len_kernel(const int64_t *__restrict__ fld, const double *__restrict__ lengths,
                               int64_t n, double *__restrict__ out)
{
    __shared__ double p[SKM_MAX_FRAGMENT_LENGTH];
    __shared__ long long total;
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int i = 0; i < SKM_MAX_FRAGMENT_LENGTH; ++i) s += fld[i];
        total = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SKM_MAX_FRAGMENT_LENGTH; i += blockDim.x)
        p[i] = __ddiv_rn((double)fld[i], (double)total);
    __syncthreads();
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double len = lengths[t];
    double acc = 0.0;
    for (int i = 0; i < SKM_MAX_FRAGMENT_LENGTH; ++i) {
        double v = __dsub_rn(len, (double)i);
        if (v < 1.0) v = 1.0;
        acc = __dadd_rn(acc, __dmul_rn(v, p[i]));  // no FMA contraction: numpy multiplies, then adds
    }
    out[t] = acc;
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// ---- bootstrap resampling --------------------------------------------------------------------
// bucket b (top 16 bits of the 64-bit uniform) starts its search at lo[b]
__global__ void multinomial_buckets_kernel(const unsigned long long *__restrict__ cum, int64_t n_classes,
                                           unsigned long long n, int bits, int32_t *__restrict__ lo)
{
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_buckets = 1LL << bits;
    if (b > n_buckets) return;
    if (b == n_buckets) {
        lo[b] = (int32_t)(n_classes - 1);
        return;
    }
    const unsigned long long u = __umul64hi((unsigned long long)b << (64 - bits), n);
    int64_t a = 0, z = n_classes;  // first index with cum > u
    while (a < z) {
        const int64_t m = (a + z) >> 1;
        if (cum[m] > u) z = m;
        else a = m + 1;
    }
    lo[b] = (int32_t)min(a, n_classes - 1);
}

__global__ void multinomial_kernel(const unsigned long long *__restrict__ cum, int64_t n_classes,
                                   unsigned long long n, int bits, const int32_t *__restrict__ lo,
                                   int64_t first_replicate, uint32_t seed_lo, uint32_t seed_hi,
                                   unsigned long long *__restrict__ out)
{
    const int64_t r = blockIdx.y;
    unsigned long long *dst = out + r * n_classes;
    const uint32_t rep = (uint32_t)(first_replicate + r);
    // one Philox block = two draws: words (x0, x1) -> draw 2j, (x2, x3) -> draw 2j + 1
    const unsigned long long n_blocks = (n + 1) >> 1;
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < n_blocks;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t x[4];
        philox4x32((uint32_t)j, (uint32_t)(j >> 32), rep, 0u, seed_lo, seed_hi, x);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (2 * j + h >= n) break;
            const unsigned long long w = ((unsigned long long)x[2 * h + 1] << 32) | x[2 * h];
            const unsigned long long u = __umul64hi(w, n);
            const uint32_t b = (uint32_t)(w >> (64 - bits));
            int64_t a = lo[b], z = (int64_t)lo[b + 1] + 1;
            if (z > n_classes) z = n_classes;
            while (a < z) {
                const int64_t m = (a + z) >> 1;
                if (cum[m] > u) z = m;
                else a = m + 1;
            }
            atomicAdd(&dst[a], 1ULL);
        }
    }
}

// ---- O(classes) resampling: a binary tree of binomial splits ------------------------------------
// Multinomial(n, count / n) factorises over a binary tree on the classes: the n reads of a node
// go left with probability (weight of the left half) / (weight of the node), i.e. the left count
// is Binomial(n_node, w_left / w_node), and so on down to single classes.  Every split is drawn
// exactly (binomial.cuh) from its own Philox stream (node, replicate), so the cost is one
// binomial draw per class and replicate whatever the read depth, replicates are reproducible and
// shardable by id, and the counts of a replicate sum to n by construction.
// Level `level` has 2^level nodes of span 2^(levels - level) classes (the class range is padded
// to a power of two with weight-0 classes); one thread per (replicate, node).
__global__ void multinomial_tree_kernel(const unsigned long long *__restrict__ cum, int64_t n_classes, int levels,
                                        int level, const int64_t *__restrict__ parent, int64_t *__restrict__ child,
                                        int64_t n_replicates, int64_t first_replicate, uint32_t k0, uint32_t k1)
{
    const int64_t nodes = 1LL << level;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= nodes * n_replicates) return;
    const int64_t r = idx >> level, i = idx & (nodes - 1);
    const int64_t span = 1LL << (levels - level);
    const int64_t lo = min(i * span, n_classes), mid = min(i * span + span / 2, n_classes), hi = min((i + 1) * span, n_classes);
    const unsigned long long w_lo = lo ? cum[lo - 1] : 0ULL, w_mid = mid ? cum[mid - 1] : 0ULL,
                             w_hi = hi ? cum[hi - 1] : 0ULL;
    const int64_t n_node = level == 0 ? (int64_t)cum[n_classes - 1] : parent[r * nodes + i];
    const unsigned long long w_left = w_mid - w_lo, w_node = w_hi - w_lo;
    int64_t left = 0;
    if (n_node > 0 && w_left > 0) {
        if (w_left == w_node) {
            left = n_node;
        } else {
            UniformStream rng{k0, k1, (uint32_t)(nodes + i), (uint32_t)(first_replicate + r), 0x54524545u, 0u};
            left = binomial_draw(n_node, (double)w_left / (double)w_node, rng);
        }
    }
    if (level == levels - 1) {  // the children are classes 2i and 2i + 1 of the output [replicate][class]
        int64_t *out = child + r * n_classes;
        if (2 * i < n_classes) out[2 * i] = left;
        if (2 * i + 1 < n_classes) out[2 * i + 1] = n_node - left;
    } else {
        *reinterpret_cast<longlong2 *>(child + (r * nodes + i) * 2) = make_longlong2(left, n_node - left);
    }
}

// Scratch memory: blocks from cudaMalloc are kept in a per-device cache when a call is done
// and handed out again to the next call (cudaMalloc / cudaFree cost milliseconds per GB-sized
// buffer and cudaFree synchronises the device; the stream-ordered pool was measured to stall
// for hundreds of ms on these sizes).  All users run on one stream at a time per device and
// synchronise it before returning, so a cached block is never still in use.
// Scratch blocks come from the virtual-memory-management API, with access for the OWNING device
// only.  Why not cudaMalloc: once peer access is on (NCCL turns it on for every NVLink peer), a
// cudaMalloc also maps the new block into every peer's address space, and that is what took the
// time of the first bootstrap call in a multi-GPU job (2 x B200, 1.1 GB of scratch for 50
// replicates: resample stage 157 ms, plan 187 ms; 13 ms and 4 ms with NCCL_P2P_DISABLE=1, i.e.
// with the same kernels and no peer mappings - profiles/r02c_alloc_peer_access.txt).  Nothing in
// this cache is ever read by a peer GPU.  The driver entry points are looked up at run time
// (cudaGetDriverEntryPoint): the library does not link against libcuda, so it still loads on a
// machine without a driver.  SKM_NO_VMM=1 goes back to cudaMalloc.
// 0 = cudaMalloc, 1 = VMM blocks with access for the owning device only.  The host layer switches it
// on when the process talks to peer GPUs (skm_scratch_local_only): with peer access on, cudaMalloc
// costs 100+ ms per GB; without, it is the steadier of the two (one GPU, EM + 100 bootstraps over
// ten fresh boxes: 180 - 220 ms with cudaMalloc, 150 - 410 ms with VMM blocks, whose first calls
// on a cold device sometimes take 50 ms each).  SKM_SCRATCH=vmm / malloc overrides.
static int g_scratch_local_only = 0;

struct Vmm {
    CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*address_free)(CUdeviceptr, size_t) = nullptr;
    CUresult (*granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
    Vmm()
    {
        if (getenv("SKM_NO_VMM")) return;
        if (const char *m = getenv("SKM_SCRATCH")) {
            if (!strcmp(m, "malloc")) return;
        }
        auto find = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult st = cudaDriverEntryPointSymbolNotFound;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess &&
                   st == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        ok = find("cuMemCreate", (void **)&create) && find("cuMemAddressReserve", (void **)&reserve) &&
             find("cuMemMap", (void **)&map) && find("cuMemSetAccess", (void **)&set_access) &&
             find("cuMemUnmap", (void **)&unmap) && find("cuMemRelease", (void **)&release) &&
             find("cuMemAddressFree", (void **)&address_free) &&
             find("cuMemGetAllocationGranularity", (void **)&granularity);
        if (!ok) cudaGetLastError();
    }
};
static Vmm &vmm_api()
{
    static Vmm v;  // resolved on first use, when the CUDA runtime is up
    return v;
}

struct BlockCache {
    struct Block {
        void *p;
        size_t bytes;
        bool own;  // a cudaMalloc of its own (can be given back to the driver), not a piece of a slab
    };
    static constexpr size_t SLAB = 256u << 20;  // small blocks are carved out of slabs: few cudaMalloc calls
    static constexpr size_t KEEP = 8ull << 30;  // own blocks kept idle per device; the largest go first beyond it
    std::mutex mu;
    std::vector<Block> free_blocks[64];
    std::vector<void *> slabs[64];
    char *slab_next[64] = {};
    size_t slab_left[64] = {};
    std::unordered_map<void *, size_t> mapped[64];  // blocks that came from the VMM API: mapped size
    // one block of device memory nobody else maps (mu held)
    cudaError_t raw_alloc(int d, size_t bytes, void **out)
    {
        Vmm &vmm = vmm_api();
        static const bool forced = getenv("SKM_SCRATCH") && !strcmp(getenv("SKM_SCRATCH"), "vmm");
        if (vmm.ok && (forced || g_scratch_local_only)) {
            CUmemAllocationProp prop = {};
            prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
            prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            prop.location.id = d;
            size_t gran = 0;
            if (vmm.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) == CUDA_SUCCESS && gran > 0) {
                const size_t size = (bytes + gran - 1) / gran * gran;
                CUmemGenericAllocationHandle h;
                const bool tr = getenv("SKM_TRACE_ALLOC") != nullptr;
                timespec ts0, ts1, ts2, ts3, ts4;
                if (tr) clock_gettime(CLOCK_MONOTONIC, &ts0);
                const CUresult made = vmm.create(&h, size, &prop, 0);
                if (tr) clock_gettime(CLOCK_MONOTONIC, &ts1);
                if (made == CUDA_ERROR_OUT_OF_MEMORY) return cudaErrorMemoryAllocation;
                if (made == CUDA_SUCCESS) {
                    CUdeviceptr va = 0;
                    bool good = vmm.reserve(&va, size, 0, 0, 0) == CUDA_SUCCESS;
                    if (tr) clock_gettime(CLOCK_MONOTONIC, &ts2);
                    bool is_mapped = false;
                    if (good) good = is_mapped = vmm.map(va, size, 0, h, 0) == CUDA_SUCCESS;
                    if (tr) clock_gettime(CLOCK_MONOTONIC, &ts3);
                    if (good) {
                        CUmemAccessDesc acc = {};
                        acc.location = prop.location;
                        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
                        good = vmm.set_access(va, size, &acc, 1) == CUDA_SUCCESS;
                    }
                    if (tr) {
                        clock_gettime(CLOCK_MONOTONIC, &ts4);
                        auto ms = [](const timespec &a, const timespec &b) {
                            return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6;
                        };
                        fprintf(stderr, "[skm alloc] %8.1f MB: create %.3f reserve %.3f map %.3f set_access %.3f ms\n",
                                size / 1048576.0, ms(ts0, ts1), ms(ts1, ts2), ms(ts2, ts3), ms(ts3, ts4));
                    }
                    vmm.release(h);  // the mapping keeps the memory until it is unmapped
                    if (good) {
                        *out = reinterpret_cast<void *>(va);
                        mapped[d][*out] = size;
                        return cudaSuccess;
                    }
                    if (is_mapped) vmm.unmap(va, size);
                    if (va) vmm.address_free(va, size);
                }
            }
        }
        return cudaMalloc(out, bytes);
    }
    void raw_free(int d, void *p)
    {
        auto it = mapped[d].find(p);
        if (it == mapped[d].end()) {
            cudaFree(p);
            return;
        }
        Vmm &vmm = vmm_api();
        // what cudaFree does: nothing in flight on the OWNING device may still use the block (the
        // calling thread may be working on another GPU)
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        cudaSetDevice(prev);
        vmm.unmap(reinterpret_cast<CUdeviceptr>(p), it->second);
        vmm.address_free(reinterpret_cast<CUdeviceptr>(p), it->second);
        mapped[d].erase(it);
    }
    // give idle own blocks back to the driver until at most `keep` bytes of them stay cached
    // (mu held; cudaFree synchronises the device, so this only runs on pressure or on request)
    size_t trim_locked(int d, size_t keep)
    {
        auto &v = free_blocks[d];
        size_t idle = 0, freed = 0;
        for (const Block &b : v)
            if (b.own) idle += b.bytes;
        while (idle > keep) {
            int big = -1;
            for (int i = 0; i < (int)v.size(); ++i)
                if (v[i].own && (big < 0 || v[i].bytes > v[big].bytes)) big = i;
            if (big < 0) break;
            raw_free(d, v[big].p);
            idle -= v[big].bytes;
            freed += v[big].bytes;
            v.erase(v.begin() + big);
        }
        return freed;
    }
    size_t trim(int device, size_t keep)
    {
        std::lock_guard<std::mutex> lock(mu);
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        const size_t freed = trim_locked(device & 63, keep);
        cudaSetDevice(prev);
        return freed;
    }
    cudaError_t take(int device, size_t bytes, void **out, size_t *got)
    {
        const int d = device & 63;
        bytes = (bytes + 255) & ~(size_t)255;
        std::lock_guard<std::mutex> lock(mu);
        auto &v = free_blocks[d];
        int best = -1;
        for (int i = 0; i < (int)v.size(); ++i)
            if (v[i].bytes >= bytes && v[i].bytes <= 2 * bytes + (1u << 20) && (best < 0 || v[i].bytes < v[best].bytes))
                best = i;
        if (best >= 0) {
            *out = v[best].p;
            *got = v[best].bytes;
            v.erase(v.begin() + best);
            return cudaSuccess;
        }
        *got = bytes;
        if (bytes <= SLAB / 4) {
            if (slab_left[d] < bytes) {
                void *slab = nullptr;
                cudaError_t e = raw_alloc(d, SLAB, &slab);
                if (e == cudaErrorMemoryAllocation && trim_locked(d, 0) > 0) {  // our own idle blocks first
                    cudaGetLastError();
                    e = raw_alloc(d, SLAB, &slab);
                }
                if (e != cudaSuccess) return e;
                slabs[d].push_back(slab);
                slab_next[d] = static_cast<char *>(slab);
                slab_left[d] = SLAB;
            }
            *out = slab_next[d];
            slab_next[d] += bytes;
            slab_left[d] -= bytes;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = raw_alloc(d, bytes, &p);
        if (e == cudaErrorMemoryAllocation && trim_locked(d, 0) > 0) {
            cudaGetLastError();
            e = raw_alloc(d, bytes, &p);
        }
        if (e != cudaSuccess) return e;
        *out = p;
        return cudaSuccess;
    }
    void give(int device, void *p, size_t bytes)
    {
        std::lock_guard<std::mutex> lock(mu);
        const int d = device & 63;
        free_blocks[d].push_back(Block{p, bytes, bytes > SLAB / 4});
        trim_locked(d, KEEP);
    }
};
static BlockCache g_blocks;

// Device memory that outlives a call (the arrays of an EM plan) from the same cache: making and
// dropping a plan per sample then costs no cudaMalloc / cudaFree (each of which synchronises the
// device: ~20 of them took 20+ ms per plan).
static std::mutex g_owned_mu;
static std::unordered_map<void *, size_t> g_owned[64];

cudaError_t dev_alloc(int device, size_t bytes, void **out)
{
    size_t got = 0;
    const cudaError_t e = g_blocks.take(device, std::max<size_t>(bytes, 256), out, &got);
    if (e != cudaSuccess) {
        *out = nullptr;
        return e;
    }
    std::lock_guard<std::mutex> lock(g_owned_mu);
    g_owned[device & 63][*out] = got;
    return cudaSuccess;
}

void dev_free(int device, void *p)
{
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_owned_mu);
        auto &m = g_owned[device & 63];
        auto it = m.find(p);
        if (it == m.end()) return;
        bytes = it->second;
        m.erase(it);
    }
    g_blocks.give(device, p, bytes);
}

void scratch_warm(int device)
{
    static std::mutex mu;
    static bool done[64] = {};
    {
        std::lock_guard<std::mutex> lock(mu);
        if (done[device & 63]) return;
        done[device & 63] = true;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    void *small = nullptr, *large = nullptr;
    if (dev_alloc(device, 1u << 20, &small) != cudaSuccess) cudaGetLastError();
    if (dev_alloc(device, 256u << 20, &large) != cudaSuccess) cudaGetLastError();
    dev_free(device, large);
    dev_free(device, small);
    cudaSetDevice(prev);
}

struct DeviceBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int device = 0;
    ~DeviceBuf() { release(); }
    void release()
    {
        if (p) g_blocks.give(device, p, bytes);
        p = nullptr;
    }
    cudaError_t alloc(size_t want, cudaStream_t = nullptr)
    {
        release();
        cudaGetDevice(&device);
        return g_blocks.take(device, std::max<size_t>(want, 256), &p, &bytes);
    }
    template <typename T>
    T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace skm

using namespace skm;

// SKM_TRACE=1: wall-clock stage times on stderr (each mark synchronises the stream)
struct Trace {
    bool on;
    cudaStream_t st;
    double t0;
    static double now()
    {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit Trace(cudaStream_t s) : on(getenv("SKM_TRACE") != nullptr), st(s), t0(0)
    {
        if (on) t0 = now();
    }
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(st);
        const double t = now();
        fprintf(stderr, "[skm trace] %-28s %9.3f ms\n", what, t - t0);
        t0 = t;
    }
};

// Large results into PAGEABLE host memory (a fresh numpy array): cudaMemcpy stages such a copy
// through the driver's own bounce buffer with one host thread, which then takes every first-touch
// page fault of the destination (160 MB of bootstrap results: 32-38 ms, 4.5 GB/s).  Here chunks
// land in two page-locked buffers at link speed and a few host threads copy each chunk on to the
// caller's buffer while the next chunk is in flight.  The stream is idle when this returns.
static std::mutex g_bounce_mu;
static unsigned char *g_bounce[2] = {nullptr, nullptr};
static cudaError_t copy_to_pageable(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    constexpr size_t CHUNK = 16u << 20;
    constexpr int THREADS = 4;
    if (bytes < 4 * CHUNK || getenv("SKM_PLAIN_D2H")) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    std::lock_guard<std::mutex> lock(g_bounce_mu);
    for (int i = 0; i < 2; ++i)
        if (!g_bounce[i] && cudaHostAlloc((void **)&g_bounce[i], CHUNK, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            g_bounce[i] = nullptr;
            const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
            return e != cudaSuccess ? e : cudaStreamSynchronize(st);
        }
    std::thread workers[2][THREADS];
    bool busy[2] = {false, false};
    auto join = [&](int b) {
        if (busy[b])
            for (auto &w : workers[b]) w.join();
        busy[b] = false;
    };
    cudaError_t err = cudaSuccess;
    int k = 0;
    for (size_t off = 0; off < bytes && err == cudaSuccess; off += CHUNK, ++k) {
        const int b = k & 1;
        const size_t n = std::min(CHUNK, bytes - off);
        join(b);  // the host copy out of this bounce buffer two chunks ago
        err = cudaMemcpyAsync(g_bounce[b], static_cast<const unsigned char *>(src) + off, n, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        if (err != cudaSuccess) break;
        const size_t piece = (n + THREADS - 1) / THREADS;
        for (int t = 0; t < THREADS; ++t) {
            const size_t lo = std::min(n, (size_t)t * piece), hi = std::min(n, lo + piece);
            unsigned char *to = static_cast<unsigned char *>(dst) + off + lo;
            const unsigned char *from = g_bounce[b] + lo;
            workers[b][t] = std::thread([to, from, lo, hi] { memcpy(to, from, hi - lo); });
        }
        busy[b] = true;
    }
    join(0);
    join(1);
    return err;
}

#define EM_TRY(expr)                                                                            \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(_e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA,          \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
    } while (0)

static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)std::max<int64_t>((n + per - 1) / per, 1); }

SKM_API int skm_scratch_local_only(int on)
{
    const int was = g_scratch_local_only;
    g_scratch_local_only = on ? 1 : 0;
    return was;
}

SKM_API int skm_release_cache(int device, int64_t *freed_bytes)
{
    if (freed_bytes) *freed_bytes = 0;
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_release_cache: no such CUDA device");
    const size_t freed = g_blocks.trim(device, 0);
    if (freed_bytes) *freed_bytes = (int64_t)freed;
    return SKM_OK;
}

SKM_API int skm_effective_lengths(const int64_t *fld, const double *lengths, int64_t n_transcripts,
                                  double *out, int buffers_on_device, int device, void *stream)
{
    if (!fld || !lengths || !out) return fail(SKM_ERR_INVALID, "skm_effective_lengths: NULL argument");
    if (n_transcripts <= 0) return SKM_OK;
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_effective_lengths: no such CUDA device (there is no CPU fallback)");
    EM_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    DeviceBuf b_fld, b_len, b_out;
    const int64_t *d_fld = fld;
    const double *d_len = lengths;
    double *d_out = out;
    if (!buffers_on_device) {
        EM_TRY(b_fld.alloc(sizeof(int64_t) * SKM_MAX_FRAGMENT_LENGTH));
        EM_TRY(b_len.alloc(sizeof(double) * (size_t)n_transcripts));
        EM_TRY(b_out.alloc(sizeof(double) * (size_t)n_transcripts));
        EM_TRY(cudaMemcpyAsync(b_fld.p, fld, sizeof(int64_t) * SKM_MAX_FRAGMENT_LENGTH, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_len.p, lengths, sizeof(double) * (size_t)n_transcripts, cudaMemcpyHostToDevice, st));
        d_fld = b_fld.as<int64_t>();
        d_len = b_len.as<double>();
        d_out = b_out.as<double>();
    }
    eff_len_kernel<<<blocks_for(n_transcripts, 128), 128, 0, st>>>(d_fld, d_len, n_transcripts, d_out);
    EM_TRY(cudaGetLastError());
    if (!buffers_on_device) {
        EM_TRY(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n_transcripts, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
    }
    return SKM_OK;
}

// ---- CSC by transcript: stable radix sort of (row, nnz index) ---------------------------------
// rows = transcripts (or sample x transcript rows); `d_row` holds the row of every nnz entry in CSR
// order.  Outputs (cudaMalloc, owned by the caller): tx_ptr[n_rows + 1], tx_class[nnz].
static int build_csc(const int64_t *d_ptr, const int32_t *d_row, int64_t C, int64_t nnz, int64_t n_rows,
                     cudaStream_t st, int64_t *tx_ptr, int32_t *tx_class, const char *who)
{
    DeviceBuf b_rowof, b_idx, b_keys_out, b_idx_out, b_hist, b_tmp, b_bad;
    EM_TRY(b_rowof.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_idx.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_keys_out.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_idx_out.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_hist.alloc(sizeof(unsigned long long) * (size_t)(n_rows + 1), st));
    EM_TRY(b_bad.alloc(sizeof(unsigned int), st));
    EM_TRY(cudaMemsetAsync(b_hist.p, 0, sizeof(unsigned long long) * (size_t)(n_rows + 1), st));
    EM_TRY(cudaMemsetAsync(b_bad.p, 0, sizeof(unsigned int), st));
    expand_rows_kernel<<<blocks_for(C, 256), 256, 0, st>>>(d_ptr, C, b_rowof.as<int32_t>());
    iota_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(b_idx.as<int32_t>(), nnz);
    histogram_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(d_row, nnz, n_rows, b_hist.as<unsigned long long>(),
                                                          b_bad.as<unsigned int>());
    {
        size_t tmp_sort = 0, tmp_scan = 0;
        int end_bit = 1;
        while ((1LL << end_bit) < n_rows) ++end_bit;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, d_row, b_keys_out.as<int32_t>(), b_idx.as<int32_t>(),
                                        b_idx_out.as<int32_t>(), (int)nnz, 0, end_bit, st);
        cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, b_hist.as<unsigned long long>(),
                                      reinterpret_cast<unsigned long long *>(tx_ptr), (int)(n_rows + 1), st);
        EM_TRY(b_tmp.alloc(std::max(tmp_sort, tmp_scan), st));
        size_t tmp = std::max(tmp_sort, tmp_scan);
        EM_TRY(cub::DeviceRadixSort::SortPairs(b_tmp.p, tmp, d_row, b_keys_out.as<int32_t>(), b_idx.as<int32_t>(),
                                               b_idx_out.as<int32_t>(), (int)nnz, 0, end_bit, st));
        tmp = std::max(tmp_sort, tmp_scan);
        EM_TRY(cub::DeviceScan::InclusiveSum(b_tmp.p, tmp, b_hist.as<unsigned long long>(),
                                             reinterpret_cast<unsigned long long *>(tx_ptr), (int)(n_rows + 1), st));
    }
    gather_i32_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(b_rowof.as<int32_t>(), b_idx_out.as<int32_t>(), nnz, tx_class);
    EM_TRY(cudaGetLastError());
    unsigned int bad = 0;
    EM_TRY(cudaMemcpyAsync(&bad, b_bad.p, sizeof(bad), cudaMemcpyDeviceToHost, st));
    EM_TRY(cudaStreamSynchronize(st));  // also: the scratch above goes back to the cache after this
    if (bad) return fail(SKM_ERR_INVALID, std::string(who) + ": transcript index out of range in class_tx");
    return SKM_OK;
}

// One cooperative launch of the fused iteration loop; *executed = iterations it ran.
static int launch_em_loop(bool samples, const EmState &s, double *xa, double *xb, int n_iters, int32_t *d_executed,
                          int device, cudaStream_t st)
{
    static std::mutex mu;
    static int blocks_per_sm[64][2];
    static int sms[64];
    const int mode = samples ? 1 : 0;
    const void *fn = samples ? (const void *)em_loop_kernel<true> : (const void *)em_loop_kernel<false>;
    const int d = device & 63;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (blocks_per_sm[d][mode] == 0) {
            int b = 0;
            EM_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, EM_LOOP_THREADS, 4096));
            if (b < 1) return fail(SKM_ERR_CUDA, "skm_em: the iteration kernel does not fit on this device");
            blocks_per_sm[d][mode] = std::min(b, EM_LOOP_BLOCKS);
            EM_TRY(cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, device));
        }
    }
    // all blocks must be resident (grid barrier): at most occupancy x SMs, no more than the rows need
    const int64_t rows = std::max(s.n_classes, s.n_tx * 8);
    const int64_t want = (rows + EM_LOOP_THREADS - 1) / EM_LOOP_THREADS;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)blocks_per_sm[d][mode] * sms[d]));
    EmState state = s;
    EmLoop lp{xa, xb, d_executed, n_iters};
    void *args[] = {&state, &lp};
    const size_t run_bytes = sizeof(uint32_t) * (size_t)((s.R + 31) / 32);
    if (run_bytes > 4096) return fail(SKM_ERR_INVALID, "skm_em: more than 32768 samples in one call");
    EM_TRY(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(EM_LOOP_THREADS), args, run_bytes, st));
    return SKM_OK;
}

// rows with more than HEAVY_ROW entries, in any order (each is summed by a block of its own)
static int select_heavy_rows(const int64_t *tx_ptr, int64_t n_rows, cudaStream_t st, int32_t **out, int32_t *n_out)
{
    *out = nullptr;
    *n_out = 0;
    int32_t *list = nullptr, *count = nullptr;
    int device = 0;
    EM_TRY(cudaGetDevice(&device));
    EM_TRY(dev_alloc(device, sizeof(int32_t) * (size_t)n_rows, (void **)&list));
    cudaError_t e = dev_alloc(device, sizeof(int32_t), (void **)&count);
    if (e == cudaSuccess) e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    if (e == cudaSuccess) {
        select_heavy_rows_kernel<<<blocks_for(n_rows, 256), 256, 0, st>>>(tx_ptr, n_rows, list, count);
        e = cudaMemcpyAsync(n_out, count, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(device, count);
    if (e != cudaSuccess) {
        dev_free(device, list);
        return fail(SKM_ERR_CUDA, std::string("skm_em: heavy rows: ") + cudaGetErrorString(e));
    }
    *out = list;
    return SKM_OK;
}

struct EmInputs {
    const skm_em_plan *plan;   // structure, both orders (device)
    const double *d_len;       // effective lengths [T]
    int R;
    int64_t max_iters;
    double *counts_cr;         // class counts in kernel layout [C][R], plan class order; scratch of the call
    const double *x_rt;        // initial guess [R][T] (ABI layout), or
    const double *x_t;         // ... one guess [T] shared by all replicates
};

// The EM proper on device-resident inputs; d_out [R][T], d_iters [R] (device).
static int em_core(const EmInputs &in, double *d_out, int32_t *d_iters, cudaStream_t st)
{
    const int64_t C = in.plan->C, T = in.plan->T;
    const int R = in.R;
    const int64_t max_iters = in.max_iters > 0 ? in.max_iters : 1000000;
    const int device = in.plan->device;
    Trace trace(st);
    // ---- state in kernel layout
    DeviceBuf b_xa, b_xb, b_inner, b_n, b_maxd, b_active, b_iters, b_nactive;
    EM_TRY(b_xa.alloc(sizeof(double) * (size_t)(T * R), st));
    EM_TRY(b_xb.alloc(sizeof(double) * (size_t)(T * R), st));
    EM_TRY(b_inner.alloc(sizeof(double) * (size_t)(C * R), st));
    EM_TRY(b_n.alloc(sizeof(double) * (size_t)R, st));
    EM_TRY(b_maxd.alloc(sizeof(unsigned long long) * 2 * (size_t)R, st));  // the fused loop alternates between two
    EM_TRY(b_active.alloc(sizeof(int32_t) * (size_t)R, st));
    EM_TRY(b_iters.alloc(sizeof(int32_t) * (size_t)R, st));
    EM_TRY(b_nactive.alloc(sizeof(int32_t) * 2, st));  // [0] = replicates still running, [1] = iterations of the last launch
    const double *d_cnt = in.counts_cr;
    if (in.x_t) broadcast_kernel<<<blocks_for(T * R, 256), 256, 0, st>>>(in.x_t, b_xa.as<double>(), T, R);
    else if (R > 1) transpose_kernel<<<blocks_for(T * R, 256), 256, 0, st>>>(in.x_rt, b_xa.as<double>(), R, T);
    else EM_TRY(cudaMemcpyAsync(b_xa.p, in.x_rt, sizeof(double) * (size_t)T, cudaMemcpyDeviceToDevice, st));
    sum_counts_kernel<<<R, EM_BLOCK, 0, st>>>(d_cnt, C, R, b_n.as<double>());
    EM_TRY(cudaMemsetAsync(b_maxd.p, 0, sizeof(unsigned long long) * 2 * (size_t)R, st));
    EM_TRY(cudaMemsetAsync(b_iters.p, 0, sizeof(int32_t) * (size_t)R, st));
    fill_i32_kernel<<<blocks_for(R, 256), 256, 0, st>>>(b_active.as<int32_t>(), R, 1);
    fill_i32_kernel<<<1, 32, 0, st>>>(b_nactive.as<int32_t>(), 1, R);

    EmState s{};
    s.class_ptr = in.plan->class_ptr;
    s.class_tx = in.plan->class_tx;
    s.tx_ptr = in.plan->tx_ptr;
    s.tx_class = in.plan->tx_class;
    s.counts = d_cnt;
    s.eff_len = in.d_len;
    s.n = b_n.as<double>();
    s.inner = b_inner.as<double>();
    s.maxd = b_maxd.as<unsigned long long>();
    s.active = b_active.as<int32_t>();
    s.iters = b_iters.as<int32_t>();
    s.n_active = b_nactive.as<int32_t>();
    s.n_classes = C;
    s.n_tx = T;
    s.R = R;
    s.heavy_rows = in.plan->heavy_rows;
    s.n_heavy = in.plan->n_heavy;
    int32_t *d_executed = b_nactive.as<int32_t>() + 1;

    trace.mark("em: state setup");
    double *cur = b_xa.as<double>(), *nxt = b_xb.as<double>();
    // One replicate: ONE cooperative launch runs the whole EM (the loop condition is evaluated on the
    // device, nothing returns to the host between iterations).  Many replicates are bandwidth-bound
    // (100 columns of x and inner do not fit L2) and stop at very different iteration counts
    // (bootstraps: mean ~40, max > 100): they run as full-occupancy E / M / decide launches enqueued
    // in groups of GROUP iterations - launches after the last replicate stopped are no-ops - and
    // whenever at most three quarters of the live columns are still running the finished ones are
    // written to the output and the state is compacted to the running columns: the work of an
    // iteration follows the replicates that still need it.
    const int GROUP = 8;
    int64_t done = 0;
    int32_t n_active = R;
    int Rc = R;                          // live columns
    std::vector<int32_t> orig((size_t)R);  // original replicate of each live column
    for (int r = 0; r < R; ++r) orig[(size_t)r] = r;
    std::vector<int32_t> final_iters((size_t)R, 0);
    std::vector<int32_t> h_active((size_t)R), h_iters((size_t)R), h_map;
    std::vector<int32_t> group_iters((size_t)R, 0);  // iteration counts of the live columns when a group starts
    std::vector<double> h_n((size_t)R);
    DeviceBuf b_map;
    EM_TRY(b_map.alloc(sizeof(int32_t) * 2 * (size_t)R, st));
    const bool may_compact = R > 1;  // the counts buffer is scratch of this call
    double *counts_buf = in.counts_cr, *inner_buf = b_inner.as<double>();
    while (n_active > 0 && done < max_iters) {
        const int64_t left = max_iters - done;
        if (R == 1) {
            int rc = launch_em_loop(false, s, cur, nxt, (int)std::min<int64_t>((int64_t)1 << 30, left), d_executed, device, st);
            if (rc) return rc;
            int32_t h2[2] = {0, 0};
            EM_TRY(cudaMemcpyAsync(h2, s.n_active, sizeof(h2), cudaMemcpyDeviceToHost, st));
            EM_TRY(cudaStreamSynchronize(st));
            n_active = h2[0];
            done += h2[1];
            if (h2[1] & 1) std::swap(cur, nxt);  // the last executed iteration wrote the other buffer
            if (h2[1] == 0) break;
            continue;
        }
        const int g = (int)std::min<int64_t>(GROUP, left);
        const unsigned grid_c = (unsigned)std::min<int64_t>((C * Rc + EM_BLOCK - 1) / EM_BLOCK, 1 << 20);
        const unsigned grid_t = (unsigned)std::min<int64_t>((T * Rc + EM_BLOCK - 1) / EM_BLOCK, 1 << 20);
        double *group_cur = cur, *group_nxt = nxt;
        for (int k = 0; k < g; ++k) {
            em_class_kernel<<<grid_c, EM_BLOCK, 0, st>>>(s, cur);
            em_tx_kernel<<<grid_t, EM_BLOCK, 0, st>>>(s, cur, nxt);
            em_decide_kernel<<<1, 128, 0, st>>>(s);
            std::swap(cur, nxt);
        }
        EM_TRY(cudaGetLastError());
        // how many iterations of the group really ran = the largest increase of a column's count
        EM_TRY(cudaMemcpyAsync(h_active.data(), s.iters, sizeof(int32_t) * (size_t)Rc, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaMemcpyAsync(&n_active, s.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
        int ran = 0;
        for (int c = 0; c < Rc; ++c) ran = std::max(ran, h_active[(size_t)c] - group_iters[(size_t)c]);
        for (int c = 0; c < Rc; ++c) group_iters[(size_t)c] = h_active[(size_t)c];
        done += ran;
        // the device stopped ping-ponging after `ran` iterations while the host kept swapping
        cur = (ran & 1) ? group_nxt : group_cur;
        nxt = (ran & 1) ? group_cur : group_nxt;
        if (ran == 0) break;
        if (!may_compact || n_active <= 0 || 4 * n_active > 3 * Rc || Rc <= 8 || done >= max_iters) continue;
        // ---- compact to the running columns ------------------------------------------------------
        EM_TRY(cudaMemcpyAsync(h_active.data(), s.active, sizeof(int32_t) * (size_t)Rc, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaMemcpyAsync(h_iters.data(), s.iters, sizeof(int32_t) * (size_t)Rc, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaMemcpyAsync(h_n.data(), s.n, sizeof(double) * (size_t)Rc, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
        std::vector<int32_t> keep, drop_cols, drop_orig, keep_orig;
        std::vector<int32_t> keep_iters;
        std::vector<double> keep_n;
        for (int c = 0; c < Rc; ++c) {
            if (h_active[(size_t)c]) {
                keep.push_back(c);
                keep_orig.push_back(orig[(size_t)c]);
                keep_iters.push_back(h_iters[(size_t)c]);
                keep_n.push_back(h_n[(size_t)c]);
            } else {
                drop_cols.push_back(c);
                drop_orig.push_back(orig[(size_t)c]);
                final_iters[(size_t)orig[(size_t)c]] = h_iters[(size_t)c];
            }
        }
        const int n_keep = (int)keep.size(), n_drop = (int)drop_cols.size();
        int32_t *d_map = b_map.as<int32_t>();
        // finished columns -> output rows
        EM_TRY(cudaMemcpyAsync(d_map, drop_cols.data(), sizeof(int32_t) * (size_t)n_drop, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(d_map + R, drop_orig.data(), sizeof(int32_t) * (size_t)n_drop, cudaMemcpyHostToDevice, st));
        extract_cols_kernel<<<blocks_for(T * n_drop, 256), 256, 0, st>>>(cur, T, Rc, n_drop, d_map, d_map + R, d_out);
        EM_TRY(cudaStreamSynchronize(st));  // the map buffer is reused just below
        // running columns: x (cur -> nxt, swap), counts (-> the inner buffer, swap), n, iters, active
        EM_TRY(cudaMemcpyAsync(d_map, keep.data(), sizeof(int32_t) * (size_t)n_keep, cudaMemcpyHostToDevice, st));
        compact_cols_kernel<<<blocks_for(T * n_keep, 256), 256, 0, st>>>(cur, nxt, T, Rc, n_keep, d_map);
        std::swap(cur, nxt);
        compact_cols_kernel<<<blocks_for(C * n_keep, 256), 256, 0, st>>>(counts_buf, inner_buf, C, Rc, n_keep, d_map);
        std::swap(counts_buf, inner_buf);
        s.counts = counts_buf;
        s.inner = inner_buf;
        EM_TRY(cudaMemcpyAsync(const_cast<double *>(s.n), keep_n.data(), sizeof(double) * (size_t)n_keep, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(s.iters, keep_iters.data(), sizeof(int32_t) * (size_t)n_keep, cudaMemcpyHostToDevice, st));
        fill_i32_kernel<<<blocks_for(n_keep, 256), 256, 0, st>>>(s.active, n_keep, 1);
        EM_TRY(cudaMemsetAsync(s.maxd, 0, sizeof(unsigned long long) * (size_t)n_keep, st));
        EM_TRY(cudaGetLastError());
        EM_TRY(cudaStreamSynchronize(st));  // host vectors above go out of scope
        orig = keep_orig;
        group_iters = keep_iters;
        Rc = n_keep;
        s.R = Rc;
    }
    // `cur` holds every live column's final x: columns that stopped earlier are copied through on
    // each later iteration
    {
        EM_TRY(cudaMemcpyAsync(h_iters.data(), s.iters, sizeof(int32_t) * (size_t)Rc, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
        for (int c = 0; c < Rc; ++c) final_iters[(size_t)orig[(size_t)c]] = h_iters[(size_t)c];
    }

    trace.mark("em: iterations");
    if (R > 1) {
        std::vector<int32_t> cols((size_t)Rc);
        for (int c = 0; c < Rc; ++c) cols[(size_t)c] = c;
        int32_t *d_map = b_map.as<int32_t>();
        EM_TRY(cudaMemcpyAsync(d_map, cols.data(), sizeof(int32_t) * (size_t)Rc, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(d_map + R, orig.data(), sizeof(int32_t) * (size_t)Rc, cudaMemcpyHostToDevice, st));
        extract_cols_kernel<<<blocks_for(T * Rc, 256), 256, 0, st>>>(cur, T, Rc, Rc, d_map, d_map + R, d_out);
        EM_TRY(cudaStreamSynchronize(st));
    } else {
        EM_TRY(cudaMemcpyAsync(d_out, cur, sizeof(double) * (size_t)T, cudaMemcpyDeviceToDevice, st));
    }
    EM_TRY(cudaGetLastError());
    if (d_iters)
        EM_TRY(cudaMemcpyAsync(d_iters, final_iters.data(), sizeof(int32_t) * (size_t)R, cudaMemcpyHostToDevice, st));
    EM_TRY(cudaStreamSynchronize(st));  // scratch buffers are released in stream order after this
    return SKM_OK;
}

// The EM's gathers (x of a class's transcripts, inner of a transcript's classes) are 8 useful
// bytes per 32-byte sector when classes come in first-seen order.  For the iterations the classes
// are renumbered by their first row (a stable sort): the classes neighbouring threads work on
// then share sectors, and a row's classes sit close together.  The caller's order stays on the
// outside: counts go through `perm` (new class i = the caller's class perm[i]) on their way in.
// Outputs into caller-provided device arrays perm[C], new_ptr[C + 1], new_rows[nnz].
static cudaError_t reorder_classes(const int64_t *class_ptr, const int32_t *class_rows, int64_t C, int64_t n_rows,
                                   cudaStream_t st, int32_t *perm, int64_t *new_ptr, int32_t *new_rows)
{
    DeviceBuf b_key, b_key2, b_idx, b_lens, b_tmp;
    cudaError_t e = b_key.alloc(sizeof(uint32_t) * (size_t)C, st);
    if (e == cudaSuccess) e = b_key2.alloc(sizeof(uint32_t) * (size_t)C, st);
    if (e == cudaSuccess) e = b_idx.alloc(sizeof(int32_t) * (size_t)C, st);
    if (e == cudaSuccess) e = b_lens.alloc(sizeof(int64_t) * (size_t)(C + 1), st);
    if (e != cudaSuccess) return e;
    size_t tmp_sort = 0, tmp_scan = 0;
    int end_bit = 1;
    while ((1LL << end_bit) < n_rows) ++end_bit;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, b_key.as<uint32_t>(), b_key2.as<uint32_t>(), b_idx.as<int32_t>(),
                                    perm, (int)C, 0, end_bit, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, b_lens.as<int64_t>(), new_ptr, (int)(C + 1), st);
    e = b_tmp.alloc(std::max(tmp_sort, tmp_scan), st);
    if (e != cudaSuccess) return e;
    class_sort_key_kernel<<<blocks_for(C, 256), 256, 0, st>>>(class_ptr, class_rows, C, b_key.as<uint32_t>(),
                                                             b_idx.as<int32_t>());
    size_t tb = std::max(tmp_sort, tmp_scan);
    e = cub::DeviceRadixSort::SortPairs(b_tmp.p, tb, b_key.as<uint32_t>(), b_key2.as<uint32_t>(), b_idx.as<int32_t>(), perm,
                                        (int)C, 0, end_bit, st);
    if (e != cudaSuccess) return e;
    permuted_lens_kernel<<<blocks_for(C + 1, 256), 256, 0, st>>>(class_ptr, perm, C, b_lens.as<int64_t>());
    tb = std::max(tmp_sort, tmp_scan);
    e = cub::DeviceScan::ExclusiveSum(b_tmp.p, tb, b_lens.as<int64_t>(), new_ptr, (int)(C + 1), st);
    if (e != cudaSuccess) return e;
    permuted_ids_kernel<<<blocks_for(C, 256), 256, 0, st>>>(class_ptr, class_rows, perm, C, new_ptr, new_rows);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the scratch above goes back to the cache
    return e;
}

// ---- plans ---------------------------------------------------------------------------------------
SKM_API void skm_em_plan_destroy(skm_em_plan *p)
{
    if (!p) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    cudaStreamSynchronize(nullptr);  // nothing of the plan is in flight on the default stream ...
    cudaDeviceSynchronize();         // ... nor on any other, before its blocks can be handed out again
    dev_free(p->device, p->class_ptr);
    dev_free(p->device, p->class_tx);
    dev_free(p->device, p->tx_ptr);
    dev_free(p->device, p->tx_class);
    dev_free(p->device, p->counts);
    dev_free(p->device, p->heavy_rows);
    dev_free(p->device, p->perm);
    cudaSetDevice(prev);
    delete p;
}

int skm::em_plan_adopt(int device, int64_t C, int64_t nnz, int64_t T, int64_t *class_ptr, int32_t *class_tx,
                       int64_t *counts, cudaStream_t st, skm_em_plan **out)
{
    skm_em_plan *p = new skm_em_plan();
    p->device = device;
    p->C = C;
    p->nnz = nnz;
    p->T = T;
    p->class_ptr = class_ptr;
    p->class_tx = class_tx;
    p->counts = counts;
    Trace trace(st);
    // classes renumbered by their first transcript for the iterations (reorder_classes)
    int64_t *s_ptr = nullptr;
    int32_t *s_tx = nullptr;
    cudaError_t e = dev_alloc(device, sizeof(int32_t) * (size_t)C, (void **)&p->perm);
    if (e == cudaSuccess) e = dev_alloc(device, sizeof(int64_t) * (size_t)(C + 1), (void **)&s_ptr);
    if (e == cudaSuccess) e = dev_alloc(device, sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1), (void **)&s_tx);
    if (e == cudaSuccess) e = dev_alloc(device, sizeof(int64_t) * (size_t)(T + 1), (void **)&p->tx_ptr);
    if (e == cudaSuccess) e = dev_alloc(device, sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1), (void **)&p->tx_class);
    if (e == cudaSuccess) e = reorder_classes(class_ptr, class_tx, C, T, st, p->perm, s_ptr, s_tx);
    if (e != cudaSuccess) {
        dev_free(device, s_ptr);
        dev_free(device, s_tx);
        skm_em_plan_destroy(p);
        return fail(e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA, std::string("skm_em_plan: ") + cudaGetErrorString(e));
    }
    dev_free(device, p->class_ptr);
    dev_free(device, p->class_tx);
    p->class_ptr = class_ptr = s_ptr;
    p->class_tx = class_tx = s_tx;
    trace.mark("em plan: class order");
    int rc = build_csc(class_ptr, class_tx, C, nnz, T, st, p->tx_ptr, p->tx_class, "skm_em_plan");
    if (rc == 0) rc = select_heavy_rows(p->tx_ptr, T, st, &p->heavy_rows, &p->n_heavy);
    trace.mark("em plan: CSC build");
    if (rc) {
        skm_em_plan_destroy(p);
        return rc;
    }
    *out = p;
    return SKM_OK;
}

static int check_em_shape(const char *who, int64_t n_classes, int64_t nnz, int64_t n_transcripts, int device)
{
    if (n_classes <= 0 || nnz <= 0 || n_transcripts <= 0) return fail(SKM_ERR_INVALID, std::string(who) + ": empty problem");
    if (nnz >= (1LL << 31) || n_classes >= (1LL << 31) || n_transcripts >= (1LL << 31))
        return fail(SKM_ERR_INVALID, std::string(who) + ": structure too large for int32 indices");
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, std::string(who) + ": no such CUDA device (there is no CPU fallback)");
    return SKM_OK;
}

SKM_API int skm_em_plan_create(const int64_t *class_ptr, const int32_t *class_tx, int64_t n_classes, int64_t nnz,
                               int64_t n_transcripts, const int64_t *counts, int buffers_on_device, int device,
                               void *stream, skm_em_plan **out)
{
    if (!out) return fail(SKM_ERR_INVALID, "skm_em_plan_create: out is NULL");
    *out = nullptr;
    if (!class_ptr || !class_tx) return fail(SKM_ERR_INVALID, "skm_em_plan_create: NULL argument");
    int rc = check_em_shape("skm_em_plan_create", n_classes, nnz, n_transcripts, device);
    if (rc) return rc;
    EM_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t *d_ptr = nullptr, *d_counts = nullptr;
    int32_t *d_tx = nullptr;
    const cudaMemcpyKind kind = buffers_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaError_t e = dev_alloc(device, sizeof(int64_t) * (size_t)(n_classes + 1), (void **)&d_ptr);
    if (e == cudaSuccess) e = dev_alloc(device, sizeof(int32_t) * (size_t)nnz, (void **)&d_tx);
    if (e == cudaSuccess && counts) e = dev_alloc(device, sizeof(int64_t) * (size_t)n_classes, (void **)&d_counts);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ptr, class_ptr, sizeof(int64_t) * (size_t)(n_classes + 1), kind, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tx, class_tx, sizeof(int32_t) * (size_t)nnz, kind, st);
    if (e == cudaSuccess && counts) e = cudaMemcpyAsync(d_counts, counts, sizeof(int64_t) * (size_t)n_classes, kind, st);
    if (e != cudaSuccess) {
        dev_free(device, d_ptr);
        dev_free(device, d_tx);
        dev_free(device, d_counts);
        return fail(e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA,
                    std::string("skm_em_plan_create: ") + cudaGetErrorString(e));
    }
    return em_plan_adopt(device, n_classes, nnz, n_transcripts, d_ptr, d_tx, d_counts, st, out);
}

SKM_API int skm_em_plan_info(const skm_em_plan *p, int64_t info[5])
{
    if (!p || !info) return fail(SKM_ERR_INVALID, "skm_em_plan_info: NULL argument");
    info[0] = p->C;
    info[1] = p->nnz;
    info[2] = p->T;
    info[3] = p->device;
    info[4] = p->counts != nullptr;
    return SKM_OK;
}

SKM_API int skm_em_plan_run(const skm_em_plan *p, const double *counts, const double *eff_len, const double *x0,
                            int64_t n_replicates, int64_t max_iters, double *out_x, int32_t *out_iters,
                            int buffers_on_device, void *stream)
{
    if (!p || !eff_len || !x0 || !out_x) return fail(SKM_ERR_INVALID, "skm_em: NULL argument");
    if (n_replicates <= 0) return fail(SKM_ERR_INVALID, "skm_em: empty problem");
    if (!counts && (!p->counts || n_replicates != 1))
        return fail(SKM_ERR_INVALID, "skm_em: counts may only be NULL for one replicate of a plan that owns its counts");
    EM_TRY(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int R = (int)n_replicates;
    const int64_t C = p->C, T = p->T;
    Trace trace(st);
    DeviceBuf b_cnt_in, b_cnt, b_len, b_x_in, b_out, b_iters;
    EmInputs in{p, eff_len, R, max_iters, nullptr, x0, nullptr};
    double *d_out = out_x;
    int32_t *d_iters = out_iters;
    // class counts: the caller's order [R][C] -> the plan's order and layout [C][R]
    EM_TRY(b_cnt.alloc(sizeof(double) * (size_t)(C * R), st));
    in.counts_cr = b_cnt.as<double>();
    if (!counts) {
        counts_to_plan_kernel<int64_t><<<blocks_for(C, 256), 256, 0, st>>>(p->counts, p->perm, b_cnt.as<double>(), C, 1);
    } else {
        const double *d_counts = counts;
        if (!buffers_on_device) {
            EM_TRY(b_cnt_in.alloc(sizeof(double) * (size_t)(C * R), st));
            EM_TRY(cudaMemcpyAsync(b_cnt_in.p, counts, sizeof(double) * (size_t)(C * R), cudaMemcpyHostToDevice, st));
            d_counts = b_cnt_in.as<double>();
        }
        counts_to_plan_kernel<double><<<blocks_for(C * R, 256), 256, 0, st>>>(d_counts, p->perm, b_cnt.as<double>(), C, R);
    }
    EM_TRY(cudaGetLastError());
    if (!buffers_on_device) {
        EM_TRY(b_len.alloc(sizeof(double) * (size_t)T, st));
        EM_TRY(b_x_in.alloc(sizeof(double) * (size_t)(T * R), st));
        EM_TRY(b_out.alloc(sizeof(double) * (size_t)(T * R), st));
        EM_TRY(b_iters.alloc(sizeof(int32_t) * (size_t)R, st));
        EM_TRY(cudaMemcpyAsync(b_len.p, eff_len, sizeof(double) * (size_t)T, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_x_in.p, x0, sizeof(double) * (size_t)(T * R), cudaMemcpyHostToDevice, st));
        in.d_len = b_len.as<double>();
        in.x_rt = b_x_in.as<double>();
        d_out = b_out.as<double>();
        d_iters = b_iters.as<int32_t>();
    }
    trace.mark("em: inputs to device");
    const int rc = em_core(in, d_out, d_iters, st);
    if (rc) return rc;
    if (!buffers_on_device) {
        EM_TRY(cudaMemcpyAsync(out_x, d_out, sizeof(double) * (size_t)(T * R), cudaMemcpyDeviceToHost, st));
        if (out_iters) EM_TRY(cudaMemcpyAsync(out_iters, d_iters, sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
    }
    trace.mark("em: results out");
    return SKM_OK;
}

SKM_API int skm_em(const int64_t *class_ptr, const int32_t *class_tx, int64_t n_classes, int64_t nnz,
                   const double *counts, const double *eff_len, int64_t n_transcripts, const double *x0,
                   int64_t n_replicates, int64_t max_iters, double *out_x, int32_t *out_iters,
                   int buffers_on_device, int device, void *stream)
{
    if (!class_ptr || !class_tx || !counts || !eff_len || !x0 || !out_x)
        return fail(SKM_ERR_INVALID, "skm_em: NULL argument");
    if (n_replicates <= 0) return fail(SKM_ERR_INVALID, "skm_em: empty problem");
    skm_em_plan *plan = nullptr;
    int rc = skm_em_plan_create(class_ptr, class_tx, n_classes, nnz, n_transcripts, nullptr, buffers_on_device, device,
                                stream, &plan);
    if (rc) return rc;
    rc = skm_em_plan_run(plan, counts, eff_len, x0, n_replicates, max_iters, out_x, out_iters, buffers_on_device, stream);
    skm_em_plan_destroy(plan);
    return rc;
}

SKM_API int skm_em_samples(const int64_t *class_ptr, const int32_t *class_tx, const int64_t *sample_class_ptr,
                           int64_t n_samples, int64_t n_classes, int64_t nnz, const double *counts,
                           const double *eff_len, int64_t n_transcripts, const double *x0, int64_t max_iters,
                           double *out_x, int32_t *out_iters, int buffers_on_device, int device, void *stream)
{
    if (!class_ptr || !class_tx || !sample_class_ptr || !counts || !eff_len || !x0 || !out_x)
        return fail(SKM_ERR_INVALID, "skm_em_samples: NULL argument");
    if (n_samples <= 0 || n_classes <= 0 || nnz <= 0 || n_transcripts <= 0)
        return fail(SKM_ERR_INVALID, "skm_em_samples: empty problem");
    const int64_t P = n_samples, C = n_classes, T = n_transcripts;
    if (nnz >= (1LL << 31) || C >= (1LL << 31) || T >= (1LL << 31) || P >= (1LL << 31) / T)
        return fail(SKM_ERR_INVALID, "skm_em_samples: structure too large for int32 indices");
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_em_samples: no such CUDA device (there is no CPU fallback)");
    EM_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = P * T;
    if (max_iters <= 0) max_iters = 1000000;

    // ---- inputs on the device --------------------------------------------------------------
    DeviceBuf b_ptr, b_tx, b_sptr, b_cnt, b_len, b_xa, b_xb, b_iters_out;
    std::vector<int64_t> h_sptr((size_t)P + 1);
    EM_TRY(b_xa.alloc(sizeof(double) * (size_t)rows, st));
    EM_TRY(b_xb.alloc(sizeof(double) * (size_t)rows, st));
    const int64_t *d_ptr = class_ptr, *d_sptr = sample_class_ptr;
    const int32_t *d_tx_local = class_tx;
    const double *d_cnt = counts, *d_len = eff_len;
    if (!buffers_on_device) {
        EM_TRY(b_ptr.alloc(sizeof(int64_t) * (size_t)(C + 1), st));
        EM_TRY(b_tx.alloc(sizeof(int32_t) * (size_t)nnz, st));
        EM_TRY(b_sptr.alloc(sizeof(int64_t) * (size_t)(P + 1), st));
        EM_TRY(b_cnt.alloc(sizeof(double) * (size_t)C, st));
        EM_TRY(b_len.alloc(sizeof(double) * (size_t)rows, st));
        EM_TRY(cudaMemcpyAsync(b_ptr.p, class_ptr, sizeof(int64_t) * (size_t)(C + 1), cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_tx.p, class_tx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_sptr.p, sample_class_ptr, sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_cnt.p, counts, sizeof(double) * (size_t)C, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_len.p, eff_len, sizeof(double) * (size_t)rows, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_xa.p, x0, sizeof(double) * (size_t)rows, cudaMemcpyHostToDevice, st));
        d_ptr = b_ptr.as<int64_t>();
        d_tx_local = b_tx.as<int32_t>();
        d_sptr = b_sptr.as<int64_t>();
        d_cnt = b_cnt.as<double>();
        d_len = b_len.as<double>();
        std::copy(sample_class_ptr, sample_class_ptr + P + 1, h_sptr.begin());
    } else {
        EM_TRY(cudaMemcpyAsync(b_xa.p, x0, sizeof(double) * (size_t)rows, cudaMemcpyDeviceToDevice, st));
        EM_TRY(cudaMemcpyAsync(h_sptr.data(), sample_class_ptr, sizeof(int64_t) * (size_t)(P + 1),
                               cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
    }
    if (h_sptr[0] != 0 || h_sptr[(size_t)P] != C)
        return fail(SKM_ERR_INVALID, "skm_em_samples: sample_class_ptr must run from 0 to n_classes");
    for (int64_t p = 0; p < P; ++p)
        if (h_sptr[(size_t)p + 1] < h_sptr[(size_t)p])
            return fail(SKM_ERR_INVALID, "skm_em_samples: sample_class_ptr is not monotone");

    // ---- global rows [sample][transcript], CSC by row ------------------------------------------
    DeviceBuf b_csample, b_rowof, b_gtx, b_txclass, b_txptr, b_bad;
    EM_TRY(b_csample.alloc(sizeof(int32_t) * (size_t)C, st));
    EM_TRY(b_rowof.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_gtx.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_txclass.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_txptr.alloc(sizeof(int64_t) * (size_t)(rows + 1), st));
    EM_TRY(b_bad.alloc(sizeof(unsigned int), st));
    EM_TRY(cudaMemsetAsync(b_bad.p, 0, sizeof(unsigned int), st));
    expand_rows_kernel<<<blocks_for(P, 256), 256, 0, st>>>(d_sptr, P, b_csample.as<int32_t>());
    expand_rows_kernel<<<blocks_for(C, 256), 256, 0, st>>>(d_ptr, C, b_rowof.as<int32_t>());
    global_tx_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(d_tx_local, b_rowof.as<int32_t>(), b_csample.as<int32_t>(),
                                                          nnz, T, b_gtx.as<int32_t>(), b_bad.as<unsigned int>());
    EM_TRY(cudaGetLastError());
    {
        unsigned int bad = 0;
        EM_TRY(cudaMemcpyAsync(&bad, b_bad.p, sizeof(bad), cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
        if (bad) return fail(SKM_ERR_INVALID, "skm_em_samples: transcript index out of range in class_tx");
    }
    b_rowof.release();
    // classes renumbered by their first row for the iterations, as in a plan (samples stay contiguous:
    // rows are sample-major; the order inside a sample is the order a plan of that sample has,
    // so results stay bit-identical to one skm_em call per sample)
    DeviceBuf b_perm, b_sptr2, b_gtx2, b_cnt2;
    EM_TRY(b_perm.alloc(sizeof(int32_t) * (size_t)C, st));
    EM_TRY(b_sptr2.alloc(sizeof(int64_t) * (size_t)(C + 1), st));
    EM_TRY(b_gtx2.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_cnt2.alloc(sizeof(double) * (size_t)C, st));
    EM_TRY(reorder_classes(d_ptr, b_gtx.as<int32_t>(), C, rows, st, b_perm.as<int32_t>(), b_sptr2.as<int64_t>(),
                           b_gtx2.as<int32_t>()));
    counts_to_plan_kernel<double><<<blocks_for(C, 256), 256, 0, st>>>(d_cnt, b_perm.as<int32_t>(), b_cnt2.as<double>(), C, 1);
    EM_TRY(cudaGetLastError());
    d_ptr = b_sptr2.as<int64_t>();
    d_cnt = b_cnt2.as<double>();
    b_gtx.release();
    int rc = build_csc(d_ptr, b_gtx2.as<int32_t>(), C, nnz, rows, st, b_txptr.as<int64_t>(), b_txclass.as<int32_t>(),
                       "skm_em_samples");
    if (rc) return rc;

    // ---- per-sample state ------------------------------------------------------------------
    DeviceBuf b_inner, b_n, b_maxd, b_active, b_iters, b_nactive;
    EM_TRY(b_inner.alloc(sizeof(double) * (size_t)C, st));
    EM_TRY(b_n.alloc(sizeof(double) * (size_t)P, st));
    EM_TRY(b_maxd.alloc(sizeof(unsigned long long) * 2 * (size_t)P, st));
    EM_TRY(b_active.alloc(sizeof(int32_t) * (size_t)P, st));
    EM_TRY(b_iters.alloc(sizeof(int32_t) * (size_t)P, st));
    EM_TRY(b_nactive.alloc(sizeof(int32_t) * 2, st));
    sum_counts_samples_kernel<<<(unsigned)P, EM_BLOCK, 0, st>>>(d_cnt, d_sptr, b_n.as<double>());
    EM_TRY(cudaMemsetAsync(b_maxd.p, 0, sizeof(unsigned long long) * 2 * (size_t)P, st));
    EM_TRY(cudaMemsetAsync(b_iters.p, 0, sizeof(int32_t) * (size_t)P, st));
    fill_i32_kernel<<<blocks_for(P, 256), 256, 0, st>>>(b_active.as<int32_t>(), P, 1);
    fill_i32_kernel<<<1, 32, 0, st>>>(b_nactive.as<int32_t>(), 1, (int32_t)P);

    EmState s{};
    s.class_ptr = d_ptr;
    s.class_tx = b_gtx2.as<int32_t>();
    s.tx_ptr = b_txptr.as<int64_t>();
    s.tx_class = b_txclass.as<int32_t>();
    s.counts = d_cnt;
    s.eff_len = d_len;
    s.n = b_n.as<double>();
    s.inner = b_inner.as<double>();
    s.maxd = b_maxd.as<unsigned long long>();
    s.active = b_active.as<int32_t>();
    s.iters = b_iters.as<int32_t>();
    s.n_active = b_nactive.as<int32_t>();
    s.n_classes = C;
    s.n_tx = rows;
    s.R = (int)P;
    s.class_sample = b_csample.as<int32_t>();
    s.tx_per_sample = T;
    int32_t *heavy = nullptr, n_heavy = 0;
    rc = select_heavy_rows(b_txptr.as<int64_t>(), rows, st, &heavy, &n_heavy);
    if (rc) return rc;
    struct FreeOnExit {
        void *p;
        int device;
        ~FreeOnExit() { dev_free(device, p); }
    } free_heavy{heavy, device};
    s.heavy_rows = heavy;
    s.n_heavy = n_heavy;

    // ---- ONE launch: every sample iterates until its own stop condition holds ----------------------
    double *cur = b_xa.as<double>(), *nxt = b_xb.as<double>();
    int64_t done = 0;
    int32_t h2[2] = {(int32_t)P, 0};
    while (h2[0] > 0 && done < max_iters) {
        rc = launch_em_loop(true, s, cur, nxt, (int)std::min<int64_t>(max_iters - done, (int64_t)1 << 30), b_nactive.as<int32_t>() + 1,
                            device, st);
        if (rc) return rc;
        EM_TRY(cudaMemcpyAsync(h2, s.n_active, sizeof(h2), cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
        done += h2[1];
        if (h2[1] & 1) std::swap(cur, nxt);
        if (h2[1] == 0) break;
    }
    EM_TRY(cudaMemcpyAsync(out_x, cur, sizeof(double) * (size_t)rows,
                           buffers_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (out_iters)
        EM_TRY(cudaMemcpyAsync(out_iters, s.iters, sizeof(int32_t) * (size_t)P,
                               buffers_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    EM_TRY(cudaStreamSynchronize(st));  // scratch buffers are released after this
    return SKM_OK;
}

// Many samples, each with its own plan (made from its mapper, device to device), in ONE set of
// launches: the plans' structures are laid end to end on the device (plan class order, so the
// class order skm_em_samples gives every sample is the one its plan has) and handed to
// skm_em_samples - no host CSR, nothing but the lengths and first guesses cross PCIe.
SKM_API int skm_em_plans_run(const skm_em_plan *const *plans, int64_t n_plans, const double *eff_len, const double *x0,
                             int64_t max_iters, double *out_x, int32_t *out_iters, int buffers_on_device, void *stream)
{
    if (!plans || !eff_len || !x0 || !out_x) return fail(SKM_ERR_INVALID, "skm_em_plans_run: NULL argument");
    if (n_plans <= 0) return fail(SKM_ERR_INVALID, "skm_em_plans_run: empty problem");
    const int64_t P = n_plans;
    int64_t C = 0, nnz = 0;
    for (int64_t k = 0; k < P; ++k) {
        const skm_em_plan *p = plans[k];
        if (!p) return fail(SKM_ERR_INVALID, "skm_em_plans_run: NULL plan");
        if (p->device != plans[0]->device || p->T != plans[0]->T)
            return fail(SKM_ERR_INVALID, "skm_em_plans_run: the plans must sit on one device and share the transcript set");
        if (!p->counts) return fail(SKM_ERR_INVALID, "skm_em_plans_run: every plan must own its class counts");
        if (p->C <= 0 || p->nnz <= 0) return fail(SKM_ERR_INVALID, "skm_em_plans_run: empty plan");
        C += p->C;
        nnz += p->nnz;
    }
    const int device = plans[0]->device;
    const int64_t T = plans[0]->T;
    EM_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    DeviceBuf b_ptr, b_tx, b_cnt, b_sptr, b_len, b_x, b_out, b_iters;
    EM_TRY(b_ptr.alloc(sizeof(int64_t) * (size_t)(C + 1), st));
    EM_TRY(b_tx.alloc(sizeof(int32_t) * (size_t)nnz, st));
    EM_TRY(b_cnt.alloc(sizeof(double) * (size_t)C, st));
    EM_TRY(b_sptr.alloc(sizeof(int64_t) * (size_t)(P + 1), st));
    std::vector<int64_t> h_sptr((size_t)P + 1, 0);
    int64_t c_off = 0, e_off = 0;
    for (int64_t k = 0; k < P; ++k) {
        const skm_em_plan *p = plans[k];
        // C + 1 entries each: the closing entry of a plan is the first entry of the next
        offset_i64_kernel<<<blocks_for(p->C + 1, 256), 256, 0, st>>>(p->class_ptr, p->C + 1, e_off, b_ptr.as<int64_t>() + c_off);
        EM_TRY(cudaMemcpyAsync(b_tx.as<int32_t>() + e_off, p->class_tx, sizeof(int32_t) * (size_t)p->nnz,
                               cudaMemcpyDeviceToDevice, st));
        counts_to_plan_kernel<int64_t><<<blocks_for(p->C, 256), 256, 0, st>>>(p->counts, p->perm, b_cnt.as<double>() + c_off, p->C, 1);
        c_off += p->C;
        e_off += p->nnz;
        h_sptr[(size_t)k + 1] = c_off;
    }
    EM_TRY(cudaGetLastError());
    EM_TRY(cudaMemcpyAsync(b_sptr.p, h_sptr.data(), sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
    const double *d_len = eff_len, *d_x = x0;
    double *d_out = out_x;
    int32_t *d_iters = out_iters;
    EM_TRY(b_iters.alloc(sizeof(int32_t) * (size_t)P, st));
    if (!d_iters || !buffers_on_device) d_iters = b_iters.as<int32_t>();
    if (!buffers_on_device) {
        EM_TRY(b_len.alloc(sizeof(double) * (size_t)(P * T), st));
        EM_TRY(b_x.alloc(sizeof(double) * (size_t)(P * T), st));
        EM_TRY(b_out.alloc(sizeof(double) * (size_t)(P * T), st));
        EM_TRY(cudaMemcpyAsync(b_len.p, eff_len, sizeof(double) * (size_t)(P * T), cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_x.p, x0, sizeof(double) * (size_t)(P * T), cudaMemcpyHostToDevice, st));
        d_len = b_len.as<double>();
        d_x = b_x.as<double>();
        d_out = b_out.as<double>();
    }
    EM_TRY(cudaStreamSynchronize(st));  // h_sptr is read by the copy above
    const int rc = skm_em_samples(b_ptr.as<int64_t>(), b_tx.as<int32_t>(), b_sptr.as<int64_t>(), P, C, nnz, b_cnt.as<double>(),
                                  d_len, T, d_x, max_iters, d_out, d_iters, 1, device, stream);
    if (rc) return rc;
    if (!buffers_on_device) {
        EM_TRY(cudaMemcpyAsync(out_x, d_out, sizeof(double) * (size_t)(P * T), cudaMemcpyDeviceToHost, st));
        if (out_iters) EM_TRY(cudaMemcpyAsync(out_iters, d_iters, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, st));
    }
    EM_TRY(cudaStreamSynchronize(st));
    return SKM_OK;
}

// Resample n = sum(counts) reads with replacement, n_replicates times, on device buffers.
static int multinomial_tree(const unsigned long long *d_cum, int64_t n_classes, int64_t n_replicates,
                            int64_t first_replicate, uint64_t seed, int64_t *d_out, cudaStream_t st)
{
    int levels = 0;
    while ((1LL << levels) < n_classes) ++levels;
    if (levels == 0) {  // one class takes every read
        EM_TRY(cudaMemcpyAsync(d_out, d_cum, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        for (int64_t r = 1; r < n_replicates; ++r)
            EM_TRY(cudaMemcpyAsync(d_out + r, d_cum, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        return SKM_OK;
    }
    DeviceBuf b_a, b_b;
    if (levels > 1) {
        EM_TRY(b_a.alloc(sizeof(int64_t) * (size_t)(n_replicates << (levels - 1)), st));
        EM_TRY(b_b.alloc(sizeof(int64_t) * (size_t)(n_replicates << (levels - 1)), st));
    }
    int64_t *parent = nullptr, *child = b_a.as<int64_t>();
    for (int level = 0; level < levels; ++level) {
        int64_t *dst = level == levels - 1 ? d_out : child;
        const int64_t threads = n_replicates << level;
        multinomial_tree_kernel<<<blocks_for(threads, 128), 128, 0, st>>>(d_cum, n_classes, levels, level, parent, dst,
                                                                       n_replicates, first_replicate, (uint32_t)seed,
                                                                       (uint32_t)(seed >> 32));
        parent = child;
        child = child == b_a.as<int64_t>() ? b_b.as<int64_t>() : b_a.as<int64_t>();
    }
    EM_TRY(cudaGetLastError());
    EM_TRY(cudaStreamSynchronize(st));  // the level buffers go back to the cache
    return SKM_OK;
}

static int multinomial_core(const int64_t *d_counts, int64_t n_classes, int64_t n_replicates, int64_t first_replicate,
                            uint64_t seed, int method, int64_t *d_out, int device, cudaStream_t st)
{
    if (method != SKM_RESAMPLE_DRAWS && method != SKM_RESAMPLE_TREE)
        return fail(SKM_ERR_INVALID, "skm_multinomial: unknown resampling method");
    DeviceBuf b_cum, b_lo, b_tmp;
    EM_TRY(b_cum.alloc(sizeof(unsigned long long) * (size_t)n_classes, st));
    // about two buckets per class: the search that follows the bucket lookup is then 0-1 steps
    int bits = 16;
    while (bits < 22 && (1LL << bits) < 2 * n_classes) ++bits;
    EM_TRY(b_lo.alloc(sizeof(int32_t) * ((size_t)(1LL << bits) + 1), st));
    size_t tmp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp, reinterpret_cast<const unsigned long long *>(d_counts),
                                  b_cum.as<unsigned long long>(), (int)n_classes, st);
    EM_TRY(b_tmp.alloc(tmp, st));
    EM_TRY(cub::DeviceScan::InclusiveSum(b_tmp.p, tmp, reinterpret_cast<const unsigned long long *>(d_counts),
                                         b_cum.as<unsigned long long>(), (int)n_classes, st));
    if (method == SKM_RESAMPLE_TREE)
        return multinomial_tree(b_cum.as<unsigned long long>(), n_classes, n_replicates, first_replicate, seed, d_out, st);
    unsigned long long n = 0;
    EM_TRY(cudaMemcpyAsync(&n, b_cum.as<unsigned long long>() + (n_classes - 1), sizeof(n), cudaMemcpyDeviceToHost, st));
    EM_TRY(cudaStreamSynchronize(st));
    EM_TRY(cudaMemsetAsync(d_out, 0, sizeof(int64_t) * (size_t)(n_classes * n_replicates), st));
    if (n > 0) {
        multinomial_buckets_kernel<<<(unsigned)(((1LL << bits) + 256) / 256), 256, 0, st>>>(
            b_cum.as<unsigned long long>(), n_classes, n, bits, b_lo.as<int32_t>());
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const unsigned gx = (unsigned)std::min<unsigned long long>((n / 2 + 256) / 256, (unsigned long long)sms * 8);
        const dim3 grid(gx, (unsigned)n_replicates);
        multinomial_kernel<<<grid, 256, 0, st>>>(b_cum.as<unsigned long long>(), n_classes, n, bits, b_lo.as<int32_t>(),
                                                first_replicate, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                reinterpret_cast<unsigned long long *>(d_out));
    }
    EM_TRY(cudaGetLastError());
    EM_TRY(cudaStreamSynchronize(st));
    return SKM_OK;
}

SKM_API int skm_multinomial(const int64_t *counts, int64_t n_classes, int64_t n_replicates,
                            int64_t first_replicate, uint64_t seed, int method, int64_t *out, int buffers_on_device,
                            int device, void *stream)
{
    if (!counts || !out) return fail(SKM_ERR_INVALID, "skm_multinomial: NULL argument");
    if (n_classes <= 0 || n_replicates <= 0) return fail(SKM_ERR_INVALID, "skm_multinomial: empty problem");
    if (n_classes >= (1LL << 31)) return fail(SKM_ERR_INVALID, "skm_multinomial: too many classes");
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_multinomial: no such CUDA device (there is no CPU fallback)");
    EM_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    DeviceBuf b_counts, b_out;
    const int64_t *d_counts = counts;
    int64_t *d_out = out;
    if (!buffers_on_device) {
        EM_TRY(b_counts.alloc(sizeof(int64_t) * (size_t)n_classes, st));
        EM_TRY(b_out.alloc(sizeof(int64_t) * (size_t)(n_classes * n_replicates), st));
        EM_TRY(cudaMemcpyAsync(b_counts.p, counts, sizeof(int64_t) * (size_t)n_classes, cudaMemcpyHostToDevice, st));
        d_counts = b_counts.as<int64_t>();
        d_out = b_out.as<int64_t>();
    }
    const int rc = multinomial_core(d_counts, n_classes, n_replicates, first_replicate, seed, method, d_out, device, st);
    if (rc) return rc;
    if (!buffers_on_device) {
        EM_TRY(cudaMemcpyAsync(out, d_out, sizeof(int64_t) * (size_t)(n_classes * n_replicates), cudaMemcpyDeviceToHost, st));
        EM_TRY(cudaStreamSynchronize(st));
    }
    return SKM_OK;
}

SKM_API int skm_em_plan_bootstrap(const skm_em_plan *p, const int64_t *counts, const double *eff_len, const double *x0,
                                  int64_t n_replicates, int64_t first_replicate, uint64_t seed, int method,
                                  int64_t max_iters, int tpm, double *out_x, int32_t *out_iters, int buffers_on_device,
                                  void *stream)
{
    if (!p || !eff_len || !x0 || !out_x) return fail(SKM_ERR_INVALID, "skm_em_bootstrap: NULL argument");
    if (n_replicates <= 0) return fail(SKM_ERR_INVALID, "skm_em_bootstrap: empty problem");
    if (!counts && !p->counts) return fail(SKM_ERR_INVALID, "skm_em_bootstrap: the plan holds no class counts");
    EM_TRY(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int R = (int)n_replicates;
    const int64_t C = p->C, T = p->T;

    DeviceBuf b_counts, b_len, b_x, b_out, b_iters, b_draws, b_cr;
    EmInputs in{p, eff_len, R, max_iters, nullptr, nullptr, x0};
    const int64_t *d_counts = counts ? counts : p->counts;
    double *d_out = out_x;
    int32_t *d_iters = out_iters;
    if (!buffers_on_device) {
        if (counts) {
            EM_TRY(b_counts.alloc(sizeof(int64_t) * (size_t)C, st));
            EM_TRY(cudaMemcpyAsync(b_counts.p, counts, sizeof(int64_t) * (size_t)C, cudaMemcpyHostToDevice, st));
            d_counts = b_counts.as<int64_t>();
        }
        EM_TRY(b_len.alloc(sizeof(double) * (size_t)T, st));
        EM_TRY(b_x.alloc(sizeof(double) * (size_t)T, st));
        EM_TRY(b_out.alloc(sizeof(double) * (size_t)(T * R), st));
        EM_TRY(b_iters.alloc(sizeof(int32_t) * (size_t)R, st));
        EM_TRY(cudaMemcpyAsync(b_len.p, eff_len, sizeof(double) * (size_t)T, cudaMemcpyHostToDevice, st));
        EM_TRY(cudaMemcpyAsync(b_x.p, x0, sizeof(double) * (size_t)T, cudaMemcpyHostToDevice, st));
        in.d_len = b_len.as<double>();
        in.x_t = b_x.as<double>();
        d_out = b_out.as<double>();
        d_iters = b_iters.as<int32_t>();
    }
    Trace trace(st);
    trace.mark("bootstrap: inputs to device");
    // resample on the device, then straight into the EM's [class][replicate] fp64 layout
    EM_TRY(b_draws.alloc(sizeof(int64_t) * (size_t)(C * R), st));
    int rc = multinomial_core(d_counts, C, R, first_replicate, seed, method, b_draws.as<int64_t>(), p->device, st);
    if (rc) return rc;
    EM_TRY(b_cr.alloc(sizeof(double) * (size_t)(C * R), st));
    counts_to_plan_kernel<int64_t><<<blocks_for(C * R, 256), 256, 0, st>>>(b_draws.as<int64_t>(), p->perm, b_cr.as<double>(), C, R);
    EM_TRY(cudaGetLastError());
    EM_TRY(cudaStreamSynchronize(st));
    b_draws.release();
    trace.mark("bootstrap: resample");
    in.counts_cr = b_cr.as<double>();
    rc = em_core(in, d_out, d_iters, st);
    if (rc) return rc;
    trace.mark("bootstrap: em_core");
    if (tpm) {
        tpm_finish_kernel<<<R, 1024, 0, st>>>(d_out, T);
        EM_TRY(cudaGetLastError());
    }
    trace.mark("bootstrap: tpm");
    if (!buffers_on_device) {
        if (out_iters) EM_TRY(cudaMemcpyAsync(out_iters, d_iters, sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, st));
        EM_TRY(copy_to_pageable(out_x, d_out, sizeof(double) * (size_t)(T * R), st));
    }
    EM_TRY(cudaStreamSynchronize(st));
    trace.mark("bootstrap: results to host");
    return SKM_OK;
}

SKM_API int skm_em_bootstrap(const int64_t *class_ptr, const int32_t *class_tx, int64_t n_classes, int64_t nnz,
                             const int64_t *counts, const double *eff_len, int64_t n_transcripts, const double *x0,
                             int64_t n_replicates, int64_t first_replicate, uint64_t seed, int method,
                             int64_t max_iters, int tpm, double *out_x, int32_t *out_iters, int buffers_on_device,
                             int device, void *stream)
{
    if (!class_ptr || !class_tx || !counts || !eff_len || !x0 || !out_x)
        return fail(SKM_ERR_INVALID, "skm_em_bootstrap: NULL argument");
    if (n_replicates <= 0) return fail(SKM_ERR_INVALID, "skm_em_bootstrap: empty problem");
    skm_em_plan *plan = nullptr;
    int rc = skm_em_plan_create(class_ptr, class_tx, n_classes, nnz, n_transcripts, nullptr, buffers_on_device, device,
                                stream, &plan);
    if (rc) return rc;
    rc = skm_em_plan_bootstrap(plan, counts, eff_len, x0, n_replicates, first_replicate, seed, method, max_iters, tpm,
                               out_x, out_iters, buffers_on_device, stream);
    skm_em_plan_destroy(plan);
    return rc;
}
