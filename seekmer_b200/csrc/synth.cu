// Device twin of seekmer_b200/synth.py::ReadSimulator (workload generation for benchmarks
// and tests; not part of the reference surface).  Integer-only, counter-based: unit i is a
// pure function of (seed, i), bit-identical to the numpy implementation.
#include "common.cuh"

namespace skm {

__device__ __forceinline__ void philox4x32_s(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
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

struct SynthArgs {
    const uint8_t *codes;
    const int64_t *tx_off;
    int64_t n_tx;
    const unsigned long long *cum;
    unsigned long long total;
    int32_t L, mu, sd, sub_thresh, n_thresh, random_pct;
    uint32_t seed_lo, seed_hi;
    int32_t paired;
    int64_t first_unit, n_units;
    uint8_t *out;
};

__global__ void synth_reads_kernel(const SynthArgs a)
{
    const int nb = (a.paired ? 2 : 1) * a.L;
    const int groups = (nb + 7) / 8;
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gid >= a.n_units * groups) return;
    const int64_t unit = gid / groups;
    const int g = (int)(gid % groups);
    const unsigned long long idx = (unsigned long long)(a.first_unit + unit);
    const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);

    uint32_t x[4], y[4], z[4];
    philox4x32_s(lo, hi, 0u, 0u, a.seed_lo, a.seed_hi, x);
    const unsigned long long u = __umul64hi(((unsigned long long)x[1] << 32) | x[0], a.total);
    int64_t lo_i = 0, hi_i = a.n_tx;  // first t with cum[t] > u
    while (lo_i < hi_i) {
        const int64_t m = (lo_i + hi_i) >> 1;
        if (a.cum[m] > u) hi_i = m;
        else lo_i = m + 1;
    }
    const int64_t t = lo_i;
    const bool swap = x[2] & 1u;
    const bool is_random = ((x[2] >> 8) % 100u) < (uint32_t)a.random_pct;
    philox4x32_s(lo, hi, 1u, 0u, a.seed_lo, a.seed_hi, y);
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (y[k] & 0xFFFFu) + (y[k] >> 16);
    long long frag = (long long)a.mu + (long long)((s * (unsigned long long)a.sd) / 53510ULL)
                     - (long long)((262140LL * a.sd) / 53510LL);
    const int64_t toff = a.tx_off[t];
    const long long tlen = a.tx_off[t + 1] - toff;
    if (frag < a.L) frag = a.L;
    if (frag > tlen) frag = tlen;
    if (!a.paired) frag = a.L;
    const unsigned long long span = (unsigned long long)(tlen - frag + 1);
    const int64_t start = (int64_t)(((unsigned long long)x[3] * span) >> 32);

    philox4x32_s(lo, hi, 2u + (uint32_t)g, 0u, a.seed_lo, a.seed_hi, z);
    uint8_t *dst = a.out + unit * (int64_t)nb + g * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int b = g * 8 + k;
        if (b >= nb) break;
        const uint32_t v = (k & 1) ? (z[k >> 1] >> 16) : (z[k >> 1] & 0xFFFFu);
        uint32_t base;
        if (is_random) {
            base = (v >> 2) & 3u;
        } else if (a.paired) {
            const bool second = b >= a.L;
            const int j = second ? b - a.L : b;
            // mate1 = first L bases; mate2 = reverse complement of the last L bases; swapped
            // pairs are emitted (mate2, mate1)
            const bool want_m2 = second != swap;
            base = want_m2 ? 3u - a.codes[toff + start + frag - 1 - j] : a.codes[toff + start + j];
        } else {
            base = swap ? 3u - a.codes[toff + start + a.L - 1 - b] : a.codes[toff + start + b];
        }
        if (v < (uint32_t)a.sub_thresh) base = (base + 1u + v % 3u) & 3u;
        uint8_t c = "ACGT"[base];
        if (v >= 65536u - (uint32_t)a.n_thresh) c = 'N';
        dst[k] = c;
    }
}

}  // namespace skm

using namespace skm;

SKM_API int skm_synth_reads(const uint8_t *tx_codes, const int64_t *tx_offsets, int64_t n_transcripts,
                            const uint64_t *cum_weights, uint64_t total_weight, int32_t read_len,
                            int32_t frag_mean, int32_t frag_sd, int32_t sub_thresh, int32_t n_thresh,
                            int32_t random_pct, uint64_t seed, int paired, int64_t first_unit,
                            int64_t n_units, uint8_t *bases, int device, void *stream)
{
    if (!tx_codes || !tx_offsets || !cum_weights || !bases)
        return fail(SKM_ERR_INVALID, "skm_synth_reads: NULL argument");
    if (n_units <= 0) return SKM_OK;
    if (skm_device_count() <= device || device < 0)
        return fail(SKM_ERR_CUDA, "skm_synth_reads: no such CUDA device");
    SKM_CUDA(cudaSetDevice(device));
    SynthArgs a{};
    a.codes = tx_codes;
    a.tx_off = tx_offsets;
    a.n_tx = n_transcripts;
    a.cum = reinterpret_cast<const unsigned long long *>(cum_weights);
    a.total = total_weight;
    a.L = read_len;
    a.mu = frag_mean;
    a.sd = frag_sd;
    a.sub_thresh = sub_thresh;
    a.n_thresh = n_thresh;
    a.random_pct = random_pct;
    a.seed_lo = (uint32_t)seed;
    a.seed_hi = (uint32_t)(seed >> 32);
    a.paired = paired ? 1 : 0;
    a.first_unit = first_unit;
    a.n_units = n_units;
    a.out = bases;
    const int nb = (paired ? 2 : 1) * read_len;
    const int64_t total = n_units * ((nb + 7) / 8);
    const int64_t blocks = (total + 255) / 256;
    if (blocks >= (1LL << 31)) return fail(SKM_ERR_INVALID, "skm_synth_reads: batch too large; split it");
    synth_reads_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    SKM_CUDA(cudaGetLastError());
    return SKM_OK;
}
