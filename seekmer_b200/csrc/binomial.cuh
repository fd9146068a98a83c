// Exact binomial sampling on a counter-based random stream, for the O(classes) bootstrap
// resampler (em.cu: multinomial_tree_kernel).  Replaces, together with the tree around it, the
// draw of scipy.stats.multinomial(n, p).rvs() at infer.py:108-111.
//
// Two published algorithms, chosen like numpy's legacy generator does:
//   n * min(p, 1 - p) <  30   inversion by sequential search from 0 (Kachitvichyanukul & Schmeiser,
//                             "Binomial random variate generation", CACM 31(2), 1988, algorithm BINV)
//   otherwise                 BTPE, the triangle / parallelogram / exponential-tails rejection
//                             sampler of the same paper (expected ~1.2 rounds, independent of n)
// Host and device build (tests/binomial_host.cpp runs goodness-of-fit tests on the CPU).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SKM_HD __host__ __device__ __forceinline__
#define SKM_HD_MEMBER __host__ __device__ __forceinline__
#else
#define SKM_HD static inline
#define SKM_HD_MEMBER inline
#endif

namespace skm {

SKM_HD void philox4x32_hd(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Pairs of uniforms in the open interval (0, 1): round j of the stream (node, replicate) is
// Philox4x32-10(counter (node, replicate, j, tag), key seed); 53 bits each, centred in their cell.
struct UniformStream {
    uint32_t k0, k1, node, replicate, tag, round;
    SKM_HD_MEMBER void next(double &u, double &v)
    {
        uint32_t x[4];
        philox4x32_hd(node, replicate, round, tag, k0, k1, x);
        round += 1;
        const uint64_t a = ((uint64_t)x[1] << 32) | x[0], b = ((uint64_t)x[3] << 32) | x[2];
        u = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        v = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
};

// BINV: X = number of successes by walking the probability mass function up from 0
SKM_HD int64_t binomial_inversion(int64_t n, double p, UniformStream &rng)
{
    const double q = 1.0 - p;
    const double qn = exp((double)n * log1p(-p));
    const double np = (double)n * p;
    double bound = np + 10.0 * sqrt(np * q + 1.0);
    if (bound > (double)n) bound = (double)n;
    double u, unused;
    rng.next(u, unused);
    int64_t x = 0;
    double px = qn;
    while (u > px) {
        x += 1;
        if ((double)x > bound) {  // numerically out of mass: start over with a fresh uniform
            x = 0;
            px = qn;
            rng.next(u, unused);
        } else {
            u -= px;
            px = ((double)(n - x + 1) * p * px) / ((double)x * q);
        }
    }
    return x;
}

// BTPE for n * p >= 30, p <= 0.5
SKM_HD int64_t binomial_btpe(int64_t n, double r, UniformStream &rng)
{
    const double q = 1.0 - r;
    const double dn = (double)n;
    const double fm = dn * r + r;
    const int64_t m = (int64_t)floor(fm);
    const double nrq = dn * r * q;
    const double p1 = floor(2.195 * sqrt(nrq) - 4.6 * q) + 0.5;
    const double xm = (double)m + 0.5;
    const double xl = xm - p1, xr = xm + p1;
    const double c = 0.134 + 20.5 / (15.3 + (double)m);
    double a = (fm - xl) / (fm - xl * r);
    const double laml = a * (1.0 + a / 2.0);
    a = (xr - fm) / (xr * q);
    const double lamr = a * (1.0 + a / 2.0);
    const double p2 = p1 * (1.0 + 2.0 * c);
    const double p3 = p2 + c / laml;
    const double p4 = p3 + c / lamr;
    for (;;) {
        double u, v;
        rng.next(u, v);
        u *= p4;
        int64_t y;
        if (u <= p1) {  // triangle: accepted outright
            return (int64_t)floor(xm - p1 * v + u);
        }
        if (u <= p2) {  // parallelograms
            const double x = xl + (u - p1) / c;
            v = v * c + 1.0 - fabs((double)m - x + 0.5) / p1;
            if (v > 1.0) continue;
            y = (int64_t)floor(x);
        } else if (u <= p3) {  // left exponential tail
            y = (int64_t)floor(xl + log(v) / laml);
            if (y < 0) continue;
            v = v * (u - p2) * laml;
        } else {  // right exponential tail
            y = (int64_t)floor(xr - log(v) / lamr);
            if (y > n) continue;
            v = v * (u - p3) * lamr;
        }
        // acceptance: v against f(y) / f(m)
        const int64_t k = y > m ? y - m : m - y;
        if (k <= 20 || (double)k >= nrq / 2.0 - 1.0) {
            // explicit evaluation by the recurrence of the probability mass function
            const double s = r / q;
            const double aa = s * (dn + 1.0);
            double f = 1.0;
            if (m < y) {
                for (int64_t i = m + 1; i <= y; ++i) f *= aa / (double)i - s;
            } else if (m > y) {
                for (int64_t i = y + 1; i <= m; ++i) f /= aa / (double)i - s;
            }
            if (v > f) continue;
            return y;
        }
        // squeezing, then the Stirling-corrected log ratio
        const double dk = (double)k;
        const double rho = (dk / nrq) * ((dk * (dk / 3.0 + 0.625) + 0.16666666666666666) / nrq + 0.5);
        const double t = -dk * dk / (2.0 * nrq);
        const double lv = log(v);
        if (lv < t - rho) return y;
        if (lv > t + rho) continue;
        const double x1 = (double)y + 1.0, f1 = (double)m + 1.0, z = dn + 1.0 - (double)m, w = dn - (double)y + 1.0;
        const double x2 = x1 * x1, f2 = f1 * f1, z2 = z * z, w2 = w * w;
        const double bound = xm * log(f1 / x1) + (dn - (double)m + 0.5) * log(z / w)
                             + (double)(y - m) * log(w * r / (x1 * q))
                             + (13680.0 - (462.0 - (132.0 - (99.0 - 140.0 / f2) / f2) / f2) / f2) / f1 / 166320.0
                             + (13680.0 - (462.0 - (132.0 - (99.0 - 140.0 / z2) / z2) / z2) / z2) / z / 166320.0
                             + (13680.0 - (462.0 - (132.0 - (99.0 - 140.0 / x2) / x2) / x2) / x2) / x1 / 166320.0
                             + (13680.0 - (462.0 - (132.0 - (99.0 - 140.0 / w2) / w2) / w2) / w2) / w / 166320.0;
        if (lv > bound) continue;
        return y;
    }
}

// X ~ Binomial(n, p)
SKM_HD int64_t binomial_draw(int64_t n, double p, UniformStream &rng)
{
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    const double r = flip ? 1.0 - p : p;
    const int64_t x = (double)n * r < 30.0 ? binomial_inversion(n, r, rng) : binomial_btpe(n, r, rng);
    return flip ? n - x : x;
}

}  // namespace skm
