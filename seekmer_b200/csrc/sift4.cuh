// Bounded SIFT4 edge check of the read mapper on register-resident windows.
//
// Reference: sift4_align_left (_mapper.pyx:404-445), sift4_align_right (:452-493),
// _match_base (:500-501).  Both reference routines compare 8 contig bases with the read
// around `offset`, tolerate up to 4 edits with cursor offsets of at most 1, and return a
// shift or 0x7FFF.
//
// One routine serves both directions.  Write the right-hand routine with cursors
// r (contig window, 0..7) and q (read, relative to `offset`, 0..8).  Substituting
//     r' = 7 - r,   q' = (offset + 7) - query_cursor
// into the left-hand routine turns every statement of it into the corresponding statement of
// the right-hand one (min <-> max, >= 0 <-> < 8, the two probe branches, both return
// expressions), with exactly two differences:
//   * the start: the left routine begins one base misaligned (`query_cursor -= 1`, :408), i.e.
//     at q' = 1 instead of q' = 0;
//   * the read bound of the first probe: `query_cursor - i >= 0` (:422) becomes
//     q' + i < offset + 8, where the right routine has q + i < length - offset (:470).
// So: mirror the two windows for the left direction, pick (q0, limit) accordingly, run the
// same loop.  The cursors never leave the three diagonals q - r in {-1, 0, +1} (a re-alignment
// returns to the main diagonal, each probe moves by one), so all base comparisons are
// precomputed as three 8-lane match masks with a handful of SWAR operations.
//
// Window frame ("unified frame"): contig base r sits at bits [15-2r : 14-2r] of `ref16`;
// read base q (q = 0..8) sits at bits [17-2q : 16-2q] of `q18`; `w9` bit (8 - q) is set when
// read base q is a wildcard (any byte other than upper-case ACGT, _mapper.pyx:501).
#pragma once

#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace skm {

constexpr int SIFT4_INVALID_SHIFT = 0x7FFF;  // _mapper.pyx:28
constexpr int SIFT4_MAX_DISTANCE = 4;        // _mapper.pyx:26

// reverse the order of n 2-bit fields held in the low 2n bits
__host__ __device__ __forceinline__ uint32_t reverse_pairs(uint32_t x, int n)
{
    uint32_t r = x;
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 4) & 0x0F0F0F0Fu) | ((r & 0x0F0F0F0Fu) << 4);
    r = ((r >> 8) & 0x00FF00FFu) | ((r & 0x00FF00FFu) << 8);
    r = (r >> 16) | (r << 16);
    return r >> (32 - 2 * n);
}

// reverse the order of n bits held in the low n bits
__host__ __device__ __forceinline__ uint32_t reverse_bits(uint32_t x, int n)
{
    uint32_t r = x;
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 4) & 0x0F0F0F0Fu) | ((r & 0x0F0F0F0Fu) << 4);
    r = ((r >> 8) & 0x00FF00FFu) | ((r & 0x00FF00FFu) << 8);
    r = (r >> 16) | (r << 16);
    return r >> (32 - n);
}

// spread the low 9 bits of w to the even bit positions (bit k -> bit 2k)
__host__ __device__ __forceinline__ uint32_t spread9(uint32_t w)
{
    w &= 0x1FFu;
    w = (w | (w << 8)) & 0x00FF00FFu;
    w = (w | (w << 4)) & 0x0F0F0F0Fu;
    w = (w | (w << 2)) & 0x33333333u;
    w = (w | (w << 1)) & 0x55555555u;
    return w;
}

// Match masks for the three diagonals.  Bit (14 - 2r) of the result for diagonal d tells
// whether contig base r matches read base r + d (equal codes, or wildcard read base).
__host__ __device__ __forceinline__ uint32_t sift4_diagonal(uint32_t ref16, uint32_t q18, uint32_t wspread,
                                                            int d)
{
    const int sh = 2 - 2 * d;  // aligns read base r+d with contig base r
    const uint32_t x = ((q18 >> sh) ^ ref16) & 0xFFFFu;
    const uint32_t eq = ~(x | (x >> 1)) & 0x5555u;
    const uint32_t wd = (wspread >> sh) & 0x5555u;
    return eq | wd;
}

// The unified loop.  q0 = 0 (right) or 1 (left); limit = number of read bases available from
// the window start in the unified frame (right: length - offset; left: offset + 8).
__host__ __device__ inline int sift4_unified(uint32_t ref16, uint32_t q18, uint32_t w9, int q0, int limit)
{
    const uint32_t ws = spread9(w9);  // wildcard of read base q at bit 16 - 2q
    const uint32_t m_lo = sift4_diagonal(ref16, q18, ws, -1);
    const uint32_t m_0 = sift4_diagonal(ref16, q18, ws, 0);
    const uint32_t m_hi = sift4_diagonal(ref16, q18, ws, 1);
    // match(r, q) for q - r in {-1, 0, 1}
#define SKM_M(r, q) ((((q) == (r) ? m_0 : ((q) > (r) ? m_hi : m_lo)) >> (14 - 2 * (r))) & 1u)
    // The common case: all 8 bases match on the main diagonal.  From q0 = 0 the loop walks it
    // straight to r = q = 8: shift 0.  From q0 = 1 the walk starts on the diagonal above; at its
    // first mismatch (r = k, q = k + 1) the cursors re-align to r = q = k + 1, the probe of that
    // cell matches at no cost (:427-436) and the main diagonal leads to r = q = 8: shift 0 -
    // unless the upper diagonal matches for r = 0..6, when q reaches 8 with r = 7: shift 1.
    if (m_0 == 0x5555u) return (q0 != 0 && (m_hi & 0x5554u) == 0x5554u) ? 1 : 0;
    int r = 0, q = q0, distance = 0;
    while (r < 8 && q < 8) {
        if (SKM_M(r, q)) {
            r += 1;
            q += 1;
            continue;
        }
        if (r != q) {
            r = q > r ? q : r;
            q = r;
        }
        for (int i = 0; i < 2; ++i) {
            if (q + i < 9 && q + i < limit && SKM_M(r, q + i)) {
                distance += i - 1;
                q += i - 1;
                r -= 1;
                break;
            }
            if (r + i < 8 && SKM_M(r + i, q)) {
                distance += i - 1;
                q -= 1;
                r += i - 1;
                break;
            }
        }
        distance += 1;
        q += 1;
        r += 1;
        if (distance > SIFT4_MAX_DISTANCE) return SIFT4_INVALID_SHIFT;
    }
#undef SKM_M
    if (r < 8) return 8 - r;
    if (q < 8) return q - 8;
    return 0;
}

}  // namespace skm
