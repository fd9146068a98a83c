// Host build of seekmer_b200/csrc/sift4.cuh for CPU unit tests (tests/test_sift4.py).
// Windows are assembled per base from ASCII with the frame definitions of the header.
#include "../seekmer_b200/csrc/sift4.cuh"

static inline uint32_t code_of(char b)
{
    const char u = b & 0xDF;
    return u == 'T' ? 3u : u == 'G' ? 2u : u == 'C' ? 1u : 0u;
}
static inline uint32_t wild_of(char b) { return !(b == 'A' || b == 'C' || b == 'G' || b == 'T'); }

extern "C" int sift4_window(const char *ref8, const char *query, int len, int offset, int left)
{
    uint32_t ref16 = 0;
    for (int r = 0; r < 8; ++r) ref16 |= code_of(ref8[r]) << (14 - 2 * r);
    uint32_t q18 = 0, w9 = 0;
    for (int q = 0; q < 9; ++q) {
        const int abs = left ? offset + 7 - q : offset + q;
        if (abs < 0 || abs >= len) continue;
        q18 |= code_of(query[abs]) << (16 - 2 * q);
        w9 |= wild_of(query[abs]) << (8 - q);
    }
    if (left) return skm::sift4_unified(skm::reverse_pairs(ref16, 8), q18, w9, 1, offset + 8);
    return skm::sift4_unified(ref16, q18, w9, 0, len - offset);
}

extern "C" unsigned reverse_pairs_c(unsigned x, int n) { return skm::reverse_pairs(x, n); }
extern "C" unsigned reverse_bits_c(unsigned x, int n) { return skm::reverse_bits(x, n); }
