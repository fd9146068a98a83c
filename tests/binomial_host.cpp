// Host build of seekmer_b200/csrc/binomial.cuh for CPU goodness-of-fit tests
// (tests/test_binomial.py), and the split tree of em.cu's multinomial_tree_kernel walked on the
// host with the same node / replicate streams.
#include <vector>

#include "../seekmer_b200/csrc/binomial.cuh"

extern "C" void binomial_batch(int64_t n, double p, uint64_t seed, int64_t count, int64_t *out)
{
    for (int64_t i = 0; i < count; ++i) {
        skm::UniformStream rng{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, (uint32_t)(i >> 32), 7u, 0u};
        out[i] = skm::binomial_draw(n, p, rng);
    }
}

// counts[n_classes] -> out[n_classes] for replicate `replicate` (level by level, like the kernel)
extern "C" void multinomial_tree_host(const int64_t *counts, int64_t n_classes, uint64_t seed, int64_t replicate,
                                      int64_t *out)
{
    std::vector<uint64_t> cum((size_t)n_classes);
    uint64_t acc = 0;
    for (int64_t c = 0; c < n_classes; ++c) cum[(size_t)c] = acc += (uint64_t)counts[c];
    int levels = 0;
    while ((1LL << levels) < n_classes) ++levels;
    if (levels == 0) {
        out[0] = (int64_t)acc;
        return;
    }
    std::vector<int64_t> parent(1, (int64_t)acc), child;
    for (int level = 0; level < levels; ++level) {
        const int64_t nodes = 1LL << level, span = 1LL << (levels - level);
        child.assign((size_t)(2 * nodes), 0);
        for (int64_t i = 0; i < nodes; ++i) {
            const int64_t lo = std::min(i * span, n_classes), mid = std::min(i * span + span / 2, n_classes),
                          hi = std::min((i + 1) * span, n_classes);
            const uint64_t w_lo = lo ? cum[(size_t)lo - 1] : 0, w_mid = mid ? cum[(size_t)mid - 1] : 0,
                           w_hi = hi ? cum[(size_t)hi - 1] : 0;
            const int64_t n_node = parent[(size_t)i];
            const uint64_t w_left = w_mid - w_lo, w_node = w_hi - w_lo;
            int64_t left = 0;
            if (n_node > 0 && w_left > 0) {
                if (w_left == w_node) {
                    left = n_node;
                } else {
                    skm::UniformStream rng{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(nodes + i),
                                           (uint32_t)replicate, 0x54524545u, 0u};
                    left = skm::binomial_draw(n_node, (double)w_left / (double)w_node, rng);
                }
            }
            child[(size_t)(2 * i)] = left;
            child[(size_t)(2 * i + 1)] = n_node - left;
        }
        parent.swap(child);
    }
    for (int64_t c = 0; c < n_classes; ++c) out[c] = parent[(size_t)c];
}
