// Shared host/device definitions for the seekmer_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/seekmer_b200.h"

#define SKM_API extern "C" __attribute__((visibility("default")))

namespace skm {

constexpr int K = SKM_KMER_SIZE;
constexpr uint64_t KMER_MASK = (1ULL << (2 * K)) - 1;
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFULL;

// ---- error plumbing -------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define SKM_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            return ::skm::fail(_e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA, \
                               std::string(#expr) + ": " + cudaGetErrorString(_e));     \
        }                                                                               \
    } while (0)

// ---- device-side index layout (the re-laid-out form of KMerIndex) ------------------
// 16-byte slot; key = canonical (min of k-mer and its reverse complement) 25-mer,
// value = contig coordinate expressed for the canonical orientation.
struct __align__(16) Slot {
    uint64_t key;
    int32_t entry;
    int32_t offset;
};

// 128-byte contig record.  First 64 bytes (one DRAM burst, what P_CONTIG reads): header plus
// the first 8 entries of the contig's target list (most lists fit; longer ones continue in
// targets[]).  Second 64 bytes: the graph LINKS of the contig - the coordinates map_kmer returns
// for the 4 k-mers that extend its last k-mer to the right and the 4 that extend its first k-mer
// to the left, probed once when the index is laid out.  A contig walk that crosses a junction
// reads the next coordinate here (8 bytes, L2-resident) instead of hashing and probing the
// 4 GB k-mer table; the other strand follows from map_kmer(rc(x)) = {~entry, offset}.
//   w0 = first_kmer | (target_count low 14 bits  << 50)
//   w1 = last_kmer  | (target_count high 14 bits << 50)
constexpr int INLINE_TARGETS = 8;
struct __align__(64) ContigRec {
    uint64_t w0;
    uint64_t w1;
    int64_t seq_offset;      // base offset into the packed sequence pool
    uint32_t target_offset;  // into targets[]
    uint32_t length;         // contig length in bases
    int32_t inline_targets[INLINE_TARGETS];  // targets[target_offset .. +8), zero padded
    int2 right_of_last[4];   // map_kmer(append(last_kmer, b)),   b = A C G T; {entry, offset}
    int2 left_of_first[4];   // map_kmer(prepend(first_kmer, b))
};

// The k-mer table is probed by BUCKET: 4 consecutive 16-byte slots = 64 bytes = one DRAM burst.
// A key lives in the first bucket from its home bucket onwards that had a free slot when it was
// inserted; buckets fill front to back, so "last slot empty" <=> "bucket not full".
constexpr int BUCKET_SLOTS = 4;

struct DevIndex {
    const Slot *table;
    uint64_t bucket_mask;     // n_slots / BUCKET_SLOTS - 1
    const ContigRec *contigs;
    const uint32_t *seq2;     // 16 bases per word, first base in the top bits
    const int32_t *targets;   // signed entries only
    int64_t n_contigs;
    int64_t n_bases;
    int64_t n_targets;
    // L2 cache policies (createpolicy values, made once per index by make_policies_kernel):
    // `pol_hot` goes with every load of contig records, links, sequences and targets, `pol_stream`
    // with the loads of table buckets and packed reads, which are used once
    uint64_t pol_hot;
    uint64_t pol_stream;
};

// status bits raised by kernels (checked by the host after each batch)
enum : uint32_t {
    ST_ARENA_FULL = 1u,
    ST_DICT_FULL = 2u,
    ST_POOL_FULL = 4u,
    ST_SHORT_READ = 8u,
    ST_KEY_COLLISION = 16u,  // two different id tuples with the same 128-bit key (never seen; refused)
};

}  // namespace skm

struct skm_index {
    int device = 0;
    skm::DevIndex d{};
    skm::Slot *table = nullptr;
    unsigned char *hot = nullptr;  // one block: contigs | seq2 | targets (the L2-resident part of the index)
    int64_t hot_bytes = 0;
    skm::ContigRec *contigs = nullptr;
    uint32_t *seq2 = nullptr;
    int32_t *targets = nullptr;
    int64_t n_slots = 0, n_kmers = 0, n_contigs = 0, n_bases = 0, n_targets = 0;
    int64_t n_transcripts = 0, max_target_count = 0, bytes = 0;
};
