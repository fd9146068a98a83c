// An EM plan: the class x transcript incidence structure resident on one device in both
// orders (CSR by class, CSC by transcript), built once and shared by the main EM, the bootstrap
// replicates and repeated calls.  A plan made from a mapper (skm_em_plan_from_mapper) also owns
// the integer class counts in the dictionary's first-seen order; nothing of it crosses PCIe.
#pragma once

#include "common.cuh"

struct skm_em_plan {
    int device = 0;
    int64_t C = 0, nnz = 0, T = 0;
    int32_t *perm = nullptr;       // [C] plan class i = the caller's class perm[i] (classes sorted by first transcript)
    int64_t *class_ptr = nullptr;  // [C + 1], plan class order
    int32_t *class_tx = nullptr;   // [nnz] transcript ids in tuple order
    int64_t *tx_ptr = nullptr;     // [T + 1]
    int32_t *tx_class = nullptr;   // [nnz] class of every entry of a transcript, in nnz order
    int64_t *counts = nullptr;     // [C] integer class counts in the CALLER's class order, or NULL
    int32_t *heavy_rows = nullptr; // transcripts with more entries than one 8-lane group should sum
    int32_t n_heavy = 0;
};

namespace skm {
// Device memory from the per-device block cache of em.cu (no cudaMalloc / cudaFree in steady
// state); what a plan is made of.
cudaError_t dev_alloc(int device, size_t bytes, void **out);
void dev_free(int device, void *p);
// Called when an index has been put on a device: the cache takes its first slab and one large
// block from the driver and keeps them, so that the allocator's one-off costs (first VMM calls of
// a process on a cold device: tens of ms each, whatever the size) are paid with the index load
// and not inside the first EM call.
void scratch_warm(int device);

// Takes ownership of class_ptr / class_tx / counts (device memory from dev_alloc on `device`;
// counts may be NULL), builds the CSC side on `stream` and returns the plan.  On failure the
// buffers are freed.
int em_plan_adopt(int device, int64_t C, int64_t nnz, int64_t T, int64_t *class_ptr, int32_t *class_tx,
                  int64_t *counts, cudaStream_t stream, skm_em_plan **out);
}  // namespace skm
