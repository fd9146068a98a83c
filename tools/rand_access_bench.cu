// Microbenchmark: what random 64-byte reads over a multi-GB table deliver on this GPU - the
// practical ceiling for the bucket probes of map_reads_kernel.  Each thread
// reads `chain` buckets; with dependent = 1 the next bucket index depends on the data of the
// previous one (a pointer chase, like probe -> contig -> link), otherwise they are independent.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/rand_access_bench tools/rand_access_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 31; x *= 0x9E3779B97F4A7C15ULL; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
    return x;
}

template <int BYTES>
__global__ void chase(const ulonglong2 *table, uint64_t mask, int chain, int dependent, unsigned long long *out,
                      long long n_threads)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (; i < n_threads; i += (long long)gridDim.x * blockDim.x) {
        uint64_t h = mix((uint64_t)i);
        for (int c = 0; c < chain; ++c) {
            const ulonglong2 *p = table + (h & mask) * (BYTES / 16);
            uint64_t v = 0;
#pragma unroll
            for (int k = 0; k < BYTES / 16; ++k) {
                const ulonglong2 s = __ldcs(p + k);
                v ^= s.x ^ s.y;
            }
            acc ^= v;
            h = mix(h + 1 + (dependent ? v : 0));
        }
    }
    if (acc == 0x1234567) *out = acc;
}

int main(int argc, char **argv)
{
    // usage: rand_access_bench [max MB]: 64-byte random reads over tables of 64 MB .. max MB
    const size_t max_bytes = (argc > 1 ? atoll(argv[1]) : 16384LL) << 20;
    void *table;
    if (cudaMalloc(&table, max_bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    cudaMemset(table, 1, max_bytes);
    unsigned long long *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const long long n = 1LL << 26;
    for (size_t bytes = 64ULL << 20; bytes <= max_bytes; bytes *= 2)
        for (int size = 32; size <= 128; size *= 2)
            for (int dep = 0; dep <= 1; ++dep) {
                if (size != 64 && dep) continue;
                const int threads_per_sm = 1024;
                const int block = 256, grid = 148 * threads_per_sm / block;
                const uint64_t mask = bytes / size - 1;
                float best = 1e9f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    if (size == 32) chase<32><<<grid, block>>>((const ulonglong2 *)table, mask, 8, dep, out, n / 8);
                    else if (size == 64) chase<64><<<grid, block>>>((const ulonglong2 *)table, mask, 8, dep, out, n / 8);
                    else chase<128><<<grid, block>>>((const ulonglong2 *)table, mask, 8, dep, out, n / 8);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    best = ms < best ? ms : best;
                }
                printf("{\"table_mb\": %zu, \"bytes_per_access\": %d, \"dependent\": %d, \"threads_per_sm\": %d, \"ms\": %.3f, "
                       "\"G_access_per_s\": %.2f, \"GB_per_s\": %.1f}\n",
                       bytes >> 20, size, dep, threads_per_sm, best, n / best / 1e6, (double)n * size / best / 1e6);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
