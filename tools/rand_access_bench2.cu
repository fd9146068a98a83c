// Microbenchmark 2: HOW a 64-byte bucket of a multi-GB table is fetched, beyond the 256 MB TLB
// reach (rand_access_bench.cu showed ~1 TB/s whatever the access size: a cost per lane request).
// Variants, all reading the same number of random 64-byte buckets (32-byte ones where stated):
//   a  4 x LDG.128 per lane                  (what probe_canonical did)
//   b  2 x LDG.256 per lane
//   c  1 x LDG.256 per lane, 32-byte buckets
//   d  4 lanes per bucket, one LDG.128 each: a warp instruction covers 8 buckets
//   e  2 lanes per bucket, one LDG.256 each: a warp instruction covers 16 buckets
//   f  cp.async.bulk of 64 bytes per lane into shared memory (mbarrier per lane)
//   g  variant a on a table mapped through the VMM API with 512 MB alignment
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/rand_access_bench2 tools/rand_access_bench2.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 31; x *= 0x9E3779B97F4A7C15ULL; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
    return x;
}

struct Quad64 { uint64_t a, b, c, d; };
__device__ __forceinline__ Quad64 ld256(const void *p)
{
    Quad64 v;
    asm volatile("ld.global.cs.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}

// mode: 0 a, 1 b, 2 c, 3 d, 4 e, 5 f
template <int MODE>
__global__ void probe(const char *table, uint64_t n_buckets64, int chain, unsigned long long *out, long long n_threads)
{
    extern __shared__ __align__(128) unsigned char smem[];
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint64_t acc = 0;
    uint32_t land = 0, bar = 0, phase = 0;
    if (MODE == 5) {
        land = (uint32_t)__cvta_generic_to_shared(smem + threadIdx.x * 64);
        bar = (uint32_t)__cvta_generic_to_shared(smem + blockDim.x * 64 + threadIdx.x * 8);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    for (; i < n_threads; i += (long long)gridDim.x * blockDim.x) {
        uint64_t h = mix((uint64_t)i);
        for (int c = 0; c < chain; ++c) {
            uint64_t v = 0;
            if (MODE == 0) {
                const ulonglong2 *p = (const ulonglong2 *)(table + (h & (n_buckets64 - 1)) * 64);
                const ulonglong2 s0 = __ldcs(p), s1 = __ldcs(p + 1), s2 = __ldcs(p + 2), s3 = __ldcs(p + 3);
                v = s0.x ^ s0.y ^ s1.x ^ s1.y ^ s2.x ^ s2.y ^ s3.x ^ s3.y;
            } else if (MODE == 1) {
                const char *p = table + (h & (n_buckets64 - 1)) * 64;
                const Quad64 s0 = ld256(p), s1 = ld256(p + 32);
                v = s0.a ^ s0.b ^ s0.c ^ s0.d ^ s1.a ^ s1.b ^ s1.c ^ s1.d;
            } else if (MODE == 2) {
                const char *p = table + (h & (n_buckets64 * 2 - 1)) * 32;
                const Quad64 s0 = ld256(p);
                v = s0.a ^ s0.b ^ s0.c ^ s0.d;
            } else if (MODE == 3) {
                // every lane has its own bucket; four rounds of eight buckets
                const uint64_t mine = (h & (n_buckets64 - 1)) * 64;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint64_t off = __shfl_sync(0xffffffffu, mine, r * 8 + (lane >> 2));
                    const ulonglong2 s = __ldcs((const ulonglong2 *)(table + off) + (lane & 3));
                    uint64_t x = s.x ^ s.y;
                    x ^= __shfl_xor_sync(0xffffffffu, x, 1);
                    x ^= __shfl_xor_sync(0xffffffffu, x, 2);
                    const uint64_t back = __shfl_sync(0xffffffffu, x, (lane & 7) * 4);
                    if ((lane >> 3) == r) v = back;
                }
            } else if (MODE == 4) {
                const uint64_t mine = (h & (n_buckets64 - 1)) * 64;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint64_t off = __shfl_sync(0xffffffffu, mine, r * 16 + (lane >> 1));
                    const Quad64 s = ld256(table + off + (lane & 1) * 32);
                    uint64_t x = s.a ^ s.b ^ s.c ^ s.d;
                    x ^= __shfl_xor_sync(0xffffffffu, x, 1);
                    const uint64_t back = __shfl_sync(0xffffffffu, x, (lane & 15) * 2);
                    if ((lane >> 4) == r) v = back;
                }
            } else {
                const char *p = table + (h & (n_buckets64 - 1)) * 64;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 64;" ::"r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                             ::"r"(land), "l"(p), "r"(bar) : "memory");
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(phase) : "memory");
                }
                phase ^= 1;
                const ulonglong2 *s = (const ulonglong2 *)(smem + threadIdx.x * 64);
                v = s[0].x ^ s[0].y ^ s[1].x ^ s[1].y ^ s[2].x ^ s[2].y ^ s[3].x ^ s[3].y;
            }
            acc ^= v;
            h = mix(h + 1);
        }
    }
    if (acc == 0x1234567) *out = acc;
}

static void *vmm_alloc(size_t bytes)
{
    if (cuInit(0) != CUDA_SUCCESS) return nullptr;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gran_min = 0, gran_rec = 0;
    cuMemGetAllocationGranularity(&gran_min, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    cuMemGetAllocationGranularity(&gran_rec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    printf("{\"vmm_granularity_min\": %zu, \"vmm_granularity_recommended\": %zu}\n", gran_min, gran_rec);
    const size_t big = 512ULL << 20;
    bytes = (bytes + big - 1) / big * big;
    CUmemGenericAllocationHandle h;
    if (cuMemCreate(&h, bytes, &prop, 0) != CUDA_SUCCESS) return nullptr;
    CUdeviceptr va;
    if (cuMemAddressReserve(&va, bytes, big, 0, 0) != CUDA_SUCCESS) return nullptr;
    if (cuMemMap(va, bytes, 0, h, 0) != CUDA_SUCCESS) return nullptr;
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (cuMemSetAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) return nullptr;
    return (void *)va;
}

int main(int argc, char **argv)
{
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 4096LL) << 20;
    cudaSetDevice(0);
    cudaFree(0);
    void *table, *vtable = nullptr;
    if (cudaMalloc(&table, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    cudaMemset(table, 1, bytes);
    vtable = vmm_alloc(bytes);
    if (vtable) cudaMemset(vtable, 1, bytes);
    else printf("{\"vmm\": \"unavailable\"}\n");
    unsigned long long *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const long long n = 1LL << 26;
    const int chain = 8;
    const char *names = "abcdefg";
    for (int tps = 768; tps <= 2048; tps = tps == 768 ? 1024 : tps * 2)
        for (int mode = 0; mode < 7; ++mode) {
            const char *t = (const char *)(mode == 6 ? vtable : table);
            if (!t) continue;
            const int block = 256, grid = 148 * tps / block;
            const uint64_t nb = bytes / 64;
            const size_t smem = 256 * 72;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                switch (mode) {
                case 0: case 6: probe<0><<<grid, block>>>(t, nb, chain, out, n / chain); break;
                case 1: probe<1><<<grid, block>>>(t, nb, chain, out, n / chain); break;
                case 2: probe<2><<<grid, block>>>(t, nb, chain, out, n / chain); break;
                case 3: probe<3><<<grid, block>>>(t, nb, chain, out, n / chain); break;
                case 4: probe<4><<<grid, block>>>(t, nb, chain, out, n / chain); break;
                case 5: probe<5><<<grid, block, smem>>>(t, nb, chain, out, n / chain); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = ms < best ? ms : best;
            }
            printf("{\"table_mb\": %zu, \"variant\": \"%c\", \"threads_per_sm\": %d, \"ms\": %.3f, \"G_buckets_per_s\": %.2f, \"err\": \"%s\"}\n",
                   bytes >> 20, names[mode], tps, best, n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
