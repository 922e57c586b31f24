// tma_stream.cu -- standalone probe (not part of the library): how fast can one CTA per SM stream HBM into shared memory with
// cp.async.bulk (1-D TMA) through an mbarrier ring?  Varies stage size, ring depth, issuing pattern and CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tma_stream tools/tma_stream.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// mode 0: tile order interleaved over CTAs (K2's), chunk = tile_bytes / stage_bytes stages per tile
__global__ void __launch_bounds__(128, 1) stream_kernel(const char* base, long long n_tiles, int tile_bytes, int stage_bytes, int n_stages, int consume) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[64], empty_bar[64];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int per_tile = tile_bytes / stage_bytes;
    const uint32_t ring = smem_u32(smem), full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
    if (warp == 0) {
        if (threadIdx.x == 0) {
            int s = 0; uint32_t ph = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const char* src = base + (size_t)t * tile_bytes;
                for (int c = 0; c < per_tile; ++c) {
                    mbar_wait(empty_s + 8u * s, ph ^ 1u);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_s + 8u * s), "r"((uint32_t)stage_bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(ring + (uint32_t)s * stage_bytes), "l"(src + (size_t)c * stage_bytes), "r"((uint32_t)stage_bytes), "r"(full_s + 8u * s) : "memory");
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0;
        uint32_t acc = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int c = 0; c < per_tile; ++c) {
                mbar_wait(full_s + 8u * s, ph);
                if (consume) {      // read the stage with the warp (ld.shared.v4), like a consumer would
                    const uint4* p = (const uint4*)(smem + (size_t)s * stage_bytes);
                    for (int u = threadIdx.x & 31; u < stage_bytes / 16; u += 32) { uint4 v = p[u]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
                }
                __syncwarp();
                if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_s + 8u * s) : "memory");
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
        }
        if (acc == 0x12345u) printf("x");
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
    const int tile_bytes = 192 * 1024;
    const long long n_tiles = 148LL * 48;                 // 1.36 GB per launch
    const int reps = 12;
    char* buf; CK(cudaMalloc(&buf, (size_t)n_tiles * tile_bytes * reps)); CK(cudaMemset(buf, 1, (size_t)n_tiles * tile_bytes * reps));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int stage_bytes, n_stages, blocks, consume; };
    const Cfg cfgs[] = {{16384, 12, 148, 0}, {16384, 6, 148, 0}, {16384, 3, 148, 0}, {16384, 2, 148, 0}, {8192, 24, 148, 0}, {8192, 8, 148, 0}, {32768, 6, 148, 0},
                        {4096, 48, 148, 0}, {65536, 3, 148, 0}, {16384, 6, 296, 0}, {16384, 3, 592, 0}, {16384, 12, 148, 1}, {16384, 6, 296, 1}, {8192, 4, 148 * 6, 0}};
    for (const Cfg& c : cfgs) {
        const size_t smem = (size_t)c.stage_bytes * c.n_stages;
        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        std::vector<float> ms;
        for (int i = 0; i < reps; ++i) {
            cudaEventRecord(e0);
            stream_kernel<<<c.blocks, 128, smem>>>(buf + (size_t)i * n_tiles * tile_bytes, n_tiles, tile_bytes, c.stage_bytes, c.n_stages, c.consume);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
        }
        std::sort(ms.begin(), ms.end());
        printf("stage %6d B x %2d stages (%3zu KiB ring), %4d CTAs, consume %d: median %.1f us -> %.0f GB/s\n", c.stage_bytes, c.n_stages, smem / 1024, c.blocks, c.consume,
               ms[reps / 2] * 1e3, (double)n_tiles * tile_bytes / (ms[reps / 2] * 1e-3) / 1e9);
    }
    return 0;
}
