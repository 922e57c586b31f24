// k2_micro.cu -- standalone micro-benchmark + timeline of the dense tile kernels K2 (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DDHR_K2_TRACE -o gpurun_out/k2_micro tools/k2_micro.cu -lcuda
//   k2_micro [rows_per_launch=37888] [queries=256] [C=768] [k_blocked_copy=1] [same_rows=0] [random_data=0]
// Launches K2 (mode 1: scratch writes) over successive sub-chunks of a synthetic dense block and prints the median
// kernel time per variant plus, for the TS variant, the clock64 timeline of CTA 0.
#include "../dhr_b200/csrc/dense_tile.cu"

#include <algorithm>
#include <vector>

namespace dhr {
void set_cuda_error(cudaError_t e, const char* what, const char*, int line) { fprintf(stderr, "cuda error %d (%s) line %d\n", (int)e, what, line); }
}

__global__ void fill_random_halves(__half* p, size_t n, unsigned seed, float scale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u ^ seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
        p[i] = __float2half_rn(((float)(x & 0xFFFF) / 32768.0f - 1.0f) * scale);
    }
}

__global__ void read_stream_kernel(const uint4* p, size_t n, unsigned* out) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *out = acc;
}

// the K2 access pattern with plain loads: block c reads tiles c, c + grid, ... of 192 KiB each (or, `range`, its own contiguous run)
__global__ void __launch_bounds__(1024) pattern_read_kernel(const uint4* p, int n_tiles, int tile_u4, int range, unsigned* out) {
    unsigned acc = 0;
    const int per = n_tiles / gridDim.x;
    for (int i = 0; i < per; ++i) {
        const int t = range ? blockIdx.x * per + i : i * gridDim.x + blockIdx.x;
        const uint4* q = p + (size_t)t * tile_u4;
        for (int u = threadIdx.x; u < tile_u4; u += 4096) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (u + j * 1024 < tile_u4)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(q + u + j * 1024));
                else v[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
        }
    }
    if (acc == 0x12345678u) *out = acc;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const long long sub = argc > 1 ? atoll(argv[1]) : 37888;
    const int nq = argc > 2 ? atoi(argv[2]) : 256;
    const int C = argc > 3 ? atoi(argv[3]) : 768;
    const bool same_rows = argc > 5 && atoi(argv[5]) != 0;     // relaunch over the same (L2-resident) rows
    const long long n_rows = sub * 40;
    dhr_index h;
    h.g.C = C; h.g.C_pad = (C + 7) / 8 * 8; h.n_rows = n_rows;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); h.num_sms = prop.multiProcessorCount;
    CK(cudaMalloc(&h.dns, (size_t)n_rows * h.g.C_pad * 2)); CK(cudaMemset(h.dns, 0x3c, (size_t)n_rows * h.g.C_pad * 2));
    if (argc <= 4 || atoi(argv[4]) != 0) {   // K-blocked corpus copy (content irrelevant for timing)
        const size_t nb = (size_t)((n_rows + 127) / 128 * 128) * ((h.g.C_pad + 63) / 64) * 64 * 2;
        CK(cudaMalloc(&h.dnst, nb)); CK(cudaMemset(h.dnst, 0x3c, nb));
    }
    void* q; CK(cudaMalloc(&q, (size_t)nq * h.g.C_pad * 2)); CK(cudaMemset(q, 0x2c, (size_t)nq * h.g.C_pad * 2));
    if (argc > 6 && atoi(argv[6]) != 0) {      // random operands instead of constants (data-dependent power / timing)
        fill_random_halves<<<1184, 256>>>((__half*)q, (size_t)nq * h.g.C_pad, 1u, 0.05f);
        fill_random_halves<<<1184, 256>>>(h.dns, (size_t)n_rows * h.g.C_pad, 2u, 0.05f);
        if (h.dnst) fill_random_halves<<<1184, 256>>>(h.dnst, (size_t)((n_rows + 127) / 128 * 128) * ((h.g.C_pad + 63) / 64) * 64, 3u, 0.05f);
        CK(cudaDeviceSynchronize());
    }
    float* scratch; CK(cudaMalloc(&scratch, (size_t)sub * 256 * 4));
    dhr::TopkState t; CK(cudaMalloc(&t.tau, 256 * 4)); CK(cudaMalloc(&t.cnt, 256 * 4));
    CK(cudaMemset(t.tau, 0x7f, 256 * 4)); CK(cudaMemset(t.cnt, 0, 256 * 4));
    const float tau_arg = argc > 7 ? (float)atof(argv[7]) : 0.f;            // finite admission threshold for the dense-only (mode 0) runs
    CK(cudaMalloc(&t.cand_score, (size_t)256 * 16384 * 4)); CK(cudaMalloc(&t.cand_row, (size_t)256 * 16384 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (h.dnst) {     // calibration: the K2 tile order with plain loads
        const int n_tiles = (int)(sub / 128), tile_u4 = ((h.g.C_pad + 63) / 64) * 16384 / 16;
        for (int mode = 0; mode < 4; ++mode) {
            const int blocks = (mode & 2) ? 296 : 148, range = mode & 1;
            std::vector<float> ms;
            for (int i = 0; i < 20; ++i) {
                cudaEventRecord(e0);
                pattern_read_kernel<<<blocks, 1024>>>((const uint4*)((const char*)h.dnst + (size_t)i * n_tiles * tile_u4 * 16), n_tiles, tile_u4, range, (unsigned*)t.cnt);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
            }
            std::sort(ms.begin(), ms.end());
            printf("read stream, K2 tile order (%d blocks x 1024 threads, %s): median %.1f us -> %.0f GB/s\n", blocks, range ? "contiguous run per block" : "tiles interleaved over blocks",
                   ms[10] * 1e3, (double)(n_tiles / blocks * blocks) * tile_u4 * 16 / (ms[10] * 1e-3) / 1e9);
        }
    }
    if (h.dnst) {     // calibration: plain vector-load read stream over the same buffer
        const size_t nb = (size_t)sub * h.g.C_pad * 2;
        for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
            std::vector<float> ms;
            for (int i = 0; i < 20; ++i) {
                cudaEventRecord(e0);
                read_stream_kernel<<<blocks, 512>>>((const uint4*)((const char*)h.dnst + (size_t)i * nb), nb / 16, (unsigned*)t.cnt);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
            }
            std::sort(ms.begin(), ms.end());
            printf("read stream (ld.global.nc.v4, %d x 512 threads): median %.1f us -> %.0f GB/s\n", blocks, ms[10] * 1e3, nb / (ms[10] * 1e-3) / 1e9);
        }
    }
    const int dbgs[] =  {0, 4, 12, 84, 84 + 2048, 64 + 2048, 0, 0, 4, 84, 0, 0};
    const int mcast[] = {1, 1, 1,  1,  1,  1,  0, 1, 1, 1, 1, 1};
    const int vars[] =  {2, 2, 2,  2,  2,  2,  1, 1, 1, 1, 2, 2};
    for (int vi = 0; vi < 10; ++vi) {
        const int variant = vars[vi];
        const int dbg = dbgs[vi];
        CK(cudaMemcpyToSymbol(dhr::g_k2_dbg, &dbg, sizeof(int)));
        h.opt_dense_variant = variant;
        h.opt_dense_multicast = mcast[vi] * 2;
        std::vector<float> ms;
        for (int i = 0; i < 40; ++i) {
            const long long r0 = same_rows ? 0 : (long long)i * sub;
            cudaEventRecord(e0);
            int rc = dhr::launch_dense_tile(&h, q, nq, r0, r0, r0 + sub, 1, scratch, 256, t, 16384, 0);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            if (rc != 0) { fprintf(stderr, "launch rc %d\n", rc); return 1; }
            float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
        }
        std::sort(ms.begin(), ms.end());
        const double flops = 2.0 * sub * nq * C;
        printf("dbg %5d multicast %d variant %d (%s): median %.1f us  min %.1f us  -> %.0f TFLOP/s, corpus read %.0f GB/s\n", dbg, mcast[vi], variant, variant == 2 ? "TS 2-CTA" : (variant ? "TS" : "SS"),
               ms[20] * 1e3, ms[0] * 1e3, flops / (ms[20] * 1e-3) / 1e12, sub * h.g.C_pad * 2.0 / (ms[20] * 1e-3) / 1e9);
    }
    // dense-only mode (filter + append, nothing passes tau): one launch over all rows, both variants
    { const int z = 0; CK(cudaMemcpyToSymbol(dhr::g_k2_dbg, &z, sizeof(int))); }
    if (argc > 7) { std::vector<float> tv(256, tau_arg); CK(cudaMemcpy(t.tau, tv.data(), 256 * 4, cudaMemcpyHostToDevice)); }
    for (int variant = 1; variant < 4; ++variant) {
        h.opt_dense_variant = variant == 3 ? 2 : 1;
        h.opt_dense_multicast = variant == 3 ? 1 : variant - 1;
        float best = 1e9f;
        for (int i = 0; i < 5; ++i) {
            CK(cudaMemset(t.cnt, 0, 256 * 4));
            cudaEventRecord(e0);
            int rc = dhr::launch_dense_tile(&h, q, nq, 0, 0, n_rows, 0, nullptr, 0, t, 16384, 0);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            if (rc != 0) { fprintf(stderr, "launch rc %d\n", rc); return 1; }
            float m; cudaEventElapsedTime(&m, e0, e1); best = std::min(best, m);
        }
        { std::vector<unsigned> cn(256); CK(cudaMemcpy(cn.data(), t.cnt, 256 * 4, cudaMemcpyDeviceToHost)); double tot = 0; for (unsigned x : cn) tot += x;
          printf("  passes per query %.1f (tau %g) | ", tot / 256, tau_arg); }
        printf("mode 0 (filter) %s multicast %d: %lld rows x %d queries in %.1f us -> %.0f TFLOP/s, corpus read %.0f GB/s\n", variant == 3 ? "TS 2-CTA" : "TS", h.opt_dense_multicast, n_rows, nq,
               best * 1e3, 2.0 * n_rows * nq * C / (best * 1e-3) / 1e12, n_rows * h.g.C_pad * 2.0 / (best * 1e-3) / 1e9);
    }
    long long tr[8][64];
    CK(cudaMemcpyFromSymbol(tr, dhr::g_k2_trace, sizeof(tr)));
    const long long t0 = tr[0][0];
    printf("TS timeline of CTA 0 (cycles from kernel entry): tmem alloc done %lld, query operand in TMEM %lld, epilogue done %lld, end %lld\n",
           tr[0][1] - t0, tr[0][2] - t0, tr[0][3] - t0, tr[0][4] - t0);
    for (int i = 0; i < 10; ++i)
        printf("  tile %d: producer start %lld | mma: wait-begin %lld acc-free %lld issued %lld | epilogue: acc-ready %lld drained %lld\n", i,
               tr[1][i] - t0, tr[2][i] - t0, tr[3][i] - t0, tr[4][i] - t0, tr[5][i] - t0, tr[6][i] - t0);
    return 0;
}
