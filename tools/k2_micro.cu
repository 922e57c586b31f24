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
    const int dbgs[] = {0, 0, 0, 4, 12, 0};
    const int mcast[] = {1, 0, 1, 1, 1, 1};
    for (int vi = 0; vi < 6; ++vi) {
        const int variant = vi == 0 ? 0 : (vi == 5 ? 2 : 1);
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
        printf("dbg %2d multicast %d variant %d (%s): median %.1f us  min %.1f us  -> %.0f TFLOP/s, corpus read %.0f GB/s\n", dbg, mcast[vi], variant, variant == 2 ? "TS 2-CTA" : (variant ? "TS" : "SS"),
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
