// k1t_v7_proto.cu -- self-checking prototype of the planned K1t v7 walk ("postings within the tile", DESIGN.md section 7,
// executable spec: tools/k1t_postings_spec.py).  NOT part of the library.  First (and so far only) run on a B200:
// bit-exact, 106 us per launch (profiles/r1_k1t_v7_proto.txt).  One command:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/k1t_v7 tools/k1t_v7_proto.cu && /tmp/k1t_v7
//
// It builds (on the host) the code-sorted tile layout for a synthetic sub-chunk of the benchmark's shape (37,888 passages,
// 128 slices x 6 values, 39 codes, 30 % empty slices) and the per-(query tile, chunk) query blocks for 256 queries, runs the
// kernel, compares acc[q][p] with a host computation that uses the same fp32 operation order (bit-exact), and prints the
// time per launch -- to be read against the current lex_tile_kernel (about 92 us per launch of the same shape, of which
// about 8 % is the dense-score init and 9 % the admission filter that this prototype replaces by a plain store).
//
// Layout per (tile of 512 passages, chunk of 4 slices):
//   off u16 [4][rt + 1]   item index (within the block) of the first item of (slice, code); [rt] = end of the slice
//   items uint4 [n]       {passage id u16 | v0, v1 | v2, v3 | v4, v5 | 0}, slices in order, inside a slice sorted by
//                         (code, passage id); all-zero slices are not stored
// Query block per (query tile of 64, chunk): uint4 [4][64] = {code u16 (0xFFFF = empty) | v0, v1 | v2, v3 | v4, v5 | 0}.
// Walk: consumer warp w owns queries 4w .. 4w+3 (accumulator rows, so no two warps touch the same row).  For each slice
// it looks up the four lists (slice, code of its query), flattens their items over the 32 lanes and does
// acc[q][p] += dot: concurrently active lanes differ in q or in p, so there are no races and no atomics.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

constexpr int PT = 512, QT = 64, SC = 4, G = 6, STAGES = 3;
constexpr int CONSUMERS = 512, THREADS = CONSUMERS + 32;

// ---- device helpers (same PTX as dhr_b200/csrc/common.cuh) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }
template <bool HI>
__device__ __forceinline__ float fma_hh(uint32_t a, uint32_t b, float c) {      // f32 += f16 * f16, same half of both words
    float d;
    if constexpr (HI)
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bh, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    else
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bl, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    return d;
}

struct Args {
    const uint8_t* lexp; const unsigned long long* blk_off; const uint32_t* blk_bytes;   // [tile][chunk]
    const uint8_t* qblocks;                                                              // [qtile][chunk] of QBLOCK bytes
    float* out;                                                                          // [qtile][tile][QT][PT]
    int n_tiles, n_chunks, n_qtiles, rt, hdr_bytes, stage_bytes, pblock_smem;
    int store;                                                                           // 0: skip the accumulator store (walk time only)
};
constexpr int QBLOCK = SC * QT * 16;

__global__ void __launch_bounds__(THREADS, 1) k1t_v7_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    float* acc = (float*)smem;                                   // [QT][PT]
    uint8_t* stages = smem + (size_t)QT * PT * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % a.n_qtiles, cta_in_q = blockIdx.x / a.n_qtiles, ctas_per_q = gridDim.x / a.n_qtiles;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMERS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == CONSUMERS / 32) {                                // ===== producer warp =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q)
                for (int c = 0; c < a.n_chunks; ++c) {
                    const size_t b = (size_t)t * a.n_chunks + c;
                    const uint32_t bytes = a.blk_bytes[b];
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    uint8_t* dst = stages + (size_t)s * a.stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], bytes + QBLOCK);
                    bulk_g2s(dst, a.lexp + a.blk_off[b], bytes, &full_bar[s]);
                    bulk_g2s(dst + a.pblock_smem, a.qblocks + ((size_t)qt * a.n_chunks + c) * QBLOCK, QBLOCK, &full_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
        }
        return;
    }
    // ===== consumers =====
    const int p = threadIdx.x;                                   // passage owned in the zero / store phases
    const int q_base = warp * 4;                                 // queries owned in the walk
    const uint32_t per = (uint32_t)a.rt + 1u;
    int s = 0; uint32_t ph = 0;
    for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q) {
#pragma unroll 16
        for (int q = 0; q < QT; ++q) acc[q * PT + p] = 0.f;
        consumer_barrier();                                      // rows are handed from passage threads to query warps
        for (int c = 0; c < a.n_chunks; ++c) {
            mbar_wait(&full_bar[s], ph);
            const uint8_t* st = stages + (size_t)s * a.stage_bytes;
            const uint16_t* off = (const uint16_t*)st;
            const uint4* items = (const uint4*)(st + a.hdr_bytes);
            const uint4* qent = (const uint4*)(st + a.pblock_smem);
#pragma unroll 1
            for (int j = 0; j < SC; ++j) {
                uint32_t start = 0, len = 0;
                if (lane < 4) {
                    const uint32_t code = qent[j * QT + q_base + lane].x & 0xFFFFu;
                    if (code < (uint32_t)a.rt) {
                        start = off[j * per + code];
                        len = (uint32_t)off[j * per + code + 1] - start;
                    }
                }
                const uint32_t s0 = __shfl_sync(0xFFFFFFFFu, start, 0), s1 = __shfl_sync(0xFFFFFFFFu, start, 1);
                const uint32_t s2 = __shfl_sync(0xFFFFFFFFu, start, 2), s3 = __shfl_sync(0xFFFFFFFFu, start, 3);
                const uint32_t c1 = __shfl_sync(0xFFFFFFFFu, len, 0);
                const uint32_t c2 = c1 + __shfl_sync(0xFFFFFFFFu, len, 1);
                const uint32_t c3 = c2 + __shfl_sync(0xFFFFFFFFu, len, 2);
                const uint32_t total = c3 + __shfl_sync(0xFFFFFFFFu, len, 3);
                for (uint32_t base = 0; base < total; base += 32) {
                    const uint32_t i = base + lane;
                    if (i < total) {
                        const bool g1 = i >= c1, g2 = i >= c2, g3 = i >= c3;
                        const uint32_t k = (g1 ? 1u : 0u) + (g2 ? 1u : 0u) + (g3 ? 1u : 0u);
                        const uint32_t idx = g3 ? s3 + (i - c3) : (g2 ? s2 + (i - c2) : (g1 ? s1 + (i - c1) : s0 + i));
                        const uint4 it = items[idx];
                        const uint4 qe = qent[j * QT + q_base + k];
                        float* ap = acc + (size_t)(q_base + k) * PT + (it.x & 0xFFFFu);
                        float te = fma_hh<true>(it.x, qe.x, 0.f);            // g = 0, 2, 4
                        te = fma_hh<true>(it.y, qe.y, te);
                        te = fma_hh<true>(it.z, qe.z, te);
                        float to = fma_hh<false>(it.y, qe.y, 0.f);           // g = 1, 3, 5
                        to = fma_hh<false>(it.z, qe.z, to);
                        to = fma_hh<false>(it.w, qe.w, to);
                        *ap = *ap + (te + to);
                    }
                }
                __syncwarp();            // another lane may update the same (q, p) in the next slice: keep the slices ordered
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        consumer_barrier();                                      // rows go back to the passage threads
        if (a.store) {
            float* o = a.out + (((size_t)qt * a.n_tiles + t) * QT) * PT + p;
#pragma unroll 8
            for (int q = 0; q < QT; ++q) o[(size_t)q * PT] = acc[q * PT + p];
        }
        consumer_barrier();
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
static uint16_t f2h(float f) { __half x = __float2half_rn(f); uint16_t h; memcpy(&h, &x, 2); return h; }

int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 37888, S = 128, rt = 39, n_queries = 256;
    const int n_tiles = (rows + PT - 1) / PT, n_chunks = S / SC, n_qtiles = n_queries / QT;
    // ---- synthetic encoded data (delade_cls recipe: |N(0,0.2)| values, 30 % empty slices, codes uniform in [0, 39)) ----
    std::vector<uint8_t> ccode((size_t)n_tiles * PT * S, 0xFF), qcode((size_t)n_queries * S, 0xFF);
    std::vector<uint16_t> cval((size_t)n_tiles * PT * S * G, 0), qval((size_t)n_queries * S * G, 0);
    auto fill = [&](std::vector<uint8_t>& code, std::vector<uint16_t>& val, size_t n_rows_real, size_t n_rows_alloc) {
        for (size_t r = 0; r < n_rows_alloc; ++r)
            for (int s = 0; s < S; ++s) {
                if (r >= n_rows_real || rnd() % 10 < 3) continue;
                code[r * S + s] = (uint8_t)(rnd() % rt);
                for (int g = 0; g < G; ++g) val[(r * S + s) * G + g] = f2h(0.01f + (float)(rnd() % 1000) / 2500.0f);
            }
    };
    fill(ccode, cval, (size_t)rows, (size_t)n_tiles * PT);
    fill(qcode, qval, (size_t)n_queries, (size_t)n_queries);
    // ---- tile layout (host) ----
    const int hdr_bytes = (SC * (rt + 1) * 2 + 15) / 16 * 16;
    std::vector<uint8_t> lexp;
    std::vector<unsigned long long> blk_off((size_t)n_tiles * n_chunks);
    std::vector<uint32_t> blk_bytes((size_t)n_tiles * n_chunks);
    uint32_t max_block = 0;
    for (int t = 0; t < n_tiles; ++t)
        for (int c = 0; c < n_chunks; ++c) {
            std::vector<uint16_t> off((size_t)hdr_bytes / 2, 0);
            std::vector<uint32_t> items;
            uint32_t n_items = 0;
            for (int j = 0; j < SC; ++j) {
                const int s = c * SC + j;
                std::vector<uint32_t> keys;                      // (code << 16) | passage: sorted by (code, passage)
                for (int pp = 0; pp < PT; ++pp) {
                    const uint8_t code = ccode[((size_t)t * PT + pp) * S + s];
                    if (code != 0xFF) keys.push_back(((uint32_t)code << 16) | (uint32_t)pp);
                }
                std::sort(keys.begin(), keys.end());
                size_t ki = 0;
                for (int code = 0; code <= rt; ++code) {
                    off[(size_t)j * (rt + 1) + code] = (uint16_t)(n_items + ki);
                    while (code < rt && ki < keys.size() && (int)(keys[ki] >> 16) == code) ++ki;
                }
                for (uint32_t key : keys) {
                    const uint32_t pp = key & 0xFFFFu;
                    const uint16_t* v = &cval[(((size_t)t * PT + pp) * S + s) * G];
                    items.push_back(pp | ((uint32_t)v[0] << 16));
                    items.push_back((uint32_t)v[1] | ((uint32_t)v[2] << 16));
                    items.push_back((uint32_t)v[3] | ((uint32_t)v[4] << 16));
                    items.push_back((uint32_t)v[5]);
                }
                n_items += (uint32_t)keys.size();
            }
            const size_t b = (size_t)t * n_chunks + c;
            blk_off[b] = lexp.size();
            blk_bytes[b] = (uint32_t)hdr_bytes + n_items * 16u;
            max_block = std::max(max_block, blk_bytes[b]);
            lexp.insert(lexp.end(), (const uint8_t*)off.data(), (const uint8_t*)off.data() + hdr_bytes);
            lexp.insert(lexp.end(), (const uint8_t*)items.data(), (const uint8_t*)items.data() + (size_t)n_items * 16);
        }
    std::vector<uint32_t> qblocks((size_t)n_qtiles * n_chunks * QBLOCK / 4, 0);
    for (int qt = 0; qt < n_qtiles; ++qt)
        for (int c = 0; c < n_chunks; ++c)
            for (int j = 0; j < SC; ++j)
                for (int q = 0; q < QT; ++q) {
                    const int s = c * SC + j, qq = qt * QT + q;
                    const uint16_t* v = &qval[((size_t)qq * S + s) * G];
                    const uint8_t code = qcode[(size_t)qq * S + s];
                    uint32_t* e = &qblocks[(((size_t)qt * n_chunks + c) * QBLOCK) / 4 + ((size_t)j * QT + q) * 4];
                    e[0] = (code == 0xFF ? 0xFFFFu : (uint32_t)code) | ((uint32_t)v[0] << 16);
                    e[1] = (uint32_t)v[1] | ((uint32_t)v[2] << 16);
                    e[2] = (uint32_t)v[3] | ((uint32_t)v[4] << 16);
                    e[3] = (uint32_t)v[5];
                }
    printf("layout: %.0f bytes per row (per-passage tile layout: %d), largest block %u bytes\n", (double)lexp.size() / rows, S * (1 + 2 * G),
           max_block);
    // ---- device ----
    Args a{};
    uint8_t* d_lexp; unsigned long long* d_off; uint32_t* d_bytes; uint8_t* d_q; float* d_out;
    const size_t out_elems = (size_t)n_qtiles * n_tiles * QT * PT;
    CK(cudaMalloc(&d_lexp, lexp.size())); CK(cudaMemcpy(d_lexp, lexp.data(), lexp.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_off, blk_off.size() * 8)); CK(cudaMemcpy(d_off, blk_off.data(), blk_off.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_bytes, blk_bytes.size() * 4)); CK(cudaMemcpy(d_bytes, blk_bytes.data(), blk_bytes.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_q, qblocks.size() * 4)); CK(cudaMemcpy(d_q, qblocks.data(), qblocks.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, out_elems * 4)); CK(cudaMemset(d_out, 0xFF, out_elems * 4));
    a.lexp = d_lexp; a.blk_off = d_off; a.blk_bytes = d_bytes; a.qblocks = d_q; a.out = d_out;
    a.n_tiles = n_tiles; a.n_chunks = n_chunks; a.n_qtiles = n_qtiles; a.rt = rt; a.hdr_bytes = hdr_bytes;
    a.pblock_smem = (int)((max_block + 127) / 128 * 128);
    a.stage_bytes = a.pblock_smem + QBLOCK;
    const size_t smem = (size_t)QT * PT * 4 + (size_t)STAGES * a.stage_bytes;
    printf("shared memory per CTA: %zu bytes (%d stages of %d)\n", smem, STAGES, a.stage_bytes);
    if (smem > 226 * 1024) { fprintf(stderr, "stage too large for %d stages\n", STAGES); return 1; }
    CK(cudaFuncSetAttribute(k1t_v7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int per_q = prop.multiProcessorCount / n_qtiles;
    if (per_q > n_tiles) per_q = n_tiles;
    const unsigned grid = (unsigned)(per_q * n_qtiles);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int store = 0; store < 2; ++store) {                   // walk only first, then with the accumulator store (checked below)
        a.store = store;
        std::vector<float> ms;
        for (int i = 0; i < 21; ++i) {
            cudaEventRecord(e0);
            k1t_v7_kernel<<<grid, THREADS, smem>>>(a);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
        }
        std::sort(ms.begin(), ms.end());
        printf("k1t_v7_kernel (%s): %d rows x %d queries, median %.1f us, min %.1f us per launch (grid %u)\n",
               store ? "with accumulator store" : "walk only", rows, n_queries, ms[10] * 1e3, ms[0] * 1e3, grid);
    }
    // ---- check against the host (same fp32 operation order: slices ascending, te/to chains, acc + (te + to)) ----
    std::vector<float> out(out_elems);
    CK(cudaMemcpy(out.data(), d_out, out_elems * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0, checked = 0, matches = 0;
    for (int qq = 0; qq < n_queries; qq += 7)
        for (int r = 0; r < n_tiles * PT; r += 3) {
            float acc = 0.f;
            for (int s = 0; s < S; ++s) {
                const uint8_t cq = qcode[(size_t)qq * S + s], cp = ccode[(size_t)r * S + s];
                if (cq == 0xFF || cp != cq) continue;
                const uint16_t* qv = &qval[((size_t)qq * S + s) * G];
                const uint16_t* pv = &cval[((size_t)r * S + s) * G];
                float te = 0.f, to = 0.f;
                for (int g = 0; g < G; g += 2) te = fmaf(h2f(pv[g]), h2f(qv[g]), te);
                for (int g = 1; g < G; g += 2) to = fmaf(h2f(pv[g]), h2f(qv[g]), to);
                acc = acc + (te + to);
                ++matches;
            }
            const float got = out[(((size_t)(qq / QT) * n_tiles + r / PT) * QT + qq % QT) * PT + r % PT];
            ++checked;
            if (memcmp(&got, &acc, 4) != 0 && !(got == acc)) { if (bad < 5) printf("mismatch q %d row %d: got %g want %g\n", qq, r, got, acc); ++bad; }
        }
    printf("checked %zu (query, passage) pairs (%zu matches): %zu mismatches -> %s\n", checked, matches, bad, bad ? "FAIL" : "OK (bit-exact)");
    return bad ? 2 : 0;
}
