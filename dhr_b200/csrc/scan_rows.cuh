// scan_rows.cu -- K1: fused GIP row scan (lexical match-and-MAC + dense tail + threshold filter).
//
// Replaces the per-query loop body of castorini/dhr retrieval/gip_retrieval.py:117-126
// (eq-mask -> multiply -> einsum -> topk) and :74-75 (dense-only) without materialising
// the [N, W] temporaries or a score vector: every row is read from HBM once per group of
// QB queries, scored in fp32 and appended to the per-query candidate list only if it beats
// the running admission threshold tau (strictly).
//
// Thread mapping: one warp scores RB=2 rows at a time.  Lane l owns lexical "units"
// l, l+32, ... of the row (a unit = lcm(G,8) fp16 values = a whole number of slices and of
// 16-byte vectors) and dense 16-byte chunks l, l+32, ...; partial sums are combined with a
// halving butterfly so that the 2*QB totals end up in distinct lanes.
//
// Two data paths with identical arithmetic:
//   variant 0  direct coalesced 128-bit ld.global.nc loads into registers
//   variant 1  TMA bulk copies (cp.async.bulk + mbarrier ring) of whole row tiles into
//              shared memory by a producer warp; consumer warps read shared memory
#pragma once
#include <type_traits>

#include "internal.h"

namespace dhr {

constexpr int kConsumerWarps = 8;
constexpr int kRB = 2;                       // rows per warp step

__host__ __device__ constexpr int unit_halves(int G) {
    return (G % 8 == 0) ? G : (G % 4 == 0) ? 2 * G : (G % 2 == 0) ? 4 * G : 8 * G;   // lcm(G, 8) for G <= 8
}

// ---- loads -------------------------------------------------------------------------------
struct GlobalSrc {
    static __device__ __forceinline__ uint4 ld16(const void* p) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ uint2 ld8(const void* p) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ uint32_t ld4(const void* p) { return __ldg((const uint32_t*)p); }
    static __device__ __forceinline__ uint32_t ld2(const void* p) { return __ldg((const uint16_t*)p); }
    static __device__ __forceinline__ uint32_t ld1(const void* p) { return __ldg((const uint8_t*)p); }
};
struct SharedSrc {
    static __device__ __forceinline__ uint4 ld16(const void* p) { return *(const uint4*)p; }
    static __device__ __forceinline__ uint2 ld8(const void* p) { return *(const uint2*)p; }
    static __device__ __forceinline__ uint32_t ld4(const void* p) { return *(const uint32_t*)p; }
    static __device__ __forceinline__ uint32_t ld2(const void* p) { return *(const uint16_t*)p; }
    static __device__ __forceinline__ uint32_t ld1(const void* p) { return *(const uint8_t*)p; }
};

// load NBYTES (1,2,4,8,16) of codes into 32-bit words
template <typename Src, int NBYTES>
__device__ __forceinline__ void load_codes(const uint8_t* p, uint32_t (&w)[(NBYTES + 3) / 4]) {
    if constexpr (NBYTES == 16) { uint4 v = Src::ld16(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    else if constexpr (NBYTES == 8) { uint2 v = Src::ld8(p); w[0] = v.x; w[1] = v.y; }
    else if constexpr (NBYTES == 4) { w[0] = Src::ld4(p); }
    else if constexpr (NBYTES == 2) { w[0] = Src::ld2(p); }
    else { w[0] = Src::ld1(p); }
}

template <typename CodeT>
__device__ __forceinline__ bool code_equal(const uint32_t* a, const uint32_t* b, int j) {
    if constexpr (sizeof(CodeT) == 1) {
        return (((a[j >> 2] ^ b[j >> 2]) >> (8 * (j & 3))) & 0xFFu) == 0u;
    } else {
        return (((a[j >> 1] ^ b[j >> 1]) >> (16 * (j & 1))) & 0xFFFFu) == 0u;
    }
}

// one lexical unit of one (row, query) pair: acc += sum_slices [codes equal] * dot_G
template <int G, typename CodeT, bool QF32>
__device__ __forceinline__ float unit_dot(const uint32_t* pw, const uint32_t* pc, const uint32_t* qw_or_f,
                                          const uint32_t* qc, float acc, bool masked) {
    constexpr int UH = unit_halves(G);
    constexpr int US = UH / G;
#pragma unroll
    for (int j = 0; j < US; ++j) {
        float t = acc;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int e = j * G + g;
            if constexpr (QF32) {
                const float qf = __uint_as_float(qw_or_f[e]);
                const float pf = (e & 1) ? half_hi_to_float(pw[e >> 1]) : half_lo_to_float(pw[e >> 1]);
                t = fmaf(pf, qf, t);
            } else {
                t = (e & 1) ? fma_h_hi(pw[e >> 1], qw_or_f[e >> 1], t) : fma_h_lo(pw[e >> 1], qw_or_f[e >> 1], t);
            }
        }
        const bool m = !masked || code_equal<CodeT>(pc, qc, j);
        acc = m ? t : acc;
    }
    return acc;
}

template <bool QF32>
__device__ __forceinline__ float chunk_dot(const uint4& p, const uint32_t* q, float acc) {
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (QF32) {
            acc = fmaf(half_lo_to_float(pw[i]), __uint_as_float(q[2 * i]), acc);
            acc = fmaf(half_hi_to_float(pw[i]), __uint_as_float(q[2 * i + 1]), acc);
        } else {
            acc = fma_h_lo(pw[i], q[i], acc);
            acc = fma_h_hi(pw[i], q[i], acc);
        }
    }
    return acc;
}

// Sum V values across the warp; lane l receives the total of value index (l >> (5 - log2 V)).
template <int V>
__device__ __forceinline__ float warp_reduce_multi(float (&v)[V], int lane) {
    static_assert(V == 1 || V == 2 || V == 4 || V == 8 || V == 16 || V == 32, "V must be a power of two");
    int off = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = upper ? v[i + n / 2] : v[i];
            const float send = upper ? v[i] : v[i + n / 2];
            v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
        }
        off >>= 1;
    }
#pragma unroll
    for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], off);
    return v[0];
}

template <int V> struct Log2 { static constexpr int value = 1 + Log2<V / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// shared-memory query block: q_lex[QB][D_pad] | q_code[QB][S_pad] | q_dns[QB][C_pad]
template <typename CodeT, int QB, bool QF32>
struct QueryBlock {
    using QT = typename std::conditional<QF32, float, __half>::type;
    const uint8_t* lex; const uint8_t* code; const uint8_t* dns;
    int lex_stride, code_stride, dns_stride;   // bytes
    static __host__ __device__ size_t bytes(int S_pad, int D_pad, int C_pad) {
        return (size_t)QB * ((size_t)D_pad * sizeof(QT) + (size_t)S_pad * sizeof(CodeT) + (size_t)C_pad * sizeof(QT));
    }
};

// cooperative copy of the group's queries into shared memory; queries beyond `nq` become
// all-zero / never-matching so they score exactly 0 and are never appended.
template <typename CodeT, int QB, bool QF32>
__device__ __forceinline__ QueryBlock<CodeT, QB, QF32> stage_queries(uint8_t* smem, const ScanArgs& a, int q0, int nq,
                                                                     int tid, int nthreads) {
    using QBk = QueryBlock<CodeT, QB, QF32>;
    using QT = typename QBk::QT;
    QBk qb;
    qb.lex_stride = a.D_pad * (int)sizeof(QT);
    qb.code_stride = a.S_pad * (int)sizeof(CodeT);
    qb.dns_stride = a.C_pad * (int)sizeof(QT);
    uint8_t* lex = smem;
    uint8_t* code = lex + (size_t)QB * qb.lex_stride;
    uint8_t* dns = code + (size_t)QB * qb.code_stride;
    qb.lex = lex; qb.code = code; qb.dns = dns;
    const uint8_t* g_lex = (const uint8_t*)a.q_lex + (size_t)q0 * qb.lex_stride;
    const uint8_t* g_code = (const uint8_t*)a.q_code + (size_t)q0 * qb.code_stride;
    const uint8_t* g_dns = (const uint8_t*)a.q_dns + (size_t)q0 * qb.dns_stride;
    const int lex_v = QB * qb.lex_stride / 16, code_v = QB * qb.code_stride / 16, dns_v = QB * qb.dns_stride / 16;
    const int lex_valid = nq * qb.lex_stride / 16, code_valid = nq * qb.code_stride / 16, dns_valid = nq * qb.dns_stride / 16;
    const uint32_t nm = CodeTraits<CodeT>::kNoMatch;
    const uint32_t nm32 = sizeof(CodeT) == 1 ? nm * 0x01010101u : nm * 0x00010001u;
    for (int i = tid; i < lex_v; i += nthreads)
        ((uint4*)lex)[i] = i < lex_valid ? ((const uint4*)g_lex)[i] : make_uint4(0, 0, 0, 0);
    for (int i = tid; i < code_v; i += nthreads)
        ((uint4*)code)[i] = i < code_valid ? ((const uint4*)g_code)[i] : make_uint4(nm32, nm32, nm32, nm32);
    for (int i = tid; i < dns_v; i += nthreads)
        ((uint4*)dns)[i] = i < dns_valid ? ((const uint4*)g_dns)[i] : make_uint4(0, 0, 0, 0);
    return qb;
}

// Score rows r and r+1 (r+1 optional) against the QB queries in shared memory; append winners.
template <typename Src, int G, typename CodeT, int QB, bool QF32>
__device__ __forceinline__ void score_row_pair(const uint8_t* lexv0, const uint8_t* lexi0, const uint8_t* dns0,
                                               int lexv_stride, int lexi_stride, int dns_stride, bool have_second,
                                               const QueryBlock<CodeT, QB, QF32>& qb, const ScanArgs& a, int lane,
                                               long long row, int q0, int nq, float my_tau) {
    constexpr int UH = unit_halves(G);
    constexpr int UW = UH / 2;                       // 32-bit words of fp16 values per unit
    constexpr int UV = UH / 8;                       // 16-byte vectors per unit
    constexpr int US = UH / G;                       // slices per unit
    constexpr int CB = US * (int)sizeof(CodeT);      // code bytes per unit
    constexpr int CW = (CB + 3) / 4;
    constexpr int QW = QF32 ? UH : UW;               // query words per unit
    constexpr int V = kRB * QB;

    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;

    const int second = have_second ? 1 : 0;
    const bool masked = a.masked != 0;

    for (int u = lane; u < a.n_units; u += 32) {
        uint32_t pw[kRB][UW];
        uint32_t pc[kRB][CW];
#pragma unroll
        for (int rb = 0; rb < kRB; ++rb) {
            const int ro = rb * second;
            const uint8_t* pv = lexv0 + (size_t)ro * lexv_stride + (size_t)u * (UH * 2);
#pragma unroll
            for (int v = 0; v < UV; ++v) {
                const uint4 x = Src::ld16(pv + 16 * v);
                pw[rb][4 * v] = x.x; pw[rb][4 * v + 1] = x.y; pw[rb][4 * v + 2] = x.z; pw[rb][4 * v + 3] = x.w;
            }
            load_codes<Src, CB>(lexi0 + (size_t)ro * lexi_stride + (size_t)u * CB, pc[rb]);
        }
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            uint32_t qw[QW];
            uint32_t qc[CW];
            const uint8_t* ql = qb.lex + (size_t)q * qb.lex_stride + (size_t)u * (QW * 4);
#pragma unroll
            for (int v = 0; v < QW / 4; ++v) {
                const uint4 x = *(const uint4*)(ql + 16 * v);
                qw[4 * v] = x.x; qw[4 * v + 1] = x.y; qw[4 * v + 2] = x.z; qw[4 * v + 3] = x.w;
            }
            load_codes<SharedSrc, CB>(qb.code + (size_t)q * qb.code_stride + (size_t)u * CB, qc);
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb)
                acc[rb * QB + q] = unit_dot<G, CodeT, QF32>(pw[rb], pc[rb], qw, qc, acc[rb * QB + q], masked);
        }
    }

    for (int c = lane; c < a.n_chunks; c += 32) {
        uint4 pv[kRB];
#pragma unroll
        for (int rb = 0; rb < kRB; ++rb) pv[rb] = Src::ld16(dns0 + (size_t)(rb * second) * dns_stride + (size_t)c * 16);
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            uint32_t qd[QF32 ? 8 : 4];
            const uint8_t* qp = qb.dns + (size_t)q * qb.dns_stride + (size_t)c * (QF32 ? 32 : 16);
            const uint4 x = *(const uint4*)qp;
            qd[0] = x.x; qd[1] = x.y; qd[2] = x.z; qd[3] = x.w;
            if constexpr (QF32) {
                const uint4 y = *(const uint4*)(qp + 16);
                qd[4] = y.x; qd[5] = y.y; qd[6] = y.z; qd[7] = y.w;
            }
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) acc[rb * QB + q] = chunk_dot<QF32>(pv[rb], qd, acc[rb * QB + q]);
        }
    }

    const float total = warp_reduce_multi<V>(acc, lane);
    constexpr int SH = 5 - Log2<V>::value;
    const int vi = lane >> SH;                        // value index this lane holds: rb * QB + q
    const int rb = vi / QB, q = vi % QB;
    if ((lane & ((1 << SH) - 1)) == 0 && q < nq && (rb == 0 || have_second)) {
        const float s = total + 0.0f;                 // -0.0 -> +0.0 so that the key order equals the float order
        if (s > my_tau) {
            const int slot = q0 + q;
            const uint32_t pos = atomicAdd(a.cnt + slot, 1u);
            if (pos < (uint32_t)a.cap) {
                a.cand_score[(size_t)slot * a.cap + pos] = s;
                a.cand_row[(size_t)slot * a.cap + pos] = (int32_t)(row + rb);
            }
        }
    }
}

// tau of the query this lane reports for (see the lane -> value mapping above)
template <int QB>
__device__ __forceinline__ float lane_tau(const ScanArgs& a, int lane, int q0, int nq) {
    constexpr int V = kRB * QB;
    constexpr int SH = 5 - Log2<V>::value;
    const int q = (lane >> SH) % QB;
    return q < nq ? a.tau[q0 + q] : INFINITY;
}

// ---------------------------------------------------------------------------------------------
// variant 0: direct loads
// ---------------------------------------------------------------------------------------------
template <int G, typename CodeT, int QB, bool QF32>
__global__ void __launch_bounds__(kConsumerWarps * 32, 2) gip_scan_direct(const __grid_constant__ ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int group = blockIdx.x % a.n_groups;
    const long long rblk = blockIdx.x / a.n_groups;
    const int q0 = group * QB;
    const int nq = min(QB, a.n_queries - q0);
    const auto qb = stage_queries<CodeT, QB, QF32>(smem, a, q0, nq, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float my_tau = lane_tau<QB>(a, lane, q0, nq);
    const long long row0 = a.row_begin + rblk * a.rows_per_cta;
    const long long row1 = min(row0 + (long long)a.rows_per_cta, a.row_end);
    const int lexv_stride = a.D_pad * 2, lexi_stride = a.S_pad * (int)sizeof(CodeT), dns_stride = a.C_pad * 2;
    for (long long r = row0 + warp * kRB; r < row1; r += kConsumerWarps * kRB) {
        score_row_pair<GlobalSrc, G, CodeT, QB, QF32>(
            (const uint8_t*)a.lexv + (size_t)r * lexv_stride, a.lexi + (size_t)r * lexi_stride,
            (const uint8_t*)a.dns + (size_t)r * dns_stride, lexv_stride, lexi_stride, dns_stride, r + 1 < row1, qb, a,
            lane, r, q0, nq, my_tau);
    }
}

// ---------------------------------------------------------------------------------------------
// variant 1: TMA bulk staging.  grid = n_groups * ctas_per_group persistent CTAs; each CTA walks
// tiles of `tile_rows` rows with stride ctas_per_group through an n_stages-deep smem ring.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxStages = 8;

template <int G, typename CodeT, int QB, bool QF32>
__global__ void __launch_bounds__((kConsumerWarps + 1) * 32, 1) gip_scan_tma(const __grid_constant__ ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];

    const int group = blockIdx.x % a.n_groups;
    const int cta_in_group = blockIdx.x / a.n_groups;
    const int ctas_per_group = gridDim.x / a.n_groups;
    const int q0 = group * QB;
    const int nq = min(QB, a.n_queries - q0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int lexv_stride = a.D_pad * 2, lexi_stride = a.S_pad * (int)sizeof(CodeT), dns_stride = a.C_pad * 2;
    const size_t qbytes = (QueryBlock<CodeT, QB, QF32>::bytes(a.S_pad, a.D_pad, a.C_pad) + 127) / 128 * 128;
    const size_t lexv_tile = ((size_t)a.tile_rows * lexv_stride + 127) / 128 * 128;
    const size_t lexi_tile = ((size_t)a.tile_rows * lexi_stride + 127) / 128 * 128;
    const size_t dns_tile = ((size_t)a.tile_rows * dns_stride + 127) / 128 * 128;
    const size_t stage_bytes = lexv_tile + lexi_tile + dns_tile;
    uint8_t* stages = smem + qbytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kConsumerWarps); }
        mbar_fence_init();
    }
    const auto qb = stage_queries<CodeT, QB, QF32>(smem, a, q0, nq, threadIdx.x, blockDim.x);
    __syncthreads();

    const long long n_rows = a.row_end - a.row_begin;
    const long long n_tiles = (n_rows + a.tile_rows - 1) / a.tile_rows;

    if (warp == kConsumerWarps) {
        // ===== producer warp: one elected lane feeds the ring =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long long t = cta_in_group; t < n_tiles; t += ctas_per_group) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const long long r0 = a.row_begin + t * a.tile_rows;
                const int rows = (int)min((long long)a.tile_rows, a.row_end - r0);
                uint8_t* dst = stages + (size_t)s * stage_bytes;
                const uint32_t b0 = (uint32_t)rows * lexv_stride, b1 = (uint32_t)rows * lexi_stride, b2 = (uint32_t)rows * dns_stride;
                mbar_arrive_expect_tx(&full_bar[s], b0 + b1 + b2);
                if (b0) bulk_g2s(dst, (const uint8_t*)a.lexv + (size_t)r0 * lexv_stride, b0, &full_bar[s]);
                if (b1) bulk_g2s(dst + lexv_tile, a.lexi + (size_t)r0 * lexi_stride, b1, &full_bar[s]);
                if (b2) bulk_g2s(dst + lexv_tile + lexi_tile, (const uint8_t*)a.dns + (size_t)r0 * dns_stride, b2, &full_bar[s]);
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== consumer warps =====
        const float my_tau = lane_tau<QB>(a, lane, q0, nq);
        int s = 0; uint32_t ph = 0;
        for (long long t = cta_in_group; t < n_tiles; t += ctas_per_group) {
            mbar_wait(&full_bar[s], ph);
            const long long r0 = a.row_begin + t * a.tile_rows;
            const int rows = (int)min((long long)a.tile_rows, a.row_end - r0);
            const uint8_t* base = stages + (size_t)s * stage_bytes;
            for (int lr = warp * kRB; lr < rows; lr += kConsumerWarps * kRB) {
                score_row_pair<SharedSrc, G, CodeT, QB, QF32>(
                    base + (size_t)lr * lexv_stride, base + lexv_tile + (size_t)lr * lexi_stride,
                    base + lexv_tile + lexi_tile + (size_t)lr * dns_stride, lexv_stride, lexi_stride, dns_stride,
                    lr + 1 < rows, qb, a, lane, r0 + lr, q0, nq, my_tau);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == a.n_stages) { s = 0; ph ^= 1u; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// rerank: exact GIP on listed rows (gip_retrieval.py:142-150, :205-215), QB = 1 arithmetic of K1.
// grid (ceil(M / warps), nq): block y scores candidates of in-flight query slot y; every valid
// candidate is appended (tau = -inf), K3 then selects.
// ---------------------------------------------------------------------------------------------
template <int G, typename CodeT, bool QF32>
__global__ void __launch_bounds__(kConsumerWarps * 32, 2)
gip_rerank_kernel(const __grid_constant__ ScanArgs a, const long long* __restrict__ cand, int n_cand, long long n_rows) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int slot = blockIdx.y;
    const auto qb = stage_queries<CodeT, 1, QF32>(smem, a, slot, 1, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kConsumerWarps + warp;
    if (j >= n_cand) return;
    const long long r = cand[(size_t)slot * n_cand + j];
    if (r < 0 || r >= n_rows) return;
    const int lexv_stride = a.D_pad * 2, lexi_stride = a.S_pad * (int)sizeof(CodeT), dns_stride = a.C_pad * 2;
    score_row_pair<GlobalSrc, G, CodeT, 1, QF32>((const uint8_t*)a.lexv + (size_t)r * lexv_stride,
                                                 a.lexi + (size_t)r * lexi_stride,
                                                 (const uint8_t*)a.dns + (size_t)r * dns_stride, lexv_stride, lexi_stride,
                                                 dns_stride, false, qb, a, lane, r, slot, 1, -INFINITY);
}

template <int G, typename CodeT>
static inline int launch_rerank_g(const ScanArgs& a, bool q_f32, const long long* d_cand, int n_cand, long long n_rows, cudaStream_t st) {
    dim3 grid((unsigned)((n_cand + kConsumerWarps - 1) / kConsumerWarps), (unsigned)a.n_queries);
    if (q_f32) {
        const size_t smem = QueryBlock<CodeT, 1, true>::bytes(a.S_pad, a.D_pad, a.C_pad);
        auto kern = gip_rerank_kernel<G, CodeT, true>;
        if (smem > 48 * 1024) DHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kConsumerWarps * 32, smem, st>>>(a, d_cand, n_cand, n_rows);
    } else {
        const size_t smem = QueryBlock<CodeT, 1, false>::bytes(a.S_pad, a.D_pad, a.C_pad);
        auto kern = gip_rerank_kernel<G, CodeT, false>;
        if (smem > 48 * 1024) DHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kConsumerWarps * 32, smem, st>>>(a, d_cand, n_cand, n_rows);
    }
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
template <int G, typename CodeT, int QB, bool QF32>
static int launch_one(const dhr_index* h, const ScanArgs& a, int variant, cudaStream_t st) {
    const long long n_rows = a.row_end - a.row_begin;
    if (n_rows <= 0 || a.n_queries <= 0) return DHR_OK;
    if (variant == 1) {
        const size_t smem = scan_tma_smem_bytes(h->g, QB, QF32, a.tile_rows, a.n_stages);
        auto kern = gip_scan_tma<G, CodeT, QB, QF32>;
        DHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long n_tiles = (n_rows + a.tile_rows - 1) / a.tile_rows;
        long long per_group = h->num_sms / a.n_groups;
        if (per_group < 1) per_group = 1;
        if (per_group > n_tiles) per_group = n_tiles;
        const unsigned grid = (unsigned)(per_group * a.n_groups);
        kern<<<grid, (kConsumerWarps + 1) * 32, smem, st>>>(a);
    } else {
        const size_t smem = QueryBlock<CodeT, QB, QF32>::bytes(a.S_pad, a.D_pad, a.C_pad);
        auto kern = gip_scan_direct<G, CodeT, QB, QF32>;
        if (smem > 48 * 1024) DHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long blocks = (n_rows + a.rows_per_cta - 1) / a.rows_per_cta;
        kern<<<(unsigned)(blocks * a.n_groups), kConsumerWarps * 32, smem, st>>>(a);
    }
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

template <int G, typename CodeT>
static int launch_qb(const dhr_index* h, const ScanArgs& a, int qb, bool q_f32, int variant, cudaStream_t st) {
#define DHR_QB_CASE(QBV)                                                                      \
    case QBV:                                                                                 \
        return q_f32 ? launch_one<G, CodeT, QBV, true>(h, a, variant, st)                     \
                     : launch_one<G, CodeT, QBV, false>(h, a, variant, st);
    switch (qb) {
        DHR_QB_CASE(1)
        DHR_QB_CASE(2)
        DHR_QB_CASE(4)
        DHR_QB_CASE(8)
        default: return DHR_ERR_INVALID;
    }
#undef DHR_QB_CASE
}


// per-G entry points, one translation unit each (scan_inst.cu compiled with -DDHR_G=<G>)
template <int G> int scan_entry(const dhr_index* h, const ScanArgs& a, int qb, bool q_f32, int variant, cudaStream_t st);
template <int G> int rerank_entry(const dhr_index* h, const ScanArgs& a, bool q_f32, const long long* d_cand, int n_cand, cudaStream_t st);

}  // namespace dhr
