// dense_tile.cu -- K2: dense [CLS] x corpus block on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the dense part of the masked row dot of castorini/dhr retrieval/gip_retrieval.py:119-120
// (the always-matching [CLS] tail, :110-113) and IP_retrieval's einsum (:74) for a tile of 64
// queries at once: D[128 passages x 64 queries] (fp32, TMEM) = A[128 x C] (corpus rows, fp16,
// K-major, TMA 128B-swizzled tiles streamed through a ring) x B[64 x C]^T (queries, resident in
// shared memory for the whole launch).  One thread issues tcgen05.mma; accumulators are double
// buffered in TMEM and drained with tcgen05.ld by four epilogue warps that either
//   - apply the strict admission threshold and append candidates (dense-only index), or
//   - write the tile to the L2-resident scratch consumed by the lexical kernel K1t (hybrid index).
// Warp roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 epilogue.
//
// Two variants.  `dense_tile_kernel` (SS: both operands in shared memory, passages = M) measured 82 cycles per
// M128 x N64 x K16 instruction -- the 4 KiB A-operand read from shared memory per instruction is amortised over only
// 64 columns.  `dense_tile_ts_kernel` (TS: A operand in TMEM) swaps the roles: the 128 in-flight queries of a CTA are
// the M dimension and live in tensor memory for the whole launch (row = TMEM lane, fp16 pairs along columns, written
// once with tcgen05.st), the corpus is the N dimension streamed as 128-passage x 64-column 16 KiB stages through a
// TMA ring, so an instruction reads only its 4 KiB B slice from shared memory and every corpus byte is reused by twice
// as many queries.  TS needs C_pad <= 768 (384 TMEM columns of A + 128 accumulator columns = 512).
#include <cuda.h>

#include "internal.h"

namespace dhr {

// Timeline hooks of the micro-benchmark harness (tools/k2_micro.cu defines DHR_K2_TRACE); compiled out of the library.
#ifdef DHR_K2_TRACE
__device__ long long g_k2_trace[8][64];
__device__ int g_k2_dbg;       // experiment bits (TS kernel): 4 = no epilogue stores, 8 = free-running MMA (no corpus loads, no stage barriers), 16 = no MMAs (corpus stream + barriers only)
#define K2_TRACE(role, idx) do { if (blockIdx.x == 0 && (idx) < 64) g_k2_trace[role][idx] = clock64(); } while (0)
#define K2_DBG() g_k2_dbg
#else
#define K2_TRACE(role, idx) do { } while (0)
#define K2_DBG() 0
#endif

constexpr int kDT_M = 128;             // passages per tile (UMMA M)
constexpr int kDT_N = 64;              // queries per tile (UMMA N)
constexpr int kDT_KB = 64;             // fp16 elements per K block = 128 bytes = one swizzle atom row
constexpr int kDT_ABytes = kDT_M * kDT_KB * 2;   // 16 KiB per A stage
constexpr int kDT_BBytes = kDT_N * kDT_KB * 2;   //  8 KiB per resident B block
constexpr int kDT_Threads = 192;
constexpr int kDT_MaxStages = 8;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// the same load delivered to the same shared-memory offset (and signalled on the same mbarrier offset) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// one lane of a converged warp; the surrounding code stays warp-uniform so tcgen05 / TMA operands live in uniform registers
// (a divergent `if (lane == 0)` issuer makes ptxas wrap every UTCHMMA in an R2UR waterfall loop: ~140 cycles per instruction)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile whose rows are 128 bytes apart (8-row groups 1024 bytes apart):
// start address >> 4 | LBO = 1 (ignored for swizzled K-major) | SBO = 1024 >> 4 | version 1 | SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16: fp16 A and B, fp32 accumulate, both K-major, M = 128, N = 64
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct DenseTileArgs {
    long long row_begin, row_end;      // rows of this launch; row_begin is a multiple of 128 relative to tile_row0
    long long tile_row0;               // first row of tile 0
    int n_tiles;                       // 128-row tiles in the launch
    int n_kblocks;                     // ceil(C_pad / 64)
    int n_stages;
    int n_qtiles;                      // 64-query tiles in flight
    int n_queries;                     // valid queries in flight (slots)
    int mode;                          // 0 = filter + append, 1 = write scratch
    float* scratch;                    // [row - tile_row0][scratch_slots]  (mode 1)
    long long scratch_slots;           // query slots per scratch row (multiple of 64)
    float* tau; uint32_t* cnt; float* cand_score; int32_t* cand_row; int cap;
};

__global__ void __launch_bounds__(kDT_Threads, 1)
dense_tile_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const DenseTileArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kDT_MaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kDT_MaxStages];
    __shared__ __align__(8) uint64_t b_bar;
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float tau_s[kDT_N];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % a.n_qtiles;
    const int cta_in_q = blockIdx.x / a.n_qtiles;
    const int ctas_per_q = gridDim.x / a.n_qtiles;

    // SWIZZLE_128B operand tiles must sit on 1024-byte boundaries of the shared address space
    uint8_t* smem_b = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // n_kblocks x 8 KiB
    uint8_t* smem_a = smem_b + (size_t)a.n_kblocks * kDT_BBytes;       // n_stages x 16 KiB

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&b_bar, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        mbar_fence_init();
    }
    if (threadIdx.x < kDT_N) {
        const int q = qt * kDT_N + threadIdx.x;
        tau_s[threadIdx.x] = (q < a.n_queries && a.mode == 0) ? a.tau[q] : INFINITY;
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 2 * kDT_N);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer (whole warp, one elected lane issues) =====
        if (elect_one()) {
            mbar_arrive_expect_tx(&b_bar, (uint32_t)a.n_kblocks * kDT_BBytes);
            for (int kb = 0; kb < a.n_kblocks; ++kb)
                tma_load_2d(smem_b + (size_t)kb * kDT_BBytes, &tmap_b, &b_bar, kb * kDT_KB, qt * kDT_N);
        }
        __syncwarp();
        int s = 0; uint32_t ph = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q) {
            const int row0 = (int)(a.tile_row0 + (long long)t * kDT_M);
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[s], kDT_ABytes);
                    tma_load_2d(smem_a + (size_t)s * kDT_ABytes, &tmap_a, &full_bar[s], kb * kDT_KB, row0);
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp in uniform control flow, one elected lane issues) =====
        constexpr uint32_t idesc = umma_idesc_f16(kDT_M, kDT_N);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        mbar_wait(&b_bar, 0);
        tc_fence_after();
        int s = 0; uint32_t ph = 0;
        int i = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            const int buf = i & 1;
            mbar_wait(&tempty_bar[buf], (((uint32_t)i >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + (uint32_t)buf * kDT_N;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + (size_t)s * kDT_ABytes);
                const uint32_t b_addr = smem_u32(smem_b + (size_t)kb * kDT_BBytes);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kDT_KB / 16; ++k) {
                        umma_f16(d_tmem, umma_smem_desc(a_addr + k * 32), umma_smem_desc(b_addr + k * 32), idesc,
                                 (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);          // frees the A stage once these MMAs have read it
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(&tfull_bar[buf]);   // accumulator tile complete
            __syncwarp();
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int p = quarter * 32 + lane;               // row of the tile owned by this thread
        int i = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            const int buf = i & 1;
            mbar_wait(&tfull_bar[buf], ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * kDT_N;
            tmem_ld_32x32(taddr, v0);
            tmem_ld_32x32(taddr + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            const long long row = a.tile_row0 + (long long)t * kDT_M + p;
            const bool row_ok = row >= a.row_begin && row < a.row_end;
            if (a.mode == 1) {
                if (row_ok) {   // scratch[row][slot]: this thread's 64 consecutive query slots = 256 contiguous bytes
                    float4* dst = (float4*)(a.scratch + (size_t)(row - a.tile_row0) * a.scratch_slots + (size_t)qt * kDT_N);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        dst[q] = make_float4(__uint_as_float(v0[4 * q]), __uint_as_float(v0[4 * q + 1]), __uint_as_float(v0[4 * q + 2]),
                                             __uint_as_float(v0[4 * q + 3]));
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        dst[8 + q] = make_float4(__uint_as_float(v1[4 * q]), __uint_as_float(v1[4 * q + 1]), __uint_as_float(v1[4 * q + 2]),
                                                 __uint_as_float(v1[4 * q + 3]));
                }
            } else if (row_ok) {
#pragma unroll
                for (int q = 0; q < kDT_N; ++q) {
                    const float sc = __uint_as_float(q < 32 ? v0[q & 31] : v1[q & 31]) + 0.0f;
                    if (sc > tau_s[q]) {
                        const int slot = qt * kDT_N + q;
                        const uint32_t pos = atomicAdd(a.cnt + slot, 1u);
                        if (pos < (uint32_t)a.cap) {
                            a.cand_score[(size_t)slot * a.cap + pos] = sc;
                            a.cand_row[(size_t)slot * a.cap + pos] = (int32_t)row;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * kDT_N);
}


// ---- 32-bit shared-address forms of the barrier / TMA / commit wrappers ------------------------------------------------------
// The per-stage loops of the TMA producer and of the MMA issuer run in ONE warp each and are latency chains of uniform-datapath
// instructions; converting generic pointers (cvta: an S2UR of the CTA id per call) and rebuilding descriptors per stage made
// them ~75 dependent instructions (~560 cycles) per stage -- more than the 256 tensor cycles of the stage's four MMAs, so the
// issuing warp, not the tensor pipe, set the pace.  With the addresses precomputed a stage is a handful of integer adds.
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_u32(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast_u32(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast_u32(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
// descriptor of the K-major 128B-swizzled operand tile at shared address `addr` (< 256 KiB, 16-byte aligned): the address field is
// additive, so desc(addr + d) = desc(addr) + (d >> 4)
__device__ __forceinline__ uint64_t umma_smem_desc_base(uint32_t addr) { return umma_smem_desc(addr); }

// ---- TS variant: queries in TMEM (M = 128), corpus streamed as the N operand -----------------------------
constexpr int kTS_M = 128;             // queries per CTA (UMMA M, one TMEM lane each)
constexpr int kTS_N = 128;             // passages per tile (UMMA N)
constexpr int kTS_BBytes = kTS_N * kDT_KB * 2;   // 16 KiB per corpus stage
constexpr int kTS_MaxStages = 12;
constexpr int kTS_MaxACols = 384;      // TMEM columns holding the query operand (C_pad <= 768)
constexpr int kTS_EpiWarps = 8;        // two warps per TMEM lane quadrant, each with half of the tile's passages
constexpr int kTS_EpiCols = kTS_N / (kTS_EpiWarps / 4);   // accumulator columns (passages) per epilogue thread
constexpr int kTS_Threads = 64 + 32 * kTS_EpiWarps;
constexpr int kTS_Prefetch = 2;         // L2 prefetch distance in tiles of one CTA
static_assert(kTS_N == kDenseTileRows && kDT_KB == kDenseTileCols, "K-blocked dense copy layout");

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// thread i of the warp writes its 32 registers to lane (base_lane + i), 32 consecutive columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct DenseTsArgs {
    long long row_begin, row_end;      // rows of this launch
    long long tile_row0;               // first row of tile 0 (row_begin rounded down to a 64-row tile from the launch base)
    long long scratch_row0;            // row of scratch line 0
    int n_tiles;                       // 128-row tiles in the launch
    int blocked;                       // corpus operand comes from the K-blocked copy
    int cluster;                       // 1: launched as clusters of two CTAs (the two query groups) sharing corpus tiles by TMA multicast
    int n_kblocks;                     // ceil(C_pad / 64)
    int n_stages;
    int n_qgroups;                     // 128-query groups in flight
    int n_queries;                     // valid queries in flight (slots)
    int mode;                          // 0 = filter + append, 1 = write scratch, 2 = scratch += scores, 3 = (scores + scratch) -> filter + append
    float* scratch;                    // [row - scratch_row0][scratch_slots]  (modes 1-3)
    long long scratch_slots;
    float* tau; uint32_t* cnt; float* cand_score; int32_t* cand_row; int cap;
    float* seg_score; int32_t* seg_row; uint32_t* seg_cnt;   // segmented candidate lists (filter modes) or nullptr: append with atomics
    const void* blocked_ptr;           // the K-blocked copy (experiments: plain bulk copies instead of tensor-map loads)
    int prefetch;                      // TMA L2 prefetch of the tile kTS_Prefetch iterations ahead (option dense_prefetch)
    const void* q16; int q_pitch, q_cols;   // lite kernel: the prepared fp16 queries (row pitch in elements, valid columns)
};

// ---- epilogue bodies shared by the cta_group::1 and cta_group::2 kernels: thread = one query, v = the scores of its kTS_EpiCols passages starting at row0 ----
// v += scratch (modes 2 and 3: partial sums of an earlier column pass, written by mode 1 / 2)
__device__ __forceinline__ void dense_ts_add_scratch(const DenseTsArgs& a, uint32_t (&v)[kTS_EpiCols / 32][32], int slot, long long row0) {
    if ((long long)slot < a.scratch_slots) {
        const float* src = a.scratch + (size_t)(row0 - a.scratch_row0) * a.scratch_slots + slot;
#pragma unroll
        for (int c = 0; c < kTS_EpiCols; ++c) {
            const long long row = row0 + c;
            if (row >= a.row_begin && row < a.row_end)
                v[c >> 5][c & 31] = __float_as_uint(__uint_as_float(v[c >> 5][c & 31]) + src[(size_t)c * a.scratch_slots]);
        }
    }
}

__device__ __forceinline__ void dense_ts_store_scratch(const DenseTsArgs& a, const uint32_t (&v)[kTS_EpiCols / 32][32], int slot, long long row0) {
    // scratch[row][slot]: for a fixed passage the 32 lanes of a warp write 32 consecutive slots (one 128-byte line)
    if ((long long)slot < a.scratch_slots) {
        float* dst = a.scratch + (size_t)(row0 - a.scratch_row0) * a.scratch_slots + slot;
        if (a.scratch_slots == kMaxInflight && row0 >= a.row_begin && row0 + kTS_EpiCols <= a.row_end) {
            // whole tile in range, compile-time row pitch: one store instruction per passage
#pragma unroll
            for (int c = 0; c < kTS_EpiCols; ++c) dst[(size_t)c * kMaxInflight] = __uint_as_float(v[c >> 5][c & 31]);
        } else {
#pragma unroll
            for (int c = 0; c < kTS_EpiCols; ++c) {
                const long long row = row0 + c;
                if (row >= a.row_begin && row < a.row_end)
                    dst[(size_t)c * a.scratch_slots] = __uint_as_float(v[c >> 5][c & 31]);
            }
        }
    }
}

__device__ __forceinline__ void dense_ts_filter_append(const DenseTsArgs& a, uint32_t (&v)[kTS_EpiCols / 32][32], int slot, long long row0, float tau_q) {
    // dense-only index: strict admission threshold.  Lanes without a valid query carry tau = +inf and never pass.  Rows pass
    // rarely once tau is set (about one (query, row) pair per warp and tile in the last chunk), so the work is kept O(passes)
    // and the code small: the lane's 128 scores are reduced to 16 group maxima and their maximum; only a lane that can pass
    // walks its groups, and a group that can pass counts its passing rows, reserves their slots with ONE atomicAdd and writes
    // them (one round trip per flagged group instead of one per appended row -- in the first chunk every row passes).
    // Measured forms (last chunk of a batch, 7.8 M rows x 256 queries, micro-benchmark pace 3.3 ms): one atomic per row 5.6 ms;
    // flat count / write passes over all 128 registers 8.3 ms (3,500 dependent instructions per warp and tile: the epilogue,
    // not the tensor pipe, was the bottleneck); whole-warp scan of one flagged lane at a time through shared memory 4.65 ms but
    // 2.3x slower in the middle chunks where many lanes pass; this form 4.9 ms and the fastest over a whole batch.
    if (!(row0 >= a.row_begin && row0 + kTS_EpiCols <= a.row_end)) {            // edge tile: rows outside the launch range never pass
#pragma unroll
        for (int c = 0; c < kTS_EpiCols; ++c) {
            const long long row = row0 + c;
            if (row < a.row_begin || row >= a.row_end) v[c >> 5][c & 31] = 0xFF800000u;      // -inf
        }
    }
    float gm[kTS_EpiCols / 8];
#pragma unroll
    for (int j = 0; j < kTS_EpiCols / 8; ++j) {
        float m = __uint_as_float(v[j >> 2][(j & 3) * 8]);
#pragma unroll
        for (int c = 1; c < 8; ++c) m = fmaxf(m, __uint_as_float(v[j >> 2][(j & 3) * 8 + c]));
        gm[j] = m;
    }
    float m = gm[0];
#pragma unroll
    for (int j = 1; j < kTS_EpiCols / 8; ++j) m = fmaxf(m, gm[j]);
    if (!(m > tau_q)) return;                                              // the common exit of (almost) every lane
    uint32_t* cnt = a.cnt + slot;
    float* cs = a.cand_score + (size_t)slot * a.cap;
    int32_t* cr = a.cand_row + (size_t)slot * a.cap;
    const int32_t r0 = (int32_t)row0;
    const uint32_t cap = (uint32_t)a.cap;
#pragma unroll
    for (int j = 0; j < kTS_EpiCols / 8; ++j) {
        if (gm[j] > tau_q) {
            uint32_t n = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) n += (__uint_as_float(v[j >> 2][(j & 3) * 8 + c]) > tau_q) ? 1u : 0u;
            uint32_t pos = atomicAdd(cnt, n);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float sc = __uint_as_float(v[j >> 2][(j & 3) * 8 + c]);
                if (sc > tau_q) {
                    if (pos < cap) { cs[pos] = sc + 0.0f; cr[pos] = r0 + 8 * j + c; }
                    ++pos;
                }
            }
        }
    }
}

// The same filter appending to this CTA's own segment of the query's candidate list: the thread is the only writer of
// (slot, segment), so its running count lives in a register (`my_cnt`, stored once when the kernel ends): no atomics, no round trip.
__device__ __forceinline__ void dense_ts_filter_append_seg(const DenseTsArgs& a, uint32_t (&v)[kTS_EpiCols / 32][32], long long row0, float tau_q,
                                                           float* seg_s, int32_t* seg_r, uint32_t& my_cnt) {
    if (!(row0 >= a.row_begin && row0 + kTS_EpiCols <= a.row_end)) {            // edge tile: rows outside the launch range never pass
#pragma unroll
        for (int c = 0; c < kTS_EpiCols; ++c) {
            const long long row = row0 + c;
            if (row < a.row_begin || row >= a.row_end) v[c >> 5][c & 31] = 0xFF800000u;      // -inf
        }
    }
    float gm[kTS_EpiCols / 8];
#pragma unroll
    for (int j = 0; j < kTS_EpiCols / 8; ++j) {
        float m = __uint_as_float(v[j >> 2][(j & 3) * 8]);
#pragma unroll
        for (int c = 1; c < 8; ++c) m = fmaxf(m, __uint_as_float(v[j >> 2][(j & 3) * 8 + c]));
        gm[j] = m;
    }
    float m = gm[0];
#pragma unroll
    for (int j = 1; j < kTS_EpiCols / 8; ++j) m = fmaxf(m, gm[j]);
    if (!(m > tau_q)) return;                                              // the common exit of (almost) every lane
    const int32_t r0 = (int32_t)row0;
    uint32_t pos = my_cnt;
#pragma unroll
    for (int j = 0; j < kTS_EpiCols / 8; ++j) {
        if (gm[j] > tau_q) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float sc = __uint_as_float(v[j >> 2][(j & 3) * 8 + c]);
                if (sc > tau_q) {
                    if (pos < (uint32_t)kSegCap) { seg_s[pos] = sc + 0.0f; seg_r[pos] = r0 + 8 * j + c; }
                    ++pos;
                }
            }
        }
    }
    my_cnt = pos;
}

__global__ void __launch_bounds__(kTS_Threads, 1)
dense_tile_ts_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_q, const DenseTsArgs a) {
    // a.cluster: the two query groups of a batch form a cluster of two CTAs that walk the same corpus tiles in lockstep; each
    // CTA loads half of every stage (64 passages) and multicasts it into both rings, so a corpus byte crosses L2->SM once for
    // 256 queries and the pair cannot drift apart (a drifting pair reads HBM twice).  Stage-empty barriers then count both
    // CTAs' MMA commits.
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kTS_MaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kTS_MaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ __align__(8) uint64_t q_bar;
    __shared__ uint32_t tmem_base_smem;

    const int k2dbg = K2_DBG();                                        // experiment bits, read once (0 in the library build)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) K2_TRACE(0, 0);
    const int qg = blockIdx.x % a.n_qgroups;
    const int cta_in_q = blockIdx.x / a.n_qgroups;
    const int ctas_per_q = gridDim.x / a.n_qgroups;
    const uint32_t a_cols = (uint32_t)a.n_kblocks * 32u;             // TMEM columns of the query operand
    uint8_t* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // SWIZZLE_128B tiles sit on 1 KiB boundaries

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], a.cluster ? 2 : 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kTS_EpiWarps); }   // only [0] is used (one accumulator)
        mbar_init(&q_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) K2_TRACE(0, 1);

    // ---- query operand -> TMEM ----
    // The 128 x C_pad fp16 query block is TMA-loaded into the (still idle) corpus ring as n_kblocks 128B-swizzled
    // [128 queries][64 columns] tiles of 16 KiB; epilogue thread (quarter, lane) owns query qg * 128 + quarter * 32 + lane,
    // reads its 128-byte row of every tile (chunk c of row r sits at chunk c ^ (r & 7): conflict-free) and writes it to
    // its TMEM lane with tcgen05.st.  Queries beyond n_queries and columns beyond C_pad are zero-filled by TMA.
    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&q_bar, (uint32_t)a.n_kblocks * (uint32_t)(kTS_M * kDT_KB * 2));
            for (int kb = 0; kb < a.n_kblocks; ++kb)
                tma_load_2d(ring + (size_t)kb * (kTS_M * kDT_KB * 2), &tmap_q, &q_bar, kb * kDT_KB, qg * kTS_M);
        }
        __syncwarp();
    } else if (warp >= 2) {
        const int quarter = warp & 3;
        const int qi = quarter * 32 + lane;
        mbar_wait(&q_bar, 0);
        for (int kb = (warp - 2) >> 2; kb < a.n_kblocks; kb += kTS_EpiWarps / 4) {
            const uint8_t* row = ring + (size_t)kb * (kTS_M * kDT_KB * 2) + (size_t)qi * 128;
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 x = *(const uint4*)(row + ((v ^ (qi & 7)) << 4));
                r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
            }
            tmem_st_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)kb * 32u, r);
        }
        tmem_st_wait();
    }
    // generic-proxy reads of the ring are complete before the async proxy (TMA) overwrites it with corpus tiles
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    if (a.cluster) cluster_sync_all();          // the peer may multicast into this ring / arrive on these barriers from here on
    else __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) K2_TRACE(0, 2);
    const uint32_t crank = a.cluster ? cluster_ctarank() : 0u;

    if (warp == 0) {
        // ===== TMA producer: corpus tiles (whole warp, one elected lane issues) =====
        int s = 0; uint32_t ph = 0; int i = 0;
        const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
        const uint32_t half_off = a.cluster ? crank * (uint32_t)(kTS_BBytes / 2) : 0u;
        const int half_rows = a.cluster ? (int)crank * (kTS_N / 2) : 0;
        const int pf_tiles = kTS_Prefetch * ctas_per_q;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            if (lane == 0) K2_TRACE(1, i);
            const int row0 = (int)(a.tile_row0 + (long long)t * kTS_N);
            const int blk0 = row0 / kTS_N * a.n_kblocks * kTS_N;                  // first row of the tile's k-block 0 in the blocked copy
            const bool do_pf = a.prefetch && qg == 0 && t + pf_tiles < a.n_tiles;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                if (k2dbg & 8) continue;
                mbar_wait_u32(empty_s + 8u * s, ph ^ 1u);
                if (elect_one()) {
                    // the ring holds exactly one corpus tile, so the demand load is issued only one tile-time ahead of its use:
                    // pull the same k-block of the tile `kTS_Prefetch` iterations further on into L2 now (HBM latency off the
                    // critical path); the two query groups share tiles, group 0 prefetches
                    if (do_pf && !(k2dbg & 64)) {
                        if (a.blocked) tma_prefetch_l2_2d(&tmap_c, 0, blk0 + (pf_tiles * a.n_kblocks + kb) * kTS_N);
                        else tma_prefetch_l2_2d(&tmap_c, kb * kDT_KB, row0 + pf_tiles * kTS_N);
                    }
                    const uint32_t bar = full_s + 8u * s;
                    const uint32_t dst = ring_s + (uint32_t)s * kTS_BBytes + half_off;
                    mbar_arrive_expect_tx_u32(bar, kTS_BBytes);                   // both halves (own + peer's multicast)
                    const int c0 = a.blocked ? 0 : kb * kDT_KB;
                    int c1 = (a.blocked ? blk0 + kb * kTS_N : row0) + half_rows;
                    if (k2dbg & 256) {          // stream-rate experiment: every CTA reads its own contiguous range of tiles
                        const int per = a.n_tiles / ctas_per_q;
                        c1 = (int)(((long long)(cta_in_q * per + (i % (per > 0 ? per : 1))) * a.n_kblocks + kb) % ((long long)a.n_tiles * a.n_kblocks)) * kTS_N + half_rows;
                    }
                    if (k2dbg & 128)            // stream-rate experiment: blocks read at one instant by all CTAs are adjacent in memory
                        c1 = (int)(((long long)(i * a.n_kblocks + kb) * ctas_per_q + cta_in_q) % ((long long)a.n_tiles * a.n_kblocks)) * kTS_N + half_rows;
                    if ((k2dbg & 32) && !a.cluster) {
                        const char* src = (const char*)a.blocked_ptr + (size_t)(blk0 + kb * kTS_N) * 128u;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst), "l"(src), "r"((uint32_t)kTS_BBytes), "r"(bar) : "memory");
                    } else
                    if (a.cluster) tma_load_2d_multicast_u32(dst, &tmap_c, bar, c0, c1, 3);
                    else tma_load_2d_u32(dst, &tmap_c, bar, c0, c1);
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp in uniform control flow, one elected lane issues) =====
        // One accumulator D[128 queries x 128 passages] (the query operand leaves 128 TMEM columns): an M128 x N128 x K16
        // instruction costs about the same ~90 cycles as an N64 one (measured per-instruction floor), so the wide tile
        // halves the tensor time per (query, passage) pair; the price is a short bubble while the epilogue drains D.
        constexpr uint32_t idesc = umma_idesc_f16(kTS_M, kTS_N);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        const uint32_t d_tmem = tmem_u + a_cols;
        const uint32_t full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar), tempty_s = smem_u32(&tempty_bar[0]), tfull_s = smem_u32(&tfull_bar[0]);
        const uint64_t desc0 = umma_smem_desc_base(smem_u32(ring));              // stage s, k-step k: desc0 + s * (16 KiB >> 4) + k * (32 B >> 4)
        int s = 0; uint32_t ph = 0;
        int i = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            if (lane == 0) K2_TRACE(2, i);
            mbar_wait_u32(tempty_s, ((uint32_t)i & 1u) ^ 1u);
            tc_fence_after();
            if (lane == 0) K2_TRACE(3, i);
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                if (!(k2dbg & 8)) mbar_wait_u32(full_s + 8u * s, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bd = desc0 + (uint64_t)((uint32_t)s * (kTS_BBytes >> 4));
                    const uint32_t at = tmem_u + (uint32_t)kb * 32u;
                    if (!(k2dbg & 16)) {
#pragma unroll
                        for (int k = 0; k < kDT_KB / 16; ++k)
                            umma_f16_ts(d_tmem, at + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (!(k2dbg & 8)) {                                // frees the corpus stage once these MMAs have read it
                        if ((k2dbg & 1024) && !a.cluster) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_s + 8u * s) : "memory");
                        else
                        if (a.cluster) umma_commit_multicast_u32(empty_s + 8u * s, 3);
                        else umma_commit_u32(empty_s + 8u * s);
                    }
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit_u32(tfull_s);                    // accumulator tile complete
            __syncwarp();
            if (lane == 0) K2_TRACE(4, i);
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4; thread = one query x 64 passages =====
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int qi = quarter * 32 + lane;
        const int slot = qg * kTS_M + qi;
        const bool q_ok = slot < a.n_queries;
        const float tau_q = (q_ok && (a.mode == 0 || a.mode == 3)) ? a.tau[slot] : INFINITY;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_cols + (uint32_t)(chalf * kTS_EpiCols);
        const bool use_seg = a.seg_cnt != nullptr && q_ok && (a.mode == 0 || a.mode == 3);
        const size_t seg_id = (size_t)(q_ok ? slot : 0) * kSegCount + (size_t)(cta_in_q * (kTS_EpiWarps / 4) + chalf);   // this thread's own list segment
        uint32_t my_cnt = 0;
        int i = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            mbar_wait(&tfull_bar[0], (uint32_t)i & 1u);
            tc_fence_after();
            if (threadIdx.x == 64) K2_TRACE(5, i);
            uint32_t v[kTS_EpiCols / 32][32];
#pragma unroll
            for (int j = 0; j < kTS_EpiCols / 32; ++j) tmem_ld_32x32(taddr + 32u * j, v[j]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)                                               // D is free again: the next tile's MMAs may start (relaxed: see mbar_arrive_cluster_relaxed)
                asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[0])) : "memory");
            if (threadIdx.x == 64) K2_TRACE(6, i);
            if ((k2dbg & 512) && a.blocked) {
                const int tp = t + ((k2dbg >> 12) & 7) * ctas_per_q;
                if (tp < a.n_tiles) {
                    const char* pbase = (const char*)a.blocked_ptr + (size_t)(a.tile_row0 / kTS_N + tp) * a.n_kblocks * kTS_BBytes;
                    const int nshare = a.cluster ? 2 : 1;
                    const int n256 = a.n_kblocks * kTS_BBytes / 256;
                    for (int u = ((int)threadIdx.x - 64) * nshare + (int)crank; u < n256; u += 32 * kTS_EpiWarps * nshare) {
                        uint32_t dummy;
                        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.u32 %0, [%1];" : "=r"(dummy) : "l"(pbase + (size_t)u * 256));
                    }
                }
            }
            const long long row0 = a.tile_row0 + (long long)t * kTS_N + chalf * kTS_EpiCols;
            if (a.mode >= 2) dense_ts_add_scratch(a, v, slot, row0);
            if (k2dbg & 4) {
            } else if (a.mode == 1 || a.mode == 2) dense_ts_store_scratch(a, v, slot, row0);
            else if (use_seg) dense_ts_filter_append_seg(a, v, row0, tau_q, a.seg_score + seg_id * kSegCap, a.seg_row + seg_id * kSegCap, my_cnt);
            else dense_ts_filter_append(a, v, slot, row0, tau_q);
        }
        if (use_seg) a.seg_cnt[seg_id] = my_cnt;
    }
    if (threadIdx.x == 64) K2_TRACE(0, 3);
    tc_fence_before();
    if (a.cluster) cluster_sync_all();          // no CTA of the pair exits while the other may still signal its barriers
    else __syncthreads();
    if (threadIdx.x == 0) K2_TRACE(0, 4);
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- TS2 variant: a CTA pair works as one M = 256 tensor-core unit (cta_group::2) ----------------------------------------
// The two 128-query groups of a batch run as a cluster of two CTAs on the two SMs of a TPC.  Each CTA keeps ITS 128 queries
// in ITS tensor memory (the A operand, M = 256 over the pair) and stages only HALF of every corpus tile (64 of the 128
// passages of a k-block, 8 KiB) in its shared memory; one thread of the leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2 M256 x N128 x K16, for which the hardware reads each CTA's half of the B operand from that CTA's
// shared memory and writes D[128 queries x 128 passages] into each CTA's own TMEM.  Against the cta_group::1 kernel this
// halves the shared-memory traffic per SM (TMA fill + operand read) and the L2 -> SM traffic per tensor instruction, which
// is what holds the single-CTA kernel at ~150 cycles per instruction against the 64-cycle floor.
//   full[s]   lives in the leader: both CTAs' TMA loads complete_tx on it (the peer through the .cta_group::2 form with the
//             leader's barrier address), the leader's producer posts the expect_tx for both halves;
//   empty[s]  in each CTA, signalled by the leader's tcgen05.commit multicast once the MMAs have read the stage;
//   tfull     in each CTA, same multicast commit after the last k-block of a tile;
//   tempty    in the leader, 8 arrivals: the four epilogue warps of both CTAs (the peer's arrive remotely).
constexpr int kTS2_HalfN = kTS_N / 2;                       // passages staged per CTA and k-block
constexpr int kTS2_StageBytes = kTS2_HalfN * kDT_KB * 2;    // 8 KiB
constexpr int kTS2_MaxStages = 24;                          // 192 KiB ring = two corpus tiles deep per CTA (also stages the query operand)

__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// The accumulator-drained signal orders no memory: the TMEM reads are complete (tcgen05.wait::ld) and fenced
// (tcgen05.fence::before_thread_sync).  A release arrive would also wait for the thread's earlier global stores (the previous
// tile's candidate / scratch writes) -- measured: the dominant part of the per-tile bubble.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA 2-D load into THIS CTA's shared memory whose completion is counted on a barrier of the CTA pair's leader
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __launch_bounds__(kTS_Threads, 1)
dense_tile_ts2_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_q, const DenseTsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kTS2_MaxStages];      // used in the leader only
    __shared__ __align__(8) uint64_t empty_bar[kTS2_MaxStages];
    __shared__ __align__(8) uint64_t tfull_bar;
    __shared__ __align__(8) uint64_t tempty_bar;                     // used in the leader only
    __shared__ __align__(8) uint64_t q_bar;
    __shared__ uint32_t tmem_base_smem;

    const int k2dbg = K2_DBG();                                        // experiment bits, read once (0 in the library build)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const int qg = (int)crank;                                       // query group of this CTA (a.n_qgroups == 2)
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t a_cols = (uint32_t)a.n_kblocks * 32u;
    uint8_t* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tfull_bar, 1);
        mbar_init(&tempty_bar, 2 * kTS_EpiWarps);
        mbar_init(&q_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_2sm(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // ---- query operand -> TMEM (as in the cta_group::1 kernel; every CTA loads its own group through the idle ring) ----
    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&q_bar, (uint32_t)a.n_kblocks * (uint32_t)(kTS_M * kDT_KB * 2));
            for (int kb = 0; kb < a.n_kblocks; ++kb)
                tma_load_2d(ring + (size_t)kb * (kTS_M * kDT_KB * 2), &tmap_q, &q_bar, kb * kDT_KB, qg * kTS_M);
        }
        __syncwarp();
    } else if (warp >= 2) {
        const int quarter = warp & 3;
        const int qi = quarter * 32 + lane;
        mbar_wait(&q_bar, 0);
        for (int kb = (warp - 2) >> 2; kb < a.n_kblocks; kb += kTS_EpiWarps / 4) {
            const uint8_t* row = ring + (size_t)kb * (kTS_M * kDT_KB * 2) + (size_t)qi * 128;
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 x = *(const uint4*)(row + ((v ^ (qi & 7)) << 4));
                r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
            }
            tmem_st_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)kb * 32u, r);
        }
        tmem_st_wait();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    cluster_sync_all();                          // barriers initialised and both query operands in place before any cross-CTA signal
    tc_fence_after();

    if (warp == 0) {
        // ===== TMA producer: this CTA's half (64 passages) of every corpus stage =====
        int s = 0; uint32_t ph = 0;
        const uint32_t full0 = mapa_shared(smem_u32(&full_bar[0]), 0u);          // the leader's full barriers (cluster address)
        const uint32_t full_local = smem_u32(full_bar), empty_s = smem_u32(empty_bar), ring_s = smem_u32(ring);
        const int half = (int)crank * kTS2_HalfN;
        const int pf_tiles = kTS_Prefetch * n_pairs;
        for (int t = pair; t < a.n_tiles; t += n_pairs) {
            const int row0 = (int)(a.tile_row0 + (long long)t * kTS_N);
            const int blk0 = row0 / kTS_N * a.n_kblocks * kTS_N;
            const bool do_pf = a.prefetch && leader && t + pf_tiles < a.n_tiles;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                if (k2dbg & 8) continue;
                mbar_wait_u32(empty_s + 8u * s, ph ^ 1u);
                if (elect_one()) {
                    if (do_pf && !(k2dbg & 64)) {
                        if (a.blocked) tma_prefetch_l2_2d(&tmap_c, 0, blk0 + (pf_tiles * a.n_kblocks + kb) * kTS_N);
                        else tma_prefetch_l2_2d(&tmap_c, kb * kDT_KB, row0 + pf_tiles * kTS_N);
                    }
                    if (leader) mbar_arrive_expect_tx_u32(full_local + 8u * s, 2u * kTS2_StageBytes);
                    const uint32_t dst = ring_s + (uint32_t)s * kTS2_StageBytes;
                    const uint32_t bar = full0 + (uint32_t)s * 8u;
                    const int c0 = a.blocked ? 0 : kb * kDT_KB;
                    const int c1 = (a.blocked ? blk0 + kb * kTS_N : row0) + half;
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(dst), "l"(&tmap_c), "r"(bar), "r"(c0), "r"(c1) : "memory");
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader CTA only =====
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * kTS_M, kTS_N);
            const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
            const uint32_t d_tmem = tmem_u + a_cols;
            const uint32_t full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar), tempty_s = smem_u32(&tempty_bar), tfull_s = smem_u32(&tfull_bar);
            const uint64_t desc0 = umma_smem_desc_base(smem_u32(ring));          // stage s, k-step k: desc0 + s * (8 KiB >> 4) + k * (32 B >> 4)
            int s = 0; uint32_t ph = 0;
            int i = 0;
            for (int t = pair; t < a.n_tiles; t += n_pairs, ++i) {
                if (!(k2dbg & 2048)) mbar_wait_u32(tempty_s, ((uint32_t)i & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < a.n_kblocks; ++kb) {
                    if (!(k2dbg & 8)) mbar_wait_u32(full_s + 8u * s, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bd = desc0 + (uint64_t)((uint32_t)s * (kTS2_StageBytes >> 4));
                        const uint32_t at = tmem_u + (uint32_t)kb * 32u;
                        if (!(k2dbg & 16)) {
#pragma unroll
                            for (int k = 0; k < kDT_KB / 16; ++k)
                                umma_f16_ts_2sm(d_tmem, at + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        if (k2dbg & 1024) {                                // experiment: software arrives instead of the commit
                            mbar_arrive_cluster(mapa_shared(empty_s + 8u * s, 0u));
                            mbar_arrive_cluster(mapa_shared(empty_s + 8u * s, 1u));
                        } else
                        if (!(k2dbg & 8))
                        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(empty_s + 8u * s), "h"((uint16_t)3) : "memory");     // both CTAs may refill the stage
                    }
                    __syncwarp();
                    if (++s == a.n_stages) { s = 0; ph ^= 1u; }
                }
                if (!(k2dbg & 2048) && elect_one())                        // accumulator tile complete in both CTAs
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(tfull_s), "h"((uint16_t)3) : "memory");
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps (both CTAs): thread = one query of this CTA's group x 64 passages (column half = (warp - 2) / 4) =====
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int qi = quarter * 32 + lane;
        const int slot = qg * kTS_M + qi;
        const bool q_ok = slot < a.n_queries;
        const float tau_q = (q_ok && (a.mode == 0 || a.mode == 3)) ? a.tau[slot] : INFINITY;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_cols + (uint32_t)(chalf * kTS_EpiCols);
        const uint32_t tempty_leader = mapa_shared(smem_u32(&tempty_bar), 0u);
        const bool use_seg = a.seg_cnt != nullptr && q_ok && (a.mode == 0 || a.mode == 3);
        const size_t seg_id = (size_t)(q_ok ? slot : 0) * kSegCount + (size_t)(pair * (kTS_EpiWarps / 4) + chalf);
        uint32_t my_cnt = 0;
        int i = 0;
        for (int t = pair; t < a.n_tiles; t += n_pairs, ++i) {
            if (k2dbg & 2048) break;
            mbar_wait(&tfull_bar, (uint32_t)i & 1u);
            tc_fence_after();
            uint32_t v[kTS_EpiCols / 32][32];
#pragma unroll
            for (int j = 0; j < kTS_EpiCols / 32; ++j) tmem_ld_32x32(taddr + 32u * j, v[j]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader);   // this warp's part of this CTA's D is free again
            const long long row0 = a.tile_row0 + (long long)t * kTS_N + chalf * kTS_EpiCols;
            if (a.mode >= 2) dense_ts_add_scratch(a, v, slot, row0);
            if (k2dbg & 4) {
            } else
            if (a.mode == 1 || a.mode == 2) dense_ts_store_scratch(a, v, slot, row0);
            else if (use_seg) dense_ts_filter_append_seg(a, v, row0, tau_q, a.seg_score + seg_id * kSegCap, a.seg_row + seg_id * kSegCap, my_cnt);
            else dense_ts_filter_append(a, v, slot, row0, tau_q);
        }
        if (use_seg) a.seg_cnt[seg_id] = my_cnt;
    }
    tc_fence_before();
    cluster_sync_all();                          // no CTA of the pair exits (or frees TMEM) while the other may still signal it
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// ---- lite variant: K2 small enough to share an SM with a K1t CTA --------------------------------------------------------
// In the hybrid path K2 (scratch mode) and K1t both wanted > 190 KB of shared memory, so they alternated on the SMs and K2's
// 20 us per sub-chunk were lost to K1t.  K1t is bound by instruction issue and shared-memory wavefronts and uses no tensor
// core, no TMEM and 43 K of the 64 K registers: this variant fits beside it -- 64-passage tiles through a 2-4 stage ring of
// 8 KiB stages (the layout of the cta_group::2 kernel's half stages), the query operand read straight from global memory
// into TMEM (no staging ring), two 64-column accumulators (no drain bubble), 192 threads under a 112-register cap, epilogue
// = plain scratch stores (modes 1 and 2 only).  Slower than the big kernel when alone, free when it runs in K1t's shadow.
constexpr int kTL_N = 64;
// waits of the lite kernel back off: its six warps share the SM's issue slots with K1t's seventeen, and K1t is issue-bound
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_u32(bar, parity)) {
        __nanosleep(128);
        if (++spins > (1u << 24)) __trap();
    }
}
constexpr int kTL_StageBytes = kTL_N * kDT_KB * 2;   // 8 KiB
constexpr int kTL_MaxStages = 8;
constexpr int kTL_Threads = 192;                     // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
static_assert(kTL_N == kTS_EpiCols, "the shared epilogue bodies take 64 columns per thread");

__global__ void __launch_bounds__(kTL_Threads, 3)
dense_tile_lite_kernel(const __grid_constant__ CUtensorMap tmap_c, const DenseTsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kTL_MaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kTL_MaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qg = blockIdx.x % a.n_qgroups;
    const int cta_in_q = blockIdx.x / a.n_qgroups;
    const int ctas_per_q = gridDim.x / a.n_qgroups;
    const uint32_t a_cols = (uint32_t)a.n_kblocks * 32u;
    uint8_t* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer: 64-passage x 64-column pieces of the K-blocked copy (8 KiB contiguous in HBM) =====
        int s = 0; uint32_t ph = 0;
        const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q) {
            const int row0 = (int)(a.tile_row0 + (long long)t * kTL_N);
            const int blk0 = row0 / kTS_N * a.n_kblocks * kTS_N + (row0 % kTS_N);   // row of the half tile in k-block 0 of its 128-row block
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait_sleep(empty_s + 8u * s, ph ^ 1u);
                if (elect_one()) {
                    const uint32_t bar = full_s + 8u * s;
                    mbar_arrive_expect_tx_u32(bar, kTL_StageBytes);
                    tma_load_2d_u32(ring_s + (uint32_t)s * kTL_StageBytes, &tmap_c, bar, 0, blk0 + kb * kTS_N);
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: M128 x N64 x K16, accumulators alternate between two 64-column buffers =====
        constexpr uint32_t idesc = umma_idesc_f16(kTS_M, kTL_N);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        const uint32_t full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar), tempty_s = smem_u32(tempty_bar), tfull_s = smem_u32(tfull_bar);
        const uint64_t desc0 = umma_smem_desc_base(smem_u32(ring));
        // the epilogue warps write the query operand and then arrive once on tempty_bar[0] and [1]: phase 0 of either barrier
        // means "operand in TMEM", every later phase "this accumulator buffer has been drained"
        int s = 0; uint32_t ph = 0;
        int i = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            const uint32_t buf = (uint32_t)i & 1u;
            mbar_wait_sleep(tempty_s + 8u * buf, ((uint32_t)i >> 1) & 1u);           // phase 0 = "queries in TMEM", then one phase per drained tile
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + a_cols + buf * (uint32_t)kTL_N;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait_sleep(full_s + 8u * s, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bd = desc0 + (uint64_t)((uint32_t)s * (kTL_StageBytes >> 4));
                    const uint32_t at = tmem_u + (uint32_t)kb * 32u;
#pragma unroll
                    for (int k = 0; k < kDT_KB / 16; ++k)
                        umma_f16_ts(d_tmem, at + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_u32(empty_s + 8u * s);
                }
                __syncwarp();
                if (++s == a.n_stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit_u32(tfull_s + 8u * buf);
            __syncwarp();
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4; thread = one query =====
        const int quarter = warp & 3;
        const int qi = quarter * 32 + lane;
        const int slot = qg * kTS_M + qi;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        {   // query operand -> TMEM straight from global memory: lane = query, column kb * 32 + j = fp16 pair (2j, 2j + 1) of k-block kb
            const uint4* qrow = (const uint4*)((const __half*)a.q16 + (size_t)slot * a.q_pitch);
            const bool q_in = slot < a.n_queries;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    uint4 x = make_uint4(0u, 0u, 0u, 0u);
                    if (q_in && kb * kDT_KB + v * 8 < a.q_cols) x = __ldg(qrow + kb * 8 + v);
                    r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                }
                tmem_st_32x32(lane_addr + (uint32_t)kb * 32u, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&tempty_bar[0]); mbar_arrive(&tempty_bar[1]); }     // phase 0 of both: operand in place
        }
        const uint32_t tempty_s = smem_u32(tempty_bar);
        int i = 0;
        for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q, ++i) {
            const uint32_t buf = (uint32_t)i & 1u;
            mbar_wait_sleep(smem_u32(&tfull_bar[buf]), ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            uint32_t v[kTS_EpiCols / 32][32];
#pragma unroll
            for (int j = 0; j < kTS_EpiCols / 32; ++j) tmem_ld_32x32(lane_addr + a_cols + buf * (uint32_t)kTL_N + 32u * j, v[j]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(tempty_s + 8u * buf) : "memory");
            const long long row0 = a.tile_row0 + (long long)t * kTL_N;
            if (a.mode >= 2) dense_ts_add_scratch(a, v, slot, row0);
            dense_ts_store_scratch(a, v, slot, row0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return nullptr; }
    fn = (EncodeTiledFn)p;
    return fn;
}

// 2-D fp16 row-major tensor [rows][cols] with row pitch `pitch_elems`; box = 64 columns x box_rows, 128B swizzle
static int make_tmap_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return DHR_ERR_CUDA;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)kDT_KB, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled", __FILE__, __LINE__);
        return DHR_ERR_CUDA;
    }
    return DHR_OK;
}

bool dense_tile_supported(const Geometry& g, int* n_stages_out) {
    if (g.C_pad <= 0) return false;
    const int nkb = (g.C_pad + kDT_KB - 1) / kDT_KB;
    const size_t budget = 220 * 1024;
    const size_t b_bytes = (size_t)nkb * kDT_BBytes;
    if (b_bytes + 2 * (size_t)kDT_ABytes > budget) return false;
    int stages = (int)((budget - b_bytes) / kDT_ABytes);
    if (stages > kDT_MaxStages) stages = kDT_MaxStages;
    if (n_stages_out) *n_stages_out = stages;
    return true;
}

bool dense_tile_ts_supported(const Geometry& g) {
    return g.C_pad > 0 && (g.C_pad + kDT_KB - 1) / kDT_KB * 32 <= kTS_MaxACols;
}

// One column pass of the TS kernels: scores[q][row] (+)= sum over `cols` fp16 columns of a corpus block and the matching
// query columns.  The corpus block is either the K-blocked copy of the dense block (`blocked`) or any row-major fp16 array
// with row pitch `c_pitch` (the dense block itself, or a column range of the lexical values for the unmasked --IP stage).
int launch_dense_pass(const dhr_index* h, const __half* blocked, const __half* rowmajor, int c_pitch, int cols, const void* q16,
                      int q_pitch, int n_queries, long long tile_row0, long long row_begin, long long row_end, int mode, float* scratch,
                      long long scratch_slots, const TopkState& t, int cap, cudaStream_t st) {
    if (row_end <= row_begin || n_queries <= 0) return DHR_OK;
    if (cols <= 0 || (cols + kDT_KB - 1) / kDT_KB * 32 > kTS_MaxACols) return DHR_ERR_UNSUPPORTED;
    CUtensorMap tmap_c, tmap_q;
    const int nkb = (cols + kDT_KB - 1) / kDT_KB;
    // corpus operand (built below): the K-blocked copy viewed as [blocks * 128 rows][64 cols] (one box = one contiguous piece
    // of HBM), or a row-major block
    DHR_TRY(make_tmap_f16(&tmap_q, q16, (uint64_t)n_queries, (uint64_t)cols, (uint64_t)q_pitch, kTS_M));
    DenseTsArgs a{};
    a.row_begin = row_begin; a.row_end = row_end;
    a.tile_row0 = row_begin / kTS_N * kTS_N;                     // absolute 128-row tiles (the K-blocked copy is tiled from row 0)
    a.blocked = blocked != nullptr;
    a.scratch_row0 = tile_row0;
    a.n_tiles = (int)((row_end - a.tile_row0 + kTS_N - 1) / kTS_N);
    a.n_kblocks = nkb;
    a.n_stages = kTS_MaxStages;
    a.n_qgroups = (n_queries + kTS_M - 1) / kTS_M;
    a.n_queries = n_queries;
    a.mode = mode;
    a.scratch = scratch; a.scratch_slots = scratch_slots;
    a.tau = t.tau; a.cnt = t.cnt; a.cand_score = t.cand_score; a.cand_row = t.cand_row; a.cap = cap;
    a.seg_score = t.seg_score; a.seg_row = t.seg_row; a.seg_cnt = t.seg_cnt;
    a.blocked_ptr = blocked;
    a.prefetch = h->opt_dense_prefetch;
    // lite form (hybrid path, scratch mode, K2 overlapped with K1t): sized by the caller to fit beside a K1t CTA on every SM
    if (h->opt_dense_lite && h->lite_stages >= 2 && mode == 1 && blocked && ((uintptr_t)q16 & 15u) == 0 && q_pitch % 8 == 0) {
        a.n_tiles = (int)((row_end - a.tile_row0 + kTL_N - 1) / kTL_N);
        a.n_stages = h->lite_stages < kTL_MaxStages ? h->lite_stages : kTL_MaxStages;
        a.q16 = q16; a.q_pitch = q_pitch; a.q_cols = cols;
        a.cluster = 0;
        DHR_TRY(make_tmap_f16(&tmap_c, blocked, (uint64_t)round_up(h->n_rows, kTS_N) * nkb, (uint64_t)kDT_KB, (uint64_t)kDT_KB, kTL_N));
        int per_q = h->num_sms / a.n_qgroups;
        if (per_q < 1) per_q = 1;
        if (per_q > a.n_tiles) per_q = a.n_tiles;
        static bool carveout_set = false;
        if (!carveout_set) {      // never let this small kernel configure an SM with a shared-memory carveout K1t does not fit in
            DHR_CUDA(cudaFuncSetAttribute(dense_tile_lite_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            carveout_set = true;
        }
        const size_t lsmem = (size_t)a.n_stages * kTL_StageBytes + 1024;
        dense_tile_lite_kernel<<<per_q * a.n_qgroups, kTL_Threads, lsmem, st>>>(tmap_c, a);
        DHR_CUDA(cudaGetLastError());
        return DHR_OK;
    }
    const size_t smem = (size_t)a.n_stages * kTS_BBytes + 1024;
    DHR_CUDA(cudaFuncSetAttribute(dense_tile_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DHR_CUDA(cudaFuncSetAttribute(dense_tile_ts2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTS2_MaxStages * kTS2_StageBytes + 1024));
    int per_q = h->num_sms / a.n_qgroups;
    if (per_q < 1) per_q = 1;
    if (per_q > a.n_tiles) per_q = a.n_tiles;
    // Dense-only searches (long launches, HBM-fed) gain 12 %; in the hybrid path K2 runs beside the tail of K1t on whatever SMs
    // free up, and a cluster needs both SMs of a pair at once (measured 1.5 % slower), so scratch mode stays unicast unless
    // asked for (dense_multicast = 2).
    a.cluster = (a.n_qgroups == 2 && ((h->opt_dense_multicast == 1 && mode == 0) || h->opt_dense_multicast == 2)) ? 1 : 0;
    // cta_group::2 form (dense_variant 2): the two query groups of a batch as one CTA pair; needs exactly two groups in flight
    // (dense_variant 3 = automatic: the pair form for the long filter-mode launches of dense-only / --IP searches, the single-CTA
    // form for the short scratch-mode launches of the hybrid path, where a cluster has to wait for both SMs of a TPC)
    const bool want_pair = h->opt_dense_variant == 2 || (h->opt_dense_variant == 3 && (mode == 0 || mode == 3));
    const bool pair_mma = want_pair && a.n_qgroups == 2 && h->num_sms >= 2;
    if (pair_mma) { a.cluster = 1; a.n_stages = kTS2_MaxStages; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(per_q * a.n_qgroups));
    cfg.blockDim = dim3(kTS_Threads);
    cfg.dynamicSmemBytes = pair_mma ? (size_t)kTS2_MaxStages * kTS2_StageBytes + 1024 : smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (a.cluster) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // pairs the device can keep resident: probed per launch configuration and device (cheap host-side query)
        int max_clusters = 0;
        {
            cudaLaunchConfig_t probe = cfg;
            probe.gridDim = dim3((unsigned)(h->num_sms / 2 * 2));
            cudaError_t pe = pair_mma ? cudaOccupancyMaxActiveClusters(&max_clusters, dense_tile_ts2_kernel, &probe)
                                      : cudaOccupancyMaxActiveClusters(&max_clusters, dense_tile_ts_kernel, &probe);
            if (pe != cudaSuccess || max_clusters < 1) { cudaGetLastError(); max_clusters = 0; }
        }
        if (max_clusters < 1) {                                        // no cluster launch on this device / partition: unicast
            if (pair_mma) return DHR_ERR_UNSUPPORTED;
            a.cluster = 0;
            cfg.attrs = nullptr; cfg.numAttrs = 0;
        } else if (per_q > max_clusters) {
            per_q = max_clusters;
            cfg.gridDim = dim3((unsigned)(per_q * 2));
        }
    }
    if (per_q * (kTS_EpiWarps / 4) > kSegCount) { a.seg_score = nullptr; a.seg_row = nullptr; a.seg_cnt = nullptr; }   // more writers than list segments: atomics
    // with multicast / the CTA-pair form the corpus map delivers half a stage (64 passages) per load
    const int box_rows = a.cluster ? kTS_N / 2 : kTS_N;
    if (blocked)
        DHR_TRY(make_tmap_f16(&tmap_c, blocked, (uint64_t)round_up(h->n_rows, kTS_N) * nkb, (uint64_t)kDT_KB, (uint64_t)kDT_KB, box_rows));
    else
        DHR_TRY(make_tmap_f16(&tmap_c, rowmajor, (uint64_t)h->n_rows, (uint64_t)cols, (uint64_t)c_pitch, box_rows));
    if (pair_mma) {
        DHR_CUDA(cudaLaunchKernelEx(&cfg, dense_tile_ts2_kernel, tmap_c, tmap_q, a));
        return DHR_OK;
    }
    DHR_CUDA(cudaLaunchKernelEx(&cfg, dense_tile_ts_kernel, tmap_c, tmap_q, a));
    return DHR_OK;
}

// Launch K2 over rows [row_begin, row_end) (row_begin aligned to 128 from tile_row0) for `n_queries` in-flight queries
// whose fp16 dense block starts at q_dns16 (row pitch C_pad).
int launch_dense_tile(const dhr_index* h, const void* q_dns16, int n_queries, long long tile_row0, long long row_begin,
                      long long row_end, int mode, float* scratch, long long scratch_slots, const TopkState& t, int cap,
                      cudaStream_t st) {
    const Geometry& g = h->g;
    int stages = 0;
    if (!dense_tile_supported(g, &stages)) return DHR_ERR_UNSUPPORTED;
    if (row_end <= row_begin || n_queries <= 0) return DHR_OK;
    if (h->opt_dense_variant >= 1 && dense_tile_ts_supported(g))
        return launch_dense_pass(h, h->dnst, h->dns, g.C_pad, g.C_pad, q_dns16, g.C_pad, n_queries, tile_row0, row_begin, row_end, mode, scratch,
                                 scratch_slots, t, cap, st);
    CUtensorMap tmap_a, tmap_b;
    DHR_TRY(make_tmap_f16(&tmap_a, h->dns, (uint64_t)h->n_rows, (uint64_t)g.C_pad, (uint64_t)g.C_pad, kDT_M));
    DHR_TRY(make_tmap_f16(&tmap_b, q_dns16, (uint64_t)n_queries, (uint64_t)g.C_pad, (uint64_t)g.C_pad, kDT_N));
    DenseTileArgs a{};
    a.row_begin = row_begin; a.row_end = row_end; a.tile_row0 = tile_row0;
    const long long first_tile = (row_begin - tile_row0) / kDT_M;
    a.tile_row0 = tile_row0 + first_tile * kDT_M;
    a.n_tiles = (int)((row_end - a.tile_row0 + kDT_M - 1) / kDT_M);
    a.n_kblocks = (g.C_pad + kDT_KB - 1) / kDT_KB;
    a.n_stages = stages;
    a.n_qtiles = (n_queries + kDT_N - 1) / kDT_N;
    a.n_queries = n_queries;
    a.mode = mode;
    a.scratch = scratch ? scratch + (size_t)(a.tile_row0 - tile_row0) * scratch_slots : nullptr;
    a.scratch_slots = scratch_slots;
    a.tau = t.tau; a.cnt = t.cnt; a.cand_score = t.cand_score; a.cand_row = t.cand_row; a.cap = cap;
    const size_t smem = (size_t)a.n_kblocks * kDT_BBytes + (size_t)stages * kDT_ABytes + 1024;
    DHR_CUDA(cudaFuncSetAttribute(dense_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_q = h->num_sms / a.n_qtiles;
    if (per_q < 1) per_q = 1;
    if (per_q > a.n_tiles) per_q = a.n_tiles;
    dense_tile_kernel<<<(unsigned)(per_q * a.n_qtiles), kDT_Threads, smem, st>>>(tmap_a, tmap_b, a);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

}  // namespace dhr
