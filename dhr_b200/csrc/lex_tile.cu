// lex_tile.cu -- K1t: tiled lexical match-and-MAC for a tile of 64 queries x 512 passages, fused with
// the admission filter.  Throughput-mode counterpart of the row scan K1 (scan_rows.cuh).
//
// Replaces, for 64 queries at once, the masked product + row dot of castorini/dhr
// retrieval/gip_retrieval.py:119-120 restricted to the lexical columns:
//     lex[q][p] = sum_s [q_idx[s] == p_idx[s]] * sum_g q_val[s,g] * p_val[s,g]
// without comparing every (query, passage, slice) triple: per query tile and slice the queries are
// bucketed by their slice index (code) -- a table tab[slice][code] = {byte offset of the first entry,
// entry count} over a packed entry list {acc byte offset of the query, G fp16 values} -- so a passage
// thread looks up the bucket of ITS code and touches only the queries that really match.  Work is
// O(matches), not O(Q*N*S).
//
// Data movement: one producer warp streams, per (passage tile, 4-slice chunk), the tiled corpus block
// (codes u8 [512][4] | values [4][512][G]) and the query-tile block (table | entries) with TMA bulk
// copies (cp.async.bulk + mbarrier ring) into shared memory; 512 consumer threads own one passage each
// and keep acc[64 queries][512 passages] fp32 in shared memory (column p is private to thread p:
// bank = lane, conflict-free, no atomics, fixed summation order).
//
// Per thread and chunk: (A) four branch-free bucket lookups (one 32-bit table word each); (B) the
// matches of all four slices are walked as one flattened list, level l = the thread's l-th match.
// The walk is branch-free inside a level (predicated shared loads/stores) and software-pipelined: the
// entry and passage-value loads of level l+1 are issued before the accumulate of level l, and the
// G-term dot is split into two independent FMA chains, so a warp is bound by issue slots and the
// shared-memory pipe rather than by the latency of one level's dependent chain.
// acc is initialised from the dense scores written by K2 (hybrid index) or zero, and after the last
// chunk each thread applies the strict threshold tau[q] and appends winners to the candidate lists.
#include "internal.h"

namespace dhr {

constexpr int kLT_PT = kLexTileRows;    // passages per tile = consumer threads
constexpr int kLT_QT = kLexTileQueries; // queries per tile
constexpr int kLT_SC = kLexTileSlices;  // slices per chunk
constexpr int kLT_MaxStages = 6;      // ring depth: as many stages as fit next to acc (small-G shapes have small stages and need depth)
constexpr int kLT_Threads = kLT_PT + 32;
constexpr int kLT_CtasPerSm = 1;
static_assert(kLT_SC == 4, "the flattened walk selects among the 4 slices of a chunk");

// entry = {acc byte offset of the query (q * PT * 4), G fp16 values, zero pad}; sized so one entry is one or two
// vector loads (8 / 16 / 16+8 bytes)
__host__ __device__ constexpr int lt_entry_words(int G) { return G <= 2 ? 2 : (G <= 6 ? 4 : 6); }
__host__ __device__ constexpr int lt_pval_words(int G) { return (G + 1) / 2; }


// Two bucket-lookup layouts, chosen by the index range rt (= largest stored code + 1):
//  * narrow (rt <= 254, e.g. DeLADE's 39 strides, uniCOIL/SPLADE int8): 8-bit codes in the tiled copy, direct table
//    tab[slice][code] of rt + 1 words;
//  * wide (rt up to 65534, e.g. densified BM25 with idx < 3466): 16-bit codes, and per slice the short list of DISTINCT
//    query codes of the tile {n, (code, table word) x n} that every passage thread compares against (warp-uniform loop;
//    a BM25 query tile has ~2 non-empty queries per slice).
constexpr int kLT_WideSliceBytes = 16 + kLT_QT * 8;

constexpr size_t kLT_StaticSmem = kLT_QT * 4 + 2048;
constexpr size_t kLT_PreferredFootprint = (size_t)204 * 1024;
constexpr size_t kLT_SmemBudget = (size_t)(227 * 1024) / kLT_CtasPerSm - 1024;   // per CTA (1 KiB reserved per CTA by the driver)

LexTileGeom lex_tile_geom(const Geometry& g, int rt) {
    LexTileGeom t{};
    t.G = g.G; t.code_bytes = g.code_bytes; t.n_chunks = g.S_pad / kLT_SC; t.rt = rt;
    t.wide = rt > 254 ? 1 : 0;
    t.tcode_bytes = t.wide ? 2 : 1;
    t.pblock_bytes = kLT_PT * kLT_SC * (t.tcode_bytes + 2 * g.G);
    t.qoff_bytes = t.wide ? kLT_SC * kLT_WideSliceBytes
                          : (int)round_up((int64_t)kLT_SC * (rt + 1) * 4, 16);   // rt buckets + one empty bucket (clamp target) per slice
    t.qblock_stride = t.qoff_bytes + kLT_SC * kLT_QT * lt_entry_words(g.G) * 4;
    t.stage_bytes = (int)round_up(t.pblock_bytes, 128) + (int)round_up(t.qblock_stride, 128);
    const size_t fixed = (size_t)kLT_QT * kLT_PT * 4 + 128 + kLT_StaticSmem;
    t.n_stages = kLT_MaxStages;
    while (t.n_stages > 2 && fixed + (size_t)t.n_stages * t.stage_bytes > kLT_SmemBudget) --t.n_stages;
    // leave >= 24 KiB of the SM's unified shared memory / L1 to the cache (scratch reads, candidate appends): at the config-2
    // shape two 31 KB stages beat three by 1.6 % (10.30 k vs 10.13 k q/s), shapes with smaller stages are unaffected
    while (t.n_stages > 2 && fixed + (size_t)t.n_stages * t.stage_bytes > kLT_PreferredFootprint) --t.n_stages;
    return t;
}

size_t lex_tile_smem_bytes(const LexTileGeom& t) {
    return (size_t)kLT_QT * kLT_PT * 4 + (size_t)t.n_stages * t.stage_bytes + 128;
}

size_t lex_tile_cta_footprint(const LexTileGeom& t, int n_stages) {
    return (size_t)kLT_QT * kLT_PT * 4 + (size_t)n_stages * t.stage_bytes + 128 + kLT_StaticSmem + 1024;
}

bool lex_tile_supported(const Geometry& g, int rt) {
    if (g.S_pad <= 0 || g.S_pad % kLT_SC != 0 || rt < 1 || rt > 65534 || g.G > 8) return false;
    if (rt > 254 && g.code_bytes != 2) return false;
    return lex_tile_smem_bytes(lex_tile_geom(g, rt)) + kLT_StaticSmem <= kLT_SmemBudget;
}

// ---- query-tile preparation: one CTA per (chunk, query tile) builds table + entries -------------
// Entries of a bucket are ordered by query id, buckets by (slice, code).
template <typename CodeT>
__global__ void __launch_bounds__(256)
lex_tile_prep_kernel(const __half* __restrict__ q_lex16, const CodeT* __restrict__ q_code, int n_queries, int S_pad, int G,
                     int rt, int qoff_bytes, int qblock_stride, uint8_t* __restrict__ qblocks, uint32_t* __restrict__ qblock_bytes) {
    extern __shared__ uint32_t prep_smem[];        // cnt[SC][rt + 1] | start[SC][rt + 1] (start doubles as the placement cursor)
    __shared__ uint32_t total_s;
    const int chunk = blockIdx.x, qt = blockIdx.y;
    const int n_chunks = gridDim.x;
    const int q0 = qt * kLT_QT;
    const int nq = min(kLT_QT, n_queries - q0);
    const int per = rt + 1;
    const int tbl = kLT_SC * per;
    uint32_t* cnt = prep_smem;
    uint32_t* start = prep_smem + tbl;
    const int EW = lt_entry_words(G);
    for (int i = threadIdx.x; i < tbl; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nq * kLT_SC; i += blockDim.x) {
        const int q = i / kLT_SC, j = i % kLT_SC;
        const uint32_t code = q_code[(size_t)(q0 + q) * S_pad + chunk * kLT_SC + j];
        if (code < (uint32_t)rt) atomicAdd(&cnt[j * per + code], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {                        // exclusive scan over (slice, code); bucket rt of every slice stays empty
        uint32_t run = 0;
        for (int i = 0; i < tbl; ++i) { start[i] = run; run += cnt[i]; }
        total_s = run;
    }
    __syncthreads();
    uint8_t* blk = qblocks + ((size_t)qt * n_chunks + chunk) * qblock_stride;
    uint32_t* tab = (uint32_t*)blk;
    for (int i = threadIdx.x; i < tbl; i += blockDim.x) tab[i] = (start[i] * (uint32_t)EW * 4u) | (cnt[i] << 16);
    __syncthreads();
    uint32_t* ent = (uint32_t*)(blk + qoff_bytes);
    if (threadIdx.x < kLT_SC) {                    // placement, sequential in q so that buckets are ordered by query id
        const int j = threadIdx.x;
        const int s = chunk * kLT_SC + j;
        for (int q = 0; q < nq; ++q) {
            const uint32_t code = q_code[(size_t)(q0 + q) * S_pad + s];
            if (code >= (uint32_t)rt) continue;
            const uint32_t pos = start[j * per + code]++;
            const __half* v = q_lex16 + ((size_t)(q0 + q) * S_pad + s) * G;
            uint32_t* e = ent + (size_t)pos * EW;
            e[0] = (uint32_t)q * (uint32_t)(kLT_PT * 4);
            for (int w = 1; w < EW; ++w) {         // value g sits in half 2 + g of the entry
                const int g0 = 2 * (w - 1), g1 = g0 + 1;
                const uint32_t lo = g0 < G ? __half_as_ushort(v[g0]) : 0u;
                const uint32_t hi = g1 < G ? __half_as_ushort(v[g1]) : 0u;
                e[w] = lo | (hi << 16);
            }
        }
    }
    if (threadIdx.x == 0)
        qblock_bytes[(size_t)qt * n_chunks + chunk] = ((uint32_t)qoff_bytes + total_s * (uint32_t)EW * 4u + 15u) & ~15u;
}

// wide layout: one warp per (chunk, query tile); lane j < 4 builds slice j sequentially (64 queries, a handful of codes)
template <typename CodeT>
__global__ void __launch_bounds__(32)
lex_tile_prep_wide_kernel(const __half* __restrict__ q_lex16, const CodeT* __restrict__ q_code, int n_queries, int S_pad, int G,
                          int qoff_bytes, int qblock_stride, uint8_t* __restrict__ qblocks, uint32_t* __restrict__ qblock_bytes) {
    __shared__ uint32_t dcode[kLT_SC][kLT_QT], dcnt[kLT_SC][kLT_QT], dpos[kLT_SC][kLT_QT];
    __shared__ uint32_t nd[kLT_SC], tot[kLT_SC];
    const int chunk = blockIdx.x, qt = blockIdx.y;
    const int n_chunks = gridDim.x;
    const int q0 = qt * kLT_QT;
    const int nq = min(kLT_QT, n_queries - q0);
    const int EW = lt_entry_words(G);
    const int j = threadIdx.x;
    uint8_t* blk = qblocks + ((size_t)qt * n_chunks + chunk) * qblock_stride;
    if (j < kLT_SC) {
        const int s = chunk * kLT_SC + j;
        uint32_t n = 0, total = 0;
        for (int q = 0; q < nq; ++q) {                 // distinct codes of the slice, in order of first appearance
            const uint32_t code = q_code[(size_t)(q0 + q) * S_pad + s];
            if (code > CodeTraits<uint16_t>::kMax) continue;
            uint32_t i = 0;
            while (i < n && dcode[j][i] != code) ++i;
            if (i == n) { dcode[j][n] = code; dcnt[j][n] = 0; ++n; }
            ++dcnt[j][i]; ++total;
        }
        nd[j] = n; tot[j] = total;
    }
    __syncwarp();
    if (j < kLT_SC) {
        const int s = chunk * kLT_SC + j;
        uint32_t base = 0;
        for (int jj = 0; jj < j; ++jj) base += tot[jj];
        uint32_t* hdr = (uint32_t*)(blk + (size_t)j * kLT_WideSliceBytes);
        hdr[0] = nd[j];
        uint32_t run = base;
        for (uint32_t i = 0; i < nd[j]; ++i) {
            hdr[4 + 2 * i] = dcode[j][i];
            hdr[4 + 2 * i + 1] = (run * (uint32_t)EW * 4u) | (dcnt[j][i] << 16);
            dpos[j][i] = run; run += dcnt[j][i];
        }
        uint32_t* ent = (uint32_t*)(blk + qoff_bytes);
        for (int q = 0; q < nq; ++q) {                 // placement in query order: buckets stay ordered by query id
            const uint32_t code = q_code[(size_t)(q0 + q) * S_pad + s];
            if (code > CodeTraits<uint16_t>::kMax) continue;
            uint32_t i = 0;
            while (dcode[j][i] != code) ++i;
            const uint32_t pos = dpos[j][i]++;
            const __half* v = q_lex16 + ((size_t)(q0 + q) * S_pad + s) * G;
            uint32_t* e = ent + (size_t)pos * EW;
            e[0] = (uint32_t)q * (uint32_t)(kLT_PT * 4);
            for (int w = 1; w < EW; ++w) {
                const int g0 = 2 * (w - 1), g1 = g0 + 1;
                const uint32_t lo = g0 < G ? __half_as_ushort(v[g0]) : 0u;
                const uint32_t hi = g1 < G ? __half_as_ushort(v[g1]) : 0u;
                e[w] = lo | (hi << 16);
            }
        }
    }
    __syncwarp();
    if (j == 0) {
        uint32_t total = 0;
        for (int jj = 0; jj < kLT_SC; ++jj) total += tot[jj];
        qblock_bytes[(size_t)qt * n_chunks + chunk] = ((uint32_t)qoff_bytes + total * (uint32_t)EW * 4u + 15u) & ~15u;
    }
}

// f32 += f16 * f16 with independent half selection of both operands
template <bool AHI, bool BHI>
__device__ __forceinline__ float fma_h_sel(uint32_t a, uint32_t b, float c) {
    float d;
    if constexpr (AHI && BHI)
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bh, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    else if constexpr (AHI && !BHI)
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bl, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    else if constexpr (!AHI && BHI)
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bh, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    else
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bl, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    return d;
}

// sum over g = g0, g0 + 2, ... < G of entry value g (half 2 + g of the entry) * passage value g: one of the two
// independent FMA chains of a match (even and odd g)
template <int G, int g>
__device__ __forceinline__ float entry_chain(const uint32_t* e, const uint32_t* pv, float t) {
    if constexpr (g < G) {
        t = fma_h_sel<(g & 1) != 0, (g & 1) != 0>(e[1 + (g >> 1)], pv[g >> 1], t);
        return entry_chain<G, g + 2>(e, pv, t);
    } else {
        return t;
    }
}

struct LexTileArgs {
    const uint8_t* lext;               // tiled corpus blocks [tile][chunk]
    const uint8_t* qblocks;            // query blocks [qtile][chunk] (stride qblock_stride)
    const uint32_t* qblock_bytes;      // bytes to copy per query block
    const uint32_t* pblock_nbytes;     // postings layout: bytes to copy per (tile, chunk) block
    long long row_begin, row_end;      // rows handled by this launch (row_begin multiple of the tile size)
    long long n_rows;
    int n_tiles;                       // tiles in [row_begin, row_end)
    int n_chunks, rt, n_stages;
    int pblock_bytes, qoff_bytes, qblock_stride, stage_bytes, pblock_smem;
    int n_qtiles;                      // query tiles in flight
    int n_queries;                     // valid in-flight queries (slots)
    const float* scratch;              // dense scores [row - scratch_row0][scratch_slots] or nullptr
    long long scratch_slots; long long scratch_row0;
    float* tau; uint32_t* cnt; float* cand_score; int32_t* cand_row; int cap;
};

// ---- predicated shared-memory access (no branches inside a match level) ---------------------------
// The loads of read-only stage data are plain (non-volatile) asm so the compiler may schedule them freely -- their
// addresses depend on data read after the stage's mbarrier wait, which orders them; the accumulator read-modify-write
// is volatile asm and keeps its program order.
template <int EW>
__device__ __forceinline__ void lds_entry(uint32_t addr, uint32_t act, uint32_t (&ew)[EW]) {
    static_assert(EW == 2 || EW == 4 || EW == 6, "entry words");
    if constexpr (EW == 2)
        asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0, %1}, [%2];\n\t}"
            : "=r"(ew[0]), "=r"(ew[1]) : "r"(addr), "r"(act));
    else if constexpr (EW == 4)
        asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
            : "=r"(ew[0]), "=r"(ew[1]), "=r"(ew[2]), "=r"(ew[3]) : "r"(addr), "r"(act));
    else
        asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %7, 0;\n\t@q ld.shared.v2.u32 {%0, %1}, [%6];\n\t@q ld.shared.v2.u32 {%2, %3}, [%6+8];\n\t"
            "@q ld.shared.v2.u32 {%4, %5}, [%6+16];\n\t}"      // 24-byte entries are only 8-byte aligned
            : "=r"(ew[0]), "=r"(ew[1]), "=r"(ew[2]), "=r"(ew[3]), "=r"(ew[4]), "=r"(ew[5]) : "r"(addr), "r"(act));
}
__device__ __forceinline__ uint32_t lds_u32_pred(uint32_t addr, uint32_t act) {
    uint32_t v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}" : "=r"(v) : "r"(addr), "r"(act));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16_pred(uint32_t addr, uint32_t act) {
    uint32_t v;
    asm("{\n\t.reg .pred q;\n\t.reg .b16 h;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.b16 h, 0;\n\t@q ld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}"
        : "=r"(v) : "r"(addr), "r"(act));
    return v;
}
// passage values of one (slice, passage): G fp16 at `addr` (4-byte aligned when G is even, 2-byte otherwise)
template <int G>
__device__ __forceinline__ void lds_pvals(uint32_t addr, uint32_t act, uint32_t (&pv)[lt_pval_words(G)]) {
    constexpr int PW = lt_pval_words(G);
    if constexpr (G % 2 == 0) {
#pragma unroll
        for (int w = 0; w < PW; ++w) pv[w] = lds_u32_pred(addr + 4 * w, act);
    } else {
#pragma unroll
        for (int w = 0; w < PW; ++w) {
            const uint32_t lo = lds_u16_pred(addr + 4 * w, act);
            const uint32_t hi = (2 * w + 1 < G) ? lds_u16_pred(addr + 4 * w + 2, act) : 0u;
            pv[w] = lo | (hi << 16);
        }
    }
}
__device__ __forceinline__ float lds_acc_pred(uint32_t addr, uint32_t act) {
    float v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.f32 %0, [%1];\n\t}" : "=f"(v) : "r"(addr), "r"(act) : "memory");
    return v;
}
__device__ __forceinline__ void sts_acc_pred(uint32_t addr, float v, uint32_t act) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.f32 [%0], %1;\n\t}" :: "r"(addr), "f"(v), "r"(act) : "memory");
}

// ---- the same with the predicate `i < n` formed inside the asm block (no materialised 0/1 flag: ptxas merges the compares) ----
template <int EW>
__device__ __forceinline__ void lds_entry_lt(uint32_t addr, uint32_t i, uint32_t n, uint32_t (&ew)[EW]) {
    static_assert(EW == 2 || EW == 4 || EW == 6, "entry words");
    if constexpr (EW == 2)
        asm("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %3, %4;\n\t@q ld.shared.v2.u32 {%0, %1}, [%2];\n\t}"
            : "=r"(ew[0]), "=r"(ew[1]) : "r"(addr), "r"(i), "r"(n));
    else if constexpr (EW == 4)
        asm("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %5, %6;\n\t@q ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
            : "=r"(ew[0]), "=r"(ew[1]), "=r"(ew[2]), "=r"(ew[3]) : "r"(addr), "r"(i), "r"(n));
    else
        asm("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %7, %8;\n\t@q ld.shared.v2.u32 {%0, %1}, [%6];\n\t@q ld.shared.v2.u32 {%2, %3}, [%6+8];\n\t"
            "@q ld.shared.v2.u32 {%4, %5}, [%6+16];\n\t}"
            : "=r"(ew[0]), "=r"(ew[1]), "=r"(ew[2]), "=r"(ew[3]), "=r"(ew[4]), "=r"(ew[5]) : "r"(addr), "r"(i), "r"(n));
}
__device__ __forceinline__ float lds_acc_lt(uint32_t addr, uint32_t i, uint32_t n) {
    float v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q ld.shared.f32 %0, [%1];\n\t}" : "=f"(v) : "r"(addr), "r"(i), "r"(n) : "memory");
    return v;
}
__device__ __forceinline__ void sts_acc_lt(uint32_t addr, float v, uint32_t i, uint32_t n) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q st.shared.f32 [%0], %1;\n\t}" :: "r"(addr), "f"(v), "r"(i), "r"(n) : "memory");
}
template <int G>
struct LexLevel {                       // operands of one match level of one thread
    uint32_t ew[lt_entry_words(G)];
    uint32_t pv[lt_pval_words(G)];
    uint32_t act;
};

template <int G, bool WIDE>
__global__ void __launch_bounds__(kLT_Threads, kLT_CtasPerSm) lex_tile_kernel(const __grid_constant__ LexTileArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kLT_MaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kLT_MaxStages];
    __shared__ __align__(16) float tau_s[kLT_QT];

    constexpr int EW = lt_entry_words(G);
    constexpr uint32_t ES = EW * 4;                                        // entry bytes
    float* acc = (float*)smem;                                            // [QT][PT]
    uint8_t* stages = smem + (size_t)kLT_QT * kLT_PT * 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % a.n_qtiles;
    const int cta_in_q = blockIdx.x / a.n_qtiles;
    const int ctas_per_q = gridDim.x / a.n_qtiles;
    const int q0 = qt * kLT_QT;
    const int nq = min(kLT_QT, a.n_queries - q0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kLT_PT / 32); }
        mbar_fence_init();
    }
    if (threadIdx.x < kLT_QT) tau_s[threadIdx.x] = threadIdx.x < nq ? a.tau[q0 + threadIdx.x] : INFINITY;
    __syncthreads();

    const long long tile0 = a.row_begin / kLT_PT;

    if (warp == kLT_PT / 32) {
        // ===== producer warp =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q) {
                const uint8_t* ptile = a.lext + (size_t)(tile0 + t) * a.n_chunks * a.pblock_bytes;
                for (int c = 0; c < a.n_chunks; ++c) {
                    const uint32_t qb = __ldg(a.qblock_bytes + (size_t)qt * a.n_chunks + c);
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    uint8_t* dst = stages + (size_t)s * a.stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)a.pblock_bytes + qb);
                    bulk_g2s(dst, ptile + (size_t)c * a.pblock_bytes, (uint32_t)a.pblock_bytes, &full_bar[s]);
                    bulk_g2s(dst + a.pblock_smem, a.qblocks + ((size_t)qt * a.n_chunks + c) * a.qblock_stride, qb, &full_bar[s]);
                    if (++s == a.n_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumers: thread p owns passage p of the tile =====
    const int p = threadIdx.x;
    const uint32_t per = (uint32_t)a.rt + 1u;
    const uint32_t acc_p = smem_u32(acc) + (uint32_t)p * 4u;               // shared address of acc[0][p]
    const uint32_t stages_s = smem_u32(stages);
    int s = 0; uint32_t ph = 0;
    for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q) {
        const long long row = (tile0 + t) * kLT_PT + p;
        const bool row_ok = row >= a.row_begin && row < a.row_end && row < a.n_rows;
        // acc init: dense scores of (row, q0..q0+63) written by K2 (256 contiguous bytes per row) or zero;
        // 8 x 128-bit loads are in flight at a time.  Slots beyond nq hold zeros (TMA zero-fills missing queries).
        if (a.scratch && row_ok) {
            const float4* src = (const float4*)(a.scratch + (size_t)(row - a.scratch_row0) * a.scratch_slots + q0);
#pragma unroll 1
            for (int qb = 0; qb < kLT_QT / 4; qb += 8) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __ldcs(src + qb + i);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float* d = acc + (size_t)(4 * (qb + i)) * kLT_PT + p;
                    d[0] = v[i].x; d[kLT_PT] = v[i].y; d[2 * kLT_PT] = v[i].z; d[3 * kLT_PT] = v[i].w;
                }
            }
        } else {
#pragma unroll 16
            for (int q = 0; q < kLT_QT; ++q) acc[q * kLT_PT + p] = 0.f;
        }
        for (int c = 0; c < a.n_chunks; ++c) {
            mbar_wait(&full_bar[s], ph);
            const uint8_t* st = stages + (size_t)s * a.stage_bytes;
            const uint32_t st_s = stages_s + (uint32_t)s * (uint32_t)a.stage_bytes;
            // ---- bucket lookups (branch-free) ----
            // narrow: the table has rt + 1 words per slice (the last one an empty bucket), so clamping the code to rt
            // resolves CODE_EMPTY without a branch.  wide: compare against the slice's distinct query codes (warp-uniform loop).
            // ea[j] / pa[j] are the shared addresses a match of slice j reads: entry of this thread's l-th match =
            // ea[slice(l)] + l * ES.
            uint32_t ea[kLT_SC], pa[kLT_SC], cum[kLT_SC + 1];
            const uint32_t ent_s = st_s + (uint32_t)a.pblock_smem + (uint32_t)a.qoff_bytes;
            const uint32_t pv_s = st_s + (uint32_t)(kLT_PT * kLT_SC * (WIDE ? 2 : 1)) + (uint32_t)p * (uint32_t)(G * 2);
            uint32_t tw[kLT_SC];
            if constexpr (WIDE) {
                const uint2 cw = *(const uint2*)(st + (size_t)p * (kLT_SC * 2));        // my four slice codes (16 bits each)
#pragma unroll
                for (int j = 0; j < kLT_SC; ++j) {
                    const uint32_t code = ((j < 2 ? cw.x : cw.y) >> (16 * (j & 1))) & 0xFFFFu;
                    const uint32_t* hdr = (const uint32_t*)(st + a.pblock_smem + j * kLT_WideSliceBytes);
                    const uint32_t n = hdr[0];
                    uint32_t w = 0;
                    for (uint32_t i = 0; i < n; ++i) {
                        const uint2 pr = *(const uint2*)(hdr + 4 + 2 * i);
                        w = pr.x == code ? pr.y : w;
                    }
                    tw[j] = w;
                }
            } else {
                const uint32_t* tab = (const uint32_t*)(st + a.pblock_smem);
                const uint32_t cw = *(const uint32_t*)(st + (size_t)p * kLT_SC);       // my four slice codes (one byte each)
#pragma unroll
                for (int j = 0; j < kLT_SC; ++j) {                           // empty slices store code rt = the always-empty bucket (no clamp)
                    const uint32_t code = __byte_perm(cw, 0u, 0x4440u + j);
                    tw[j] = tab[j * per + code];
                }
            }
            cum[0] = 0;
#pragma unroll
            for (int j = 0; j < kLT_SC; ++j) {
                ea[j] = ent_s + (tw[j] & 0xFFFFu) - cum[j] * ES;
                pa[j] = pv_s + (uint32_t)(j * kLT_PT * G * 2);
                cum[j + 1] = cum[j] + (tw[j] >> 16);
            }
            const uint32_t mine = cum[kLT_SC];
            const uint32_t levels = __reduce_max_sync(0xFFFFFFFFu, mine);

            // ---- flattened, software-pipelined match walk ----
            auto fetch = [&](uint32_t l, LexLevel<G>& o) {
                const bool g1 = l >= cum[1], g2 = l >= cum[2], g3 = l >= cum[3];
                const uint32_t e = g3 ? ea[3] : (g2 ? ea[2] : (g1 ? ea[1] : ea[0]));
                const uint32_t v = g3 ? pa[3] : (g2 ? pa[2] : (g1 ? pa[1] : pa[0]));
                o.act = l < mine ? 1u : 0u;
                lds_entry<EW>(e + l * ES, o.act, o.ew);
                lds_pvals<G>(v, o.act, o.pv);
            };
            auto apply = [&](const LexLevel<G>& o) {
                const uint32_t addr = acc_p + o.ew[0];
                const float cur = lds_acc_pred(addr, o.act);
                const float te = entry_chain<G, 0>(o.ew, o.pv, 0.f);
                const float to = entry_chain<G, 1>(o.ew, o.pv, 0.f);
                sts_acc_pred(addr, cur + (te + to), o.act);
            };
            LexLevel<G> x, y;
            fetch(0, x);
#pragma unroll 1
            for (uint32_t l = 0; l < levels; l += 2) {
                fetch(l + 1, y);
                apply(x);
                if (l + 1 >= levels) break;
                fetch(l + 2, x);
                apply(y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == a.n_stages) { s = 0; ph ^= 1u; }
        }
        // admission filter
        if (row_ok) {
#pragma unroll 1
            for (int qb = 0; qb < nq; qb += 4) {
                const float4 tq = *(const float4*)(tau_s + qb);                    // tau_s is +inf beyond nq
                const float tv[4] = {tq.x, tq.y, tq.z, tq.w};
                float sv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) sv[i] = acc[(qb + i) * kLT_PT + p] + 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (sv[i] > tv[i]) {
                        const int slot = q0 + qb + i;
                        const uint32_t pos = atomicAdd(a.cnt + slot, 1u);
                        if (pos < (uint32_t)a.cap) {
                            a.cand_score[(size_t)slot * a.cap + pos] = sv[i];
                            a.cand_row[(size_t)slot * a.cap + pos] = (int32_t)row;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// =====================================================================================================
// K1p: the same tile (64 queries x 512 passages) walked from a POSTINGS layout of the corpus tile.
//
// K1t gives every passage a thread and lets it walk the queries of its bucket; per-thread match counts are
// Poisson, so ~40 % of the lanes of a level work, and every thread pays the lookup of 4 slices (~100
// instructions per chunk) whether they match or not.  Here the index stores, per (tile, slice), only the
// NON-EMPTY passages, sorted by their slice code: item = {passage u16 | code u8, G fp16 values}.  A warp takes
// 32 consecutive items of the slice's list: its lanes hold the same or neighbouring codes, so they read the
// SAME bucket (broadcast loads of table word and entries), run the same number of levels (bucket size), and no
// lane is spent on an empty slice.  Every level is: entry (broadcast) -> acc[q][passage] += dot.  Within a slice
// the passages of all items are distinct, so the warps of the CTA never touch the same accumulator; between
// slices a CTA barrier orders the read-modify-writes of one (query, passage) pair (fixed summation order: by
// slice).  The table word and the item of the next slice are loaded before the barrier (stage data is
// read-only), which takes their latency off the per-slice critical path.
// =====================================================================================================

// ---- K1p pipeline stages (force-inlined; every slice of a chunk has its own PItem / PRun registers) ----------------------------
template <int EW> struct PItem { uint32_t it[EW]; uint32_t tw; };
struct PRun { uint32_t cnt, eb, accp, M; };
constexpr int kLP_Group = 4;               // bucket entries handled as one group of independent accumulator updates

template <int EW>
__device__ __forceinline__ void k1p_item(uint32_t st_s, uint32_t hdr, uint32_t p, PItem<EW>& o) {         // S1
    lds_entry_lt<EW>(st_s + 16u + ((hdr >> 16) + p) * (uint32_t)(EW * 4), p, hdr & 0xFFFFu, o.it);
}
template <int EW>
__device__ __forceinline__ void k1p_table(uint32_t tab_s, uint32_t slice_off, uint32_t hdr, uint32_t p, PItem<EW>& o) {   // S2
    const uint32_t code = (o.it[0] >> 16) & 0xFFu;
    o.tw = 0u;                                                              // slots beyond the item count: empty bucket
    asm("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q ld.shared.u32 %0, [%1];\n\t}"
        : "+r"(o.tw) : "r"(tab_s + (slice_off + code) * 4u), "r"(p), "r"(hdr & 0xFFFFu));
}
template <int EW>
__device__ __forceinline__ void k1p_load_group(const PRun& r, uint32_t i0, uint32_t (&e)[kLP_Group][EW]) {
#pragma unroll
    for (int u = 0; u < kLP_Group; ++u) lds_entry_lt<EW>(r.eb + (i0 + u) * (uint32_t)(EW * 4), i0 + u, r.cnt, e[u]);
}
template <int G, int EW>
__device__ __forceinline__ void k1p_apply_group(const PRun& r, const PItem<EW>& item, uint32_t i0, const uint32_t (&e)[kLP_Group][EW]) {
    uint32_t ad[kLP_Group]; float cur[kLP_Group], res[kLP_Group];
#pragma unroll
    for (int u = 0; u < kLP_Group; ++u) { ad[u] = r.accp + e[u][0]; cur[u] = lds_acc_lt(ad[u], i0 + u, r.cnt); }
#pragma unroll
    for (int u = 0; u < kLP_Group; ++u) res[u] = entry_chain<G, 0>(e[u], item.it + 1, cur[u]) + entry_chain<G, 1>(e[u], item.it + 1, 0.f);
#pragma unroll
    for (int u = 0; u < kLP_Group; ++u) sts_acc_lt(ad[u], res[u], i0 + u, r.cnt);
}
template <int EW>
__device__ __forceinline__ void k1p_bucket(uint32_t ent_s, uint32_t acc_s, const PItem<EW>& item, PRun& r, uint32_t (&e)[kLP_Group][EW]) {   // S3
    r.cnt = item.tw >> 16;
    r.eb = ent_s + (item.tw & 0xFFFFu);
    r.accp = acc_s + (item.it[0] & 0xFFFFu) * 4u;                          // acc[0][passage of the item]
    r.M = __reduce_max_sync(0xFFFFFFFFu, r.cnt);
    k1p_load_group<EW>(r, 0u, e);
}
template <int G, int EW>
__device__ __forceinline__ void k1p_update(const PRun& r, const PItem<EW>& item, const uint32_t (&e)[kLP_Group][EW]) {      // S4
    k1p_apply_group<G, EW>(r, item, 0u, e);
#pragma unroll 1
    for (uint32_t i = kLP_Group; i < r.M; ++i) {                          // larger buckets (warp-uniform trip count, rare): one entry at a time
        uint32_t e1[EW];
        lds_entry_lt<EW>(r.eb + i * (uint32_t)(EW * 4), i, r.cnt, e1);
        const uint32_t ad = r.accp + e1[0];
        const float cur = lds_acc_lt(ad, i, r.cnt);
        sts_acc_lt(ad, entry_chain<G, 0>(e1, item.it + 1, cur) + entry_chain<G, 1>(e1, item.it + 1, 0.f), i, r.cnt);
    }
}

constexpr int kLP_Threads = kLT_PT;     // 512: no separate producer warp (a 544-thread CTA is capped at 96 registers per thread, 512 at 128)

template <int G>
__global__ void __launch_bounds__(kLP_Threads, 1) lex_post_kernel(const __grid_constant__ LexTileArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kLT_MaxStages];
    __shared__ __align__(16) float tau_s[kLT_QT];

    constexpr int EW = lt_entry_words(G);                                  // entry and item share the format {word 0, G fp16}
    float* acc = (float*)smem;                                            // [QT][PT]
    uint8_t* stages = smem + (size_t)kLT_QT * kLT_PT * 4;

    const int qt = blockIdx.x % a.n_qtiles;
    const int cta_in_q = blockIdx.x / a.n_qtiles;
    const int ctas_per_q = gridDim.x / a.n_qtiles;
    const int q0 = qt * kLT_QT;
    const int nq = min(kLT_QT, a.n_queries - q0);
    const int p = threadIdx.x;                                             // passage of this thread for init / filter; item slot for the walk
    const uint32_t up = (uint32_t)p;

    if (p == 0) {
        for (int s = 0; s < a.n_stages; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    if (p < kLT_QT) tau_s[p] = p < nq ? a.tau[q0 + p] : INFINITY;
    __syncthreads();

    const long long tile0 = a.row_begin / kLT_PT;
    const int my_tiles = cta_in_q < a.n_tiles ? (a.n_tiles - cta_in_q + ctas_per_q - 1) / ctas_per_q : 0;
    const int n_work = my_tiles * a.n_chunks;                               // work item w = (tile w / n_chunks of this CTA, chunk w % n_chunks)
    // The CTA is its own producer: all warps pass a CTA barrier after every slice, so when a chunk is done nobody reads its
    // stage any more and thread 0 refills it with the chunk n_stages ahead (TMA bulk copies, completion on the stage's mbarrier).
    auto issue = [&](int w) {
        const int t = cta_in_q + (w / a.n_chunks) * ctas_per_q, c = w % a.n_chunks, s = w % a.n_stages;
        const size_t blk = (size_t)(tile0 + t) * a.n_chunks + c;
        const uint32_t qb = __ldg(a.qblock_bytes + (size_t)qt * a.n_chunks + c);
        const uint32_t pb = __ldg(a.pblock_nbytes + blk);
        uint8_t* dst = stages + (size_t)s * a.stage_bytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the stage's generic-proxy reads are done (barrier) before TMA rewrites it
        mbar_arrive_expect_tx(&full_bar[s], pb + qb);
        bulk_g2s(dst, a.lext + blk * (size_t)a.pblock_bytes, pb, &full_bar[s]);
        bulk_g2s(dst + a.pblock_smem, a.qblocks + ((size_t)qt * a.n_chunks + c) * a.qblock_stride, qb, &full_bar[s]);
    };
    if (p == 0)
        for (int w = 0; w < a.n_stages && w < n_work; ++w) issue(w);

    const uint32_t per = (uint32_t)a.rt + 1u;
    const uint32_t acc_s = smem_u32(acc);
    const uint32_t stages_s = smem_u32(stages);
    int w = 0;
    for (int t = cta_in_q; t < a.n_tiles; t += ctas_per_q) {
        const long long row = (tile0 + t) * kLT_PT + p;
        const bool row_ok = row >= a.row_begin && row < a.row_end && row < a.n_rows;
        if (a.scratch && row_ok) {
            const float4* src = (const float4*)(a.scratch + (size_t)(row - a.scratch_row0) * a.scratch_slots + q0);
#pragma unroll 1
            for (int qb = 0; qb < kLT_QT / 4; qb += 8) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __ldcs(src + qb + i);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float* d = acc + (size_t)(4 * (qb + i)) * kLT_PT + p;
                    d[0] = v[i].x; d[kLT_PT] = v[i].y; d[2 * kLT_PT] = v[i].z; d[3 * kLT_PT] = v[i].w;
                }
            }
        } else {
#pragma unroll 16
            for (int q = 0; q < kLT_QT; ++q) acc[q * kLT_PT + p] = 0.f;
        }
        __syncthreads();                                                   // the walk touches every column of acc
        for (int c = 0; c < a.n_chunks; ++c, ++w) {
            const int s = w % a.n_stages;
            mbar_wait(&full_bar[s], (uint32_t)(w / a.n_stages) & 1u);
            const uint32_t st_s = stages_s + (uint32_t)s * (uint32_t)a.stage_bytes;
            const uint32_t tab_s = st_s + (uint32_t)a.pblock_smem;
            const uint32_t ent_s = tab_s + (uint32_t)a.qoff_bytes;
            uint32_t hdr[kLT_SC];                                          // per slice: (first item << 16) | item count
            {   // a C++ load: ordered behind the barrier wait by its memory clobber (every later stage read depends on it)
                const uint4 hv = *(const uint4*)(stages + (size_t)s * a.stage_bytes);
                hdr[0] = hv.x; hdr[1] = hv.y; hdr[2] = hv.z; hdr[3] = hv.w;
            }
            // Software pipeline over the four slices of the chunk.  Per slice: S1 item load -> S2 table word of its code -> S3 bucket
            // size, warp maximum, first group of entries -> S4 accumulator read-modify-writes.  Only S4 touches acc and has to wait
            // for the barrier that ends the previous slice; S1..S3 read stage data only and are issued ahead of their consumer (S3 of
            // the next slice right before the barrier, so its entry loads fly across it): after a barrier the critical path is
            // LDS acc -> FMA chain -> STS.
            PItem<EW> i0, i1, i2, i3;
            PRun r0, r1, r2, r3;
            uint32_t en[kLP_Group][EW];
            k1p_item<EW>(st_s, hdr[0], up, i0);
            k1p_item<EW>(st_s, hdr[1], up, i1);
            k1p_table<EW>(tab_s, 0u, hdr[0], up, i0);
            k1p_bucket<EW>(ent_s, acc_s, i0, r0, en);
            // slice 0
            k1p_item<EW>(st_s, hdr[2], up, i2);
            k1p_table<EW>(tab_s, per, hdr[1], up, i1);
            k1p_update<G, EW>(r0, i0, en);
            k1p_bucket<EW>(ent_s, acc_s, i1, r1, en);
            __syncthreads();
            // slice 1
            k1p_item<EW>(st_s, hdr[3], up, i3);
            k1p_table<EW>(tab_s, 2u * per, hdr[2], up, i2);
            k1p_update<G, EW>(r1, i1, en);
            k1p_bucket<EW>(ent_s, acc_s, i2, r2, en);
            __syncthreads();
            // slice 2
            k1p_table<EW>(tab_s, 3u * per, hdr[3], up, i3);
            k1p_update<G, EW>(r2, i2, en);
            k1p_bucket<EW>(ent_s, acc_s, i3, r3, en);
            __syncthreads();
            // slice 3
            k1p_update<G, EW>(r3, i3, en);
            __syncthreads();                                               // chunk done by every warp: its stage may be refilled
            if (p == 0 && w + a.n_stages < n_work) issue(w + a.n_stages);
        }
        // admission filter (thread p = passage p again; the barrier above ordered all accumulator writes)
        if (row_ok) {
#pragma unroll 1
            for (int qb = 0; qb < nq; qb += 4) {
                const float4 tq = *(const float4*)(tau_s + qb);
                const float tv[4] = {tq.x, tq.y, tq.z, tq.w};
                float sv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) sv[i] = acc[(qb + i) * kLT_PT + p] + 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (sv[i] > tv[i]) {
                        const int slot = q0 + qb + i;
                        const uint32_t pos = atomicAdd(a.cnt + slot, 1u);
                        if (pos < (uint32_t)a.cap) {
                            a.cand_score[(size_t)slot * a.cap + pos] = sv[i];
                            a.cand_row[(size_t)slot * a.cap + pos] = (int32_t)row;
                        }
                    }
                }
            }
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
// postings layout: block = 16-byte header (4 x {first item << 16 | count}) + up to 4 x 512 items of lt_entry_words(G) words
LexTileGeom lex_post_geom(const Geometry& g, int rt) {
    LexTileGeom t = lex_tile_geom(g, rt);
    t.pblock_bytes = (int)round_up(16 + (int64_t)kLT_PT * kLT_SC * lt_entry_words(g.G) * 4, 128);
    t.stage_bytes = t.pblock_bytes + (int)round_up(t.qblock_stride, 128);
    const size_t fixed = (size_t)kLT_QT * kLT_PT * 4 + 128 + kLT_StaticSmem;
    t.n_stages = kLT_MaxStages;
    while (t.n_stages > 1 && fixed + (size_t)t.n_stages * t.stage_bytes > kLT_SmemBudget) --t.n_stages;
    return t;
}

bool lex_post_supported(const Geometry& g, int rt) {
    if (g.S_pad <= 0 || g.S_pad % kLT_SC != 0 || rt < 1 || rt > 254 || g.G > 8) return false;   // 8-bit codes only
    const LexTileGeom t = lex_post_geom(g, rt);
    return t.n_stages >= 2;
}

template <int G>
static int launch_lex_post_t(const dhr_index* h, const LexTileArgs& a, size_t smem, cudaStream_t st) {
    auto kern = lex_post_kernel<G>;
    DHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_q = h->num_sms / a.n_qtiles;
    if (per_q < 1) per_q = 1;
    if (per_q > a.n_tiles) per_q = a.n_tiles;
    kern<<<(unsigned)(per_q * a.n_qtiles), kLP_Threads, smem, st>>>(a);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

int launch_lex_post(const dhr_index* h, const LexTileGeom& t, const uint8_t* qblocks, const uint32_t* qblock_bytes, int n_queries,
                    long long row_begin, long long row_end, const float* scratch, long long scratch_slots, long long scratch_row0,
                    const TopkState& tk, int cap, cudaStream_t st) {
    if (row_end <= row_begin || n_queries <= 0) return DHR_OK;
    if (!h->lexp) return DHR_ERR_STATE;
    LexTileArgs a{};
    a.lext = h->lexp; a.qblocks = qblocks; a.qblock_bytes = qblock_bytes; a.pblock_nbytes = h->lexp_nbytes;
    a.row_begin = row_begin; a.row_end = row_end; a.n_rows = h->n_rows;
    a.n_tiles = (int)((row_end - row_begin / kLT_PT * kLT_PT + kLT_PT - 1) / kLT_PT);
    a.n_chunks = t.n_chunks; a.rt = t.rt; a.n_stages = t.n_stages;
    a.pblock_bytes = t.pblock_bytes; a.qoff_bytes = t.qoff_bytes; a.qblock_stride = t.qblock_stride;
    a.stage_bytes = t.stage_bytes; a.pblock_smem = t.pblock_bytes;
    a.n_qtiles = (n_queries + kLT_QT - 1) / kLT_QT;
    a.n_queries = n_queries;
    a.scratch = scratch; a.scratch_slots = scratch_slots; a.scratch_row0 = scratch_row0;
    a.tau = tk.tau; a.cnt = tk.cnt; a.cand_score = tk.cand_score; a.cand_row = tk.cand_row; a.cap = cap;
    const size_t smem = (size_t)kLT_QT * kLT_PT * 4 + (size_t)t.n_stages * t.stage_bytes + 128;
    switch (h->g.G) {
        case 1: return launch_lex_post_t<1>(h, a, smem, st);
        case 2: return launch_lex_post_t<2>(h, a, smem, st);
        case 3: return launch_lex_post_t<3>(h, a, smem, st);
        case 4: return launch_lex_post_t<4>(h, a, smem, st);
        case 5: return launch_lex_post_t<5>(h, a, smem, st);
        case 6: return launch_lex_post_t<6>(h, a, smem, st);
        case 7: return launch_lex_post_t<7>(h, a, smem, st);
        case 8: return launch_lex_post_t<8>(h, a, smem, st);
        default: return DHR_ERR_UNSUPPORTED;
    }
}

int lex_post_entry_words(int G) { return lt_entry_words(G); }

int launch_lex_tile_prep(const dhr_index* h, const LexTileGeom& t, const void* q_lex16, const void* q_code, int n_queries,
                         uint8_t* qblocks, uint32_t* qblock_bytes, cudaStream_t st) {
    const Geometry& g = h->g;
    const int n_qtiles = (n_queries + kLT_QT - 1) / kLT_QT;
    if (n_qtiles == 0) return DHR_OK;
    dim3 grid((unsigned)t.n_chunks, (unsigned)n_qtiles);
    if (t.wide) {
        if (g.code_bytes != 2) return DHR_ERR_UNSUPPORTED;
        lex_tile_prep_wide_kernel<uint16_t><<<grid, 32, 0, st>>>((const __half*)q_lex16, (const uint16_t*)q_code, n_queries, g.S_pad, g.G,
                                                                 t.qoff_bytes, t.qblock_stride, qblocks, qblock_bytes);
        DHR_CUDA(cudaGetLastError());
        return DHR_OK;
    }
    const size_t smem = (size_t)2 * kLT_SC * (t.rt + 1) * sizeof(uint32_t);
    if (g.code_bytes == 1)
        lex_tile_prep_kernel<uint8_t><<<grid, 256, smem, st>>>((const __half*)q_lex16, (const uint8_t*)q_code, n_queries, g.S_pad, g.G,
                                                               t.rt, t.qoff_bytes, t.qblock_stride, qblocks, qblock_bytes);
    else
        lex_tile_prep_kernel<uint16_t><<<grid, 256, smem, st>>>((const __half*)q_lex16, (const uint16_t*)q_code, n_queries, g.S_pad, g.G,
                                                                t.rt, t.qoff_bytes, t.qblock_stride, qblocks, qblock_bytes);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

template <int G, bool WIDE>
static int launch_lex_tile_t(const dhr_index* h, const LexTileArgs& a, size_t smem, cudaStream_t st) {
    auto kern = lex_tile_kernel<G, WIDE>;
    DHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_q = h->num_sms * kLT_CtasPerSm / a.n_qtiles;
    if (per_q < 1) per_q = 1;
    if (per_q > a.n_tiles) per_q = a.n_tiles;
    kern<<<(unsigned)(per_q * a.n_qtiles), kLT_Threads, smem, st>>>(a);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

template <bool WIDE>
static int launch_lex_tile_g(const dhr_index* h, const LexTileArgs& a, size_t smem, cudaStream_t st) {
    switch (h->g.G) {
        case 1: return launch_lex_tile_t<1, WIDE>(h, a, smem, st);
        case 2: return launch_lex_tile_t<2, WIDE>(h, a, smem, st);
        case 3: return launch_lex_tile_t<3, WIDE>(h, a, smem, st);
        case 4: return launch_lex_tile_t<4, WIDE>(h, a, smem, st);
        case 5: return launch_lex_tile_t<5, WIDE>(h, a, smem, st);
        case 6: return launch_lex_tile_t<6, WIDE>(h, a, smem, st);
        case 7: return launch_lex_tile_t<7, WIDE>(h, a, smem, st);
        case 8: return launch_lex_tile_t<8, WIDE>(h, a, smem, st);
        default: return DHR_ERR_UNSUPPORTED;
    }
}

int launch_lex_tile(const dhr_index* h, const LexTileGeom& t, const uint8_t* qblocks, const uint32_t* qblock_bytes, int n_queries,
                    long long row_begin, long long row_end, const float* scratch, long long scratch_slots, long long scratch_row0,
                    const TopkState& tk, int cap, cudaStream_t st) {
    if (row_end <= row_begin || n_queries <= 0) return DHR_OK;
    LexTileArgs a{};
    a.lext = h->lext; a.qblocks = qblocks; a.qblock_bytes = qblock_bytes;
    a.row_begin = row_begin; a.row_end = row_end; a.n_rows = h->n_rows;
    a.n_tiles = (int)((row_end - row_begin / kLT_PT * kLT_PT + kLT_PT - 1) / kLT_PT);
    a.n_chunks = t.n_chunks; a.rt = t.rt; a.n_stages = t.n_stages;
    a.pblock_bytes = t.pblock_bytes; a.qoff_bytes = t.qoff_bytes; a.qblock_stride = t.qblock_stride;
    a.stage_bytes = t.stage_bytes; a.pblock_smem = (int)round_up(t.pblock_bytes, 128);
    a.n_qtiles = (n_queries + kLT_QT - 1) / kLT_QT;
    a.n_queries = n_queries;
    a.scratch = scratch; a.scratch_slots = scratch_slots; a.scratch_row0 = scratch_row0;
    a.tau = tk.tau; a.cnt = tk.cnt; a.cand_score = tk.cand_score; a.cand_row = tk.cand_row; a.cap = cap;
    if (!h->lext) return DHR_ERR_STATE;
    return t.wide ? launch_lex_tile_g<true>(h, a, lex_tile_smem_bytes(t), st) : launch_lex_tile_g<false>(h, a, lex_tile_smem_bytes(t), st);
}

}  // namespace dhr
