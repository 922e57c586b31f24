// search.cu -- dhr_search / dhr_rerank orchestration: query preparation, chunked scan with a
// tightening admission threshold, exact selection, overflow-proof fallback.
//
// Replaces the body of GIP_retrieval (gip_retrieval.py:88-165, exact branch and the --IP first
// stage) and IP_retrieval (:60-85).  Selection without a score matrix:
//   rows are scanned in chunks of increasing row ranges [b0,b1), [b1,b2), ...; the first chunk
//   (<= cap-k rows) admits every row; after each chunk K3 keeps the best k candidates of each
//   query and publishes tau = k-th best score; later chunks admit a row only if score > tau
//   (strict: later rows have larger row ids, so they lose ties -- exactly the (score desc, row
//   asc) rule).  Chunks grow geometrically so a random-order corpus admits ~k*(growth-1) rows per
//   chunk; if a query still overflows its cap-slot buffer (adversarial order) it is re-run with
//   uniform chunks of cap-k rows, which cannot overflow.
#include <algorithm>
#include <math.h>
#include <vector>

#include "internal.h"

namespace dhr {

// ---- query preparation ------------------------------------------------------------------------
// one thread per (query, padded slice): values in fp16 and fp32, code; flags[2] |= 1 if any value
// is not exactly representable in fp16 (then the scan uses the fp32-query kernels).
template <typename CodeT>
__global__ void prep_lexical_kernel(int n, int S, int G, int S_pad, int val_dtype, const void* vals, long long vstride,
                                    int idx_dtype, const void* idx, long long istride, __half* q16, float* q32, CodeT* qc,
                                    int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * S_pad) return;
    const int q = (int)(i / S_pad), s = (int)(i % S_pad);
    __half* o16 = q16 + ((size_t)q * S_pad + s) * G;
    float* o32 = q32 + ((size_t)q * S_pad + s) * G;
    CodeT* oc = qc + (size_t)q * S_pad + s;
    if (s >= S) {
        for (int g = 0; g < G; ++g) { o16[g] = __float2half_rn(0.f); o32[g] = 0.f; }
        *oc = (CodeT)CodeTraits<CodeT>::kNoMatch;
        return;
    }
    bool nonzero = false, inexact = false;
    for (int g = 0; g < G; ++g) {
        const size_t off = (size_t)q * vstride + (size_t)s * G + g;
        const float f = val_dtype == DHR_VAL_F16 ? __half2float(((const __half*)vals)[off]) : ((const float*)vals)[off];
        const __half hv = __float2half_rn(f);
        o16[g] = hv; o32[g] = f;
        nonzero |= (f != 0.f);
        inexact |= (__half2float(hv) != f);
    }
    if (inexact) atomicOr(flags + 2, 1);
    uint32_t code = CodeTraits<CodeT>::kNoMatch;
    if (nonzero && idx) {
        long long v;
        switch (idx_dtype) {
            case DHR_IDX_U8:  v = ((const uint8_t*)idx)[(size_t)q * istride + s]; break;
            case DHR_IDX_I8:  v = ((const int8_t*)idx)[(size_t)q * istride + s]; break;
            case DHR_IDX_I16: v = ((const int16_t*)idx)[(size_t)q * istride + s]; break;
            case DHR_IDX_U16: v = ((const uint16_t*)idx)[(size_t)q * istride + s]; break;
            case DHR_IDX_I32: v = ((const int32_t*)idx)[(size_t)q * istride + s]; break;
            default:          v = ((const long long*)idx)[(size_t)q * istride + s]; break;
        }
        if (v >= 0 && v <= (long long)CodeTraits<CodeT>::kMax) code = (uint32_t)v;
    }
    *oc = (CodeT)code;
}

__global__ void prep_dense_kernel(int n, int D, int C, int C_pad, int val_dtype, const void* vals, long long vstride,
                                  float lamda, __half* q16, float* q32, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * C_pad) return;
    const int q = (int)(i / C_pad), c = (int)(i % C_pad);
    float f = 0.f;
    if (c < C) {
        const size_t off = (size_t)q * vstride + D + c;
        f = val_dtype == DHR_VAL_F16 ? __half2float(((const __half*)vals)[off]) : ((const float*)vals)[off];
        f = lamda * f;                                   // gip_retrieval.py:281-283, fp32 on the CPU path
    }
    const __half hv = __float2half_rn(f);
    q16[(size_t)q * C_pad + c] = hv;
    q32[(size_t)q * C_pad + c] = f;
    if (__half2float(hv) != f) atomicOr(flags + 2, 1);
}

int launch_prep_queries(dhr_index* h, int n, int val_dtype, const void* d_vals, int64_t vstride, int idx_dtype,
                        const void* d_idx, int64_t istride, float lamda, cudaStream_t st) {
    const Geometry& g = h->g;
    if (g.S_pad > 0) {
        const long long total = (long long)n * g.S_pad;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        if (g.code_bytes == 1)
            prep_lexical_kernel<uint8_t><<<blocks, 256, 0, st>>>(n, g.S, g.G, g.S_pad, val_dtype, d_vals, vstride, idx_dtype, d_idx,
                                                                 istride, (__half*)h->q_lex16, (float*)h->q_lex32,
                                                                 (uint8_t*)h->q_code, h->d_flags);
        else
            prep_lexical_kernel<uint16_t><<<blocks, 256, 0, st>>>(n, g.S, g.G, g.S_pad, val_dtype, d_vals, vstride, idx_dtype, d_idx,
                                                                  istride, (__half*)h->q_lex16, (float*)h->q_lex32,
                                                                  (uint16_t*)h->q_code, h->d_flags);
        DHR_CUDA(cudaGetLastError());
        h->stats.n_prep_launches++; h->stats.n_kernel_launches++;
    }
    if (g.C_pad > 0) {
        const long long total = (long long)n * g.C_pad;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        prep_dense_kernel<<<blocks, 256, 0, st>>>(n, g.S * g.G, g.C, g.C_pad, val_dtype, d_vals, vstride, lamda,
                                                  (__half*)h->q_dns16, (float*)h->q_dns32, h->d_flags);
        DHR_CUDA(cudaGetLastError());
        h->stats.n_prep_launches++; h->stats.n_kernel_launches++;
    }
    return DHR_OK;
}

// Slot state at the start of a call.  Between the batches of a call the final-pass select hands every slot back clean
// (topk.cu), so this runs once per dhr_search / dhr_rerank and not once per batch.
__global__ void init_slots_kernel(TopkState t, int n_slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) { t.tau[i] = -INFINITY; t.cnt[i] = 0u; t.overflow[i] = 0u; }
}
void launch_init_slots(const TopkState& t, cudaStream_t st) { init_slots_kernel<<<1, kMaxInflight, 0, st>>>(t, kMaxInflight); }

// ---- workspace ----------------------------------------------------------------------------------
static int ensure_query_workspace(dhr_index* h, int n) {
    if (n <= h->q_capacity) return DHR_OK;
    const Geometry& g = h->g;
    void** bufs[] = {&h->q_lex16, &h->q_lex32, &h->q_dns16, &h->q_dns32, &h->q_code};
    for (void** b : bufs) if (*b) { cudaFree(*b); *b = nullptr; }
    h->q_capacity = 0;
    const size_t nn = (size_t)n + kMaxInflight;   // slack: a group may read (and discard) rows past the last query
    if (g.D_pad > 0) {
        DHR_CUDA(cudaMalloc(&h->q_lex16, nn * g.D_pad * 2));
        DHR_CUDA(cudaMalloc(&h->q_lex32, nn * g.D_pad * 4));
        DHR_CUDA(cudaMalloc(&h->q_code, nn * g.S_pad * g.code_bytes));
    }
    if (g.C_pad > 0) {
        DHR_CUDA(cudaMalloc(&h->q_dns16, nn * g.C_pad * 2));
        DHR_CUDA(cudaMalloc(&h->q_dns32, nn * g.C_pad * 4));
    }
    h->q_capacity = n;
    return DHR_OK;
}

static int ensure_topk_state(dhr_index* h, int lane = 0) {
    TopkState& t = lane ? h->topk1 : h->topk;
    if (t.tau) return DHR_OK;
    DHR_CUDA(cudaMalloc(&t.tau, kMaxInflight * sizeof(float)));
    DHR_CUDA(cudaMalloc(&t.cnt, kMaxInflight * sizeof(uint32_t)));
    DHR_CUDA(cudaMalloc(&t.overflow, kMaxInflight * sizeof(uint32_t)));
    DHR_CUDA(cudaMalloc(&t.cand_score, (size_t)kMaxInflight * kCandCap * sizeof(float)));
    DHR_CUDA(cudaMalloc(&t.cand_row, (size_t)kMaxInflight * kCandCap * sizeof(int32_t)));
    return DHR_OK;
}

// device staging of host outputs.  The [Q,k] arrays and the [Q] counts grow independently: a later call with more queries and
// a smaller k must not reuse a counts buffer sized for fewer queries.
static int ensure_out_buffers(dhr_index* h, size_t n_queries, int k) {
    const size_t need = n_queries * (size_t)k;
    if (need > h->out_capacity || !h->d_out_scores) {
        if (h->d_out_scores) cudaFree(h->d_out_scores);
        if (h->d_out_rows) cudaFree(h->d_out_rows);
        h->d_out_scores = nullptr; h->d_out_rows = nullptr; h->out_capacity = 0;
        DHR_CUDA(cudaMalloc(&h->d_out_scores, need * sizeof(float)));
        DHR_CUDA(cudaMalloc(&h->d_out_rows, need * sizeof(int64_t)));
        h->out_capacity = need;
    }
    if (n_queries > h->out_q_capacity || !h->d_out_counts) {
        if (h->d_out_counts) cudaFree(h->d_out_counts);
        h->d_out_counts = nullptr; h->out_q_capacity = 0;
        DHR_CUDA(cudaMalloc(&h->d_out_counts, (n_queries + 1) * sizeof(int32_t)));
        h->out_q_capacity = n_queries;
    }
    return DHR_OK;
}

// the paths that append with atomics (K1, K1t, the multi-launch --IP passes, rerank) must not hand K3 the segment pointers
static TopkState without_segments(TopkState t) {
    t.seg_score = nullptr; t.seg_row = nullptr; t.seg_cnt = nullptr;
    return t;
}

// segmented candidate lists of the tensor-core filter epilogue (one set per batch lane), attached to a copy of the lane's TopkState
static int ensure_segments(dhr_index* h, int lane) {
    TopkState& t = lane ? h->topk1 : h->topk;
    if (t.seg_cnt) return DHR_OK;
    const size_t n = (size_t)kMaxInflight * kSegCount * kSegCap;
    DHR_CUDA(cudaMalloc(&t.seg_score, n * sizeof(float)));
    DHR_CUDA(cudaMalloc(&t.seg_row, n * sizeof(int32_t)));
    DHR_CUDA(cudaMalloc(&t.seg_cnt, (size_t)kMaxInflight * kSegCount * sizeof(uint32_t)));
    DHR_CUDA(cudaMemset(t.seg_cnt, 0, (size_t)kMaxInflight * kSegCount * sizeof(uint32_t)));
    return DHR_OK;
}

static int ensure_overflow_flags(dhr_index* h, size_t n_queries) {
    if (n_queries <= h->overflow_capacity && h->d_overflow) return DHR_OK;
    if (h->d_overflow) cudaFree(h->d_overflow);
    h->d_overflow = nullptr; h->overflow_capacity = 0;
    DHR_CUDA(cudaMalloc(&h->d_overflow, (n_queries + 1) * sizeof(uint32_t)));
    h->overflow_capacity = n_queries;
    return DHR_OK;
}

// chunk boundaries (row indices).  growth <= 1 selects the overflow-proof uniform schedule.
// A chunk [a, b) scanned with tau = k-th best of the first a rows admits ~k*(b/a - 1) rows of a random-order corpus; the
// growth bound keeps that near half of the cap - k free slots.  Among the growths that give the minimal number of chunks
// (= selects per batch) the smallest is used, and the chunks after the first are whole multiples of `quantum` rows (the
// tile path's sub-launch size) so that only the last launch of a batch is ragged.
static std::vector<long long> chunk_schedule(long long n_rows, int k, int cap, bool safe, long long align = 1, long long quantum = 0) {
    std::vector<long long> b;
    b.push_back(0);
    const long long first = std::min<long long>(n_rows, ((long long)cap - k) / align * align);
    if (n_rows == 0) return b;
    b.push_back(first);
    if (safe) {
        while (b.back() < n_rows) b.push_back(std::min<long long>(n_rows, b.back() + (cap - k) / align * align));
        return b;
    }
    double gmax = 1.0 + 0.55 * (double)(cap - k) / (double)k;
    gmax = std::min(9.5, std::max(1.25, gmax));
    double growth = gmax;
    if (first > 0 && n_rows > first) {
        const double ratio = (double)n_rows / (double)first;
        const double n = ceil(log(ratio) / log(gmax) - 1e-9);
        growth = std::min(gmax, std::max(1.25, pow(ratio, 1.0 / std::max(1.0, n)) * 1.0001));
    }
    while (b.back() < n_rows) {
        long long next = (long long)ceil((double)b.back() * growth);
        if (next >= n_rows || (double)(n_rows - next) < 0.04 * (double)next) { b.push_back(n_rows); break; }   // no sliver chunk behind an aligned-down boundary
        next = next / align * align;
        if (quantum > 0 && next - b.back() >= quantum) {                 // whole sub-launches: nearest multiple, never beyond gmax
            const long long len = next - b.back();
            long long q = (len + quantum / 2) / quantum * quantum;
            while (q > quantum && (double)(b.back() + q) > (double)b.back() * gmax * 1.02) q -= quantum;
            next = b.back() + q;
        }
        if (next <= b.back()) next = b.back() + align;
        b.push_back(std::min(n_rows, next));
    }
    return b;
}

struct QuerySet {
    bool f32;                       // use fp32 query arrays
    const uint8_t* lex; const uint8_t* code; const uint8_t* dns;
    size_t lex_stride, code_stride, dns_stride;   // bytes per query
};

static QuerySet query_set(const dhr_index* h, bool f32) {
    const Geometry& g = h->g;
    QuerySet q;
    q.f32 = f32;
    q.lex = (const uint8_t*)(f32 ? h->q_lex32 : h->q_lex16);
    q.dns = (const uint8_t*)(f32 ? h->q_dns32 : h->q_dns16);
    q.code = (const uint8_t*)h->q_code;
    q.lex_stride = (size_t)g.D_pad * (f32 ? 4 : 2);
    q.dns_stride = (size_t)g.C_pad * (f32 ? 4 : 2);
    q.code_stride = (size_t)g.S_pad * g.code_bytes;
    return q;
}

// scan + select for queries [base, base+nq) of the prepared query set
static int run_batch(dhr_index* h, const QuerySet& qs, int base, int nq, int k, bool masked, bool safe, int qb,
                     SelectOut so, cudaStream_t st) {
    const Geometry& g = h->g;
    TopkState t = without_segments(h->topk);
    so.base = base;
    const std::vector<long long> bounds = chunk_schedule(h->n_rows, k, kCandCap, safe);
    ScanArgs a{};
    a.lexv = h->lexv; a.lexi = h->lexi; a.dns = h->dns;
    a.S_pad = g.S_pad; a.D_pad = g.D_pad; a.C_pad = g.C_pad; a.n_units = g.n_units; a.n_chunks = g.n_chunks;
    a.q_lex = qs.lex + (size_t)base * qs.lex_stride;
    a.q_code = qs.code + (size_t)base * qs.code_stride;
    a.q_dns = qs.dns + (size_t)base * qs.dns_stride;
    a.n_queries = nq;
    a.n_groups = (nq + qb - 1) / qb;
    a.masked = masked ? 1 : 0;
    a.tau = t.tau; a.cnt = t.cnt; a.cand_score = t.cand_score; a.cand_row = t.cand_row; a.cap = kCandCap;
    a.rows_per_cta = 128;
    a.tile_rows = 16;
    int variant = h->opt_scan_variant;
    if (variant == 1) {
        int stages = 4;
        while (stages > 1 && scan_tma_smem_bytes(g, qb, qs.f32, a.tile_rows, stages) > 200 * 1024) --stages;
        if (scan_tma_smem_bytes(g, qb, qs.f32, a.tile_rows, stages) > 200 * 1024 || stages < 2) variant = 0;   // rows too wide
        a.n_stages = stages;
    }
    const size_t direct_smem = (size_t)qb * (qs.lex_stride + qs.code_stride + qs.dns_stride);
    if (variant == 0 && direct_smem > 200 * 1024) return DHR_ERR_UNSUPPORTED;
    const size_t n_chunks = bounds.size() - 1;
    for (size_t c = 0; c < n_chunks; ++c) {
        a.row_begin = bounds[c];
        a.row_end = bounds[c + 1];
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        if (h->opt_profile) { e0 = h->events.get(); e1 = h->events.get(); e2 = h->events.get(); cudaEventRecord(e0, st); }
        DHR_TRY(launch_scan(h, a, qb, qs.f32, variant, st));
        if (h->opt_profile) cudaEventRecord(e1, st);
        const bool final_pass = (c + 1 == n_chunks);
        DHR_TRY(launch_select(t, nq, k, kCandCap, final_pass, so, st));
        if (h->opt_profile) cudaEventRecord(e2, st);
        h->stats.n_scan_launches++;
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches += 2;
        h->stats.corpus_passes += (double)(a.row_end - a.row_begin) * a.n_groups / (double)std::max<int64_t>(1, h->n_rows);
        h->stats.alg_bytes += (double)(a.row_end - a.row_begin) * a.n_groups * (double)g.row_bytes();
    }
    if (n_chunks == 0) {   // empty index: all padding
        DHR_TRY(launch_select(t, nq, k, kCandCap, true, so, st));
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches++;
    }
    h->stats.scan_variant = variant;
    return DHR_OK;
}

// dense-only index on the tensor-core tile kernel (K2): same chunk schedule, 128-row aligned
static int run_batch_dense_tile(dhr_index* h, const QuerySet& qs, int base, int nq, int k, SelectOut so, cudaStream_t st, int L) {
    DHR_TRY(ensure_segments(h, L));
    TopkState t = L ? h->topk1 : h->topk;
    so.base = base;
    const std::vector<long long> bounds = chunk_schedule(h->n_rows, k, kCandCap, false, 128);
    const size_t n_chunks = bounds.size() - 1;
    const void* q16 = qs.dns + (size_t)base * qs.dns_stride;
    for (size_t c = 0; c < n_chunks; ++c) {
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        if (h->opt_profile) { e0 = h->events.get(); e1 = h->events.get(); e2 = h->events.get(); cudaEventRecord(e0, st); }
        if (t.seg_cnt) DHR_CUDA(cudaMemsetAsync(t.seg_cnt, 0, (size_t)kMaxInflight * kSegCount * sizeof(uint32_t), st));
        DHR_TRY(launch_dense_tile(h, q16, nq, 0, bounds[c], bounds[c + 1], 0, nullptr, 0, t, kCandCap, st));
        if (h->opt_profile) cudaEventRecord(e1, st);
        DHR_TRY(launch_select(t, nq, k, kCandCap, c + 1 == n_chunks, so, st));
        if (h->opt_profile) cudaEventRecord(e2, st);
        h->stats.n_scan_launches++;
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches += 2;
        h->stats.corpus_passes += (double)(bounds[c + 1] - bounds[c]) * ((nq + 127) / 128) / (double)std::max<int64_t>(1, h->n_rows);   // K2 (TS): one pass per 128 queries
        h->stats.dense_flops += 2.0 * (double)(bounds[c + 1] - bounds[c]) * (double)nq * (double)h->g.C;
        h->stats.alg_bytes += (double)(bounds[c + 1] - bounds[c]) * ((nq + 127) / 128) * (double)h->g.C_pad * 2.0;
    }
    if (n_chunks == 0) {
        DHR_TRY(launch_select(t, nq, k, kCandCap, true, so, st));
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches++;
    }
    h->stats.scan_variant = 2;
    return DHR_OK;
}

// hybrid / lexical index on the tile kernels: per sub-chunk of rows K2 writes the dense scores of the
// in-flight queries to an L2-resident scratch, K1t adds the lexical part and filters.
constexpr long long kTileSubRows = 37888;        // 74 K1t tiles of 512 rows x 4 query tiles = 2 work items per CTA on 148 SMs

static int ensure_tile_workspace(dhr_index* h, const LexTileGeom& t, int n_queries) {
    const size_t n_qtiles = (size_t)(n_queries + kLexTileQueries - 1) / kLexTileQueries + 2;
    const size_t qb_need = n_qtiles * t.n_chunks * (size_t)t.qblock_stride;
    if (qb_need > h->qblocks_bytes) {
        if (h->qblocks) cudaFree(h->qblocks);
        h->qblocks = nullptr; h->qblocks_bytes = 0;
        DHR_CUDA(cudaMalloc(&h->qblocks, qb_need));
        h->qblocks_bytes = qb_need;
    }
    const size_t nb_need = n_qtiles * t.n_chunks * sizeof(uint32_t);
    if (nb_need > h->qblock_bytes_cap) {
        if (h->qblock_bytes) cudaFree(h->qblock_bytes);
        h->qblock_bytes = nullptr; h->qblock_bytes_cap = 0;
        DHR_CUDA(cudaMalloc(&h->qblock_bytes, nb_need));
        h->qblock_bytes_cap = nb_need;
    }
    return DHR_OK;
}

// streams, events and (hybrid index) the two scratch buffers of a batch lane
static int ensure_lane(dhr_index* h, int L, bool need_scratch) {
    dhr_index::TileLane& ln = h->lane[L];
    const size_t sc_need = need_scratch ? (size_t)2 * kMaxInflight * kTileSubRows * sizeof(float) : 0;
    if (sc_need > ln.scratch_bytes) {
        if (ln.scratch) cudaFree(ln.scratch);
        ln.scratch = nullptr; ln.scratch_bytes = 0;
        DHR_CUDA(cudaMalloc(&ln.scratch, sc_need));
        ln.scratch_bytes = sc_need;
    }
    if (!ln.aux) {
        // option stream_priority: the streams that carry K1t launches (and the selects) outrank K2's stream, so that the block
        // scheduler places K1t's big CTAs first and K2's fill what is left of an SM
        int least = 0, greatest = 0;
        if (h->opt_stream_priority) DHR_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        if (L == 1) DHR_CUDA(cudaStreamCreateWithPriority(&ln.main, cudaStreamNonBlocking, greatest));
        DHR_CUDA(cudaStreamCreateWithPriority(&ln.aux, cudaStreamNonBlocking, least));
        DHR_CUDA(cudaStreamCreateWithPriority(&ln.aux2, cudaStreamNonBlocking, greatest));
        DHR_CUDA(cudaEventCreateWithFlags(&ln.ev_join, cudaEventDisableTiming));
        DHR_CUDA(cudaEventCreateWithFlags(&ln.ev_sel, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            DHR_CUDA(cudaEventCreateWithFlags(&ln.ev_k2_done[i], cudaEventDisableTiming));
            DHR_CUDA(cudaEventCreateWithFlags(&ln.ev_k1_done[i], cudaEventDisableTiming));
        }
        DHR_CUDA(cudaEventCreateWithFlags(&ln.ev_fork, cudaEventDisableTiming));
        DHR_CUDA(cudaEventCreateWithFlags(&ln.ev_done, cudaEventDisableTiming));
    }
    if (!h->ev_lanes_fork) DHR_CUDA(cudaEventCreateWithFlags(&h->ev_lanes_fork, cudaEventDisableTiming));
    return DHR_OK;
}

static int run_batch_hybrid_tile(dhr_index* h, const LexTileGeom& lt, const QuerySet& qs, int base, int nq, int k,
                                 SelectOut so, cudaStream_t st, int L) {
    const Geometry& g = h->g;
    TopkState t = without_segments(L ? h->topk1 : h->topk);
    dhr_index::TileLane& ln = h->lane[L];
    so.base = base;
    const std::vector<long long> bounds = chunk_schedule(h->n_rows, k, kCandCap, false, kLexTileRows, kTileSubRows);
    const size_t n_chunks = bounds.size() - 1;
    const void* q16 = qs.dns + (size_t)base * qs.dns_stride;
    const size_t qt0 = (size_t)base / kLexTileQueries;
    const uint8_t* qblocks = h->qblocks + qt0 * lt.n_chunks * (size_t)lt.qblock_stride;
    const uint32_t* qbytes = h->qblock_bytes + qt0 * lt.n_chunks;
    // Stream plan.  K2 runs on an auxiliary stream one sub-chunk ahead of K1t (two scratch buffers); K2 (mode 1) reads only
    // the prepared queries, so it does not depend on the selects between chunks.  The K1t launches of one chunk are
    // independent of each other (they append to the candidate lists with atomics and read a tau that only the select
    // between chunks changes), so they alternate between the caller's stream and a second auxiliary stream: the CTAs of
    // the next launch fill the SMs that the tail of the previous one leaves idle.  The select of a chunk joins both.
    const bool dense = g.C_pad > 0;
    const bool overlap = h->opt_overlap != 0;
    const bool k2_overlap = dense && overlap;
    const size_t sc_half = (size_t)kMaxInflight * kTileSubRows;
    std::vector<std::pair<long long, long long>> subs;
    std::vector<size_t> sub_chunk;
    for (size_t c = 0; c < n_chunks; ++c)
        for (long long r0 = bounds[c]; r0 < bounds[c + 1]; r0 += kTileSubRows) {
            subs.emplace_back(r0, std::min(bounds[c + 1], r0 + kTileSubRows));
            sub_chunk.push_back(c);
        }
    cudaStream_t k2s = k2_overlap ? ln.aux : st;
    cudaStream_t k1s[2] = {st, overlap ? ln.aux2 : st};
    if (overlap) {
        DHR_CUDA(cudaEventRecord(ln.ev_fork, st));                   // queries prepared, slots initialised, previous batch done
        if (k2_overlap) DHR_CUDA(cudaStreamWaitEvent(k2s, ln.ev_fork, 0));
        DHR_CUDA(cudaStreamWaitEvent(k1s[1], ln.ev_fork, 0));
    }
    auto launch_k2 = [&](size_t i) -> int {
        const int b = (int)(i & 1);
        if (k2_overlap && i >= 2) DHR_CUDA(cudaStreamWaitEvent(k2s, ln.ev_k1_done[b], 0));     // K1t(i-2) has read this buffer
        DHR_TRY(launch_dense_tile(h, q16, nq, subs[i].first, subs[i].first, subs[i].second, 1, ln.scratch + (k2_overlap ? b * sc_half : 0),
                                  kMaxInflight, t, kCandCap, k2s));
        if (k2_overlap) DHR_CUDA(cudaEventRecord(ln.ev_k2_done[b], k2s));
        h->stats.n_kernel_launches++;
        return DHR_OK;
    };
    if (k2_overlap && !subs.empty()) DHR_TRY(launch_k2(0));
    size_t si = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        if (h->opt_profile) { e0 = h->events.get(); e1 = h->events.get(); e2 = h->events.get(); cudaEventRecord(e0, st); }
        bool used_aux2 = false;
        for (; si < subs.size() && sub_chunk[si] == c; ++si) {
            const int b = (int)(si & 1);
            cudaStream_t ks = k1s[b];
            const long long r0 = subs[si].first, r1 = subs[si].second;
            if (dense) {
                if (k2_overlap) {
                    if (si + 1 < subs.size()) DHR_TRY(launch_k2(si + 1));
                    DHR_CUDA(cudaStreamWaitEvent(ks, ln.ev_k2_done[b], 0));
                } else {
                    DHR_TRY(launch_k2(si));
                    if (ks != st) { DHR_CUDA(cudaEventRecord(ln.ev_k2_done[b], st)); DHR_CUDA(cudaStreamWaitEvent(ks, ln.ev_k2_done[b], 0)); }
                }
            }
            if (h->lexp)
                DHR_TRY(launch_lex_post(h, lt, qblocks, qbytes, nq, r0, r1, dense ? ln.scratch + (k2_overlap ? b * sc_half : 0) : nullptr, kMaxInflight,
                                        r0, t, kCandCap, ks));
            else
                DHR_TRY(launch_lex_tile(h, lt, qblocks, qbytes, nq, r0, r1, dense ? ln.scratch + (k2_overlap ? b * sc_half : 0) : nullptr, kMaxInflight,
                                        r0, t, kCandCap, ks));
            if (overlap) DHR_CUDA(cudaEventRecord(ln.ev_k1_done[b], ks));
            if (ks != st) used_aux2 = true;
            h->stats.n_kernel_launches++;
            h->stats.n_scan_launches++;
        }
        if (used_aux2) {                                             // the select needs every K1t launch of the chunk
            DHR_CUDA(cudaEventRecord(ln.ev_join, k1s[1]));
            DHR_CUDA(cudaStreamWaitEvent(st, ln.ev_join, 0));
        }
        if (h->opt_profile) cudaEventRecord(e1, st);
        DHR_TRY(launch_select(t, nq, k, kCandCap, c + 1 == n_chunks, so, st));
        if (h->opt_profile) cudaEventRecord(e2, st);
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches++;
        h->stats.corpus_passes += (double)(bounds[c + 1] - bounds[c]) * ((nq + kLexTileQueries - 1) / kLexTileQueries) /
                                  (double)std::max<int64_t>(1, h->n_rows);
        h->stats.dense_flops += 2.0 * (double)(bounds[c + 1] - bounds[c]) * (double)nq * (double)g.C;
        h->stats.alg_bytes += (double)(bounds[c + 1] - bounds[c]) *
                              (((nq + kLexTileQueries - 1) / kLexTileQueries) * ((double)g.D_pad * 2.0 + (double)g.S_pad * lt.tcode_bytes) +
                               ((nq + 127) / 128) * (double)g.C_pad * 2.0);
        if (overlap && c + 1 < n_chunks) {                           // the next chunk's K1t launches read the new tau
            DHR_CUDA(cudaEventRecord(ln.ev_sel, st));
            DHR_CUDA(cudaStreamWaitEvent(k1s[1], ln.ev_sel, 0));
        }
    }
    if (n_chunks == 0) {
        DHR_TRY(launch_select(t, nq, k, kCandCap, true, so, st));
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches++;
    }
    h->stats.scan_variant = 3;
    h->stats.lex_layout = h->lexp ? 1.0 : 0.0;
    return DHR_OK;
}

static int stage_queries_to_device(dhr_index* h, int n, int val_dtype, const void* q_vals, int64_t vstride, int idx_dtype,
                                   const void* q_idx, int64_t istride, const void** d_vals, int64_t* d_vstride,
                                   const void** d_idx, int64_t* d_istride, cudaStream_t st) {
    const Geometry& g = h->g;
    const int W = g.S * g.G + g.C;
    const size_t vsz = val_dtype == DHR_VAL_F16 ? 2 : 4;
    if (is_device_pointer(q_vals)) { *d_vals = q_vals; *d_vstride = vstride; }
    else {
        DHR_TRY(ensure_device_buffer(&h->stage_a, &h->stage_a_bytes, (size_t)n * W * vsz));
        DHR_CUDA(cudaMemcpy2DAsync(h->stage_a, (size_t)W * vsz, q_vals, (size_t)vstride * vsz, (size_t)W * vsz, (size_t)n,
                                   cudaMemcpyHostToDevice, st));
        *d_vals = h->stage_a; *d_vstride = W;
    }
    *d_idx = nullptr; *d_istride = 0;
    if (g.S > 0 && q_idx) {
        size_t isz = 1;
        switch (idx_dtype) { case DHR_IDX_I16: case DHR_IDX_U16: isz = 2; break; case DHR_IDX_I32: isz = 4; break; case DHR_IDX_I64: isz = 8; break; default: break; }
        if (is_device_pointer(q_idx)) { *d_idx = q_idx; *d_istride = istride; }
        else {
            DHR_TRY(ensure_device_buffer(&h->stage_b, &h->stage_b_bytes, (size_t)n * g.S * isz));
            DHR_CUDA(cudaMemcpy2DAsync(h->stage_b, (size_t)g.S * isz, q_idx, (size_t)istride * isz, (size_t)g.S * isz, (size_t)n,
                                       cudaMemcpyHostToDevice, st));
            *d_idx = h->stage_b; *d_istride = g.S;
        }
    }
    return DHR_OK;
}

static int validate_query_args(const dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t vstride,
                               int q_idx_dtype, const void* q_idx, int64_t istride, int k, bool masked) {
    if (!h) return DHR_ERR_INVALID;
    if (!h->finalized) return DHR_ERR_STATE;
    const Geometry& g = h->g;
    const int W = g.S * g.G + g.C;
    if (n_queries < 0 || k < 1) return DHR_ERR_INVALID;
    if (k > DHR_MAX_K) return DHR_ERR_UNSUPPORTED;
    if (n_queries == 0) return DHR_OK;
    if (!q_vals || vstride < W) return DHR_ERR_INVALID;
    if (q_val_dtype != DHR_VAL_F16 && q_val_dtype != DHR_VAL_F32) return DHR_ERR_INVALID;
    if (g.S > 0 && masked) {
        if (!q_idx || istride < g.S) return DHR_ERR_INVALID;
        if (q_idx_dtype < DHR_IDX_U8 || q_idx_dtype > DHR_IDX_I64) return DHR_ERR_INVALID;
    }
    return DHR_OK;
}

// Unmasked (--IP, gip_retrieval.py:139) first stage of an index with a lexical part: a plain [Q, W] x [W, N] inner product over
// ALL columns, i.e. a GEMM -> the tensor-core kernel K2, one column pass per <= 768 columns: the lexical value columns from the
// row-major array (TMA 2-D over a column window), then the dense block from its K-blocked copy; partial sums travel between
// the passes of a sub-chunk through the L2-resident scratch, the last pass adds them and applies the admission filter.
static int run_batch_unmasked_tile(dhr_index* h, const QuerySet& qs, int base, int nq, int k, SelectOut so, cudaStream_t st) {
    const Geometry& g = h->g;
    TopkState t = without_segments(h->topk);
    so.base = base;
    const std::vector<long long> bounds = chunk_schedule(h->n_rows, k, kCandCap, false, kDenseTileRows, kTileSubRows);
    const size_t n_chunks = bounds.size() - 1;
    struct Pass { const __half* blocked; const __half* rowmajor; int pitch, cols; const void* q; int q_pitch; };
    std::vector<Pass> passes;
    for (int c0 = 0; c0 < g.D_pad; c0 += kDensePassMaxCols)
        passes.push_back({nullptr, h->lexv + c0, g.D_pad, std::min(kDensePassMaxCols, g.D_pad - c0),
                          (const __half*)(qs.lex + (size_t)base * qs.lex_stride) + c0, g.D_pad});
    if (g.C_pad > 0)
        passes.push_back({h->dnst, h->dns, g.C_pad, g.C_pad, qs.dns + (size_t)base * qs.dns_stride, g.C_pad});
    const int np = (int)passes.size();
    for (size_t c = 0; c < n_chunks; ++c) {
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        if (h->opt_profile) { e0 = h->events.get(); e1 = h->events.get(); e2 = h->events.get(); cudaEventRecord(e0, st); }
        for (long long r0 = bounds[c]; r0 < bounds[c + 1]; r0 += kTileSubRows) {
            const long long r1 = std::min(bounds[c + 1], r0 + kTileSubRows);
            for (int p = 0; p < np; ++p) {
                const int mode = np == 1 ? 0 : (p == 0 ? 1 : (p == np - 1 ? 3 : 2));
                DHR_TRY(launch_dense_pass(h, passes[p].blocked, passes[p].rowmajor, passes[p].pitch, passes[p].cols, passes[p].q,
                                          passes[p].q_pitch, nq, r0, r0, r1, mode, h->lane[0].scratch, kMaxInflight, t, kCandCap, st));
                h->stats.n_kernel_launches++;
                h->stats.n_scan_launches++;
                h->stats.alg_bytes += (double)(r1 - r0) * ((nq + 127) / 128) * (double)passes[p].cols * 2.0;
            }
        }
        if (h->opt_profile) cudaEventRecord(e1, st);
        DHR_TRY(launch_select(t, nq, k, kCandCap, c + 1 == n_chunks, so, st));
        if (h->opt_profile) cudaEventRecord(e2, st);
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches++;
        h->stats.corpus_passes += (double)(bounds[c + 1] - bounds[c]) * ((nq + 127) / 128) / (double)std::max<int64_t>(1, h->n_rows);
        h->stats.dense_flops += 2.0 * (double)(bounds[c + 1] - bounds[c]) * (double)nq * (double)(g.S * g.G + g.C);
    }
    if (n_chunks == 0) {
        DHR_TRY(launch_select(t, nq, k, kCandCap, true, so, st));
        h->stats.n_select_launches++;
        h->stats.n_kernel_launches++;
    }
    h->stats.scan_variant = 4;
    return DHR_OK;
}

// ---- the search proper: shared by dhr_search (synchronous, host or device outputs) and dhr_search_keys (stream-ordered,
// device-resident packed keys, no host synchronisation on the common path) -------------------------------------------------
struct SearchRequest {
    int n_queries, q_val_dtype; const void* q_vals; int64_t vstride;
    int q_idx_dtype; const void* q_idx; int64_t istride;
    float lamda; int k; bool masked;
};

// re-run the queries whose candidate buffer overflowed (adversarial row order) with the overflow-proof schedule
static int rerun_overflowed(dhr_index* h, const std::vector<uint32_t>& flags, int n_queries, int k, bool masked, bool f32,
                            SelectOut so, cudaStream_t st, int* n_rerun) {
    const QuerySet qs = query_set(h, f32);
    int n = 0;
    so.overflow = nullptr;
    for (int q = 0; q < n_queries; ++q) {
        if (!flags[(size_t)q]) continue;
        ++n;
        h->stats.n_fallback_queries++;
        DHR_TRY(ensure_rowmajor(h));
        DHR_TRY(run_batch(h, qs, q, 1, k, masked, true, 1, so, st));
    }
    if (n_rerun) *n_rerun = n;
    return DHR_OK;
}

// Enqueue the whole search on `st`.  `so` holds DEVICE output pointers.  Returns with work in flight; the only host
// synchronisation is the read of the "queries need fp32" flag, which is skipped when the queries are fp16 and lamda == 1
// (always exactly representable: the reference's on-disk format).
static int enqueue_search(dhr_index* h, const SearchRequest& r, SelectOut so, bool record_batches, bool* used_f32, cudaStream_t st) {
    const Geometry& g = h->g;
    const int n_queries = r.n_queries, k = r.k;
    h->stats = dhr_stats{};
    h->stats.n_queries = n_queries;
    h->stats.bytes_per_pass = (double)h->n_rows * (double)g.row_bytes();
    h->events.reset();

    DHR_TRY(ensure_query_workspace(h, n_queries));
    DHR_TRY(ensure_topk_state(h));
    DHR_TRY(ensure_overflow_flags(h, (size_t)n_queries));
    const bool exact_f16 = r.q_val_dtype == DHR_VAL_F16 && r.lamda == 1.0f;
    if (!exact_f16) DHR_CUDA(cudaMemsetAsync(h->d_flags + 2, 0, sizeof(int), st));
    DHR_CUDA(cudaMemsetAsync(h->d_overflow, 0, (size_t)n_queries * sizeof(uint32_t), st));

    if (h->opt_profile) { h->ev_begin = h->events.get(); h->ev_end = h->events.get(); cudaEventRecord(h->ev_begin, st); }

    const void* d_vals; const void* d_idx; int64_t d_vs, d_is;
    DHR_TRY(stage_queries_to_device(h, n_queries, r.q_val_dtype, r.q_vals, r.vstride, r.q_idx_dtype, r.masked ? r.q_idx : nullptr,
                                    r.istride, &d_vals, &d_vs, &d_idx, &d_is, st));
    DHR_TRY(launch_prep_queries(h, n_queries, r.q_val_dtype, d_vals, d_vs, r.q_idx_dtype, d_idx, d_is, r.lamda, st));
    int need_f32 = 0;
    if (!exact_f16) {
        DHR_CUDA(cudaMemcpyAsync(&need_f32, h->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
        DHR_CUDA(cudaStreamSynchronize(st));
    }
    *used_f32 = need_f32 != 0;
    launch_init_slots(h->topk, st);
    DHR_CUDA(cudaGetLastError());
    h->stats.n_kernel_launches++;

    so.overflow = h->d_overflow;
    so.row_offset = h->row_offset;
    const QuerySet qs = query_set(h, need_f32 != 0);
    int qb = h->opt_query_block;
    int groups = h->opt_query_groups;
    if (qb * groups > kMaxScanInflight) groups = kMaxScanInflight / qb;
    int slots = qb * groups;
    // tensor-core tile path: dense-only index, or the unmasked (--IP) first stage of any index through K2 over the
    // lexical + dense columns is not available (K2 reads the dense block only) -> dense-only; queries exactly fp16
    const bool tile_dense = h->opt_tile_mode && g.S == 0 && !qs.f32 && dense_tile_supported(g, nullptr);
    if (tile_dense) { slots = kMaxInflight; qb = 64; groups = kMaxInflight / 64; }
    const int rt = std::max(1, h->max_code + 1);
    const bool post = h->lexp && lex_post_supported(g, rt);                // postings layout -> K1p, tiled layout -> K1t
    const bool tile_hybrid = h->opt_tile_mode && g.S > 0 && r.masked && !qs.f32 && (post || (h->lext && lex_tile_supported(g, rt))) &&
                             (g.C_pad == 0 || dense_tile_supported(g, nullptr));
    LexTileGeom lt{};
    if (tile_hybrid) {
        lt = post ? lex_post_geom(g, rt) : lex_tile_geom(g, rt);
        // K2 beside K1t on the same SM (dense_tile.cu, lite form): K1t gives up ring stages until two 8 KiB K2 stages (+ 4 KiB for K2's
        // alignment slack, static and reserved shared memory) fit into what its CTA leaves of the SM
        if (!post && h->opt_lex_stages >= 2 && h->opt_lex_stages < lt.n_stages) lt.n_stages = h->opt_lex_stages;
        h->lite_stages = 0;
        if (!post && g.C_pad > 0 && h->opt_dense_lite && h->opt_overlap && h->opt_dense_variant >= 1 && dense_tile_ts_supported(g)) {
            auto room = [&](int stages) { return h->smem_per_sm - (long long)lex_tile_cta_footprint(lt, stages) - 4096; };
            int stages = lt.n_stages;
            while (stages > 2 && room(stages) < 2 * 8192) --stages;
            if (room(stages) >= 2 * 8192) { lt.n_stages = stages; h->lite_stages = (int)std::min<long long>(8, room(stages) / 8192); }
        }
        DHR_TRY(ensure_tile_workspace(h, lt, n_queries));
        DHR_TRY(ensure_lane(h, 0, g.C_pad > 0));
        DHR_TRY(launch_lex_tile_prep(h, lt, h->q_lex16, h->q_code, n_queries, h->qblocks, h->qblock_bytes, st));
        h->stats.n_prep_launches++; h->stats.n_kernel_launches++;
        slots = kMaxInflight; qb = kLexTileQueries; groups = kMaxInflight / kLexTileQueries;
    }
    // unmasked search of an index with a lexical part (--IP first stage): all columns through K2
    const bool tile_unmasked = h->opt_tile_mode && g.S > 0 && !r.masked && !qs.f32 && h->opt_dense_variant >= 1 &&
                               (g.C_pad == 0 || dense_tile_ts_supported(g));
    if (tile_unmasked) {
        DHR_TRY(ensure_rowmajor(h));                                      // the lexical columns are read from the row-major array
        DHR_TRY(ensure_lane(h, 0, true));
        slots = kMaxInflight; qb = 128; groups = kMaxInflight / 128;
    }
    if (!tile_dense && !tile_hybrid) DHR_TRY(ensure_rowmajor(h));          // the row scan K1 reads the row-major arrays
    else if (g.C_pad > 0 && (!h->dnst || h->opt_dense_variant < 1 || !dense_tile_ts_supported(g)))
        DHR_TRY(ensure_rowmajor(h));                                      // K2's SS variant streams the row-major dense block
    h->stats.query_block = qb;
    h->stats.query_groups = groups;
    h->batch_size = slots;
    // two batch lanes (tile paths): odd batches run on lane 1's streams with lane 1's selection state and scratch
    const bool two_lanes = (tile_hybrid || tile_dense) && h->opt_lanes >= 2 && n_queries > slots;
    if (two_lanes) {
        DHR_TRY(ensure_lane(h, 0, tile_hybrid && g.C_pad > 0));
        DHR_TRY(ensure_lane(h, 1, tile_hybrid && g.C_pad > 0));
        DHR_TRY(ensure_topk_state(h, 1));
        launch_init_slots(h->topk1, st);
        DHR_CUDA(cudaGetLastError());
        h->stats.n_kernel_launches++;
        DHR_CUDA(cudaEventRecord(h->ev_lanes_fork, st));                  // queries prepared, both lanes' slots initialised
        DHR_CUDA(cudaStreamWaitEvent(h->lane[1].main, h->ev_lanes_fork, 0));
    }
    int b = 0;
    for (int base = 0; base < n_queries; base += slots, ++b) {
        const int nq = std::min(slots, n_queries - base);
        const int L = two_lanes ? (b & 1) : 0;
        cudaStream_t ms = L ? h->lane[1].main : st;
        if (tile_dense) DHR_TRY(run_batch_dense_tile(h, qs, base, nq, k, so, ms, L));
        else if (tile_unmasked) DHR_TRY(run_batch_unmasked_tile(h, qs, base, nq, k, so, st));
        else if (tile_hybrid) DHR_TRY(run_batch_hybrid_tile(h, lt, qs, base, nq, k, so, ms, L));
        else DHR_TRY(run_batch(h, qs, base, nq, k, r.masked, false, qb, so, st));
        if (record_batches) {
            while ((int)h->batch_events.size() <= b) {
                cudaEvent_t e;
                DHR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->batch_events.push_back(e);
            }
            DHR_CUDA(cudaEventRecord(h->batch_events[(size_t)b], ms));
        }
    }
    if (two_lanes) {                                                       // the caller's stream continues after both lanes
        DHR_CUDA(cudaEventRecord(h->lane[1].ev_done, h->lane[1].main));
        DHR_CUDA(cudaStreamWaitEvent(st, h->lane[1].ev_done, 0));
    }
    h->n_batches = b;
    if (h->opt_profile) cudaEventRecord(h->ev_end, st);
    return DHR_OK;
}

static void collect_profile(dhr_index* h) {
    if (!h->opt_profile || !h->ev_begin) return;
    // events were taken in triples (scan begin, scan end, select end) after the two bracket events
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end);
    h->stats.total_ms = ms;
    for (size_t i = 2; i + 2 < h->events.used; i += 3) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, h->events.ev[i], h->events.ev[i + 1]);
        cudaEventElapsedTime(&b, h->events.ev[i + 1], h->events.ev[i + 2]);
        h->stats.scan_ms += a;
        h->stats.select_ms += b;
    }
    h->ev_begin = h->ev_end = nullptr;
}

}  // namespace dhr

using namespace dhr;

extern "C" int dhr_search(dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t vstride, int q_idx_dtype,
                          const void* q_idx, int64_t istride, float lamda, int k, unsigned flags, float* out_scores,
                          int64_t* out_rows, int32_t* out_counts, void* stream) {
    const bool masked = !(flags & DHR_SEARCH_UNMASKED);
    DHR_TRY(validate_query_args(h, n_queries, q_val_dtype, q_vals, vstride, q_idx_dtype, q_idx, istride, k, masked));
    if (!out_scores || !out_rows) return DHR_ERR_INVALID;
    if (n_queries == 0) return DHR_OK;
    if (h->pending.active) return DHR_ERR_STATE;                          // a dhr_search_keys call awaits dhr_search_complete
    DHR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;

    const bool out_dev = is_device_pointer(out_scores) && is_device_pointer(out_rows) &&
                         (!out_counts || is_device_pointer(out_counts));
    SelectOut so{};
    so.scores = out_scores; so.rows = out_rows; so.counts = out_counts;
    if (!out_dev) {
        DHR_TRY(ensure_out_buffers(h, (size_t)n_queries, k));
        so.scores = h->d_out_scores; so.rows = h->d_out_rows; so.counts = h->d_out_counts;
    }
    const SearchRequest r{n_queries, q_val_dtype, q_vals, vstride, q_idx_dtype, q_idx, istride, lamda, k, masked};
    bool f32 = false;
    int status = enqueue_search(h, r, so, false, &f32, st);
    const bool enqueue_failed = status != DHR_OK;                        // its CUDA error text is already recorded
    // one synchronisation: results (host outputs) and the per-query overflow flags come back together
    std::vector<uint32_t> h_overflow((size_t)n_queries, 0u);
    auto copy_out = [&]() -> int {
        if (out_dev) return DHR_OK;
        const size_t n = (size_t)n_queries * k;
        DHR_CUDA(cudaMemcpyAsync(out_scores, so.scores, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        DHR_CUDA(cudaMemcpyAsync(out_rows, so.rows, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        if (out_counts) DHR_CUDA(cudaMemcpyAsync(out_counts, so.counts, (size_t)n_queries * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        return DHR_OK;
    };
    if (status == DHR_OK) status = copy_out();
    if (status == DHR_OK && cudaMemcpyAsync(h_overflow.data(), h->d_overflow, (size_t)n_queries * sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        status = DHR_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess && status == DHR_OK) status = DHR_ERR_CUDA;
    if (status == DHR_OK) {
        int n_rerun = 0;
        status = rerun_overflowed(h, h_overflow, n_queries, k, masked, f32, so, st, &n_rerun);
        if (status == DHR_OK && n_rerun > 0) {
            status = copy_out();
            if (cudaStreamSynchronize(st) != cudaSuccess && status == DHR_OK) status = DHR_ERR_CUDA;
        }
    }
    if (status == DHR_ERR_CUDA && !enqueue_failed) set_cuda_error(cudaGetLastError(), "dhr_search", __FILE__, __LINE__);
    if (status != DHR_OK) return status;
    collect_profile(h);
    return DHR_OK;
}

// ---- stream-ordered variant for the sharded search --------------------------------------------------------------------
extern "C" int dhr_search_keys(dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t vstride, int q_idx_dtype,
                               const void* q_idx, int64_t istride, float lamda, int k, unsigned flags, uint64_t* out_keys,
                               void* stream) {
    const bool masked = !(flags & DHR_SEARCH_UNMASKED);
    DHR_TRY(validate_query_args(h, n_queries, q_val_dtype, q_vals, vstride, q_idx_dtype, q_idx, istride, k, masked));
    if (!out_keys || !is_device_pointer(out_keys)) return DHR_ERR_INVALID;
    if (h->pending.active) return DHR_ERR_STATE;
    if ((uint64_t)h->row_offset + (uint64_t)h->n_rows >= 0xFFFFFFFFull) return DHR_ERR_UNSUPPORTED;   // global rows are packed in 32 bits
    if (n_queries == 0) return DHR_OK;
    DHR_CUDA(cudaSetDevice(h->device));
    SelectOut so{};
    so.keys = out_keys;
    const SearchRequest r{n_queries, q_val_dtype, q_vals, vstride, q_idx_dtype, q_idx, istride, lamda, k, masked};
    bool f32 = false;
    const int status = enqueue_search(h, r, so, true, &f32, (cudaStream_t)stream);
    if (status != DHR_OK) return status;
    h->pending.active = true; h->pending.n_queries = n_queries; h->pending.k = k; h->pending.masked = masked; h->pending.f32 = f32;
    h->pending.out_keys = out_keys;
    return DHR_OK;
}

extern "C" int dhr_search_batches(const dhr_index* h, int* batch_size, int* n_batches) {
    if (!h) return DHR_ERR_INVALID;
    if (batch_size) *batch_size = h->batch_size;
    if (n_batches) *n_batches = h->n_batches;
    return DHR_OK;
}

extern "C" int dhr_search_wait_batch(dhr_index* h, int batch, void* stream) {
    if (!h || !h->pending.active || batch < 0 || batch >= h->n_batches) return DHR_ERR_INVALID;
    DHR_CUDA(cudaSetDevice(h->device));
    DHR_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->batch_events[(size_t)batch], 0));
    return DHR_OK;
}

extern "C" int dhr_search_complete(dhr_index* h, int* n_rerun, void* stream) {
    if (!h) return DHR_ERR_INVALID;
    if (n_rerun) *n_rerun = 0;
    if (!h->pending.active) return DHR_ERR_STATE;
    DHR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n_queries = h->pending.n_queries;
    std::vector<uint32_t> h_overflow((size_t)n_queries, 0u);
    h->pending.active = false;
    DHR_CUDA(cudaMemcpyAsync(h_overflow.data(), h->d_overflow, (size_t)n_queries * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DHR_CUDA(cudaStreamSynchronize(st));
    SelectOut so{};
    so.keys = h->pending.out_keys; so.row_offset = h->row_offset;
    int n = 0;
    DHR_TRY(rerun_overflowed(h, h_overflow, n_queries, h->pending.k, h->pending.masked, h->pending.f32, so, st, &n));
    if (n > 0) DHR_CUDA(cudaStreamSynchronize(st));
    if (n_rerun) *n_rerun = n;
    collect_profile(h);
    return DHR_OK;
}

extern "C" int dhr_rerank(dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t vstride, int q_idx_dtype,
                          const void* q_idx, int64_t istride, float lamda, const int64_t* cand_rows, int n_cand, int k,
                          float* out_scores, int64_t* out_rows, int32_t* out_counts, void* stream) {
    DHR_TRY(validate_query_args(h, n_queries, q_val_dtype, q_vals, vstride, q_idx_dtype, q_idx, istride, k, true));
    if (!out_scores || !out_rows || !cand_rows || n_cand < 1) return DHR_ERR_INVALID;
    if (n_cand > kCandCap) return DHR_ERR_UNSUPPORTED;
    if (n_queries == 0) return DHR_OK;
    if (h->pending.active) return DHR_ERR_STATE;
    DHR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Geometry& g = h->g;
    h->stats = dhr_stats{};
    h->stats.n_queries = n_queries;
    DHR_TRY(ensure_query_workspace(h, n_queries));
    DHR_TRY(ensure_topk_state(h));
    DHR_TRY(ensure_rowmajor(h));
    DHR_CUDA(cudaMemsetAsync(h->d_flags + 2, 0, sizeof(int), st));
    const void* d_vals; const void* d_idx; int64_t d_vs, d_is;
    DHR_TRY(stage_queries_to_device(h, n_queries, q_val_dtype, q_vals, vstride, q_idx_dtype, q_idx, istride, &d_vals, &d_vs,
                                    &d_idx, &d_is, st));
    DHR_TRY(launch_prep_queries(h, n_queries, q_val_dtype, d_vals, d_vs, q_idx_dtype, d_idx, d_is, lamda, st));
    int need_f32 = 0;
    if (!(q_val_dtype == DHR_VAL_F16 && lamda == 1.0f)) {
        DHR_CUDA(cudaMemcpyAsync(&need_f32, h->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
        DHR_CUDA(cudaStreamSynchronize(st));
    }

    const bool out_dev = is_device_pointer(out_scores) && is_device_pointer(out_rows) &&
                         (!out_counts || is_device_pointer(out_counts));
    SelectOut so{};
    so.scores = out_scores; so.rows = out_rows; so.counts = out_counts; so.row_offset = h->row_offset;
    if (!out_dev) {
        DHR_TRY(ensure_out_buffers(h, (size_t)n_queries, k));
        so.scores = h->d_out_scores; so.rows = h->d_out_rows; so.counts = h->d_out_counts;
    }
    const long long* d_cand = (const long long*)cand_rows;
    if (!is_device_pointer(cand_rows)) {
        DHR_TRY(ensure_device_buffer(&h->stage_c, &h->stage_c_bytes, (size_t)n_queries * n_cand * sizeof(long long)));
        DHR_CUDA(cudaMemcpyAsync(h->stage_c, cand_rows, (size_t)n_queries * n_cand * sizeof(long long), cudaMemcpyHostToDevice, st));
        d_cand = (const long long*)h->stage_c;
    }
    const QuerySet qs = query_set(h, need_f32 != 0);
    int status = DHR_OK;
    launch_init_slots(h->topk, st);
    h->stats.n_kernel_launches++;
    for (int base = 0; base < n_queries && status == DHR_OK; base += kMaxInflight) {
        const int nq = std::min(kMaxInflight, n_queries - base);
        TopkState t = without_segments(h->topk);
        h->stats.n_kernel_launches += 2;
        ScanArgs a{};
        a.lexv = h->lexv; a.lexi = h->lexi; a.dns = h->dns;
        a.S_pad = g.S_pad; a.D_pad = g.D_pad; a.C_pad = g.C_pad; a.n_units = g.n_units; a.n_chunks = g.n_chunks;
        a.q_lex = qs.lex + (size_t)base * qs.lex_stride;
        a.q_code = qs.code + (size_t)base * qs.code_stride;
        a.q_dns = qs.dns + (size_t)base * qs.dns_stride;
        a.n_queries = nq; a.n_groups = nq; a.masked = 1;
        a.tau = t.tau; a.cnt = t.cnt; a.cand_score = t.cand_score; a.cand_row = t.cand_row; a.cap = kCandCap;
        status = launch_rerank(h, a, qs.f32, d_cand + (size_t)base * n_cand, n_cand, st);
        so.base = base;
        if (status == DHR_OK) status = launch_select(t, nq, k, kCandCap, true, so, st);
        h->stats.n_scan_launches++;
        h->stats.n_select_launches++;
    }
    if (status == DHR_OK && !out_dev) {
        const size_t n = (size_t)n_queries * k;
        if (cudaMemcpyAsync(out_scores, so.scores, n * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(out_rows, so.rows, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            (out_counts && cudaMemcpyAsync(out_counts, so.counts, (size_t)n_queries * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess))
            status = DHR_ERR_CUDA;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && status == DHR_OK) status = DHR_ERR_CUDA;
    if (status == DHR_ERR_CUDA) set_cuda_error(cudaGetLastError(), "dhr_rerank", __FILE__, __LINE__);
    return status;
}
