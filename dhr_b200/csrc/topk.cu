// topk.cu -- K3: per-query exact selection over the candidate list, and the multi-shard merge.
//
// Replaces torch.topk / torch.argsort of gip_retrieval.py:75,123 and the per-query
// argsort of retrieval/merge.result.py:39.  Order is total: (score desc, row asc), encoded in
// one 64-bit key so a single descending bitonic sort in shared memory yields the answer.
#include <algorithm>

#include "internal.h"

namespace dhr {

constexpr int kSelectThreads = 1024;

// descending bitonic sort of m (power of two) keys in shared memory
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int m, int tid, int nthreads) {
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (m >> 1); t += nthreads) {
                const int lo = ((t / stride) * (stride << 1)) + (t % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                const bool swap = desc ? (a < b) : (a > b);
                if (swap) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// One CTA per in-flight query slot.  Keeps the best k of the candidates appended so far in slots [0, k) (sorted),
// publishes the new strict admission threshold tau (the k-th best score: rows scanned later have larger row ids, so a
// later row with score == tau loses the tie) and, on the final pass, writes the result row.
//
// Selection before sorting: the n <= 16384 candidate keys stay in registers (16 per thread); an MSB-first radix select
// (8 passes of 8 bits; per-warp private histograms: candidate scores share their high bytes, so a CTA-wide
// histogram would serialise all warps on one bin) finds the k-th largest key -- keys are unique because
// they embed the row -- then only the k keys >= it are compacted into shared memory and bitonic-sorted.  A chunk
// admits ~7k rows next to the k kept ones, so this sorts 1024 keys instead of 8192-16384, and the small footprint
// (8 * pow2(k) bytes) lets two CTAs share an SM.
constexpr int kSelectKeysPerThread = kCandCap / kSelectThreads;
static_assert(kCandCap % kSelectThreads == 0, "candidate keys are distributed evenly over the select CTA");

__global__ void __launch_bounds__(kSelectThreads, 1)
topk_select_kernel(TopkState t, int k, int cap, int final_pass, long long row_offset, float* out_scores,
                   long long* out_rows, int* out_counts, int out_base) {
    extern __shared__ __align__(16) unsigned long long keys[];          // [m] selected keys, m = pow2 >= min(n, k)
    __shared__ uint32_t whist[kSelectThreads / 32][256];              // one histogram per warp
    __shared__ uint32_t hist[256];
    __shared__ uint32_t wtot[kSelectThreads / 32];
    __shared__ uint32_t sh_digit, sh_need;
    const int slot = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t raw = t.cnt[slot];
    const int n = (int)min(raw, (uint32_t)cap);
    if (raw > (uint32_t)cap && tid == 0) t.overflow[slot] = 1u;
    float* cs = t.cand_score + (size_t)slot * cap;
    int32_t* cr = t.cand_row + (size_t)slot * cap;
    unsigned long long r[kSelectKeysPerThread];
#pragma unroll
    for (int j = 0; j < kSelectKeysPerThread; ++j) {
        const int i = tid + j * kSelectThreads;
        r[j] = i < n ? make_key(cs[i], (uint32_t)cr[i]) : 0ull;
    }
    const int keep = min(n, k);
    unsigned long long kth = 0ull;                                       // n <= k: everything is kept
    if (n > k) {
        unsigned long long prefix = 0ull;
        uint32_t need = (uint32_t)k;                                     // rank (from the top) inside the current prefix group
        for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
            for (int b = 0; b < 8; ++b) whist[warp][lane * 8 + b] = 0u;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kSelectKeysPerThread; ++j) {
                const int i = tid + j * kSelectThreads;
                const bool active = i < n && (shift == 56 || (r[j] >> (shift + 8)) == prefix);
                const uint32_t digit = (uint32_t)(r[j] >> shift) & 0xFFu;
                if (active) atomicAdd(&whist[warp][digit], 1u);       // conflicts stay inside the warp (MATCH.ANY grouping measured slower)
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t sum = 0;
#pragma unroll 8
                for (int w = 0; w < kSelectThreads / 32; ++w) sum += whist[w][tid];
                hist[tid] = sum;
            }
            __syncthreads();
            if (warp == 0) {                                             // digit of the need-th key, counting from bin 255 down
                uint32_t c[8], sum = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; sum += c[b]; }
                uint32_t suf = sum;                                      // inclusive suffix sum over lanes
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t v = __shfl_down_sync(0xFFFFFFFFu, suf, off);
                    if (lane + off < 32) suf += v;
                }
                const uint32_t above = suf - sum;
                if (above < need && suf >= need) {
                    uint32_t rem = need - above;
                    int d = 0;
#pragma unroll
                    for (int b = 7; b >= 0; --b) {
                        if (rem != 0u) {
                            if (c[b] >= rem) { d = lane * 8 + b; sh_need = rem; rem = 0u; }
                            else rem -= c[b];
                        }
                    }
                    sh_digit = (uint32_t)d;
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (unsigned long long)sh_digit;
            need = sh_need;
        }
        kth = prefix;
    }
    // compaction of the kept keys, then a small sort
    int m = 2;
    while (m < keep) m <<= 1;
    for (int i = tid; i < m; i += kSelectThreads) keys[i] = 0ull;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < kSelectKeysPerThread; ++j) mine += (tid + j * kSelectThreads < n && r[j] >= kth) ? 1u : 0u;
    uint32_t incl = mine;                                                // block-wide exclusive scan (no atomics)
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t wv = wtot[lane];
        uint32_t wi = wv;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, off);
            if (lane >= off) wi += v;
        }
        wtot[lane] = wi - wv;
    }
    __syncthreads();
    uint32_t pos = wtot[warp] + incl - mine;
#pragma unroll
    for (int j = 0; j < kSelectKeysPerThread; ++j)
        if (tid + j * kSelectThreads < n && r[j] >= kth) keys[pos++] = r[j];
    bitonic_sort_desc(keys, m, tid, kSelectThreads);
    for (int i = tid; i < keep; i += kSelectThreads) {
        const unsigned long long key = keys[i];
        cs[i] = key_score(key);
        cr[i] = (int32_t)key_row(key);
    }
    if (tid == 0) {
        t.cnt[slot] = (uint32_t)keep;
        t.tau[slot] = (n >= k) ? key_score(keys[k - 1]) : -INFINITY;
    }
    if (final_pass) {
        float* os = out_scores + (size_t)(out_base + slot) * k;
        long long* orow = out_rows + (size_t)(out_base + slot) * k;
        for (int i = tid; i < k; i += kSelectThreads) {
            if (i < keep) {
                const unsigned long long key = keys[i];
                os[i] = key_score(key);
                orow[i] = (long long)key_row(key) + row_offset;
            } else {
                os[i] = -INFINITY;
                orow[i] = -1;
            }
        }
        if (tid == 0 && out_counts) out_counts[out_base + slot] = keep;
    }
}

int launch_select(const TopkState& t, int n_slots, int k, int cap, bool final_pass, int64_t row_offset,
                  float* out_scores, int64_t* out_rows, int32_t* out_counts, int out_base, cudaStream_t st) {
    if (n_slots <= 0) return DHR_OK;
    if (cap > kCandCap) return DHR_ERR_INVALID;
    int m = 2;
    while (m < std::min(k, cap)) m <<= 1;
    const size_t smem = (size_t)m * sizeof(unsigned long long);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        DHR_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    topk_select_kernel<<<n_slots, kSelectThreads, smem, st>>>(t, k, cap, final_pass ? 1 : 0, (long long)row_offset,
                                                              out_scores, (long long*)out_rows, out_counts, out_base);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

// ---------------------------------------------------------------------------------------------
// shard merge: [P, Q, k] -> [Q, k].  Rows are global ids (int64); they are ranked through their
// position in the gathered list only when scores tie, so the key uses the 64-bit row directly:
// two-key sort = sort by (score desc, row asc) with a 96-bit comparison done as key + payload.
// ---------------------------------------------------------------------------------------------
struct MergeItem { unsigned int s; unsigned int pad; long long row; };

__device__ __forceinline__ bool merge_before(const MergeItem& a, const MergeItem& b) {
    // true if a ranks strictly before b
    if (a.s != b.s) return a.s > b.s;
    return (unsigned long long)a.row < (unsigned long long)b.row;   // padding rows (-1) compare last
}

__global__ void __launch_bounds__(kSelectThreads, 1)
topk_merge_kernel(int P, int Q, int k, const float* scores, const long long* rows, float* out_scores, long long* out_rows) {
    extern __shared__ __align__(16) unsigned char raw_smem[];
    MergeItem* items = (MergeItem*)raw_smem;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n = P * k;
    int m = 2;
    while (m < n) m <<= 1;
    for (int i = tid; i < m; i += kSelectThreads) {
        MergeItem it;
        it.pad = 0;
        if (i < n) {
            const int p = i / k, j = i % k;
            const size_t src = ((size_t)p * Q + q) * k + j;
            const long long r = rows[src];
            it.row = r;
            it.s = r >= 0 ? float_to_ordered(scores[src] + 0.0f) : 0u;
        } else {
            it.row = -1;
            it.s = 0u;
        }
        items[i] = it;
    }
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (m >> 1); t += kSelectThreads) {
                const int lo = ((t / stride) * (stride << 1)) + (t % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const MergeItem a = items[lo], b = items[hi];
                const bool swap = desc ? merge_before(b, a) : merge_before(a, b);
                if (swap) { items[lo] = b; items[hi] = a; }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < k; i += kSelectThreads) {
        const MergeItem it = i < m ? items[i] : MergeItem{0u, 0u, -1};
        const bool valid = it.row >= 0;
        out_scores[(size_t)q * k + i] = valid ? ordered_to_float(it.s) : -INFINITY;
        out_rows[(size_t)q * k + i] = valid ? it.row : -1;
    }
}

}  // namespace dhr

using namespace dhr;

extern "C" int dhr_topk_merge(int device, int n_parts, int n_queries, int k, const float* scores, const int64_t* rows,
                              float* out_scores, int64_t* out_rows, void* stream) {
    if (n_parts <= 0 || n_queries < 0 || k <= 0 || !scores || !rows || !out_scores || !out_rows) return DHR_ERR_INVALID;
    if (n_queries == 0) return DHR_OK;
    long long n = (long long)n_parts * k;
    long long m = 2;
    while (m < n) m <<= 1;
    const size_t smem = (size_t)m * sizeof(MergeItem);
    if (smem > 200 * 1024) return DHR_ERR_UNSUPPORTED;
    DHR_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t in_elems = (size_t)n_parts * n_queries * k, out_elems = (size_t)n_queries * k;
    const bool in_dev = is_device_pointer(scores) && is_device_pointer(rows);
    const bool out_dev = is_device_pointer(out_scores) && is_device_pointer(out_rows);
    float* d_s = nullptr; long long* d_r = nullptr; float* d_os = nullptr; long long* d_or = nullptr;
    int status = DHR_OK;
    auto cleanup = [&]() {
        if (!in_dev) { cudaFree(d_s); cudaFree(d_r); }
        if (!out_dev) { cudaFree(d_os); cudaFree(d_or); }
    };
#define MERGE_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_cuda_error(_e, #expr, __FILE__, __LINE__); cleanup(); return DHR_ERR_CUDA; } } while (0)
    if (in_dev) { d_s = (float*)scores; d_r = (long long*)rows; }
    else {
        MERGE_CUDA(cudaMalloc(&d_s, in_elems * 4)); MERGE_CUDA(cudaMalloc(&d_r, in_elems * 8));
        MERGE_CUDA(cudaMemcpyAsync(d_s, scores, in_elems * 4, cudaMemcpyHostToDevice, st));
        MERGE_CUDA(cudaMemcpyAsync(d_r, rows, in_elems * 8, cudaMemcpyHostToDevice, st));
    }
    if (out_dev) { d_os = out_scores; d_or = (long long*)out_rows; }
    else { MERGE_CUDA(cudaMalloc(&d_os, out_elems * 4)); MERGE_CUDA(cudaMalloc(&d_or, out_elems * 8)); }
    MERGE_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<n_queries, kSelectThreads, smem, st>>>(n_parts, n_queries, k, d_s, d_r, d_os, d_or);
    MERGE_CUDA(cudaGetLastError());
    if (!out_dev) {
        MERGE_CUDA(cudaMemcpyAsync(out_scores, d_os, out_elems * 4, cudaMemcpyDeviceToHost, st));
        MERGE_CUDA(cudaMemcpyAsync(out_rows, d_or, out_elems * 8, cudaMemcpyDeviceToHost, st));
    }
    MERGE_CUDA(cudaStreamSynchronize(st));
#undef MERGE_CUDA
    cleanup();
    return status;
}
