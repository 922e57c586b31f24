// topk.cu -- K3: per-query exact selection over the candidate list, and the multi-shard merge.
//
// Replaces torch.topk / torch.argsort of gip_retrieval.py:75,123 and the per-query
// argsort of retrieval/merge.result.py:39.  Order is total: (score desc, row asc), encoded in
// one 64-bit key so a single descending bitonic sort in shared memory yields the answer.
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

// One CTA per in-flight query slot.  Sorts the candidates appended so far, keeps the best k in
// slots [0, k), publishes the new strict admission threshold tau (the k-th best score: rows
// scanned later have larger row ids, so a later row with score == tau loses the tie) and, on
// the final pass, writes the result row.
__global__ void __launch_bounds__(kSelectThreads, 1)
topk_select_kernel(TopkState t, int k, int cap, int final_pass, long long row_offset, float* out_scores,
                   long long* out_rows, int* out_counts, int out_base) {
    extern __shared__ __align__(16) unsigned long long keys[];
    const int slot = blockIdx.x;
    const int tid = threadIdx.x;
    const uint32_t raw = t.cnt[slot];
    const int n = (int)min(raw, (uint32_t)cap);
    if (raw > (uint32_t)cap && tid == 0) t.overflow[slot] = 1u;
    int m = 2;
    while (m < n) m <<= 1;
    float* cs = t.cand_score + (size_t)slot * cap;
    int32_t* cr = t.cand_row + (size_t)slot * cap;
    for (int i = tid; i < m; i += kSelectThreads)
        keys[i] = i < n ? make_key(cs[i], (uint32_t)cr[i]) : 0ull;
    bitonic_sort_desc(keys, m, tid, kSelectThreads);
    const int keep = min(n, k);
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
    static bool attr_set = false;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    if (!attr_set) {
        DHR_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
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
