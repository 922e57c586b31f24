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
topk_select_kernel(TopkState t, int k, int cap, int final_pass, const SelectOut o) {
    const long long row_offset = o.row_offset;
    float* out_scores = o.scores; long long* out_rows = (long long*)o.rows; int* out_counts = o.counts;
    unsigned long long* out_keys = (unsigned long long*)o.keys;
    const int out_base = o.base;
    extern __shared__ __align__(16) unsigned long long keys[];          // [m] selected keys, m = pow2 >= min(n, k)
    __shared__ uint32_t whist[kSelectThreads / 32][256];              // one histogram per warp
    __shared__ uint32_t hist[256];
    __shared__ uint32_t wtot[kSelectThreads / 32];
    __shared__ uint32_t sh_digit, sh_need, sh_bin;
    __shared__ unsigned long long wmin[kSelectThreads / 32];
    const int slot = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* cs = t.cand_score + (size_t)slot * cap;
    int32_t* cr = t.cand_row + (size_t)slot * cap;
    uint32_t raw = t.cnt[slot];
    bool seg_over = false;
    if (t.seg_cnt) {
        // gather the per-CTA segments of the tensor-core filter epilogue behind the kept candidates: exclusive scan of the segment
        // counts (warp 0), then one warp per segment copies its entries (coalesced) into the flat list
        __shared__ uint32_t seg_off[kSegCount + 1];
        __shared__ uint32_t seg_flag;
        if (tid == 0) seg_flag = 0u;
        __syncthreads();
        if (warp == 0) {
            uint32_t carry = 0;
            for (int r0 = 0; r0 < kSegCount; r0 += 32) {
                const uint32_t c = t.seg_cnt[(size_t)slot * kSegCount + r0 + lane];
                if (c > (uint32_t)kSegCap) seg_flag = 1u;
                const uint32_t v = min(c, (uint32_t)kSegCap);
                uint32_t inc = v;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, off);
                    if (lane >= off) inc += y;
                }
                seg_off[r0 + lane] = carry + inc - v;
                carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
            if (lane == 0) seg_off[kSegCount] = carry;
        }
        __syncthreads();
        for (int sg = warp; sg < kSegCount; sg += kSelectThreads / 32) {
            const uint32_t o0 = seg_off[sg], ns = seg_off[sg + 1] - o0;
            const float* ss = t.seg_score + ((size_t)slot * kSegCount + sg) * kSegCap;
            const int32_t* sr = t.seg_row + ((size_t)slot * kSegCount + sg) * kSegCap;
            for (uint32_t i = lane; i < ns; i += 32) {
                const uint32_t dst = raw + o0 + i;
                if (dst < (uint32_t)cap) { cs[dst] = ss[i]; cr[dst] = sr[i]; }
            }
        }
        __syncthreads();
        raw += seg_off[kSegCount];
        seg_over = seg_flag != 0u;
    }
    const int n = (int)min(raw, (uint32_t)cap);
    const bool overflowed = raw > (uint32_t)cap || seg_over || t.overflow[slot] != 0u;   // sticky over the chunks of a batch
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
                            if (c[b] >= rem) { d = lane * 8 + b; sh_need = rem; sh_bin = c[b]; rem = 0u; }
                            else rem -= c[b];
                        }
                    }
                    sh_digit = (uint32_t)d;
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (unsigned long long)sh_digit;
            need = sh_need;
            if (sh_bin == need) {                                        // the whole bin is kept: every key >= prefix followed by zeros is
                prefix <<= shift;                                        // in the top k, the remaining digits need not be resolved
                break;                                                   // (block-uniform: sh_bin / sh_need are shared)
            }
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
        if (tid + j * kSelectThreads < n && r[j] >= kth) {
            if (pos < (uint32_t)m) keys[pos] = r[j];                       // keys are unique for scans; duplicated rerank candidates must not run past m
            ++pos;
        }
    // Intermediate passes only need the SET of kept candidates and the k-th best score: no sort.  The k-th key is the minimum
    // of the kept keys (the radix select may have stopped early, so `kth` itself need not be a key).
    unsigned long long kmin = ~0ull;
    if (!final_pass) {
#pragma unroll
        for (int j = 0; j < kSelectKeysPerThread; ++j)
            if (tid + j * kSelectThreads < n && r[j] >= kth) kmin = r[j] < kmin ? r[j] : kmin;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o2 = __shfl_xor_sync(0xFFFFFFFFu, kmin, off);
            kmin = o2 < kmin ? o2 : kmin;
        }
        if (lane == 0) wmin[warp] = kmin;
        __syncthreads();
    } else {
        bitonic_sort_desc(keys, m, tid, kSelectThreads);
    }
    for (int i = tid; i < keep; i += kSelectThreads) {
        const unsigned long long key = keys[i];
        cs[i] = key_score(key);
        cr[i] = (int32_t)key_row(key);
    }
    __syncthreads();                                                     // every thread has read cnt / overflow of this slot
    if (tid == 0) {
        if (final_pass) {
            // the slot is handed to the next batch clean (no separate init / carry kernels between batches)
            t.cnt[slot] = 0u; t.tau[slot] = -INFINITY; t.overflow[slot] = 0u;
            if (o.overflow && overflowed) o.overflow[out_base + slot] = 1u;
        } else {
            unsigned long long km = wmin[0];
            for (int w = 1; w < kSelectThreads / 32; ++w) km = wmin[w] < km ? wmin[w] : km;
            t.cnt[slot] = (uint32_t)keep;
            t.tau[slot] = (n >= k) ? key_score(km) : -INFINITY;
            t.overflow[slot] = overflowed ? 1u : 0u;
        }
    }
    if (final_pass && out_keys) {
        // exchange format of the sharded search: (ordered score << 32) | (0xFFFFFFFF - GLOBAL row); 0 = padding.  Descending
        // key order is (score desc, global row asc) on every shard, so the merge is a plain 64-bit merge.
        unsigned long long* ok = out_keys + (size_t)(out_base + slot) * k;
        for (int i = tid; i < k; i += kSelectThreads) ok[i] = i < keep ? keys[i] - (unsigned long long)row_offset : 0ull;
        if (tid == 0 && out_counts) out_counts[out_base + slot] = keep;
    } else if (final_pass) {
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

int launch_select(const TopkState& t, int n_slots, int k, int cap, bool final_pass, const SelectOut& o, cudaStream_t st) {
    if (n_slots <= 0) return DHR_OK;
    if (cap > kCandCap) return DHR_ERR_INVALID;
    int m = 2;
    while (m < std::min(k, cap)) m <<= 1;
    const size_t smem = (size_t)m * sizeof(unsigned long long);
    // the attribute is per device, so it is set on every launch (a process may hold indexes on several GPUs); the kernel's
    // 33 KiB of static shared memory count against the 48 KiB default, so even small k needs the opt-in
    DHR_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_select_kernel<<<n_slots, kSelectThreads, smem, st>>>(t, k, cap, final_pass ? 1 : 0, o);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

// ---------------------------------------------------------------------------------------------
// shard merge: P per-shard top-k lists of one query -> top-k of their union (replaces the per-query
// argsort of retrieval/merge.result.py:39 and is the step after the NCCL all-gather).
//
// Progressive bitonic merge, one CTA per query: the running best list (m = pow2 >= k items, descending)
// sits in the first half of a 2m-item shared buffer; the next shard's list is loaded into the second half;
// since both halves are sorted descending, one "flip" stage (compare i with 2m-1-i) followed by the
// log2(m) half-cleaner stages sorts the first half descending -- log2(2m) stages per shard instead of a
// full sort of P*k items, and shared memory is O(2m) whatever P.  A list that is NOT sorted on arrival
// (dhr_topk_merge accepts arbitrary input) is detected while loading and bitonic-sorted first.
//
// Two item types share the code: 64-bit packed keys (the sharded search's exchange format) and
// {ordered score, int64 row} pairs (generic entry point; rows rank by value when scores tie).
// ---------------------------------------------------------------------------------------------
struct MergeItem { unsigned int s; unsigned int pad; long long row; };

struct KeyTraits {
    typedef unsigned long long Item;
    __device__ static __forceinline__ Item pad() { return 0ull; }
    __device__ static __forceinline__ bool before(const Item& a, const Item& b) { return a > b; }   // a ranks strictly before b
};
struct PairTraits {
    typedef MergeItem Item;
    __device__ static __forceinline__ Item pad() { return MergeItem{0u, 0u, -1}; }
    __device__ static __forceinline__ bool before(const Item& a, const Item& b) {
        if (a.s != b.s) return a.s > b.s;
        return (unsigned long long)a.row < (unsigned long long)b.row;                                 // padding rows (-1) compare last
    }
};

constexpr int kMergeThreads = 256;

// descending bitonic sort of items[0, m)
template <typename T>
__device__ __forceinline__ void merge_sort_desc(typename T::Item* items, int m, int tid) {
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (m >> 1); t += kMergeThreads) {
                const int lo = ((t / stride) * (stride << 1)) + (t % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const typename T::Item a = items[lo], b = items[hi];
                const bool swap = desc ? T::before(b, a) : T::before(a, b);
                if (swap) { items[lo] = b; items[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// items[0, m) and items[m, 2m) both descending -> items[0, m) = the m best of the union, descending
template <typename T>
__device__ __forceinline__ void merge_two_desc(typename T::Item* items, int m, int tid) {
    __syncthreads();
    for (int t = tid; t < m; t += kMergeThreads) {                        // flip stage: the first half now holds the m best (bitonic)
        const typename T::Item a = items[t], b = items[2 * m - 1 - t];
        if (T::before(b, a)) { items[t] = b; items[2 * m - 1 - t] = a; }
    }
    for (int stride = m >> 1; stride > 0; stride >>= 1) {                 // half-cleaners on the first half only
        __syncthreads();
        for (int t = tid; t < (m >> 1); t += kMergeThreads) {
            const int lo = ((t / stride) * (stride << 1)) + (t % stride);
            const int hi = lo + stride;
            const typename T::Item a = items[lo], b = items[hi];
            if (T::before(b, a)) { items[lo] = b; items[hi] = a; }
        }
    }
    __syncthreads();
}

template <typename T, typename Load>
__device__ __forceinline__ void merge_parts(typename T::Item* items, int P, int k, int m, int tid, Load load) {
    __shared__ int unsorted;
    for (int p = 0; p < P; ++p) {
        typename T::Item* dst = items + (p == 0 ? 0 : m);
        if (tid == 0) unsorted = 0;
        __syncthreads();
        for (int i = tid; i < m; i += kMergeThreads) dst[i] = i < k ? load(p, i) : T::pad();
        __syncthreads();
        for (int i = tid; i + 1 < k; i += kMergeThreads)
            if (T::before(dst[i + 1], dst[i])) unsorted = 1;
        __syncthreads();
        if (unsorted) merge_sort_desc<T>(dst, m, tid);                    // block-uniform branch
        if (p > 0) merge_two_desc<T>(items, m, tid);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kMergeThreads)
merge_keys_kernel(int P, int k, int m, const unsigned long long* __restrict__ keys, long long part_stride, float* out_scores,
                  long long* out_rows, unsigned long long* out_keys) {
    extern __shared__ __align__(16) unsigned char raw_smem[];
    unsigned long long* items = (unsigned long long*)raw_smem;
    const int q = blockIdx.x, tid = threadIdx.x;
    merge_parts<KeyTraits>(items, P, k, m, tid,
                           [&](int p, int i) { return keys[(size_t)p * part_stride + (size_t)q * k + i]; });
    for (int i = tid; i < k; i += kMergeThreads) {
        const unsigned long long key = items[i];
        if (out_keys) out_keys[(size_t)q * k + i] = key;
        if (out_scores) {
            out_scores[(size_t)q * k + i] = key ? key_score(key) : -INFINITY;
            out_rows[(size_t)q * k + i] = key ? (long long)key_row(key) : -1;
        }
    }
}

__global__ void __launch_bounds__(kMergeThreads)
merge_pairs_kernel(int P, int Q, int k, int m, const float* __restrict__ scores, const long long* __restrict__ rows,
                   float* out_scores, long long* out_rows) {
    extern __shared__ __align__(16) unsigned char raw_smem[];
    MergeItem* items = (MergeItem*)raw_smem;
    const int q = blockIdx.x, tid = threadIdx.x;
    merge_parts<PairTraits>(items, P, k, m, tid, [&](int p, int i) {
        const size_t src = ((size_t)p * Q + q) * k + i;
        const long long r = rows[src];
        MergeItem it;
        it.pad = 0u; it.row = r;
        it.s = r >= 0 ? float_to_ordered(scores[src] + 0.0f) : 0u;
        if (r < 0) it.row = -1;
        return it;
    });
    for (int i = tid; i < k; i += kMergeThreads) {
        const MergeItem it = items[i];
        const bool valid = it.row >= 0;
        out_scores[(size_t)q * k + i] = valid ? ordered_to_float(it.s) : -INFINITY;
        out_rows[(size_t)q * k + i] = valid ? it.row : -1;
    }
}

static int merge_pow2(int k) { int m = 2; while (m < k) m <<= 1; return m; }

}  // namespace dhr

using namespace dhr;

extern "C" int dhr_merge_keys(int device, int n_parts, int n_queries, int k, const uint64_t* keys, int64_t part_stride,
                              float* out_scores, int64_t* out_rows, uint64_t* out_keys, void* stream) {
    if (n_parts <= 0 || n_queries < 0 || k <= 0 || !keys || part_stride < (int64_t)n_queries * k) return DHR_ERR_INVALID;
    if ((!out_scores) != (!out_rows) || (!out_scores && !out_keys)) return DHR_ERR_INVALID;
    if (n_queries == 0) return DHR_OK;
    if (!is_device_pointer(keys) || (out_scores && (!is_device_pointer(out_scores) || !is_device_pointer(out_rows))) ||
        (out_keys && !is_device_pointer(out_keys)))
        return DHR_ERR_INVALID;                                           // device-resident exchange buffers only (stream-ordered, no sync)
    const int m = merge_pow2(k);
    const size_t smem = (size_t)2 * m * sizeof(unsigned long long);
    if (smem > 200 * 1024) return DHR_ERR_UNSUPPORTED;                    // k <= 8192
    DHR_CUDA(cudaSetDevice(device));
    if (smem > 48 * 1024) DHR_CUDA(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_keys_kernel<<<n_queries, kMergeThreads, smem, (cudaStream_t)stream>>>(n_parts, k, m, (const unsigned long long*)keys,
                                                                                (long long)part_stride, out_scores, (long long*)out_rows,
                                                                                (unsigned long long*)out_keys);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}

extern "C" int dhr_topk_merge(int device, int n_parts, int n_queries, int k, const float* scores, const int64_t* rows,
                              float* out_scores, int64_t* out_rows, void* stream) {
    if (n_parts <= 0 || n_queries < 0 || k <= 0 || !scores || !rows || !out_scores || !out_rows) return DHR_ERR_INVALID;
    if (n_queries == 0) return DHR_OK;
    const int m = merge_pow2(k);
    const size_t smem = (size_t)2 * m * sizeof(MergeItem);
    if (smem > 200 * 1024) return DHR_ERR_UNSUPPORTED;                    // k <= 4096 (any number of parts)
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
    if (smem > 48 * 1024) MERGE_CUDA(cudaFuncSetAttribute(merge_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_pairs_kernel<<<n_queries, kMergeThreads, smem, st>>>(n_parts, n_queries, k, m, d_s, d_r, d_os, d_or);
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
