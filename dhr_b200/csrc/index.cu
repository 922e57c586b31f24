// index.cu -- index lifetime: HBM-resident layout, row ingestion, validation.
//
// Replaces the load path of castorini/dhr retrieval/gip_retrieval.py:289-315 (pickle.load ->
// shard slice -> .cuda()).  The caller hands rows in the reference's layout -- values
// [n, S*G + C] (fp16 as stored by encode.py:156,165 / densify_corpus.py:68-72, or the fp32 copy
// the CPU path makes at gip_retrieval.py:313) and slice indices [n, S] -- and the index keeps
// three row-major device arrays:
//    lexv [N][D_pad] fp16    lexical values, D_pad = S_pad*G
//    lexi [N][S_pad] codes   uint8 or uint16; all-zero slices are stored as CODE_EMPTY
//    dns  [N][C_pad] fp16    dense [CLS] block
#include <algorithm>
#include <mutex>
#include <string.h>
#include <string>

#include "internal.h"

namespace dhr {

static thread_local std::string g_last_cuda_error;

void set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), what, file, line);
    g_last_cuda_error = buf;
    cudaGetLastError();   // clear the sticky-less error state
}

bool is_device_pointer(const void* p) {
    if (!p) return false;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int ensure_device_buffer(void** p, size_t* cur, size_t need) {
    if (*cur >= need && *p) return DHR_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cur = 0; }
    if (need == 0) return DHR_OK;
    DHR_CUDA(cudaMalloc(p, need));
    *cur = need;
    return DHR_OK;
}

cudaEvent_t EventPool::get() {
    if (used == ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ev.push_back(e);
    }
    return ev[used++];
}
void EventPool::destroy() {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
    used = 0;
}

static int idx_dtype_size(int dt) {
    switch (dt) {
        case DHR_IDX_U8: case DHR_IDX_I8: return 1;
        case DHR_IDX_I16: case DHR_IDX_U16: return 2;
        case DHR_IDX_I32: return 4;
        case DHR_IDX_I64: return 8;
        default: return 0;
    }
}

__device__ __forceinline__ long long load_index_value(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case DHR_IDX_U8:  return ((const uint8_t*)base)[off];
        case DHR_IDX_I8:  return ((const int8_t*)base)[off];
        case DHR_IDX_I16: return ((const int16_t*)base)[off];
        case DHR_IDX_U16: return ((const uint16_t*)base)[off];
        case DHR_IDX_I32: return ((const int32_t*)base)[off];
        default:          return ((const long long*)base)[off];
    }
}

__device__ __forceinline__ __half load_value_as_half(const void* base, int dtype, size_t off, bool* lossy) {
    if (dtype == DHR_VAL_F16) return ((const __half*)base)[off];
    const float f = ((const float*)base)[off];
    const __half h = __float2half_rn(f);
    if (__half2float(h) != f && f == f) *lossy = true;
    return h;
}

// one thread per (row, padded slice): G values + the code
template <typename CodeT>
__global__ void ingest_lexical_kernel(long long n, int S, int G, int S_pad, int W, int val_dtype, const void* vals,
                                      long long vstride, int idx_dtype, const void* idx, long long istride,
                                      __half* lexv, CodeT* lexi, long long row_base, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * S_pad) return;
    const long long r = i / S_pad;
    const int s = (int)(i % S_pad);
    __half* out = lexv + ((size_t)(row_base + r) * S_pad + s) * G;
    CodeT* oc = lexi + (size_t)(row_base + r) * S_pad + s;
    if (s >= S) {
        for (int g = 0; g < G; ++g) out[g] = __float2half_rn(0.f);
        *oc = (CodeT)CodeTraits<CodeT>::kEmpty;
        return;
    }
    bool lossy = false, nonzero = false;
    for (int g = 0; g < G; ++g) {
        const __half h = load_value_as_half(vals, val_dtype, (size_t)r * vstride + (size_t)s * G + g, &lossy);
        out[g] = h;
        nonzero |= (__half_as_ushort(h) & 0x7FFFu) != 0;
    }
    if (lossy) atomicOr(flags + 0, 1);
    const long long v = load_index_value(idx, idx_dtype, (size_t)r * istride + s);
    uint32_t code = CodeTraits<CodeT>::kEmpty;
    if (nonzero) {
        if (v < 0 || v > (long long)CodeTraits<CodeT>::kMax) { atomicOr(flags + 1, 1); code = CodeTraits<CodeT>::kNoMatch; }
        else { code = (uint32_t)v; atomicMax(flags + 3, (int)code + 1); }
    }
    *oc = (CodeT)code;
}

__global__ void ingest_dense_kernel(long long n, int D, int C, int C_pad, int val_dtype, const void* vals, long long vstride,
                                    __half* dns, long long row_base, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C_pad) return;
    const long long r = i / C_pad;
    const int c = (int)(i % C_pad);
    bool lossy = false;
    __half h = __float2half_rn(0.f);
    if (c < C) h = load_value_as_half(vals, val_dtype, (size_t)r * vstride + D + c, &lossy);
    dns[(size_t)(row_base + r) * C_pad + c] = h;
    if (lossy) atomicOr(flags + 0, 1);
}

// Tiled lexical copy for K1t, built once at finalize from the row-major arrays:
//   [tile of 512 rows][chunk of 4 slices]{ codes TCode [512][4] | vals fp16 [4][512][G] }
// TCode is uint8 when the largest stored code is <= 253 (CODE_EMPTY/NOMATCH map to 0xFF), else uint16 (0xFFFF).
template <typename CodeT, typename TCode>
__global__ void build_lext_kernel(long long n_rows, long long n_rows_pad, int S_pad, int G, uint32_t empty_code, const __half* __restrict__ lexv,
                                  const CodeT* __restrict__ lexi, uint8_t* __restrict__ lext) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows_pad * S_pad) return;
    const long long r = i / S_pad;
    const int s = (int)(i % S_pad);
    const long long tile = r / kLexTileRows;
    const int tp = (int)(r % kLexTileRows), chunk = s / kLexTileSlices, tj = s % kLexTileSlices;
    const size_t pblock = (size_t)kLexTileRows * kLexTileSlices * (sizeof(TCode) + 2 * (size_t)G);
    uint8_t* blk = lext + ((size_t)tile * (S_pad / kLexTileSlices) + chunk) * pblock;
    TCode* tc = (TCode*)blk + (size_t)tp * kLexTileSlices + tj;
    __half* tv = (__half*)(blk + (size_t)kLexTileRows * kLexTileSlices * sizeof(TCode)) + ((size_t)tj * kLexTileRows + tp) * G;
    // empty slices store `empty_code`: the narrow layout uses rt (= largest code + 1), the index of the always-empty bucket of the
    // query tables, so K1t needs no clamp; the wide layout uses 0xFFFF, which no query code equals
    const uint32_t kTEmpty = empty_code;
    if (r >= n_rows) {
        *tc = (TCode)kTEmpty;
        for (int g = 0; g < G; ++g) tv[g] = __float2half_rn(0.f);
        return;
    }
    const uint32_t code = lexi[(size_t)r * S_pad + s];
    *tc = code > CodeTraits<TCode>::kMax ? (TCode)kTEmpty : (TCode)code;
    const __half* src = lexv + ((size_t)r * S_pad + s) * G;
    for (int g = 0; g < G; ++g) tv[g] = src[g];
}

// Postings copy for K1p, built once at finalize: one CTA per (tile of 512 rows, chunk of 4 slices), thread = row.  Per slice a
// counting sort by code over the non-empty rows; item = {row-in-tile | code << 16, G fp16 values} of EW words.  The order of
// items with the same code is the order of the atomic ranks (arbitrary); a slice holds each row once, so the scores do not
// depend on it.
template <typename CodeT>
__global__ void __launch_bounds__(kLexTileRows)
build_lexp_kernel(long long n_rows, int S_pad, int G, int EW, int rt, int block_stride, const __half* __restrict__ lexv,
                  const CodeT* __restrict__ lexi, uint8_t* __restrict__ lexp, uint32_t* __restrict__ nbytes) {
    __shared__ uint32_t cnt[256], start[256];
    __shared__ uint32_t slice_n;
    const int chunk = blockIdx.x, n_chunks = gridDim.x;
    const long long tile = blockIdx.y;
    const int p = threadIdx.x;
    const long long row = tile * kLexTileRows + p;
    uint8_t* blk = lexp + ((size_t)tile * n_chunks + chunk) * (size_t)block_stride;
    uint32_t* hdr = (uint32_t*)blk;
    uint32_t* items = (uint32_t*)(blk + 16);
    uint32_t base = 0;
    for (int j = 0; j < kLexTileSlices; ++j) {
        const int s = chunk * kLexTileSlices + j;
        uint32_t code = 0xFFFFFFFFu;
        if (row < n_rows) code = lexi[(size_t)row * S_pad + s];
        const bool valid = code < (uint32_t)rt;                           // CODE_EMPTY / CODE_NOMATCH are above every stored code
        if (p < 256) cnt[p] = 0u;
        __syncthreads();
        uint32_t rank = 0;
        if (valid) rank = atomicAdd(&cnt[code], 1u);
        __syncthreads();
        if (p == 0) {
            uint32_t run = 0;
            for (int c = 0; c < rt; ++c) { start[c] = run; run += cnt[c]; }
            slice_n = run;
            hdr[j] = (base << 16) | run;
        }
        __syncthreads();
        if (valid) {
            uint32_t* it = items + (size_t)(base + start[code] + rank) * EW;
            const __half* v = lexv + ((size_t)row * S_pad + s) * G;
            it[0] = (uint32_t)p | (code << 16);
            for (int w = 1; w < EW; ++w) {
                const int g0 = 2 * (w - 1), g1 = g0 + 1;
                const uint32_t lo = g0 < G ? __half_as_ushort(v[g0]) : 0u;
                const uint32_t hi = g1 < G ? __half_as_ushort(v[g1]) : 0u;
                it[w] = lo | (hi << 16);
            }
        }
        base += slice_n;
        __syncthreads();
    }
    if (p == 0) nbytes[(size_t)tile * n_chunks + chunk] = (16u + base * (uint32_t)EW * 4u + 15u) & ~15u;
}

// inverse: rebuild the row-major arrays from the postings (rows of empty slices were pre-filled with value 0 / CODE_EMPTY)
template <typename CodeT>
__global__ void __launch_bounds__(kLexTileRows)
unbuild_lexp_kernel(long long n_rows, int S_pad, int G, int EW, int block_stride, const uint8_t* __restrict__ lexp, __half* __restrict__ lexv,
                    CodeT* __restrict__ lexi) {
    const int chunk = blockIdx.x, n_chunks = gridDim.x;
    const long long tile = blockIdx.y;
    const uint8_t* blk = lexp + ((size_t)tile * n_chunks + chunk) * (size_t)block_stride;
    const uint32_t* hdr = (const uint32_t*)blk;
    const uint32_t* items = (const uint32_t*)(blk + 16);
    for (int j = 0; j < kLexTileSlices; ++j) {
        const uint32_t n = hdr[j] & 0xFFFFu, first = hdr[j] >> 16;
        const int s = chunk * kLexTileSlices + j;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t* it = items + (size_t)(first + i) * EW;
            const long long row = tile * kLexTileRows + (it[0] & 0xFFFFu);
            if (row >= n_rows) continue;
            lexi[(size_t)row * S_pad + s] = (CodeT)((it[0] >> 16) & 0xFFu);
            __half* dst = lexv + ((size_t)row * S_pad + s) * G;
            for (int g = 0; g < G; ++g) dst[g] = __ushort_as_half((unsigned short)((it[1 + (g >> 1)] >> (16 * (g & 1))) & 0xFFFFu));
        }
    }
}

template <typename CodeT>
__global__ void fill_empty_lexical_kernel(long long total_slices, int G, __half* __restrict__ lexv, CodeT* __restrict__ lexi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_slices) return;
    lexi[i] = (CodeT)CodeTraits<CodeT>::kEmpty;
    for (int g = 0; g < G; ++g) lexv[(size_t)i * G + g] = __float2half_rn(0.f);
}

// K-blocked copy of the dense block for K2 (TS variant): [tile of 128 rows][k-block][128 rows][64 cols], so that one
// 16 KiB TMA stage is one contiguous piece of HBM (the row-major block would be read as 128-byte pieces 2*C_pad apart,
// which measured ~2.2 TB/s).  One thread moves one 16-byte vector; rows >= n_rows and columns >= C_pad are zero.
__global__ void build_dnst_kernel(long long n_rows, long long n_rows_pad, int C_pad, int n_kblocks, const __half* __restrict__ dns,
                                  __half* __restrict__ dnst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long vec_per_row = (long long)n_kblocks * (kDenseTileCols / 8);
    if (i >= n_rows_pad * vec_per_row) return;
    const long long r = i / vec_per_row;
    const int v = (int)(i % vec_per_row);
    const int kb = v / (kDenseTileCols / 8), cv = v % (kDenseTileCols / 8);
    const int col = kb * kDenseTileCols + cv * 8;
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (r < n_rows && col < C_pad) x = *(const uint4*)(dns + (size_t)r * C_pad + col);
    const long long tile = r / kDenseTileRows;
    const int tr = (int)(r % kDenseTileRows);
    *(uint4*)(dnst + (((size_t)tile * n_kblocks + kb) * kDenseTileRows + tr) * kDenseTileCols + cv * 8) = x;
}

// inverse of build_lext_kernel / build_dnst_kernel: rebuild the row-major arrays from the tiled copies (option "rowmajor")
template <typename CodeT, typename TCode>
__global__ void unbuild_lext_kernel(long long n_rows, int S_pad, int G, uint32_t empty_code, const uint8_t* __restrict__ lext, __half* __restrict__ lexv,
                                    CodeT* __restrict__ lexi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * S_pad) return;
    const long long r = i / S_pad;
    const int s = (int)(i % S_pad);
    const long long tile = r / kLexTileRows;
    const int tp = (int)(r % kLexTileRows), chunk = s / kLexTileSlices, tj = s % kLexTileSlices;
    const size_t pblock = (size_t)kLexTileRows * kLexTileSlices * (sizeof(TCode) + 2 * (size_t)G);
    const uint8_t* blk = lext + ((size_t)tile * (S_pad / kLexTileSlices) + chunk) * pblock;
    const TCode tc = ((const TCode*)blk)[(size_t)tp * kLexTileSlices + tj];
    const __half* tv = (const __half*)(blk + (size_t)kLexTileRows * kLexTileSlices * sizeof(TCode)) + ((size_t)tj * kLexTileRows + tp) * G;
    const uint32_t kTEmpty = empty_code;
    lexi[(size_t)r * S_pad + s] = (uint32_t)tc == kTEmpty ? (CodeT)CodeTraits<CodeT>::kEmpty : (CodeT)tc;
    __half* dst = lexv + ((size_t)r * S_pad + s) * G;
    for (int g = 0; g < G; ++g) dst[g] = tv[g];
}

__global__ void unbuild_dnst_kernel(long long n_rows, int C_pad, int n_kblocks, const __half* __restrict__ dnst, __half* __restrict__ dns) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long vec_per_row = C_pad / 8;
    if (i >= n_rows * vec_per_row) return;
    const long long r = i / vec_per_row;
    const int col = (int)(i % vec_per_row) * 8;
    const int kb = col / kDenseTileCols, cv = (col % kDenseTileCols) / 8;
    const long long tile = r / kDenseTileRows;
    const int tr = (int)(r % kDenseTileRows);
    *(uint4*)(dns + (size_t)r * C_pad + col) =
        *(const uint4*)(dnst + (((size_t)tile * n_kblocks + kb) * kDenseTileRows + tr) * kDenseTileCols + cv * 8);
}

static bool tiled_copies_complete(const dhr_index* h) {
    const Geometry& g = h->g;
    return (g.D_pad == 0 || h->lext || h->lexp) && (g.C_pad == 0 || h->dnst);
}

static int drop_rowmajor_unchecked(dhr_index* h);
int drop_rowmajor(dhr_index* h) {
    if (!h->finalized) return DHR_OK;
    return drop_rowmajor_unchecked(h);
}
static int drop_rowmajor_unchecked(dhr_index* h) {
    if (!tiled_copies_complete(h) || h->n_rows == 0) return DHR_OK;   // nothing to rebuild from: keep them
    DHR_CUDA(cudaSetDevice(h->device));
    DHR_CUDA(cudaDeviceSynchronize());
    if (h->lexv) { cudaFree(h->lexv); h->lexv = nullptr; }
    if (h->lexi) { cudaFree(h->lexi); h->lexi = nullptr; }
    if (h->dns) { cudaFree(h->dns); h->dns = nullptr; }
    return DHR_OK;
}

int ensure_rowmajor(dhr_index* h) {
    const Geometry& g = h->g;
    const bool have = (g.D_pad == 0 || (h->lexv && h->lexi)) && (g.C_pad == 0 || h->dns);
    if (have) return DHR_OK;
    if (!tiled_copies_complete(h)) return DHR_ERR_STATE;
    DHR_CUDA(cudaSetDevice(h->device));
    const size_t rows = (size_t)(h->capacity > 0 ? h->capacity : 1);
    if (g.D_pad > 0 && !h->lexv) {
        DHR_CUDA(cudaMalloc(&h->lexv, rows * g.D_pad * 2));
        DHR_CUDA(cudaMalloc(&h->lexi, rows * g.S_pad * g.code_bytes));
        const long long total = h->n_rows * g.S_pad;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        const bool wide = std::max(1, h->max_code + 1) > 254;
        if (!h->lext) {                                                   // postings layout: pre-fill, then scatter the items
            const LexTileGeom lp = lex_post_geom(g, std::max(1, h->max_code + 1));
            const int EW = lex_post_entry_words(g.G);
            dim3 grid((unsigned)(g.S_pad / kLexTileSlices), (unsigned)((h->n_rows + kLexTileRows - 1) / kLexTileRows));
            if (g.code_bytes == 1) {
                fill_empty_lexical_kernel<uint8_t><<<blocks, 256>>>(total, g.G, h->lexv, (uint8_t*)h->lexi);
                unbuild_lexp_kernel<uint8_t><<<grid, kLexTileRows>>>(h->n_rows, g.S_pad, g.G, EW, lp.pblock_bytes, h->lexp, h->lexv, (uint8_t*)h->lexi);
            } else {
                fill_empty_lexical_kernel<uint16_t><<<blocks, 256>>>(total, g.G, h->lexv, (uint16_t*)h->lexi);
                unbuild_lexp_kernel<uint16_t><<<grid, kLexTileRows>>>(h->n_rows, g.S_pad, g.G, EW, lp.pblock_bytes, h->lexp, h->lexv, (uint16_t*)h->lexi);
            }
        } else if (wide) unbuild_lext_kernel<uint16_t, uint16_t><<<blocks, 256>>>(h->n_rows, g.S_pad, g.G, 0xFFFFu, h->lext, h->lexv, (uint16_t*)h->lexi);
        else if (g.code_bytes == 1) unbuild_lext_kernel<uint8_t, uint8_t><<<blocks, 256>>>(h->n_rows, g.S_pad, g.G, (uint32_t)std::max(1, h->max_code + 1), h->lext, h->lexv, (uint8_t*)h->lexi);
        else unbuild_lext_kernel<uint16_t, uint8_t><<<blocks, 256>>>(h->n_rows, g.S_pad, g.G, (uint32_t)std::max(1, h->max_code + 1), h->lext, h->lexv, (uint16_t*)h->lexi);
        DHR_CUDA(cudaGetLastError());
    }
    if (g.C_pad > 0 && !h->dns) {
        DHR_CUDA(cudaMalloc(&h->dns, rows * g.C_pad * 2));
        const int nkb = (g.C_pad + kDenseTileCols - 1) / kDenseTileCols;
        const long long total = h->n_rows * (g.C_pad / 8);
        unbuild_dnst_kernel<<<(unsigned)((total + 255) / 256), 256>>>(h->n_rows, g.C_pad, nkb, h->dnst, h->dns);
        DHR_CUDA(cudaGetLastError());
    }
    DHR_CUDA(cudaDeviceSynchronize());
    h->stats.rowmajor_rebuilds++;
    return DHR_OK;
}

static int ingest_device(dhr_index* h, long long n, int val_dtype, const void* d_vals, long long vstride, int idx_dtype,
                         const void* d_idx, long long istride) {
    const Geometry& g = h->g;
    const int W = g.S * g.G + g.C;
    if (g.S_pad > 0) {
        const long long total = n * g.S_pad;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        if (g.code_bytes == 1)
            ingest_lexical_kernel<uint8_t><<<blocks, 256>>>(n, g.S, g.G, g.S_pad, W, val_dtype, d_vals, vstride, idx_dtype, d_idx,
                                                            istride, h->lexv, (uint8_t*)h->lexi, h->n_rows, h->d_flags);
        else
            ingest_lexical_kernel<uint16_t><<<blocks, 256>>>(n, g.S, g.G, g.S_pad, W, val_dtype, d_vals, vstride, idx_dtype, d_idx,
                                                             istride, h->lexv, (uint16_t*)h->lexi, h->n_rows, h->d_flags);
        DHR_CUDA(cudaGetLastError());
    }
    if (g.C_pad > 0) {
        const long long total = n * g.C_pad;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        ingest_dense_kernel<<<blocks, 256>>>(n, g.S * g.G, g.C, g.C_pad, val_dtype, d_vals, vstride, h->dns, h->n_rows, h->d_flags);
        DHR_CUDA(cudaGetLastError());
    }
    return DHR_OK;
}

}  // namespace dhr

using namespace dhr;

extern "C" {

int dhr_version(void) { return DHR_B200_VERSION; }

const char* dhr_strerror(int status) {
    switch (status) {
        case DHR_OK: return "ok";
        case DHR_ERR_INVALID: return "invalid argument";
        case DHR_ERR_CUDA: return "CUDA runtime error";
        case DHR_ERR_NOMEM: return "out of memory";
        case DHR_ERR_UNSUPPORTED: return "unsupported shape or size";
        case DHR_ERR_LOSSY: return "fp32 corpus value is not representable in fp16";
        case DHR_ERR_IDX_RANGE: return "corpus slice index outside the code range";
        case DHR_ERR_STATE: return "invalid call order for this index";
        case DHR_ERR_NO_DEVICE: return "no usable CUDA device";
        default: return "unknown status";
    }
}

const char* dhr_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

int dhr_device_count(int* count) {
    if (!count) return DHR_ERR_INVALID;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_cuda_error(e, "cudaGetDeviceCount", __FILE__, __LINE__); *count = 0; return DHR_ERR_NO_DEVICE; }
    *count = n;
    return DHR_OK;
}

int dhr_index_create(dhr_index** out, int device, int64_t cap_rows, int n_slices, int group, int n_dense, int idx_dtype,
                     int64_t row_offset, unsigned flags) {
    if (!out || cap_rows < 0 || n_slices < 0 || n_dense < 0 || group < 1 || row_offset < 0) return DHR_ERR_INVALID;
    if (n_slices == 0 && n_dense == 0) return DHR_ERR_INVALID;
    if (group > DHR_MAX_GROUP) return DHR_ERR_UNSUPPORTED;
    if (cap_rows > 0x7FFFFFF0ll) return DHR_ERR_UNSUPPORTED;
    if (n_slices > 0 && idx_dtype_size(idx_dtype) == 0) return DHR_ERR_INVALID;
    int ndev = 0;
    DHR_TRY(dhr_device_count(&ndev));
    if (device < 0 || device >= ndev) return DHR_ERR_NO_DEVICE;
    DHR_CUDA(cudaSetDevice(device));

    dhr_index* h = new (std::nothrow) dhr_index();
    if (!h) return DHR_ERR_NOMEM;
    h->device = device;
    h->capacity = cap_rows;
    h->row_offset = row_offset;
    h->idx_dtype = n_slices > 0 ? idx_dtype : DHR_IDX_NONE;
    Geometry& g = h->g;
    g.S = n_slices; g.G = group; g.C = n_dense;
    g.S_pad = (int)round_up(n_slices, 16);
    g.D_pad = g.S_pad * group;
    g.C_pad = (int)round_up(n_dense, 8);
    const bool narrow = (flags & DHR_INDEX_NARROW_CODES) != 0;
    h->keep_rowmajor = (flags & DHR_INDEX_KEEP_ROWMAJOR) != 0;
    h->use_postings = (flags & DHR_INDEX_LEX_POSTINGS) != 0;
    g.code_bytes = (idx_dtype_size(h->idx_dtype) == 1 || narrow || n_slices == 0) ? 1 : 2;
    g.unit_halves = (group % 8 == 0) ? group : (group % 4 == 0) ? 2 * group : (group % 2 == 0) ? 4 * group : 8 * group;
    g.unit_slices = g.unit_halves / group;
    g.n_units = g.S_pad / g.unit_slices;
    g.n_chunks = g.C_pad / 8;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) { h->num_sms = prop.multiProcessorCount; h->smem_per_sm = (long long)prop.sharedMemPerMultiprocessor; }
    int status = DHR_OK;
    auto fail = [&](int s) { dhr_index_close(h); return s; };
    const size_t rows = (size_t)(cap_rows > 0 ? cap_rows : 1);
    if (g.D_pad > 0) {
        if (cudaMalloc(&h->lexv, rows * g.D_pad * 2) != cudaSuccess) return fail(DHR_ERR_NOMEM);
        if (cudaMalloc(&h->lexi, rows * g.S_pad * g.code_bytes) != cudaSuccess) return fail(DHR_ERR_NOMEM);
    }
    if (g.C_pad > 0 && cudaMalloc(&h->dns, rows * g.C_pad * 2) != cudaSuccess) return fail(DHR_ERR_NOMEM);
    if (cudaMalloc(&h->d_flags, 4 * sizeof(int)) != cudaSuccess) return fail(DHR_ERR_NOMEM);
    if (cudaMemset(h->d_flags, 0, 4 * sizeof(int)) != cudaSuccess) return fail(DHR_ERR_CUDA);
    *out = h;
    return status;
}

int dhr_index_append(dhr_index* h, int64_t n, int val_dtype, const void* vals, int64_t vstride, int idx_dtype,
                     const void* idx, int64_t istride) {
    if (!h || n < 0) return DHR_ERR_INVALID;
    if (h->finalized) return DHR_ERR_STATE;
    if (n == 0) return DHR_OK;
    const Geometry& g = h->g;
    const int W = g.S * g.G + g.C;
    if (!vals || vstride < W || (val_dtype != DHR_VAL_F16 && val_dtype != DHR_VAL_F32)) return DHR_ERR_INVALID;
    if (g.S > 0 && (!idx || istride < g.S || idx_dtype_size(idx_dtype) == 0)) return DHR_ERR_INVALID;
    if (h->n_rows + n > h->capacity) return DHR_ERR_INVALID;
    DHR_CUDA(cudaSetDevice(h->device));
    const size_t vsz = val_dtype == DHR_VAL_F16 ? 2 : 4;
    const size_t isz = g.S > 0 ? (size_t)idx_dtype_size(idx_dtype) : 0;
    const bool v_dev = is_device_pointer(vals);
    const bool i_dev = g.S > 0 ? is_device_pointer(idx) : true;
    if (v_dev && i_dev) {
        DHR_TRY(ingest_device(h, n, val_dtype, vals, vstride, idx_dtype, idx, istride));
        DHR_CUDA(cudaDeviceSynchronize());
        h->n_rows += n;
        return DHR_OK;
    }
    // host source: stage in pieces of <= 64 MiB of values (packed, row stride = W / S)
    const size_t row_in = (size_t)W * vsz;
    long long piece = (long long)((64ull << 20) / (row_in ? row_in : 1));
    if (piece < 1) piece = 1;
    for (long long r0 = 0; r0 < n; r0 += piece) {
        const long long m = (n - r0 < piece) ? (n - r0) : piece;
        const void* dv = nullptr; const void* di = nullptr;
        long long dvs = vstride, dis = istride;
        if (v_dev) dv = (const uint8_t*)vals + (size_t)r0 * vstride * vsz;
        else {
            DHR_TRY(ensure_device_buffer(&h->stage_a, &h->stage_a_bytes, (size_t)piece * row_in));
            DHR_CUDA(cudaMemcpy2D(h->stage_a, row_in, (const uint8_t*)vals + (size_t)r0 * vstride * vsz, (size_t)vstride * vsz,
                                  row_in, (size_t)m, cudaMemcpyHostToDevice));
            dv = h->stage_a; dvs = W;
        }
        if (g.S > 0) {
            if (i_dev) di = (const uint8_t*)idx + (size_t)r0 * istride * isz;
            else {
                DHR_TRY(ensure_device_buffer(&h->stage_b, &h->stage_b_bytes, (size_t)piece * g.S * isz));
                DHR_CUDA(cudaMemcpy2D(h->stage_b, (size_t)g.S * isz, (const uint8_t*)idx + (size_t)r0 * istride * isz,
                                      (size_t)istride * isz, (size_t)g.S * isz, (size_t)m, cudaMemcpyHostToDevice));
                di = h->stage_b; dis = g.S;
            }
        }
        DHR_TRY(ingest_device(h, m, val_dtype, dv, dvs, idx_dtype, di, dis));
        DHR_CUDA(cudaDeviceSynchronize());
        h->n_rows += m;
    }
    return DHR_OK;
}

int dhr_index_finalize(dhr_index* h) {
    if (!h) return DHR_ERR_INVALID;
    if (h->finalized) return DHR_OK;
    DHR_CUDA(cudaSetDevice(h->device));
    int flags[4] = {0, 0, 0, 0};
    DHR_CUDA(cudaMemcpy(flags, h->d_flags, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[0]) return DHR_ERR_LOSSY;
    if (flags[1]) return DHR_ERR_IDX_RANGE;
    h->max_code = flags[3] - 1;
    const int rt_fin = std::max(1, h->max_code + 1);
    if (h->g.S_pad > 0 && h->n_rows > 0 && h->use_postings && lex_post_supported(h->g, rt_fin)) {
        // postings layout (K1p): non-empty passages of every (tile, slice) sorted by code
        const Geometry& g = h->g;
        const LexTileGeom lp = lex_post_geom(g, rt_fin);
        const long long n_tiles = (h->n_rows + kLexTileRows - 1) / kLexTileRows;
        const int n_chunks = g.S_pad / kLexTileSlices;
        h->lexp_bytes = (size_t)n_tiles * n_chunks * (size_t)lp.pblock_bytes;
        if (cudaMalloc(&h->lexp, h->lexp_bytes) != cudaSuccess) { cudaGetLastError(); h->lexp = nullptr; h->lexp_bytes = 0; }
        if (h->lexp && cudaMalloc(&h->lexp_nbytes, (size_t)n_tiles * n_chunks * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError(); cudaFree(h->lexp); h->lexp = nullptr; h->lexp_bytes = 0; h->lexp_nbytes = nullptr;
        }
        if (h->lexp) {
            dim3 grid((unsigned)n_chunks, (unsigned)n_tiles);
            const int EW = lex_post_entry_words(g.G);
            if (g.code_bytes == 1)
                build_lexp_kernel<uint8_t><<<grid, kLexTileRows>>>(h->n_rows, g.S_pad, g.G, EW, rt_fin, lp.pblock_bytes, h->lexv, (const uint8_t*)h->lexi, h->lexp, h->lexp_nbytes);
            else
                build_lexp_kernel<uint16_t><<<grid, kLexTileRows>>>(h->n_rows, g.S_pad, g.G, EW, rt_fin, lp.pblock_bytes, h->lexv, (const uint16_t*)h->lexi, h->lexp, h->lexp_nbytes);
            DHR_CUDA(cudaGetLastError());
            DHR_CUDA(cudaDeviceSynchronize());
            h->lex_layout = 1;
        }
    }
    if (h->g.S_pad > 0 && h->n_rows > 0 && !h->lexp && lex_tile_supported(h->g, rt_fin)) {
        const Geometry& g = h->g;
        const LexTileGeom lt = lex_tile_geom(g, rt_fin);
        const long long rows_pad = round_up(h->n_rows, kLexTileRows);
        h->lext_bytes = (size_t)rows_pad * g.S_pad * (lt.tcode_bytes + 2 * (size_t)g.G);
        if (cudaMalloc(&h->lext, h->lext_bytes) != cudaSuccess) { cudaGetLastError(); h->lext = nullptr; h->lext_bytes = 0; }   // tile path simply stays off
        if (h->lext) {
            const long long total = rows_pad * g.S_pad;
            const unsigned blocks = (unsigned)((total + 255) / 256);
            if (lt.wide)
                build_lext_kernel<uint16_t, uint16_t><<<blocks, 256>>>(h->n_rows, rows_pad, g.S_pad, g.G, 0xFFFFu, h->lexv, (const uint16_t*)h->lexi, h->lext);
            else if (g.code_bytes == 1)
                build_lext_kernel<uint8_t, uint8_t><<<blocks, 256>>>(h->n_rows, rows_pad, g.S_pad, g.G, (uint32_t)rt_fin, h->lexv, (const uint8_t*)h->lexi, h->lext);
            else
                build_lext_kernel<uint16_t, uint8_t><<<blocks, 256>>>(h->n_rows, rows_pad, g.S_pad, g.G, (uint32_t)rt_fin, h->lexv, (const uint16_t*)h->lexi, h->lext);
            DHR_CUDA(cudaGetLastError());
            DHR_CUDA(cudaDeviceSynchronize());
        }
    }
    if (h->g.C_pad > 0 && h->n_rows > 0 && dense_tile_ts_supported(h->g)) {
        const Geometry& g = h->g;
        const int nkb = (g.C_pad + kDenseTileCols - 1) / kDenseTileCols;
        const long long rows_pad = round_up(h->n_rows, kDenseTileRows);
        h->dnst_bytes = (size_t)rows_pad * nkb * kDenseTileCols * sizeof(__half);
        if (cudaMalloc(&h->dnst, h->dnst_bytes) != cudaSuccess) { cudaGetLastError(); h->dnst = nullptr; h->dnst_bytes = 0; }   // K2 then reads the row-major block
        if (h->dnst) {
            const long long total = rows_pad * nkb * (kDenseTileCols / 8);
            build_dnst_kernel<<<(unsigned)((total + 255) / 256), 256>>>(h->n_rows, rows_pad, g.C_pad, nkb, h->dns, h->dnst);
            DHR_CUDA(cudaGetLastError());
            DHR_CUDA(cudaDeviceSynchronize());
        }
    }
    // the row-major arrays only serve K1 / K4 / the overflow fallback: with complete tiled copies they are dropped (half the
    // HBM footprint) and rebuilt on first use, unless the index was created with DHR_INDEX_KEEP_ROWMAJOR
    if (!h->keep_rowmajor) DHR_TRY(drop_rowmajor_unchecked(h));
    // the ingest staging buffers are not needed any more
    if (h->stage_a) { cudaFree(h->stage_a); h->stage_a = nullptr; h->stage_a_bytes = 0; }
    if (h->stage_b) { cudaFree(h->stage_b); h->stage_b = nullptr; h->stage_b_bytes = 0; }
    h->finalized = true;
    return DHR_OK;
}

int dhr_index_open(dhr_index** out, int device, int64_t n_rows, int n_slices, int group, int n_dense, int val_dtype,
                   const void* vals, int64_t vstride, int idx_dtype, const void* idx, int64_t istride, int64_t row_offset,
                   unsigned flags) {
    if (!out) return DHR_ERR_INVALID;
    dhr_index* h = nullptr;
    DHR_TRY(dhr_index_create(&h, device, n_rows, n_slices, group, n_dense, idx_dtype, row_offset, flags));
    int s = dhr_index_append(h, n_rows, val_dtype, vals, vstride, idx_dtype, idx, istride);
    if (s == DHR_OK) s = dhr_index_finalize(h);
    if (s != DHR_OK) { dhr_index_close(h); return s; }
    *out = h;
    return DHR_OK;
}

int dhr_index_close(dhr_index* h) {
    if (!h) return DHR_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    void* bufs[] = {h->lexv, h->lexi, h->dns, h->dnst, h->lext, h->qblocks, h->qblock_bytes, h->lane[0].scratch, h->lane[1].scratch, h->lexp, h->lexp_nbytes, h->d_flags, h->stage_a, h->stage_b, h->q_lex16, h->q_lex32, h->q_dns16, h->q_dns32,
                    h->q_code, h->topk.tau, h->topk.cnt, h->topk.overflow, h->topk.cand_score, h->topk.cand_row,
                    h->topk1.tau, h->topk1.cnt, h->topk1.overflow, h->topk1.cand_score, h->topk1.cand_row,
                    h->topk.seg_score, h->topk.seg_row, h->topk.seg_cnt, h->topk1.seg_score, h->topk1.seg_row, h->topk1.seg_cnt,
                    h->d_out_scores, h->d_out_rows, h->d_out_counts, h->d_overflow, h->stage_c};
    for (void* b : bufs) if (b) cudaFree(b);
    for (cudaEvent_t e : h->batch_events) cudaEventDestroy(e);
    for (int L = 0; L < 2; ++L) {
        dhr_index::TileLane& ln = h->lane[L];
        cudaEvent_t evs[] = {ln.ev_k2_done[0], ln.ev_k2_done[1], ln.ev_k1_done[0], ln.ev_k1_done[1], ln.ev_fork, ln.ev_join, ln.ev_sel, ln.ev_done};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
        if (ln.main) cudaStreamDestroy(ln.main);
        if (ln.aux) cudaStreamDestroy(ln.aux);
        if (ln.aux2) cudaStreamDestroy(ln.aux2);
    }
    if (h->ev_lanes_fork) cudaEventDestroy(h->ev_lanes_fork);
    h->events.destroy();
    cudaGetLastError();
    delete h;
    return DHR_OK;
}

int dhr_index_rows(const dhr_index* h, int64_t* n_rows) {
    if (!h || !n_rows) return DHR_ERR_INVALID;
    *n_rows = h->n_rows;
    return DHR_OK;
}

int dhr_index_row_bytes(const dhr_index* h, int64_t* bytes) {
    if (!h || !bytes) return DHR_ERR_INVALID;
    *bytes = h->g.row_bytes();
    return DHR_OK;
}

int dhr_index_device_bytes(const dhr_index* h, int64_t* bytes) {
    if (!h || !bytes) return DHR_ERR_INVALID;
    const Geometry& g = h->g;
    const size_t rows = (size_t)(h->capacity > 0 ? h->capacity : 1);
    size_t b = 0;
    if (h->lexv) b += rows * g.D_pad * 2;
    if (h->lexi) b += rows * g.S_pad * g.code_bytes;
    if (h->dns) b += rows * g.C_pad * 2;
    b += h->lext_bytes + h->lexp_bytes + h->dnst_bytes + h->qblocks_bytes + h->qblock_bytes_cap + h->lane[0].scratch_bytes + h->lane[1].scratch_bytes + h->stage_a_bytes + h->stage_b_bytes +
         h->stage_c_bytes;
    if (h->topk.tau) b += (size_t)kMaxInflight * (12 + (size_t)kCandCap * 8);
    if (h->topk1.tau) b += (size_t)kMaxInflight * (12 + (size_t)kCandCap * 8);
    if (h->topk.seg_cnt) b += (size_t)kMaxInflight * kSegCount * (4 + (size_t)kSegCap * 8);
    if (h->topk1.seg_cnt) b += (size_t)kMaxInflight * kSegCount * (4 + (size_t)kSegCap * 8);
    if (h->q_capacity > 0) b += ((size_t)h->q_capacity + kMaxInflight) * ((size_t)g.D_pad * 6 + (size_t)g.S_pad * g.code_bytes + (size_t)g.C_pad * 6);
    b += h->out_capacity * 12 + h->out_q_capacity * 4 + h->overflow_capacity * 4;
    *bytes = (int64_t)b;
    return DHR_OK;
}

int dhr_index_set_option(dhr_index* h, const char* name, int64_t value) {
    if (!h || !name) return DHR_ERR_INVALID;
    if (!strcmp(name, "rowmajor")) {
        if (!h->finalized) return DHR_ERR_STATE;
        return value ? ensure_rowmajor(h) : drop_rowmajor(h);
    }
    if (!strcmp(name, "scan_variant")) { if (value < 0 || value > 1) return DHR_ERR_INVALID; h->opt_scan_variant = (int)value; return DHR_OK; }
    if (!strcmp(name, "query_block")) {
        if (value != 1 && value != 2 && value != 4 && value != 8) return DHR_ERR_INVALID;
        h->opt_query_block = (int)value; return DHR_OK;
    }
    if (!strcmp(name, "query_groups")) { if (value < 1 || value > kMaxScanInflight) return DHR_ERR_INVALID; h->opt_query_groups = (int)value; return DHR_OK; }
    if (!strcmp(name, "tile_mode")) { h->opt_tile_mode = value != 0; return DHR_OK; }
    if (!strcmp(name, "stream_priority")) { if (value < 0 || value > 1) return DHR_ERR_INVALID; h->opt_stream_priority = (int)value; return DHR_OK; }
    if (!strcmp(name, "lex_stages")) { if (value < 0 || value > 8 || value == 1) return DHR_ERR_INVALID; h->opt_lex_stages = (int)value; return DHR_OK; }
    if (!strcmp(name, "dense_lite")) { if (value < 0 || value > 1) return DHR_ERR_INVALID; h->opt_dense_lite = (int)value; return DHR_OK; }
    if (!strcmp(name, "dense_prefetch")) { if (value < 0 || value > 1) return DHR_ERR_INVALID; h->opt_dense_prefetch = (int)value; return DHR_OK; }
    if (!strcmp(name, "dense_multicast")) { if (value < 0 || value > 2) return DHR_ERR_INVALID; h->opt_dense_multicast = (int)value; return DHR_OK; }
    if (!strcmp(name, "lanes")) { if (value < 1 || value > 2) return DHR_ERR_INVALID; h->opt_lanes = (int)value; return DHR_OK; }
    if (!strcmp(name, "overlap")) { h->opt_overlap = value != 0; return DHR_OK; }
    if (!strcmp(name, "dense_variant")) { if (value < 0 || value > 3) return DHR_ERR_INVALID; h->opt_dense_variant = (int)value; return DHR_OK; }
    if (!strcmp(name, "profile")) { h->opt_profile = value != 0; return DHR_OK; }
    return DHR_ERR_INVALID;
}

int dhr_index_get_stats(const dhr_index* h, dhr_stats* out) {
    if (!h || !out) return DHR_ERR_INVALID;
    *out = h->stats;
    return DHR_OK;
}

}  // extern "C"
