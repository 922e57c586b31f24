// densify.cu -- next-row n3: the neural densify op of castorini/dhr tevatron/DHR/utils.py:5-22 fused with
// the storage conversion of tevatron/driver/encode.py:155-170,180-195.
//   lexical_reps [B, V] -> drop the first `remove_dims` vocabulary ids -> view(B, R, dims) -> max over R
//   value  = max_r x[b, remove + r*dims + d]   (stored fp16, like value_encoded)
//   index  = argmax_r (first maximum)          (stored uint8, like index_encoded)
// One thread per (b, d): the R strided reads of neighbouring d are coalesced; the op is HBM-bound (reads B*V values once).
#include "internal.h"

namespace dhr {

template <typename T>
__global__ void densify_kernel(int batch, int dims, int R, int remove_dims, const T* __restrict__ reps, long long stride,
                               __half* __restrict__ out_vals, long long val_stride, uint8_t* __restrict__ out_idx, long long idx_stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)batch * dims) return;
    const int b = (int)(i / dims), d = (int)(i % dims);
    const T* row = reps + (size_t)b * stride + remove_dims + d;
    float best = -INFINITY;
    int arg = 0;
    for (int r = 0; r < R; ++r) {
        float v;
        if constexpr (sizeof(T) == 2) v = __half2float(row[(size_t)r * dims]); else v = row[(size_t)r * dims];
        if (v > best || r == 0) { if (v > best || r == 0) { best = v; arg = r; } }
    }
    out_vals[(size_t)b * val_stride + d] = __float2half_rn(best);
    out_idx[(size_t)b * idx_stride + d] = (uint8_t)arg;
}

}  // namespace dhr

using namespace dhr;

extern "C" int dhr_densify(int device, int batch, int vocab, int dims, int remove_dims, int val_dtype, const void* reps,
                           int64_t reps_row_stride, void* out_vals_f16, int64_t out_val_row_stride, uint8_t* out_idx,
                           int64_t out_idx_row_stride, void* stream) {
    if (batch < 0 || dims <= 0 || remove_dims < 0 || vocab <= remove_dims || !reps || !out_vals_f16 || !out_idx) return DHR_ERR_INVALID;
    if ((vocab - remove_dims) % dims != 0) return DHR_ERR_INVALID;          // utils.py:15-16
    const int R = (vocab - remove_dims) / dims;
    if (R > 256) return DHR_ERR_UNSUPPORTED;                                 // index is stored as uint8 (encode.py:157,166)
    if (reps_row_stride < vocab || out_val_row_stride < dims || out_idx_row_stride < dims) return DHR_ERR_INVALID;
    if (val_dtype != DHR_VAL_F16 && val_dtype != DHR_VAL_F32) return DHR_ERR_INVALID;
    if (!is_device_pointer(reps) || !is_device_pointer(out_vals_f16) || !is_device_pointer(out_idx)) return DHR_ERR_INVALID;
    if (batch == 0) return DHR_OK;
    DHR_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)batch * dims;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (val_dtype == DHR_VAL_F16)
        densify_kernel<__half><<<blocks, 256, 0, st>>>(batch, dims, R, remove_dims, (const __half*)reps, reps_row_stride,
                                                       (__half*)out_vals_f16, out_val_row_stride, out_idx, out_idx_row_stride);
    else
        densify_kernel<float><<<blocks, 256, 0, st>>>(batch, dims, R, remove_dims, (const float*)reps, reps_row_stride,
                                                      (__half*)out_vals_f16, out_val_row_stride, out_idx, out_idx_row_stride);
    DHR_CUDA(cudaGetLastError());
    return DHR_OK;
}
