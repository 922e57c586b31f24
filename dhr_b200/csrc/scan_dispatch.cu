// scan_dispatch.cu -- runtime dispatch of K1 over G (values per slice); kernels live in scan_inst.cu.
#include "internal.h"

namespace dhr {

template <int G> int scan_entry(const dhr_index* h, const ScanArgs& a, int qb, bool q_f32, int variant, cudaStream_t st);
template <int G> int rerank_entry(const dhr_index* h, const ScanArgs& a, bool q_f32, const long long* d_cand, int n_cand, cudaStream_t st);

size_t scan_tma_smem_bytes(const Geometry& g, int query_block, bool q_f32, int tile_rows, int n_stages) {
    const size_t qt = q_f32 ? 4 : 2;
    size_t qbytes = (size_t)query_block * ((size_t)g.D_pad * qt + (size_t)g.S_pad * g.code_bytes + (size_t)g.C_pad * qt);
    qbytes = (qbytes + 127) / 128 * 128;
    auto r128 = [](size_t x) { return (x + 127) / 128 * 128; };
    const size_t stage = r128((size_t)tile_rows * g.D_pad * 2) + r128((size_t)tile_rows * g.S_pad * g.code_bytes) +
                         r128((size_t)tile_rows * g.C_pad * 2);
    return qbytes + stage * n_stages;
}


int launch_scan(const dhr_index* h, const ScanArgs& a, int query_block, bool q_f32, int variant, cudaStream_t st) {
    switch (h->g.G) {
        case 1: return scan_entry<1>(h, a, query_block, q_f32, variant, st);
        case 2: return scan_entry<2>(h, a, query_block, q_f32, variant, st);
        case 3: return scan_entry<3>(h, a, query_block, q_f32, variant, st);
        case 4: return scan_entry<4>(h, a, query_block, q_f32, variant, st);
        case 5: return scan_entry<5>(h, a, query_block, q_f32, variant, st);
        case 6: return scan_entry<6>(h, a, query_block, q_f32, variant, st);
        case 7: return scan_entry<7>(h, a, query_block, q_f32, variant, st);
        case 8: return scan_entry<8>(h, a, query_block, q_f32, variant, st);
        default: return DHR_ERR_UNSUPPORTED;
    }
}

int launch_rerank(const dhr_index* h, const ScanArgs& a, bool q_f32, const long long* d_cand, int n_cand, cudaStream_t st) {
    if (a.n_queries <= 0 || n_cand <= 0) return DHR_OK;
    switch (h->g.G) {
        case 1: return rerank_entry<1>(h, a, q_f32, d_cand, n_cand, st);
        case 2: return rerank_entry<2>(h, a, q_f32, d_cand, n_cand, st);
        case 3: return rerank_entry<3>(h, a, q_f32, d_cand, n_cand, st);
        case 4: return rerank_entry<4>(h, a, q_f32, d_cand, n_cand, st);
        case 5: return rerank_entry<5>(h, a, q_f32, d_cand, n_cand, st);
        case 6: return rerank_entry<6>(h, a, q_f32, d_cand, n_cand, st);
        case 7: return rerank_entry<7>(h, a, q_f32, d_cand, n_cand, st);
        case 8: return rerank_entry<8>(h, a, q_f32, d_cand, n_cand, st);
        default: return DHR_ERR_UNSUPPORTED;
    }
}

}  // namespace dhr
