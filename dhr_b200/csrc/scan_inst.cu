// scan_inst.cu -- instantiates K1 (scan + rerank) for one value of G; built once per G in parallel.
#include "scan_rows.cuh"

#ifndef DHR_G
#error "compile with -DDHR_G=<values per slice>"
#endif

namespace dhr {

template <>
int scan_entry<DHR_G>(const dhr_index* h, const ScanArgs& a, int qb, bool q_f32, int variant, cudaStream_t st) {
    if (h->g.code_bytes == 1) return launch_qb<DHR_G, uint8_t>(h, a, qb, q_f32, variant, st);
    return launch_qb<DHR_G, uint16_t>(h, a, qb, q_f32, variant, st);
}

template <>
int rerank_entry<DHR_G>(const dhr_index* h, const ScanArgs& a, bool q_f32, const long long* d_cand, int n_cand, cudaStream_t st) {
    if (h->g.code_bytes == 1) return launch_rerank_g<DHR_G, uint8_t>(a, q_f32, d_cand, n_cand, h->n_rows, st);
    return launch_rerank_g<DHR_G, uint16_t>(a, q_f32, d_cand, n_cand, h->n_rows, st);
}

}  // namespace dhr
