// internal.h -- host-side structures shared by the translation units of libdhr_b200.so
#pragma once
#include <vector>

#include "common.cuh"

namespace dhr {

constexpr int kCandCap = 16384;          // candidate slots per in-flight query (fits a 128 KiB smem sort)
constexpr int kMaxInflight = 256;        // query slots per super-batch
constexpr int kMaxScanInflight = 64;     // row-scan path: query_block * query_groups

struct Geometry {
    int S = 0, G = 1, C = 0;             // slices, values per slice, dense columns (user shape)
    int S_pad = 0, D_pad = 0, C_pad = 0; // padded strides (elements) of the resident arrays
    int code_bytes = 1;                  // 1 = uint8 codes, 2 = uint16 codes
    int unit_halves = 8;                 // lcm(G, 8): fp16 values one lane handles per lexical unit
    int unit_slices = 8;                 // unit_halves / G
    int n_units = 0;                     // S_pad / unit_slices
    int n_chunks = 0;                    // C_pad / 8 (16-byte dense chunks)
    int64_t row_bytes() const { return (int64_t)D_pad * 2 + (int64_t)S_pad * code_bytes + (int64_t)C_pad * 2; }
};

// per-query selection state of one in-flight slot
struct TopkState {
    float*    tau = nullptr;             // [slots] strict admission threshold
    uint32_t* cnt = nullptr;             // [slots] candidates appended (may exceed cap -> overflow)
    uint32_t* overflow = nullptr;        // [slots] sticky overflow flag
    float*    cand_score = nullptr;      // [slots][cap]
    int32_t*  cand_row = nullptr;        // [slots][cap]
    // Segmented candidate lists of the tensor-core filter epilogue (K2, thread = query): every CTA appends to its OWN segment of a
    // slot with a register counter -- no atomics, no round trips -- and K3 gathers the segments behind the kept candidates.
    float*    seg_score = nullptr;       // [slots][kSegCount][kSegCap]
    int32_t*  seg_row = nullptr;
    uint32_t* seg_cnt = nullptr;         // [slots][kSegCount] appended per segment (may exceed kSegCap -> overflow); zeroed before a scan launch
};
constexpr int kSegCount = 320;           // >= 2 x CTAs that can work on one query group in one launch (two epilogue warps per TMEM lane quadrant, each its own writer); multiple of 32
constexpr int kSegCap = 256;             // per (slot, CTA) and chunk: the first chunk hands a CTA <= 2 tiles of 128 rows, later chunks ~100 rows

struct EventPool {
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    cudaEvent_t get();
    void reset() { used = 0; }
    void destroy();
};

}  // namespace dhr

struct dhr_index {
    int device = 0;
    int64_t capacity = 0, n_rows = 0, row_offset = 0;
    dhr::Geometry g;
    int idx_dtype = DHR_IDX_NONE;
    bool finalized = false;
    bool use_postings = false;           // DHR_INDEX_LEX_POSTINGS: lexical copy as postings (K1p, experimental) instead of the tiled layout (K1t)
    bool keep_rowmajor = false;          // DHR_INDEX_KEEP_ROWMAJOR: never drop lexv / lexi / dns at finalize
    // resident arrays (device)
    __half* lexv = nullptr;              // [capacity][D_pad]
    uint8_t* lexi = nullptr;             // [capacity][S_pad] codes (uint8 or uint16)
    __half* dns = nullptr;               // [capacity][C_pad]
    __half* dnst = nullptr;              // K-blocked copy of the dense block for K2 (TS): [tile of 128 rows][k-block of 64 cols][128][64], built at finalize
    size_t dnst_bytes = 0;
    uint8_t* lext = nullptr;             // tiled lexical copy for K1t: [tile of 512 rows][4-slice chunk]{codes u8|u16 [512][4] | vals [4][512][G]}
    size_t lext_bytes = 0;
    uint8_t* lexp = nullptr;             // postings copy for K1p: [tile][chunk] fixed-stride blocks (see lex_post_geom)
    size_t lexp_bytes = 0;
    uint32_t* lexp_nbytes = nullptr;     // [tiles * chunks] bytes actually used by each block (what the producer copies)
    int lex_layout = 0;                  // 0 = tiled copy (K1t), 1 = postings (K1p)
    int max_code = -1;                   // largest slice code stored (known after finalize)
    int* d_flags = nullptr;              // [4] device-side validation flags (lossy, idx range, query needs fp32, spare)
    // staging for host -> device appends / queries
    void* stage_a = nullptr; size_t stage_a_bytes = 0;
    void* stage_b = nullptr; size_t stage_b_bytes = 0;
    // query workspace
    void* q_lex16 = nullptr; void* q_lex32 = nullptr; void* q_dns16 = nullptr; void* q_dns32 = nullptr; void* q_code = nullptr;
    int q_capacity = 0;
    dhr::TopkState topk;
    // tile-mode workspace
    uint8_t* qblocks = nullptr; size_t qblocks_bytes = 0;
    uint32_t* qblock_bytes = nullptr; size_t qblock_bytes_cap = 0;
    // Two batch lanes: consecutive 256-query batches alternate between two independent sets of selection state, scratch, streams
    // and events, so the select / drain at a chunk boundary of one batch is covered by scan launches of the other.
    struct TileLane {
        cudaStream_t main = nullptr;         // lane 1 only (lane 0 runs on the caller's stream)
        cudaStream_t aux = nullptr;          // K2 launches of the hybrid tile path
        cudaStream_t aux2 = nullptr;         // every second K1t launch of a chunk (launches of one chunk are independent)
        float* scratch = nullptr; size_t scratch_bytes = 0;   // two sub-chunk buffers (K2 of sub-chunk i+1 overlaps K1t of sub-chunk i)
        cudaEvent_t ev_k2_done[2] = {nullptr, nullptr}, ev_k1_done[2] = {nullptr, nullptr}, ev_fork = nullptr, ev_join = nullptr, ev_sel = nullptr;
        cudaEvent_t ev_done = nullptr;       // lane 1: all its batches enqueued so far have finished
    };
    TileLane lane[2];
    dhr::TopkState topk1;                    // selection state of lane 1 (lane 0 uses `topk`)
    cudaEvent_t ev_lanes_fork = nullptr;
    int opt_lanes = 2;                       // 1 = batches strictly one after the other
    float* d_out_scores = nullptr; int64_t* d_out_rows = nullptr; int32_t* d_out_counts = nullptr;
    size_t out_capacity = 0, out_q_capacity = 0;                   // [Q,k] elements / [Q] counts of the host-output staging
    uint32_t* d_overflow = nullptr; size_t overflow_capacity = 0;   // [Q] per-query overflow flags of the current search
    void* stage_c = nullptr; size_t stage_c_bytes = 0;              // rerank candidate staging
    // stream-ordered search (dhr_search_keys): per-batch completion events and what dhr_search_complete needs
    std::vector<cudaEvent_t> batch_events;
    int batch_size = 0, n_batches = 0;
    struct { bool active = false; int n_queries = 0, k = 0; bool masked = true, f32 = false; uint64_t* out_keys = nullptr; } pending;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // options
    int opt_scan_variant = 1;            // TMA bulk staging (measured faster than direct loads at QB=1 and QB=8)
    int opt_query_block = 8;
    int opt_query_groups = 8;
    int opt_profile = 0;
    int opt_tile_mode = 1;               // use the tensor-core tile kernels when the shape allows
    int opt_overlap = 1;                 // hybrid tile path: run K2 on a second stream, one sub-chunk ahead of K1t
    int opt_stream_priority = 0;         // 1: lane streams that carry K1t are created with the highest priority (set before the first search)
    int opt_lex_stages = 0;              // K1t ring depth cap (0 = as many stages as fit)
    int opt_dense_lite = 0;              // hybrid path: 1 = K2 in its small-footprint form that co-resides with a K1t CTA on an SM (measured slower: 9.1-9.7 k vs 10.3 k q/s); 0 = the big kernel, alternating
    int lite_stages = 0;                 // ring depth the lite K2 may use beside the current K1t geometry (set per batch by the hybrid runner; 0 = does not fit)
    int opt_dense_prefetch = 0;          // K2: TMA L2 prefetch two tiles ahead of the demand loads (measured slower on B200 once the ring holds a whole tile: 0)
    int opt_dense_multicast = 1;         // K2 (TS): the two query groups of a batch share corpus tiles as a cluster of two CTAs (0 off, 1 dense-only searches, 2 always)
    int opt_dense_variant = 3;           // K2: 0 = both operands in shared memory (SS), 1 = queries in TMEM (TS), 2 = TS as a CTA pair (cta_group::2, M = 256), 3 = auto (2 for filter-mode launches, 1 for scratch-mode ones)
    int num_sms = 148;
    long long smem_per_sm = 233472;      // shared memory of one SM (B200: 228 KiB)
    dhr_stats stats{};
    dhr::EventPool events;
};

namespace dhr {

// scan launch description (one chunk of rows, one super-batch of queries)
struct ScanArgs {
    const __half* lexv; const uint8_t* lexi; const __half* dns;
    int S_pad, D_pad, C_pad, n_units, n_chunks;
    long long row_begin, row_end;
    const void* q_lex; const void* q_code; const void* q_dns;   // rows of the first query of the super-batch
    int n_queries;                       // valid queries in the super-batch
    int n_groups;                        // ceil(n_queries / QB)
    int masked;
    float* tau; uint32_t* cnt; float* cand_score; int32_t* cand_row; int cap;
    int rows_per_cta;
    int tile_rows;                       // TMA variant: rows per smem stage
    int n_stages;
};

int launch_scan(const dhr_index* h, const ScanArgs& a, int query_block, bool q_f32, int variant, cudaStream_t st);
size_t scan_tma_smem_bytes(const Geometry& g, int query_block, bool q_f32, int tile_rows, int n_stages);

// where the final pass of a batch writes: either (scores, rows, counts) or packed 64-bit keys (sharded exchange format)
struct SelectOut {
    float* scores = nullptr; int64_t* rows = nullptr; int32_t* counts = nullptr;
    uint64_t* keys = nullptr;            // [Q][k] (ordered score << 32) | (0xFFFFFFFF - global row), 0 = padding
    uint32_t* overflow = nullptr;        // [Q] set to 1 for queries whose candidate buffer overflowed in any chunk
    int base = 0;                        // first query of the batch
    int64_t row_offset = 0;
};
int launch_select(const TopkState& t, int n_slots, int k, int cap, bool final_pass, const SelectOut& o, cudaStream_t st);
void launch_init_slots(const TopkState& t, cudaStream_t st);

int launch_prep_queries(dhr_index* h, int n, int val_dtype, const void* vals, int64_t vstride, int idx_dtype,
                        const void* idx, int64_t istride, float lamda, cudaStream_t st);

int launch_rerank(const dhr_index* h, const ScanArgs& a, bool q_f32, const long long* d_cand, int n_cand, cudaStream_t st);

struct LexTileGeom {
    int G, code_bytes, n_chunks, rt;
    int wide, tcode_bytes, n_stages;     // bucket-lookup layout (lex_tile.cu), code width of the tiled copy, smem ring depth
    int pblock_bytes, qoff_bytes, qblock_stride, stage_bytes;
};
constexpr int kLexTileRows = 512;        // passages per K1t tile (= consumer threads)
constexpr int kLexTileQueries = 64;      // queries per K1t tile (acc[64][512] fp32 = 128 KiB, 16 consumer warps per SM)
constexpr int kLexTileSlices = 4;        // slices per chunk
LexTileGeom lex_tile_geom(const Geometry& g, int rt);
size_t lex_tile_cta_footprint(const LexTileGeom& t, int n_stages);   // shared memory one K1t CTA takes from its SM (dynamic + static + driver reserve)
// postings layout of the corpus tile (K1p, lex_tile.cu): per (tile of 512 rows, chunk of 4 slices) a fixed-stride block holding a
// 16-byte header and the NON-EMPTY passages of each slice sorted by code, item = {passage | code << 16, G fp16}
LexTileGeom lex_post_geom(const Geometry& g, int rt);
bool lex_post_supported(const Geometry& g, int rt);
int lex_post_entry_words(int G);
int launch_lex_post(const dhr_index* h, const LexTileGeom& t, const uint8_t* qblocks, const uint32_t* qblock_bytes, int n_queries,
                    long long row_begin, long long row_end, const float* scratch, long long scratch_slots, long long scratch_row0,
                    const TopkState& tk, int cap, cudaStream_t st);
bool lex_tile_supported(const Geometry& g, int rt);
int launch_lex_tile_prep(const dhr_index* h, const LexTileGeom& t, const void* q_lex16, const void* q_code, int n_queries,
                         uint8_t* qblocks, uint32_t* qblock_bytes, cudaStream_t st);
int launch_lex_tile(const dhr_index* h, const LexTileGeom& t, const uint8_t* qblocks, const uint32_t* qblock_bytes, int n_queries,
                    long long row_begin, long long row_end, const float* scratch, long long scratch_slots, long long scratch_row0,
                    const TopkState& tk, int cap, cudaStream_t st);

bool dense_tile_supported(const Geometry& g, int* n_stages_out);
bool dense_tile_ts_supported(const Geometry& g);
constexpr int kDenseTileRows = 128;      // rows per K2 (TS) corpus tile and per block of the K-blocked copy
constexpr int kDenseTileCols = 64;       // fp16 columns per k-block (128 bytes)
int launch_dense_tile(const dhr_index* h, const void* q_dns16, int n_queries, long long tile_row0, long long row_begin,
                      long long row_end, int mode, float* scratch, long long scratch_slots, const TopkState& t, int cap,
                      cudaStream_t st);

// one column pass of the tensor-core kernel over any fp16 column block (mode: 0 filter, 1 scratch =, 2 scratch +=, 3 (+ scratch) -> filter)
int launch_dense_pass(const dhr_index* h, const __half* blocked, const __half* rowmajor, int c_pitch, int cols, const void* q16,
                      int q_pitch, int n_queries, long long tile_row0, long long row_begin, long long row_end, int mode, float* scratch,
                      long long scratch_slots, const TopkState& t, int cap, cudaStream_t st);
constexpr int kDensePassMaxCols = 768;   // query operand of one pass must fit 384 TMEM columns

int ensure_device_buffer(void** p, size_t* cur, size_t need);
// row-major arrays (lexv / lexi / dns) serve the row scan K1, the rerank kernel K4 and the overflow fallback; they can be
// dropped once the tiled copies exist (option "rowmajor" = 0) and are rebuilt from the tiled copies on first use
int ensure_rowmajor(dhr_index* h);
int drop_rowmajor(dhr_index* h);
bool is_device_pointer(const void* p);

}  // namespace dhr
