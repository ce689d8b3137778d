// wrenc_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Plain C++17 restatement of the all-intra RD-search hot path of hjmkt/wrenc (Rust, read-only at
// /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or execute anything under oracle/.  The product (wrenc_b200/) never does.
//
// PARITY UNPINNED: the reference holds no golden vectors / known-answer tests for this path
//  and cannot be compiled here (no rustc/cargo).  The restatement is pinned only by
// (a) the source text it follows line by line (file:line cited at each function),
// (b) spec identities and the libm known-answer constants of SURVEY.md §5.9-H4 (tests/test_oracle_kat.py),
// (c) an independent decode of its own bitstream (oracle/wrenc_decode.cpp) reproducing its reconstruction.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace wo {

enum TreeType { SINGLE_TREE = 0, DUAL_TREE_LUMA = 1, DUAL_TREE_CHROMA = 2 };
enum { MODE_PLANAR = 0, MODE_DC = 1, MODE_LT_CCLM = 81, MODE_L_CCLM = 82, MODE_T_CCLM = 83 };  // common.rs:139-142

// Every tunable the hot path reads from ectx.extra_params, with the defaults that are live when
// dep_quant_used_flag=true and trellis=true (block_splitter.rs:21-44,187-375,594-693; quantizer.rs:16-19,650-683).
struct Tuning {
    double lv_pow_dq_trellis = 0.48592678233563835;
    double lv_offset_dq_trellis = 0.15150746310196822;
    float non_planar_offset = 2.2153597f;
    float mpm_idx_offset = 1.3660221f;
    float mpm_remainder_mult = 0.5007182f;
    float mpm_remainder_offset = 2.2973304f;
    float planar_offset = 0.9626864f;
    float header_bits = 1.1772872f;
    float chroma_header_bits = 1.309252f;
    float qp_div = 4.4043665f;
    float lambda_mul = 1.1282581f;
    float cclm_pow = 0.4587651f;
    float mpm_idx_pow = 0.40271285f;
    float mpm_remainder_pow = 0.34385094f;
    float cclm_mode_idx_offset = 2.1f;
    float non_cclm_offset = 0.89f;
    float cclm_offset = 0.53f;
    bool has_a = false;  // extra param "a": overrides the chroma lambda multiplier (block_splitter.rs:775-778)
    float a = 0.f;
    double quant_lv_pow = 0.5004010166085378;
    double quant_qp_div = 5.218413785332902;
    double quant_lambda_mul = 1.2709404305806742;
    int64_t quant_lambda_offset = 11;
    // parse "k=v,k=v" (main.rs:202-216).  Returns false on malformed input.
    bool parse(const char *extra_params, std::string *err);
};

// Constants derived once per (qp, tuning) on the host with libm pow/powf.
struct Consts {
    int qp = 32;
    int64_t lv[1024];       // lv_dq_trellis_table  block_splitter.rs:45-53
    int64_t dq[1024];       // Quantizer::dq_table  quantizer.rs:15-26
    int64_t lambda_q = 0;   // quantizer.rs:683
    float lambda_rd = 0;    // block_splitter.rs:472
    float lambda_rd_c = 0;  // block_splitter.rs:775-778
    int32_t ls = 0;         // quantizer.rs:617-622 (uniform, m=16)
    // header-bit tables, already "* 16384.0) as i64" (block_splitter.rs:377-406, 695-712)
    // luma kind: 0 = planar, 1..5 = mpm idx 0..4, 6..66 = mpm remainder 0..60
    // cclm kind: 0 = not CCLM, 1..3 = cclm_mode_idx 0..2
    int64_t hdr_single[67][4];
    int64_t hdr_dual_luma[67];
    int64_t hdr_chroma[4];
    void init(int qp, const Tuning &t);
};

struct Plane {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    void alloc(int w_, int h_) { w = w_; h = h_; d.assign((size_t)w * h, 0); }
    uint8_t &at(int x, int y) { return d[(size_t)y * w + x]; }
    uint8_t at(int x, int y) const { return d[(size_t)y * w + x]; }
};

// Per-CTU decision record — the flat form both the oracle and the CUDA path emit for parity checks.
struct CtuRecord {
    uint32_t split_mask;      // bit 0: 32x32 split; bits 1..4: 16x16 #i split (z-order); bits 5..20: 8x8 #i split
    uint8_t luma_mode[64];    // luma intra mode of the CU covering each 4x4 block, raster 8x8 grid
    uint8_t chroma_mode[16];  // chroma pred mode actually used (luma-derived DM value or 81..83) per 8x8 luma block, raster 4x4 grid
    float cost;               // RD cost returned by split_ct for the CTU root
};

struct Picture {
    int W = 0, H = 0;
    Plane orig[3], rec[3];
    std::vector<int16_t> coef[3];  // final quantised levels, TB-local raster stored at the TB's position
    std::vector<uint8_t> mode_map;  // final luma mode per 4x4 block (W/4 x H/4), filled CTU by CTU
    std::vector<CtuRecord> records;
    void init(int W_, int H_, const uint8_t *y, const uint8_t *cb, const uint8_t *cr);
};

// A transform unit as the predictor sees it (ctu.rs:324-367): luma geometry + tree type + the two
// tree-position availability flags (ctu.rs:2083-2188) + the TU's copy of the CU modes.
struct TU {
    int x, y, w;  // luma position / size (square)
    int tree;
    bool ar, bl;  // is_above_right_available / is_below_left_available
    int mode[3];  // tu.cu_intra_pred_mode
};

// ---- block ops (exported for per-block parity tests) ----
void build_refs(const Picture &p, const TU &tu, int c, int mode, int16_t *left /*2N+1*/, int16_t *above /*2N*/,
                int16_t *leftF, int16_t *aboveF);
void predict(const Picture &p, const TU &tu, int c, uint8_t *pred /*N*N*/);
void fwd_dct(const int16_t *res, int log2n, int16_t *coef);
void inv_dct(const int16_t *deq, int log2n, int16_t *out);
void quantize_dq(const Consts &k, const int16_t *coef, int log2n, int16_t *q);
void dequantize(const Consts &k, const int16_t *q, int log2n, int16_t *d);
int64_t rate_levels(const Consts &k, const int16_t *q, int log2n);
const int16_t *dct_matrix(int log2n);  // N x N, row-major [i][x]
const uint16_t *scan_order(int log2n); // coding-order-reversed? no: forward diag scan: idx -> (y<<8|x)

// ---- the search ----
struct Encoder {
    Consts k;
    int max_depth = 3;
    Picture *pic = nullptr;
    // statistics (for the bench normaliser cross-check)
    uint64_t n_pipelines = 0, n_predictions = 0;
    // search one picture completely (all CTUs, raster order); fills pic->rec, coef, records, mode_map
    void search_picture(Picture &p);
    // search a single CTU (ctu-aligned luma x,y); previous CTUs must be final
    void search_ctu(Picture &p, int cx, int cy);

  private:
    friend struct SearchImpl;
};

// CABAC-coded slice_data() of a searched picture (phase 2; wrenc_oracle_cabac.cpp)
std::vector<uint8_t> code_slice_data(const Consts &k, Picture &p);
// test hooks (wrenc_oracle_cabac.cpp): bin string in the product's entry format; the reference engine over any bin string
std::vector<uint8_t> code_slice_data_traced(const Consts &k, Picture &p, std::vector<uint16_t> &bins);
std::vector<uint8_t> code_bin_string(int slice_qp, const uint16_t *entries, size_t n);

}  // namespace wo
