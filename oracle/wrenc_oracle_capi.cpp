// C API over the CPU ORACLE for ctypes (tests/, smoke(), bench.py cpu_baseline only — NOT product code).
#include <algorithm>
#include <cstring>
#include <string>

#include "wrenc_oracle.hpp"

using namespace wo;

extern "C" {

struct wo_handle {
    Tuning t;
    Encoder enc;
    std::string err;
};

wo_handle *wo_create(int qp, int max_depth, const char *extra_params) {
    wo_handle *h = new wo_handle();
    if (!h->t.parse(extra_params, &h->err)) {
        delete h;
        return nullptr;
    }
    h->enc.k.init(qp, h->t);
    h->enc.max_depth = max_depth;
    return h;
}
void wo_destroy(wo_handle *h) { delete h; }

// constants for known-answer tests
void wo_consts(wo_handle *h, int64_t *lv8, int64_t *dq8, int64_t *lambda_q, float *lambda_rd, int32_t *ls) {
    for (int i = 0; i < 8; i++) {
        lv8[i] = h->enc.k.lv[i];
        dq8[i] = h->enc.k.dq[i];
    }
    *lambda_q = h->enc.k.lambda_q;
    *lambda_rd = h->enc.k.lambda_rd;
    *ls = h->enc.k.ls;
}
void wo_hdr_tables(wo_handle *h, int64_t *single /*67*4*/, int64_t *dual /*67*/, int64_t *chroma /*4*/) {
    memcpy(single, h->enc.k.hdr_single, sizeof(h->enc.k.hdr_single));
    memcpy(dual, h->enc.k.hdr_dual_luma, sizeof(h->enc.k.hdr_dual_luma));
    memcpy(chroma, h->enc.k.hdr_chroma, sizeof(h->enc.k.hdr_chroma));
}
void wo_dct_matrix(int log2n, int16_t *out) { memcpy(out, dct_matrix(log2n), sizeof(int16_t) << (2 * log2n)); }
void wo_scan(int log2n, uint16_t *out) { memcpy(out, scan_order(log2n), sizeof(uint16_t) << (2 * log2n)); }

// Search (+ optional CABAC) of one picture.  Planes are tightly packed I420.
// records: W/32*H/32 CtuRecord (88 bytes each); coef_*: int16 planes; slice_data may be NULL (search only).
long wo_encode_picture(wo_handle *h, int W, int H, const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint8_t *rec_y,
                       uint8_t *rec_cb, uint8_t *rec_cr, int16_t *coef_y, int16_t *coef_cb, int16_t *coef_cr, void *records,
                       uint8_t *slice_data, size_t cap) {
    Picture p;
    p.init(W, H, y, cb, cr);
    h->enc.search_picture(p);
    uint8_t *rec[3] = {rec_y, rec_cb, rec_cr};
    int16_t *coef[3] = {coef_y, coef_cb, coef_cr};
    for (int c = 0; c < 3; c++) {
        if (rec[c]) memcpy(rec[c], p.rec[c].d.data(), p.rec[c].d.size());
        if (coef[c]) memcpy(coef[c], p.coef[c].data(), p.coef[c].size() * sizeof(int16_t));
    }
    if (records) memcpy(records, p.records.data(), p.records.size() * sizeof(CtuRecord));
    if (slice_data) {
        std::vector<uint8_t> sd = code_slice_data(h->enc.k, p);
        if (sd.size() > cap) return -1;
        memcpy(slice_data, sd.data(), sd.size());
        return (long)sd.size();
    }
    return 0;
}

uint64_t wo_num_pipelines(wo_handle *h) { return h->enc.n_pipelines; }
uint64_t wo_num_predictions(wo_handle *h) { return h->enc.n_predictions; }

// ---- per-block entry points (parity tests of the per-block kernels) ----
// Predict one component of a TU from a picture's reconstruction planes.
void wo_predict(int W, int H, const uint8_t *rec_y, const uint8_t *rec_cb, const uint8_t *rec_cr, int x, int y, int w, int tree, int ar,
                int bl, int c, int mode, uint8_t *pred) {
    Picture p;
    p.W = W;
    p.H = H;
    const uint8_t *src[3] = {rec_y, rec_cb, rec_cr};
    for (int i = 0; i < 3; i++) {
        int pw = i ? W / 2 : W, ph = i ? H / 2 : H;
        p.rec[i].alloc(pw, ph);
        memcpy(p.rec[i].d.data(), src[i], (size_t)pw * ph);
    }
    TU tu{x, y, w, tree, ar != 0, bl != 0, {mode, mode, mode}};
    predict(p, tu, c, pred);
}
void wo_fwd_dct(const int16_t *res, int log2n, int16_t *coef) { fwd_dct(res, log2n, coef); }
void wo_inv_dct(const int16_t *deq, int log2n, int16_t *out) { inv_dct(deq, log2n, out); }
void wo_quantize(wo_handle *h, const int16_t *coef, int log2n, int16_t *q) { quantize_dq(h->enc.k, coef, log2n, q); }
void wo_dequantize(wo_handle *h, const int16_t *q, int log2n, int16_t *d) { dequantize(h->enc.k, q, log2n, d); }
int64_t wo_rate(wo_handle *h, const int16_t *q, int log2n) { return rate_levels(h->enc.k, q, log2n); }
}

// ---- test hook: the bin string of a picture in the product's entry format (tests of the product's arithmetic coder) ----
extern "C" long wo_trace_bins(wo_handle *h, int W, int H, const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint16_t *bins, size_t cap,
                              uint8_t *slice_data, size_t sd_cap, long *sd_len) {
    Picture p;
    p.init(W, H, y, cb, cr);
    h->enc.search_picture(p);
    std::vector<uint16_t> b;
    std::vector<uint8_t> sd = code_slice_data_traced(h->enc.k, p, b);
    if (sd_len) *sd_len = (long)sd.size();
    if (slice_data && sd.size() <= sd_cap) memcpy(slice_data, sd.data(), sd.size());
    if (bins) memcpy(bins, b.data(), std::min(cap, b.size()) * sizeof(uint16_t));
    return (long)b.size();
}
