// wrenc_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code; see wrenc_oracle.hpp header).
// Literal, sequential restatement of wrenc's search path.  Every function cites the reference lines it follows
// (paths relative to /root/reference/src).  PARITY UNPINNED (no reference golden vectors exist; see header).
#include "wrenc_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace wo {

// ------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------

// VVC/HEVC 32-point DCT-II magnitudes c[j] = round-ish(64*sqrt(2)*cos(j*pi/64)), j=1..31; c[16]=64, c[32]=0.
// These are the even rows of the reference's 64-point table (transformer.rs:934-1234, sub-sampled at 1212-1221);
// tools/check_dct_structure.py verifies the identity against the reference text.
static const int kCos32[33] = {91, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                               61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};

static int16_t g_T[6][32 * 32];
static uint16_t g_scan[6][1024];  // forward scan index k = sb*16+pos -> (y<<8)|x   (ctu.rs:14-81)
static bool g_init = false;

static void diag_order(int bw, int bh, std::vector<std::pair<int, int>> &out) {
    // ctu.rs:53-77: up-right diagonal scan
    out.assign((size_t)bw * bh, {0, 0});
    int i = 0, x = 0, y = 0;
    bool stop = false;
    while (!stop) {
        while (y >= 0) {
            if (x < bw && y < bh) {
                out[i] = {x, y};
                i++;
            }
            y--;
            x++;
        }
        y = x;
        x = 0;
        if (i >= bw * bh) stop = true;
    }
}

static void init_tables() {
    if (g_init) return;
    for (int l = 2; l <= 5; l++) {
        int n = 1 << l;
        for (int i = 0; i < n; i++)
            for (int x = 0; x < n; x++) {
                int v;
                if (i == 0) {
                    v = 64;
                } else {
                    int m = (i * (2 * x + 1) * (32 / n)) % 128;  // angle in units of pi/64
                    int s = 1;
                    if (m > 64) m = 128 - m;
                    if (m > 32) {
                        m = 64 - m;
                        s = -1;
                    }
                    v = s * kCos32[m];
                }
                g_T[l][i * n + x] = (int16_t)v;
            }
        std::vector<std::pair<int, int>> co, so;
        diag_order(4, 4, co);
        diag_order(n / 4, n / 4, so);
        for (int sb = 0; sb < (n / 4) * (n / 4); sb++)
            for (int p = 0; p < 16; p++) {
                int x = so[sb].first * 4 + co[p].first;
                int y = so[sb].second * 4 + co[p].second;
                g_scan[l][sb * 16 + p] = (uint16_t)((y << 8) | x);
            }
    }
    g_init = true;
}

const int16_t *dct_matrix(int log2n) {
    init_tables();
    return g_T[log2n];
}
const uint16_t *scan_order(int log2n) {
    init_tables();
    return g_scan[log2n];
}

// common.rs:145-151 (entries 14+2 .. 14+66)
static const int kAngle[67] = {0,   0,   32,  29,  26,  23,  20,  18,  16,  14,  12,  10,  8,   6,   4,   3,   2,
                               1,   0,   -1,  -2,  -3,  -4,  -6,  -8,  -10, -12, -14, -16, -18, -20, -23, -26, -29,
                               -32, -29, -26, -23, -20, -18, -16, -14, -12, -10, -8,  -6,  -4,  -3,  -2,  -1,  0,
                               1,   2,   3,   4,   6,   8,   10,  12,  14,  16,  18,  20,  23,  26,  29,  32};
// common.rs:153-186 (VVC Table 25 fC)
static const int kFC[32][4] = {
    {0, 64, 0, 0},    {-1, 63, 2, 0},   {-2, 62, 4, 0},   {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
    {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
    {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
    {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
    {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
    {0, 4, 62, -2},   {0, 2, 63, -1}};
// common.rs:188-221 (fG): {16-(p>>1), 32-(p>>1), 16+(p>>1), p>>1}
static inline void fG(int p, int f[4]) {
    int h = p >> 1;
    f[0] = 16 - h;
    f[1] = 32 - h;
    f[2] = 16 + h;
    f[3] = h;
}

bool Tuning::parse(const char *s, std::string *err) {
    if (!s || !*s) return true;
    std::string str(s);
    size_t pos = 0;
    while (pos <= str.size()) {
        size_t e = str.find(',', pos);
        if (e == std::string::npos) e = str.size();
        std::string kv = str.substr(pos, e - pos);
        pos = e + 1;
        size_t eq = kv.find('=');
        if (eq == std::string::npos || kv.find('=', eq + 1) != std::string::npos) {
            if (err) *err = "Invalid extra-params: " + str;
            return false;
        }
        std::string key = kv.substr(0, eq), val = kv.substr(eq + 1);
        const char *v = val.c_str();
        if (key == "lv_pow_dq_trellis") lv_pow_dq_trellis = strtod(v, nullptr);
        else if (key == "lv_offset_dq_trellis") lv_offset_dq_trellis = strtod(v, nullptr);
        else if (key == "non_planar_offset_dq_trellis") non_planar_offset = strtof(v, nullptr);
        else if (key == "mpm_idx_offset_dq_trellis") mpm_idx_offset = strtof(v, nullptr);
        else if (key == "mpm_remainder_mult_dq_trellis") mpm_remainder_mult = strtof(v, nullptr);
        else if (key == "mpm_remainder_offset_dq_trellis") mpm_remainder_offset = strtof(v, nullptr);
        else if (key == "planer_offset_dq_trellis") planar_offset = strtof(v, nullptr);
        else if (key == "header_bits_dq_trellis") header_bits = strtof(v, nullptr);
        else if (key == "chroma_header_bits_dq_trellis") chroma_header_bits = strtof(v, nullptr);
        else if (key == "qp_div_dq_trellis") qp_div = strtof(v, nullptr);
        else if (key == "lambda_mul_dq_trellis") lambda_mul = strtof(v, nullptr);
        else if (key == "cclm_pow") cclm_pow = strtof(v, nullptr);
        else if (key == "mpm_idx_pow") mpm_idx_pow = strtof(v, nullptr);
        else if (key == "mpm_remainder_pow") mpm_remainder_pow = strtof(v, nullptr);
        else if (key == "cclm_mode_idx_offset_dq_trellis") cclm_mode_idx_offset = strtof(v, nullptr);
        else if (key == "non_cclm_offset_dq_trellis") non_cclm_offset = strtof(v, nullptr);
        else if (key == "cclm_offset_dq_trellis") cclm_offset = strtof(v, nullptr);
        else if (key == "a") { has_a = true; a = strtof(v, nullptr); }
        else if (key == "quant_lv_pow") quant_lv_pow = strtod(v, nullptr);
        else if (key == "quant_qp_div_trellis") quant_qp_div = strtod(v, nullptr);
        else if (key == "quant_lambda_mul_trellis") quant_lambda_mul = strtod(v, nullptr);
        else if (key == "quant_lambda_offset_trellis") quant_lambda_offset = strtoll(v, nullptr, 10);
        // any other key is stored but never read on the live (dep-quant + trellis) path
        if (e == str.size()) break;
    }
    return true;
}

void Consts::init(int qp_, const Tuning &t) {
    init_tables();
    qp = qp_;
    for (int i = 0; i < 1024; i++) {
        lv[i] = (int64_t)(std::pow((double)i + t.lv_offset_dq_trellis, t.lv_pow_dq_trellis) * 16384.0);  // block_splitter.rs:51-52
        dq[i] = (int64_t)std::pow((double)(i * 16384), t.quant_lv_pow);                                   // quantizer.rs:20-22
    }
    lambda_q = (int64_t)(std::pow(2.0, (double)qp / t.quant_qp_div) * t.quant_lambda_mul) + t.quant_lambda_offset;  // quantizer.rs:683
    lambda_rd = powf(2.0f, (float)qp / t.qp_div) * t.lambda_mul;                                                   // block_splitter.rs:472
    lambda_rd_c = t.has_a ? powf(2.0f, (float)qp / t.qp_div) * t.a : lambda_rd;                                    // block_splitter.rs:775-778
    static const int kLevelScale[6] = {40, 45, 51, 57, 64, 72};  // quantizer.rs:8 (rect_non_ts_flag = 0)
    ls = (16 * kLevelScale[(qp + 1) % 6]) << ((qp + 1) / 6);     // quantizer.rs:617-622, 325-333 (m = 16)
    // header-bit tables (block_splitter.rs:377-406)
    for (int lk = 0; lk < 67; lk++) {
        for (int ck = 0; ck < 5; ck++) {
            float cclm_bits;
            if (ck == 0) cclm_bits = t.non_cclm_offset;
            else if (ck <= 3) cclm_bits = t.cclm_offset + powf((float)(ck - 1) + t.cclm_mode_idx_offset, t.cclm_pow);
            else cclm_bits = 0.0f;  // DUAL_TREE_LUMA, cclm flag false
            float luma;
            if (lk == 0) luma = t.planar_offset;
            else if (lk <= 5) luma = t.non_planar_offset + powf((float)(lk - 1) + t.mpm_idx_offset, t.mpm_idx_pow);
            else luma = t.non_planar_offset + t.mpm_remainder_mult * powf((float)(lk - 6) + t.mpm_remainder_offset, t.mpm_remainder_pow);
            float mode_bits = luma + cclm_bits;
            if (ck <= 3) hdr_single[lk][ck] = (int64_t)((t.header_bits + mode_bits) * 16384.0f);
            else hdr_dual_luma[lk] = (int64_t)((t.header_bits / 3.0f + mode_bits) * 16384.0f);
        }
    }
    for (int ck = 0; ck < 4; ck++) {  // block_splitter.rs:695-712
        float mode_bits = ck == 0 ? t.non_cclm_offset : t.cclm_offset + powf((float)(ck - 1) + t.cclm_mode_idx_offset, t.cclm_pow);
        hdr_chroma[ck] = (int64_t)((t.chroma_header_bits + mode_bits) * 16384.0f);
    }
}

void Picture::init(int W_, int H_, const uint8_t *y, const uint8_t *cb, const uint8_t *cr) {
    W = W_;
    H = H_;
    const uint8_t *src[3] = {y, cb, cr};
    for (int c = 0; c < 3; c++) {
        int w = c ? W / 2 : W, h = c ? H / 2 : H;
        orig[c].alloc(w, h);
        rec[c].alloc(w, h);
        memcpy(orig[c].d.data(), src[c], (size_t)w * h);
        coef[c].assign((size_t)w * h, 0);
    }
    mode_map.assign((size_t)(W / 4) * (H / 4), 0);
    records.assign((size_t)(W / 32) * (H / 32), CtuRecord{});
}

// ------------------------------------------------------------------------------------------------
// availability (encoder_context.rs:918-956; check_pred_mode_y=false, WPP off)
// ------------------------------------------------------------------------------------------------
static inline bool nb_avail(const Picture &p, int xc, int yc, int xn, int yn, int w, int h, bool ar, bool bl) {
    return xn >= 0 && yn >= 0 && xn < p.W && yn < p.H && (((xn >> 5) <= (xc >> 5)) || ((yn >> 5) < (yc >> 5))) &&
           ((yn >> 5) < (yc >> 5) + 1) && (xn < xc + w || ar) && (yn < yc + h || bl);
}

// ------------------------------------------------------------------------------------------------
// intra prediction
// ------------------------------------------------------------------------------------------------

// intra_predictor.rs:146-353  set_left_and_above_ref_samples (ref_idx = 0, no ISP)
void build_refs(const Picture &p, const TU &tu, int c, int mode, int16_t *left, int16_t *above, int16_t *leftF,
                int16_t *aboveF) {
    const Plane &rec = p.rec[c];
    int cs = c != 0;
    int xt = tu.x >> cs, yt = tu.y >> cs, n = tu.w >> cs;
    int ref_w = 2 * n, ref_h = 2 * n;
    bool ref_filter_flag = (mode == 0 || mode == 2 || mode == 34 || mode == 66);  // :185-188
    int nl = ref_h + 1, na = ref_w;
    for (int i = 0; i < nl; i++) left[i] = -1;
    for (int i = 0; i < na; i++) above[i] = -1;
    bool available = true;
    int x_nb_cmp = xt - 1;
    int x_nb_y = x_nb_cmp << cs;
    for (int y = -1; y <= ref_h - 1; y++) {  // :205-228
        int y_nb_cmp = yt + y;
        int y_nb_y = y_nb_cmp << cs;
        if (y == -1 || y % 4 == 0) available = nb_avail(p, tu.x, tu.y, x_nb_y, y_nb_y, tu.w, tu.w, tu.ar, tu.bl);
        if (available) left[y + 1] = rec.at(x_nb_cmp, y_nb_cmp);
    }
    int y_nb_cmp = yt - 1;
    int y_nb_y = y_nb_cmp << cs;
    for (int x = 0; x <= ref_w - 1; x++) {  // :239-261
        int xn = xt + x;
        int xny = xn << cs;
        if (x == 0 || x % 4 == 0) available = nb_avail(p, tu.x, tu.y, xny, y_nb_y, tu.w, tu.w, tu.ar, tu.bl);
        if (available) above[x] = rec.at(xn, y_nb_cmp);
    }
    // substitution :263-302
    bool la = true, aa = true;
    for (int i = 0; i < nl; i++) la &= left[i] < 0;
    for (int i = 0; i < na; i++) aa &= above[i] < 0;
    if (la && aa) {
        for (int i = 0; i < nl; i++) left[i] = 128;
        for (int i = 0; i < na; i++) above[i] = 128;
    } else {
        if (left[nl - 1] < 0) {
            bool found = false;
            for (int i = nl - 2; i >= 0; i--)
                if (left[i] >= 0) {
                    left[nl - 1] = left[i];
                    found = true;
                    break;
                }
            if (!found)
                for (int i = 0; i < na; i++)
                    if (above[i] >= 0) {
                        left[nl - 1] = above[i];
                        break;
                    }
        }
        for (int y = ref_h - 2; y >= -1; y--)
            if (left[y + 1] < 0) left[y + 1] = left[y + 2];
    }
    if (above[0] < 0) above[0] = left[0];
    for (int x = 1; x <= ref_w - 1; x++)
        if (above[x] < 0) above[x] = above[x - 1];
    // [1 2 1] filter :304-352
    bool filter_flag = n * n > 32 && c == 0 && ref_filter_flag;
    if (filter_flag) {
        leftF[0] = (int16_t)((left[1] + 2 * left[0] + above[0] + 2) >> 2);
        for (int y = 0; y < ref_h - 1; y++) leftF[1 + y] = (int16_t)((left[2 + y] + 2 * left[1 + y] + left[y] + 2) >> 2);
        leftF[ref_h] = left[ref_h];
        aboveF[0] = (int16_t)((left[0] + 2 * above[0] + above[1] + 2) >> 2);
        for (int x = 0; x < ref_w - 2; x++) aboveF[1 + x] = (int16_t)((above[x] + 2 * above[x + 1] + above[x + 2] + 2) >> 2);
        aboveF[ref_w - 1] = above[ref_w - 1];
    } else {
        memcpy(leftF, left, sizeof(int16_t) * nl);
        memcpy(aboveF, above, sizeof(int16_t) * na);
    }
}

static inline int ilog2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) l++;
    return l;
}

// intra_predictor.rs:355-757  position_dependent_prediction_sample_filter
// ars = above (x index), lrs = left without corner (y index), sized 2N; pred is N x N in/out
static void pdpc(const int16_t *ars, const int16_t *lrs, int alrs, uint8_t *pred, int n, int pred_mode, int inv_angle) {
    static const int W0[3][12] = {{32, 8, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                  {32, 16, 8, 4, 2, 1, 0, 0, 0, 0, 0, 0},
                                  {32, 32, 16, 16, 8, 8, 4, 4, 2, 2, 1, 1}};
    auto W = [&](int ns, int i) { return i < 12 ? W0[ns][i] : 0; };
    int l2 = ilog2(n);
    int n_scale;
    if (pred_mode > 50) n_scale = std::min(l2 - ilog2(3 * inv_angle - 2) + 8, 2);
    else if (pred_mode > 1 && pred_mode < 18) n_scale = std::min(l2 - ilog2(3 * inv_angle - 2) + 8, 2);
    else n_scale = (l2 + l2 - 2) >> 2;
    std::vector<int> ref_l(n * n, 0), ref_t(n * n, 0), w_l(n, 0), w_t(n, 0);
    if (pred_mode < 2) {
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                ref_l[y * n + x] = lrs[y];
                ref_t[y * n + x] = ars[x];
            }
        for (int i = 0; i < n; i++) w_l[i] = w_t[i] = W(n_scale, i);
    } else if (pred_mode == 18 || pred_mode == 50) {
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                ref_l[y * n + x] = lrs[y] - alrs + pred[y * n + x];
                ref_t[y * n + x] = ars[x] - alrs + pred[y * n + x];
            }
        for (int i = 0; i < n; i++) {
            w_l[i] = pred_mode == 50 ? W(n_scale, i) : 0;
            w_t[i] = pred_mode == 18 ? W(n_scale, i) : 0;
        }
    } else if (pred_mode < 18 && n_scale >= 0) {
        for (int y = 0; y < n; y++) {
            int dxi = ((y + 1) * inv_angle + 256) >> 9;
            for (int x = 0; x < n; x++) {
                if (y < (3 << n_scale)) {
                    assert(x + dxi < 2 * n);
                    ref_t[y * n + x] = ars[x + dxi];
                }
            }
        }
        for (int i = 0; i < n; i++) w_t[i] = W(n_scale, i);  // returned as (w_l=ZERO, w_t=WEIGHTS) :438-441
    } else if (pred_mode > 50 && n_scale >= 0) {
        for (int x = 0; x < n; x++) {
            int dyi = ((x + 1) * inv_angle + 256) >> 9;
            for (int y = 0; y < n; y++) {
                if (x < (3 << n_scale)) {
                    assert(y + dyi < 2 * n);
                    ref_l[y * n + x] = lrs[y + dyi];
                }
            }
        }
        for (int i = 0; i < n; i++) w_l[i] = W(n_scale, i);
    }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {  // :746-754 (i16 arithmetic; no overflow for N<=32)
            int16_t v = (int16_t)(ref_l[y * n + x] * w_l[x] + ref_t[y * n + x] * w_t[y] + (64 - w_t[y] - w_l[x]) * pred[y * n + x] + 32);
            int r = v >> 6;
            pred[y * n + x] = (uint8_t)std::min(255, std::max(0, r));
        }
}

// intra_predictor.rs:1604-2055  predict_cclm (4:2:0, both collocated flags false)
static void predict_cclm(const Picture &p, const TU &tu, int c, uint8_t *pred) {
    int mode = tu.mode[c];
    int tw = tu.w / 2, th = tu.w / 2, tx = tu.x / 2, ty = tu.y / 2;
    bool avail_l = nb_avail(p, tu.x, tu.y, tu.x - 1, tu.y, tu.w, tu.w, false, false);
    bool avail_t = nb_avail(p, tu.x, tu.y, tu.x, tu.y - 1, tu.w, tu.w, false, false);
    int num_top_right = 0;
    bool avail_tr = true;
    if (mode == MODE_T_CCLM) {
        for (int x = tw; x < 2 * tw; x++) {
            if (!avail_tr) break;
            avail_tr = nb_avail(p, tu.x, tu.y, tu.x + x * 2, tu.y - 1, tu.w, tu.w, tu.ar, tu.bl);
            if (avail_tr) num_top_right++;
        }
    }
    int num_below_left = 0;
    bool avail_bl = true;
    if (mode == MODE_L_CCLM) {
        for (int y = th; y < 2 * th; y++) {
            if (!avail_bl) break;
            avail_bl = nb_avail(p, tu.x, tu.y, tu.x - 1, tu.y + y * 2, tu.w, tu.w, tu.ar, tu.bl);
            if (avail_bl) num_below_left++;
        }
    }
    int num_samp_t, num_samp_l;
    if (mode == MODE_LT_CCLM) {
        num_samp_t = avail_t ? tw : 0;
        num_samp_l = avail_l ? th : 0;
    } else {
        num_samp_t = (avail_t && mode == MODE_T_CCLM) ? tw + std::min(num_top_right, th) : 0;
        num_samp_l = (avail_l && mode == MODE_L_CCLM) ? th + std::min(num_below_left, tw) : 0;
    }
    bool b_ctu_boundary = (tu.y & 31) == 0;
    int is4 = !(avail_t && avail_l && mode == MODE_LT_CCLM);
    int start_pos_t = num_samp_t >> (2 + is4);
    int pick_step_t = std::max(1, num_samp_t >> (1 + is4));
    int cnt_t = 0, cnt_l = 0;
    int pick_t[4], pick_l[4];
    if (avail_t && (mode == MODE_LT_CCLM || mode == MODE_T_CCLM)) {
        cnt_t = std::min((1 + is4) << 1, num_samp_t);
        for (int i = 0; i < cnt_t; i++) pick_t[i] = start_pos_t + i * pick_step_t;
    }
    int start_pos_l = num_samp_l >> (2 + is4);
    int pick_step_l = std::max(1, num_samp_l >> (1 + is4));
    if (avail_l && (mode == MODE_LT_CCLM || mode == MODE_L_CCLM)) {
        cnt_l = std::min((1 + is4) << 1, num_samp_l);
        for (int i = 0; i < cnt_l; i++) pick_l[i] = start_pos_l + i * pick_step_l;
    }
    if (num_samp_l == 0 && num_samp_t == 0) {
        for (int i = 0; i < tw * th; i++) pred[i] = 128;
        return;
    }
    int dim = 2 * tu.w + 3;
    std::vector<long> py((size_t)dim * dim, 0);
    const int ox = 3, oy = 3;
    auto P = [&](int y, int x) -> long & { return py[(size_t)(y + oy) * dim + (x + ox)]; };
    const Plane &ry = p.rec[0];
    for (int y = 0; y < tu.w; y++)
        for (int x = 0; x < tu.w; x++) P(y, x) = ry.at(tu.x + x, tu.y + y);
    if (avail_l)
        for (int y = (avail_t ? -1 : 0); y < 2 * std::max(num_samp_l, th); y++)
            for (int x = -3; x <= -1; x++) P(y, x) = ry.at(tu.x + x, tu.y + y);
    if (!avail_t)
        for (int y = -2; y <= -1; y++)
            for (int x = -2; x < tu.w; x++) P(y, x) = P(0, x);
    if (avail_t)
        for (int y = -3; y <= -1; y++)
            for (int x = (avail_l ? -1 : 0); x < 2 * std::max(num_samp_t, tw); x++) P(y, x) = ry.at(tu.x + x, tu.y + y);
    if (!avail_l)
        for (int y = -2; y < 2 * th; y++) P(y, -1) = P(y, 0);
    std::vector<long> pds((size_t)tw * th);
    for (int y = 0; y < th; y++)
        for (int x = 0; x < tw; x++) {  // :1854-1868
            int sx = 2 * x, sy = 2 * y;
            pds[y * tw + x] = (P(sy, sx - 1) + P(sy + 1, sx - 1) + P(sy, sx) * 2 + P(sy + 1, sx) * 2 + P(sy, sx + 1) + P(sy + 1, sx + 1) + 4) >> 3;
        }
    long sel_y[4] = {0, 0, 0, 0}, sel_c[4] = {0, 0, 0, 0};
    assert(cnt_t + cnt_l == 4);  // the reference would index out of bounds (panic) otherwise (:1967-1972)
    const Plane &rc = p.rec[c];
    if (num_samp_t > 0) {
        for (int i = 0; i < cnt_t; i++) sel_c[i] = rc.at(tx + pick_t[i], ty - 1);
        for (int i = 0; i < cnt_t; i++) {
            int sx = 2 * pick_t[i];
            if (!b_ctu_boundary)
                sel_y[i] = (P(-1, sx - 1) + P(-2, sx - 1) + P(-1, sx) * 2 + P(-2, sx) * 2 + P(-1, sx + 1) + P(-2, sx + 1) + 4) >> 3;
            else
                sel_y[i] = (P(-1, sx - 1) + P(-1, sx) * 2 + P(-1, sx + 1) + 2) >> 2;
        }
    }
    if (num_samp_l > 0) {
        for (int i = cnt_t; i < cnt_t + cnt_l; i++) sel_c[i] = rc.at(tx - 1, ty + pick_l[i - cnt_t]);
        for (int i = cnt_t; i < cnt_t + cnt_l; i++) {
            int sx = -2, sy = 2 * pick_l[i - cnt_t];
            sel_y[i] = (P(sy, sx - 1) + P(sy + 1, sx - 1) + P(sy, sx) * 2 + P(sy + 1, sx) * 2 + P(sy, sx + 1) + P(sy + 1, sx + 1) + 4) >> 3;
        }
    }
    int mn[2] = {0, 2}, mx[2] = {1, 3};  // :1973-1990
    if (sel_y[mn[0]] > sel_y[mn[1]]) std::swap(mn[0], mn[1]);
    if (sel_y[mx[0]] > sel_y[mx[1]]) std::swap(mx[0], mx[1]);
    if (sel_y[mn[0]] > sel_y[mx[1]]) {
        std::swap(mn[0], mx[0]);
        std::swap(mn[1], mx[1]);
    }
    if (sel_y[mn[1]] > sel_y[mx[0]]) std::swap(mn[1], mx[0]);
    long max_y = (sel_y[mx[0]] + sel_y[mx[1]] + 1) >> 1;
    long max_c = (sel_c[mx[0]] + sel_c[mx[1]] + 1) >> 1;
    long min_y = (sel_y[mn[0]] + sel_y[mn[1]] + 1) >> 1;
    long min_c = (sel_c[mn[0]] + sel_c[mn[1]] + 1) >> 1;
    long diff = max_y - min_y;
    long a, b;
    int k;
    if (diff != 0) {  // :1994-2031
        long diff_c = max_c - min_c;
        int x = ilog2((int)diff);
        long norm_diff = ((diff << 4) >> x) & 15;
        x += norm_diff != 0;
        int y = std::labs(diff_c) > 0 ? ilog2((int)std::labs(diff_c)) + 1 : 0;
        static const int div_sig[16] = {0, 7, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1, 1, 1, 0};
        a = diff_c == 0 ? 0 : (diff_c * (div_sig[norm_diff] | 8) + (1L << (y - 1))) >> y;
        if (3 + x - y < 1) {
            k = 1;
            a = a < 0 ? -15 : (a > 0 ? 15 : 0);
        } else {
            k = 3 + x - y;
        }
        b = min_c - ((a * min_y) >> k);
    } else {
        a = 0;
        k = 0;
        b = min_c;
    }
    for (int i = 0; i < tw * th; i++) {
        long v = ((pds[i] * a) >> k) + b;
        pred[i] = (uint8_t)std::min(255L, std::max(0L, v));
    }
}

// intra_predictor.rs:56-144 predict (dispatch) + 759-1146 planar + 1148-1285 DC + 1287-1602 angular
void predict(const Picture &p, const TU &tu, int c, uint8_t *pred) {
    int mode = tu.mode[c];
    if (mode > 66) {
        predict_cclm(p, tu, c, pred);
        return;
    }
    int cs = c != 0;
    int n = tu.w >> cs;
    int l2 = ilog2(n);
    int16_t left[66], above[65], leftF[66], aboveF[65];
    build_refs(p, tu, c, mode, left, above, leftF, aboveF);
    const int16_t *lrs = leftF + 1;
    const int16_t *ars = aboveF;
    int alrs = leftF[0];
    if (mode == MODE_PLANAR) {
        int ars_r = ars[n], lrs_b = lrs[n];
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                int16_t pv = (int16_t)((n - 1 - y) * ars[x] + (y + 1) * lrs_b);
                int16_t ph = (int16_t)((n - 1 - x) * lrs[y] + (x + 1) * ars_r);
                int16_t v = (int16_t)(pv + ph + n);
                pred[y * n + x] = (uint8_t)(v >> (l2 + 1));
            }
        pdpc(ars, lrs, alrs, pred, n, 0, 0);
        return;
    }
    if (mode == MODE_DC) {
        int16_t s = (int16_t)n;
        for (int i = 0; i < n; i++) s = (int16_t)(s + ars[i]);
        for (int i = 0; i < n; i++) s = (int16_t)(s + lrs[i]);
        uint8_t dc = (uint8_t)(s >> (l2 + 1));
        for (int i = 0; i < n * n; i++) pred[i] = dc;
        pdpc(ars, lrs, alrs, pred, n, 1, 0);
        return;
    }
    // angular :1364-1559
    bool ref_filter_flag = (mode == 2 || mode == 34 || mode == 66);
    bool filter_flag;
    if (ref_filter_flag) {
        filter_flag = false;
    } else {
        int md = std::min(std::abs(mode - 50), std::abs(mode - 18));
        static const int thr[6] = {0, 0, 24, 14, 2, 0};
        filter_flag = md > thr[l2];
    }
    int ang = kAngle[mode];
    int inv_angle = ang > 0 ? (512 * 32 + ang / 2) / ang : (ang < 0 ? -((512 * 32 + (-ang) / 2) / -ang) : 0);
    const int16_t *lfull = leftF;  // index 0 = corner
    std::vector<int> refx;
    auto RX = [&](int idx) -> int { return idx < 0 ? refx[(int)refx.size() + idx] : refx[idx]; };
    if (mode >= 34) {
        refx.assign(n + 2, 0);
        refx[0] = alrs;
        for (int x = 0; x <= n; x++) refx[x + 1] = ars[x];
        if (ang < 0) {
            for (int x = -n; x <= -1; x++) refx.push_back(lfull[std::min((x * inv_angle + 256) >> 9, n)]);
        } else {
            for (int x = n + 2; x < 2 * n; x++) refx.push_back(ars[x - 1]);
            for (int i = 1; i <= 3; i++) refx.push_back(ars[2 * n - 1]);
        }
        for (int y = 0; y < n; y++) {
            int i_idx = ((y + 1) * ang) >> 5;
            int i_fact = ((y + 1) * ang) & 31;
            for (int x = 0; x < n; x++) {
                if (c == 0) {
                    int f[4];
                    if (filter_flag) fG(i_fact, f);
                    else memcpy(f, kFC[i_fact], sizeof(f));
                    long s = 0;
                    for (int i = 0; i < 4; i++) s += (long)f[i] * RX(x + i_idx + i);
                    long v = (s + 32) >> 6;
                    pred[y * n + x] = (uint8_t)std::min(255L, std::max(0L, v));
                } else if (i_fact != 0) {
                    pred[y * n + x] = (uint8_t)(((32 - i_fact) * RX(x + i_idx + 1) + i_fact * RX(x + i_idx + 2) + 16) >> 5);
                } else {
                    pred[y * n + x] = (uint8_t)RX(x + i_idx + 1);
                }
            }
        }
    } else {
        refx.assign(n + 2, 0);
        for (int x = 0; x <= n + 1; x++) refx[x] = lfull[x];
        if (ang < 0) {
            for (int x = -n; x <= -1; x++) {
                int idx = std::min((x * inv_angle + 256) >> 9, n);
                refx.push_back(idx == 0 ? alrs : ars[idx - 1]);
            }
        } else {
            for (int x = n + 2; x <= 2 * n; x++) refx.push_back(lfull[x]);
            for (int i = 1; i <= 2; i++) refx.push_back(lfull[2 * n]);
        }
        for (int x = 0; x < n; x++) {
            int i_idx = ((x + 1) * ang) >> 5;
            int i_fact = ((x + 1) * ang) & 31;
            for (int y = 0; y < n; y++) {
                if (c == 0) {
                    int f[4];
                    if (filter_flag) fG(i_fact, f);
                    else memcpy(f, kFC[i_fact], sizeof(f));
                    long s = 0;
                    for (int i = 0; i < 4; i++) s += (long)f[i] * RX(y + i_idx + i);
                    long v = (s + 32) >> 6;
                    pred[y * n + x] = (uint8_t)std::min(255L, std::max(0L, v));
                } else if (i_fact != 0) {
                    pred[y * n + x] = (uint8_t)(((32 - i_fact) * RX(y + i_idx + 1) + i_fact * RX(y + i_idx + 2) + 16) >> 5);
                } else {
                    pred[y * n + x] = (uint8_t)RX(y + i_idx + 1);
                }
            }
        }
    }
    if (mode <= 18 || mode >= 50) pdpc(ars, lrs, alrs, pred, n, mode, inv_angle);  // :1571-1591
}

// ------------------------------------------------------------------------------------------------
// transform (transformer.rs:2040-2378, 2380-2737; DCT-II only)
// ------------------------------------------------------------------------------------------------
void fwd_dct(const int16_t *res, int l2, int16_t *coef) {
    int n = 1 << l2;
    const int16_t *T = dct_matrix(l2);
    std::vector<int32_t> h((size_t)n * n), t((size_t)n * n);
    int s1 = l2 - 1, d1 = 1 << (s1 - 1);
    for (int y = 0; y < n; y++)
        for (int i = 0; i < n; i++) {
            int32_t s = 0;
            for (int x = 0; x < n; x++) s += (int32_t)T[i * n + x] * res[y * n + x];
            h[y * n + i] = (s + d1) >> s1;
        }
    int s2 = l2 + 6, d2 = 1 << (s2 - 1);
    for (int x = 0; x < n; x++)
        for (int i = 0; i < n; i++) {
            int32_t s = 0;
            for (int y = 0; y < n; y++) s += (int32_t)T[i * n + y] * h[y * n + x];
            t[i * n + x] = (s + d2) >> s2;
        }
    for (int i = 0; i < n * n; i++) coef[i] = (int16_t)t[i];
}

void inv_dct(const int16_t *deq, int l2, int16_t *out) {
    int n = 1 << l2;
    const int16_t *T = dct_matrix(l2);
    std::vector<int32_t> v((size_t)n * n);
    for (int x = 0; x < n; x++)
        for (int y = 0; y < n; y++) {
            int32_t s = 0;
            for (int i = 0; i < n; i++) s += (int32_t)T[i * n + y] * deq[i * n + x];
            s = (s + 64) >> 7;
            v[y * n + x] = std::min(32767, std::max(-32768, s));
        }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            int32_t s = 0;
            for (int i = 0; i < n; i++) s += (int32_t)T[i * n + x] * v[y * n + i];
            out[y * n + x] = (int16_t)((s + 2048) >> 12);
        }
}

// ------------------------------------------------------------------------------------------------
// dependent quantisation (quantizer.rs:338-517 search_dq, 519-759 quantize, 761-1079 dequantize)
// ------------------------------------------------------------------------------------------------
namespace {
struct DqEntry {
    int64_t a;
    int16_t q;
    int64_t cost;
    bool set;
};
struct DqSearch {
    const Consts &k;
    const int16_t *t;
    int l2, n, sh;
    int32_t off;
    const uint16_t *scan;
    std::vector<DqEntry> memo;  // [kidx][state]
    DqSearch(const Consts &k_, const int16_t *t_, int l2_) : k(k_), t(t_), l2(l2_) {
        n = 1 << l2;
        sh = l2 + 4;  // quantizer.rs:558-559 (bit_depth 8, dep_quant)
        off = (1 << sh) >> 1;
        scan = scan_order(l2);
        memo.assign((size_t)n * n * 4, DqEntry{0, 0, 0, false});
    }
    int64_t dq_cost(int64_t dist, int64_t bits) const {
        if (bits < 0 || bits >= 1024) {
            fprintf(stderr, "oracle: dq_table index %lld out of range (the reference would panic)\n", (long long)bits);
            abort();
        }
        return 128 * dist + k.lambda_q * k.dq[bits];
    }
    // kidx = sb*16 + pos (forward scan index)
    DqEntry search(int kidx, int q_state, bool itz) {
        static const int TR[4][2] = {{0, 2}, {2, 0}, {1, 3}, {3, 1}};  // encoder_context.rs:339
        DqEntry &m = memo[(size_t)kidx * 4 + q_state];
        if (m.set) return m;
        int xc = scan[kidx] & 255, yc = scan[kidx] >> 8;
        int32_t tc = t[yc * n + xc];
        int32_t lsc = k.ls;
        int64_t a;
        int16_t q;
        int64_t cost;
        if (kidx == 0) {  // last_scan_pos == 0 && last_sub_block == 0  (:367-409)
            if (tc == 0) {
                cost = dq_cost(0, 1 - (int64_t)itz);
                a = 0;
                q = 0;
            } else {
                uint64_t delta = q_state > 1;
                int32_t s = (tc << sh) - off;
                if (tc < 0) s = -s;
                uint64_t a0 = (uint64_t)(s / lsc / 2);
                int16_t q0 = (int16_t)(2 * a0 - delta);  // usize wrap, H3
                if (tc < 0) q0 = (int16_t)-q0;
                int32_t dq0 = ((int32_t)q0 * lsc + off) >> sh;
                int32_t d0 = std::abs(tc - dq0);
                int64_t cost0 = dq_cost(d0, (int64_t)(a0 + 1) * (int64_t)(a0 != 0 || !itz));
                uint64_t a1 = a0 + 1;
                int16_t q1 = (int16_t)(2 * a1 - delta);
                if (tc < 0) q1 = (int16_t)-q1;
                int32_t dq1 = ((int32_t)q1 * lsc + off) >> sh;
                int32_t d1 = std::abs(tc - dq1);
                int64_t cost1 = dq_cost(d1, (int64_t)(a1 + 1));
                if (cost0 <= cost1) {
                    a = (int64_t)a0;
                    q = q0;
                    cost = cost0;
                } else {
                    a = (int64_t)a1;
                    q = q1;
                    cost = cost1;
                }
            }
        } else {
            int next = kidx - 1;
            if (tc == 0) {
                int nq = TR[q_state][0];
                DqEntry c = search(next, nq, itz);
                cost = c.cost + dq_cost(0, 1 - (int64_t)itz);
                a = 0;
                q = 0;
            } else {
                int32_t s = (tc << sh) - off;
                if (tc < 0) s = -s;
                int32_t delta = q_state > 1;
                int64_t a0 = (s / lsc + delta) / 2;
                int nq0 = TR[q_state][a0 & 1];
                int32_t q0 = a0 > 0 ? 2 * (int32_t)a0 - delta : 0;
                if (tc < 0) q0 = -q0;
                int32_t dq0 = (q0 * lsc + off) >> sh;
                int32_t d0 = std::abs(tc - dq0);
                int64_t cost0 = (a0 == 0 && itz) ? dq_cost(d0, 0) : dq_cost(d0, a0 + 1);
                DqEntry c0 = search(next, nq0, itz && a0 == 0);
                cost0 += c0.cost;
                int64_t a1 = a0 + 1;
                int nq1 = TR[q_state][a1 & 1];
                long q1 = 2 * (long)a1 - (q_state > 1);
                if (tc < 0) q1 = -q1;
                int32_t dq1 = ((int32_t)q1 * lsc + off) >> sh;
                int32_t d1 = std::abs(tc - dq1);
                int64_t cost1 = dq_cost(d1, a1 + 1);
                DqEntry c1 = search(next, nq1, false);
                cost1 += c1.cost;
                if (cost0 <= cost1) {
                    a = a0;
                    q = (int16_t)q0;
                    cost = cost0;
                } else {
                    a = a1;
                    q = (int16_t)q1;
                    cost = cost1;
                }
            }
        }
        if ((kidx & 15) == 0 && itz && a == 0) cost -= k.lambda_q * k.dq[1];  // :512-514
        DqEntry &mm = memo[(size_t)kidx * 4 + q_state];
        mm = DqEntry{a, q, cost, true};
        return mm;
    }
};
}  // namespace

void quantize_dq(const Consts &k, const int16_t *coef, int l2, int16_t *qout) {
    static const int TR[4][2] = {{0, 2}, {2, 0}, {1, 3}, {3, 1}};
    int n = 1 << l2;
    DqSearch S(k, coef, l2);
    int q_state = 0;
    bool itz = true;
    for (int kidx = n * n - 1; kidx >= 0; kidx--) {  // quantizer.rs:686-721
        DqEntry e = S.search(kidx, q_state, itz);
        itz = itz && e.a == 0;
        int xc = S.scan[kidx] & 255, yc = S.scan[kidx] >> 8;
        qout[yc * n + xc] = e.q;
        q_state = TR[q_state][e.a & 1];
    }
}

void dequantize(const Consts &k, const int16_t *q, int l2, int16_t *d) {
    int n = 1 << l2, sh = l2 + 4, off = (1 << sh) >> 1;
    for (int i = 0; i < n * n; i++) {
        int32_t v = ((int32_t)q[i] * k.ls + off) >> sh;  // quantizer.rs:1074-1075
        d[i] = (int16_t)std::min(32767, std::max(-32768, v));
    }
}

// block_splitter.rs:415-460 / 727-763: dep-quant rate walk of one TB
int64_t rate_levels(const Consts &k, const int16_t *q, int l2) {
    static const int TR[4][2] = {{0, 2}, {2, 0}, {1, 3}, {3, 1}};
    int n = 1 << l2;
    const uint16_t *scan = scan_order(l2);
    int64_t sum = 0;
    int q_state = 0;
    bool itz = true;
    for (int kidx = n * n - 1; kidx >= 0; kidx--) {
        int xc = scan[kidx] & 255, yc = scan[kidx] >> 8;
        int qc = std::abs((int)q[yc * n + xc]);
        if (qc == 0) {
            sum += itz ? 0 : k.lv[0];
            q_state = TR[q_state][0];
        } else {
            int a = (qc + (q_state > 1)) / 2;
            if (a >= 1024) {
                fprintf(stderr, "oracle: lv_table index out of range (the reference would panic)\n");
                abort();
            }
            sum += k.lv[a];
            q_state = TR[q_state][a & 1];
        }
        itz = itz && qc == 0;
    }
    return sum;
}

// ------------------------------------------------------------------------------------------------
// the search (block_splitter.rs)
// ------------------------------------------------------------------------------------------------
namespace {
struct Node {
    int x, y, w, tree;
    bool ar, bl;
    bool is_root = false;
    int cu_mode[3] = {0, 0, 0};  // CodingUnit.intra_pred_mode (ctu.rs:1241,1328)
    int tu_mode[3] = {0, 0, 0};  // TransformUnit.cu_intra_pred_mode
    bool split = false;
    std::vector<Node> ch;
    std::vector<int16_t> q[3];  // tu.quantized_transformed_coeffs
};
}  // namespace

struct SearchImpl {
    Encoder &E;
    Picture &P;
    const Consts &K;
    int ctu_x, ctu_y;
    int root_mode = 0;  // luma mode of the CTU root's CU as seen by tile.get_cu during the search (H1)

    SearchImpl(Encoder &e, Picture &p, int cx, int cy) : E(e), P(p), K(e.k), ctu_x(cx), ctu_y(cy) {}

    static bool is_cclm(int m) { return m >= 81; }
    bool active(const Node &n, int c) const {
        return n.tree == DUAL_TREE_LUMA ? c == 0 : (n.tree == DUAL_TREE_CHROMA ? c != 0 : true);
    }
    // ctu.rs:1372-1381
    void set_mode(Node &n, const int m[3]) {
        n.cu_mode[0] = m[0];
        int derived = is_cclm(m[1]) ? m[1] : m[0];  // ctu.rs:1672-1733 with intra_chroma_pred_mode == 4
        n.cu_mode[1] = n.cu_mode[2] = derived;
        n.tu_mode[0] = m[0];
        n.tu_mode[1] = m[1];
        n.tu_mode[2] = m[2];
        if (n.is_root) root_mode = n.cu_mode[0];
    }
    TU make_tu(const Node &n) const {
        TU t{n.x, n.y, n.w, n.tree, n.ar, n.bl, {n.tu_mode[0], n.tu_mode[1], n.tu_mode[2]}};
        return t;
    }
    // luma mode of the CU that tile.get_cu(px,py) returns while this CTU is being searched (H1)
    int cu_mode_at(int px, int py, bool *exists) const {
        if (px < 0 || py < 0 || px >= P.W || py >= P.H) {
            *exists = false;
            return 0;
        }
        *exists = true;
        if ((px >> 5) == (ctu_x >> 5) && (py >> 5) == (ctu_y >> 5)) return root_mode;
        return P.mode_map[(size_t)(py >> 2) * (P.W / 4) + (px >> 2)];
    }
    // ctu.rs:1498-1635: returns luma kind (0 planar, 1..5 mpm idx, 6.. remainder)
    int luma_kind(const Node &n) const {
        int mode = n.cu_mode[0];
        if (mode == MODE_PLANAR) return 0;
        bool ex;
        int lm = cu_mode_at(n.x - 1, n.y + n.w - 1, &ex);
        int left = ex ? lm : MODE_PLANAR;
        int am = cu_mode_at(n.x + n.w - 1, n.y - 1, &ex);
        int above = (ex && !(n.y - 1 < ((n.y >> 5) << 5))) ? am : MODE_PLANAR;
        int cand[5];
        mpm_list(left, above, cand);
        for (int i = 0; i < 5; i++)
            if (cand[i] == mode) return 1 + i;
        std::sort(cand, cand + 5);
        int rem;
        if (mode > cand[4]) rem = mode - 6;
        else if (mode > cand[3]) rem = mode - 5;
        else if (mode > cand[2]) rem = mode - 4;
        else if (mode > cand[1]) rem = mode - 3;
        else if (mode > cand[0]) rem = mode - 2;
        else rem = mode - 1;
        return 6 + rem;
    }
    static void mpm_list(int left, int above, int cand[5]) {  // ctu.rs:1530-1601
        if (left == above && left > MODE_DC) {
            int m = left;
            int v[5] = {m, 2 + (m + 61) % 64, 2 + (m - 1) % 64, 2 + (m + 60) % 64, 2 + m % 64};
            memcpy(cand, v, sizeof(v));
        } else if (left != above && (left > MODE_DC || above > MODE_DC)) {
            int mn = std::min(left, above), mx = std::max(left, above);
            if (mn > MODE_DC) {
                int d = mx - mn;
                if (d == 1) {
                    int v[5] = {left, above, 2 + (mn + 61) % 64, 2 + (mx - 1) % 64, 2 + (mn + 60) % 64};
                    memcpy(cand, v, sizeof(v));
                } else if (d >= 62) {
                    int v[5] = {left, above, 2 + (mn - 1) % 64, 2 + (mx + 61) % 64, 2 + mn % 64};
                    memcpy(cand, v, sizeof(v));
                } else if (d == 2) {
                    int v[5] = {left, above, 2 + (mn - 1) % 64, 2 + (mn + 61) % 64, 2 + (mx - 1) % 64};
                    memcpy(cand, v, sizeof(v));
                } else {
                    int v[5] = {left, above, 2 + (mn + 61) % 64, 2 + (mn - 1) % 64, 2 + (mx + 61) % 64};
                    memcpy(cand, v, sizeof(v));
                }
            } else {
                int v[5] = {mx, 2 + (mx + 61) % 64, 2 + (mx - 1) % 64, 2 + (mx + 60) % 64, 2 + mx % 64};
                memcpy(cand, v, sizeof(v));
            }
        } else {
            int v[5] = {MODE_DC, 50, 18, 46, 54};
            memcpy(cand, v, sizeof(v));
        }
    }

    // predict one component into `pred`, write nothing else
    void do_predict(const Node &n, int c, std::vector<uint8_t> &pred) {
        int cn = c ? n.w / 2 : n.w;
        pred.resize((size_t)cn * cn);
        TU t = make_tu(n);
        predict(P, t, c, pred.data());
        E.n_predictions++;
    }
    // predict -> transform -> quantize -> dequantize -> inverse -> recon (block_splitter.rs:148-183); returns ssd
    uint64_t pipeline(Node &n, int c, bool want_ssd) {
        int cs = c != 0;
        int cn = n.w >> cs, cx = n.x >> cs, cy = n.y >> cs;
        int l2 = ilog2(cn);
        std::vector<uint8_t> pred;
        do_predict(n, c, pred);
        std::vector<int16_t> res((size_t)cn * cn), coef(res.size()), deq(res.size()), itr(res.size());
        for (int y = 0; y < cn; y++)
            for (int x = 0; x < cn; x++) res[y * cn + x] = (int16_t)((int)P.orig[c].at(cx + x, cy + y) - (int)pred[y * cn + x]);
        fwd_dct(res.data(), l2, coef.data());
        n.q[c].resize(res.size());
        quantize_dq(K, coef.data(), l2, n.q[c].data());
        dequantize(K, n.q[c].data(), l2, deq.data());
        inv_dct(deq.data(), l2, itr.data());
        uint64_t ssd = 0;
        for (int y = 0; y < cn; y++)
            for (int x = 0; x < cn; x++) {
                int16_t r16 = (int16_t)((int16_t)pred[y * cn + x] + itr[y * cn + x]);
                int rec = std::min(255, std::max(0, (int)r16));
                P.rec[c].at(cx + x, cy + y) = (uint8_t)rec;
                int d = rec - (int)P.orig[c].at(cx + x, cy + y);
                ssd += (uint64_t)(d * d);
            }
        E.n_pipelines++;
        (void)want_ssd;
        return ssd;
    }

    // block_splitter.rs:64-108
    float aux_cost(const int m[3], Node &n) {
        set_mode(n, m);
        uint64_t sad = 0;
        std::vector<uint8_t> pred;
        for (int c = 0; c < 3; c++)
            if (active(n, c)) {
                do_predict(n, c, pred);
                int cs = c != 0, cn = n.w >> cs, cx = n.x >> cs, cy = n.y >> cs;
                for (int y = 0; y < cn; y++)
                    for (int x = 0; x < cn; x++) sad += (uint64_t)std::abs((int)pred[y * cn + x] - (int)P.orig[c].at(cx + x, cy + y));
            }
        return (float)sad;
    }
    // block_splitter.rs:110-474
    float pred_cost(const int m[3], Node &n) {
        set_mode(n, m);
        int lk = luma_kind(n);
        bool cclm_flag = is_cclm(n.cu_mode[1]);
        int cclm_idx = cclm_flag ? n.cu_mode[1] - MODE_LT_CCLM : 0;
        uint64_t ssd = 0;
        for (int c = 0; c < 3; c++)
            if (active(n, c)) ssd += pipeline(n, c, true);
        int64_t header;
        if (n.tree == SINGLE_TREE) header = K.hdr_single[lk][cclm_flag ? 1 + cclm_idx : 0];
        else if (n.tree == DUAL_TREE_LUMA) header = cclm_flag ? 0 /*unreachable*/ : K.hdr_dual_luma[lk];
        else {
            fprintf(stderr, "oracle: get_intra_pred_cost on DUAL_TREE_CHROMA is unreachable in the reference\n");
            abort();
        }
        int64_t level = 0;
        for (int c = 0; c < 3; c++)
            if (active(n, c)) level += rate_levels(K, n.q[c].data(), ilog2(c ? n.w / 2 : n.w));
        level += header;
        return (float)ssd + K.lambda_rd * ((float)level / 16384.0f);
    }
    // block_splitter.rs:476-522
    float chroma_aux_cost(int mode, Node &n) {
        int m[3] = {n.cu_mode[0], mode, mode};
        set_mode(n, m);
        uint64_t sad = 0;
        std::vector<uint8_t> pred;
        for (int c = 1; c < 3; c++)
            if (active(n, c)) {
                do_predict(n, c, pred);
                int cn = n.w / 2, cx = n.x / 2, cy = n.y / 2;
                for (int y = 0; y < cn; y++)
                    for (int x = 0; x < cn; x++) sad += (uint64_t)std::abs((int)pred[y * cn + x] - (int)P.orig[c].at(cx + x, cy + y));
            }
        return (float)sad;
    }
    // block_splitter.rs:524-780
    float chroma_pred_cost(int mode, Node &n) {
        int m[3] = {n.cu_mode[0], mode, mode};
        set_mode(n, m);
        bool cclm_flag = is_cclm(n.cu_mode[1]);
        int cclm_idx = cclm_flag ? n.cu_mode[1] - MODE_LT_CCLM : 0;
        uint64_t ssd = 0;
        for (int c = 1; c < 3; c++)
            if (active(n, c)) ssd += pipeline(n, c, true);
        if (n.tree == DUAL_TREE_LUMA) abort();
        int64_t level = 0;
        for (int c = 1; c < 3; c++) level += rate_levels(K, n.q[c].data(), ilog2(n.w / 2));
        level += K.hdr_chroma[cclm_flag ? 1 + cclm_idx : 0];
        return (float)ssd + K.lambda_rd_c * ((float)level / 16384.0f);
    }

    void save_rec(const Node &n, int c, std::vector<uint8_t> &buf) {
        int cs = c != 0, cn = n.w >> cs, cx = n.x >> cs, cy = n.y >> cs;
        buf.resize((size_t)cn * cn);
        for (int y = 0; y < cn; y++)
            for (int x = 0; x < cn; x++) buf[y * cn + x] = P.rec[c].at(cx + x, cy + y);
    }
    void restore_rec(const Node &n, int c, const std::vector<uint8_t> &buf) {
        int cs = c != 0, cn = n.w >> cs, cx = n.x >> cs, cy = n.y >> cs;
        for (int y = 0; y < cn; y++)
            for (int x = 0; x < cn; x++) P.rec[c].at(cx + x, cy + y) = buf[y * cn + x];
    }

    int best_cclm(Node &n) {  // block_splitter.rs:841-854 / 1041-1054
        float lt = chroma_aux_cost(MODE_LT_CCLM, n);
        float t = chroma_aux_cost(MODE_T_CCLM, n);
        float l = chroma_aux_cost(MODE_L_CCLM, n);
        if (lt <= t && lt <= l) return MODE_LT_CCLM;
        if (t <= l) return MODE_T_CCLM;
        return MODE_L_CCLM;
    }

    // block_splitter.rs:782-1154
    float split_ct(Node &n, Node *parent, int max_depth) {
        if (max_depth == 0) {
            if (n.tree == DUAL_TREE_CHROMA) {  // :794-885
                // luma CU covering the parent's centre sample (ctu.rs:2372-2396: first matching child)
                int px = parent->x + parent->w / 2, py = parent->y + parent->w / 2;
                int dm = 0;
                for (const Node &s : parent->ch)
                    if (px >= s.x && px < s.x + s.w && py >= s.y && py < s.y + s.w) {
                        dm = s.cu_mode[1];  // get_intra_chroma_pred_mode...: derived chroma mode of the luma CU
                        break;
                    }
                int cclm_mode = best_cclm(n);
                float cclm_cost = chroma_pred_cost(cclm_mode, n);
                std::vector<uint8_t> save[3];
                save_rec(n, 1, save[1]);
                save_rec(n, 2, save[2]);
                float current_cost = chroma_pred_cost(dm, n);
                float mn = std::min(current_cost, cclm_cost);
                int idx = (current_cost == mn) ? 0 : 1;
                if (idx == 1) {
                    int m[3] = {cclm_mode, cclm_mode, cclm_mode};
                    set_mode(n, m);
                    restore_rec(n, 1, save[1]);
                    restore_rec(n, 2, save[2]);
                }
                return mn;
            }
            static const int cand_modes[15] = {0, 1, 2, 7, 13, 18, 23, 29, 34, 39, 45, 50, 55, 60, 66};
            float cand_costs[15];
            for (int i = 0; i < 15; i++) {
                int m[3] = {cand_modes[i], cand_modes[i], cand_modes[i]};
                cand_costs[i] = cand_modes[i] <= 1 ? pred_cost(m, n) : aux_cost(m, n);
            }
            float min_dir_cost = cand_costs[2];
            int min_dir_idx = 2;
            for (int i = 3; i < 15; i++)
                if (cand_costs[i] < min_dir_cost) {
                    min_dir_cost = cand_costs[i];
                    min_dir_idx = i;
                }
            auto step_search = [&](int current_mode, int step, float current_cost, bool aux, float *out_cost) -> int {
                if (!aux) {
                    int m[3] = {current_mode, current_mode, current_mode};
                    current_cost = pred_cost(m, n);
                }
                while (step > 0) {
                    float cost0, cost1;
                    if (current_mode < 2 + step) cost0 = 3.402823466e+38f;
                    else {
                        int m[3] = {current_mode - step, current_mode - step, current_mode - step};
                        cost0 = aux ? aux_cost(m, n) : pred_cost(m, n);
                    }
                    if (current_mode + step > 66) cost1 = 3.402823466e+38f;
                    else {
                        int m[3] = {current_mode + step, current_mode + step, current_mode + step};
                        cost1 = aux ? aux_cost(m, n) : pred_cost(m, n);
                    }
                    float mn = std::min(std::min(current_cost, cost0), cost1);
                    if (current_cost == mn) {
                    } else if (cost0 == mn) {
                        current_mode -= step;
                        current_cost = cost0;
                    } else {
                        current_mode += step;
                        current_cost = cost1;
                    }
                    step /= 2;
                }
                *out_cost = current_cost;
                return current_mode;
            };
            float tmpc;
            int dir_mode = step_search(cand_modes[min_dir_idx], 2, min_dir_cost, true, &tmpc);
            float dir_cost;
            dir_mode = step_search(dir_mode, 1, min_dir_cost, false, &dir_cost);
            int cm[3] = {0, 1, dir_mode};
            float cc[3] = {cand_costs[0], cand_costs[1], dir_cost};
            float min_cost = std::min(std::min(cc[0], cc[1]), cc[2]);
            int mi = cc[0] == min_cost ? 0 : (cc[1] == min_cost ? 1 : 2);
            int mode = cm[mi];
            {
                int m[3] = {mode, mode, mode};
                set_mode(n, m);
            }
            if (active(n, 0)) pipeline(n, 0, false);  // :989-1037 luma redo
            if (n.tree != DUAL_TREE_LUMA) {           // cclm_enabled_flag = true (sps.rs:320)
                float current_cost = chroma_pred_cost(mode, n);
                int cclm_mode = best_cclm(n);
                float cclm_cost = chroma_pred_cost(cclm_mode, n);
                float mn = std::min(current_cost, cclm_cost);
                if (current_cost == mn) {
                    int m[3] = {mode, mode, mode};
                    set_mode(n, m);
                    min_cost = pred_cost(m, n);
                } else {
                    int m[3] = {mode, cclm_mode, cclm_mode};
                    min_cost = pred_cost(m, n);
                }
            } else if (mode <= 1) {
                int m[3] = {mode, mode, mode};
                min_cost = pred_cost(m, n);
            }
            return min_cost;
        }
        float no_split_cost = split_ct(n, parent, 0);
        Node sp = n;  // the clone (block_splitter.rs:1081-1084); shares the CU until split() clears cus
        std::vector<uint8_t> save[3];
        for (int c = 0; c < 3; c++)
            if (active(n, c)) save_rec(n, c, save[c]);
        // CodingTree::split(SPLIT_QT) ctu.rs:1960-2064
        sp.split = true;
        sp.is_root = false;  // children are new CTs; the clone itself no longer owns a CU
        sp.ch.clear();
        int child_tree = n.w == 8 ? DUAL_TREE_LUMA : n.tree;
        for (int i = 0; i < 4; i++) {
            Node c;
            c.w = n.w / 2;
            c.x = n.x + (i % 2) * c.w;
            c.y = n.y + (i / 2) * c.w;
            c.tree = child_tree;
            // availability ctu.rs:2083-2188
            if (c.x + c.w >= P.W) c.ar = false;
            else if (i == 0) c.ar = 0 < c.y;
            else if (i == 1) c.ar = n.ar;
            else if (i == 2) c.ar = true;
            else c.ar = false;
            if (c.y + c.w >= P.H) c.bl = false;
            else if (i == 1 || i == 3) c.bl = false;
            else if (i == 0) c.bl = 0 < c.x;
            else c.bl = n.bl;
            sp.ch.push_back(c);
        }
        if (n.w == 8) {  // local dual tree: extra chroma CT (ctu.rs:2031-2055)
            Node c;
            c.w = n.w;
            c.x = n.x;
            c.y = n.y;
            c.tree = DUAL_TREE_CHROMA;
            c.ar = (c.x + c.w >= P.W) ? false : n.ar;  // same size as parent: inherits (ctu.rs:2131-2134)
            c.bl = (c.y + c.w >= P.H) ? false : n.bl;  // x == ct.x and y+h == ct.y+ct.h: inherits (ctu.rs:2104-2107)
            sp.ch.push_back(c);
        }
        float split_cost = 0.0f;
        for (size_t i = 0; i < sp.ch.size(); i++) split_cost += split_ct(sp.ch[i], &sp, max_depth - 1);
        if (split_cost > no_split_cost) {
            for (int c = 0; c < 3; c++)
                if (active(n, c)) restore_rec(n, c, save[c]);
            return no_split_cost;
        }
        bool was_root = n.is_root;
        n = sp;  // *ct = split_ct.clone()
        n.is_root = was_root;
        return split_cost;
    }

    // second pass in coding order (ctu_encoder.rs:1421-1461) + record extraction
    void finalize(Node &n, CtuRecord &r, int depth, int zidx) {
        if (n.split) {
            if (depth == 0) r.split_mask |= 1u;
            else if (depth == 1) r.split_mask |= 1u << (1 + zidx);
            else if (depth == 2) r.split_mask |= 1u << (5 + zidx);
            for (size_t i = 0; i < n.ch.size(); i++) finalize(n.ch[i], r, depth + 1, zidx * 4 + (int)(i & 3));
            return;
        }
        for (int c = 0; c < 3; c++)
            if (active(n, c)) {
                std::vector<uint8_t> before;
                save_rec(n, c, before);
                pipeline(n, c, false);
                E.n_pipelines--;  // the second pass is excluded from the nominal work count
                std::vector<uint8_t> after;
                save_rec(n, c, after);
                if (before != after) {
                    fprintf(stderr, "oracle: H7 violated (second pass changed the reconstruction) at (%d,%d) w=%d c=%d\n", n.x, n.y, n.w, c);
                    abort();
                }
                int cs = c != 0, cn = n.w >> cs, cx = n.x >> cs, cy = n.y >> cs, pw = P.orig[c].w;
                // TB-local raster inside the TB's own area of the coefficient plane
                for (int y = 0; y < cn; y++)
                    for (int x = 0; x < cn; x++) P.coef[c][(size_t)(cy + y) * pw + cx + x] = n.q[c][y * cn + x];
            }
        if (active(n, 0)) {
            for (int y = 0; y < n.w; y += 4)
                for (int x = 0; x < n.w; x += 4) {
                    int bx = (n.x - ctu_x + x) >> 2, by = (n.y - ctu_y + y) >> 2;
                    r.luma_mode[by * 8 + bx] = (uint8_t)n.tu_mode[0];
                    P.mode_map[(size_t)((n.y + y) >> 2) * (P.W / 4) + ((n.x + x) >> 2)] = (uint8_t)n.tu_mode[0];
                }
        }
        if (active(n, 1)) {
            for (int y = 0; y < n.w; y += 8)
                for (int x = 0; x < n.w; x += 8) {
                    int bx = (n.x - ctu_x + x) >> 3, by = (n.y - ctu_y + y) >> 3;
                    r.chroma_mode[by * 4 + bx] = (uint8_t)n.tu_mode[1];
                }
        }
    }
};

void Encoder::search_ctu(Picture &p, int cx, int cy) {
    SearchImpl S(*this, p, cx, cy);
    Node root;
    root.x = cx;
    root.y = cy;
    root.w = 32;
    root.tree = SINGLE_TREE;
    root.is_root = true;
    // ctu.rs:2083-2092,2114-2117: root below-left false; :2124-2127,2183-2186: root above-right
    root.bl = false;
    root.ar = (cx + 32 >= p.W) ? false : (0 < cy && cx + 32 < p.W);
    float cost = S.split_ct(root, nullptr, max_depth);
    CtuRecord &r = p.records[(size_t)(cy / 32) * (p.W / 32) + cx / 32];
    memset(&r, 0, sizeof(r));
    r.cost = cost;
    S.finalize(root, r, 0, 0);
}

void Encoder::search_picture(Picture &p) {
    pic = &p;
    for (int cy = 0; cy < p.H; cy += 32)
        for (int cx = 0; cx < p.W; cx += 32) search_ctu(p, cx, cy);
}

}  // namespace wo
