// wrenc_oracle_cabac.cpp — CPU ORACLE (test infrastructure, NOT product code; see wrenc_oracle.hpp header).
// Restatement of the reference's syntax writer + CABAC engine for the I-slice subset the all-intra path emits:
//   CtuEncoder::encode_coding_tree / encode_coding_unit / encode_transform_unit / encode_residual  (ctu_encoder.rs:227-2269)
//   BoolCoder arithmetic engine, context init, binarisations, ctxInc derivations                  (bool_coder.rs)
//   end_of_slice_one_bit + byte alignment of SliceEncoder::encode                                 (slice_encoder.rs:380-388,419)
// It works on the searched Picture (final levels, records, mode map) and drives the arithmetic coder bin by bin, like the
// reference.  Pinned by data the reference produced: wrapped in the product's header writers, its output reproduces the sizes of
// all 16 .vvc files of tools/evaluation/summary.json to the byte (tests/test_reference_pin.py).  The trace hooks at the end of
// the file hand the bin string and the engine to the tests of the product's arithmetic coder (tests/host/cabac_engine_host_test.cpp).
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wrenc_oracle.hpp"
#include "cabac_tables.inc"

namespace wo {
namespace {

static const int TRS[4][2] = {{0, 2}, {2, 0}, {1, 3}, {3, 1}};  // encoder_context.rs:339
static const int kRice[32] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3};  // cabac_contexts.rs:919-921

struct BoolCoder {
    std::vector<bool> bits;
    uint16_t p[CTX_TOTAL][2];
    unsigned range = 510, offset = 0;
    int outstanding = 0;
    bool first_bit = true;
    std::vector<uint16_t> *trace = nullptr;  // test hook: the bin string in the product's entry format (ctx | bin << 9 | bypass << 10)

    void init(int slice_qp) {  // bool_coder.rs:1073-1111
        for (int i = 0; i < CTX_TOTAL; i++) {
            int init_value = kCabacInitValue[i];
            int m = (init_value >> 3) - 4, n = (init_value & 7) * 18 + 1;
            int pre = std::min(127, std::max(1, ((m * (std::min(63, std::max(0, slice_qp)) - 16)) >> 1) + n));
            p[i][0] = (uint16_t)(pre << 3);
            p[i][1] = (uint16_t)(pre << 7);
        }
        range = 510;
        offset = 0;
        outstanding = 0;
        first_bit = true;
    }
    void flush_bin(bool b) {  // :174-191
        if (!first_bit) bits.push_back(b);
        first_bit = false;
        while (outstanding > 0) {
            bits.push_back(!b);
            outstanding--;
        }
    }
    void flush_trailing(bool b) {  // :193-199
        bits.push_back(b);
        while (outstanding > 0) {
            bits.push_back(!b);
            outstanding--;
        }
    }
    void renorm() {  // :157-171
        while (range < 256) {
            if (offset < 256) flush_bin(false);
            else if (offset >= 512) {
                offset -= 512;
                flush_bin(true);
            } else {
                offset -= 256;
                outstanding++;
            }
            range <<= 1;
            offset <<= 1;
        }
    }
    void bypass(bool b) {  // :202-216
        if (trace) trace->push_back((uint16_t)(((unsigned)b << 9) | (1u << 10)));
        offset <<= 1;
        if (b) offset += range;
        if (offset >= 1024) {
            flush_bin(true);
            offset -= 1024;
        } else if (offset < 512) flush_bin(false);
        else {
            offset -= 512;
            outstanding++;
        }
    }
    void decision(int ctx, bool b) {  // :254-296 + :136-154
        if (trace) trace->push_back((uint16_t)((unsigned)ctx | ((unsigned)b << 9)));
        unsigned q_range_idx = range >> 5;
        unsigned p_state = p[ctx][1] + 16u * p[ctx][0];
        unsigned val_mps = p_state >> 14;
        unsigned lps = ((q_range_idx * ((val_mps == 0 ? p_state : 32767u - p_state) >> 9)) >> 1) + 4;
        if (val_mps == (unsigned)b) range -= lps;
        else {
            offset += range - lps;
            range = lps;
        }
        renorm();
        int shift_idx = kCabacShiftIdx[ctx];
        int s0 = (shift_idx >> 2) + 2, s1 = (shift_idx & 3) + 3 + s0;
        p[ctx][0] = (uint16_t)(p[ctx][0] - (p[ctx][0] >> s0) + ((1023 * (int)b) >> s0));
        p[ctx][1] = (uint16_t)(p[ctx][1] - (p[ctx][1] >> s1) + ((16383 * (int)b) >> s1));
    }
    void stop_one_bit() {  // :218-235 with bin == 1
        range -= 2;
        offset += range;
        range = 2;
        renorm();
        flush_bin(((offset >> 9) & 1) != 0);
        unsigned two = ((offset >> 7) & 3) | 1;
        flush_trailing(((two >> 1) & 1) != 0);
        flush_trailing((two & 1) != 0);
        first_bit = true;
        outstanding = 0;
    }
    void bypass_fixed(unsigned v, int nbits) {
        for (int i = nbits - 1; i >= 0; i--) bypass(((v >> i) & 1) != 0);
    }
};

struct SliceWriter {
    const Consts &K;
    Picture &P;
    BoolCoder c;
    bool is_cu_qp_delta_coded = false;
    bool mts_dc_only = true, mts_zero_out_sig_coeff_flag = true;

    SliceWriter(const Consts &k, Picture &p) : K(k), P(p) {}

    const CtuRecord &record_at(int px, int py) const { return P.records[(size_t)(py / 32) * (P.W / 32) + px / 32]; }
    // width/height of the coding tree leaf that holds the CU covering luma sample (px, py)   (ctu.rs:2209-2245)
    int ct_size_at(int px, int py) const {
        const CtuRecord &r = record_at(px, py);
        int lx = px % 32, ly = py % 32;
        if (!(r.split_mask & 1u)) return 32;
        int a = (ly / 16) * 2 + lx / 16;
        if (!((r.split_mask >> (1 + a)) & 1u)) return 16;
        int b = ((ly % 16) / 8) * 2 + (lx % 16) / 8;
        if (!((r.split_mask >> (5 + a * 4 + b)) & 1u)) return 8;
        return 4;
    }
    int luma_mode_at(int px, int py) const { return P.mode_map[(size_t)(py / 4) * (P.W / 4) + px / 4]; }
    int16_t level(int c, int x, int y) const { return P.coef[c][(size_t)y * P.orig[c].w + x]; }

    // ---- residual_coding, ctu_encoder.rs:1786-2269 (dep_quant on, transform skip never chosen)
    void encode_residual(int c_idx, int x0, int y0, int log2) {
        const int n = 1 << log2, nsb = n / 4;
        const uint16_t *scan = scan_order(log2);  // forward scan index -> (y << 8) | x
        auto qc = [&](int x, int y) { return (int)level(c_idx, x0 + x, y0 + y); };
        std::vector<int> abs_level((size_t)n * n, 0), abs_level_pass1((size_t)n * n, 0);
        // get_last_sig_coeff_pos (ctu.rs:867-899)
        int last_k = n * n - 1;
        while (last_k > 0 && qc(scan[last_k] & 255, scan[last_k] >> 8) == 0) last_k--;
        const int last_x = scan[last_k] & 255, last_y = scan[last_k] >> 8;
        auto split_prefix = [](int v, int &prefix, int &suffix) {  // ctu_encoder.rs:1818-1849
            if (v <= 3) {
                prefix = v;
                suffix = 0;
                return;
            }
            int pre, suf, suffix_bits = 1;
            for (;;) {
                pre = v >> suffix_bits;
                suf = v - (pre << suffix_bits);
                if (pre < 4) break;
                suffix_bits++;
            }
            prefix = ((suffix_bits + 1) << 1) + (pre & 1);
            suffix = suf;
        };
        int xp, xsuf, yp, ysuf;
        split_prefix(last_x, xp, xsuf);
        split_prefix(last_y, yp, ysuf);
        auto last_prefix = [&](int base, int prefix) {  // TR cMax = 2*log2 - 1; ctxInc bool_coder.rs:2053-2083
            static const int OFFSET_Y[6] = {0, 0, 3, 6, 10, 15};
            int ctx_offset, ctx_shift;
            if (c_idx == 0) {
                ctx_offset = OFFSET_Y[log2 - 1];
                ctx_shift = (log2 + 1) >> 2;
            } else {
                ctx_offset = 20;
                ctx_shift = std::min(2, std::max(0, (1 << log2) >> 3));
            }
            int c_max = (log2 << 1) - 1;
            for (int bin_idx = 0; bin_idx < prefix; bin_idx++) c.decision(base + (bin_idx >> ctx_shift) + ctx_offset, true);
            if (prefix < c_max) c.decision(base + (prefix >> ctx_shift) + ctx_offset, false);
        };
        last_prefix(CTX_LAST_X, xp);
        last_prefix(CTX_LAST_Y, yp);
        if (xp > 3) c.bypass_fixed((unsigned)xsuf, (xp >> 1) - 1);
        if (yp > 3) c.bypass_fixed((unsigned)ysuf, (yp >> 1) - 1);
        int rem_bins_pass1 = ((1 << (2 * log2)) * 7) >> 2;
        const int last_subblock = last_k / 16, last_scan_pos = last_k % 16;
        if ((last_subblock > 0 || last_scan_pos > 0) && c_idx == 0) mts_dc_only = false;
        auto sb_coded = [&](int xs, int ys) {  // tu.get_sb_coded_flag
            for (int yy = 0; yy < 4; yy++)
                for (int xx = 0; xx < 4; xx++)
                    if (qc(xs * 4 + xx, ys * 4 + yy) != 0) return true;
            return false;
        };
        auto template_sum = [&](const std::vector<int> &a, int x, int y, int *num_sig) {  // bool_coder.rs:2151-2246 / 1133-1174
            int sum = 0, num = 0;
            auto add = [&](int xx, int yy) {
                int v = a[(size_t)yy * n + xx];
                sum += v;
                num += std::min((int)(qc(xx, yy) != 0), v);
            };
            if (x < n - 1) {
                add(x + 1, y);
                if (x < n - 2) add(x + 2, y);
                if (y < n - 1) add(x + 1, y + 1);
            }
            if (y < n - 1) {
                add(x, y + 1);
                if (y < n - 2) add(x, y + 2);
            }
            if (num_sig) *num_sig = num;
            return sum;
        };
        auto rice_coded = [&](int value, int c_rice_param) {  // encode_abs_remainder / encode_dec_abs_level, bool_coder.rs:1384-1465
            int c_max = 6 << c_rice_param;
            int prefix_val = std::min(c_max, value);
            std::vector<bool> bins;
            int pv = prefix_val >> c_rice_param;
            if (pv < (c_max >> c_rice_param)) {
                bins.assign(pv, true);
                bins.push_back(false);
            } else {
                bins.assign(c_max >> c_rice_param, true);
            }
            if (c_max > prefix_val && c_rice_param > 0) {
                int suffix_val = prefix_val - (pv << c_rice_param);
                for (int i = c_rice_param - 1; i >= 0; i--) bins.push_back(((suffix_val >> i) & 1) != 0);
            }
            bool all = bins.size() == 6;
            for (bool b : bins) all = all && b;
            if (all) {  // limited k-th order exp-Golomb, bool_coder.rs:1305-1331
                int symbol_val = value - c_max, k = c_rice_param + 1;
                int code_value = symbol_val >> k, pre_ext_len = 0;
                while (pre_ext_len < 11 && code_value > (2 << pre_ext_len) - 2) {
                    pre_ext_len++;
                    bins.push_back(true);
                }
                int escape_length;
                if (pre_ext_len == 11) escape_length = 15;
                else {
                    bins.push_back(false);
                    escape_length = pre_ext_len + k;
                }
                symbol_val -= ((1 << pre_ext_len) - 1) << k;
                while (escape_length > 0) {
                    escape_length--;
                    bins.push_back(((symbol_val >> escape_length) & 1) == 1);
                }
            }
            for (bool b : bins) c.bypass(b);
        };
        int q_state = 0;
        std::vector<std::pair<int, int>> sb_order((size_t)nsb * nsb);
        for (int i = 0; i < nsb * nsb; i++) sb_order[i] = {(scan[i * 16] & 255) / 4, (scan[i * 16] >> 8) / 4};
        for (int i = last_subblock; i >= 0; i--) {
            const int x_s = sb_order[i].first, y_s = sb_order[i].second;
            int abs_levels[16];
            {
                int st = q_state;
                for (int nn = 15; nn >= 0; nn--) {
                    int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                    int v = std::abs(qc(x, y));
                    if (v != 0 && (v & 1) != (st > 1)) {
                        fprintf(stderr, "oracle: level parity does not match the dep-quant state (the reference asserts, ctu_encoder.rs:1966-1971)\n");
                        abort();
                    }
                    abs_levels[nn] = (v + (st > 1)) / 2;
                    st = TRS[st][abs_levels[nn] & 1];
                }
            }
            bool infer_sb_dc_sig_coeff_flag = false;
            const bool sb_coded_flag = sb_coded(x_s, y_s) || (x_s == 0 && y_s == 0);
            if (i < last_subblock && i > 0) {
                int csbf_ctx = 0;  // bool_coder.rs:2102-2149
                if (x_s < nsb - 1) csbf_ctx += sb_coded(x_s + 1, y_s);
                if (y_s < nsb - 1) csbf_ctx += sb_coded(x_s, y_s + 1);
                c.decision(CTX_SB_CODED + (c_idx == 0 ? std::min(csbf_ctx, 1) : 2 + std::min(csbf_ctx, 1)), sb_coded_flag);
                infer_sb_dc_sig_coeff_flag = true;
            }
            if (sb_coded_flag && (x_s > 3 || y_s > 3) && c_idx == 0) mts_zero_out_sig_coeff_flag = false;
            const int first_pos_mode0 = i == last_subblock ? last_scan_pos : 15;
            int first_pos_mode1 = first_pos_mode0;
            for (int nn = first_pos_mode0; nn >= 0; nn--) {
                if (rem_bins_pass1 < 4) break;
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                const bool at_last = x == last_x && y == last_y;
                const bool sig_coeff_flag = qc(x, y) != 0 || at_last || ((x & 3) == 0 && (y & 3) == 0 && infer_sb_dc_sig_coeff_flag && sb_coded_flag);
                if (sb_coded_flag && (nn > 0 || !infer_sb_dc_sig_coeff_flag) && !at_last) {
                    int loc_sum = template_sum(abs_level_pass1, x, y, nullptr);
                    int d = x + y, ctx_inc;  // bool_coder.rs:2248-2290
                    if (c_idx == 0) ctx_inc = 12 * std::max(0, q_state - 1) + std::min((loc_sum + 1) >> 1, 3) + (d < 2 ? 8 : (d < 5 ? 4 : 0));
                    else ctx_inc = 36 + 8 * std::max(0, q_state - 1) + std::min((loc_sum + 1) >> 1, 3) + (d < 2 ? 4 : 0);
                    c.decision(CTX_SIG + ctx_inc, sig_coeff_flag);
                    rem_bins_pass1--;
                    if (sig_coeff_flag) infer_sb_dc_sig_coeff_flag = false;
                }
                const int al = abs_levels[nn];
                const bool gtx0 = al > 1, gtx1 = al > 3, par = al > 1 && al % 2 == 1;
                if (sig_coeff_flag) {
                    int num_sig = 0;
                    int loc_sum = template_sum(abs_level_pass1, x, y, &num_sig);
                    int ctx_offset = std::min(loc_sum - num_sig, 4), d = x + y, ctx_inc;  // bool_coder.rs:2292-2371
                    if (at_last) ctx_inc = c_idx == 0 ? 0 : 21;
                    else if (c_idx == 0) ctx_inc = 1 + ctx_offset + (d == 0 ? 15 : (d < 3 ? 10 : (d < 10 ? 5 : 0)));
                    else ctx_inc = 22 + ctx_offset + (d == 0 ? 5 : 0);
                    c.decision(CTX_GTX + ctx_inc, gtx0);
                    rem_bins_pass1--;
                    if (gtx0) {
                        c.decision(CTX_PAR + ctx_inc, par);
                        rem_bins_pass1--;
                        c.decision(CTX_GTX + ctx_inc + 32, gtx1);
                        rem_bins_pass1--;
                    }
                }
                const int pass1 = (int)sig_coeff_flag + (int)par + (int)gtx0 + 2 * (int)gtx1;
                abs_level_pass1[(size_t)y * n + x] = pass1;
                assert((pass1 & 1) == (al & 1));
                q_state = TRS[q_state][pass1 & 1];
                first_pos_mode1 = nn - 1;
            }
            for (int nn = first_pos_mode0; nn > first_pos_mode1; nn--) {
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                int abs_remainder = 0;
                if (abs_levels[nn] > 3) {
                    abs_remainder = (abs_levels[nn] - abs_level_pass1[(size_t)y * n + x]) / 2;
                    int loc_sum_abs = std::min(31, std::max(0, template_sum(abs_level, x, y, nullptr) - 4 * 5));
                    rice_coded(abs_remainder, kRice[loc_sum_abs]);
                }
                abs_level[(size_t)y * n + x] = abs_level_pass1[(size_t)y * n + x] + 2 * abs_remainder;
                assert(abs_level[(size_t)y * n + x] == abs_levels[nn]);
            }
            for (int nn = first_pos_mode1; nn >= 0; nn--) {
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                abs_level[(size_t)y * n + x] = abs_levels[nn];
                if (sb_coded_flag) {
                    int loc_sum_abs = std::min(31, std::max(0, template_sum(abs_level, x, y, nullptr)));
                    int c_rice_param = kRice[loc_sum_abs];
                    int zero_pos = (q_state < 2 ? 1 : 2) << c_rice_param;  // ctu.rs:739-782
                    int v = abs_levels[nn];
                    int dec_abs_level = v == 0 ? zero_pos : (zero_pos >= v ? v - 1 : v);
                    rice_coded(dec_abs_level, c_rice_param);
                }
                q_state = TRS[q_state][abs_levels[nn] & 1];
            }
            for (int nn = 15; nn >= 0; nn--)  // coeff_sign_flag, bypass; sign data hiding is off
                if (abs_levels[nn] > 0) c.bypass(qc(scan[i * 16 + nn] & 255, scan[i * 16 + nn] >> 8) < 0);
        }
    }

    // ---- intra luma mode syntax with the final-tree neighbours (ctu.rs:1498-1635, ctu_encoder.rs:755-804)
    void encode_luma_mode(int px, int py, int size) {
        const int mode = luma_mode_at(px, py);
        if (mode == MODE_PLANAR) {
            c.decision(CTX_MPM_FLAG, true);
            c.decision(CTX_NOT_PLANAR + 1, false);
            return;
        }
        int left = px - 1 >= 0 ? luma_mode_at(px - 1, py + size - 1) : MODE_PLANAR;
        int above = (py - 1 >= 0 && !(py - 1 < (py / 32) * 32)) ? luma_mode_at(px + size - 1, py - 1) : MODE_PLANAR;
        std::vector<int> cand;
        if (left == above && left > MODE_DC) cand = {left, 2 + (left + 61) % 64, 2 + (left - 1) % 64, 2 + (left + 60) % 64, 2 + left % 64};
        else if (left != above && (left > MODE_DC || above > MODE_DC)) {
            int mn = std::min(left, above), mx = std::max(left, above);
            if (mn > MODE_DC) {
                int d = mx - mn;
                if (d == 1) cand = {left, above, 2 + (mn + 61) % 64, 2 + (mx - 1) % 64, 2 + (mn + 60) % 64};
                else if (d >= 62) cand = {left, above, 2 + (mn - 1) % 64, 2 + (mx + 61) % 64, 2 + mn % 64};
                else if (d == 2) cand = {left, above, 2 + (mn - 1) % 64, 2 + (mn + 61) % 64, 2 + (mx - 1) % 64};
                else cand = {left, above, 2 + (mn + 61) % 64, 2 + (mn - 1) % 64, 2 + (mx + 61) % 64};
            } else cand = {mx, 2 + (mx + 61) % 64, 2 + (mx - 1) % 64, 2 + (mx + 60) % 64, 2 + mx % 64};
        } else cand = {MODE_DC, 50, 18, 46, 54};
        int pos = -1;
        for (int i = 0; i < 5 && pos < 0; i++)
            if (cand[i] == mode) pos = i;
        if (pos >= 0) {
            c.decision(CTX_MPM_FLAG, true);
            c.decision(CTX_NOT_PLANAR + 1, true);
            for (int i = 0; i < pos; i++) c.bypass(true);  // intra_luma_mpm_idx: TR cMax 4, bypass
            if (pos < 4) c.bypass(false);
        } else {
            c.decision(CTX_MPM_FLAG, false);
            std::sort(cand.begin(), cand.end());
            int rem = mode - 1;
            for (int j = 4; j >= 0; j--)
                if (mode > cand[j]) {
                    rem = mode - (j + 2);
                    break;
                }
            // truncated binary cMax 60 (bool_coder.rs:1246-1255)
            int nsym = 61, k = 5, u = (1 << (k + 1)) - nsym;
            if (rem < u) c.bypass_fixed((unsigned)rem, k);
            else c.bypass_fixed((unsigned)(rem + u), k + 1);
        }
    }

    bool any_level(int c_idx, int x0, int y0, int n) const {
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++)
                if (level(c_idx, x0 + x, y0 + y) != 0) return true;
        return false;
    }

    void encode_coding_unit(int px, int py, int size, int tree) {
        if (tree != DUAL_TREE_CHROMA) encode_luma_mode(px, py, size);
        if (tree != DUAL_TREE_LUMA) {  // ctu_encoder.rs:806-874; is_cclm_enabled is true here (ctu.rs:1383-1409)
            const CtuRecord &r = record_at(px, py);
            int cm = r.chroma_mode[((py % 32) / 8) * 4 + (px % 32) / 8];
            bool cclm_mode_flag = cm >= MODE_LT_CCLM;
            c.decision(CTX_CCLM_FLAG, cclm_mode_flag);
            if (cclm_mode_flag) {
                int idx = cm - MODE_LT_CCLM;  // TR cMax 2: first bin context 0, second bypass
                c.decision(CTX_CCLM_IDX, idx > 0);
                if (idx > 0) c.bypass(idx > 1);
            } else {
                c.decision(CTX_CHROMA_PRED, false);  // intra_chroma_pred_mode == 4: single bin 0 (bool_coder.rs:1333-1346)
            }
        }
        mts_dc_only = true;  // ctu_encoder.rs:1213-1216
        mts_zero_out_sig_coeff_flag = true;
        // transform_unit, ctu_encoder.rs:1542-1782
        const int log2 = ilog2_(size);
        bool y_coded = tree != DUAL_TREE_CHROMA && any_level(0, px, py, size);
        bool cb_coded = tree != DUAL_TREE_LUMA && any_level(1, px / 2, py / 2, size / 2);
        bool cr_coded = tree != DUAL_TREE_LUMA && any_level(2, px / 2, py / 2, size / 2);
        if (tree == SINGLE_TREE || tree == DUAL_TREE_CHROMA) {
            c.decision(CTX_TU_CB + 0, cb_coded);
            c.decision(CTX_TU_CR + (cb_coded ? 1 : 0), cr_coded);
        }
        if (tree == SINGLE_TREE || tree == DUAL_TREE_LUMA) c.decision(CTX_TU_Y + 0, y_coded);
        bool chroma_available = tree != DUAL_TREE_LUMA;
        if ((y_coded || (chroma_available && (cb_coded || cr_coded))) && tree != DUAL_TREE_CHROMA && !is_cu_qp_delta_coded) {
            c.decision(CTX_QP_DELTA_ABS + 0, false);  // cu_qp_delta_abs = 0: TR prefix "0"
            is_cu_qp_delta_coded = true;
        }
        if (y_coded) {
            c.decision(CTX_TS_FLAG + 0, false);
            encode_residual(0, px, py, log2);
        }
        if (cb_coded) {
            c.decision(CTX_TS_FLAG + 1, false);
            encode_residual(1, px / 2, py / 2, log2 - 1);
        }
        if (cr_coded) {
            c.decision(CTX_TS_FLAG + 1, false);
            encode_residual(2, px / 2, py / 2, log2 - 1);
        }
        if (tree != DUAL_TREE_CHROMA && mts_zero_out_sig_coeff_flag && !mts_dc_only) c.decision(CTX_MTS + 0, false);  // ctu_encoder.rs:1299-1318
    }
    static int ilog2_(int v) {
        int l = 0;
        while ((1 << (l + 1)) <= v) l++;
        return l;
    }

    void encode_coding_tree(int px, int py, int size) {
        const CtuRecord &r = record_at(px, py);
        const int leaf = ct_size_at(px, py);
        const bool split_cu_flag = leaf < size;
        if (size > 4) {  // allow_split_qt (encoder_context.rs:958-971); BT/TT never allowed
            bool cond_l = px - 1 >= 0 && ct_size_at(px - 1, py) < size;  // bool_coder.rs:2659-2744
            bool cond_a = py - 1 >= 0 && ct_size_at(px, py - 1) < size;
            c.decision(CTX_SPLIT_CU + (int)cond_l + (int)cond_a, split_cu_flag);
        }
        (void)r;
        if (!split_cu_flag) {
            encode_coding_unit(px, py, size, SINGLE_TREE);
            return;
        }
        if (size == 8) {  // local dual tree (ctu.rs:2031-2055): four luma CUs, then the chroma CU
            for (int i = 0; i < 4; i++) encode_coding_unit(px + (i % 2) * 4, py + (i / 2) * 4, 4, DUAL_TREE_LUMA);
            encode_coding_unit(px, py, 8, DUAL_TREE_CHROMA);
            return;
        }
        for (int i = 0; i < 4; i++) encode_coding_tree(px + (i % 2) * (size / 2), py + (i / 2) * (size / 2), size / 2);
    }

    static std::vector<uint8_t> finish(BoolCoder &c) {
        c.stop_one_bit();  // end_of_slice_one_bit (slice_encoder.rs:380-388)
        std::vector<uint8_t> out((c.bits.size() + 7) / 8, 0);  // bins.byte_align(): zero padding (slice_encoder.rs:419)
        for (size_t i = 0; i < c.bits.size(); i++)
            if (c.bits[i]) out[i / 8] |= (uint8_t)(0x80u >> (i % 8));
        return out;
    }

    std::vector<uint8_t> run(std::vector<uint16_t> *trace = nullptr) {
        c.trace = trace;
        c.init(K.qp);  // first CTU of the picture: ctu_encoder.rs:38-47
        for (int cy = 0; cy < P.H; cy += 32)
            for (int cx = 0; cx < P.W; cx += 32) {
                is_cu_qp_delta_coded = false;  // quantisation group = CTU (ctu_encoder.rs:305-310)
                encode_coding_tree(cx, cy, 32);
            }
        c.trace = nullptr;
        return finish(c);
    }
};

}  // namespace

// Test hooks for the product's arithmetic coder (wrenc_b200/csrc/cabac_engine.cuh): the bin string of a searched picture in
// the product's 16-bit entry format, and the reference engine run over an arbitrary bin string (+ end_of_slice_one_bit and
// byte alignment, as SliceWriter::run does).
std::vector<uint8_t> code_slice_data_traced(const Consts &k, Picture &p, std::vector<uint16_t> &bins) {
    SliceWriter w(k, p);
    bins.clear();
    return w.run(&bins);
}
std::vector<uint8_t> code_bin_string(int slice_qp, const uint16_t *entries, size_t n) {
    BoolCoder c;
    c.init(slice_qp);
    for (size_t i = 0; i < n; i++) {
        const unsigned e = entries[i];
        if (e & 1024u) c.bypass(((e >> 9) & 1) != 0);
        else c.decision((int)(e & 511u), ((e >> 9) & 1) != 0);
    }
    return SliceWriter::finish(c);
}

std::vector<uint8_t> code_slice_data(const Consts &k, Picture &p) {
    SliceWriter w(k, p);
    return w.run();
}

}  // namespace wo
