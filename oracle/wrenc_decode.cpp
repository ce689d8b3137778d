// wrenc_decode.cpp — CPU ORACLE self-check (test infrastructure, NOT product code; see wrenc_oracle.hpp header).
//
// A minimal VVC intra decoder for exactly the subset the all-intra path emits: it parses a picture's
// slice_data() with a standard CABAC *decoding* engine (VVC 9.3.4.3: 9-bit offset, range 510, two-window probability states),
// rebuilds the coding trees, modes and transform coefficient levels, and reconstructs the picture with the oracle's block
// operations (prediction, dequantisation, inverse DCT).  What it pins that nothing else here can: the bitstream is a
// well-formed arithmetic code, every context index the encoder used is derivable by a decoder from already decoded data,
// the dependent-quantisation state machine / MPM / chroma DM derivations are symmetric, and the decoded reconstruction
// equals the encoder's (--reconst) — the stand-in for the reference's VTM integration test (scripts/intergration_test.sh).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "wrenc_oracle.hpp"
#include "cabac_tables.inc"

namespace wo {
namespace {

static const int TRD[4][2] = {{0, 2}, {2, 0}, {1, 3}, {3, 1}};
static const int kRiceD[32] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3};

struct BitReader {
    const uint8_t *d;
    size_t n, pos = 0;  // pos in bits
    int bit() {
        if (pos >= n * 8) return 0;  // reading past the end returns zeros (the trailing alignment bits are zeros anyway)
        int b = (d[pos >> 3] >> (7 - (pos & 7))) & 1;
        pos++;
        return b;
    }
};

struct CabacDecoder {
    BitReader br;
    uint16_t p[CTX_TOTAL][2];
    unsigned range = 510, offset = 0;
    void init(int qp, const uint8_t *data, size_t len) {
        br.d = data;
        br.n = len;
        br.pos = 0;
        for (int i = 0; i < CTX_TOTAL; i++) {
            int iv = kCabacInitValue[i], m = (iv >> 3) - 4, nn = (iv & 7) * 18 + 1;
            int pre = std::min(127, std::max(1, ((m * (std::min(63, std::max(0, qp)) - 16)) >> 1) + nn));
            p[i][0] = (uint16_t)(pre << 3);
            p[i][1] = (uint16_t)(pre << 7);
        }
        range = 510;
        offset = 0;
        for (int i = 0; i < 9; i++) offset = (offset << 1) | (unsigned)br.bit();
    }
    int decision(int ctx) {
        unsigned q = range >> 5, ps = p[ctx][1] + 16u * p[ctx][0], mps = ps >> 14;
        unsigned lps = ((q * ((mps ? 32767u - ps : ps) >> 9)) >> 1) + 4;
        range -= lps;
        int bin;
        if (offset >= range) {
            bin = !mps;
            offset -= range;
            range = lps;
        } else bin = (int)mps;
        while (range < 256) {
            range <<= 1;
            offset = (offset << 1) | (unsigned)br.bit();
        }
        int si = kCabacShiftIdx[ctx], s0 = (si >> 2) + 2, s1 = (si & 3) + 3 + s0;
        p[ctx][0] = (uint16_t)(p[ctx][0] - (p[ctx][0] >> s0) + ((1023 * bin) >> s0));
        p[ctx][1] = (uint16_t)(p[ctx][1] - (p[ctx][1] >> s1) + ((16383 * bin) >> s1));
        return bin;
    }
    int bypass() {
        offset = (offset << 1) | (unsigned)br.bit();
        if (offset >= range) {
            offset -= range;
            return 1;
        }
        return 0;
    }
    unsigned bypass_bits(int n) {
        unsigned v = 0;
        for (int i = 0; i < n; i++) v = (v << 1) | (unsigned)bypass();
        return v;
    }
    int terminate() {
        range -= 2;
        if (offset >= range) return 1;
        while (range < 256) {
            range <<= 1;
            offset = (offset << 1) | (unsigned)br.bit();
        }
        return 0;
    }
};

struct Decoder {
    Consts K;
    Picture P;  // rec / coef / mode_map / records are filled while decoding
    CabacDecoder c;
    std::vector<uint8_t> size_map;  // luma CU size per 4x4 block (split_cu_flag contexts)
    bool is_cu_qp_delta_coded = false, mts_dc_only = true, mts_zero_out = true;

    int size_at(int px, int py) const { return size_map[(size_t)(py / 4) * (P.W / 4) + px / 4]; }
    int mode_at(int px, int py) const { return P.mode_map[(size_t)(py / 4) * (P.W / 4) + px / 4]; }
    int16_t &lev(int ci, int x, int y) { return P.coef[ci][(size_t)y * P.orig[ci].w + x]; }

    static void fail(const char *what) { throw std::runtime_error(what); }

    int rice_value(int c_rice_param) {  // inverse of abs_remainder / dec_abs_level binarisation
        int prefix = 0;
        while (prefix < 6 && c.bypass()) prefix++;
        if (prefix < 6) return (prefix << c_rice_param) + (int)c.bypass_bits(c_rice_param);
        int k = c_rice_param + 1, pre_ext_len = 0;
        while (pre_ext_len < 11 && c.bypass()) pre_ext_len++;
        int escape_length = pre_ext_len == 11 ? 15 : pre_ext_len + k;
        int v = (int)c.bypass_bits(escape_length) + (((1 << pre_ext_len) - 1) << k);
        return (6 << c_rice_param) + v;
    }

    void decode_residual(int c_idx, int x0, int y0, int log2) {
        const int n = 1 << log2, nsb = n / 4;
        const uint16_t *scan = scan_order(log2);
        std::vector<int> abs_level((size_t)n * n, 0), pass1((size_t)n * n, 0);
        std::vector<uint8_t> sbc((size_t)nsb * nsb, 0);
        auto last_prefix = [&](int base) {
            static const int OFFSET_Y[6] = {0, 0, 3, 6, 10, 15};
            int off, shift;
            if (c_idx == 0) { off = OFFSET_Y[log2 - 1]; shift = (log2 + 1) >> 2; }
            else { off = 20; shift = std::min(2, std::max(0, n >> 3)); }
            int c_max = (log2 << 1) - 1, v = 0;
            while (v < c_max && c.decision(base + (v >> shift) + off)) v++;
            return v;
        };
        int xp = last_prefix(CTX_LAST_X), yp = last_prefix(CTX_LAST_Y);
        auto finish = [&](int prefix) {
            if (prefix <= 3) return prefix;
            int bits = (prefix >> 1) - 1;
            int suffix = (int)c.bypass_bits(bits);
            return (1 << bits) * (2 + (prefix & 1)) + suffix;  // VVC 7.4.12.11: LastSignificantCoeffX
        };
        const int last_x = finish(xp), last_y = finish(yp);
        if (last_x >= n || last_y >= n) fail("last significant position outside the block");
        int last_k = -1;
        for (int k = 0; k < n * n; k++)
            if ((scan[k] & 255) == last_x && (scan[k] >> 8) == last_y) last_k = k;
        int rem_bins = ((n * n) * 7) >> 2;
        const int last_sb = last_k / 16, last_pos = last_k % 16;
        if ((last_sb > 0 || last_pos > 0) && c_idx == 0) mts_dc_only = false;
        auto tsum = [&](const std::vector<int> &a, int x, int y, int *num) {
            int s = 0, nz = 0;
            auto add = [&](int xx, int yy) { int v = a[(size_t)yy * n + xx]; s += v; nz += v > 0; };
            if (x < n - 1) { add(x + 1, y); if (x < n - 2) add(x + 2, y); if (y < n - 1) add(x + 1, y + 1); }
            if (y < n - 1) { add(x, y + 1); if (y < n - 2) add(x, y + 2); }
            if (num) *num = nz;
            return s;
        };
        int q_state = 0;
        for (int i = last_sb; i >= 0; i--) {
            const int xs = (scan[i * 16] & 255) / 4, ys = (scan[i * 16] >> 8) / 4;
            bool coded = true, infer = false;
            if (i < last_sb && i > 0) {
                int csbf = 0;
                if (xs < nsb - 1) csbf += sbc[(size_t)ys * nsb + xs + 1];
                if (ys < nsb - 1) csbf += sbc[(size_t)(ys + 1) * nsb + xs];
                coded = c.decision(CTX_SB_CODED + (c_idx == 0 ? std::min(csbf, 1) : 2 + std::min(csbf, 1))) != 0;
                infer = true;
            }
            sbc[(size_t)ys * nsb + xs] = coded;
            if (coded && (xs > 3 || ys > 3) && c_idx == 0) mts_zero_out = false;
            const int start_state = q_state;
            int al[16] = {0};
            const int fp0 = i == last_sb ? last_pos : 15;
            int fp1 = fp0;
            for (int nn = fp0; nn >= 0; nn--) {
                if (rem_bins < 4) break;
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                const bool at_last = x == last_x && y == last_y;
                bool sig;
                if (coded && (nn > 0 || !infer) && !at_last) {
                    int s = tsum(pass1, x, y, nullptr), d = x + y, ci;
                    if (c_idx == 0) ci = 12 * std::max(0, q_state - 1) + std::min((s + 1) >> 1, 3) + (d < 2 ? 8 : (d < 5 ? 4 : 0));
                    else ci = 36 + 8 * std::max(0, q_state - 1) + std::min((s + 1) >> 1, 3) + (d < 2 ? 4 : 0);
                    sig = c.decision(CTX_SIG + ci) != 0;
                    rem_bins--;
                    if (sig) infer = false;
                } else sig = coded && (at_last || (nn == 0 && infer));
                int gt1 = 0, par = 0, gt3 = 0;
                if (sig) {
                    int num = 0, s = tsum(pass1, x, y, &num), d = x + y, ci;
                    if (at_last) ci = c_idx == 0 ? 0 : 21;
                    else if (c_idx == 0) ci = 1 + std::min(s - num, 4) + (d == 0 ? 15 : (d < 3 ? 10 : (d < 10 ? 5 : 0)));
                    else ci = 22 + std::min(s - num, 4) + (d == 0 ? 5 : 0);
                    gt1 = c.decision(CTX_GTX + ci);
                    rem_bins--;
                    if (gt1) {
                        par = c.decision(CTX_PAR + ci);
                        rem_bins--;
                        gt3 = c.decision(CTX_GTX + ci + 32);
                        rem_bins--;
                    }
                }
                const int p1 = (int)sig + par + gt1 + 2 * gt3;
                pass1[(size_t)y * n + x] = p1;
                al[nn] = p1;
                q_state = TRD[q_state][p1 & 1];
                fp1 = nn - 1;
            }
            for (int nn = fp0; nn > fp1; nn--) {
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                if (al[nn] >= 4 && (pass1[(size_t)y * n + x] >> 1) >= 2) {  // gt3 set: sig + par + gt1 + 2 = 4 or 5
                    int s = std::min(31, std::max(0, tsum(abs_level, x, y, nullptr) - 20));
                    al[nn] += 2 * rice_value(kRiceD[s]);
                }
                abs_level[(size_t)y * n + x] = al[nn];
            }
            for (int nn = fp1; nn >= 0; nn--) {
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                if (coded) {
                    int s = std::min(31, std::max(0, tsum(abs_level, x, y, nullptr)));
                    int rice = kRiceD[s], zero_pos = (q_state < 2 ? 1 : 2) << rice;
                    int dec = rice_value(rice);
                    al[nn] = dec == zero_pos ? 0 : (dec < zero_pos ? dec + 1 : dec);
                }
                abs_level[(size_t)y * n + x] = al[nn];
                q_state = TRD[q_state][al[nn] & 1];
            }
            // signs, then TransCoeffLevel through the dependent-quantisation state machine
            int sign[16] = {0};
            for (int nn = 15; nn >= 0; nn--)
                if (al[nn] > 0) sign[nn] = c.bypass();
            int st = start_state;
            for (int nn = 15; nn >= 0; nn--) {
                const int x = scan[i * 16 + nn] & 255, y = scan[i * 16 + nn] >> 8;
                if (al[nn] > 0) {
                    int v = 2 * al[nn] - (st > 1 ? 1 : 0);
                    lev(c_idx, x0 + x, y0 + y) = (int16_t)(sign[nn] ? -v : v);
                }
                st = TRD[st][al[nn] & 1];
            }
            if (st != q_state && i < last_sb) fail("dependent quantisation state desynchronised");
            q_state = st;
        }
    }

    int decode_luma_mode(int px, int py, int size) {
        if (c.decision(CTX_MPM_FLAG)) {
            if (!c.decision(CTX_NOT_PLANAR + 1)) return MODE_PLANAR;
            int idx = 0;
            while (idx < 4 && c.bypass()) idx++;
            int cand[5];
            mpm(px, py, size, cand);
            return cand[idx];
        }
        int v = (int)c.bypass_bits(5);
        if (v >= 3) v = ((v << 1) | c.bypass()) - 3;  // truncated binary, cMax 60
        int cand[5];
        mpm(px, py, size, cand);
        std::sort(cand, cand + 5);
        int mode = v + 1;  // VVC 8.4.2: IntraPredModeY = remainder + 1, then + 1 for every candidate <= it (ascending)
        for (int i = 0; i < 5; i++)
            if (mode >= cand[i]) mode++;
        return mode;
    }
    void mpm(int px, int py, int size, int cand[5]) const {
        int left = px - 1 >= 0 ? mode_at(px - 1, py + size - 1) : MODE_PLANAR;
        int above = (py - 1 >= 0 && (py % 32) != 0) ? mode_at(px + size - 1, py - 1) : MODE_PLANAR;
        auto set = [&](int a, int b, int cc, int d, int e) { cand[0] = a; cand[1] = b; cand[2] = cc; cand[3] = d; cand[4] = e; };
        if (left == above && left > MODE_DC) set(left, 2 + (left + 61) % 64, 2 + (left - 1) % 64, 2 + (left + 60) % 64, 2 + left % 64);
        else if (left != above && (left > MODE_DC || above > MODE_DC)) {
            int mn = std::min(left, above), mx = std::max(left, above);
            if (mn > MODE_DC) {
                int d = mx - mn;
                if (d == 1) set(left, above, 2 + (mn + 61) % 64, 2 + (mx - 1) % 64, 2 + (mn + 60) % 64);
                else if (d >= 62) set(left, above, 2 + (mn - 1) % 64, 2 + (mx + 61) % 64, 2 + mn % 64);
                else if (d == 2) set(left, above, 2 + (mn - 1) % 64, 2 + (mn + 61) % 64, 2 + (mx - 1) % 64);
                else set(left, above, 2 + (mn + 61) % 64, 2 + (mn - 1) % 64, 2 + (mx + 61) % 64);
            } else set(mx, 2 + (mx + 61) % 64, 2 + (mx - 1) % 64, 2 + (mx + 60) % 64, 2 + mx % 64);
        } else set(MODE_DC, 50, 18, 46, 54);
    }

    void reconstruct(const TU &tu, int c_idx, bool coded) {
        const int cs = c_idx != 0, n = tu.w >> cs, cx = tu.x >> cs, cy = tu.y >> cs;
        int l2 = 0;
        while ((1 << (l2 + 1)) <= n) l2++;
        std::vector<uint8_t> pred((size_t)n * n);
        predict(P, tu, c_idx, pred.data());
        std::vector<int16_t> q((size_t)n * n, 0), d((size_t)n * n, 0), r((size_t)n * n, 0);
        if (coded) {
            for (int y = 0; y < n; y++)
                for (int x = 0; x < n; x++) q[(size_t)y * n + x] = lev(c_idx, cx + x, cy + y);
            dequantize(K, q.data(), l2, d.data());
            inv_dct(d.data(), l2, r.data());
        }
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                int v = (int)pred[(size_t)y * n + x] + r[(size_t)y * n + x];
                P.rec[c_idx].at(cx + x, cy + y) = (uint8_t)std::min(255, std::max(0, v));
            }
    }

    void decode_cu(int px, int py, int size, int tree, bool ar, bool bl, CtuRecord &rec, int ctu_x, int ctu_y) {
        int luma_mode = 0;
        if (tree != DUAL_TREE_CHROMA) {
            luma_mode = decode_luma_mode(px, py, size);
            for (int y = 0; y < size; y += 4)
                for (int x = 0; x < size; x += 4) {
                    P.mode_map[(size_t)((py + y) / 4) * (P.W / 4) + (px + x) / 4] = (uint8_t)luma_mode;
                    rec.luma_mode[((py - ctu_y + y) / 4) * 8 + (px - ctu_x + x) / 4] = (uint8_t)luma_mode;
                }
        }
        int chroma_mode = 0;
        if (tree != DUAL_TREE_LUMA) {
            if (c.decision(CTX_CCLM_FLAG)) {
                int idx = c.decision(CTX_CCLM_IDX);
                if (idx) idx += c.bypass();
                chroma_mode = MODE_LT_CCLM + idx;
            } else {
                if (c.decision(CTX_CHROMA_PRED)) fail("intra_chroma_pred_mode != 4 is never emitted");
                // DM: luma mode at the centre of the chroma block (VVC 8.4.3); for SINGLE_TREE that is the CU's own mode
                chroma_mode = tree == SINGLE_TREE ? luma_mode : mode_at(px + size / 2, py + size / 2);
            }
            for (int y = 0; y < size; y += 8)
                for (int x = 0; x < size; x += 8) rec.chroma_mode[((py - ctu_y + y) / 8) * 4 + (px - ctu_x + x) / 8] = (uint8_t)chroma_mode;
        }
        mts_dc_only = true;
        mts_zero_out = true;
        int l2 = 0;
        while ((1 << (l2 + 1)) <= size) l2++;
        bool cb = false, cr = false, yc = false;
        if (tree != DUAL_TREE_LUMA) {
            cb = c.decision(CTX_TU_CB) != 0;
            cr = c.decision(CTX_TU_CR + (cb ? 1 : 0)) != 0;
        }
        if (tree != DUAL_TREE_CHROMA) yc = c.decision(CTX_TU_Y) != 0;
        if ((yc || cb || cr) && tree != DUAL_TREE_CHROMA && !is_cu_qp_delta_coded) {
            if (c.decision(CTX_QP_DELTA_ABS)) fail("cu_qp_delta_abs != 0 is never emitted");
            is_cu_qp_delta_coded = true;
        }
        TU tu{px, py, size, tree, ar, bl, {luma_mode, chroma_mode, chroma_mode}};
        if (yc) {
            if (c.decision(CTX_TS_FLAG)) fail("transform_skip_flag is never set");
            decode_residual(0, px, py, l2);
        }
        if (tree != DUAL_TREE_CHROMA) reconstruct(tu, 0, yc);  // luma first: CCLM reads it
        if (cb) {
            if (c.decision(CTX_TS_FLAG + 1)) fail("transform_skip_flag is never set");
            decode_residual(1, px / 2, py / 2, l2 - 1);
        }
        if (cr) {
            if (c.decision(CTX_TS_FLAG + 1)) fail("transform_skip_flag is never set");
            decode_residual(2, px / 2, py / 2, l2 - 1);
        }
        if (tree != DUAL_TREE_LUMA) {
            reconstruct(tu, 1, cb);
            reconstruct(tu, 2, cr);
        }
        if (tree != DUAL_TREE_CHROMA && mts_zero_out && !mts_dc_only)
            if (c.decision(CTX_MTS)) fail("mts_idx != 0 is never emitted");
    }

    struct Nd { int x, y, w; bool ar, bl; };
    Nd child(const Nd &p, int i) const {  // availability flags of a QT child (ctu.rs:2083-2188)
        Nd q;
        q.w = p.w / 2; q.x = p.x + (i % 2) * q.w; q.y = p.y + (i / 2) * q.w;
        if (q.x + q.w >= P.W) q.ar = false;
        else if (i == 0) q.ar = 0 < q.y;
        else if (i == 1) q.ar = p.ar;
        else q.ar = i == 2;
        if (q.y + q.w >= P.H) q.bl = false;
        else if (i == 1 || i == 3) q.bl = false;
        else if (i == 0) q.bl = 0 < q.x;
        else q.bl = p.bl;
        return q;
    }
    void mark_size(const Nd &n, int size) {
        for (int y = 0; y < n.w; y += 4)
            for (int x = 0; x < n.w; x += 4) size_map[(size_t)((n.y + y) / 4) * (P.W / 4) + (n.x + x) / 4] = (uint8_t)size;
    }
    void decode_tree(const Nd &n, CtuRecord &rec, int ctu_x, int ctu_y, int bit) {
        bool cl = n.x - 1 >= 0 && size_at(n.x - 1, n.y) < n.w;
        bool ca = n.y - 1 >= 0 && size_at(n.x, n.y - 1) < n.w;
        bool split = c.decision(CTX_SPLIT_CU + (int)cl + (int)ca) != 0;
        if (!split) {
            mark_size(n, n.w);
            decode_cu(n.x, n.y, n.w, SINGLE_TREE, n.ar, n.bl, rec, ctu_x, ctu_y);
            return;
        }
        rec.split_mask |= 1u << bit;
        if (n.w == 8) {
            mark_size(n, 4);
            for (int i = 0; i < 4; i++) {
                Nd q = child(n, i);
                decode_cu(q.x, q.y, 4, DUAL_TREE_LUMA, q.ar, q.bl, rec, ctu_x, ctu_y);
            }
            bool ar = (n.x + n.w >= P.W) ? false : n.ar, bl = (n.y + n.w >= P.H) ? false : n.bl;
            decode_cu(n.x, n.y, 8, DUAL_TREE_CHROMA, ar, bl, rec, ctu_x, ctu_y);
            return;
        }
        // a 32 or 16 node that splits: its area is not final yet; neighbours to the right/below see the children as they decode.
        // Mark the whole node with the children's size first so that not-yet-decoded siblings read a defined value that is
        // never consulted (contexts only look left/above, i.e. at blocks decoded earlier).
        for (int i = 0; i < 4; i++) {
            Nd q = child(n, i);
            int cbit = n.w == 32 ? 1 + i : 5 + 4 * (bit - 1) + i;
            decode_tree(q, rec, ctu_x, ctu_y, cbit);
        }
    }

    void run(int qp, int W, int H, const uint8_t *data, size_t len) {
        Tuning t;
        K.init(qp, t);
        std::vector<uint8_t> zy((size_t)W * H, 0), zc((size_t)W * H / 4, 0);
        P.init(W, H, zy.data(), zc.data(), zc.data());
        size_map.assign((size_t)(W / 4) * (H / 4), 0);
        c.init(qp, data, len);
        for (int cy = 0; cy < H; cy += 32)
            for (int cx = 0; cx < W; cx += 32) {
                CtuRecord &rec = P.records[(size_t)(cy / 32) * (W / 32) + cx / 32];
                memset(&rec, 0, sizeof(rec));
                is_cu_qp_delta_coded = false;
                Nd root{cx, cy, 32, (cx + 32 >= W) ? false : (0 < cy && cx + 32 < W), false};
                decode_tree(root, rec, cx, cy, 0);
            }
        if (!c.terminate()) fail("end_of_slice_one_bit missing");
    }
};

}  // namespace
}  // namespace wo

extern "C" {
// Decodes one picture's slice_data().  Returns 0 on success, -1 on a malformed stream (message on stderr).
// records: (H/32)*(W/32) CtuRecord entries with cost = 0; coef_*: int16 planes; rec_*: reconstruction.
int wo_decode_picture(int qp, int W, int H, const uint8_t *data, size_t len, uint8_t *rec_y, uint8_t *rec_cb, uint8_t *rec_cr, int16_t *coef_y,
                      int16_t *coef_cb, int16_t *coef_cr, void *records, size_t *bits_consumed) {
    wo::Decoder d;
    try {
        d.run(qp, W, H, data, len);
    } catch (const std::exception &e) {
        fprintf(stderr, "wo_decode_picture: %s\n", e.what());
        return -1;
    }
    uint8_t *rec[3] = {rec_y, rec_cb, rec_cr};
    int16_t *coef[3] = {coef_y, coef_cb, coef_cr};
    for (int c = 0; c < 3; c++) {
        if (rec[c]) memcpy(rec[c], d.P.rec[c].d.data(), d.P.rec[c].d.size());
        if (coef[c]) memcpy(coef[c], d.P.coef[c].data(), d.P.coef[c].size() * sizeof(int16_t));
    }
    if (records) memcpy(records, d.P.records.data(), d.P.records.size() * sizeof(wo::CtuRecord));
    if (bits_consumed) *bits_consumed = d.c.br.pos;
    return 0;
}
}
