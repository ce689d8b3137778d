// slice_coder.cu — phase 2 of the hot path: the decided trees are turned into the CABAC-coded slice_data() of each picture on
// the device (reference src/ctu_encoder.rs:227-2269 syntax order, src/bool_coder.rs:136-296 arithmetic coder).
//
// Two kernels, because every context index of the emitted subset depends only on the decisions (levels, modes, tree), never
// on the arithmetic coder's state:
//   wrenc_b200_syntax_kernel  one thread per CTU: walks the CTU's coding tree in syntax order and writes the bin string as
//                             16-bit entries (context index | bin value | bypass flag) into the CTU's slot of the bin arena;
//   wrenc_b200_cabac_kernel   one thread per picture: initialises the 253 contexts from the slice QP, runs the range coder
//                             over the CTUs' bin strings in raster order, terminates, byte-aligns.
// I-slice subset actually emitted: split_cu_flag; intra_luma_mpm_flag / not_planar / mpm_idx / mpm_remainder;
// cclm_mode_flag / cclm_mode_idx / intra_chroma_pred_mode(=4); tu_cb/cr/y_coded_flag; cu_qp_delta_abs(=0) once per CTU;
// transform_skip_flag(=0); residual_coding with dependent quantisation; mts_idx(=0); end_of_slice_one_bit.
#include <cuda_runtime.h>
#include <stdint.h>

#define WB_CABAC_TABLE static __constant__ const
#include "cabac_tables.h"
#include "search_kernel_api.h"

namespace wb {

// 4x4 up-right diagonal scan (ctu.rs:53-77): x | y << 2
__constant__ uint8_t c_diag4[16] = {0 | 0 << 2, 0 | 1 << 2, 1 | 0 << 2, 0 | 2 << 2, 1 | 1 << 2, 2 | 0 << 2, 0 | 3 << 2, 1 | 2 << 2,
                                    2 | 1 << 2, 3 | 0 << 2, 1 | 3 << 2, 2 | 2 << 2, 3 | 1 << 2, 2 | 3 << 2, 3 | 2 << 2, 3 | 3 << 2};
__constant__ uint8_t c_rice[32] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3};

struct SbOrder {  // sub-block diagonal orders for 1x1, 2x2, 4x4, 8x8 sub-block grids: x | y << 4
    uint8_t o[1 + 4 + 16 + 64];
};
__device__ __forceinline__ int sb_off(int l2) { return l2 == 2 ? 0 : (l2 == 3 ? 1 : (l2 == 4 ? 5 : 21)); }

__device__ void build_sb_order(SbOrder &T) {
    for (int l2 = 2; l2 <= 5; l2++) {
        const int nsb = 1 << (l2 - 2);
        int i = 0, x = 0, y = 0;
        while (i < nsb * nsb) {
            while (y >= 0) {
                if (x < nsb && y < nsb) T.o[sb_off(l2) + i++] = (uint8_t)(x | (y << 4));
                y--; x++;
            }
            y = x; x = 0;
        }
    }
}

struct Sink {
    uint16_t *p;
    int n, cap;  // entries beyond cap are counted, not stored (staging pass)
    __device__ __forceinline__ void put(unsigned e) {
        if (n < cap) p[n] = (uint16_t)e;
        n++;
    }
    __device__ __forceinline__ void ctx(int c, int b) { put((unsigned)c | ((unsigned)(b & 1) << 9)); }
    __device__ __forceinline__ void byp(int b) { put(((unsigned)(b & 1) << 9) | (1u << 10)); }
    __device__ __forceinline__ void byp_bits(unsigned v, int nb) {
        for (int i = nb - 1; i >= 0; i--) byp((v >> i) & 1);
    }
};

struct PicView {
    int W, H, Wc, Hc;
    const int16_t *lev[3];
    const CtuRecord *rec;
    const uint8_t *mode_map;
};

__device__ __forceinline__ int tr_state(int s, int par) { return 2 * ((par & 1) ^ (s & 1)) + (s >> 1); }  // encoder_context.rs:339

// size of the luma CU that covers the 4x4 block (bx, by) of a CTU with the given split mask
__device__ __forceinline__ int leaf_size(unsigned mask, int bx, int by) {
    if (!(mask & 1)) return 32;
    const int a = ((by >> 2) << 1) | (bx >> 2);
    if (!((mask >> (1 + a)) & 1)) return 16;
    const int b = (((by >> 1) & 1) << 1) | ((bx >> 1) & 1);
    if (!((mask >> (5 + 4 * a + b)) & 1)) return 8;
    return 4;
}
__device__ __forceinline__ int cu_size_at(const PicView &P, int px, int py) {
    const CtuRecord &r = P.rec[(py >> 5) * P.Wc + (px >> 5)];
    return leaf_size(r.split_mask, (px & 31) >> 2, (py & 31) >> 2);
}
__device__ __forceinline__ int luma_mode_at(const PicView &P, int px, int py) { return P.mode_map[(size_t)(py >> 2) * (P.W >> 2) + (px >> 2)]; }

struct TuState {
    bool mts_dc_only, mts_zero_out, qp_delta_coded;
};

// residual_coding() of one transform block (ctu_encoder.rs:1786-2269), regular (non transform-skip) path with dep-quant
__device__ void code_residual(Sink &S, const SbOrder &SO, const int16_t *q, int stride, int c_idx, int l2, TuState &ts, uint16_t *pass1, uint16_t *absl) {
    const int n = 1 << l2, nn = n * n, nsbw = n >> 2;
    const uint8_t *sbo = SO.o + sb_off(l2);
    auto pos_of = [&](int k, int &x, int &y) {
        const int sb = k >> 4, p = k & 15;
        x = ((sbo[sb] & 15) << 2) + (c_diag4[p] & 3);
        y = ((sbo[sb] >> 4) << 2) + (c_diag4[p] >> 2);
    };
    // last significant coefficient in coding order (ctu.rs:867-899)
    int last_k = 0, lx = 0, ly = 0;
    for (int k = nn - 1; k >= 0; k--) {
        int x, y;
        pos_of(k, x, y);
        if (q[y * stride + x] != 0 || k == 0) { last_k = k; lx = x; ly = y; break; }
    }
    // last_sig_coeff_{x,y}_{prefix,suffix} (ctu_encoder.rs:1818-1849, bool_coder.rs:2053-2083)
    int pre[2], suf[2], sbits[2];
    for (int d = 0; d < 2; d++) {
        const int v = d ? ly : lx;
        if (v <= 3) { pre[d] = v; suf[d] = 0; sbits[d] = 0; }
        else {
            int b = 1, p, s;
            for (;;) { p = v >> b; s = v - (p << b); if (p < 4) break; b++; }
            pre[d] = ((b + 1) << 1) + (p & 1); suf[d] = s; sbits[d] = (pre[d] >> 1) - 1;
        }
    }
    {
        const int cmax = (l2 << 1) - 1;
        int off, shift;
        if (c_idx == 0) { off = l2 == 2 ? 0 : (l2 == 3 ? 3 : (l2 == 4 ? 6 : 10)); shift = (l2 + 1) >> 2; }
        else { off = 20; shift = min(2, max(0, n >> 3)); }
        for (int d = 0; d < 2; d++) {
            const int base = d ? CTX_LAST_Y : CTX_LAST_X;
            for (int i = 0; i < pre[d]; i++) S.ctx(base + (i >> shift) + off, 1);
            if (pre[d] < cmax) S.ctx(base + (pre[d] >> shift) + off, 0);
        }
        for (int d = 0; d < 2; d++)
            if (pre[d] > 3) S.byp_bits((unsigned)suf[d], sbits[d]);
    }
    int rem = (nn * 7) >> 2;
    const int last_sb = last_k >> 4, last_pos = last_k & 15;
    if ((last_sb > 0 || last_pos > 0) && c_idx == 0) ts.mts_dc_only = false;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) { pass1[y * n + x] = 0; absl[y * n + x] = 0; }
    auto sb_nonzero = [&](int xs, int ys) {
        for (int yy = 0; yy < 4; yy++)
            for (int xx = 0; xx < 4; xx++)
                if (q[((ys << 2) + yy) * stride + (xs << 2) + xx] != 0) return true;
        return false;
    };
    auto loc_sums = [&](const uint16_t *a, int x, int y, int &num) {
        int sum = 0; num = 0;
        if (x < n - 1) {
            int v = a[y * n + x + 1]; sum += v; num += v > 0;
            if (x < n - 2) { v = a[y * n + x + 2]; sum += v; num += v > 0; }
            if (y < n - 1) { v = a[(y + 1) * n + x + 1]; sum += v; num += v > 0; }
        }
        if (y < n - 1) {
            int v = a[(y + 1) * n + x]; sum += v; num += v > 0;
            if (y < n - 2) { v = a[(y + 2) * n + x]; sum += v; num += v > 0; }
        }
        return sum;
    };
    auto code_rem = [&](int value, int rice) {  // abs_remainder / dec_abs_level binarisation (bool_coder.rs:1384-1465), all bypass
        const int cmax = 6 << rice;
        const int pv = min(cmax, value);
        const int pre_ = pv >> rice;
        if (pre_ < 6) {
            for (int i = 0; i < pre_; i++) S.byp(1);
            S.byp(0);
            if (rice > 0) S.byp_bits((unsigned)(pv - (pre_ << rice)), rice);
        } else {
            for (int i = 0; i < 6; i++) S.byp(1);
            // limited k-th order exp-Golomb escape, k = rice + 1, maxPreExtLen 11, truncSuffixLen 15 (bool_coder.rs:1278-1303)
            int sym = value - cmax;
            const int k = rice + 1;
            const int cv = sym >> k;
            int pel = 0;
            while (pel < 11 && cv > (2 << pel) - 2) { pel++; S.byp(1); }
            int esc;
            if (pel == 11) esc = 15;
            else { S.byp(0); esc = pel + k; }
            sym -= ((1 << pel) - 1) << k;
            S.byp_bits((unsigned)sym, esc);
        }
    };
    int qstate = 0;
    for (int i = last_sb; i >= 0; i--) {
        const int xs = sbo[i] & 15, ys = sbo[i] >> 4;
        int a[16];
        {
            int st = qstate;
            for (int p = 15; p >= 0; p--) {
                const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
                const int v = q[y * stride + x];
                a[p] = ((v < 0 ? -v : v) + (st > 1)) >> 1;
                st = tr_state(st, a[p]);
            }
        }
        const bool sbcoded = sb_nonzero(xs, ys) || i == 0;
        bool infer = false;
        if (i < last_sb && i > 0) {
            int csbf = 0;
            if (xs < nsbw - 1) csbf += sb_nonzero(xs + 1, ys);
            if (ys < nsbw - 1) csbf += sb_nonzero(xs, ys + 1);
            S.ctx(CTX_SB_CODED + (c_idx == 0 ? min(csbf, 1) : 2 + min(csbf, 1)), sbcoded);
            infer = true;
        }
        if (sbcoded && (xs > 3 || ys > 3) && c_idx == 0) ts.mts_zero_out = false;
        const int fp0 = i == last_sb ? last_pos : 15;
        int fp1 = fp0;
        for (int p = fp0; p >= 0; p--) {
            if (rem < 4) break;
            const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
            const bool is_last = x == lx && y == ly;
            const bool sig = q[y * stride + x] != 0 || is_last || (p == 0 && infer && sbcoded);
            if (sbcoded && (p > 0 || !infer) && !is_last) {
                int num;
                const int sum = loc_sums(pass1, x, y, num);
                const int d = x + y;
                int ci;
                if (c_idx == 0) ci = 12 * max(0, qstate - 1) + min(3, (sum + 1) >> 1) + (d < 2 ? 8 : (d < 5 ? 4 : 0));
                else ci = 36 + 8 * max(0, qstate - 1) + min(3, (sum + 1) >> 1) + (d < 2 ? 4 : 0);
                S.ctx(CTX_SIG + ci, sig);
                rem--;
                if (sig) infer = false;
            }
            const int al = a[p];
            const bool gt1 = al > 1, gt3 = al > 3, par = al > 1 && (al & 1);
            if (sig) {
                int num;
                const int sum = loc_sums(pass1, x, y, num);
                const int d = x + y;
                int ci;
                if (is_last) ci = c_idx == 0 ? 0 : 21;
                else if (c_idx == 0) ci = 1 + min(4, sum - num) + (d == 0 ? 15 : (d < 3 ? 10 : (d < 10 ? 5 : 0)));
                else ci = 22 + min(4, sum - num) + (d == 0 ? 5 : 0);
                S.ctx(CTX_GTX + ci, gt1);
                rem--;
                if (gt1) {
                    S.ctx(CTX_PAR + ci, par);
                    rem--;
                    S.ctx(CTX_GTX + ci + 32, gt3);
                    rem--;
                }
            }
            const int p1 = (int)sig + (int)par + (int)gt1 + 2 * (int)gt3;
            pass1[y * n + x] = (uint16_t)p1;
            qstate = tr_state(qstate, p1);
            fp1 = p - 1;
        }
        for (int p = fp0; p > fp1; p--) {  // abs_remainder of the positions coded in pass 1
            const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
            if (a[p] > 3) {
                int num;
                const int sum = loc_sums(absl, x, y, num);
                const int rice = c_rice[min(31, max(0, sum - 20))];
                code_rem((a[p] - pass1[y * n + x]) >> 1, rice);
            }
            absl[y * n + x] = (uint16_t)a[p];
        }
        for (int p = fp1; p >= 0; p--) {  // dec_abs_level of the rest
            const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
            absl[y * n + x] = (uint16_t)a[p];
            if (sbcoded) {
                int num;
                const int sum = loc_sums(absl, x, y, num);
                const int rice = c_rice[min(31, max(0, sum))];
                const int zero_pos = (qstate < 2 ? 1 : 2) << rice;
                const int v = a[p];
                const int dec = v == 0 ? zero_pos : (zero_pos >= v ? v - 1 : v);
                code_rem(dec, rice);
            }
            qstate = tr_state(qstate, a[p]);
        }
        for (int p = 15; p >= 0; p--) {  // coeff_sign_flag (sign data hiding off)
            if (a[p] > 0) {
                const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
                S.byp(q[y * stride + x] < 0);
            }
        }
    }
}

__device__ bool block_nonzero(const int16_t *q, int stride, int n) {
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            if (q[y * stride + x] != 0) return true;
    return false;
}

// MPM list of the syntax pass (ctu.rs:1498-1635): true neighbours from the final mode map
__device__ void code_luma_mode(Sink &S, const PicView &P, int px, int py, int size, int mode) {
    if (mode == 0) { S.ctx(CTX_MPM_FLAG, 1); S.ctx(CTX_NOT_PLANAR + 1, 0); return; }
    const int left = px > 0 ? luma_mode_at(P, px - 1, py + size - 1) : 0;
    const int above = (py > 0 && (py & 31) != 0) ? luma_mode_at(P, px + size - 1, py - 1) : 0;
    int cand[5];
    if (left == above && left > 1) {
        const int m = left;
        cand[0] = m; cand[1] = 2 + (m + 61) % 64; cand[2] = 2 + (m - 1) % 64; cand[3] = 2 + (m + 60) % 64; cand[4] = 2 + m % 64;
    } else if (left != above && (left > 1 || above > 1)) {
        const int mn = min(left, above), mx = max(left, above);
        if (mn > 1) {
            const int d = mx - mn;
            cand[0] = left; cand[1] = above;
            if (d == 1) { cand[2] = 2 + (mn + 61) % 64; cand[3] = 2 + (mx - 1) % 64; cand[4] = 2 + (mn + 60) % 64; }
            else if (d >= 62) { cand[2] = 2 + (mn - 1) % 64; cand[3] = 2 + (mx + 61) % 64; cand[4] = 2 + mn % 64; }
            else if (d == 2) { cand[2] = 2 + (mn - 1) % 64; cand[3] = 2 + (mn + 61) % 64; cand[4] = 2 + (mx - 1) % 64; }
            else { cand[2] = 2 + (mn + 61) % 64; cand[3] = 2 + (mn - 1) % 64; cand[4] = 2 + (mx + 61) % 64; }
        } else {
            cand[0] = mx; cand[1] = 2 + (mx + 61) % 64; cand[2] = 2 + (mx - 1) % 64; cand[3] = 2 + (mx + 60) % 64; cand[4] = 2 + mx % 64;
        }
    } else {
        cand[0] = 1; cand[1] = 50; cand[2] = 18; cand[3] = 46; cand[4] = 54;
    }
    int idx = -1;
    for (int i = 4; i >= 0; i--)
        if (cand[i] == mode) idx = i;
    if (idx >= 0) {
        S.ctx(CTX_MPM_FLAG, 1);
        S.ctx(CTX_NOT_PLANAR + 1, 1);
        for (int i = 0; i < idx; i++) S.byp(1);  // TR cMax 4, all bypass
        if (idx < 4) S.byp(0);
    } else {
        S.ctx(CTX_MPM_FLAG, 0);
        int below = 0;
        for (int i = 0; i < 5; i++) below += cand[i] < mode;
        const int remv = mode - 1 - below;
        // truncated binary, cMax 60: n = 61, k = 5, u = 3 (bool_coder.rs:1246-1255)
        if (remv < 3) S.byp_bits((unsigned)remv, 5);
        else S.byp_bits((unsigned)(remv + 3), 6);
    }
}

// coding_unit() + transform_unit() of one CU (ctu_encoder.rs:440-1321, 1414-1784); (x, y) CTU-relative luma position
__device__ void code_cu(Sink &S, const SbOrder &SO, const PicView &P, const CtuRecord &rec, int ctu_x, int ctu_y, int x, int y, int size, int tree, TuState &ts,
                        uint16_t *pass1, uint16_t *absl) {
    const int px = ctu_x + x, py = ctu_y + y;
    if (tree != DUAL_TREE_CHROMA) code_luma_mode(S, P, px, py, size, rec.luma_mode[(y >> 2) * 8 + (x >> 2)]);
    if (tree != DUAL_TREE_LUMA) {
        const int cm = rec.chroma_mode[(y >> 3) * 4 + (x >> 3)];
        const bool cclm = cm >= MODE_LT_CCLM;
        S.ctx(CTX_CCLM_FLAG, cclm);
        if (cclm) {
            const int idx = cm - MODE_LT_CCLM;  // TR cMax 2: bin 0 context coded, bin 1 bypass
            S.ctx(CTX_CCLM_IDX, idx > 0);
            if (idx > 0) S.byp(idx > 1);
        } else {
            S.ctx(CTX_CHROMA_PRED, 0);  // intra_chroma_pred_mode == 4 (DM)
        }
    }
    ts.mts_dc_only = true;
    ts.mts_zero_out = true;
    const int cw = P.W >> 1;
    const int16_t *qy = P.lev[0] + (size_t)py * P.W + px;
    const int16_t *qcb = P.lev[1] + (size_t)(py >> 1) * cw + (px >> 1), *qcr = P.lev[2] + (size_t)(py >> 1) * cw + (px >> 1);
    const int l2 = 31 - __clz(size);
    const bool ycbf = tree != DUAL_TREE_CHROMA && block_nonzero(qy, P.W, size);
    const bool cbcbf = tree != DUAL_TREE_LUMA && block_nonzero(qcb, cw, size >> 1);
    const bool crcbf = tree != DUAL_TREE_LUMA && block_nonzero(qcr, cw, size >> 1);
    if (tree != DUAL_TREE_LUMA) {
        S.ctx(CTX_TU_CB, cbcbf);
        S.ctx(CTX_TU_CR + (cbcbf ? 1 : 0), crcbf);
    }
    if (tree != DUAL_TREE_CHROMA) S.ctx(CTX_TU_Y, ycbf);
    if ((ycbf || cbcbf || crcbf) && tree != DUAL_TREE_CHROMA && !ts.qp_delta_coded) {
        S.ctx(CTX_QP_DELTA_ABS, 0);  // cu_qp_delta_abs == 0
        ts.qp_delta_coded = true;
    }
    if (ycbf) {
        S.ctx(CTX_TS_FLAG, 0);
        code_residual(S, SO, qy, P.W, 0, l2, ts, pass1, absl);
    }
    if (cbcbf) {
        S.ctx(CTX_TS_FLAG + 1, 0);
        code_residual(S, SO, qcb, cw, 1, l2 - 1, ts, pass1, absl);
    }
    if (crcbf) {
        S.ctx(CTX_TS_FLAG + 1, 0);
        code_residual(S, SO, qcr, cw, 2, l2 - 1, ts, pass1, absl);
    }
    if (tree != DUAL_TREE_CHROMA && ts.mts_zero_out && !ts.mts_dc_only) S.ctx(CTX_MTS, 0);  // mts_idx == 0
}

// coding_tree() (ctu_encoder.rs:227-438); QT only, local dual tree at 8x8 -> 4x4.  Written as nested loops over the three
// quad-tree levels (no device recursion: the stack frame stays statically sized).
__device__ __forceinline__ bool code_split_flag(Sink &S, const PicView &P, const CtuRecord &rec, int px, int py, int size, int bit) {
    // allow_split_qt holds for 32, 16, 8 (encoder_context.rs:958-971); ctxInc bool_coder.rs:2659-2744
    const bool split = (rec.split_mask >> bit) & 1;
    const bool cl = px > 0 && cu_size_at(P, px - 1, py) < size;
    const bool ca = py > 0 && cu_size_at(P, px, py - 1) < size;
    S.ctx(CTX_SPLIT_CU + (int)cl + (int)ca, split);
    return split;
}

__device__ void code_ctu(Sink &S, const SbOrder &SO, const PicView &P, const CtuRecord &rec, int ctu_x, int ctu_y, TuState &ts, uint16_t *pass1, uint16_t *absl) {
    if (!code_split_flag(S, P, rec, ctu_x, ctu_y, 32, 0)) {
        code_cu(S, SO, P, rec, ctu_x, ctu_y, 0, 0, 32, SINGLE_TREE, ts, pass1, absl);
        return;
    }
    for (int a = 0; a < 4; a++) {
        const int x16 = (a & 1) * 16, y16 = (a >> 1) * 16;
        if (!code_split_flag(S, P, rec, ctu_x + x16, ctu_y + y16, 16, 1 + a)) {
            code_cu(S, SO, P, rec, ctu_x, ctu_y, x16, y16, 16, SINGLE_TREE, ts, pass1, absl);
            continue;
        }
        for (int b = 0; b < 4; b++) {
            const int x8 = x16 + (b & 1) * 8, y8 = y16 + (b >> 1) * 8;
            if (!code_split_flag(S, P, rec, ctu_x + x8, ctu_y + y8, 8, 5 + 4 * a + b)) {
                code_cu(S, SO, P, rec, ctu_x, ctu_y, x8, y8, 8, SINGLE_TREE, ts, pass1, absl);
                continue;
            }
            for (int i = 0; i < 4; i++) code_cu(S, SO, P, rec, ctu_x, ctu_y, x8 + (i & 1) * 4, y8 + (i >> 1) * 4, 4, DUAL_TREE_LUMA, ts, pass1, absl);
            code_cu(S, SO, P, rec, ctu_x, ctu_y, x8, y8, 8, DUAL_TREE_CHROMA, ts, pass1, absl);
        }
    }
}

extern "C" __global__ void __launch_bounds__(64) wrenc_b200_syntax_kernel(SyntaxParams Q) {
    __shared__ SbOrder SO;
    if (threadIdx.x == 0) build_sb_order(SO);
    __syncthreads();
    const int nctu = Q.Wc * Q.Hc;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)Q.n_pics * nctu) return;
    const int pic = (int)(gid / nctu), ctu = (int)(gid - (long long)pic * nctu);
    PicView P;
    P.W = Q.W; P.H = Q.H; P.Wc = Q.Wc; P.Hc = Q.Hc;
    const size_t ps = (size_t)Q.W * Q.H * 3 / 2;
    P.lev[0] = Q.lev + (size_t)pic * ps;
    P.lev[1] = P.lev[0] + (size_t)Q.W * Q.H;
    P.lev[2] = P.lev[1] + (size_t)(Q.W >> 1) * (Q.H >> 1);
    P.rec = Q.records + (size_t)pic * nctu;
    P.mode_map = Q.mode_map + (size_t)pic * (Q.W >> 2) * (Q.H >> 2);
    // First pass (Q.bins == nullptr): count the CTU's bins and keep the first stage_cap of them in the CTU's staging slot.
    // Second pass: only the CTUs whose string did not fit are walked again and written at their scanned arena offset (the
    // others are copied there by wrenc_b200_bin_compact_kernel).
    Sink S;
    S.n = 0;
    if (!Q.bins) {
        S.p = Q.stage + (size_t)gid * Q.stage_cap;
        S.cap = Q.stage_cap;
    } else {
        if (Q.bin_count[gid] <= Q.stage_cap) return;
        if (Q.bin_offset[gid] + (unsigned long long)Q.bin_count[gid] > Q.bins_cap) return;  // arena too small: the coder kernel reports it
        S.p = Q.bins + Q.bin_offset[gid];
        S.cap = 0x7fffffff;
    }
    uint16_t pass1[1024], absl[1024];
    TuState ts;
    ts.qp_delta_coded = false;  // quantisation group = CTU (cu_qp_delta_subdiv 0, ctu_encoder.rs:305-310)
    ts.mts_dc_only = true;
    ts.mts_zero_out = true;
    const int cx = (ctu % Q.Wc) * 32, cy = (ctu / Q.Wc) * 32;
    code_ctu(S, SO, P, P.rec[ctu], cx, cy, ts, pass1, absl);
    if (!Q.bins) Q.bin_count[gid] = S.n;
}

// staged bin strings -> their arena offsets, one warp per CTU (strings longer than the staging slot are written by the second
// syntax pass instead)
extern "C" __global__ void __launch_bounds__(256) wrenc_b200_bin_compact_kernel(SyntaxParams Q) {
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gid >= (long long)Q.n_pics * Q.Wc * Q.Hc) return;
    const int cnt = Q.bin_count[gid];
    if (cnt > Q.stage_cap) return;
    if (Q.bin_offset[gid] + (unsigned long long)cnt > Q.bins_cap) return;  // arena too small: the coder kernel reports it
    const uint16_t *src = Q.stage + (size_t)gid * Q.stage_cap;
    uint16_t *dst = Q.bins + Q.bin_offset[gid];
    for (int i = lane; i < cnt; i += 32) dst[i] = src[i];
}

// exclusive prefix sum of the per-CTU bin counts over all pictures (one block; the counts are a few hundred thousand ints)
extern "C" __global__ void __launch_bounds__(1024) wrenc_b200_bin_scan_kernel(const int *count, unsigned long long *offset, long long n, unsigned long long *total) {
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x;
    const long long chunk = (n + 1023) / 1024, b = t * chunk, e = min(n, b + chunk);
    unsigned long long s = 0;
    for (long long i = b; i < e; i++) s += (unsigned long long)count[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < 1024; i++) { unsigned long long v = part[i]; part[i] = acc; acc += v; }
        *total = acc;
    }
    __syncthreads();
    unsigned long long acc = part[t];
    for (long long i = b; i < e; i++) { offset[i] = acc; acc += (unsigned long long)count[i]; }
}

// ---------------------------------------------------------------------------------------------------------------
// arithmetic coder (bool_coder.rs:136-296, 1073-1111) — one thread per picture
// ---------------------------------------------------------------------------------------------------------------
struct BitOut {
    uint8_t *p;
    size_t n, cap;
    unsigned acc;
    int nb;
    __device__ __forceinline__ void bit(int b) {
        acc = (acc << 1) | (unsigned)(b & 1);
        if (++nb == 8) {
            if (p && n < cap) p[n] = (uint8_t)acc;
            n++;
            nb = 0;
            acc = 0;
        }
    }
};
struct Engine {
    unsigned low, range;
    int outstanding;
    bool first;
    BitOut out;
    __device__ __forceinline__ void put(int b) {  // flush_cabac_bin: the very first bit of the slice is dropped
        if (!first) out.bit(b);
        first = false;
        while (outstanding > 0) { out.bit(!b); outstanding--; }
    }
    __device__ __forceinline__ void renorm() {
        while (range < 256) {
            if (low < 256) put(0);
            else if (low >= 512) { low -= 512; put(1); }
            else { low -= 256; outstanding++; }
            range <<= 1;
            low <<= 1;
        }
    }
};

// One WARP per picture.  The coder itself is sequential, so every lane runs it redundantly on identical state (context
// states in shared memory, engine state in registers; lane 0 alone stores the output bytes); what the warp buys is the
// memory side: the bin strings are fetched 32 entries at a time with one coalesced load, double-buffered, and handed to the
// coder by shuffles, instead of one dependent 2-byte global load per bin.
extern "C" __global__ void __launch_bounds__(32) wrenc_b200_cabac_kernel(SyntaxParams Q) {
    const int pic = blockIdx.x, lane = threadIdx.x;
    if (pic >= Q.n_pics) return;
    const int nctu = Q.Wc * Q.Hc;
    __shared__ uint16_t p0[CTX_TOTAL], p1[CTX_TOTAL];
    __shared__ uint8_t sh0[CTX_TOTAL], sh1[CTX_TOTAL];
    for (int i = lane; i < CTX_TOTAL; i += 32) {  // init_ctx_table (bool_coder.rs:1073-1093)
        const int iv = kCabacInitValue[i];
        const int m = (iv >> 3) - 4, nn = (iv & 7) * 18 + 1;
        const int pre = min(127, max(1, ((m * (min(63, max(0, Q.qp)) - 16)) >> 1) + nn));
        p0[i] = (uint16_t)(pre << 3);
        p1[i] = (uint16_t)(pre << 7);
        const int si = kCabacShiftIdx[i];
        sh0[i] = (uint8_t)((si >> 2) + 2);
        sh1[i] = (uint8_t)((si & 3) + 3 + (si >> 2) + 2);
    }
    __syncwarp();
    Engine E;
    E.low = 0; E.range = 510; E.outstanding = 0; E.first = true;
    E.out.p = lane == 0 ? Q.out + (size_t)pic * Q.out_cap : nullptr;  // lanes other than 0 only count
    E.out.n = 0; E.out.cap = Q.out_cap; E.out.acc = 0; E.out.nb = 0;
    int overflow = 0;
    // the CTUs' bin strings lie back to back in the arena (exclusive scan of the counts), in raster order
    const size_t g0 = (size_t)pic * nctu;
    const unsigned long long beg = Q.bin_offset[g0];
    const unsigned long long end = Q.bin_offset[g0 + nctu - 1] + (unsigned long long)Q.bin_count[g0 + nctu - 1];
    if (end > Q.bins_cap) {  // this picture's strings did not fit the arena (sized from earlier batches): nothing was written for it
        if (lane == 0) Q.out_len[pic] = -2;
        return;
    }
    const uint16_t *b = Q.bins + beg;
    const long long total = (long long)(end - beg);
    unsigned nxt = lane < total ? b[lane] : 0u;
    for (long long base = 0; base < total; base += 32) {
        const unsigned cur = nxt;
        const long long pf = base + 32 + lane;
        nxt = pf < total ? b[pf] : 0u;  // prefetch the next 32 entries while this batch is coded
        const int cnt = (int)min(32ll, total - base);
        for (int i = 0; i < cnt; i++) {
            const unsigned e = __shfl_sync(0xffffffffu, cur, i);
            const int bin = (e >> 9) & 1;
            if (e & 1024u) {  // bypass (bool_coder.rs:202-216)
                E.low <<= 1;
                if (bin) E.low += E.range;
                if (E.low >= 1024) { E.put(1); E.low -= 1024; }
                else if (E.low < 512) E.put(0);
                else { E.low -= 512; E.outstanding++; }
            } else {  // context coded (bool_coder.rs:254-296)
                const int ci = e & 511;
                const unsigned q0 = p0[ci], q1 = p1[ci], s0 = sh0[ci], s1 = sh1[ci];
                const unsigned ps = q1 + 16u * q0;
                const unsigned mps = ps >> 14;
                const unsigned lps = ((((E.range >> 5) * ((mps ? 32767u - ps : ps) >> 9)) >> 1) + 4);
                if ((unsigned)bin == mps) E.range -= lps;
                else { E.low += E.range - lps; E.range = lps; }
                E.renorm();
                __syncwarp();  // every lane has read the old state before anyone stores the (identical) new one
                p0[ci] = (uint16_t)(q0 - (q0 >> s0) + ((1023 * bin) >> s0));
                p1[ci] = (uint16_t)(q1 - (q1 >> s1) + ((16383 * bin) >> s1));
                __syncwarp();
            }
        }
    }
    // end_of_slice_one_bit = 1 (bool_coder.rs:218-235), then byte alignment with zeros (slice_encoder.rs:419)
    E.range -= 2;
    E.low += E.range;
    E.range = 2;
    E.renorm();
    E.put((E.low >> 9) & 1);
    {
        const unsigned two = ((E.low >> 7) & 3) | 1;
        for (int k = 1; k >= 0; k--) {  // flush_cabac_trailing_bin
            const int bb = (two >> k) & 1;
            E.out.bit(bb);
            while (E.outstanding > 0) { E.out.bit(!bb); E.outstanding--; }
        }
    }
    while (E.out.nb != 0) E.out.bit(0);
    if (E.out.n > E.out.cap) overflow = 1;
    if (lane == 0) Q.out_len[pic] = overflow ? -1 : (int)E.out.n;
}

cudaError_t launch_syntax(const SyntaxParams &Q, cudaStream_t stream) {  // Q.bins == nullptr: counting pass
    const long long total = (long long)Q.n_pics * Q.Wc * Q.Hc;
    wrenc_b200_syntax_kernel<<<(unsigned)((total + 63) / 64), 64, 0, stream>>>(Q);
    return cudaGetLastError();
}
cudaError_t launch_bin_scan(const SyntaxParams &Q, unsigned long long *d_total, cudaStream_t stream) {
    const long long total = (long long)Q.n_pics * Q.Wc * Q.Hc;
    wrenc_b200_bin_scan_kernel<<<1, 1024, 0, stream>>>(Q.bin_count, Q.bin_offset, total, d_total);
    return cudaGetLastError();
}
cudaError_t launch_bin_compact(const SyntaxParams &Q, cudaStream_t stream) {
    const long long total = (long long)Q.n_pics * Q.Wc * Q.Hc;
    wrenc_b200_bin_compact_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, stream>>>(Q);
    return cudaGetLastError();
}
cudaError_t launch_cabac(const SyntaxParams &Q, cudaStream_t stream) {
    wrenc_b200_cabac_kernel<<<Q.n_pics, 32, 0, stream>>>(Q);
    return cudaGetLastError();
}

}  // namespace wb
