// syntax_walk.cuh — the syntax walk of the slice coder: coding_tree / coding_unit / transform_unit / residual_coding of one CTU
// (reference src/ctu_encoder.rs:227-2269) as a bin string of 16-bit entries (context index | bin << 9 | bypass << 10), written so
// that it compiles for the device (wrenc_b200_syntax_kernel, one thread per CTU) and for the host
// (tests/host/syntax_walk_host_test.cpp walks searched pictures with it and compares every entry with the bins the oracle's
// syntax writer hands to its arithmetic coder, without a GPU).
// I-slice subset actually emitted: split_cu_flag; intra_luma_mpm_flag / not_planar / mpm_idx / mpm_remainder;
// cclm_mode_flag / cclm_mode_idx / intra_chroma_pred_mode(=4); tu_cb/cr/y_coded_flag; cu_qp_delta_abs(=0) once per CTU;
// transform_skip_flag(=0); residual_coding with dependent quantisation; mts_idx(=0).  (end_of_slice_one_bit is the coder's.)
#pragma once
#include <stdint.h>

#ifndef WB_NZMAP
#define WB_NZMAP 1  // 1: the walk reads a per-CTU map of non-zero 4x4 level blocks instead of scanning the level planes
#endif
#ifndef WB_SYN_OUTLINE
#define WB_SYN_OUTLINE 1  // 1: coding_unit / residual_coding / luma mode are real functions (the walk is 28 k instructions when everything is inlined)
#endif
#ifndef WB_SYN_ROLL
#define WB_SYN_ROLL 1  // 1: the fixed-trip loops of residual_coding stay rolled
#endif
#if defined(__CUDACC__)
#define SW_CONST __constant__
#define SW_FN __device__
#define SW_FN_INLINE __device__ __forceinline__
#if WB_SYN_OUTLINE
#define SW_FN_NOINLINE __device__ __noinline__
#else
#define SW_FN_NOINLINE __device__
#endif
#if WB_SYN_ROLL
#define WB_SYN_UNROLL1 _Pragma("unroll 1")
#else
#define WB_SYN_UNROLL1
#endif
#ifndef WB_CABAC_TABLE
#define WB_CABAC_TABLE static __constant__ const
#endif
#else  // host build (tests): plain functions, tables in ordinary memory
#include <algorithm>
#define SW_CONST static const
#define SW_FN inline
#define SW_FN_INLINE inline
#define SW_FN_NOINLINE inline
#define WB_SYN_UNROLL1
#ifndef WB_CABAC_TABLE
#define WB_CABAC_TABLE static const
#endif
#endif
#include "cabac_tables.h"
#include "search_kernel_api.h"

namespace wb {

#if !defined(__CUDACC__)
using std::max;
using std::min;
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
#endif

// 4x4 up-right diagonal scan (ctu.rs:53-77): x | y << 2
SW_CONST uint8_t c_diag4[16] = {0 | 0 << 2, 0 | 1 << 2, 1 | 0 << 2, 0 | 2 << 2, 1 | 1 << 2, 2 | 0 << 2, 0 | 3 << 2, 1 | 2 << 2,
                                    2 | 1 << 2, 3 | 0 << 2, 1 | 3 << 2, 2 | 2 << 2, 3 | 1 << 2, 2 | 3 << 2, 3 | 2 << 2, 3 | 3 << 2};
SW_CONST uint8_t c_rice[32] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3};

struct SbOrder {  // sub-block diagonal orders for 1x1, 2x2, 4x4, 8x8 sub-block grids: x | y << 4
    uint8_t o[1 + 4 + 16 + 64];
};
SW_FN_INLINE int sb_off(int l2) { return l2 == 2 ? 0 : (l2 == 3 ? 1 : (l2 == 4 ? 5 : 21)); }

SW_FN void build_sb_order(SbOrder &T) {
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
    SW_FN_INLINE void put(unsigned e) {
        if (n < cap) p[n] = (uint16_t)e;
        n++;
    }
    SW_FN_INLINE void ctx(int c, int b) { put((unsigned)c | ((unsigned)(b & 1) << 9)); }
    SW_FN_INLINE void byp(int b) { put(((unsigned)(b & 1) << 9) | (1u << 10)); }
    SW_FN_INLINE void byp_bits(unsigned v, int nb) {
        WB_SYN_UNROLL1
        for (int i = nb - 1; i >= 0; i--) byp((v >> i) & 1);
    }
};

struct PicView {
    int W, H, Wc, Hc;
    const int16_t *lev[3];
    const CtuRecord *rec;
    const uint8_t *mode_map;
};

SW_FN_INLINE int tr_state(int s, int par) { return 2 * ((par & 1) ^ (s & 1)) + (s >> 1); }  // encoder_context.rs:339

// size of the luma CU that covers the 4x4 block (bx, by) of a CTU with the given split mask
SW_FN_INLINE int leaf_size(unsigned mask, int bx, int by) {
    if (!(mask & 1)) return 32;
    const int a = ((by >> 2) << 1) | (bx >> 2);
    if (!((mask >> (1 + a)) & 1)) return 16;
    const int b = (((by >> 1) & 1) << 1) | ((bx >> 1) & 1);
    if (!((mask >> (5 + 4 * a + b)) & 1)) return 8;
    return 4;
}
SW_FN_INLINE int cu_size_at(const PicView &P, int px, int py) {
    const CtuRecord &r = P.rec[(py >> 5) * P.Wc + (px >> 5)];
    return leaf_size(r.split_mask, (px & 31) >> 2, (py & 31) >> 2);
}
SW_FN_INLINE int luma_mode_at(const PicView &P, int px, int py) { return P.mode_map[(size_t)(py >> 2) * (P.W >> 2) + (px >> 2)]; }

struct TuState {
    bool mts_dc_only, mts_zero_out, qp_delta_coded;
};

// residual_coding() of one transform block (ctu_encoder.rs:1786-2269), regular (non transform-skip) path with dep-quant
// (nzm, rs, b0): the CTU's map of non-zero 4x4 blocks of this component, its row stride in bits and the bit of the TB's first block
SW_FN_NOINLINE void code_residual(Sink &S, const SbOrder &SO, const int16_t *q, int stride, int c_idx, int l2, TuState &ts, uint8_t *pass1, uint8_t *absl,
                              unsigned long long nzm, int rs, int b0) {
    const int n = 1 << l2, nn = n * n, nsbw = n >> 2;
    const uint8_t *sbo = SO.o + sb_off(l2);
    auto pos_of = [&](int k, int &x, int &y) {
        const int sb = k >> 4, p = k & 15;
        x = ((sbo[sb] & 15) << 2) + (c_diag4[p] & 3);
        y = ((sbo[sb] >> 4) << 2) + (c_diag4[p] >> 2);
    };
#if WB_NZMAP
    auto sb_nonzero = [&](int xs, int ys) { return ((nzm >> (b0 + ys * rs + xs)) & 1ull) != 0ull; };
#else
    auto sb_nonzero = [&](int xs, int ys) {
        for (int yy = 0; yy < 4; yy++)
            for (int xx = 0; xx < 4; xx++)
                if (q[((ys << 2) + yy) * stride + (xs << 2) + xx] != 0) return true;
        return false;
    };
#endif
    // last significant coefficient in coding order (ctu.rs:867-899)
    int last_k = 0, lx = 0, ly = 0;
    {
        int k = nn - 1;
#if WB_NZMAP
        while (k > 15 && !sb_nonzero(sbo[k >> 4] & 15, sbo[k >> 4] >> 4)) k -= 16;  // whole sub-blocks without a level: the map, not 16 loads
#endif
        for (; k >= 0; k--) {
            int x, y;
            pos_of(k, x, y);
            if (q[y * stride + x] != 0 || k == 0) { last_k = k; lx = x; ly = y; break; }
        }
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
            WB_SYN_UNROLL1
            for (int i = 0; i < pre[d]; i++) S.ctx(base + (i >> shift) + off, 1);
            if (pre[d] < cmax) S.ctx(base + (pre[d] >> shift) + off, 0);
        }
        for (int d = 0; d < 2; d++)
            if (pre[d] > 3) S.byp_bits((unsigned)suf[d], sbits[d]);
    }
    int rem = (nn * 7) >> 2;
    const int last_sb = last_k >> 4, last_pos = last_k & 15;
    if ((last_sb > 0 || last_pos > 0) && c_idx == 0) ts.mts_dc_only = false;
    // per-position history of this TB, one byte each (thread-local): pass1 <= 5; absolute levels saturate at 255, which leaves every
    // use exact (they only enter the five-neighbour sums that select a Rice parameter, and those saturate at 31 resp. 51)
    for (int i = 0; i < nn / 4; i++) { reinterpret_cast<uint32_t *>(pass1)[i] = 0u; reinterpret_cast<uint32_t *>(absl)[i] = 0u; }
    auto loc_sums = [&](const uint8_t *a, int x, int y, int &num) {
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
            WB_SYN_UNROLL1
            for (int i = 0; i < pre_; i++) S.byp(1);
            S.byp(0);
            if (rice > 0) S.byp_bits((unsigned)(pv - (pre_ << rice)), rice);
        } else {
            WB_SYN_UNROLL1
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
#if WB_NZMAP
        if (i < last_sb && i > 0 && !sb_nonzero(xs, ys)) {
            // a sub-block without levels codes its sb_coded_flag and nothing else: every absolute level is 0 whatever the
            // quantiser state, pass1 / absl stay 0, and 16 transitions with parity 0 take the state back to where it was
            int csbf = 0;
            if (xs < nsbw - 1) csbf += sb_nonzero(xs + 1, ys);
            if (ys < nsbw - 1) csbf += sb_nonzero(xs, ys + 1);
            S.ctx(CTX_SB_CODED + (c_idx == 0 ? min(csbf, 1) : 2 + min(csbf, 1)), 0);
            continue;
        }
#endif
        int a[16];
        {
            int st = qstate;
            WB_SYN_UNROLL1
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
        WB_SYN_UNROLL1
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
            pass1[y * n + x] = (uint8_t)p1;
            qstate = tr_state(qstate, p1);
            fp1 = p - 1;
        }
        WB_SYN_UNROLL1
        for (int p = fp0; p > fp1; p--) {  // abs_remainder of the positions coded in pass 1
            const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
            if (a[p] > 3) {
                int num;
                const int sum = loc_sums(absl, x, y, num);
                const int rice = c_rice[min(31, max(0, sum - 20))];
                code_rem((a[p] - pass1[y * n + x]) >> 1, rice);
            }
            absl[y * n + x] = (uint8_t)min(a[p], 255);
        }
        WB_SYN_UNROLL1
        for (int p = fp1; p >= 0; p--) {  // dec_abs_level of the rest
            const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
            absl[y * n + x] = (uint8_t)min(a[p], 255);
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
        WB_SYN_UNROLL1
        for (int p = 15; p >= 0; p--) {  // coeff_sign_flag (sign data hiding off)
            if (a[p] > 0) {
                const int x = (xs << 2) + (c_diag4[p] & 3), y = (ys << 2) + (c_diag4[p] >> 2);
                S.byp(q[y * stride + x] < 0);
            }
        }
    }
}

SW_FN bool block_nonzero(const int16_t *q, int stride, int n) {
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            if (q[y * stride + x] != 0) return true;
    return false;
}

// MPM list of the syntax pass (ctu.rs:1498-1635): true neighbours from the final mode map
SW_FN_NOINLINE void code_luma_mode(Sink &S, const PicView &P, int px, int py, int size, int mode) {
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
SW_FN_NOINLINE void code_cu(Sink &S, const SbOrder &SO, const PicView &P, const CtuRecord &rec, const NzMap &nz, int ctu_x, int ctu_y, int x, int y, int size, int tree,
                        TuState &ts, uint8_t *pass1, uint8_t *absl) {
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
    const int by0 = (y >> 2) * 8 + (x >> 2), bc0 = (y >> 3) * 4 + (x >> 3);  // first 4x4 block of the CU in the luma / chroma maps
#if WB_NZMAP
    const int nb = size >> 2, nbc = max(1, size >> 3);  // 4x4 blocks per CU row, luma / chroma
    unsigned long long my = 0ull;
    unsigned mc = 0u;
    for (int r = 0; r < nb; r++) my |= ((1ull << nb) - 1ull) << (by0 + r * 8);
    for (int r = 0; r < nbc; r++) mc |= ((1u << nbc) - 1u) << (bc0 + r * 4);
    const bool ycbf = tree != DUAL_TREE_CHROMA && (nz.y & my) != 0ull;
    const bool cbcbf = tree != DUAL_TREE_LUMA && ((unsigned)nz.cb & mc) != 0u;
    const bool crcbf = tree != DUAL_TREE_LUMA && ((unsigned)nz.cr & mc) != 0u;
#else
    const bool ycbf = tree != DUAL_TREE_CHROMA && block_nonzero(qy, P.W, size);
    const bool cbcbf = tree != DUAL_TREE_LUMA && block_nonzero(qcb, cw, size >> 1);
    const bool crcbf = tree != DUAL_TREE_LUMA && block_nonzero(qcr, cw, size >> 1);
#endif
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
        code_residual(S, SO, qy, P.W, 0, l2, ts, pass1, absl, nz.y, 8, by0);
    }
    if (cbcbf) {
        S.ctx(CTX_TS_FLAG + 1, 0);
        code_residual(S, SO, qcb, cw, 1, l2 - 1, ts, pass1, absl, nz.cb, 4, bc0);
    }
    if (crcbf) {
        S.ctx(CTX_TS_FLAG + 1, 0);
        code_residual(S, SO, qcr, cw, 2, l2 - 1, ts, pass1, absl, nz.cr, 4, bc0);
    }
    if (tree != DUAL_TREE_CHROMA && ts.mts_zero_out && !ts.mts_dc_only) S.ctx(CTX_MTS, 0);  // mts_idx == 0
}

// coding_tree() (ctu_encoder.rs:227-438); QT only, local dual tree at 8x8 -> 4x4.  Written as nested loops over the three
// quad-tree levels (no device recursion: the stack frame stays statically sized).
SW_FN_INLINE bool code_split_flag(Sink &S, const PicView &P, const CtuRecord &rec, int px, int py, int size, int bit) {
    // allow_split_qt holds for 32, 16, 8 (encoder_context.rs:958-971); ctxInc bool_coder.rs:2659-2744
    const bool split = (rec.split_mask >> bit) & 1;
    const bool cl = px > 0 && cu_size_at(P, px - 1, py) < size;
    const bool ca = py > 0 && cu_size_at(P, px, py - 1) < size;
    S.ctx(CTX_SPLIT_CU + (int)cl + (int)ca, split);
    return split;
}

SW_FN void code_ctu(Sink &S, const SbOrder &SO, const PicView &P, const CtuRecord &rec, const NzMap &nz, int ctu_x, int ctu_y, TuState &ts, uint8_t *pass1, uint8_t *absl) {
    if (!code_split_flag(S, P, rec, ctu_x, ctu_y, 32, 0)) {
        code_cu(S, SO, P, rec, nz, ctu_x, ctu_y, 0, 0, 32, SINGLE_TREE, ts, pass1, absl);
        return;
    }
    for (int a = 0; a < 4; a++) {
        const int x16 = (a & 1) * 16, y16 = (a >> 1) * 16;
        if (!code_split_flag(S, P, rec, ctu_x + x16, ctu_y + y16, 16, 1 + a)) {
            code_cu(S, SO, P, rec, nz, ctu_x, ctu_y, x16, y16, 16, SINGLE_TREE, ts, pass1, absl);
            continue;
        }
        for (int b = 0; b < 4; b++) {
            const int x8 = x16 + (b & 1) * 8, y8 = y16 + (b >> 1) * 8;
            if (!code_split_flag(S, P, rec, ctu_x + x8, ctu_y + y8, 8, 5 + 4 * a + b)) {
                code_cu(S, SO, P, rec, nz, ctu_x, ctu_y, x8, y8, 8, SINGLE_TREE, ts, pass1, absl);
                continue;
            }
            for (int i = 0; i < 4; i++) code_cu(S, SO, P, rec, nz, ctu_x, ctu_y, x8 + (i & 1) * 4, y8 + (i >> 1) * 4, 4, DUAL_TREE_LUMA, ts, pass1, absl);
            code_cu(S, SO, P, rec, nz, ctu_x, ctu_y, x8, y8, 8, DUAL_TREE_CHROMA, ts, pass1, absl);
        }
    }
}

}  // namespace wb
