// search.cu — CTU tree search kernel (persistent, wavefront work list) + launch wrapper.  See search_kernel.cuh.
#include <cfloat>
#include <cstdlib>

#include "search_kernel.cuh"

namespace wb {

// ---------------------------------------------------------------------------------------------------------------
// leaf evaluation (block_splitter.rs:886-1078) for SINGLE_TREE 32/16/8 and DUAL_TREE_LUMA 4x4 nodes
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long luma_hdr(const Shared &S, const DevTables *tab, const CtuGeom &g, const Node &nd, int mode, int ck, int root_mode) {
    int lk = luma_kind(S, g, nd, mode, root_mode);
    return nd.tree == SINGLE_TREE ? tab->hdr_single[lk][ck] : tab->hdr_dual[lk];
}

__device__ __forceinline__ void fill_lm(Shared &S, const Node &nd, int mode, int tid) {
    int cells = nd.w >> 2;
    for (int i = tid; i < cells * cells; i += NTHREADS) {
        int yy = i / cells, xx = i - yy * cells;
        S.lm[((nd.y >> 2) + yy) * 8 + (nd.x >> 2) + xx] = (uint8_t)mode;
    }
}
__device__ __forceinline__ void fill_cm(Shared &S, const Node &nd, int mode, int tid) {
    int cells = nd.w >> 3;
    for (int i = tid; i < cells * cells; i += NTHREADS) {
        int yy = i / cells, xx = i - yy * cells;
        S.cm[((nd.y >> 3) + yy) * 4 + (nd.x >> 3) + xx] = (uint8_t)mode;
    }
}

__device__ __noinline__ float leaf_eval(Shared &S, const SearchParams &P, const CtuGeom &g, const Node &nd, int &root_mode, bool is_root) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DevTables *tab = P.tab;
    const WarpScratch ws = warp_scratch(S, warp);
    const int ncomp = nd.tree == DUAL_TREE_LUMA ? 1 : 3;
    // ---- phase 0: reference samples
    for (int t = warp; t < ncomp; t += NW) build_refs(S, g, nd, t, lane);
    __syncthreads();
    // ---- phase 1: planar / DC full evaluations, 13 coarse angular SADs
    {
        const int nfull = 2 * ncomp, ntask = nfull + 13 * ncomp;
        for (int t = warp; t < ntask; t += NW) {
            if (t < nfull) {
                int mode, c;
                if (t < 2) { mode = t; c = 0; }
                else { mode = (t - 2) >> 1; c = 1 + ((t - 2) & 1); }
                unsigned ssd; int rate;
                full_task(S, tab, g, nd, c, mode, false, ws, lane, ssd, rate);
                if (lane == 0) { S.r_ssd[t] = ssd; S.r_rate[t] = rate; }
            } else {
                int u = t - nfull;
                int c = u / 13, mi = u - c * 13;
                unsigned sad = sad_task(S, g, nd, c, c_cand15[2 + mi], ws, lane);
                if (lane == 0) S.r_sad[t] = sad;
            }
        }
    }
    __syncthreads();
    float cost_pl, cost_dc;
    int cur;
    float cur_cost;
    {
        unsigned ssd0 = S.r_ssd[0], ssd1 = S.r_ssd[1];
        long long r0 = S.r_rate[0], r1 = S.r_rate[1];
        if (ncomp == 3) {
            ssd0 += S.r_ssd[2] + S.r_ssd[3]; r0 += (long long)S.r_rate[2] + S.r_rate[3];
            ssd1 += S.r_ssd[4] + S.r_ssd[5]; r1 += (long long)S.r_rate[4] + S.r_rate[5];
        }
        cost_pl = rd_cost(ssd0, r0 + luma_hdr(S, tab, g, nd, 0, 0, root_mode), tab->lambda_rd);
        cost_dc = rd_cost(ssd1, r1 + luma_hdr(S, tab, g, nd, 1, 0, root_mode), tab->lambda_rd);
        const int nfull = 2 * ncomp;
        int best = 0;
        float bc = 0.f;
        for (int i = 0; i < 13; i++) {
            unsigned s = S.r_sad[nfull + i];
            if (ncomp == 3) s += S.r_sad[nfull + 13 + i] + S.r_sad[nfull + 26 + i];
            float c = __uint2float_rn(s);
            if (i == 0 || c < bc) { bc = c; best = i; }
        }
        cur = c_cand15[2 + best];
        cur_cost = bc;
    }
    __syncthreads();  // results consumed before the next phase overwrites them
    // ---- phases 2,3: SAD refinement +-2, +-1 (step_search aux=true, block_splitter.rs:905-973)
    for (int step = 2; step >= 1; step >>= 1) {
        const bool v0 = !(cur < 2 + step), v1 = !(cur + step > 66);
        for (int t = warp; t < 2 * ncomp; t += NW) {
            int cand = t / ncomp, c = t - cand * ncomp;
            if (cand == 0 ? v0 : v1) {
                unsigned sad = sad_task(S, g, nd, c, cand == 0 ? cur - step : cur + step, ws, lane);
                if (lane == 0) S.r_sad[t] = sad;
            }
        }
        __syncthreads();
        float c0 = FLT_MAX, c1 = FLT_MAX;
        if (v0) { unsigned s = 0; for (int c = 0; c < ncomp; c++) s += S.r_sad[c]; c0 = __uint2float_rn(s); }
        if (v1) { unsigned s = 0; for (int c = 0; c < ncomp; c++) s += S.r_sad[ncomp + c]; c1 = __uint2float_rn(s); }
        float mn = fminf(fminf(cur_cost, c0), c1);
        if (cur_cost == mn) {
        } else if (c0 == mn) { cur -= step; cur_cost = c0; }
        else { cur += step; cur_cost = c1; }
        __syncthreads();
    }
    // ---- phase 4: full evaluation of dir, dir-1, dir+1 (step_search aux=false)
    int dir = cur;
    float dir_cost;
    {
        const bool v0 = !(dir < 3), v1 = !(dir + 1 > 66);
        const int ntask = 3 * ncomp;
        for (int t = warp; t < ntask; t += NW) {
            int cand, c;
            if (t < 3) { cand = t; c = 0; }
            else { cand = (t - 3) >> 1; c = 1 + ((t - 3) & 1); }
            bool valid = cand == 0 || (cand == 1 ? v0 : v1);
            if (valid) {
                int mode = cand == 0 ? dir : (cand == 1 ? dir - 1 : dir + 1);
                unsigned ssd; int rate;
                full_task(S, tab, g, nd, c, mode, false, ws, lane, ssd, rate);
                if (lane == 0) { S.r_ssd[t] = ssd; S.r_rate[t] = rate; }
            }
        }
        __syncthreads();
        float cc[3];
#pragma unroll
        for (int cand = 0; cand < 3; cand++) {
            bool valid = cand == 0 || (cand == 1 ? v0 : v1);
            cc[cand] = FLT_MAX;
            if (valid) {
                unsigned ssd = S.r_ssd[cand];
                long long r = S.r_rate[cand];
                if (ncomp == 3) { ssd += S.r_ssd[3 + 2 * cand] + S.r_ssd[4 + 2 * cand]; r += (long long)S.r_rate[3 + 2 * cand] + S.r_rate[4 + 2 * cand]; }
                int mode = cand == 0 ? dir : (cand == 1 ? dir - 1 : dir + 1);
                cc[cand] = rd_cost(ssd, r + luma_hdr(S, tab, g, nd, mode, 0, root_mode), tab->lambda_rd);
            }
        }
        float mn = fminf(fminf(cc[0], cc[1]), cc[2]);
        if (cc[0] == mn) { dir_cost = cc[0]; }
        else if (cc[1] == mn) { dir -= 1; dir_cost = cc[1]; }
        else { dir += 1; dir_cost = cc[2]; }
        __syncthreads();
    }
    // ---- winner among planar, DC, dir (first minimum)
    float min_cost = fminf(fminf(cost_pl, cost_dc), dir_cost);
    const int mode = cost_pl == min_cost ? 0 : (cost_dc == min_cost ? 1 : dir);
    if (is_root) root_mode = mode;
    // ---- phase 5: luma redo (commit) + chroma DM full evaluation (commit)
    for (int t = warp; t < ncomp; t += NW) {
        unsigned ssd; int rate;
        full_task(S, tab, g, nd, t, mode, true, ws, lane, ssd, rate);
        if (lane == 0) { S.r_ssd[t] = ssd; S.r_rate[t] = rate; }
    }
    fill_lm(S, nd, mode, tid);
    __syncthreads();
    if (ncomp == 1) return min_cost;
    const unsigned ssdY = S.r_ssd[0], ssdDM = S.r_ssd[1] + S.r_ssd[2];
    const long long rateY = S.r_rate[0], rateDM = (long long)S.r_rate[1] + S.r_rate[2];
    const float cost_dm = rd_cost(ssdDM, rateDM + tab->hdr_chroma[0], tab->lambda_rd_c);
    // ---- phase 5b: CCLM down-sampled luma of the committed luma reconstruction
    if (warp == NW - 1) cclm_downsample(S, g, nd, lane);
    __syncthreads();
    // ---- phase 6: CCLM SADs in the order LT, T, L
    for (int t = warp; t < 6; t += NW) {
        const int mi = t >> 1, c = 1 + (t & 1);
        const int cm = mi == 0 ? MODE_LT_CCLM : (mi == 1 ? MODE_T_CCLM : MODE_L_CCLM);
        unsigned sad = sad_task(S, g, nd, c, cm, ws, lane);
        if (lane == 0) S.r_sad[t] = sad;
    }
    __syncthreads();
    int cclm_mode;
    {
        float lt = __uint2float_rn(S.r_sad[0] + S.r_sad[1]), t = __uint2float_rn(S.r_sad[2] + S.r_sad[3]), l = __uint2float_rn(S.r_sad[4] + S.r_sad[5]);
        if (lt <= t && lt <= l) cclm_mode = MODE_LT_CCLM;
        else if (t <= l) cclm_mode = MODE_T_CCLM;
        else cclm_mode = MODE_L_CCLM;
    }
    // ---- phase 7: CCLM full evaluation (no commit)
    for (int t = warp; t < 2; t += NW) {
        unsigned ssd; int rate;
        full_task(S, tab, g, nd, 1 + t, cclm_mode, false, ws, lane, ssd, rate);
        if (lane == 0) { S.r_ssd[8 + t] = ssd; S.r_rate[8 + t] = rate; }
    }
    __syncthreads();
    const unsigned ssdCC = S.r_ssd[8] + S.r_ssd[9];
    const long long rateCC = (long long)S.r_rate[8] + S.r_rate[9];
    const int ck = 1 + (cclm_mode - MODE_LT_CCLM);
    const float cost_cclm = rd_cost(ssdCC, rateCC + tab->hdr_chroma[ck], tab->lambda_rd_c);
    const float cmn = fminf(cost_dm, cost_cclm);
    float final_cost;
    if (cost_dm == cmn) {
        fill_cm(S, nd, mode, tid);
        final_cost = rd_cost(ssdY + ssdDM, rateY + rateDM + luma_hdr(S, tab, g, nd, mode, 0, root_mode), tab->lambda_rd);
    } else {
        // ---- phase 8: commit the CCLM chroma
        for (int t = warp; t < 2; t += NW) {
            unsigned ssd; int rate;
            full_task(S, tab, g, nd, 1 + t, cclm_mode, true, ws, lane, ssd, rate);
        }
        fill_cm(S, nd, cclm_mode, tid);
        final_cost = rd_cost(ssdY + ssdCC, rateY + rateCC + luma_hdr(S, tab, g, nd, mode, ck, root_mode), tab->lambda_rd);
    }
    __syncthreads();
    return final_cost;
}

// ---------------------------------------------------------------------------------------------------------------
// the 8x8 DUAL_TREE_CHROMA coding tree that follows four 4x4 luma CUs (block_splitter.rs:794-885)
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ float chroma_ct_eval(Shared &S, const SearchParams &P, const CtuGeom &g, const Node &nd) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DevTables *tab = P.tab;
    const WarpScratch ws = warp_scratch(S, warp);
    // luma CU covering the parent's centre sample = the bottom-right 4x4 (ctu.rs:2372-2396)
    const int dm = S.lm[((nd.y >> 2) + 1) * 8 + (nd.x >> 2) + 1];
    for (int t = warp; t < 3; t += NW) {
        if (t < 2) build_refs(S, g, nd, 1 + t, lane);
        else cclm_downsample(S, g, nd, lane);
    }
    __syncthreads();
    for (int t = warp; t < 8; t += NW) {
        if (t < 2) {
            unsigned ssd; int rate;
            full_task(S, tab, g, nd, 1 + t, dm, true, ws, lane, ssd, rate);
            if (lane == 0) { S.r_ssd[t] = ssd; S.r_rate[t] = rate; }
        } else {
            const int u = t - 2;
            const int mi = u >> 1, c = 1 + (u & 1);
            const int cm = mi == 0 ? MODE_LT_CCLM : (mi == 1 ? MODE_T_CCLM : MODE_L_CCLM);
            unsigned sad = sad_task(S, g, nd, c, cm, ws, lane);
            if (lane == 0) S.r_sad[u] = sad;
        }
    }
    __syncthreads();
    int cclm_mode;
    {
        float lt = __uint2float_rn(S.r_sad[0] + S.r_sad[1]), t = __uint2float_rn(S.r_sad[2] + S.r_sad[3]), l = __uint2float_rn(S.r_sad[4] + S.r_sad[5]);
        if (lt <= t && lt <= l) cclm_mode = MODE_LT_CCLM;
        else if (t <= l) cclm_mode = MODE_T_CCLM;
        else cclm_mode = MODE_L_CCLM;
    }
    for (int t = warp; t < 2; t += NW) {
        unsigned ssd; int rate;
        full_task(S, tab, g, nd, 1 + t, cclm_mode, false, ws, lane, ssd, rate);
        if (lane == 0) { S.r_ssd[8 + t] = ssd; S.r_rate[8 + t] = rate; }
    }
    __syncthreads();
    const float cost_dm = rd_cost(S.r_ssd[0] + S.r_ssd[1], (long long)S.r_rate[0] + S.r_rate[1] + tab->hdr_chroma[0], tab->lambda_rd_c);
    const int ck = 1 + (cclm_mode - MODE_LT_CCLM);
    const float cost_cclm = rd_cost(S.r_ssd[8] + S.r_ssd[9], (long long)S.r_rate[8] + S.r_rate[9] + tab->hdr_chroma[ck], tab->lambda_rd_c);
    const float mn = fminf(cost_dm, cost_cclm);
    if (cost_dm == mn) {
        fill_cm(S, nd, dm, tid);
    } else {
        for (int t = warp; t < 2; t += NW) {
            unsigned ssd; int rate;
            full_task(S, tab, g, nd, 1 + t, cclm_mode, true, ws, lane, ssd, rate);
        }
        fill_cm(S, nd, cclm_mode, tid);
    }
    __syncthreads();
    return mn;
}

// ---------------------------------------------------------------------------------------------------------------
// no-split state save / restore (block_splitter.rs:1085-1109, 1125-1145) — here also levels and modes, because the
// CUDA path keeps the tree as flat arrays instead of cloning CodingTree objects
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sv_off_y(int d) { return d == 0 ? 0 : (d == 1 ? 1024 : 1280); }
__device__ __forceinline__ int sv_off_c(int d) { return d == 0 ? 0 : (d == 1 ? 256 : 320); }

__device__ __noinline__ void save_node(Shared &S, const Node &nd, int d, int tid) {
    const int w = nd.w, oy = sv_off_y(d), oc = sv_off_c(d);
    for (int i = tid; i < w * w; i += NTHREADS) {
        int y = i / w, x = i - y * w;
        S.svRecY[oy + i] = RY(S, nd.x + x, nd.y + y);
        S.svLvY[oy + i] = S.lvY[(nd.y + y) * 32 + nd.x + x];
    }
    const int cw = w >> 1, bx = nd.x >> 1, by = nd.y >> 1;
    for (int i = tid; i < 2 * cw * cw; i += NTHREADS) {
        int c = i >= cw * cw, j = i - c * cw * cw;
        int y = j / cw, x = j - y * cw;
        S.svRecC[c][oc + j] = RC(S, 1 + c, bx + x, by + y);
        S.svLvC[c][oc + j] = S.lvC[c][(by + y) * 16 + bx + x];
    }
    for (int i = tid; i < 64; i += NTHREADS) S.svLm[d][i] = S.lm[i];
    for (int i = tid; i < 16; i += NTHREADS) S.svCm[d][i] = S.cm[i];
}
__device__ __noinline__ void restore_node(Shared &S, const Node &nd, int d, int tid) {
    const int w = nd.w, oy = sv_off_y(d), oc = sv_off_c(d);
    for (int i = tid; i < w * w; i += NTHREADS) {
        int y = i / w, x = i - y * w;
        RY(S, nd.x + x, nd.y + y) = S.svRecY[oy + i];
        S.lvY[(nd.y + y) * 32 + nd.x + x] = S.svLvY[oy + i];
    }
    const int cw = w >> 1, bx = nd.x >> 1, by = nd.y >> 1;
    for (int i = tid; i < 2 * cw * cw; i += NTHREADS) {
        int c = i >= cw * cw, j = i - c * cw * cw;
        int y = j / cw, x = j - y * cw;
        RC(S, 1 + c, bx + x, by + y) = S.svRecC[c][oc + j];
        S.lvC[c][(by + y) * 16 + bx + x] = S.svLvC[c][oc + j];
    }
    const int cells = w >> 2;
    for (int i = tid; i < cells * cells; i += NTHREADS) {
        int yy = i / cells, xx = i - yy * cells;
        int idx = ((nd.y >> 2) + yy) * 8 + (nd.x >> 2) + xx;
        S.lm[idx] = S.svLm[d][idx];
    }
    const int cc = w >> 3;
    for (int i = tid; i < cc * cc; i += NTHREADS) {
        int yy = i / cc, xx = i - yy * cc;
        int idx = ((nd.y >> 3) + yy) * 4 + (nd.x >> 3) + xx;
        S.cm[idx] = S.svCm[d][idx];
    }
}

// CodingTree::split(SPLIT_QT) child geometry + availability flags (ctu.rs:1960-2064, 2083-2188; H9)
__device__ __forceinline__ Node qt_child(const CtuGeom &g, const Node &p, int i, int tree) {
    Node c;
    c.w = p.w >> 1;
    c.x = p.x + (i & 1) * c.w;
    c.y = p.y + (i >> 1) * c.w;
    c.tree = tree;
    const int ax = g.cx + c.x, ay = g.cy + c.y;
    if (ax + c.w >= g.W) c.ar = false;
    else if (i == 0) c.ar = 0 < ay;
    else if (i == 1) c.ar = p.ar;
    else if (i == 2) c.ar = true;
    else c.ar = false;
    if (ay + c.w >= g.H) c.bl = false;
    else if (i == 1 || i == 3) c.bl = false;
    else if (i == 0) c.bl = 0 < ax;
    else c.bl = p.bl;
    return c;
}

// ---------------------------------------------------------------------------------------------------------------
// one CTU: split_ct(root, max_depth)  (block_splitter.rs:782-1154)
// ---------------------------------------------------------------------------------------------------------------
__device__ void ctu_search(Shared &S, const SearchParams &P, const CtuGeom &g, unsigned &split_mask_out, float &cost_out) {
    const int tid = threadIdx.x;
    const int md = P.max_depth;
    int root_mode = 0;
    unsigned mask = 0;
    Node n32;
    n32.x = 0; n32.y = 0; n32.w = 32; n32.tree = SINGLE_TREE;
    n32.bl = false;
    n32.ar = (g.cx + 32 >= g.W) ? false : (0 < g.cy && g.cx + 32 < g.W);
    float cost32 = leaf_eval(S, P, g, n32, root_mode, true);
    if (md >= 1) {
        save_node(S, n32, 0, tid);
        __syncthreads();
        const unsigned mask32 = mask;
        float split32 = 0.0f;
        for (int a = 0; a < 4; a++) {
            Node n16 = qt_child(g, n32, a, SINGLE_TREE);
            float cost16 = leaf_eval(S, P, g, n16, root_mode, false);
            if (md >= 2) {
                save_node(S, n16, 1, tid);
                __syncthreads();
                const unsigned mask16 = mask;
                float split16 = 0.0f;
                for (int b = 0; b < 4; b++) {
                    Node n8 = qt_child(g, n16, b, SINGLE_TREE);
                    float cost8 = leaf_eval(S, P, g, n8, root_mode, false);
                    if (md >= 3) {
                        save_node(S, n8, 2, tid);
                        __syncthreads();
                        float split8 = 0.0f;
                        for (int c = 0; c < 4; c++) {
                            Node n4 = qt_child(g, n8, c, DUAL_TREE_LUMA);
                            split8 = __fadd_rn(split8, leaf_eval(S, P, g, n4, root_mode, false));
                        }
                        Node nc = n8;  // local dual tree: chroma CT of the parent's size (ctu.rs:2031-2055)
                        nc.tree = DUAL_TREE_CHROMA;
                        nc.ar = (g.cx + nc.x + nc.w >= g.W) ? false : n8.ar;
                        nc.bl = (g.cy + nc.y + nc.w >= g.H) ? false : n8.bl;
                        split8 = __fadd_rn(split8, chroma_ct_eval(S, P, g, nc));
                        if (split8 > cost8) {
                            restore_node(S, n8, 2, tid);
                            __syncthreads();
                        } else {
                            mask |= 1u << (5 + a * 4 + b);
                            cost8 = split8;
                        }
                    }
                    split16 = __fadd_rn(split16, cost8);
                }
                if (split16 > cost16) {
                    restore_node(S, n16, 1, tid);
                    mask = mask16;
                    __syncthreads();
                } else {
                    mask |= 1u << (1 + a);
                    cost16 = split16;
                }
            }
            split32 = __fadd_rn(split32, cost16);
        }
        if (split32 > cost32) {
            restore_node(S, n32, 0, tid);
            mask = mask32;
            __syncthreads();
        } else {
            mask |= 1u;
            cost32 = split32;
        }
    }
    split_mask_out = mask;
    cost_out = cost32;
}

// ---------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ void init_tables(Shared &S, const DevTables *tab, int tid) {
    for (int l2 = 2; l2 <= 5; l2++) {
        const int n = 1 << l2, o = tab_off(l2);
        for (int e = tid; e < n * n; e += NTHREADS) {
            int i = e >> l2, x = e & (n - 1);
            int v;
            if (i == 0) v = 64;
            else {
                int m = (i * (2 * x + 1) * (32 / n)) % 128;  // angle in units of pi/64
                int s = 1;
                if (m > 64) m = 128 - m;
                if (m > 32) { m = 64 - m; s = -1; }
                v = s * c_cos32[m];
            }
            S.tb.T[o + i * n + x] = (int8_t)v;
            S.tb.Tt[o + x * n + i] = (int8_t)v;
        }
    }
    // scan tables: sub-block order and in-sub-block order are both up-right diagonal scans (ctu.rs:53-77)
    if (tid < 4) {
        const int l2 = 2 + tid, n = 1 << l2, nsb = n >> 2, o = tab_off(l2);
        // diagonal order of the 4x4 coefficients
        uint8_t cx4[16], cy4[16];
        {
            int i = 0, x = 0, y = 0;
            while (i < 16) {
                while (y >= 0) {
                    if (x < 4 && y < 4) { cx4[i] = (uint8_t)x; cy4[i] = (uint8_t)y; i++; }
                    y--; x++;
                }
                y = x; x = 0;
            }
        }
        int sb = 0, x = 0, y = 0;
        while (sb < nsb * nsb) {
            while (y >= 0) {
                if (x < nsb && y < nsb) {
                    for (int p = 0; p < 16; p++) S.tb.scan[o + sb * 16 + p] = (uint16_t)((y * 4 + cy4[p]) * n + x * 4 + cx4[p]);
                    sb++;
                }
                y--; x++;
            }
            y = x; x = 0;
        }
    }
    for (int i = tid; i < 64; i += NTHREADS) { S.tb.ldq[i] = tab->ldq[i]; S.tb.lv[i] = tab->lv[i]; }
    if (tid < 32) {
        unsigned pk = 0;
        for (int i = 0; i < 4; i++) pk |= ((unsigned)(uint8_t)c_fC[tid][i]) << (8 * i);
        S.tb.fc[tid] = (int)pk;
    }
}

__device__ __forceinline__ int ld_relaxed(const int *p) {  // polling load: no L1 invalidation per poll (the acquire fence follows the loop)
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

extern "C" __global__ void __launch_bounds__(NTHREADS, WB_MINB) wrenc_b200_search_kernel(SearchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &S = *reinterpret_cast<Shared *>(smem_raw);
    const int tid = threadIdx.x;
    init_tables(S, P.tab, tid);
    __syncthreads();
    const int W = P.W, H = P.H, Wc = P.Wc;
    const size_t pic_samples = (size_t)W * H * 3 / 2;
    const int cw = W >> 1, chh = H >> 1;
    for (;;) {
        if (tid == 0) S.item = (int)atomicAdd(P.counter, 1u);
        __syncthreads();
        const int item = S.item;
        if (item >= P.n_items) break;
        const uint32_t it = P.items[item];
        const int pic = it >> 16, cyi = (it >> 8) & 255, cxi = it & 255;
        CtuGeom g;
        g.cx = cxi * 32; g.cy = cyi * 32; g.W = W; g.H = H;
        int *done = P.done + (size_t)pic * Wc * P.Hc;
        // wavefront dependencies: left CTU and above-right CTU (above when in the last column) must be final
        if (tid == 0) {
            if (cxi > 0) while (ld_relaxed(&done[cyi * Wc + cxi - 1]) != P.epoch) __nanosleep(1000);
            if (cyi > 0) {
                const int ax = min(cxi + 1, Wc - 1);
                while (ld_relaxed(&done[(cyi - 1) * Wc + ax]) != P.epoch) __nanosleep(1000);
            }
            __threadfence();  // acquire side: order the halo loads after the flag observation
        }
        __syncthreads();
        const uint8_t *oY = P.orig + (size_t)pic * pic_samples, *oCb = oY + (size_t)W * H, *oCr = oCb + (size_t)cw * chh;
        uint8_t *rY = P.rec + (size_t)pic * pic_samples, *rCb = rY + (size_t)W * H, *rCr = rCb + (size_t)cw * chh;
        // ---- stage the CTU: source samples, neighbouring reconstruction, left-CTU modes
        {
            for (int i = tid; i < 256; i += NTHREADS) {  // 32 rows x 8 words
                int y = i >> 3, x4 = (i & 7) * 4;
                *reinterpret_cast<uint32_t *>(&S.orgY[y * 32 + x4]) = __ldg(reinterpret_cast<const uint32_t *>(oY + (size_t)(g.cy + y) * W + g.cx + x4));
            }
            for (int i = tid; i < 128; i += NTHREADS) {
                int c = i >> 6, t = i & 63, yy = t >> 2, xx = (t & 3) * 4;
                const uint8_t *src = (c ? oCr : oCb) + (size_t)((g.cy >> 1) + yy) * cw + (g.cx >> 1) + xx;
                *reinterpret_cast<uint32_t *>(&S.orgC[c][yy * 16 + xx]) = __ldg(reinterpret_cast<const uint32_t *>(src));
            }
            // luma halo: rows -2,-1 (cols -4..63) and cols -4..-1 of rows 0..31
            for (int i = tid; i < 2 * RY_STRIDE + 32 * 4; i += NTHREADS) {
                int xx, yy;
                if (i < 2 * RY_STRIDE) { yy = -2 + i / RY_STRIDE; xx = -4 + i % RY_STRIDE; }
                else { int j = i - 2 * RY_STRIDE; yy = j >> 2; xx = -4 + (j & 3); }
                int ax = g.cx + xx, ay = g.cy + yy;
                uint8_t v = 0;
                if (ax >= 0 && ax < W && ay >= 0) v = __ldcg(rY + (size_t)ay * W + ax);
                RY(S, xx, yy) = v;
            }
            for (int i = tid; i < 2 * (RC_STRIDE + 16 * 4); i += NTHREADS) {
                int c = i >= (RC_STRIDE + 16 * 4), j = i - c * (RC_STRIDE + 16 * 4);
                int xx, yy;
                if (j < RC_STRIDE) { yy = -1; xx = -4 + j; }
                else { int q = j - RC_STRIDE; yy = q >> 2; xx = -4 + (q & 3); }
                int ax = (g.cx >> 1) + xx, ay = (g.cy >> 1) + yy;
                uint8_t v = 0;
                if (ax >= 0 && ax < cw && ay >= 0) v = __ldcg((c ? rCr : rCb) + (size_t)ay * cw + ax);
                RC(S, 1 + c, xx, yy) = v;
            }
            if (tid < 8) S.leftModes[tid] = g.cx > 0 ? __ldcg(P.mode_map + (size_t)pic * (W >> 2) * (H >> 2) + (size_t)((g.cy >> 2) + tid) * (W >> 2) + (g.cx >> 2) - 1) : 0;
            for (int i = tid; i < 1024; i += NTHREADS) S.lvY[i] = 0;
            for (int i = tid; i < 256; i += NTHREADS) { S.lvC[0][i] = 0; S.lvC[1][i] = 0; }
            for (int i = tid; i < 64; i += NTHREADS) S.lm[i] = 0;
            for (int i = tid; i < 16; i += NTHREADS) S.cm[i] = 0;
        }
        __syncthreads();
        unsigned split_mask;
        float cost;
        ctu_search(S, P, g, split_mask, cost);
        __syncthreads();
        // ---- write back: reconstruction, levels, modes, record
        {
            for (int i = tid; i < 256; i += NTHREADS) {
                int y = i >> 3, x4 = (i & 7) * 4;
                uint32_t v = (uint32_t)RY(S, x4, y) | ((uint32_t)RY(S, x4 + 1, y) << 8) | ((uint32_t)RY(S, x4 + 2, y) << 16) | ((uint32_t)RY(S, x4 + 3, y) << 24);
                *reinterpret_cast<uint32_t *>(rY + (size_t)(g.cy + y) * W + g.cx + x4) = v;
            }
            for (int i = tid; i < 128; i += NTHREADS) {
                int c = i >> 6, t = i & 63, yy = t >> 2, xx = (t & 3) * 4;
                uint32_t u = (uint32_t)RC(S, 1 + c, xx, yy) | ((uint32_t)RC(S, 1 + c, xx + 1, yy) << 8) | ((uint32_t)RC(S, 1 + c, xx + 2, yy) << 16) |
                             ((uint32_t)RC(S, 1 + c, xx + 3, yy) << 24);
                *reinterpret_cast<uint32_t *>((c ? rCr : rCb) + (size_t)((g.cy >> 1) + yy) * cw + (g.cx >> 1) + xx) = u;
            }
            int16_t *lY = P.lev + (size_t)pic * pic_samples, *lCb = lY + (size_t)W * H, *lCr = lCb + (size_t)cw * chh;
            for (int i = tid; i < 512; i += NTHREADS) {  // luma levels, 2 per iteration
                int yy = i >> 4, xx = (i & 15) * 2;
                *reinterpret_cast<uint32_t *>(lY + (size_t)(g.cy + yy) * W + g.cx + xx) = *reinterpret_cast<const uint32_t *>(&S.lvY[yy * 32 + xx]);
            }
            for (int i = tid; i < 256; i += NTHREADS) {
                int c = i >> 7, t = i & 127, yy = t >> 3, xx = (t & 7) * 2;
                *reinterpret_cast<uint32_t *>((c ? lCr : lCb) + (size_t)((g.cy >> 1) + yy) * cw + (g.cx >> 1) + xx) =
                    *reinterpret_cast<const uint32_t *>(&S.lvC[c][yy * 16 + xx]);
            }
            CtuRecord *r = P.records + (size_t)pic * Wc * P.Hc + cyi * Wc + cxi;
            for (int i = tid; i < 64; i += NTHREADS) {
                r->luma_mode[i] = S.lm[i];
                P.mode_map[(size_t)pic * (W >> 2) * (H >> 2) + (size_t)((g.cy >> 2) + (i >> 3)) * (W >> 2) + (g.cx >> 2) + (i & 7)] = S.lm[i];
            }
            for (int i = tid; i < 16; i += NTHREADS) r->chroma_mode[i] = S.cm[i];
            if (tid == 0) { r->split_mask = split_mask; r->cost = cost; }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&done[cyi * Wc + cxi], P.epoch);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-block entry points (parity tests of the block ops; north_star correctness level 1)
// ---------------------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(NTHREADS, 1) wrenc_b200_block_kernel(BlockParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &S = *reinterpret_cast<Shared *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    init_tables(S, P.tab, tid);
    __syncthreads();
    const WarpScratch ws = warp_scratch(S, 0);
    if (P.op == 0) {  // predict one component of one TU from a picture's reconstruction
        CtuGeom g;
        g.cx = P.x & ~31; g.cy = P.y & ~31; g.W = P.W; g.H = P.H;
        const int cw = P.W >> 1, chh = P.H >> 1;
        const uint8_t *rY = P.rec, *rCb = rY + (size_t)P.W * P.H, *rCr = rCb + (size_t)cw * chh;
        for (int i = tid; i < RY_ROWS * RY_STRIDE; i += NTHREADS) {
            int yy = i / RY_STRIDE - RY_Y0, xx = i % RY_STRIDE - RY_X0;
            int ax = g.cx + xx, ay = g.cy + yy;
            S.recY[i] = (ax >= 0 && ax < P.W && ay >= 0 && ay < P.H) ? rY[(size_t)ay * P.W + ax] : 0;
        }
        for (int i = tid; i < 2 * RC_ROWS * RC_STRIDE; i += NTHREADS) {
            int c = i >= RC_ROWS * RC_STRIDE, j = i - c * RC_ROWS * RC_STRIDE;
            int yy = j / RC_STRIDE - RC_Y0, xx = j % RC_STRIDE - RC_X0;
            int ax = (g.cx >> 1) + xx, ay = (g.cy >> 1) + yy;
            S.recC[c][j] = (ax >= 0 && ax < cw && ay >= 0 && ay < chh) ? (c ? rCr : rCb)[(size_t)ay * cw + ax] : 0;
        }
        __syncthreads();
        Node nd;
        nd.x = P.x - g.cx; nd.y = P.y - g.cy; nd.w = P.w; nd.tree = P.tree; nd.ar = P.ar != 0; nd.bl = P.bl != 0;
        if (warp == 0) {
            const int c = P.c, n = nd.w >> (c != 0);
            if (P.mode > 66) cclm_downsample(S, g, nd, lane);
            else build_refs(S, g, nd, c, lane);
            __syncwarp();
            PredCtx pc;
            pred_setup(S, g, nd, c, P.mode, ws.refx, lane, pc);
            for (int i = lane; i < n * n; i += 32) P.out8[i] = (uint8_t)pred_sample(S, pc, i % n, i / n);
        }
        return;
    }
    if (warp != 0) return;
    const int l2 = P.l2, n = 1 << l2, nn = n * n, to = tab_off(l2);
    for (int blk = blockIdx.x; blk < P.count; blk += gridDim.x) {
        const int16_t *in = P.in + (size_t)blk * nn;
        int16_t *out = P.out16 + (size_t)blk * nn;
        for (int i = lane; i < nn; i += 32) ws.A[i] = in[i];
        __syncwarp();
        if (P.op == 1) {
            mm_rows(S.tb.Tt + to, ws.A, ws.B, n, l2, 1 << (l2 - 2), l2 - 1, false, lane);
            __syncwarp();
            mm_cols(S.tb.T + to, ws.B, ws.A, n, l2, 1 << (l2 + 5), l2 + 6, false, lane);
            __syncwarp();
            for (int i = lane; i < nn; i += 32) out[i] = ws.A[i];
        } else if (P.op == 2) {
            mm_cols(S.tb.Tt + to, ws.A, ws.B, n, l2, 64, 7, true, lane);  // Tt rows are T columns: V[y][x] = sum_i T[i][y] D[i][x]
            __syncwarp();
            mm_rows(S.tb.T + to, ws.B, ws.A, n, l2, 2048, 12, false, lane);  // R[y][x] = sum_i T[i][x] V[y][i]
            __syncwarp();
            for (int i = lane; i < nn; i += 32) out[i] = ws.A[i];
        } else if (P.op == 3) {
            int rate; bool any;
            trellis(S, P.tab, ws.A, l2, ws.Wd, ws.B, lane, rate, any);
            for (int i = lane; i < nn; i += 32) out[i] = ws.B[i];
            if (lane == 0) P.outi[blk] = rate;
        } else if (P.op == 4) {
            const int sh = l2 + 4, off = 1 << (sh - 1), ls = P.tab->ls;
            for (int i = lane; i < nn; i += 32) out[i] = (int16_t)min(32767, max(-32768, ((int)ws.A[i] * ls + off) >> sh));
        }
        __syncwarp();
    }
}

cudaError_t launch_block(const BlockParams &P, int grid, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(wrenc_b200_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared));
    if (e != cudaSuccess) return e;
    wrenc_b200_block_kernel<<<grid, NTHREADS, sizeof(Shared), stream>>>(P);
    return cudaGetLastError();
}

size_t search_smem_bytes() { return sizeof(Shared); }

static size_t smem_request() {  // dev-time knob: WRENC_B200_SMEM_PAD bytes of extra dynamic shared memory lower the CTAs per SM
    size_t pad = 0;
    if (const char *e = getenv("WRENC_B200_SMEM_PAD")) pad = (size_t)atol(e);
    return sizeof(Shared) + pad;
}

cudaError_t launch_search(const SearchParams &P, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wrenc_b200_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_request());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    wrenc_b200_search_kernel<<<grid, NTHREADS, smem_request(), stream>>>(P);
    return cudaGetLastError();
}

int search_ctas_per_sm() {
    int n = 0;
    cudaFuncSetAttribute(wrenc_b200_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_request());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wrenc_b200_search_kernel, NTHREADS, smem_request());
    return n;
}

}  // namespace wb
