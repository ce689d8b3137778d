// search.cu — CTU tree search kernel (persistent, wavefront work list) + launch wrapper.  See search_kernel.cuh.
#include <cfloat>
#include <cstdlib>

#include "search_kernel.cuh"

namespace wb {

// Dev-time phase timer (-DWB_PROFILE builds only, tools/phase_profile.py): thread 0 of every CTA adds the cycles between two
// barriers to g_prof[slot]; slot = 16 * node kind (0: 32x32, 1: 16x16, 2: 8x8, 3: 4x4 luma, 4: chroma CT, 5: outside the tree) + phase.
#ifdef WB_PROFILE
__device__ unsigned long long g_prof[128];
#define WB_PROF(slot)                                                              \
    do {                                                                           \
        if (threadIdx.x == 0) {                                                    \
            const long long t_ = clock64();                                        \
            atomicAdd(&g_prof[slot], (unsigned long long)(t_ - S.prof_last));      \
            S.prof_last = t_;                                                      \
        }                                                                          \
    } while (0)
#else
#define WB_PROF(slot)
#endif

// ---------------------------------------------------------------------------------------------------------------
// Lock-step execution of KC independent CTUs per CTA.
//
// The search is exhaustive (no early termination), so every CTU walks the same node sequence; only the data-dependent
// decisions differ.  A CTA therefore runs the phases of KC CTUs together: the task list of a phase is the union of the
// CTUs' tasks (t-major, so the long tasks are listed first), and the decision after a phase is taken by one thread per CTU
// (32x32 / 16x16 CUs, leaf_eval) or by one warp per CTU together with the follow-up work (CUs up to 8x8, small_eval) and
// published through shared memory.
// ---------------------------------------------------------------------------------------------------------------
struct NodeId {
    int depth, a, b, c;  // position in the quad tree: child indices at depth 1, 2, 3
    int tree;            // SINGLE_TREE, DUAL_TREE_LUMA (depth 3) or DUAL_TREE_CHROMA (the 8x8 chroma CT: depth 2 geometry)
};

// CodingTree::split(SPLIT_QT) child geometry + availability flags (ctu.rs:1960-2064, 2083-2188; H9)
__device__ __forceinline__ Node qt_child(const CtuGeom g, const Node p, int i, int tree) {
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

// geometry and availability flags of a node of one CTU (the flags depend on the CTU's position in the picture)
__device__ __forceinline__ Node make_node(const CtuGeom g, const NodeId &id) {
    Node n;
    n.x = 0; n.y = 0; n.w = 32; n.tree = SINGLE_TREE;
    n.bl = false;
    n.ar = (g.cx + 32 >= g.W) ? false : (0 < g.cy && g.cx + 32 < g.W);
    if (id.depth >= 1) n = qt_child(g, n, id.a, SINGLE_TREE);
    if (id.depth >= 2) n = qt_child(g, n, id.b, SINGLE_TREE);
    if (id.tree == DUAL_TREE_CHROMA) {  // local dual tree: chroma CT of the 8x8 parent's size (ctu.rs:2031-2055)
        Node c = n;
        c.tree = DUAL_TREE_CHROMA;
        c.ar = (g.cx + c.x + c.w >= g.W) ? false : n.ar;
        c.bl = (g.cy + c.y + c.w >= g.H) ? false : n.bl;
        return c;
    }
    if (id.depth >= 3) n = qt_child(g, n, id.c, DUAL_TREE_LUMA);
    return n;
}

__device__ __forceinline__ long long luma_hdr(const Ctx S, const DevTables *tab, const Node nd, int mode, int ck) {
    int lk = luma_kind(S, S.c->g, nd, mode, S.c->root_mode);
    return nd.tree == SINGLE_TREE ? tab->hdr_single[lk][ck] : tab->hdr_dual[lk];
}

__device__ __forceinline__ void fill_lm(const Ctx S, const Node nd, int mode, int lane) {
    int cells = nd.w >> 2;
#pragma unroll RU
    for (int i = lane; i < cells * cells; i += 32) {
        int yy = i >> ilog2i(cells), xx = i & (cells - 1);
        S.c->lm[((nd.y >> 2) + yy) * 8 + (nd.x >> 2) + xx] = (uint8_t)mode;
    }
}
__device__ __forceinline__ void fill_cm(const Ctx S, const Node nd, int mode, int lane) {
    int cells = nd.w >> 3;
#pragma unroll RU
    for (int i = lane; i < cells * cells; i += 32) {
        int yy = i >> ilog2i(cells), xx = i & (cells - 1);
        S.c->cm[((nd.y >> 3) + yy) * 4 + (nd.x >> 3) + xx] = (uint8_t)mode;
    }
}

// Tasks of a phase are handed out in list order (the large luma tasks first) to whichever warp is free: a shared-memory
// ticket per phase (S.ticket[phase parity], reset two phases later); a phase with at most one task per warp is assigned
// statically.  The long tasks of a phase (the 32x32 luma pipelines of the root) lead the list and have their own ticket, which
// the warps drain before they join the general one: longest tasks first.  (Until round 2 only NBIG = 8 warps owned a scratch
// large enough for them; now every warp does.)
#ifndef WB_NT_NOINLINE
#define WB_NT_NOINLINE 0
#endif
#if WB_NT_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
int next_task(Shared &S, int &slot, int nbig, int ntot, int prev, int warp, int lane) {
    if (nbig == 0 && ntot <= NW) return prev < 0 ? warp : ntot;  // at most one task per warp: no ticket needed
    if (nbig > 0 && warp < NBIG && prev < nbig) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.ticket_big[slot], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t < nbig) return t;
    }
    int t = 0;
    if (lane == 0) t = atomicAdd(&S.ticket[slot], 1);
    return nbig + __shfl_sync(0xffffffffu, t, 0);
}
// node geometry + availability flags of the current node, cached per CTU (x | y << 6 | w << 12 | tree << 18 | ar << 20 | bl << 21)
__device__ __forceinline__ unsigned pack_node(const Node n) {
    return (unsigned)n.x | ((unsigned)n.y << 6) | ((unsigned)n.w << 12) | ((unsigned)n.tree << 18) | ((unsigned)n.ar << 20) | ((unsigned)n.bl << 21);
}
__device__ __forceinline__ Node unpack_node(unsigned p) {
    Node n;
    n.x = p & 63; n.y = (p >> 6) & 63; n.w = (p >> 12) & 63; n.tree = (p >> 18) & 3; n.ar = (p >> 20) & 1; n.bl = (p >> 21) & 1;
    return n;
}
#define WB_FOR_TASKS(ntask)                                                                            \
    for (int tt = next_task(S, S_slot, nst, (ntask) * KC, -1, warp, lane); tt < (ntask) * KC; tt = next_task(S, S_slot, nst, (ntask) * KC, tt, warp, lane)) \
        if (S.c[tt % KC].active)
// called by every thread between two phases (after the barrier that ends a phase): switch to the other ticket and
// clear the one used two phases ago
#define WB_NEXT_PHASE()                                       \
    do {                                                      \
        S_slot ^= 1;                                          \
        if (threadIdx.x == 0) { S.ticket[S_slot ^ 1] = 0; S.ticket_big[S_slot ^ 1] = 0; } \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// leaf evaluation (block_splitter.rs:886-1078) for SINGLE_TREE 32/16/8 and DUAL_TREE_LUMA 4x4 nodes; the cost of
// CTU k is left in S.c[k].leaf_cost
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void leaf_eval(Shared &S, const SearchParams &P, const NodeId id, int &S_slot) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DevTables *tab = P.tab;
    const WarpScratch ws = warp_scratch(S, warp);
    constexpr int ncomp = 3;  // SINGLE_TREE 32x32 / 16x16 CUs (smaller CUs: small_eval)
    const bool is_root = id.depth == 0;
    int nst = 0;  // number of leading long tasks (32x32 luma pipelines of the root) with their own ticket, see next_task
    constexpr int sb = 0;  // every full evaluation's outcome is kept in a candidate slot of the CTU's global scratch
    [[maybe_unused]] const int pk_ = 16 * id.depth;
    if (tid < KC && S.c[tid].active) S.c[tid].node = pack_node(make_node(S.c[tid].g, id));
    __syncthreads();
    WB_PROF(pk_ + 0);
    // ---- phase 0: reference samples
    WB_FOR_TASKS(ncomp) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        build_refs(V, V.c->g, unpack_node(V.c->node), t, lane);
    }
    __syncthreads();
    WB_PROF(pk_ + 1);
    WB_NEXT_PHASE();
    // ---- phase 1: planar / DC full evaluations, 13 coarse angular SADs
    {
        {
            const int nfull = 2 * ncomp, ntask = nfull + 13 * ncomp;
            nst = is_root ? 2 * KC : 0;
            WB_FOR_TASKS(ntask) {
                const int k = tt % KC, t = tt / KC;
                Ctx V{&S.tb, &S.c[k]};
                const Node nd = unpack_node(V.c->node);
                if (t < nfull) {
                    int mode, c;
                    if (t < 2) { mode = t; c = 0; }
                    else { mode = (t - 2) >> 1; c = 1 + ((t - 2) & 1); }
                    unsigned ssd; int rate;
                    full_task(V, tab, V.c->g, nd, c, mode, ws, lane, ssd, rate, sb + mode);
                    if (lane == 0) { V.c->pd_ssd[mode][c] = ssd; V.c->pd_rate[mode][c] = rate; }
                } else {
                    int u = t - nfull;
                    int c = u / 13, mi = u - c * 13;
                    unsigned sad = sad_task(V, V.c->g, nd, c, c_cand15[2 + mi], ws, lane);
                    if (lane == 0) V.c->r_sad[t] = sad;
                }
            }
        }
        __syncthreads();
        WB_PROF(pk_ + 2);
        WB_NEXT_PHASE();
        nst = 0;
        if (tid < KC && S.c[tid].active) {
            CtuCtx &C = S.c[tid];
            const int nfull = 2 * ncomp;
            int best = 0;
            float bc = 0.f;
            for (int i = 0; i < 13; i++) {
                unsigned s = C.r_sad[nfull + i];
                if (ncomp == 3) s += C.r_sad[nfull + 13 + i] + C.r_sad[nfull + 26 + i];
                float c = __uint2float_rn(s);
                if (i == 0 || c < bc) { bc = c; best = i; }
            }
            C.cur = c_cand15[2 + best];
            C.cur_cost = bc;
            C.v0 = !(C.cur < 2 + 2);
            C.v1 = !(C.cur + 2 > 66);
        }
        __syncthreads();
        WB_PROF(pk_ + 3);
        // ---- phases 2,3: SAD refinement +-2, +-1 (step_search aux=true, block_splitter.rs:905-973)
        for (int step = 2; step >= 1; step >>= 1) {
            WB_FOR_TASKS(2 * ncomp) {
                const int k = tt % KC, t = tt / KC;
                Ctx V{&S.tb, &S.c[k]};
                int cand = t / ncomp, c = t - cand * ncomp;
                if (cand == 0 ? V.c->v0 : V.c->v1) {
                    unsigned sad = sad_task(V, V.c->g, unpack_node(V.c->node), c, cand == 0 ? V.c->cur - step : V.c->cur + step, ws, lane);
                    if (lane == 0) V.c->r_sad[t] = sad;
                }
            }
            __syncthreads();
            WB_PROF(pk_ + 4);
            WB_NEXT_PHASE();
            if (tid < KC && S.c[tid].active) {
                CtuCtx &C = S.c[tid];
                float c0 = FLT_MAX, c1 = FLT_MAX;
                if (C.v0) { unsigned s = 0; for (int c = 0; c < ncomp; c++) s += C.r_sad[c]; c0 = __uint2float_rn(s); }
                if (C.v1) { unsigned s = 0; for (int c = 0; c < ncomp; c++) s += C.r_sad[ncomp + c]; c1 = __uint2float_rn(s); }
                float mn = fminf(fminf(C.cur_cost, c0), c1);
                if (C.cur_cost == mn) {
                } else if (c0 == mn) { C.cur -= step; C.cur_cost = c0; }
                else { C.cur += step; C.cur_cost = c1; }
                const int ns = step >> 1;
                if (ns > 0) { C.v0 = !(C.cur < 2 + ns); C.v1 = !(C.cur + ns > 66); }
                else { C.dir = C.cur; C.v0 = !(C.dir < 3); C.v1 = !(C.dir + 1 > 66); }
            }
            __syncthreads();
            WB_PROF(pk_ + 5);
        }
    }
    // ---- phase 4: full evaluation of dir, dir-1, dir+1 (step_search aux=false)
    nst = is_root ? 3 * KC : 0;
    WB_FOR_TASKS(9) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        int cand, c;
        if (t < 3) { cand = t; c = 0; }
        else { cand = (t - 3) >> 1; c = 1 + ((t - 3) & 1); }
        bool valid = cand == 0 || (cand == 1 ? V.c->v0 : V.c->v1);
        if (valid) {
            const int dir = V.c->dir;
            int mode = cand == 0 ? dir : (cand == 1 ? dir - 1 : dir + 1);
            unsigned ssd; int rate;
            full_task(V, tab, V.c->g, unpack_node(V.c->node), c, mode, ws, lane, ssd, rate, sb + 2 + cand);
            if (lane == 0) { V.c->r_ssd[t] = ssd; V.c->r_rate[t] = rate; }
        }
    }
    __syncthreads();
    WB_PROF(pk_ + 6);
    WB_NEXT_PHASE();
    if (tid < KC && S.c[tid].active) {
        Ctx V{&S.tb, &S.c[tid]};
        CtuCtx &C = *V.c;
        const Node nd = unpack_node(C.node);
        float cc[3];
        for (int cand = 0; cand < 3; cand++) {
            bool valid = cand == 0 || (cand == 1 ? C.v0 : C.v1);
            cc[cand] = FLT_MAX;
            if (valid) {
                unsigned ssd = C.r_ssd[cand];
                long long r = C.r_rate[cand];
                if (ncomp == 3) { ssd += C.r_ssd[3 + 2 * cand] + C.r_ssd[4 + 2 * cand]; r += (long long)C.r_rate[3 + 2 * cand] + C.r_rate[4 + 2 * cand]; }
                int mode = cand == 0 ? C.dir : (cand == 1 ? C.dir - 1 : C.dir + 1);
                cc[cand] = rd_cost(ssd, r + luma_hdr(V, tab, nd, mode, 0), tab->lambda_rd);
            }
        }
        {   // planar and DC (evaluated in phase 1)
            unsigned ssd0 = C.pd_ssd[0][0], ssd1 = C.pd_ssd[1][0];
            long long r0 = C.pd_rate[0][0], r1 = C.pd_rate[1][0];
            if (ncomp == 3) {
                ssd0 += C.pd_ssd[0][1] + C.pd_ssd[0][2]; r0 += (long long)C.pd_rate[0][1] + C.pd_rate[0][2];
                ssd1 += C.pd_ssd[1][1] + C.pd_ssd[1][2]; r1 += (long long)C.pd_rate[1][1] + C.pd_rate[1][2];
            }
            C.cost_pl = rd_cost(ssd0, r0 + luma_hdr(V, tab, nd, 0, 0), tab->lambda_rd);
            C.cost_dc = rd_cost(ssd1, r1 + luma_hdr(V, tab, nd, 1, 0), tab->lambda_rd);
        }
        float mn = fminf(fminf(cc[0], cc[1]), cc[2]);
        if (cc[0] == mn) { C.dir_cost = cc[0]; C.dir_cand = 0; }
        else if (cc[1] == mn) { C.dir -= 1; C.dir_cost = cc[1]; C.dir_cand = 1; }
        else { C.dir += 1; C.dir_cost = cc[2]; C.dir_cand = 2; }
        // ---- winner among planar, DC, dir (first minimum)
        C.min_cost = fminf(fminf(C.cost_pl, C.cost_dc), C.dir_cost);
        C.mode = C.cost_pl == C.min_cost ? 0 : (C.cost_dc == C.min_cost ? 1 : C.dir);
        if (is_root) C.root_mode = C.mode;
        C.leaf_cost = C.min_cost;
    }
    __syncthreads();
    WB_PROF(pk_ + 7);
    // ---- phase 5: luma redo (commit) + chroma DM full evaluation (commit)
    nst = 0;
    WB_FOR_TASKS(ncomp) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = unpack_node(V.c->node);
        {   // the winner's luma and its same-mode (DM) chroma were evaluated in phase 1 or 4: copy them out of their slot
            const int md = V.c->mode, dc = V.c->dir_cand;
            if (is_root) commit_root_slot(V, t, md <= 1 ? md : 2 + dc, lane);
            else commit_slot(V, nd, t, md <= 1 ? md : 2 + dc, lane);
            if (lane == 0) {
                if (md <= 1) { V.c->fin_ssd[t] = V.c->pd_ssd[md][t]; V.c->fin_rate[t] = V.c->pd_rate[md][t]; }
                else { const int r = t == 0 ? dc : 3 + 2 * dc + (t - 1); V.c->fin_ssd[t] = V.c->r_ssd[r]; V.c->fin_rate[t] = V.c->r_rate[r]; }
            }
        }
        if (t == 0) fill_lm(V, nd, V.c->mode, lane);
    }
    __syncthreads();
    WB_PROF(pk_ + 8);
    WB_NEXT_PHASE();
    // ---- phases 5b + 6: CCLM down-sampled luma of the committed luma reconstruction, then the CCLM mode by SAD (order LT,
    //      T, L), both by one warp per CTU
    nst = 0;
    WB_FOR_TASKS(1) {
        const int k = tt % KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = unpack_node(V.c->node);
        cclm_downsample(V, V.c->g, nd, lane);
        __syncwarp();
        const int cm = cclm_search(V, V.c->g, nd, lane);
        if (lane == 0) V.c->cclm_mode = cm;
    }
    __syncthreads();
    WB_PROF(pk_ + 10);
    WB_NEXT_PHASE();
    // ---- phase 7: CCLM full evaluation (no commit)
        WB_FOR_TASKS(2) {
            const int k = tt % KC, t = tt / KC;
            Ctx V{&S.tb, &S.c[k]};
            unsigned ssd; int rate;
            full_task(V, tab, V.c->g, unpack_node(V.c->node), 1 + t, V.c->cclm_mode, ws, lane, ssd, rate, sb + 5);
            if (lane == 0) { V.c->r_ssd[8 + t] = ssd; V.c->r_rate[8 + t] = rate; }
        }
    __syncthreads();
    WB_PROF(pk_ + 12);
    WB_NEXT_PHASE();
    if (tid < KC && S.c[tid].active) {
        Ctx V{&S.tb, &S.c[tid]};
        CtuCtx &C = *V.c;
        const Node nd = unpack_node(C.node);
        const unsigned ssdY = C.fin_ssd[0], ssdDM = C.fin_ssd[1] + C.fin_ssd[2];
        const long long rateY = C.fin_rate[0], rateDM = (long long)C.fin_rate[1] + C.fin_rate[2];
        const float cost_dm = rd_cost(ssdDM, rateDM + tab->hdr_chroma[0], tab->lambda_rd_c);
        const unsigned ssdCC = C.r_ssd[8] + C.r_ssd[9];
        const long long rateCC = (long long)C.r_rate[8] + C.r_rate[9];
        const int ck = 1 + (C.cclm_mode - MODE_LT_CCLM);
        const float cost_cclm = rd_cost(ssdCC, rateCC + tab->hdr_chroma[ck], tab->lambda_rd_c);
        const float cmn = fminf(cost_dm, cost_cclm);
        if (cost_dm == cmn) {
            C.cclm_wins = 0;
            C.leaf_cost = rd_cost(ssdY + ssdDM, rateY + rateDM + luma_hdr(V, tab, nd, C.mode, 0), tab->lambda_rd);
        } else {
            C.cclm_wins = 1;
            C.leaf_cost = rd_cost(ssdY + ssdCC, rateY + rateCC + luma_hdr(V, tab, nd, C.mode, ck), tab->lambda_rd);
        }
    }
    __syncthreads();
    WB_PROF(pk_ + 13);
    // ---- phase 8: commit the CCLM chroma where it won; publish the chroma mode
        WB_FOR_TASKS(2) {
            const int k = tt % KC, t = tt / KC;
            Ctx V{&S.tb, &S.c[k]};
            const Node nd = unpack_node(V.c->node);
            if (V.c->cclm_wins) {  // evaluated in phase 7 with unchanged inputs
                if (is_root) commit_root_slot(V, 1 + t, 5, lane);
                else commit_slot(V, nd, 1 + t, 5, lane);
            }
            if (t == 0) fill_cm(V, nd, V.c->cclm_wins ? V.c->cclm_mode : V.c->mode, lane);
        }
    __syncthreads();
    WB_PROF(pk_ + 14);
    WB_NEXT_PHASE();
}

// ---------------------------------------------------------------------------------------------------------------
// CUs up to 8x8: the 8x8 SINGLE_TREE CU, the 4x4 DUAL_TREE_LUMA CU and the 8x8 DUAL_TREE_CHROMA coding tree.  These are the
// most numerous nodes and their blocks are too small to fill a phase with work, so the evaluation is cut into few, fat
// phases: whatever depends only on one CTU's previous result (decision -> commit -> CCLM down-sampling -> CCLM mode search)
// runs back to back in ONE warp per CTU instead of being separated by block-wide barriers.
//   A  node geometry + reference samples                                    (one warp per component)
//   B  coarse direction search in parts + refinement by the last arriver || planar / DC full evaluations
//   C  dir, dir-1, dir+1 full evaluations
//   D  one warp per CTU: RD decision among planar, DC, dir (costs of the 5 candidates on 5 lanes), commit of the winner from
//      its slot, mode map; 4x4 CU: cost to the parent, done.  8x8 CU: CCLM down-sampled luma, CCLM mode by SAD
//   E  8x8 CU: CCLM full evaluation (Cb | Cr by the two half-warps)
//   F  8x8 CU, one warp per CTU: DM vs CCLM, commit, chroma mode
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void small_eval(Shared &S, const SearchParams &P, const NodeId id, int &S_slot) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DevTables *tab = P.tab;
    const WarpScratch ws = warp_scratch(S, warp);
    const bool cu4 = id.depth == 3;  // 4x4 luma CU (else the 8x8 SINGLE_TREE CU)
    const int ncomp = cu4 ? 1 : 3;
    const int nparts = cu4 ? 1 : 4;  // 4x4 luma CU: eight modes per pass, the whole direction search is one warp task
    constexpr int nst = 0;
    [[maybe_unused]] const int pk_ = 16 * id.depth;
    // ---- A
    WB_FOR_TASKS(ncomp) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = make_node(V.c->g, id);
        if (t == 0 && lane == 0) V.c->node = pack_node(nd);
        if (t == 0 && !cu4) build_refs(V, V.c->g, nd, 0, lane);
        else build_refs4(V, V.c->g, nd, t, lane);
    }
    __syncthreads();
    WB_PROF(pk_ + 1);
    WB_NEXT_PHASE();
    // ---- B
    WB_FOR_TASKS(cu4 ? nparts + 1 : nparts + 4) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = unpack_node(V.c->node);
        if (!cu4 && t < 2) {  // 8x8 luma planar, DC
            unsigned ssd; int rate;
            full_task(V, tab, V.c->g, nd, 0, t, ws, lane, ssd, rate, t);
            if (lane == 0) { V.c->pd_ssd[t][0] = ssd; V.c->pd_rate[t][0] = rate; }
        } else if (t < (cu4 ? 0 : 2) + nparts) {
            dir_search_part(V, nd, nparts - 1 - (t - (cu4 ? 0 : 2)), nparts, reinterpret_cast<uint8_t *>(ws.refx), lane);  // the last part is the longest: it goes first
        } else {  // 4x4 TBs: planar | DC of the luma CU, or Cb | Cr of one mode
            const int half = lane >> 4;
            const int mode = cu4 ? half : t - (2 + nparts), c = cu4 ? 0 : 1 + half;
            unsigned ssd; int rate;
            full_pair4(V, tab, V.c->g, nd, c, mode, false, mode, ws, lane, ssd, rate);
            if ((lane & 15) == 0) { V.c->pd_ssd[mode][c] = ssd; V.c->pd_rate[mode][c] = rate; }
        }
    }
    __syncthreads();
    WB_PROF(pk_ + 2);
    WB_NEXT_PHASE();
    // ---- C
    WB_FOR_TASKS(cu4 ? 2 : 6) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = unpack_node(V.c->node);
        const int dir = V.c->dir;
        if (!cu4 && t < 3) {
            const bool valid = t == 0 || (t == 1 ? V.c->v0 : V.c->v1);
            if (valid) {
                unsigned ssd; int rate;
                full_task(V, tab, V.c->g, nd, 0, t == 0 ? dir : (t == 1 ? dir - 1 : dir + 1), ws, lane, ssd, rate, 2 + t);
                if (lane == 0) { V.c->r_ssd[t] = ssd; V.c->r_rate[t] = rate; }
            }
        } else {
            const int half = lane >> 4;
            const int cand = cu4 ? 2 * t + half : t - 3, c = cu4 ? 0 : 1 + half;
            const bool valid = cand == 0 || (cand == 1 ? V.c->v0 : (cand == 2 && V.c->v1));
            // both halves of the warp run full_pair4 in step: the 4x4 CU's second half may have no candidate (dir-1 invalid, or the
            // empty partner of dir+1) and then runs along on the first half's mode without storing anything
            const int cand0 = cu4 ? 2 * t : cand;
            const bool valid0 = cand0 == 0 || (cand0 == 1 ? V.c->v0 : V.c->v1);
            if (valid0) {
                const int cm = valid ? cand : cand0;
                unsigned ssd; int rate;
                full_pair4(V, tab, V.c->g, nd, c, cm == 0 ? dir : (cm == 1 ? dir - 1 : dir + 1), false, 2 + cm, ws, lane, ssd, rate, valid);
                const int ri = c == 0 ? cand : 3 + 2 * cand + (c - 1);
                if (valid && (lane & 15) == 0) { V.c->r_ssd[ri] = ssd; V.c->r_rate[ri] = rate; }
            }
        }
    }
    __syncthreads();
    WB_PROF(pk_ + 6);
    WB_NEXT_PHASE();
    // ---- D
    WB_FOR_TASKS(1) {
        const int k = tt % KC;
        Ctx V{&S.tb, &S.c[k]};
        CtuCtx &C = *V.c;
        const Node nd = unpack_node(C.node);
        int dir = C.dir;
        // 4x4 CU: fetch the five candidate slots (16 samples each, two slots per warp load) while the candidates are priced,
        // so that the commit below does not wait for an L2 round trip after the decision
        int pr0 = 0, pr1 = 0, pr2 = 0, pl0 = 0, pl1 = 0, pl2 = 0;
        if (cu4) {
            const int h = lane >> 4, i = lane & 15;
            const uint8_t *gr = C.groot + h * ROOT_SLOT_SAMPLES + i;
            const int16_t *gv = reinterpret_cast<const int16_t *>(C.groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES) + h * ROOT_SLOT_SAMPLES + i;
            pr0 = __ldcg(gr); pl0 = __ldcg(gv);
            pr1 = __ldcg(gr + 2 * ROOT_SLOT_SAMPLES); pl1 = __ldcg(gv + 2 * ROOT_SLOT_SAMPLES);
            if (h == 0) { pr2 = __ldcg(gr + 4 * ROOT_SLOT_SAMPLES); pl2 = __ldcg(gv + 4 * ROOT_SLOT_SAMPLES); }
        }
        // lane j < 5 prices candidate j: planar, DC, dir, dir-1, dir+1 (block_splitter.rs:472-473)
        float cost = FLT_MAX;
        if (lane < 5) {
            const int cand = lane - 2;
            const bool valid = lane < 3 || (cand == 1 ? C.v0 : C.v1);
            if (valid) {
                unsigned ssd;
                long long r;
                if (lane < 2) {
                    ssd = C.pd_ssd[lane][0]; r = C.pd_rate[lane][0];
                    if (ncomp == 3) { ssd += C.pd_ssd[lane][1] + C.pd_ssd[lane][2]; r += (long long)C.pd_rate[lane][1] + C.pd_rate[lane][2]; }
                } else {
                    ssd = C.r_ssd[cand]; r = C.r_rate[cand];
                    if (ncomp == 3) { ssd += C.r_ssd[3 + 2 * cand] + C.r_ssd[4 + 2 * cand]; r += (long long)C.r_rate[3 + 2 * cand] + C.r_rate[4 + 2 * cand]; }
                }
                const int mode = lane < 2 ? lane : (cand == 0 ? dir : (cand == 1 ? dir - 1 : dir + 1));
                cost = rd_cost(ssd, r + luma_hdr(V, tab, nd, mode, 0), tab->lambda_rd);
            }
        }
        const float cost_pl = __shfl_sync(0xffffffffu, cost, 0), cost_dc = __shfl_sync(0xffffffffu, cost, 1);
        const float cc0 = __shfl_sync(0xffffffffu, cost, 2), cc1 = __shfl_sync(0xffffffffu, cost, 3), cc2 = __shfl_sync(0xffffffffu, cost, 4);
        const float mn = fminf(fminf(cc0, cc1), cc2);
        float dir_cost;
        int dir_cand;
        if (cc0 == mn) { dir_cost = cc0; dir_cand = 0; }
        else if (cc1 == mn) { dir -= 1; dir_cost = cc1; dir_cand = 1; }
        else { dir += 1; dir_cost = cc2; dir_cand = 2; }
        // winner among planar, DC, dir (first minimum)
        const float min_cost = fminf(fminf(cost_pl, cost_dc), dir_cost);
        const int mode = cost_pl == min_cost ? 0 : (cost_dc == min_cost ? 1 : dir);
        if (lane == 0) {
            C.mode = mode;
            C.leaf_cost = min_cost;
            if (cu4) C.split[2] = __fadd_rn(C.split[2], min_cost);  // the 4x4 CU is final: its cost goes to the 8x8 parent's split cost
        }
        // the winner's luma and same-mode (DM) chroma were evaluated in phase B or C: copy them out of the slot
        const int slot = mode <= 1 ? mode : 2 + dir_cand;
        if (cu4) {
            const int rs = slot >> 1, i = lane & 15;
            const int r = rs == 0 ? pr0 : (rs == 1 ? pr1 : pr2), l = rs == 0 ? pl0 : (rs == 1 ? pl1 : pl2);
            if ((lane >> 4) == (slot & 1)) {
                RY(V, nd.x + (i & 3), nd.y + (i >> 2)) = (uint8_t)r;
                C.lvY[(nd.y + (i >> 2)) * 32 + nd.x + (i & 3)] = (int16_t)l;
            }
            __syncwarp();
        } else {
            for (int t = 0; t < ncomp; t++) commit_slot(V, nd, t, slot, lane);
        }
        if (!cu4 && lane < 3) {
            const int t = lane;
            if (mode <= 1) { C.fin_ssd[t] = C.pd_ssd[mode][t]; C.fin_rate[t] = C.pd_rate[mode][t]; }
            else { const int r = t == 0 ? dir_cand : 3 + 2 * dir_cand + (t - 1); C.fin_ssd[t] = C.r_ssd[r]; C.fin_rate[t] = C.r_rate[r]; }
        }
        fill_lm(V, nd, mode, lane);
        if (!cu4) {
            __syncwarp();
            cclm_downsample(V, C.g, nd, lane);
            __syncwarp();
            const int cm = cclm_search(V, C.g, nd, lane);
            if (lane == 0) C.cclm_mode = cm;
        }
    }
    __syncthreads();
    WB_PROF(pk_ + 7);
    WB_NEXT_PHASE();
    if (cu4) return;
    // ---- E
    WB_FOR_TASKS(1) {
        const int k = tt % KC, h = lane >> 4;
        Ctx V{&S.tb, &S.c[k]};
        unsigned ssd; int rate;
        full_pair4(V, tab, V.c->g, unpack_node(V.c->node), 1 + h, V.c->cclm_mode, false, 5, ws, lane, ssd, rate);
        if ((lane & 15) == 0) { V.c->r_ssd[8 + h] = ssd; V.c->r_rate[8 + h] = rate; }
    }
    __syncthreads();
    WB_PROF(pk_ + 12);
    WB_NEXT_PHASE();
    // ---- F
    WB_FOR_TASKS(1) {
        const int k = tt % KC;
        Ctx V{&S.tb, &S.c[k]};
        CtuCtx &C = *V.c;
        const Node nd = unpack_node(C.node);
        // fetch the CCLM evaluation's slot (Cb | Cr, 16 samples each) while the costs are computed: the commit below then does not
        // wait for an L2 round trip after the decision
        const int pc_ = 1 + (lane >> 4), pi_ = lane & 15;
        const int pso_ = 5 * ROOT_SLOT_SAMPLES + (pc_ == 1 ? 1024 : 1280) + pi_;
        const int prec_ = __ldcg(C.groot + pso_);
        const int plev_ = __ldcg(reinterpret_cast<const int16_t *>(C.groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES) + pso_);
        const unsigned ssdY = C.fin_ssd[0], ssdDM = C.fin_ssd[1] + C.fin_ssd[2];
        const long long rateY = C.fin_rate[0], rateDM = (long long)C.fin_rate[1] + C.fin_rate[2];
        const float cost_dm = rd_cost(ssdDM, rateDM + tab->hdr_chroma[0], tab->lambda_rd_c);
        const unsigned ssdCC = C.r_ssd[8] + C.r_ssd[9];
        const long long rateCC = (long long)C.r_rate[8] + C.r_rate[9];
        const int cclm_mode = C.cclm_mode, mode = C.mode;
        const int ck = 1 + (cclm_mode - MODE_LT_CCLM);
        const float cost_cclm = rd_cost(ssdCC, rateCC + tab->hdr_chroma[ck], tab->lambda_rd_c);
        const bool cclm_wins = !(cost_dm == fminf(cost_dm, cost_cclm));  // tie -> DM
        __syncwarp();
        if (lane == 0) {
            C.cclm_wins = cclm_wins;
            C.leaf_cost = cclm_wins ? rd_cost(ssdY + ssdCC, rateY + rateCC + luma_hdr(V, tab, nd, mode, ck), tab->lambda_rd)
                                    : rd_cost(ssdY + ssdDM, rateY + rateDM + luma_hdr(V, tab, nd, mode, 0), tab->lambda_rd);
        }
        if (cclm_wins) {  // evaluated in phase E with unchanged inputs: copy it out of its slot
            const int bx_ = nd.x >> 1, by_ = nd.y >> 1;
            RC(V, pc_, bx_ + (pi_ & 3), by_ + (pi_ >> 2)) = (uint8_t)prec_;
            C.lvC[pc_ - 1][(by_ + (pi_ >> 2)) * 16 + bx_ + (pi_ & 3)] = (int16_t)plev_;
            __syncwarp();
        }
        fill_cm(V, nd, cclm_wins ? cclm_mode : mode, lane);
    }
    __syncthreads();
    WB_PROF(pk_ + 14);
    WB_NEXT_PHASE();
}

// ---------------------------------------------------------------------------------------------------------------
// the 8x8 DUAL_TREE_CHROMA coding tree that follows four 4x4 luma CUs (block_splitter.rs:794-885): DM (the mode of the
// luma CU covering the centre, ctu.rs:2372-2396 = the bottom-right 4x4) against the best-SAD CCLM mode
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void chroma_ct_eval(Shared &S, const SearchParams &P, const NodeId id, int &S_slot) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int nst = 0;
    const DevTables *tab = P.tab;
    const WarpScratch ws = warp_scratch(S, warp);
    // ---- A: reference samples of Cb, Cr; CCLM down-sampled luma
    WB_FOR_TASKS(3) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = make_node(V.c->g, id);
        if (t == 0 && lane == 0) V.c->node = pack_node(nd);
        if (t < 2) build_refs4(V, V.c->g, nd, 1 + t, lane);
        else cclm_downsample(V, V.c->g, nd, lane);
    }
    __syncthreads();
    WB_PROF(64 + 1);
    WB_NEXT_PHASE();
    // ---- B: DM full evaluation (committed) || CCLM mode by SAD
    WB_FOR_TASKS(2) {
        const int k = tt % KC, t = tt / KC;
        Ctx V{&S.tb, &S.c[k]};
        const Node nd = unpack_node(V.c->node);
        if (t == 0) {
            const int dm = V.c->lm[((nd.y >> 2) + 1) * 8 + (nd.x >> 2) + 1], h = lane >> 4;
            unsigned ssd; int rate;
            full_pair4(V, tab, V.c->g, nd, 1 + h, dm, true, -1, ws, lane, ssd, rate);
            if ((lane & 15) == 0) { V.c->r_ssd[h] = ssd; V.c->r_rate[h] = rate; }
        } else {
            const int cm = cclm_search(V, V.c->g, nd, lane);
            if (lane == 0) V.c->cclm_mode = cm;
        }
    }
    __syncthreads();
    WB_PROF(64 + 2);
    WB_NEXT_PHASE();
    // ---- C: CCLM full evaluation (not committed)
    WB_FOR_TASKS(1) {
        const int k = tt % KC, h = lane >> 4;
        Ctx V{&S.tb, &S.c[k]};
        unsigned ssd; int rate;
        full_pair4(V, tab, V.c->g, unpack_node(V.c->node), 1 + h, V.c->cclm_mode, false, 5, ws, lane, ssd, rate);
        if ((lane & 15) == 0) { V.c->r_ssd[8 + h] = ssd; V.c->r_rate[8 + h] = rate; }
    }
    __syncthreads();
    WB_PROF(64 + 4);
    WB_NEXT_PHASE();
    // ---- D: one warp per CTU: decision (tie -> DM), commit of the CCLM chroma if it won, chroma mode, cost to the 8x8 parent
    WB_FOR_TASKS(1) {
        const int k = tt % KC;
        Ctx V{&S.tb, &S.c[k]};
        CtuCtx &C = *V.c;
        const Node nd = unpack_node(C.node);
        // fetch the CCLM evaluation's slot (Cb | Cr, 16 samples each) while the costs are computed: the commit below then does not
        // wait for an L2 round trip after the decision
        const int pc_ = 1 + (lane >> 4), pi_ = lane & 15;
        const int pso_ = 5 * ROOT_SLOT_SAMPLES + (pc_ == 1 ? 1024 : 1280) + pi_;
        const int prec_ = __ldcg(C.groot + pso_);
        const int plev_ = __ldcg(reinterpret_cast<const int16_t *>(C.groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES) + pso_);
        const int cclm_mode = C.cclm_mode;
        const float cost_dm = rd_cost(C.r_ssd[0] + C.r_ssd[1], (long long)C.r_rate[0] + C.r_rate[1] + tab->hdr_chroma[0], tab->lambda_rd_c);
        const int ck = 1 + (cclm_mode - MODE_LT_CCLM);
        const float cost_cclm = rd_cost(C.r_ssd[8] + C.r_ssd[9], (long long)C.r_rate[8] + C.r_rate[9] + tab->hdr_chroma[ck], tab->lambda_rd_c);
        const float mn = fminf(cost_dm, cost_cclm);
        const bool cclm_wins = !(cost_dm == mn);
        const int dm = C.lm[((nd.y >> 2) + 1) * 8 + (nd.x >> 2) + 1];
        __syncwarp();
        if (lane == 0) {
            C.cclm_wins = cclm_wins;
            C.leaf_cost = mn;
            C.split[2] = __fadd_rn(C.split[2], mn);
        }
        if (cclm_wins) {  // evaluated in phase C with unchanged inputs (CCLM reads the luma and the neighbours outside the block)
            const int bx_ = nd.x >> 1, by_ = nd.y >> 1;
            RC(V, pc_, bx_ + (pi_ & 3), by_ + (pi_ >> 2)) = (uint8_t)prec_;
            C.lvC[pc_ - 1][(by_ + (pi_ >> 2)) * 16 + bx_ + (pi_ & 3)] = (int16_t)plev_;
            __syncwarp();
        }
        fill_cm(V, nd, cclm_wins ? cclm_mode : dm, lane);
    }
    __syncthreads();
    WB_PROF(64 + 6);
    WB_NEXT_PHASE();
}

// ---------------------------------------------------------------------------------------------------------------
// no-split state save / restore (block_splitter.rs:1085-1109, 1125-1145) — here also levels and modes, because the
// CUDA path keeps the tree as flat arrays instead of cloning CodingTree objects
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sv_off_y(int d) { return d == 0 ? 0 : (d == 1 ? 1024 : 1280); }
__device__ __forceinline__ int sv_off_c(int d) { return d == 0 ? 0 : (d == 1 ? 256 : 320); }

// The saved samples and levels live in the CTU's global scratch (gsave: recY | recCb | recCr | lvY | lvCb | lvCr): written
// once per node that may split, read back only where the no-split state wins.
__device__ __noinline__ void save_node(const Ctx S, const Node nd, int d, int tid, int nthr) {
    WB_SHARED_CTX(S);
    const int w = nd.w, oy = sv_off_y(d), oc = sv_off_c(d);
    uint8_t *svRecY = S.c->gsave;
    int16_t *svLvY = reinterpret_cast<int16_t *>(S.c->gsave + SAVE_SAMPLES);
#pragma unroll RU
    for (int i = tid; i < w * w; i += nthr) {
        int y = i >> ilog2i(w), x = i & (w - 1);
        svRecY[oy + i] = RY(S, nd.x + x, nd.y + y);
        svLvY[oy + i] = S.c->lvY[(nd.y + y) * 32 + nd.x + x];
    }
    const int cw = w >> 1, bx = nd.x >> 1, by = nd.y >> 1;
#pragma unroll RU
    for (int i = tid; i < 2 * cw * cw; i += nthr) {
        int c = i >= cw * cw, j = i & (cw * cw - 1);
        int y = j >> ilog2i(cw), x = j & (cw - 1);
        svRecY[SAVE_Y + c * SAVE_C + oc + j] = RC(S, 1 + c, bx + x, by + y);
        svLvY[SAVE_Y + c * SAVE_C + oc + j] = S.c->lvC[c][(by + y) * 16 + bx + x];
    }
#pragma unroll RU
    for (int i = tid; i < 64; i += nthr) S.c->svLm[d][i] = S.c->lm[i];
#pragma unroll RU
    for (int i = tid; i < 16; i += nthr) S.c->svCm[d][i] = S.c->cm[i];
}
__device__ __noinline__ void restore_node(const Ctx S, const Node nd, int d, int tid, int nthr) {
    WB_SHARED_CTX(S);
    const int w = nd.w, oy = sv_off_y(d), oc = sv_off_c(d);
    const uint8_t *svRecY = S.c->gsave;
    const int16_t *svLvY = reinterpret_cast<const int16_t *>(S.c->gsave + SAVE_SAMPLES);
#pragma unroll RU
    for (int i = tid; i < w * w; i += nthr) {
        int y = i >> ilog2i(w), x = i & (w - 1);
        RY(S, nd.x + x, nd.y + y) = __ldcg(svRecY + oy + i);
        S.c->lvY[(nd.y + y) * 32 + nd.x + x] = __ldcg(svLvY + oy + i);
    }
    const int cw = w >> 1, bx = nd.x >> 1, by = nd.y >> 1;
#pragma unroll RU
    for (int i = tid; i < 2 * cw * cw; i += nthr) {
        int c = i >= cw * cw, j = i & (cw * cw - 1);
        int y = j >> ilog2i(cw), x = j & (cw - 1);
        RC(S, 1 + c, bx + x, by + y) = __ldcg(svRecY + SAVE_Y + c * SAVE_C + oc + j);
        S.c->lvC[c][(by + y) * 16 + bx + x] = __ldcg(svLvY + SAVE_Y + c * SAVE_C + oc + j);
    }
    const int cells = w >> 2;
#pragma unroll RU
    for (int i = tid; i < cells * cells; i += nthr) {
        int yy = i >> ilog2i(cells), xx = i & (cells - 1);
        int idx = ((nd.y >> 2) + yy) * 8 + (nd.x >> 2) + xx;
        S.c->lm[idx] = S.c->svLm[d][idx];
    }
    const int cc = w >> 3;
#pragma unroll RU
    for (int i = tid; i < cc * cc; i += nthr) {
        int yy = i >> ilog2i(cc), xx = i & (cc - 1);
        int idx = ((nd.y >> 3) + yy) * 4 + (nd.x >> 3) + xx;
        S.c->cm[idx] = S.c->svCm[d][idx];
    }
}

// after the leaf evaluation of a node that may still split: remember its cost and state.  The CTUs of the batch are handled
// side by side, NTHREADS / KC threads each.
constexpr int GT = NTHREADS / KC;  // threads per CTU in the save / restore passes
__device__ void begin_split(Shared &S, const NodeId &id, int d) {
    const int k = threadIdx.x / GT, gt = threadIdx.x % GT;
    if (k < KC && S.c[k].active) {
        Ctx V{&S.tb, &S.c[k]};
        CtuCtx &C = *V.c;
        save_node(V, make_node(C.g, id), d, gt, GT);
        if (gt == 0) {
            C.cost[d] = C.leaf_cost;
            C.split[d] = 0.0f;
            if (d < 2) C.mask_sv[d] = C.mask;
        }
    }
    __syncthreads();
}
// split_cost > no_split_cost keeps the no-split state (block_splitter.rs:1125); otherwise the split is adopted.  The node's
// final cost is added to its parent's running split cost (f32, child order) when there is a parent (dp >= 0).
__device__ void end_split(Shared &S, const NodeId &id, int d, int bit, int dp) {
    const int k = threadIdx.x / GT, gt = threadIdx.x % GT;
    if (k < KC && S.c[k].active) {
        Ctx V{&S.tb, &S.c[k]};
        CtuCtx &C = *V.c;
        const float split = C.split[d], nosplit = C.cost[d];
        const bool restore = split > nosplit;
        if (restore) restore_node(V, make_node(C.g, id), d, gt, GT);
        if (gt == 0) {
            if (restore) {
                if (d < 2) C.mask = C.mask_sv[d];
            } else {
                C.mask |= 1u << bit;
            }
            const float cost = restore ? nosplit : split;
            C.leaf_cost = cost;
            if (dp >= 0) C.split[dp] = __fadd_rn(C.split[dp], cost);
        }
    }
    __syncthreads();
}
// add the cost of the node just finished (leaf_cost) to its parent's running split cost (f32, child order)
__device__ __forceinline__ void add_to_parent(Shared &S, int d) {
    const int tid = threadIdx.x;
    if (tid < KC && S.c[tid].active) S.c[tid].split[d] = __fadd_rn(S.c[tid].split[d], S.c[tid].leaf_cost);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// split_ct(root, max_depth) of KC CTUs  (block_splitter.rs:782-1154)
// ---------------------------------------------------------------------------------------------------------------
__device__ void ctu_search(Shared &S, const SearchParams &P, int &S_slot) {
    const int md = P.max_depth;
    NodeId n32{0, 0, 0, 0, SINGLE_TREE};
    leaf_eval(S, P, n32, S_slot);
    if (md >= 1) {
        begin_split(S, n32, 0);
        for (int a = 0; a < 4; a++) {
            NodeId n16{1, a, 0, 0, SINGLE_TREE};
            leaf_eval(S, P, n16, S_slot);
            if (md >= 2) {
                begin_split(S, n16, 1);
                for (int b = 0; b < 4; b++) {
                    NodeId n8{2, a, b, 0, SINGLE_TREE};
                    small_eval(S, P, n8, S_slot);
                    if (md >= 3) {
                        begin_split(S, n8, 2);
                        for (int c = 0; c < 4; c++) {  // the 4x4 CUs and the chroma CT add their cost to split[2] themselves
                            NodeId n4{3, a, b, c, DUAL_TREE_LUMA};
                            small_eval(S, P, n4, S_slot);
                        }
                        NodeId nc{2, a, b, 0, DUAL_TREE_CHROMA};
                        chroma_ct_eval(S, P, nc, S_slot);
                        end_split(S, n8, 2, 5 + a * 4 + b, 1);
                    } else add_to_parent(S, 1);
                }
                end_split(S, n16, 1, 1 + a, 0);
            } else add_to_parent(S, 0);
        }
        end_split(S, n32, 0, 0, -1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ void init_tables(Shared &S, const DevTables *tab, int tid) {
    auto dct = [](int i, int x, int n) -> int {  // DCT-II matrix entry T[i][x] of size n (transformer.rs:934-1234)
        if (i == 0) return 64;
        int m = (i * (2 * x + 1) * (32 / n)) % 128;  // angle in units of pi/64
        int sg = 1;
        if (m > 64) m = 128 - m;
        if (m > 32) { m = 64 - m; sg = -1; }
        return sg * c_cos32[m];
    };
    for (int l2 = 2; l2 <= 5; l2++) {
        const int n = 1 << l2, o = tab_off(l2);
        for (int e = tid; e < n * n; e += NTHREADS) {
            const int i = e >> l2, x = e & (n - 1);
            const int v = dct(i, x, n);
            S.tb.T[o + i * n + x] = (int8_t)v;
            S.tb.Tt[o + x * n + i] = (int8_t)v;
        }
        for (int e = tid; e < n * n / 4; e += NTHREADS) {
            const int q4 = e >> l2, j = e & (n - 1);
            unsigned qr = 0, qc = 0;
            for (int b = 0; b < 4; b++) {
                qr |= (unsigned)(uint8_t)(int8_t)dct(j, 4 * q4 + b, n) << (8 * b);
                qc |= (unsigned)(uint8_t)(int8_t)dct(4 * q4 + b, j, n) << (8 * b);
            }
            S.tb.Qr[o / 4 + e] = (int32_t)qr;
            S.tb.Qc[o / 4 + e] = (int32_t)qc;
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
    for (int i = tid; i < WB_TAB_N; i += NTHREADS) { S.tb.ldq[i] = tab->ldq[i]; S.tb.lv[i] = tab->lv[i]; }
    if (tid < 32) {
        unsigned pk = 0;
        for (int i = 0; i < 4; i++) pk |= ((unsigned)(uint8_t)c_fC[tid][i]) << (8 * i);
        S.tb.fc[tid] = (int)pk;
    }
    if (tid == 0) S.tb.ls_recip = (uint32_t)(0x100000000ull / (unsigned long long)tab->ls) + 1u;
    if (tid == 32) a4::fill_tap_tables(S.tb.taps, c_fC);
    for (int m = tid; m < 68; m += NTHREADS) {
        const int ang = m < 67 ? c_angle[m] : 0;
        const int inv = ang > 0 ? (512 * 32 + ang / 2) / ang : (ang < 0 ? -((512 * 32 + (-ang) / 2) / -ang) : 0);
        S.tb.ang[m] = (int8_t)ang;
        S.tb.invang[m] = (int16_t)inv;
        for (int l2 = 2; l2 <= 5; l2++) {  // intra_predictor.rs:1342-1366 (filter choice), 355-420 (PDPC nScale)
            unsigned ap = 0;
            if (m >= 2 && m <= 66) {
                if (!(m == 2 || m == 34 || m == 66)) {
                    const int md = min(abs(m - 50), abs(m - 18));
                    const int thr = l2 == 2 ? 24 : (l2 == 3 ? 14 : (l2 == 4 ? 2 : 0));
                    if (md > thr) ap |= 1u;
                }
                if (m <= 18 || m >= 50) {
                    const int ns = (m == 18 || m == 50) ? (2 * l2 - 2) >> 2 : min(l2 - ilog2i(3 * inv - 2) + 8, 2);
                    if (ns >= 0) ap |= 8u | ((unsigned)ns << 1);
                }
            }
            S.tb.angp[l2 - 2][m] = (uint8_t)ap;
        }
    }
}

// ---- TMA (bulk async copy) staging of the source blocks: global -> shared, completion on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WB_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WB_MBAR_DONE;\n"
        "bra WB_MBAR_WAIT;\n"
        "WB_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
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
    if (tid == 0) {
        S.ticket[0] = 0; S.ticket[1] = 0; S.ticket_big[0] = 0; S.ticket_big[1] = 0;
        mbar_init(&S.tma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int S_slot = 0;
    unsigned tma_phase = 0;
#ifdef WB_PROFILE
    if (tid == 0) S.prof_last = clock64();
#endif
    __syncthreads();
    const int W = P.W, H = P.H, Wc = P.Wc;
    const size_t pic_samples = (size_t)W * H * 3 / 2;
    const int cw = W >> 1, chh = H >> 1;
    for (;;) {
        if (tid == 0) S.item0 = (int)atomicAdd(P.counter, 1u) * KC;
        __syncthreads();
        const int item0 = S.item0;
        if (item0 >= P.n_items) break;
        WB_PROF(80);
        // ---- one thread per CTU: decode the item, wait for the wavefront dependencies (left CTU and above-right CTU,
        //      above when in the last column, of the same picture must be final)
        if (tid < KC) {
            CtuCtx &C = S.c[tid];
            const uint32_t it = P.items[item0 + tid];
            C.active = it != 0xffffffffu;
            if (C.active) {
                C.pic = it >> 16; C.cyi = (it >> 8) & 255; C.cxi = it & 255;
                C.g.cx = C.cxi * 32; C.g.cy = C.cyi * 32; C.g.W = W; C.g.H = H;
                C.mask = 0; C.root_mode = 0; C.dir_cnt = 0;
                C.groot = P.root_slots + ((size_t)blockIdx.x * KC + tid) * CTU_SCRATCH_BYTES;
                C.gsave = C.groot + ROOT_SLOT_BYTES;
                int *done = P.done + (size_t)C.pic * Wc * P.Hc;
                if (C.cxi > 0) while (ld_relaxed(&done[C.cyi * Wc + C.cxi - 1]) != P.epoch) __nanosleep(200);
                if (C.cyi > 0) {
                    const int ax = min(C.cxi + 1, Wc - 1);
                    while (ld_relaxed(&done[(C.cyi - 1) * Wc + ax]) != P.epoch) __nanosleep(200);
                }
            }
            __threadfence();  // acquire side: order the halo loads after the flag observation
        }
        __syncthreads();
        WB_PROF(81);
        // ---- stage the CTUs: source samples (TMA), neighbouring reconstruction, left-CTU modes
        if (tid == 0) {
            int nact = 0;
            for (int k = 0; k < KC; k++) nact += S.c[k].active;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the source buffers vs the async writes
            mbar_arrive_expect_tx(&S.tma_bar, (unsigned)nact * 1536u);
        }
        __syncthreads();
        for (int k = 0; k < KC; k++) {
            CtuCtx &C = S.c[k];
            if (!C.active) continue;
            Ctx V{&S.tb, &C};
            const CtuGeom g = C.g;
            const int pic = C.pic;
            const uint8_t *oY = P.orig + (size_t)pic * pic_samples, *oCb = oY + (size_t)W * H, *oCr = oCb + (size_t)cw * chh;
            const uint8_t *rY = P.rec + (size_t)pic * pic_samples, *rCb = rY + (size_t)W * H, *rCr = rCb + (size_t)cw * chh;
            // source block: 32 luma rows of 32 bytes + 2 x 16 chroma rows of 16 bytes, one TMA bulk copy per row, all
            // completing on S.tma_bar (armed above with the byte count of the whole batch)
            for (int i = tid; i < 64; i += NTHREADS) {
                if (i < 32) tma_bulk_g2s(&C.orgY[i * 32], oY + (size_t)(g.cy + i) * W + g.cx, 32, &S.tma_bar);
                else {
                    const int c = (i - 32) >> 4, yy = (i - 32) & 15;
                    tma_bulk_g2s(&C.orgC[c][yy * 16], (c ? oCr : oCb) + (size_t)((g.cy >> 1) + yy) * cw + (g.cx >> 1), 16, &S.tma_bar);
                }
            }
            // luma halo: rows -2,-1 (cols -4..63) and cols -4..-1 of rows 0..31
            for (int i = tid; i < 2 * RY_STRIDE + 32 * 4; i += NTHREADS) {
                int xx, yy;
                if (i < 2 * RY_STRIDE) { yy = -2 + i / RY_STRIDE; xx = -4 + i % RY_STRIDE; }
                else { int j = i - 2 * RY_STRIDE; yy = j >> 2; xx = -4 + (j & 3); }
                int ax = g.cx + xx, ay = g.cy + yy;
                uint8_t v = 0;
                if (ax >= 0 && ax < W && ay >= 0) v = __ldcg(rY + (size_t)ay * W + ax);
                RY(V, xx, yy) = v;
            }
            for (int i = tid; i < 2 * (RC_STRIDE + 16 * 4); i += NTHREADS) {
                int c = i >= (RC_STRIDE + 16 * 4), j = i - c * (RC_STRIDE + 16 * 4);
                int xx, yy;
                if (j < RC_STRIDE) { yy = -1; xx = -4 + j; }
                else { int q = j - RC_STRIDE; yy = q >> 2; xx = -4 + (q & 3); }
                int ax = (g.cx >> 1) + xx, ay = (g.cy >> 1) + yy;
                uint8_t v = 0;
                if (ax >= 0 && ax < cw && ay >= 0) v = __ldcg((c ? rCr : rCb) + (size_t)ay * cw + ax);
                RC(V, 1 + c, xx, yy) = v;
            }
            for (int i = tid; i < 8; i += NTHREADS)
                C.leftModes[i] = g.cx > 0 ? __ldcg(P.mode_map + (size_t)pic * (W >> 2) * (H >> 2) + (size_t)((g.cy >> 2) + i) * (W >> 2) + (g.cx >> 2) - 1) : 0;
            for (int i = tid; i < 1024; i += NTHREADS) C.lvY[i] = 0;
            for (int i = tid; i < 256; i += NTHREADS) { C.lvC[0][i] = 0; C.lvC[1][i] = 0; }
            for (int i = tid; i < 64; i += NTHREADS) C.lm[i] = 0;
            for (int i = tid; i < 16; i += NTHREADS) C.cm[i] = 0;
        }
        mbar_wait(&S.tma_bar, tma_phase);  // all source bytes have landed
        tma_phase ^= 1;
        __syncthreads();
        // transposed copy of the source blocks (the horizontal angular modes predict column-wise, four samples of a column per lane)
        for (int i = tid; i < KC * 1536; i += NTHREADS) {
            const int k = i / 1536, j = i - k * 1536;
            CtuCtx &C = S.c[k];
            if (!C.active) continue;
            if (j < 1024) C.orgYT[(j & 31) * 32 + (j >> 5)] = C.orgY[j];
            else {
                const int c = (j - 1024) >> 8, q = (j - 1024) & 255;
                C.orgCT[c][(q & 15) * 16 + (q >> 4)] = C.orgC[c][q];
            }
        }
        __syncthreads();
        WB_PROF(82);
        ctu_search(S, P, S_slot);
        __syncthreads();
        WB_PROF(83);
        // ---- write back: reconstruction, levels, modes, record
        for (int k = 0; k < KC; k++) {
            CtuCtx &C = S.c[k];
            if (!C.active) continue;
            Ctx V{&S.tb, &C};
            const CtuGeom g = C.g;
            const int pic = C.pic;
            uint8_t *rY = P.rec + (size_t)pic * pic_samples, *rCb = rY + (size_t)W * H, *rCr = rCb + (size_t)cw * chh;
            for (int i = tid; i < 256; i += NTHREADS) {
                int y = i >> 3, x4 = (i & 7) * 4;
                uint32_t v = (uint32_t)RY(V, x4, y) | ((uint32_t)RY(V, x4 + 1, y) << 8) | ((uint32_t)RY(V, x4 + 2, y) << 16) | ((uint32_t)RY(V, x4 + 3, y) << 24);
                *reinterpret_cast<uint32_t *>(rY + (size_t)(g.cy + y) * W + g.cx + x4) = v;
            }
            for (int i = tid; i < 128; i += NTHREADS) {
                int c = i >> 6, t = i & 63, yy = t >> 2, xx = (t & 3) * 4;
                uint32_t u = (uint32_t)RC(V, 1 + c, xx, yy) | ((uint32_t)RC(V, 1 + c, xx + 1, yy) << 8) | ((uint32_t)RC(V, 1 + c, xx + 2, yy) << 16) |
                             ((uint32_t)RC(V, 1 + c, xx + 3, yy) << 24);
                *reinterpret_cast<uint32_t *>((c ? rCr : rCb) + (size_t)((g.cy >> 1) + yy) * cw + (g.cx >> 1) + xx) = u;
            }
            int16_t *lY = P.lev + (size_t)pic * pic_samples, *lCb = lY + (size_t)W * H, *lCr = lCb + (size_t)cw * chh;
            for (int i = tid; i < 512; i += NTHREADS) {  // luma levels, 2 per iteration
                int yy = i >> 4, xx = (i & 15) * 2;
                *reinterpret_cast<uint32_t *>(lY + (size_t)(g.cy + yy) * W + g.cx + xx) = *reinterpret_cast<const uint32_t *>(&C.lvY[yy * 32 + xx]);
            }
            for (int i = tid; i < 256; i += NTHREADS) {
                int c = i >> 7, t = i & 127, yy = t >> 3, xx = (t & 7) * 2;
                *reinterpret_cast<uint32_t *>((c ? lCr : lCb) + (size_t)((g.cy >> 1) + yy) * cw + (g.cx >> 1) + xx) =
                    *reinterpret_cast<const uint32_t *>(&C.lvC[c][yy * 16 + xx]);
            }
            CtuRecord *r = P.records + (size_t)pic * Wc * P.Hc + C.cyi * Wc + C.cxi;
            for (int i = tid; i < 64; i += NTHREADS) {
                r->luma_mode[i] = C.lm[i];
                P.mode_map[(size_t)pic * (W >> 2) * (H >> 2) + (size_t)((g.cy >> 2) + (i >> 3)) * (W >> 2) + (g.cx >> 2) + (i & 7)] = C.lm[i];
            }
            for (int i = tid; i < 16; i += NTHREADS) r->chroma_mode[i] = C.cm[i];
            if (tid == 0) { r->split_mask = C.mask; r->cost = C.leaf_cost; }
        }
        __threadfence();
        __syncthreads();
        if (tid < KC && S.c[tid].active) {
            const CtuCtx &C = S.c[tid];
            st_release(P.done + (size_t)C.pic * Wc * P.Hc + C.cyi * Wc + C.cxi, P.epoch);
        }
        __syncthreads();
        WB_PROF(84);
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
    Ctx V{&S.tb, &S.c[0]};
    if (P.op == 0) {  // predict one component of one TU from a picture's reconstruction
        CtuGeom g;
        g.cx = P.x & ~31; g.cy = P.y & ~31; g.W = P.W; g.H = P.H;
        const int cw = P.W >> 1, chh = P.H >> 1;
        const uint8_t *rY = P.rec, *rCb = rY + (size_t)P.W * P.H, *rCr = rCb + (size_t)cw * chh;
        for (int i = tid; i < RY_ROWS * RY_STRIDE; i += NTHREADS) {
            int yy = i / RY_STRIDE - RY_Y0, xx = i % RY_STRIDE - RY_X0;
            int ax = g.cx + xx, ay = g.cy + yy;
            S.c[0].recY[i] = (ax >= 0 && ax < P.W && ay >= 0 && ay < P.H) ? rY[(size_t)ay * P.W + ax] : 0;
        }
        for (int i = tid; i < 2 * RC_ROWS * RC_STRIDE; i += NTHREADS) {
            int c = i >= RC_ROWS * RC_STRIDE, j = i - c * RC_ROWS * RC_STRIDE;
            int yy = j / RC_STRIDE - RC_Y0, xx = j % RC_STRIDE - RC_X0;
            int ax = (g.cx >> 1) + xx, ay = (g.cy >> 1) + yy;
            S.c[0].recC[c][j] = (ax >= 0 && ax < cw && ay >= 0 && ay < chh) ? (c ? rCr : rCb)[(size_t)ay * cw + ax] : 0;
        }
        __syncthreads();
        Node nd;
        nd.x = P.x - g.cx; nd.y = P.y - g.cy; nd.w = P.w; nd.tree = P.tree; nd.ar = P.ar != 0; nd.bl = P.bl != 0;
        if (warp == 0) {
            const int c = P.c, n = nd.w >> (c != 0);
            if (P.mode > 66) cclm_downsample(V, g, nd, lane);
            else build_refs(V, g, nd, c, lane);
            __syncwarp();
            predict_block(V, g, nd, c, P.mode, ws.refx, ws.pred, lane);  // the search kernel's own prediction path (the SAD it returns is not used here)
            __syncwarp();
            for (int i = lane; i < n * n; i += 32) P.out8[i] = ws.pred[i];
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
            mm_rows_q<true>(S.tb.Qr + to / 4, ws.A, ws.B, n, l2, 1 << (l2 - 2), l2 - 1, lane);
            __syncwarp();
            mm_cols_q(S.tb.T + to, reinterpret_cast<const int32_t *>(ws.B), ws.A, n, l2, 1 << (l2 + 5), l2 + 6, false, lane);
            __syncwarp();
            for (int i = lane; i < nn; i += 32) out[i] = ws.A[i];
        } else if (P.op == 2) {
            for (int o = lane; o < nn / 2; o += 32) {  // pair-packed layout of the column pass
                const int ip = o >> l2, x = o & (n - 1);
                reinterpret_cast<int32_t *>(ws.B)[o] = ((int)ws.A[(2 * ip) * n + x] & 0xffff) | ((int)ws.A[(2 * ip + 1) * n + x] << 16);
            }
            __syncwarp();
            mm_cols_q(S.tb.Tt + to, reinterpret_cast<const int32_t *>(ws.B), ws.A, n, l2, 64, 7, true, lane);  // V[y][x] = sum_i T[i][y] D[i][x]
            __syncwarp();
            for (int i = lane; i < nn; i += 32) ws.B[i] = ws.A[i];
            __syncwarp();
            mm_rows_q<false>(S.tb.Qc + to / 4, ws.B, ws.A, n, l2, 2048, 12, lane);  // R[y][x] = sum_i T[i][x] V[y][i]
            __syncwarp();
            for (int i = lane; i < nn; i += 32) out[i] = ws.A[i];
        } else if (P.op == 3 || P.op == 5) {  // 3: the routine the search uses for this size; 5: trellis() for every size
            int rate; bool any;
#if WB_CHAIN8
            if (l2 == 3 && P.op == 3) trellis8_chain(V, P.tab, ws.A, reinterpret_cast<uint16_t *>(ws.A + 128), ws.A, reinterpret_cast<int4 *>(ws.B), lane, rate, any);
            else
#endif
#if WB_CHAIN16
            if (l2 == 4 && P.op == 3) trellis16_chain(V, P.tab, ws.A, reinterpret_cast<uint8_t *>(ws.A + 256), reinterpret_cast<int4 *>(ws.A + 384), lane, rate, any);
            else
#endif
            trellis(V, P.tab, ws.A, l2, ws.Wd, ws.A, lane, rate, any);  // levels in place, as the search does
            for (int i = lane; i < nn; i += 32) out[i] = ws.A[i];
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

#ifdef WB_PROFILE
extern "C" int wrenc_b200_debug_prof(unsigned long long *out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_prof, sizeof(g_prof));
    if (e == cudaSuccess && reset) {
        static unsigned long long zero[128];
        e = cudaMemcpyToSymbol(g_prof, zero, sizeof(zero));
    }
    return e == cudaSuccess ? 0 : -1;
}
#endif

size_t search_smem_bytes() { return sizeof(Shared); }
int search_ctus_per_cta() { return KC; }

static size_t smem_request() {  // dev-time knob: WRENC_B200_SMEM_PAD bytes of extra dynamic shared memory lower the CTAs per SM
    size_t pad = 0;
    if (const char *e = getenv("WRENC_B200_SMEM_PAD")) pad = (size_t)atol(e);
    return sizeof(Shared) + pad;
}

// The dynamic shared memory opt-in is a per-device function attribute: set once per device, not once per process.
static cudaError_t ensure_smem_attr() {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(wrenc_b200_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_request());
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    return cudaSuccess;
}

cudaError_t launch_search(const SearchParams &P, int grid, cudaStream_t stream, const void *persist_base, size_t persist_bytes) {
    cudaError_t e = ensure_smem_attr();
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem_request();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (persist_base && persist_bytes) {  // per-launch attribute: the caller's stream attributes are left alone
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void *>(persist_base);
        attr[0].val.accessPolicyWindow.num_bytes = persist_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    e = cudaLaunchKernelEx(&cfg, wrenc_b200_search_kernel, P);
    if (e != cudaSuccess && cfg.numAttrs) {  // best effort: without the window the launch only costs DRAM traffic
        cudaGetLastError();
        cfg.numAttrs = 0;
        e = cudaLaunchKernelEx(&cfg, wrenc_b200_search_kernel, P);
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

int search_ctas_per_sm() {
    int n = 0;
    cudaFuncSetAttribute(wrenc_b200_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_request());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wrenc_b200_search_kernel, NTHREADS, smem_request());
    return n;
}

}  // namespace wb
