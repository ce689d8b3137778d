// search_kernel.cuh — the all-intra RD search of wrenc's CTUs, hand-written for sm_100a: block functions (one warp or half a
// warp per block) and the shared-memory layout.  The phase control that drives them is in search.cu.
//
// What it computes (reference file:line, all under /root/reference/src):
//   split_ct quad-tree RD decision 32->16->8->4 (+ local dual-tree chroma CT)      block_splitter.rs:782-1154
//   leaf evaluation: 15 coarse modes, SAD step search, 3 RD evals, chroma DM vs CCLM  block_splitter.rs:886-1078, 794-885
//   full evaluation (predict, fwd DCT, dep-quant trellis, dequant, inv DCT, recon, SSD, rate)  block_splitter.rs:110-474,524-780
//   intra prediction incl. reference substitution/filter, PDPC, CCLM                intra_predictor.rs:56-2055
//   DCT-II 4..32 by matrix multiply (dp2a)                                          transformer.rs:2040-2378,2380-2737
//   dependent quantisation (memoised DFS == Viterbi with first-visit flags, H2/H3)  quantizer.rs:338-759, dequantize 761-1079
//   MPM derivation for the mode-bit estimate (H1: in-CTU neighbours see the root CU) ctu.rs:1498-1635
//
// Execution model: a persistent grid, one CTA of WB_NW warps per SM; each CTA pulls batches of WB_K
// mutually independent CTUs from a work list sorted in wavefront order (key = x + 2y, all pictures in lock step), waits on
// the left and above-right CTU's done flags, stages the CTUs' source (TMA) and neighbouring reconstruction in shared memory
// and runs the whole tree search of the batch there in lock step.  The candidates of one decision phase are independent
// tasks, one warp (4x4 TBs: half a warp) each.  What is written once and read back at most once (candidate slots, saved
// no-split states) lives in a per-CTA global scratch.  All arithmetic is exact integer except the RD cost, which is IEEE f32
// with explicit _rn intrinsics (no FMA contraction) in the reference's operation order.
//
// Code size is a first-order effect here: the kernel is ~15 000 instructions (234 kB) against a 32 kB
// instruction cache, and lock step is what keeps the 16 warps of an SM in the same few kB at any moment.  Hence the rolled copy
// loops (WB_RU), the constant-mask collectives, and the build knobs below, each of which records what it measured.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ang4.cuh"
#include "search_kernel_api.h"

namespace wb {

#ifndef WB_NW
#define WB_NW 16
#endif
#ifndef WB_MINB
#define WB_MINB 1
#endif
#ifndef WB_U5
#define WB_U5 5  // unroll of the 15-step state-per-lane pass of full_pair4
#endif
#ifndef WB_U2
#define WB_U2 1  // unroll of the dp2a inner loops (measured: 1 beats 2 by 0.8 %, code size)
#endif
#ifndef WB_RU
#define WB_RU 1  // unroll of the copy / element-wise loops that are not on a critical path: rolled (4, what the compiler does by itself, is 7 % more code and measured 2.2 % slower)
#endif
#ifndef WB_CHAIN16
#define WB_CHAIN16 0  // 1: 16x16 TBs quantised by trellis16_chain too (bit-exact; measured 4.6 % slower: a 255-step chain is longer than the chunked pass)
#endif
#ifndef WB_CHAIN8
#define WB_CHAIN8 1  // 8x8 TBs quantised by trellis8_chain (one sequential state-per-lane chain) instead of trellis()
#endif
#ifndef WB_U8C
#define WB_U8C 5  // unroll of the 15 steps per sub-block of trellis8_chain
#endif
constexpr int U5 = WB_U5, U2 = WB_U2, RU = WB_RU, U8C = WB_U8C;
constexpr int NW = WB_NW;    // warps per CTA
constexpr int NTHREADS = NW * 32;
#ifndef WB_K
#define WB_K 8
#endif
constexpr int KC = WB_K;     // CTUs searched in lock step by one CTA
#ifndef WB_NBIG
#define WB_NBIG WB_NW
#endif
constexpr int NBIG = NW < WB_NBIG ? NW : WB_NBIG;  // warps that take the leading ('big') tasks of a phase first: all of them since every warp owns a 5 kB scratch


// ---------------------------------------------------------------------------------------------------------------
// constant tables
// ---------------------------------------------------------------------------------------------------------------
__constant__ int8_t c_angle[67] = {0,   0,   32,  29,  26,  23,  20,  18,  16,  14,  12,  10,  8,   6,   4,   3,   2,
                                   1,   0,   -1,  -2,  -3,  -4,  -6,  -8,  -10, -12, -14, -16, -18, -20, -23, -26, -29,
                                   -32, -29, -26, -23, -20, -18, -16, -14, -12, -10, -8,  -6,  -4,  -3,  -2,  -1,  0,
                                   1,   2,   3,   4,   6,   8,   10,  12,  14,  16,  18,  20,  23,  26,  29,  32};
// VVC Table 25, fC (common.rs:153-186); fG is the closed form {16-(p>>1), 32-(p>>1), 16+(p>>1), p>>1}
__constant__ int8_t c_fC[32][4] = {
    {0, 64, 0, 0},    {-1, 63, 2, 0},   {-2, 62, 4, 0},   {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
    {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
    {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
    {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
    {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
    {0, 4, 62, -2},   {0, 2, 63, -1}};
// DCT-II magnitudes c[j] = 32-point basis value at angle j*pi/64 (even rows of the 64-point VVC matrix, transformer.rs:934-1234)
__constant__ int8_t c_cos32[33] = {91, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                   61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
__constant__ uint8_t c_divsig[16] = {0, 7, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1, 1, 1, 0};
__constant__ uint8_t c_cand15[15] = {0, 1, 2, 7, 13, 18, 23, 29, 34, 39, 45, 50, 55, 60, 66};

__host__ __device__ constexpr int tab_off(int l2) { return l2 == 2 ? 0 : (l2 == 3 ? 16 : (l2 == 4 ? 80 : 336)); }
constexpr int TAB_TOTAL = 16 + 64 + 256 + 1024;

// ---------------------------------------------------------------------------------------------------------------
// shared memory layout
// ---------------------------------------------------------------------------------------------------------------
constexpr int RY_STRIDE = 68, RY_X0 = 4, RY_Y0 = 2, RY_ROWS = 34;  // luma window: rows -2..31, cols -4..63
constexpr int RC_STRIDE = 36, RC_X0 = 4, RC_Y0 = 1, RC_ROWS = 17;  // chroma window: rows -1..15, cols -4..31

struct Tables {
    int8_t T[TAB_TOTAL];       // DCT matrix per size, row-major  T[i*n + x]
    int8_t Tt[TAB_TOTAL];      // transposed                      Tt[x*n + i]
    int32_t Qr[TAB_TOTAL / 4]; // packed for the row passes: Qr[x4*n + i] = bytes T[i][4*x4 + 0..3]   (forward horizontal pass)
    int32_t Qc[TAB_TOTAL / 4]; //                            Qc[i4*n + x] = bytes T[4*i4 + 0..3][x]   (inverse horizontal pass)
    uint16_t scan[TAB_TOTAL];  // forward scan index k (sub-block diagonal, then 4x4 diagonal; ctu.rs:14-81) -> raster offset
#ifndef WB_TAB_N
#define WB_TAB_N 1024  // entries of the two rate tables kept in shared memory (1024 = all: no global-memory fallback branch in the trellis; 64: +8 kB of L1, measured 1.6 % slower)
#endif
    int32_t ldq[WB_TAB_N];
    int32_t lv[WB_TAB_N];
    int32_t fc[32];            // fC taps packed as 4 x int8
    int16_t invang[68];        // inverse angle per mode (sign of the angle kept; intra_predictor.rs:1330-1341)
    int8_t ang[68];            // intraPredAngle per mode
    uint32_t ls_recip;         // floor(2^32 / ls) + 1: the trellis divides by the (launch-constant) level scale, see div_ls
    uint8_t angp[4][68];       // per block size (log2 - 2) and angular mode: bit 0 fG taps (luma), bits 1-2 PDPC nScale, bit 3 PDPC applies
    a4::TapTables taps;        // fC / fG / doubled chroma taps, packed 4 x int8 per iFact (ang4.cuh)
};

struct WarpScratch {  // pointers into the scratch pool
    int16_t *A, *B;
    uint16_t *Wd;
    uint8_t *pred;
    int16_t *refx;  // 3n+3 entries, index n+idx for idx in [-n, 2n+2]
};

constexpr int MAXTASK = 48;
constexpr int LN_OFF = 40, LN_SIZE = 120;  // byte reference lines: index range [-32 - slack, 2 * 32 + 3 + slack]

struct CtuGeom {
    int cx, cy;  // absolute luma position of the CTU
    int W, H;
};

// Everything that belongs to ONE CTU while it is searched.  A CTA searches WB_K independent CTUs in lock step.
struct alignas(16) CtuCtx {
    uint8_t orgY[1024];       // source block; filled by TMA bulk copies (16-byte aligned rows)
    uint8_t orgC[2][256];
    uint8_t recY[RY_ROWS * RY_STRIDE];
    uint8_t recC[2][RC_ROWS * RC_STRIDE];
    int16_t lvY[1024];
    int16_t lvC[2][256];
    uint8_t lm[64], cm[16];
    uint8_t svLm[3][64], svCm[3][16];  // modes of the no-split state saved per depth (0: 32x32, 1: 16x16, 2: 8x8); samples and levels: gsave
    // reference samples of the current node: [comp][raw|filtered]
    int16_t seq[3][132];      // after build_refs: the substituted samples in one line, left[2n] ... left[1], corner, above[0] ... above[2n-1]
    int16_t seqF[132];        // the same line of the [1 2 1] filtered luma references
    int16_t refL[3][2][68];
    int16_t refA[3][2][64];
    // the same reference samples as BYTES in two mirrored lines per component (ang4.cuh): ln[v][0] = up (corner, above...),
    // ln[v][1] = dn (corner, left...), v = 0 luma, 1 luma [1 2 1]-filtered, 2 Cb, 3 Cr; index 0 at byte LN_OFF, valid [-n, 2n+3]
    alignas(8) uint8_t ln[4][2][LN_SIZE];
    alignas(16) uint8_t orgYT[1024];     // the source block transposed (orgYT[x * 32 + y]): the horizontal angular modes predict column-wise
    alignas(16) uint8_t orgCT[2][256];
    uint8_t pds[256];         // CCLM down-sampled luma of the current node
    uint8_t leftModes[8];     // final luma modes of the left CTU's right-most 4x4 column
    // task results
    uint32_t r_ssd[MAXTASK];
    int32_t r_rate[MAXTASK];
    uint32_t r_sad[MAXTASK];
    // control state: written by the CTU's decision thread, read by everyone after the next barrier
    CtuGeom g;
    int active, pic, cxi, cyi;
    unsigned node;  // packed geometry + availability flags of the node being evaluated (pack_node)
    int root_mode;
    unsigned mask, mask_sv[2];
    float cost[3], split[3], leaf_cost;
    int restore;
    // leaf-evaluation state
    float cost_pl, cost_dc, cur_cost, dir_cost, min_cost, cost_dm;
    int cur, dir, mode, cclm_mode, v0, v1, cclm_wins, dir_cand;
    uint8_t *groot;           // this CTU's global scratch (SearchParams::root_slots): candidate slots of the CU being evaluated
    uint8_t *gsave;           //   and the no-split states saved per depth
    unsigned dir_part[4];     // direction search of CUs up to 8x8: first minimum of each part of the coarse modes, and the arrival counter
    int dir_cnt;
    // results of the planar / DC evaluations (phase 1) and of the winner (phase 5), per component
    unsigned pd_ssd[2][3], fin_ssd[3];
    int pd_rate[2][3], fin_rate[3];
};

struct Shared {
    Tables tb;
    CtuCtx c[WB_K];
    int item0;
#ifdef WB_PROFILE
    long long prof_last;
#endif
    unsigned long long tma_bar;  // mbarrier the TMA bulk copies of the source blocks complete on
    int ticket[2];            // dynamic task tickets of the current / previous phase
    int ticket_big[2];        //   and of the leading long tasks of a phase (32x32 luma pipelines of the root), handed out first
    // per-warp scratch
    // per-warp scratch, the same for every warp (5 kB: a 32x32 luma pipeline fits): A residual -> coefficients -> levels (the
    // dependent quantisation writes the levels in place) -> reconstruction residual; B transform intermediate, then the
    // trellis words of trellis() / the 1 kB cost table of trellis8_chain, then the dequantised block; P the prediction
    struct alignas(16) WarpBuf {
        int16_t A[1024];
        int16_t B[1024];
        uint8_t P[1024];
    } wbuf[NW];
    alignas(16) int16_t refx[NW][100];  // per-warp scratch line: projected references of the negative-angle modes (as bytes)
};

static_assert(offsetof(CtuCtx, lvY) % 8 == 0 && offsetof(CtuCtx, lvC) % 8 == 0 && sizeof(CtuCtx) % 8 == 0, "commit_root_slot stores the levels as 64-bit words");
static_assert(offsetof(Shared, wbuf) % 16 == 0 && sizeof(Shared::WarpBuf) % 16 == 0 && offsetof(Tables, Tt) % 4 == 0,
              "full_pair4 reads these arrays as 32-bit words, trellis8_chain stores 16-byte rows");

struct Ctx {  // what the per-CTU device functions see
    Tables *tb;
    CtuCtx *c;
};

struct Node {
    int x, y, w;  // CTU-relative luma position / size
    int tree;
    bool ar, bl;
};

// Every pointer the block functions receive points into shared memory; telling the compiler lets it emit LDS/STS instead of
// generic loads/stores in the out-of-line functions.
#define WB_SHARED_PTR(p) __builtin_assume(__isShared(p))
#define WB_SHARED_CTX(S) do { WB_SHARED_PTR((S).c); WB_SHARED_PTR((S).tb); } while (0)

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ilog2i(int v) { return 31 - __clz(v); }
__device__ __forceinline__ int clip8(int v) { return min(255, max(0, v)); }
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sumu(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over one half of the warp (two independent 4x4 blocks per warp); mask = the half's lanes
__device__ __forceinline__ int half_sum(int v, unsigned mask) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint8_t &RY(const Ctx S, int x, int y) { return S.c->recY[(y + RY_Y0) * RY_STRIDE + x + RY_X0]; }
__device__ __forceinline__ uint8_t &RC(const Ctx S, int c, int x, int y) { return S.c->recC[c - 1][(y + RC_Y0) * RC_STRIDE + x + RC_X0]; }
__device__ __forceinline__ int rec_at(const Ctx S, int c, int x, int y) { return c == 0 ? RY(S, x, y) : RC(S, c, x, y); }
__device__ __forceinline__ int org_at(const Ctx S, int c, int x, int y) { return c == 0 ? S.c->orgY[y * 32 + x] : S.c->orgC[c - 1][y * 16 + x]; }

// encoder_context.rs:918-956 derive_neighbouring_block_availability, CTU-relative luma coordinates (H9)
__device__ __forceinline__ bool nb_avail(const CtuGeom g, const Node nd, int xn, int yn, bool ar, bool bl) {
    int ax = g.cx + xn, ay = g.cy + yn;
    return ax >= 0 && ay >= 0 && ax < g.W && ay < g.H && (((xn >> 5) <= 0) || ((yn >> 5) < 0)) && ((yn >> 5) < 1) &&
           (xn < nd.x + nd.w || ar) && (yn < nd.y + nd.w || bl);
}

__device__ __forceinline__ float rd_cost(unsigned ssd, long long level, float lambda) {
    // block_splitter.rs:472-473 / 779: ssd as f32 + lambda * (level as f32 / 16384.0)
    // (the division by 2^14 is an exact scaling, so it is done as a multiplication by 2^-14: bit-identical, no division sequence)
    return __fadd_rn(__uint2float_rn(ssd), __fmul_rn(lambda, __fmul_rn(__ll2float_rn(level), 6.103515625e-05f)));
}

// ---------------------------------------------------------------------------------------------------------------
// reference samples (intra_predictor.rs:146-353), one warp per component
// ---------------------------------------------------------------------------------------------------------------
// byte reference lines (ang4.cuh): store left sample li (li = 0: the corner) / above sample ai of line variant v into both
// mirrored lines, with the three copies of the last sample that stand for ref[min(idx, 2n)]
__device__ __forceinline__ int ln_variant(int c, int filt) { return c == 0 ? filt : 1 + c; }
__device__ __forceinline__ void ln_put_left(const Ctx S, int v, int li, int n, int val) {
    uint8_t *up = S.c->ln[v][0] + LN_OFF, *dn = S.c->ln[v][1] + LN_OFF;
    dn[li] = (uint8_t)val;
    if (li <= n) up[-li] = (uint8_t)val;
    if (li == 2 * n) { dn[li + 1] = (uint8_t)val; dn[li + 2] = (uint8_t)val; dn[li + 3] = (uint8_t)val; }
}
__device__ __forceinline__ void ln_put_above(const Ctx S, int v, int ai, int n, int val) {
    uint8_t *up = S.c->ln[v][0] + LN_OFF, *dn = S.c->ln[v][1] + LN_OFF;
    up[1 + ai] = (uint8_t)val;
    if (1 + ai <= n) dn[-(1 + ai)] = (uint8_t)val;
    if (ai == 2 * n - 1) { up[2 * n + 1] = (uint8_t)val; up[2 * n + 2] = (uint8_t)val; up[2 * n + 3] = (uint8_t)val; }
}

__device__ __noinline__ void build_refs(const Ctx S, const CtuGeom g, const Node nd, int c, int lane) {
    WB_SHARED_CTX(S);
    const int cs = c != 0;
    const int n = nd.w >> cs, xt = nd.x >> cs, yt = nd.y >> cs;
    const int nl = 2 * n + 1, na = 2 * n, tot = nl + na;
    int16_t *seq = S.c->seq[c];
    const int rounds = (tot + 31) >> 5;
    // sequence order: left[nl-1] ... left[0], above[0] ... above[na-1].  Pass 1: the available samples (-1: not available) and the
    // first available position.  (Rolled loops on purpose: this function runs at the start of every node, and what it costs is the
    // instruction-cache lines it touches)
    int first = -1;
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const int j = r * 32 + lane;
        int v = -1;
        if (j < tot) {
            if (j < nl) {
                const int y = (nl - 1 - j) - 1;
                const int yr = (y == -1) ? -1 : (y & ~3);
                if (nb_avail(g, nd, (xt - 1) << cs, (yt + yr) << cs, nd.ar, nd.bl)) v = rec_at(S, c, xt - 1, yt + y);
            } else {
                const int x = j - nl;
                const int xr = x & ~3;
                if (nb_avail(g, nd, (xt + xr) << cs, (yt - 1) << cs, nd.ar, nd.bl)) v = rec_at(S, c, xt + x, yt - 1);
            }
            seq[j] = (int16_t)v;
        }
        const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
        if (first < 0 && m) first = r * 32 + __ffs(m) - 1;
    }
    __syncwarp();
    // Pass 2: substitution = the nearest available sample at or before the position, else the first available one, else 128
    // (intra_predictor.rs:214-301).  A source position is always an available one, which this pass leaves unchanged.
    int16_t *L = S.c->refL[c][0], *A = S.c->refA[c][0];
    int prev = -1;  // last available position of the rounds before this one
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const int j = r * 32 + lane;
        const int own = j < tot ? (int)seq[j] : -1;
        const unsigned m = __ballot_sync(0xffffffffu, own >= 0);
        int val = 128;
        if (j < tot && first >= 0) {
            const unsigned mm = m & (0xffffffffu >> (31 - lane));
            const int src = mm ? r * 32 + 31 - __clz(mm) : (prev >= 0 ? prev : first);
            val = seq[src];
        }
        __syncwarp();
        if (j < tot) {
            if (j < nl) { L[nl - 1 - j] = (int16_t)val; ln_put_left(S, ln_variant(c, 0), nl - 1 - j, n, val); }
            else { A[j - nl] = (int16_t)val; ln_put_above(S, ln_variant(c, 0), j - nl, n, val); }
            seq[j] = (int16_t)val;
        }
        if (m) prev = r * 32 + 31 - __clz(m);
    }
    __syncwarp();
    if (c == 0 && n >= 8) {  // [1 2 1] filtered copy, used by modes 0,2,34,66 (intra_predictor.rs:304-352)
        int16_t *LF = S.c->refL[0][1], *AF = S.c->refA[0][1];
#pragma unroll RU
        for (int i = lane; i < nl; i += 32) {
            int v;
            if (i == 0) v = (L[1] + 2 * L[0] + A[0] + 2) >> 2;
            else if (i == nl - 1) v = L[i];
            else v = (L[i + 1] + 2 * L[i] + L[i - 1] + 2) >> 2;
            LF[i] = (int16_t)v;
            S.c->seqF[nl - 1 - i] = (int16_t)v;
            ln_put_left(S, 1, i, n, v);
        }
#pragma unroll RU
        for (int i = lane; i < na; i += 32) {
            int v;
            if (i == 0) v = (L[0] + 2 * A[0] + A[1] + 2) >> 2;
            else if (i == na - 1) v = A[i];
            else v = (A[i - 1] + 2 * A[i] + A[i + 1] + 2) >> 2;
            AF[i] = (int16_t)v;
            S.c->seqF[nl + i] = (int16_t)v;
            ln_put_above(S, 1, i, n, v);
        }
    }
    __syncwarp();
}

// The same for a 4x4 block (any component): 9 left + 8 above samples, one per lane, substitution by shuffles.
__device__ __noinline__ void build_refs4(const Ctx S, const CtuGeom g, const Node nd, int c, int lane) {
    WB_SHARED_CTX(S);
    const int cs = c != 0, xt = nd.x >> cs, yt = nd.y >> cs;
    constexpr int nl = 9, tot = 17;
    const int j = lane;
    int v = -1;
    if (j < tot) {
        if (j < nl) {
            const int y = (nl - 1 - j) - 1;
            const int yr = (y == -1) ? -1 : (y & ~3);
            if (nb_avail(g, nd, (xt - 1) << cs, (yt + yr) << cs, nd.ar, nd.bl)) v = rec_at(S, c, xt - 1, yt + y);
        } else {
            const int x = j - nl;
            if (nb_avail(g, nd, (xt + (x & ~3)) << cs, (yt - 1) << cs, nd.ar, nd.bl)) v = rec_at(S, c, xt + x, yt - 1);
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, v >= 0);
    const unsigned mm = mask & (0xffffffffu >> (31 - lane));
    const int src = mm ? 31 - __clz(mm) : (mask ? __ffs(mask) - 1 : 0);
    int val = __shfl_sync(0xffffffffu, v, src);
    if (!mask) val = 128;
    if (j < tot) {
        if (j < nl) { S.c->refL[c][0][nl - 1 - j] = (int16_t)val; ln_put_left(S, ln_variant(c, 0), nl - 1 - j, 4, val); }
        else { S.c->refA[c][0][j - nl] = (int16_t)val; ln_put_above(S, ln_variant(c, 0), j - nl, 4, val); }
        S.c->seq[c][j] = (int16_t)val;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// prediction (per-sample evaluation after a per-task setup)
// ---------------------------------------------------------------------------------------------------------------
struct PredCtx {
    int kind;  // 0 planar, 1 dc, 3 cclm (angular modes: ang4.cuh)
    int mode, c, n, l2;
    const int16_t *lf;  // left incl. corner at [0]
    const int16_t *ab;  // above
    int dc, nscale;
    int a, k, b;        // cclm
    bool cclm128;
};

__device__ __forceinline__ int pdpc_w(int ns, int i) {
    int s = (2 * i) >> ns;
    return s > 5 ? 0 : (32 >> s);
}

// CCLM luma accessor with the replication rules of intra_predictor.rs:1775-1818 (only the reachable cases)
__device__ __forceinline__ int cclm_py(const Ctx S, int bx, int by, bool avail_l, int y, int x) {
    if (x < 0 && !avail_l) x = 0;
    return RY(S, bx + x, by + y);
}
__device__ __forceinline__ int cclm_ds6(const Ctx S, int bx, int by, bool avail_l, int sy, int sx) {
    return (cclm_py(S, bx, by, avail_l, sy, sx - 1) + cclm_py(S, bx, by, avail_l, sy + 1, sx - 1) + 2 * cclm_py(S, bx, by, avail_l, sy, sx) +
            2 * cclm_py(S, bx, by, avail_l, sy + 1, sx) + cclm_py(S, bx, by, avail_l, sy, sx + 1) + cclm_py(S, bx, by, avail_l, sy + 1, sx + 1) + 4) >> 3;
}

// down-sampled luma of the node (intra_predictor.rs:1854-1868), one warp
__device__ __noinline__ void cclm_downsample(const Ctx S, const CtuGeom g, const Node nd, int lane) {
    WB_SHARED_CTX(S);
    Node tmp = nd;
    bool avail_l = nb_avail(g, tmp, nd.x - 1, nd.y, false, false);
    int tw = nd.w >> 1;
#pragma unroll RU
    for (int i = lane; i < tw * tw; i += 32) {
        int y = i >> ilog2i(tw), x = i & (tw - 1);
        S.c->pds[i] = (uint8_t)cclm_ds6(S, nd.x, nd.y, avail_l, 2 * y, 2 * x);
    }
}

// derive (a,k,b) for one CCLM mode / component (intra_predictor.rs:1604-2031); uniform across the warp
__device__ __noinline__ void cclm_params(const Ctx S, const CtuGeom g, const Node nd, int c, int mode, PredCtx &pc) {
    WB_SHARED_CTX(S);
    const int tw = nd.w >> 1, th = tw, tx = nd.x >> 1, ty = nd.y >> 1;
    bool avail_l = nb_avail(g, nd, nd.x - 1, nd.y, false, false);
    bool avail_t = nb_avail(g, nd, nd.x, nd.y - 1, false, false);
    int num_tr = 0, num_bl = 0;
    if (mode == MODE_T_CCLM) {
        for (int x = tw; x < 2 * tw; x++) {
            if (!nb_avail(g, nd, nd.x + x * 2, nd.y - 1, nd.ar, nd.bl)) break;
            num_tr++;
        }
    }
    if (mode == MODE_L_CCLM) {
        for (int y = th; y < 2 * th; y++) {
            if (!nb_avail(g, nd, nd.x - 1, nd.y + y * 2, nd.ar, nd.bl)) break;
            num_bl++;
        }
    }
    int num_t, num_l;
    if (mode == MODE_LT_CCLM) {
        num_t = avail_t ? tw : 0;
        num_l = avail_l ? th : 0;
    } else {
        num_t = (avail_t && mode == MODE_T_CCLM) ? tw + min(num_tr, th) : 0;
        num_l = (avail_l && mode == MODE_L_CCLM) ? th + min(num_bl, tw) : 0;
    }
    pc.cclm128 = (num_l == 0 && num_t == 0);
    pc.a = 0; pc.k = 0; pc.b = 128;
    if (pc.cclm128) return;
    const bool ctu_boundary = nd.y == 0;  // (tu.y & 31) == 0
    const int is4 = !(avail_t && avail_l && mode == MODE_LT_CCLM);
    int cnt_t = 0, cnt_l = 0;
    int start_t = num_t >> (2 + is4), step_t = max(1, num_t >> (1 + is4));
    int start_l = num_l >> (2 + is4), step_l = max(1, num_l >> (1 + is4));
    if (avail_t && (mode == MODE_LT_CCLM || mode == MODE_T_CCLM)) cnt_t = min((1 + is4) << 1, num_t);
    if (avail_l && (mode == MODE_LT_CCLM || mode == MODE_L_CCLM)) cnt_l = min((1 + is4) << 1, num_l);
    int sy[4] = {0, 0, 0, 0}, sc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < cnt_t) {
            int p = start_t + i * step_t;
            sc[i] = RC(S, c, tx + p, ty - 1);
            int sx = 2 * p;
            if (!ctu_boundary)
                sy[i] = (cclm_py(S, nd.x, nd.y, avail_l, -1, sx - 1) + cclm_py(S, nd.x, nd.y, avail_l, -2, sx - 1) +
                         2 * cclm_py(S, nd.x, nd.y, avail_l, -1, sx) + 2 * cclm_py(S, nd.x, nd.y, avail_l, -2, sx) +
                         cclm_py(S, nd.x, nd.y, avail_l, -1, sx + 1) + cclm_py(S, nd.x, nd.y, avail_l, -2, sx + 1) + 4) >> 3;
            else
                sy[i] = (cclm_py(S, nd.x, nd.y, avail_l, -1, sx - 1) + 2 * cclm_py(S, nd.x, nd.y, avail_l, -1, sx) +
                         cclm_py(S, nd.x, nd.y, avail_l, -1, sx + 1) + 2) >> 2;
        } else if (i < cnt_t + cnt_l) {
            int p = start_l + (i - cnt_t) * step_l;
            sc[i] = RC(S, c, tx - 1, ty + p);
            sy[i] = cclm_ds6(S, nd.x, nd.y, avail_l, 2 * p, -2);
        }
    }
    int mn0 = 0, mn1 = 2, mx0 = 1, mx1 = 3, t;
    if (sy[mn0] > sy[mn1]) { t = mn0; mn0 = mn1; mn1 = t; }
    if (sy[mx0] > sy[mx1]) { t = mx0; mx0 = mx1; mx1 = t; }
    if (sy[mn0] > sy[mx1]) { t = mn0; mn0 = mx0; mx0 = t; t = mn1; mn1 = mx1; mx1 = t; }
    if (sy[mn1] > sy[mx0]) { t = mn1; mn1 = mx0; mx0 = t; }
    int max_y = (sy[mx0] + sy[mx1] + 1) >> 1, max_c = (sc[mx0] + sc[mx1] + 1) >> 1;
    int min_y = (sy[mn0] + sy[mn1] + 1) >> 1, min_c = (sc[mn0] + sc[mn1] + 1) >> 1;
    int diff = max_y - min_y;
    if (diff != 0) {
        int diff_c = max_c - min_c;
        int x = ilog2i(diff);
        int norm = ((diff << 4) >> x) & 15;
        x += norm != 0;
        int adc = abs(diff_c);
        int y = adc > 0 ? ilog2i(adc) + 1 : 0;
        int a = diff_c == 0 ? 0 : (diff_c * (c_divsig[norm] | 8) + (1 << (y - 1))) >> y;
        int k;
        if (3 + x - y < 1) {
            k = 1;
            a = a < 0 ? -15 : (a > 0 ? 15 : 0);
        } else {
            k = 3 + x - y;
        }
        pc.a = a;
        pc.k = k;
        pc.b = min_c - ((a * min_y) >> k);
    } else {
        pc.a = 0;
        pc.k = 0;
        pc.b = min_c;
    }
}

// per-task setup of the non-angular modes: picks the reference arrays, DC value, CCLM parameters
__device__ __forceinline__ void pred_setup(const Ctx S, const CtuGeom g, const Node nd, int c, int mode, int lane, PredCtx &pc) {
    WB_SHARED_CTX(S);
    const int cs = c != 0;
    const int n = nd.w >> cs;
    pc.mode = mode; pc.c = c; pc.n = n; pc.l2 = ilog2i(n);
    pc.nscale = (2 * pc.l2 - 2) >> 2; pc.dc = 0;
    pc.a = pc.k = pc.b = 0; pc.cclm128 = false;
    if (mode > 66) {
        pc.kind = 3;
        pc.lf = pc.ab = nullptr;
        cclm_params(S, g, nd, c, mode, pc);
        return;
    }
    const int filt = (c == 0 && n >= 8 && mode == 0) ? 1 : 0;
    pc.lf = S.c->refL[c][filt];
    pc.ab = S.c->refA[c][filt];
    if (mode == MODE_PLANAR) {
        pc.kind = 0;
        return;
    }
    pc.kind = 1;
    int s = 0;
#pragma unroll RU
    for (int i = lane; i < n; i += 32) s += pc.ab[i] + pc.lf[1 + i];
    s = warp_sum(s) + n;
    pc.dc = (s >> (pc.l2 + 1)) & 255;
}

// planar (0), DC (1) incl. PDPC, CCLM (3)
__device__ __forceinline__ int pred_sample(const Ctx S, const PredCtx &pc, int x, int y) {
    const int n = pc.n;
    if (pc.kind == 3) {
        if (pc.cclm128) return 128;
        return clip8(((S.c->pds[y * n + x] * pc.a) >> pc.k) + pc.b);
    }
    const int16_t *lrs = pc.lf + 1, *ars = pc.ab;
    int p;
    if (pc.kind == 0) {
        int pv = (n - 1 - y) * ars[x] + (y + 1) * lrs[n];
        int ph = (n - 1 - x) * lrs[y] + (x + 1) * ars[n];
        p = ((pv + ph + n) >> (pc.l2 + 1)) & 255;
    } else {
        p = pc.dc;
    }
    const int wl = pdpc_w(pc.nscale, x), wt = pdpc_w(pc.nscale, y);
    const int v = (int16_t)(lrs[y] * wl + ars[x] * wt + (64 - wt - wl) * p + 32);
    return clip8(v >> 6);
}

// Angular prediction of one sample straight from the reference line (no per-task projection array): with
// E[j] = above[j-1] (j > 0), corner (j = 0), left[-j] (j < 0), the projected reference of intra_predictor.rs:1398-1414 /
// 1480-1495 is ref[idx] = E[+-f(idx)], f(idx) = min(idx, 2n) for idx >= 0 and -min((idx * invAngle + 256) >> 9, n) below
// (+ for the vertical modes, - for the horizontal ones).  mode, c, n may differ from lane to lane.
__device__ __forceinline__ int ang_sample_direct(const Ctx S, int c, int n, int l2, int mode, int x, int y) {
    const int ang = S.tb->ang[mode], inv = S.tb->invang[mode];
    const bool vertical = mode >= 34;
    const int filt = (c == 0 && n >= 8 && (mode == 2 || mode == 34 || mode == 66)) ? 1 : 0;
    const int16_t *lf = S.c->refL[c][filt], *ab = S.c->refA[c][filt];
    const int16_t *e0 = (filt ? S.c->seqF : S.c->seq[c]) + 2 * n;
    const int t = vertical ? y : x, u = vertical ? x : y;
    const int prod = (t + 1) * ang;
    const int ifact = prod & 31, base = u + (prod >> 5);
    const int sg = vertical ? 1 : -1;
    auto tap = [&](int idx) -> int {  // ang >= 0: idx >= 0 always (no projection from the other side)
        const int f = (ang >= 0 || idx >= 0) ? min(idx, 2 * n) : -min((idx * inv + 256) >> 9, n);
        return e0[sg * f];
    };
    const unsigned ap = S.tb->angp[l2 - 2][mode];  // mode / size dependent switches, tabulated once per CTA (init_tables)
    int p;
    if (c == 0) {
        int f0, f1, f2, f3;
        if (ap & 1u) {
            const int h = ifact >> 1;
            f0 = 16 - h; f1 = 32 - h; f2 = 16 + h; f3 = h;
        } else {
            const int pk = S.tb->fc[ifact];
            f0 = (int8_t)(pk & 255); f1 = (int8_t)((pk >> 8) & 255); f2 = (int8_t)((pk >> 16) & 255); f3 = pk >> 24;
        }
        p = clip8((f0 * tap(base) + f1 * tap(base + 1) + f2 * tap(base + 2) + f3 * tap(base + 3) + 32) >> 6);
    } else if (ifact != 0) {
        p = (((32 - ifact) * tap(base + 1) + ifact * tap(base + 2) + 16) >> 5) & 255;
    } else {
        p = tap(base + 1) & 255;
    }
    if (ap & 8u) {  // PDPC (intra_predictor.rs:355-757)
        const int ns = (ap >> 1) & 3;
        const int16_t *lrs = lf + 1, *ars = ab;
        int refl = 0, reft = 0, wl = 0, wt = 0;
        if (mode == 18 || mode == 50) {
            const int corner = lf[0];
            refl = lrs[y] - corner + p; reft = ars[x] - corner + p;
            if (mode == 50) wl = pdpc_w(ns, x); else wt = pdpc_w(ns, y);
        } else if (mode < 18) {
            if (y < (3 << ns)) reft = ars[x + (((y + 1) * inv + 256) >> 9)];
            wt = pdpc_w(ns, y);
        } else {
            if (x < (3 << ns)) refl = lrs[y + (((x + 1) * inv + 256) >> 9)];
            wl = pdpc_w(ns, x);
        }
        const int v = (int16_t)(refl * wl + reft * wt + (64 - wt - wl) * p + 32);
        p = clip8(v >> 6);
    }
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// angular prediction, four samples per lane (ang4.cuh)
// ---------------------------------------------------------------------------------------------------------------
// per (mode, component, block size) constants; mode may differ from lane to lane
__device__ __forceinline__ a4::Mode a4_setup(const Ctx S, int c, int n, int l2, int mode) {
    a4::Mode m;
    const unsigned ap = S.tb->angp[l2 - 2][mode];  // mode / size dependent switches, tabulated once per CTA (init_tables)
    m.ang = S.tb->ang[mode];
    m.inv = S.tb->invang[mode];
    const int v = ln_variant(c, (c == 0 && n >= 8 && (mode == 2 || mode == 34 || mode == 66)) ? 1 : 0);
    const int vert = mode >= 34;
    m.main = S.c->ln[v][vert ? 0 : 1] + LN_OFF;
    m.side = S.c->ln[v][vert ? 1 : 0] + LN_OFF;
    m.taps = c ? 2 : (int)(ap & 1u);
    m.pdpc = (ap & 8u) ? ((mode == 18 || mode == 50) ? 1 : 2) : 0;
    m.ns = (int)((ap >> 1) & 3u);
    return m;
}

// source samples matching a quad of a (t, u) line: rows of the block for the vertical modes, rows of the TRANSPOSED block for the
// horizontal ones (bx, by: block origin in its component plane)
__device__ __forceinline__ unsigned org_quad(const Ctx S, int c, int mode, int bx, int by, int t, int u0) {
    if (mode >= 34) return *reinterpret_cast<const unsigned *>((c ? S.c->orgC[c - 1] + ((by + t) << 4) : S.c->orgY + ((by + t) << 5)) + bx + u0);
    return *reinterpret_cast<const unsigned *>((c ? S.c->orgCT[c - 1] + ((bx + t) << 4) : S.c->orgYT + ((bx + t) << 5)) + by + u0);
}

// SAD of ONE angular mode over the three components of an 8x8 SINGLE_TREE CU in one pass: lanes 0-15 the 16 quads of the luma
// block, lanes 16-19 / 20-23 the four quads of the Cb / Cr 4x4 blocks.  scr: the warp's scratch line (>= 80 bytes).
__device__ __noinline__ unsigned dir_sad8(const Ctx S, const Node nd, int mode, uint8_t *scr, int lane) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(scr);
    const int c = lane < 16 ? 0 : (lane < 20 ? 1 : 2);
    const int n = c ? 4 : 8, l2 = c ? 2 : 3;
    const int t = c ? (lane & 3) : (lane >> 1), u0 = c ? 0 : (lane & 1) * 4;
    a4::Mode m = a4_setup(S, c, n, l2, mode);
    if (m.ang < 0) {  // uniform: project the side line below index 0 (luma 18 elements, chroma 10 each) into the scratch line
        uint8_t *pl = scr + 12, *pcb = scr + 36, *pcr = scr + 56;  // index 0 of the three projected lines ([-n-3, n+6] each)
        if (lane < 18) {
            const a4::Mode my = a4_setup(S, 0, 8, 3, mode);
            a4::project_elem(pl, my.main, my.side, 8, my.inv, lane);
        } else if (lane < 28) {
            const a4::Mode my = a4_setup(S, 1, 4, 2, mode);
            a4::project_elem(pcb, my.main, my.side, 4, my.inv, lane - 18);
        }
        if (lane < 10) {
            const a4::Mode my = a4_setup(S, 2, 4, 2, mode);
            a4::project_elem(pcr, my.main, my.side, 4, my.inv, lane);
        }
        __syncwarp();
        m.main = c == 0 ? pl : (c == 1 ? pcb : pcr);
    }
    unsigned s = 0;
    if (lane < 24) s = a4::sad4(a4::quad(m, S.tb->taps, t, u0), org_quad(S, c, mode, nd.x >> (c != 0), nd.y >> (c != 0), t, u0), 0u);
    s = warp_sumu(s);
    __syncwarp();  // the scratch line is rewritten by the next mode
    return s;
}

// SADs of up to EIGHT angular modes of a 4x4 luma CU in one pass: four lanes (one quad = one line each) per mode; lanes
// 4k .. 4k+3 evaluate `mode` (per lane; mode < 0: idle).  Returns the SAD of the lane's group.  scr >= 168 bytes.
__device__ __noinline__ unsigned dir_sad4(const Ctx S, const Node nd, int mode, uint8_t *scr, int lane) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(scr);
    const int slot = lane >> 2, t = lane & 3;
    const bool act = mode >= 2;
    a4::Mode m = a4_setup(S, 0, 4, 2, act ? mode : 2);
    if (act && m.ang < 0) {  // per group: 10 projected elements, spread over the group's four lanes
        uint8_t *pr = scr + 7 + 20 * slot;  // index 0; [-7, 10] used
        a4::project_elem(pr, m.main, m.side, 4, m.inv, t);
        a4::project_elem(pr, m.main, m.side, 4, m.inv, t + 4);
        if (t < 2) a4::project_elem(pr, m.main, m.side, 4, m.inv, t + 8);
        m.main = pr;
    }
    __syncwarp();
    unsigned s = 0;
    if (act) s = a4::sad4(a4::quad(m, S.tb->taps, t, 0), org_quad(S, 0, mode, nd.x, nd.y, t, 0), 0u);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    __syncwarp();
    return s;
}

__device__ __forceinline__ unsigned warp_minu(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// The SAD-driven direction search of one CU up to 8x8 (block_splitter.rs:887-973: 13 coarse angular modes, then
// step_search +-2 and +-1 on the summed SAD, first-minimum rules H6).  The SADs are integers below 2^24, so the reference's
// f32 comparisons are integer comparisons.  Leaves dir, v0 (dir-1 valid), v1 (dir+1 valid) in the CTU context.
// 8x8 CU: the coarse modes are split over `nparts` warp tasks, one mode per pass (dir_sad8); each part leaves its first minimum
// as (sad << 4 | index), and the warp that finishes last (shared-memory counter) takes the overall first minimum and runs the
// two refinement steps.  4x4 luma CU (nparts = 1): eight modes per pass (dir_sad4): 13 coarse modes in two passes, each
// refinement step in one.
__device__ __noinline__ void dir_search_part(const Ctx S, const Node nd, int part, int nparts, uint8_t *scr, int lane) {
    WB_SHARED_CTX(S);
    const bool luma_only = nd.tree == DUAL_TREE_LUMA;
    unsigned bp = 0xffffffffu;
    if (luma_only) {
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            const int i = pass * 8 + (lane >> 2);
            const unsigned s = dir_sad4(S, nd, i < 13 ? c_cand15[2 + i] : -1, scr, lane);
            if (i < 13) bp = min(bp, (s << 4) | (unsigned)i);
        }
        bp = warp_minu(bp);
    } else {
        // coarse modes [lo, hi) of this part (nparts = 4: constant divisor)
        const int lo = part * 13 / 4, hi = (part + 1) * 13 / 4;
#pragma unroll 1
        for (int i = lo; i < hi; i++) bp = min(bp, (dir_sad8(S, nd, c_cand15[2 + i], scr, lane) << 4) | (unsigned)i);
        int last = 0;
        if (lane == 0) {
            S.c->dir_part[part] = bp;
            __threadfence_block();
            last = atomicAdd(&S.c->dir_cnt, 1) == nparts - 1;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) return;
        __threadfence_block();
        for (int q = 0; q < nparts; q++) bp = min(bp, *(volatile unsigned *)&S.c->dir_part[q]);
        if (lane == 0) S.c->dir_cnt = 0;
    }
    int cur = c_cand15[2 + (bp & 15u)];
    unsigned cur_cost = bp >> 4;
#pragma unroll 1
    for (int step = 2; step >= 1; step >>= 1) {
        const bool v0 = !(cur < 2 + step), v1 = !(cur + step > 66);
        unsigned c0 = 0xffffffffu, c1 = 0xffffffffu;
        if (luma_only) {  // group 0: cur - step, group 1: cur + step
            const int g = lane >> 2;
            const unsigned s = dir_sad4(S, nd, g == 0 ? (v0 ? cur - step : -1) : (g == 1 ? (v1 ? cur + step : -1) : -1), scr, lane);
            const unsigned s0 = __shfl_sync(0xffffffffu, s, 0), s1 = __shfl_sync(0xffffffffu, s, 4);
            if (v0) c0 = s0;
            if (v1) c1 = s1;
        } else {
            if (v0) c0 = dir_sad8(S, nd, cur - step, scr, lane);
            if (v1) c1 = dir_sad8(S, nd, cur + step, scr, lane);
        }
        const unsigned mn = min(min(cur_cost, c0), c1);
        if (cur_cost == mn) {
        } else if (c0 == mn) { cur -= step; cur_cost = c0; }
        else { cur += step; cur_cost = c1; }
    }
    if (lane == 0) {
        S.c->dir = cur;
        S.c->v0 = !(cur < 3);
        S.c->v1 = !(cur + 1 > 66);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// transforms (transformer.rs:2040-2378 forward, 2380-2737 inverse), one warp per TB, direct matrix multiply
// ---------------------------------------------------------------------------------------------------------------
// Both passes multiply 16-bit samples by 8-bit matrix entries and run on the packed dot-product instruction (dp2a: two
// 16x8-bit products per instruction, exact 32-bit accumulation): every item produces two outputs, loads four samples and
// four coefficients per 32/64-bit shared-memory access, and so spends about one instruction per multiply-add.
//
// Row pass: out[y][i] = (sum_x M[x][i] * in[y][x] + rnd) >> sh for the row pair (2yp, 2yp+1); Q[x4*n + i] = bytes M[4*x4+0..3][i].
// PACK: the two results are stored as one word P[yp*n + i] = (row 2yp | row 2yp+1 << 16), the layout the column pass reads.
template <bool PACK>
__device__ __noinline__ void mm_rows_q(const int32_t *Q, const int16_t *in, void *out, int n, int l2, int rnd, int sh, int lane) {
    WB_SHARED_PTR(Q); WB_SHARED_PTR(in); WB_SHARED_PTR(out);
    const int items = (n >> 1) << l2, nq = n >> 2;
#pragma unroll 1
    for (int o = lane; o < items; o += 32) {
        const int yp = o >> l2, i = o & (n - 1);
        const int2 *r0 = reinterpret_cast<const int2 *>(in + (2 * yp) * n), *r1 = reinterpret_cast<const int2 *>(in + (2 * yp + 1) * n);
        const int32_t *q = Q + i;
        int s0 = 0, s1 = 0;
#pragma unroll U2
        for (int x4 = 0; x4 < nq; x4++) {
            const int w = q[x4 * n];
            const int2 a = r0[x4], b = r1[x4];
            s0 = __dp2a_lo(a.x, w, s0); s0 = __dp2a_hi(a.y, w, s0);
            s1 = __dp2a_lo(b.x, w, s1); s1 = __dp2a_hi(b.y, w, s1);
        }
        s0 = (s0 + rnd) >> sh; s1 = (s1 + rnd) >> sh;
        if (PACK) reinterpret_cast<int32_t *>(out)[yp * n + i] = (s0 & 0xffff) | (s1 << 16);
        else {
            reinterpret_cast<int16_t *>(out)[(2 * yp) * n + i] = (int16_t)s0;
            reinterpret_cast<int16_t *>(out)[(2 * yp + 1) * n + i] = (int16_t)s1;
        }
    }
}
// Column pass: out[i][x] = (sum_y M[i][y] * in[y][x] + rnd) >> sh for the output row pair (2ip, 2ip+1); M row-major int8,
// in pair-packed P[(y/2)*n + x] = (in[y][x] | in[y+1][x] << 16).
__device__ __noinline__ void mm_cols_q(const int8_t *M, const int32_t *P, int16_t *out, int n, int l2, int rnd, int sh, bool clamp16, int lane) {
    WB_SHARED_PTR(M); WB_SHARED_PTR(P); WB_SHARED_PTR(out);
    const int items = (n >> 1) << l2, nq = n >> 2;
#pragma unroll 1
    for (int o = lane; o < items; o += 32) {
        const int ip = o >> l2, x = o & (n - 1);
        const int32_t *m0 = reinterpret_cast<const int32_t *>(M + (2 * ip) * n), *m1 = reinterpret_cast<const int32_t *>(M + (2 * ip + 1) * n);
        const int32_t *pp = P + x;
        int s0 = 0, s1 = 0;
#pragma unroll U2
        for (int y4 = 0; y4 < nq; y4++) {
            const int w0 = m0[y4], w1 = m1[y4];
            const int p0 = pp[(2 * y4) * n], p1 = pp[(2 * y4 + 1) * n];
            s0 = __dp2a_lo(p0, w0, s0); s0 = __dp2a_hi(p1, w0, s0);
            s1 = __dp2a_lo(p0, w1, s1); s1 = __dp2a_hi(p1, w1, s1);
        }
        s0 = (s0 + rnd) >> sh; s1 = (s1 + rnd) >> sh;
        if (clamp16) { s0 = min(32767, max(-32768, s0)); s1 = min(32767, max(-32768, s1)); }
        out[(2 * ip) * n + x] = (int16_t)s0;
        out[(2 * ip + 1) * n + x] = (int16_t)s1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dependent quantisation (quantizer.rs:338-517 search_dq + 686-721 walk) and the rate walk (block_splitter.rs:415-460)
// ---------------------------------------------------------------------------------------------------------------
#if WB_TAB_N >= 1024
#define WB_LDQ(i) (S.tb->ldq[min((i), 1023)])
#define WB_LV(i) (S.tb->lv[min((i), 1023)])
#else
#define WB_LDQ(i) ((i) < WB_TAB_N ? S.tb->ldq[(i)] : __ldg(&tab->ldq[min((i), 1023)]))
#define WB_LV(i) ((i) < WB_TAB_N ? S.tb->lv[(i)] : __ldg(&tab->lv[min((i), 1023)]))
#endif

// Next-state maps of the walk are kept as 4 bytes (byte s = next state of state s), so that composing two maps is one byte
// permute: the selector of __byte_perm is the first map in nibble form.
constexpr unsigned MAP_ID = 0x03020100u;
// s / ls for s < 2^25 (|tc| << sh plus the rounding offset) without a division sequence: the estimate by the tabulated
// reciprocal is floor(s / ls) or one more; one multiply-subtract decides.
__device__ __forceinline__ unsigned div_ls(const Ctx S, unsigned s, unsigned ls) {
    unsigned q = __umulhi(s, S.tb->ls_recip);
    if ((int)(s - q * ls) < 0) q--;
    return q;
}
__device__ __forceinline__ unsigned map_nib(unsigned m) { return __byte_perm(m | (m >> 4), 0u, 0x4420u); }  // bytes -> nibbles
__device__ __forceinline__ unsigned map_compose(unsigned a, unsigned b) { return __byte_perm(b, 0u, map_nib(a)); }  // apply a first, then b

// Local candidate costs of one scan position (quantizer.rs:436-503): for delta = (state > 1) in {0,1} the two candidate
// levels a0 = (x + delta) / 2 and a1 = a0 + 1, cost_i = 128 * |tc - dequant(q_i)| + lambda * dq_table[bits_i].
struct LC {
    int L00, L10, L01, L11;  // [candidate][delta]
    int L0s0;                // candidate 0 as seen by state 0 (first-visit is_trailing_zeros flag, H2)
    unsigned pk;             // bit0: parity of a0 (delta 0), bit1: parity of a0 (delta 1), bit2: sub-block-start adjustment applies to state 0
};
constexpr int TR_INF = 1 << 28;

#ifndef WB_LC_NOINLINE
#define WB_LC_NOINLINE 0
#endif
#if WB_LC_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
LC local_costs(const Ctx S, const DevTables *__restrict__ tab, int tc, unsigned w, int k, int kstar, int ls, int sh, int off, int ldq1) {
    LC r;
    const bool flagged = k > kstar;
    if (w & 2048u) {
        const unsigned x = w & 2047u;
        int a0 = (int)(x >> 1);
        int q0 = 2 * a0, q1 = 2 * a0 + 2;
        if (tc < 0) { q0 = -q0; q1 = -q1; }
        int d0 = abs(tc - ((q0 * ls + off) >> sh));
        int d1 = abs(tc - ((q1 * ls + off) >> sh));
        r.L00 = 128 * d0 + WB_LDQ(a0 + 1);
        r.L10 = 128 * d1 + WB_LDQ(a0 + 2);
        r.L0s0 = (flagged && a0 == 0) ? 128 * d0 : r.L00;
        r.pk = a0 & 1;
        a0 = (int)((x + 1) >> 1);
        q0 = a0 > 0 ? 2 * a0 - 1 : 0;
        q1 = 2 * a0 + 1;
        if (tc < 0) { q0 = -q0; q1 = -q1; }
        d0 = abs(tc - ((q0 * ls + off) >> sh));
        d1 = abs(tc - ((q1 * ls + off) >> sh));
        r.L01 = 128 * d0 + WB_LDQ(a0 + 1);
        r.L11 = 128 * d1 + WB_LDQ(a0 + 2);
        r.pk |= (a0 & 1) << 1;
    } else {
        r.L00 = r.L01 = ldq1;
        r.L10 = r.L11 = TR_INF;
        r.L0s0 = flagged ? 0 : ldq1;
        r.pk = 0;
    }
    if (flagged && (k & 15) == 0) r.pk |= 4;
    return r;
}

// One trellis step on the 4 state costs (states 0,1 move to {0,2}, states 2,3 to {1,3}; encoder_context.rs:339).
// Returns the 4 decision bits (1 = candidate a1; ties keep a0, quantizer.rs:505).
__device__ __forceinline__ unsigned vstep(const LC &l, int ldq1, int &C0, int &C1, int &C2, int &C3) {
    int X = (l.pk & 1) ? C2 : C0, Y = (l.pk & 1) ? C0 : C2;
    int c0 = l.L0s0 + X, c1 = l.L10 + Y;
    const bool d_0 = c1 < c0;
    int n0 = d_0 ? c1 : c0;
    if ((l.pk & 4) && !d_0) n0 -= ldq1;  // quantizer.rs:512-514, applied after the comparison
    c0 = l.L00 + Y; c1 = l.L10 + X;
    const bool d_1 = c1 < c0;
    const int n1 = d_1 ? c1 : c0;
    X = (l.pk & 2) ? C3 : C1; Y = (l.pk & 2) ? C1 : C3;
    c0 = l.L01 + X; c1 = l.L11 + Y;
    const bool d_2 = c1 < c0;
    const int n2 = d_2 ? c1 : c0;
    c0 = l.L01 + Y; c1 = l.L11 + X;
    const bool d_3 = c1 < c0;
    const int n3 = d_3 ? c1 : c0;
    C0 = n0; C1 = n1; C2 = n2; C3 = n3;
    return (unsigned)d_0 | ((unsigned)d_1 << 1) | ((unsigned)d_2 << 2) | ((unsigned)d_3 << 3);
}

// next-state map of one position for the walk: next(s) = 2 * (parity(a_s) ^ (s & 1)) + (s >> 1), a_s = a0(delta(s)) + decision bit s.
// pos_q: the 4 bits parity(a_s) ^ (s & 1), spread to the bytes by a multiplication; pos_map_nib: the map as a permute selector.
__device__ __forceinline__ unsigned pos_q(unsigned pk, unsigned dec) { return ((((pk & 1u) * 3u) | ((pk & 2u) * 6u)) ^ dec ^ 0xAu) & 0xFu; }
__device__ __forceinline__ unsigned pos_map(unsigned pk, unsigned dec, bool nz) {
    if (!nz) return 0x03010200u;  // a = 0 for every state: 0->0, 1->2, 2->1, 3->3
    return (((pos_q(pk, dec) * 0x00204081u) & 0x01010101u) << 1) + 0x01010000u;
}
__device__ __forceinline__ unsigned pos_map_nib(unsigned pk, unsigned dec, bool nz) { return map_nib(pos_map(pk, dec, nz)); }

// coef (raster, n x n) -> lev (raster).  Returns rate (sum of lv[] per block_splitter.rs:415-460) and whether any level != 0.
//
// The reference's memoised DFS (quantizer.rs:338-517) equals a backward Viterbi pass from DC upwards in which the entries
// (k, state 0) for k above the highest position k* whose state-0 lower candidate is non-zero carry the first-visit
// is_trailing_zeros flag.  Every step is (min,+)-linear in the 4 state costs except the post-comparison
// adjustment at sub-block starts, so the pass is split into chunks: (B) each lane folds the steps of its chunk into a 4x4
// (min,+) matrix, (C) the chunk heads (which carry the non-linear step) and matrices are applied in sequence, one chunk per
// iteration, (D) each lane replays its chunk from its true entry costs to record the decisions, (F) the walk from the last
// position (quantizer.rs:686-721) is a prefix scan over per-chunk next-state maps followed by a per-lane walk.  Costs are
// int32 relative to the running minimum; the decisions are identical to the reference's i64 comparisons.
// Why int32 is enough: a step adds at most 128 * |tc - dequant(q)| + ldq[bits] with |tc - dequant| <= 2 quantiser steps + 1
// < 2^15.2 (x is the clamped quotient, so the candidates bracket tc) and ldq <= 2^26 (HostConsts::init rejects larger tables;
// the default tuning stays below 2^24.5 at QP 63); every state reaches every other within two steps, so after the per-chunk
// renormalisation the four state costs are at most 2 steps apart (< 2^27.1), and a chunk of 16 steps adds < 2^30.1 even at the
// theoretical bound (< 2^26 with the default tables): all sums stay below 2^31.  TR_INF = 2^28 marks the absent candidate a1 of
// a zero coefficient: it only has to exceed cost differences (<= 3 steps), never absolute path costs.
__device__ __noinline__ void trellis(const Ctx S, const DevTables *__restrict__ tab, const int16_t *coef, int l2, uint16_t *Wd, int16_t *lev, int lane,
                        int &rate_out, bool &any_out) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(coef); WB_SHARED_PTR(Wd); WB_SHARED_PTR(lev);
    const int n = 1 << l2, nn = n * n, sh = l2 + 4, off = 1 << (sh - 1);
    const int ls = tab->ls;
    const uint16_t *scan = S.tb->scan + tab_off(l2);
    const int ldq1 = S.tb->ldq[1];
    // ---- A: x = S / ls per position; k* (H2)
    int kstar = -1;
    bool anytc = false;
#pragma unroll 1
    for (int k = lane; k < nn; k += 32) {
        int tc = coef[scan[k]];
        unsigned x = 0, nz = tc != 0;
        if (nz) {
            unsigned s = tc > 0 ? ((unsigned)tc << sh) - (unsigned)off : ((unsigned)(-tc) << sh) + (unsigned)off;
            x = min(div_ls(S, s, (unsigned)ls), 2047u);
            anytc = true;
            if (x >= 2) kstar = k;
        }
        Wd[k] = (uint16_t)(x | (nz << 11));
    }
    kstar = warp_max(kstar);
    anytc = __any_sync(0xffffffffu, anytc);
    __syncwarp();
    if (!anytc) {  // every candidate level is 0 and every position is a trailing zero: levels 0, rate 0
#pragma unroll RU
        for (int k = lane; k < nn; k += 32) lev[k] = 0;
        rate_out = 0;
        any_out = false;
        __syncwarp();
        return;
    }
    // ---- DC leaf (quantizer.rs:367-409), computed uniformly
    int Lf0, Lf1, Lf2, Lf3;
    unsigned leafdec = 0;
    {
        const int tc = coef[0];
        const unsigned x = Wd[0] & 2047u;
        int Cs[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const bool itz = (s == 0) && (kstar < 0);
            int cost;
            if (tc == 0) {
                cost = itz ? -ldq1 : ldq1;
            } else {
                const int delta = s > 1;
                int a0 = (int)(x >> 1);
                int q0 = (int)(int16_t)(2 * a0 - delta);  // H3: usize wrap gives -1 for a0 == 0, delta == 1
                if (tc < 0) q0 = -q0;
                int d0 = abs(tc - ((q0 * ls + off) >> sh));
                int bits0 = (a0 != 0 || !itz) ? a0 + 1 : 0;
                int cost0 = 128 * d0 + WB_LDQ(bits0);
                int a1 = a0 + 1;
                int q1 = 2 * a1 - delta;
                if (tc < 0) q1 = -q1;
                int d1 = abs(tc - ((q1 * ls + off) >> sh));
                int cost1 = 128 * d1 + WB_LDQ(a1 + 1);
                if (cost0 <= cost1) {
                    cost = cost0;
                    if (itz && a0 == 0) cost -= ldq1;
                } else {
                    cost = cost1;
                    leafdec |= 1u << s;
                }
            }
            Cs[s] = cost;
        }
        Lf0 = Cs[0]; Lf1 = Cs[1]; Lf2 = Cs[2]; Lf3 = Cs[3];
    }
    const bool tiny = nn == 16;  // 4x4 TBs (the most numerous): one position per lane, plain sequential pass over 15 steps
    const int csl = tiny ? 0 : (nn >= 1024 ? 4 : (nn >= 256 ? 3 : 2)), CS = 1 << csl, nch = nn >> csl, rounds = (nch + 31) >> 5;
    int carry0 = 0, carry1 = 0, carry2 = 0, carry3 = 0;
    unsigned cm0 = MAP_ID, cm1 = MAP_ID;  // walk map of this lane's chunk in round 0 / 1
    if (tiny) {
        LC l;
        l.L00 = l.L01 = l.L0s0 = 0; l.L10 = l.L11 = TR_INF; l.pk = 0;
        unsigned w = 0;
        if (lane < 16) {
            w = Wd[lane];
            if (lane > 0) l = local_costs(S, tab, coef[scan[lane]], w, lane, kstar, ls, sh, off, ldq1);
        }
        int C0 = Lf0, C1 = Lf1, C2 = Lf2, C3 = Lf3;
        unsigned mydec = leafdec;
#pragma unroll 1
        for (int j = 1; j < 16; j++) {
            LC b;
            b.L00 = __shfl_sync(0xffffffffu, l.L00, j); b.L10 = __shfl_sync(0xffffffffu, l.L10, j);
            b.L01 = __shfl_sync(0xffffffffu, l.L01, j); b.L11 = __shfl_sync(0xffffffffu, l.L11, j);
            b.L0s0 = __shfl_sync(0xffffffffu, l.L0s0, j); b.pk = __shfl_sync(0xffffffffu, l.pk, j);
            const unsigned dec = vstep(b, ldq1, C0, C1, C2, C3);
            if (lane == j) mydec = dec;
        }
        if (lane < 16) {
            Wd[lane] = (uint16_t)(w | (mydec << 12));
            const unsigned pk = lane == 0 ? ((w >> 1) & 1u) * 3u : l.pk;
            cm0 = pos_map(pk, mydec, (w & 2048u) != 0);
        }
    } else
    for (int r = 0; r < rounds; r++) {
        const int c = r * 32 + lane;
        const bool vc = c < nch;
        const int k0 = c * CS;
        // ---- B: chunk head costs and the (min,+) matrix of the remaining CS-1 steps
        LC head;
        head.L00 = head.L01 = head.L0s0 = 0; head.L10 = head.L11 = TR_INF; head.pk = 0;
        int M[4][4];
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
            for (int t = 0; t < 4; t++) M[s][t] = s == t ? 0 : TR_INF;
        bool sep_head = false;
        if (vc) {
#pragma unroll 1
            for (int i = 0; i < CS; i++) {
                const int k = k0 + i;
                const LC l = local_costs(S, tab, coef[scan[k]], Wd[k], k, kstar, ls, sh, off, ldq1);
                if (i == 0) {
                    head = l;
                    // the chunk's first step stays separate only where it is not (min,+)-linear: the DC leaf and the
                    // sub-block starts that carry the post-comparison adjustment; everywhere else it is folded into M
                    sep_head = k == 0 || (l.pk & 4u);
                    if (sep_head) continue;
                }
                // new state 0 = min(a00 + old 0, a02 + old 2), state 1 = min(a10 + old 0, a12 + old 2), states 2, 3 from old 1, 3:
                // which candidate each old state feeds depends on the parity of a0 (selected once per position, not per column)
                const bool p0 = l.pk & 1, p1 = l.pk & 2;
                const int a00 = p0 ? l.L10 : l.L0s0, a02 = p0 ? l.L0s0 : l.L10, a10 = p0 ? l.L00 : l.L10, a12 = p0 ? l.L10 : l.L00;
                const int a21 = p1 ? l.L11 : l.L01, a23 = p1 ? l.L01 : l.L11, a31 = p1 ? l.L01 : l.L11, a33 = p1 ? l.L11 : l.L01;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int n0 = min(a00 + M[0][t], a02 + M[2][t]), n1 = min(a10 + M[0][t], a12 + M[2][t]);
                    const int n2 = min(a21 + M[1][t], a23 + M[3][t]), n3 = min(a31 + M[1][t], a33 + M[3][t]);
                    M[0][t] = n0; M[1][t] = n1; M[2][t] = n2; M[3][t] = n3;
                }
            }
        }
#ifndef WB_PAIR16
#define WB_PAIR16 0  // 1: 16x16 TBs: the matrices of the two chunks of a sub-block are combined before the serial pass (16 steps instead of 32); bit-exact, measured 1.2 % slower
#endif
        // ---- B2 (16x16 TBs: 32 chunks of 8, the sub-block starts - the only steps that may be non-linear - are the even chunks):
        //      the even lanes combine their matrix with their right neighbour's, the serial pass then runs over pairs
        const bool pair = WB_PAIR16 && nn == 256;
        int A0[4][4];  // the chunk's own matrix, for the way back down
        if (pair) {
            int N[4][4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int b0 = __shfl_down_sync(0xffffffffu, M[s][0], 1), b1 = __shfl_down_sync(0xffffffffu, M[s][1], 1);
                const int b2 = __shfl_down_sync(0xffffffffu, M[s][2], 1), b3 = __shfl_down_sync(0xffffffffu, M[s][3], 1);
#pragma unroll
                for (int t = 0; t < 4; t++) N[s][t] = min(min(b0 + M[0][t], b1 + M[1][t]), min(b2 + M[2][t], b3 + M[3][t]));
            }
#pragma unroll
            for (int s = 0; s < 4; s++)
#pragma unroll
                for (int t = 0; t < 4; t++) { A0[s][t] = M[s][t]; M[s][t] = N[s][t]; }
        }
        const int cstep = pair ? 2 : 1;
        // ---- C: apply head + matrix chunk (pair) by chunk (pair); lane j owns chunk r*32 + j
        int O0 = 0, O1 = 0, O2 = 0, O3 = 0, my0 = 0, my1 = 0, my2 = 0, my3 = 0;
        int X0 = 0, X1 = 0, X2 = 0, X3 = 0;  // what the owner's matrix applies to (= its entry costs after its separate head, if any)
        const int cnt = min(32, nch - r * 32);
        const unsigned sepmask = __ballot_sync(0xffffffffu, sep_head);
#pragma unroll 1
        for (int j = 0; j < cnt; j += cstep) {
            int I0, I1, I2, I3;
            if (j == 0) { I0 = carry0; I1 = carry1; I2 = carry2; I3 = carry3; }
            else {
                I0 = __shfl_sync(0xffffffffu, O0, j - cstep); I1 = __shfl_sync(0xffffffffu, O1, j - cstep);
                I2 = __shfl_sync(0xffffffffu, O2, j - cstep); I3 = __shfl_sync(0xffffffffu, O3, j - cstep);
            }
            int H0 = I0, H1 = I1, H2 = I2, H3 = I3;
            if ((sepmask >> j) & 1u) {  // uniform: chunk j has a separate head (every lane applies its own, only lane j's result is used)
                if (r == 0 && j == 0) { H0 = Lf0; H1 = Lf1; H2 = Lf2; H3 = Lf3; }
                else vstep(head, ldq1, H0, H1, H2, H3);
            }
            int T0 = min(min(M[0][0] + H0, M[0][1] + H1), min(M[0][2] + H2, M[0][3] + H3));
            int T1 = min(min(M[1][0] + H0, M[1][1] + H1), min(M[1][2] + H2, M[1][3] + H3));
            int T2 = min(min(M[2][0] + H0, M[2][1] + H1), min(M[2][2] + H2, M[2][3] + H3));
            int T3 = min(min(M[3][0] + H0, M[3][1] + H1), min(M[3][2] + H2, M[3][3] + H3));
            const int mn = min(min(T0, T1), min(T2, T3));
            O0 = T0 - mn; O1 = T1 - mn; O2 = T2 - mn; O3 = T3 - mn;
            if (lane == j) { my0 = I0; my1 = I1; my2 = I2; my3 = I3; X0 = H0; X1 = H1; X2 = H2; X3 = H3; }
        }
        carry0 = __shfl_sync(0xffffffffu, O0, cnt - cstep); carry1 = __shfl_sync(0xffffffffu, O1, cnt - cstep);
        carry2 = __shfl_sync(0xffffffffu, O2, cnt - cstep); carry3 = __shfl_sync(0xffffffffu, O3, cnt - cstep);
        if (pair) {  // the way back down: entry costs of the odd chunk = the even chunk's own matrix applied to what it started from
            int E0 = min(min(A0[0][0] + X0, A0[0][1] + X1), min(A0[0][2] + X2, A0[0][3] + X3));
            int E1 = min(min(A0[1][0] + X0, A0[1][1] + X1), min(A0[1][2] + X2, A0[1][3] + X3));
            int E2 = min(min(A0[2][0] + X0, A0[2][1] + X1), min(A0[2][2] + X2, A0[2][3] + X3));
            int E3 = min(min(A0[3][0] + X0, A0[3][1] + X1), min(A0[3][2] + X2, A0[3][3] + X3));
            const int mn = min(min(E0, E1), min(E2, E3));
            E0 = __shfl_up_sync(0xffffffffu, E0 - mn, 1); E1 = __shfl_up_sync(0xffffffffu, E1 - mn, 1);
            E2 = __shfl_up_sync(0xffffffffu, E2 - mn, 1); E3 = __shfl_up_sync(0xffffffffu, E3 - mn, 1);
            if (lane & 1) { my0 = E0; my1 = E1; my2 = E2; my3 = E3; }
        }
        // ---- D: replay the chunk from its true entry costs, record decisions and the chunk's walk map
        if (vc) {
            int C0 = my0, C1 = my1, C2 = my2, C3 = my3;
            unsigned cm = MAP_ID;
#pragma unroll 1
            for (int i = 0; i < CS; i++) {
                const int k = k0 + i;
                const unsigned w = Wd[k];
                unsigned dec, pk;
                if (k == 0) {
                    C0 = Lf0; C1 = Lf1; C2 = Lf2; C3 = Lf3;
                    dec = leafdec;
                    pk = ((w >> 1) & 1u) * 3u;  // DC: a0 = x / 2 for both deltas
                } else {
                    const LC l = local_costs(S, tab, coef[scan[k]], w, k, kstar, ls, sh, off, ldq1);
                    dec = vstep(l, ldq1, C0, C1, C2, C3);
                    pk = l.pk;
                }
                Wd[k] = (uint16_t)(w | (dec << 12));
                cm = __byte_perm(cm, 0u, pos_map_nib(pk, dec, (w & 2048u) != 0));  // this position first, then the ones below
            }
            if (r == 0) cm0 = cm; else cm1 = cm;
        }
    }
    __syncwarp();
    // ---- F: walk from the last scan position with state 0 (quantizer.rs:686-721) + rate (block_splitter.rs:415-460)
    unsigned state = 0;
    bool seen_nz = false;
    int rate = 0;
    const int lv0 = S.tb->lv[0];
    for (int r = rounds - 1; r >= 0; r--) {
        const int c = r * 32 + lane;
        const bool vc = c < nch;
        const int k0 = c * CS;
        unsigned inc = vc ? (r == 0 ? cm0 : cm1) : MAP_ID;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_down_sync(0xffffffffu, inc, d);
            if (lane + d < 32) inc = map_compose(t, inc);
        }
        const unsigned exc = __shfl_down_sync(0xffffffffu, inc, 1);
        unsigned s = (lane == 31) ? state : ((exc >> (8 * state)) & 3u);
        const unsigned all = __shfl_sync(0xffffffffu, inc, 0);
        state = (all >> (8 * state)) & 3u;
        int lead = 0, irate = 0;
        bool has = false;
        if (vc) {
#pragma unroll 1
            for (int i = CS - 1; i >= 0; i--) {
                const int k = k0 + i;
                const unsigned w = Wd[k];
                const unsigned x = w & 2047u, dec = w >> 12;
                const int delta = s > 1;
                int q = 0;
                unsigned a = 0;
                if (w & 2048u) {
                    a = ((k == 0) ? (x >> 1) : ((x + delta) >> 1)) + ((dec >> s) & 1u);
                    if (k == 0) q = (int)(int16_t)(2 * (int)a - delta);
                    else q = a > 0 ? 2 * (int)a - delta : 0;
                    if (coef[scan[k]] < 0) q = -q;
                }
                lev[scan[k]] = (int16_t)q;
                if (q != 0) {
                    irate += WB_LV((abs(q) + delta) >> 1);
                    has = true;
                } else if (has) irate += lv0;
                else lead++;
                s = 2u * ((a & 1u) ^ (s & 1u)) + (s >> 1);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (vc) {
            rate += irate;
            if (seen_nz || (lane < 31 && (bal >> (lane + 1)) != 0)) rate += lead * lv0;
        }
        seen_nz = seen_nz || bal != 0;
    }
    rate_out = warp_sum(rate);
    any_out = seen_nz;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Dependent quantisation of ONE 8x8 TB (128 of the 165 TBs >= 8x8 of a CTU).  Same arithmetic as trellis(), organised the
// other way round: instead of cutting the TB into chunks and paying for the (min,+) matrices, the chunk-by-chunk pass and the
// replay, the backward Viterbi pass runs as one sequential chain with one trellis STATE per lane (lanes 0-3) over local costs
// that all 32 lanes tabulated beforehand (one 32-byte row per position: the costs of candidate a0 / a1 as each of the four
// states sees them).  A chain step is two shuffles (the two predecessor states), one 8-byte table read and ~8 integer instructions; the whole routine is
// about a third of trellis() in code and in executed instructions.
//   coef: coefficients (raster); Wd: 64 x / non-zero words; lc: 1 kB cost table (may overlap lev: it is dead before the walk)
//   lev (out): levels (raster).  Cost bound: see trellis(); the chain is renormalised at every sub-block start.
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void trellis8_chain(const Ctx S, const DevTables *__restrict__ tab, const int16_t *coef, uint16_t *Wd, int16_t *lev, int4 *lc, int lane,
                                            int &rate_out, bool &any_out) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(coef); WB_SHARED_PTR(Wd); WB_SHARED_PTR(lev); WB_SHARED_PTR(lc);
    constexpr int sh = 7, off = 64;
    const int ls = tab->ls, ldq1 = S.tb->ldq[1];
    const uint16_t *scan = S.tb->scan + tab_off(3);
    const int sc_lo = scan[lane], sc_hi = scan[lane + 32];  // this lane's two scan positions: lane and lane + 32
    // ---- A: x = S / ls, k* (H2), local costs of every position -> table, parity / flag masks of the chain
    const int tcl = coef[sc_lo], tch = coef[sc_hi];
    unsigned xl = 0, xh = 0;
    if (tcl != 0) xl = min(div_ls(S, tcl > 0 ? ((unsigned)tcl << sh) - (unsigned)off : ((unsigned)(-tcl) << sh) + (unsigned)off, (unsigned)ls), 2047u);
    if (tch != 0) xh = min(div_ls(S, tch > 0 ? ((unsigned)tch << sh) - (unsigned)off : ((unsigned)(-tch) << sh) + (unsigned)off, (unsigned)ls), 2047u);
    if (!__any_sync(0xffffffffu, (tcl | tch) != 0)) {  // every level is 0, rate 0 (as in trellis())
        reinterpret_cast<uint32_t *>(lev)[lane] = 0u;
        rate_out = 0;
        any_out = false;
        __syncwarp();
        return;
    }
    const int kstar = warp_max(xh >= 2 ? lane + 32 : (xl >= 2 ? lane : -1));
    const unsigned wl = xl | ((unsigned)(tcl != 0) << 11), wh = xh | ((unsigned)(tch != 0) << 11);
    Wd[lane] = (uint16_t)wl;
    Wd[lane + 32] = (uint16_t)wh;
    unsigned p1l, p2l, p1h, p2h;
    {
        const LC ll = local_costs(S, tab, tcl, wl, lane, kstar, ls, sh, off, ldq1);
        lc[2 * lane] = make_int4(ll.L0s0, ll.L10, ll.L00, ll.L10);  // two 16-byte halves per position: (a0, a1) costs as seen by states 0 | 1 and 2 | 3
        lc[2 * lane + 1] = make_int4(ll.L01, ll.L11, ll.L01, ll.L11);
        p1l = __ballot_sync(0xffffffffu, ll.pk & 1u); p2l = __ballot_sync(0xffffffffu, ll.pk & 2u);
    }
    {
        const LC lh = local_costs(S, tab, tch, wh, lane + 32, kstar, ls, sh, off, ldq1);
        lc[2 * (lane + 32)] = make_int4(lh.L0s0, lh.L10, lh.L00, lh.L10);
        lc[2 * (lane + 32) + 1] = make_int4(lh.L01, lh.L11, lh.L01, lh.L11);
        p1h = __ballot_sync(0xffffffffu, lh.pk & 1u); p2h = __ballot_sync(0xffffffffu, lh.pk & 2u);
    }
    const int tc0 = __shfl_sync(0xffffffffu, tcl, 0);
    const unsigned x0 = __shfl_sync(0xffffffffu, xl, 0);
    __syncwarp();
    // ---- B: the chain, lane s = state s
    unsigned dlo = 0, dhi = 0;
    if (lane < 4) {
        const int s = lane;
        const unsigned inv = (s & 1) ? 0xffffffffu : 0u;
        const unsigned msw_lo = (s < 2 ? p1l : p2l) ^ inv, msw_hi = (s < 2 ? p1h : p2h) ^ inv;
        const unsigned adj = s == 0 ? ((unsigned)(16 > kstar) | ((unsigned)(32 > kstar) << 1) | ((unsigned)(48 > kstar) << 2)) : 0u;  // quantizer.rs:512-514 at k = 16, 32, 48
        const char *lp = reinterpret_cast<const char *>(lc) + s * 8;  // this state's (a0, a1) pair inside a position's 32-byte row
        int C;
        {   // DC leaf (quantizer.rs:367-409) for this lane's state
            const bool itz = (s == 0) && (kstar < 0);
            if (tc0 == 0) {
                C = itz ? -ldq1 : ldq1;
            } else {
                const int delta = s > 1;
                const int A0 = (int)(x0 >> 1);
                int q0 = (int)(int16_t)(2 * A0 - delta);  // H3: usize wrap gives -1 for a0 == 0, delta == 1
                if (tc0 < 0) q0 = -q0;
                const int d0 = abs(tc0 - ((q0 * ls + off) >> sh));
                const int bits0 = (A0 != 0 || !itz) ? A0 + 1 : 0;
                const int cost0 = 128 * d0 + WB_LDQ(bits0);
                const int A1 = A0 + 1;
                int q1 = 2 * A1 - delta;
                if (tc0 < 0) q1 = -q1;
                const int d1 = abs(tc0 - ((q1 * ls + off) >> sh));
                const int cost1 = 128 * d1 + WB_LDQ(A1 + 1);
                if (cost0 <= cost1) {
                    C = cost0;
                    if (itz && A0 == 0) C -= ldq1;
                } else {
                    C = cost1;
                    dlo = 1u;
                }
            }
        }
        // states 0,1 continue from {0,2}, states 2,3 from {1,3}; which of the two feeds candidate a0 depends on the parity of a0
        const int srcP = s >> 1, srcQ = srcP + 2;
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const unsigned msw = h ? msw_hi : msw_lo;
            unsigned dec = 0;
            const char *lph = lp + h * 1024;
#pragma unroll 1
            for (int q4 = 0; q4 < 2; q4++) {  // one sub-block of 16 positions per iteration
                const int jb = q4 * 16;
                int j0 = jb + 1;
                if (h | q4) {  // sub-block start: renormalise, then the step with the post-comparison adjustment
                    int mn = min(C, __shfl_xor_sync(0xFu, C, 1));
                    mn = min(mn, __shfl_xor_sync(0xFu, mn, 2));
                    C -= mn;
                    const int2 L = *reinterpret_cast<const int2 *>(lph + jb * 32);
                    const int P = __shfl_sync(0xFu, C, srcP), Q = __shfl_sync(0xFu, C, srcQ);
                    const bool sw = (msw >> jb) & 1u;
                    const int c0 = L.x + (sw ? Q : P), c1 = L.y + (sw ? P : Q);
                    const bool d = c1 < c0;  // ties keep a0 (quantizer.rs:505)
                    C = d ? c1 : c0;
                    if (((adj >> (2 * h + q4 - 1)) & 1u) && !d) C -= ldq1;
                    dec |= (unsigned)d << jb;
                }
#pragma unroll U8C
                for (int j = j0; j < jb + 16; j++) {
                    const int2 L = *reinterpret_cast<const int2 *>(lph + j * 32);
                    const int P = __shfl_sync(0xFu, C, srcP), Q = __shfl_sync(0xFu, C, srcQ);
                    const bool sw = (msw >> j) & 1u;
                    const int c0 = L.x + (sw ? Q : P), c1 = L.y + (sw ? P : Q);
                    const bool d = c1 < c0;
                    C = d ? c1 : c0;
                    dec |= (unsigned)d << j;
                }
            }
            if (h) dhi = dec; else dlo |= dec;
        }
    }
    __syncwarp();
    // ---- F: walk from the last scan position with state 0 (quantizer.rs:686-721) + rate (block_splitter.rs:415-460); this lane
    //      again owns positions lane and lane + 32.  (The cost table is dead from here on: lev may overlap it.)
    unsigned mdl = 0, mdh = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        mdl |= ((__shfl_sync(0xffffffffu, dlo, q) >> lane) & 1u) << q;
        mdh |= ((__shfl_sync(0xffffffffu, dhi, q) >> lane) & 1u) << q;
    }
    const unsigned pkl = lane == 0 ? ((xl >> 1) & 1u) * 3u : (((xl >> 1) & 1u) | ((((xl + 1) >> 1) & 1u) << 1));
    const unsigned pkh = ((xh >> 1) & 1u) | ((((xh + 1) >> 1) & 1u) << 1);
    unsigned inc = pos_map(pkh, mdh, tch != 0);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned u = __shfl_down_sync(0xffffffffu, inc, d);
        if (lane + d < 32) inc = map_compose(u, inc);
    }
    unsigned exc = __shfl_down_sync(0xffffffffu, inc, 1);
    const unsigned sth = lane == 31 ? 0u : (exc & 3u);            // state entering position lane + 32
    const unsigned mid = __shfl_sync(0xffffffffu, inc, 0) & 3u;   // state entering position 31
    inc = pos_map(pkl, mdl, tcl != 0);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned u = __shfl_down_sync(0xffffffffu, inc, d);
        if (lane + d < 32) inc = map_compose(u, inc);
    }
    exc = __shfl_down_sync(0xffffffffu, inc, 1);
    const unsigned stl = lane == 31 ? mid : ((exc >> (8 * mid)) & 3u);
    int qh = 0, ql = 0;
    const int dh = sth > 1, dl = stl > 1;
    if (tch != 0) {
        const unsigned a = ((xh + dh) >> 1) + ((mdh >> sth) & 1u);
        qh = a > 0 ? 2 * (int)a - dh : 0;
        if (tch < 0) qh = -qh;
    }
    if (tcl != 0) {
        const unsigned a = (lane == 0 ? (xl >> 1) : ((xl + dl) >> 1)) + ((mdl >> stl) & 1u);
        if (lane == 0) ql = (int)(int16_t)(2 * (int)a - dl);
        else ql = a > 0 ? 2 * (int)a - dl : 0;
        if (tcl < 0) ql = -ql;
    }
    __syncwarp();
    lev[sc_hi] = (int16_t)qh;
    lev[sc_lo] = (int16_t)ql;
    // a zero level costs lv[0] iff a non-zero level precedes it in the walk (= sits at a higher scan position)
    const unsigned bh = __ballot_sync(0xffffffffu, qh != 0), bl = __ballot_sync(0xffffffffu, ql != 0);
    const int lv0 = S.tb->lv[0];
    int r = 0;
    if (qh != 0) r += WB_LV((abs(qh) + dh) >> 1);
    else if (lane < 31 && (bh >> (lane + 1)) != 0) r += lv0;
    if (ql != 0) r += WB_LV((abs(ql) + dl) >> 1);
    else if (bh != 0 || (lane < 31 && (bl >> (lane + 1)) != 0)) r += lv0;
    rate_out = warp_sum(r);
    any_out = (bh | bl) != 0;
    __syncwarp();
}

// The same for ONE 16x16 TB (256 positions, 16 sub-blocks), in place: coef -> levels.  The warp's whole 5 kB scratch is laid out
// by the caller: coefficients / levels 512 B | fl: one flag byte per position (bit 0 / 1 parity of a0 for delta 0 / 1, bit 2:
// state 0 sees a0 = 0 without its rate) | lc: 4 kB cost table | (prediction, 256 B).  All 32 lanes run the chain, as eight
// identical groups of four state lanes: nothing diverges, and group w simply keeps the decision word of positions 32w .. 32w+31,
// so the 4 x 256 decision bits need no memory.  x = S / ls is recomputed where it is needed instead of being stored.
__device__ __noinline__ void trellis16_chain(const Ctx S, const DevTables *__restrict__ tab, int16_t *coef, uint8_t *fl, int4 *lc, int lane, int &rate_out, bool &any_out) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(coef); WB_SHARED_PTR(fl); WB_SHARED_PTR(lc);
    constexpr int sh = 8, off = 128;
    const int ls = tab->ls, ldq1 = S.tb->ldq[1];
    const uint16_t *scan = S.tb->scan + tab_off(4);
    auto xq_of = [&](int tc) -> unsigned {
        return tc == 0 ? 0u : min(div_ls(S, tc > 0 ? ((unsigned)tc << sh) - (unsigned)off : ((unsigned)(-tc) << sh) + (unsigned)off, (unsigned)ls), 2047u);
    };
    // ---- A1: k* (H2), any coefficient at all?
    int kstar = -1;
    bool anytc = false;
#pragma unroll 1
    for (int k = lane; k < 256; k += 32) {
        const int tc = coef[scan[k]];
        anytc |= tc != 0;
        if (xq_of(tc) >= 2) kstar = k;
    }
    kstar = warp_max(kstar);
    if (!__any_sync(0xffffffffu, anytc)) {  // every level is 0, rate 0 (as in trellis())
#pragma unroll 1
        for (int k = lane; k < 128; k += 32) reinterpret_cast<uint32_t *>(coef)[k] = 0u;
        rate_out = 0;
        any_out = false;
        __syncwarp();
        return;
    }
    // ---- A2: local costs of every position -> table, flag bytes
#pragma unroll 1
    for (int k = lane; k < 256; k += 32) {
        const int tc = coef[scan[k]];
        const unsigned x = xq_of(tc);
        const LC l = local_costs(S, tab, tc, x | ((unsigned)(tc != 0) << 11), k, kstar, ls, sh, off, ldq1);
        lc[k] = make_int4(l.L00, l.L10, l.L01, l.L11);
        fl[k] = (uint8_t)((l.pk & 3u) | ((k > kstar && (x >> 1) == 0) ? 4u : 0u));
    }
    __syncwarp();
    // ---- B: the chain, lane (w, s): state s; group w keeps decision word w
    unsigned keep = 0;
    {
        const int s = lane & 3, w_own = lane >> 2;
        const char *lp = reinterpret_cast<const char *>(lc) + (s >> 1) * 8;
        const int fsh = s < 2 ? 0 : 1;  // which parity bit of the flag byte this state looks at
        const unsigned finv = s & 1;
        int C;
        unsigned dec = 0;
        {   // DC leaf (quantizer.rs:367-409) for this lane's state
            const int tc0 = coef[0];
            const unsigned x0 = xq_of(tc0);
            const bool itz = (s == 0) && (kstar < 0);
            if (tc0 == 0) {
                C = itz ? -ldq1 : ldq1;
            } else {
                const int delta = s > 1;
                const int A0 = (int)(x0 >> 1);
                int q0 = (int)(int16_t)(2 * A0 - delta);  // H3: usize wrap gives -1 for a0 == 0, delta == 1
                if (tc0 < 0) q0 = -q0;
                const int d0 = abs(tc0 - ((q0 * ls + off) >> sh));
                const int bits0 = (A0 != 0 || !itz) ? A0 + 1 : 0;
                const int cost0 = 128 * d0 + WB_LDQ(bits0);
                const int A1 = A0 + 1;
                int q1 = 2 * A1 - delta;
                if (tc0 < 0) q1 = -q1;
                const int d1 = abs(tc0 - ((q1 * ls + off) >> sh));
                const int cost1 = 128 * d1 + WB_LDQ(A1 + 1);
                if (cost0 <= cost1) {
                    C = cost0;
                    if (itz && A0 == 0) C -= ldq1;
                } else {
                    C = cost1;
                    dec = 1u;
                }
            }
        }
        // states 0,1 continue from {0,2}, states 2,3 from {1,3}; which of the two feeds candidate a0 depends on the parity of a0
        const int srcP = (lane & ~3) | (s >> 1), srcQ = srcP + 2;
#pragma unroll 1
        for (int sb = 0; sb < 16; sb++) {  // one sub-block of 16 positions per iteration
            const int jb = sb * 16;
            int j0 = jb + 1;
            if (sb) {  // sub-block start: renormalise, then the step with the post-comparison adjustment (quantizer.rs:512-514)
                int mn = min(C, __shfl_xor_sync(0xffffffffu, C, 1));
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 2));
                C -= mn;
                const int2 L = *reinterpret_cast<const int2 *>(lp + jb * 16);
                const unsigned f = fl[jb];
                const int P = __shfl_sync(0xffffffffu, C, srcP), Q = __shfl_sync(0xffffffffu, C, srcQ);
                const bool sw = (((f >> fsh) ^ finv) & 1u) != 0;
                const int La = L.x - ((s == 0 && (f & 4u)) ? ldq1 : 0);
                const int c0 = La + (sw ? Q : P), c1 = L.y + (sw ? P : Q);
                const bool d = c1 < c0;  // ties keep a0 (quantizer.rs:505)
                C = d ? c1 : c0;
                if (s == 0 && jb > kstar && !d) C -= ldq1;
                dec |= (unsigned)d << (jb & 31);
            }
#pragma unroll 5
            for (int j = j0; j < jb + 16; j++) {
                const int2 L = *reinterpret_cast<const int2 *>(lp + j * 16);
                const unsigned f = fl[j];
                const int P = __shfl_sync(0xffffffffu, C, srcP), Q = __shfl_sync(0xffffffffu, C, srcQ);
                const bool sw = (((f >> fsh) ^ finv) & 1u) != 0;
                const int La = L.x - ((s == 0 && (f & 4u)) ? ldq1 : 0);
                const int c0 = La + (sw ? Q : P), c1 = L.y + (sw ? P : Q);
                const bool d = c1 < c0;
                C = d ? c1 : c0;
                dec |= (unsigned)d << (j & 31);
            }
            if (sb & 1) {  // a decision word is complete
                if (w_own == (sb >> 1)) keep = dec;
                dec = 0;
            }
        }
    }
    __syncwarp();
    // ---- F: walk from the last scan position with state 0 (quantizer.rs:686-721) + rate (block_splitter.rs:415-460), 32 positions
    //      at a time from the top; this lane owns position 32 i + lane of segment i
    unsigned state = 0;  // state entering the segment's highest position
    bool seen_nz = false;
    int rate = 0;
    const int lv0 = S.tb->lv[0];
#pragma unroll 1
    for (int i = 7; i >= 0; i--) {
        const int k = 32 * i + lane;
        const int sc = scan[k];
        const int tc = coef[sc];
        const unsigned x = xq_of(tc);
        unsigned md = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) md |= ((__shfl_sync(0xffffffffu, keep, 4 * i + q) >> lane) & 1u) << q;
        const unsigned pk = k == 0 ? ((x >> 1) & 1u) * 3u : (((x >> 1) & 1u) | ((((x + 1) >> 1) & 1u) << 1));
        unsigned inc = pos_map(pk, md, tc != 0);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned u = __shfl_down_sync(0xffffffffu, inc, d);
            if (lane + d < 32) inc = map_compose(u, inc);
        }
        const unsigned exc = __shfl_down_sync(0xffffffffu, inc, 1);
        const unsigned st = lane == 31 ? state : ((exc >> (8 * state)) & 3u);  // state entering this lane's position
        state = (__shfl_sync(0xffffffffu, inc, 0) >> (8 * state)) & 3u;
        const int dl = st > 1;
        int q = 0;
        if (tc != 0) {
            const unsigned a = (k == 0 ? (x >> 1) : ((x + dl) >> 1)) + ((md >> st) & 1u);
            if (k == 0) q = (int)(int16_t)(2 * (int)a - dl);
            else q = a > 0 ? 2 * (int)a - dl : 0;
            if (tc < 0) q = -q;
        }
        coef[sc] = (int16_t)q;  // in place: this thread read the coefficient above
        // a zero level costs lv[0] iff a non-zero level precedes it in the walk (= sits at a higher scan position)
        const unsigned bal = __ballot_sync(0xffffffffu, q != 0);
        if (q != 0) rate += WB_LV((abs(q) + dl) >> 1);
        else if (seen_nz || (lane < 31 && (bal >> (lane + 1)) != 0)) rate += lv0;
        seen_nz = seen_nz || bal != 0;
    }
    rate_out = warp_sum(rate);
    any_out = seen_nz;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// tasks
// ---------------------------------------------------------------------------------------------------------------
__device__ WarpScratch warp_scratch(Shared &S, int warp) {
    WarpScratch ws;
    ws.A = S.wbuf[warp].A; ws.B = S.wbuf[warp].B; ws.Wd = reinterpret_cast<uint16_t *>(S.wbuf[warp].B); ws.pred = S.wbuf[warp].P;
    ws.refx = S.refx[warp];
    return ws;
}

// prediction of one (mode, component) block of size >= 4 by one warp.  Writes the samples to pred_out (may be null; raster,
// 4-byte aligned) and returns the SAD against the source block.  Angular modes: four samples per lane and iteration (ang4.cuh);
// planar / DC / CCLM: one sample per lane (pred_sample).
__device__ __noinline__ unsigned predict_block(const Ctx S, const CtuGeom g, const Node nd, int c, int mode, int16_t *refx, uint8_t *pred_out, int lane) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(refx);
    if (pred_out) WB_SHARED_PTR(pred_out);
    const int cs = c != 0, n = nd.w >> cs, bx = nd.x >> cs, by = nd.y >> cs, l2 = ilog2i(n);
    unsigned sad = 0;
    if (mode >= 2 && mode <= 66) {  // most predictions: the 13 coarse + 4 refinement SADs, 3 of the 5 full evaluations
        a4::Mode m = a4_setup(S, c, n, l2, mode);
        if (m.ang < 0) {  // project the side line below index 0 into the warp's scratch line ([-n-3, n+6] around index 0)
            uint8_t *pr = reinterpret_cast<uint8_t *>(refx) + 36;
#pragma unroll RU
            for (int i = lane; i < 2 * n + 2; i += 32) a4::project_elem(pr, m.main, m.side, n, m.inv, i);
            __syncwarp();
            m.main = pr;
        }
        const int qsh = l2 - 2, nquads = n << qsh;
        const bool vert = mode >= 34;
#pragma unroll 1
        for (int q = lane; q < nquads; q += 32) {
            const int t = q >> qsh, u0 = (q & ((1 << qsh) - 1)) << 2;
            const unsigned p4 = a4::quad(m, S.tb->taps, t, u0);
            sad = a4::sad4(p4, org_quad(S, c, mode, bx, by, t, u0), sad);
            if (pred_out) {
                if (vert) *reinterpret_cast<unsigned *>(pred_out + (t << l2) + u0) = p4;
                else {
                    uint8_t *d = pred_out + (u0 << l2) + t;
                    d[0] = (uint8_t)p4; d[n] = (uint8_t)(p4 >> 8); d[2 * n] = (uint8_t)(p4 >> 16); d[3 * n] = (uint8_t)(p4 >> 24);
                }
            }
        }
        sad = warp_sumu(sad);
        __syncwarp();
        return sad;
    }
    PredCtx pc_mem;  // cclm_params takes its address
    pred_setup(S, g, nd, c, mode, lane, pc_mem);
    const PredCtx pc = pc_mem;  // private copy whose address never escapes: the per-sample loop keeps it in registers
    const uint8_t *org = cs ? S.c->orgC[c - 1] + by * 16 + bx : S.c->orgY + by * 32 + bx;  // source block, row stride 16 / 32
    const int osh = cs ? 4 : 5;
#pragma unroll 1
    for (int i = lane; i < n * n; i += 32) {
        int y = i >> l2, x = i & (n - 1);
        int p = pred_sample(S, pc, x, y);
        if (pred_out) pred_out[i] = (uint8_t)p;
        sad += abs(p - (int)org[(y << osh) + x]);
    }
    return warp_sumu(sad);
}

// SAD of one (mode, component) (block_splitter.rs:64-108 / 476-522)
__device__ __forceinline__ unsigned sad_task(const Ctx S, const CtuGeom g, const Node nd, int c, int mode, const WarpScratch ws, int lane) {
    return predict_block(S, g, nd, c, mode, ws.refx, nullptr, lane);
}

// full evaluation of one (mode, component): block_splitter.rs:148-183 + rate 415-460.  The outcome (reconstruction, levels)
// goes to candidate slot `slot` of the CTU's global scratch; the winner is committed from there (commit_slot).
__device__ __noinline__ void full_task(const Ctx S, const DevTables *__restrict__ tab, const CtuGeom g, const Node nd, int c, int mode,
                          const WarpScratch ws, int lane, unsigned &ssd_out, int &rate_out, int slot) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(ws.A); WB_SHARED_PTR(ws.B); WB_SHARED_PTR(ws.Wd); WB_SHARED_PTR(ws.pred); WB_SHARED_PTR(ws.refx);
    const int cs = c != 0, n = nd.w >> cs, bx = nd.x >> cs, by = nd.y >> cs, l2 = ilog2i(n), nn = n * n;
    // (a 16x16 TB keeps its prediction in the last 256 bytes of the warp's scratch: everything between the coefficients and it is the
    //  cost table of trellis16_chain)
    uint8_t *const pred = ws.pred + ((WB_CHAIN16 && l2 == 4) ? 768 : 0);
    const bool anysad = predict_block(S, g, nd, c, mode, ws.refx, pred, lane) != 0;
    __syncwarp();
    int16_t *A = ws.A, *B = ws.B;
    bool anyres = anysad;
    const uint8_t *org = cs ? S.c->orgC[c - 1] + by * 16 + bx : S.c->orgY + by * 32 + bx;  // source block, row stride 16 / 32
    const int osh = cs ? 4 : 5;
#pragma unroll 1
    for (int i = lane; i < nn; i += 32) {
        int y = i >> l2, x = i & (n - 1);
        A[i] = (int16_t)((int)org[(y << osh) + x] - (int)pred[i]);
    }
    anyres = __any_sync(0xffffffffu, anyres);
    __syncwarp();
    const int to = tab_off(l2);
    bool anylev = false;
    int rate = 0;
    if (anyres) {
        mm_rows_q<true>(S.tb->Qr + to / 4, A, B, n, l2, 1 << (l2 - 2), l2 - 1, lane);
        __syncwarp();
        mm_cols_q(S.tb->T + to, reinterpret_cast<const int32_t *>(B), A, n, l2, 1 << (l2 + 5), l2 + 6, false, lane);
        __syncwarp();
        // dependent quantisation: coefficients -> levels, in place in A (every position's coefficient is read by the thread that
        // writes its level); B holds the trellis words / the cost table meanwhile
#if WB_CHAIN8
        if (l2 == 3) trellis8_chain(S, tab, A, reinterpret_cast<uint16_t *>(A + 128), A, reinterpret_cast<int4 *>(B), lane, rate, anylev);
        else
#endif
#if WB_CHAIN16
        if (l2 == 4) trellis16_chain(S, tab, A, reinterpret_cast<uint8_t *>(A + 256), reinterpret_cast<int4 *>(A + 384), lane, rate, anylev);
        else
#endif
        trellis(S, tab, A, l2, reinterpret_cast<uint16_t *>(B), A, lane, rate, anylev);
    } else {
#pragma unroll RU
        for (int i = lane; i < nn; i += 32) A[i] = 0;
        __syncwarp();
    }
    // candidate slot (planar, DC, dir, dir-1, dir+1, CCLM): the evaluation's outcome goes to the CTU's global scratch, block-local raster
    const int soff = c == 0 ? 0 : (c == 1 ? 1024 : 1280);
    uint8_t *gRec = nullptr;
    if (slot >= 0) {
        gRec = S.c->groot + slot * ROOT_SLOT_SAMPLES + soff;
        int16_t *gLv = reinterpret_cast<int16_t *>(S.c->groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES) + slot * ROOT_SLOT_SAMPLES + soff;
#pragma unroll RU
        for (int i = lane; i < nn / 2; i += 32) reinterpret_cast<uint32_t *>(gLv)[i] = reinterpret_cast<const uint32_t *>(A)[i];  // two levels per store
    }
    if (anylev) {
        const int sh = l2 + 4, off = 1 << (sh - 1), ls = tab->ls;
        // dequantise (quantizer.rs:1074-1075) into the pair-packed layout of the column pass
#pragma unroll RU
        for (int o = lane; o < nn / 2; o += 32) {
            const int ip = o >> l2, x = o & (n - 1);
            const int d0 = min(32767, max(-32768, ((int)A[(2 * ip) * n + x] * ls + off) >> sh));
            const int d1 = min(32767, max(-32768, ((int)A[(2 * ip + 1) * n + x] * ls + off) >> sh));
            reinterpret_cast<int32_t *>(B)[o] = (d0 & 0xffff) | (d1 << 16);
        }
        __syncwarp();
        // vertical: V[y][x] = clamp16((sum_i T[i][y] * D[i][x] + 64) >> 7)   (Tt's rows are T's columns)
        mm_cols_q(S.tb->Tt + to, reinterpret_cast<const int32_t *>(B), A, n, l2, 64, 7, true, lane);
        __syncwarp();
        // horizontal: R[y][x] = (sum_i T[i][x] * V[y][i] + 2048) >> 12
        mm_rows_q<false>(S.tb->Qc + to / 4, A, B, n, l2, 2048, 12, lane);
        __syncwarp();
    }
    unsigned ssd = 0;
#pragma unroll 1
    for (int i = lane; i < nn; i += 32) {
        int y = i >> l2, x = i & (n - 1);
        int res = anylev ? (int)B[i] : 0;
        int rec = clip8((int)(int16_t)((int)pred[i] + res));
        int d = rec - (int)org[(y << osh) + x];
        ssd += (unsigned)(d * d);
        if (slot >= 0) gRec[i] = (uint8_t)rec;
    }
    ssd_out = warp_sumu(ssd);
    rate_out = rate;
    __syncwarp();
}

// Full evaluation of a 4x4 TB by HALF a warp (lanes 0-15 and 16-31 evaluate two independent TBs of the same node: two modes
// of a 4x4 luma CU, or the Cb and Cr blocks of one chroma mode).  Same arithmetic as full_task, but the block lives in
// registers (lane gl = raster sample 4y + x), the 4-point transforms exchange operands by shuffles, and the dependent
// quantisation runs one trellis STATE per lane (lanes 0-3 of the half) over the 15 sequential steps, reading the local costs
// that the position lanes tabulated in shared memory.  c, mode, commit and slot may differ between the halves, but BOTH halves
// always run in step (all 32 lanes call; the prediction kind - angular / planar-DC / CCLM - must be the same): a half whose TB
// is not wanted passes on = false, evaluates the TB it is given and stores nothing.  That keeps every shuffle / ballot a
// full-warp collective with a constant mask: with per-half masks each of the ~45 shuffles carried its own convergence code
// (325 of the function's 1 586 instructions).
constexpr unsigned long long INV_SCAN4 = 0xFDA6EB73C8419520ull;  // nibble r = scan position of raster offset r (inverse of the 4x4 diagonal scan)
__device__ __forceinline__ int dot4_s8(int packed, int a0, int a1, int a2, int a3) {
    return (int)(int8_t)(packed & 255) * a0 + (int)(int8_t)((packed >> 8) & 255) * a1 + (int)(int8_t)((packed >> 16) & 255) * a2 + (packed >> 24) * a3;
}
__device__ __noinline__ void full_pair4(const Ctx S, const DevTables *__restrict__ tab, const CtuGeom g, const Node nd, int c, int mode, bool commit, int slot,
                                        const WarpScratch ws, int lane, unsigned &ssd_out, int &rate_out, bool on = true) {
    WB_SHARED_CTX(S);
    WB_SHARED_PTR(ws.A); WB_SHARED_PTR(ws.B);
    const int hb = lane & 16, gl = lane & 15, half = hb >> 4;
    constexpr unsigned hm = 0xffffffffu;  // the two halves run in step (see above): every collective is a full-warp one with a constant mask
    if (!on) { commit = false; slot = -1; }
    const int cs = c != 0, bx = nd.x >> cs, by = nd.y >> cs;
    const int x = gl & 3, y = gl >> 2;
    int p;  // prediction sample, straight from the reference samples (no projection array, no per-task setup pass)
    if (mode >= 2 && mode <= 66) {
        p = ang_sample_direct(S, c, 4, 2, mode, x, y);
    } else if (mode > 66) {
        PredCtx pc;
        cclm_params(S, g, nd, c, mode, pc);
        p = pc.cclm128 ? 128 : clip8(((S.c->pds[gl] * pc.a) >> pc.k) + pc.b);
    } else {  // planar / DC with PDPC, n = 4 (nScale 0, unfiltered references)
        const int16_t *lrs = S.c->refL[c][0] + 1, *ars = S.c->refA[c][0];
        if (mode == MODE_PLANAR) {
            p = (((3 - y) * ars[x] + (y + 1) * lrs[4] + (3 - x) * lrs[y] + (x + 1) * ars[4] + 4) >> 3) & 255;
        } else {
            p = ((ars[0] + ars[1] + ars[2] + ars[3] + lrs[0] + lrs[1] + lrs[2] + lrs[3] + 4) >> 3) & 255;
        }
        const int wl = pdpc_w(0, x), wt = pdpc_w(0, y);
        p = clip8((int)(int16_t)(lrs[y] * wl + ars[x] * wt + (64 - wt - wl) * p + 32) >> 6);
    }
    const int org = org_at(S, c, bx + x, by + y);
    const int res = org - p;
    // ---- forward DCT (transformer.rs:2040-2378 with n = 4: shifts 1 and 8)
    const int rowq = 4 * y, colq = x;
    const int Trow_x = *reinterpret_cast<const int *>(S.tb->T + 4 * x);    // T[x][0..3]
    const int Trow_y = *reinterpret_cast<const int *>(S.tb->T + 4 * y);    // T[y][0..3]
    const int Tcol_x = *reinterpret_cast<const int *>(S.tb->Tt + 4 * x);   // T[0..3][x]
    const int Tcol_y = *reinterpret_cast<const int *>(S.tb->Tt + 4 * y);   // T[0..3][y]
    int a0 = __shfl_sync(hm, res, hb + rowq), a1 = __shfl_sync(hm, res, hb + rowq + 1), a2 = __shfl_sync(hm, res, hb + rowq + 2), a3 = __shfl_sync(hm, res, hb + rowq + 3);
    const int b1 = (int)(int16_t)((dot4_s8(Trow_x, a0, a1, a2, a3) + 1) >> 1);           // B[y][i = x]
    a0 = __shfl_sync(hm, b1, hb + colq); a1 = __shfl_sync(hm, b1, hb + colq + 4); a2 = __shfl_sync(hm, b1, hb + colq + 8); a3 = __shfl_sync(hm, b1, hb + colq + 12);
    const int coef = (int)(int16_t)((dot4_s8(Trow_y, a0, a1, a2, a3) + 128) >> 8);       // coef[i = y][x]
    // ---- dependent quantisation: the 4x4 case of trellis() (quantizer.rs:338-517, 686-721; rate block_splitter.rs:415-460)
    const int sh = 6, off = 32, ls = tab->ls, ldq1 = S.tb->ldq[1];
    const int k = gl;  // scan position owned by this lane
    const int tc = __shfl_sync(hm, coef, hb + S.tb->scan[k]);
    unsigned xq = 0;
    const unsigned nz = tc != 0;
    if (nz) {
        const unsigned sc = tc > 0 ? ((unsigned)tc << sh) - (unsigned)off : ((unsigned)(-tc) << sh) + (unsigned)off;
        xq = min(div_ls(S, sc, (unsigned)ls), 2047u);
    }
    const unsigned w = xq | (nz << 11);
    const unsigned bstar = (__ballot_sync(hm, xq >= 2) >> hb) & 0xffffu;
    const int kstar = bstar ? 31 - __clz(bstar) : -1;
    const bool anytc = ((__ballot_sync(hm, nz) >> hb) & 0xffffu) != 0;
    int q = 0, rate = 0;
    bool anylev = false;
    if (__any_sync(hm, anytc)) {  // warp-uniform: a half without coefficients quantises zeros to zeros (rate 0)
        LC l;
        l.L00 = l.L01 = l.L0s0 = 0; l.L10 = l.L11 = TR_INF; l.pk = 0;
        if (k > 0) l = local_costs(S, tab, tc, w, k, kstar, ls, sh, off, ldq1);  // (k & 15) != 0: no sub-block-start adjustment inside a 4x4 TB
        int *tbl = reinterpret_cast<int *>(half ? ws.B : ws.A);  // [position][state] -> (cost of candidate a0, cost of candidate a1)
        tbl[8 * k + 0] = l.L0s0; tbl[8 * k + 1] = l.L10;
        tbl[8 * k + 2] = l.L00;  tbl[8 * k + 3] = l.L10;
        tbl[8 * k + 4] = l.L01;  tbl[8 * k + 5] = l.L11;
        tbl[8 * k + 6] = l.L01;  tbl[8 * k + 7] = l.L11;
        const unsigned m1 = (__ballot_sync(hm, l.pk & 1u) >> hb) & 0xffffu, m2 = (__ballot_sync(hm, l.pk & 2u) >> hb) & 0xffffu;
        const int tc0 = __shfl_sync(hm, tc, hb);
        const unsigned x0 = __shfl_sync(hm, xq, hb);
        __syncwarp(hm);
        unsigned dec = 0;
        if (gl < 4) {
            const int s = gl;
            constexpr unsigned m4 = 0x000F000Fu;  // the state lanes of both halves
            // DC leaf (quantizer.rs:367-409) for this lane's state
            int C;
            {
                const bool itz = (s == 0) && (kstar < 0);
                if (tc0 == 0) {
                    C = itz ? -ldq1 : ldq1;
                } else {
                    const int delta = s > 1;
                    const int A0 = (int)(x0 >> 1);
                    int q0 = (int)(int16_t)(2 * A0 - delta);  // H3: usize wrap gives -1 for a0 == 0, delta == 1
                    if (tc0 < 0) q0 = -q0;
                    const int d0 = abs(tc0 - ((q0 * ls + off) >> sh));
                    const int bits0 = (A0 != 0 || !itz) ? A0 + 1 : 0;
                    const int cost0 = 128 * d0 + WB_LDQ(bits0);
                    const int A1 = A0 + 1;
                    int q1 = 2 * A1 - delta;
                    if (tc0 < 0) q1 = -q1;
                    const int d1 = abs(tc0 - ((q1 * ls + off) >> sh));
                    const int cost1 = 128 * d1 + WB_LDQ(A1 + 1);
                    if (cost0 <= cost1) {
                        C = cost0;
                        if (itz && A0 == 0) C -= ldq1;
                    } else {
                        C = cost1;
                        dec = 1u;
                    }
                }
            }
            // states 0,1 continue from {0,2}, states 2,3 from {1,3}; which of the two feeds candidate a0 depends on the parity of a0
            const int srcP = hb + (s >> 1), srcQ = srcP + 2;
            const unsigned msw = (s < 2 ? m1 : m2) ^ ((s & 1) ? 0xffffu : 0u);
            const int *tp = tbl + 2 * s;
#pragma unroll U5
            for (int j = 1; j < 16; j++) {
                const int La = tp[8 * j], Lb = tp[8 * j + 1];
                const int P = __shfl_sync(m4, C, srcP), Q = __shfl_sync(m4, C, srcQ);
                const bool sw = (msw >> j) & 1u;
                const int c0 = La + (sw ? Q : P), c1 = Lb + (sw ? P : Q);
                const bool d = c1 < c0;  // ties keep a0 (quantizer.rs:505)
                C = d ? c1 : c0;
                dec |= (unsigned)d << j;
            }
        }
        __syncwarp(hm);
        unsigned mydec = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) mydec |= ((__shfl_sync(hm, dec, hb + s) >> k) & 1u) << s;
        const unsigned pk = k == 0 ? ((w >> 1) & 1u) * 3u : l.pk;
        unsigned inc = pos_map(pk, mydec, nz != 0);
        // walk from the last scan position with state 0
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const unsigned t = __shfl_down_sync(hm, inc, d, 16);
            if (gl + d < 16) inc = map_compose(t, inc);
        }
        const unsigned exc = __shfl_down_sync(hm, inc, 1, 16);
        const unsigned s = gl == 15 ? 0u : (exc & 3u);
        const int delta = s > 1;
        if (nz) {
            const unsigned a = ((k == 0) ? (xq >> 1) : ((xq + delta) >> 1)) + ((mydec >> s) & 1u);
            if (k == 0) q = (int)(int16_t)(2 * (int)a - delta);
            else q = a > 0 ? 2 * (int)a - delta : 0;
            if (tc < 0) q = -q;
        }
        const bool has = q != 0;
        const unsigned bal = (__ballot_sync(hm, has) >> hb) & 0xffffu;
        if (has) rate = WB_LV((abs(q) + delta) >> 1);
        else if ((bal >> (gl + 1)) != 0) rate = S.tb->lv[0];
        rate = half_sum(rate, hm);
        anylev = bal != 0;
    }
    // levels back to raster order
    const int lev = __shfl_sync(hm, q, hb + (int)((INV_SCAN4 >> (4 * gl)) & 15ull));
    const int soff = slot * ROOT_SLOT_SAMPLES + (c == 0 ? 0 : (c == 1 ? 1024 : 1280)) + gl;
    if (commit) {
        if (c == 0) S.c->lvY[(by + y) * 32 + bx + x] = (int16_t)lev;
        else S.c->lvC[c - 1][(by + y) * 16 + bx + x] = (int16_t)lev;
    }
    if (slot >= 0) reinterpret_cast<int16_t *>(S.c->groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES)[soff] = (int16_t)lev;
    int r = 0;
    if (__any_sync(hm, anylev)) {  // warp-uniform (all-zero levels give a zero residual); dequantise (quantizer.rs:1074-1075), inverse DCT (transformer.rs:2380-2737 with n = 4)
        const int dq = min(32767, max(-32768, (lev * ls + off) >> sh));
        a0 = __shfl_sync(hm, dq, hb + colq); a1 = __shfl_sync(hm, dq, hb + colq + 4); a2 = __shfl_sync(hm, dq, hb + colq + 8); a3 = __shfl_sync(hm, dq, hb + colq + 12);
        const int v = min(32767, max(-32768, (dot4_s8(Tcol_y, a0, a1, a2, a3) + 64) >> 7));   // V[y][x] = sum_i T[i][y] D[i][x]
        a0 = __shfl_sync(hm, v, hb + rowq); a1 = __shfl_sync(hm, v, hb + rowq + 1); a2 = __shfl_sync(hm, v, hb + rowq + 2); a3 = __shfl_sync(hm, v, hb + rowq + 3);
        r = (int)(int16_t)((dot4_s8(Tcol_x, a0, a1, a2, a3) + 2048) >> 12);                    // R[y][x] = sum_i T[i][x] V[y][i]
    }
    const int rec = clip8((int)(int16_t)(p + r));
    const int d = rec - org;
    if (commit) {
        if (c == 0) RY(S, bx + x, by + y) = (uint8_t)rec;
        else RC(S, c, bx + x, by + y) = (uint8_t)rec;
    }
    if (slot >= 0) S.c->groot[soff] = (uint8_t)rec;
    ssd_out = (unsigned)half_sum(d * d, hm);
    rate_out = rate;
    __syncwarp(hm);
}

// SAD-driven choice among the three CCLM modes (block_splitter.rs:1041-1054 / 812-850), one warp, chroma blocks of any size.
// The (a, k, b) derivation is scalar work, so the six (mode, component) combinations are derived at once, one per lane
// (lanes 0-5; the other lanes repeat them), and handed to the sample lanes (0-15 Cb, 16-31 Cr, 16 samples per iteration) by
// shuffles.  Order LT, T, L with the reference's tie rules (LT unless strictly worse, then T).
__device__ __noinline__ int cclm_search(const Ctx S, const CtuGeom g, const Node nd, int lane) {
    WB_SHARED_CTX(S);
    const int combo = min(lane & 7, 5);  // mode index * 2 + component - 1
    PredCtx pc;
    {
        const int mi = combo >> 1;
        cclm_params(S, g, nd, 1 + (combo & 1), mi == 0 ? MODE_LT_CCLM : (mi == 1 ? MODE_T_CCLM : MODE_L_CCLM), pc);
    }
    const int pa = pc.cclm128 ? 0 : pc.a, pk = pc.cclm128 ? 0 : pc.k, pb = pc.cclm128 ? 128 : pc.b;  // a = 0: the prediction is b
    const int cidx = lane >> 4, n = nd.w >> 1, l2 = ilog2i(n), bx = nd.x >> 1, by = nd.y >> 1;
    int a[3], k[3], b[3];
#pragma unroll
    for (int mi = 0; mi < 3; mi++) {
        const int src = 2 * mi + cidx;
        a[mi] = __shfl_sync(0xffffffffu, pa, src); k[mi] = __shfl_sync(0xffffffffu, pk, src); b[mi] = __shfl_sync(0xffffffffu, pb, src);
    }
    const uint8_t *org = S.c->orgC[cidx] + by * 16 + bx;
    unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll 1
    for (int i = lane & 15; i < n * n; i += 16) {
        const int o = (int)org[((i >> l2) << 4) + (i & (n - 1))], ds = S.c->pds[i];
        s0 += (unsigned)abs(clip8(((ds * a[0]) >> k[0]) + b[0]) - o);
        s1 += (unsigned)abs(clip8(((ds * a[1]) >> k[1]) + b[1]) - o);
        s2 += (unsigned)abs(clip8(((ds * a[2]) >> k[2]) + b[2]) - o);
    }
    s0 = warp_sumu(s0); s1 = warp_sumu(s1); s2 = warp_sumu(s2);
    if (s0 <= s1 && s0 <= s2) return MODE_LT_CCLM;
    return s1 <= s2 ? MODE_T_CCLM : MODE_L_CCLM;
}

// The winner of a node was already evaluated with unchanged inputs (the reference repeats that evaluation,
// block_splitter.rs:989-1037 / 1062-1076, with identical results): copy its reconstruction and levels out of its slot
// (global scratch, written by other warps of this CTA before the last block barrier; read around L1).
__device__ __noinline__ void commit_slot(const Ctx S, const Node nd, int c, int slot, int lane) {
    WB_SHARED_CTX(S);
    const int cs = c != 0, n = nd.w >> cs, bx = nd.x >> cs, by = nd.y >> cs, l2 = ilog2i(n), nn = n * n;
    const int soff = slot * ROOT_SLOT_SAMPLES + (c == 0 ? 0 : (c == 1 ? 1024 : 1280));
    const uint8_t *gRec = S.c->groot + soff;
    const int16_t *gLv = reinterpret_cast<const int16_t *>(S.c->groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES) + soff;
    int16_t *dst = c == 0 ? S.c->lvY : S.c->lvC[c - 1];
    const int stride = c == 0 ? 32 : 16;
#pragma unroll RU
    for (int i = lane; i < nn; i += 32) {
        int y = i >> l2, x = i & (n - 1);
        dst[(by + y) * stride + bx + x] = __ldcg(gLv + i);
        const uint8_t r = __ldcg(gRec + i);
        if (c == 0) RY(S, bx + x, by + y) = r;
        else RC(S, c, bx + x, by + y) = r;
    }
    __syncwarp();
}

// The same for the 32x32 root CU, four samples per lane and iteration.  (The slots were written by other warps of this CTA
// before the last block barrier; read around L1.)
__device__ __noinline__ void commit_root_slot(const Ctx S, int c, int slot, int lane) {
    WB_SHARED_CTX(S);
    const int n = c == 0 ? 32 : 16, l2 = c == 0 ? 5 : 4, nn = n * n;
    const int soff = c == 0 ? 0 : (c == 1 ? 1024 : 1280);
    const uint8_t *gRec = S.c->groot + slot * ROOT_SLOT_SAMPLES + soff;
    const int16_t *gLv = reinterpret_cast<const int16_t *>(S.c->groot + ROOT_SLOTS * ROOT_SLOT_SAMPLES) + slot * ROOT_SLOT_SAMPLES + soff;
    int16_t *dst = c == 0 ? S.c->lvY : S.c->lvC[c - 1];
    for (int i = lane; i < nn / 4; i += 32) {  // 4 samples per lane and iteration
        const unsigned r4 = __ldcg(reinterpret_cast<const unsigned *>(gRec) + i);
        const uint2 l4 = __ldcg(reinterpret_cast<const uint2 *>(gLv) + i);
        const int y = (4 * i) >> l2, x = (4 * i) & (n - 1);
        *reinterpret_cast<uint2 *>(dst + y * n + x) = l4;  // the CTU level arrays have the block's own stride at the root
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint8_t r = (uint8_t)(r4 >> (8 * j));
            if (c == 0) RY(S, x + j, y) = r;
            else RC(S, c, x + j, y) = r;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// MPM derivation for the mode-bit estimate (ctu.rs:1498-1635) with the H1 neighbour semantics
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ int luma_kind(const Ctx S, const CtuGeom g, const Node nd, int mode, int root_mode) {
    WB_SHARED_CTX(S);
    if (mode == MODE_PLANAR) return 0;
    int left, above;
    if (nd.x == 0) left = g.cx > 0 ? S.c->leftModes[(nd.y + nd.w - 1) >> 2] : MODE_PLANAR;
    else left = root_mode;
    above = nd.y == 0 ? MODE_PLANAR : root_mode;
    int cand[5];
    if (left == above && left > MODE_DC) {
        int m = left;
        cand[0] = m; cand[1] = 2 + (m + 61) % 64; cand[2] = 2 + (m - 1) % 64; cand[3] = 2 + (m + 60) % 64; cand[4] = 2 + m % 64;
    } else if (left != above && (left > MODE_DC || above > MODE_DC)) {
        int mn = min(left, above), mx = max(left, above);
        if (mn > MODE_DC) {
            int d = mx - mn;
            cand[0] = left; cand[1] = above;
            if (d == 1) { cand[2] = 2 + (mn + 61) % 64; cand[3] = 2 + (mx - 1) % 64; cand[4] = 2 + (mn + 60) % 64; }
            else if (d >= 62) { cand[2] = 2 + (mn - 1) % 64; cand[3] = 2 + (mx + 61) % 64; cand[4] = 2 + mn % 64; }
            else if (d == 2) { cand[2] = 2 + (mn - 1) % 64; cand[3] = 2 + (mn + 61) % 64; cand[4] = 2 + (mx - 1) % 64; }
            else { cand[2] = 2 + (mn + 61) % 64; cand[3] = 2 + (mn - 1) % 64; cand[4] = 2 + (mx + 61) % 64; }
        } else {
            cand[0] = mx; cand[1] = 2 + (mx + 61) % 64; cand[2] = 2 + (mx - 1) % 64; cand[3] = 2 + (mx + 60) % 64; cand[4] = 2 + mx % 64;
        }
    } else {
        cand[0] = MODE_DC; cand[1] = 50; cand[2] = 18; cand[3] = 46; cand[4] = 54;
    }
#pragma unroll
    for (int i = 0; i < 5; i++)
        if (cand[i] == mode) return 1 + i;
    // remainder = mode - 1 - #(candidates < mode)  (sorted-candidate ladder of ctu.rs:1603-1633)
    int below = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) below += cand[i] < mode;
    return 6 + mode - 1 - below;
}

}  // namespace wb
