// slice_coder.cu — phase 2 of the hot path: the decided trees are turned into the CABAC-coded slice_data() of each picture on
// the device (reference src/ctu_encoder.rs:227-2269 syntax order, src/bool_coder.rs:136-296 arithmetic coder).
//
// Every context index of the emitted subset depends only on the decisions (levels, modes, tree), never on the arithmetic
// coder's state, so the stage is a syntax walk that produces bin strings and an arithmetic coder that consumes them:
//   wrenc_b200_nzmap_kernel        one warp per CTU: map of the CTU's non-zero 4x4 level blocks (coalesced 16-byte loads);
//   wrenc_b200_syntax_kernel       one thread per CTU: walks the CTU's coding tree in syntax order and writes the bin string as
//                                  16-bit entries (context index | bin value | bypass flag): counting + staging pass, then a
//                                  second pass for the CTUs whose string outgrew its staging slot;
//   wrenc_b200_bin_scan_kernel / wrenc_b200_bin_compact_kernel   arena offsets of the strings, staged strings -> arena;
//   wrenc_b200_cabac_kernel        three warps per picture (context states / interval widths / code value and bytes) over the
//                                  CTUs' bin strings in raster order; terminates, byte-aligns.  The engine is cabac_engine.cuh.
// I-slice subset actually emitted: split_cu_flag; intra_luma_mpm_flag / not_planar / mpm_idx / mpm_remainder;
// cclm_mode_flag / cclm_mode_idx / intra_chroma_pred_mode(=4); tu_cb/cr/y_coded_flag; cu_qp_delta_abs(=0) once per CTU;
// transform_skip_flag(=0); residual_coding with dependent quantisation; mts_idx(=0); end_of_slice_one_bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "syntax_walk.cuh"
#include "cabac_engine.cuh"

namespace wb {

// Map of the non-zero 4x4 level blocks, one warp per CTU, coalesced: lane r reads luma row r (64 bytes) and one chroma row (lanes
// 0-15 Cb, 16-31 Cr, 32 bytes); a ballot per block column gathers the rows.  The syntax walk (one THREAD per CTU, scattered 2-byte
// loads) then answers "is this CU / sub-block coded" and "where is the last coded sub-block" from the map instead of scanning.
extern "C" __global__ void __launch_bounds__(256) wrenc_b200_nzmap_kernel(SyntaxParams Q) {
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nctu = Q.Wc * Q.Hc;
    if (gid >= (long long)Q.n_pics * nctu) return;
    const int pic = (int)(gid / nctu), ctu = (int)(gid - (long long)pic * nctu);
    const int cx = (ctu % Q.Wc) * 32, cy = (ctu / Q.Wc) * 32, cw = Q.W >> 1;
    const size_t ps = (size_t)Q.W * Q.H * 3 / 2;
    const int16_t *ly = Q.lev + (size_t)pic * ps, *lcb = ly + (size_t)Q.W * Q.H, *lcr = lcb + (size_t)cw * (Q.H >> 1);
    const uint4 *ry = reinterpret_cast<const uint4 *>(ly + (size_t)(cy + lane) * Q.W + cx);  // 32 levels = 4 x 16 bytes, 8 per block column pair
    unsigned fy = 0;
    for (int j = 0; j < 4; j++) {
        const uint4 v = ry[j];
        fy |= ((v.x | v.y) != 0u ? 1u : 0u) << (2 * j);
        fy |= ((v.z | v.w) != 0u ? 1u : 0u) << (2 * j + 1);
    }
    const int16_t *pc = (lane < 16 ? lcb : lcr) + (size_t)((cy >> 1) + (lane & 15)) * cw + (cx >> 1);
    const uint4 *rc = reinterpret_cast<const uint4 *>(pc);
    unsigned fc = 0;
    for (int j = 0; j < 2; j++) {
        const uint4 v = rc[j];
        fc |= ((v.x | v.y) != 0u ? 1u : 0u) << (2 * j);
        fc |= ((v.z | v.w) != 0u ? 1u : 0u) << (2 * j + 1);
    }
    unsigned long long my = 0ull;
    for (int bx = 0; bx < 8; bx++) {
        const unsigned rows = __ballot_sync(0xffffffffu, (fy >> bx) & 1u);
        for (int by = 0; by < 8; by++)
            if ((rows >> (4 * by)) & 15u) my |= 1ull << (by * 8 + bx);
    }
    unsigned mcb = 0, mcr = 0;
    for (int bx = 0; bx < 4; bx++) {
        const unsigned rows = __ballot_sync(0xffffffffu, (fc >> bx) & 1u);
        for (int by = 0; by < 4; by++) {
            if ((rows >> (4 * by)) & 15u) mcb |= 1u << (by * 4 + bx);
            if ((rows >> (16 + 4 * by)) & 15u) mcr |= 1u << (by * 4 + bx);
        }
    }
    if (lane == 0) {
        NzMap m;
        m.y = my; m.cb = (unsigned short)mcb; m.cr = (unsigned short)mcr; m.pad = 0;
        Q.nzmap[gid] = m;
    }
}

#ifndef WB_SYN_MINB
#define WB_SYN_MINB 16  // resident blocks of 64 threads per SM the register allocation aims for (64 registers per thread)
#endif
extern "C" __global__ void __launch_bounds__(64, WB_SYN_MINB) wrenc_b200_syntax_kernel(SyntaxParams Q) {
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
    __align__(4) uint8_t pass1[1024], absl[1024];
    TuState ts;
    ts.qp_delta_coded = false;  // quantisation group = CTU (cu_qp_delta_subdiv 0, ctu_encoder.rs:305-310)
    ts.mts_dc_only = true;
    ts.mts_zero_out = true;
    const int cx = (ctu % Q.Wc) * 32, cy = (ctu / Q.Wc) * 32;
    NzMap nz;
#if WB_NZMAP
    nz = Q.nzmap[gid];
#else
    nz.y = 0; nz.cb = 0; nz.cr = 0; nz.pad = 0;
#endif
    code_ctu(S, SO, P, P.rec[ctu], nz, cx, cy, ts, pass1, absl);
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
// arithmetic coder (bool_coder.rs:136-296, 1073-1111) — the engine itself is cabac_engine.cuh
// ---------------------------------------------------------------------------------------------------------------
struct WarpEnv {  // what ce::code_batch needs: the context index of entry i of the batch, context words in shared memory
    unsigned cur;
    unsigned *ctx;
    __device__ __forceinline__ unsigned ci(int i) const { return __shfl_sync(0xffffffffu, cur, i) & 511u; }
    __device__ __forceinline__ unsigned load(unsigned c) const { return ctx[c]; }
    __device__ __forceinline__ void store(unsigned c, unsigned w) {
        __syncwarp();  // every lane has read the old word (its own load and the prefetch of the next bin) before anyone stores the (identical) new one
        ctx[c] = w;
    }
};

#ifndef WB_CABAC2
#define WB_CABAC2 2  // 2: two warps per picture (token records through shared memory); 1: one warp, token words by shuffle; 0: ce::code_batch (all sequential)
#endif

#if WB_CABAC2 == 2
__device__ __forceinline__ void nbar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
constexpr int CABAC_SUBS = 8;                         // 32-entry batches per slot hand-over
constexpr int CABAC_SLOT = CABAC_SUBS * 32 + 8;       // records per slot (+ the no-op padding of the last round + the range warp's look-ahead)
enum { BAR_REC_FULL = 1, BAR_REC_EMPTY = 3, BAR_OP_FULL = 5, BAR_OP_EMPTY = 7, BAR_TERM = 9 };  // + slot; 64 participants each (32 arrive, 32 wait)

// THREE warps per picture, one per stage of the data flow.  What is sequential by nature is only the interval width: the
// probability state a context-coded bin sees depends on the bin history of its context alone, and the code value (low) never
// feeds back into the widths.
//   warp 1, states: fetches the bin string 32 entries at a time (one coalesced load, double-buffered); lane i owns entry i;
//     match.any groups the lanes by context, the group's first lane takes the context word from shared memory and the adapted
//     word travels down the group by shuffles (as many rounds as the most frequent context of the batch has entries), the
//     group's last lane stores it back; every walked entry - a context-coded bin, or the start of a run of up to 8 bypass bins -
//     becomes one ready-to-use record (ce::TokRec), written densely into one of two shared-memory slots;
//   warp 0, range: walks the records; per record the dependent chain range -> LPS width -> one-shift renormalisation and nothing
//     else (ce::step_range); leaves what the code value needs (ce::LowOp: addend, shift, bypass term) in a second pair of slots;
//   warp 2, low: accumulates the code value in a 64-bit window, four records per round, resolves carries and stores the bytes
//     (ce::step_low / flush / finish_low; lane 0 stores).
// Warps 0 and 2 run every lane redundantly on identical registers.  The slots are handed over with named barriers (full /
// empty per slot and pair) every CABAC_SUBS batches, so each stage works one slot ahead of the next.
extern "C" __global__ void __launch_bounds__(96) wrenc_b200_cabac_kernel(SyntaxParams Q) {
    const int pic = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (pic >= Q.n_pics) return;
    const int nctu = Q.Wc * Q.Hc;
    __shared__ unsigned ctx[CTX_TOTAL];
    __shared__ __align__(16) ce::TokRec rec[2][CABAC_SLOT];
    __shared__ __align__(8) ce::LowOp ops[2][CABAC_SLOT];
    __shared__ int rec_count[2], op_count[2];
    __shared__ unsigned term_add;
    // the CTUs' bin strings lie back to back in the arena (exclusive scan of the counts), in raster order
    const size_t g0 = (size_t)pic * nctu;
    const unsigned long long beg = Q.bin_offset[g0];
    const unsigned long long end = Q.bin_offset[g0 + nctu - 1] + (unsigned long long)Q.bin_count[g0 + nctu - 1];
    if (end > Q.bins_cap) {  // this picture's strings did not fit the arena (sized from earlier batches): nothing was written for it
        if (threadIdx.x == 0) Q.out_len[pic] = -2;
        return;
    }
    const uint16_t *b = Q.bins + beg;
    const long long total = (long long)(end - beg);
    if (warp == 1) {
        for (int i = lane; i < CTX_TOTAL; i += 32) ctx[i] = ce::ctx_init_word(kCabacInitValue[i], kCabacShiftIdx[i], Q.qp);  // init_ctx_table (bool_coder.rs:1073-1093)
        __syncwarp();
        unsigned nxt = lane < total ? b[lane] : 0u;
        int slot = 0;
        for (long long sbase = 0; sbase < total; sbase += 32 * CABAC_SUBS, slot ^= 1) {
            int fill = 0;  // records of this slot so far
            for (long long base = sbase; base < min(total, sbase + 32 * CABAC_SUBS); base += 32) {
                const unsigned e = nxt;
                const long long pf = base + 32 + lane;
                nxt = pf < total ? b[pf] : 0u;  // prefetch the next 32 entries
                const int cnt = (int)min(32ll, total - base);
                const unsigned bypm = __ballot_sync(0xffffffffu, (e & 1024u) != 0u);  // entries at and above cnt are 0
                const unsigned binm = __ballot_sync(0xffffffffu, (e & 512u) != 0u);
                const bool is_ctx = lane < cnt && !(e & 1024u);
                const unsigned bin = (e >> 9) & 1u, ci = e & 511u;
                // lanes of one context, in entry order: rank within the group, predecessor, last of the group
                const unsigned peers = __match_any_sync(0xffffffffu, is_ctx ? ci : 512u + (unsigned)lane);
                const unsigned before = peers & ((1u << lane) - 1u);
                const int rank = __popc(before);
                const int pred = before ? 31 - __clz((int)before) : lane;
                const int rounds = (int)__reduce_max_sync(0xffffffffu, (unsigned)rank);
                unsigned w = ctx[is_ctx ? ci : 0u];
                unsigned adapted = ce::adapt(w, bin);
                for (int r = 1; r <= rounds; r++) {  // after round r the lanes of rank <= r hold the word their bin sees
                    const unsigned t = __shfl_sync(0xffffffffu, adapted, pred);
                    if (rank == r) w = t;
                    adapted = ce::adapt(w, bin);
                }
                __syncwarp();  // every lane has read its group's word before the group's last lane replaces it
                if (is_ctx && (peers >> lane) == 1u) ctx[ci] = adapted;
                __syncwarp();
                const bool walked = lane < cnt && ce::tok_walked(bypm, lane);
                const unsigned wm = __ballot_sync(0xffffffffu, walked);
                const int idx = fill + __popc(wm & ((1u << lane) - 1u));
                if (base == sbase && sbase >= 2ll * 32 * CABAC_SUBS) nbar_sync(BAR_REC_EMPTY + slot);  // the range warp is done with this slot's previous records
                if (walked) rec[slot][idx] = (e & 1024u) ? ce::rec_bypass(bypm, binm, lane) : ce::rec_ctx(w, bin);
                fill += __popc(wm);
            }
            if (lane < 3) rec[slot][fill + lane] = ce::rec_nop();  // the walkers take four records per round
            if (lane == 0) rec_count[slot] = fill;
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            nbar_arrive(BAR_REC_FULL + slot);
        }
    } else if (warp == 0) {
        unsigned range = 510u;
        int slot = 0;
        for (long long sbase = 0; sbase < total; sbase += 32 * CABAC_SUBS, slot ^= 1) {
            nbar_sync(BAR_REC_FULL + slot);
            if (sbase >= 2ll * 32 * CABAC_SUBS) nbar_sync(BAR_OP_EMPTY + slot);  // the low warp is done with this slot's previous records
            const ce::TokRec *rs = rec[slot];
            ce::LowOp *os = ops[slot];
            const int count = rec_count[slot];
            // four records per round, the next round's records loaded before this round's chain (the slot is padded with no-ops
            // and four more entries, so the look-ahead never leaves it)
            ce::TokRec cur[4], nxt4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) cur[u] = rs[u];
            for (int j = 0; j < count; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) nxt4[u] = rs[j + 4 + u];
                ce::LowOp o[4];
#pragma unroll
                for (int u = 0; u < 4; u++) o[u] = ce::step_range(range, cur[u]);
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < 4; u++) os[j + u] = o[u];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) cur[u] = nxt4[u];
            }
            if (lane == 0) op_count[slot] = count;
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            nbar_arrive(BAR_OP_FULL + slot);
            nbar_arrive(BAR_REC_EMPTY + slot);
        }
        if (lane == 0) term_add = range - 2u;  // end_of_slice_one_bit = 1 (bool_coder.rs:218-235)
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        nbar_arrive(BAR_TERM);
    } else {
        ce::Arith E;
        E.init(lane == 0 ? Q.out + (size_t)pic * Q.out_cap : nullptr, Q.out_cap);  // lanes other than 0 only count
        int slot = 0;
        for (long long sbase = 0; sbase < total; sbase += 32 * CABAC_SUBS, slot ^= 1) {
            nbar_sync(BAR_OP_FULL + slot);
            const ce::LowOp *os = ops[slot];
            const int count = op_count[slot];
            for (int j = 0; j < count; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) ce::step_low(E, os[j + u]);
                E.flush();
            }
            nbar_arrive(BAR_OP_EMPTY + slot);
        }
        nbar_sync(BAR_TERM);
        // flush, stop bit, byte alignment with zeros (slice_encoder.rs:419)
        const size_t n = E.finish_low(term_add);
        if (lane == 0) Q.out_len[pic] = n > Q.out_cap ? -1 : (int)n;
    }
}
#else
// One WARP per picture; the bin strings are fetched 32 entries at a time with one coalesced load, double-buffered, and two
// ballots turn a batch into a bypass mask and a bin mask.  What is sequential by nature is only the interval (range / low) update.
// The probability state a context-coded bin sees depends on the bin history of its context alone, so it is resolved by the
// lanes in parallel: lane i owns entry i, match.any groups the lanes by context, the group's first lane takes the context word
// from shared memory and the adapted word travels down the group by shuffles (as many rounds as the most frequent context of
// the batch has entries), the group's last lane stores it back.  Every lane then packs its entry into one token word - LPS
// probability index and MPS flag, or a whole run of up to 8 bypass bins - and the sequential part (every lane redundantly on
// identical registers; lane 0 alone stores the output bytes) walks the tokens: per context-coded bin one shuffle and the chain
// range -> LPS width -> one-shift renormalisation (no per-bit loop, no outstanding-bit counter: cabac_engine.cuh).
extern "C" __global__ void __launch_bounds__(32) wrenc_b200_cabac_kernel(SyntaxParams Q) {
    const int pic = blockIdx.x, lane = threadIdx.x;
    if (pic >= Q.n_pics) return;
    const int nctu = Q.Wc * Q.Hc;
    __shared__ unsigned ctx[CTX_TOTAL];
    for (int i = lane; i < CTX_TOTAL; i += 32) ctx[i] = ce::ctx_init_word(kCabacInitValue[i], kCabacShiftIdx[i], Q.qp);  // init_ctx_table (bool_coder.rs:1073-1093)
    __syncwarp();
    ce::Arith E;
    E.init(lane == 0 ? Q.out + (size_t)pic * Q.out_cap : nullptr, Q.out_cap);  // lanes other than 0 only count
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
#if !WB_CABAC2
    WarpEnv env;
    env.ctx = ctx;
#endif
    for (long long base = 0; base < total; base += 32) {
        const unsigned e = nxt;
        const long long pf = base + 32 + lane;
        nxt = pf < total ? b[pf] : 0u;  // prefetch the next 32 entries while this batch is coded
        const int cnt = (int)min(32ll, total - base);
        const unsigned bypm = __ballot_sync(0xffffffffu, (e & 1024u) != 0u);  // entries at and above cnt are 0
        const unsigned binm = __ballot_sync(0xffffffffu, (e & 512u) != 0u);
#if WB_CABAC2
        const bool is_ctx = lane < cnt && !(e & 1024u);
        const unsigned bin = (e >> 9) & 1u, ci = e & 511u;
        // lanes of one context, in entry order: rank within the group, predecessor, last of the group
        const unsigned peers = __match_any_sync(0xffffffffu, is_ctx ? ci : 512u + (unsigned)lane);
        const unsigned before = peers & ((1u << lane) - 1u);
        const int rank = __popc(before);
        const int pred = before ? 31 - __clz((int)before) : lane;
        const int rounds = (int)__reduce_max_sync(0xffffffffu, (unsigned)rank);
        unsigned w = ctx[is_ctx ? ci : 0u];
        unsigned adapted = ce::adapt(w, bin);
        for (int r = 1; r <= rounds; r++) {  // after round r the lanes of rank <= r hold the word their bin sees
            const unsigned t = __shfl_sync(0xffffffffu, adapted, pred);
            if (rank == r) w = t;
            adapted = ce::adapt(w, bin);
        }
        __syncwarp();  // every lane has read its group's word before the group's last lane replaces it
        if (is_ctx && (peers >> lane) == 1u) ctx[ci] = adapted;
        __syncwarp();
        const unsigned tok = (e & 1024u) ? ce::token_bypass(bypm, binm, lane) : ce::token_ctx(w, bin, lane);
        ce::run_tokens(E, [&](int i) { return __shfl_sync(0xffffffffu, tok, i); }, cnt);
#else
        env.cur = e;
        ce::code_batch(E, env, bypm, binm, cnt);
#endif
    }
    // end_of_slice_one_bit = 1 (bool_coder.rs:218-235), then byte alignment with zeros (slice_encoder.rs:419)
    const size_t n = E.finish();
    if (lane == 0) Q.out_len[pic] = n > Q.out_cap ? -1 : (int)n;
}
#endif

int syntax_first_pass_kernels() { return 1 + WB_NZMAP; }
cudaError_t launch_syntax(const SyntaxParams &Q, cudaStream_t stream) {  // Q.bins == nullptr: non-zero map + counting pass
    const long long total = (long long)Q.n_pics * Q.Wc * Q.Hc;
#if WB_NZMAP
    if (!Q.bins) wrenc_b200_nzmap_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, stream>>>(Q);
#endif
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
    wrenc_b200_cabac_kernel<<<Q.n_pics, WB_CABAC2 == 2 ? 96 : 32, 0, stream>>>(Q);
    return cudaGetLastError();
}

}  // namespace wb
