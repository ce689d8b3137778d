// cabac_engine.cuh — the arithmetic coder of the slice coder (bool_coder.rs:136-296), restated for a latency-bound device loop
// and written so that it compiles for the device (wrenc_b200_cabac_kernel) and for the host (tests/host/cabac_engine_host_test.cpp
// runs it against the oracle's bit-by-bit engine on random and real bin strings without a GPU).
//
// The reference keeps a 10-bit `offset`, emits one bit per renormalisation step and counts outstanding bits (put-bit form).  The
// coded string depends only on the sequence of interval updates, so the same bytes come out of the register form used here:
//   * `low` is a window of the code value (64 bits wide: up to 32 bits may be added before bytes have to leave), `bits_left`
//     counts the free bits below bit 32; a renormalisation is ONE shift by n = clz(range) - 23 bits (no loop), and whole bytes
//     leave the window in flush() / write_out() until at least 12 bits are free again;
//   * a carry out of the window is resolved against one buffered byte and a count of 0xff bytes behind it (write_out / finish);
//   * a run of up to 8 bypass bins is one update  low = (low << k) + range * value  (k single steps are exactly that);
//   * the probability pair of a context and its two adaptation shifts are ONE 32-bit word
//         q0 (10 bits) | q1 (14 bits) << 10 | shift0 << 24 | shift1 << 28        (bool_coder.rs:136-154, 1073-1093)
//     so that a context-coded bin costs one load and one store.
// The dropped first bit of the reference (flush_cabac_bin) is the top bit of its 10-bit offset, which is always 0: the register form
// starts with 9 bits in the window (bits_left = 23) and never holds it.  end_of_slice_one_bit, the two trailing bits with the
// forced stop bit and the zero padding (bool_coder.rs:218-235, slice_encoder.rs:419) are finish().
// Three forms of the same coder, each checked on the host: code_batch (all sequential: one load and one store per context-coded
// bin), run_tokens (every entry packed into one word by its own lane, one branch-free step per word) and the record list of the
// shipped kernel (TokRec / step_range / step_low: the interval-width side and the code-value side of a step, run by different warps).
#pragma once
#include <stdint.h>

#define CE_UNLIKELY(x) __builtin_expect(!!(x), 0)
#if defined(__CUDACC__)
#define CE_HD __host__ __device__ __forceinline__
#else
#define CE_HD inline
#endif

namespace ce {

#if defined(__CUDA_ARCH__)
CE_HD int clz32(unsigned v) { return __clz((int)v); }
CE_HD int ffs32(unsigned v) { return __ffs((int)v); }
CE_HD unsigned brev32(unsigned v) { return __brev(v); }
#else
CE_HD int clz32(unsigned v) { return v ? __builtin_clz(v) : 32; }
CE_HD int ffs32(unsigned v) { return __builtin_ffs((int)v); }
CE_HD unsigned brev32(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
#endif

// context word: init_ctx_table (bool_coder.rs:1073-1093)
CE_HD unsigned ctx_init_word(int init_value, int shift_idx, int slice_qp) {
    const int m = (init_value >> 3) - 4, nn = (init_value & 7) * 18 + 1;
    const int qp = slice_qp < 0 ? 0 : (slice_qp > 63 ? 63 : slice_qp);
    int pre = ((m * (qp - 16)) >> 1) + nn;
    pre = pre < 1 ? 1 : (pre > 127 ? 127 : pre);
    const unsigned s0 = (unsigned)((shift_idx >> 2) + 2), s1 = (unsigned)((shift_idx & 3) + 3) + s0;
    return (unsigned)(pre << 3) | ((unsigned)(pre << 7) << 10) | (s0 << 24) | (s1 << 28);
}

// probability token of a bin coded with context word w: qlps | is_mps << 5
CE_HD unsigned ctx_token(unsigned w, unsigned bin) {
    const unsigned ps = ((w >> 10) & 16383u) + 16u * (w & 1023u);
    const unsigned mps = ps >> 14;
    return ((mps ? 32767u - ps : ps) >> 9) | (bin == mps ? 32u : 0u);
}
// the context word after coding `bin` (bool_coder.rs:136-154: two estimators with their own adaptation shifts)
CE_HD unsigned adapt(unsigned w, unsigned bin) {
    const unsigned q0 = w & 1023u, q1 = (w >> 10) & 16383u, s0 = (w >> 24) & 7u, s1 = w >> 28;
    const unsigned n0 = q0 - (q0 >> s0) + (bin ? 1023u >> s0 : 0u);
    const unsigned n1 = q1 - (q1 >> s1) + (bin ? 16383u >> s1 : 0u);
    return (w & 0xff000000u) | n0 | (n1 << 10);
}

struct Arith {
    unsigned long long low;  // 64-bit window: up to 32 bits may be added before whole bytes have to leave it (flush)
    unsigned range, buf_byte;
    int bits_left, n_buf;
    uint8_t *p;  // nullptr: count only
    size_t n, cap;

    CE_HD void init(uint8_t *out, size_t out_cap) {
        low = 0; range = 510; buf_byte = 0xff; bits_left = 23; n_buf = 0;
        p = out; n = 0; cap = out_cap;
    }
    CE_HD void emit(unsigned b) {
        if (p && n < cap) p[n] = (uint8_t)b;
        n++;
    }
    // the top byte of the window leaves it; a carry (bit 8 of lead) goes into the buffered byte and the 0xff run behind it
    CE_HD void write_out() {
        const unsigned lead = (unsigned)(low >> (24 - bits_left));
        bits_left += 8;
        low &= 0xffffffffffffffffull >> (32 + bits_left);
        if (lead == 0xffu) n_buf++;
        else if (n_buf > 0) {
            const unsigned carry = lead >> 8;
            emit(buf_byte + carry);
            buf_byte = lead & 0xffu;
            const unsigned fill = (0xffu + carry) & 0xffu;
#pragma unroll 1
            while (n_buf > 1) { emit(fill); n_buf--; }
        } else {
            n_buf = 1;
            buf_byte = lead;
        }
    }
    // interval update of a context-coded bin (bool_coder.rs:254-296) from the bin's probability token: the LPS probability index
    // qlps = (p < 0.5 ? p : 1 - p) >> 9 of the context word before the bin, and whether the bin is the more probable symbol
    CE_HD void decision_tok(unsigned qlps, bool is_mps) {
        const unsigned lps = (((range >> 5) * qlps) >> 1) + 4u;
        const unsigned rmps = range - lps;
        const unsigned r = is_mps ? rmps : lps;
        const unsigned add = is_mps ? 0u : rmps;
        const int nsh = clz32(r) - 23;  // r < 512; the MPS interval is at least 128 wide, the LPS interval at least 4
        low = (low + add) << nsh;
        range = r << nsh;
        bits_left -= nsh;
        flush();
    }
    // whole bytes out of the window until at least 12 bits are free below bit 32 again (bits_left >= -20 on entry)
    CE_HD void flush() {
#pragma unroll 1
        while (bits_left < 12) write_out();
    }
    // context-coded bin; returns the adapted context word
    CE_HD unsigned decision(unsigned w, unsigned bin) {
        const unsigned t = ctx_token(w, bin);
        decision_tok(t & 31u, (t & 32u) != 0u);
        return adapt(w, bin);
    }
    // k <= 8 bypass bins, first bin in the most significant bit of v (bool_coder.rs:202-216, k times)
    CE_HD void bypass(unsigned v, int k) {
        low = (low << k) + range * v;
        bits_left -= k;
        flush();
    }
    // end_of_slice_one_bit = 1, flush, stop bit, zero padding to the byte boundary; returns the byte count
    CE_HD size_t finish() { return finish_low(range - 2u); }
    // the same from the low side alone: term_add = range - 2 of the interval before end_of_slice_one_bit
    CE_HD size_t finish_low(unsigned term_add) {
        low = (low + term_add) << 7;
        range = 2u << 7;
        bits_left -= 7;
        flush();
        if (low >> (32 - bits_left)) {
            emit(buf_byte + 1);
#pragma unroll 1
            while (n_buf > 1) { emit(0x00u); n_buf--; }
            low -= 1ull << (32 - bits_left);
        } else {
            if (n_buf > 0) emit(buf_byte);
#pragma unroll 1
            while (n_buf > 1) { emit(0xffu); n_buf--; }
        }
        int k = 24 - bits_left;                         // bits of the code value still in the window (above bit 8)
        unsigned v = ((unsigned)(low >> 8) << 1) | 1u;            // + the stop bit the reference forces into its last trailing bit
        k += 1;
        const int pad = (8 - (k & 7)) & 7;
        v <<= pad;
        k += pad;
#pragma unroll 1
        for (int sh = k - 8; sh >= 0; sh -= 8) emit((v >> sh) & 0xffu);
        return n;
    }
};

// One batch of up to 32 bin-string entries (16 bits each: context index | bin << 9 | bypass << 10), given as two masks over the
// batch (bypass flags, bin values; bits at and above cnt are 0) and an environment that hands out the context index of entry i
// and loads / stores context words.  The state of the next context-coded bin is fetched before the current bin is coded (and
// replaced by the adapted word when it is the same context), so that its latency is off the dependent chain.
template <class Env>
CE_HD void code_batch(Arith &E, Env &env, unsigned bypm, unsigned binm, int cnt) {
    const unsigned valid = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
    const unsigned ctxm = ~bypm & valid;
    int i = 0;
    bool have = false;
    unsigned nci = 0, nw = 0;
    while (i < cnt) {
        if ((bypm >> i) & 1u) {
            const unsigned run = ~(bypm >> i);
            int k = run ? ffs32(run) - 1 : 32;
            k = k > 8 ? 8 : k;
            const unsigned v = brev32(binm >> i) >> (32 - k);
            E.bypass(v, k);
            i += k;
            continue;
        }
        unsigned ci, w;
        if (have) { ci = nci; w = nw; }
        else { ci = env.ci(i); w = env.load(ci); }
        const unsigned rest = (ctxm >> i) >> 1;
        have = rest != 0u;
        if (have) {
            nci = env.ci(i + ffs32(rest));
            nw = env.load(nci);
        }
        const unsigned neww = E.decision(w, (binm >> i) & 1u);
        env.store(ci, neww);
        if (have && nci == ci) nw = neww;
        i++;
    }
}

// ---- the same batch as a token program (wrenc_b200_cabac_kernel): every entry of the batch is turned into one word by its own
// lane, in parallel, and the sequential part only walks the words with ONE branch-free update per word:
//     lps  = (((range >> 5) * qlps) >> 1) + c4          c4 = 4 for a context-coded bin, 0 for a bypass run (qlps = 0: lps = 0)
//     r    = mps ? range - lps : lps                    bypass runs carry mps = 1: r = range, nothing is added to low
//     n    = clz(r) - 23                                0 for a bypass run
//     low  = ((low + (mps ? 0 : range - lps)) << (n + k)) + range * v        k, v = length and value of the bypass run, 0 otherwise
//     range = r << n
// word layout: qlps (bits 0-4) | mps << 5 | bypass << 6 | k << 7 | v << 11 | next << 19   (next = position of the next word to
// walk: own position + 1, or + k for a bypass run; a word with k = 0, v = 0, bypass = mps = 1 changes nothing: TOK_NOP)
enum : unsigned { TOK_BYPASS = 64u, TOK_NOP = 32u | 64u | (32u << 19) };
CE_HD unsigned token_ctx(unsigned w, unsigned bin, int pos) { return ctx_token(w, bin) | ((unsigned)(pos + 1) << 19); }
CE_HD unsigned token_bypass(unsigned bypm, unsigned binm, int pos) {  // bit pos of bypm is set
    const unsigned run = ~(bypm >> pos);
    int k = run ? ffs32(run) - 1 : 32;
    k = k > 8 ? 8 : k;
    const unsigned v = brev32(binm >> pos) >> (32 - k);
    return 32u | (unsigned)TOK_BYPASS | ((unsigned)k << 7) | (v << 11) | ((unsigned)(pos + k) << 19);
}
CE_HD void step(Arith &E, unsigned t) {
    const unsigned qlps = t & 31u, c4 = ((t >> 4) & 4u) ^ 4u, k = (t >> 7) & 15u, v = (t >> 11) & 255u;
    const bool mps = (t & 32u) != 0u;
    const unsigned lps = (((E.range >> 5) * qlps) >> 1) + c4;
    const unsigned rmps = E.range - lps;
    const unsigned r = mps ? rmps : lps;
    const unsigned add = mps ? 0u : rmps;
    const int nsh = clz32(r) - 23;
    const int sh = nsh + (int)k;
    E.low = ((E.low + add) << sh) + (unsigned long long)(E.range * v);
    E.range = r << nsh;
    E.bits_left -= sh;
}
// get(i): the word of entry i (i < 32).  The word of the next entry is fetched before the current one is coded; the walk is
// unrolled by four with no branch inside (words past the end are TOK_NOP; four words add at most 32 bits to the 64-bit window,
// whole bytes leave it once per round).
template <class Get>
CE_HD void run_tokens(Arith &E, Get get, int cnt) {
    int i = 0;
    unsigned t = cnt > 0 ? get(0) : (unsigned)TOK_NOP;
    while (i < cnt) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int ni = (int)(t >> 19);
            unsigned tn = get(ni & 31);
            tn = ni >= cnt ? (unsigned)TOK_NOP : tn;
            step(E, t);
            t = tn;
            i = ni;
        }
        E.flush();
    }
}

// ---- the token program as a COMPACTED record list (wrenc_b200_cabac_kernel, two warps per picture): the producer warp writes one
// ready-to-use record per walked entry (no packing, nothing to decode) into shared memory, densely, so that the consumer walks
// consecutive records: four branch-free steps per round with every load address known in advance.
struct TokRec {
    unsigned qlps, c4, k, v;  // step() operands (see above)
    unsigned mps, pad0, pad1, pad2;
};
// entry pos of a batch is walked when it is context-coded or starts a (sub-)run of 8 bypass bins: bins 0, 8, 16, 24 of a run
CE_HD bool tok_walked(unsigned bypm, int pos) {
    if (!((bypm >> pos) & 1u)) return true;
    const unsigned below = ~bypm & ((1u << pos) - 1u);  // context-coded entries before pos
    const int start = below ? 32 - clz32(below) : 0;    // first entry of the bypass run that holds pos
    return ((pos - start) & 7) == 0;
}
CE_HD TokRec rec_nop() { TokRec t = {0u, 0u, 0u, 0u, 1u, 0u, 0u, 0u}; return t; }
CE_HD TokRec rec_ctx(unsigned w, unsigned bin) {
    const unsigned c = ctx_token(w, bin);
    TokRec t = {c & 31u, 4u, 0u, 0u, (c >> 5) & 1u, 0u, 0u, 0u};
    return t;
}
CE_HD TokRec rec_bypass(unsigned bypm, unsigned binm, int pos) {
    const unsigned b = token_bypass(bypm, binm, pos);
    TokRec t = {0u, 0u, (b >> 7) & 15u, (b >> 11) & 255u, 1u, 0u, 0u, 0u};
    return t;
}
// One step, split where the data flow splits: the RANGE side is the only sequential dependency between steps (range -> LPS width ->
// renormalisation shift -> range); what it leaves for the LOW side is  add | sh << 16  and  range * v  (LowOp), and the low side
// (window, carries, bytes) never feeds back.  The kernel runs the two sides in different warps.
struct LowOp {
    unsigned add_sh, rv;
};
CE_HD LowOp step_range(unsigned &range, const TokRec &t) {
    const unsigned lps = (((range >> 5) * t.qlps) >> 1) + t.c4;
    const unsigned rmps = range - lps;
    const bool mps = t.mps != 0u;
    const unsigned r = mps ? rmps : lps;
    const unsigned add = mps ? 0u : rmps;
    const int nsh = clz32(r) - 23;
    LowOp o = {add | ((unsigned)(nsh + (int)t.k) << 16), range * t.v};
    range = r << nsh;
    return o;
}
CE_HD void step_low(Arith &E, const LowOp &o) {
    const int sh = (int)(o.add_sh >> 16);
    E.low = ((E.low + (o.add_sh & 0xffffu)) << sh) + (unsigned long long)o.rv;
    E.bits_left -= sh;
}
CE_HD void step_rec(Arith &E, const TokRec &t) { step_low(E, step_range(E.range, t)); }
// get(j): record j of the batch's list, padded with rec_nop() to a multiple of four
template <class Get>
CE_HD void walk_records(Arith &E, Get get, int count) {
    for (int j = 0; j < count; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) step_rec(E, get(j + u));
        E.flush();
    }
}

}  // namespace ce
