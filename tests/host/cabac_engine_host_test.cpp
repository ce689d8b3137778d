// Host check of wrenc_b200/csrc/cabac_engine.cuh (the register-form arithmetic coder of wrenc_b200_cabac_kernel: one-shift
// renormalisation, byte-wise output with carry resolution, merged bypass runs, packed context words, prefetched context state)
// against the CPU oracle's bit-by-bit engine (oracle/wrenc_oracle_cabac.cpp, bool_coder.rs:136-296), on random bin strings built to
// provoke long carry / 0xff runs and on the real bin strings of searched pictures.  Test infrastructure: built and run by
// tests/test_cabac_engine_host.py (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../oracle/wrenc_oracle.hpp"
#include "../../oracle/cabac_tables.inc"
#include "../../wrenc_b200/csrc/cabac_engine.cuh"

struct HostEnv {
    const uint16_t *e;  // the batch
    unsigned *ctx;
    unsigned ci(int i) const { return e[i] & 511u; }
    unsigned load(unsigned c) const { return ctx[c]; }
    void store(unsigned c, unsigned w) { ctx[c] = w; }
};

// the token-program form: every entry becomes one word (here: in entry order, which is what the kernel's warp-parallel state
// propagation has to reproduce), then the sequential walk
static std::vector<uint8_t> code_tokens(int qp, const std::vector<uint16_t> &bins) {
    std::vector<unsigned> ctx(CTX_TOTAL);
    for (int i = 0; i < CTX_TOTAL; i++) ctx[i] = ce::ctx_init_word(kCabacInitValue[i], kCabacShiftIdx[i], qp);
    std::vector<uint8_t> out(bins.size() / 4 + 64);
    ce::Arith E;
    E.init(out.data(), out.size());
    for (size_t base = 0; base < bins.size(); base += 32) {
        const int cnt = (int)std::min<size_t>(32, bins.size() - base);
        unsigned bypm = 0, binm = 0, tok[32] = {};
        for (int i = 0; i < cnt; i++) {
            bypm |= (unsigned)((bins[base + i] >> 10) & 1) << i;
            binm |= (unsigned)((bins[base + i] >> 9) & 1) << i;
        }
        for (int i = 0; i < cnt; i++) {
            const unsigned e = bins[base + i], bin = (e >> 9) & 1u;
            if (e & 1024u) tok[i] = ce::token_bypass(bypm, binm, i);
            else {
                tok[i] = ce::token_ctx(ctx[e & 511u], bin, i);
                ctx[e & 511u] = ce::adapt(ctx[e & 511u], bin);
            }
        }
        ce::run_tokens(E, [&](int i) { return tok[i]; }, cnt);
    }
    const size_t n = E.finish();
    if (n > out.size()) { fprintf(stderr, "host buffer too small\n"); exit(2); }
    out.resize(n);
    return out;
}

// the compacted record list of the two-warp kernel: walked entries only, padded with no-ops to a multiple of four
static std::vector<uint8_t> code_records(int qp, const std::vector<uint16_t> &bins) {
    std::vector<unsigned> ctx(CTX_TOTAL);
    for (int i = 0; i < CTX_TOTAL; i++) ctx[i] = ce::ctx_init_word(kCabacInitValue[i], kCabacShiftIdx[i], qp);
    std::vector<uint8_t> out(bins.size() / 4 + 64);
    ce::Arith E;
    E.init(out.data(), out.size());
    for (size_t base = 0; base < bins.size(); base += 32) {
        const int cnt = (int)std::min<size_t>(32, bins.size() - base);
        unsigned bypm = 0, binm = 0;
        for (int i = 0; i < cnt; i++) {
            bypm |= (unsigned)((bins[base + i] >> 10) & 1) << i;
            binm |= (unsigned)((bins[base + i] >> 9) & 1) << i;
        }
        ce::TokRec rec[36];
        int count = 0;
        for (int i = 0; i < cnt; i++) {
            const unsigned e = bins[base + i], bin = (e >> 9) & 1u;
            const bool walked = ce::tok_walked(bypm, i);
            if (e & 1024u) {
                if (walked) rec[count++] = ce::rec_bypass(bypm, binm, i);
            } else {
                if (!walked) { fprintf(stderr, "context-coded entry not walked\n"); exit(2); }
                rec[count++] = ce::rec_ctx(ctx[e & 511u], bin);
                ctx[e & 511u] = ce::adapt(ctx[e & 511u], bin);
            }
        }
        for (int i = 0; i < 3; i++) rec[count + i] = ce::rec_nop();
        ce::walk_records(E, [&](int j) { return rec[j]; }, count);
    }
    const size_t n = E.finish();
    if (n > out.size()) { fprintf(stderr, "host buffer too small\n"); exit(2); }
    out.resize(n);
    return out;
}

static std::vector<uint8_t> code_fast(int qp, const std::vector<uint16_t> &bins) {
    std::vector<unsigned> ctx(CTX_TOTAL);
    for (int i = 0; i < CTX_TOTAL; i++) ctx[i] = ce::ctx_init_word(kCabacInitValue[i], kCabacShiftIdx[i], qp);
    std::vector<uint8_t> out(bins.size() / 4 + 64);
    ce::Arith E;
    E.init(out.data(), out.size());
    for (size_t base = 0; base < bins.size(); base += 32) {
        const int cnt = (int)std::min<size_t>(32, bins.size() - base);
        unsigned bypm = 0, binm = 0;
        for (int i = 0; i < cnt; i++) {
            bypm |= (unsigned)((bins[base + i] >> 10) & 1) << i;
            binm |= (unsigned)((bins[base + i] >> 9) & 1) << i;
        }
        HostEnv env{bins.data() + base, ctx.data()};
        ce::code_batch(E, env, bypm, binm, cnt);
    }
    const size_t n = E.finish();
    if (n > out.size()) { fprintf(stderr, "host buffer too small\n"); exit(2); }
    out.resize(n);
    return out;
}

static long g_checked = 0, g_bad = 0, g_bins = 0;
static void check(int qp, const std::vector<uint16_t> &bins, const char *what) {
    const std::vector<uint8_t> want = wo::code_bin_string(qp, bins.data(), bins.size());
    const std::vector<uint8_t> got = code_fast(qp, bins);
    g_checked++;
    g_bins += (long)bins.size();
    if (want != code_tokens(qp, bins)) {
        if (g_bad < 5) fprintf(stderr, "MISMATCH (token program) %s: %zu bins\n", what, bins.size());
        g_bad++;
    }
    if (want != code_records(qp, bins)) {
        if (g_bad < 5) fprintf(stderr, "MISMATCH (record list) %s: %zu bins\n", what, bins.size());
        g_bad++;
    }
    if (want != got) {
        if (g_bad < 5) fprintf(stderr, "MISMATCH %s: %zu bins, want %zu bytes, got %zu\n", what, bins.size(), want.size(), got.size());
        g_bad++;
    }
}

int main(int argc, char **argv) {
    const int seeds = argc > 1 ? atoi(argv[1]) : 400;
    // ---- random strings
    for (int seed = 0; seed < seeds; seed++) {
        std::mt19937 rng(777u + (unsigned)seed);
        auto rnd = [&](int n) { return (int)(rng() % (unsigned)n); };
        const int qp = seed % 3 == 0 ? rnd(64) : 32;
        const int style = seed % 8;
        const int len = style == 7 ? rnd(40) : 1 + rnd(seed % 5 == 0 ? 60000 : 4000);
        std::vector<uint16_t> bins;
        bins.reserve(len + 64);
        // per-context bias: some contexts almost always 1 / 0 (probability states run to their ends: long MPS runs, small LPS intervals)
        int bias[CTX_TOTAL];
        for (int i = 0; i < CTX_TOTAL; i++) bias[i] = style == 1 ? (rnd(2) ? 1000 : 0) : (style == 2 ? 500 : rnd(1001));
        const int nctx = style == 3 ? 2 : CTX_TOTAL;
        while ((int)bins.size() < len) {
            const int kind = rnd(100);
            if (style == 4 || kind < 25) {  // bypass run; style 4: bypass only, mostly ones (carries ripple through 0xff runs)
                int run = 1 + rnd(kind < 5 ? 70 : 12);
                const int ones = style == 4 ? 97 : (style == 5 ? 3 : rnd(101));
                while (run-- > 0) bins.push_back((uint16_t)(1024u | ((unsigned)(rnd(100) < ones) << 9)));
            } else {
                int run = 1 + rnd(style == 6 ? 40 : 6);  // style 6: long runs on one context (state forwarding across entries)
                const int c = rnd(nctx);
                while (run-- > 0) {
                    const int cc = style == 6 ? c : rnd(nctx);
                    bins.push_back((uint16_t)((unsigned)cc | ((unsigned)(rnd(1000) < bias[cc]) << 9)));
                }
            }
        }
        bins.resize(len);
        check(qp, bins, "random");
    }
    // ---- many short strings ending in a run of bypass ones / LPS-heavy tails: a carry resolved in finish() behind a 0xff run
    for (int seed = 0; seed < seeds * 100; seed++) {
        std::mt19937 rng(31337u + (unsigned)seed);
        auto rnd = [&](int n) { return (int)(rng() % (unsigned)n); };
        std::vector<uint16_t> bins;
        const int head = rnd(60), tail = rnd(60);
        for (int i = 0; i < head; i++) bins.push_back((uint16_t)((unsigned)rnd(CTX_TOTAL) | ((unsigned)rnd(2) << 9)));
        for (int i = 0; i < tail; i++) bins.push_back((uint16_t)(1024u | ((unsigned)(rnd(100) < 95) << 9)));
        if (seed & 1) bins.push_back((uint16_t)((unsigned)rnd(CTX_TOTAL) | ((unsigned)rnd(2) << 9)));
        check(20 + seed % 30, bins, "short");
    }
    {  // empty string, single entries
        check(32, {}, "empty");
        for (unsigned e : {0u, 512u, 1024u, 1536u, 252u, 252u | 512u}) check(27, {(uint16_t)e}, "single");
    }
    // ---- real strings: searched pictures (synthetic content and noise, three QPs), the engine must reproduce slice_data()
    const int W = 96, H = 64;
    for (int t = 0; t < 4; t++) {
        const int qp = t == 0 ? 32 : (t == 1 ? 22 : (t == 2 ? 37 : 12));
        std::mt19937 rng(99u + (unsigned)t);
        std::vector<uint8_t> y(W * H), cb(W * H / 4), cr(W * H / 4);
        for (int j = 0; j < H; j++)
            for (int i = 0; i < W; i++) y[j * W + i] = (uint8_t)(t == 3 ? rng() : (128 + 60 * ((i / 7 + j / 5 + t) & 1) + (int)(rng() % 9) + ((i * j) >> 6)));
        for (auto &v : cb) v = (uint8_t)(t == 3 ? rng() : 100 + rng() % 30);
        for (auto &v : cr) v = (uint8_t)(t == 3 ? rng() : 140 + rng() % 20);
        wo::Tuning tu;
        wo::Encoder enc;
        enc.k.init(qp, tu);
        enc.max_depth = 3;
        wo::Picture pic;
        pic.init(W, H, y.data(), cb.data(), cr.data());
        enc.search_picture(pic);
        std::vector<uint16_t> bins;
        const std::vector<uint8_t> want = wo::code_slice_data_traced(enc.k, pic, bins);
        const std::vector<uint8_t> got = code_fast(qp, bins);
        g_checked++;
        g_bins += (long)bins.size();
        long nbyp = 0;
        for (uint16_t e : bins) nbyp += (e >> 10) & 1;
        printf("picture qp %d: %zu bins (%ld bypass), %zu bytes\n", qp, bins.size(), nbyp, want.size());
        if (want != got || want != code_tokens(qp, bins) || want != code_records(qp, bins)) { fprintf(stderr, "MISMATCH real picture qp %d\n", qp); g_bad++; }
    }
    // optional: raw I420 clips (the reference's test clips decoded by tools/decode_assets.py) after the seed count: file W H qp [...],
    // first frame cropped to multiples of 32: the real bin string of a natural picture through all three forms of the coder
    for (int a = 2; a + 3 < argc; a += 4) {
        const int W0 = atoi(argv[a + 1]), H0 = atoi(argv[a + 2]), qp = atoi(argv[a + 3]);
        const int Wc = W0 / 32 * 32, Hc = H0 / 32 * 32;
        FILE *f = fopen(argv[a], "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
        std::vector<uint8_t> raw((size_t)W0 * H0 * 3 / 2);
        if (fread(raw.data(), 1, raw.size(), f) != raw.size()) { fprintf(stderr, "short read %s\n", argv[a]); return 2; }
        fclose(f);
        std::vector<uint8_t> y((size_t)Wc * Hc), cb((size_t)Wc * Hc / 4), cr((size_t)Wc * Hc / 4);
        for (int j = 0; j < Hc; j++) memcpy(&y[(size_t)j * Wc], &raw[(size_t)j * W0], Wc);
        for (int j = 0; j < Hc / 2; j++) {
            memcpy(&cb[(size_t)j * (Wc / 2)], &raw[(size_t)W0 * H0 + (size_t)j * (W0 / 2)], Wc / 2);
            memcpy(&cr[(size_t)j * (Wc / 2)], &raw[(size_t)W0 * H0 * 5 / 4 + (size_t)j * (W0 / 2)], Wc / 2);
        }
        wo::Tuning tu;
        wo::Encoder enc;
        enc.k.init(qp, tu);
        enc.max_depth = 3;
        wo::Picture pic;
        pic.init(Wc, Hc, y.data(), cb.data(), cr.data());
        enc.search_picture(pic);
        std::vector<uint16_t> bins;
        const std::vector<uint8_t> want = wo::code_slice_data_traced(enc.k, pic, bins);
        g_checked++;
        g_bins += (long)bins.size();
        printf("%s qp %d: %zu bins, %zu bytes\n", argv[a], qp, bins.size(), want.size());
        if (want != code_fast(qp, bins) || want != code_tokens(qp, bins) || want != code_records(qp, bins)) { fprintf(stderr, "MISMATCH clip %s qp %d\n", argv[a], qp); g_bad++; }
    }
    printf("%ld strings, %ld bins, %ld mismatches\n", g_checked, g_bins, g_bad);
    return g_bad ? 1 : 0;
}
