// Host check of wrenc_b200/csrc/syntax_walk.cuh (the slice coder's syntax walk: coding_tree / coding_unit / transform_unit /
// residual_coding of one CTU as a bin string, with coded-block / coded-sub-block decisions taken from the per-CTU map of non-zero
// 4x4 level blocks) against the CPU oracle's syntax writer (oracle/wrenc_oracle_cabac.cpp, ctu_encoder.rs:227-2269): pictures are
// searched by the oracle, walked CTU by CTU with the product's code, and every bin-string entry (context index, bin, bypass flag)
// must equal what the oracle hands to its arithmetic coder.  Test infrastructure: built and run by
// tests/test_syntax_walk_host.py (g++, no GPU); the same header is compiled into wrenc_b200_syntax_kernel.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../oracle/wrenc_oracle.hpp"
#include "../../wrenc_b200/csrc/syntax_walk.cuh"

// the definition of the map wrenc_b200_nzmap_kernel computes (its warp-parallel form is checked on the GPU by output equality)
static wb::NzMap nzmap_of_ctu(const wo::Picture &p, int cx, int cy) {
    wb::NzMap m{};
    const int cw = p.W / 2;
    for (int by = 0; by < 8; by++)
        for (int bx = 0; bx < 8; bx++) {
            bool nz = false;
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) nz |= p.coef[0][(size_t)(cy + by * 4 + y) * p.W + cx + bx * 4 + x] != 0;
            if (nz) m.y |= 1ull << (by * 8 + bx);
        }
    for (int c = 1; c <= 2; c++) {
        unsigned v = 0;
        for (int by = 0; by < 4; by++)
            for (int bx = 0; bx < 4; bx++) {
                bool nz = false;
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++) nz |= p.coef[c][(size_t)(cy / 2 + by * 4 + y) * cw + cx / 2 + bx * 4 + x] != 0;
                if (nz) v |= 1u << (by * 4 + bx);
            }
        (c == 1 ? m.cb : m.cr) = (unsigned short)v;
    }
    return m;
}

static long g_pics = 0, g_ctus = 0, g_bins = 0, g_bad = 0;

static void check_picture(int W, int H, int qp, int depth, const std::vector<uint8_t> &y, const std::vector<uint8_t> &cb, const std::vector<uint8_t> &cr, const char *what) {
    wo::Tuning tu;
    wo::Encoder enc;
    enc.k.init(qp, tu);
    enc.max_depth = depth;
    wo::Picture pic;
    pic.init(W, H, y.data(), cb.data(), cr.data());
    enc.search_picture(pic);
    std::vector<uint16_t> want;
    wo::code_slice_data_traced(enc.k, pic, want);

    static_assert(sizeof(wb::CtuRecord) == sizeof(wo::CtuRecord), "record layouts");
    wb::PicView P;
    P.W = W; P.H = H; P.Wc = W / 32; P.Hc = H / 32;
    for (int c = 0; c < 3; c++) P.lev[c] = pic.coef[c].data();
    P.rec = reinterpret_cast<const wb::CtuRecord *>(pic.records.data());
    P.mode_map = pic.mode_map.data();
    wb::SbOrder SO;
    wb::build_sb_order(SO);
    std::vector<uint16_t> got(want.size() + 4096);
    size_t n = 0;
    for (int ctu = 0; ctu < P.Wc * P.Hc; ctu++) {
        const int cx = (ctu % P.Wc) * 32, cy = (ctu / P.Wc) * 32;
        wb::Sink S;
        S.p = got.data() + n;
        S.n = 0;
        S.cap = (int)(got.size() - n);
        wb::TuState ts;
        ts.qp_delta_coded = false;
        ts.mts_dc_only = true;
        ts.mts_zero_out = true;
        alignas(4) uint8_t pass1[1024], absl[1024];
        const wb::NzMap nz = nzmap_of_ctu(pic, cx, cy);
        wb::code_ctu(S, SO, P, P.rec[ctu], nz, cx, cy, ts, pass1, absl);
        if (S.n > S.cap) { fprintf(stderr, "%s: string longer than the oracle's whole picture\n", what); g_bad++; return; }
        n += (size_t)S.n;
        g_ctus++;
    }
    g_pics++;
    g_bins += (long)want.size();
    if (n != want.size() || memcmp(got.data(), want.data(), n * sizeof(uint16_t)) != 0) {
        size_t k = 0;
        while (k < n && k < want.size() && got[k] == want[k]) k++;
        fprintf(stderr, "MISMATCH %s (%dx%d qp %d depth %d): %zu entries vs %zu, first difference at %zu\n", what, W, H, qp, depth, n, want.size(), k);
        g_bad++;
    }
}

int main(int argc, char **argv) {
    struct Cfg { int W, H, qp, depth, kind; };
    const Cfg cfgs[] = {{96, 64, 32, 3, 0}, {96, 64, 22, 3, 0}, {96, 64, 37, 3, 0}, {96, 64, 27, 3, 1}, {64, 64, 32, 0, 0}, {64, 64, 32, 1, 0}, {64, 96, 32, 2, 0},
                        {128, 96, 17, 3, 2}, {96, 64, 12, 3, 3}, {64, 32, 45, 3, 1}, {64, 64, 63, 3, 0}, {160, 96, 30, 3, 1}, {32, 32, 26, 3, 3}, {32, 96, 32, 3, 1}};
    int t = 0;
    for (const Cfg &c : cfgs) {
        std::mt19937 rng(4242u + (unsigned)t++);
        std::vector<uint8_t> y((size_t)c.W * c.H), cb((size_t)c.W * c.H / 4), cr((size_t)c.W * c.H / 4);
        for (int j = 0; j < c.H; j++)
            for (int i = 0; i < c.W; i++) {
                int v;
                if (c.kind == 0) v = 128 + 60 * ((i / 7 + j / 5) & 1) + (int)(rng() % 9) + ((i * j) >> 6);       // edges + light noise
                else if (c.kind == 1) v = 90 + (i * 3 + j * 2) % 120 + (int)(rng() % 5) * ((i / 16 + j / 16) & 1);  // gradients, noisy tiles
                else if (c.kind == 2) v = 128 + (int)(rng() % 64) - 32 + 40 * ((i / 3) & 1);                         // busy
                else v = (int)(rng() & 255);                                                                        // noise
                y[(size_t)j * c.W + i] = (uint8_t)std::min(255, std::max(0, v));
            }
        for (auto &v : cb) v = (uint8_t)(c.kind == 3 ? rng() : 100 + rng() % (c.kind == 2 ? 60 : 12));
        for (auto &v : cr) v = (uint8_t)(c.kind == 3 ? rng() : 140 + rng() % (c.kind == 2 ? 50 : 8));
        check_picture(c.W, c.H, c.qp, c.depth, y, cb, cr, "synthetic");
    }
    // optional: raw I420 clips (the reference's test clips decoded by tools/decode_assets.py): file W H qp [file W H qp ...], first frame, cropped to multiples of 32
    for (int a = 1; a + 3 < argc; a += 4) {
        const int W0 = atoi(argv[a + 1]), H0 = atoi(argv[a + 2]), qp = atoi(argv[a + 3]);
        const int W = W0 / 32 * 32, H = H0 / 32 * 32;
        FILE *f = fopen(argv[a], "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
        std::vector<uint8_t> raw((size_t)W0 * H0 * 3 / 2);
        if (fread(raw.data(), 1, raw.size(), f) != raw.size()) { fprintf(stderr, "short read %s\n", argv[a]); return 2; }
        fclose(f);
        std::vector<uint8_t> y((size_t)W * H), cb((size_t)W * H / 4), cr((size_t)W * H / 4);
        for (int j = 0; j < H; j++) memcpy(&y[(size_t)j * W], &raw[(size_t)j * W0], W);
        for (int j = 0; j < H / 2; j++) {
            memcpy(&cb[(size_t)j * (W / 2)], &raw[(size_t)W0 * H0 + (size_t)j * (W0 / 2)], W / 2);
            memcpy(&cr[(size_t)j * (W / 2)], &raw[(size_t)W0 * H0 * 5 / 4 + (size_t)j * (W0 / 2)], W / 2);
        }
        check_picture(W, H, qp, 3, y, cb, cr, argv[a]);
    }
    printf("%ld pictures, %ld CTUs, %ld bin-string entries, %ld mismatches\n", g_pics, g_ctus, g_bins, g_bad);
    return g_bad ? 1 : 0;
}
