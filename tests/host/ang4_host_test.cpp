// Host check of wrenc_b200/csrc/ang4.cuh (the four-samples-per-lane angular predictor of the search kernel) against the CPU
// oracle's prediction, for every angular mode x block size x component x neighbour-availability pattern.  Test infrastructure:
// built and run by tests/test_ang4_host.py (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../oracle/wrenc_oracle.hpp"
#include "../../wrenc_b200/csrc/ang4.cuh"

static const int8_t kAngle[67] = {0,   0,   32,  29,  26,  23,  20,  18,  16,  14,  12,  10,  8,   6,   4,   3,   2,
                                  1,   0,   -1,  -2,  -3,  -4,  -6,  -8,  -10, -12, -14, -16, -18, -20, -23, -26, -29,
                                  -32, -29, -26, -23, -20, -18, -16, -14, -12, -10, -8,  -6,  -4,  -3,  -2,  -1,  0,
                                  1,   2,   3,   4,   6,   8,   10,  12,  14,  16,  18,  20,  23,  26,  29,  32};
static const int8_t kFC[32][4] = {
    {0, 64, 0, 0},    {-1, 63, 2, 0},   {-2, 62, 4, 0},   {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
    {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
    {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
    {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
    {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
    {0, 4, 62, -2},   {0, 2, 63, -1}};

static int ilog2i(int v) { int l = 0; while ((1 << (l + 1)) <= v) l++; return l; }

int main(int argc, char **argv) {
    const int seeds = argc > 1 ? atoi(argv[1]) : 3;
    a4::TapTables T;
    a4::fill_tap_tables(T, kFC);
    long checked = 0, bad = 0;
    const int W = 128, H = 96;
    for (int seed = 0; seed < seeds; seed++) {
        srand(1234 + seed);
        std::vector<uint8_t> y(W * H), cb(W * H / 4), cr(W * H / 4);
        for (auto &v : y) v = (uint8_t)(seed == 0 ? rand() : (rand() % 3 == 0 ? rand() : 100 + rand() % 40));
        for (auto &v : cb) v = (uint8_t)rand();
        for (auto &v : cr) v = (uint8_t)(seed == 1 ? (rand() & 1 ? 255 : 0) : rand());
        wo::Picture pic;
        pic.init(W, H, y.data(), cb.data(), cr.data());
        for (int c = 0; c < 3; c++) pic.rec[c].d = pic.orig[c].d;  // prediction reads the reconstruction planes
        for (int w = 4; w <= 32; w *= 2) {
            // positions: picture corner / edges / interior, CTU-aligned and not
            const int pos[][2] = {{0, 0}, {32, 0}, {0, 32}, {32, 32}, {64, 32}, {96, 64}, {32 + w % 32, 32}, {96, 0}, {0, 64}, {64 + (32 - w), 64 - w + (w == 32 ? 0 : 0)}};
            for (auto &ps : pos) {
                int x0 = ps[0] / w * w, y0 = ps[1] / w * w;
                if (x0 + w > W || y0 + w > H) continue;
                for (int flags = 0; flags < 4; flags++) {
                    for (int c = 0; c < 3; c++) {
                        if (c > 0 && w < 8) continue;
                        const int n = c ? w / 2 : w, l2 = ilog2i(n);
                        for (int mode = 2; mode <= 66; mode++) {
                            wo::TU tu{x0, y0, w, wo::SINGLE_TREE, (flags & 1) != 0, (flags & 2) != 0, {mode, mode, mode}};
                            std::vector<uint8_t> want(n * n), got(n * n);
                            wo::predict(pic, tu, c, want.data());
                            int16_t left[65], above[64], leftF[65], aboveF[64];
                            wo::build_refs(pic, tu, c, mode, left, above, leftF, aboveF);
                            const bool filt = c == 0 && n >= 8 && (mode == 2 || mode == 34 || mode == 66);
                            const int16_t *L = filt ? leftF : left, *A = filt ? aboveF : above;
                            // the two mirrored byte lines, index range [-n, 2n+3], 8 bytes of slack on either side
                            std::vector<uint8_t> upb(3 * n + 4 + 16, 0xAA), dnb(3 * n + 4 + 16, 0x55);
                            uint8_t *up = upb.data() + 8 + n, *dn = dnb.data() + 8 + n;
                            up[0] = dn[0] = (uint8_t)L[0];
                            for (int k = 1; k <= 2 * n; k++) { up[k] = (uint8_t)A[k - 1]; dn[k] = (uint8_t)L[k]; }
                            for (int k = 2 * n + 1; k <= 2 * n + 3; k++) { up[k] = up[2 * n]; dn[k] = dn[2 * n]; }
                            for (int k = 1; k <= n; k++) { up[-k] = dn[k]; dn[-k] = up[k]; }
                            const int ang = kAngle[mode];
                            const int inv = ang > 0 ? (512 * 32 + ang / 2) / ang : (ang < 0 ? -((512 * 32 + (-ang) / 2) / -ang) : 0);
                            const bool vertical = mode >= 34;
                            a4::Mode m;
                            m.ang = ang; m.inv = inv;
                            m.main = vertical ? up : dn;
                            m.side = vertical ? dn : up;
                            m.taps = c ? 2 : 0;
                            if (c == 0 && !(mode == 2 || mode == 34 || mode == 66)) {
                                const int md = std::min(abs(mode - 50), abs(mode - 18));
                                const int thr = l2 == 2 ? 24 : (l2 == 3 ? 14 : (l2 == 4 ? 2 : 0));
                                if (md > thr) m.taps = 1;
                            }
                            m.pdpc = 0; m.ns = 0;
                            if (mode <= 18 || mode >= 50) {
                                const int ns = (mode == 18 || mode == 50) ? (2 * l2 - 2) >> 2 : std::min(l2 - ilog2i(3 * inv - 2) + 8, 2);
                                if (ns >= 0) { m.pdpc = (mode == 18 || mode == 50) ? 1 : 2; m.ns = ns; }
                            }
                            std::vector<uint8_t> projb(2 * n + 4 + 16, 0x77);
                            if (ang < 0) {
                                uint8_t *pr = projb.data() + 8 + n;
                                for (int i = 0; i < 2 * n + 4; i++) a4::project_elem(pr, m.main, m.side, n, inv, i);
                                m.main = pr;
                            }
                            for (int t = 0; t < n; t++)
                                for (int u0 = 0; u0 < n; u0 += 4) {
                                    const unsigned q = a4::quad(m, T, t, u0);
                                    for (int i = 0; i < 4; i++) {
                                        const int px = vertical ? u0 + i : t, py = vertical ? t : u0 + i;
                                        got[py * n + px] = (uint8_t)(q >> (8 * i));
                                    }
                                }
                            checked++;
                            if (memcmp(want.data(), got.data(), n * n)) {
                                if (bad++ < 10) {
                                    int k = 0;
                                    while (want[k] == got[k]) k++;
                                    fprintf(stderr, "MISMATCH seed %d w %d pos %d,%d flags %d c %d mode %d: sample %d,%d want %d got %d\n", seed, w, x0, y0, flags, c, mode, k % n, k / n,
                                            want[k], got[k]);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    printf("ang4: %ld blocks checked, %ld mismatches\n", checked, bad);
    return bad ? 1 : 0;
}
