// ang4.cuh — angular intra prediction, FOUR samples per lane (intra_predictor.rs:1287-1602 + PDPC 355-757), written so that the
// per-lane arithmetic compiles for the device (search kernel) and for the host (tests/host/ang4_host_test.cpp checks it against
// the oracle for every mode, size, component and reference-sample pattern without a GPU).
//
// Frame of reference: every angular mode predicts along "lines": for the vertical modes (>= 34) a line is a row (t = y, u = x),
// for the horizontal modes (< 34) a column (t = x, u = y).  Sample (t, u) is a 4-tap (luma) / 2-tap (chroma) filter over the
// projected reference  ref[u + iIdx(t) + 0..3],  iIdx = ((t+1)*angle) >> 5,  iFact = ((t+1)*angle) & 31.  With the reference
// samples held as BYTES in two mirrored lines per component
//     up[k]: up[0] = corner, up[k] = above[k-1] (k > 0),   dn[k]: dn[0] = corner, dn[k] = left[k] (k > 0)
// (each padded with three copies of its last sample: ref[min(idx, 2n)]), the projected reference of a non-negative angle is
// the MAIN line itself (up for vertical, dn for horizontal modes) and four consecutive samples of a line read seven
// consecutive bytes: three aligned 32-bit loads, funnel shifts, and one packed dot product (dp4a, unsigned samples x signed
// taps) per sample.  A negative angle needs the other (SIDE) line projected through the inverse angle below index 0: the
// caller builds main[-n .. -1] once per (mode, block) in a scratch line (ang4_project).
// The chroma 2-tap filter ((32-f)*r1 + f*r2 + 16) >> 5 is written as the 4-tap {0, 64-2f, 2f, 0} with the luma rounding
// (x + 32) >> 6: floor((X + 16) / 32) == floor((2X + 32) / 64).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define A4_HD __host__ __device__ __forceinline__
#else
#define A4_HD inline
#endif

namespace a4 {

#if defined(__CUDA_ARCH__)
A4_HD unsigned fsr(unsigned lo, unsigned hi, unsigned bits) { return __funnelshift_r(lo, hi, bits); }
A4_HD int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
A4_HD unsigned sad4(unsigned a, unsigned b, unsigned c) { return __vsadu4(a, b) + c; }
A4_HD int imin(int a, int b) { return min(a, b); }
A4_HD int imax(int a, int b) { return max(a, b); }
#else
A4_HD unsigned fsr(unsigned lo, unsigned hi, unsigned bits) {
    bits &= 31u;
    return bits ? (lo >> bits) | (hi << (32u - bits)) : lo;
}
A4_HD int dp4a_us(unsigned a, int b, int c) {
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 255u) * (int)(int8_t)((unsigned)b >> (8 * i));
    return c;
}
A4_HD unsigned sad4(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) {
        int d = (int)((a >> (8 * i)) & 255u) - (int)((b >> (8 * i)) & 255u);
        c += (unsigned)(d < 0 ? -d : d);
    }
    return c;
}
A4_HD int imin(int a, int b) { return a < b ? a : b; }
A4_HD int imax(int a, int b) { return a > b ? a : b; }
#endif

A4_HD int clip255(int v) { return imin(255, imax(0, v)); }

// Tap tables, packed 4 x int8 (tap i in byte i): [0] fC (VVC table 25, common.rs:153-186), [1] fG = {16-h, 32-h, 16+h, h} with
// h = iFact >> 1, [2] the chroma 2-tap filter doubled {0, 64-2f, 2f, 0}.
struct TapTables {
    int t[3][32];
};
A4_HD void fill_tap_tables(TapTables &T, const int8_t (*fC)[4]) {
    for (int f = 0; f < 32; f++) {
        unsigned a = 0;
        for (int i = 0; i < 4; i++) a |= (unsigned)(uint8_t)fC[f][i] << (8 * i);
        const int h = f >> 1;
        T.t[0][f] = (int)a;
        T.t[1][f] = (int)((unsigned)(16 - h) | ((unsigned)(32 - h) << 8) | ((unsigned)(16 + h) << 16) | ((unsigned)h << 24));
        T.t[2][f] = (int)(((unsigned)(64 - 2 * f) << 8) | ((unsigned)(2 * f) << 16));
    }
}

// Per (mode, component, block size) constants of one prediction.
struct Mode {
    const uint8_t *main;  // main[idx], idx in [-n, 2n+3]; for a negative angle a projected scratch copy (ang4_project)
    const uint8_t *side;  // the other line (PDPC): side[0] = corner, side[1 + i] = i-th sample
    int ang;              // intraPredAngle
    int inv;              // inverse angle with the angle's sign (intra_predictor.rs:1390-1396)
    int taps;             // 0 fC, 1 fG, 2 chroma
    int pdpc;             // 0 none, 1 modes 18 / 50, 2 other angular modes with PDPC
    int ns;               // PDPC nScale
};

// Projection of the side line below index 0 for a negative angle (intra_predictor.rs:1398-1414 / 1480-1495):
// dst[idx] = side[min((idx * inv + 256) >> 9, n)] for idx in [-n, -1] (inv < 0), dst[idx] = main_line[idx] for idx in [0, n+3].
// `i` runs over [0, 2n+4): one call per element (the caller spreads the elements over lanes).
A4_HD void project_elem(uint8_t *dst, const uint8_t *main_line, const uint8_t *side, int n, int inv, int i) {
    const int idx = i - n;
    dst[idx] = idx < 0 ? side[imin((idx * inv + 256) >> 9, n)] : main_line[idx];
}

// Four samples (t, u0 .. u0+3) of one line, packed little endian (sample u0 in byte 0).
A4_HD unsigned quad(const Mode &m, const TapTables &T, int t, int u0) {
    const int prod = (t + 1) * m.ang;
    const int iidx = prod >> 5, ifact = prod & 31;
    const uint8_t *p = m.main + iidx + u0;
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
#if defined(__CUDA_ARCH__)
    __builtin_assume(__isShared(w));  // the reference lines live in shared memory: LDS, not generic loads
#endif
    const unsigned sh = (unsigned)(a & 3) * 8u;
    const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
    const unsigned lo = fsr(w0, w1, sh), hi = fsr(w1, w2, sh);  // bytes 0-3 and 4-7 of the window
    const int taps = T.t[m.taps][ifact];
    int s0 = dp4a_us(lo, taps, 32), s1 = dp4a_us(fsr(lo, hi, 8), taps, 32), s2 = dp4a_us(fsr(lo, hi, 16), taps, 32), s3 = dp4a_us(fsr(lo, hi, 24), taps, 32);
    int p0 = clip255(s0 >> 6), p1 = clip255(s1 >> 6), p2 = clip255(s2 >> 6), p3 = clip255(s3 >> 6);
    if (m.pdpc == 1) {  // modes 18 / 50: ref = side[1 + t] - corner + pred, weight by u (intra_predictor.rs:355-757)
        const int d = (int)m.side[1 + t] - (int)m.side[0];
        int wgt;
        wgt = (2 * u0) >> m.ns;       wgt = wgt > 5 ? 0 : 32 >> wgt; p0 = clip255(((d + p0) * wgt + (64 - wgt) * p0 + 32) >> 6);
        wgt = (2 * u0 + 2) >> m.ns;   wgt = wgt > 5 ? 0 : 32 >> wgt; p1 = clip255(((d + p1) * wgt + (64 - wgt) * p1 + 32) >> 6);
        wgt = (2 * u0 + 4) >> m.ns;   wgt = wgt > 5 ? 0 : 32 >> wgt; p2 = clip255(((d + p2) * wgt + (64 - wgt) * p2 + 32) >> 6);
        wgt = (2 * u0 + 6) >> m.ns;   wgt = wgt > 5 ? 0 : 32 >> wgt; p3 = clip255(((d + p3) * wgt + (64 - wgt) * p3 + 32) >> 6);
    } else if (m.pdpc == 2 && u0 < (3 << m.ns)) {  // side sample at t + ((u+1)*invAngle + 256 >> 9), weight 32 >> (2u >> nScale), u < 3 << nScale
        const uint8_t *sp = m.side + 1 + t;
        const int lim = 3 << m.ns;
        if (u0 < lim)     { const int wgt = 32 >> ((2 * u0) >> m.ns);     p0 = clip255(((int)sp[((u0 + 1) * m.inv + 256) >> 9] * wgt + (64 - wgt) * p0 + 32) >> 6); }
        if (u0 + 1 < lim) { const int wgt = 32 >> ((2 * u0 + 2) >> m.ns); p1 = clip255(((int)sp[((u0 + 2) * m.inv + 256) >> 9] * wgt + (64 - wgt) * p1 + 32) >> 6); }
        if (u0 + 2 < lim) { const int wgt = 32 >> ((2 * u0 + 4) >> m.ns); p2 = clip255(((int)sp[((u0 + 3) * m.inv + 256) >> 9] * wgt + (64 - wgt) * p2 + 32) >> 6); }
        if (u0 + 3 < lim) { const int wgt = 32 >> ((2 * u0 + 6) >> m.ns); p3 = clip255(((int)sp[((u0 + 4) * m.inv + 256) >> 9] * wgt + (64 - wgt) * p3 + 32) >> 6); }
    }
    return (unsigned)p0 | ((unsigned)p1 << 8) | ((unsigned)p2 << 16) | ((unsigned)p3 << 24);
}

}  // namespace a4
