// search_kernel_api.h — what the host side (api.cu) needs to know about the search kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tuning.h"

namespace wb {

struct CtuRecord {
    uint32_t split_mask;
    uint8_t luma_mode[64];
    uint8_t chroma_mode[16];
    float cost;
};
static_assert(sizeof(CtuRecord) == 88, "record layout is part of the C ABI");

struct SearchParams {
    int W, H, Wc, Hc;      // luma size, CTUs per row / column
    int max_depth;
    int n_items;           // number of slots (batches x CTUs per CTA)
    int epoch;             // value a done flag takes when its CTU is final in THIS launch
    const uint8_t *orig;   // [pic][W*H*3/2] I420
    uint8_t *rec;          // same geometry
    int16_t *lev;          // same geometry, int16 per sample
    uint8_t *mode_map;     // [pic][(W/4)*(H/4)] final luma mode per 4x4 (left-CTU MPM lookups, H1)
    CtuRecord *records;    // [pic][Wc*Hc]
    int *done;             // [pic][Wc*Hc]
    const uint32_t *items; // work list: batches of search_ctus_per_cta() slots, pic<<16 | cy<<8 | cx or 0xffffffff (empty slot)
    unsigned int *counter; // work-list cursor
    uint8_t *root_slots;   // [CTA][CTU of the batch][CTU_SCRATCH_BYTES]: candidate slots of the CU being evaluated + saved no-split states
                           // (written and read back by the same CTA, L2 resident; keeps the per-CTU shared-memory context small)
    const DevTables *tab;
};

// per CTU: 6 slots (planar, DC, dir, dir-1, dir+1, CCLM) x 1536 samples (Y 1024, Cb 256, Cr 256): reconstruction u8, then levels i16
constexpr int ROOT_SLOT_SAMPLES = 1536, ROOT_SLOTS = 6, ROOT_SLOT_BYTES = ROOT_SLOTS * ROOT_SLOT_SAMPLES * 3;
// no-split state saved per depth (32x32, 16x16, 8x8): Y 1024 + 256 + 64 samples, Cb / Cr 256 + 64 + 16 each: reconstruction u8, then levels i16
constexpr int SAVE_Y = 1024 + 256 + 64, SAVE_C = 256 + 64 + 16, SAVE_SAMPLES = SAVE_Y + 2 * SAVE_C;
constexpr int CTU_SCRATCH_BYTES = ROOT_SLOT_BYTES + SAVE_SAMPLES * 3;
static_assert(CTU_SCRATCH_BYTES % 16 == 0 && ROOT_SLOT_BYTES % 16 == 0 && SAVE_SAMPLES % 8 == 0, "alignment of the per-CTU global scratch");
enum { SINGLE_TREE = 0, DUAL_TREE_LUMA = 1, DUAL_TREE_CHROMA = 2 };
enum { MODE_PLANAR = 0, MODE_DC = 1, MODE_LT_CCLM = 81, MODE_L_CCLM = 82, MODE_T_CCLM = 83 };

struct NzMap {  // per CTU: which 4x4 blocks of the level planes hold a non-zero level (luma bit by * 8 + bx, chroma bit by * 4 + bx)
    unsigned long long y;
    unsigned short cb, cr;
    unsigned int pad;
};
struct SyntaxParams {  // slice_coder.cu: decided trees -> CABAC-coded slice_data() per picture
    int W, H, Wc, Hc, n_pics, qp;
    const int16_t *lev;        // [pic][W*H*3/2]
    const CtuRecord *records;  // [pic][Wc*Hc]
    const uint8_t *mode_map;   // [pic][(W/4)*(H/4)]
    NzMap *nzmap;              // [pic][ctu] written by the counting pass's first kernel, read by both syntax passes
    uint16_t *bins;            // bin arena (entries: ctx index | bin << 9 | bypass << 10); nullptr = counting + staging pass
    unsigned long long bins_cap;  // entries the arena holds: strings that would end beyond it are not written and their picture reports
                               //   out_len = -2, so that the host can grow the arena and run the passes again WITHOUT a mid-path sync
    uint16_t *stage;           // [pic][ctu][stage_cap]: the counting pass keeps the first stage_cap entries of every CTU's bin string here, so that
    int stage_cap;             //   only CTUs with longer strings are walked a second time; the others are copied to their arena offset
    int *bin_count;            // [pic][ctu] entries of the CTU's bin string
    unsigned long long *bin_offset;  // [pic][ctu] start of the CTU's bin string in the arena (exclusive scan of bin_count)
    uint8_t *out;              // [pic][out_cap] bytes
    size_t out_cap;
    int *out_len;              // [pic] bytes written, -1 on overflow
};
cudaError_t launch_syntax(const SyntaxParams &Q, cudaStream_t stream);  // Q.bins == nullptr: non-zero map + counting pass
int syntax_first_pass_kernels();  // kernels launch_syntax enqueues for the counting pass
cudaError_t launch_bin_scan(const SyntaxParams &Q, unsigned long long *d_total, cudaStream_t stream);
cudaError_t launch_bin_compact(const SyntaxParams &Q, cudaStream_t stream);
cudaError_t launch_cabac(const SyntaxParams &Q, cudaStream_t stream);

struct BlockParams {
    int op;  // 0 predict, 1 forward DCT, 2 inverse DCT, 3 dep-quant (+rate) by the routine the search uses for that size, 4 dequantise,
             // 5 dep-quant of every size by trellis()
    int W, H, x, y, w, tree, ar, bl, c, mode;
    const uint8_t *rec;   // I420 reconstruction picture (op 0)
    uint8_t *out8;
    int l2, count;
    const int16_t *in;
    int16_t *out16;
    int *outi;
    const DevTables *tab;
};
cudaError_t launch_block(const BlockParams &P, int grid, cudaStream_t stream);
size_t search_smem_bytes();
int search_ctas_per_sm();
int search_ctus_per_cta();  // CTUs one CTA searches in lock step; the work list is made of batches of this many slots
// persist_base/persist_bytes: L2 persisting access window for this launch only (the per-CTA scratch), or nullptr/0
cudaError_t launch_search(const SearchParams &P, int grid, cudaStream_t stream, const void *persist_base = nullptr, size_t persist_bytes = 0);

}  // namespace wb
