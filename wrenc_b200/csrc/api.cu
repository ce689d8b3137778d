// api.cu — the C ABI of include/wrenc_b200.h: handle, picture batching, work-list construction, launches, copies.
// Host-side counterpart of the reference's per-picture driver (src/main.rs:294-402) for the part that moves behind the FFI.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wrenc_b200.h"
#include "search_kernel_api.h"

using namespace wb;

static thread_local std::string g_create_err;

struct wrenc_b200 {
    wrenc_b200_config cfg{};
    std::string extra;
    std::string err;
    HostConsts hc;
    int W = 0, H = 0, Wc = 0, Hc = 0, B = 1;
    size_t pic_samples = 0;
    int sm_count = 0, ctas_per_sm = 0, grid = 0;
    cudaStream_t stream = nullptr;
    DevTables *d_tab = nullptr;
    // workspace shared by both entry points (sized for ws_pics pictures)
    int ws_pics = 0;
    uint8_t *d_mode_map = nullptr;
    uint8_t *d_root_slots = nullptr;  // per CTA x lock-step CTU: candidate slots of the 32x32 root CU
    int *d_done = nullptr;
    uint32_t *d_items = nullptr;
    size_t items_cap = 0;
    unsigned int *d_counter = nullptr;
    int items_for = -1;  // n_pictures the uploaded work list was built for
    int n_items = 0;
    int epoch = 0;
    // batch buffers of the host-plane path
    uint8_t *d_orig = nullptr, *d_rec = nullptr;
    int16_t *d_lev = nullptr;
    CtuRecord *d_rec_ctu = nullptr;
    uint8_t *h_orig = nullptr, *h_rec = nullptr;
    int16_t *h_lev = nullptr;
    CtuRecord *h_records = nullptr;
    std::vector<uint64_t> pic_ids;
    int n_filled = 0, n_returned = 0, last_returned = -1;
    bool launched = false;
    // phase 2: slice_data coder buffers (sized for coder_pics pictures)
    int coder_pics = 0;
    uint16_t *d_bins = nullptr;
    size_t bins_cap = 0;
    uint16_t *d_stage = nullptr;  // per-CTU staging slots of the bin strings (stage_cap() entries each)
    int *d_bin_count = nullptr;
    unsigned long long *d_bin_offset = nullptr, *d_bin_total = nullptr;
    uint8_t *d_out = nullptr;
    int *d_out_len = nullptr;
    size_t out_cap = 0;
    uint8_t *h_out = nullptr;
    int *h_out_len = nullptr;
    cudaEvent_t ev_done = nullptr;
    unsigned long long launches = 0;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return WRENC_B200_ECUDA;                                                               \
        }                                                                                          \
    } while (0)

static int ensure_workspace(wrenc_b200 *h, int n_pics) {
    if (n_pics <= h->ws_pics) return 0;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_mode_map); cudaFree(h->d_done);
    h->d_mode_map = nullptr; h->d_done = nullptr;
    size_t nctu = (size_t)h->Wc * h->Hc * n_pics;
    CK(cudaMalloc(&h->d_mode_map, (size_t)(h->W / 4) * (h->H / 4) * n_pics));
    CK(cudaMalloc(&h->d_done, nctu * sizeof(int)));
    CK(cudaMemsetAsync(h->d_done, 0, nctu * sizeof(int), h->stream));
    CK(cudaMemsetAsync(h->d_mode_map, 0, (size_t)(h->W / 4) * (h->H / 4) * n_pics, h->stream));
    if (!h->d_root_slots) CK(cudaMalloc(&h->d_root_slots, (size_t)h->grid * search_ctus_per_cta() * CTU_SCRATCH_BYTES));
    h->ws_pics = n_pics;
    h->items_for = -1;
    h->epoch = 0;
    return 0;
}

// Work list in wavefront order.  A CTU (x,y) depends on (x-1,y) and (x+1,y-1); both have a smaller key x+2y, so every
// dependency of an item precedes it in the list and a persistent grid that hands items out in list order cannot
// deadlock.  Pictures are staggered so that the ramp-up of one overlaps the ramp-down of another.
static int ensure_items(wrenc_b200 *h, int n_pics) {
    if (h->items_for == n_pics) return 0;
    const int Wc = h->Wc, Hc = h->Hc;
    // Items of one key level never depend on each other; an item of level K+1 depends on two level-K items of its own
    // picture.  All pictures advance in lock step (stagger 0): a level then holds n_pics x (diagonal length) items, so once
    // it is several grids wide every dependency was handed out whole CTU latencies earlier and nobody waits.  Measured on
    // B200 (profiles/r1_variants.txt): staggering the pictures only lengthens the ramp; WRENC_B200_LEVEL_FACTOR=f re-enables
    // it (pictures are staggered so that a level holds about f x grid items) for experiments with a bounded working set.
    double stagger = 0.0;
    if (const char *e = getenv("WRENC_B200_LEVEL_FACTOR")) {
        const double factor = atof(e);
        const double avg_diag = (double)Wc * Hc / (Wc + 2 * Hc);
        if (factor > 0 && n_pics * avg_diag > factor * h->grid) stagger = (double)Wc * Hc / (factor * h->grid);
    }
    struct It { int key, pic, cy, cx; };
    std::vector<It> v;
    v.reserve((size_t)n_pics * Wc * Hc);
    for (int p = 0; p < n_pics; p++)
        for (int cy = 0; cy < Hc; cy++)
            for (int cx = 0; cx < Wc; cx++) v.push_back({cx + 2 * cy + (int)(stagger * p), p, cy, cx});
    std::stable_sort(v.begin(), v.end(), [](const It &a, const It &b) { return a.key < b.key; });
    // Batches of KC mutually independent CTUs (one CTA searches a batch in lock step).  Items of one key level never
    // depend on each other, so a batch never crosses a level boundary; short levels leave empty slots.
    const int KC = search_ctus_per_cta();
    std::vector<uint32_t> items;
    items.reserve(v.size() + (size_t)KC * ((size_t)(Wc + 2 * Hc + n_pics) + (size_t)h->grid * 64));
    // A level too short to give every CTA a full batch (the ramps of the wavefront) is spread over as many batches as there
    // are CTAs: a batch with fewer active CTUs has fewer tasks per phase and finishes sooner.
    for (size_t i = 0; i < v.size();) {
        size_t e = i;
        while (e < v.size() && v[e].key == v[i].key) e++;
        const size_t n = e - i;
        size_t nb = (n + KC - 1) / KC;
        if (nb < (size_t)h->grid) nb = std::min((size_t)h->grid, n);
        for (size_t b = 0; b < nb; b++) {
            const size_t lo = i + n * b / nb, hi = i + n * (b + 1) / nb;  // hi - lo <= KC because nb >= n / KC
            for (size_t q = lo; q < hi; q++) items.push_back(((uint32_t)v[q].pic << 16) | ((uint32_t)v[q].cy << 8) | (uint32_t)v[q].cx);
            for (size_t q = hi - lo; q < (size_t)KC; q++) items.push_back(0xffffffffu);
        }
        i = e;
    }
    if (items.size() > h->items_cap) {
        cudaFree(h->d_items);
        h->d_items = nullptr;
        CK(cudaMalloc(&h->d_items, items.size() * sizeof(uint32_t)));
        h->items_cap = items.size();
    }
    CK(cudaMemcpyAsync(h->d_items, items.data(), items.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));  // `items` is pageable and goes out of scope
    h->items_for = n_pics;
    h->n_items = (int)items.size();
    return 0;
}

static int enqueue_search(wrenc_b200 *h, int n_pics, const uint8_t *d_yuv, uint8_t *d_rec, int16_t *d_lev, CtuRecord *d_records, cudaStream_t st) {
    int rc = ensure_workspace(h, n_pics);
    if (rc) return rc;
    rc = ensure_items(h, n_pics);
    if (rc) return rc;
    SearchParams P;
    P.W = h->W; P.H = h->H; P.Wc = h->Wc; P.Hc = h->Hc;
    P.max_depth = h->cfg.max_split_depth;
    P.n_items = h->n_items;
    P.epoch = ++h->epoch;
    P.orig = d_yuv; P.rec = d_rec; P.lev = d_lev; P.mode_map = h->d_mode_map; P.records = d_records;
    P.done = h->d_done; P.items = h->d_items; P.counter = h->d_counter; P.tab = h->d_tab;
    P.root_slots = h->d_root_slots;
    CK(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), st));
    int grid = std::min(h->grid, h->n_items / search_ctus_per_cta());
    // The CTU scratch (40 MB for 148 x 8 CTUs) is rewritten for every CU: keep it resident in L2 (persisting access window)
    // so that it is not written back to HBM over and over.  Best effort: a failure only costs DRAM traffic.
    {
        const size_t bytes = (size_t)h->grid * search_ctus_per_cta() * CTU_SCRATCH_BYTES;
        static bool limit_set = false;
        if (!limit_set) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes + (4u << 20));
            limit_set = true;
        }
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = h->d_root_slots;
        av.accessPolicyWindow.num_bytes = bytes;
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    CK(launch_search(P, grid, st));
    {   // the window applies to the search kernel only
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    h->launches++;
    return 0;
}

// Entries per CTU staging slot: CTUs whose bin string is longer (busy content at low QP) are walked a second time.
// (WRENC_B200_STAGE_CAP: dev-time override, lets the tests force the second walk.)
static int stage_cap() {
    static int cap = 0;
    if (!cap) {
        const char *e = getenv("WRENC_B200_STAGE_CAP");
        cap = e && atoi(e) > 0 ? atoi(e) : 1024;
    }
    return cap;
}

static int ensure_coder(wrenc_b200 *h, int n_pics) {
    if (n_pics <= h->coder_pics) return 0;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_bin_count); cudaFree(h->d_bin_offset); cudaFree(h->d_out); cudaFree(h->d_out_len); cudaFree(h->d_stage);
    h->d_bin_count = nullptr; h->d_bin_offset = nullptr; h->d_out = nullptr; h->d_out_len = nullptr; h->d_stage = nullptr;
    const size_t nctu = (size_t)h->Wc * h->Hc * n_pics;
    h->out_cap = (size_t)h->W * h->H * 3 / 2;
    CK(cudaMalloc(&h->d_bin_count, nctu * sizeof(int)));
    CK(cudaMalloc(&h->d_stage, nctu * stage_cap() * sizeof(uint16_t)));
    CK(cudaMalloc(&h->d_bin_offset, nctu * sizeof(unsigned long long)));
    CK(cudaMalloc(&h->d_out, (size_t)n_pics * h->out_cap));
    CK(cudaMalloc(&h->d_out_len, (size_t)n_pics * sizeof(int)));
    if (!h->d_bin_total) CK(cudaMalloc(&h->d_bin_total, sizeof(unsigned long long)));
    h->coder_pics = n_pics;
    return 0;
}

// CABAC-code the pictures the last search on this handle decided (the mode map lives in the handle's workspace).
// Pass 1 counts the bins of every CTU, an exclusive scan turns the counts into arena offsets, the host reads the total
// (the one synchronisation of this path) and grows the arena if needed, pass 2 writes the bin strings, then one thread per
// picture runs the arithmetic coder.
static int enqueue_coder(wrenc_b200 *h, int n_pics, const int16_t *d_lev, const CtuRecord *d_records, uint8_t *d_out, size_t out_cap, int *d_out_len,
                         cudaStream_t st) {
    int rc = ensure_coder(h, n_pics);
    if (rc) return rc;
    SyntaxParams Q;
    Q.W = h->W; Q.H = h->H; Q.Wc = h->Wc; Q.Hc = h->Hc; Q.n_pics = n_pics; Q.qp = h->cfg.qp;
    Q.lev = d_lev; Q.records = d_records; Q.mode_map = h->d_mode_map;
    Q.bins = nullptr; Q.bin_count = h->d_bin_count; Q.bin_offset = h->d_bin_offset;
    Q.stage = h->d_stage; Q.stage_cap = stage_cap();
    Q.out = d_out; Q.out_cap = out_cap; Q.out_len = d_out_len;
    CK(launch_syntax(Q, st));
    CK(launch_bin_scan(Q, h->d_bin_total, st));
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, h->d_bin_total, sizeof(total), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (total + 1 > h->bins_cap) {
        cudaFree(h->d_bins);
        h->d_bins = nullptr;
        h->bins_cap = (size_t)(total + total / 4 + 1024);
        CK(cudaMalloc(&h->d_bins, h->bins_cap * sizeof(uint16_t)));
    }
    Q.bins = h->d_bins;
    CK(launch_syntax(Q, st));       // only the CTUs that did not fit their staging slot
    CK(launch_bin_compact(Q, st));  // everybody else: staged string -> arena offset
    CK(launch_cabac(Q, st));
    h->launches += 5;
    return 0;
}

extern "C" {

const char *wrenc_b200_version(void) { return "wrenc_b200 0.1 (sm_100a)"; }

const char *wrenc_b200_last_error(const wrenc_b200 *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int wrenc_b200_create(const wrenc_b200_config *cfg, wrenc_b200 **out) {
    if (!out) return WRENC_B200_EINVAL;
    *out = nullptr;
    if (!cfg) { g_create_err = "null config"; return WRENC_B200_EINVAL; }
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->width % 32 || cfg->height % 32 || cfg->width > 8160 || cfg->height > 8160) {
        g_create_err = "width/height must be positive multiples of 32 (at most 8160)";
        return WRENC_B200_EINVAL;
    }
    if (cfg->qp < 0 || cfg->qp > 63 || cfg->max_split_depth < 0 || cfg->max_split_depth > 3 || cfg->pictures_in_flight < 0) {
        g_create_err = "qp must be 0..63, max_split_depth 0..3";
        return WRENC_B200_EINVAL;
    }
    wrenc_b200 *h = new wrenc_b200();
    h->cfg = *cfg;
    if (cfg->extra_params) h->extra = cfg->extra_params;
    h->cfg.extra_params = nullptr;
    Tuning t;
    if (!t.parse(h->extra.c_str(), g_create_err) || !h->hc.init(cfg->qp, t, g_create_err)) {
        delete h;
        return WRENC_B200_EINVAL;
    }
    h->W = cfg->width; h->H = cfg->height; h->Wc = h->W / 32; h->Hc = h->H / 32;
    h->B = std::max(1, cfg->pictures_in_flight);
    if (h->B > 65535) h->B = 65535;
    h->pic_samples = (size_t)h->W * h->H * 3 / 2;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        g_create_err = "no usable CUDA device (this library has no CPU fallback)";
        cudaGetLastError();
        delete h;
        return WRENC_B200_ENODEV;
    }
    auto fail = [&](const char *what, cudaError_t e, int code) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(e);
        wrenc_b200_destroy(h);
        return code;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail("cudaSetDevice", e, WRENC_B200_ENODEV);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e, WRENC_B200_ENODEV);
    if (prop.major != 10) {
        g_create_err = "device is not sm_100 (Blackwell B200); the kernels are built for sm_100a only";
        delete h;
        return WRENC_B200_ENODEV;
    }
    h->sm_count = prop.multiProcessorCount;
    h->ctas_per_sm = search_ctas_per_sm();
    if (h->ctas_per_sm <= 0) return fail("search kernel does not fit on the device", cudaGetLastError(), WRENC_B200_ECUDA);
    h->grid = h->sm_count * h->ctas_per_sm;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e, WRENC_B200_ECUDA);
    if ((e = cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e, WRENC_B200_ECUDA);
    if ((e = cudaMalloc(&h->d_tab, sizeof(DevTables))) != cudaSuccess) return fail("cudaMalloc", e, WRENC_B200_ECUDA);
    if ((e = cudaMemcpy(h->d_tab, &h->hc.t, sizeof(DevTables), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e, WRENC_B200_ECUDA);
    if ((e = cudaMalloc(&h->d_counter, sizeof(unsigned int))) != cudaSuccess) return fail("cudaMalloc", e, WRENC_B200_ECUDA);
    *out = h;
    return WRENC_B200_OK;
}

void wrenc_b200_destroy(wrenc_b200 *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_tab); cudaFree(h->d_root_slots); cudaFree(h->d_mode_map); cudaFree(h->d_done); cudaFree(h->d_items); cudaFree(h->d_counter);
    cudaFree(h->d_orig); cudaFree(h->d_rec); cudaFree(h->d_lev); cudaFree(h->d_rec_ctu);
    cudaFree(h->d_bins); cudaFree(h->d_stage); cudaFree(h->d_bin_count); cudaFree(h->d_bin_offset); cudaFree(h->d_bin_total); cudaFree(h->d_out); cudaFree(h->d_out_len);
    cudaFreeHost(h->h_out); cudaFreeHost(h->h_out_len);
    cudaFreeHost(h->h_orig); cudaFreeHost(h->h_rec); cudaFreeHost(h->h_lev); cudaFreeHost(h->h_records);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

static int ensure_batch_buffers(wrenc_b200 *h) {
    if (h->d_orig) return 0;
    CK(cudaSetDevice(h->cfg.device));
    const size_t B = h->B, ps = h->pic_samples, nctu = (size_t)h->Wc * h->Hc;
    CK(cudaMalloc(&h->d_orig, B * ps));
    CK(cudaMalloc(&h->d_rec, B * ps));
    CK(cudaMalloc(&h->d_lev, B * ps * sizeof(int16_t)));
    CK(cudaMalloc(&h->d_rec_ctu, B * nctu * sizeof(CtuRecord)));
    CK(cudaHostAlloc(&h->h_orig, B * ps, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h->h_records, B * nctu * sizeof(CtuRecord), cudaHostAllocDefault));
    if (h->cfg.want_recon) CK(cudaHostAlloc(&h->h_rec, B * ps, cudaHostAllocDefault));
    if (h->cfg.want_decisions) CK(cudaHostAlloc(&h->h_lev, B * ps * sizeof(int16_t), cudaHostAllocDefault));
    h->pic_ids.assign(B, 0);
    return 0;
}

int wrenc_b200_submit(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr) {
    if (!h || !y || !cb || !cr) return WRENC_B200_EINVAL;
    int rc = ensure_batch_buffers(h);
    if (rc) return rc;
    if (h->launched || h->n_filled >= h->B) {
        h->err = "pictures_in_flight pictures are pending; call wrenc_b200_receive first";
        return WRENC_B200_EFULL;
    }
    CK(cudaSetDevice(h->cfg.device));
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4;
    uint8_t *dst = h->h_orig + (size_t)h->n_filled * ps;
    memcpy(dst, y, ny);
    memcpy(dst + ny, cb, nc);
    memcpy(dst + ny + nc, cr, nc);
    CK(cudaMemcpyAsync(h->d_orig + (size_t)h->n_filled * ps, dst, ps, cudaMemcpyHostToDevice, h->stream));
    h->pic_ids[h->n_filled] = pic_idx;
    h->n_filled++;
    return WRENC_B200_OK;
}

int wrenc_b200_submit_pinned(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr) {
    if (!h || !y || !cb || !cr) return WRENC_B200_EINVAL;
    int rc = ensure_batch_buffers(h);
    if (rc) return rc;
    if (h->launched || h->n_filled >= h->B) {
        h->err = "pictures_in_flight pictures are pending; call wrenc_b200_receive first";
        return WRENC_B200_EFULL;
    }
    CK(cudaSetDevice(h->cfg.device));
    const uint8_t *src[3] = {y, cb, cr};
    for (int i = 0; i < 3; i++) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, src[i]) != cudaSuccess || a.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            h->err = "wrenc_b200_submit_pinned: plane is not in page-locked host memory";
            return WRENC_B200_EINVAL;
        }
    }
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4;
    uint8_t *dst = h->d_orig + (size_t)h->n_filled * ps;
    CK(cudaMemcpyAsync(dst, y, ny, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dst + ny, cb, nc, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dst + ny + nc, cr, nc, cudaMemcpyHostToDevice, h->stream));
    h->pic_ids[h->n_filled] = pic_idx;
    h->n_filled++;
    return WRENC_B200_OK;
}

int wrenc_b200_flush(wrenc_b200 *h) {
    if (!h) return WRENC_B200_EINVAL;
    if (h->launched || h->n_filled == 0) return WRENC_B200_OK;
    CK(cudaSetDevice(h->cfg.device));
    const int n = h->n_filled;
    const size_t ps = h->pic_samples, nctu = (size_t)h->Wc * h->Hc;
    int rc = enqueue_search(h, n, h->d_orig, h->d_rec, h->d_lev, h->d_rec_ctu, h->stream);
    if (rc) return rc;
    if (h->cfg.want_slice_data) {
        rc = ensure_coder(h, h->B);
        if (rc) return rc;
        if (!h->h_out) {
            CK(cudaHostAlloc(&h->h_out, (size_t)h->B * h->out_cap, cudaHostAllocDefault));
            CK(cudaHostAlloc(&h->h_out_len, (size_t)h->B * sizeof(int), cudaHostAllocDefault));
        }
        rc = enqueue_coder(h, n, h->d_lev, h->d_rec_ctu, h->d_out, h->out_cap, h->d_out_len, h->stream);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->h_out_len, h->d_out_len, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));  // lengths first, then exactly the coded bytes of every picture
        for (int i = 0; i < n; i++)
            if (h->h_out_len[i] > 0)
                CK(cudaMemcpyAsync(h->h_out + (size_t)i * h->out_cap, h->d_out + (size_t)i * h->out_cap, (size_t)h->h_out_len[i], cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaMemcpyAsync(h->h_records, h->d_rec_ctu, n * nctu * sizeof(CtuRecord), cudaMemcpyDeviceToHost, h->stream));
    if (h->cfg.want_recon) CK(cudaMemcpyAsync(h->h_rec, h->d_rec, n * ps, cudaMemcpyDeviceToHost, h->stream));
    if (h->cfg.want_decisions) CK(cudaMemcpyAsync(h->h_lev, h->d_lev, n * ps * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev_done, h->stream));
    h->launched = true;
    h->n_returned = 0;
    return WRENC_B200_OK;
}

int wrenc_b200_receive(wrenc_b200 *h, uint64_t *pic_idx, const uint8_t **slice_data, size_t *len, const uint8_t **rec_y, const uint8_t **rec_cb,
                       const uint8_t **rec_cr) {
    if (!h) return WRENC_B200_EINVAL;
    if (!h->launched) {
        if (h->n_filled == 0) {
            h->err = "nothing submitted";
            return WRENC_B200_EAGAIN;
        }
        int rc = wrenc_b200_flush(h);
        if (rc) return rc;
    }
    CK(cudaSetDevice(h->cfg.device));
    if (h->n_returned == 0) CK(cudaEventSynchronize(h->ev_done));
    const int i = h->n_returned;
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4;
    if (pic_idx) *pic_idx = h->pic_ids[i];
    if (slice_data) *slice_data = nullptr;
    if (len) *len = 0;
    if (h->cfg.want_slice_data) {
        if (h->h_out_len[i] < 0) {
            h->err = "slice_data coder overflow (bin arena or output buffer too small for this picture)";
            return WRENC_B200_EOVERFLOW;
        }
        if (slice_data) *slice_data = h->h_out + (size_t)i * h->out_cap;
        if (len) *len = (size_t)h->h_out_len[i];
    }
    const uint8_t *r = h->cfg.want_recon ? h->h_rec + (size_t)i * ps : nullptr;
    if (rec_y) *rec_y = r;
    if (rec_cb) *rec_cb = r ? r + ny : nullptr;
    if (rec_cr) *rec_cr = r ? r + ny + nc : nullptr;
    h->last_returned = i;
    h->n_returned++;
    if (h->n_returned == h->n_filled) {
        h->launched = false;
        h->n_filled = 0;
    }
    return WRENC_B200_OK;
}

int wrenc_b200_decisions(wrenc_b200 *h, const wrenc_b200_ctu_record **records, const int16_t **lev_y, const int16_t **lev_cb, const int16_t **lev_cr) {
    if (!h) return WRENC_B200_EINVAL;
    if (h->last_returned < 0) {
        h->err = "no picture received yet";
        return WRENC_B200_EAGAIN;
    }
    const int i = h->last_returned;
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4, nctu = (size_t)h->Wc * h->Hc;
    if (records) *records = reinterpret_cast<const wrenc_b200_ctu_record *>(h->h_records + (size_t)i * nctu);
    const int16_t *l = h->cfg.want_decisions ? h->h_lev + (size_t)i * ps : nullptr;
    if (lev_y) *lev_y = l;
    if (lev_cb) *lev_cb = l ? l + ny : nullptr;
    if (lev_cr) *lev_cr = l ? l + ny + nc : nullptr;
    return WRENC_B200_OK;
}

int wrenc_b200_pending(const wrenc_b200 *h) { return h ? h->n_filled - (h->launched ? h->n_returned : 0) : 0; }

int wrenc_b200_search_resident(wrenc_b200 *h, int32_t n_pictures, const uint8_t *d_yuv, uint8_t *d_rec, int16_t *d_levels,
                               wrenc_b200_ctu_record *d_records, void *stream) {
    if (!h || n_pictures <= 0 || n_pictures > 65535 || !d_yuv || !d_rec || !d_levels || !d_records) return WRENC_B200_EINVAL;
    if (((uintptr_t)d_yuv & 15) || ((uintptr_t)d_rec & 3) || ((uintptr_t)d_levels & 3)) {
        h->err = "d_yuv must be 16-byte aligned (TMA bulk copies), d_rec / d_levels 4-byte aligned";
        return WRENC_B200_EINVAL;
    }
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    // workspace (re)allocation and work-list upload happen on the handle's stream and are synchronised there
    int rc = enqueue_search(h, n_pictures, d_yuv, d_rec, d_levels, reinterpret_cast<CtuRecord *>(d_records), st);
    if (rc) return rc;
    return 1;
}

int wrenc_b200_code_resident(wrenc_b200 *h, int32_t n_pictures, const int16_t *d_levels, const wrenc_b200_ctu_record *d_records, uint8_t *d_out,
                             size_t out_cap, int32_t *d_out_len, void *stream) {
    if (!h || n_pictures <= 0 || !d_levels || !d_records || !d_out || !d_out_len || out_cap == 0) return WRENC_B200_EINVAL;
    if (n_pictures > h->ws_pics) {
        h->err = "wrenc_b200_code_resident must follow wrenc_b200_search_resident of the same pictures on this handle";
        return WRENC_B200_EINVAL;
    }
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    int rc = enqueue_coder(h, n_pictures, d_levels, reinterpret_cast<const CtuRecord *>(d_records), d_out, out_cap, d_out_len, st);
    return rc ? rc : 5;
}

size_t wrenc_b200_workspace_bytes(const wrenc_b200 *h, int32_t n) {
    if (!h || n <= 0) return 0;
    // mode map + done flags + work list per picture, the CTU scratch of the grid (independent of n), tables, counter
    return (size_t)(h->W / 4) * (h->H / 4) * n + (size_t)h->Wc * h->Hc * n * 8 + (size_t)h->grid * search_ctus_per_cta() * CTU_SCRATCH_BYTES +
           sizeof(DevTables) + 4;
}

// nal.rs:210-299 (write_byte_stream_nal_unit_bins + write_nal_unit_bins), host only
int64_t wrenc_b200_write_nal(int32_t nuh_layer_id, int32_t nal_unit_type, int32_t nuh_temporal_id, const uint8_t *payload, size_t len,
                             uint8_t *out, size_t cap) {
    if ((len && !payload) || nuh_layer_id < 0 || nuh_layer_id > 63 || nal_unit_type < 0 || nal_unit_type > 31 || nuh_temporal_id < 0 ||
        nuh_temporal_id > 6)
        return WRENC_B200_EINVAL;
    // size first: every 00 00 0x found while idx + 3 < len costs one extra byte
    size_t extra = 0;
    for (size_t idx = 0; idx + 3 < len;) {
        if (payload[idx] == 0 && payload[idx + 1] == 0 && payload[idx + 2] <= 3) { extra++; idx += 2; }
        else idx++;
    }
    const size_t need = 6 + 2 + len + extra;
    if (!out || cap < need) return -(int64_t)need;
    size_t n = 0;
    out[n++] = 0; out[n++] = 0; out[n++] = 0;   // "header_bytes" (nal.rs:218)
    out[n++] = 0; out[n++] = 0; out[n++] = 1;   // start_code_prefix_one_3bytes
    out[n++] = (uint8_t)(nuh_layer_id & 63);    // forbidden_zero_bit, nuh_reserved_zero_bit, nuh_layer_id
    out[n++] = (uint8_t)((nal_unit_type << 3) | ((nuh_temporal_id + 1) & 7));
    size_t idx = 0;
    while (idx + 3 < len) {
        if (payload[idx] == 0 && payload[idx + 1] == 0 && payload[idx + 2] <= 3) {
            out[n++] = 0; out[n++] = 0; out[n++] = 3;  // the two zeros, then emulation_prevention_three_byte
            idx += 2;
        } else {
            out[n++] = payload[idx++];
        }
    }
    while (idx < len) out[n++] = payload[idx++];
    return (int64_t)n;
}

int wrenc_b200_get_consts(const wrenc_b200 *h, wrenc_b200_consts *out) {
    if (!h || !out) return WRENC_B200_EINVAL;
    out->lambda_q = h->hc.lambda_q;
    out->lambda_rd = h->hc.t.lambda_rd;
    out->lambda_rd_chroma = h->hc.t.lambda_rd_c;
    out->ls = h->hc.t.ls;
    for (int i = 0; i < 8; i++) { out->lv[i] = h->hc.lv64[i]; out->dq[i] = h->hc.dq64[i]; }
    return WRENC_B200_OK;
}

int wrenc_b200_derive_consts(int32_t qp, const char *extra_params, wrenc_b200_consts *out, int64_t *hdr_single /*67*4 or NULL*/,
                             int64_t *hdr_dual /*67 or NULL*/, int64_t *hdr_chroma /*4 or NULL*/) {
    if (!out || qp < 0 || qp > 63) return WRENC_B200_EINVAL;
    Tuning t;
    HostConsts hc;
    if (!t.parse(extra_params, g_create_err) || !hc.init(qp, t, g_create_err)) return WRENC_B200_EINVAL;
    out->lambda_q = hc.lambda_q;
    out->lambda_rd = hc.t.lambda_rd;
    out->lambda_rd_chroma = hc.t.lambda_rd_c;
    out->ls = hc.t.ls;
    for (int i = 0; i < 8; i++) { out->lv[i] = hc.lv64[i]; out->dq[i] = hc.dq64[i]; }
    if (hdr_single) for (int i = 0; i < 67; i++) for (int j = 0; j < 4; j++) hdr_single[i * 4 + j] = hc.t.hdr_single[i][j];
    if (hdr_dual) for (int i = 0; i < 67; i++) hdr_dual[i] = hc.t.hdr_dual[i];
    if (hdr_chroma) for (int i = 0; i < 4; i++) hdr_chroma[i] = hc.t.hdr_chroma[i];
    return WRENC_B200_OK;
}

// ---- per-block entry points (host pointers; parity tests of the block kernels) ----
static int run_block(wrenc_b200 *h, BlockParams &P, const void *in, size_t in_bytes, void *out, size_t out_bytes, int *outi, int n_outi, bool is_pred) {
    CK(cudaSetDevice(h->cfg.device));
    void *d_in = nullptr, *d_out = nullptr;
    int *d_i = nullptr;
    CK(cudaMalloc(&d_in, in_bytes));
    CK(cudaMalloc(&d_out, out_bytes));
    CK(cudaMalloc(&d_i, sizeof(int) * std::max(1, n_outi)));
    CK(cudaMemcpy(d_in, in, in_bytes, cudaMemcpyHostToDevice));
    if (is_pred) { P.rec = (const uint8_t *)d_in; P.out8 = (uint8_t *)d_out; }
    else { P.in = (const int16_t *)d_in; P.out16 = (int16_t *)d_out; }
    P.outi = d_i;
    P.tab = h->d_tab;
    cudaError_t e = launch_block(P, is_pred ? 1 : std::min(P.count, 1024), h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && outi) e = cudaMemcpy(outi, d_i, sizeof(int) * n_outi, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_i);
    if (e != cudaSuccess) { h->err = std::string("block kernel: ") + cudaGetErrorString(e); return WRENC_B200_ECUDA; }
    return WRENC_B200_OK;
}

int wrenc_b200_block_predict(wrenc_b200 *h, const uint8_t *rec_i420, int x, int y, int w, int tree, int ar, int bl, int c, int mode, uint8_t *pred) {
    if (!h || !rec_i420 || !pred || (w != 4 && w != 8 && w != 16 && w != 32) || x < 0 || y < 0 || x + w > h->W || y + w > h->H || (x % w) || (y % w) || c < 0 || c > 2 ||
        (c > 0 && w < 8) || mode < 0 || (mode > 66 && (mode < 81 || mode > 83 || c == 0)))
        return WRENC_B200_EINVAL;
    BlockParams P{};
    P.op = 0; P.W = h->W; P.H = h->H; P.x = x; P.y = y; P.w = w; P.tree = tree; P.ar = ar; P.bl = bl; P.c = c; P.mode = mode;
    int n = c ? w / 2 : w;
    return run_block(h, P, rec_i420, h->pic_samples, pred, (size_t)n * n, nullptr, 0, true);
}

static int block_i16(wrenc_b200 *h, int op, const int16_t *in, int log2n, int count, int16_t *out, int *rates) {
    if (!h || !in || !out || log2n < 2 || log2n > 5 || count <= 0) return WRENC_B200_EINVAL;
    BlockParams P{};
    P.op = op; P.l2 = log2n; P.count = count;
    size_t bytes = ((size_t)count << (2 * log2n)) * sizeof(int16_t);
    return run_block(h, P, in, bytes, out, bytes, rates, rates ? count : 0, false);
}
int wrenc_b200_block_fwd_dct(wrenc_b200 *h, const int16_t *res, int log2n, int count, int16_t *coef) { return block_i16(h, 1, res, log2n, count, coef, nullptr); }
int wrenc_b200_block_inv_dct(wrenc_b200 *h, const int16_t *deq, int log2n, int count, int16_t *out) { return block_i16(h, 2, deq, log2n, count, out, nullptr); }
int wrenc_b200_block_quantize(wrenc_b200 *h, const int16_t *coef, int log2n, int count, int16_t *levels, int32_t *rates) {
    return block_i16(h, 3, coef, log2n, count, levels, rates);
}
int wrenc_b200_block_dequantize(wrenc_b200 *h, const int16_t *levels, int log2n, int count, int16_t *out) { return block_i16(h, 4, levels, log2n, count, out, nullptr); }

}  // extern "C"
