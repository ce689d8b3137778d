// api.cu — the C ABI of include/wrenc_b200.h: handle, picture batching, work-list construction, launches, copies.
// Host-side counterpart of the reference's per-picture driver (src/main.rs:294-402) for the part that moves behind the FFI.
//
// Host-plane path (submit / receive): a ring of batch slots, each with its own picture, decision and coder buffers.  Pictures
// are copied to the device as they are submitted (copy stream); a slot that fills up is launched at once: search kernel on the
// search stream, syntax + CABAC kernels on the coder stream, device-to-host copies on their own stream, chained by events.  The
// search of batch n+1 therefore overlaps the H2D copies of batch n+2 and the D2H copies of batch n, and receive returns picture i
// as soon as ITS bytes have landed.  (The coder kernels of batch n are only ORDERED independently of the next search: on one GPU they
// cannot run under it, because the search grid holds every register of every SM - DESIGN.md section 1.)  Nothing in the launch path
// synchronises with the host.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wrenc_b200.h"
#include "search_kernel_api.h"

using namespace wb;

static thread_local std::string g_create_err;

// Buffers one search + coder pass needs besides the pictures: final luma modes per 4x4, wavefront flags, work-list cursor and
// the slice coder's per-CTU counts / staging slots / bin arena.
struct Workspace {
    int pics = 0;
    uint8_t *d_mode_map = nullptr;
    int *d_done = nullptr;
    unsigned int *d_counter = nullptr;
    int epoch = 0;
    // slice_data coder
    int coder_pics = 0;
    uint16_t *d_bins = nullptr;
    size_t bins_cap = 0;
    NzMap *d_nzmap = nullptr;     // per-CTU map of the non-zero 4x4 level blocks
    uint16_t *d_stage = nullptr;  // per-CTU staging slots of the bin strings (stage_cap() entries each)
    int *d_bin_count = nullptr;
    unsigned long long *d_bin_offset = nullptr, *d_bin_total = nullptr;
};

struct ItemList {
    uint32_t *d_items = nullptr;
    int n_items = 0;
};

// One batch of the host-plane path.
struct Slot {
    Workspace ws;
    uint8_t *d_orig = nullptr, *d_rec = nullptr;
    int16_t *d_lev = nullptr;
    CtuRecord *d_rec_ctu = nullptr;
    uint8_t *d_out = nullptr;
    int *d_out_len = nullptr;
    // pinned host side
    uint8_t *h_orig = nullptr, *h_rec = nullptr;
    int16_t *h_lev = nullptr;
    CtuRecord *h_records = nullptr;
    uint8_t *h_out = nullptr;  // the pictures' slice_data back to back
    size_t h_out_cap = 0;
    int *h_out_len = nullptr;
    unsigned long long *h_bin_total = nullptr;
    std::vector<size_t> out_off;
    std::vector<uint64_t> pic_ids;
    std::vector<cudaEvent_t> ev_pic;  // picture i's slice_data, records, reconstruction and levels have landed
    cudaEvent_t ev_h2d = nullptr, ev_search = nullptr, ev_len = nullptr, ev_fixed = nullptr;
    int n_filled = 0, n_returned = 0;
    bool launched = false, bytes_enqueued = false, allocated = false;
};

struct wrenc_b200 {
    wrenc_b200_config cfg{};
    std::string extra;
    std::string err;
    HostConsts hc;
    int W = 0, H = 0, Wc = 0, Hc = 0, B = 1;
    size_t pic_samples = 0, out_cap = 0;
    int sm_count = 0, ctas_per_sm = 0, grid = 0;
    cudaStream_t stream = nullptr;  // search stream (also the default stream of the resident entry points)
    cudaStream_t st_h2d = nullptr, st_coder = nullptr, st_d2h = nullptr, st_bytes = nullptr;  // copies in, coder, fixed-size outputs, slice_data bytes
    DevTables *d_tab = nullptr;
    uint8_t *d_root_slots = nullptr;  // per CTA x lock-step CTU: candidate slots + saved no-split states (one search at a time per handle)
    std::map<int, ItemList> items;    // wavefront work list per n_pictures: uploaded once, never overwritten while a kernel may read it
    Workspace rws;                    // workspace of the resident entry points
    std::vector<Slot> slots;
    int fill = 0, recv = 0;           // slot being filled / slot being received
    int last_slot = -1, last_pic = -1;
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

static void free_workspace(Workspace &w) {
    cudaFree(w.d_mode_map); cudaFree(w.d_done); cudaFree(w.d_counter);
    cudaFree(w.d_bins); cudaFree(w.d_stage); cudaFree(w.d_nzmap); cudaFree(w.d_bin_count); cudaFree(w.d_bin_offset); cudaFree(w.d_bin_total);
    w = Workspace();
}

// (Re)allocation synchronises the device (cudaFree / cudaMalloc); it happens when n_pics grows, never in steady state.
static int ensure_workspace(wrenc_b200 *h, Workspace &w, int n_pics) {
    if (n_pics <= w.pics) return 0;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaDeviceSynchronize());
    cudaFree(w.d_mode_map); cudaFree(w.d_done);
    w.d_mode_map = nullptr; w.d_done = nullptr;
    const size_t nctu = (size_t)h->Wc * h->Hc * n_pics;
    CK(cudaMalloc(&w.d_mode_map, (size_t)(h->W / 4) * (h->H / 4) * n_pics));
    CK(cudaMalloc(&w.d_done, nctu * sizeof(int)));
    CK(cudaMemset(w.d_done, 0, nctu * sizeof(int)));
    CK(cudaMemset(w.d_mode_map, 0, (size_t)(h->W / 4) * (h->H / 4) * n_pics));
    if (!w.d_counter) CK(cudaMalloc(&w.d_counter, sizeof(unsigned int)));
    if (!h->d_root_slots) CK(cudaMalloc(&h->d_root_slots, (size_t)h->grid * search_ctus_per_cta() * CTU_SCRATCH_BYTES));
    w.pics = n_pics;
    w.epoch = 0;
    return 0;
}

// Work list in wavefront order.  A CTU (x,y) depends on (x-1,y) and (x+1,y-1); both have a smaller key x+2y, so every
// dependency of an item precedes it in the list and a persistent grid that hands items out in list order cannot
// deadlock.  One list per n_pictures, uploaded with a blocking copy the first time that batch size is seen and kept until the
// handle is destroyed: a list a running kernel may still be reading is never overwritten.
static int get_items(wrenc_b200 *h, int n_pics, ItemList &out) {
    auto it = h->items.find(n_pics);
    if (it != h->items.end()) { out = it->second; return 0; }
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
    ItemList L;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMalloc(&L.d_items, items.size() * sizeof(uint32_t)));
    CK(cudaMemcpy(L.d_items, items.data(), items.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    L.n_items = (int)items.size();
    h->items[n_pics] = L;
    out = L;
    return 0;
}

static int enqueue_search(wrenc_b200 *h, Workspace &w, int n_pics, const uint8_t *d_yuv, uint8_t *d_rec, int16_t *d_lev, CtuRecord *d_records, cudaStream_t st) {
    int rc = ensure_workspace(h, w, n_pics);
    if (rc) return rc;
    ItemList L;
    rc = get_items(h, n_pics, L);
    if (rc) return rc;
    SearchParams P;
    P.W = h->W; P.H = h->H; P.Wc = h->Wc; P.Hc = h->Hc;
    P.max_depth = h->cfg.max_split_depth;
    P.n_items = L.n_items;
    P.epoch = ++w.epoch;
    P.orig = d_yuv; P.rec = d_rec; P.lev = d_lev; P.mode_map = w.d_mode_map; P.records = d_records;
    P.done = w.d_done; P.items = L.d_items; P.counter = w.d_counter; P.tab = h->d_tab;
    P.root_slots = h->d_root_slots;
    CK(cudaMemsetAsync(w.d_counter, 0, sizeof(unsigned int), st));
    const int grid = std::min(h->grid, L.n_items / search_ctus_per_cta());
    // The CTU scratch (40 MB for 148 x 8 CTUs) is rewritten for every CU: keep it resident in L2 (persisting access window,
    // a per-launch attribute) so that it is not written back to HBM over and over.
    CK(launch_search(P, grid, st, h->d_root_slots, (size_t)h->grid * search_ctus_per_cta() * CTU_SCRATCH_BYTES));
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
// WRENC_B200_ARENA_ENTRIES_PER_CTU: dev-time override of the initial bin-arena size (lets the tests force the grow-and-retry path)
static size_t arena_entries_per_ctu() {
    const char *e = getenv("WRENC_B200_ARENA_ENTRIES_PER_CTU");
    return e && atol(e) > 0 ? (size_t)atol(e) : 640;
}

static int ensure_coder(wrenc_b200 *h, Workspace &w, int n_pics) {
    if (n_pics <= w.coder_pics) return 0;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaDeviceSynchronize());
    cudaFree(w.d_bin_count); cudaFree(w.d_bin_offset); cudaFree(w.d_stage); cudaFree(w.d_nzmap);
    w.d_bin_count = nullptr; w.d_bin_offset = nullptr; w.d_stage = nullptr; w.d_nzmap = nullptr;
    const size_t nctu = (size_t)h->Wc * h->Hc * n_pics;
    CK(cudaMalloc(&w.d_bin_count, nctu * sizeof(int)));
    CK(cudaMalloc(&w.d_stage, nctu * stage_cap() * sizeof(uint16_t)));
    CK(cudaMalloc(&w.d_nzmap, nctu * sizeof(NzMap)));
    CK(cudaMalloc(&w.d_bin_offset, nctu * sizeof(unsigned long long)));
    if (!w.d_bin_total) CK(cudaMalloc(&w.d_bin_total, sizeof(unsigned long long)));
    if (w.bins_cap < nctu * arena_entries_per_ctu()) {  // first guess; grown from the measured totals (grow_arena)
        cudaFree(w.d_bins);
        w.d_bins = nullptr;
        w.bins_cap = nctu * arena_entries_per_ctu();
        CK(cudaMalloc(&w.d_bins, w.bins_cap * sizeof(uint16_t)));
    }
    w.coder_pics = n_pics;
    return 0;
}

static int grow_arena(wrenc_b200 *h, Workspace &w, unsigned long long total) {
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaDeviceSynchronize());
    cudaFree(w.d_bins);
    w.d_bins = nullptr;
    w.bins_cap = (size_t)(total + total / 4 + 1024);
    CK(cudaMalloc(&w.d_bins, w.bins_cap * sizeof(uint16_t)));
    return 0;
}

// CABAC-code the pictures the last search with this workspace decided (its mode map).  Pass 1 counts the bins of every CTU
// and stages the short strings, an exclusive scan turns the counts into arena offsets (and leaves the total in d_bin_total),
// pass 2 writes the long strings, the compaction moves the staged ones, then three warps per picture run the arithmetic coder (states, interval widths, code value).
// No host synchronisation: the arena is sized from earlier batches; pictures whose strings would not fit report
// out_len = -2 and the caller grows the arena (d_bin_total) and calls again with first_pass = false.
static int enqueue_coder(wrenc_b200 *h, Workspace &w, int n_pics, const int16_t *d_lev, const CtuRecord *d_records, uint8_t *d_out, size_t out_cap, int *d_out_len,
                         cudaStream_t st, bool first_pass = true) {
    int rc = ensure_coder(h, w, n_pics);
    if (rc) return rc;
    SyntaxParams Q;
    Q.W = h->W; Q.H = h->H; Q.Wc = h->Wc; Q.Hc = h->Hc; Q.n_pics = n_pics; Q.qp = h->cfg.qp;
    Q.lev = d_lev; Q.records = d_records; Q.mode_map = w.d_mode_map; Q.nzmap = w.d_nzmap;
    Q.bins = nullptr; Q.bins_cap = w.bins_cap; Q.bin_count = w.d_bin_count; Q.bin_offset = w.d_bin_offset;
    Q.stage = w.d_stage; Q.stage_cap = stage_cap();
    Q.out = d_out; Q.out_cap = out_cap; Q.out_len = d_out_len;
    if (first_pass) {
        CK(launch_syntax(Q, st));
        CK(launch_bin_scan(Q, w.d_bin_total, st));
        h->launches += 1 + syntax_first_pass_kernels();
    }
    Q.bins = w.d_bins;
    CK(launch_syntax(Q, st));       // only the CTUs that did not fit their staging slot
    CK(launch_bin_compact(Q, st));  // everybody else: staged string -> arena offset
    CK(launch_cabac(Q, st));
    h->launches += 3;
    return 0;
}

static void free_slot(Slot &s) {
    free_workspace(s.ws);
    cudaFree(s.d_orig); cudaFree(s.d_rec); cudaFree(s.d_lev); cudaFree(s.d_rec_ctu); cudaFree(s.d_out); cudaFree(s.d_out_len);
    cudaFreeHost(s.h_orig); cudaFreeHost(s.h_rec); cudaFreeHost(s.h_lev); cudaFreeHost(s.h_records); cudaFreeHost(s.h_out); cudaFreeHost(s.h_out_len);
    cudaFreeHost(s.h_bin_total);
    for (cudaEvent_t e : s.ev_pic) cudaEventDestroy(e);
    if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
    if (s.ev_search) cudaEventDestroy(s.ev_search);
    if (s.ev_len) cudaEventDestroy(s.ev_len);
    if (s.ev_fixed) cudaEventDestroy(s.ev_fixed);
    s = Slot();
}

extern "C" {

const char *wrenc_b200_version(void) { return "wrenc_b200 0.2 (sm_100a)"; }

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
    h->out_cap = h->pic_samples * 2;  // slice_data of a picture: twice its raw size (noise at QP < 6 codes to more than 8 bits per sample)
    int nslots = 2;  // batches in flight: one being searched while the other is filled / coded / received
    if (const char *e = getenv("WRENC_B200_SLOTS")) nslots = std::min(8, std::max(1, atoi(e)));
    h->slots.resize(nslots);
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
    // L2 carve-out for the search kernel's persisting access window: a per-DEVICE limit, set by every handle for its own device
    // (best effort: without it the window is ignored and the scratch only costs DRAM traffic)
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)h->grid * search_ctus_per_cta() * CTU_SCRATCH_BYTES + (4u << 20));
    cudaGetLastError();
    cudaStream_t *sts[5] = {&h->stream, &h->st_h2d, &h->st_coder, &h->st_d2h, &h->st_bytes};
    for (cudaStream_t *s : sts)
        if ((e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e, WRENC_B200_ECUDA);
    if ((e = cudaMalloc(&h->d_tab, sizeof(DevTables))) != cudaSuccess) return fail("cudaMalloc", e, WRENC_B200_ECUDA);
    if ((e = cudaMemcpy(h->d_tab, &h->hc.t, sizeof(DevTables), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e, WRENC_B200_ECUDA);
    *out = h;
    return WRENC_B200_OK;
}

void wrenc_b200_destroy(wrenc_b200 *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (Slot &s : h->slots) free_slot(s);
    free_workspace(h->rws);
    for (auto &kv : h->items) cudaFree(kv.second.d_items);
    cudaFree(h->d_tab); cudaFree(h->d_root_slots);
    cudaStream_t sts[5] = {h->stream, h->st_h2d, h->st_coder, h->st_d2h, h->st_bytes};
    for (cudaStream_t s : sts)
        if (s) cudaStreamDestroy(s);
    cudaCtxResetPersistingL2Cache();  // hand the persisting lines of the scratch back
    cudaGetLastError();
    delete h;
}

static int ensure_slot(wrenc_b200 *h, Slot &s) {
    if (s.allocated) return 0;
    CK(cudaSetDevice(h->cfg.device));
    const size_t B = h->B, ps = h->pic_samples, nctu = (size_t)h->Wc * h->Hc;
    CK(cudaMalloc(&s.d_orig, B * ps));
    CK(cudaMalloc(&s.d_rec, B * ps));
    CK(cudaMalloc(&s.d_lev, B * ps * sizeof(int16_t)));
    CK(cudaMalloc(&s.d_rec_ctu, B * nctu * sizeof(CtuRecord)));
    CK(cudaHostAlloc(&s.h_orig, B * ps, cudaHostAllocDefault));
    CK(cudaHostAlloc(&s.h_records, B * nctu * sizeof(CtuRecord), cudaHostAllocDefault));
    if (h->cfg.want_recon) CK(cudaHostAlloc(&s.h_rec, B * ps, cudaHostAllocDefault));
    if (h->cfg.want_decisions) CK(cudaHostAlloc(&s.h_lev, B * ps * sizeof(int16_t), cudaHostAllocDefault));
    if (h->cfg.want_slice_data) {
        CK(cudaMalloc(&s.d_out, B * h->out_cap));
        CK(cudaMalloc(&s.d_out_len, B * sizeof(int)));
        CK(cudaHostAlloc(&s.h_out_len, B * sizeof(int), cudaHostAllocDefault));
        CK(cudaHostAlloc(&s.h_bin_total, sizeof(unsigned long long), cudaHostAllocDefault));
        s.h_out_cap = std::max<size_t>(B * ps / 8, 1 << 16);  // grown when a batch codes to more
        CK(cudaHostAlloc(&s.h_out, s.h_out_cap, cudaHostAllocDefault));
    }
    CK(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_search, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_len, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_fixed, cudaEventDisableTiming));
    s.ev_pic.resize(B);
    for (size_t i = 0; i < B; i++) CK(cudaEventCreateWithFlags(&s.ev_pic[i], cudaEventDisableTiming));
    s.pic_ids.assign(B, 0);
    s.out_off.assign(B + 1, 0);
    int rc = ensure_workspace(h, s.ws, h->B);
    if (rc) return rc;
    if (h->cfg.want_slice_data && (rc = ensure_coder(h, s.ws, h->B))) return rc;
    s.allocated = true;
    return 0;
}

// Launch the batch in slot s: search after its H2D copies; the fixed-size outputs (records, reconstruction, levels) are copied
// back right after the search on their own stream; the coder runs on the coder stream and leaves the coded lengths on the host.
static int launch_slot(wrenc_b200 *h, Slot &s) {
    if (s.launched || s.n_filled == 0) return 0;
    CK(cudaSetDevice(h->cfg.device));
    const int n = s.n_filled;
    const size_t ps = h->pic_samples, nctu = (size_t)h->Wc * h->Hc;
    CK(cudaEventRecord(s.ev_h2d, h->st_h2d));
    CK(cudaStreamWaitEvent(h->stream, s.ev_h2d, 0));
    int rc = enqueue_search(h, s.ws, n, s.d_orig, s.d_rec, s.d_lev, s.d_rec_ctu, h->stream);
    if (rc) return rc;
    CK(cudaEventRecord(s.ev_search, h->stream));
    CK(cudaStreamWaitEvent(h->st_d2h, s.ev_search, 0));
    for (int i = 0; i < n; i++) {
        CK(cudaMemcpyAsync(s.h_records + (size_t)i * nctu, s.d_rec_ctu + (size_t)i * nctu, nctu * sizeof(CtuRecord), cudaMemcpyDeviceToHost, h->st_d2h));
        if (h->cfg.want_recon) CK(cudaMemcpyAsync(s.h_rec + (size_t)i * ps, s.d_rec + (size_t)i * ps, ps, cudaMemcpyDeviceToHost, h->st_d2h));
        if (h->cfg.want_decisions) CK(cudaMemcpyAsync(s.h_lev + (size_t)i * ps, s.d_lev + (size_t)i * ps, ps * sizeof(int16_t), cudaMemcpyDeviceToHost, h->st_d2h));
        if (!h->cfg.want_slice_data) CK(cudaEventRecord(s.ev_pic[i], h->st_d2h));  // nothing else to wait for: the picture is complete
    }
    CK(cudaEventRecord(s.ev_fixed, h->st_d2h));
    s.bytes_enqueued = !h->cfg.want_slice_data;
    if (h->cfg.want_slice_data) {
        CK(cudaStreamWaitEvent(h->st_coder, s.ev_search, 0));
        rc = enqueue_coder(h, s.ws, n, s.d_lev, s.d_rec_ctu, s.d_out, h->out_cap, s.d_out_len, h->st_coder);
        if (rc) return rc;
        CK(cudaMemcpyAsync(s.h_out_len, s.d_out_len, n * sizeof(int), cudaMemcpyDeviceToHost, h->st_coder));
        CK(cudaMemcpyAsync(s.h_bin_total, s.ws.d_bin_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st_coder));
        CK(cudaEventRecord(s.ev_len, h->st_coder));
    }
    s.launched = true;
    s.n_returned = 0;
    return 0;
}

// Second half of a launched batch, entered once the coded lengths are on the host (receive waits for them; submit only polls):
// grow-and-retry of the bin arena if this batch outgrew it, then per picture the D2H copy of exactly its coded bytes on the
// byte stream (which carries no cross-batch dependency), followed by that picture's event.
static int enqueue_outputs(wrenc_b200 *h, Slot &s, bool wait) {
    if (!s.launched || s.bytes_enqueued) return 0;
    CK(cudaSetDevice(h->cfg.device));
    const int n = s.n_filled;
    for (;;) {
        if (wait) CK(cudaEventSynchronize(s.ev_len));
        else {
            cudaError_t q = cudaEventQuery(s.ev_len);
            if (q == cudaErrorNotReady) return 0;
            CK(q);
        }
        bool arena_short = false;
        for (int i = 0; i < n; i++) arena_short |= s.h_out_len[i] == -2;
        if (!arena_short) break;
        int rc = grow_arena(h, s.ws, *s.h_bin_total);
        if (rc) return rc;
        rc = enqueue_coder(h, s.ws, n, s.d_lev, s.d_rec_ctu, s.d_out, h->out_cap, s.d_out_len, h->st_coder, false);
        if (rc) return rc;
        CK(cudaMemcpyAsync(s.h_out_len, s.d_out_len, n * sizeof(int), cudaMemcpyDeviceToHost, h->st_coder));
        CK(cudaEventRecord(s.ev_len, h->st_coder));
        wait = true;
    }
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        s.out_off[i] = total;
        total += s.h_out_len[i] > 0 ? (size_t)s.h_out_len[i] : 0;
    }
    s.out_off[n] = total;
    if (total > s.h_out_cap) {
        cudaFreeHost(s.h_out);
        s.h_out = nullptr;
        s.h_out_cap = total + total / 4;
        CK(cudaHostAlloc(&s.h_out, s.h_out_cap, cudaHostAllocDefault));
    }
    CK(cudaStreamWaitEvent(h->st_bytes, s.ev_len, 0));
    CK(cudaStreamWaitEvent(h->st_bytes, s.ev_fixed, 0));
    for (int i = 0; i < n; i++) {
        if (s.h_out_len[i] > 0)
            CK(cudaMemcpyAsync(s.h_out + s.out_off[i], s.d_out + (size_t)i * h->out_cap, (size_t)s.h_out_len[i], cudaMemcpyDeviceToHost, h->st_bytes));
        CK(cudaEventRecord(s.ev_pic[i], h->st_bytes));
    }
    s.bytes_enqueued = true;
    return 0;
}

static int submit_common(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr, bool pinned) {
    if (!h || !y || !cb || !cr) return WRENC_B200_EINVAL;
    Slot &s = h->slots[h->fill];
    if (s.launched) {  // every slot holds a batch that has not been received completely
        h->err = "all batch slots hold pictures that have not been received; call wrenc_b200_receive first";
        return WRENC_B200_EFULL;
    }
    int rc = ensure_slot(h, s);
    if (rc) return rc;
    CK(cudaSetDevice(h->cfg.device));
    if (pinned) {
        const uint8_t *src[3] = {y, cb, cr};
        for (int i = 0; i < 3; i++) {
            cudaPointerAttributes a{};
            if (cudaPointerGetAttributes(&a, src[i]) != cudaSuccess || a.type != cudaMemoryTypeHost) {
                cudaGetLastError();
                h->err = "wrenc_b200_submit_pinned: plane is not in page-locked host memory";
                return WRENC_B200_EINVAL;
            }
        }
    }
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4;
    uint8_t *dst = s.d_orig + (size_t)s.n_filled * ps;
    if (pinned) {
        CK(cudaMemcpyAsync(dst, y, ny, cudaMemcpyHostToDevice, h->st_h2d));
        CK(cudaMemcpyAsync(dst + ny, cb, nc, cudaMemcpyHostToDevice, h->st_h2d));
        CK(cudaMemcpyAsync(dst + ny + nc, cr, nc, cudaMemcpyHostToDevice, h->st_h2d));
    } else {
        uint8_t *stg = s.h_orig + (size_t)s.n_filled * ps;
        memcpy(stg, y, ny);
        memcpy(stg + ny, cb, nc);
        memcpy(stg + ny + nc, cr, nc);
        CK(cudaMemcpyAsync(dst, stg, ps, cudaMemcpyHostToDevice, h->st_h2d));
    }
    s.pic_ids[s.n_filled] = pic_idx;
    s.n_filled++;
    if (s.n_filled == h->B) {  // the batch is complete: launch it now and move on to the next slot
        rc = launch_slot(h, s);
        if (rc) return rc;
        h->fill = (h->fill + 1) % (int)h->slots.size();
    }
    // keep the output side of earlier batches moving while the caller is still submitting
    for (Slot &o : h->slots)
        if (o.launched && !o.bytes_enqueued && (rc = enqueue_outputs(h, o, false))) return rc;
    return WRENC_B200_OK;
}

int wrenc_b200_submit(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr) {
    return submit_common(h, pic_idx, y, cb, cr, false);
}
int wrenc_b200_submit_pinned(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr) {
    return submit_common(h, pic_idx, y, cb, cr, true);
}

void *wrenc_b200_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void wrenc_b200_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
    cudaGetLastError();
}

int wrenc_b200_flush(wrenc_b200 *h) {
    if (!h) return WRENC_B200_EINVAL;
    Slot &s = h->slots[h->fill];
    if (s.launched || s.n_filled == 0) return WRENC_B200_OK;
    int rc = launch_slot(h, s);
    if (rc) return rc;
    h->fill = (h->fill + 1) % (int)h->slots.size();
    return WRENC_B200_OK;
}

int wrenc_b200_receive(wrenc_b200 *h, uint64_t *pic_idx, const uint8_t **slice_data, size_t *len, const uint8_t **rec_y, const uint8_t **rec_cb,
                       const uint8_t **rec_cr) {
    if (!h) return WRENC_B200_EINVAL;
    Slot &s = h->slots[h->recv];
    if (!s.launched) {
        if (s.n_filled == 0) {
            h->err = "nothing submitted";
            return WRENC_B200_EAGAIN;
        }
        int rc = wrenc_b200_flush(h);  // the oldest pending pictures sit in a partly filled slot
        if (rc) return rc;
    }
    CK(cudaSetDevice(h->cfg.device));
    int rc = enqueue_outputs(h, s, true);
    if (rc) return rc;
    const int i = s.n_returned;
    CK(cudaEventSynchronize(s.ev_pic[i]));
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4;
    if (pic_idx) *pic_idx = s.pic_ids[i];
    if (slice_data) *slice_data = nullptr;
    if (len) *len = 0;
    int ret = WRENC_B200_OK;
    if (h->cfg.want_slice_data) {
        if (s.h_out_len[i] < 0) {  // the picture is still consumed: decisions and reconstruction stay reachable, the handle stays usable
            h->err = "slice_data of this picture did not fit the output buffer (twice its raw size)";
            ret = WRENC_B200_EOVERFLOW;
        } else {
            if (slice_data) *slice_data = s.h_out + s.out_off[i];
            if (len) *len = (size_t)s.h_out_len[i];
        }
    }
    const uint8_t *r = h->cfg.want_recon ? s.h_rec + (size_t)i * ps : nullptr;
    if (rec_y) *rec_y = r;
    if (rec_cb) *rec_cb = r ? r + ny : nullptr;
    if (rec_cr) *rec_cr = r ? r + ny + nc : nullptr;
    h->last_slot = h->recv;
    h->last_pic = i;
    s.n_returned++;
    if (s.n_returned == s.n_filled) {  // the slot is free again (its host buffers stay untouched until it is launched again)
        s.launched = false;
        s.n_filled = 0;
        h->recv = (h->recv + 1) % (int)h->slots.size();
    }
    return ret;
}

int wrenc_b200_decisions(wrenc_b200 *h, const wrenc_b200_ctu_record **records, const int16_t **lev_y, const int16_t **lev_cb, const int16_t **lev_cr) {
    if (!h) return WRENC_B200_EINVAL;
    if (h->last_slot < 0) {
        h->err = "no picture received yet";
        return WRENC_B200_EAGAIN;
    }
    const Slot &s = h->slots[h->last_slot];
    const int i = h->last_pic;
    const size_t ps = h->pic_samples, ny = (size_t)h->W * h->H, nc = ny / 4, nctu = (size_t)h->Wc * h->Hc;
    if (records) *records = reinterpret_cast<const wrenc_b200_ctu_record *>(s.h_records + (size_t)i * nctu);
    const int16_t *l = h->cfg.want_decisions ? s.h_lev + (size_t)i * ps : nullptr;
    if (lev_y) *lev_y = l;
    if (lev_cb) *lev_cb = l ? l + ny : nullptr;
    if (lev_cr) *lev_cr = l ? l + ny + nc : nullptr;
    return WRENC_B200_OK;
}

int wrenc_b200_pending(const wrenc_b200 *h) {
    if (!h) return 0;
    int n = 0;
    for (const Slot &s : h->slots) n += s.n_filled - (s.launched ? s.n_returned : 0);
    return n;
}

// Allocates the workspace and uploads the work list for n_pictures, so that the following search_resident / code_resident
// calls with at most that many pictures neither allocate nor block.
int wrenc_b200_prepare(wrenc_b200 *h, int32_t n_pictures) {
    if (!h || n_pictures <= 0 || n_pictures > 65535) return WRENC_B200_EINVAL;
    int rc = ensure_workspace(h, h->rws, n_pictures);
    if (rc) return rc;
    ItemList L;
    if ((rc = get_items(h, n_pictures, L))) return rc;
    return ensure_coder(h, h->rws, n_pictures);
}

int wrenc_b200_search_resident(wrenc_b200 *h, int32_t n_pictures, const uint8_t *d_yuv, uint8_t *d_rec, int16_t *d_levels,
                               wrenc_b200_ctu_record *d_records, void *stream) {
    if (!h || n_pictures <= 0 || n_pictures > 65535 || !d_yuv || !d_rec || !d_levels || !d_records) return WRENC_B200_EINVAL;
    if (((uintptr_t)d_yuv & 15) || ((uintptr_t)d_rec & 3) || ((uintptr_t)d_levels & 3)) {
        h->err = "d_yuv must be 16-byte aligned (TMA bulk copies), d_rec / d_levels 4-byte aligned";
        return WRENC_B200_EINVAL;
    }
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    int rc = enqueue_search(h, h->rws, n_pictures, d_yuv, d_rec, d_levels, reinterpret_cast<CtuRecord *>(d_records), st);
    if (rc) return rc;
    return 1;
}

int wrenc_b200_code_resident(wrenc_b200 *h, int32_t n_pictures, const int16_t *d_levels, const wrenc_b200_ctu_record *d_records, uint8_t *d_out,
                             size_t out_cap, int32_t *d_out_len, void *stream) {
    if (!h || n_pictures <= 0 || !d_levels || !d_records || !d_out || !d_out_len || out_cap == 0) return WRENC_B200_EINVAL;
    if (n_pictures > h->rws.pics) {
        h->err = "wrenc_b200_code_resident must follow wrenc_b200_search_resident of the same pictures on this handle";
        return WRENC_B200_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(d_levels) & 15u) {  // the non-zero map kernel reads the level rows with 16-byte loads
        h->err = "wrenc_b200_code_resident: d_levels must be 16-byte aligned";
        return WRENC_B200_EINVAL;
    }
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    int rc = enqueue_coder(h, h->rws, n_pictures, d_levels, reinterpret_cast<const CtuRecord *>(d_records), d_out, out_cap, d_out_len, st);
    return rc ? rc : 4 + syntax_first_pass_kernels();
}

// After a code_resident call whose d_out_len reported -2 (bin arena too small for that batch): grows the arena to the total the
// call measured and codes the same pictures again.  Blocks (reads the total back).  Returns launches enqueued (3) or <0.
int wrenc_b200_code_resident_retry(wrenc_b200 *h, int32_t n_pictures, const int16_t *d_levels, const wrenc_b200_ctu_record *d_records, uint8_t *d_out,
                                   size_t out_cap, int32_t *d_out_len, void *stream) {
    if (!h || n_pictures <= 0 || n_pictures > h->rws.coder_pics || !d_levels || !d_records || !d_out || !d_out_len || out_cap == 0) return WRENC_B200_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, h->rws.d_bin_total, sizeof(total), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (total > h->rws.bins_cap) {
        int rc = grow_arena(h, h->rws, total);
        if (rc) return rc;
    }
    int rc = enqueue_coder(h, h->rws, n_pictures, d_levels, reinterpret_cast<const CtuRecord *>(d_records), d_out, out_cap, d_out_len, st, false);
    return rc ? rc : 3;
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
    // 8x8 TBs have two routines in the kernel (trellis8_chain, used by the search, and the general trellis()): WRENC_B200_BLOCK_GENERAL=1
    // lets the parity tests reach the second one for 8x8 blocks too
    const char *e = getenv("WRENC_B200_BLOCK_GENERAL");
    return block_i16(h, e && atoi(e) ? 5 : 3, coef, log2n, count, levels, rates);
}
int wrenc_b200_block_dequantize(wrenc_b200 *h, const int16_t *levels, int log2n, int count, int16_t *out) { return block_i16(h, 4, levels, log2n, count, out, nullptr); }

}  // extern "C"
