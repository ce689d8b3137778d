// wrenc_b200_cli — the reference's command line (src/main.rs:85-115) over the C ABI: reads raw I420 frames, runs the RD search
// and the slice_data coder on the GPU (wrenc_b200_submit / wrenc_b200_receive, streaming), wraps the returned bytes with the
// parameter-set / picture-header / slice-header writers (wrenc_b200_write_parameter_sets / wrenc_b200_write_picture) and writes
// a complete .vvc byte stream — the same file `wrenc -i IN --input-size WxH --output-size WxH --num-pictures N --qp Q -o OUT`
// writes (main.rs:117-403).  Flags, their meaning and the error behaviour follow the reference: bad sizes / unreadable files
// print "error: ..." and exit 0 (main.rs:127-133,165-185,321-324); only what the reference cannot hit (no GPU) exits 1.
// Extra flag: --pictures-in-flight B (pictures per batch; two batches are in flight).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../include/wrenc_b200.h"

static bool parse_size(const char *s, int *w, int *h) {  // WIDTHxHEIGHT (main.rs:163-187)
    char *end = nullptr;
    long a = strtol(s, &end, 10);
    if (end == s || *end != 'x') return false;
    const char *p = end + 1;
    long b = strtol(p, &end, 10);
    if (end == p || *end != '\0' || a <= 0 || b <= 0) return false;
    *w = (int)a;
    *h = (int)b;
    return true;
}

static void usage() {
    fprintf(stderr,
            "wrenc_b200_cli -i <input|-> -o <output|-> [-r <reconst>] --input-size WxH --output-size WxH --num-pictures N\n"
            "               [--qp Q] [--max-split-depth D] [--extra-params K=V[,K=V...]] [--pictures-in-flight B]\n");
}

int main(int argc, char **argv) {
    std::string input, output, reconst, in_size, out_size, extra;
    long num_pictures = -1, qp = -1, depth = 3, batch = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", name); usage(); exit(2); }
            return argv[++i];
        };
        if (a == "-i" || a == "--input") input = val("--input");
        else if (a == "-o" || a == "--output") output = val("--output");
        else if (a == "-r" || a == "--reconst") reconst = val("--reconst");
        else if (a == "--input-size") in_size = val("--input-size");
        else if (a == "--output-size") out_size = val("--output-size");
        else if (a == "--num-pictures") num_pictures = atol(val("--num-pictures"));
        else if (a == "--qp") qp = atol(val("--qp"));
        else if (a == "--max-split-depth") depth = atol(val("--max-split-depth"));
        else if (a == "--extra-params") extra = val("--extra-params");
        else if (a == "--pictures-in-flight") batch = atol(val("--pictures-in-flight"));
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str()); usage(); return 2; }
    }
    if (input.empty() || output.empty() || in_size.empty() || out_size.empty() || num_pictures < 0) { usage(); return 2; }
    int iw = 0, ih = 0, W = 0, H = 0;
    if (!parse_size(in_size.c_str(), &iw, &ih)) { fprintf(stderr, "error: Invalid input-size: %s\n", in_size.c_str()); return 0; }
    if (!parse_size(out_size.c_str(), &W, &H)) { fprintf(stderr, "error: Invalid output-size: %s\n", out_size.c_str()); return 0; }
    (void)iw; (void)ih;  // like the reference, frames are read with the OUTPUT size (main.rs:318-349)
    FILE *fin = input == "-" ? stdin : fopen(input.c_str(), "rb");
    if (!fin) { fprintf(stderr, "error: failed to open input file: %s\n", input.c_str()); return 0; }
    FILE *fout = output == "-" ? stdout : fopen(output.c_str(), "wb");
    if (!fout) { fprintf(stderr, "error: failed to open output file: %s\n", output.c_str()); return 0; }
    FILE *frec = nullptr;
    if (!reconst.empty() && !(frec = fopen(reconst.c_str(), "wb"))) { fprintf(stderr, "error: failed to open reconst file: %s\n", reconst.c_str()); return 0; }

    wrenc_b200_config cfg{};
    cfg.width = W; cfg.height = H;
    cfg.qp = qp < 0 ? 26 : (int)qp;  // no --qp: QP 26 (ctu.rs:382,1265)
    cfg.max_split_depth = (int)depth;
    cfg.device = getenv("WRENC_B200_DEVICE") ? atoi(getenv("WRENC_B200_DEVICE")) : 0;
    cfg.pictures_in_flight = (int)(batch > 0 ? batch : (num_pictures < 64 ? (num_pictures > 0 ? num_pictures : 1) : 64));
    cfg.want_recon = frec != nullptr;
    cfg.want_decisions = 0;
    cfg.want_slice_data = 1;
    cfg.extra_params = extra.empty() ? nullptr : extra.c_str();
    wrenc_b200 *h = nullptr;
    int rc = wrenc_b200_create(&cfg, &h);
    if (rc == WRENC_B200_EINVAL) { fprintf(stderr, "error: %s\n", wrenc_b200_last_error(nullptr)); return 0; }
    if (rc != 0) { fprintf(stderr, "error: %s\n", wrenc_b200_last_error(nullptr)); return 1; }

    std::vector<uint8_t> buf(4096);
    int64_t n = wrenc_b200_write_parameter_sets(W, H, (int)qp, buf.data(), buf.size());  // VPS, SPS, PPS (main.rs:223-260)
    if (n < 0) { fprintf(stderr, "error: parameter sets\n"); return 1; }
    fwrite(buf.data(), 1, (size_t)n, fout);

    const size_t ny = (size_t)W * H, nc = ny / 4, ps = ny + 2 * nc;
    std::vector<uint8_t> frame(ps);
    long submitted = 0, received = 0;
    bool eof = false;
    auto drain_one = [&]() -> bool {
        uint64_t idx = 0;
        const uint8_t *sd = nullptr, *ry = nullptr, *rcb = nullptr, *rcr = nullptr;
        size_t len = 0;
        int r = wrenc_b200_receive(h, &idx, &sd, &len, &ry, &rcb, &rcr);
        if (r != 0) { fprintf(stderr, "error: %s\n", wrenc_b200_last_error(h)); return false; }
        if (buf.size() < len + len / 2 + 64) buf.resize(len + len / 2 + 64);
        int64_t m = wrenc_b200_write_picture((int)qp, idx, sd, len, buf.data(), buf.size());  // PH NAL + IDR_W_RADL slice NAL (main.rs:297-316,380-389)
        if (m < 0) { fprintf(stderr, "error: picture %llu\n", (unsigned long long)idx); return false; }
        fwrite(buf.data(), 1, (size_t)m, fout);
        if (frec) {  // --reconst (main.rs:391-401)
            fwrite(ry, 1, ny, frec); fwrite(rcb, 1, nc, frec); fwrite(rcr, 1, nc, frec);
        }
        received++;
        return true;
    };
    while (submitted < num_pictures && !eof) {
        if (fread(frame.data(), 1, ps, fin) != ps) {  // the reference prints the read error and exits 0 (main.rs:321-324)
            fprintf(stderr, "error: failed to read picture %ld\n", submitted);
            eof = true;
            break;
        }
        for (;;) {
            rc = wrenc_b200_submit(h, (uint64_t)submitted, frame.data(), frame.data() + ny, frame.data() + ny + nc);
            if (rc != WRENC_B200_EFULL) break;
            if (!drain_one()) return 1;  // both batches are in flight: write the oldest picture first
        }
        if (rc != 0) { fprintf(stderr, "error: %s\n", wrenc_b200_last_error(h)); return 1; }
        submitted++;
    }
    while (received < submitted)
        if (!drain_one()) return 1;
    wrenc_b200_destroy(h);
    if (fout != stdout) fclose(fout); else fflush(fout);
    if (frec) fclose(frec);
    if (fin != stdin) fclose(fin);
    return 0;
}
