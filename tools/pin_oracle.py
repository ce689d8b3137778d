#!/usr/bin/env python3
"""Pin the CPU oracle against the only reference-PRODUCED data in the container.

/root/reference/tools/evaluation/summary.json holds what the real wrenc binary (commit 1d5b5ec) produced for the two CIF
clips in /root/reference/assets at QP 20,23,...,41, --max-split-depth 3 (presets.json): the size of the whole .vvc file in
bytes (`du -b`, wrenc_fixed_qp.sh:4) and, per frame, PSNR-Y/U/V of the VTM-decoded stream against the decoded clip, printed
by ffmpeg's psnr filter with two decimals (psnr.sh:9-11).  This script runs the same 16 encodes with the oracle on the same
input (tools/decode_assets.py: the H.264-decoded yuv420p planes, exact), assembles the byte stream with the product's header
writers (wrenc_b200_write_parameter_sets / wrenc_b200_write_picture) and compares:
  * total file bytes (exact equality expected),
  * all 30 x 3 per-frame PSNR values per operating point at the two printed decimals.
The result is written to tests/golden/reference_pin.json (committed); tests/test_reference_pin.py re-checks a subset live and
the committed table in full.  Test infrastructure, not product code.

Usage: pin_oracle.py [--jobs N] [--points bus:32,mobile:32,...] [--out FILE]
"""
import argparse
import ctypes as C
import hashlib
import json
import math
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
W, H, FRAMES = 352, 288, 30
QPS = [20, 23, 26, 29, 32, 35, 38, 41]
SUMMARY = "/root/reference/tools/evaluation/summary.json"
CLIP_FILES = {"bus": "bus_352x288_30fps_30fr.mp4", "mobile": "mobile_352x288_30fps_30fr.mp4"}


def clip_path(name):
    return os.path.join(ROOT, "tests", "golden", "_assets", name + "_cif.yuv")


def load_clip(name):
    raw = np.fromfile(clip_path(name), np.uint8).reshape(FRAMES, -1)
    n = W * H
    return [(r[:n].reshape(H, W), r[n:n * 5 // 4].reshape(H // 2, W // 2), r[n * 5 // 4:].reshape(H // 2, W // 2)) for r in raw]


def ffmpeg_psnr_2dp(rec, org):
    """ffmpeg vf_psnr: mse = sum((a-b)^2) / (w*h) per plane; psnr = 10*log10(255^2 / mse); printed with %0.2f (inf if mse == 0)."""
    d = rec.astype(np.int64) - org.astype(np.int64)
    sse = int((d * d).sum())
    if sse == 0:
        return float("inf")
    mse = sse / d.size
    return float("%0.2f" % (10.0 * math.log10(255.0 * 255.0 / mse)))


def _job(a):
    name, qp, i = a
    from oracle_lib import Oracle
    f = load_clip(name)[i]
    o = Oracle(qp, 3).encode_picture(*f, want_slice_data=True)
    psnr = [ffmpeg_psnr_2dp(o["rec"][c], f[c]) for c in range(3)]
    rec_sha = hashlib.sha256(b"".join(p.tobytes() for p in o["rec"])).hexdigest()
    return o["slice_data"], psnr, rec_sha


def product_lib():
    import wrenc_b200
    L = wrenc_b200.load_library()
    L.wrenc_b200_write_parameter_sets.restype = C.c_int64
    L.wrenc_b200_write_parameter_sets.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
    L.wrenc_b200_write_picture.restype = C.c_int64
    L.wrenc_b200_write_picture.argtypes = [C.c_int32, C.c_uint64, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    return L


def assemble_vvc(L, qp, slice_datas, width=W, height=H):
    """Complete byte stream: parameter sets once, then PH + slice NAL per picture (main.rs:223-260, 294-389)."""
    buf = C.create_string_buffer(4096)
    n = L.wrenc_b200_write_parameter_sets(width, height, qp, buf, len(buf))
    assert n > 0
    out = [buf.raw[:n]]
    for i, sd in enumerate(slice_datas):
        cap = len(sd) + len(sd) // 2 + 64
        b = C.create_string_buffer(cap)
        n = L.wrenc_b200_write_picture(qp, i, sd, len(sd), b, cap)
        assert n > 0
        out.append(b.raw[:n])
    return b"".join(out)


def reference_points():
    s = json.load(open(SUMMARY))
    pts = {}
    for preset in s["results"]:
        if preset["preset"] != "wrenc_fixed_qp":
            continue
        for r in preset["results"]:
            assert r["parameters"] == {"max_split_depth": 3}
            for v in r["results"]:
                clip = v["video"].split("_")[0]
                for q in v["results"]:
                    pf = q["metrics"]["PSNR"]["per_frame"]
                    pts[(clip, q["qp"])] = dict(bytes=q["bytes"], psnr=[[f["Y"], f["U"], f["V"]] for f in pf])
    return dict(commit=s["commit_id"], date=s["date"], points=pts)


def run_point(ex, L, name, qp, frames=FRAMES):
    res = list(ex.map(_job, [(name, qp, i) for i in range(frames)]))
    sds = [r[0] for r in res]
    out = dict(slice_data_bytes=[len(s) for s in sds], psnr=[r[1] for r in res], rec_sha256=[r[2] for r in res],
               slice_data_sha256=[hashlib.sha256(s).hexdigest() for s in sds])
    if frames == FRAMES:
        vvc = assemble_vvc(L, qp, sds)
        out["file_bytes"] = len(vvc)
        out["file_sha256"] = hashlib.sha256(vvc).hexdigest()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--points", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_pin.json"))
    a = ap.parse_args()
    ref = reference_points()
    want = [(c, q) for c in ("bus", "mobile") for q in QPS]
    if a.points:
        want = [(p.split(":")[0], int(p.split(":")[1])) for p in a.points.split(",")]
    L = product_lib()
    table = {}
    bad = 0
    with ProcessPoolExecutor(a.jobs) as ex:
        for clip, qp in want:
            r = run_point(ex, L, clip, qp)
            e = ref["points"][(clip, qp)]
            n_psnr = sum(1 for f in range(FRAMES) for c in range(3) if r["psnr"][f][c] == e["psnr"][f][c])
            r["reference_file_bytes"] = e["bytes"]
            r["reference_psnr"] = e["psnr"]
            r["psnr_values_equal"] = n_psnr
            table["%s:%d" % (clip, qp)] = r
            ok = r["file_bytes"] == e["bytes"] and n_psnr == 3 * FRAMES
            bad += not ok
            print("%-7s qp %2d  file bytes %8d  reference %8d  diff %+5d   per-frame PSNR equal %2d/90  %s" %
                  (clip, qp, r["file_bytes"], e["bytes"], r["file_bytes"] - e["bytes"], n_psnr, "OK" if ok else "MISMATCH"), flush=True)
    doc = dict(source="/root/reference/tools/evaluation/summary.json", reference_commit=ref["commit"], reference_date=ref["date"],
               input="tools/decode_assets.py (libavcodec H.264 decode of /root/reference/assets/*.mp4, yuv420p as is)",
               generator="tools/pin_oracle.py", points=table)
    if not a.points:
        with open(a.out, "w") as f:
            json.dump(doc, f, indent=1, sort_keys=True)
        print("wrote", a.out)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
