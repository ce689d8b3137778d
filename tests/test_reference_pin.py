"""CPU: the oracle is pinned by data the REAL reference produced.

/root/reference/tools/evaluation/summary.json records, for the reference binary at commit 1d5b5ec on its two CIF clips
(assets/*.mp4), QP 20..41 step 3, --max-split-depth 3: the size of the complete .vvc file and the per-frame PSNR-Y/U/V of the
VTM-decoded stream (two decimals).  tools/pin_oracle.py reproduces all 16 operating points with oracle/ + the product's header
writers on the exactly decoded clips (tools/decode_assets.py) and commits the outcome as tests/golden/reference_pin.json:
16/16 file sizes equal to the byte and 1440/1440 PSNR values equal.  Here:
  * the committed table is checked against summary.json itself (when /root/reference is present) and for internal consistency;
  * a subset is re-encoded live, so the table cannot go stale against the oracle: 2 frames of each clip at all 8 QPs
    (committed as tests/golden/clips/cif_clips_2frames.npz) and the complete 30-frame encodes of three operating points."""
import hashlib
import json
import os
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

import wrenc_b200

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pin_oracle  # noqa: E402

PIN = json.load(open(os.path.join(HERE, "golden", "reference_pin.json")))
QPS = pin_oracle.QPS


def _have_clips():
    if all(os.path.exists(pin_oracle.clip_path(n)) for n in ("bus", "mobile")):
        return True
    if os.path.isdir("/root/reference/assets"):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "decode_assets.py"), "--out", os.path.dirname(pin_oracle.clip_path("bus"))],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return True
    return False


def test_committed_table_is_the_reference_summary_and_all_points_match():
    pts = PIN["points"]
    assert sorted(pts) == sorted("%s:%d" % (c, q) for c in ("bus", "mobile") for q in QPS)
    if os.path.exists(pin_oracle.SUMMARY):
        ref = pin_oracle.reference_points()
        assert ref["commit"] == PIN["reference_commit"] == "1d5b5ec"
        for (clip, qp), e in ref["points"].items():
            p = pts["%s:%d" % (clip, qp)]
            assert p["reference_file_bytes"] == e["bytes"] and p["reference_psnr"] == e["psnr"]
    n_psnr = 0
    for k, p in pts.items():
        assert p["file_bytes"] == p["reference_file_bytes"], f"{k}: .vvc size {p['file_bytes']} != reference {p['reference_file_bytes']}"
        assert p["psnr"] == p["reference_psnr"], f"{k}: per-frame PSNR differs from the reference's"
        n_psnr += sum(len(f) for f in p["psnr"])
        hdr = p["file_bytes"] - sum(p["slice_data_bytes"])
        assert 600 < hdr < 2500  # parameter sets + 30 x (PH NAL + slice NAL framing) + emulation-prevention bytes
    assert n_psnr == 1440
    # spot values of BASELINE.md §1
    assert pts["bus:32"]["reference_file_bytes"] == 301521 and pts["mobile:32"]["reference_file_bytes"] == 525645


def _encode_subset(a):
    clip, i, qp = a
    from oracle_lib import Oracle
    g = np.load(os.path.join(HERE, "golden", "clips", "cif_clips_2frames.npz"))
    f = tuple(g[f"{clip}_{i}_{k}"] for k in ("y", "cb", "cr"))
    o = Oracle(qp, 3).encode_picture(*f, want_slice_data=True)
    return (clip, i, qp, hashlib.sha256(o["slice_data"]).hexdigest(), hashlib.sha256(b"".join(p.tobytes() for p in o["rec"])).hexdigest(),
            [pin_oracle.ffmpeg_psnr_2dp(o["rec"][c], f[c]) for c in range(3)])


def test_live_two_frames_of_each_clip_at_all_qps():
    jobs = [(c, i, q) for c in ("bus", "mobile") for i in (0, 1) for q in QPS]
    with ProcessPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        for clip, i, qp, sd_sha, rec_sha, psnr in ex.map(_encode_subset, jobs):
            p = PIN["points"]["%s:%d" % (clip, qp)]
            assert psnr == p["reference_psnr"][i], f"{clip} qp {qp} frame {i}: PSNR {psnr} != reference {p['reference_psnr'][i]}"
            assert sd_sha == p["slice_data_sha256"][i] and rec_sha == p["rec_sha256"][i], "committed pin table is stale against the oracle"


@pytest.mark.parametrize("point", ["bus:32", "mobile:32", "bus:20"])
def test_live_complete_encode_has_the_reference_file_size(point):
    """30 frames, oracle search + CABAC, product header writers: the .vvc is exactly as long as the reference's
    (bus QP32: 301 521 B, summary.json:1571-1574; QP20 also exercises sh_qp_delta = -6)."""
    if not _have_clips():
        pytest.skip("decoded clips absent and /root/reference/assets not available to decode them")
    clip, qp = point.split(":")
    with ProcessPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        r = pin_oracle.run_point(ex, wrenc_b200.load_library(), clip, int(qp))
    p = PIN["points"][point]
    assert r["file_bytes"] == p["reference_file_bytes"]
    assert r["psnr"] == p["reference_psnr"]
    assert r["file_sha256"] == p["file_sha256"]
