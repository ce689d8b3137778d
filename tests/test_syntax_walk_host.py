"""CPU: the slice coder's syntax walk (wrenc_b200/csrc/syntax_walk.cuh: coding_tree / coding_unit / transform_unit /
residual_coding of a CTU as a bin string, coded-block and coded-sub-block decisions from the map of non-zero 4x4 level blocks) is
compiled for the HOST and compared, entry by entry, with the bins the oracle's syntax writer hands to its arithmetic coder
(oracle/wrenc_oracle_cabac.cpp, ctu_encoder.rs:227-2269): 14 synthetic pictures (QP 12..63, depth 0..3, edges / gradients / busy /
noise, 1-CTU to 15-CTU pictures) and, when tools/decode_assets.py has decoded them, the first frame of the reference's two clips
at two QPs.  No GPU: the same header is compiled into wrenc_b200_syntax_kernel."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_syntax_walk_matches_oracle_bins(tmp_path):
    exe = tmp_path / "syntax_walk_host_test"
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", str(exe), os.path.join(ROOT, "tests", "host", "syntax_walk_host_test.cpp"),
                           os.path.join(ROOT, "oracle", "wrenc_oracle.cpp"), os.path.join(ROOT, "oracle", "wrenc_oracle_cabac.cpp"), "-I", ROOT])
    args, n_pics = [], 14
    for clip in ("bus", "mobile"):
        path = os.path.join(ROOT, "tests", "golden", "_assets", clip + "_cif.yuv")
        if os.path.exists(path):
            for qp in (23, 35):
                args += [path, "352", "288", str(qp)]
                n_pics += 1
    out = subprocess.run([str(exe)] + args, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"{n_pics} pictures" in out.stdout and " 0 mismatches" in out.stdout
