"""CPU: the arithmetic coder of the slice coder kernel (wrenc_b200/csrc/cabac_engine.cuh: one-shift renormalisation, byte-wise
output with carry resolution, merged bypass runs, packed context words, the token program) is compiled for the HOST and
compared with the oracle's bit-by-bit engine (oracle/wrenc_oracle_cabac.cpp, bool_coder.rs:136-296) on 40 000 random bin strings
built to provoke carries through 0xff runs, and on the real bin strings of four searched synthetic pictures (QP 12 noise to QP 37) and of the first frame of the
reference's two clips at QP 20 / 32 / 41.
No GPU: the same header is included by the CUDA kernel."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabac_engine_matches_oracle_engine(tmp_path):
    exe = tmp_path / "cabac_engine_host_test"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "cabac_engine_host_test.cpp"),
                           os.path.join(ROOT, "oracle", "wrenc_oracle.cpp"), os.path.join(ROOT, "oracle", "wrenc_oracle_cabac.cpp"), "-I", ROOT])
    args, n = ["400"], 40411
    for clip in ("bus", "mobile"):  # the reference's clips, when tools/decode_assets.py has decoded them (build() does): real bin strings
        path = os.path.join(ROOT, "tests", "golden", "_assets", clip + "_cif.yuv")
        if os.path.exists(path):
            for qp in (20, 32, 41):
                args += [path, "352", "288", str(qp)]
                n += 1
    out = subprocess.run([str(exe)] + args, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"{n} strings" in out.stdout and " 0 mismatches" in out.stdout
