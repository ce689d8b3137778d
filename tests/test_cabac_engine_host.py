"""CPU: the arithmetic coder of the slice coder kernel (wrenc_b200/csrc/cabac_engine.cuh: one-shift renormalisation, byte-wise
output with carry resolution, merged bypass runs, packed context words, the token program) is compiled for the HOST and
compared with the oracle's bit-by-bit engine (oracle/wrenc_oracle_cabac.cpp, bool_coder.rs:136-296) on 40 000 random bin strings
built to provoke carries through 0xff runs, and on the real bin strings of four searched pictures (QP 12 noise to QP 37).
No GPU: the same header is included by the CUDA kernel."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabac_engine_matches_oracle_engine(tmp_path):
    exe = tmp_path / "cabac_engine_host_test"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "cabac_engine_host_test.cpp"),
                           os.path.join(ROOT, "oracle", "wrenc_oracle.cpp"), os.path.join(ROOT, "oracle", "wrenc_oracle_cabac.cpp"), "-I", ROOT])
    out = subprocess.run([str(exe), "400"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "40411 strings" in out.stdout and " 0 mismatches" in out.stdout
