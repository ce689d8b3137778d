"""CPU: the four-samples-per-lane angular predictor the search kernel uses (wrenc_b200/csrc/ang4.cuh: byte reference lines,
funnel-shifted windows, packed dot products, PDPC) is compiled for the HOST and compared with the oracle's prediction
(oracle/wrenc_oracle.cpp `predict`, intra_predictor.rs:1287-1602) for all 65 angular modes x block sizes 4..32 x components x
neighbour-availability patterns x picture positions.  No GPU: the same header is included by the CUDA kernel."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ang4_matches_oracle_prediction(tmp_path):
    exe = tmp_path / "ang4_host_test"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "ang4_host_test.cpp"),
                           os.path.join(ROOT, "oracle", "wrenc_oracle.cpp"), "-I", ROOT])
    out = subprocess.run([str(exe), "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout and "78000 blocks" in out.stdout
