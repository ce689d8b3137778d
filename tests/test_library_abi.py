"""CPU: the C-ABI library loads, exports every symbol include/wrenc_b200.h declares, validates arguments, derives the
search constants on the host (no GPU needed) and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import wrenc_b200
from wrenc_b200.encoder import derive_consts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "wrenc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wrenc_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = wrenc_b200.load_library()
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(wrenc_b200.EXPORTS) == names


def test_library_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {wrenc_b200.LIB_PATH} 2>/dev/null").read()
    if out.strip():
        assert "sm_100a" in out and "sm_90" not in out


def test_host_constants_match_known_answers():
    for qp, lrd, lq, ls in [(22, 35.980415, 34, 9216), (27, 79.03267, 56, 16384), (32, 173.59894, 100, 29184), (37, 381.31808, 184, 52224)]:
        c = derive_consts(qp)
        assert np.float32(c["lambda_rd"]) == np.float32(lrd) and c["lambda_q"] == lq and c["ls"] == ls
        assert c["dq"].tolist() == [0, 128, 181, 222, 257, 287, 314, 340]
        assert c["lv"].tolist() == [6548, 17546, 23774, 28619, 32720, 36338, 39610, 42618]


def test_host_constants_match_the_oracle_tables():
    import ctypes as C
    from oracle_lib import Oracle
    for qp, extra in [(32, None), (26, None), (30, "lambda_mul_dq_trellis=1.5,quant_lambda_offset_trellis=7,cclm_pow=0.5,a=0.9,mpm_idx_pow=0.3")]:
        c = derive_consts(qp, extra)
        o = Oracle(qp, 3, extra)
        oc = o.consts()
        assert oc["lambda_q"] == c["lambda_q"] and oc["ls"] == c["ls"] and np.float32(oc["lambda_rd"]) == np.float32(c["lambda_rd"])
        hs, hd, hc = np.zeros((67, 4), np.int64), np.zeros(67, np.int64), np.zeros(4, np.int64)
        o.L.wo_hdr_tables(C.c_void_p(o.h), hs.ctypes.data_as(C.c_void_p), hd.ctypes.data_as(C.c_void_p), hc.ctypes.data_as(C.c_void_p))
        assert np.array_equal(hs, c["hdr_single"]) and np.array_equal(hd, c["hdr_dual"]) and np.array_equal(hc, c["hdr_chroma"])


def test_extra_params_errors_mirror_the_reference():
    with pytest.raises(ValueError):
        derive_consts(32, "lambda_mul_dq_trellis")  # main.rs:206-214: "Invalid extra-params"
    with pytest.raises(ValueError):
        derive_consts(32, "a=1=2")
    derive_consts(32, "some_unknown_key=1")  # stored, never read


def test_invalid_config_is_rejected_before_touching_cuda():
    for kw in (dict(width=100, height=64), dict(width=64, height=0), dict(width=64, height=64, qp=64), dict(width=64, height=64, max_split_depth=4)):
        args = dict(width=64, height=64, qp=32, max_split_depth=3)
        args.update(kw)
        with pytest.raises(ValueError):
            wrenc_b200.SearchEncoder(**args)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the failure path is covered on the CPU box")
    with pytest.raises(wrenc_b200.WrencB200Error):
        wrenc_b200.SearchEncoder(64, 64)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wrenc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "wrenc_oracle" not in txt and "oracle/" not in txt.replace("touches oracle/", ""), f
