"""Generates tests/golden/*.npz from the CPU oracle (oracle/, the C++ restatement of the reference — the Rust reference
cannot be built or imported in this environment).  The fixtures pin the oracle against drift and give the
GPU tests reference outputs that need no CPU search at run time.   Usage: python tests/golden/gen_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle  # noqa: E402
from wrenc_b200.synth import random_frame, synth_frame  # noqa: E402

CASES = [  # name, W, H, qp, depth, kind, seed/frame, extra
    ("synth_96x64_qp32_d3", 96, 64, 32, 3, "synth", 0, None),
    ("synth_96x64_qp22_d3", 96, 64, 22, 3, "synth", 1, None),
    ("synth_96x64_qp27_d3", 96, 64, 27, 3, "synth", 2, None),
    ("synth_96x64_qp37_d3", 96, 64, 37, 3, "synth", 3, None),
    ("synth_64x96_qp32_d2", 64, 96, 32, 2, "synth", 4, None),
    ("synth_64x64_qp32_d1", 64, 64, 32, 1, "synth", 5, None),
    ("synth_64x64_qp32_d0", 64, 64, 32, 0, "synth", 6, None),
    ("random_64x64_qp32_d3", 64, 64, 32, 3, "random", 7, None),
    ("random_32x32_qp26_d3", 32, 32, 26, 3, "random", 8, None),
    ("synth_96x64_qp30_extra", 96, 64, 30, 3, "synth", 9, "lambda_mul_dq_trellis=1.5,quant_lambda_offset_trellis=7,cclm_pow=0.5"),
    ("flat_64x64_qp32_d3", 64, 64, 32, 3, "flat", 10, None),
]


def make_frame(kind, W, H, s):
    if kind == "synth":
        return synth_frame(W, H, frame=s)
    if kind == "random":
        return random_frame(W, H, s)
    return (np.full((H, W), 77, np.uint8), np.full((H // 2, W // 2), 130, np.uint8), np.full((H // 2, W // 2), 120, np.uint8))


def main():
    for name, W, H, qp, depth, kind, s, extra in CASES:
        y, cb, cr = make_frame(kind, W, H, s)
        o = Oracle(qp, depth, extra).encode_picture(y, cb, cr, want_slice_data=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), y=y, cb=cb, cr=cr, rec_y=o["rec"][0], rec_cb=o["rec"][1], rec_cr=o["rec"][2],
                            coef_y=o["coef"][0], coef_cb=o["coef"][1], coef_cr=o["coef"][2], records=o["records"].view(np.uint8), slice_data=np.frombuffer(o["slice_data"], np.uint8),
                            qp=qp, depth=depth, extra=extra or "")
        print(name, "ok", sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.startswith(name)))


if __name__ == "__main__":
    main()
