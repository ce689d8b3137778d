#!/usr/bin/env python3
"""Golden hashes for bench.py's parity check of its TIMED outputs: frame 0 of rank 0 of every bench configuration, encoded
by the CPU oracle (search + CABAC) -> sha256 of slice_data and of the I420 reconstruction.  Writes tests/golden/bench_golden.json.
Test infrastructure (the oracle is pinned against the reference's own output files: tests/test_reference_pin.py)."""
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def job(name):
    import bench
    from oracle_lib import Oracle
    from wrenc_b200.synth import synth_frame
    cfg = bench.CONFIGS[name]
    f = synth_frame(cfg["W"], cfg["H"], seed=cfg["seed"], frame=0)
    o = Oracle(cfg["qp"], bench.DEPTH).encode_picture(*f, want_slice_data=True)
    return name, dict(width=cfg["W"], height=cfg["H"], qp=cfg["qp"], seed=cfg["seed"], frame=0, slice_data_bytes=len(o["slice_data"]),
                      slice_data_sha256=hashlib.sha256(o["slice_data"]).hexdigest(),
                      rec_sha256=hashlib.sha256(b"".join(p.tobytes() for p in o["rec"])).hexdigest())


if __name__ == "__main__":
    import bench
    with ProcessPoolExecutor(4) as ex:
        out = dict(ex.map(job, sorted(bench.CONFIGS)))
    path = os.path.join(ROOT, "tests", "golden", "bench_golden.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
