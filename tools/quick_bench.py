"""Dev-time: resident-search throughput (W, H, QP, F from the environment; default 1080p QP32) for the library selected by
WRENC_B200_LIB, plus a golden parity check."""
import glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wrenc_b200
W, H, F = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1088)), int(os.environ.get("F", 48))
QP = int(os.environ.get("QP", 32))
NCTU = (W // 32) * (H // 32)
ok = True
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))[:: int(os.environ.get("GSTEP", 3))]:
    g = np.load(path)
    h, w = g["y"].shape
    enc = wrenc_b200.SearchEncoder(w, h, qp=int(g["qp"]), max_split_depth=int(g["depth"]), pictures_in_flight=1, extra_params=str(g["extra"]) or None)
    r = enc.encode_pictures([(g["y"], g["cb"], g["cr"])])[0]
    enc.close()
    same = all(np.array_equal(r["rec"][c], g["rec_" + k]) and np.array_equal(r["coef"][c], g["coef_" + k]) for c, k in enumerate(("y", "cb", "cr"))) and r["records"].tobytes() == g["records"].tobytes()
    ok &= same
nu = 6
frames = [wrenc_b200.synth_frame(W, H, seed=0xB2000002, frame=f) for f in range(nu)]
host = np.stack([np.concatenate([a.ravel() for a in f]) for f in frames])
dev = torch.device("cuda")
d_yuv = torch.from_numpy(host).to(dev).repeat((F + nu - 1) // nu, 1)[:F].contiguous()
d_rec = torch.empty_like(d_yuv); d_lev = torch.empty(d_yuv.shape, dtype=torch.int16, device=dev)
d_records = torch.empty((F * NCTU, 88), dtype=torch.uint8, device=dev)
enc = wrenc_b200.SearchEncoder(W, H, qp=QP, device=0, pictures_in_flight=1, want_recon=False, want_decisions=False)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
enc.search_resident(F, d_yuv, d_rec, d_lev, d_records, st.cuda_stream); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); enc.search_resident(F, d_yuv, d_rec, d_lev, d_records, st.cuda_stream); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("%s parity=%s  %dx%d QP%d F=%d  %.1f ms  %.1f frames/s  %.0f CTU/s" % (os.path.basename(os.environ.get("WRENC_B200_LIB", "default")), ok, W, H, QP, F, ms, F / ms * 1e3, F * NCTU / ms * 1e3), flush=True)
