"""Dev-time: where the search kernel's cycles go, per node size and phase (needs a -DWB_PROFILE build selected by
WRENC_B200_LIB; see the WB_PROF marks in wrenc_b200/csrc/search.cu).  Same workload as quick_bench.py."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wrenc_b200
W, H, F = 1920, 1088, int(os.environ.get("F", 240))
nu = 6
frames = [wrenc_b200.synth_frame(W, H, seed=0xB2000002, frame=f) for f in range(nu)]
host = np.stack([np.concatenate([a.ravel() for a in f]) for f in frames])
dev = torch.device("cuda")
d_yuv = torch.from_numpy(host).to(dev).repeat((F + nu - 1) // nu, 1)[:F].contiguous()
d_rec = torch.empty_like(d_yuv); d_lev = torch.empty(d_yuv.shape, dtype=torch.int16, device=dev)
d_records = torch.empty((F * 2040, 88), dtype=torch.uint8, device=dev)
enc = wrenc_b200.SearchEncoder(W, H, qp=32, device=0, pictures_in_flight=1, want_recon=False, want_decisions=False)
lib = ctypes.CDLL(os.environ["WRENC_B200_LIB"])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
buf = (ctypes.c_ulonglong * 128)()
enc.search_resident(F, d_yuv, d_rec, d_lev, d_records, st.cuda_stream); torch.cuda.synchronize()
lib.wrenc_b200_debug_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); enc.search_resident(F, d_yuv, d_rec, d_lev, d_records, st.cuda_stream); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
lib.wrenc_b200_debug_prof(buf, 0)
v = np.array(list(buf), dtype=np.float64)
tot = v.sum()
print("F=%d  %.1f ms  %.0f CTU/s   sum of CTA cycles %.3e (= %.1f ms x 148 CTAs at 1.965 GHz)" % (F, ms, F * 2040 / ms * 1e3, tot, tot / 148 / 1.965e6))
leaf = ["setup", "refs", "ph1 pl/dc full+13 SAD", "dec1", "ph2,3 SAD", "dec2,3", "ph4 3 full", "dec4", "ph5 commit", "ph5b ds", "ph6 cclm SAD", "dec6", "ph7 cclm full", "dec7", "ph8 commit", "-"]
cct = ["setup", "refs+ds", "DM full+cclm SAD", "dec", "cclm full", "dec", "commit"] + ["-"] * 9
misc = ["ticket", "wavefront wait", "staging", "search tail", "write-back"] + ["-"] * 11
for kind, name in enumerate(["32x32", "16x16", "8x8", "4x4 luma", "chroma CT", "outside"]):
    blk = v[16 * kind: 16 * kind + 16]
    print("%-10s %5.1f %%" % (name, 100 * blk.sum() / tot))
    names = leaf if kind < 4 else (cct if kind == 4 else misc)
    for i in range(16):
        if blk[i] > 0:
            print("    %-24s %5.2f %%" % (names[i], 100 * blk[i] / tot))
