"""Small profiling workload: F frames of WxH through the host-plane path (one search-kernel launch)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wrenc_b200
W, H, F = int(os.environ.get("W", 352)), int(os.environ.get("H", 288)), int(os.environ.get("F", 8))
frames = [wrenc_b200.synth_frame(W, H, frame=f) for f in range(min(F, 4))]
frames = [frames[i % len(frames)] for i in range(F)]
enc = wrenc_b200.SearchEncoder(W, H, qp=32, pictures_in_flight=F, want_recon=False, want_decisions=False)
for rep in range(int(os.environ.get("REPS", 2))):
    t = time.time(); enc.encode_pictures(frames); dt = time.time() - t
    print("rep", rep, "%.3f s" % dt, "%.0f CTU/s" % (F * (W // 32) * (H // 32) / dt))
