"""Dev-time: time of the slice-coder stage alone (syntax + scan + compact + CABAC kernels, wrenc_b200_code_resident) for the
library selected by WRENC_B200_LIB, on the bench's 1080p content (F pictures, default 240 and 24), CUDA events on the launching
stream; frame 0's slice_data must hash to the oracle's committed result (tests/golden/bench_golden.json)."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wrenc_b200
from bench import CONFIGS, synth_host
name = os.environ.get("CFG", "1080p")
cfg = CONFIGS[name]
W, H, qp = cfg["W"], cfg["H"], cfg["qp"]
NCTU = (W // 32) * (H // 32)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_golden.json")))[name]
dev = torch.device("cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for F in [int(f) for f in os.environ.get("FS", "240 24").split()]:
    nu = min(F, cfg["unique"])
    host = synth_host(cfg, nu, 0)
    pb = W * H * 3 // 2
    d_yuv = torch.from_numpy(host).to(dev).repeat((F + nu - 1) // nu, 1)[:F].contiguous()
    d_rec = torch.empty_like(d_yuv); d_lev = torch.empty(d_yuv.shape, dtype=torch.int16, device=dev)
    d_records = torch.empty((F * NCTU, 88), dtype=torch.uint8, device=dev)
    d_out = torch.empty((F, pb), dtype=torch.uint8, device=dev); d_len = torch.empty(F, dtype=torch.int32, device=dev)
    enc = wrenc_b200.SearchEncoder(W, H, qp=qp, device=0, pictures_in_flight=1, want_recon=False, want_decisions=False, want_slice_data=True)
    enc.prepare(F)
    enc.search_resident(F, d_yuv, d_rec, d_lev, d_records, st.cuda_stream)
    enc.code_resident(F, d_lev, d_records, d_out, pb, d_len, st.cuda_stream); torch.cuda.synchronize()
    if int(d_len.min().item()) == -2:
        enc.code_resident_retry(F, d_lev, d_records, d_out, pb, d_len, st.cuda_stream); torch.cuda.synchronize()
    ms = []
    for _ in range(int(os.environ.get("REPS", 5))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); enc.code_resident(F, d_lev, d_records, d_out, pb, d_len, st.cuda_stream); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    n0 = int(d_len[0]); sd = d_out[0, :n0].cpu().numpy().tobytes()
    ok = hashlib.sha256(sd).hexdigest() == gold["slice_data_sha256"] and int(d_len.min().item()) > 0
    same = all(int(d_len[i]) == int(d_len[i % nu]) and torch.equal(d_out[i, :int(d_len[i])], d_out[i % nu, :int(d_len[i])]) for i in range(nu, min(F, 3 * nu)))
    print("%s %s F=%d coder stage %s ms (min %.2f)  frame0==oracle %s repeats identical %s  coded %d B" % (
        os.path.basename(os.environ.get("WRENC_B200_LIB", "default")), name, F, " ".join("%.2f" % m for m in ms), min(ms), ok, same, int(d_len.to(torch.int64).sum())), flush=True)
    enc.close(); del d_yuv, d_rec, d_lev, d_records, d_out, d_len; torch.cuda.empty_cache()
