"""Dev-time GPU diagnosis: block-op parity, then full-search parity with a report of the first differing CTU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import wrenc_b200
from oracle_lib import Oracle

def main():
    W, H, qp = int(os.environ.get("W", 96)), int(os.environ.get("H", 64)), int(os.environ.get("QP", 32))
    md = int(os.environ.get("MD", 3))
    rng = np.random.default_rng(1)
    ora = Oracle(qp, md)
    enc = wrenc_b200.SearchEncoder(W, H, qp=qp, max_split_depth=md, pictures_in_flight=4)
    print("consts", enc.consts()["lambda_q"], ora.consts()["lambda_q"])
    ok = True
    for l2 in (2, 3, 4, 5):
        n = 1 << l2
        res = rng.integers(-255, 256, (16, n, n)).astype(np.int16)
        res[0] = 0; res[1] = 255; res[2] = -255
        g = enc.block_fwd_dct(res)
        o = np.stack([ora.fwd_dct(b) for b in res])
        print("fwd", n, np.array_equal(g, o)); ok &= np.array_equal(g, o)
        coefs = o.copy()
        coefs[3] = (rng.normal(0, 3, (n, n))).astype(np.int16)
        coefs[4] = (rng.normal(0, 30, (n, n))).astype(np.int16)
        coefs[5] = 0; coefs[5, 0, 0] = 1
        coefs[6] = 0; coefs[6, 0, 0] = -37
        coefs[7] = (rng.normal(0, 300, (n, n)) * (rng.random((n, n)) < 0.1)).astype(np.int16)
        gq, gr = enc.block_quantize(coefs)
        oq = np.stack([ora.quantize(b) for b in coefs])
        orr = np.array([ora.rate(b) for b in oq])
        eq = np.array_equal(gq, oq)
        print("quant", n, eq, "rate", np.array_equal(gr, orr)); ok &= eq
        if not eq:
            bad = [i for i in range(16) if not np.array_equal(gq[i], oq[i])]
            print("  bad blocks", bad)
            i = bad[0]; d = np.argwhere(gq[i] != oq[i]); print("  first diffs", d[:5], gq[i][tuple(d[0])], oq[i][tuple(d[0])])
        gd = enc.block_dequantize(oq)
        od = np.stack([ora.dequantize(b) for b in oq])
        print("deq", n, np.array_equal(gd, od)); ok &= np.array_equal(gd, od)
        gi = enc.block_inv_dct(od)
        oi = np.stack([ora.inv_dct(b) for b in od])
        print("inv", n, np.array_equal(gi, oi)); ok &= np.array_equal(gi, oi)
    # prediction
    y, cb, cr = wrenc_b200.random_frame(W, H, 5)
    rec = [y, cb, cr]
    nbad = 0
    for (x0, y0, w) in [(32, 32, 32), (32, 32, 16), (48, 32, 16), (32, 48, 16), (40, 40, 8), (36, 36, 4), (0, 0, 32), (0, 32, 16), (64, 0, 8), (64, 32, 32), (88, 56, 8)]:
        if x0 + w > W or y0 + w > H: continue
        for ar in (0, 1):
            for bl in (0, 1):
                for c in ((0, 1, 2) if w >= 8 else (0,)):
                    modes = list(range(67)) + ([81, 82, 83] if c else [])
                    for m in modes:
                        tree = 0 if w >= 8 else 1
                        g = enc.block_predict(rec, x0, y0, w, tree, ar, bl, c, m)
                        o = ora.predict(rec, x0, y0, w, tree, ar, bl, c, m)
                        if not np.array_equal(g, o):
                            nbad += 1
                            if nbad <= 12: print("pred mismatch", (x0, y0, w), "ar", ar, "bl", bl, "c", c, "mode", m, np.argwhere(g != o)[:3].tolist())
    print("pred mismatches", nbad); ok &= nbad == 0
    # full search
    for kind in ("synth", "random"):
        frames = [wrenc_b200.synth_frame(W, H, frame=f) if kind == "synth" else wrenc_b200.random_frame(W, H, 100 + f) for f in range(2)]
        t = time.time(); res = enc.encode_pictures(frames); dt = time.time() - t
        for f, ((yy, cbb, crr), r) in enumerate(zip(frames, res)):
            o = ora.encode_picture(yy, cbb, crr, want_slice_data=True)
            sdsame = o['slice_data'] == r['slice_data']
            print('   slice_data', 'same' if sdsame else 'DIFF', len(o['slice_data']), len(r['slice_data']))
            if not sdsame:
                a, b = o['slice_data'], r['slice_data']
                k = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))
                print('   first differing byte', k)
            ok &= sdsame
            same = all(np.array_equal(o["rec"][c], r["rec"][c]) for c in range(3)) and all(np.array_equal(o["coef"][c], r["coef"][c]) for c in range(3)) and o["records"].tobytes() == r["records"].tobytes()
            print(kind, "frame", f, "bit-exact" if same else "MISMATCH", "gpu %.3fs" % dt)
            ok &= same
            if not same:
                Wc = W // 32
                for i in range(len(o["records"])):
                    a, b = o["records"][i], r["records"][i]
                    cx, cy = (i % Wc) * 32, (i // Wc) * 32
                    recsame = all(np.array_equal(o["rec"][c][cy >> (c > 0):(cy + 32) >> (c > 0), cx >> (c > 0):(cx + 32) >> (c > 0)], r["rec"][c][cy >> (c > 0):(cy + 32) >> (c > 0), cx >> (c > 0):(cx + 32) >> (c > 0)]) for c in range(3))
                    if a.tobytes() != b.tobytes() or not recsame:
                        print(" first bad CTU", i, (cx, cy), "split", hex(a["split_mask"]), hex(b["split_mask"]), "cost", a["cost"], b["cost"], "recsame", recsame)
                        print("  luma o", a["luma_mode"].reshape(8, 8).tolist()); print("  luma g", b["luma_mode"].reshape(8, 8).tolist())
                        print("  chroma o", a["chroma_mode"].tolist()); print("  chroma g", b["chroma_mode"].tolist())
                        break
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
