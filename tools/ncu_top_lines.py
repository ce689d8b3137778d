"""Dev-time: top CUDA-C lines by executed warp instructions from `ncu --page source --csv --print-source cuda,sass` output.
Usage: python tools/ncu_top_lines.py src.csv [N] [file-substring]"""
import csv, sys, collections, os
rows = csv.reader(open(sys.argv[1]))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
flt = sys.argv[3] if len(sys.argv) > 3 else ""
cur, hdr = None, None
agg = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name") or hdr is None or not r[0].strip().isdigit(): continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (os.path.basename(cur), int(r[0]))
    try:
        agg[key] += int(d.get("Instructions Executed", 0)); thr[key] += int(d.get("Thread Instructions Executed", 0)); smp[key] += int(d.get("# Samples", 0))
    except ValueError: pass
    src[key] = r[1]
tot = sum(agg.values()); ts = sum(smp.values())
print("total", tot)
for key, v in agg.most_common():
    if flt and flt not in key[0]: continue
    if N <= 0: break
    N -= 1
    print("%5.2f%% s%5.2f%% thr %4.1f %s:%d  %s" % (100 * v / tot, 100 * smp[key] / ts, thr[key] / max(1, v), key[0], key[1], src[key].strip()[:150]))
