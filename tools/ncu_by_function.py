"""Dev-time: aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA-C function.
Usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_by_function.py src.csv
Source lines are mapped to the enclosing function of the file on disk (lines of force-inlined helpers count for the helper)."""
import csv, re, sys, collections
ROOT = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
def func_map(path):
    starts = []
    pat = re.compile(r"^(?:extern \"C\" )?(?:static )?__(?:device|global)__.*?\b(\w+)\s*\(")
    for i, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and not line.rstrip().endswith(";"):
            starts.append((i, m.group(1)))
    return starts
def lookup(starts, ln):
    name = "(file scope)"
    for s, n in starts:
        if s <= ln: name = n
        else: break
    return name
rows = csv.reader(open(sys.argv[1]))
cur, hdr, starts = None, None, []
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur = __import__("os").environ.get("SRC_PREFIX", "") + r[1]; starts = func_map(cur) if __import__("os").path.exists(cur) else []; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name") or hdr is None or not r[0].strip().isdigit(): continue
    fn = lookup(starts, int(r[0]))
    d = dict(zip(hdr[4:], r[4:]))
    a = agg[fn]
    for k, v in d.items():
        if k in ("Instructions Executed", "Thread Instructions Executed", "# Samples") or (k.startswith("stall_") and "Not Issued" not in k):
            try: a[k] += int(v)
            except ValueError: pass
tot_i = sum(a["Instructions Executed"] for a in agg.values()); tot_s = sum(a["# Samples"] for a in agg.values())
keys = ["stall_selected", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_sleep", "stall_dispatch", "stall_mio", "stall_lg"]
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("%-22s %6s %6s %5s | " % ("function", "inst%", "smpl%", "thr") + " ".join("%7s" % k[6:13] for k in keys))
for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["Instructions Executed"]):
    if a["Instructions Executed"] < tot_i * 0.001: continue
    thr = a["Thread Instructions Executed"] / max(1, a["Instructions Executed"])
    print("%-22s %6.2f %6.2f %5.1f | " % (fn[:22], 100 * a["Instructions Executed"] / tot_i, 100 * a["# Samples"] / tot_s, thr) + " ".join("%7d" % a[k] for k in keys))
