#!/usr/bin/env python
"""bench.py — all-intra search throughput on B200 (BASELINE.json metric: 1080p all-intra frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]

A step = one pass of the hot path (the whole RD search of every CTU, then the CABAC coding of every picture) over F synthetic 1920x1088 I420 frames per GPU at
QP32, --max-split-depth 3 (BASELINE.json configs[2]).  Frames are independent IDR pictures, so ranks shard picture
ranges and there is no data-path collective ("scaling": "weak": every rank searches its own F frames per step).

  value        frames/s of the whole hot path (search kernel + syntax/CABAC kernels), inputs resident in HBM, timed with
               CUDA events on the launching stream, max over ranks
  e2e          same metric through the C-ABI submit/receive calls with HOST planes: H2D of every frame and D2H of the
               per-CTU records + quantised levels inside the timed region
  roofline     INT32 issue roofline of the search kernel: 8 290 304 nominal integer ops per CTU
               against the IMAD rate measured live on this GPU (2 ops per multiply-add); roofline_hbm shows why HBM is
               not the bound
  cpu_baseline the CPU oracle (oracle/, a C++ restatement of the reference: the Rust reference cannot be built here)
               on all host cores, one process per core, on a bounded sample (kind "port")
  --impl reference   the same oracle arm as a stand-alone run (rank 0 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, QP, DEPTH = 1920, 1088, 32, 3
CTUS_PER_FRAME = (W // 32) * (H // 32)
OPS_PER_CTU = 8290304          # SURVEY.md §8(d) nominal integer ops per CTU (transforms 4 358 144 + trellis 3 932 160)
ALG_BYTES_PER_CTU = 1536 + 1536 + 3072 + 88   # source read + recon write + level write + record
NCU_DRAM_BYTES_PER_CTU = 16990  # dram__bytes_read.sum + dram__bytes_write.sum per CTU, ncu capture of the final build (profiles/README.md)
METRIC = "1080p all-intra frames/s (RD search + CABAC slice_data, byte-identical vs oracle)"
UNIT = "frames/s"


def synth_frames(n_unique, seed=0xB2000002):
    from wrenc_b200.synth import synth_frame
    return [synth_frame(W, H, seed=seed, frame=f) for f in range(n_unique)]


# ------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle): one process per core, each searching `rows` CTU rows of a 1080p frame
# ------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    idx, rows = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from wrenc_b200.synth import synth_frame
    y, cb, cr = synth_frame(W, H, seed=0xB2000002, frame=idx)
    hh = rows * 32
    o = Oracle(QP, DEPTH)
    t0 = time.perf_counter()
    o.encode_picture(y[:hh], cb[:hh // 2], cr[:hh // 2], want_slice_data=True)  # search + syntax/CABAC, like the GPU arm
    return time.perf_counter() - t0


def cpu_arm_step(pool, cores, rows):
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(i, rows) for i in range(cores)])
    dt = time.perf_counter() - t0
    ctus = cores * rows * (W // 32)
    return (ctus / CTUS_PER_FRAME) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    cores = os.cpu_count() or 1
    rows = 4
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_arm_step(pool, cores, 1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_arm_step(pool, cores, rows)
        dt = time.perf_counter() - t0
    frames = args.steps * cores * rows * (W // 32) / CTUS_PER_FRAME
    v = frames / dt
    sample = f"{cores} processes x {rows} CTU rows (1920x{rows * 32}) of a 1920x1088 QP32 frame per step; frames = CTUs/2040"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f32-cost",
            "data": "synthetic", "config": workload_config(args, args.frames),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, frames):
    return {"workload": "synthetic 1920x1088 yuv420p 8-bit all-intra QP32 max-split-depth 3 (BASELINE.json configs[2])",
            "frames_per_gpu_per_step": frames, "ctus_per_frame": CTUS_PER_FRAME, "qp": QP, "max_split_depth": DEPTH,
            "l2": "inputs larger than L2 (frames_per_gpu_per_step x 3.1 MB source + 9.4 MB outputs per frame)",
            "parallelism": f"picture ranges sharded over {args.gpus} GPU(s), CTU wavefronts of all pictures interleaved per GPU"}


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(p[0])); mx = max(mx, float(p[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import wrenc_b200
    from wrenc_b200.encoder import measure_int32_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    F = args.frames
    n_unique = min(F, args.unique)
    frames = synth_frames(n_unique, seed=0xB2000002 + 977 * rank)
    pic_bytes = W * H * 3 // 2
    host = np.empty((n_unique, pic_bytes), np.uint8)
    for i, (y, cb, cr) in enumerate(frames):
        host[i] = np.concatenate([y.ravel(), cb.ravel(), cr.ravel()])
    reps = (F + n_unique - 1) // n_unique
    d_yuv = torch.from_numpy(host).to(dev).repeat(reps, 1)[:F].contiguous()
    d_rec = torch.empty((F, pic_bytes), dtype=torch.uint8, device=dev)
    d_lev = torch.empty((F, pic_bytes), dtype=torch.int16, device=dev)
    d_records = torch.empty((F * CTUS_PER_FRAME, 88), dtype=torch.uint8, device=dev)

    enc = wrenc_b200.SearchEncoder(W, H, qp=QP, max_split_depth=DEPTH, device=local, pictures_in_flight=args.e2e_batch,
                                   want_recon=False, want_decisions=False, want_slice_data=True)
    out_cap = pic_bytes
    d_out = torch.empty((F, out_cap), dtype=torch.uint8, device=dev)
    d_out_len = torch.empty(F, dtype=torch.int32, device=dev)
    tstream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: its handle goes through the C ABI
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    search_ev = []

    def step():  # the whole hot path: RD search of every CTU, then syntax + CABAC coding of every picture
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = enc.search_resident(F, d_yuv, d_rec, d_lev, d_records, stream)
        e1.record()
        search_ev.append((e0, e1))
        return n + enc.code_resident(F, d_lev, d_records, d_out, out_cap, d_out_len, stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    ev[0].record()
    for k in range(args.steps):
        launches += step()
        ev[k + 1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [a.elapsed_time(b) for a, b in search_ev[-args.steps:]]  # the search kernel alone (dominant kernel)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * F * args.steps / (elapsed_ms * 1e-3)

    # ---- parity spot check of the timed outputs (size-independent property: deterministic, equal for repeated frames)
    if F > n_unique:
        a = d_rec[0]
        b = d_rec[n_unique]
        assert torch.equal(a, b), "repeated input frame produced a different reconstruction"
        la, lb = int(d_out_len[0]), int(d_out_len[n_unique])
        assert la == lb and la > 0 and torch.equal(d_out[0, :la], d_out[n_unique, :lb]), "repeated input frame produced different slice_data"
    coded_bytes = int(d_out_len.to(torch.int64).sum().item())
    assert int(d_out_len.min().item()) > 0, "slice_data coder reported an overflow"

    # ---- e2e: host planes through submit/receive
    # the step's inputs come from pinned host memory (the buffer a YUV reader would fill): submit_pinned, no staging copy
    host_pinned_t = torch.from_numpy(host).pin_memory()
    hp = host_pinned_t.numpy()
    planes = [(hp[i, :W * H].reshape(H, W), hp[i, W * H:W * H * 5 // 4].reshape(H // 2, W // 2), hp[i, W * H * 5 // 4:].reshape(H // 2, W // 2))
              for i in range(n_unique)]
    Fe = args.e2e_frames

    def e2e_step():
        got, cost, coded = 0, 0.0, []
        i = 0
        while got < Fe:
            nb = min(args.e2e_batch, Fe - i) if i < Fe else 0
            for _ in range(nb):
                y, cb, cr = planes[i % n_unique]
                enc.submit(i, y, cb, cr, pinned=True)
                i += 1
            while enc.pending():
                r = enc.receive(copy=False)
                cost += float(r["records"]["cost"][0]) + len(r["slice_data"])
                coded.append(r["slice_data"])
                got += 1
        if dist is not None:  # the job's only exchange: ordered gather of the per-picture byte buffers on the writer rank
            from wrenc_b200.sharding import gather_in_order
            allb = gather_in_order(coded, dst=0, device=dev)
            if rank == 0:
                assert len(allb) == world * Fe
        return cost

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Fe * args.e2e_steps / float(te.item())

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline (rank 0's GPU): INT32 issue peak measured live
    imad = measure_int32_peak(local)
    peak_ops = 2.0 * imad
    launch_ms = statistics.mean(kernel_ms)
    ctus = F * CTUS_PER_FRAME
    achieved_ops = OPS_PER_CTU * ctus / (launch_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = ALG_BYTES_PER_CTU * ctus / (launch_ms * 1e-3) / 1e9
    roof = {"bound": "int32", "achieved": achieved_ops / 1e12, "peak": peak_ops / 1e12, "unit": "Tops/s", "frac": achieved_ops / peak_ops,
            "traffic": NCU_DRAM_BYTES_PER_CTU * ctus, "traffic_note": "dram__bytes_read+write of one ncu capture of this kernel (130 560-CTU launch, profiles/README.md) scaled per CTU to this launch",
            "kernel": "wrenc_b200_search_kernel", "launch_ms": launch_ms, "units_per_launch": ctus, "ops_per_unit": OPS_PER_CTU, "share_of_step": launch_ms / (elapsed_ms / args.steps),
            "peak_source": "IMAD-chain microbenchmark run live in bench.py (2 ops per multiply-add); not in MEASURED_PEAKS.json"}
    roof_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "bytes_per_unit": ALG_BYTES_PER_CTU, "ncu_dram_bytes_per_unit": NCU_DRAM_BYTES_PER_CTU, "ncu_note": "dram__bytes_read+write of one ncu capture (130 560-CTU launch, profiles/README.md): the algorithmic bytes plus what still spills of the per-CTA scratch (candidate slots, saved states; kept in L2 by a persisting access window)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}

    # ---- CPU baseline on a bounded sample
    cpu = None
    if not args.no_cpu:
        import multiprocessing as mp
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        cores = os.cpu_count() or 1
        rows = args.cpu_rows
        with mp.get_context("fork").Pool(cores) as pool:
            v, dt = cpu_arm_step(pool, cores, rows)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} processes x {rows} CTU rows (1920x{rows * 32}) of a 1920x1088 QP32 frame, {dt:.1f} s wall; frames = CTUs/2040"}

    h2d = Fe * pic_bytes
    d2h = Fe * (CTUS_PER_FRAME * 88) + int(coded_bytes / F * Fe)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/f32-cost", "data": f"synthetic ({n_unique} unique frames per GPU, repeated to {F})",
            "config": workload_config(args, F), "ctus_per_s": value * CTUS_PER_FRAME, "coded_bytes_per_frame": coded_bytes / F,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "frames_per_step": Fe,
                    "steps": args.e2e_steps, "batch": args.e2e_batch, "note": "pinned host planes -> submit_pinned/receive (H2D of every frame in the timed region); D2H = CABAC-coded slice_data of every picture + CTU records"},
            "gpu_launches": launches, "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=240, help="frames per GPU per step")
    ap.add_argument("--unique", type=int, default=12, help="unique synthetic frames generated on the host (repeated on device)")
    ap.add_argument("--e2e-frames", type=int, default=240)
    ap.add_argument("--e2e-batch", type=int, default=240)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-rows", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
