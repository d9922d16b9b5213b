#!/usr/bin/env python
"""bench.py -- frames/s of the per-frame pixel pipeline on B200 (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic frames: B frames (B = a ring of distinct frames larger than L2), i.e. for every
frame  [depth composite] -> text overlay -> RGB->YUV420P (scene) + GRAY8->YUV420P (depth).

  value      frames/s with the inputs already resident in HBM: one batched launch set per
             step through nes_gpu_convert_batch_device, CUDA events on the library's stream.
  e2e        the same metric through the reference-facing call with HOST buffers
             (nes_gpu_submit / nes_gpu_wait, pinned memory, 3 frames in flight): H2D and D2H
             are inside the timed region.
  roofline   algorithmic bytes per launch / average launch duration vs the measured HBM peak.
  cpu_baseline  the reference's CPU path (real libswscale + FreeType through oracle/_ref, or
             the C port) on this host, bounded sample, rank 0 at N=1 only.

--impl reference times only the CPU path (the reference has no GPU code).
Multi-GPU: one process per GPU (torchrun), sessions are independent -> weak scaling, no
data-path collective; the barrier / max-over-ranks use torch.distributed.
"""
from __future__ import annotations

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

METRIC = "frames/sec per GPU and box (1/2/4/8 B200) at 1080p/4K; p50 frame latency"
UNIT = "frames/s"
DEFAULT_WORKLOAD = "c2_1080p_2src_composite"  # BASELINE.json configs[1]
RING_TARGET_BYTES = 400 << 20  # distinct frames per step: > 3x the 126 MB L2


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 100 ms in the background."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.samples = []  # (t, sm, smmax, power, [reasons])
        self.windows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                sm, smmax, pw = float(p[0]), float(p[1]), float(p[2])
            except (ValueError, IndexError):
                continue
            reasons = [n for n, v in zip(names, p[3:7]) if v.lower().startswith("active")]
            self.samples.append((time.time(), sm, smmax, pw, reasons))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        inside = [s for s in self.samples if any(t0 - 0.05 <= s[0] <= t1 + 0.05 for t0, t1 in self.windows)]
        note = None
        if not inside:
            mid = [0.5 * (a + b) for a, b in self.windows] or [time.time()]
            inside = sorted(self.samples, key=lambda s: min(abs(s[0] - m) for m in mid))[:3]
            note = "timed regions shorter than the sampling period; nearest samples used"
        reasons = sorted({r for s in inside for r in s[4]})
        out = {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": inside[0][2], "power_w_max": max(s[3] for s in inside),
               "reasons": reasons, "samples": len(inside)}
        if note:
            out["note"] = note
        return out


# ----------------------------------------------------------------------------------------
# CPU reference arm (the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------
def cpu_reference(wl_name: str, threads: int, budget_s: float, flags: int | None = None, max_frames: int | None = None):
    """Times the reference's CPU path on this host: per frame [composite (C port; the reference
    has none)] -> 4x render_string_to_frame with FT_Load_Char per character
    (render_text.cc:81-110) -> sws_getContext + sws_scale + sws_freeContext for scene and depth
    (type_managers.cc:143-155).  flags None = 0 = the reference as shipped."""
    from oracle import oracle as O
    import ngp_encode_server_b200 as n
    wl = n.synth.WORKLOADS[wl_name]
    P = O.Port()
    R = O.Ref()
    kind = "reference" if (R.have_sws and R.have_ft) else "port"
    sws_flags = 0 if flags is None else flags
    font = os.path.join(ROOT, "tests", "golden", "Aileron-Regular.ttf")
    glyphs = None if kind == "reference" else O.GlyphTable.load(os.path.join(ROOT, "tests", "golden", "glyphs_aileron20.npz"))
    frames = [n.synth.make_sources(wl, f) for f in range(2)]
    runs = [n.synth.text_runs(wl, f) for f in range(2)]
    counts = [0] * threads
    t_first = [None] * threads
    start_evt = threading.Event()
    deadline = [0.0]

    def work(tid):
        tctx = R.text_new(font) if kind == "reference" else None
        i = 0
        start_evt.wait()
        while True:
            srcs = frames[i & 1]
            if wl["n_src"] > 1:
                comp, cdep = P.composite([s[0] for s in srcs], [s[1] for s in srcs], wl["fmt"])
            else:
                comp, cdep = srcs[0][0].copy(), srcs[0][1]
            bpp = comp.shape[2]
            for pos, txt in runs[i & 1]:  # the reference's stamp loop (RGB24 as written; 4-byte pixels: same loop, pixel stride 4)
                if kind == "reference":
                    R.text_render(tctx, comp, pos, txt) if bpp == 3 else R.text_render4(tctx, comp, pos, txt, wl["fmt"])
                else:
                    P.render_string(comp, pos, txt, glyphs) if bpp == 3 else P.render_string4(comp, pos, txt, glyphs, wl["fmt"])
            if kind == "reference":
                R.sws_convert(comp, wl["fmt"], wl["wd"], wl["hd"], flags=sws_flags)
                R.sws_convert(cdep, "gray", wl["wd"], wl["hd"], flags=sws_flags)
            else:
                P.rgb_to_yuv420p(comp, wl["fmt"], wl["wd"], wl["hd"])
                P.gray_to_yuv420p(cdep, wl["wd"], wl["hd"])
            i += 1
            counts[tid] = i
            if time.perf_counter() >= deadline[0] or (max_frames and i >= max_frames):
                break
        if tctx:
            R.text_free(tctx)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ths:
        t.start()
    t0 = time.perf_counter()
    deadline[0] = t0 + budget_s
    start_evt.set()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = sum(counts)
    return {"value": total / dt, "unit": UNIT, "cores": threads, "kind": kind, "frames": total, "seconds": dt,
            "sample": f"{total} frames of {wl_name} in {dt:.1f} s on {threads} thread(s) of {os.cpu_count()} host cpus; "
                      f"libswscale flags={'0 (reference as shipped)' if sws_flags == 0 else hex(sws_flags)}; "
                      + ("real libswscale 9.1.100 + FreeType 2.14.3 via oracle/_ref" if kind == "reference" else "C port (oracle/liboracle_port.so)")
                      + ("; composite = C port (no reference implementation exists)" if wl["n_src"] > 1 else "")}


def bind_to_gpu_numa(local: int, world: int):
    """One process per GPU: run (and first-touch its pinned buffers) on CPUs close to that GPU.
    NVML gives the GPU's ideal CPU set; ranks that share a set split it.  Returns a short note."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        ideal = None
        try:
            import pynvml
            pynvml.nvmlInit()
            words = (max(allowed) // 64) + 1
            masks = [pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), words) for i in range(world)]
            sets = [sorted(c for c in allowed if (m[c // 64] >> (c % 64)) & 1) for m in masks]
            if sets[local]:
                peers = [i for i in range(world) if sets[i] == sets[local]]
                k, n = peers.index(local), len(peers)
                share = sets[local][k * len(sets[local]) // n:(k + 1) * len(sets[local]) // n]
                ideal = share or sets[local]
        except Exception:
            ideal = None
        if ideal is None:
            ideal = allowed[local * len(allowed) // world:(local + 1) * len(allowed) // world] or allowed
        os.sched_setaffinity(0, ideal)
        return f"cpus {ideal[0]}-{ideal[-1]}"
    except (AttributeError, OSError):
        return "unbound"


def reference_threads(sessions: int) -> int:
    """The reference runs one hot-path thread per eye per session (main.cpp:274-282)."""
    return max(1, min(os.cpu_count() or 1, 2 * sessions))


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--in-flight", type=int, default=3, help="frames in flight on the e2e path (= session ring depth)")
    ap.add_argument("--warmup-seconds", type=float, default=1.0, help="minimum wall time of untimed warm-up (clock ramp)")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version
    # banner, build output) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import __graft_entry__ as g
    if rank == 0 or world == 1:
        g.build()
    import ngp_encode_server_b200 as n
    wl = dict(n.synth.WORKLOADS[args.workload])
    alg = n.synth.algorithmic_bytes(wl)
    config = {"workload": args.workload, "description": wl["desc"], "src": [wl["w"], wl["h"]], "dst": [wl["wd"], wl["hd"]], "pix_fmt": wl["fmt"],
              "sources_per_frame": wl["n_src"], "overlay": wl["text"], "algorithmic_bytes_per_frame": alg, "sessions_per_gpu": 1}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sessions = max(1, args.gpus) * (1 if wl["sessions"] == 1 else max(1, wl["sessions"] // max(1, args.gpus)))
        thr = reference_threads(sessions)
        w = cpu_reference(args.workload, thr, 1.0, max_frames=warmup)  # warm-up: W frames per thread
        per_step_s = thr / max(w["value"], 1e-9)  # a step = one frame on every thread
        steps = max(1, min(args.steps, int(60.0 / per_step_s)))
        r = cpu_reference(args.workload, thr, 1e9, max_frames=steps)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": 1000.0 * r["seconds"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": thr, "kind": r["kind"], "sample": r["sample"] + f"; step = one frame on each of {thr} threads (one hot-path thread per eye per session, main.cpp:274-282)"},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the pixel pipeline has no CPU fallback"})
        return 2
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa(local, world) if world > 1 else "unbound (single process)"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local) if rank == 0 else None  # one nvidia-smi poller per box, not per rank
    s = n.Session(device=local, max_width=max(wl["w"], wl["wd"]), max_height=max(wl["h"], wl["hd"]), max_sources=wl["n_src"], ring_depth=args.in_flight)
    metrics, bitmaps = n.synth.load_glyph_table()
    s.atlas_set(metrics, bitmaps)
    bpp = n.PIX_BPP[wl["fmt"]]
    w, h, wd, hd = wl["w"], wl["h"], wl["wd"], wl["hd"]
    in_bytes = wl["n_src"] * (bpp + 1) * w * h
    ysz, csz = n.align32(wd) * hd, n.align32(wd // 2) * (hd // 2)
    out_bytes = 2 * (ysz + 2 * csz)
    B = max(2, min(64, int(np.ceil(RING_TARGET_BYTES / (in_bytes + out_bytes)))))
    composite_resize = False  # composite + resize is one fused kernel now: every workload takes the batched path

    # ---- B distinct frames: pinned host copies + device copies
    host_frames, dev_frames, fins_dev, fouts_dev, fins_host, fouts_host, runs_made = [], [], [], [], [], [], []
    for f in range(B):
        srcs = n.synth.make_sources(wl, f)
        hs, ds, src_dev, src_host = [], [], [], []
        for px, dep in srcs:
            hp, hd_ = s.host_array(px.nbytes), s.host_array(dep.nbytes)
            hp[:] = px.reshape(-1); hd_[:] = dep.reshape(-1)
            dp, dd = s.device_alloc(px.nbytes), s.device_alloc(dep.nbytes)
            s.h2d(dp, hp); s.h2d(dd, hd_)
            hs.append((hp, hd_)); ds.append((dp, dd))
            src_dev.append(((dp, px.nbytes), (dd, dep.nbytes), 0, 0))
            src_host.append((hp, hd_, 0, 0))
        host_frames.append(hs); dev_frames.append(ds)
        fins_dev.append(n.Session.frame_in(wl["fmt"], w, h, src_dev, mem=n.NES_MEM_DEVICE))
        fins_host.append(n.Session.frame_in(wl["fmt"], w, h, src_host))
        d_s, d_d = s.device_alloc(ysz + 2 * csz), s.device_alloc(ysz + 2 * csz)
        fo = n.nes_frame_out(); fo.width, fo.height, fo.mem = wd, hd, n.NES_MEM_DEVICE
        for p, (off, ls) in enumerate([(0, n.align32(wd)), (ysz, n.align32(wd // 2)), (ysz + csz, n.align32(wd // 2))]):
            fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off, ls, d_d + off, ls
        fouts_dev.append(fo)
        sc = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s)
        dp_ = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s)
        fouts_host.append((n.api._frame_out(sc, dp_), sc, dp_))
        runs_made.append(n.Session.make_runs(n.synth.text_runs(wl, f)))
    runs_list = [n.synth.text_runs(wl, f) for f in range(B)]

    stream = torch.cuda.ExternalStream(s.stream, device=torch.device("cuda", local))

    # ---- value: HBM-resident, one batched launch set per step ---------------------------
    if composite_resize:
        def step_dev():
            ts = []
            for f in range(B):
                if len(ts) == args.in_flight:
                    s.wait(ts.pop(0))
                ts.append(s.submit_prepared(fins_dev[f], runs_made[f], fouts_dev[f]))
            for t in ts:
                s.wait(t)
    else:
        prepared = s.prepare_batch(fins_dev, runs_list, fouts_dev)

        def step_dev():
            s.run_batch(prepared, sync=False)

    # W warm-up steps, and at least ~1 s of them: an idle B200 sits at 120 MHz and needs a few
    # hundred ms of load to reach its boost clock (a 45 ms timed region right after 5 short steps
    # measured 2-4x slow, run to run)
    tw0 = time.perf_counter()
    n_warm = 0
    while n_warm < warmup or time.perf_counter() - tw0 < args.warmup_seconds:
        step_dev()
        n_warm += 1
        if n_warm % 64 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    l0 = s.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    h0 = time.perf_counter()
    for _ in range(args.steps):
        step_dev()
    host_issue_ms = 1000.0 * (time.perf_counter() - h0)
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    t1 = time.time()
    barrier()
    if sampler:
        sampler.window(t0, t1)
    dev_ms = e0.elapsed_time(e1)
    launches = s.launches - l0
    if world > 1:
        tt = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms_max = float(tt.item())
    else:
        dev_ms_max = dev_ms
    value = world * B * args.steps / (dev_ms_max / 1000.0)
    ms_per_step = dev_ms_max / args.steps

    # single-frame launches (what one streaming session without batching sees)
    single = None
    if not composite_resize:
        prep1 = [s.prepare_batch([fins_dev[f]], [runs_list[f]], [fouts_dev[f]]) for f in range(B)]
        for f in range(B):
            s.run_batch(prep1[f])
        torch.cuda.synchronize()
        reps = max(1, min(args.steps, 200))
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(reps):
            for f in range(B):
                s.run_batch(prep1[f])
        e3.record(stream)
        e3.synchronize()
        single = B * reps / (e2.elapsed_time(e3) / 1000.0)

    # ---- e2e: host pinned buffers through submit/wait, 3 frames in flight ----------------
    e2e = None
    lat_p50 = None
    lat_banded = None
    if not args.no_e2e:
        def step_host():
            ts = []
            for f in range(B):
                if len(ts) == args.in_flight:
                    s.wait(ts.pop(0))
                ts.append(s.submit_prepared(fins_host[f], runs_made[f], fouts_host[f][0]))
            for t in ts:
                s.wait(t)

        for _ in range(3):
            step_host()
        e_steps = max(3, min(args.steps, int(np.ceil(3.0 / max(1e-4, B * (in_bytes + out_bytes) / 40e9)))))
        torch.cuda.synchronize()
        barrier()
        t0 = time.time()
        p0 = time.perf_counter()
        for _ in range(e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - p0
        t1 = time.time()
        barrier()
        if sampler:
            sampler.window(t0, t1)
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": world * B * e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * in_bytes, "d2h_bytes_per_step": B * out_bytes,
               "steps": e_steps, "frames_in_flight": args.in_flight, "host_memory": "pinned (nes_gpu_host_alloc)"}
        lats = []
        for i in range(30):
            p = time.perf_counter()
            s.wait(s.submit_prepared(fins_host[i % B], runs_made[i % B], fouts_host[i % B][0]))
            lats.append(1000.0 * (time.perf_counter() - p))
        lat_p50 = statistics.median(lats)
        # the same with the low-latency mode (frame uploaded / converted / downloaded in 2 row bands)
        lat_banded = None
        if w == wd and h == hd:
            s.set_latency_bands(env_int("NES_BENCH_BANDS", 2))
            lats = []
            for i in range(30):
                p = time.perf_counter()
                s.wait(s.submit_prepared(fins_host[i % B], runs_made[i % B], fouts_host[i % B][0]))
                lats.append(1000.0 * (time.perf_counter() - p))
            lat_banded = statistics.median(lats)
            s.set_latency_bands(1)
        tm = s.last_timing()
        e2e["last_frame_us"] = {k: round(v, 1) for k, v in tm.items() if k.endswith("_us")}
        e2e["pcie_share"] = round((tm["h2d_us"] + tm["d2h_us"]) / max(tm["total_us"], 1e-9), 3)

    clocks = None
    if sampler:
        sampler.stop()
        clocks = sampler.summary()

    # ---- roofline of the dominant kernel --------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    launch_s = (dev_ms / 1000.0) / args.steps  # this rank's average launch-set duration (one dominant launch per step)
    achieved = (alg * B / launch_s) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    kernel = f"k_resize_tiles<{bpp}>" if (w != wd or h != hd) else f"k_frame_strips<{bpp}>"
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "kernel": kernel, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg * B, "launch_us": round(launch_s * 1e6, 2)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": dict(config, frames_per_step=B, l2="inputs larger than L2: ring of %d distinct frames, %.0f MB touched per step" % (B, B * (in_bytes + out_bytes) / 1e6)),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "p50_frame_latency_ms": None if lat_p50 is None else round(lat_p50, 3),
            "p50_frame_latency_ms_2_bands": None if (lat_p50 is None or lat_banded is None) else round(lat_banded, 3),
            "single_frame_launch_fps": None if single is None else round(single, 1),
            "host_issue_ms_per_step": round(host_issue_ms / args.steps, 4), "host_binding": numa_note}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s.close()
        thr = reference_threads(1)
        cb = cpu_reference(args.workload, thr, args.cpu_seconds)
        c1 = cpu_reference(args.workload, 1, max(3.0, args.cpu_seconds / 3))
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": thr, "kind": cb["kind"], "sample": cb["sample"], "per_core_value": c1["value"]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
