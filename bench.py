#!/usr/bin/env python
"""bench.py -- frames/s of the per-frame pixel pipeline on B200 (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One JSON line on stdout (rank 0).  Per frame the hot path is
    [depth composite] -> text overlay -> RGB->YUV420P (scene) + GRAY8->YUV420P (depth).
A "step" is one pass over a batch of synthetic frames: R back-to-back launches of a prepared batch of B
device-resident frames (B distinct frames: a ring larger than 3x L2; R sized so that a step lasts >= 5 ms).

  value      frames/s with the inputs already resident in HBM (nes_gpu_batch_run, CUDA events on the library's
             stream; the descriptor table of the batch is prepared once -- per-frame host work is what `e2e` and
             `single_frame_api_fps` measure).
  e2e        the same metric through the reference-facing call with HOST buffers (nes_gpu_submit / nes_gpu_wait,
             pinned memory, 3 frames in flight): H2D and D2H are inside the timed region (>= 2 s of it), next to
             `pcie_ceiling` = plain cudaMemcpyAsync of the same buffers at this GPU count.
  roofline   algorithmic bytes per launch / average launch duration vs the measured HBM peak.
  cpu_baseline  the reference's CPU path (real libswscale + FreeType through oracle/_ref, or the C port) on this
             host, bounded sample, rank 0 at N=1 only.
  verified   one output frame of every measured workload is compared, outside the timed regions, with the hashes
             tests/golden/make_golden.py took from the real libswscale / FreeType (and with the C port).
  workloads  at N=1 the same set of numbers for the other BASELINE configs (4K, 7680x2160, 1080p sessions, 4x4K->1440p).

--impl reference times only the CPU path (the reference has no GPU code) and never loads libnes_gpu.so.
Multi-GPU: one process per GPU (torchrun), sessions are independent -> weak scaling, no data-path collective; the
barrier / max-over-ranks use torch.distributed.
"""
from __future__ import annotations

import argparse
import hashlib
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
EXTRA_WORKLOADS = ["4k_rgb24", "c3_7680x2160_sbs", "c4_1080p_sessions", "c5_4k_4src_to_1440p"]
RING_TARGET_BYTES = 400 << 20  # distinct frames per launch: > 3x the 126 MB L2
STEP_TARGET_MS = 5.0
E2E_TARGET_S = 2.0


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def sha16(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


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
# CPU reference arm (the only place bench.py executes oracle/ as the thing measured)
# ----------------------------------------------------------------------------------------
def cpu_reference(synth, wl_name: str, threads: int, budget_s: float, flags: int | None = None, max_frames: int | None = None):
    """Times the reference's CPU path on this host: per frame [composite (C port; the reference
    has none)] -> render_string_to_frame per overlay with FT_Load_Char per character
    (render_text.cc:81-110) -> sws_getContext + sws_scale + sws_freeContext for scene and depth
    (type_managers.cc:143-155).  flags None = 0 = the reference as shipped."""
    from oracle import oracle as O
    wl = synth.WORKLOADS[wl_name]
    P = O.Port()
    R = O.Ref()
    kind = "reference" if (R.have_sws and R.have_ft) else "port"
    sws_flags = 0 if flags is None else flags
    font = os.path.join(ROOT, "tests", "golden", "Aileron-Regular.ttf")
    glyphs = None if kind == "reference" else O.GlyphTable.load(os.path.join(ROOT, "tests", "golden", "glyphs_aileron20.npz"))
    frames = [synth.make_sources(wl, f) for f in range(2)]
    runs = [synth.text_runs(wl, f) for f in range(2)]
    counts = [0] * threads
    start_evt = threading.Event()
    deadline = [0.0]

    def work(tid):
        tctx = R.text_new(font) if kind == "reference" else None
        i = 0
        start_evt.wait()
        while True:
            if kind == "reference":
                O.expected_frame(frames[i & 1], wl["fmt"], runs[i & 1], wl["wd"], wl["hd"], P, ref=R, tctx=tctx, flags=sws_flags)
            else:
                O.expected_frame(frames[i & 1], wl["fmt"], runs[i & 1], wl["wd"], wl["hd"], P, glyphs)
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


def downstream_encode(n_frames: int = 60):
    """The stage after the hot path (send_frame_thread -> avcodec_send_frame, encode.cpp:133-165,
    type_managers.cc:47-110), timed separately as BASELINE.md asks.  The reference encodes H.264 with libx264; the
    libavcodec bundled in this image has no H.264 encoder, so a clearly labelled SUBSTITUTE (its mpeg4 encoder, the
    reference's bitrate / GOP / frame-rate defaults, main.cpp:109-123) is timed on converted 1080p planes, one thread."""
    try:
        from ngp_encode_server_b200 import avhandoff
        return avhandoff.time_substitute_encoder(1920, 1080, n_frames)
    except Exception as e:  # noqa: BLE001
        return {"status": "not measurable here: " + repr(e)[:160]}


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


def workload_config(synth, name: str) -> dict:
    """The `config` object: identical on both arms (nothing run-dependent in it)."""
    wl = synth.WORKLOADS[name]
    return {"workload": name, "description": wl["desc"], "src": [wl["w"], wl["h"]], "dst": [wl["wd"], wl["hd"]], "pix_fmt": wl["fmt"],
            "sources_per_frame": wl["n_src"], "overlay": wl["text"], "algorithmic_bytes_per_frame": synth.algorithmic_bytes(wl), "sessions_per_gpu": 1}


# ----------------------------------------------------------------------------------------
# one workload on this rank's GPU
# ----------------------------------------------------------------------------------------
class Harness:
    def __init__(self, n, torch, dist, rank, world, local, sampler, args):
        self.n, self.torch, self.dist = n, torch, dist
        self.rank, self.world, self.local = rank, world, local
        self.sampler, self.args = sampler, args
        self.golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json"))).get("configs", {})
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        self.traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def measure(self, name: str, steps: int, warmup: int, full: bool):
        """-> dict with value / roofline / e2e / verified ... for workload `name`.  full: also latency, single-frame
        and copy-ceiling figures (the main workload); the extra workloads get value, roofline, e2e, verified."""
        n, torch, args = self.n, self.torch, self.args
        wl = dict(n.synth.WORKLOADS[name])
        alg = n.synth.algorithmic_bytes(wl)
        bpp = n.PIX_BPP[wl["fmt"]]
        w, h, wd, hd = wl["w"], wl["h"], wl["wd"], wl["hd"]
        in_bytes = wl["n_src"] * (bpp + 1) * w * h
        ysz, csz = n.align32(wd) * hd, n.align32(wd // 2) * (hd // 2)
        out_bytes = 2 * (ysz + 2 * csz)
        B = max(2, min(64, int(np.ceil(RING_TARGET_BYTES / (in_bytes + out_bytes)))))
        s = n.Session(device=self.local, max_width=max(w, wd), max_height=max(h, hd), max_sources=wl["n_src"], ring_depth=args.in_flight)
        metrics, bitmaps = n.synth.load_glyph_table()
        s.atlas_set(metrics, bitmaps)
        stream = torch.cuda.ExternalStream(s.stream, device=torch.device("cuda", self.local))

        # ---- B distinct frames: pinned host copies + device copies; two sets of device outputs (consecutive launches
        # overlap on the device, they must not share destination planes)
        fins_dev, fins_host, fouts_host, runs_made, runs_list, host_srcs = [], [], [], [], [], []
        fouts_dev = [[], []]
        out_dev_ptrs = [[], []]
        for f in range(B):
            srcs = n.synth.make_sources(wl, f)
            src_dev, src_host = [], []
            for px, dep in srcs:
                hp, hd_ = s.host_array(px.nbytes), s.host_array(dep.nbytes)
                hp[:] = px.reshape(-1); hd_[:] = dep.reshape(-1)
                dp, dd = s.device_alloc(px.nbytes), s.device_alloc(dep.nbytes)
                s.h2d(dp, hp); s.h2d(dd, hd_)
                src_dev.append(((dp, px.nbytes), (dd, dep.nbytes), 0, 0))
                src_host.append((hp, hd_, 0, 0))
            host_srcs.append(src_host)
            fins_dev.append(n.Session.frame_in(wl["fmt"], w, h, src_dev, mem=n.NES_MEM_DEVICE))
            fins_host.append(n.Session.frame_in(wl["fmt"], w, h, src_host))
            for k in range(2):
                d_s, d_d = s.device_alloc(ysz + 2 * csz), s.device_alloc(ysz + 2 * csz)
                fo = n.nes_frame_out(); fo.width, fo.height, fo.mem = wd, hd, n.NES_MEM_DEVICE
                for p, (off, ls) in enumerate([(0, n.align32(wd)), (ysz, n.align32(wd // 2)), (ysz + csz, n.align32(wd // 2))]):
                    fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off, ls, d_d + off, ls
                fouts_dev[k].append(fo)
                out_dev_ptrs[k].append((d_s, d_d))
            sc = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s)
            dp_ = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s)
            fouts_host.append((n.api._frame_out(sc, dp_), sc, dp_))
            runs_list.append(n.synth.text_runs(wl, f))
            runs_made.append(n.Session.make_runs(runs_list[-1]))

        # ---- value: HBM-resident, prepared batches, R launches per step --------------------------------------
        batches = [s.batch_prepare(fins_dev, runs_list, fouts_dev[k]) for k in range(2)]
        seq = [0]

        def launch():
            s.batch_run(batches[seq[0] & 1])
            seq[0] += 1

        # warm-up: W steps and at least ~1 s (an idle B200 sits at a low clock and needs load to boost); it also
        # calibrates R (launches per step)
        for _ in range(8):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(16):
            launch()
        e1.record(stream)
        e1.synchronize()
        launch_ms_est = max(e0.elapsed_time(e1) / 16, 1e-3)
        R = max(1, int(np.ceil(STEP_TARGET_MS / launch_ms_est)))
        if self.world > 1:  # the same R on every rank
            t = torch.tensor([R], device="cuda", dtype=torch.int64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            R = int(t.item())

        def step_dev():
            for _ in range(R):
                launch()

        tw0 = time.perf_counter()
        n_warm = 0
        while n_warm < warmup or time.perf_counter() - tw0 < args.warmup_seconds:
            step_dev()
            n_warm += 1
            if n_warm % 8 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        l0 = s.launches
        t0 = time.time()
        e0.record(stream)
        h0 = time.perf_counter()
        for _ in range(steps):
            step_dev()
        host_issue_ms = 1000.0 * (time.perf_counter() - h0)
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        t1 = time.time()
        self.barrier()
        if self.sampler:
            self.sampler.window(t0, t1)
        dev_ms = e0.elapsed_time(e1)
        launches = s.launches - l0
        dev_ms_max = self.max_over_ranks(dev_ms)
        value = self.world * B * R * steps / (dev_ms_max / 1000.0)
        out = {"value": value, "unit": UNIT, "ms_per_step": dev_ms_max / steps, "steps": steps, "frames_per_step": B * R, "frames_per_launch": B,
               "launches_per_step": R, "gpu_launches": int(launches),
               "l2": "inputs larger than L2: every launch reads a ring of %d distinct frames, %.0f MB touched per launch" % (B, B * (in_bytes + out_bytes) / 1e6),
               "host_issue_ms_per_step": round(host_issue_ms / steps, 4)}

        # ---- roofline of the dominant kernel (this rank's own launches) ----------------------------------------
        launch_s = (dev_ms / 1000.0) / (steps * R)
        achieved = (alg * B / launch_s) / 1e9
        kernel = (f"k_resize_strips<{bpp}>" if (w != wd or h != hd) else f"k_frame_strips<{bpp}>")
        out["roofline"] = {"bound": "hbm", "achieved": round(achieved, 1), "peak": self.peak, "unit": "GB/s", "frac": round(achieved / self.peak, 4),
                           "traffic": self.traffic.get(name), "kernel": kernel, "peak_source": self.peak_src, "algorithmic_bytes_per_launch": alg * B,
                           "launch_us": round(launch_s * 1e6, 2),
                           "note": "launches of consecutive steps overlap on the device (programmatic dependent launch): launch_us = timed region / launches"}

        # ---- verification of what was just timed: frame 0 of the batch against the committed golden hashes ------
        ver = {"golden": "tests/golden/golden.json configs (real libswscale 9.1.100 + FreeType 2.14.3)"}
        sc0 = n.FrameManager(n.FrameContext(wd, hd, "yuv420p")); dp0 = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"))
        s.d2h(sc0.buffer, out_dev_ptrs[0][0][0]); s.d2h(dp0.buffer, out_dev_ptrs[0][0][1])
        sc1 = n.FrameManager(n.FrameContext(wd, hd, "yuv420p")); dp1 = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"))
        s.d2h(sc1.buffer, out_dev_ptrs[1][0][0]); s.d2h(dp1.buffer, out_dev_ptrs[1][0][1])
        g = self.golden.get(name)
        ver["device_scene_sha16"], ver["device_depth_sha16"] = sha16(sc0.cropped()), sha16(dp0.cropped())
        ok = g is not None and ver["device_scene_sha16"] == g["scene"] and ver["device_depth_sha16"] == g["depth"]
        ok = ok and sc1.cropped() == sc0.cropped() and dp1.cropped() == dp0.cropped()
        try:  # the C port as a second checker (test infrastructure; only compares, outside every timed region)
            from oracle import oracle as O
            P = O.Port()
            gl = O.GlyphTable.load(os.path.join(ROOT, "tests", "golden", "glyphs_aileron20.npz"))
            f_chk = B - 1
            ws, wdp = O.expected_frame(n.synth.make_sources(wl, f_chk), wl["fmt"], runs_list[f_chk], wd, hd, P, gl)
            scl = n.FrameManager(n.FrameContext(wd, hd, "yuv420p")); dpl = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"))
            s.d2h(scl.buffer, out_dev_ptrs[0][f_chk][0]); s.d2h(dpl.buffer, out_dev_ptrs[0][f_chk][1])
            ver["port_equal_frame"] = f_chk
            ver["port_equal"] = bool(scl.cropped() == ws.cropped() and dpl.cropped() == wdp.cropped())
            ok = ok and ver["port_equal"]
        except Exception as e:  # noqa: BLE001
            ver["port_equal"] = None
            ver["port_note"] = repr(e)[:120]

        # ---- single-frame launches: what one streaming session without batching sees --------------------------
        if True:
            singles = [s.batch_prepare([fins_dev[f]], [runs_list[f]], [fouts_dev[0][f]]) for f in range(B)]
            for b in singles:
                s.batch_run(b)
            torch.cuda.synchronize()
            reps = max(1, int(np.ceil((2000 if full else 300) / B)))
            e0.record(stream)
            for _ in range(reps):
                for b in singles:
                    s.batch_run(b)
            e1.record(stream)
            e1.synchronize()
            out["single_frame_launch_fps"] = round(B * reps / (e0.elapsed_time(e1) / 1000.0), 1)
            # the same through the per-call API (descriptor built, uploaded and launched per frame)
            prep1 = [s.prepare_batch([fins_dev[f]], [runs_list[f]], [fouts_dev[0][f]]) for f in range(B)]
            for p in prep1:
                s.run_batch(p)
            torch.cuda.synchronize()
            reps = max(1, int(np.ceil((1000 if full else 300) / B)))
            best = None
            for _try in range(3):  # a short loop of host calls: one scheduling hiccup of the box would be the whole reading
                e0.record(stream)
                p0 = time.perf_counter()
                for _ in range(reps):
                    for p in prep1:
                        s.run_batch(p)
                host_s = time.perf_counter() - p0
                e1.record(stream)
                e1.synchronize()
                dev_s = e0.elapsed_time(e1) / 1000.0
                if best is None or dev_s < best[0]:
                    best = (dev_s, host_s)
            out["single_frame_api_fps"] = round(B * reps / best[0], 1)
            out["single_frame_api_host_us"] = round(1e6 * best[1] / (B * reps), 2)
            for b in singles:
                s.batch_free(b)

        # ---- e2e: host pinned buffers through submit/wait, frames in flight ------------------------------------
        if not args.no_e2e:
            def step_host():
                ts = []
                for f in range(B):
                    if len(ts) == args.in_flight:
                        s.wait(ts.pop(0))
                    ts.append(s.submit_prepared(fins_host[f], runs_made[f], fouts_host[f][0]))
                for t in ts:
                    s.wait(t)

            p0 = time.perf_counter()
            step_host()
            step_host()
            est = (time.perf_counter() - p0) / 2
            target = E2E_TARGET_S if full else 0.6 * E2E_TARGET_S
            e_steps = max(3, int(np.ceil(target / max(est, 1e-4))))
            if self.world > 1:
                t = torch.tensor([e_steps], device="cuda", dtype=torch.int64)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                e_steps = int(t.item())
            torch.cuda.synchronize()
            self.barrier()
            t0 = time.time()
            p0 = time.perf_counter()
            for _ in range(e_steps):
                step_host()
            torch.cuda.synchronize()
            dt = time.perf_counter() - p0
            t1 = time.time()
            self.barrier()
            if self.sampler:
                self.sampler.window(t0, t1)
            dt = self.max_over_ranks(dt)
            e2e = {"value": self.world * B * e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * in_bytes, "d2h_bytes_per_step": B * out_bytes,
                   "steps": e_steps, "frames_per_step": B, "seconds": round(dt, 3), "frames_in_flight": args.in_flight, "host_memory": "pinned (nes_gpu_host_alloc)"}
            # the e2e path's own output of frame 0 against the same golden hashes
            ver["e2e_scene_sha16"], ver["e2e_depth_sha16"] = sha16(fouts_host[0][1].cropped()), sha16(fouts_host[0][2].cropped())
            ok = ok and g is not None and ver["e2e_scene_sha16"] == g["scene"] and ver["e2e_depth_sha16"] == g["depth"]

            # un-banded latency of one frame, then the stage times of its last frame
            lats = []
            for i in range(30):
                p = time.perf_counter()
                s.wait(s.submit_prepared(fins_host[i % B], runs_made[i % B], fouts_host[i % B][0]))
                lats.append(1000.0 * (time.perf_counter() - p))
            out["p50_frame_latency_ms"] = round(statistics.median(lats), 3)
            tm = s.last_timing()
            e2e["last_frame_us"] = {k: round(v, 1) for k, v in tm.items() if k.endswith("_us")}
            e2e["pcie_share"] = round((tm["h2d_us"] + tm["d2h_us"]) / max(tm["total_us"], 1e-9), 3)
            if full and w == wd and h == hd:
                s.set_latency_bands(env_int("NES_BENCH_BANDS", 2))
                lats = []
                for i in range(30):
                    p = time.perf_counter()
                    s.wait(s.submit_prepared(fins_host[i % B], runs_made[i % B], fouts_host[i % B][0]))
                    lats.append(1000.0 * (time.perf_counter() - p))
                out["p50_frame_latency_ms_2_bands"] = round(statistics.median(lats), 3)
                s.set_latency_bands(1)

            # ---- copy ceiling: the same pinned buffers through plain cudaMemcpyAsync (no kernels, no library), both
            # directions at once, at this GPU count -- what the host <-> device links give this rank
            ceiling = self.copy_ceiling(host_srcs, fouts_host, 1.0 if full else 0.5)
            e2e["pcie_ceiling"] = ceiling
            if ceiling.get("value"):
                e2e["frac_of_copy_ceiling"] = round(e2e["value"] / ceiling["value"], 3)
            out["e2e"] = e2e
        ver["ok"] = bool(ok)
        out["verified"] = bool(ok)
        out["verify"] = ver
        for b in batches:
            s.batch_free(b)
        s.close()
        if wl.get("sessions", 1) > 1 and not args.no_e2e:
            out["sessions"] = self.measure_sessions(name, 2.0 if full else 1.2)
            out["verified"] = bool(out["verified"] and out["sessions"]["verified"])
        return out

    def measure_sessions(self, name: str, seconds: float):
        """BASELINE config 4: `sessions` concurrent client sessions sharded s -> GPU s % n (this rank serves its shard),
        one host thread per 8 sessions, every session submits single frames from pinned host buffers through
        nes_gpu_submit / nes_gpu_wait with 2 frames in flight; the rank's mux coalesces the ready frames of all its
        sessions into shared launches.  -> aggregate frames/s (H2D + D2H inside), mux statistics."""
        import threading
        n, torch = self.n, self.torch
        wl = n.synth.WORKLOADS[name]
        total_sessions = wl["sessions"]
        mine = n.shard.sessions_of_rank(self.rank, self.world, total_sessions)
        w, h, wd, hd = wl["w"], wl["h"], wl["wd"], wl["hd"]
        bpp = n.PIX_BPP[wl["fmt"]]
        metrics, bitmaps = n.synth.load_glyph_table()
        mux = n.Mux(device=self.local, max_batch=64)
        sess = []
        base_frames = [n.synth.make_sources(wl, f) for f in range(2)]  # two distinct frames, copied per session
        for sid in mine:
            s_ = n.Session(device=self.local, max_width=max(w, wd), max_height=max(h, hd), max_sources=wl["n_src"], ring_depth=2)
            s_.atlas_set(metrics, bitmaps)
            mux.attach(s_)
            slots = []
            for f in range(2):
                src_host = []
                for px, dep in base_frames[f]:
                    hp, hd_ = s_.host_array(px.nbytes), s_.host_array(dep.nbytes)
                    hp[:] = px.reshape(-1); hd_[:] = dep.reshape(-1)
                    src_host.append((hp, hd_, 0, 0))
                fin = n.Session.frame_in(wl["fmt"], w, h, src_host)
                sc = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s_)
                dp = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s_)
                runs = n.Session.make_runs(n.synth.text_runs(wl, f))
                slots.append((fin, runs, n.api._frame_out(sc, dp), sc, dp, src_host))
            sess.append((s_, slots))
        n_threads = max(1, len(sess) // 8)
        counts = [0] * n_threads
        stop = threading.Event()
        go = threading.Event()
        errors = []

        def drive(tid):
            try:
                my = sess[tid::n_threads]
                tickets = [[None, None] for _ in my]
                go.wait()
                it = 0
                while not stop.is_set():
                    f = it & 1
                    for i, (s_, slots) in enumerate(my):
                        if tickets[i][f] is not None:
                            s_.wait(tickets[i][f])
                            counts[tid] += 1
                        tickets[i][f] = s_.submit_prepared(slots[f][0], slots[f][1], slots[f][2])
                    it += 1
                for i, (s_, slots) in enumerate(my):
                    for f in range(2):
                        if tickets[i][f] is not None:
                            s_.wait(tickets[i][f])
                            counts[tid] += 1
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))
                stop.set()

        ths = [threading.Thread(target=drive, args=(t,)) for t in range(n_threads)]
        for t in ths:
            t.start()
        torch.cuda.synchronize()
        self.barrier()
        t0 = time.time()
        p0 = time.perf_counter()
        go.set()
        # ramp: until every session has been through both of its frame slots twice (the first submits of a session allocate its
        # device buffers -- cudaMalloc synchronises the device -- which is start-up, not the steady state this measures)
        ramp_to = time.perf_counter() + 15.0
        while min(counts) < 4 * max(1, len(sess) // n_threads) and time.perf_counter() < ramp_to and not stop.is_set():
            time.sleep(0.05)
        c0, p1 = sum(counts), time.perf_counter()
        time.sleep(seconds)
        c1, p2 = sum(counts), time.perf_counter()
        stop.set()
        for t in ths:
            t.join()
        t1 = time.time()
        if self.sampler:
            self.sampler.window(t0, t1)
        fps_rank = (c1 - c0) / (p2 - p1)
        if self.world > 1:
            tt = torch.tensor([fps_rank], device="cuda", dtype=torch.float64)
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.SUM)
            fps = float(tt.item())
        else:
            fps = fps_rank
        self.barrier()
        st = mux.stats()
        # one output of the last round against the golden hash of frame 0 / 1 (frame 0's is committed)
        g = self.golden.get(name)
        s0, slots0 = sess[0]
        ok = (not errors) and g is not None and sha16(slots0[0][3].cropped()) == g["scene"] and sha16(slots0[0][4].cropped()) == g["depth"]
        in_bytes = wl["n_src"] * (bpp + 1) * w * h
        out_bytes = 2 * (n.align32(wd) * hd + 2 * n.align32(wd // 2) * (hd // 2))
        res = {"value": fps, "unit": UNIT, "sessions": total_sessions, "sessions_this_rank": len(mine), "host_threads_per_rank": n_threads, "frames_in_flight_per_session": 2,
               "seconds": round(p2 - p1, 3), "h2d_bytes_per_frame": in_bytes, "d2h_bytes_per_frame": out_bytes, "verified": bool(ok),
               "mux": {"frames": st["frames"], "launch_sets": st["launch_sets"], "launches": st["launches"], "max_batch": st["max_batch"],
                       "mean_batch": round(st["frames"] / max(st["launch_sets"], 1), 2)},
               "what": "sessions s -> GPU s % n, nes_gpu_submit / nes_gpu_wait per frame from pinned host buffers, one nes_gpu_mux per GPU"}
        if errors:
            res["errors"] = errors[:3]
        for s_, _ in sess:
            s_.close()
        mux.close()
        return res

    def copy_ceiling(self, host_srcs, fouts_host, seconds: float):
        torch = self.torch
        try:
            dev = torch.device("cuda", self.local)
            srcs = [[torch.from_numpy(a) for pair in frame for a in pair[:2]] for frame in host_srcs]
            outs = [[torch.from_numpy(sc.buffer), torch.from_numpy(dp.buffer)] for _, sc, dp in fouts_host]
            pinned = all(t.is_pinned() for fr in srcs for t in fr) and all(t.is_pinned() for fr in outs for t in fr)
            d_in = [torch.empty_like(t, device=dev) for t in srcs[0]]
            d_out = [torch.empty_like(t, device=dev) for t in outs[0]]
            s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            B = len(srcs)

            def one_pass():
                for f in range(B):
                    with torch.cuda.stream(s_in):
                        for d, t in zip(d_in, srcs[f]):
                            d.copy_(t, non_blocking=True)
                    with torch.cuda.stream(s_out):
                        for d, t in zip(d_out, outs[f]):
                            t.copy_(d, non_blocking=True)

            one_pass()
            torch.cuda.synchronize()
            p0 = time.perf_counter()
            one_pass()
            torch.cuda.synchronize()
            est = time.perf_counter() - p0
            reps = max(2, int(np.ceil(seconds / max(est, 1e-4))))
            if self.world > 1:
                t = torch.tensor([reps], device="cuda", dtype=torch.int64)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                reps = int(t.item())
            self.barrier()
            p0 = time.perf_counter()
            for _ in range(reps):
                one_pass()
            torch.cuda.synchronize()
            dt = self.max_over_ranks(time.perf_counter() - p0)
            self.barrier()
            in_b = sum(t.numel() for t in srcs[0])
            out_b = sum(t.numel() for t in outs[0])
            return {"value": self.world * B * reps / dt, "unit": UNIT, "h2d_gbs_per_gpu": round(B * reps * in_b / dt / 1e9, 2),
                    "d2h_gbs_per_gpu": round(B * reps * out_b / dt / 1e9, 2), "pinned": bool(pinned),
                    "what": "cudaMemcpyAsync of every source buffer of a frame (one stream) and of its two output images (another stream), no kernels"}
        except Exception as e:  # noqa: BLE001
            return {"value": None, "error": repr(e)[:200]}


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `workloads` block (the other BASELINE configs)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
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

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # only the checker is built and loaded here: the synthetic workloads come from the package's pure-python
        # module, libnes_gpu.so is neither built nor opened by this arm
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, stdout=sys.stderr)
        from ngp_encode_server_b200 import synth
        wl = synth.WORKLOADS[args.workload]
        sessions = max(1, args.gpus) * (1 if wl["sessions"] == 1 else max(1, wl["sessions"] // max(1, args.gpus)))
        thr = reference_threads(sessions)
        w = cpu_reference(synth, args.workload, thr, 1.0, max_frames=warmup)  # warm-up: W frames per thread
        per_step_s = thr / max(w["value"], 1e-9)  # a step = one frame on every thread
        steps = max(1, min(args.steps, int(60.0 / per_step_s)))
        r = cpu_reference(synth, args.workload, thr, 1e9, max_frames=steps)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": 1000.0 * r["seconds"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": workload_config(synth, args.workload),
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": thr, "kind": r["kind"], "sample": r["sample"] + f"; step = one frame on each of {thr} threads (one hot-path thread per eye per session, main.cpp:274-282)"},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import __graft_entry__ as g
    if rank == 0 or world == 1:
        g.build()
    import ngp_encode_server_b200 as n
    import torch
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the pixel pipeline has no CPU fallback"})
        return 2
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa(local, world) if world > 1 else "unbound (single process)"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    n.lib()
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("NES_BENCH_NO_SAMPLER")) else None  # one nvidia-smi poller per box, not per rank
    H = Harness(n, torch, dist, rank, world, local, sampler, args)
    main_res = H.measure(args.workload, args.steps, warmup, full=True)

    line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(n.synth, args.workload)}
    for k in ("frames_per_step", "frames_per_launch", "launches_per_step", "l2", "roofline", "e2e", "gpu_launches", "verified", "verify", "sessions",
              "p50_frame_latency_ms", "p50_frame_latency_ms_2_bands", "single_frame_launch_fps", "single_frame_api_fps", "single_frame_api_host_us",
              "host_issue_ms_per_step"):
        if k in main_res:
            line[k] = main_res[k]
    line["host_binding"] = numa_note
    line["composite_note"] = ("the synthetic composite sources are transparent on complementary stripes (SURVEY.md §8 d): exactly one source is "
                              "valid per pixel, the depth compare never has to break a tie in this workload (tests cover overlapping sources)")

    # ---- the other BASELINE configs (N=1 only: the driver's scaling runs stay short) -----------------------------------
    if world == 1 and not args.no_extra and args.workload == DEFAULT_WORKLOAD:
        extra = {}
        for name in EXTRA_WORKLOADS:
            try:
                r = H.measure(name, max(5, min(args.steps, 40)), 3, full=False)
                extra[name] = {k: r[k] for k in ("value", "unit", "ms_per_step", "frames_per_step", "frames_per_launch", "roofline", "e2e", "verified",
                                                 "p50_frame_latency_ms", "gpu_launches", "sessions", "single_frame_launch_fps", "single_frame_api_fps",
                                                 "single_frame_api_host_us") if k in r}
                extra[name]["config"] = workload_config(n.synth, name)
            except Exception as e:  # noqa: BLE001
                extra[name] = {"error": repr(e)[:300]}
        line["workloads"] = extra

    clocks = None
    if sampler:
        sampler.stop()
        clocks = sampler.summary()
    line["clocks"] = clocks

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        thr = reference_threads(1)
        cb = cpu_reference(n.synth, args.workload, thr, args.cpu_seconds)
        c1 = cpu_reference(n.synth, args.workload, 1, max(3.0, args.cpu_seconds / 3))
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": thr, "kind": cb["kind"], "sample": cb["sample"], "per_core_value": c1["value"]}
        if "workloads" in line:
            for name in EXTRA_WORKLOADS:
                if "error" in line["workloads"].get(name, {"error": 1}):
                    continue
                c = cpu_reference(n.synth, name, thr, 4.0)
                line["workloads"][name]["cpu_baseline"] = {"value": c["value"], "unit": UNIT, "cores": thr, "kind": c["kind"], "sample": c["sample"]}
        line["downstream_encode"] = downstream_encode()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
