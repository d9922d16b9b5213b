#!/usr/bin/env python
"""Diagnostic: per-CTA timeline of k_frame_strips (trace build of the library).

    make -C ngp-encode-server_b200/csrc EXTRA=-DNES_TRACE OUT=$PWD/ngp-encode-server_b200/libnes_gpu_trace.so
    NES_GPU_LIB=ngp-encode-server_b200/libnes_gpu_trace.so python tools/diag_trace.py --workload c2_1080p_2src_composite --frames 16

Writes gpurun_out/trace_<workload>_<frames>.json: per CTA start / first stage full / end (ns from the
earliest start), SM id, chunks processed."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def device_batch(n, s, wl, B, distinct=None):
    """B device-resident frames of workload wl (ring of `distinct` different frames) -> (fins, runs, fouts)"""
    bpp = n.PIX_BPP[wl["fmt"]]
    w, h, wd, hd = wl["w"], wl["h"], wl["wd"], wl["hd"]
    ysz, csz = n.align32(wd) * hd, n.align32(wd // 2) * (hd // 2)
    distinct = distinct or B
    fins, fouts, runs = [], [], []
    srcs_dev = []
    for f in range(distinct):
        srcs = n.synth.make_sources(wl, f)
        sd = []
        for px, dep in srcs:
            dp, dd = s.device_alloc(px.nbytes), s.device_alloc(dep.nbytes)
            s.h2d(dp, px.reshape(-1)); s.h2d(dd, dep.reshape(-1))
            sd.append(((dp, px.nbytes), (dd, dep.nbytes), 0, 0))
        srcs_dev.append(sd)
    for f in range(B):
        fins.append(n.Session.frame_in(wl["fmt"], w, h, srcs_dev[f % distinct], mem=n.NES_MEM_DEVICE))
        d_s, d_d = s.device_alloc(ysz + 2 * csz), s.device_alloc(ysz + 2 * csz)
        fo = n.nes_frame_out(); fo.width, fo.height, fo.mem = wd, hd, n.NES_MEM_DEVICE
        for p, (off, ls) in enumerate([(0, n.align32(wd)), (ysz, n.align32(wd // 2)), (ysz + csz, n.align32(wd // 2))]):
            fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off, ls, d_d + off, ls
        fouts.append(fo)
        runs.append(n.synth.text_runs(wl, f % distinct))
    return fins, runs, fouts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2_1080p_2src_composite")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--mode", default="prepared", choices=["prepared", "convert"])
    ap.add_argument("--text", default=None, help="override the workload's overlay: none | reference | dense")
    ap.add_argument("--nsrc", type=int, default=0, help="override the number of sources")
    args = ap.parse_args()
    import ngp_encode_server_b200 as n
    import torch
    wl = dict(n.synth.WORKLOADS[args.workload])
    if args.text:
        wl["text"] = args.text
    if args.nsrc:
        wl["n_src"] = args.nsrc
    s = n.Session(device=0, max_width=max(wl["w"], wl["wd"]), max_height=max(wl["h"], wl["hd"]), max_sources=wl["n_src"])
    m, b = n.synth.load_glyph_table()
    s.atlas_set(m, b)
    if args.frames <= 0:  # the ring bench.py uses: distinct frames > 3x L2
        bpp = n.PIX_BPP[wl["fmt"]]
        per = wl["n_src"] * (bpp + 1) * wl["w"] * wl["h"] + 2 * (n.align32(wl["wd"]) * wl["hd"] + 2 * n.align32(wl["wd"] // 2) * (wl["hd"] // 2))
        args.frames = max(2, min(64, int(np.ceil((400 << 20) / per))))
    fins, runs, fouts = device_batch(n, s, wl, args.frames)
    import time
    prep = s.prepare_batch(fins, runs, fouts)
    handle = s.batch_prepare(fins, runs, fouts) if args.mode == "prepared" else None
    step = (lambda: s.batch_run(handle)) if handle else (lambda: s.run_batch(prep))
    stream = torch.cuda.ExternalStream(s.stream)
    for _ in range(200):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        step()
    host_us = 1e6 * (time.perf_counter() - t0) / args.reps
    e1.record(stream)
    torch.cuda.synchronize()
    us = 1000.0 * e0.elapsed_time(e1) / args.reps
    print(f"mode {args.mode}: host issue {host_us:.1f} us per call, device {us:.1f} us per launch", file=sys.stderr)
    run_batch_traced = step
    L = n.lib()
    out = {"workload": args.workload, "frames": args.frames, "launch_us": us, "host_issue_us": host_us, "mode": args.mode}
    if hasattr(L, "nes_debug_read_trace"):
        torch.cuda.synchronize()
        L.nes_debug_clear_trace_units()
        run_batch_traced()
        torch.cuda.synchronize()
        ub = (C.c_ulonglong * 32768)()
        L.nes_debug_read_trace_units(ub, 32768)
        u = np.array(ub[:], dtype=np.uint64).reshape(-1, 2)
        buf = (C.c_ulonglong * 4096)()
        L.nes_debug_read_trace(buf, 4096)
        a = np.array(buf[:], dtype=np.uint64).reshape(-1, 4)
        a = a[a[:, 2] > 0]
        t0 = int(a[:, 0].min())
        out["ctas"] = [[int(r[0]) - t0, int(r[1]) - t0, int(r[2]) - t0, int(r[3]) & 0xFFFFFFFF, int(r[3]) >> 32] for r in a]
        nz = np.nonzero(u[:, 0])[0]
        out["units"] = [[int(i), int(u[i, 0]) - t0, int(u[i, 1]) & 0xFFFFFFFF, int(u[i, 1]) >> 32] for i in nz]  # unit, start ns, cta, stamp chunks
        st, fi, en = a[:, 0].astype(np.int64) - t0, a[:, 1].astype(np.int64) - t0, a[:, 2].astype(np.int64) - t0
        print(f"{args.workload} x{args.frames}: launch {us:.1f} us; ctas {len(a)}; start p50 {np.median(st)/1e3:.1f} max {st.max()/1e3:.1f} us; "
              f"first-full - start p50 {np.median(fi-st)/1e3:.1f} max {(fi-st).max()/1e3:.1f} us; end min {en.min()/1e3:.1f} p50 {np.median(en)/1e3:.1f} max {en.max()/1e3:.1f} us; "
              f"mean idle tail {(en.max()-en).mean()/1e3:.1f} us", file=sys.stderr)
    else:
        print(f"{args.workload} x{args.frames}: launch {us:.1f} us (no trace in this build)", file=sys.stderr)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"trace_{args.workload}_{args.frames}.json"), "w") as f:
        json.dump(out, f)
    s.close()


if __name__ == "__main__":
    main()
