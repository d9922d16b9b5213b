#!/usr/bin/env python
"""config 4 through the mux, alone (no other measurement in the process): frames/s and mux statistics."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.2)
    ap.add_argument("--pre", type=int, default=0, help="single-frame per-call API launches on another session first")
    ap.add_argument("--pre-close", type=int, default=1)
    ap.add_argument("--full", type=int, default=0)
    ap.add_argument("--wl", default="c4_1080p_sessions")
    ap.add_argument("--measure", default="", help="run Harness.measure first: 'e2e' (with the e2e part) or 'dev' (without)")
    a = ap.parse_args()
    import torch
    import ngp_encode_server_b200 as n
    n.lib()
    args = argparse.Namespace(no_e2e=False, in_flight=3)
    H = bench.Harness(n, torch, None, 0, 1, 0, None, args)
    if a.pre:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import diag_trace
        wl = dict(n.synth.WORKLOADS["c4_1080p_sessions"])
        s = n.Session(device=0, max_width=wl["w"], max_height=wl["h"], max_sources=1)
        m, b = n.synth.load_glyph_table()
        s.atlas_set(m, b)
        fins, runs, fouts = diag_trace.device_batch(n, s, wl, 4)
        preps = [s.prepare_batch([fins[f]], [runs[f]], [fouts[f]]) for f in range(4)]
        for i in range(a.pre):
            s.run_batch(preps[i & 3])
        torch.cuda.synchronize()
        print("pre done", a.pre, file=sys.stderr)
        if a.pre_close:
            s.close()
    if a.measure:
        args2 = argparse.Namespace(no_e2e=(a.measure != "e2e"), in_flight=3, no_cpu_baseline=True, cpu_seconds=1.0, warmup_seconds=0.2)
        H2 = bench.Harness(n, torch, None, 0, 1, 0, None, args2)
        H2.measure_sessions = lambda *x, **k: {"verified": True}
        r0 = H2.measure(a.wl, 5, 3, full=bool(a.full))
        print("measure done", round(r0["value"]), file=sys.stderr)
    r = H.measure_sessions("c4_1080p_sessions", a.seconds)
    print(json.dumps({k: r[k] for k in ("value", "mux", "verified")}))


if __name__ == "__main__":
    main()
