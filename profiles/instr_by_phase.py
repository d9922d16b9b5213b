#!/usr/bin/env python
"""Warp instructions and stall samples per PHASE of a kernel: joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv)
with nvdisasm -gi line info of the cubin (innermost line + the line it was inlined at) and groups by line ranges of the kernel file.
usage: instr_by_phase.py <source_page.csv> <cubin> <kernel-mangled-substring> <frame_strips|resize_strips> <pixels per launch>"""
import csv, re, subprocess, sys
from collections import defaultdict

page, cubin, kname, which, pixels = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
SRC = {"frame_strips": "frame_strips.cu", "resize_strips": "resize_strips.cu"}[which]


def anchors(path):
    """line numbers of the comment anchors that delimit the phases (so the table survives edits)"""
    out = {}
    for i, l in enumerate(open(path), 1):
        out.setdefault(l.strip()[:60], i)
    return out


import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = open(os.path.join(ROOT, "ngp-encode-server_b200", "csrc", SRC)).read().split("\n")


def find(sub, start=0):
    for i in range(start, len(lines)):
        if sub in lines[i]:
            return i + 1
    raise KeyError(sub)


if which == "resize_strips":
    P = find("=== producer warp ==="); V = find("=== vertical warps ==="); H = find("=== horizontal warps ===")
    vl = find("if (it < total_l) {", V); vc = find("// chroma (U and V share the filter)", V)
    hov = find("unit_overlay(jobs[c.job]", H); hrow = find("---- per source row: composite", H); hh = find("---- horizontal pass of this row", H)
    hconv = find("const bool need_l = y >= c.lr0", H)
    END = find("// host side: planning and launch")
    sel = (find("__device__ __forceinline__ void select4") - 1, find("// Horizontal pass of one source row for NL slots") - 1)
    hl = (find("__device__ __forceinline__ void h_luma") - 1, find("__device__ __forceinline__ void h_chroma") - 2)
    hc = (find("__device__ __forceinline__ void h_chroma") - 1, find("// barrier among the horizontal warps only") - 1)
    ov = (find("__device__ __forceinline__ void unit_overlay"), find("}  // namespace", find("__device__ __forceinline__ void unit_overlay")))
    RANGES = [(P, V - 1, "producer warp (TMA issue, chunk contexts)"), (V, vl - 1, "V: chunk set-up"), (vl, vc - 1, "V: luma + depth vertical pass"), (vc, H - 1, "V: chroma vertical pass"),
              (H, hov - 1, "H: chunk set-up (waits, filter registers)"), (hov, hov, "H: unit overlay build"), (hov + 1, hconv - 1, "H: row load, composite select, overlay bits"),
              (hconv, hh - 1, "H: colour conversion -> row buffer"), (hh, END - 1, "H: horizontal pass -> rings")]
    INNER = [(sel, "H: row load, composite select, overlay bits"), (hl, "H: horizontal pass -> rings"), (hc, "H: horizontal pass -> rings"), (ov, "H: unit overlay build")]
else:
    B = find("__device__ __forceinline__ void frame_strips_body")
    P = find("=== producer warp ===", B); C = find("=== consumer warps ===", B)
    fill = find("---- fill our rows ourselves", C); mat = find("---- staged composite under text", C); st = find("---- text overlay, stamped", C)
    pa = find("---- phase A:", C); pb = find("---- phase B:", C)
    END = find("__global__ void __launch_bounds__(CTA_THREADS, 3)", pb) - 2
    sel = (find("__device__ __forceinline__ void select_staged") - 1, find("__device__ unsigned long long g_trace") - 1)
    RANGES = [(P, C - 1, "producer warp (TMA issue, chunk contexts)"), (C, fill - 1, "chunk set-up (waits, context)"), (fill, mat - 1, "rows not staged by TMA"), (mat, st - 1, "composite materialised under text"),
              (st, pa - 1, "glyph stamp"), (pa, pb - 1, "phase A: Y, depth Y, pair-summed chroma -> ring"), (pb, END, "phase B: 8-tap vertical chroma -> U, V")]
    INNER = [(sel, "phase A: composite select")]

dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
# every instruction is preceded by its inline chain, innermost frame first: one "//## File ... line N [inlined at ... line M]" per frame
loc = {}
chain, on, fresh, prev_op = [], False, True, ""
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        on = kname in m.group(1); chain = []; continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        op = m.group(2).strip()
        poll = ("SYNCS.PHASECHK" in op or "NANOSLEEP" in op) or ("BRA" in op and ("SYNCS.PHASECHK" in prev_op or "NANOSLEEP" in prev_op))
        loc[int(m.group(1), 16)] = (list(chain), poll)
        prev_op, fresh = op, True


def phase(entry):
    if entry is None:
        return "other"
    frames, poll = entry
    if poll:
        return "mbarrier wait loops (all roles)"
    for f in frames:  # innermost first: a helper with its own row in the table wins
        if f[0] == SRC:
            for (lo, hi), name in INNER:
                if lo <= f[1] <= hi:
                    return name
    for f in frames:  # else the innermost frame inside the kernel body
        if f[0] == SRC:
            for lo, hi, name in RANGES:
                if lo <= f[1] <= hi:
                    return name
    return "other"


rows = list(csv.reader(open(page)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = defaultdict(lambda: [0, 0])
base = None
ti = ts = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    n, s = int(r[ii] or 0), int(r[isamp] or 0)
    g = phase(loc.get(a - base))
    agg[g][0] += n; agg[g][1] += s; ti += n; ts += s
print(f"| phase | warp instructions | thread-instructions per pixel | share of instructions | share of warp time (stall samples) |")
print("|---|---|---|---|---|")
for g, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {g} | {n/1e6:.2f} M | {32*n/pixels:.1f} | {100*n/ti:.1f} % | {100*s/max(ts,1):.1f} % |")
print(f"| **total** | {ti/1e6:.2f} M | {32*ti/pixels:.1f} | | |")
