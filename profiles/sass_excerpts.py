#!/usr/bin/env python
"""SASS evidence of the in-tree library: per kernel instruction / mnemonic counts and the producer's TMA excerpt.
usage: python profiles/sass_excerpts.py [libnes_gpu.so] > profiles/r02_sass_excerpts.txt"""
import os, re, subprocess, sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ngp-encode-server_b200", "libnes_gpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("# SASS evidence, round 2 (cuobjdump -sass ngp-encode-server_b200/libnes_gpu.so; sm_100a)")
print("# per kernel: instruction count, and the counts of the mnemonics that show the Blackwell data path")
print("#   UTMALDG = cp.async.bulk.tensor (TMA tile load)   SYNCS = mbarrier ops   IDP.2A = dp2a   PREEXIT = griddepcontrol.launch_dependents")
print("#   NANOSLEEP.SYNCS = the suspend-time hint of mbarrier.try_wait   (no UTMASTG: outputs are 4/8/16-byte STG from registers -- the kernels are issue-bound, not store-bound)")
print()
fn, body = None, {}
for l in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        fn = m.group(1); body[fn] = []; continue
    if fn and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", l):
        body[fn].append(l)
shown = set()
for fn, lines in body.items():
    ops = Counter()
    forms = set()
    for l in lines:
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m: continue
        op = m.group(1)
        ops[op.split(".")[0]] += 1
        if op.startswith("IDP.2A"): ops["IDP.2A"] += 1
        if op.startswith("SYNCS") or op.startswith("NANOSLEEP"): forms.add(op)
    if not any(k in fn for k in ("k_frame_strips", "k_resize_strips", "k_resize_tiles", "k_depth16")):
        continue
    print(fn)
    print("    instructions %d  UTMALDG %d  UTMASTG %d  SYNCS %d  IDP.2A %d  LDS %d  STS %d  LDG %d  STG %d  IMAD %d  BAR %d  LDL %d  PREEXIT %d  NANOSLEEP %d  HMMA %d" % (
        len(lines), ops["UTMALDG"], ops["UTMASTG"], ops["SYNCS"], ops["IDP.2A"], ops["LDS"], ops["STS"], ops["LDG"], ops["STG"], ops["IMAD"], ops["BAR"], ops["LDL"],
        ops["PREEXIT"], ops["NANOSLEEP"], ops["HMMA"]))
    base = re.sub(r"ILi\dE(Li\dE)?", "", fn)
    if ops["UTMALDG"] and base not in shown:
        shown.add(base)
        i = next(k for k, l in enumerate(lines) if "UTMALDG" in l)
        print("    excerpt (the producer's tensor-map copy and the mbarrier around it):")
        for l in lines[max(0, i - 6):i + 3]:
            print("      " + l.strip())
        print("    mbarrier / sleep forms used: " + ", ".join(sorted(forms)))
    if ops["PREEXIT"] and base + "#pdl" not in shown:
        shown.add(base + "#pdl")
        i = next(k for k, l in enumerate(lines) if "PREEXIT" in l)
        print("    first instructions (programmatic dependent launch trigger):")
        for l in lines[max(0, i - 2):i + 2]:
            print("      " + l.strip())
