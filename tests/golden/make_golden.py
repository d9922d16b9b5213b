"""Regenerates the committed golden fixtures from the REAL libraries of this image
(libswscale 9.1.100 / FreeType 2.14.3 through oracle/_ref/libnes_ref.so, i.e. the
reference's glue of type_managers.cc:143-155 and render_text.cc:10-111 run here).

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/glyphs_aileron20.npz, tests/golden/golden.json.
The GPU box does not need the libraries: tests compare against these files and against
the C restatement (oracle/liboracle_port.so), which test_oracle.py pins to the same files.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

FONT = os.path.join(HERE, "Aileron-Regular.ttf")


def sha16(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


def main():
    R = O.Ref()
    assert R.have_sws and R.have_ft, "bundled libswscale / FreeType not found"
    out = {"swscale": R.swscale_version(), "flags": "SWS_BITEXACT|SWS_ACCURATE_RND", "convert": [], "overlay": [], "formats": []}
    t = R.text_new(FONT)
    out["freetype"] = R.freetype_version(t)
    out["font_sha16"] = sha16(open(FONT, "rb").read())
    gt = R.glyph_table(t)
    gt.save(os.path.join(HERE, "glyphs_aileron20.npz"))

    sizes = [(64, 32, 64, 32), (96, 54, 64, 36), (130, 46, 130, 46), (1280, 720, 1280, 720), (1920, 1080, 1920, 1080),
             (3840, 2160, 3840, 2160), (3840, 2160, 2560, 1440), (384, 216, 256, 144), (256, 144, 384, 216), (200, 100, 120, 90),
             (128, 72, 192, 108), (96, 54, 64, 54), (640, 360, 426, 240), (1280, 720, 1920, 1080), (1920, 1080, 1280, 720)]
    for (w, h, wd, hd) in sizes:
        rgb, dep = O.synth_rgb(w, h), O.synth_depth(w, h)
        s = R.sws_convert(rgb, "rgb24", wd, hd)
        d = R.sws_convert(dep, "gray", wd, hd)
        out["convert"].append({"src": [w, h], "dst": [wd, hd], "scene": sha16(s.cropped()), "depth": sha16(d.cropped()),
                               "y0": s.y[0, :4].tolist(), "u0": s.u[0, :4].tolist(), "v0": s.v[0, :4].tolist(), "dy0": d.y[0, :4].tolist()})
    # pixel-format variants give byte-identical YUV on the same pixels
    rgb = O.synth_rgb(320, 180)
    for fmt in ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"]:
        for (wd, hd) in [(320, 180), (214, 120)]:
            s = R.sws_convert(O.to_fmt(rgb, fmt), fmt, wd, hd)
            out["formats"].append({"fmt": fmt, "src": [320, 180], "dst": [wd, hd], "scene": sha16(s.cropped())})
    # gray LUT
    lut = R.sws_convert(np.tile(np.arange(256, dtype=np.uint8), (4, 1)), "gray")
    out["gray_lut"] = lut.y[0, :256].tolist()
    out["gray_lut_sha16"] = sha16(lut.y[0, :256].tobytes())
    # overlay + convert (encode.cpp:76-98 order)
    for (w, h) in [(1280, 720), (1920, 1080)]:
        surf = np.ascontiguousarray(O.synth_rgb(w, h))
        stamped = 0
        for pos, txt in O.reference_strings():
            stamped += R.text_render(t, surf, pos, txt)
        s = R.sws_convert(surf, "rgb24")
        out["overlay"].append({"size": [w, h], "stamped": int((surf != O.synth_rgb(w, h)).any(axis=2).sum()), "stamp_calls": int(stamped),
                               "rgb": sha16(surf.tobytes()), "yuv": sha16(s.cropped())})
    # BASELINE configs at their stated sizes, frame 0 of the benchmark's synthetic workloads
    # (composite through the C port -- the reference has none --, text through the real FreeType,
    # conversion through the real libswscale)
    import ngp_encode_server_b200 as n
    P = O.Port()
    out["configs"] = {}
    for name, wl in n.synth.WORKLOADS.items():
        srcs = n.synth.make_sources(wl, 0)
        runs = n.synth.text_runs(wl, 0)
        sc, dp = O.expected_frame(srcs, wl["fmt"], runs, wl["wd"], wl["hd"], P, ref=R, tctx=t)
        out["configs"][name] = {"frame": 0, "src": [wl["w"], wl["h"]], "dst": [wl["wd"], wl["hd"]], "fmt": wl["fmt"], "n_src": wl["n_src"],
                                "n_runs": len(runs), "scene": sha16(sc.cropped()), "depth": sha16(dp.cropped())}
    R.text_free(t)
    # constant frame
    c = R.sws_convert(np.full((16, 16, 3), 200, np.uint8), "rgb24")
    out["const200"] = [int(c.y[0, 0]), int(c.u[0, 0]), int(c.v[0, 0])]
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
