"""Synthetic inputs and workloads of the benchmark (SURVEY.md §8 d, BASELINE.json configs).

Closed-form frames so that the GPU arm, the CPU reference arm and the tests all see the
same bytes without shipping data:  frame f, pixel (x, y)
    R = (3x + 5y + 7f) & 255      G = ((x*x >> 3) + 11y + 13f) & 255
    B = (((x ^ y) * 7) + 17f) & 255      depth = (((x + 2y + 3f) * 5) >> 1) & 255
    composite tests: source k of n is transparent (A = 0) on stripes ((x >> 6) + k) % n != 0
"""
from __future__ import annotations

import os

import numpy as np

KINITIAL_CAMERA_MATRIX = [1.0, 0.0, 0.0, 0.5, 0.0, -1.0, 0.0, 0.5, 0.0, 0.0, -1.0, 0.5]  # camera_manager.h:18-19

CH_OFF = {"rgb24": (0, 1, 2, -1), "bgr24": (2, 1, 0, -1), "rgba": (0, 1, 2, 3), "bgra": (2, 1, 0, 3), "argb": (1, 2, 3, 0), "abgr": (3, 2, 1, 0)}


def synth_rgb(w: int, h: int, f: int = 0) -> np.ndarray:
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    r = (3 * x + 5 * y + 7 * f) & 255
    g = (((x * x) >> 3) + 11 * y + 13 * f) & 255
    b = (((x ^ y) * 7) + 17 * f) & 255
    return np.stack([r, g, b], axis=-1).astype(np.uint8)


def synth_depth(w: int, h: int, f: int = 0) -> np.ndarray:
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    return ((((x + 2 * y + 3 * f) * 5) >> 1) & 255).astype(np.uint8)


def synth_alpha(w: int, h: int, k: int, n: int) -> np.ndarray:
    x = np.arange(w, dtype=np.int64)[None, :]
    a = np.where(((x >> 6) + k) % n != 0, 0, 255).astype(np.uint8)
    return np.broadcast_to(a, (h, w)).copy()


def to_fmt(rgb: np.ndarray, fmt: str, alpha: np.ndarray | None = None) -> np.ndarray:
    ro, go, bo, ao = CH_OFF[fmt]
    h, w = rgb.shape[:2]
    out = np.zeros((h, w, 3 if ao < 0 else 4), np.uint8)
    out[..., ro], out[..., go], out[..., bo] = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    if ao >= 0:
        out[..., ao] = 255 if alpha is None else alpha
    return out


def reference_strings(index: int = 0, timestamp: str = "12:34:56.789", is_left: bool = True, matrix=None):
    """The four overlays of process_frame_thread in call order (encode.cpp:57-97): (position, text)."""
    from .api import RENDER_POSITION_CENTER, RENDER_POSITION_LEFT_BOTTOM, RENDER_POSITION_LEFT_TOP, RENDER_POSITION_RIGHT_TOP, format_camera_matrix
    return [
        (RENDER_POSITION_CENTER, format_camera_matrix(matrix or KINITIAL_CAMERA_MATRIX)),
        (RENDER_POSITION_LEFT_BOTTOM, b"index=" + str(index).encode()),
        (RENDER_POSITION_LEFT_TOP, timestamp.encode()),
        (RENDER_POSITION_RIGHT_TOP, b"direction=left" if is_left else b"direction=right"),
    ]


def dense_text(lines: int = 64, cols: int = 120) -> bytes:
    """BASELINE config 5's "dense per-frame text overlay": lines x cols printable characters."""
    return b"\n".join(bytes((33 + (i * 7 + j * 3) % 90) for i in range(cols)) for j in range(lines))


def load_glyph_table(path: str | None = None):
    """The committed Aileron 20 px glyph table (tests/golden/glyphs_aileron20.npz) ->
    (metrics int32 [256,5], bitmaps list)."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "glyphs_aileron20.npz")
    z = np.load(path)
    offs, cov = z["offsets"], z["coverage"]
    return z["metrics"].astype(np.int32), [np.ascontiguousarray(cov[offs[b]: offs[b + 1]]) for b in range(256)]


# name -> description of one "frame" of the workload
WORKLOADS = {
    # BASELINE.json configs[1]
    "c2_1080p_2src_composite": dict(w=1920, h=1080, wd=1920, hd=1080, fmt="rgba", n_src=2, text="reference", sessions=1,
                                    desc="2 x (1920x1080 RGBA + GRAY8) -> depth composite -> 4-string overlay -> scene+depth YUV420P 1080p"),
    # BASELINE.json configs[2]
    "c3_7680x2160_sbs": dict(w=7680, h=2160, wd=7680, hd=2160, fmt="rgb24", n_src=1, text="reference_sbs", sessions=1,
                             desc="7680x2160 side-by-side RGB24 + GRAY8 -> two 4-string overlays -> scene+depth YUV420P same size"),
    # BASELINE.json configs[3] (per session; sessions are sharded across GPUs)
    "c4_1080p_sessions": dict(w=1920, h=1080, wd=1920, hd=1080, fmt="rgb24", n_src=1, text="reference", sessions=64,
                              desc="1920x1080 RGB24 + GRAY8 -> 4-string overlay -> scene+depth YUV420P, 64 sessions sharded over the GPUs"),
    # BASELINE.json configs[4]
    "c5_4k_4src_to_1440p": dict(w=3840, h=2160, wd=2560, hd=1440, fmt="rgba", n_src=4, text="dense", sessions=1,
                                desc="4 x (3840x2160 RGBA + GRAY8) -> composite -> dense overlay -> bicubic resize -> scene+depth YUV420P 2560x1440"),
    # north_star target resolution (SURVEY.md §8 d "plain 4K RGB24+depth same-size")
    "4k_rgb24": dict(w=3840, h=2160, wd=3840, hd=2160, fmt="rgb24", n_src=1, text="reference", sessions=1,
                     desc="3840x2160 RGB24 + GRAY8 -> 4-string overlay -> scene+depth YUV420P same size"),
}


def algorithmic_bytes(wl: dict) -> int:
    """SURVEY.md §8 d: every input byte read once, scene + depth YUV420P written once."""
    bpp = 3 if wl["fmt"] in ("rgb24", "bgr24") else 4
    return wl["n_src"] * (bpp + 1) * wl["w"] * wl["h"] + 2 * (wl["wd"] * wl["hd"] * 3 // 2)


def text_runs(wl: dict, f: int):
    """(position, text[, view]) runs of frame f.  A view (x, y, w, h) makes that sub-rectangle the frame the
    reference's render_string_to_frame sees (nes_text_run.view_*): config 3 stamps one 4-string set per eye."""
    from .api import RENDER_POSITION_LEFT_TOP
    if wl["text"] == "reference":
        return reference_strings(index=f)
    if wl["text"] == "reference_sbs":  # SURVEY.md §8 d config 3: two 4-string sets, one per half
        half = wl["w"] // 2
        left = [(p, t, (0, 0, half, wl["h"])) for p, t in reference_strings(index=f, is_left=True)]
        right = [(p, t, (half, 0, half, wl["h"])) for p, t in reference_strings(index=f, is_left=False)]
        return left + right
    if wl["text"] == "dense":
        return [(RENDER_POSITION_LEFT_TOP, dense_text())]
    return []


def make_sources(wl: dict, f: int):
    """-> [(pixels uint8 [h,w,bpp], depth uint8 [h,w])] for frame f"""
    out = []
    for k in range(wl["n_src"]):
        rgb = synth_rgb(wl["w"], wl["h"], f * wl["n_src"] + k)
        alpha = synth_alpha(wl["w"], wl["h"], k, wl["n_src"]) if wl["n_src"] > 1 else None
        out.append((to_fmt(rgb, wl["fmt"], alpha), synth_depth(wl["w"], wl["h"], f * wl["n_src"] + k)))
    return out
