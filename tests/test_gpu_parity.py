"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
(oracle/liboracle_port.so, pinned to libswscale 9.1.100 / FreeType 2.14.3 by
tests/test_oracle.py) and against the committed golden vectors.  Bit-exact everywhere:
this is integer / byte work."""
import os

import numpy as np
import pytest

from conftest import sha16

pytestmark = pytest.mark.gpu


def run_gpu(N, s, fmt, srcs, w, h, wd=None, hd=None, runs=None, want_depth=True, pinned=False):
    """srcs: [(rgb[h,w,bpp], depth[h,w] or None)] -> (scene FrameManager, depth FrameManager or None)"""
    wd, hd = wd or w, hd or h
    scene = N.FrameManager(N.FrameContext(wd, hd, "yuv420p"), session=s if pinned else None)
    depth = N.FrameManager(N.FrameContext(wd, hd, "yuv420p"), session=s if pinned else None) if want_depth else None
    keep = []
    sources = []
    for rgb, dep in srcs:
        rgb = np.ascontiguousarray(rgb)
        keep.append(rgb)
        if dep is not None:
            dep = np.ascontiguousarray(dep)
            keep.append(dep)
        sources.append((rgb.reshape(-1), None if dep is None else dep.reshape(-1), 0, 0))
    fin = N.Session.frame_in(fmt, w, h, sources)
    s.convert(fin, runs, N.api._frame_out(scene, depth))
    return scene, depth


def first_diff(a: bytes, b: bytes):
    x, y = np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8)
    if x.size != y.size:
        return f"size {x.size} vs {y.size}"
    d = np.nonzero(x != y)[0]
    return "equal" if d.size == 0 else f"{d.size} bytes differ, first at {d[0]}: got {x[d[0]]} want {y[d[0]]}"


# ------------------------------------------------------------------ same-size path
@pytest.mark.parametrize("w,h", [(64, 32), (130, 46), (258, 70), (256, 32), (512, 64), (8, 8), (4, 4), (6, 10), (1280, 720), (1920, 1080), (254, 38)])
def test_same_size_rgb24_and_depth(N, O, port, session, w, h):
    rng = np.random.default_rng(w * 10007 + h)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
    sc, dp = run_gpu(N, session, "rgb24", [(rgb, dep)], w, h)
    assert sc.cropped() == port.rgb_to_yuv420p(rgb, "rgb24").cropped(), first_diff(sc.cropped(), port.rgb_to_yuv420p(rgb, "rgb24").cropped())
    assert dp.cropped() == port.gray_to_yuv420p(dep).cropped(), first_diff(dp.cropped(), port.gray_to_yuv420p(dep).cropped())


def test_golden_vectors(N, O, session, golden):
    for c in golden["convert"]:
        (w, h), (wd, hd) = c["src"], c["dst"]
        sc, dp = run_gpu(N, session, "rgb24", [(O.synth_rgb(w, h), O.synth_depth(w, h))], w, h, wd, hd)
        assert sha16(sc.cropped()) == c["scene"], c
        assert sha16(dp.cropped()) == c["depth"], c
        assert sc.planes[0][0, :4].tolist() == c["y0"] and sc.planes[1][0, :4].tolist() == c["u0"] and sc.planes[2][0, :4].tolist() == c["v0"]


@pytest.mark.parametrize("fmt", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
def test_pixel_formats(N, O, port, session, golden, fmt):
    rgb = O.synth_rgb(320, 180)
    img = O.to_fmt(rgb, fmt)
    for c in golden["formats"]:
        if c["fmt"] != fmt:
            continue
        sc, _ = run_gpu(N, session, fmt, [(img, None)], 320, 180, *c["dst"], want_depth=False)
        assert sha16(sc.cropped()) == c["scene"], c
    rng = np.random.default_rng(5)
    bpp = N.PIX_BPP[fmt]
    rnd = rng.integers(0, 256, (94, 386, bpp), dtype=np.uint8)
    sc, _ = run_gpu(N, session, fmt, [(rnd, None)], 386, 94, want_depth=False)
    assert sc.cropped() == port.rgb_to_yuv420p(rnd, fmt).cropped()


def test_extremes_and_constant(N, port, session, golden):
    ext = np.zeros((64, 512, 3), np.uint8)
    ext[:, :128] = (255, 0, 0); ext[:, 128:256] = (0, 0, 255); ext[:, 256:384] = (0, 255, 0); ext[::2, 384:] = 255
    sc, _ = run_gpu(N, session, "rgb24", [(ext, None)], 512, 64, want_depth=False)
    assert sc.cropped() == port.rgb_to_yuv420p(ext, "rgb24").cropped()
    c, _ = run_gpu(N, session, "rgb24", [(np.full((16, 16, 3), 200, np.uint8), None)], 16, 16, want_depth=False)
    assert [int(c.planes[0][0, 0]), int(c.planes[1][0, 0]), int(c.planes[2][0, 0])] == golden["const200"]
    lut_in = np.tile(np.arange(256, dtype=np.uint8), (4, 1))
    _, d = run_gpu(N, session, "rgb24", [(np.zeros((4, 256, 3), np.uint8), lut_in)], 256, 4)
    assert d.planes[0][0, :256].tolist() == golden["gray_lut"]
    assert (d.planes[1][:, :128] == 128).all() and (d.planes[2][:, :128] == 128).all()


def test_strided_and_pinned_buffers(N, O, port, session):
    """Row strides wider than the row, pinned vs pageable host memory, unaligned source base."""
    w, h = 322, 58
    rng = np.random.default_rng(11)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
    want_s, want_d = port.rgb_to_yuv420p(rgb, "rgb24").cropped(), port.gray_to_yuv420p(dep).cropped()
    # strided pageable
    rs, ds = w * 3 + 10, w + 6
    big = np.zeros((h, rs), np.uint8); big[:, : w * 3] = rgb.reshape(h, -1)
    bigd = np.zeros((h, ds), np.uint8); bigd[:, :w] = dep
    scene = N.FrameManager(N.FrameContext(w, h, "yuv420p")); depth = N.FrameManager(N.FrameContext(w, h, "yuv420p"))
    fin = N.Session.frame_in("rgb24", w, h, [(big.reshape(-1), bigd.reshape(-1), rs, ds)])
    session.convert(fin, None, N.api._frame_out(scene, depth))
    assert scene.cropped() == want_s and depth.cropped() == want_d
    # pinned in and out, source base offset by 1 byte (protobuf payloads are unaligned)
    pin = session.host_array(w * h * 3 + 1); pin[1:] = rgb.reshape(-1)
    pind = session.host_array(w * h + 3); pind[3:] = dep.reshape(-1)
    scene = N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=session); depth = N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=session)
    fin = N.Session.frame_in("rgb24", w, h, [(pin[1:], pind[3:], 0, 0)])
    session.convert(fin, None, N.api._frame_out(scene, depth))
    assert scene.cropped() == want_s and depth.cropped() == want_d
    t = session.last_timing()
    assert t["n_launches"] >= 1 and t["kernels_us"] > 0


# ------------------------------------------------------------------ overlay
def test_overlay_golden(N, O, session, golden):
    for c in golden["overlay"]:
        w, h = c["size"]
        sc, _ = run_gpu(N, session, "rgb24", [(O.synth_rgb(w, h), None)], w, h, runs=O.reference_strings(), want_depth=False)
        assert sha16(sc.cropped()) == c["yuv"], c


@pytest.mark.parametrize("w,h,fmt", [(640, 360, "rgb24"), (200, 120, "rgb24"), (1280, 720, "bgra"), (700, 300, "argb")])
def test_overlay_vs_oracle(N, O, port, glyphs, session, w, h, fmt):
    rng = np.random.default_rng(w)
    rgb = rng.integers(0, 128, (h, w, 3), dtype=np.uint8)
    runs = O.reference_strings(index=987654321, is_left=False) + [(O.POS_RIGHT_BOTTOM, b"clipped at the right edge \xe9\xff~ and beyond the frame")]
    surf = rgb.copy()
    for pos, txt in runs:
        port.render_string(surf, pos, txt, glyphs)
    want = port.rgb_to_yuv420p(O.to_fmt(surf, fmt), fmt)
    sc, _ = run_gpu(N, session, fmt, [(O.to_fmt(rgb, fmt), None)], w, h, runs=runs, want_depth=False)
    assert sc.cropped() == want.cropped(), first_diff(sc.cropped(), want.cropped())


def test_dense_overlay(N, O, port, glyphs, session):
    """64 lines x 120 characters tiled over the frame (BASELINE config 5's overlay): > HIT_CAP glyphs."""
    w, h = 1920, 1080
    rgb = O.synth_rgb(w, h, 2)
    line = bytes((33 + (i * 7) % 90) for i in range(120))
    text = b"\n".join(line for _ in range(64))
    runs = [(O.POS_LEFT_TOP, text)]
    surf = np.ascontiguousarray(rgb.copy())
    port.render_string(surf, O.POS_LEFT_TOP, text, glyphs)
    want = port.rgb_to_yuv420p(surf, "rgb24")
    sc, _ = run_gpu(N, session, "rgb24", [(rgb, None)], w, h, runs=runs, want_depth=False)
    assert sc.cropped() == want.cropped(), first_diff(sc.cropped(), want.cropped())


def test_text_layout_matches_oracle_stamp(N, O, port, glyphs, session):
    w, h = 400, 200
    for pos, txt in O.reference_strings(index=42):
        placed = session.text_layout(w, h, pos, txt)
        a = np.zeros((h, w, 3), np.uint8)
        for (x, y, code) in placed:
            gw, gr = int(glyphs.metrics[code][0]), int(glyphs.metrics[code][1])
            bm = glyphs.bitmaps[code].reshape(gr, gw)
            for q in range(gr):
                for p in range(gw):
                    if bm[q, p] and 0 <= x + p < w and 0 <= y + q < h:
                        a[y + q, x + p] = 255
        b = np.zeros((h, w, 3), np.uint8)
        port.render_string(b, pos, txt, glyphs)
        assert np.array_equal(a, b)


# ------------------------------------------------------------------ composite
@pytest.mark.parametrize("n,w,h,fmt", [(2, 640, 360, "rgba"), (4, 322, 94, "bgra"), (3, 256, 64, "argb"), (2, 130, 46, "rgb24"), (2, 1920, 1080, "rgba")])
def test_composite_same_size(N, O, port, session, n, w, h, fmt):
    rng = np.random.default_rng(n * 100 + w)
    srcs, rgbs, deps = [], [], []
    for k in range(n):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        alpha = O.synth_alpha(w, h, k, n) if k else np.where(rng.integers(0, 4, (h, w)) == 0, 0, 255).astype(np.uint8)
        img = O.to_fmt(rgb, fmt, alpha)
        dep = rng.integers(0, 8, (h, w), dtype=np.uint8) * 32 if k % 2 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        srcs.append((img, dep)); rgbs.append(img); deps.append(np.ascontiguousarray(dep))
    comp, cdep = port.composite(rgbs, deps, fmt)
    want_s, want_d = port.rgb_to_yuv420p(comp, fmt), port.gray_to_yuv420p(cdep)
    sc, dp = run_gpu(N, session, fmt, srcs, w, h)
    assert sc.cropped() == want_s.cropped(), first_diff(sc.cropped(), want_s.cropped())
    assert dp.cropped() == want_d.cropped(), first_diff(dp.cropped(), want_d.cropped())


def test_composite_overlay_config2(N, O, port, glyphs, session):
    """BASELINE config 2 in small: 2 RGBA+depth sources -> composite -> 4 overlays -> scene+depth."""
    w, h, n = 960, 540, 2
    srcs, rgbs, deps = [], [], []
    for k in range(n):
        img = O.to_fmt(O.synth_rgb(w, h, k), "rgba", O.synth_alpha(w, h, k, n))
        dep = O.synth_depth(w, h, k)
        srcs.append((img, dep)); rgbs.append(img); deps.append(dep)
    comp, cdep = port.composite(rgbs, deps, "rgba")
    surf = np.ascontiguousarray(comp[..., :3])
    for pos, txt in O.reference_strings():
        port.render_string(surf, pos, txt, glyphs)
    want_s, want_d = port.rgb_to_yuv420p(surf, "rgb24"), port.gray_to_yuv420p(cdep)
    sc, dp = run_gpu(N, session, "rgba", srcs, w, h, runs=O.reference_strings())
    assert sc.cropped() == want_s.cropped() and dp.cropped() == want_d.cropped()


# ------------------------------------------------------------------ resize path
@pytest.mark.parametrize("w,h,wd,hd", [(96, 54, 64, 36), (384, 216, 256, 144), (256, 144, 384, 216), (200, 100, 120, 90), (128, 72, 192, 108),
                                       (96, 54, 64, 54), (640, 360, 426, 240), (322, 182, 214, 120), (160, 90, 480, 270), (640, 360, 212, 120), (64, 64, 64, 32), (64, 32, 128, 32)])
def test_resize_vs_oracle(N, O, port, session, w, h, wd, hd):
    rng = np.random.default_rng(w + wd)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
    sc, dp = run_gpu(N, session, "rgb24", [(rgb, dep)], w, h, wd, hd)
    want_s, want_d = port.rgb_to_yuv420p(rgb, "rgb24", wd, hd), port.gray_to_yuv420p(dep, wd, hd)
    assert sc.cropped() == want_s.cropped(), first_diff(sc.cropped(), want_s.cropped())
    assert dp.cropped() == want_d.cropped(), first_diff(dp.cropped(), want_d.cropped())


def test_resize_composite_overlay_config5_small(N, O, port, glyphs, session):
    """BASELINE config 5 at 1/4 scale: 4 RGBA+depth sources -> composite -> dense overlay -> 3:2 downscale."""
    w, h, wd, hd, n = 960, 540, 640, 360, 4
    srcs, rgbs, deps = [], [], []
    for k in range(n):
        img = O.to_fmt(O.synth_rgb(w, h, k), "rgba", O.synth_alpha(w, h, k, n))
        dep = O.synth_depth(w, h, k)
        srcs.append((img, dep)); rgbs.append(img); deps.append(dep)
    comp, cdep = port.composite(rgbs, deps, "rgba")
    text = b"\n".join(bytes((40 + (i * 5 + j) % 80) for i in range(60)) for j in range(20))
    surf = np.ascontiguousarray(comp[..., :3])
    port.render_string(surf, O.POS_LEFT_TOP, text, glyphs)
    want_s, want_d = port.rgb_to_yuv420p(surf, "rgb24", wd, hd), port.gray_to_yuv420p(cdep, wd, hd)
    sc, dp = run_gpu(N, session, "rgba", srcs, w, h, wd, hd, runs=[(O.POS_LEFT_TOP, text)])
    assert sc.cropped() == want_s.cropped(), first_diff(sc.cropped(), want_s.cropped())
    assert dp.cropped() == want_d.cropped()


@pytest.mark.parametrize("ns", [2, 3, 5])
def test_resize_with_shallow_substage_rings(ns):
    """k_resize_strips walks a chunk of 4 sub-stages through a ring of `ns` slots (2 ... 8, whatever shared memory is left after the
    rings); the slot / parity arithmetic of the three warp roles must hold for every ns, also when a chunk wraps the ring more
    than once.  NES_RZ_NS is read once per process, so the resize tests are re-run in a child process per value."""
    import subprocess
    import sys
    env = dict(os.environ, NES_RZ_NS=str(ns))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "test_resize_vs_oracle or test_resize_composite_overlay_config5_small or test_nv12_output or test_batch_mixed_jobs"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


# ------------------------------------------------------------------ API behaviour
def test_errors(N, session):
    rgb = np.zeros(64 * 32 * 3, np.uint8)
    scene = N.FrameManager(N.FrameContext(64, 32, "yuv420p"))
    fo = N.api._frame_out(scene, None)
    for bad, status in [(dict(w=63, h=32), N.NES_ERR_INVALID_ARG), (dict(w=64, h=34), N.NES_ERR_SHORT_BUFFER), (dict(w=20000, h=32), N.NES_ERR_TOO_LARGE)]:
        fin = N.Session.frame_in("rgb24", bad["w"], bad["h"], [(rgb, None, 0, 0)])
        with pytest.raises(N.NesGpuError) as e:
            session.convert(fin, None, fo)
        assert e.value.status == status
    with pytest.raises(N.NesGpuError) as e:
        session.wait(123456)
    assert e.value.status == N.NES_ERR_BAD_TICKET
    s2 = N.Session(device=0, max_width=64, max_height=32, max_sources=1)
    fin = N.Session.frame_in("rgb24", 64, 32, [(rgb, None, 0, 0)])
    with pytest.raises(N.NesGpuError) as e:
        s2.convert(fin, [(0, b"x")], fo)
    assert e.value.status == N.NES_ERR_NO_ATLAS
    s2.convert(fin, None, fo)  # the failed call left the session usable
    s2.close()


def test_async_ring_and_batch(N, O, port, session):
    """submit/wait with 3 frames in flight, ring-full error, and the batched device entry point."""
    w, h = 512, 96
    frames = [(O.synth_rgb(w, h, f), O.synth_depth(w, h, f)) for f in range(5)]
    want = [(port.rgb_to_yuv420p(r, "rgb24").cropped(), port.gray_to_yuv420p(d).cropped()) for r, d in frames]
    outs, tickets, keep = [], [], []
    for f in range(3):
        sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=session), N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=session)
        rgb, dep = np.ascontiguousarray(frames[f][0]).reshape(-1), np.ascontiguousarray(frames[f][1]).reshape(-1)
        keep += [rgb, dep]
        fin = N.Session.frame_in("rgb24", w, h, [(rgb, dep, 0, 0)])
        tickets.append(session.submit(fin, None, N.api._frame_out(sc, dp)))
        outs.append((sc, dp))
    with pytest.raises(N.NesGpuError) as e:
        session.submit(fin, None, N.api._frame_out(sc, dp))
    assert e.value.status == N.NES_ERR_BUSY
    for f in (2, 0, 1):
        session.wait(tickets[f])
        assert outs[f][0].cropped() == want[f][0] and outs[f][1].cropped() == want[f][1]
    # batched, device resident
    ysz, csz = N.align32(w) * h, N.align32(w // 2) * (h // 2)
    fins, fouts, devs = [], [], []
    for f in range(5):
        d_rgb, d_dep, d_s, d_d = (session.device_alloc(n) for n in (w * h * 3, w * h, ysz + 2 * csz, ysz + 2 * csz))
        session.h2d(d_rgb, frames[f][0]); session.h2d(d_dep, frames[f][1])
        fins.append(N.Session.frame_in("rgb24", w, h, [((d_rgb, w * h * 3), (d_dep, w * h), 0, 0)], mem=N.NES_MEM_DEVICE))
        fo = N.nes_frame_out(); fo.width, fo.height, fo.mem = w, h, N.NES_MEM_DEVICE
        for p, (off, ls) in enumerate([(0, N.align32(w)), (ysz, N.align32(w // 2)), (ysz + csz, N.align32(w // 2))]):
            fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off, ls, d_d + off, ls
        fouts.append(fo); devs.append((d_rgb, d_dep, d_s, d_d))
    before = session.launches
    session.convert_batch_device(fins, None, fouts, sync=True)
    assert session.launches - before == 1  # five frames, one launch
    for f in range(5):
        sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p")), N.FrameManager(N.FrameContext(w, h, "yuv420p"))
        session.d2h(sc.buffer, devs[f][2]); session.d2h(dp.buffer, devs[f][3])
        assert sc.cropped() == want[f][0] and dp.cropped() == want[f][1]
        for p in devs[f]:
            session.device_free(p)


def test_prepared_batch(N, O, port, glyphs, session):
    """nes_gpu_batch_prepare / run / free: a descriptor table built once and launched several times back to back
    (consecutive launches overlap on the device); mixed text / no text frames, every output checked after each round
    of runs with fresh output buffers (zeroed in between)."""
    w, h = 768, 200
    n = 6
    frames = [(O.synth_rgb(w, h, f), O.synth_depth(w, h, f)) for f in range(n)]
    runs = [O.reference_strings(index=f) if f % 2 == 0 else None for f in range(n)]
    want = []
    for f in range(n):
        sc, dp = O.expected_frame([frames[f]], "rgb24", runs[f], w, h, port, glyphs)
        want.append((sc.cropped(), dp.cropped()))
    ysz, csz = N.align32(w) * h, N.align32(w // 2) * (h // 2)
    fins, fouts, devs = [], [], []
    for f in range(n):
        d_rgb, d_dep, d_s, d_d = (session.device_alloc(k) for k in (w * h * 3, w * h, ysz + 2 * csz, ysz + 2 * csz))
        session.h2d(d_rgb, frames[f][0]); session.h2d(d_dep, frames[f][1])
        fins.append(N.Session.frame_in("rgb24", w, h, [((d_rgb, w * h * 3), (d_dep, w * h), 0, 0)], mem=N.NES_MEM_DEVICE))
        fo = N.nes_frame_out(); fo.width, fo.height, fo.mem = w, h, N.NES_MEM_DEVICE
        for p, (off, ls) in enumerate([(0, N.align32(w)), (ysz, N.align32(w // 2)), (ysz + csz, N.align32(w // 2))]):
            fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off, ls, d_d + off, ls
        fouts.append(fo); devs.append((d_rgb, d_dep, d_s, d_d))
    b = session.batch_prepare(fins, [r or [] for r in runs], fouts)
    zero = np.zeros(ysz + 2 * csz, np.uint8)
    try:
        for rounds in (1, 5):
            for f in range(n):
                session.h2d(devs[f][2], zero); session.h2d(devs[f][3], zero)
            before = session.launches
            for _ in range(rounds):
                session.batch_run(b)
            session.batch_run(b, sync=True)
            assert session.launches - before == rounds + 1
            for f in range(n):
                sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p")), N.FrameManager(N.FrameContext(w, h, "yuv420p"))
                session.d2h(sc.buffer, devs[f][2]); session.d2h(dp.buffer, devs[f][3])
                assert sc.cropped() == want[f][0], (f, first_diff(sc.cropped(), want[f][0]))
                assert dp.cropped() == want[f][1], f
    finally:
        session.batch_free(b)
        for d in devs:
            for p in d:
                session.device_free(p)


def test_reference_call_sequence(N, O, port, glyphs, session):
    """The reference's own flow (server.cpp:172-194 -> encode.cpp:55-98) through the mirrored
    classes: wire bytes -> RenderedFrame -> 4x render_string_to_frame -> convert_frame."""
    import os
    from conftest import FONT
    w, h = 640, 360
    rgb, dep = O.synth_rgb(w, h, 7), O.synth_depth(w, h, 7)
    msg = O.pack_rendered_frame(7, True, w, h, O.KINITIAL_CAMERA_MATRIX, rgb.tobytes(), dep.tobytes())
    s = N.Session(device=0, max_width=w, max_height=h, max_sources=1)
    try:
        if N.find_freetype() is None:
            pytest.skip("no FreeType binary in this image")
        etctx = N.RenderTextContext(FONT, s)
        frame = N.RenderedFrame(msg, "rgb24", "gray", (w, h), (w, h), s)
        N.api.process_frame(frame, etctx, "12:34:56.789")
        with pytest.raises(RuntimeError):
            frame.convert_frame()
        surf = np.ascontiguousarray(rgb.copy())
        for pos, txt in O.reference_strings(index=7):
            port.render_string(surf, pos, txt, glyphs)
        assert frame.converted_frame_scene().cropped() == port.rgb_to_yuv420p(surf, "rgb24").cropped()
        assert frame.converted_frame_depth().cropped() == port.gray_to_yuv420p(dep).cropped()
        # SwsContextManager on its own (type_managers.cc:143-155), scene and depth
        dst = N.FrameManager(N.FrameContext(w, h, "yuv420p"))
        N.SwsContextManager(N.FrameManager(N.FrameContext(w, h, "rgb24"), np.ascontiguousarray(rgb).reshape(-1)), dst, s)
        assert dst.cropped() == port.rgb_to_yuv420p(rgb, "rgb24").cropped()
        dst = N.FrameManager(N.FrameContext(w, h, "yuv420p"))
        N.SwsContextManager(N.FrameManager(N.FrameContext(w, h, "gray"), np.ascontiguousarray(dep).reshape(-1)), dst, s)
        assert dst.cropped() == port.gray_to_yuv420p(dep).cropped()
    finally:
        s.close()


# ------------------------------------------------------------------ BASELINE configs at their stated sizes
@pytest.mark.parametrize("name", ["c2_1080p_2src_composite", "c3_7680x2160_sbs", "c4_1080p_sessions", "c5_4k_4src_to_1440p", "4k_rgb24"])
def test_config_size_golden(N, O, port, glyphs, session, golden, name):
    """Every BASELINE config at full size, frame 0 of the benchmark's own synthetic workload, through the C ABI
    with host buffers: the hashes are the real libswscale / FreeType ones (tests/golden/make_golden.py) and the
    planes equal the port byte for byte.  Config 3 carries one overlay set per eye (text run views), config 5 is
    4 x (4K RGBA + depth) -> composite -> dense overlay -> 2560x1440."""
    c, wl = golden["configs"][name], N.synth.WORKLOADS[name]
    srcs, runs = N.synth.make_sources(wl, c["frame"]), N.synth.text_runs(wl, c["frame"])
    sc, dp = run_gpu(N, session, wl["fmt"], srcs, wl["w"], wl["h"], wl["wd"], wl["hd"], runs=runs, pinned=True)
    want_s, want_d = O.expected_frame(srcs, wl["fmt"], runs, wl["wd"], wl["hd"], port, glyphs)
    assert sc.cropped() == want_s.cropped(), first_diff(sc.cropped(), want_s.cropped())
    assert dp.cropped() == want_d.cropped(), first_diff(dp.cropped(), want_d.cropped())
    assert sha16(sc.cropped()) == c["scene"] and sha16(dp.cropped()) == c["depth"]


def test_text_run_views(N, O, port, glyphs, session):
    """nes_text_run.view_*: pen placement and clipping relative to a sub-rectangle (an eye of a side-by-side
    frame); text that runs over the view's edge must not spill into the neighbouring eye, views partly outside
    the frame are cut to it."""
    w, h = 1024, 288
    rgb = O.synth_rgb(w, h, 4)
    long_line = b"a long line that certainly runs over the right edge of a 512 pixel wide eye, and then some more"
    runs = [(O.POS_RIGHT_TOP, long_line, (0, 0, 512, h)), (O.POS_LEFT_BOTTOM, b"left eye", (0, 0, 512, h)),
            (O.POS_CENTER, O.format_matrix_text(O.KINITIAL_CAMERA_MATRIX), (512, 0, 512, h)), (O.POS_RIGHT_TOP, long_line, (512, 0, 512, h)),
            (O.POS_LEFT_TOP, b"odd view", (301, 33, 333, 111)), (O.POS_RIGHT_BOTTOM, b"view sticking out", (900, 200, 400, 300)),
            (O.POS_LEFT_TOP, b"whole frame")]
    surf = np.ascontiguousarray(rgb.copy())
    # the last view is cut to the frame for the oracle (pen placement still uses the full 400 x 300 view)
    for r in runs:
        if len(r) > 2 and r[2][0] + r[2][2] > w:
            x, y, vw, vh = r[2]
            big = np.zeros((vh, vw, 3), np.uint8)
            big[:h - y, :w - x] = surf[y:, x:]
            port.render_string(big, r[0], r[1], glyphs)
            surf[y:, x:] = big[:h - y, :w - x]
        else:
            O.stamp_runs(surf, [r], "rgb24", port, glyphs)
    sc, _ = run_gpu(N, session, "rgb24", [(rgb, None)], w, h, runs=runs, want_depth=False)
    want = port.rgb_to_yuv420p(surf, "rgb24")
    assert sc.cropped() == want.cropped(), first_diff(sc.cropped(), want.cropped())
    assert (surf != rgb).any()


# ------------------------------------------------------------------ full BASELINE sizes: properties + golden
def test_full_size_properties(N, O, session, golden):
    """At 4K / 7680x2160 the oracle is too slow for a fuzz; check golden hashes (4K) and
    size-independent properties: tiling invariance (a frame converted whole equals the same
    rows converted as a 64-row-aligned crop away from the crop's borders), depth LUT
    pointwise-ness, and idempotence of the stamp."""
    w, h = 7680, 2160
    rgb, dep = O.synth_rgb(w, h, 1), O.synth_depth(w, h, 1)
    sc, dp = run_gpu(N, session, "rgb24", [(rgb, dep)], w, h, pinned=True)
    lut = np.array(golden["gray_lut"], np.uint8)
    assert np.array_equal(dp.planes[0][:, :w], lut[dep])
    # crop rows [512, 1024): interior chroma rows (away from the crop's folded borders) must agree
    crop, _ = run_gpu(N, session, "rgb24", [(rgb[512:1024], None)], w, 512, want_depth=False)
    assert np.array_equal(sc.planes[0][512:1024, :w], crop.planes[0][:, :w])
    assert np.array_equal(sc.planes[1][256 + 2:512 - 2, : w // 2], crop.planes[1][2:-2, : w // 2])
    assert np.array_equal(sc.planes[2][256 + 2:512 - 2, : w // 2], crop.planes[2][2:-2, : w // 2])
    # left/right halves are independent eyes (side-by-side): converting the left half alone matches
    left, _ = run_gpu(N, session, "rgb24", [(np.ascontiguousarray(rgb[:, :3840]), None)], 3840, h, want_depth=False)
    assert np.array_equal(sc.planes[0][:, :3840], left.planes[0][:, :3840])
    assert np.array_equal(sc.planes[1][:, :1920], left.planes[1][:, :1920])


def test_cpp_shim_process_frame(N, O, port, glyphs, tmp_path):
    """The reference's process_frame_thread body (encode.cpp:55-98) compiled in C++ against
    include/nes_gpu_shim.hpp: wire bytes -> RenderedFrame -> 4 overlays -> convert_frame()."""
    import subprocess
    from conftest import FONT
    from test_host import build_process_frame
    if N.find_freetype() is None:
        pytest.skip("no FreeType binary in this image")
    exe = build_process_frame(tmp_path)
    w, h = 1280, 720
    rgb, dep = O.synth_rgb(w, h, 9), O.synth_depth(w, h, 9)
    (tmp_path / "m.bin").write_bytes(O.pack_rendered_frame(9, False, w, h, O.KINITIAL_CAMERA_MATRIX, rgb.tobytes(), dep.tobytes()))
    avutil = N.avhandoff.bundled_ffmpeg()["avutil"]
    if avutil is None:
        pytest.skip("no libavutil in this image (to_avframe is part of the flow)")
    r = subprocess.run([exe, str(tmp_path / "m.bin"), FONT, N.find_freetype(), str(w), str(h), "12:34:56.789", str(tmp_path / "o")], capture_output=True, text=True,
                       env=dict(os.environ, NES_AVUTIL_SO=avutil))
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "ok index=9 left=0"
    surf = np.ascontiguousarray(rgb.copy())
    for pos, txt in O.reference_strings(index=9, is_left=False):
        port.render_string(surf, pos, txt, glyphs)
    want_s, want_d = port.rgb_to_yuv420p(surf, "rgb24").cropped(), port.gray_to_yuv420p(dep).cropped()
    assert (tmp_path / "o.scene.yuv").read_bytes() == want_s
    assert (tmp_path / "o.depth.yuv").read_bytes() == want_d
    assert (tmp_path / "o.sws.yuv").read_bytes() == want_s
    assert (tmp_path / "o.zero.yuv").read_bytes() == want_s   # zero-copy constructor (wire bytes borrowed)
    # the same flow with the process-wide multiplexer (nes_shim::enable_mux / NES_GPU_MUX=1): the thread's session hands
    # its frames to the per-GPU dispatcher
    r = subprocess.run([exe, str(tmp_path / "m.bin"), FONT, N.find_freetype(), str(w), str(h), "12:34:56.789", str(tmp_path / "m")], capture_output=True, text=True,
                       env=dict(os.environ, NES_AVUTIL_SO=avutil, NES_GPU_MUX="1"))
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "m.scene.yuv").read_bytes() == want_s and (tmp_path / "m.depth.yuv").read_bytes() == want_d


@pytest.mark.parametrize("w,h,wd,hd,pinned", [(640, 360, 640, 360, True), (322, 94, 322, 94, False), (1920, 1080, 1920, 1080, True), (384, 216, 256, 144, True),
                                              (256, 144, 384, 216, False), (1920, 1080, 1280, 720, True), (64, 8, 64, 8, False)])
def test_depth16(N, O, port, glyphs, session, w, h, wd, hd, pinned):
    """16-bit depth input (GRAY16LE, nes_frame_in.depth_fmt): the depth image equals libswscale's (16-bit scaler + ordered
    dither, pinned in test_oracle.py); the scene of the same frame (with its overlays) is unaffected."""
    rng = np.random.default_rng(w + h)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dep = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    dep[0, :8] = [0, 1, 255, 256, 32767, 32768, 65534, 65535]
    runs = O.reference_strings(index=16)
    want_s, _ = O.expected_frame([(rgb, np.zeros((h, w), np.uint8))], "rgb24", runs, wd, hd, port, glyphs)
    want_d = port.gray16_to_yuv420p(dep, wd, hd)
    scene = N.FrameManager(N.FrameContext(wd, hd, "yuv420p"), session=session if pinned else None)
    depth = N.FrameManager(N.FrameContext(wd, hd, "yuv420p"), session=session if pinned else None)
    if pinned:
        hr, hd_ = session.host_array(rgb.nbytes), session.host_array(dep.nbytes)
        hr[:] = rgb.reshape(-1); hd_[:] = dep.reshape(-1).view(np.uint8)
    else:
        hr, hd_ = np.ascontiguousarray(rgb).reshape(-1), np.ascontiguousarray(dep).reshape(-1).view(np.uint8)
    fin = N.Session.frame_in("rgb24", w, h, [(hr, hd_, 0, 0)], depth_fmt="gray16le")
    session.convert(fin, runs, N.api._frame_out(scene, depth))
    assert depth.cropped() == want_d.cropped(), first_diff(depth.cropped(), want_d.cropped())
    assert scene.cropped() == want_s.cropped(), first_diff(scene.cropped(), want_s.cropped())


def test_depth16_rejects_composites(N, O, session):
    w, h = 64, 32
    srcs = [(O.to_fmt(O.synth_rgb(w, h, k), "rgba").reshape(-1), np.zeros(w * h * 2, np.uint8), 0, 0) for k in range(2)]
    fin = N.Session.frame_in("rgba", w, h, srcs, depth_fmt="gray16le")
    sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p")), N.FrameManager(N.FrameContext(w, h, "yuv420p"))
    with pytest.raises(N.NesGpuError) as e:
        session.convert(fin, None, N.api._frame_out(sc, dp))
    assert e.value.status == N.NES_ERR_INVALID_ARG


def test_mux_many_sessions(N, O, port, glyphs):
    """nes_gpu_mux: 12 client sessions (two pixel formats, with and without text, one of them resizing) driven by 4 host
    threads; every frame equals the oracle, and the dispatcher coalesced frames of several sessions into shared launches."""
    import threading
    mux = N.Mux(device=0, max_batch=32)
    n_sess, frames_per = 12, 10
    sessions, errors = [], []
    try:
        for k in range(n_sess):
            s = N.Session(device=0, max_width=640, max_height=360, max_sources=2, ring_depth=2)
            s.atlas_set(glyphs.metrics, glyphs.bitmaps)
            mux.attach(s)
            sessions.append(s)

        def drive(tid):
            try:
                mine = [k for k in range(n_sess) if k % 4 == tid]
                state = {}
                for k in mine:
                    s = sessions[k]
                    w, h = (640, 360) if k % 3 else (320, 200)
                    wd, hd = (426, 240) if k == 5 else (w, h)
                    fmt = "rgba" if k % 2 else "rgb24"
                    nsrc = 2 if k % 2 else 1
                    state[k] = dict(w=w, h=h, wd=wd, hd=hd, fmt=fmt, nsrc=nsrc, pending=None)
                for f in range(frames_per + 1):
                    for k in mine:
                        st_, s = state[k], sessions[k]
                        if st_["pending"] is not None:  # collect the previous frame of this session
                            t, sc, dp, want, keep = st_["pending"]
                            s.wait(t)
                            assert sc.cropped() == want[0].cropped(), (k, f - 1, first_diff(sc.cropped(), want[0].cropped()))
                            assert dp.cropped() == want[1].cropped(), (k, f - 1)
                            st_["pending"] = None
                        if f == frames_per:
                            continue
                        wl = dict(w=st_["w"], h=st_["h"], fmt=st_["fmt"], n_src=st_["nsrc"])
                        srcs = N.synth.make_sources(wl, 100 * k + f)
                        runs = O.reference_strings(index=f) if (k + f) % 2 else None
                        want = O.expected_frame(srcs, st_["fmt"], runs, st_["wd"], st_["hd"], port, glyphs)
                        sc = N.FrameManager(N.FrameContext(st_["wd"], st_["hd"], "yuv420p"), session=s)
                        dp = N.FrameManager(N.FrameContext(st_["wd"], st_["hd"], "yuv420p"), session=s)
                        keep = [(np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(d).reshape(-1)) for a, d in srcs]
                        fin = N.Session.frame_in(st_["fmt"], st_["w"], st_["h"], [(a, d, 0, 0) for a, d in keep])
                        t = s.submit(fin, runs, N.api._frame_out(sc, dp))
                        st_["pending"] = (t, sc, dp, want, keep)
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        ths = [threading.Thread(target=drive, args=(t,)) for t in range(4)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert not errors, errors[:3]
        st = mux.stats()
        assert st["frames"] == n_sess * frames_per
        assert st["launch_sets"] <= st["frames"] and st["max_batch"] >= 1
        assert mux.error() == ""
    finally:
        for s in sessions:
            s.close()
        mux.close()


def test_mux_destroyed_before_its_sessions(N, O, port, glyphs):
    """Either order of destruction: a session whose mux is gone is back on its own streams and converts on its own;
    a session destroyed first leaves the mux's list (the mux then closes cleanly)."""
    w, h = 320, 200
    mux = N.Mux(device=0, max_batch=8)
    s1 = N.Session(device=0, max_width=w, max_height=h, max_sources=1, ring_depth=2)
    s2 = N.Session(device=0, max_width=w, max_height=h, max_sources=1, ring_depth=2)
    try:
        for s in (s1, s2):
            s.atlas_set(glyphs.metrics, glyphs.bitmaps)
            mux.attach(s)
        wl = dict(w=w, h=h, fmt="rgb24", n_src=1)

        def convert_and_check(s, f):
            srcs = N.synth.make_sources(wl, f)
            runs = O.reference_strings(index=f)
            want = O.expected_frame(srcs, "rgb24", runs, w, h, port, glyphs)
            sc = N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=s)
            dp = N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=s)
            keep = [(np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(d).reshape(-1)) for a, d in srcs]
            fin = N.Session.frame_in("rgb24", w, h, [(a, d, 0, 0) for a, d in keep])
            s.wait(s.submit(fin, runs, N.api._frame_out(sc, dp)))
            assert sc.cropped() == want[0].cropped() and dp.cropped() == want[1].cropped()

        convert_and_check(s1, 1)
        convert_and_check(s2, 2)
        assert mux.stats()["frames"] == 2
        s2.close()           # a session goes first ...
        convert_and_check(s1, 3)
        mux.close()          # ... then the mux, with s1 still attached
        convert_and_check(s1, 4)  # s1 is un-multiplexed now
    finally:
        s1.close()
        s2.close()
        mux.close()


def test_reference_encode_cpp_text_runs(N, O, port, glyphs, tmp_path):
    """The reference's OWN process_frame_thread + send_frame_thread text (extracted from /root/reference in the build
    container into the git-ignored tests/cpp/_ref/, which travels to the GPU box like oracle/_ref) run against the shim on the GPU: one frame goes through the four overlays, convert_frame()
    and the to_avframe() hand-off.  The timestamp overlay is wall-clock text: the planes are compared with the oracle
    outside the rows it can touch; the depth planes entirely."""
    import subprocess
    from conftest import FONT
    import test_host
    if not test_host.reference_text_available():
        pytest.skip("the reference's text was not extracted in the build container (tests/cpp/process_frame.cpp is the committed restatement)")
    if N.find_freetype() is None:
        pytest.skip("no FreeType binary in this image")
    exe = test_host.build_encode_text_driver(tmp_path)
    w, h = 1280, 720
    rgb, dep = O.synth_rgb(w, h, 5), O.synth_depth(w, h, 5)
    (tmp_path / "m.bin").write_bytes(O.pack_rendered_frame(0, True, w, h, O.KINITIAL_CAMERA_MATRIX, rgb.tobytes(), dep.tobytes()))
    avutil = N.avhandoff.bundled_ffmpeg()["avutil"]
    if avutil is None:
        pytest.skip("no libavutil in this image (to_avframe is part of the flow)")
    r = subprocess.run([exe, str(tmp_path / "m.bin"), FONT, N.find_freetype(), str(w), str(h), str(tmp_path / "o")], capture_output=True, text=True,
                       env=dict(os.environ, NES_AVUTIL_SO=avutil))
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().startswith("ok index=0 sent=1+1")
    surf = np.ascontiguousarray(rgb.copy())
    for pos, txt in O.reference_strings(index=0, is_left=True):
        if pos != O.POS_LEFT_TOP:
            port.render_string(surf, pos, txt, glyphs)
    want = port.rgb_to_yuv420p(surf, "rgb24")
    got = np.frombuffer((tmp_path / "o.scene.yuv").read_bytes(), np.uint8)
    gy = got[: w * h].reshape(h, w)
    assert np.array_equal(gy[80:], want.y[80:, :w])                 # below the timestamp line (LEFT_TOP: baseline y = 50)
    assert np.array_equal(gy[:80, 400:], want.y[:80, 400:w])        # right of it
    assert (gy[:80, :400] != want.y[:80, :400]).any()               # and the timestamp was stamped
    assert (tmp_path / "o.depth.yuv").read_bytes() == port.gray_to_yuv420p(dep).cropped()


# ------------------------------------------------------------------ wider coverage of the kernels' paths
def _rgba_sources(O, rng, n, w, h, fmt="rgba"):
    srcs, rgbs, deps = [], [], []
    for k in range(n):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        alpha = np.where(rng.integers(0, 3, (h, w)) == 0, 0, 255).astype(np.uint8)
        img = O.to_fmt(rgb, fmt, alpha)
        dep = np.ascontiguousarray(rng.integers(0, 16, (h, w), dtype=np.uint8) * 16)
        srcs.append((img, dep)); rgbs.append(img); deps.append(dep)
    return srcs, rgbs, deps


@pytest.mark.parametrize("n,w,h,wd,hd", [(5, 512, 64, 512, 64), (8, 320, 48, 320, 48), (6, 384, 216, 256, 144)])
def test_composite_more_than_four_sources(N, O, port, glyphs, n, w, h, wd, hd):
    """5..8 renderer inputs: the fused kernel stages at most 4 sources by TMA, beyond that the
    consumer warps composite into the stage themselves; the resize kernel loops past 4."""
    s = N.Session(device=0, max_width=w, max_height=h, max_sources=8)
    try:
        rng = np.random.default_rng(n)
        srcs, rgbs, deps = _rgba_sources(O, rng, n, w, h)
        comp, cdep = port.composite(rgbs, deps, "rgba")
        sc, dp = run_gpu(N, s, "rgba", srcs, w, h, wd, hd)
        assert sc.cropped() == port.rgb_to_yuv420p(comp, "rgba", wd, hd).cropped()
        assert dp.cropped() == port.gray_to_yuv420p(cdep, wd, hd).cropped()
    finally:
        s.close()


@pytest.mark.parametrize("w,h,wd,hd", [(1536, 384, 256, 64), (960, 540, 320, 180), (128, 96, 640, 480), (720, 486, 1280, 720)])
def test_resize_large_ratios(N, O, port, session, w, h, wd, hd):
    """6:1 and 3:1 reductions (24- and 11-tap filters: smaller destination tiles are picked so that
    the source window fits shared memory) and 5x enlargements."""
    rng = np.random.default_rng(w * 3 + hd)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
    sc, dp = run_gpu(N, session, "rgb24", [(rgb, dep)], w, h, wd, hd)
    want_s, want_d = port.rgb_to_yuv420p(rgb, "rgb24", wd, hd), port.gray_to_yuv420p(dep, wd, hd)
    assert sc.cropped() == want_s.cropped(), first_diff(sc.cropped(), want_s.cropped())
    assert dp.cropped() == want_d.cropped(), first_diff(dp.cropped(), want_d.cropped())


def _device_job(N, session, fmt, srcs, w, h, wd, hd):
    """Upload sources, allocate device planes; returns (frame_in, frame_out, (scene_ptr, depth_ptr), all_ptrs)."""
    bpp = N.PIX_BPP[fmt]
    ptrs, sources = [], []
    for img, dep in srcs:
        d_rgb, d_dep = session.device_alloc(w * h * bpp), session.device_alloc(w * h)
        session.h2d(d_rgb, np.ascontiguousarray(img)); session.h2d(d_dep, np.ascontiguousarray(dep))
        ptrs += [d_rgb, d_dep]
        sources.append(((d_rgb, w * h * bpp), (d_dep, w * h), 0, 0))
    ysz, csz = N.align32(wd) * hd, N.align32(wd // 2) * (hd // 2)
    d_s, d_d = session.device_alloc(ysz + 2 * csz), session.device_alloc(ysz + 2 * csz)
    ptrs += [d_s, d_d]
    fo = N.nes_frame_out(); fo.width, fo.height, fo.mem = wd, hd, N.NES_MEM_DEVICE
    for p, (off, ls) in enumerate([(0, N.align32(wd)), (ysz, N.align32(wd // 2)), (ysz + csz, N.align32(wd // 2))]):
        fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off, ls, d_d + off, ls
    return N.Session.frame_in(fmt, w, h, sources, mem=N.NES_MEM_DEVICE), fo, (d_s, d_d), ptrs


def test_batch_mixed_jobs(N, O, port, glyphs, session):
    """One batched call mixing what a multi-session server submits together: 3-byte and 4-byte
    pixels, 1 / 2 / 4 sources (one sub-stage slot size per launch), same-size and resized outputs
    (composite + overlay + resize is a single fused launch), with and without text."""
    rng = np.random.default_rng(77)
    text = [(O.POS_LEFT_TOP, b"mixed batch\nline two 0123456789"), (O.POS_CENTER, b"centre")]
    specs = [("rgb24", 1, 512, 96, 512, 96, None), ("rgba", 2, 512, 96, 512, 96, text), ("rgba", 1, 768, 64, 768, 64, None),
             ("rgba", 4, 480, 270, 320, 180, text), ("rgb24", 1, 384, 216, 256, 144, text), ("bgra", 3, 256, 48, 256, 48, None)]
    jobs, wants = [], []
    for fmt, n, w, h, wd, hd, runs in specs:
        if n == 1 and N.PIX_BPP[fmt] == 3:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
            srcs, comp, cdep = [(img, dep)], img, dep
        else:
            srcs, rgbs, deps = _rgba_sources(O, rng, n, w, h, fmt)
            # a single source is converted as it is (alpha only matters to the composite)
            comp, cdep = port.composite(rgbs, deps, fmt) if n > 1 else (rgbs[0], deps[0])
        surf = np.ascontiguousarray(comp.copy())
        for pos, txt in runs or []:
            (port.render_string if surf.shape[2] == 3 else (lambda a, p, t, g: port.render_string4(a, p, t, g, fmt)))(surf, pos, txt, glyphs)
        wants.append((port.rgb_to_yuv420p(surf, fmt, wd, hd).cropped(), port.gray_to_yuv420p(np.ascontiguousarray(cdep), wd, hd).cropped()))
        jobs.append(_device_job(N, session, fmt, srcs, w, h, wd, hd) + (wd, hd, runs))
    before = session.launches
    session.convert_batch_device([j[0] for j in jobs], [j[6] for j in jobs], [j[1] for j in jobs], sync=True)
    assert session.launches - before == 4  # two pixel-size classes x (same-size kernel, resize kernel)
    for (fin, fo, (d_s, d_d), ptrs, wd, hd, runs), (want_s, want_d), spec in zip(jobs, wants, specs):
        sc, dp = N.FrameManager(N.FrameContext(wd, hd, "yuv420p")), N.FrameManager(N.FrameContext(wd, hd, "yuv420p"))
        session.d2h(sc.buffer, d_s); session.d2h(dp.buffer, d_d)
        assert sc.cropped() == want_s, (spec[:6], first_diff(sc.cropped(), want_s))
        assert dp.cropped() == want_d, (spec[:6], first_diff(dp.cropped(), want_d))
        for p in ptrs:
            session.device_free(p)


def test_batch_many_units_per_cta(N, O, port, session):
    """24 720p frames in one launch: far more work units than resident CTAs, so every CTA walks
    several (strip, segment) units back to back through the same sub-stage ring and chroma ring."""
    w, h, nf = 1280, 720, 24
    base_rgb, base_dep = O.synth_rgb(w, h, 3), O.synth_depth(w, h, 3)
    jobs = []
    for f in range(nf):
        rgb = np.roll(base_rgb, 37 * f, axis=1); dep = np.roll(base_dep, 11 * f, axis=0)
        jobs.append(_device_job(N, session, "rgb24", [(rgb, dep)], w, h, w, h) + (rgb, dep))
    session.convert_batch_device([j[0] for j in jobs], None, [j[1] for j in jobs], sync=True)
    for f in (0, 7, 23):
        fin, fo, (d_s, d_d), ptrs, rgb, dep = jobs[f]
        sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p")), N.FrameManager(N.FrameContext(w, h, "yuv420p"))
        session.d2h(sc.buffer, d_s); session.d2h(dp.buffer, d_d)
        assert sc.cropped() == port.rgb_to_yuv420p(np.ascontiguousarray(rgb), "rgb24").cropped()
        assert dp.cropped() == port.gray_to_yuv420p(np.ascontiguousarray(dep)).cropped()
    for j in jobs:
        for p in j[3]:
            session.device_free(p)


def test_ingest_ring_zero_copy(N, O, port, glyphs, session):
    """server.cpp:91-112,172-194 replaced end to end: the wire message is received into a pinned
    slot, parsed in place, and the frame is DMA'd from inside the slot (payload offsets inside a
    protobuf message are unaligned) -- no host copy of the pixels."""
    w, h = 640, 360
    ring = N.IngestRing(3, 8 + 64 + w * h * 4)
    try:
        slots = []
        for f in range(3):
            rgb, dep = O.synth_rgb(w, h, f), O.synth_depth(w, h, f)
            msg = O.pack_rendered_frame(f, bool(f & 1), w, h, O.KINITIAL_CAMERA_MATRIX, rgb.tobytes(), dep.tobytes())
            slot, buf = ring.acquire()
            buf[: len(msg)] = np.frombuffer(msg, np.uint8)  # stands for recv(fd, buf, len)
            info, src = ring.commit(slot, len(msg))
            assert (info.index, info.width, info.height, bool(info.is_left)) == (f, w, h, bool(f & 1))
            fin = N.nes_frame_in(); fin.n_sources, fin.pix_fmt, fin.width, fin.height, fin.mem = 1, N.PIX_FMT["rgb24"], w, h, N.NES_MEM_HOST
            fin.src[0] = src
            sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=session), N.FrameManager(N.FrameContext(w, h, "yuv420p"), session=session)
            runs = O.reference_strings(index=f)
            slots.append((slot, session.submit(fin, runs, N.api._frame_out(sc, dp)), sc, dp, rgb, dep, runs, fin))
        with pytest.raises(N.NesGpuError) as e:
            ring.acquire()
        assert e.value.status == N.NES_ERR_BUSY
        for slot, ticket, sc, dp, rgb, dep, runs, _ in slots:
            session.wait(ticket)
            ring.release(slot)
            surf = np.ascontiguousarray(rgb.copy())
            for pos, txt in runs:
                port.render_string(surf, pos, txt, glyphs)
            assert sc.cropped() == port.rgb_to_yuv420p(surf, "rgb24").cropped()
            assert dp.cropped() == port.gray_to_yuv420p(dep).cropped()
        slot, buf = ring.acquire()  # released slots come back
        buf[:16] = 0
        with pytest.raises(N.NesGpuError):
            ring.commit(slot, 16)  # truncated message
    finally:
        ring.close()


def _nv12(yuv) -> bytes:
    """The oracle's planar result re-laid as NV12: same samples, chroma interleaved."""
    y, u, v = yuv.planes()
    uv = np.empty((u.shape[0], 2 * u.shape[1]), np.uint8)
    uv[:, 0::2], uv[:, 1::2] = u, v
    return y.tobytes() + uv.tobytes()


@pytest.mark.parametrize("fmt,n,w,h,wd,hd", [("rgb24", 1, 1280, 720, 1280, 720), ("rgba", 2, 640, 360, 640, 360), ("rgb24", 1, 322, 94, 322, 94),
                                             ("rgb24", 1, 384, 216, 256, 144), ("rgba", 3, 480, 270, 320, 180), ("bgra", 1, 200, 100, 300, 150)])
def test_nv12_output(N, O, port, glyphs, session, fmt, n, w, h, wd, hd):
    """NV12 destination (Y + interleaved UV): the same samples as YUV420P, from both kernels, with
    composite, overlay and depth stream; full-width, ragged-width and unaligned strips."""
    rng = np.random.default_rng(w + n)
    if n == 1:
        img = rng.integers(0, 256, (h, w, N.PIX_BPP[fmt]), dtype=np.uint8)
        dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
        srcs, comp, cdep = [(img, dep)], img, dep
    else:
        srcs, rgbs, deps = _rgba_sources(O, rng, n, w, h, fmt)
        comp, cdep = port.composite(rgbs, deps, fmt)
    runs = O.reference_strings(index=5)
    surf = np.ascontiguousarray(comp.copy())
    for pos, txt in runs:
        (port.render_string if surf.shape[2] == 3 else (lambda a, p, t, g: port.render_string4(a, p, t, g, fmt)))(surf, pos, txt, glyphs)
    want_s, want_d = _nv12(port.rgb_to_yuv420p(surf, fmt, wd, hd)), _nv12(port.gray_to_yuv420p(np.ascontiguousarray(cdep), wd, hd))
    scene, depth = N.FrameManager(N.FrameContext(wd, hd, "nv12"), session=session), N.FrameManager(N.FrameContext(wd, hd, "nv12"))
    sources = [(np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(d).reshape(-1), 0, 0) for a, d in srcs]
    fin = N.Session.frame_in(fmt, w, h, sources)
    session.convert(fin, runs, N.api._frame_out(scene, depth))
    assert scene.cropped() == want_s, first_diff(scene.cropped(), want_s)
    assert depth.cropped() == want_d, first_diff(depth.cropped(), want_d)


def test_side_by_side_stereo_packing(N, O, port, session):
    """Two eyes into one side-by-side frame (BASELINE config 3's layout) in a single launch: each eye
    is a job whose destination planes start half a frame to the right and share the full line size."""
    w, h = 1280, 720
    ysz, csz = N.align32(2 * w) * h, N.align32(w) * (h // 2)
    d_s, d_d = session.device_alloc(ysz + 2 * csz), session.device_alloc(ysz + 2 * csz)
    fins, fouts, ptrs, eyes = [], [], [d_s, d_d], []
    for eye in range(2):
        rgb, dep = O.synth_rgb(w, h, 10 + eye), O.synth_depth(w, h, 10 + eye)
        d_rgb, d_dep = session.device_alloc(w * h * 3), session.device_alloc(w * h)
        session.h2d(d_rgb, rgb); session.h2d(d_dep, dep)
        ptrs += [d_rgb, d_dep]; eyes.append((rgb, dep))
        fins.append(N.Session.frame_in("rgb24", w, h, [((d_rgb, w * h * 3), (d_dep, w * h), 0, 0)], mem=N.NES_MEM_DEVICE))
        fo = N.nes_frame_out(); fo.width, fo.height, fo.mem = w, h, N.NES_MEM_DEVICE
        for p, (off, ls, xoff) in enumerate([(0, N.align32(2 * w), eye * w), (ysz, N.align32(w), eye * w // 2), (ysz + csz, N.align32(w), eye * w // 2)]):
            fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + off + xoff, ls, d_d + off + xoff, ls
        fouts.append(fo)
    before = session.launches
    session.convert_batch_device(fins, None, fouts, sync=True)
    assert session.launches - before == 1
    sbs_s, sbs_d = N.FrameManager(N.FrameContext(2 * w, h, "yuv420p")), N.FrameManager(N.FrameContext(2 * w, h, "yuv420p"))
    session.d2h(sbs_s.buffer, d_s); session.d2h(sbs_d.buffer, d_d)
    for eye, (rgb, dep) in enumerate(eyes):
        ws, wd_ = port.rgb_to_yuv420p(rgb, "rgb24").planes(), port.gray_to_yuv420p(dep).planes()
        for got, want in ((sbs_s, ws), (sbs_d, wd_)):
            assert np.array_equal(got.planes[0][:, eye * w:(eye + 1) * w], want[0])
            assert np.array_equal(got.planes[1][:, eye * w // 2:(eye + 1) * w // 2], want[1])
            assert np.array_equal(got.planes[2][:, eye * w // 2:(eye + 1) * w // 2], want[2])
    for p in ptrs:
        session.device_free(p)


def test_stress_random_sequence(N, O, port, glyphs, session):
    """300 back-to-back frames of random shape / format / source count / text / destination size
    through submit+wait with frames in flight (exercises the self re-arming work counters, the
    sub-stage ring across launches with different slot sizes, tensor-map caching and both kernels);
    every 10th frame is checked against the oracle."""
    rng = np.random.default_rng(2024)
    shapes = [(64, 32), (130, 46), (256, 64), (322, 94), (512, 96), (640, 360), (768, 130), (1280, 720)]
    pending = []
    for it in range(300):
        w, h = shapes[rng.integers(len(shapes))]
        resize = rng.integers(4) == 0
        wd, hd = ((w * 2 // 3) & ~1, (h * 2 // 3) & ~1) if resize else (w, h)
        fmt = ["rgb24", "rgba", "bgra", "argb"][rng.integers(4)]
        n = 1 if fmt == "rgb24" else int(rng.integers(1, 5))
        if n == 1:
            img = rng.integers(0, 256, (h, w, N.PIX_BPP[fmt]), dtype=np.uint8)
            dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
            srcs, comp, cdep = [(img, dep)], img, dep
        else:
            srcs, rgbs, deps = _rgba_sources(O, rng, n, w, h, fmt)
            comp, cdep = (rgbs, deps), None
        runs = O.reference_strings(index=it) if rng.integers(2) else None
        scene, depth = N.FrameManager(N.FrameContext(wd, hd, "yuv420p"), session=session), N.FrameManager(N.FrameContext(wd, hd, "yuv420p"), session=session)
        sources = [(np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(d).reshape(-1), 0, 0) for a, d in srcs]
        fin = N.Session.frame_in(fmt, w, h, sources)
        if len(pending) == 3:
            _finish(N, O, port, glyphs, session, pending.pop(0))
        ticket = session.submit(fin, runs, N.api._frame_out(scene, depth))
        pending.append((ticket, it, fmt, n, comp, cdep, runs, wd, hd, scene, depth, sources, fin))
    while pending:
        _finish(N, O, port, glyphs, session, pending.pop(0))


def _finish(N, O, port, glyphs, session, item):
    ticket, it, fmt, n, comp, cdep, runs, wd, hd, scene, depth, sources, fin = item
    session.wait(ticket)
    if it % 10:
        return
    if n > 1:
        comp, cdep = port.composite(comp[0], comp[1], fmt)
    surf = np.ascontiguousarray(comp.copy())
    for pos, txt in runs or []:
        (port.render_string if surf.shape[2] == 3 else (lambda a, p, t, g: port.render_string4(a, p, t, g, fmt)))(surf, pos, txt, glyphs)
    want_s, want_d = port.rgb_to_yuv420p(surf, fmt, wd, hd).cropped(), port.gray_to_yuv420p(np.ascontiguousarray(cdep), wd, hd).cropped()
    assert scene.cropped() == want_s, (it, fmt, n, wd, hd, first_diff(scene.cropped(), want_s))
    assert depth.cropped() == want_d, (it, fmt, n, wd, hd, first_diff(depth.cropped(), want_d))


def test_two_sessions_two_threads(N, O, port, glyphs):
    """The reference runs one process_frame_thread per eye (main.cpp:274-282): two host threads,
    each with its own session on the same GPU, converting concurrently."""
    import threading
    w, h = 640, 360
    errors = []

    def eye(is_left):
        try:
            s = N.Session(device=0, max_width=w, max_height=h, max_sources=2)
            s.atlas_set(glyphs.metrics, glyphs.bitmaps)
            for f in range(40):
                rgb, dep = O.synth_rgb(w, h, f + 100 * is_left), O.synth_depth(w, h, f + 100 * is_left)
                runs = O.reference_strings(index=f, is_left=bool(is_left))
                sc, dp = run_gpu(N, s, "rgb24", [(rgb, dep)], w, h, runs=runs, pinned=True)
                if f % 8 == 0:
                    surf = np.ascontiguousarray(rgb.copy())
                    for pos, txt in runs:
                        port.render_string(surf, pos, txt, glyphs)
                    assert sc.cropped() == port.rgb_to_yuv420p(surf, "rgb24").cropped()
                    assert dp.cropped() == port.gray_to_yuv420p(dep).cropped()
            s.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ths = [threading.Thread(target=eye, args=(k,)) for k in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("bands", [2, 4, 8])
def test_latency_bands(N, O, port, glyphs, bands):
    """Low-latency mode: the frame goes up, through the kernel and down in row bands; same bytes out,
    `bands` launches per frame; composites, text crossing band borders, NV12, and a frame too small to split."""
    s = N.Session(device=0, max_width=1920, max_height=1080, max_sources=2)
    try:
        s.atlas_set(glyphs.metrics, glyphs.bitmaps)
        s.set_latency_bands(bands)
        rng = np.random.default_rng(bands)
        for fmt, n, w, h, out_fmt in [("rgb24", 1, 1280, 720, "yuv420p"), ("rgba", 2, 1920, 1080, "yuv420p"), ("rgb24", 1, 640, 360, "nv12"), ("rgb24", 1, 64, 32, "yuv420p")]:
            if n == 1:
                img = rng.integers(0, 256, (h, w, N.PIX_BPP[fmt]), dtype=np.uint8)
                dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
                srcs, comp, cdep = [(img, dep)], img, dep
            else:
                srcs, rgbs, deps = _rgba_sources(O, rng, n, w, h, fmt)
                comp, cdep = port.composite(rgbs, deps, fmt)
            text = b"\n".join(bytes((40 + (i * 5 + j) % 80) for i in range(w // 24)) for j in range(h // 22))
            runs = [(O.POS_LEFT_TOP, text)]
            surf = np.ascontiguousarray(comp.copy())
            (port.render_string if surf.shape[2] == 3 else (lambda a, p, t, g: port.render_string4(a, p, t, g, fmt)))(surf, O.POS_LEFT_TOP, text, glyphs)
            ws, wd_ = port.rgb_to_yuv420p(surf, fmt), port.gray_to_yuv420p(np.ascontiguousarray(cdep))
            want_s, want_d = (_nv12(ws), _nv12(wd_)) if out_fmt == "nv12" else (ws.cropped(), wd_.cropped())
            scene, depth = N.FrameManager(N.FrameContext(w, h, out_fmt), session=s), N.FrameManager(N.FrameContext(w, h, out_fmt), session=s)
            sources = []
            for a, d in srcs:
                ha, hd_ = s.host_array(a.nbytes), s.host_array(d.nbytes)
                ha[:] = np.ascontiguousarray(a).reshape(-1); hd_[:] = np.ascontiguousarray(d).reshape(-1)
                sources.append((ha, hd_, 0, 0))
            fin = N.Session.frame_in(fmt, w, h, sources)
            before = s.launches
            s.convert(fin, runs, N.api._frame_out(scene, depth))
            assert scene.cropped() == want_s, (fmt, n, w, h, first_diff(scene.cropped(), want_s))
            assert depth.cropped() == want_d, (fmt, n, w, h, first_diff(depth.cropped(), want_d))
            assert 1 <= s.launches - before <= bands
            if h >= 360:
                assert s.launches - before == bands
    finally:
        s.close()


@pytest.mark.parametrize("fmt,n,w,h", [("rgb24", 1, 272, 36), ("rgb24", 1, 528, 14), ("rgba", 2, 272, 50), ("rgba", 1, 1936, 62), ("rgb24", 1, 4112, 34),
                                       ("bgra", 3, 784, 12), ("rgb24", 1, 16, 16), ("rgba", 4, 48, 132)])
def test_tma_path_ragged_strips_and_short_frames(N, O, port, session, fmt, n, w, h):
    """Widths that are multiples of 16 (rows staged by TMA) but leave a last strip of 16..240 pixels
    (generic store path, zero-filled box columns), frames of 12..16 rows (a single chunk with rows
    above and below the frame), and frames taller than wide."""
    rng = np.random.default_rng(w * 3 + h)
    if n == 1:
        img = rng.integers(0, 256, (h, w, N.PIX_BPP[fmt]), dtype=np.uint8)
        dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
        srcs, comp, cdep = [(img, dep)], img, dep
    else:
        srcs, rgbs, deps = _rgba_sources(O, rng, n, w, h, fmt)
        comp, cdep = port.composite(rgbs, deps, fmt)
    sc, dp = run_gpu(N, session, fmt, srcs, w, h, pinned=True)
    want_s, want_d = port.rgb_to_yuv420p(np.ascontiguousarray(comp), fmt).cropped(), port.gray_to_yuv420p(np.ascontiguousarray(cdep)).cropped()
    assert sc.cropped() == want_s, first_diff(sc.cropped(), want_s)
    assert dp.cropped() == want_d, first_diff(dp.cropped(), want_d)


def test_strided_device_sources(N, O, port, session):
    """Device-resident sources whose row stride is wider than the row (a renderer writing into a padded
    surface): the tensor maps carry the stride; also an odd byte offset (falls back to the non-TMA path)."""
    w, h = 640, 96
    rng = np.random.default_rng(9)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
    want_s, want_d = port.rgb_to_yuv420p(rgb, "rgb24").cropped(), port.gray_to_yuv420p(dep).cropped()
    for rs, ds, off in [(w * 3 + 64, w + 32, 0), (w * 3 + 16, w + 16, 0), (w * 3 + 64, w + 32, 1)]:
        big = np.zeros((h, rs), np.uint8); big[:, : w * 3] = rgb.reshape(h, -1)
        bigd = np.zeros((h, ds), np.uint8); bigd[:, :w] = dep
        d_rgb, d_dep = session.device_alloc(big.nbytes + 16), session.device_alloc(bigd.nbytes + 16)
        session.h2d(d_rgb + off, big); session.h2d(d_dep + off, bigd)
        ysz, csz = N.align32(w) * h, N.align32(w // 2) * (h // 2)
        d_s, d_d = session.device_alloc(ysz + 2 * csz), session.device_alloc(ysz + 2 * csz)
        fin = N.Session.frame_in("rgb24", w, h, [((d_rgb + off, big.nbytes), (d_dep + off, bigd.nbytes), rs, ds)], mem=N.NES_MEM_DEVICE)
        fo = N.nes_frame_out(); fo.width, fo.height, fo.mem = w, h, N.NES_MEM_DEVICE
        for p, (o, ls) in enumerate([(0, N.align32(w)), (ysz, N.align32(w // 2)), (ysz + csz, N.align32(w // 2))]):
            fo.scene[p], fo.scene_linesize[p], fo.depth[p], fo.depth_linesize[p] = d_s + o, ls, d_d + o, ls
        session.convert_batch_device([fin], None, [fo], sync=True)
        sc, dp = N.FrameManager(N.FrameContext(w, h, "yuv420p")), N.FrameManager(N.FrameContext(w, h, "yuv420p"))
        session.d2h(sc.buffer, d_s); session.d2h(dp.buffer, d_d)
        assert sc.cropped() == want_s, (rs, ds, off, first_diff(sc.cropped(), want_s))
        assert dp.cropped() == want_d, (rs, ds, off)
        for p in (d_rgb, d_dep, d_s, d_d):
            session.device_free(p)
