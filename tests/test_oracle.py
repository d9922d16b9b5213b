"""CPU tests (no GPU): pin the oracle to the golden vectors / the real libraries, and check
the product's host-side pieces (filter tables, unpack, text layout, ABI exports)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import FONT, sha16


# ---------------------------------------------------------------- oracle vs golden vectors
def test_port_matches_golden_convert(O, port, golden):
    for c in golden["convert"]:
        (w, h), (wd, hd) = c["src"], c["dst"]
        if w * h > 1920 * 1080:
            continue  # 4K cases run in test_port_matches_golden_4k
        s = port.rgb_to_yuv420p(O.synth_rgb(w, h), "rgb24", wd, hd)
        d = port.gray_to_yuv420p(O.synth_depth(w, h), wd, hd)
        assert sha16(s.cropped()) == c["scene"], c
        assert sha16(d.cropped()) == c["depth"], c
        assert s.y[0, :4].tolist() == c["y0"] and s.u[0, :4].tolist() == c["u0"] and s.v[0, :4].tolist() == c["v0"]


def test_port_matches_golden_4k(O, port, golden):
    for c in golden["convert"]:
        (w, h), (wd, hd) = c["src"], c["dst"]
        if w * h <= 1920 * 1080:
            continue
        s = port.rgb_to_yuv420p(O.synth_rgb(w, h), "rgb24", wd, hd)
        assert sha16(s.cropped()) == c["scene"], c


def test_port_pixel_formats(O, port, golden):
    rgb = O.synth_rgb(320, 180)
    for c in golden["formats"]:
        s = port.rgb_to_yuv420p(O.to_fmt(rgb, c["fmt"]), c["fmt"], *c["dst"])
        assert sha16(s.cropped()) == c["scene"], c


def test_port_gray_lut_and_constant(O, port, golden):
    lut = port.gray_to_yuv420p(np.tile(np.arange(256, dtype=np.uint8), (4, 1)))
    assert lut.y[0, :256].tolist() == golden["gray_lut"]
    assert sha16(lut.y[0, :256].tobytes()) == golden["gray_lut_sha16"] == "a73ff4dd73bef9ae"
    assert (lut.u[:, :128] == 128).all() and (lut.v[:, :128] == 128).all()
    c = port.rgb_to_yuv420p(np.full((16, 16, 3), 200, np.uint8), "rgb24")
    assert [int(c.y[0, 0]), int(c.u[0, 0]), int(c.v[0, 0])] == golden["const200"] == [188, 128, 128]


def test_port_overlay_golden(O, port, glyphs, golden):
    for c in golden["overlay"]:
        w, h = c["size"]
        surf = np.ascontiguousarray(O.synth_rgb(w, h))
        n = sum(port.render_string(surf, pos, txt, glyphs) for pos, txt in O.reference_strings())
        assert n == c["stamp_calls"]
        assert int((surf != O.synth_rgb(w, h)).any(axis=2).sum()) == c["stamped"]
        assert sha16(surf.tobytes()) == c["rgb"]
        assert sha16(port.rgb_to_yuv420p(surf, "rgb24").cropped()) == c["yuv"]


def test_survey_filter_tables(port):
    """SURVEY.md Appendix A.2/A.3 reference tables."""
    c, p = port.init_filter(3840, 2560, 1 << 14)
    assert c.shape[1] == 6
    assert c[1000].tolist() == [-819, 1567, 10266, 6280, -758, -152] and p[1000] == 1498
    assert c[1001].tolist() == [-152, -758, 6280, 10266, 1567, -819] and p[1001] == 1499
    assert c[0].tolist() == [11014, 6280, -758, -152, 0, 0] and p[0] == 0
    assert c[1].tolist() == [-752, 6223, 10171, 1554, -812, 0] and p[1] == 0
    assert c[2558].tolist() == [0, -819, 1567, 10266, 6280, -910] and p[2558] == 3834
    assert c[2559].tolist() == [0, 0, -152, -758, 6280, 11014] and p[2559] == 3834
    c, p = port.init_filter(2160, 1440, 1 << 12)
    assert c[700].tolist() == [-205, 392, 2566, 1571, -190, -38] and p[700] == 1048
    assert c[0].tolist() == [2753, 1571, -190, -38, 0, 0] and c[1439].tolist() == [0, 0, -38, -190, 1571, 2753] and p[1439] == 2154
    c, p = port.init_filter(2160, 720, 1 << 12)
    assert c.shape[1] == 11 and c[100].tolist() == [-61, -121, 0, 475, 1072, 1366, 1072, 475, 0, -121, -61] and p[100] == 296
    assert c[0].tolist() == [1365, 1366, 1072, 475, 0, -121, -61, 0, 0, 0, 0]
    assert c[719].tolist() == [0, 0, 0, 0, -61, -121, 0, 475, 1072, 1366, 1365] and p[719] == 2149
    c, p = port.init_filter(2160, 1080, 1 << 12)  # the same-size chroma vertical filter
    assert c[500].tolist() == [-58, -172, 492, 1786, 1786, 492, -172, -58] and p[500] == 997
    assert c[0].tolist() == [2048, 1786, 492, -172, -58, 0, 0, 0] and c[1].tolist() == [-230, 492, 1786, 1786, 492, -172, -58, 0]


# ---------------------------------------------------------------- oracle vs the real libraries
def test_port_vs_real_swscale_random(O, port, ref):
    rng = np.random.default_rng(1)
    for (w, h, wd, hd) in [(64, 32, 64, 32), (130, 46, 130, 46), (258, 70, 258, 70), (96, 54, 64, 36), (128, 72, 192, 108), (200, 100, 120, 90), (322, 182, 214, 120), (160, 90, 480, 270), (640, 360, 212, 120)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        dep = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert port.rgb_to_yuv420p(rgb, "rgb24", wd, hd).cropped() == ref.sws_convert(rgb, "rgb24", wd, hd).cropped(), (w, h, wd, hd)
        assert port.gray_to_yuv420p(dep, wd, hd).cropped() == ref.sws_convert(dep, "gray", wd, hd).cropped(), (w, h, wd, hd)
    rgba = rng.integers(0, 256, (90, 160, 4), dtype=np.uint8)
    for fmt in ["rgba", "bgra", "argb", "abgr"]:
        assert port.rgb_to_yuv420p(rgba, fmt).cropped() == ref.sws_convert(rgba, fmt).cropped(), fmt
    # extremes (saturated primaries exercise the 15-bit clamps)
    ext = np.zeros((32, 64, 3), np.uint8)
    ext[:, :16] = (255, 0, 0); ext[:, 16:32] = (0, 0, 255); ext[:, 32:48] = (0, 255, 0); ext[::2, 48:] = 255
    assert port.rgb_to_yuv420p(ext, "rgb24").cropped() == ref.sws_convert(ext, "rgb24").cropped()


def test_port_overlay_vs_real_freetype(O, port, ref, glyphs):
    t = ref.text_new(FONT)
    try:
        live = ref.glyph_table(t)
        assert np.array_equal(live.metrics, glyphs.metrics)
        assert all(np.array_equal(a, b) for a, b in zip(live.bitmaps, glyphs.bitmaps))
        for (w, h) in [(640, 360), (200, 120), (1280, 720)]:
            a = np.ascontiguousarray(O.synth_rgb(w, h)); b = a.copy()
            for pos, txt in O.reference_strings(index=123456) + [(O.POS_RIGHT_BOTTOM, b"clipped at the right edge \xe9\xff~")]:
                assert ref.text_render(t, a, pos, txt) == port.render_string(b, pos, txt, glyphs)
            assert np.array_equal(a, b), (w, h)
    finally:
        ref.text_free(t)


# ---------------------------------------------------------------- product host logic vs oracle
def test_product_filter_tables_match_oracle(N, port):
    for (s, d, one) in [(3840, 2560, 1 << 14), (1920, 1280, 1 << 14), (2160, 1440, 1 << 12), (2160, 720, 1 << 12), (2160, 1080, 1 << 12),
                        (128, 192, 1 << 14), (96, 64, 1 << 14), (200, 120, 1 << 14), (100, 90, 1 << 12), (54, 54, 1 << 12), (640, 426, 1 << 14),
                        (360, 240, 1 << 12), (160, 480, 1 << 14), (90, 270, 1 << 12), (1280, 1920, 1 << 14), (720, 1080, 1 << 12), (640, 212, 1 << 14),
                        (6, 4, 1 << 14), (4, 8, 1 << 12), (1000, 100, 1 << 14)]:
        c, p = N.filter_table(s, d, one)
        c2, p2 = port.init_filter(s, d, one)
        assert c.shape == c2.shape and np.array_equal(c, c2) and np.array_equal(p, p2), (s, d, one)
        assert (c.astype(np.int64).sum(axis=1) == one).all()


def test_product_unpack_matches_oracle(N, O):
    rng = np.random.default_rng(7)
    frame, depth = rng.integers(0, 256, 64 * 32 * 3, dtype=np.uint8).tobytes(), rng.integers(0, 256, 64 * 32, dtype=np.uint8).tobytes()
    orders = [("index", "camera", "is_left", "frame", "depth"), ("depth", "frame", "is_left", "camera", "index"), ("camera", "depth", "index", "frame", "is_left")]
    for order in orders:
        for (idx, left) in [(0, False), (5, True), (2**40 + 3, True)]:
            msg = O.pack_rendered_frame(idx, left, 64, 32, O.KINITIAL_CAMERA_MATRIX, frame, depth, field_order=order)
            a, b = N.unpack_rendered_frame(msg), O.unpack_rendered_frame(msg)
            for k in ("index", "is_left", "cam_is_left", "width", "height", "matrix", "frame", "depth"):
                assert a[k] == b[k], (k, order)
            fo, fl = a["frame"]
            assert msg[fo:fo + fl] == frame
            do, dl = a["depth"]
            assert msg[do:do + dl] == depth
            assert a["consumed"] == len(msg)
    # no prefix, empty payloads, unknown fields, truncation
    msg = O.pack_rendered_frame(1, False, 0, 0, [], b"", b"", prefix=False)
    assert N.unpack_rendered_frame(msg, prefix=False)["frame"] == (0, 0)
    good = O.pack_rendered_frame(9, True, 64, 32, [1.0] * 12, frame, depth)
    extra = good[:8] + b"\x78\x05" + b"\xa2\x06\x03abc" + good[8:]  # unknown varint field 15, unknown bytes field 100
    extra = (len(extra) - 8).to_bytes(8, "little") + extra[8:]
    assert N.unpack_rendered_frame(extra)["index"] == 9
    for cut in (4, 9, 40, len(good) - 1):
        with pytest.raises(N.NesGpuError) as e:
            N.unpack_rendered_frame(good[:cut])
        assert e.value.status == N.NES_ERR_PARSE
    bad = good[:8] + b"\x0b" + good[9:]  # wire type 3 (group) is not part of nes.proto
    with pytest.raises(N.NesGpuError):
        N.unpack_rendered_frame(bad)


def _nesproto():
    """nesproto.* message classes from the real protobuf runtime: a dynamic FileDescriptorProto that
    restates /root/reference/proto/nes.proto:1-25 (no protoc in this image)."""
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    T = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="nes.proto", package="nesproto", syntax="proto3")
    cam = fd.message_type.add(name="Camera")
    cam.field.add(name="is_left", number=2, type=T.TYPE_BOOL, label=T.LABEL_OPTIONAL)
    cam.field.add(name="width", number=3, type=T.TYPE_UINT32, label=T.LABEL_OPTIONAL)
    cam.field.add(name="height", number=4, type=T.TYPE_UINT32, label=T.LABEL_OPTIONAL)
    cam.field.add(name="matrix", number=12, type=T.TYPE_FLOAT, label=T.LABEL_REPEATED)
    rf = fd.message_type.add(name="RenderedFrame")
    rf.field.add(name="index", number=1, type=T.TYPE_UINT64, label=T.LABEL_OPTIONAL)
    rf.field.add(name="camera", number=2, type=T.TYPE_MESSAGE, type_name=".nesproto.Camera", label=T.LABEL_OPTIONAL)
    rf.field.add(name="is_left", number=3, type=T.TYPE_BOOL, label=T.LABEL_OPTIONAL)
    rf.field.add(name="frame", number=6, type=T.TYPE_BYTES, label=T.LABEL_OPTIONAL)
    rf.field.add(name="depth", number=7, type=T.TYPE_BYTES, label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        f = message_factory.MessageFactory(pool)
        return f.GetPrototype(pool.FindMessageTypeByName("nesproto.Camera")), f.GetPrototype(pool.FindMessageTypeByName("nesproto.RenderedFrame")), pb.__version__
    return get(pool.FindMessageTypeByName("nesproto.Camera")), get(pool.FindMessageTypeByName("nesproto.RenderedFrame")), pb.__version__


def test_product_unpack_of_real_protobuf_messages(N, O):
    """nes_unpack_rendered_frame against messages serialised by the REAL protobuf runtime (what the renderer
    sends, server.cpp:172-175), not by the oracle's own packer: proto3 defaults (zero scalars omitted), packed
    repeated float, fields in shuffled order (concatenated partial messages merge on parse), last-one-wins."""
    Camera, RenderedFrame, version = _nesproto()
    rng = np.random.default_rng(42)
    w, h = 96, 54
    frame, depth = rng.integers(0, 256, w * h * 3, dtype=np.uint8).tobytes(), rng.integers(0, 256, w * h, dtype=np.uint8).tobytes()
    cases = [dict(index=0, is_left=False, cam_left=False, w=w, h=h, matrix=O.KINITIAL_CAMERA_MATRIX),
             dict(index=7, is_left=True, cam_left=True, w=w, h=h, matrix=[0.0] * 12),
             dict(index=2**63 + 11, is_left=True, cam_left=False, w=w, h=h, matrix=[float(i) - 5.5 for i in range(12)]),
             dict(index=300, is_left=False, cam_left=True, w=0, h=0, matrix=[])]
    for c in cases:
        m = RenderedFrame(index=c["index"], is_left=c["is_left"], frame=frame, depth=depth)
        m.camera.is_left, m.camera.width, m.camera.height = c["cam_left"], c["w"], c["h"]
        m.camera.matrix.extend(c["matrix"])
        wire = m.SerializeToString()
        # round trip through the real parser first: what we compare against is what protobuf itself reads back
        back = RenderedFrame.FromString(wire)
        for prefix in (False, True):
            msg = (len(wire).to_bytes(8, "little") + wire) if prefix else wire
            u = N.unpack_rendered_frame(msg, prefix=prefix)
            assert u["index"] == back.index and u["is_left"] == back.is_left and u["cam_is_left"] == back.camera.is_left
            assert u["width"] == back.camera.width and u["height"] == back.camera.height
            assert u["matrix"] == pytest.approx(list(back.camera.matrix), abs=0) and len(u["matrix"]) == len(back.camera.matrix)
            fo, fl = u["frame"]
            do, dl = u["depth"]
            assert msg[fo:fo + fl] == back.frame and msg[do:do + dl] == back.depth
            assert u["consumed"] == len(msg)
            o = O.unpack_rendered_frame(msg, prefix=prefix)
            for k in ("index", "is_left", "cam_is_left", "width", "height", "frame", "depth"):
                assert u[k] == o[k], k
    # shuffled field order: serialise single-field messages and concatenate them (protobuf merges on parse);
    # the camera is split in two partial sub-messages, the index appears twice (last wins)
    parts = {
        "index_old": RenderedFrame(index=1).SerializeToString(),
        "depth": RenderedFrame(depth=depth).SerializeToString(),
        "cam_a": RenderedFrame(camera=Camera(width=w, matrix=[1.0, 2.0, 3.0])).SerializeToString(),
        "frame": RenderedFrame(frame=frame).SerializeToString(),
        "is_left": RenderedFrame(is_left=True).SerializeToString(),
        "cam_b": RenderedFrame(camera=Camera(height=h, is_left=True, matrix=[4.0])).SerializeToString(),
        "index": RenderedFrame(index=99).SerializeToString(),
    }
    for seed in range(6):
        order = list(parts)
        np.random.default_rng(seed).shuffle(order)
        if order.index("index_old") > order.index("index"):
            i, j = order.index("index_old"), order.index("index")
            order[i], order[j] = order[j], order[i]
        wire = b"".join(parts[k] for k in order)
        back = RenderedFrame.FromString(wire)
        u = N.unpack_rendered_frame(wire, prefix=False)
        assert (u["index"], u["is_left"], u["cam_is_left"], u["width"], u["height"]) == (99, True, True, w, h) == (back.index, back.is_left, back.camera.is_left, back.camera.width, back.camera.height)
        assert u["matrix"] == list(back.camera.matrix), order
        fo, fl = u["frame"]
        do, dl = u["depth"]
        assert wire[fo:fo + fl] == frame and wire[do:do + dl] == depth
    # unpacked (non-packed) repeated float is legal wire too: tag 0x65 (field 12, fixed32) per element
    import struct
    cam = b"\x18" + bytes([w]) + b"".join(b"\x65" + struct.pack("<f", v) for v in (1.5, -2.5))
    wire = b"\x12" + bytes([len(cam)]) + cam
    assert list(RenderedFrame.FromString(wire).camera.matrix) == [1.5, -2.5]
    assert N.unpack_rendered_frame(wire, prefix=False)["matrix"] == [1.5, -2.5]


def test_port_matches_config_size_goldens(N, O, port, glyphs, golden):
    """The C port (text from the golden glyph table) reproduces, at the BASELINE configs' full sizes, the hashes
    tests/golden/make_golden.py took from the real libswscale / FreeType: 1080p 2-source composite + overlays,
    7680x2160 with one overlay set per eye, 1080p sessions, 4 x 4K composite + dense overlay -> 1440p, 4K."""
    assert set(golden["configs"]) == set(N.synth.WORKLOADS)
    for name, c in golden["configs"].items():
        wl = N.synth.WORKLOADS[name]
        sc, dp = O.expected_frame(N.synth.make_sources(wl, c["frame"]), wl["fmt"], N.synth.text_runs(wl, c["frame"]), wl["wd"], wl["hd"], port, glyphs)
        assert sha16(sc.cropped()) == c["scene"], name
        assert sha16(dp.cropped()) == c["depth"], name


def test_port_gray16_vs_real_libswscale(O, port, ref):
    """GRAY16LE -> YUV420P (16-bit depth input, SURVEY.md §8 f rank 4): the port equals the real libswscale, same size
    and resized; and the 8x8 ordered dither libswscale applies to a >8-bit source is re-derived entry by entry from the
    live library (same-size output = (p + d[y&7][x&7]) >> 7 with p the range-compressed 15-bit sample)."""
    import ctypes as C
    rng = np.random.default_rng(16)
    for (w, h, wd, hd) in [(64, 32, 64, 32), (320, 180, 320, 180), (384, 216, 256, 144), (256, 144, 384, 216), (200, 100, 120, 90), (96, 54, 64, 54)]:
        img = rng.integers(0, 65536, (h, w), dtype=np.uint16)
        assert port.gray16_to_yuv420p(img, wd, hd).cropped() == ref.sws_convert(img, "gray16le", wd, hd).cropped(), (w, h, wd, hd)
    full = np.arange(65536, dtype=np.uint16).reshape(16, 4096)
    assert port.gray16_to_yuv420p(full).cropped() == ref.sws_convert(full, "gray16le").cropped()
    img = rng.integers(0, 65536, (64, 4096), dtype=np.uint16)
    out = ref.sws_convert(img, "gray16le").y[:64, :4096].astype(np.int64)
    p = ((np.minimum(img.astype(np.int64) >> 1, 32767)) * 14071 + 33561472) >> 14
    port.L.nes_oracle_dither8x8_128.restype = C.POINTER(C.c_uint8)
    table = np.ctypeslib.as_array(port.L.nes_oracle_dither8x8_128(), shape=(8, 8))
    for yy in range(8):
        for xx in range(8):
            pp, oo = p[yy::8, xx::8].ravel(), out[yy::8, xx::8].ravel()
            m = (oo > 0) & (oo < 255)
            lo, hi = (128 * oo - pp)[m].max(), (128 * oo - pp + 127)[m].min()
            assert lo == hi == table[yy, xx], (yy, xx, lo, hi, table[yy, xx])


def test_format_camera_matrix(N, O):
    assert N.format_camera_matrix(O.KINITIAL_CAMERA_MATRIX) == O.format_matrix_text(O.KINITIAL_CAMERA_MATRIX)
    assert N.format_camera_matrix(O.KINITIAL_CAMERA_MATRIX).startswith(b"+1.00000 +0.00000 +0.00000 +0.50000 \n+0.00000 -1.00000")


def test_abi_exports_every_declared_symbol(N):
    L = ctypes.CDLL(N.LIB_PATH)
    hdr = open(os.path.join(os.path.dirname(N.LIB_PATH), "..", "include", "nes_gpu.h")).read()
    import re
    declared = set(re.findall(r"NES_API\s+[\w\s\*]+?\b(nes_\w+)\s*\(", hdr))
    assert declared == set(N.ABI_SYMBOLS), declared ^ set(N.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.nes_gpu_abi_version() == 2


def test_ctypes_mirrors_match_the_header(N, tmp_path):
    """Compile a C program against include/nes_gpu.h and compare struct sizes with the ctypes mirrors."""
    import subprocess
    names = ["nes_gpu_cfg", "nes_glyph", "nes_text_run", "nes_source", "nes_frame_in", "nes_frame_out", "nes_timing", "nes_placed_glyph", "nes_unpacked_frame"]
    src = '#include <stdio.h>\n#include "nes_gpu.h"\nint main(void){' + "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}\n"
    c = tmp_path / "sz.c"
    c.write_text(src)
    inc = os.path.join(os.path.dirname(N.LIB_PATH), "..", "include")
    subprocess.run(["gcc", "-std=c11", "-I", inc, str(c), "-o", str(tmp_path / "sz")], check=True)
    out = subprocess.run([str(tmp_path / "sz")], check=True, capture_output=True, text=True).stdout.split()
    for n, sz in zip(names, out):
        assert ctypes.sizeof(getattr(N, n)) == int(sz), n


def test_no_cpu_fallback(N):
    """Without a CUDA device the compute entry points must fail loudly, never fall back."""
    if N.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(N.NesGpuError) as e:
        N.Session()
    assert e.value.status == N.NES_ERR_CUDA
    assert N.strerror(N.NES_ERR_CUDA).startswith("CUDA error")


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ngp-encode-server_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt and "libnes_ref" not in txt, f


def test_product_font_rasterise_matches_golden_glyphs(N, glyphs):
    """The product's FreeType loader (csrc/text.cc) yields the same 256 glyphs as the golden table
    (made with the oracle's shim over the same FreeType 2.14.3)."""
    if N.find_freetype() is None:
        pytest.skip("no FreeType binary in this image")
    metrics, bitmaps = N.font_rasterise(FONT)
    for b in range(256):
        if glyphs.metrics[b][0] * glyphs.metrics[b][1] == 0:
            assert bitmaps[b].size == 0 and metrics[b][4] == glyphs.metrics[b][4], b
        else:
            assert metrics[b].tolist() == glyphs.metrics[b].tolist(), b
            assert np.array_equal(bitmaps[b], glyphs.bitmaps[b]), b
    with pytest.raises(N.NesGpuError) as e:
        N.font_rasterise("/nonexistent/font.ttf")
    assert e.value.status == N.NES_ERR_FREETYPE
