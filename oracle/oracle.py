"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of the CPU checker:

* ``Port``  -- oracle/liboracle_port.so, the plain-C restatement of the arithmetic
  on the reference's hot path (swscale_port.c, overlay_port.c; each function cites
  the reference file:line it follows).
* ``Ref``   -- oracle/_ref/libnes_ref.so, the reference's glue
  (type_managers.cc:143-155, render_text.cc:10-111) restated against the REAL
  libswscale 9.1.100 / FreeType 2.14.3 binaries bundled in this image
  (opencv_python_headless.libs / pillow.libs).  Used to pin ``Port``.
* ``unpack_rendered_frame`` -- python restatement of the wire format the
  reference parses in src/server.cpp:91-112,175 (8-byte size prefix +
  nesproto.RenderedFrame, proto/nes.proto:18-25).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import struct
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libnes_ref.so")

SWS_BICUBIC = 4
SWS_ACCURATE_RND = 0x40000
SWS_BITEXACT = 0x80000
PARITY_FLAGS = SWS_BITEXACT | SWS_ACCURATE_RND

POS_LEFT_TOP, POS_LEFT_BOTTOM, POS_RIGHT_TOP, POS_RIGHT_BOTTOM, POS_CENTER = range(5)

# pixel format name -> (bytes per pixel, r_off, g_off, b_off, a_off)
PIXFMT = {
    "rgb24": (3, 0, 1, 2, -1),
    "bgr24": (3, 2, 1, 0, -1),
    "rgba": (4, 0, 1, 2, 3),
    "bgra": (4, 2, 1, 0, 3),
    "argb": (4, 1, 2, 3, 0),
    "abgr": (4, 3, 2, 1, 0),
}


def build(force: bool = False) -> None:
    """Compile the checker with the committed recipe (oracle/Makefile)."""
    if force:
        subprocess.run(["make", "-C", HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def _u8p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def align32(x: int) -> int:
    return (x + 31) & ~31


class Yuv:
    """Three planes laid out like the reference's av_image_alloc(..., align=32) block
    (type_managers.cc:119-121): one buffer, Y then U then V, strides aligned to 32."""

    def __init__(self, w: int, h: int):
        self.w, self.h = w, h
        self.cw, self.ch = (w + 1) // 2, (h + 1) // 2
        self.ys, self.cs = align32(w), align32(self.cw)
        self.buf = np.zeros(self.ys * h + 2 * self.cs * self.ch, np.uint8)
        self.y = self.buf[: self.ys * h].reshape(h, self.ys)
        o = self.ys * h
        self.u = self.buf[o : o + self.cs * self.ch].reshape(self.ch, self.cs)
        o += self.cs * self.ch
        self.v = self.buf[o : o + self.cs * self.ch].reshape(self.ch, self.cs)

    def cropped(self) -> bytes:
        return (
            self.y[:, : self.w].tobytes() + self.u[:, : self.cw].tobytes() + self.v[:, : self.cw].tobytes()
        )

    def planes(self):
        return self.y[:, : self.w], self.u[:, : self.cw], self.v[:, : self.cw]


class _Glyph(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("rows", C.c_int32),
        ("left", C.c_int32),
        ("top", C.c_int32),
        ("advance", C.c_int32),
        ("_pad", C.c_int32),
        ("buffer", C.POINTER(C.c_uint8)),
    ]


class GlyphTable:
    """256 glyphs indexed by byte value, as FT_Load_Char(face, (char)byte, FT_LOAD_RENDER)
    leaves them (render_text.cc:88).  metrics[b] = (width, rows, left, top, advance)."""

    def __init__(self, metrics: np.ndarray, bitmaps: list):
        self.metrics = metrics  # int32 [256,5]
        self.bitmaps = bitmaps  # list of uint8 arrays (rows*width)
        self._arr = (_Glyph * 256)()
        for b in range(256):
            g = self._arr[b]
            g.width, g.rows, g.left, g.top, g.advance = (int(v) for v in metrics[b])
            bm = bitmaps[b]
            g.buffer = _u8p(bm) if bm.size else None

    def save(self, path: str) -> None:
        offs = np.zeros(257, np.int64)
        for b in range(256):
            offs[b + 1] = offs[b] + self.bitmaps[b].size
        cov = np.concatenate([bm.ravel() for bm in self.bitmaps]) if offs[-1] else np.zeros(0, np.uint8)
        np.savez_compressed(path, metrics=self.metrics, offsets=offs, coverage=cov)

    @staticmethod
    def load(path: str) -> "GlyphTable":
        z = np.load(path)
        offs = z["offsets"]
        cov = z["coverage"]
        bitmaps = [np.ascontiguousarray(cov[offs[b] : offs[b + 1]]) for b in range(256)]
        return GlyphTable(z["metrics"].astype(np.int32), bitmaps)


class Port:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        self.L = L
        L.nes_oracle_init_filter.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.POINTER(C.c_int32))]
        L.nes_oracle_init_filter.restype = C.c_int
        L.nes_oracle_free.argtypes = [C.c_void_p]
        L.nes_oracle_rgb_to_yuv420p.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_uint8), C.c_int] * 3
        L.nes_oracle_rgb_to_yuv420p.restype = C.c_int
        L.nes_oracle_gray_to_yuv420p.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_uint8), C.c_int] * 3
        L.nes_oracle_gray_to_yuv420p.restype = C.c_int
        L.nes_oracle_gray16_to_yuv420p.argtypes = [C.POINTER(C.c_uint16), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_uint8), C.c_int] * 3
        L.nes_oracle_gray16_to_yuv420p.restype = C.c_int
        L.nes_oracle_render_string.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.POINTER(_Glyph)]
        L.nes_oracle_render_string.restype = C.c_long
        L.nes_oracle_render_string4.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.POINTER(_Glyph), C.c_int]
        L.nes_oracle_render_string4.restype = C.c_long
        L.nes_oracle_composite.argtypes = [C.c_int, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int]
        L.nes_oracle_composite.restype = None

    def init_filter(self, src: int, dst: int, one: int):
        """-> (coef int16 [dst,size], pos int32 [dst])"""
        pc = C.POINTER(C.c_int16)()
        pp = C.POINTER(C.c_int32)()
        size = self.L.nes_oracle_init_filter(src, dst, one, C.byref(pc), C.byref(pp))
        coef = np.ctypeslib.as_array(pc, shape=(dst, size)).copy()
        pos = np.ctypeslib.as_array(pp, shape=(dst,)).copy()
        self.L.nes_oracle_free(pc)
        self.L.nes_oracle_free(pp)
        return coef, pos

    def rgb_to_yuv420p(self, img: np.ndarray, fmt: str, wd: int | None = None, hd: int | None = None) -> Yuv:
        """img: uint8 [H, W, bpp] (C contiguous rows; row stride = img.strides[0])"""
        bpp, ro, go, bo, _ = PIXFMT[fmt]
        h, w = img.shape[:2]
        assert img.shape[2] == bpp and img.dtype == np.uint8
        wd, hd = wd or w, hd or h
        out = Yuv(wd, hd)
        r = self.L.nes_oracle_rgb_to_yuv420p(_u8p(img), img.strides[0], bpp, ro, go, bo, w, h, wd, hd, _u8p(out.y), out.ys, _u8p(out.u), out.cs, _u8p(out.v), out.cs)
        if r:
            raise ValueError("nes_oracle_rgb_to_yuv420p: bad arguments")
        return out

    def gray_to_yuv420p(self, img: np.ndarray, wd: int | None = None, hd: int | None = None) -> Yuv:
        h, w = img.shape
        wd, hd = wd or w, hd or h
        out = Yuv(wd, hd)
        r = self.L.nes_oracle_gray_to_yuv420p(_u8p(img), img.strides[0], w, h, wd, hd, _u8p(out.y), out.ys, _u8p(out.u), out.cs, _u8p(out.v), out.cs)
        if r:
            raise ValueError("nes_oracle_gray_to_yuv420p: bad arguments")
        return out

    def gray16_to_yuv420p(self, img: np.ndarray, wd: int | None = None, hd: int | None = None) -> Yuv:
        """img: uint16 [H, W] (GRAY16LE)"""
        h, w = img.shape
        assert img.dtype == np.uint16
        wd, hd = wd or w, hd or h
        out = Yuv(wd, hd)
        r = self.L.nes_oracle_gray16_to_yuv420p(img.ctypes.data_as(C.POINTER(C.c_uint16)), img.strides[0], w, h, wd, hd, _u8p(out.y), out.ys, _u8p(out.u), out.cs, _u8p(out.v), out.cs)
        if r:
            raise ValueError("nes_oracle_gray16_to_yuv420p: bad arguments")
        return out

    def render_string(self, surface: np.ndarray, position: int, text: bytes, glyphs: GlyphTable) -> int:
        """surface: uint8 [H, W, 3] contiguous RGB24, modified in place."""
        h, w = surface.shape[:2]
        assert surface.flags["C_CONTIGUOUS"] and surface.shape[2] == 3
        return self.L.nes_oracle_render_string(_u8p(surface), w, h, position, text, len(text), glyphs._arr)

    def render_string4(self, surface: np.ndarray, position: int, text: bytes, glyphs: GlyphTable, fmt: str) -> int:
        """surface: uint8 [H, W, 4] contiguous, colour bytes stamped white, alpha untouched."""
        h, w = surface.shape[:2]
        assert surface.flags["C_CONTIGUOUS"] and surface.shape[2] == 4
        return self.L.nes_oracle_render_string4(_u8p(surface), w, h, position, text, len(text), glyphs._arr, 1 if PIXFMT[fmt][4] == 0 else 0)

    def composite(self, rgbs: list, depths: list, fmt: str):
        """rgbs[k]: uint8 [H,W,bpp]; depths[k]: uint8 [H,W] -> (rgb, depth)"""
        bpp, _, _, _, ao = PIXFMT[fmt]
        n = len(rgbs)
        h, w = depths[0].shape
        pr = (C.POINTER(C.c_uint8) * n)(*[_u8p(a) for a in rgbs])
        sr = (C.c_int * n)(*[a.strides[0] for a in rgbs])
        pd = (C.POINTER(C.c_uint8) * n)(*[_u8p(a) for a in depths])
        sd = (C.c_int * n)(*[a.strides[0] for a in depths])
        out = np.zeros((h, w, bpp), np.uint8)
        outd = np.zeros((h, w), np.uint8)
        self.L.nes_oracle_composite(n, pr, sr, bpp, max(ao, 0), pd, sd, w, h, _u8p(out), out.strides[0], _u8p(outd), outd.strides[0])
        return out, outd


def _site_packages() -> str:
    return sysconfig.get_paths()["purelib"]


def find_bundled_libs() -> dict:
    """Locate the wheel-bundled FFmpeg 8.0.1 / FreeType 2.14.3 binaries."""
    sp = _site_packages()
    d1 = os.path.join(sp, "opencv_python_headless.libs")
    d2 = os.path.join(sp, "pillow.libs")

    def one(d, pat):
        m = sorted(glob.glob(os.path.join(d, pat)))
        return m[0] if m else None

    return {
        "sws_deps": [p for p in (one(d1, "libcrypto-*"), one(d1, "libdrm-*")) if p],
        "avutil": one(d1, "libavutil-*"),
        "swscale": one(d1, "libswscale-*"),
        "ft_deps": [p for p in (one(d2, "libpng16-*"), one(d2, "libbrotlicommon-*"), one(d2, "libbrotlidec-*")) if p],
        "freetype": one(d2, "libfreetype-*"),
    }


class Ref:
    """The reference glue over the real libswscale / FreeType (oracle/ref_shim.c)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build()
        L = C.CDLL(REF_SO)
        self.L = L
        libs = find_bundled_libs()
        self.libs = libs
        L.nes_ref_dlopen.argtypes = [C.c_char_p]
        L.nes_ref_bind_swscale.argtypes = [C.c_char_p, C.c_char_p]
        L.nes_ref_bind_freetype.argtypes = [C.c_char_p]
        L.nes_ref_sws_convert.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_uint8), C.c_int] * 3
        L.nes_ref_text_new.argtypes = [C.c_char_p]
        L.nes_ref_text_new.restype = C.c_void_p
        L.nes_ref_text_free.argtypes = [C.c_void_p]
        L.nes_ref_freetype_version.argtypes = [C.c_void_p]
        L.nes_ref_text_render.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_int]
        L.nes_ref_text_render.restype = C.c_long
        L.nes_ref_text_render4.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.c_int]
        L.nes_ref_text_render4.restype = C.c_long
        L.nes_ref_text_glyph.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_int]
        L.nes_ref_swscale_version.restype = C.c_uint
        self.have_sws = False
        self.have_ft = False
        if libs["swscale"] and libs["avutil"]:
            for p in libs["sws_deps"]:
                L.nes_ref_dlopen(p.encode())
            self.have_sws = L.nes_ref_bind_swscale(libs["swscale"].encode(), libs["avutil"].encode()) == 0
        if libs["freetype"]:
            for p in libs["ft_deps"]:
                L.nes_ref_dlopen(p.encode())
            self.have_ft = L.nes_ref_bind_freetype(libs["freetype"].encode()) == 0

    def swscale_version(self) -> str:
        v = self.L.nes_ref_swscale_version()
        return f"{v >> 16}.{(v >> 8) & 255}.{v & 255}"

    def sws_convert(self, img: np.ndarray, fmt: str, wd: int | None = None, hd: int | None = None, flags: int = PARITY_FLAGS) -> Yuv:
        """types::SwsContextManager(source, dest): fmt in rgb24/bgr24/rgba/bgra/argb/abgr/gray."""
        h, w = img.shape[:2]
        wd, hd = wd or w, hd or h
        out = Yuv(wd, hd)
        r = self.L.nes_ref_sws_convert(img.ctypes.data_as(C.POINTER(C.c_uint8)), img.strides[0], fmt.encode(), w, h, wd, hd, flags, _u8p(out.y), out.ys, _u8p(out.u), out.cs, _u8p(out.v), out.cs)
        if r:
            raise RuntimeError(f"nes_ref_sws_convert failed: {r}")
        return out

    def text_new(self, font_path: str):
        t = self.L.nes_ref_text_new(font_path.encode())
        if not t:
            raise RuntimeError("FreeType could not open " + font_path)
        return t

    def text_free(self, t) -> None:
        self.L.nes_ref_text_free(t)

    def freetype_version(self, t) -> int:
        return self.L.nes_ref_freetype_version(t)

    def text_render(self, t, surface: np.ndarray, position: int, text: bytes) -> int:
        h, w = surface.shape[:2]
        assert surface.flags["C_CONTIGUOUS"] and surface.shape[2] == 3
        return self.L.nes_ref_text_render(t, _u8p(surface), w, h, position, text, len(text))

    def text_render4(self, t, surface: np.ndarray, position: int, text: bytes, fmt: str) -> int:
        h, w = surface.shape[:2]
        assert surface.flags["C_CONTIGUOUS"] and surface.shape[2] == 4
        return self.L.nes_ref_text_render4(t, _u8p(surface), w, h, position, text, len(text), 1 if PIXFMT[fmt][4] == 0 else 0)

    def glyph_table(self, t) -> GlyphTable:
        metrics = np.zeros((256, 5), np.int32)
        bitmaps = []
        buf = np.zeros(1 << 16, np.uint8)
        m5 = (C.c_int32 * 5)()
        for b in range(256):
            if b == 10:  # '\n' is never rasterised by the reference (render_text.cc:82-86)
                bitmaps.append(np.zeros(0, np.uint8))
                continue
            self.L.nes_ref_text_glyph(t, b, m5, _u8p(buf), buf.size)
            metrics[b] = list(m5)
            n = m5[0] * m5[1]
            bitmaps.append(buf[:n].copy())
        return GlyphTable(metrics, bitmaps)


# --------------------------------------------------------------------------
# wire format (src/server.cpp:91-112 length prefix; proto/nes.proto:18-25)
# --------------------------------------------------------------------------
def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def pack_rendered_frame(index: int, is_left: bool, width: int, height: int, matrix, frame: bytes, depth: bytes, prefix: bool = True, field_order=("index", "camera", "is_left", "frame", "depth")) -> bytes:
    """Serialise nesproto.RenderedFrame the way protobuf does (proto3: zero scalars omitted)."""
    cam = b""
    if is_left:
        cam += b"\x10\x01"
    if width:
        cam += b"\x18" + _varint(width)
    if height:
        cam += b"\x20" + _varint(height)
    if len(matrix):
        m = struct.pack("<%df" % len(matrix), *matrix)
        cam += b"\x62" + _varint(len(m)) + m
    parts = {
        "index": (b"\x08" + _varint(index)) if index else b"",
        "camera": b"\x12" + _varint(len(cam)) + cam,
        "is_left": b"\x18\x01" if is_left else b"",
        "frame": (b"\x32" + _varint(len(frame)) + frame) if len(frame) else b"",
        "depth": (b"\x3a" + _varint(len(depth)) + depth) if len(depth) else b"",
    }
    msg = b"".join(parts[k] for k in field_order)
    return (struct.pack("<Q", len(msg)) + msg) if prefix else msg


def _read_varint(buf: bytes, p: int):
    v = 0
    s = 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            return v, p


def unpack_rendered_frame(buf: bytes, prefix: bool = True) -> dict:
    """Python restatement of socket_receive_blocking_lpf + ParseFromString
    (server.cpp:91-112,175): returns the scalar fields and (offset, length) of the two
    bytes fields inside ``buf``.  Last occurrence of a field wins (protobuf semantics)."""
    p = 0
    end = len(buf)
    if prefix:
        (n,) = struct.unpack_from("<Q", buf, 0)
        p = 8
        end = 8 + n
        if end > len(buf):
            raise ValueError("truncated")
    out = {"index": 0, "is_left": False, "width": 0, "height": 0, "cam_is_left": False, "matrix": [], "frame": (0, 0), "depth": (0, 0)}
    while p < end:
        tag, p = _read_varint(buf, p)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, p = _read_varint(buf, p)
            if field == 1:
                out["index"] = v
            elif field == 3:
                out["is_left"] = bool(v)
        elif wt == 2:
            ln, p = _read_varint(buf, p)
            if p + ln > end:
                raise ValueError("truncated")
            if field == 2:
                q, qe = p, p + ln
                while q < qe:
                    t2, q = _read_varint(buf, q)
                    f2, w2 = t2 >> 3, t2 & 7
                    if w2 == 0:
                        v, q = _read_varint(buf, q)
                        if f2 == 2:
                            out["cam_is_left"] = bool(v)
                        elif f2 == 3:
                            out["width"] = v
                        elif f2 == 4:
                            out["height"] = v
                    elif w2 == 2:
                        l2, q = _read_varint(buf, q)
                        if f2 == 12:
                            out["matrix"] += list(struct.unpack_from("<%df" % (l2 // 4), buf, q))
                        q += l2
                    elif w2 == 5:
                        if f2 == 12:
                            out["matrix"].append(struct.unpack_from("<f", buf, q)[0])
                        q += 4
                    elif w2 == 1:
                        q += 8
                    else:
                        raise ValueError("bad wire type")
            elif field == 6:
                out["frame"] = (p, ln)
            elif field == 7:
                out["depth"] = (p, ln)
            p += ln
        elif wt == 5:
            p += 4
        elif wt == 1:
            p += 8
        else:
            raise ValueError("bad wire type")
    return out


# --------------------------------------------------------------------------
# synthetic inputs (SURVEY.md §8 d)
# --------------------------------------------------------------------------
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
    """Composite tests: source k is transparent on stripes ((x>>6)+k)%n != 0."""
    x = np.arange(w, dtype=np.int64)[None, :]
    a = np.where(((x >> 6) + k) % n != 0, 0, 255).astype(np.uint8)
    return np.broadcast_to(a, (h, w)).copy()


def to_fmt(rgb: np.ndarray, fmt: str, alpha: np.ndarray | None = None) -> np.ndarray:
    bpp, ro, go, bo, ao = PIXFMT[fmt]
    h, w = rgb.shape[:2]
    out = np.zeros((h, w, bpp), np.uint8)
    out[..., ro], out[..., go], out[..., bo] = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    if ao >= 0:
        out[..., ao] = 255 if alpha is None else alpha
    return out


KINITIAL_CAMERA_MATRIX = [1.0, 0.0, 0.0, 0.5, 0.0, -1.0, 0.0, 0.5, 0.0, 0.0, -1.0, 0.5]


def format_matrix_text(matrix) -> bytes:
    """encode.cpp:57-74: 12 camera floats + "0 0 0 1", each '%+.5f ' (showpos, fixed,
    setw 7 / fill '0' never pads because +d.ddddd is already 8 chars), '\\n' every 4."""
    s = ""
    for i, v in enumerate(matrix, 1):
        s += "%+.5f " % v
        if i % 4 == 0:
            s += "\n"
    s += "%+.5f %+.5f %+.5f %+.5f " % (0.0, 0.0, 0.0, 1.0)
    return s.encode()


def reference_strings(index: int = 0, timestamp: str = "12:34:56.789", is_left: bool = True, matrix=None):
    """The four overlays of encode.cpp:76-97 in call order: (position, text)."""
    return [
        (POS_CENTER, format_matrix_text(matrix or KINITIAL_CAMERA_MATRIX)),
        (POS_LEFT_BOTTOM, b"index=" + str(index).encode()),
        (POS_LEFT_TOP, timestamp.encode()),
        (POS_RIGHT_TOP, b"direction=left" if is_left else b"direction=right"),
    ]


# --------------------------------------------------------------------------
# one whole frame of the path, the way the reference would run it
# --------------------------------------------------------------------------
def stamp_runs(surface: np.ndarray, runs, fmt: str, port: "Port", glyphs: "GlyphTable | None" = None, ref: "Ref | None" = None, tctx=None) -> None:
    """render_string_to_frame (render_text.cc:35-111) for every run, in call order, in place.
    A run is (position, text) or (position, text, (x, y, w, h)): with a view the reference's loop runs on
    that sub-frame (copied out contiguous -- the loop assumes stride == width*bpp, render_text.cc:101 --
    stamped, copied back), i.e. once per eye of a side-by-side frame.  With ``ref``/``tctx`` the glyphs come
    from the real FreeType per character, else from the golden glyph table through the C port."""
    bpp = surface.shape[2]
    for r in runs or []:
        pos, txt = r[0], r[1]
        view = r[2] if len(r) > 2 and r[2][2] > 0 else None
        tgt = surface if view is None else np.ascontiguousarray(surface[view[1]:view[1] + view[3], view[0]:view[0] + view[2]])
        if ref is not None:
            ref.text_render(tctx, tgt, pos, txt) if bpp == 3 else ref.text_render4(tctx, tgt, pos, txt, fmt)
        else:
            port.render_string(tgt, pos, txt, glyphs) if bpp == 3 else port.render_string4(tgt, pos, txt, glyphs, fmt)
        if view is not None:
            surface[view[1]:view[1] + view[3], view[0]:view[0] + view[2]] = tgt


def expected_frame(srcs, fmt: str, runs, wd: int, hd: int, port: "Port", glyphs: "GlyphTable | None" = None, ref: "Ref | None" = None, tctx=None, flags: int = PARITY_FLAGS):
    """[depth composite (C port: the reference has none)] -> overlays -> scene and depth
    conversion (type_managers.cc:143-155).  srcs: [(pixels [h,w,bpp], depth [h,w])].  With ``ref`` the text
    and the conversion run in the real FreeType / libswscale, else in the port.  -> (Yuv scene, Yuv depth)"""
    if len(srcs) > 1:
        comp, cdep = port.composite([s[0] for s in srcs], [s[1] for s in srcs], fmt)
    else:
        comp, cdep = np.ascontiguousarray(srcs[0][0]).copy(), np.ascontiguousarray(srcs[0][1])
    stamp_runs(comp, runs, fmt, port, glyphs, ref, tctx)
    if ref is not None:
        return ref.sws_convert(comp, fmt, wd, hd, flags=flags), ref.sws_convert(cdep, "gray", wd, hd, flags=flags)
    return port.rgb_to_yuv420p(comp, fmt, wd, hd), port.gray_to_yuv420p(cdep, wd, hd)
