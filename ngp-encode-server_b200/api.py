"""Host-side mirror of the reference's hot-path interface over the C ABI (include/nes_gpu.h).

The product is ``libnes_gpu.so`` (CUDA kernels + C ABI).  This module only binds it with
ctypes and re-creates, in Python, the four reference entry points the path sits behind so
that tests and benchmarks read like the reference's own call sequence
(/root/reference/src/encode.cpp:76-98):

    FrameManager / FrameContext        include/base/video/type_managers.h:159-247
    SwsContextManager(source, dest)    src/base/video/type_managers.cc:143-155
    RenderTextContext                  src/base/video/render_text.cc:10-113
    RenderedFrame                      include/base/video/rendered_frame.h:15-69

There is no CPU fallback: every compute call goes to the GPU library and raises
``NesGpuError`` when it fails (no device, missing library, CUDA error).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NES_GPU_LIB") or os.path.join(HERE, "libnes_gpu.so")  # NES_GPU_LIB: diagnostic builds of the same library

NES_MAX_SOURCES = 8
NES_MEM_HOST, NES_MEM_DEVICE = 0, 1
NES_OUT_YUV420P, NES_OUT_NV12 = 0, 1

PIX_FMT = {"rgb24": 0, "bgr24": 1, "rgba": 2, "bgra": 3, "argb": 4, "abgr": 5}
PIX_BPP = {"rgb24": 3, "bgr24": 3, "rgba": 4, "bgra": 4, "argb": 4, "abgr": 4}

# RenderTextContext::RenderPosition (render_text.h:17-23)
RENDER_POSITION_LEFT_TOP, RENDER_POSITION_LEFT_BOTTOM, RENDER_POSITION_RIGHT_TOP, RENDER_POSITION_RIGHT_BOTTOM, RENDER_POSITION_CENTER = range(5)

NES_OK = 0
NES_ERR_INVALID_ARG, NES_ERR_SHORT_BUFFER, NES_ERR_TOO_LARGE, NES_ERR_CUDA = -1, -2, -3, -4
NES_ERR_NO_MEMORY, NES_ERR_BAD_TICKET, NES_ERR_PARSE, NES_ERR_NO_ATLAS, NES_ERR_FREETYPE, NES_ERR_BUSY, NES_ERR_UNSUPPORTED = -5, -6, -7, -8, -9, -10, -11


class NesGpuError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"{what}: {status} ({detail})" if detail else f"{what}: {status}")


class nes_gpu_cfg(C.Structure):
    _fields_ = [("device", C.c_int), ("max_width", C.c_int), ("max_height", C.c_int), ("max_sources", C.c_int), ("ring_depth", C.c_int), ("max_glyphs", C.c_int)]


class nes_glyph(C.Structure):
    _fields_ = [("code", C.c_int32), ("width", C.c_int32), ("rows", C.c_int32), ("left", C.c_int32), ("top", C.c_int32), ("advance", C.c_int32), ("pitch", C.c_int32), ("reserved", C.c_int32), ("coverage", C.c_void_p)]


class nes_text_run(C.Structure):
    _fields_ = [("position", C.c_int32), ("len", C.c_int32), ("text", C.c_char_p),
                ("view_x", C.c_int32), ("view_y", C.c_int32), ("view_w", C.c_int32), ("view_h", C.c_int32)]


class nes_source(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("depth", C.c_void_p), ("rgb_stride", C.c_int32), ("depth_stride", C.c_int32), ("rgb_bytes", C.c_uint64), ("depth_bytes", C.c_uint64)]


class nes_frame_in(C.Structure):
    _fields_ = [("n_sources", C.c_int32), ("pix_fmt", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("mem", C.c_int32), ("depth_fmt", C.c_int32), ("src", nes_source * NES_MAX_SOURCES)]


class nes_frame_out(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mem", C.c_int32), ("pix_fmt", C.c_int32),
                ("scene", C.c_void_p * 3), ("scene_linesize", C.c_int32 * 3), ("reserved2", C.c_int32),
                ("depth", C.c_void_p * 3), ("depth_linesize", C.c_int32 * 3), ("reserved3", C.c_int32)]


class nes_timing(C.Structure):
    _fields_ = [("h2d_us", C.c_float), ("kernels_us", C.c_float), ("d2h_us", C.c_float), ("total_us", C.c_float), ("n_launches", C.c_int32), ("reserved", C.c_int32)]


class nes_mux_stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("launch_sets", C.c_uint64), ("launches", C.c_uint64), ("max_batch", C.c_uint64)]


class nes_placed_glyph(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("code", C.c_int32), ("reserved", C.c_int32),
                ("clip_x", C.c_int32), ("clip_y", C.c_int32), ("clip_w", C.c_int32), ("clip_h", C.c_int32)]


class nes_unpacked_frame(C.Structure):
    _fields_ = [("index", C.c_uint64), ("is_left", C.c_int32), ("cam_is_left", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("n_matrix", C.c_int32), ("matrix", C.c_float * 16), ("frame_off", C.c_uint64), ("frame_len", C.c_uint64),
                ("depth_off", C.c_uint64), ("depth_len", C.c_uint64), ("consumed", C.c_uint64)]


# every symbol include/nes_gpu.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "nes_gpu_abi_version", "nes_gpu_strerror", "nes_gpu_session_error", "nes_gpu_device_count",
    "nes_gpu_session_create", "nes_gpu_session_destroy", "nes_gpu_session_stream", "nes_gpu_session_launches", "nes_gpu_session_set_latency_bands",
    "nes_gpu_host_alloc", "nes_gpu_host_free", "nes_gpu_device_alloc", "nes_gpu_device_free",
    "nes_gpu_memcpy_h2d", "nes_gpu_memcpy_d2h", "nes_gpu_atlas_set", "nes_gpu_atlas_load_font",
    "nes_font_rasterise", "nes_gpu_submit", "nes_gpu_wait", "nes_gpu_convert", "nes_gpu_convert_batch_device", "nes_gpu_last_timing",
    "nes_gpu_batch_prepare", "nes_gpu_batch_run", "nes_gpu_batch_free",
    "nes_avframe_wrap", "nes_avframe_free", "nes_avframe_ref_count", "nes_avframe_error",
    "nes_gpu_mux_create", "nes_gpu_mux_destroy", "nes_gpu_mux_attach", "nes_gpu_mux_stats", "nes_gpu_mux_error",
    "nes_gpu_filter_table", "nes_gpu_text_layout", "nes_unpack_rendered_frame",
    "nes_ingest_ring_create", "nes_ingest_ring_destroy", "nes_ingest_acquire", "nes_ingest_commit", "nes_ingest_release",
]

_lib = None


def lib() -> C.CDLL:
    """Load libnes_gpu.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NesGpuError(NES_ERR_CUDA, "libnes_gpu.so is not built", "run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.nes_gpu_abi_version.restype = i32
    L.nes_gpu_strerror.restype = C.c_char_p
    L.nes_gpu_strerror.argtypes = [i32]
    L.nes_gpu_session_error.restype = C.c_char_p
    L.nes_gpu_session_error.argtypes = [vp]
    L.nes_gpu_session_create.argtypes = [C.POINTER(nes_gpu_cfg), C.POINTER(vp)]
    L.nes_gpu_session_destroy.argtypes = [vp]
    L.nes_gpu_session_destroy.restype = None
    L.nes_gpu_session_stream.argtypes = [vp]
    L.nes_gpu_session_stream.restype = vp
    L.nes_gpu_session_launches.argtypes = [vp]
    L.nes_gpu_session_set_latency_bands.argtypes = [vp, i32]
    L.nes_gpu_session_launches.restype = u64
    L.nes_gpu_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.nes_gpu_host_free.argtypes = [vp]
    L.nes_gpu_host_free.restype = None
    L.nes_gpu_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.nes_gpu_device_free.argtypes = [vp, vp]
    L.nes_gpu_device_free.restype = None
    L.nes_gpu_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.nes_gpu_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.nes_gpu_atlas_set.argtypes = [vp, C.POINTER(nes_glyph), i32]
    L.nes_gpu_atlas_load_font.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.nes_font_rasterise.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(nes_glyph), vp, u64, C.POINTER(u64)]
    L.nes_gpu_submit.argtypes = [vp, C.POINTER(nes_frame_in), C.POINTER(nes_text_run), i32, C.POINTER(nes_frame_out), C.POINTER(u64)]
    L.nes_gpu_wait.argtypes = [vp, u64]
    L.nes_gpu_convert.argtypes = [vp, C.POINTER(nes_frame_in), C.POINTER(nes_text_run), i32, C.POINTER(nes_frame_out)]
    L.nes_gpu_convert_batch_device.argtypes = [vp, i32, C.POINTER(nes_frame_in), C.POINTER(C.POINTER(nes_text_run)), C.POINTER(i32), C.POINTER(nes_frame_out), i32]
    L.nes_gpu_last_timing.argtypes = [vp, C.POINTER(nes_timing)]
    L.nes_gpu_batch_prepare.argtypes = [vp, i32, C.POINTER(nes_frame_in), C.POINTER(C.POINTER(nes_text_run)), C.POINTER(i32), C.POINTER(nes_frame_out), C.POINTER(vp)]
    L.nes_gpu_batch_run.argtypes = [vp, vp, i32]
    L.nes_gpu_batch_free.argtypes = [vp, vp]
    L.nes_gpu_batch_free.restype = None
    L.nes_gpu_filter_table.argtypes = [i32, i32, i32, C.POINTER(C.c_int16), i32, C.POINTER(C.c_int32), i32]
    L.nes_gpu_text_layout.argtypes = [vp, i32, i32, C.POINTER(nes_text_run), C.POINTER(nes_placed_glyph), i32]
    L.nes_unpack_rendered_frame.argtypes = [vp, u64, i32, C.POINTER(nes_unpacked_frame)]
    L.nes_ingest_ring_create.argtypes = [i32, u64, C.POINTER(vp)]
    L.nes_ingest_ring_destroy.argtypes = [vp]
    L.nes_ingest_ring_destroy.restype = None
    L.nes_ingest_acquire.argtypes = [vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(u64)]
    L.nes_ingest_commit.argtypes = [vp, i32, u64, i32, i32, C.POINTER(nes_unpacked_frame), C.POINTER(nes_source)]
    L.nes_ingest_release.argtypes = [vp, i32]
    L.nes_gpu_mux_create.argtypes = [i32, i32, C.POINTER(vp)]
    L.nes_gpu_mux_destroy.argtypes = [vp]
    L.nes_gpu_mux_destroy.restype = None
    L.nes_gpu_mux_attach.argtypes = [vp, vp]
    L.nes_gpu_mux_stats.argtypes = [vp, C.POINTER(nes_mux_stats)]
    L.nes_gpu_mux_error.argtypes = [vp]
    L.nes_gpu_mux_error.restype = C.c_char_p
    L.nes_avframe_wrap.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(i32), i32, i32, i32, C.c_int64, vp, vp, C.POINTER(vp)]
    L.nes_avframe_free.argtypes = [C.POINTER(vp)]
    L.nes_avframe_free.restype = None
    L.nes_avframe_ref_count.argtypes = [vp]
    L.nes_avframe_error.restype = C.c_char_p
    _lib = L
    return L


def strerror(status: int) -> str:
    return lib().nes_gpu_strerror(status).decode()


def device_count() -> int:
    return lib().nes_gpu_device_count()


def align32(x: int) -> int:
    return (x + 31) & ~31


def find_freetype() -> str | None:
    """The FreeType binary of this image (pillow wheel), SURVEY.md Appendix B."""
    env = os.environ.get("NES_FREETYPE_SO")
    if env and os.path.exists(env):
        return env
    m = sorted(glob.glob(os.path.join(sysconfig.get_paths()["purelib"], "pillow.libs", "libfreetype-*")))
    return m[0] if m else None


# --------------------------------------------------------------------------------------
# host-only helpers (no GPU needed)
# --------------------------------------------------------------------------------------
def filter_table(src: int, dst: int, one: int):
    """libswscale initFilter() table for the default bicubic scaler -> (coef[dst,size], pos[dst])."""
    L = lib()
    size = L.nes_gpu_filter_table(src, dst, one, None, 0, None, 0)
    if size < 0:
        raise NesGpuError(size, "nes_gpu_filter_table", strerror(size))
    coef = np.zeros((dst, size), np.int16)
    pos = np.zeros(dst, np.int32)
    r = L.nes_gpu_filter_table(src, dst, one, coef.ctypes.data_as(C.POINTER(C.c_int16)), coef.size, pos.ctypes.data_as(C.POINTER(C.c_int32)), pos.size)
    if r < 0:
        raise NesGpuError(r, "nes_gpu_filter_table", strerror(r))
    return coef, pos


def unpack_rendered_frame(buf, prefix: bool = True) -> dict:
    """Locate the fields of a (length-prefixed) nesproto.RenderedFrame without copying."""
    a = np.frombuffer(buf, np.uint8)
    u = nes_unpacked_frame()
    r = lib().nes_unpack_rendered_frame(a.ctypes.data, a.size, 1 if prefix else 0, C.byref(u))
    if r:
        raise NesGpuError(r, "nes_unpack_rendered_frame", strerror(r))
    return {
        "index": u.index, "is_left": bool(u.is_left), "cam_is_left": bool(u.cam_is_left), "width": u.width, "height": u.height,
        "matrix": [u.matrix[i] for i in range(u.n_matrix)], "frame": (u.frame_off, u.frame_len), "depth": (u.depth_off, u.depth_len),
        "consumed": u.consumed,
    }


class IngestRing:
    """Pinned receive ring (nes_ingest_*): recv() a length-prefixed RenderedFrame straight into a
    slot, commit (parsed in place), submit the returned source, release after wait.  Replaces the
    three payload copies of server.cpp:91-112,175 + rendered_frame.cc:5-27."""

    def __init__(self, n_slots: int, slot_bytes: int):
        self.L = lib()
        h = C.c_void_p()
        r = self.L.nes_ingest_ring_create(n_slots, slot_bytes, C.byref(h))
        if r:
            raise NesGpuError(r, "nes_ingest_ring_create", strerror(r))
        self.h = h

    def acquire(self):
        """-> (slot, uint8 view of the slot's pinned memory)"""
        slot, buf, cap = C.c_int(), C.c_void_p(), C.c_uint64()
        r = self.L.nes_ingest_acquire(self.h, C.byref(slot), C.byref(buf), C.byref(cap))
        if r:
            raise NesGpuError(r, "nes_ingest_acquire", strerror(r))
        return slot.value, np.ctypeslib.as_array((C.c_uint8 * cap.value).from_address(buf.value))

    def commit(self, slot: int, length: int, bytes_per_pixel: int = 3, prefix: bool = True):
        """-> (nes_unpacked_frame, nes_source pointing into the slot)"""
        info, src = nes_unpacked_frame(), nes_source()
        r = self.L.nes_ingest_commit(self.h, slot, length, 1 if prefix else 0, bytes_per_pixel, C.byref(info), C.byref(src))
        if r:
            raise NesGpuError(r, "nes_ingest_commit", strerror(r))
        return info, src

    def release(self, slot: int):
        r = self.L.nes_ingest_release(self.h, slot)
        if r:
            raise NesGpuError(r, "nes_ingest_release", strerror(r))

    def close(self):
        if self.h:
            self.L.nes_ingest_ring_destroy(self.h)
            self.h = None


def font_rasterise(font_path: str, freetype_so: str | None = None):
    """Rasterise codes 0..255 like RenderTextContext does per character (render_text.cc:12-32,88)
    -> (metrics int32 [256,5] = width, rows, left, top, advance; bitmaps list of uint8 arrays)."""
    ft = freetype_so or find_freetype()
    arr = (nes_glyph * 256)()
    cov = np.zeros(1 << 20, np.uint8)
    used = C.c_uint64()
    r = lib().nes_font_rasterise(ft.encode() if ft else None, font_path.encode(), arr, cov.ctypes.data, cov.size, C.byref(used))
    if r:
        raise NesGpuError(r, "nes_font_rasterise", strerror(r))
    metrics = np.zeros((256, 5), np.int32)
    bitmaps = []
    for b in range(256):
        g = arr[b]
        metrics[b] = (g.width, g.rows, g.left, g.top, g.advance)
        n = g.width * g.rows
        off = (g.coverage - cov.ctypes.data) if g.coverage else 0
        bitmaps.append(cov[off:off + n].copy())
    return metrics, bitmaps


def format_camera_matrix(matrix) -> bytes:
    """The matrix text process_frame_thread builds (encode.cpp:57-74): every value as
    showpos/fixed/precision 5 followed by a space, newline after every 4th camera value,
    then the constant row "0 0 0 1"."""
    s = ""
    for i, v in enumerate(matrix, 1):
        s += "%+.5f " % v
        if i % 4 == 0:
            s += "\n"
    s += "%+.5f %+.5f %+.5f %+.5f " % (0.0, 0.0, 0.0, 1.0)
    return s.encode()


# --------------------------------------------------------------------------------------
# session + reference-shaped classes
# --------------------------------------------------------------------------------------
class Session:
    """One hot-path worker (the GPU equivalent of one process_frame_thread)."""

    def __init__(self, device: int = 0, max_width: int = 7680, max_height: int = 4320, max_sources: int = 4, ring_depth: int = 3, max_glyphs: int = 8192):
        self.L = lib()
        self.h = C.c_void_p()
        cfg = nes_gpu_cfg(device, max_width, max_height, max_sources, ring_depth, max_glyphs)
        r = self.L.nes_gpu_session_create(C.byref(cfg), C.byref(self.h))
        if r:
            self.h = C.c_void_p()
            raise NesGpuError(r, "nes_gpu_session_create", strerror(r))
        self.device = device
        self._keep = {}
        self._dev = set()

    def close(self):
        """Destroys the session and frees the pinned buffers host_array() handed out (arrays obtained from
        host_array must not be touched afterwards)."""
        if self.h:
            for p in list(self._dev):
                self.L.nes_gpu_device_free(self.h, p)
            self._dev.clear()
            self.L.nes_gpu_session_destroy(self.h)
            self.h = C.c_void_p()
            for p in self._keep.values():
                self.L.nes_gpu_host_free(p)
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, r: int, what: str):
        if r:
            raise NesGpuError(r, what, strerror(r) + " | " + self.L.nes_gpu_session_error(self.h).decode())

    @property
    def stream(self) -> int:
        return self.L.nes_gpu_session_stream(self.h)

    @property
    def launches(self) -> int:
        return self.L.nes_gpu_session_launches(self.h)

    def set_latency_bands(self, bands: int):
        """Low-latency mode: pinned-to-pinned same-size frames go through in `bands` row bands
        (download of a band overlaps the upload of the next)."""
        self._check(self.L.nes_gpu_session_set_latency_bands(self.h, bands), "nes_gpu_session_set_latency_bands")

    # -- memory ------------------------------------------------------------------
    def host_array(self, nbytes: int) -> np.ndarray:
        """Pinned host bytes as a numpy array (freed with the session object that owns it)."""
        p = C.c_void_p()
        r = self.L.nes_gpu_host_alloc(nbytes, C.byref(p))
        if r:
            raise NesGpuError(r, "nes_gpu_host_alloc", strerror(r))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, np.uint8)
        self._keep[arr.ctypes.data] = p
        return arr

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.L.nes_gpu_device_alloc(self.h, nbytes, C.byref(p)), "nes_gpu_device_alloc")
        self._dev.add(p.value)
        return p.value

    def device_free(self, p: int):
        self._dev.discard(p)
        self.L.nes_gpu_device_free(self.h, p)

    def h2d(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self.L.nes_gpu_memcpy_h2d(self.h, dst, src.ctypes.data, src.nbytes), "nes_gpu_memcpy_h2d")

    def d2h(self, dst: np.ndarray, src: int):
        self._check(self.L.nes_gpu_memcpy_d2h(self.h, dst.ctypes.data, src, dst.nbytes), "nes_gpu_memcpy_d2h")

    # -- atlas -------------------------------------------------------------------
    def atlas_set(self, metrics: np.ndarray, bitmaps: list):
        """metrics int32 [256,5] = (width, rows, left, top, advance); bitmaps[b] = rows*width bytes."""
        arr = (nes_glyph * 256)()
        keep = []
        for b in range(256):
            w, rows, left, top, adv = (int(v) for v in metrics[b])
            bm = np.ascontiguousarray(bitmaps[b], np.uint8)
            keep.append(bm)
            arr[b] = nes_glyph(b, w, rows, left, top, adv, w, 0, bm.ctypes.data if bm.size else None)
        self._check(self.L.nes_gpu_atlas_set(self.h, arr, 256), "nes_gpu_atlas_set")

    def atlas_load_font(self, font_path: str, freetype_so: str | None = None):
        ft = freetype_so or find_freetype()
        self._check(self.L.nes_gpu_atlas_load_font(self.h, ft.encode() if ft else None, font_path.encode()), "nes_gpu_atlas_load_font")

    def text_layout(self, w: int, h: int, position: int, text: bytes, view=(0, 0, 0, 0)):
        run = nes_text_run(position, len(text), text, *view)
        cap = max(len(text), 1)
        out = (nes_placed_glyph * cap)()
        n = self.L.nes_gpu_text_layout(self.h, w, h, C.byref(run), out, cap)
        if n < 0:
            self._check(n, "nes_gpu_text_layout")
        return [(out[i].x, out[i].y, out[i].code) for i in range(n)]

    # -- descriptors -------------------------------------------------------------
    @staticmethod
    def make_runs(runs):
        """runs: [(position, bytes)] or [(position, bytes, (view_x, view_y, view_w, view_h))] -> (ctypes array, n)"""
        n = len(runs or [])
        arr = (nes_text_run * max(n, 1))()
        for i, r in enumerate(runs or []):
            arr[i] = nes_text_run(r[0], len(r[1]), r[1], *(r[2] if len(r) > 2 else (0, 0, 0, 0)))
        return arr, n

    @staticmethod
    def frame_in(fmt: str, width: int, height: int, sources, mem: int = NES_MEM_HOST, depth_fmt: str = "gray") -> nes_frame_in:
        """sources: [(rgb, depth_or_None, rgb_stride, depth_stride)] with numpy arrays (host) or
        (ptr, nbytes) tuples (device).  depth_fmt "gray16le": the depth arrays hold 2 bytes per sample (stride in bytes)."""
        fi = nes_frame_in()
        fi.n_sources, fi.pix_fmt, fi.width, fi.height, fi.mem = len(sources), PIX_FMT[fmt], width, height, mem
        fi.depth_fmt = {"gray": 0, "gray8": 0, "gray16le": 1}[depth_fmt]
        for k, (rgb, depth, rs, ds) in enumerate(sources):
            s = fi.src[k]
            if isinstance(rgb, np.ndarray):
                s.rgb, s.rgb_bytes = rgb.ctypes.data, rgb.nbytes
            else:
                s.rgb, s.rgb_bytes = rgb
            if depth is None:
                s.depth, s.depth_bytes = None, 0
            elif isinstance(depth, np.ndarray):
                s.depth, s.depth_bytes = depth.ctypes.data, depth.nbytes
            else:
                s.depth, s.depth_bytes = depth
            s.rgb_stride, s.depth_stride = rs, ds
        return fi

    # -- hot path ----------------------------------------------------------------
    def submit(self, fin: nes_frame_in, runs, fout: nes_frame_out) -> int:
        arr, n = runs if isinstance(runs, tuple) else self.make_runs(runs)
        t = C.c_uint64()
        self._check(self.L.nes_gpu_submit(self.h, C.byref(fin), arr, n, C.byref(fout), C.byref(t)), "nes_gpu_submit")
        return t.value

    def wait(self, ticket: int):
        self._check(self.L.nes_gpu_wait(self.h, ticket), "nes_gpu_wait")

    def convert(self, fin: nes_frame_in, runs, fout: nes_frame_out):
        arr, n = runs if isinstance(runs, tuple) else self.make_runs(runs)
        self._check(self.L.nes_gpu_convert(self.h, C.byref(fin), arr, n, C.byref(fout)), "nes_gpu_convert")

    def convert_batch_device(self, fins, runs_list, fouts, sync: bool = True):
        n = len(fins)
        a_in = (nes_frame_in * n)(*fins)
        a_out = (nes_frame_out * n)(*fouts)
        if runs_list:
            made = [self.make_runs(r) for r in runs_list]
            a_runs = (C.POINTER(nes_text_run) * n)(*[C.cast(m[0], C.POINTER(nes_text_run)) for m in made])
            a_n = (C.c_int * n)(*[m[1] for m in made])
        else:
            made, a_runs, a_n = None, None, None
        self._check(self.L.nes_gpu_convert_batch_device(self.h, n, a_in, a_runs, a_n, a_out, 1 if sync else 0), "nes_gpu_convert_batch_device")
        return (a_in, a_out, made)  # keep-alive for async callers

    def prepare_batch(self, fins, runs_list, fouts):
        """Build the ctypes argument arrays of convert_batch_device once (a streaming caller
        re-submits the same descriptors every step)."""
        n = len(fins)
        a_in = (nes_frame_in * n)(*fins)
        a_out = (nes_frame_out * n)(*fouts)
        made = [self.make_runs(r) for r in runs_list] if runs_list else None
        a_runs = (C.POINTER(nes_text_run) * n)(*[C.cast(m[0], C.POINTER(nes_text_run)) for m in made]) if made else None
        a_n = (C.c_int * n)(*[m[1] for m in made]) if made else None
        return (n, a_in, a_runs, a_n, a_out, made, runs_list)

    def run_batch(self, prepared, sync: bool = False):
        n, a_in, a_runs, a_n, a_out = prepared[:5]
        r = self.L.nes_gpu_convert_batch_device(self.h, n, a_in, a_runs, a_n, a_out, 1 if sync else 0)
        if r:
            self._check(r, "nes_gpu_convert_batch_device")

    def batch_prepare(self, fins, runs_list, fouts):
        """nes_gpu_batch_prepare: descriptor table of device-resident frames built + uploaded once -> handle."""
        n, a_in, a_runs, a_n, a_out = self.prepare_batch(fins, runs_list, fouts)[:5]
        h = C.c_void_p()
        self._check(self.L.nes_gpu_batch_prepare(self.h, n, a_in, a_runs, a_n, a_out, C.byref(h)), "nes_gpu_batch_prepare")
        return h

    def batch_run(self, handle, sync: bool = False):
        r = self.L.nes_gpu_batch_run(self.h, handle, 1 if sync else 0)
        if r:
            self._check(r, "nes_gpu_batch_run")

    def batch_free(self, handle):
        self.L.nes_gpu_batch_free(self.h, handle)

    def submit_prepared(self, fin, runs_made, fout) -> int:
        t = C.c_uint64()
        r = self.L.nes_gpu_submit(self.h, C.byref(fin), runs_made[0], runs_made[1], C.byref(fout), C.byref(t))
        if r:
            self._check(r, "nes_gpu_submit")
        return t.value

    def last_timing(self) -> dict:
        t = nes_timing()
        self._check(self.L.nes_gpu_last_timing(self.h, C.byref(t)), "nes_gpu_last_timing")
        return {"h2d_us": t.h2d_us, "kernels_us": t.kernels_us, "d2h_us": t.d2h_us, "total_us": t.total_us, "n_launches": t.n_launches}


class Mux:
    """nes_gpu_mux: the client sessions of one GPU share one dispatcher -- nes_gpu_submit on an attached session only
    stages the frame, the ready frames of all sessions go out in one launch (BASELINE config 4)."""

    def __init__(self, device: int = 0, max_batch: int = 64):
        self.L = lib()
        self.h = C.c_void_p()
        r = self.L.nes_gpu_mux_create(device, max_batch, C.byref(self.h))
        if r:
            self.h = C.c_void_p()
            raise NesGpuError(r, "nes_gpu_mux_create", strerror(r))
        self.device = device

    def attach(self, session: "Session"):
        r = self.L.nes_gpu_mux_attach(self.h, session.h)
        if r:
            raise NesGpuError(r, "nes_gpu_mux_attach", strerror(r))

    def stats(self) -> dict:
        st = nes_mux_stats()
        r = self.L.nes_gpu_mux_stats(self.h, C.byref(st))
        if r:
            raise NesGpuError(r, "nes_gpu_mux_stats", strerror(r))
        return {"frames": st.frames, "launch_sets": st.launch_sets, "launches": st.launches, "max_batch": st.max_batch}

    def error(self) -> str:
        return self.L.nes_gpu_mux_error(self.h).decode()

    def close(self):
        """Sessions still attached go back to their own streams (either order of destruction is fine)."""
        if self.h:
            self.L.nes_gpu_mux_destroy(self.h)
            self.h = C.c_void_p()


class FrameContext:
    """types::FrameManager::FrameContext (type_managers.h:172-184)."""

    def __init__(self, width: int, height: int, pix_fmt: str):
        self.width, self.height, self.pix_fmt = width, height, pix_fmt


class FrameManager:
    """types::FrameManager (type_managers.h:159-247, type_managers.cc:116-141).

    Owning mode (buffer is None) allocates the planes like av_image_alloc(..., align=32): one
    contiguous block, strides aligned to 32, from pinned memory when a session is given.
    Borrowing mode wraps an external buffer with tight line sizes (av_image_get_linesize)."""

    kBufferSizeAlignValueBytes = 32

    def __init__(self, context: FrameContext, buffer: np.ndarray | None = None, session: Session | None = None):
        self.context = context
        self._owner = session  # pinned planes live as long as the session that allocated them
        w, h = context.width, context.height
        self.text_runs = []  # overlays queued by RenderTextContext, applied on the device copy
        if context.pix_fmt == "yuv420p":
            cw, ch = (w + 1) // 2, (h + 1) // 2
            self.linesize = [align32(w), align32(cw), align32(cw)]
            sizes = [self.linesize[0] * h, self.linesize[1] * ch, self.linesize[2] * ch]
            total = sum(sizes)
            if buffer is None:
                buffer = session.host_array(total) if session is not None else np.zeros(total, np.uint8)
            self.buffer = buffer
            o1, o2 = sizes[0], sizes[0] + sizes[1]
            self.planes = [buffer[:o1].reshape(h, self.linesize[0]), buffer[o1:o2].reshape(ch, self.linesize[1]), buffer[o2:o2 + sizes[2]].reshape(ch, self.linesize[2])]
            self.data = [p.ctypes.data for p in self.planes]
        elif context.pix_fmt == "nv12":
            # Y plane + one interleaved UV plane (the layout hardware encoders take)
            cw, ch = (w + 1) // 2, (h + 1) // 2
            self.linesize = [align32(w), align32(2 * cw), 0]
            sizes = [self.linesize[0] * h, self.linesize[1] * ch]
            if buffer is None:
                buffer = session.host_array(sum(sizes)) if session is not None else np.zeros(sum(sizes), np.uint8)
            self.buffer = buffer
            self.planes = [buffer[:sizes[0]].reshape(h, self.linesize[0]), buffer[sizes[0]:sizes[0] + sizes[1]].reshape(ch, self.linesize[1])]
            self.data = [p.ctypes.data for p in self.planes] + [0]
        else:
            bpp = 1 if context.pix_fmt == "gray" else PIX_BPP[context.pix_fmt]
            if buffer is None:
                ls = align32(w * bpp)
                buffer = session.host_array(ls * h) if session is not None else np.zeros(ls * h, np.uint8)
                self.linesize = [ls]
            else:
                self.linesize = [w * bpp]
            self.buffer = buffer
            self.data = [buffer.ctypes.data]

    def cropped(self) -> bytes:
        """Y||U||V (YUV420P) or Y||UV (NV12) without stride padding."""
        w, h = self.context.width, self.context.height
        cw = (w + 1) // 2
        if self.context.pix_fmt == "nv12":
            return self.planes[0][:, :w].tobytes() + self.planes[1][:, :2 * cw].tobytes()
        return self.planes[0][:, :w].tobytes() + self.planes[1][:, :cw].tobytes() + self.planes[2][:, :cw].tobytes()


def _frame_out(scene: FrameManager, depth: FrameManager | None) -> nes_frame_out:
    fo = nes_frame_out()
    fo.width, fo.height, fo.mem = scene.context.width, scene.context.height, NES_MEM_HOST
    fo.pix_fmt = NES_OUT_NV12 if scene.context.pix_fmt == "nv12" else NES_OUT_YUV420P
    for p in range(3):
        fo.scene[p] = scene.data[p]
        fo.scene_linesize[p] = scene.linesize[p]
        if depth is not None:
            fo.depth[p] = depth.data[p]
            fo.depth_linesize[p] = depth.linesize[p]
    return fo


class SwsContextManager:
    """types::SwsContextManager(source, dest) (type_managers.cc:143-155): converts on
    construction.  source is a packed RGB(A) frame or a GRAY8 frame, dest a YUV420P frame.
    Text runs queued on the source frame are stamped on the device copy first."""

    def __init__(self, source: FrameManager, dest: FrameManager, session: Session):
        sc, w, h = source.context, source.context.width, source.context.height
        if sc.pix_fmt == "gray":
            # the C ABI converts depth next to a scene; a lone GRAY8 conversion rides with a
            # 3-byte dummy scene of the same size whose output is discarded
            dummy = np.zeros(w * h * 3, np.uint8)
            fin = Session.frame_in("rgb24", w, h, [(dummy, source.buffer, 0, source.linesize[0])])
            scratch = FrameManager(FrameContext(dest.context.width, dest.context.height, "yuv420p"))
            session.convert(fin, None, _frame_out(scratch, dest))
        else:
            fin = Session.frame_in(sc.pix_fmt, w, h, [(source.buffer, None, source.linesize[0], 0)])
            session.convert(fin, source.text_runs, _frame_out(dest, None))


class RenderTextContext:
    """RenderTextContext (render_text.h:15-39).  The constructor rasterises the font once
    into the session's device atlas; render_string_to_frame queues the run on the frame (the
    stamp itself happens on the device copy inside the next conversion, in call order)."""

    def __init__(self, font_location: str, session: Session):
        self.session = session
        session.atlas_load_font(font_location)

    def render_string_to_frame(self, frame: FrameManager, opt: int, content):
        if isinstance(content, str):
            content = content.encode()
        frame.text_runs.append((opt, content))


class RenderedFrame:
    """RenderedFrame (rendered_frame.h:15-69, rendered_frame.cc:5-27) built straight from the
    wire bytes (server.cpp:172-194): the scene/depth payloads are views into ``message``."""

    def __init__(self, message, pix_fmt_scene: str, pix_fmt_depth: str, codec_scene: tuple, codec_depth: tuple, session: Session, prefix: bool = True):
        self.session = session
        self._msg = np.frombuffer(message, np.uint8)
        self.fields = unpack_rendered_frame(self._msg, prefix)
        w, h = self.fields["width"], self.fields["height"]
        fo, fl = self.fields["frame"]
        do, dl = self.fields["depth"]
        self.m_source_avframe_scene = FrameManager(FrameContext(w, h, pix_fmt_scene), self._msg[fo:fo + fl])
        self.m_converted_avframe_scene = FrameManager(FrameContext(codec_scene[0], codec_scene[1], "yuv420p"), session=session)
        self.m_source_avframe_depth = FrameManager(FrameContext(w, h, pix_fmt_depth), self._msg[do:do + dl])
        self.m_converted_avframe_depth = FrameManager(FrameContext(codec_depth[0], codec_depth[1], "yuv420p"), session=session)
        self.m_converted = False

    def index(self) -> int:
        return self.fields["index"]

    def is_left(self) -> bool:
        return self.fields["is_left"]

    def get_cam(self) -> dict:
        return self.fields

    def source_frame_scene(self) -> FrameManager:
        return self.m_source_avframe_scene

    def converted_frame_scene(self) -> FrameManager:
        return self.m_converted_avframe_scene

    def converted_frame_depth(self) -> FrameManager:
        return self.m_converted_avframe_depth

    def convert_frame(self):
        if self.m_converted:
            raise RuntimeError("Tried to convert a converted RenderedFrame.")
        src, dep = self.m_source_avframe_scene, self.m_source_avframe_depth
        fin = Session.frame_in(src.context.pix_fmt, src.context.width, src.context.height, [(src.buffer, dep.buffer, src.linesize[0], dep.linesize[0])])
        self.session.convert(fin, src.text_runs, _frame_out(self.m_converted_avframe_scene, self.m_converted_avframe_depth))
        self.m_converted = True


def process_frame(frame: RenderedFrame, etctx: RenderTextContext, timestamp: str):
    """Body of process_frame_thread for one frame (encode.cpp:55-98)."""
    scene = frame.source_frame_scene()
    etctx.render_string_to_frame(scene, RENDER_POSITION_CENTER, format_camera_matrix(frame.get_cam()["matrix"]))
    etctx.render_string_to_frame(scene, RENDER_POSITION_LEFT_BOTTOM, b"index=" + str(frame.index()).encode())
    etctx.render_string_to_frame(scene, RENDER_POSITION_LEFT_TOP, timestamp)
    etctx.render_string_to_frame(scene, RENDER_POSITION_RIGHT_TOP, "direction=left" if frame.is_left() else "direction=right")
    frame.convert_frame()
