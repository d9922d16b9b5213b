"""Encoder hand-off helpers (host side, ctypes): the stage right after the hot path.

The reference hands the converted planes to libavcodec in send_frame_thread
(/root/reference/src/encode.cpp:133-165: ``codecctx->send_frame(frame.to_avframe().get())``; codec set-up in
src/base/video/type_managers.cc:47-110, H.264 through libx264).  Two things live here:

* ``wrap_frame`` -- ``nes_avframe_wrap`` of the C ABI: a FrameManager's planes as a ref-counted AVFrame
  (the encoder takes a reference instead of copying; replaces type_managers.h:187-239).
* ``Encoder`` / ``time_substitute_encoder`` -- libavcodec bound with ctypes, configured through AVOptions with the
  reference's defaults (400 kbit/s, GOP 250, 30 fps: main.cpp:109-123).  The libavcodec bundled in this image
  (opencv wheel, FFmpeg 8.0.1) has NO H.264 encoder, so the downstream stage can only be timed with a SUBSTITUTE
  (mpeg4), always reported under that label and never mixed into the hot-path figures.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sysconfig
import time

import numpy as np

from . import api

AV_PIX_FMT_YUV420P = 0
_EAGAIN = -11


def bundled_ffmpeg() -> dict:
    """Paths of the FFmpeg shared objects shipped inside the opencv wheel."""
    env = os.environ.get("NES_FFMPEG_LIBS")
    d = env or os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")

    def one(pat):
        m = sorted(glob.glob(os.path.join(d, pat)))
        return m[0] if m else None

    return {"dir": d, "deps": [p for p in (one("libcrypto-*"), one("libssl-*"), one("libdrm-*"), one("libvpx-*"), one("libaom-*")) if p],
            "avutil": one("libavutil-*") or one("libavutil.so*"), "swresample": one("libswresample-*") or one("libswresample.so*"),
            "avcodec": one("libavcodec-*") or one("libavcodec.so*")}


_libs = None


def _load():
    global _libs
    if _libs is not None:
        return _libs
    p = bundled_ffmpeg()
    if not p["avutil"] or not p["avcodec"]:
        raise api.NesGpuError(api.NES_ERR_UNSUPPORTED, "libavutil / libavcodec not found", p["dir"])
    for dep in p["deps"]:
        C.CDLL(dep, C.RTLD_GLOBAL)
    avutil = C.CDLL(p["avutil"], C.RTLD_GLOBAL)
    if p["swresample"]:
        C.CDLL(p["swresample"], C.RTLD_GLOBAL)
    avcodec = C.CDLL(p["avcodec"], C.RTLD_GLOBAL)
    vp = C.c_void_p
    avcodec.avcodec_find_encoder_by_name.restype = vp
    avcodec.avcodec_find_encoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_alloc_context3.restype = vp
    avcodec.avcodec_alloc_context3.argtypes = [vp]
    avcodec.avcodec_open2.argtypes = [vp, vp, vp]
    avcodec.avcodec_free_context.argtypes = [C.POINTER(vp)]
    avcodec.avcodec_send_frame.argtypes = [vp, vp]
    avcodec.avcodec_receive_packet.argtypes = [vp, vp]
    avcodec.av_packet_alloc.restype = vp
    avcodec.av_packet_unref.argtypes = [vp]
    avcodec.av_packet_free.argtypes = [C.POINTER(vp)]
    avcodec.avcodec_version.restype = C.c_uint
    avutil.av_opt_set.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    avutil.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int]

    class Q(C.Structure):
        _fields_ = [("num", C.c_int), ("den", C.c_int)]

    avutil.av_opt_set_q.argtypes = [vp, C.c_char_p, Q, C.c_int]
    _libs = (avutil, avcodec, Q, p)
    return _libs


def wrap_frame(fm: "api.FrameManager", pts: int = 0):
    """FrameManager (YUV420P) -> AVFrame* whose buffers reference the FrameManager's planes (nes_avframe_wrap).
    Returns (frame pointer, keep-alive); free with ``free_frame``."""
    L = api.lib()
    p = bundled_ffmpeg()
    planes = (C.c_void_p * 3)(*fm.data[:3])
    ls = (C.c_int * 3)(*fm.linesize[:3])
    out = C.c_void_p()
    r = L.nes_avframe_wrap(p["avutil"].encode() if p["avutil"] else None, planes, ls, fm.context.width, fm.context.height, AV_PIX_FMT_YUV420P, pts,
                           None, None, C.byref(out))
    if r:
        raise api.NesGpuError(r, "nes_avframe_wrap", api.strerror(r) + " | " + L.nes_avframe_error().decode())
    return out, (planes, ls, fm)


def free_frame(frame):
    api.lib().nes_avframe_free(C.byref(frame))


class Encoder:
    """A libavcodec video encoder opened like AVCodecContextManager::codec_ctx_init (type_managers.cc:47-87) but
    through AVOptions (no FFmpeg headers here)."""

    def __init__(self, name: str, width: int, height: int, bit_rate: int = 400000, fps: int = 30, keyint: int = 250):
        avutil, avcodec, Q, _ = _load()
        self.avutil, self.avcodec = avutil, avcodec
        codec = avcodec.avcodec_find_encoder_by_name(name.encode())
        if not codec:
            raise api.NesGpuError(api.NES_ERR_UNSUPPORTED, "encoder not available in this libavcodec", name)
        self.ctx = C.c_void_p(avcodec.avcodec_alloc_context3(codec))
        for r in (avutil.av_opt_set(self.ctx, b"video_size", b"%dx%d" % (width, height), 0), avutil.av_opt_set(self.ctx, b"pixel_format", b"yuv420p", 0),
                  avutil.av_opt_set_q(self.ctx, b"time_base", Q(1, fps), 0), avutil.av_opt_set_int(self.ctx, b"b", bit_rate, 0),
                  avutil.av_opt_set_int(self.ctx, b"g", keyint, 0)):
            if r < 0:
                raise api.NesGpuError(api.NES_ERR_UNSUPPORTED, "av_opt_set failed", str(r))
        r = avcodec.avcodec_open2(self.ctx, codec, None)
        if r < 0:
            raise api.NesGpuError(api.NES_ERR_UNSUPPORTED, "avcodec_open2 failed", str(r))
        self.pkt = C.c_void_p(avcodec.av_packet_alloc())
        self.name = name

    def send(self, frame) -> int:
        """avcodec_send_frame + drain of the ready packets (receive_packet_handler, encode.cpp:214-246) -> packets."""
        r = self.avcodec.avcodec_send_frame(self.ctx, frame)
        if r < 0 and r != _EAGAIN:
            raise api.NesGpuError(api.NES_ERR_UNSUPPORTED, "avcodec_send_frame failed", str(r))
        n = 0
        while self.avcodec.avcodec_receive_packet(self.ctx, self.pkt) >= 0:
            n += 1
            self.avcodec.av_packet_unref(self.pkt)
        return n

    def close(self):
        if self.ctx:
            self.avcodec.av_packet_free(C.byref(self.pkt))
            self.avcodec.avcodec_free_context(C.byref(self.ctx))
            self.ctx = C.c_void_p()


def time_substitute_encoder(width: int, height: int, n_frames: int = 60, name: str = "mpeg4") -> dict:
    """Frames/s of one encoder thread on converted planes (a moving synthetic YUV420P sequence), wrapped with
    nes_avframe_wrap exactly as the hand-off does."""
    _, avcodec, _, paths = _load()
    enc = Encoder(name, width, height)
    frames = []
    yy, xx = np.mgrid[0:height, 0:api.align32(width)]
    for f in range(4):
        fm = api.FrameManager(api.FrameContext(width, height, "yuv420p"))
        fm.planes[0][:] = ((xx * 3 + yy * 5 + f * 7) & 255).astype(np.uint8)
        fm.planes[1][:] = 128
        fm.planes[2][:] = 128
        frames.append(fm)
    pk = 0
    t0 = time.perf_counter()
    for i in range(n_frames):
        fr, keep = wrap_frame(frames[i % 4], pts=i)
        pk += enc.send(fr)
        free_frame(fr)
    dt = time.perf_counter() - t0
    enc.close()
    v = avcodec.avcodec_version()
    return {"status": "substitute", "encoder": name, "libavcodec": f"{v >> 16}.{(v >> 8) & 255}.{v & 255}", "size": [width, height], "frames": n_frames, "packets": pk,
            "value": n_frames / dt, "unit": "frames/s", "threads": 1,
            "note": "the reference encodes H.264 with libx264 (type_managers.cc:47-87); this image's libavcodec has no H.264 encoder, so its mpeg4 encoder "
                    "stands in, with the reference's defaults (400 kbit/s, GOP 250, 30 fps); planes handed over as ref-counted AVFrames (nes_avframe_wrap). "
                    "A separate downstream stage: not part of `value` / `e2e`."}
