"""CPU tests of the host side: C++ shim compiles against the C ABI, session sharding across
ranks (world_size 2 over gloo), bench.py's reference arm prints a well-formed line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_process_frame(tmp_path):
    exe = str(tmp_path / "process_frame")
    lib_dir = os.path.join(ROOT, "ngp-encode-server_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "process_frame.cpp"),
                    "-o", exe, "-L", lib_dir, "-l:libnes_gpu.so", f"-Wl,-rpath,{lib_dir}"], check=True)
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(N, O, tmp_path):
    """include/nes_gpu_shim.hpp + the reference's process_frame_thread body compile and link against
    libnes_gpu.so; without a GPU the run must fail with an error (no CPU fallback), after the
    host-only parts (FreeType rasterise, wire unpack) succeeded."""
    exe = build_process_frame(tmp_path)
    if N.device_count() > 0 or N.find_freetype() is None:
        pytest.skip("needs a GPU-less box with FreeType")
    msg = O.pack_rendered_frame(3, True, 64, 32, O.KINITIAL_CAMERA_MATRIX, O.synth_rgb(64, 32).tobytes(), O.synth_depth(64, 32).tobytes())
    (tmp_path / "m.bin").write_bytes(msg)
    r = subprocess.run([exe, str(tmp_path / "m.bin"), os.path.join(ROOT, "tests", "golden", "Aileron-Regular.ttf"), N.find_freetype(), "64", "32", "12:34:56.789", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA error" in r.stderr, (r.returncode, r.stderr)
    # truncated payload: rejected by the shim before any GPU work (the reference would read out of bounds)
    bad = O.pack_rendered_frame(3, True, 64, 32, O.KINITIAL_CAMERA_MATRIX, b"x" * 100, b"y" * 10)
    (tmp_path / "b.bin").write_bytes(bad)
    r = subprocess.run([exe, str(tmp_path / "b.bin"), os.path.join(ROOT, "tests", "golden", "Aileron-Regular.ttf"), N.find_freetype(), "64", "32", "t", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "payload shorter" in r.stderr


REFERENCE_ENCODE_CPP = "/root/reference/src/encode.cpp"


def extract_reference_encode_text(dst_path):
    """The reference's own text of timestamp(), process_frame_thread and send_frame_thread
    (/root/reference/src/encode.cpp:23-212), verbatim, into `dst_path` -- read at test time, never committed."""
    src = open(REFERENCE_ENCODE_CPP).read()
    a, b = src.index("std::string timestamp() {"), src.index("int receive_packet_handler(")
    text = src[a:b]
    assert "void process_frame_thread(" in text and "void send_frame_thread(" in text and "frame->get_cam().matrix()" in text
    assert ".to_avframe().get()" in text
    with open(dst_path, "w") as f:
        f.write(text)


# where the extracted text is parked so that it travels to the GPU box (git-ignored, like oracle/_ref: built from the
# reference's sources where they lie, never part of the history)
REF_TEXT_DIR = os.path.join(ROOT, "tests", "cpp", "_ref")


def reference_text_available() -> bool:
    return os.path.exists(REFERENCE_ENCODE_CPP) or os.path.exists(os.path.join(REF_TEXT_DIR, "encode_text.inc"))


def build_encode_text_driver(tmp_path):
    if os.path.exists(REFERENCE_ENCODE_CPP):
        os.makedirs(REF_TEXT_DIR, exist_ok=True)
        extract_reference_encode_text(os.path.join(REF_TEXT_DIR, "encode_text.inc"))
    exe = str(tmp_path / "encode_text_driver")
    lib_dir = os.path.join(ROOT, "ngp-encode-server_b200")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"), "-I", REF_TEXT_DIR,
                    os.path.join(ROOT, "tests", "cpp", "encode_text_driver.cpp"), "-o", exe, "-L", lib_dir, "-l:libnes_gpu.so", f"-Wl,-rpath,{lib_dir}", "-pthread"],
                   check=True)
    return exe


@pytest.mark.skipif(not os.path.exists(REFERENCE_ENCODE_CPP), reason="the reference tree is only mounted in the build container")
def test_reference_encode_cpp_text_compiles_against_the_shim(N, O, tmp_path):
    """Signature fidelity of the drop-in: the UNMODIFIED text of the reference's process_frame_thread and
    send_frame_thread (src/encode.cpp:23-212) compiles and links against include/nes_gpu_shim.hpp (+ stand-ins for the
    queues / logger / codec manager that are out of scope): get_cam().matrix(), the RenderedFrame constructor taking the
    parsed message and the two codec managers, render_string_to_frame, convert_frame(), to_avframe().get().
    Without a GPU the run stops at the first conversion with a CUDA error (no CPU fallback)."""
    exe = build_encode_text_driver(tmp_path)
    if N.device_count() > 0 or N.find_freetype() is None:
        return
    msg = O.pack_rendered_frame(0, True, 64, 32, O.KINITIAL_CAMERA_MATRIX, O.synth_rgb(64, 32).tobytes(), O.synth_depth(64, 32).tobytes())
    (tmp_path / "m.bin").write_bytes(msg)
    r = subprocess.run([exe, str(tmp_path / "m.bin"), os.path.join(ROOT, "tests", "golden", "Aileron-Regular.ttf"), N.find_freetype(), "64", "32", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA error" in r.stderr, (r.returncode, r.stderr)


def test_avframe_handoff(N):
    """nes_avframe_wrap (replaces FrameManager::AVFrameWrapper / to_avframe, type_managers.h:187-239): the planes of a
    converted frame as a ref-counted AVFrame of the REAL libavutil bundled in this image -- fields where libavutil
    expects them, no copy (the frame's data pointers are the FrameManager's), release callback on the last unref,
    and a real encoder (libavcodec mpeg4; no H.264 encoder exists here) accepts it."""
    import ctypes as C
    import numpy as np
    av = N.avhandoff
    paths = av.bundled_ffmpeg()
    if not paths["avutil"] or not paths["avcodec"]:
        pytest.skip("no bundled FFmpeg in this image")
    w, h = 640, 360
    fm = N.FrameManager(N.FrameContext(w, h, "yuv420p"))
    yy, xx = np.mgrid[0:h, 0:fm.linesize[0]]
    fm.planes[0][:] = ((3 * xx + 5 * yy) & 255).astype(np.uint8)
    fm.planes[1][:] = 100
    fm.planes[2][:] = 160
    L = N.lib()
    released = []
    CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
    cb = CB(lambda opaque, base: released.append(base))
    planes = (C.c_void_p * 3)(*fm.data[:3])
    ls = (C.c_int * 3)(*fm.linesize[:3])
    fr = C.c_void_p()
    r = L.nes_avframe_wrap(paths["avutil"].encode(), planes, ls, w, h, 0, 42, C.cast(cb, C.c_void_p), None, C.byref(fr))
    assert r == 0, L.nes_avframe_error()
    # the public head of AVFrame (libavutil/frame.h): data[8], linesize[8], extended_data, width, height, nb_samples, format
    base = fr.value
    assert [C.c_uint64.from_address(base + 8 * i).value for i in range(3)] == list(fm.data[:3])          # no copy
    assert [C.c_int.from_address(base + 64 + 4 * i).value for i in range(3)] == list(fm.linesize[:3])
    assert C.c_uint64.from_address(base + 96).value == base                                                  # extended_data == data
    assert (C.c_int.from_address(base + 104).value, C.c_int.from_address(base + 108).value, C.c_int.from_address(base + 116).value) == (w, h, 0)
    assert L.nes_avframe_ref_count(fr) == 1
    # libavutil itself agrees: av_frame_clone takes references to the same buffers (no new allocation of planes)
    avutil = C.CDLL(paths["avutil"])
    avutil.av_frame_clone.restype = C.c_void_p
    avutil.av_frame_clone.argtypes = [C.c_void_p]
    avutil.av_frame_free.argtypes = [C.POINTER(C.c_void_p)]
    clone = C.c_void_p(avutil.av_frame_clone(fr))
    assert clone.value and C.c_uint64.from_address(clone.value).value == fm.data[0] and L.nes_avframe_ref_count(fr) == 2
    # a real encoder takes it
    enc = av.Encoder("mpeg4", w, h)
    assert enc.send(fr) >= 0
    enc.close()
    L.nes_avframe_free(C.byref(fr))
    assert fr.value is None and released == []        # the clone still references the planes
    avutil.av_frame_free(C.byref(clone))
    assert released == [fm.data[0]]                   # last reference gone: the block goes back to its owner, once
    # bad arguments
    assert L.nes_avframe_wrap(None, planes, ls, 0, h, 0, 0, None, None, C.byref(fr)) == N.NES_ERR_INVALID_ARG
    t = av.time_substitute_encoder(320, 180, 8)
    assert t["status"] == "substitute" and t["encoder"] == "mpeg4" and t["packets"] >= 1 and t["value"] > 0


def test_shard_partition(N):
    for world in (1, 2, 4, 8):
        shards = [N.shard.sessions_of_rank(r, world, 64) for r in range(world)]
        assert sorted(s for sh in shards for s in sh) == list(range(64))
        assert all(len(sh) == 64 // world for sh in shards)
        assert all(N.shard.device_for_session(s, world) == r for r, sh in enumerate(shards) for s in sh)
    assert N.shard.aggregate_fps([100, 100], [1.0, 2.0]) == 100.0
    with pytest.raises(ValueError):
        N.shard.sessions_of_rank(2, 2, 64)


WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["NES_ROOT"])
import torch, torch.distributed as dist
import ngp_encode_server_b200 as n
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = n.shard.sessions_of_rank(rank, world, 64)
got = [None] * world
dist.all_gather_object(got, mine)
flat = sorted(s for sh in got for s in sh)
t = n.shard.max_over_ranks(1.0 + rank, dist)     # slowest rank defines the step time
frames = [len(sh) * 10 for sh in got]
fps = n.shard.aggregate_fps(frames, [t] * world)
dist.barrier()
if rank == 0:
    print(json.dumps({"ok": flat == list(range(64)), "t": t, "fps": fps, "world": world}))
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    """The N>1 host path (shard plan, barrier, max-over-ranks) with two CPU processes over gloo."""
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, NES_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731", str(w)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d == {"ok": True, "t": 2.0, "fps": 640 / 2.0, "world": 2}


def test_bench_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", "--workload", "c2_1080p_2src_composite"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"] == "c2_1080p_2src_composite" and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0


def test_gray_luma_dp2a_constants():
    """k_frame_strips computes the GRAY8 range compression with one dp2a per pixel:
    (d*56282 + 1081500) >> 16 must equal libswscale's ((((d<<7)*14071 + 33561472) >> 14) + 64) >> 7
     for every byte, and stay below 2^24 (Y is byte 2 of the sum)."""
    for d in range(256):
        want = ((((d << 7) * 14071 + 33561472) >> 14) + 64) >> 7
        s = d * 56282 + 1081500
        assert s < (1 << 24) and (s >> 16) == want == (d * 219 + 127) // 255 + 16


def test_ingest_ring_fails_loudly_without_gpu(N):
    """The receive ring is page-locked memory for DMA: without a CUDA device it must refuse
    (NES_ERR_CUDA), not hand out pageable memory silently."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(N.NesGpuError) as e:
        N.IngestRing(2, 1 << 20)
    assert e.value.status == N.NES_ERR_CUDA
