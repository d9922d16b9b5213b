// avframe.cc -- encoder hand-off: converted planes as a REF-COUNTED AVFrame.
//
// The reference builds a throw-away AVFrame per send (FrameManager::AVFrameWrapper / to_avframe,
// /root/reference/include/base/video/type_managers.h:187-239: av_frame_alloc + av_image_alloc -- a buffer that is
// orphaned at once -- + av_image_fill_pointers over the FrameManager's planes) and hands it to avcodec_send_frame
// (src/encode.cpp:136-137,164-165).  The frame is not ref-counted, so libavcodec copies every plane before it
// returns.  Here the planes (pinned memory the D2H copy landed in) are wrapped with av_buffer_create: the encoder
// takes a reference instead of a copy, and the release callback returns the block to its owner when the last
// reference is dropped.
//
// libavutil is bound at run time (dlopen, hand-declared prototypes: this library needs no FFmpeg headers).  AVFrame is
// a public struct whose layout depends on the libavutil major version; the field offsets used here are those of
// libavutil 60 (FFmpeg 8) and are VERIFIED against the loaded library before first use (a frame obtained from
// av_frame_get_buffer must show its buffer where we expect buf[0], its line size where we expect linesize[0]):
// an unknown layout fails with NES_ERR_UNSUPPORTED instead of corrupting memory.  A build with the FFmpeg headers
// (include/nes_gpu_shim.hpp with NES_SHIM_WITH_LIBAV) uses the struct directly and never calls this file.
#include <dirent.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "nes_gpu.h"

namespace {

struct AvBufferRef {  // libavutil/buffer.h (stable since 2013)
  void *buffer;
  uint8_t *data;
  size_t size;
};

struct Layout {
  int data, linesize, extended_data, width, height, format, pts, buf;
};
constexpr Layout kLavu60{0, 64, 96, 104, 108, 116, 136, 184};

struct Api {
  void *h = nullptr;
  unsigned (*version)() = nullptr;
  void *(*frame_alloc)() = nullptr;
  void (*frame_free)(void **) = nullptr;
  int (*frame_get_buffer)(void *, int) = nullptr;
  AvBufferRef *(*buffer_create)(uint8_t *, size_t, void (*)(void *, uint8_t *), void *, int) = nullptr;
  int (*buffer_get_ref_count)(const AvBufferRef *) = nullptr;
  Layout L{};
  int status = NES_ERR_UNSUPPORTED;
  bool tried = false;
  std::string why;
};

// A wheel-bundled libavutil (opencv_python_headless.libs/) names its dependencies by hashed sonames that live next
// to it and carries no RUNPATH: load those siblings first (the same situation as FreeType in text.cc).
void preload_siblings(const char *so_path) {
  const std::string path(so_path);
  const size_t slash = path.rfind('/');
  if (slash == std::string::npos) return;
  const std::string dir = path.substr(0, slash);
  DIR *d = opendir(dir.c_str());
  if (!d) return;
  std::vector<std::string> names;
  while (dirent *e = readdir(d)) names.push_back(e->d_name);
  closedir(d);
  for (const char *prefix : {"libcrypto", "libdrm"})
    for (const std::string &n : names)
      if (n.compare(0, strlen(prefix), prefix) == 0) dlopen((dir + "/" + n).c_str(), RTLD_NOW | RTLD_GLOBAL);
  dlerror();
}

template <class T>
T &at(void *frame, int off) { return *reinterpret_cast<T *>(static_cast<uint8_t *>(frame) + off); }

Api &api(const char *so) {
  static Api a;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  // bound once; a failed attempt is retried when the caller names a library (a later call may know the path)
  if (a.status == NES_OK || (a.tried && !(so && *so))) return a;
  a.tried = true;
  a.why.clear();
  [&] {
    const char *cands[] = {so, getenv("NES_AVUTIL_SO"), "libavutil.so.60", "libavutil.so"};
    for (const char *c : cands) {
      if (!c || !*c) continue;
      preload_siblings(c);
      if ((a.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL))) break;
      if (const char *e = dlerror()) a.why = e;
    }
    if (!a.h) { a.status = NES_ERR_UNSUPPORTED; a.why = "cannot dlopen libavutil: " + a.why; return; }
    a.version = (unsigned (*)())dlsym(a.h, "avutil_version");
    a.frame_alloc = (void *(*)())dlsym(a.h, "av_frame_alloc");
    a.frame_free = (void (*)(void **))dlsym(a.h, "av_frame_free");
    a.frame_get_buffer = (int (*)(void *, int))dlsym(a.h, "av_frame_get_buffer");
    a.buffer_create = (AvBufferRef * (*)(uint8_t *, size_t, void (*)(void *, uint8_t *), void *, int)) dlsym(a.h, "av_buffer_create");
    a.buffer_get_ref_count = (int (*)(const AvBufferRef *))dlsym(a.h, "av_buffer_get_ref_count");
    if (!a.version || !a.frame_alloc || !a.frame_free || !a.frame_get_buffer || !a.buffer_create || !a.buffer_get_ref_count) {
      a.why = "libavutil symbols missing";
      return;
    }
    const unsigned major = a.version() >> 16;
    if (major != 60) { a.why = "AVFrame layout of libavutil " + std::to_string(major) + " is not known to this build"; return; }
    a.L = kLavu60;
    // verify the layout against the live library: a 64x16 YUV420P frame with library-allocated buffers
    void *f = a.frame_alloc();
    if (!f) { a.why = "av_frame_alloc failed"; return; }
    at<int>(f, a.L.width) = 64; at<int>(f, a.L.height) = 16; at<int>(f, a.L.format) = 0 /* AV_PIX_FMT_YUV420P */;
    bool ok = a.frame_get_buffer(f, 32) == 0;
    if (ok) {
      uint8_t *d0 = at<uint8_t *>(f, a.L.data);
      AvBufferRef *b0 = at<AvBufferRef *>(f, a.L.buf);
      ok = d0 != nullptr && b0 != nullptr && at<int>(f, a.L.linesize) >= 64 && at<int>(f, a.L.linesize + 4) >= 32 &&
           at<uint8_t **>(f, a.L.extended_data) == &at<uint8_t *>(f, a.L.data);
      // buf[0] must be a live AVBufferRef that owns plane 0 (its data pointer is at or below data[0], inside the block)
      ok = ok && b0->data != nullptr && b0->data <= d0 && d0 < b0->data + b0->size && a.buffer_get_ref_count(b0) == 1;
    }
    a.frame_free(&f);
    if (!ok) { a.why = "AVFrame layout check failed against the loaded libavutil"; return; }
    a.status = NES_OK;
  }();
  return a;
}

struct Release {  // one per wrapped frame, shared by its plane buffers
  std::atomic<int> refs;
  void (*fn)(void *, uint8_t *);
  void *opaque;
  uint8_t *base;
};
void plane_free(void *opaque, uint8_t *) {
  Release *r = static_cast<Release *>(opaque);
  if (r->refs.fetch_sub(1) == 1) {
    if (r->fn) r->fn(r->opaque, r->base);
    delete r;
  }
}

}  // namespace

extern "C" {

int nes_avframe_wrap(const char *avutil_so, uint8_t *const planes[3], const int linesize[3], int width, int height, int av_pix_fmt, int64_t pts,
                     void (*release)(void *opaque, uint8_t *base), void *opaque, void **out_frame) {
  if (!planes || !linesize || !out_frame || width <= 0 || height <= 0) return NES_ERR_INVALID_ARG;
  *out_frame = nullptr;
  Api &a = api(avutil_so);
  if (a.status != NES_OK) return a.status;
  int n_planes = 0;
  while (n_planes < 3 && planes[n_planes]) n_planes++;
  if (n_planes == 0) return NES_ERR_INVALID_ARG;
  void *f = a.frame_alloc();
  if (!f) return NES_ERR_NO_MEMORY;
  Release *r = new (std::nothrow) Release{{n_planes}, release, opaque, planes[0]};
  if (!r) { a.frame_free(&f); return NES_ERR_NO_MEMORY; }
  at<int>(f, a.L.width) = width; at<int>(f, a.L.height) = height; at<int>(f, a.L.format) = av_pix_fmt;
  at<int64_t>(f, a.L.pts) = pts;
  for (int p = 0; p < n_planes; p++) {
    const int rows = p == 0 ? height : (height + 1) / 2;  // 4:2:0 chroma planes; a packed / gray frame has one plane
    AvBufferRef *b = a.buffer_create(planes[p], (size_t)linesize[p] * rows, plane_free, r, 0);
    if (!b) {
      // planes already attached are released with the frame; account for the ones that never will be
      for (int k = p; k < n_planes; k++) plane_free(r, nullptr);
      a.frame_free(&f);
      return NES_ERR_NO_MEMORY;
    }
    at<AvBufferRef *>(f, a.L.buf + 8 * p) = b;
    at<uint8_t *>(f, a.L.data + 8 * p) = planes[p];
    at<int>(f, a.L.linesize + 4 * p) = linesize[p];
  }
  *out_frame = f;
  return NES_OK;
}

void nes_avframe_free(void **frame) {
  if (!frame || !*frame) return;
  Api &a = api(nullptr);
  if (a.status == NES_OK) a.frame_free(frame);
}

int nes_avframe_ref_count(void *frame) {
  if (!frame) return NES_ERR_INVALID_ARG;
  Api &a = api(nullptr);
  if (a.status != NES_OK) return a.status;
  AvBufferRef *b = at<AvBufferRef *>(frame, a.L.buf);
  return b ? a.buffer_get_ref_count(b) : 0;
}

const char *nes_avframe_error(void) { return api(nullptr).why.c_str(); }

}  // extern "C"
