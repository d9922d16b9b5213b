// nes_gpu_shim.hpp -- header-only C++ mirror of the reference's hot-path classes on top of
// the C ABI (nes_gpu.h), so that the reference's process_frame_thread (src/encode.cpp:42-119)
// and socket_client_thread (src/server.cpp:170-201) keep their call sequence:
//
//     etctx->render_string_to_frame(frame->source_frame_scene(), RENDER_POSITION_CENTER, text);   x4
//     frame->convert_frame();
//     encode_queue->insert(frame_index, std::move(frame));
//
// Classes and the reference declarations they mirror (same names, same argument meaning, same
// exceptions):
//     types::FrameManager        include/base/video/type_managers.h:159-247, type_managers.cc:116-141
//     types::SwsContextManager   include/base/video/type_managers.h:253-264, type_managers.cc:143-155
//     RenderTextContext          include/base/video/render_text.h:15-39,    render_text.cc:10-113
//     RenderedFrame              include/base/video/rendered_frame.h:15-69, rendered_frame.cc:5-27
//
// Differences a maintainer should know (INTEGRATION.md has the full list):
//   * render_string_to_frame does not touch the host pixels: it queues the run on the frame and
//     the stamp happens on the device copy inside convert_frame(), in call order, before the
//     colour conversion -- exactly the order of encode.cpp:76-98.  Nobody reads the source
//     frame after convert_frame() in the reference, so the result is the same.
//   * RenderedFrame has the reference's constructor (a parsed nesproto::RenderedFrame + the two codec managers,
//     rendered_frame.h:17-20) and a zero-copy one from the wire bytes.  Without protobuf (NES_SHIM_WITH_PROTOBUF
//     undefined) a minimal nesproto::Camera / nesproto::RenderedFrame with the generated accessors
//     (index(), is_left(), camera().matrix(), frame(), depth(), ParseFromString) stands in, parsed by
//     nes_unpack_rendered_frame.
//   * FrameManager::to_avframe() returns a REF-COUNTED AVFrame over the (pinned) planes: avcodec_send_frame
//     takes a reference instead of copying, nothing leaks (the reference's wrapper orphans an av_image_alloc
//     block per call, type_managers.h:201-218).
//   * every thread that converts gets its own nes_gpu_session (lazily, device from
//     NES_GPU_DEVICE or nes_shim::set_thread_device); errors become std::runtime_error like
//     the reference's.
#ifndef NES_GPU_SHIM_HPP_
#define NES_GPU_SHIM_HPP_

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "nes_gpu.h"

#ifdef NES_SHIM_WITH_LIBAV
extern "C" {
#include <libavutil/buffer.h>
#include <libavutil/frame.h>
#include <libavutil/pixfmt.h>
}
#else
struct AVFrame;  // opaque here: created and freed by libavutil through nes_avframe_wrap / nes_avframe_free
// the few AVPixelFormat values the path uses (libavutil/pixfmt.h numbering)
enum AVPixelFormat {
  AV_PIX_FMT_NONE = -1,
  AV_PIX_FMT_YUV420P = 0,
  AV_PIX_FMT_RGB24 = 2,
  AV_PIX_FMT_BGR24 = 3,
  AV_PIX_FMT_GRAY8 = 8,
  AV_PIX_FMT_ARGB = 25,
  AV_PIX_FMT_RGBA = 26,
  AV_PIX_FMT_ABGR = 27,
  AV_PIX_FMT_BGRA = 28
};
#define AV_NUM_DATA_POINTERS 8
#endif

namespace nes_shim {

inline void check(int st, const char *what, nes_gpu_session *s = nullptr) {
  if (st == NES_OK) return;
  std::string msg = std::string(what) + ": " + nes_gpu_strerror(st);
  if (s && st == NES_ERR_CUDA) msg += std::string(" (") + nes_gpu_session_error(s) + ")";
  throw std::runtime_error{msg};
}

inline int to_nes_fmt(AVPixelFormat f) {
  switch (f) {
    case AV_PIX_FMT_RGB24: return NES_PIX_RGB24;
    case AV_PIX_FMT_BGR24: return NES_PIX_BGR24;
    case AV_PIX_FMT_RGBA: return NES_PIX_RGBA;
    case AV_PIX_FMT_BGRA: return NES_PIX_BGRA;
    case AV_PIX_FMT_ARGB: return NES_PIX_ARGB;
    case AV_PIX_FMT_ABGR: return NES_PIX_ABGR;
    default: throw std::runtime_error{"nes_shim: unsupported source pixel format"};
  }
}
inline int bytes_per_pixel(AVPixelFormat f) {
  return f == AV_PIX_FMT_GRAY8 ? 1 : (f == AV_PIX_FMT_RGB24 || f == AV_PIX_FMT_BGR24) ? 3 : 4;
}

struct SessionLimits {
  int max_width = 7680, max_height = 4320, max_sources = 1, ring_depth = 3;
};
inline SessionLimits &limits() { static SessionLimits l; return l; }
inline int &thread_device() {
  static thread_local int dev = std::getenv("NES_GPU_DEVICE") ? std::atoi(std::getenv("NES_GPU_DEVICE")) : 0;
  return dev;
}
inline void set_thread_device(int d) { thread_device() = d; }

// Process-wide multiplexer (nes_gpu_mux), one per GPU: with nes_shim::enable_mux() (or NES_GPU_MUX=1) the sessions of all
// converting threads -- both eyes, every client session of the process -- hand their frames to one dispatcher per GPU,
// which launches whatever is ready together (BASELINE config 4).  Off by default: one session per thread, one launch per frame.
inline bool &mux_enabled() {
  static bool on = std::getenv("NES_GPU_MUX") != nullptr && std::atoi(std::getenv("NES_GPU_MUX")) != 0;
  return on;
}
inline void enable_mux(bool on = true) { mux_enabled() = on; }
inline nes_gpu_mux *device_mux(int device) {
  static std::mutex mu;
  static std::vector<std::pair<int, nes_gpu_mux *>> muxes;  // live for the process (sessions must go first)
  std::lock_guard<std::mutex> lk(mu);
  for (auto &m : muxes)
    if (m.first == device) return m.second;
  nes_gpu_mux *m = nullptr;
  check(nes_gpu_mux_create(device, 0, &m), "nes_gpu_mux_create");
  muxes.emplace_back(device, m);
  return m;
}

// One session per converting thread (the reference runs one process_frame_thread per eye).
class ThreadSession {
 public:
  static ThreadSession &get() {
    static thread_local ThreadSession s;
    return s;
  }
  nes_gpu_session *handle() {
    if (!m_s) {
      nes_gpu_cfg cfg{thread_device(), limits().max_width, limits().max_height, limits().max_sources, limits().ring_depth, 0};
      check(nes_gpu_session_create(&cfg, &m_s), "nes_gpu_session_create");
      if (mux_enabled()) check(nes_gpu_mux_attach(device_mux(cfg.device), m_s), "nes_gpu_mux_attach");
    }
    return m_s;
  }
  const void *atlas_owner = nullptr;  // RenderTextContext whose font is loaded in this session
  ~ThreadSession() { nes_gpu_session_destroy(m_s); }

 private:
  nes_gpu_session *m_s = nullptr;
};

}  // namespace nes_shim

#ifndef NES_SHIM_WITH_PROTOBUF
// Stand-ins for the classes protoc generates from proto/nes.proto:4-25 (only what the hot path touches).
namespace nesproto {
class Camera {
 public:
  bool is_left() const { return m_is_left; }
  uint32_t width() const { return m_width; }
  uint32_t height() const { return m_height; }
  const std::vector<float> &matrix() const { return m_matrix; }
  int matrix_size() const { return (int)m_matrix.size(); }
  void set_is_left(bool v) { m_is_left = v; }
  void set_width(uint32_t v) { m_width = v; }
  void set_height(uint32_t v) { m_height = v; }
  void add_matrix(float v) { m_matrix.push_back(v); }

 private:
  bool m_is_left = false;
  uint32_t m_width = 0, m_height = 0;
  std::vector<float> m_matrix;
};
class RenderedFrame {
 public:
  uint64_t index() const { return m_index; }
  bool is_left() const { return m_is_left; }
  const Camera &camera() const { return m_camera; }
  Camera *mutable_camera() { return &m_camera; }
  const std::string &frame() const { return m_frame; }
  const std::string &depth() const { return m_depth; }
  void set_index(uint64_t v) { m_index = v; }
  void set_is_left(bool v) { m_is_left = v; }
  void set_frame(std::string v) { m_frame = std::move(v); }
  void set_depth(std::string v) { m_depth = std::move(v); }
  // server.cpp:175: the message without the 8-byte length prefix
  bool ParseFromArray(const void *data, int size) {
    nes_unpacked_frame u;
    if (nes_unpack_rendered_frame(static_cast<const uint8_t *>(data), (uint64_t)size, 0, &u) != NES_OK) return false;
    m_index = u.index; m_is_left = u.is_left != 0;
    m_camera = Camera();
    m_camera.set_is_left(u.cam_is_left != 0); m_camera.set_width((uint32_t)u.width); m_camera.set_height((uint32_t)u.height);
    for (int i = 0; i < u.n_matrix; i++) m_camera.add_matrix(u.matrix[i]);
    const char *p = static_cast<const char *>(data);
    m_frame.assign(p + u.frame_off, u.frame_len);
    m_depth.assign(p + u.depth_off, u.depth_len);
    return true;
  }
  bool ParseFromString(const std::string &s) { return ParseFromArray(s.data(), (int)s.size()); }

 private:
  uint64_t m_index = 0;
  bool m_is_left = false;
  Camera m_camera;
  std::string m_frame, m_depth;
};
}  // namespace nesproto
#endif

namespace types {

// types::FrameManager (type_managers.h:159-247)
class FrameManager {
 public:
  static constexpr unsigned kBufferSizeAlignValueBytes = 32;

  struct FrameData {
    uint8_t *data[AV_NUM_DATA_POINTERS] = {0};
    int linesize[AV_NUM_DATA_POINTERS] = {0};
  };
  struct FrameContext {
    FrameContext(unsigned width, unsigned height, AVPixelFormat pix_fmt) : width(width), height(height), pix_fmt(pix_fmt) {}
    // FrameContext(types::AVCodecContextManager::CodecInfoProvider) of the reference (type_managers.h:176-179):
    // anything whose operator-> yields width / height / pix_fmt
    template <class CodecInfoProvider, class = decltype(std::declval<CodecInfoProvider &>()->width)>
    FrameContext(CodecInfoProvider codecinfo) : width(codecinfo->width), height(codecinfo->height), pix_fmt(codecinfo->pix_fmt) {}
    unsigned width;
    unsigned height;
    AVPixelFormat pix_fmt;
  };
  struct TextRun {
    int position;
    std::string content;
  };

  // buffer == nullptr: own the planes, laid out like av_image_alloc(..., align 32)
  // (type_managers.cc:119-121) but in pinned memory so the D2H copy lands in them directly.
  // buffer != nullptr: borrow it with tight line sizes (type_managers.cc:127-133).
  FrameManager(FrameContext context, uint8_t *buffer = nullptr) : m_context(context) {
    const unsigned w = context.width, h = context.height;
    auto align32 = [](unsigned v) { return (int)((v + 31u) & ~31u); };
    if (buffer == nullptr) {
      size_t total;
      if (context.pix_fmt == AV_PIX_FMT_YUV420P) {
        const unsigned cw = (w + 1) / 2, ch = (h + 1) / 2;
        m_data.linesize[0] = align32(w); m_data.linesize[1] = m_data.linesize[2] = align32(cw);
        total = (size_t)m_data.linesize[0] * h + 2 * (size_t)m_data.linesize[1] * ch;
      } else {
        m_data.linesize[0] = align32(w * nes_shim::bytes_per_pixel(context.pix_fmt));
        total = (size_t)m_data.linesize[0] * h;
      }
      void *p = nullptr;
      bool pinned = false;
      if (nes_gpu_host_alloc(total + 32, &p) == NES_OK) {
        pinned = true;
      } else if (!(p = std::malloc(total + 32))) {
        throw std::runtime_error{"Failed to allocate frame data."};
      }
      // the block is shared with every AVFrame made from it (to_avframe): freed when the last owner lets go
      m_block = std::shared_ptr<uint8_t>(static_cast<uint8_t *>(p), [pinned](uint8_t *q) { if (pinned) nes_gpu_host_free(q); else std::free(q); });
      m_data.data[0] = static_cast<uint8_t *>(p);
      if (context.pix_fmt == AV_PIX_FMT_YUV420P) {
        m_data.data[1] = m_data.data[0] + (size_t)m_data.linesize[0] * h;
        m_data.data[2] = m_data.data[1] + (size_t)m_data.linesize[1] * ((h + 1) / 2);
      }
    } else {
      if (context.pix_fmt == AV_PIX_FMT_YUV420P) {
        m_data.linesize[0] = (int)w; m_data.linesize[1] = m_data.linesize[2] = (int)((w + 1) / 2);
        m_data.data[1] = buffer + (size_t)w * h;
        m_data.data[2] = m_data.data[1] + (size_t)m_data.linesize[1] * ((h + 1) / 2);
      } else {
        m_data.linesize[0] = (int)(w * nes_shim::bytes_per_pixel(context.pix_fmt));
      }
      m_data.data[0] = buffer;
    }
  }
  FrameManager(const FrameManager &) = delete;
  FrameManager &operator=(const FrameManager &) = delete;

  inline FrameContext &context() { return m_context; }
  inline FrameData &data() { return m_data; }
  inline std::vector<TextRun> &text_runs() { return m_runs; }

  // FrameManager::AVFrameWrapper (type_managers.h:187-231): an AVFrame over this frame's planes, freed with the
  // wrapper.  Unlike the reference's it is ref-counted (av_buffer_create over the shared block), so
  // avcodec_send_frame keeps a reference instead of copying the planes, and it allocates nothing it then orphans.
  class AVFrameWrapper {
   public:
    AVFrameWrapper(FrameData &data, FrameContext &context, std::shared_ptr<uint8_t> block) {
      auto *keep = new std::shared_ptr<uint8_t>(std::move(block));  // dropped by the release callback
      auto release = [](void *opaque, uint8_t *) { delete static_cast<std::shared_ptr<uint8_t> *>(opaque); };
#ifdef NES_SHIM_WITH_LIBAV
      m_avframe = av_frame_alloc();
      if (m_avframe == nullptr) { delete keep; throw std::runtime_error{"Failed to allocate AVFrame."}; }
      m_avframe->format = context.pix_fmt; m_avframe->width = (int)context.width; m_avframe->height = (int)context.height;
      const int planes = context.pix_fmt == AV_PIX_FMT_YUV420P ? 3 : 1;
      auto *count = new std::pair<int, std::shared_ptr<uint8_t> *>(planes, keep);
      for (int p = 0; p < planes; p++) {
        const int rows = p == 0 ? (int)context.height : ((int)context.height + 1) / 2;
        m_avframe->buf[p] = av_buffer_create(data.data[p], (size_t)data.linesize[p] * rows,
                                             [](void *o, uint8_t *) { auto *c = static_cast<std::pair<int, std::shared_ptr<uint8_t> *> *>(o); if (--c->first == 0) { delete c->second; delete c; } },
                                             count, 0);
        m_avframe->data[p] = data.data[p]; m_avframe->linesize[p] = data.linesize[p];
      }
      (void)release;
#else
      void *f = nullptr;
      uint8_t *const planes[3] = {data.data[0], data.data[1], data.data[2]};
      const int st = nes_avframe_wrap(nullptr, planes, data.linesize, (int)context.width, (int)context.height, (int)context.pix_fmt, 0, release, keep, &f);
      if (st != NES_OK) {
        delete keep;
        throw std::runtime_error{std::string("AVFrameWrapper: Failed to allocate AVFrame data: ") + nes_gpu_strerror(st) + " (" + nes_avframe_error() + ")"};
      }
      m_avframe = static_cast<AVFrame *>(f);
#endif
    }
    AVFrameWrapper(const AVFrameWrapper &) = delete;
    AVFrameWrapper(AVFrameWrapper &&o) noexcept : m_avframe(o.m_avframe) { o.m_avframe = nullptr; }
    inline AVFrame *get() { return m_avframe; }
    ~AVFrameWrapper() {
#ifdef NES_SHIM_WITH_LIBAV
      av_frame_free(&m_avframe);
#else
      void *f = m_avframe;
      nes_avframe_free(&f);
#endif
    }

   private:
    AVFrame *m_avframe = nullptr;
  };
  inline AVFrameWrapper to_avframe() {
    if (!m_block) throw std::runtime_error{"to_avframe: the frame borrows its buffer (only converted frames are handed to the encoder)"};
    return AVFrameWrapper(m_data, m_context, m_block);
  }

  ~FrameManager() {}

 private:
  FrameData m_data;
  FrameContext m_context;
  std::vector<TextRun> m_runs;  // overlays queued by RenderTextContext::render_string_to_frame
  std::shared_ptr<uint8_t> m_block;  // owning mode: the planes' block (null when the buffer is borrowed)
};

namespace detail {
inline void fill_out(nes_frame_out &fo, FrameManager *scene, FrameManager *depth) {
  std::memset(&fo, 0, sizeof(fo));
  fo.width = (int)scene->context().width; fo.height = (int)scene->context().height; fo.mem = NES_MEM_HOST;
  for (int p = 0; p < 3; p++) {
    fo.scene[p] = scene->data().data[p]; fo.scene_linesize[p] = scene->data().linesize[p];
    if (depth) { fo.depth[p] = depth->data().data[p]; fo.depth_linesize[p] = depth->data().linesize[p]; }
  }
}
// scene (+ optional depth) conversion with the scene's queued text runs
inline void convert(FrameManager &scene_src, FrameManager *depth_src, FrameManager &scene_dst, FrameManager *depth_dst) {
  nes_gpu_session *s = nes_shim::ThreadSession::get().handle();
  const int w = (int)scene_src.context().width, h = (int)scene_src.context().height;
  nes_frame_in fi;
  std::memset(&fi, 0, sizeof(fi));
  fi.n_sources = 1; fi.pix_fmt = nes_shim::to_nes_fmt(scene_src.context().pix_fmt); fi.width = w; fi.height = h; fi.mem = NES_MEM_HOST;
  fi.src[0].rgb = scene_src.data().data[0]; fi.src[0].rgb_stride = scene_src.data().linesize[0];
  fi.src[0].rgb_bytes = (uint64_t)scene_src.data().linesize[0] * h;
  if (depth_src) {
    fi.src[0].depth = depth_src->data().data[0]; fi.src[0].depth_stride = depth_src->data().linesize[0];
    fi.src[0].depth_bytes = (uint64_t)depth_src->data().linesize[0] * h;
  }
  nes_frame_out fo;
  fill_out(fo, &scene_dst, depth_src ? depth_dst : nullptr);
  std::vector<nes_text_run> runs;
  for (auto &r : scene_src.text_runs()) runs.push_back(nes_text_run{r.position, (int32_t)r.content.size(), r.content.data()});
  nes_shim::check(nes_gpu_convert(s, &fi, runs.data(), (int)runs.size(), &fo), "nes_gpu_convert", s);
}
}  // namespace detail

// types::SwsContextManager (type_managers.cc:143-155): converts on construction.
class SwsContextManager {
 public:
  SwsContextManager(FrameManager &source, FrameManager &dest) {
    if (dest.context().pix_fmt != AV_PIX_FMT_YUV420P) throw std::runtime_error{"Failed to allocate sws_context."};
    if (source.context().pix_fmt == AV_PIX_FMT_GRAY8) {
      // the ABI converts depth beside a scene: a lone GRAY8 frame rides with a blank RGB24 scene
      const unsigned w = source.context().width, h = source.context().height;
      FrameManager blank(FrameManager::FrameContext(w, h, AV_PIX_FMT_RGB24));
      std::memset(blank.data().data[0], 0, (size_t)blank.data().linesize[0] * h);
      FrameManager sink(FrameManager::FrameContext(dest.context().width, dest.context().height, AV_PIX_FMT_YUV420P));
      detail::convert(blank, &source, sink, &dest);
    } else {
      detail::convert(source, nullptr, dest, nullptr);
    }
  }
  ~SwsContextManager() {}
};

}  // namespace types

// RenderTextContext (render_text.h:15-39)
class RenderTextContext {
 public:
  enum RenderPosition {
    RENDER_POSITION_LEFT_TOP,
    RENDER_POSITION_LEFT_BOTTOM,
    RENDER_POSITION_RIGHT_TOP,
    RENDER_POSITION_RIGHT_BOTTOM,
    RENDER_POSITION_CENTER
  };

  // Rasterises the font once (FT_Init_FreeType / FT_New_Face / FT_Set_Char_Size(0, 20*64, 0, 0) /
  // FT_Load_Char(FT_LOAD_RENDER), render_text.cc:12-32,88) into a host glyph table.
  RenderTextContext(std::string font_location, std::string freetype_so = std::string()) : m_coverage(1 << 20) {
    uint64_t used = 0;
    const int st = nes_font_rasterise(freetype_so.empty() ? nullptr : freetype_so.c_str(), font_location.c_str(), m_glyphs,
                                      m_coverage.data(), m_coverage.size(), &used);
    if (st != NES_OK) throw std::runtime_error{std::string("EncodeTextContext: Failed to init font face: ") + nes_gpu_strerror(st)};
  }

  // Same signature as the reference; the stamp is deferred to the device copy (see file header).
  void render_string_to_frame(types::FrameManager &frame, RenderTextContext::RenderPosition opt, std::string content) {
    nes_shim::ThreadSession &ts = nes_shim::ThreadSession::get();
    if (ts.atlas_owner != this) {  // first use on this thread: upload the atlas to its session
      nes_shim::check(nes_gpu_atlas_set(ts.handle(), m_glyphs, 256), "nes_gpu_atlas_set", ts.handle());
      ts.atlas_owner = this;
    }
    frame.text_runs().push_back(types::FrameManager::TextRun{(int)opt, std::move(content)});
  }

 private:
  nes_glyph m_glyphs[256];
  std::vector<uint8_t> m_coverage;
};

// RenderedFrame (rendered_frame.h:15-69) built from the wire bytes the renderer sent
// (server.cpp:91-112 framing + nes.proto:18-25), without copying the payload.
class RenderedFrame {
 public:
  // The reference's constructor (rendered_frame.h:17-20, rendered_frame.cc:5-27): the parsed message is copied
  // into the object, the source frames borrow its two byte strings, the destination frames take their size from
  // the codec managers (anything with get_codec_info()-> width / height / pix_fmt).
  template <class CodecContextManager>
  RenderedFrame(nesproto::RenderedFrame frame, AVPixelFormat pix_fmt_scene, AVPixelFormat pix_fmt_depth,
                std::shared_ptr<CodecContextManager> ctxmgr_scene, std::shared_ptr<CodecContextManager> ctxmgr_depth)
      : m_frame_response(std::move(frame)),
        m_source_avframe_scene(types::FrameManager::FrameContext(m_frame_response.camera().width(), m_frame_response.camera().height(), pix_fmt_scene),
                               (uint8_t *)m_frame_response.frame().data()),
        m_converted_avframe_scene(types::FrameManager::FrameContext(ctxmgr_scene->get_codec_info())),
        m_source_avframe_depth(types::FrameManager::FrameContext(m_frame_response.camera().width(), m_frame_response.camera().height(), pix_fmt_depth),
                               (uint8_t *)m_frame_response.depth().data()),
        m_converted_avframe_depth(types::FrameManager::FrameContext(ctxmgr_depth->get_codec_info())),
        m_converted(false) {
    check_payload(m_frame_response.frame().size(), m_frame_response.depth().size(), pix_fmt_scene);
  }

  // Zero-copy variant: built straight from the wire bytes (server.cpp:91-112 framing + nes.proto:18-25); `message`
  // must outlive the object (the payload is borrowed, not copied; scalar fields and the camera are extracted).
  RenderedFrame(const uint8_t *message, size_t len, bool has_length_prefix, AVPixelFormat pix_fmt_scene, AVPixelFormat pix_fmt_depth,
                unsigned dst_width, unsigned dst_height)
      : m_frame_response(scalars(message, len, has_length_prefix)),
        m_source_avframe_scene(types::FrameManager::FrameContext(m_frame_response.camera().width(), m_frame_response.camera().height(), pix_fmt_scene),
                               const_cast<uint8_t *>(message) + m_wire.frame_off),
        m_converted_avframe_scene(types::FrameManager::FrameContext(dst_width, dst_height, AV_PIX_FMT_YUV420P)),
        m_source_avframe_depth(types::FrameManager::FrameContext(m_frame_response.camera().width(), m_frame_response.camera().height(), pix_fmt_depth),
                               const_cast<uint8_t *>(message) + m_wire.depth_off),
        m_converted_avframe_depth(types::FrameManager::FrameContext(dst_width, dst_height, AV_PIX_FMT_YUV420P)),
        m_converted(false) {
    check_payload(m_wire.frame_len, m_wire.depth_len, pix_fmt_scene);
  }

  inline void convert_frame() {
    if (m_converted) throw std::runtime_error{"Tried to convert a converted RenderedFrame."};
    types::detail::convert(m_source_avframe_scene, &m_source_avframe_depth, m_converted_avframe_scene, &m_converted_avframe_depth);
    m_converted = true;
  }

  inline uint64_t index() const { return m_frame_response.index(); }
  inline bool is_left() const { return m_frame_response.is_left(); }
  inline const nesproto::Camera &get_cam() const { return m_frame_response.camera(); }
  inline types::FrameManager &source_frame_scene() { return m_source_avframe_scene; }
  inline types::FrameManager &converted_frame_scene() { return m_converted_avframe_scene; }
  inline types::FrameManager &converted_frame_depth() { return m_converted_avframe_depth; }

 private:
  // scalar fields + camera of a wire message (the two byte strings stay where they are: m_wire keeps their offsets)
  nesproto::RenderedFrame scalars(const uint8_t *message, size_t len, bool prefix) {
    nes_shim::check(nes_unpack_rendered_frame(message, len, prefix ? 1 : 0, &m_wire), "nes_unpack_rendered_frame");
    nesproto::RenderedFrame f;
    f.set_index(m_wire.index); f.set_is_left(m_wire.is_left != 0);
    nesproto::Camera *c = f.mutable_camera();
    c->set_is_left(m_wire.cam_is_left != 0); c->set_width((uint32_t)m_wire.width); c->set_height((uint32_t)m_wire.height);
    for (int i = 0; i < m_wire.n_matrix; i++) c->add_matrix(m_wire.matrix[i]);
    return f;
  }
  void check_payload(uint64_t frame_len, uint64_t depth_len, AVPixelFormat pix_fmt_scene) const {
    const uint64_t px = (uint64_t)m_frame_response.camera().width() * m_frame_response.camera().height();
    // the reference trusts camera.width/height (rendered_frame.cc:14-25); a short payload is an
    // out-of-bounds read there, an exception here
    if (frame_len < px * nes_shim::bytes_per_pixel(pix_fmt_scene) || depth_len < px)
      throw std::runtime_error{"RenderedFrame: payload shorter than width*height"};
  }
  nes_unpacked_frame m_wire{};             // declared before m_frame_response: scalars() fills it
  nesproto::RenderedFrame m_frame_response;
  types::FrameManager m_source_avframe_scene;
  types::FrameManager m_converted_avframe_scene;
  types::FrameManager m_source_avframe_depth;
  types::FrameManager m_converted_avframe_depth;
  bool m_converted;
};

#endif  // NES_GPU_SHIM_HPP_
