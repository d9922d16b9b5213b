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
//   * RenderedFrame is built from the wire bytes (zero-copy unpack) instead of a parsed
//     nesproto::RenderedFrame; the accessors index()/is_left()/camera matrix come from the same
//     fields.  With protobuf available, pass msg.frame().data() etc. to the pointer constructor.
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
#include <libavutil/pixfmt.h>
}
#else
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
    }
    return m_s;
  }
  const void *atlas_owner = nullptr;  // RenderTextContext whose font is loaded in this session
  ~ThreadSession() { nes_gpu_session_destroy(m_s); }

 private:
  nes_gpu_session *m_s = nullptr;
};

}  // namespace nes_shim

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
      if (nes_gpu_host_alloc(total + 32, &p) == NES_OK) {
        m_pinned = true;
      } else if (!(p = std::malloc(total + 32))) {
        throw std::runtime_error{"Failed to allocate frame data."};
      }
      m_data.data[0] = static_cast<uint8_t *>(p);
      if (context.pix_fmt == AV_PIX_FMT_YUV420P) {
        m_data.data[1] = m_data.data[0] + (size_t)m_data.linesize[0] * h;
        m_data.data[2] = m_data.data[1] + (size_t)m_data.linesize[1] * ((h + 1) / 2);
      }
    } else {
      m_should_free_buffer = false;
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

  ~FrameManager() {
    if (m_should_free_buffer && m_data.data[0]) {
      if (m_pinned) nes_gpu_host_free(m_data.data[0]);
      else std::free(m_data.data[0]);
    }
  }

 private:
  FrameData m_data;
  FrameContext m_context;
  std::vector<TextRun> m_runs;  // overlays queued by RenderTextContext::render_string_to_frame
  bool m_should_free_buffer = true;
  bool m_pinned = false;
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
  // `message` must outlive the object (the reference copies it; here it is borrowed).
  RenderedFrame(const uint8_t *message, size_t len, bool has_length_prefix, AVPixelFormat pix_fmt_scene, AVPixelFormat pix_fmt_depth,
                unsigned dst_width, unsigned dst_height)
      : m_fields(unpack(message, len, has_length_prefix)),
        m_source_avframe_scene(types::FrameManager::FrameContext(m_fields.width, m_fields.height, pix_fmt_scene),
                               const_cast<uint8_t *>(message) + m_fields.frame_off),
        m_converted_avframe_scene(types::FrameManager::FrameContext(dst_width, dst_height, AV_PIX_FMT_YUV420P)),
        m_source_avframe_depth(types::FrameManager::FrameContext(m_fields.width, m_fields.height, pix_fmt_depth),
                               const_cast<uint8_t *>(message) + m_fields.depth_off),
        m_converted_avframe_depth(types::FrameManager::FrameContext(dst_width, dst_height, AV_PIX_FMT_YUV420P)),
        m_converted(false) {
    const uint64_t px = (uint64_t)m_fields.width * m_fields.height;
    // the reference trusts camera.width/height (rendered_frame.cc:14-25); a short payload is an
    // out-of-bounds read there, an exception here
    if (m_fields.frame_len < px * nes_shim::bytes_per_pixel(pix_fmt_scene) || m_fields.depth_len < px)
      throw std::runtime_error{"RenderedFrame: payload shorter than width*height"};
  }

  inline void convert_frame() {
    if (m_converted) throw std::runtime_error{"Tried to convert a converted RenderedFrame."};
    types::detail::convert(m_source_avframe_scene, &m_source_avframe_depth, m_converted_avframe_scene, &m_converted_avframe_depth);
    m_converted = true;
  }

  inline uint64_t index() const { return m_fields.index; }
  inline bool is_left() const { return m_fields.is_left != 0; }
  inline const nes_unpacked_frame &get_cam() const { return m_fields; }  // width, height, matrix[n_matrix]
  inline types::FrameManager &source_frame_scene() { return m_source_avframe_scene; }
  inline types::FrameManager &converted_frame_scene() { return m_converted_avframe_scene; }
  inline types::FrameManager &converted_frame_depth() { return m_converted_avframe_depth; }

 private:
  static nes_unpacked_frame unpack(const uint8_t *message, size_t len, bool prefix) {
    nes_unpacked_frame u;
    nes_shim::check(nes_unpack_rendered_frame(message, len, prefix ? 1 : 0, &u), "nes_unpack_rendered_frame");
    return u;
  }
  nes_unpacked_frame m_fields;
  types::FrameManager m_source_avframe_scene;
  types::FrameManager m_converted_avframe_scene;
  types::FrameManager m_source_avframe_depth;
  types::FrameManager m_converted_avframe_depth;
  bool m_converted;
};

#endif  // NES_GPU_SHIM_HPP_
