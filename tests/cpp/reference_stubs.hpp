// tests/cpp/reference_stubs.hpp -- TEST INFRASTRUCTURE.  The parts of the reference that sit either side of the
// hot path and are out of scope here: minimal stand-ins with the reference's
// interfaces so that its own encode.cpp text compiles against include/nes_gpu_shim.hpp:
//   FrameQueue / FrameMap        include/base/video/frame_queue.h:17-35, frame_map.h:17-42
//   LockTimeout                  include/base/exceptions/lock_timeout.h:9-13
//   ScopedTimer                  include/base/scoped_timer.h:9-23
//   tlog::info() / error()       include/base/logging.h:6-8 (tinylogger streams)
//   types::AVCodecContextManager include/base/video/type_managers.h:69-141 (CodecInitInfo, get_codec_info, send_frame)
#ifndef NES_TEST_REFERENCE_STUBS_HPP_
#define NES_TEST_REFERENCE_STUBS_HPP_
#include <atomic>
#include <cerrno>
#include <chrono>
#include <exception>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <shared_mutex>
#include <sstream>
#include <string>

#include "nes_gpu_shim.hpp"

#ifndef AVERROR
#define AVERROR(e) (-(e))
#endif
#ifndef AVERROR_EOF
#define AVERROR_EOF (-541478725)
#endif

class LockTimeout : public std::exception {
  virtual const char *what() const throw() { return "Waiting for lock timed out."; }
};

class ScopedTimer {
 public:
  using clock = std::chrono::steady_clock;
  using time_format = std::chrono::milliseconds;
  ScopedTimer() : _start(clock::now()) {}
  inline time_format elapsed() { return std::chrono::duration_cast<time_format>(clock::now() - _start); }

 private:
  std::chrono::time_point<clock> _start;
};

namespace tlog {
struct Line {
  std::ostringstream s;
  template <class T>
  Line &operator<<(const T &v) { s << v; return *this; }
  ~Line() { std::cerr << s.str() << "\n"; }
};
inline Line info() { return Line(); }
inline Line error() { return Line(); }
}  // namespace tlog

struct AVPacket;

namespace types {
// Only what the hot path and send_frame_thread use: the destination size and send_frame.
class AVCodecContextManager {
 public:
  struct CodecInitInfo {
    AVPixelFormat pix_fmt;
    unsigned width, height;
  };
  class CodecInfoProvider {
   public:
    CodecInfoProvider(CodecInitInfo &info, std::shared_mutex &mutex) : m_info(std::make_shared<CodecInitInfo>(info)), m_lock(mutex) {}
    inline CodecInitInfo *operator->() { return m_info.get(); }

   private:
    std::shared_ptr<CodecInitInfo> m_info;
    std::shared_lock<std::shared_mutex> m_lock;
  };
  AVCodecContextManager(unsigned width, unsigned height) : m_info{AV_PIX_FMT_YUV420P, width, height} {}
  inline CodecInfoProvider get_codec_info() { return CodecInfoProvider{m_info, m_mutex}; }
  // records what the encoder would have been handed
  int send_frame(AVFrame *frm) {
    sent++;
    last_ref_count = nes_avframe_ref_count(frm);
    return 0;
  }
  int receive_packet(AVPacket *) { return AVERROR(EAGAIN); }
  int sent = 0, last_ref_count = 0;

 private:
  CodecInitInfo m_info;
  mutable std::shared_mutex m_mutex;
};
}  // namespace types

// Single-threaded stand-ins: pop() / get_delete() on an empty container raise LockTimeout like the reference
// does after its 1 s wait (frame_queue.cc:25-39, frame_map.cc:26-53) and ask the driver to stop.
class FrameQueue {
 public:
  using element = std::unique_ptr<RenderedFrame>;
  explicit FrameQueue(std::atomic<bool> &stop) : m_stop(stop) {}
  void push(element &&el) { m_queue.push(std::move(el)); }
  element pop() {
    if (m_queue.empty()) { m_stop = true; throw LockTimeout{}; }
    element e = std::move(m_queue.front());
    m_queue.pop();
    return e;
  }

 private:
  std::queue<element> m_queue;
  std::atomic<bool> &m_stop;
};

class FrameMap {
 public:
  using element = std::unique_ptr<RenderedFrame>;
  using keytype = std::uint64_t;
  explicit FrameMap(std::atomic<bool> &stop) : m_stop(stop) {}
  void insert(keytype index, element &&el) { m_map[index] = std::move(el); }
  element get_delete(keytype index) {
    auto it = m_map.find(index);
    if (it == m_map.end()) { m_stop = true; throw LockTimeout{}; }
    element e = std::move(it->second);
    m_map.erase(it);
    return e;
  }
  std::size_t size() const { return m_map.size(); }

 private:
  std::map<keytype, element> m_map;
  std::atomic<bool> &m_stop;
};
#endif
