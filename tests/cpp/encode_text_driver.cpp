// tests/cpp/encode_text_driver.cpp -- drives the reference's OWN process_frame_thread and send_frame_thread
// (the text of /root/reference/src/encode.cpp from `std::string timestamp()` up to `receive_packet_handler`,
// extracted at test time into encode_text.inc, never committed) compiled against include/nes_gpu_shim.hpp and the
// stand-ins of reference_stubs.hpp.
//
//   encode_text_driver <wire message file> <font.ttf> <freetype.so> <dst_w> <dst_h> <out prefix>
//
// Builds a RenderedFrame with the reference's constructor (server.cpp:175,193-194), lets process_frame_thread take
// it through the four overlays and convert_frame(), then send_frame_thread hand both converted frames to the
// (stand-in) encoder through to_avframe().  The timestamp overlay is wall-clock text, so the planes are compared by
// the Python test only away from it.
#include <atomic>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>

#include "reference_stubs.hpp"

#include "encode_text.inc"

static void dump(types::FrameManager &f, const std::string &path) {
  std::ofstream o(path, std::ios::binary);
  const unsigned w = f.context().width, h = f.context().height;
  for (unsigned y = 0; y < h; y++) o.write((const char *)f.data().data[0] + (size_t)y * f.data().linesize[0], w);
  for (int p = 1; p < 3; p++)
    for (unsigned y = 0; y < (h + 1) / 2; y++) o.write((const char *)f.data().data[p] + (size_t)y * f.data().linesize[p], (w + 1) / 2);
}

int main(int argc, char **argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: see file header\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    std::string wire((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    auto etctx = std::make_shared<RenderTextContext>(argv[2], argv[3]);
    const unsigned dw = (unsigned)std::atoi(argv[4]), dh = (unsigned)std::atoi(argv[5]);
    auto ctx_scene = std::make_shared<types::AVCodecContextManager>(dw, dh);
    auto ctx_depth = std::make_shared<types::AVCodecContextManager>(dw, dh);
    std::atomic<bool> shutdown{false};
    auto fq = std::make_shared<FrameQueue>(shutdown);
    auto fm = std::make_shared<FrameMap>(shutdown);

    // server.cpp:172-194: the message (after its 8-byte length prefix) -> ParseFromString -> RenderedFrame
    nesproto::RenderedFrame frame;
    if (!frame.ParseFromString(wire.substr(8))) { std::fprintf(stderr, "ParseFromString failed\n"); return 3; }
    std::unique_ptr<RenderedFrame> frame_o = std::make_unique<RenderedFrame>(frame, AV_PIX_FMT_RGB24, AV_PIX_FMT_GRAY8, ctx_scene, ctx_depth);
    const uint64_t index = frame_o->index();
    fq->push(std::move(frame_o));

    process_frame_thread(ctx_scene, fq, fm, etctx, shutdown);   // the reference's text
    if (fm->size() != 1) { std::fprintf(stderr, "process_frame_thread did not insert the frame\n"); return 4; }
    // keep the planes for the test before the encoder stage consumes the frame
    {
      auto f = fm->get_delete(index);
      dump(f->converted_frame_scene(), std::string(argv[6]) + ".scene.yuv");
      dump(f->converted_frame_depth(), std::string(argv[6]) + ".depth.yuv");
      fm->insert(0, std::move(f));   // send_frame_thread starts at frame_index 0 (encode.cpp:127)
    }
    shutdown = false;
    send_frame_thread(ctx_scene, ctx_depth, fm, shutdown);      // the reference's text
    if (ctx_scene->sent != 1 || ctx_depth->sent != 1) { std::fprintf(stderr, "send_frame_thread did not send both frames\n"); return 5; }
    std::printf("ok index=%llu sent=%d+%d refcount=%d\n", (unsigned long long)index, ctx_scene->sent, ctx_depth->sent, ctx_scene->last_ref_count);
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
