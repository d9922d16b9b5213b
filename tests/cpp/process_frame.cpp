// tests/cpp/process_frame.cpp -- the body of the reference's process_frame_thread
// (/root/reference/src/encode.cpp:55-98) compiled against include/nes_gpu_shim.hpp instead of the
// reference's own headers: same calls, same order, same text formatting.
//
//   process_frame <wire message file> <font.ttf> <freetype.so> <dst_w> <dst_h> <timestamp> <out prefix>
//
// Writes <out>.scene.yuv and <out>.depth.yuv (Y||U||V without stride padding).
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <memory>
#include <sstream>
#include <vector>

#include "nes_gpu_shim.hpp"

static void dump(types::FrameManager &f, const std::string &path) {
  std::ofstream o(path, std::ios::binary);
  const unsigned w = f.context().width, h = f.context().height;
  for (unsigned y = 0; y < h; y++) o.write((const char *)f.data().data[0] + (size_t)y * f.data().linesize[0], w);
  for (int p = 1; p < 3; p++)
    for (unsigned y = 0; y < (h + 1) / 2; y++) o.write((const char *)f.data().data[p] + (size_t)y * f.data().linesize[p], (w + 1) / 2);
}

int main(int argc, char **argv) {
  if (argc < 8) { std::fprintf(stderr, "usage: see file header\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<uint8_t> wire((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    auto etctx = std::make_shared<RenderTextContext>(argv[2], argv[3]);
    const unsigned dw = (unsigned)std::atoi(argv[4]), dh = (unsigned)std::atoi(argv[5]);
    // server.cpp:172-194: ParseFromString of the message, then the reference's RenderedFrame constructor
    // (RGB24 scene + GRAY8 depth, destination size from the codec managers)
    struct Codec {
      struct Info { unsigned width, height; AVPixelFormat pix_fmt; };
      Info info;
      Info *get_codec_info() { return &info; }
    };
    auto ctx_scene = std::make_shared<Codec>(Codec{{dw, dh, AV_PIX_FMT_YUV420P}}), ctx_depth = std::make_shared<Codec>(Codec{{dw, dh, AV_PIX_FMT_YUV420P}});
    nesproto::RenderedFrame msg;
    if (!msg.ParseFromArray(wire.data() + 8, (int)wire.size() - 8)) { std::fprintf(stderr, "ParseFromArray failed\n"); return 4; }
    std::unique_ptr<RenderedFrame> frame = std::make_unique<RenderedFrame>(msg, AV_PIX_FMT_RGB24, AV_PIX_FMT_GRAY8, ctx_scene, ctx_depth);

    // ---- encode.cpp:55-98 ----------------------------------------------------------------
    uint64_t frame_index = frame->index();
    std::stringstream cam_matrix;
    int idx = 0;
    for (auto it : frame->get_cam().matrix()) {
      idx++;
      cam_matrix << std::fixed << std::showpos << std::setw(7) << std::setprecision(5) << std::setfill('0') << it << ' ';
      if (idx % 4 == 0) cam_matrix << '\n';
    }
    cam_matrix << std::fixed << std::showpos << std::setw(7) << std::setprecision(5) << std::setfill('0') << 0.f << ' ' << 0.f << ' ' << 0.f << ' ' << 1.f << ' ';

    etctx->render_string_to_frame(frame->source_frame_scene(), RenderTextContext::RenderPosition::RENDER_POSITION_CENTER, cam_matrix.str());
    etctx->render_string_to_frame(frame->source_frame_scene(), RenderTextContext::RenderPosition::RENDER_POSITION_LEFT_BOTTOM,
                                  std::string("index=") + std::to_string(frame->index()));
    etctx->render_string_to_frame(frame->source_frame_scene(), RenderTextContext::RenderPosition::RENDER_POSITION_LEFT_TOP, argv[6]);
    std::string direction = frame->is_left() ? "direction=left" : "direction=right";
    etctx->render_string_to_frame(frame->source_frame_scene(), RenderTextContext::RenderPosition::RENDER_POSITION_RIGHT_TOP, direction);
    frame->convert_frame();
    // ---------------------------------------------------------------------------------------
    (void)frame_index;
    bool threw = false;
    try { frame->convert_frame(); } catch (const std::runtime_error &) { threw = true; }
    if (!threw) { std::fprintf(stderr, "second convert_frame() did not throw\n"); return 3; }

    dump(frame->converted_frame_scene(), std::string(argv[7]) + ".scene.yuv");
    dump(frame->converted_frame_depth(), std::string(argv[7]) + ".depth.yuv");

    // the zero-copy constructor (wire bytes, payload borrowed) gives the same planes
    {
      RenderedFrame z(wire.data(), wire.size(), true, AV_PIX_FMT_RGB24, AV_PIX_FMT_GRAY8, dw, dh);
      for (auto &r : frame->source_frame_scene().text_runs()) z.source_frame_scene().text_runs().push_back(r);
      z.convert_frame();
      dump(z.converted_frame_scene(), std::string(argv[7]) + ".zero.yuv");
    }
    // encoder hand-off (encode.cpp:136-137): to_avframe() is a ref-counted AVFrame over the converted planes
    {
      auto wrapped = frame->converted_frame_scene().to_avframe();
      if (nes_avframe_ref_count(wrapped.get()) != 1) { std::fprintf(stderr, "to_avframe: unexpected reference count\n"); return 5; }
    }

    // types::SwsContextManager on its own (type_managers.cc:143-155)
    types::FrameManager dst(types::FrameManager::FrameContext(dw, dh, AV_PIX_FMT_YUV420P));
    { types::SwsContextManager sws(frame->source_frame_scene(), dst); }
    dump(dst, std::string(argv[7]) + ".sws.yuv");
    std::printf("ok index=%llu left=%d\n", (unsigned long long)frame->index(), (int)frame->is_left());
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
