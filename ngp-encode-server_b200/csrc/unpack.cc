// unpack.cc -- zero-copy unpack of a length-prefixed nesproto.RenderedFrame.
//
// Replaces, for the pixel path, the chain
//   socket_receive_blocking_lpf   /root/reference/src/server.cpp:91-112 (8-byte native
//                                 size_t length, then the payload)
//   RenderedFrame::ParseFromString  src/server.cpp:175
//   RenderedFrame ctor              src/base/video/rendered_frame.cc:5-27
// which deep-copies the payload three times.  Here the two `bytes` fields are only
// located (offset + length inside the receive buffer) so that the H2D copy can read
// them in place.  Schema: proto/nes.proto:4-25
//   Camera        { bool is_left = 2; uint32 width = 3; uint32 height = 4; repeated float matrix = 12; }
//   RenderedFrame { uint64 index = 1; Camera camera = 2; bool is_left = 3; bytes frame = 6; bytes depth = 7; }
// proto3 wire rules honoured: fields in any order, zero scalars omitted, last value of
// a scalar wins, repeated float packed or unpacked, embedded messages merged, unknown
// fields skipped.
#include <cstdint>
#include <cstring>

#include "nes_gpu.h"

namespace {

struct Reader {
  const uint8_t *p;
  const uint8_t *end;
  bool ok = true;

  bool varint(uint64_t *v) {
    uint64_t r = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) return ok = false;
      const uint8_t b = *p++;
      r |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) { *v = r; return true; }
    }
    return ok = false;
  }
  bool skip(uint64_t n) {
    if ((uint64_t)(end - p) < n) return ok = false;
    p += n;
    return true;
  }
  bool skip_field(unsigned wire_type) {
    uint64_t t;
    switch (wire_type) {
      case 0: return varint(&t);
      case 1: return skip(8);
      case 2: return varint(&t) && skip(t);
      case 5: return skip(4);
      default: return ok = false;  // groups (3,4) are not used by nes.proto
    }
  }
};

bool parse_camera(Reader r, nes_unpacked_frame *out) {
  while (r.p < r.end) {
    uint64_t tag, v;
    if (!r.varint(&tag)) return false;
    const unsigned field = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
    if (wt == 0 && (field == 2 || field == 3 || field == 4)) {
      if (!r.varint(&v)) return false;
      if (field == 2) out->cam_is_left = v != 0;
      else if (field == 3) out->width = (int32_t)(uint32_t)v;
      else out->height = (int32_t)(uint32_t)v;
    } else if (field == 12 && wt == 2) {
      if (!r.varint(&v) || (uint64_t)(r.end - r.p) < v) return false;
      for (uint64_t i = 0; i + 4 <= v; i += 4) {
        if (out->n_matrix < 16) std::memcpy(&out->matrix[out->n_matrix], r.p + i, 4);
        out->n_matrix++;
      }
      r.p += v;
    } else if (field == 12 && wt == 5) {
      if ((r.end - r.p) < 4) return false;
      if (out->n_matrix < 16) std::memcpy(&out->matrix[out->n_matrix], r.p, 4);
      out->n_matrix++;
      r.p += 4;
    } else if (!r.skip_field(wt)) {
      return false;
    }
  }
  return true;
}

}  // namespace

extern "C" int nes_unpack_rendered_frame(const uint8_t *buf, uint64_t len, int has_length_prefix,
                                         nes_unpacked_frame *out) {
  if (!buf || !out) return NES_ERR_INVALID_ARG;
  std::memset(out, 0, sizeof(*out));
  uint64_t start = 0, end = len;
  if (has_length_prefix) {
    if (len < 8) return NES_ERR_PARSE;
    uint64_t n;
    std::memcpy(&n, buf, 8);  // native-endian size_t, server.cpp:46-60,91-100
    if (n > len - 8) return NES_ERR_PARSE;
    start = 8;
    end = 8 + n;
  }
  out->consumed = end;
  Reader r{buf + start, buf + end};
  while (r.p < r.end) {
    uint64_t tag, v;
    if (!r.varint(&tag)) return NES_ERR_PARSE;
    const unsigned field = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
    if (field == 1 && wt == 0) {
      if (!r.varint(&v)) return NES_ERR_PARSE;
      out->index = v;
    } else if (field == 3 && wt == 0) {
      if (!r.varint(&v)) return NES_ERR_PARSE;
      out->is_left = v != 0;
    } else if ((field == 2 || field == 6 || field == 7) && wt == 2) {
      if (!r.varint(&v) || (uint64_t)(r.end - r.p) < v) return NES_ERR_PARSE;
      if (field == 2) {
        if (!parse_camera(Reader{r.p, r.p + v}, out)) return NES_ERR_PARSE;
      } else if (field == 6) {
        out->frame_off = (uint64_t)(r.p - buf);
        out->frame_len = v;
      } else {
        out->depth_off = (uint64_t)(r.p - buf);
        out->depth_len = v;
      }
      r.p += v;
    } else if (!r.skip_field(wt)) {
      return NES_ERR_PARSE;
    }
  }
  if (out->n_matrix > 16) out->n_matrix = 16;
  return NES_OK;
}
