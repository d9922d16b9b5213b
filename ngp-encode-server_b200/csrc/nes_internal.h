// nes_internal.h -- device-visible job descriptors shared by kernels.cu and session.cu.
// Not part of the public ABI (include/nes_gpu.h is).
#ifndef NES_INTERNAL_H_
#define NES_INTERNAL_H_

#include <stdint.h>

#include "nes_gpu.h"

namespace nes {

// Same-size fused tile geometry (kernels.cu, k_frame_tiles): one CTA owns a
// TILE_W x TILE_H block of source pixels and loads 3 halo rows above and below
// for the 8-tap vertical chroma filter.
constexpr int TILE_W = 256;
constexpr int TILE_H = 30;  // divides 720/1080/1440/2160: no ragged last tile row
constexpr int HALO = 3;
constexpr int TILE_ROWS = TILE_H + 2 * HALO;
constexpr int CTA_THREADS = 256;
constexpr int HIT_CAP = 256;  // glyph rect tests per overlay chunk
constexpr int MASK_WORDS = 128;  // per-job bitmap of tiles touched by text (4096 tiles)

// Resize tile geometry (k_resize_tiles): destination pixels per CTA.
constexpr int RS_TILE_W = 64;
constexpr int RS_TILE_H = 16;
// Below this source height libswscale's vertical chroma filter has fewer than 8 taps
// (initFilter clamps the size to srcH-2), so the fused same-size kernel does not apply.
constexpr int MIN_FUSED_H = 12;

struct DevSource {
  const uint8_t *rgb;
  const uint8_t *depth;
  int32_t rgb_stride;
  int32_t depth_stride;
};

// One glyph of a text run placed in the frame (top-left of its bitmap).
struct DevPlaced {
  int32_t x, y;
  int32_t w, h;
  int32_t pitch;
  uint32_t atlas_off;
};

struct DevFilter {
  const int16_t *coef;  // [dst][size]
  const int32_t *pos;   // [dst]
  int32_t size;
  int32_t pad;
};

struct DevJob {
  DevSource src[NES_MAX_SOURCES];
  int32_t n_src;
  int32_t bpp;       // 3 or 4
  int32_t rgb_base;  // byte offset of the first colour byte inside a pixel (0, or 1 for ARGB/ABGR)
  int32_t a_off;     // byte offset of alpha (bpp 4) or -1
  int32_t cy[3], cu[3], cv[3];  // BT.601 coefficients per colour byte position
  // the same coefficients packed for dp2a on a pixel word (bytes 0,1 -> [0]; bytes 2,3 -> [1]),
  // 16-bit signed halves; bytes that are not colour get 0
  uint32_t ky[2], ku[2], kv[2];
  int32_t W, H, Wd, Hd;
  uint8_t *sy, *su, *sv;  // scene planes
  int32_t sys, sus, svs;
  int32_t out_vec;        // 1: all destination pointers/strides 16-byte aligned
  uint8_t *dy, *du, *dv;  // depth planes (dy == nullptr: no depth stream)
  int32_t dys, dus, dvs;
  int32_t in_vec;         // 1: all source pointers/strides 16-byte aligned
  const DevPlaced *glyphs;
  const uint8_t *atlas;
  int32_t n_glyphs;
  int32_t tma_ok;     // single source, 16-byte aligned rows: tiles are staged with bulk async copies
  int32_t use_mask;   // tile_mask valid (tiles_x*tiles_y <= 32*MASK_WORDS)
  int32_t pad1;
  uint32_t tile_mask[MASK_WORDS];  // bit t set: tile t (or its halo rows) intersects a placed glyph
  int32_t tiles_x, tiles_y, tile_base;
  // resize only
  DevFilter hl, hc, vl, vc;
  int32_t half;  // chroma horizontally pair-summed before the H pass
  int32_t csW;   // chroma source width fed to the H pass
  int32_t rs_smem;  // shared memory the resize kernel needs for this job's worst tile (host use)
  int32_t general;  // 1: k_resize_tiles (any size change, or H < 12 where libswscale's chroma filter is truncated)
  // composite scratch (resize of a composite goes through a scratch frame)
  uint8_t *scratch_rgb;
  uint8_t *scratch_depth;
};

// Launchers (kernels.cu).  jobs: device pointer to n_jobs descriptors.
int launch_frame_tiles(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream);
int launch_resize_tiles(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream);
int launch_composite(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream);
int kernels_init();  // opt-in shared memory sizes; returns cudaError_t
int frame_tiles_init();

}  // namespace nes
#endif
