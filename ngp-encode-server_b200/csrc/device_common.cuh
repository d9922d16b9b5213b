// device_common.cuh -- constants and device helpers shared by the kernels.
#ifndef NES_DEVICE_COMMON_CUH_
#define NES_DEVICE_COMMON_CUH_
#include <cuda_runtime.h>
#include <stdint.h>

#include "nes_internal.h"

namespace nes {

// BT.601 limited range, 15-bit
// Y = (RY*R + GY*G + BY*B + (32<<14) + (1<<8)) >> 9  -> 14 bit; *2 -> 15 bit;
// 8-bit out = (y15 + 64) >> 7.  Folded: Y = (S + (32<<14) + (1<<8) + (64<<8)) >> 15
// (nested floors; the low bit cleared by "*2" cannot carry because 64 is even;
// no clip is needed: 16 <= Y <= 251).
constexpr int Y_BIAS = (32 << 14) + (1 << 8) + (64 << 8);
// chroma of a horizontal pixel pair: u14 = (RU*r2 + GU*g2 + BU*b2 + (256<<15) + (1<<9)) >> 10,
// u15 = 2*u14 = (S >> 9) & ~1 (S > 0 always; 2*u14 <= 30720 so min(.,32767) never fires).
constexpr int C_BIAS = (256 << 15) + (1 << 9);
// per-pixel chroma (resize path without pair sum): (..., + (256<<14) + (1<<8)) >> 9
constexpr int C1_BIAS = (256 << 14) + (1 << 8);
// GRAY8 -> limited range luma: ((((d<<7)*14071 + 33561472) >> 14) + 64) >> 7 == (d*219 + 127)/255 + 16;
// frame_strips.cu computes it with one dp2a per pixel, resize_tiles.cu with the two-step form.

__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ const DevJob *find_job(const DevJob *jobs, int n_jobs, int bid, int *tile) {
  int j = 0;
  // jobs are few (<= a few hundred); tile_base is a prefix sum
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_base <= bid) lo = mid; else hi = mid - 1;
  }
  j = lo;
  *tile = bid - jobs[j].tile_base;
  return jobs + j;
}

// ---------------------------------------------------------------------------
// depth-select composite of one pixel (self-defined semantics, DESIGN.md §composite;
// oracle/overlay_port.c nes_oracle_composite): among valid sources (bpp 4: alpha != 0;
// bpp 3: always) the smallest depth wins, ties to the lowest source index; no valid
// source -> pixel bytes 0, depth 255.
// ---------------------------------------------------------------------------
template <int BPP>
__device__ __forceinline__ void composite_px(const DevJob &jb, int x, int y, uint8_t *out_px, uint32_t *out_d) {
  int best = -1;
  uint32_t best_d = 256;
  for (int k = 0; k < jb.n_src; k++) {
    const uint8_t *p = jb.src[k].rgb + (size_t)y * jb.src[k].rgb_stride + (size_t)x * BPP;
    const uint32_t d = jb.src[k].depth ? jb.src[k].depth[(size_t)y * jb.src[k].depth_stride + x] : 0u;
    const bool valid = (BPP == 3) ? true : (p[jb.a_off] != 0);
    if (valid && d < best_d) { best = k; best_d = d; }
  }
  if (best < 0) {
#pragma unroll
    for (int c = 0; c < BPP; c++) out_px[c] = 0;
    *out_d = 255;
  } else {
    const uint8_t *p = jb.src[best].rgb + (size_t)y * jb.src[best].rgb_stride + (size_t)x * BPP;
#pragma unroll
    for (int c = 0; c < BPP; c++) out_px[c] = p[c];
    *out_d = best_d;
  }
}

}  // namespace nes
#endif
