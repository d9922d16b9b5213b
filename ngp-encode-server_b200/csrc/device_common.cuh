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
// GRAY8 -> limited range luma: ((((d<<7)*14071 + 33561472) >> 14) + 64) >> 7
//   == (d*1801088 + 34610048) >> 21  (nested floors)
constexpr int G_MUL = 14071 << 7;
constexpr int G_ADD = 33561472 + (64 << 14);

__device__ __forceinline__ uint32_t byte_at(uint32_t w, int p) { return __byte_perm(w, 0u, 0x4440u | (uint32_t)p); }
__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ uint32_t gray_y(uint32_t d) { return (d * (uint32_t)G_MUL + (uint32_t)G_ADD) >> 21; }
__device__ __forceinline__ uint32_t gray_y4(uint32_t w) {
  return gray_y(w & 255u) | (gray_y((w >> 8) & 255u) << 8) | (gray_y((w >> 16) & 255u) << 16) | (gray_y(w >> 24) << 24);
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

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

// 4 pixels of RGBA-family sources at once (aligned fast path).
__device__ __forceinline__ void composite_px4(const DevJob &jb, int x, int y, uint4 *out_px, uint32_t *out_d4) {
  uint4 best_px = make_uint4(0, 0, 0, 0);
  uint32_t bd0 = 256, bd1 = 256, bd2 = 256, bd3 = 256;
  const int ash = jb.a_off * 8;
  for (int k = 0; k < jb.n_src; k++) {
    const uint4 p = __ldg((const uint4 *)(jb.src[k].rgb + (size_t)y * jb.src[k].rgb_stride + (size_t)x * 4));
    const uint32_t dw = __ldg((const uint32_t *)(jb.src[k].depth + (size_t)y * jb.src[k].depth_stride + x));
    const uint32_t d0 = dw & 255u, d1 = (dw >> 8) & 255u, d2 = (dw >> 16) & 255u, d3 = dw >> 24;
    if (((p.x >> ash) & 255u) && d0 < bd0) { bd0 = d0; best_px.x = p.x; }
    if (((p.y >> ash) & 255u) && d1 < bd1) { bd1 = d1; best_px.y = p.y; }
    if (((p.z >> ash) & 255u) && d2 < bd2) { bd2 = d2; best_px.z = p.z; }
    if (((p.w >> ash) & 255u) && d3 < bd3) { bd3 = d3; best_px.w = p.w; }
  }
  *out_px = best_px;
  *out_d4 = min(bd0, 255u) | (min(bd1, 255u) << 8) | (min(bd2, 255u) << 16) | (min(bd3, 255u) << 24);
}

// ---------------------------------------------------------------------------
// glyph stamp into a shared-memory pixel tile.  Reference semantics
// (render_text.cc:94-106): every bitmap pixel with coverage != 0 that falls inside the
// frame becomes (255,255,255).  All stamps write the same value, so overlapping glyphs
// and concurrent warps are order-free.
//   tile origin (ox, oy) in frame coordinates, tile extent cols [cx0,cx1) rows [ry0,ry1)
// ---------------------------------------------------------------------------
template <int BPP>
__device__ __forceinline__ void stamp_glyphs(const DevJob &jb, uint8_t *s_px, int row_bytes, int ox, int oy, int cx0,
                                             int cx1, int ry0, int ry1, int *s_hits, int *s_nhits) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  for (int base = 0; base < jb.n_glyphs; base += HIT_CAP) {
    if (tid == 0) *s_nhits = 0;
    __syncthreads();
    for (int g = base + tid; g < min(base + HIT_CAP, jb.n_glyphs); g += blockDim.x) {
      const DevPlaced pg = jb.glyphs[g];
      if (pg.x < cx1 && pg.x + pg.w > cx0 && pg.y < ry1 && pg.y + pg.h > ry0) s_hits[atomicAdd(s_nhits, 1)] = g;
    }
    __syncthreads();
    const int nh = *s_nhits;
    for (int h = warp; h < nh; h += nwarps) {
      const DevPlaced pg = jb.glyphs[s_hits[h]];
      const uint8_t *cov = jb.atlas + pg.atlas_off;
      const int q0 = max(0, ry0 - pg.y), q1 = min(pg.h, ry1 - pg.y);
      const int p0 = max(0, cx0 - pg.x), p1 = min(pg.w, cx1 - pg.x);
      for (int q = q0; q < q1; q++) {
        for (int p = p0 + lane; p < p1; p += 32) {
          if (cov[q * pg.pitch + p]) {
            uint8_t *px = s_px + (pg.y + q - oy) * row_bytes + (pg.x + p - ox) * BPP + (BPP == 4 ? jb.rgb_base : 0);
            px[0] = 255; px[1] = 255; px[2] = 255;
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace nes
#endif
