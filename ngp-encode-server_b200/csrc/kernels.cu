// kernels.cu -- sm_100a kernels of the per-frame pixel pipeline.
//
// What they compute (bit-exact integer restatement of libswscale's C path as the
// reference drives it, /root/reference/src/base/video/type_managers.cc:143-155,
// called from RenderedFrame::convert_frame, include/base/video/rendered_frame.h:24-33,
// after the overlay of src/base/video/render_text.cc:81-110):
//
//   k_frame_tiles   same-size path, ONE launch per batch of frames:
//                   [depth-select composite of N sources] -> glyph stamp overlay in
//                   shared memory -> Y (pointwise) + pair-summed chroma -> 8-tap
//                   vertical bicubic -> U,V; and depth GRAY8 -> Y (range
//                   compression) with U=V=128.
//   k_composite     composite to a scratch frame (only used ahead of a resize).
//   k_resize_tiles  general bicubic resize path (horizontal + vertical polyphase
//                   with libswscale's initFilter tables).
//
// All of this is HBM-bound u8/int32 work: no tensor cores.  The design rules are
// coalesced 128-bit global accesses, cp.async staging of the packed-pixel tile in
// shared memory (where the overlay is applied and the halo rows are shared by the
// whole CTA), conflict-free shared-memory access patterns, and a grid that is
// many waves of small CTAs (4 CTAs/SM).
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_common.cuh"
#include "nes_internal.h"

namespace nes {

// ---------------------------------------------------------------------------
// k_composite: composite N sources to a scratch frame (same pixel format + GRAY8).
// Only used when a composite is followed by a resize.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_composite(const DevJob *__restrict__ jobs, int n_jobs) {
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs[j];
    if (jb.n_src <= 1 || !jb.scratch_rgb) continue;  // only general-path composites get a scratch frame
    const int W = jb.W, H = jb.H;
    const size_t n = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      const int y = (int)(i / W), x = (int)(i % W);
      uint32_t d;
      if (jb.bpp == 4) composite_px<4>(jb, x, y, jb.scratch_rgb + i * 4, &d);
      else composite_px<3>(jb, x, y, jb.scratch_rgb + i * 3, &d);
      if (jb.scratch_depth) jb.scratch_depth[i] = (uint8_t)d;
    }
  }
}

// ---------------------------------------------------------------------------
// k_resize_tiles: general bicubic resize.
// One CTA produces an RS_TILE_W x RS_TILE_H block of destination luma, the matching
// chroma block and the matching depth-luma block.
//   stage A  source pixels (+overlay) -> y14 / u14,v14 (int16) in shared memory
//   stage H  horizontal polyphase -> 15-bit rows in shared memory
//   stage V  vertical polyphase -> 8-bit planes
// Source windows are bounded by the filter tables; shared memory is sized by the
// host from the worst-case window (launch_resize_tiles).
// ---------------------------------------------------------------------------
struct ResizeWin {
  int sx0, sx1, sy0, sy1;  // source window (pixels / rows), half-open
};

__device__ __forceinline__ int filt_lo(const DevFilter &f, int i) { return f.pos[i]; }
__device__ __forceinline__ int filt_hi(const DevFilter &f, int i) { return f.pos[i] + f.size; }

__global__ void __launch_bounds__(256) k_resize_tiles(const DevJob *__restrict__ jobs, int n_jobs, int smem_cap) {
  extern __shared__ __align__(16) uint8_t smem[];
  int tile;
  const DevJob *jp = find_job(jobs, n_jobs, blockIdx.x, &tile);
  if (!jp->general) return;
  const DevJob &jb = *jp;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tx = tile % jb.tiles_x, ty = tile / jb.tiles_x;
  const int W = jb.W, H = jb.H, Wd = jb.Wd, Hd = jb.Hd;
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  const int bpp = jb.bpp;
  const uint8_t *src = (jb.n_src > 1) ? jb.scratch_rgb : jb.src[0].rgb;
  const int sstride = (jb.n_src > 1) ? W * bpp : jb.src[0].rgb_stride;
  const uint8_t *dsrc = (jb.n_src > 1) ? jb.scratch_depth : jb.src[0].depth;
  const int dstride = (jb.n_src > 1) ? W : jb.src[0].depth_stride;

  // destination ranges of this tile
  const int dx0 = tx * RS_TILE_W, dx1 = min(dx0 + RS_TILE_W, Wd);
  const int dy0 = ty * RS_TILE_H, dy1 = min(dy0 + RS_TILE_H, Hd);
  const int cx0 = dx0 >> 1, cx1 = min((dx1 + 1) >> 1, cdW);
  const int cy0 = dy0 >> 1, cy1 = min((dy1 + 1) >> 1, cdH);

  // source windows: luma rows/cols, chroma rows / chroma-source cols
  // filter positions are monotone except at fixed-up borders: scan for the window
  int lr0 = H, lr1 = 0;
  for (int i = dy0; i < dy1; i++) { lr0 = min(lr0, jb.vl.pos[i]); lr1 = max(lr1, jb.vl.pos[i] + jb.vl.size); }
  int lc0 = W, lc1 = 0;
  for (int i = dx0; i < dx1; i++) { lc0 = min(lc0, jb.hl.pos[i]); lc1 = max(lc1, jb.hl.pos[i] + jb.hl.size); }
  int cr0 = H, cr1 = 0;
  for (int i = cy0; i < cy1; i++) { cr0 = min(cr0, jb.vc.pos[i]); cr1 = max(cr1, jb.vc.pos[i] + jb.vc.size); }
  int cc0 = jb.csW, cc1 = 0;
  for (int i = cx0; i < cx1; i++) { cc0 = min(cc0, jb.hc.pos[i]); cc1 = max(cc1, jb.hc.pos[i] + jb.hc.size); }
  // chroma-source columns -> pixel columns
  const int pc0 = jb.half ? cc0 * 2 : cc0, pc1 = jb.half ? cc1 * 2 : cc1;
  // union window of source pixels
  const int wx0 = min(lc0, pc0), wx1 = max(lc1, pc1);
  const int wy0 = min(lr0, cr0), wy1 = max(lr1, cr1);
  const int ww = wx1 - wx0, wh = wy1 - wy0;
  const int lw = lc1 - lc0, cw = cc1 - cc0;  // widths of the 14-bit rows fed to the H pass
  const int dw = dx1 - dx0, dcw = cx1 - cx0;

  // shared memory carve-up
  const int px_bytes = (wh * ww * bpp + 15) & ~15;
  uint8_t *s_px = smem;
  int16_t *s_y14 = (int16_t *)(smem + px_bytes);                 // [lr1-lr0][lw]
  const int y14_n = ((lr1 - lr0) * lw + 7) & ~7;
  int16_t *s_u14 = s_y14 + y14_n;                                // [cr1-cr0][cw]
  const int c14_n = ((cr1 - cr0) * cw + 7) & ~7;
  int16_t *s_v14 = s_u14 + c14_n;
  int16_t *s_hy = s_v14 + c14_n;                                 // [lr1-lr0][dw]
  const int hy_n = ((lr1 - lr0) * dw + 7) & ~7;
  int16_t *s_hu = s_hy + hy_n;                                   // [cr1-cr0][dcw]
  const int hc_n = ((cr1 - cr0) * dcw + 7) & ~7;
  int16_t *s_hv = s_hu + hc_n;
  int *s_hits = (int *)(s_hv + hc_n);
  int *s_nhits = s_hits + HIT_CAP;
  const int need = (int)((uint8_t *)(s_nhits + 4) - smem);
  if (need > smem_cap) { __trap(); }

  // ---- load source window ------------------------------------------------------
  {
    const int rowb = ww * bpp;
    for (int r = tid / 32; r < wh; r += nthr / 32) {
      const uint8_t *g = src + (size_t)(wy0 + r) * sstride + (size_t)wx0 * bpp;
      uint8_t *s = s_px + r * rowb;
      for (int i = tid & 31; i < rowb; i += 32) s[i] = g[i];
    }
  }
  __syncthreads();
  if (jb.n_glyphs > 0) {
    if (bpp == 3) stamp_glyphs<3>(jb, s_px, ww * 3, wx0, wy0, wx0, wx1, wy0, wy1, s_hits, s_nhits);
    else stamp_glyphs<4>(jb, s_px, ww * 4, wx0, wy0, wx0, wx1, wy0, wy1, s_hits, s_nhits);
  }

  // ---- stage A: 14-bit planes ---------------------------------------------------
  const int rb = jb.rgb_base;
  for (int i = tid; i < (lr1 - lr0) * lw; i += nthr) {
    const int r = i / lw, x = i % lw;
    const uint8_t *p = s_px + ((lr0 + r - wy0) * ww + (lc0 + x - wx0)) * bpp + (bpp == 4 ? rb : 0);
    s_y14[i] = (int16_t)((jb.cy[0] * p[0] + jb.cy[1] * p[1] + jb.cy[2] * p[2] + (32 << 14) + (1 << 8)) >> 9);
  }
  for (int i = tid; i < (cr1 - cr0) * cw; i += nthr) {
    const int r = i / cw, x = i % cw;
    if (jb.half) {
      const uint8_t *p = s_px + ((cr0 + r - wy0) * ww + (2 * (cc0 + x) - wx0)) * bpp + (bpp == 4 ? rb : 0);
      const int s0 = p[0] + p[bpp], s1 = p[1] + p[bpp + 1], s2 = p[2] + p[bpp + 2];
      s_u14[i] = (int16_t)((jb.cu[0] * s0 + jb.cu[1] * s1 + jb.cu[2] * s2 + C_BIAS) >> 10);
      s_v14[i] = (int16_t)((jb.cv[0] * s0 + jb.cv[1] * s1 + jb.cv[2] * s2 + C_BIAS) >> 10);
    } else {
      const uint8_t *p = s_px + ((cr0 + r - wy0) * ww + (cc0 + x - wx0)) * bpp + (bpp == 4 ? rb : 0);
      s_u14[i] = (int16_t)((jb.cu[0] * p[0] + jb.cu[1] * p[1] + jb.cu[2] * p[2] + C1_BIAS) >> 9);
      s_v14[i] = (int16_t)((jb.cv[0] * p[0] + jb.cv[1] * p[1] + jb.cv[2] * p[2] + C1_BIAS) >> 9);
    }
  }
  __syncthreads();

  // ---- stage H: horizontal polyphase, hScale16To15 (>>13, clamp 32767) ----------
  for (int i = tid; i < (lr1 - lr0) * dw; i += nthr) {
    const int r = i / dw, x = i % dw;
    const int16_t *f = jb.hl.coef + (size_t)(dx0 + x) * jb.hl.size;
    const int16_t *s = s_y14 + r * lw + (jb.hl.pos[dx0 + x] - lc0);
    int v = 0;
    for (int j = 0; j < jb.hl.size; j++) v += (int)s[j] * f[j];
    s_hy[i] = (int16_t)min(v >> 13, 32767);
  }
  for (int i = tid; i < (cr1 - cr0) * dcw; i += nthr) {
    const int r = i / dcw, x = i % dcw;
    const int16_t *f = jb.hc.coef + (size_t)(cx0 + x) * jb.hc.size;
    const int off = r * cw + (jb.hc.pos[cx0 + x] - cc0);
    int u = 0, v = 0;
    for (int j = 0; j < jb.hc.size; j++) { u += (int)s_u14[off + j] * f[j]; v += (int)s_v14[off + j] * f[j]; }
    s_hu[i] = (int16_t)min(u >> 13, 32767);
    s_hv[i] = (int16_t)min(v >> 13, 32767);
  }
  __syncthreads();

  // ---- stage V: vertical polyphase to 8 bit (yuv2planeX / yuv2plane1) -----------
  for (int i = tid; i < (dy1 - dy0) * dw; i += nthr) {
    const int r = i / dw, x = i % dw;
    const int yy = dy0 + r;
    int out;
    if (jb.vl.size == 1) {
      out = clip8((s_hy[(jb.vl.pos[yy] - lr0) * dw + x] + 64) >> 7);
    } else {
      const int16_t *f = jb.vl.coef + (size_t)yy * jb.vl.size;
      int v = 64 << 12;
      const int base = (jb.vl.pos[yy] - lr0) * dw + x;
      for (int j = 0; j < jb.vl.size; j++) v += (int)s_hy[base + j * dw] * f[j];
      out = clip8(v >> 19);
    }
    jb.sy[(size_t)yy * jb.sys + dx0 + x] = (uint8_t)out;
  }
  for (int i = tid; i < (cy1 - cy0) * dcw; i += nthr) {
    const int r = i / dcw, x = i % dcw;
    const int yy = cy0 + r;
    int ou, ov;
    if (jb.vc.size == 1) {
      const int o = (jb.vc.pos[yy] - cr0) * dcw + x;
      ou = clip8((s_hu[o] + 64) >> 7);
      ov = clip8((s_hv[o] + 64) >> 7);
    } else {
      const int16_t *f = jb.vc.coef + (size_t)yy * jb.vc.size;
      int u = 64 << 12, v = 64 << 12;
      const int base = (jb.vc.pos[yy] - cr0) * dcw + x;
      for (int j = 0; j < jb.vc.size; j++) { u += (int)s_hu[base + j * dcw] * f[j]; v += (int)s_hv[base + j * dcw] * f[j]; }
      ou = clip8(u >> 19);
      ov = clip8(v >> 19);
    }
    jb.su[(size_t)yy * jb.sus + cx0 + x] = (uint8_t)ou;
    jb.sv[(size_t)yy * jb.svs + cx0 + x] = (uint8_t)ov;
  }

  // ---- depth: hScale8To15 (>>7) -> range compression -> vertical; U=V=128 --------
  if (jb.dy) {
    __syncthreads();  // s_hy is reused
    for (int i = tid; i < (lr1 - lr0) * dw; i += nthr) {
      const int r = i / dw, x = i % dw;
      const int16_t *f = jb.hl.coef + (size_t)(dx0 + x) * jb.hl.size;
      const uint8_t *s = dsrc + (size_t)(lr0 + r) * dstride + jb.hl.pos[dx0 + x];
      int v = 0;
      for (int j = 0; j < jb.hl.size; j++) v += (int)s[j] * f[j];
      v = min(v >> 7, 32767);
      s_hy[i] = (int16_t)((v * 14071 + 33561472) >> 14);
    }
    __syncthreads();
    for (int i = tid; i < (dy1 - dy0) * dw; i += nthr) {
      const int r = i / dw, x = i % dw;
      const int yy = dy0 + r;
      int out;
      if (jb.vl.size == 1) {
        out = clip8((s_hy[(jb.vl.pos[yy] - lr0) * dw + x] + 64) >> 7);
      } else {
        const int16_t *f = jb.vl.coef + (size_t)yy * jb.vl.size;
        int v = 64 << 12;
        const int base = (jb.vl.pos[yy] - lr0) * dw + x;
        for (int j = 0; j < jb.vl.size; j++) v += (int)s_hy[base + j * dw] * f[j];
        out = clip8(v >> 19);
      }
      jb.dy[(size_t)yy * jb.dys + dx0 + x] = (uint8_t)out;
    }
    for (int i = tid; i < (cy1 - cy0) * dcw; i += nthr) {
      const int r = i / dcw, x = i % dcw;
      jb.du[(size_t)(cy0 + r) * jb.dus + cx0 + x] = 128;
      jb.dv[(size_t)(cy0 + r) * jb.dvs + cx0 + x] = 128;
    }
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int g_resize_smem_cap = 0;

int kernels_init() {
  cudaError_t e;
  if (int r = frame_strips_init()) return r;
  g_resize_smem_cap = 200 * 1024;
  e = cudaFuncSetAttribute(k_resize_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, g_resize_smem_cap);
  if (e != cudaSuccess) return (int)e;
  return 0;
}

int launch_composite(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream) {
  bool any = false;
  for (int j = 0; j < n_jobs; j++)
    if (jobs_host[j].n_src > 1 && jobs_host[j].scratch_rgb) any = true;
  if (!any) return 0;
  k_composite<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(jobs_dev, n_jobs);
  return 1;
}

int launch_resize_tiles(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream) {
  int total = 0, smem = 0;
  bool any = false;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs_host[j];
    total = jb.tile_base + jb.tiles_x * jb.tiles_y;
    if (jb.general) { any = true; smem = jb.rs_smem > smem ? jb.rs_smem : smem; }
  }
  if (!any || total == 0) return 0;
  smem = (smem + 1023) & ~1023;
  if (smem > g_resize_smem_cap) return -1;
  k_resize_tiles<<<total, 256, smem, (cudaStream_t)stream>>>(jobs_dev, n_jobs, smem);
  return 1;
}

}  // namespace nes
