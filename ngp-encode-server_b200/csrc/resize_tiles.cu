// resize_tiles.cu -- k_resize_tiles: the general bicubic path (any size change, and frames
// under 12 rows where libswscale truncates the vertical chroma filter).
//
// What it computes: libswscale's generic C path as the reference drives it
// (/root/reference/src/base/video/type_managers.cc:143-155 via rendered_frame.h:24-33, after the
// overlay of render_text.cc:81-110); integer spec in SURVEY.md Appendix A.3 / A.4:
//   scene : RGB(A) -> 14-bit Y / (pair-summed) U,V -> horizontal polyphase (>>13, 15 bit)
//           -> vertical polyphase (>>19) -> 8 bit planes
//   depth : GRAY8 -> horizontal polyphase (>>7) -> range compression -> vertical; U = V = 128
// with the filter tables of csrc/filter.cc (bit-identical to initFilter's).
//
// Shape (HBM-bound in principle: every source byte is read once from DRAM, halos hit L2):
//   * one CTA per destination tile (rs_tw x rs_th luma samples, chosen per size pair so that
//     two CTAs fit an SM); the source window of a tile comes from per-tile-column / per-tile-row
//     tables built on the host with the filter (no scanning on the device);
//   * stage A reads the window straight from global memory, 4 pixels per lane (16-byte loads of
//     every source + 4 depth bytes), does the depth-select composite in registers and the glyph
//     overlay from a per-tile bit mask, and writes 14-bit planes (int16) + depth bytes to shared
//     memory: no packed-pixel tile, no scratch frame, the composite costs no extra pass;
//   * stage H: lane = destination column (its taps live in registers), warp = source row;
//   * stage V: lane = 4 adjacent destination columns (int32 rows, 16-byte shared loads), 4-byte stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_common.cuh"
#include "nes_internal.h"

namespace nes {

namespace {

__device__ __forceinline__ int dp2a_lo_s(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_s(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// coefficient . colour bytes of one pixel word (k0: bytes 0,1; k1: bytes 2,3; non-colour bytes carry 0)
__device__ __forceinline__ int dot_px(uint32_t k0, uint32_t k1, uint32_t px, int acc) { return dp2a_hi_s(k1, px, dp2a_lo_s(k0, px, acc)); }

__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
// 4 pixels at columns x..x+3 of row y as pixel words (3-byte pixels: r | g<<8 | b<<16 | junk<<24,
// the junk byte has coefficient 0) + their 4 depth bytes, after the depth-select composite.
template <int BPP>
__device__ __forceinline__ void load_group(const DevJob &jb, int x, int y, bool want_depth, uint32_t (&px)[4], uint32_t *d4) {
  const int W = jb.W;
  if (jb.in_vec && x + 4 <= W && (BPP == 4 || jb.n_src == 1)) {
    if (jb.n_src == 1) {
      const uint8_t *p = jb.src[0].rgb + (size_t)y * jb.src[0].rgb_stride + (size_t)x * BPP;
      if (BPP == 4) {
        const uint4 q = __ldg((const uint4 *)p);
        px[0] = q.x; px[1] = q.y; px[2] = q.z; px[3] = q.w;
      } else {
        const uint32_t w0 = __ldg((const uint32_t *)p), w1 = __ldg((const uint32_t *)p + 1), w2 = __ldg((const uint32_t *)p + 2);
        px[0] = w0;
        px[1] = __funnelshift_r(w0, w1, 24);
        px[2] = __funnelshift_r(w1, w2, 16);
        px[3] = w2 >> 8;
      }
      *d4 = want_depth ? __ldg((const uint32_t *)(jb.src[0].depth + (size_t)y * jb.src[0].depth_stride + x)) : 0u;
    } else {
      // depth-select composite (device_common.cuh composite_px for the semantics): all loads of the
      // first four sources are issued before the first compare
      uint4 q[4];
      uint32_t dw[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < jb.n_src) {
          q[k] = __ldg((const uint4 *)(jb.src[k].rgb + (size_t)y * jb.src[k].rgb_stride + (size_t)x * 4));
          dw[k] = __ldg((const uint32_t *)(jb.src[k].depth + (size_t)y * jb.src[k].depth_stride + x));
        }
      const uint32_t am = 0xFFu << (8 * jb.a_off);
      uint32_t bd[4] = {256, 256, 256, 256};
      px[0] = px[1] = px[2] = px[3] = 0;
      auto consider = [&](const uint4 &p, uint32_t d) {
        const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t di = (d >> (8 * i)) & 255u;
          const bool take = (w[i] & am) != 0 && di < bd[i];
          bd[i] = take ? di : bd[i];
          px[i] = take ? w[i] : px[i];
        }
      };
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < jb.n_src) consider(q[k], dw[k]);
      for (int k = 4; k < jb.n_src; k++)
        consider(__ldg((const uint4 *)(jb.src[k].rgb + (size_t)y * jb.src[k].rgb_stride + (size_t)x * 4)),
                 __ldg((const uint32_t *)(jb.src[k].depth + (size_t)y * jb.src[k].depth_stride + x)));
      *d4 = min(bd[0], 255u) | (min(bd[1], 255u) << 8) | (min(bd[2], 255u) << 16) | (min(bd[3], 255u) << 24);
    }
    return;
  }
  // unaligned sources, 3-byte composites, the ragged right edge: per pixel
  *d4 = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    px[i] = 0;
    if (x + i >= W) continue;
    uint8_t b[4] = {0, 0, 0, 0};
    uint32_t d = 0;
    if (jb.n_src == 1) {
      const uint8_t *p = jb.src[0].rgb + (size_t)y * jb.src[0].rgb_stride + (size_t)(x + i) * BPP;
#pragma unroll
      for (int c = 0; c < BPP; c++) b[c] = p[c];
      if (want_depth) d = jb.src[0].depth[(size_t)y * jb.src[0].depth_stride + x + i];
    } else {
      composite_px<BPP>(jb, x + i, y, b, &d);
    }
    px[i] = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
    *d4 |= d << (8 * i);
  }
}

// Per-source addressing staged in shared memory once per tile (the job descriptor lives in global
// memory; stage A would otherwise re-read it for every group of pixels).
struct SrcDesc {
  const uint8_t *rgb, *depth;
  int32_t rgb_stride, depth_stride;
};

// Aligned fast path of load_group for 4-byte pixels: the lane's column offset is fixed, sources come
// from shared memory, the first source is taken wherever it is valid without a compare.
__device__ __forceinline__ void load_group_fast4(const SrcDesc *sd, int n_src, uint32_t a_mask, int xb /* x * 4 */, int x, int y, uint32_t (&px)[4],
                                                 uint32_t *d4) {
  uint4 q[4];
  uint32_t dw[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (k < n_src) {
      q[k] = __ldg((const uint4 *)(sd[k].rgb + (uint32_t)(y * sd[k].rgb_stride + xb)));
      dw[k] = __ldg((const uint32_t *)(sd[k].depth + (uint32_t)(y * sd[k].depth_stride + x)));
    }
  if (n_src == 1) {  // a single source is converted as it is (alpha only matters to the composite)
    px[0] = q[0].x; px[1] = q[0].y; px[2] = q[0].z; px[3] = q[0].w;
    *d4 = dw[0];
    return;
  }
  uint32_t bd[4];
  {
    const uint32_t w[4] = {q[0].x, q[0].y, q[0].z, q[0].w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool valid = (w[i] & a_mask) != 0;
      bd[i] = valid ? __byte_perm(dw[0], 0u, 0x4440 + i) : 256u;
      px[i] = valid ? w[i] : 0u;
    }
  }
  auto consider = [&](const uint4 &p, uint32_t d) {
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t di = __byte_perm(d, 0u, 0x4440 + i);
      const bool take = (w[i] & a_mask) != 0 && di < bd[i];
      bd[i] = take ? di : bd[i];
      px[i] = take ? w[i] : px[i];
    }
  };
#pragma unroll
  for (int k = 1; k < 4; k++)
    if (k < n_src) consider(q[k], dw[k]);
  for (int k = 4; k < n_src; k++)
    consider(__ldg((const uint4 *)(sd[k].rgb + (uint32_t)(y * sd[k].rgb_stride + xb))), __ldg((const uint32_t *)(sd[k].depth + (uint32_t)(y * sd[k].depth_stride + x))));
  const uint32_t lo = __byte_perm(min(bd[0], 255u), min(bd[1], 255u), 0x0040), hi = __byte_perm(min(bd[2], 255u), min(bd[3], 255u), 0x0040);
  *d4 = __byte_perm(lo, hi, 0x5410);
}

struct Tile {
  int dx0, dw, dwp, dy0, dh;      // destination luma block (dwp: row stride of the H-pass output, multiple of 4)
  int cx0, dcw, dcwp, cy0, dch;   // destination chroma block
  int lc0, lc1, lr0, lr1;         // luma source columns / rows
  int cc0, cc1, cr0, cr1;         // chroma source columns (chroma-source units) / rows
  int wx0, ww, wy0, wh;           // union source window in pixels (wx0, ww multiples of 4)
  int cww;                        // chroma plane row stride (ww/2 when pair-summed, else ww)
};

__device__ __forceinline__ int lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return (int)v;
}
__device__ __forceinline__ int lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return (int)v;
}
__device__ __forceinline__ int lds_s16(uint32_t a) {
  int v;
  asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_32(uint32_t a, int v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

// two carve-ups are the same launch only if every region offset agrees
__host__ __device__ inline bool same_layout(const RsLayout &a, const RsLayout &b) {
  return a.y14 == b.y14 && a.u14 == b.u14 && a.v14 == b.v14 && a.hy == b.hy && a.hu == b.hu && a.hv == b.hv && a.dep == b.dep && a.mask == b.mask &&
         a.hits == b.hits && a.vtab == b.vtab && a.src == b.src && a.total == b.total;
}

enum { H_SCENE = 0, H_DEPTH = 1 };

// Horizontal polyphase of NP planes that share one filter (NP = 2: U and V):
//   out[row][dx] = fin( sum_j in[row][pos[dx] + j] * f[dx][j] )
// lane = destination column (its T taps and its position live in registers), warp = source row;
// the tap loads are one shared load each at an immediate offset.  ESZ: bytes per input sample
// (2: 14-bit planes, 1: depth bytes).  Output is int32 so that the vertical pass needs no unpacking.
//   H_SCENE: hScale16To15  min(v >> 13, 32767)
//   H_DEPTH: hScale8To15   min(v >> 7, 32767), then the GRAY8 range compression (Appendix A.4)
template <int T, int NP, int ESZ, int MODE>
__device__ __forceinline__ void hpass(const DevFilter &f, int d0, int n_dst, int n_rows, const uint32_t (&in)[NP], int in_stride, int x_org,
                                      const uint32_t (&out)[NP], int out_stride, int lane, int warp) {
  constexpr int NWARP = RS_THREADS / 32;
  for (int db = 0; db < n_dst; db += 32) {
    const int dx = db + lane;
    if (dx >= n_dst) continue;
    const int pos = f.pos[d0 + dx] - x_org;
    const int16_t *cf = f.coef + (size_t)(d0 + dx) * f.size;
    int c[T > 0 ? T : 1];
    if (T > 0) {
#pragma unroll
      for (int j = 0; j < T; j++) c[j] = j < f.size ? (int)cf[j] : 0;
    }
    uint32_t a[NP], o[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { a[p] = in[p] + warp * in_stride + pos * ESZ; o[p] = out[p] + warp * out_stride + dx * 4; }
    for (int row = warp; row < n_rows; row += NWARP) {
#pragma unroll
      for (int p = 0; p < NP; p++) {
        int v = 0;
        if (T > 0) {
#pragma unroll
          for (int j = 0; j < T; j++) v += (ESZ == 2 ? lds_u16(a[p] + 2 * j) : lds_u8(a[p] + j)) * c[j];
        } else {
          for (int j = 0; j < f.size; j++) v += (ESZ == 2 ? lds_u16(a[p] + 2 * j) : lds_u8(a[p] + j)) * (int)cf[j];
        }
        if (MODE == H_SCENE) v = min(v >> 13, 32767);
        else { v = min(v >> 7, 32767); v = (v * 14071 + 33561472) >> 14; }
        sts_32(o[p], v);
        a[p] += NWARP * in_stride; o[p] += NWARP * out_stride;
      }
    }
  }
}

template <int NP, int ESZ, int MODE>
__device__ __forceinline__ void hpass_any(const DevFilter &f, int d0, int n_dst, int n_rows, const uint32_t (&in)[NP], int in_stride, int x_org,
                                          const uint32_t (&out)[NP], int out_stride, int lane, int warp) {
  if (f.size <= 4) hpass<4, NP, ESZ, MODE>(f, d0, n_dst, n_rows, in, in_stride, x_org, out, out_stride, lane, warp);
  else if (f.size <= 6) hpass<6, NP, ESZ, MODE>(f, d0, n_dst, n_rows, in, in_stride, x_org, out, out_stride, lane, warp);
  else if (f.size <= 8) hpass<8, NP, ESZ, MODE>(f, d0, n_dst, n_rows, in, in_stride, x_org, out, out_stride, lane, warp);
  else if (f.size <= 12) hpass<12, NP, ESZ, MODE>(f, d0, n_dst, n_rows, in, in_stride, x_org, out, out_stride, lane, warp);
  else hpass<0, NP, ESZ, MODE>(f, d0, n_dst, n_rows, in, in_stride, x_org, out, out_stride, lane, warp);
}

// Vertical polyphase of NP int32 planes (row stride `stride` bytes) to 8-bit planes: 4 adjacent
// columns per thread (16-byte shared loads, 4-byte stores).  vtab: this tile's rows of the filter
// staged in shared memory, per destination row [pos - r0, coef[0..size)] as int32 / int16.
template <int NP>
__device__ __forceinline__ void vpass(int size, uint32_t vtab, int vrow_bytes, const uint32_t (&in)[NP], int stride, int d0y, int n_rows, int n_cols,
                                      uint8_t *const (&dst)[NP], const int (&dst_stride)[NP], bool vec, int tid, bool interleave = false) {
  const int ngrp = (n_cols + 3) >> 2;
  const int total = n_rows * ngrp;
  const int sh = (ngrp & (ngrp - 1)) == 0 ? __ffs(ngrp) - 1 : -1;
  for (int idx = tid; idx < total; idx += RS_THREADS) {
    const int ry = sh >= 0 ? idx >> sh : idx / ngrp;
    const int gx = idx - ry * ngrp;
    const uint32_t vt = vtab + ry * vrow_bytes;
    int pos;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(pos) : "r"(vt));
    uint32_t a[NP];
    int v[NP][4];
#pragma unroll
    for (int p = 0; p < NP; p++) {
      a[p] = in[p] + pos * stride + gx * 16;
      v[p][0] = v[p][1] = v[p][2] = v[p][3] = size == 1 ? 64 : 64 << 12;
    }
    if (size == 1) {
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const int4 w = lds_v4(a[p]);
        v[p][0] = (v[p][0] + w.x) >> 7; v[p][1] = (v[p][1] + w.y) >> 7; v[p][2] = (v[p][2] + w.z) >> 7; v[p][3] = (v[p][3] + w.w) >> 7;
      }
    } else {
      for (int j = 0; j < size; j++) {
        const int c = lds_s16(vt + 4 + 2 * j);
#pragma unroll
        for (int p = 0; p < NP; p++) {
          const int4 w = lds_v4(a[p]);
          v[p][0] += w.x * c; v[p][1] += w.y * c; v[p][2] += w.z * c; v[p][3] += w.w * c;
          a[p] += stride;
        }
      }
#pragma unroll
      for (int p = 0; p < NP; p++) { v[p][0] >>= 19; v[p][1] >>= 19; v[p][2] >>= 19; v[p][3] >>= 19; }
    }
    if (NP == 2 && interleave) {  // NV12: the two planes go out as U0 V0 U1 V1 ... into dst[0]
      uint8_t *q = dst[0] + (size_t)(d0y + ry) * dst_stride[0] + 8 * gx;
      uint8_t b[8];
#pragma unroll
      for (int k = 0; k < 4; k++) { b[2 * k] = (uint8_t)clip8(v[0][k]); b[2 * k + 1] = (uint8_t)clip8(v[NP - 1][k]); }
      if (vec && 4 * gx + 4 <= n_cols) {
        *(uint2 *)q = make_uint2(b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24), b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24));
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (4 * gx + (k >> 1) < n_cols) q[k] = b[k];
      }
      continue;
    }
#pragma unroll
    for (int p = 0; p < NP; p++) {
      uint8_t *q = dst[p] + (size_t)(d0y + ry) * dst_stride[p] + 4 * gx;
      if (vec && 4 * gx + 4 <= n_cols) {
        *(uint32_t *)q = (uint32_t)clip8(v[p][0]) | ((uint32_t)clip8(v[p][1]) << 8) | ((uint32_t)clip8(v[p][2]) << 16) | ((uint32_t)clip8(v[p][3]) << 24);
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (4 * gx + k < n_cols) q[k] = (uint8_t)clip8(v[p][k]);
      }
    }
  }
}

// stage this tile's rows of a vertical filter: per destination row {int32 pos - r0, int16 coef[size]}
__device__ __forceinline__ void stage_vtab(const DevFilter &f, int d0y, int n_rows, int r0, uint8_t *tab, int vrow_bytes, int tid) {
  const int per = 1 + f.size;
  for (int i = tid; i < n_rows * per; i += RS_THREADS) {
    const int ry = i / per, k = i - ry * per;
    if (k == 0) *(int32_t *)(tab + ry * vrow_bytes) = f.pos[d0y + ry] - r0;
    else *(int16_t *)(tab + ry * vrow_bytes + 4 + 2 * (k - 1)) = f.coef[(size_t)(d0y + ry) * f.size + k - 1];
  }
}

}  // namespace

template <int BPP>
__global__ void __launch_bounds__(RS_THREADS, 3) k_resize_tiles(const DevJob *__restrict__ jobs, int n_jobs, const RsLayout L) {
  extern __shared__ __align__(16) uint8_t smem[];
  int tile;
  const DevJob *jp = find_job(jobs, n_jobs, blockIdx.x, &tile);
  // a launch serves the jobs of one pixel class that share one shared-memory carve-up (kernel parameter:
  // constant-bank operands; derived per tile in registers it cost 6 % of the kernel's instructions)
  if (!jp->general || jp->rz_ok || jp->bpp != BPP || !same_layout(jp->rs_lay, L)) return;
  const DevJob &jb = *jp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = RS_THREADS / 32;
  const int tx = tile % jb.tiles_x, ty = tile / jb.tiles_x;
  const int W = jb.W, Wd = jb.Wd, Hd = jb.Hd;
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  const bool half = jb.half != 0;
  const bool want_depth = jb.dy != nullptr;

  Tile t;
  t.dx0 = tx * jb.rs_tw; t.dw = min(jb.rs_tw, Wd - t.dx0); t.dwp = (t.dw + 3) & ~3;
  t.dy0 = ty * jb.rs_th; t.dh = min(jb.rs_th, Hd - t.dy0);
  t.cx0 = t.dx0 >> 1; t.dcw = min((t.dx0 + t.dw + 1) >> 1, cdW) - t.cx0; t.dcwp = (t.dcw + 3) & ~3;
  t.cy0 = t.dy0 >> 1; t.dch = min((t.dy0 + t.dh + 1) >> 1, cdH) - t.cy0;
  {
    const int4 wx = __ldg((const int4 *)jb.rs_win_x + tx), wy = __ldg((const int4 *)jb.rs_win_y + ty);
    t.lc0 = wx.x; t.lc1 = wx.y; t.cc0 = wx.z; t.cc1 = wx.w;
    t.lr0 = wy.x; t.lr1 = wy.y; t.cr0 = wy.z; t.cr1 = wy.w;
  }
  const int pc0 = half ? t.cc0 * 2 : t.cc0, pc1 = half ? t.cc1 * 2 : t.cc1;
  t.wx0 = min(t.lc0, pc0) & ~3;
  t.ww = ((max(t.lc1, pc1) - t.wx0) + 3) & ~3;
  t.wy0 = min(t.lr0, t.cr0);
  t.wh = max(t.lr1, t.cr1) - t.wy0;
  t.cww = half ? t.ww >> 1 : t.ww;
  const int nl = t.lr1 - t.lr0, nc = t.cr1 - t.cr0;

  int16_t *s_y14 = (int16_t *)(smem + L.y14);  // [wh][ww]
  int16_t *s_u14 = (int16_t *)(smem + L.u14);  // [wh][cww]
  int16_t *s_v14 = (int16_t *)(smem + L.v14);
  uint8_t *s_dep = smem + L.dep;               // [wh][ww]
  const int mw = (t.ww >> 5) + 1;
  uint32_t *s_mask = (uint32_t *)(smem + L.mask);  // [wh][mw] overlay bits
  int *s_hits = (int *)(smem + L.hits);
  int *s_nhits = s_hits + HIT_CAP;

  if (tid < jb.n_src) {
    SrcDesc *sdw = (SrcDesc *)(smem + L.src);
    sdw[tid].rgb = jb.src[tid].rgb; sdw[tid].depth = jb.src[tid].depth;
    sdw[tid].rgb_stride = jb.src[tid].rgb_stride; sdw[tid].depth_stride = jb.src[tid].depth_stride;
  }
  __syncthreads();

  // ---- overlay bit mask of the window (render_text.cc:94-106: coverage != 0 -> white) --------
  bool has_text = false;
  if (jb.n_glyphs > 0) {
    for (int i = tid; i < t.wh * mw; i += RS_THREADS) s_mask[i] = 0;
    const int x1 = t.wx0 + t.ww, y1 = t.wy0 + t.wh;
    int any = 0;
    // only the row bands this window can touch (the host bucketed the glyph list by band)
    const int b0 = max(t.wy0 - jb.glyph_max_h, 0) >> jb.glyph_band_shift, b1 = (y1 - 1) >> jb.glyph_band_shift;
    const int g_begin = jb.glyph_band[b0], g_end = jb.glyph_band[b1 + 1];
    for (int base = g_begin; base < g_end; base += HIT_CAP) {
      if (tid == 0) *s_nhits = 0;
      __syncthreads();
      for (int g = base + tid; g < min(base + HIT_CAP, g_end); g += RS_THREADS) {
        const DevPlaced pg = jb.glyphs[g];
        if (pg.x < x1 && pg.x + pg.w > t.wx0 && pg.y < y1 && pg.y + pg.h > t.wy0) s_hits[atomicAdd(s_nhits, 1)] = g;
      }
      __syncthreads();
      const int nh = *s_nhits;
      any |= nh;
      for (int h = warp; h < nh; h += NWARP) {
        const DevPlaced pg = jb.glyphs[s_hits[h]];
        const int q0 = max(0, t.wy0 - pg.y), q1 = min(pg.h, y1 - pg.y);
        const int p0 = max(0, t.wx0 - pg.x), p1 = min(pg.w, x1 - pg.x);
        // the atlas holds one bit per glyph pixel (DevPlaced): a lane takes one word of one visible row and ORs
        // it, shifted to the window's column grid, into the tile's overlay mask
        const int bit_lo = pg.bit0 + p0, bit_hi = pg.bit0 + p1;
        const int w_lo = bit_lo >> 5, w_hi = (bit_hi - 1) >> 5, nw = w_hi - w_lo + 1;
        for (int i = lane; i < (q1 - q0) * nw; i += 32) {
          const int qq = i / nw, wi = w_lo + (i - qq * nw), q = q0 + qq;
          uint32_t m = __ldg(jb.atlas + pg.mask_off + (uint32_t)(q * pg.wpr + wi));
          const int lo = max(bit_lo - 32 * wi, 0), hi = min(bit_hi - 32 * wi, 32);
          m &= (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (32 - hi));
          if (m == 0) continue;
          const int xb = pg.x - pg.bit0 + 32 * wi - t.wx0;  // window column of bit 0 of this word (negative only for masked-off bits)
          const int wd = xb >> 5, sh = xb & 31;
          uint32_t *mrow = s_mask + (pg.y + q - t.wy0) * mw;
          const uint32_t lo_part = m << sh;
          if (lo_part && wd >= 0) atomicOr(&mrow[wd], lo_part);
          if (sh) {
            const uint32_t hi_part = m >> (32 - sh);
            if (hi_part) atomicOr(&mrow[wd + 1], hi_part);
          }
        }
      }
      __syncthreads();
    }
    has_text = any != 0;
  }

  // ---- stage A: source window -> 14-bit planes + depth bytes ------------------------------------
  {
    const uint32_t ky0 = jb.ky[0], ky1 = jb.ky[1], ku0 = jb.ku[0], ku1 = jb.ku[1], kv0 = jb.kv[0], kv1 = jb.kv[1];
    const uint32_t white = BPP == 4 ? (0x00FFFFFFu << (8 * jb.rgb_base)) : 0x00FFFFFFu;
    const int ngrp = t.ww >> 2;
    // the aligned fast path needs every source's depth plane (a lone source without a depth stream has none)
    const bool fast = BPP == 4 && jb.in_vec && (jb.n_src > 1 || want_depth);
    const int n_src = jb.n_src;
    const uint32_t a_mask = jb.a_off >= 0 ? (0xFFu << (8 * jb.a_off)) : 0xFFFFFFFFu;
    const SrcDesc *sd = (const SrcDesc *)(smem + L.src);
    // a lane owns one column of 4-pixel groups (two when the window is wider than 128 pixels)
    for (int g = lane; g < ngrp; g += 32) {
      const int x = t.wx0 + 4 * g;
      const bool inside = x + 4 <= W;
      const uint32_t mshift = (4 * g) & 31;
      const uint32_t *mrow = s_mask + (g >> 3);
      for (int r = warp; r < t.wh; r += NWARP) {
        const int y = t.wy0 + r;
        uint32_t px[4], d4;
        if (fast && inside) load_group_fast4(sd, n_src, a_mask, x * 4, x, y, px, &d4);
        else load_group<BPP>(jb, x, y, want_depth, px, &d4);
        if (has_text) {
          const uint32_t m = (mrow[r * mw] >> mshift) & 15u;
#pragma unroll
          for (int i = 0; i < 4; i++)
            if ((m >> i) & 1u) px[i] |= white;
        }
        int yv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) yv[i] = dot_px(ky0, ky1, px[i], (32 << 14) + (1 << 8)) >> 9;
        *(uint2 *)(s_y14 + r * t.ww + 4 * g) = make_uint2(pack16(yv[0], yv[1]), pack16(yv[2], yv[3]));
        if (half) {
          int u[2], v[2];
#pragma unroll
          for (int p = 0; p < 2; p++) {
            u[p] = dot_px(ku0, ku1, px[2 * p + 1], dot_px(ku0, ku1, px[2 * p], C_BIAS)) >> 10;
            v[p] = dot_px(kv0, kv1, px[2 * p + 1], dot_px(kv0, kv1, px[2 * p], C_BIAS)) >> 10;
          }
          *(uint32_t *)(s_u14 + r * t.cww + 2 * g) = pack16(u[0], u[1]);
          *(uint32_t *)(s_v14 + r * t.cww + 2 * g) = pack16(v[0], v[1]);
        } else {
          int u[4], v[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            u[i] = dot_px(ku0, ku1, px[i], C1_BIAS) >> 9;
            v[i] = dot_px(kv0, kv1, px[i], C1_BIAS) >> 9;
          }
          *(uint2 *)(s_u14 + r * t.cww + 4 * g) = make_uint2(pack16(u[0], u[1]), pack16(u[2], u[3]));
          *(uint2 *)(s_v14 + r * t.cww + 4 * g) = make_uint2(pack16(v[0], v[1]), pack16(v[2], v[3]));
        }
        if (want_depth) *(uint32_t *)(s_dep + r * t.ww + 4 * g) = d4;
      }
    }
  }
  __syncthreads();

  // ---- stage H: horizontal polyphase, hScale16To15 (>>13, clamp 32767) -> int32 rows --------------
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  const int vl_row = (4 + 2 * jb.vl.size + 3) & ~3, vc_row = (4 + 2 * jb.vc.size + 3) & ~3;
  stage_vtab(jb.vl, t.dy0, t.dh, t.lr0, smem + L.vtab, vl_row, tid);
  stage_vtab(jb.vc, t.cy0, t.dch, t.cr0, smem + L.vtab + t.dh * vl_row, vc_row, tid);
  {
    const uint32_t in_y[1] = {sb + L.y14 + (uint32_t)((t.lr0 - t.wy0) * t.ww) * 2}, out_y[1] = {sb + L.hy};
    hpass_any<1, 2, H_SCENE>(jb.hl, t.dx0, t.dw, nl, in_y, t.ww * 2, t.wx0, out_y, t.dwp * 4, lane, warp);
    const int corg = half ? (t.wx0 >> 1) : t.wx0;
    const uint32_t in_c[2] = {sb + L.u14 + (uint32_t)((t.cr0 - t.wy0) * t.cww) * 2, sb + L.v14 + (uint32_t)((t.cr0 - t.wy0) * t.cww) * 2};
    const uint32_t out_c[2] = {sb + L.hu, sb + L.hv};
    hpass_any<2, 2, H_SCENE>(jb.hc, t.cx0, t.dcw, nc, in_c, t.cww * 2, corg, out_c, t.dcwp * 4, lane, warp);
  }
  __syncthreads();

  // ---- stage V: vertical polyphase to 8 bit (yuv2planeX / yuv2plane1) ----------------------------
  const bool vec = jb.out_vec != 0;
  {
    const uint32_t in_y[1] = {sb + L.hy};
    uint8_t *const dst_y[1] = {jb.sy + t.dx0};
    const int ds_y[1] = {jb.sys};
    vpass<1>(jb.vl.size, sb + L.vtab, vl_row, in_y, t.dwp * 4, t.dy0, t.dh, t.dw, dst_y, ds_y, vec, tid);
    const uint32_t in_c[2] = {sb + L.hu, sb + L.hv};
    const bool nv12 = jb.nv12 != 0;
    uint8_t *const dst_c[2] = {jb.su + (nv12 ? 2 * t.cx0 : t.cx0), nv12 ? nullptr : jb.sv + t.cx0};
    const int ds_c[2] = {jb.sus, jb.svs};
    vpass<2>(jb.vc.size, sb + L.vtab + t.dh * vl_row, vc_row, in_c, t.dcwp * 4, t.cy0, t.dch, t.dcw, dst_c, ds_c, vec, tid, nv12);
  }

  // ---- depth: hScale8To15 (>>7) -> range compression -> vertical; U = V = 128 ----------------------
  if (want_depth) {
    __syncthreads();  // the luma H rows are reused
    const uint32_t in_d[1] = {sb + L.dep + (uint32_t)((t.lr0 - t.wy0) * t.ww)}, out_d[1] = {sb + L.hy};
    hpass_any<1, 1, H_DEPTH>(jb.hl, t.dx0, t.dw, nl, in_d, t.ww, t.wx0, out_d, t.dwp * 4, lane, warp);
    __syncthreads();
    uint8_t *const dst_d[1] = {jb.dy + t.dx0};
    const int ds_d[1] = {jb.dys};
    vpass<1>(jb.vl.size, sb + L.vtab, vl_row, out_d, t.dwp * 4, t.dy0, t.dh, t.dw, dst_d, ds_d, vec, tid);
    // depth chroma is constant 128; NV12: one plane of 2*dcw bytes per row
    const bool nv12 = jb.nv12 != 0;
    const int rowb = nv12 ? 2 * t.dcw : t.dcw, xoff = nv12 ? 2 * t.cx0 : t.cx0;
    const int ngrp = (rowb + 3) >> 2;
    for (int idx = tid; idx < t.dch * ngrp; idx += RS_THREADS) {
      const int ry = idx / ngrp, gx = idx - ry * ngrp;
      uint8_t *pu = jb.du + (size_t)(t.cy0 + ry) * jb.dus + xoff + 4 * gx;
      uint8_t *pv = nv12 ? nullptr : jb.dv + (size_t)(t.cy0 + ry) * jb.dvs + xoff + 4 * gx;
      if (vec && 4 * gx + 4 <= rowb) {
        *(uint32_t *)pu = 0x80808080u;
        if (pv) *(uint32_t *)pv = 0x80808080u;
      } else {
        for (int k = 0; k < 4; k++)
          if (4 * gx + k < rowb) { pu[k] = 128; if (pv) pv[k] = 128; }
      }
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
  if (int r = resize_strips_init()) return r;
  g_resize_smem_cap = RS_SMEM_MAX;
  e = cudaFuncSetAttribute(k_resize_tiles<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_resize_smem_cap);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_resize_tiles<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_resize_smem_cap);
  if (e != cudaSuccess) return (int)e;
  return 0;
}

int launch_resize_tiles(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream) {
  int total = 0;
  for (int j = 0; j < n_jobs; j++) total = jobs_host[j].tile_base + jobs_host[j].tiles_x * jobs_host[j].tiles_y;
  if (total == 0) return 0;
  // one launch per (pixel class, carve-up): CTAs of other jobs return at once.  A batch normally holds one
  // size pair, i.e. one launch.
  int launches = 0;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs_host[j];
    if (!jb.general || jb.rz_ok) continue;
    bool seen = false;
    for (int i = 0; i < j && !seen; i++) {
      const DevJob &o = jobs_host[i];
      seen = o.general && !o.rz_ok && o.bpp == jb.bpp && same_layout(o.rs_lay, jb.rs_lay);
    }
    if (seen) continue;
    const int sm = (jb.rs_lay.total + 1023) & ~1023;
    if (sm > g_resize_smem_cap) return -1;
    if (jb.bpp == 3) k_resize_tiles<3><<<total, RS_THREADS, sm, (cudaStream_t)stream>>>(jobs_dev, n_jobs, jb.rs_lay);
    else k_resize_tiles<4><<<total, RS_THREADS, sm, (cudaStream_t)stream>>>(jobs_dev, n_jobs, jb.rs_lay);
    launches++;
  }
  return launches;
}

}  // namespace nes
