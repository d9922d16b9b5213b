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

// ---------------------------------------------------------------------------
// k_frame_tiles: same-size fused path.
// shared memory: s_in  [TILE_ROWS][TILE_W*BPP] packed pixels (tile + halo rows)
//                s_uv  [TILE_ROWS][TILE_W/2]   u15 | v15<<16 per (source row, chroma col)
//                s_hits[HIT_CAP], s_nhits
// ---------------------------------------------------------------------------
template <int BPP>
struct FrameTileSmem {
  static constexpr int ROWB = TILE_W * BPP;
  static constexpr int IN_BYTES = TILE_ROWS * ROWB;
  static constexpr int UV_BYTES = TILE_ROWS * (TILE_W / 2) * 4;
  static constexpr int HIT_BYTES = HIT_CAP * 4 + 16;
  static constexpr int TOTAL = IN_BYTES + UV_BYTES + HIT_BYTES;
};

template <int BPP>
__global__ void __launch_bounds__(CTA_THREADS, (BPP == 3 ? 4 : 3))
k_frame_tiles(const DevJob *__restrict__ jobs, int n_jobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  using L = FrameTileSmem<BPP>;
  uint8_t *s_in = smem;
  uint32_t *s_uv = (uint32_t *)(smem + L::IN_BYTES);
  int *s_hits = (int *)(smem + L::IN_BYTES + L::UV_BYTES);
  int *s_nhits = s_hits + HIT_CAP;

  int tile;
  const DevJob *jp = find_job(jobs, n_jobs, blockIdx.x, &tile);
  if (jp->bpp != BPP || jp->general) return;  // other template / general-path job
  const DevJob &jb = *jp;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tile % jb.tiles_x, ty = tile / jb.tiles_x;
  const int W = jb.W, H = jb.H;
  const int x0 = tx * TILE_W, y0 = ty * TILE_H;
  const int tw = min(TILE_W, W - x0), th = min(TILE_H, H - y0);
  const int oy = y0 - HALO;                       // frame row of tile row 0
  const int ya = max(oy, 0), yb = min(y0 + th + HALO, H);  // rows actually present
  const bool composite = jb.n_src > 1;
  const bool vec_in = jb.in_vec != 0, vec_out = jb.out_vec != 0;

  // ---- stage 1: packed pixels of the tile (+halo) into shared memory ----------
  if (!composite) {
    const uint8_t *src = jb.src[0].rgb;
    const int stride = jb.src[0].rgb_stride;
    const int nbytes = tw * BPP;
    for (int y = ya + warp; y < yb; y += CTA_THREADS / 32) {
      const uint8_t *g = src + (size_t)y * stride + (size_t)x0 * BPP;
      uint8_t *s = s_in + (y - oy) * L::ROWB;
      if (vec_in) {
        const int nvec = nbytes >> 4;
        for (int i = lane; i < nvec; i += 32) cp_async16(s + i * 16, g + i * 16);
        for (int i = (nvec << 4) + lane; i < nbytes; i += 32) s[i] = g[i];
      } else {
        for (int i = lane; i < nbytes; i += 32) s[i] = g[i];
      }
    }
    // depth stream is pointwise: do it while the cp.async traffic is in flight
    if (jb.dy) {
      const uint8_t *dsrc = jb.src[0].depth;
      const int dstride = jb.src[0].depth_stride;
      for (int r = warp; r < th; r += CTA_THREADS / 32) {
        const int y = y0 + r;
        const uint8_t *g = dsrc + (size_t)y * dstride + x0;
        uint8_t *o = jb.dy + (size_t)y * jb.dys + x0;
        const int x = lane * 8;
        if (vec_in && vec_out && x + 8 <= tw) {
          const uint2 d = __ldg((const uint2 *)(g + x));
          *(uint2 *)(o + x) = make_uint2(gray_y4(d.x), gray_y4(d.y));
        } else {
          for (int i = x; i < min(x + 8, tw); i++) o[i] = (uint8_t)gray_y(g[i]);
        }
      }
    }
    cp_async_wait_all();
  } else {
    // composite: select per pixel among the sources, keep the winner's bytes in the
    // tile and convert the winner's depth on the fly (core rows only).
    for (int y = ya + warp; y < yb; y += CTA_THREADS / 32) {
      uint8_t *s = s_in + (y - oy) * L::ROWB;
      const bool core = (y >= y0) && (y < y0 + th);
      if (BPP == 4 && vec_in && vec_out && (tw & 3) == 0) {
        for (int x = lane * 4; x < tw; x += 128) {
          uint4 px; uint32_t d4;
          composite_px4(jb, x0 + x, y, &px, &d4);
          *(uint4 *)(s + x * 4) = px;
          if (core && jb.dy) *(uint32_t *)(jb.dy + (size_t)y * jb.dys + x0 + x) = gray_y4(d4);
        }
      } else {
        for (int x = lane; x < tw; x += 32) {
          uint32_t d;
          composite_px<BPP>(jb, x0 + x, y, s + x * BPP, &d);
          if (core && jb.dy) jb.dy[(size_t)y * jb.dys + x0 + x] = (uint8_t)gray_y(d);
        }
      }
    }
  }
  // depth chroma planes are constant 128
  if (jb.dy) {
    for (int r = warp; r < (th >> 1); r += CTA_THREADS / 32) {
      const int ci = (y0 >> 1) + r;
      uint8_t *ou = jb.du + (size_t)ci * jb.dus + (x0 >> 1);
      uint8_t *ov = jb.dv + (size_t)ci * jb.dvs + (x0 >> 1);
      const int c = lane * 4;
      if (vec_out && c + 4 <= (tw >> 1)) {
        *(uint32_t *)(ou + c) = 0x80808080u;
        *(uint32_t *)(ov + c) = 0x80808080u;
      } else {
        for (int i = c; i < min(c + 4, tw >> 1); i++) { ou[i] = 128; ov[i] = 128; }
      }
    }
  }
  __syncthreads();

  // ---- stage 2: text overlay, stamped into the shared tile ---------------------
  if (jb.n_glyphs > 0) stamp_glyphs<BPP>(jb, s_in, L::ROWB, x0, oy, x0, x0 + tw, ya, yb, s_hits, s_nhits);

  // ---- stage 3: per source row: Y out, pair-summed chroma to s_uv -------------
  {
    const int cy0 = jb.cy[0], cy1 = jb.cy[1], cy2 = jb.cy[2];
    const int cu0 = jb.cu[0], cu1 = jb.cu[1], cu2 = jb.cu[2];
    const int cv0 = jb.cv[0], cv1 = jb.cv[1], cv2 = jb.cv[2];
    const int rb = jb.rgb_base;
    const int x = lane * 8;
    if (x < tw) {
      for (int y = ya + warp; y < yb; y += CTA_THREADS / 32) {
        const int tr = y - oy;
        uint32_t c[8][3];
        if (BPP == 3) {
          const uint2 *p = (const uint2 *)(s_in + tr * L::ROWB + lane * 24);
          const uint2 a = p[0], b = p[1], d = p[2];
          const uint32_t w[6] = {a.x, a.y, b.x, b.y, d.x, d.y};
#pragma unroll
          for (int k = 0; k < 8; k++)
#pragma unroll
            for (int j = 0; j < 3; j++) c[k][j] = byte_at(w[(3 * k + j) >> 2], (3 * k + j) & 3);
        } else {
          const uint4 *p = (const uint4 *)(s_in + tr * L::ROWB + lane * 32);
          const uint4 a = p[0], b = p[1];
          const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int k = 0; k < 8; k++)
#pragma unroll
            for (int j = 0; j < 3; j++) c[k][j] = byte_at(w[k], rb + j);
        }
        if (y >= y0 && y < y0 + th) {
          uint32_t yv[8];
#pragma unroll
          for (int k = 0; k < 8; k++)
            yv[k] = (uint32_t)(cy0 * (int)c[k][0] + cy1 * (int)c[k][1] + cy2 * (int)c[k][2] + Y_BIAS) >> 15;
          uint8_t *o = jb.sy + (size_t)y * jb.sys + x0 + x;
          if (vec_out && x + 8 <= tw) {
            uint2 v;
            v.x = yv[0] | (yv[1] << 8) | (yv[2] << 16) | (yv[3] << 24);
            v.y = yv[4] | (yv[5] << 8) | (yv[6] << 16) | (yv[7] << 24);
            *(uint2 *)o = v;
          } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
              if (x + k < tw) o[k] = (uint8_t)yv[k];
          }
        }
        uint32_t uv[4];
#pragma unroll
        for (int p2 = 0; p2 < 4; p2++) {
          const int s0 = (int)(c[2 * p2][0] + c[2 * p2 + 1][0]);
          const int s1 = (int)(c[2 * p2][1] + c[2 * p2 + 1][1]);
          const int s2 = (int)(c[2 * p2][2] + c[2 * p2 + 1][2]);
          const uint32_t u = ((uint32_t)(cu0 * s0 + cu1 * s1 + cu2 * s2 + C_BIAS) >> 9) & 0xFFFEu;
          const uint32_t v = ((uint32_t)(cv0 * s0 + cv1 * s1 + cv2 * s2 + C_BIAS) >> 9) & 0xFFFEu;
          uv[p2] = u | (v << 16);
        }
        *(uint4 *)(s_uv + tr * (TILE_W / 2) + lane * 4) = make_uint4(uv[0], uv[1], uv[2], uv[3]);
      }
    }
  }
  __syncthreads();

  // ---- stage 4: 8-tap vertical bicubic on chroma (edge taps fold = clamped rows) --
  // T = [-58,-172,492,1786,1786,492,-172,-58]/4096, symmetric: pair the taps first (the
  // packed u|v words add without carry: 2*32767 < 65536).
  {
    const int c = lane * 4;  // chroma column inside the tile
    if (c < (tw >> 1)) {
      for (int r = warp; r < (th >> 1); r += CTA_THREADS / 32) {
        const int ci = (y0 >> 1) + r;
        uint32_t t[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int sr = min(max(2 * ci - 3 + j, 0), H - 1) - oy;
          const uint4 q = *(const uint4 *)(s_uv + sr * (TILE_W / 2) + c);
          t[j][0] = q.x; t[j][1] = q.y; t[j][2] = q.z; t[j][3] = q.w;
        }
        uint32_t ub = 0, vb = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t a07 = t[0][k] + t[7][k];
          const uint32_t a16 = t[1][k] + t[6][k];
          const uint32_t a25 = t[2][k] + t[5][k];
          const uint32_t a34 = t[3][k] + t[4][k];
          const int su = (64 << 12) + 1786 * (int)(a34 & 0xFFFFu) + 492 * (int)(a25 & 0xFFFFu) -
                         172 * (int)(a16 & 0xFFFFu) - 58 * (int)(a07 & 0xFFFFu);
          const int sv = (64 << 12) + 1786 * (int)(a34 >> 16) + 492 * (int)(a25 >> 16) - 172 * (int)(a16 >> 16) -
                         58 * (int)(a07 >> 16);
          ub |= (uint32_t)clip8(su >> 19) << (8 * k);
          vb |= (uint32_t)clip8(sv >> 19) << (8 * k);
        }
        uint8_t *ou = jb.su + (size_t)ci * jb.sus + (x0 >> 1) + c;
        uint8_t *ov = jb.sv + (size_t)ci * jb.svs + (x0 >> 1) + c;
        if (vec_out && c + 4 <= (tw >> 1)) {
          *(uint32_t *)ou = ub;
          *(uint32_t *)ov = vb;
        } else {
          for (int k = 0; k < 4; k++)
            if (c + k < (tw >> 1)) { ou[k] = (uint8_t)(ub >> (8 * k)); ov[k] = (uint8_t)(vb >> (8 * k)); }
        }
      }
    }
  }
}

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
  e = cudaFuncSetAttribute(k_frame_tiles<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrameTileSmem<3>::TOTAL);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_frame_tiles<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrameTileSmem<4>::TOTAL);
  if (e != cudaSuccess) return (int)e;
  g_resize_smem_cap = 200 * 1024;
  e = cudaFuncSetAttribute(k_resize_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, g_resize_smem_cap);
  if (e != cudaSuccess) return (int)e;
  return 0;
}

int launch_frame_tiles(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream) {
  int total = 0;
  bool any3 = false, any4 = false;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs_host[j];
    total = jb.tile_base + jb.tiles_x * jb.tiles_y;
    if (!jb.general) (jb.bpp == 3 ? any3 : any4) = true;
  }
  if (total == 0) return 0;
  int launches = 0;
  if (any3) {
    k_frame_tiles<3><<<total, CTA_THREADS, FrameTileSmem<3>::TOTAL, (cudaStream_t)stream>>>(jobs_dev, n_jobs);
    launches++;
  }
  if (any4) {
    k_frame_tiles<4><<<total, CTA_THREADS, FrameTileSmem<4>::TOTAL, (cudaStream_t)stream>>>(jobs_dev, n_jobs);
    launches++;
  }
  return launches;
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
