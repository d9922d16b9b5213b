// frame_tiles.cu -- k_frame_tiles: the fused same-size frame kernel (the hot kernel).
//
// Per frame, in ONE launch for a whole batch of frames:
//   [depth-select composite of N sources] -> glyph stamp overlay -> Y (pointwise)
//   + horizontally pair-summed chroma -> 8-tap vertical bicubic -> U,V planes, and the depth
//   stream GRAY8 -> Y (range compression) with U = V = 128.
// Arithmetic: libswscale's C path as driven by the reference
// (/root/reference/src/base/video/type_managers.cc:143-155 via rendered_frame.h:24-33, after
// the overlay of render_text.cc:81-110); integer spec in SURVEY.md Appendix A.2 / A.4.
//
// Shape of the kernel (HBM-bound u8/int32 streaming work, no tensor cores):
//   * persistent CTAs (3 per SM), each walks tiles  blockIdx.x, +gridDim.x, ...
//   * a tile is TILE_W x TILE_H source pixels plus 3 halo rows above and below (the 8-tap
//     vertical chroma filter); its packed-pixel rows are staged in shared memory by bulk
//     async copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), double
//     buffered: warp 0 issues the copies of the NEXT tile before the CTA starts computing
//     the current one, so the loads of tile t+1 overlap the arithmetic of tile t;
//   * the overlay is stamped into the shared tile (only tiles whose bit is set in the job's
//     tile mask look at the glyph list at all);
//   * phase A (warp per source row): Y with two dp2a per pixel, pair-summed chroma with
//     eight dp2a per pixel pair, 128-bit coalesced Y stores; the 15-bit chroma rows are
//     written IN PLACE over the pixel row the warp has just consumed;
//   * phase B (warp per chroma row): symmetric 8-tap filter on packed (u|v<<16) words,
//     conflict-free 128-bit shared loads, coalesced 32-bit U/V stores;
//   * depth is pointwise: its loads are issued before phase A and consumed after phase B.
// The composite variant selects among the sources on the fly while filling the shared tile.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_common.cuh"
#include "nes_internal.h"

namespace nes {

namespace {

// Everything a CTA needs to know about one tile, written to shared memory by lane 0 of
// warp 0 while the previous tile is being computed (so no thread reads the job descriptor
// from global memory on the critical path).  Plane pointers are pre-offset to the tile.
struct TileCtx {
  int32_t valid, tma, stamp, n_src;
  int32_t x0, y0, tw, th, ya, yb, H, job;
  int32_t vec_in, vec_out, depth_vec, a_shift;
  uint32_t kya, kyb, kua, kub, kva, kvb;
  int32_t sys, sus, svs, dys, dus, dvs;
  uint8_t *sy, *su, *sv, *dy, *du, *dv;  // + tile column offset
  const uint8_t *rgb[NES_MAX_SOURCES];   // + tile column offset
  const uint8_t *dep[NES_MAX_SOURCES];
  int32_t rs[NES_MAX_SOURCES], ds[NES_MAX_SOURCES];
};

template <int BPP>
struct TileSmem {
  static constexpr int ROWB = TILE_W * BPP;
  static constexpr int PX_BYTES = TILE_ROWS * ROWB;
  static constexpr int OFF_CTX = 2 * PX_BYTES;
  static constexpr int OFF_BAR = OFF_CTX + 2 * (int)sizeof(TileCtx);
  static constexpr int OFF_Q = OFF_BAR + 16;
  static constexpr int OFF_HITS = OFF_Q + 16;
  static constexpr int TOTAL = OFF_HITS + HIT_CAP * 4 + 16;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk async copy (TMA engine, no tensor map), completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// explicit shared-space accesses (32-bit shared addresses; keeps the hot loops off generic LD/ST)
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ int atoms_inc(uint32_t a) {
  int v;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1   (a: signed 16-bit halves, b: unsigned bytes)
__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// d = c + a.lo16 * b.byte2 + a.hi16 * b.byte3
__device__ __forceinline__ int dp2a_hi(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// max(min(v, 255), 0) in one instruction
__device__ __forceinline__ uint32_t clip8_relu(int v) {
  int d;
  asm("min.s32.relu %0, %1, %2;" : "=r"(d) : "r"(v), "r"(255));
  return (uint32_t)d;
}

// GRAY8 -> limited-range luma for the 4 bytes of a word:
//   Y = (d*219 + 127)/255 + 16, two pixels per multiply in 16-bit lanes;
//   floor(t/255) == (t + (t >> 8) + 1) >> 8 for every t = d*219 + 127, d in 0..255
__device__ __forceinline__ uint32_t gray_y4_packed(uint32_t w) {
  const uint32_t p01 = __byte_perm(w, 0u, 0x4140), p23 = __byte_perm(w, 0u, 0x4342);  // d0 | d1<<16 ; d2 | d3<<16
  const uint32_t t01 = p01 * 219u + 0x007F007Fu, t23 = p23 * 219u + 0x007F007Fu;
  const uint32_t s01 = t01 + __byte_perm(t01, 0u, 0x4341) + 0x10011001u;  // + (t>>8) + 1 + (16<<8) per lane
  const uint32_t s23 = t23 + __byte_perm(t23, 0u, 0x4341) + 0x10011001u;
  return __byte_perm(s01, s23, 0x7531);  // byte 1 of every 16-bit lane
}

// Which job does global tile `t` belong to (tile_base is a prefix sum over the batch).
__device__ __forceinline__ int job_of_tile(const DevJob *jobs, int n_jobs, int t) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_base <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

}  // namespace

// One warp-wide grab from a shared work counter.
__device__ __forceinline__ int grab(uint32_t q_addr, int lane) {
  int v = 0;
  if (lane == 0) v = atoms_inc(q_addr);
  return __shfl_sync(0xffffffffu, v, 0);
}

// Depth-select composite of the 8 pixels a lane owns in a row (4 at column 4*lane, 4 at
// 128 + 4*lane) straight from the N sources: the winner's pixel words and depth bytes.
// Semantics: DESIGN.md "composite" / oracle/overlay_port.c nes_oracle_composite.
template <int N>
__device__ __forceinline__ void composite_fetch(const TileCtx &c, int y, int lane, int n_src, uint32_t (&p)[8], uint32_t (&d4)[2]) {
  uint32_t bd[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { p[i] = 0; bd[i] = 256; }
  const int ash = c.a_shift;
  const int cnt = N > 0 ? N : n_src;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : NES_MAX_SOURCES); k++) {
    if (k >= cnt) break;
    const uint8_t *rp = c.rgb[k] + (size_t)y * c.rs[k] + lane * 16;
    const uint8_t *dp = c.dep[k] + (size_t)y * c.ds[k] + lane * 4;
    uint4 q[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    uint32_t dw[2] = {0, 0};
    if (lane * 4 < c.tw) { q[0] = __ldg((const uint4 *)rp); dw[0] = __ldg((const uint32_t *)dp); }
    if (128 + lane * 4 < c.tw) { q[1] = __ldg((const uint4 *)(rp + 512)); dw[1] = __ldg((const uint32_t *)(dp + 128)); }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t w[4] = {q[h].x, q[h].y, q[h].z, q[h].w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t d = (dw[h] >> (8 * i)) & 255u;
        const bool take = ((w[i] >> ash) & 255u) != 0 && d < bd[4 * h + i];
        bd[4 * h + i] = take ? d : bd[4 * h + i];
        p[4 * h + i] = take ? w[i] : p[4 * h + i];
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; h++)
    d4[h] = min(bd[4 * h], 255u) | (min(bd[4 * h + 1], 255u) << 8) | (min(bd[4 * h + 2], 255u) << 16) | (min(bd[4 * h + 3], 255u) << 24);
}

template <int BPP>
__global__ void __launch_bounds__(CTA_THREADS, 3)
k_frame_tiles(const DevJob *__restrict__ jobs, int n_jobs, int total_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  using L = TileSmem<BPP>;
  TileCtx *s_ctx = (TileCtx *)(smem + L::OFF_CTX);
  uint64_t *s_bar = (uint64_t *)(smem + L::OFF_BAR);
  int *s_q = (int *)(smem + L::OFF_Q);  // [0] phase A rows, [1] phase B rows, [2] fill rows
  int *s_hits = (int *)(smem + L::OFF_HITS);
  int *s_nhits = s_hits + HIT_CAP;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = CTA_THREADS / 32;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t qA = smem_base + L::OFF_Q, qB = qA + 4, qF = qA + 8;

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_q[0] = s_q[1] = s_q[2] = 0;
  }
  __syncthreads();

  // warp 0: describe tile `t` in s_ctx[b] and start its bulk copies into buffer b
  auto prefetch = [&](int t, int b) {
    int tma = 0, ya = 0, yb = 0, tw = 0, y0 = 0;
    const uint8_t *src = nullptr;
    int stride = 0;
    if (lane == 0) {
      const int j = job_of_tile(jobs, n_jobs, t);
      const DevJob *jp = jobs + j;
      TileCtx &c = s_ctx[b];
      c.job = j;
      c.valid = (jp->bpp == BPP) && !jp->general;
      const int local = t - jp->tile_base;
      const int tx = local % jp->tiles_x, ty = local / jp->tiles_x;
      const int x0 = tx * TILE_W;
      y0 = ty * TILE_H;
      tw = min(TILE_W, jp->W - x0);
      const int th = min(TILE_H, jp->H - y0);
      ya = max(y0 - HALO, 0); yb = min(y0 + th + HALO, jp->H);
      c.x0 = x0; c.y0 = y0; c.tw = tw; c.th = th; c.ya = ya; c.yb = yb; c.H = jp->H;
      c.n_src = jp->n_src;
      tma = c.valid && jp->tma_ok;
      c.tma = tma;
      c.stamp = jp->n_glyphs > 0 && (!jp->use_mask || ((jp->tile_mask[local >> 5] >> (local & 31)) & 1u));
      c.vec_in = jp->in_vec; c.vec_out = jp->out_vec;
      c.depth_vec = jp->dy && jp->n_src == 1 && jp->in_vec && jp->out_vec && (tw & 7) == 0;
      c.a_shift = jp->a_off > 0 ? 8 * jp->a_off : 0;
      c.kya = jp->ky[0]; c.kyb = jp->ky[1]; c.kua = jp->ku[0]; c.kub = jp->ku[1]; c.kva = jp->kv[0]; c.kvb = jp->kv[1];
      c.sys = jp->sys; c.sus = jp->sus; c.svs = jp->svs; c.dys = jp->dys; c.dus = jp->dus; c.dvs = jp->dvs;
      c.sy = jp->sy + x0; c.su = jp->su + (x0 >> 1); c.sv = jp->sv + (x0 >> 1);
      c.dy = jp->dy ? jp->dy + x0 : nullptr;
      c.du = jp->dy ? jp->du + (x0 >> 1) : nullptr;
      c.dv = jp->dy ? jp->dv + (x0 >> 1) : nullptr;
      for (int k = 0; k < jp->n_src; k++) {
        c.rgb[k] = jp->src[k].rgb + (size_t)x0 * BPP;
        c.dep[k] = jp->src[k].depth ? jp->src[k].depth + x0 : nullptr;
        c.rs[k] = jp->src[k].rgb_stride; c.ds[k] = jp->src[k].depth_stride;
      }
      src = c.rgb[0]; stride = c.rs[0];
    }
    tma = __shfl_sync(0xffffffffu, tma, 0);
    if (tma) {
      ya = __shfl_sync(0xffffffffu, ya, 0); yb = __shfl_sync(0xffffffffu, yb, 0);
      tw = __shfl_sync(0xffffffffu, tw, 0); y0 = __shfl_sync(0xffffffffu, y0, 0);
      stride = __shfl_sync(0xffffffffu, stride, 0);
      src = (const uint8_t *)__shfl_sync(0xffffffffu, (unsigned long long)src, 0);
      uint8_t *buf = smem + b * L::PX_BYTES;
      const uint32_t rowb = (uint32_t)tw * BPP;
      const int oy = y0 - HALO;
      for (int y = ya + lane; y < yb; y += 32) bulk_g2s(buf + (y - oy) * L::ROWB, src + (size_t)y * stride, rowb, &s_bar[b]);
      if (lane == 0) mbar_arrive_expect_tx(&s_bar[b], rowb * (uint32_t)(yb - ya));
    } else if (lane == 0) {
      mbar_arrive_expect_tx(&s_bar[b], 0);  // nothing in flight: the CTA fills the tile itself
    }
  };

  if (warp == 0 && (int)blockIdx.x < total_tiles) prefetch(blockIdx.x, 0);

  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
    const int cur = it & 1;
    if (warp == 0 && tile + (int)gridDim.x < total_tiles) prefetch(tile + gridDim.x, cur ^ 1);
    mbar_wait(&s_bar[cur], (uint32_t)(it >> 1) & 1u);

    const TileCtx &c = s_ctx[cur];
    if (c.valid) {
      uint8_t *s_px = smem + cur * L::PX_BYTES;
      const int x0 = c.x0, y0 = c.y0, tw = c.tw, th = c.th, ya = c.ya, yb = c.yb;
      const int oy = y0 - HALO;
      const int H = c.H;
      const bool vec_out = c.vec_out != 0;
      const int n_src = c.n_src;
      uint8_t *const dy = c.dy;
      // composite tiles without text never touch shared memory for pixels
      const bool direct = (BPP == 4) && n_src > 1 && !c.stamp && c.vec_in && vec_out && (tw & 3) == 0;

      // ---- depth loads of a single-source tile (consumed after phase B) ------------------
      constexpr int DROWS = (TILE_H + NW - 1) / NW;
      uint2 dreg[DROWS];
      const bool depth_vec = c.depth_vec != 0;  // CTA-uniform
      const bool depth_lane = depth_vec && lane * 8 < tw;
      if (depth_lane) {
        const uint8_t *dsrc = c.dep[0] + lane * 8;
        const int dstride = c.ds[0];
#pragma unroll
        for (int i = 0; i < DROWS; i++) {
          const int r = warp + i * NW;
          if (r < th) dreg[i] = __ldg((const uint2 *)(dsrc + (y0 + r) * dstride));
        }
      }

      // ---- fill the tile ourselves when it was not staged by bulk copies ------------------
      if (!c.tma && !direct) {
        for (int y = ya + grab(qF, lane); y < yb; y = ya + grab(qF, lane)) {
          uint8_t *s = s_px + (y - oy) * L::ROWB;
          if (n_src == 1) {
            const uint8_t *g = c.rgb[0] + (size_t)y * c.rs[0];
            const int nbytes = tw * BPP;
            for (int i = lane; i < nbytes; i += 32) s[i] = g[i];
          } else {
            const bool core = (y >= y0) && (y < y0 + th);
            if (BPP == 4 && c.vec_in && vec_out && (tw & 3) == 0) {
              uint32_t p[8], d4[2];
              composite_fetch<0>(c, y, lane, n_src, p, d4);
              *(uint4 *)(s + lane * 16) = make_uint4(p[0], p[1], p[2], p[3]);
              *(uint4 *)(s + 512 + lane * 16) = make_uint4(p[4], p[5], p[6], p[7]);
              if (core && dy) {
                if (lane * 4 < tw) *(uint32_t *)(dy + (size_t)y * c.dys + lane * 4) = gray_y4_packed(d4[0]);
                if (128 + lane * 4 < tw) *(uint32_t *)(dy + (size_t)y * c.dys + 128 + lane * 4) = gray_y4_packed(d4[1]);
              }
            } else {
              const DevJob &jb = jobs[c.job];
              for (int x = lane; x < tw; x += 32) {
                uint32_t d;
                composite_px<BPP>(jb, x0 + x, y, s + x * BPP, &d);
                if (core && dy) dy[(size_t)y * c.dys + x] = (uint8_t)gray_y(d);
              }
            }
          }
        }
      }
      if ((!c.tma && !direct) || c.stamp) __syncthreads();

      // ---- text overlay, stamped into the shared tile ---------------------------------------
      if (c.stamp) stamp_glyphs<BPP>(jobs[c.job], s_px, L::ROWB, x0, oy, x0, x0 + tw, ya, yb, s_hits, s_nhits);

      // ---- phase A: per source row: Y out, pair-summed chroma (u15 | v15<<16) in place --------
      {
        const uint32_t kya = c.kya, kyb = c.kyb, kua = c.kua, kub = c.kub, kva = c.kva, kvb = c.kvb;
        uint8_t *const sy = c.sy;
        const size_t sys = (size_t)c.sys;
        const uint32_t px_addr = smem_base + cur * L::PX_BYTES;
        int y = ya + grab(qA, lane);
        while (y < yb) {
          const int y_next = ya + grab(qA, lane);  // overlaps the queue round trip with this row's arithmetic
          const uint32_t row = px_addr + (y - oy) * L::ROWB;
          const bool core = (y >= y0) && (y < y0 + th);
          uint32_t p[8];  // pixel words: colour bytes of pixel k in the positions ky/ku/kv expect
          if (BPP == 3) {
            // lane owns pixels 8*lane .. 8*lane+7 (24 bytes; 8-byte loads at 24-byte stride are conflict free)
            const uint2 a = lds64(row + lane * 24), b = lds64(row + lane * 24 + 8), d = lds64(row + lane * 24 + 16);
            p[0] = a.x;
            p[1] = __funnelshift_r(a.x, a.y, 24);
            p[2] = __funnelshift_r(a.y, b.x, 16);
            p[3] = __funnelshift_r(b.x, b.y, 8);
            p[4] = b.y;
            p[5] = __funnelshift_r(b.y, d.x, 24);
            p[6] = __funnelshift_r(d.x, d.y, 16);
            p[7] = d.y >> 8;
            __syncwarp();  // every lane has read its pixels before anyone overwrites the row
          } else if (direct) {
            uint32_t d4[2];
            if (n_src == 2) composite_fetch<2>(c, y, lane, 2, p, d4);
            else if (n_src == 4) composite_fetch<4>(c, y, lane, 4, p, d4);
            else composite_fetch<0>(c, y, lane, n_src, p, d4);
            if (core && dy) {
              if (lane * 4 < tw) *(uint32_t *)(dy + (size_t)y * c.dys + lane * 4) = gray_y4_packed(d4[0]);
              if (128 + lane * 4 < tw) *(uint32_t *)(dy + (size_t)y * c.dys + 128 + lane * 4) = gray_y4_packed(d4[1]);
            }
          } else {
            // lane owns pixels 4*lane..+3 and 128+4*lane..+3 (16-byte loads at 16-byte stride)
            const uint4 a = lds128(row + lane * 16), b = lds128(row + 512 + lane * 16);
            p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
            __syncwarp();
          }
          uint32_t uv[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            int su = dp2a_lo(kua, p[2 * j], C_BIAS); su = dp2a_hi(kub, p[2 * j], su);
            su = dp2a_lo(kua, p[2 * j + 1], su); su = dp2a_hi(kub, p[2 * j + 1], su);
            int sv = dp2a_lo(kva, p[2 * j], C_BIAS); sv = dp2a_hi(kvb, p[2 * j], sv);
            sv = dp2a_lo(kva, p[2 * j + 1], sv); sv = dp2a_hi(kvb, p[2 * j + 1], sv);
            uv[j] = (((uint32_t)su >> 9) & 0xFFFEu) | (((uint32_t)sv << 7) & 0xFFFE0000u);
          }
          if (BPP == 3) {
            sts128(row + lane * 16, uv[0], uv[1], uv[2], uv[3]);  // chroma cols 4*lane..+3
          } else {
            sts64(row + lane * 8, uv[0], uv[1]);         // chroma cols 2*lane, 2*lane+1
            sts64(row + 256 + lane * 8, uv[2], uv[3]);   // chroma cols 64+2*lane, +1
          }
          if (core) {
            uint32_t yv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) yv[k] = (uint32_t)dp2a_hi(kyb, p[k], dp2a_lo(kya, p[k], Y_BIAS)) >> 15;
            const uint32_t w0 = yv[0] | (yv[1] << 8) | (yv[2] << 16) | (yv[3] << 24);
            const uint32_t w1 = yv[4] | (yv[5] << 8) | (yv[6] << 16) | (yv[7] << 24);
            uint8_t *o = sy + y * (int)sys;
            if (BPP == 3) {
              const int x = lane * 8;
              if (vec_out && x + 8 <= tw) {
                *(uint2 *)(o + x) = make_uint2(w0, w1);
              } else {
#pragma unroll
                for (int k = 0; k < 8; k++)
                  if (x + k < tw) o[x + k] = (uint8_t)yv[k];
              }
            } else {
              const int xa = lane * 4, xb = 128 + lane * 4;
              if (vec_out && xa + 4 <= tw) *(uint32_t *)(o + xa) = w0;
              else
#pragma unroll
                for (int k = 0; k < 4; k++)
                  if (xa + k < tw) o[xa + k] = (uint8_t)yv[k];
              if (vec_out && xb + 4 <= tw) *(uint32_t *)(o + xb) = w1;
              else
#pragma unroll
                for (int k = 0; k < 4; k++)
                  if (xb + k < tw) o[xb + k] = (uint8_t)yv[4 + k];
            }
          }
          y = y_next;
        }
      }
      __syncthreads();
      if (tid == 0) { s_q[0] = 0; s_q[2] = 0; }  // next tile's queues (nobody is in them now)

      // ---- phase B: 8-tap vertical bicubic on chroma; edge taps fold = clamped row index ------
      // T = [-58,-172,492,1786,1786,492,-172,-58]/4096 is symmetric: pair the taps first (packed
      // u|v<<16 words add without carry: 2*32767 < 65536).
      {
        const int cc = lane * 4;  // chroma column inside the tile
        uint8_t *const su_ = c.su, *const sv_ = c.sv;
        const size_t sus = (size_t)c.sus, svs = (size_t)c.svs;
        const int crows = th >> 1;
        const uint32_t px_addr = smem_base + cur * L::PX_BYTES + cc * 4;
        const bool interior = (ya == oy) && (yb == y0 + th + HALO);  // no clamped tap rows in this tile
        uint8_t *const du = c.du, *const dv = c.dv;
        const int dus = c.dus, dvs = c.dvs;
        int r = grab(qB, lane);
        while (r < crows) {
          const int r_next = grab(qB, lane);
          const int ci = (y0 >> 1) + r;
          const int rr = r;
          r = r_next;
          if (cc >= (tw >> 1)) continue;
          uint32_t t[8][4];
          if (interior) {
            const uint32_t base = px_addr + (uint32_t)(2 * rr) * L::ROWB;  // tile row of tap 0 = 2*rr
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const uint4 q = lds128(base + j * L::ROWB);
              t[j][0] = q.x; t[j][1] = q.y; t[j][2] = q.z; t[j][3] = q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int sr = min(max(2 * ci - 3 + j, 0), H - 1) - oy;
              const uint4 q = lds128(px_addr + sr * L::ROWB);
              t[j][0] = q.x; t[j][1] = q.y; t[j][2] = q.z; t[j][3] = q.w;
            }
          }
          uint32_t ub = 0, vb = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint32_t a07 = t[0][k] + t[7][k], a16 = t[1][k] + t[6][k], a25 = t[2][k] + t[5][k], a34 = t[3][k] + t[4][k];
            const int au = (64 << 12) + 1786 * (int)(a34 & 0xFFFFu) + 492 * (int)(a25 & 0xFFFFu) - 172 * (int)(a16 & 0xFFFFu) - 58 * (int)(a07 & 0xFFFFu);
            const int av = (64 << 12) + 1786 * (int)(a34 >> 16) + 492 * (int)(a25 >> 16) - 172 * (int)(a16 >> 16) - 58 * (int)(a07 >> 16);
            ub |= clip8_relu(au >> 19) << (8 * k);
            vb |= clip8_relu(av >> 19) << (8 * k);
          }
          uint8_t *ou = su_ + ci * (int)sus + cc;
          uint8_t *ov = sv_ + ci * (int)svs + cc;
          if (vec_out && cc + 4 <= (tw >> 1)) {
            *(uint32_t *)ou = ub;
            *(uint32_t *)ov = vb;
            if (dy) {  // depth chroma planes are constant 128
              *(uint32_t *)(du + ci * dus + cc) = 0x80808080u;
              *(uint32_t *)(dv + ci * dvs + cc) = 0x80808080u;
            }
          } else {
            for (int k = 0; k < 4; k++)
              if (cc + k < (tw >> 1)) {
                ou[k] = (uint8_t)(ub >> (8 * k)); ov[k] = (uint8_t)(vb >> (8 * k));
                if (dy) { du[ci * dus + cc + k] = 128; dv[ci * dvs + cc + k] = 128; }
              }
          }
        }
      }

      // ---- depth stream: Y = range-compressed gray (U = V = 128 went out with phase B) ----------
      if (dy && n_src == 1) {
        const int dys = c.dys;
        if (depth_vec) {
          if (depth_lane) {
            uint8_t *o = dy + lane * 8;
#pragma unroll
            for (int i = 0; i < DROWS; i++) {
              const int r = warp + i * NW;
              if (r < th) *(uint2 *)(o + (y0 + r) * dys) = make_uint2(gray_y4_packed(dreg[i].x), gray_y4_packed(dreg[i].y));
            }
          }
        } else {
          const uint8_t *dsrc = c.dep[0];
          const int dstride = c.ds[0];
          for (int r = warp; r < th; r += NW)
            for (int x = lane; x < tw; x += 32) dy[(y0 + r) * dys + x] = (uint8_t)gray_y(dsrc[(size_t)(y0 + r) * dstride + x]);
        }
      }
    }
    // the buffer of this iteration is refilled by bulk copies issued at the top of the next
    // iteration: order our generic-proxy writes (in-place chroma, stamps) before them
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) s_q[1] = 0;
  }
}

static int g_ctas_per_sm[2] = {0, 0};
static int g_num_sms = 0;

int frame_tiles_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_frame_tiles<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileSmem<3>::TOTAL);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_frame_tiles<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileSmem<4>::TOTAL);
  if (e != cudaSuccess) return (int)e;
  int dev = 0;
  cudaGetDevice(&dev);
  e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_ctas_per_sm[0], k_frame_tiles<3>, CTA_THREADS, TileSmem<3>::TOTAL);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_ctas_per_sm[1], k_frame_tiles<4>, CTA_THREADS, TileSmem<4>::TOTAL);
  if (e != cudaSuccess) return (int)e;
  if (g_ctas_per_sm[0] < 1 || g_ctas_per_sm[1] < 1) return (int)cudaErrorLaunchOutOfResources;
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
    const int grid = total < g_num_sms * g_ctas_per_sm[0] ? total : g_num_sms * g_ctas_per_sm[0];
    k_frame_tiles<3><<<grid, CTA_THREADS, TileSmem<3>::TOTAL, (cudaStream_t)stream>>>(jobs_dev, n_jobs, total);
    launches++;
  }
  if (any4) {
    const int grid = total < g_num_sms * g_ctas_per_sm[1] ? total : g_num_sms * g_ctas_per_sm[1];
    k_frame_tiles<4><<<grid, CTA_THREADS, TileSmem<4>::TOTAL, (cudaStream_t)stream>>>(jobs_dev, n_jobs, total);
    launches++;
  }
  return launches;
}

}  // namespace nes
