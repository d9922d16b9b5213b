// strips_common.cuh -- device helpers shared by the strip-organised kernels (frame_strips.cu, resize_strips.cu):
// mbarrier / TMA wrappers, explicit shared-space accesses, dp2a flavours, packing helpers.
#ifndef NES_STRIPS_COMMON_CUH_
#define NES_STRIPS_COMMON_CUH_
#include <cuda_runtime.h>
#include <stdint.h>

#include "nes_internal.h"

namespace nes {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same on a precomputed 32-bit shared address (the consumer loop keeps the barrier base in a register)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// the same with a suspend-time hint (SASS: NANOSLEEP.SYNCS between tries -- the warp is woken by the barrier's update)
__device__ __forceinline__ void mbar_wait_hint_a(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(hint_ns) : "memory");
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
// global -> shared 2D tensor-map copy (TMA): box = 8 rows x one strip of u32 elements at element
// coordinates (x, y); rows / columns outside the tensor are zero-filled and still counted
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const TMap *map, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
// explicit shared-space accesses (32-bit shared addresses; keeps the hot loops off generic LD/ST)
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
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
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// barrier among the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * CONSUMER_WARPS) : "memory"); }

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
// the same with unsigned 16-bit halves (luma coefficients doubled: 2*16519 > 32767)
__device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t dp2a_hi_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// a: unsigned 16-bit halves (packed chroma pair sums), b: SIGNED bytes (small filter taps)
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// max(min(v, 255), 0) in one instruction
__device__ __forceinline__ uint32_t clip8_relu(int v) {
  int d;
  asm("min.s32.relu %0, %1, %2;" : "=r"(d) : "r"(v), "r"(255));
  return (uint32_t)d;
}
// bytes 2 of four sums -> one word (the doubled luma sums carry Y in byte 2, byte 3 is 0)
__device__ __forceinline__ uint32_t pack_b2(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}

// GRAY8 -> limited-range luma for the 4 bytes of a word:
//   Y = (d*219 + 127)/255 + 16 == (d*56282 + 1081500) >> 16 for every d in 0..255 (exhaustive
//   check in tests/test_host.py); one dp2a per pixel straight from the packed word, Y is byte 2
__device__ __forceinline__ uint32_t gray_y4_packed(uint32_t w) {
  constexpr uint32_t A = 56282u, B = 1081500u;
  return pack_b2(dp2a_lo_uu(A, w, B), dp2a_lo_uu(A << 16, w, B), dp2a_hi_uu(A, w, B), dp2a_hi_uu(A << 16, w, B));
}

// pair-summed chroma of a horizontal pixel pair -> packed 14-bit (u | v<<16); su, sv include C_BIAS
__device__ __forceinline__ uint32_t pack_uv14(int su, int sv) { return ((uint32_t)su >> 10) | (((uint32_t)sv << 6) & 0xFFFF0000u); }

// store 4 (or 8) output bytes at column x of a row that holds tw valid columns
__device__ __forceinline__ void store4(uint8_t *row, int x, uint32_t w, int tw, bool vec) {
  if (vec && x + 4 <= tw) *(uint32_t *)(row + x) = w;
  else
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (x + k < tw) row[x + k] = (uint8_t)(w >> (8 * k));
}
__device__ __forceinline__ void store8(uint8_t *row, int x, uint32_t w0, uint32_t w1, int tw, bool vec) {
  if (vec && x + 8 <= tw) *(uint2 *)(row + x) = make_uint2(w0, w1);
  else { store4(row, x, w0, tw, false); store4(row, x + 4, w1, tw, false); }
}
__device__ __forceinline__ void stg32(uint8_t *p, uint32_t w) { *(uint32_t *)p = w; }
__device__ __forceinline__ void stg64(uint8_t *p, uint32_t w0, uint32_t w1) { *(uint2 *)p = make_uint2(w0, w1); }

// Glyph stamp into the staged rows of source 0 (consumer warps only).  Reference semantics
// (render_text.cc:94-106): every bitmap pixel with coverage != 0 inside the frame becomes
// (255,255,255).  All stamps write the same value: overlapping glyphs are order-free.
// Chunk-local row r lives in the (r / 8)-th sub-stage of the chunk, row r % 8.
// The glyph list is bucketed by row band on the host, so only the glyphs near the chunk's rows are tested
// (one thread each); the descriptors that hit are staged in shared memory; then a warp takes a glyph and a
// lane one of its rows: ONE load of the row's bit mask (1 bit per pixel, DevPlaced) and a loop over its set bits.
template <int BPP, int SW, int HITS = STRIP_HITS>
__device__ __forceinline__ void stamp_chunk(const DevJob &jb, const DevPlaced *glyph_list, uint8_t *stage0, int qc, int ns, int slot_bytes, int x0, int x1, int yc0, int ra,
                                            int rb, int rgb_base, DevPlaced *s_hits, int *s_nhits) {
  constexpr int ROWB = SW * BPP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int g_begin = 0, g_end = jb.n_glyphs;
  if (jb.glyph_band_shift >= 0) {
    const int b0 = max(ra - jb.glyph_max_h, 0) >> jb.glyph_band_shift, b1 = (rb - 1) >> jb.glyph_band_shift;
    g_begin = jb.glyph_band[b0]; g_end = jb.glyph_band[b1 + 1];
  }
  const DevPlaced *__restrict__ glyphs = glyph_list ? glyph_list : jb.glyphs;  // (the list may travel in the kernel's parameter block)
  const uint32_t *__restrict__ atlas = jb.atlas;
  for (int base = g_begin; base < g_end; base += HITS) {
    if (tid == 0) *s_nhits = 0;
    consumer_sync();
    if (tid < HITS && base + tid < g_end) {
      const DevPlaced pg = glyphs[base + tid];
      if (pg.x < x1 && pg.x + pg.w > x0 && pg.y < rb && pg.y + pg.h > ra) s_hits[atomicAdd(s_nhits, 1)] = pg;
    }
    consumer_sync();
    const int nh = *s_nhits;
    for (int h = warp; h < nh; h += CONSUMER_WARPS) {
      const DevPlaced pg = s_hits[h];
      const int q0 = max(0, ra - pg.y), q1 = min(pg.h, rb - pg.y);   // visible rows inside the chunk
      const int p0 = max(0, x0 - pg.x), p1 = min(pg.w, x1 - pg.x);   // visible columns inside the strip
      const int bit_lo = pg.bit0 + p0, bit_hi = pg.bit0 + p1;        // bit range of a row's mask
      const int w_lo = bit_lo >> 5, w_hi = (bit_hi - 1) >> 5;
      const int nw = w_hi - w_lo + 1;
      for (int i = lane; i < (q1 - q0) * nw; i += 32) {
        const int qq = i / nw, wi = w_lo + (i - qq * nw), q = q0 + qq;
        uint32_t m = __ldg(atlas + pg.mask_off + (uint32_t)(q * pg.wpr + wi));
        // keep bits [bit_lo, bit_hi) of this word
        const int lo = max(bit_lo - 32 * wi, 0), hi = min(bit_hi - 32 * wi, 32);
        m &= (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (32 - hi));
        if (m == 0) continue;
        const int r = pg.y + q - yc0;
        int slot = qc + (r >> 3);  // chunk-local row r lives in sub-stage r / 8 after the chunk's first one (the ring may wrap)
        if (slot >= ns) slot -= ns;
        uint8_t *row = stage0 + slot * slot_bytes + (r & (SUB_ROWS - 1)) * ROWB + (BPP == 4 ? rgb_base : 0);
        const int xbase = pg.x - pg.bit0 + 32 * wi - x0;  // strip-local column of bit 0 of this word
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          uint8_t *px = row + (xbase + b) * BPP;
          px[0] = 255; px[1] = 255; px[2] = 255;
        }
      }
    }
    consumer_sync();
  }
}


}  // namespace
}  // namespace nes
#endif
