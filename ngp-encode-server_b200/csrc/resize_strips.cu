// resize_strips.cu -- k_resize_strips: the bicubic resize path in the strip / ring organisation of
// k_frame_strips (any size change whose filters fit the limits below; the rest stays with k_resize_tiles).
//
// What it computes: libswscale's generic C path as the reference drives it
// (/root/reference/src/base/video/type_managers.cc:143-155 via rendered_frame.h:24-33, after the overlay of
// render_text.cc:81-110); integer spec in SURVEY.md Appendix A.3 / A.4, filter tables of csrc/filter.cc:
//   scene : [depth-select composite] -> glyph stamp -> 14-bit Y / (pair-summed) U,V -> horizontal polyphase
//           (>>13, 15 bit) -> vertical polyphase (>>19) -> 8-bit planes
//   depth : GRAY8 -> horizontal polyphase (>>7) -> range compression -> vertical; U = V = 128
//
// Shape (HBM-bound in principle; what bounds it in practice is instruction issue, so the organisation is about
// touching every source pixel once, keeping per-pixel instruction counts low and never making one warp wait for another):
//   * work unit = one SEGMENT (rz_seg_rows destination rows) of one column STRIP (rz_dw destination columns)
//     of one frame; ONE persistent CTA per SM takes units from a global counter.  The strip's source window is at most
//     RZ_BOXW = 128 pixels wide: one 4-pixel group per lane;
//   * the source window is walked top to bottom in CHUNKS of 32 source rows; rows travel through a ring of 8-row
//     SUB-STAGES filled by the PRODUCER warp with 2D tensor-map TMA copies (packed pixels and GRAY8 depth of every
//     staged source; SASS UTMALDG), exactly like k_frame_strips;
//   * 16 HORIZONTAL warps: a warp owns one source row of a sub-stage END TO END: composite select among the staged
//     sources in registers -> overlay bits (one shared-memory word) -> 14-bit Y / U / V and depth bytes
//     into a 1 KB per-warp row buffer -> horizontal pass of that row (lane = destination column, its taps live in
//     registers for the whole unit) -> 15-bit samples into four rings (Y, depth, U, V); the sub-stage is released at once;
//   * 6 VERTICAL warps run one chunk behind: every destination row whose taps are all in the rings (4 adjacent columns
//     per thread, 16-byte ring loads, 4-byte stores).  The vertical halo is paid once per segment, nothing is staged twice;
//   * the roles meet only at mbarriers (sub-stage full / empty, chunk horizontally done, chunk vertically done): there is
//     no CTA-wide barrier anywhere in the loop;
//   * the text overlay of a unit is ONE BIT PER PIXEL of its source window, built in shared memory by the horizontal
//     warps when the unit starts (one thread tests a placed glyph, the hits are staged as descriptors, a lane ORs one
//     word of one glyph row of the 1-bit atlas); the rows apply it in registers after the composite.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "device_common.cuh"
#include "nes_internal.h"
#include "strips_common.cuh"

namespace nes {

namespace {

enum { RM_ROWS = 0, RM_SELECT = 1 };

constexpr int RZ_ROWBUF = 1024;        // per-warp row buffer: y14 | u14 | v14 (u16, 144 each) | depth bytes (144)
constexpr int RZ_RB_Y = 0, RZ_RB_U = 288, RZ_RB_V = 576, RZ_RB_D = 864;
constexpr int RZ_NS_MAX = 8;
constexpr int RZ_NCTX = 8;             // chunk contexts in flight: the producer runs <= 2 chunks ahead of the horizontal warps, the vertical warps <= 2 behind
constexpr int RZ_HITS = 96;            // glyph descriptors staged per pass of a unit's overlay build
constexpr int RZ_NL = RZ_MAX_DW / 32;  // 32-column slots of a destination strip
constexpr int RZ_SUBS = RZ_CH / RZ_SUB;
constexpr int RZ_HG = RZ_H_WARPS / RZ_SUB;  // groups of horizontal warps (group g takes sub-stages g, g + RZ_HG, ... of a chunk)
static_assert(RZ_H_WARPS % RZ_SUB == 0 && RZ_SUBS % RZ_HG == 0, "a horizontal warp takes whole sub-stage rows");
constexpr int RZ_ROWS_PER_WARP = RZ_SUBS / RZ_HG;

// Everything the horizontal and vertical warps need to know about one chunk (written by lane 0 of the producer warp).
struct RzCtx {
  int32_t last;       // last chunk this CTA processes
  int32_t first;      // first chunk of a unit: (re)load the horizontal filter registers
  int32_t mode, n_src, n_staged, job;
  int32_t unit_text;  // (first chunk) some placed glyph may touch the unit's window: build its overlay bits
  int32_t stamp;      // the overlay bits of this chunk's rows are not all zero
  int32_t wx0;        // source column of window column 0 (4-byte pixels: multiple of 4; 3-byte pixels: of 16 -- TMA box rows start on 16-byte boundaries)
  int32_t dshift;     // window column 0 inside the staged depth rows (their box starts at wx0 & ~15)
  int32_t yc0, ra, rb;          // source row of chunk-local row 0; rows of this chunk that are needed [ra, rb)
  int32_t lr0, lr1, cr0, cr1;   // source rows the luma / chroma vertical filters of this unit read
  int32_t ya, yb, ca, cb;       // destination rows to emit after this chunk (luma, chroma)
  int32_t rbase_y, rbase_c;     // luma / chroma ring slot of chunk-local row 0 (the rings run on across units)
  int32_t dx0, dw, cx0, dcw;    // destination strip: luma / chroma columns
  int32_t half, dep_staged, nv12, rgb_base;
  int32_t hls, hcs, vls, vcs;   // filter sizes
  uint32_t a_mask;
  uint32_t ky[2], ku[2], kv[2];
  const int16_t *hl_coef, *hc_coef, *vl_coef, *vc_coef;
  const int32_t *hl_pos, *hc_pos, *vl_pos, *vc_pos;
  int32_t sys, sus, svs, dys, dus, dvs;
  uint8_t *sy, *su, *sv, *dy, *du, *dv;  // plane bases (not offset)
};

struct RzSmem {
  int ringY, ringD, ringU, ringV, rowbuf, ctx, bar, hits, ovl, stage;
};
constexpr int RZ_CTX_BYTES = ((int)sizeof(RzCtx) + 15) & ~15;
// nry / nrc: rows of the luma (+ depth) and chroma rings.  The vertical pass of chunk j reads back to RZ_CH + (filter
// size) - 1 rows before the end of chunk j while the horizontal warps may already be writing chunk j + 1 (they wait for
// the vertical pass of chunk j - 1 before that): 2 * RZ_CH - 1 + (filter size) rows.
__host__ __device__ inline int rz_ring_rows(int vsize) { return (2 * RZ_CH - 1 + vsize + 7) & ~7; }
// ovl_rows: source rows of the tallest unit of the launch that carries text (0: none does)
__host__ __device__ inline RzSmem rz_smem(int dwp, int nry, int nrc, int ovl_rows) {
  RzSmem L;
  int o = 0;
  L.ringY = o; o += nry * dwp * 4;
  L.ringD = o; o += nry * dwp * 4;
  L.ringU = o; o += nrc * (dwp / 2) * 4;
  L.ringV = o; o += nrc * (dwp / 2) * 4;
  L.rowbuf = o; o += RZ_H_WARPS * RZ_ROWBUF;
  L.ctx = o; o += RZ_NCTX * RZ_CTX_BYTES;
  L.bar = o; o += (2 * RZ_NS_MAX + 4) * 8;  // full[], empty[], hdone[2], vdone[2]
  L.hits = o; o += ovl_rows > 0 ? RZ_HITS * (int)sizeof(DevPlaced) + 16 : 0;
  L.ovl = o; o += ovl_rows * (RZ_BOXW / 8);  // one bit per pixel of the unit's source window
  L.stage = (o + 1023) & ~1023;
  return L;
}

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
__device__ __forceinline__ int4 lds_v4s(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
// coefficient . colour bytes of one pixel word (k0: bytes 0,1; k1: bytes 2,3; non-colour bytes carry 0)
__device__ __forceinline__ int dot_px(uint32_t k0, uint32_t k1, uint32_t px, int acc) { return dp2a_hi(k1, px, dp2a_lo(k0, px, acc)); }

__device__ __forceinline__ int job_of_rz_unit(const DevJob *jobs, int n_jobs, int cls, int u) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].rz_unit_base[cls] <= u) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Depth-select composite of the 4 pixels a lane owns (columns 4*lane..+3 of the window) from the N staged
// sources: the winner's pixel words and depth bytes (DESIGN.md "composite"; oracle nes_oracle_composite).
// px / dep: shared addresses of source 0's row; source k's row is k*RZ_SUB rows further.
template <int ROWB>
__device__ __forceinline__ void select4(uint32_t px, uint32_t dep, int n_src, uint32_t a_mask, int lane, uint32_t (&p)[4], uint32_t &d4) {
  uint32_t bd[4];
  {
    const uint4 q = lds128(px + lane * 16);
    const uint32_t dw = lds32(dep + lane * 4);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool valid = (w[i] & a_mask) != 0;
      bd[i] = valid ? __byte_perm(dw, 0u, 0x4440 + i) : 256u;
      p[i] = valid ? w[i] : 0u;
    }
  }
#pragma unroll
  for (int k = 1; k < TMA_MAX_SOURCES; k++) {
    if (k >= n_src) break;
    const uint4 q = lds128(px + (uint32_t)(k * RZ_SUB) * ROWB + lane * 16);
    // a renderer's partial view leaves most of a row transparent: a source without one valid pixel in this row of
    // the window is skipped by the whole warp
    if (!__any_sync(0xffffffffu, ((q.x | q.y | q.z | q.w) & a_mask) != 0)) continue;
    const uint32_t dw = lds32(dep + (uint32_t)(k * RZ_SUB) * RZ_DEPB + lane * 4);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t d = __byte_perm(dw, 0u, 0x4440 + i);
      const bool take = (w[i] & a_mask) != 0 && d < bd[i];
      bd[i] = take ? d : bd[i];
      p[i] = take ? w[i] : p[i];
    }
  }
  const uint32_t lo = __byte_perm(min(bd[0], 255u), min(bd[1], 255u), 0x0040), hi = __byte_perm(min(bd[2], 255u), min(bd[3], 255u), 0x0040);
  d4 = __byte_perm(lo, hi, 0x5410);
}

// Horizontal pass of one source row for NL slots of 32 luma columns (+ the depth plane, same filter): all chains
// advance together, tap by tap.
template <int NL, int T>
__device__ __forceinline__ void h_luma(uint32_t rowbuf, const int (&pos_l)[RZ_NL], const int (&cf_l)[RZ_NL][T], bool want_depth, int dw, int lane, uint32_t oy, uint32_t od) {
  int v[NL], d[NL];
  uint32_t a[NL], ad[NL];
#pragma unroll
  for (int cc = 0; cc < NL; cc++) { v[cc] = 0; d[cc] = 0; a[cc] = rowbuf + RZ_RB_Y + pos_l[cc] * 2; ad[cc] = rowbuf + RZ_RB_D + pos_l[cc]; }
  if (want_depth) {
#pragma unroll
    for (int jj = 0; jj < T; jj++) {
#pragma unroll
      for (int cc = 0; cc < NL; cc++) {
        const int yv = lds_u16(a[cc] + 2 * jj), dv = lds_u8(ad[cc] + jj);
        v[cc] += yv * cf_l[cc][jj];
        d[cc] += dv * cf_l[cc][jj];
      }
    }
  } else {
#pragma unroll
    for (int jj = 0; jj < T; jj++) {
#pragma unroll
      for (int cc = 0; cc < NL; cc++) v[cc] += lds_u16(a[cc] + 2 * jj) * cf_l[cc][jj];
    }
  }
#pragma unroll
  for (int cc = 0; cc < NL; cc++) {
    const bool on = lane + 32 * cc < dw;
    const int y15 = min(v[cc] >> 13, 32767);
    if (on) sts32(oy + 128 * cc, (uint32_t)y15);
    if (want_depth) {
      int g = min(d[cc] >> 7, 32767);
      g = (g * 14071 + 33561472) >> 14;
      if (on) sts32(od + 128 * cc, (uint32_t)g);
    }
  }
}
template <int NC, int T>
__device__ __forceinline__ void h_chroma(uint32_t rowbuf, const int (&pos_c)[2], const int (&cf_c)[2][T], int dcw, int lane, uint32_t ou, uint32_t ov) {
  int u[NC], v[NC];
  uint32_t au[NC], av[NC];
#pragma unroll
  for (int cc = 0; cc < NC; cc++) { u[cc] = 0; v[cc] = 0; au[cc] = rowbuf + RZ_RB_U + pos_c[cc] * 2; av[cc] = rowbuf + RZ_RB_V + pos_c[cc] * 2; }
#pragma unroll
  for (int jj = 0; jj < T; jj++) {
#pragma unroll
    for (int cc = 0; cc < NC; cc++) {
      const int uu = lds_u16(au[cc] + 2 * jj), vv = lds_u16(av[cc] + 2 * jj);
      u[cc] += uu * cf_c[cc][jj];
      v[cc] += vv * cf_c[cc][jj];
    }
  }
#pragma unroll
  for (int cc = 0; cc < NC; cc++) {
    if (lane + 32 * cc < dcw) {
      sts32(ou + 128 * cc, (uint32_t)min(u[cc] >> 13, 32767));
      sts32(ov + 128 * cc, (uint32_t)min(v[cc] >> 13, 32767));
    }
  }
}

// barrier among the horizontal warps only
__device__ __forceinline__ void h_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * RZ_H_WARPS) : "memory"); }

// Overlay bits of one unit's source window (all horizontal warps; called when the unit starts): s_ovl[r1 - r0][RZ_BOXW / 32],
// window column x of source row y is bit x & 31 of word (y - r0) * 4 + (x >> 5).  The glyph list is bucketed by row band on
// the host; one thread tests one glyph, the hits are staged as whole descriptors, then a warp takes a glyph and a lane one
// word of one of its rows (the atlas holds one bit per glyph pixel) and ORs it in, shifted to the window's column grid.
// (render_text.cc:94-106: a bitmap pixel with coverage != 0 turns the frame pixel white.)
__device__ __forceinline__ void unit_overlay(const DevJob &jb, uint32_t *s_ovl, int x0, int r0, int r1, DevPlaced *s_hits, int *s_nhits) {
  constexpr int NT = 32 * RZ_H_WARPS;
  constexpr int MW = RZ_BOXW / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x1 = x0 + RZ_BOXW;
  h_sync();  // every horizontal warp is done with the previous unit's bits
  for (int i = tid; i < (r1 - r0) * MW; i += NT) s_ovl[i] = 0;
  if (tid == 0) { s_nhits[0] = 0; s_nhits[1] = 0; }
  h_sync();
  int g_begin = 0, g_end = jb.n_glyphs;
  if (jb.glyph_band_shift >= 0) {
    const int b0 = max(r0 - jb.glyph_max_h, 0) >> jb.glyph_band_shift, b1 = (r1 - 1) >> jb.glyph_band_shift;
    g_begin = jb.glyph_band[b0]; g_end = jb.glyph_band[b1 + 1];
  }
  const DevPlaced *__restrict__ glyphs = jb.glyphs;
  const uint32_t *__restrict__ atlas = jb.atlas;
  auto or_rows = [&](const DevPlaced &pg, int first, int step) {
    const int q0 = max(0, r0 - pg.y), q1 = min(pg.h, r1 - pg.y);
    const int p0 = max(0, x0 - pg.x), p1 = min(pg.w, x1 - pg.x);
    const int bit_lo = pg.bit0 + p0, bit_hi = pg.bit0 + p1;
    const int w_lo = bit_lo >> 5, w_hi = (bit_hi - 1) >> 5, nw = w_hi - w_lo + 1;
    for (int i = first; i < (q1 - q0) * nw; i += step) {
      const int qq = i / nw, wi = w_lo + (i - qq * nw), q = q0 + qq;
      uint32_t m = __ldg(atlas + pg.mask_off + (uint32_t)(q * pg.wpr + wi));
      const int lo = max(bit_lo - 32 * wi, 0), hi = min(bit_hi - 32 * wi, 32);
      m &= (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (32 - hi));
      if (m == 0) continue;
      const int xb = pg.x - pg.bit0 + 32 * wi - x0;  // window column of bit 0 of this word (negative only for masked-off bits)
      const int wd = xb >> 5, sh = xb & 31;
      uint32_t *mrow = s_ovl + (pg.y + q - r0) * MW;
      const uint32_t lo_part = m << sh;
      if (lo_part && wd >= 0) atomicOr(&mrow[wd], lo_part);
      if (sh) {
        const uint32_t hi_part = m >> (32 - sh);
        if (hi_part) atomicOr(&mrow[wd + 1], hi_part);
      }
    }
  };
  int pass = 0;
  for (int base = g_begin; base < g_end; base += NT, pass++) {
    int *cnt = s_nhits + (pass & 1);
    if (base + tid < g_end) {
      const DevPlaced pg = glyphs[base + tid];
      if (pg.x < x1 && pg.x + pg.w > x0 && pg.y < r1 && pg.y + pg.h > r0) {
        const int i = atomicAdd(cnt, 1);
        if (i < RZ_HITS) s_hits[i] = pg;
        else or_rows(pg, 0, 1);  // more hits than staging slots (tiny glyphs): this thread does the whole glyph
      }
    }
    h_sync();
    const int nh = min(*cnt, RZ_HITS);
    if (tid == 0) s_nhits[(pass + 1) & 1] = 0;  // (the other counter: last read before the previous pass's second barrier)
    for (int h = warp; h < nh; h += RZ_H_WARPS) or_rows(s_hits[h], lane, 32);
    h_sync();
  }
}

}  // namespace

// T: horizontal taps held in registers (>= the longest horizontal filter of the launch; shorter filters are
// padded with zero coefficients).
template <int BPP, int T>
__global__ void __launch_bounds__(RZ_THREADS, 1)
k_resize_strips(const DevJob *__restrict__ jobs, int n_jobs, int total_units, uint32_t *__restrict__ counters, int ns, int slot_bytes, int dwp, int nry, int nrc,
                int ovl_rows) {
  asm volatile("griddepcontrol.launch_dependents;");  // see k_frame_strips: consecutive launches overlap
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int ROWB = RZ_BOXW * BPP;
  const RzSmem L = rz_smem(dwp, nry, nrc, ovl_rows);
  uint64_t *s_full = (uint64_t *)(smem + L.bar);
  uint64_t *s_empty = s_full + RZ_NS_MAX;
  uint64_t *s_hdone = s_empty + RZ_NS_MAX;  // [2]: the horizontal warps have written chunk j's rows into the rings (j & 1)
  uint64_t *s_vdone = s_hdone + 2;          // [2]: the vertical warps have emitted chunk j's destination rows
  DevPlaced *s_hits = (DevPlaced *)(smem + L.hits);
  int *s_nhits = (int *)(s_hits + RZ_HITS);  // [2]: alternate from pass to pass (the idle one is re-armed meanwhile)
  uint32_t *s_ovl = (uint32_t *)(smem + L.ovl);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int i = 0; i < RZ_NS_MAX; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], RZ_SUB); }
    // (every thread of a role arrives itself: the hand-over is then a plain per-thread happens-before, which is also what compute-sanitizer's racecheck can follow)
    for (int i = 0; i < 2; i++) { mbar_init(&s_hdone[i], 32 * RZ_H_WARPS); mbar_init(&s_vdone[i], 32 * RZ_V_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == RZ_H_WARPS + RZ_V_WARPS) {
    // =========================== producer warp ===========================================
    int cur_u = (int)blockIdx.x, next_u = 0;
    if (lane == 0) next_u = (int)gridDim.x + (int)atomicAdd(&counters[0], 1u);
    next_u = __shfl_sync(0xffffffffu, next_u, 0);
    int q = 0, par = 0, round0 = 1;
    int chunk_it = 0;
    int rbase_y = 0, rbase_c = 0;  // the rings run on from unit to unit
    while (cur_u < total_units) {
      const int j = job_of_rz_unit(jobs, n_jobs, BPP - 3, cur_u);
      const DevJob *jp = jobs + j;
      const int local = cur_u - jp->rz_unit_base[BPP - 3];
      const int strip = local % jp->rz_strips_x, seg = local / jp->rz_strips_x;
      const int Wd = jp->Wd, Hd = jp->Hd, S = jp->rz_seg_rows, n_src = jp->n_src;
      const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
      const int half = jp->half;
      const int dx0 = strip * jp->rz_dw, dw = min(jp->rz_dw, Wd - dx0);
      const int cx0 = dx0 >> 1, dcw = min((dx0 + dw + 1) >> 1, cdW) - cx0;
      const int dy0 = seg * S, dy1 = min(dy0 + S, Hd);
      const int cy0 = dy0 >> 1, cy1 = min((dy1 + 1) >> 1, cdH);
      const int32_t *vlp = jp->vl.pos, *vcp = jp->vc.pos;
      const int vls = jp->vl.size, vcs = jp->vc.size;
      // filter positions are non-decreasing: the unit's source window follows from its first / last samples
      const int cpos0 = jp->hc.pos[cx0];
      const int wx0 = min(jp->hl.pos[dx0], half ? 2 * cpos0 : cpos0) & (BPP == 4 ? ~3 : ~15);  // box rows start on 16-byte boundaries
      const int wxd = wx0 & ~15;
      const int lr0 = vlp[dy0], lr1 = vlp[dy1 - 1] + vls, cr0 = vcp[cy0], cr1 = vcp[cy1 - 1] + vcs;
      const int r0 = min(lr0, cr0), r1 = max(lr1, cr1);
      const int nchunks = (r1 - r0 + RZ_CH - 1) / RZ_CH;
      const int n_staged = n_src;
      const int dep_staged = (jp->dy != nullptr) || n_src > 1;
      const uint32_t tx_bytes = (uint32_t)(n_staged * RZ_SUB) * (uint32_t)(ROWB + (dep_staged ? RZ_DEPB : 0));
      const TMap *my_map = nullptr;
      uint32_t my_off = 0;
      int my_x = 0;
      if (lane < n_staged) {
        my_map = &jp->tmap_px[lane]; my_off = (uint32_t)(lane * RZ_SUB) * ROWB; my_x = (wx0 * BPP) >> 2;
      } else if (dep_staged && lane < 2 * n_staged) {
        my_map = &jp->tmap_dep[lane - n_staged];
        my_off = (uint32_t)(n_staged * RZ_SUB) * ROWB + (uint32_t)((lane - n_staged) * RZ_SUB) * RZ_DEPB;
        my_x = wxd >> 2;
      }
      const int strips256 = (jp->W + STRIP_W - 1) / STRIP_W;
      const int nbands = (jp->H + (1 << MASK_BAND_SHIFT) - 1) >> MASK_BAND_SHIFT;
      // text: does any placed glyph touch a (32-row band, 256-pixel strip) cell of this unit's source window?
      int unit_text = 0;
      if (jp->n_glyphs > 0) {
        unit_text = 1;
        if (jp->use_mask) {
          const int s0 = wx0 / STRIP_W, s1 = min((wx0 + RZ_BOXW - 1) / STRIP_W, strips256 - 1);
          unsigned any = 0;
          for (int band = (r0 >> MASK_BAND_SHIFT) + lane; band <= (r1 - 1) >> MASK_BAND_SHIFT && band < nbands; band += 32)
            for (int st = s0; st <= s1; st++) {
              const int bit = band * strips256 + st;
              any |= (jp->tile_mask[bit >> 5] >> (bit & 31)) & 1u;
            }
          unit_text = __any_sync(0xffffffffu, any != 0);
        }
        if (unit_text && r1 - r0 > ovl_rows) __trap();  // the host sized the overlay rows from the same filters: cannot happen
      }
      int ly = dy0, cy = cy0;  // emission cursors
      for (int k = 0; k < nchunks; k++, chunk_it++) {
        const int yc0 = r0 + k * RZ_CH;
        const int ra = yc0, rb = min(yc0 + RZ_CH, r1);
        const bool last_k = (k == nchunks - 1);
        // destination rows whose last tap row is now staged (positions are non-decreasing: count the leading hits)
        const int ya = ly, ca = cy;
        for (;;) {
          const int c = ly + lane;
          const bool ok = c < dy1 && (last_k || vlp[c] + vls <= rb);
          const unsigned b = __ballot_sync(0xffffffffu, ok);
          const int n = (b == 0xffffffffu) ? 32 : __ffs(~b) - 1;
          ly += n;
          if (n < 32) break;
        }
        for (;;) {
          const int c = cy + lane;
          const bool ok = c < cy1 && (last_k || vcp[c] + vcs <= rb);
          const unsigned b = __ballot_sync(0xffffffffu, ok);
          const int n = (b == 0xffffffffu) ? 32 : __ffs(~b) - 1;
          cy += n;
          if (n < 32) break;
        }
        if (lane == 0) {
          RzCtx &c = *(RzCtx *)(smem + L.ctx + (chunk_it & (RZ_NCTX - 1)) * RZ_CTX_BYTES);
          int stamp = 0;
          if (unit_text) {
            stamp = 1;
            if (jp->use_mask) {
              stamp = 0;
              const int s0 = wx0 / STRIP_W, s1 = min((wx0 + RZ_BOXW - 1) / STRIP_W, strips256 - 1);
              for (int band = ra >> MASK_BAND_SHIFT; band <= (rb - 1) >> MASK_BAND_SHIFT && band < nbands; band++)
                for (int st = s0; st <= s1; st++) {
                  const int bit = band * strips256 + st;
                  stamp |= (jp->tile_mask[bit >> 5] >> (bit & 31)) & 1u;
                }
            }
          }
          c.last = last_k && next_u >= total_units;
          c.first = (k == 0);
          c.mode = n_src > 1 ? RM_SELECT : RM_ROWS;
          c.n_src = n_src; c.n_staged = n_staged; c.job = j;
          c.unit_text = unit_text; c.stamp = stamp;
          c.wx0 = wx0; c.dshift = wx0 - wxd; c.yc0 = yc0; c.ra = ra; c.rb = rb;
          c.lr0 = lr0; c.lr1 = lr1; c.cr0 = cr0; c.cr1 = cr1;
          c.ya = ya; c.yb = ly; c.ca = ca; c.cb = cy;
          c.rbase_y = rbase_y; c.rbase_c = rbase_c;
          c.dx0 = dx0; c.dw = dw; c.cx0 = cx0; c.dcw = dcw;
          c.half = half; c.dep_staged = dep_staged; c.nv12 = jp->nv12; c.rgb_base = jp->rgb_base;
          c.hls = jp->hl.size; c.hcs = jp->hc.size; c.vls = vls; c.vcs = vcs;
          c.a_mask = jp->a_off >= 0 ? (0xFFu << (8 * jp->a_off)) : 0xFFFFFFFFu;
          c.ky[0] = jp->ky[0]; c.ky[1] = jp->ky[1]; c.ku[0] = jp->ku[0]; c.ku[1] = jp->ku[1]; c.kv[0] = jp->kv[0]; c.kv[1] = jp->kv[1];
          c.hl_coef = jp->hl.coef; c.hc_coef = jp->hc.coef; c.vl_coef = jp->vl.coef; c.vc_coef = jp->vc.coef;
          c.hl_pos = jp->hl.pos; c.hc_pos = jp->hc.pos; c.vl_pos = vlp; c.vc_pos = vcp;
          c.sys = jp->sys; c.sus = jp->sus; c.svs = jp->svs; c.dys = jp->dys; c.dus = jp->dus; c.dvs = jp->dvs;
          c.sy = jp->sy; c.su = jp->su; c.sv = jp->sv; c.dy = jp->dy; c.du = jp->du; c.dv = jp->dv;
        }
        __syncwarp();
#pragma unroll 1
        for (int sub = 0; sub < RZ_SUBS; sub++) {
          if (!round0) mbar_wait_hint_a(smem_u32(&s_empty[q]), (uint32_t)(par ^ 1), 2000u);
          const int ys = yc0 + sub * RZ_SUB;
          const bool wanted = ys < rb;
          if (wanted) {
            if (lane == 0) mbar_arrive_expect_tx(&s_full[q], tx_bytes);
            __syncwarp();
            if (my_map) tma_load_2d(smem_base + L.stage + (uint32_t)(q * slot_bytes) + my_off, my_map, my_x, ys, &s_full[q]);
          } else if (lane == 0) {
            mbar_arrive(&s_full[q]);
          }
          if (++q == ns) { q = 0; par ^= 1; round0 = 0; }
        }
        rbase_y += RZ_CH;
        if (rbase_y >= nry) rbase_y -= nry;
        rbase_c += RZ_CH;
        if (rbase_c >= nrc) rbase_c -= nrc;
      }
      cur_u = next_u;
      if (lane == 0 && cur_u < total_units) next_u = (int)gridDim.x + (int)atomicAdd(&counters[0], 1u);
      next_u = __shfl_sync(0xffffffffu, next_u, 0);
    }
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) { counters[0] = 0; counters[1] = 0; __threadfence(); }
    }
    return;
  }

  const uint32_t ringY = smem_base + L.ringY, ringD = smem_base + L.ringD, ringU = smem_base + L.ringU, ringV = smem_base + L.ringV;
  const int rowbY = dwp * 4, rowbC = (dwp / 2) * 4;
  const uint32_t hdone0 = smem_u32(s_hdone), vdone0 = smem_u32(s_vdone);

  if (warp >= RZ_H_WARPS) {
    // ============================= vertical warps ===========================================
    // chunk j: wait until every horizontal warp has written its rows of chunk j, emit the destination rows the
    // producer listed for it, then tell the horizontal warps (they wait for chunk j before writing chunk j + 2)
    const int vtid = tid - 32 * RZ_H_WARPS;
    constexpr int VT = 32 * RZ_V_WARPS;
    for (int j = 0;; j++) {
      mbar_wait_hint_a(hdone0 + (j & 1) * 8, (uint32_t)((j >> 1) & 1), 2000u);
      const RzCtx &c = *(const RzCtx *)(smem + L.ctx + (j & (RZ_NCTX - 1)) * RZ_CTX_BYTES);
      const int last = c.last;
      const int yc0 = c.yc0, dw = c.dw, dcw = c.dcw;
      const bool want_depth = c.dy != nullptr;
      const int rbase_y = c.rbase_y, rbase_c = c.rbase_c;
      // (the launch only takes jobs with 16-byte aligned destination planes: whole groups go out as words)
      const int ya = c.ya, yb = c.yb, vls = c.vls;
      const int gl = (dw + 3) >> 2;
      const int total_l = (yb - ya) * gl;
      const uint32_t rcp_l = gl > 1 ? 0xFFFFFFFFu / (uint32_t)gl + 1u : 0u;  // idx / gl == umulhi(idx, rcp) for every idx here
      const int ca = c.ca, cb = c.cb, vcs = c.vcs;
      const int gc = (dcw + 3) >> 2;
      const int total_c = (cb - ca) * gc;
      const uint32_t rcp_c = gc > 1 ? 0xFFFFFFFFu / (uint32_t)gc + 1u : 0u;
      const bool nv12 = c.nv12 != 0;
      uint8_t *const sy = c.sy + c.dx0, *const dyp = want_depth ? c.dy + c.dx0 : nullptr;
      for (int it = vtid; it < total_l + total_c; it += VT) {
        if (it < total_l) {
          // luma (+ depth luma, same filter)
          const int idx = it;
          const int ry = gl > 1 ? (int)__umulhi((uint32_t)idx, rcp_l) : idx, g = idx - ry * gl;
          const int dyy = ya + ry;
          const int pos = __ldg(c.vl_pos + dyy);
          int slot = rbase_y + pos - yc0;
          if (slot < 0) slot += nry;
          if (slot >= nry) slot -= nry;
          int a[4], d[4];
          if (vls == 1) {
            const int4 w = lds_v4s(ringY + slot * rowbY + g * 16);
            a[0] = (w.x + 64) >> 7; a[1] = (w.y + 64) >> 7; a[2] = (w.z + 64) >> 7; a[3] = (w.w + 64) >> 7;
            if (want_depth) {
              const int4 e = lds_v4s(ringD + slot * rowbY + g * 16);
              d[0] = (e.x + 64) >> 7; d[1] = (e.y + 64) >> 7; d[2] = (e.z + 64) >> 7; d[3] = (e.w + 64) >> 7;
            }
          } else {
            const int16_t *cf = c.vl_coef + (size_t)dyy * vls;
            a[0] = a[1] = a[2] = a[3] = 64 << 12;
            d[0] = d[1] = d[2] = d[3] = 64 << 12;
            int kn = (int)__ldg(cf);
            for (int jj = 0; jj < vls; jj++) {
              const int k = kn;
              if (jj + 1 < vls) kn = (int)__ldg(cf + jj + 1);  // the next tap's coefficient is in flight while this tap is applied
              const int4 w = lds_v4s(ringY + slot * rowbY + g * 16);
              a[0] += w.x * k; a[1] += w.y * k; a[2] += w.z * k; a[3] += w.w * k;
              if (want_depth) {
                const int4 e = lds_v4s(ringD + slot * rowbY + g * 16);
                d[0] += e.x * k; d[1] += e.y * k; d[2] += e.z * k; d[3] += e.w * k;
              }
              if (++slot == nry) slot = 0;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) { a[k] >>= 19; d[k] >>= 19; }
          }
          const uint32_t yw = clip8_relu(a[0]) | (clip8_relu(a[1]) << 8) | (clip8_relu(a[2]) << 16) | (clip8_relu(a[3]) << 24);
          uint8_t *o = sy + (size_t)dyy * c.sys + 4 * g;
          if (4 * g + 4 <= dw) stg32(o, yw);
          else
            for (int k = 0; 4 * g + k < dw; k++) o[k] = (uint8_t)(yw >> (8 * k));
          if (want_depth) {
            const uint32_t gw = clip8_relu(d[0]) | (clip8_relu(d[1]) << 8) | (clip8_relu(d[2]) << 16) | (clip8_relu(d[3]) << 24);
            uint8_t *od = dyp + (size_t)dyy * c.dys + 4 * g;
            if (4 * g + 4 <= dw) stg32(od, gw);
            else
              for (int k = 0; 4 * g + k < dw; k++) od[k] = (uint8_t)(gw >> (8 * k));
          }
        } else {
          // chroma (U and V share the filter); depth chroma is constant 128
          const int idx = it - total_l;
          const int ry = gc > 1 ? (int)__umulhi((uint32_t)idx, rcp_c) : idx, g = idx - ry * gc;
          const int cyy = ca + ry;
          const int pos = __ldg(c.vc_pos + cyy);
          int slot = rbase_c + pos - yc0;
          if (slot < 0) slot += nrc;
          if (slot >= nrc) slot -= nrc;
          int u[4], v[4];
          if (vcs == 1) {
            const int4 w = lds_v4s(ringU + slot * rowbC + g * 16), e = lds_v4s(ringV + slot * rowbC + g * 16);
            u[0] = (w.x + 64) >> 7; u[1] = (w.y + 64) >> 7; u[2] = (w.z + 64) >> 7; u[3] = (w.w + 64) >> 7;
            v[0] = (e.x + 64) >> 7; v[1] = (e.y + 64) >> 7; v[2] = (e.z + 64) >> 7; v[3] = (e.w + 64) >> 7;
          } else {
            const int16_t *cf = c.vc_coef + (size_t)cyy * vcs;
            u[0] = u[1] = u[2] = u[3] = 64 << 12;
            v[0] = v[1] = v[2] = v[3] = 64 << 12;
            int kn = (int)__ldg(cf);
            for (int jj = 0; jj < vcs; jj++) {
              const int k = kn;
              if (jj + 1 < vcs) kn = (int)__ldg(cf + jj + 1);
              const int4 w = lds_v4s(ringU + slot * rowbC + g * 16), e = lds_v4s(ringV + slot * rowbC + g * 16);
              u[0] += w.x * k; u[1] += w.y * k; u[2] += w.z * k; u[3] += w.w * k;
              v[0] += e.x * k; v[1] += e.y * k; v[2] += e.z * k; v[3] += e.w * k;
              if (++slot == nrc) slot = 0;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) { u[k] >>= 19; v[k] >>= 19; }
          }
          const uint32_t ub = clip8_relu(u[0]) | (clip8_relu(u[1]) << 8) | (clip8_relu(u[2]) << 16) | (clip8_relu(u[3]) << 24);
          const uint32_t vb = clip8_relu(v[0]) | (clip8_relu(v[1]) << 8) | (clip8_relu(v[2]) << 16) | (clip8_relu(v[3]) << 24);
          const bool full = 4 * g + 4 <= dcw;
          if (nv12) {
            // U0 V0 U1 V1 | U2 V2 U3 V3: chroma column x sits at byte 2x of the UV row
            const uint32_t w0 = __byte_perm(ub, vb, 0x5140), w1 = __byte_perm(ub, vb, 0x7362);
            uint8_t *o = c.su + (size_t)cyy * c.sus + 2 * c.cx0 + 8 * g;
            if (full) stg64(o, w0, w1);
            else
              for (int k = 0; 4 * g + (k >> 1) < dcw; k++) o[k] = (uint8_t)((k < 4 ? w0 : w1) >> (8 * (k & 3)));
            if (want_depth) {
              uint8_t *od = c.du + (size_t)cyy * c.dus + 2 * c.cx0 + 8 * g;
              if (full) stg64(od, 0x80808080u, 0x80808080u);
              else
                for (int k = 0; 4 * g + (k >> 1) < dcw; k++) od[k] = 128;
            }
          } else {
            uint8_t *ou = c.su + (size_t)cyy * c.sus + c.cx0 + 4 * g, *ov = c.sv + (size_t)cyy * c.svs + c.cx0 + 4 * g;
            if (full) { stg32(ou, ub); stg32(ov, vb); }
            else
              for (int k = 0; 4 * g + k < dcw; k++) { ou[k] = (uint8_t)(ub >> (8 * k)); ov[k] = (uint8_t)(vb >> (8 * k)); }
            if (want_depth) {
              uint8_t *du_ = c.du + (size_t)cyy * c.dus + c.cx0 + 4 * g, *dv_ = c.dv + (size_t)cyy * c.dvs + c.cx0 + 4 * g;
              if (full) { stg32(du_, 0x80808080u); stg32(dv_, 0x80808080u); }
              else
                for (int k = 0; 4 * g + k < dcw; k++) { du_[k] = 128; dv_[k] = 128; }
            }
          }
        }
      }
      mbar_arrive_a(vdone0 + (j & 1) * 8);  // this thread is done with the rings and the context of chunk j
      if (last) break;
    }
    return;
  }

  // ============================= horizontal warps ===========================================
  const uint32_t full0 = smem_base + L.bar, empty0 = full0 + RZ_NS_MAX * 8;
  const uint32_t rowbuf = smem_base + L.rowbuf + warp * RZ_ROWBUF;
  const int hgrp = warp / RZ_SUB, wrow = warp % RZ_SUB;  // this warp takes row wrow of the sub-stages hgrp, hgrp + RZ_HG, ... of every chunk
  // horizontal filters of the current unit: lane = destination column (+32 per slot)
  int pos_l[RZ_NL], pos_c[2];
  int cf_l[RZ_NL][T], cf_c[2][T];
#pragma unroll
  for (int c = 0; c < RZ_NL; c++) {
    pos_l[c] = 0;
#pragma unroll
    for (int jj = 0; jj < T; jj++) cf_l[c][jj] = 0;
  }
#pragma unroll
  for (int c = 0; c < 2; c++) {
    pos_c[c] = 0;
#pragma unroll
    for (int jj = 0; jj < T; jj++) cf_c[c][jj] = 0;
  }
  int qc = 0, parc = 0;  // ring slot and parity of the chunk's first sub-stage
  for (int j = 0;; j++) {
    int q[RZ_ROWS_PER_WARP], par[RZ_ROWS_PER_WARP];
#pragma unroll
    for (int i = 0; i < RZ_ROWS_PER_WARP; i++) {
      int a = qc + hgrp + i * RZ_HG, p = parc;
      while (a >= ns) { a -= ns; p ^= 1; }
      q[i] = a; par[i] = p;
    }
    mbar_wait_hint_a(full0 + q[0] * 8, (uint32_t)par[0], 2000u);
    const RzCtx &c = *(const RzCtx *)(smem + L.ctx + (j & (RZ_NCTX - 1)) * RZ_CTX_BYTES);
    const int last = c.last;
    {
      const uint32_t dep_off = (uint32_t)(c.n_staged * RZ_SUB) * ROWB + (uint32_t)c.dshift;
      const int wx0 = c.wx0, yc0 = c.yc0, ra = c.ra, rb = c.rb;
      const int dw = c.dw, dcw = c.dcw;
      const bool half = c.half != 0, dep_staged = c.dep_staged != 0;
      const bool want_depth = c.dy != nullptr;
      const int mode = c.mode;

      if (c.first) {
        const int hls = c.hls, hcs = c.hcs;
        const int corg = half ? (wx0 >> 1) : wx0;
#pragma unroll
        for (int cc = 0; cc < RZ_NL; cc++) {
          const int col = lane + 32 * cc;
          const bool on = col < dw;
          pos_l[cc] = on ? __ldg(c.hl_pos + c.dx0 + col) - wx0 : 0;
          const int16_t *cf = c.hl_coef + (size_t)(c.dx0 + (on ? col : 0)) * hls;
#pragma unroll
          for (int jj = 0; jj < T; jj++) cf_l[cc][jj] = (on && jj < hls) ? (int)__ldg(cf + jj) : 0;
        }
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
          const int col = lane + 32 * cc;
          const bool on = col < dcw;
          pos_c[cc] = on ? __ldg(c.hc_pos + c.cx0 + col) - corg : 0;
          const int16_t *cf = c.hc_coef + (size_t)(c.cx0 + (on ? col : 0)) * hcs;
#pragma unroll
          for (int jj = 0; jj < T; jj++) cf_c[cc][jj] = (on && jj < hcs) ? (int)__ldg(cf + jj) : 0;
        }
      }
      // ---- text overlay of the unit: one bit per pixel of its source window (nothing is written to the staged rows)
      if (c.first && c.unit_text) unit_overlay(jobs[c.job], s_ovl, wx0, min(c.lr0, c.cr0), max(c.lr1, c.cr1), s_hits, s_nhits);
      // the ring rows this chunk overwrites were last read by the vertical pass of chunk j - 2
      if (j >= 2) mbar_wait_hint_a(vdone0 + (j & 1) * 8, (uint32_t)(((j - 2) >> 1) & 1), 2000u);

      // ---- per source row: composite -> overlay bits -> 14-bit planes (row buffer) -> horizontal pass -> rings ----------
      const uint32_t ky0 = c.ky[0], ky1 = c.ky[1], ku0 = c.ku[0], ku1 = c.ku[1], kv0 = c.kv[0], kv1 = c.kv[1];
      const int rbase_y = c.rbase_y, rbase_c = c.rbase_c;
      const bool stamped = c.stamp != 0;
      const int ovl_r0 = min(c.lr0, c.cr0);
#pragma unroll
      for (int i = 0; i < RZ_ROWS_PER_WARP; i++) {
        const int r = (hgrp + i * RZ_HG) * RZ_SUB + wrow;
        const int y = yc0 + r;
        if (i > 0) mbar_wait_hint_a(full0 + q[i] * 8, (uint32_t)par[i], 2000u);
        if (y >= ra && y < rb) {
          const uint32_t sb = smem_base + L.stage + (uint32_t)(q[i] * slot_bytes);
          const uint32_t row = sb + wrow * ROWB;
          const uint32_t drow = sb + dep_off + wrow * RZ_DEPB;
          // text overlay (render_text.cc:94-106: coverage != 0 -> white): the 4 bits of this lane's pixels
          uint32_t obits = 0;
          if (stamped) obits = s_ovl[(y - ovl_r0) * (RZ_BOXW / 32) + (lane >> 3)] >> ((lane & 7) * 4);
          uint32_t p[4], d4 = 0;
          if (BPP == 4) {
            if (mode == RM_SELECT) {
              select4<ROWB>(row, drow, c.n_src, c.a_mask, lane, p, d4);
            } else {
              const uint4 a = lds128(row + lane * 16);
              p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w;
              if (dep_staged) d4 = lds32(drow + lane * 4);
            }
          } else {
            // 4 packed 3-byte pixels = 3 words -> pixel words r | g<<8 | b<<16 | junk<<24 (the junk byte has coefficient 0)
            const uint32_t w0 = lds32(row + lane * 12), w1 = lds32(row + lane * 12 + 4), w2 = lds32(row + lane * 12 + 8);
            p[0] = w0;
            p[1] = __funnelshift_r(w0, w1, 24);
            p[2] = __funnelshift_r(w1, w2, 16);
            p[3] = w2 >> 8;
            if (dep_staged) d4 = lds32(drow + lane * 4);
          }
          if (stamped) {
            const uint32_t white = BPP == 4 ? (0x00FFFFFFu << (8 * c.rgb_base)) : 0x00FFFFFFu;
#pragma unroll
            for (int k = 0; k < 4; k++)
              if ((obits >> k) & 1u) p[k] |= white;
          }
          const bool need_l = y >= c.lr0 && y < c.lr1, need_c = y >= c.cr0 && y < c.cr1;
          if (need_l) {
            int yv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) yv[k] = dot_px(ky0, ky1, p[k], (32 << 14) + (1 << 8)) >> 9;
            sts64(rowbuf + RZ_RB_Y + lane * 8, pack16(yv[0], yv[1]), pack16(yv[2], yv[3]));
            if (want_depth) sts32(rowbuf + RZ_RB_D + lane * 4, d4);
          }
          if (need_c) {
            if (half) {
              int u[2], v[2];
#pragma unroll
              for (int k = 0; k < 2; k++) {
                u[k] = dot_px(ku0, ku1, p[2 * k + 1], dot_px(ku0, ku1, p[2 * k], C_BIAS)) >> 10;
                v[k] = dot_px(kv0, kv1, p[2 * k + 1], dot_px(kv0, kv1, p[2 * k], C_BIAS)) >> 10;
              }
              sts32(rowbuf + RZ_RB_U + lane * 4, pack16(u[0], u[1]));
              sts32(rowbuf + RZ_RB_V + lane * 4, pack16(v[0], v[1]));
            } else {
              int u[4], v[4];
#pragma unroll
              for (int k = 0; k < 4; k++) {
                u[k] = dot_px(ku0, ku1, p[k], C1_BIAS) >> 9;
                v[k] = dot_px(kv0, kv1, p[k], C1_BIAS) >> 9;
              }
              sts64(rowbuf + RZ_RB_U + lane * 8, pack16(u[0], u[1]), pack16(u[2], u[3]));
              sts64(rowbuf + RZ_RB_V + lane * 8, pack16(v[0], v[1]), pack16(v[2], v[3]));
            }
          }
          __syncwarp();
          int slot = rbase_y + r, slot_c = rbase_c + r;
          if (slot >= nry) slot -= nry;
          if (slot_c >= nrc) slot_c -= nrc;
          // ---- horizontal pass of this row: hScale16To15 (>>13, clamp) / hScale8To15 (>>7) + range compression.
          // The slots of a plane (32 destination columns each) and the luma / depth pair run as independent
          // accumulation chains inside one tap loop (the chains hide each other's multiply latency).
          if (need_l) {
            const uint32_t oy = ringY + slot * rowbY + lane * 4, od = ringD + slot * rowbY + lane * 4;
            const int nl = (dw + 31) >> 5;  // warp-uniform
            if (nl == 1) h_luma<1, T>(rowbuf, pos_l, cf_l, want_depth, dw, lane, oy, od);
            else if (nl == 2) h_luma<2, T>(rowbuf, pos_l, cf_l, want_depth, dw, lane, oy, od);
            else h_luma<3, T>(rowbuf, pos_l, cf_l, want_depth, dw, lane, oy, od);
          }
          if (need_c) {
            const uint32_t ou = ringU + slot_c * rowbC + lane * 4, ov = ringV + slot_c * rowbC + lane * 4;
            if (dcw <= 32) h_chroma<1, T>(rowbuf, pos_c, cf_c, dcw, lane, ou, ov);
            else h_chroma<2, T>(rowbuf, pos_c, cf_c, dcw, lane, ou, ov);
          }
        }
        __syncwarp();  // the row buffer is rewritten by this warp's next row; every lane is done with the sub-stage
        if (lane == 0) mbar_arrive_a(empty0 + q[i] * 8);
      }
    }
    mbar_arrive_a(hdone0 + (j & 1) * 8);  // this thread's ring samples of chunk j are written
    if (last) break;
    qc += RZ_SUBS;
    while (qc >= ns) { qc -= ns; parc ^= 1; }
  }
}

// ---------------------------------------------------------------------------
// host side: planning and launch
// ---------------------------------------------------------------------------
static int g_rz_sms = 0, g_rz_optin = 0;
static bool g_rz_pdl = true;

template <int BPP, int T>
static cudaError_t rz_set_attr() {
  return cudaFuncSetAttribute(k_resize_strips<BPP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_rz_optin);
}

int resize_strips_init() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e = cudaDeviceGetAttribute(&g_rz_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&g_rz_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return (int)e;
  if ((e = rz_set_attr<3, 4>()) != cudaSuccess || (e = rz_set_attr<3, 6>()) != cudaSuccess || (e = rz_set_attr<3, 8>()) != cudaSuccess ||
      (e = rz_set_attr<4, 4>()) != cudaSuccess || (e = rz_set_attr<4, 6>()) != cudaSuccess || (e = rz_set_attr<4, 8>()) != cudaSuccess)
    return (int)e;
  if (const char *v = getenv("NES_NO_PDL")) g_rz_pdl = atoi(v) == 0;
  return 0;
}

struct RzConfig {
  int ns, slot, smem, dwp, taps, nry, nrc, ovl_rows;
};
// Launch shape of one pixel class: rings sized for the widest destination strip and the longest vertical filters of
// the launch, sub-stage slots for the most sources; one CTA per SM, the rest of its shared memory is the sub-stage ring.
static RzConfig rz_config(const DevJob *jobs, int n_jobs, int bpp) {
  int dwp = 16, staged = 1, taps = 1, vl = 1, vc = 1;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs[j];
    if (!jb.rz_ok || jb.bpp != bpp) continue;
    dwp = std::max(dwp, (jb.rz_dw + 15) & ~15);
    staged = std::max(staged, jb.n_src);
    taps = std::max(taps, std::max(jb.hl.size, jb.hc.size));
    vl = std::max(vl, jb.vl.size);
    vc = std::max(vc, jb.vc.size);
  }
  // overlay bits: one row of RZ_BOXW bits per source row of the tallest unit that carries text.  A unit of S destination
  // rows reads at most S * ratio source rows plus the reach of the vertical filters (luma and chroma windows overlap)
  int ovl_rows = 0;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs[j];
    if (!jb.rz_ok || jb.bpp != bpp || jb.n_glyphs <= 0) continue;
    const double ratio = (double)jb.H / jb.Hd;
    const int rows = (int)std::ceil((jb.rz_seg_rows + 2) * ratio) + 2 * std::max(jb.vl.size, jb.vc.size) + (int)std::ceil(2 * ratio) + 8;
    ovl_rows = std::max(ovl_rows, std::min(rows, jb.H));
  }
  const int nry = rz_ring_rows(vl), nrc = rz_ring_rows(vc);
  const RzSmem L = rz_smem(dwp, nry, nrc, ovl_rows);
  const int slot = staged * RZ_SUB * (RZ_BOXW * bpp + RZ_DEPB);
  RzConfig c{0, slot, 0, dwp, taps <= 4 ? 4 : taps <= 6 ? 6 : 8, nry, nrc, ovl_rows};
  const int budget = (g_rz_optin > 0 ? g_rz_optin : 232448) - L.stage;
  // NES_RZ_NS (tests): cap the sub-stage ring, e.g. 2 or 3 slots for a 4-sub-stage chunk (what a launch with little shared memory left gets)
  static const int ns_cap = [] { const char *v = getenv("NES_RZ_NS"); const int x = v ? atoi(v) : RZ_NS_MAX; return std::min(std::max(x, 2), RZ_NS_MAX); }();
  const int ns = std::min(ns_cap, budget / slot);
  if (ns >= 2) { c.ns = ns; c.smem = L.stage + ns * slot; }
  return c;
}

// One segment height (destination rows) per pixel class: the vertical halo (about the longest vertical filter, in
// source rows, per segment) against the one-unit tail of the persistent grid.  Unit numbering: a prefix sum over the
// launch in which the jobs of the other class (and the jobs k_resize_tiles keeps) take no units.
void plan_resize_strips(DevJob *jobs, int n_jobs) {
  const int sms = g_rz_sms > 0 ? g_rz_sms : 148;
  for (int cls = 0; cls < 2; cls++) {
    const int bpp = 3 + cls;
    bool any = false;
    for (int j = 0; j < n_jobs; j++) any = any || (jobs[j].rz_ok && jobs[j].bpp == bpp);
    int best_s = 32;
    if (any) {
      // Candidates: the heights that cut a frame into k equal segments (multiples of 16).  Units of a launch cost about the same
      // and are handed out dynamically, so the launch takes ceil(units / CTAs) rounds of one unit: source rows of the segment +
      // the vertical halo and the ragged last chunk + what a unit costs before its first row (filter registers, overlay bits,
      // pipeline fill), in source rows.
      const int grid = sms;
      int hd_max = 0;
      for (int j = 0; j < n_jobs; j++)
        if (jobs[j].rz_ok && jobs[j].bpp == bpp) hd_max = std::max(hd_max, jobs[j].Hd);
      double best_cost = 1e30;
      for (int k = 1; k <= std::max(1, hd_max / 32); k++) {  // (segments under 32 rows never pay: a unit costs ~50 rows before its first one)
        const int S = std::max(16, ((hd_max + k - 1) / k + 15) & ~15);
        if (k > 1 && S == std::max(16, ((hd_max + k - 2) / (k - 1) + 15) & ~15)) continue;  // same height as the previous k
        double unit_max = 0;
        long units = 0;
        for (int j = 0; j < n_jobs; j++) {
          const DevJob &jb = jobs[j];
          if (!jb.rz_ok || jb.bpp != bpp) continue;
          const int strips = (jb.Wd + jb.rz_dw - 1) / jb.rz_dw, segs = (jb.Hd + S - 1) / S;
          const double ratio = (double)jb.H / jb.Hd;
          const double halo = std::max(jb.vl.size, jb.vc.size) + RZ_CH / 2;
          units += (long)strips * segs;
          unit_max = std::max(unit_max, std::min(S, jb.Hd) * ratio + halo + 24.0);
        }
        const double cost = (double)((units + grid - 1) / grid) * unit_max;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_s = S; }
      }
    }
    static const int seg_override = [] { const char *v = getenv("NES_RZ_SEG"); return v ? atoi(v) : 0; }();  // experiments: fixed segment height
    if (seg_override >= 16) best_s = seg_override;
    int base = 0;
    for (int j = 0; j < n_jobs; j++) {
      DevJob &jb = jobs[j];
      jb.rz_unit_base[cls] = base;
      if (!jb.rz_ok || jb.bpp != bpp) continue;
      jb.rz_seg_rows = best_s;
      jb.rz_strips_x = (jb.Wd + jb.rz_dw - 1) / jb.rz_dw;
      jb.rz_segs_y = (jb.Hd + best_s - 1) / best_s;
      jb.rz_units = jb.rz_strips_x * jb.rz_segs_y;
      base += jb.rz_units;
    }
  }
}

template <int BPP, int T>
static cudaError_t rz_launch_one(int grid, const RzConfig &c, cudaStream_t st, const DevJob *jobs, int n_jobs, int total, uint32_t *ctr) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(RZ_THREADS);
  cfg.dynamicSmemBytes = (size_t)c.smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_rz_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_resize_strips<BPP, T>, jobs, n_jobs, total, ctr, c.ns, c.slot, c.dwp, c.nry, c.nrc, c.ovl_rows);
}

int launch_resize_strips(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, uint32_t *counters, uint64_t *seq, void *stream) {
  int launches = 0;
  cudaStream_t st = (cudaStream_t)stream;
  for (int bpp = 3; bpp <= 4; bpp++) {
    int total = 0;
    for (int j = 0; j < n_jobs; j++)
      if (jobs_host[j].rz_ok && jobs_host[j].bpp == bpp) total += jobs_host[j].rz_units;
    if (total == 0) continue;
    const RzConfig c = rz_config(jobs_host, n_jobs, bpp);
    if (c.ns < 2) return -1;
    const int grid = std::min(total, g_rz_sms);
    uint32_t *ctr = counters + 2 * ((*seq)++ % COUNTER_SLOTS);
    cudaError_t e;
    if (bpp == 3) e = c.taps == 4 ? rz_launch_one<3, 4>(grid, c, st, jobs_dev, n_jobs, total, ctr) : c.taps == 6 ? rz_launch_one<3, 6>(grid, c, st, jobs_dev, n_jobs, total, ctr) : rz_launch_one<3, 8>(grid, c, st, jobs_dev, n_jobs, total, ctr);
    else e = c.taps == 4 ? rz_launch_one<4, 4>(grid, c, st, jobs_dev, n_jobs, total, ctr) : c.taps == 6 ? rz_launch_one<4, 6>(grid, c, st, jobs_dev, n_jobs, total, ctr) : rz_launch_one<4, 8>(grid, c, st, jobs_dev, n_jobs, total, ctr);
    if (e != cudaSuccess) return -1;
    launches++;
  }
  return launches;
}

}  // namespace nes
