// frame_strips.cu -- k_frame_strips: the fused same-size frame kernel (the hot kernel).
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
//   * work unit = one SEGMENT (seg_rows output rows) of one 256-pixel column STRIP of one
//     frame; persistent CTAs fetch units from a global counter;
//   * a CTA walks its segment top to bottom in CHUNKS of 16 source rows (32 for 3-byte pixels);
//   * source rows travel through a ring of 8-row SUB-STAGES in shared memory.  Warp 8 is the
//     PRODUCER: per sub-stage it issues one 2D tensor-map TMA copy per staged plane
//     (cp.async.bulk.tensor -> SASS UTMALDG: the packed pixels of every source and their GRAY8
//     depth rows; rows outside the frame are zero-filled by the TMA unit) completing on the
//     sub-stage's "full" mbarrier, and it describes every chunk in shared memory (ChunkCtx: all
//     pointers pre-offset, coefficients, row ranges), so no consumer reads the job descriptor
//     on the critical path.  The ring is 2-8 sub-stages deep (by source count), i.e. the loads
//     run up to four chunks ahead of the arithmetic;
//   * the 8 CONSUMER warps own one row of every sub-stage: phase A turns the row into Y (two
//     dp2a per pixel, coefficients doubled so that the result is byte 2 of the sum) and into
//     14-bit pair-summed chroma (u | v<<16; six dp2a per pixel pair straight from the raw words
//     for 3-byte pixels, eight for 4-byte pixels) written to a separate CHROMA RING of 40 rows,
//     then releases the sub-stage ("empty" mbarrier) at once -- the ring, not the stage, carries
//     the 6 rows the vertical filter needs from chunk to chunk, so the 3+3 halo rows are
//     computed once per segment and nothing is copied between chunks;
//   * phase B (warp per chroma row, after the one consumer barrier of the chunk): the taps
//     [-58,-172,492,1786,1786,492,-172,-58]/4096 are symmetric and even: rows are pair-added on
//     the packed words, the three small taps are dp2a on the packed word (no unpacking), the
//     big one is one IMAD per channel; fast warps run ahead into phase A of the next chunk;
//   * composite: the select among the N staged sources happens in registers (tiles without
//     text) or is materialised into source 0's rows first (tiles a glyph touches).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "device_common.cuh"
#include "nes_internal.h"
#include "strips_common.cuh"

namespace nes {

namespace {

enum { MODE_ROWS = 0, MODE_SELECT = 1, MODE_MATERIALIZED = 2 };

constexpr int DEP_ROWB = STRIP_W;  // bytes of a staged depth row
// Source rows per compute chunk, per pixel-size class.  32 rows halve the per-chunk bookkeeping but the
// chroma ring (RING_ROWS = 40) then needs a second consumer barrier per chunk (phase A of the next chunk
// would overwrite rows phase B still reads).
#ifndef NES_CHUNK_ROWS_BPP3
#define NES_CHUNK_ROWS_BPP3 32  // measured: 4K 0.60 -> 0.64 of HBM peak against 16
#endif
template <int BPP>
struct Geo {
  static constexpr int CH = BPP == 3 ? NES_CHUNK_ROWS_BPP3 : CHUNK_ROWS;
  static constexpr int NSUB = CH / SUB_ROWS;
  static constexpr bool SECOND_BARRIER = 2 * CH + 6 > RING_ROWS;
};
// The chroma ring has RING_ROWS logical slots; slots [0, RING_MIRROR) are written twice (also
// RING_ROWS rows further) so that the 8 consecutive rows a vertical tap window reads never wrap.
constexpr int RING_MIRROR = 8;
constexpr int RING_PHYS = RING_ROWS + 8;

// Everything the consumer warps need to know about one chunk, written to shared memory by
// lane 0 of the producer warp while earlier chunks are being computed.  Plane pointers are
// pre-offset to the strip.
struct ChunkCtx {
  int32_t last;      // last chunk this CTA processes
  int32_t mode;      // MODE_*
  int32_t tma;       // rows were staged by TMA (else the consumers fill them)
  int32_t stamp;     // a glyph may intersect this chunk
  int32_t n_src, job;
  int32_t n_staged;  // sources laid out in a sub-stage (n_src when staged by TMA, else 1)
  int32_t x0, tw;
  int32_t yc0;       // frame row of chunk-local row 0 (negative above the frame)
  int32_t ra, rb;    // source rows of this chunk that exist / are needed [ra, rb)
  int32_t ya, yb;    // luma rows to emit [ya, yb)
  int32_t cA, cB;    // chroma rows to emit [cA, cB)
  int32_t H;
  int32_t edge;      // a vertical tap of [cA, cB) is clamped at the frame border
  int32_t rbase;     // chroma-ring slot of chunk-local row 0
  int32_t fullw;     // 16-byte aligned planes, strip width a multiple of 32: vector stores, lanes past the strip idle
  int32_t vec_out, rgb_base, dep_staged;
  int32_t nv12;      // chroma interleaved into one plane (su / du)
#ifdef NES_TRACE
  int32_t unit, k;   // work unit and chunk index inside it
#endif
  uint32_t a_mask;   // alpha byte mask of a pixel word (composite validity)
  uint32_t ky[4], ku[3], kv[3];
  int32_t sys, sus, svs, dys, dus, dvs;
  uint8_t *sy, *su, *sv, *dy, *du, *dv;  // + strip column offset
};

template <int BPP>
struct StripSmem {
  static constexpr int ROWB = STRIP_W * BPP;
  static constexpr int RING_ROWB = (STRIP_W / 2) * 4;
  // fixed part first, the sub-stage ring (ns x slot_bytes, chosen per launch) at the end
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_CTX = OFF_RING + RING_PHYS * RING_ROWB;
  static constexpr int CTX_BYTES = ((int)sizeof(ChunkCtx) + 15) & ~15;
  static constexpr int OFF_BAR = OFF_CTX + NCTX * CTX_BYTES;
  static constexpr int OFF_HITS = OFF_BAR + 2 * NS_MAX * 8;
  static constexpr int OFF_STAGE = (OFF_HITS + STRIP_HITS * (int)sizeof(DevPlaced) + 16 + 1023) & ~1023;
  // bytes of one sub-stage holding n sources: 8 pixel rows + 8 depth rows per source
  static constexpr int slot_bytes(int n) { return n * SUB_ROWS * (ROWB + DEP_ROWB); }
};


// Which job of this kernel's bpp class does work unit `u` of numbering phase `ph` belong to (unit_base is a
// prefix sum over the batch in which the jobs of the other class take no units): the last job whose base is <= u.
__device__ __forceinline__ int job_of_unit(const DevJob *jobs, int n_jobs, int cls, int ph, int u) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].unit_base[cls][ph] <= u) lo = mid; else hi = mid - 1;
  }
  return lo;
}


// Depth-select composite of the 8 pixels a lane owns in a sub-stage row (4 at column 4*lane, 4
// at 128 + 4*lane) from the N staged sources: the winner's pixel words and depth bytes.
// px / dep: shared addresses of source 0's row; source k's row is k*SUB_ROWS rows further.
// Semantics: DESIGN.md "composite" / oracle/overlay_port.c nes_oracle_composite.
template <int ROWB>
__device__ __forceinline__ void select_staged(uint32_t px, uint32_t dep, int n_src, uint32_t a_mask, int lane, uint32_t (&p)[8],
                                              uint32_t (&d4)[2]) {
  uint32_t bd[8];
  {  // source 0: taken wherever it is valid (nothing to compare against yet)
    const uint4 q[2] = {lds128(px + lane * 16), lds128(px + 512 + lane * 16)};
    const uint32_t dw[2] = {lds32(dep + lane * 4), lds32(dep + 128 + lane * 4)};
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t w[4] = {q[h].x, q[h].y, q[h].z, q[h].w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool valid = (w[i] & a_mask) != 0;
        bd[4 * h + i] = valid ? __byte_perm(dw[h], 0u, 0x4440 + i) : 256u;
        p[4 * h + i] = valid ? w[i] : 0u;
      }
    }
  }
#pragma unroll
  for (int k = 1; k < TMA_MAX_SOURCES; k++) {
    if (k >= n_src) break;
    const uint32_t rp = px + (uint32_t)(k * SUB_ROWS) * ROWB + lane * 16;
    const uint32_t dp = dep + (uint32_t)(k * SUB_ROWS) * DEP_ROWB + lane * 4;
    const uint4 q[2] = {lds128(rp), lds128(rp + 512)};
    const uint32_t dw[2] = {lds32(dp), lds32(dp + 128)};
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t w[4] = {q[h].x, q[h].y, q[h].z, q[h].w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t d = __byte_perm(dw[h], 0u, 0x4440 + i);
        const bool take = (w[i] & a_mask) != 0 && d < bd[4 * h + i];
        bd[4 * h + i] = take ? d : bd[4 * h + i];
        p[4 * h + i] = take ? w[i] : p[4 * h + i];
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t lo = __byte_perm(min(bd[4 * h], 255u), min(bd[4 * h + 1], 255u), 0x0040);
    const uint32_t hi = __byte_perm(min(bd[4 * h + 2], 255u), min(bd[4 * h + 3], 255u), 0x0040);
    d4[h] = __byte_perm(lo, hi, 0x5410);
  }
}

}  // namespace

#ifdef NES_TRACE
// diagnostic build only (make EXTRA=-DNES_TRACE OUT=...): per CTA {start, first stage full, end (ns, globaltimer), smid | chunks << 32}
__device__ unsigned long long g_trace[4 * 1024];
__device__ unsigned long long g_trace_units[2 * 16384];  // per unit {time its first chunk was taken up, cta | stamp chunks << 32}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
#endif

// The kernel body.  `jobs` is either the launch's descriptor table in global memory or (single-frame launches) the
// descriptor inside the kernel's own parameter block; `inline_glyphs` is that block's placed-glyph list (else null: the
// job's own device list is used).
template <int BPP>
__device__ __forceinline__ void frame_strips_body(const DevJob *__restrict__ jobs, const DevPlaced *__restrict__ inline_glyphs, int n_jobs, int unit_begin, int total_units,
                                                  int text_units, uint32_t *__restrict__ counters, int ns, int slot_bytes) {
  // programmatic dependent launch: the next launch on this stream (other frames: nothing of ours is its input)
  // may fill the SMs as our CTAs drain instead of waiting for the whole grid
  asm volatile("griddepcontrol.launch_dependents;");
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = StripSmem<BPP>;
  constexpr int ROWB = L::ROWB;
  constexpr int RING_ROWB = L::RING_ROWB;
  constexpr int CLS = BPP - 3;
  constexpr int NW = CONSUMER_WARPS;
  using G = Geo<BPP>;
  uint64_t *s_full = (uint64_t *)(smem + L::OFF_BAR);  // [NS_MAX]
  uint64_t *s_empty = s_full + NS_MAX;                  // [NS_MAX]
  DevPlaced *s_hits = (DevPlaced *)(smem + L::OFF_HITS);
  int *s_nhits = (int *)(s_hits + STRIP_HITS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t smem_base = smem_u32(smem);

#ifdef NES_TRACE
  if (tid == 0 && blockIdx.x < 1024) g_trace[4 * blockIdx.x] = gtime();
#endif
  if (tid == 0) {
    for (int i = 0; i < NS_MAX; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {
    // =========================== producer warp ===========================================
    // the first unit is static (grid <= total_units: every CTA has work), the rest come from a
    // global counter
    // (a launch may cover only the unit range [unit_begin, total_units): banded low-latency submits)
    int cur_u = unit_begin + (int)blockIdx.x, next_u = 0;
    if (lane == 0) next_u = unit_begin + (int)gridDim.x + (int)atomicAdd(&counters[0], 1u);
    next_u = __shfl_sync(0xffffffffu, next_u, 0);
    int q = 0, par = 0, round0 = 1;  // sub-stage cursor: slot, parity of its use count, first trip round the ring
    int chunk_it = 0, rbase = 0;
    while (cur_u < total_units) {
      // ---- unit geometry (all lanes compute it: cheap, keeps the copy code uniform) ---------
      // units [0, text_units) are the segments with text of every job (phase 0), the rest follow (phase 1)
      const int ph = cur_u < text_units ? 0 : 1;
      const int pu = ph ? cur_u - text_units : cur_u;
      const int j = job_of_unit(jobs, n_jobs, CLS, ph, pu);
      const DevJob *jp = jobs + j;
      const int local = pu - jp->unit_base[CLS][ph];
      const int strip = local % jp->strips_x;
      const int seg = jp->seg_order[(ph ? jp->n_text_segs : 0) + local / jp->strips_x];
      const int W = jp->W, H = jp->H, S = jp->seg_rows, n_src = jp->n_src;
      const int x0 = strip * STRIP_W, tw = min(STRIP_W, W - x0);
      const int Y0 = seg * S, Y1 = min(Y0 + S, H);
      const int L0 = Y0 - HALO, need_end = min(Y1 + HALO, H);
      const int nchunks = (need_end - L0 + G::CH - 1) / G::CH;
      const int tma = jp->tma_ok;
      const int n_staged = tma ? n_src : 1;
      const int dep_staged = (jp->dy != nullptr) || n_src > 1;
      const int nbands = (H + (1 << MASK_BAND_SHIFT) - 1) >> MASK_BAND_SHIFT;
      const uint32_t tx_bytes = (uint32_t)(n_staged * SUB_ROWS) * (uint32_t)(ROWB + (dep_staged ? DEP_ROWB : 0));
      // which plane this lane copies: lanes [0, n) the pixel planes, lanes [n, 2n) the depth planes
      const TMap *my_map = nullptr;
      uint32_t my_off = 0;
      int my_x = 0;
      if (tma) {
        if (lane < n_staged) {
          my_map = &jp->tmap_px[lane]; my_off = (uint32_t)(lane * SUB_ROWS) * ROWB; my_x = (x0 * BPP) >> 2;
        } else if (dep_staged && lane < 2 * n_staged) {
          my_map = &jp->tmap_dep[lane - n_staged];
          my_off = (uint32_t)(n_staged * SUB_ROWS) * ROWB + (uint32_t)((lane - n_staged) * SUB_ROWS) * DEP_ROWB;
          my_x = x0 >> 2;
        }
      }
      for (int k = 0; k < nchunks; k++, chunk_it++) {
        const int yc0 = L0 + k * G::CH;
        const int ra = max(yc0, 0), rb = min(yc0 + G::CH, need_end);
        const bool last_k = (k == nchunks - 1);
        if (lane == 0) {
          // the context slot was last used NCTX chunks ago; the consumers are at most
          // NS_MAX/2 + 2 chunks behind (they must release sub-stages for us to get here)
          ChunkCtx &c = *(ChunkCtx *)(smem + L::OFF_CTX + (chunk_it & (NCTX - 1)) * L::CTX_BYTES);
          int stamp = 0;
          if (jp->n_glyphs > 0) {
            stamp = 1;
            if (jp->use_mask) {
              stamp = 0;
              for (int band = ra >> MASK_BAND_SHIFT; band <= (rb - 1) >> MASK_BAND_SHIFT && band < nbands; band++) {
                const int bit = band * jp->strips_x + strip;
                stamp |= (jp->tile_mask[bit >> 5] >> (bit & 31)) & 1u;
              }
            }
          }
#ifdef NES_TRACE
          c.unit = cur_u; c.k = k;
#endif
          c.last = last_k && next_u >= total_units;
          c.tma = tma;
          c.stamp = stamp;
          c.mode = (tma && n_src > 1) ? (stamp ? MODE_MATERIALIZED : MODE_SELECT) : MODE_ROWS;
          c.n_src = n_src; c.job = j; c.n_staged = n_staged;
          c.x0 = x0; c.tw = tw; c.yc0 = yc0; c.ra = ra; c.rb = rb;
          c.ya = max(ra, Y0); c.yb = min(rb, Y1);
          const int cA = (k == 0) ? (Y0 >> 1) : (Y0 >> 1) + ((k * G::CH) >> 1) - 3;
          const int cB = last_k ? (Y1 >> 1) : (Y0 >> 1) + (((k + 1) * G::CH) >> 1) - 3;
          c.cA = cA; c.cB = cB; c.H = H;
          c.edge = (2 * cA - 3 < 0) || (2 * (cB - 1) + 4 > H - 1);
          c.rbase = rbase;
          c.vec_out = jp->out_vec;
          c.fullw = jp->out_vec && (tw & 31) == 0;
          c.a_mask = jp->a_off >= 0 ? (0xFFu << (8 * jp->a_off)) : 0xFFFFFFFFu;
          c.rgb_base = jp->rgb_base;
          c.dep_staged = dep_staged;
          if (BPP == 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) c.ky[i] = jp->ky3[i];
#pragma unroll
            for (int i = 0; i < 3; i++) { c.ku[i] = jp->ku3[i]; c.kv[i] = jp->kv3[i]; }
          } else {
            c.ky[0] = jp->ky2[0]; c.ky[1] = jp->ky2[1];
            c.ku[0] = jp->ku[0]; c.ku[1] = jp->ku[1]; c.kv[0] = jp->kv[0]; c.kv[1] = jp->kv[1];
          }
          c.sys = jp->sys; c.sus = jp->sus; c.svs = jp->svs; c.dys = jp->dys; c.dus = jp->dus; c.dvs = jp->dvs;
          const int nv12 = jp->nv12;
          const int xc = nv12 ? x0 : (x0 >> 1);  // byte offset of the strip inside a chroma row
          c.nv12 = nv12;
          c.sy = jp->sy + x0; c.su = jp->su + xc; c.sv = nv12 ? nullptr : jp->sv + xc;
          c.dy = jp->dy ? jp->dy + x0 : nullptr;
          c.du = jp->dy ? jp->du + xc : nullptr;
          c.dv = (jp->dy && !nv12) ? jp->dv + xc : nullptr;
        }
        __syncwarp();
#pragma unroll 1
        for (int sub = 0; sub < G::NSUB; sub++) {
          if (!round0) mbar_wait(&s_empty[q], (uint32_t)(par ^ 1));
          const int ys = yc0 + sub * SUB_ROWS;
          const bool wanted = tma && ys < rb && ys + SUB_ROWS > ra;  // else nothing of this sub-stage is read
          if (wanted) {
            if (lane == 0) mbar_arrive_expect_tx(&s_full[q], tx_bytes);
            __syncwarp();
            if (my_map) tma_load_2d(smem_base + L::OFF_STAGE + (uint32_t)(q * slot_bytes) + my_off, my_map, my_x, ys, &s_full[q]);
          } else if (lane == 0) {
            mbar_arrive(&s_full[q]);  // nothing in flight (the consumers fill the rows themselves, or skip them)
          }
          if (++q == ns) { q = 0; par ^= 1; round0 = 0; }
        }
        rbase += G::CH;
        if (rbase >= RING_ROWS) rbase -= RING_ROWS;
      }
      cur_u = next_u;
      if (lane == 0 && cur_u < total_units) next_u = unit_begin + (int)gridDim.x + (int)atomicAdd(&counters[0], 1u);
      next_u = __shfl_sync(0xffffffffu, next_u, 0);
    }
    // the last CTA to run out of work re-arms the counters for the next launch on this stream
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) { counters[0] = 0; counters[1] = 0; __threadfence(); }
    }
    return;
  }

  // ============================= consumer warps ===========================================
  const uint32_t ring0 = smem_base + L::OFF_RING;
  const uint32_t full0 = smem_base + L::OFF_BAR, empty0 = full0 + NS_MAX * 8;
  int qc = 0, parc = 0;  // sub-stage cursor of the chunk's first sub-stage: slot, parity of its use count
  for (int chunk_it = 0;; chunk_it++) {
    // the chunk's sub-stages (the ring may wrap between them)
    int q[G::NSUB], par[G::NSUB];
    q[0] = qc; par[0] = parc;
#pragma unroll
    for (int i = 1; i < G::NSUB; i++) {
      q[i] = q[i - 1] + 1; par[i] = par[i - 1];
      if (q[i] == ns) { q[i] = 0; par[i] ^= 1; }
    }
    mbar_wait_a(full0 + q[0] * 8, (uint32_t)par[0]);
#ifdef NES_TRACE
    if (tid == 0 && chunk_it == 0 && blockIdx.x < 1024) g_trace[4 * blockIdx.x + 1] = gtime();
#endif
    const ChunkCtx &c = *(const ChunkCtx *)(smem + L::OFF_CTX + (chunk_it & (NCTX - 1)) * L::CTX_BYTES);
#ifdef NES_TRACE
    if (tid == 0 && c.unit < 16384) {
      if (c.k == 0) { g_trace_units[2 * c.unit] = gtime(); g_trace_units[2 * c.unit + 1] = blockIdx.x; }
      if (c.stamp) g_trace_units[2 * c.unit + 1] += 1ull << 32;
    }
#endif
    const int last = c.last;
    {
      uint32_t sb[G::NSUB];
#pragma unroll
      for (int i = 0; i < G::NSUB; i++) sb[i] = smem_base + L::OFF_STAGE + (uint32_t)(q[i] * slot_bytes);
      const uint32_t dep_off = (uint32_t)(c.n_staged * SUB_ROWS) * ROWB;
      const int x0 = c.x0, tw = c.tw, yc0 = c.yc0, ra = c.ra, rb = c.rb, ya = c.ya, yb = c.yb;
      const int n_src = c.n_src;
      int mode = c.mode;
      const bool fullw = c.fullw != 0, vec_out = c.vec_out != 0;
      const bool dep_staged = c.dep_staged != 0;
      uint8_t *const dy = c.dy;
      const int dys = c.dys;

      if (!c.tma || mode == MODE_MATERIALIZED || c.stamp) {
        // whole-chunk work on the staged rows: needs all of its sub-stages
#pragma unroll
        for (int i = 1; i < G::NSUB; i++) mbar_wait_a(full0 + q[i] * 8, (uint32_t)par[i]);
        // ---- fill our rows ourselves when they were not staged by TMA ------------------------
        if (!c.tma) {
          const DevJob &jb = jobs[c.job];
#pragma unroll
          for (int i = 0; i < G::NSUB; i++) {
            const int y = yc0 + warp + i * SUB_ROWS;
            if (y < ra || y >= rb) continue;
            uint8_t *const gi = smem + L::OFF_STAGE + q[i] * slot_bytes;
            uint8_t *sp = gi + warp * ROWB;
            uint8_t *sd = gi + SUB_ROWS * ROWB + warp * DEP_ROWB;
            if (n_src == 1) {
              const uint8_t *gp = jb.src[0].rgb + (size_t)y * jb.src[0].rgb_stride + (size_t)x0 * BPP;
              const int nbytes = tw * BPP;
              for (int bb = lane; bb < nbytes; bb += 32) sp[bb] = gp[bb];
              if (dep_staged) {
                const uint8_t *gd = jb.src[0].depth + (size_t)y * jb.src[0].depth_stride + x0;
                for (int x = lane; x < tw; x += 32) sd[x] = gd[x];
              }
            } else {
              for (int x = lane; x < tw; x += 32) {
                uint32_t d;
                composite_px<BPP>(jb, x0 + x, y, sp + x * BPP, &d);
                sd[x] = (uint8_t)d;
              }
            }
          }
        }
        // ---- staged composite under text: materialise the select into source 0's rows ----------
        if (BPP == 4 && mode == MODE_MATERIALIZED) {
#pragma unroll
          for (int i = 0; i < G::NSUB; i++) {
            const int y = yc0 + warp + i * SUB_ROWS;
            if (y < ra || y >= rb) continue;
            const uint32_t sbi = sb[i];
            const uint32_t px = sbi + warp * ROWB;
            const uint32_t dp = sbi + dep_off + warp * DEP_ROWB;
            uint32_t p[8], d4[2];
            select_staged<ROWB>(px, dp, n_src, c.a_mask, lane, p, d4);
            __syncwarp();
            sts128(px + lane * 16, p[0], p[1], p[2], p[3]);
            sts128(px + 512 + lane * 16, p[4], p[5], p[6], p[7]);
            sts32(dp + lane * 4, d4[0]);
            sts32(dp + 128 + lane * 4, d4[1]);
          }
          mode = MODE_ROWS;
        }
        __syncwarp();
        // ---- text overlay, stamped into the staged rows of source 0 ----------------------------
        if (c.stamp) stamp_chunk<BPP, STRIP_W>(jobs[c.job], inline_glyphs, smem + L::OFF_STAGE, qc, ns, slot_bytes, x0, x0 + tw, yc0, ra, rb, c.rgb_base, s_hits, s_nhits);
        fence_proxy_async();  // our generic-proxy writes to the stage come before the TMA refill
      }

      // ---- phase A: per source row: Y out, pair-summed chroma (u14 | v14<<16) to the ring ------
      {
        uint8_t *const sy = c.sy;
        const int sys = c.sys;
        const int rbase = c.rbase;
#pragma unroll
        for (int i = 0; i < G::NSUB; i++) {
          const int r = warp + i * SUB_ROWS;
          const int y = yc0 + r;
          if (i >= 1) mbar_wait_a(full0 + q[i] * 8, (uint32_t)par[i]);
          if (y >= ra && y < rb) {
            const uint32_t row = sb[i] + warp * ROWB;
            const uint32_t drow = sb[i] + dep_off + warp * DEP_ROWB;
            int slot = rbase + r;
            if (slot >= RING_ROWS) slot -= RING_ROWS;
            const uint32_t crow = ring0 + slot * RING_ROWB;
            const bool mirror = slot < RING_MIRROR;
            const bool core = (y >= ya) && (y < yb);
            uint32_t uv[4];
            if (BPP == 3) {
              // lane owns pixels 8*lane .. 8*lane+7 = 6 raw words (8-byte loads at 24-byte stride are
              // conflict free); the coefficient pairs are laid out per byte phase, so nothing is unpacked
              const uint2 a = lds64(row + lane * 24), b = lds64(row + lane * 24 + 8), d = lds64(row + lane * 24 + 16);
              const uint32_t w[6] = {a.x, a.y, b.x, b.y, d.x, d.y};
              const uint32_t u01 = c.ku[0], u20 = c.ku[1], u12 = c.ku[2], v01 = c.kv[0], v20 = c.kv[1], v12 = c.kv[2];
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const uint32_t w0 = w[3 * h], w1 = w[3 * h + 1], w2 = w[3 * h + 2];
                const int su0 = dp2a_lo(u12, w1, dp2a_hi(u20, w0, dp2a_lo(u01, w0, C_BIAS)));
                const int sv0 = dp2a_lo(v12, w1, dp2a_hi(v20, w0, dp2a_lo(v01, w0, C_BIAS)));
                const int su1 = dp2a_hi(u12, w2, dp2a_lo(u20, w2, dp2a_hi(u01, w1, C_BIAS)));
                const int sv1 = dp2a_hi(v12, w2, dp2a_lo(v20, w2, dp2a_hi(v01, w1, C_BIAS)));
                uv[2 * h] = pack_uv14(su0, sv0);
                uv[2 * h + 1] = pack_uv14(su1, sv1);
              }
              sts128(crow + lane * 16, uv[0], uv[1], uv[2], uv[3]);  // chroma cols 4*lane..+3
              if (mirror) sts128(crow + RING_ROWS * RING_ROWB + lane * 16, uv[0], uv[1], uv[2], uv[3]);
              if (core) {
                const uint32_t y01 = c.ky[0], y2_ = c.ky[1], y_0 = c.ky[2], y12 = c.ky[3];
                uint32_t sm[8];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  const uint32_t w0 = w[3 * h], w1 = w[3 * h + 1], w2 = w[3 * h + 2];
                  sm[4 * h] = dp2a_hi_uu(y2_, w0, dp2a_lo_uu(y01, w0, 2 * Y_BIAS));
                  sm[4 * h + 1] = dp2a_lo_uu(y12, w1, dp2a_hi_uu(y_0, w0, 2 * Y_BIAS));
                  sm[4 * h + 2] = dp2a_lo_uu(y2_, w2, dp2a_hi_uu(y01, w1, 2 * Y_BIAS));
                  sm[4 * h + 3] = dp2a_hi_uu(y12, w2, dp2a_lo_uu(y_0, w2, 2 * Y_BIAS));
                }
                const uint32_t yw0 = pack_b2(sm[0], sm[1], sm[2], sm[3]), yw1 = pack_b2(sm[4], sm[5], sm[6], sm[7]);
                const uint32_t o = (uint32_t)(y * sys + lane * 8);
                if (fullw) { if (lane * 8 < tw) stg64(sy + o, yw0, yw1); }
                else store8(sy + y * sys, lane * 8, yw0, yw1, tw, vec_out);
                if (dy) {
                  const uint2 dd = lds64(drow + lane * 8);
                  const uint32_t g0 = gray_y4_packed(dd.x), g1 = gray_y4_packed(dd.y);
                  if (fullw) { if (lane * 8 < tw) stg64(dy + (uint32_t)(y * dys + lane * 8), g0, g1); }
                  else store8(dy + y * dys, lane * 8, g0, g1, tw, vec_out);
                }
              }
            } else {
              // lane owns pixels 4*lane..+3 and 128+4*lane..+3 (16-byte accesses at 16-byte stride)
              uint32_t p[8];
              uint32_t d4[2] = {0, 0};
              if (mode == MODE_SELECT) {
                select_staged<ROWB>(row, drow, n_src, c.a_mask, lane, p, d4);
              } else {
                const uint4 a = lds128(row + lane * 16), b = lds128(row + 512 + lane * 16);
                p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
                if (dep_staged) { d4[0] = lds32(drow + lane * 4); d4[1] = lds32(drow + 128 + lane * 4); }
              }
              const uint32_t kua = c.ku[0], kub = c.ku[1], kva = c.kv[0], kvb = c.kv[1];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                int su = dp2a_lo(kua, p[2 * j], C_BIAS); su = dp2a_hi(kub, p[2 * j], su);
                su = dp2a_lo(kua, p[2 * j + 1], su); su = dp2a_hi(kub, p[2 * j + 1], su);
                int sv = dp2a_lo(kva, p[2 * j], C_BIAS); sv = dp2a_hi(kvb, p[2 * j], sv);
                sv = dp2a_lo(kva, p[2 * j + 1], sv); sv = dp2a_hi(kvb, p[2 * j + 1], sv);
                uv[j] = pack_uv14(su, sv);
              }
              sts64(crow + lane * 8, uv[0], uv[1]);        // chroma cols 2*lane, 2*lane+1
              sts64(crow + 256 + lane * 8, uv[2], uv[3]);  // chroma cols 64+2*lane, +1
              if (mirror) {
                sts64(crow + RING_ROWS * RING_ROWB + lane * 8, uv[0], uv[1]);
                sts64(crow + RING_ROWS * RING_ROWB + 256 + lane * 8, uv[2], uv[3]);
              }
              if (core) {
                const uint32_t kya = c.ky[0], kyb = c.ky[1];
                uint32_t sm[8];
#pragma unroll
                for (int k = 0; k < 8; k++) sm[k] = dp2a_hi_uu(kyb, p[k], dp2a_lo_uu(kya, p[k], 2 * Y_BIAS));
                const uint32_t yw0 = pack_b2(sm[0], sm[1], sm[2], sm[3]), yw1 = pack_b2(sm[4], sm[5], sm[6], sm[7]);
                if (fullw) {
                  uint8_t *o = sy + (uint32_t)(y * sys + lane * 4);
                  if (lane * 4 < tw) stg32(o, yw0);
                  if (128 + lane * 4 < tw) stg32(o + 128, yw1);
                } else {
                  uint8_t *o = sy + y * sys;
                  store4(o, lane * 4, yw0, tw, vec_out); store4(o, 128 + lane * 4, yw1, tw, vec_out);
                }
                if (dy) {
                  const uint32_t g0 = gray_y4_packed(d4[0]), g1 = gray_y4_packed(d4[1]);
                  if (fullw) {
                    uint8_t *od = dy + (uint32_t)(y * dys + lane * 4);
                    if (lane * 4 < tw) stg32(od, g0);
                    if (128 + lane * 4 < tw) stg32(od + 128, g1);
                  } else {
                    uint8_t *od = dy + y * dys;
                    store4(od, lane * 4, g0, tw, vec_out); store4(od, 128 + lane * 4, g1, tw, vec_out);
                  }
                }
              }
            }
          }
          // this warp is done with its row of the sub-stage: release it to the producer
          __syncwarp();
          if (lane == 0) mbar_arrive_a(empty0 + q[i] * 8);
        }
      }
      consumer_sync();

      // ---- phase B: 8-tap vertical bicubic on chroma; edge taps fold = clamped row index ------
      // T/2 = [-29,-86,246,893,893,246,-86,-29] on 14-bit samples, out = (2^16 + sum) >> 17:
      // the symmetric rows are added on the packed words first (2*15360 < 32768: no carry), the
      // small taps are dp2a on the packed word (bytes (t,0,0,t): .lo picks u, .hi picks v;
      // 246 = 2*123 on the doubled sum), the big one is one IMAD per channel.
      {
        const int cc = lane * 4;  // chroma column inside the strip
        const int cw = tw >> 1;
        const int cA = c.cA, cB = c.cB;
        const int rbase = c.rbase;
        uint8_t *const du = c.du, *const dv = c.dv;
        const int dus = c.dus, dvs = c.dvs;
        // depth chroma planes are constant 128: 16-byte stores, one
        // warp-store covers 4 rows x 128 bytes of one plane (NV12: 2 rows x 256 bytes of the UV plane)
        if (dy && fullw) {
          if (!c.nv12) {
            const int ngroups = (cB - cA + 3) >> 2;
            for (int t = warp; t < 2 * ngroups; t += NW) {
              const int row = cA + 4 * (t >> 1) + (lane >> 3);
              if (row < cB && (lane & 7) * 16 < cw) {
                uint8_t *o = (t & 1) ? dv + (uint32_t)(row * dvs) : du + (uint32_t)(row * dus);
                *(uint4 *)(o + (lane & 7) * 16) = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
              }
            }
          } else {
            for (int t = warp; 2 * t < cB - cA; t += NW) {
              const int row = cA + 2 * t + (lane >> 4);
              if (row < cB && (lane & 15) * 16 < 2 * cw)
                *(uint4 *)(du + (uint32_t)(row * dus) + (lane & 15) * 16) = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
            }
          }
        }
        if (cc < cw) {
          uint8_t *const su_ = c.su, *const sv_ = c.sv;
          const int sus = c.sus, svs = c.svs;
          const bool edge = c.edge != 0;
          const uint32_t col = ring0 + cc * 4;
#pragma unroll 1
          for (int ci = cA + warp; ci < cB; ci += NW) {
            uint32_t t[8][4];
            if (!edge) {
              int s0 = rbase + (2 * ci - 3 - yc0);  // >= rbase - 6; after the wrap s0 is in [0, RING_ROWS]: rows s0..s0+7 <= 47 (mirror)
              if (s0 < 0) s0 += RING_ROWS;
              if (s0 > RING_ROWS) s0 -= RING_ROWS;
              const uint32_t base = col + (uint32_t)(s0 * RING_ROWB);
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const uint4 qv = lds128(base + j * RING_ROWB);
                t[j][0] = qv.x; t[j][1] = qv.y; t[j][2] = qv.z; t[j][3] = qv.w;
              }
            } else {
              const int H = c.H;
#pragma unroll
              for (int j = 0; j < 8; j++) {
                int sj = rbase + min(max(2 * ci - 3 + j, 0), H - 1) - yc0;
                if (sj < 0) sj += RING_ROWS;
                if (sj >= RING_ROWS) sj -= RING_ROWS;
                const uint4 qv = lds128(col + (uint32_t)(sj * RING_ROWB));
                t[j][0] = qv.x; t[j][1] = qv.y; t[j][2] = qv.z; t[j][3] = qv.w;
              }
            }
            uint32_t us[4], vs[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint32_t a07 = t[0][k] + t[7][k], a16 = t[1][k] + t[6][k], a25 = (t[2][k] + t[5][k]) * 2u, a34 = t[3][k] + t[4][k];
              int au = dp2a_lo_us(a07, 0xE30000E3u, 1 << 16);  // -29
              au = dp2a_lo_us(a16, 0xAA0000AAu, au);           // -86
              au = dp2a_lo_us(a25, 0x7B00007Bu, au);           // 123 * 2
              au += 893 * (int)(a34 & 0xFFFFu);
              int av = dp2a_hi_us(a07, 0xE30000E3u, 1 << 16);
              av = dp2a_hi_us(a16, 0xAA0000AAu, av);
              av = dp2a_hi_us(a25, 0x7B00007Bu, av);
              av += 893 * (int)(a34 >> 16);
              us[k] = clip8_relu(au >> 17);
              vs[k] = clip8_relu(av >> 17);
            }
            const uint32_t ub = __byte_perm(__byte_perm(us[0], us[1], 0x0040), __byte_perm(us[2], us[3], 0x0040), 0x5410);
            const uint32_t vb = __byte_perm(__byte_perm(vs[0], vs[1], 0x0040), __byte_perm(vs[2], vs[3], 0x0040), 0x5410);
            if (c.nv12) {
              // U0 V0 U1 V1 | U2 V2 U3 V3: chroma column cc sits at byte 2*cc of the UV row
              const uint32_t w0 = __byte_perm(ub, vb, 0x5140), w1 = __byte_perm(ub, vb, 0x7362);
              if (fullw) {
                stg64(su_ + (uint32_t)(ci * sus + 2 * cc), w0, w1);
              } else {
                store8(su_ + ci * sus, 2 * cc, w0, w1, 2 * cw, vec_out);
                if (dy) store8(du + ci * dus, 2 * cc, 0x80808080u, 0x80808080u, 2 * cw, vec_out);
              }
            } else if (fullw) {
              stg32(su_ + (uint32_t)(ci * sus + cc), ub);
              stg32(sv_ + (uint32_t)(ci * svs + cc), vb);
            } else {
              store4(su_ + ci * sus, cc, ub, cw, vec_out);
              store4(sv_ + ci * svs, cc, vb, cw, vec_out);
              if (dy) {
                store4(du + ci * dus, cc, 0x80808080u, cw, vec_out);
                store4(dv + ci * dvs, cc, 0x80808080u, cw, vec_out);
              }
            }
          }
        }
      }
    }
#ifdef NES_TRACE
    if (last && tid == 0 && blockIdx.x < 1024) {
      g_trace[4 * blockIdx.x + 2] = gtime();
      g_trace[4 * blockIdx.x + 3] = (unsigned long long)smid() | ((unsigned long long)(chunk_it + 1) << 32);
    }
#endif
    if (last) break;
    if (G::SECOND_BARRIER) consumer_sync();  // phase A of the next chunk reuses ring rows phase B was reading
    qc = q[G::NSUB - 1] + 1; parc = par[G::NSUB - 1];
    if (qc == ns) { qc = 0; parc ^= 1; }
  }
}

template <int BPP>
__global__ void __launch_bounds__(CTA_THREADS, 3)
k_frame_strips(const DevJob *__restrict__ jobs, int n_jobs, int unit_begin, int total_units, int text_units, uint32_t *__restrict__ counters, int ns, int slot_bytes) {
  frame_strips_body<BPP>(jobs, nullptr, n_jobs, unit_begin, total_units, text_units, counters, ns, slot_bytes);
}

// Single-frame launches: the descriptor, its tensor maps and its placed glyphs travel in the kernel's parameter block
// (__grid_constant__: the producer warp and the TMA unit read them in place), so a frame costs no descriptor upload,
// no event and no wait on the host -- and consecutive frames of a stream are adjacent launches that overlap on the device.
template <int BPP>
__global__ void __launch_bounds__(CTA_THREADS, 3)
k_frame_strips_1(const __grid_constant__ JobPack pack, int unit_begin, int total_units, int text_units, uint32_t *__restrict__ counters, int ns, int slot_bytes) {
  frame_strips_body<BPP>(&pack.job, pack.glyphs, 1, unit_begin, total_units, text_units, counters, ns, slot_bytes);
}

#ifdef NES_TRACE
extern "C" __attribute__((visibility("default"))) int nes_debug_read_trace(unsigned long long *out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * (size_t)n);
}
extern "C" __attribute__((visibility("default"))) int nes_debug_read_trace_units(unsigned long long *out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_units, sizeof(unsigned long long) * (size_t)n);
}
extern "C" __attribute__((visibility("default"))) int nes_debug_clear_trace_units() {
  void *p = nullptr;
  cudaGetSymbolAddress(&p, g_trace_units);
  return (int)cudaMemset(p, 0, sizeof(g_trace_units));
}
#endif

static int g_num_sms = 0;
static int g_smem_per_sm = 0, g_smem_reserved = 1024;
static int g_force_ctas = 0;  // NES_STRIPS_CTAS=2|3 (experiments)

int frame_strips_init() {
  cudaError_t e;
  int dev = 0;
  cudaGetDevice(&dev);
  e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&g_smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceGetAttribute(&g_smem_reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
  int optin = 0;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return (int)e;
  const int two = std::min(optin, g_smem_per_sm / 2 - g_smem_reserved);
  e = cudaFuncSetAttribute(k_frame_strips<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, two);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_frame_strips<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, two);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_frame_strips_1<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, two);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_frame_strips_1<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, two);
  if (e != cudaSuccess) return (int)e;
  if (const char *v = getenv("NES_STRIPS_CTAS")) g_force_ctas = atoi(v);
  return 0;
}

// Launch shape of one pixel-size class: the sub-stage slot is the largest any job of the launch
// needs (so that a slot index means the same shared-memory range for every unit a CTA walks
// through); three CTAs per SM when at least two slots fit a third of the SM's shared memory
// (latency hiding by occupancy), else two CTAs with a deeper ring.
struct StripsConfig {
  int ns, slot, smem, ctas;
};
static StripsConfig strips_config(int cls, int staged) {
  const int off = cls == 0 ? StripSmem<3>::OFF_STAGE : StripSmem<4>::OFF_STAGE;
  const int slot = cls == 0 ? StripSmem<3>::slot_bytes(staged) : StripSmem<4>::slot_bytes(staged);
  const int smem_sm = g_smem_per_sm > 0 ? g_smem_per_sm : 233472;
  StripsConfig c{0, slot, 0, 0};
  for (int ctas = 3; ctas >= 1; ctas--) {
    if (g_force_ctas && ctas > g_force_ctas) continue;
    const int budget = smem_sm / ctas - g_smem_reserved - off;
    const int ns = std::min(NS_MAX, budget / slot);
    if (ns >= std::max(2, cls == 0 ? Geo<3>::NSUB : Geo<4>::NSUB)) { c.ns = ns; c.ctas = ctas; c.smem = off + ns * slot; break; }
  }
  return c;
}
static void launch_staged(const DevJob *jobs, int n_jobs, int staged[2]) {
  staged[0] = staged[1] = 1;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs[j];
    if (!jb.general && jb.tma_ok && jb.n_src > staged[jb.bpp - 3]) staged[jb.bpp - 3] = jb.n_src;
  }
}

// Host-side planning of a launch: one segment height per bpp class (trade: 6 halo rows per
// segment against the one-unit tail of the persistent grid; seg_rows + 6 is a multiple of the
// chunk height so that no staged row is wasted), unit numbering.  Jobs of the other class
// (and general jobs) take no units.  Segments that carry text (the job's tile_mask) are numbered first, over
// the whole launch: a chunk with text costs several plain ones, and a unit that is started last is the tail.
void plan_frame_strips(DevJob *jobs, int n_jobs, bool text_first) {
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  int staged[2];
  launch_staged(jobs, n_jobs, staged);
  for (int cls = 0; cls < 2; cls++) {
    const int grid = sms * std::max(1, strips_config(cls, staged[cls]).ctas);
    const int ch = cls == 0 ? Geo<3>::CH : Geo<4>::CH;
    int s_min = std::max(ch, 32) - 2 * HALO;
    for (int j = 0; j < n_jobs; j++)
      if (!jobs[j].general && jobs[j].bpp == 3 + cls)
        while ((jobs[j].H + s_min - 1) / s_min > MAX_SEGS) s_min += ch;
    int best_s = s_min;
    double best_cost = 1e30;
    for (int S = s_min; S <= std::max(256, s_min); S += ch) {
      double work = 0;
      long units = 0;
      for (int j = 0; j < n_jobs; j++) {
        const DevJob &jb = jobs[j];
        if (jb.general || jb.bpp != 3 + cls) continue;
        const int strips = (jb.W + STRIP_W - 1) / STRIP_W, segs = (jb.H + S - 1) / S;
        units += (long)strips * segs;
        work += (double)strips * (jb.H + segs * 6.2);  // what a segment costs beyond its rows (halo rows, pipeline fill), in rows: fitted on B200
      }
      if (units == 0) break;
      const double unit_cost = S + 6.2;
      // more units than CTAs: the work per CTA plus a small share of one unit for the ragged end -- consecutive launches overlap
      // (programmatic dependent launch), so the tail of a launch is mostly filled by the next one
      const double cost = units <= grid ? unit_cost : work / grid + 0.1 * unit_cost;
      if (cost < best_cost - 1e-9) { best_cost = cost; best_s = S; }
    }
    static const int seg_override = [] { const char *v = getenv("NES_STRIPS_SEG"); return v ? atoi(v) : 0; }();  // experiments: fixed segment height
    if (seg_override >= s_min) best_s = s_min + (seg_override - s_min) / ch * ch;
    int base[2] = {0, 0};
    for (int j = 0; j < n_jobs; j++) {
      DevJob &jb = jobs[j];
      jb.unit_base[cls][0] = base[0];
      jb.unit_base[cls][1] = base[1];
      if (jb.general || jb.bpp != 3 + cls) continue;
      const int S = best_s;
      jb.seg_rows = S;
      jb.strips_x = (jb.W + STRIP_W - 1) / STRIP_W;
      jb.segs_y = (jb.H + S - 1) / S;
      jb.n_units = jb.strips_x * jb.segs_y;
      // segments with text first
      int n_text = 0, n_plain = 0;
      uint8_t plain[MAX_SEGS];
      const int nbands = (jb.H + (1 << MASK_BAND_SHIFT) - 1) >> MASK_BAND_SHIFT;
      for (int sg = 0; sg < jb.segs_y; sg++) {
        bool text = false;
        if (text_first && jb.n_glyphs > 0 && jb.use_mask) {
          const int r0 = std::max(sg * S - HALO, 0), r1 = std::min((sg + 1) * S + HALO, jb.H);
          for (int band = r0 >> MASK_BAND_SHIFT; band <= (r1 - 1) >> MASK_BAND_SHIFT && band < nbands && band < 256 && !text; band++)
            text = (jb.band_text[band >> 5] >> (band & 31)) & 1u;
        }
        if (text) jb.seg_order[n_text++] = (uint8_t)sg; else plain[n_plain++] = (uint8_t)sg;
      }
      for (int i = 0; i < n_plain; i++) jb.seg_order[n_text + i] = plain[i];
      jb.n_text_segs = n_text;
      base[0] += n_text * jb.strips_x;
      base[1] += n_plain * jb.strips_x;
    }
  }
}

static bool g_pdl = true;  // NES_NO_PDL=1: plain stream-ordered launches (A/B experiments)

template <int BPP>
static cudaError_t launch_one(int grid, int smem, cudaStream_t st, const DevJob *jobs, int n_jobs, int u0, int u1, int text_units, uint32_t *counters, int ns,
                              int slot) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_frame_strips<BPP>, jobs, n_jobs, u0, u1, text_units, counters, ns, slot);
}

static bool g_inline = true;  // NES_NO_INLINE=1: single-frame launches read their descriptor from the table like batches do

template <int BPP>
static cudaError_t launch_one_inline(int grid, int smem, cudaStream_t st, const JobPack &pack, int u0, int u1, int text_units, uint32_t *counters, int ns, int slot) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_frame_strips_1<BPP>, pack, u0, u1, text_units, counters, ns, slot);
}

// A single same-size frame whose placed glyphs fit the parameter block can be launched without a device-resident descriptor.
bool frame_strips_inline_ok(const DevJob *jobs_host, int n_jobs) {
  static const bool once = [] { if (const char *v = getenv("NES_NO_INLINE")) g_inline = atoi(v) == 0; return true; }();
  (void)once;
  return g_inline && n_jobs == 1 && !jobs_host[0].general && jobs_host[0].n_glyphs <= INLINE_GLYPHS;
}

int launch_frame_strips(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, uint32_t *counters, uint64_t *seq, void *stream, int unit_begin, int unit_end,
                        const DevPlaced *glyphs_host) {
  int total[2] = {0, 0}, text[2] = {0, 0}, staged[2];
  launch_staged(jobs_host, n_jobs, staged);
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs_host[j];
    if (!jb.general) { total[jb.bpp - 3] += jb.n_units; text[jb.bpp - 3] += jb.n_text_segs * jb.strips_x; }
  }
  static const bool once = [] { if (const char *v = getenv("NES_NO_PDL")) g_pdl = atoi(v) == 0; return true; }();
  (void)once;
  int launches = 0;
  for (int cls = 0; cls < 2; cls++) {
    if (total[cls] == 0) continue;
    // a unit range only makes sense for a single job (its units are numbered from 0, in frame order)
    const int u0 = (n_jobs == 1 && unit_end > 0) ? unit_begin : 0, u1 = (n_jobs == 1 && unit_end > 0) ? std::min(unit_end, total[cls]) : total[cls];
    if (u1 <= u0) continue;
    const StripsConfig c = strips_config(cls, staged[cls]);
    if (c.ns < 2) return -1;
    const int grid = std::min(u1 - u0, g_num_sms * c.ctas);
    // every launch gets its own self re-arming counter pair: consecutive launches overlap (see the kernel's first line)
    uint32_t *ctr = counters + 2 * ((*seq)++ % COUNTER_SLOTS);
    cudaError_t e;
    if (glyphs_host != nullptr && frame_strips_inline_ok(jobs_host, n_jobs)) {
      static thread_local JobPack pack;  // ~9 KB: copied into the launch's parameter block by cudaLaunchKernelEx
      pack.job = jobs_host[0];
      if (jobs_host[0].n_glyphs > 0) std::memcpy(pack.glyphs, glyphs_host, sizeof(DevPlaced) * (size_t)jobs_host[0].n_glyphs);
      e = cls == 0 ? launch_one_inline<3>(grid, c.smem, (cudaStream_t)stream, pack, u0, u1, text[cls], ctr, c.ns, c.slot)
                   : launch_one_inline<4>(grid, c.smem, (cudaStream_t)stream, pack, u0, u1, text[cls], ctr, c.ns, c.slot);
    } else {
      e = cls == 0 ? launch_one<3>(grid, c.smem, (cudaStream_t)stream, jobs_dev, n_jobs, u0, u1, text[cls], ctr, c.ns, c.slot)
                   : launch_one<4>(grid, c.smem, (cudaStream_t)stream, jobs_dev, n_jobs, u0, u1, text[cls], ctr, c.ns, c.slot);
    }
    if (e != cudaSuccess) return -1;
    launches++;
  }
  return launches;
}

}  // namespace nes
