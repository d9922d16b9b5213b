// nes_internal.h -- device-visible job descriptors shared by the kernel files and session.cu.
// Not part of the public ABI (include/nes_gpu.h is).
#ifndef NES_INTERNAL_H_
#define NES_INTERNAL_H_

#include <stdint.h>

#include "nes_gpu.h"

namespace nes {

// Same-size fused kernel geometry (frame_strips.cu, k_frame_strips).  A frame is cut into
// column strips of STRIP_W pixels; a strip is cut into segments of `seg_rows` output rows
// (one work unit = one segment of one strip); a CTA walks a segment top to bottom in chunks of
// CHUNK_ROWS source rows.  Source rows travel through a ring of SUB_ROWS-row sub-stages filled
// by 2D tensor-map TMA; the pair-summed chroma rows live in a separate ring of RING_ROWS rows,
// so the 3+3 halo rows of the 8-tap vertical filter are paid once per segment, not per chunk.
constexpr int STRIP_W = 256;
constexpr int HALO = 3;
constexpr int CHUNK_ROWS = 16;     // source rows per compute chunk (= 2 sub-stages)
constexpr int SUB_ROWS = 8;        // source rows per TMA sub-stage (one row per consumer warp)
constexpr int RING_ROWS = 40;      // chroma ring: >= 6 carried rows + 2 chunks
constexpr int NS_MAX = 8;          // sub-stages in the ring (launch-wide, >= 2)
constexpr int NCTX = 8;            // chunk contexts in flight
constexpr int TMA_MAX_SOURCES = 4; // composites of more sources are filled by the consumer warps
constexpr int CONSUMER_WARPS = 8;
constexpr int CTA_THREADS = 32 * (CONSUMER_WARPS + 1);  // + one producer warp (TMA, chunk contexts)
constexpr int HIT_CAP = 256;  // glyph rect tests per pass (resize tiles)
constexpr int STRIP_HITS = 256;  // glyph rect tests per pass of an overlay chunk (k_frame_strips; the hits are staged as whole descriptors)
constexpr int MAX_SEGS = 192;   // segments of one frame (k_frame_strips work numbering)
constexpr int MASK_WORDS = 128;  // per-job bitmap: (32-row band, strip) cells touched by text
constexpr int MASK_BAND_SHIFT = 5;

// Resize kernel (resize_tiles.cu, k_resize_tiles): one CTA per destination tile of rs_tw x rs_th
// luma samples (per size pair, chosen on the host so that two CTAs fit an SM when possible).
constexpr int RS_THREADS = 256;
constexpr int RS_SMEM_MAX = 200 * 1024;   // one tile must fit this
constexpr int RS_SMEM_GOAL = 72 * 1024;   // three CTAs per SM
constexpr int GLYPH_BANDS = 256;          // placed glyphs of a resize job are bucketed by row band (counting sort on the host)

// Strip-organised resize kernel (resize_strips.cu, k_resize_strips): a work unit is one segment (rz_seg_rows
// destination rows) of one column strip (rz_dw destination columns) of one frame; its source window is at most
// RZ_BOXW pixels wide (one 4-pixel group per lane) and is walked top to bottom in chunks of RZ_CH source rows through
// the same TMA sub-stage ring as k_frame_strips.  One persistent CTA per SM with three warp roles: RZ_H_WARPS take
// source rows from packed pixels to horizontally filtered 15-bit samples in rings, RZ_V_WARPS run the vertical pass
// one chunk behind them, one producer warp issues the TMA copies; the roles meet only at mbarriers.
constexpr int RZ_BOXW = 128;
constexpr int RZ_DEPB = RZ_BOXW + 16;  // bytes of a staged depth row: its box starts on the 16-pixel boundary at or below the window origin
constexpr int RZ_SUB = 8;        // source rows per TMA sub-stage
constexpr int RZ_CH = 32;        // source rows per chunk (two rows per horizontal-pass warp)
constexpr int RZ_H_WARPS = 16, RZ_V_WARPS = 6;  // 23 warps with the producer: 6 per scheduler, 80 registers each
constexpr int RZ_THREADS = 32 * (RZ_H_WARPS + RZ_V_WARPS + 1);
constexpr int RZ_MAX_DW = 96;    // destination columns per strip (three 32-column slots per lane)
constexpr int RZ_MAX_TH = 8;     // horizontal taps kept in registers
constexpr int RZ_MAX_TV = 16;    // vertical taps
constexpr int RZ_MIN_WD = 64, RZ_MIN_HD = 32;  // smaller destinations go to k_resize_tiles

#if defined(__CUDACC__)
#define NES_HD __host__ __device__
#else
#define NES_HD
#endif
// Shared-memory layout of one resize tile (byte offsets; every region 16-byte aligned).
struct RsLayout {
  int y14, u14, v14, hy, hu, hv, dep, mask, hits, vtab, src, total;
};
// wh x ww: union source window (ww multiple of 4), cww: chroma plane row stride, nl / nc: luma /
// chroma source rows fed to the vertical pass, dwp / dcwp: padded destination widths
// dh / dch: destination rows of the tile, vls / vcs: vertical filter sizes (their rows are staged)
NES_HD inline RsLayout rs_layout(int wh, int ww, int cww, int nl, int nc, int dwp, int dcwp, int dh, int dch, int vls, int vcs, int hit_cap) {
  RsLayout L;
  int o = 0;
  auto take = [&](int bytes) { const int at = o; o += (bytes + 16 + 15) & ~15; return at; };  // +16: tap loops may over-read
  L.y14 = take(wh * ww * 2);
  L.u14 = take(wh * cww * 2);
  L.v14 = take(wh * cww * 2);
  L.hy = take(nl * dwp * 4);  // H-pass output rows are int32
  L.hu = take(nc * dcwp * 4);
  L.hv = take(nc * dcwp * 4);
  L.dep = take(wh * ww);
  L.mask = take(wh * ((ww >> 5) + 1) * 4);
  L.hits = take(hit_cap * 4 + 16);
  L.vtab = take(dh * ((4 + 2 * vls + 3) & ~3) + dch * ((4 + 2 * vcs + 3) & ~3));
  L.src = take(NES_MAX_SOURCES * 24);  // staged per-source pointers and strides
  L.total = o;
  return L;
}
// Below this source height libswscale's vertical chroma filter has fewer than 8 taps
// (initFilter clamps the size to srcH-2), so the fused same-size kernel does not apply.
constexpr int MIN_FUSED_H = 12;

struct DevSource {
  const uint8_t *rgb;
  const uint8_t *depth;
  int32_t rgb_stride;
  int32_t depth_stride;
};

// One glyph of a text run placed in the frame, already clipped to the run's view.  The atlas on the device
// holds one BIT per bitmap pixel (coverage != 0, the only thing render_text.cc:100 looks at): glyph row q is
// `wpr` 32-bit words; visible column p of visible row q is bit (bit0 + p) of the row that starts at word
// mask_off + q * wpr.
struct DevPlaced {
  int32_t x, y;       // frame position of the first visible bitmap pixel
  int32_t w, h;       // visible size
  uint32_t mask_off;  // word index of the first visible row's mask inside the atlas
  uint16_t wpr;       // mask words per bitmap row
  uint16_t bit0;      // bit index of the first visible column inside a row's mask
};

struct DevFilter {
  const int16_t *coef;  // [dst][size]
  const int32_t *pos;   // [dst]
  int32_t size;
  int32_t pad;
};

// A CUtensorMap (128 bytes, 64-byte aligned) without dragging <cuda.h> into this header.
struct alignas(64) TMap {
  uint64_t opaque[16];
};

struct alignas(64) DevJob {
  // 2D tensor maps of the staged planes (u32 elements): packed pixels and GRAY8 depth per source
  TMap tmap_px[TMA_MAX_SOURCES];
  TMap tmap_dep[TMA_MAX_SOURCES];
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
  const uint32_t *atlas;  // 1 bit per glyph pixel (see DevPlaced)
  int32_t n_glyphs;
  int32_t tma_ok;     // 16-byte aligned rows, <= TMA_MAX_SOURCES sources: rows are staged by tensor-map TMA
  int32_t use_mask;   // tile_mask valid (tiles_x*tiles_y <= 32*MASK_WORDS)
  int32_t nv12;       // chroma goes to one interleaved plane (su / du; sv / dv unused)
  uint32_t tile_mask[MASK_WORDS];  // bit (band * strips_x + strip) set: a placed glyph intersects that cell
  uint32_t band_text[8];           // bit band set: some cell of that 32-row band is (host: segment ordering)
  int32_t tiles_x, tiles_y, tile_base;  // k_resize_tiles work (general jobs only; 0 tiles otherwise)
  // k_frame_strips work (same-size jobs only).  The units of a launch are numbered in two phases: first the
  // segments that carry text, of every job (they take up to 3x longer: started first they never form the
  // tail of the launch), then all the others.  seg_order lists this job's segments in that order.
  int32_t strips_x, segs_y;
  int32_t seg_rows;    // output rows per segment (even)
  int32_t unit_base[2][2];  // [bpp-3][phase]: units of that class and phase ahead of this job in the launch
  int32_t n_units;
  int32_t n_text_segs;      // the first n_text_segs entries of seg_order are phase 0
  uint8_t seg_order[MAX_SEGS];
  // dp2a operands for packed 3-byte pixels read as raw words (frame_strips.cu phase A)
  uint32_t ky3[4], ku3[3], kv3[3];
  uint32_t ky2[2];     // luma coefficients x2, unsigned (4-byte pixels): Y lands in byte 2 of the sum
  // resize only
  DevFilter hl, hc, vl, vc;
  int32_t half;  // chroma horizontally pair-summed before the H pass
  int32_t csW;   // chroma source width fed to the H pass
  int32_t rs_smem;  // shared memory the resize kernel needs for this job's worst tile (host use)
  int32_t general;  // 1: k_resize_tiles (any size change, or H < 12 where libswscale's chroma filter is truncated)
  // placed glyphs sorted by band = clamp(y, 0, H-1) >> glyph_band_shift: band b holds glyphs
  // [glyph_band[b], glyph_band[b+1]); a tile only tests the bands its source window can touch
  int32_t glyph_band_shift, glyph_max_h;  // shift < 0: the list is short and unsorted, scan it whole
  int32_t glyph_band[GLYPH_BANDS + 1];
  int32_t rs_tw, rs_th;     // destination tile of the resize kernel
  RsLayout rs_lay;          // shared-memory carve-up for the largest tile of this size pair (same bases for every tile)
  const int32_t *rs_win_x;  // [tiles_x][4]: luma source columns [lc0, lc1), chroma source columns [cc0, cc1)
  const int32_t *rs_win_y;  // [tiles_y][4]: luma source rows [lr0, lr1), chroma source rows [cr0, cr1)
  // 16-bit depth stream (depth16.cu; host use only: these kernels take their arguments by value).  d16_src != null: the
  // job's own depth stream is off (dy == null) and this image is written by k_depth16_* after the scene kernels
  const uint8_t *d16_src;
  int32_t d16_stride;
  uint8_t *d16_y, *d16_u, *d16_v;
  int32_t d16_ys, d16_us, d16_vs;
  // k_resize_strips work (general jobs it can take: rz_ok; the others keep their tiles)
  int32_t rz_ok;
  int32_t rz_dw;                         // destination columns per strip (multiple of 16, <= RZ_MAX_DW; 0: the size pair does not fit)
  int32_t rz_strips_x, rz_seg_rows, rz_segs_y;
  int32_t rz_unit_base[2], rz_units;     // [bpp-3]: units of the rz jobs of that pixel class ahead of this job in the launch
};

// Parameter block of a single-frame launch of k_frame_strips (k_frame_strips_1): descriptor + placed glyphs by value.
constexpr int INLINE_GLYPHS = 192;
struct alignas(64) JobPack {
  DevJob job;
  DevPlaced glyphs[INLINE_GLYPHS];
};

// Launchers (frame_strips.cu, resize_tiles.cu).  jobs: device pointer to n_jobs descriptors.
// unit_end > 0 (single-job launches only): process only the units [unit_begin, unit_end) of the job
// counters: COUNTER_SLOTS pairs; *seq: the session's launch sequence number (picks the pair)
// glyphs_host != null: the caller has the job's placed glyphs on the host and did NOT upload descriptor / glyphs when
// frame_strips_inline_ok() said so -- a single same-size frame then travels in the kernel's parameter block.
int launch_frame_strips(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, uint32_t *counters, uint64_t *seq, void *stream, int unit_begin = 0, int unit_end = 0,
                        const DevPlaced *glyphs_host = nullptr);
bool frame_strips_inline_ok(const DevJob *jobs_host, int n_jobs);
// Assigns seg_rows / seg_order / unit_base of the same-size jobs of a launch (host).  text_first = false keeps the
// segments in frame order (banded submits launch unit ranges that must be row bands).
void plan_frame_strips(DevJob *jobs_host, int n_jobs, bool text_first = true);
constexpr int COUNTER_SLOTS = 64;  // work counters of k_frame_strips: one self re-arming {next unit, CTAs done} pair per launch in flight
int launch_resize_tiles(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, void *stream);
// Assigns rz_seg_rows / rz_unit_base of the jobs k_resize_strips takes (host), then the launch itself.
void plan_resize_strips(DevJob *jobs_host, int n_jobs);
int launch_resize_strips(const DevJob *jobs_dev, const DevJob *jobs_host, int n_jobs, uint32_t *counters, uint64_t *seq, void *stream);
int resize_strips_init();
// GRAY16LE depth images of the jobs that carry one (depth16.cu); returns launches or -1
int launch_depth16(const DevJob *jobs_host, int n_jobs, void *stream);
int kernels_init();  // opt-in shared memory sizes; returns cudaError_t
int frame_strips_init();

}  // namespace nes
#endif
