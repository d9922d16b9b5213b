// session.cu -- host runtime behind the C ABI (include/nes_gpu.h).
//
// A session is what one process_frame_thread of the reference owns
// (/root/reference/src/encode.cpp:42-119, one per eye, main.cpp:274-282): a CUDA device,
// three streams (H2D | kernels | D2H) and a ring of frame slots so that the upload of
// frame f+1, the kernels of frame f and the download of frame f-1 overlap.  Per frame
// the host does: pen arithmetic for the text runs (text.cc), one descriptor (DevJob),
// at most one pinned staging memcpy each way (none when the caller's buffers are pinned),
// and 1 kernel launch (frame_strips.cu or resize_tiles.cu).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "filter.h"
#include "nes_gpu.h"
#include "nes_internal.h"
#include "text.h"

using namespace nes;

namespace {

constexpr int kMaxBatch = 256;
constexpr int kBatchRing = 8;
constexpr int kMaxBands = 8;
constexpr int kBatchGlyphFactor = 8;  // placed glyphs per batched launch = this x cfg.max_glyphs

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct FilterSet {  // device tables for one (W,H,Wd,Hd)
  DevFilter hl{}, hc{}, vl{}, vc{};
  FilterTable h_hl, h_hc, h_vl, h_vc;
  void *blob = nullptr;
  int half = 0, csW = 0;
  int tw = 0, th = 0;      // destination tile of the resize kernel for this size pair
  int smem_need = 0;       // shared memory of its worst tile
  RsLayout layout{};       // carve-up for the largest tile dimensions (one layout for all tiles)
  const int32_t *win_x = nullptr, *win_y = nullptr;  // device: per tile column / row source windows
  int rz_dw[2] = {0, 0};   // [bpp-3]: destination strip width of k_resize_strips for this size pair (0: it cannot take it)
};

struct StagedCopy {  // pinned staging -> caller memory, done in wait()
  uint8_t *dst;
  const uint8_t *src;
  size_t bytes;
};

struct PlaneLayout {  // where one YUV420P image sits inside a linear buffer
  size_t off[3];
  size_t bytes[3];
  size_t total;
};

// what the download of a frame needs to know (kept with the slot: a multiplexed frame is downloaded by the dispatcher)
struct Download {
  nes_frame_out out{};
  PlaneLayout ps{}, pd{};
  bool direct = false, want_depth = false;
  int nimg = 1, nplanes = 3;
};

struct Slot {
  Download dl;
  int deferred = 0;    // staged by nes_gpu_submit, launched and downloaded by the session's mux
  int dispatched = 0;  // (guarded by the mux's done_mu) the mux has enqueued kernels + download: e_out is recorded
  int dl_status = 0;   // status of the dispatcher's part
  uint8_t *d_in = nullptr, *d_out = nullptr;
  size_t d_in_cap = 0, d_out_cap = 0;
  uint8_t *h_in = nullptr, *h_out = nullptr;
  size_t h_in_cap = 0, h_out_cap = 0;
  DevJob *h_job = nullptr, *d_job = nullptr;
  DevPlaced *h_glyphs = nullptr, *d_glyphs = nullptr;
  cudaEvent_t e_start = nullptr, e_in = nullptr, e_k0 = nullptr, e_k1 = nullptr, e_out = nullptr;
  cudaEvent_t e_band_in[kMaxBands] = {}, e_band_k[kMaxBands] = {};  // banded submits: rows uploaded / converted per band
  uint64_t ticket = 0;
  bool busy = false;
  int n_launches = 0;
  std::vector<StagedCopy> staged;
};

struct RunCacheEntry {  // one laid-out text run (place_text)
  uint64_t atlas_gen = ~0ull;
  int W = 0, H = 0, position = 0, view[4] = {0, 0, 0, 0};
  std::string text;
  std::vector<DevPlaced> placed;
};

struct BatchTables {
  DevJob *h_jobs = nullptr, *d_jobs = nullptr;
  DevPlaced *h_glyphs = nullptr, *d_glyphs = nullptr;
  cudaEvent_t done = nullptr, up = nullptr;
  bool used = false;
  int n = 0;  // frames described
  bool inline_job = false;  // a single frame launched from its kernel's parameter block (nothing was uploaded)
};

}  // namespace

struct nes_gpu_mux;
static int mux_enqueue(nes_gpu_mux *m, nes_gpu_session *s, int slot_index);
static void mux_forget(nes_gpu_mux *m, nes_gpu_session *s);
static void mux_wait_dispatched(nes_gpu_mux *m, nes_gpu_session *s, int slot_index);

struct nes_gpu_session {
  nes_gpu_cfg cfg{};
  std::mutex mu;
  cudaStream_t st_in = nullptr, st_k = nullptr, st_out = nullptr;
  // the streams the session created (st_in / st_out point at the mux's pooled copy streams while it is attached: a device
  // has a handful of hardware queues, and hundreds of streams waiting on each other's events serialise on them)
  cudaStream_t own_in = nullptr, own_out = nullptr;
  std::vector<Slot> slots;
  BatchTables batch[kBatchRing];
  uint64_t batch_seq = 0;
  uint64_t next_ticket = 1;
  uint64_t launches = 0;
  uint64_t strips_seq = 0;  // k_frame_strips launches issued (selects the work-counter pair)
  HostAtlas atlas;
  uint32_t *d_atlas = nullptr;     // glyph bit masks (HostAtlas::mask)
  uint32_t *d_counters = nullptr;  // k_frame_strips work counters: COUNTER_SLOTS self re-arming pairs, handed out round-robin per launch
  std::map<std::tuple<int, int, int, int>, FilterSet> filters;
  // tensor maps of staged planes, keyed by (pointer, stride, width in bytes, rows, box bytes): a streaming
  // session cycles through a handful of ring buffers, so encoding happens once per buffer
  std::map<std::tuple<const void *, int, int, int, int>, TMap> tmaps;
  nes_timing last{};
  std::string err;
  int sticky = 0;
  std::vector<nes_placed_glyph> scratch_placed;
  std::vector<RunCacheEntry> run_cache = std::vector<RunCacheEntry>(16);
  size_t run_cache_next = 0;
  uint64_t atlas_gen = 0;  // bumped by every atlas upload (invalidates run_cache)
  std::vector<DevPlaced> scratch_banded;
  int latency_bands = 1;  // > 1: upload / convert / download a frame in that many row bands (nes_gpu_session_set_latency_bands)
  nes_gpu_mux *mux = nullptr;  // attached: submit only stages, the mux launches the ready frames of all its sessions together
};

namespace {

#define CU_TRY(s, call)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      (s)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
      (s)->sticky = NES_ERR_CUDA;                                                            \
      return NES_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

int ensure_dev(nes_gpu_session *s, uint8_t **p, size_t *cap, size_t need) {
  if (need <= *cap) return NES_OK;
  if (*p) CU_TRY(s, cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  need = align_up(need, 1 << 20);
  {
    const cudaError_t e = cudaMalloc((void **)p, need);
    if (e == cudaErrorMemoryAllocation) {  // a transient condition of the device, not a broken session
      cudaGetLastError();
      *p = nullptr;
      s->err = "cudaMalloc: out of device memory (" + std::to_string(need) + " bytes)";
      return NES_ERR_NO_MEMORY;
    }
    CU_TRY(s, e);
  }
  CU_TRY(s, cudaMemset(*p, 0, need));
  *cap = need;
  return NES_OK;
}

int ensure_host(nes_gpu_session *s, uint8_t **p, size_t *cap, size_t need) {
  if (need <= *cap) return NES_OK;
  if (*p) CU_TRY(s, cudaFreeHost(*p));
  *p = nullptr;
  *cap = 0;
  need = align_up(need, 1 << 20);
  {
    const cudaError_t e = cudaHostAlloc((void **)p, need, cudaHostAllocDefault);
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();
      *p = nullptr;
      s->err = "cudaHostAlloc: out of pinned host memory (" + std::to_string(need) + " bytes)";
      return NES_ERR_NO_MEMORY;
    }
    CU_TRY(s, e);
  }
  *cap = need;
  return NES_OK;
}

bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int fmt_info(int fmt, int *bpp, int *base, int *a_off, bool *bgr) {
  switch (fmt) {
    case NES_PIX_RGB24: *bpp = 3; *base = 0; *a_off = -1; *bgr = false; return 0;
    case NES_PIX_BGR24: *bpp = 3; *base = 0; *a_off = -1; *bgr = true; return 0;
    case NES_PIX_RGBA: *bpp = 4; *base = 0; *a_off = 3; *bgr = false; return 0;
    case NES_PIX_BGRA: *bpp = 4; *base = 0; *a_off = 3; *bgr = true; return 0;
    case NES_PIX_ARGB: *bpp = 4; *base = 1; *a_off = 0; *bgr = false; return 0;
    case NES_PIX_ABGR: *bpp = 4; *base = 1; *a_off = 0; *bgr = true; return 0;
    default: return -1;
  }
}

// Source windows of the resize tiles for one tile size: per tile column {lc0, lc1, cc0, cc1}, per tile
// row {lr0, lr1, cr0, cr1} (filter positions are monotone except at fixed-up borders, so scan), and
// the shared memory the worst tile needs (rs_layout, the same function the kernel carves with).
int resize_windows(const FilterSet &fs, int Wd, int Hd, int tw, int th, std::vector<int32_t> *wx, std::vector<int32_t> *wy, RsLayout *lay) {
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  wx->clear(); wy->clear();
  for (int dx0 = 0; dx0 < Wd; dx0 += tw) {
    const int dx1 = std::min(dx0 + tw, Wd);
    const int cx0 = dx0 >> 1, cx1 = std::min((dx1 + 1) >> 1, cdW);
    int lc0 = 1 << 30, lc1 = 0, cc0 = 1 << 30, cc1 = 0;
    for (int i = dx0; i < dx1; i++) { lc0 = std::min(lc0, fs.h_hl.pos[i]); lc1 = std::max(lc1, fs.h_hl.pos[i] + fs.hl.size); }
    for (int i = cx0; i < cx1; i++) { cc0 = std::min(cc0, fs.h_hc.pos[i]); cc1 = std::max(cc1, fs.h_hc.pos[i] + fs.hc.size); }
    wx->insert(wx->end(), {lc0, lc1, cc0, cc1});
  }
  for (int dy0 = 0; dy0 < Hd; dy0 += th) {
    const int dy1 = std::min(dy0 + th, Hd);
    const int cy0 = dy0 >> 1, cy1 = std::min((dy1 + 1) >> 1, cdH);
    int lr0 = 1 << 30, lr1 = 0, cr0 = 1 << 30, cr1 = 0;
    for (int i = dy0; i < dy1; i++) { lr0 = std::min(lr0, fs.h_vl.pos[i]); lr1 = std::max(lr1, fs.h_vl.pos[i] + fs.vl.size); }
    for (int i = cy0; i < cy1; i++) { cr0 = std::min(cr0, fs.h_vc.pos[i]); cr1 = std::max(cr1, fs.h_vc.pos[i] + fs.vc.size); }
    wy->insert(wy->end(), {lr0, lr1, cr0, cr1});
  }
  // one carve-up for every tile: each region sized for the largest value of its dimensions over all tiles
  int m_wh = 0, m_ww = 0, m_cww = 0, m_nl = 0, m_nc = 0, m_dwp = 0, m_dcwp = 0, m_dh = 0, m_dch = 0;
  for (size_t ty = 0; ty * 4 < wy->size(); ty++)
    for (size_t tx = 0; tx * 4 < wx->size(); tx++) {
      const int32_t *X = &(*wx)[tx * 4], *Y = &(*wy)[ty * 4];
      const int dx0 = (int)tx * tw, dw = std::min(tw, Wd - dx0), dcw = std::min((dx0 + dw + 1) >> 1, cdW) - (dx0 >> 1);
      const int pc0 = fs.half ? X[2] * 2 : X[2], pc1 = fs.half ? X[3] * 2 : X[3];
      const int wx0 = std::min(X[0], pc0) & ~3, ww = ((std::max(X[1], pc1) - wx0) + 3) & ~3;
      const int wy0 = std::min(Y[0], Y[2]), wh = std::max(Y[1], Y[3]) - wy0;
      const int dy0 = (int)ty * th, dh = std::min(th, Hd - dy0), dch = std::min((dy0 + dh + 1) >> 1, cdH) - (dy0 >> 1);
      m_wh = std::max(m_wh, wh); m_ww = std::max(m_ww, ww); m_cww = std::max(m_cww, fs.half ? ww >> 1 : ww);
      m_nl = std::max(m_nl, Y[1] - Y[0]); m_nc = std::max(m_nc, Y[3] - Y[2]);
      m_dwp = std::max(m_dwp, (dw + 3) & ~3); m_dcwp = std::max(m_dcwp, (dcw + 3) & ~3);
      m_dh = std::max(m_dh, dh); m_dch = std::max(m_dch, dch);
    }
  *lay = rs_layout(m_wh, m_ww, m_cww, m_nl, m_nc, m_dwp, m_dcwp, m_dh, m_dch, fs.vl.size, fs.vc.size, HIT_CAP);
  const int worst = lay->total;
  return worst;
}

int get_filters(nes_gpu_session *s, int W, int H, int Wd, int Hd, FilterSet **out) {
  auto key = std::make_tuple(W, H, Wd, Hd);
  auto it = s->filters.find(key);
  if (it != s->filters.end()) { *out = &it->second; return NES_OK; }
  FilterSet fs;
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  fs.half = (Wd >> 1) <= (W >> 1);
  fs.csW = fs.half ? (W >> 1) : W;
  if (build_filter(W, Wd, 1 << 14, &fs.h_hl) < 0 || build_filter(fs.csW, cdW, 1 << 14, &fs.h_hc) < 0 ||
      build_filter(H, Hd, 1 << 12, &fs.h_vl) < 0 || build_filter(H, cdH, 1 << 12, &fs.h_vc) < 0)
    return NES_ERR_INVALID_ARG;
  const FilterTable *t[4] = {&fs.h_hl, &fs.h_hc, &fs.h_vl, &fs.h_vc};
  DevFilter *d[4] = {&fs.hl, &fs.hc, &fs.vl, &fs.vc};
  size_t total = 0, off[4][2];
  for (int i = 0; i < 4; i++) {
    off[i][0] = total; total = align_up(total + t[i]->coef.size() * 2 + 64, 256);  // +64: tap loops may over-read a row
    off[i][1] = total; total = align_up(total + t[i]->pos.size() * 4, 256);
  }
  std::vector<uint8_t> host(total, 0);
  for (int i = 0; i < 4; i++) {
    std::memcpy(&host[off[i][0]], t[i]->coef.data(), t[i]->coef.size() * 2);
    std::memcpy(&host[off[i][1]], t[i]->pos.data(), t[i]->pos.size() * 4);
  }
  CU_TRY(s, cudaMalloc(&fs.blob, total));
  CU_TRY(s, cudaMemcpy(fs.blob, host.data(), total, cudaMemcpyHostToDevice));
  for (int i = 0; i < 4; i++) {
    d[i]->coef = (const int16_t *)((uint8_t *)fs.blob + off[i][0]);
    d[i]->pos = (const int32_t *)((uint8_t *)fs.blob + off[i][1]);
    d[i]->size = t[i]->size;
  }
  // destination tile: the largest that leaves room for two CTAs per SM, else the largest that fits at all
  static const int kTiles[][2] = {{128, 32}, {64, 32}, {64, 16}, {32, 16}, {32, 8}, {16, 8}, {16, 4}, {8, 4}};
  std::vector<int32_t> wx, wy;
  int pick = -1;
  for (int pass = 0; pass < 2 && pick < 0; pass++)
    for (int i = 0; i < (int)(sizeof(kTiles) / sizeof(kTiles[0])); i++) {
      const int need = resize_windows(fs, Wd, Hd, kTiles[i][0], kTiles[i][1], &wx, &wy, &fs.layout);
      if (need <= (pass == 0 ? RS_SMEM_GOAL : RS_SMEM_MAX)) { pick = i; fs.smem_need = need; break; }
    }
  if (pick < 0) { cudaFree(fs.blob); return NES_ERR_TOO_LARGE; }  // scale ratio beyond what one tile can stage
  fs.tw = kTiles[pick][0]; fs.th = kTiles[pick][1];
  {
    int32_t *dw_ = nullptr;
    CU_TRY(s, cudaMalloc((void **)&dw_, (wx.size() + wy.size()) * 4 + 32));
    CU_TRY(s, cudaMemcpy(dw_, wx.data(), wx.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(s, cudaMemcpy(dw_ + wx.size(), wy.data(), wy.size() * 4, cudaMemcpyHostToDevice));
    fs.win_x = dw_; fs.win_y = dw_ + wx.size();
  }
  // k_resize_strips: the widest destination strip (multiple of 16) whose source window -- luma taps and chroma taps of
  // every column of the strip, from an origin TMA can start a box row at (4 pixels for 4-byte pixels, 16 for 3-byte
  // ones) -- fits RZ_BOXW pixels, for every strip of the frame
  fs.rz_dw[0] = fs.rz_dw[1] = 0;
  if (fs.hl.size <= RZ_MAX_TH && fs.hc.size <= RZ_MAX_TH && fs.vl.size <= RZ_MAX_TV && fs.vc.size <= RZ_MAX_TV && Wd >= RZ_MIN_WD && Hd >= RZ_MIN_HD) {
    const int taps = std::max(fs.hl.size, fs.hc.size);
    const int T = taps <= 4 ? 4 : taps <= 6 ? 6 : 8;  // taps the kernel reads per sample (zero-padded)
    for (int cls = 0; cls < 2; cls++) {
      const int amask = cls == 0 ? ~15 : ~3;
      static const int max_dw = [] { const char *v = getenv("NES_RZ_MAXDW"); const int x = v ? atoi(v) : RZ_MAX_DW; return std::min(std::max(x & ~15, 16), RZ_MAX_DW); }();
      for (int dw = max_dw; dw >= 16 && !fs.rz_dw[cls]; dw -= 16) {
        bool fits = true;
        for (int dx0 = 0; dx0 < Wd && fits; dx0 += dw) {
          const int dx1 = std::min(dx0 + dw, Wd), cx0 = dx0 >> 1, cx1 = std::min((dx1 + 1) >> 1, cdW);
          const int cmul = fs.half ? 2 : 1;
          const int wx0 = std::min(fs.h_hl.pos[dx0], cmul * fs.h_hc.pos[cx0]) & amask;
          // real taps must stay inside the window; the zero-padded ones may read the row buffers' 16-sample slack
          fits = (fs.h_hl.pos[dx1 - 1] + fs.hl.size - wx0 <= RZ_BOXW) && (cmul * (fs.h_hc.pos[cx1 - 1] + fs.hc.size) - wx0 <= RZ_BOXW) &&
                 (fs.h_hl.pos[dx1 - 1] + T - wx0 <= RZ_BOXW + 16) && (fs.h_hc.pos[cx1 - 1] + T - wx0 / cmul <= RZ_BOXW / cmul + 16);
        }
        if (fits) fs.rz_dw[cls] = dw;
      }
    }
  }
  auto ins = s->filters.emplace(key, std::move(fs));
  *out = &ins.first->second;
  return NES_OK;
}

PlaneLayout yuv_layout(const int32_t ls[3], int Hd, size_t base, bool nv12) {
  PlaneLayout p;
  const int cH = (Hd + 1) >> 1;
  p.off[0] = base; p.bytes[0] = (size_t)ls[0] * Hd;
  p.off[1] = p.off[0] + p.bytes[0]; p.bytes[1] = (size_t)ls[1] * cH;
  p.off[2] = p.off[1] + p.bytes[1]; p.bytes[2] = nv12 ? 0 : (size_t)ls[2] * cH;
  p.total = p.bytes[0] + p.bytes[1] + p.bytes[2];
  return p;
}

int validate(const nes_gpu_session *s, const nes_frame_in *in, const nes_frame_out *out, int bpp) {
  if (in->n_sources < 1 || in->n_sources > NES_MAX_SOURCES || in->n_sources > s->cfg.max_sources) return NES_ERR_INVALID_ARG;
  const int W = in->width, H = in->height, Wd = out->width, Hd = out->height;
  if (W < 4 || H < 4 || Wd < 4 || Hd < 4 || ((W | H | Wd | Hd) & 1)) return NES_ERR_INVALID_ARG;
  if (W > s->cfg.max_width || H > s->cfg.max_height || Wd > s->cfg.max_width || Hd > s->cfg.max_height) return NES_ERR_TOO_LARGE;
  if (in->mem != NES_MEM_HOST && in->mem != NES_MEM_DEVICE) return NES_ERR_INVALID_ARG;
  if (out->mem != NES_MEM_HOST && out->mem != NES_MEM_DEVICE) return NES_ERR_INVALID_ARG;
  const bool want_depth = out->depth[0] != nullptr;
  if (in->depth_fmt != NES_DEPTH_GRAY8 && in->depth_fmt != NES_DEPTH_GRAY16LE) return NES_ERR_INVALID_ARG;
  if (in->depth_fmt == NES_DEPTH_GRAY16LE && in->n_sources > 1) return NES_ERR_INVALID_ARG;  // no 16-bit depth composite
  const int dbs = in->depth_fmt == NES_DEPTH_GRAY16LE ? 2 : 1;  // bytes per depth sample
  for (int k = 0; k < in->n_sources; k++) {
    const nes_source &sr = in->src[k];
    if (!sr.rgb) return NES_ERR_INVALID_ARG;
    const int64_t rs = sr.rgb_stride ? sr.rgb_stride : W * bpp;
    if (rs < (int64_t)W * bpp) return NES_ERR_INVALID_ARG;
    if (sr.rgb_bytes < (uint64_t)(rs * (H - 1) + (int64_t)W * bpp)) return NES_ERR_SHORT_BUFFER;
    if (want_depth || in->n_sources > 1) {
      if (!sr.depth) return NES_ERR_INVALID_ARG;
      const int64_t ds = sr.depth_stride ? sr.depth_stride : (int64_t)W * dbs;
      if (ds < (int64_t)W * dbs || (dbs == 2 && (ds & 1))) return NES_ERR_INVALID_ARG;
      if (sr.depth_bytes < (uint64_t)(ds * (H - 1) + (int64_t)W * dbs)) return NES_ERR_SHORT_BUFFER;
    }
  }
  // the kernels index planes with 32-bit offsets
  if ((int64_t)W * 4 * H >= (int64_t)1 << 31 || (int64_t)out->scene_linesize[0] * Hd >= (int64_t)1 << 31 ||
      (want_depth && (int64_t)out->depth_linesize[0] * Hd >= (int64_t)1 << 31))
    return NES_ERR_TOO_LARGE;
  for (int k = 0; k < in->n_sources; k++)
    if ((int64_t)in->src[k].rgb_stride * H >= (int64_t)1 << 31 || (int64_t)in->src[k].depth_stride * H >= (int64_t)1 << 31) return NES_ERR_TOO_LARGE;
  if (out->pix_fmt != NES_OUT_YUV420P && out->pix_fmt != NES_OUT_NV12) return NES_ERR_INVALID_ARG;
  const bool nv12 = out->pix_fmt == NES_OUT_NV12;
  // line sizes: at least a row, at most a generous multiple of the session's widest row (side-by-side packing
  // passes the packed frame's line size for each eye); an absurd value would otherwise surface as an allocation failure
  const int64_t maxls = 16 * (int64_t)std::max(s->cfg.max_width, 64);
  for (int p = 0; p < (nv12 ? 2 : 3); p++) {
    const int minls = p ? (nv12 ? 2 : 1) * ((Wd + 1) / 2) : Wd;
    if (!out->scene[p] || out->scene_linesize[p] < minls) return NES_ERR_INVALID_ARG;
    if (out->scene_linesize[p] > maxls) return NES_ERR_TOO_LARGE;
    if (want_depth && (!out->depth[p] || out->depth_linesize[p] < minls)) return NES_ERR_INVALID_ARG;
    if (want_depth && out->depth_linesize[p] > maxls) return NES_ERR_TOO_LARGE;
  }
  return NES_OK;
}

bool aligned16(const void *p, int stride) { return (((uintptr_t)p | (uintptr_t)stride) & 15) == 0; }

// Fill the colour / geometry part of a job.
void job_common(DevJob *jb, const nes_frame_in *in, const nes_frame_out *out, int bpp, int base, int a_off, bool bgr) {
  std::memset(jb, 0, sizeof(*jb));
  jb->n_src = in->n_sources;
  jb->bpp = bpp;
  jb->rgb_base = base;
  jb->a_off = a_off;
  const int RY = 8414, GY = 16519, BY = 3208, RU = -4865, GU = -9528, BU = 14392, RV = 14392, GV = -12061, BV = -2332;
  const int r = bgr ? 2 : 0, b = bgr ? 0 : 2;
  jb->cy[r] = RY; jb->cy[1] = GY; jb->cy[b] = BY;
  jb->cu[r] = RU; jb->cu[1] = GU; jb->cu[b] = BU;
  jb->cv[r] = RV; jb->cv[1] = GV; jb->cv[b] = BV;
  // dp2a operands: 16-bit coefficient pairs for pixel-word bytes (0,1) and (2,3)
  int cb[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int j = 0; j < 3; j++) { cb[0][base + j] = jb->cy[j]; cb[1][base + j] = jb->cu[j]; cb[2][base + j] = jb->cv[j]; }
  uint32_t *k[3] = {jb->ky, jb->ku, jb->kv};
  for (int c = 0; c < 3; c++) {
    k[c][0] = ((uint32_t)cb[c][0] & 0xFFFFu) | ((uint32_t)cb[c][1] << 16);
    k[c][1] = ((uint32_t)cb[c][2] & 0xFFFFu) | ((uint32_t)cb[c][3] << 16);
  }
  // luma coefficients doubled and unsigned: Y comes out as byte 2 of the sum (frame_strips.cu)
  jb->ky2[0] = ((uint32_t)(2 * cb[0][0]) & 0xFFFFu) | ((uint32_t)(2 * cb[0][1]) << 16);
  jb->ky2[1] = ((uint32_t)(2 * cb[0][2]) & 0xFFFFu) | ((uint32_t)(2 * cb[0][3]) << 16);
  if (bpp == 3) {
    // 3-byte pixels are consumed as raw words: coefficient pairs per byte phase
    auto pk = [](int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); };
    const int *y = jb->cy, *u = jb->cu, *v = jb->cv;
    jb->ky3[0] = pk(2 * y[0], 2 * y[1]); jb->ky3[1] = pk(2 * y[2], 0); jb->ky3[2] = pk(0, 2 * y[0]); jb->ky3[3] = pk(2 * y[1], 2 * y[2]);
    jb->ku3[0] = pk(u[0], u[1]); jb->ku3[1] = pk(u[2], u[0]); jb->ku3[2] = pk(u[1], u[2]);
    jb->kv3[0] = pk(v[0], v[1]); jb->kv3[1] = pk(v[2], v[0]); jb->kv3[2] = pk(v[1], v[2]);
  }
  jb->W = in->width; jb->H = in->height; jb->Wd = out->width; jb->Hd = out->height;
}

// Bitmap of the (32-row band, strip) cells of the fused kernel that a placed glyph touches, so
// that every other chunk skips the overlay stage without looking at the glyph list.
void job_tile_mask(DevJob *jb, const DevPlaced *placed) {
  jb->use_mask = 0;
  std::memset(jb->band_text, 0, sizeof(jb->band_text));
  if ((jb->general && !jb->rz_ok) || jb->n_glyphs <= 0) return;
  const int strips = (jb->W + STRIP_W - 1) / STRIP_W;
  const int nbands = (jb->H + (1 << MASK_BAND_SHIFT) - 1) >> MASK_BAND_SHIFT;
  if (strips * nbands > 32 * MASK_WORDS) return;
  std::memset(jb->tile_mask, 0, sizeof(jb->tile_mask));
  for (int g = 0; g < jb->n_glyphs; g++) {
    const DevPlaced &p = placed[g];
    const int x0 = std::max(p.x, 0), x1 = std::min(p.x + p.w, jb->W), y0 = std::max(p.y, 0), y1 = std::min(p.y + p.h, jb->H);
    if (x0 >= x1 || y0 >= y1) continue;
    for (int band = y0 >> MASK_BAND_SHIFT; band <= (y1 - 1) >> MASK_BAND_SHIFT; band++)
      for (int st = x0 / STRIP_W; st <= (x1 - 1) / STRIP_W; st++) {
        const int t = band * strips + st;
        jb->tile_mask[t >> 5] |= 1u << (t & 31);
        jb->band_text[band >> 5] |= 1u << (band & 31);
      }
  }
  jb->use_mask = 1;
}

// Resize tiles of the general jobs of a launch (the same-size jobs are planned by plan_frame_strips).
void job_tiles(DevJob *jb, int tile_base) {
  jb->tiles_x = jb->tiles_y = 0;
  if (jb->general && !jb->rz_ok) {
    jb->tiles_x = (jb->Wd + jb->rs_tw - 1) / jb->rs_tw;
    jb->tiles_y = (jb->Hd + jb->rs_th - 1) / jb->rs_th;
  }
  jb->tile_base = tile_base;
}

// 16-bit depth: the scene kernels run without a depth stream, k_depth16_* (depth16.cu) writes the depth image.
void job_depth16(DevJob *jb) {
  jb->d16_src = jb->src[0].depth; jb->d16_stride = jb->src[0].depth_stride;
  jb->d16_y = jb->dy; jb->d16_u = jb->du; jb->d16_v = jb->dv;
  jb->d16_ys = jb->dys; jb->d16_us = jb->dus; jb->d16_vs = jb->dvs;
  jb->dy = jb->du = jb->dv = nullptr;
  jb->src[0].depth = nullptr;
}

void job_alignment(DevJob *jb) {
  bool iv = true;
  for (int k = 0; k < jb->n_src; k++) {
    iv = iv && aligned16(jb->src[k].rgb, jb->src[k].rgb_stride);
    if (jb->src[k].depth) iv = iv && aligned16(jb->src[k].depth, jb->src[k].depth_stride);
  }
  bool ov = aligned16(jb->sy, jb->sys) && aligned16(jb->su, jb->sus) && (jb->nv12 || aligned16(jb->sv, jb->svs));
  if (jb->dy) ov = ov && aligned16(jb->dy, jb->dys) && aligned16(jb->du, jb->dus) && (jb->nv12 || aligned16(jb->dv, jb->dvs));
  jb->in_vec = iv;
  jb->out_vec = ov;
  // rows staged by tensor-map TMA: 16-byte aligned rows; composites only for 4-byte pixels
  jb->tma_ok = iv && (jb->W % 16) == 0 && jb->n_src <= TMA_MAX_SOURCES && (jb->n_src == 1 || jb->bpp == 4);
  // k_resize_strips: the size pair must fit its limits (rz_dw, get_filters) and the planes must suit TMA / word stores
  jb->rz_ok = jb->general && jb->rz_dw > 0 && iv && ov && (jb->W % 4) == 0 && jb->n_src <= TMA_MAX_SOURCES && (jb->n_src == 1 || jb->bpp == 4);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// 2D tensor map over a plane of `rows` rows of `width_bytes` bytes (u32 elements), box =
// SUB_ROWS rows x one strip (box_bytes).  Rows / columns outside the plane read as zero.
int plane_tmap(nes_gpu_session *s, const uint8_t *base, int stride, int width_bytes, int rows, int box_bytes, TMap *out) {
  static_assert(sizeof(TMap) == sizeof(CUtensorMap) && alignof(TMap) >= alignof(CUtensorMap), "TMap mirrors CUtensorMap");
  const auto key = std::make_tuple((const void *)base, stride, width_bytes, rows, box_bytes);
  auto it = s->tmaps.find(key);
  if (it != s->tmaps.end()) { *out = it->second; return NES_OK; }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { s->err = "cuTensorMapEncodeTiled unavailable"; return NES_ERR_CUDA; }
  CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)(width_bytes / 4), (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)stride};
  const cuuint32_t box[2] = {(cuuint32_t)(box_bytes / 4), (cuuint32_t)SUB_ROWS};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { s->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return NES_ERR_CUDA; }
  if (s->tmaps.size() > 4096) s->tmaps.clear();
  TMap t;
  std::memcpy(&t, &m, sizeof(t));
  s->tmaps.emplace(key, t);
  *out = t;
  return NES_OK;
}

// Tensor maps of every plane the fused kernel stages for this job.
int job_tmaps(nes_gpu_session *s, DevJob *jb) {
  if (jb->general ? !jb->rz_ok : !jb->tma_ok) return NES_OK;
  const bool dep = jb->dy != nullptr || jb->n_src > 1;
  const int boxw = jb->general ? RZ_BOXW : STRIP_W;  // pixels per staged row
  for (int k = 0; k < jb->n_src; k++) {
    int st = plane_tmap(s, jb->src[k].rgb, jb->src[k].rgb_stride, jb->W * jb->bpp, jb->H, boxw * jb->bpp, &jb->tmap_px[k]);
    if (st) return st;
    if (dep && (st = plane_tmap(s, jb->src[k].depth, jb->src[k].depth_stride, jb->W, jb->H, jb->general ? RZ_DEPB : boxw, &jb->tmap_dep[k]))) return st;
  }
  return NES_OK;
}

// Text runs -> placed glyph descriptors at dst[0..].  Returns count or negative status.
// A streaming session repeats most of its strings from frame to frame (the camera matrix, "direction=..."): the
// placed form of the last few distinct runs is kept and copied instead of laid out again.
int place_text(nes_gpu_session *s, int W, int H, const nes_text_run *runs, int n_runs, DevPlaced *dst, int cap) {
  if (n_runs <= 0) return 0;
  if (!runs) return NES_ERR_INVALID_ARG;
  if (!s->atlas.valid) return NES_ERR_NO_ATLAS;
  int n = 0;
  for (int r = 0; r < n_runs; r++) {
    const nes_text_run &run = runs[r];
    if (run.len > 0 && !run.text) return NES_ERR_INVALID_ARG;
    if (run.len <= 0) continue;
    RunCacheEntry *hit = nullptr;
    for (RunCacheEntry &e : s->run_cache)
      if (e.atlas_gen == s->atlas_gen && e.W == W && e.H == H && e.position == run.position && e.view[0] == run.view_x && e.view[1] == run.view_y &&
          e.view[2] == run.view_w && e.view[3] == run.view_h && (int)e.text.size() == run.len && std::memcmp(e.text.data(), run.text, (size_t)run.len) == 0) {
        hit = &e;
        break;
      }
    if (!hit) {
      s->scratch_placed.clear();
      layout_run(s->atlas, W, H, run, &s->scratch_placed);
      hit = &s->run_cache[s->run_cache_next++ % s->run_cache.size()];
      hit->atlas_gen = s->atlas_gen; hit->W = W; hit->H = H; hit->position = run.position;
      hit->view[0] = run.view_x; hit->view[1] = run.view_y; hit->view[2] = run.view_w; hit->view[3] = run.view_h;
      hit->text.assign(run.text, (size_t)run.len);
      hit->placed.clear();
      for (const nes_placed_glyph &pg : s->scratch_placed) {
        const HostGlyph &g = s->atlas.glyph[pg.code];
        // pre-clipped to the run's view (hence to the frame): the device only clips to its own tile / chunk
        hit->placed.push_back(DevPlaced{pg.x + pg.clip_x, pg.y + pg.clip_y, pg.clip_w, pg.clip_h, g.mask_off + (uint32_t)(pg.clip_y * g.wpr), (uint16_t)g.wpr, (uint16_t)pg.clip_x});
      }
    }
    const int m = (int)hit->placed.size();
    if (n + m > cap) return NES_ERR_TOO_LARGE;
    if (m) std::memcpy(dst + n, hit->placed.data(), (size_t)m * sizeof(DevPlaced));
    n += m;
  }
  return n;
}

// Bucket the placed glyphs by row band (stable counting sort; stamps are order-free) so that a tile / chunk
// tests only the glyphs near its source rows instead of the whole list.
void band_glyphs(DevJob *jb, DevPlaced *gl, int n, std::vector<DevPlaced> *tmp) {
  // a short list is scanned whole by an overlay chunk of the fused kernel (one pass of STRIP_HITS tests): no sort
  if (!jb->general && n <= STRIP_HITS) { jb->glyph_band_shift = -1; return; }
  int shift = 5;
  while (((jb->H - 1) >> shift) >= GLYPH_BANDS) shift++;
  jb->glyph_band_shift = shift;
  jb->glyph_max_h = 0;
  int count[GLYPH_BANDS + 1] = {0};
  auto band_of = [&](const DevPlaced &p) { return std::min(std::max(p.y, 0), jb->H - 1) >> shift; };
  for (int i = 0; i < n; i++) { count[band_of(gl[i]) + 1]++; jb->glyph_max_h = std::max(jb->glyph_max_h, gl[i].h); }
  for (int b = 0; b < GLYPH_BANDS; b++) count[b + 1] += count[b];
  for (int b = 0; b <= GLYPH_BANDS; b++) jb->glyph_band[b] = count[b];
  tmp->assign(gl, gl + n);
  for (int i = 0; i < n; i++) gl[count[band_of((*tmp)[i])]++] = (*tmp)[i];
}

// D2H of one converted frame on the session's download stream (which already waits for the kernels): straight into the
// caller's planes when they are pinned and laid out like av_image_alloc, else through pinned staging (copied out in wait).
int enqueue_download(nes_gpu_session *s, Slot &sl) {
  const Download &d = sl.dl;
  if (d.out.mem != NES_MEM_HOST) return NES_OK;
  const size_t total = d.pd.off[0] + (d.want_depth ? d.pd.total : 0);
  if (d.direct) {
    CU_TRY(s, cudaMemcpyAsync(d.out.scene[0], sl.d_out + d.ps.off[0], d.ps.total, cudaMemcpyDeviceToHost, s->st_out));
    if (d.want_depth) CU_TRY(s, cudaMemcpyAsync(d.out.depth[0], sl.d_out + d.pd.off[0], d.pd.total, cudaMemcpyDeviceToHost, s->st_out));
  } else {
    int st;
    if ((st = ensure_host(s, &sl.h_out, &sl.h_out_cap, total))) return st;
    CU_TRY(s, cudaMemcpyAsync(sl.h_out, sl.d_out, total, cudaMemcpyDeviceToHost, s->st_out));
    for (int im = 0; im < d.nimg; im++) {
      uint8_t *const *pl = im ? d.out.depth : d.out.scene;
      const PlaneLayout &L = im ? d.pd : d.ps;
      for (int p = 0; p < d.nplanes; p++) sl.staged.push_back(StagedCopy{pl[p], sl.h_out + L.off[p], L.bytes[p]});
    }
  }
  return NES_OK;
}

int run_kernels(nes_gpu_session *s, const DevJob *d_jobs, const DevJob *h_jobs, int n, cudaStream_t st, const DevPlaced *glyphs_host = nullptr) {
  int l = 0;
  const int r0 = launch_frame_strips(d_jobs, h_jobs, n, s->d_counters, &s->strips_seq, st, 0, 0, glyphs_host);
  if (r0 > 0) l += r0;
  const int r1 = launch_resize_strips(d_jobs, h_jobs, n, s->d_counters, &s->strips_seq, st);
  if (r1 > 0) l += r1;
  const int r = launch_resize_tiles(d_jobs, h_jobs, n, st);
  if (r > 0) l += r;
  const int r2 = launch_depth16(h_jobs, n, st);
  if (r2 > 0) l += r2;
  if (r0 < 0 || r1 < 0 || r < 0 || r2 < 0) {  // a launch failed: the frame was not converted
    if (s->err.empty()) s->err = "kernel launch failed";
    s->sticky = NES_ERR_CUDA;
  }
  s->launches += (uint64_t)l;
  return l;
}

}  // namespace

// ============================================================================
extern "C" {

int nes_gpu_abi_version(void) { return NES_ABI_VERSION; }

const char *nes_gpu_strerror(int st) {
  switch (st) {
    case NES_OK: return "ok";
    case NES_ERR_INVALID_ARG: return "invalid argument";
    case NES_ERR_SHORT_BUFFER: return "source buffer shorter than stride*height";
    case NES_ERR_TOO_LARGE: return "frame or glyph list exceeds the session limits";
    case NES_ERR_CUDA: return "CUDA error (see nes_gpu_session_error)";
    case NES_ERR_NO_MEMORY: return "out of memory";
    case NES_ERR_BAD_TICKET: return "unknown ticket";
    case NES_ERR_PARSE: return "malformed RenderedFrame message";
    case NES_ERR_NO_ATLAS: return "text submitted before a glyph atlas was set";
    case NES_ERR_FREETYPE: return "FreeType unavailable or font could not be opened";
    case NES_ERR_BUSY: return "frame ring full; wait on an older ticket";
    case NES_ERR_UNSUPPORTED: return "optional run-time dependency missing or of an unknown version";
    default: return "unknown status";
  }
}

const char *nes_gpu_session_error(nes_gpu_session *s) { return s ? s->err.c_str() : ""; }

int nes_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int nes_gpu_session_create(const nes_gpu_cfg *cfg, nes_gpu_session **out) {
  if (!cfg || !out) return NES_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->max_width < 4 || cfg->max_height < 4) return NES_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return NES_ERR_CUDA; }
  if (cfg->device < 0 || cfg->device >= ndev) return NES_ERR_INVALID_ARG;
  nes_gpu_session *s = new (std::nothrow) nes_gpu_session();
  if (!s) return NES_ERR_NO_MEMORY;
  s->cfg = *cfg;
  if (s->cfg.max_sources < 1) s->cfg.max_sources = 1;
  if (s->cfg.max_sources > NES_MAX_SOURCES) s->cfg.max_sources = NES_MAX_SOURCES;
  if (s->cfg.ring_depth < 1) s->cfg.ring_depth = 3;
  if (s->cfg.max_glyphs < 1) s->cfg.max_glyphs = 8192;
  auto fail = [&](int st) { nes_gpu_session_destroy(s); return st; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(NES_ERR_CUDA);
  if (kernels_init() != 0) return fail(NES_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&s->own_in, cudaStreamNonBlocking) != cudaSuccess) return fail(NES_ERR_CUDA);
  s->st_in = s->own_in;
  if (cudaStreamCreateWithFlags(&s->st_k, cudaStreamNonBlocking) != cudaSuccess) return fail(NES_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&s->own_out, cudaStreamNonBlocking) != cudaSuccess) return fail(NES_ERR_CUDA);
  s->st_out = s->own_out;
  if (cudaMalloc((void **)&s->d_counters, 2 * COUNTER_SLOTS * sizeof(uint32_t)) != cudaSuccess) return fail(NES_ERR_CUDA);
  if (cudaMemset(s->d_counters, 0, 2 * COUNTER_SLOTS * sizeof(uint32_t)) != cudaSuccess) return fail(NES_ERR_CUDA);
  s->slots.resize((size_t)s->cfg.ring_depth);
  const size_t gl_bytes = (size_t)s->cfg.max_glyphs * sizeof(DevPlaced);
  for (Slot &sl : s->slots) {
    cudaEvent_t *ev[5] = {&sl.e_start, &sl.e_in, &sl.e_k0, &sl.e_k1, &sl.e_out};
    for (cudaEvent_t *e : ev)
      if (cudaEventCreate(e) != cudaSuccess) return fail(NES_ERR_CUDA);
    for (int b = 0; b < kMaxBands; b++)
      if (cudaEventCreateWithFlags(&sl.e_band_in[b], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&sl.e_band_k[b], cudaEventDisableTiming) != cudaSuccess)
        return fail(NES_ERR_CUDA);
    if (cudaHostAlloc((void **)&sl.h_job, sizeof(DevJob), cudaHostAllocDefault) != cudaSuccess) return fail(NES_ERR_CUDA);
    if (cudaMalloc((void **)&sl.d_job, sizeof(DevJob)) != cudaSuccess) return fail(NES_ERR_CUDA);
    if (cudaHostAlloc((void **)&sl.h_glyphs, gl_bytes, cudaHostAllocDefault) != cudaSuccess) return fail(NES_ERR_CUDA);
    if (cudaMalloc((void **)&sl.d_glyphs, gl_bytes) != cudaSuccess) return fail(NES_ERR_CUDA);
  }
  *out = s;
  return NES_OK;
}

void nes_gpu_session_destroy(nes_gpu_session *s) {
  if (!s) return;
  cudaSetDevice(s->cfg.device);
  if (s->st_in) cudaStreamSynchronize(s->st_in);
  if (s->st_k) cudaStreamSynchronize(s->st_k);
  if (s->st_out) cudaStreamSynchronize(s->st_out);
  if (s->mux) mux_forget(s->mux, s);  // back on its own streams, off the mux's list
  if (s->st_in) cudaStreamSynchronize(s->st_in);
  if (s->st_out) cudaStreamSynchronize(s->st_out);
  for (Slot &sl : s->slots) {
    cudaFree(sl.d_in); cudaFree(sl.d_out);
    cudaFreeHost(sl.h_in); cudaFreeHost(sl.h_out);
    cudaFreeHost(sl.h_job); cudaFree(sl.d_job);
    cudaFreeHost(sl.h_glyphs); cudaFree(sl.d_glyphs);
    cudaEvent_t ev[5] = {sl.e_start, sl.e_in, sl.e_k0, sl.e_k1, sl.e_out};
    for (cudaEvent_t e : ev)
      if (e) cudaEventDestroy(e);
    for (int b = 0; b < kMaxBands; b++) {
      if (sl.e_band_in[b]) cudaEventDestroy(sl.e_band_in[b]);
      if (sl.e_band_k[b]) cudaEventDestroy(sl.e_band_k[b]);
    }
  }
  for (BatchTables &b : s->batch) {
    cudaFreeHost(b.h_jobs); cudaFree(b.d_jobs); cudaFreeHost(b.h_glyphs); cudaFree(b.d_glyphs);
    if (b.done) cudaEventDestroy(b.done);
    if (b.up) cudaEventDestroy(b.up);
  }
  for (auto &kv : s->filters) { cudaFree(kv.second.blob); cudaFree((void *)kv.second.win_x); }
  cudaFree(s->d_atlas);
  cudaFree(s->d_counters);
  if (s->own_in) cudaStreamDestroy(s->own_in);
  if (s->st_k) cudaStreamDestroy(s->st_k);
  if (s->own_out) cudaStreamDestroy(s->own_out);
  cudaGetLastError();
  delete s;
}

void *nes_gpu_session_stream(nes_gpu_session *s) { return s ? (void *)s->st_k : nullptr; }

int nes_gpu_session_set_latency_bands(nes_gpu_session *s, int bands) {
  if (!s || bands < 1 || bands > kMaxBands) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  s->latency_bands = bands;
  return NES_OK;
}
uint64_t nes_gpu_session_launches(nes_gpu_session *s) { return s ? s->launches : 0; }

int nes_gpu_host_alloc(size_t bytes, void **out) {
  if (!out || bytes == 0) return NES_ERR_INVALID_ARG;
  if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); *out = nullptr; return NES_ERR_CUDA; }
  return NES_OK;
}
void nes_gpu_host_free(void *p) {
  if (p) { cudaFreeHost(p); cudaGetLastError(); }
}

int nes_gpu_device_alloc(nes_gpu_session *s, size_t bytes, void **out) {
  if (!s || !out || bytes == 0) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  CU_TRY(s, cudaMalloc(out, bytes));
  return NES_OK;
}
void nes_gpu_device_free(nes_gpu_session *s, void *p) {
  if (!s || !p) return;
  std::lock_guard<std::mutex> lk(s->mu);
  cudaSetDevice(s->cfg.device);
  cudaFree(p);
  cudaGetLastError();
}
int nes_gpu_memcpy_h2d(nes_gpu_session *s, void *dst, const void *src, size_t bytes) {
  if (!s || !dst || !src) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  CU_TRY(s, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return NES_OK;
}
int nes_gpu_memcpy_d2h(nes_gpu_session *s, void *dst, const void *src, size_t bytes) {
  if (!s || !dst || !src) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  CU_TRY(s, cudaStreamSynchronize(s->st_k));
  CU_TRY(s, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return NES_OK;
}

// ---- atlas ------------------------------------------------------------------
static int upload_atlas(nes_gpu_session *s) {
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  CU_TRY(s, cudaStreamSynchronize(s->st_k));
  if (s->d_atlas) { CU_TRY(s, cudaFree(s->d_atlas)); s->d_atlas = nullptr; }
  build_masks(&s->atlas);
  s->atlas_gen++;
  const size_t n = std::max<size_t>(s->atlas.mask.size(), 4) * sizeof(uint32_t);
  CU_TRY(s, cudaMalloc((void **)&s->d_atlas, n + 64));  // + slack: a row's last word may be read as part of a wider load
  CU_TRY(s, cudaMemset(s->d_atlas, 0, n + 64));
  if (!s->atlas.mask.empty())
    CU_TRY(s, cudaMemcpy(s->d_atlas, s->atlas.mask.data(), s->atlas.mask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  s->atlas.valid = true;
  return NES_OK;
}

int nes_gpu_atlas_set(nes_gpu_session *s, const nes_glyph *glyphs, int n) {
  if (!s || (n > 0 && !glyphs) || n < 0) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  HostAtlas a;
  for (int i = 0; i < n; i++) {
    const nes_glyph &g = glyphs[i];
    if (g.code < 0 || g.code > 255 || g.width < 0 || g.rows < 0) return NES_ERR_INVALID_ARG;
    if (g.width > 0 && g.rows > 0 && (!g.coverage || g.pitch < g.width)) return NES_ERR_INVALID_ARG;
    HostGlyph &h = a.glyph[g.code];
    h.width = g.width; h.rows = g.rows; h.left = g.left; h.top = g.top; h.advance = g.advance;
    h.pitch = g.width;
    h.offset = (uint32_t)a.coverage.size();
    for (int q = 0; q < g.rows; q++) a.coverage.insert(a.coverage.end(), g.coverage + (size_t)q * g.pitch, g.coverage + (size_t)q * g.pitch + g.width);
  }
  s->atlas = std::move(a);
  return upload_atlas(s);
}

int nes_gpu_atlas_load_font(nes_gpu_session *s, const char *freetype_so, const char *font_path) {
  if (!s || !font_path) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  HostAtlas a;
  const int st = rasterise_font(freetype_so, font_path, &a, &s->err);
  if (st != NES_OK) return st;
  s->atlas = std::move(a);
  return upload_atlas(s);
}

int nes_font_rasterise(const char *freetype_so, const char *font_path, nes_glyph *glyphs, uint8_t *coverage,
                       uint64_t coverage_cap, uint64_t *coverage_used) {
  if (!font_path || !glyphs) return NES_ERR_INVALID_ARG;
  HostAtlas a;
  std::string err;
  const int st = rasterise_font(freetype_so, font_path, &a, &err);
  if (st != NES_OK) return st;
  if (coverage_used) *coverage_used = a.coverage.size();
  if (a.coverage.size() > coverage_cap || (!coverage && !a.coverage.empty())) return NES_ERR_TOO_LARGE;
  if (!a.coverage.empty()) std::memcpy(coverage, a.coverage.data(), a.coverage.size());
  for (int c = 0; c < 256; c++) {
    const HostGlyph &g = a.glyph[c];
    glyphs[c] = nes_glyph{c, g.width, g.rows, g.left, g.top, g.advance, g.width, 0, (g.width > 0 && g.rows > 0) ? coverage + g.offset : nullptr};
  }
  return NES_OK;
}

int nes_gpu_text_layout(nes_gpu_session *s, int frame_w, int frame_h, const nes_text_run *run, nes_placed_glyph *out, int cap) {
  if (!s || !run || (cap > 0 && !out)) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  if (!s->atlas.valid) return NES_ERR_NO_ATLAS;
  std::vector<nes_placed_glyph> v;
  layout_run(s->atlas, frame_w, frame_h, *run, &v);
  for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
  return (int)v.size();
}

int nes_gpu_filter_table(int src, int dst, int one, int16_t *coef, int coef_cap, int32_t *pos, int pos_cap) {
  FilterTable t;
  const int size = build_filter(src, dst, one, &t);
  if (size < 0) return NES_ERR_INVALID_ARG;
  if (coef) {
    if ((size_t)coef_cap < t.coef.size()) return NES_ERR_TOO_LARGE;
    std::memcpy(coef, t.coef.data(), t.coef.size() * 2);
  }
  if (pos) {
    if ((size_t)pos_cap < t.pos.size()) return NES_ERR_TOO_LARGE;
    std::memcpy(pos, t.pos.data(), t.pos.size() * 4);
  }
  return size;
}

// ---- the hot path -------------------------------------------------------------
int nes_gpu_submit(nes_gpu_session *s, const nes_frame_in *in, const nes_text_run *runs, int n_runs,
                   const nes_frame_out *out, uint64_t *ticket) {
  if (!s || !in || !out || !ticket) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->sticky) return s->sticky;
  int bpp, base, a_off; bool bgr;
  if (fmt_info(in->pix_fmt, &bpp, &base, &a_off, &bgr)) return NES_ERR_INVALID_ARG;
  int st = validate(s, in, out, bpp);
  if (st) return st;
  CU_TRY(s, cudaSetDevice(s->cfg.device));

  Slot &sl = s->slots[s->next_ticket % s->slots.size()];
  if (sl.busy) return NES_ERR_BUSY;
  const int W = in->width, H = in->height, Wd = out->width, Hd = out->height;
  const bool want_depth = out->depth[0] != nullptr;
  const bool need_depth_in = want_depth || in->n_sources > 1;
  const bool resize = (W != Wd) || (H != Hd) || H < MIN_FUSED_H;
  const size_t dbs = in->depth_fmt == NES_DEPTH_GRAY16LE ? 2 : 1;  // bytes per depth sample

  // text -> placed glyphs (pinned)
  const int n_gl = place_text(s, W, H, runs, n_runs, sl.h_glyphs, s->cfg.max_glyphs);
  if (n_gl < 0) return n_gl;

  DevJob *jb = sl.h_job;
  job_common(jb, in, out, bpp, base, a_off, bgr);
  jb->glyphs = sl.d_glyphs;
  jb->atlas = s->d_atlas;
  jb->n_glyphs = n_gl;

  sl.staged.clear();
  CU_TRY(s, cudaEventRecord(sl.e_start, s->st_in));
  bool banded = false, pinned_upload = false;
  struct PinnedUpload { const uint8_t *rgb, *depth; uint8_t *d_rgb, *d_depth; size_t rs, ds; } up[NES_MAX_SOURCES] = {};

  // ---- inputs ----
  if (in->mem == NES_MEM_DEVICE) {
    for (int k = 0; k < in->n_sources; k++) {
      jb->src[k].rgb = in->src[k].rgb;
      jb->src[k].rgb_stride = in->src[k].rgb_stride ? in->src[k].rgb_stride : W * bpp;
      jb->src[k].depth = need_depth_in ? in->src[k].depth : nullptr;
      jb->src[k].depth_stride = in->src[k].depth_stride ? in->src[k].depth_stride : (int)(W * dbs);
    }
  } else {
    const size_t rs_dev = align_up((size_t)W * bpp, 16), ds_dev = align_up((size_t)W * dbs, 16);
    const size_t rgb_sz = align_up(rs_dev * H, 256), dep_sz = need_depth_in ? align_up(ds_dev * H, 256) : 0;
    const size_t need = (rgb_sz + dep_sz) * in->n_sources;
    if ((st = ensure_dev(s, &sl.d_in, &sl.d_in_cap, need))) return st;
    bool all_pinned = true;
    for (int k = 0; k < in->n_sources; k++) {
      all_pinned = all_pinned && is_pinned(in->src[k].rgb);
      if (need_depth_in) all_pinned = all_pinned && is_pinned(in->src[k].depth);
    }
    if (!all_pinned && (st = ensure_host(s, &sl.h_in, &sl.h_in_cap, need))) return st;
    // banded low-latency submit: pinned host buffers both ways, same-size path; the uploads are issued band by
    // band further down, interleaved with the launches
    banded = s->latency_bands > 1 && all_pinned && !resize && out->mem == NES_MEM_HOST && !s->mux;
    for (int k = 0; k < in->n_sources; k++) {
      const nes_source &sr = in->src[k];
      const size_t o_rgb = (rgb_sz + dep_sz) * k, o_dep = o_rgb + rgb_sz;
      const size_t rs = sr.rgb_stride ? sr.rgb_stride : (size_t)W * bpp;
      const size_t ds = sr.depth_stride ? sr.depth_stride : (size_t)W * dbs;
      jb->src[k].rgb = sl.d_in + o_rgb;
      jb->src[k].rgb_stride = (int)rs_dev;
      jb->src[k].depth = need_depth_in ? sl.d_in + o_dep : nullptr;
      jb->src[k].depth_stride = (int)ds_dev;
      if (all_pinned) {
        // issued below (whole frame, or band by band): remember where the rows are
        up[k] = PinnedUpload{sr.rgb, need_depth_in ? sr.depth : nullptr, sl.d_in + o_rgb, sl.d_in + o_dep, rs, ds};
        pinned_upload = true;
      } else {
        // the one host memcpy of the payload: caller memory -> pinned staging
        if (rs == rs_dev) std::memcpy(sl.h_in + o_rgb, sr.rgb, rs * (H - 1) + (size_t)W * bpp);
        else for (int y = 0; y < H; y++) std::memcpy(sl.h_in + o_rgb + y * rs_dev, sr.rgb + y * rs, (size_t)W * bpp);
        if (need_depth_in) {
          if (ds == ds_dev) std::memcpy(sl.h_in + o_dep, sr.depth, ds * (H - 1) + W * dbs);
          else for (int y = 0; y < H; y++) std::memcpy(sl.h_in + o_dep + y * ds_dev, sr.depth + y * ds, W * dbs);
        }
      }
    }
    if (!all_pinned) CU_TRY(s, cudaMemcpyAsync(sl.d_in, sl.h_in, need, cudaMemcpyHostToDevice, s->st_in));
  }

  // ---- outputs ----
  PlaneLayout ps{}, pd{};
  if (out->mem == NES_MEM_DEVICE) {
    jb->sy = out->scene[0]; jb->su = out->scene[1]; jb->sv = out->scene[2];
    if (want_depth) { jb->dy = out->depth[0]; jb->du = out->depth[1]; jb->dv = out->depth[2]; }
  } else {
    ps = yuv_layout(out->scene_linesize, Hd, 0, out->pix_fmt == NES_OUT_NV12);
    pd = yuv_layout(out->depth_linesize, Hd, align_up(ps.total, 256), out->pix_fmt == NES_OUT_NV12);
    const size_t need = pd.off[0] + (want_depth ? pd.total : 0);
    if ((st = ensure_dev(s, &sl.d_out, &sl.d_out_cap, need))) return st;
    jb->sy = sl.d_out + ps.off[0]; jb->su = sl.d_out + ps.off[1]; jb->sv = sl.d_out + ps.off[2];
    if (want_depth) { jb->dy = sl.d_out + pd.off[0]; jb->du = sl.d_out + pd.off[1]; jb->dv = sl.d_out + pd.off[2]; }
  }
  jb->sys = out->scene_linesize[0]; jb->sus = out->scene_linesize[1]; jb->svs = out->scene_linesize[2];
  jb->dys = out->depth_linesize[0]; jb->dus = out->depth_linesize[1]; jb->dvs = out->depth_linesize[2];
  jb->nv12 = out->pix_fmt == NES_OUT_NV12;
  if (in->depth_fmt == NES_DEPTH_GRAY16LE && want_depth) job_depth16(jb);
  // (a banded submit launches unit ranges of one kernel: a 16-bit depth frame goes through whole)
  if (jb->d16_src) banded = false;

  jb->general = resize;
  if (n_gl > 0) band_glyphs(jb, sl.h_glyphs, n_gl, &s->scratch_banded);
  if (resize) {
    FilterSet *fs;
    if ((st = get_filters(s, W, H, Wd, Hd, &fs))) return st;
    jb->hl = fs->hl; jb->hc = fs->hc; jb->vl = fs->vl; jb->vc = fs->vc;
    jb->half = fs->half; jb->csW = fs->csW; jb->rs_smem = fs->smem_need;
    jb->rs_tw = fs->tw; jb->rs_th = fs->th; jb->rs_win_x = fs->win_x; jb->rs_win_y = fs->win_y; jb->rs_lay = fs->layout;
    jb->rz_dw = getenv("NES_NO_RZ") ? 0 : fs->rz_dw[bpp - 3];
  }
  job_alignment(jb);
  job_tiles(jb, 0);
  if ((st = job_tmaps(s, jb))) return st;
  job_tile_mask(jb, sl.h_glyphs);
  // (banded submits launch unit ranges that must be whole row bands: segments stay in frame order)
  plan_frame_strips(jb, 1, /*text_first=*/!banded);
  plan_resize_strips(jb, 1);

  // a single same-size frame travels in its kernel's parameter block (k_frame_strips_1): no descriptor / glyph upload.
  // (a multiplexed session's frame is launched from the mux's table, which points at the slot's device glyph list)
  const bool inline_job = !s->mux && frame_strips_inline_ok(jb, 1);
  if (!inline_job) {
    if (n_gl > 0) CU_TRY(s, cudaMemcpyAsync(sl.d_glyphs, sl.h_glyphs, (size_t)n_gl * sizeof(DevPlaced), cudaMemcpyHostToDevice, s->st_in));
    if (!s->mux) CU_TRY(s, cudaMemcpyAsync(sl.d_job, sl.h_job, sizeof(DevJob), cudaMemcpyHostToDevice, s->st_in));
  }

  // rows [y0, y1) of every (pinned) source -> the device copy
  const size_t rs_dev_ = align_up((size_t)W * bpp, 16), ds_dev_ = align_up((size_t)W * dbs, 16);
  auto upload_rows = [&](int y0, int y1) -> int {
    if (y1 <= y0) return NES_OK;
    for (int k = 0; k < in->n_sources; k++) {
      const PinnedUpload &u = up[k];
      const size_t rows = (size_t)(y1 - y0);
      if (u.rs == rs_dev_) CU_TRY(s, cudaMemcpyAsync(u.d_rgb + y0 * rs_dev_, u.rgb + y0 * u.rs, y1 == H ? (rows - 1) * u.rs + (size_t)W * bpp : rows * u.rs, cudaMemcpyHostToDevice, s->st_in));
      else CU_TRY(s, cudaMemcpy2DAsync(u.d_rgb + y0 * rs_dev_, rs_dev_, u.rgb + y0 * u.rs, u.rs, (size_t)W * bpp, rows, cudaMemcpyHostToDevice, s->st_in));
      if (u.depth) {
        if (u.ds == ds_dev_) CU_TRY(s, cudaMemcpyAsync(u.d_depth + y0 * ds_dev_, u.depth + y0 * u.ds, y1 == H ? (rows - 1) * u.ds + (size_t)W * dbs : rows * u.ds, cudaMemcpyHostToDevice, s->st_in));
        else CU_TRY(s, cudaMemcpy2DAsync(u.d_depth + y0 * ds_dev_, ds_dev_, u.depth + y0 * u.ds, u.ds, (size_t)W * dbs, rows, cudaMemcpyHostToDevice, s->st_in));
      }
    }
    return NES_OK;
  };

  // destination planes: straight into the caller's memory when it is pinned and laid out like av_image_alloc
  const int nimg = want_depth ? 2 : 1, nplanes = jb->nv12 ? 2 : 3;
  bool direct = out->mem == NES_MEM_HOST;
  for (int im = 0; direct && im < nimg; im++) {
    uint8_t *const *pl = im ? out->depth : out->scene;
    const PlaneLayout &L = im ? pd : ps;
    const bool contiguous = pl[1] == pl[0] + L.bytes[0] && (jb->nv12 || pl[2] == pl[1] + L.bytes[1]);
    direct = contiguous && is_pinned(pl[0]);
  }
  banded = banded && direct && jb->segs_y >= 2;
  sl.dl.out = *out; sl.dl.ps = ps; sl.dl.pd = pd; sl.dl.direct = direct; sl.dl.want_depth = want_depth; sl.dl.nimg = nimg; sl.dl.nplanes = nplanes;
  sl.deferred = 0;

  if (s->mux) {
    // ---- multiplexed session: the frame is staged (uploads enqueued, descriptor built); the mux's dispatcher launches it
    // together with the ready frames of its other sessions and enqueues the download
    if (pinned_upload && (st = upload_rows(0, H))) return st;
    CU_TRY(s, cudaEventRecord(sl.e_in, s->st_in));
    sl.deferred = 1; sl.dispatched = 0; sl.dl_status = NES_OK;
    sl.busy = true;
    sl.ticket = s->next_ticket++;
    *ticket = sl.ticket;
    return mux_enqueue(s->mux, s, (int)(&sl - s->slots.data()));
  }
  if (banded) {
    // ---- low-latency path: upload, convert and download the frame in row bands, so that the download of
    // band b overlaps the upload of band b+1 and only the last band's kernel + download follow the last
    // uploaded byte.  Bands are whole segments; a band's kernel needs HALO source rows below its last row.
    const int nb = std::min(s->latency_bands, (int)jb->segs_y), S = jb->seg_rows;
    int uploaded = 0;
    sl.n_launches = 0;
    for (int b = 0; b < nb; b++) {
      const int seg_lo = b * jb->segs_y / nb, seg_hi = (b + 1) * jb->segs_y / nb;
      const int r0 = seg_lo * S, r1 = std::min(seg_hi * S, H);
      const int up_end = (b == nb - 1) ? H : std::min(r1 + HALO, H);
      if ((st = upload_rows(uploaded, up_end))) return st;
      uploaded = up_end;
      CU_TRY(s, cudaEventRecord(sl.e_band_in[b], s->st_in));
      CU_TRY(s, cudaStreamWaitEvent(s->st_k, sl.e_band_in[b], 0));
      if (b == 0) CU_TRY(s, cudaEventRecord(sl.e_k0, s->st_k));
      const int l = launch_frame_strips(sl.d_job, sl.h_job, 1, s->d_counters, &s->strips_seq, s->st_k, seg_lo * jb->strips_x, seg_hi * jb->strips_x,
                                        inline_job ? sl.h_glyphs : nullptr);
      if (l < 0) { s->err = "launch_frame_strips failed"; return NES_ERR_CUDA; }
      sl.n_launches += l;
      s->launches += (uint64_t)l;
      CU_TRY(s, cudaGetLastError());
      CU_TRY(s, cudaEventRecord(sl.e_band_k[b], s->st_k));
      CU_TRY(s, cudaStreamWaitEvent(s->st_out, sl.e_band_k[b], 0));
      for (int im = 0; im < nimg; im++) {
        uint8_t *const *pl = im ? out->depth : out->scene;
        const int32_t *ls = im ? out->depth_linesize : out->scene_linesize;
        const PlaneLayout &L = im ? pd : ps;
        for (int p = 0; p < nplanes; p++) {
          const size_t y0 = p ? r0 / 2 : r0, y1 = p ? r1 / 2 : r1;
          CU_TRY(s, cudaMemcpyAsync(pl[p] + y0 * ls[p], sl.d_out + L.off[p] + y0 * ls[p], (y1 - y0) * ls[p], cudaMemcpyDeviceToHost, s->st_out));
        }
      }
    }
    CU_TRY(s, cudaEventRecord(sl.e_in, s->st_in));
    CU_TRY(s, cudaEventRecord(sl.e_k1, s->st_k));
  } else {
    if (pinned_upload && (st = upload_rows(0, H))) return st;
    CU_TRY(s, cudaEventRecord(sl.e_in, s->st_in));

    // ---- kernels ----
    CU_TRY(s, cudaStreamWaitEvent(s->st_k, sl.e_in, 0));
    CU_TRY(s, cudaEventRecord(sl.e_k0, s->st_k));
    sl.n_launches = run_kernels(s, sl.d_job, sl.h_job, 1, s->st_k, inline_job ? sl.h_glyphs : nullptr);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaEventRecord(sl.e_k1, s->st_k));

    // ---- download ----
    CU_TRY(s, cudaStreamWaitEvent(s->st_out, sl.e_k1, 0));
    if ((st = enqueue_download(s, sl))) return st;
  }
  CU_TRY(s, cudaEventRecord(sl.e_out, s->st_out));

  sl.busy = true;
  sl.ticket = s->next_ticket++;
  *ticket = sl.ticket;
  return NES_OK;
}

int nes_gpu_wait(nes_gpu_session *s, uint64_t ticket) {
  if (!s) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  Slot *sl = nullptr;
  for (Slot &c : s->slots)
    if (c.busy && c.ticket == ticket) sl = &c;
  if (!sl) return NES_ERR_BAD_TICKET;
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  if (sl->deferred) {
    mux_wait_dispatched(s->mux, s, (int)(sl - s->slots.data()));
    if (sl->dl_status != NES_OK) { sl->busy = false; return sl->dl_status; }
  }
  CU_TRY(s, cudaEventSynchronize(sl->e_out));
  for (const StagedCopy &c : sl->staged) std::memcpy(c.dst, c.src, c.bytes);
  sl->staged.clear();
  float h2d = 0, k = 0, d2h = 0, tot = 0;
  cudaEventElapsedTime(&h2d, sl->e_start, sl->e_in);
  if (!sl->deferred) {  // a multiplexed frame shares its launch with other sessions' frames: no per-frame kernel interval
    cudaEventElapsedTime(&k, sl->e_k0, sl->e_k1);
    cudaEventElapsedTime(&d2h, sl->e_k1, sl->e_out);
  }
  cudaEventElapsedTime(&tot, sl->e_start, sl->e_out);
  cudaGetLastError();
  s->last.h2d_us = h2d * 1000.f; s->last.kernels_us = k * 1000.f; s->last.d2h_us = d2h * 1000.f; s->last.total_us = tot * 1000.f;
  s->last.n_launches = sl->n_launches;
  sl->busy = false;
  CU_TRY(s, cudaGetLastError());
  return NES_OK;
}

int nes_gpu_convert(nes_gpu_session *s, const nes_frame_in *in, const nes_text_run *runs, int n_runs, const nes_frame_out *out) {
  uint64_t t;
  const int st = nes_gpu_submit(s, in, runs, n_runs, out, &t);
  if (st) return st;
  return nes_gpu_wait(s, t);
}

int nes_gpu_last_timing(nes_gpu_session *s, nes_timing *t) {
  if (!s || !t) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  *t = s->last;
  return NES_OK;
}

// Descriptor table of a batch of device-resident frames -> bt (pinned + device copies, uploaded on the copy
// stream; bt.up is recorded behind the upload).
static int build_batch(nes_gpu_session *s, BatchTables &bt, int n_frames, const nes_frame_in *in, const nes_text_run *const *runs, const int *n_runs,
                       const nes_frame_out *out, bool allow_inline = false) {
  const int gl_cap = s->cfg.max_glyphs * kBatchGlyphFactor;
  const size_t gl_bytes = (size_t)gl_cap * sizeof(DevPlaced);
  if (!bt.h_jobs) {
    CU_TRY(s, cudaHostAlloc((void **)&bt.h_jobs, sizeof(DevJob) * kMaxBatch, cudaHostAllocDefault));
    CU_TRY(s, cudaMalloc((void **)&bt.d_jobs, sizeof(DevJob) * kMaxBatch));
    CU_TRY(s, cudaHostAlloc((void **)&bt.h_glyphs, gl_bytes, cudaHostAllocDefault));
    CU_TRY(s, cudaMalloc((void **)&bt.d_glyphs, gl_bytes));
    CU_TRY(s, cudaEventCreateWithFlags(&bt.done, cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&bt.up, cudaEventDisableTiming));
  }
  if (bt.used) CU_TRY(s, cudaEventSynchronize(bt.done));  // tables still referenced by an older launch
  int tile_base = 0, gl_used = 0;
  for (int f = 0; f < n_frames; f++) {
    int bpp, base, a_off; bool bgr;
    if (fmt_info(in[f].pix_fmt, &bpp, &base, &a_off, &bgr)) return NES_ERR_INVALID_ARG;
    if (in[f].mem != NES_MEM_DEVICE || out[f].mem != NES_MEM_DEVICE) return NES_ERR_INVALID_ARG;
    int st = validate(s, &in[f], &out[f], bpp);
    if (st) return st;
    const int W = in[f].width, H = in[f].height;
    const bool general = W != out[f].width || H != out[f].height || H < MIN_FUSED_H;
    const bool want_depth = out[f].depth[0] != nullptr;
    DevJob *jb = &bt.h_jobs[f];
    job_common(jb, &in[f], &out[f], bpp, base, a_off, bgr);
    const int n_gl = place_text(s, W, H, runs ? runs[f] : nullptr, (runs && n_runs) ? n_runs[f] : 0, bt.h_glyphs + gl_used, std::min(s->cfg.max_glyphs, gl_cap - gl_used));
    if (n_gl < 0) return n_gl;
    jb->glyphs = bt.d_glyphs + gl_used;
    jb->atlas = s->d_atlas;
    jb->n_glyphs = n_gl;
    gl_used += n_gl;
    for (int k = 0; k < in[f].n_sources; k++) {
      jb->src[k].rgb = in[f].src[k].rgb;
      jb->src[k].rgb_stride = in[f].src[k].rgb_stride ? in[f].src[k].rgb_stride : W * bpp;
      jb->src[k].depth = (want_depth || in[f].n_sources > 1) ? in[f].src[k].depth : nullptr;
      jb->src[k].depth_stride = in[f].src[k].depth_stride ? in[f].src[k].depth_stride : W * (in[f].depth_fmt == NES_DEPTH_GRAY16LE ? 2 : 1);
    }
    jb->sy = out[f].scene[0]; jb->su = out[f].scene[1]; jb->sv = out[f].scene[2];
    jb->sys = out[f].scene_linesize[0]; jb->sus = out[f].scene_linesize[1]; jb->svs = out[f].scene_linesize[2];
    if (want_depth) {
      jb->dy = out[f].depth[0]; jb->du = out[f].depth[1]; jb->dv = out[f].depth[2];
      jb->dys = out[f].depth_linesize[0]; jb->dus = out[f].depth_linesize[1]; jb->dvs = out[f].depth_linesize[2];
    }
    jb->nv12 = out[f].pix_fmt == NES_OUT_NV12;
    if (in[f].depth_fmt == NES_DEPTH_GRAY16LE && want_depth) job_depth16(jb);
    jb->general = general;
    if (n_gl > 0) band_glyphs(jb, bt.h_glyphs + gl_used - n_gl, n_gl, &s->scratch_banded);
    if (general) {
      FilterSet *fs;
      if ((st = get_filters(s, W, H, out[f].width, out[f].height, &fs))) return st;
      jb->hl = fs->hl; jb->hc = fs->hc; jb->vl = fs->vl; jb->vc = fs->vc;
      jb->half = fs->half; jb->csW = fs->csW; jb->rs_smem = fs->smem_need;
      jb->rs_tw = fs->tw; jb->rs_th = fs->th; jb->rs_win_x = fs->win_x; jb->rs_win_y = fs->win_y; jb->rs_lay = fs->layout;
      jb->rz_dw = getenv("NES_NO_RZ") ? 0 : fs->rz_dw[bpp - 3];
    }
    job_alignment(jb);
    job_tiles(jb, tile_base);
    tile_base += jb->tiles_x * jb->tiles_y;
    if ((st = job_tmaps(s, jb))) return st;
    job_tile_mask(jb, bt.h_glyphs + gl_used - n_gl);
  }
  plan_frame_strips(bt.h_jobs, n_frames);
  plan_resize_strips(bt.h_jobs, n_frames);
  bt.n = n_frames;
  bt.inline_job = allow_inline && frame_strips_inline_ok(bt.h_jobs, n_frames);
  if (bt.inline_job) return NES_OK;  // the one frame travels in its kernel's parameter block: nothing to upload
  // descriptor upload on the copy stream, so that it overlaps the kernels of the previous batch
  if (gl_used > 0) CU_TRY(s, cudaMemcpyAsync(bt.d_glyphs, bt.h_glyphs, (size_t)gl_used * sizeof(DevPlaced), cudaMemcpyHostToDevice, s->st_in));
  CU_TRY(s, cudaMemcpyAsync(bt.d_jobs, bt.h_jobs, sizeof(DevJob) * n_frames, cudaMemcpyHostToDevice, s->st_in));
  CU_TRY(s, cudaEventRecord(bt.up, s->st_in));
  return NES_OK;
}

int nes_gpu_convert_batch_device(nes_gpu_session *s, int n_frames, const nes_frame_in *in, const nes_text_run *const *runs,
                                 const int *n_runs, const nes_frame_out *out, int sync) {
  if (!s || !in || !out || n_frames < 1 || n_frames > kMaxBatch) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->sticky) return s->sticky;
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  BatchTables &bt = s->batch[s->batch_seq++ % kBatchRing];
  const int st = build_batch(s, bt, n_frames, in, runs, n_runs, out, /*allow_inline=*/true);
  if (st) return st;
  if (!bt.inline_job) CU_TRY(s, cudaStreamWaitEvent(s->st_k, bt.up, 0));
  run_kernels(s, bt.d_jobs, bt.h_jobs, n_frames, s->st_k, bt.inline_job ? bt.h_glyphs : nullptr);
  CU_TRY(s, cudaGetLastError());
  CU_TRY(s, cudaEventRecord(bt.done, s->st_k));
  bt.used = true;
  if (sync) CU_TRY(s, cudaStreamSynchronize(s->st_k));
  return NES_OK;
}

// ---- prepared batches: descriptors built and uploaded once, launched many times -------------------------
struct nes_gpu_batch {
  BatchTables bt;
};

int nes_gpu_batch_prepare(nes_gpu_session *s, int n_frames, const nes_frame_in *in, const nes_text_run *const *runs, const int *n_runs,
                          const nes_frame_out *out, nes_gpu_batch **batch) {
  if (!s || !in || !out || !batch || n_frames < 1 || n_frames > kMaxBatch) return NES_ERR_INVALID_ARG;
  *batch = nullptr;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->sticky) return s->sticky;
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  nes_gpu_batch *b = new (std::nothrow) nes_gpu_batch();
  if (!b) return NES_ERR_NO_MEMORY;
  const int st = build_batch(s, b->bt, n_frames, in, runs, n_runs, out);
  // the table is resident before the call returns: nes_gpu_batch_run is then nothing but the launches, back to
  // back on the compute stream (no event in between: consecutive launches overlap, frame_strips.cu)
  cudaError_t e = st == NES_OK ? cudaStreamSynchronize(s->st_in) : cudaSuccess;
  if (st != NES_OK || e != cudaSuccess) {
    BatchTables &t = b->bt;
    cudaFreeHost(t.h_jobs); cudaFree(t.d_jobs); cudaFreeHost(t.h_glyphs); cudaFree(t.d_glyphs);
    if (t.done) cudaEventDestroy(t.done);
    if (t.up) cudaEventDestroy(t.up);
    delete b;
    if (st != NES_OK) return st;
    CU_TRY(s, e);
  }
  *batch = b;
  return NES_OK;
}

int nes_gpu_batch_run(nes_gpu_session *s, nes_gpu_batch *b, int sync) {
  if (!s || !b) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->sticky) return s->sticky;
  CU_TRY(s, cudaSetDevice(s->cfg.device));
  run_kernels(s, b->bt.d_jobs, b->bt.h_jobs, b->bt.n, s->st_k);
  CU_TRY(s, cudaGetLastError());
  if (sync) CU_TRY(s, cudaStreamSynchronize(s->st_k));
  return NES_OK;
}

void nes_gpu_batch_free(nes_gpu_session *s, nes_gpu_batch *b) {
  if (!s || !b) return;
  std::lock_guard<std::mutex> lk(s->mu);
  cudaSetDevice(s->cfg.device);
  cudaStreamSynchronize(s->st_k);
  BatchTables &t = b->bt;
  cudaFreeHost(t.h_jobs); cudaFree(t.d_jobs); cudaFreeHost(t.h_glyphs); cudaFree(t.d_glyphs);
  if (t.done) cudaEventDestroy(t.done);
  if (t.up) cudaEventDestroy(t.up);
  cudaGetLastError();
  delete b;
}

}  // extern "C"

// ============================================================================
// nes_gpu_mux: many client sessions of one GPU, one launch set for all their ready frames
// (BASELINE config 4: 64 concurrent 1080p sessions; the reference runs one process per session, main.cpp:133-171,
// and one process_frame_thread per eye, :274-282).  A session attached to a mux keeps its own upload / download
// streams, frame slots, atlas and caches; nes_gpu_submit stages the frame (descriptor, H2D copies) on the caller's
// thread and hands the slot to the dispatcher thread, which gathers whatever is ready -- never waiting for a batch to
// fill -- into ONE descriptor table and ONE launch of k_frame_strips / k_resize_strips, then enqueues each frame's
// download on its session's stream.  nes_gpu_wait is unchanged for the caller.
// ============================================================================
struct nes_gpu_mux {
  int device = 0, max_batch = 64;
  cudaStream_t st_k = nullptr;
  uint32_t *d_counters = nullptr;
  uint64_t strips_seq = 0;
  // copy streams shared by the attached sessions (session i uploads on st_up[i % kCopyStreams], downloads on st_down[...])
  static constexpr int kCopyStreams = 4;
  cudaStream_t st_up[kCopyStreams] = {}, st_down[kCopyStreams] = {};
  std::mutex att_mu;
  std::vector<nes_gpu_session *> attached;
  uint64_t attach_seq = 0;
  static constexpr int kTables = 4;
  BatchTables tables[kTables];
  uint64_t table_seq = 0;
  std::mutex q_mu;
  std::condition_variable q_cv;
  std::vector<std::pair<nes_gpu_session *, int>> pending;
  bool stop = false;
  std::mutex done_mu;
  std::condition_variable done_cv;
  std::thread worker;
  uint64_t frames = 0, launch_sets = 0, launches = 0, max_batch_seen = 0;
  std::string err;
};

static int mux_enqueue(nes_gpu_mux *m, nes_gpu_session *s, int slot_index) {
  {
    std::lock_guard<std::mutex> lk(m->q_mu);
    m->pending.emplace_back(s, slot_index);
  }
  m->q_cv.notify_one();
  return NES_OK;
}

static void mux_wait_dispatched(nes_gpu_mux *m, nes_gpu_session *s, int slot_index) {
  std::unique_lock<std::mutex> lk(m->done_mu);
  m->done_cv.wait(lk, [&] { return s->slots[(size_t)slot_index].dispatched != 0; });
}

static void mux_dispatch(nes_gpu_mux *m, std::vector<std::pair<nes_gpu_session *, int>> &batch) {
  const int n = (int)batch.size();
  BatchTables &bt = m->tables[m->table_seq++ % nes_gpu_mux::kTables];
  int status = NES_OK;
  auto cu = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess && status == NES_OK) { status = NES_ERR_CUDA; m->err = std::string(what) + ": " + cudaGetErrorString(e); }
  };
  if (bt.used) cu(cudaEventSynchronize(bt.done), "cudaEventSynchronize(table)");
  // one descriptor table: the staged jobs of every session, re-planned as one launch
  int tile_base = 0;
  for (int i = 0; i < n; i++) {
    Slot &sl = batch[i].first->slots[(size_t)batch[i].second];
    std::memcpy(&bt.h_jobs[i], sl.h_job, sizeof(DevJob));
    job_tiles(&bt.h_jobs[i], tile_base);
    tile_base += bt.h_jobs[i].tiles_x * bt.h_jobs[i].tiles_y;
  }
  plan_frame_strips(bt.h_jobs, n);
  plan_resize_strips(bt.h_jobs, n);
  cu(cudaMemcpyAsync(bt.d_jobs, bt.h_jobs, sizeof(DevJob) * (size_t)n, cudaMemcpyHostToDevice, m->st_k), "cudaMemcpyAsync(table)");
  for (int i = 0; i < n; i++) cu(cudaStreamWaitEvent(m->st_k, batch[i].first->slots[(size_t)batch[i].second].e_in, 0), "cudaStreamWaitEvent(upload)");
  int l = 0;
  if (status == NES_OK) {
    const int r0 = launch_frame_strips(bt.d_jobs, bt.h_jobs, n, m->d_counters, &m->strips_seq, m->st_k);
    const int r1 = launch_resize_strips(bt.d_jobs, bt.h_jobs, n, m->d_counters, &m->strips_seq, m->st_k);
    const int r2 = launch_resize_tiles(bt.d_jobs, bt.h_jobs, n, m->st_k);
    const int r3 = launch_depth16(bt.h_jobs, n, m->st_k);
    if (r0 < 0 || r1 < 0 || r2 < 0 || r3 < 0) { status = NES_ERR_CUDA; m->err = "kernel launch failed"; }
    l = std::max(r0, 0) + std::max(r1, 0) + std::max(r2, 0) + std::max(r3, 0);
    cu(cudaGetLastError(), "launch");
  }
  cu(cudaEventRecord(bt.done, m->st_k), "cudaEventRecord(done)");
  bt.used = true;
  for (int i = 0; i < n; i++) {
    nes_gpu_session *s = batch[i].first;
    Slot &sl = s->slots[(size_t)batch[i].second];
    int st = status;
    if (st == NES_OK) {
      cu(cudaStreamWaitEvent(s->st_out, bt.done, 0), "cudaStreamWaitEvent(done)");
      st = enqueue_download(s, sl);
    }
    cu(cudaEventRecord(sl.e_out, s->st_out), "cudaEventRecord(out)");
    sl.n_launches = l;
    sl.dl_status = st != NES_OK ? st : status;
  }
  {
    std::lock_guard<std::mutex> lk(m->done_mu);
    for (int i = 0; i < n; i++) batch[i].first->slots[(size_t)batch[i].second].dispatched = 1;
    m->frames += (uint64_t)n; m->launch_sets++; m->launches += (uint64_t)l;
    m->max_batch_seen = std::max<uint64_t>(m->max_batch_seen, (uint64_t)n);
  }
  m->done_cv.notify_all();
}

// The session leaves the mux (its own destruction, or the mux's): back on the streams it created.
static void mux_forget(nes_gpu_mux *m, nes_gpu_session *s) {
  std::lock_guard<std::mutex> lk(m->att_mu);
  m->attached.erase(std::remove(m->attached.begin(), m->attached.end(), s), m->attached.end());
  s->st_in = s->own_in; s->st_out = s->own_out;
  s->mux = nullptr;
}

static void mux_worker(nes_gpu_mux *m) {
  cudaSetDevice(m->device);
  std::vector<std::pair<nes_gpu_session *, int>> batch;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(m->q_mu);
      m->q_cv.wait(lk, [&] { return m->stop || !m->pending.empty(); });
      if (m->pending.empty() && m->stop) return;
      const size_t take = std::min(m->pending.size(), (size_t)m->max_batch);
      batch.assign(m->pending.begin(), m->pending.begin() + (long)take);
      m->pending.erase(m->pending.begin(), m->pending.begin() + (long)take);
    }
    mux_dispatch(m, batch);
  }
}

extern "C" {

int nes_gpu_mux_create(int device, int max_batch, nes_gpu_mux **out) {
  if (!out) return NES_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return NES_ERR_CUDA; }
  if (device < 0 || device >= ndev) return NES_ERR_INVALID_ARG;
  nes_gpu_mux *m = new (std::nothrow) nes_gpu_mux();
  if (!m) return NES_ERR_NO_MEMORY;
  m->device = device;
  m->max_batch = max_batch < 1 ? 64 : std::min(max_batch, kMaxBatch);
  bool ok = cudaSetDevice(device) == cudaSuccess && kernels_init() == 0 && cudaStreamCreateWithFlags(&m->st_k, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < nes_gpu_mux::kCopyStreams; i++)
    ok = ok && cudaStreamCreateWithFlags(&m->st_up[i], cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&m->st_down[i], cudaStreamNonBlocking) == cudaSuccess;
  ok = ok &&
            cudaMalloc((void **)&m->d_counters, 2 * COUNTER_SLOTS * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemset(m->d_counters, 0, 2 * COUNTER_SLOTS * sizeof(uint32_t)) == cudaSuccess;
  for (BatchTables &bt : m->tables) {
    ok = ok && cudaHostAlloc((void **)&bt.h_jobs, sizeof(DevJob) * (size_t)m->max_batch, cudaHostAllocDefault) == cudaSuccess &&
         cudaMalloc((void **)&bt.d_jobs, sizeof(DevJob) * (size_t)m->max_batch) == cudaSuccess &&
         cudaEventCreateWithFlags(&bt.done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    nes_gpu_mux_destroy(m);
    return NES_ERR_CUDA;
  }
  m->worker = std::thread(mux_worker, m);
  *out = m;
  return NES_OK;
}

void nes_gpu_mux_destroy(nes_gpu_mux *m) {
  if (!m) return;
  if (m->worker.joinable()) {
    {
      std::lock_guard<std::mutex> lk(m->q_mu);
      m->stop = true;
    }
    m->q_cv.notify_all();
    m->worker.join();
  }
  cudaSetDevice(m->device);
  if (m->st_k) cudaStreamSynchronize(m->st_k);
  for (int i = 0; i < nes_gpu_mux::kCopyStreams; i++) {
    if (m->st_up[i]) cudaStreamSynchronize(m->st_up[i]);
    if (m->st_down[i]) cudaStreamSynchronize(m->st_down[i]);
  }
  // sessions that are still attached go back to their own streams (they stay usable, un-multiplexed)
  for (;;) {
    nes_gpu_session *s = nullptr;
    {
      std::lock_guard<std::mutex> lk(m->att_mu);
      if (!m->attached.empty()) s = m->attached.back();
    }
    if (!s) break;
    std::lock_guard<std::mutex> lk(s->mu);
    mux_forget(m, s);
  }
  for (BatchTables &bt : m->tables) {
    cudaFreeHost(bt.h_jobs); cudaFree(bt.d_jobs);
    if (bt.done) cudaEventDestroy(bt.done);
  }
  cudaFree(m->d_counters);
  if (m->st_k) cudaStreamDestroy(m->st_k);
  for (int i = 0; i < nes_gpu_mux::kCopyStreams; i++) {
    if (m->st_up[i]) cudaStreamDestroy(m->st_up[i]);
    if (m->st_down[i]) cudaStreamDestroy(m->st_down[i]);
  }
  cudaGetLastError();
  delete m;
}

int nes_gpu_mux_attach(nes_gpu_mux *m, nes_gpu_session *s) {
  if (!m || !s) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->cfg.device != m->device) return NES_ERR_INVALID_ARG;
  for (const Slot &sl : s->slots)
    if (sl.busy) return NES_ERR_BUSY;
  if (s->mux == m) return NES_OK;
  if (s->mux) mux_forget(s->mux, s);
  std::lock_guard<std::mutex> lk2(m->att_mu);
  const uint64_t i = m->attach_seq++;
  s->mux = m;
  s->st_in = m->st_up[i % nes_gpu_mux::kCopyStreams];
  s->st_out = m->st_down[i % nes_gpu_mux::kCopyStreams];
  m->attached.push_back(s);
  return NES_OK;
}

int nes_gpu_mux_stats(nes_gpu_mux *m, nes_mux_stats *out) {
  if (!m || !out) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(m->done_mu);
  out->frames = m->frames; out->launch_sets = m->launch_sets; out->launches = m->launches; out->max_batch = m->max_batch_seen;
  return NES_OK;
}

const char *nes_gpu_mux_error(nes_gpu_mux *m) { return m ? m->err.c_str() : ""; }

}  // extern "C"
