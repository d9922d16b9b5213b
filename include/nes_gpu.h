/*
 * nes_gpu.h -- C ABI of libnes_gpu.so: the B200 (sm_100a) implementation of
 * ngp-encode-server's per-frame pixel pipeline
 *
 *     unpack -> [depth composite] -> text overlay -> RGB->YUV420P / GRAY8->YUV420P
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has
 * no FFI layer of its own; the path sits behind four C++ entry points, and each
 * function below names the one it replaces (paths relative to the reference
 * tree, /root/reference):
 *
 *   types::SwsContextManager(src, dst)          src/base/video/type_managers.cc:143-155
 *   RenderedFrame::convert_frame()              include/base/video/rendered_frame.h:24-33
 *   RenderTextContext::render_string_to_frame   src/base/video/render_text.cc:35-111
 *   socket_receive_blocking_lpf + ParseFromString + RenderedFrame ctor
 *                                               src/server.cpp:91-112,175,193-194
 *                                               src/base/video/rendered_frame.cc:5-27
 *
 * include/nes_gpu_shim.hpp rebuilds the four reference signatures on top of this
 * ABI so encode.cpp / server.cpp compile unchanged.
 *
 * Conventions
 *   - plain C types only; every function returns 0 (NES_OK) or a negative
 *     nes_status, never throws, never aborts.
 *   - the caller owns every host pointer it passes; they must stay valid until
 *     nes_gpu_wait() returns for the ticket (or until the synchronous call
 *     returns).  The library owns device buffers, streams, events, pinned
 *     staging and the glyph atlas.
 *   - the library never writes to source buffers (the overlay is applied to the
 *     device copy; the reference stamps the host buffer in place,
 *     render_text.cc:100-103 -- nobody reads it afterwards).
 *   - calls on different sessions are fully concurrent; calls on one session are
 *     serialised by an internal mutex.
 *   - arithmetic contract: output planes are bit-exact with libswscale's C path
 *     (SWS_BITEXACT|SWS_ACCURATE_RND, default bicubic scaler, BT.601 limited
 *     range), the stamp overlay and the depth composite are exact integer work.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns NES_ERR_CUDA.
 */
#ifndef NES_GPU_H_
#define NES_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NES_API __attribute__((visibility("default")))
#else
#define NES_API
#endif

#define NES_ABI_VERSION 2
#define NES_MAX_SOURCES 8

typedef enum nes_status {
  NES_OK = 0,
  NES_ERR_INVALID_ARG = -1,   /* null pointer, odd/too small size, bad enum            */
  NES_ERR_SHORT_BUFFER = -2,  /* rgb_bytes/depth_bytes smaller than stride*height       */
  NES_ERR_TOO_LARGE = -3,     /* frame exceeds the session's max_width/max_height       */
  NES_ERR_CUDA = -4,          /* CUDA runtime error (sticky per session; see strerror)  */
  NES_ERR_NO_MEMORY = -5,
  NES_ERR_BAD_TICKET = -6,
  NES_ERR_PARSE = -7,         /* malformed / truncated nesproto.RenderedFrame           */
  NES_ERR_NO_ATLAS = -8,      /* text runs submitted before nes_gpu_atlas_set           */
  NES_ERR_FREETYPE = -9,      /* FreeType could not be loaded / font could not be opened */
  NES_ERR_BUSY = -10,         /* ring full: wait on an older ticket first               */
  NES_ERR_UNSUPPORTED = -11   /* optional run-time dependency missing or of an unknown version */
} nes_status;

/* Pixel formats accepted on the scene input (AV_PIX_FMT_* names in the
 * reference; server.cpp:193-194 uses RGB24 + GRAY8).  Alpha only matters to the
 * composite (alpha == 0 marks an invalid pixel); the conversion ignores it. */
typedef enum nes_pix_fmt {
  NES_PIX_RGB24 = 0,
  NES_PIX_BGR24 = 1,
  NES_PIX_RGBA = 2,
  NES_PIX_BGRA = 3,
  NES_PIX_ARGB = 4,
  NES_PIX_ABGR = 5
} nes_pix_fmt;

/* RenderTextContext::RenderPosition, include/base/video/render_text.h:17-23 */
typedef enum nes_text_pos {
  NES_TEXT_LEFT_TOP = 0,
  NES_TEXT_LEFT_BOTTOM = 1,
  NES_TEXT_RIGHT_TOP = 2,
  NES_TEXT_RIGHT_BOTTOM = 3,
  NES_TEXT_CENTER = 4
} nes_text_pos;

typedef enum nes_mem_kind {
  NES_MEM_HOST = 0,  /* host pointers; H2D / D2H copies are part of the call          */
  NES_MEM_DEVICE = 1 /* device pointers on the session's device; kernels only          */
} nes_mem_kind;

typedef struct nes_gpu_session nes_gpu_session;

typedef struct nes_gpu_cfg {
  int device;      /* CUDA device ordinal                                              */
  int max_width;   /* largest source or destination width this session will see       */
  int max_height;
  int max_sources; /* 1..NES_MAX_SOURCES renderer inputs per frame                    */
  int ring_depth;  /* frames in flight (H2D | kernels | D2H overlap); 0 -> 3           */
  int max_glyphs;  /* placed glyphs per frame; 0 -> 8192                              */
} nes_gpu_cfg;

/* One rasterised glyph: exactly what FT_Load_Char(face, ch, FT_LOAD_RENDER)
 * leaves in face->glyph (render_text.cc:88-110). */
typedef struct nes_glyph {
  int32_t code;    /* (unsigned char) value, 0..255                                   */
  int32_t width;   /* bitmap.width                                                    */
  int32_t rows;    /* bitmap.rows                                                     */
  int32_t left;    /* bitmap_left                                                     */
  int32_t top;     /* bitmap_top                                                      */
  int32_t advance; /* advance.x >> 6                                                  */
  int32_t pitch;   /* bytes between coverage rows                                     */
  int32_t reserved;
  const uint8_t *coverage; /* rows * pitch bytes, 8-bit coverage                      */
} nes_glyph;

/* One render_string_to_frame(frame, position, content) call (render_text.h:27-29).
 * view_*: the sub-rectangle of the source frame that plays the role of the reference's
 * `frame` for this call -- pen placement uses view_w x view_h, the stamp is clipped to the
 * view (render_text.cc:98 clips to the frame it was given).  view_w == 0 (the default of a
 * zero-initialised struct) means the whole frame.  A side-by-side stereo frame
 * (BASELINE config 3) gives each eye its own set of runs: views (0,0,W/2,H) and (W/2,0,W/2,H). */
typedef struct nes_text_run {
  int32_t position; /* nes_text_pos                                                   */
  int32_t len;      /* bytes in text                                                  */
  const char *text; /* not NUL-terminated necessarily; '\n' starts a new line (+20 px) */
  int32_t view_x, view_y, view_w, view_h;
} nes_text_run;

/* One renderer's output for this frame (nes.proto:18-25 fields frame, depth). */
typedef struct nes_source {
  const uint8_t *rgb;   /* packed pixels                                              */
  const uint8_t *depth; /* GRAY8, may be NULL when the frame has no depth stream      */
  int32_t rgb_stride;   /* bytes per row; 0 -> tight (width * bytes_per_pixel)        */
  int32_t depth_stride; /* 0 -> tight (width)                                         */
  uint64_t rgb_bytes;   /* bytes readable at rgb   (validated; the reference does not) */
  uint64_t depth_bytes; /* bytes readable at depth                                    */
} nes_source;

typedef struct nes_frame_in {
  int32_t n_sources; /* 1 = plain convert; >1 = depth-select composite first          */
  int32_t pix_fmt;   /* nes_pix_fmt of every source                                   */
  int32_t width;     /* source size (Camera.width/height, rendered_frame.cc:14-25)    */
  int32_t height;
  int32_t mem;       /* nes_mem_kind of the source pointers                           */
  int32_t depth_fmt; /* nes_depth_fmt: 0 = GRAY8 (the reference, server.cpp:193-194), 1 = GRAY16LE */
  nes_source src[NES_MAX_SOURCES];
} nes_frame_in;

/* Sample format of the depth planes.  GRAY16LE (2 bytes per sample, depth_stride in bytes) is converted like libswscale
 * converts it: the 16-bit horizontal scaler and the 8x8 ordered dither of its vertical scaler.  Only frames with ONE
 * source can carry 16-bit depth (a 16-bit depth composite is not defined): NES_ERR_INVALID_ARG otherwise. */
typedef enum nes_depth_fmt {
  NES_DEPTH_GRAY8 = 0,
  NES_DEPTH_GRAY16LE = 1
} nes_depth_fmt;

/* Destination pixel layouts.  YUV420P is what the reference hands to libavcodec
 * (three planes).  NV12 (Y plane + one interleaved U0 V0 U1 V1 ... plane: [1] with
 * linesize >= 2*ceil(width/2), [2] ignored) is the layout hardware encoders take; the
 * sample values are identical, only the chroma storage differs. */
typedef enum nes_out_fmt {
  NES_OUT_YUV420P = 0,
  NES_OUT_NV12 = 1
} nes_out_fmt;

/* Destination: two YUV420P images laid out like FrameManager::FrameData
 * (type_managers.h:166-169) after av_image_alloc(..., align 32).  depth[0] may be
 * NULL to skip the depth stream. */
typedef struct nes_frame_out {
  int32_t width;  /* encoder size (CodecInitInfo, type_managers.h:74-97)              */
  int32_t height;
  int32_t mem;    /* nes_mem_kind of the plane pointers                               */
  int32_t pix_fmt; /* nes_out_fmt: 0 = YUV420P (the reference's encoder input), 1 = NV12 */
  uint8_t *scene[3];
  int32_t scene_linesize[3];
  int32_t reserved2;
  uint8_t *depth[3];
  int32_t depth_linesize[3];
  int32_t reserved3;
} nes_frame_out;

/* Per-stage device times of the last completed frame of the session, in
 * microseconds (CUDA events on the session's streams).  Replaces ScopedTimer,
 * include/base/scoped_timer.h:9-23 / encode.cpp:55,102-112. */
typedef struct nes_timing {
  float h2d_us;
  float kernels_us;
  float d2h_us;
  float total_us;     /* first H2D byte -> last D2H byte                              */
  int32_t n_launches; /* kernels launched for that frame                              */
  int32_t reserved;
} nes_timing;

/* ---- library / error ---------------------------------------------------- */
NES_API int nes_gpu_abi_version(void);
NES_API const char *nes_gpu_strerror(int status);
/* Text of the last CUDA error seen by the session ("" if none). */
NES_API const char *nes_gpu_session_error(nes_gpu_session *s);
NES_API int nes_gpu_device_count(void);

/* ---- session ------------------------------------------------------------ */
NES_API int nes_gpu_session_create(const nes_gpu_cfg *cfg, nes_gpu_session **out);
NES_API void nes_gpu_session_destroy(nes_gpu_session *s);
/* cudaStream_t of the compute stream (for event timing by a harness). */
NES_API void *nes_gpu_session_stream(nes_gpu_session *s);
/* Low-latency mode (1..8, default 1 = off): a frame submitted from pinned host memory to pinned host
 * planes, same size in and out, is uploaded, converted and downloaded in `bands` row bands, so the
 * download of a band overlaps the upload of the next and only the last band's kernel and download
 * follow the last uploaded byte.  Measured on B200 / PCIe Gen5: p50 latency -16 % for a 4K frame with
 * 2 bands (1.14 -> 0.96 ms), no gain at 1080p (the extra per-copy overheads eat it); more bands are
 * slower.  Throughput with several frames in flight is unchanged. */
NES_API int nes_gpu_session_set_latency_bands(nes_gpu_session *s, int bands);
/* Total kernel launches issued by the session since creation. */
NES_API uint64_t nes_gpu_session_launches(nes_gpu_session *s);

/* Pinned host memory for zero-copy DMA (a FrameManager can allocate its planes
 * here instead of av_image_alloc, type_managers.cc:119-121). */
NES_API int nes_gpu_host_alloc(size_t bytes, void **out);
NES_API void nes_gpu_host_free(void *p);
/* Plain device memory on the session's device (NES_MEM_DEVICE callers). */
NES_API int nes_gpu_device_alloc(nes_gpu_session *s, size_t bytes, void **out);
NES_API void nes_gpu_device_free(nes_gpu_session *s, void *p);
NES_API int nes_gpu_memcpy_h2d(nes_gpu_session *s, void *dst_dev, const void *src_host, size_t bytes);
NES_API int nes_gpu_memcpy_d2h(nes_gpu_session *s, void *dst_host, const void *src_dev, size_t bytes);

/* ---- overlay: glyph atlas + text (replaces RenderTextContext) ------------ */
/* Upload the rasterised glyph set (replaces the per-character FT_Load_Char of
 * render_text.cc:88; glyphs are rasterised once, not per frame). */
NES_API int nes_gpu_atlas_set(nes_gpu_session *s, const nes_glyph *glyphs, int n);
/* Convenience: dlopen FreeType (path may be NULL -> search), open `font`,
 * FT_Set_Char_Size(0, 20*64, 0, 0) like render_text.cc:12-32, rasterise codes
 * 0..255 and call nes_gpu_atlas_set. */
NES_API int nes_gpu_atlas_load_font(nes_gpu_session *s, const char *freetype_so, const char *font_path);

/* Host-only half of the above (no session, no GPU): fills glyphs[0..255]; their coverage
 * pointers point into `coverage` (pitch == width).  Returns 0, NES_ERR_FREETYPE, or
 * NES_ERR_TOO_LARGE when coverage_cap is too small (*coverage_used = bytes needed). */
NES_API int nes_font_rasterise(const char *freetype_so, const char *font_path, nes_glyph *glyphs,
                               uint8_t *coverage, uint64_t coverage_cap, uint64_t *coverage_used);

/* ---- the hot path -------------------------------------------------------- */
/* Asynchronous: stage + enqueue H2D, kernels, D2H for one frame.  `runs` are the
 * render_string_to_frame calls to apply to the scene BEFORE conversion, in call
 * order (encode.cpp:76-97). */
NES_API int nes_gpu_submit(nes_gpu_session *s, const nes_frame_in *in, const nes_text_run *runs,
                           int n_runs, const nes_frame_out *out, uint64_t *ticket);
/* Block until the frame of `ticket` is in the caller's destination planes. */
NES_API int nes_gpu_wait(nes_gpu_session *s, uint64_t ticket);
/* submit + wait: the body of RenderedFrame::convert_frame() with the overlay
 * calls folded in. */
NES_API int nes_gpu_convert(nes_gpu_session *s, const nes_frame_in *in, const nes_text_run *runs,
                            int n_runs, const nes_frame_out *out);
/* Many independent frames (sessions / eyes) in ONE set of launches.  All
 * pointers must be NES_MEM_DEVICE.  runs_per_frame may be NULL. */
NES_API int nes_gpu_convert_batch_device(nes_gpu_session *s, int n_frames, const nes_frame_in *in,
                                         const nes_text_run *const *runs, const int *n_runs,
                                         const nes_frame_out *out, int sync);
/* The same split in two: nes_gpu_batch_prepare validates the frames, lays out the text and leaves the descriptor
 * table resident on the device; nes_gpu_batch_run is then only the kernel launches (a stable set of device-resident
 * frames -- a ring of session buffers -- is converted again and again without per-frame host work, and consecutive
 * runs overlap on the device: the next launch fills the SMs the previous one drains).  The frames' pointers, sizes
 * and text are those given at prepare time.  Free with nes_gpu_batch_free before destroying the session. */
typedef struct nes_gpu_batch nes_gpu_batch;
NES_API int nes_gpu_batch_prepare(nes_gpu_session *s, int n_frames, const nes_frame_in *in,
                                  const nes_text_run *const *runs, const int *n_runs,
                                  const nes_frame_out *out, nes_gpu_batch **batch);
NES_API int nes_gpu_batch_run(nes_gpu_session *s, nes_gpu_batch *batch, int sync);
NES_API void nes_gpu_batch_free(nes_gpu_session *s, nes_gpu_batch *batch);
NES_API int nes_gpu_last_timing(nes_gpu_session *s, nes_timing *t);

/* ---- host-side pieces, exported so they can be checked without a GPU ------ */
/* libswscale initFilter() tables for the bicubic default scaler (the tables the
 * resize kernels consume).  one = 1<<14 (horizontal) or 1<<12 (vertical).
 * Returns the filter size (>0) or a negative status; writes up to coef_cap
 * coefficients (dst*size, row-major) and dst positions. */
NES_API int nes_gpu_filter_table(int src, int dst, int one, int16_t *coef, int coef_cap,
                                 int32_t *pos, int pos_cap);

/* Placement of one text run: fills x/y (top-left in the frame), glyph code per
 * placed glyph.  Mirrors the pen arithmetic of render_text.cc:47-110.  Returns
 * the number of glyphs placed or a negative status. */
typedef struct nes_placed_glyph {
  int32_t x, y;   /* frame position of bitmap pixel (0,0)                             */
  int32_t code;   /* index into the atlas                                             */
  int32_t reserved;
  int32_t clip_x, clip_y, clip_w, clip_h; /* part of the bitmap inside the run's view (bitmap coordinates) */
} nes_placed_glyph;
NES_API int nes_gpu_text_layout(nes_gpu_session *s, int frame_w, int frame_h, const nes_text_run *run,
                                nes_placed_glyph *out, int cap);

/* Zero-copy unpack of one length-prefixed nesproto.RenderedFrame
 * (server.cpp:91-112 framing: 8-byte native size_t length, then the message;
 * proto/nes.proto:18-25).  Offsets are relative to `buf`. */
typedef struct nes_unpacked_frame {
  uint64_t index;
  int32_t is_left;
  int32_t cam_is_left;
  int32_t width;
  int32_t height;
  int32_t n_matrix;
  float matrix[16];
  uint64_t frame_off, frame_len; /* bytes field 6                                     */
  uint64_t depth_off, depth_len; /* bytes field 7                                     */
  uint64_t consumed;             /* bytes of buf used (8 + message length)            */
} nes_unpacked_frame;
NES_API int nes_unpack_rendered_frame(const uint8_t *buf, uint64_t len, int has_length_prefix,
                                      nes_unpacked_frame *out);

/* ---- pinned receive ring: zero-copy ingest ----------------------------------
 * Replaces the three payload copies of the reference's ingest chain
 * (socket_receive_blocking_lpf server.cpp:91-112 -> ParseFromString :175 ->
 * RenderedFrame ctor rendered_frame.cc:5-27): recv() each length-prefixed message
 * straight into a slot of page-locked memory, commit it (parsed in place), pass the
 * returned nes_source (pointers into the slot) to nes_gpu_submit, release the slot
 * after nes_gpu_wait.  Thread-safe; slots are handed out round-robin. */
typedef struct nes_ingest_ring nes_ingest_ring;
NES_API int nes_ingest_ring_create(int n_slots, uint64_t slot_bytes, nes_ingest_ring **out);
NES_API void nes_ingest_ring_destroy(nes_ingest_ring *r);
/* Next free slot to receive into; NES_ERR_BUSY when every slot is still referenced. */
NES_API int nes_ingest_acquire(nes_ingest_ring *r, int *slot, uint8_t **buf, uint64_t *cap);
/* `len` bytes of one message are in the slot: locate the payload in place, check it against
 * Camera.width/height (NES_ERR_SHORT_BUFFER; the reference does not check), fill *src. */
NES_API int nes_ingest_commit(nes_ingest_ring *r, int slot, uint64_t len, int has_length_prefix,
                              int bytes_per_pixel, nes_unpacked_frame *info, nes_source *src);
NES_API int nes_ingest_release(nes_ingest_ring *r, int slot);

/* ---- many client sessions on one GPU: one launch for all their ready frames --------------
 * BASELINE config 4 (64 concurrent sessions sharded over the GPUs; the reference is one process per session,
 * main.cpp:133-171, one process_frame_thread per eye, :274-282).  Create one mux per GPU and attach that GPU's
 * sessions: nes_gpu_submit on an attached session stages the frame on the caller's thread (descriptor, H2D copies on
 * one of the mux's pooled copy streams) and hands it to the mux's dispatcher thread, which gathers whatever frames are
 * ready -- it never waits for a batch to fill -- into one descriptor table and one kernel launch, then enqueues every
 * frame's download.  nes_gpu_wait / nes_gpu_convert work as before.  A destroyed session detaches itself; destroying
 * the mux hands the sessions still attached back to their own streams (they stay usable, un-multiplexed). */
typedef struct nes_gpu_mux nes_gpu_mux;
typedef struct nes_mux_stats {
  uint64_t frames;      /* frames dispatched                                             */
  uint64_t launch_sets; /* dispatches (one descriptor table each)                        */
  uint64_t launches;    /* kernel launches                                               */
  uint64_t max_batch;   /* most frames in one dispatch                                   */
} nes_mux_stats;
NES_API int nes_gpu_mux_create(int device, int max_batch /* frames per launch; 0 -> 64 */, nes_gpu_mux **out);
NES_API void nes_gpu_mux_destroy(nes_gpu_mux *m);
NES_API int nes_gpu_mux_attach(nes_gpu_mux *m, nes_gpu_session *s);
NES_API int nes_gpu_mux_stats(nes_gpu_mux *m, nes_mux_stats *out);
NES_API const char *nes_gpu_mux_error(nes_gpu_mux *m);

/* ---- encoder hand-off: planes as a ref-counted AVFrame ---------------------------------
 * Replaces FrameManager::AVFrameWrapper / to_avframe() (type_managers.h:187-239), whose frame is not
 * ref-counted -- avcodec_send_frame copies every plane (encode.cpp:136-137,164-165) -- and whose
 * av_image_alloc block leaks.  The planes are wrapped with av_buffer_create (libavutil bound with dlopen:
 * avutil_so may be NULL -> NES_AVUTIL_SO / the default soname): the encoder takes a reference instead of a copy;
 * `release(opaque, planes[0])` runs when the last reference (the returned frame's, or the encoder's) is dropped.
 * planes[1..2] may be NULL (one-plane formats).  av_pix_fmt is libavutil's AVPixelFormat value (0 = YUV420P).
 * *out_frame is an AVFrame*; free it with nes_avframe_free (= av_frame_free).  Returns NES_ERR_UNSUPPORTED when
 * libavutil is missing or its AVFrame layout is not the one this build knows (checked against the live library). */
NES_API int nes_avframe_wrap(const char *avutil_so, uint8_t *const planes[3], const int linesize[3], int width, int height,
                             int av_pix_fmt, int64_t pts, void (*release)(void *opaque, uint8_t *base), void *opaque,
                             void **out_frame);
NES_API void nes_avframe_free(void **frame);
/* av_buffer_get_ref_count of the frame's first plane (tests: the encoder holds a reference, not a copy). */
NES_API int nes_avframe_ref_count(void *frame);
NES_API const char *nes_avframe_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NES_GPU_H_ */
