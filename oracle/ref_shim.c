/*
 * oracle/ref_shim.c -> oracle/_ref/libnes_ref.so -- TEST INFRASTRUCTURE.
 *
 * "The reference, run here": the reference's hot path is ~40 lines of glue
 * around two third-party libraries (libswscale, FreeType).  Its own build
 * (CMake + libav, FreeType, websocketpp, protobuf development headers) cannot run
 * in this image, but the BINARIES of both libraries exist (FFmpeg 8.0.1
 * libswscale 9.1.100 in the opencv wheel, FreeType 2.14.3 in the pillow wheel).
 * This shim restates the glue -- nothing else -- with hand-declared prototypes
 * and calls the real libraries through dlopen:
 *
 *   nes_ref_sws_convert     = types::SwsContextManager ctor+dtor
 *                             /root/reference/src/base/video/type_managers.cc:143-155
 *                             (sws_getContext(..., flags, 0,0,0); sws_scale over all
 *                             rows; sws_freeContext) -- a fresh context per call.
 *   nes_ref_text_new        = RenderTextContext ctor, src/base/video/render_text.cc:10-33
 *                             (FT_Init_FreeType, FT_New_Face, FT_Set_Char_Size(0,20*64,0,0))
 *   nes_ref_text_render     = RenderTextContext::render_string_to_frame,
 *                             render_text.cc:35-111 (FT_Load_Char(FT_LOAD_RENDER) per
 *                             character, no cache, stamp 255 where coverage != 0)
 *   nes_ref_text_glyph      = what FT_Load_Char leaves in face->glyph (for building the
 *                             glyph table handed to oracle/overlay_port.c and to the
 *                             product's atlas in tests)
 *
 * It is used (a) to validate oracle/swscale_port.c and overlay_port.c, (b) as the
 * CPU baseline of kind "reference" in bench.py.  Never linked by the product.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NES_REF_API __attribute__((visibility("default")))

/* ---- libswscale / libavutil prototypes (public API, declared by hand) ---- */
struct SwsContext;
typedef struct SwsContext *(*fn_sws_getContext)(int, int, int, int, int, int, int, void *, void *, const double *);
typedef int (*fn_sws_scale)(struct SwsContext *, const uint8_t *const[], const int[], int, int, uint8_t *const[], const int[]);
typedef void (*fn_sws_freeContext)(struct SwsContext *);
typedef int (*fn_av_get_pix_fmt)(const char *);
typedef unsigned (*fn_version)(void);

static fn_sws_getContext p_sws_getContext;
static fn_sws_scale p_sws_scale;
static fn_sws_freeContext p_sws_freeContext;
static fn_av_get_pix_fmt p_av_get_pix_fmt;
static fn_version p_swscale_version;

/* ---- FreeType public structs, mirrored for LP64 (freetype/freetype.h, ftimage.h) ---- */
typedef struct { void *data; void (*finalizer)(void *); } FT_Generic_;
typedef struct { long xMin, yMin, xMax, yMax; } FT_BBox_;
typedef struct { long x, y; } FT_Vector_;
typedef struct { long width, height, horiBearingX, horiBearingY, horiAdvance, vertBearingX, vertBearingY, vertAdvance; } FT_Glyph_Metrics_;
typedef struct {
  unsigned int rows, width;
  int pitch;
  unsigned char *buffer;
  unsigned short num_grays;
  unsigned char pixel_mode, palette_mode;
  void *palette;
} FT_Bitmap_;
typedef struct FT_GlyphSlotRec_m {
  void *library, *face;
  struct FT_GlyphSlotRec_m *next;
  unsigned int glyph_index;
  FT_Generic_ generic;
  FT_Glyph_Metrics_ metrics;
  long linearHoriAdvance, linearVertAdvance;
  FT_Vector_ advance;
  int format;
  FT_Bitmap_ bitmap;
  int bitmap_left, bitmap_top;
  /* ... (rest unused) */
} FT_GlyphSlotRec_m;
typedef struct {
  long num_faces, face_index, face_flags, style_flags, num_glyphs;
  char *family_name, *style_name;
  int num_fixed_sizes;
  void *available_sizes;
  int num_charmaps;
  void *charmaps;
  FT_Generic_ generic;
  FT_BBox_ bbox;
  unsigned short units_per_EM;
  short ascender, descender, height, max_advance_width, max_advance_height, underline_position, underline_thickness;
  FT_GlyphSlotRec_m *glyph;
  /* ... (rest unused) */
} FT_FaceRec_m;

typedef int (*fn_FT_Init_FreeType)(void **);
typedef int (*fn_FT_New_Face)(void *, const char *, long, FT_FaceRec_m **);
typedef int (*fn_FT_Set_Char_Size)(FT_FaceRec_m *, long, long, unsigned, unsigned);
typedef int (*fn_FT_Load_Char)(FT_FaceRec_m *, unsigned long, int32_t);
typedef int (*fn_FT_Done_FreeType)(void *);
typedef void (*fn_FT_Library_Version)(void *, int *, int *, int *);
static fn_FT_Init_FreeType p_FT_Init_FreeType;
static fn_FT_New_Face p_FT_New_Face;
static fn_FT_Set_Char_Size p_FT_Set_Char_Size;
static fn_FT_Load_Char p_FT_Load_Char;
static fn_FT_Done_FreeType p_FT_Done_FreeType;
static fn_FT_Library_Version p_FT_Library_Version;
#define FT_LOAD_RENDER_ (1L << 2)

/* dlopen a dependency (or the library itself) with RTLD_GLOBAL; returns 0 on success */
NES_REF_API int nes_ref_dlopen(const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "nes_ref_dlopen(%s): %s\n", path, dlerror()); return -1; }
  return 0;
}

NES_REF_API int nes_ref_bind_swscale(const char *swscale_path, const char *avutil_path) {
  void *hu = dlopen(avutil_path, RTLD_NOW | RTLD_GLOBAL);
  void *hs = hu ? dlopen(swscale_path, RTLD_NOW | RTLD_GLOBAL) : NULL;
  if (!hu || !hs) { fprintf(stderr, "nes_ref_bind_swscale: %s\n", dlerror()); return -1; }
  p_sws_getContext = (fn_sws_getContext)dlsym(hs, "sws_getContext");
  p_sws_scale = (fn_sws_scale)dlsym(hs, "sws_scale");
  p_sws_freeContext = (fn_sws_freeContext)dlsym(hs, "sws_freeContext");
  p_swscale_version = (fn_version)dlsym(hs, "swscale_version");
  p_av_get_pix_fmt = (fn_av_get_pix_fmt)dlsym(hu, "av_get_pix_fmt");
  return (p_sws_getContext && p_sws_scale && p_sws_freeContext && p_av_get_pix_fmt) ? 0 : -2;
}

NES_REF_API unsigned nes_ref_swscale_version(void) { return p_swscale_version ? p_swscale_version() : 0; }

NES_REF_API int nes_ref_pix_fmt(const char *name) { return p_av_get_pix_fmt ? p_av_get_pix_fmt(name) : -1; }

/*
 * type_managers.cc:143-155.  src_fmt/dst is given by FFmpeg pixel-format NAME so no
 * enum value is hard-coded.  dst is YUV420P with the caller's three planes/strides.
 */
NES_REF_API int nes_ref_sws_convert(const uint8_t *src, int src_stride, const char *src_fmt, int W, int H,
                                    int Wd, int Hd, int flags, uint8_t *dy, int ys, uint8_t *du, int us,
                                    uint8_t *dv, int vs) {
  if (!p_sws_getContext) return -100;
  const int sf = p_av_get_pix_fmt(src_fmt), df = p_av_get_pix_fmt("yuv420p");
  if (sf < 0 || df < 0) return -101;
  struct SwsContext *ctx = p_sws_getContext(W, H, sf, Wd, Hd, df, flags, 0, 0, 0);
  if (!ctx) return -102; /* "Failed to allocate sws_context." */
  const uint8_t *const sdata[4] = {src, 0, 0, 0};
  const int sstride[4] = {src_stride, 0, 0, 0};
  uint8_t *const ddata[4] = {dy, du, dv, 0};
  const int dstride[4] = {ys, us, vs, 0};
  p_sws_scale(ctx, sdata, sstride, 0, H, ddata, dstride);
  p_sws_freeContext(ctx);
  return 0;
}

/* ------------------------------ text ------------------------------------ */
typedef struct { void *library; FT_FaceRec_m *face; } nes_ref_text;

NES_REF_API int nes_ref_bind_freetype(const char *freetype_path) {
  void *h = dlopen(freetype_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "nes_ref_bind_freetype: %s\n", dlerror()); return -1; }
  p_FT_Init_FreeType = (fn_FT_Init_FreeType)dlsym(h, "FT_Init_FreeType");
  p_FT_New_Face = (fn_FT_New_Face)dlsym(h, "FT_New_Face");
  p_FT_Set_Char_Size = (fn_FT_Set_Char_Size)dlsym(h, "FT_Set_Char_Size");
  p_FT_Load_Char = (fn_FT_Load_Char)dlsym(h, "FT_Load_Char");
  p_FT_Done_FreeType = (fn_FT_Done_FreeType)dlsym(h, "FT_Done_FreeType");
  p_FT_Library_Version = (fn_FT_Library_Version)dlsym(h, "FT_Library_Version");
  return (p_FT_Init_FreeType && p_FT_New_Face && p_FT_Set_Char_Size && p_FT_Load_Char && p_FT_Done_FreeType) ? 0 : -2;
}

/* render_text.cc:10-33 */
NES_REF_API void *nes_ref_text_new(const char *font_location) {
  if (!p_FT_Init_FreeType) return NULL;
  nes_ref_text *t = (nes_ref_text *)calloc(1, sizeof(*t));
  if (p_FT_Init_FreeType(&t->library)) { free(t); return NULL; }
  if (p_FT_New_Face(t->library, font_location, 0, &t->face)) { p_FT_Done_FreeType(t->library); free(t); return NULL; }
  if (p_FT_Set_Char_Size(t->face, 0, 20 * 64, 0, 0)) { p_FT_Done_FreeType(t->library); free(t); return NULL; }
  return t;
}

NES_REF_API void nes_ref_text_free(void *ctx) {
  nes_ref_text *t = (nes_ref_text *)ctx;
  if (!t) return;
  p_FT_Done_FreeType(t->library);
  free(t);
}

NES_REF_API int nes_ref_freetype_version(void *ctx) {
  nes_ref_text *t = (nes_ref_text *)ctx;
  int a = 0, b = 0, c = 0;
  if (t && p_FT_Library_Version) p_FT_Library_Version(t->library, &a, &b, &c);
  return a * 10000 + b * 100 + c;
}

/* render_text.cc:35-111; returns the number of pixels stamped */
static long text_render_bpp(void *ctx, uint8_t *surface, uint32_t width, uint32_t height, int opt,
                            const char *content, int len, int bpp, int c_off) {
  nes_ref_text *t = (nes_ref_text *)ctx;
  FT_GlyphSlotRec_m *slot = t->face->glyph;
  int pen_x, pen_y;
  int x_box = 300, y_box = 100, margin = 50;
  long stamped = 0;
  switch (opt) {
    case 0: pen_x = margin; pen_y = margin; break;
    case 1: pen_x = margin; pen_y = height - y_box + margin; break;
    case 2: pen_x = width - x_box + margin; pen_y = margin; break;
    case 3: pen_x = width - x_box + margin; pen_y = height - y_box + margin; break;
    case 4: pen_x = width / 2 - x_box; pen_y = height / 2 - y_box; break;
    default: pen_x = margin; pen_y = margin; break;
  }
  int orig_pen_x = pen_x;
  for (int n = 0; n < len; n++) {
    char ch = content[n];
    if (ch == '\n') { pen_x = orig_pen_x; pen_y = pen_y + 20; continue; }
    (void)p_FT_Load_Char(t->face, (unsigned long)ch, FT_LOAD_RENDER_); /* failure: stale slot is drawn */
    int i, j, p, q;
    int x_max = pen_x + slot->bitmap_left + slot->bitmap.width;
    int y_max = pen_y - slot->bitmap_top + slot->bitmap.rows;
    for (j = pen_y - slot->bitmap_top, q = 0; j < y_max; j++, q++) {
      for (i = pen_x + slot->bitmap_left, p = 0; i < x_max; i++, p++) {
        if (i < 0 || j < 0 || (uint32_t)i >= width || (uint32_t)j >= height) continue;
        if (slot->bitmap.buffer[q * slot->bitmap.width + p]) {
          surface[(j * width + i) * bpp + c_off] = 255;
          surface[(j * width + i) * bpp + c_off + 1] = 255;
          surface[(j * width + i) * bpp + c_off + 2] = 255;
          stamped++;
        }
      }
    }
    pen_x += slot->advance.x >> 6;
  }
  return stamped;
}

NES_REF_API long nes_ref_text_render(void *ctx, uint8_t *surface, uint32_t width, uint32_t height, int opt,
                                     const char *content, int len) {
  return text_render_bpp(ctx, surface, width, height, opt, content, len, 3, 0);
}

/* same loop on a 4-byte pixel surface (BASELINE configs 2/5 feed RGBA; the reference itself is RGB24 only) */
NES_REF_API long nes_ref_text_render4(void *ctx, uint8_t *surface, uint32_t width, uint32_t height, int opt,
                                      const char *content, int len, int c_off) {
  return text_render_bpp(ctx, surface, width, height, opt, content, len, 4, c_off);
}

/*
 * Glyph as FT_Load_Char(face, (unsigned long)(char)byte, FT_LOAD_RENDER) leaves it.
 * out5 = {width, rows, left, top, advance.x>>6}; coverage copied (pitch-aware) into
 * buf (cap bytes).  Returns FT error code, or -1 if buf too small.
 */
NES_REF_API int nes_ref_text_glyph(void *ctx, int byte, int32_t *out5, uint8_t *buf, int cap) {
  nes_ref_text *t = (nes_ref_text *)ctx;
  const char ch = (char)byte;
  int err = p_FT_Load_Char(t->face, (unsigned long)ch, FT_LOAD_RENDER_);
  FT_GlyphSlotRec_m *slot = t->face->glyph;
  out5[0] = (int32_t)slot->bitmap.width;
  out5[1] = (int32_t)slot->bitmap.rows;
  out5[2] = slot->bitmap_left;
  out5[3] = slot->bitmap_top;
  out5[4] = (int32_t)(slot->advance.x >> 6);
  const int need = (int)(slot->bitmap.width * slot->bitmap.rows);
  if (need > cap) return -1;
  /* the reference indexes buffer[q*width+p], i.e. assumes pitch == width */
  for (unsigned q = 0; q < slot->bitmap.rows; q++)
    for (unsigned p = 0; p < slot->bitmap.width; p++)
      buf[q * slot->bitmap.width + p] = slot->bitmap.buffer[q * slot->bitmap.width + p];
  return err;
}
