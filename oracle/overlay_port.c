/*
 * oracle/overlay_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's text overlay and of the (self-defined)
 * depth composite.
 *
 *  - nes_oracle_render_string follows RenderTextContext::render_string_to_frame,
 *    /root/reference/src/base/video/render_text.cc:35-111: pen placement by the
 *    5-value RenderPosition enum (include/base/video/render_text.h:17-23), per
 *    character bitmap stamp "coverage != 0 -> (255,255,255)" into a tightly
 *    packed RGB24 surface with bounds check, pen_x += advance.x >> 6, '\n' ->
 *    pen_x = start, pen_y += 20.  Glyph bitmaps are an INPUT (they come from the
 *    real FreeType, oracle/ref_freetype.py, exactly what FT_Load_Char(ch,
 *    FT_LOAD_RENDER) leaves in face->glyph at 20 pt / 72 dpi).
 *  - nes_oracle_composite: the reference has NO depth composite (SURVEY.md §0
 *    item 1, §8 a5); BASELINE.json's north_star asks for one.  Semantics are
 *    defined by this repo (DESIGN.md "composite"): per pixel pick, among valid
 *    sources (4-byte formats: alpha != 0; 3-byte formats: always valid), the one
 *    with the smallest depth, ties to the lowest source index; no valid source
 *    -> RGB (0,0,0), depth 255.  Parity for this stage is therefore self-defined.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this file.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define NES_ORACLE_API __attribute__((visibility("default")))

typedef struct {
  int32_t width;   /* slot->bitmap.width  */
  int32_t rows;    /* slot->bitmap.rows   */
  int32_t left;    /* slot->bitmap_left   */
  int32_t top;     /* slot->bitmap_top    */
  int32_t advance; /* slot->advance.x >> 6 */
  int32_t _pad;
  const uint8_t *buffer; /* rows*width coverage bytes (pitch == width) */
} nes_oracle_glyph;

/* render_text.h:17-23 */
enum { POS_LEFT_TOP = 0, POS_LEFT_BOTTOM = 1, POS_RIGHT_TOP = 2, POS_RIGHT_BOTTOM = 3, POS_CENTER = 4 };

/* glyphs: 256 entries indexed by (unsigned char)ch.  Returns pixels stamped. */
static long render_string_bpp(uint8_t *surface, uint32_t width, uint32_t height, int position, const char *content,
                              int len, const nes_oracle_glyph *glyphs, int bpp, int c_off) {
  int pen_x, pen_y;
  const int x_box = 300, y_box = 100, margin = 50; /* render_text.cc:48-50 */
  long stamped = 0;
  switch (position) {
    case POS_LEFT_TOP: pen_x = margin; pen_y = margin; break;
    case POS_LEFT_BOTTOM: pen_x = margin; pen_y = (int)(height - y_box + margin); break;
    case POS_RIGHT_TOP: pen_x = (int)(width - x_box + margin); pen_y = margin; break;
    case POS_RIGHT_BOTTOM: pen_x = (int)(width - x_box + margin); pen_y = (int)(height - y_box + margin); break;
    case POS_CENTER: pen_x = (int)(width / 2 - x_box); pen_y = (int)(height / 2 - y_box); break;
    default: pen_x = margin; pen_y = margin; break;
  }
  const int orig_pen_x = pen_x;
  for (int n = 0; n < len; n++) {
    const char ch = content[n];
    if (ch == '\n') { pen_x = orig_pen_x; pen_y += 20; continue; }
    const nes_oracle_glyph *g = &glyphs[(unsigned char)ch];
    const int x_max = pen_x + g->left + g->width;
    const int y_max = pen_y - g->top + g->rows;
    int i, j, p, q;
    for (j = pen_y - g->top, q = 0; j < y_max; j++, q++) {
      for (i = pen_x + g->left, p = 0; i < x_max; i++, p++) {
        if (i < 0 || j < 0 || (uint32_t)i >= width || (uint32_t)j >= height) continue;
        if (g->buffer[q * g->width + p]) {
          uint8_t *px = surface + ((size_t)j * width + (size_t)i) * bpp + c_off;
          px[0] = 255; px[1] = 255; px[2] = 255;
          stamped++;
        }
      }
    }
    pen_x += g->advance;
  }
  return stamped;
}

/* render_text.cc:35-111 as written: tightly packed RGB24 surface */
NES_ORACLE_API long nes_oracle_render_string(uint8_t *surface, uint32_t width, uint32_t height,
                                             int position, const char *content, int len,
                                             const nes_oracle_glyph *glyphs) {
  return render_string_bpp(surface, width, height, position, content, len, glyphs, 3, 0);
}

/* the same stamp on a 4-byte pixel surface (colour bytes at c_off..c_off+2, alpha untouched):
 * the reference only ever sees RGB24 (server.cpp:193); BASELINE configs 2/5 feed RGBA */
NES_ORACLE_API long nes_oracle_render_string4(uint8_t *surface, uint32_t width, uint32_t height,
                                              int position, const char *content, int len,
                                              const nes_oracle_glyph *glyphs, int c_off) {
  return render_string_bpp(surface, width, height, position, content, len, glyphs, 4, c_off);
}

/*
 * Depth-select composite of n sources.  rgb[k]: packed pixels, bpp 3 or 4,
 * a_off = byte offset of alpha inside a 4-byte pixel (ignored for bpp 3).
 * Output keeps the source pixel format (alpha copied from the winner; (0,0,0,0)
 * when no source is valid) plus one GRAY8 depth plane.
 */
NES_ORACLE_API void nes_oracle_composite(int n, const uint8_t *const *rgb, const int *rgb_stride, int bpp,
                                         int a_off, const uint8_t *const *depth, const int *depth_stride,
                                         int W, int H, uint8_t *out_rgb, int out_rgb_stride,
                                         uint8_t *out_depth, int out_depth_stride) {
  /* row-wise "running best" form of the per-pixel definition above (same result: a source
   * replaces the current winner only when strictly nearer, so ties keep the lowest index);
   * written so a plain C compiler vectorises it -- this is the CPU baseline's composite. */
  int16_t *best = (int16_t *)malloc(sizeof(int16_t) * (size_t)W);
  for (int y = 0; y < H; y++) {
    uint8_t *restrict o = out_rgb + (size_t)y * out_rgb_stride;
    memset(o, 0, (size_t)W * bpp);
    for (int x = 0; x < W; x++) best[x] = 256;
    for (int k = 0; k < n; k++) {
      const uint8_t *restrict p = rgb[k] + (size_t)y * rgb_stride[k];
      const uint8_t *restrict d = depth[k] + (size_t)y * depth_stride[k];
      if (bpp == 4) {
        const uint32_t *restrict p4 = (const uint32_t *)p;
        uint32_t *restrict o4 = (uint32_t *)o;
        if (((uintptr_t)p4 | (uintptr_t)o4) & 3) {
          for (int x = 0; x < W; x++)
            if (p[4 * x + a_off] != 0 && d[x] < best[x]) { best[x] = d[x]; memcpy(o + 4 * x, p + 4 * x, 4); }
        } else {
          const int sh = 8 * a_off; /* little-endian host */
          for (int x = 0; x < W; x++) {
            const int take = (((p4[x] >> sh) & 255u) != 0) & (d[x] < best[x]);
            best[x] = take ? d[x] : best[x];
            o4[x] = take ? p4[x] : o4[x];
          }
        }
      } else {
        for (int x = 0; x < W; x++)
          if (d[x] < best[x]) { best[x] = d[x]; o[3 * x] = p[3 * x]; o[3 * x + 1] = p[3 * x + 1]; o[3 * x + 2] = p[3 * x + 2]; }
      }
    }
    uint8_t *restrict od = out_depth + (size_t)y * out_depth_stride;
    for (int x = 0; x < W; x++) od[x] = (uint8_t)(best[x] > 255 ? 255 : best[x]);
  }
  free(best);
}
