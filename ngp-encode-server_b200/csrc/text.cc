// text.cc -- host side of the text overlay: glyph atlas + pen placement.
//
// The reference rasterises every character of every string on every frame
// (FT_Load_Char(FT_LOAD_RENDER), /root/reference/src/base/video/render_text.cc:88) under
// a global mutex (:38).  Here the 256 possible `char` values are rasterised ONCE with the
// same FreeType calls (FT_Init_FreeType / FT_New_Face / FT_Set_Char_Size(0, 20*64, 0, 0),
// render_text.cc:12-32) into an atlas that lives in device memory; per frame the host
// only runs the pen arithmetic of render_text.cc:47-110 and emits a list of placed
// glyph rectangles that the kernels stamp inside their shared-memory tiles.
#include <dirent.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "text.h"

namespace nes {

// pen start by RenderPosition (render_text.cc:47-77): x_box 300, y_box 100, margin 50
static void pen_start(int pos, int W, int H, int *px, int *py) {
  const int x_box = 300, y_box = 100, margin = 50;
  switch (pos) {
    case NES_TEXT_LEFT_BOTTOM: *px = margin; *py = H - y_box + margin; break;
    case NES_TEXT_RIGHT_TOP: *px = W - x_box + margin; *py = margin; break;
    case NES_TEXT_RIGHT_BOTTOM: *px = W - x_box + margin; *py = H - y_box + margin; break;
    case NES_TEXT_CENTER: *px = (int)((unsigned)W / 2) - x_box; *py = (int)((unsigned)H / 2) - y_box; break;
    case NES_TEXT_LEFT_TOP:
    default: *px = margin; *py = margin; break;
  }
}

int layout_run(const HostAtlas &atlas, int W, int H, const nes_text_run &run, std::vector<nes_placed_glyph> *out) {
  // the run's view is what the reference calls `frame` (an eye of a side-by-side frame, or the whole frame)
  int vx = 0, vy = 0, vw = W, vh = H;
  if (run.view_w > 0 && run.view_h > 0) {
    vx = run.view_x; vy = run.view_y; vw = run.view_w; vh = run.view_h;
    // a view sticking out of the frame is cut to it (nothing outside the frame exists to be stamped)
    if (vx < 0 || vy < 0 || vx >= W || vy >= H) return 0;
  }
  // visible part of the view inside the frame, in view coordinates
  const int cx1 = (vx + vw > W) ? W - vx : vw, cy1 = (vy + vh > H) ? H - vy : vh;
  int pen_x, pen_y;
  pen_start(run.position, vw, vh, &pen_x, &pen_y);
  const int line_x = pen_x;
  int placed = 0;
  for (int n = 0; n < run.len; n++) {
    const unsigned char ch = (unsigned char)run.text[n];
    if (ch == '\n') {  // render_text.cc:82-86
      pen_x = line_x;
      pen_y += 20;
      continue;
    }
    const HostGlyph &g = atlas.glyph[ch];
    if (g.width > 0 && g.rows > 0) {
      const int gx = pen_x + g.left, gy = pen_y - g.top;  // view coordinates
      // keep only the part of the bitmap inside the view (the reference bounds-checks per pixel, :98)
      const int p0 = gx < 0 ? -gx : 0, q0 = gy < 0 ? -gy : 0;
      const int p1 = gx + g.width > cx1 ? cx1 - gx : g.width, q1 = gy + g.rows > cy1 ? cy1 - gy : g.rows;
      if (p0 < p1 && q0 < q1) {
        out->push_back(nes_placed_glyph{vx + gx, vy + gy, (int32_t)ch, 0, p0, q0, p1 - p0, q1 - q0});
        placed++;
      }
    }
    pen_x += g.advance;  // render_text.cc:109
  }
  return placed;
}

void build_masks(HostAtlas *atlas) {
  atlas->mask.clear();
  for (HostGlyph &g : atlas->glyph) {
    g.wpr = (g.width + 31) / 32;
    g.mask_off = (uint32_t)atlas->mask.size();
    if (g.width <= 0 || g.rows <= 0) { g.wpr = 0; continue; }
    atlas->mask.resize(atlas->mask.size() + (size_t)g.wpr * g.rows, 0u);
    for (int q = 0; q < g.rows; q++)
      for (int p = 0; p < g.width; p++)
        if (atlas->coverage[g.offset + (size_t)q * g.pitch + p]) atlas->mask[g.mask_off + (size_t)q * g.wpr + (p >> 5)] |= 1u << (p & 31);
  }
}

// ---- FreeType through dlopen (public API, LP64 layouts of freetype.h / ftimage.h) ----
namespace {
struct FtGeneric { void *data; void (*finalizer)(void *); };
struct FtBBox { long xMin, yMin, xMax, yMax; };
struct FtVector { long x, y; };
struct FtGlyphMetrics { long width, height, horiBearingX, horiBearingY, horiAdvance, vertBearingX, vertBearingY, vertAdvance; };
struct FtBitmap {
  unsigned rows, width;
  int pitch;
  unsigned char *buffer;
  unsigned short num_grays;
  unsigned char pixel_mode, palette_mode;
  void *palette;
};
struct FtGlyphSlot {
  void *library, *face;
  FtGlyphSlot *next;
  unsigned glyph_index;
  FtGeneric generic;
  FtGlyphMetrics metrics;
  long linearHoriAdvance, linearVertAdvance;
  FtVector advance;
  int format;
  FtBitmap bitmap;
  int bitmap_left, bitmap_top;
};
struct FtFace {
  long num_faces, face_index, face_flags, style_flags, num_glyphs;
  char *family_name, *style_name;
  int num_fixed_sizes;
  void *available_sizes;
  int num_charmaps;
  void *charmaps;
  FtGeneric generic;
  FtBBox bbox;
  unsigned short units_per_EM;
  short ascender, descender, height, max_advance_width, max_advance_height, underline_position, underline_thickness;
  FtGlyphSlot *glyph;
};
constexpr int kFtLoadRender = 1 << 2;
constexpr unsigned char kFtPixelModeGray = 2;

// A wheel-bundled libfreetype (pillow.libs/) names its dependencies by hashed sonames
// that live next to it and carries no RUNPATH: load those siblings first.
void preload_siblings(const char *so_path) {
  const std::string path(so_path);
  const size_t slash = path.rfind('/');
  if (slash == std::string::npos) return;
  const std::string dir = path.substr(0, slash);
  DIR *d = opendir(dir.c_str());
  if (!d) return;
  std::vector<std::string> names;
  while (dirent *e = readdir(d)) names.push_back(e->d_name);
  closedir(d);
  for (const char *prefix : {"libbrotlicommon", "libbrotlidec", "libpng16"})
    for (const std::string &n : names)
      if (n.compare(0, strlen(prefix), prefix) == 0) dlopen((dir + "/" + n).c_str(), RTLD_NOW | RTLD_GLOBAL);
  dlerror();
}
}  // namespace

int rasterise_font(const char *freetype_so, const char *font_path, HostAtlas *atlas, std::string *err) {
  const char *candidates[] = {freetype_so, getenv("NES_FREETYPE_SO"), "libfreetype.so.6", "libfreetype.so"};
  void *h = nullptr;
  std::string why = "not found";
  for (const char *c : candidates) {
    if (!c || !*c) continue;
    preload_siblings(c);
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
    if (const char *e = dlerror()) why = e;
  }
  if (!h) { *err = "cannot dlopen FreeType: " + why; return NES_ERR_FREETYPE; }
  auto init = (int (*)(void **))dlsym(h, "FT_Init_FreeType");
  auto new_face = (int (*)(void *, const char *, long, FtFace **))dlsym(h, "FT_New_Face");
  auto set_size = (int (*)(FtFace *, long, long, unsigned, unsigned))dlsym(h, "FT_Set_Char_Size");
  auto load_char = (int (*)(FtFace *, unsigned long, int32_t))dlsym(h, "FT_Load_Char");
  auto done = (int (*)(void *))dlsym(h, "FT_Done_FreeType");
  if (!init || !new_face || !set_size || !load_char || !done) { *err = "FreeType symbols missing"; return NES_ERR_FREETYPE; }
  void *lib = nullptr;
  FtFace *face = nullptr;
  if (init(&lib)) { *err = "FT_Init_FreeType failed"; return NES_ERR_FREETYPE; }
  if (new_face(lib, font_path, 0, &face)) { done(lib); *err = std::string("FT_New_Face failed for ") + font_path; return NES_ERR_FREETYPE; }
  if (set_size(face, 0, 20 * 64, 0, 0)) { done(lib); *err = "FT_Set_Char_Size failed"; return NES_ERR_FREETYPE; }
  atlas->coverage.clear();
  for (int c = 0; c < 256; c++) {
    HostGlyph &g = atlas->glyph[c];
    g = HostGlyph{};
    if (c == '\n') continue;  // never rasterised (render_text.cc:82-86)
    // the reference passes a (signed) char: bytes >= 0x80 become huge code points -> .notdef
    const int rc = load_char(face, (unsigned long)(char)c, kFtLoadRender);
    const FtGlyphSlot *slot = face->glyph;
    // a failed load leaves the slot holding the previous code's bitmap (the reference would re-stamp whatever
    // the previous character of the string left there, render_text.cc:88-93: not reproducible from a static
    // atlas): store an empty glyph with no advance -- a documented deviation
    if (rc != 0) continue;
    g.width = (int)slot->bitmap.width;
    g.rows = (int)slot->bitmap.rows;
    g.left = slot->bitmap_left;
    g.top = slot->bitmap_top;
    g.advance = (int)(slot->advance.x >> 6);
    g.pitch = g.width;
    g.offset = (uint32_t)atlas->coverage.size();
    if (g.width > 0 && g.rows > 0 && slot->bitmap.pixel_mode == kFtPixelModeGray) {
      // the reference indexes buffer[q*width + p] (render_text.cc:99), i.e. it assumes
      // pitch == width (true for 8-bit gray bitmaps); copy with that same indexing
      for (int q = 0; q < g.rows; q++)
        for (int p = 0; p < g.width; p++) atlas->coverage.push_back(slot->bitmap.buffer[q * g.width + p]);
    } else {
      g.width = g.rows = 0;
    }
  }
  done(lib);
  atlas->valid = true;
  return NES_OK;
}

}  // namespace nes
