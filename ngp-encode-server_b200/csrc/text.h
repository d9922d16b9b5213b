// text.h -- host glyph atlas + text layout, see text.cc
#ifndef NES_TEXT_H_
#define NES_TEXT_H_
#include <cstdint>
#include <string>
#include <vector>

#include "nes_gpu.h"

namespace nes {

struct HostGlyph {
  int width = 0, rows = 0, left = 0, top = 0, advance = 0, pitch = 0;
  uint32_t offset = 0;    // into HostAtlas::coverage
  uint32_t mask_off = 0;  // into HostAtlas::mask (words)
  int wpr = 0;            // mask words per bitmap row
};

struct HostAtlas {
  bool valid = false;
  HostGlyph glyph[256];
  std::vector<uint8_t> coverage;
  // what the device stamps from: one bit per bitmap pixel, set where coverage != 0 (render_text.cc:100 tests
  // nothing else), rows padded to whole 32-bit words
  std::vector<uint32_t> mask;
};

// (Re)builds atlas->mask and the glyphs' mask_off / wpr from the coverage bytes.
void build_masks(HostAtlas *atlas);

// Pen arithmetic of RenderTextContext::render_string_to_frame (render_text.cc:47-110):
// appends one placed glyph per drawable character.  Returns glyphs placed.
int layout_run(const HostAtlas &atlas, int W, int H, const nes_text_run &run, std::vector<nes_placed_glyph> *out);

// Rasterise codes 0..255 with FreeType at 20 pt / 72 dpi (render_text.cc:12-32,88).
int rasterise_font(const char *freetype_so, const char *font_path, HostAtlas *atlas, std::string *err);

}  // namespace nes
#endif
