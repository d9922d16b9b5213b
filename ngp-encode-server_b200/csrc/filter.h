// filter.h -- polyphase filter tables (host), see filter.cc
#ifndef NES_FILTER_H_
#define NES_FILTER_H_
#include <cstdint>
#include <vector>

namespace nes {

struct FilterTable {
  int size = 0;               // taps per destination sample
  int dst = 0;                // destination samples
  std::vector<int16_t> coef;  // [dst][size]
  std::vector<int32_t> pos;   // [dst] first source sample of each row of taps
};

// libswscale-compatible bicubic table for src -> dst samples, rows summing to `one`.
// Returns the filter size or -1.
int build_filter(int src, int dst, int one, FilterTable *out);

}  // namespace nes
#endif
