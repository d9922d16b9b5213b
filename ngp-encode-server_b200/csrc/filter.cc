// filter.cc -- host-side polyphase filter tables for the resize kernels.
//
// The reference resizes by calling sws_getContext(..., flags=0, NULL, NULL, NULL)
// (/root/reference/src/base/video/type_managers.cc:145-147), i.e. libswscale's default
// bicubic scaler (B=0, C=0.6).  The tables libswscale builds for that call are fully
// determined by (srcSize, dstSize, one); this file computes the same tables with the
// same integer arithmetic (int64, truncating division, error-diffusion normalisation;
// SURVEY.md Appendix A.3) so the device kernels reproduce the C path bit for bit.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "filter.h"

namespace nes {

namespace {

inline int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }

inline int floor_log2(unsigned v) {
  int n = 0;
  while (v > 1) { v >>= 1; ++n; }
  return n;
}

// bicubic kernel weight for distance d (30-bit fixed point), B = 0, C = 0.6
inline int64_t bicubic_weight(int64_t d) {
  const int64_t one24 = 1 << 24, one30 = (int64_t)1 << 30;
  const int64_t C = (int64_t)(0.6 * (1 << 24));
  if (d >= ((int64_t)1 << 31)) return 0;
  const int64_t dd = (d * d) >> 30;
  const int64_t ddd = (dd * d) >> 30;
  if (d < one30) return (12 * one24 - 6 * C) * ddd + (-18 * one24 + 6 * C) * dd + 6 * one24 * one30;
  return (-6 * C) * ddd + (30 * C) * dd + (-48 * C) * d + (24 * C) * one30;
}

}  // namespace

int build_filter(int src, int dst, int one, FilterTable *out) {
  if (src < 1 || dst < 1 || one < 1) return -1;
  const int64_t x_inc = (((int64_t)src << 16) + (dst >> 1)) / dst;
  const int64_t fone = (int64_t)1 << (54 - std::min(floor_log2((unsigned)(src / dst)), 8));
  std::vector<int32_t> pos((size_t)dst);
  std::vector<int64_t> w;  // [dst][taps]
  int taps;

  if (iabs64(x_inc - 0x10000) < 10) {
    taps = 1;
    w.assign((size_t)dst, fone);
    for (int i = 0; i < dst; i++) pos[i] = i;
  } else {
    taps = (x_inc <= 0x10000) ? 5 : 1 + (4 * src + dst - 1) / dst;
    taps = std::max(std::min(taps, src - 2), 1);
    w.assign((size_t)dst * taps, 0);
    int64_t centre = ((128 * x_inc) >> 7) - (((int64_t)128 * 0x10000) >> 7);
    for (int i = 0; i < dst; i++, centre += 2 * x_inc) {
      int xx = (int)((centre - (int64_t)(taps - 2) * 65536) / 131072);
      pos[i] = xx;
      for (int j = 0; j < taps; j++, xx++) {
        int64_t d = iabs64((int64_t)xx * 131072 - centre) << 13;
        if (x_inc > 0x10000) d = d * dst / src;
        w[(size_t)i * taps + j] = bicubic_weight(d) / (((int64_t)1 << 54) / fone);
      }
    }
  }

  // drop negligible taps: shift them out on the left, count them on the right
  const double cutoff = 0.002 * (double)fone;
  int kept = 0;
  for (int i = dst - 1; i >= 0; i--) {
    int64_t *f = &w[(size_t)i * taps];
    int64_t acc = 0;
    for (int j = 0; j < taps; j++) {
      acc += iabs64(f[0]);
      if ((double)acc > cutoff) break;
      if (i < dst - 1 && pos[i] >= pos[i + 1]) break;
      std::rotate(f, f + 1, f + taps);
      f[taps - 1] = 0;
      pos[i]++;
    }
    int need = taps;
    acc = 0;
    for (int j = taps - 1; j > 0; j--) {
      acc += iabs64(f[j]);
      if ((double)acc > cutoff) break;
      need--;
    }
    kept = std::max(kept, need);
  }

  // fold taps that fall outside the source onto the border sample
  std::vector<int64_t> v((size_t)dst * kept);
  for (int i = 0; i < dst; i++)
    for (int j = 0; j < kept; j++) v[(size_t)i * kept + j] = w[(size_t)i * taps + j];
  for (int i = 0; i < dst; i++) {
    int64_t *f = &v[(size_t)i * kept];
    if (pos[i] < 0) {
      for (int j = 1; j < kept; j++) {
        const int left = std::max(j + pos[i], 0);
        f[left] += f[j];
        f[j] = 0;
      }
      pos[i] = 0;
    }
    if (pos[i] + kept > src) {
      const int shift = pos[i] + std::min(kept - src, 0);
      int64_t spill = 0;
      for (int j = kept - 1; j >= 0; j--)
        if (pos[i] + j >= src) { spill += f[j]; f[j] = 0; }
      for (int j = kept - 1; j >= 0; j--) f[j] = (j < shift) ? 0 : f[j - shift];
      pos[i] -= shift;
      f[src - 1 - pos[i]] += spill;
    }
  }

  // normalise every row to `one`, diffusing the rounding error along the row
  out->size = kept;
  out->dst = dst;
  out->coef.assign((size_t)dst * kept, 0);
  out->pos = pos;
  for (int i = 0; i < dst; i++) {
    const int64_t *f = &v[(size_t)i * kept];
    int64_t sum = 0;
    for (int j = 0; j < kept; j++) sum += f[j];
    sum = (sum + one / 2) / one;
    if (sum == 0) sum = 1;
    int64_t err = 0;
    for (int j = 0; j < kept; j++) {
      const int64_t t = f[j] + err;
      const int64_t q = t >= 0 ? (t + (sum >> 1)) / sum : (t - (sum >> 1)) / sum;
      out->coef[(size_t)i * kept + j] = (int16_t)q;
      err = t - q * sum;
    }
  }
  return kept;
}

}  // namespace nes
