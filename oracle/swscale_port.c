/*
 * oracle/swscale_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar, one thread) of the arithmetic the reference
 * runs on its per-frame hot path.  The reference itself contains only glue:
 *
 *   types::SwsContextManager   /root/reference/src/base/video/type_managers.cc:143-155
 *       sws_getContext(w,h,fmt, W,H,YUV420P, flags=0, NULL,NULL,NULL) + one sws_scale()
 *   RenderedFrame::convert_frame  /root/reference/include/base/video/rendered_frame.h:24-33
 *       scene RGB24 -> YUV420P and depth GRAY8 -> YUV420P
 *
 * The arithmetic lives in libswscale, an un-vendored, un-pinned system
 * dependency of the reference (README.md:17,32; CMakeLists.txt:44-45).  This
 * file restates libswscale's published C path (libswscale/input.c rgb24ToY_c /
 * rgb24ToUV_half_c / rgb24ToUV_c, swscale.c hScale16To15_c / hScale8To15_c /
 * lumRangeFromJpeg, utils.c initFilter (bicubic), output.c yuv2planeX_8_c /
 * yuv2plane1_8_c) for the pinned version libswscale 9.1.100 (FFmpeg 8.0.1), as
 * run with SWS_BITEXACT|SWS_ACCURATE_RND and the default (bicubic) scaler.
 *
 * Pinning: tests/test_oracle_swscale.py checks this port byte-for-byte against
 * the live libswscale 9.1.100 bundled in this image (oracle/ref_swscale.py) and
 * against the golden SHA-256 vectors in tests/golden/ (SURVEY.md Appendix C).
 * The reference has no tests of its own, so those are the only pins available.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this file.  The product (libnes_gpu.so) never links or calls it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NES_ORACLE_API __attribute__((visibility("default")))

static inline int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static inline int64_t i64abs(int64_t v) { return v < 0 ? -v : v; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static int log2floor_u(unsigned v) { /* av_log2: floor(log2(v)), av_log2(0)=0 */
  int n = 0;
  while (v > 1) { v >>= 1; n++; }
  return n;
}

/* ------------------------------------------------------------------------
 * initFilter (bicubic, B=0, C=0.6, no src/dst filter vectors), libswscale
 * utils.c.  Returns filter size; *out_coef is dstW*size int16, *out_pos dstW.
 * filter_align==1 is the C path; >1 only appends zero taps.
 * ---------------------------------------------------------------------- */
NES_ORACLE_API int nes_oracle_init_filter(int src_w, int dst_w, int one,
                                          int16_t **out_coef, int32_t **out_pos) {
  const int64_t x_inc = (((int64_t)src_w << 16) + (dst_w >> 1)) / dst_w;
  const int src_pos = 128, dst_pos = 128;
  const int64_t fone = (int64_t)1 << (54 - imin(log2floor_u((unsigned)(src_w / dst_w)), 8));
  int filter_size;
  int64_t *filter = NULL;
  int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(dst_w + 3));
  int i, j;

  if (i64abs(x_inc - 0x10000) < 10 && src_pos == dst_pos) {
    filter_size = 1;
    filter = (int64_t *)malloc(sizeof(int64_t) * (size_t)dst_w);
    for (i = 0; i < dst_w; i++) { filter[i] = fone; pos[i] = i; }
  } else {
    const int size_factor = 4; /* bicubic */
    int64_t x_dst_in_src;
    if (x_inc <= (1 << 16)) filter_size = 1 + size_factor;
    else filter_size = 1 + (size_factor * src_w + dst_w - 1) / dst_w;
    filter_size = imin(filter_size, src_w - 2);
    filter_size = imax(filter_size, 1);
    filter = (int64_t *)malloc(sizeof(int64_t) * (size_t)dst_w * (size_t)filter_size);
    x_dst_in_src = ((dst_pos * x_inc) >> 7) - ((src_pos * (int64_t)0x10000) >> 7);
    for (i = 0; i < dst_w; i++) {
      int xx = (int)((x_dst_in_src - (filter_size - 2) * ((int64_t)1 << 16)) / (1 << 17));
      pos[i] = xx;
      for (j = 0; j < filter_size; j++) {
        int64_t d = i64abs(((int64_t)xx * (1 << 17)) - x_dst_in_src) << 13;
        int64_t coeff;
        const int64_t B = 0;
        const int64_t C = (int64_t)(0.6 * (1 << 24));
        if (x_inc > (1 << 16)) d = d * dst_w / src_w;
        if (d >= ((int64_t)1 << 31)) {
          coeff = 0;
        } else {
          int64_t dd = (d * d) >> 30;
          int64_t ddd = (dd * d) >> 30;
          if (d < ((int64_t)1 << 30))
            coeff = (12 * (1 << 24) - 9 * B - 6 * C) * ddd +
                    (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                    (6 * (1 << 24) - 2 * B) * ((int64_t)1 << 30);
          else
            coeff = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd +
                    (-12 * B - 48 * C) * d + (8 * B + 24 * C) * ((int64_t)1 << 30);
        }
        coeff /= ((int64_t)1 << 54) / fone;
        filter[(size_t)i * filter_size + j] = coeff;
        xx++;
      }
      x_dst_in_src += 2 * x_inc;
    }
  }

  /* reduce: shift near-zero taps out on the left, count them on the right */
  const int filter2_size = filter_size;
  int min_filter_size = 0;
  const double cut = 0.002 * (double)fone; /* SWS_MAX_REDUCE_CUTOFF */
  for (i = dst_w - 1; i >= 0; i--) {
    int min = filter2_size;
    int64_t cut_off = 0;
    int64_t *f = filter + (size_t)i * filter2_size;
    for (j = 0; j < filter2_size; j++) {
      int k;
      cut_off += i64abs(f[0]);
      if ((double)cut_off > cut) break;
      if (i < dst_w - 1 && pos[i] >= pos[i + 1]) break;
      for (k = 1; k < filter2_size; k++) f[k - 1] = f[k];
      f[k - 1] = 0;
      pos[i]++;
    }
    cut_off = 0;
    for (j = filter2_size - 1; j > 0; j--) {
      cut_off += i64abs(f[j]);
      if ((double)cut_off > cut) break;
      min--;
    }
    if (min > min_filter_size) min_filter_size = min;
  }
  filter_size = min_filter_size; /* filterAlign == 1 on the C path */

  int64_t *fr = (int64_t *)malloc(sizeof(int64_t) * (size_t)dst_w * (size_t)filter_size);
  for (i = 0; i < dst_w; i++)
    for (j = 0; j < filter_size; j++)
      fr[(size_t)i * filter_size + j] = (j >= filter2_size) ? 0 : filter[(size_t)i * filter2_size + j];
  free(filter);

  /* fix borders */
  for (i = 0; i < dst_w; i++) {
    int64_t *f = fr + (size_t)i * filter_size;
    if (pos[i] < 0) {
      for (j = 1; j < filter_size; j++) {
        int left = imax(j + pos[i], 0);
        f[left] += f[j];
        f[j] = 0;
      }
      pos[i] = 0;
    }
    if (pos[i] + filter_size > src_w) {
      int shift = pos[i] + imin(filter_size - src_w, 0);
      int64_t acc = 0;
      for (j = filter_size - 1; j >= 0; j--) {
        if (pos[i] + j >= src_w) { acc += f[j]; f[j] = 0; }
      }
      for (j = filter_size - 1; j >= 0; j--) {
        if (j < shift) f[j] = 0;
        else f[j] = f[j - shift];
      }
      pos[i] -= shift;
      f[src_w - 1 - pos[i]] += acc;
    }
  }

  /* normalise with error diffusion */
  int16_t *coef = (int16_t *)calloc((size_t)dst_w * (size_t)filter_size + 8, sizeof(int16_t));
  for (i = 0; i < dst_w; i++) {
    int64_t error = 0, sum = 0;
    int64_t *f = fr + (size_t)i * filter_size;
    for (j = 0; j < filter_size; j++) sum += f[j];
    sum = (sum + one / 2) / one;
    if (!sum) sum = 1;
    for (j = 0; j < filter_size; j++) {
      int64_t v = f[j] + error;
      int int_v = (int)(v >= 0 ? (v + (sum >> 1)) / sum : (v - (sum >> 1)) / sum);
      coef[(size_t)i * filter_size + j] = (int16_t)int_v;
      error = v - int_v * sum;
    }
  }
  free(fr);
  *out_coef = coef;
  *out_pos = pos;
  return filter_size;
}

NES_ORACLE_API void nes_oracle_free(void *p) { free(p); }

/* BT.601 limited range, 15-bit (libswscale ff_yuv2rgb_coeffs[SWS_CS_DEFAULT] ->
 * rgb2yuv table), SURVEY.md Appendix A.1 */
enum { RY = 8414, GY = 16519, BY = 3208, RU = -4865, GU = -9528, BU = 14392, RV = 14392, GV = -12061, BV = -2332 };

/* horizontal polyphase, 16-bit input (RGB-derived planes): hScale16To15_c, sh=13 */
static void hscale16to15(int16_t *dst, int dst_w, const int16_t *src, const int16_t *filter,
                         const int32_t *pos, int size) {
  for (int i = 0; i < dst_w; i++) {
    int val = 0;
    for (int j = 0; j < size; j++) val += (int)src[pos[i] + j] * filter[size * i + j];
    dst[i] = (int16_t)imin(val >> 13, (1 << 15) - 1);
  }
}
/* horizontal polyphase, 8-bit input (gray): hScale8To15_c */
static void hscale8to15(int16_t *dst, int dst_w, const uint8_t *src, const int16_t *filter,
                        const int32_t *pos, int size) {
  for (int i = 0; i < dst_w; i++) {
    int val = 0;
    for (int j = 0; j < size; j++) val += (int)src[pos[i] + j] * filter[size * i + j];
    dst[i] = (int16_t)imin(val >> 7, (1 << 15) - 1);
  }
}

/* vertical polyphase to 8 bit: yuv2planeX_8_c (dither = constant 64) / yuv2plane1_8_c */
static void vscale_plane(uint8_t *dst, int dst_stride, int dst_w, int dst_h, const int16_t *plane15,
                         int plane_w, const int16_t *vf, const int32_t *vpos, int vsize) {
  for (int i = 0; i < dst_h; i++) {
    uint8_t *d = dst + (size_t)i * dst_stride;
    if (vsize == 1) {
      const int16_t *p = plane15 + (size_t)vpos[i] * plane_w;
      for (int x = 0; x < dst_w; x++) d[x] = (uint8_t)clip8((p[x] + 64) >> 7);
    } else {
      for (int x = 0; x < dst_w; x++) {
        int val = 64 << 12;
        for (int j = 0; j < vsize; j++)
          val += (int)plane15[(size_t)(vpos[i] + j) * plane_w + x] * vf[(size_t)i * vsize + j];
        d[x] = (uint8_t)clip8(val >> 19);
      }
    }
  }
}

static int xinc_is_unscaled(int s, int d) {
  int64_t x_inc = (((int64_t)s << 16) + (d >> 1)) / d;
  return i64abs(x_inc - 0x10000) < 10;
}

/*
 * Packed RGB (3 or 4 bytes/pixel, byte offsets r_off/g_off/b_off inside a pixel)
 * -> YUV420P with optional bicubic resize.  All of W,H,Wd,Hd must be even.
 * Returns 0, or -1 on bad arguments.
 */
NES_ORACLE_API int nes_oracle_rgb_to_yuv420p(const uint8_t *src, int src_stride, int bpp, int r_off,
                                             int g_off, int b_off, int W, int H, int Wd, int Hd,
                                             uint8_t *dy, int ys, uint8_t *du, int us, uint8_t *dv,
                                             int vs) {
  if (W < 4 || H < 4 || Wd < 2 || Hd < 2 || (W | H | Wd | Hd) & 1 || (bpp != 3 && bpp != 4)) return -1;
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  const int half = (Wd >> 1) <= (W >> 1);
  const int csW = half ? (W >> 1) : W;

  int16_t *hf_l, *hf_c, *vf_l, *vf_c;
  int32_t *hp_l, *hp_c, *vp_l, *vp_c;
  const int hs_l = nes_oracle_init_filter(W, Wd, 1 << 14, &hf_l, &hp_l);
  const int hs_c = nes_oracle_init_filter(csW, cdW, 1 << 14, &hf_c, &hp_c);
  const int vs_l = nes_oracle_init_filter(H, Hd, 1 << 12, &vf_l, &vp_l);
  const int vs_c = nes_oracle_init_filter(H, cdH, 1 << 12, &vf_c, &vp_c);

  int16_t *y14 = (int16_t *)malloc(sizeof(int16_t) * (size_t)(W + 16));
  int16_t *u14 = (int16_t *)malloc(sizeof(int16_t) * (size_t)(csW + 16));
  int16_t *v14 = (int16_t *)malloc(sizeof(int16_t) * (size_t)(csW + 16));
  int16_t *py = (int16_t *)malloc(sizeof(int16_t) * (size_t)Wd * H);
  int16_t *pu = (int16_t *)malloc(sizeof(int16_t) * (size_t)cdW * H);
  int16_t *pv = (int16_t *)malloc(sizeof(int16_t) * (size_t)cdW * H);

  for (int y = 0; y < H; y++) {
    const uint8_t *row = src + (size_t)y * src_stride;
    for (int x = 0; x < W; x++) { /* rgb24ToY_c */
      int r = row[x * bpp + r_off], g = row[x * bpp + g_off], b = row[x * bpp + b_off];
      y14[x] = (int16_t)((RY * r + GY * g + BY * b + (32 << 14) + (1 << 8)) >> 9);
    }
    if (half) { /* rgb24ToUV_half_c */
      for (int c = 0; c < csW; c++) {
        const uint8_t *p0 = row + (size_t)(2 * c) * bpp, *p1 = p0 + bpp;
        int r = p0[r_off] + p1[r_off], g = p0[g_off] + p1[g_off], b = p0[b_off] + p1[b_off];
        u14[c] = (int16_t)((RU * r + GU * g + BU * b + (256 << 15) + (1 << 9)) >> 10);
        v14[c] = (int16_t)((RV * r + GV * g + BV * b + (256 << 15) + (1 << 9)) >> 10);
      }
    } else { /* rgb24ToUV_c */
      for (int x = 0; x < W; x++) {
        int r = row[x * bpp + r_off], g = row[x * bpp + g_off], b = row[x * bpp + b_off];
        u14[x] = (int16_t)((RU * r + GU * g + BU * b + (256 << 14) + (1 << 8)) >> 9);
        v14[x] = (int16_t)((RV * r + GV * g + BV * b + (256 << 14) + (1 << 8)) >> 9);
      }
    }
    hscale16to15(py + (size_t)y * Wd, Wd, y14, hf_l, hp_l, hs_l);
    hscale16to15(pu + (size_t)y * cdW, cdW, u14, hf_c, hp_c, hs_c);
    hscale16to15(pv + (size_t)y * cdW, cdW, v14, hf_c, hp_c, hs_c);
  }
  vscale_plane(dy, ys, Wd, Hd, py, Wd, vf_l, vp_l, vs_l);
  vscale_plane(du, us, cdW, cdH, pu, cdW, vf_c, vp_c, vs_c);
  vscale_plane(dv, vs, cdW, cdH, pv, cdW, vf_c, vp_c, vs_c);

  free(y14); free(u14); free(v14); free(py); free(pu); free(pv);
  free(hf_l); free(hf_c); free(vf_l); free(vf_c);
  free(hp_l); free(hp_c); free(vp_l); free(vp_c);
  return 0;
}

/*
 * GRAY8 -> YUV420P (gray is full range in libswscale, so luma is range
 * compressed to 16..235: lumRangeFromJpeg with coeff 14071, offset 33561472 in
 * 9.1.100; U=V=128).  SURVEY.md Appendix A.4.
 */
NES_ORACLE_API int nes_oracle_gray_to_yuv420p(const uint8_t *src, int src_stride, int W, int H, int Wd,
                                              int Hd, uint8_t *dy, int ys, uint8_t *du, int us,
                                              uint8_t *dv, int vs) {
  if (W < 4 || H < 4 || Wd < 2 || Hd < 2 || (W | H | Wd | Hd) & 1) return -1;
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  int16_t *hf_l, *vf_l;
  int32_t *hp_l, *vp_l;
  const int hs_l = nes_oracle_init_filter(W, Wd, 1 << 14, &hf_l, &hp_l);
  const int vs_l = nes_oracle_init_filter(H, Hd, 1 << 12, &vf_l, &vp_l);
  int16_t *py = (int16_t *)malloc(sizeof(int16_t) * (size_t)Wd * H);
  for (int y = 0; y < H; y++) {
    int16_t *p = py + (size_t)y * Wd;
    hscale8to15(p, Wd, src + (size_t)y * src_stride, hf_l, hp_l, hs_l);
    for (int x = 0; x < Wd; x++) p[x] = (int16_t)((p[x] * 14071 + 33561472) >> 14);
  }
  vscale_plane(dy, ys, Wd, Hd, py, Wd, vf_l, vp_l, vs_l);
  for (int i = 0; i < cdH; i++) {
    memset(du + (size_t)i * us, 128, (size_t)cdW);
    memset(dv + (size_t)i * vs, 128, (size_t)cdW);
  }
  free(py); free(hf_l); free(vf_l); free(hp_l); free(vp_l);
  (void)xinc_is_unscaled;
  return 0;
}

/* Ordered dither libswscale applies in the vertical scaler when the source has more than 8 bits per sample and the
 * destination 8 (swscale.c: should_dither = is16BPS(srcFormat) -> lumDither8 = ff_dither_8x8_128[dstY & 7]; output.c
 * yuv2plane1_8_c / yuv2planeX_8_c add dither[(x + offset) & 7], <<12 for the filtered case).  The table is
 * libswscale/output.c ff_dither_8x8_128; tests/test_oracle.py re-derives every entry from the live library. */
static const uint8_t kDither8x8_128[8][8] = {
    {36, 68, 60, 92, 34, 66, 58, 90},  {100, 4, 124, 28, 98, 2, 122, 26}, {52, 84, 44, 76, 50, 82, 42, 74}, {116, 20, 108, 12, 114, 18, 106, 10},
    {32, 64, 56, 88, 38, 70, 62, 94},  {96, 0, 120, 24, 102, 6, 126, 30}, {48, 80, 40, 72, 54, 86, 46, 78}, {112, 16, 104, 8, 118, 22, 110, 14}};
NES_ORACLE_API const uint8_t *nes_oracle_dither8x8_128(void) { return &kDither8x8_128[0][0]; }

static void vscale_plane_dither(uint8_t *dst, int dst_stride, int dst_w, int dst_h, const int16_t *plane15, int plane_w, const int16_t *vf,
                                const int32_t *vpos, int vsize) {
  for (int i = 0; i < dst_h; i++) {
    uint8_t *d = dst + (size_t)i * dst_stride;
    const uint8_t *dith = kDither8x8_128[i & 7];
    if (vsize == 1) {
      const int16_t *p = plane15 + (size_t)vpos[i] * plane_w;
      for (int x = 0; x < dst_w; x++) d[x] = (uint8_t)clip8((p[x] + dith[x & 7]) >> 7);
    } else {
      for (int x = 0; x < dst_w; x++) {
        int val = dith[x & 7] << 12;
        for (int j = 0; j < vsize; j++) val += (int)plane15[(size_t)(vpos[i] + j) * plane_w + x] * vf[(size_t)i * vsize + j];
        d[x] = (uint8_t)clip8(val >> 19);
      }
    }
  }
}

/* horizontal polyphase, 16-bit gray input: hScale16To15_c with sh = depth - 1 = 15 (libswscale swscale.c; the
 * RGB-derived planes above use sh = 13) */
static void hscale_gray16to15(int16_t *dst, int dst_w, const uint16_t *src, const int16_t *filter, const int32_t *pos, int size) {
  for (int i = 0; i < dst_w; i++) {
    int val = 0;
    for (int j = 0; j < size; j++) val += (int)src[pos[i] + j] * filter[size * i + j];
    dst[i] = (int16_t)imin(val >> 15, (1 << 15) - 1);
  }
}

/*
 * GRAY16LE -> YUV420P: the 16-bit depth variant of the reference's depth stream (server.cpp:193-194 fixes GRAY8 today;
 * SURVEY.md §8 f rank 4).  Same chain as GRAY8 with the 16-bit horizontal scaler: hScale16To15 (>>15) ->
 * lumRangeFromJpeg (14071 / 33561472) -> vertical WITH libswscale's 8x8 ordered dither (a 16-bit source going to 8 bits);
 * U = V = 128.  Pinned against the real libswscale in
 * tests/test_oracle.py.
 */
NES_ORACLE_API int nes_oracle_gray16_to_yuv420p(const uint16_t *src, int src_stride_bytes, int W, int H, int Wd, int Hd, uint8_t *dy, int ys,
                                                uint8_t *du, int us, uint8_t *dv, int vs) {
  if (W < 4 || H < 4 || Wd < 2 || Hd < 2 || (W | H | Wd | Hd) & 1) return -1;
  const int cdW = (Wd + 1) >> 1, cdH = (Hd + 1) >> 1;
  int16_t *hf_l, *vf_l;
  int32_t *hp_l, *vp_l;
  const int hs_l = nes_oracle_init_filter(W, Wd, 1 << 14, &hf_l, &hp_l);
  const int vs_l = nes_oracle_init_filter(H, Hd, 1 << 12, &vf_l, &vp_l);
  int16_t *py = (int16_t *)malloc(sizeof(int16_t) * (size_t)Wd * H);
  for (int y = 0; y < H; y++) {
    int16_t *p = py + (size_t)y * Wd;
    hscale_gray16to15(p, Wd, (const uint16_t *)((const uint8_t *)src + (size_t)y * src_stride_bytes), hf_l, hp_l, hs_l);
    for (int x = 0; x < Wd; x++) p[x] = (int16_t)((p[x] * 14071 + 33561472) >> 14);
  }
  vscale_plane_dither(dy, ys, Wd, Hd, py, Wd, vf_l, vp_l, vs_l);
  for (int i = 0; i < cdH; i++) {
    memset(du + (size_t)i * us, 128, (size_t)cdW);
    memset(dv + (size_t)i * vs, 128, (size_t)cdW);
  }
  free(py); free(hf_l); free(vf_l); free(hp_l); free(vp_l);
  return 0;
}
