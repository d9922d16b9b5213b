// depth16.cu -- k_depth16_*: the depth stream from 16-bit samples (GRAY16LE -> YUV420P).
//
// The reference fixes the depth stream to GRAY8 (/root/reference/src/server.cpp:193-194); a 16-bit depth input is
// the "next" item of SURVEY.md §8 f (rank 4).  It is libswscale's same chain as GRAY8 with the 16-bit scaler
// (hScale16To15_c, sh = 15) and -- because the source has more than 8 bits -- the 8x8 ordered dither of the vertical
// scaler (swscale.c should_dither, output.c ff_dither_8x8_128) instead of the constant 64:
//   x15 = min((sum_j s[hpos+j] * hf[j]) >> 15, 32767) -> (x15 * 14071 + 33561472) >> 14 ->
//   out = clip8(((d[y&7][x&7] << 12) + sum_j p[vpos+j] * vf[j]) >> 19)        (same size: clip8((p + d) >> 7))
// U = V = 128.  Oracle: oracle/swscale_port.c nes_oracle_gray16_to_yuv420p (pinned to the real library).
//
// Scope: frames with ONE source (a 16-bit depth composite is not defined).  The scene of such a frame goes through
// the usual kernels with its depth stream switched off; these two small kernels then write the depth image.  They are
// plain streaming kernels (the same-size one moves 3 bytes per pixel: 16 B loads, 8 B stores).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "device_common.cuh"
#include "nes_internal.h"

namespace nes {

namespace {

__constant__ uint8_t c_dither[8][8] = {{36, 68, 60, 92, 34, 66, 58, 90},  {100, 4, 124, 28, 98, 2, 122, 26}, {52, 84, 44, 76, 50, 82, 42, 74},
                                       {116, 20, 108, 12, 114, 18, 106, 10}, {32, 64, 56, 88, 38, 70, 62, 94},  {96, 0, 120, 24, 102, 6, 126, 30},
                                       {48, 80, 40, 72, 54, 86, 46, 78},   {112, 16, 104, 8, 118, 22, 110, 14}};

__device__ __forceinline__ int range15(int x15) { return (x15 * 14071 + 33561472) >> 14; }

// same size: one thread = 8 adjacent pixels of one row
__global__ void __launch_bounds__(256) k_depth16_same(const uint8_t *__restrict__ src, int src_stride, int W, int H, uint8_t *__restrict__ dy, int dys,
                                                      uint8_t *__restrict__ du, uint8_t *__restrict__ dv, int dus, int dvs, int nv12) {
  const int groups = (W + 7) >> 3;
  const int total = groups * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / groups, g = idx - y * groups, x = 8 * g;
    const uint16_t *row = (const uint16_t *)(src + (size_t)y * src_stride) + x;
    uint16_t s[8];
    if (x + 8 <= W && ((((uintptr_t)row) & 15) == 0)) {
      const uint4 q = __ldg((const uint4 *)row);
      s[0] = q.x & 0xFFFF; s[1] = q.x >> 16; s[2] = q.y & 0xFFFF; s[3] = q.y >> 16; s[4] = q.z & 0xFFFF; s[5] = q.z >> 16; s[6] = q.w & 0xFFFF; s[7] = q.w >> 16;
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) s[k] = x + k < W ? __ldg(row + k) : 0;
    }
    uint8_t o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = (uint8_t)clip8((range15(min((int)s[k] >> 1, 32767)) + c_dither[y & 7][k]) >> 7);  // x & 7 == k
    uint8_t *out = dy + (size_t)y * dys + x;
    if (x + 8 <= W && ((((uintptr_t)out) & 7) == 0)) {
      *(uint2 *)out = make_uint2(o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24), o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24));
    } else {
      for (int k = 0; k < 8 && x + k < W; k++) out[k] = o[k];
    }
    // chroma of the depth image is constant 128: the thread of an even row fills its 4 chroma columns
    if (!(y & 1)) {
      const int cy = y >> 1, cx = x >> 1, cw = (W + 1) >> 1;
      for (int k = 0; k < 4 && cx + k < cw; k++) {
        if (nv12) { du[(size_t)cy * dus + 2 * (cx + k)] = 128; du[(size_t)cy * dus + 2 * (cx + k) + 1] = 128; }
        else { du[(size_t)cy * dus + cx + k] = 128; dv[(size_t)cy * dvs + cx + k] = 128; }
      }
    }
  }
}

// resize: one thread = one destination pixel (vsize x hsize taps straight from global memory / L2)
__global__ void __launch_bounds__(256) k_depth16_resize(const uint8_t *__restrict__ src, int src_stride, int Wd, int Hd, DevFilter hl, DevFilter vl,
                                                        uint8_t *__restrict__ dy, int dys, uint8_t *__restrict__ du, uint8_t *__restrict__ dv, int dus, int dvs,
                                                        int nv12) {
  const int total = Wd * Hd;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / Wd, x = idx - y * Wd;
    const int hpos = __ldg(hl.pos + x), vpos = __ldg(vl.pos + y);
    const int16_t *hf = hl.coef + (size_t)x * hl.size, *vf = vl.coef + (size_t)y * vl.size;
    const int d = c_dither[y & 7][x & 7];
    int acc = d << 12, one = 0;
    for (int j = 0; j < vl.size; j++) {
      const uint16_t *row = (const uint16_t *)(src + (size_t)(vpos + j) * src_stride) + hpos;
      int v = 0;
      for (int i = 0; i < hl.size; i++) v += (int)__ldg(row + i) * (int)__ldg(hf + i);
      const int p = range15(min(v >> 15, 32767));
      acc += p * (int)__ldg(vf + j);
      one = p;
    }
    dy[(size_t)y * dys + x] = (uint8_t)clip8(vl.size == 1 ? (one + d) >> 7 : acc >> 19);
    if (!(y & 1) && !(x & 1)) {
      const int cy = y >> 1, cx = x >> 1;
      if (nv12) { du[(size_t)cy * dus + 2 * cx] = 128; du[(size_t)cy * dus + 2 * cx + 1] = 128; }
      else { du[(size_t)cy * dus + cx] = 128; dv[(size_t)cy * dvs + cx] = 128; }
    }
  }
}

}  // namespace

int launch_depth16(const DevJob *jobs_host, int n_jobs, void *stream) {
  int launches = 0;
  for (int j = 0; j < n_jobs; j++) {
    const DevJob &jb = jobs_host[j];
    if (!jb.d16_src) continue;
    cudaStream_t st = (cudaStream_t)stream;
    if (jb.W == jb.Wd && jb.H == jb.Hd) {
      const int total = ((jb.W + 7) >> 3) * jb.H;
      k_depth16_same<<<std::min((total + 255) / 256, 148 * 8), 256, 0, st>>>(jb.d16_src, jb.d16_stride, jb.W, jb.H, jb.d16_y, jb.d16_ys, jb.d16_u, jb.d16_v, jb.d16_us, jb.d16_vs, jb.nv12);
    } else {
      const long total = (long)jb.Wd * jb.Hd;
      k_depth16_resize<<<(int)std::min<long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(jb.d16_src, jb.d16_stride, jb.Wd, jb.Hd, jb.hl, jb.vl, jb.d16_y, jb.d16_ys, jb.d16_u, jb.d16_v,
                                                                                          jb.d16_us, jb.d16_vs, jb.nv12);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    launches++;
  }
  return launches;
}

}  // namespace nes
