// ingest.cc -- pinned receive ring: the zero-copy replacement of the reference's ingest chain
//   socket_receive_blocking_lpf   /root/reference/src/server.cpp:91-112  (recv into a std::string)
//   RenderedFrame::ParseFromString  src/server.cpp:175                    (copy #2, into the message)
//   RenderedFrame ctor              src/base/video/rendered_frame.cc:5-27 (copy #3, message moved + wrapped)
// The socket thread recv()s each length-prefixed message straight into a slot of page-locked
// memory; nes_ingest_commit locates the two `bytes` fields in place (unpack.cc) and hands back
// nes_source descriptors that point INTO the slot, so nes_gpu_submit DMAs the payload from where
// the NIC stack put it: no host memcpy of the pixels at all.  A slot stays referenced until
// nes_ingest_release (after nes_gpu_wait of the frame that used it).
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <new>
#include <vector>

#include "nes_gpu.h"

struct nes_ingest_ring {
  std::mutex mu;
  uint8_t *base = nullptr;
  uint64_t slot_bytes = 0;
  std::vector<int> state;  // 0 free, 1 acquired (receiving), 2 committed (referenced by a frame)
  int next = 0;
};

extern "C" {

int nes_ingest_ring_create(int n_slots, uint64_t slot_bytes, nes_ingest_ring **out) {
  if (!out || n_slots < 1 || n_slots > 1024 || slot_bytes < 64) return NES_ERR_INVALID_ARG;
  *out = nullptr;
  nes_ingest_ring *r = new (std::nothrow) nes_ingest_ring();
  if (!r) return NES_ERR_NO_MEMORY;
  r->slot_bytes = (slot_bytes + 4095) & ~(uint64_t)4095;
  if (cudaHostAlloc((void **)&r->base, r->slot_bytes * (uint64_t)n_slots, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    delete r;
    return NES_ERR_CUDA;
  }
  r->state.assign((size_t)n_slots, 0);
  *out = r;
  return NES_OK;
}

void nes_ingest_ring_destroy(nes_ingest_ring *r) {
  if (!r) return;
  if (r->base) { cudaFreeHost(r->base); cudaGetLastError(); }
  delete r;
}

int nes_ingest_acquire(nes_ingest_ring *r, int *slot, uint8_t **buf, uint64_t *cap) {
  if (!r || !slot || !buf || !cap) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(r->mu);
  const int n = (int)r->state.size();
  for (int i = 0; i < n; i++) {
    const int s = (r->next + i) % n;
    if (r->state[(size_t)s] == 0) {
      r->state[(size_t)s] = 1;
      r->next = (s + 1) % n;
      *slot = s;
      *buf = r->base + r->slot_bytes * (uint64_t)s;
      *cap = r->slot_bytes;
      return NES_OK;
    }
  }
  return NES_ERR_BUSY;
}

int nes_ingest_commit(nes_ingest_ring *r, int slot, uint64_t len, int has_length_prefix, int bytes_per_pixel, nes_unpacked_frame *info,
                      nes_source *src) {
  if (!r || !info || !src || bytes_per_pixel < 3 || bytes_per_pixel > 4) return NES_ERR_INVALID_ARG;
  const uint8_t *buf;
  {
    std::lock_guard<std::mutex> lk(r->mu);
    if (slot < 0 || slot >= (int)r->state.size() || r->state[(size_t)slot] != 1) return NES_ERR_INVALID_ARG;
    if (len > r->slot_bytes) return NES_ERR_TOO_LARGE;
    buf = r->base + r->slot_bytes * (uint64_t)slot;
  }
  const int st = nes_unpack_rendered_frame(buf, len, has_length_prefix, info);
  if (st != NES_OK) return st;
  // the reference trusts Camera.width/height against the payload (rendered_frame.cc:14-25); we check
  const uint64_t px = (uint64_t)info->width * (uint64_t)info->height;
  if (info->width < 1 || info->height < 1 || info->frame_len < px * (uint64_t)bytes_per_pixel) return NES_ERR_SHORT_BUFFER;
  if (info->depth_len != 0 && info->depth_len < px) return NES_ERR_SHORT_BUFFER;
  src->rgb = buf + info->frame_off;
  src->rgb_stride = 0;
  src->rgb_bytes = info->frame_len;
  src->depth = info->depth_len ? buf + info->depth_off : nullptr;
  src->depth_stride = 0;
  src->depth_bytes = info->depth_len;
  std::lock_guard<std::mutex> lk(r->mu);
  r->state[(size_t)slot] = 2;
  return NES_OK;
}

int nes_ingest_release(nes_ingest_ring *r, int slot) {
  if (!r) return NES_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(r->mu);
  if (slot < 0 || slot >= (int)r->state.size() || r->state[(size_t)slot] == 0) return NES_ERR_INVALID_ARG;
  r->state[(size_t)slot] = 0;
  return NES_OK;
}

}  // extern "C"
