// How long does nanosleep.u32 N really suspend a warp on this GPU?  (diagnostic; build: nvcc -gencode arch=compute_100a,code=sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned n, unsigned long long *out) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < 16; i++) asm volatile("nanosleep.u32 %0;" ::"r"(n));
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / 16;
}
int main() {
  unsigned long long *d, h[4];
  cudaMalloc(&d, sizeof(h));
  for (unsigned n : {0u, 32u, 64u, 128u, 256u, 512u, 1024u, 2048u, 4096u, 16384u, 65536u}) {
    probe<<<4, 32>>>(n, d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("nanosleep %6u -> %llu %llu %llu %llu ns per call\n", n, h[0], h[1], h[2], h[3]);
  }
  return 0;
}
