// micro-benchmark: cost of CTA barriers on sm_100a (cycles per barrier), 1024-thread CTA
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out) {
  __shared__ unsigned s_m[8];
  const int tid = threadIdx.x;
  if (tid < 8) s_m[tid] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < 100; ++i) __syncthreads();
  long long t1 = clock64();
  if (tid == 0) out[0] = (t1 - t0) / 100;
  // 8 warps on barrier 1 while 24 warps wait on barrier 0
  if (tid < 256) {
    t0 = clock64();
    for (int i = 0; i < 100; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / 100;
    t0 = clock64();
    unsigned acc = 0;
    for (int i = 0; i < 100; ++i) {
      unsigned h = 0;
      for (int w = 0; w < 8; ++w) h |= s_m[w] & (tid + i);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (h == 12345u) atomicOr(&s_m[tid >> 5], 1u << (tid & 31));
      int r;
      asm volatile("{\n .reg .pred p, q;\n setp.ne.s32 q, %1, 0;\n bar.red.or.pred p, 1, 256, q;\n selp.s32 %0, 1, 0, p;\n}\n" : "=r"(r) : "r"((int)(h == 777u)) : "memory");
      acc += r;
    }
    t1 = clock64();
    if (tid == 0) { out[2] = (t1 - t0) / 100; out[3] = acc; }
  }
  __syncthreads();
  // one warp alone: 100 dependent LDS
  if (tid < 32) {
    t0 = clock64();
    unsigned v = tid & 7;
    for (int i = 0; i < 100; ++i) v = s_m[v & 7] + (v & 7);
    t1 = clock64();
    if (tid == 0) { out[4] = (t1 - t0) / 100; out[5] = v; }
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  k<<<1, 1024>>>(d); k<<<1, 1024>>>(d);
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("syncthreads(1024) %lld cyc | bar.sync 1,256 (others parked) %lld cyc | relax sweep %lld cyc | dependent LDS %lld cyc\n", h[0], h[1], h[2], h[4]);
  return 0;
}
