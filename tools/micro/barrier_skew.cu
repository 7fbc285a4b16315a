// micro-benchmark: wake-up latency of warps parked at a CTA barrier when the LAST warp arrives late (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int work) {
  __shared__ long long s_arrive;
  __shared__ volatile int s_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_flag = 0;
  __syncthreads();
  // --- test 1: __syncthreads, warp 0 arrives `work` cycles late
  long long lat0 = 0, lat1 = 0;
  for (int rep = 0; rep < 20; ++rep) {
    if (warp == 0) {
      long long t = clock64();
      while (clock64() - t < work) { }
      if (lane == 0) s_arrive = clock64();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) lat0 += t1 - s_arrive;
    if (tid == 32 * 20) lat1 += t1 - s_arrive;
    __syncthreads();
  }
  if (tid == 0) out[0] = lat0 / 20;
  if (tid == 32 * 20) out[1] = lat1 / 20;
  // --- test 2: named barrier among 8 warps, warp 0 late, others (8..31) parked at barrier 0
  lat0 = lat1 = 0;
  if (tid < 256) {
    for (int rep = 0; rep < 20; ++rep) {
      if (warp == 0) {
        long long t = clock64();
        while (clock64() - t < work) { }
        if (lane == 0) s_arrive = clock64();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const long long t1 = clock64();
      if (tid == 0) lat0 += t1 - s_arrive;
      if (tid == 32 * 5) lat1 += t1 - s_arrive;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (tid == 0) out[2] = lat0 / 20;
    if (tid == 32 * 5) out[3] = lat1 / 20;
  }
  __syncthreads();
  // --- test 3: spin on a shared flag with __nanosleep, warp 0 is the producer
  lat1 = 0;
  for (int rep = 1; rep <= 20; ++rep) {
    if (warp == 0) {
      long long t = clock64();
      while (clock64() - t < work) { }
      if (lane == 0) { s_arrive = clock64(); __threadfence_block(); s_flag = rep; }
    } else {
      while (s_flag != rep) __nanosleep(32);
    }
    const long long t1 = clock64();
    if (tid == 32 * 20) lat1 += t1 - s_arrive;
    __syncthreads();
  }
  if (tid == 32 * 20) out[4] = lat1 / 20;
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int work : {0, 500, 2000, 10000}) {
    cudaMemset(d, 0, 64);
    k<<<1, 1024>>>(d, work); k<<<1, 1024>>>(d, work);
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("late by %5d cyc: syncthreads release: late warp %lld, parked warp %lld | bar.sync 1,256: late warp %lld, parked warp %lld | flag spin+nanosleep: %lld cyc\n",
           work, h[0], h[1], h[2], h[3], h[4]);
  }
  return 0;
}
