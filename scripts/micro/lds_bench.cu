// Cost of scattered 16-byte shared-memory look-ups (LDS.128, one random record per lane) by
// table layout: which lanes share a pass decides what replication can remove.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o lds_bench lds_bench.cu && ./lds_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// LAYOUT 0: records 32 bytes apart, first half read (the production curve tables)
//        1: records 16 bytes apart
//        2: eight copies, record r of copy g at chunk 8 r + g, g = lane & 7
//        3: sixteen copies ... g = lane & 15 (chunk 16 r + g)
//        4: 32 copies, g = lane (no two lanes ever share a bank group)
//        5: every lane the same record (broadcast)
//        6: no load
template <int LAYOUT>
__global__ void __launch_bounds__(1024, 1) k(int iters, int nrec, double* out) {
  extern __shared__ __align__(16) double2 tab[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = make_double2((double)i, 1.0);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    s = mix(s + 0x9e3779b9u);
    const int r = (int)(s % (uint32_t)nrec);
    int idx;
    if (LAYOUT == 0) idx = 2 * r;
    if (LAYOUT == 1) idx = r;
    if (LAYOUT == 2) idx = 8 * r + (lane & 7);
    if (LAYOUT == 3) idx = 16 * r + (lane & 15);
    if (LAYOUT == 4) idx = 32 * r + lane;
    if (LAYOUT == 5) idx = (int)(mix(it) % (uint32_t)nrec);
    if (LAYOUT == 6) { acc += (double)(r & 3); continue; }
    const double2 v = tab[idx];
    acc += v.x + v.y;
  }
  if (acc == 1.2345) out[0] = acc;
}

template <int L>
float run(int iters, int nrec, double* out, int sms) {
  cudaFuncSetAttribute(k<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<L><<<sms, 1024, 8192 * 16>>>(iters, nrec, out);
  cudaEventRecord(a);
  k<L><<<sms, 1024, 8192 * 16>>>(iters, nrec, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8);
  const int iters = 4000, nrec = 250;      // 250 records: the size of a curve's bucket table
  const char* names[7] = {"32-byte stride (production)", "16-byte stride", "8 copies (lane & 7)",
                          "16 copies (lane & 15)", "32 copies (lane)", "broadcast", "no load"};
  float t[7] = {run<0>(iters, nrec, out, sms), run<1>(iters, nrec, out, sms), run<2>(iters, nrec, out, sms),
                run<3>(iters, nrec, out, sms), run<4>(iters, nrec, out, sms), run<5>(iters, nrec, out, sms),
                run<6>(iters, nrec, out, sms)};
  for (int m = 0; m < 7; ++m)
    printf("{\"layout\": \"%s\", \"ms\": %.3f, \"cycles_per_warp_LDS128\": %.2f}\n", names[m], t[m],
           (t[m] - t[6]) * 1e-3 * 1.965e9 / (32.0 * iters));
  return 0;
}
