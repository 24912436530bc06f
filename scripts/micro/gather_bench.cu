// Throughput of scattered (one line per lane) gathers from an L2-resident table on one GPU, by
// instruction kind: what a lane-gather costs the SM's L1TEX path.  Build & run:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

struct __align__(32) D4 { double a, b, c, d; };
__device__ __forceinline__ D4 ldg256(const void* p) {
  D4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// MODE 0: LDG.64   1: LDG.128   2: LDG.256   3: 2 x LDG.256 of one 64-byte row
//      4: LDGSTS.128 + LDS.128   5: 4 x LDGSTS.128 of one 64-byte row + 4 LDS.128   6: no load
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const char* tab, uint32_t mask64, int iters, double* out) {
  extern __shared__ __align__(16) unsigned char sm[];
  uint4* stg = reinterpret_cast<uint4*>(sm) + threadIdx.x;
  double acc = 0.0;
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    s = mix(s + 0x9e3779b9u);
    const char* p = tab + (size_t)(s & mask64) * 64;   // a random 64-byte row
    if (MODE == 0) { acc += __ldg(reinterpret_cast<const double*>(p)); }
    if (MODE == 1) { const double2 v = __ldg(reinterpret_cast<const double2*>(p)); acc += v.x + v.y; }
    if (MODE == 2) { const D4 v = ldg256(p); acc += v.a + v.d; }
    if (MODE == 3) { const D4 v = ldg256(p), w = ldg256(p + 32); acc += v.a + w.d; }
    if (MODE == 4 || MODE == 5) {
      const int n = MODE == 4 ? 1 : 4;
      for (int c = 0; c < n; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(stg + c * 1024)), "l"(p + 16 * c) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      for (int c = 0; c < n; ++c) { const uint4 v = stg[c * 1024]; acc += (double)v.x; }
    }
    if (MODE == 6) acc += (double)(s & 7);
  }
  if (acc == 1.2345) out[0] = acc;
}

template <int MODE>
float run(const char* tab, uint32_t mask, int iters, double* out, int sms) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<sms, 1024, 65536>>>(tab, mask, iters, out);
  cudaEventRecord(a);
  k<MODE><<<sms, 1024, 65536>>>(tab, mask, iters, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const char* names[7] = {"LDG.64", "LDG.128", "LDG.256", "2xLDG.256 (64 B row)", "LDGSTS.128+LDS", "4xLDGSTS.128+4LDS (64 B row)", "no load"};
  for (int mb : {2, 16, 64}) {
    const size_t bytes = (size_t)mb << 20;
    char* tab; double* out;
    cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes); cudaMalloc(&out, 8);
    const uint32_t mask = (uint32_t)(bytes / 64 - 1);
    const int iters = 2000;
    float t[7];
    t[0] = run<0>(tab, mask, iters, out, sms); t[1] = run<1>(tab, mask, iters, out, sms);
    t[2] = run<2>(tab, mask, iters, out, sms); t[3] = run<3>(tab, mask, iters, out, sms);
    t[4] = run<4>(tab, mask, iters, out, sms); t[5] = run<5>(tab, mask, iters, out, sms);
    t[6] = run<6>(tab, mask, iters, out, sms);
    for (int m = 0; m < 7; ++m) {
      const double rows = (double)sms * 1024 * iters;
      printf("{\"table_MB\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"ns_per_row_per_SM\": %.4f, \"rows_per_s\": %.4g}\n", mb,
             names[m], t[m], 1e6 * t[m] / (1024.0 * iters), rows / (t[m] * 1e-3));
    }
    cudaFree(tab); cudaFree(out);
  }
  return 0;
}
