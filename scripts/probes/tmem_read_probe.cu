// Probe: TMEM -> register read bandwidth per SM (tcgen05.ld), by warp count and shape.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../oa_transformer_b200/csrc/oat_ptx.cuh"
using namespace oat;

// mode 0: 32x32b.x32   mode 1: 32x32b.x16   mode 2: 16x256b.x4 (two per 32 lanes)
__global__ void __launch_bounds__(512, 1) probe(int nwarps, int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = slot;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = ((it + warp) * 32) & 255;
      if (mode == 0) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t + lane_off + col, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc += __uint_as_float(v[e]);
      } else if (mode == 1) {
        uint32_t v[16], w[16];
        tmem_ld_32x32b_x16(t + lane_off + col, v);
        tmem_ld_32x32b_x16(t + lane_off + col + 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc += __uint_as_float(v[e]) + __uint_as_float(w[e]);
      } else {
        uint32_t v[16], w[16];
        tmem_ld_16x256b_x4(t + lane_off + col, v);
        tmem_ld_16x256b_x4(t + lane_off + (16u << 16) + col, w);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc += __uint_as_float(v[e]) + __uint_as_float(w[e]);
      }
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  if (acc == 1234.5f) *sink = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(t); }
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 16 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8, 12, 16}) {
      probe<<<1, 512>>>(nw, mode, iters, d, sink);
      probe<<<1, 512>>>(nw, mode, iters, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      const double bytes = double(nw) * iters * 32 * 32 * 4;
      printf("mode %d (%s) warps %2d: %lld clk, %.1f B/clk per SM, %.1f clk per 4 KB warp-load\n", mode,
             mode == 0 ? "32x32b.x32" : mode == 1 ? "2 x 32x32b.x16" : "2 x 16x256b.x4", nw, mx, bytes / mx, double(mx) / iters);
    }
  return 0;
}
