// Fused multi-tensor AdamW with the semantics of transformers.AdamW (the optimizer the reference builds through
// config.initialize('optimizer', transformers, ...), OATrans/train_dist_multi.py:66; transformers 4.6 optimization.py):
//   m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;  p -= step_size * m / (sqrt(v) + eps) ;  p -= lr * wd * p
//   step_size = lr * sqrt(1 - b2^t) / (1 - b1^t) when correct_bias, else lr.
// One launch updates every parameter tensor: a device-resident table gives (p, g, m, v) pointers and a prefix sum of
// 1024-element chunks; CTAs grid-stride over the chunks (binary search chunk -> tensor). HBM-bound: 28 B per parameter.
#include "oat_host.h"

namespace oat {

constexpr int kOptChunk = 1024;   // elements per chunk = 256 threads x float4

__global__ void __launch_bounds__(256) adamw_multi_kernel(const long long* __restrict__ table,      // [n][4]: p, g, m, v
                                                         const long long* __restrict__ chunk_prefix,  // [n + 1]
                                                         const long long* __restrict__ sizes,         // [n]
                                                         int n, long long total_chunks, float lr, float b1, float b2,
                                                         float omb1, float omb2, float eps, float wd, float step_size) {
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n - 1;                       // last tensor whose first chunk is <= chunk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    float* p = reinterpret_cast<float*>(table[4 * lo + 0]);
    const float* g = reinterpret_cast<const float*>(table[4 * lo + 1]);
    float* m = reinterpret_cast<float*>(table[4 * lo + 2]);
    float* v = reinterpret_cast<float*>(table[4 * lo + 3]);
    const long long size = sizes[lo];
    const long long base = (chunk - chunk_prefix[lo]) * kOptChunk + threadIdx.x * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 && base + 4 <= size;
    float pv[4], gv[4], mv[4], vv[4];
    if (vec) {
      const float4 a = *reinterpret_cast<const float4*>(p + base), b = *reinterpret_cast<const float4*>(g + base),
                   c = *reinterpret_cast<const float4*>(m + base), d = *reinterpret_cast<const float4*>(v + base);
      pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
      mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w; vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = base + i < size;
        pv[i] = ok ? p[base + i] : 0.f; gv[i] = ok ? g[base + i] : 0.f;
        mv[i] = ok ? m[base + i] : 0.f; vv[i] = ok ? v[base + i] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mv[i] = b1 * mv[i] + omb1 * gv[i];
      vv[i] = b2 * vv[i] + omb2 * gv[i] * gv[i];
      pv[i] = pv[i] - step_size * (mv[i] / (sqrtf(vv[i]) + eps));
      if (wd > 0.f) pv[i] = pv[i] - lr * wd * pv[i];
    }
    if (vec) {
      *reinterpret_cast<float4*>(p + base) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      *reinterpret_cast<float4*>(m + base) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v + base) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (base + i < size) { p[base + i] = pv[i]; m[base + i] = mv[i]; v[base + i] = vv[i]; }
      }
    }
  }
}

}  // namespace oat

extern "C" int oat_adamw_multi(const int64_t* table, const int64_t* chunk_prefix, const int64_t* sizes, int32_t n,
                               int64_t total_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                               int32_t step, int32_t correct_bias, oat_stream_t stream) {
  using namespace oat;
  OAT_REQUIRE(table != nullptr && chunk_prefix != nullptr && sizes != nullptr && n > 0 && total_chunks > 0 && step >= 1,
              "oat_adamw_multi: bad arguments (n=%d chunks=%lld step=%d)", n, (long long)total_chunks, step);
  float step_size = lr;
  if (correct_bias) {
    const double c1 = 1.0 - pow(static_cast<double>(beta1), step), c2 = 1.0 - pow(static_cast<double>(beta2), step);
    step_size = static_cast<float>(lr * sqrt(c2) / c1);
  }
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const unsigned grid = static_cast<unsigned>(total_chunks < cap ? total_chunks : cap);
  adamw_multi_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(table),
                                                          reinterpret_cast<const long long*>(chunk_prefix),
                                                          reinterpret_cast<const long long*>(sizes), n, total_chunks, lr,
                                                          beta1, beta2, static_cast<float>(1.0 - static_cast<double>(beta1)),
                                                          static_cast<float>(1.0 - static_cast<double>(beta2)), eps, weight_decay, step_size);
  return check_launch("adamw_multi_kernel");
}
