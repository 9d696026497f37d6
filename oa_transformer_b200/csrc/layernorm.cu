// LayerNorm forward / backward over the embedding dimension - HBM-bound row kernels, one warp per row,
// 128-bit loads, fp32 statistics by warp shuffle, values held in registers between the two passes.
//
// Reference call sites: nn.LayerNorm(eps=1e-6) norm1/norm2/norm3/norm in OATrans/model/video_transformer.py:164-174,346
// (pre-LN), and HF DistilBERT's LayerNorm(eps=1e-12) after the embeddings and after each residual add (post-LN).
// Forward writes the bf16 operand of the next GEMM (and optionally the fp32 value for a post-LN residual stream);
// backward fuses the residual-gradient adds and emits the bf16 copy the next dgrad/wgrad GEMM consumes.
#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int kLnWarps = 8;

template <int NV>  // D = NV * 128
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, long long rows, __nv_bfloat16* __restrict__ y_bf16,
                     long long ldy, float* __restrict__ y_f32, long long ldyf, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, __nv_bfloat16* __restrict__ y_split, long long ldys,
                     long long split_period) {
  constexpr int D = NV * 128;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  // split-bf16 copy [hi | hi | lo] of every split_period-th row (the CLS rows of the video tower, period = T)
  __nv_bfloat16* ys = (y_split != nullptr && row % split_period == 0) ? y_split + (row / split_period) * ldys : nullptr;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean_out != nullptr) mean_out[row] = mean;
    if (rstd_out != nullptr) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
    const float4 b = reinterpret_cast<const float4*>(beta)[i * 32 + lane];
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y_f32 != nullptr) reinterpret_cast<float4*>(y_f32 + row * ldyf)[i * 32 + lane] = o;
    if (y_bf16 != nullptr) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y_bf16 + row * ldy)[i * 32 + lane] = pk;
    }
    if (ys != nullptr) {
      uint2 hi, lo;
      hi.x = pack_bf16x2(o.x, o.y);
      hi.y = pack_bf16x2(o.z, o.w);
      const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y);
      lo.x = pack_bf16x2(o.x - h0.x, o.y - h0.y);
      lo.y = pack_bf16x2(o.z - h1.x, o.w - h1.y);
      reinterpret_cast<uint2*>(ys)[i * 32 + lane] = hi;
      reinterpret_cast<uint2*>(ys + D)[i * 32 + lane] = hi;
      reinterpret_cast<uint2*>(ys + 2 * D)[i * 32 + lane] = lo;
    }
  }
}

// dx = add1 + add2 + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  dy = dy_bf16 + dy_f32
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy   (per-lane register partials -> smem -> one atomic per column)
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy_bf16, long long lddyb, const float* __restrict__ dy_f32,
                     long long lddyf, const float* __restrict__ x, long long ldx, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, long long rows,
                     const float* __restrict__ add1, const float* __restrict__ add2, long long ldadd,
                     float* __restrict__ dx, long long lddx, __nv_bfloat16* __restrict__ dx_bf16, long long lddxb,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
  constexpr int D = NV * 128;
  __shared__ float red[kLnWarps][D];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 gm[NV], dg[NV], db[NV], ds[NV];   // ds: column sums of dx (the bias gradient of the producer GEMM)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gm[i] = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = static_cast<long long>(blockIdx.x) * kLnWarps + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = reinterpret_cast<const float4*>(x + row * ldx)[i * 32 + lane];
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dy_bf16 != nullptr) {
        const uint2 raw = reinterpret_cast<const uint2*>(dy_bf16 + row * lddyb)[i * 32 + lane];
        const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
        d = make_float4(a.x, a.y, b.x, b.y);
      }
      if (dy_f32 != nullptr) {
        const float4 f = reinterpret_cast<const float4*>(dy_f32 + row * lddyf)[i * 32 + lane];
        d.x += f.x; d.y += f.y; d.z += f.z; d.w += f.w;
      }
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rs * (g[i].x - s1 - xh[i].x * s2);
      o.y = rs * (g[i].y - s1 - xh[i].y * s2);
      o.z = rs * (g[i].z - s1 - xh[i].z * s2);
      o.w = rs * (g[i].w - s1 - xh[i].w * s2);
      if (add1 != nullptr) {
        const float4 a = reinterpret_cast<const float4*>(add1 + row * ldadd)[i * 32 + lane];
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      if (add2 != nullptr) {
        const float4 a = reinterpret_cast<const float4*>(add2 + row * ldadd)[i * 32 + lane];
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      ds[i].x += o.x; ds[i].y += o.y; ds[i].z += o.z; ds[i].w += o.w;
      if (dx != nullptr) reinterpret_cast<float4*>(dx + row * lddx)[i * 32 + lane] = o;
      if (dx_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(dx_bf16 + row * lddxb)[i * 32 + lane] = pk;
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr && dxsum == nullptr) return;
  // cross-warp reduction of the per-lane column partials, then one atomic per column per CTA
  for (int pass = 0; pass < 3; ++pass) {
    float4* src = pass == 0 ? dg : (pass == 1 ? db : ds);
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) reinterpret_cast<float4*>(red[warp])[i * 32 + lane] = src[i];
    __syncthreads();
    if (dst != nullptr) {
      for (int c = threadIdx.x; c < D; c += kLnWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) s += red[w][c];
        atomicAdd(dst + c, s);
      }
    }
  }
}

// Same arithmetic with the row operands staged through shared memory by cp.async, double-buffered per warp: while a
// warp reduces / writes row r, the 14 bytes per column of row r + stride (x, dy, the two residual-gradient addends)
// are already in flight - 86 KB of loads per SM instead of what 8 warps x one row can keep outstanding from
// registers (the register version sits at 239 registers, one 8-warp CTA per SM, ~70 % of the HBM roofline).
// Slots per stage: x fp32 | f0 fp32 (add1, or dy_f32 for the post-LN text tower) | f1 fp32 (add2) | dy bf16.
__device__ __forceinline__ void ln_cp16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32, 1)
layernorm_bwd_staged_kernel(const __nv_bfloat16* __restrict__ dy_bf16, long long lddyb, const float* __restrict__ dy_f32,
                            long long lddyf, const float* __restrict__ x, long long ldx, const float* __restrict__ mean,
                            const float* __restrict__ rstd, const float* __restrict__ gamma, long long rows,
                            const float* __restrict__ add1, const float* __restrict__ add2, long long ldadd,
                            float* __restrict__ dx, long long lddx, __nv_bfloat16* __restrict__ dx_bf16, long long lddxb,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
  constexpr int D = NV * 128;
  constexpr int kRowBytes = D * 14;
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t ln_smem[];
  float(*red)[D] = reinterpret_cast<float(*)[D]>(ln_smem);                       // [kLnWarps][D]
  float* s_gamma = reinterpret_cast<float*>(ln_smem + kLnWarps * D * 4);           // [D]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t* stage = ln_smem + (kLnWarps + 1) * D * 4 + warp * 2 * kRowBytes;
  const float* f0 = dy_f32 != nullptr ? dy_f32 : add1;
  const long long ldf0 = dy_f32 != nullptr ? lddyf : ldadd;
  const float* f1 = dy_f32 != nullptr ? nullptr : add2;
  const bool f0_is_dy = dy_f32 != nullptr;
  for (int c = threadIdx.x; c < D; c += kLnWarps * 32) s_gamma[c] = gamma[c];
  float4 dg[NV], db[NV], ds[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto issue = [&](long long row, int buf) {
    uint8_t* b = stage + buf * kRowBytes;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      ln_cp16(b + c * 16, x + row * ldx + c * 4);
      if (f0 != nullptr) ln_cp16(b + D * 4 + c * 16, f0 + row * ldf0 + c * 4);
      if (f1 != nullptr) ln_cp16(b + D * 8 + c * 16, f1 + row * ldadd + c * 4);
    }
    if (dy_bf16 != nullptr) {
      for (int c = lane; c < NV * 16; c += 32) ln_cp16(b + D * 12 + c * 16, dy_bf16 + row * lddyb + c * 8);
    }
  };
  const long long stride = static_cast<long long>(gridDim.x) * kLnWarps;
  long long row = static_cast<long long>(blockIdx.x) * kLnWarps + warp;
  if (row < rows) issue(row, 0);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  __syncthreads();                                            // s_gamma
  int buf = 0;
  for (; row < rows; row += stride, buf ^= 1) {
    if (row + stride < rows) issue(row + stride, buf ^ 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncwarp();
    const uint8_t* b = stage + buf * kRowBytes;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      const float4 xv = *reinterpret_cast<const float4*>(b + c * 16);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dy_bf16 != nullptr) {
        const uint2 raw = *reinterpret_cast<const uint2*>(b + D * 12 + c * 8);
        const float2 a = unpack_bf16x2(raw.x), bb = unpack_bf16x2(raw.y);
        d = make_float4(a.x, a.y, bb.x, bb.y);
      }
      if (f0_is_dy) {
        const float4 f = *reinterpret_cast<const float4*>(b + D * 4 + c * 16);
        d.x += f.x; d.y += f.y; d.z += f.z; d.w += f.w;
      }
      const float4 gm = *reinterpret_cast<const float4*>(s_gamma + c * 4);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      float4 o;
      o.x = rs * (g[i].x - s1 - xh[i].x * s2);
      o.y = rs * (g[i].y - s1 - xh[i].y * s2);
      o.z = rs * (g[i].z - s1 - xh[i].z * s2);
      o.w = rs * (g[i].w - s1 - xh[i].w * s2);
      if (!f0_is_dy && f0 != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(b + D * 4 + c * 16);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      if (f1 != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(b + D * 8 + c * 16);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      ds[i].x += o.x; ds[i].y += o.y; ds[i].z += o.z; ds[i].w += o.w;
      if (dx != nullptr) reinterpret_cast<float4*>(dx + row * lddx)[c] = o;
      if (dx_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(dx_bf16 + row * lddxb)[c] = pk;
      }
    }
    __syncwarp();                                             // this stage is refilled by the next iteration's prefetch
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  if (dgamma == nullptr && dbeta == nullptr && dxsum == nullptr) return;
  for (int pass = 0; pass < 3; ++pass) {
    float4* src = pass == 0 ? dg : (pass == 1 ? db : ds);
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) reinterpret_cast<float4*>(red[warp])[i * 32 + lane] = src[i];
    __syncthreads();
    if (dst != nullptr) {
      for (int c = threadIdx.x; c < D; c += kLnWarps * 32) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) acc += red[w][c];
        atomicAdd(dst + c, acc);
      }
    }
  }
}

template <int NV>
static int launch_ln_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                         long long rows, void* y_bf16, long long ldy, float* y_f32, long long ldyf, float* mean,
                         float* rstd, void* y_split, long long ldys, long long split_period, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((rows + kLnWarps - 1) / kLnWarps);
  cudaError_t e = launch_pdl(layernorm_fwd_kernel<NV>, dim3(grid), dim3(kLnWarps * 32), 0, s, x, ldx, gamma, beta, eps, rows,
                             reinterpret_cast<__nv_bfloat16*>(y_bf16), ldy, y_f32, ldyf, mean, rstd,
                             reinterpret_cast<__nv_bfloat16*>(y_split), ldys, split_period);
  if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "layernorm_fwd_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("layernorm_fwd_kernel");
}

template <int NV>
static int launch_ln_bwd(const void* dyb, long long lddyb, const float* dyf, long long lddyf, const float* x,
                         long long ldx, const float* mean, const float* rstd, const float* gamma, long long rows,
                         const float* add1, const float* add2, long long ldadd, float* dx, long long lddx,
                         void* dxb, long long lddxb, float* dgamma, float* dbeta, float* dxsum, cudaStream_t s) {
  constexpr int D = NV * 128;
  long long want = (rows + kLnWarps - 1) / kLnWarps;
  const bool aligned = ldx % 4 == 0 && (dyb == nullptr || (lddyb % 8 == 0 && (reinterpret_cast<uintptr_t>(dyb) & 15) == 0)) &&
                       (dyf == nullptr || (lddyf % 4 == 0 && (reinterpret_cast<uintptr_t>(dyf) & 15) == 0)) &&
                       ((add1 == nullptr && add2 == nullptr) || ldadd % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(add1) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(add2) & 15) == 0 && !(dyf != nullptr && (add1 != nullptr || add2 != nullptr));
  constexpr int smem = (kLnWarps + 1) * D * 4 + kLnWarps * 2 * D * 14;
  if (smem <= 227 * 1024 && aligned && rows >= 4 * kLnWarps) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(layernorm_bwd_staged_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "layernorm_bwd smem attr: %s", cudaGetErrorString(e));
      attr_done = true;
    }
    const long long cap1 = static_cast<long long>(num_sms());
    const unsigned grid1 = static_cast<unsigned>(want < cap1 ? want : cap1);
    cudaError_t e = launch_pdl(layernorm_bwd_staged_kernel<NV>, dim3(grid1), dim3(kLnWarps * 32), smem, s,
                               reinterpret_cast<const __nv_bfloat16*>(dyb), lddyb, dyf, lddyf, x, ldx, mean, rstd, gamma,
                               rows, add1, add2, ldadd, dx, lddx, reinterpret_cast<__nv_bfloat16*>(dxb), lddxb, dgamma,
                               dbeta, dxsum);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "layernorm_bwd_staged_kernel launch: %s", cudaGetErrorString(e));
    return check_launch("layernorm_bwd_staged_kernel");
  }
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  layernorm_bwd_kernel<NV><<<grid, kLnWarps * 32, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dyb), lddyb, dyf, lddyf, x, ldx, mean, rstd, gamma, rows, add1, add2,
      ldadd, dx, lddx, reinterpret_cast<__nv_bfloat16*>(dxb), lddxb, dgamma, dbeta, dxsum);
  return check_launch("layernorm_bwd_kernel");
}

}  // namespace oat

extern "C" int oat_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                 int64_t rows, int32_t D, void* y_bf16, int64_t ldy, float* y_f32, int64_t ldyf,
                                 float* mean, float* rstd, void* y_split, int64_t ldys, int64_t split_period,
                                 oat_stream_t stream) {
  using namespace oat;
  OAT_REQUIRE(rows >= 0 && D > 0 && D % 128 == 0 && D <= 1024, "oat_layernorm_fwd: D=%d must be a multiple of 128, <= 1024", D);
  OAT_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldyf % 4 == 0, "oat_layernorm_fwd: leading dims must be multiples of 4");
  OAT_REQUIRE(y_split == nullptr || (split_period >= 1 && ldys % 4 == 0 && ldys >= 3 * D),
              "oat_layernorm_fwd: y_split needs split_period >= 1 and a pitch >= 3*D that is a multiple of 4");
  if (y_split == nullptr) split_period = 1;
  if (rows == 0) return OAT_OK;
  cudaStream_t s = as_stream(stream);
  switch (D / 128) {
    case 1: return launch_ln_fwd<1>(x, ldx, gamma, beta, eps, rows, y_bf16, ldy, y_f32, ldyf, mean, rstd, y_split, ldys, split_period, s);
    case 2: return launch_ln_fwd<2>(x, ldx, gamma, beta, eps, rows, y_bf16, ldy, y_f32, ldyf, mean, rstd, y_split, ldys, split_period, s);
    case 4: return launch_ln_fwd<4>(x, ldx, gamma, beta, eps, rows, y_bf16, ldy, y_f32, ldyf, mean, rstd, y_split, ldys, split_period, s);
    case 6: return launch_ln_fwd<6>(x, ldx, gamma, beta, eps, rows, y_bf16, ldy, y_f32, ldyf, mean, rstd, y_split, ldys, split_period, s);
    case 8: return launch_ln_fwd<8>(x, ldx, gamma, beta, eps, rows, y_bf16, ldy, y_f32, ldyf, mean, rstd, y_split, ldys, split_period, s);
    default: return set_error(OAT_ERR_ARG, "oat_layernorm_fwd: unsupported D=%d (128, 256, 512, 768, 1024)", D);
  }
}

extern "C" int oat_layernorm_bwd(const void* dy_bf16, int64_t lddyb, const float* dy_f32, int64_t lddyf,
                                 const float* x, int64_t ldx, const float* mean, const float* rstd,
                                 const float* gamma, int64_t rows, int32_t D, const float* add1, const float* add2,
                                 int64_t ldadd, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb, float* dgamma,
                                 float* dbeta, float* dxsum, oat_stream_t stream) {
  using namespace oat;
  OAT_REQUIRE(rows >= 0 && D > 0 && D % 128 == 0 && D <= 1024, "oat_layernorm_bwd: D=%d must be a multiple of 128, <= 1024", D);
  OAT_REQUIRE(dy_bf16 != nullptr || dy_f32 != nullptr, "oat_layernorm_bwd: no incoming gradient");
  if (rows == 0) return OAT_OK;
  cudaStream_t s = as_stream(stream);
#define OAT_LN_BWD(NV)                                                                                              \
  return launch_ln_bwd<NV>(dy_bf16, lddyb, dy_f32, lddyf, x, ldx, mean, rstd, gamma, rows, add1, add2, ldadd, dx, \
                           lddx, dx_bf16, lddxb, dgamma, dbeta, dxsum, s)
  switch (D / 128) {
    case 1: OAT_LN_BWD(1);
    case 2: OAT_LN_BWD(2);
    case 4: OAT_LN_BWD(4);
    case 6: OAT_LN_BWD(6);
    case 8: OAT_LN_BWD(8);
    default: return set_error(OAT_ERR_ARG, "oat_layernorm_bwd: unsupported D=%d (128, 256, 512, 768, 1024)", D);
  }
#undef OAT_LN_BWD
}
