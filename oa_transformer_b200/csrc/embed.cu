// Data-movement kernels around the contractions: fp32 -> bf16 operand packing, patch extraction (im2col for the
// k16/s16 patch "convolution"), token assembly (CLS + positional + temporal embeddings, frame-major patch tokens
// followed by that frame's object tokens), their gradients, column sums for bias gradients, and the DistilBERT
// embedding lookup / scatter. All HBM-bound: 128-bit accesses, grid sized by rows, no shared-memory staging needed
// (every element is touched once).
//
// Reference call sites: VideoPatchEmbed.forward OATrans/model/video_transformer.py:71-76; forward_features
// :303-325 (token index 1 + f*n + i, pos_embed[1+i] tiled over frames, temporal_embed[f] repeated inside a frame);
// object_embed OATrans/model/oa_video_transformer_region.py:250,257-261; HF DistilBERT Embeddings.
#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

// ------------------------------------------------------------------------------------------------ cast / pack
__global__ void cast_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                 long long ldd, long long rows, int cols, int cols_padded, int relu) {
  const int vec_per_row = cols_padded >> 2;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec_per_row;
    const int c = static_cast<int>(idx - r * vec_per_row) << 2;
    float v[4];
    const float* s = src + r * lds + c;
    if (c + 3 < cols && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
      const float4 f = *reinterpret_cast<const float4*>(s);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (c + k < cols) ? s[k] : 0.f;
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    uint2 pk;
    pk.x = pack_bf16x2(v[0], v[1]);
    pk.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(dst + r * ldd + c) = pk;
  }
}

// Split-bf16 activation operand: dst[r] = [hi | hi | lo] (each `cols` wide), hi = bf16(x), lo = bf16(x - hi), optional
// ReLU first. Against a weight row [hi | lo | hi] one K-concatenated GEMM yields hi.hi + hi.lo + lo.hi.
__global__ void split3_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                   long long ldd, long long rows, int cols, int relu) {
  const int vec_per_row = cols >> 2;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec_per_row;
    const int c = static_cast<int>(idx - r * vec_per_row) << 2;
    float4 f = *reinterpret_cast<const float4*>(src + r * lds + c);
    if (relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
    uint2 hi, lo;
    hi.x = pack_bf16x2(f.x, f.y);
    hi.y = pack_bf16x2(f.z, f.w);
    const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y);
    lo.x = pack_bf16x2(f.x - h0.x, f.y - h0.y);
    lo.y = pack_bf16x2(f.z - h1.x, f.w - h1.y);
    __nv_bfloat16* d = dst + r * ldd + c;
    *reinterpret_cast<uint2*>(d) = hi;
    *reinterpret_cast<uint2*>(d + cols) = hi;
    *reinterpret_cast<uint2*>(d + 2 * cols) = lo;
  }
}

// Element dropout (text tower, training mode): see include/oat.h. 4 elements per thread.
__global__ void dropout_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ residual,
                                   long long ldr, float* __restrict__ out, long long ldo,
                                   __nv_bfloat16* __restrict__ out_bf16, long long ldob,
                                   __nv_bfloat16* __restrict__ out3, long long ld3, long long rows, int cols, float inv_keep,
                                   uint32_t thresh, uint64_t seed, uint32_t site) {
  const int vec_per_row = cols >> 2;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec_per_row;
    const int c = static_cast<int>(idx - r * vec_per_row) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = dropout_keep(seed, site, static_cast<uint64_t>(r) * cols + c + k, thresh) ? o[k] * inv_keep : 0.f;
    if (residual != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(residual + r * ldr + c);
      o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
    }
    if (out != nullptr) *reinterpret_cast<float4*>(out + r * ldo + c) = make_float4(o[0], o[1], o[2], o[3]);
    uint2 hi;
    hi.x = pack_bf16x2(o[0], o[1]);
    hi.y = pack_bf16x2(o[2], o[3]);
    if (out_bf16 != nullptr) *reinterpret_cast<uint2*>(out_bf16 + r * ldob + c) = hi;
    if (out3 != nullptr) {
      const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y);
      uint2 lo;
      lo.x = pack_bf16x2(o[0] - h0.x, o[1] - h0.y);
      lo.y = pack_bf16x2(o[2] - h1.x, o[3] - h1.y);
      __nv_bfloat16* d = out3 + r * ld3 + c;
      *reinterpret_cast<uint2*>(d) = hi;
      *reinterpret_cast<uint2*>(d + cols) = hi;
      *reinterpret_cast<uint2*>(d + 2 * cols) = lo;
    }
  }
}

__global__ void dropout_bwd_kernel(const float* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ dyb,
                                   long long lddyb, float* __restrict__ dx, long long lddx,
                                   __nv_bfloat16* __restrict__ dxb, long long lddxb, long long rows, int cols,
                                   float inv_keep, uint32_t thresh, uint64_t seed, uint32_t site) {
  const int vec_per_row = cols >> 2;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec_per_row;
    const int c = static_cast<int>(idx - r * vec_per_row) << 2;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (dy != nullptr) {
      const float4 v = *reinterpret_cast<const float4*>(dy + r * lddy + c);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    if (dyb != nullptr) {
      const uint2 raw = *reinterpret_cast<const uint2*>(dyb + r * lddyb + c);
      const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
      o[0] += a.x; o[1] += a.y; o[2] += b.x; o[3] += b.y;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = dropout_keep(seed, site, static_cast<uint64_t>(r) * cols + c + k, thresh) ? o[k] * inv_keep : 0.f;
    if (dx != nullptr) *reinterpret_cast<float4*>(dx + r * lddx + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (dxb != nullptr) {
      uint2 pk;
      pk.x = pack_bf16x2(o[0], o[1]);
      pk.y = pack_bf16x2(o[2], o[3]);
      *reinterpret_cast<uint2*>(dxb + r * lddxb + c) = pk;
    }
  }
}

__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, long long n, uint32_t thresh, uint64_t seed, uint32_t site) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x)
    keep[idx] = dropout_keep(seed, site, static_cast<uint64_t>(idx), thresh) ? 1 : 0;
}

// Weight-gradient GEMMs run against the activation matrix extended by a column of ones ([tokens, cols | 1 0 ... 0]): the
// product dY^T [X | 1] carries the bias gradient (column sums of dY) in column `cols` for free - the tensor core reads dY
// anyway, a separate column-sum kernel would read it again. This kernel moves the result out of the fp32 scratch
// [rows, ld]: dW += scratch[:, :cols], db += scratch[:, cols], and re-zeroes the scratch for the next accumulate.
__global__ void unpack_wgrad_kernel(float* __restrict__ scratch, long long ld, int cols, float* __restrict__ dw,
                                    long long ldw, float* __restrict__ db, long long rows) {
  const int vec_per_row = (cols >> 2) + 1;            // cols / 4 weight vectors + the vector that starts with the bias sum
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec_per_row;
    const int v = static_cast<int>(idx - r * vec_per_row);
    float4* sp = reinterpret_cast<float4*>(scratch + r * ld) + v;
    const float4 x = *sp;
    *sp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < (cols >> 2)) {
      float4* wp = reinterpret_cast<float4*>(dw + r * ldw) + v;
      float4 w = *wp;
      w.x += x.x; w.y += x.y; w.z += x.z; w.w += x.w;
      *wp = w;
    } else {
      db[r] += x.x;
    }
  }
}

// Many casts in one launch (the bf16 operand copies of every weight of a tower, re-packed each step): a device table
// row = {src, dst, rows, cols, cols_padded, src pitch, dst pitch, kind}, chunk_prefix = prefix sum of the
// 1024-element chunks of each tensor's padded (rows x cols_padded) extent. kind 0: bf16 copy; 1: plain fp32 copy (bias
// packing); 2: split-bf16 weight copy, dst row = [hi | lo | hi] with segments cols_padded wide (pitch >= 3 * cols_padded).
__global__ void __launch_bounds__(256) cast_multi_kernel(const long long* __restrict__ table,
                                                        const long long* __restrict__ chunk_prefix, int n,
                                                        long long total_chunks) {
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const long long* t = table + 8 * lo;
    const float* src = reinterpret_cast<const float*>(t[0]);
    const long long rows = t[2], cols = t[3], colsp = t[4], lds = t[5], ldd = t[6];
    const long long idx = (chunk - chunk_prefix[lo]) * 1024 + threadIdx.x * 4;      // over rows x cols_padded
    if (idx >= rows * colsp) continue;
    const long long r = idx / colsp;
    const long long c = idx - r * colsp;                                            // multiple of 4 (cols_padded % 4 == 0)
    float v[4];
    const float* sp = src + r * lds + c;
    if (c + 3 < cols && ((reinterpret_cast<uintptr_t>(sp) & 15) == 0)) {
      const float4 f = *reinterpret_cast<const float4*>(sp);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (c + k < cols) ? sp[k] : 0.f;
    }
    if (t[7] == 1) {
      float* d = reinterpret_cast<float*>(t[1]) + r * ldd + c;
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k] = v[k];
    } else if (t[7] == 2) {
      uint2 hi, lo;
      hi.x = pack_bf16x2(v[0], v[1]);
      hi.y = pack_bf16x2(v[2], v[3]);
      const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y);
      lo.x = pack_bf16x2(v[0] - h0.x, v[1] - h0.y);
      lo.y = pack_bf16x2(v[2] - h1.x, v[3] - h1.y);
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(t[1]) + r * ldd + c;
      *reinterpret_cast<uint2*>(d) = hi;
      *reinterpret_cast<uint2*>(d + colsp) = lo;
      *reinterpret_cast<uint2*>(d + 2 * colsp) = hi;
    } else {
      uint2 pk;
      pk.x = pack_bf16x2(v[0], v[1]);
      pk.y = pack_bf16x2(v[2], v[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(t[1]) + r * ldd + c) = pk;
    }
  }
}

// dx[r, c] = relu_mask(x[r, c]) * dy_bf16[r, c]  (fp32 out) - backward of the ReLU in txt_proj (oa_model.py:68)
__global__ void relu_bwd_kernel(const float* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ dy,
                                long long lddy, float* __restrict__ dx, long long lddx, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / cols;
    const int c = static_cast<int>(idx - r * cols);
    dx[r * lddx + c] = x[r * ldx + c] > 0.f ? __bfloat162float(dy[r * lddy + c]) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ patches
// video fp32 [BF, C, H, W] -> bf16 [BF*gh*gw, C*P*P], k = c*P*P + i*P + j (the flattened Conv2d weight order)
__global__ void im2col_kernel(const float* __restrict__ video, __nv_bfloat16* __restrict__ out, long long BF, int C,
                              int H, int W, int P) {
  const int gw = W / P, gh = H / P;
  const int w8 = W >> 3;
  const long long total = BF * C * H * w8;
  const int K = C * P * P;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x8 = static_cast<int>(idx % w8);
    long long rest = idx / w8;
    const int y = static_cast<int>(rest % H); rest /= H;
    const int c = static_cast<int>(rest % C);
    const long long bf = rest / C;
    const float* s = video + ((bf * C + c) * H + y) * static_cast<long long>(W) + x8 * 8;
    const float4 a = *reinterpret_cast<const float4*>(s);
    const float4 b = *reinterpret_cast<const float4*>(s + 4);
    const int x = x8 * 8;
    const int gx = x / P, j = x - gx * P, gy = y / P, i = y - gy * P;
    uint4 pk;
    pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
    pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + ((bf * gh + gy) * gw + gx) * static_cast<long long>(K) + c * P * P + i * P + j) = pk;
  }
}

// ------------------------------------------------------------------------------------------------ token assembly
// x[b, 0]           = cls_token + pos_embed[0]
// x[b, 1+f*n+i]     = patch[(b*F+f)*N+i] + pos_embed[1+i] + temporal[f] (+ type[0])          i <  N
// x[b, 1+f*n+N+o]   = object[(b*F+f)*O+o] + temporal[f] (+ type[1])                          o <  O
__global__ void assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ object,
                                       const float* __restrict__ cls_token, const float* __restrict__ pos_embed,
                                       const float* __restrict__ temporal, const float* __restrict__ type_embed,
                                       float* __restrict__ x, int B, int F, int N, int O, int D) {
  const int n = N + O, T = 1 + F * n, d4 = D >> 2;
  const long long total = static_cast<long long>(B) * T * d4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4) << 2;
    const long long bt = idx / d4;
    const int tok = static_cast<int>(bt % T);
    const int b = static_cast<int>(bt / T);
    float4 v;
    if (tok == 0) {
      const float4 a = *reinterpret_cast<const float4*>(cls_token + c);
      const float4 p = *reinterpret_cast<const float4*>(pos_embed + c);
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else {
      const int f = (tok - 1) / n, i = (tok - 1) - f * n;
      const float4 te = *reinterpret_cast<const float4*>(temporal + static_cast<long long>(f) * D + c);
      if (i < N) {
        const float4 a = *reinterpret_cast<const float4*>(patch + ((static_cast<long long>(b) * F + f) * N + i) * D + c);
        const float4 p = *reinterpret_cast<const float4*>(pos_embed + static_cast<long long>(1 + i) * D + c);
        v = make_float4(a.x + p.x + te.x, a.y + p.y + te.y, a.z + p.z + te.z, a.w + p.w + te.w);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(object + ((static_cast<long long>(b) * F + f) * O + (i - N)) * D + c);
        v = make_float4(a.x + te.x, a.y + te.y, a.z + te.z, a.w + te.w);
      }
      if (type_embed != nullptr) {
        const float4 ty = *reinterpret_cast<const float4*>(type_embed + (i < N ? 0 : D) + c);
        v.x += ty.x; v.y += ty.y; v.z += ty.z; v.w += ty.w;
      }
    }
    *reinterpret_cast<float4*>(x + bt * D + c) = v;
  }
}

// Gradient of the assembly: scatters dx (fp32 [B,T,D]) into bf16 operand buffers for the patch / object embedding
// weight gradients and reduces the embedding-table gradients over the batch (one thread per (token, 4 dims),
// looping over b, then a handful of atomics).
__global__ void assemble_tokens_bwd_kernel(const float* __restrict__ dx, __nv_bfloat16* __restrict__ dpatch,
                                           __nv_bfloat16* __restrict__ dobject, float* __restrict__ dcls,
                                           float* __restrict__ dpos, float* __restrict__ dtemporal,
                                           float* __restrict__ dtype_embed, int B, int F, int N, int O, int D) {
  const int n = N + O, T = 1 + F * n, d4 = D >> 2;
  const long long total = static_cast<long long>(T) * d4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4) << 2;
    const int tok = static_cast<int>(idx / d4);
    const int f = tok == 0 ? 0 : (tok - 1) / n, i = tok == 0 ? 0 : (tok - 1) - f * n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 g = *reinterpret_cast<const float4*>(dx + (static_cast<long long>(b) * T + tok) * D + c);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      if (tok != 0) {
        uint2 pk;
        pk.x = pack_bf16x2(g.x, g.y);
        pk.y = pack_bf16x2(g.z, g.w);
        if (i < N) {
          if (dpatch != nullptr)
            *reinterpret_cast<uint2*>(dpatch + ((static_cast<long long>(b) * F + f) * N + i) * D + c) = pk;
        } else if (dobject != nullptr) {
          *reinterpret_cast<uint2*>(dobject + ((static_cast<long long>(b) * F + f) * O + (i - N)) * D + c) = pk;
        }
      }
    }
    auto add4 = [&](float* p) {
      atomicAdd(p + 0, acc.x); atomicAdd(p + 1, acc.y); atomicAdd(p + 2, acc.z); atomicAdd(p + 3, acc.w);
    };
    if (tok == 0) {
      if (dcls != nullptr) add4(dcls + c);
      if (dpos != nullptr) add4(dpos + c);
    } else {
      if (dtemporal != nullptr) add4(dtemporal + static_cast<long long>(f) * D + c);
      if (i < N && dpos != nullptr) add4(dpos + static_cast<long long>(1 + i) * D + c);
      if (dtype_embed != nullptr) add4(dtype_embed + (i < N ? 0 : D) + c);
    }
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[c] += sum_r x[r, c]   (bias gradients). Each CTA reduces one slab of rows; a thread owns 8 adjacent columns (one
// 16-byte load per row). The slab height is chosen so that the number of slabs is a multiple of the SM count (2 per SM
// at M = 59424): with fixed 256-row slabs 233 CTAs landed 2-1 on the SMs (79 % balance). 64-row slabs were tried too:
// 4x the atomics onto the same `cols` addresses made the kernel 1.7x slower in the step.
// The loads go through cp.async into a per-thread ring in shared memory (8 rows deep): ptxas otherwise keeps only ~3
// register loads in flight per thread (it schedules for 32 registers), which held the kernel at 60 % of the HBM roofline.
// A thread only ever reads the 16 bytes it copied itself, so no block-level synchronisation is needed.
// Narrow inputs (the dq slice of a dqkv buffer: 768 columns = 96 threads per row) would leave an SM with ~6 warps: the
// CTA then runs several row lanes (lane y takes rows y, y + lanes, ...; 4 x 96 threads), reduced through shared memory
// at the end, so the number of atomics per column stays what it was.
constexpr int kColsumRing = 8;
__global__ void __launch_bounds__(512) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                          long long rows, int cols, float* __restrict__ out,
                                                          int rows_per_cta, int tpr) {
  extern __shared__ __align__(16) uint8_t colsum_ring[];          // [kColsumRing][blockDim.x] x 16 B
  pdl_launch_dependents();
  pdl_wait();
  const int lanes = blockDim.x / tpr;                             // row lanes of this CTA
  const int tx = threadIdx.x % tpr, ly = threadIdx.x / tpr;
  const int c = (blockIdx.y * tpr + tx) * 8;
  const bool active = c < cols;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta + ly;
  const long long r1 = min(rows, static_cast<long long>(blockIdx.x + 1) * rows_per_cta);
  const uint32_t slot0 = static_cast<uint32_t>(__cvta_generic_to_shared(colsum_ring)) + threadIdx.x * 16;
  const uint32_t slot_stride = blockDim.x * 16;
  const long long step = static_cast<long long>(lanes) * ld;
  const __nv_bfloat16* src = x + r0 * ld + c;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (active) {
#pragma unroll
    for (int u = 0; u < kColsumRing; ++u) {
      if (r0 + static_cast<long long>(u) * lanes < r1)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(slot0 + u * slot_stride), "l"(src + u * step) : "memory");
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    int slot = 0;
    long long i = 0;
    for (long long r = r0; r < r1; r += lanes, ++i) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kColsumRing - 1) : "memory");
      const uint4 v = *reinterpret_cast<const uint4*>(colsum_ring + (slot * blockDim.x + threadIdx.x) * 16);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
      }
      if (r + static_cast<long long>(kColsumRing) * lanes < r1)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(slot0 + slot * slot_stride), "l"(src + (i + kColsumRing) * step) : "memory");
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      slot = slot + 1 == kColsumRing ? 0 : slot + 1;
    }
  }
  if (lanes > 1) {                                                // CTA-uniform
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();                                              // every ring slot is dead: the ring becomes [lanes][tpr][8] floats
    float* red = reinterpret_cast<float*>(colsum_ring);
    if (ly > 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red[(ly * tpr + tx) * 8 + k] = acc[k];
    }
    __syncthreads();
    if (ly > 0) return;
    for (int l = 1; l < lanes; ++l) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += red[(l * tpr + tx) * 8 + k];
    }
  }
  if (!active) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) atomicAdd(out + c + k, acc[k]);
}

// out[c] += sum_k v[k] * W[k][c] (fp32): a [1, K] x [K, N] product, one CTA per (128 columns, 32 k rows), partial sums by
// atomics. Used for the v part of the qkv bias gradient, db_v = db_proj . W_proj (see oat.h).
__global__ void __launch_bounds__(128) vecmat_f32_kernel(const float* __restrict__ v, const float* __restrict__ W, long long ldw,
                                                         int K, int N, float* __restrict__ out) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int k0 = blockIdx.y * 32;
  if (c >= N) return;
  float acc = 0.f;
#pragma unroll 8
  for (int k = k0; k < min(K, k0 + 32); ++k) acc = fmaf(__ldg(v + k), __ldg(W + static_cast<long long>(k) * ldw + c), acc);
  atomicAdd(out + c, acc);
}

// ------------------------------------------------------------------------------------------------ text embeddings
// out[b*L + l] = word_emb[ids[b*L+l]] + pos_emb[l]   (fp32; the LayerNorm kernel follows)
__global__ void text_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                  const float* __restrict__ pos, float* __restrict__ out, long long rows, int L, int D) {
  const int d4 = D >> 2;
  const long long total = rows * d4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4) << 2;
    const long long r = idx / d4;
    const float4 w = *reinterpret_cast<const float4*>(word + ids[r] * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + (r % L) * D + c);
    *reinterpret_cast<float4*>(out + r * D + c) = make_float4(w.x + p.x, w.y + p.y, w.z + p.z, w.w + p.w);
  }
}
__global__ void text_embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dsum,
                                      float* __restrict__ dword, float* __restrict__ dpos, long long rows, int L, int D) {
  const long long total = rows * D;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % D);
    const long long r = idx / D;
    const float g = dsum[idx];
    if (dword != nullptr) atomicAdd(dword + ids[r] * D + c, g);
    if (dpos != nullptr) atomicAdd(dpos + (r % L) * D + c, g);
  }
}

static unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace oat

using namespace oat;

extern "C" int oat_cast_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t cols,
                             int32_t cols_padded, int32_t relu, oat_stream_t stream) {
  OAT_REQUIRE(cols > 0 && cols_padded >= cols && cols_padded % 4 == 0 && ldd % 4 == 0 && ldd >= cols_padded,
              "oat_cast_bf16: cols=%d cols_padded=%d ldd=%lld (padded width and pitch must be multiples of 4)", cols,
              cols_padded, (long long)ldd);
  if (rows <= 0) return OAT_OK;
  const long long total = rows * (cols_padded / 4);
  cast_bf16_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, cols_padded, relu);
  return check_launch("cast_bf16_kernel");
}

extern "C" int oat_split3_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t cols,
                               int32_t relu, oat_stream_t stream) {
  OAT_REQUIRE(cols > 0 && cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ldd >= 3LL * cols,
              "oat_split3_bf16: cols=%d lds=%lld ldd=%lld (multiples of 4, ldd >= 3*cols)", cols, (long long)lds,
              (long long)ldd);
  OAT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
              "oat_split3_bf16: src must be 16-byte and dst 8-byte aligned");
  if (rows <= 0) return OAT_OK;
  const long long total = rows * (cols / 4);
  split3_bf16_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, relu);
  return check_launch("split3_bf16_kernel");
}

extern "C" int oat_dropout_fwd(const float* x, int64_t ldx, const float* residual, int64_t ldr, float* out, int64_t ldo,
                               void* out_bf16, int64_t ldob, void* out_split3, int64_t ld3, int64_t rows, int32_t cols,
                               float p, uint64_t seed, uint32_t site, oat_stream_t stream) {
  OAT_REQUIRE(x != nullptr && cols > 0 && cols % 4 == 0 && p >= 0.f && p < 1.f, "oat_dropout_fwd: bad arguments (cols %% 4, 0 <= p < 1)");
  OAT_REQUIRE(ldx % 4 == 0 && ldr % 4 == 0 && ldo % 4 == 0 && ldob % 4 == 0 && ld3 % 4 == 0, "oat_dropout_fwd: pitches must be multiples of 4");
  OAT_REQUIRE(out_split3 == nullptr || ld3 >= 3LL * cols, "oat_dropout_fwd: split output needs a pitch >= 3 * cols");
  if (rows <= 0) return OAT_OK;
  const long long total = rows * (cols / 4);
  dropout_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      x, ldx, residual, ldr, out, ldo, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldob,
      reinterpret_cast<__nv_bfloat16*>(out_split3), ld3, rows, cols, 1.0f / (1.0f - p),
      static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0), seed, site);
  return check_launch("dropout_fwd_kernel");
}

extern "C" int oat_dropout_bwd(const float* dy_f32, int64_t lddy, const void* dy_bf16, int64_t lddyb, float* dx_f32,
                               int64_t lddx, void* dx_bf16, int64_t lddxb, int64_t rows, int32_t cols, float p,
                               uint64_t seed, uint32_t site, oat_stream_t stream) {
  OAT_REQUIRE((dy_f32 != nullptr || dy_bf16 != nullptr) && (dx_f32 != nullptr || dx_bf16 != nullptr) && cols > 0 &&
              cols % 4 == 0 && p >= 0.f && p < 1.f, "oat_dropout_bwd: bad arguments");
  OAT_REQUIRE(lddy % 4 == 0 && lddyb % 4 == 0 && lddx % 4 == 0 && lddxb % 4 == 0, "oat_dropout_bwd: pitches must be multiples of 4");
  if (rows <= 0) return OAT_OK;
  const long long total = rows * (cols / 4);
  dropout_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      dy_f32, lddy, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), lddyb, dx_f32, lddx,
      reinterpret_cast<__nv_bfloat16*>(dx_bf16), lddxb, rows, cols, 1.0f / (1.0f - p),
      static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0), seed, site);
  return check_launch("dropout_bwd_kernel");
}

extern "C" int oat_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint32_t site, oat_stream_t stream) {
  OAT_REQUIRE(keep != nullptr && n > 0 && p >= 0.f && p < 1.f, "oat_dropout_mask: bad arguments");
  dropout_mask_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
      keep, n, static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0), seed, site);
  return check_launch("dropout_mask_kernel");
}

extern "C" int oat_unpack_wgrad(float* scratch, int64_t ld, int32_t cols, float* dw, int64_t ldw, float* db, int64_t rows,
                                oat_stream_t stream) {
  OAT_REQUIRE(scratch != nullptr && dw != nullptr && db != nullptr && cols > 0 && cols % 4 == 0 && ld >= cols + 4 &&
              ld % 4 == 0 && ldw % 4 == 0, "oat_unpack_wgrad: bad arguments (cols %% 4 == 0, ld >= cols + 4)");
  OAT_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0,
              "oat_unpack_wgrad: scratch and dw must be 16-byte aligned");
  if (rows <= 0) return OAT_OK;
  const long long total = rows * ((cols >> 2) + 1);
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(scratch, ld, cols, dw, ldw, db, rows);
  return check_launch("unpack_wgrad_kernel");
}

extern "C" int oat_cast_multi(const int64_t* table, const int64_t* chunk_prefix, int32_t n, int64_t total_chunks,
                              oat_stream_t stream) {
  OAT_REQUIRE(table != nullptr && chunk_prefix != nullptr && n > 0 && total_chunks > 0, "oat_cast_multi: bad arguments");
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const unsigned grid = static_cast<unsigned>(total_chunks < cap ? total_chunks : cap);
  cast_multi_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(table),
                                                         reinterpret_cast<const long long*>(chunk_prefix), n, total_chunks);
  return check_launch("cast_multi_kernel");
}

extern "C" int oat_relu_bwd(const float* x, int64_t ldx, const void* dy_bf16, int64_t lddy, float* dx, int64_t lddx,
                            int64_t rows, int32_t cols, oat_stream_t stream) {
  if (rows <= 0) return OAT_OK;
  relu_bwd_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(
      x, ldx, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), lddy, dx, lddx, rows, cols);
  return check_launch("relu_bwd_kernel");
}

extern "C" int oat_im2col_patches(const float* video, void* out_bf16, int64_t BF, int32_t C, int32_t H, int32_t W,
                                  int32_t P, oat_stream_t stream) {
  OAT_REQUIRE(P % 8 == 0 && H % P == 0 && W % P == 0, "oat_im2col_patches: H=%d W=%d must be multiples of P=%d (P %% 8 == 0)", H, W, P);
  if (BF <= 0) return OAT_OK;
  const long long total = BF * C * H * (W / 8);
  im2col_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(video, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                    BF, C, H, W, P);
  return check_launch("im2col_kernel");
}

extern "C" int oat_assemble_tokens(const float* patch, const float* object, const float* cls_token,
                                   const float* pos_embed, const float* temporal_embed, const float* type_embed,
                                   float* x, int32_t B, int32_t F, int32_t N, int32_t O, int32_t D,
                                   oat_stream_t stream) {
  OAT_REQUIRE(D % 4 == 0 && B > 0 && F > 0 && N > 0 && O >= 0, "oat_assemble_tokens: bad geometry");
  OAT_REQUIRE(O == 0 || object != nullptr, "oat_assemble_tokens: object tokens requested without object embeddings");
  const long long total = static_cast<long long>(B) * (1 + F * (N + O)) * (D / 4);
  assemble_tokens_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(patch, object, cls_token, pos_embed,
                                                                             temporal_embed, type_embed, x, B, F, N, O, D);
  return check_launch("assemble_tokens_kernel");
}

extern "C" int oat_assemble_tokens_bwd(const float* dx, void* dpatch_bf16, void* dobject_bf16, float* dcls,
                                       float* dpos, float* dtemporal, float* dtype_embed, int32_t B, int32_t F,
                                       int32_t N, int32_t O, int32_t D, oat_stream_t stream) {
  OAT_REQUIRE(D % 4 == 0 && B > 0 && F > 0 && N > 0 && O >= 0, "oat_assemble_tokens_bwd: bad geometry");
  const long long total = static_cast<long long>(1 + F * (N + O)) * (D / 4);
  assemble_tokens_bwd_kernel<<<grid_for(total, 128), 128, 0, as_stream(stream)>>>(
      dx, reinterpret_cast<__nv_bfloat16*>(dpatch_bf16), reinterpret_cast<__nv_bfloat16*>(dobject_bf16), dcls, dpos,
      dtemporal, dtype_embed, B, F, N, O, D);
  return check_launch("assemble_tokens_bwd_kernel");
}

extern "C" int oat_colsum_bf16(const void* x_bf16, int64_t ld, int64_t rows, int32_t cols, float* out,
                               oat_stream_t stream) {
  OAT_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "oat_colsum_bf16: cols and ld must be multiples of 8");
  if (rows <= 0) return OAT_OK;
  // one block spans all columns when it can (2304 -> 288 threads, 3072 -> 384): no nearly-empty second column block
  const int want = ((cols / 8 + 31) / 32) * 32;
  const int tpr = want <= 384 ? want : 256;          // threads per row; ring: 8 x threads x 16 B <= 48 KB
  int lanes = 384 / tpr;                             // row lanes: fill the CTA up to 384 threads
  if (lanes < 1) lanes = 1;
  if (lanes > 4) lanes = 4;
  const int threads = tpr * lanes;
  // slabs: a multiple of the SM count, about 200-256 rows each
  const long long sms = num_sms();
  long long k = (rows + 256 * sms - 1) / (256 * sms);
  if (k < 1) k = 1;
  long long rows_per = (rows + k * sms - 1) / (k * sms);
  if (rows_per < 16) rows_per = 16;
  dim3 grid(static_cast<unsigned>((rows + rows_per - 1) / rows_per), static_cast<unsigned>((cols / 8 + tpr - 1) / tpr));
  if (launch_pdl(colsum_bf16_kernel, grid, dim3(threads), static_cast<size_t>(kColsumRing) * threads * 16, as_stream(stream),
                 reinterpret_cast<const __nv_bfloat16*>(x_bf16), ld, rows, cols, out, static_cast<int>(rows_per), tpr) != cudaSuccess)
    return check_launch("colsum_bf16_kernel");
  return check_launch("colsum_bf16_kernel");
}

extern "C" int oat_vecmat_f32(const float* v, const float* W, int64_t ldw, int32_t K, int32_t N, float* out,
                              oat_stream_t stream) {
  OAT_REQUIRE(v != nullptr && W != nullptr && out != nullptr && ldw >= N, "oat_vecmat_f32: bad arguments");
  if (K <= 0 || N <= 0) return OAT_OK;
  vecmat_f32_kernel<<<dim3((N + 127) / 128, (K + 31) / 32), 128, 0, as_stream(stream)>>>(v, W, ldw, K, N, out);
  return check_launch("vecmat_f32_kernel");
}

extern "C" int oat_text_embed(const int64_t* ids, const float* word_emb, const float* pos_emb, float* out,
                              int64_t rows, int32_t L, int32_t D, oat_stream_t stream) {
  OAT_REQUIRE(D % 4 == 0 && L > 0, "oat_text_embed: bad geometry");
  if (rows <= 0) return OAT_OK;
  text_embed_kernel<<<grid_for(rows * (D / 4), 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(ids), word_emb, pos_emb, out, rows, L, D);
  return check_launch("text_embed_kernel");
}

extern "C" int oat_text_embed_bwd(const int64_t* ids, const float* dsum, float* dword, float* dpos, int64_t rows,
                                  int32_t L, int32_t D, oat_stream_t stream) {
  if (rows <= 0) return OAT_OK;
  text_embed_bwd_kernel<<<grid_for(rows * D, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(ids), dsum, dword, dpos, rows, L, D);
  return check_launch("text_embed_bwd_kernel");
}
