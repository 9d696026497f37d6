// Variant heads on top of the towers and the backward of the object -> patch attention (X4):
//
//   * oat_object_patch_attn_bwd : gradient of oat_object_patch_attn (csrc/xattn.cu) w.r.t. q, k, v for the three
//     score -> weight modes (mask pooling OATrans/model/oa_model_global_local.py:178, sigmoid region similarity
//     OATrans/model/oa_model_region_mem.py:147-151, softmax).
//   * oat_token_pool(_bwd)      : out = a * cls + b * mean_l tok[:, l]  - `(video_embeddings + mean(region, 1)) / 2`
//     (oa_model_region_mem.py:119) and `torch.mean(region_feat, dim=1)` (trainer/trainer_global_local.py:207).
//   * oat_bce_sum               : scale * BCELoss(reduction='sum')(p, t) and its gradient - the region loss
//     `0.1 * criterion(region_sim, patch_mask) / rows` (trainer/trainer_region_mem.py:97,161-167), with torch's
//     clamps (log terms at -100, the gradient's denominator at 1e-12).
//
// All fp32 SIMT: these heads are a few MFLOP per sample (36 x 196 x 256 MACs), latency-bound; the tensors are read
// once with coalesced accesses. Deterministic (no floating-point atomics) so that multi-rank runs agree bit for bit.
#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int kHWarps = 8;

// One CTA per (b, o): dW = dweights + dout . v^T ; dS by mode ; dq = dS . k.   Writes dS to scratch for kernel B.
__global__ void __launch_bounds__(kHWarps * 32)
xattn_bwd_rows_kernel(const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ weights,
                      const float* __restrict__ dweights, const float* __restrict__ dout, float* __restrict__ dq,
                      float* __restrict__ ds, int O, int L, int C, int Cv, int mode) {
  extern __shared__ float sh[];          // [L] dW -> dS, [kHWarps] scratch
  float* red = sh + L;
  const int b = blockIdx.x / O, o = blockIdx.x - b * O;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = static_cast<long long>(b) * O + o;
  const float* w = weights + row * L;
  const float* vb = v != nullptr ? v + static_cast<long long>(b) * L * Cv : nullptr;
  const float* dor = dout != nullptr ? dout + row * Cv : nullptr;
  for (int l = warp; l < L; l += kHWarps) {
    float s = 0.f;
    if (dor != nullptr && vb != nullptr)
      for (int c = lane; c < Cv; c += 32) s = fmaf(dor[c], vb[static_cast<long long>(l) * Cv + c], s);
    s = warp_sum(s);
    if (lane == 0) sh[l] = s + (dweights != nullptr ? dweights[row * L + l] : 0.f);
  }
  __syncthreads();
  if (mode == 0) return;                 // the masks are data: nothing flows to q / k
  if (mode == 1) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) { const float p = w[l]; sh[l] = sh[l] * p * (1.0f - p); }
  } else {
    float dot = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) dot = fmaf(w[l], sh[l], dot);
    dot = warp_sum(dot);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    dot = 0.f;
    for (int i = 0; i < kHWarps; ++i) dot += red[i];
    const float scl = rsqrtf(static_cast<float>(C));
    for (int l = threadIdx.x; l < L; l += blockDim.x) sh[l] = w[l] * (sh[l] - dot) * scl;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) ds[row * L + l] = sh[l];
  if (dq != nullptr) {
    const float* kb = k + static_cast<long long>(b) * L * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc = fmaf(sh[l], kb[static_cast<long long>(l) * C + c], acc);
      dq[row * C + c] = acc;
    }
  }
}

// One CTA per (b, l): dk[b, l] = sum_o dS[b, o, l] q[b, o] ; dv[b, l] = sum_o w[b, o, l] dout[b, o].
__global__ void __launch_bounds__(128)
xattn_bwd_cols_kernel(const float* __restrict__ q, const float* __restrict__ weights, const float* __restrict__ ds,
                      const float* __restrict__ dout, float* __restrict__ dk, float* __restrict__ dv, int O, int L,
                      int C, int Cv, int mode) {
  const int b = blockIdx.x / L, l = blockIdx.x - b * L;
  const long long base = static_cast<long long>(b) * O;
  if (dk != nullptr && mode != 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float acc = 0.f;
      for (int o = 0; o < O; ++o) acc = fmaf(ds[(base + o) * L + l], q[(base + o) * C + c], acc);
      dk[(static_cast<long long>(b) * L + l) * C + c] = acc;
    }
  }
  if (dv != nullptr && dout != nullptr) {
    for (int c = threadIdx.x; c < Cv; c += blockDim.x) {
      float acc = 0.f;
      for (int o = 0; o < O; ++o) acc = fmaf(weights[(base + o) * L + l], dout[(base + o) * Cv + c], acc);
      dv[(static_cast<long long>(b) * L + l) * Cv + c] = acc;
    }
  }
}

// out[b, c] = a * cls[b, c] + bcoef * mean_l tok[b, l, c];  tok rows: tok + b * ld_batch + l * ld_tok.
__global__ void token_pool_kernel(const float* __restrict__ cls, long long ld_cls, const float* __restrict__ tok,
                                  long long ld_batch, long long ld_tok, float* __restrict__ out, int B, int L, int P,
                                  float a, float bcoef) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * P) return;
  const int b = idx / P, c = idx - b * P;
  const float* t = tok + b * ld_batch + c;
  float acc = 0.f;
  for (int l = 0; l < L; ++l) acc += t[l * ld_tok];
  float r = bcoef * (acc / static_cast<float>(L));
  if (cls != nullptr) r = fmaf(a, cls[b * ld_cls + c], r);
  out[static_cast<long long>(b) * P + c] = r;
}

// dcls = a * dout ; dtok[b, l, c] = bcoef / L * dout[b, c]
__global__ void token_pool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dcls, float* __restrict__ dtok,
                                      long long ld_batch, long long ld_tok, int B, int L, int P, float a, float bcoef) {
  const long long total = static_cast<long long>(B) * (L + 1) * P;
  const float bl = bcoef / static_cast<float>(L);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % P);
    const long long r = idx / P;
    const int l = static_cast<int>(r % (L + 1));
    const int b = static_cast<int>(r / (L + 1));
    const float g = dout[static_cast<long long>(b) * P + c];
    if (l == L) {
      if (dcls != nullptr) dcls[static_cast<long long>(b) * P + c] = a * g;
    } else {
      dtok[b * ld_batch + l * ld_tok + c] = bl * g;
    }
  }
}

// loss = scale * sum_i -(t log p + (1 - t) log(1 - p)), log clamped at -100; dp = scale * (p - t) / max(p (1 - p), 1e-12)
// (torch.nn.functional.binary_cross_entropy). One CTA: a fixed summation tree, identical on every rank.
__global__ void __launch_bounds__(1024)
bce_sum_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n, float scale,
               float* __restrict__ loss, float* __restrict__ dp) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = p[i], y = t[i];
    const float lp = fmaxf(logf(x), -100.0f), lq = fmaxf(logf(1.0f - x), -100.0f);
    acc -= y * lp + (1.0f - y) * lq;
    if (dp != nullptr) dp[i] = scale * (x - y) / fmaxf((1.0f - x) * x, 1e-12f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = scale * v;
  }
}

}  // namespace oat

using namespace oat;

extern "C" int oat_object_patch_attn_bwd(const float* q, const float* k, const float* v, const float* weights,
                                         const float* dweights, const float* dout, float* dq, float* dk, float* dv,
                                         float* ds_scratch, int32_t B, int32_t O, int32_t L, int32_t C, int32_t Cv,
                                         int32_t mode, oat_stream_t stream) {
  OAT_REQUIRE(B > 0 && O > 0 && L > 0, "oat_object_patch_attn_bwd: empty problem");
  OAT_REQUIRE(mode >= 0 && mode <= 2, "oat_object_patch_attn_bwd: mode must be 0 (mask), 1 (sigmoid) or 2 (softmax)");
  OAT_REQUIRE(weights != nullptr, "oat_object_patch_attn_bwd: the forward weights (or masks) are required");
  OAT_REQUIRE(dweights != nullptr || dout != nullptr, "oat_object_patch_attn_bwd: no incoming gradient");
  OAT_REQUIRE(dout == nullptr || (v != nullptr && Cv > 0), "oat_object_patch_attn_bwd: dout needs v");
  OAT_REQUIRE(mode == 0 || (q != nullptr && k != nullptr && ds_scratch != nullptr && C > 0),
              "oat_object_patch_attn_bwd: modes 1/2 need q, k and the (B, O, L) scratch");
  const size_t smem = (static_cast<size_t>(L) + kHWarps) * sizeof(float);
  OAT_REQUIRE(smem <= 48 * 1024, "oat_object_patch_attn_bwd: L=%d too large", L);
  cudaStream_t s = as_stream(stream);
  if (mode != 0) {
    xattn_bwd_rows_kernel<<<B * O, kHWarps * 32, smem, s>>>(k, v, weights, dweights, dout, dq, ds_scratch, O, L, C, Cv,
                                                           mode);
    int rc = check_launch("xattn_bwd_rows_kernel");
    if (rc != OAT_OK) return rc;
  }
  if ((dk != nullptr && mode != 0) || (dv != nullptr && dout != nullptr)) {
    xattn_bwd_cols_kernel<<<B * L, 128, 0, s>>>(q, weights, ds_scratch, dout, dk, dv, O, L, C, Cv, mode);
    return check_launch("xattn_bwd_cols_kernel");
  }
  return OAT_OK;
}

extern "C" int oat_token_pool(const float* cls, int64_t ld_cls, const float* tok, int64_t ld_batch, int64_t ld_tok,
                              float* out, int32_t B, int32_t L, int32_t P, float a, float b, oat_stream_t stream) {
  OAT_REQUIRE(B > 0 && L > 0 && P > 0 && tok != nullptr && out != nullptr, "oat_token_pool: bad arguments");
  const int total = B * P;
  token_pool_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(cls, ld_cls, tok, ld_batch, ld_tok, out, B, L, P,
                                                                       a, b);
  return check_launch("token_pool_kernel");
}

extern "C" int oat_token_pool_bwd(const float* dout, float* dcls, float* dtok, int64_t ld_batch, int64_t ld_tok,
                                  int32_t B, int32_t L, int32_t P, float a, float b, oat_stream_t stream) {
  OAT_REQUIRE(B > 0 && L > 0 && P > 0 && dout != nullptr && dtok != nullptr, "oat_token_pool_bwd: bad arguments");
  const long long total = static_cast<long long>(B) * (L + 1) * P;
  const long long want = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  token_pool_bwd_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(
      dout, dcls, dtok, ld_batch, ld_tok, B, L, P, a, b);
  return check_launch("token_pool_bwd_kernel");
}

extern "C" int oat_bce_sum(const float* p, const float* target, int64_t n, float scale, float* loss, float* dp,
                           oat_stream_t stream) {
  OAT_REQUIRE(n > 0 && p != nullptr && target != nullptr && loss != nullptr, "oat_bce_sum: bad arguments");
  bce_sum_kernel<<<1, 1024, 0, as_stream(stream)>>>(p, target, n, scale, loss, dp);
  return check_launch("bce_sum_kernel");
}
