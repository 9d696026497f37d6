// Object -> patch attention and the bbox -> patch-mask bookkeeping that feeds it.
//
// One kernel, three score -> weight modes, optional weights . V:
//   mode 0 "mask"    : weights = given binary patch masks; out = masks @ v
//                      (einsum 'b o l, b l c -> b o c', OATrans/model/oa_model_global_local.py:178)
//   mode 1 "sigmoid" : weights = sigmoid(q . k)     (OATrans/model/oa_model_region_mem.py:147-151)
//   mode 2 "softmax" : weights = softmax(q . k * C^-0.5)
//                      (Visualization/Cross_Modality_Transformer_Visualization/visualize.py:155-168)
// q (B, O, C), k (B, L, C), v (B, L, Cv) fp32; weights (B, O, L) and out (B, O, Cv) fp32. 21.7 MFLOP per frame at
// O = 36, L = 196, C = 768: fp32 SIMT, one CTA per (batch, object); k / v rows are shared by the O CTAs of a sample
// through L2. Masks: patch_all_masks_from_bbox, OATrans/base/base_dataset_global_local.py:348-356 - float64
// arithmetic like numpy so that int() / ceil() land on the same integers (bit-exact).
#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int kXWarps = 8;

__global__ void __launch_bounds__(kXWarps * 32)
object_patch_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                         const float* __restrict__ masks, float* __restrict__ weights, float* __restrict__ out, int O,
                         int L, int C, int Cv, int mode) {
  extern __shared__ float sw[];          // [L] weights of this (b, o) + [kXWarps] scratch
  float* red = sw + L;
  const int b = blockIdx.x / O, o = blockIdx.x - b * O;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* kb = k + static_cast<long long>(b) * L * C;
  if (mode == 0) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) sw[l] = masks[(static_cast<long long>(b) * O + o) * L + l];
  } else {
    const float* qr = q + (static_cast<long long>(b) * O + o) * C;
    const float scl = mode == 2 ? rsqrtf(static_cast<float>(C)) : 1.0f;
    for (int l = warp; l < L; l += kXWarps) {
      float s = 0.f;
      for (int c = lane * 4; c < C; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(qr + c);
        const float4 x = *reinterpret_cast<const float4*>(kb + static_cast<long long>(l) * C + c);
        s += (a.x * x.x + a.y * x.y) + (a.z * x.z + a.w * x.w);
      }
      s = warp_sum(s) * scl;
      if (lane == 0) sw[l] = mode == 1 ? 1.0f / (1.0f + expf(-s)) : s;
    }
  }
  __syncthreads();
  if (mode == 2) {
    float mx = -INFINITY;
    for (int l = threadIdx.x; l < L; l += blockDim.x) mx = fmaxf(mx, sw[l]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < kXWarps; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float e = expf(sw[l] - mx);
      sw[l] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < kXWarps; ++w) sum += red[w];
    const float inv = 1.0f / sum;
    for (int l = threadIdx.x; l < L; l += blockDim.x) sw[l] *= inv;
    __syncthreads();
  }
  if (weights != nullptr)
    for (int l = threadIdx.x; l < L; l += blockDim.x) weights[(static_cast<long long>(b) * O + o) * L + l] = sw[l];
  if (out != nullptr && v != nullptr) {
    const float* vb = v + static_cast<long long>(b) * L * Cv;
    for (int c = threadIdx.x; c < Cv; c += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc = fmaf(sw[l], vb[static_cast<long long>(l) * Cv + c], acc);
      out[(static_cast<long long>(b) * O + o) * Cv + c] = acc;
    }
  }
}

// boxes fp64 [n, stride] (x1, y1, x2, y2 in [0, 1] first) -> masks fp32 [n, g*g], row-major patch grid
__global__ void patch_masks_kernel(const double* __restrict__ boxes, int stride, float* __restrict__ masks, int n, int g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * g * g) return;
  const int box = idx / (g * g), cell = idx - box * g * g, r = cell / g, c = cell - r * g;
  const double* bx = boxes + static_cast<long long>(box) * stride;
  const double x1 = bx[0] * g, y1 = bx[1] * g, x2 = bx[2] * g, y2 = bx[3] * g;
  // Python: mask[int(y1):ceil(y2), int(x1):ceil(x2)] = 1 ; int() truncates toward zero, slices clamp to [0, g]
  const long long r0 = static_cast<long long>(y1), r1 = static_cast<long long>(ceil(y2));
  const long long c0 = static_cast<long long>(x1), c1 = static_cast<long long>(ceil(x2));
  auto norm = [g](long long i) { return i < 0 ? (i + g < 0 ? 0LL : i + g) : (i > g ? static_cast<long long>(g) : i); };
  const long long rs = norm(r0), re = norm(r1), cs = norm(c0), ce = norm(c1);
  masks[idx] = (r >= rs && r < re && c >= cs && c < ce) ? 1.0f : 0.0f;
}

// base/base_dataset_region_mem.py:233-247: for each of the `para` selected boxes, the union of the masks of EVERY box
// of the same object class; boxes are scaled by the grid size first (fp64, like numpy) and cut with int() / ceil().
__global__ void patch_masks_same_class_kernel(const double* __restrict__ boxes, int stride, const int* __restrict__ classes,
                                              const int* __restrict__ sel, float* __restrict__ masks, int n, int para,
                                              int g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= para * g * g) return;
  const int j = idx / (g * g), cell = idx - j * g * g, r = cell / g, c = cell - r * g;
  const int cls = classes[sel[j]];
  auto norm = [g](long long i) { return i < 0 ? (i + g < 0 ? 0LL : i + g) : (i > g ? static_cast<long long>(g) : i); };
  float v = 0.0f;
  for (int i = 0; i < n; ++i) {
    if (classes[i] != cls) continue;
    const double* bx = boxes + static_cast<long long>(i) * stride;
    const double x1 = bx[0] * g, y1 = bx[1] * g, x2 = bx[2] * g, y2 = bx[3] * g;
    const long long rs = norm(static_cast<long long>(y1)), re = norm(static_cast<long long>(ceil(y2)));
    const long long cs = norm(static_cast<long long>(x1)), ce = norm(static_cast<long long>(ceil(x2)));
    if (r >= rs && r < re && c >= cs && c < ce) v = 1.0f;
  }
  masks[idx] = v;
}

// base/base_dataset_global_local.py:395-405: ends[i] = running sum of int(token_len[indices[i]]); total = the sum.
__global__ void object_tags_masks_kernel(const double* __restrict__ lens, const long long* __restrict__ indices,
                                         float* __restrict__ ends, int* __restrict__ total, int k) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long end = 0;
  for (int i = 0; i < k; ++i) {
    end += static_cast<long long>(lens[indices[i]]);      // int() truncates toward zero
    ends[i] = static_cast<float>(end);
  }
  total[0] = static_cast<int>(end);
}

// base/base_dataset.py:593-650 (read_object_from_disk after np.load): confidence order, optional class de-duplication
// (v = 2), numpy 'edge' padding to top_k - which pads BOTH axes, so with m < top_k regions the feature part is
// 2048 + (top_k - m) wide (reproduced as the reference computes it) - and box geometry scaled by the image size.
// One CTA per output row; n is small (<= 100 regions), so the order is found by counting.
__global__ void __launch_bounds__(256)
region_features_topk_kernel(const float* __restrict__ x, const float* __restrict__ bbox, const float* __restrict__ conf,
                            const long long* __restrict__ ids, int n, int fdim, int top_k, int v, float image_w,
                            float image_h, float* __restrict__ out, long long ld_out, int* __restrict__ m_out) {
  __shared__ int order[128];        // order[r]: index of the r-th most confident region (np.argsort(conf)[::-1])
  __shared__ int uniq[128];         // v = 2: np.unique(ids, return_index=True)[1] - first index of each class, by class id
  __shared__ int m_sh;
  const int r = blockIdx.x;
  if (threadIdx.x == 0) m_sh = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int pos = 0;
    for (int j = 0; j < n; ++j) pos += (conf[j] > conf[i] || (conf[j] == conf[i] && j > i)) ? 1 : 0;
    order[pos] = i;
    if (v == 2) {
      bool first = true;
      for (int j = 0; j < i; ++j) first = first && ids[j] != ids[i];
      if (first) {
        int rank = 0;
        for (int j = 0; j < n; ++j) {
          bool jf = ids[j] < ids[i];
          for (int q = 0; jf && q < j; ++q) jf = ids[q] != ids[j];
          rank += jf ? 1 : 0;
        }
        uniq[rank] = i;
        atomicAdd(&m_sh, 1);
      }
    }
  }
  __syncthreads();
  const int m = v == 2 ? m_sh : n;
  const int res = top_k > m ? top_k - m : 0;
  if (r == 0 && threadIdx.x == 0) m_out[0] = m;
  const int rr = r < m ? r : m - 1;                        // 'edge' rows
  const int src = v == 2 ? order[uniq[rr]] : order[rr];
  const int fw = fdim + res;                               // 'edge' columns of the feature block
  for (int c = threadIdx.x; c < fw; c += blockDim.x) out[r * ld_out + c] = x[static_cast<long long>(src) * fdim + (c < fdim ? c : fdim - 1)];
  if (threadIdx.x == 0) {
    const float* b = bbox + static_cast<long long>(src) * 4;
    const float bw = b[2] - b[0], bh = b[3] - b[1];
    const float sw = __fdiv_rn(bw, image_w), sh = __fdiv_rn(bh, image_h);
    const float sx = __fdiv_rn(b[0], image_w), sy = __fdiv_rn(b[1], image_h);
    float* o = out + r * ld_out + fw;
    o[0] = sx; o[1] = sy; o[2] = __fadd_rn(sx, sw); o[3] = __fadd_rn(sy, sh); o[4] = sw; o[5] = sh;
  }
}

}  // namespace oat

using namespace oat;

extern "C" int oat_patch_masks_same_class(const double* boxes, int32_t stride, const int32_t* classes,
                                          const int32_t* sel, float* masks, int32_t n, int32_t para, int32_t grid,
                                          oat_stream_t stream) {
  OAT_REQUIRE(n > 0 && para > 0 && grid > 0 && stride >= 4 && boxes != nullptr && classes != nullptr && sel != nullptr,
              "oat_patch_masks_same_class: bad arguments");
  const int total = para * grid * grid;
  patch_masks_same_class_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(boxes, stride, classes, sel, masks, n,
                                                                                   para, grid);
  return check_launch("patch_masks_same_class_kernel");
}

extern "C" int oat_object_tags_masks(const double* token_lens, const int64_t* indices, float* ends, int32_t* total,
                                     int32_t k, oat_stream_t stream) {
  OAT_REQUIRE(k > 0 && token_lens != nullptr && indices != nullptr && ends != nullptr && total != nullptr,
              "oat_object_tags_masks: bad arguments");
  object_tags_masks_kernel<<<1, 32, 0, as_stream(stream)>>>(token_lens, reinterpret_cast<const long long*>(indices), ends,
                                                          total, k);
  return check_launch("object_tags_masks_kernel");
}

extern "C" int oat_region_features_topk(const float* x, const float* bbox, const float* conf, const int64_t* ids,
                                        int32_t n, int32_t feat_dim, int32_t top_k, int32_t v, int32_t image_w,
                                        int32_t image_h, float* out, int64_t ld_out, int32_t* m_out,
                                        oat_stream_t stream) {
  OAT_REQUIRE(n > 0 && n <= 128 && top_k > 0 && feat_dim > 0 && (v == 1 || v == 2), "oat_region_features_topk: bad arguments (n <= 128)");
  OAT_REQUIRE(ld_out >= feat_dim + top_k + 6, "oat_region_features_topk: ld_out must hold feat_dim + top_k + 6 columns");
  OAT_REQUIRE(v == 1 || ids != nullptr, "oat_region_features_topk: v = 2 needs the object class ids");
  region_features_topk_kernel<<<top_k, 256, 0, as_stream(stream)>>>(x, bbox, conf, reinterpret_cast<const long long*>(ids),
                                                                  n, feat_dim, top_k, v, static_cast<float>(image_w),
                                                                  static_cast<float>(image_h), out, ld_out, m_out);
  return check_launch("region_features_topk_kernel");
}

extern "C" int oat_object_patch_attn(const float* q, const float* k, const float* v, const float* masks,
                                     float* weights, float* out, int32_t B, int32_t O, int32_t L, int32_t C,
                                     int32_t Cv, int32_t mode, oat_stream_t stream) {
  OAT_REQUIRE(B > 0 && O > 0 && L > 0, "oat_object_patch_attn: empty problem");
  OAT_REQUIRE(mode >= 0 && mode <= 2, "oat_object_patch_attn: mode must be 0 (mask), 1 (sigmoid) or 2 (softmax)");
  OAT_REQUIRE(mode != 0 || masks != nullptr, "oat_object_patch_attn: mode 0 needs masks");
  OAT_REQUIRE(mode == 0 || (q != nullptr && k != nullptr && C > 0 && C % 4 == 0), "oat_object_patch_attn: q/k needed, C %% 4 == 0");
  OAT_REQUIRE(weights != nullptr || (out != nullptr && v != nullptr), "oat_object_patch_attn: nothing to compute");
  const size_t smem = (static_cast<size_t>(L) + kXWarps) * sizeof(float);
  OAT_REQUIRE(smem <= 48 * 1024, "oat_object_patch_attn: L=%d too large", L);
  object_patch_attn_kernel<<<B * O, kXWarps * 32, smem, as_stream(stream)>>>(q, k, v, masks, weights, out, O, L, C, Cv, mode);
  return check_launch("object_patch_attn_kernel");
}

extern "C" int oat_patch_masks_from_bbox(const double* boxes, int32_t stride, float* masks, int32_t n, int32_t grid,
                                         oat_stream_t stream) {
  OAT_REQUIRE(n >= 0 && grid > 0 && stride >= 4, "oat_patch_masks_from_bbox: bad arguments");
  if (n == 0) return OAT_OK;
  const int total = n * grid * grid;
  patch_masks_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(boxes, stride, masks, n, grid);
  return check_launch("patch_masks_kernel");
}
