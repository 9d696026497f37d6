// Batch-wide similarity + symmetric InfoNCE (forward and backward).
//
// Reference: sim_matrix, OATrans/model/model.py:164-172 (rows L2-normalised with the norm clamped at eps, then
// a_n @ b_n^T; rows = text, columns = video, trainer_dist.py:161) and NormSoftmaxLoss.forward,
// OATrans/model/loss.py:13-25 (log_softmax(x/T) over rows and over columns, mean of the two diagonals, negated).
//
// Pipeline (all on the caller's stream; the (Bg x Bg) matrix is tiny next to the towers, so the goal is accuracy
// and few launches, not tensor throughput):
//   1. normalise: one warp per row -> fp32 unit rows + a bf16 "hi | hi | lo" / "hi | lo | hi" split packing so
//      that the tcgen05 bf16 GEMM (K = 3*P) reproduces the fp32 dot product to ~2^-16 relative;
//   2. sims = A' . B'^T on the tcgen05 GEMM (oat_gemm_bf16);
//   3. row and column log-sum-exp of sims/T (one CTA per row / column);
//   4. loss scalar + dL/dsims = (softmax_row + softmax_col - 2 I) / (Bg T);
//   5. dL/d(unit rows) by fp32 SIMT products, then through the normalisation to dL/d(embeddings).
// The gradient rows of the LOCAL samples are sliced by the caller: the all-gather backward is a slice with no
// reduction (trainer_dist.py:40-45).
#include <string.h>

#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

// a [rows, P] fp32 -> an [rows, P] fp32 (unit rows), inv_norm [rows], pack [rows, 3P] bf16
// order = 0: hi | hi | lo     order = 1: hi | lo | hi
__global__ void loss_normalize_kernel(const float* __restrict__ a, float* __restrict__ an, float* __restrict__ inv_norm,
                                      __nv_bfloat16* __restrict__ pack, int rows, int P, float eps, int order) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = a + static_cast<long long>(row) * P;
  float ss = 0.f;
  for (int c = lane; c < P; c += 32) ss += r[c] * r[c];
  const float nrm = sqrtf(warp_sum(ss));
  const float inv = 1.0f / fmaxf(nrm, eps);
  if (lane == 0) inv_norm[row] = inv;
  for (int c = lane; c < P; c += 32) {
    const float v = r[c] * inv;
    an[static_cast<long long>(row) * P + c] = v;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = pack + static_cast<long long>(row) * 3 * P;
    o[c] = hi;
    o[P + c] = order == 0 ? hi : lo;
    o[2 * P + c] = order == 0 ? lo : hi;
  }
}

__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;
}

// blockIdx < n: row LSE of sims[i, :]/T ; blockIdx >= n: column LSE of sims[:, j]/T
__global__ void loss_lse_kernel(const float* __restrict__ sims, int n, int ld, float inv_t,
                                float* __restrict__ row_lse, float* __restrict__ col_lse) {
  __shared__ float sh[32];
  const bool is_col = blockIdx.x >= n;
  const int k = is_col ? blockIdx.x - n : blockIdx.x;
  const long long stride = is_col ? ld : 1;
  const float* p = sims + (is_col ? k : static_cast<long long>(k) * ld);
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, p[i * stride] * inv_t);
  mx = block_max(mx, sh);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += expf(p[i * stride] * inv_t - mx);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) (is_col ? col_lse : row_lse)[k] = mx + logf(s);
}

// dsims[i,j] = (exp(s/T - row_lse_i) + exp(s/T - col_lse_j) - 2 [i==j]) / (n T); loss = -(1/n) sum_i (2 s_ii/T - row_lse_i - col_lse_i)
__global__ void loss_grad_logits_kernel(const float* __restrict__ sims, int n, int ld, float inv_t,
                                        float* __restrict__ row_lse, const float* __restrict__ col_lse,
                                        float* __restrict__ dsims) {
  const int i = blockIdx.x;
  const float rl = row_lse[i];
  const float g = inv_t / n;
  __syncthreads();                       // every thread holds rl before the slot is recycled below
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float z = sims[static_cast<long long>(i) * ld + j] * inv_t;
    float d = expf(z - rl) + expf(z - col_lse[j]);
    if (i == j) {
      d -= 2.f;
      // this row's term of the loss takes over the row_lse slot (only block i reads row_lse[i]); the fixed-order sum
      // below makes the loss bit-identical on every rank of a data-parallel run (no floating-point atomics)
      row_lse[i] = -(2.f * z - rl - col_lse[j]) / n;
    }
    if (dsims != nullptr) dsims[static_cast<long long>(i) * ld + j] = d * g;
  }
}

__global__ void __launch_bounds__(256) loss_sum_kernel(const float* __restrict__ terms, int n, float* __restrict__ loss) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += terms[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) loss[0] = acc;
}

// side 0: dan[i,:] = sum_j dsims[i,j] bn[j,:] ; side 1: dbn[j,:] = sum_i dsims[i,j] an[i,:]
// then through x_n = x * inv (inv = 1/max(|x|, eps)): dx = inv * (dxn - x_n (x_n . dxn)) if the clamp is inactive,
// else dx = inv * dxn.
__global__ void loss_grad_embed_kernel(const float* __restrict__ dsims, int n_other, int ld, int P, const float* __restrict__ other_n,
                                       const float* __restrict__ self_n, const float* __restrict__ inv_norm, float eps,
                                       int side, float* __restrict__ dself) {
  __shared__ float sh[32];
  const int r = blockIdx.x;
  const float inv = inv_norm[r];
  const bool clamped = inv >= (1.0f / eps);
  for (int c0 = 0; c0 < P; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    float acc = 0.f;
    if (c < P) {
      for (int k = 0; k < n_other; ++k) {
        const float w = side == 0 ? dsims[static_cast<long long>(r) * ld + k] : dsims[static_cast<long long>(k) * ld + r];
        acc += w * other_n[static_cast<long long>(k) * P + c];
      }
    }
    // the projection needs the full-row dot product: only valid when P <= blockDim.x (checked on the host)
    const float xn = c < P ? self_n[static_cast<long long>(r) * P + c] : 0.f;
    const float dot = block_sum(acc * xn, sh);
    if (c < P) dself[static_cast<long long>(r) * P + c] = clamped ? inv * acc : inv * (acc - xn * dot);
  }
}

}  // namespace oat

using namespace oat;

namespace {
struct SimWs {
  float *an, *bn, *inv_a, *inv_b, *sims_pad;
  __nv_bfloat16 *pa, *pb;
  size_t mp;
};
size_t pad4(size_t v) { return (v + 3) & ~static_cast<size_t>(3); }
size_t sim_ws_bytes(size_t n, size_t m, size_t P) {
  const size_t mp = pad4(m);
  size_t f = n * P + mp * P + pad4(n) + mp + n * mp;
  size_t bytes = f * 4 + (pad4(n) + mp) * 3 * P * 2;
  return (bytes + 255) & ~static_cast<size_t>(255);
}
SimWs carve(void* ws, size_t n, size_t m, size_t P) {
  SimWs w;
  w.mp = pad4(m);
  w.an = reinterpret_cast<float*>(ws);
  w.bn = w.an + n * P;
  w.inv_a = w.bn + w.mp * P;
  w.inv_b = w.inv_a + pad4(n);
  w.sims_pad = w.inv_b + w.mp;
  w.pa = reinterpret_cast<__nv_bfloat16*>(w.sims_pad + n * w.mp);
  w.pb = w.pa + pad4(n) * 3 * P;
  return w;
}
}  // namespace

extern "C" size_t oat_sim_workspace_bytes(int32_t n, int32_t m, int32_t P) { return sim_ws_bytes(n, m, P); }

// sims[n, m] = unit(text) . unit(video)^T ; the workspace keeps the unit rows and 1/norms for the backward.
extern "C" int oat_sim_matrix_fwd(const float* text, const float* video, int32_t n, int32_t m, int32_t P, float eps,
                                  float* sims, void* workspace, size_t workspace_bytes, oat_stream_t stream) {
  OAT_REQUIRE(n > 0 && m > 0 && P > 0 && P % 8 == 0 && P <= 1024, "oat_sim_matrix_fwd: n=%d m=%d P=%d (P multiple of 8, <= 1024)", n, m, P);
  OAT_REQUIRE(workspace != nullptr && workspace_bytes >= sim_ws_bytes(n, m, P), "oat_sim_matrix_fwd: workspace too small");
  OAT_REQUIRE(sims != nullptr, "oat_sim_matrix_fwd: null output");
  cudaStream_t s = as_stream(stream);
  SimWs w = carve(workspace, n, m, P);
  if (w.mp != static_cast<size_t>(m)) {  // zero the padding rows of the packed video operand (padded sims columns)
    cudaError_t e0 = cudaMemsetAsync(w.pb + static_cast<size_t>(m) * 3 * P, 0, (w.mp - m) * 3 * P * 2, s);
    if (e0 != cudaSuccess) return set_error(OAT_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e0));
  }
  const int wpb = 8;
  loss_normalize_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, s>>>(text, w.an, w.inv_a, w.pa, n, P, eps, 0);
  loss_normalize_kernel<<<(m + wpb - 1) / wpb, wpb * 32, 0, s>>>(video, w.bn, w.inv_b, w.pb, m, P, eps, 1);
  int rc = check_launch("loss_normalize_kernel");
  if (rc != OAT_OK) return rc;
  oat_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = w.pa; g.lda = 3 * P; g.a_major = 0;
  g.B = w.pb; g.ldb = 3 * P; g.b_major = 0;
  g.M = n; g.N = static_cast<int32_t>(w.mp); g.K = 3 * P;
  g.alpha = 1.0f; g.scale = 1.0f;
  g.out_f32 = w.sims_pad; g.ld_f32 = static_cast<int64_t>(w.mp);
  rc = oat_gemm_bf16(&g, stream);
  if (rc != OAT_OK) return rc;
  cudaError_t e = cudaMemcpy2DAsync(sims, sizeof(float) * m, w.sims_pad, sizeof(float) * w.mp, sizeof(float) * m, n,
                                    cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
  return OAT_OK;
}

// dtext[n,P], dvideo[m,P] from dsims[n,m] (pitch m) through the product and the normalisation.
extern "C" int oat_sim_matrix_bwd(const float* dsims, int32_t n, int32_t m, int32_t P, float eps, float* dtext,
                                  float* dvideo, void* workspace, size_t workspace_bytes, oat_stream_t stream) {
  OAT_REQUIRE(n > 0 && m > 0 && P > 0 && P <= 1024, "oat_sim_matrix_bwd: bad sizes");
  OAT_REQUIRE(workspace != nullptr && workspace_bytes >= sim_ws_bytes(n, m, P), "oat_sim_matrix_bwd: workspace too small");
  cudaStream_t s = as_stream(stream);
  SimWs w = carve(workspace, n, m, P);
  const int threads = ((P + 31) / 32) * 32;
  if (dtext != nullptr) loss_grad_embed_kernel<<<n, threads, 0, s>>>(dsims, m, m, P, w.bn, w.an, w.inv_a, eps, 0, dtext);
  if (dvideo != nullptr) loss_grad_embed_kernel<<<m, threads, 0, s>>>(dsims, n, m, P, w.an, w.bn, w.inv_b, eps, 1, dvideo);
  return check_launch("loss_grad_embed_kernel");
}

// loss = NormSoftmaxLoss(sims) and dL/dsims (same pitch as sims). scratch: fp32 [2n].
extern "C" int oat_norm_softmax_loss(const float* sims, int32_t n, int64_t ld, float temperature, float* loss,
                                     float* dsims, float* scratch, oat_stream_t stream) {
  OAT_REQUIRE(n > 0 && ld >= n && temperature > 0.f, "oat_norm_softmax_loss: bad arguments");
  OAT_REQUIRE(loss != nullptr && scratch != nullptr, "oat_norm_softmax_loss: null loss/scratch");
  cudaStream_t s = as_stream(stream);
  const float inv_t = 1.0f / temperature;
  loss_lse_kernel<<<2 * n, 256, 0, s>>>(sims, n, static_cast<int>(ld), inv_t, scratch, scratch + n);
  loss_grad_logits_kernel<<<n, 256, 0, s>>>(sims, n, static_cast<int>(ld), inv_t, scratch, scratch + n, dsims);
  loss_sum_kernel<<<1, 256, 0, s>>>(scratch, n, loss);
  return check_launch("loss_grad_logits_kernel / loss_sum_kernel");
}

extern "C" size_t oat_infonce_workspace_bytes(int32_t n, int32_t P) {
  // sim workspace | sims [n,n] | dsims [n,n] | lse scratch [2n]
  return sim_ws_bytes(n, n, P) + ((static_cast<size_t>(2) * n * n + 2 * static_cast<size_t>(n)) * 4 + 255 & ~static_cast<size_t>(255));
}

// Fused convenience: sims -> loss -> gradients of the gathered embeddings in one call (same kernels as above).
extern "C" int oat_infonce_fwd_bwd(const float* text, const float* video, int32_t n, int32_t P, float temperature,
                                   float eps, float* sims_out, float* loss, float* dtext, float* dvideo,
                                   void* workspace, size_t workspace_bytes, oat_stream_t stream) {
  OAT_REQUIRE(workspace != nullptr && workspace_bytes >= oat_infonce_workspace_bytes(n, P),
              "oat_infonce_fwd_bwd: workspace too small");
  const size_t simb = sim_ws_bytes(n, n, P);
  float* sims = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + simb);
  float* dsims = sims + static_cast<size_t>(n) * n;
  float* scratch = dsims + static_cast<size_t>(n) * n;
  int rc = oat_sim_matrix_fwd(text, video, n, n, P, eps, sims, workspace, simb, stream);
  if (rc != OAT_OK) return rc;
  if (sims_out != nullptr) {
    cudaError_t e = cudaMemcpyAsync(sims_out, sims, sizeof(float) * n * n, cudaMemcpyDeviceToDevice, as_stream(stream));
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
  }
  const bool want_grad = dtext != nullptr || dvideo != nullptr;
  rc = oat_norm_softmax_loss(sims, n, n, temperature, loss, want_grad ? dsims : nullptr, scratch, stream);
  if (rc != OAT_OK || !want_grad) return rc;
  return oat_sim_matrix_bwd(dsims, n, n, P, eps, dtext, dvideo, workspace, simb, stream);
}

// ---------------------------------------------------------------------------------------------- retrieval ranks
// One CTA per ground-truth pair i: counts over row i (text -> video) and column i (video -> text) of the matrix.
namespace oat {
__global__ void __launch_bounds__(256) retrieval_ranks_kernel(const float* __restrict__ sims, int n, long long ld,
                                                              float* __restrict__ t2v, float* __restrict__ v2t) {
  __shared__ int s_cnt[3];
  const int i = blockIdx.x;
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const float gt = sims[static_cast<long long>(i) * ld + i];
  int row_gt = 0, col_gt = 0, col_eq = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float r = sims[static_cast<long long>(i) * ld + j];
    const float c = sims[static_cast<long long>(j) * ld + i];
    row_gt += r > gt;
    col_gt += c > gt;
    col_eq += c == gt;
  }
  row_gt = __reduce_add_sync(0xffffffffu, row_gt);
  col_gt = __reduce_add_sync(0xffffffffu, col_gt);
  col_eq = __reduce_add_sync(0xffffffffu, col_eq);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt[0], row_gt);
    atomicAdd(&s_cnt[1], col_gt);
    atomicAdd(&s_cnt[2], col_eq);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (t2v != nullptr) t2v[i] = static_cast<float>(s_cnt[0]);
    if (v2t != nullptr) v2t[i] = static_cast<float>(s_cnt[1]) + 0.5f * static_cast<float>(s_cnt[2] - 1);
  }
}
}  // namespace oat

extern "C" int oat_retrieval_ranks(const float* sims, int32_t n, int64_t ld, float* t2v_rank, float* v2t_rank,
                                   oat_stream_t stream) {
  using namespace oat;
  OAT_REQUIRE(sims != nullptr && n > 0 && ld >= n, "oat_retrieval_ranks: bad arguments (n=%d ld=%lld)", n, (long long)ld);
  OAT_REQUIRE(t2v_rank != nullptr || v2t_rank != nullptr, "oat_retrieval_ranks: no output");
  retrieval_ranks_kernel<<<n, 256, 0, as_stream(stream)>>>(sims, n, ld, t2v_rank, v2t_rank);
  return check_launch("retrieval_ranks_kernel");
}
