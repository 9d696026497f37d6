// Skinny GEMM: C[M <= 64, N] = A(M,K) . B(N,K)^T, both operands K-major bf16, fp32 accumulate - the few-row products
// of the path: the split-bf16 CLS rows of the video tower ([B, 3K] x [N, 3K], engine.py), the two projections
// (oa_model.py:68-75: [B, 768] -> [B, 256]) and their input gradients.
//
// With M = batch rows the contraction is a weight STREAM ([N, K] read once, <= 64 output rows): the 128-row tcgen05 tile
// would leave a handful of CTAs pulling the whole weight through one SM each (or need split-K with floating-point
// atomics). Here every CTA owns 16 output columns over the full K: N/16 CTAs stream their 16 weight rows with cp.async
// (128-column chunks). The kernel is a latency chain (one chunk per ring turn), so the CTA runs TWO k-groups of 4 warps,
// each with its own 4-stage ring over every other chunk; the warps multiply on bf16 mma.sync m16n8k16, group 1 hands its
// partial sums to group 0 through shared memory, and the fused epilogue of oat_gemm_bf16 (alpha, bias, column scale,
// GELU + derivative / x aux / ReLU, fp32 residual, accumulate) is applied to the fragments. One owner per output
// element and a fixed summation order: bit-reproducible, no atomics.
#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

namespace {

constexpr int kSkM = 64;          // most rows a problem may have (4 m-tiles of 16)
constexpr int kSkN = 16;          // columns per CTA
constexpr int kSkK = 128;         // k chunk
constexpr int kSkPitch = kSkK + 8;                // bf16 elements: 272-byte rows keep ldmatrix conflict-free
constexpr int kSkStages = 4;      // cp.async ring depth per k-group: three chunks in flight hide the L2 / HBM latency
constexpr int kSkGroups = 2;      // k-groups per CTA: group g streams chunks g, g + 2, ... through its own ring
constexpr int kSkThreads = kSkGroups * 128;

__device__ __forceinline__ void sk_cp16(uint32_t smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(g) : "memory");
}
__device__ __forceinline__ void sk_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sk_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sk_group_sync(int group) {          // the 128 threads of one k-group
  asm volatile("bar.sync %0, 128;\n" ::"r"(group + 1) : "memory");
}
__device__ __forceinline__ void sk_ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sk_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct SkinnyParams {
  const __nv_bfloat16* A; long long lda;
  const __nv_bfloat16* B; long long ldb;
  int M, N, K;
  float alpha;
  const float* bias;
  int scale_cols; float scale;
  int act;
  const __nv_bfloat16* aux; long long ld_aux;
  const float* residual; long long ldr;
  float* out_f32; long long ld_f32;
  __nv_bfloat16* out_bf16; long long ld_bf16;
  __nv_bfloat16* out2; long long ld2;
  int accumulate;
};

// MROWS = 32 or 64 staged A rows. Work split inside a k-group (4 warps): MROWS = 64: warp w owns m-tile w and both 8-column
// tiles; MROWS = 32: warp w owns m-tile (w & 1) and column tile (w >> 1) - half the MMA chain per warp.
template <int MROWS>
__global__ void __launch_bounds__(kSkThreads) skinny_gemm_kernel(const SkinnyParams p) {
  constexpr int kStage = (MROWS + kSkN) * kSkPitch;             // elements per stage
  constexpr int kNT = MROWS == 64 ? 2 : 1;                      // column tiles per warp
  extern __shared__ __align__(16) __nv_bfloat16 sk_smem[];
  const int group = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * kSkN;
  const int chunks = (p.K + kSkK - 1) / kSkK;
  const int my_chunks = chunks > group ? (chunks - group + kSkGroups - 1) / kSkGroups : 0;   // chunk j of the group = group + j * kSkGroups
  __nv_bfloat16* ring = sk_smem + group * (kSkStages * kStage);
  const int mt = MROWS == 64 ? warp : (warp & 1);
  const int nt0 = MROWS == 64 ? 0 : (warp >> 1);

  auto stage = [&](int j, int s) {
    __nv_bfloat16* As = ring + s * kStage;
    __nv_bfloat16* Bs = As + MROWS * kSkPitch;
    const int k0 = (group + j * kSkGroups) * kSkK;
    // (MROWS + 16) rows x 16 vectors of 8 bf16
    for (int idx = tid; idx < (MROWS + kSkN) * (kSkK / 8); idx += 128) {
      const int r = idx / (kSkK / 8), v = idx - r * (kSkK / 8);
      const int k = k0 + v * 8;
      const bool is_a = r < MROWS;
      const int row = is_a ? r : n0 + (r - MROWS);
      __nv_bfloat16* dst = (is_a ? As + r * kSkPitch : Bs + (r - MROWS) * kSkPitch) + v * 8;
      const bool ok = is_a ? row < p.M : row < p.N;        // rows beyond M / N stay zero from the clear below
      if (!ok) continue;
      if (k < p.K) {
        sk_cp16(smem_u32(dst), (is_a ? p.A + static_cast<long long>(row) * p.lda : p.B + static_cast<long long>(row) * p.ldb) + k);
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);      // k tail of the last chunk
      }
    }
  };
  // rows that no load ever writes (A rows >= M, B rows >= N) must read as zero: clear the group's ring once
  for (int idx = tid; idx < kSkStages * kStage / 8; idx += 128) reinterpret_cast<uint4*>(ring)[idx] = make_uint4(0, 0, 0, 0);
  sk_group_sync(group);

  float acc[kNT][4];
#pragma unroll
  for (int i = 0; i < kNT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

#pragma unroll
  for (int j = 0; j < kSkStages - 1; ++j) {
    if (j < my_chunks) stage(j, j);
    sk_commit();
  }
  for (int j = 0; j < my_chunks; ++j) {
    const int s = j % kSkStages;
    sk_wait<kSkStages - 2>();                 // chunk j has landed (groups complete in order)
    sk_group_sync(group);                     // ... for every thread of the k-group, and the stage refilled below is no longer being read
    if (j + kSkStages - 1 < my_chunks) stage(j + kSkStages - 1, (j + kSkStages - 1) % kSkStages);
    sk_commit();
    if (mt * 16 < p.M) {
      const uint32_t a_base = smem_u32(ring + s * kStage);
      const uint32_t b_base = a_base + MROWS * kSkPitch * 2;
#pragma unroll
      for (int ks = 0; ks < kSkK / 16; ++ks) {
        uint32_t a[4], b[4];
        sk_ldsm4(a_base + ((mt * 16 + (lane & 15)) * kSkPitch + ks * 16 + (lane >> 4) * 8) * 2, a);
        // B rows = output columns: matrices (n 0-7, k lo), (n 0-7, k hi), (n 8-15, k lo), (n 8-15, k hi)
        sk_ldsm4(b_base + (((lane & 7) + ((lane >> 4) << 3)) * kSkPitch + ks * 16 + ((lane >> 3) & 1) * 8) * 2, b);
        if constexpr (kNT == 2) {
          sk_mma(acc[0], a, b[0], b[1]);
          sk_mma(acc[1], a, b[2], b[3]);
        } else {
          sk_mma(acc[0], a, nt0 == 0 ? b[0] : b[2], nt0 == 0 ? b[1] : b[3]);
        }
      }
    }
  }

  // ---- k-groups > 0 hand their partial sums to group 0 through shared memory (fixed order: bit-reproducible)
  __syncthreads();                            // every ring is dead from here on
  float* red = reinterpret_cast<float*>(sk_smem);               // [kSkGroups - 1][128][kNT * 4]
  if (group > 0) {
#pragma unroll
    for (int i = 0; i < kNT; ++i)
      *reinterpret_cast<float4*>(red + (((group - 1) * 128 + tid) * kNT + i) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  if (group > 0) return;
#pragma unroll
  for (int g2 = 1; g2 < kSkGroups; ++g2) {
#pragma unroll
    for (int i = 0; i < kNT; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(red + (((g2 - 1) * 128 + tid) * kNT + i) * 4);
      acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
    }
  }

  // ---- epilogue on the fragments: lane (g, t) holds rows g / g+8 and columns 2t, 2t+1 of each 8-column tile
  if (mt * 16 >= p.M) return;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) {
    const int col = n0 + (nt0 + nt) * 8 + 2 * t;
    if (col >= p.N) continue;                    // N is a multiple of 4 and col is even: col + 1 < N as well
    float b0 = 0.f, b1 = 0.f;
    if (p.bias != nullptr) { b0 = p.bias[col]; b1 = p.bias[col + 1]; }
    const float sc = col < p.scale_cols ? p.scale : 1.0f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = mt * 16 + g + half * 8;
      if (row >= p.M) continue;
      float x0 = (acc[nt][2 * half] * p.alpha + b0) * sc, x1 = (acc[nt][2 * half + 1] * p.alpha + b1) * sc;
      if (p.act == 1) {
        float d0, d1;
        gelu_fwd_grad(x0, x0, d0);
        gelu_fwd_grad(x1, x1, d1);
        *reinterpret_cast<uint32_t*>(p.out2 + static_cast<long long>(row) * p.ld2 + col) = pack_bf16x2(d0, d1);
      } else if (p.act == 2) {
        const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.aux + static_cast<long long>(row) * p.ld_aux + col));
        x0 *= a.x; x1 *= a.y;
      } else if (p.act == 3) {
        x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f);
      }
      if (p.residual != nullptr) {
        const float2 r = *reinterpret_cast<const float2*>(p.residual + static_cast<long long>(row) * p.ldr + col);
        x0 += r.x; x1 += r.y;
      }
      if (p.out_f32 != nullptr) {
        float2* o = reinterpret_cast<float2*>(p.out_f32 + static_cast<long long>(row) * p.ld_f32 + col);
        if (p.accumulate) { const float2 old = *o; x0 += old.x; x1 += old.y; }
        *o = make_float2(x0, x1);
      }
      if (p.out_bf16 != nullptr)
        *reinterpret_cast<uint32_t*>(p.out_bf16 + static_cast<long long>(row) * p.ld_bf16 + col) = pack_bf16x2(x0, x1);
    }
  }
}

}  // namespace

// K-major x K-major problems with at most 64 rows whose tensors allow the vector accesses above.
bool skinny_gemm_supported(const oat_gemm_args* a) {
  if (a->M > kSkM || a->a_major != 0 || a->b_major != 0) return false;
  if (a->K % 8 != 0 || a->lda % 8 != 0 || a->ldb % 8 != 0 || a->N % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) != 0 || (reinterpret_cast<uintptr_t>(a->B) & 15) != 0) return false;
  // accumulate products (the [hi | lo] x [lo | hi] corrections) stream a long K into a few columns: split-K over the whole
  // chip on the tcgen05 kernel is ~10x faster there (measured: 1.5 vs 37 us for K = 6144, N = 768), at the price of
  // floating-point atomics
  if (a->accumulate) return false;
  if (a->act == 4) return false;       // row dots live in the tcgen05 kernel's epilogue
  auto even = [](const void* p, long long ld, int esz) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % (2 * esz)) == 0 && ld % 2 == 0); };
  return even(a->out_f32, a->ld_f32, 4) && even(a->residual, a->ldr, 4) && even(a->out_bf16, a->ld_bf16, 2) &&
         even(a->out2_bf16, a->ld2, 2) && even(a->aux_bf16, a->ld_aux, 2);
}

int launch_skinny_gemm(const oat_gemm_args* a, cudaStream_t stream) {
  SkinnyParams p;
  p.A = reinterpret_cast<const __nv_bfloat16*>(a->A); p.lda = a->lda;
  p.B = reinterpret_cast<const __nv_bfloat16*>(a->B); p.ldb = a->ldb;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.alpha = a->alpha; p.bias = a->bias; p.scale_cols = a->scale_cols; p.scale = a->scale; p.act = a->act;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux_bf16); p.ld_aux = a->ld_aux;
  p.residual = a->residual; p.ldr = a->ldr;
  p.out_f32 = a->out_f32; p.ld_f32 = a->ld_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16); p.ld_bf16 = a->ld_bf16;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2_bf16); p.ld2 = a->ld2;
  p.accumulate = a->accumulate;
  const int grid = (a->N + kSkN - 1) / kSkN;
  if (a->M <= 32) {
    constexpr int smem = kSkGroups * kSkStages * (32 + kSkN) * kSkPitch * 2;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "skinny_gemm smem attr: %s", cudaGetErrorString(e));
      attr_set = true;
    }
    skinny_gemm_kernel<32><<<grid, kSkThreads, smem, stream>>>(p);
  } else {
    constexpr int smem = kSkGroups * kSkStages * (64 + kSkN) * kSkPitch * 2;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "skinny_gemm smem attr: %s", cudaGetErrorString(e));
      attr_set = true;
    }
    skinny_gemm_kernel<64><<<grid, kSkThreads, smem, stream>>>(p);
  }
  return check_launch("skinny_gemm_kernel");
}

}  // namespace oat
