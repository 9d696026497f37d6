// Space attention forward on the 5th-generation tensor cores (tcgen05 + TMEM), with the global CLS query fused in.
//
// Reference: VarAttention.forward, OATrans/model/video_transformer.py:99-135, '(b f) n d' grouping (:112, pattern
// :275-278): every token of frame f attends to [CLS] + the n tokens of its own frame; the CLS query attends to
// every key (:108-110). One attention group = (batch b, frame f, head h): n queries, n + 1 keys, head dim 64.
//
// Persistent kernel, one CTA per SM, 10 warps:
//   warp 0      TMA producer: Q / K / V head slices of the group (n rows x 128 B each, straight out of the qkv GEMM
//               output, 128B-swizzled) into a 2-stage shared-memory ring; the CLS token's q / k / v rows are appended
//               as row n of each matrix (so the CLS query rides along as one more query row and the CLS key as one
//               more key - nothing is materialised n times as the reference does at :115-119)
//   warp 1      tcgen05.mma issuer: S = Q K^T per 128-query tile into TMEM (fp32, N = keys rounded up to 16), then
//               O = P V with the A operand (P, bf16) read straight from TMEM where the softmax warps left it
//   warps 2-5   softmax + epilogue of query tile 0 (one thread per query row, TMEM lane = row)
//   warps 6-9   softmax + epilogue of query tile 1
// TMEM (512 columns): per tile a 256-column region; S occupies [0, nkp), P overwrites S in place as packed bf16 in
// [0, nkp/2), O accumulates in [128, 192) once S is dead. The whole key row of a query lives in TMEM, so the softmax is
// exact two-pass (row max, then exp / sum) - no online rescaling. While one tile is in softmax the tensor core works on
// the other tile / the next group.
//
// The CLS query (row n of tile 1) produces one partial (max, sum, unnormalised O) per (b, h, f); a small combine
// kernel merges the F partials (the CLS key is counted by frame 0 only).
#include <cuda.h>
#include <stdlib.h>

#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box1);

namespace {

constexpr int SD = 64;                         // head dim
constexpr int kSpThreads = 320;
constexpr int kTileBytes = 128 * 128;          // 128 rows x 128 B
constexpr int kMatBytes = 256 * 128;           // up to 256 rows of one operand
constexpr int kStageBytes = 3 * kMatBytes;     // Q | K | V
constexpr int kSpSmem = 2 * kStageBytes + 1024 + 256;
constexpr int kClsPart = 2 + SD;               // (max2, sum, o[64]) per (b, h, f)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct SpaceGeom {
  int B, T, H, F, n, nk, nkp, groups;
  long long ld_qkv, ld_out;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  float* lse;
  float* cls_part;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ld32_wait(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld_32x32b_x32(taddr, v);
  tmem_ld_wait();
}

// Dynamic group scheduler (forward and pipelined backward): groups are handed out by an atomic counter instead of a
// fixed stride - with a fixed stride the slowest SMs (whole TPCs ~20 % behind the median: memory-side placement) set the
// kernel time (measured on the backward: 417-439 k cycles for the slowest CTA against a median of 343 k). Each launch
// takes the next {next, done} counter pair of a small pool, so launches in flight on different streams never share
// one; the last CTA of a launch to finish re-arms its pair (no memset between launches).
constexpr int kSchedSlots = 32;
__device__ unsigned int g_sched[2][kSchedSlots][2];      // [forward | backward][slot][next, done]
static unsigned int next_sched_slot(int which) {
  static unsigned int counters[2] = {0, 0};
  return __atomic_fetch_add(&counters[which], 1u, __ATOMIC_RELAXED) % kSchedSlots;
}

template <int KS>   // KS = 16-key steps of the P V product = padded key count / 16
__global__ void __launch_bounds__(kSpThreads, 1)
attn_space_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const SpaceGeom G, const int sched_slot) {
  unsigned int* const sched = g_sched[0][sched_slot];
  constexpr int NCH = (KS + 1) / 2;            // 32-key chunks of a score row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes);
  uint64_t* full = bars;           // [2] producer -> MMA
  uint64_t* empty = bars + 2;      // [2] MMA + 8 softmax warps -> producer
  uint64_t* s_full = bars + 4;     // [2] MMA -> softmax (S tile ready)
  uint64_t* p_full = bars + 6;     // [2] softmax -> MMA (P in TMEM)
  uint64_t* o_full = bars + 8;     // [2] MMA -> softmax (O ready)
  uint64_t* buf_free = bars + 10;  // [2] softmax -> MMA (TMEM region drained)
  uint64_t* sched_full = bars + 12; // [4] producer -> everyone: group index of an iteration published
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  int* sched_g = reinterpret_cast<int*>(tmem_slot + 1);   // [4] ring of group indices (-1: no more work)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HDIM = G.H * SD;
  pdl_launch_dependents();

  // operand rows that no load ever writes (key rows nk..nkp, query rows n+1..255) must be finite: zero everything once
  for (int i = tid; i < 2 * kStageBytes / 16; i += kSpThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) tma_prefetch_desc(&tmap);
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 9);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&buf_free[i], 4);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&sched_full[i], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                   // shared memory cleared, barriers and TMEM ready: now wait for the qkv GEMM
  const uint32_t tmem_base = *tmem_slot;
  // Groups are handed out by an atomic counter (see g_sched); the producer draws them one iteration ahead and publishes
  // them in a small ring.
  auto group_of = [&](int i) -> int {
    mbar_wait(&sched_full[i & 3], (i >> 2) & 1);
    return sched_g[i & 3];
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    auto fetch = [&](int i) -> int {
      int g = 0;
      if (lane == 0) {
        g = static_cast<int>(atomicAdd(&sched[0], 1u));
        if (g >= G.groups) g = -1;
        sched_g[i & 3] = g;
        mbar_arrive(&sched_full[i & 3]);
      }
      return __shfl_sync(0xffffffffu, g, 0);
    };
    int g = fetch(0);
    for (int i = 0; g >= 0; ++i) {
      const int g_next = fetch(i + 1);
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      const int row0 = b * G.T + 1 + f * G.n;
      uint8_t* qs = smem + s * kStageBytes;
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[s], 3u * G.n * 128u);
        tma_load_2d(qs, &tmap, &full[s], h * SD, row0);
        tma_load_2d(qs + kMatBytes, &tmap, &full[s], HDIM + h * SD, row0);
        tma_load_2d(qs + 2 * kMatBytes, &tmap, &full[s], 2 * HDIM + h * SD, row0);
      }
      if (lane < 24) {
        // the CLS token's q / k / v head slices become row n of each operand (same 128B swizzle as the TMA rows)
        const int m = lane >> 3, c = lane & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(G.qkv + static_cast<long long>(b) * G.T * G.ld_qkv +
                                                          m * HDIM + h * SD + c * 8);
        *reinterpret_cast<uint4*>(qs + m * kMatBytes + G.n * 128 + ((c ^ (G.n & 7)) << 4)) = val;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      g = g_next;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane issues everything)
    // Literal TMEM addresses (this CTA owns all 512 columns, so the allocation starts at 0) and an elected block keep
    // every tcgen05.mma operand in uniform registers: the 15 P V instructions of a tile go out back to back.
    if (tmem_base != 0) __trap();
    constexpr uint32_t idesc_s = make_idesc_bf16(128, static_cast<uint32_t>(KS * 16), 0u, 0u);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, SD, 0u, 1u);   // B = V is MN-major ([key][d], d contiguous)
    auto issue_pv = [&](int j, uint32_t v_addr, uint32_t par, int release_stage) {
      mbar_wait(&p_full[j], par);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = j * 256 + 128;
        const uint32_t a_tmem = j * 256;
        const uint64_t bdesc = make_smem_desc_sw128(v_addr, kMatBytes, 1024);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          tc_mma_bf16_ts(d_tmem, a_tmem + ks * 8, bdesc + ks * 128, idesc_o, ks > 0 ? 1u : 0u);
        tc_commit(&o_full[j]);
        if (release_stage >= 0) tc_commit(&empty[release_stage]);
      }
      __syncwarp();
    };
    uint32_t prev_v = 0;
    int prev_stage = -1;
    int i = 0;
    for (;; ++i) {
      if (group_of(i) < 0) break;
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1, par = i & 1;
      const uint32_t q_addr = smem_u32(smem + s * kStageBytes);
      const uint32_t k_addr = q_addr + kMatBytes, v_addr = k_addr + kMatBytes;
      mbar_wait(&full[s], ph);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        mbar_wait(&buf_free[j], par ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc_sw128(q_addr + j * kTileBytes, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(k_addr, 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(j * 256, adesc + k * 2, bdesc + k * 2, idesc_s, k > 0 ? 1u : 0u);
          tc_commit(&s_full[j]);
        }
        __syncwarp();
        if (j == 0) {
          if (i > 0) issue_pv(1, prev_v, (i - 1) & 1, prev_stage);
        } else {
          issue_pv(0, v_addr, par, -1);
        }
      }
      prev_v = v_addr;
      prev_stage = s;
    }
    if (i > 0) issue_pv(1, prev_v, (i - 1) & 1, prev_stage);
  } else {
    // ------------------------------------------------------------------ softmax + epilogue, one thread per query row
    const int j = (warp - 2) >> 2;             // query tile = TMEM region
    const int q = warp & 3;                    // TMEM lane quarter this warp may address
    const int r_tile = q * 32 + lane;
    const int r = j * 128 + r_tile;            // query index inside the group; r == n is the CLS query
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 256;
    for (int i = 0;; ++i) {
      const int g = group_of(i);
      if (g < 0) break;
      const int s = i & 1;
      const uint32_t par = i & 1;
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      // keys: [frame tokens 0..n-1, CLS]; the CLS query counts the CLS key in frame 0 only
      const int limit = (r == G.n && f != 0) ? G.n : G.nk;
      mbar_wait(&s_full[j], par);
      tc_fence_after();
      // ONE pass over the score row. Reading S out of TMEM is what bounds this kernel (64 B/clk per SM: a 128 x 240 fp32
      // tile costs ~1.9 k clk per pass), so the row maximum is not found in a pass of its own: chunk by chunk (32 keys),
      // p = 2^(s*log2e - m2) against a running reference m2 that starts as the maximum of the first chunk and is only
      // moved when a later chunk exceeds it by more than 2^8 (then the few P chunks already written are re-scaled in
      // TMEM - rare, and exact: softmax is shift-invariant and bf16 / fp32 carry the exponent). P never exceeds 2^8.
      // A tcgen05.ld takes ~220 clk to come back: the next chunk's load is issued right after the wait for the current
      // one (tcgen05.wait::ld has no groups), so it is in flight while the current chunk is computed.
      float m2 = -INFINITY, sum = 0.f;
      auto chunk = [&](const int c, uint32_t (&v)[32]) {
        float cm = -INFINITY;
        if (c < NCH - 1) {
#pragma unroll
          for (int e = 0; e < 32; ++e) cm = fmaxf(cm, __uint_as_float(v[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) cm = fmaxf(cm, (c * 32 + e < limit) ? __uint_as_float(v[e]) : -INFINITY);
        }
        const float cm2 = cm * kLog2e;
        if (c == 0) {
          m2 = cm2;
        } else {
          const bool need = cm2 > m2 + 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float m2n = need ? cm2 : m2;
            const float sc = ex2_approx(m2 - m2n);          // exactly 1 for the rows that keep their reference
            sum *= sc;
            tmem_st_wait();
#pragma unroll
            for (int cc = 0; cc < c; ++cc) {
              uint32_t old[16];
              tmem_ld_32x32b_x16(t_row + cc * 16, old);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float2 f = unpack_bf16x2(old[e]);
                old[e] = pack_bf16x2(f.x * sc, f.y * sc);
              }
              tmem_st_32x32b_x16(t_row + cc * 16, old);
            }
            m2 = m2n;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), kLog2e, -m2));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), kLog2e, -m2));
          if (c == NCH - 1) {
            if (c * 32 + e >= limit) p0 = 0.f;
            if (c * 32 + e + 1 >= limit) p1 = 0.f;
          }
          sum += p0 + p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(t_row + c * 16, pk);
      };
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(t_row, va);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();
        if (c & 1) {
          if (c + 1 < NCH) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, va);
          chunk(c, vb);
        } else {
          if (c + 1 < NCH) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, vb);
          chunk(c, va);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j]);

      // epilogue: O (unnormalised) out of TMEM, then hand the region back to the MMA warp
      mbar_wait(&o_full[j], par);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      ld32_wait(t_row + 128, o0);
      ld32_wait(t_row + 160, o1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&buf_free[j]);

      if (r == G.n) {
        float* pp = G.cls_part + ((static_cast<long long>(b) * G.H + h) * G.F + f) * kClsPart;
        pp[0] = m2;
        pp[1] = sum;
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          pp[2 + d] = __uint_as_float(o0[d]);
          pp[34 + d] = __uint_as_float(o1[d]);
        }
      }
      const float inv = 1.0f / sum;
      // stage this row as bf16 in the (dead) Q tile of the stage, then write whole 128-byte rows coalesced
      uint8_t* stg = smem + s * kStageBytes + j * kTileBytes;
      {
        uint8_t* rowp = stg + r_tile * 128;
        const int sw = r_tile & 7;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o0[8 * c + 0]) * inv, __uint_as_float(o0[8 * c + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = w;
          w.x = pack_bf16x2(__uint_as_float(o1[8 * c + 0]) * inv, __uint_as_float(o1[8 * c + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o1[8 * c + 2]) * inv, __uint_as_float(o1[8 * c + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o1[8 * c + 4]) * inv, __uint_as_float(o1[8 * c + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o1[8 * c + 6]) * inv, __uint_as_float(o1[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(rowp + (((c + 4) ^ sw) << 4)) = w;
        }
      }
      __syncwarp();
      const long long tok0 = static_cast<long long>(b) * G.T + 1 + f * G.n;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = q * 32 + it * 4 + (lane >> 3), ch = lane & 7;
        if (j * 128 + rr < G.n) {
          const uint4 w = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
          *reinterpret_cast<uint4*>(G.out + (tok0 + j * 128 + rr) * G.ld_out + h * SD + ch * 8) = w;
        }
      }
      if (r < G.n && G.lse != nullptr)
        G.lse[(static_cast<long long>(b) * G.H + h) * G.T + 1 + f * G.n + r] = (m2 + log2f(sum)) * kLn2;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {   // last CTA out re-arms the pair for a later launch
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// merge the F per-frame partials of the CLS query: out row 0 of every (b, h)
__global__ void __launch_bounds__(SD) attn_cls_combine_kernel(const SpaceGeom G) {
  const int bh = blockIdx.x, b = bh / G.H, h = bh % G.H, d = threadIdx.x;
  const float* part = G.cls_part + static_cast<long long>(bh) * G.F * kClsPart;
  float M = -INFINITY;
  for (int f = 0; f < G.F; ++f) M = fmaxf(M, part[f * kClsPart]);
  float L = 0.f, o = 0.f;
  for (int f = 0; f < G.F; ++f) {
    const float w = exp2f(part[f * kClsPart] - M);
    L = fmaf(part[f * kClsPart + 1], w, L);
    o = fmaf(part[f * kClsPart + 2 + d], w, o);
  }
  G.out[static_cast<long long>(b) * G.T * G.ld_out + h * SD + d] = __float2bfloat16_rn(o / L);
  if (d == 0 && G.lse != nullptr) G.lse[static_cast<long long>(bh) * G.T] = (M + log2f(L)) * kLn2;
}

template <int KS>
int launch_space_tc(const CUtensorMap& tm, const SpaceGeom& G, cudaStream_t s) {
  auto kern = attn_space_tc_fwd_kernel<KS>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_space_tc smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  const int sms = num_sms();
  const int grid = G.groups < sms ? G.groups : sms;
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kSpThreads), kSpSmem, s, tm, G, static_cast<int>(next_sched_slot(0)));
  if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_space_tc_fwd_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("attn_space_tc_fwd_kernel");
}


// =============================================================================================== backward
// Key-major ("transposed") formulation so that the three gradient contractions whose M dimension is the key index read
// their A operand without a transpose: per key tile kt (128 keys, TMEM lane = key) and query half qh (128 queries)
//   S^T  = K[kt] Q[qh]^T            TMEM cols [  0,128)      dP^T = V[kt] dO[qh]^T      TMEM cols [128,256)
//   math (8 warps, thread = key row, 64 query columns each):  P^T = 2^(S^T log2e - lse2[q]),  dS^T = P^T (dP^T - delta[q])
//        P^T  -> TMEM (bf16, in place over the S^T columns already consumed)
//        dS^T -> shared memory, [key][query] rows of 64 queries (128 B), 128B-swizzled: read by the tensor core both as a
//                K-major A operand (dK) and as an MN-major A operand (dQ)
//   dV[kt]  += P^T  dO[qh]     (A from TMEM, B = dO rows MN-major)     TMEM cols [256,320)
//   dK[kt]  += dS^T Q[qh]      (A K-major smem, B = Q rows MN-major)   TMEM cols [320,384)
//   dQ[qh]  += dS   K[kt]      (A MN-major smem, B = K rows MN-major)  TMEM cols [384,448) / [448,512)
// The CLS query is query row n, the CLS key is key row n; their gradients are reduced across the F groups of a (b, h)
// with fp32 atomics into cls_acc (same contract as the mma.sync kernel: dq already scaled). Rows beyond the valid ones
// are zero in shared memory, so they contribute nothing and need no masking; the only masked cell is the
// (CLS query, CLS key) pair, which frame 0 alone accounts for.
struct SpaceBwdGeom {
  int B, T, H, F, n, groups;
  long long ld_qkv, ld_out, ld_dout, ld_dqkv;
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  const float* lse;
  __nv_bfloat16* dqkv;
  float* cls_acc;
  float scale;
  const float* delta;     // optional [H][ld_delta]: rowsum(dO * O) from the dO-producing GEMM (pipelined kernel only)
  long long ld_delta;
};

constexpr int kBwdDsBytes = 2 * kTileBytes;                       // dS^T of one unit: 2 blocks of [128 keys x 64 queries]
constexpr int kBwdOperandBytes = 4 * kMatBytes;                   // Q | K | V | dO
constexpr int kBwdSmem = kBwdOperandBytes + kBwdDsBytes + 2 * 256 * 4 + 1024 + 256;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kSpThreads, 1)
attn_space_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                         const SpaceBwdGeom G) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + kMatBytes;
  uint8_t* Vs = smem + 2 * kMatBytes;
  uint8_t* Ds = smem + 3 * kMatBytes;                 // dO
  uint8_t* dSs = smem + kBwdOperandBytes;             // dS^T unit buffer / epilogue staging
  float* lse2_s = reinterpret_cast<float*>(dSs + kBwdDsBytes);   // [256]
  float* del_s = lse2_s + 256;                                   // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(del_s + 256);
  uint64_t* full = bars;           // producer -> MMA + math        (per group)
  uint64_t* empty = bars + 1;      // MMA -> producer               (per group)
  uint64_t* st_full = bars + 2;    // MMA -> math: S^T, dP^T ready  (per unit)
  uint64_t* math_done = bars + 3;  // math -> MMA: P^T, dS^T ready  (per unit)
  uint64_t* acc_full = bars + 4;   // MMA -> math: dV, dK complete  (per key tile)
  uint64_t* acc_free = bars + 5;   // math -> MMA: dV, dK drained   (per key tile)
  uint64_t* dq_full = bars + 6;    // MMA -> math: dQ complete      (per group)
  uint64_t* dq_free = bars + 7;    // math -> MMA: dQ drained       (per group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HDIM = G.H * SD;
  const int n = G.n;
  pdl_launch_dependents();

  for (int i = tid; i < (kBwdOperandBytes + kBwdDsBytes) / 16; i += kSpThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmap_qkv); tma_prefetch_desc(&tmap_do); }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  } else if (warp == 1 && lane == 0) {
    mbar_init(full, 2);
    mbar_init(empty, 1);
    mbar_init(st_full, 1);
    mbar_init(math_done, 8);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 8);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    int i = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++i) {
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      const int row0 = b * G.T + 1 + f * n;
      mbar_wait(empty, (i & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(full, 4u * n * 128u);
        tma_load_2d(Qs, &tmap_qkv, full, h * SD, row0);
        tma_load_2d(Ks, &tmap_qkv, full, HDIM + h * SD, row0);
        tma_load_2d(Vs, &tmap_qkv, full, 2 * HDIM + h * SD, row0);
        tma_load_2d(Ds, &tmap_do, full, h * SD, row0);
      }
      {
        // CLS token rows (q, k, v, dO) -> row n of each operand
        const int m = lane >> 3, c = lane & 7;
        const __nv_bfloat16* src = (m < 3)
            ? G.qkv + static_cast<long long>(b) * G.T * G.ld_qkv + m * HDIM + h * SD + c * 8
            : G.dout + static_cast<long long>(b) * G.T * G.ld_dout + h * SD + c * 8;
        const uint4 val = *reinterpret_cast<const uint4*>(src);
        *reinterpret_cast<uint4*>(smem + m * kMatBytes + n * 128 + ((c ^ (n & 7)) << 4)) = val;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_st = make_idesc_bf16(128, 128, 0u, 0u);   // S^T, dP^T: both operands K-major
    const uint32_t idesc_kn = make_idesc_bf16(128, SD, 0u, 1u);    // dV, dK: A K-major (TMEM / smem), B MN-major
    const uint32_t idesc_mn = make_idesc_bf16(128, SD, 1u, 1u);    // dQ: A MN-major smem, B MN-major
    const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs), d_addr = smem_u32(Ds);
    const uint32_t ds_addr = smem_u32(dSs);
    const uint32_t t_st = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 320,
                   t_dq = tmem_base + 384;
    int i = 0;
    uint32_t unit = 0, ktile = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++i) {
      mbar_wait(full, i & 1);
      tc_fence_after();
      for (int kt = 0; kt < 2; ++kt, ++ktile) {
        for (int qh = 0; qh < 2; ++qh, ++unit) {
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(t_st, make_smem_desc_sw128(k_addr + kt * kTileBytes + k * 32, 0, 1024),
                          make_smem_desc_sw128(q_addr + qh * kTileBytes + k * 32, 0, 1024), idesc_st, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(t_dp, make_smem_desc_sw128(v_addr + kt * kTileBytes + k * 32, 0, 1024),
                          make_smem_desc_sw128(d_addr + qh * kTileBytes + k * 32, 0, 1024), idesc_st, k > 0 ? 1u : 0u);
            tc_commit(st_full);
          }
          __syncwarp();
          if (qh == 0) {                       // first accumulation into dV / dK of this key tile overwrites them
            mbar_wait(acc_free, (ktile & 1) ^ 1);
            if (kt == 0) mbar_wait(dq_free, (i & 1) ^ 1);
          }
          mbar_wait(math_done, unit & 1);
          tc_fence_after();
          if (lane == 0) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {   // dV += P^T dO : k = 16 queries per step
              const uint32_t a_t = t_st + (ks < 4 ? ks * 8 : 64 + (ks - 4) * 8);
              tc_mma_bf16_ts(t_dv, a_t, make_smem_desc_sw128(d_addr + (qh * 128 + ks * 16) * 128, kMatBytes, 1024),
                             idesc_kn, (qh > 0 || ks > 0) ? 1u : 0u);
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {   // dK += dS^T Q
              const uint64_t adesc = make_smem_desc_sw128(ds_addr + (ks >> 2) * kTileBytes + (ks & 3) * 32, 0, 1024);
              tc_mma_bf16(t_dk, adesc, make_smem_desc_sw128(q_addr + (qh * 128 + ks * 16) * 128, kMatBytes, 1024),
                          idesc_kn, (qh > 0 || ks > 0) ? 1u : 0u);
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {   // dQ += dS K : k = 16 keys per step, A = dS^T read MN-major
              const uint64_t adesc = make_smem_desc_sw128(ds_addr + ks * 2048, kTileBytes, 1024);
              tc_mma_bf16(t_dq + qh * SD, adesc,
                          make_smem_desc_sw128(k_addr + (kt * 128 + ks * 16) * 128, kMatBytes, 1024), idesc_mn,
                          (kt > 0 || ks > 0) ? 1u : 0u);
            }
            if (qh == 1) {
              tc_commit(acc_full);
              if (kt == 1) { tc_commit(dq_full); tc_commit(empty); }
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ math warps
    const int mw = warp - 2;                  // 0..7
    const int q4 = warp & 3;                  // TMEM lane quarter
    const int hh = mw >> 2;                   // which 64-query half of a unit / which accumulator in the epilogues
    const int r_tile = q4 * 32 + lane;        // key row inside the key tile (math) / row inside the tile (epilogues)
    const uint32_t lane_off = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t t_st = tmem_base + lane_off + hh * 64, t_dp = tmem_base + lane_off + 128 + hh * 64;
    uint8_t* ds_row = dSs + hh * kTileBytes + r_tile * 128;
    const int sw = r_tile & 7;
    uint8_t* stg = dSs + mw * 4096;           // this warp's 32 x 128 B staging rows
    const int mt = mw * 32 + lane;            // 0..255: query row this thread prepares lse / delta for
    int i = 0;
    uint32_t unit = 0, ktile = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++i) {
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      const long long tok_base = static_cast<long long>(b) * G.T;
      const long long tok0 = tok_base + 1 + f * n;
      mbar_wait(full, i & 1);
      // ---- delta[q] = dO[q] . O[q], lse2[q] = lse[q] * log2e for the 256 query rows (row n = CLS query)
      {
        float dl = 0.f, l2 = 0.f;
        if (mt <= n) {
          const long long tok = (mt == n) ? tok_base : tok0 + mt;
          const __nv_bfloat16* op = G.out + tok * G.ld_out + h * SD;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 o = *reinterpret_cast<const uint4*>(op + c * 8);
            const uint4 d = *reinterpret_cast<const uint4*>(Ds + mt * 128 + ((c ^ (mt & 7)) << 4));
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 x = unpack_bf16x2(ow[k]), y = unpack_bf16x2(dw[k]);
              dl = fmaf(x.x, y.x, dl);
              dl = fmaf(x.y, y.y, dl);
            }
          }
          l2 = G.lse[(static_cast<long long>(b) * G.H + h) * G.T + (tok - tok_base)] * kLog2e;
        }
        del_s[mt] = dl;
        lse2_s[mt] = l2;
      }
      named_bar_sync(1, 256);

      for (int kt = 0; kt < 2; ++kt, ++ktile) {
        const int key = kt * 128 + r_tile;
        // the (CLS query, CLS key) cell is counted by frame 0 only
        const int kill = (key == n && f != 0) ? n : -1;
        for (int qh = 0; qh < 2; ++qh, ++unit) {
          mbar_wait(st_full, unit & 1);
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32b_x32(t_st + cc * 32, sv);
            tmem_ld_32x32b_x32(t_dp + cc * 32, dv);
            tmem_ld_wait();
            const int qa0 = qh * 128 + hh * 64 + cc * 32;
            float pv[32], dsv[32];
#pragma unroll
            for (int e4 = 0; e4 < 32; e4 += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(lse2_s + qa0 + e4);
              const float4 d4 = *reinterpret_cast<const float4*>(del_s + qa0 + e4);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float p = ex2_approx(fmaf(__uint_as_float(sv[e4 + k]), kLog2e, -ls[k]));
                pv[e4 + k] = p;
                dsv[e4 + k] = p * (__uint_as_float(dv[e4 + k]) - dl[k]);
              }
            }
            if (kill >= qa0 && kill < qa0 + 32) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (qa0 + e == kill) { pv[e] = 0.f; dsv[e] = 0.f; }
            }
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(pv[2 * e], pv[2 * e + 1]);
            tmem_st_32x32b_x16(t_st + cc * 16, pk);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint4 w;
              w.x = pack_bf16x2(dsv[8 * c4 + 0], dsv[8 * c4 + 1]);
              w.y = pack_bf16x2(dsv[8 * c4 + 2], dsv[8 * c4 + 3]);
              w.z = pack_bf16x2(dsv[8 * c4 + 4], dsv[8 * c4 + 5]);
              w.w = pack_bf16x2(dsv[8 * c4 + 6], dsv[8 * c4 + 7]);
              *reinterpret_cast<uint4*>(ds_row + (((cc * 4 + c4) ^ sw) << 4)) = w;
            }
          }
          tmem_st_wait();
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(math_done);
        }
        // ---- dV (hh == 0) / dK (hh == 1) of this key tile: TMEM -> bf16 rows -> global, CLS key row -> atomics
        mbar_wait(acc_full, ktile & 1);
        tc_fence_after();
        {
          uint32_t a0[32], a1[32];
          const uint32_t t_acc = tmem_base + lane_off + 256 + hh * 64;
          tmem_ld_32x32b_x32(t_acc, a0);
          tmem_ld_32x32b_x32(t_acc + 32, a1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
          if (key == n && G.cls_acc != nullptr) {
            float* acc = G.cls_acc + (static_cast<long long>(b) * G.H + h) * 3 * SD + (hh == 0 ? 2 * SD : SD);
#pragma unroll
            for (int d = 0; d < 32; ++d) {
              atomicAdd(acc + d, __uint_as_float(a0[d]));
              atomicAdd(acc + 32 + d, __uint_as_float(a1[d]));
            }
          }
          uint8_t* rowp = stg + lane * 128;
          const int s7 = lane & 7;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(a0[8 * c + 0]), __uint_as_float(a0[8 * c + 1]));
            w.y = pack_bf16x2(__uint_as_float(a0[8 * c + 2]), __uint_as_float(a0[8 * c + 3]));
            w.z = pack_bf16x2(__uint_as_float(a0[8 * c + 4]), __uint_as_float(a0[8 * c + 5]));
            w.w = pack_bf16x2(__uint_as_float(a0[8 * c + 6]), __uint_as_float(a0[8 * c + 7]));
            *reinterpret_cast<uint4*>(rowp + ((c ^ s7) << 4)) = w;
            w.x = pack_bf16x2(__uint_as_float(a1[8 * c + 0]), __uint_as_float(a1[8 * c + 1]));
            w.y = pack_bf16x2(__uint_as_float(a1[8 * c + 2]), __uint_as_float(a1[8 * c + 3]));
            w.z = pack_bf16x2(__uint_as_float(a1[8 * c + 4]), __uint_as_float(a1[8 * c + 5]));
            w.w = pack_bf16x2(__uint_as_float(a1[8 * c + 6]), __uint_as_float(a1[8 * c + 7]));
            *reinterpret_cast<uint4*>(rowp + (((c + 4) ^ s7) << 4)) = w;
          }
          __syncwarp();
          __nv_bfloat16* dst = G.dqkv + (hh == 0 ? 2 * HDIM : HDIM) + h * SD;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), ch = lane & 7;
            const int kj = kt * 128 + q4 * 32 + rr;
            if (kj < n) {
              const uint4 w = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
              *reinterpret_cast<uint4*>(dst + (tok0 + kj) * G.ld_dqkv + ch * 8) = w;
            }
          }
        }
        named_bar_sync(1, 256);   // staging rows drained before the next unit's dS^T lands on them
      }
      // ---- dQ: tile hh (queries hh*128 ..), scaled; CLS query row -> atomics
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      {
        uint32_t a0[32], a1[32];
        const uint32_t t_acc = tmem_base + lane_off + 384 + hh * 64;
        tmem_ld_32x32b_x32(t_acc, a0);
        tmem_ld_32x32b_x32(t_acc + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
        const int qrow = hh * 128 + r_tile;
        if (qrow == n && G.cls_acc != nullptr) {
          float* acc = G.cls_acc + (static_cast<long long>(b) * G.H + h) * 3 * SD;
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            atomicAdd(acc + d, __uint_as_float(a0[d]) * G.scale);
            atomicAdd(acc + 32 + d, __uint_as_float(a1[d]) * G.scale);
          }
        }
        uint8_t* rowp = stg + lane * 128;
        const int s7 = lane & 7;
        const float sc = G.scale;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(a0[8 * c + 0]) * sc, __uint_as_float(a0[8 * c + 1]) * sc);
          w.y = pack_bf16x2(__uint_as_float(a0[8 * c + 2]) * sc, __uint_as_float(a0[8 * c + 3]) * sc);
          w.z = pack_bf16x2(__uint_as_float(a0[8 * c + 4]) * sc, __uint_as_float(a0[8 * c + 5]) * sc);
          w.w = pack_bf16x2(__uint_as_float(a0[8 * c + 6]) * sc, __uint_as_float(a0[8 * c + 7]) * sc);
          *reinterpret_cast<uint4*>(rowp + ((c ^ s7) << 4)) = w;
          w.x = pack_bf16x2(__uint_as_float(a1[8 * c + 0]) * sc, __uint_as_float(a1[8 * c + 1]) * sc);
          w.y = pack_bf16x2(__uint_as_float(a1[8 * c + 2]) * sc, __uint_as_float(a1[8 * c + 3]) * sc);
          w.z = pack_bf16x2(__uint_as_float(a1[8 * c + 4]) * sc, __uint_as_float(a1[8 * c + 5]) * sc);
          w.w = pack_bf16x2(__uint_as_float(a1[8 * c + 6]) * sc, __uint_as_float(a1[8 * c + 7]) * sc);
          *reinterpret_cast<uint4*>(rowp + (((c + 4) ^ s7) << 4)) = w;
        }
        __syncwarp();
        __nv_bfloat16* dst = G.dqkv + h * SD;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3), ch = lane & 7;
          const int qj = hh * 128 + q4 * 32 + rr;
          if (qj < n) {
            const uint4 w = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
            *reinterpret_cast<uint4*>(dst + (tok0 + qj) * G.ld_dqkv + ch * 8) = w;
          }
        }
      }
      named_bar_sync(1, 256);   // staging drained; also orders this group's lse/delta reads before the next group's writes
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// =============================================================================================== backward, pipelined
// Same key-major formulation as above, re-cut so that the tensor pipe, the math warps, the loads and the stores all run
// at the same time (the kernel above runs them one after the other: measured 33 k clk per group, of which 2.4 k math +
// 2.5 k MMA + handshakes per unit, 4 k delta prologue and 4.7 k exposed loads).
//   * sub-unit = (key tile kt of 128 keys) x (64 queries): S^T / dP^T are N = 64 accumulators, DOUBLE-buffered in TMEM
//     (2 x 128 columns), so the MMAs of sub-unit v+1 and the gradient MMAs of v-1 run under the math of v;
//   * dS^T goes through a ring of four [128 keys x 64 queries] shared-memory blocks; dQ is issued once per PAIR of
//     sub-units (M = 128 queries) from two adjacent blocks;
//   * 16 warps: TMA producer, MMA issuer, 2 warps that prepare lse2 / delta of the NEXT group from global memory,
//     8 math warps, 4 epilogue warps (dV, dK, dQ out of TMEM -> bf16 -> coalesced stores) that run beside the math;
//   * operands are loaded in three phases ({K0,V0}, {Q0,dO0}, {K1,V1,Q1,dO1,CLS rows}) and released tile by tile, so
//     the next group's first two phases land while the current group is still in its second key tile.
// Tensor-pipe budget (scripts/probes/mma_rate_probe.cu): every M = 128 instruction with N <= 128 costs ~68 clk (the A
// operand feed), 82 from TMEM, 88 for an MN-major A: ~11.5 k clk per group for the 160 instructions - about the time
// HBM needs for the group's 238 KB (10.3 k clk at 1/148 of 6.4 TB/s).
#ifdef OAT_SPACE_DBG
__device__ long long g_dbg[8192];
#if OAT_SPACE_DBG >= 2   // light: one timestamp per group (slot 0 only), negligible overhead
#define DBG2(it, slot) do { if ((slot) == 0 && blockIdx.x == 0 && lane == 0 && (it) < 30) g_dbg[(it) * 128] = clock64(); } while (0)
#else
#define DBG2(it, slot) do { if (blockIdx.x == 0 && lane == 0 && (it) < 8) g_dbg[(it) * 128 + (slot)] = clock64(); } while (0)
#endif
#else
#define DBG2(it, slot) do { } while (0)
#endif
#ifdef OAT_SPACE_DBG
#define DBGK(slot) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_dbg[3900 + (slot)] = clock64() - t_kernel_start; } while (0)
#else
#define DBGK(slot) do { } while (0)
#endif
constexpr int kB2Threads = 512;
constexpr int kB2RingBytes = 4 * (kTileBytes);                   // dS^T ring: 4 blocks of [128 keys x 64 queries] bf16
constexpr int kB2StageBytes = 4 * 4096;                          // epilogue staging: 32 rows x 128 B per epilogue warp
constexpr int kB2Smem = kBwdOperandBytes + kB2RingBytes + kB2StageBytes + 6 * 256 * 4 + 1024 + 512;

__global__ void __launch_bounds__(kB2Threads, 1)
attn_space_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_qkv_a, const __grid_constant__ CUtensorMap tmap_qkv_b,
                          const __grid_constant__ CUtensorMap tmap_qkv_q, const __grid_constant__ CUtensorMap tmap_qkv_p,
                          const __grid_constant__ CUtensorMap tmap_do_q, const __grid_constant__ CUtensorMap tmap_do_p,
                          const __grid_constant__ CUtensorMap tmap_o_q, const __grid_constant__ CUtensorMap tmap_o_p,
                          const __grid_constant__ CUtensorMap tmap_dqkv_a, const __grid_constant__ CUtensorMap tmap_dqkv_b,
                          const SpaceBwdGeom G, const int sched_slot) {
  unsigned int* const sched = g_sched[1][sched_slot];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + kMatBytes;
  uint8_t* Vs = smem + 2 * kMatBytes;
  uint8_t* Ds = smem + 3 * kMatBytes;                                   // dO
  uint8_t* ring = smem + kBwdOperandBytes;                              // dS^T blocks
  uint8_t* stage = ring + kB2RingBytes;                                 // epilogue staging
  float* lse2_s = reinterpret_cast<float*>(stage + kB2StageBytes);      // [3][256]
  float* del_s = lse2_s + 768;                                          // [3][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(del_s + 768);
  uint64_t* full_kv = bars;          // [2] producer -> MMA: {K,V} rows of key tile 0 / 1
  uint64_t* empty_kv = bars + 2;     // [2] MMA -> producer
  uint64_t* full_qd = bars + 4;      // [4] producer -> MMA: {Q,dO} rows of 64-query block 0..3
  uint64_t* empty_qd = bars + 8;     // [4] MMA -> producer
  uint64_t* st_full = bars + 12;     // [2] MMA -> math: S^T, dP^T of a sub-unit ready (per TMEM buffer)
  uint64_t* math_done = bars + 14;   // [2] math -> MMA: P^T, dS^T in TMEM, dS^T in the ring
  uint64_t* acc_full = bars + 16;    // MMA -> epilogue: dV, dK of a key tile complete
  uint64_t* acc_free = bars + 17;    // epilogue -> MMA
  uint64_t* dq_full = bars + 18;     // [2] MMA -> epilogue: dQ of the query half that completes after sub-unit 5 / after 7
  uint64_t* dq_free = bars + 20;     // [2] epilogue -> MMA
  uint64_t* dl_full = bars + 22;     // [3] delta warps -> math: lse2 / delta of a group ready
  uint64_t* dl_free = bars + 25;     // [3] math -> delta warps
  uint64_t* sched_full = bars + 28;  // [8] producer -> everyone: group index of an iteration published
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);
  int* sched_g = reinterpret_cast<int*>(tmem_slot + 1);               // [8] ring of group indices (-1: no more work)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HDIM = G.H * SD;
  const int n = G.n;
  const int rem = n - 128;           // token rows of the second tile (the CLS row follows them)
  pdl_launch_dependents();
#ifdef OAT_SPACE_DBG
  const long long t_kernel_start = clock64();
#endif

  for (int i = tid; i < (kBwdOperandBytes + kB2RingBytes) / 16; i += kB2Threads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 5) DBGK(0);
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv_a); tma_prefetch_desc(&tmap_qkv_b);
      tma_prefetch_desc(&tmap_qkv_q); tma_prefetch_desc(&tmap_qkv_p);
      tma_prefetch_desc(&tmap_do_q); tma_prefetch_desc(&tmap_do_p);
      tma_prefetch_desc(&tmap_o_q); tma_prefetch_desc(&tmap_o_p);
      tma_prefetch_desc(&tmap_dqkv_a); tma_prefetch_desc(&tmap_dqkv_b);
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  } else if (warp == 1 && lane == 0) {
    mbar_init(&full_kv[0], 1); mbar_init(&full_kv[1], 2);                 // the CLS rows of k, v arrive by hand
    for (int k = 0; k < 4; ++k) mbar_init(&full_qd[k], k == (n >> 6) ? 2 : 1);   // ... and so do those of q, dO
    for (int k = 0; k < 2; ++k) mbar_init(&empty_kv[k], 1);
    for (int k = 0; k < 4; ++k) mbar_init(&empty_qd[k], 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(&st_full[k], 1);
      mbar_init(&math_done[k], 8);
    }
    for (int k = 0; k < 3; ++k) {
      mbar_init(&dl_full[k], 2);
      mbar_init(&dl_free[k], 8);
    }
    mbar_init(acc_full, 1); mbar_init(acc_free, 4);
    for (int k = 0; k < 2; ++k) { mbar_init(&dq_full[k], 1); mbar_init(&dq_free[k], 4); }
    for (int k = 0; k < 8; ++k) mbar_init(&sched_full[k], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 5) DBGK(1);
  pdl_wait();
  if (warp == 5) DBGK(2);
  const uint32_t tmem_base = *tmem_slot;
  // group of iteration i (every role but the producer, which fetches them one iteration ahead of its own loads)
  auto group_of = [&](int i) -> int {
    mbar_wait(&sched_full[i & 7], (i >> 3) & 1);
    return sched_g[i & 7];
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // Operand sets in the order the group needs them (see the MMA issuer for the schedule): {K0,V0}, the two 64-query
    // blocks of the first query half, those of the other half, {K1,V1}. Q and dO are loaded and released per 64-query
    // block: the blocks the next group opens with are the ones this group stops reading 3-4 sub-units before its end.
    const int cq = n >> 6;                                   // 64-query block that holds the CLS query row
    const int part_q = rem >= 64 ? 3 : 2;                    // the block that is cut short by the CLS row ...
    const int part_rows = rem >= 64 ? rem - 64 : rem;        // ... and its token rows
    auto fetch = [&](int i) -> int {              // next group from the global counter, published in ring slot i & 3
      int g = 0;
      if (lane == 0) {
        g = static_cast<int>(atomicAdd(&sched[0], 1u));
        if (g >= G.groups) g = -1;
        sched_g[i & 7] = g;
        mbar_arrive(&sched_full[i & 7]);
      }
      return __shfl_sync(0xffffffffu, g, 0);
    };
    // Two groups ahead: the delta warps work that far in front of the math warps (their ~10 k clk per group - four
    // dependent load / reduce rounds in two warps - would otherwise sit right on the critical path), and the MMA warp
    // looks one group ahead.
    int g = fetch(0);
    int g_next = g >= 0 ? fetch(1) : -1;
    DBGK(3);
    for (int i = 0; g >= 0; ++i) {
      const int g_next2 = g_next >= 0 ? fetch(i + 2) : -1;
      if (g_next >= 0 && lane == 0) {
        // Pull the next group's q / k / v / dO / O head slices into L2 now, a whole group time before they are read.
        // The operand tiles are single-buffered (their TMA loads start only when the previous group releases them) and
        // the delta warps read O / dO with plain loads, four dependent round trips per group: from DRAM both are
        // latency-bound, and on the SMs with the longest path to memory the delta warps set the pace of the kernel
        // (measured: 398 k cycles per CTA with them, 304 k with their loads removed).
        const int hn = g_next % G.H, rn = g_next / G.H, fn = rn % G.F, bn = rn / G.F;
        const int rown = bn * G.T + 1 + fn * n;
#pragma unroll
        for (int pq = 0; pq < 4; ++pq) {
          const int rows = pq == part_q ? part_rows : (pq < part_q ? 64 : 0);
          if (rows > 0) {
            const bool pt = pq == part_q;
#pragma unroll
            for (int m = 0; m < 3; ++m) tma_prefetch_2d(pt ? &tmap_qkv_p : &tmap_qkv_q, m * HDIM + hn * SD, rown + pq * 64);
            tma_prefetch_2d(pt ? &tmap_do_p : &tmap_do_q, hn * SD, rown + pq * 64);
            if (G.delta == nullptr) tma_prefetch_2d(pt ? &tmap_o_p : &tmap_o_q, hn * SD, rown + pq * 64);   // O feeds delta only
          }
        }
      }
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      const int row0 = b * G.T + 1 + f * n;
      const uint32_t pe = (i & 1) ^ 1;
      const int flip = i & 1;
      // CLS token rows -> row n of two operands: (q, dO) with their 64-query block, (k, v) with key tile 1
      auto cls_rows = [&](int m_lo, int m_hi) {
        if (lane < 16) {
          const int m = lane < 8 ? m_lo : m_hi, c = lane & 7;
          const __nv_bfloat16* src = (m < 3)
              ? G.qkv + static_cast<long long>(b) * G.T * G.ld_qkv + m * HDIM + h * SD + c * 8
              : G.dout + static_cast<long long>(b) * G.T * G.ld_dout + h * SD + c * 8;
          const uint4 val = *reinterpret_cast<const uint4*>(src);
          *reinterpret_cast<uint4*>(smem + m * kMatBytes + n * 128 + ((c ^ (n & 7)) << 4)) = val;
        }
        fence_proxy_async_smem();
        __syncwarp();
      };
      auto load_qd = [&](int pq) {
        mbar_wait(&empty_qd[pq], pe);
        const int rows = pq == part_q ? part_rows : (pq < part_q ? 64 : 0);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_qd[pq], 2u * rows * 128u);
          if (rows > 0) {
            tma_load_2d(Qs + pq * 8192, pq == part_q ? &tmap_qkv_p : &tmap_qkv_q, &full_qd[pq], h * SD, row0 + pq * 64);
            tma_load_2d(Ds + pq * 8192, pq == part_q ? &tmap_do_p : &tmap_do_q, &full_qd[pq], h * SD, row0 + pq * 64);
          }
        }
        if (pq == cq) {
          cls_rows(0, 3);
          if (lane == 0) mbar_arrive(&full_qd[pq]);
        }
      };
      mbar_wait(&empty_kv[0], pe);
      DBG2(i, 100);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_kv[0], 2u * kTileBytes);
        tma_load_2d(Ks, &tmap_qkv_a, &full_kv[0], HDIM + h * SD, row0);
        tma_load_2d(Vs, &tmap_qkv_a, &full_kv[0], 2 * HDIM + h * SD, row0);
      }
      load_qd(flip * 2);
      load_qd(flip * 2 + 1);
      DBG2(i, 101);
      load_qd((flip ^ 1) * 2);
      load_qd((flip ^ 1) * 2 + 1);
      DBG2(i, 102);
      mbar_wait(&empty_kv[1], pe);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_kv[1], 2u * rem * 128u);
        if (rem > 0) {
          tma_load_2d(Ks + kTileBytes, &tmap_qkv_b, &full_kv[1], HDIM + h * SD, row0 + 128);
          tma_load_2d(Vs + kTileBytes, &tmap_qkv_b, &full_kv[1], 2 * HDIM + h * SD, row0 + 128);
        }
      }
      cls_rows(1, 2);
      if (lane == 0) mbar_arrive(&full_kv[1]);
      g = g_next;
      g_next = g_next2;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // One thread feeds ~160 instructions per group to the tensor pipe, which needs ~70 clk for each: the issue path
    // itself has to stay well under that. The eight sub-units of a group are unrolled, so every operand offset, TMEM
    // column and barrier parity below is a compile-time constant added to a descriptor built once.
    constexpr uint32_t idesc_sd = make_idesc_bf16(128, 64, 0u, 0u);    // S^T, dP^T: both operands K-major
    constexpr uint32_t idesc_kn = make_idesc_bf16(128, SD, 0u, 1u);    // dV, dK: A K-major (TMEM / smem), B MN-major
    constexpr uint32_t idesc_mn = make_idesc_bf16(128, SD, 1u, 1u);    // dQ: A MN-major smem, B MN-major
    const uint64_t kK = make_smem_desc_sw128(smem_u32(Ks), 0, 1024), kV = make_smem_desc_sw128(smem_u32(Vs), 0, 1024),
                   kQ = make_smem_desc_sw128(smem_u32(Qs), 0, 1024), kD = make_smem_desc_sw128(smem_u32(Ds), 0, 1024),
                   kR = make_smem_desc_sw128(smem_u32(ring), 0, 1024);
    const uint64_t mK = make_smem_desc_sw128(smem_u32(Ks), kMatBytes, 1024), mQ = make_smem_desc_sw128(smem_u32(Qs), kMatBytes, 1024),
                   mD = make_smem_desc_sw128(smem_u32(Ds), kMatBytes, 1024),
                   mR = make_smem_desc_sw128(smem_u32(ring), kTileBytes, 1024);
    // This CTA owns the SM's whole tensor memory (512 columns, 1 CTA / SM), so the allocation starts at column 0, lane 0.
    // Using the literal address keeps every tcgen05.mma operand in uniform registers: with a base loaded from shared
    // memory the compiler wraps each instruction in an elect / R2UR.BROADCAST loop (~9 instructions per MMA), and the
    // issue thread, not the tensor pipe, sets the pace of these 67-clock instructions.
    if (tmem_base != 0) __trap();
    constexpr uint32_t tmem0 = 0;
    constexpr uint32_t t_dv = tmem0 + 256, t_dk = tmem0 + 320, t_dq = tmem0 + 384;
    auto off = [](uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); };   // start-address field: bits [0, 14)

    // Schedule of a group: sub-unit v = (key tile v >> 2) x (64 queries). The four 128-query steps visit the query
    // halves in the order 0 1 1 0 in even group iterations and 1 0 0 1 in odd ones (kQH ^ flip): the operand tiles a
    // group reads first ({K0,V0} and its first query half) are then exactly the ones the previous group stopped reading
    // two steps / one step before its end, so their loads are never exposed, and the tiles released last ({K1,V1} and
    // the other query half) are needed one and two steps into the next group.
    // The issue thread's own instruction stream is the critical path of this kernel (the tensor pipe needs ~45-80 clk per
    // instruction, the thread needs ~6 clk per instruction of its own and shares its scheduler with three busy warps):
    // everything is unrolled to literal operands, and a sub-unit is ONE elected block - waits first, then up to 24
    // tcgen05.mma back to back.
    // MMAs of S^T and dP^T of sub-unit v (group iteration `it`) into TMEM buffer v & 1; the caller has waited for the operands
    auto sdp_mmas = [&](int it, int v) {
      const int kt = v >> 2, slot = (0x6 >> (v >> 1)) & 1, flip = it & 1;      // slot (dQ accumulator): 0 1 1 0
      const uint32_t qoff = static_cast<uint32_t>(((slot ^ flip) * 2 + (v & 1)) * 8192);
      const uint32_t t_s = tmem0 + (v & 1) * 128;
      const uint64_t bq = off(kQ, qoff), bd = off(kD, qoff);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(t_s, off(kK, kt * kTileBytes + k * 32), bq + k * 2, idesc_sd, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(t_s + 64, off(kV, kt * kTileBytes + k * 32), bd + k * 2, idesc_sd, k > 0 ? 1u : 0u);
      tc_commit(&st_full[v & 1]);
    };
    auto sdp_wait = [&](int it, int v) {            // operand sets sub-unit v is the first to read
      const int flip = it & 1;
      if (v == 0 || v == 4) mbar_wait(&full_kv[v >> 2], it & 1);
      if (v < 4) mbar_wait(&full_qd[((v >> 1) ^ flip) * 2 + (v & 1)], it & 1);
    };
    auto sdp_ready = [&](int it, int v) {           // the same, as a non-blocking question
      const int flip = it & 1;
      bool ok = true;
      if (v == 0 || v == 4) ok = mbar_test_wait(&full_kv[v >> 2], it & 1);
      if (v < 4) ok = ok && mbar_test_wait(&full_qd[((v >> 1) ^ flip) * 2 + (v & 1)], it & 1);
      return __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
    };
    // Schedule of a group: sub-unit v = (key tile v >> 2) x (64 queries). The four 128-query steps visit the query
    // halves in the order 0 1 1 0 in even group iterations and 1 0 0 1 in odd ones (slot ^ flip): the operand tiles a
    // group reads first ({K0,V0} and its first query half) are then the ones the previous group stopped reading
    // two steps / one step before its end, and the tiles released last ({K1,V1} and the other query half) are needed
    // one and two steps into the next group.
    // Gradient products of sub-unit v, with S^T / dP^T of sub-unit v + 2 (same TMEM buffer) slipped in right behind the
    // eight dV / dK instructions that read P^T / dS^T out of that buffer (A operands from TMEM: no shared-memory
    // fetch); the dQ products run under the math warps' next sub-unit.
    auto issue_grads = [&](int it, int v, bool more) {
      const int kt = v >> 2, slot = (0x6 >> (v >> 1)) & 1, flip = it & 1;
      const uint32_t qoff = static_cast<uint32_t>(((slot ^ flip) * 2 + (v & 1)) * 8192);
      const int it2 = v < 6 ? it : it + 1, v2 = (v + 2) & 7;       // the sub-unit whose S^T / dP^T go out with this one
      if (v == 0) DBG2(it, 0);
      // operands of a new group that have not landed yet must not hold this group's last products back
      bool sdp_now = more;
      if (more && v2 <= 4) sdp_now = sdp_ready(it2, v2);
      DBG2(it, v * 4 + 2);
      mbar_wait(&math_done[v & 1], (v >> 1) & 1);
      DBG2(it, v * 4 + 3);
      if ((v & 3) == 0) mbar_wait(acc_free, (kt & 1) ^ 1);       // the first accumulation into dV / dK of a key tile overwrites them
      // ... and so does the first one into a dQ accumulator. The accumulator follows the query half (half 0 -> columns
      // 384.., half 1 -> 448..), so the one this group finishes LAST (sub-unit 7) is the one the next group starts
      // SECOND (sub-unit 3): the epilogue warps get three sub-units to drain it instead of one.
      if (v == 1) mbar_wait(&dq_free[0], (it & 1) ^ 1);           // half `flip`: the previous group's early one
      if (v == 3) mbar_wait(&dq_free[1], (it & 1) ^ 1);           // half `1 ^ flip`: the previous group's late one
      DBG2(it, 104 + v);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t t_p = tmem0 + (v & 1) * 128;               // P^T (bf16): queries 0..31 at +0, 32..63 at +32; dS^T at +64
        const uint64_t bd = off(mD, qoff), bq = off(mQ, qoff);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {                           // k = 16 queries per step; both A operands from TMEM
          const uint32_t a_col = ks < 2 ? ks * 8 : 32 + (ks - 2) * 8;
          const uint32_t first = ((v & 3) > 0 || ks > 0) ? 1u : 0u;
          tc_mma_bf16_ts(t_dv, t_p + a_col, bd + ks * 128, idesc_kn, first);        // dV += P^T dO
          tc_mma_bf16_ts(t_dk, t_p + 64 + a_col, bq + ks * 128, idesc_kn, first);   // dK += dS^T Q
        }
        if ((v & 3) == 3) tc_commit(acc_full);     // dV, dK of this key tile are complete: the epilogue warps may start
        // the last reads of a 64-query block of Q / dO: blocks of accumulator 1's half after sub-units 4, 5, the other
        // half after 6, 7
        if (v >= 4) tc_commit(&empty_qd[((v < 6 ? 1 : 0) ^ flip) * 2 + (v & 1)]);
        if (sdp_now) sdp_mmas(it2, v2);
        if (v & 1) {                                               // dQ of the pair (v - 1, v): 128 queries
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)                           // k = 16 keys per step, A = dS^T read MN-major
            tc_mma_bf16(t_dq + (slot ^ flip) * SD, off(mR, ((v - 1) & 3) * kTileBytes + ks * 2048),
                        off(mK, (kt * 128 + ks * 16) * 128), idesc_mn, (kt > 0 || ks > 0) ? 1u : 0u);
        }
        if (v == 3) tc_commit(&empty_kv[0]);
        if (v == 5) tc_commit(&dq_full[0]);                                      // query half 1 ^ flip is complete
        if (v == 7) { tc_commit(&dq_full[1]); tc_commit(&empty_kv[1]); }
      }
      __syncwarp();
      DBG2(it, 80 + v);
      if (more && !sdp_now) {
        sdp_wait(it2, v2);
        tc_fence_after();
        if (elect_one()) sdp_mmas(it2, v2);
        __syncwarp();
      }
    };
    auto issue_sdp = [&](int it, int v) {
      sdp_wait(it, v);
      tc_fence_after();
      if (elect_one()) sdp_mmas(it, v);
      __syncwarp();
    };
    // (only THIS role is unrolled over the sub-units; the math and epilogue warps loop - with every role unrolled the
    // kernel was 130 KB of code and 16-20 % of the math warps' stall samples were instruction fetches)
    bool have = group_of(0) >= 0;
    DBGK(4);
    if (have) { issue_sdp(0, 0); issue_sdp(0, 1); }
    DBGK(5);
    for (int it = 0; have; ++it) {
      const bool more = group_of(it + 1) >= 0;
#pragma unroll
      for (int v = 0; v < 8; ++v) issue_grads(it, v, v < 6 || more);
      have = more;
    }
    DBGK(6);
  } else if (warp < 4) {
    // ------------------------------------------------------------------ lse2 / delta of the next group (global only)
    const int t64 = (warp - 2) * 32 + lane;
    const int part = t64 & 3, rsub = t64 >> 2;                    // 4 threads per query row, 16 rows per pass
    for (int i = 0;; ++i) {
      const int g = group_of(i);
      if (g < 0) break;
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      const long long tok_base = static_cast<long long>(b) * G.T;
      const long long tok0 = tok_base + 1 + f * n;
      const int db = i % 3, dph = (i / 3) & 1;       // lse2 / delta buffer of this group and its phase
      mbar_wait(&dl_free[db], dph ^ 1);
      if (warp == 2) DBG2(i, 60);
      float* l2 = lse2_s + db * 256;
      float* dl = del_s + db * 256;
      const float* lse_g = G.lse + (static_cast<long long>(b) * G.H + h) * G.T;
      if (G.delta != nullptr) {
        // delta came out of the epilogue of the GEMM that produced dO (oat_gemm_bf16 act 4): two 4-byte loads per query
        // row instead of 256 bytes of O and dO (their loads cost a fifth of this kernel's cycles)
        const float* del_g = G.delta + static_cast<long long>(h) * G.ld_delta + tok_base;
        float lv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = u * 64 + t64;
          const int rc = r < n ? r : n;
          const long long tl = (rc == n) ? 0 : 1 + f * n + rc;
          lv[u] = __ldg(lse_g + tl);
          dv[u] = __ldg(del_g + tl);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = u * 64 + t64;
          dl[r] = r <= n ? -dv[u] : 0.f;
          l2[r] = r <= n ? -lv[u] * kLog2e : 0.f;
        }
      } else
#pragma unroll 1
      for (int p0 = 0; p0 < 16; p0 += 4) {
        // 4 rows per thread in flight: all 17 loads of a batch are issued before the first use (no branches: rows past
        // the CLS query row are clamped onto it and zeroed afterwards)
        uint4 o[4][2], d[4][2];
        float lv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = (p0 + u) * 16 + rsub;
          const int rc = r < n ? r : n;
          const long long tok = (rc == n) ? tok_base : tok0 + rc;
          const __nv_bfloat16* op = G.out + tok * G.ld_out + h * SD + part * 16;
          const __nv_bfloat16* dp = G.dout + tok * G.ld_dout + h * SD + part * 16;
          o[u][0] = __ldg(reinterpret_cast<const uint4*>(op));
          o[u][1] = __ldg(reinterpret_cast<const uint4*>(op + 8));
          d[u][0] = __ldg(reinterpret_cast<const uint4*>(dp));
          d[u][1] = __ldg(reinterpret_cast<const uint4*>(dp + 8));
          lv[u] = __ldg(lse_g + (tok - tok_base));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = (p0 + u) * 16 + rsub;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ow[4] = {o[u][k].x, o[u][k].y, o[u][k].z, o[u][k].w};
            const uint32_t dw[4] = {d[u][k].x, d[u][k].y, d[u][k].z, d[u][k].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(ow[e]), y = unpack_bf16x2(dw[e]);
              acc = fmaf(x.x, y.x, acc);
              acc = fmaf(x.y, y.y, acc);
            }
          }
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          if (part == 0) {
            dl[r] = r <= n ? -acc : 0.f;                    // negated: the math warps add them
            l2[r] = r <= n ? -lv[u] * kLog2e : 0.f;
          }
        }
      }
      __syncwarp();
      if (warp == 2) DBG2(i, 61);
      if (lane == 0) mbar_arrive(&dl_full[db]);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ math warps
    const int q4 = warp & 3;                  // TMEM lane quarter
    const int hh = (warp - 4) >> 2;           // which 32-query half of a sub-unit
    const int tr = lane >> 2, tq = lane & 3;  // fragment coordinates: key rows tr, tr + 8 (+16, +24), query pair tq
    const uint32_t la0 = static_cast<uint32_t>(q4 * 32) << 16, la1 = static_cast<uint32_t>(q4 * 32 + 16) << 16;
    for (int i = 0;; ++i) {
      const int g = group_of(i);
      if (g < 0) break;
      const int f = (g / G.H) % G.F;
      const int flip = i & 1;
      const int db = i % 3, dph = (i / 3) & 1;
      const float* l2 = lse2_s + db * 256;
      const float* dl = del_s + db * 256;
      mbar_wait(&dl_full[db], dph);
      // This warp's block of a sub-unit: 32 keys (TMEM lanes q4*32 ..) x 32 queries (columns hh*32 ..), read in the
      // mma-fragment layout (16x256b): a thread holds 4 keys x 8 queries, so it needs lse / delta of 8 queries only
      // (eight 8-byte shared-memory loads per sub-unit; with one key row per thread it would be 64 values = 16 broadcast
      // LDS.128, 4 k wavefronts per group - a quarter of the kernel's shared-memory time).
      // A tcgen05.ld takes ~220 clk to come back, so the two 16-key halves of a block are software-pipelined: the
      // loads of one half are in flight while the other half is computed (tcgen05.wait::ld has no groups: a load is
      // issued right AFTER the wait that precedes the compute step it overlaps).
      uint32_t S0[16], D0[16], S1[16], D1[16];
      auto ld_half = [&](int v, int kh, uint32_t (&Sx)[16], uint32_t (&Dx)[16]) {
        const uint32_t t_s = tmem_base + (v & 1) * 128 + hh * 32 + (kh ? la1 : la0);   // S^T; dP^T at +64
        tmem_ld_16x256b_x4(t_s, Sx);
        tmem_ld_16x256b_x4(t_s + 64, Dx);
      };
      // one half: P^T, dS^T of 2 key rows x 8 queries per thread -> TMEM (A operands of dV, dK) and the ring (dQ)
      auto compute_half = [&](int v, int kh, int qq, uint32_t (&Sx)[16], uint32_t (&Dx)[16], const uint64_t (&nls)[4],
                              const uint64_t (&ndl)[4]) {
        const int kt = v >> 2;
        // the (CLS query, CLS key) cell is counted by frame 0 only: a score of -inf makes P and dS exactly 0 there
        if (f != 0 && (n >> 7) == kt && ((n & 127) >> 4) == q4 * 2 + kh && (n >> 5) == qq * 2 + hh) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int key = kt * 128 + q4 * 32 + kh * 16 + tr + 8 * ((k >> 1) & 1);
            const int qry = qq * 64 + hh * 32 + 2 * tq + 8 * (k >> 2) + (k & 1);
            if (key == n && qry == n) Sx[k] = 0xff800000u;
          }
        }
        const uint64_t l2e = f2_pack(kLog2e, kLog2e);
        uint8_t* blk = ring + (v & 3) * kTileBytes + (q4 * 32 + kh * 16 + tr) * 128 + 4 * tq;
        uint32_t pp[8], dd[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int rs = 0; rs < 2; ++rs) {
            const int k = 4 * j + 2 * rs;
            float x0, x1, d0, d1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(Sx[k]), __uint_as_float(Sx[k + 1])), l2e, nls[j]), x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            const uint64_t t = f2_add(f2_pack(__uint_as_float(Dx[k]), __uint_as_float(Dx[k + 1])), ndl[j]);
            f2_unpack(f2_mul(f2_pack(p0, p1), t), d0, d1);
            pp[2 * j + rs] = pack_bf16x2(p0, p1);
            dd[2 * j + rs] = pack_bf16x2(d0, d1);
            // dS^T row (key) in the ring block, 16-byte chunk hh*4 + j, word tq: conflict-free (8 rows x 4 words)
            *reinterpret_cast<uint32_t*>(blk + rs * 1024 + (((hh * 4 + j) ^ tr) << 4)) = dd[2 * j + rs];
          }
        }
        const uint32_t t_s = tmem_base + (v & 1) * 128 + hh * 32 + (kh ? la1 : la0);
        tmem_st_16x128b_x4(t_s, pp);                                      // P^T  (A operand of dV)
        tmem_st_16x128b_x4(t_s + 64, dd);                                 // dS^T (A operand of dK)
      };
      mbar_wait(&st_full[0], 0);
      tc_fence_after();
      ld_half(0, 0, S0, D0);
#pragma unroll 1
      for (int v = 0; v < 8; ++v) {
        const int qq = (((0x6 >> (v >> 1)) & 1) ^ flip) * 2 + (v & 1);    // 64-query block this sub-unit covers
        const int qb = qq * 64 + hh * 32 + 2 * tq;                        // first of this thread's queries
        uint64_t nls[4], ndl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          nls[j] = *reinterpret_cast<const uint64_t*>(l2 + qb + 8 * j);    // (-lse2[q], -lse2[q + 1])
          ndl[j] = *reinterpret_cast<const uint64_t*>(dl + qb + 8 * j);    // (-delta[q], -delta[q + 1])
        }
        tmem_ld_wait();
        ld_half(v, 1, S1, D1);
        compute_half(v, 0, qq, S0, D0, nls, ndl);
        tmem_ld_wait();
        // first half of the next sub-unit, if the tensor pipe has already delivered it (it usually has)
        bool early = false;
        if (v < 7) early = __all_sync(0xffffffffu, mbar_test_wait(&st_full[(v + 1) & 1], ((v + 1) >> 1) & 1));
        if (early) {
          tc_fence_after();
          ld_half(v + 1, 0, S0, D0);
        }
        compute_half(v, 1, qq, S1, D1, nls, ndl);
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (warp == 4) DBG2(i, 34 + v * 3);
        if (lane == 0) mbar_arrive(&math_done[v & 1]);
        if (v < 7 && !early) {
          mbar_wait(&st_full[(v + 1) & 1], ((v + 1) >> 1) & 1);
          tc_fence_after();
          ld_half(v + 1, 0, S0, D0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dl_free[db]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: dV, dK per key tile, dQ per group
    const int q4 = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q4 * 32) << 16;
    uint8_t* stg = stage + q4 * 4096;
    const int s7 = lane & 7;
    // one 64-column accumulator of this warp's 32 rows: TMEM -> registers (bf16 pairs), fp32 row kept for the CLS atomics
    // (parking that row in shared memory to add it after the hand-back was tried: the dependent LDS -> RED chain of the
    // one thread that owns the row stalls its whole warp far longer than 64 back-to-back REDs from registers)
    auto drain = [&](uint32_t taddr, float sc, uint32_t (&pk)[32], float* cls_dst) {
      uint32_t a[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        tmem_ld_32x32b_x32(taddr + hf * 32, a);
        tmem_ld_wait();
        if (cls_dst != nullptr) {
#pragma unroll
          for (int d = 0; d < 32; ++d) atomicAdd(cls_dst + hf * 32 + d, __uint_as_float(a[d]) * sc);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e)
          pk[hf * 16 + e] = pack_bf16x2(__uint_as_float(a[2 * e]) * sc, __uint_as_float(a[2 * e + 1]) * sc);
      }
    };
    // 32 rows x 128 B into the warp's staging rows (128B swizzle), then ONE bulk-tensor store of the rows that are token
    // rows of this group (whole 32-row box, or the n % 32 row box for the tile that the CLS row cuts)
    auto store_rows = [&](const uint32_t (&pk)[32], int col, int row_global, int row_first) {
      if (lane == 0) bulk_wait_read<0>();        // the previous store has finished reading the staging rows
      __syncwarp();
      uint8_t* rowp = stg + lane * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(rowp + ((c ^ s7) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (row_first + 32 <= n) tma_store_2d(&tmap_dqkv_a, stg, col, row_global);
        else if (row_first < n) tma_store_2d(&tmap_dqkv_b, stg, col, row_global);
        bulk_commit();
      }
    };
    for (int i = 0;; ++i) {
      const int g = group_of(i);
      if (g < 0) break;
      const int h = g % G.H, rest = g / G.H, f = rest % G.F, b = rest / G.F;
      const int row0 = b * G.T + 1 + f * n;
      float* cls = G.cls_acc != nullptr ? G.cls_acc + (static_cast<long long>(b) * G.H + h) * 3 * SD : nullptr;
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        const int row_first = kt * 128 + q4 * 32;
        const bool is_cls = (row_first + lane == n) && cls != nullptr;
        uint32_t pv[32], pk[32];
        if (kt == 1) {
          // dQ accumulator 1 (query half 1 ^ flip) has been complete since sub-unit 5: it goes out while the second key
          // tile is finished
          const int qrow = ((i & 1) ^ 1) * 128 + q4 * 32;
          mbar_wait(&dq_full[0], i & 1);
          tc_fence_after();
          drain(tmem_base + lane_off + 384 + ((i & 1) ^ 1) * 64, G.scale, pv, (qrow + lane == n && cls != nullptr) ? cls : nullptr);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dq_free[0]);
          store_rows(pv, h * SD, row0 + qrow, qrow);
        }
        mbar_wait(acc_full, kt & 1);
        if (warp == 12) DBG2(i, 64 + kt * 3);
        tc_fence_after();
        drain(tmem_base + lane_off + 256, 1.0f, pv, is_cls ? cls + 2 * SD : nullptr);
        drain(tmem_base + lane_off + 320, 1.0f, pk, is_cls ? cls + SD : nullptr);
        tc_fence_before();
        __syncwarp();
        if (warp == 12) DBG2(i, 65 + kt * 3);
        if (lane == 0) mbar_arrive(acc_free);      // the accumulators are handed back BEFORE the stores: the MMA warp needs
        store_rows(pv, 2 * HDIM + h * SD, row0 + row_first, row_first);   // them again one sub-unit into the next key tile
        if (kt == 1) {
          // dQ accumulator 0 (query half `flip`) completes with the group's last instruction; read it (and hand it back)
          // before dK goes out
          const int qrow = (i & 1) * 128 + q4 * 32;
          mbar_wait(&dq_full[1], i & 1);
          if (warp == 12) DBG2(i, 70);
          tc_fence_after();
          drain(tmem_base + lane_off + 384 + (i & 1) * 64, G.scale, pv, (qrow + lane == n && cls != nullptr) ? cls : nullptr);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dq_free[1]);
          store_rows(pk, HDIM + h * SD, row0 + row_first, row_first);
          store_rows(pv, h * SD, row0 + qrow, qrow);
          if (warp == 12) DBG2(i, 71);
        } else {
          store_rows(pk, HDIM + h * SD, row0 + row_first, row_first);
        }
        if (warp == 12) DBG2(i, 66 + kt * 3);
      }
    }
    if (warp == 12) DBGK(7);
    if (lane == 0) bulk_wait<0>();               // all stores complete before the CTA exits
    if (warp == 12) DBGK(8);
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {   // last CTA out: every CTA has drawn its end-of-work ticket
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
#ifdef OAT_SPACE_DBG
  if (tid == 0) g_dbg[4096 + blockIdx.x] = clock64() - t_kernel_start;
#endif
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool space_tc_fwd_supported(const oat_attn_args* a) {
  return a->mode == 0 && a->key_mask == nullptr && a->cls_acc != nullptr && a->n >= 128 && a->n + 1 <= 256 &&
         a->ld_qkv % 8 == 0 && a->ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a->out) & 15) == 0;
}

// cls_acc must hold B*H*F*66 floats (forward workspace for the CLS-query partials)
int launch_space_tc_fwd(const oat_attn_args* a, cudaStream_t s) {
  SpaceGeom G;
  G.B = a->B; G.T = a->T; G.H = a->H; G.F = a->F; G.n = a->n;
  G.nk = a->n + 1;
  G.nkp = (G.nk + 15) & ~15;
  G.groups = a->B * a->F * a->H;
  G.ld_qkv = a->ld_qkv; G.ld_out = a->ld_out;
  G.qkv = reinterpret_cast<const __nv_bfloat16*>(a->qkv);
  G.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  G.lse = a->lse;
  G.cls_part = a->cls_acc;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, a->qkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_qkv, a->n);
  if (rc != OAT_OK) return rc;
  switch (G.nkp >> 4) {
    case 9: rc = launch_space_tc<9>(tm, G, s); break;
    case 10: rc = launch_space_tc<10>(tm, G, s); break;
    case 11: rc = launch_space_tc<11>(tm, G, s); break;
    case 12: rc = launch_space_tc<12>(tm, G, s); break;
    case 13: rc = launch_space_tc<13>(tm, G, s); break;
    case 14: rc = launch_space_tc<14>(tm, G, s); break;
    case 15: rc = launch_space_tc<15>(tm, G, s); break;
    case 16: rc = launch_space_tc<16>(tm, G, s); break;
    default: return set_error(OAT_ERR_ARG, "attn_space_tc: unsupported key count %d", G.nk);
  }
  if (rc != OAT_OK) return rc;
  attn_cls_combine_kernel<<<a->B * a->H, SD, 0, s>>>(G);
  return check_launch("attn_cls_combine_kernel");
}

bool space_tc_bwd_supported(const oat_attn_args* a) {
  return a->mode == 0 && a->key_mask == nullptr && a->cls_acc != nullptr && a->n >= 128 && a->n + 1 <= 256 &&
         a->ld_qkv % 8 == 0 && a->ld_out % 8 == 0 && a->ld_dout % 8 == 0 && a->ld_dqkv % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a->dout) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->dqkv) & 15) == 0;
}

// cls_acc ([B*H][3][64] fp32) must be zeroed by the caller; the caller also runs the finalize kernel afterwards
int launch_space_tc_bwd(const oat_attn_args* a, cudaStream_t s) {
  SpaceBwdGeom G;
  G.B = a->B; G.T = a->T; G.H = a->H; G.F = a->F; G.n = a->n;
  G.groups = a->B * a->F * a->H;
  G.ld_qkv = a->ld_qkv; G.ld_out = a->ld_out; G.ld_dout = a->ld_dout; G.ld_dqkv = a->ld_dqkv;
  G.qkv = reinterpret_cast<const __nv_bfloat16*>(a->qkv);
  G.out = reinterpret_cast<const __nv_bfloat16*>(a->out);
  G.dout = reinterpret_cast<const __nv_bfloat16*>(a->dout);
  G.lse = a->lse;
  G.dqkv = reinterpret_cast<__nv_bfloat16*>(a->dqkv);
  G.cls_acc = a->cls_acc;
  G.scale = a->scale;
  G.delta = a->delta; G.ld_delta = a->ld_delta;
  const int sms = num_sms();
  const int grid = G.groups < sms ? G.groups : sms;
  static const bool legacy = getenv("OAT_SPACE_BWD_V1") != nullptr;
  if (legacy) {
    CUtensorMap tq, td;
    int rc = make_tmap_bf16_2d(&tq, a->qkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_qkv, a->n);
    if (rc != OAT_OK) return rc;
    rc = make_tmap_bf16_2d(&td, a->dout, 1ull * a->H * SD, 1ull * a->B * a->T, a->ld_dout, a->n);
    if (rc != OAT_OK) return rc;
    static bool done = false;
    if (!done) {
      cudaError_t e = cudaFuncSetAttribute(attn_space_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
      if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_space_tc_bwd smem attr: %s", cudaGetErrorString(e));
      done = true;
    }
    cudaError_t e = launch_pdl(attn_space_tc_bwd_kernel, dim3(grid), dim3(kSpThreads), kBwdSmem, s, tq, td, G);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_space_tc_bwd_kernel launch: %s", cudaGetErrorString(e));
    return check_launch("attn_space_tc_bwd_kernel");
  }
  // pipelined kernel: token rows [0, 128) and [128, n) of a group are separate boxes (tiles are released one by one)
  CUtensorMap tqa, tqb, tqq, tqp, tdq, tdp, toq, top, tsa, tsb;
  const int rem_rows = a->n - 128;
  const int rem = rem_rows > 0 ? rem_rows : 1;
  const int part = (rem_rows >= 64 ? rem_rows - 64 : rem_rows) > 0 ? (rem_rows >= 64 ? rem_rows - 64 : rem_rows) : 1;
  const int tail = (a->n & 31) ? (a->n & 31) : 32;
  int rc = make_tmap_bf16_2d(&tqa, a->qkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_qkv, 128);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tqb, a->qkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_qkv, rem);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tqq, a->qkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_qkv, 64);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tqp, a->qkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_qkv, part);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tdq, a->dout, 1ull * a->H * SD, 1ull * a->B * a->T, a->ld_dout, 64);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tdp, a->dout, 1ull * a->H * SD, 1ull * a->B * a->T, a->ld_dout, part);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&toq, a->out, 1ull * a->H * SD, 1ull * a->B * a->T, a->ld_out, 64);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&top, a->out, 1ull * a->H * SD, 1ull * a->B * a->T, a->ld_out, part);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tsa, a->dqkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_dqkv, 32);
  if (rc == OAT_OK) rc = make_tmap_bf16_2d(&tsb, a->dqkv, 3ull * a->H * SD, 1ull * a->B * a->T, a->ld_dqkv, tail);
  if (rc != OAT_OK) return rc;
  static bool done2 = false;
  if (!done2) {
    cudaError_t e = cudaFuncSetAttribute(attn_space_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB2Smem);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_space_tc_bwd2 smem attr: %s", cudaGetErrorString(e));
    done2 = true;
  }
  cudaError_t e = launch_pdl(attn_space_tc_bwd2_kernel, dim3(grid), dim3(kB2Threads), kB2Smem, s, tqa, tqb, tqq, tqp, tdq, tdp, toq, top, tsa, tsb, G,
                              static_cast<int>(next_sched_slot(1)));
  if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_space_tc_bwd2_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("attn_space_tc_bwd2_kernel");
}

}  // namespace oat

#ifdef OAT_SPACE_DBG
extern "C" int oat_debug_timeline(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, oat::g_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif
