// Time attention (sequence length F + 1 <= 17). Reference: VarAttention.forward with the '(b n) f d' grouping,
// OATrans/model/video_transformer.py:112-122: token (f, i) attends to [CLS] + tokens (f', i) of every frame f'.
//
// A group (b, slot i, head h) has F queries and F + 1 keys of 64 dims: 2-8 FLOP per byte, an HBM-bound gather. The
// kernels touch every 128-byte head slice exactly once and keep the instruction count per token row low:
//   * CTA = 4 warps = 128 token rows of one (batch, head): row = group_in_warp * Fp + frame (Fp = F rounded up to a
//     power of two), i.e. 128 / Fp consecutive slots. Every warp stages the head slices of its 32 rows into shared
//     memory with cp.async, 8 lanes per 128-byte slice (full lines on the global side, all loads of the CTA in flight
//     at once), into 128B-pitch rows with an XOR-swizzled 16-byte unit order (conflict-free ldmatrix and row stores).
//   * The arithmetic runs on bf16 m16n8k16 MMAs (fp32 accumulate): a warp's 32 rows are two independent 16-row tiles
//     (groups never straddle a tile since Fp | 16); the 16 x 16 score block of a tile is computed densely and masked
//     to its block diagonal, so one code path serves every F <= 16. The CLS key is a 17th key column and the CLS
//     query a 17th query row, both expressed as MMAs against 8-row matrices whose row 0 holds the CLS vector.
//   * Results go back through the (dead) staged rows so that the global stores are full 128-byte lines too.
// A first SIMT version (one thread per token row, fp32 FMAs) needed ~3 k / ~6 k instructions per row (forward /
// backward) and ran 3x / 5x off the HBM roofline; the MMA formulation needs ~10x fewer.
#include <stdlib.h>

#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int TD = 64;            // head dim
constexpr int kTimeWarps = 4;     // 128 threads = 128 token rows staged per CTA

struct TimeGeom {
  int B, T, H, F, n, Fp, lgFp, gpc, chunks;   // Fp = 1 << lgFp; gpc: groups per CTA, chunks: CTAs per (b, h)
  long long ld_qkv, ld_out, ld_dout, ld_dqkv;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  float* lse;
  const __nv_bfloat16* dout;
  __nv_bfloat16* dqkv;
  float scale;
  float* cls_acc;                       // [B*H][3][64]: dq_cls (unscaled), dk_cls, dv_cls
  float* cls_part;                      // forward: [B*H][chunks*warps][2+64] partials of the CLS query (or null)
  const float* delta;                   // backward, optional: [H][ld_delta] rowsum(dO * O) from the dO-producing GEMM
  long long ld_delta;
};

// ------------------------------------------------------------------------------------------------ shared-memory rows
// A staged row is one head slice (64 bf16 = 128 B = 8 chunks of 16 B). Rows are packed at a 128-byte pitch and chunk c
// of row r lives at chunk position (c ^ (r & 7)): the 8 lanes that stage one row write one full 128-byte line (global
// side: one coalesced line per 8 lanes), a thread that reads ITS OWN row hits 8 distinct bank groups across a
// quarter-warp, and ldmatrix reads of 8 consecutive rows are conflict-free.
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return static_cast<uint32_t>(r) * 128u + ((static_cast<uint32_t>(c ^ (r & 7))) << 4); }

__device__ __forceinline__ void cp_async16_t(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_t() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Token of local row lr (0..31) of warp `warp` in CTA chunk `chunk`; -1 when the row is padding.
//   lr = group_in_warp * Fp + frame,  slot = chunk * gpc + warp * (32 / Fp) + group_in_warp,  token = 1 + frame * n + slot
__device__ __forceinline__ int row_token(const TimeGeom& G, int chunk, int warp, int lr) {
  const int gl = lr >> G.lgFp, i = lr & (G.Fp - 1);
  const int pos = chunk * G.gpc + ((warp * 32) >> G.lgFp) + gl;
  return (i < G.F && pos < G.n) ? 1 + i * G.n + pos : -1;
}

// Coalesced per-warp staging: in pass `it` lanes 8q..8q+7 move the 8 chunks of local row it*4 + q.
// tok[it] caches the token of the row this lane helps with in pass `it` (used again by the coalesced stores).
__device__ __forceinline__ void stage_rows_warp(uint8_t* arr, const __nv_bfloat16* src_col, long long ld, const int (&tok)[8],
                                                int warp, int lane) {
  const int c = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = warp * 32 + it * 4 + (lane >> 3);
    uint8_t* dst = arr + sw_off(r, c);
    if (tok[it] >= 0) cp_async16_t(smem_u32(dst), src_col + static_cast<long long>(tok[it]) * ld + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void store_rows_warp(const uint8_t* arr, __nv_bfloat16* dst_col, long long ld, const int (&tok)[8],
                                                int warp, int lane) {
  const int c = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = warp * 32 + it * 4 + (lane >> 3);
    if (tok[it] >= 0)
      *reinterpret_cast<uint4*>(dst_col + static_cast<long long>(tok[it]) * ld + c * 8) = *reinterpret_cast<const uint4*>(arr + sw_off(r, c));
  }
}

// several arrays that share the row pitch (q | k | v slices of one qkv row, dq | dk | dv of one dqkv row): one 64-bit row
// offset per pass instead of one per array and pass
template <int NA>
__device__ __forceinline__ void stage_rows_warp_n(uint8_t* const (&arr)[NA], const __nv_bfloat16* const (&src_col)[NA],
                                                  long long ld, const int (&tok)[8], int warp, int lane) {
  const int c = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = warp * 32 + it * 4 + (lane >> 3);
    const uint32_t off = sw_off(r, c);
    if (tok[it] >= 0) {
      const long long roff = static_cast<long long>(tok[it]) * ld + c * 8;
#pragma unroll
      for (int a = 0; a < NA; ++a) cp_async16_t(smem_u32(arr[a] + off), src_col[a] + roff);
    } else {
#pragma unroll
      for (int a = 0; a < NA; ++a) *reinterpret_cast<uint4*>(arr[a] + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
template <int NA>
__device__ __forceinline__ void store_rows_warp_n(const uint8_t* const (&arr)[NA], __nv_bfloat16* const (&dst_col)[NA],
                                                  long long ld, const int (&tok)[8], int warp, int lane) {
  const int c = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = warp * 32 + it * 4 + (lane >> 3);
    if (tok[it] >= 0) {
      const uint32_t off = sw_off(r, c);
      const long long roff = static_cast<long long>(tok[it]) * ld + c * 8;
#pragma unroll
      for (int a = 0; a < NA; ++a) *reinterpret_cast<uint4*>(dst_col[a] + roff) = *reinterpret_cast<const uint4*>(arr[a] + off);
    }
  }
}

constexpr int kRows = kTimeWarps * 32;
constexpr int kArr = kRows * 128;            // bytes of one staged array (Q, K, V or dO rows of the CTA)
constexpr int kClsPartT = 2 + TD;            // per-tile partial of the CLS query: running max, sum, 64 weighted-V sums
constexpr int kClsParts = 2 * kTimeWarps;    // partials per CTA (two 16-row tiles per warp)

// ------------------------------------------------------------------------------------------------ backward (tensor cores)
// The SIMT formulation above issues ~6 k instructions per token row in the backward; here the same arithmetic runs on
// bf16 m16n8k16 MMAs. A warp owns 32 staged token rows = two independent 16-row tiles (groups never straddle a tile:
// Fp | 16). Per tile the (16 x 16) score block  S = Q K^T  is computed densely and masked to its block diagonal
// (same group, valid rows), so one code path serves F <= 16. The CLS key is a 17th key column, the CLS query a 17th
// query row; both are expressed as MMAs against 8-row matrices whose row 0 holds the CLS vector (rows 1-7 zero):
//   query side   S, dP = dO V^T, s_i0 = q_i.k_cls, dp_i0 = dO_i.v_cls  ->  P = exp(S - lse_i), dS = P (dP - delta_i)
//                dQ  = dS K  + ds_i0 k_cls
//   key side     P^T, dS^T by movmatrix;  s_cj = k_j.q_cls, dp_cj = v_j.dO_cls -> P_cj, dS_cj
//                dK  = dS^T Q + dS_cj q_cls          dV = P^T dO + P_cj dO_cls
//   CLS rows     dQ_cls += sum_j dS_cj k_j,  dK_cls += sum_i ds_i0 q_i,  dV_cls += sum_i p_i0 dO_i  (row-vector x matrix
//                MMAs whose A operand has a single non-zero row) -> per-warp shared accumulators -> one global atomic per CTA.
// One ldmatrix.x4 of a staged (16 rows x 16 dims) block is both the A fragment of that block and the B fragments of
// its two 8-row halves, so Q, K, V, dO are read from shared memory once for all score-type products.
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// transpose an 8x8 b16 matrix held one 32-bit register per lane (row = lane / 4, columns 2 * (lane % 4), +1)
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(y) : "r"(x));
  return y;
}

// non-transposed fragments of the (16 rows x 64 dims) tile at staged row R0: f[ks] = {rows 0-7 | dims lo, rows 8-15 | lo,
// rows 0-7 | hi, rows 8-15 | hi} of k-step ks. As A operand: f[ks]; as B operand of the 8-row half h: (f[ks][h], f[ks][2 + h]).
__device__ __forceinline__ void load_tile_frags(uint32_t arr, int R0, int lane, uint32_t (&f)[4][4]) {
  const int row = R0 + (lane & 7) + ((lane & 8) ? 8 : 0);
  const uint32_t rowaddr = arr + static_cast<uint32_t>(row) * 128u;
  const uint32_t x = static_cast<uint32_t>(row & 7);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm4(rowaddr + (((2u * ks + (lane >> 4)) ^ x) << 4), f[ks]);
}
// B fragments of the 8-row CLS matrix `mat` (row 0 = vector) as an [n = row][k = dim] operand: b[ks][0..1]
__device__ __forceinline__ void load_cls_frags(uint32_t mat, int lane, uint32_t (&b)[4][2]) {
  const uint32_t rowaddr = mat + static_cast<uint32_t>(lane & 7) * 128u;
  const uint32_t x = static_cast<uint32_t>(lane & 7);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t r[4];
    ldsm4(rowaddr + (((4u * h + (lane >> 3)) ^ x) << 4), r);
    b[2 * h][0] = r[0]; b[2 * h][1] = r[1]; b[2 * h + 1][0] = r[2]; b[2 * h + 1][1] = r[3];
  }
}
// acc[j] (+)= A . X[k = tile row][n = dims 8j..8j+7]  and  acc2[j] (+)= A2 . X   for the 16-row tile at R0 (transposed loads)
__device__ __forceinline__ void mma_rows_t(uint32_t arr, int R0, int lane, const uint32_t (&A)[4], float (&acc)[8][4],
                                           const uint32_t (&A2)[4], float (&acc2)[8][4]) {
  const int row = R0 + (lane & 7) + ((lane & 8) ? 8 : 0);
  const uint32_t rowaddr = arr + static_cast<uint32_t>(row) * 128u;
  const uint32_t x = static_cast<uint32_t>(row & 7);
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    uint32_t r[4];
    ldsm4_t(rowaddr + (((static_cast<uint32_t>(j) + (lane >> 4)) ^ x) << 4), r);
    mma_bf16(acc[j], A, r[0], r[1]);
    mma_bf16(acc[j + 1], A, r[2], r[3]);
    mma_bf16(acc2[j], A2, r[0], r[1]);
    mma_bf16(acc2[j + 1], A2, r[2], r[3]);
  }
}
// acc[j] += E . C[k = matrix row][n = dims]: rank-1 update with the CLS vector in row 0 of the 8-row matrix (E: column 0)
__device__ __forceinline__ void mma_cls_t(uint32_t mat, int lane, const uint32_t (&E)[4], float (&acc)[8][4]) {
  const uint32_t rowaddr = mat + static_cast<uint32_t>(lane & 7) * 128u;
  const uint32_t x = static_cast<uint32_t>(lane & 7);
#pragma unroll
  for (int j = 0; j < 8; j += 4) {
    uint32_t r[4];
    ldsm4_t(rowaddr + (((static_cast<uint32_t>(j) + (lane >> 3)) ^ x) << 4), r);
#pragma unroll
    for (int i = 0; i < 4; ++i) mma_bf16(acc[j + i], E, r[i], 0u);
  }
}
// rows g / g + 8 of the accumulator tile -> bf16 -> staged rows (conflict-free 32-bit stores)
template <bool SCALED>
__device__ __forceinline__ void store_acc_rows(uint8_t* arr, int R0, int lane, const float (&acc)[8][4], float mul = 1.f) {
  const int g = lane >> 2, t = lane & 3;
  const int ra = R0 + g, rb = ra + 8;
  uint8_t* pa = arr + static_cast<uint32_t>(ra) * 128u + 4 * t;
  uint8_t* pb = arr + static_cast<uint32_t>(rb) * 128u + 4 * t;
  const uint32_t x = static_cast<uint32_t>(ra & 7);        // (ra & 7) == (rb & 7)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t o = (static_cast<uint32_t>(j) ^ x) << 4;
    if (SCALED) {
      *reinterpret_cast<uint32_t*>(pa + o) = pack_bf16x2(acc[j][0] * mul, acc[j][1] * mul);
      *reinterpret_cast<uint32_t*>(pb + o) = pack_bf16x2(acc[j][2] * mul, acc[j][3] * mul);
    } else {
      *reinterpret_cast<uint32_t*>(pa + o) = pack_bf16x2(acc[j][0], acc[j][1]);
      *reinterpret_cast<uint32_t*>(pb + o) = pack_bf16x2(acc[j][2], acc[j][3]);
    }
  }
}
// row 0 of the accumulator tile (lanes 0-3) -> this warp's private shared accumulators. Lane t owns dims 8j + 2t, +1 of
// its warp's slot for the whole kernel, so a plain read-add-write does (shared fp32 atomicAdd is a CAS loop: 96 of them
// per warp were a third of the kernel's stall samples).
__device__ __forceinline__ void add_row0(float* dst, int lane, const float (&acc)[8][4]) {
  if (lane < 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2* p = reinterpret_cast<float2*>(dst + 8 * j + 2 * lane);
      float2 v = *p;
      v.x += acc[j][0];
      v.y += acc[j][1];
      *p = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward (tensor cores)
// Same tiling as the backward below: a warp owns two independent 16-row tiles; S = Q K^T is computed densely per
// tile and masked to the group diagonal; the CLS key is a 17th column (MMA against the 8-row k_cls matrix), the CLS
// value a rank-1 MMA update. CLS query (video_transformer.py:108-110: attends to every token): every tile also scores
// its 16 keys against q_cls and reduces (max, sum, sum p.v) to one partial in `cls_part`, merged by
// attn_time_cls_combine_kernel; without a workspace the caller runs the separate all-keys pass instead.
template <int MINB>
__global__ void __launch_bounds__(kRows, MINB) attn_time_fwd_kernel(const TimeGeom G) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm_time_raw[];
  uint8_t* Qs = sm_time_raw;
  uint8_t* Ks = Qs + kArr;
  uint8_t* Vs = Ks + kArr;
  uint8_t* Cm = Vs + kArr;                                     // 3 CLS matrices [8][128 B]: q, k, v
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // head-fastest block order: the H CTAs that read the 128-byte head slices of the SAME token rows (one 4.6 KB qkv row
  // holds all heads) are neighbours in launch order, so they hit the same DRAM pages at about the same time
  const int h = blockIdx.x % G.H, rest_ = blockIdx.x / G.H;
  const int chunk = rest_ % G.chunks, b = rest_ / G.chunks;
  const int bh = b * G.H + h;
  const int HD3 = G.H * TD;
  const long long row0 = static_cast<long long>(b) * G.T;
  const __nv_bfloat16* base = G.qkv + row0 * G.ld_qkv + h * TD;
  int tok[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) tok[it] = row_token(G, chunk, warp, it * 4 + (lane >> 3));
  {
    uint8_t* const arrs[3] = {Ks, Qs, Vs};
    const __nv_bfloat16* const srcs[3] = {base + HD3, base, base + 2 * HD3};
    stage_rows_warp_n<3>(arrs, srcs, G.ld_qkv, tok, warp, lane);
  }
  if (threadIdx.x < 96) {  // CLS matrices: row 0 <- vector, rows 1-7 <- 0
    const int m = threadIdx.x >> 5, rr = (threadIdx.x >> 3) & 3, c = threadIdx.x & 7;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 4 + rr;
      uint8_t* dst = Cm + m * 1024 + r * 128 + ((c ^ r) << 4);
      if (r == 0) cp_async16_t(smem_u32(dst), base + m * HD3 + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const int my_tok = row_token(G, chunk, warp, lane);
  const uint32_t vm = __ballot_sync(0xffffffffu, my_tok >= 0);
  cp_async_wait_all_t();
  __syncthreads();

  const uint32_t aQ = smem_u32(Qs), aK = smem_u32(Ks), aV = smem_u32(Vs);
  const uint32_t mQ = smem_u32(Cm), mK = mQ + 1024, mV = mQ + 2048;
  const int g = lane >> 2, t = lane & 3;
  const int lg = G.lgFp;
  uint32_t bq[4][2], bk[4][2];
  load_cls_frags(mQ, lane, bq);
  load_cls_frags(mK, lane, bk);
  float* lse_out = G.lse != nullptr ? G.lse + (static_cast<long long>(b) * G.H + h) * G.T : nullptr;

#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int R0 = warp * 32 + mt * 16;
    const int wr = mt * 16;
    uint32_t Pa[4], E0[4], Rc[4];
    float lse_a, lse_b, mc, lc;
    {
      uint32_t fq[4][4], fk[4][4];
      load_tile_frags(aQ, R0, lane, fq);
      load_tile_frags(aK, R0, lane, fk);
      float S[2][4] = {}, s0[4] = {}, sc[4] = {};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_bf16(S[nt], fq[ks], fk[ks][nt], fk[ks][2 + nt]);
        mma_bf16(s0, fq[ks], bk[ks][0], bk[ks][1]);            // q_i . k_cls (column 0)
        mma_bf16(sc, fk[ks], bq[ks][0], bq[ks][1]);            // k_j . q_cls (column 0)
      }
      const int ra = wr + g, rb = ra + 8;
      const bool va = (vm >> ra) & 1u, vb = (vm >> rb) & 1u;
      // the CLS-key score of rows g / g+8 lives in the t == 0 lane of the quad
      const float s0a = __shfl_sync(0xffffffffu, s0[0], lane & ~3), s0b = __shfl_sync(0xffffffffu, s0[2], lane & ~3);
      float ma = s0a, mb = s0b;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = wr + nt * 8 + 2 * t + e;
          const bool vk = (vm >> key) & 1u;
          if (!(va && vk && ((ra >> lg) == (key >> lg)))) S[nt][e] = -INFINITY;
          if (!(vb && vk && ((rb >> lg) == (key >> lg)))) S[nt][2 + e] = -INFINITY;
          ma = fmaxf(ma, S[nt][e]);
          mb = fmaxf(mb, S[nt][2 + e]);
        }
      }
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
      float la = 0.f, lb = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          S[nt][e] = __expf(S[nt][e] - ma); la += S[nt][e];
          S[nt][2 + e] = __expf(S[nt][2 + e] - mb); lb += S[nt][2 + e];
        }
      }
      la += __shfl_xor_sync(0xffffffffu, la, 1); la += __shfl_xor_sync(0xffffffffu, la, 2);
      lb += __shfl_xor_sync(0xffffffffu, lb, 1); lb += __shfl_xor_sync(0xffffffffu, lb, 2);
      const float p0a = __expf(s0a - ma), p0b = __expf(s0b - mb);
      la += p0a; lb += p0b;
      const float ia = 1.f / la, ib = 1.f / lb;
      lse_a = ma + __logf(la); lse_b = mb + __logf(lb);
      Pa[0] = pack_bf16x2(S[0][0] * ia, S[0][1] * ia); Pa[1] = pack_bf16x2(S[0][2] * ib, S[0][3] * ib);
      Pa[2] = pack_bf16x2(S[1][0] * ia, S[1][1] * ia); Pa[3] = pack_bf16x2(S[1][2] * ib, S[1][3] * ib);
      E0[0] = t == 0 ? pack_bf16x2(p0a * ia, 0.f) : 0u; E0[1] = t == 0 ? pack_bf16x2(p0b * ib, 0.f) : 0u; E0[2] = 0u; E0[3] = 0u;
      // CLS query vs the 16 keys of this tile (t == 0 lanes hold keys g, g+8)
      const float ca = (t == 0 && va) ? sc[0] : -INFINITY, cb = (t == 0 && vb) ? sc[2] : -INFINITY;
      mc = warp_max(fmaxf(ca, cb));
      const float pa = mc > -INFINITY ? __expf(ca - mc) : 0.f, pb = mc > -INFINITY ? __expf(cb - mc) : 0.f;
      lc = warp_sum(pa + pb);
      Rc[0] = movm_t(pack_bf16x2(pa, 0.f)); Rc[1] = 0u; Rc[2] = movm_t(pack_bf16x2(pb, 0.f)); Rc[3] = 0u;
    }
    float acc[8][4], acc2[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f; }
    mma_rows_t(aV, R0, lane, Pa, acc, Rc, acc2);               // O = P V ; CLS partial sum_j p_cj v_j (row 0 of acc2)
    mma_cls_t(mV, lane, E0, acc);                              // + p_i0 v_cls
    if (G.cls_part != nullptr) {
      float* part = G.cls_part + ((static_cast<long long>(bh) * G.chunks + chunk) * kClsParts + warp * 2 + mt) * kClsPartT;
      if (lane == 0) { part[0] = mc; part[1] = lc; }
      if (lane < 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          part[2 + 8 * j + 2 * lane] = acc2[j][0];
          part[2 + 8 * j + 2 * lane + 1] = acc2[j][1];
        }
      }
    }
    if (lse_out != nullptr && t == 0) {
      const int ta = row_token(G, chunk, warp, wr + g), tb = row_token(G, chunk, warp, wr + g + 8);
      if (ta >= 0) lse_out[ta] = lse_a;
      if (tb >= 0) lse_out[tb] = lse_b;
    }
    __syncwarp();                                              // every lane has its Q fragments of this tile
    store_acc_rows<false>(Qs, R0, lane, acc);
  }
  __syncwarp();
  store_rows_warp(Qs, G.out + row0 * G.ld_out + h * TD, G.ld_out, tok, warp, lane);
}


// Merges the per-tile partials of the CLS query with the (CLS query, CLS key) pair; writes out row 0 and lse[0].
// One CTA per (batch, head): 8 slices of 64 threads walk the partials 8-way interleaved (all loads of a slice are
// independent: the kernel is a latency chain otherwise), warp 0 meanwhile scores the (CLS, CLS) pair, then slice 0
// combines through shared memory.
constexpr int kCombSlices = 8;
__global__ void __launch_bounds__(kCombSlices * TD) attn_time_cls_combine_kernel(const TimeGeom G) {
  __shared__ float sM[kCombSlices], sL[kCombSlices], sO[kCombSlices][TD], sScc;
  const int bh = blockIdx.x, b = bh / G.H, h = bh - b * G.H;
  const int d = threadIdx.x & (TD - 1), slice = threadIdx.x >> 6;
  const int HD3 = G.H * TD;
  const int parts = G.chunks * kClsParts;
  const float* part = G.cls_part + static_cast<long long>(bh) * parts * kClsPartT;
  const __nv_bfloat16* base = G.qkv + static_cast<long long>(b) * G.T * G.ld_qkv + h * TD;
  if (threadIdx.x < 32) {                                       // q_cls . k_cls, two dims per lane
    const float2 qc = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + 2 * threadIdx.x));
    const float2 kc = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + HD3 + 2 * threadIdx.x));
    const float scc = warp_sum(qc.x * kc.x + qc.y * kc.y);
    if (threadIdx.x == 0) sScc = scc;
  }
  // slice-local max, then the running sums relative to it
  float m = -INFINITY;
#pragma unroll 8
  for (int p = slice; p < parts; p += kCombSlices) m = fmaxf(m, part[p * kClsPartT]);
  float l = 0.f, o = 0.f;
  if (m > -INFINITY) {
#pragma unroll 8
    for (int p = slice; p < parts; p += kCombSlices) {
      const float w = __expf(part[p * kClsPartT] - m);         // exp(-inf) = 0 for empty partials
      l = fmaf(part[p * kClsPartT + 1], w, l);
      o = fmaf(part[p * kClsPartT + 2 + d], w, o);
    }
  }
  if (d == 0) { sM[slice] = m; sL[slice] = l; }
  sO[slice][d] = o;
  __syncthreads();
  if (slice == 0) {
    const float scc = sScc;
    float M = scc;
#pragma unroll
    for (int s2 = 0; s2 < kCombSlices; ++s2) M = fmaxf(M, sM[s2]);
    const float pcc = __expf(scc - M);
    float L = pcc;
    float O = bf16_round(pcc) * __bfloat162float(base[2 * HD3 + d]);
#pragma unroll
    for (int s2 = 0; s2 < kCombSlices; ++s2) {
      const float w = sM[s2] > -INFINITY ? __expf(sM[s2] - M) : 0.f;
      L = fmaf(sL[s2], w, L);
      O = fmaf(sO[s2][d], w, O);
    }
    G.out[static_cast<long long>(b) * G.T * G.ld_out + h * TD + d] = __float2bfloat16_rn(O / L);
    if (d == 0 && G.lse != nullptr) G.lse[static_cast<long long>(bh) * G.T] = M + __logf(L);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// EXT_DELTA: delta = rowsum(dO * O) arrives from the GEMM that produced dO (oat_gemm_bf16 act 4) - O is not read at all.
template <bool EXT_DELTA>
__global__ void __launch_bounds__(kRows, 3) attn_time_bwd_kernel(const TimeGeom G) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm_time_raw[];
  uint8_t* Qs = sm_time_raw;
  uint8_t* Ks = Qs + kArr;
  uint8_t* Vs = Ks + kArr;
  uint8_t* Ds = Vs + kArr;                                                  // dO rows
  uint8_t* Cm = Ds + kArr;                                                  // 4 CLS matrices [8][128 B]: q, k, v, dO
  float* sLse = reinterpret_cast<float*>(Cm + 4 * 1024);                    // [kRows]
  float* sDelta = sLse + kRows;                                             // [kRows]
  float* sAcc = sDelta + kRows;                                             // [warp][3][64]: dq_cls, dk_cls, dv_cls partials
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // head-fastest block order: the H CTAs that read the 128-byte head slices of the SAME token rows (one 4.6 KB qkv row
  // holds all heads) are neighbours in launch order, so they hit the same DRAM pages at about the same time
  const int h = blockIdx.x % G.H, rest_ = blockIdx.x / G.H;
  const int chunk = rest_ % G.chunks, b = rest_ / G.chunks;
  const int bh = b * G.H + h;
  const int HD3 = G.H * TD;
  const long long row0 = static_cast<long long>(b) * G.T;
  const __nv_bfloat16* base = G.qkv + row0 * G.ld_qkv + h * TD;
  const __nv_bfloat16* dbase = G.dout + row0 * G.ld_dout + h * TD;
  const __nv_bfloat16* obase = G.out + row0 * G.ld_out + h * TD;
  const float* lrow = G.lse + (static_cast<long long>(b) * G.H + h) * G.T;
  int tok[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) tok[it] = row_token(G, chunk, warp, it * 4 + (lane >> 3));
  stage_rows_warp(Ds, dbase, G.ld_dout, tok, warp, lane);
  {
    uint8_t* const arrs[3] = {Vs, Ks, Qs};
    const __nv_bfloat16* const srcs[3] = {base + 2 * HD3, base + HD3, base};
    stage_rows_warp_n<3>(arrs, srcs, G.ld_qkv, tok, warp, lane);
  }
  {  // CLS matrices: row 0 <- vector, rows 1-7 <- 0
    const int m = threadIdx.x >> 5, rr = (threadIdx.x >> 3) & 3, c = threadIdx.x & 7;   // 4 matrices x (4 rows x 8 chunks) per pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 4 + rr;
      uint8_t* dst = Cm + m * 1024 + r * 128 + ((c ^ r) << 4);
      if (r == 0) cp_async16_t(smem_u32(dst), (m < 3 ? base + m * HD3 : dbase) + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  for (int i = threadIdx.x; i < kTimeWarps * 3 * TD; i += kRows) sAcc[i] = 0.f;
  float* wAcc = sAcc + warp * 3 * TD;
  uint4 oraw[8];
  if constexpr (!EXT_DELTA) {  // O chunks of the rows this lane helps with (coalesced), for delta = dO . O
    const int c = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it)
      oraw[it] = tok[it] >= 0 ? *reinterpret_cast<const uint4*>(obase + static_cast<long long>(tok[it]) * G.ld_out + c * 8)
                              : make_uint4(0u, 0u, 0u, 0u);
  }
  const int my_tok = row_token(G, chunk, warp, lane);
  const uint32_t vm = __ballot_sync(0xffffffffu, my_tok >= 0);              // valid rows of this warp
  sLse[threadIdx.x] = my_tok >= 0 ? lrow[my_tok] : 0.f;
  // CLS query statistics: lse_c, delta_c = dO_cls . O_cls (lane holds dims 2 * lane, +1)
  const float lse_c = lrow[0];
  float delta_c;
  if constexpr (EXT_DELTA) {
    const float* drow = G.delta + static_cast<long long>(h) * G.ld_delta + row0;
    sDelta[threadIdx.x] = my_tok >= 0 ? drow[my_tok] : 0.f;
    delta_c = drow[0];
  } else {
    const float2 dc = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dbase + 2 * lane));
    const float2 oc = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(obase + 2 * lane));
    delta_c = warp_sum(dc.x * oc.x + dc.y * oc.y);
  }
  cp_async_wait_all_t();
  __syncthreads();                                              // CLS matrices + sAcc zeroing are CTA-wide
  if constexpr (!EXT_DELTA) {
    const int c = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = warp * 32 + it * 4 + (lane >> 3);
      const uint4 dv = *reinterpret_cast<const uint4*>(Ds + sw_off(rr, c));
      const uint32_t x[4] = {dv.x, dv.y, dv.z, dv.w}, y[4] = {oraw[it].x, oraw[it].y, oraw[it].z, oraw[it].w};
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        part = fmaf(__uint_as_float(x[k] << 16), __uint_as_float(y[k] << 16), part);
        part = fmaf(__uint_as_float(x[k] & 0xffff0000u), __uint_as_float(y[k] & 0xffff0000u), part);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (c == 0) sDelta[rr] = part;
    }
  }
  __syncwarp();

  const uint32_t aQ = smem_u32(Qs), aK = smem_u32(Ks), aV = smem_u32(Vs), aD = smem_u32(Ds);
  const uint32_t mQ = smem_u32(Cm), mK = mQ + 1024, mV = mQ + 2048, mD = mQ + 3072;
  const int g = lane >> 2, t = lane & 3;
  const int lg = G.lgFp;
  uint32_t bq[4][2], bk[4][2], bv[4][2], bd[4][2];              // CLS vectors as B operands of the score-type products
  load_cls_frags(mQ, lane, bq);
  load_cls_frags(mK, lane, bk);
  load_cls_frags(mV, lane, bv);
  load_cls_frags(mD, lane, bd);

#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int R0 = warp * 32 + mt * 16;                         // first staged row of the tile
    const int wr = mt * 16;                                     // same, within the warp (validity mask bits)
    uint32_t Pa[4], dSa[4], E0[4], Ep[4], Ec[4], Epc[4];
    {
      uint32_t fq[4][4], fk[4][4], fv[4][4], fd[4][4];
      load_tile_frags(aQ, R0, lane, fq);
      load_tile_frags(aK, R0, lane, fk);
      load_tile_frags(aV, R0, lane, fv);
      load_tile_frags(aD, R0, lane, fd);
      float S[2][4] = {}, dP[2][4] = {}, s0[4] = {}, dp0[4] = {}, sc[4] = {}, dpc[4] = {};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma_bf16(S[nt], fq[ks], fk[ks][nt], fk[ks][2 + nt]);
          mma_bf16(dP[nt], fd[ks], fv[ks][nt], fv[ks][2 + nt]);
        }
        mma_bf16(s0, fq[ks], bk[ks][0], bk[ks][1]);            // q_i . k_cls       (column 0)
        mma_bf16(dp0, fd[ks], bv[ks][0], bv[ks][1]);           // dO_i . v_cls
        mma_bf16(sc, fk[ks], bq[ks][0], bq[ks][1]);            // k_j . q_cls
        mma_bf16(dpc, fv[ks], bd[ks][0], bd[ks][1]);           // v_j . dO_cls
      }
      // ---- element-wise: rows ra = g, rb = g + 8 of the tile; columns (keys) nt * 8 + 2t + {0, 1}
      const int ra = wr + g, rb = ra + 8;
      const bool va = (vm >> ra) & 1u, vb = (vm >> rb) & 1u;
      const float lse_a = sLse[warp * 32 + ra], lse_b = sLse[warp * 32 + rb];
      const float del_a = sDelta[warp * 32 + ra], del_b = sDelta[warp * 32 + rb];
      float P[2][4], dS[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = wr + nt * 8 + 2 * t + e;
          const bool vk = (vm >> key) & 1u;
          const bool oka = va && vk && ((ra >> lg) == (key >> lg));
          const bool okb = vb && vk && ((rb >> lg) == (key >> lg));
          P[nt][e] = oka ? __expf(S[nt][e] - lse_a) : 0.f;
          P[nt][2 + e] = okb ? __expf(S[nt][2 + e] - lse_b) : 0.f;
          dS[nt][e] = P[nt][e] * (dP[nt][e] - del_a);
          dS[nt][2 + e] = P[nt][2 + e] * (dP[nt][2 + e] - del_b);
        }
      }
      Pa[0] = pack_bf16x2(P[0][0], P[0][1]); Pa[1] = pack_bf16x2(P[0][2], P[0][3]);
      Pa[2] = pack_bf16x2(P[1][0], P[1][1]); Pa[3] = pack_bf16x2(P[1][2], P[1][3]);
      dSa[0] = pack_bf16x2(dS[0][0], dS[0][1]); dSa[1] = pack_bf16x2(dS[0][2], dS[0][3]);
      dSa[2] = pack_bf16x2(dS[1][0], dS[1][1]); dSa[3] = pack_bf16x2(dS[1][2], dS[1][3]);
      // CLS key column (valid in the t == 0 lanes: column 0 of the MMA tile) and CLS query row (same lanes, keys g, g+8)
      const bool c0 = (t == 0);
      const float p0a = (c0 && va) ? __expf(s0[0] - lse_a) : 0.f, p0b = (c0 && vb) ? __expf(s0[2] - lse_b) : 0.f;
      const float ds0a = p0a * (dp0[0] - del_a), ds0b = p0b * (dp0[2] - del_b);
      const float pca = (c0 && va) ? __expf(sc[0] - lse_c) : 0.f, pcb = (c0 && vb) ? __expf(sc[2] - lse_c) : 0.f;
      const float dsca = pca * (dpc[0] - delta_c), dscb = pcb * (dpc[2] - delta_c);
      E0[0] = pack_bf16x2(ds0a, 0.f); E0[1] = pack_bf16x2(ds0b, 0.f); E0[2] = 0u; E0[3] = 0u;
      Ep[0] = pack_bf16x2(p0a, 0.f); Ep[1] = pack_bf16x2(p0b, 0.f); Ep[2] = 0u; Ep[3] = 0u;
      Ec[0] = pack_bf16x2(dsca, 0.f); Ec[1] = pack_bf16x2(dscb, 0.f); Ec[2] = 0u; Ec[3] = 0u;
      Epc[0] = pack_bf16x2(pca, 0.f); Epc[1] = pack_bf16x2(pcb, 0.f); Epc[2] = 0u; Epc[3] = 0u;
    }
    // transposes for the key side, and the single-row A operands of the CLS-row reductions
    uint32_t PTa[4], dSTa[4], Rk[4], Rv[4], Rq[4];
    PTa[0] = movm_t(Pa[0]); PTa[1] = movm_t(Pa[2]); PTa[2] = movm_t(Pa[1]); PTa[3] = movm_t(Pa[3]);
    dSTa[0] = movm_t(dSa[0]); dSTa[1] = movm_t(dSa[2]); dSTa[2] = movm_t(dSa[1]); dSTa[3] = movm_t(dSa[3]);
    Rk[0] = movm_t(E0[0]); Rk[1] = 0u; Rk[2] = movm_t(E0[1]); Rk[3] = 0u;      // row 0 = ds_i0 over queries i
    Rv[0] = movm_t(Ep[0]); Rv[1] = 0u; Rv[2] = movm_t(Ep[1]); Rv[3] = 0u;      // row 0 = p_i0
    Rq[0] = movm_t(Ec[0]); Rq[1] = 0u; Rq[2] = movm_t(Ec[1]); Rq[3] = 0u;      // row 0 = dS_cj over keys j

    float acc[8][4], acc2[8][4];
    // ---- dQ = dS K + ds_i0 k_cls (-> V rows, dead since their fragments were loaded) ; dQ_cls += dS_c. K
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f; }
    mma_rows_t(aK, R0, lane, dSa, acc, Rq, acc2);
    mma_cls_t(mK, lane, E0, acc);
    add_row0(wAcc, lane, acc2);
    __syncwarp();                                               // every lane has loaded its V / K fragments of this tile
    store_acc_rows<true>(Vs, R0, lane, acc, G.scale);
    // ---- dK = dS^T Q + dS_cj q_cls (-> K rows) ; dK_cls += ds_.0 Q
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f; }
    mma_rows_t(aQ, R0, lane, dSTa, acc, Rk, acc2);
    mma_cls_t(mQ, lane, Ec, acc);
    add_row0(wAcc + TD, lane, acc2);
    store_acc_rows<false>(Ks, R0, lane, acc);
    // ---- dV = P^T dO + P_cj dO_cls (-> Q rows) ; dV_cls += p_.0 dO
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f; }
    mma_rows_t(aD, R0, lane, PTa, acc, Rv, acc2);
    mma_cls_t(mD, lane, Epc, acc);
    add_row0(wAcc + 2 * TD, lane, acc2);
    __syncwarp();                                               // every lane is done with the Q rows of this tile
    store_acc_rows<false>(Qs, R0, lane, acc);
  }
  __syncwarp();
  __nv_bfloat16* gbase = G.dqkv + row0 * G.ld_dqkv + h * TD;
  {
    const uint8_t* const arrs[3] = {Vs, Ks, Qs};                            // dq (already scaled), dk, dv
    __nv_bfloat16* const dsts[3] = {gbase, gbase + HD3, gbase + 2 * HD3};
    store_rows_warp_n<3>(arrs, dsts, G.ld_dqkv, tok, warp, lane);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * TD; i += kRows) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kTimeWarps; ++w) v += sAcc[w * 3 * TD + i];
    atomicAdd(G.cls_acc + static_cast<long long>(bh) * 3 * TD + i, v);
  }
}

// Adds the (CLS query, CLS key) pair and writes row 0 of dqkv: dq = scale * (acc_q + dS_cc k_c), dk = acc_k + dS_cc q_c,
// dv = acc_v + P_cc dO_c. One warp per (batch, head); lane holds dims 2*lane, 2*lane+1.
__global__ void attn_time_cls_finalize_kernel(const TimeGeom G) {
  const int bh = blockIdx.x, b = bh / G.H, h = bh - b * G.H, lane = threadIdx.x;
  const int HD3 = G.H * TD;
  const long long row0 = static_cast<long long>(b) * G.T;
  auto ld2 = [&](const __nv_bfloat16* p) { return unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p + 2 * lane)); };
  const float2 q = ld2(G.qkv + row0 * G.ld_qkv + h * TD), k = ld2(G.qkv + row0 * G.ld_qkv + HD3 + h * TD),
               v = ld2(G.qkv + row0 * G.ld_qkv + 2 * HD3 + h * TD), dO = ld2(G.dout + row0 * G.ld_dout + h * TD),
               o = ld2(G.out + row0 * G.ld_out + h * TD);
  const float s = warp_sum(q.x * k.x + q.y * k.y);
  const float dp = warp_sum(dO.x * v.x + dO.y * v.y);
  const float delta = warp_sum(dO.x * o.x + dO.y * o.y);
  const float p = __expf(s - G.lse[(static_cast<long long>(b) * G.H + h) * G.T]);
  const float ds = p * (dp - delta);
  const float* acc = G.cls_acc + static_cast<long long>(bh) * 3 * TD;
  __nv_bfloat16* dst = G.dqkv + row0 * G.ld_dqkv + h * TD + 2 * lane;
  *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2((acc[2 * lane] + ds * k.x) * G.scale, (acc[2 * lane + 1] + ds * k.y) * G.scale);
  *reinterpret_cast<uint32_t*>(dst + HD3) = pack_bf16x2(acc[TD + 2 * lane] + ds * q.x, acc[TD + 2 * lane + 1] + ds * q.y);
  *reinterpret_cast<uint32_t*>(dst + 2 * HD3) = pack_bf16x2(acc[2 * TD + 2 * lane] + p * dO.x, acc[2 * TD + 2 * lane + 1] + p * dO.y);
}

static TimeGeom make_time_geom(const oat_attn_args* a) {
  TimeGeom G;
  G.B = a->B; G.T = a->T; G.H = a->H; G.F = a->F; G.n = a->n;
  int fp = 1;
  while (fp < a->F) fp <<= 1;
  G.Fp = fp;
  G.lgFp = 0;
  while ((1 << G.lgFp) < fp) ++G.lgFp;
  G.gpc = kTimeWarps * (32 / fp);
  G.chunks = (a->n + G.gpc - 1) / G.gpc;
  G.ld_qkv = a->ld_qkv; G.ld_out = a->ld_out; G.ld_dout = a->ld_dout; G.ld_dqkv = a->ld_dqkv;
  G.qkv = reinterpret_cast<const __nv_bfloat16*>(a->qkv);
  G.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  G.lse = a->lse;
  G.dout = reinterpret_cast<const __nv_bfloat16*>(a->dout);
  G.dqkv = reinterpret_cast<__nv_bfloat16*>(a->dqkv);
  G.scale = a->scale; G.cls_acc = a->cls_acc;
  G.cls_part = nullptr;
  G.delta = a->delta; G.ld_delta = a->ld_delta;
  return G;
}

// fp32 words of workspace the fused CLS query of the time forward needs (one partial per warp of every CTA)
long long time_fwd_workspace_floats(int B, int H, int F, int n) {
  int fp = 1;
  while (fp < F) fp <<= 1;
  const int gpc = kTimeWarps * (32 / fp);
  const long long chunks = (n + gpc - 1) / gpc;
  return static_cast<long long>(B) * H * chunks * kClsParts * kClsPartT;
}

template <typename K>
static int set_smem_once(K kern, int bytes, bool* done, const char* who) {
  if (*done) return OAT_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "%s smem attr: %s", who, cudaGetErrorString(e));
  *done = true;
  return OAT_OK;
}

// Time attention forward. With a workspace in args->cls_acc (time_fwd_workspace_floats) the CLS query is fused
// (partials + combine kernel) and *cls_done is set; otherwise the caller runs attn_cls_fwd_kernel afterwards.
int launch_time_fwd(const oat_attn_args* a, cudaStream_t s, bool* cls_done) {
  TimeGeom G = make_time_geom(a);
  G.cls_part = a->cls_acc;
  *cls_done = G.cls_part != nullptr;
  const int grid = a->B * a->H * G.chunks;
  constexpr int smem = 3 * kArr + 3 * 1024;
  static bool done3 = false, done4 = false;
  // 4 CTAs / SM (128 registers) measured 5 % faster than 3 (144 registers): more loads in flight per SM
  static const bool four = getenv("OAT_TIME_FWD_CTAS") == nullptr || atoi(getenv("OAT_TIME_FWD_CTAS")) != 3;
  int rc;
  if (four) {
    if ((rc = set_smem_once(attn_time_fwd_kernel<4>, smem, &done4, "attn_time_fwd")) != OAT_OK) return rc;
    if (launch_pdl(attn_time_fwd_kernel<4>, dim3(grid), dim3(kRows), smem, s, G) != cudaSuccess) return check_launch("attn_time_fwd_kernel");
  } else {
    if ((rc = set_smem_once(attn_time_fwd_kernel<3>, smem, &done3, "attn_time_fwd")) != OAT_OK) return rc;
    if (launch_pdl(attn_time_fwd_kernel<3>, dim3(grid), dim3(kRows), smem, s, G) != cudaSuccess) return check_launch("attn_time_fwd_kernel");
  }
  rc = check_launch("attn_time_fwd_kernel");
  if (rc != OAT_OK || !*cls_done) return rc;
  attn_time_cls_combine_kernel<<<a->B * a->H, kCombSlices * TD, 0, s>>>(G);
  return check_launch("attn_time_cls_combine_kernel");
}

// full time-attention backward (patch rows + CLS row); cls_acc must be zeroed by the caller beforehand
int launch_time_bwd(const oat_attn_args* a, cudaStream_t s) {
  const TimeGeom G = make_time_geom(a);
  const int grid = a->B * a->H * G.chunks;
  constexpr int smem = 4 * kArr + 4 * 1024 + (2 * kRows + kTimeWarps * 3 * TD) * static_cast<int>(sizeof(float));
  static bool done0 = false, done1 = false;
  int rc;
  if (G.delta != nullptr) {
    if ((rc = set_smem_once(attn_time_bwd_kernel<true>, smem, &done1, "attn_time_bwd")) != OAT_OK) return rc;
    if (launch_pdl(attn_time_bwd_kernel<true>, dim3(grid), dim3(kRows), smem, s, G) != cudaSuccess) return check_launch("attn_time_bwd_kernel");
  } else {
    if ((rc = set_smem_once(attn_time_bwd_kernel<false>, smem, &done0, "attn_time_bwd")) != OAT_OK) return rc;
    if (launch_pdl(attn_time_bwd_kernel<false>, dim3(grid), dim3(kRows), smem, s, G) != cudaSuccess) return check_launch("attn_time_bwd_kernel");
  }
  rc = check_launch("attn_time_bwd_kernel");
  if (rc != OAT_OK) return rc;
  attn_time_cls_finalize_kernel<<<a->B * a->H, 32, 0, s>>>(G);
  return check_launch("attn_time_cls_finalize_kernel");
}

}  // namespace oat
