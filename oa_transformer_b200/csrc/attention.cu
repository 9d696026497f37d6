// Divided space / time attention, the global CLS attention, and masked text self-attention: forward + backward.
//
// Reference: VarAttention.forward, OATrans/model/video_transformer.py:99-135 (+ attn() :28-32): q is pre-scaled
// (:105, done by the qkv GEMM epilogue here), the CLS query attends to every key (:110), patch queries attend to
// [CLS] + their own frame (space, 'b (f n) d -> (b f) n d') or [CLS] + their own spatial slot across frames (time,
// '-> (b n) f d') (:112-122). Text: HF DistilBERT MultiHeadSelfAttention with a key-padding mask.
//
// Layout: one CTA per attention group (batch, frame|slot, head). The group's Q/K/V (and dO) rows - 128-byte head
// slices gathered straight out of the [B*T, 3*H*64] qkv GEMM output, the CLS row included once instead of being
// materialised n times as the reference does (:115-119) - are staged in shared memory with cp.async; the
// contractions run on bf16 tensor-core MMA (m16n8k16, fp32 accumulate) from ldmatrix fragments; softmax is fp32
// with quad-shuffle row reductions; results leave through shared memory as full 128-byte rows. Nothing of size
// (n x n) ever touches HBM. Backward recomputes P from the saved log-sum-exp; the CLS query rides along as one extra
// query row of every group (its dQ and the CLS key's dK/dV are reduced across groups with fp32 atomics).
//
// The per-group work is HBM/latency-bound for time attention (sequence length F+1) and small-tile tensor work for
// space attention; the tcgen05/TMEM variant of the space kernel is the next step.
#include <stdlib.h>

#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int HD = 64;       // head dim
constexpr int PITCH = 72;    // smem row pitch in bf16 (144 B): conflict-free ldmatrix
constexpr int PITCH_B = PITCH * 2;

struct AttnGeom {
  int mode, B, T, H, F, n;
  long long ld_qkv, ld_out, ld_dout, ld_dqkv;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  float* lse;
  const int* key_mask;
  const __nv_bfloat16* dout;
  __nv_bfloat16* dqkv;
  float scale;
  float* cls_acc;
  // mode 2: dropout on the softmax weights (see oat_attn_args)
  uint32_t drop_thresh, drop_site;
  float drop_inv_keep;
  unsigned long long drop_seed;
};

struct Group {
  int b, h, g;           // batch, head, frame (space) / slot (time)
  int nq, nk, has_cls;
  int q0, qs;            // query i -> token q0 + i*qs ; key j (j >= has_cls) -> token q0 + (j-has_cls)*qs
};

__device__ __forceinline__ Group decode_group(const AttnGeom& G, int idx) {
  Group r;
  r.h = idx % G.H;
  int rest = idx / G.H;
  if (G.mode == 0) {
    r.g = rest % G.F; r.b = rest / G.F;
    r.nq = G.n; r.nk = G.n + 1; r.has_cls = 1; r.q0 = 1 + r.g * G.n; r.qs = 1;
  } else if (G.mode == 1) {
    r.g = rest % G.n; r.b = rest / G.n;
    r.nq = G.F; r.nk = G.F + 1; r.has_cls = 1; r.q0 = 1 + r.g; r.qs = G.n;
  } else {
    r.g = 0; r.b = rest;
    r.nq = G.T; r.nk = G.T; r.has_cls = 0; r.q0 = 0; r.qs = 1;
  }
  return r;
}
__device__ __forceinline__ int key_token(const Group& g, int j) {
  return (g.has_cls && j == 0) ? 0 : g.q0 + (j - g.has_cls) * g.qs;
}
__device__ __forceinline__ int query_token(const Group& g, int i) {  // i == nq -> the CLS query (backward only)
  return (i >= g.nq) ? 0 : g.q0 + i * g.qs;
}

__device__ __forceinline__ void cp_async16(uint32_t smem, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A-operand fragments (16 rows x 64) of the tile starting at smem row `row0`
__device__ __forceinline__ void load_a_frags(uint32_t base, int row0, int lane, uint32_t (&a)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(base + (row0 + (lane & 15)) * PITCH_B + (ks * 16 + (lane >> 4) * 8) * 2, a[ks]);
}
// C[16 x 32] += A(16 x 64) . M[rows r0..r0+31][64]^T   (contraction over the 64 head dims; M rows are the N dim)
__device__ __forceinline__ void mma_a_rowsT(float (&c)[4][4], const uint32_t (&a)[4][4], uint32_t mbase, int r0,
                                            int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      ldsm_x4(mbase + (r0 + np * 16 + (lane & 7) + ((lane >> 4) << 3)) * PITCH_B + (ks * 16 + ((lane >> 3) & 1) * 8) * 2, b);
      mma16816(c[2 * np], a[ks], b[0], b[1]);
      mma16816(c[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// C[16 x 64] += P(16 x 32, as two k16 A fragments) . M[rows r0..r0+31][64]   (contraction over the 32 rows)
__device__ __forceinline__ void mma_p_rows(float (&c)[8][4], const uint32_t (&pa)[2][4], uint32_t mbase, int r0,
                                           int lane) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(mbase + (r0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH_B + (dp * 16 + (lane >> 4) * 8) * 2, b);
      mma16816(c[2 * dp], pa[kk], b[0], b[1]);
      mma16816(c[2 * dp + 1], pa[kk], b[2], b[3]);
    }
  }
}
// fp32 C fragments of a 16 x 32 tile -> bf16 A fragments for two k16 steps
__device__ __forceinline__ void c_to_a(const float (&c)[4][4], uint32_t (&a)[2][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    a[kk][0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
    a[kk][1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
    a[kk][2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    a[kk][3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// stage `count` head-slice rows (128 B each) of one matrix into smem rows [0, count); zero rows [count, rows_alloc)
template <typename TokFn>
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src_base, long long ld, int count,
                                           int rows_alloc, TokFn tok_of, int tid, int nthreads) {
  for (int idx = tid; idx < rows_alloc * 8; idx += nthreads) {
    const int r = idx >> 3, c = idx & 7;
    __nv_bfloat16* d = dst + r * PITCH + c * 8;
    if (r < count) {
      cp_async16(smem_u32(d), src_base + static_cast<long long>(tok_of(r)) * ld + c * 8);
    } else {
      *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
    }
  }
}
// write a warp's 16 x 64 fp32 tile (C fragments) as bf16 into its staging rows, then out as 128-byte rows
template <typename RowFn>
__device__ __forceinline__ void store_tile_bf16(const float (&c)[8][4], __nv_bfloat16* stg, int lane, int valid_rows,
                                                RowFn gptr_of_row) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(stg + g * PITCH + nt * 8 + 2 * t) = pack_bf16x2(c[nt][0], c[nt][1]);
    *reinterpret_cast<uint32_t*>(stg + (g + 8) * PITCH + nt * 8 + 2 * t) = pack_bf16x2(c[nt][2], c[nt][3]);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 4 + (lane >> 3), ch = lane & 7;
    if (r < valid_rows) {
      __nv_bfloat16* gp = gptr_of_row(r);
      if (gp != nullptr) *reinterpret_cast<uint4*>(gp + ch * 8) = *reinterpret_cast<const uint4*>(stg + r * PITCH + ch * 8);
    }
  }
  __syncwarp();
}

// =============================================================================================== forward
template <int ROWS, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_kernel(const AttnGeom G) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Ks = Qs + ROWS * PITCH;
  __nv_bfloat16* Vs = Ks + ROWS * PITCH;
  const Group gr = decode_group(G, blockIdx.x);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HDIM = G.H * HD;
  const __nv_bfloat16* base = G.qkv + static_cast<long long>(gr.b) * G.T * G.ld_qkv + gr.h * HD;
  stage_rows(Qs, base, G.ld_qkv, gr.nq, ROWS, [&](int i) { return query_token(gr, i); }, tid, NWARPS * 32);
  stage_rows(Ks, base + HDIM, G.ld_qkv, gr.nk, ROWS, [&](int j) { return key_token(gr, j); }, tid, NWARPS * 32);
  stage_rows(Vs, base + 2 * HDIM, G.ld_qkv, gr.nk, ROWS, [&](int j) { return key_token(gr, j); }, tid, NWARPS * 32);
  cp_async_wait_all();
  __syncthreads();

  const uint32_t qb = smem_u32(Qs), kb = smem_u32(Ks), vb = smem_u32(Vs);
  const int g = lane >> 2, t = lane & 3;
  const int* kmask = (G.key_mask != nullptr) ? G.key_mask + static_cast<long long>(gr.b) * G.T : nullptr;
  const int q_tiles = (gr.nq + 15) >> 4, k_chunks = (gr.nk + 31) >> 5;
  for (int qt = warp; qt < q_tiles; qt += NWARPS) {
    uint32_t qa[4][4];
    load_a_frags(qb, qt * 16, lane, qa);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    for (int kc = 0; kc < k_chunks; ++kc) {
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      mma_a_rowsT(s, qa, kb, kc * 32, lane);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = kc * 32 + nt * 8 + 2 * t + e;
          const bool ok = j < gr.nk && (kmask == nullptr || kmask[j] != 0);
          if (!ok) { s[nt][e] = -INFINITY; s[nt][2 + e] = -INFINITY; }
        }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float c0 = (mn0 == -INFINITY) ? 1.f : __expf(m0 - mn0), c1 = (mn1 == -INFINITY) ? 1.f : __expf(m1 - mn1);
      const float b0 = (mn0 == -INFINITY) ? 0.f : mn0, b1 = (mn1 == -INFINITY) ? 0.f : mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = __expf(s[nt][0] - b0); s[nt][1] = __expf(s[nt][1] - b0);
        s[nt][2] = __expf(s[nt][2] - b1); s[nt][3] = __expf(s[nt][3] - b1);
        rs0 += s[nt][0] + s[nt][1]; rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
      m0 = mn0; m1 = mn1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      if (G.drop_thresh != 0u) {
        // dropout acts on the normalised weights: the row sum above keeps every weight, the P.V product only the kept ones
        const unsigned long long row_base = (static_cast<unsigned long long>(gr.b) * G.H + gr.h) * G.T;
        const unsigned long long r0 = (row_base + qt * 16 + g) * G.T, r1 = (row_base + qt * 16 + g + 8) * G.T;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = kc * 32 + nt * 8 + 2 * t + e;
            s[nt][e] = dropout_keep(G.drop_seed, G.drop_site, r0 + j, G.drop_thresh) ? s[nt][e] * G.drop_inv_keep : 0.f;
            s[nt][2 + e] = dropout_keep(G.drop_seed, G.drop_site, r1 + j, G.drop_thresh) ? s[nt][2 + e] * G.drop_inv_keep : 0.f;
          }
        }
      }
      uint32_t pa[2][4];
      c_to_a(s, pa);
      mma_p_rows(o, pa, vb, kc * 32, lane);
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] *= i0; o[i][1] *= i0; o[i][2] *= i1; o[i][3] *= i1; }
    const int valid = min(16, gr.nq - qt * 16);
    if (G.lse != nullptr && t == 0) {
      float* lrow = G.lse + (static_cast<long long>(gr.b) * G.H + gr.h) * G.T;
      if (g < valid) lrow[query_token(gr, qt * 16 + g)] = m0 + __logf(l0);
      if (g + 8 < valid) lrow[query_token(gr, qt * 16 + g + 8)] = m1 + __logf(l1);
    }
    __syncwarp();
    // this warp's Q rows are dead now: reuse them as the output staging tile
    store_tile_bf16(o, Qs + qt * 16 * PITCH, lane, valid, [&](int r) {
      return G.out + (static_cast<long long>(gr.b) * G.T + query_token(gr, qt * 16 + r)) * G.ld_out + gr.h * HD;
    });
  }
}

// CLS query: one CTA per (batch, head); scores for all T keys in smem, block softmax, weighted V sum.
constexpr int kClsWarps = 8;
__global__ void __launch_bounds__(kClsWarps * 32) attn_cls_fwd_kernel(const AttnGeom G) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  float* sc = reinterpret_cast<float*>(smem_attn);        // [T]
  float* red = sc + ((G.T + 3) & ~3);                      // [kClsWarps * 64] + [2 * kClsWarps]
  const int b = blockIdx.x / G.H, h = blockIdx.x % G.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HDIM = G.H * HD;
  const __nv_bfloat16* base = G.qkv + static_cast<long long>(b) * G.T * G.ld_qkv + h * HD;
  const float2 q = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + 2 * lane));
  float mx = -INFINITY;
  for (int j = warp; j < G.T; j += kClsWarps) {
    const float2 k = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + static_cast<long long>(j) * G.ld_qkv + HDIM + 2 * lane));
    const float s = warp_sum(q.x * k.x + q.y * k.y);
    if (lane == 0) sc[j] = s;
    mx = fmaxf(mx, s);
  }
  float* wred = red + kClsWarps * 64;
  if (lane == 0) wred[warp] = mx;
  __syncthreads();
  mx = wred[0];
#pragma unroll
  for (int w = 1; w < kClsWarps; ++w) mx = fmaxf(mx, wred[w]);
  float sum = 0.f;
  for (int j = tid; j < G.T; j += kClsWarps * 32) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) wred[kClsWarps + warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < kClsWarps; ++w) sum += wred[kClsWarps + w];
  const float inv = 1.f / sum;
  float o0 = 0.f, o1 = 0.f;
  for (int j = warp; j < G.T; j += kClsWarps) {
    const float p = bf16_round(sc[j] * inv);
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + static_cast<long long>(j) * G.ld_qkv + 2 * HDIM + 2 * lane));
    o0 += p * v.x; o1 += p * v.y;
  }
  red[warp * 64 + 2 * lane] = o0;
  red[warp * 64 + 2 * lane + 1] = o1;
  __syncthreads();
  if (tid < 64) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kClsWarps; ++w) a += red[w * 64 + tid];
    G.out[static_cast<long long>(b) * G.T * G.ld_out + h * HD + tid] = __float2bfloat16_rn(a);
    if (tid == 0 && G.lse != nullptr) G.lse[(static_cast<long long>(b) * G.H + h) * G.T] = mx + __logf(sum);
  }
}

// =============================================================================================== backward
template <int ROWS, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_kernel(const AttnGeom G) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Ks = Qs + ROWS * PITCH;
  __nv_bfloat16* Vs = Ks + ROWS * PITCH;
  __nv_bfloat16* Ds = Vs + ROWS * PITCH;                  // dO
  __nv_bfloat16* Stg = Ds + ROWS * PITCH;                 // [NWARPS][16][PITCH]
  float* lse_s = reinterpret_cast<float*>(Stg + NWARPS * 16 * PITCH);  // [ROWS]
  float* del_s = lse_s + ROWS;                                           // [ROWS]
  const Group gr = decode_group(G, blockIdx.x);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HDIM = G.H * HD;
  const int nqe = gr.nq + gr.has_cls;                     // the CLS query is row nq
  const long long row0 = static_cast<long long>(gr.b) * G.T;
  const __nv_bfloat16* base = G.qkv + row0 * G.ld_qkv + gr.h * HD;
  const __nv_bfloat16* dbase = G.dout + row0 * G.ld_dout + gr.h * HD;
  stage_rows(Qs, base, G.ld_qkv, nqe, ROWS, [&](int i) { return query_token(gr, i); }, tid, NWARPS * 32);
  stage_rows(Ks, base + HDIM, G.ld_qkv, gr.nk, ROWS, [&](int j) { return key_token(gr, j); }, tid, NWARPS * 32);
  stage_rows(Vs, base + 2 * HDIM, G.ld_qkv, gr.nk, ROWS, [&](int j) { return key_token(gr, j); }, tid, NWARPS * 32);
  stage_rows(Ds, dbase, G.ld_dout, nqe, ROWS, [&](int i) { return query_token(gr, i); }, tid, NWARPS * 32);
  // delta_i = dO_i . O_i (8 lanes per row, 16 B each), lse_i
  const __nv_bfloat16* obase = G.out + row0 * G.ld_out + gr.h * HD;
  const float* lrow = G.lse + (static_cast<long long>(gr.b) * G.H + gr.h) * G.T;
  for (int idx = tid; idx < ROWS * 8; idx += NWARPS * 32) {
    const int r = idx >> 3, c = idx & 7;
    float part = 0.f;
    if (r < nqe) {
      const long long tok = query_token(gr, r);
      const uint4 a = *reinterpret_cast<const uint4*>(dbase + tok * G.ld_dout + c * 8);
      const uint4 o = *reinterpret_cast<const uint4*>(obase + tok * G.ld_out + c * 8);
      const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* op = reinterpret_cast<const uint32_t*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = unpack_bf16x2(ap[k]), y = unpack_bf16x2(op[k]);
        part += x.x * y.x + x.y * y.y;
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (c == 0) {
      del_s[r] = part;
      lse_s[r] = (r < nqe) ? lrow[query_token(gr, r)] : 0.f;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const uint32_t qb = smem_u32(Qs), kb = smem_u32(Ks), vb = smem_u32(Vs), db = smem_u32(Ds);
  const int g = lane >> 2, t = lane & 3;
  const int* kmask = (G.key_mask != nullptr) ? G.key_mask + static_cast<long long>(gr.b) * G.T : nullptr;
  const bool own_cls_pair = (gr.g == 0);   // the (CLS query, CLS key) pair is counted by one group only
  __nv_bfloat16* stg = Stg + warp * 16 * PITCH;
  float* acc = (G.cls_acc != nullptr) ? G.cls_acc + (static_cast<long long>(gr.b) * G.H + gr.h) * 3 * HD : nullptr;
  const int q_tiles = (nqe + 15) >> 4, k_tiles = (gr.nk + 15) >> 4;
  const int k_chunks = (gr.nk + 31) >> 5, q_chunks = (nqe + 31) >> 5;

  // ---------------- pass A: dQ (query-tile major)
  for (int qt = warp; qt < q_tiles; qt += NWARPS) {
    uint32_t qa[4][4], da[4][4];
    load_a_frags(qb, qt * 16, lane, qa);
    load_a_frags(db, qt * 16, lane, da);
    const int i0 = qt * 16 + g, i1 = i0 + 8;
    const float ls0 = lse_s[i0], ls1 = lse_s[i1], dl0 = del_s[i0], dl1 = del_s[i1];
    const bool cls0 = gr.has_cls && i0 == gr.nq, cls1 = gr.has_cls && i1 == gr.nq;
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kc = 0; kc < k_chunks; ++kc) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
      mma_a_rowsT(s, qa, kb, kc * 32, lane);
      mma_a_rowsT(dp, da, vb, kc * 32, lane);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = kc * 32 + nt * 8 + 2 * t + e;
          const bool ok = j < gr.nk && (kmask == nullptr || kmask[j] != 0);
          const bool pair_ok = !(gr.has_cls && j == 0) || own_cls_pair;   // only matters for the CLS query rows
          const float p0 = (ok && i0 < nqe && (!cls0 || pair_ok)) ? __expf(s[nt][e] - ls0) : 0.f;
          const float p1 = (ok && i1 < nqe && (!cls1 || pair_ok)) ? __expf(s[nt][2 + e] - ls1) : 0.f;
          float dp0 = dp[nt][e], dp1 = dp[nt][2 + e];
          if (G.drop_thresh != 0u) {      // dP = keep(dO . V^T) / (1 - p); delta already holds sum_j P dP
            const unsigned long long rb = (static_cast<unsigned long long>(gr.b) * G.H + gr.h) * G.T;
            dp0 = dropout_keep(G.drop_seed, G.drop_site, (rb + i0) * G.T + j, G.drop_thresh) ? dp0 * G.drop_inv_keep : 0.f;
            dp1 = dropout_keep(G.drop_seed, G.drop_site, (rb + i1) * G.T + j, G.drop_thresh) ? dp1 * G.drop_inv_keep : 0.f;
          }
          s[nt][e] = p0 * (dp0 - dl0);
          s[nt][2 + e] = p1 * (dp1 - dl1);
        }
      }
      uint32_t dsa[2][4];
      c_to_a(s, dsa);
      mma_p_rows(dq, dsa, kb, kc * 32, lane);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { dq[i][0] *= G.scale; dq[i][1] *= G.scale; dq[i][2] *= G.scale; dq[i][3] *= G.scale; }
    if (gr.has_cls && (qt * 16 + 15 >= gr.nq) && (qt * 16 <= gr.nq) && acc != nullptr) {
      // the CLS query row lives in this tile: reduce its partial dQ across groups
      if (cls0) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { atomicAdd(acc + nt * 8 + 2 * t, dq[nt][0]); atomicAdd(acc + nt * 8 + 2 * t + 1, dq[nt][1]); }
      }
      if (cls1) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { atomicAdd(acc + nt * 8 + 2 * t, dq[nt][2]); atomicAdd(acc + nt * 8 + 2 * t + 1, dq[nt][3]); }
      }
    }
    const int valid = min(16, gr.nq - qt * 16);   // patch queries only
    store_tile_bf16(dq, stg, lane, valid, [&](int r) {
      return G.dqkv + (row0 + query_token(gr, qt * 16 + r)) * G.ld_dqkv + gr.h * HD;
    });
  }

  // ---------------- pass B: dK, dV (key-tile major)
  for (int kt = warp; kt < k_tiles; kt += NWARPS) {
    uint32_t ka[4][4], va[4][4];
    load_a_frags(kb, kt * 16, lane, ka);
    load_a_frags(vb, kt * 16, lane, va);
    const int j0 = kt * 16 + g, j1 = j0 + 8;
    const bool kok0 = j0 < gr.nk && (kmask == nullptr || kmask[j0] != 0);
    const bool kok1 = j1 < gr.nk && (kmask == nullptr || kmask[j1] != 0);
    const bool kcls0 = gr.has_cls && j0 == 0;     // j1 >= 8 is never the CLS key
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qc = 0; qc < q_chunks; ++qc) {
      float st[4][4], dpt[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
      mma_a_rowsT(st, ka, qb, qc * 32, lane);     // S^T = K Q^T
      mma_a_rowsT(dpt, va, db, qc * 32, lane);    // dP^T = V dO^T
      float pt[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = qc * 32 + nt * 8 + 2 * t + e;
          const bool qok = i < nqe;
          const bool qcls = gr.has_cls && i == gr.nq;
          const float ls = lse_s[i], dl = del_s[i];
          const float p0 = (qok && kok0 && !(qcls && kcls0 && !own_cls_pair)) ? __expf(st[nt][e] - ls) : 0.f;
          const float p1 = (qok && kok1) ? __expf(st[nt][2 + e] - ls) : 0.f;
          float m0k = 1.f, m1k = 1.f;
          if (G.drop_thresh != 0u) {      // weight (query i, key j0 / j1): the same draw as in the forward pass
            const unsigned long long rb = ((static_cast<unsigned long long>(gr.b) * G.H + gr.h) * G.T + i) * G.T;
            m0k = dropout_keep(G.drop_seed, G.drop_site, rb + j0, G.drop_thresh) ? G.drop_inv_keep : 0.f;
            m1k = dropout_keep(G.drop_seed, G.drop_site, rb + j1, G.drop_thresh) ? G.drop_inv_keep : 0.f;
          }
          pt[nt][e] = p0 * m0k; pt[nt][2 + e] = p1 * m1k;
          st[nt][e] = p0 * (dpt[nt][e] * m0k - dl);
          st[nt][2 + e] = p1 * (dpt[nt][2 + e] * m1k - dl);
        }
      }
      uint32_t pa[2][4], dsa[2][4];
      c_to_a(pt, pa);
      c_to_a(st, dsa);
      mma_p_rows(dv, pa, db, qc * 32, lane);      // dV += P^T dO
      mma_p_rows(dk, dsa, qb, qc * 32, lane);     // dK += dS^T Q
    }
    if (gr.has_cls && kt == 0 && g == 0 && acc != nullptr) {
      // CLS key row (j = 0): reduce across groups
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        atomicAdd(acc + HD + nt * 8 + 2 * t, dk[nt][0]); atomicAdd(acc + HD + nt * 8 + 2 * t + 1, dk[nt][1]);
        atomicAdd(acc + 2 * HD + nt * 8 + 2 * t, dv[nt][0]); atomicAdd(acc + 2 * HD + nt * 8 + 2 * t + 1, dv[nt][1]);
      }
    }
    const int valid = min(16, gr.nk - kt * 16);
    auto krow = [&](int r, int which) -> __nv_bfloat16* {
      const int j = kt * 16 + r;
      if (gr.has_cls && j == 0) return nullptr;
      return G.dqkv + (row0 + key_token(gr, j)) * G.ld_dqkv + which * HDIM + gr.h * HD;
    };
    store_tile_bf16(dk, stg, lane, valid, [&](int r) { return krow(r, 1); });
    store_tile_bf16(dv, stg, lane, valid, [&](int r) { return krow(r, 2); });
  }
}

// fp32 cross-group accumulators [B*H][3][64] -> row 0 (CLS token) of dqkv
__global__ void attn_cls_finalize_kernel(const AttnGeom G) {
  const int bh = blockIdx.x, b = bh / G.H, h = bh % G.H;
  const int i = threadIdx.x;  // 0..191
  const int which = i / HD, d = i % HD;
  G.dqkv[static_cast<long long>(b) * G.T * G.ld_dqkv + which * G.H * HD + h * HD + d] =
      __float2bfloat16_rn(G.cls_acc[static_cast<long long>(bh) * 3 * HD + i]);
}

static AttnGeom to_geom(const oat_attn_args* a) {
  AttnGeom G;
  G.mode = a->mode; G.B = a->B; G.T = a->T; G.H = a->H; G.F = a->F; G.n = a->n;
  G.ld_qkv = a->ld_qkv; G.ld_out = a->ld_out; G.ld_dout = a->ld_dout; G.ld_dqkv = a->ld_dqkv;
  G.qkv = reinterpret_cast<const __nv_bfloat16*>(a->qkv);
  G.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  G.lse = a->lse; G.key_mask = a->key_mask;
  G.dout = reinterpret_cast<const __nv_bfloat16*>(a->dout);
  G.dqkv = reinterpret_cast<__nv_bfloat16*>(a->dqkv);
  G.scale = a->scale; G.cls_acc = a->cls_acc;
  const bool drop = a->mode == 2 && a->dropout_p > 0.f && a->dropout_p < 1.f;
  G.drop_thresh = drop ? static_cast<uint32_t>(static_cast<double>(a->dropout_p) * 4294967296.0) : 0u;
  G.drop_inv_keep = drop ? 1.0f / (1.0f - a->dropout_p) : 1.0f;
  G.drop_site = a->dropout_site;
  G.drop_seed = a->dropout_seed;
  return G;
}

static int check_geom(const oat_attn_args* a, const char* who, int* rows_needed, int* groups) {
  if (a == nullptr) return set_error(OAT_ERR_ARG, "%s: null args", who);
  if (a->B <= 0 || a->H <= 0 || a->T <= 0) return set_error(OAT_ERR_ARG, "%s: bad B/H/T", who);
  if (a->ld_qkv % 8 != 0 || a->ld_out % 8 != 0) return set_error(OAT_ERR_ARG, "%s: row pitches must be multiples of 8", who);
  if (a->mode == 0 || a->mode == 1) {
    if (a->F <= 0 || a->n <= 0 || a->T != 1 + a->F * a->n)
      return set_error(OAT_ERR_ARG, "%s: T=%d must equal 1 + F*n (F=%d, n=%d)", who, a->T, a->F, a->n);
    *rows_needed = (a->mode == 0 ? a->n : a->F) + 1;
    *groups = a->B * a->H * (a->mode == 0 ? a->F : a->n);
  } else if (a->mode == 2) {
    *rows_needed = a->T;
    *groups = a->B * a->H;
  } else {
    return set_error(OAT_ERR_ARG, "%s: unknown mode %d", who, a->mode);
  }
  if (*rows_needed > 256)
    return set_error(OAT_ERR_ARG, "%s: %d keys per group exceed the 256-row shared-memory tile", who, *rows_needed);
  return OAT_OK;
}

template <int ROWS, int NWARPS>
static int launch_fwd(const AttnGeom& G, int groups, cudaStream_t s) {
  const int smem = 3 * ROWS * PITCH_B;
  auto kern = attn_fwd_kernel<ROWS, NWARPS>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_fwd smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  kern<<<groups, NWARPS * 32, smem, s>>>(G);
  return check_launch("attn_fwd_kernel");
}
template <int ROWS, int NWARPS>
static int launch_bwd(const AttnGeom& G, int groups, cudaStream_t s) {
  const int smem = 4 * ROWS * PITCH_B + NWARPS * 16 * PITCH_B + 2 * ROWS * 4;
  auto kern = attn_bwd_kernel<ROWS, NWARPS>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_bwd smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  kern<<<groups, NWARPS * 32, smem, s>>>(G);
  return check_launch("attn_bwd_kernel");
}

// SIMT gather kernels for time attention (attention_time.cu): sequence length F + 1 <= 17
int launch_time_fwd(const oat_attn_args* a, cudaStream_t s, bool* cls_done);
long long time_fwd_workspace_floats(int B, int H, int F, int n);
int launch_time_bwd(const oat_attn_args* a, cudaStream_t s);
constexpr int kTimeSimtMaxF = 16;
// tcgen05 / TMEM space attention with the CLS query fused (attention_space_tc.cu): 128 <= n <= 255
bool space_tc_fwd_supported(const oat_attn_args* a);
int launch_space_tc_fwd(const oat_attn_args* a, cudaStream_t s);
bool space_tc_bwd_supported(const oat_attn_args* a);
int launch_space_tc_bwd(const oat_attn_args* a, cudaStream_t s);

}  // namespace oat

extern "C" size_t oat_attn_fwd_workspace_floats(int32_t mode, int32_t B, int32_t H, int32_t F, int32_t n) {
  if (B <= 0 || H <= 0 || F <= 0 || n <= 0) return 0;
  if (mode == 0) return static_cast<size_t>(B) * H * F * 66;
  if (mode == 1 && F <= oat::kTimeSimtMaxF) return static_cast<size_t>(oat::time_fwd_workspace_floats(B, H, F, n));
  return 0;
}

extern "C" int oat_attn_fwd(const oat_attn_args* a, oat_stream_t stream) {
  using namespace oat;
  int rows = 0, groups = 0;
  int rc = check_geom(a, "oat_attn_fwd", &rows, &groups);
  if (rc != OAT_OK) return rc;
  OAT_REQUIRE(a->qkv != nullptr && a->out != nullptr, "oat_attn_fwd: null qkv/out");
  const AttnGeom G = to_geom(a);
  cudaStream_t s = as_stream(stream);
  if (space_tc_fwd_supported(a)) return launch_space_tc_fwd(a, s);
  if (a->mode == 1 && a->F <= kTimeSimtMaxF) {
    bool cls_done = false;
    rc = launch_time_fwd(a, s, &cls_done);
    if (rc != OAT_OK || cls_done) return rc;
  } else if (rows <= 32) rc = launch_fwd<32, 2>(G, groups, s);
  else if (rows <= 64) rc = launch_fwd<64, 4>(G, groups, s);
  else if (rows <= 128) rc = launch_fwd<128, 4>(G, groups, s);
  else rc = launch_fwd<256, 8>(G, groups, s);
  if (rc != OAT_OK) return rc;
  if (a->mode != 2) {
    const int smem = (((a->T + 3) & ~3) + kClsWarps * 64 + 2 * kClsWarps) * 4;
    static int max_set = 0;
    if (smem > 48 * 1024 && smem > max_set) {
      cudaError_t e = cudaFuncSetAttribute(attn_cls_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "attn_cls smem attr: %s", cudaGetErrorString(e));
      max_set = smem;
    }
    attn_cls_fwd_kernel<<<a->B * a->H, kClsWarps * 32, smem, s>>>(G);
    rc = check_launch("attn_cls_fwd_kernel");
  }
  return rc;
}

extern "C" int oat_attn_bwd(const oat_attn_args* a, oat_stream_t stream) {
  using namespace oat;
  int rows = 0, groups = 0;
  int rc = check_geom(a, "oat_attn_bwd", &rows, &groups);
  if (rc != OAT_OK) return rc;
  OAT_REQUIRE(a->qkv != nullptr && a->out != nullptr && a->dout != nullptr && a->dqkv != nullptr && a->lse != nullptr,
              "oat_attn_bwd: null tensor");
  OAT_REQUIRE(a->ld_dout % 8 == 0 && a->ld_dqkv % 8 == 0, "oat_attn_bwd: row pitches must be multiples of 8");
  OAT_REQUIRE(a->mode == 2 || a->cls_acc != nullptr, "oat_attn_bwd: cls_acc workspace required for space/time modes");
  const AttnGeom G = to_geom(a);
  cudaStream_t s = as_stream(stream);
  if (a->mode != 2) {
    cudaError_t e = cudaMemsetAsync(a->cls_acc, 0, sizeof(float) * a->B * a->H * 3 * HD, s);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    if (a->mode == 1 && a->F <= kTimeSimtMaxF) return launch_time_bwd(a, s);
    if (space_tc_bwd_supported(a) && !a->key_mask && getenv("OAT_SPACE_BWD_LEGACY") == nullptr) {
      rc = launch_space_tc_bwd(a, s);
      if (rc != OAT_OK) return rc;
      attn_cls_finalize_kernel<<<a->B * a->H, 3 * HD, 0, s>>>(G);
      return check_launch("attn_cls_finalize_kernel");
    }
    rows += 1;  // the CLS query row
    if (rows > 256) return set_error(OAT_ERR_ARG, "oat_attn_bwd: group too large for the 256-row tile");
  }
  if (rows <= 32) rc = launch_bwd<32, 2>(G, groups, s);
  else if (rows <= 64) rc = launch_bwd<64, 4>(G, groups, s);
  else if (rows <= 128) rc = launch_bwd<128, 8>(G, groups, s);
  else rc = launch_bwd<256, 8>(G, groups, s);
  if (rc != OAT_OK) return rc;
  if (a->mode != 2) {
    attn_cls_finalize_kernel<<<a->B * a->H, 3 * HD, 0, s>>>(G);
    rc = check_launch("attn_cls_finalize_kernel");
  }
  return rc;
}
