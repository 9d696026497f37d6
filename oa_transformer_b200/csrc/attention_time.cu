// Time attention (sequence length F + 1 <= 17) as an HBM-bound gather kernel: one THREAD per (batch, slot, head,
// frame) row, fp32 SIMT math, 128-bit row loads. Reference: VarAttention.forward with '(b n) f d' grouping,
// OATrans/model/video_transformer.py:112-122: token (f, i) attends to [CLS] + tokens (f', i) of every frame f'.
//
// A group (b, slot i, head h) has F queries and F + 1 keys of 64 dims: 2-8 FLOP per byte, far too small for 128-row
// tensor-core tiles, so the goal is to touch every 128-byte head slice once and keep the SM busy:
//   lane = group_in_warp * Fp + frame   (Fp = F rounded up to a power of two; 32 / Fp groups per warp)
//   every warp stages the head slices of its 32 token rows into shared memory with cp.async, 8 lanes per 128-byte
//   slice (full lines on the global side, all loads of a CTA in flight at once); every lane then owns one token row
//   and reads the other rows of its group from shared memory (broadcast). Result rows go back through shared memory
//   so that the global stores are full lines too.
// Backward computes the (F x (F+1)) probability / dS rows once on the query side, hands them to the key side through
// shared memory, and reduces the three CLS-row vectors (dQ of the CLS query, dK / dV of the CLS key) with a warp
// transpose-reduce -> shared memory -> one global atomic per component per CTA.
#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int TD = 64;            // head dim
constexpr int kTimeWarps = 4;     // 128 threads = 128 token rows staged per CTA
constexpr int TP = 72;            // smem row pitch (bf16): 144 B keeps the per-group broadcast reads conflict-free

struct TimeGeom {
  int B, T, H, F, n, Fp, gpc, chunks;   // gpc: groups per CTA, chunks: CTAs per (b, h)
  long long ld_qkv, ld_out, ld_dout, ld_dqkv;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  float* lse;
  const __nv_bfloat16* dout;
  __nv_bfloat16* dqkv;
  float scale;
  float* cls_acc;                       // [B*H][3][64]: dq_cls (unscaled), dk_cls, dv_cls
  float* cls_part;                      // forward: [B*H][chunks*warps][2+64] partials of the CLS query (or null)
};

__device__ __forceinline__ void load_row(const __nv_bfloat16* p, uint4 (&r)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) r[c] = reinterpret_cast<const uint4*>(p)[c];
}
__device__ __forceinline__ void unpack_row(const uint4 (&r)[8], float (&f)[TD]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t w[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[c * 8 + 2 * k] = __uint_as_float(w[k] << 16);
      f[c * 8 + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
}
// dot(f, row) and axpy with a packed bf16 row
__device__ __forceinline__ float dot_row(const float (&f)[TD], const uint4 (&r)[8]) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t w[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
    a0 = fmaf(f[c * 8 + 0], __uint_as_float(w[0] << 16), a0);
    a1 = fmaf(f[c * 8 + 1], __uint_as_float(w[0] & 0xffff0000u), a1);
    a2 = fmaf(f[c * 8 + 2], __uint_as_float(w[1] << 16), a2);
    a3 = fmaf(f[c * 8 + 3], __uint_as_float(w[1] & 0xffff0000u), a3);
    a0 = fmaf(f[c * 8 + 4], __uint_as_float(w[2] << 16), a0);
    a1 = fmaf(f[c * 8 + 5], __uint_as_float(w[2] & 0xffff0000u), a1);
    a2 = fmaf(f[c * 8 + 6], __uint_as_float(w[3] << 16), a2);
    a3 = fmaf(f[c * 8 + 7], __uint_as_float(w[3] & 0xffff0000u), a3);
  }
  return (a0 + a1) + (a2 + a3);
}
__device__ __forceinline__ void axpy_row(float a, const uint4 (&r)[8], float (&acc)[TD]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t w[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[c * 8 + 2 * k] = fmaf(a, __uint_as_float(w[k] << 16), acc[c * 8 + 2 * k]);
      acc[c * 8 + 2 * k + 1] = fmaf(a, __uint_as_float(w[k] & 0xffff0000u), acc[c * 8 + 2 * k + 1]);
    }
  }
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* p, const float (&f)[TD], float mul) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v;
    v.x = pack_bf16x2(f[c * 8 + 0] * mul, f[c * 8 + 1] * mul);
    v.y = pack_bf16x2(f[c * 8 + 2] * mul, f[c * 8 + 3] * mul);
    v.z = pack_bf16x2(f[c * 8 + 4] * mul, f[c * 8 + 5] * mul);
    v.w = pack_bf16x2(f[c * 8 + 6] * mul, f[c * 8 + 7] * mul);
    reinterpret_cast<uint4*>(p)[c] = v;
  }
}

// Sum a 64-vector over the 32 lanes of a warp; afterwards lane L holds components (2 * rev-index(L), +1) in v[0..1]
// and `base` tells which. 62 shuffles instead of 320.
__device__ __forceinline__ int warp_transpose_reduce(float (&v)[TD], int lane) {
  int base = 0;
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int width = 16 >> step;      // lane distance
    const int half = 32 >> step;       // components kept
    const bool upper = (lane & width) != 0;
#pragma unroll
    for (int c = 0; c < half; ++c) {
      const float send = upper ? v[c] : v[c + half];
      const float keep = upper ? v[c + half] : v[c];
      v[c] = keep + __shfl_xor_sync(0xffffffffu, send, width);
    }
    base += upper ? half : 0;
  }
  return base;                         // v[0], v[1] are components base, base + 1
}
__device__ __forceinline__ float dot_packed(const uint4 (&x)[8], const uint4 (&y)[8]) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t xw[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
    const uint32_t yw[4] = {y[c].x, y[c].y, y[c].z, y[c].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a0 = fmaf(__uint_as_float(xw[k] << 16), __uint_as_float(yw[k] << 16), a0);
      a1 = fmaf(__uint_as_float(xw[k] & 0xffff0000u), __uint_as_float(yw[k] & 0xffff0000u), a1);
    }
  }
  return a0 + a1;
}

// ------------------------------------------------------------------------------------------------ shared-memory rows
// A staged row is one head slice (64 bf16 = 128 B = 8 chunks of 16 B). Rows are packed at a 128-byte pitch and chunk c
// of row r lives at chunk position (c ^ (r & 7)): the 8 lanes that stage one row write one full 128-byte line (global
// side: one coalesced line per 8 lanes), a thread that reads ITS OWN row hits 8 distinct bank groups across a
// quarter-warp, and the lanes of a group that read the SAME row get a broadcast.
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return static_cast<uint32_t>(r) * 128u + ((static_cast<uint32_t>(c ^ (r & 7))) << 4); }
__device__ __forceinline__ void load_row_sw(const uint8_t* arr, int r, uint4 (&raw)[8]) {
  const uint8_t* row = arr + static_cast<uint32_t>(r) * 128u;
  const uint32_t x = static_cast<uint32_t>(r & 7) << 4;
#pragma unroll
  for (int c = 0; c < 8; ++c) raw[c] = *reinterpret_cast<const uint4*>(row + ((static_cast<uint32_t>(c) << 4) ^ x));
}
__device__ __forceinline__ void store_row_sw(uint8_t* arr, int r, const float (&f)[TD], float mul) {
  uint8_t* row = arr + static_cast<uint32_t>(r) * 128u;
  const uint32_t x = static_cast<uint32_t>(r & 7) << 4;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v;
    v.x = pack_bf16x2(f[c * 8 + 0] * mul, f[c * 8 + 1] * mul);
    v.y = pack_bf16x2(f[c * 8 + 2] * mul, f[c * 8 + 3] * mul);
    v.z = pack_bf16x2(f[c * 8 + 4] * mul, f[c * 8 + 5] * mul);
    v.w = pack_bf16x2(f[c * 8 + 6] * mul, f[c * 8 + 7] * mul);
    *reinterpret_cast<uint4*>(row + ((static_cast<uint32_t>(c) << 4) ^ x)) = v;
  }
}
__device__ __forceinline__ void load_row_lin(const uint8_t* row, uint4 (&raw)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) raw[c] = reinterpret_cast<const uint4*>(row)[c];
}

__device__ __forceinline__ void cp_async16_t(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_t() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Token of local row lr (0..31) of warp `warp` in CTA chunk `chunk`; -1 when the row is padding.
//   lr = group_in_warp * Fp + frame,  slot = chunk * gpc + warp * (32 / Fp) + group_in_warp,  token = 1 + frame * n + slot
__device__ __forceinline__ int row_token(const TimeGeom& G, int chunk, int warp, int lr) {
  const int gl = lr / G.Fp, i = lr - gl * G.Fp;
  const int pos = chunk * G.gpc + warp * (32 / G.Fp) + gl;
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

constexpr int kRows = kTimeWarps * 32;
constexpr int kArr = kRows * 128;            // bytes of one staged array (Q, K, V or dO rows of the CTA)
constexpr int kClsPartT = 2 + TD;            // per-warp partial of the CLS query: running max, sum, 64 weighted-V sums

// ------------------------------------------------------------------------------------------------ forward
// Every warp is autonomous (no block-wide barrier): it stages its 32 token rows of Q, K, V (+ the three CLS rows),
// each lane then runs the softmax row of its own token against [CLS key] + the F keys of its group, writes the
// output row back over its Q row and the warp stores the 32 output rows with full 128-byte lines.
// CLS query (video_transformer.py:108-110: attends to every token): every lane also scores its own key row against
// q_cls; the warp reduces (max, sum, sum p.v) to one partial in `cls_part`, merged by attn_time_cls_combine_kernel.
template <int KMAX>   // F + 1 <= KMAX
__global__ void __launch_bounds__(kRows, 4) attn_time_fwd_kernel(const TimeGeom G) {
  extern __shared__ __align__(128) uint8_t sm_time_raw[];
  uint8_t* Qs = sm_time_raw;
  uint8_t* Ks = Qs + kArr;
  uint8_t* Vs = Ks + kArr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* Cs = Vs + kArr + warp * (3 * 128);                  // this warp's copy of the CLS q / k / v rows
  const int bh = blockIdx.x / G.chunks, chunk = blockIdx.x - bh * G.chunks;
  const int b = bh / G.H, h = bh - b * G.H;
  const int HD3 = G.H * TD;
  const long long row0 = static_cast<long long>(b) * G.T;
  const __nv_bfloat16* base = G.qkv + row0 * G.ld_qkv + h * TD;
  int tok[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) tok[it] = row_token(G, chunk, warp, it * 4 + (lane >> 3));
  stage_rows_warp(Ks, base + HD3, G.ld_qkv, tok, warp, lane);
  stage_rows_warp(Vs, base + 2 * HD3, G.ld_qkv, tok, warp, lane);
  stage_rows_warp(Qs, base, G.ld_qkv, tok, warp, lane);
  if (lane < 24) cp_async16_t(smem_u32(Cs + (lane >> 3) * 128 + (lane & 7) * 16), base + (lane >> 3) * HD3 + (lane & 7) * 8);
  const int my_tok = row_token(G, chunk, warp, lane);
  const bool valid = my_tok >= 0;
  const int r = threadIdx.x;                                   // this lane's row in the staged arrays
  const int g0 = r - (lane % G.Fp);                            // row of frame 0 of this lane's group
  cp_async_wait_all_t();
  __syncwarp();

  uint4 raw[8];
  float a[TD];
  // ---- CLS query partial over this warp's 32 keys
  if (G.cls_part != nullptr) {
    uint4 kr[8];
    load_row_sw(Ks, r, kr);
    load_row_lin(Cs, raw);
    const float sc = valid ? dot_packed(kr, raw) : -INFINITY;
    const float mw = warp_max(sc);
    const float p = valid ? __expf(sc - mw) : 0.f;             // a warp without valid rows never gets here with mw finite
    const float lw = warp_sum(p);
    load_row_sw(Vs, r, raw);
    unpack_row(raw, a);
    const float pb = bf16_round(p);
#pragma unroll
    for (int d = 0; d < TD; ++d) a[d] *= pb;
    const int cb = warp_transpose_reduce(a, lane);
    float* part = G.cls_part + (static_cast<long long>(bh) * G.chunks * kTimeWarps + chunk * kTimeWarps + warp) * kClsPartT;
    const bool any = mw > -INFINITY;
    if (lane == 0) { part[0] = mw; part[1] = any ? lw : 0.f; }
    part[2 + cb] = any ? a[0] : 0.f;
    part[2 + cb + 1] = any ? a[1] : 0.f;
  }
  // ---- patch queries
  float lse_val = 0.f;
  {
    load_row_sw(Qs, r, raw);
    unpack_row(raw, a);                                        // a = q_i
    float s[KMAX];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (j <= G.F) {
        if (j == 0) load_row_lin(Cs + 128, raw); else load_row_sw(Ks, g0 + j - 1, raw);
        s[j] = dot_row(a, raw);
        m = fmaxf(m, s[j]);
      } else {
        s[j] = -INFINITY;
      }
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      s[j] = (j <= G.F) ? __expf(s[j] - m) : 0.f;
      l += s[j];
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (j <= G.F) {
        if (j == 0) load_row_lin(Cs + 256, raw); else load_row_sw(Vs, g0 + j - 1, raw);
        axpy_row(bf16_round(s[j] * inv), raw, a);
      }
    }
    lse_val = m + __logf(l);
  }
  store_row_sw(Qs, r, a, 1.f);                                 // own Q row is dead: it becomes the output staging row
  if (valid && G.lse != nullptr) G.lse[(static_cast<long long>(b) * G.H + h) * G.T + my_tok] = lse_val;
  __syncwarp();
  store_rows_warp(Qs, G.out + row0 * G.ld_out + h * TD, G.ld_out, tok, warp, lane);
}

// Merges the per-warp partials of the CLS query with the (CLS query, CLS key) pair; writes out row 0 and lse[0].
__global__ void __launch_bounds__(TD) attn_time_cls_combine_kernel(const TimeGeom G) {
  const int bh = blockIdx.x, b = bh / G.H, h = bh - b * G.H, d = threadIdx.x;
  const int HD3 = G.H * TD;
  const int parts = G.chunks * kTimeWarps;
  const float* part = G.cls_part + static_cast<long long>(bh) * parts * kClsPartT;
  const __nv_bfloat16* base = G.qkv + static_cast<long long>(b) * G.T * G.ld_qkv + h * TD;
  float scc = 0.f;
  for (int k = 0; k < TD; ++k) scc = fmaf(__bfloat162float(base[k]), __bfloat162float(base[HD3 + k]), scc);
  float M = scc;
  for (int p = 0; p < parts; ++p) M = fmaxf(M, part[p * kClsPartT]);
  const float pcc = __expf(scc - M);
  float L = pcc;
  float o = bf16_round(pcc) * __bfloat162float(base[2 * HD3 + d]);
  for (int p = 0; p < parts; ++p) {
    const float w = __expf(part[p * kClsPartT] - M);           // exp(-inf) = 0 for empty partials
    L = fmaf(part[p * kClsPartT + 1], w, L);
    o = fmaf(part[p * kClsPartT + 2 + d], w, o);
  }
  G.out[static_cast<long long>(b) * G.T * G.ld_out + h * TD + d] = __float2bfloat16_rn(o / L);
  if (d == 0 && G.lse != nullptr) G.lse[static_cast<long long>(bh) * G.T] = M + __logf(L);
}

// ------------------------------------------------------------------------------------------------ backward
// Per warp: stage Q, K, V, dO rows (coalesced), O rows straight into registers for delta. Query side: lane = query row
// (P / dS row, dQ, the CLS-key shares); P and dS go to shared memory (over the V rows, dead by then) for the key side:
// lane = key row (dK, dV). The CLS query's row of P / dS against this lane's key is recomputed from the staged CLS rows.
// Gradient rows are written over dead staged rows and stored by the warp as full 128-byte lines.
template <int KMAX>
__global__ void __launch_bounds__(kRows, 3) attn_time_bwd_kernel(const TimeGeom G) {
  extern __shared__ __align__(128) uint8_t sm_time_raw[];
  uint8_t* Qs = sm_time_raw;
  uint8_t* Ks = Qs + kArr;
  uint8_t* Vs = Ks + kArr;
  uint8_t* Ds = Vs + kArr;                                                  // dO rows
  uint8_t* Cs = Ds + kArr;                                                  // CLS rows: q, k, v, dO, O  [5][128 B]
  float* sDelta = reinterpret_cast<float*>(Cs + 5 * 128);                   // [kRows]
  float* sAcc = sDelta + kRows;                                             // [3][64] CTA accumulators for the CLS rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sP = reinterpret_cast<float*>(Vs + warp * (32 * 128));             // [32][KMAX-1] over this warp's V rows
  float* sDS = sP + 32 * (KMAX - 1);                                        // (2 * 32 * 16 * 4 B = 4 KB at KMAX = 17)
  const int bh = blockIdx.x / G.chunks, chunk = blockIdx.x - bh * G.chunks;
  const int b = bh / G.H, h = bh - b * G.H;
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
  stage_rows_warp(Vs, base + 2 * HD3, G.ld_qkv, tok, warp, lane);
  stage_rows_warp(Ks, base + HD3, G.ld_qkv, tok, warp, lane);
  stage_rows_warp(Qs, base, G.ld_qkv, tok, warp, lane);
  if (threadIdx.x < 40) {
    const int rr = threadIdx.x >> 3, c = threadIdx.x & 7;
    const __nv_bfloat16* src = rr < 3 ? base + rr * HD3 : (rr == 3 ? dbase : obase);
    cp_async16_t(smem_u32(Cs + rr * 128 + c * 16), src + c * 8);
  }
  for (int t = threadIdx.x; t < 3 * TD; t += kRows) sAcc[t] = 0.f;
  uint4 raw[8];
  {  // O chunks of the rows this lane helps with (coalesced), for delta = dO . O
    const int c = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it)
      raw[it] = tok[it] >= 0 ? *reinterpret_cast<const uint4*>(obase + static_cast<long long>(tok[it]) * G.ld_out + c * 8)
                             : make_uint4(0u, 0u, 0u, 0u);
  }
  const int my_tok = row_token(G, chunk, warp, lane);
  const bool valid = my_tok >= 0;
  const float lse = valid ? lrow[my_tok] : 0.f;
  const float lse_c = lrow[0];
  const int r = threadIdx.x;
  const int fi = lane % G.Fp;                                   // frame index of this lane
  const int g0 = r - fi;
  cp_async_wait_all_t();
  __syncthreads();                                              // CLS rows + sAcc zeroing are CTA-wide
  {
    const int c = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = warp * 32 + it * 4 + (lane >> 3);
      const uint4 dv = *reinterpret_cast<const uint4*>(Ds + sw_off(rr, c));
      const uint32_t x[4] = {dv.x, dv.y, dv.z, dv.w}, y[4] = {raw[it].x, raw[it].y, raw[it].z, raw[it].w};
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
  const float delta = sDelta[r];

  float a[TD];
  float p[KMAX], ds[KMAX];
  // ================= CLS query against this lane's key: P_cj, dS_cj and the dQ_cls share dS_cj * k_j =================
  float pc = 0.f, dsc = 0.f;
  {
    uint4 kv[8];
    load_row_sw(Ks, r, kv);                                     // k_own
    load_row_lin(Cs, raw);                                      // q_cls
    const float sc = dot_packed(kv, raw);
    unpack_row(kv, a);                                          // a = k_own
    load_row_sw(Vs, r, kv);                                     // v_own
    load_row_lin(Cs + 3 * 128, raw);                            // dO_cls
    const float dp = dot_packed(kv, raw);
    load_row_lin(Cs + 4 * 128, kv);                             // O_cls
    const float delta_c = dot_packed(raw, kv);
    if (valid) {
      pc = __expf(sc - lse_c);
      dsc = pc * (dp - delta_c);
    }
#pragma unroll
    for (int d = 0; d < TD; ++d) a[d] *= dsc;
    const int cb = warp_transpose_reduce(a, lane);
    atomicAdd(&sAcc[cb], a[0]); atomicAdd(&sAcc[cb + 1], a[1]);
  }
  // ================= query side: row i of P / dS, dQ_i, shares of dK_cls / dV_cls =================
  load_row_sw(Ds, r, raw);
  unpack_row(raw, a);                                           // a = dO_i
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {                              // dP_ij - delta_i = dO_i . v_j - delta_i
    ds[j] = 0.f;
    if (j <= G.F) {
      if (j == 0) load_row_lin(Cs + 2 * 128, raw); else load_row_sw(Vs, g0 + j - 1, raw);
      ds[j] = dot_row(a, raw) - delta;
    }
  }
  uint4 qr[8];
  load_row_sw(Qs, r, qr);                                       // own q row
  {
    // P_i0 first (q_i . k_cls), so that dO_i can be consumed in place: dV_cls share = P_i0 * dO_i
    load_row_lin(Cs + 128, raw);
    const float p0 = valid ? bf16_round(__expf(dot_packed(qr, raw) - lse)) : 0.f;
#pragma unroll
    for (int d = 0; d < TD; ++d) a[d] *= p0;
    const int cb = warp_transpose_reduce(a, lane);
    atomicAdd(&sAcc[2 * TD + cb], a[0]); atomicAdd(&sAcc[2 * TD + cb + 1], a[1]);
  }
  unpack_row(qr, a);                                            // a = q_i
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {                              // P_ij = exp(q_i . k_j - lse_i), dS = P (dP - delta)
    p[j] = 0.f;
    if (j <= G.F) {
      if (j == 0) load_row_lin(Cs + 128, raw); else load_row_sw(Ks, g0 + j - 1, raw);
      p[j] = valid ? __expf(dot_row(a, raw) - lse) : 0.f;
    }
    ds[j] *= p[j];
  }
  __syncwarp();                                                 // every lane is done with the V rows: they become P / dS
#pragma unroll
  for (int j = 1; j < KMAX; ++j) {
    sP[lane * (KMAX - 1) + j - 1] = p[j];
    sDS[lane * (KMAX - 1) + j - 1] = ds[j];
  }
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] *= ds[0];                   // dK_cls share = dS_i0 * q_i (in place)
  {
    const int cb = warp_transpose_reduce(a, lane);
    atomicAdd(&sAcc[TD + cb], a[0]); atomicAdd(&sAcc[TD + cb + 1], a[1]);
  }
  // dQ_i = sum_j dS_ij k_j
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j <= G.F) {
      if (j == 0) load_row_lin(Cs + 128, raw); else load_row_sw(Ks, g0 + j - 1, raw);
      axpy_row(ds[j], raw, a);
    }
  }
  __syncwarp();                                                 // K rows are dead: dQ rows take their place
  store_row_sw(Ks, r, a, G.scale);
  __syncwarp();
  store_rows_warp(Ks, G.dqkv + row0 * G.ld_dqkv + h * TD, G.ld_dqkv, tok, warp, lane);
  __syncwarp();

  // ================= key side: this lane's token as key j = fi + 1 =================
  // dK_j = sum_i dS_ij q_i + dS_cj q_cls
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
  for (int i = 0; i < KMAX - 1; ++i) {
    if (i < G.F) {
      load_row_sw(Qs, g0 + i, raw);
      axpy_row(valid ? sDS[(lane - fi + i) * (KMAX - 1) + fi] : 0.f, raw, a);
    }
  }
  load_row_lin(Cs, raw);
  axpy_row(dsc, raw, a);
  store_row_sw(Ks, r, a, 1.f);
  // dV_j = sum_i P_ij dO_i + P_cj dO_cls   (P rounded to bf16 like the forward's P.V operand)
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
  for (int i = 0; i < KMAX - 1; ++i) {
    if (i < G.F) {
      load_row_sw(Ds, g0 + i, raw);
      axpy_row(valid ? bf16_round(sP[(lane - fi + i) * (KMAX - 1) + fi]) : 0.f, raw, a);
    }
  }
  load_row_lin(Cs + 3 * 128, raw);
  axpy_row(bf16_round(pc), raw, a);
  __syncwarp();                                                 // every lane is done with the Q rows
  store_row_sw(Qs, r, a, 1.f);
  __syncwarp();
  store_rows_warp(Ks, G.dqkv + row0 * G.ld_dqkv + HD3 + h * TD, G.ld_dqkv, tok, warp, lane);
  store_rows_warp(Qs, G.dqkv + row0 * G.ld_dqkv + 2 * HD3 + h * TD, G.ld_dqkv, tok, warp, lane);
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * TD; t += kRows) atomicAdd(G.cls_acc + static_cast<long long>(bh) * 3 * TD + t, sAcc[t]);
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
  return G;
}

// fp32 words of workspace the fused CLS query of the time forward needs (one partial per warp of every CTA)
long long time_fwd_workspace_floats(int B, int H, int F, int n) {
  int fp = 1;
  while (fp < F) fp <<= 1;
  const int gpc = kTimeWarps * (32 / fp);
  const long long chunks = (n + gpc - 1) / gpc;
  return static_cast<long long>(B) * H * chunks * kTimeWarps * kClsPartT;
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
  constexpr int smem = 3 * kArr + kTimeWarps * 3 * 128;
  static bool d5 = false, d9 = false, d17 = false;
  int rc;
  if (a->F + 1 <= 5) {
    if ((rc = set_smem_once(attn_time_fwd_kernel<5>, smem, &d5, "attn_time_fwd")) != OAT_OK) return rc;
    attn_time_fwd_kernel<5><<<grid, kRows, smem, s>>>(G);
  } else if (a->F + 1 <= 9) {
    if ((rc = set_smem_once(attn_time_fwd_kernel<9>, smem, &d9, "attn_time_fwd")) != OAT_OK) return rc;
    attn_time_fwd_kernel<9><<<grid, kRows, smem, s>>>(G);
  } else {
    if ((rc = set_smem_once(attn_time_fwd_kernel<17>, smem, &d17, "attn_time_fwd")) != OAT_OK) return rc;
    attn_time_fwd_kernel<17><<<grid, kRows, smem, s>>>(G);
  }
  rc = check_launch("attn_time_fwd_kernel");
  if (rc != OAT_OK || !*cls_done) return rc;
  attn_time_cls_combine_kernel<<<a->B * a->H, TD, 0, s>>>(G);
  return check_launch("attn_time_cls_combine_kernel");
}

// full time-attention backward (patch rows + CLS row); cls_acc must be zeroed by the caller beforehand
int launch_time_bwd(const oat_attn_args* a, cudaStream_t s) {
  const TimeGeom G = make_time_geom(a);
  const int grid = a->B * a->H * G.chunks;
  constexpr int smem = 4 * kArr + 5 * 128 + (kRows + 3 * TD) * static_cast<int>(sizeof(float));
  static bool d5 = false, d9 = false, d17 = false;
  int rc;
  if (a->F + 1 <= 5) {
    if ((rc = set_smem_once(attn_time_bwd_kernel<5>, smem, &d5, "attn_time_bwd")) != OAT_OK) return rc;
    attn_time_bwd_kernel<5><<<grid, kRows, smem, s>>>(G);
  } else if (a->F + 1 <= 9) {
    if ((rc = set_smem_once(attn_time_bwd_kernel<9>, smem, &d9, "attn_time_bwd")) != OAT_OK) return rc;
    attn_time_bwd_kernel<9><<<grid, kRows, smem, s>>>(G);
  } else {
    if ((rc = set_smem_once(attn_time_bwd_kernel<17>, smem, &d17, "attn_time_bwd")) != OAT_OK) return rc;
    attn_time_bwd_kernel<17><<<grid, kRows, smem, s>>>(G);
  }
  rc = check_launch("attn_time_bwd_kernel");
  if (rc != OAT_OK) return rc;
  attn_time_cls_finalize_kernel<<<a->B * a->H, 32, 0, s>>>(G);
  return check_launch("attn_time_cls_finalize_kernel");
}

}  // namespace oat
