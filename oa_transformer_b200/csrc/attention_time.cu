// Time attention (sequence length F + 1 <= 17) as an HBM-bound gather kernel: one THREAD per (batch, slot, head,
// frame) row, fp32 SIMT math, 128-bit row loads. Reference: VarAttention.forward with '(b n) f d' grouping,
// OATrans/model/video_transformer.py:112-122: token (f, i) attends to [CLS] + tokens (f', i) of every frame f'.
//
// A group (b, slot i, head h) has F queries and F + 1 keys of 64 dims: 2-8 FLOP per byte, far too small for 128-row
// tensor-core tiles, so the goal is to touch every 128-byte head slice once and keep the SM busy:
//   lane = group_in_warp * Fp + frame   (Fp = F rounded up to a power of two; 32 / Fp groups per warp)
//   every lane owns one token row and stages that token's head slices into shared memory with cp.async (all loads
//   of a CTA are in flight at once); the other rows of its group are then read from shared memory (broadcast).
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

struct TimeLane {
  int b, h, pos, i, gl;   // gl: group index within the warp
  bool valid;
  long long row0;         // b * T
  int tok;
};
__device__ __forceinline__ TimeLane decode_lane(const TimeGeom& G, int lane, int warp) {
  TimeLane L;
  const int bh = blockIdx.x / G.chunks, chunk = blockIdx.x - bh * G.chunks;
  L.b = bh / G.H; L.h = bh - L.b * G.H;
  L.gl = lane / G.Fp; L.i = lane - L.gl * G.Fp;
  L.pos = chunk * G.gpc + warp * (32 / G.Fp) + L.gl;
  L.valid = L.i < G.F && L.pos < G.n;
  L.row0 = static_cast<long long>(L.b) * G.T;
  L.tok = 1 + L.i * G.n + L.pos;
  return L;
}

__device__ __forceinline__ void cp_async16_t(void* smem, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_t() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void stage_row(__nv_bfloat16* dst, const __nv_bfloat16* src) {
#pragma unroll
  for (int c = 0; c < 8; ++c) cp_async16_t(dst + c * 8, src + c * 8);
}

// ------------------------------------------------------------------------------------------------ forward
template <int KMAX>   // F + 1 <= KMAX
__global__ void __launch_bounds__(kTimeWarps * 32) attn_time_fwd_kernel(const TimeGeom G) {
  __shared__ __align__(16) __nv_bfloat16 Ks[kTimeWarps * 32 * TP];
  __shared__ __align__(16) __nv_bfloat16 Vs[kTimeWarps * 32 * TP];
  __shared__ __align__(16) __nv_bfloat16 Cs[2 * TP];           // CLS key / value rows of this (b, h)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const TimeLane L = decode_lane(G, lane, warp);
  const int HD3 = G.H * TD;
  const __nv_bfloat16* base = G.qkv + L.row0 * G.ld_qkv + L.h * TD;
  const long long own = static_cast<long long>(L.valid ? L.tok : 0);
  stage_row(Ks + threadIdx.x * TP, base + own * G.ld_qkv + HD3);
  stage_row(Vs + threadIdx.x * TP, base + own * G.ld_qkv + 2 * HD3);
  if (threadIdx.x < 16) cp_async16_t(Cs + (threadIdx.x >> 3) * TP + (threadIdx.x & 7) * 8,
                                     base + (1 + (threadIdx.x >> 3)) * HD3 + (threadIdx.x & 7) * 8);
  uint4 raw[8];
  float q[TD];
  load_row(base + own * G.ld_qkv, raw);
  unpack_row(raw, q);
  cp_async_wait_all_t();
  __syncthreads();
  if (!L.valid) return;
  const int g0 = threadIdx.x - L.i;          // smem row of frame 0 of this lane's group
  float s[KMAX];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j <= G.F) {
      load_row(j == 0 ? Cs : Ks + (g0 + j - 1) * TP, raw);
      s[j] = dot_row(q, raw);
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
  float o[TD];
#pragma unroll
  for (int d = 0; d < TD; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j <= G.F) {
      load_row(j == 0 ? Cs + TP : Vs + (g0 + j - 1) * TP, raw);
      axpy_row(bf16_round(s[j] * inv), raw, o);
    }
  }
  store_row_bf16(G.out + (L.row0 + L.tok) * G.ld_out + L.h * TD, o, 1.f);
  if (G.lse != nullptr) G.lse[(static_cast<long long>(L.b) * G.H + L.h) * G.T + L.tok] = m + __logf(l);
}

// ------------------------------------------------------------------------------------------------ backward
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

template <int KMAX>
__global__ void __launch_bounds__(kTimeWarps * 32, 4) attn_time_bwd_kernel(const TimeGeom G) {
  extern __shared__ __align__(16) uint8_t sm_time_raw[];
  constexpr int kRows = kTimeWarps * 32;
  // two row buffers, used twice: (K, V) for the query side, then (Q, dO) for the key side
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(sm_time_raw);     // [kRows][TP]
  __nv_bfloat16* Ys = Xs + kRows * TP;
  __nv_bfloat16* Cs = Ys + kRows * TP;                                     // CLS rows: q, k, v, dO, O  [5][TP]
  float* sm_f = reinterpret_cast<float*>(Cs + 5 * TP);
  float* sP = sm_f + (threadIdx.x >> 5) * (2 * 32 * KMAX);                 // per warp: P[lane][key], dS likewise
  float* sDS = sP + 32 * KMAX;
  float* sAcc = sm_f + kTimeWarps * (2 * 32 * KMAX);                       // [3][64] CTA accumulators for the CLS rows

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const TimeLane L = decode_lane(G, lane, warp);
  const int HD3 = G.H * TD;
  const __nv_bfloat16* base = G.qkv + L.row0 * G.ld_qkv + L.h * TD;
  const __nv_bfloat16* dbase = G.dout + L.row0 * G.ld_dout + L.h * TD;
  const __nv_bfloat16* obase = G.out + L.row0 * G.ld_out + L.h * TD;
  const float* lrow = G.lse + (static_cast<long long>(L.b) * G.H + L.h) * G.T;
  const long long own = static_cast<long long>(L.valid ? L.tok : 0);

  // stage K, V rows of this CTA's 128 tokens and the CLS rows; every load of the CTA is in flight at once
  stage_row(Xs + threadIdx.x * TP, base + own * G.ld_qkv + HD3);
  stage_row(Ys + threadIdx.x * TP, base + own * G.ld_qkv + 2 * HD3);
  if (threadIdx.x < 40) {
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    const __nv_bfloat16* src = r < 3 ? base + r * HD3 : (r == 3 ? dbase : obase);
    cp_async16_t(Cs + r * TP + c * 8, src + c * 8);
  }
  for (int t = threadIdx.x; t < 3 * TD; t += blockDim.x) sAcc[t] = 0.f;
  uint4 raw[8];
  float a[TD];                 // fp32 row whose role changes per phase
  float p[KMAX], ds[KMAX];
  load_row(dbase + own * G.ld_dout, raw);                  // own dO row
  unpack_row(raw, a);                                      // a = dO_i
  load_row(obase + own * G.ld_out, raw);                   // own O row (only needed for delta)
  const float delta = dot_row(a, raw);
  const float lse = lrow[own];
  cp_async_wait_all_t();
  __syncthreads();
  const int g0 = threadIdx.x - L.i;                        // smem row of frame 0 of this lane's group

  // ================= query side: row i of P / dS, dQ_i, shares of dK_cls / dV_cls =================
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {                         // dP_ij - delta_i = dO_i . v_j - delta_i
    ds[j] = 0.f;
    if (j <= G.F) {
      load_row(j == 0 ? Cs + 2 * TP : Ys + (g0 + j - 1) * TP, raw);
      ds[j] = dot_row(a, raw) - delta;
    }
  }
  uint4 qr[8];
  load_row(base + own * G.ld_qkv, qr);                     // own q row
  {
    // P_i0 first (q_i . k_cls), so that dO_i can be consumed in place: dV_cls share = P_i0 * dO_i
    load_row(Cs + TP, raw);
    const float p0 = L.valid ? bf16_round(__expf(dot_packed(qr, raw) - lse)) : 0.f;
#pragma unroll
    for (int d = 0; d < TD; ++d) a[d] *= p0;
    warp_transpose_reduce(a, lane);
    atomicAdd(&sAcc[2 * TD + 2 * lane], a[0]); atomicAdd(&sAcc[2 * TD + 2 * lane + 1], a[1]);
  }
  unpack_row(qr, a);                                       // a = q_i
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {                         // P_ij = exp(q_i . k_j - lse_i), dS = P (dP - delta)
    p[j] = 0.f;
    if (j <= G.F) {
      load_row(j == 0 ? Cs + TP : Xs + (g0 + j - 1) * TP, raw);
      p[j] = L.valid ? __expf(dot_row(a, raw) - lse) : 0.f;
    }
    ds[j] *= p[j];
    sP[lane * KMAX + j] = p[j];
    sDS[lane * KMAX + j] = ds[j];
  }
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] *= ds[0];              // dK_cls share = dS_i0 * q_i (in place)
  warp_transpose_reduce(a, lane);
  atomicAdd(&sAcc[TD + 2 * lane], a[0]); atomicAdd(&sAcc[TD + 2 * lane + 1], a[1]);
  // dQ_i = sum_j dS_ij k_j
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j <= G.F) {
      load_row(j == 0 ? Cs + TP : Xs + (g0 + j - 1) * TP, raw);
      axpy_row(ds[j], raw, a);
    }
  }
  if (L.valid) store_row_bf16(G.dqkv + (L.row0 + L.tok) * G.ld_dqkv + L.h * TD, a, G.scale);
  __syncthreads();                                         // everyone is done with K / V: restage Q and dO
  stage_row(Xs + threadIdx.x * TP, base + own * G.ld_qkv);
  stage_row(Ys + threadIdx.x * TP, dbase + own * G.ld_dout);

  // ================= key side: this lane's token as key j = i + 1 =================
  float pc = 0.f, dsc = 0.f;
  {
    uint4 kv[8];
    load_row(base + own * G.ld_qkv + HD3, kv);             // k_own (L1/L2 hit: staged a moment ago)
    load_row(Cs, raw);                                     // q_cls
    const float sc = dot_packed(kv, raw);
    load_row(base + own * G.ld_qkv + 2 * HD3, kv);         // v_own
    load_row(Cs + 3 * TP, raw);                            // dO_cls
    const float dp = dot_packed(kv, raw);
    load_row(Cs + 4 * TP, kv);                             // O_cls
    const float delta_c = dot_packed(raw, kv);
    if (L.valid) {
      pc = __expf(sc - lrow[0]);
      dsc = pc * (dp - delta_c);
    }
  }
  cp_async_wait_all_t();
  __syncthreads();
  const int jk = L.i + 1;
  // dK_j = sum_i dS_ij q_i + dS_cj q_cls
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
  for (int i = 0; i < KMAX - 1; ++i) {
    if (i < G.F) {
      load_row(Xs + (g0 + i) * TP, raw);
      axpy_row(L.valid ? sDS[(lane - L.i + i) * KMAX + jk] : 0.f, raw, a);
    }
  }
  load_row(Cs, raw);
  axpy_row(dsc, raw, a);
  if (L.valid) store_row_bf16(G.dqkv + (L.row0 + L.tok) * G.ld_dqkv + HD3 + L.h * TD, a, 1.f);
  // dV_j = sum_i P_ij dO_i + P_cj dO_cls   (P rounded to bf16 like the forward's P.V operand)
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] = 0.f;
#pragma unroll
  for (int i = 0; i < KMAX - 1; ++i) {
    if (i < G.F) {
      load_row(Ys + (g0 + i) * TP, raw);
      axpy_row(L.valid ? bf16_round(sP[(lane - L.i + i) * KMAX + jk]) : 0.f, raw, a);
    }
  }
  load_row(Cs + 3 * TP, raw);
  axpy_row(bf16_round(pc), raw, a);
  if (L.valid) store_row_bf16(G.dqkv + (L.row0 + L.tok) * G.ld_dqkv + 2 * HD3 + L.h * TD, a, 1.f);
  // dQ_cls share = dS_cj * k_j
  load_row(base + own * G.ld_qkv + HD3, raw);
  unpack_row(raw, a);
#pragma unroll
  for (int d = 0; d < TD; ++d) a[d] *= dsc;
  warp_transpose_reduce(a, lane);
  atomicAdd(&sAcc[2 * lane], a[0]); atomicAdd(&sAcc[2 * lane + 1], a[1]);
  __syncthreads();
  const int bh = blockIdx.x / G.chunks;
  for (int t = threadIdx.x; t < 3 * TD; t += blockDim.x) atomicAdd(G.cls_acc + static_cast<long long>(bh) * 3 * TD + t, sAcc[t]);
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
  return G;
}

// patch-query part of the time attention forward (the CLS query is handled by attn_cls_fwd_kernel in attention.cu)
int launch_time_fwd(const oat_attn_args* a, cudaStream_t s) {
  const TimeGeom G = make_time_geom(a);
  const int grid = a->B * a->H * G.chunks;
  if (a->F + 1 <= 5) attn_time_fwd_kernel<5><<<grid, kTimeWarps * 32, 0, s>>>(G);
  else if (a->F + 1 <= 9) attn_time_fwd_kernel<9><<<grid, kTimeWarps * 32, 0, s>>>(G);
  else attn_time_fwd_kernel<17><<<grid, kTimeWarps * 32, 0, s>>>(G);
  return check_launch("attn_time_fwd_kernel");
}

// full time-attention backward (patch rows + CLS row); cls_acc must be zeroed by the caller beforehand
int launch_time_bwd(const oat_attn_args* a, cudaStream_t s) {
  const TimeGeom G = make_time_geom(a);
  const int grid = a->B * a->H * G.chunks;
  auto smem_for = [](int kmax) {
    return static_cast<int>((2 * kTimeWarps * 32 + 5) * TP * 2 + (kTimeWarps * 2 * 32 * kmax + 3 * TD) * sizeof(float));
  };
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_time_bwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(5));
    cudaFuncSetAttribute(attn_time_bwd_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(9));
    cudaFuncSetAttribute(attn_time_bwd_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(17));
    attr_done = true;
  }
  if (a->F + 1 <= 5) attn_time_bwd_kernel<5><<<grid, kTimeWarps * 32, smem_for(5), s>>>(G);
  else if (a->F + 1 <= 9) attn_time_bwd_kernel<9><<<grid, kTimeWarps * 32, smem_for(9), s>>>(G);
  else attn_time_bwd_kernel<17><<<grid, kTimeWarps * 32, smem_for(17), s>>>(G);
  int rc = check_launch("attn_time_bwd_kernel");
  if (rc != OAT_OK) return rc;
  attn_time_cls_finalize_kernel<<<a->B * a->H, 32, 0, s>>>(G);
  return check_launch("attn_time_cls_finalize_kernel");
}

}  // namespace oat
