// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A(M,K) . B(N,K)^T  (fp32 accumulate in TMEM)
//
//   warp 0      : TMEM allocation, then TMA producer (one elected lane)
//   warp 1      : mbarrier init, then tcgen05.mma issuer (one elected lane)
//   warps 2..9  : epilogue (two warps per TMEM lane quarter, alternating column chunks). Specialised epilogues keep
//                 the accumulator row of a thread in registers (tcgen05.ld 32x32b: lane = row), apply bias / q-scale /
//                 GELU (+ stored derivative) / x aux / + fp32 residual there, write the row into a 128B-swizzled
//                 4 KB shared-memory box and hand it to a TMA bulk-tensor STORE (reduce-add for split-K / gradient
//                 accumulation); residual and aux boxes arrive by TMA load, prefetched one chunk ahead. ~3
//                 instructions per element, and every global access is a full 128-byte row segment.
//                 The generic epilogue (any runtime switch of oat_gemm_args, unaligned outputs) transposes through
//                 shared memory and uses plain vector loads/stores.
//
// Operand tiles are staged by TMA into 128B-swizzled shared memory, 4 stages of (128 x 64) + (BLOCK_N x 64) bf16.
// Either operand may be K-major (contraction dim contiguous in global memory: activations x weights^T) or
// MN-major (row dim contiguous: the weight-gradient contraction over tokens, and dX = dY . W without a
// transposed weight copy).  The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Split-K (for the weight gradients, whose output has < 148 tiles)
// reduces with fp32 atomics straight into the gradient buffer.
// CTA pairs (cluster of 2, cta_group::2) take their tiles from a per-launch atomic counter (dynamic tile scheduler, see
// the kernel body): a pair that becomes resident late - an NCCL reduction or a side-stream kernel held its SMs when the
// grid launched - owes one tile, not a whole static share.
// Epilogue EPI_ROWDOT (act 4): besides the bf16 output, the dot of every stored output row with an aux row per
// 64-column block - delta = rowsum(dO * O) per head for the attention backward, out of the accumulator row the epilogue
// thread already holds.
//
// This one kernel carries every dense contraction of the hot path (reference call sites:
// OATrans/model/video_transformer.py:102,133 (qkv/proj), :46-49 (Mlp), :69 (patch conv as GEMM),
// OATrans/model/oa_model.py:68-75 (projections), OATrans/model/model.py:164-172 (sim matrix),
// HF DistilBERT linears) and their autograd counterparts.
#include <cuda.h>
#include <stdlib.h>

#include <atomic>

#include "oat_host.h"
#include "oat_ptx.cuh"

namespace oat {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kChunk = 16;                       // accumulator columns per tcgen05.ld
constexpr int kStgPitch = 20;                    // floats; 80-byte rows keep STS.128 / LDS.128 conflict-free
constexpr int kStagingFloats = 32 * kStgPitch;

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, split_k, k_blocks;
  const float* bias;
  const float* residual;
  long long ldr;
  float* out_f32;
  long long ld_f32;
  __nv_bfloat16* out_bf16;
  long long ld_bf16;
  __nv_bfloat16* out2_bf16;
  long long ld2;
  const __nv_bfloat16* aux_bf16;
  long long ld_aux;
  int act;
  int scale_cols;
  float scale;
  float alpha;
  int accumulate;
  float* rowdot;      // EPI_ROWDOT: rowdot[(col / 64) * ld_rowdot + row] = sum over the 64-column block of bf16(out) * aux
  long long ld_rowdot;
  unsigned int* sched;  // CTA pairs: {next tile, pairs done} of this launch (dynamic tile scheduler); null = static stride
  int idle_wait;      // epilogue warps sleep between polls while the k-loop runs (experiment knob OAT_GEMM_IDLE_WAIT)
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// one 128-bit reduction instead of four scalar atomics (split-K / gradient accumulation epilogue)
__device__ __forceinline__ void red_add_v4(float* gptr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(gptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Epilogue specialisations (compile-time, so the per-element instruction count stays small - the K=768 GEMMs give the
// epilogue only ~6 k cycles per 128x256 tile): the generic one keeps every runtime switch of oat_gemm_args.
enum EpiMode { EPI_GENERIC = 0, EPI_BF16 = 1, EPI_F32_RES = 2, EPI_GELU = 3, EPI_MULAUX = 4, EPI_ATOMIC = 5, EPI_ROWDOT = 6 };

constexpr int kBoxBytes = 32 * 128;              // one epilogue box: 32 rows x 128 B (64 bf16 or 32 fp32 columns)

// TWO: CTA pair (cluster of 2, tcgen05 cta_group::2). The pair computes a 256 x BLOCK_N tile: each CTA stages its own
// 128 rows of A and HALF of the B tile, so the L2 -> shared-memory traffic per MMA drops from 48 KB to 32 KB per
// 128x256x64 step (the 1-CTA kernel is bound by exactly that traffic: ~7 KB/clk chip-wide at 1.2 PFLOP/s).
template <int BLOCK_N, int EPI, bool TWO>
struct GemmSmem {
  static constexpr bool kTmaEpi = EPI != EPI_GENERIC;
  // boxes per epilogue warp: output (+ second output for GELU') (+ residual / aux input)
  static constexpr int kBoxes = !kTmaEpi ? 0 : (EPI == EPI_BF16 || EPI == EPI_ATOMIC) ? 1 : 2;   // ROWDOT: output + aux input
  static constexpr int kEpiBytes = kTmaEpi ? kEpiWarps * kBoxes * kBoxBytes : kEpiWarps * kStagingFloats * 4;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kLoadN = TWO ? BLOCK_N / 2 : BLOCK_N;          // B rows staged by this CTA
  static constexpr int kBBytes = kLoadN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // the TMA ring takes what the epilogue boxes leave of the 227 KB
  static constexpr int kStagesFit = (232448 - 2048 - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kBarrierBytes = (2 * kStages + 4 + kEpiWarps + 4) * 8 + 16 + 16;   // + scheduler: 4 barriers, tile ring
  static constexpr int kTotal = kStages * kStageBytes + kEpiBytes + kBarrierBytes + 1024;  // +1024 align slack
};

__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float4 v) {
  uint2 raw;
  raw.x = pack_bf16x2(v.x, v.y);
  raw.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = raw;
}

// tmap_o: output box map (bf16 {64,32} or fp32 {32,32}); tmap_x: second output (GELU') or residual / aux input.
template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool TWO>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_x,
                 const GemmParams p) {
  using S = GemmSmem<BLOCK_N, EPI, TWO>;
  static_assert(!TWO || S::kTmaEpi, "CTA pairs are implemented for the TMA-box epilogues only");
  constexpr int kStages = S::kStages;
  constexpr int kPair = TWO ? 2 : 1;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // 256 or 512: power of two
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();      // the next kernel's prologue may overlap this grid's tail (it waits before touching memory)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                // [kStages][16 KB]
  uint8_t* smem_b = smem + kStages * S::kABytes;          // [kStages][kLoadN*128 B]
  uint8_t* epi_smem = smem + kStages * S::kStageBytes;    // boxes (TMA epilogues) or transpose staging (generic)
  float* staging = reinterpret_cast<float*>(epi_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + S::kEpiBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * kStages; // [2]        MMA -> epilogue
  uint64_t* tmem_empty_bar = tmem_full_bar + 2; // [2]        epilogue -> MMA
  uint64_t* in_bar = tmem_empty_bar + 2;        // [kEpiWarps] residual / aux box landed (TMA epilogues)
  uint64_t* sched_full = in_bar + kEpiWarps;    // [2] scheduler -> roles: tile index of iteration it in ring[it & 1]
  uint64_t* sched_empty = sched_full + 2;       // [2] roles of BOTH CTAs -> scheduler (the leader's pair is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_empty + 2);
  volatile int* tile_ring = reinterpret_cast<volatile int*>(tmem_slot + 4);   // [2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_a);
      tma_prefetch_desc(&tmap_b);
      if constexpr (S::kTmaEpi) tma_prefetch_desc(&tmap_o);
      if constexpr (S::kTmaEpi && S::kBoxes == 2) tma_prefetch_desc(&tmap_x);
    }
    __syncwarp();
    if constexpr (TWO) tmem_alloc_pair<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], kPair);             // pair: the producers of both CTAs arrive on the leader's barrier
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kPair * kEpiWarps);   // pair: the epilogue warps of both CTAs release the leader
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&in_bar[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sched_full[i], 1);                         // leader: the scheduler's arrive; peer: its producer arms 4 tx bytes
      mbar_init(&sched_empty[i], 2 * (1 + kEpiWarps));      // MMA warp + epilogue warps (leader), producer + epilogue warps (peer)
    }
    fence_mbar_init();
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all();          // barrier inits of both CTAs visible before any remote arrive
  else __syncthreads();
  tc_fence_after();
  pdl_wait();                   // prologue done; from here on the kernel reads what earlier kernels in the stream wrote
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;
  const int worker = TWO ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int num_workers = TWO ? static_cast<int>(cluster_count_x()) : static_cast<int>(gridDim.x);

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int num_tiles = tiles_mn * p.split_k;
  const int kb_per_split = (p.k_blocks + p.split_k - 1) / p.split_k;

  // ---- tile scheduler. Static: iteration `it` of worker w is tile w + it * num_workers. Dynamic (CTA pairs, p.sched): the
  // leader's producer thread draws tiles from an atomic counter one iteration ahead and publishes them through a two-slot
  // ring in both CTAs (local store + arrive; st.async + complete_tx into the peer), every other role reads the slot and
  // releases it on the leader's sched_empty barrier. Pairs that become resident late (another kernel - an NCCL reduction,
  // a side-stream GEMM - holds their SMs when the grid launches) then find the counter exhausted and exit instead of
  // running a full static share after everybody else has finished.
  // The first tile of a pair stays static (tile = pair index: no counter round trip in front of the first load); a late
  // pair then still owes exactly that one tile.
  const bool dyn = TWO && p.sched != nullptr;
  auto consumer_tile = [&](int it) -> int {
    if (!dyn || it == 0) {
      const long long t = static_cast<long long>(worker) + static_cast<long long>(it) * num_workers;
      return t < num_tiles ? static_cast<int>(t) : -1;
    }
    const int j = it - 1;                      // ring use j serves iteration j + 1
    mbar_wait(&sched_full[j & 1], (j >> 1) & 1);
    return tile_ring[j & 1];
  };
  auto consumer_release = [&](int it) {       // ONE thread per role, after every thread of the role has read the slot
    if (!dyn || it == 0) return;
    const int j = it - 1;
    if (cta_rank == 0) mbar_arrive(&sched_empty[j & 1]);
    else mbar_arrive_cluster(mapa_u32(smem_u32(&sched_empty[j & 1]), 0u));
  };
  auto peer_arm = [&](int it) {               // peer's producer thread: its sched_full expects the leader's 4-byte st.async
    if (dyn && it > 0) mbar_arrive_expect_tx(&sched_full[(it - 1) & 1], 4);
  };
  auto draw_tile = [&]() -> int {
    const unsigned int t = atomicAdd(p.sched, 1u) + static_cast<unsigned int>(num_workers);
    return t < static_cast<unsigned int>(num_tiles) ? static_cast<int>(t) : -1;
  };
  auto publish_tile = [&](int it, int t) {     // leader's producer thread only, it >= 1
    const int j = it - 1;
    const int slot = j & 1;
    mbar_wait(&sched_empty[slot], ((j >> 1) & 1) ^ 1);
    tile_ring[slot] = t;
    st_async_u32(mapa_u32(smem_u32(const_cast<int*>(&tile_ring[slot])), 1u), static_cast<uint32_t>(t),
                 mapa_u32(smem_u32(&sched_full[slot]), 1u));
    mbar_arrive(&sched_full[slot]);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const bool fetcher = dyn && cta_rank == 0;
      int tile = consumer_tile(0), next = -1;
      for (int it = 0; tile >= 0; ++it) {
        if (fetcher) next = draw_tile();        // in flight while this tile's loads are issued
        const int split = tile / tiles_mn;
        const int mn = tile - split * tiles_mn;
        const int m_blk = mn / p.num_n_blocks, n_blk = mn - m_blk * p.num_n_blocks;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * S::kABytes;
          uint8_t* sb = smem_b + stage * S::kBBytes;
          const int m_row = (m_blk * kPair + static_cast<int>(cta_rank)) * BLOCK_M;        // this CTA's 128 rows of A / D
          const int n_row = n_blk * BLOCK_N + static_cast<int>(cta_rank) * S::kLoadN;     // this CTA's part of the B tile
          if constexpr (!TWO) {
            mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_row);
            } else {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i)
                tma_load_2d(sa + i * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m_row + i * 64, kb * BLOCK_K);
            }
            if constexpr (!B_MN) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_row);
            } else {
#pragma unroll
              for (int i = 0; i < S::kLoadN / 64; ++i)
                tma_load_2d(sb + i * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n_row + i * 64, kb * BLOCK_K);
            }
          } else {
            // both CTAs count their bytes on the LEADER's barrier (the leader's MMA thread consumes both halves)
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0u);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            else mbar_arrive_cluster(lead_bar);
            if constexpr (!A_MN) {
              tma_load_2d_pair(sa, &tmap_a, lead_bar, kb * BLOCK_K, m_row);
            } else {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i)
                tma_load_2d_pair(sa + i * (BLOCK_K * 128), &tmap_a, lead_bar, m_row + i * 64, kb * BLOCK_K);
            }
            if constexpr (!B_MN) {
              tma_load_2d_pair(sb, &tmap_b, lead_bar, kb * BLOCK_K, n_row);
            } else {
#pragma unroll
              for (int i = 0; i < S::kLoadN / 64; ++i)
                tma_load_2d_pair(sb + i * (BLOCK_K * 128), &tmap_b, lead_bar, n_row + i * 64, kb * BLOCK_K);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (fetcher) {
          publish_tile(it + 1, next);
          tile = next;
        } else {
          peer_arm(it + 1);
          tile = consumer_tile(it + 1);
          consumer_release(it + 1);
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(kPair * BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    // K-major: 8-row groups 1024 B apart; k-step of 16 elements = 32 B inside the swizzle row.
    // MN-major: 8-k-row groups 1024 B apart, 64-wide MN atoms BLOCK_K*128 B apart; k-step of 16 rows = 2048 B.
    constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 0, b_lbo = B_MN ? BLOCK_K * 128 : 0;
    constexpr uint32_t a_kstep = A_MN ? UMMA_K * 128 : UMMA_K * 2, b_kstep = B_MN ? UMMA_K * 128 : UMMA_K * 2;
    uint32_t stage = 0, phase = 0;
    for (int local_iter = 0;; ++local_iter) {
      const int tile = consumer_tile(local_iter);
      __syncwarp();
      if (lane == 0) consumer_release(local_iter);
      if (tile < 0) break;
      const int split = tile / tiles_mn;
      const int kb0 = split * kb_per_split;
      const int kb1 = min(p.k_blocks, kb0 + kb_per_split);
      const uint32_t acc = local_iter & 1;
      const uint32_t acc_phase = (local_iter >> 1) & 1;
      // Narrow last column block: when only a few columns of the tile exist (the "ones" column that turns a weight-
      // gradient GEMM into weight + bias gradient, engine.py) the instruction's N shrinks to 16 instead of multiplying
      // BLOCK_N - n zero columns. Pair mode feeds instruction columns [0, N/2) from the leader and [N/2, N) from the
      // peer's (all out-of-range, zero) half tile, so it is only taken there when the valid columns fit in the first 8.
      uint32_t idesc_t = idesc;
      {
        const int mn = tile - split * tiles_mn;
        const int n_blk = mn % p.num_n_blocks;
        const int n_valid = p.N - n_blk * BLOCK_N;
        if (n_valid <= (TWO ? 8 : 16)) idesc_t = make_idesc_bf16(kPair * BLOCK_M, 16, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      }
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * S::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * S::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            if constexpr (TWO) tc_mma_bf16_pair(d_tmem, adesc, bdesc, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
            else tc_mma_bf16(d_tmem, adesc, bdesc, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if constexpr (TWO) {
            tc_commit_pair(&empty_bar[stage]);                       // frees the slot in both CTAs
            if (kb == kb1 - 1) tc_commit_pair(&tmem_full_bar[acc]);  // both epilogues may start
          } else {
            tc_commit(&empty_bar[stage]);                       // smem slot free once these MMAs retire
            if (kb == kb1 - 1) tc_commit(&tmem_full_bar[acc]);  // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // pair, non-leader CTA: the leader issues the MMAs for both
  } else if constexpr (S::kTmaEpi) {
    // ------------------------------------------------------------------ epilogue warps, TMA boxes
    // lane = accumulator row (tcgen05.ld 32x32b), CW consecutive columns per chunk in registers. The row is written
    // into a 128B-swizzled box (16-byte unit j of row r at unit j ^ (r & 7): conflict-free per quarter-warp and
    // exactly the layout SWIZZLE_128B tensor maps expect) and stored / reduce-added by one TMA instruction.
    const int q = warp & 3;                 // TMEM lane quarter this warp may address
    const int half = (warp - 2) >> 2;       // which of the two interleaved chunk streams of that quarter
    constexpr bool kF32Out = EPI == EPI_F32_RES || EPI == EPI_ATOMIC;
    constexpr int CW = kF32Out ? 32 : 64;   // columns per 128-byte box row
    constexpr int kChunksT = BLOCK_N / CW;
    constexpr bool kHasIn = EPI == EPI_F32_RES || EPI == EPI_MULAUX || EPI == EPI_ROWDOT;
    uint8_t* box_o = epi_smem + (warp - 2) * (S::kBoxes * kBoxBytes);
    uint8_t* box_x = box_o + kBoxBytes;     // second output (GELU') or input (residual / aux); only if kBoxes == 2
    const uint32_t row_sw = static_cast<uint32_t>(lane & 7) << 4;
    const uint32_t my_o = smem_u32(box_o) + lane * 128;
    const uint32_t my_x = smem_u32(box_x) + lane * 128;
    uint64_t* my_in_bar = &in_bar[warp - 2];
    uint32_t in_count = 0;                  // input boxes consumed so far (mbarrier parity)
    for (int local_iter = 0;; ++local_iter) {
      const int tile = consumer_tile(local_iter);
      __syncwarp();
      if (lane == 0) consumer_release(local_iter);
      if (tile < 0) break;
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks, n_blk = mn - m_blk * p.num_n_blocks;
      const uint32_t acc = local_iter & 1;
      const uint32_t acc_phase = (local_iter >> 1) & 1;
      const int row0 = (m_blk * kPair + static_cast<int>(cta_rank)) * BLOCK_M + q * 32;
      const int col0 = n_blk * BLOCK_N;
      const bool first_split = (split == 0);
      const bool use_bias = (EPI == EPI_BF16 || EPI == EPI_F32_RES || EPI == EPI_GELU || EPI == EPI_ROWDOT) && p.bias != nullptr && first_split;
      const bool use_in = kHasIn && (EPI == EPI_MULAUX || EPI == EPI_ROWDOT || (p.residual != nullptr && first_split));
      if (use_in && lane == 0) {              // first input box of the tile: in flight while the MMAs finish
        mbar_arrive_expect_tx(my_in_bar, kBoxBytes);
        tma_load_2d(box_x, &tmap_x, my_in_bar, col0 + half * CW, row0);
      }
      if (p.idle_wait) mbar_wait_idle(&tmem_full_bar[acc], acc_phase);
      else mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < kChunksT; c += 2) {
        const bool last = (c + 2 >= kChunksT);
        const int col = col0 + c * CW;
        uint32_t v[CW];
        {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + c * CW;
          tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          if constexpr (CW == 64) tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
        }
        if (last) {
          // this warp has read everything it needs from the accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (TWO) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0u));
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        uint4 xin[8];
        if constexpr (kHasIn) {
          if (use_in) {
            mbar_wait(my_in_bar, in_count & 1);
            ++in_count;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = lds128(my_x + ((static_cast<uint32_t>(j) << 4) ^ row_sw));
              xin[j] = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
            }
            __syncwarp();                   // every lane has its row: the box may be refilled
            if (!last && lane == 0) {
              mbar_arrive_expect_tx(my_in_bar, kBoxBytes);
              tma_load_2d(box_x, &tmap_x, my_in_bar, col + 2 * CW, row0);
            }
          }
        }
        // ---- element-wise part, in registers
        if (use_bias) {
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
            v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b4.x);
            v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b4.y);
            v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b4.z);
            v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b4.w);
          }
        }
        if constexpr (EPI == EPI_BF16) {
          if (col < p.scale_cols) {         // q columns (scale_cols is a multiple of 4; usually of CW too)
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (col + j < p.scale_cols) v[j] = __float_as_uint(__uint_as_float(v[j]) * p.scale);
          }
        }
        if constexpr (EPI == EPI_F32_RES) {
          if (use_in) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + __uint_as_float(xin[j].x));
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + __uint_as_float(xin[j].y));
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + __uint_as_float(xin[j].z));
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + __uint_as_float(xin[j].w));
            }
          }
        }
        if constexpr (EPI == EPI_MULAUX) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t w[4] = {xin[j].x, xin[j].y, xin[j].z, xin[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              v[8 * j + 2 * k] = __float_as_uint(__uint_as_float(v[8 * j + 2 * k]) * __uint_as_float(w[k] << 16));
              v[8 * j + 2 * k + 1] = __float_as_uint(__uint_as_float(v[8 * j + 2 * k + 1]) * __uint_as_float(w[k] & 0xffff0000u));
            }
          }
        }
        // ---- the previous TMA store out of this warp's box(es) must have finished reading them
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        if constexpr (kF32Out) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(my_o + ((static_cast<uint32_t>(j) << 4) ^ row_sw), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else if constexpr (EPI == EPI_GELU) {
          // GELU(erf) and its derivative from one shared exponential; the derivative (bf16) is what backward needs
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g[8], d[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2)
              gelu_fwd_grad2(__uint_as_float(v[8 * j + k]), __uint_as_float(v[8 * j + k + 1]), g[k], g[k + 1], d[k], d[k + 1]);
            const uint32_t off = (static_cast<uint32_t>(j) << 4) ^ row_sw;
            sts128(my_o + off, pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
            sts128(my_x + off, pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
          }
        } else if constexpr (EPI == EPI_ROWDOT) {
          // the output row as stored (bf16) dotted with the aux row over this 64-column block: delta = dO . O of the
          // attention backward (one head per block), so that kernel need not read O at all
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t o4[4] = {pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])),
                                    pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                    pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                    pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]))};
            const uint32_t w[4] = {xin[j].x, xin[j].y, xin[j].z, xin[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              dot = fmaf(__uint_as_float(o4[k] << 16), __uint_as_float(w[k] << 16), dot);
              dot = fmaf(__uint_as_float(o4[k] & 0xffff0000u), __uint_as_float(w[k] & 0xffff0000u), dot);
            }
            sts128(my_o + ((static_cast<uint32_t>(j) << 4) ^ row_sw), o4[0], o4[1], o4[2], o4[3]);
          }
          if (row0 + lane < p.M) p.rowdot[static_cast<long long>(col >> 6) * p.ld_rowdot + row0 + lane] = dot;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(my_o + ((static_cast<uint32_t>(j) << 4) ^ row_sw),
                   pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])),
                   pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                   pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                   pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
        }
        fence_proxy_async_smem();           // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if constexpr (EPI == EPI_ATOMIC) tma_reduce_add_2d(&tmap_o, box_o, col, row0);
          else tma_store_2d(&tmap_o, box_o, col, row0);
          if constexpr (EPI == EPI_GELU) tma_store_2d(&tmap_x, box_x, col, row0);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait<0>();          // the boxes must outlive the last store's reads
  } else {
    // ------------------------------------------------------------------ epilogue warps, generic
    const int q = warp & 3;                 // TMEM lane quarter this warp may address
    const int half = (warp - 2) >> 2;       // which of the two interleaved chunk streams of that quarter
    float* stg = staging + (warp - 2) * kStagingFloats;
    const uint32_t stg_w = smem_u32(stg) + lane * (kStgPitch * 4);                      // this lane's row (write side)
    const uint32_t stg_r = smem_u32(stg) + ((lane & 7) * kStgPitch + (lane >> 3) * 4) * 4;  // read side, +it*8 rows
    const int rr = lane & 7;                // row within an 8-row group handled per read-back iteration
    const int cc = (lane >> 3) * 4;         // 4-column group within the 16-column chunk
    constexpr int kChunks = BLOCK_N / kChunk;
    constexpr bool kGeneric = EPI == EPI_GENERIC;
    for (int local_iter = 0;; ++local_iter) {
      const int tile = consumer_tile(local_iter);
      if (tile < 0) break;
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks, n_blk = mn - m_blk * p.num_n_blocks;
      const uint32_t acc = local_iter & 1;
      const uint32_t acc_phase = (local_iter >> 1) & 1;
      const int row_base = m_blk * BLOCK_M + q * 32 + rr;
      const int col_lane = n_blk * BLOCK_N + cc;          // + c * kChunk
      const bool first_split = (split == 0);
      const bool use_bias = p.bias != nullptr && first_split;
      const bool use_res = (kGeneric || EPI == EPI_F32_RES) && p.residual != nullptr && first_split;
      const bool use_aux = EPI == EPI_MULAUX || (kGeneric && p.act == 2);
      // per-row element offsets of the 4 rows this lane finishes per chunk (row = row_base + it*8)
      bool row_ok[4];
      long long off_f32[4], off_bf16[4], off_res[4], off_x2[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const long long row = row_base + it * 8;
        row_ok[it] = row < p.M;
        off_f32[it] = row * p.ld_f32 + col_lane;
        off_bf16[it] = row * p.ld_bf16 + col_lane;
        off_res[it] = row * p.ldr + col_lane;
        off_x2[it] = (EPI == EPI_GELU || (kGeneric && p.act == 1)) ? row * p.ld2 + col_lane : row * p.ld_aux + col_lane;
      }

      float4 res_cur[4], res_nxt[4];
      uint2 aux_cur[4], aux_nxt[4];
      auto prefetch = [&](int c, float4 (&res)[4], uint2 (&aux)[4]) {
        const bool col_ok = col_lane + c * kChunk < p.N;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const bool ok = row_ok[it] && col_ok;
          if (use_res) res[it] = ok ? *reinterpret_cast<const float4*>(p.residual + off_res[it] + c * kChunk)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          if (use_aux) aux[it] = ok ? *reinterpret_cast<const uint2*>(p.aux_bf16 + off_x2[it] + c * kChunk)
                                    : make_uint2(0u, 0u);
        }
      };
      prefetch(half, res_cur, aux_cur);          // global loads may start before the accumulator is ready
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < kChunks; c += 2) {
        const bool last = (c + 2 >= kChunks);
        if (!last) prefetch(c + 2, res_nxt, aux_nxt);
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + c * kChunk, v);
        tmem_ld_wait();
        if (last) {
          // this warp has read everything it needs from the accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(stg_w + j * 16, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col = col_lane + c * kChunk;
        if (col < p.N) {
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI != EPI_ATOMIC && EPI != EPI_MULAUX && use_bias) bias4 = *reinterpret_cast<const float4*>(p.bias + col);
          const bool scaled = (kGeneric || EPI == EPI_BF16) && col < p.scale_cols;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!row_ok[it]) continue;
            float4 x = lds128(stg_r + it * (8 * kStgPitch * 4));
            if (kGeneric) { x.x *= p.alpha; x.y *= p.alpha; x.z *= p.alpha; x.w *= p.alpha; }
            x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
            if (scaled) { x.x *= p.scale; x.y *= p.scale; x.z *= p.scale; x.w *= p.scale; }
            if (EPI == EPI_GELU || (kGeneric && p.act == 1)) {
              // GELU(erf) and its derivative from one shared exponential; the derivative (bf16) is what backward needs
              float4 d;
              gelu_fwd_grad(x.x, x.x, d.x); gelu_fwd_grad(x.y, x.y, d.y);
              gelu_fwd_grad(x.z, x.z, d.z); gelu_fwd_grad(x.w, x.w, d.w);
              st_bf16x4(p.out2_bf16 + off_x2[it] + c * kChunk, d);
            } else if (use_aux) {
              const float2 a0 = unpack_bf16x2(aux_cur[it].x), a1 = unpack_bf16x2(aux_cur[it].y);
              x.x *= a0.x; x.y *= a0.y; x.z *= a1.x; x.w *= a1.y;
            } else if (kGeneric && p.act == 3) {
              x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
            }
            if (use_res) { x.x += res_cur[it].x; x.y += res_cur[it].y; x.z += res_cur[it].z; x.w += res_cur[it].w; }
            if (EPI == EPI_ATOMIC || (kGeneric && p.accumulate && p.out_f32 != nullptr)) {
              red_add_v4(p.out_f32 + off_f32[it] + c * kChunk, x);
            } else if (EPI == EPI_F32_RES || (kGeneric && p.out_f32 != nullptr)) {
              *reinterpret_cast<float4*>(p.out_f32 + off_f32[it] + c * kChunk) = x;
            }
            if (EPI == EPI_BF16 || EPI == EPI_GELU || EPI == EPI_MULAUX || (kGeneric && p.out_bf16 != nullptr))
              st_bf16x4(p.out_bf16 + off_bf16[it] + c * kChunk, x);
          }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) { res_cur[it] = res_nxt[it]; aux_cur[it] = aux_nxt[it]; }
      }
    }
  }

  tc_fence_before();
  if constexpr (TWO) cluster_sync_all();          // the peer may still be reading operands / signalling barriers here
  else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_pair<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
  if constexpr (TWO) {
    // the last pair out re-arms the counter pair for a later launch (no memset between launches)
    if (dyn && cta_rank == 0 && threadIdx.x == 0) {
      if (atomicAdd(p.sched + 1, 1u) == static_cast<unsigned int>(num_workers) - 1u) {
        atomicExch(p.sched, 0u);
        atomicExch(p.sched + 1, 0u);
      }
    }
  }
}

// Counter pairs {next tile, pairs done} of the dynamic tile scheduler: every CTA-pair launch takes the next pair of the pool
// (launches in flight on different streams never share one) and its last pair out re-arms it, so no memset is needed.
constexpr int kGemmSchedSlots = 256;
__device__ unsigned int g_gemm_sched[kGemmSchedSlots][2];

// ---------------------------------------------------------------------------------------------- host side
static unsigned int* next_gemm_sched_slot() {
  static const bool on = [] { const char* e = getenv("OAT_GEMM_DYNAMIC"); return e == nullptr || atoi(e) != 0; }();
  if (!on) return nullptr;
  static unsigned int* base = [] {
    void* ptr = nullptr;
    return cudaGetSymbolAddress(&ptr, g_gemm_sched) == cudaSuccess ? static_cast<unsigned int*>(ptr) : nullptr;
  }();
  static std::atomic<unsigned int> seq{0};
  return base == nullptr ? nullptr : base + 2 * (seq.fetch_add(1) % kGemmSchedSlots);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D tensor, dim0 contiguous (length d0), dim1 with stride ld elements; box = {128 B of dim0, box1}; 128B swizzle.
static int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box1,
                        bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(OAT_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const uint64_t esz = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * esz) % 16 != 0)
    return set_error(OAT_ERR_ARG, "TMA operand needs 16-byte aligned base and row pitch (ptr=%p ld=%llu)", ptr,
                     (unsigned long long)ld);
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(OAT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return OAT_OK;
}
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box1) {
  return make_tmap_2d(out, ptr, d0, d1, ld, box1, false);
}

static int pick_split_k(int tiles_mn, int k_blocks, int sms, int requested) {
  if (requested > 0) return requested < k_blocks ? requested : (k_blocks > 0 ? k_blocks : 1);
  if (tiles_mn >= sms || k_blocks < 16) return 1;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 64 && s * 8 <= k_blocks; ++s) {
    // every split must own at least one k-block
    const int per = (k_blocks + s - 1) / s;
    if ((s - 1) * per >= k_blocks) continue;
    const long long t = 1LL * tiles_mn * s;
    const long long waves = (t + sms - 1) / sms;
    const double eff = double(t) / double(waves * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  return best;
}

static bool tma_ok(const void* ptr, long long ld, int esz) {
  return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * esz) % 16 == 0;
}

// The specialised (TMA-box) epilogues need 64-column granularity and 16-byte aligned tensors; anything else is generic.
static int pick_epi(const oat_gemm_args* a) {
  // the reduce-add boxes of the accumulate epilogue are clipped by the TMA unit, so a ragged N (the ones column of the
  // weight + bias gradient GEMM: N = K_in + 16) stays on the fast path; the other epilogues read bias / aux in whole boxes
  const bool ragged_ok = a->accumulate && a->N % 4 == 0;
  if (a->alpha != 1.0f || (a->N % 64 != 0 && !ragged_ok)) return EPI_GENERIC;
  if (a->bias != nullptr && (reinterpret_cast<uintptr_t>(a->bias) & 15) != 0) return EPI_GENERIC;
  const bool f32 = a->out_f32 != nullptr, b16 = a->out_bf16 != nullptr;
  const bool f32_ok = tma_ok(a->out_f32, a->ld_f32, 4), b16_ok = tma_ok(a->out_bf16, a->ld_bf16, 2);
  if (a->accumulate)
    return (f32_ok && !b16 && a->act == 0 && a->bias == nullptr && a->residual == nullptr && a->scale_cols == 0) ? EPI_ATOMIC : EPI_GENERIC;
  if (a->act == 1)
    return (b16_ok && !f32 && a->residual == nullptr && a->scale_cols == 0 && tma_ok(a->out2_bf16, a->ld2, 2)) ? EPI_GELU : EPI_GENERIC;
  if (a->act == 2)
    return (b16_ok && !f32 && a->residual == nullptr && a->bias == nullptr && a->scale_cols == 0 && tma_ok(a->aux_bf16, a->ld_aux, 2)) ? EPI_MULAUX : EPI_GENERIC;
  if (a->act == 4)      // validated by the caller (oat_gemm_bf16): TMA-able bf16 output and aux, no fp32 output / residual / scale
    return EPI_ROWDOT;
  if (a->act != 0) return EPI_GENERIC;
  if (b16_ok && !f32 && a->residual == nullptr) return EPI_BF16;
  if (f32_ok && !b16 && a->scale_cols == 0 && (a->residual == nullptr || tma_ok(a->residual, a->ldr, 4))) return EPI_F32_RES;
  return EPI_GENERIC;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool TWO>
static int launch_gemm(const oat_gemm_args* a, cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N, EPI, TWO>;
  constexpr int kPair = TWO ? 2 : 1;
  CUtensorMap ta, tb, to, tx;
  int rc;
  if (!A_MN) rc = make_tmap_bf16_2d(&ta, a->A, a->K, a->M, a->lda, BLOCK_M);
  else rc = make_tmap_bf16_2d(&ta, a->A, a->M, a->K, a->lda, BLOCK_K);
  if (rc != OAT_OK) return rc;
  if (!B_MN) rc = make_tmap_bf16_2d(&tb, a->B, a->K, a->N, a->ldb, S::kLoadN);
  else rc = make_tmap_bf16_2d(&tb, a->B, a->N, a->K, a->ldb, BLOCK_K);
  if (rc != OAT_OK) return rc;
  to = ta;
  tx = ta;
  if (EPI == EPI_BF16 || EPI == EPI_GELU || EPI == EPI_MULAUX || EPI == EPI_ROWDOT) rc = make_tmap_2d(&to, a->out_bf16, a->N, a->M, a->ld_bf16, 32, false);
  if (EPI == EPI_F32_RES || EPI == EPI_ATOMIC) rc = make_tmap_2d(&to, a->out_f32, a->N, a->M, a->ld_f32, 32, true);
  if (rc != OAT_OK) return rc;
  if (EPI == EPI_GELU) rc = make_tmap_2d(&tx, a->out2_bf16, a->N, a->M, a->ld2, 32, false);
  if (EPI == EPI_MULAUX || EPI == EPI_ROWDOT) rc = make_tmap_2d(&tx, a->aux_bf16, a->N, a->M, a->ld_aux, 32, false);
  if (EPI == EPI_F32_RES && a->residual != nullptr) rc = make_tmap_2d(&tx, a->residual, a->N, a->M, a->ldr, 32, true);
  if (rc != OAT_OK) return rc;

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_m_blocks = (a->M + kPair * BLOCK_M - 1) / (kPair * BLOCK_M);
  p.num_n_blocks = (a->N + BLOCK_N - 1) / BLOCK_N;
  p.k_blocks = (a->K + BLOCK_K - 1) / BLOCK_K;
  const int workers = num_sms() / kPair;          // CTAs, or CTA pairs
  const bool can_split = a->accumulate && a->out_f32 != nullptr && a->out_bf16 == nullptr && a->act == 0;
  // a narrow last column block (N = 16 instruction, see the MMA issuer) costs ~1/16 of a tile: balance the split on the
  // full-width tiles only
  const int n_tail = a->N % BLOCK_N;
  const bool narrow_last = n_tail != 0 && n_tail <= (TWO ? 8 : 16) && p.num_n_blocks > 1;
  const int tiles_for_split = p.num_m_blocks * (p.num_n_blocks - (narrow_last ? 1 : 0));
  p.split_k = (a->split_k == 0 && !can_split) ? 1 : pick_split_k(tiles_for_split, p.k_blocks, workers, a->split_k);
  {  // no empty splits
    const int per = (p.k_blocks + p.split_k - 1) / p.split_k;
    p.split_k = (p.k_blocks + per - 1) / per;
  }
  if (p.split_k > 1 && !can_split)
    return set_error(OAT_ERR_ARG, "split-K needs accumulate=1 into an fp32 output and no activation");
  p.bias = a->bias; p.residual = a->residual; p.ldr = a->ldr;
  p.out_f32 = a->out_f32; p.ld_f32 = a->ld_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16); p.ld_bf16 = a->ld_bf16;
  p.out2_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out2_bf16); p.ld2 = a->ld2;
  p.aux_bf16 = reinterpret_cast<const __nv_bfloat16*>(a->aux_bf16); p.ld_aux = a->ld_aux;
  p.act = a->act; p.scale_cols = a->scale_cols; p.scale = a->scale; p.alpha = a->alpha;
  p.accumulate = a->accumulate;
  p.rowdot = a->rowdot; p.ld_rowdot = a->ld_rowdot;
  p.sched = TWO ? next_gemm_sched_slot() : nullptr;
  {
    const char* e = getenv("OAT_GEMM_IDLE_WAIT");
    p.idle_wait = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }

  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, EPI, TWO>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", S::kTotal, cudaGetErrorString(e));
    attr_set = true;
  }
  const long long tiles = 1LL * p.num_m_blocks * p.num_n_blocks * p.split_k;
  const int nwork = static_cast<int>(tiles < workers ? tiles : workers);
  if constexpr (!TWO) {
    cudaError_t e = launch_pdl(kern, dim3(nwork), dim3(kGemmThreads), S::kTotal, stream, ta, tb, to, tx, p);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "gemm_bf16_kernel launch: %s", cudaGetErrorString(e));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nwork, 1, 1);
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tx, p);
    if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "gemm_bf16_kernel (CTA pair) launch: %s", cudaGetErrorString(e));
  }
  return check_launch("gemm_bf16_kernel");
}

// CTA pairs for every problem with more than one 128-row block; OAT_GEMM_2CTA=0 keeps every launch on single CTAs.
static bool use_pairs(const oat_gemm_args* a, int epi) {
  const char* env = getenv("OAT_GEMM_2CTA");
  const bool enabled = env == nullptr || atoi(env) != 0;
  return enabled && epi != EPI_GENERIC && a->M > BLOCK_M;
}

bool skinny_gemm_supported(const oat_gemm_args* a);          // gemm_skinny.cu: <= 64 rows, K-major x K-major
int launch_skinny_gemm(const oat_gemm_args* a, cudaStream_t stream);

}  // namespace oat

extern "C" int oat_gemm_bf16(const oat_gemm_args* a, oat_stream_t stream) {
  using namespace oat;
  OAT_REQUIRE(a != nullptr, "oat_gemm_bf16: null args");
  OAT_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "oat_gemm_bf16: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  OAT_REQUIRE(a->N % 4 == 0, "oat_gemm_bf16: N=%d must be a multiple of 4", a->N);
  OAT_REQUIRE(a->out_f32 != nullptr || a->out_bf16 != nullptr, "oat_gemm_bf16: no output");
  OAT_REQUIRE(a->act != 1 || a->out2_bf16 != nullptr, "oat_gemm_bf16: GELU forward needs out2_bf16");
  OAT_REQUIRE(a->act != 2 || a->aux_bf16 != nullptr, "oat_gemm_bf16: GELU backward needs aux_bf16");
  OAT_REQUIRE(a->scale_cols % 4 == 0, "oat_gemm_bf16: scale_cols must be a multiple of 4");
  if (a->act == 4) {
    OAT_REQUIRE(a->rowdot != nullptr && a->aux_bf16 != nullptr && a->out_bf16 != nullptr && a->out_f32 == nullptr &&
                    a->residual == nullptr && a->scale_cols == 0 && a->accumulate == 0 && a->alpha == 1.0f,
                "oat_gemm_bf16: act 4 (row dots) needs rowdot, aux_bf16 and a bf16 output only");
    OAT_REQUIRE(a->N % 256 == 0 && (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 && (a->ld_bf16 * 2) % 16 == 0 &&
                    (reinterpret_cast<uintptr_t>(a->aux_bf16) & 15) == 0 && (a->ld_aux * 2) % 16 == 0 &&
                    (a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0),
                "oat_gemm_bf16: act 4 needs N %% 256 == 0 and 16-byte aligned out_bf16 / aux_bf16 rows");
    OAT_REQUIRE(a->ld_rowdot >= a->M, "oat_gemm_bf16: ld_rowdot must be >= M");
  }
  cudaStream_t s = as_stream(stream);
  const bool amn = a->a_major != 0, bmn = a->b_major != 0;
  {
    // few-row products (CLS rows, projections) are weight streams: one CTA per 16 output columns, no split-K atomics
    static const bool skinny_on = [] { const char* e = getenv("OAT_GEMM_SKINNY"); return e == nullptr || atoi(e) != 0; }();
    if (skinny_on && skinny_gemm_supported(a)) return launch_skinny_gemm(a, s);
  }
  // 256-wide tiles unless the problem is narrow
  const bool wide = (a->N % 256 == 0) || a->N > 512;
  const int code = (amn ? 2 : 0) | (bmn ? 1 : 0);
  if (!wide) {
    switch (code) {
      case 0: return launch_gemm<128, false, false, EPI_GENERIC, false>(a, s);
      case 1: return launch_gemm<128, false, true, EPI_GENERIC, false>(a, s);
      case 2: return launch_gemm<128, true, false, EPI_GENERIC, false>(a, s);
      default: return launch_gemm<128, true, true, EPI_GENERIC, false>(a, s);
    }
  }
  const int epi = pick_epi(a);
  const bool two = use_pairs(a, epi);
#define OAT_GEMM_EPI(AM, BM, E) return two ? launch_gemm<256, AM, BM, E, true>(a, s) : launch_gemm<256, AM, BM, E, false>(a, s)
#define OAT_GEMM_CASE(AM, BM)                                                         \
  switch (epi) {                                                                      \
    case EPI_BF16: OAT_GEMM_EPI(AM, BM, EPI_BF16);                                    \
    case EPI_F32_RES: OAT_GEMM_EPI(AM, BM, EPI_F32_RES);                              \
    case EPI_GELU: OAT_GEMM_EPI(AM, BM, EPI_GELU);                                    \
    case EPI_MULAUX: OAT_GEMM_EPI(AM, BM, EPI_MULAUX);                                \
    case EPI_ATOMIC: OAT_GEMM_EPI(AM, BM, EPI_ATOMIC);                                \
    case EPI_ROWDOT: OAT_GEMM_EPI(AM, BM, EPI_ROWDOT);                                \
    default: return launch_gemm<256, AM, BM, EPI_GENERIC, false>(a, s);               \
  }
  switch (code) {
    case 0: OAT_GEMM_CASE(false, false)
    case 1: OAT_GEMM_CASE(false, true)
    case 2: OAT_GEMM_CASE(true, false)
    default: OAT_GEMM_CASE(true, true)
  }
#undef OAT_GEMM_CASE
#undef OAT_GEMM_EPI
}
