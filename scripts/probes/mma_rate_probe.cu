// Probe: how fast does one thread get tcgen05.mma instructions of the attention kernels' shapes through the tensor pipe?
// For each variant: R repetitions of a block of MMAs, clock64 around the issue loop (issue time) and around
// issue + commit + mbarrier wait (completion time). Operands are whatever is in shared memory (zeros) - timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mma_rate scripts/probes/mma_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../oa_transformer_b200/csrc/oat_ptx.cuh"

using namespace oat;

constexpr int kSmem = 200 * 1024;

struct Res { long long issue, total; };

// variant: 0 SS N=128 K/K one chain | 1 SS N=128 two chains | 2 SS N=64 (A K-major, B MN-major) one chain
//          3 SS N=64 three chains | 4 TS N=64 one chain | 5 TS N=64 + SS N=64 + SS N=64(A MN-major) interleaved (bwd grads)
//          6 SS N=64 A MN-major one chain | 7 SS N=240 one chain (fwd S) | 8 TS N=64 chain of 15 (fwd PV)
//          9 SS N=64 K/K one chain | 10 SS N=256 K/K one chain
__global__ void __launch_bounds__(384, 1) probe(int variant, int reps, int noise, Res* out, int fill = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  for (int i = threadIdx.x; i < (kSmem - 2048) / 16; i += blockDim.x) {
    // fill 1: pseudo-random bf16 values in (-2, 2) (exponent bits 0x3f80 region), so the products are ordinary numbers
    uint32_t h = i * 2654435761u;
    auto w = [&](uint32_t x) { x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15; return (x & 0x807f807fu) | 0x3f003f00u; };
    reinterpret_cast<uint4*>(smem)[i] = fill ? make_uint4(w(h), w(h + 1), w(h + 2), w(h + 3)) : make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t a = smem_u32(smem), b = a + 32768, c = a + 65536, d = a + 98304;
    const uint32_t i128 = make_idesc_bf16(128, 128, 0, 0), i64kn = make_idesc_bf16(128, 64, 0, 1),
                   i64mn = make_idesc_bf16(128, 64, 1, 1), i240 = make_idesc_bf16(128, 240, 0, 0),
                   i64kk = make_idesc_bf16(128, 64, 0, 0), i256 = make_idesc_bf16(128, 256, 0, 0);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      switch (variant) {
        case 0:
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_bf16(t, make_smem_desc_sw128(a + (k & 3) * 32, 0, 1024), make_smem_desc_sw128(b + (k & 3) * 32, 0, 1024), i128, k > 0);
          break;
        case 1:
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_bf16(t, make_smem_desc_sw128(a + k * 32, 0, 1024), make_smem_desc_sw128(b + k * 32, 0, 1024), i128, k > 0);
            tc_mma_bf16(t + 128, make_smem_desc_sw128(c + k * 32, 0, 1024), make_smem_desc_sw128(d + k * 32, 0, 1024), i128, k > 0);
          }
          break;
        case 2:
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_bf16(t, make_smem_desc_sw128(a + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(b + k * 2048, 32768, 1024), i64kn, k > 0);
          break;
        case 3:
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            tc_mma_bf16(t, make_smem_desc_sw128(a + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(b + k * 2048, 32768, 1024), i64kn, k > 0);
            tc_mma_bf16(t + 64, make_smem_desc_sw128(a + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(c + k * 2048, 32768, 1024), i64kn, k > 0);
            tc_mma_bf16(t + 128, make_smem_desc_sw128(a + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(d + k * 2048, 32768, 1024), i64kn, k > 0);
          }
          break;
        case 4:
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_bf16_ts(t + 256, t + k * 8, make_smem_desc_sw128(b + k * 2048, 32768, 1024), i64kn, k > 0);
          break;
        case 5:
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            tc_mma_bf16_ts(t + 256, t + k * 8, make_smem_desc_sw128(b + k * 2048, 32768, 1024), i64kn, k > 0);
            tc_mma_bf16(t + 320, make_smem_desc_sw128(a + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(c + k * 2048, 32768, 1024), i64kn, k > 0);
            tc_mma_bf16(t + 384, make_smem_desc_sw128(a + k * 2048, 16384, 1024),
                        make_smem_desc_sw128(d + k * 2048, 32768, 1024), i64mn, k > 0);
          }
          break;
        case 6:
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_bf16(t + 384, make_smem_desc_sw128(a + k * 2048, 16384, 1024),
                        make_smem_desc_sw128(d + k * 2048, 32768, 1024), i64mn, k > 0);
          break;
        case 7:
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(t, make_smem_desc_sw128(a + k * 32, 0, 1024), make_smem_desc_sw128(b + k * 32, 0, 1024), i240, k > 0);
          break;
        case 8:
#pragma unroll
          for (int k = 0; k < 15; ++k)
            tc_mma_bf16_ts(t + 256, t + k * 8, make_smem_desc_sw128(b + k * 2048, 32768, 1024), i64kn, k > 0);
          break;
        case 9:
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_bf16(t, make_smem_desc_sw128(a + (k & 3) * 32, 0, 1024), make_smem_desc_sw128(b + (k & 3) * 32, 0, 1024), i64kk, k > 0);
          break;
        case 11: {
          // one group of the pipelined space-attention backward, exactly as its MMA warp issues it (160 instructions)
          const uint32_t Q = a, K = a + 32768, V = a + 65536, D = a + 98304, R = a + 131072;
          const uint32_t isd = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int vn = (v + 1) & 7, ktn = vn >> 2, qqn = vn & 3, kt = v >> 2, qq = v & 3;
            const uint32_t ts = t + (vn & 1) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(ts, make_smem_desc_sw128(K + ktn * 16384 + k * 32, 0, 1024), make_smem_desc_sw128(Q + qqn * 8192 + k * 32, 0, 1024), isd, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(ts + 64, make_smem_desc_sw128(V + ktn * 16384 + k * 32, 0, 1024), make_smem_desc_sw128(D + qqn * 8192 + k * 32, 0, 1024), isd, k > 0);
            const uint32_t tp = t + (v & 1) * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tc_mma_bf16_ts(t + 256, tp + (ks < 2 ? ks * 8 : 32 + (ks - 2) * 8), make_smem_desc_sw128(D + (qq * 64 + ks * 16) * 128, 32768, 1024), i64kn, (qq > 0 || ks > 0));
              tc_mma_bf16(t + 320, make_smem_desc_sw128(R + (v & 3) * 16384 + ks * 32, 0, 1024), make_smem_desc_sw128(Q + (qq * 64 + ks * 16) * 128, 32768, 1024), i64kn, (qq > 0 || ks > 0));
            }
            if (qq & 1) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc_mma_bf16(t + 384 + (qq >> 1) * 64, make_smem_desc_sw128(R + ((v - 1) & 3) * 16384 + ks * 2048, 16384, 1024),
                            make_smem_desc_sw128(K + (kt * 128 + ks * 16) * 128, 32768, 1024), i64mn, (kt > 0 || ks > 0));
            }
          }
        } break;
        case 10:
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_bf16(t, make_smem_desc_sw128(a + (k & 3) * 32, 0, 1024), make_smem_desc_sw128(b + (k & 3) * 32, 0, 1024), i256, k > 0);
          break;
      }
    }
    long long t1 = clock64();
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out->issue = t1 - t0;
    out->total = t2 - t0;
    stop = 1;
  } else if (threadIdx.x >= 128 && noise != 0) {
    // noise warps (8): 1 = tcgen05.ld of 64 columns in a loop, 2 = st.shared.v4 in a loop, 3 = ex2 / fma arithmetic,
    // 4 = tcgen05.ld + arithmetic + tcgen05.st + st.shared (the math warps' mix)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t taddr = t + (static_cast<uint32_t>((w & 3) * 32) << 16) + 128 + (w >> 2 & 1) * 64;
    uint8_t* my = smem + 131072 + (w - 4) * 2048 + lane * 64;
    float acc = 0.f;
    uint32_t it = 0;
    while (!stop && ++it < 200000) {
      if (noise == 1 || noise == 4) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc += __uint_as_float(v[e]);
        if (noise == 4) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float a, b;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(__uint_as_float(v[2 * e])));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(__uint_as_float(v[2 * e + 1])));
            pk[e] = pack_bf16x2(a, b);
          }
          tmem_st_32x32b_x16(taddr, pk);
          tmem_st_wait();
          *reinterpret_cast<uint4*>(my) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(my + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      } else if (noise == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) *reinterpret_cast<uint4*>(my + e * 16) = make_uint4(it, it, it, it);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc)); acc = fmaf(y, 0.5f, 0.25f); }
      }
    }
    if (acc == 1234.5f) out->issue = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(t); }
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  Res* d;
  cudaMalloc(&d, sizeof(Res));
  const char* names[] = {"SS N=128 K/K, 1 chain, 8 MMA", "SS N=128 K/K, 2 chains, 8 MMA", "SS N=64 K/MN, 1 chain, 8 MMA",
                         "SS N=64 K/MN, 3 chains, 24 MMA", "TS N=64, 1 chain, 8 MMA", "bwd grads TS+SS+SS(MN A), 24 MMA",
                         "SS N=64 MN/MN, 1 chain, 8 MMA", "SS N=240 K/K, 1 chain, 4 MMA", "TS N=64, 1 chain, 15 MMA",
                         "SS N=64 K/K, 1 chain, 8 MMA", "SS N=256 K/K, 1 chain, 8 MMA", "space bwd group, 160 MMA"};
  const int per[] = {8, 8, 8, 24, 8, 24, 8, 4, 15, 8, 8, 160};
  for (int v : {9, 5, 11}) {
    for (int noise = 0; noise < 10; noise += 1) {
      const int reps = 16;
      const int fill = noise >= 5;
      Res h;
      for (int it = 0; it < 2; ++it) {
        probe<<<1, 384, kSmem>>>(v, reps, noise % 5, d, fill);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-36s noise %d (+5 = random operands): issue %7lld clk (%.1f / MMA)   complete %7lld clk (%.1f / MMA)\n", names[v], noise, h.issue,
             double(h.issue) / (reps * per[v]), h.total, double(h.total) / (reps * per[v]));
    }
  }
  return 0;
}
