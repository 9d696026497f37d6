// Tensor-core issue-rate probe: back-to-back tcgen05.mma (bf16, M=128 per CTA, N columns, K=16) on garbage operands,
// single CTA vs CTA pair (cta_group::2), with an optional commit every `every` MMAs. Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I oa_transformer_b200/csrc -o build/mma_pair_probe scripts/probes/mma_pair_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "oat_ptx.cuh"
using namespace oat;

template <bool TWO>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, int N, int every, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;  // bf16 1.0
  fence_proxy_async_smem();
  if (warp == 0) {
    if constexpr (TWO) tmem_alloc_pair<512>(&slot); else tmem_alloc<512>(&slot);
  } else if (warp == 1 && lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(TWO ? 256 : 128, N, 0u, 0u);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t ad = make_smem_desc_sw128(a_addr + (i & 3) * 32, 0, 1024);
        const uint64_t bd = make_smem_desc_sw128(b_addr + (i & 3) * 32, 0, 1024);
        if constexpr (TWO) tc_mma_bf16_pair(tmem, ad, bd, idesc, i > 0 ? 1u : 0u);
        else tc_mma_bf16(tmem, ad, bd, idesc, i > 0 ? 1u : 0u);
        if (every > 0 && (i % every) == every - 1) {
          if constexpr (TWO) tc_commit_pair(&bars[1]); else tc_commit(&bars[1]);
        }
      }
      if constexpr (TWO) tc_commit_pair(&bars[0]); else tc_commit(&bars[0]);
      while (!mbar_try_wait(&bars[0], ph)) {}
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_pair<512>(tmem); else tmem_dealloc<512>(tmem);
  }
}

template <bool TWO>
static void run(int n_mma, int N, int every, int ctas) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 1024);
  cudaMemset(d, 0, sizeof(long long) * 1024);
  auto k = probe<TWO>;
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TWO ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, n_mma, N, every, d);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[1024];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("{\"pair\": %d, \"N\": %d, \"commit_every\": %d, \"ctas\": %d, \"cycles_per_mma\": %.1f}\n", TWO ? 1 : 0, N, every,
         ctas, double(mx) / n_mma);
  cudaFree(d);
}

int main() {
  const int n = 2048;
  for (int ctas : {2, 148}) {
    run<false>(n, 256, 0, ctas);
    run<false>(n, 256, 4, ctas);
    run<true>(n, 256, 0, ctas);
    run<true>(n, 256, 4, ctas);
    run<true>(n, 256, 16, ctas);
    run<true>(n, 128, 0, ctas);
    run<true>(n, 128, 4, ctas);
  }
  return 0;
}
