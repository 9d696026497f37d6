// Probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b.x4 and tcgen05.st.16x128b.x4.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../oa_transformer_b200/csrc/oat_ptx.cuh"
using namespace oat;

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out_ld, uint32_t* out_st) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = slot;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  // pattern via 32x32b: lane l (global lane index), column c  <-  l * 1000 + c
  {
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = (warp * 32 + lane) * 1000 + c;
    tmem_st_32x32b_x16(t + lane_off, reinterpret_cast<uint32_t(&)[16]>(v[0]));
    tmem_st_32x32b_x16(t + lane_off + 16, reinterpret_cast<uint32_t(&)[16]>(v[16]));
    tmem_st_wait();
  }
  __syncwarp();
  // read back with 16x256b.x4 at lanes +0 and +16
  for (int kh = 0; kh < 2; ++kh) {
    uint32_t r[16];
    const uint32_t addr = t + lane_off + (static_cast<uint32_t>(kh * 16) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
    tmem_ld_wait();
    for (int k = 0; k < 16; ++k) out_ld[((warp * 2 + kh) * 32 + lane) * 16 + k] = r[k];
  }
  __syncwarp();
  // write with 16x128b.x4: value encodes (thread, reg); read back with 32x32b to see where each landed
  for (int kh = 0; kh < 2; ++kh) {
    uint32_t r[8];
    for (int k = 0; k < 8; ++k) r[k] = 100000 + kh * 10000 + lane * 100 + k;
    const uint32_t addr = t + lane_off + (static_cast<uint32_t>(kh * 16) << 16) + 64;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(addr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
  }
  tmem_st_wait();
  __syncwarp();
  {
    uint32_t v[16];
    tmem_ld_32x32b_x16(t + lane_off + 64, v);
    tmem_ld_wait();
    for (int c = 0; c < 16; ++c) out_st[(warp * 32 + lane) * 16 + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(t); }
}

int main() {
  uint32_t *d_ld, *d_st;
  cudaMalloc(&d_ld, 4 * 2 * 32 * 16 * 4);
  cudaMalloc(&d_st, 128 * 16 * 4);
  probe<<<1, 128>>>(d_ld, d_st);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t ld[4 * 2 * 32 * 16], st[128 * 16];
  cudaMemcpy(ld, d_ld, sizeof(ld), cudaMemcpyDeviceToHost);
  cudaMemcpy(st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
  // check the assumed mapping of the load: reg 4j + 2*rs + cs  <->  lane kh*16 + t/4 + 8*rs, column 8j + 2*(t%4) + cs
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int kh = 0; kh < 2; ++kh) for (int t = 0; t < 32; ++t) for (int k = 0; k < 16; ++k) {
    const int j = k >> 2, rs = (k >> 1) & 1, cs = k & 1;
    const uint32_t expect = (w * 32 + kh * 16 + t / 4 + 8 * rs) * 1000 + 8 * j + 2 * (t % 4) + cs;
    if (ld[((w * 2 + kh) * 32 + t) * 16 + k] != expect) ++bad;
  }
  printf("16x256b.x4 load: %d mismatches vs assumed mapping\n", bad);
  if (bad) for (int t = 0; t < 8; ++t) { printf("t%d:", t); for (int k = 0; k < 16; ++k) printf(" %u", ld[t * 16 + k]); printf("\n"); }
  // store: assumed reg 2j + rs  ->  lane kh*16 + t/4 + 8*rs, column 4j + t%4
  bad = 0;
  for (int l = 0; l < 128; ++l) for (int c = 0; c < 16; ++c) {
    const int li = l % 32, kh = li / 16, r16 = li % 16, rs = r16 / 8, tq = r16 % 8, j = c / 4, tm = c % 4;
    const uint32_t expect = 100000 + kh * 10000 + (tq * 4 + tm) * 100 + 2 * j + rs;
    if (st[l * 16 + c] != expect) ++bad;
  }
  printf("16x128b.x4 store: %d mismatches vs assumed mapping\n", bad);
  if (bad) for (int l = 0; l < 10; ++l) { printf("lane%d:", l); for (int c = 0; c < 16; ++c) printf(" %u", st[l * 16 + c]); printf("\n"); }
  return 0;
}
