/* oat.h - C ABI of liboat.so: the sm_100a kernels behind OA-Transformer's video-text dual-encoder hot path.
 *
 * The reference (FingerRec/OA-Transformer) is pure PyTorch and has no FFI; this ABI is what its Python plugin
 * surface (OATrans/model/oa_model.py:FrozenInTime, OATrans/model/video_transformer.py, OATrans/model/loss.py,
 * OATrans/trainer/trainer_dist.py) binds through ctypes in oa_transformer_b200/_lib.py. Each entry point cites the
 * reference line(s) whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs. No torch / C++ types cross the boundary.
 *   - every pointer is DEVICE memory owned by the caller; the library never allocates, frees, or synchronises.
 *   - work is enqueued on the caller's stream (a cudaStream_t passed as void*); 0 = legacy default stream.
 *   - returns OAT_OK (0) or a negative code; oat_last_error() returns the thread-local message.
 *   - bf16 tensors are passed as void* (uint16 storage); "ld" arguments are leading dimensions in ELEMENTS.
 *   - there is no CPU fallback: on a non-sm_100 device every launch fails with OAT_ERR_CUDA.
 */
#ifndef OAT_H_
#define OAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OAT_OK 0
#define OAT_ERR_ARG (-1)
#define OAT_ERR_CUDA (-2)
#define OAT_ERR_ARCH (-3)

typedef void* oat_stream_t;

/* ---- library ---------------------------------------------------------------------------------------------- */
int oat_version(void);
const char* oat_last_error(void);
/* OAT_OK iff the current CUDA device is compute capability 10.x (B200); OAT_ERR_ARCH otherwise. */
int oat_device_check(void);

/* ---- dense contraction (tcgen05 / TMEM / TMA) ---------------------------------------------------------------
 * C[M,N] = alpha * A(M,K) . B(N,K)^T, bf16 operands, fp32 accumulation, fused epilogue.
 *   a_major = 0: A stored [M][lda], K contiguous.   a_major = 1: A stored [K][lda], M contiguous.
 *   b_major = 0: B stored [N][ldb], K contiguous.   b_major = 1: B stored [K][ldb], N contiguous.
 * epilogue order: alpha, +bias[N], first scale_cols columns *= scale, activation, +residual[M][ldr] (fp32), store.
 *   act 0: none | 1: GELU(erf) forward - the bf16 pre-activation goes to out2_bf16, GELU of that rounded value to
 *   the outputs | 2: multiply by GELU'(aux_bf16[M][ld_aux]) | 3: ReLU.
 *   accumulate = 1: atomically add into out_f32 (gradient accumulation / split-K). split_k = 0 lets the library pick.
 * Replaces: nn.Linear at video_transformer.py:102,133,46-49; Conv2d-as-GEMM :69; oa_model.py:68-75; torch.mm in
 * model/model.py:171; DistilBERT linears; and the autograd dgrad/wgrad of each. */
typedef struct oat_gemm_args {
  const void* A; int64_t lda; int32_t a_major;
  const void* B; int64_t ldb; int32_t b_major;
  int32_t M, N, K;
  float alpha;
  const float* bias;
  int32_t scale_cols; float scale;
  int32_t act;
  const void* aux_bf16; int64_t ld_aux;
  const float* residual; int64_t ldr;
  float* out_f32; int64_t ld_f32;
  void* out_bf16; int64_t ld_bf16;
  void* out2_bf16; int64_t ld2;
  int32_t accumulate;
  int32_t split_k;
} oat_gemm_args;
int oat_gemm_bf16(const oat_gemm_args* args, oat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OAT_H_ */
