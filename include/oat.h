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
 *   act 0: none | 1: GELU(erf) forward - GELU of the fp32 accumulator goes to the outputs and its derivative
 *   GELU'(x) (bf16) to out2_bf16, which is exactly the aux operand of the backward GEMM | 2: multiply by aux_bf16[M][ld_aux] (the stored GELU') | 3: ReLU
 *   | 4: row dots - the bf16 output is unchanged and rowdot[(c / 64) * ld_rowdot + r] = sum over the 64-column block c / 64 of
 *   bf16(out[r][.]) * aux_bf16[r][.] (fp32). With out = dO (the dgrad of the attention output projection) and aux = O this is
 *   delta = rowsum(dO * O) per head of the softmax backward (what autograd derives for video_transformer.py:122-131), handed to
 *   oat_attn_bwd through oat_attn_args.delta so that the attention backward does not read O. Needs N % 256 == 0, a bf16
 *   output only, 16-byte aligned out_bf16 / aux_bf16 rows.
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
  float* rowdot; int64_t ld_rowdot;   /* act 4 only: [N / 64][ld_rowdot >= M] fp32 */
} oat_gemm_args;
int oat_gemm_bf16(const oat_gemm_args* args, oat_stream_t stream);

/* ---- LayerNorm over the embedding dimension (D in {128,256,512,768,1024}) ------------------------------------
 * Forward: y = (x - mean) * rstd * gamma + beta per row; writes bf16 (GEMM operand) and/or fp32 outputs and the
 * row statistics needed by backward. Replaces nn.LayerNorm(eps=1e-6) at video_transformer.py:164,167,174,346 and
 * DistilBERT's LayerNorm(eps=1e-12). Rows may be strided (ldx) so that only the CLS rows are normalised for :351.
 * Backward: dy = dy_bf16 (+ dy_f32); dx = add1 + add2 + LN'(dy); dgamma/dbeta are ACCUMULATED (atomics), and so is
 * dxsum[c] += sum_rows dx[row, c] (optional: the bias gradient of the GEMM that produced the LayerNorm input).
 * y_split (optional, bf16, pitch ldys >= 3*D): rows r with r % split_period == 0 are also written to
 * y_split[r / split_period] as the split-bf16 operand [hi | hi | lo] (hi = bf16(y), lo = bf16(y - hi)); see
 * oat_split3_bf16. With split_period = T these are the CLS rows of the video tower. */
int oat_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int64_t rows,
                      int32_t D, void* y_bf16, int64_t ldy, float* y_f32, int64_t ldyf, float* mean, float* rstd,
                      void* y_split, int64_t ldys, int64_t split_period, oat_stream_t stream);
int oat_layernorm_bwd(const void* dy_bf16, int64_t lddyb, const float* dy_f32, int64_t lddyf, const float* x,
                      int64_t ldx, const float* mean, const float* rstd, const float* gamma, int64_t rows, int32_t D,
                      const float* add1, const float* add2, int64_t ldadd, float* dx, int64_t lddx, void* dx_bf16,
                      int64_t lddxb, float* dgamma, float* dbeta, float* dxsum, oat_stream_t stream);

/* ---- attention ------------------------------------------------------------------------------------------------
 * qkv: bf16 [B*T, 3*H*64] as produced by the qkv GEMM (q already scaled by 64^-0.5, columns q | k | v, head-major).
 * mode 0 "space": T = 1 + F*n; token (f,i) attends to [CLS] + the n tokens of frame f.
 * mode 1 "time" : token (f,i) attends to [CLS] + tokens (f',i) of every frame f'.
 *   In both, the CLS query (token 0) attends to all T keys.           (video_transformer.py:99-135)
 * mode 2 "plain": every token attends to every key with key_mask[b*T+j] != 0 (DistilBERT self-attention).
 * out: bf16 [B*T, H*64]; lse: fp32 [B*H*T] log-sum-exp per query (saved for backward).
 * Backward consumes dout (bf16, same layout as out) and writes dqkv (bf16, same layout as qkv; the q part is the
 * gradient w.r.t. the UNSCALED projection, i.e. multiplied by `scale`). cls_acc: fp32 [B*H*3*64] scratch.
 * Forward, modes 0/1: cls_acc may instead point to oat_attn_fwd_workspace_floats() fp32 words of scratch; the CLS
 * query is then fused into the space / time kernel (per-CTA partials + a combine kernel) instead of a separate pass
 * over all T keys. */
typedef struct oat_attn_args {
  int32_t mode, B, T, H, F, n;
  const void* qkv; int64_t ld_qkv;
  void* out; int64_t ld_out;
  float* lse;
  const int32_t* key_mask;
  const void* dout; int64_t ld_dout;
  void* dqkv; int64_t ld_dqkv;
  float scale;
  float* cls_acc;
  /* mode 2 only: dropout on the softmax weights (HF DistilBERT attention_dropout, active because the reference keeps
   * text_model.train(), model/oa_model.py:28). Weight (b, h, i, j) is dropped iff the Philox draw of element
   * ((b*H + h)*T + i)*T + j at site dropout_site is < dropout_p * 2^32; kept weights are scaled by 1 / (1 - p). */
  float dropout_p;
  uint32_t dropout_site;
  uint64_t dropout_seed;
  /* backward, modes 0/1, optional: delta[h * ld_delta + b*T + t] = sum_d dout[b*T + t][h*64 + d] * out[b*T + t][h*64 + d]
   * precomputed by the GEMM that produced dout (oat_gemm_bf16 act 4). NULL: the kernels compute it from out and dout. */
  const float* delta; int64_t ld_delta;
} oat_attn_args;
size_t oat_attn_fwd_workspace_floats(int32_t mode, int32_t B, int32_t H, int32_t F, int32_t n);
int oat_attn_fwd(const oat_attn_args* args, oat_stream_t stream);
int oat_attn_bwd(const oat_attn_args* args, oat_stream_t stream);

/* ---- operand packing and token bookkeeping (HBM-bound) ---------------------------------------------------------
 * oat_cast_bf16: dst[r, c] = bf16(src[r, c]) for c < cols, 0 for cols <= c < cols_padded (optional ReLU first:
 *   the ReLU of txt_proj, oa_model.py:68). Used for weights, region features (2054 -> padded pitch) and CLS rows.
 * oat_relu_bwd: dx = (x > 0) ? dy : 0.
 * oat_im2col_patches: video fp32 [BF, C, H, W] -> bf16 [BF*(H/P)*(W/P), C*P*P] in Conv2d weight order
 *   (VideoPatchEmbed, video_transformer.py:69-76: kernel = stride = P makes the convolution a GEMM).
 * oat_assemble_tokens: x[b,0] = cls + pos[0]; x[b,1+f*n+i] = patch + pos[1+i] + temporal[f] (i < N);
 *   x[b,1+f*n+N+o] = object + temporal[f]; n = N + O; optional token-type rows [2, D]
 *   (forward_features, video_transformer.py:303-325; oa_video_transformer_region.py:250-261).
 * oat_assemble_tokens_bwd: scatters dx to bf16 dpatch / dobject and ACCUMULATES dcls, dpos, dtemporal, dtype.
 * oat_colsum_bf16: out[c] += sum_r x[r, c]  (bias gradients).
 * oat_text_embed(_bwd): DistilBERT word + position embedding sum and its scatter-add gradient. */
int oat_cast_bf16(const float* src, int64_t lds, void* dst_bf16, int64_t ldd, int64_t rows, int32_t cols,
                  int32_t cols_padded, int32_t relu, oat_stream_t stream);
/* Split-bf16 operands: the rows the logits depend on directly (the video tower's CLS rows, the text tower, the two
 * projections of oa_model.py:68-75) take three bf16 MMAs per product instead of one, x.w ~= hi.hi + hi.lo + lo.hi with
 * hi = bf16(v), lo = bf16(v - hi), as ONE K-concatenated oat_gemm_bf16: activation rows [hi | hi | lo] (this call,
 * dst pitch ldd >= 3*cols, optional ReLU first) against weight rows [hi | lo | hi] (oat_cast_multi kind 2). */
int oat_split3_bf16(const float* src, int64_t lds, void* dst_bf16, int64_t ldd, int64_t rows, int32_t cols,
                    int32_t relu, oat_stream_t stream);
/* Many casts in one launch. table[n][8] (device memory) = {src fp32*, dst*, rows, cols, cols_padded, src pitch,
 * dst pitch, kind}; chunk_prefix[n+1] = prefix sum of ceil(rows * cols_padded / 1024). kind 0: same element semantics as
 * oat_cast_bf16; kind 1: plain fp32 copies (used to pack q/k/v biases); kind 2: split-bf16 weight copy, dst row =
 * [hi | lo | hi] in segments of cols_padded (dst pitch >= 3 * cols_padded). */
int oat_cast_multi(const int64_t* table, const int64_t* chunk_prefix, int32_t n, int64_t total_chunks,
                   oat_stream_t stream);
int oat_relu_bwd(const float* x, int64_t ldx, const void* dy_bf16, int64_t lddy, float* dx, int64_t lddx,
                 int64_t rows, int32_t cols, oat_stream_t stream);
/* Element dropout of the text tower in training mode (HF DistilBERT `dropout`: after the embedding LayerNorm and on
 * the FFN output before the residual add; the reference leaves text_model.train() on, model/oa_model.py:28).
 * Masks are Philox4x32-10 draws of (seed, site, element index = row * cols + col): nothing is stored.
 * oat_dropout_fwd: out = keep(x) / (1 - p) + residual (optional); also as bf16 and / or split-bf16 [hi | hi | lo].
 * oat_dropout_bwd: dx = keep(dy_f32 + dy_bf16) / (1 - p) as fp32 and / or bf16 (either input / output may be NULL).
 * oat_dropout_mask: keep[idx] (uint8) for idx < n - what the oracle is given so that both sides drop the same elements. */
int oat_dropout_fwd(const float* x, int64_t ldx, const float* residual, int64_t ldr, float* out, int64_t ldo,
                    void* out_bf16, int64_t ldob, void* out_split3, int64_t ld3, int64_t rows, int32_t cols, float p,
                    uint64_t seed, uint32_t site, oat_stream_t stream);
int oat_dropout_bwd(const float* dy_f32, int64_t lddy, const void* dy_bf16, int64_t lddyb, float* dx_f32, int64_t lddx,
                    void* dx_bf16, int64_t lddxb, int64_t rows, int32_t cols, float p, uint64_t seed, uint32_t site,
                    oat_stream_t stream);
int oat_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint32_t site, oat_stream_t stream);
int oat_im2col_patches(const float* video, void* out_bf16, int64_t BF, int32_t C, int32_t H, int32_t W, int32_t P,
                       oat_stream_t stream);
int oat_assemble_tokens(const float* patch, const float* object, const float* cls_token, const float* pos_embed,
                        const float* temporal_embed, const float* type_embed, float* x, int32_t B, int32_t F,
                        int32_t N, int32_t O, int32_t D, oat_stream_t stream);
int oat_assemble_tokens_bwd(const float* dx, void* dpatch_bf16, void* dobject_bf16, float* dcls, float* dpos,
                            float* dtemporal, float* dtype_embed, int32_t B, int32_t F, int32_t N, int32_t O,
                            int32_t D, oat_stream_t stream);
int oat_colsum_bf16(const void* x_bf16, int64_t ld, int64_t rows, int32_t cols, float* out, oat_stream_t stream);
/* out[c] += sum_k v[k] * W[k*ldw + c] (fp32 row vector times matrix). The qkv bias gradient of VarAttention
 * (video_transformer.py:102) needs a column sum over all token rows of dq only: softmax makes the rows of dS sum to
 * zero, so sum_j dK_j = sum_i q_i (sum_j dS_ij) = 0, and its rows of P sum to one, so sum_j dV_j = sum_i dO_i = db_proj . W_proj
 * (dO = dY_proj . W_proj, video_transformer.py:133) - this call, on the proj bias gradient and the fp32 proj weight. */
int oat_vecmat_f32(const float* v, const float* W, int64_t ldw, int32_t K, int32_t N, float* out, oat_stream_t stream);
/* Bias gradient for free: a weight-gradient GEMM dY^T . [X | 1 0 .. 0] (activation rows extended by a ones column, pitch
 * cols + 16) leaves dW in columns [0, cols) and the column sums of dY - the bias gradient - in column `cols` of its fp32
 * scratch output [rows, ld]; oat_gemm_bf16 multiplies that last, 16-wide column block with an N = 16 instruction. This
 * call adds both into their gradient tensors (dw [rows, ldw], db [rows]) and re-zeroes the scratch. */
int oat_unpack_wgrad(float* scratch, int64_t ld, int32_t cols, float* dw, int64_t ldw, float* db, int64_t rows,
                     oat_stream_t stream);
int oat_text_embed(const int64_t* ids, const float* word_emb, const float* pos_emb, float* out, int64_t rows,
                   int32_t L, int32_t D, oat_stream_t stream);
int oat_text_embed_bwd(const int64_t* ids, const float* dsum, float* dword, float* dpos, int64_t rows, int32_t L,
                       int32_t D, oat_stream_t stream);

/* ---- similarity matrix + symmetric InfoNCE, forward and backward ------------------------------------------------
 * text, video: fp32 [n, P] GATHERED embeddings (rows = the global batch). Writes sims [n, n] (optional), the scalar
 * loss, and dL/dtext, dL/dvideo [n, P] (optional). The caller slices its local rows: the all-gather backward is a
 * slice without reduction (trainer_dist.py:40-45). Replaces model/model.py:164-172 + model/loss.py:13-25. */
size_t oat_sim_workspace_bytes(int32_t n, int32_t m, int32_t P);
/* sims[n, m] (pitch m) = unit(text[n,P]) . unit(video[m,P])^T; the workspace keeps what the backward needs. */
int oat_sim_matrix_fwd(const float* text, const float* video, int32_t n, int32_t m, int32_t P, float eps,
                       float* sims, void* workspace, size_t workspace_bytes, oat_stream_t stream);
int oat_sim_matrix_bwd(const float* dsims, int32_t n, int32_t m, int32_t P, float eps, float* dtext, float* dvideo,
                       void* workspace, size_t workspace_bytes, oat_stream_t stream);
/* loss[0] = NormSoftmaxLoss(sims[n,n]); dsims (optional, same pitch) = dL/dsims; scratch: fp32 [2n]. */
int oat_norm_softmax_loss(const float* sims, int32_t n, int64_t ld, float temperature, float* loss, float* dsims,
                          float* scratch, oat_stream_t stream);
size_t oat_infonce_workspace_bytes(int32_t n, int32_t P);
int oat_infonce_fwd_bwd(const float* text, const float* video, int32_t n, int32_t P, float temperature, float eps,
                        float* sims_out, float* loss, float* dtext, float* dvideo, void* workspace,
                        size_t workspace_bytes, oat_stream_t stream);

/* ---- object -> patch attention (three score->weight modes) and bbox -> patch masks ---------------------------------
 * mode 0: weights = masks (B,O,L), out = masks @ v            (oa_model_global_local.py:178)
 * mode 1: weights = sigmoid(q . k^T)                          (oa_model_region_mem.py:147-151)
 * mode 2: weights = softmax(q . k^T * C^-0.5)                 (Visualization/.../visualize.py:155-168)
 * q (B,O,C), k (B,L,C), v (B,L,Cv) fp32; weights (B,O,L) and/or out (B,O,Cv) fp32 (either may be NULL).
 * oat_patch_masks_from_bbox: boxes fp64 [n, stride] (x1,y1,x2,y2 in [0,1]) -> masks fp32 [n, grid*grid], exactly
 * mask[int(y1*g):ceil(y2*g), int(x1*g):ceil(x2*g)] = 1 of base/base_dataset_global_local.py:348-356 (bit-exact). */
int oat_object_patch_attn(const float* q, const float* k, const float* v, const float* masks, float* weights,
                          float* out, int32_t B, int32_t O, int32_t L, int32_t C, int32_t Cv, int32_t mode,
                          oat_stream_t stream);
int oat_patch_masks_from_bbox(const double* boxes, int32_t stride, float* masks, int32_t n, int32_t grid,
                              oat_stream_t stream);
/* The other bookkeeping in front of the path, bit-exact with the reference:
 * oat_patch_masks_same_class: base/base_dataset_region_mem.py:233-247 - masks[j] = union over every box i with
 *   classes[i] == classes[sel[j]] of the box's patch rectangle (boxes fp64 [n, stride], scaled by `grid` and cut with
 *   int() / ceil() as above); sel[para] are the indices random.sample drew on the host.
 * oat_object_tags_masks: base/base_dataset_global_local.py:395-405 - ends[i] = running sum of int(token_lens[indices[i]])
 *   (token_lens fp64 as np.loadtxt reads them), total[0] = the sum.
 * oat_region_features_topk: base/base_dataset.py:593-650 after np.load - rows by descending confidence, v = 2 keeps the
 *   first region of every object class (np.unique order), np.pad(..., 'edge') to top_k rows - which also pads the
 *   columns, so with m < top_k regions the row is feat_dim + (top_k - m) + 6 wide, as the reference produces it -
 *   then [x1/W, y1/H, x1/W + w/W, y1/H + h/H, w/W, h/H]. out: fp32 [top_k, ld_out], ld_out >= feat_dim + top_k + 6;
 *   m_out[0] = number of regions before padding. n <= 128. */
int oat_patch_masks_same_class(const double* boxes, int32_t stride, const int32_t* classes, const int32_t* sel,
                               float* masks, int32_t n, int32_t para, int32_t grid, oat_stream_t stream);
int oat_object_tags_masks(const double* token_lens, const int64_t* indices, float* ends, int32_t* total, int32_t k,
                          oat_stream_t stream);
int oat_region_features_topk(const float* x, const float* bbox, const float* conf, const int64_t* ids, int32_t n,
                             int32_t feat_dim, int32_t top_k, int32_t v, int32_t image_w, int32_t image_h, float* out,
                             int64_t ld_out, int32_t* m_out, oat_stream_t stream);
/* Backward of oat_object_patch_attn. weights = what the forward returned (the masks for mode 0); dweights (B,O,L)
 * and/or dout (B,O,Cv) are the incoming gradients (either may be NULL); dq (B,O,C), dk (B,L,C), dv (B,L,Cv) are
 * written (any may be NULL); ds_scratch: fp32 (B,O,L), needed for modes 1/2. Mode 0 propagates to v only (the masks
 * are data). Replaces autograd through oa_model_global_local.py:178 / oa_model_region_mem.py:147-151. */
int oat_object_patch_attn_bwd(const float* q, const float* k, const float* v, const float* weights,
                              const float* dweights, const float* dout, float* dq, float* dk, float* dv,
                              float* ds_scratch, int32_t B, int32_t O, int32_t L, int32_t C, int32_t Cv, int32_t mode,
                              oat_stream_t stream);

/* ---- variant heads ----------------------------------------------------------------------------------
 * oat_token_pool: out[b] = a * cls[b] + bcoef * mean_l tok[b, l]  (cls may be NULL); token row (b, l) starts at
 *   tok + b * ld_batch + l * ld_tok, so a [:, 1:] slice of a (B, T, P) tensor is passed without a copy.
 *   `video_embeddings = (video_embeddings + mean(video_region_feature, 1)) / 2` (oa_model_region_mem.py:119) is
 *   a = bcoef = 0.5; `torch.mean(region_feat, dim=1)` (trainer/trainer_global_local.py:207) is cls = NULL, bcoef = 1.
 * oat_token_pool_bwd: dcls = a * dout (optional), dtok[b, l] = bcoef / L * dout[b].
 * oat_bce_sum: loss[0] = scale * BCELoss(reduction='sum')(p, target) and dp = d loss / d p (optional), with
 *   torch's clamps (log at -100, gradient denominator at 1e-12): the region loss `0.1 * BCE_sum / rows`
 *   (trainer/trainer_region_mem.py:97,161-167) is scale = 0.1 / rows. One CTA, fixed summation order. */
int oat_token_pool(const float* cls, int64_t ld_cls, const float* tok, int64_t ld_batch, int64_t ld_tok, float* out,
                   int32_t B, int32_t L, int32_t P, float a, float bcoef, oat_stream_t stream);
int oat_token_pool_bwd(const float* dout, float* dcls, float* dtok, int64_t ld_batch, int64_t ld_tok, int32_t B,
                       int32_t L, int32_t P, float a, float bcoef, oat_stream_t stream);
int oat_bce_sum(const float* p, const float* target, int64_t n, float scale, float* loss, float* dp,
                oat_stream_t stream);

/* ---- retrieval ranks on a square similarity matrix (rows = text queries, columns = videos) ------------------------------
 * t2v_rank[i] = #{j : sims[i][j] > sims[i][i]}                       ties broken optimistically (model/metric.py:62-69)
 * v2t_rank[i] = #{j : sims[j][i] > sims[i][i]} + (ties_i - 1) / 2     tied ranks averaged        (model/metric.py:153,183)
 * i.e. the column of the ground-truth pair in the sorted distance row, exactly as np.sort + np.where produce it;
 * R@K / MedR / MeanR (cols2metrics, metric.py:281-291) are then counts over these n numbers. Integer bookkeeping:
 * bit-exact with the reference. */
int oat_retrieval_ranks(const float* sims, int32_t n, int64_t ld, float* t2v_rank, float* v2t_rank,
                        oat_stream_t stream);

/* ---- optimizer step (the step right after the path) ------------------------------------------------
 * Fused multi-tensor AdamW with transformers.AdamW semantics (train_dist_multi.py:66 builds the optimizer from the
 * `transformers` module): table[n][4] = device pointers (p, g, m, v), all fp32; sizes[n] = element counts;
 * chunk_prefix[n+1] = prefix sum of ceil(size / 1024). All three tables live in DEVICE memory. `step` is 1-based. */
int oat_adamw_multi(const int64_t* table, const int64_t* chunk_prefix, const int64_t* sizes, int32_t n,
                    int64_t total_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int32_t step, int32_t correct_bias, oat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OAT_H_ */
