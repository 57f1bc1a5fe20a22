/* egovlp_b200 — C-ABI of the B200-native EgoVLP hot path (libegovlp_b200.so).
 *
 * The reference (showlab/EgoVLP) has no FFI layer: its hot path is Python over torch ops
 * (SURVEY.md section 8b).  This header is the boundary a maintainer binds instead of those torch
 * ops: plain device pointers, sizes and a CUstream/cudaStream_t passed as void*.  Every entry point
 * cites the reference call site(s) it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host_*; the caller owns every buffer; kernels
 *     never allocate or free; work is enqueued on `stream` and the call returns immediately.
 *   - return 0 on success, <0 on error (EGOVLP_ERR_*); egovlp_last_error() gives the message
 *     (thread-local).  No global state except per-process function attributes.
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t); "row-major [R, C] ld" = element (r,c) at r*ld+c.
 *   - token layout of the video tower: x[b, s, :] with s = 0 the CLS token and s = 1 + f*N + n the
 *     patch n of frame f (frame-major), exactly the reference's (model/video_transformer.py:305-310).
 */
#ifndef EGOVLP_B200_H_
#define EGOVLP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGOVLP_ERR_ARG (-1)
#define EGOVLP_ERR_CUDA (-2)
#define EGOVLP_ERR_UNSUPPORTED (-3)

const char* egovlp_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int egovlp_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM  D[m,n] = epi( sum_k A[m,k] * B[n,k] ),  bf16 operands, fp32 accumulation (tcgen05/TMEM).
 * Replaces nn.Linear forward/backward everywhere on the path: model/video_transformer.py:41-52
 * (Mlp), :88-89,103,135 (qkv/proj), :70,76 (patch-embed conv as GEMM); DistilBERT q/k/v/out_lin,
 * ffn.lin1/lin2; model/model.py:72-79 (projections); and autograd's dgrad/wgrad of each.
 *   a_mn_major = 0: A stored [M, K] (ld = lda);  1: A stored [K, M] (m contiguous)
 *   b_mn_major = 0: B stored [N, K] (ld = ldb);  1: B stored [K, N] (n contiguous)
 * Epilogue, applied in this order to v = alpha * acc:
 *   v += bias[n];  if (n < col_scale_ncols) v *= col_scale;  out2[m,n] = bf16(v) (if out2);
 *   act 1: v = gelu_erf(v)   act 2: v *= gelu_erf'(aux[m,n]);   v += residual[m,n] (fp32);
 *   out_mode 0: out(bf16) = v;  1: out(fp32) = v;  2: atomicAdd(out(fp32), v) (needed for split_k>1)
 * Constraints: N % 32 == 0, lda/ldb/ldo % 8 == 0, 16B-aligned bases.
 */
typedef struct egovlp_gemm_epilogue {
  const float* bias;     /* [N] fp32 or NULL */
  const float* residual; /* [M, ldr] fp32 or NULL */
  const void* aux;       /* [M, ldaux] bf16, for act == 2 */
  void* out;             /* [M, ldo] bf16 (out_mode 0) or fp32 (1, 2) */
  void* out2;            /* [M, ldo2] bf16 or NULL */
  long long ldr, ldaux, ldo, ldo2;
  int out_mode;
  int act;
  float alpha;
  float col_scale;
  int col_scale_ncols;
} egovlp_gemm_epilogue;

int egovlp_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb,
                     int M, int N, int K, const egovlp_gemm_epilogue* epi, int split_k, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (nn.LayerNorm; model/video_transformer.py:146,156,159,228,253 eps 1e-6;
 * DistilBERT sa_layer_norm/output_layer_norm/embeddings.LayerNorm eps 1e-12).
 * x fp32 [rows, D] (row stride ldx) -> y bf16 [rows, D] and/or y32 fp32 [rows, D]; mean/rstd fp32 [rows]
 * are saved for the backward.  Optional fused residual: x_eff = x + add (fp32 [rows, D]), and x_eff is
 * written to sum_out (post-LN DistilBERT: LN(sa + x)).  D % 4 == 0 and D <= 1024.
 */
int egovlp_layernorm_fwd(const float* x, long long ldx, const float* add, float* sum_out, const float* gamma,
                         const float* beta, void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int D,
                         float eps, void* stream);
/* Backward.  dx = LNbwd(dy) [+ add1] [+ add2]  (fp32 [rows, D], row stride D), optionally also stored as bf16
 * (dx_bf16) for use as a GEMM operand.  add1/add2 carry the residual-stream gradients that bypass the LN
 * (SpaceTimeBlock: dsr = dy + LN2bwd, dx = dsr + dtr + LN3bwd).  dgamma/dbeta (fp32 [D]) are ACCUMULATED
 * with atomicAdd -- zero them first for a plain gradient; either may be NULL. */
int egovlp_layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* gamma,
                         const float* mean, const float* rstd, const float* add1, const float* add2, float* dx,
                         void* dx_bf16, float* dgamma, float* dbeta, int rows, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Divided space-time attention core of VarAttention.forward (model/video_transformer.py:104-133) and its
 * autograd.  Head dim must be 64 (D = 64*H).  mode 0 = time ('b (f n) d -> (b n) f d'), 1 = space ('(b f) n d').
 *   qkv   bf16 [B*S, 3*D], S = 1 + T*N, columns [q | k | v] each (head, 64); q ALREADY scaled by 64^-0.5
 *         (the QKV GEMM epilogue applies it, :106)
 *   out   bf16 [B*S, D]  = cat(cls_out, attended patches) with heads merged (:130-133), ready for proj
 *   lse   fp32 [B, H, S] log-sum-exp of every query row (saved for the backward)
 *   cls_part fp32 workspace of egovlp_divided_attn_workspace_floats() floats (CLS-query partials)
 * Semantics kept from the reference: the CLS query attends over ALL S keys (:112); every patch query attends
 * over its group's keys plus the CLS key/value (:117-124).
 * Backward: dqkv bf16 [B*S, 3*D] receives d(q_prescale), dk, dv for every token (q gradient multiplied by
 * q_scale); dcls_ws is an fp32 workspace of B*H*3*64 floats (zeroed internally).
 */
long long egovlp_divided_attn_workspace_floats(int B, int T, int N, int H, int mode);
int egovlp_divided_attn_fwd(const void* qkv, void* out, float* lse, float* cls_part, int B, int T, int N, int H,
                            int mode, void* stream);
int egovlp_divided_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                            float* dcls_ws, int B, int T, int N, int H, int mode, float q_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise / reduction helpers on the path.
 */
/* fp32 -> bf16 cast (weights: fp32 master -> bf16 GEMM operand). */
int egovlp_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* out[n] += sum_m dy[m, n]  (bias gradients).  dy bf16 or fp32 [M, N] row stride ld. */
int egovlp_colsum_accum(const void* dy, int dy_is_fp32, long long ld, float* out, int M, int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGOVLP_B200_H_ */
