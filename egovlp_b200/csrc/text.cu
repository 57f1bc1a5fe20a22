// Text-tower specific kernels (DistilBERT, transformers modeling_distilbert.py; call sites model/model.py:117-138):
// embedding gather (+positions), key-padding-masked self-attention for short sequences (L <= 128), and the
// CLS -> ReLU gather in front of txt_proj (model/model.py:73-75,125).  The Linear / LayerNorm / GELU work of the
// tower runs on the shared tcgen05 GEMM and LayerNorm kernels.  <0.4% of the step's FLOPs: fp32 CUDA-core math.
#include "common.cuh"
#include "egovlp_b200.h"

namespace egovlp {
namespace {

constexpr int HD = 64;

// out[tok, :] = word[ids[tok], :] + pos[tok % L, :]   (fp32)
__global__ void text_embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                      const float* __restrict__ pos, float* __restrict__ out, int ntok, int L, int D) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= (long long)ntok * D) return;
  const int tok = i / D, d = i % D;
  const float4 w = *reinterpret_cast<const float4*>(word + ids[tok] * D + d);
  const float4 p = *reinterpret_cast<const float4*>(pos + (long long)(tok % L) * D + d);
  *reinterpret_cast<float4*>(out + i) = make_float4(w.x + p.x, w.y + p.y, w.z + p.z, w.w + p.w);
}
__global__ void text_embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dsum,
                                      float* __restrict__ dword, float* __restrict__ dpos, int ntok, int L, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ntok * D) return;
  const int tok = i / D, d = i % D;
  const float g = dsum[i];
  atomicAdd(dword + ids[tok] * D + d, g);
  atomicAdd(dpos + (long long)(tok % L) * D + d, g);
}

// One CTA per (b, h).  qkv bf16 [B*L, 3*D] (q pre-scaled), mask int64 [B, L] (0 = padded key).
template <bool BWD>
__global__ void __launch_bounds__(256)
text_attn_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask, bf16* __restrict__ out,
                 const bf16* __restrict__ dout, bf16* __restrict__ dqkv, int B, int L, int H, float q_scale) {
  extern __shared__ float sm[];
  const int D = H * HD, b = blockIdx.x / H, h = blockIdx.x % H;
  const int LP = L + 1, RS = HD + 1;
  float* q = sm;                    // [L][65]
  float* k = q + L * RS;
  float* v = k + L * RS;
  float* dO = v + L * RS;           // bwd only
  float* P = BWD ? dO + L * RS : v + L * RS;   // [L][L+1]
  float* keyok = P + L * LP;        // [L]
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < L * HD; i += blockDim.x) {
    const int r = i / HD, d = i % HD;
    const bf16* row = qkv + ((long long)(b * L + r)) * 3 * D + h * HD + d;
    q[r * RS + d] = __bfloat162float(row[0]);
    k[r * RS + d] = __bfloat162float(row[D]);
    v[r * RS + d] = __bfloat162float(row[2 * D]);
    if (BWD) dO[r * RS + d] = __bfloat162float(dout[((long long)(b * L + r)) * D + h * HD + d]);
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) keyok[i] = mask[(long long)b * L + i] != 0 ? 1.f : 0.f;
  __syncthreads();
  // P = softmax(q k^T + key mask)
  for (int i = warp; i < L; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) {
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < HD; ++d) s += q[i * RS + d] * k[j * RS + d];
      s = keyok[j] != 0.f ? s : -INFINITY;
      P[i * LP + j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = __expf(P[i * LP + j] - mx);
      P[i * LP + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < L; j += 32) P[i * LP + j] *= inv;
  }
  __syncthreads();
  if (!BWD) {
    for (int i = warp; i < L; i += nw) {
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < L; ++j) {
        const float p = P[i * LP + j];
        o0 += p * v[j * RS + lane];
        o1 += p * v[j * RS + lane + 32];
      }
      bf16* dst = out + ((long long)(b * L + i)) * D + h * HD;
      dst[lane] = __float2bfloat16(o0);
      dst[lane + 32] = __float2bfloat16(o1);
    }
    return;
  }
  // ---- backward ----
  // dV_j = sum_i P_ij dO_i
  for (int j = warp; j < L; j += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < L; ++i) {
      const float p = P[i * LP + j];
      a0 += p * dO[i * RS + lane];
      a1 += p * dO[i * RS + lane + 32];
    }
    bf16* dst = dqkv + ((long long)(b * L + j)) * 3 * D + 2 * D + h * HD;
    dst[lane] = __float2bfloat16(a0);
    dst[lane + 32] = __float2bfloat16(a1);
  }
  __syncthreads();
  // dS = P * (dP - delta), in place
  for (int i = warp; i < L; i += nw) {
    float delta = 0.f;
    float dp_local[4];   // L <= 128 -> at most 4 keys per lane
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + jj * 32;
      float dp = 0.f;
      if (j < L) {
#pragma unroll 16
        for (int d = 0; d < HD; ++d) dp += dO[i * RS + d] * v[j * RS + d];
        delta += P[i * LP + j] * dp;
      }
      dp_local[jj] = dp;
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + jj * 32;
      if (j < L) P[i * LP + j] *= (dp_local[jj] - delta);
    }
  }
  __syncthreads();
  // dQ_i = q_scale * sum_j dS_ij K_j ;  dK_j = sum_i dS_ij Q_i
  for (int i = warp; i < L; i += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float s = P[i * LP + j];
      a0 += s * k[j * RS + lane];
      a1 += s * k[j * RS + lane + 32];
    }
    bf16* dst = dqkv + ((long long)(b * L + i)) * 3 * D + h * HD;
    dst[lane] = __float2bfloat16(a0 * q_scale);
    dst[lane + 32] = __float2bfloat16(a1 * q_scale);
  }
  for (int j = warp; j < L; j += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < L; ++i) {
      const float s = P[i * LP + j];
      a0 += s * q[i * RS + lane];
      a1 += s * q[i * RS + lane + 32];
    }
    bf16* dst = dqkv + ((long long)(b * L + j)) * 3 * D + D + h * HD;
    dst[lane] = __float2bfloat16(a0);
    dst[lane + 32] = __float2bfloat16(a1);
  }
}

// out[r, :] = bf16(relu(x[r * row_stride, :]))
__global__ void relu_rows_fwd_kernel(const float* __restrict__ x, long long row_stride, bf16* __restrict__ out, int rows,
                                     int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  const int r = i / D, d = i % D;
  out[i] = __float2bfloat16(fmaxf(0.f, x[(long long)r * row_stride + d]));
}
// dx[r * row_stride, :] = dh[r, :] * (x > 0)    (other rows of dx untouched)
__global__ void relu_rows_bwd_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ dh,
                                     float* __restrict__ dx, int rows, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  const int r = i / D, d = i % D;
  const long long o = (long long)r * row_stride + d;
  dx[o] = x[o] > 0.f ? dh[i] : 0.f;
}

size_t text_attn_smem(int L, bool bwd) { return (size_t)((bwd ? 4 : 3) * L * (HD + 1) + L * (L + 1) + L) * sizeof(float); }

}  // namespace
}  // namespace egovlp

using namespace egovlp;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int egovlp_text_embed_fwd(const long long* input_ids, const float* word_emb, const float* pos_emb,
                                     float* out, int B, int L, int D, void* stream) {
  EGOVLP_CHECK_ARG(input_ids && word_emb && pos_emb && out && B > 0 && L > 0 && D % 4 == 0, "text_embed_fwd: bad args");
  const long long n = (long long)B * L * D / 4;
  text_embed_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(input_ids, word_emb, pos_emb, out, B * L, L, D);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
extern "C" int egovlp_text_embed_bwd(const long long* input_ids, const float* dsum, float* dword, float* dpos, int B,
                                     int L, int D, void* stream) {
  EGOVLP_CHECK_ARG(input_ids && dsum && dword && dpos && B > 0 && L > 0, "text_embed_bwd: bad args");
  const long long n = (long long)B * L * D;
  text_embed_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(input_ids, dsum, dword, dpos, B * L, L, D);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
extern "C" int egovlp_text_attn_fwd(const void* qkv, const long long* attention_mask, void* out, int B, int L, int H,
                                    void* stream) {
  EGOVLP_CHECK_ARG(qkv && attention_mask && out && B > 0 && H > 0, "text_attn_fwd: bad args");
  EGOVLP_CHECK_ARG(L > 0 && L <= 128, "text_attn: L=%d unsupported (1..128)", L);
  const size_t smem = text_attn_smem(L, false);
  auto kern = text_attn_kernel<false>;
  EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B * H, 256, smem, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), attention_mask,
                                        reinterpret_cast<bf16*>(out), nullptr, nullptr, B, L, H, 1.f);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
extern "C" int egovlp_text_attn_bwd(const void* qkv, const long long* attention_mask, const void* dout, void* dqkv,
                                    int B, int L, int H, float q_scale, void* stream) {
  EGOVLP_CHECK_ARG(qkv && attention_mask && dout && dqkv && B > 0 && H > 0, "text_attn_bwd: bad args");
  EGOVLP_CHECK_ARG(L > 0 && L <= 128, "text_attn: L=%d unsupported (1..128)", L);
  const size_t smem = text_attn_smem(L, true);
  auto kern = text_attn_kernel<true>;
  EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B * H, 256, smem, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), attention_mask, nullptr,
                                        reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dqkv), B, L, H,
                                        q_scale);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
extern "C" int egovlp_relu_rows_fwd(const float* x, long long row_stride, void* out_bf16, int rows, int D,
                                    void* stream) {
  EGOVLP_CHECK_ARG(x && out_bf16 && rows > 0 && D > 0, "relu_rows_fwd: bad args");
  const long long n = (long long)rows * D;
  relu_rows_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, row_stride, reinterpret_cast<bf16*>(out_bf16), rows, D);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
extern "C" int egovlp_relu_rows_bwd(const float* x, long long row_stride, const float* dh, float* dx, int rows, int D,
                                    void* stream) {
  EGOVLP_CHECK_ARG(x && dh && dx && rows > 0 && D > 0, "relu_rows_bwd: bad args");
  const long long n = (long long)rows * D;
  relu_rows_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, row_stride, dh, dx, rows, D);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
