// Fused multi-tensor AdamW with HuggingFace transformers.AdamW semantics (the optimizer the reference's configs name:
// configs/pt/egoclip.json:49-54 via run/train_egoclip.py:72-73): eps added OUTSIDE the bias correction,
// decoupled weight decay applied after the update with the un-corrected lr.
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= step_size * m / (sqrt(v) + eps);  p -= lr * wd * p
// One launch for all tensors: a chunk table maps each CTA to (tensor, offset).  HBM-bound: 28 B/param.
#include "common.cuh"
#include "egovlp_b200.h"

namespace egovlp {
namespace {

constexpr int CHUNK = 4096;   // elements per CTA

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const egovlp_adamw_tensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                   const int* __restrict__ chunk_offset, float lr, float beta1, float beta2, float eps, float wd,
                   float step_size, const float* __restrict__ grad_scale) {
  const egovlp_adamw_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const long long base = (long long)chunk_offset[blockIdx.x] * CHUNK;
  const float gs = grad_scale ? *grad_scale : 1.f;
  float* p = t.param; const float* g = t.grad; float* m = t.exp_avg; float* v = t.exp_avg_sq;
  bf16* sh = reinterpret_cast<bf16*>(t.shadow_bf16);      // optional bf16 GEMM-operand copy refreshed in the same pass
  const long long end = min(t.numel, base + CHUNK);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(sh) & 7) == 0;
  for (long long i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
    if (vec && i + 3 < end) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = ga[k] * gs;
        ma[k] = beta1 * ma[k] + (1.f - beta1) * gk;
        va[k] = beta2 * va[k] + (1.f - beta2) * gk * gk;
        pa[k] -= step_size * ma[k] / (sqrtf(va[k]) + eps);
        if (wd > 0.f) pa[k] -= lr * wd * pa[k];
      }
      *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
      if (sh) *reinterpret_cast<uint2*>(sh + i) = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
    } else {
      for (long long j = i; j < min(end, i + 4); ++j) {
        const float gk = g[j] * gs;
        const float mj = beta1 * m[j] + (1.f - beta1) * gk, vj = beta2 * v[j] + (1.f - beta2) * gk * gk;
        float pj = p[j] - step_size * mj / (sqrtf(vj) + eps);
        if (wd > 0.f) pj -= lr * wd * pj;
        m[j] = mj; v[j] = vj; p[j] = pj;
        if (sh) sh[j] = __float2bfloat16_rn(pj);
      }
    }
  }
}

// fp32 -> bf16 for many tensors in one launch (same chunk-table scheme): refreshes every bf16 GEMM-operand copy of the
// fp32 master weights at the top of a training forward, whatever optimizer touched them.
__global__ void __launch_bounds__(256)
cast_multi_kernel(const egovlp_cast_tensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                  const int* __restrict__ chunk_offset) {
  const egovlp_cast_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const long long base = (long long)chunk_offset[blockIdx.x] * CHUNK;
  const long long end = min(t.numel, base + CHUNK);
  const float* s = t.src; bf16* d = reinterpret_cast<bf16*>(t.dst_bf16);
  const bool vec = (reinterpret_cast<uintptr_t>(s) & 15) == 0 && (reinterpret_cast<uintptr_t>(d) & 7) == 0;
  for (long long i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
    if (vec && i + 3 < end) {
      const float4 x = *reinterpret_cast<const float4*>(s + i);
      *reinterpret_cast<uint2*>(d + i) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
    } else {
      for (long long j = i; j < min(end, i + 4); ++j) d[j] = __float2bfloat16_rn(s[j]);
    }
  }
}

}  // namespace
}  // namespace egovlp

using namespace egovlp;

extern "C" int egovlp_adamw_chunk_elems(void) { return CHUNK; }

extern "C" int egovlp_adamw_multi(const egovlp_adamw_tensor* tensors_dev, const int* chunk_tensor_dev,
                                  const int* chunk_offset_dev, int n_chunks, float lr, float beta1, float beta2,
                                  float eps, float weight_decay, float step_size, const float* grad_scale_dev,
                                  void* stream) {
  EGOVLP_CHECK_ARG(tensors_dev && chunk_tensor_dev && chunk_offset_dev && n_chunks >= 0, "adamw_multi: bad args");
  if (n_chunks == 0) return EGOVLP_OK;
  adamw_multi_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tensors_dev, chunk_tensor_dev, chunk_offset_dev, lr, beta1, beta2, eps, weight_decay, step_size, grad_scale_dev);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}

extern "C" int egovlp_cast_multi_f32_to_bf16(const egovlp_cast_tensor* tensors_dev, const int* chunk_tensor_dev,
                                             const int* chunk_offset_dev, int n_chunks, void* stream) {
  EGOVLP_CHECK_ARG(tensors_dev && chunk_tensor_dev && chunk_offset_dev && n_chunks >= 0, "cast_multi: bad args");
  if (n_chunks == 0) return EGOVLP_OK;
  cast_multi_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tensors_dev, chunk_tensor_dev,
                                                                                   chunk_offset_dev);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
