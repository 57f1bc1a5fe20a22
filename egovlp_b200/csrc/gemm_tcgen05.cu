// Persistent warp-specialised bf16 GEMM on tcgen05 / TMEM, operands staged by TMA (128B swizzle).
//
//   D[m,n] = epilogue( sum_k A[m,k] * B[n,k] )
//
// Replaces every nn.Linear / einsum-free contraction the reference reaches through cuBLAS sgemm:
//   qkv / proj  (model/video_transformer.py:88-89,103,135), Mlp.fc1/fc2 (:41-52), patch-embed conv-as-GEMM
//   (:70,76), DistilBERT q/k/v/out_lin + ffn (transformers modeling_distilbert.py), projections
//   (model/model.py:72-79) and all their backward dgrad / wgrad contractions (autograd in the reference).
//
// Layout: A and B are bf16 in HBM.  "K-major" = the contraction index is contiguous (x[M,K], W[N,K]);
// "MN-major" = the m/n index is contiguous (stored [K, M] / [K, N]) which is what dgrad (W as B) and wgrad
// (dy and x as A and B, contraction over tokens) need -- no transposed copies are ever materialised.
// One CTA per SM, 128 x BLOCK_N output tile, 64-deep k-blocks, 4-6 stage TMA->smem ring, fp32 accumulators
// double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one lane), 2 = TMEM allocator (2 and 3 also sum the A tiles' columns in
// the wgrad form), 4..11 = epilogue.  setmaxnreg gives the control warpgroup 64 registers and the epilogue warps 216.
// Epilogue per warp: 32 rows x BLOCK_N/2 columns in 32-column chunks, software-pipelined (TMEM chunk c+1 and the
// per-element global operands of chunk c+1 are requested while chunk c is computed; the first chunk of the NEXT tile
// during the last chunk of this one).  The CTA-pair kernels are additionally instantiated with compile-time specialised
// epilogues (EpiMode) for the forms that carry the training / inference step; the host picks the mode from the
// epilogue descriptor, everything else takes the generic epilogue (same arithmetic, same order: bit-identical results).
#include <stdlib.h>

#include "common.cuh"
#include "egovlp_b200.h"

namespace egovlp {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
#ifndef EGOVLP_EPI16
#define EGOVLP_EPI16 0
#endif
// Epilogue warps: 8 (two per scheduler).  -DEGOVLP_EPI16=1 builds the specialised CTA-pair kernels with 16 (four per
// scheduler, 32 rows x 64 columns each, 5 smem stages, no look-ahead): measured 3-10 % SLOWER on every hot form
// (fc1 1.075 vs 1.040 ms, fc2-dgrad 1.052 vs 0.941, proj 0.379 vs 0.337) -- the step runs at the board's power cap
// (SM clock 1.6-1.7 of 1.965 GHz), where time follows the energy of the instructions and bytes, not their latency.
__host__ __device__ constexpr int epi_warps(bool two, int mode) { return (two && mode != 0 && EGOVLP_EPI16) ? 16 : 8; }
constexpr int MN_ATOM_BYTES = BLOCK_K * 128;  // one 64(mn) x BLOCK_K(k) MN-major slab

struct EpiParams {
  const float* bias;
  const float* residual;
  const bf16* aux;
  void* out;
  bf16* out2;
  long long ldr, ldaux, ldo, ldo2;
  int out_mode;  // 0 bf16 store, 1 fp32 store, 2 fp32 atomic add
  int act;       // 0 none, 1 gelu(erf), 2 multiply by gelu'(aux), 3 gelu(erf) with out2 = gelu', 4 multiply by aux
  float alpha;
  float col_scale;
  int col_scale_ncols;
  int res_row_mod;  // 0: residual row = m; else residual row = m % res_row_mod (broadcast table)
  float* colsum;    // optional fp32 [N]: accumulates the column sums of the stored values (bias gradient)
  float* colsum_a;  // optional fp32 [M], MN-major A / MN-major B (wgrad) only: accumulates sum_k A[k, m], i.e. the bias
                    // gradient of the Linear whose weight gradient this GEMM computes, from the A tiles already in smem
  int debug_no_loads;  // profiling aid (EGOVLP_GEMM_DEBUG_NOLOADS=1): skip the TMA loads, MMAs run on stale smem
};

// TWO = CTA pair (cta_group::2): a 256 x 256 tile per cluster, each CTA stages its 128 rows of A and HALF of B
// (128 of the 256 n-rows).  Every smem byte then feeds twice the MMA work, which is what the 128 B/clk shared
// memory port needs: in single-CTA mode TMA writes (96 B/clk) + UMMA operand reads (96 B/clk) oversubscribe it.
template <int BLOCK_N, bool TWO, int EW = 8>
struct Cfg {
  static constexpr int B_ROWS = TWO ? BLOCK_N / 2 : BLOCK_N;
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGES = (BLOCK_N == 256 && !TWO) ? 4 : (EW == 16 ? 5 : 6);   // 16 x 4 KB of staging costs a stage
  static constexpr int NUM_THREADS = (4 + EW) * 32;
  static constexpr int TILE_M = TWO ? 2 * BLOCK_M : BLOCK_M;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int EPI_STAGE_BYTES = EW * 4096;
  static constexpr int SMEM_BYTES =
      STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// GELU (exact-erf form) in the epilogue.  Phi(x) = 0.5 (1 + erf(x / sqrt2)) through Abramowitz-Stegun 7.1.26
// (|err| <= 1.5e-7 on erf).  With u = k x, k = sqrt(log2(e) / 2) (so that exp(-x^2 / 2) = 2^(-u^2)),
// t = 1 / (1 + p |x| / sqrt2) and e' = c exp(-x^2 / 2) = 2^(log2 c - u^2), c = 1 / (k sqrt(2 pi)):
//     r = 0.5 erf(|x| / sqrt2) = 0.5 + e' t (b1 + t (b2 + t (b3 + t (b4 + t b5))))      (b = -a / (2 c))
//     Phi(x) = 0.5 + copysign(r, x)       GELU = x Phi       GELU' = Phi + x phi = Phi + u e'
// The fc1 epilogue has to fit in the 6144-cycle MMA time of a K = 768 tile with two warps per scheduler, i.e. ~24 issue
// slots per element including the loads, packs and stores, so the arithmetic is packed two elements per instruction
// (FFMA2 / FMUL2 / FADD2): per PAIR 4 MUFU + 2 FFMA (|u| needs the scalar form's abs modifier) + 2 LOP3 (copysign) +
// 11 packed instructions, = 9.5 issue slots per element for GELU and GELU' together (17.4 in the scalar form).
constexpr float GELU_K = 0.84932180028801907f;          // sqrt(log2(e) / 2)
constexpr float GELU_PT = 0.27273749f;                  // 0.3275911 / sqrt2 / GELU_K
constexpr float GELU_LOG2C = -1.0901312f;               // log2(c), c = 0.39894228 / GELU_K = 0.46971865
constexpr float GELU_B1 = -0.5f * 0.254829592f / 0.46971865f, GELU_B2 = 0.5f * 0.284496736f / 0.46971865f,
                GELU_B3 = -0.5f * 1.421413741f / 0.46971865f, GELU_B4 = 0.5f * 1.453152027f / 0.46971865f,
                GELU_B5 = -0.5f * 1.061405429f / 0.46971865f;
struct GeluConsts {       // packed constants, built once per warp (register pairs)
  f32x2 k, nk, l2c, b1, b2, b3, b4, b5, half;
  __device__ __forceinline__ GeluConsts()
      : k(pk2(GELU_K, GELU_K)), nk(pk2(-GELU_K, -GELU_K)), l2c(pk2(GELU_LOG2C, GELU_LOG2C)), b1(pk2(GELU_B1, GELU_B1)),
        b2(pk2(GELU_B2, GELU_B2)), b3(pk2(GELU_B3, GELU_B3)), b4(pk2(GELU_B4, GELU_B4)), b5(pk2(GELU_B5, GELU_B5)),
        half(pk2(0.5f, 0.5f)) {}
};
// two elements: Phi (returned), u = k x and e' for the derivative
__device__ __forceinline__ f32x2 gelu_phi2(const GeluConsts& gc, float x0, float x1, f32x2 x, f32x2& u, f32x2& e) {
  u = mul2(x, gc.k);
  float u0, u1, a0, a1, t0, t1, e0, e1, r0, r1;
  up2(u, u0, u1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(fabsf(u0), GELU_PT, 1.f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(fabsf(u1), GELU_PT, 1.f)));
  up2(fma2(mul2(x, gc.nk), u, gc.l2c), a0, a1);          // log2 c - u^2
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f32x2 t = pk2(t0, t1);
  e = pk2(e0, e1);
  f32x2 p = fma2(t, gc.b5, gc.b4);
  p = fma2(p, t, gc.b3);
  p = fma2(p, t, gc.b2);
  p = fma2(p, t, gc.b1);
  up2(fma2(mul2(p, t), e, gc.half), r0, r1);
  return add2(pk2(copysignf(r0, x0), copysignf(r1, x1)), gc.half);
}
__device__ __forceinline__ void gelu2(const GeluConsts& gc, float& x0, float& x1) {            // in place
  const f32x2 x = pk2(x0, x1);
  f32x2 u, e;
  up2(mul2(x, gelu_phi2(gc, x0, x1, x, u, e)), x0, x1);
}
// GELU(x) and GELU'(x) from ONE evaluation of Phi (the fc1 epilogue stores the derivative for the backward instead of
// the pre-activation: the dgrad-fc2 epilogue then only multiplies, act = 4); both returned packed as bf16x2
__device__ __forceinline__ void gelu_and_grad2(const GeluConsts& gc, float x0, float x1, uint32_t& g_bf, uint32_t& d_bf) {
  const f32x2 x = pk2(x0, x1);
  f32x2 u, e;
  const f32x2 phi = gelu_phi2(gc, x0, x1, x, u, e);
  float g0, g1, d0, d1;
  up2(mul2(x, phi), g0, g1);
  up2(fma2(u, e, phi), d0, d1);
  g_bf = pack_bf16x2(g0, g1);
  d_bf = pack_bf16x2(d0, d1);
}
// scalar form of the derivative for the legacy act = 2 epilogue (recompute GELU' from the stored pre-activation)
__device__ __forceinline__ float gelu_grad_fast(float x) {     // Phi(x) + x phi(x)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.f)));
  const float e = exp2f(x * x * -0.72134752044448170f);            // exp(-x^2 / 2)
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float phi_cdf = 0.5f + copysignf(fmaf(-(p * t), e, 0.5f), x);
  return fmaf(x * 0.3989422804014327f, e, phi_cdf);
}

// Epilogue staging (warp-private, 32 rows x 128 B).  bf16: a row's 32 values = 4 x 16B chunks placed at slot
// (c ^ ((r>>1)&3)) + 4*((r&1)^which) so that both the row-owner writes and the 4-lanes-per-row reads are conflict-free;
// `which` = 0 / 1 picks one of two disjoint halves of the buffer, so two bf16 outputs (GELU and GELU') can be staged
// together and leave after ONE warp barrier.
__device__ __forceinline__ void stage_bf16_rows_packed(uint32_t stg, int lane, const uint32_t (&w)[16], int which) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t a = stg + lane * 128 + ((((c ^ ((lane >> 1) & 3)) + 4 * ((lane & 1) ^ which))) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]),
                 "r"(w[4 * c + 3]));
  }
}
__device__ __forceinline__ void stage_bf16_rows(uint32_t stg, int lane, const float (&v)[32], int which) {
  uint32_t w[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  stage_bf16_rows_packed(stg, lane, w, which);
}
__device__ __forceinline__ void store_bf16_coalesced(const uint8_t* stg_gen, int lane, bf16* out, long long ldo, int row0,
                                                     int n0, int M, int which) {
  const int c = lane & 3;
  uint4 x[4];      // all four shared-memory reads first: a read -> store pair per row made every store wait out a read
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = 8 * i + (lane >> 2);
    x[i] = *reinterpret_cast<const uint4*>(stg_gen + rl * 128 + (((c ^ ((rl >> 1) & 3)) + 4 * ((rl & 1) ^ which)) << 4));
  }
  bf16* p = out + (long long)(row0 + (lane >> 2)) * ldo + n0 + c * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (row0 + 8 * i + (lane >> 2) < M) *reinterpret_cast<uint4*>(p) = x[i];
    p += 8 * ldo;
  }
}

// MODE specialises the epilogue at compile time for the step's four hot forms (the generic epilogue decides everything
// per chunk from runtime fields: ~100 branches and ~120 integer instructions per 32-column chunk, which made the fc1,
// fc2-dgrad and residual GEMMs epilogue-bound and spilled the kernel out of the instruction cache):
enum EpiMode {
  EPI_GENERIC = 0,   // everything EpiParams can express
  EPI_BF16 = 1,      // alpha, bias, optional column scale -> bf16                      (qkv, plain dgrad)
  EPI_ACT3 = 2,      // bias -> GELU -> bf16, GELU' -> bf16 out2                         (Mlp.fc1 forward)
  EPI_MUL_AUX = 3,   // x bf16 aux -> bf16                                               (Mlp.fc2 input gradient)
  EPI_RES_F32 = 4,   // bias + fp32 residual -> fp32                                     (proj / fc2 forward)
  EPI_ACT1 = 5,      // bias -> GELU -> bf16                                             (Mlp.fc1, inference)
};

template <int BLOCK_N, bool A_MN, bool B_MN, bool TWO, int MODE>
__global__ void __launch_bounds__((4 + epi_warps(TWO, MODE)) * 32, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                         int K, int num_m_blocks, int num_n_blocks, int kb_per_split, int num_splits, EpiParams ep) {
  constexpr int EW = epi_warps(TWO, MODE);
  using C = Cfg<BLOCK_N, TWO, EW>;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;       // CTA of the pair; rank 0 issues the MMAs
  const int worker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t epi_stage = sB + STAGES * C::B_STAGE_BYTES;
  const uint32_t bars = epi_stage + C::EPI_STAGE_BYTES;
  const uint32_t full_bar = bars;                    // STAGES x 8B
  const uint32_t empty_bar = bars + 8 * STAGES;      // STAGES x 8B
  const uint32_t tfull_bar = bars + 16 * STAGES;     // 2 x 8B
  const uint32_t tempty_bar = tfull_bar + 16;        // 2 x 8B
  const uint32_t tmem_slot = tempty_bar + 16;        // 4B
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_units = num_tiles * num_splits;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      // a stage is released by the MMA commit and, when the A tiles are column-summed, by the two summing warps too
      mbar_init(empty_bar + 8 * s, (A_MN && B_MN && !TWO && ep.colsum_a) ? 3 : 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, EW * (TWO ? 2 : 1));   // pair: both CTAs' epilogues report to rank 0
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (TWO) tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
    else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  if (TWO) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Register split (setmaxnreg): the control warpgroup (TMA, MMA issue, TMEM allocation, wgrad column sums) gives
  // registers back, the two epilogue warpgroups take them for their one-chunk-ahead operand prefetch:
  // 128 x 64 + 256 x 216 <= the CTA's launch allocation of 384 x 168 (16 epilogue warps: 128 x 56 + 512 x 104 <= 640 x 96).
  if (warp < 4) {
  if (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = worker; unit < num_units; unit += num_workers) {
        const int split = unit / num_tiles, tile = unit - split * num_tiles;
        const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
        const int kb0 = split * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
        const int m_row = m_blk * C::TILE_M + (int)rank * BLOCK_M;          // this CTA's 128 rows of A
        const int n_row = n_blk * BLOCK_N + (int)rank * C::B_ROWS;           // this CTA's rows of B (all, or its half)
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          // pair: rank 0 expects the bytes of BOTH CTAs; every load credits rank 0's barrier
          if (ep.debug_no_loads) {
            if (!TWO || rank == 0) mbar_arrive(fb);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (!TWO || rank == 0) mbar_expect_tx(fb, (A_STAGE_BYTES + C::B_STAGE_BYTES) * (TWO ? 2 : 1));
          const uint32_t a_dst = sA + stage * A_STAGE_BYTES, b_dst = sB + stage * C::B_STAGE_BYTES;
          auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1) {
            if (TWO) tma_load_2d_2sm(dst, tm, fb, c0, c1);
            else tma_load_2d(dst, tm, fb, c0, c1);
          };
          if (!A_MN) {
            load(a_dst, &tmA, kb * BLOCK_K, m_row);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) load(a_dst + i * MN_ATOM_BYTES, &tmA, m_row + i * 64, kb * BLOCK_K);
          }
          if (!B_MN) {
            load(b_dst, &tmB, kb * BLOCK_K, n_row);
          } else {
#pragma unroll
            for (int i = 0; i < C::B_ROWS / 64; ++i) load(b_dst + i * MN_ATOM_BYTES, &tmB, n_row + i * 64, kb * BLOCK_K);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (rank 0 of a pair issues for both SMs) =====================
    constexpr uint32_t idesc = make_idesc_bf16(C::TILE_M, BLOCK_N, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int unit = worker; unit < num_units; unit += num_workers, ++it) {
      const int split = unit / num_tiles;
      const int kb0 = split * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_src = sA + stage * A_STAGE_BYTES, b_src = sB + stage * C::B_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_src + k * (UMMA_K * 128), MN_ATOM_BYTES, 1024)
                                        : make_smem_desc_sw128(a_src + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_src + k * (UMMA_K * 128), MN_ATOM_BYTES, 1024)
                                        : make_smem_desc_sw128(b_src + k * (UMMA_K * 2), 16, 1024);
            if (TWO) umma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (TWO) {
            umma_commit_2cta(empty_bar + 8 * stage);                 // frees the stage in both CTAs
            if (kb == kb1 - 1) umma_commit_2cta(tfull_bar + 8 * acc);
          } else {
            umma_commit(empty_bar + 8 * stage);
            if (kb == kb1 - 1) umma_commit(tfull_bar + 8 * acc);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if ((warp == 2 || warp == 3) && A_MN && B_MN && !TWO && ep.colsum_a != nullptr) {
    // ===================== column sums of A (wgrad only): bias gradient for free =====================
    // A stage holds A as two slabs [64 k-rows][64 m-columns] (128 B rows, 16-byte chunks XOR-swizzled by row & 7).
    // Warp 2 sums slab 0, warp 3 slab 1: lane = (row & 3 group, chunk); each lane keeps 8 fp32 column sums.  The
    // units of one (k-range, m-block) -- one per n-block, all staging the same A tiles -- share the work: unit n_blk sums
    // the k-blocks with kb % num_n_blocks == n_blk, so every A tile is summed exactly once across the grid and no unit
    // carries the whole shared-memory read load (all of it on the n_blk = 0 units cost +25 % on the GEMM).
    const int slab = warp - 2, chunk = lane & 7, rsub = lane >> 3;
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = worker; unit < num_units; unit += num_workers) {
      const int split = unit / num_tiles, tile = unit - split * num_tiles;
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      const int kb0 = split * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        if (kb % num_n_blocks == n_blk) {     // this unit's share of the k-blocks
          const uint32_t slab_base = sA + stage * A_STAGE_BYTES + slab * MN_ATOM_BYTES;
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int r = rsub + 4 * i;
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                         : "r"(slab_base + r * 128 + ((chunk ^ (r & 7)) << 4)));
            acc[0] += __uint_as_float(v0 << 16); acc[1] += __uint_as_float(v0 & 0xffff0000u);
            acc[2] += __uint_as_float(v1 << 16); acc[3] += __uint_as_float(v1 & 0xffff0000u);
            acc[4] += __uint_as_float(v2 << 16); acc[5] += __uint_as_float(v2 & 0xffff0000u);
            acc[6] += __uint_as_float(v3 << 16); acc[7] += __uint_as_float(v3 & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
          acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
        }
        const int col = m_blk * BLOCK_M + slab * 64 + chunk * 8;
        if (lane < 8 && col < M) {            // M % 8 == 0 is checked by the host
          red_add_v4(ep.colsum_a + col, acc[0], acc[1], acc[2], acc[3]);
          red_add_v4(ep.colsum_a + col + 4, acc[4], acc[5], acc[6], acc[7]);
        }
      }
    }
  }
  } else {
    if (EW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    // Software-pipelined per 32-column chunk: the TMEM load of chunk c+1 is issued as soon as chunk c has been copied
    // out of its registers, and the per-element global operands of chunk c+1 (fp32 residual, bf16 aux) are requested the
    // moment chunk c has consumed its own -- across tile boundaries too -- so neither latency is paid per chunk (with
    // two epilogue warps per scheduler there is nobody else to hide it).
    const int e = warp - 4;
    const int q = warp & 3;       // TMEM lane quarter this warp may read
    const int half = e >> 2;      // which slice of the BLOCK_N columns (2 slices with 8 warps, 4 with 16)
    constexpr int COLS_PER_WARP = BLOCK_N / (EW / 4);
    constexpr bool LD_AHEAD = EW == 8;      // with 4 warps per scheduler the other warps cover the TMEM latency
    constexpr int NCH = COLS_PER_WARP / 32;
    const uint32_t stg = epi_stage + e * 4096;
    const uint8_t* stg_gen = smem_gen + (stg - smem_base);
    const GeluConsts gc;
    // epilogue fields: compile-time constants in the specialised modes
    constexpr bool GEN = MODE == EPI_GENERIC;
    const int e_act = GEN ? ep.act : (MODE == EPI_ACT3 ? 3 : MODE == EPI_MUL_AUX ? 4 : MODE == EPI_ACT1 ? 1 : 0);
    const int e_out_mode = GEN ? ep.out_mode : (MODE == EPI_RES_F32 ? 1 : 0);
    const float* e_residual = (GEN || MODE == EPI_RES_F32) ? ep.residual : nullptr;
    const bf16* e_aux = (GEN || MODE == EPI_MUL_AUX) ? ep.aux : nullptr;
    float* e_colsum = GEN ? ep.colsum : nullptr;
    bf16* e_out2 = (GEN || MODE == EPI_ACT3) ? ep.out2 : nullptr;
    const int e_col_scale_ncols = (GEN || MODE == EPI_BF16) ? ep.col_scale_ncols : 0;
    const int e_res_row_mod = GEN ? ep.res_row_mod : 0;
    // act 4 with a plain bf16 output: multiply in the row-per-lane register domain (aux fetched coalesced and turned into
    // the row-owner layout through the staging buffer) and leave through the cheap staged bf16 store -- the "wide" float4
    // path costs ~3.5x the instructions per element and made the fc2 input-gradient GEMM epilogue-bound
    const bool narrow4 = e_act == 4 && e_out_mode == 0 && e_residual == nullptr && e_colsum == nullptr;
    const bool wide = !narrow4 && (e_out_mode != 0 || e_residual != nullptr || e_act == 2 || e_act == 4 ||
                                   e_colsum != nullptr);
    const bool wide_aux = wide && (e_act == 2 || e_act == 4);
    const int c4 = lane & 7, wr = lane >> 3;         // wide path: lane -> (row within a group of 4, float4 column)
    float4 pres[8];        // prefetched residual (wide path)
    uint32_t paux[16];     // prefetched aux: 8 x uint2 (wide path) or 4 x uint4 in the coalesced layout (narrow4)
    int pre_tag = -1;      // (unit, chunk) whose operands the prefetch registers hold
    auto prefetch = [&](int prow0, int pn0) {
      if (narrow4) {
        // coalesced: 4 lanes cover a row's 64 bytes (8 rows per request); the row-owner layout the multiply needs is
        // produced by a pass through the staging buffer (one 16-byte load per lane touching 32 different lines cost
        // the L1 four times the tag lookups)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int grow = min(prow0 + 8 * i + (lane >> 2), M - 1);
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(e_aux + (long long)grow * ep.ldaux + pn0 + (lane & 3) * 8));
          paux[4 * i] = t.x; paux[4 * i + 1] = t.y; paux[4 * i + 2] = t.z; paux[4 * i + 3] = t.w;
        }
      } else if (wide) {
        const int col = pn0 + c4 * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int grow = min(prow0 + 4 * i + wr, M - 1);          // rows clamped, no branches
          if (e_residual) {
            const int rrow = e_res_row_mod ? grow % e_res_row_mod : grow;
            pres[i] = __ldg(reinterpret_cast<const float4*>(e_residual + (long long)rrow * ep.ldr + col));
          }
          if (wide_aux) {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(e_aux + (long long)grow * ep.ldaux + col));
            paux[2 * i] = t.x; paux[2 * i + 1] = t.y;
          }
        }
      }
    };
    auto tile_row0 = [&](int unit) {
      const int tile = unit % num_tiles;
      return (tile / num_n_blocks) * C::TILE_M + (int)rank * BLOCK_M + q * 32;
    };
    auto tile_n0 = [&](int unit) {
      const int tile = unit % num_tiles;
      return (tile % num_n_blocks) * BLOCK_N + half * COLS_PER_WARP;
    };
    const bool has_pre = narrow4 || (wide && (e_residual != nullptr || wide_aux));
    if (LD_AHEAD && has_pre && worker < num_units && tile_n0(worker) < N) {
      prefetch(tile_row0(worker), tile_n0(worker));
      pre_tag = worker * NCH;
    }
    uint32_t r[32];
    // wait for accumulator stage of the it-th tile of this CTA and request this warp's first 32 columns of it
    auto first_chunk = [&](int it_) {
      const int acc_ = it_ & 1;
      mbar_wait(tfull_bar + 8 * acc_, (it_ >> 1) & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + acc_ * BLOCK_N + half * COLS_PER_WARP, r);
    };
    if (LD_AHEAD && worker < num_units) first_chunk(0);
    int it = 0;
    for (int unit = worker; unit < num_units; unit += num_workers, ++it) {
      const int acc = it & 1;
      const int row0 = tile_row0(unit);     // first of this warp's 32 rows
      const int nt0 = tile_n0(unit);
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N + half * COLS_PER_WARP;
      if (!LD_AHEAD) first_chunk(it);
      // Each chunk: TMEM -> registers in row-owner layout (lane = row, 32 consecutive columns) -> per-column math
      // -> warp-private swizzled smem transpose -> coalesced layout (a row's 64/128 B handled by 4/8 adjacent lanes)
      // -> per-element operands (residual, aux) and full-sector global stores.
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int n0 = nt0 + 32 * c;
        const bool valid = n0 < N;     // warp-uniform
        if (!LD_AHEAD && c > 0) tmem_ld_32x32b_x32(t_row + 32 * c, r);
        // operands of this chunk: already in flight, unless this is a cold start or a 16-warp kernel (no look-ahead)
        if (valid && has_pre && pre_tag != unit * NCH + c) prefetch(row0, n0);
        float4 bq[8];
        if (valid && ep.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = __ldg(b4 + j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        float v[32];
        {
          const f32x2 al = pk2(ep.alpha, ep.alpha);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            up2(fma2(pk2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), al, pk2(bq[j].x, bq[j].y)), v[4 * j],
                v[4 * j + 1]);
            up2(fma2(pk2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), al, pk2(bq[j].z, bq[j].w)),
                v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (c + 1 < NCH) {
          if (LD_AHEAD) tmem_ld_32x32b_x32(t_row + 32 * (c + 1), r);       // next chunk, in flight during this chunk's math
        } else {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (TWO && rank != 0) mbar_arrive_remote(mapa_shared(tempty_bar + 8 * acc, 0));
            else mbar_arrive(tempty_bar + 8 * acc);
          }
          // the first chunk of the next tile travels during this tile's last chunk (its accumulator is normally complete
          // already: the MMA warp runs one tile ahead of an epilogue that is the bottleneck)
          if (LD_AHEAD && unit + num_workers < num_units) first_chunk(it + 1);
        }
        if (!valid) continue;
        // where the next chunk's operands come from
        int nunit = unit, nc = c + 1;
        if (nc == NCH) { nunit = unit + num_workers; nc = 0; }
        const bool nvalid = LD_AHEAD && has_pre && nunit < num_units && tile_n0(nunit) + 32 * nc < N;
        const int nrow0 = nvalid ? tile_row0(nunit) : 0, nn0 = nvalid ? tile_n0(nunit) + 32 * nc : 0;

        if (n0 < e_col_scale_ncols) {
          const f32x2 sc = pk2(ep.col_scale, ep.col_scale);
#pragma unroll
          for (int j = 0; j < 16; ++j) up2(mul2(pk2(v[2 * j], v[2 * j + 1]), sc), v[2 * j], v[2 * j + 1]);
        }
        if (e_act == 3) {                     // out = GELU(v), out2 = GELU'(v): both staged side by side, one round trip
          uint32_t og[16], od[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) gelu_and_grad2(gc, v[2 * j], v[2 * j + 1], og[j], od[j]);
          stage_bf16_rows_packed(stg, lane, og, 0);
          if (e_out2) stage_bf16_rows_packed(stg, lane, od, 1);
          __syncwarp();
          store_bf16_coalesced(stg_gen, lane, reinterpret_cast<bf16*>(ep.out), ep.ldo, row0, n0, M, 0);
          if (e_out2) store_bf16_coalesced(stg_gen, lane, e_out2, ep.ldo2, row0, n0, M, 1);
          __syncwarp();   // staging buffer is reused by the next chunk
          continue;
        }
        if (e_out2) {      // pre-activation (act 1) or a second copy
          stage_bf16_rows(stg, lane, v, 1);
          __syncwarp();
          store_bf16_coalesced(stg_gen, lane, e_out2, ep.ldo2, row0, n0, M, 1);
          if (wide) __syncwarp();        // the fp32 staging below overwrites the whole buffer
        }
        if (e_act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) gelu2(gc, v[2 * j], v[2 * j + 1]);
        }
        if (narrow4) {
          // aux: coalesced registers -> staging half 1 -> this lane's own row
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = 8 * i + (lane >> 2), cc = lane & 3;
            const uint32_t a = stg + rl * 128 + ((((cc ^ ((rl >> 1) & 3)) + 4 * ((rl & 1) ^ 1))) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(paux[4 * i]), "r"(paux[4 * i + 1]),
                         "r"(paux[4 * i + 2]), "r"(paux[4 * i + 3]));
          }
          __syncwarp();
          if (nvalid) { prefetch(nrow0, nn0); pre_tag = nunit * NCH + nc; }      // registers are free again
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint4 a4 = *reinterpret_cast<const uint4*>(
                stg_gen + lane * 128 + ((((cc ^ ((lane >> 1) & 3)) + 4 * ((lane & 1) ^ 1))) << 4));
            const uint32_t w[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = unpack_bf16x2(w[i]);
              const int j = 4 * cc + i;
              up2(mul2(pk2(v[2 * j], v[2 * j + 1]), pk2(f.x, f.y)), v[2 * j], v[2 * j + 1]);
            }
          }
        }
        if (!wide) {
          stage_bf16_rows(stg, lane, v, 0);
          __syncwarp();
          store_bf16_coalesced(stg_gen, lane, reinterpret_cast<bf16*>(ep.out), ep.ldo, row0, n0, M, 0);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                         "f"(v[4 * j + 2]), "f"(v[4 * j + 3]));
          }
          __syncwarp();
          const int col = n0 + c4 * 4;
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + wr;
            x[i] = *reinterpret_cast<const float4*>(stg_gen + rl * 128 + ((c4 ^ (rl & 7)) << 4));
            if (e_act == 2) {
              const float2 p0 = unpack_bf16x2(paux[2 * i]), p1 = unpack_bf16x2(paux[2 * i + 1]);
              x[i].x *= gelu_grad_fast(p0.x); x[i].y *= gelu_grad_fast(p0.y);
              x[i].z *= gelu_grad_fast(p1.x); x[i].w *= gelu_grad_fast(p1.y);
            } else if (e_act == 4) {          // aux already holds GELU'(pre-activation)
              const float2 p0 = unpack_bf16x2(paux[2 * i]), p1 = unpack_bf16x2(paux[2 * i + 1]);
              x[i].x *= p0.x; x[i].y *= p0.y; x[i].z *= p1.x; x[i].w *= p1.y;
            }
            if (e_residual) { x[i].x += pres[i].x; x[i].y += pres[i].y; x[i].z += pres[i].z; x[i].w += pres[i].w; }
          }
          // this chunk's operands are consumed: request the next chunk's before the stores go out
          if (nvalid) { prefetch(nrow0, nn0); pre_tag = nunit * NCH + nc; }
          float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int grow = row0 + 4 * i + wr;
            if (grow < M) {
              cs.x += x[i].x; cs.y += x[i].y; cs.z += x[i].z; cs.w += x[i].w;
              if (e_out_mode == 0) {
                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + (long long)grow * ep.ldo + col) =
                    make_uint2(pack_bf16x2(x[i].x, x[i].y), pack_bf16x2(x[i].z, x[i].w));
              } else if (e_out_mode == 1) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (long long)grow * ep.ldo + col) = x[i];
              } else {
                red_add_v4(reinterpret_cast<float*>(ep.out) + (long long)grow * ep.ldo + col, x[i].x, x[i].y, x[i].z, x[i].w);
              }
            }
          }
          if (e_colsum) {     // lanes l, l+8, l+16, l+24 hold the same 4 columns: fold, then one vector atomic per column group
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
            if (lane < 8) red_add_v4(e_colsum + col, cs.x, cs.y, cs.z, cs.w);
          }
        }
        __syncwarp();   // staging buffer is reused by the next chunk
      }
    }
  }

  tc_fence_before();
  if (TWO) cluster_sync_all();      // the peer may still be reading this CTA's smem / signalling its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool TWO, int MODE = EPI_GENERIC>
int launch(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int splits,
           const EpiParams& ep, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, TWO, epi_warps(TWO, MODE)>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, BLOCK_M, BLOCK_K);
  else       rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, BLOCK_K, 64);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d_bf16(&tmB, B, N, K, ldb, C::B_ROWS, BLOCK_K);
  else       rc = make_tmap_2d_bf16(&tmB, B, K, N, ldb, BLOCK_K, 64);
  if (rc) return rc;
  const int num_m_blocks = (M + C::TILE_M - 1) / C::TILE_M, num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  splits = max(1, min(splits, num_kb));
  const int kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per_split - 1) / kb_per_split;  // no empty splits
  const int units = num_m_blocks * num_n_blocks * splits;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, A_MN, B_MN, TWO, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  const int workers = min(units, TWO ? num_sms() / 2 : num_sms());
  cfg.gridDim = dim3(TWO ? 2 * workers : workers);
  cfg.blockDim = dim3(C::NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  if (TWO) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  EGOVLP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, N, K, num_m_blocks, num_n_blocks, kb_per_split, splits, ep));
  return EGOVLP_OK;
}

template <int BLOCK_N, bool TWO>
int dispatch_major(int a_mn, int b_mn, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                   int splits, const EpiParams& ep, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch<BLOCK_N, false, false, TWO>(A, lda, B, ldb, M, N, K, splits, ep, stream);
  if (!a_mn && b_mn) return launch<BLOCK_N, false, true, TWO>(A, lda, B, ldb, M, N, K, splits, ep, stream);
  if (a_mn && b_mn) return launch<BLOCK_N, true, true, TWO>(A, lda, B, ldb, M, N, K, splits, ep, stream);
  return launch<BLOCK_N, true, false, TWO>(A, lda, B, ldb, M, N, K, splits, ep, stream);
}

// Which specialised epilogue (if any) computes exactly what `ep` asks for; EGOVLP_GEMM_GENERIC_EPI=1 keeps every call on
// the generic one (the kernel tests run both and compare).
inline int epi_mode(const EpiParams& ep) {
  const char* g = getenv("EGOVLP_GEMM_GENERIC_EPI");
  if (g && g[0] == '1') return EPI_GENERIC;
  if (ep.colsum || ep.res_row_mod) return EPI_GENERIC;
  const bool no_scale = ep.col_scale_ncols == 0;
  if (ep.act == 0 && ep.out_mode == 0 && !ep.residual && !ep.out2) return EPI_BF16;
  if (ep.act == 3 && ep.out_mode == 0 && !ep.residual && ep.out2 && no_scale) return EPI_ACT3;
  if (ep.act == 4 && ep.out_mode == 0 && !ep.residual && !ep.out2 && no_scale) return EPI_MUL_AUX;
  if (ep.act == 0 && ep.out_mode == 1 && ep.residual && !ep.out2 && no_scale) return EPI_RES_F32;
  if (ep.act == 1 && ep.out_mode == 0 && !ep.residual && !ep.out2 && no_scale) return EPI_ACT1;
  return EPI_GENERIC;
}

// EGOVLP_GEMM_1CTA=1 keeps every shape on the single-CTA kernels (tests exercise both)
inline bool force_one_cta() {
  const char* e = getenv("EGOVLP_GEMM_1CTA");
  return e && e[0] == '1';
}

}  // namespace

}  // namespace egovlp

using namespace egovlp;

extern "C" int egovlp_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                                long long ldb, int M, int N, int K, const egovlp_gemm_epilogue* e, int split_k,
                                void* stream) {
  EGOVLP_CHECK_ARG(A && B && e && e->out, "gemm: null pointer");
  EGOVLP_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  EGOVLP_CHECK_ARG(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  EGOVLP_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: leading dimensions must be multiples of 8 (16B TMA strides)");
  EGOVLP_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                   "gemm: operands must be 16B aligned");
  EGOVLP_CHECK_ARG(e->out_mode >= 0 && e->out_mode <= 2 && e->act >= 0 && e->act <= 4, "gemm: bad epilogue mode");
  EGOVLP_CHECK_ARG(split_k <= 1 || e->out_mode == 2, "gemm: split_k > 1 needs out_mode=2 (fp32 atomic accumulate)");
  EGOVLP_CHECK_ARG((e->act != 2 && e->act != 4) || e->aux, "gemm: act=2/4 needs aux");
  EGOVLP_CHECK_ARG(e->ldo % 8 == 0, "gemm: ldo must be a multiple of 8");
  EpiParams ep;
  ep.bias = e->bias; ep.residual = e->residual; ep.aux = reinterpret_cast<const bf16*>(e->aux);
  ep.out = e->out; ep.out2 = reinterpret_cast<bf16*>(e->out2);
  ep.ldr = e->ldr; ep.ldaux = e->ldaux; ep.ldo = e->ldo; ep.ldo2 = e->ldo2;
  ep.out_mode = e->out_mode; ep.act = e->act; ep.alpha = e->alpha;
  ep.col_scale = e->col_scale; ep.col_scale_ncols = e->col_scale_ncols; ep.res_row_mod = e->res_row_mod;
  ep.colsum = e->colsum;
  ep.colsum_a = e->colsum_a;
  EGOVLP_CHECK_ARG(!e->colsum_a || (a_mn_major && b_mn_major && N % 256 == 0 && M % 8 == 0 &&
                                    (reinterpret_cast<uintptr_t>(e->colsum_a) & 15) == 0),
                   "gemm: colsum_a needs the MN/MN (wgrad) form with N % 256 == 0, M % 8 == 0 and a 16B-aligned vector");
  {
    const char* dbg = getenv("EGOVLP_GEMM_DEBUG_NOLOADS");
    ep.debug_no_loads = (dbg && dbg[0] == '1') ? 1 : 0;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // CTA-pair 256 x 256 tiles for the K-major-A shapes (fwd, dgrad); the token-contraction wgrad (both operands
  // MN-major, split-K) measured 3-8 % faster on single-CTA 128 x 256 tiles (tools/bench_gemm_modes.py)
  if (N % 256 == 0 && !force_one_cta() && !(a_mn_major && b_mn_major)) {
    if (!a_mn_major) {       // the step's hot forms get a compile-time specialised epilogue
      const int mode = epi_mode(ep);
#define EGOVLP_GEMM_MODE(MODE)                                                                               \
  return b_mn_major ? launch<256, false, true, true, MODE>(A, lda, B, ldb, M, N, K, split_k, ep, st)          \
                    : launch<256, false, false, true, MODE>(A, lda, B, ldb, M, N, K, split_k, ep, st)
      if (mode == EPI_BF16) { EGOVLP_GEMM_MODE(EPI_BF16); }
      if (mode == EPI_ACT3) { EGOVLP_GEMM_MODE(EPI_ACT3); }
      if (mode == EPI_MUL_AUX) { EGOVLP_GEMM_MODE(EPI_MUL_AUX); }
      if (mode == EPI_RES_F32) { EGOVLP_GEMM_MODE(EPI_RES_F32); }
      if (mode == EPI_ACT1) { EGOVLP_GEMM_MODE(EPI_ACT1); }
#undef EGOVLP_GEMM_MODE
    }
    return dispatch_major<256, true>(a_mn_major, b_mn_major, A, lda, B, ldb, M, N, K, split_k, ep, st);
  }
  if (N % 256 == 0) return dispatch_major<256, false>(a_mn_major, b_mn_major, A, lda, B, ldb, M, N, K, split_k, ep, st);
  return dispatch_major<128, false>(a_mn_major, b_mn_major, A, lda, B, ldb, M, N, K, split_k, ep, st);
}
