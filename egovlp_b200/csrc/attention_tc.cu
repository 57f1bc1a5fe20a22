// Space attention forward on tcgen05 / TMEM (the N-length attention of VarAttention, model/video_transformer.py:
// 109-133 in '(b f) n d' mode): S = Q K^T and O = P V are UMMA tiles, Q/K/V arrive by TMA, the softmax runs on
// the TMEM accumulator rows.  Same inputs / outputs / CLS semantics as the mma.sync kernels in attention.cu.
//
// One persistent CTA per SM loops over the (b, head, frame) groups with double-buffered Q/K/V tiles:
//   warp 0   TMA producer: per group 3 boxes of N patch rows + 3 one-row boxes for the CLS q/k/v (row N of each tile)
//   warp 1   MMA issuer:   S_t[128 x NKP] = Q_t K^T for the two 128-row query tiles (4 UMMAs each, K = 64), then, as
//            soon as the softmax warps have written P_t (bf16, 128B-swizzled K-major tile in smem),
//            O[128 x 64] = P_t V (NKP/16 UMMAs, V read as an MN-major B operand straight from the [key, d] tile)
//   warp 2   TMEM allocator (512 columns: S_0 | S_1 | O)
//   warps 4-7 one thread per query row: two passes over the S row in TMEM (max; exp2 / sum / bf16 P), then the
//            O row: normalise, store 128 contiguous bytes; the CLS query row leaves its (max, sum, acc) partial.
// Rows >= N+1 of a tile are padding: their results are never stored (UMMA rows are independent).
#include <stdlib.h>

#include "common.cuh"
#include "egovlp_b200.h"

namespace egovlp {
namespace {

constexpr int HD = 64;
constexpr int ROWB = 128;                 // bytes per head-row
constexpr float LOG2E = 1.4426950408889634f;
constexpr int TILE_ROWS = 208;            // rows reserved per Q/K/V tile (NKP <= 208 -> 26 KB, 1024-aligned)
constexpr int TILE_BYTES = TILE_ROWS * ROWB;
constexpr int P_BLOCK_BYTES = 128 * ROWB; // one 64-key column block of P for 128 rows
constexpr int P_BYTES = 4 * P_BLOCK_BYTES;
constexpr int TC_THREADS = 384;             // TMA, MMA, TMEM-alloc, spare + 8 softmax warps
constexpr int S1_COL = 224, O_COL = 448;  // TMEM columns: S_0 at 0, S_1 at 224, O at 448 (NKP <= 224)

// Bounded mbarrier wait without a printf call site (a call makes every live register of the caller spill around it) and
// with the clock looked at only every 1024 polls: a waiting warp costs almost no issue slots.
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 1024; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a protocol bug fails the launch instead of hanging the GPU
  }
}
#define mbar_wait mbar_wait_hot

struct TcGeom {
  int B, H, T, N, S, D, NK, NKP, groups;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
space_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_cls,
                         bf16* __restrict__ out, float* __restrict__ lse_out, float* __restrict__ cls_part, TcGeom G) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // [buf0: Q K V][buf1: Q K V][P][barriers]
  const uint32_t sP = base + 6 * TILE_BYTES;
  const uint32_t bars = sP + P_BYTES;
  const uint32_t full_bar = bars;            // 2
  const uint32_t empty_bar = bars + 16;      // 2
  const uint32_t sfull_bar = bars + 32;      // 2 (per query tile)
  const uint32_t sfree_bar = bars + 48;      // 2
  const uint32_t pready_bar = bars + 64;
  const uint32_t pfree_bar = bars + 72;
  const uint32_t ofull_bar = bars + 80;      // 2 (per query tile)
  const uint32_t ofree_bar = bars + 96;      // O_0 region free (O_1 aliases S_0: released through sfree[0])
  const uint32_t tmem_slot = bars + 104;
  const uint32_t xchg = bars + 128;          // 2 x [2][128] floats: row max / row sum exchange between column halves
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NT = 2;                       // query tiles (the host only routes 128 < N+1 <= 208 here)

  // zero the padding rows NK .. TILE_ROWS-1 of every tile once (TMA never writes them): V pad rows must be finite
  for (int i = threadIdx.x; i < 6 * (TILE_ROWS - G.NK) * 8; i += TC_THREADS) {
    const int c = i & 7, rr = (i >> 3) % (TILE_ROWS - G.NK), tile = (i >> 3) / (TILE_ROWS - G.NK);
    *reinterpret_cast<uint4*>(gen + tile * TILE_BYTES + (G.NK + rr) * ROWB + c * 16) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_rows);
    tma_prefetch_desc(&tm_cls);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
      mbar_init(sfull_bar + 8 * i, 1);
      mbar_init(sfree_bar + 8 * i, 8);
    }
    mbar_init(pready_bar, 8);
    mbar_init(pfree_bar, 1);
    mbar_init(ofull_bar, 1);
    mbar_init(ofull_bar + 8, 1);
    mbar_init(ofree_bar, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();       // the zero fill above must be visible to the UMMA (async proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int it = 0;
      for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++it) {
        const int buf = it & 1;
        const int f = g % G.T, h = (g / G.T) % G.H, b = g / (G.T * G.H);
        mbar_wait(empty_bar + 8 * buf, ((it >> 1) & 1) ^ 1);
        const uint32_t fb = full_bar + 8 * buf;
        mbar_expect_tx(fb, 3u * (uint32_t)G.NK * ROWB);
        const uint32_t q = base + buf * 3 * TILE_BYTES;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          tma_load_4d(q + w * TILE_BYTES, &tm_rows, fb, 0, 1 + f * G.N, w * G.H + h, b);
          tma_load_4d(q + w * TILE_BYTES + G.N * ROWB, &tm_cls, fb, 0, 0, w * G.H + h, b);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc_s = make_idesc_bf16(128, G.NKP, false, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, false, true);     // B = V, MN-major (d contiguous)
    int it = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t q = base + buf * 3 * TILE_BYTES, k = q + TILE_BYTES, v = k + TILE_BYTES;
      mbar_wait(full_bar + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      for (int t = 0; t < NT; ++t) {
        mbar_wait(sfree_bar + 8 * t, (it & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_smem_desc_sw128(q + t * 128 * ROWB + kk * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(k + kk * 32, 16, 1024);
            umma_bf16_ss(tmem + (t ? S1_COL : 0), ad, bd, idesc_s, kk > 0);
          }
          umma_commit(sfull_bar + 8 * t);
        }
        __syncwarp();
      }
      for (int t = 0; t < NT; ++t) {
        const int n = it * NT + t;
        mbar_wait(pready_bar, n & 1);
        if (t == 0) mbar_wait(ofree_bar, (it & 1) ^ 1);         // O_0 of the previous group has been read
        tc_fence_after();
        if (lane == 0) {
          // O_0 has its own columns; O_1 reuses the first columns of S_0, which every softmax warp read long ago
          // (they signalled P_1 after it); the next group's S_0 waits for the O_1 epilogue through sfree[0]
          const uint32_t o_tmem = tmem + (t == 0 ? O_COL : 0);
          for (int ks = 0; ks < G.NKP / 16; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(sP + (ks >> 2) * P_BLOCK_BYTES + (ks & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(v + ks * 16 * ROWB, 8192, 1024);
            umma_bf16_ss(o_tmem, ad, bd, idesc_o, ks > 0);
          }
          umma_commit(pfree_bar);
          umma_commit(ofull_bar + 8 * t);
          if (t == NT - 1) umma_commit(empty_bar + 8 * buf);     // every UMMA reading this Q/K/V buffer is done
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ======================= softmax + epilogue: TWO threads per query row =======================
    // warps 4-7 take the first half of the key columns, warps 8-11 the second half of the same TMEM lanes; each
    // thread keeps its half row (up to 112 fp32) in registers, so S is read from TMEM once; the two halves exchange
    // their max / sum through smem and a 64-thread named barrier.
    const int qd = warp & 3, half = (warp - 4) >> 2;
    const int r_in_tile = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const int HW = ((G.NKP / 8 + 1) / 2) * 8;             // columns per half, multiple of 8 (104 for 208)
    const int col0 = half * HW, ncol = half ? G.NKP - HW : HW;
    float* xmax = reinterpret_cast<float*>(gen + (xchg - base));      // [2][128]
    float* xsum = xmax + 256;
    int it = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++it) {
      const int f = g % G.T, h = (g / G.T) % G.H, b = g / (G.T * G.H);
      float mx0 = 0.f, mx1 = 0.f, sum0 = 0.f, sum1 = 0.f;
      // ---- softmax of both query tiles first; their P V products run underneath, the O rows are read afterwards
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int n = it * 2 + t;
        const int row = t * 128 + r_in_tile;             // query row of the group (N = CLS query)
        // the CLS key (column N) is visible to every patch query, and to the CLS query in the first frame only
        const int vis = (row == G.N && f != 0) ? G.N : G.NK;
        const uint32_t s_addr = tmem + lane_base + (t ? S1_COL : 0) + col0;
        mbar_wait(sfull_bar + 8 * t, it & 1);
        tc_fence_after();
        uint32_t r[112];
        tmem_ld_32x32b_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld_32x32b_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_32x32b_x32(s_addr + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
        tmem_ld_32x32b_x16(s_addr + 96, *reinterpret_cast<uint32_t(*)[16]>(&r[96]));
        tmem_ld_wait();
        if (t == 1) {           // S_1 is free again (S_0's columns stay reserved: O_1 lands there)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sfree_bar + 8);
        }
        // half 0 holds patch keys only (dense: no predicates, packed FFMA2 / FADD2 math); half 1 ends with the CLS key
        // and the padding columns (masked per element).  Padding ROWS compute garbage-but-finite values nobody stores.
        float mx = -INFINITY;
        if (half == 0) {
#pragma unroll
          for (int j = 0; j < 112; ++j)
            if (j < ncol) mx = fmaxf(mx, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 112; ++j) {
            const bool ok = j < ncol && col0 + j < vis;
            r[j] = ok ? r[j] : 0xff800000u;
            mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
        xmax[half * 128 + r_in_tile] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        mx = fmaxf(mx, xmax[(half ^ 1) * 128 + r_in_tile]);
        const float ms = mx * LOG2E;
        const f32x2 nms = pk2(-ms, -ms), l2e = pk2(LOG2E, LOG2E);
        mbar_wait(pfree_bar, (n & 1) ^ 1);               // the previous P V product has consumed the P tile
        f32x2 sum2 = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 14; ++c) {
          if (8 * c < ncol) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float e0, e1;
              up2(fma2(pk2(__uint_as_float(r[8 * c + 2 * j]), __uint_as_float(r[8 * c + 2 * j + 1])), l2e, nms), e0, e1);
              e0 = exp2f(e0);                            // masked entries are -inf -> exp2(-inf) = 0
              e1 = exp2f(e1);
              sum2 = add2(sum2, pk2(e0, e1));
              pk[j] = pack_bf16x2(e0, e1);
            }
            const int col = col0 + 8 * c;
            const uint32_t a = sP + (col >> 6) * P_BLOCK_BYTES + r_in_tile * ROWB + ((((col & 63) >> 3) ^ (r_in_tile & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]));
          }
        }
        float sum, sum_hi;
        up2(sum2, sum, sum_hi);
        sum += sum_hi;
        xsum[half * 128 + r_in_tile] = sum;
        tc_fence_before();
        fence_proxy_async_smem();                        // P (generic-proxy stores) -> visible to the UMMA reads
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        sum += xsum[(half ^ 1) * 128 + r_in_tile];
        if (lane == 0) mbar_arrive(pready_bar);
        if (t == 0) { mx0 = mx; sum0 = sum; } else { mx1 = mx; sum1 = sum; }
      }
      // ---- epilogues: this thread's 32 columns of each O row
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int row = t * 128 + r_in_tile;
        mbar_wait(ofull_bar + 8 * t, it & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem + lane_base + (t == 0 ? O_COL : 0) + half * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t == 0 ? ofree_bar : sfree_bar);   // O_1 read -> S_0's columns are free again
        const float mx = t ? mx1 : mx0, sum = t ? sum1 : sum0;
        if (row < G.N) {
          const float inv = 1.f / sum;
          const long long tok = (long long)b * G.S + 1 + f * G.N + row;
          uint4* dst = reinterpret_cast<uint4*>(out + tok * G.D + h * HD + half * 32);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[c] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv),
                                pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv),
                                pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv),
                                pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv));
          if (half == 0) lse_out[((long long)(b * G.H + h)) * G.S + 1 + f * G.N + row] = mx + logf(sum);
        } else if (row == G.N) {
          float* dst = cls_part + (((long long)(b * G.H + h)) * G.T + f) * 66;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[half * 32 + j] = __uint_as_float(o[j]);
          if (half == 0) { dst[64] = mx; dst[65] = sum; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// geometry the tcgen05 kernel covers: two query tiles, keys padded to <= 208
// Opt-in (EGOVLP_ATTN_TC=1).  Measured at B=16, T=16, N=196, H=12 (tools/bench_misc.py): this kernel 0.185 ms vs
// 0.163 ms for the mma.sync span kernel -- at head_dim 64 and 197 keys the tcgen05 formulation is bound by
// TMEM -> register reads of S (277 KB per group at ~64 B/clk/SM) and MUFU exp2, not by the MMAs, while mma.sync
// keeps S in registers.  The default dispatch therefore stays on the span kernel; see DESIGN.md section 7.
bool space_attn_tc_supported(int N) {
  const char* e = getenv("EGOVLP_ATTN_TC");
  if (!(e && e[0] == '1')) return false;
  const int NK = N + 1;
  return NK > 128 && NK <= TILE_ROWS;
}

int space_attn_fwd_tc(const void* qkv, void* out, float* lse, float* cls_part, int B, int T, int N, int H,
                      cudaStream_t st) {
  TcGeom G;
  G.B = B; G.H = H; G.T = T; G.N = N; G.S = 1 + T * N; G.D = H * HD; G.NK = N + 1; G.NKP = (G.NK + 15) / 16 * 16;
  G.groups = B * H * T;
  const uint64_t W = 3ull * G.D;
  const uint64_t dims[4] = {HD, (uint64_t)G.S, (uint64_t)(3 * H), (uint64_t)B};
  const uint64_t strides[4] = {1, W, HD, (uint64_t)G.S * W};
  const uint32_t box_rows[4] = {HD, (uint32_t)N, 1, 1};
  const uint32_t box_cls[4] = {HD, 1, 1, 1};
  CUtensorMap tm_rows, tm_cls;
  int rc = make_tmap_nd_bf16(&tm_rows, qkv, 4, dims, strides, box_rows, true);
  if (rc) return rc;
  rc = make_tmap_nd_bf16(&tm_cls, qkv, 4, dims, strides, box_cls, true);
  if (rc) return rc;
  const int smem = 6 * TILE_BYTES + P_BYTES + 128 + 2048 + 1024;
  static bool attr = false;
  if (!attr) {
    EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(space_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int grid = G.groups < num_sms() ? G.groups : num_sms();
  space_attn_fwd_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tm_rows, tm_cls, reinterpret_cast<bf16*>(out), lse, cls_part, G);
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}

}  // namespace egovlp
