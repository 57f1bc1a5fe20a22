// Space attention BACKWARD on tcgen05 / TMEM (the N-length attention of VarAttention in '(b f) n d' mode,
// model/video_transformer.py:109-133, differentiated): all five contractions of a (b, head, frame) group
//     S = Q K^T     dP = dO V^T     dV = P^T dO     dK = dS^T Q     dQ = dS K          (P = exp(S - lse), dS = P o (dP - delta))
// are UMMA tiles with TMEM accumulators; Q / K / V / dO arrive by TMA, P and dS go through 128B-swizzled shared memory
// ONCE and serve three operand roles (the same bytes are a K-major A tile for dQ and an MN-major A tile for dV / dK), so
// every contraction is computed exactly once per group -- the mma.sync kernel in attention.cu recomputes S and dP in
// both of its phases (7 GEMM units instead of 5).  Same inputs / outputs / CLS semantics as fast_attn_bwd_kernel<false>.
//
// Group rows: 0..N-1 = the frame's patches, row N = the CLS token (as a key for every patch query; as a query it sees the
// patch keys of the frame and, in frame 0 only, the CLS key -- its lse / delta are the global ones), rows N+1..NKP-1
// zero padding.  One persistent CTA per SM loops over groups, each in four (key tile kt, query tile qt) steps of a
// 128 x 128 block (second tiles are 128 x (NKP-128)):
//   warp 0     TMA producer: a ring of 6 tile slots [NKP rows x 128 B]; Q, K, dO, V of a group (patch rows + the CLS row
//              box each), so the next group's Q / K are in flight while this one computes
//   warp 1     MMA issuer (one thread): S and dP of the next step are issued as soon as the softmax warps have pulled the
//              current ones out of TMEM; dV after P is in smem, dK / dQ after dS
//   warp 2     TMEM allocator: 512 columns = S 128 | dP 128 | dQ_0 64 | dQ_1 64 | dK 64 | dV 64
//   warps 4-11 softmax: two threads per query row (64 key columns each): TMEM -> registers, P = exp2(S log2e - lse2) ->
//              smem (packed FFMA2 + MUFU, no predicates on the dense first key tile), dS = P (dP - delta) -> smem
//   warps 12-15 drain warpgroup: per-row lse2 / delta = rowsum(dO o O) of the NEXT group from global memory (double
//              buffered), and the accumulator drains (dK / dV per key tile, dQ per group) straight to dqkv as 64-byte row
//              pieces, CLS-row gradients by fp32 atomics -- so the softmax warps never wait for a drain or a prologue.
// setmaxnreg splits the CTA's registers 72 / 176 / 88 per thread between the control, softmax and drain warpgroups.
// (16 softmax warps -- four threads per row -- were measured SLOWER: 1.82 ms vs 1.46 ms at B = 64.)
#include <stdlib.h>

#include "common.cuh"
#include "egovlp_b200.h"

namespace egovlp {
namespace {

constexpr int HD = 64;
constexpr int ROWB = 128;                  // bytes per head row
constexpr float LOG2E = 1.4426950408889634f;
constexpr int TILE_ROWS = 208;
constexpr int TILE_BYTES = TILE_ROWS * ROWB;          // 26 KB, 1024-aligned
constexpr int NSLOT = 6;
constexpr int PB_BYTES = 2 * 128 * ROWB;              // [2 key blocks of 64][128 query rows][128 B]
constexpr int KBLK_BYTES = 128 * ROWB;                // one 64-key block
constexpr int LSD_FLOATS = 2 * 2 * TILE_ROWS;         // [parity][lse2 | delta][row]
constexpr int THREADS = 512;                          // warpgroups: control (TMA / MMA / TMEM) | softmax x2 | drain
constexpr int S_COL = 0, DP_COL = 128, DQ_COL = 256, DK_COL = 384, DV_COL = 448;

struct BwdGeom {
  int B, H, T, N, S, D, NK, NKP, W1, groups;        // W1 = NKP - 128: width of the second key / query tile
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(r));
}
__device__ __forceinline__ void st_shared_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// Bounded mbarrier wait for the hot roles: try_wait suspends in hardware; the clock is only looked at every 1024
// polls (a protocol bug traps the launch after a few seconds instead of hanging), so a waiting warp costs almost no issue slots (the generic mbar_wait spends ~8 instructions per poll, which
// showed up as 18 M instructions each for the TMA and the MMA warp in the first profile of this kernel).
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 1024; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s.  No printf: a call site makes every live register spill around it
  }
}
#define mbar_wait mbar_wait_hot

// Optional timeline (EGOVLP_ATTN_BWD_TRACE=<file>): CTA 0 stamps (event, group, step, clock) of its first groups into a
// device buffer that the host dumps after the launch -- the tool the pipeline of this kernel was tuned with.
struct Tracer {                                          // per-role private cursor: plain stores, no atomics
  unsigned long long* p;
  int n;
};
__device__ __forceinline__ Tracer make_tracer(unsigned long long* tr, int role, bool on) {
  Tracer t;
  t.p = (tr != nullptr && blockIdx.x == 0 && on) ? tr + 1 + role * 1000 : nullptr;
  t.n = 0;
  return t;
}
__device__ __forceinline__ void trace_ev(Tracer& t, int ev, int gi, int it) {
  if (t.p != nullptr && t.n < 1000) {
    t.p[t.n++] = ((unsigned long long)(ev & 0xff) << 56) | ((unsigned long long)(gi & 0xff) << 48) |
                 ((unsigned long long)(it & 0xf) << 44) | (clock64() & 0xfffffffffffull);
  }
}

// arrive (one per warp) after this warp's TMEM reads / generic-proxy smem writes are complete
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

struct SoftmaxArgs {
  float lse2, delta;
  uint32_t s_addr, dp_addr, p_row, ds_row, sw;
  int col0, ncol, vis_cols, lane;
  uint32_t parity, s_full, s_free, p_freeb, p_ready, dp_full, ds_freeb;
  bool active;
  Tracer* trc;
  int gi, it;
};

// One (key tile, query tile) step of a softmax thread: S (TMEM) -> P = exp2(S log2e - lse2) -> bf16 smem, then
// dP (TMEM) -> dS = P (dP - delta) -> bf16 smem.  MASKED = the short second key tile (<= 40 columns per thread, CLS key
// and padding keys masked per element); the first key tile is dense: 64 columns, packed FFMA2 / FMUL2 / FADD2 math and no
// predicates (the first profile of this kernel was instruction-issue bound: 12 instructions per element in this loop).
template <bool MASKED>
__device__ __forceinline__ void softmax_step(const SoftmaxArgs& A) {
  constexpr int NCH = MASKED ? 5 : 8;                    // 8-column chunks handled by a thread
  const f32x2 nl2 = pk2(-A.lse2, -A.lse2), l2e = pk2(LOG2E, LOG2E), ndel = pk2(-A.delta, -A.delta);
  mbar_wait(A.s_full, A.parity);
  tc_fence_after();
  trace_ev(*A.trc, 20, A.gi, A.it);               // softmax: s_full seen
  mbar_wait(A.p_freeb, A.parity ^ 1);                    // the previous step's dV has consumed the P buffer (long ago)
  {
    uint32_t sv[64];
    if (A.active) {
      // always two wide loads (columns past ncol are stale TMEM, never used): narrow tcgen05.ld shapes pay a fixed
      // per-instruction cost that dominated the first version of this kernel
      tmem_ld32(A.s_addr, sv);
      tmem_ld32(A.s_addr + 32, sv + 32);
      tmem_ld_wait();
    }
    warp_arrive(A.s_free, A.lane);
    trace_ev(*A.trc, 21, A.gi, A.it);             // softmax: S in registers
    if (A.active) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (!MASKED || 8 * c < A.ncol) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float e0, e1;
            up2(fma2(pk2(__uint_as_float(sv[8 * c + 2 * j]), __uint_as_float(sv[8 * c + 2 * j + 1])), l2e, nl2), e0, e1);
            e0 = exp2f(e0);
            e1 = exp2f(e1);
            if (MASKED) {
              e0 = 8 * c + 2 * j < A.vis_cols ? e0 : 0.f;
              e1 = 8 * c + 2 * j + 1 < A.vis_cols ? e1 : 0.f;
            }
            pk[j] = pack_bf16x2(e0, e1);
          }
          const int col = A.col0 + 8 * c;
          st_shared_v4(A.p_row + (col >> 6) * KBLK_BYTES + ((((col & 63) >> 3) ^ A.sw) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
  }
  fence_proxy_async_smem();                              // generic-proxy stores -> visible to the UMMA reads
  warp_arrive(A.p_ready, A.lane);
  trace_ev(*A.trc, 22, A.gi, A.it);               // softmax: P written
  // dP -> dS = P (dP - delta), 32 columns at a time; P is read back from this thread's own smem row (bf16, exactly what
  // dV consumes) instead of being held in 32 registers across the wait.  Masked / padded elements have P == 0 and a
  // finite dP, so no predicate is needed here.
  mbar_wait(A.dp_full, A.parity);
  tc_fence_after();
  mbar_wait(A.ds_freeb, A.parity ^ 1);                   // the previous step's dK / dQ have consumed the dS buffer
  trace_ev(*A.trc, 23, A.gi, A.it);               // softmax: dp_full + ds_free seen
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    if (A.active && (!MASKED || 32 * hh < A.ncol)) {
      uint32_t dp[32];
      tmem_ld32(A.dp_addr + 32 * hh, dp);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (4 * hh + c < NCH && (!MASKED || 32 * hh + 8 * c < A.ncol)) {
          const int col = A.col0 + 32 * hh + 8 * c;
          const uint32_t off = (col >> 6) * KBLK_BYTES + ((((col & 63) >> 3) ^ A.sw) << 4);
          uint32_t pk[4], dsp[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3]) : "r"(A.p_row + off));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float d0, d1;
            up2(mul2(pk2(__uint_as_float(pk[j] << 16), __uint_as_float(pk[j] & 0xffff0000u)),
                     add2(pk2(__uint_as_float(dp[8 * c + 2 * j]), __uint_as_float(dp[8 * c + 2 * j + 1])), ndel)), d0, d1);
            dsp[j] = pack_bf16x2(d0, d1);
          }
          st_shared_v4(A.ds_row + off, dsp[0], dsp[1], dsp[2], dsp[3]);
        }
      }
    }
  }
}

// 32 fp32 accumulator values of one row -> 64 contiguous bytes of bf16
__device__ __forceinline__ void store_row32(const uint32_t (&v)[32], float scale, bf16* dst) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int c = 0; c < 4; ++c)
    d[c] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * c]) * scale, __uint_as_float(v[8 * c + 1]) * scale),
                      pack_bf16x2(__uint_as_float(v[8 * c + 2]) * scale, __uint_as_float(v[8 * c + 3]) * scale),
                      pack_bf16x2(__uint_as_float(v[8 * c + 4]) * scale, __uint_as_float(v[8 * c + 5]) * scale),
                      pack_bf16x2(__uint_as_float(v[8 * c + 6]) * scale, __uint_as_float(v[8 * c + 7]) * scale));
}
__device__ __forceinline__ void atomic_row32(const uint32_t (&v)[32], float* dst) {
#pragma unroll
  for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
}

__global__ void __launch_bounds__(THREADS, 1)
space_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_cls,
                         const __grid_constant__ CUtensorMap tm_do_rows, const __grid_constant__ CUtensorMap tm_do_cls,
                         const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse_in,
                         bf16* __restrict__ dqkv, float* __restrict__ dcls, float q_scale, BwdGeom G,
                         unsigned long long* __restrict__ trace) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sP = base + NSLOT * TILE_BYTES;
  const uint32_t sDS = sP + PB_BYTES;
  const uint32_t lsd_off = NSLOT * TILE_BYTES + 2 * PB_BYTES;
  float* lsd = reinterpret_cast<float*>(gen + lsd_off);
  const uint32_t bars = base + lsd_off + LSD_FLOATS * 4;
  const uint32_t tile_full = bars, tile_empty = bars + 48;
  const uint32_t s_full = bars + 96, s_free = bars + 104, dp_full = bars + 112, dp_free = bars + 120;
  const uint32_t p_ready = bars + 128, p_freeb = bars + 136, ds_ready = bars + 144, ds_freeb = bars + 152;
  const uint32_t dkv_full = bars + 160, dkv_free = bars + 168, dq_full = bars + 176, dq_free = bars + 184;
  const uint32_t lsd_ready = bars + 192, lsd_taken = bars + 200;
  const uint32_t tmem_slot = bars + 208;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the padding rows NK .. TILE_ROWS-1 of every slot once (TMA never writes them): they enter dV / dK / dQ as
  // contraction rows with zero P / dS and must therefore be finite
  for (int i = threadIdx.x; i < NSLOT * (TILE_ROWS - G.NK) * 8; i += THREADS) {
    const int c = i & 7, rr = (i >> 3) % (TILE_ROWS - G.NK), slot = (i >> 3) / (TILE_ROWS - G.NK);
    *reinterpret_cast<uint4*>(gen + slot * TILE_BYTES + (G.NK + rr) * ROWB + c * 16) = make_uint4(0, 0, 0, 0);
  }
  // P / dS start as zeros: their never-written corners are read as garbage ROWS / lanes only, but keep them finite
  for (int i = threadIdx.x; i < 2 * PB_BYTES / 16; i += THREADS)
    *reinterpret_cast<uint4*>(gen + NSLOT * TILE_BYTES + i * 16) = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_rows); tma_prefetch_desc(&tm_cls);
    tma_prefetch_desc(&tm_do_rows); tma_prefetch_desc(&tm_do_cls);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(tile_full + 8 * i, 1); mbar_init(tile_empty + 8 * i, 1); }
    mbar_init(s_full, 1);   mbar_init(s_free, 8);
    mbar_init(dp_full, 1);  mbar_init(dp_free, 8);
    mbar_init(p_ready, 8);  mbar_init(p_freeb, 1);
    mbar_init(ds_ready, 8); mbar_init(ds_freeb, 1);
    mbar_init(dkv_full, 1); mbar_init(dkv_free, 4);
    mbar_init(dq_full, 1);  mbar_init(dq_free, 4);
    mbar_init(lsd_ready, 4); mbar_init(lsd_taken, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const int ksteps1 = G.W1 / 16;                        // contraction steps over the second (short) tile

  // register budget: the three single-thread roles give registers back, the softmax warps take them (the CTA keeps its launch allocation of 512 x 128: control 72, softmax 176, drain 88 -- setmaxnreg can only hand out what the CTA itself released)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      Tracer trc = make_tracer(trace, 0, true);
      int gi = 0;
      for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++gi) {
        const int f = g % G.T, h = (g / G.T) % G.H, b = g / (G.T * G.H);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {                    // Q, K, dO, V
          const int c = 4 * gi + j, slot = c % NSLOT, u = c / NSLOT;
          mbar_wait(tile_empty + 8 * slot, (u & 1) ^ 1);
          trace_ev(trc, 1, gi, j);                    // TMA: tile j issued
          const uint32_t fb = tile_full + 8 * slot, dst = base + slot * TILE_BYTES;
          mbar_expect_tx(fb, (uint32_t)G.NK * ROWB);
          if (j == 2) {
            tma_load_4d(dst, &tm_do_rows, fb, 0, 1 + f * G.N, h, b);
            tma_load_4d(dst + G.N * ROWB, &tm_do_cls, fb, 0, 0, h, b);
          } else {
            const int w = j == 0 ? 0 : (j == 1 ? 1 : 2);
            tma_load_4d(dst, &tm_rows, fb, 0, 1 + f * G.N, w * G.H + h, b);
            tma_load_4d(dst + G.N * ROWB, &tm_cls, fb, 0, 0, w * G.H + h, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc_s0 = make_idesc_bf16(128, 128, false, false);
    const uint32_t idesc_s1 = make_idesc_bf16(128, G.W1, false, false);
    constexpr uint32_t idesc_kv = make_idesc_bf16(128, HD, true, true);      // A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
    constexpr uint32_t idesc_q = make_idesc_bf16(128, HD, false, true);      // A = dS (K-major), B = K (MN-major)
    Tracer trc = make_tracer(trace, 1, lane == 0);
    int gi = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++gi) {
      const int c0 = 4 * gi;
      const uint32_t tq = base + ((c0 + 0) % NSLOT) * TILE_BYTES, tk = base + ((c0 + 1) % NSLOT) * TILE_BYTES;
      const uint32_t tdo = base + ((c0 + 2) % NSLOT) * TILE_BYTES, tv = base + ((c0 + 3) % NSLOT) * TILE_BYTES;

      auto issue_s = [&](int it) {                       // S[qt rows, kt keys] = Q K^T
        const int kt = it >> 1, qt = it & 1;
        if (lane == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(tmem + S_COL, make_smem_desc_sw128(tq + qt * 128 * ROWB + kk * 32, 16, 1024),
                         make_smem_desc_sw128(tk + kt * 128 * ROWB + kk * 32, 16, 1024), kt ? idesc_s1 : idesc_s0, kk > 0);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int it) {                      // dP = dO V^T
        const int kt = it >> 1, qt = it & 1;
        if (lane == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(tmem + DP_COL, make_smem_desc_sw128(tdo + qt * 128 * ROWB + kk * 32, 16, 1024),
                         make_smem_desc_sw128(tv + kt * 128 * ROWB + kk * 32, 16, 1024), kt ? idesc_s1 : idesc_s0, kk > 0);
          umma_commit(dp_full);
        }
        __syncwarp();
      };

      // first step of the group: S needs Q, K; dP needs dO, V
      {
        const int n = 4 * gi;
        mbar_wait(tile_full + 8 * ((c0 + 0) % NSLOT), ((c0 + 0) / NSLOT) & 1);
        mbar_wait(tile_full + 8 * ((c0 + 1) % NSLOT), ((c0 + 1) / NSLOT) & 1);
        mbar_wait(s_free, (n & 1) ^ 1);
        tc_fence_after();
        trace_ev(trc, 10, gi, 0);                 // MMA: S(0) issue
        issue_s(0);
        mbar_wait(tile_full + 8 * ((c0 + 2) % NSLOT), ((c0 + 2) / NSLOT) & 1);
        mbar_wait(tile_full + 8 * ((c0 + 3) % NSLOT), ((c0 + 3) / NSLOT) & 1);
        mbar_wait(dp_free, (n & 1) ^ 1);
        tc_fence_after();
        trace_ev(trc, 11, gi, 0);                 // MMA: dP(0) issue
        issue_dp(0);
      }
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        const int n = 4 * gi + it, kt = it >> 1, qt = it & 1;
        const int ks_q = qt ? ksteps1 : 8;               // contraction over this query tile's rows
        const int ks_k = kt ? ksteps1 : 8;               // contraction over this key tile's rows
        // ---- dV[kt] (+)= P^T dO
        mbar_wait(p_ready, n & 1);
        trace_ev(trc, 12, gi, it);                // MMA: p_ready seen
        if (qt == 0) mbar_wait(dkv_free, ((2 * gi + kt) & 1) ^ 1);       // previous key tile's dK / dV drained
        tc_fence_after();
        trace_ev(trc, 13, gi, it);                // MMA: dV issue
        if (lane == 0) {
          for (int j = 0; j < ks_q; ++j)
            umma_bf16_ss(tmem + DV_COL, make_smem_desc_sw128(sP + j * 2048, KBLK_BYTES, 1024),
                         make_smem_desc_sw128(tdo + (qt * 128 + j * 16) * ROWB, 8192, 1024), idesc_kv, (qt | j) != 0);
          umma_commit(p_freeb);
        }
        __syncwarp();
        trace_ev(trc, 16, gi, it);                // MMA: dV issued
        // ---- S of the next step as soon as the current S has left TMEM
        if (it < 3) {
          mbar_wait(s_free, ((n + 1) & 1) ^ 1);
          tc_fence_after();
          trace_ev(trc, 10, gi, it + 1);
          issue_s(it + 1);
        }
        // ---- dK[kt] (+)= dS^T Q,  dQ[qt] (+)= dS K
        mbar_wait(ds_ready, n & 1);
        trace_ev(trc, 14, gi, it);                // MMA: ds_ready seen
        if (it == 0) mbar_wait(dq_free, (gi & 1) ^ 1);                   // previous group's dQ drained
        tc_fence_after();
        trace_ev(trc, 15, gi, it);                // MMA: dK / dQ issue
        if (lane == 0) {
          for (int j = 0; j < ks_q; ++j)
            umma_bf16_ss(tmem + DK_COL, make_smem_desc_sw128(sDS + j * 2048, KBLK_BYTES, 1024),
                         make_smem_desc_sw128(tq + (qt * 128 + j * 16) * ROWB, 8192, 1024), idesc_kv, (qt | j) != 0);
          for (int j = 0; j < ks_k; ++j)
            umma_bf16_ss(tmem + DQ_COL + 64 * qt, make_smem_desc_sw128(sDS + (j >> 2) * KBLK_BYTES + (j & 3) * 32, 16, 1024),
                         make_smem_desc_sw128(tk + (kt * 128 + j * 16) * ROWB, 8192, 1024), idesc_q, (kt | j) != 0);
          umma_commit(ds_freeb);
          if (qt == 1) umma_commit(dkv_full);
          if (it == 3) {
            umma_commit(dq_full);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_commit(tile_empty + 8 * ((c0 + j) % NSLOT));
          }
        }
        __syncwarp();
        trace_ev(trc, 17, gi, it);                // MMA: dK / dQ issued
        if (it < 3) {
          mbar_wait(dp_free, ((n + 1) & 1) ^ 1);
          tc_fence_after();
          trace_ev(trc, 11, gi, it + 1);
          issue_dp(it + 1);
        }
      }
    }
  }
  } else if (warp < 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // ============================== softmax / dS ==============================
    const int qd = warp & 3, hf = (warp - 4) >> 2;
    const int r_in = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const int half1 = ((G.W1 / 8 + 1) / 2) * 8;          // columns of half 0 in the short key tile (multiple of 8)
    Tracer trc = make_tracer(trace, 2, warp == 4 && lane == 0);
    int gi = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++gi) {
      const int f = g % G.T;
      const float* lse2_s = lsd + (gi & 1) * 2 * TILE_ROWS;
      const float* delta_s = lse2_s + TILE_ROWS;
      mbar_wait(lsd_ready, gi & 1);                      // the drain warpgroup has prepared this group's lse2 / delta
      __syncwarp();
      if (lane == 0) mbar_arrive(lsd_taken);             // ... and may now prepare the next one (one phase ahead at most)
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        const int n = 4 * gi + it, kt = it >> 1, qt = it & 1;
        const int row = qt * 128 + r_in;                 // query row of the group
        const bool active = !(qt == 1 && qd * 32 >= G.W1);          // this warp's rows are beyond the padded tile
        const bool rvalid = row < G.NK;
        // keys this query may attend: all NK, except the CLS query outside frame 0 (no CLS key)
        const int nvis = rvalid ? ((row == G.N && f != 0) ? G.N : G.NK) : 0;
        const int ncol = kt ? (hf ? G.W1 - half1 : half1) : 64;     // this thread's columns of the key tile
        const int col0 = kt ? hf * half1 : hf * 64;
        // invalid (padding) rows: lse2 = +inf makes every P of the row exp2(-inf) = 0, hence dS = 0, without a mask
        SoftmaxArgs sa;
        sa.lse2 = rvalid ? lse2_s[row] : INFINITY;
        sa.delta = rvalid ? delta_s[row] : 0.f;
        sa.s_addr = tmem + lane_base + S_COL + col0;
        sa.dp_addr = tmem + lane_base + DP_COL + col0;
        sa.p_row = sP + r_in * ROWB;
        sa.ds_row = sDS + r_in * ROWB;
        sa.sw = r_in & 7;
        sa.col0 = col0;
        sa.ncol = ncol;
        sa.vis_cols = nvis - kt * 128 - col0;            // columns [0, vis_cols) of this thread are visible to its row
        sa.active = active;
        sa.lane = lane;
        sa.parity = n & 1;
        sa.s_full = s_full; sa.s_free = s_free; sa.p_freeb = p_freeb; sa.p_ready = p_ready;
        sa.dp_full = dp_full; sa.ds_freeb = ds_freeb;
        sa.trc = &trc; sa.gi = gi; sa.it = it;
        if (kt == 0) softmax_step<false>(sa);
        else softmax_step<true>(sa);
        fence_proxy_async_smem();
        warp_arrive(dp_free, lane);
        if (lane == 0) mbar_arrive(ds_ready);            // (ordered after the fences + __syncwarp of warp_arrive)
        trace_ev(trc, 24, gi, it);               // softmax: dS written
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // ============================== drain warpgroup ==============================
    // (1) per-row lse (log2 units) and delta = rowsum(dO o O) of the NEXT group, straight from global memory (so it
    //     does not wait for the TMA ring), double-buffered by group parity; (2) the accumulator drains: dK / dV after
    //     each key tile, dQ at the end of the group -- 64-byte row pieces to dqkv, CLS rows by fp32 atomics.  The softmax
    //     warps never wait for either.
    const int qd = warp & 3;
    const int r_in = qd * 32 + lane;                     // TMEM lane = accumulator row owned by this thread
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const int t128 = threadIdx.x - 384;                  // 0..127
    Tracer trc = make_tracer(trace, 3, warp == 12 && lane == 0);

    auto prepare_rows = [&](int g, int par) {
      const int f = g % G.T, h = (g / G.T) % G.H, b = g / (G.T * G.H);
      float* lse2_s = lsd + par * 2 * TILE_ROWS;
      float* delta_s = lse2_s + TILE_ROWS;
#pragma unroll 1
      for (int row = t128; row < G.NK; row += 128) {
        const long long tok = (long long)b * G.S + (row < G.N ? 1 + f * G.N + row : 0);
        const uint4* orow = reinterpret_cast<const uint4*>(out + tok * G.D + h * HD);
        const uint4* drow = reinterpret_cast<const uint4*>(dout + tok * G.D + h * HD);
        const float l = __ldg(lse_in + ((long long)(b * G.H + h)) * G.S + (row < G.N ? 1 + f * G.N + row : 0));
        float d = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                   // 64 B of O and dO at a time (72 registers per thread here)
          uint4 ov[4], dv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { ov[i] = __ldg(orow + 4 * hh + i); dv[i] = __ldg(drow + 4 * hh + i); }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t ou[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w}, du[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 a = unpack_bf16x2(ou[k]), bb = unpack_bf16x2(du[k]);
              d = fmaf(a.x, bb.x, d);
              d = fmaf(a.y, bb.y, d);
            }
          }
        }
        lse2_s[row] = l * LOG2E;
        delta_s[row] = d;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(lsd_ready);             // mbarrier arrive = release: the smem writes above are visible
    };
    auto drain_row = [&](uint32_t col, float scale, bf16* dst_row, float* dst_cls, bool is_row, bool is_cls) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t acc[32];
        tmem_ld32(tmem + lane_base + col + hf * 32, acc);
        tmem_ld_wait();
        if (is_row) store_row32(acc, scale, dst_row + hf * 32);
        else if (is_cls) atomic_row32(acc, dst_cls + hf * 32);
      }
    };

    if ((int)blockIdx.x < G.groups) prepare_rows(blockIdx.x, 0);
    int gi = 0;
    for (int g = blockIdx.x; g < G.groups; g += gridDim.x, ++gi) {
      const int f = g % G.T, h = (g / G.T) % G.H, b = g / (G.T * G.H);
      // the next group's lse2 / delta: its buffer (parity gi + 1) was last read in group gi - 1, whose dQ this warpgroup
      // has already drained, i.e. every softmax warp is past it
      if (g + (int)gridDim.x < G.groups) {
        mbar_wait(lsd_taken, gi & 1);                    // every softmax warp has seen phase gi of lsd_ready
        trace_ev(trc, 30, gi, 0);
        prepare_rows(g + gridDim.x, (gi + 1) & 1);
        trace_ev(trc, 31, gi, 0);  // drain WG: next group's lse / delta ready
      }
      bf16* base_row = dqkv + ((long long)b * G.S + 1 + f * G.N) * (3 * G.D) + h * HD;
      float* cls = dcls + ((long long)(b * G.H + h) * 3) * HD;
#pragma unroll 1
      for (int kt = 0; kt < 2; ++kt) {
        mbar_wait(dkv_full, (2 * gi + kt) & 1);
        tc_fence_after();
        trace_ev(trc, 32, gi, kt);  // drain WG: dkv_full seen
        const int key = kt * 128 + r_in;
        if (!(kt == 1 && qd * 32 >= G.W1)) {
          drain_row(DK_COL, 1.f, base_row + (long long)key * (3 * G.D) + G.D, cls + HD, key < G.N, key == G.N);
          drain_row(DV_COL, 1.f, base_row + (long long)key * (3 * G.D) + 2 * G.D, cls + 2 * HD, key < G.N, key == G.N);
        }
        warp_arrive(dkv_free, lane);
        trace_ev(trc, 33, gi, kt);  // drain WG: dK / dV stored
      }
      mbar_wait(dq_full, gi & 1);
      tc_fence_after();
      trace_ev(trc, 34, gi, 0);    // drain WG: dq_full seen
#pragma unroll 1
      for (int qt = 0; qt < 2; ++qt) {
        const int row = qt * 128 + r_in;                 // the CLS query (row N): raw sum, scaled by cls_grad_finalize_kernel
        if (qt == 0 || qd * 32 < G.W1)
          drain_row(DQ_COL + 64 * qt, q_scale, base_row + (long long)row * (3 * G.D), cls, row < G.N, row == G.N);
      }
      warp_arrive(dq_free, lane);
      trace_ev(trc, 35, gi, 0);    // drain WG: dQ stored
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

// Geometry the kernel covers: two key / query tiles, 128 < N + 1 <= 208.  EGOVLP_ATTN_TC_BWD=0 keeps the mma.sync kernel.
bool space_attn_bwd_tc_supported(int N) {
  const char* e = getenv("EGOVLP_ATTN_TC_BWD");
  if (e && e[0] == '0') return false;
  const int NK = N + 1;
  return NK > 128 && NK <= TILE_ROWS;
}

int space_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dcls,
                      int B, int T, int N, int H, float q_scale, cudaStream_t st) {
  BwdGeom G;
  G.B = B; G.H = H; G.T = T; G.N = N; G.S = 1 + T * N; G.D = H * HD; G.NK = N + 1; G.NKP = (G.NK + 15) / 16 * 16;
  G.W1 = G.NKP - 128; G.groups = B * H * T;
  CUtensorMap tm_rows, tm_cls, tm_do_rows, tm_do_cls;
  {
    const uint64_t W = 3ull * G.D;
    const uint64_t dims[4] = {HD, (uint64_t)G.S, (uint64_t)(3 * H), (uint64_t)B};
    const uint64_t strides[4] = {1, W, HD, (uint64_t)G.S * W};
    const uint32_t box_rows[4] = {HD, (uint32_t)N, 1, 1}, box_cls[4] = {HD, 1, 1, 1};
    int rc = make_tmap_nd_bf16(&tm_rows, qkv, 4, dims, strides, box_rows, true);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tm_cls, qkv, 4, dims, strides, box_cls, true);
    if (rc) return rc;
  }
  {
    const uint64_t W = (uint64_t)G.D;
    const uint64_t dims[4] = {HD, (uint64_t)G.S, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[4] = {1, W, HD, (uint64_t)G.S * W};
    const uint32_t box_rows[4] = {HD, (uint32_t)N, 1, 1}, box_cls[4] = {HD, 1, 1, 1};
    int rc = make_tmap_nd_bf16(&tm_do_rows, dout, 4, dims, strides, box_rows, true);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tm_do_cls, dout, 4, dims, strides, box_cls, true);
    if (rc) return rc;
  }
  const int smem = NSLOT * TILE_BYTES + 2 * PB_BYTES + LSD_FLOATS * 4 + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(space_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int grid = G.groups < num_sms() ? G.groups : num_sms();
  const char* trace_path = getenv("EGOVLP_ATTN_BWD_TRACE");
  unsigned long long* trace = nullptr;
  if (trace_path && trace_path[0]) {
    EGOVLP_CHECK_CUDA(cudaMalloc(&trace, 4001 * sizeof(unsigned long long)));
    EGOVLP_CHECK_CUDA(cudaMemsetAsync(trace, 0, 4001 * sizeof(unsigned long long), st));
  }
  space_attn_bwd_tc_kernel<<<grid, THREADS, smem, st>>>(tm_rows, tm_cls, tm_do_rows, tm_do_cls,
                                                        reinterpret_cast<const bf16*>(out),
                                                        reinterpret_cast<const bf16*>(dout), lse,
                                                        reinterpret_cast<bf16*>(dqkv), dcls, q_scale, G, trace);
  EGOVLP_CHECK_LAUNCH();
  if (trace) {                                             // debugging aid: synchronous dump of the timeline
    static unsigned long long host[4001];
    EGOVLP_CHECK_CUDA(cudaStreamSynchronize(st));
    EGOVLP_CHECK_CUDA(cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int i = 0; i < 4000; ++i)
        if (host[1 + i])
          fprintf(f, "%llu %llu %llu %llu\n", host[1 + i] >> 56, (host[1 + i] >> 48) & 0xff, (host[1 + i] >> 44) & 0xf,
                  host[1 + i] & 0xfffffffffffull);
      fclose(f);
    }
  }
  return EGOVLP_OK;
}

}  // namespace egovlp
