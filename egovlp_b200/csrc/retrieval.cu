// Retrieval-evaluation kernels: per-query ranking metrics of the EPIC-Kitchens MIR evaluation,
// reference utils/nDCG.py:3-45,96-139 (calculate_DCG / calculate_k_counts / calculate_nDCG) and utils/mAP.py:4-44
// (calculate_mAP), which the reference runs as numpy argsort + fancy indexing on the host
// (model/metric.py:257-299, run/test_epic.py:137-157).
//
// One CTA ranks one query row entirely in shared memory: the row of similarities is loaded once (coalesced), sorted
// descending together with its column indices by a bitonic network, and the relevancy row is gathered through the
// sorted indices; DCG and average precision are then block reductions / one block scan in fp64.  HBM traffic is the
// algorithmic minimum (one read of the similarity row and of the relevancy row), there is no [rows, cols] index
// matrix in global memory, and nothing is copied to the host.
#include "common.cuh"

namespace egovlp {
namespace {

constexpr int RANK_THREADS = 512;
constexpr int RANK_MAX_COLS = 16384;

// strict total order of the ranking: larger similarity first; ties by column index (tie_hi: larger index first, i.e.
// a stable ascending argsort reversed as in nDCG.py:32; otherwise smaller index first, a stable argsort of -sim as in
// mAP.py:25)
__device__ __forceinline__ bool ranks_before(float ka, unsigned short ia, float kb, unsigned short ib, bool tie_hi) {
  return ka > kb || (ka == kb && (tie_hi ? ia > ib : ia < ib));
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < RANK_THREADS / 32; ++w) s += red[w];
  return s;
}

template <typename RelT>
__global__ void __launch_bounds__(RANK_THREADS)
rank_metrics_kernel(const float* __restrict__ sim, long long ld_sim, const RelT* __restrict__ rel, long long ld_rel,
                    const int* __restrict__ k_counts, int cols, int np2, int tie_hi, double* __restrict__ dcg,
                    double* __restrict__ ap) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* key = reinterpret_cast<float*>(smem);
  unsigned short* idx = reinterpret_cast<unsigned short*>(key + np2);
  double* red = reinterpret_cast<double*>(smem + (size_t)np2 * 6);      // np2 * 6 is a multiple of 8 (np2 >= 512)
  double* scan = red + RANK_THREADS / 32;
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* srow = sim + (long long)row * ld_sim;
  const RelT* rrow = rel + (long long)row * ld_rel;

  int n_pos = 0;                    // relevant items of this query (k of nDCG.py:47-75 when no k_counts is given)
  for (int c = tid; c < np2; c += RANK_THREADS) {
    key[c] = c < cols ? srow[c] : -INFINITY;
    idx[c] = (unsigned short)(c < cols ? c : 0xFFFF);
    if (c < cols && (double)rrow[c] > 0.0) ++n_pos;
  }
  const int k_row = (int)(block_sum((double)n_pos, red) + 0.5);
  __syncthreads();

  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (np2 >> 1); t += RANK_THREADS) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo + j;
        const float ka = key[lo], kb = key[hi];
        const unsigned short ia = idx[lo], ib = idx[hi];
        const bool up = (lo & k) == 0;
        const bool swap = up ? ranks_before(kb, ib, ka, ia, tie_hi) : ranks_before(ka, ia, kb, ib, tie_hi);
        if (swap) { key[lo] = kb; key[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
      }
      __syncthreads();
    }
  }

  // rank i (0-based) holds column idx[i].  Each thread owns `per` consecutive ranks.
  const int per = np2 / RANK_THREADS, i0 = tid * per;
  double d_part = 0.0, c_part = 0.0;
  int n_one = 0;
  for (int i = i0; i < i0 + per && i < cols; ++i) {
    const double r = (double)rrow[idx[i]];
    const double kc = k_counts ? (double)k_counts[(long long)row * cols + i] : (i < k_row ? 1.0 : 0.0);
    d_part += r * kc / log2((double)i + 2.0);
    c_part += r;
  }
  const double d_tot = block_sum(d_part, red);
  // exclusive scan of the per-thread relevancy sums -> running cumulative relevancy (mAP.py:31)
  __syncthreads();
  scan[tid] = c_part;
  __syncthreads();
  for (int o = 1; o < RANK_THREADS; o <<= 1) {
    const double v = tid >= o ? scan[tid - o] : 0.0;
    __syncthreads();
    scan[tid] += v;
    __syncthreads();
  }
  double cum = scan[tid] - c_part, a_part = 0.0;
  for (int i = i0; i < i0 + per && i < cols; ++i) {
    const double r = (double)rrow[idx[i]];
    cum += r;
    if (r == 1.0) { a_part += cum / (double)(i + 1); ++n_one; }
  }
  const double a_tot = block_sum(a_part, red);
  const double ones = block_sum((double)n_one, red);
  if (tid == 0) {
    if (dcg) dcg[row] = d_tot;
    if (ap) ap[row] = a_tot / ones;            // 0 / 0 -> NaN, as numpy (a query without relevant items)
  }
}

}  // namespace
}  // namespace egovlp

using namespace egovlp;

extern "C" int egovlp_rank_metrics(const float* sim, long long ld_sim, const void* rel, int rel_is_f64, long long ld_rel,
                                   const int* k_counts, int rows, int cols, int tie_mode, double* dcg, double* ap,
                                   void* stream) {
  EGOVLP_CHECK_ARG(sim && rel && rows >= 0 && cols > 0 && (dcg || ap), "rank_metrics: bad args");
  EGOVLP_CHECK_ARG(cols <= RANK_MAX_COLS, "rank_metrics: more than 16384 gallery items per query are not supported");
  EGOVLP_CHECK_ARG(ld_sim >= cols && ld_rel >= cols, "rank_metrics: leading dimensions smaller than cols");
  if (rows == 0) return EGOVLP_OK;
  int np2 = RANK_THREADS;
  while (np2 < cols) np2 <<= 1;
  const size_t smem = (size_t)np2 * 6 + (RANK_THREADS / 32 + RANK_THREADS) * sizeof(double);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rel_is_f64) {
    EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(rank_metrics_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rank_metrics_kernel<double><<<rows, RANK_THREADS, smem, st>>>(sim, ld_sim, static_cast<const double*>(rel), ld_rel,
                                                                   k_counts, cols, np2, tie_mode != 0, dcg, ap);
  } else {
    EGOVLP_CHECK_CUDA(cudaFuncSetAttribute(rank_metrics_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rank_metrics_kernel<float><<<rows, RANK_THREADS, smem, st>>>(sim, ld_sim, static_cast<const float*>(rel), ld_rel,
                                                                  k_counts, cols, np2, tie_mode != 0, dcg, ap);
  }
  EGOVLP_CHECK_LAUNCH();
  return EGOVLP_OK;
}
