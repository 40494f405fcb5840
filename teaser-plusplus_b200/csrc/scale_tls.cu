// Unknown-scale variant of stage 1 (Params::estimate_scaling = true, the reference's default):
//   TLSScaleSolver::solveForScale         teaser/src/registration.cc:410-425
//   ScalarTLSEstimator::estimate over K   teaser/src/registration.cc:21-88
// Per TIM k=(i,j), i<j in the reference's order k = i*N - i(i+1)/2 + (j-i-1) (registration.cc:531):
//   ratio_k = ||dst_j-dst_i|| / ||src_j-src_i||,  alpha_k = beta * (1/||src_j-src_i||)
// then a scalar TLS over the K ratios gives the scale; the inlier predicate |ratio - s| <= alpha is
// re-evaluated (identically) by the graph tile kernel when it emits the bitset.
//
// Round-1 implementation: one CTA per problem sorts the 2K interval end points with a bitonic network and
// one thread sweeps them in the reference's accumulation order (bit-exact vs the oracle up to the order of
// tied end points).  Correct for every size but latency-bound; api.cu limits it to n <= kMaxScaleN.  The
// HBM-bound formulation for large K (device radix sort + segmented scans) is the next step.
#include "tls_device.cuh"
#include "tzr_internal.cuh"

namespace tzr {

__global__ void __launch_bounds__(256) scale_pairs_kernel(Batch bt, double* X, double* Rg) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int n = bt.n;
  const long long K = (long long)n * (n - 1) / 2;
  const double* src = bt.src + (size_t)b * n * 3;
  const double* dst = bt.dst + (size_t)b * n * 3;
  double* x = X + (size_t)b * K;
  double* r = Rg + (size_t)b * K;
  const long long seg = (long long)i * n - (long long)i * (i + 1) / 2;  // registration.cc:531
  for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
    const double d1 = tim_norm_exact(src, i, j);
    const double d2 = tim_norm_exact(dst, i, j);
    const long long k = seg + (j - i - 1);
    x[k] = __ddiv_rn(d2, d1);                        // raw_scales = v2_dist / v1_dist        :420
    r[k] = __dmul_rn(bt.beta, __ddiv_rn(1.0, d1));   // alphas = beta * v1_dist.cwiseInverse() :422
  }
}

__global__ void __launch_bounds__(1024) scale_tls_kernel(Batch bt, const double* X, const double* Rg, double* key,
                                                          int32_t* idx, long long npad) {
  const int b = blockIdx.x;
  const int n = bt.n;
  const long long K = (long long)n * (n - 1) / 2;
  const double* x = X + (size_t)b * K;
  const double* r = Rg + (size_t)b * K;
  double* kk = key + (size_t)b * npad;
  int32_t* ii = idx + (size_t)b * npad;
  __shared__ double s_est;
  for (long long q = threadIdx.x; q < npad; q += blockDim.x) {
    double k;
    int id;
    if (q < 2 * K) {
      const long long t = q >> 1;
      k = (q & 1) ? x[t] + r[t] : x[t] - r[t];  // registration.cc:36-37
      id = (int)q;
    } else {
      k = 1.0 / 0.0;
      id = 0x7fffffff;
    }
    kk[q] = k;
    ii[q] = id;
  }
  __syncthreads();
  bitonic_sort_block(kk, ii, (int)npad, 1);
  if (threadIdx.x == 0) s_est = tls_sweep(x, r, 0.0, K, ii);
  __syncthreads();
  if (threadIdx.x == 0) bt.sol[b].scale = s_est;
}

int launch_scale_estimation(const Batch& bt, double* X, double* Rg, double* key, int32_t* idx, long long npad,
                            cudaStream_t st) {
  if (bt.n < 2) return 0;
  dim3 g1((unsigned)(bt.n - 1), (unsigned)bt.B);
  scale_pairs_kernel<<<g1, 256, 0, st>>>(bt, X, Rg);
  scale_tls_kernel<<<bt.B, 1024, 0, st>>>(bt, X, Rg, key, idx, npad);
  return 2;
}

}  // namespace tzr
