// Device-side scalar TLS (adaptive voting) shared by the translation, scale and stand-alone kernels.
// ScalarTLSEstimator::estimate  (teaser/src/registration.cc:21-88)
#pragma once
#include "tzr_internal.cuh"

namespace tzr {
namespace {

// ---- scalar TLS (registration.cc:21-88): bitonic sort of the 2M interval endpoints + sweep --------
// Arrays for `nax` independent estimators are processed together.  key/idx: nax * npad entries.
// payload idx: 2*i (lower endpoint, enters the consensus set) or 2*i+1 (upper endpoint, leaves it).
__device__ inline bool key_less(double ka, int ia, double kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

__device__ void bitonic_sort_block(double* key, int* idx, int npad, int nax) {
  const int half = npad >> 1;
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < half * nax; t += blockDim.x) {
        const int ax = t / half, q = t - ax * half;
        const int lo = 2 * q - (q & (j - 1));  // index with bit j cleared
        const int hi = lo + j;
        double* kk = key + (size_t)ax * npad;
        int* ii = idx + (size_t)ax * npad;
        const bool up = ((lo & k) == 0);
        const double ka = kk[lo], kb = kk[hi];
        const int ia = ii[lo], ib = ii[hi];
        const bool sw = up ? key_less(kb, ib, ka, ia) : key_less(ka, ia, kb, ib);
        if (sw) {
          kk[lo] = kb;
          kk[hi] = ka;
          ii[lo] = ib;
          ii[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
}

// Sequential sweep over the sorted endpoints (one thread), exactly the reference's accumulation order.
// X: measurements, ranges: per-measurement range (or nullptr -> const_range).
__device__ double tls_sweep(const double* __restrict__ X, const double* __restrict__ ranges, double const_range,
                            long long M, const int* __restrict__ idx) {
  double ranges_inverse_sum = 0.0;
  for (long long i = 0; i < M; ++i) ranges_inverse_sum += ranges ? ranges[i] : const_range;  // ranges.sum() :51
  double dot_X_weights = 0.0, dot_weights_consensus = 0.0;
  long long consensus = 0;
  double sum_xi = 0.0, sum_xi_square = 0.0;
  double best_cost = 0.0, best_xhat = 0.0;
  const long long nr = 2 * M;
  for (long long i = 0; i < nr; ++i) {  // :58-75
    const int e = idx[i];
    const long long id = e >> 1;
    const double eps = (e & 1) ? -1.0 : 1.0;
    const double r = ranges ? ranges[id] : const_range;
    const double w = 1.0 / (r * r);
    const double x = X[id];
    consensus += (e & 1) ? -1 : 1;
    dot_weights_consensus += eps * w;
    dot_X_weights += eps * w * x;
    ranges_inverse_sum -= eps * r;
    sum_xi += eps * x;
    sum_xi_square += eps * x * x;
    const double x_hat = dot_X_weights / dot_weights_consensus;
    const double residual = (double)consensus * x_hat * x_hat + sum_xi_square - 2.0 * sum_xi * x_hat;
    const double x_cost = residual + ranges_inverse_sum;
    if (i == 0 || x_cost < best_cost) {  // minCoeff: first strict minimum, NaN never wins after i=0
      best_cost = x_cost;
      best_xhat = x_hat;
    }
  }
  return best_xhat;
}

__device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}


}  // namespace
}  // namespace tzr
