// Certification downstream of solve(): teaser::DRSCertifier::certify (reference teaser/src/certification.cc:40-190)
// on the device, FP64 throughout.  n = 4N + 4; all matrices are dense n x n, column-major.
//
//   cert_blocks_kernel   per measurement k: the 4x4 blocks of Q_cost (getQCost, :233-291) already rotated by the
//                        block-diagonal Omega (:293-314: Q_bar = D^T Q D acts block-wise), the 4x4 block of the initial
//                        multiplier guess (getLambdaGuess, :448-529) and the terms of mu = x^T Q x (:92).
//   cert_reduce_kernel   sums those per-k terms (top-left lambda block, mu).
//   cert_fill_kernel     assembles M_init = Q_bar - mu J_bar - lambda_bar_init (:100).
//   sym_kernel           B = (M + M^T) / 2 (getNearestPSD linalg.h:90; computeSubOptimalityGap :193).
//   [eigendecomposition of B: cusolverDnDsyevd, loaded lazily with dlopen — the one library call of this stage, in
//    the role cuBLAS has for plain GEMMs; see DESIGN.md 3.10]
//   psd_gemm_kernel      M_PSD = V diag(max(w, 0)) V^T over the positive eigenpairs only (linalg.h:93-98).
//   affine_in_kernel     W = 2 M_PSD - M - M_init (:139).
//   dual_bw_kernel .. dual_diag_kernel   getOptimalDualProjection (:316-446) with the sparse inverse map A_inv
//                        (getLinearProjection, :531-655: (N(N+1)/2)^2 entries in the reference) applied in closed form:
//                        out[a,b] = (x+2y) B[a,b] + y (th_a (Rs[b]-Cs[b]) - th_b (Rs[a]-Cs[a])) with two O(N^2) sums.
//   update_kernel        M_affine = M_init + W_dual (:153); M += gamma_tau (M_affine - M_PSD) (:181).
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "tzr_internal.cuh"

namespace tzr {

namespace {

// ---- lazily bound cuSOLVER entry points (keeps libteaser_b200.so loadable without libcusolver) ----------------
struct Cusolver {
  void* lib = nullptr;
  cusolverStatus_t (*create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*set_stream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*syevd_buf)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double*, int,
                                const double*, int*) = nullptr;
  cusolverStatus_t (*syevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double*, int, double*,
                            double*, int, int*) = nullptr;
  bool ok = false;
  std::string err;
};

Cusolver& cusolver() {
  static Cusolver c = [] {
    Cusolver s;
    const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11"};
    for (const char* nme : names) {
      s.lib = dlopen(nme, RTLD_NOW | RTLD_LOCAL);
      if (s.lib) break;
    }
    if (!s.lib) {
      s.err = std::string("cannot load libcusolver: ") + dlerror();
      return s;
    }
    s.create = (decltype(s.create))dlsym(s.lib, "cusolverDnCreate");
    s.destroy = (decltype(s.destroy))dlsym(s.lib, "cusolverDnDestroy");
    s.set_stream = (decltype(s.set_stream))dlsym(s.lib, "cusolverDnSetStream");
    s.syevd_buf = (decltype(s.syevd_buf))dlsym(s.lib, "cusolverDnDsyevd_bufferSize");
    s.syevd = (decltype(s.syevd))dlsym(s.lib, "cusolverDnDsyevd");
    s.ok = s.create && s.destroy && s.set_stream && s.syevd_buf && s.syevd;
    if (!s.ok) s.err = "libcusolver lacks the Dsyevd entry points";
    return s;
  }();
  return c;
}

// ---- small dense helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mat4_mul(const double* A, const double* B, double* C) {  // column-major 4x4
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += A[r + 4 * k] * B[k + 4 * c];
      C[r + 4 * c] = s;
    }
}

// Omega^T X Omega for a column-major 4x4 X
__device__ void rotate_block(const double* om, const double* X, double* out) {
  double omT[16], t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) omT[r + 4 * c] = om[c + 4 * r];
  mat4_mul(omT, X, t);
  mat4_mul(t, om, out);
}

__constant__ signed char kP[9][16] = {  // certification.cc:242-252
    {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1},  {0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0},
    {0, 0, 1, 0, 0, 0, 0, -1, 1, 0, 0, 0, 0, -1, 0, 0},  {0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, -1, 0, 0, -1, 0},
    {-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1},  {0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0},
    {0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 0, 0},    {0, 0, 0, -1, 0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0},
    {-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};

__device__ __forceinline__ void hat(const double* u, double* H) {  // column-major 3x3 (linalg.h:20-29)
  H[0] = 0;      H[3] = -u[2];  H[6] = u[1];
  H[1] = u[2];   H[4] = 0;      H[7] = -u[0];
  H[2] = -u[1];  H[5] = u[0];   H[8] = 0;
}

// per-k outputs, 16 doubles each (column-major 4x4): offA (rotated first-row/column block), diagB (rotated diagonal
// block), lam (the un-negated "current_block" of getLambdaGuess); terms[k] = theta-weighted contribution to mu
__global__ void cert_blocks_kernel(const double* __restrict__ src, const double* __restrict__ dst,
                                   const double* __restrict__ theta, int N, const double* __restrict__ Rcm,
                                   const double* __restrict__ q_xyzw, double nbs, double* offA, double* diagB,
                                   double* lam, double* mu_terms) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const double v1[3] = {src[3 * k], src[3 * k + 1], src[3 * k + 2]};
  const double v2[3] = {dst[3 * k], dst[3 * k + 1], dst[3 * k + 2]};
  // P_k = reshape(P' vec(v2 v1'), [4,4]) (column-major vec and reshape, :268-271)
  double A9[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) A9[r + 3 * c] = v2[r] * v1[c];
  double Pk[16];
  for (int e = 0; e < 16; ++e) {
    double s = 0;
    for (int m = 0; m < 9; ++m) s += (double)kP[m][e] * A9[m];
    Pk[e] = s;
  }
  const double nn = (v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]) + (v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  const double ck = 0.5 * (nn - nbs), ck2 = 0.5 * (nn + nbs);
  double Ak[16], Bk[16];
  for (int e = 0; e < 16; ++e) {
    const double id = (e % 5 == 0) ? 1.0 : 0.0;
    Ak[e] = -0.5 * Pk[e] + ck / 2 * id;
    Bk[e] = -Pk[e] + ck2 * id;
  }
  const double x = q_xyzw[0], y = q_xyzw[1], z = q_xyzw[2], w = q_xyzw[3];
  const double om[16] = {w, z, -y, -x, -z, w, x, -y, y, -x, w, -z, x, y, z, w};  // getOmega1 (:293-303), column-major
  rotate_block(om, Ak, offA + 16 * (size_t)k);
  rotate_block(om, Bk, diagB + 16 * (size_t)k);
  // mu = x^T Q x, x = kron([1, theta], q): blocks (0,k), (k,0) hold A_k, (k,k) holds B_k
  const double qv[4] = {x, y, z, w};
  double qAq = 0, qBq = 0;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      qAq += qv[r] * Ak[r + 4 * c] * qv[c];
      qBq += qv[r] * Bk[r + 4 * c] * qv[c];
    }
  const double th = theta[k];
  mu_terms[k] = 2.0 * th * qAq + th * th * qBq;
  // getLambdaGuess (:448-529)
  double R[9];
  for (int e = 0; e < 9; ++e) R[e] = Rcm[e];  // column-major
  double Rs[3], d[3], xi[3];
  for (int r = 0; r < 3; ++r) Rs[r] = R[r] * v1[0] + R[r + 3] * v1[1] + R[r + 6] * v1[2];
  for (int r = 0; r < 3; ++r) d[r] = v2[r] - Rs[r];
  for (int r = 0; r < 3; ++r) xi[r] = R[3 * r] * d[0] + R[3 * r + 1] * d[1] + R[3 * r + 2] * d[2];  // R^T d
  double sh[9], xh[9], shsh[9], xhsh[9];
  hat(v1, sh);
  hat(xi, xh);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      double a = 0, b = 0;
      for (int m = 0; m < 3; ++m) {
        a += sh[r + 3 * m] * sh[m + 3 * c];
        b += xh[r + 3 * m] * sh[m + 3 * c];
      }
      shsh[r + 3 * c] = a;
      xhsh[r + 3 * c] = b;
    }
  const double sxi = v1[0] * xi[0] + v1[1] * xi[1] + v1[2] * xi[2];
  const double xx = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2];
  const bool inl = th > 0;
  const double c44 = inl ? (-0.75 * xx - 0.25 * nbs) : (-0.25 * xx - 0.75 * nbs);
  const double cxx = inl ? 0.75 : 0.25, cvec = inl ? -1.5 : -0.5;
  double blk[16];
  for (int e = 0; e < 16; ++e) blk[e] = 0;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      const double id = r == c ? 1.0 : 0.0;
      blk[r + 4 * c] = shsh[r + 3 * c] - 0.5 * sxi * id + 0.5 * xhsh[r + 3 * c] + 0.5 * xi[r] * v1[c] - cxx * xx * id -
                       0.25 * nbs * id;
    }
  for (int r = 0; r < 3; ++r) {
    const double v = cvec * (xh[r] * v1[0] + xh[r + 3] * v1[1] + xh[r + 6] * v1[2]);
    blk[r + 12] = v;  // column 3
    blk[3 + 4 * r] = v;  // row 3
  }
  blk[15] = c44;
  for (int e = 0; e < 16; ++e) lam[16 * (size_t)k + e] = blk[e];
}

// one CTA: top[16] = sum_k lam_k, *mu = sum_k mu_terms[k] (k ascending per thread, then a fixed tree: deterministic)
__global__ void __launch_bounds__(256) cert_reduce_kernel(const double* lam, const double* mu_terms, int N,
                                                         double* top, double* mu) {
  __shared__ double sh[256];
  for (int e = 0; e <= 16; ++e) {
    double s = 0;
    for (int k = threadIdx.x; k < N; k += 256) s += (e < 16) ? lam[16 * (size_t)k + e] : mu_terms[k];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (e < 16) top[e] = sh[0];
      else *mu = sh[0];
    }
    __syncthreads();
  }
}

// M_init = Q_bar - mu J_bar - lambda_bar_init, one thread per entry
__global__ void cert_fill_kernel(int N, const double* offA, const double* diagB, const double* lam, const double* top,
                                 const double* mu, double* M_init) {
  const int n = 4 * N + 4;
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int r = (int)(e % n), c = (int)(e / n);
  const int bi = r >> 2, bj = c >> 2, lr = r & 3, lc = c & 3;
  double v = 0;
  if (bi == 0 && bj == 0) {
    v = -((lr == lc) ? *mu : 0.0) - top[lr + 4 * lc];  // Q_bar(0,0) = 0; -mu J; lambda(0,0) = +sum of blocks
  } else if (bi == 0) {
    v = offA[16 * (size_t)(bj - 1) + lr + 4 * lc];
  } else if (bj == 0) {
    v = offA[16 * (size_t)(bi - 1) + lr + 4 * lc];
  } else if (bi == bj) {
    v = diagB[16 * (size_t)(bi - 1) + lr + 4 * lc] + lam[16 * (size_t)(bi - 1) + lr + 4 * lc];  // lambda = -blk
  }
  M_init[e] = v;
}

__global__ void sym_kernel(const double* __restrict__ M, int n, double* __restrict__ B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int r = (int)(e % n), c = (int)(e / n);
  B[e] = (M[e] + M[(size_t)r * n + c]) / 2;
}

// first index with w > 0 (w ascending) and Vs = V * diag(max(w, 0)) for those columns
__global__ void scale_pos_kernel(const double* __restrict__ V, const double* __restrict__ w, int n,
                                 double* __restrict__ Vs) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int c = (int)(e / n);
  const double wc = w[c];
  Vs[e] = wc > 0 ? V[e] * wc : 0.0;
}

__global__ void first_pos_kernel(const double* w, int n, int* first_pos) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int lo = 0, hi = n;  // w ascending
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (w[mid] > 0) hi = mid;
      else lo = mid + 1;
    }
    *first_pos = lo;
  }
}

// C (n x n) = Vs[:, k0:] * V[:, k0:]^T, 64 x 64 tile per CTA, 16 x 16 threads x (4 x 4), k in slabs of 16
constexpr int kGT = 64, kGK = 16;
__global__ void __launch_bounds__(256) psd_gemm_kernel(const double* __restrict__ Vs, const double* __restrict__ V,
                                                      int n, const int* __restrict__ first_pos,
                                                      double* __restrict__ C) {
  __shared__ double As[kGK][kGT + 1], Bs[kGK][kGT + 1];
  const int k0 = *first_pos;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.x * kGT, c0 = blockIdx.y * kGT;
  double acc[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  for (int kb = k0; kb < n; kb += kGK) {
    for (int e = threadIdx.x; e < kGK * kGT; e += 256) {
      const int kk = e / kGT, i = e % kGT;
      const int k = kb + kk;
      As[kk][i] = (k < n && r0 + i < n) ? Vs[(size_t)k * n + r0 + i] : 0.0;
      Bs[kk][i] = (k < n && c0 + i < n) ? V[(size_t)k * n + c0 + i] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        b[i] = Bs[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (r < n && c < n) C[(size_t)c * n + r] = acc[i][j];
    }
}

__global__ void affine_in_kernel(const double* __restrict__ Mpsd, const double* __restrict__ M,
                                 const double* __restrict__ Minit, long long nn, double* __restrict__ W) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nn) W[e] = 2 * Mpsd[e] - M[e] - Minit[e];
}

// ---- getOptimalDualProjection -----------------------------------------------------------------------------------
// pair (i, j), 0 <= i < j <= N, running index as in the reference's double loop (:342-343)
__device__ __forceinline__ long long pair_id(int i, int j, int N1) {  // N1 = N + 1 blocks
  return (long long)i * N1 - (long long)i * (i + 1) / 2 + (j - i - 1);
}

#define WAT(r, c) W[(size_t)(c) * n + (r)]

// bW[p][0..2] = -th_ij C + D - E + th_ij F (:349-368)
__global__ void dual_bw_kernel(const double* __restrict__ W, const double* __restrict__ thp, int N,
                               double* __restrict__ bW) {
  const int n = 4 * N + 4, N1 = N + 1;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i >= N || j <= i || j >= N1) return;
  const double tij = thp[i] * thp[j];
  const long long p = pair_id(i, j, N1);
  for (int c = 0; c < 3; ++c) {
    const double Cc = WAT(4 * i + 3, 4 * i + c), Dc = WAT(4 * j + 3, 4 * i + c);
    const double Ec = WAT(4 * i + 3, 4 * j + c), Fc = WAT(4 * j + 3, 4 * j + c);
    bW[3 * p + c] = ((-tij * Cc + Dc) + (-1.0) * Ec) + tij * Fc;
  }
}

// D[v][c] = Rs[v][c] - Cs[v][c]: Rs[v] = sum_{j>v} th_j B[v,j], Cs[v] = sum_{i<v} th_i B[i,v]   (one thread per (v,c))
__global__ void dual_sums_kernel(const double* __restrict__ bW, const double* __restrict__ thp, int N,
                                 double* __restrict__ D) {
  const int N1 = N + 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * N1) return;
  const int v = t / 3, c = t % 3;
  double rs = 0, cs = 0;
  for (int j = v + 1; j < N1; ++j) rs += thp[j] * bW[3 * pair_id(v, j, N1) + c];
  for (int i = 0; i < v; ++i) cs += thp[i] * bW[3 * pair_id(i, v, N1) + c];
  D[t] = rs - cs;
}

// off-diagonal blocks of W_dual (both triangles) (:381-415), A_inv applied in closed form
__global__ void dual_off_kernel(const double* __restrict__ W, const double* __restrict__ thp,
                                const double* __restrict__ bW, const double* __restrict__ D, int N,
                                double* __restrict__ Wd) {
  const int n = 4 * N + 4, N1 = N + 1;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i >= N || j <= i || j >= N1) return;
  const double y = 1.0 / (2.0 * N + 6.0), x = (N + 1.0) * y;
  const long long p = pair_id(i, j, N1);
  double yd[3];
  for (int c = 0; c < 3; ++c)
    yd[c] = (x + 2 * y) * bW[3 * p + c] + y * (thp[i] * D[3 * j + c] - thp[j] * D[3 * i + c]);
  double blk[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) blk[r + 4 * c] = (WAT(4 * i + r, 4 * j + c) - WAT(4 * i + c, 4 * j + r)) / 2;
  for (int r = 0; r < 3; ++r) {
    blk[r + 12] = yd[r];
    blk[3 + 4 * r] = -yd[r];
  }
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      Wd[(size_t)(4 * j + c) * n + 4 * i + r] = blk[r + 4 * c];
      Wd[(size_t)(4 * i + r) * n + 4 * j + c] = blk[r + 4 * c];  // transpose (:416-417)
    }
}

// diagonal blocks (:419-434): one warp per block row; then diag3[i][9] holds the 3x3 corner for the mean
__global__ void dual_diag_kernel(const double* __restrict__ W, const double* __restrict__ thp, int N,
                                 double* __restrict__ Wd, double* __restrict__ diag3) {
  const int n = 4 * N + 4, N1 = N + 1;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= N1) return;
  double rs[4] = {0, 0, 0, 0};
  for (int j = lane; j < N1; j += 32) {
    if (j == i) continue;  // the diagonal block is still zero when the reference sums the block row
    const double tj = thp[j];
    for (int r = 0; r < 4; ++r) rs[r] += tj * Wd[(size_t)(4 * j + 3) * n + 4 * i + r];
  }
  for (int r = 0; r < 4; ++r)
    for (int off = 16; off > 0; off >>= 1) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], off);
  if (lane == 0) {
    const double ti = thp[i];
    double blk[16];
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) blk[r + 4 * c] = WAT(4 * i + r, 4 * i + c);
    for (int r = 0; r < 4; ++r) blk[r + 12] = -ti * rs[r];
    for (int c = 0; c < 4; ++c) blk[3 + 4 * c] = -ti * rs[c];
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) Wd[(size_t)(4 * i + c) * n + 4 * i + r] = blk[r + 4 * c];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) diag3[9 * (size_t)i + r + 3 * c] = blk[r + 4 * c];
  }
}

// one CTA: mean of the 3x3 corners, subtracted from every diagonal block (:435-445)
__global__ void __launch_bounds__(256) dual_mean_kernel(const double* __restrict__ diag3, int N,
                                                       double* __restrict__ Wd) {
  __shared__ double sh[256];
  __shared__ double mean[9];
  const int n = 4 * N + 4, N1 = N + 1;
  for (int e = 0; e < 9; ++e) {
    double s = 0;
    for (int i = threadIdx.x; i < N1; i += 256) s += diag3[9 * (size_t)i + e];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) mean[e] = sh[0] / (double)N1;
    __syncthreads();
  }
  for (int t = threadIdx.x; t < 9 * N1; t += 256) {
    const int i = t / 9, e = t % 9, r = e % 3, c = e / 3;
    Wd[(size_t)(4 * i + c) * n + 4 * i + r] -= mean[e];
  }
}

__global__ void affine_out_kernel(const double* __restrict__ Minit, const double* __restrict__ Wd, long long nn,
                                  double* __restrict__ Maff) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nn) Maff[e] = Minit[e] + Wd[e];
}

__global__ void update_kernel(double* __restrict__ M, const double* __restrict__ Maff,
                              const double* __restrict__ Mpsd, long long nn, double gamma_tau) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nn) M[e] += gamma_tau * (Maff[e] - Mpsd[e]);
}

inline int grid1(long long items, int threads) { return (int)((items + threads - 1) / threads); }

}  // namespace

// ---- host-side driver objects ------------------------------------------------------------------------------------
struct CertWork {
  // device pointers carved from one scratch allocation
  double *src, *dst, *theta, *thp, *R, *q, *offA, *diagB, *lam, *mu_terms, *top, *mu;
  double *M_init, *M, *B, *w, *Vs, *M_psd, *W, *Wd, *M_aff, *bW, *D, *diag3, *work;
  int *first_pos, *info;
  int lwork;
};

size_t cert_scratch_bytes(int N, int lwork) {
  const size_t n = 4 * (size_t)N + 4, nn = n * n, N1 = (size_t)N + 1;
  size_t d = 0;
  d += 3 * N + 3 * N + N + N1 + 9 + 4;              // src dst theta thp R q
  d += 16 * (size_t)N * 3 + N + 16 + 1;             // offA diagB lam mu_terms top mu
  d += nn * 8;                                      // M_init M B Vs M_psd W Wd M_aff
  d += n;                                           // w
  d += 3 * N1 * (N1 - 1) / 2 + 3 * N1 + 9 * N1;     // bW D diag3
  d += (size_t)lwork;
  return d * sizeof(double) + 64 * 34 + 256;
}

static CertWork carve(void* scratch, int N, int lwork) {
  const size_t n = 4 * (size_t)N + 4, nn = n * n, N1 = (size_t)N + 1;
  char* p = (char*)scratch;
  auto take = [&](size_t doubles) {
    double* r = (double*)p;
    p += ((doubles * sizeof(double) + 63) / 64) * 64;
    return r;
  };
  CertWork w;
  w.src = take(3 * N); w.dst = take(3 * N); w.theta = take(N); w.thp = take(N1); w.R = take(9); w.q = take(4);
  w.offA = take(16 * (size_t)N); w.diagB = take(16 * (size_t)N); w.lam = take(16 * (size_t)N);
  w.mu_terms = take(N); w.top = take(16); w.mu = take(1);
  w.M_init = take(nn); w.M = take(nn); w.B = take(nn); w.Vs = take(nn); w.M_psd = take(nn); w.W = take(nn);
  w.Wd = take(nn); w.M_aff = take(nn); w.w = take(n);
  w.bW = take(3 * N1 * (N1 - 1) / 2); w.D = take(3 * N1); w.diag3 = take(9 * N1);
  w.work = take((size_t)lwork);
  w.first_pos = (int*)p;
  w.info = w.first_pos + 1;
  w.lwork = lwork;
  return w;
}

// getOptimalDualProjection on device buffers (W -> Wd); 5 launches
static int launch_dual_projection(const CertWork& w, int N, cudaStream_t st) {
  const int N1 = N + 1;
  const size_t n = 4 * (size_t)N + 4;
  cudaMemsetAsync(w.Wd, 0, n * n * sizeof(double), st);
  dim3 gp(grid1(N1, 128), N);
  dual_bw_kernel<<<gp, 128, 0, st>>>(w.W, w.thp, N, w.bW);
  dual_sums_kernel<<<grid1(3 * N1, 128), 128, 0, st>>>(w.bW, w.thp, N, w.D);
  dual_off_kernel<<<gp, 128, 0, st>>>(w.W, w.thp, w.bW, w.D, N, w.Wd);
  dual_diag_kernel<<<grid1(N1, 4), 128, 0, st>>>(w.W, w.thp, N, w.Wd, w.diag3);
  dual_mean_kernel<<<1, 256, 0, st>>>(w.diag3, N, w.Wd);
  return 5;
}

// Returns TZR_OK or a negative status; err receives a message.  All pointers host.  mode: 0 full certify,
// 1 initial matrix only (M_init_out n*n, mu_out), 2 dual projection only (W_in n*n -> Wd_out n*n; theta = N values).
int certify_device(int mode, double noise_bound, double cbar2, double sub_optimality, double max_iterations,
                   double gamma_tau, const double* R_cm, const double* src, const double* dst, const double* theta,
                   int N, int* is_optimal, double* best_subopt, int* n_iters, double* traj, int traj_cap,
                   double* M_init_out, double* mu_out, const double* W_in, double* Wd_out, void** scratch,
                   size_t* scratch_cap, void** solver_handle, int64_t* launches, cudaStream_t st, std::string* err) {
  const size_t n = 4 * (size_t)N + 4, nn = n * n;
  if (n > 32768) return TZR_ERR_TOO_LARGE;
  int lwork = 0;
  Cusolver& cs = cusolver();
  cusolverDnHandle_t& handle = *reinterpret_cast<cusolverDnHandle_t*>(solver_handle);  // owned by the context
  if (mode == 0) {
    if (!cs.ok) {
      *err = cs.err;
      return TZR_ERR_UNSUPPORTED;
    }
    if (!handle && cs.create(&handle) != CUSOLVER_STATUS_SUCCESS) {
      *err = "cusolverDnCreate failed";
      return TZR_ERR_CUDA;
    }
    cs.set_stream(handle, st);
    if (cs.syevd_buf(handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, nullptr, (int)n, nullptr,
                     &lwork) != CUSOLVER_STATUS_SUCCESS) {
      *err = "cusolverDnDsyevd_bufferSize failed";
      return TZR_ERR_CUDA;
    }
  }
  const size_t need = cert_scratch_bytes(N, lwork);
  if (need > *scratch_cap) {
    if (*scratch) cudaFree(*scratch);
    *scratch = nullptr;
    *scratch_cap = 0;
    if (cudaMalloc(scratch, need) != cudaSuccess) {
      *err = "cudaMalloc of the certifier workspace failed";
      cudaGetLastError();
      return TZR_ERR_ALLOC;
    }
    *scratch_cap = need;
  }
  CertWork w = carve(*scratch, N, lwork);
  std::vector<double> thp((size_t)N + 1);
  thp[0] = 1.0;
  for (int i = 0; i < N; ++i) thp[i + 1] = theta[i];
  cudaMemcpyAsync(w.thp, thp.data(), (N + 1) * sizeof(double), cudaMemcpyHostToDevice, st);
  if (mode == 2) {
    cudaMemcpyAsync(w.W, W_in, nn * sizeof(double), cudaMemcpyHostToDevice, st);
    *launches += launch_dual_projection(w, N, st);
    cudaMemcpyAsync(Wd_out, w.Wd, nn * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return TZR_ERR_CUDA;
    return TZR_OK;
  }
  // quaternion of R (Eigen::Quaterniond(R).normalize(), :66-69) on the host: 20 flops
  double q[4];
  {
    const double* R = R_cm;  // column-major: R(r,c) = R[r + 3c]
    auto Rm = [&](int r, int c) { return R[r + 3 * c]; };
    const double t = Rm(0, 0) + Rm(1, 1) + Rm(2, 2);
    double x, y, z, ww;
    if (t > 0) {
      double s = std::sqrt(t + 1.0);
      ww = 0.5 * s;
      s = 0.5 / s;
      x = (Rm(2, 1) - Rm(1, 2)) * s;
      y = (Rm(0, 2) - Rm(2, 0)) * s;
      z = (Rm(1, 0) - Rm(0, 1)) * s;
    } else {
      int i = 0;
      if (Rm(1, 1) > Rm(0, 0)) i = 1;
      if (Rm(2, 2) > Rm(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      double s = std::sqrt(Rm(i, i) - Rm(j, j) - Rm(k, k) + 1.0);
      double v[3];
      v[i] = 0.5 * s;
      s = 0.5 / s;
      ww = (Rm(k, j) - Rm(j, k)) * s;
      v[j] = (Rm(j, i) + Rm(i, j)) * s;
      v[k] = (Rm(k, i) + Rm(i, k)) * s;
      x = v[0]; y = v[1]; z = v[2];
    }
    const double nrm = std::sqrt(x * x + y * y + z * z + ww * ww);
    q[0] = x / nrm; q[1] = y / nrm; q[2] = z / nrm; q[3] = ww / nrm;
  }
  cudaMemcpyAsync(w.src, src, 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(w.dst, dst, 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(w.theta, theta, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(w.R, R_cm, 9 * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(w.q, q, 4 * sizeof(double), cudaMemcpyHostToDevice, st);
  const double nbs = cbar2 * noise_bound * noise_bound;
  cert_blocks_kernel<<<grid1(N, 128), 128, 0, st>>>(w.src, w.dst, w.theta, N, w.R, w.q, nbs, w.offA, w.diagB, w.lam,
                                                    w.mu_terms);
  cert_reduce_kernel<<<1, 256, 0, st>>>(w.lam, w.mu_terms, N, w.top, w.mu);
  cert_fill_kernel<<<grid1((long long)nn, 256), 256, 0, st>>>(N, w.offA, w.diagB, w.lam, w.top, w.mu, w.M_init);
  *launches += 3;
  double mu = 0;
  cudaMemcpyAsync(&mu, w.mu, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (mode == 1) {
    cudaMemcpyAsync(M_init_out, w.M_init, nn * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return TZR_ERR_CUDA;
    *mu_out = mu;
    return TZR_OK;
  }
  cudaMemcpyAsync(w.M, w.M_init, nn * sizeof(double), cudaMemcpyDeviceToDevice, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return TZR_ERR_CUDA;
  const int ge = grid1((long long)nn, 256);
  double best = INFINITY;
  int iters = 0;
  std::vector<double> wh(n);
  for (size_t iter = 0; (double)iter < max_iterations; ++iter) {
    // nearest PSD matrix (linalg.h:84-99)
    sym_kernel<<<ge, 256, 0, st>>>(w.M, (int)n, w.B);
    if (cs.syevd(handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, w.B, (int)n, w.w, w.work,
                 lwork, w.info) != CUSOLVER_STATUS_SUCCESS) {
      *err = "cusolverDnDsyevd failed";
      return TZR_ERR_CUDA;
    }
    first_pos_kernel<<<1, 32, 0, st>>>(w.w, (int)n, w.first_pos);
    scale_pos_kernel<<<ge, 256, 0, st>>>(w.B, w.w, (int)n, w.Vs);
    dim3 gg((unsigned)((n + kGT - 1) / kGT), (unsigned)((n + kGT - 1) / kGT));
    psd_gemm_kernel<<<gg, 256, 0, st>>>(w.Vs, w.B, (int)n, w.first_pos, w.M_psd);
    affine_in_kernel<<<ge, 256, 0, st>>>(w.M_psd, w.M, w.M_init, (long long)nn, w.W);
    *launches += 5 + launch_dual_projection(w, N, st);
    affine_out_kernel<<<ge, 256, 0, st>>>(w.M_init, w.Wd, (long long)nn, w.M_aff);
    // sub-optimality gap (:192-231): smallest eigenvalue of sym(M_affine)
    sym_kernel<<<ge, 256, 0, st>>>(w.M_aff, (int)n, w.B);
    if (cs.syevd(handle, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, w.B, (int)n, w.w, w.work,
                 lwork, w.info) != CUSOLVER_STATUS_SUCCESS) {
      *err = "cusolverDnDsyevd (values) failed";
      return TZR_ERR_CUDA;
    }
    *launches += 2;
    double min_eig = 0;
    int info = 0;
    cudaMemcpyAsync(&min_eig, w.w, sizeof(double), cudaMemcpyDeviceToHost, st);  // ascending: w[0] is the minimum
    cudaMemcpyAsync(&info, w.info, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) {
      *err = cudaGetErrorString(cudaGetLastError());
      return TZR_ERR_CUDA;
    }
    double gap;
    if (info != 0) gap = INFINITY;  // "Failed to find the minimal eigenvalue" (:221-225)
    else if (min_eig > 0) gap = 0;
    else gap = (-min_eig * (N + 1)) / mu;
    if ((int)iter < traj_cap && traj) traj[iter] = gap;
    iters = (int)iter + 1;
    if (gap < best) best = gap;
    if (gap < sub_optimality) break;
    update_kernel<<<ge, 256, 0, st>>>(w.M, w.M_aff, w.M_psd, (long long)nn, gamma_tau);
    *launches += 1;
  }
  *is_optimal = best < sub_optimality ? 1 : 0;
  *best_subopt = best;
  *n_iters = iters;
  if (cudaStreamSynchronize(st) != cudaSuccess) return TZR_ERR_CUDA;
  return TZR_OK;
}

void certify_release(void* solver_handle) {
  if (solver_handle && cusolver().ok) cusolver().destroy((cusolverDnHandle_t)solver_handle);
}

}  // namespace tzr
