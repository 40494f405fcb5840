// Stage 1 of solve(): TIMs + scale-consistency test + inlier graph, fused.
//
// Replaces (reference, /root/reference):
//   RobustRegistrationSolver::computeTIMs            teaser/src/registration.cc:512-551  (x2)
//   ScaleInliersSelector::solveForScale              teaser/src/registration.cc:427-443
//   inlier_graph_.addEdge loop                       teaser/src/registration.cc:614-619
//
// The reference materialises 2 x (3 x K) doubles of TIMs, 2 x K norms and a K-byte mask
// (K = N(N-1)/2; 65 B per pair).  Here a pair is a register-resident predicate and the only
// output is one bit in a packed symmetric adjacency bitset.
//
// Predicate (must be bit-identical to the reference's IEEE-double evaluation without FMA
// contraction):  | sqrt(|s_j-s_i|^2) - sqrt(|d_j-d_i|^2) | <= beta.
// FP64 sqrt makes a pure-double kernel FP64-issue-bound, so every pair is first classified by a
// sqrt-free FP32 interval test on centred single-precision copies of the points:
//     with a=|ds|^2, b=|dd|^2, t=a-b, s=a+b:   |sqrt(a)-sqrt(b)| <= g  <=>  t^2 <= 2 g^2 s - g^4   (s >= g^2)
// evaluated for g = beta-delta ("surely an edge") and g = beta+delta ("surely not"), where delta
// bounds every FP32 error of the pipeline (conversion, differences, squares).
// Only pairs inside the 2*delta band (typically < 1e-4 of all pairs) are re-evaluated with the
// exact double sequence, so the bitset is identical to a pure-FP64 evaluation (verified on device
// by the TZR_FLAG_VERIFY mode and against the oracle in tests/).
#include "tzr_internal.cuh"

namespace tzr {

// ------------------------------------------------------------------------------------------------
// exact predicate: the reference's operation sequence in IEEE double, no contraction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tim_norm_exact(const double* __restrict__ p, int i, int j) {
  const double dx = __dsub_rn(p[3 * j + 0], p[3 * i + 0]);
  const double dy = __dsub_rn(p[3 * j + 1], p[3 * i + 1]);
  const double dz = __dsub_rn(p[3 * j + 2], p[3 * i + 2]);
  // src.array().square().colwise().sum(): (x^2 + y^2) + z^2   (registration.cc:434-437)
  const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return __dsqrt_rn(s);
}

__device__ __forceinline__ bool edge_exact(const double* __restrict__ src, const double* __restrict__ dst, int i,
                                           int j, double beta) {
  const double d1 = tim_norm_exact(src, i, j);
  const double d2 = tim_norm_exact(dst, i, j);
  return fabs(__dsub_rn(d1, d2)) <= beta;  // (v1_dist - v2_dist).abs() <= beta   registration.cc:442
}

// ------------------------------------------------------------------------------------------------
// prep: bounding boxes, centred float copies, filter constants.  One CTA per problem.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256) prep_kernel(Batch bt) {
  const int b = blockIdx.x;
  const int n = bt.n;
  const double* src = bt.src + (size_t)b * n * 3;
  const double* dst = bt.dst + (size_t)b * n * 3;
  __shared__ double s_red[8][12];
  __shared__ int s_bad[8];
  __shared__ double s_c[6];
  double mn[6], mx[6];
  for (int k = 0; k < 6; ++k) {
    mn[k] = 1.0 / 0.0;
    mx[k] = -1.0 / 0.0;
  }
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    for (int k = 0; k < 3; ++k) {
      const double a = src[3 * i + k], c = dst[3 * i + k];
      bad |= !isfinite(a) | !isfinite(c);
      mn[k] = fmin(mn[k], a);
      mx[k] = fmax(mx[k], a);
      mn[3 + k] = fmin(mn[3 + k], c);
      mx[3 + k] = fmax(mx[3 + k], c);
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < 6; ++k) {
    mn[k] = warp_min(mn[k]);
    mx[k] = warp_max(mx[k]);
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    for (int k = 0; k < 6; ++k) {
      s_red[w][k] = mn[k];
      s_red[w][6 + k] = mx[k];
    }
    s_bad[w] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int anybad = 0;
    for (int k = 0; k < 6; ++k) {
      double a = s_red[0][k], c = s_red[0][6 + k];
      for (int q = 1; q < 8; ++q) {
        a = fmin(a, s_red[q][k]);
        c = fmax(c, s_red[q][6 + k]);
      }
      mn[k] = a;
      mx[k] = c;
    }
    for (int q = 0; q < 8; ++q) anybad |= s_bad[q];
    double Ms = 0, Md = 0;
    GraphConsts gc;
    for (int k = 0; k < 3; ++k) {
      gc.cs[k] = 0.5 * (mn[k] + mx[k]);
      gc.cd[k] = 0.5 * (mn[3 + k] + mx[3 + k]);
      Ms = fmax(Ms, 0.5 * (mx[k] - mn[k]));
      Md = fmax(Md, 0.5 * (mx[3 + k] - mn[3 + k]));
      s_c[k] = gc.cs[k];
      s_c[3 + k] = gc.cd[k];
    }
    const double beta = bt.beta;
    gc.beta = beta;
    // delta: bound on the FP32 error of |D1 - D2| (<= ~70 u32 (Ms+Md)); 256 u32 (...) used.
    const double u32 = 5.9604644775390625e-08;  // 2^-24
    const double delta = 256.0 * u32 * (Ms + Md + beta);
    const double gam1 = beta - delta, gam2 = beta + delta;
    const double up = 1.0 + 9.5367431640625e-07, dn = 1.0 - 9.5367431640625e-07;  // 1 +- 2^-20
    int use64 = anybad || !(Ms < 1e8) || !(Md < 1e8) || !(gam2 > 1e-8) || !isfinite(beta) || (bt.flags_dbg & 1u);
    if (gam1 > 0) {
      gc.c1 = (float)(2.0 * gam1 * gam1 * dn);
      gc.g1 = (float)(gam1 * gam1 * gam1 * gam1 * up);
    } else {
      gc.c1 = 0.f;
      gc.g1 = __int_as_float(0x7f800000);  // +inf -> "surely an edge" can never fire
    }
    gc.c2 = (float)(2.0 * gam2 * gam2 * up);
    gc.g2 = (float)(gam2 * gam2 * gam2 * gam2 * dn);
    gc.smin = (float)(16.0 * gam2 * gam2 * up);
    gc.use_fp64 = use64;
    bt.gc[b] = gc;
    bt.n_edges2[b] = 0ull;
  }
  __syncthreads();
  float4* sf = bt.sf + (size_t)b * n;
  float4* df = bt.df + (size_t)b * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float4 a, c;
    a.x = (float)(src[3 * i + 0] - s_c[0]);
    a.y = (float)(src[3 * i + 1] - s_c[1]);
    a.z = (float)(src[3 * i + 2] - s_c[2]);
    a.w = 0.f;
    c.x = (float)(dst[3 * i + 0] - s_c[3]);
    c.y = (float)(dst[3 * i + 1] - s_c[4]);
    c.z = (float)(dst[3 * i + 2] - s_c[5]);
    c.w = 0.f;
    sf[i] = a;
    df[i] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// graph tile kernel: one CTA per 128x128 tile of the upper triangle (I <= J) of one problem.
// 8 warps; warp w owns rows [32*(w/2), +32) x cols [64*(w%2), +64) of the tile: lane l keeps the two
// column points (l, l+32) in registers and sweeps the 32 row points broadcast from shared memory.
// Row words come from __ballot_sync; the transposed (column) words are accumulated per lane, so the
// symmetric half of the bitset costs no extra predicate evaluations.
// ------------------------------------------------------------------------------------------------
struct PairEval {
  bool sure, amb;
};

__device__ __forceinline__ PairEval classify(const float4 is, const float4 id, const float4 js, const float4 jd,
                                             const float c1, const float g1, const float c2, const float g2,
                                             const float smin) {
  const float ax = js.x - is.x, ay = js.y - is.y, az = js.z - is.z;
  const float bx = jd.x - id.x, by = jd.y - id.y, bz = jd.z - id.z;
  const float a = fmaf(az, az, fmaf(ay, ay, ax * ax));
  const float b = fmaf(bz, bz, fmaf(by, by, bx * bx));
  const float t = a - b, s = a + b;
  const float tt = t * t;
  const float r1 = fmaf(c1, s, -g1);
  const float r2 = fmaf(c2, s, -g2);
  PairEval e;
  const bool big = s >= smin;
  e.sure = big && (tt <= r1);
  const bool non = big && (tt > r2);
  e.amb = !(e.sure || non);
  return e;
}

__global__ void __launch_bounds__(kGraphThreads, 3) graph_tile_kernel(Batch bt) {
  const int b = blockIdx.y;
  const int n = bt.n;
  const int nt = (n + kTile - 1) / kTile;
  // decode upper-triangular tile index -> (I, J), I <= J
  const int p = blockIdx.x;
  int I = (int)floor(((2.0 * nt + 1.0) - sqrt((2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)p)) * 0.5);
  if (I < 0) I = 0;
  while (I > 0 && (long long)I * nt - (long long)I * (I - 1) / 2 > p) --I;
  while ((long long)(I + 1) * nt - (long long)(I + 1) * I / 2 <= p) ++I;
  const int J = I + (p - (int)((long long)I * nt - (long long)I * (I - 1) / 2));

  __shared__ float4 s_is[kTile];
  __shared__ float4 s_id[kTile];
  __shared__ __align__(16) uint32_t s_row[kTile][4];
  __shared__ __align__(16) uint32_t s_col[kTile][4];

  const GraphConsts* gcp = bt.gc + b;
  const float c1 = gcp->c1, g1 = gcp->g1, c2 = gcp->c2, g2 = gcp->g2, smin = gcp->smin;
  const bool force64 = gcp->use_fp64 != 0;
  const double beta = gcp->beta;
  const bool verify = (bt.flags_dbg & 2u) != 0;

  const float4* sf = bt.sf + (size_t)b * n;
  const float4* df = bt.df + (size_t)b * n;
  const double* src = bt.src + (size_t)b * n * 3;
  const double* dst = bt.dst + (size_t)b * n * 3;

  const int tid = threadIdx.x;
  {
    const int t = tid & (kTile - 1);
    const int i = I * kTile + t;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < kTile)
      s_is[t] = (i < n) ? sf[i] : z;
    else
      s_id[t] = (i < n) ? df[i] : z;
  }
  const int w = tid >> 5, lane = tid & 31;
  const int ri = w >> 1, ch = w & 1;
  const int j0 = J * kTile + 64 * ch + lane, j1 = j0 + 32;
  const bool vj0 = j0 < n, vj1 = j1 < n;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 j0s = vj0 ? sf[j0] : z4, j0d = vj0 ? df[j0] : z4;
  const float4 j1s = vj1 ? sf[j1] : z4, j1d = vj1 ? df[j1] : z4;
  __syncthreads();

  uint32_t rowA = 0, rowB = 0, colA = 0, colB = 0;
  const int ibase = I * kTile + 32 * ri;
#pragma unroll
  for (int ii = 0; ii < 32; ++ii) {
    const int i = ibase + ii;
    if (i >= n) break;  // warp-uniform
    const float4 is = s_is[32 * ri + ii], id = s_id[32 * ri + ii];
    PairEval e0 = classify(is, id, j0s, j0d, c1, g1, c2, g2, smin);
    PairEval e1 = classify(is, id, j1s, j1d, c1, g1, c2, g2, smin);
    const bool ok0 = vj0 && (i != j0), ok1 = vj1 && (i != j1);
    bool p0 = ok0 && e0.sure, p1 = ok1 && e1.sure;
    bool a0 = ok0 && (e0.amb || force64), a1 = ok1 && (e1.amb || force64);
    if (__any_sync(0xffffffffu, a0 || a1)) {
      if (a0) p0 = edge_exact(src, dst, i, j0, beta);
      if (a1) p1 = edge_exact(src, dst, i, j1, beta);
      if (bt.rechecks) {
        const unsigned m0 = __ballot_sync(0xffffffffu, a0), m1 = __ballot_sync(0xffffffffu, a1);
        if (lane == 0) atomicAdd(bt.rechecks, (unsigned long long)(__popc(m0) + __popc(m1)));
      }
    }
    if (verify) {
      int bad = 0;
      if (ok0 && !a0) bad += (edge_exact(src, dst, i, j0, beta) != p0);
      if (ok1 && !a1) bad += (edge_exact(src, dst, i, j1, beta) != p1);
      if (bad) atomicAdd(bt.mismatches, (unsigned long long)bad);
    }
    const uint32_t m0 = __ballot_sync(0xffffffffu, p0);
    const uint32_t m1 = __ballot_sync(0xffffffffu, p1);
    if (lane == ii) {
      rowA = m0;
      rowB = m1;
    }
    colA |= p0 ? (1u << ii) : 0u;
    colB |= p1 ? (1u << ii) : 0u;
  }
  s_row[32 * ri + lane][2 * ch + 0] = rowA;
  s_row[32 * ri + lane][2 * ch + 1] = rowB;
  s_col[64 * ch + lane][ri] = colA;
  s_col[64 * ch + 32 + lane][ri] = colB;
  __syncthreads();

  uint32_t* adj32 = reinterpret_cast<uint32_t*>(bt.adj) + (size_t)b * n * pitch32(n);
  const int P32 = pitch32(n);
  if (tid < kTile) {
    const int row = I * kTile + tid;
    if (row < n) {
      const uint4 v = *reinterpret_cast<const uint4*>(&s_row[tid][0]);
      *reinterpret_cast<uint4*>(adj32 + (size_t)row * P32 + 4 * J) = v;
    }
  } else if (I != J) {
    const int t = tid - kTile;
    const int row = J * kTile + t;
    if (row < n) {
      const uint4 v = *reinterpret_cast<const uint4*>(&s_col[t][0]);
      *reinterpret_cast<uint4*>(adj32 + (size_t)row * P32 + 4 * I) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// degree kernel: deg[v] = popcount(row v); n_edges2[b] = sum of degrees.  One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) degree_kernel(Batch bt) {
  const int b = blockIdx.y;
  const int n = bt.n;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  __shared__ int s_sum[8];
  int d = 0;
  if (row < n) {
    const int P = pitch64(n);
    const uint64_t* r = bt.adj + ((size_t)b * n + row) * P;
    for (int x = lane; x < P; x += 32) d += __popcll(r[x]);
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) bt.deg[(size_t)b * n + row] = d;
  }
  if (lane == 0) s_sum[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int q = 0; q < 8; ++q) s += s_sum[q];
    if (s) atomicAdd(bt.n_edges2 + b, (unsigned long long)s);
  }
}

void launch_prep(const Batch& bt, cudaStream_t st) { prep_kernel<<<bt.B, 256, 0, st>>>(bt); }

void launch_graph(const Batch& bt, cudaStream_t st) {
  const int nt = (bt.n + kTile - 1) / kTile;
  dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)bt.B);
  graph_tile_kernel<<<grid, kGraphThreads, 0, st>>>(bt);
}

void launch_degree(const Batch& bt, cudaStream_t st) {
  dim3 grid((unsigned)((bt.n + 7) / 8), (unsigned)bt.B);
  degree_kernel<<<grid, 256, 0, st>>>(bt);
}

}  // namespace tzr
