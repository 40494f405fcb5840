// Correspondence generation upstream of solve(): teaser::Matcher::calculateCorrespondences
// (reference teaser/src/matcher.cc:21-337) on the device.
//
//   nn_kernel            exact FP32 1-NN of every query feature among a database of features, brute force, with
//                        flann::L2<float>'s accumulation order (groups of four, then the tail) so that distances —
//                        and therefore the argmin — are the ones the reference's exact KD-tree search compares.
//                        Result merged across database segments with one 64-bit atomicMin on (distance bits, index):
//                        the lowest index wins among equal distances.
//   mean/normalize       normalizePoints (matcher.cc:55-113): float, sequential accumulation for the mean.
//   corres_kernel        i_to_j / corres_ij ++ corres_ji / cross check (matcher.cc:152-215) with an ordered compaction.
//   tuple_kernel         the tuple test (matcher.cc:223-281), one thread per trial, counter-based draws.
//   finalize             swap back, sort (CUB radix sort, the only library call), unique (matcher.cc:283-297).
//
// The kNN over 33-D descriptors is 0.8 G multiply-adds at 5000 x 5000: microseconds on the FP32 pipes and exact,
// which the tensor-core route (TF32/BF16 products) is not — it would need the same "filter + exact recheck" split as
// the graph stage for no measurable gain at these sizes.
#include <algorithm>

#include <cub/cub.cuh>

#include "tzr_internal.cuh"

namespace tzr {

namespace {

constexpr int kNnTile = 64;      // queries per CTA and database points per smem tile
constexpr int kNnThreads = 256;  // 16 x 16 threads, 4 x 4 pairs each

__device__ __forceinline__ unsigned long long pack_dist(float d, int idx) {
  unsigned int bits = (d != d) ? 0x7fc00000u : __float_as_uint(d);  // distances are >= 0: bit order == value order
  return ((unsigned long long)bits << 32) | (unsigned int)idx;
}

// dynamic smem: Qs[dim][64] | Ds[dim][64]
__global__ void __launch_bounds__(kNnThreads) nn_kernel(const float* __restrict__ query, int nq,
                                                        const float* __restrict__ db, int ndb, int dim,
                                                        int db_per_seg, unsigned long long* __restrict__ best) {
  extern __shared__ float nn_smem[];
  float* Qs = nn_smem;
  float* Ds = nn_smem + (size_t)dim * kNnTile;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * kNnTile;
  const int seg_lo = blockIdx.y * db_per_seg, seg_hi = min(ndb, seg_lo + db_per_seg);
  // transposing loads: consecutive threads take consecutive points (conflict-free smem stores; the strided global
  // reads of one 64 x dim tile stay within 64*dim*4 contiguous bytes and are served by L1 after the first touch)
  for (int e = tid; e < dim * kNnTile; e += kNnThreads) {
    const int k = e / kNnTile, q = e - k * kNnTile;
    Qs[e] = (q0 + q < nq) ? query[(size_t)(q0 + q) * dim + k] : 0.f;
  }
  unsigned long long mine[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  const int dim4 = dim & ~3;
  for (int d0 = seg_lo; d0 < seg_hi; d0 += kNnTile) {
    __syncthreads();
    for (int e = tid; e < dim * kNnTile; e += kNnThreads) {
      const int k = e / kNnTile, p = e - k * kNnTile;
      Ds[e] = (d0 + p < seg_hi) ? db[(size_t)(d0 + p) * dim + k] : 0.f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    for (int k = 0; k < dim4; k += 4) {
      float g[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + (k + kk) * kNnTile + ty * 4);
        const float4 dv = *reinterpret_cast<const float4*>(Ds + (k + kk) * kNnTile + tx * 4);
        const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float df = __fsub_rn(qa[u], da[v]);
            const float sq = __fmul_rn(df, df);
            g[u][v] = (kk == 0) ? sq : __fadd_rn(g[u][v], sq);  // ((d0^2 + d1^2) + d2^2) + d3^2
          }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = __fadd_rn(acc[u][v], g[u][v]);
    }
    for (int k = dim4; k < dim; ++k) {
      const float4 qv = *reinterpret_cast<const float4*>(Qs + k * kNnTile + ty * 4);
      const float4 dv = *reinterpret_cast<const float4*>(Ds + k * kNnTile + tx * 4);
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float df = __fsub_rn(qa[u], da[v]);
          acc[u][v] = __fadd_rn(acc[u][v], __fmul_rn(df, df));
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int p = d0 + tx * 4 + v;
      if (p < seg_hi) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned long long key = pack_dist(acc[u][v], p);
          mine[u] = key < mine[u] ? key : mine[u];
        }
      }
    }
  }
  // the 16 threads sharing ty are 16 consecutive lanes of one warp
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    unsigned long long m = mine[u];
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, off);
      m = o < m ? o : m;
    }
    const int q = q0 + ty * 4 + u;
    if (tx == 0 && q < nq) atomicMin(&best[q], m);
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, long long n, unsigned long long v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// lanes 0..5: (cloud, component) -> sequential float sum in index order, then / n  (matcher.cc:62-72)
__global__ void mean_kernel(const float* pts_a, int na, const float* pts_b, int nb, float* mean6) {
  const int lane = threadIdx.x;
  if (lane >= 6) return;
  const int c = lane / 3, k = lane - 3 * c;
  const float* p = c ? pts_b : pts_a;
  const int n = c ? nb : na;
  float s = 0.f;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[3 * (size_t)(i + u) + k];
#pragma unroll
    for (int u = 0; u < 8; ++u) s = __fadd_rn(s, v[u]);
  }
  for (; i < n; ++i) s = __fadd_rn(s, p[3 * (size_t)i + k]);
  mean6[lane] = __fdiv_rn(s, (float)n);
}

// subtract the mean in place and fold max ||p|| into scale_bits (non-negative floats: uint order == float order)
__global__ void center_kernel(float* pts_a, int na, float* pts_b, int nb, const float* mean6, unsigned int* scale_bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float t = 0.f;
  if (i < na + nb) {
    const int c = i >= na;
    float* p = (c ? pts_b + 3 * (size_t)(i - na) : pts_a + 3 * (size_t)i);
    const float x = __fsub_rn(p[0], mean6[3 * c]), y = __fsub_rn(p[1], mean6[3 * c + 1]),
                z = __fsub_rn(p[2], mean6[3 * c + 2]);
    p[0] = x;
    p[1] = y;
    p[2] = z;
    t = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    if (!(t == t)) t = 0.f;  // NaN never wins `temp > max_scale`
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, off));
  if ((threadIdx.x & 31) == 0 && t > 0.f) atomicMax(scale_bits, __float_as_uint(t));
}

__global__ void rescale_kernel(float* pts_a, int na, float* pts_b, int nb, const unsigned int* scale_bits,
                               int use_absolute_scale, float* gscale_out) {
  const float g = use_absolute_scale ? 1.0f : __uint_as_float(*scale_bits);
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0) *gscale_out = g;
  if (g == 1.0f) return;
  if (i < 3ll * na) pts_a[i] = __fdiv_rn(pts_a[i], g);
  else if (i < 3ll * (na + nb)) pts_b[i - 3ll * na] = __fdiv_rn(pts_b[i - 3ll * na], g);
}

__global__ void hit_kernel(const unsigned long long* best_j, int nj, uint8_t* hit) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nj) hit[(unsigned int)best_j[j]] = 1;
}

// Block-wide ordered compaction helper: returns the exclusive prefix of `flag` over the CTA's threads and the total.
__device__ int block_excl_scan(int flag, int* total, int* s_warp /*[33]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned int bal = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  __syncthreads();
  if (lane == 0) s_warp[wid] = __popc(bal);
  __syncthreads();
  if (wid == 0) {
    int v = lane < nw ? s_warp[lane] : 0;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    s_warp[lane] = incl - v;
    if (lane == 31) s_warp[32] = incl;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + in_warp;
}

// One CTA.  Slots (matcher.cc:152-215): without cross check, slot i < n_i is (i, i_to_j[i]) when i was hit, slot
// n_i + j is (nn_of_j[j], j); with cross check, slot i is (i, i_to_j[i]) when hit and nn_of_j[i_to_j[i]] == i.
// corres[] receives the valid slots in slot order (the order the reference's vectors have), ncorr their number.
__global__ void __launch_bounds__(1024) corres_kernel(const unsigned long long* best_j, const unsigned long long* best_i,
                                                      const uint8_t* hit, int n_i, int n_j, int crosscheck,
                                                      int2* corres, int* ncorr) {
  __shared__ int s_warp[33];
  const int slots = crosscheck ? n_i : n_i + n_j;
  int base = 0;
  for (int s0 = 0; s0 < slots; s0 += blockDim.x) {
    const int s = s0 + threadIdx.x;
    int flag = 0;
    int2 pr = make_int2(0, 0);
    if (s < slots) {
      if (s < n_i) {
        if (hit[s]) {
          const int j = (int)(unsigned int)best_i[s];
          pr = make_int2(s, j);
          flag = crosscheck ? ((int)(unsigned int)best_j[j] == s) : 1;
        }
      } else {
        const int j = s - n_i;
        pr = make_int2((int)(unsigned int)best_j[j], j);
        flag = 1;
      }
    }
    int total;
    const int pos = block_excl_scan(flag, &total, s_warp);
    if (flag) corres[base + pos] = pr;
    base += total;
  }
  if (threadIdx.x == 0) *ncorr = base;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ float dist3(const float* a, const float* b) {
  const float x = __fsub_rn(a[0], b[0]), y = __fsub_rn(a[1], b[1]), z = __fsub_rn(a[2], b[2]);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// matcher.cc:223-281: 100 * ncorr trials; a passing triple keeps its three correspondences.
__global__ void tuple_kernel(const int2* corres, const int* ncorr_p, const float* pts_i, const float* pts_j,
                             float scale, uint64_t seed, uint8_t* keep) {
  const long long ncorr = *ncorr_p;
  const long long trials = ncorr * 100;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < trials;
       t += (long long)gridDim.x * blockDim.x) {
    long long r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = (long long)(splitmix64(seed + 3ull * (uint64_t)t + k) >> 33) % ncorr;
    const int2 c0 = corres[r[0]], c1 = corres[r[1]], c2 = corres[r[2]];
    const float *a0 = pts_i + 3 * (size_t)c0.x, *a1 = pts_i + 3 * (size_t)c1.x, *a2 = pts_i + 3 * (size_t)c2.x;
    const float *b0 = pts_j + 3 * (size_t)c0.y, *b1 = pts_j + 3 * (size_t)c1.y, *b2 = pts_j + 3 * (size_t)c2.y;
    const float li0 = dist3(a0, a1), li1 = dist3(a1, a2), li2 = dist3(a2, a0);
    const float lj0 = dist3(b0, b1), lj1 = dist3(b1, b2), lj2 = dist3(b2, b0);
    if ((__fmul_rn(li0, scale) < lj0) && (lj0 < __fdiv_rn(li0, scale)) && (__fmul_rn(li1, scale) < lj1) &&
        (lj1 < __fdiv_rn(li1, scale)) && (__fmul_rn(li2, scale) < lj2) && (lj2 < __fdiv_rn(li2, scale))) {
      keep[r[0]] = 1;
      keep[r[1]] = 1;
      keep[r[2]] = 1;
    }
  }
}

// keys[c] = (first << 32 | second) of kept correspondences, in the caller's (src, dst) order; ~0 elsewhere
__global__ void keys_kernel(const int2* corres, const int* ncorr_p, const uint8_t* keep, int use_keep, int swapped,
                            unsigned long long* keys, int cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cap) return;
  unsigned long long k = ~0ull;
  if (c < *ncorr_p && (!use_keep || keep[c])) {
    const int2 pr = corres[c];
    const unsigned int a = swapped ? pr.y : pr.x, b = swapped ? pr.x : pr.y;
    k = ((unsigned long long)a << 32) | b;
  }
  keys[c] = k;
}

// One CTA: std::unique over the sorted keys (matcher.cc:295-296) -> pairs[2*c], pairs[2*c+1]; count.
__global__ void __launch_bounds__(1024) unique_kernel(const unsigned long long* keys, int cap, int32_t* pairs,
                                                      int* count) {
  __shared__ int s_warp[33];
  int base = 0;
  for (int s0 = 0; s0 < cap; s0 += blockDim.x) {
    const int s = s0 + threadIdx.x;
    int flag = 0;
    unsigned long long k = ~0ull;
    if (s < cap) {
      k = keys[s];
      flag = (k != ~0ull) && (s == 0 || keys[s - 1] != k);
    }
    int total;
    const int pos = block_excl_scan(flag, &total, s_warp);
    if (flag) {
      pairs[2 * (size_t)(base + pos)] = (int32_t)(k >> 32);
      pairs[2 * (size_t)(base + pos) + 1] = (int32_t)(k & 0xffffffffu);
    }
    base += total;
  }
  if (threadIdx.x == 0) *count = base;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int launch_feature_nn(const float* query, int nq, const float* db, int ndb, int dim, unsigned long long* best,
                      int num_sms, cudaStream_t st) {
  const size_t smem = (size_t)2 * dim * kNnTile * sizeof(float);
  if (smem > 48 * 1024)  // per device and cheap: no process-wide "already set" flag (one context per GPU is allowed)
    cudaFuncSetAttribute(nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kMatchMaxDim * kNnTile * 4);
  fill_u64_kernel<<<std::min(1024, (nq + 255) / 256), 256, 0, st>>>(best, nq, ~0ull);
  const int qt = (nq + kNnTile - 1) / kNnTile;
  const int db_tiles = (ndb + kNnTile - 1) / kNnTile;
  int segs = std::max(1, std::min(db_tiles, (4 * num_sms + qt - 1) / qt));  // ~4 CTAs per SM in flight
  const int tiles_per_seg = (db_tiles + segs - 1) / segs;
  segs = (db_tiles + tiles_per_seg - 1) / tiles_per_seg;
  nn_kernel<<<dim3(qt, segs), kNnThreads, smem, st>>>(query, nq, db, ndb, dim, tiles_per_seg * kNnTile, best);
  return 2;
}

// Scratch layout (all device): best_j[n_j] u64 | best_i[n_i] u64 | keys[cap] u64 | keys_sorted[cap] u64 |
// corres[cap] int2 | hit[n_i] | keep[cap] | mean6 | scale_bits | ncorr | cub temp
size_t match_scratch_bytes(int ns, int nd) {
  const size_t n_i = std::max(ns, nd), n_j = std::min(ns, nd), cap = n_i + n_j;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                 (int)cap);
  return align256(n_j * 8) + align256(n_i * 8) + 2 * align256(cap * 8) + align256(cap * 8) + align256(n_i) +
         align256(cap) + 256 + align256(cub_bytes) + 1024;
}

// All inputs device-resident; pts are normalised IN PLACE (pass copies).  pairs: (ns + nd) x 2 int32; count and
// gscale: device scalars.  Returns the number of kernel launches, or a negative tzr_status.
int launch_match(float* src_pts, int ns, float* dst_pts, int nd, const float* src_feat, const float* dst_feat, int dim,
                 int use_absolute_scale, int use_crosscheck, int use_tuple_test, float tuple_scale, uint64_t seed,
                 void* scratch, int32_t* pairs, int* count, float* gscale, int num_sms, cudaStream_t st) {
  if (dim < 1 || dim > kMatchMaxDim || ns < 1 || nd < 1) return TZR_ERR_INVALID_ARG;
  const bool swapped = nd > ns;  // the larger cloud is "i" (matcher.cc:121-126)
  const int n_i = swapped ? nd : ns, n_j = swapped ? ns : nd;
  const float* feat_i = swapped ? dst_feat : src_feat;
  const float* feat_j = swapped ? src_feat : dst_feat;
  float* pts_i = swapped ? dst_pts : src_pts;
  float* pts_j = swapped ? src_pts : dst_pts;
  const int cap = n_i + n_j;
  char* w = (char*)scratch;
  auto take = [&](size_t bytes) {
    char* p = w;
    w += align256(bytes);
    return p;
  };
  unsigned long long* best_j = (unsigned long long*)take((size_t)n_j * 8);
  unsigned long long* best_i = (unsigned long long*)take((size_t)n_i * 8);
  unsigned long long* keys = (unsigned long long*)take((size_t)cap * 8);
  unsigned long long* keys_sorted = (unsigned long long*)take((size_t)cap * 8);
  int2* corres = (int2*)take((size_t)cap * 8);
  uint8_t* hit = (uint8_t*)take((size_t)n_i);
  uint8_t* keep = (uint8_t*)take((size_t)cap);
  float* mean6 = (float*)take(256);
  unsigned int* scale_bits = (unsigned int*)(mean6 + 8);
  int* ncorr = (int*)(mean6 + 9);
  void* cub_temp = (void*)w;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                 cap);
  int nl = 0;
  const bool tuple = use_tuple_test && tuple_scale != 0.f;
  // normalizePoints only feeds the tuple test, but global_scale_ is part of the observable state: always computed
  cudaMemsetAsync(mean6, 0, 64, st);
  cudaMemsetAsync(hit, 0, (size_t)n_i, st);
  mean_kernel<<<1, 32, 0, st>>>(src_pts, ns, dst_pts, nd, mean6);
  center_kernel<<<(ns + nd + 255) / 256, 256, 0, st>>>(src_pts, ns, dst_pts, nd, mean6, scale_bits);
  rescale_kernel<<<(3 * (ns + nd) + 255) / 256, 256, 0, st>>>(src_pts, ns, dst_pts, nd, scale_bits,
                                                             use_absolute_scale, gscale);
  nl += 3;
  nl += launch_feature_nn(feat_j, n_j, feat_i, n_i, dim, best_j, num_sms, st);  // NN of every j among the i features
  hit_kernel<<<(n_j + 255) / 256, 256, 0, st>>>(best_j, n_j, hit);
  // the reference searches the reverse NN lazily for hit i only (:157-161); computing all of them costs the same
  // launch and the unused ones are ignored by corres_kernel
  nl += launch_feature_nn(feat_i, n_i, feat_j, n_j, dim, best_i, num_sms, st);
  corres_kernel<<<1, 1024, 0, st>>>(best_j, best_i, hit, n_i, n_j, use_crosscheck, corres, ncorr);
  nl += 2;
  if (tuple) {
    cudaMemsetAsync(keep, 0, (size_t)cap, st);
    tuple_kernel<<<num_sms * 8, 256, 0, st>>>(corres, ncorr, pts_i, pts_j, tuple_scale, seed, keep);
    nl += 1;
  }
  keys_kernel<<<(cap + 255) / 256, 256, 0, st>>>(corres, ncorr, keep, tuple ? 1 : 0, swapped ? 1 : 0, keys, cap);
  cub::DeviceRadixSort::SortKeys(cub_temp, cub_bytes, keys, keys_sorted, cap, 0, 64, st);
  unique_kernel<<<1, 1024, 0, st>>>(keys_sorted, cap, pairs, count);
  nl += 3;
  return nl;
}

}  // namespace tzr
