// Stage 1 of solve() on the 5th-generation tensor cores: TIMs + scale-consistency test + inlier graph.
//
// Replaces (reference, /root/reference), exactly like graph_build.cu:
//   RobustRegistrationSolver::computeTIMs            teaser/src/registration.cc:512-551  (x2)
//   ScaleInliersSelector::solveForScale              teaser/src/registration.cc:427-443
//   inlier_graph_.addEdge loop                       teaser/src/registration.cc:614-619
//
// The CUDA-core kernel (graph_build.cu) spends ~20 issue slots per pair: twelve FP32 operations for the two squared
// TIM norms, two MUFU.SQRT, compares, ballots.  Here the squared norms come from the tensor cores instead:
//     a_ij = |s_i - s_j|^2 = n_i + n_j - 2 s_i.s_j      (and b_ij for the destination cloud)
// is a K = 3 (+2 for the norms) contraction.  Every centred coordinate (and every norm) is split into three tf32
// pieces h + m + l (11 significant bits each, so the split carries 33 bits of the FP64 value), and the products
// hh', hm', mh', mm', hl', lh' plus the six norm pieces are laid out as a K = 24 tf32 GEMM (three K = 8 tcgen05.mma
// steps per cloud): one 128x128 tile of a and of b lands in tensor memory per 6 MMAs, issued by one thread.  The
// operand tiles (tc_prep_kernel writes them once per problem in the exact shared-memory image the MMA descriptor
// wants: no-swizzle K-major planes) are staged by the TMA engine (cp.async.bulk -> UBLKCP) into a double-buffered
// shared-memory ring, completion on mbarriers.
//
// Epilogue (8 warps, tcgen05.ld 32 lanes x 32 columns): with t = a-b, s = a+b, q = sqrt(ab) (one MUFU per pair)
//     d = t^2 - beta^2 (s + 2q) = (sqrt a + sqrt b)^2 (g^2 - beta^2),      g = |sqrt a - sqrt b|
// so the pair is an edge iff d <= 0.  This form is well conditioned (both sides are equal at the threshold, so the
// FP32 evaluation error is a few ulp of beta^2 (sqrt a + sqrt b)^2, no cancellation of D^2-sized terms against
// beta^2).  The sign bit of d is shifted straight into the row word (no compare, no ballot); min |d| and min ab over
// the 32 pairs of a thread are tracked with two 3-input FMNMX, and only if min|d| <= theta or min ab <= prisk
// (prep_kernel) the warp revisits its chunk and re-evaluates the flagged pairs with the reference's
// exact FP64 sequence.  6 issue slots + 1 MUFU per pair instead of ~20 + 2.
// Output (packed symmetric bitset, fused degrees) is bit-identical to graph_build.cu's and to the oracle's.
#include "tc_ptx.cuh"
#include "tzr_internal.cuh"

namespace tzr {

using namespace tc;

constexpr int kTcEpiWarps = 8;                       // warp w: TMEM lanes 32*(w&3).., columns 64*(w>>2)..
constexpr int kTcThreads = 32 * (kTcEpiWarps + 1);   // + 1 producer warp (TMA + MMA issue by one elected lane)
constexpr int kTcPlaneBytes = 128 * 16;              // 128 rows x 4 tf32
constexpr int kTcCloudBytes = 6 * kTcPlaneBytes;     // 6 planes = K 24
constexpr int kTcRoleBytes = 2 * kTcCloudBytes;      // src + dst
constexpr int kTcBlockBytes = 2 * kTcRoleBytes;      // A role + B role of one 128-point block
constexpr int kTcSmemBytes = 4 * kTcRoleBytes + 256; // A x2, B x2, barriers

size_t tc_operand_bytes(int B, int n) { return (size_t)B * ((n + kTile - 1) / kTile) * kTcBlockBytes; }

// ------------------------------------------------------------------------------------------------
// operand tiles.  One thread per point; block (blk, b).
// A-role plane p of a cloud holds [c_x, c_y, c_z, w] per row with (piece, w) = (h,N0) (h,N1) (m,N2) (m,1) (h,1) (l,1);
// B-role: -2 x pieces (h,m,h,m,l,h) with w = 1,1,1,N0,N1,N2, so that sum_k A_ik B_jk = N_i + N_j - 2 (hh'+hm'+mh'+mm'+hl'+lh').
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_trunc(float f) { return __uint_as_float(__float_as_uint(f) & 0xFFFFE000u); }
__device__ __forceinline__ void split3(double v, float& h, float& m, float& l) {
  h = tf32_trunc((float)v);
  const double r1 = v - (double)h;
  m = tf32_trunc((float)r1);
  const double r2 = r1 - (double)m;
  l = tf32_trunc((float)r2);
}

__global__ void __launch_bounds__(128) tc_prep_kernel(Batch bt) {
  const int b = blockIdx.y, blk = blockIdx.x, r = threadIdx.x;
  const GraphConsts* gc = bt.gc + b;
  if (!gc->use_tc) return;
  const int n = bt.n, nt = (n + kTile - 1) / kTile;
  const int j = blk * kTile + r;
  float4* out = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bt.opnd) + ((size_t)b * nt + blk) * kTcBlockBytes);
  const double* src = bt.src + (size_t)b * n * 3;
  const double* dst = bt.dst + (size_t)b * n * 3;
#pragma unroll
  for (int cloud = 0; cloud < 2; ++cloud) {
    float c[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    double nrm = 0;
    if (j < n) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double v = cloud == 0 ? (src[3 * j + k] - gc->cs[k]) * gc->s_hat : (dst[3 * j + k] - gc->cd[k]);
        split3(v, c[k][0], c[k][1], c[k][2]);
        const double rep = ((double)c[k][0] + (double)c[k][1]) + (double)c[k][2];
        nrm += rep * rep;
      }
    }
    float N0, N1, N2;
    split3(nrm, N0, N1, N2);
    float4* A = out + cloud * (kTcCloudBytes / 16);
    float4* Bq = out + (kTcRoleBytes / 16) + cloud * (kTcCloudBytes / 16);
    const float one = j < n ? 1.f : 0.f;  // rows past n: all-zero operands (finite results, masked later)
    A[0 * 128 + r] = make_float4(c[0][0], c[1][0], c[2][0], N0);
    A[1 * 128 + r] = make_float4(c[0][0], c[1][0], c[2][0], N1);
    A[2 * 128 + r] = make_float4(c[0][1], c[1][1], c[2][1], N2);
    A[3 * 128 + r] = make_float4(c[0][1], c[1][1], c[2][1], one);
    A[4 * 128 + r] = make_float4(c[0][0], c[1][0], c[2][0], one);
    A[5 * 128 + r] = make_float4(c[0][2], c[1][2], c[2][2], one);
    Bq[0 * 128 + r] = make_float4(-2.f * c[0][0], -2.f * c[1][0], -2.f * c[2][0], one);
    Bq[1 * 128 + r] = make_float4(-2.f * c[0][1], -2.f * c[1][1], -2.f * c[2][1], one);
    Bq[2 * 128 + r] = make_float4(-2.f * c[0][0], -2.f * c[1][0], -2.f * c[2][0], one);
    Bq[3 * 128 + r] = make_float4(-2.f * c[0][1], -2.f * c[1][1], -2.f * c[2][1], N0);
    Bq[4 * 128 + r] = make_float4(-2.f * c[0][2], -2.f * c[1][2], -2.f * c[2][2], N1);
    Bq[5 * 128 + r] = make_float4(-2.f * c[0][0], -2.f * c[1][0], -2.f * c[2][0], N2);
  }
}

// ------------------------------------------------------------------------------------------------
// tile schedule: a work item is a strip = (problem b, row block I, column blocks J0 .. J0+S-1 of the upper triangle).
// CTA c walks items c, c + gridDim.x, ...; the producer lane and the epilogue warps iterate the same sequence.
// ------------------------------------------------------------------------------------------------
struct TileIter {
  int item, step, total, spp, S, nt;
  int b, I, J, J1;
  bool first;
  __device__ __forceinline__ void init(int start, int step_, int total_, int spp_, int S_, int nt_) {
    item = start - step_;
    step = step_;
    total = total_;
    spp = spp_;
    S = S_;
    nt = nt_;
    b = I = 0;
    J = J1 = 0;
    first = false;
  }
  __device__ __forceinline__ bool next() {
    if (J + 1 < J1) {
      ++J;
      first = false;
      return true;
    }
    item += step;
    if (item >= total) return false;
    b = item / spp;
    int p = item - b * spp;
    I = 0;
    while (true) {
      const int ng = (nt - I + S - 1) / S;
      if (p < ng) break;
      p -= ng;
      ++I;
    }
    J = I + p * S;
    J1 = min(nt, J + S);
    first = true;
    return true;
  }
  __device__ __forceinline__ bool last_of_strip() const { return J + 1 == J1; }
};

__host__ __device__ inline int tc_strips_per_problem(int n, int S) {
  const int nt = (n + kTile - 1) / kTile;
  int total = 0;
  for (int I = 0; I < nt; ++I) total += (nt - I + S - 1) / S;
  return total;
}

// packed FP32x2 (two pairs per instruction; FADD2 / FMUL2 / FFMA2)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pk2f(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float sqrt_approx_tc(float x) {  // MUFU.SQRT, max relative error 2^-23 (part of theta)
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// d and p = a*b for two pairs; shared by the sweep and the (rare) revisit so both see identical values
__device__ __forceinline__ void tc_pair2(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, f32x2 two, f32x2 nbeta2,
                                         float& d0, float& d1, float& p0, float& p1) {
  const f32x2 A = pk2(a0, a1), B = pk2(b0, b1);
  const f32x2 t = sub2(A, B), s = add2(A, B), p = mul2(A, B);
  upk2(p, p0, p1);
  const f32x2 q = pk2f(sqrt_approx_tc(p0), sqrt_approx_tc(p1));
  const f32x2 w = fma2(q, two, s), t2 = mul2(t, t), d = fma2(w, nbeta2, t2);
  upk2(d, d0, d1);
}

// 32x32 bit transpose across the lanes of a warp (lane l passes row l, receives column l)
__device__ __forceinline__ uint32_t tc_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int j = 16 >> st;
    const uint32_t m = st == 0 ? 0x0000FFFFu : st == 1 ? 0x00FF00FFu : st == 2 ? 0x0F0F0F0Fu : st == 3 ? 0x33333333u
                                                                                                       : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    if ((lane & j) == 0)
      x ^= (((x >> j) ^ y) & m) << j;
    else
      x ^= ((y >> j) ^ x) & m;
  }
  return x;
}

// barrier slots in shared memory
enum { kBarAFull0 = 0, kBarAFull1, kBarBFull0, kBarBFull1, kBarBEmpty0, kBarBEmpty1, kBarTmemFull, kBarTmemEmpty, kNumBars };

template <bool kVerify>
__global__ void __launch_bounds__(kTcThreads, 2) graph_tc_kernel(Batch bt, int S, int spp, int total_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kTcRoleBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4 * kTcRoleBytes + 8 * kNumBars);
  const uint32_t sA0 = smem_u32(smem), sB0 = smem_u32(smem + 2 * kTcRoleBytes);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = bt.n, nt = (n + kTile - 1) / kTile;
  if (tid == 0) {
    for (int i = 0; i < kNumBars; ++i) mbar_init(bar(i), i == kBarTmemEmpty ? kTcEpiWarps : 1);
    mbar_fence_init();
  }
  if (warp == kTcEpiWarps) tmem_alloc<256>(smem_u32(tmem_slot));
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == kTcEpiWarps) {
    // ================= producer: TMA loads one tile ahead, MMA issue =================
    if (lane == 0) {
      const uint8_t* opnd = reinterpret_cast<const uint8_t*>(bt.opnd);
      TileIter ld, mm;
      ld.init(blockIdx.x, gridDim.x, total_items, spp, S, nt);
      mm.init(blockIdx.x, gridDim.x, total_items, spp, S, nt);
      uint32_t n_loaded = 0, n_strips_loaded = 0, n_mma = 0, n_strips = 0, ap = 0;
      auto issue_load = [&](const TileIter& t) {
        const uint32_t st = n_loaded & 1u, use = n_loaded >> 1;
        if (use > 0) mbar_wait(bar(kBarBEmpty0 + st), (use - 1) & 1u);  // MMAs that read this stage are complete
        if (t.first) {
          const uint32_t a = n_strips_loaded & 1u;
          mbar_arrive_expect_tx(bar(kBarAFull0 + a), kTcRoleBytes);
          bulk_g2s(sA0 + a * kTcRoleBytes, opnd + ((size_t)t.b * nt + t.I) * kTcBlockBytes, kTcRoleBytes, bar(kBarAFull0 + a));
          ++n_strips_loaded;
        }
        mbar_arrive_expect_tx(bar(kBarBFull0 + st), kTcRoleBytes);
        bulk_g2s(sB0 + st * kTcRoleBytes, opnd + ((size_t)t.b * nt + t.J) * kTcBlockBytes + kTcRoleBytes, kTcRoleBytes,
                 bar(kBarBFull0 + st));
        ++n_loaded;
      };
      auto next_tc = [&](TileIter& t) {  // next tile of a problem that takes the tensor-core path
        while (t.next())
          if (bt.gc[t.b].use_tc) return true;
        return false;
      };
      const uint32_t idesc = make_idesc_tf32(128, 128);
      bool have_ld = next_tc(ld);
      if (have_ld) issue_load(ld);
      while (next_tc(mm)) {
        have_ld = have_ld && next_tc(ld);
        if (have_ld) issue_load(ld);
        const uint32_t st = n_mma & 1u, use = n_mma >> 1;
        if (mm.first) {
          ap = n_strips & 1u;
          mbar_wait(bar(kBarAFull0 + ap), (n_strips >> 1) & 1u);
          ++n_strips;
        }
        mbar_wait(bar(kBarBFull0 + st), use & 1u);
        if (n_mma > 0) mbar_wait(bar(kBarTmemEmpty), (n_mma - 1) & 1u);  // the epilogue has drained the previous tile
        fence_after_sync();
#pragma unroll
        for (int cloud = 0; cloud < 2; ++cloud)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint64_t da = make_smem_desc(sA0 + ap * kTcRoleBytes + cloud * kTcCloudBytes + s * 2 * kTcPlaneBytes,
                                               kTcPlaneBytes, 128);
            const uint64_t db = make_smem_desc(sB0 + st * kTcRoleBytes + cloud * kTcCloudBytes + s * 2 * kTcPlaneBytes,
                                               kTcPlaneBytes, 128);
            mma_tf32(tbase + 128u * cloud, da, db, idesc, s > 0);
          }
        mma_commit(bar(kBarBEmpty0 + st));
        mma_commit(bar(kBarTmemFull));
        ++n_mma;
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    TileIter ti;
    ti.init(blockIdx.x, gridDim.x, total_items, spp, S, nt);
    uint32_t n_t = 0;
    int rdeg = 0;
    const int P32 = pitch32(n);
    const f32x2 two = pk2f(2.f, 2.f);
    while (ti.next()) {
      const GraphConsts* gcp = bt.gc + ti.b;
      if (!gcp->use_tc) continue;
      const int b = ti.b, I = ti.I, J = ti.J;
      const float beta2 = gcp->tc_beta2, theta = gcp->tc_theta, prisk = gcp->tc_prisk;
      const f32x2 nbeta2 = pk2f(-beta2, -beta2);
      const int i = I * kTile + 32 * q + lane;
      uint32_t* adj32 = reinterpret_cast<uint32_t*>(bt.adj) + (size_t)b * n * P32;
      int* degp = bt.deg + (size_t)b * n;
      mbar_wait(bar(kBarTmemFull), n_t & 1u);
      fence_after_sync();
      uint32_t word0 = 0u, word1 = 0u;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = 64 * h + 32 * c;
        const int j0 = J * kTile + col0;
        uint32_t ra[32], rb[32];
        tmem_ld32(tbase + lane_base + (uint32_t)col0, ra);
        tmem_ld32(tbase + lane_base + 128u + (uint32_t)col0, rb);
        tmem_wait_ld();
        if (c == 1) {  // both chunks of this warp are in registers: hand the accumulator back to the MMA issuer
          fence_before_sync();
          if (lane == 0) mbar_arrive(bar(kBarTmemEmpty));
        }
        // ---- sweep: bit k of word = sign(d_k), i.e. pair (i, j0+k) classified as an edge
        uint32_t word = 0u;
        float m1 = __int_as_float(0x7f800000), m2 = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 30; k >= 0; k -= 2) {
          float d0, d1, p0, p1;
          tc_pair2(ra[k], ra[k + 1], rb[k], rb[k + 1], two, nbeta2, d0, d1, p0, p1);
          word = __funnelshift_l(__float_as_uint(d1), word, 1);
          word = __funnelshift_l(__float_as_uint(d0), word, 1);
          m1 = fminf(m1, fminf(fabsf(d0), fabsf(d1)));
          m2 = fminf(m2, fminf(p0, p1));
        }
        // validity of the pairs of this thread: columns < n, row < n, i != j
        uint32_t vmask = j0 + 32 <= n ? 0xffffffffu : (j0 >= n ? 0u : ((1u << (n - j0)) - 1u));
        if (i >= n) vmask = 0u;
        if (i >= j0 && i < j0 + 32) vmask &= ~(1u << (i - j0));
        const bool flagged = !(m1 > theta) || !(m2 > prisk);
        if (kVerify || __any_sync(0xffffffffu, flagged && vmask != 0u)) {
          // ---- rare: find the undecided pairs of this thread and re-evaluate them with the exact FP64 sequence
          uint32_t fmask = 0u;
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            float d0, d1, p0, p1;
            tc_pair2(ra[k], ra[k + 1], rb[k], rb[k + 1], two, nbeta2, d0, d1, p0, p1);
            if (!(fabsf(d0) > theta) || !(p0 > prisk)) fmask |= 1u << k;
            if (!(fabsf(d1) > theta) || !(p1 > prisk)) fmask |= 2u << k;
          }
          fmask &= vmask;
          const double* src = bt.src + (size_t)b * n * 3;
          const double* dst = bt.dst + (size_t)b * n * 3;
          const double beta = gcp->beta;
          const bool scale_mode = bt.scale_mode != 0;
          const double s_hat = scale_mode ? bt.sol[b].scale : 1.0;
          if (kVerify) {  // every DECIDED pair is re-evaluated exactly; disagreements are counted (must stay 0)
            uint32_t vm = vmask & ~fmask;
            int bad = 0;
            while (vm) {
              const int k = __ffs(vm) - 1;
              vm &= vm - 1;
              const bool ex = scale_mode ? edge_exact_scale(src, dst, i, j0 + k, beta, s_hat) : edge_exact(src, dst, i, j0 + k, beta);
              bad += (ex != (((word >> k) & 1u) != 0u));
            }
            if (bad) atomicAdd(bt.mismatches, (unsigned long long)bad);
          }
          int nre = 0;
          while (fmask) {
            const int k = __ffs(fmask) - 1;
            fmask &= fmask - 1;
            const bool ex = scale_mode ? edge_exact_scale(src, dst, i, j0 + k, beta, s_hat) : edge_exact(src, dst, i, j0 + k, beta);
            word = (word & ~(1u << k)) | ((ex ? 1u : 0u) << k);
            ++nre;
          }
          if (bt.rechecks) {
            nre = __reduce_add_sync(0xffffffffu, nre);
            if (lane == 0 && nre) atomicAdd(bt.rechecks, (unsigned long long)nre);
          }
        }
        word &= vmask;
        if (c == 0)
          word0 = word;
        else
          word1 = word;
        rdeg += __popc(word);
        if (I != J) {  // transposed half: lane l holds column j0+l over rows I*128 + 32q .. +31
          const uint32_t colw = tc_transpose32(word, lane);
          const int jc = j0 + lane;
          if (jc < n) {
            adj32[(size_t)jc * P32 + 4 * I + q] = colw;
            if (colw) atomicAdd(degp + jc, __popc(colw));
          }
        }
      }
      if (i < n) *reinterpret_cast<uint2*>(adj32 + (size_t)i * P32 + 4 * J + 2 * h) = make_uint2(word0, word1);
      if (ti.last_of_strip()) {
        if (i < n && rdeg) atomicAdd(degp + i, rdeg);
        rdeg = 0;
      }
      ++n_t;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kTcEpiWarps) tmem_dealloc<256>(tbase);
}

int launch_graph_tc(const Batch& bt, cudaStream_t st, int num_sms) {
  static bool attr_done_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done_dev[dev & 63]) {
    cudaFuncSetAttribute(graph_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    cudaFuncSetAttribute(graph_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    attr_done_dev[dev & 63] = true;
  }
  const int nt = (bt.n + kTile - 1) / kTile;
  dim3 pg((unsigned)nt, (unsigned)bt.B);
  tc_prep_kernel<<<pg, 128, 0, st>>>(bt);
  // strip length: long strips amortise the A tile, short ones balance small batches
  const int ctas = 2 * num_sms;
  int S = 8;
  while (S > 1 && (long long)bt.B * tc_strips_per_problem(bt.n, S) < 4LL * ctas) S >>= 1;
  const int spp = tc_strips_per_problem(bt.n, S);
  const long long total = (long long)bt.B * spp;
  const int grid = (int)(total < ctas ? total : ctas);
  if (bt.flags_dbg & 2u)
    graph_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, st>>>(bt, S, spp, (int)total);
  else
    graph_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, st>>>(bt, S, spp, (int)total);
  return 2;
}

}  // namespace tzr
