// factor.cuh — exact GP inference for one cluster on the device:
//   Ky = K(x,x) + (noise + 1e-8) I           (k_build; GPy ExactGaussianInference + kern.K)
//   L  = chol(Ky)      (blocked right-looking, DMMA trailing updates)      [GPy: dpotrf]
//   M  = L^-1          (blocked, same launches; replaces GPy's dtrtri/dpotri + dtrtrs)
//   z  = M r, alpha = M^T z, LML = -(m ln 2pi + 2 sum ln L_ii + z.z)/2       [GPy: dpotrs]
//   gradient of LML w.r.t. theta, contracted tile by tile from W = M^T M    [GPy: dL_dK]
// Reference call sites: mogp/obsmodel.py:46,56,65,73,92,94.  Formula-level fidelity
// (SURVEY.md §2.3): r^2 by the expansion x^2 + x'^2 - 2xx' with clip, sqrt then square,
// the 1e-8 jitter.
#pragma once
#include "tile.cuh"

namespace mogp {

constexpr double GPY_JITTER = 1e-8;
constexpr double LOG_2_PI = 1.8378770664093454835606594728112;

struct Theta {
  double A, k0, k1, noise;
  int kernel;     // 0 rbf, 1 linear+bias
  int mean_func;  // NegLinear on/off
};

// r^2 exactly as GPy's Stationary._unscaled_dist evaluates it (no FMA contraction).
__device__ __forceinline__ double r2_expansion(double xi, double xj) {
  double r2 = __dadd_rn(__dmul_rn(-2.0, __dmul_rn(xi, xj)), __dadd_rn(__dmul_rn(xi, xi), __dmul_rn(xj, xj)));
  return r2 > 0.0 ? r2 : 0.0;
}

// k(xi, xj); `same` marks a diagonal element of the symmetric matrix (GPy forces r = 0 there).
// scaled_r2 (optional) returns (dist / lengthscale)^2 for the lengthscale gradient.
__device__ __forceinline__ double kern_eval(const Theta& th, double xi, double xj, bool same, double* scaled_r2) {
  if (th.kernel == 0) {
    double r = same ? 0.0 : sqrt(r2_expansion(xi, xj)) / th.k1;
    double rr = __dmul_rn(r, r);
    if (scaled_r2) *scaled_r2 = rr;
    return th.k0 * exp(-0.5 * rr);
  }
  double xx = __dmul_rn(xi, xj);
  if (scaled_r2) *scaled_r2 = xx;
  return __dadd_rn(__dmul_rn(th.k0, xx), th.k1);
}

__device__ __forceinline__ double kern_diag(const Theta& th, double x) {
  return th.kernel == 0 ? th.k0 : __dadd_rn(__dmul_rn(th.k0, __dmul_rn(x, x)), th.k1);
}

// ---------------------------------------------------------------------------------------------
// Ky (lower) into L, zero accumulator into M.  Rows >= m are identity padding.
// grid (nb, nb): x = block column, y = block row; CTAs above the diagonal exit.
// (theta is read from device memory - the cluster's stat block - so that a captured CUDA graph of the whole
// evaluation can be relaunched at new hyper-parameters without touching its nodes)
// wt (optional): multiplicity of each point.  n identical observations (same input, here also the same residual mean) are
// ONE pseudo-observation with noise (noise + 1e-8) / n (plus closed-form terms in the log marginal, engine.cu): the refit's
// evaluations run on the cluster with its repeated inputs collapsed - every patient brings the same onset anchor, ~13 % of a
// cluster's rows - and on nothing else.
__global__ void __launch_bounds__(NTHREADS) k_build(double* __restrict__ L, double* __restrict__ M, int64_t ld,
                                                    const double* __restrict__ x, int m, const Theta* __restrict__ thp,
                                                    double jitter, const double* __restrict__ wt) {
  const int bj = blockIdx.x, bi = blockIdx.y;
  if (bj > bi) return;
  const Theta th = *thp;
  __shared__ double sxi[TB], sxj[TB];
  if (threadIdx.x < TB) {
    int i = bi * TB + threadIdx.x, j = bj * TB + threadIdx.x;
    sxi[threadIdx.x] = i < m ? x[i] : 0.0;
    sxj[threadIdx.x] = j < m ? x[j] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TB * TB; idx += NTHREADS) {
    int a = idx >> 6, b = idx & 63;
    int i = bi * TB + a, j = bj * TB + b;
    double v = 0.0;
    if (j <= i) {
      if (i < m) {
        v = kern_eval(th, sxi[a], sxj[b], i == j, nullptr);
        if (i == j) v = (wt ? v + __dadd_rn(th.noise, GPY_JITTER) / wt[i] : __dadd_rn(v, __dadd_rn(th.noise, GPY_JITTER))) + jitter;
      } else {
        v = (i == j) ? 1.0 : 0.0;
      }
    }
    L[(int64_t)i * ld + j] = v;
    M[(int64_t)i * ld + j] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Panel step p in two launches:
//   k_diag  (1 CTA)     : L_pp = chol(A_pp) (kept in shared memory only), M_pp = L_pp^-1 -> M
//   k_panel (nb-1 CTAs) : b > p : L_bp = A_bp M_pp^T          (TRSM by the inverse of the diagonal tile)
//                         b < p : M_pb = -M_pp Acc_pb         (finalise block-row p of the inverse factor)
// The serial 64x64 factorisation runs once per panel on one SM instead of once per CTA: with several clusters
// refitted concurrently (one stream each) that SM-time is what bounds the throughput.
// The diagonal tile of L keeps A_pp (L_pp is never needed again: log-determinants come from the diagonal of M).
constexpr int DIAG_SMEM = (TB * LD_C + TB * LD_C + 9 * TB) * (int)sizeof(double);
constexpr int PANEL_SMEM = (TB * LD_C + TB * LD_R) * (int)sizeof(double);

__global__ void __launch_bounds__(NTHREADS) k_diag(const double* __restrict__ L, double* __restrict__ M, int64_t ld, int p,
                                                   int* __restrict__ info) {
  extern __shared__ __align__(16) double smem[];
  double* sD = smem;                 // [64][LD_C]  A_pp -> L_pp
  double* sM = sD + TB * LD_C;       // [64][LD_C]  M_pp
  double* col = sM + TB * LD_C;      // publish buffers
  const int64_t dpp = (int64_t)p * TB * ld + (int64_t)p * TB;
  const long long t0 = clock64();
  load_tile<LD_C>(sD, L + dpp, ld);
  double* rdiag = col + 8 * TB;      // [64]   (col: [2][4][64] for trtri_tile, [2][64] for potrf_tile)
  int fail = potrf_tile<LD_C>(sD, col, rdiag);
  const long long t1 = clock64();
  if (fail && threadIdx.x == 0) atomicCAS(info, 0, p * TB + fail);
  trtri_tile<LD_C>(sD, sM, col, rdiag);
  const long long t2 = clock64();
  if (threadIdx.x == 0 && p == 0) {  // diagnostics: cycles of the two serial phases (stat[14], stat[15] via info + 12)
    reinterpret_cast<double*>(info)[6] = (double)(t1 - t0);
    reinterpret_cast<double*>(info)[7] = (double)(t2 - t1);
  }
  for (int idx = threadIdx.x; idx < TB * TB; idx += NTHREADS) {
    int i = idx >> 6, j = idx & 63;
    M[dpp + (int64_t)i * ld + j] = sM[i * LD_C + j];
  }
}

// grid nb - 1: blockIdx.x skips b == p.
__global__ void __launch_bounds__(NTHREADS) k_panel(double* __restrict__ L, double* __restrict__ M, int64_t ld, int p) {
  extern __shared__ __align__(16) double smem[];
  double* sM = smem;                 // [64][LD_C]  M_pp
  double* sT = sM + TB * LD_C;       // [64][LD_R]  the CTA's own tile
  const int b = (int)blockIdx.x < p ? (int)blockIdx.x : (int)blockIdx.x + 1;
  const int64_t dpp = (int64_t)p * TB * ld + (int64_t)p * TB;
  load_tile<LD_C>(sM, M + dpp, ld);
  if (b > p) load_tile<LD_C>(sT, L + (int64_t)b * TB * ld + (int64_t)p * TB, ld);
  else load_tile<LD_R>(sT, M + (int64_t)p * TB * ld + (int64_t)b * TB, ld);
  __syncthreads();
  Acc64 acc;
  acc.zero();
  if (b > p) {
    warp_mma<2, 4, true, true>(acc, sT, sM, warp_rb64(), warp_cb64());  // A_bp * M_pp^T
    double* out = L + (int64_t)b * TB * ld + (int64_t)p * TB;
    for_each_acc64(acc, [&](int r, int c, double& v) { out[(int64_t)r * ld + c] = v; });
  } else {
    warp_mma<2, 4, true, false>(acc, sM, sT, warp_rb64(), warp_cb64());  // M_pp * Acc_pb
    double* out = M + (int64_t)p * TB * ld + (int64_t)b * TB;
    for_each_acc64(acc, [&](int r, int c, double& v) { out[(int64_t)r * ld + c] = -v; });
  }
}

// ---------------------------------------------------------------------------------------------
// After panel p: grid (nb, nb), x = block column j, y = block row i (> p):
//   j > p, j <= i : A_ij -= L_ip L_jp^T                (Cholesky trailing update)
//   j <= p        : Acc_ij += L_ip M_pj                (right-looking update of the inverse factor)
constexpr int TRAIL_SMEM = (TB * LD_C + TB * LD_R) * (int)sizeof(double);

// Two-level blocking: with one trailing update per 64-column panel every output tile is read and written once per
// panel and the evaluation is bound by HBM (measured: ~2.9 GB of traffic per evaluation at m = 2 600, 3.7 TB/s in
// aggregate over the concurrent refits).  Panels are therefore grouped FACTOR_KB at a time: after panel p only the
// tiles the rest of the group needs are updated (columns < P1 of the Cholesky part, rows < P1 of the inverse part);
// everything below / to the right gets ONE update per group with the contraction over all its panels (k_trail_blk).
constexpr int FACTOR_KB = 4;
__global__ void __launch_bounds__(NTHREADS) k_trail(double* __restrict__ L, double* __restrict__ M, int64_t ld, int p, int P1) {
  const int bj = blockIdx.x, bi = blockIdx.y + p + 1;
  if (bj > bi) return;
  if (bj > p ? bj >= P1 : bi >= P1) return;  // deferred to the group's update
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;             // L_ip   [64][LD_C]
  double* sB = sA + TB * LD_C;   // L_jp [64][LD_C]  or  M_pj [64][LD_R]
  load_tile<LD_C>(sA, L + (int64_t)bi * TB * ld + (int64_t)p * TB, ld);
  Acc64 acc;
  acc.zero();
  if (bj > p) {
    load_tile<LD_C>(sB, L + (int64_t)bj * TB * ld + (int64_t)p * TB, ld);
    __syncthreads();
    warp_mma<2, 4, true, true>(acc, sA, sB, warp_rb64(), warp_cb64());
    double* out = L + (int64_t)bi * TB * ld + (int64_t)bj * TB;
    const bool diag = (bi == bj);
    for_each_acc64(acc, [&](int r, int c, double& v) {
      if (!diag || c <= r) out[(int64_t)r * ld + c] -= v;
    });
  } else {
    load_tile<LD_R>(sB, M + (int64_t)p * TB * ld + (int64_t)bj * TB, ld);
    __syncthreads();
    warp_mma<2, 4, true, false>(acc, sA, sB, warp_rb64(), warp_cb64());
    double* out = M + (int64_t)bi * TB * ld + (int64_t)bj * TB;
    for_each_acc64(acc, [&](int r, int c, double& v) { out[(int64_t)r * ld + c] += v; });
  }
}

// After the panels [P0, P1) of a group: rows bi >= P1, two row tiles per CTA (they share the second operand):
//   bj >= P1 : A_ij -= sum_p L_ip L_jp^T          bj < P1 : Acc_ij += sum_{p >= max(P0, bj)} L_ip M_pj
// The contraction over the group's panels runs in 32-deep slices through a two-stage cp.async ring: the next slice is
// in flight while the current one is multiplied (ncu on the synchronous version: 30 % of the warp stalls on the global
// loads of the operand tiles, tensor pipe 51 % of the elapsed cycles), and the two stages of 2-row-tile CTAs still fit
// twice on an SM.
constexpr int TSL = 32;              // contraction slice
constexpr int LD_S = 36;             // smem ld of a K-contiguous slice [64][32] (4 mod 16)
constexpr int TSL_TILE = TB * LD_S;  // doubles per operand slice (>= 32 * LD_R for a [32][64] slice of M)
static_assert(TSL * LD_R <= TSL_TILE, "slice of M does not fit");
constexpr int TRAILB_SMEM = 2 * 3 * TSL_TILE * (int)sizeof(double);
template <int ROWS>  // row tiles per CTA: 2 (shared second operand) or 1
__global__ void __launch_bounds__(NTHREADS) k_trail_blk(double* __restrict__ L, double* __restrict__ M, int64_t ld, int P0, int P1,
                                                        int nb) {
  const int bj = blockIdx.x, bi0 = P1 + ROWS * (int)blockIdx.y, bi1 = bi0 + 1;
  const bool do0 = bj <= bi0, do1 = ROWS == 2 && bi1 < nb && bj <= bi1;
  if (!do0 && !do1) return;
  extern __shared__ __align__(16) double smem[];
  const bool lpart = bj >= P1;
  const int p_first = lpart ? P0 : max(P0, bj);
  const int nslice = (P1 - p_first) * (TB / TSL);
  auto issue = [&](int sl, int stage) {
    double* sA0 = smem + stage * (3 * TSL_TILE);
    double* sA1 = sA0 + TSL_TILE;
    double* sB = sA1 + TSL_TILE;
    const int p = p_first + sl / (TB / TSL);
    const int64_t k0 = (int64_t)p * TB + (sl % (TB / TSL)) * TSL;
#pragma unroll
    for (int q = 0; q < TB * TSL / 2 / NTHREADS; ++q) {  // [64][32] slices, 16 chunks of 16 bytes per row
      const int idx = threadIdx.x + q * NTHREADS, r = idx >> 4, c = (idx & 15) << 1;
      if (do0) cp_async16(sA0 + r * LD_S + c, L + ((int64_t)bi0 * TB + r) * ld + k0 + c);
      if (do1) cp_async16(sA1 + r * LD_S + c, L + ((int64_t)bi1 * TB + r) * ld + k0 + c);
      if (lpart) cp_async16(sB + r * LD_S + c, L + ((int64_t)bj * TB + r) * ld + k0 + c);
    }
    if (!lpart) {
#pragma unroll
      for (int q = 0; q < TSL * TB / 2 / NTHREADS; ++q) {  // [32][64] slice of M: 32 chunks per row
        const int idx = threadIdx.x + q * NTHREADS, k = idx >> 5, c = (idx & 31) << 1;
        cp_async16(sB + k * LD_R + c, M + (k0 + k) * ld + (int64_t)bj * TB + c);
      }
    }
  };
  Acc64 acc0, acc1;
  acc0.zero();
  acc1.zero();
  if (nslice > 0) issue(0, 0);
  cp_async_commit();
  for (int sl = 0; sl < nslice; ++sl) {
    cp_async_wait<0>();
    __syncthreads();  // slice sl has landed for everybody; the other stage is free
    if (sl + 1 < nslice) issue(sl + 1, (sl + 1) & 1);
    cp_async_commit();
    const double* sA0 = smem + (sl & 1) * (3 * TSL_TILE);
    const double* sA1 = sA0 + TSL_TILE;
    const double* sB = sA1 + TSL_TILE;
    if (lpart) {
      if (do0) warp_mma<2, 4, true, true, TSL, LD_S, LD_S>(acc0, sA0, sB, warp_rb64(), warp_cb64());
      if (do1) warp_mma<2, 4, true, true, TSL, LD_S, LD_S>(acc1, sA1, sB, warp_rb64(), warp_cb64());
    } else {
      if (do0) warp_mma<2, 4, true, false, TSL, LD_S, LD_R>(acc0, sA0, sB, warp_rb64(), warp_cb64());
      if (do1) warp_mma<2, 4, true, false, TSL, LD_S, LD_R>(acc1, sA1, sB, warp_rb64(), warp_cb64());
    }
  }
  cp_async_wait<0>();
  if (lpart) {
    if (do0) {
      double* out = L + (int64_t)bi0 * TB * ld + (int64_t)bj * TB;
      const bool diag = bi0 == bj;
      for_each_acc64(acc0, [&](int r, int c, double& v) {
        if (!diag || c <= r) out[(int64_t)r * ld + c] -= v;
      });
    }
    if (do1) {
      double* out = L + (int64_t)bi1 * TB * ld + (int64_t)bj * TB;
      const bool diag = bi1 == bj;
      for_each_acc64(acc1, [&](int r, int c, double& v) {
        if (!diag || c <= r) out[(int64_t)r * ld + c] -= v;
      });
    }
  } else {
    if (do0) {
      double* out = M + (int64_t)bi0 * TB * ld + (int64_t)bj * TB;
      for_each_acc64(acc0, [&](int r, int c, double& v) { out[(int64_t)r * ld + c] += v; });
    }
    if (do1) {
      double* out = M + (int64_t)bi1 * TB * ld + (int64_t)bj * TB;
      for_each_acc64(acc1, [&](int r, int c, double& v) { out[(int64_t)r * ld + c] += v; });
    }
  }
}

// The same group update with 128 x 128 output blocks (two row tiles x two column tiles per CTA): every operand slice
// loaded serves four tile products instead of two (the update streams L and M from HBM when ~30 clusters are refitted at
// once: arithmetic intensity is what it is short of), a warp owns 32 x 64 of the block (32 DMMAs per 12 fragment
// loads), and the contraction runs through a two-stage cp.async ring in 32-deep slices - the next slice is in flight
// while the current one is multiplied.  Fragments of the K-contiguous operands (rows of L) come from 128-bit loads (lane
// t supplies k = 2t, 2t+1 to two consecutive DMMAs, as in the scoring kernel); the rows of M of the inverse part are
// stored [k][column] and read with the matching k.
constexpr int TBIG_BK = 32;                       // contraction slice
constexpr int TBIG_LDA = 40;                      // smem ld of a K-contiguous slice [128][32] (8 mod 16)
constexpr int TBIG_LDM = 130;                     // smem ld of a slice of M [32][128] (2 mod 16: k = 2t, 2t+1 conflict free)
constexpr int TBIG_STAGE = 2 * 128 * TBIG_LDA;    // doubles per stage: A slice | B slice
constexpr int TBIG_SMEM = 2 * TBIG_STAGE * (int)sizeof(double);
__global__ void __launch_bounds__(NTHREADS, 1) k_trail_big(double* __restrict__ L, double* __restrict__ M, int64_t ld, int P0, int P1,
                                                           int nb) {
  const int bj0 = 2 * (int)blockIdx.x, bi0 = P1 + 2 * (int)blockIdx.y;  // first column / row tile of the block
  if (bi0 >= nb || bj0 > bi0 + 1 || bj0 >= nb) return;
  const bool lpart = bj0 >= P1;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int rb = 32 * (warp >> 1), cb = 64 * (warp & 1);
  const int rows_ok = min(128, (nb - bi0) * TB), cols_ok = min(128, (nb - bj0) * TB);  // the block may overhang the matrix
  const int p_first = lpart ? P0 : max(P0, bj0);
  const int nslice = (P1 - p_first) * (TB / TBIG_BK);
  auto issue = [&](int sl, int stage) {
    double* sA = smem + stage * TBIG_STAGE;
    double* sB = sA + 128 * TBIG_LDA;
    const int p = p_first + sl / (TB / TBIG_BK), k0 = p * TB + (sl % (TB / TBIG_BK)) * TBIG_BK;  // panel, first contraction index
#pragma unroll
    for (int q = 0; q < 128 * TBIG_BK / 2 / NTHREADS; ++q) {
      const int idx = tid + q * NTHREADS, r = idx >> 4, c = (idx & 15) << 1;
      double* d = sA + r * TBIG_LDA + c;
      if (r < rows_ok) cp_async16(d, L + (int64_t)(bi0 * TB + r) * ld + k0 + c);
      else *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
    }
    if (lpart) {
#pragma unroll
      for (int q = 0; q < 128 * TBIG_BK / 2 / NTHREADS; ++q) {
        const int idx = tid + q * NTHREADS, r = idx >> 4, c = (idx & 15) << 1;
        double* d = sB + r * TBIG_LDA + c;
        if (r < cols_ok) cp_async16(d, L + (int64_t)(bj0 * TB + r) * ld + k0 + c);
        else *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
      }
    } else {
#pragma unroll
      for (int q = 0; q < TBIG_BK * 128 / 2 / NTHREADS; ++q) {
        const int idx = tid + q * NTHREADS, k = idx >> 6, c = (idx & 63) << 1;
        double* d = sB + k * TBIG_LDM + c;
        // M_pj is lower triangular by tiles: nothing above the diagonal (never written: do not read it)
        const bool live = c < cols_ok && (bj0 + (c >> 6)) <= p;
        if (live) cp_async16(d, M + (int64_t)(k0 + k) * ld + bj0 * TB + c);
        else *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
      }
    }
  };
  Acc<4, 8> acc;
  acc.zero();
  if (nslice > 0) issue(0, 0);
  cp_async_commit();
  for (int sl = 0; sl < nslice; ++sl) {
    cp_async_wait<0>();
    __syncthreads();  // slice sl has landed for everybody; the other stage is free
    if (sl + 1 < nslice) issue(sl + 1, (sl + 1) & 1);
    cp_async_commit();
    const double* sA = smem + (sl & 1) * TBIG_STAGE;
    const double* sB = sA + 128 * TBIG_LDA;
#pragma unroll
    for (int kb = 0; kb < TBIG_BK; kb += 8) {
      double2 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double2*>(sA + (rb + 8 * i + g) * TBIG_LDA + kb + 2 * t);
      if (lpart) {
        double2 b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const double2*>(sB + (cb + 8 * j + g) * TBIG_LDA + kb + 2 * t);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].y, b[j].y);
      } else {
        double b0[8], b1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b0[j] = sB[(kb + 2 * t) * TBIG_LDM + cb + 8 * j + g];
          b1[j] = sB[(kb + 2 * t + 1) * TBIG_LDM + cb + 8 * j + g];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].x, b0[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].y, b1[j]);
      }
    }
  }
  cp_async_wait<0>();
  double* out = lpart ? L : M;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = rb + 8 * i + g, c = cb + 8 * j + 2 * t;
      if (r >= rows_ok || c >= cols_ok) continue;
      const int bi = bi0 + (r >> 6), bj = bj0 + (c >> 6);
      if (bj > bi) continue;  // a tile above the diagonal
      double2* o = reinterpret_cast<double2*>(out + (int64_t)(bi0 * TB + r) * ld + bj0 * TB + c);
      double2 v = *o;
      if (lpart) {
        const int gr = bi0 * TB + r, gc = bj0 * TB + c;  // inside a diagonal tile only the lower part is kept
        if (gc <= gr) v.x -= acc.v[i][j][0];
        if (gc + 1 <= gr) v.y -= acc.v[i][j][1];
      } else {
        v.x += acc.v[i][j][0];
        v.y += acc.v[i][j][1];
      }
      *o = v;
    }
}

// ---------------------------------------------------------------------------------------------
// z = M r with r = y + A x (residual of the NegLinear mean, neg_linear.py:22-23). Warp per row.
__global__ void __launch_bounds__(NTHREADS) k_z(const double* __restrict__ M, int64_t ld, const double* __restrict__ x,
                                                const double* __restrict__ y, int m, const Theta* __restrict__ thp,
                                                double* __restrict__ z) {
  const Theta th = *thp;
  const int i = blockIdx.x * (NTHREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= m) return;
  const double A = th.mean_func ? th.A : 0.0;
  double acc = 0.0;
  for (int j = lane; j <= i; j += 32) acc += M[(int64_t)i * ld + j] * (y[j] + A * x[j]);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) z[i] = acc;
}

// alpha = M^T z in two deterministic stages.  Stage 1: CTA (J, C) sums rows [256 C, 256 C + 256) of the 64-column
// strip J (each warp reads whole 512-byte row segments) into apart[C][col]; stage 2 (inside k_lml) adds the
// row chunks in order.
constexpr int ALPHA_ROWS = 256;
__global__ void __launch_bounds__(NTHREADS) k_alpha_part(const double* __restrict__ M, int64_t ld,
                                                         const double* __restrict__ z, int m, int mpad,
                                                         double* __restrict__ apart) {
  __shared__ double2 red[NTHREADS / 32][32];
  const int J = blockIdx.x, C = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = max(C * ALPHA_ROWS, J * TB), r1 = min((C + 1) * ALPHA_ROWS, m);
  double2 acc = make_double2(0.0, 0.0);
  for (int i = r0 + w; i < r1; i += NTHREADS / 32) {
    const double2 v = *reinterpret_cast<const double2*>(M + (int64_t)i * ld + J * TB + 2 * lane);
    const double zi = z[i];
    acc.x += v.x * zi;
    acc.y += v.y * zi;
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0) {
    double2 s = make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < NTHREADS / 32; ++q) {
      s.x += red[q][lane].x;
      s.y += red[q][lane].y;
    }
    *reinterpret_cast<double2*>(apart + (int64_t)C * mpad + J * TB + 2 * lane) = s;
  }
}

// alpha[i] = sum of the row-chunk partials (fixed order).
__global__ void __launch_bounds__(256) k_alpha_reduce(const double* __restrict__ apart, int nchunk, int mpad,
                                                      double* __restrict__ alpha, int m) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  double a = 0.0;
  for (int c = i / ALPHA_ROWS; c < nchunk; ++c) a += apart[(int64_t)c * mpad + i];
  alpha[i] = a;
}

// One CTA: alpha = sum of the row-chunk partials; quad = z.z, logdet = 2 sum ln L_ii = -2 sum ln M_ii,
// sxa = sum x_i alpha_i, LML.  out[0..3].
__global__ void __launch_bounds__(1024) k_lml(const double* __restrict__ M, int64_t ld, const double* __restrict__ z,
                                              const double* __restrict__ x, const double* __restrict__ apart, int nchunk,
                                              int mpad, double* __restrict__ alpha, int m, double* __restrict__ out) {
  __shared__ double red[3][1024];
  double q = 0.0, ld2 = 0.0, sxa = 0.0;
  for (int i = threadIdx.x; i < m; i += 1024) {
    double a = 0.0;
    for (int c = i / ALPHA_ROWS; c < nchunk; ++c) a += apart[(int64_t)c * mpad + i];  // chunks above row i hold zeros
    alpha[i] = a;
    q += z[i] * z[i];
    ld2 -= log(M[(int64_t)i * ld + i]);
    sxa += x[i] * a;
  }
  red[0][threadIdx.x] = q;
  red[1][threadIdx.x] = ld2;
  red[2][threadIdx.x] = sxa;
  __syncthreads();
  for (int s = 512; s; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
      red[2][threadIdx.x] += red[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double logdet = 2.0 * red[1][0];
    out[0] = 0.5 * (-(double)m * LOG_2_PI - logdet - red[0][0]);  // log marginal likelihood
    out[1] = red[0][0];                                           // r^T Ky^-1 r
    out[2] = logdet;
    out[3] = red[2][0];                                           // x . alpha  (dLML/dA = -x.alpha)
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient contraction.  W = M^T M is never materialised: CTA (bj, bi, kc) forms the partial tile
//   W_ij^(kc) = sum_{k in chunk kc, k >= bi} M_k,bi^T M_k,bj
// and contracts it (plus the alpha alpha^T term, on chunk 0 only) with dK/dtheta built on the fly:
//   S0 = sum dL_dK_ij * K_ij        (-> d/d variance [rbf], d/d bias [linear: K_ij := 1])
//   S1 = sum dL_dK_ij * K_ij * r2   (-> d/d lengthscale [rbf], d/d variances [linear: x_i x_j])
//   S2 = trace(dL_dK)               (-> d/d noise)
// with dL_dK = (alpha alpha^T - W)/2 (GPy ExactGaussianInference).  Partials are written per CTA
// and summed in a fixed order by k_grad_final, so the result is run-to-run deterministic.
constexpr int GRAD_KC = 64;  // contraction chunk (row tiles per CTA; 4 -> 64: 334 -> 245 us at m = 2600, every chunk CTA
                            // repeats the kernel-derivative epilogue of its tile pair); grad_kc() in engine.cu may override it
constexpr int GRAD_SMEM = (2 * TB * LD_R) * (int)sizeof(double);

__global__ void __launch_bounds__(NTHREADS) k_grad(const double* __restrict__ M, int64_t ld, int nb,
                                                   const double* __restrict__ x, const double* __restrict__ alpha,
                                                   int m, const Theta* __restrict__ thp, double* __restrict__ partial,
                                                   int kchunk, const double* __restrict__ wt) {
  const Theta th = *thp;
  const int bj = blockIdx.x, bi = blockIdx.y, kc = blockIdx.z;
  const int cta = (kc * gridDim.y + bi) * gridDim.x + bj;
  const int k0 = bi + kc * kchunk;
  __shared__ double red[3][NTHREADS / 32];
  if (bj > bi || k0 >= nb) {
    if (threadIdx.x < 3) partial[(int64_t)cta * 3 + threadIdx.x] = 0.0;
    return;
  }
  extern __shared__ __align__(16) double smem[];
  Acc64 acc;
  acc.zero();
  const int k1 = min(nb, k0 + kchunk);
  // the rows of the contraction stream through a two-stage cp.async ring in 32-row slices of the two column strips
  constexpr int GSL = 32, GST = 2 * GSL * LD_R;  // slice depth, doubles per stage
  const int nslice = (k1 - k0) * (TB / GSL);
  auto issue = [&](int sl, int stage) {
    double* sA = smem + stage * GST;
    double* sB = sA + GSL * LD_R;
    const int64_t r0 = (int64_t)(k0 + sl / (TB / GSL)) * TB + (sl % (TB / GSL)) * GSL;
#pragma unroll
    for (int q = 0; q < GSL * TB / 2 / NTHREADS; ++q) {
      const int idx = threadIdx.x + q * NTHREADS, k = idx >> 5, c = (idx & 31) << 1;
      cp_async16(sA + k * LD_R + c, M + (r0 + k) * ld + (int64_t)bi * TB + c);
      cp_async16(sB + k * LD_R + c, M + (r0 + k) * ld + (int64_t)bj * TB + c);
    }
  };
  issue(0, 0);
  cp_async_commit();
  for (int sl = 0; sl < nslice; ++sl) {
    cp_async_wait<0>();
    __syncthreads();  // slice sl has landed for everybody; the other stage is free
    if (sl + 1 < nslice) issue(sl + 1, (sl + 1) & 1);
    cp_async_commit();
    const double* sA = smem + (sl & 1) * GST;
    warp_mma<2, 4, false, false, GSL, LD_R, LD_R>(acc, sA, sA + GSL * LD_R, warp_rb64(), warp_cb64());
  }
  cp_async_wait<0>();
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for_each_acc64(acc, [&](int r, int c, double& w) {
    int i = bi * TB + r, j = bj * TB + c;
    if (i < m && j <= i) {
      double aa = (kc == 0) ? alpha[i] * alpha[j] : 0.0;
      double dl = 0.5 * (aa - w);
      double f = (i == j) ? 1.0 : 2.0;
      double r2;
      double kij = kern_eval(th, x[i], x[j], i == j, &r2);
      if (th.kernel == 1) kij = 1.0;
      s0 += f * dl * kij;
      s1 += f * dl * kij * r2;
      if (i == j) s2 += wt ? dl / wt[i] : dl;  // d Ky_ii / d noise = 1 / multiplicity
    }
  });
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
    red[2][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NTHREADS / 32; ++q) s += red[threadIdx.x][q];
    partial[(int64_t)cta * 3 + threadIdx.x] = s;
  }
}

// grad_theta[4] = dLML/d[A, k0, k1, noise] from the partial sums (one CTA, fixed order).
__global__ void __launch_bounds__(256) k_grad_final(const double* __restrict__ partial, int n_partial,
                                                    const Theta* __restrict__ thp, const double* __restrict__ lml_out,
                                                    double* __restrict__ grad) {
  const Theta th = *thp;
  __shared__ double red[3][256];
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n_partial; i += 256)
    for (int q = 0; q < 3; ++q) s[q] += partial[(int64_t)i * 3 + q];
  for (int q = 0; q < 3; ++q) red[q][threadIdx.x] = s[q];
  __syncthreads();
  for (int st = 128; st; st >>= 1) {
    if (threadIdx.x < st)
      for (int q = 0; q < 3; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    grad[0] = th.mean_func ? -lml_out[3] : 0.0;  // neg_linear.py:25-26 with dL_dm = alpha
    if (th.kernel == 0) {
      grad[1] = red[0][0] / th.k0;  // Stationary.update_gradients_full: sum(K * dL_dK) / variance
      grad[2] = red[1][0] / th.k1;  // -sum(dL_dr * r) / lengthscale, dK_dr = -r K
    } else {
      grad[1] = red[1][0];  // Linear: sum(dL_dK * x x^T)
      grad[2] = red[0][0];  // Bias:   sum(dL_dK)
    }
    grad[3] = red[2][0];  // Gaussian.exact_inference_gradients: trace(dL_dK)
  }
}

}  // namespace mogp
