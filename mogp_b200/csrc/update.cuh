// update.cuh — rank-n up/down-dates of a cluster's inverse Cholesky factor M at fixed theta, and the
// leave-block-out ("own cluster") score that avoids most down-dates altogether.
//
// The reference refactorises the whole cluster (O(m^3), GPy set_XY) every time a patient leaves
// (mogp/obsmodel.py:68-73, mogp_constrained.py:239-242) or joins (obsmodel.py:87-92).  With
// Ky = L L^T, M = L^-1, W = Ky^-1 = M^T M and a patient's n rows at offset p (block B):
//
//  * own-cluster score without removing the patient:  y_B | y_rest ~ N(y_B - C alpha_B, C), C = W_BB^-1,
//    W_BB = sum_k b_k b_k^T with b_k = M[k, p..p+n)  (k_rm_gram + k_own_final);
//  * row deletion: M' (rows/cols of B removed) = u_i (M[i,j] - g_i . h_{i-1}[:,j]) for i > B, where
//    G_i = sum_{k<=i} b_k b_k^T, g_i = G_{i-1}^-1 b_i, u_i = (1 + b_i.g_i)^-1/2, h_i[:,j] = sum_{k<=i} b_k M[k,j]
//    (the closed form of the Givens sweep that re-triangularises M after a column block is dropped);
//    k_rm_gram -> k_rm_prep -> k_rm_apply, tile by tile with DMMA for the in-tile prefix;
//  * append: M' = [[M, 0], [-Ls^-1 T^T M, Ls^-1]], T = M Kx (kept by the scoring pass),
//    Ls = chol(Kss + (noise+1e-8) I - T^T T);  k_append_small -> k_append_rows -> k_append_finalize.
#pragma once
#include "score.cuh"

namespace mogp {

constexpr int NP = 16;                  // max rows appended / removed / left out per operation
constexpr int NG = NP * (NP + 1) / 2;   // packed lower-triangular n x n
constexpr int LD_W = 84;                // smem ld of the [64][64+NP] operand (84 = 4 mod 16)

__device__ __forceinline__ void tri_unpack(int t, int& a, int& b) {  // t -> (a >= b)
  a = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= t) ++a;
  while (a * (a + 1) / 2 > t) --a;
  b = t - a * (a + 1) / 2;
}
__device__ __forceinline__ int tri_idx(int a, int b) { return a * (a + 1) / 2 + b; }

// In-place Cholesky of the packed SPD matrix G (n x n, element (a,b) at G[tri_idx(a,b)*stride]).
// Returns false on a non-positive pivot.
__device__ __forceinline__ bool chol_packed(double* G, int n, int stride) {
  for (int k = 0; k < n; ++k) {
    double d = G[tri_idx(k, k) * stride];
    for (int q = 0; q < k; ++q) {
      double l = G[tri_idx(k, q) * stride];
      d -= l * l;
    }
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    G[tri_idx(k, k) * stride] = d;
    for (int i = k + 1; i < n; ++i) {
      double s = G[tri_idx(i, k) * stride];
      for (int q = 0; q < k; ++q) s -= G[tri_idx(i, q) * stride] * G[tri_idx(k, q) * stride];
      G[tri_idx(i, k) * stride] = s / d;
    }
  }
  return true;
}
// Solve (L L^T) g = b with the packed factor; b overwritten by g.
__device__ __forceinline__ void chol_solve_packed(const double* L, int n, int stride, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int q = 0; q < i; ++q) s -= L[tri_idx(i, q) * stride] * b[q];
    b[i] = s / L[tri_idx(i, i) * stride];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int q = i + 1; q < n; ++q) s -= L[tri_idx(q, i) * stride] * b[q];
    b[i] = s / L[tri_idx(i, i) * stride];
  }
}

// Geometry of a block deletion: rows/cols [p, p+n) leave an m x m factor.  "New" indices i' address the
// (m-n) x (m-n) result: old row = i' (< p) or i' + n.
struct RmGeom {
  int m, n, p;
  int mnew;   // m - n
  int I0;     // first 64-row tile (new space) that holds updated rows: p / 64
  int nbn;    // tiles of the result: ceil(mnew / 64), at least 1
  int64_t ld;
};

// Workspace (doubles): Bt[nrows][NP] | Gt[nrows][NP] | u[nrows] | Pgram[nT][NG] | Gc[NG] | own[8] | gsol[NP]
struct RmWork {
  double *Bt, *Gt, *u, *Pgram, *Gc, *own;
  double* gsol = nullptr;  // W_BB^-1 alpha_B (k_rm_prep, last tile) for the closed-form alpha of the result
  double* S = nullptr;     // [row tile][64][64]: strict_lower(G_I B_I^T), once per row tile (k_rm_prep) for every column tile
};

// ---------------------------------------------------------------------------------------------
// Per 64-row tile (new space, tiles I0..nbn-1): stage b rows, partial Gram of the tile, carriers' Gram.
__device__ __forceinline__ void rm_gram_tile(const double* __restrict__ M, const RmGeom& g, const RmWork& w, int bx,
                                             bool write_bt) {
  __shared__ double sb[TB][NP + 1];
  const int I = g.I0 + bx;
  for (int idx = threadIdx.x; idx < TB * NP; idx += 256) {
    int r = idx / NP, a = idx % NP;
    int ip = I * TB + r;
    double v = 0.0;
    if (a < g.n && ip >= g.p && ip < g.mnew) v = M[(int64_t)(ip + g.n) * g.ld + g.p + a];
    sb[r][a] = v;
    if (write_bt) w.Bt[((int64_t)bx * TB + r) * NP + a] = v;
  }
  __syncthreads();
  if (threadIdx.x < NG) {
    int a, b;
    tri_unpack(threadIdx.x, a, b);
    double s = 0.0;
#pragma unroll 8
    for (int r = 0; r < TB; ++r) s += sb[r][a] * sb[r][b];
    w.Pgram[(int64_t)bx * NG + threadIdx.x] = s;
    if (bx == 0) {  // carriers: rows p..p+n-1 of the block columns (lower triangular M_BB)
      double c = 0.0;
      if (a < g.n)
        for (int q = a; q < g.n; ++q) c += M[(int64_t)(g.p + q) * g.ld + g.p + a] * M[(int64_t)(g.p + q) * g.ld + g.p + b];
      w.Gc[threadIdx.x] = (a < g.n) ? c : (a == b ? 1.0 : 0.0);  // unused rows: identity keeps chol happy
    }
  }
}
__global__ void __launch_bounds__(256) k_rm_gram(const double* __restrict__ M, RmGeom g, RmWork w) {
  rm_gram_tile(M, g, w, (int)blockIdx.x, true);
}

// Leave-block-out score of the block's own patient under the cluster that still contains it.
//   out[0] = sum_i log N(y_i | mu_i, v_i), out[1] = predictive mean at visit thr_index, out[2] = ok flag
__device__ __forceinline__ void own_final_body(const double* __restrict__ alpha, const double* __restrict__ yB, const RmGeom& g,
                                               const RmWork& w, int nT, const Theta& th, int thr_index,
                                               double* __restrict__ out) {
  __shared__ double G[NG], C[NP][NP], v[NP];
  __shared__ int ok;
  if (threadIdx.x < NG) {
    double s = w.Gc[threadIdx.x];
#pragma unroll 8
    for (int I = 0; I < nT; ++I) s += w.Pgram[(int64_t)I * NG + threadIdx.x];
    G[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) ok = chol_packed(G, g.n, 1) ? 1 : 0;
  __syncthreads();
  if (threadIdx.x < g.n) {  // column threadIdx.x of C = W_BB^-1
    double e[NP];
    for (int i = 0; i < g.n; ++i) e[i] = (i == (int)threadIdx.x) ? 1.0 : 0.0;
    chol_solve_packed(G, g.n, 1, e);
    for (int i = 0; i < g.n; ++i) C[i][threadIdx.x] = e[i];
  }
  __syncthreads();
  if (threadIdx.x < g.n) {
    double s = 0.0;
    for (int q = 0; q < g.n; ++q) s += C[threadIdx.x][q] * alpha[g.p + q];
    v[threadIdx.x] = s;  // y_i - mu_i
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int i = 0; i < g.n; ++i) {
      // C_ii = k** + noise + 1e-8 - q_i; GPy clips v* = k** - q_i at 1e-15 and then adds the noise
      double vstar = C[i][i] - th.noise - GPY_JITTER;
      double vn = vstar > 1e-15 ? C[i][i] - GPY_JITTER : 1e-15 + th.noise;
      total += -0.5 * LOG_2_PI - 0.5 * log(vn) - 0.5 * (v[i] * v[i]) / vn;
    }
    out[0] = total;
    out[1] = (thr_index >= 0 && thr_index < g.n) ? yB[thr_index] - v[thr_index] : nan("");
    out[2] = ok ? 1.0 : 0.0;
  }
}
__global__ void __launch_bounds__(256) k_own_final(const double* __restrict__ alpha, const double* __restrict__ yB, RmGeom g,
                                                   RmWork w, int nT, Theta th, int thr_index, double* __restrict__ out) {
  own_final_body(alpha, yB, g, w, nT, th, thr_index, out);
}

// The same two kernels for several patients at once (the own-cluster scores of a look-ahead batch): one launch pair
// instead of one pair per patient.  The jobs travel in the kernel parameters.
constexpr int OWN_BATCH = 16;
struct OwnJob {
  const double* M;
  const double* alpha;
  const double* yB;
  double* out;   // [3]
  RmGeom g;
  RmWork w;      // only Pgram / Gc are used
  Theta th;
  int nT;
};
struct OwnJobs {
  OwnJob j[OWN_BATCH];
};
__global__ void __launch_bounds__(256) k_own_gram_batch(const __grid_constant__ OwnJobs jobs) {
  const OwnJob& jb = jobs.j[blockIdx.y];
  if ((int)blockIdx.x >= jb.nT) return;
  rm_gram_tile(jb.M, jb.g, jb.w, (int)blockIdx.x, false);
}
__global__ void __launch_bounds__(256) k_own_final_batch(const __grid_constant__ OwnJobs jobs, int thr_index) {
  const OwnJob& jb = jobs.j[blockIdx.x];
  own_final_body(jb.alpha, jb.yB, jb.g, jb.w, jb.nT, jb.th, thr_index, jb.out);
}

// ---------------------------------------------------------------------------------------------
// Rank-n corrections of CACHED scores (look-ahead pipeline, chain.inl).  When a patient's block B leaves cluster a,
// or joins cluster b, the scores of the patients still ahead in the batch under a / b follow from what the scoring
// pass left on the device for them - T_q = M k_q, s_q = |T_q|^2 = k_q^T Ky^-1 k_q, mu_q - in O(m n) per visit,
// without waiting for the O(m^2) update of the factor itself (which proceeds on its own streams):
//   leave:  u = (M^T T_q)_B,  G = W_BB = sum_k b_k b_k^T:   s_q' = s_q - u^T G^-1 u,   mu_q' = mu_q - u^T G^-1 alpha_B
//           (block elimination of (K + s2 I)^-1; the same quantities as the leave-block-out score above);
//   join:   t = Ls^-1 (k(x_B, x_q) - T_B^T T_q)  (the rows the append adds to T_q):   s_q' = s_q + |t|^2,
//           mu_q' = mu_q + t . z_B,  Ls and z_B = Ls^-1 (y_B - mu_B) from k_append_small.
// One CTA per patient ahead; out[2 i] = its new score, out[2 i + 1] = its new mean at the threshold visit.
struct AlgPos {
  const double *T, *ss, *mu;  // the patient's columns: T[c * stride + k], ss[c], mu[c]
  int64_t stride;
  const double *x, *y;        // its visits
  int ncols;
  int thr;                    // which of those columns is the threshold visit (-1: none)
  // optional: the pair's score as its scoring pass left it ON THE DEVICE.  -inf = excluded before the product (k_prune): only
  // the threshold column carries T / ss / mu.  (Corrections queued for a batch whose pass is still in flight cannot know
  // on the host which pairs were excluded; within a batch the host narrows the fields itself and leaves this null.)
  const double* flag;
};
// The columns a correction works on: all of the patient's, or the threshold column alone of an excluded pair.
__device__ __forceinline__ AlgPos alg_eff(const AlgPos& P) {
  AlgPos E = P;
  if (P.flag && *P.flag == -INFINITY && P.thr >= 0 && P.thr < P.ncols) {
    E.T += (int64_t)P.thr * P.stride;
    E.ss += P.thr;
    E.mu += P.thr;
    E.x += P.thr;
    E.y += P.thr;
    E.ncols = 1;
    E.thr = 0;
  }
  return E;
}
struct AlgJobs {
  AlgPos pos[OWN_BATCH];
  Theta th;
  int thr_index;
  double* out;
  // leave
  const double* M;      // the factor BEFORE the deletion (its buffer is not reused until the host has read the result)
  const double* alpha;
  int64_t ld;
  int p, n, m;
  // join
  const double* Tn;     // T of the joining block, Tn[a * stride_n + k]
  int64_t stride_n;
  const double *Li, *znew, *xn;
  // leave: optional copy of what the correction reads from the cluster (block columns of rows p.., alpha_B), so that
  // it can be replayed for the NEXT batch's positions after the cluster's buffer has moved on (chain.inl)
  double* logB;         // [(k - p) * NP + a]
  double* logAlpha;     // [NP]
  // maintain: besides the corrected scores, leave the patients' T / ss / mu ON THE DEVICE as a pass over the changed cluster
  // would: the corrections write ss', mu' back (and a join the n rows it adds to every T column, at row m - needs
  // m + n <= stride); a leave's T follows in k_T_leave.  The cached columns then serve further corrections and appends.
  int maintain;
};

__device__ __forceinline__ void alg_finish(const AlgJobs& J, const AlgPos& P, const double* ssn, const double* mun,
                                           double* lpd /* smem [NP] */) {
  const int q = threadIdx.x;
  if (q < P.ncols) {
    double v = kern_diag(J.th, P.x[q]) - ssn[q];
    v = v > 1e-15 ? v : 1e-15;
    const double vn = v + J.th.noise, d = P.y[q] - mun[q];
    lpd[q] = -0.5 * LOG_2_PI - 0.5 * log(vn) - 0.5 * (d * d) / vn;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the scoring epilogue adds the visits by a shuffle tree over 32 lanes (lane q = visit q): same order here
    double tr[32];
    for (int i = 0; i < 32; ++i) tr[i] = i < P.ncols ? lpd[i] : 0.0;
    for (int o = 16; o; o >>= 1)
      for (int i = 0; i < o; ++i) tr[i] += tr[i + o];
    J.out[2 * blockIdx.x] = tr[0];
    J.out[2 * blockIdx.x + 1] = (P.thr >= 0 && P.thr < P.ncols) ? mun[P.thr] : nan("");
  }
}

constexpr int ALG_THREADS = 512;
constexpr int ALG_ENTRIES = NG + NP * NP;  // G (packed) | U[a][q]
// leave, step 1: grid (row tiles from p, patients ahead): partial G and U over 64 rows of the block's columns
__global__ void __launch_bounds__(256) k_alg_leave_part(const __grid_constant__ AlgJobs J, double* __restrict__ partial) {
  const AlgPos P = alg_eff(J.pos[blockIdx.y]);
  __shared__ double sb[TB][NP + 1], sT[TB][NP + 1];
  const int tid = threadIdx.x, n = J.n, nc = P.ncols;
  const int k0 = J.p + blockIdx.x * TB;
  for (int idx = tid; idx < TB * NP; idx += 256) {
    const int r = idx / NP, a = idx % NP, k = k0 + r;
    double v = 0.0;
    if (a < n && k < J.m && J.p + a <= k) v = J.M[(int64_t)k * J.ld + J.p + a];
    sb[r][a] = v;
    if (J.logB && blockIdx.y == 0 && k < J.m) J.logB[(int64_t)(k - J.p) * NP + a] = v;
  }
  if (J.logAlpha && blockIdx.y == 0 && blockIdx.x == 0 && tid < NP) J.logAlpha[tid] = tid < n ? J.alpha[J.p + tid] : 0.0;
  if (nc == 0) return;  // log only
  for (int idx = tid; idx < TB * NP; idx += 256) {
    const int c = idx / TB, r = idx % TB, k = k0 + r;
    sT[r][c] = (c < nc && k < J.m) ? P.T[(int64_t)c * P.stride + k] : 0.0;
  }
  __syncthreads();
  double* out = partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * ALG_ENTRIES;
  for (int e = tid; e < ALG_ENTRIES; e += 256) {
    double acc = 0.0;
    if (e < NG) {
      int ea, eb;
      tri_unpack(e, ea, eb);
      if (ea < n) {
#pragma unroll 8
        for (int r = 0; r < TB; ++r) acc += sb[r][ea] * sb[r][eb];
      }
    } else {
      const int ea = (e - NG) / NP, eb = (e - NG) % NP;
      if (ea < n && eb < nc) {
#pragma unroll 8
        for (int r = 0; r < TB; ++r) acc += sb[r][ea] * sT[r][eb];
      }
    }
    out[e] = acc;
  }
}
// leave, step 2: one CTA per patient ahead: add the partials in a fixed order, G^-1 u, new moments, new score
__global__ void __launch_bounds__(256) k_alg_leave_fin(const __grid_constant__ AlgJobs J, const double* __restrict__ partial,
                                                       int ntiles) {
  const AlgPos P = alg_eff(J.pos[blockIdx.x]);
  __shared__ double G[NG], U[NP][NP + 1], ssn[NP], mun[NP], lpd[NP];
  __shared__ int ok;
  const int tid = threadIdx.x, n = J.n, nc = P.ncols;
  for (int e = tid; e < ALG_ENTRIES; e += 256) {
    double acc = 0.0;
#pragma unroll 8
    for (int t = 0; t < ntiles; ++t) acc += partial[((int64_t)blockIdx.x * ntiles + t) * ALG_ENTRIES + e];
    if (e < NG) {
      int ea, eb;
      tri_unpack(e, ea, eb);
      G[e] = (ea < n) ? acc : (ea == eb ? 1.0 : 0.0);  // unused rows: identity keeps chol happy
    } else {
      U[(e - NG) / NP][(e - NG) % NP] = acc;
    }
  }
  __syncthreads();
  if (tid == 0) ok = chol_packed(G, n, 1) ? 1 : 0;
  __syncthreads();
  if (tid < nc) {
    double g[NP];
    for (int a = 0; a < n; ++a) g[a] = U[a][tid];
    chol_solve_packed(G, n, 1, g);
    double ds = 0.0, dm = 0.0;
    for (int a = 0; a < n; ++a) {
      ds += g[a] * U[a][tid];
      dm += g[a] * J.alpha[J.p + a];
    }
    ssn[tid] = ok ? P.ss[tid] - ds : nan("");
    mun[tid] = P.mu[tid] - dm;
    if (J.maintain) {
      const_cast<double*>(P.ss)[tid] = ssn[tid];
      const_cast<double*>(P.mu)[tid] = mun[tid];
    }
  }
  __syncthreads();
  alg_finish(J, P, ssn, mun, lpd);
}

// join, step 1: grid (row tiles, patients ahead): partial C[a][q] = sum_k T_B[a][k] T_q[k] over 64 rows
__global__ void __launch_bounds__(256) k_alg_join_part(const __grid_constant__ AlgJobs J, double* __restrict__ partial) {
  const AlgPos P = alg_eff(J.pos[blockIdx.y]);
  __shared__ double sa[TB][NP + 1], sT[TB][NP + 1];
  const int tid = threadIdx.x, n = J.n, nc = P.ncols;
  const int64_t len = P.stride < J.stride_n ? P.stride : J.stride_n;  // both are the cluster's padded size
  const int64_t k0 = (int64_t)blockIdx.x * TB;
  for (int idx = tid; idx < TB * NP; idx += 256) {
    const int c = idx / TB, r = idx % TB;
    const int64_t k = k0 + r;
    sa[r][c] = (c < n && k < len) ? J.Tn[(int64_t)c * J.stride_n + k] : 0.0;
    sT[r][c] = (c < nc && k < len) ? P.T[(int64_t)c * P.stride + k] : 0.0;
  }
  __syncthreads();
  const int ea = tid / NP, eb = tid % NP;  // NP * NP = 256 entries
  double acc = 0.0;
  if (ea < n && eb < nc) {
#pragma unroll 8
    for (int r = 0; r < TB; ++r) acc += sa[r][ea] * sT[r][eb];
  }
  partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (NP * NP) + tid] = acc;
}
// join, step 2 (after k_append_small): one CTA per patient ahead
__global__ void __launch_bounds__(256) k_alg_join_fin(const __grid_constant__ AlgJobs J, const double* __restrict__ partial,
                                                      int ntiles) {
  const AlgPos P = alg_eff(J.pos[blockIdx.x]);
  __shared__ double C[NP][NP + 1], ssn[NP], mun[NP], lpd[NP];
  const int tid = threadIdx.x, n = J.n, nc = P.ncols;
  {
    double acc = 0.0;
#pragma unroll 8
    for (int t = 0; t < ntiles; ++t) acc += partial[((int64_t)blockIdx.x * ntiles + t) * (NP * NP) + tid];
    C[tid / NP][tid % NP] = acc;
  }
  __syncthreads();
  if (tid < nc) {
    double w[NP];
    for (int a = 0; a < n; ++a) w[a] = kern_eval(J.th, J.xn[a], P.x[tid], false, nullptr) - C[a][tid];
    double s2 = 0.0, dm = 0.0;
    for (int i = 0; i < n; ++i) {
      double t = 0.0;
      for (int c = 0; c <= i; ++c) t += J.Li[i * NP + c] * w[c];
      s2 += t * t;
      dm += t * J.znew[i];
      if (J.maintain) const_cast<double*>(P.T)[(int64_t)tid * P.stride + J.m + i] = t;  // the rows the append adds to T_q
    }
    ssn[tid] = P.ss[tid] + s2;
    mun[tid] = P.mu[tid] + dm;
    if (J.maintain) {
      const_cast<double*>(P.ss)[tid] = ssn[tid];
      const_cast<double*>(P.mu)[tid] = mun[tid];
    }
  }
  __syncthreads();
  alg_finish(J, P, ssn, mun, lpd);
}

// Per tile: G_{i-1} for each of its rows (running Gram, snapshots in shared memory), then one thread per
// row solves g_i = G_{i-1}^-1 b_i and u_i = (1 + b_i.g_i)^-1/2.
constexpr int RM_PREP_SMEM = (NG * (TB + 1) + 2 * TB * (NP + 1)) * (int)sizeof(double);
__global__ void __launch_bounds__(256) k_rm_prep(RmGeom g, RmWork w, const double* __restrict__ alpha_old) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double Gfull[NG];
  double* snap = smem;                  // [NG][TB+1]
  double* sb = snap + NG * (TB + 1);    // [TB][NP+1]
  double* sg = sb + TB * (NP + 1);      // [TB][NP+1]: g_i
  const int I = g.I0 + blockIdx.x;
  for (int idx = threadIdx.x; idx < TB * NP; idx += 256) {
    int r = idx / NP, a = idx % NP;
    sb[r * (NP + 1) + a] = w.Bt[((int64_t)blockIdx.x * TB + r) * NP + a];
  }
  __syncthreads();
  if (threadIdx.x < NG) {
    int a, b;
    tri_unpack(threadIdx.x, a, b);
    double run = w.Gc[threadIdx.x];
#pragma unroll 8
    for (int J = 0; J < (int)blockIdx.x; ++J) run += w.Pgram[(int64_t)J * NG + threadIdx.x];
    for (int k = 0; k < TB; ++k) {
      snap[threadIdx.x * (TB + 1) + k] = run;
      run += sb[k * (NP + 1) + a] * sb[k * (NP + 1) + b];
    }
    Gfull[threadIdx.x] = (a < g.n) ? run : (a == b ? 1.0 : 0.0);  // the last tile ends with G = W_BB
  }
  __syncthreads();
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255 && w.gsol) {
    // alpha of the result in closed form: alpha'_R = alpha_R - W_RB W_BB^-1 alpha_B (k_rm_alpha); here W_BB^-1 alpha_B
    double v[NP];
    const bool ok = chol_packed(Gfull, g.n, 1);
    for (int q = 0; q < g.n; ++q) v[q] = alpha_old[g.p + q];
    if (ok) chol_solve_packed(Gfull, g.n, 1, v);
    for (int q = 0; q < NP; ++q) w.gsol[q] = (q < g.n) ? (ok ? v[q] : nan("")) : 0.0;
  }
  if (threadIdx.x < TB) {
    const int r = threadIdx.x, ip = I * TB + r;
    double gv[NP];
    double u = 1.0;
#pragma unroll
    for (int a = 0; a < NP; ++a) gv[a] = 0.0;
    if (ip >= g.p && ip < g.mnew) {
      double* Gr = snap + r;  // element t at Gr[t*(TB+1)]
      bool ok = chol_packed(Gr, g.n, TB + 1);
      for (int a = 0; a < g.n; ++a) gv[a] = sb[r * (NP + 1) + a];
      if (ok) {
        chol_solve_packed(Gr, g.n, TB + 1, gv);
        double s = 0.0;
        for (int a = 0; a < g.n; ++a) s += gv[a] * sb[r * (NP + 1) + a];
        u = 1.0 / sqrt(1.0 + s);
      } else {
        u = nan("");
      }
    }
    for (int a = 0; a < NP; ++a) {
      const double ga = (a < g.n) ? gv[a] : 0.0;
      w.Gt[((int64_t)blockIdx.x * TB + r) * NP + a] = ga;
      sg[r * (NP + 1) + a] = ga;
    }
    w.u[(int64_t)blockIdx.x * TB + r] = u;
  }
  __syncthreads();
  if (w.S) {  // S_I = strict_lower(G_I B_I^T): the in-tile part of the prefix, the same for every column tile
    double* So = w.S + (int64_t)blockIdx.x * (TB * TB);
    for (int idx = threadIdx.x; idx < TB * TB; idx += 256) {
      const int r = idx >> 6, k = idx & 63;
      double sv = 0.0;
      if (k < r) {
#pragma unroll
        for (int a = 0; a < NP; ++a) sv += sg[r * (NP + 1) + a] * sb[k * (NP + 1) + a];
      }
      So[idx] = sv;
    }
  }
}

// Element (ip, jp) of the result's source: the old factor with rows/cols [p, p+n) skipped, identity padding
// beyond mnew, zero strictly above the diagonal (tiles above the diagonal are never written: do not read them).
__device__ __forceinline__ double rm_src(const double* __restrict__ Mo, const RmGeom& g, int ip, int jp) {
  if (ip >= g.mnew || jp >= g.mnew) return (ip == jp) ? 1.0 : 0.0;
  const int i = ip < g.p ? ip : ip + g.n, j = jp < g.p ? jp : jp + g.n;
  if (j > i) return 0.0;
  return Mo[(int64_t)i * g.ld + j];
}

// h_{i-1}[:, j] is a prefix sum over rows; it is formed in three fully parallel steps over the lower tiles
// (I >= J, I >= I0) of the result:
//   k_rm_hpart : P[I][J] = B_I^T * old tile (I, J)                                (n x 64 per tile)
//   k_rm_hscan : H[I][J] = carriers' term + sum_{I' < I} P[I'][J]                  (exclusive scan down each J)
//   k_rm_apply : new tile = diag(u) (old - S_I old - G_I H[I][J]),  S_I = strict_lower(G_I B_I^T)   (DMMA)
// P and H are stored as [I - I0][J][NP][64].
__global__ void __launch_bounds__(NTHREADS) k_rm_hpart(const double* __restrict__ Mo, RmGeom g, RmWork w,
                                                       double* __restrict__ P) {
  const int J = blockIdx.x, I = g.I0 + blockIdx.y;
  if (J > I) return;
  __shared__ double sT[TB][TB + 1];
  __shared__ double sb[TB][NP + 1];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
    int r = idx >> 6, c = idx & 63;
    sT[r][c] = rm_src(Mo, g, I * TB + r, J * TB + c);
  }
  for (int idx = tid; idx < TB * NP; idx += NTHREADS) {
    int r = idx / NP, a = idx % NP;
    sb[r][a] = w.Bt[((int64_t)blockIdx.y * TB + r) * NP + a];
  }
  __syncthreads();
  double* out = P + ((int64_t)blockIdx.y * g.nbn + J) * (NP * TB);
  for (int idx = tid; idx < NP * TB; idx += NTHREADS) {
    int q = idx >> 6, c = idx & 63;
    double s = 0.0;
    if (q < g.n) {
#pragma unroll 8
      for (int k = 0; k < TB; ++k) s += sb[k][q] * sT[k][c];
    }
    out[idx] = s;
  }
}

// grid nbn (one CTA per column block), NP*64 threads: thread (q, c) walks down the row tiles.
__global__ void __launch_bounds__(NP * TB) k_rm_hscan(const double* __restrict__ Mo, RmGeom g, const double* __restrict__ P,
                                                      double* __restrict__ H) {
  const int J = blockIdx.x, q = threadIdx.x >> 6, c = threadIdx.x & 63;
  const int Ifirst = max(J, g.I0);
  // carriers: rows p..p+n-1 contribute b_k[q] M[k, j]
  double run = 0.0;
  const int jp = J * TB + c;
  if (q < g.n && jp < g.mnew) {
    const int j = jp < g.p ? jp : jp + g.n;
    for (int a = q; a < g.n; ++a)
      if (j <= g.p + a) run += Mo[(int64_t)(g.p + a) * g.ld + g.p + q] * Mo[(int64_t)(g.p + a) * g.ld + j];
  }
  for (int I = Ifirst; I < g.nbn; ++I) {
    const int64_t off = ((int64_t)(I - g.I0) * g.nbn + J) * (NP * TB) + threadIdx.x;
    H[off] = run;
    run += P[off];
  }
}

constexpr int RM_APPLY_SMEM = (TB * LD_W + (TB + NP) * LD_R + TB) * (int)sizeof(double);
__global__ void __launch_bounds__(NTHREADS) k_rm_apply(const double* __restrict__ Mo, double* __restrict__ Mn, RmGeom g,
                                                       RmWork w, const double* __restrict__ H) {
  const int J = blockIdx.x, I = blockIdx.y, tid = threadIdx.x;
  if (J > I) return;
  const int64_t ld = g.ld;
  if (I < g.I0) {  // rows above the deleted block: plain copy
    for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
      int r = idx >> 6, c = idx & 63;
      Mn[(int64_t)(I * TB + r) * ld + J * TB + c] = rm_src(Mo, g, I * TB + r, J * TB + c);
    }
    return;
  }
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;                       // [64][LD_W]: S (cols 0..63) | G (cols 64..64+NP)
  double* sB = sA + TB * LD_W;             // [64+NP][LD_R]: old tile (rows 0..63) ; H (rows 64..64+NP)
  double* su = sB + (TB + NP) * LD_R;      // [64]
  const int64_t wrow = (int64_t)(I - g.I0) * TB;
  const double* Hs = H + ((int64_t)(I - g.I0) * g.nbn + J) * (NP * TB);
  const double* Ss = w.S + (int64_t)(I - g.I0) * (TB * TB);
  for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
    int r = idx >> 6, c = idx & 63;
    sB[r * LD_R + c] = rm_src(Mo, g, I * TB + r, J * TB + c);
    sA[r * LD_W + c] = Ss[idx];
  }
  for (int idx = tid; idx < NP * TB; idx += NTHREADS) sB[(TB + (idx >> 6)) * LD_R + (idx & 63)] = Hs[idx];
  for (int idx = tid; idx < TB * NP; idx += NTHREADS) {
    int r = idx / NP, a = idx % NP;
    sA[r * LD_W + TB + a] = w.Gt[(wrow + r) * NP + a];
  }
  if (tid < TB) su[tid] = w.u[wrow + tid];
  __syncthreads();
  Acc64 acc;
  acc.zero();
  warp_mma<2, 4, true, false, TB + NP, LD_W, LD_R>(acc, sA, sB, warp_rb64(), warp_cb64());
  for_each_acc64(acc, [&](int r, int c, double& v) {
    Mn[(int64_t)(I * TB + r) * ld + J * TB + c] = su[r] * (sB[r * LD_R + c] - v);
  });
}

// The deletion in ONE pass over the factor (replaces k_rm_hpart + k_rm_hscan + k_rm_apply + k_rm_vec + k_rm_alpha): one CTA
// per 64-column strip J walks down the row tiles I = max(J, I0) .. nbn-1 with the prefix h (n x 64) carried along:
//     new tile = diag(u) (old - [S_I | G_I] [old ; H]),   then   H += B_I^T old
// (both products on the FP64 tensor pipe), the next row tile's operands already in flight (cp.async double buffer; the
// sources are shifted by the deleted block, so the old tile is gathered in 8-byte pieces).  Every element of M is read
// once and written once.  The rows above the block are a plain copy (same CTA, before the walk).  After the last tile H
// is W_{B, strip}: alpha of the result in closed form, alpha'_j = alpha_j - W_{B,j} . (W_BB^-1 alpha_B), and x, y of the
// strip move to their new places.
// Strips are RM_SW = 16 columns wide: the walk of a strip is a chain of dependent row tiles (each needs the prefix of
// the one above), its length - not the traffic - is what a waiting re-scoring pass sees, and a narrow strip makes every
// link short (64 x 80 x 16 products) while 4 nb strips keep every SM busy.
constexpr int LD_BT = 20;  // smem ld of the [64][NP] block rows of the second product
constexpr int RM_SW = 16;  // strip width
constexpr int LD_SW = 20;  // smem ld of the strip's [64 + NP][RM_SW] operand (4 mod 16: conflict-free fragment loads)
struct RmStripStage {      // one row tile's operands in shared memory
  double sB[(TB + NP) * LD_SW];  // old tile slice (rows 0..63) ; H (rows 64..64+NP)
  double sA[TB * LD_W];          // S (cols 0..63) | G (cols 64..64+NP)
  double sBt[TB * LD_BT];        // b rows of the tile: [k][a]
  double su[TB];
};
constexpr int RM_STRIP_SMEM = 2 * (int)sizeof(RmStripStage);

__global__ void __launch_bounds__(NTHREADS) k_rm_strip(const double* __restrict__ Mo, double* __restrict__ Mn, RmGeom g, RmWork w,
                                                       const double* __restrict__ vec_old, double* __restrict__ vec_new,
                                                       int64_t cap) {
  extern __shared__ __align__(16) unsigned char rm_smem_raw[];
  RmStripStage* st = reinterpret_cast<RmStripStage*>(rm_smem_raw);
  __shared__ double sH[NP][RM_SW];  // the prefix carried down the strip
  const int strip = blockIdx.x, tid = threadIdx.x;
  const int col0 = strip * RM_SW, J = col0 / TB;  // first column, block column
  const int64_t ld = g.ld;
  // rows above the deleted block: neither rows nor columns move (j <= i < p)
  for (int I = J; I < g.I0; ++I)
    for (int idx = tid; idx < TB * RM_SW / 2; idx += NTHREADS) {
      const int r = idx / (RM_SW / 2), c = (idx % (RM_SW / 2)) * 2;
      const int i = I * TB + r, j = col0 + c;
      const int64_t o = (int64_t)i * ld + j;
      double2 v = *reinterpret_cast<const double2*>(Mo + o);
      if (j > i) v.x = 0.0;  // (inside the diagonal tile the upper part is kept at zero)
      if (j + 1 > i) v.y = 0.0;
      *reinterpret_cast<double2*>(Mn + o) = v;
    }
  const int Ifirst = max(J, g.I0);
  // carriers: rows p..p+n-1 of the block's own columns contribute b_k[q] M[k, j] to every prefix
  {
    const int q = tid / RM_SW, c = tid % RM_SW, jp = col0 + c;  // NP * RM_SW = 256 = NTHREADS entries
    double run = 0.0;
    if (q < g.n && jp < g.mnew) {
      const int j = jp < g.p ? jp : jp + g.n;
      for (int a = q; a < g.n; ++a)
        if (j <= g.p + a) run += Mo[(int64_t)(g.p + a) * ld + g.p + q] * Mo[(int64_t)(g.p + a) * ld + j];
    }
    sH[q][c] = run;
  }
  auto issue = [&](int I, RmStripStage& S) {  // operands of row tile I (cp.async; constants stored directly)
    const int bx = I - g.I0;
    for (int idx = tid; idx < TB * RM_SW; idx += NTHREADS) {
      const int r = idx / RM_SW, c = idx % RM_SW, ip = I * TB + r, jp = col0 + c;
      double* dst = S.sB + r * LD_SW + c;
      if (ip >= g.mnew || jp >= g.mnew) {
        *dst = (ip == jp) ? 1.0 : 0.0;
      } else {
        const int i = ip < g.p ? ip : ip + g.n, j = jp < g.p ? jp : jp + g.n;
        if (j > i) *dst = 0.0;
        else cp_async8(dst, Mo + (int64_t)i * ld + j);
      }
    }
    const double* Ss = w.S + (int64_t)bx * (TB * TB);
    for (int idx = tid; idx < TB * TB / 2; idx += NTHREADS) {
      const int r = idx >> 5, c = (idx & 31) << 1;
      cp_async16(S.sA + r * LD_W + c, Ss + r * TB + c);
    }
    for (int idx = tid; idx < TB * NP / 2; idx += NTHREADS) {
      const int r = idx / (NP / 2), a = (idx % (NP / 2)) * 2;
      cp_async16(S.sA + r * LD_W + TB + a, w.Gt + ((int64_t)bx * TB + r) * NP + a);
      cp_async16(S.sBt + r * LD_BT + a, w.Bt + ((int64_t)bx * TB + r) * NP + a);
    }
    if (tid < TB / 2) cp_async16(S.su + 2 * tid, w.u + (int64_t)bx * TB + 2 * tid);
  };
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  if (Ifirst < g.nbn) issue(Ifirst, st[0]);
  cp_async_commit();
  int cur = 0;
  for (int I = Ifirst; I < g.nbn; ++I, cur ^= 1) {
    RmStripStage& S = st[cur];
    if (I + 1 < g.nbn) issue(I + 1, st[cur ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();  // this thread's copies of tile I have landed ...
    S.sB[(TB + tid / RM_SW) * LD_SW + tid % RM_SW] = sH[tid / RM_SW][tid % RM_SW];
    __syncthreads();     // ... everybody's have, and H is in place
    // new = diag(u) (old - [S | G] [old ; H]): 64 x 16, warp w owns rows [8w, 8w + 8)
    Acc<1, 2> acc;
    acc.zero();
    {
      // S is STRICTLY lower triangular: rows [8w, 8w + 8) only see k < 8w + 8 of it; then the G H part (k = 64 .. 64 + NP)
      const double* arow = S.sA + (8 * warp + gq) * LD_W;
      auto kstep = [&](int kb) {
        const double a = arow[kb + tq];
        const double b0 = S.sB[(kb + tq) * LD_SW + gq], b1 = S.sB[(kb + tq) * LD_SW + 8 + gq];
        dmma884(acc.v[0][0][0], acc.v[0][0][1], a, b0);
        dmma884(acc.v[0][1][0], acc.v[0][1][1], a, b1);
      };
      for (int kb = 0; kb < 8 * warp + 8; kb += 4) kstep(kb);
#pragma unroll
      for (int kb = TB; kb < TB + NP; kb += 4) kstep(kb);
    }
    {
      const int r = 8 * warp + gq;
      double* orow = Mn + (int64_t)(I * TB + r) * ld + col0;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 8 * j + 2 * tq;
        orow[c] = S.su[r] * (S.sB[r * LD_SW + c] - acc.v[0][j][0]);
        orow[c + 1] = S.su[r] * (S.sB[r * LD_SW + c + 1] - acc.v[0][j][1]);
      }
    }
    // H += B_I^T old: 16 x 16 in four 8 x 8 blocks (warps 0..3)
    if (warp < 4) {
      Acc<1, 1> hp;
      hp.zero();
      warp_mma<1, 1, false, false, TB, LD_BT, LD_SW>(hp, S.sBt, S.sB, 8 * (warp >> 1), 8 * (warp & 1));
      const int q = 8 * (warp >> 1) + gq, c = 8 * (warp & 1) + 2 * tq;
      sH[q][c] += hp.v[0][0][0];
      sH[q][c + 1] += hp.v[0][0][1];
    }
    __syncthreads();  // the stage may be refilled, H is complete for the next tile
  }
  cp_async_wait<0>();
  // closed-form alpha of the result and the strip's share of x, y  (vec = x | y | alpha | z, `cap` doubles each)
  if (tid < RM_SW) {
    const int jp = col0 + tid;
    if (jp < g.mnew) {
      const int j = jp < g.p ? jp : jp + g.n;
      double sdot = 0.0;
      for (int a = 0; a < g.n; ++a) sdot += sH[a][tid] * w.gsol[a];
      vec_new[jp] = vec_old[j];
      vec_new[cap + jp] = vec_old[cap + j];
      vec_new[2 * cap + jp] = vec_old[2 * cap + j] - sdot;
    }
  }
  if (strip == 0 && tid < 32) vec_new[4 * cap + tid] = vec_old[4 * cap + tid];  // the stat block travels with the buffer
}

// The same walk over 64-column strips (one CTA per block column): a quarter of the CTAs, four times the work per link.  The
// chain of dependent row tiles takes longer than with 16-column strips, but the deletion holds far fewer SMs while it runs
// (every CTA of it displaces a CTA of the scoring pass in flight): the better trade once nobody waits for the deletion
// itself (cached columns kept current, chain.inl).  MOGP_RM_SW=64.
struct RmStrip64Stage {      // one row tile's operands in shared memory
  double sB[(TB + NP) * LD_R];   // old tile (rows 0..63) ; H (rows 64..64+NP)
  double sA[TB * LD_W];          // S (cols 0..63) | G (cols 64..64+NP)
  double sBt[TB * LD_BT];        // b rows of the tile: [k][a]
  double su[TB];
};
constexpr int RM_STRIP64_SMEM = 2 * (int)sizeof(RmStrip64Stage);

__global__ void __launch_bounds__(NTHREADS) k_rm_strip64(const double* __restrict__ Mo, double* __restrict__ Mn, RmGeom g, RmWork w,
                                                       const double* __restrict__ vec_old, double* __restrict__ vec_new,
                                                       int64_t cap) {
  extern __shared__ __align__(16) unsigned char rm64_smem_raw[];
  RmStrip64Stage* st = reinterpret_cast<RmStrip64Stage*>(rm64_smem_raw);
  __shared__ double sH[NP][TB];  // the prefix carried down the strip
  const int J = blockIdx.x, tid = threadIdx.x;
  const int64_t ld = g.ld;
  // rows above the deleted block: neither rows nor columns move (j <= i < p)
  for (int I = J; I < g.I0; ++I)
    for (int idx = tid; idx < TB * TB / 2; idx += NTHREADS) {
      const int r = idx >> 5, c = (idx & 31) << 1;
      const int64_t o = (int64_t)(I * TB + r) * ld + J * TB + c;
      double2 v = *reinterpret_cast<const double2*>(Mo + o);
      if (I == J) {  // (tiles above the diagonal hold nothing; inside the diagonal tile the upper part is kept at zero)
        if (c > r) v.x = 0.0;
        if (c + 1 > r) v.y = 0.0;
      }
      *reinterpret_cast<double2*>(Mn + o) = v;
    }
  const int Ifirst = max(J, g.I0);
  // carriers: rows p..p+n-1 of the block's own columns contribute b_k[q] M[k, j] to every prefix
  for (int idx = tid; idx < NP * TB; idx += NTHREADS) {
    const int q = idx >> 6, c = idx & 63, jp = J * TB + c;
    double run = 0.0;
    if (q < g.n && jp < g.mnew) {
      const int j = jp < g.p ? jp : jp + g.n;
      for (int a = q; a < g.n; ++a)
        if (j <= g.p + a) run += Mo[(int64_t)(g.p + a) * ld + g.p + q] * Mo[(int64_t)(g.p + a) * ld + j];
    }
    sH[q][c] = run;
  }
  auto issue = [&](int I, RmStrip64Stage& S) {  // operands of row tile I (cp.async; constants stored directly)
    const int bx = I - g.I0;
    for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
      const int r = idx >> 6, c = idx & 63, ip = I * TB + r, jp = J * TB + c;
      double* dst = S.sB + r * LD_R + c;
      if (ip >= g.mnew || jp >= g.mnew) {
        *dst = (ip == jp) ? 1.0 : 0.0;
      } else {
        const int i = ip < g.p ? ip : ip + g.n, j = jp < g.p ? jp : jp + g.n;
        if (j > i) *dst = 0.0;
        else cp_async8(dst, Mo + (int64_t)i * ld + j);
      }
    }
    const double* Ss = w.S + (int64_t)bx * (TB * TB);
    for (int idx = tid; idx < TB * TB / 2; idx += NTHREADS) {
      const int r = idx >> 5, c = (idx & 31) << 1;
      cp_async16(S.sA + r * LD_W + c, Ss + r * TB + c);
    }
    for (int idx = tid; idx < TB * NP / 2; idx += NTHREADS) {
      const int r = idx / (NP / 2), a = (idx % (NP / 2)) * 2;
      cp_async16(S.sA + r * LD_W + TB + a, w.Gt + ((int64_t)bx * TB + r) * NP + a);
      cp_async16(S.sBt + r * LD_BT + a, w.Bt + ((int64_t)bx * TB + r) * NP + a);
    }
    if (tid < TB / 2) cp_async16(S.su + 2 * tid, w.u + (int64_t)bx * TB + 2 * tid);
  };
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  if (Ifirst < g.nbn) issue(Ifirst, st[0]);
  cp_async_commit();
  int cur = 0;
  for (int I = Ifirst; I < g.nbn; ++I, cur ^= 1) {
    RmStrip64Stage& S = st[cur];
    if (I + 1 < g.nbn) issue(I + 1, st[cur ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();  // this thread's copies of tile I have landed ...
    for (int idx = tid; idx < NP * TB; idx += NTHREADS) S.sB[(TB + (idx >> 6)) * LD_R + (idx & 63)] = sH[idx >> 6][idx & 63];
    __syncthreads();     // ... everybody's have, and H is in place
    Acc64 acc;
    acc.zero();
    warp_mma<2, 4, true, false, TB + NP, LD_W, LD_R>(acc, S.sA, S.sB, warp_rb64(), warp_cb64());
    for_each_acc64(acc, [&](int r, int c, double& v) {
      Mn[(int64_t)(I * TB + r) * ld + J * TB + c] = S.su[r] * (S.sB[r * LD_R + c] - v);
    });
    // H += B_I^T old: 16 x 64, warp w owns columns [8w, 8w + 8)
    Acc<2, 1> hp;
    hp.zero();
    warp_mma<2, 1, false, false, TB, LD_BT, LD_R>(hp, S.sBt, S.sB, 0, warp * 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int q = 8 * i + gq, c = warp * 8 + 2 * tq;
      sH[q][c] += hp.v[i][0][0];
      sH[q][c + 1] += hp.v[i][0][1];
    }
    __syncthreads();  // the stage may be refilled, H is complete for the next tile
  }
  cp_async_wait<0>();
  // closed-form alpha of the result and the strip's share of x, y  (vec = x | y | alpha | z, `cap` doubles each)
  if (tid < TB) {
    const int jp = J * TB + tid;
    if (jp < g.mnew) {
      const int j = jp < g.p ? jp : jp + g.n;
      double sdot = 0.0;
      for (int a = 0; a < g.n; ++a) sdot += sH[a][tid] * w.gsol[a];
      vec_new[jp] = vec_old[j];
      vec_new[cap + jp] = vec_old[cap + j];
      vec_new[2 * cap + jp] = vec_old[2 * cap + j] - sdot;
    }
  }
  if (J == 0 && tid < 32) vec_new[4 * cap + tid] = vec_old[4 * cap + tid];  // the stat block travels with the buffer
}

// The same row operations applied to cached columns T_q = M k_q of the patients still ahead (look-ahead pipeline, chain.inl):
// with M' E = G M (E drops the block's columns) T_q' = M' k_q' = G T_q - the contributions of the block's entries of k_q cancel
// (g_i . G_{i-1} k_B = b_i . k_B) - so T_q' [i'] = u_i (T_q[i] - g_i . h_{i-1}), h_i = sum_{k <= i} b_k T_q[k] from the block's own
// rows on: a column of T is transformed exactly like a column of M left of the block.  One CTA per patient (<= NP = RM_SW
// columns: one strip), in place (a row tile is written where rows already consumed were read from), O(m n) per column
// instead of the O(m^2) product against the updated factor.  w: the deletion's operands (a private copy: the deletion's own
// workspace moves on); J.M / J.ld / J.p address the block's columns of the old factor (the logged copy).
__global__ void __launch_bounds__(NTHREADS) k_T_leave(const __grid_constant__ AlgJobs J, RmGeom g, RmWork w) {
  extern __shared__ __align__(16) unsigned char rm_smem_raw[];
  RmStripStage* st = reinterpret_cast<RmStripStage*>(rm_smem_raw);
  __shared__ double sH[NP][RM_SW];
  const AlgPos P = alg_eff(J.pos[blockIdx.x]);
  const int tid = threadIdx.x, nc = P.ncols;
  if (nc <= 0) return;
  double* T = const_cast<double*>(P.T);
  const int64_t str = P.stride;
  {  // carriers: rows p .. p+n-1 (b_k = M[k, p .. p+n), lower triangular inside the block)
    const int q = tid / RM_SW, c = tid % RM_SW;
    double run = 0.0;
    if (q < g.n && c < nc)
      for (int a = q; a < g.n; ++a) run += J.M[(int64_t)(g.p + a) * J.ld + J.p + q] * T[(int64_t)c * str + g.p + a];
    sH[q][c] = run;
  }
  auto issue = [&](int I, RmStripStage& S) {
    const int bx = I - g.I0;
    for (int idx = tid; idx < TB * RM_SW; idx += NTHREADS) {
      const int c = idx / TB, r = idx % TB, ip = I * TB + r;  // consecutive threads: consecutive rows of one column
      double* dst = S.sB + r * LD_SW + c;
      if (ip >= g.mnew || c >= nc) *dst = 0.0;
      else cp_async8(dst, T + (int64_t)c * str + (ip < g.p ? ip : ip + g.n));
    }
    const double* Ss = w.S + (int64_t)bx * (TB * TB);
    for (int idx = tid; idx < TB * TB / 2; idx += NTHREADS) {
      const int r = idx >> 5, c = (idx & 31) << 1;
      cp_async16(S.sA + r * LD_W + c, Ss + r * TB + c);
    }
    for (int idx = tid; idx < TB * NP / 2; idx += NTHREADS) {
      const int r = idx / (NP / 2), a = (idx % (NP / 2)) * 2;
      cp_async16(S.sA + r * LD_W + TB + a, w.Gt + ((int64_t)bx * TB + r) * NP + a);
      cp_async16(S.sBt + r * LD_BT + a, w.Bt + ((int64_t)bx * TB + r) * NP + a);
    }
    if (tid < TB / 2) cp_async16(S.su + 2 * tid, w.u + (int64_t)bx * TB + 2 * tid);
  };
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  issue(g.I0, st[0]);
  cp_async_commit();
  int cur = 0;
  for (int I = g.I0; I < g.nbn; ++I, cur ^= 1) {
    RmStripStage& S = st[cur];
    if (I + 1 < g.nbn) issue(I + 1, st[cur ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    S.sB[(TB + tid / RM_SW) * LD_SW + tid % RM_SW] = sH[tid / RM_SW][tid % RM_SW];
    __syncthreads();
    Acc<1, 2> acc;
    acc.zero();
    {
      const double* arow = S.sA + (8 * warp + gq) * LD_W;
      auto kstep = [&](int kb) {
        const double a = arow[kb + tq];
        const double b0 = S.sB[(kb + tq) * LD_SW + gq], b1 = S.sB[(kb + tq) * LD_SW + 8 + gq];
        dmma884(acc.v[0][0][0], acc.v[0][0][1], a, b0);
        dmma884(acc.v[0][1][0], acc.v[0][1][1], a, b1);
      };
      for (int kb = 0; kb < 8 * warp + 8; kb += 4) kstep(kb);
#pragma unroll
      for (int kb = TB; kb < TB + NP; kb += 4) kstep(kb);
    }
    {
      const int r = 8 * warp + gq, ip = I * TB + r;
      if (ip >= g.p) {  // (rows above the block stay as they are)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * j + 2 * tq + e;
            if (c < nc) T[(int64_t)c * str + ip] = ip < g.mnew ? S.su[r] * (S.sB[r * LD_SW + c] - acc.v[0][j][e]) : 0.0;
          }
      }
    }
    if (warp < 4) {
      Acc<1, 1> hp;
      hp.zero();
      warp_mma<1, 1, false, false, TB, LD_BT, LD_SW>(hp, S.sBt, S.sB, 8 * (warp >> 1), 8 * (warp & 1));
      const int q = 8 * (warp >> 1) + gq, c = 8 * (warp & 1) + 2 * tq;
      sH[q][c] += hp.v[0][0][0];
      sH[q][c + 1] += hp.v[0][0][1];
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  // the result is n rows shorter: rows [mnew, m) beyond the last tile written become padding (zero) again
  for (int idx = tid; idx < nc * g.n; idx += NTHREADS) {
    const int c = idx / g.n, ip = g.mnew + idx % g.n;
    if (ip >= g.nbn * TB && ip < str) T[(int64_t)c * str + ip] = 0.0;
  }
}

// x, y of the result (new order = old order with the block dropped).
__global__ void k_rm_vec(const double* __restrict__ xo, const double* __restrict__ yo, double* __restrict__ xn,
                         double* __restrict__ yn, RmGeom g) {
  int ip = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= g.mnew) return;
  int i = ip < g.p ? ip : ip + g.n;
  xn[ip] = xo[i];
  yn[ip] = yo[i];
}

// alpha of the result without another pass over the factor: W_{B,j} = sum_k M[k,B] M[k,j] is the LAST prefix of the
// deletion's row sums (H + P of the last row tile), so alpha'_j = alpha_j - W_{B,j} . (W_BB^-1 alpha_B).
__global__ void k_rm_alpha(const double* __restrict__ alpha_old, double* __restrict__ alpha_new, const double* __restrict__ H,
                           const double* __restrict__ P, RmGeom g, const double* __restrict__ gsol) {
  const int jp = blockIdx.x * blockDim.x + threadIdx.x;
  if (jp >= g.mnew) return;
  const int j = jp < g.p ? jp : jp + g.n;
  const int64_t base = ((int64_t)(g.nbn - 1 - g.I0) * g.nbn + (jp >> 6)) * (NP * TB) + (jp & 63);
  double s = 0.0;
  for (int a = 0; a < g.n; ++a) s += (H[base + (int64_t)a * TB] + P[base + (int64_t)a * TB]) * gsol[a];
  alpha_new[jp] = alpha_old[j] - s;
}

// ---------------------------------------------------------------------------------------------
// Identity padding for block rows/cols [from, to) of M (to, from multiples of 64).
__global__ void k_pad_identity(double* __restrict__ M, int64_t ld, int from, int to) {
  const int64_t total = (int64_t)to * to - (int64_t)from * from;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int i, j;
    const int64_t top = (int64_t)from * (to - from);  // rows < from, cols [from, to)
    if (idx < top) {
      i = (int)(idx / (to - from));
      j = from + (int)(idx % (to - from));
    } else {
      int64_t q = idx - top;
      i = from + (int)(q / to);
      j = (int)(q % to);
    }
    M[(int64_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
}

// Workspace of an append: Li [NP*NP] | znew [NP] | gpart [nblk][NG] | part [KS][NP][mpad]
constexpr int APPEND_KS = 8;
constexpr int AP_GRAM_ROWS = 256;
struct ApWork {
  double *Li, *znew, *gpart, *part;
};

// Partial Gram matrices of T (n columns x mpad rows, T[c*tstride + r]): gpart[blk][(a,b)] over 256 rows.
__global__ void __launch_bounds__(256) k_append_gram(const double* __restrict__ T, int mpad, int64_t tstride, int n, ApWork w) {
  __shared__ double sT[NP][AP_GRAM_ROWS + 1];
  const int r0 = blockIdx.x * AP_GRAM_ROWS;
  for (int idx = threadIdx.x; idx < NP * AP_GRAM_ROWS; idx += 256) {
    int a = idx / AP_GRAM_ROWS, r = idx % AP_GRAM_ROWS;
    sT[a][r] = (a < n && r0 + r < mpad) ? T[(int64_t)a * tstride + r0 + r] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < NG) {
    int a, b;
    tri_unpack(threadIdx.x, a, b);
    double s = 0.0;
    if (a < n) {
#pragma unroll 8
      for (int r = 0; r < AP_GRAM_ROWS; ++r) s += sT[a][r] * sT[b][r];
    }
    w.gpart[(int64_t)blockIdx.x * NG + threadIdx.x] = s;
  }
}

// One CTA: S = Kss + (noise + 1e-8) I - T^T T, Ls = chol(S), Li = Ls^-1, znew = Li (y_B - mu_B).
//   mu[c] = predictive mean at x_B[c] (mean function included).
__global__ void __launch_bounds__(256) k_append_small(const double* __restrict__ mu, int nblk, const double* __restrict__ xB,
                                                      const double* __restrict__ yB, int n, Theta th, ApWork w,
                                                      int* __restrict__ info) {
  __shared__ double S[NP][NP + 1], Li[NP][NP + 1];
  if (threadIdx.x < NG) {
    int a, b;
    tri_unpack(threadIdx.x, a, b);
    if (a < n) {
      double s = 0.0;
#pragma unroll 8
      for (int q = 0; q < nblk; ++q) s += w.gpart[(int64_t)q * NG + threadIdx.x];
      double k = kern_eval(th, xB[a], xB[b], a == b, nullptr);
      if (a == b) k = __dadd_rn(k, __dadd_rn(th.noise, GPY_JITTER));
      S[a][b] = k - s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool ok = true;
    for (int k = 0; k < n && ok; ++k) {
      double d = S[k][k];
      for (int q = 0; q < k; ++q) d -= S[k][q] * S[k][q];
      if (!(d > 0.0)) { ok = false; break; }
      d = sqrt(d);
      S[k][k] = d;
      for (int i = k + 1; i < n; ++i) {
        double s = S[i][k];
        for (int q = 0; q < k; ++q) s -= S[i][q] * S[k][q];
        S[i][k] = s / d;
      }
    }
    if (!ok) atomicCAS(info, 0, 1);
    // Li = Ls^-1 (lower), column by column
    for (int c = 0; c < n; ++c)
      for (int i = 0; i < n; ++i) {
        double s = (i == c) ? 1.0 : 0.0;
        if (i < c) { Li[i][c] = 0.0; continue; }
        for (int q = c; q < i; ++q) s -= S[i][q] * Li[q][c];
        Li[i][c] = ok ? s / S[i][i] : nan("");
      }
  }
  __syncthreads();
  if (threadIdx.x < NP * NP) {
    int i = threadIdx.x / NP, c = threadIdx.x % NP;
    w.Li[threadIdx.x] = (i < n && c < n) ? Li[i][c] : 0.0;
  }
  if (threadIdx.x < NP) {
    double s = 0.0;
    if ((int)threadIdx.x < n)
      for (int q = 0; q <= (int)threadIdx.x; ++q) s += Li[threadIdx.x][q] * (yB[q] - mu[q]);
    w.znew[threadIdx.x] = s;
  }
}

// part[s][i][j] = sum over row blocks kb = J + s, J + s + KS, ... of (Li T^T)[i][kb-block] * M[kb-block][J-block].
// The tiles of M (one read of the factor in all) and the matching slices of T stream through a two-stage cp.async
// ring: the next row block is in flight while the current one is multiplied.
constexpr int AP_STAGE = TB * LD_R + NP * LD_C;  // doubles per stage: M tile [k][c] | raw T slice [q][k]
constexpr int AP_ROWS_SMEM = (2 * AP_STAGE + NP * LD_C) * (int)sizeof(double);
__global__ void __launch_bounds__(NTHREADS) k_append_rows(const double* __restrict__ M, int64_t ld, int nb, int mpad,
                                                          const double* __restrict__ T, int64_t tstride, int n, ApWork w) {
  extern __shared__ __align__(16) double smem[];
  double* sC = smem + 2 * AP_STAGE;  // (Li T^T) tile [i][k], LD_C
  __shared__ double sLi[NP][NP + 1];
  const int J = blockIdx.x, s = blockIdx.y;
  if (threadIdx.x < NP * NP) sLi[threadIdx.x / NP][threadIdx.x % NP] = w.Li[threadIdx.x];
  Acc<2, 1> acc;
  acc.zero();
  const int warp = threadIdx.x >> 5;
  auto issue = [&](int kb, int stage) {
    double* sT = smem + stage * AP_STAGE;
    double* sR = sT + TB * LD_R;
    load_tile_async<LD_R>(sT, M + (int64_t)kb * TB * ld + (int64_t)J * TB, ld);
    for (int idx = threadIdx.x; idx < NP * TB / 2; idx += NTHREADS) {
      const int q = idx >> 5, k = (idx & 31) << 1;
      if (q < n) cp_async16(sR + q * LD_C + k, T + (int64_t)q * tstride + kb * TB + k);
    }
  };
  int stage = 0;
  if (J + s < nb) issue(J + s, 0);
  cp_async_commit();
  for (int kb = J + s; kb < nb; kb += APPEND_KS, stage ^= 1) {
    cp_async_wait<0>();
    __syncthreads();  // this row block has landed for everybody; the other stage and sC are free
    if (kb + APPEND_KS < nb) issue(kb + APPEND_KS, stage ^ 1);
    cp_async_commit();
    const double* sT = smem + stage * AP_STAGE;
    const double* sR = sT + TB * LD_R;
    for (int idx = threadIdx.x; idx < NP * TB; idx += NTHREADS) {
      int i = idx >> 6, k = idx & 63;
      double v = 0.0;
      for (int q = 0; q <= i && q < n; ++q) v += sLi[i][q] * sR[q * LD_C + k];
      sC[i * LD_C + k] = v;
    }
    __syncthreads();
    warp_mma<2, 1, true, false>(acc, sC, sT, 0, warp * 8);
  }
  cp_async_wait<0>();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int r = 8 * i + g, c = J * TB + warp * 8 + 2 * t;
    double* o = w.part + ((int64_t)s * NP + r) * mpad + c;
    o[0] = acc.v[i][0][0];
    o[1] = acc.v[i][0][1];
  }
}

// Write the n new rows of M, extend x, y, z.  grid: ceil(mpad_new / 256).
__global__ void k_append_finalize(double* __restrict__ M, int64_t ld, int m, int n, int mpad_old, int mpad_new, ApWork w,
                                  const double* __restrict__ xB, const double* __restrict__ yB, double* __restrict__ x,
                                  double* __restrict__ y, double* __restrict__ z, double* __restrict__ alpha) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= mpad_new) return;
  // alpha' = M'^T z' = [alpha + R^T z_B ; Ls^-T z_B] with R the new rows: no pass over the factor
  double da = 0.0;
  for (int i = 0; i < n; ++i) {
    double v;
    if (j < m) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < APPEND_KS; ++q) s += w.part[((int64_t)q * NP + i) * mpad_old + j];
      v = -s;
    } else if (j < m + n) {
      v = w.Li[i * NP + (j - m)];
    } else {
      v = 0.0;
    }
    M[(int64_t)(m + i) * ld + j] = v;
    da += v * w.znew[i];
  }
  if (j < m + n) alpha[j] = (j < m ? alpha[j] : 0.0) + da;
  if (j < n) {
    x[m + j] = xB[j];
    y[m + j] = yB[j];
    z[m + j] = w.znew[j];
  }
}

inline void update_set_attributes() {
  cudaFuncSetAttribute(k_rm_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_PREP_SMEM);
  cudaFuncSetAttribute(k_rm_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_APPLY_SMEM);
  cudaFuncSetAttribute(k_rm_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_STRIP_SMEM);
  cudaFuncSetAttribute(k_rm_strip64, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_STRIP64_SMEM);
  cudaFuncSetAttribute(k_T_leave, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_STRIP_SMEM);
  cudaFuncSetAttribute(k_append_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_ROWS_SMEM);
}

}  // namespace mogp
