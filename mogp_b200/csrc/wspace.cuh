// wspace.cuh - the cluster's log marginal likelihood and its gradient in WEIGHT SPACE, for the optimiser's evaluations of
// the hyper-parameter refit (refit.inl).  The reference evaluates GPRegression's objective by a dense Cholesky of the
// m x m covariance at every point L-BFGS-B asks for (mogp/obsmodel.py:94 -> GPy ExactGaussianInference): O(m^3), ~370 times
// per sweep at m ~ 2 600.  On ONE input dimension the RBF kernel is numerically low rank on the cluster's interval:
// interpolated at r Chebyshev nodes t,   k(x, x') ~= c(x)^T K_tt c(x'),   c = the Lagrange basis of the nodes, so
//     Ky = C K_tt C^T + s I        (C = [c(x_1) .. c(x_m)]^T, m x r;  s = noise + 1e-8).
// Once per refit (theta-independent): S = C^T C, C^T y, C^T x, y.y, x.y, x.x.  Per evaluation: K_tt ~= R^T R by a pivoted
// Cholesky stopped at 1e-15 of the largest pivot (R: rho x r, rho ~ 17..35 - the kernel's numerical rank), Phi = C R^T,
//     A = s I + R S R^T  (rho x rho, eigenvalues >= s whatever the data),   b = R C^T res,   v = A^-1 b,   res = y + a x,
//     log|Ky|         = (m - rho) log s + log|A|
//     res^T Ky^-1 res = (res.res - b.v) / s
//     dLML/dtheta_K   = 1/2 [ u^T dK u - sum(dK o W) ],   u = C^T Ky^-1 res = (C^T res - (R S)^T v) / s,
//                                                          W = C^T Ky^-1 C   = (S - (R S)^T A^-1 (R S)) / s
//     dLML/ds         = 1/2 [ |Ky^-1 res|^2 - tr Ky^-1 ] = 1/2 [ (res.res - b.v - s v.v)/s^2 - (m - rho + s tr A^-1)/s ]
//     dLML/da         = -(x.res - (R C^T x).v) / s                                              (mean function -a x)
// - small symmetric positive definite algebra, O(r^3) per evaluation whatever m, and no inverse of anything the data could
// make singular.  Against the dense restatement (oracle/weight_space.py, tests/test_weight_space.py): LML to 1e-12
// relative, gradient to 1e-11 of its norm while the interval is at most 9 length scales long (r = 48); shorter length
// scales are evaluated the dense way.  The factor the scoring path needs is produced by the closing dense evaluation at
// x_opt, as before.
//
// The arithmetic is written as phases of a bulk-synchronous CTA (every phase: independent work items; barriers between
// phases), so that the SAME functions run under sequential loops on the host (tests/native/ws_host.cpp).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define WS_HD __host__ __device__ __forceinline__
#else
#define WS_HD inline
#endif

namespace mogp {

constexpr int WS_R = 48;                 // Chebyshev nodes per cluster
constexpr double WS_MAX_SPAN = 9.0;      // (hi - lo) / lengthscale up to which r = 48 reproduces the dense result to 1e-12
constexpr double WS_PIVOT_TOL = 1e-15;   // the pivoted Cholesky of K_tt stops below this fraction of the largest pivot
constexpr int WS_RR = WS_R * WS_R;

// One lane's workspace (doubles).
struct WsWork {
  double* base;
  WS_HD double* t() const { return base; }                          // [r] nodes
  WS_HD double* S() const { return base + WS_R; }                   // [r][r] C^T C
  WS_HD double* cy() const { return S() + WS_RR; }                  // [r] C^T y
  WS_HD double* cx() const { return cy() + WS_R; }                  // [r] C^T x
  WS_HD double* sc() const { return cx() + WS_R; }                  // [8] y.y, x.y, x.x, -, m ; [5..7] pivot scratch: p, pivot, continue
  WS_HD double* K() const { return sc() + 8; }                      // [r][r] K_tt
  WS_HD double* E() const { return K() + WS_RR; }                   // dK/dvariance
  WS_HD double* Q() const { return E() + WS_RR; }                   // dK/dlengthscale
  WS_HD double* R() const { return Q() + WS_RR; }                   // [rho][r] K_tt ~= R^T R
  WS_HD double* RS() const { return R() + WS_RR; }                  // [rho][r] R S
  WS_HD double* A() const { return RS() + WS_RR; }                  // [rho][rho] (row stride r) s I + R S R^T, then its Cholesky factor
  WS_HD double* X() const { return A() + WS_RR; }                   // [rho][r] A^-1 R S
  WS_HD double* W() const { return X() + WS_RR; }                   // [r][r] C^T Ky^-1 C
  WS_HD double* dg() const { return W() + WS_RR; }                  // [r] residual diagonal of the pivoted Cholesky
  WS_HD double* cres() const { return dg() + WS_R; }                // [r] C^T res
  WS_HD double* b() const { return cres() + WS_R; }                 // [rho]
  WS_HD double* v() const { return b() + WS_R; }                    // [rho] A^-1 b
  WS_HD double* out() const { return v() + WS_R; }                  // [8] lml, dA, dvariance, dlengthscale, dnoise, info, rho
};
constexpr int64_t WS_WORK_ELEMS = WS_R + WS_RR + 2 * WS_R + 8 + 8 * (int64_t)WS_RR + 4 * WS_R + 8;

struct WsNodes {  // Chebyshev points of the second kind on [lo, hi] and their barycentric weights (host-computed)
  double t[WS_R], w[WS_R];
};
inline void ws_make_nodes(double lo, double hi, WsNodes* nd) {
  for (int j = 0; j < WS_R; ++j) {
    nd->t[j] = 0.5 * (lo + hi) + 0.5 * (hi - lo) * cos(M_PI * j / (WS_R - 1));
    nd->w[j] = ((j & 1) ? -1.0 : 1.0) * ((j == 0 || j == WS_R - 1) ? 0.5 : 1.0);
  }
}

// ---- setup ---------------------------------------------------------------------------------------------------------
// row k of C: the Lagrange basis of the nodes at x_k (barycentric formula; exact at a node)
WS_HD void ws_basis_row(const WsNodes& nd, double x, double* c) {
  int hit = -1;
  double sum = 0.0;
  for (int j = 0; j < WS_R; ++j) {
    const double dd = x - nd.t[j];
    if (dd == 0.0) hit = j;
    c[j] = nd.w[j] / (dd == 0.0 ? 1.0 : dd);
    sum += c[j];
  }
  for (int j = 0; j < WS_R; ++j) c[j] = hit >= 0 ? (j == hit ? 1.0 : 0.0) : c[j] / sum;
}
// entry e of S = C^T C (e < r r), of C^T y / C^T x (next 2 r), of the three scalar sums (next 3): sums over the m rows
constexpr int WS_GRAM_ENTRIES = WS_RR + 2 * WS_R + 3;
WS_HD void ws_gram_entry(const double* C, const double* x, const double* y, int m, int e, const WsWork& w) {
  double acc = 0.0;
  if (e < WS_RR) {
    const int i = e / WS_R, j = e % WS_R;
    for (int k = 0; k < m; ++k) acc += C[(int64_t)k * WS_R + i] * C[(int64_t)k * WS_R + j];
    w.S()[e] = acc;
  } else if (e < WS_RR + WS_R) {
    const int i = e - WS_RR;
    for (int k = 0; k < m; ++k) acc += C[(int64_t)k * WS_R + i] * y[k];
    w.cy()[i] = acc;
  } else if (e < WS_RR + 2 * WS_R) {
    const int i = e - WS_RR - WS_R;
    for (int k = 0; k < m; ++k) acc += C[(int64_t)k * WS_R + i] * x[k];
    w.cx()[i] = acc;
  } else {
    const int q = e - WS_RR - 2 * WS_R;
    for (int k = 0; k < m; ++k) acc += (q == 0 ? y[k] * y[k] : q == 1 ? x[k] * y[k] : x[k] * x[k]);
    w.sc()[q] = acc;
    if (q == 0) w.sc()[4] = (double)m;
  }
}

// ---- one evaluation at theta = (a, variance, lengthscale, noise) ------------------------------------------------
struct WsTheta {
  double a, var, ls, noise;
  int mean_func;
};
// phase (r r items): the node kernel and its derivatives; (r items, e < r): the residual diagonal and C^T res
WS_HD void ws_ph_kernel(const WsWork& w, const WsTheta& th, int e) {
  const int i = e / WS_R, j = e % WS_R;
  const double dd = w.t()[i] - w.t()[j], r2 = dd * dd / (th.ls * th.ls);
  const double ex = exp(-0.5 * r2);
  w.K()[e] = th.var * ex;
  w.E()[e] = ex;
  w.Q()[e] = th.var * ex * r2 / th.ls;
  w.R()[e] = 0.0;
  if (e < WS_R) {
    w.dg()[e] = th.var;
    w.cres()[e] = w.cy()[e] + (th.mean_func ? th.a * w.cx()[e] : 0.0);
  }
}
// pivoted Cholesky K ~= R^T R, step k.  (one item) the pivot: the largest residual diagonal, first index on ties
WS_HD void ws_pc_pivot(const WsWork& w, const WsTheta& th, int k) {
  int p = 0;
  for (int j = 1; j < WS_R; ++j)
    if (w.dg()[j] > w.dg()[p]) p = j;
  const double dp = w.dg()[p];
  const bool go = dp > WS_PIVOT_TOL * th.var;
  w.sc()[5] = (double)p;
  w.sc()[6] = go ? sqrt(dp) : 0.0;
  w.sc()[7] = go ? 1.0 : 0.0;
  if (!go) w.out()[6] = (double)k;  // rho
}
// (r items) row k of R and the residual diagonal
WS_HD void ws_pc_row(const WsWork& w, int k, int j) {
  const int p = (int)w.sc()[5];
  const double piv = w.sc()[6];
  double s = w.K()[p * WS_R + j];
  for (int q = 0; q < k; ++q) s -= w.R()[q * WS_R + p] * w.R()[q * WS_R + j];
  s /= piv;
  w.R()[k * WS_R + j] = s;
  w.dg()[j] = j == p ? 0.0 : w.dg()[j] - s * s;
}
// (rho r items) R S
WS_HD void ws_ph_rs(const WsWork& w, int e) {
  const int i = e / WS_R, j = e % WS_R;
  double a = 0.0;
  for (int k = 0; k < WS_R; ++k) a += w.R()[i * WS_R + k] * w.S()[k * WS_R + j];
  w.RS()[e] = a;
}
// (rho rho items) A = s I + R S R^T ; (j == 0 also: b = R C^T res)
WS_HD void ws_ph_a(const WsWork& w, const WsTheta& th, int i, int j) {
  double a = 0.0;
  for (int k = 0; k < WS_R; ++k) a += w.RS()[i * WS_R + k] * w.R()[j * WS_R + k];
  w.A()[i * WS_R + j] = a + (i == j ? th.noise + 1e-8 : 0.0);
  if (j == 0) {
    double bb = 0.0;
    for (int k = 0; k < WS_R; ++k) bb += w.R()[i * WS_R + k] * w.cres()[k];
    w.b()[i] = bb;
  }
}
// in-place lower Cholesky of the leading n x n block of a matrix with row stride WS_R; false on a non-positive pivot
WS_HD bool ws_chol(double* A, int n) {
  for (int k = 0; k < n; ++k) {
    double dgk = A[k * WS_R + k];
    for (int q = 0; q < k; ++q) dgk -= A[k * WS_R + q] * A[k * WS_R + q];
    if (!(dgk > 0.0)) return false;
    dgk = sqrt(dgk);
    A[k * WS_R + k] = dgk;
    for (int i = k + 1; i < n; ++i) {
      double s = A[i * WS_R + k];
      for (int q = 0; q < k; ++q) s -= A[i * WS_R + q] * A[k * WS_R + q];
      A[i * WS_R + k] = s / dgk;
    }
  }
  return true;
}
WS_HD void ws_solve(const double* Lf, int n, double* b, int stride) {  // b <- (L L^T)^-1 b, b[i * stride]
  for (int i = 0; i < n; ++i) {
    double s = b[i * stride];
    for (int q = 0; q < i; ++q) s -= Lf[i * WS_R + q] * b[q * stride];
    b[i * stride] = s / Lf[i * WS_R + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i * stride];
    for (int q = i + 1; q < n; ++q) s -= Lf[q * WS_R + i] * b[q * stride];
    b[i * stride] = s / Lf[i * WS_R + i];
  }
}
// (one item) Cholesky of A (in place in Af: the workspace's A, or a shared-memory copy of it), v = A^-1 b
WS_HD void ws_ph_chol(const WsWork& w, double* Af, int rho) {
  const bool ok = rho > 0 && ws_chol(Af, rho);
  w.out()[5] = ok ? 0.0 : 1.0;
  if (!ok) return;
  for (int i = 0; i < rho; ++i) w.v()[i] = w.b()[i];
  ws_solve(Af, rho, w.v(), 1);
}
// (r items) column j of X = A^-1 R S
WS_HD void ws_ph_x(const WsWork& w, const double* Af, int rho, int j) {
  for (int i = 0; i < rho; ++i) w.X()[i * WS_R + j] = w.RS()[i * WS_R + j];
  ws_solve(Af, rho, w.X() + j, WS_R);
}
// (r r items) W = (S - (R S)^T X) / s
WS_HD void ws_ph_w(const WsWork& w, const WsTheta& th, int rho, int e) {
  const int i = e / WS_R, j = e % WS_R;
  double a = 0.0;
  for (int k = 0; k < rho; ++k) a += w.RS()[k * WS_R + i] * w.X()[k * WS_R + j];
  w.W()[e] = (w.S()[e] - a) / (th.noise + 1e-8);
}
// (one item) the sums
WS_HD void ws_ph_finish(const WsWork& w, const WsTheta& th, const double* Af, int rho) {
  double* out = w.out();
  if (out[5] != 0.0) return;
  const double s = th.noise + 1e-8, m = w.sc()[4];
  const double a = th.mean_func ? th.a : 0.0;
  const double rr = w.sc()[0] + 2.0 * a * w.sc()[1] + a * a * w.sc()[2];
  const double xres = w.sc()[1] + a * w.sc()[2];
  double logdet = 0.0, bv = 0.0, vv = 0.0, trAi = 0.0, rxv = 0.0;
  for (int i = 0; i < rho; ++i) {
    logdet += 2.0 * log(Af[i * WS_R + i]);
    bv += w.b()[i] * w.v()[i];
    vv += w.v()[i] * w.v()[i];
    double rx = 0.0;
    for (int k = 0; k < WS_R; ++k) rx += w.R()[i * WS_R + k] * w.cx()[k];
    rxv += rx * w.v()[i];
  }
  // tr A^-1 = sum_i |L^-1 e_i|^2
  for (int i = 0; i < rho; ++i) {
    double col[WS_R];
    for (int q = 0; q < rho; ++q) col[q] = q == i ? 1.0 : 0.0;
    for (int q = i; q < rho; ++q) {
      double sacc = col[q];
      for (int p = i; p < q; ++p) sacc -= Af[q * WS_R + p] * col[p];
      col[q] = sacc / Af[q * WS_R + q];
      trAi += col[q] * col[q];
    }
  }
  double u[WS_R];
  for (int j = 0; j < WS_R; ++j) {
    double acc = 0.0;
    for (int k = 0; k < rho; ++k) acc += w.RS()[k * WS_R + j] * w.v()[k];
    u[j] = (w.cres()[j] - acc) / s;
  }
  double qv = 0.0, ql = 0.0, tv = 0.0, tl = 0.0;
  for (int i = 0; i < WS_R; ++i) {
    double rv = 0.0, rl = 0.0;
    for (int j = 0; j < WS_R; ++j) {
      rv += w.E()[i * WS_R + j] * u[j];
      rl += w.Q()[i * WS_R + j] * u[j];
      tv += w.E()[i * WS_R + j] * w.W()[i * WS_R + j];
      tl += w.Q()[i * WS_R + j] * w.W()[i * WS_R + j];
    }
    qv += u[i] * rv;
    ql += u[i] * rl;
  }
  const double LOG2PI = 1.8378770664093453;
  out[0] = -0.5 * (m * LOG2PI + (m - rho) * log(s) + logdet + (rr - bv) / s);
  out[1] = th.mean_func ? -(xres - rxv) / s : 0.0;
  out[2] = 0.5 * (qv - tv);
  out[3] = 0.5 * (ql - tl);
  out[4] = 0.5 * ((rr - bv - s * vv) / (s * s) - (m - rho + s * trAi) / s);
}

#ifdef __CUDACC__
// C = Lagrange basis values, one thread per row
__global__ void __launch_bounds__(128) k_ws_basis(const __grid_constant__ WsNodes nd, const double* __restrict__ x, int m,
                                                  double* __restrict__ C, WsWork w) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < WS_R) w.t()[threadIdx.x] = nd.t[threadIdx.x];
  if (k >= m) return;
  double c[WS_R];
  ws_basis_row(nd, x[k], c);
  for (int j = 0; j < WS_R; ++j) C[(int64_t)k * WS_R + j] = c[j];
}
// S = C^T C, C^T y, C^T x, the scalar sums: one entry per thread, fixed summation order
__global__ void __launch_bounds__(128) k_ws_gram(const double* __restrict__ C, const double* __restrict__ x,
                                                 const double* __restrict__ y, int m, WsWork w) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < WS_GRAM_ENTRIES) ws_gram_entry(C, x, y, m, e, w);
}
// one evaluation: one CTA, the phases above between barriers
__global__ void __launch_bounds__(256) k_ws_eval(WsWork w, const WsTheta th) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double sA[WS_RR];  // the sequential factorisation works on a shared-memory copy (latency of every dependent load)
  if (tid == 0) w.out()[6] = (double)WS_R;
  for (int e = tid; e < WS_RR; e += nt) ws_ph_kernel(w, th, e);
  __syncthreads();
  for (int k = 0; k < WS_R; ++k) {
    if (tid == 0) ws_pc_pivot(w, th, k);
    __syncthreads();
    if (w.sc()[7] == 0.0) break;  // (uniform: everybody reads the same flag after the barrier)
    for (int j = tid; j < WS_R; j += nt) ws_pc_row(w, k, j);
    __syncthreads();
  }
  __syncthreads();
  const int rho = (int)w.out()[6];
  for (int e = tid; e < rho * WS_R; e += nt) ws_ph_rs(w, e);
  __syncthreads();
  for (int e = tid; e < rho * rho; e += nt) ws_ph_a(w, th, e / rho, e % rho);
  __syncthreads();
  for (int e = tid; e < rho * rho; e += nt) sA[(e / rho) * WS_R + e % rho] = w.A()[(e / rho) * WS_R + e % rho];
  __syncthreads();
  if (tid == 0) ws_ph_chol(w, sA, rho);
  __syncthreads();
  if (w.out()[5] != 0.0) return;
  for (int j = tid; j < WS_R; j += nt) ws_ph_x(w, sA, rho, j);
  __syncthreads();
  for (int e = tid; e < WS_RR; e += nt) ws_ph_w(w, th, rho, e);
  __syncthreads();
  if (tid == 0) ws_ph_finish(w, th, sA, rho);
}
#endif

}  // namespace mogp
