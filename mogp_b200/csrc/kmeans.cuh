// kmeans.cuh — the sampler's initial partition on the device (mogp/mogp_constrained.py:182-207):
//   KMeans(n_clusters=K, random_state=rand_seed, n_init=10).fit(X[:, 1]).labels_
// for the ONE-dimensional data the reference clusters (each patient's first-visit time).  The arithmetic follows
// scikit-learn's dense float64 path so that the labels - not just the partition - come out the same:
//   * data centred by its mean; tol = var(x) * 1e-4;
//   * k-means++ seeding (sklearn.cluster._kmeans._kmeans_plusplus): first centre by choice(n) (host, NumPy's cdf
//     arithmetic), every further centre the best of 2 + int(ln K) candidates drawn with probability proportional to the
//     squared distance to the closest centre so far: searchsorted(cumsum(closest), u * potential), squared distances as
//     sklearn's euclidean_distances evaluates them (-2 c x + c^2 + x^2, clipped at 0);
//   * Lloyd iterations: labels = argmin_j (c_j^2 - 2 x c_j) (first minimum), centres = means, stop when the labels
//     repeat or the squared centre shift <= tol, then one more E-step; inertia; the best of n_init runs wins.
// The uniforms sklearn would take from RandomState(random_state) come in from the host, as every other draw of the
// chain does (they do not depend on the data: n_init * (1 + (K - 1) * trials) doubles).
// One CTA per initialisation (the n_init runs are independent once the draws are fixed), every reduction in a fixed
// order: the result is deterministic.  Summation orders differ from NumPy's (pairwise) and BLAS's, so a candidate /
// label could differ from sklearn's only if a draw or a point sat within rounding of a boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mogp {

constexpr int KM_THREADS = 1024;
__host__ __device__ __forceinline__ int64_t km_min(int64_t a, int64_t b) { return a < b ? a : b; }

// sum over the CTA, fixed order: per-thread partials -> warp shuffle tree -> warps in order.  All threads get it.
__device__ __forceinline__ double km_block_sum(double v, double* s_warp /* [32] */) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // s_warp may still be read from the previous call
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll 8
  for (int w = 0; w < KM_THREADS / 32; ++w) s += s_warp[w];
  return s;
}

__device__ __forceinline__ double km_d2(double c, double x) {  // sklearn euclidean_distances, one feature
  double t = __dmul_rn(-2.0, __dmul_rn(c, x));
  t = __dadd_rn(t, __dmul_rn(c, c));
  t = __dadd_rn(t, __dmul_rn(x, x));
  return t > 0.0 ? t : 0.0;
}

// mean and variance of x (one CTA), xc = x - mean
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_center(const double* __restrict__ x, int64_t n, double* __restrict__ xc,
                                                              double* __restrict__ stat /* mean, var */) {
  __shared__ double s_warp[32];
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += KM_THREADS) a += x[i];
  const double mean = km_block_sum(a, s_warp) / (double)n;
  double v = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += KM_THREADS) {
    const double d = x[i] - mean;
    xc[i] = d;
    v += d * d;
  }
  const double var = km_block_sum(v, s_warp) / (double)n;
  if (threadIdx.x == 0) {
    stat[0] = mean;
    stat[1] = var;
  }
}

// k-means++ seeding, one CTA per initialisation.  u: the run's uniforms AFTER the first-centre draw.
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_seed(const double* __restrict__ xc, int64_t n, int K, int trials,
                                                            const double* __restrict__ u_all, int64_t per_init,
                                                            const int64_t* __restrict__ first_id, double* __restrict__ closest_all,
                                                            double* __restrict__ centers_all) {
  __shared__ double s_warp[32];
  __shared__ double s_pref[KM_THREADS + 1];
  __shared__ int64_t s_cand[16];
  __shared__ double s_pot[16];
  const int run = blockIdx.x, tid = threadIdx.x;
  const double* u = u_all + (int64_t)run * per_init + 1;
  double* closest = closest_all + (int64_t)run * n;
  double* centers = centers_all + (int64_t)run * K;
  const int64_t L = (n + KM_THREADS - 1) / KM_THREADS, a0 = km_min(n, (int64_t)tid * L), a1 = km_min(n, a0 + L);
  const double c0 = xc[first_id[run]];
  double part = 0.0;
  for (int64_t i = tid; i < n; i += KM_THREADS) {
    const double d = km_d2(c0, xc[i]);
    closest[i] = d;
    part += d;
  }
  double pot = km_block_sum(part, s_warp);
  if (tid == 0) centers[0] = c0;
  for (int c = 1; c < K; ++c) {
    // cumsum(closest): contiguous chunk per thread, exclusive scan of the chunk sums (fixed order)
    double loc = 0.0;
    for (int64_t i = a0; i < a1; ++i) loc += closest[i];
    __syncthreads();
    s_pref[tid + 1] = loc;
    if (tid == 0) s_pref[0] = 0.0;
    __syncthreads();
    if (tid == 0) {
      double run_sum = 0.0;
      for (int t = 1; t <= KM_THREADS; ++t) {  // 1024 sequential adds: the order NumPy's cumsum uses across chunk ends
        run_sum += s_pref[t];
        s_pref[t] = run_sum;
      }
    }
    if (tid < trials) s_cand[tid] = n - 1;  // (np.clip: a draw beyond the last cumulative sum)
    __syncthreads();
    for (int t = 0; t < trials; ++t) {
      const double rv = u[(int64_t)(c - 1) * trials + t] * pot;
      const double lo = s_pref[tid], hi = s_pref[tid + 1];
      if (a1 > a0 && (tid == 0 || lo < rv) && rv <= hi) {  // the first index whose cumulative sum reaches rv is in this chunk
        double acc = lo;
        int64_t idx = a1 - 1;
        for (int64_t i = a0; i < a1; ++i) {
          acc += closest[i];
          if (acc >= rv) {
            idx = i;
            break;
          }
        }
        s_cand[t] = idx;
      }
    }
    __syncthreads();
    for (int t = 0; t < trials; ++t) {
      const double cc = xc[s_cand[t]];
      double p = 0.0;
      for (int64_t i = tid; i < n; i += KM_THREADS) p += fmin(closest[i], km_d2(cc, xc[i]));
      const double tot = km_block_sum(p, s_warp);
      if (tid == 0) s_pot[t] = tot;
    }
    __syncthreads();
    int best = 0;
    for (int t = 1; t < trials; ++t)
      if (s_pot[t] < s_pot[best]) best = t;  // np.argmin: the first minimum
    pot = s_pot[best];
    const double cb = xc[s_cand[best]];
    for (int64_t i = tid; i < n; i += KM_THREADS) closest[i] = fmin(closest[i], km_d2(cb, xc[i]));
    if (tid == 0) centers[c] = cb;
    __syncthreads();
  }
}

// Lloyd iterations of one initialisation per CTA.  Dynamic shared memory: centers[K] | newc[K] | csq[K] | part[S * K] |
// cnt[S * K] (S = segments per cluster of the M-step).
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_lloyd(const double* __restrict__ xc, int64_t n, int K, int max_iter,
                                                             const double* __restrict__ stat, double tol_rel,
                                                             double* __restrict__ centers_all, int32_t* __restrict__ labels_all,
                                                             int32_t* __restrict__ labels_old_all, double* __restrict__ inertia_all,
                                                             int32_t* __restrict__ iters_all) {
  extern __shared__ __align__(16) double km_smem[];
  __shared__ double s_warp[32];
  __shared__ int s_flag, s_empty;
  const int run = blockIdx.x, tid = threadIdx.x;
  const int S = K >= KM_THREADS ? 1 : KM_THREADS / K;
  double* cen = km_smem;
  double* newc = cen + K;
  double* csq = newc + K;
  double* part = csq + K;
  double* cntp = part + (int64_t)S * K;
  double* centers = centers_all + (int64_t)run * K;
  int32_t* labels = labels_all + (int64_t)run * n;
  int32_t* labels_old = labels_old_all + (int64_t)run * n;
  const double tol = stat[1] * tol_rel;
  for (int j = tid; j < K; j += KM_THREADS) cen[j] = centers[j];
  for (int64_t i = tid; i < n; i += KM_THREADS) labels_old[i] = -1;
  __syncthreads();
  auto e_step = [&]() -> int {  // labels = argmin_j c_j^2 - 2 x c_j; returns whether any label differs from labels_old
    for (int j = tid; j < K; j += KM_THREADS) csq[j] = cen[j] * cen[j];
    __syncthreads();
    int changed = 0;
    for (int64_t i = tid; i < n; i += KM_THREADS) {
      const double x = xc[i];
      double bestv = csq[0] - 2.0 * (x * cen[0]);
      int bj = 0;
      for (int j = 1; j < K; ++j) {
        const double v = csq[j] - 2.0 * (x * cen[j]);
        if (v < bestv) {
          bestv = v;
          bj = j;
        }
      }
      labels[i] = bj;
      changed |= (bj != labels_old[i]);
    }
    if (tid == 0) s_flag = 0;
    __syncthreads();
    if (changed) s_flag = 1;
    __syncthreads();
    return s_flag;
  };
  int it = 0;
  bool strict = false;
  for (; it < max_iter; ++it) {
    const int changed = e_step();
    // M-step: cluster j, segment s sums its share of the points in index order; segments added in order
    const int64_t seg_len = (n + S - 1) / S;
    for (int q = tid; q < S * K; q += KM_THREADS) {
      const int j = q % K, s = q / K;
      const int64_t b0 = km_min(n, (int64_t)s * seg_len), b1 = km_min(n, b0 + seg_len);
      double sum = 0.0, cnt = 0.0;
      for (int64_t i = b0; i < b1; ++i)
        if (labels[i] == j) {
          sum += xc[i];
          cnt += 1.0;
        }
      part[(int64_t)s * K + j] = sum;
      cntp[(int64_t)s * K + j] = cnt;
    }
    if (tid == 0) s_empty = 0;
    __syncthreads();
    for (int j = tid; j < K; j += KM_THREADS) {
      double sum = 0.0, cnt = 0.0;
      for (int s = 0; s < S; ++s) {
        sum += part[(int64_t)s * K + j];
        cnt += cntp[(int64_t)s * K + j];
      }
      part[j] = sum;   // (segment 0's slots now hold the totals)
      cntp[j] = cnt;
      if (cnt == 0.0) s_empty = 1;
    }
    __syncthreads();
    if (s_empty && tid == 0) {
      // sklearn _relocate_empty_clusters_dense: the points farthest from their centres become the centres of the empty
      // clusters (one each, taken out of the cluster they were in).  Rare; done by one thread.
      for (int j = 0; j < K; ++j) {
        if (cntp[j] != 0.0) continue;
        double far = -1.0;
        int64_t fi = -1;
        for (int64_t i = 0; i < n; ++i) {
          const int l = labels[i];
          if (cntp[l] <= 1.0) continue;  // never empties another cluster
          const double d = (xc[i] - cen[l]) * (xc[i] - cen[l]);
          if (d > far) {
            far = d;
            fi = i;
          }
        }
        if (fi < 0) break;
        const int l = labels[fi];
        part[l] -= xc[fi];
        cntp[l] -= 1.0;
        part[j] = xc[fi];
        cntp[j] = 1.0;
        labels[fi] = j;
      }
    }
    __syncthreads();
    double sh = 0.0;
    for (int j = tid; j < K; j += KM_THREADS) {
      const double nc = cntp[j] > 0.0 ? part[j] / cntp[j] : cen[j];
      newc[j] = nc;
      sh += (nc - cen[j]) * (nc - cen[j]);
    }
    const double shift = km_block_sum(sh, s_warp);
    for (int j = tid; j < K; j += KM_THREADS) cen[j] = newc[j];
    __syncthreads();
    if (!changed) {  // np.array_equal(labels, labels_old): strict convergence
      strict = true;
      ++it;
      break;
    }
    if (shift <= tol) {
      ++it;
      break;
    }
    for (int64_t i = tid; i < n; i += KM_THREADS) labels_old[i] = labels[i];
    __syncthreads();
  }
  if (!strict) e_step();  // labels that match the final centres
  double in = 0.0;
  for (int64_t i = tid; i < n; i += KM_THREADS) {
    const double d = xc[i] - cen[labels[i]];
    in += d * d;
  }
  const double inertia = km_block_sum(in, s_warp);
  for (int j = tid; j < K; j += KM_THREADS) centers[j] = cen[j];
  if (tid == 0) {
    inertia_all[run] = inertia;
    iters_all[run] = it;
  }
}

}  // namespace mogp
