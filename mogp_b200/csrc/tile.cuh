// tile.cuh — FP64 building blocks shared by every kernel of the engine (sm_100a).
//
//  * dmma884: the native Blackwell FP64 tensor-core op (SASS DMMA.8x8x4) through
//    mma.sync.aligned.m8n8k4 (tcgen05/UMMA has no FP64 type).
//  * 64x64 tiles staged in shared memory with two paddings chosen so that the DMMA
//    fragment loads are bank-conflict free:
//        LD_C = 68 when the contraction index is the COLUMN of the stored tile,
//        LD_R = 68 when the contraction index is the ROW of the stored tile.
//  * in-shared-memory Cholesky and triangular inverse of one 64x64 diagonal tile.
//
// Matrices are row-major, lower triangular, leading dimension `ld` (multiple of 64);
// upper triangles of diagonal tiles are kept at zero so no kernel needs a mask.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mogp {

constexpr int TB = 64;     // tile edge
constexpr int LD_C = 68;   // smem ld, contraction over the column index of the stored tile
constexpr int LD_R = 68;   // smem ld, contraction over the row index of the stored tile (4t+g banks: conflict free)
constexpr int NTHREADS = 256;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Copy a ROWS x 64 tile (row-major, leading dimension ld) from global to shared memory [ROWS][LD].
template <int LD, int ROWS = TB>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g, int64_t ld) {
#pragma unroll
  for (int q = 0; q < (ROWS * TB / 2) / NTHREADS; ++q) {
    int idx = threadIdx.x + q * NTHREADS;
    int r = idx >> 5, c = (idx & 31) << 1;
    double2 v = *reinterpret_cast<const double2*>(g + (int64_t)r * ld + c);
    *reinterpret_cast<double2*>(s + r * LD + c) = v;
  }
}

// Accumulators of one warp for a (8*MS) x (8*NS) patch.
template <int MS, int NS>
struct Acc {
  double v[MS][NS][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) v[i][j][0] = v[i][j][1] = 0.0;
  }
};

// acc += A(rows rb.., k 0..KDIM) * B(k, cols cb..) for one warp.
//   A_COL: A element (r,k) is sA[r*LDA + k]   else sA[k*LDA + r]
//   B_COL: B element (k,n) is sB[n*LDB + k]   else sB[k*LDB + n]
// (LDA / LDB must be = 4 mod 16 for conflict-free fragment loads: 68, 84, ...)
template <int MS, int NS, bool A_COL, bool B_COL, int KDIM = TB, int LDA = 68, int LDB = 68>
__device__ __forceinline__ void warp_mma(Acc<MS, NS>& acc, const double* __restrict__ sA,
                                         const double* __restrict__ sB, int rb, int cb) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 4
  for (int kb = 0; kb < KDIM; kb += 4) {
    double a[MS], b[NS];
#pragma unroll
    for (int i = 0; i < MS; ++i)
      a[i] = A_COL ? sA[(rb + 8 * i + g) * LDA + kb + t] : sA[(kb + t) * LDA + rb + 8 * i + g];
#pragma unroll
    for (int j = 0; j < NS; ++j)
      b[j] = B_COL ? sB[(cb + 8 * j + g) * LDB + kb + t] : sB[(kb + t) * LDB + cb + 8 * j + g];
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
  }
}

// The standard 64x64 CTA tile: 8 warps as 4 (rows) x 2 (cols), each warp 16 x 32.
using Acc64 = Acc<2, 4>;
__device__ __forceinline__ int warp_rb64() { return ((threadIdx.x >> 5) >> 1) * 16; }
__device__ __forceinline__ int warp_cb64() { return ((threadIdx.x >> 5) & 1) * 32; }

// Visit every accumulator element of the 64x64 CTA tile: f(row, col, value&).
template <typename F>
__device__ __forceinline__ void for_each_acc64(Acc64& acc, F f) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rb = warp_rb64(), cb = warp_cb64();
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f(rb + 8 * i + g, cb + 8 * j + 2 * t, acc.v[i][j][0]);
      f(rb + 8 * i + g, cb + 8 * j + 2 * t + 1, acc.v[i][j][1]);
    }
}

// In-place lower Cholesky of the 64x64 tile s[64][LD] (lower part valid on entry).
// Returns 0 or (1 + first non-positive pivot column); result uniform across the CTA.
// The strict upper triangle is zeroed.
template <int LD>
__device__ int potrf_tile(double* __restrict__ s, double* __restrict__ col) {
  const int tid = threadIdx.x;
  int fail = 0;
  for (int k = 0; k < TB; ++k) {
    __syncthreads();
    double dkk = s[k * LD + k];
    if (!(dkk > 0.0)) {
      if (fail == 0) fail = k + 1;
      dkk = 1.0;  // keep going with a harmless pivot; the caller discards the result
    }
    double d = sqrt(dkk);
    if (tid < TB && tid >= k) {
      double c = (tid == k) ? d : s[tid * LD + k] / d;
      col[tid] = c;
    }
    __syncthreads();
    if (tid < TB && tid >= k) s[tid * LD + k] = col[tid];
    // trailing update: thread owns the 4x4 patch (i = 4*ti + a, j = 4*tj + b)
    const int ti = tid >> 4, tj = tid & 15;
    if (4 * ti + 3 > k && tj <= ti) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        int i = 4 * ti + a;
        double ci = col[i];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          int j = 4 * tj + b;
          if (j > k && j <= i) s[i * LD + j] -= ci * col[j];
        }
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
    int i = idx >> 6, j = idx & 63;
    if (j > i) s[i * LD + j] = 0.0;
  }
  __syncthreads();
  return fail;
}

// sM = inverse of the lower-triangular tile sL (both [64][LD]); upper triangle of sM zeroed.
// Column c is solved by the four lanes 4c..4c+3 (forward substitution, split dot products).
template <int LD>
__device__ void trtri_tile(const double* __restrict__ sL, double* __restrict__ sM) {
  const int tid = threadIdx.x, c = tid >> 2, q = tid & 3;
  for (int i = 0; i < TB; ++i) {
    // warp-uniform trip count is not required: lanes of a quad stay together (same c)
    double part = 0.0;
    if (i > c) {
      for (int j = c + q; j < i; j += 4) part += sL[i * LD + j] * sM[j * LD + c];
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (q == 0) {
      double v;
      if (i < c) v = 0.0;
      else if (i == c) v = 1.0 / sL[i * LD + i];
      else v = -part / sL[i * LD + i];
      sM[i * LD + c] = v;
    }
    __syncwarp();
  }
  __syncthreads();
}

}  // namespace mogp
