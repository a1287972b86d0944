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

// ---- TMA bulk copies + mbarrier (sm_90+/sm_100a): cp.async.bulk moves whole 512-byte tile rows from global
// to (padded) shared-memory rows and signals an mbarrier with the byte count; SASS: UBLKCP + SYNCS. ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// ---- cp.async (LDGSTS) 16-byte copies: global -> padded shared rows, no register staging ----
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Asynchronous version of load_tile: ROWS x 64 tile into s[ROWS][LD].
template <int LD, int ROWS = TB>
__device__ __forceinline__ void load_tile_async(double* __restrict__ s, const double* __restrict__ g, int64_t ld) {
#pragma unroll
  for (int q = 0; q < (ROWS * TB / 2 + NTHREADS - 1) / NTHREADS; ++q) {
    int idx = threadIdx.x + q * NTHREADS;
    if ((ROWS * TB / 2) % NTHREADS == 0 || idx < ROWS * TB / 2) {
      int r = idx >> 5, c = (idx & 31) << 1;
      cp_async16(s + r * LD + c, g + (int64_t)r * ld + c);
    }
  }
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

// ---- 64x64 diagonal tile: Cholesky and triangular inverse, patch-distributed in registers -------------------
// All 256 threads take part: thread (ti, tj) = (tid >> 4, tid & 15) keeps the 4x4 patch rows 4ti.., cols 4tj.. in
// registers (statically indexed).  Each of the 64 column steps publishes one column (or row) through a
// double-buffered 64-entry shared array, costs ONE block barrier, and updates every patch with 16 FMAs.
// The loop over the 16 patch columns is a real loop (compact code: the fully unrolled row-per-thread variant was
// instruction-fetch bound), the 4 columns inside a patch are unrolled so register indices stay static.
//
// In-place lower Cholesky of s[64][LD] (lower part valid on entry); strict upper triangle zeroed.
// Returns 0 or (1 + first non-positive pivot column), uniform across the CTA.
template <int LD>
__device__ int potrf_tile(double* __restrict__ s, double* __restrict__ buf /* >= 128 doubles */) {
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;
  __syncthreads();  // the tile was just staged by other threads
  double a[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) a[e][f] = (4 * tj + f <= 4 * ti + e) ? s[(4 * ti + e) * LD + 4 * tj + f] : 0.0;
  for (int kb = 0; kb < 16; ++kb) {
#pragma unroll
    for (int kf = 0; kf < 4; ++kf) {
      const int k = 4 * kb + kf;
      double* cb = buf + (k & 1) * TB;
      if (tj == kb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) cb[4 * ti + e] = a[e][kf];
      }
      __syncthreads();
      double dk = cb[k];
      if (!(dk > 0.0)) {
        if (tid == 0 && s_fail == 0) s_fail = k + 1;
        dk = 1.0;  // keep going with a harmless pivot; the caller discards the result
      }
      if (ti >= kb && tj <= ti) {
        const double inv = 1.0 / dk;
        double ci[4], cj[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ci[e] = cb[4 * ti + e] * inv;
          cj[e] = cb[4 * tj + e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int i = 4 * ti + e, j = 4 * tj + f;
            if (j > k && j <= i) a[e][f] -= ci[e] * cj[f];
          }
        if (tj == kb) {
          const double rs = rsqrt(dk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * ti + e;
            a[e][kf] = (i == k) ? sqrt(dk) : (i > k ? a[e][kf] * rs : a[e][kf]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) s[(4 * ti + e) * LD + 4 * tj + f] = (4 * tj + f <= 4 * ti + e) ? a[e][f] : 0.0;
  __syncthreads();
  return s_fail;
}

// sM = inverse of the lower-triangular tile sL (both [64][LD]); upper triangle of sM zeroed.
// Right-looking on X (= I at the start): step j scales row j by 1/L_jj, publishes it, and every row i > j
// subtracts L_ij * row j.
template <int LD>
__device__ void trtri_tile(const double* __restrict__ sL, double* __restrict__ sM, double* __restrict__ buf /* >= 128 */) {
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  double x[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) x[e][f] = (4 * ti + e == 4 * tj + f) ? 1.0 : 0.0;
  for (int jb = 0; jb < 16; ++jb) {
#pragma unroll
    for (int jf = 0; jf < 4; ++jf) {
      const int j = 4 * jb + jf;
      double* rb = buf + (j & 1) * TB;
      if (ti == jb) {
        const double inv = 1.0 / sL[j * LD + j];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          x[jf][f] *= inv;
          rb[4 * tj + f] = x[jf][f];
        }
      }
      __syncthreads();
      if (ti >= jb && tj <= jb) {  // rows below j; row j has entries only in columns <= j
        double r[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) r[f] = rb[4 * tj + f];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * ti + e;
          if (i > j) {
            const double l = sL[i * LD + j];
#pragma unroll
            for (int f = 0; f < 4; ++f) x[e][f] -= l * r[f];
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) sM[(4 * ti + e) * LD + 4 * tj + f] = (4 * tj + f <= 4 * ti + e) ? x[e][f] : 0.0;
  __syncthreads();
}

}  // namespace mogp
