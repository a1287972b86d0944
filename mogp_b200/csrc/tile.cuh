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

// Barrier among the first 64 threads of the CTA only (warps 0 and 1), id 1.
__device__ __forceinline__ void bar64() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

// In-place lower Cholesky of the 64x64 tile s[64][LD] (lower part valid on entry).
// Returns 0 or (1 + first non-positive pivot column); result uniform across the CTA.
// The strict upper triangle is zeroed.
//
// Row-per-thread in registers: thread r < 64 keeps row r (64 doubles, statically indexed).  Step k: every
// thread publishes its a[k] (one shared column, double-buffered), one 64-thread barrier, then
// a[j] -= a[k] * col[j] / col[k] for k < j <= r.  One barrier and one reciprocal per column; the other six warps
// wait at the closing block barrier.  (The 128-barrier shared-memory version of this routine was 70 % of the
// refit time: ncu, profiles/r1_v0_launches_refit_region.csv.)
template <int LD>
__device__ int potrf_tile(double* __restrict__ s, double* __restrict__ col /* >= 128 doubles */) {
  const int r = threadIdx.x;
  __shared__ int s_fail;
  if (r == 0) s_fail = 0;
  __syncthreads();
  if (r < TB) {
    double a[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) a[j] = (j <= r) ? s[r * LD + j] : 0.0;
#pragma unroll
    for (int k = 0; k < TB; ++k) {
      double* cb = col + (k & 1) * TB;
      cb[r] = a[k];
      bar64();
      double dk = cb[k];
      if (!(dk > 0.0)) {
        if (r == k && s_fail == 0) s_fail = k + 1;
        dk = 1.0;  // keep going with a harmless pivot; the caller discards the result
      }
      const double inv = 1.0 / dk;
      const double t = a[k] * inv;
#pragma unroll
      for (int j = k + 1; j < TB; ++j)
        if (j <= r) a[j] -= t * cb[j];
      a[k] = (r == k) ? sqrt(dk) : a[k] * rsqrt(dk);
    }
#pragma unroll
    for (int j = 0; j < TB; ++j) s[r * LD + j] = (j <= r) ? a[j] : 0.0;
  }
  __syncthreads();
  return s_fail;
}

// sM = inverse of the lower-triangular tile sL (both [64][LD]); upper triangle of sM zeroed.
// Column-per-thread in registers: thread c < 64 solves L x = e_c by forward substitution with x[0..63] statically
// indexed (x_i = 0 for i < c falls out of the recurrence, so all threads run the same straight-line code);
// the entries of L are shared-memory broadcasts; four partial sums per row shorten the dependent chain.
template <int LD>
__device__ void trtri_tile(const double* __restrict__ sL, double* __restrict__ sM) {
  const int c = threadIdx.x;
  if (c < TB) {
    double x[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
      for (int j = 0; j + 3 < i; j += 4) {
        p0 += sL[i * LD + j] * x[j];
        p1 += sL[i * LD + j + 1] * x[j + 1];
        p2 += sL[i * LD + j + 2] * x[j + 2];
        p3 += sL[i * LD + j + 3] * x[j + 3];
      }
#pragma unroll
      for (int j = (i / 4) * 4; j < i; ++j) p0 += sL[i * LD + j] * x[j];
      const double rhs = (i == c) ? 1.0 : 0.0;
      x[i] = (rhs - ((p0 + p1) + (p2 + p3))) / sL[i * LD + i];
    }
#pragma unroll
    for (int i = 0; i < TB; ++i) sM[i * LD + c] = x[i];
  }
  __syncthreads();
}

}  // namespace mogp
