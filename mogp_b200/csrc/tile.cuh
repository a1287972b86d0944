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

// ---- mbarriers and TMA (sm_90+/sm_100a) ----
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
// TMA tile load: one box of the 2-D tensor map (coordinates: c0 = column, c1 = row of its first element) into shared
// memory, completion counted in bytes on the mbarrier (SASS: UTMALDG).  Issued by ONE thread.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
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
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {  // (8-byte form: sources off a 16-byte boundary)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Asynchronous version of load_tile: ROWS x 64 tile into s[ROWS][LD].
template <int LD, int ROWS = TB, int NT = NTHREADS>
__device__ __forceinline__ void load_tile_async(double* __restrict__ s, const double* __restrict__ g, int64_t ld) {
#pragma unroll
  for (int q = 0; q < (ROWS * TB / 2 + NT - 1) / NT; ++q) {
    int idx = threadIdx.x + q * NT;
    if ((ROWS * TB / 2) % NT == 0 || idx < ROWS * TB / 2) {
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

// cp.async completion tracked by an mbarrier: the arrive fires when all cp.async this thread has issued so far have
// landed (.noinc: it counts as one of the barrier's expected arrivals).
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// The same product with 16-byte fragment loads: DMMA contracts 4 k-indices per instruction and any 4 will do as long as
// A and B agree, so lane t takes k = kb + 2t and kb + 2t + 1 from ONE 128-bit load and feeds them to two consecutive
// DMMAs (k = kb + {0,2,4,6}, then kb + {1,3,5,7}): half the shared-memory load instructions per DMMA.  Both operands
// are stored with the contraction index contiguous ([row][LD], [col][LD]); LD = 8 mod 16 keeps the 128-bit loads of a
// quarter warp (two rows x 64 bytes) on distinct banks.
template <int MS, int NS, int KDIM, int LD>
__device__ __forceinline__ void warp_mma_k8(Acc<MS, NS>& acc, const double* __restrict__ sA, const double* __restrict__ sB,
                                            int rb, int cb) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kb = 0; kb < KDIM; kb += 8) {
    double2 a[MS], b[NS];
#pragma unroll
    for (int i = 0; i < MS; ++i) a[i] = *reinterpret_cast<const double2*>(sA + (rb + 8 * i + g) * LD + kb + 2 * t);
#pragma unroll
    for (int j = 0; j < NS; ++j) b[j] = *reinterpret_cast<const double2*>(sB + (cb + 8 * j + g) * LD + kb + 2 * t);
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].y, b[j].y);
  }
}

// The same with only the first NA (compile time) of the NS column groups live: the others cost neither fragment loads
// nor DMMAs.  Column base 0.  warp_mma_k8_first dispatches a run-time count (uniform over the warp) to these
// straight-line bodies - per-instruction predicates around the DMMAs cost more issue slots than the skipped work saves.
template <int MS, int NS, int NA, int KDIM, int LD>
__device__ __forceinline__ void warp_mma_k8_na(Acc<MS, NS>& acc, const double* __restrict__ sA, const double* __restrict__ sB,
                                               int rb) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kb = 0; kb < KDIM; kb += 8) {
    double2 a[MS], b[NA];
#pragma unroll
    for (int i = 0; i < MS; ++i) a[i] = *reinterpret_cast<const double2*>(sA + (rb + 8 * i + g) * LD + kb + 2 * t);
#pragma unroll
    for (int j = 0; j < NA; ++j) b[j] = *reinterpret_cast<const double2*>(sB + (8 * j + g) * LD + kb + 2 * t);
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NA; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NA; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i].y, b[j].y);
  }
}
template <int MS, int NS, int KDIM, int LD>
__device__ __forceinline__ void warp_mma_k8_first(Acc<MS, NS>& acc, const double* __restrict__ sA, const double* __restrict__ sB,
                                                  int rb, int nsa) {
  switch (nsa) {
    case 1: warp_mma_k8_na<MS, NS, 1, KDIM, LD>(acc, sA, sB, rb); break;
    case 2: if constexpr (NS >= 2) warp_mma_k8_na<MS, NS, 2, KDIM, LD>(acc, sA, sB, rb); break;
    case 3: if constexpr (NS >= 3) warp_mma_k8_na<MS, NS, 3, KDIM, LD>(acc, sA, sB, rb); break;
    case 4: if constexpr (NS >= 4) warp_mma_k8_na<MS, NS, 4, KDIM, LD>(acc, sA, sB, rb); break;
    case 5: if constexpr (NS >= 5) warp_mma_k8_na<MS, NS, 5, KDIM, LD>(acc, sA, sB, rb); break;
    case 6: if constexpr (NS >= 6) warp_mma_k8_na<MS, NS, 6, KDIM, LD>(acc, sA, sB, rb); break;
    case 7: if constexpr (NS >= 7) warp_mma_k8_na<MS, NS, 7, KDIM, LD>(acc, sA, sB, rb); break;
    default: if constexpr (NS >= 8) warp_mma_k8_na<MS, NS, 8, KDIM, LD>(acc, sA, sB, rb); break;
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
// double-buffered 64-entry shared array, costs ONE barrier, and updates every patch with 16 FMAs.  The loop over
// the 16 patch columns is a real loop (compact code: a fully unrolled row-per-thread variant was instruction-fetch
// bound), the 4 columns inside a patch are unrolled so register indices stay static.
//
// The 64 pivots are a serial dependency chain, and FP64 latency on this part is long (measured ~770 cycles per
// pivot with sqrt + divide on the chain).  So: the factorisation runs as L D L^T - per pivot only a reciprocal
// (hardware seed + two Newton steps = 4 dependent FMAs) - the square roots are taken for all columns at once at
// the end; and warps whose rows are all above the pivot leave the loop, so the barrier shrinks as it goes
// (named barrier with the live thread count).

// 1/d to <= 1 ulp: rcp.approx seed (~2^-23) and two Newton-Raphson steps.
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}
// barrier among the warps first..7 of the CTA (they are the ones still in the loop)
__device__ __forceinline__ void bar_live(int first_warp) {
  asm volatile("bar.sync 1, %0;" ::"r"((8 - first_warp) * 32) : "memory");
}

// In-place lower Cholesky of s[64][LD] (lower part valid on entry); strict upper triangle zeroed.
// rdiag[k] = 1 / L_kk.  Returns 0 or (1 + first non-positive pivot column), uniform across the CTA.
//
// Four columns per barrier: the owners of patch column kb publish their (not yet eliminated) 4-column block; after
// ONE barrier every live thread redoes the tiny 4x4 L D L^T of the diagonal patch in registers (four reciprocals -
// the serial chain), eliminates the rows it needs through those four pivots, and applies the rank-4 update to its
// own patch.  16 barriers instead of 64; the pivot chain (64 reciprocals) is what remains.
template <int LD>
__device__ int potrf_tile(double* __restrict__ s, double* __restrict__ buf /* >= 512 doubles */,
                          double* __restrict__ rdiag /* 64 */) {
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15, warp = tid >> 5;
  __shared__ int s_fail;
  __shared__ double dpiv[TB];  // the pivots d_k of L D L^T
  if (tid == 0) s_fail = 0;
  __syncthreads();  // the tile was just staged by other threads
  double a[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) a[e][f] = (4 * tj + f <= 4 * ti + e) ? s[(4 * ti + e) * LD + 4 * tj + f] : 0.0;
  // warp w holds patch rows ti = 2w, 2w+1: it is needed while kb <= 2w + 1
  for (int kb = 0; kb <= 2 * warp + 1; ++kb) {
    const int first_warp = kb >> 1;
    double* pb = buf + (kb & 1) * (4 * TB);  // [64 rows][4]: the column block before its own elimination
    if (tj == kb && ti >= kb) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < 4; ++q) pb[(4 * ti + e) * 4 + q] = a[e][q];
    }
    bar_live(first_warp);
    if (ti >= kb && tj >= kb && tj <= ti) {
      // 4x4 L D L^T of the diagonal patch (unscaled columns): D4[e][q], e >= q; u[q] = 1 / d_q
      double D4[4][4], u[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < 4; ++q) D4[e][q] = (q <= e) ? pb[(4 * kb + e) * 4 + q] : 0.0;
      bool bad = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double dq = D4[q][q];
        if (!(dq > 0.0)) {
          bad = true;
          if (ti == kb && tj == kb && s_fail == 0) s_fail = 4 * kb + q + 1;
          dq = 1.0;  // keep going with a harmless pivot; the caller discards the result
          D4[q][q] = dq;
        }
        u[q] = fast_rcp(dq);
#pragma unroll
        for (int e = q + 1; e < 4; ++e) {
          const double t = D4[e][q] * u[q];
#pragma unroll
          for (int f = q + 1; f <= e; ++f) D4[e][f] -= t * D4[f][q];
        }
      }
      (void)bad;
      if (ti == kb) {  // the diagonal patch itself (tj == kb here)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int q = 0; q <= e; ++q) a[e][q] = D4[e][q];
#pragma unroll
        for (int q = 0; q < 4; ++q) dpiv[4 * kb + q] = D4[q][q];
      } else {
        // eliminate the rows this thread needs through the four pivots: row r of the block, values v[0..3]:
        //   v[f] -= v[q] * D4[f][q] * u[q]  for f > q
        double ci[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int q = 0; q < 4; ++q) ci[e][q] = pb[(4 * ti + e) * 4 + q];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const double t = ci[e][q] * u[q];
#pragma unroll
            for (int f = q + 1; f < 4; ++f) ci[e][f] -= t * D4[f][q];
          }
        }
        if (tj == kb) {  // owner of the column block: these are its final (unscaled) columns
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int q = 0; q < 4; ++q) a[e][q] = ci[e][q];
        } else {
          double cj[4][4];
#pragma unroll
          for (int f = 0; f < 4; ++f) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cj[f][q] = pb[(4 * tj + f) * 4 + q];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const double t = cj[f][q] * u[q];
#pragma unroll
              for (int g = q + 1; g < 4; ++g) cj[f][g] -= t * D4[g][q];
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              if (4 * tj + f <= 4 * ti + e) {
                double acc = a[e][f];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc -= (ci[e][q] * u[q]) * cj[f][q];
                a[e][f] = acc;
              }
            }
        }
      }
    }
  }
  __syncthreads();
  // L = (unscaled columns) * D^-1/2
  if (tid < TB) rdiag[tid] = rsqrt(dpiv[tid]);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int i = 4 * ti + e, j = 4 * tj + f;
      s[i * LD + j] = (j <= i) ? a[e][f] * rdiag[j] : 0.0;
    }
  __syncthreads();
  return s_fail;
}

// sM = inverse of the lower-triangular tile sL (both [64][LD]); upper triangle of sM zeroed.
// Right-looking on X (= I at the start), four rows per barrier: the threads that own patch row jb finish its four
// rows locally (row j scaled by 1/L_jj = rdiag[j], rows j+1.. of the patch reduced by it), publish them, and every
// patch row below applies the rank-4 update x_i -= sum_q L_{i,4jb+q} * row_q.  16 barriers instead of 64.
template <int LD>
__device__ void trtri_tile(const double* __restrict__ sL, double* __restrict__ sM, double* __restrict__ buf /* >= 512 */,
                           const double* __restrict__ rdiag) {
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15, warp = tid >> 5;
  double x[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) x[e][f] = (4 * ti + e == 4 * tj + f) ? 1.0 : 0.0;
  for (int jb = 0; jb <= 2 * warp + 1; ++jb) {
    const int first_warp = jb >> 1;
    double* rb = buf + (jb & 1) * (4 * TB);  // [4][64]
    if (ti == jb) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 4 * jb + q;
        const double inv = rdiag[j];
#pragma unroll
        for (int f = 0; f < 4; ++f) x[q][f] *= inv;
#pragma unroll
        for (int e = q + 1; e < 4; ++e) {
          const double l = sL[(4 * jb + e) * LD + j];
#pragma unroll
          for (int f = 0; f < 4; ++f) x[e][f] -= l * x[q][f];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int f = 0; f < 4; ++f) rb[q * TB + 4 * tj + f] = x[q][f];
    }
    bar_live(first_warp);
    if (ti > jb && tj <= jb) {  // patch rows below; rows of block jb have entries only in columns <= 4jb+3
      double r[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int f = 0; f < 4; ++f) r[q][f] = rb[q * TB + 4 * tj + f];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double l = sL[(4 * ti + e) * LD + 4 * jb + q];
#pragma unroll
          for (int f = 0; f < 4; ++f) x[e][f] -= l * r[q][f];
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) sM[(4 * ti + e) * LD + 4 * tj + f] = (4 * tj + f <= 4 * ti + e) ? x[e][f] : 0.0;
  __syncthreads();
}

}  // namespace mogp
