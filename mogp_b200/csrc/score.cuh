// score.cuh — per-(patient, cluster) predictive log-likelihood (GPobs.calc_score,
// mogp/obsmodel.py:75-81 -> GPy GP.log_predictive_density -> Posterior._raw_predict), the
// threshold mean (get_pred_at_loc, obsmodel.py:83-85) and the new-cluster marginal
// (GPobs.__init__ on one patient, obsmodel.py:56).
//
// For cluster c with inverse factor M_c (m x m) and query visits x*_1..x*_R (all patients' visits
// packed as columns):
//   KxT[c][col][j] = k(x_c[j], x*_col)                       k_kx   (cross covariance, built once)
//   mu[c][col]     = sum_j KxT alpha_c[j]  - A_c x*_col      k_kx
//   T = M_c Kx  (DMMA), part[c][rb][col] = sum_{i in row block rb} T[i][col]^2      k_score_gemm (<= 16 columns) /
//                                                                                   k_score_wide (24..64 columns)
//   v* = kdiag - sum_rb part, clipped at 1e-15;  lpd = -ln(2pi)/2 - ln(v*+s2)/2 - (y-mu)^2/(2(v*+s2))
//   score[p][c] = sum over the patient's visits                                    k_score_epi
// Only the diagonal predictive variance is used, visits are scored independently.
#pragma once
#include "factor.cuh"

namespace mogp {

// A CUtensorMap as the device sees it (128 opaque bytes, 64-byte aligned): the host encodes one per cluster factor
// (cuTensorMapEncodeTiled, engine.cu) and hands the kernel an array of them, one per slot of the launch.
struct alignas(64) TensorMap {
  unsigned char opaque[128];
};

// Per-slot view used by the batched scoring kernels (array of these lives in device memory).
struct SlotView {
  const double* M;      // inverse factor, row-major lower, ld
  const double* x;      // cluster inputs
  const double* alpha;  // Ky^-1 r
  int64_t ld;
  int m, mpad;
  Theta th;
  int64_t kx_off;    // offset (doubles) of this slot's KxT block in the scratch: [col][mpad]
  int64_t part_off;  // offset (doubles) of this slot's partial block: [rb][Rpad]
};

// Up to SLOT_PACK views travel inside the kernel parameters (no host->device copy on the critical path of a Gibbs
// step, where the GPU is idle until the first kernel arrives); larger batches read them from device memory.
constexpr int SLOT_PACK = 32;
struct SlotPack {
  SlotView v[SLOT_PACK];
};
#define SLOT_ARGS const SlotView* __restrict__ slots, const __grid_constant__ SlotPack pack, int use_pack
#define SLOT_AT(i) (use_pack ? pack.v[(i)] : slots[(i)])

// Cross covariance, built once per scoring pass: KxT[slot][col][j] = k(x_slot[j], x*_col) (zero in the padding),
// and the row-block partials of Kx^T alpha.  grid (ceil(mpad_max/256), ceil(Rpad/16), n_slots): one thread per
// cluster point j, 16 query columns per CTA, coalesced stores along j.
constexpr int KX_ROWS = 256, KX_COLS = 16;
__global__ void __launch_bounds__(KX_ROWS) k_kx(SLOT_ARGS, const double* __restrict__ xs, int R,
                                                int Rpad, int max_rblk, double* __restrict__ kxT,
                                                double* __restrict__ mupart) {
  const SlotView sv = SLOT_AT(blockIdx.z);
  const int j = blockIdx.x * KX_ROWS + threadIdx.x;
  if (blockIdx.x * KX_ROWS >= sv.mpad) return;
  __shared__ double wred[KX_ROWS / 32][KX_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * KX_COLS;
  const bool live = j < sv.m;
  const double xj = live ? sv.x[j] : 0.0, aj = live ? sv.alpha[j] : 0.0;
#pragma unroll 4
  for (int q = 0; q < KX_COLS; ++q) {
    const int col = c0 + q;
    if (col >= Rpad) break;
    double v = 0.0;
    if (live && col < R) v = kern_eval(sv.th, xj, xs[col], false, nullptr);
    if (j < sv.mpad) kxT[sv.kx_off + (int64_t)col * sv.mpad + j] = v;
    double p = v * aj;
#pragma unroll
    for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (lane == 0) wred[warp][q] = p;
  }
  __syncthreads();
  if (threadIdx.x < KX_COLS && c0 + threadIdx.x < Rpad) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < KX_ROWS / 32; ++w) s += wred[w][threadIdx.x];
    mupart[((int64_t)blockIdx.z * max_rblk + blockIdx.x) * Rpad + c0 + threadIdx.x] = s;
  }
}

// Launches with few clusters (re-scoring the one or two clusters a move changed) have too few (slot, row block)
// items for 148 SMs, and the heaviest item is a chain of ~40 dependent tile products.  They run with the
// contraction split: a CTA takes at most kc tiles of its row block and stores its partial T in a scratch,
// k_score_reduce adds the partials of a row block in a fixed order and emits T and the column sums of squares.
struct ChunkPlan {
  int kc = 0;        // tiles per CTA (0 = whole row block per CTA, epilogue in the same kernel)
  int maxch = 1;     // chunks of the longest row block
  double* Tpart = nullptr;  // [slot: kx_off * maxch][chunk][col][mpad]
  // 0: longest row blocks first (best balance when the pass has the GPU to itself).  1: long and short row blocks
  // alternate in launch order, so that CTAs retire all the time instead of in waves of equal length - the pass of
  // the next batch runs beside short high-priority kernels that need an SM NOW (look-ahead pipeline)
  int zigzag = 0;
};

__global__ void __launch_bounds__(256) k_score_reduce(SLOT_ARGS, int Rpad, const ChunkPlan cp, double* __restrict__ part,
                                                      double* __restrict__ Tout) {
  const SlotView sv = SLOT_AT(blockIdx.x);
  const int nb = sv.mpad / TB;
  const int rb = (int)blockIdx.y;
  if (rb >= nb) return;
  __shared__ double sq[64][TB + 1];  // [col][row]
  const int nch = rb / cp.kc + 1;
  const double* base = cp.Tpart + sv.kx_off * cp.maxch;
  const int row = threadIdx.x & 63;
  for (int c = threadIdx.x >> 6; c < Rpad; c += 4) {
    double v = 0.0;
    for (int q = 0; q < nch; ++q) v += base[((int64_t)q * Rpad + c) * sv.mpad + rb * TB + row];
    if (Tout) Tout[sv.kx_off + (int64_t)c * sv.mpad + rb * TB + row] = v;
    sq[c][row] = v * v;
  }
  __syncthreads();
  if ((int)threadIdx.x < Rpad) {
    double s = 0.0;
    for (int r = 0; r < TB; ++r) s += sq[threadIdx.x][r];
    part[sv.part_off + (int64_t)rb * Rpad + threadIdx.x] = s;
  }
}

// T = M Kx for a 64-row block and a BN-column tile; emits column sums of squares per row block.
// grid (n_col_tiles, n_slots, max_nb).  BN in {8, 16}: the HBM-bound widths (wider tiles: k_score_wide below).
template <int BN>
struct ScoreShape;
template <>
struct ScoreShape<8> { static constexpr int MS = 1, NS = 1, WM = 8, WN = 1; };
template <>
struct ScoreShape<16> { static constexpr int MS = 1, NS = 2, WM = 8, WN = 1; };

// Stages of the tile pipeline by column-tile width (one CTA per SM; a stage = one 64x64 tile of M + BN x 64 of KxT).
template <int BN>
struct ScoreStages { static constexpr int S = 5; };
template <int BN>
struct ScoreSmem {
  static constexpr int STAGE = (TB + BN) * LD_C;  // doubles per stage
  static constexpr int BYTES = ScoreStages<BN>::S * STAGE * (int)sizeof(double);
};

// The product is HBM-bound for the few columns of one patient (2 n flops per 8 bytes of M): the tiles of M are
// streamed through an S-stage ring of cp.async (LDGSTS, L1-bypassing) copies into padded rows, so the DMMA
// fragment loads stay conflict free; S-1 tiles (~40 KB each) are in flight per SM while the 8 warps run the
// DMMAs of the current one; one block barrier per tile.
// (Measured on B200: the same ring fed by 512-byte cp.async.bulk row copies + mbarriers ran 1.45x SLOWER than
// even the unpipelined loop - per-copy overhead of the TMA unit; a single bulk copy per tile needs tile-major
// storage of M, see DESIGN.md "next".  A persistent one-CTA-per-SM variant whose ring never drains between work
// items was also measured: 169 us vs 162 us per launch - the hardware CTA scheduler balances the ~1800 items
// better than a static snake deal, and the per-tile cursor logic costs more than the pipeline refills save.)
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) k_score_gemm(SLOT_ARGS, const double* __restrict__ kxT, int Rpad,
                                                            double* __restrict__ part, double* __restrict__ Tout,
                                                            const ChunkPlan cp) {
  using S = ScoreShape<BN>;
  constexpr int NS = ScoreStages<BN>::S;
  constexpr int STAGE = ScoreSmem<BN>::STAGE;
  const SlotView sv = SLOT_AT(blockIdx.y);
  const int nb = sv.mpad / TB;
  // heavy row blocks first, ACROSS the slots (the slot index varies fastest in launch order): block row rb needs
  // rb+1 tile products, so the CTA scheduler sees the work items longest-first
  const int zi = cp.kc > 0 ? (int)blockIdx.z / cp.maxch : (int)blockIdx.z;
  const int chunk = cp.kc > 0 ? (int)blockIdx.z % cp.maxch : 0;
  const int rb = nb - 1 - zi;
  if (rb < 0) return;
  const int t_begin = chunk * cp.kc;
  if (cp.kc > 0 && t_begin > rb) return;
  const int c0 = blockIdx.x * BN;
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[S::WM][BN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp / S::WN, wn = warp % S::WN;
  const int rbase = wm * 8 * S::MS, cbase = wn * 8 * S::NS;
  const double* Mrow = sv.M + (int64_t)rb * TB * sv.ld + (int64_t)t_begin * TB;
  const double* kx = kxT + sv.kx_off + (int64_t)c0 * sv.mpad + (int64_t)t_begin * TB;
  const int ntiles = cp.kc > 0 ? min(cp.kc, rb + 1 - t_begin) : rb + 1;
  auto issue = [&](int tile) {
    if (tile < ntiles) {
      double* sA = smem + (tile % NS) * STAGE;
      load_tile_async<LD_C>(sA, Mrow + (int64_t)tile * TB, sv.ld);
      load_tile_async<LD_C, BN>(sA + TB * LD_C, kx + (int64_t)tile * TB, sv.mpad);
    }
    cp_async_commit();  // (possibly empty: keeps the group count uniform)
  };
#pragma unroll
  for (int tile = 0; tile < NS - 1; ++tile) issue(tile);
  Acc<S::MS, S::NS> acc;
  acc.zero();
  for (int tile = 0; tile < ntiles; ++tile) {
    cp_async_wait<NS - 2>();  // this thread's copies of `tile` have landed
    __syncthreads();          // ... everybody's have, and everybody is done with tile-1: refill its stage
    issue(tile + NS - 1);
    const double* sA = smem + (tile % NS) * STAGE;
    warp_mma<S::MS, S::NS, true, true>(acc, sA, sA + TB * LD_C, rbase, cbase);
  }
  cp_async_wait<0>();
  // (rows beyond m are padding, exactly 0 - whatever an append running beside a look-ahead pass has written there meanwhile:
  // see wide_item)
#pragma unroll
  for (int i = 0; i < S::MS; ++i)
    if (rb * TB + rbase + 8 * i + g >= sv.m) {
#pragma unroll
      for (int j = 0; j < S::NS; ++j) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
    }
  if (cp.kc > 0) {  // split contraction: this CTA's partial product goes to the scratch, k_score_reduce finishes
    double* o = cp.Tpart + sv.kx_off * cp.maxch + (int64_t)chunk * Rpad * sv.mpad;
#pragma unroll
    for (int i = 0; i < S::MS; ++i)
#pragma unroll
      for (int j = 0; j < S::NS; ++j) {
        int r = rb * TB + rbase + 8 * i + g, c = c0 + cbase + 8 * j + 2 * t;
        o[(int64_t)c * sv.mpad + r] = acc.v[i][j][0];
        o[(int64_t)(c + 1) * sv.mpad + r] = acc.v[i][j][1];
      }
    return;
  }
  if (Tout) {  // optional: keep T (= L^-1 Kx) for the rank-n append of the chosen cluster
    double* o = Tout + sv.kx_off;
#pragma unroll
    for (int i = 0; i < S::MS; ++i)
#pragma unroll
      for (int j = 0; j < S::NS; ++j) {
        int r = rb * TB + rbase + 8 * i + g, c = c0 + cbase + 8 * j + 2 * t;
        o[(int64_t)c * sv.mpad + r] = acc.v[i][j][0];
        o[(int64_t)(c + 1) * sv.mpad + r] = acc.v[i][j][1];
      }
  }
  // column sums of squares over this warp's rows, then over the WM warps (fixed order)
#pragma unroll
  for (int j = 0; j < S::NS; ++j) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int i = 0; i < S::MS; ++i) {
      s0 += acc.v[i][j][0] * acc.v[i][j][0];
      s1 += acc.v[i][j][1] * acc.v[i][j][1];
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (g == 0) {
      red[wm][cbase + 8 * j + 2 * t] = s0;
      red[wm][cbase + 8 * j + 2 * t + 1] = s1;
    }
  }
  __syncthreads();
  if (threadIdx.x < BN) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < S::WM; ++q) s += red[q][threadIdx.x];
    part[sv.part_off + (int64_t)rb * Rpad + c0 + threadIdx.x] = s;
  }
}

// The same product for wider column tiles (several patients per pass, batch prediction): there it is bound by the
// FP64 tensor pipe, not by HBM.  16 warps = 4 row groups x 4 k groups: a warp owns 16 rows x BN columns of the
// output and a quarter of every tile's contraction range (4 of the 16 DMMA k-steps), which keeps the fragment loads
// at (2 + NS) / (2 NS) shared-memory loads per DMMA for ANY BN = 8 NS, and gives each scheduler four warps to cover
// the fragment-load latency with (ncu, 8-warp version at BN = 32: tensor pipe 67 % busy, two warps per scheduler,
// 34 % of the stall samples on fixed-latency dependencies).  The four k-group partials are added once per CTA, in a
// fixed order, through the (then idle) stage buffers.
constexpr int WIDE_THREADS = 512;
constexpr int LD_K = 72;  // smem ld of both operands of the wide kernel (8 mod 16: conflict-free 128-bit fragment loads)
template <int NS>
struct WideSmem {
  static constexpr int BN = 8 * NS;
  static constexpr int STAGE = (TB + BN) * LD_K;  // doubles per stage
  static constexpr int S = (4 * STAGE * 8 + 64 <= 227 * 1024) ? 4 : 3;
  static constexpr int BYTES = S * STAGE * (int)sizeof(double) + 64;  // + full/empty mbarriers of the ring
};

// cmap (optional, shared memory): the pass computes only ncol columns of this slot, cmap[c'] = the query column behind
// computed column c' (pairs the monotonicity rule excludes keep their threshold column only, see k_prune).  Column
// groups of 8 beyond ncol are skipped altogether (no operand loads, no DMMAs); T and the sums of squares land at the
// query columns' own positions, so everything downstream addresses them as in an unpruned pass.
template <int NS>
__device__ __forceinline__ void wide_item(const SlotView& sv, int rb, int chunk, int c0, const double* __restrict__ kxT, int Rpad,
                                          double* __restrict__ part, double* __restrict__ Tout, const ChunkPlan& cp,
                                          double* smem, double (*red)[8 * NS], const int* cmap, int ncol, const void* tmap) {
  constexpr int BN = 8 * NS, MS = 2, WM = 4;
  constexpr int NST = WideSmem<NS>::S, STAGE = WideSmem<NS>::STAGE;
  const int t_begin = chunk * cp.kc;
  if (cp.kc > 0 && t_begin > rb) return;
  const int nsa = cmap ? (ncol + 7) >> 3 : NS;  // live column groups (uniform over the CTA)
  if (nsa == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, kg = warp >> 2;  // row group, k group
  const int rbase = wm * 16;
  const double* Mrow = sv.M + (int64_t)rb * TB * sv.ld + (int64_t)t_begin * TB;
  const double* kx = kxT + sv.kx_off + (int64_t)c0 * sv.mpad + (int64_t)t_begin * TB;
  const int ntiles = cp.kc > 0 ? min(cp.kc, rb + 1 - t_begin) : rb + 1;
  // Ring of NST stages, synchronised per stage instead of per CTA: full[s] completes when every thread's copies of
  // the tile in stage s have landed (cp.async -> mbarrier), empty[s] when all 16 warps are done reading it.  The tile
  // DIST ahead is requested while the current one is multiplied, into the stage that was consumed NST - DIST tiles
  // ago, so a warp may run a whole tile ahead of the slowest one without anybody waiting: the tensor pipe keeps
  // working across tile boundaries (ncu on the __syncthreads version: 65 % busy, warps parked at the barrier).
  constexpr int DIST = NST - 2 > 0 ? NST - 2 : 1;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE);
  uint64_t* empty = full + NST;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NST; ++q) {
      mbar_init(full + q, WIDE_THREADS + (tmap ? 1 : 0));  // every thread's cp.async arrival (+ the TMA issuer's expect_tx)
      mbar_init(empty + q, WIDE_THREADS / 32);
    }
    mbar_init_fence();
  }
  __syncthreads();
  auto issue = [&](int tile) {
    if (tile < ntiles) {
      const int st = tile % NST, use = tile / NST;
      if (use > 0) mbar_wait(empty + st, (use - 1) & 1);
      double* sA = smem + st * STAGE;
      if (tmap) {
        // the 64 x 64 tile of M by TMA: one thread, one instruction - a 64-row x 72-double box of the 2-D map over the
        // row-major factor lands as the padded [64][LD_K] rows the fragment loads expect (the 8 spill columns fall
        // into the padding; beyond the matrix edge the unit fills zeros)
        if (threadIdx.x == 0) {
          mbar_expect_tx(full + st, TB * LD_K * (uint32_t)sizeof(double));
          tma_load_2d(sA, tmap, (t_begin + tile) * TB, rb * TB, full + st);
        }
      } else {
        load_tile_async<LD_K, TB, WIDE_THREADS>(sA, Mrow + (int64_t)tile * TB, sv.ld);
      }
      if (!cmap) {
        load_tile_async<LD_K, BN, WIDE_THREADS>(sA + TB * LD_K, kx + (int64_t)tile * TB, sv.mpad);
      } else {  // the live groups' columns, gathered through the map
        const double* kxb = kxT + sv.kx_off + (int64_t)(t_begin + tile) * TB;
#pragma unroll
        for (int q = 0; q < (BN * TB / 2 + WIDE_THREADS - 1) / WIDE_THREADS; ++q) {
          const int idx = threadIdx.x + q * WIDE_THREADS;
          const int r = idx >> 5, c = (idx & 31) << 1;
          if (r < 8 * nsa) cp_async16(sA + TB * LD_K + r * LD_K + c, kxb + (int64_t)cmap[r] * sv.mpad + c);
        }
      }
      cp_async_mbar_arrive(full + st);
    }
  };
#pragma unroll
  for (int tile = 0; tile < DIST; ++tile) issue(tile);
  Acc<MS, NS> acc;
  acc.zero();
  for (int tile = 0; tile < ntiles; ++tile) {
    issue(tile + DIST);
    const int st = tile % NST;
    mbar_wait(full + st, (tile / NST) & 1);
    const double* sA = smem + st * STAGE;
    warp_mma_k8_first<MS, NS, TB / 4, LD_K>(acc, sA + kg * (TB / 4), sA + TB * LD_K + kg * (TB / 4), rbase, nsa);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + st);
  }
  __syncthreads();  // every warp is done with the stage buffers: reuse them for the k-group reduction
  constexpr int LDP = BN + 2;  // partial tile [64][LDP] per k group 1..3
  if (kg > 0) {
    double* o = smem + (kg - 1) * (TB * LDP);
#pragma unroll
    for (int i = 0; i < MS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (j >= nsa) continue;
        const int r = rbase + 8 * i + g, c = 8 * j + 2 * t;
        *reinterpret_cast<double2*>(o + r * LDP + c) = make_double2(acc.v[i][j][0], acc.v[i][j][1]);
      }
  }
  __syncthreads();
  if (kg == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double* o = smem + q * (TB * LDP);
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (j >= nsa) continue;
          const int r = rbase + 8 * i + g, c = 8 * j + 2 * t;
          const double2 v = *reinterpret_cast<const double2*>(o + r * LDP + c);
          acc.v[i][j][0] += v.x;
          acc.v[i][j][1] += v.y;
        }
    }
    // Rows beyond the cluster's m are padding: identity rows of M times zero rows of Kx, exactly 0.  A pass of the look-ahead
    // pipeline runs while patients join its clusters, and an append writes its new rows of M INTO that padding: whatever the
    // tiles held when they were loaded, the pass reports the cluster it was launched for (the rows a join adds to the cached
    // columns are written by the correction chained behind the pass, chain.inl).
#pragma unroll
    for (int i = 0; i < MS; ++i)
      if (rb * TB + rbase + 8 * i + g >= sv.m) {
#pragma unroll
        for (int j = 0; j < NS; ++j) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
      }
    if (cp.kc > 0) {  // split contraction: partial product to the scratch (k_score_reduce finishes)
      double* o = cp.Tpart + sv.kx_off * cp.maxch + (int64_t)chunk * Rpad * sv.mpad;
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const int r = rb * TB + rbase + 8 * i + g, c = c0 + 8 * j + 2 * t;
          o[(int64_t)c * sv.mpad + r] = acc.v[i][j][0];
          o[(int64_t)(c + 1) * sv.mpad + r] = acc.v[i][j][1];
        }
    } else if (Tout) {
      double* o = Tout + sv.kx_off;
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (j >= nsa) continue;
          const int r = rb * TB + rbase + 8 * i + g, cl = 8 * j + 2 * t;
          if (!cmap) {
            o[(int64_t)(c0 + cl) * sv.mpad + r] = acc.v[i][j][0];
            o[(int64_t)(c0 + cl + 1) * sv.mpad + r] = acc.v[i][j][1];
          } else {
            if (cl < ncol) o[(int64_t)cmap[cl] * sv.mpad + r] = acc.v[i][j][0];
            if (cl + 1 < ncol) o[(int64_t)cmap[cl + 1] * sv.mpad + r] = acc.v[i][j][1];
          }
        }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      if (j >= nsa) continue;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int i = 0; i < MS; ++i) {
        s0 += acc.v[i][j][0] * acc.v[i][j][0];
        s1 += acc.v[i][j][1] * acc.v[i][j][1];
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (g == 0) {
        red[wm][8 * j + 2 * t] = s0;
        red[wm][8 * j + 2 * t + 1] = s1;
      }
    }
  }
  if (cp.kc > 0) return;
  __syncthreads();
  if ((int)threadIdx.x < (cmap ? ncol : BN)) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < WM; ++q) s += red[q][threadIdx.x];
    part[sv.part_off + (int64_t)rb * Rpad + (cmap ? cmap[threadIdx.x] : c0 + (int)threadIdx.x)] = s;
  }
}

// colmap / ncols (optional, one column tile only): per-slot lists of the columns to compute, see k_prune.
template <int NS>
__global__ void __launch_bounds__(WIDE_THREADS, 1) k_score_wide(SLOT_ARGS, const double* __restrict__ kxT, int Rpad,
                                                               double* __restrict__ part, double* __restrict__ Tout,
                                                               const ChunkPlan cp, const int* __restrict__ colmap,
                                                               const int* __restrict__ ncols, const TensorMap* __restrict__ tmaps) {
  constexpr int BN = 8 * NS;
  extern __shared__ __align__(128) double smem_wide[];  // (128-byte aligned: destination of the TMA boxes)
  double* smem = smem_wide;
  __shared__ double red[4][BN];
  __shared__ int s_cmap[BN];
  const SlotView sv = SLOT_AT(blockIdx.y);
  const int nb = sv.mpad / TB;
  const int zi = cp.kc > 0 ? (int)blockIdx.z / cp.maxch : (int)blockIdx.z;
  const int chunk = cp.kc > 0 ? (int)blockIdx.z % cp.maxch : 0;
  if (zi >= nb) return;
  const int rb = cp.zigzag ? ((zi & 1) ? (zi >> 1) : nb - 1 - (zi >> 1)) : nb - 1 - zi;  // heavy row blocks first, across the slots
  int ncol = BN;
  if (colmap) {
    if (threadIdx.x < BN) s_cmap[threadIdx.x] = colmap[(int64_t)blockIdx.y * Rpad + threadIdx.x];
    ncol = ncols[blockIdx.y];
    __syncthreads();
  }
  wide_item<NS>(sv, rb, chunk, blockIdx.x * BN, kxT, Rpad, part, Tout, cp, smem, red, colmap ? s_cmap : nullptr, ncol,
                tmaps ? static_cast<const void*>(tmaps + blockIdx.y) : nullptr);
}

// The monotonicity rule (mogp_constrained.py:254-266) needs only the predictive MEAN at the patient's first real visit:
// a pair with Y[n, idx] - mean > threshold gets probability 0 and the reference never calls calc_score for it.  The mean
// is O(m) per visit from k_kx's partial sums, the score O(m^2).  So before the O(m^2) product of a batch pass, one CTA
// per slot decides which (patient, slot) pairs are excluded - from exactly the sums the epilogue will add up, in the
// same order, so both agree bit for bit - and lists the columns the product has to compute for that slot: every column
// of the patients still in the race, and the threshold column alone of the excluded ones (its T, sum of squares and
// mean are what the rank-n corrections of a later move need to re-derive that mean).
constexpr int PRUNE_MAX_PAT = 64;
__global__ void __launch_bounds__(64) k_prune(SLOT_ARGS, int n_slots, const int64_t* __restrict__ off, int n_pat,
                                              const double* __restrict__ xs, int Rpad, const double* __restrict__ mupart,
                                              int max_rblk, int thr_index, const double* __restrict__ ythr, double threshold,
                                              int* __restrict__ colmap, int* __restrict__ ncols,
                                              unsigned char* __restrict__ flag) {
  const int c = blockIdx.x, p = threadIdx.x;
  const SlotView sv = SLOT_AT(c);
  const int nrb = (sv.mpad + KX_ROWS - 1) / KX_ROWS;
  __shared__ unsigned char sflag[PRUNE_MAX_PAT];
  if (p < n_pat) {
    const int64_t c0 = off[p] - off[0], c1 = off[p + 1] - off[0];
    unsigned char f = 0;
    if (thr_index >= 0 && c0 + thr_index < c1) {
      const int64_t col = c0 + thr_index;
      double m = 0.0;
      for (int rb = 0; rb < nrb; ++rb) m += mupart[((int64_t)c * max_rblk + rb) * Rpad + col];
      if (sv.th.mean_func) m += xs[col] * (-sv.th.A);
      f = (ythr[p] - m > threshold) ? 1 : 0;
    }
    sflag[p] = f;
    flag[(int64_t)p * n_slots + c] = f;
  }
  __syncthreads();
  if (p == 0) {
    int k = 0;
    int* cm = colmap + (int64_t)c * Rpad;
    for (int q = 0; q < n_pat; ++q) {
      const int c0 = (int)(off[q] - off[0]), c1 = (int)(off[q + 1] - off[0]);
      for (int col = c0; col < c1; ++col)
        if (!sflag[q] || col == c0 + thr_index) cm[k++] = col;
    }
    ncols[c] = k;
    const int fill = k > 0 ? cm[0] : 0;
    for (; k < Rpad; ++k) cm[k] = fill;
  }
}

// ---- patient table from the NaN-padded N x T matrices (mogp_constrained.py:246-247) -------------------------------
// counts per row, exclusive scan (one CTA walks the rows in chunks of 1024: not a hot path), compaction per row.
__global__ void k_pack_count(const double* __restrict__ X, const double* __restrict__ Y, int64_t N, int T,
                             int64_t* __restrict__ cnt, int* __restrict__ mismatch) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int cx = 0, cy = 0;
  for (int t = 0; t < T; ++t) {
    cx += !isnan(X[n * T + t]);
    cy += !isnan(Y[n * T + t]);
  }
  cnt[n] = cx;
  if (cx != cy) atomicExch(mismatch, 1);
}
__global__ void __launch_bounds__(1024) k_pack_scan(const int64_t* __restrict__ cnt, int64_t N, int64_t* __restrict__ off) {
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < N; base += 1024) {
    const int64_t i = base + threadIdx.x;
    int64_t v = i < N ? cnt[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive Hillis-Steele scan of the chunk
      int64_t add = (int)threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < N) off[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) off[N] = carry;
}
__global__ void k_pack_scatter(const double* __restrict__ X, const double* __restrict__ Y, int64_t N, int T,
                               const int64_t* __restrict__ off, int thr_col, double* __restrict__ x, double* __restrict__ y,
                               double* __restrict__ ythr) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t ox = off[n], oy = off[n];
  for (int t = 0; t < T; ++t) {
    const double xv = X[n * T + t], yv = Y[n * T + t];
    if (!isnan(xv)) x[ox++] = xv;
    if (!isnan(yv)) y[oy++] = yv;
  }
  ythr[n] = (thr_col >= 0 && thr_col < T) ? Y[n * T + thr_col] : nan("");
}

// Per (patient, slot): finish the predictive moments and sum the per-visit log densities.  One warp per pair:
// lanes take the patient's visits (columns), each sums its column's row-block partials in a fixed order, the
// visit log densities are added across the warp by a shuffle tree (deterministic).
// Optional per-column outputs (mean, var incl. noise, lpd) serve predict / log_predictive_density.
__global__ void __launch_bounds__(256) k_score_epi(SLOT_ARGS, int n_slots,
                                                   const int64_t* __restrict__ off, int64_t n_pat,
                                                   const double* __restrict__ xs, const double* __restrict__ ys,
                                                   int Rpad, const double* __restrict__ part,
                                                   const double* __restrict__ mupart, int max_rblk,
                                                   double* __restrict__ mu, int thr_index,
                                                   double* __restrict__ score, double* __restrict__ mu_thr,
                                                   double* __restrict__ col_mean, double* __restrict__ col_var,
                                                   double* __restrict__ col_lpd, double* __restrict__ col_ss,
                                                   const unsigned char* __restrict__ pruned) {
  const int64_t gid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gid >= n_pat * n_slots) return;
  const bool excl = pruned && pruned[gid];  // excluded by the monotonicity rule: only its threshold column was computed
  const int64_t p = gid / n_slots;
  const int c = (int)(gid % n_slots);
  const SlotView sv = SLOT_AT(c);
  const int nb = sv.mpad / TB;
  const int nrb = (sv.mpad + KX_ROWS - 1) / KX_ROWS;
  const int64_t c0 = off[p] - off[0], c1 = off[p + 1] - off[0];
  double total = 0.0;
  for (int64_t col = c0 + lane; col < c1; col += 32) {
    if (excl && col != c0 + thr_index) continue;
    double ss = 0.0;
    for (int rb = 0; rb < nb; ++rb) ss += part[sv.part_off + (int64_t)rb * Rpad + col];
    double v = kern_diag(sv.th, xs[col]) - ss;
    v = v > 1e-15 ? v : 1e-15;  // Posterior._raw_predict clips the variance
    const double vn = v + sv.th.noise;
    double m = 0.0;
    for (int rb = 0; rb < nrb; ++rb) m += mupart[((int64_t)c * max_rblk + rb) * Rpad + col];
    if (sv.th.mean_func) m += xs[col] * (-sv.th.A);  // + m(x*) = -A x*  (neg_linear.py:22-23)
    mu[(int64_t)c * Rpad + col] = m;
    if (col_ss) col_ss[(int64_t)c * Rpad + col] = ss;  // k*^T Ky^-1 k* before the clip (rank-n corrections start here)
    const double d = ys ? ys[col] - m : 0.0;
    const double lpd = -0.5 * LOG_2_PI - 0.5 * log(vn) - 0.5 * (d * d) / vn;
    total += lpd;
    if (col_mean) col_mean[(int64_t)c * Rpad + col] = m;
    if (col_var) col_var[(int64_t)c * Rpad + col] = vn;
    if (col_lpd) col_lpd[(int64_t)c * Rpad + col] = lpd;
    if (mu_thr && col == c0 + thr_index) mu_thr[gid] = m;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (lane == 0) {
    if (score) score[gid] = excl ? -INFINITY : total;
    if (mu_thr && !(thr_index >= 0 && c0 + thr_index < c1)) mu_thr[gid] = nan("");
  }
}

// New-cluster marginal: one CTA (64 threads) per patient, n <= 64 visits, all in shared memory.
//   ll = log N(y | -A x, K + (noise + 1e-8) I),  A = |a_draw|   (GPobs.__init__, obsmodel.py:44-56)
__global__ void __launch_bounds__(64) k_new_ll(const int64_t* __restrict__ off, const double* __restrict__ xs,
                                               const double* __restrict__ ys, const double* __restrict__ a_draw,
                                               Theta th, double* __restrict__ ll) {
  __shared__ double S[64][65];
  __shared__ double r[64], xv[64];
  const int64_t p = blockIdx.x;
  const int64_t c0 = off[p] - off[0];
  const int n = (int)(off[p + 1] - off[p]);
  const int tid = threadIdx.x;
  const double A = th.mean_func ? fabs(a_draw[p]) : 0.0;
  if (n > 64 || n <= 0) {
    if (tid == 0) ll[p] = nan("");
    return;
  }
  if (tid < n) {
    xv[tid] = xs[c0 + tid];
    r[tid] = ys[c0 + tid] + A * xv[tid];
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += 64) {
    int i = idx / n, j = idx % n;
    if (j <= i) {
      double v = kern_eval(th, xv[i], xv[j], i == j, nullptr);
      if (i == j) v = __dadd_rn(v, __dadd_rn(th.noise, GPY_JITTER));
      S[i][j] = v;
    }
  }
  __syncthreads();
  // right-looking Cholesky, thread per row
  bool bad = false;
  for (int k = 0; k < n; ++k) {
    double dkk = S[k][k];
    if (!(dkk > 0.0)) { bad = true; dkk = 1.0; }
    double d = sqrt(dkk);
    __syncthreads();
    if (tid == k) S[k][k] = d;
    if (tid > k && tid < n) S[tid][k] /= d;
    __syncthreads();
    if (tid > k && tid < n) {
      double lik = S[tid][k];
      for (int j = k + 1; j <= tid; ++j) S[tid][j] -= lik * S[j][k];
    }
    __syncthreads();
  }
  // forward solve z = L^-1 r (thread 0; n <= 64), quad = z.z, logdet
  if (tid == 0) {
    double quad = 0.0, lsum = 0.0;
    for (int i = 0; i < n; ++i) {
      double s = r[i];
      for (int j = 0; j < i; ++j) s -= S[i][j] * r[j];
      s /= S[i][i];
      r[i] = s;
      quad += s * s;
      lsum += log(S[i][i]);
    }
    ll[p] = bad ? nan("") : 0.5 * (-(double)n * LOG_2_PI - 2.0 * lsum - quad);
  }
}

}  // namespace mogp
