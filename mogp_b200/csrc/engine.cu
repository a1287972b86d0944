// engine.cu — host side of the C ABI declared in include/mogp_b200.h.
//
// One engine = one CUDA stream + a table of cluster slots.  A slot is the device image of one
// reference `GPobs` object (mogp/obsmodel.py:6-94): its data (x, y), its hyper-parameters theta and the
// result of exact GP inference at theta — the INVERSE Cholesky factor M = chol(K + (noise+1e-8) I)^-1,
// alpha = Ky^-1 (y - m(x)) and the log marginal likelihood.  Scoring consumes M as a product
// (T = M Kx, DMMA) instead of GPy's triangular solve (dtrtrs), which has no parallel slack.
//
// Nothing here falls back to the CPU: every numerical entry point launches kernels of this library
// and fails with MOGP_ECUDA when there is no device.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>
#include <deque>

#include "../../include/mogp_b200.h"
#include "lbfgsb.hpp"
#include "score.cuh"
#include "update.cuh"
#include "kmeans.cuh"
#include "wspace.cuh"

using namespace mogp;

constexpr int ALG_OUT_SLOTS = 2 + 30;  // outputs of the two within-batch corrections + of a replay (one region per slot)
constexpr int REPLAY_MAX = 24;         // corrections queued ahead for the next batch's positions, per batch

// The join-side score correction a commit launches right after k_append_small (append_core), see update.cuh.
struct AlgHook {
  AlgJobs jobs;
  int npos = 0, ntiles = 0;
  bool fired = false, on_quick = false;
  int32_t slot = -1;  // the cluster joined and the batch buffer its cached columns live in (t_wait / t_mark)
  int bi = 0;
};
// One rank-n change of the batch being processed, kept so that the same correction can be applied to the NEXT batch's
// positions when their pass (launched before the change) is absorbed.
struct ChangeLog {
  int32_t slot = -1;
  int kind = 0;            // 0 leave, 1 join
  AlgJobs J;               // everything but pos[] / out; leave: M / alpha point at the logged copies
  int ntiles = 0;
  int64_t epoch_before = 0;
};
// The leave-side half of keeping the cached T current (k_T_leave): filled by the step that moves a patient, fired by
// remove_core as soon as the deletion's operands exist (k_rm_prep).  Target 0 = the positions still ahead in the batch being
// processed, target 1 = the positions of the next batch (their pass, launched before the move, is corrected behind it).
constexpr int NBUF = 4;       // batch buffers of the look-ahead pipeline: the batch being processed + up to two passes in flight
                              // + one that stays untouched until the corrections reading it have run
constexpr int NTGT = 3;       // targets of a move's corrections: the current batch + the passes in flight
struct LeaveHook {
  AlgJobs J[NTGT];
  int npos[NTGT] = {0, 0, 0};
  int bi[NTGT] = {0, 0, 0};   // batch buffer each target's columns live in
  int32_t slot = -1;
  bool fired[NTGT] = {false, false, false};
};
constexpr int TS_RING = 64;   // private copies of a deletion's operands for k_T_leave (three batches of moves in flight)
constexpr int TEV_RING = 1024; // events ordering the readers / writers of the cached columns of one (batch, slot)
namespace {

thread_local std::string g_create_error;

inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

struct Cluster {
  bool used = false;
  int32_t gen = 0;   // bumped every time the slot is freed: part of the handle the ABI hands out (slot_handle)
  mogp_model_spec spec{};
  double theta[4] = {0, 0, 0, 0};
  int64_t m = 0;     // data points
  int64_t cap = 0;   // allocated edge of M (multiple of 64) = leading dimension
  double* M = nullptr;    // one pooled buffer: M[cap*cap] followed by vec
  double* vec = nullptr;  // x | y | alpha | z (cap each) | stat[STAT]   (= M + cap*cap)
  bool theta_dirty = true;  // device copy of theta (stat + 16) is stale
  std::vector<double> hx, hy;     // host mirror of the data (pickling, jitter mean, get_data)
  std::vector<int64_t> members;   // chain-managed clusters: patient ids in data order
  bool valid = false;             // factor / alpha / lml correspond to (data, theta)
  double lml = 0.0, quad = 0.0, logdet = 0.0;
  double grad[4] = {0, 0, 0, 0};
  bool grad_valid = false;
  // a collapsed shadow of another cluster (refit lanes): wt = device multiplicities of its points, corr_* = the closed-form
  // terms that turn its log marginal into the full cluster's (objective_end)
  const double* wt = nullptr;
  double corr_n1 = 0.0, corr_logn = 0.0, corr_ss = 0.0;
  TensorMap tmap;                 // 2-D tensor map over M (64-row x 72-double boxes) for the scoring kernel's TMA loads ...
  const double* tmap_M = nullptr; // ... encoded for this buffer
  int64_t tmap_cap = 0;           // ... and this leading dimension
  bool side_busy = false;         // a row deletion of this cluster may still be running on the side stream
  bool stat_on_device = false;    // lml / quad / logdet of the last rank update not fetched yet
  bool lml_dirty = false;         // ... and not even reduced yet (k_lml runs when somebody asks)
  bool z_dirty = false;           // z = M r not recomputed after a row deletion (k_z runs when somebody asks)
  double* x() const { return vec; }
  double* y() const { return vec + cap; }
  double* alpha() const { return vec + 2 * cap; }
  double* z() const { return vec + 3 * cap; }
  double* stat() const { return vec + 4 * cap; }  // [0..3] lml, quad, logdet, x.alpha; [8] info; [10..13] grad; [16..] Theta
  Theta* theta_dev() const { return reinterpret_cast<Theta*>(vec + 4 * cap + 16); }
};

constexpr int STAT = 32;  // doubles in a cluster's stat block

struct ScorePlan {
  int64_t R = 0, Rpad = 0;
  int BN = 8;
  int n_slots = 0;
  int max_nb = 0;
  int64_t kx_total = 0, part_total = 0;
  int max_rblk = 0;  // row blocks (of KX_ROWS) of the largest slot
  double *kxT = nullptr, *part = nullptr, *mupart = nullptr, *mu = nullptr, *T = nullptr, *ss = nullptr;
};

struct Chain {
  bool active = false;
  mogp_chain_config cfg{};
  int64_t N = 0;
  std::vector<int32_t> z;
  std::vector<int64_t> Nk;
  std::vector<int32_t> slot_of_id;
  int64_t pending = -1;      // patient currently removed (between begin and commit)
  int32_t pending_prev = -1; // its previous cluster id
  int32_t pending_lazy = 0;  // fast mode: the patient's rows are still inside its previous cluster
  int64_t pending_p = -1;    // ... at this row offset
  bool gram_ready = false;   // k_rm_gram has run for (previous cluster, pending_p)
  std::vector<int32_t> score_slots;  // slots scored by the batched pass of this step, in order
  ScorePlan plan;                    // where their T / mu live in the scratch
  double a_dev_val = 0.0;
  double last_margin = 0.0;
  std::vector<int32_t> act_ids;
  std::vector<double> act_score;
  // ---- look-ahead window of mogp_chain_sweep (fast mode): several consecutive patients are scored against every
  // active cluster in ONE pass over the factors; a cached score stays valid until its cluster changes ----
  struct WinEntry {
    double score = 0.0, muthr = 0.0;
    bool pruned = false;         // excluded by the monotonicity rule before the product: only its threshold column has T / ss / mu
    int64_t stride = 0;          // rows between two columns of T (the slot's padded size when the columns were computed)
    const double* T = nullptr;   // the patient's columns of T = M Kx under that slot (stride = mpad when scored)
    const double* mu = nullptr;  // ... their predictive means
    const double* ss = nullptr;  // ... and their k*^T Ky^-1 k* (device, as mu)
  };
  struct Batch {                   // consecutive sweep positions scored by one pass
    int64_t s0 = 0, cnt = 0;
    std::vector<int32_t> slots;    // slots the pass scored ...
    std::vector<int64_t> epoch;    // ... their change counters when it was launched ...
    std::vector<int64_t> mpad;     // ... and their padded sizes (layout of T in the pass's scratch)
    std::vector<int64_t> msz;      // ... and their sizes (accounting of the pruned columns)
    bool prof = false;             // the pass was launched with profiling on
    std::vector<int32_t> own_slot; // [cnt] slot whose leave-block-out score was launched with the pass (-1: none)
    std::vector<int64_t> own_epoch;
    ScorePlan pl;
  };
  struct Window {
    bool open = false;             // a pipelined sweep is running: cluster changes must be recorded (win_invalidate)
    bool in_step = false;          // the pending patient is served from the cache (locate_T)
    int64_t n = 0;                 // steps of the sweep
    std::vector<int64_t> pat;      // patient ids, in sweep order
    std::vector<int64_t> col0;     // [n + 1] first column of each position in the packed query
    int64_t cur = 0;               // position being processed
    int64_t end_launched = 0;      // positions below have a pass launched (their cache entries exist or are on the way)
    std::vector<std::vector<WinEntry>> e;  // [position][slot]
    int64_t batch_end = 0;         // end of the batch being processed
    std::vector<int64_t> slot_upto;   // [slot]: the cached SCORES are current for the positions in [cur, slot_upto)
    std::vector<char> slot_T_ok;      // [slot]: ... and so are the cached T / mu / ss on the device
    std::vector<char> slot_alg;       // [slot]: the slot took a rank-n score correction in this batch (one per batch)
    std::vector<int64_t> slot_epoch;  // [slot]: incremented whenever the cluster in that slot changes
    std::vector<double> new_ll;    // [position]
    std::vector<double> own_ll, own_mu, own_flag;  // [position] leave-block-out score under the patient's own cluster
    std::vector<char> own_ok;
    Batch b[NBUF];
    std::vector<int> flight;       // buffers of the passes launched and not yet absorbed (batches after the current one), oldest first
    // rank-n corrections of the positions of a pass in flight, queued behind that pass at the moment a move of the current
    // batch changes a cluster (replay_launch); harvested when that batch is absorbed
    struct AsyncItem {
      int32_t slot = -1;
      int64_t epoch_before = 0;
      int out_idx = 0;
    };
    std::vector<AsyncItem> async_items[NBUF];  // per pass in flight
    int64_t rpw_used[NBUF] = {};   // doubles used in that pass's share of the engine's d_rpw
    int cur_bi = 0;                // buffer index of the batch being processed
    int64_t blog_used = 0;         // doubles used in the engine's d_blog (this batch's half: blog_base, blog_half)
    int64_t blog_base = 0, blog_half = 0;
    // cached columns kept current on the device (MOGP_TMAINT): [batch buffer][slot] index of the event recorded after the last
    // kernel that reads or writes them (-1: none pending); rp_epoch[buffer][slot]: change counter up to which the pass in
    // flight in that buffer has been brought by corrections queued behind it (-1: the chain of corrections is broken);
    // rp_last[buffer][slot]: output block of the last of those corrections
    std::vector<int> slot_tev[NBUF];
    std::vector<int64_t> rp_epoch[NBUF];
    std::vector<int> rp_last[NBUF];
  } win;
  int64_t win_passes = 0, win_rescores = 0, win_patients = 0, win_algebra = 0, win_replayed = 0;  // accounting
  int64_t win_pairs = 0, win_pruned = 0;   // (patient, cluster) pairs of the window passes / those excluded before the product
  double t_open = 0.0, t_ensure = 0.0, t_commit = 0.0, t_alg = 0.0;
  double t_chain = 0.0;          // host seconds waiting for the corrections chained behind a pass when its batch starts
  int64_t win_chained = 0;       // corrections chained behind passes in flight
  double max_corr_err = 0.0;     // MOGP_CHECK_CORRECTIONS: worst |rank-n corrected - re-scored| (relative)
  int64_t corr_checked = 0;         // host seconds spent in the three phases of windowed steps
};

// Device scratch + slot-view staging of a scoring pass.  Several passes can be in flight (the next batch's pass on the
// look-ahead stream, a re-scoring pass on the main stream), each on its own resources.
struct ScoreRes {
  double* scr = nullptr;
  int64_t scr_elems = 0, scr_used = 0;
  bool bump = false;               // bump-allocate: blocks stay alive (T = M Kx of earlier passes is still referenced)
  std::vector<double*> retired;    // bump blocks outgrown while still referenced
  SlotView* d_slots = nullptr;     // views of launches with more than SLOT_PACK slots
  int64_t slots_cap = 0;
  int zigzag = 0;                  // launch order of the wide kernel's row blocks (ChunkPlan::zigzag)
  SlotView* h_slots = nullptr;     // own pinned staging of those views (null: the engine's upload staging)
  int64_t hslots_cap = 0;
  TensorMap* d_tmaps = nullptr;    // tensor maps of the launch's slots (TMA loads of the factor tiles)
  int64_t tmaps_cap = 0;
  TensorMap* h_tmaps = nullptr;    // own pinned staging of those (passes on the look-ahead stream)
  int64_t htmaps_cap = 0;
  // a pass of the look-ahead pipeline: its short preamble / epilogue run on `pre` (high priority), handed over by ev_pre0 /
  // ev_pre1; the long product waits for `wide_after` (the product of the pass before it: passes in flight do not share the
  // SMs, the later one only fills the earlier one's tail) and records `wide_done`
  cudaStream_t pre = nullptr;
  cudaEvent_t ev_pre0 = nullptr, ev_pre1 = nullptr, wide_done = nullptr, wide_after = nullptr;
  cudaEvent_t pre_done = nullptr;  // the preamble has read the clusters' x / alpha (an append updates alpha in place: it waits)
};

// One batch of the look-ahead pipeline (chain.inl): outputs of its scoring pass, device and pinned host.
struct BatchBuf {
  ScoreRes res;
  double* d_out = nullptr;
  int64_t out_elems = 0;
  double* h_out = nullptr;
  int64_t hout_elems = 0;
  cudaEvent_t scored = nullptr;    // the pass's kernels have finished (T / ss / mu of the batch are on the device)
  cudaEvent_t done = nullptr;      // the pass and its read-back have finished (host spins on it)
  cudaEvent_t done_blk = nullptr;  // the same for engines in blocking-sync mode (host sleeps: many chains per process)
};

}  // namespace

// Small append workspaces: up to AP_RING appends may be queued on the main stream while the next one's Gram / Ls are
// already computed on the quick stream (a batch holds at most 16 patients, and every batch start waits for the main
// stream).  A slot serves clusters of up to AP_RING_NBLK * 256 points; larger ones use the shared workspace.
constexpr int AP_RING = 64, AP_RING_NBLK = 256;
constexpr int64_t AP_RING_SLOT = NP * NP + NP + (int64_t)AP_RING_NBLK * NG;

struct mogp_engine {
  int device = 0;
  cudaStream_t stream = nullptr;   // the stream every helper launches on (normally `main_stream`)
  cudaStream_t main_stream = nullptr, side = nullptr;  // side: independent work of a Gibbs step (SideScope)
  cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_block = nullptr;
  bool blocking_sync = false;
  std::vector<std::pair<int64_t, double*>> pending_release;  // buffers the side stream may still be reading
  // Factor buffers retired while scoring passes are in flight (look-ahead pipeline): a pass launched before the retirement
  // may still be reading the old factor - and what it computes from it is USED (corrected by the chained rank-n kernels) -
  // so the buffer goes back to the pool only when every pass launched before has been absorbed.
  struct HeldBuf {
    int64_t cap;
    double* buf;
    int64_t tag;  // passes launched when it was retired
  };
  std::vector<HeldBuf> held;
  int64_t pass_launched = 0, pass_absorbed = 0;
  bool side_pending = false;       // some cluster has side_busy set
  double* d_apart2 = nullptr;  // alpha partials of the side stream
  int64_t apart2_elems = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  std::vector<Cluster> cl;
  // patient table
  int64_t P = 0, V = 0;
  std::vector<int64_t> h_off;
  std::vector<double> h_x, h_y, h_ythr;
  int64_t* d_off = nullptr;
  double *d_px = nullptr, *d_py = nullptr;
  // factorisation workspace
  double* d_L = nullptr;
  int64_t L_elems = 0;
  double* d_partial = nullptr;
  int64_t partial_elems = 0;
  double* d_apart = nullptr;  // row-chunk partials of alpha = M^T z
  int64_t apart_elems = 0;
  // scoring scratch
  ScoreRes res_main;        // patient-by-patient steps, predict, append
  ScoreRes res_bump[NBUF];  // re-scoring passes of the look-ahead pipeline (one arena per batch buffer)
  BatchBuf bb[NBUF];        // the batches in flight
  cudaStream_t ahead = nullptr, ahead2 = nullptr;  // low priority: the pass being launched (= ahead_ring[its buffer]) / own-cluster scores
  cudaStream_t ahead_ring[NBUF] = {};              // one look-ahead stream per batch buffer
  cudaEvent_t ev_fork = nullptr, ev_ahead2 = nullptr;
  cudaEvent_t last_wide = nullptr;                 // wide_done of the pass launched last (the next one's wide_after)
  cudaStream_t quick = nullptr, quick2 = nullptr;  // high priority: rank-n score corrections after a move (join / leave side)
  cudaEvent_t ev_q2 = nullptr;
  double* d_algw2 = nullptr;      // partial sums of the join-side correction
  int64_t algw2_elems = 0;
  cudaEvent_t ev_alg = nullptr, ev_apq = nullptr;
  double* d_algw = nullptr;       // partial sums of the leave-side correction
  int64_t algw_elems = 0;
  double* d_apring = nullptr;     // ring of small append workspaces (Li | z_B | Gram partials): several appends in flight
  int apring_next = 0;
  double* d_alg = nullptr;        // their outputs [ALG_OUT_SLOTS][OWN_BATCH positions][2]
  double* h_alg = nullptr;        // ... pinned
  double* d_blog = nullptr;       // logged block columns / alpha_B of the current batch's leave-side changes
  int64_t blog_elems = 0;
  double* d_replay = nullptr;     // outputs of the corrections queued ahead [REPLAY_MAX][OWN_BATCH][2] ...
  double* h_replay = nullptr;     // ... pinned
  double* d_rpw = nullptr;        // ... their partial sums
  int64_t rpw_elems = 0;
  cudaEvent_t ev_replay[NBUF] = {};
  int64_t rpw_share = 0;          // doubles of d_rpw per batch buffer
  AlgHook* alg_hook = nullptr;  // set while a commit should launch the append-side correction (append_core)
  LeaveHook* leave_hook = nullptr;  // ... and the leave-side update of the cached columns (remove_core)
  cudaStream_t ts0 = nullptr;     // high priority: k_T_leave for the current batch's positions
  cudaStream_t ts1[NBUF] = {};    // ... for the positions of a pass in flight (per batch buffer: each waits for its own pass)
  cudaStream_t tq[NBUF] = {};     // high priority: the corrections queued behind a pass in flight
  cudaEvent_t tev[TEV_RING] = {};
  int tev_next = 0;
  double* d_tsw = nullptr;        // ring of private copies of a deletion's operands (Bt | Gt | u | ... | S)
  int64_t tsw_slot = 0;           // doubles per ring entry
  int tsw_next = 0;
  cudaEvent_t tsw_done[TS_RING][NTGT] = {};  // the k_T_leave launches reading entry k have finished
  cudaEvent_t tsw_cp = nullptr;
  cudaEvent_t ev_ts[2 * NBUF] = {};  // batch_launch: the new pass reuses a buffer the kernels on the ts1 / tq streams worked on
  cudaEvent_t ev_ts_end[NBUF] = {};  // [batch buffer] ts0 at the end of the batch it held: buffer and arena are reused NBUF batches later
  double* d_upd2 = nullptr;  // row-deletion workspace of the look-ahead streams (own-cluster scores)
  int64_t upd2_elems = 0;
  double *d_own1 = nullptr, *d_own2 = nullptr;  // Gram partials of the batched own-cluster scores (main-side / look-ahead)
  int64_t own1_elems = 0, own2_elems = 0;
  double* d_wq = nullptr;  // packed query of the whole sweep, in sweep order: x | y | a_draw
  int64_t wq_elems = 0;
  int64_t* d_woff = nullptr;
  int64_t woff_elems = 0;
  double* d_adraw = nullptr;  // the newcomer's randn draw of the current Gibbs step
  double* d_q = nullptr;  // query staging: x | y | a_draw
  int64_t q_elems = 0;
  int64_t* d_qoff = nullptr;
  int64_t qoff_elems = 0;
  double* d_out = nullptr;  // outputs staged on the device before the copy back
  int64_t out_elems = 0;
  // pinned host staging: pin = small read-backs (always synchronised at once); pin_dn = bulk read-backs
  // (same); pin_up = uploads, bump-allocated so that no range is rewritten before the stream has consumed it
  char* pin = nullptr;
  size_t pin_bytes = 0;
  char* pin_dn = nullptr;
  size_t dn_bytes = 0;
  char* pin_up = nullptr;
  size_t up_bytes = 0, up_off = 0;
  Chain chain;
  // lanes of the refit driver (mogp_refit_clusters): light worker engines (one stream + factorisation workspace each),
  // all driven by the caller's thread
  bool light = false;
  double* d_wt = nullptr;  // a lane's copy of the multiplicities of the collapsed cluster it is evaluating
  int64_t wt_elems = 0;
  double* d_ws = nullptr;   // a lane's weight-space workspace (wspace.cuh: WsWork) ...
  int64_t ws_elems = 0;
  double* d_wsC = nullptr;  // ... and the Lagrange basis values of the cluster's inputs [m][WS_R]
  int64_t wsC_elems = 0;
  std::vector<mogp_engine*> lanes;
  std::vector<cudaEvent_t> lane_ev;
  int64_t refit_evals = 0, refit_launched = 0;   // objective evaluations requested / actually run on the device
  // membership probabilities of every step of the last mogp_chain_sweep (self.p[n], mogp_constrained.py:276)
  bool record_probs = false;
  std::vector<int64_t> rec_off;
  std::vector<int32_t> rec_ids;
  std::vector<double> rec_p;
  // update workspace (update.cuh)
  double* d_upd = nullptr;
  int64_t upd_elems = 0;
  double* d_apw = nullptr;   // append workspace
  int64_t apw_elems = 0;
  std::vector<std::pair<int64_t, double*>> pool;  // free cluster buffers by capacity
  // CUDA graph of one objective evaluation (inference + gradient, ~3 nb + 8 kernels), relaunched for every
  // L-BFGS-B evaluation of the same cluster; keyed by the buffers it touches
  cudaGraph_t eval_graph = nullptr;
  cudaGraphExec_t eval_exec = nullptr;
  const void* eval_key[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t eval_kernels = 0;
  bool eval_with_grad = true;  // the graph launched last includes the gradient kernels
  // accounting (bench.py): bytes moved by the C ABI calls, and device time of the scoring GEMM
  int64_t bytes_h2d = 0, bytes_d2h = 0;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop), recycled
  size_t prof_used = 0;
  double prof_union_ms = 0.0;  // union of the profiled whole passes' (start, stop) intervals
  // class 0: whole passes (a batch or a patient against every active cluster, batch prediction); class 1: the
  // re-scoring passes of the look-ahead pipeline (one or two clusters, latency bound) - kept apart in the accounts
  std::vector<char> prof_tag;
  int prof_cls = 0;
  double prof_ms_c[2] = {0.0, 0.0}, prof_bytes_c[2] = {0.0, 0.0}, prof_flops_c[2] = {0.0, 0.0}, prof_pairs = 0.0;
  int64_t prof_launches_c[2] = {0, 0};
};

namespace {

#define FAIL(h, code, ...)                                  \
  do {                                                      \
    char _b[512];                                           \
    snprintf(_b, sizeof _b, __VA_ARGS__);                   \
    (h)->err = _b;                                          \
    return (code);                                          \
  } while (0)

#define CK(h, call)                                                                               \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) FAIL(h, _e == cudaErrorMemoryAllocation ? MOGP_ENOMEM : MOGP_ECUDA,    \
                                "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != MOGP_OK) return _rc; \
  } while (0)

#define LAUNCHED(h) ((h)->launches++)

// MOGP_DEBUG_POISON=1 fills every fresh device buffer with NaN bit patterns, so that a read of memory
// no kernel has written shows up in the results (debugging aid; off by default).
bool poison_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MOGP_DEBUG_POISON");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

template <typename T>
int ensure(mogp_engine* h, T** p, int64_t* have, int64_t need) {
  if (need <= *have) return MOGP_OK;
  if (*p) CK(h, cudaFree(*p));
  *p = nullptr;
  *have = 0;
  int64_t want = need + need / 4;
  CK(h, cudaMalloc((void**)p, sizeof(T) * (size_t)want));
  if (poison_enabled()) CK(h, cudaMemsetAsync(*p, 0xFF, sizeof(T) * (size_t)want, h->stream));
  *have = want;
  return MOGP_OK;
}

int ensure_pin(mogp_engine* h, size_t need) {
  if (need <= h->pin_bytes) return MOGP_OK;
  if (h->pin) CK(h, cudaFreeHost(h->pin));
  h->pin = nullptr;
  h->pin_bytes = 0;
  size_t want = need + need / 2 + 4096;
  CK(h, cudaMallocHost((void**)&h->pin, want));
  h->pin_bytes = want;
  return MOGP_OK;
}

// Every wait on the stream goes through here: afterwards the upload staging can be reused from the start.
cudaError_t sync_stream(mogp_engine* h) {
  cudaError_t e;
  if (h->blocking_sync) {  // worker engines: sleep instead of spinning, many of them share the host cores
    e = cudaEventRecord(h->ev_block, h->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(h->ev_block);
  } else {
    e = cudaStreamSynchronize(h->stream);
  }
  h->up_off = 0;
  return e;
}

// Host wait for everything queued on stream s (one of the engine's helper streams), in the engine's waiting mode.
cudaError_t sync_helper_stream(mogp_engine* h, cudaStream_t s) {
  if (!h->blocking_sync) return cudaStreamSynchronize(s);
  cudaError_t e = cudaEventRecord(h->ev_block, s);
  return e == cudaSuccess ? cudaEventSynchronize(h->ev_block) : e;
}

// Launches made while a SideScope is alive go to the side stream, ordered after everything already queued on
// the main stream.  join_side() makes the main stream wait for them (no host wait) and returns the buffers the
// side stream was still reading to the pool.
struct SideScope {
  mogp_engine* h;
  explicit SideScope(mogp_engine* h_) : h(h_) {
    cudaEventRecord(h->ev_main, h->main_stream);
    cudaStreamWaitEvent(h->side, h->ev_main, 0);
    h->stream = h->side;
  }
  ~SideScope() {
    cudaEventRecord(h->ev_side, h->side);
    h->stream = h->main_stream;
  }
};
void pool_put(mogp_engine* h, int64_t cap, double* buf);
void join_side(mogp_engine* h) {
  cudaStreamWaitEvent(h->main_stream, h->ev_side, 0);
  for (auto& pb : h->pending_release) pool_put(h, pb.first, pb.second);
  h->pending_release.clear();
  if (h->side_pending) {
    for (auto& c : h->cl) c.side_busy = false;
    h->side_pending = false;
  }
}
// Before main-stream work on cluster c: a deletion of its rows queued on the side stream (look-ahead pipeline: the
// main stream does not wait for those by default) must be ordered first.
inline void order_after_side(mogp_engine* h, const Cluster& c) {
  if (c.side_busy) join_side(h);
}

// A fresh range of the upload staging buffer; valid until the next sync_stream().
int stage_alloc(mogp_engine* h, size_t bytes, char** out) {
  bytes = (bytes + 255) / 256 * 256;
  if (h->up_off + bytes > h->up_bytes) {
    CK(h, sync_stream(h));
    if (bytes > h->up_bytes) {
      if (h->pin_up) CK(h, cudaFreeHost(h->pin_up));
      h->pin_up = nullptr;
      h->up_bytes = 0;
      size_t want = std::max<size_t>(bytes * 2, (size_t)1 << 20);
      CK(h, cudaMallocHost((void**)&h->pin_up, want));
      h->up_bytes = want;
    }
  }
  *out = h->pin_up + h->up_off;
  h->up_off += bytes;
  return MOGP_OK;
}

int ensure_dn(mogp_engine* h, size_t need) {
  if (need <= h->dn_bytes) return MOGP_OK;
  if (h->pin_dn) CK(h, cudaFreeHost(h->pin_dn));
  h->pin_dn = nullptr;
  h->dn_bytes = 0;
  size_t want = need + need / 2 + 4096;
  CK(h, cudaMallocHost((void**)&h->pin_dn, want));
  h->dn_bytes = want;
  return MOGP_OK;
}

Theta make_theta(const Cluster& c) {
  Theta th;
  th.A = c.theta[0];
  th.k0 = c.theta[1];
  th.k1 = c.theta[2];
  th.noise = c.theta[3];
  th.kernel = c.spec.kernel;
  th.mean_func = c.spec.mean_func;
  return th;
}

int check_slot(mogp_engine* h, int32_t slot) {  // slot = INDEX into h->cl (internal callers)
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // every cluster entry point comes through here; callers may be on any host thread
  if (slot < 0 || slot >= (int32_t)h->cl.size() || !h->cl[slot].used) FAIL(h, MOGP_EINVAL, "bad cluster slot %d", slot);
  if (h->cl[slot].side_busy && h->stream == h->main_stream) join_side(h);
  return MOGP_OK;
}

// What crosses the ABI is a HANDLE: the slot's index in the low 16 bits, its generation above.  A slot index is reused
// after mogp_cluster_free / a new chain; a caller still holding the old handle (a GPobs view of a cluster that has since
// been emptied, or of a previous chain) gets MOGP_EINVAL instead of silently reading the cluster that now lives there.
constexpr int32_t SLOT_INDEX_BITS = 16, SLOT_INDEX_MASK = (1 << SLOT_INDEX_BITS) - 1, SLOT_GEN_MASK = 0x7FFF;
inline int32_t slot_handle(const mogp_engine* h, int32_t idx) {
  return idx < 0 ? idx : (idx | ((h->cl[idx].gen & SLOT_GEN_MASK) << SLOT_INDEX_BITS));
}
// handle -> index, checked (replaces *slot)
int slot_in(mogp_engine* h, int32_t* slot) {
  if (!h) return MOGP_EINVAL;
  const int32_t hd = *slot;
  const int32_t idx = hd & SLOT_INDEX_MASK;
  if (hd < 0 || idx >= (int32_t)h->cl.size() || !h->cl[idx].used) FAIL(h, MOGP_EINVAL, "bad cluster slot %d", hd);
  if (((hd >> SLOT_INDEX_BITS) & SLOT_GEN_MASK) != (h->cl[idx].gen & SLOT_GEN_MASK))
    FAIL(h, MOGP_EINVAL, "stale cluster handle %d: the slot has been freed and reused since it was handed out", hd);
  *slot = idx;
  return check_slot(h, idx);
}

// Capacity classes (x1.25 steps, multiples of 64) so that buffers can be pooled and ping-ponged.
int64_t cap_class(int64_t need) {
  static int64_t forced = -1;  // MOGP_DEBUG_CAP: force one capacity (debugging aid)
  if (forced < 0) {
    const char* e = getenv("MOGP_DEBUG_CAP");
    forced = e ? atoll(e) : 0;
  }
  if (forced >= need) return round_up(forced, TB);
  int64_t c = TB;
  while (c < need) c = round_up(c + c / 4, TB);
  return c;
}

int pool_get(mogp_engine* h, int64_t cap, double** out);
void pool_put(mogp_engine* h, int64_t cap, double* buf);

int cluster_reserve(mogp_engine* h, Cluster& c, int64_t m_new, bool keep) {
  int64_t need = round_up(std::max<int64_t>(m_new, 1), TB);
  if (need <= c.cap) return MOGP_OK;
  const int64_t cap = cap_class(need + (keep ? need / 8 : 0));
  double* buf = nullptr;
  TRY(pool_get(h, cap, &buf));
  double* M = buf;
  double* vec = buf + cap * cap;
  CK(h, cudaMemsetAsync(vec, 0, sizeof(double) * (size_t)(4 * cap + STAT), h->stream));
  const bool copied = keep && c.cap > 0;
  if (copied) {
    const int64_t mp = round_up(c.m, TB);
    CK(h, cudaMemcpy2DAsync(M, cap * sizeof(double), c.M, c.cap * sizeof(double), mp * sizeof(double), mp,
                            cudaMemcpyDeviceToDevice, h->stream));
    for (int q = 0; q < 4; ++q)
      CK(h, cudaMemcpyAsync(vec + q * cap, c.vec + q * c.cap, sizeof(double) * c.m, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpyAsync(vec + 4 * cap, c.vec + 4 * c.cap, sizeof(double) * STAT, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (c.M) pool_put(h, c.cap, c.M);  // stream-ordered reuse: every consumer runs on h->stream
  c.M = M;
  c.vec = vec;
  c.cap = cap;
  if (!copied) c.theta_dirty = true;  // a fresh buffer holds no device copy of theta
  return MOGP_OK;
}

// Bring the device copy of theta (read by k_build / k_z / k_grad / k_grad_final) up to date.
int sync_theta(mogp_engine* h, Cluster& c) {
  if (!c.theta_dirty) return MOGP_OK;
  char* up = nullptr;
  TRY(stage_alloc(h, sizeof(Theta), &up));
  Theta th = make_theta(c);
  memcpy(up, &th, sizeof(Theta));
  CK(h, cudaMemcpyAsync(c.theta_dev(), up, sizeof(Theta), cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)sizeof(Theta);
  c.theta_dirty = false;
  return MOGP_OK;
}

// alpha = M^T z (two stages) and the log marginal likelihood from the current M and z.
int launch_alpha_lml(mogp_engine* h, Cluster& c) {
  const int64_t mpad = round_up(c.m, TB);
  const int nchunk = (int)((c.m + ALPHA_ROWS - 1) / ALPHA_ROWS);
  TRY(ensure(h, &h->d_apart, &h->apart_elems, (int64_t)nchunk * mpad));
  k_alpha_part<<<dim3((unsigned)(mpad / TB), (unsigned)nchunk), NTHREADS, 0, h->stream>>>(c.M, c.cap, c.z(), (int)c.m,
                                                                                           (int)mpad, h->d_apart);
  LAUNCHED(h);
  k_lml<<<1, 1024, 0, h->stream>>>(c.M, c.cap, c.z(), c.x(), h->d_apart, nchunk, (int)mpad, c.alpha(), (int)c.m, c.stat());
  LAUNCHED(h);
  return MOGP_OK;
}

// Row tiles of the contraction W = M^T M per CTA of the gradient kernel: every chunk CTA repeats the kernel-derivative
// epilogue of its tile pair, so fewer, longer chunks do less redundant work (MOGP_GRAD_KC: tuning).
int grad_kc() {
  static const int v = [] {
    const char* e = getenv("MOGP_GRAD_KC");
    const int k = e ? atoi(e) : GRAD_KC;
    return k >= 1 ? k : GRAD_KC;
  }();
  return v;
}

// Blocked right-looking Cholesky + inverse factor, panels grouped FACTOR_KB at a time (factor.cuh).
void launch_factor(mogp_engine* h, double* L, double* M, int64_t ld, int nb, int* info) {
  static const int kb = [] {
    const char* e = getenv("MOGP_FACTOR_KB");  // panels per group (tuning)
    const int v = e ? atoi(e) : FACTOR_KB;
    return v >= 1 ? v : FACTOR_KB;
  }();
  for (int P0 = 0; P0 < nb; P0 += kb) {
    const int P1 = std::min(P0 + kb, nb);
    for (int p = P0; p < P1; ++p) {
      k_diag<<<1, NTHREADS, DIAG_SMEM, h->stream>>>(L, M, ld, p, info);
      LAUNCHED(h);
      if (nb > 1) {
        k_panel<<<nb - 1, NTHREADS, PANEL_SMEM, h->stream>>>(L, M, ld, p);
        LAUNCHED(h);
      }
      if (p + 1 < nb && !(p == P1 - 1 && P1 < nb)) {  // (the group's last panel leaves everything to k_trail_blk)
        k_trail<<<dim3(nb, nb - p - 1), NTHREADS, TRAIL_SMEM, h->stream>>>(L, M, ld, p, P1);
        LAUNCHED(h);
      }
    }
    if (P1 < nb) {
      // 2 (default): two row tiles per CTA sharing the second operand; 1: one; 4: 128 x 128 blocks (k_trail_big - measured
      // slower here, 898 vs 657 us per evaluation at m = 2 600 and 0.50 vs 0.44 s per 30-cluster refit: one 8-warp CTA
      // per SM with 164 KB of operands leaves the prologue / read-modify-write epilogue of every block uncovered, and
      // its ~200-block launches quantise badly on 148 SMs; kept for experiments, profiles/README.md)
      static const int rows = getenv("MOGP_TRAIL_ROWS") ? atoi(getenv("MOGP_TRAIL_ROWS")) : 2;
      if (rows == 4 && (P1 & 1) == 0)
        k_trail_big<<<dim3((nb + 1) / 2, (nb - P1 + 1) / 2), NTHREADS, TBIG_SMEM, h->stream>>>(L, M, ld, P0, P1, nb);
      else if (rows == 1) k_trail_blk<1><<<dim3(nb, nb - P1), NTHREADS, TRAILB_SMEM, h->stream>>>(L, M, ld, P0, P1, nb);
      else k_trail_blk<2><<<dim3(nb, (nb - P1 + 1) / 2), NTHREADS, TRAILB_SMEM, h->stream>>>(L, M, ld, P0, P1, nb);
      LAUNCHED(h);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Exact inference at (data, theta): ExactGaussianInference.inference as reached by GPy's
// GPRegression.__init__ / set_XY / every optimiser evaluation (obsmodel.py:46,73,92,94).
// jitchol's retry ladder (GPy.util.linalg.jitchol: diag.mean()*1e-6, x10 per try, 5 tries).
int infer(mogp_engine* h, Cluster& c) {
  c.valid = false;
  c.grad_valid = false;
  if (c.m <= 0) FAIL(h, MOGP_ESTATE, "cluster has no data");
  const int64_t mpad = round_up(c.m, TB);
  const int nb = (int)(mpad / TB);
  const int64_t ld = c.cap;
  TRY(ensure(h, &h->d_L, &h->L_elems, ld * mpad));
  TRY(sync_theta(h, c));
  const Theta* th = c.theta_dev();
  double* st = c.stat();
  int* info = reinterpret_cast<int*>(st + 8);
  TRY(ensure_pin(h, 256));
  double jitter = 0.0;
  for (int attempt = 0;; ++attempt) {
    CK(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
    k_build<<<dim3(nb, nb), NTHREADS, 0, h->stream>>>(h->d_L, c.M, ld, c.x(), (int)c.m, th, jitter, c.wt);
    LAUNCHED(h);
    launch_factor(h, h->d_L, c.M, ld, nb, info);
    k_z<<<(unsigned)((c.m + 7) / 8), NTHREADS, 0, h->stream>>>(c.M, ld, c.x(), c.y(), (int)c.m, th, c.z());
    LAUNCHED(h);
    TRY(launch_alpha_lml(h, c));
    CK(h, cudaMemcpyAsync(h->pin, st, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->stream));
    CK(h, sync_stream(h));
    h->bytes_d2h += 72;
    CK(h, cudaGetLastError());
    const double* hs = reinterpret_cast<const double*>(h->pin);
    int hinfo;
    memcpy(&hinfo, hs + 8, sizeof(int));
    if (hinfo == 0 && isfinite(hs[0])) {
      c.lml = hs[0];
      c.quad = hs[1];
      c.logdet = hs[2];
      c.valid = true;
      c.stat_on_device = false;
      c.z_dirty = false;
      c.lml_dirty = false;
      return MOGP_OK;
    }
    // jitchol: fail at once on a non-positive diagonal, else retry with growing jitter
    double dsum = 0.0;
    bool nonpos = false;
    for (int64_t i = 0; i < c.m; ++i) {
      double d = (c.spec.kernel == MOGP_KERNEL_RBF ? c.theta[1] : c.theta[1] * c.hx[i] * c.hx[i] + c.theta[2]) +
                 c.theta[3] + GPY_JITTER;
      if (!(d > 0.0)) nonpos = true;
      dsum += d;
    }
    if (nonpos || attempt >= 5 || !isfinite(dsum))
      FAIL(h, MOGP_ENOTPD, "covariance not positive definite (m=%lld, first bad pivot %d, lml %g, theta %g %g %g %g)",
           (long long)c.m, hinfo, hs[0], c.theta[0], c.theta[1], c.theta[2], c.theta[3]);
    jitter = (attempt == 0) ? dsum / (double)c.m * 1e-6 : jitter * 10.0;
  }
}

// Gradient of the log marginal likelihood w.r.t. natural theta at the current factor.
int fetch_stat(mogp_engine* h, Cluster& c);

int infer_grad(mogp_engine* h, Cluster& c) {
  if (!c.valid) TRY(infer(h, c));
  TRY(fetch_stat(h, c));  // x.alpha of a rank-updated cluster is reduced lazily
  const int64_t mpad = round_up(c.m, TB);
  const int nb = (int)(mpad / TB);
  const int nz = (nb + grad_kc() - 1) / grad_kc();
  const int64_t n_partial = (int64_t)nb * nb * nz;
  TRY(ensure(h, &h->d_partial, &h->partial_elems, n_partial * 3));
  TRY(ensure_pin(h, 256));  // (a worker engine may get here without ever having run an inference itself)
  TRY(sync_theta(h, c));
  const Theta* th = c.theta_dev();
  double* st = c.stat();
  k_grad<<<dim3(nb, nb, nz), NTHREADS, GRAD_SMEM, h->stream>>>(c.M, c.cap, nb, c.x(), c.alpha(), (int)c.m, th,
                                                                h->d_partial, grad_kc(), c.wt);
  LAUNCHED(h);
  k_grad_final<<<1, 256, 0, h->stream>>>(h->d_partial, (int)n_partial, th, st, st + 10);
  LAUNCHED(h);
  CK(h, cudaMemcpyAsync(h->pin, st + 10, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, sync_stream(h));
  CK(h, cudaGetLastError());
  memcpy(c.grad, h->pin, sizeof(double) * 4);
  c.grad_valid = true;
  return MOGP_OK;
}

int upload_data(mogp_engine* h, Cluster& c, int64_t m, const double* x, const double* y) {
  TRY(cluster_reserve(h, c, m, false));
  c.hx.assign(x, x + m);
  c.hy.assign(y, y + m);
  c.m = m;
  char* up = nullptr;
  TRY(stage_alloc(h, sizeof(double) * 2 * m, &up));
  double* stage = reinterpret_cast<double*>(up);
  memcpy(stage, x, sizeof(double) * m);
  memcpy(stage + m, y, sizeof(double) * m);
  CK(h, cudaMemcpyAsync(c.x(), stage, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(c.y(), stage + m, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)sizeof(double) * 2 * m;
  c.valid = false;
  c.grad_valid = false;
  return MOGP_OK;
}

// ---- paramz pieces of the optimiser objective (host scalars) -----------------------
constexpr double LIM_VAL = 36.0;

double logexp_f(double t) {  // paramz Logexp.f
  if (t > LIM_VAL) return t;
  double tc = t < -709.782712893384 ? -709.782712893384 : t;
  return log1p(exp(tc));
}
double logexp_finv(double th) { return th > LIM_VAL ? th : log(expm1(th)); }  // Logexp.finv
double logexp_gradfactor(double th) { return th > LIM_VAL ? 1.0 : -expm1(-th); }

struct Gamma {
  double a, b;
  bool on;
};
// Gamma.from_EV(E, V): a = E^2/V, b = E/V   (obsmodel.py:28,32,41,47,54)
Gamma from_EV(double E, double V) { return Gamma{E * E / V, E / V, true}; }

void priors_of(const mogp_model_spec& s, Gamma pr[4], bool fixed[4]) {
  const Gamma none{0, 0, false};
  pr[0] = s.mean_func ? from_EV(2.0 / 3.0, 0.2) : none;
  fixed[0] = !s.mean_func;  // absent parameter
  pr[1] = s.signal_variance_fix ? none : from_EV(1.0, 0.5);
  fixed[1] = s.signal_variance_fix != 0;
  pr[2] = s.kernel == MOGP_KERNEL_RBF ? from_EV(4.0, 9.0) : none;  // Bias variance: positive, no prior
  fixed[2] = false;
  pr[3] = s.noise_variance_fix ? none : from_EV(0.75, 0.25 * 0.25);
  fixed[3] = s.noise_variance_fix != 0;
}

// ---- scoring --------------------------------------------------------------------------------------------
// Profiling of the dominant kernel (k_score_gemm): CUDA events on the launching stream around every launch.
void prof_drain(mogp_engine* h) {
  if (h->prof_used == 0) return;
  for (size_t i = 1; i < h->prof_used; i += 2) cudaEventSynchronize(h->prof_ev[i]);  // (several streams: every stop event)
  // Two passes of the look-ahead pipeline can be in flight, the later one queued on the SMs behind the earlier one: the
  // time between a launch's start and stop events then includes its wait.  Besides the per-launch sum, keep the UNION of the
  // whole passes' intervals - the time during which the GPU was working on at least one of them.
  std::vector<std::pair<float, float>> iv;
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) != cudaSuccess) continue;
    h->prof_ms_c[(int)h->prof_tag[i]] += ms;
    if (h->prof_tag[i] != 0) continue;
    float t0 = 0.f;
    if (cudaEventElapsedTime(&t0, h->prof_ev[0], h->prof_ev[i]) == cudaSuccess) iv.emplace_back(t0, t0 + ms);
  }
  std::sort(iv.begin(), iv.end());
  float lo = 0.f, hi = -1.f;
  for (const auto& q : iv) {
    if (hi < lo || q.first > hi) {
      if (hi >= lo) h->prof_union_ms += hi - lo;
      lo = q.first;
      hi = q.second;
    } else {
      hi = std::max(hi, q.second);
    }
  }
  if (hi >= lo) h->prof_union_ms += hi - lo;
  h->prof_used = 0;
}
void prof_mark(mogp_engine* h, bool stop) {
  if (!h->prof_on) return;
  if (!stop && h->prof_used + 2 > 8192) prof_drain(h);
  if (h->prof_used >= h->prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->prof_ev.push_back(e);
    h->prof_tag.push_back(0);
  }
  h->prof_tag[h->prof_used] = (char)h->prof_cls;
  cudaEventRecord(h->prof_ev[h->prof_used++], h->stream);
}

// ---- tensor maps for the TMA loads of the scoring kernel --------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {  // the driver entry point, through the runtime (no link against libcuda); null: no TMA path
  static const TmapEncodeFn fn = [] {
    const char* e = getenv("MOGP_TMA");
    if (e && e[0] == '0') return (TmapEncodeFn) nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) {
      cudaGetLastError();
      return (TmapEncodeFn) nullptr;
    }
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// The map of cluster c's factor: a 2-D FP64 tensor (cap x cap, row pitch cap), boxes of 64 rows x LD_K = 72 doubles - the
// padded shared-memory row of k_score_wide, so a box lands exactly as the [64][LD_K] tile its fragment loads expect.
bool cluster_tmap(Cluster& c) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc || !c.M) return false;
  if (c.tmap_M == c.M && c.tmap_cap == c.cap) return true;
  static_assert(sizeof(TensorMap) == sizeof(CUtensorMap), "tensor map size");
  const cuuint64_t dims[2] = {(cuuint64_t)c.cap, (cuuint64_t)c.cap};
  const cuuint64_t strides[1] = {(cuuint64_t)c.cap * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)LD_K, (cuuint32_t)TB}, estr[2] = {1, 1};
  if (enc(reinterpret_cast<CUtensorMap*>(&c.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, c.M, dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    c.tmap_M = nullptr;
    return false;
  }
  c.tmap_M = c.M;
  c.tmap_cap = c.cap;
  return true;
}

template <int BN>
void launch_score_gemm(mogp_engine* h, const ScorePlan& pl, double* Tout, const SlotView* d_slots, const SlotPack& pack,
                       int use_pack, const ChunkPlan& cp) {
  const int smem = ScoreSmem<BN>::BYTES;  // (opted into in mogp_engine_create, per device)
  dim3 grid((unsigned)(pl.Rpad / BN), (unsigned)pl.n_slots, (unsigned)(pl.max_nb * cp.maxch));
  prof_mark(h, false);
  k_score_gemm<BN><<<grid, NTHREADS, smem, h->stream>>>(d_slots, pack, use_pack, pl.kxT, (int)pl.Rpad, pl.part, Tout, cp);
  prof_mark(h, true);
  LAUNCHED(h);
}

template <int NS>
void launch_score_wide(mogp_engine* h, const ScorePlan& pl, double* Tout, const SlotView* d_slots, const SlotPack& pack,
                       int use_pack, const ChunkPlan& cp, const int* colmap = nullptr, const int* ncols = nullptr,
                       const TensorMap* tmaps = nullptr) {
  const int smem = WideSmem<NS>::BYTES;  // (opted into in mogp_engine_create, per device)
  dim3 grid((unsigned)(pl.Rpad / (8 * NS)), (unsigned)pl.n_slots, (unsigned)(pl.max_nb * cp.maxch));
  prof_mark(h, false);
  k_score_wide<NS><<<grid, WIDE_THREADS, smem, h->stream>>>(d_slots, pack, use_pack, pl.kxT, (int)pl.Rpad, pl.part, Tout, cp, colmap,
                                                            ncols, tmaps);
  prof_mark(h, true);
  LAUNCHED(h);
}

// Device-resident query (off_dev[n_pat+1] with off_dev[0] = first column, xs/ys indexed from 0 = that
// column).  Runs k_kx -> k_score_gemm -> k_score_epi; outputs stay on the device.
// prune (optional; batch passes of the look-ahead pipeline): apply the monotonicity rule on the device BEFORE the O(m^2)
// product - pairs it excludes (score -inf) keep only their threshold column in the product (k_prune).
struct PruneSpec {
  const double* ythr_dev;  // [n_pat] the raw cells Y[n, index_to_eval] of the query's patients
  double threshold;
};
int score_device(mogp_engine* h, int64_t n_pat, int64_t R, const int64_t* off_dev, const double* xs_dev,
                 const double* ys_dev, int32_t n_slots, const int32_t* slots, int32_t thr_index, double* score_dev,
                 double* mu_thr_dev, double* col_mean, double* col_var, double* col_lpd, bool keep_T, ScorePlan* plan_out,
                 ScoreRes* res_in = nullptr, const PruneSpec* prune = nullptr) {
  ScoreRes& res = res_in ? *res_in : h->res_main;
  h->prof_cls = res.bump ? 1 : 0;
  if (n_slots <= 0 || n_pat <= 0 || R <= 0) return MOGP_OK;
  ScorePlan pl;
  pl.R = R;
  pl.BN = R <= 64 ? (int)round_up(R, 8) : 64;  // one column tile of 8..64 columns; wider queries: tiles of 64
  pl.Rpad = round_up(R, pl.BN);
  pl.n_slots = n_slots;
  for (int s = 0; s < n_slots; ++s) {  // (inference synchronises: do it before staging anything)
    TRY(check_slot(h, slots[s]));
    if (!h->cl[slots[s]].valid) TRY(infer(h, h->cl[slots[s]]));
  }
  const int use_pack = n_slots <= SLOT_PACK ? 1 : 0;
  SlotPack pack;
  SlotView* hv = pack.v;
  if (!use_pack) {
    if (res_in && res_in != &h->res_main && !res.bump) {  // a pass on the look-ahead stream: own pinned staging
      if (res.hslots_cap < n_slots) {
        if (res.h_slots) CK(h, cudaFreeHost(res.h_slots));
        res.h_slots = nullptr;
        res.hslots_cap = 0;
        CK(h, cudaMallocHost((void**)&res.h_slots, sizeof(SlotView) * (size_t)(n_slots + 16)));
        res.hslots_cap = n_slots + 16;
      }
      hv = res.h_slots;
    } else {
      char* up = nullptr;
      TRY(stage_alloc(h, sizeof(SlotView) * (size_t)n_slots, &up));
      hv = reinterpret_cast<SlotView*>(up);
    }
  }
  for (int s = 0; s < n_slots; ++s) {
    Cluster& c = h->cl[slots[s]];
    SlotView v;
    v.M = c.M;
    v.x = c.x();
    v.alpha = c.alpha();
    v.ld = c.cap;
    v.m = (int)c.m;
    v.mpad = (int)round_up(c.m, TB);
    v.th = make_theta(c);
    v.kx_off = pl.kx_total;
    v.part_off = pl.part_total;
    pl.kx_total += pl.Rpad * (int64_t)v.mpad;
    pl.part_total += (int64_t)(v.mpad / TB) * pl.Rpad;
    pl.max_nb = std::max(pl.max_nb, v.mpad / TB);
    pl.max_rblk = std::max(pl.max_rblk, (v.mpad + KX_ROWS - 1) / KX_ROWS);
    hv[s] = v;
    if (h->prof_on) {  // B_pair = 8 (m^2/2 + 3m) + 16 n + 8 per (patient, cluster); F = n m (m + 10).
      // One launch scores n_pat patients against the cluster: the factor term is charged ONCE per launch (what the
      // launch must move), the per-pair terms per patient; prof_pairs counts the (patient, cluster) units.
      const double m = (double)c.m;
      h->prof_bytes_c[h->prof_cls] += 8.0 * (m * m / 2 + 3 * m) + 8.0 * (double)n_pat + 16.0 * (double)R;
      h->prof_flops_c[h->prof_cls] += (double)R * m * (m + 10.0);
      if (h->prof_cls == 0) h->prof_pairs += (double)n_pat;
    }
  }
  if (h->prof_on) h->prof_launches_c[h->prof_cls] += 1;
  TRY(ensure(h, &res.d_slots, &res.slots_cap, n_slots));
  const int64_t mu_total = (int64_t)n_slots * pl.Rpad, mupart_total = mu_total * pl.max_rblk;
  // few (slot, row block) items and one column tile: split the contraction across CTAs (see ChunkPlan)
  ChunkPlan cp;
  static const int split_kc = getenv("MOGP_RESCORE_KC") ? atoi(getenv("MOGP_RESCORE_KC")) : 8;  // tiles per CTA (0: no split)
  if (split_kc > 0 && pl.Rpad == pl.BN && (int64_t)n_slots * pl.max_nb <= 2 * 148 && pl.max_nb > split_kc) {
    cp.kc = split_kc;
    cp.maxch = (pl.max_nb + cp.kc - 1) / cp.kc;
  }
  cp.zigzag = cp.kc == 0 ? res.zigzag : 0;
  const int64_t tpart_total = cp.kc > 0 ? pl.kx_total * cp.maxch : 0;
  const bool do_prune = prune && thr_index >= 0 && cp.kc == 0 && pl.Rpad == pl.BN && pl.BN >= 24 && n_pat <= PRUNE_MAX_PAT;
  // colmap[n_slots][Rpad] | ncols[n_slots] (ints) | flag[n_pat * n_slots] (bytes), carved out of the same scratch
  const int64_t prune_ints = do_prune ? (int64_t)n_slots * pl.Rpad + n_slots : 0;
  const int64_t prune_total = do_prune ? (prune_ints + 1) / 2 + (n_pat * n_slots + 7) / 8 + 2 : 0;
  const int64_t need = pl.kx_total * (keep_T ? 2 : 1) + pl.part_total + 2 * mu_total + mupart_total + tpart_total + prune_total;
  if (res.bump) {
    if (res.scr_used + need > res.scr_elems) {  // outgrown: earlier blocks are still referenced by the pipeline
      if (res.scr) res.retired.push_back(res.scr);
      res.scr = nullptr;
      res.scr_elems = res.scr_used = 0;
      const int64_t want = std::max<int64_t>(4 * need, (int64_t)1 << 21);
      CK(h, cudaMalloc((void**)&res.scr, sizeof(double) * (size_t)want));
      if (poison_enabled()) CK(h, cudaMemsetAsync(res.scr, 0xFF, sizeof(double) * (size_t)want, h->stream));
      res.scr_elems = want;
    }
    pl.kxT = res.scr + res.scr_used;
    res.scr_used += need;
  } else {
    TRY(ensure(h, &res.scr, &res.scr_elems, need));
    pl.kxT = res.scr;
  }
  pl.part = pl.kxT + pl.kx_total;
  pl.mu = pl.part + pl.part_total;
  pl.mupart = pl.mu + mu_total;
  pl.T = keep_T ? pl.mupart + mupart_total : nullptr;
  cp.Tpart = cp.kc > 0 ? pl.mupart + mupart_total + (keep_T ? pl.kx_total : 0) : nullptr;
  pl.ss = pl.mupart + mupart_total + (keep_T ? pl.kx_total : 0) + tpart_total;
  // (pl.ss has mu_total entries and is the second "mu_total" of `need`; the pruning lists come after it)
  int* d_colmap = do_prune ? reinterpret_cast<int*>(pl.ss + mu_total) : nullptr;
  int* d_ncols = do_prune ? d_colmap + (int64_t)n_slots * pl.Rpad : nullptr;
  unsigned char* d_flag = do_prune ? reinterpret_cast<unsigned char*>(pl.ss + mu_total + (prune_ints + 1) / 2 + 1) : nullptr;
  if (!use_pack) {
    CK(h, cudaMemcpyAsync(res.d_slots, hv, sizeof(SlotView) * (size_t)n_slots, cudaMemcpyHostToDevice, h->stream));
    h->bytes_h2d += (int64_t)sizeof(SlotView) * n_slots;
  }
  // tensor maps of the slots' factors for the wide kernel's TMA loads (cached per cluster, re-encoded when its buffer moves)
  const TensorMap* d_tmaps = nullptr;
  if (pl.BN >= 24 && tmap_encoder()) {
    TensorMap* hm = nullptr;
    if (res_in && res_in != &h->res_main && !res.bump) {  // a pass on the look-ahead stream: own pinned staging
      if (res.htmaps_cap < n_slots) {
        if (res.h_tmaps) CK(h, cudaFreeHost(res.h_tmaps));
        res.h_tmaps = nullptr;
        res.htmaps_cap = 0;
        CK(h, cudaMallocHost((void**)&res.h_tmaps, sizeof(TensorMap) * (size_t)(n_slots + 16)));
        res.htmaps_cap = n_slots + 16;
      }
      hm = res.h_tmaps;
    } else {
      char* up = nullptr;
      TRY(stage_alloc(h, sizeof(TensorMap) * (size_t)n_slots, &up));
      hm = reinterpret_cast<TensorMap*>(up);
    }
    bool ok = true;
    for (int s = 0; s < n_slots && ok; ++s) {
      Cluster& c = h->cl[slots[s]];
      ok = cluster_tmap(c);
      if (ok) hm[s] = c.tmap;
    }
    if (ok) {
      TRY(ensure(h, &res.d_tmaps, &res.tmaps_cap, n_slots));
      CK(h, cudaMemcpyAsync(res.d_tmaps, hm, sizeof(TensorMap) * (size_t)n_slots, cudaMemcpyHostToDevice, h->stream));
      h->bytes_h2d += (int64_t)sizeof(TensorMap) * n_slots;
      d_tmaps = res.d_tmaps;
    }
  }
  // A pass on the look-ahead (lowest priority) stream: its short preamble - the cross covariances and the pruning lists -
  // runs on the high-priority stream instead, so that it does not queue behind everything else on a busy GPU; only the
  // long product itself yields to the chain's urgent kernels.  (MOGP_PRE_HI=0: everything on the look-ahead stream.)
  static const bool pre_hi_on = !(getenv("MOGP_PRE_HI") && getenv("MOGP_PRE_HI")[0] == '0');
  const bool pre_hi = pre_hi_on && res.pre && res_in && res_in != &h->res_main && !res.bump;
  cudaStream_t pre = pre_hi ? res.pre : h->stream;
  if (pre_hi) {
    CK(h, cudaEventRecord(res.ev_pre0, h->stream));
    CK(h, cudaStreamWaitEvent(pre, res.ev_pre0, 0));
  }
  k_kx<<<dim3((unsigned)pl.max_rblk, (unsigned)((pl.Rpad + KX_COLS - 1) / KX_COLS), (unsigned)n_slots), KX_ROWS, 0, pre>>>(
      res.d_slots, pack, use_pack, xs_dev, (int)R, (int)pl.Rpad, pl.max_rblk, pl.kxT, pl.mupart);
  LAUNCHED(h);
  if (do_prune) {
    k_prune<<<(unsigned)n_slots, 64, 0, pre>>>(res.d_slots, pack, use_pack, n_slots, off_dev, (int)n_pat, xs_dev, (int)pl.Rpad,
                                               pl.mupart, pl.max_rblk, thr_index, prune->ythr_dev, prune->threshold, d_colmap,
                                               d_ncols, d_flag);
    LAUNCHED(h);
  }
  if (res.pre_done) CK(h, cudaEventRecord(res.pre_done, pre));
  if (pre_hi) {
    CK(h, cudaEventRecord(res.ev_pre1, pre));
    CK(h, cudaStreamWaitEvent(h->stream, res.ev_pre1, 0));
  }
  if (res.wide_after) CK(h, cudaStreamWaitEvent(h->stream, res.wide_after, 0));
  switch (pl.BN) {  // <= 16 columns: bound by HBM (8-warp ring); wider: bound by the FP64 tensor pipe (16 warps)
    case 8: launch_score_gemm<8>(h, pl, pl.T, res.d_slots, pack, use_pack, cp); break;
    case 16: launch_score_gemm<16>(h, pl, pl.T, res.d_slots, pack, use_pack, cp); break;
    case 24: launch_score_wide<3>(h, pl, pl.T, res.d_slots, pack, use_pack, cp, d_colmap, d_ncols, d_tmaps); break;
    case 32: launch_score_wide<4>(h, pl, pl.T, res.d_slots, pack, use_pack, cp, d_colmap, d_ncols, d_tmaps); break;
    case 40: launch_score_wide<5>(h, pl, pl.T, res.d_slots, pack, use_pack, cp, d_colmap, d_ncols, d_tmaps); break;
    case 48: launch_score_wide<6>(h, pl, pl.T, res.d_slots, pack, use_pack, cp, d_colmap, d_ncols, d_tmaps); break;
    case 56: launch_score_wide<7>(h, pl, pl.T, res.d_slots, pack, use_pack, cp, d_colmap, d_ncols, d_tmaps); break;
    default: launch_score_wide<8>(h, pl, pl.T, res.d_slots, pack, use_pack, cp, d_colmap, d_ncols, d_tmaps); break;
  }
  if (cp.kc > 0) {
    k_score_reduce<<<dim3((unsigned)n_slots, (unsigned)pl.max_nb), 256, 0, h->stream>>>(res.d_slots, pack, use_pack, (int)pl.Rpad, cp,
                                                                                      pl.part, pl.T);
    LAUNCHED(h);
  }
  const int64_t pairs = n_pat * n_slots;
  if (res.wide_done) CK(h, cudaEventRecord(res.wide_done, h->stream));
  if (pre_hi) {  // the short epilogue of a look-ahead pass: high priority again (the host is usually waiting for it)
    CK(h, cudaEventRecord(res.ev_pre0, h->stream));
    CK(h, cudaStreamWaitEvent(pre, res.ev_pre0, 0));
  }
  k_score_epi<<<(unsigned)((pairs + 7) / 8), 256, 0, pre>>>(res.d_slots, pack, use_pack, n_slots, off_dev, n_pat, xs_dev, ys_dev,
                                                                     (int)pl.Rpad, pl.part, pl.mupart, pl.max_rblk, pl.mu,
                                                                     thr_index, score_dev,
                                                                     mu_thr_dev, col_mean, col_var, col_lpd, pl.ss, d_flag);
  LAUNCHED(h);
  if (pre_hi) {
    CK(h, cudaEventRecord(res.ev_pre1, pre));
    CK(h, cudaStreamWaitEvent(h->stream, res.ev_pre1, 0));
  }
  CK(h, cudaGetLastError());
  if (plan_out) *plan_out = pl;
  return MOGP_OK;
}

// Upload a host CSR query into the staging buffers (d_qoff, d_q = x | y | a_draw).
int stage_query(mogp_engine* h, int64_t n_pat, const int64_t* off, const double* x, const double* y, const double* a_draw) {
  if (n_pat <= 0 || !off || !x) FAIL(h, MOGP_EINVAL, "empty query");
  const int64_t base = off[0], R = off[n_pat] - off[0];
  if (R <= 0) FAIL(h, MOGP_EINVAL, "query has no visits");
  for (int64_t p = 0; p < n_pat; ++p)
    if (off[p + 1] < off[p]) FAIL(h, MOGP_EINVAL, "offsets must be non-decreasing");
  const size_t bytes = sizeof(int64_t) * (n_pat + 1) + sizeof(double) * (2 * R + n_pat);
  TRY(ensure(h, &h->d_qoff, &h->qoff_elems, n_pat + 1));
  TRY(ensure(h, &h->d_q, &h->q_elems, 2 * R + n_pat));
  char* up = nullptr;
  TRY(stage_alloc(h, bytes, &up));
  int64_t* so = reinterpret_cast<int64_t*>(up);
  double* sd = reinterpret_cast<double*>(so + n_pat + 1);
  memcpy(so, off, sizeof(int64_t) * (n_pat + 1));
  memcpy(sd, x + base, sizeof(double) * R);
  if (y) memcpy(sd + R, y + base, sizeof(double) * R);
  else memset(sd + R, 0, sizeof(double) * R);
  if (a_draw) memcpy(sd + 2 * R, a_draw, sizeof(double) * n_pat);
  else memset(sd + 2 * R, 0, sizeof(double) * n_pat);
  CK(h, cudaMemcpyAsync(h->d_qoff, so, sizeof(int64_t) * (n_pat + 1), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_q, sd, sizeof(double) * (2 * R + n_pat), cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)bytes;
  return MOGP_OK;
}

int copy_back(mogp_engine* h, double* dst, const double* src_dev, int64_t n) {
  TRY(ensure_dn(h, sizeof(double) * (size_t)n));
  CK(h, cudaMemcpyAsync(h->pin_dn, src_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(h, sync_stream(h));
  h->bytes_d2h += (int64_t)sizeof(double) * n;
  memcpy(dst, h->pin_dn, sizeof(double) * n);
  return MOGP_OK;
}


// ---- cluster slots (internal: indices; the extern "C" wrappers translate handles) ----------------------------------
int upload_data(mogp_engine* h, Cluster& c, int64_t m, const double* x, const double* y);
int cluster_create_i(mogp_engine* h, const mogp_model_spec* spec, const double theta[4], int32_t* slot_out) {
  if (!spec || !theta || !slot_out) FAIL(h, MOGP_EINVAL, "null argument");
  if (spec->kernel != MOGP_KERNEL_RBF && spec->kernel != MOGP_KERNEL_LINEAR) FAIL(h, MOGP_EINVAL, "unknown kernel %d", spec->kernel);
  int32_t slot = -1;
  for (size_t i = 0; i < h->cl.size(); ++i)
    if (!h->cl[i].used) {
      slot = (int32_t)i;
      break;
    }
  if (slot < 0) {
    if ((int32_t)h->cl.size() > SLOT_INDEX_MASK) FAIL(h, MOGP_ENOMEM, "more than %d cluster slots", SLOT_INDEX_MASK + 1);
    h->cl.emplace_back();
    slot = (int32_t)h->cl.size() - 1;
  }
  Cluster& c = h->cl[slot];
  c.used = true;
  c.spec = *spec;
  memcpy(c.theta, theta, sizeof(double) * 4);
  c.theta_dirty = true;
  c.m = 0;
  c.valid = c.grad_valid = false;
  c.hx.clear();
  c.hy.clear();
  c.members.clear();
  *slot_out = slot;
  return MOGP_OK;
}

int cluster_free_i(mogp_engine* h, int32_t slot) {
  TRY(check_slot(h, slot));
  Cluster& c = h->cl[slot];
  if (c.M) pool_put(h, c.cap, c.M);  // stream-ordered reuse by the next cluster of that capacity class
  c.M = c.vec = nullptr;
  c.cap = 0;
  c.used = false;
  c.gen += 1;  // handles to the cluster that lived here are stale from now on
  c.m = 0;
  c.valid = c.grad_valid = false;
  c.hx.clear();
  c.hy.clear();
  c.members.clear();
  return MOGP_OK;
}

int cluster_set_data_i(mogp_engine* h, int32_t slot, int64_t m, const double* x, const double* y) {
  TRY(check_slot(h, slot));
  if (m <= 0 || !x || !y) FAIL(h, MOGP_EINVAL, "cluster data must be non-empty");
  Cluster& c = h->cl[slot];
  c.members.clear();
  TRY(upload_data(h, c, m, x, y));
  return infer(h, c);
}

int cluster_set_members_i(mogp_engine* h, int32_t slot, int64_t n_members, const int64_t* ids) {
  TRY(check_slot(h, slot));
  if (n_members <= 0 || !ids) FAIL(h, MOGP_EINVAL, "member list must be non-empty");
  if (h->P <= 0) FAIL(h, MOGP_ESTATE, "no patient table");
  std::vector<double> xs, ys;
  for (int64_t i = 0; i < n_members; ++i) {
    if (ids[i] < 0 || ids[i] >= h->P) FAIL(h, MOGP_EINVAL, "patient id %lld out of range", (long long)ids[i]);
    xs.insert(xs.end(), h->h_x.begin() + h->h_off[ids[i]], h->h_x.begin() + h->h_off[ids[i] + 1]);
    ys.insert(ys.end(), h->h_y.begin() + h->h_off[ids[i]], h->h_y.begin() + h->h_off[ids[i] + 1]);
  }
  if (xs.empty()) FAIL(h, MOGP_EINVAL, "members have no visits");
  Cluster& c = h->cl[slot];
  TRY(upload_data(h, c, (int64_t)xs.size(), xs.data(), ys.data()));
  c.members.assign(ids, ids + n_members);
  return infer(h, c);
}

// New-cluster marginal of a row too long for k_new_ll's shared-memory Cholesky (more than 64 visits): the general
// cluster path - a temporary slot and exact inference (GPobs(x_n, y_n).ll, obsmodel.py:56, for any row length).
int newcomer_ll_general(mogp_engine* h, const mogp_model_spec& spec, const double th4[4], int64_t n, const double* x,
                        const double* y, double* ll) {
  int32_t slot = -1;
  TRY(cluster_create_i(h, &spec, th4, &slot));
  int rc = upload_data(h, h->cl[slot], n, x, y);
  if (rc == MOGP_OK) rc = infer(h, h->cl[slot]);
  if (rc == MOGP_OK) *ll = h->cl[slot].lml;
  const std::string keep = h->err;
  cluster_free_i(h, slot);
  h->err = keep;
  return rc;
}

}  // namespace

#include "chain.inl"

// =================================================================================================
extern "C" {

int mogp_version(void) { return 100; }

// light = a lane of the refit driver: the main stream and the factorisation workspaces only (no look-ahead streams,
// no append ring, no correction buffers).
static int engine_create_impl(int32_t device, bool light, mogp_engine** out) {
  if (!out) return MOGP_EINVAL;
  *out = nullptr;
  // One hardware work queue per stream of the look-ahead pipeline (default 8: streams share queues, and a stream waiting for
  // a scoring pass blocks the urgent kernels queued behind it).  Effective when this library makes the process' first CUDA
  // call; otherwise the host program sets it before it creates the context.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return MOGP_ECUDA;
  }
  if (device < 0 || device >= n) {
    g_create_error = "device index out of range";
    return MOGP_EINVAL;
  }
  mogp_engine* h = new mogp_engine();
  h->device = device;
  h->light = light;
  int prio_least = 0, prio_greatest = 0;
  if (cudaSetDevice(device) == cudaSuccess) cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  // main / side carry the dependent chain of a Gibbs step (highest priority); the look-ahead streams carry the next
  // batch's scoring pass, which fills the SMs the chain leaves idle (lowest priority)
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      (!light && (cudaStreamCreateWithPriority(&h->ahead2, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
                  cudaStreamCreateWithPriority(&h->quick, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
                  cudaStreamCreateWithPriority(&h->quick2, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
                  cudaMalloc((void**)&h->d_apring, sizeof(double) * AP_RING * AP_RING_SLOT) != cudaSuccess ||
                  cudaMalloc((void**)&h->d_alg, sizeof(double) * ALG_OUT_SLOTS * 2 * OWN_BATCH) != cudaSuccess ||
                  cudaMallocHost((void**)&h->h_alg, sizeof(double) * ALG_OUT_SLOTS * 2 * OWN_BATCH) != cudaSuccess ||
                  cudaMalloc((void**)&h->d_replay, sizeof(double) * NBUF * REPLAY_MAX * 2 * OWN_BATCH) != cudaSuccess ||
                  cudaMallocHost((void**)&h->h_replay, sizeof(double) * NBUF * REPLAY_MAX * 2 * OWN_BATCH) != cudaSuccess ||
                  cudaStreamCreateWithPriority(&h->ts0, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
                  cudaEventCreateWithFlags(&h->tsw_cp, cudaEventDisableTiming) != cudaSuccess)) ||
      cudaEventCreateWithFlags(&h->ev_q2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_alg, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_apq, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_ahead2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_block, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    g_create_error = std::string("engine init failed: ") + cudaGetErrorString(cudaGetLastError());
    delete h;
    return MOGP_ECUDA;
  }
  h->main_stream = h->stream;
  for (int i = 0; i < NBUF; ++i) {
    h->res_bump[i].bump = true;
    bool ok = cudaEventCreateWithFlags(&h->bb[i].done, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->bb[i].done_blk, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
    if (ok && !light)
      ok = cudaEventCreateWithFlags(&h->bb[i].scored, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_replay[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaStreamCreateWithPriority(&h->ahead_ring[i], cudaStreamNonBlocking, prio_least) == cudaSuccess &&
           cudaStreamCreateWithPriority(&h->bb[i].res.pre, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->bb[i].res.ev_pre0, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->bb[i].res.ev_pre1, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->bb[i].res.wide_done, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->bb[i].res.pre_done, cudaEventDisableTiming) == cudaSuccess &&
           cudaStreamCreateWithPriority(&h->tq[i], cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
           cudaStreamCreateWithPriority(&h->ts1[i], cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
    if (!ok) {
      g_create_error = std::string("engine init failed: ") + cudaGetErrorString(cudaGetLastError());
      delete h;
      return MOGP_ECUDA;
    }
  }
  cudaMalloc((void**)&h->d_adraw, sizeof(double) * 8);
  cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM);
  cudaFuncSetAttribute(k_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
  cudaFuncSetAttribute(k_trail, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAIL_SMEM);
  cudaFuncSetAttribute(k_trail_blk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAILB_SMEM);
  cudaFuncSetAttribute(k_trail_blk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAILB_SMEM);
  cudaFuncSetAttribute(k_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM);
  cudaFuncSetAttribute(k_trail_big, cudaFuncAttributeMaxDynamicSharedMemorySize, TBIG_SMEM);
  // (function attributes are per device: set for every kernel whenever an engine is created, on its device)
  cudaFuncSetAttribute(k_score_gemm<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScoreSmem<8>::BYTES);
  cudaFuncSetAttribute(k_score_gemm<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScoreSmem<16>::BYTES);
  cudaFuncSetAttribute(k_score_wide<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideSmem<3>::BYTES);
  cudaFuncSetAttribute(k_score_wide<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideSmem<4>::BYTES);
  cudaFuncSetAttribute(k_score_wide<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideSmem<5>::BYTES);
  cudaFuncSetAttribute(k_score_wide<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideSmem<6>::BYTES);
  cudaFuncSetAttribute(k_score_wide<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideSmem<7>::BYTES);
  cudaFuncSetAttribute(k_score_wide<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideSmem<8>::BYTES);
  update_set_attributes();
  *out = h;
  return MOGP_OK;
}

int mogp_engine_create(int32_t device, mogp_engine** out) { return engine_create_impl(device, false, out); }

void mogp_engine_destroy(mogp_engine* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  sync_stream(h);
  for (mogp_engine* l : h->lanes) mogp_engine_destroy(l);
  h->lanes.clear();
  for (cudaEvent_t e : h->lane_ev) cudaEventDestroy(e);
  h->lane_ev.clear();
  for (auto& c : h->cl)
    if (c.M) cudaFree(c.M);
  cudaFree(h->d_off);
  cudaFree(h->d_px);
  cudaFree(h->d_py);
  cudaFree(h->d_L);
  cudaFree(h->d_partial);
  cudaFree(h->d_apart);
  auto free_res = [](ScoreRes& r) {
    cudaFree(r.scr);
    for (double* p : r.retired) cudaFree(p);
    cudaFree(r.d_slots);
    if (r.h_slots) cudaFreeHost(r.h_slots);
    cudaFree(r.d_tmaps);
    if (r.h_tmaps) cudaFreeHost(r.h_tmaps);
  };
  free_res(h->res_main);
  for (int i = 0; i < NBUF; ++i) {
    free_res(h->res_bump[i]);
    free_res(h->bb[i].res);
    cudaFree(h->bb[i].d_out);
    if (h->bb[i].h_out) cudaFreeHost(h->bb[i].h_out);
    if (h->bb[i].done) cudaEventDestroy(h->bb[i].done);
    if (h->bb[i].scored) cudaEventDestroy(h->bb[i].scored);
    if (h->bb[i].done_blk) cudaEventDestroy(h->bb[i].done_blk);
    if (h->ev_replay[i]) cudaEventDestroy(h->ev_replay[i]);
    if (h->ahead_ring[i]) cudaStreamDestroy(h->ahead_ring[i]);
    if (h->bb[i].res.pre) cudaStreamDestroy(h->bb[i].res.pre);
    for (cudaEvent_t e : {h->bb[i].res.ev_pre0, h->bb[i].res.ev_pre1, h->bb[i].res.wide_done, h->bb[i].res.pre_done})
      if (e) cudaEventDestroy(e);
    if (h->tq[i]) cudaStreamDestroy(h->tq[i]);
    if (h->ts1[i]) cudaStreamDestroy(h->ts1[i]);
  }
  cudaFree(h->d_adraw);
  cudaFree(h->d_wt);
  cudaFree(h->d_ws);
  cudaFree(h->d_wsC);
  cudaFree(h->d_upd2);
  cudaFree(h->d_own1);
  cudaFree(h->d_own2);
  for (cudaEvent_t e : {h->ev_main, h->ev_side, h->ev_block, h->ev_fork, h->ev_ahead2, h->ev_alg, h->ev_apq, h->ev_q2})
    if (e) cudaEventDestroy(e);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->quick) cudaStreamDestroy(h->quick);
  if (h->quick2) cudaStreamDestroy(h->quick2);
  cudaFree(h->d_algw2);
  cudaFree(h->d_alg);
  cudaFree(h->d_blog);
  cudaFree(h->d_replay);
  cudaFree(h->d_rpw);
  if (h->h_replay) cudaFreeHost(h->h_replay);
  if (h->ts0) cudaStreamDestroy(h->ts0);
  if (h->tsw_cp) cudaEventDestroy(h->tsw_cp);
  for (auto& e : h->tev)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_ts)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_ts_end)
    if (e) cudaEventDestroy(e);
  for (auto& e2 : h->tsw_done)
    for (auto& e : e2)
      if (e) cudaEventDestroy(e);
  cudaFree(h->d_tsw);
  cudaFree(h->d_algw);
  cudaFree(h->d_apring);
  if (h->h_alg) cudaFreeHost(h->h_alg);
  if (h->ahead2) cudaStreamDestroy(h->ahead2);
  cudaFree(h->d_wq);
  cudaFree(h->d_woff);
  cudaFree(h->d_q);
  cudaFree(h->d_qoff);
  cudaFree(h->d_out);
  cudaFree(h->d_upd);
  cudaFree(h->d_apw);
  for (auto& pb : h->pool) cudaFree(pb.second);
  for (auto& hb : h->held) cudaFree(hb.buf);
  if (h->pin) cudaFreeHost(h->pin);
  if (h->pin_dn) cudaFreeHost(h->pin_dn);
  if (h->pin_up) cudaFreeHost(h->pin_up);
  if (h->eval_exec) cudaGraphExecDestroy(h->eval_exec);
  if (h->eval_graph) cudaGraphDestroy(h->eval_graph);
  for (auto e : h->prof_ev) cudaEventDestroy(e);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
}

const char* mogp_last_error(const mogp_engine* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int64_t mogp_launch_count(const mogp_engine* h) { return h ? h->launches : 0; }

int mogp_engine_set_blocking_sync(mogp_engine* h, int32_t on) {
  if (!h) return MOGP_EINVAL;
  h->blocking_sync = on != 0;
  return MOGP_OK;
}

int mogp_sync(mogp_engine* h) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  CK(h, sync_stream(h));
  return MOGP_OK;
}
int mogp_timer_start(mogp_engine* h) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  CK(h, cudaEventRecord(h->ev0, h->stream));
  return MOGP_OK;
}
int mogp_timer_stop(mogp_engine* h, double* ms_out) {
  if (!h || !ms_out) return MOGP_EINVAL;
  CK(h, cudaEventRecord(h->ev1, h->stream));
  CK(h, cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *ms_out = ms;
  return MOGP_OK;
}

// ---- patients ---------------------------------------------------------------------------------------------
static int upload_patients(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x, const double* y) {
  const int64_t V = off[n_patients];
  char* up = nullptr;
  TRY(stage_alloc(h, sizeof(int64_t) * (n_patients + 1) + sizeof(double) * 2 * V, &up));
  int64_t* so = reinterpret_cast<int64_t*>(up);
  double* sd = reinterpret_cast<double*>(so + n_patients + 1);
  memcpy(so, off, sizeof(int64_t) * (n_patients + 1));
  memcpy(sd, x, sizeof(double) * V);
  memcpy(sd + V, y, sizeof(double) * V);
  CK(h, cudaMemcpyAsync(h->d_off, so, sizeof(int64_t) * (n_patients + 1), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_px, sd, sizeof(double) * V, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_py, sd + V, sizeof(double) * V, cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)(sizeof(int64_t) * (n_patients + 1) + sizeof(double) * 2 * V);
  return MOGP_OK;
}

int mogp_set_patients(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x, const double* y,
                      const double* y_thr) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n_patients <= 0 || !off || !x || !y || off[0] != 0) FAIL(h, MOGP_EINVAL, "bad patient table");
  for (int64_t p = 0; p < n_patients; ++p)
    if (off[p + 1] < off[p]) FAIL(h, MOGP_EINVAL, "offsets must be non-decreasing");
  const int64_t V = off[n_patients];
  h->P = n_patients;
  h->V = V;
  h->h_off.assign(off, off + n_patients + 1);
  h->h_x.assign(x, x + V);
  h->h_y.assign(y, y + V);
  if (y_thr) h->h_ythr.assign(y_thr, y_thr + n_patients);
  else h->h_ythr.clear();
  cudaFree(h->d_off);
  cudaFree(h->d_px);
  cudaFree(h->d_py);
  h->d_off = nullptr;
  h->d_px = h->d_py = nullptr;
  CK(h, cudaMalloc((void**)&h->d_off, sizeof(int64_t) * (n_patients + 1)));
  CK(h, cudaMalloc((void**)&h->d_px, sizeof(double) * std::max<int64_t>(V, 1)));
  CK(h, cudaMalloc((void**)&h->d_py, sizeof(double) * std::max<int64_t>(V, 1)));
  TRY(upload_patients(h, n_patients, off, x, y));
  CK(h, sync_stream(h));
  h->chain.active = false;
  return MOGP_OK;
}

int mogp_set_patients_matrix(mogp_engine* h, int64_t n_patients, int32_t T, const double* X, const double* Y, int32_t thr_col,
                             int64_t* off_out, double* x_out, double* y_out, int64_t* n_visits_out) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n_patients <= 0 || T <= 0 || !X || !Y) FAIL(h, MOGP_EINVAL, "bad patient matrices");
  const int64_t cells = n_patients * (int64_t)T;
  double *dX = nullptr, *dthr = nullptr;
  int64_t* dcnt = nullptr;
  int* dflag = nullptr;
  struct Guard {
    double *&a, *&b;
    int64_t*& c;
    int*& d;
    ~Guard() {
      cudaFree(a);
      cudaFree(b);
      cudaFree(c);
      cudaFree(d);
    }
  } guard{dX, dthr, dcnt, dflag};
  CK(h, cudaMalloc((void**)&dX, sizeof(double) * 2 * cells));
  CK(h, cudaMalloc((void**)&dthr, sizeof(double) * n_patients));
  CK(h, cudaMalloc((void**)&dcnt, sizeof(int64_t) * n_patients));
  CK(h, cudaMalloc((void**)&dflag, sizeof(int)));
  double* dY = dX + cells;
  CK(h, cudaMemcpyAsync(dX, X, sizeof(double) * cells, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(dY, Y, sizeof(double) * cells, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemsetAsync(dflag, 0, sizeof(int), h->stream));
  h->bytes_h2d += (int64_t)sizeof(double) * 2 * cells;
  cudaFree(h->d_off);
  cudaFree(h->d_px);
  cudaFree(h->d_py);
  h->d_off = nullptr;
  h->d_px = h->d_py = nullptr;
  h->P = 0;
  h->chain.active = false;
  CK(h, cudaMalloc((void**)&h->d_off, sizeof(int64_t) * (n_patients + 1)));
  CK(h, cudaMalloc((void**)&h->d_px, sizeof(double) * cells));
  CK(h, cudaMalloc((void**)&h->d_py, sizeof(double) * cells));
  const unsigned gb = (unsigned)((n_patients + 255) / 256);
  k_pack_count<<<gb, 256, 0, h->stream>>>(dX, dY, n_patients, T, dcnt, dflag);
  LAUNCHED(h);
  k_pack_scan<<<1, 1024, 0, h->stream>>>(dcnt, n_patients, h->d_off);
  LAUNCHED(h);
  k_pack_scatter<<<gb, 256, 0, h->stream>>>(dX, dY, n_patients, T, h->d_off, thr_col, h->d_px, h->d_py, dthr);
  LAUNCHED(h);
  int flag = 0;
  h->h_off.assign(n_patients + 1, 0);
  h->h_ythr.assign(n_patients, 0.0);
  CK(h, cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(h->h_off.data(), h->d_off, sizeof(int64_t) * (n_patients + 1), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(h->h_ythr.data(), dthr, sizeof(double) * n_patients, cudaMemcpyDeviceToHost, h->stream));
  CK(h, sync_stream(h));
  CK(h, cudaGetLastError());
  if (flag) FAIL(h, MOGP_EINVAL, "X and Y have different numbers of non-NaN cells in some row");
  const int64_t V = h->h_off[n_patients];
  h->h_x.assign(V, 0.0);
  h->h_y.assign(V, 0.0);
  if (V > 0) {
    CK(h, cudaMemcpyAsync(h->h_x.data(), h->d_px, sizeof(double) * V, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->h_y.data(), h->d_py, sizeof(double) * V, cudaMemcpyDeviceToHost, h->stream));
    CK(h, sync_stream(h));
  }
  h->bytes_d2h += (int64_t)(sizeof(int64_t) * (n_patients + 1) + sizeof(double) * (2 * V + n_patients));
  h->P = n_patients;
  h->V = V;
  if (thr_col < 0) h->h_ythr.clear();
  if (off_out) memcpy(off_out, h->h_off.data(), sizeof(int64_t) * (n_patients + 1));
  if (x_out && V > 0) memcpy(x_out, h->h_x.data(), sizeof(double) * V);
  if (y_out && V > 0) memcpy(y_out, h->h_y.data(), sizeof(double) * V);
  if (n_visits_out) *n_visits_out = V;
  return MOGP_OK;
}

/* Re-send the same patient table (same offsets) from the host without disturbing the chain: the per-sweep
   host->device copy of an end-to-end run whose data live in host memory. */
int mogp_refresh_patients(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x, const double* y) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n_patients != h->P || !off || !x || !y || off[n_patients] != h->V) FAIL(h, MOGP_EINVAL, "patient table shape changed");
  if (memcmp(off, h->h_off.data(), sizeof(int64_t) * (n_patients + 1)) != 0) FAIL(h, MOGP_EINVAL, "patient offsets changed");
  return upload_patients(h, n_patients, off, x, y);
}

int mogp_profile(mogp_engine* h, int32_t enable) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  prof_drain(h);
  h->prof_on = enable != 0;
  return MOGP_OK;
}

int mogp_profile_read(mogp_engine* h, int64_t* launches, double* ms, double* alg_bytes, double* alg_flops, int32_t reset) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  prof_drain(h);
  if (launches) *launches = h->prof_launches_c[0];
  if (ms) *ms = h->prof_ms_c[0];
  if (alg_bytes) *alg_bytes = h->prof_bytes_c[0];
  if (alg_flops) *alg_flops = h->prof_flops_c[0];
  if (reset) {
    // (the re-scoring class is read through mogp_chain_counters before the reset)
    for (int c = 0; c < 2; ++c) {
      h->prof_launches_c[c] = 0;
      h->prof_ms_c[c] = h->prof_bytes_c[c] = h->prof_flops_c[c] = 0.0;
    }
    h->prof_pairs = 0.0;
    h->prof_union_ms = 0.0;
  }
  return MOGP_OK;
}

int mogp_chain_counters(mogp_engine* h, double* out, int32_t n_out) {
  if (!h || !out) return MOGP_EINVAL;
  prof_drain(h);
  const double v[21] = {(double)h->chain.win_passes, (double)h->chain.win_rescores, (double)h->chain.win_patients, h->prof_pairs,
                        h->chain.t_open, h->chain.t_ensure, h->chain.t_commit,
                        (double)h->prof_launches_c[1], h->prof_ms_c[1], h->prof_bytes_c[1], h->prof_flops_c[1],
                        (double)h->chain.win_algebra, h->chain.t_alg, (double)h->chain.win_replayed,
                        h->chain.max_corr_err, (double)h->chain.corr_checked, (double)h->chain.win_pairs,
                        (double)h->chain.win_pruned, h->chain.t_chain, (double)h->chain.win_chained, h->prof_union_ms};
  for (int i = 0; i < n_out && i < 21; ++i) out[i] = v[i];
  return MOGP_OK;
}

int mogp_transfer_bytes(mogp_engine* h, int64_t* h2d, int64_t* d2h) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (h2d) *h2d = h->bytes_h2d;
  if (d2h) *d2h = h->bytes_d2h;
  return MOGP_OK;
}

// ---- clusters -----------------------------------------------------------------------------------------------
int mogp_cluster_create(mogp_engine* h, const mogp_model_spec* spec, const double theta[4], int32_t* slot_out) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  int32_t idx = -1;
  TRY(cluster_create_i(h, spec, theta, &idx));
  *slot_out = slot_handle(h, idx);
  return MOGP_OK;
}

int mogp_cluster_free(mogp_engine* h, int32_t slot) {
  TRY(slot_in(h, &slot));
  return cluster_free_i(h, slot);
}

int mogp_cluster_set_theta(mogp_engine* h, int32_t slot, const double theta[4]) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  for (int i = 1; i < 4; ++i)
    if (!(theta[i] > 0.0) && !(i == 2 && c.spec.kernel == MOGP_KERNEL_LINEAR && theta[i] >= 0.0))
      FAIL(h, MOGP_EINVAL, "theta[%d] = %g must be positive", i, theta[i]);
  memcpy(c.theta, theta, sizeof(double) * 4);
  c.theta_dirty = true;
  c.valid = c.grad_valid = false;
  if (c.m > 0) return infer(h, c);
  return MOGP_OK;
}

int mogp_cluster_get_theta(mogp_engine* h, int32_t slot, double theta[4]) {
  TRY(slot_in(h, &slot));
  memcpy(theta, h->cl[slot].theta, sizeof(double) * 4);
  return MOGP_OK;
}

int mogp_cluster_set_data(mogp_engine* h, int32_t slot, int64_t m, const double* x, const double* y) {
  TRY(slot_in(h, &slot));
  return cluster_set_data_i(h, slot, m, x, y);
}

int mogp_cluster_set_members(mogp_engine* h, int32_t slot, int64_t n_members, const int64_t* ids) {
  TRY(slot_in(h, &slot));
  return cluster_set_members_i(h, slot, n_members, ids);
}

int mogp_cluster_append(mogp_engine* h, int32_t slot, int64_t n, const double* x, const double* y) {
  TRY(slot_in(h, &slot));
  if (n <= 0 || !x || !y) FAIL(h, MOGP_EINVAL, "nothing to append");
  Cluster& c = h->cl[slot];
  if (c.m == 0) return cluster_set_data_i(h, slot, n, x, y);
  if (!c.valid) TRY(infer(h, c));
  return append_points(h, c, n, x, y);
}

int mogp_cluster_remove_rows(mogp_engine* h, int32_t slot, int64_t p, int64_t n) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  if (p < 0 || n <= 0 || p + n > c.m) FAIL(h, MOGP_EINVAL, "rows [%lld, %lld) outside the cluster (m=%lld)", (long long)p, (long long)(p + n), (long long)c.m);
  if (n == c.m) FAIL(h, MOGP_EINVAL, "cannot remove every row; free the cluster instead");
  if (!c.valid) TRY(infer(h, c));
  c.members.clear();
  int64_t left = n;
  while (left > 0) {  // last chunk first so that p stays valid
    const int nn = (int)std::min<int64_t>(NP, left);
    const int64_t q = p + left - nn;
    TRY(rm_gram(h, c, q, nn));
    TRY(remove_core(h, c, q, nn));
    left -= nn;
  }
  return fetch_stat(h, c);
}

int mogp_cluster_leave_out_score(mogp_engine* h, int32_t slot, int64_t p, int64_t n, int32_t thr_index, double* score,
                                 double* mu_thr) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  if (p < 0 || n <= 0 || n > NP || p + n > c.m || n == c.m) FAIL(h, MOGP_EINVAL, "bad leave-out block");
  if (!c.valid) TRY(infer(h, c));
  TRY(rm_gram(h, c, p, (int)n));
  TRY(own_score(h, c, p, (int)n, c.y() + p, thr_index));
  RmGeom g = make_geom(c, p, (int)n);
  RmWork w;
  TRY(rm_work(h, g, &w));
  double out[3];
  TRY(copy_back(h, out, w.own, 3));
  if (out[2] != 1.0) FAIL(h, MOGP_ENOTPD, "leave-out block not positive definite");
  if (score) *score = out[0];
  if (mu_thr) *mu_thr = out[1];
  return MOGP_OK;
}

int mogp_cluster_get_factor(mogp_engine* h, int32_t slot, double* M_out, double* alpha_out) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  if (!c.valid) TRY(infer(h, c));
  if (M_out) {
    CK(h, cudaMemcpy2DAsync(M_out, c.m * sizeof(double), c.M, c.cap * sizeof(double), c.m * sizeof(double), c.m,
                            cudaMemcpyDeviceToHost, h->stream));
    CK(h, sync_stream(h));
  }
  if (alpha_out) {
    CK(h, cudaMemcpyAsync(alpha_out, c.alpha(), c.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(h, sync_stream(h));
  }
  return MOGP_OK;
}

int mogp_cluster_debug_stat(mogp_engine* h, int32_t slot, double* out32) {
  TRY(slot_in(h, &slot));
  return copy_back(h, out32, h->cl[slot].stat(), STAT);
}

int64_t mogp_cluster_size(mogp_engine* h, int32_t slot) {
  if (slot_in(h, &slot) != MOGP_OK) return -1;
  return h->cl[slot].m;
}

int mogp_cluster_get_data(mogp_engine* h, int32_t slot, double* x, double* y) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  if (x) memcpy(x, c.hx.data(), sizeof(double) * c.m);
  if (y) memcpy(y, c.hy.data(), sizeof(double) * c.m);
  return MOGP_OK;
}

int mogp_cluster_loglik(mogp_engine* h, int32_t slot, double* lml) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  if (!c.valid) TRY(infer(h, c));
  TRY(fetch_stat(h, c));
  *lml = c.lml;
  return MOGP_OK;
}

int mogp_cluster_lml_grad(mogp_engine* h, int32_t slot, double grad_theta[4]) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  if (!c.grad_valid) TRY(infer_grad(h, c));
  memcpy(grad_theta, c.grad, sizeof(double) * 4);
  return MOGP_OK;
}

int mogp_cluster_n_free(mogp_engine* h, int32_t slot) {
  if (slot_in(h, &slot) != MOGP_OK) return MOGP_EINVAL;
  Gamma pr[4];
  bool fixed[4];
  priors_of(h->cl[slot].spec, pr, fixed);
  int n = 0;
  for (int i = 0; i < 4; ++i) n += !fixed[i];
  return n;
}

int mogp_cluster_get_optimizer_array(mogp_engine* h, int32_t slot, double* t, int32_t* n_t) {
  TRY(slot_in(h, &slot));
  Cluster& c = h->cl[slot];
  Gamma pr[4];
  bool fixed[4];
  priors_of(c.spec, pr, fixed);
  int n = 0;
  for (int i = 0; i < 4; ++i)
    if (!fixed[i]) t[n++] = logexp_finv(c.theta[i]);
  if (n_t) *n_t = n;
  return MOGP_OK;
}

bool graphs_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MOGP_NO_GRAPH");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

// One objective evaluation (exact inference + gradient at the cluster's current theta) as a CUDA graph: captured
// once per (cluster buffers, size), relaunched for each of the ~20 evaluations of a refit.  theta travels through
// device memory (sync_theta), so the nodes never change.  Split in two so that one host thread can keep the
// evaluations of many clusters in flight (refit driver below): eval_launch queues the work on h->stream without
// waiting, eval_collect reads the results once the stream has been waited for.  *ok = false when the factorisation
// failed: the caller falls back to the eager path with GPy's jitter ladder.
// with_grad = false: factor, alpha, LML only - the closing f(x_opt) of a refit needs the factor and the objective, not the
// gradient (a third of the evaluation's flops)
static int eval_launch(mogp_engine* h, Cluster& c, bool* launched, bool with_grad = true) {
  *launched = false;
  const int64_t mpad = round_up(c.m, TB);
  const int nb = (int)(mpad / TB);
  const int64_t ld = c.cap;
  const int nz = (nb + grad_kc() - 1) / grad_kc();
  const int64_t n_partial = (int64_t)nb * nb * nz;
  const int nchunk = (int)((c.m + ALPHA_ROWS - 1) / ALPHA_ROWS);
  TRY(ensure(h, &h->d_L, &h->L_elems, ld * mpad));
  TRY(ensure(h, &h->d_apart, &h->apart_elems, (int64_t)nchunk * mpad));
  TRY(ensure(h, &h->d_partial, &h->partial_elems, n_partial * 3));
  TRY(ensure_pin(h, 256));
  TRY(sync_theta(h, c));
  double* st = c.stat();
  const void* key[8] = {c.M, c.vec, h->d_L, h->d_apart, h->d_partial, reinterpret_cast<const void*>((intptr_t)c.m), c.wt,
                        reinterpret_cast<const void*>((intptr_t)(with_grad ? 1 : 0))};
  if (!h->eval_exec || memcmp(key, h->eval_key, sizeof key) != 0) {
    if (h->eval_graph) cudaGraphDestroy(h->eval_graph);
    h->eval_graph = nullptr;
    const int64_t launches0 = h->launches;
    CK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int* info = reinterpret_cast<int*>(st + 8);
    const Theta* th = c.theta_dev();
    cudaMemsetAsync(info, 0, sizeof(int), h->stream);
    k_build<<<dim3(nb, nb), NTHREADS, 0, h->stream>>>(h->d_L, c.M, ld, c.x(), (int)c.m, th, 0.0, c.wt);
    LAUNCHED(h);
    launch_factor(h, h->d_L, c.M, ld, nb, info);
    k_z<<<(unsigned)((c.m + 7) / 8), NTHREADS, 0, h->stream>>>(c.M, ld, c.x(), c.y(), (int)c.m, th, c.z());
    LAUNCHED(h);
    int rc = launch_alpha_lml(h, c);
    if (with_grad) {
      k_grad<<<dim3(nb, nb, nz), NTHREADS, GRAD_SMEM, h->stream>>>(c.M, c.cap, nb, c.x(), c.alpha(), (int)c.m, th, h->d_partial, grad_kc(), c.wt);
      LAUNCHED(h);
      k_grad_final<<<1, 256, 0, h->stream>>>(h->d_partial, (int)n_partial, th, st, st + 10);
      LAUNCHED(h);
    }
    cudaMemcpyAsync(h->pin, st, sizeof(double) * 14, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t ce = cudaStreamEndCapture(h->stream, &h->eval_graph);
    h->eval_kernels = h->launches - launches0;
    h->launches = launches0;  // nothing ran yet
    if (rc != MOGP_OK) return rc;
    if (ce != cudaSuccess || !h->eval_graph) {
      cudaGetLastError();
      if (h->eval_exec) cudaGraphExecDestroy(h->eval_exec);
      h->eval_exec = nullptr;
      return MOGP_OK;  // not launched: eager path
    }
    // the same topology with other buffers (the next cluster of the same tile count): update the executable graph
    // in place, which is much cheaper than instantiating a new one
    bool have = false;
    if (h->eval_exec) {
      cudaGraphExecUpdateResultInfo ui;
      if (cudaGraphExecUpdate(h->eval_exec, h->eval_graph, &ui) == cudaSuccess) {
        have = true;
      } else {
        cudaGetLastError();
        cudaGraphExecDestroy(h->eval_exec);
        h->eval_exec = nullptr;
      }
    }
    if (!have && cudaGraphInstantiate(&h->eval_exec, h->eval_graph, 0) != cudaSuccess) {
      cudaGetLastError();
      h->eval_exec = nullptr;
      return MOGP_OK;
    }
    memcpy(h->eval_key, key, sizeof key);
  }
  CK(h, cudaGraphLaunch(h->eval_exec, h->stream));
  h->eval_with_grad = with_grad;
  *launched = true;
  return MOGP_OK;
}

static int eval_collect(mogp_engine* h, Cluster& c, bool* ok) {
  *ok = false;
  h->launches += h->eval_kernels;
  h->bytes_d2h += 14 * 8;
  const double* hs = reinterpret_cast<const double*>(h->pin);
  int hinfo;
  memcpy(&hinfo, hs + 8, sizeof(int));
  if (hinfo != 0 || !isfinite(hs[0])) return MOGP_OK;  // eager path decides (jitter ladder / not PD)
  c.lml = hs[0];
  c.quad = hs[1];
  c.logdet = hs[2];
  if (h->eval_with_grad) memcpy(c.grad, hs + 10, sizeof(double) * 4);
  c.valid = true;
  c.grad_valid = h->eval_with_grad;
  c.stat_on_device = false;
  c.z_dirty = false;
  c.lml_dirty = false;
  *ok = true;
  return MOGP_OK;
}

// hw supplies the stream and the workspaces, c is the cluster (of hw itself or of another engine on the device).
// objective_begin maps the optimiser coordinates to theta and, when a full evaluation is needed, queues it on
// hw->stream without waiting (*launched); objective_end - after that stream has been waited for - finishes:
// eager fallback with the jitter ladder if the queued factorisation failed, priors, Jacobian, chain rule.
static int objective_begin(mogp_engine* hw, Cluster& c, const double* t, int32_t n_t, bool want_grad, bool* launched,
                           bool grad_in_graph = true) {
  mogp_engine* h = hw;
  *launched = false;
  Gamma pr[4];
  bool fixed[4];
  priors_of(c.spec, pr, fixed);
  int nfree = 0;
  for (int i = 0; i < 4; ++i) nfree += !fixed[i];
  if (t) {
    if (n_t != nfree) FAIL(h, MOGP_EINVAL, "expected %d optimiser coordinates, got %d", nfree, n_t);
    // paramz re-runs inference whenever the optimiser array is set; the result at an unchanged theta is the
    // same, so the factor (and gradient) of the last evaluation are kept when t maps to the current theta -
    // L-BFGS-B's f(x_opt) re-evaluation and the final set (obsmodel.py:94) usually hit this.
    int q = 0;
    bool same = true;
    double th_new[4];
    for (int i = 0; i < 4; ++i) {
      th_new[i] = fixed[i] ? c.theta[i] : logexp_f(t[q++]);
      same = same && th_new[i] == c.theta[i];
    }
    if (!same) {
      memcpy(c.theta, th_new, sizeof th_new);
      c.theta_dirty = true;
      c.valid = c.grad_valid = false;
    }
  }
  if (want_grad && !c.valid && c.m > 0 && graphs_enabled()) TRY(eval_launch(h, c, launched, grad_in_graph));
  return MOGP_OK;
}

static int objective_end(mogp_engine* hw, Cluster& c, bool launched, double* f, double* g) {
  mogp_engine* h = hw;
  Gamma pr[4];
  bool fixed[4];
  priors_of(c.spec, pr, fixed);
  if (launched) {
    bool ok = false;
    TRY(eval_collect(h, c, &ok));
    (void)ok;  // on failure valid stays false and the eager path below runs
  }
  if (!c.valid) TRY(infer(h, c));
  if (g && !c.grad_valid) TRY(infer_grad(h, c));
  // log prior + log Jacobian of the Logexp transform for every parameter that has a prior
  double lp = 0.0, dlp[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    if (!pr[i].on) continue;
    const double th = c.theta[i];
    lp += -lgamma(pr[i].a) + pr[i].a * log(pr[i].b) + (pr[i].a - 1.0) * log(th) - pr[i].b * th;
    lp += logexp_finv(th) - th;
    dlp[i] = (pr[i].a - 1.0) / th - pr[i].b + 1.0 / expm1(th);
  }
  TRY(fetch_stat(h, c));
  // a collapsed shadow: n identical observations = one with noise s / n, times (2 pi s)^-(n-1)/2 n^-1/2 exp(-SS / 2s) per
  // group, s = noise + 1e-8 (zeros for an ordinary cluster)
  const double sn = c.theta[3] + GPY_JITTER;
  const double corr = -(0.5 * c.corr_n1 * log(2.0 * M_PI * sn) + c.corr_logn + 0.5 * c.corr_ss / sn);
  const double dcorr = -(0.5 * c.corr_n1 / sn - 0.5 * c.corr_ss / (sn * sn));
  if (f) *f = -(c.lml + corr) - lp;
  if (g) {
    int q = 0;
    for (int i = 0; i < 4; ++i)
      if (!fixed[i]) g[q++] = -(c.grad[i] + (i == 3 ? dcorr : 0.0) + dlp[i]) * logexp_gradfactor(c.theta[i]);
  }
  return MOGP_OK;
}

static int objective_impl(mogp_engine* hw, Cluster& c, const double* t, int32_t n_t, double* f, double* g) {
  bool launched = false;
  TRY(objective_begin(hw, c, t, n_t, g != nullptr, &launched));
  if (launched) CK(hw, sync_stream(hw));
  return objective_end(hw, c, launched, f, g);
}

int mogp_cluster_objective(mogp_engine* h, int32_t slot, const double* t, int32_t n_t, double* f, double* g) {
  TRY(slot_in(h, &slot));
  return objective_impl(h, h->cl[slot], t, n_t, f, g);
}

int mogp_cluster_objective_on(mogp_engine* worker, mogp_engine* owner, int32_t slot, const double* t, int32_t n_t, double* f,
                              double* g) {
  if (!worker || !owner) return MOGP_EINVAL;
  cudaSetDevice(worker->device);
  if (worker->device != owner->device) FAIL(worker, MOGP_EINVAL, "worker and owner engines must be on the same device");
  // (the owner is only read here: no check_slot / join_side on it from a worker's thread)
  const int32_t idx = slot & SLOT_INDEX_MASK;
  if (slot < 0 || idx >= (int32_t)owner->cl.size() || !owner->cl[idx].used ||
      ((slot >> SLOT_INDEX_BITS) & SLOT_GEN_MASK) != (owner->cl[idx].gen & SLOT_GEN_MASK))
    FAIL(worker, MOGP_EINVAL, "bad cluster slot %d", slot);
  return objective_impl(worker, owner->cl[idx], t, n_t, f, g);
}

#include "refit.inl"

// ---- k-means initialisation (mogp_constrained.py:182-207), see kmeans.cuh ------------------------------------------
int mogp_kmeans_init(mogp_engine* h, int64_t n, const double* x, int32_t K, int32_t n_init, int32_t max_iter, double tol,
                     const double* u, int64_t n_u, int32_t* labels_out, double* centers_out, double* inertia_out,
                     int32_t* best_run_out, int32_t* n_iter_out) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n <= 0 || !x || K <= 0 || K > n || K > 4096 || n_init <= 0 || max_iter <= 0 || !u || !labels_out)
    FAIL(h, MOGP_EINVAL, "bad k-means arguments (n=%lld, K=%d)", (long long)n, K);
  const int trials = 2 + (int)log((double)K);  // n_local_trials of sklearn's _kmeans_plusplus
  const int64_t per = 1 + (int64_t)(K - 1) * trials;
  if (trials > 16 || n_u < per * n_init) FAIL(h, MOGP_EINVAL, "k-means needs %lld uniforms, got %lld", (long long)(per * n_init), (long long)n_u);
  // first centre of every run: RandomState.choice(n, p = 1/n): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
  std::vector<int64_t> first(n_init);
  {
    std::vector<double> cdf(n);
    const double p = 1.0 / (double)n;
    double run = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      run += p;
      cdf[i] = run;
    }
    const double last = cdf[n - 1];
    for (int64_t i = 0; i < n; ++i) cdf[i] /= last;
    for (int r = 0; r < n_init; ++r) {
      const double uv = u[(int64_t)r * per];
      first[r] = std::min<int64_t>(n - 1, std::upper_bound(cdf.begin(), cdf.end(), uv) - cdf.begin());
    }
  }
  double *d_x = nullptr, *d_u = nullptr;
  int64_t* d_first = nullptr;
  int32_t* d_lab = nullptr;
  struct Guard {
    double *&a, *&b;
    int64_t*& c;
    int32_t*& d;
    ~Guard() {
      cudaFree(a);
      cudaFree(b);
      cudaFree(c);
      cudaFree(d);
    }
  } guard{d_x, d_u, d_first, d_lab};
  // d_x: x | xc | stat[2] | closest[n_init][n] | centers[n_init][K] | inertia[n_init]
  const int64_t nx = 2 * n + 2 + (int64_t)n_init * n + (int64_t)n_init * K + n_init;
  CK(h, cudaMalloc((void**)&d_x, sizeof(double) * nx));
  CK(h, cudaMalloc((void**)&d_u, sizeof(double) * per * n_init));
  CK(h, cudaMalloc((void**)&d_first, sizeof(int64_t) * n_init));
  CK(h, cudaMalloc((void**)&d_lab, sizeof(int32_t) * (2 * (int64_t)n_init * n + n_init)));
  double* d_xc = d_x + n;
  double* d_stat = d_xc + n;
  double* d_closest = d_stat + 2;
  double* d_centers = d_closest + (int64_t)n_init * n;
  double* d_inertia = d_centers + (int64_t)n_init * K;
  int32_t* d_lab_old = d_lab + (int64_t)n_init * n;
  int32_t* d_iters = d_lab_old + (int64_t)n_init * n;
  CK(h, cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(d_u, u, sizeof(double) * per * n_init, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(d_first, first.data(), sizeof(int64_t) * n_init, cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)(sizeof(double) * (n + per * n_init) + sizeof(int64_t) * n_init);
  k_kmeans_center<<<1, KM_THREADS, 0, h->stream>>>(d_x, n, d_xc, d_stat);
  LAUNCHED(h);
  k_kmeans_seed<<<n_init, KM_THREADS, 0, h->stream>>>(d_xc, n, K, trials, d_u, per, d_first, d_closest, d_centers);
  LAUNCHED(h);
  const int S = K >= KM_THREADS ? 1 : KM_THREADS / K;
  const size_t smem = sizeof(double) * (3 * (size_t)K + 2 * (size_t)S * K);
  if (smem > 48 * 1024) CK(h, cudaFuncSetAttribute(k_kmeans_lloyd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_kmeans_lloyd<<<n_init, KM_THREADS, smem, h->stream>>>(d_xc, n, K, max_iter, d_stat, tol, d_centers, d_lab, d_lab_old, d_inertia,
                                                          d_iters);
  LAUNCHED(h);
  CK(h, cudaGetLastError());
  std::vector<int32_t> lab((size_t)n_init * n), iters(n_init);
  std::vector<double> inertia(n_init), centers((size_t)n_init * K);
  double stat[2];
  CK(h, cudaMemcpyAsync(lab.data(), d_lab, sizeof(int32_t) * (size_t)n_init * n, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(iters.data(), d_iters, sizeof(int32_t) * n_init, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(inertia.data(), d_inertia, sizeof(double) * n_init, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(centers.data(), d_centers, sizeof(double) * (size_t)n_init * K, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(stat, d_stat, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
  CK(h, sync_stream(h));
  h->bytes_d2h += (int64_t)(sizeof(int32_t) * ((size_t)n_init * n + n_init) + sizeof(double) * (n_init + (size_t)n_init * K + 2));
  // KMeans.fit: keep a run if its inertia is lower AND its clustering is not the same up to a relabelling
  auto same_clustering = [&](const int32_t* a, const int32_t* b) {
    std::vector<int32_t> map(K, -1);
    for (int64_t i = 0; i < n; ++i) {
      if (map[a[i]] == -1) map[a[i]] = b[i];
      else if (map[a[i]] != b[i]) return false;
    }
    return true;
  };
  int best = 0;
  for (int r = 1; r < n_init; ++r)
    if (inertia[r] < inertia[best] && !same_clustering(lab.data() + (size_t)r * n, lab.data() + (size_t)best * n)) best = r;
  memcpy(labels_out, lab.data() + (size_t)best * n, sizeof(int32_t) * n);
  if (centers_out)
    for (int j = 0; j < K; ++j) centers_out[j] = centers[(size_t)best * K + j] + stat[0];  // back on the data's scale
  if (inertia_out) *inertia_out = inertia[best];
  if (best_run_out) *best_run_out = best;
  if (n_iter_out) *n_iter_out = iters[best];
  return MOGP_OK;
}

int mogp_cluster_log_predictive_density(mogp_engine* h, int32_t slot, int64_t n, const double* xs, const double* ys,
                                        double* lpd) {
  TRY(slot_in(h, &slot));
  if (n <= 0 || !xs || !ys || !lpd) FAIL(h, MOGP_EINVAL, "bad query");
  int64_t off[2] = {0, n};
  TRY(stage_query(h, 1, off, xs, ys, nullptr));
  ScorePlan pl;
  TRY(ensure(h, &h->d_out, &h->out_elems, 3 * round_up(n, 64) + 8));
  const int64_t Rp = round_up(n, 64);
  TRY(score_device(h, 1, n, h->d_qoff, h->d_q, h->d_q + n, 1, &slot, -1, nullptr, nullptr, nullptr, nullptr, h->d_out, false, &pl));
  (void)Rp;
  return copy_back(h, lpd, h->d_out, n);
}

int mogp_cluster_predict(mogp_engine* h, int32_t slot, int64_t n, const double* xs, double* mean, double* var) {
  TRY(slot_in(h, &slot));
  if (n <= 0 || !xs) FAIL(h, MOGP_EINVAL, "bad query");
  int64_t off[2] = {0, n};
  TRY(stage_query(h, 1, off, xs, nullptr, nullptr));
  ScorePlan pl;
  const int64_t Rp = round_up(n, 64);
  TRY(ensure(h, &h->d_out, &h->out_elems, 2 * Rp + 8));
  TRY(score_device(h, 1, n, h->d_qoff, h->d_q, nullptr, 1, &slot, -1, nullptr, nullptr, h->d_out, h->d_out + Rp, nullptr, false, &pl));
  if (mean) TRY(copy_back(h, mean, h->d_out, n));
  if (var) TRY(copy_back(h, var, h->d_out + Rp, n));
  return MOGP_OK;
}

// ---- batched pair scoring ------------------------------------------------------------------------------------------
int mogp_score_pairs_dev(mogp_engine* h, int64_t n_patients, int64_t n_visits, const int64_t* off_dev, const double* x_dev,
                         const double* y_dev, int32_t n_slots, const int32_t* slots, int32_t thr_index, double* score_dev,
                         double* mu_thr_dev) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n_patients <= 0 || n_visits <= 0 || !off_dev || !x_dev || !y_dev || n_slots <= 0 || !slots || !score_dev)
    FAIL(h, MOGP_EINVAL, "bad arguments");
  std::vector<int32_t> idx(slots, slots + n_slots);
  for (int32_t& q : idx) TRY(slot_in(h, &q));
  return score_device(h, n_patients, n_visits, off_dev, x_dev, y_dev, n_slots, idx.data(), thr_index, score_dev, mu_thr_dev,
                      nullptr, nullptr, nullptr, false, nullptr);
}

int mogp_score_pairs(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x, const double* y, int32_t n_slots,
                     const int32_t* slots, int32_t thr_index, double* score, double* mu_thr) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n_patients <= 0 || !off || !x || !y || n_slots <= 0 || !slots || !score) FAIL(h, MOGP_EINVAL, "bad arguments");
  // Chunk the patients so the cross-covariance scratch stays bounded (~2 GiB).
  std::vector<int32_t> idx(slots, slots + n_slots);
  for (int32_t& q : idx) TRY(slot_in(h, &q));
  slots = idx.data();
  int64_t msum = 0;
  for (int s = 0; s < n_slots; ++s) msum += round_up(h->cl[slots[s]].m, TB);
  const int64_t max_cols = std::max<int64_t>(64, ((int64_t)1 << 28) / std::max<int64_t>(msum, 1));
  int64_t p0 = 0;
  while (p0 < n_patients) {
    int64_t p1 = p0 + 1;
    while (p1 < n_patients && off[p1 + 1] - off[p0] <= max_cols) ++p1;
    const int64_t np = p1 - p0, R = off[p1] - off[p0];
    if (R <= 0) {  // patients without visits score 0 (empty sum)
      for (int64_t q = p0 * n_slots; q < p1 * n_slots; ++q) {
        score[q] = 0.0;
        if (mu_thr) mu_thr[q] = NAN;
      }
      p0 = p1;
      continue;
    }
    TRY(stage_query(h, np, off + p0, x, y, nullptr));
    const int64_t pairs = np * n_slots;
    TRY(ensure(h, &h->d_out, &h->out_elems, 2 * pairs + np));
    // MOGP_PROBE_PRUNE=k (diagnostic, timing probes of the pruned product on its own): every k-th patient is treated as
    // excluded by the monotonicity rule under every cluster (its score comes back -inf)
    static const int probe = getenv("MOGP_PROBE_PRUNE") ? atoi(getenv("MOGP_PROBE_PRUNE")) : 0;
    PruneSpec ps{h->d_out + 2 * pairs, 0.0};
    if (probe > 0 && thr_index >= 0) {
      char* up = nullptr;
      TRY(stage_alloc(h, sizeof(double) * np, &up));
      double* yt = reinterpret_cast<double*>(up);
      for (int64_t q = 0; q < np; ++q) yt[q] = (q % probe == 0) ? 1e300 : -1e300;
      CK(h, cudaMemcpyAsync(h->d_out + 2 * pairs, yt, sizeof(double) * np, cudaMemcpyHostToDevice, h->stream));
    }
    TRY(score_device(h, np, R, h->d_qoff, h->d_q, h->d_q + R, n_slots, slots, thr_index, h->d_out,
                     mu_thr ? h->d_out + pairs : nullptr, nullptr, nullptr, nullptr, false, nullptr, nullptr,
                     (probe > 0 && thr_index >= 0) ? &ps : nullptr));
    TRY(copy_back(h, score + p0 * n_slots, h->d_out, pairs));
    if (mu_thr) TRY(copy_back(h, mu_thr + p0 * n_slots, h->d_out + pairs, pairs));
    p0 = p1;
  }
  return MOGP_OK;
}

int mogp_new_cluster_ll(mogp_engine* h, const mogp_model_spec* spec, const double theta[4], int64_t n_patients, const int64_t* off,
                        const double* x, const double* y, const double* a_draw, double* ll) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!spec || !theta || n_patients <= 0 || !off || !x || !y || !ll) FAIL(h, MOGP_EINVAL, "bad arguments");
  if (spec->mean_func && !a_draw) FAIL(h, MOGP_EINVAL, "mean_func needs the randn draws of the NegLinear mapping");
  bool any_long = false;
  for (int64_t p = 0; p < n_patients; ++p) {
    if (off[p + 1] - off[p] <= 0) FAIL(h, MOGP_EINVAL, "new-cluster model needs at least one visit per patient");
    any_long = any_long || off[p + 1] - off[p] > 64;
  }
  TRY(stage_query(h, n_patients, off, x, y, a_draw));
  const int64_t R = off[n_patients] - off[0];
  Theta th;
  th.A = 0.0;
  th.k0 = theta[1];
  th.k1 = theta[2];
  th.noise = theta[3];
  th.kernel = spec->kernel;
  th.mean_func = spec->mean_func;
  TRY(ensure(h, &h->d_out, &h->out_elems, n_patients));
  k_new_ll<<<(unsigned)n_patients, 64, 0, h->stream>>>(h->d_qoff, h->d_q, h->d_q + R, h->d_q + 2 * R, th, h->d_out);
  LAUNCHED(h);
  CK(h, cudaGetLastError());
  TRY(copy_back(h, ll, h->d_out, n_patients));
  if (any_long)  // rows longer than the kernel's shared-memory tile (64 visits): the general cluster path
    for (int64_t p = 0; p < n_patients; ++p) {
      const int64_t n = off[p + 1] - off[p];
      if (n <= 64) continue;
      const double th4[4] = {spec->mean_func ? fabs(a_draw[p]) : 0.0, theta[1], theta[2], theta[3]};
      int rc = newcomer_ll_general(h, *spec, th4, n, x + off[p], y + off[p], &ll[p]);
      if (rc == MOGP_ENOTPD) ll[p] = NAN;
      else if (rc != MOGP_OK) return rc;
    }
  return MOGP_OK;
}

// ---- L-BFGS-B state machine, reverse communication (tests against SciPy; the refit driver uses the same class) ----
struct mogp_lbfgsb {
  Lbfgsb opt;
};
int mogp_lbfgsb_create(int32_t n, int32_t m, double factr, double pgtol, int32_t maxls, const double* x0, mogp_lbfgsb** out) {
  if (!out || !x0 || n <= 0 || n > Lbfgsb::NMAX || m <= 0 || m > Lbfgsb::MMAX || maxls <= 0) return MOGP_EINVAL;
  mogp_lbfgsb* s = new mogp_lbfgsb();
  s->opt.init(n, x0, m, factr, pgtol, maxls);
  *out = s;
  return MOGP_OK;
}
int mogp_lbfgsb_step(mogp_lbfgsb* s, double* x, double* f, double* g) {
  if (!s || !x || !f || !g) return MOGP_EINVAL;
  Lbfgsb& o = s->opt;
  if (o.task == Lbfgsb::T_FG) {
    o.f = *f;
    memcpy(o.g, g, sizeof(double) * o.n);
  }
  const int task = (int)o.step();
  memcpy(x, o.x, sizeof(double) * o.n);
  *f = o.f;
  memcpy(g, o.g, sizeof(double) * o.n);
  return task;
}
void mogp_lbfgsb_destroy(mogp_lbfgsb* s) { delete s; }

}  // extern "C"
