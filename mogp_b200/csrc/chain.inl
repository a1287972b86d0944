// chain.inl — host orchestration of the structural updates and of the collapsed Gibbs step
// (mogp/mogp_constrained.py:231-284 + mogp/allocmodel.py:22-34).  Included by engine.cu.
namespace {

int env_int_once(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// ---- cached columns kept current on the device (look-ahead pipeline, MOGP_TMAINT) ----
// A move changes two clusters.  The columns T_q = M k_q (and ss, mu) the scoring passes left on the device for the patients
// still ahead are brought to the changed cluster by O(m n) kernels - k_alg_leave_fin / k_alg_join_fin write ss', mu' (and the
// rows a join adds) back, k_T_leave applies a deletion's row operations - instead of being recomputed by an O(m^2) product
// against the updated factor.  Readers and writers of the columns of one (batch, slot) run on several streams; each waits
// for the event of the one before it (t_wait) and leaves its own (t_mark).
// Default: on for an engine that has the GPU to itself.  An engine in blocking-sync mode is one of many chains sharing a GPU
// (bench.py's 64-chain leg, 8 host threads per process): the other chains already fill the SMs while this one waits, its
// clusters are small (a re-scoring pass is cheap) and the extra streams and kernels cost more than they save - measured
// 5.2 chain-sweeps/s with columns kept current and two passes in flight, 8.1 without.
bool tmaint_on(const mogp_engine* h) {  // (read per call: tests toggle it)
  const char* e = getenv("MOGP_TMAINT");
  return e ? e[0] != '0' : !h->blocking_sync;
}
int t_wait(mogp_engine* h, cudaStream_t s, int bi, int32_t slot) {
  const std::vector<int>& v = h->chain.win.slot_tev[bi];
  if (slot < 0 || (size_t)slot >= v.size() || v[slot] < 0) return MOGP_OK;
  CK(h, cudaStreamWaitEvent(s, h->tev[v[slot]], 0));
  return MOGP_OK;
}
int t_mark(mogp_engine* h, cudaStream_t s, int bi, int32_t slot) {
  std::vector<int>& v = h->chain.win.slot_tev[bi];
  if (slot < 0) return MOGP_OK;
  if ((size_t)slot >= v.size()) v.resize((size_t)slot + 1, -1);
  const int k = h->tev_next;
  h->tev_next = (k + 1) % TEV_RING;
  if (!h->tev[k]) CK(h, cudaEventCreateWithFlags(&h->tev[k], cudaEventDisableTiming));
  CK(h, cudaEventRecord(h->tev[k], s));
  v[slot] = k;
  return MOGP_OK;
}

// remove_core, right after k_rm_prep: k_T_leave for the targets the step asked for, from a private copy of the deletion's
// operands (the workspace itself is rewritten by the next patient's Gram).
int fire_leave_hook(mogp_engine* h, const RmGeom& g, const RmWork& w, int nT) {
  LeaveHook* hk = h->leave_hook;
  h->leave_hook = nullptr;
  bool any = false;
  for (int t = 0; t < NTGT; ++t) any = any || hk->npos[t] > 0;
  if (!any) return MOGP_OK;
  const int64_t used = (w.S + (int64_t)nT * TB * TB) - w.Bt;
  if (!h->d_tsw || used > h->tsw_slot) return MOGP_OK;  // not fired: the step falls back to re-scoring
  const int k = h->tsw_next;
  h->tsw_next = (k + 1) % TS_RING;
  for (int t = 0; t < NTGT; ++t)
    if (h->tsw_done[k][t]) CK(h, cudaStreamWaitEvent(h->stream, h->tsw_done[k][t], 0));
  double* dst = h->d_tsw + (int64_t)k * h->tsw_slot;
  CK(h, cudaMemcpyAsync(dst, w.Bt, sizeof(double) * (size_t)used, cudaMemcpyDeviceToDevice, h->stream));
  CK(h, cudaEventRecord(h->tsw_cp, h->stream));
  RmWork wc = w;
  wc.Bt = dst;
  wc.Gt = dst + (w.Gt - w.Bt);
  wc.u = dst + (w.u - w.Bt);
  wc.Pgram = dst + (w.Pgram - w.Bt);
  wc.Gc = dst + (w.Gc - w.Bt);
  wc.own = dst + (w.own - w.Bt);
  wc.gsol = dst + (w.gsol - w.Bt);
  wc.S = dst + (w.S - w.Bt);
  for (int t = 0; t < NTGT; ++t) {
    if (hk->npos[t] <= 0) continue;
    cudaStream_t s = t == 0 ? h->ts0 : h->ts1[hk->bi[t]];
    CK(h, cudaStreamWaitEvent(s, h->tsw_cp, 0));
    TRY(t_wait(h, s, hk->bi[t], hk->slot));
    k_T_leave<<<(unsigned)hk->npos[t], NTHREADS, RM_STRIP_SMEM, s>>>(hk->J[t], g, wc);
    LAUNCHED(h);
    TRY(t_mark(h, s, hk->bi[t], hk->slot));
    if (!h->tsw_done[k][t]) CK(h, cudaEventCreateWithFlags(&h->tsw_done[k][t], cudaEventDisableTiming));
    CK(h, cudaEventRecord(h->tsw_done[k][t], s));
    hk->fired[t] = true;
  }
  return MOGP_OK;
}

// ---- buffer pool: clusters own one buffer of cap*cap + 4*cap + 16 doubles; row deletions ping-pong ----
int pool_get(mogp_engine* h, int64_t cap, double** out) {
  for (size_t i = 0; i < h->pool.size(); ++i)
    if (h->pool[i].first == cap) {
      *out = h->pool[i].second;
      h->pool.erase(h->pool.begin() + i);
      if (poison_enabled()) CK(h, cudaMemsetAsync(*out, 0xFF, sizeof(double) * (size_t)(cap * cap + 4 * cap + STAT), h->stream));
      return MOGP_OK;
    }
  CK(h, cudaMalloc((void**)out, sizeof(double) * (size_t)(cap * cap + 4 * cap + STAT)));
  if (poison_enabled()) CK(h, cudaMemsetAsync(*out, 0xFF, sizeof(double) * (size_t)(cap * cap + 4 * cap + STAT), h->stream));
  return MOGP_OK;
}
void pool_put(mogp_engine* h, int64_t cap, double* buf) {
  if (!buf) return;
  // a pass in flight may be reading it (see HeldBuf); without chained corrections what a pass computes for a cluster that
  // changed meanwhile is discarded, and the buffer can go round at once (stream order covers everything that is used)
  if (h->chain.win.open && h->pass_launched > h->pass_absorbed && tmaint_on(h)) {
    h->held.push_back({cap, buf, h->pass_launched});
    return;
  }
  h->pool.emplace_back(cap, buf);
}
// after a pass has been absorbed (or when nothing is in flight any more: all = true)
void pool_unhold(mogp_engine* h, bool all) {
  size_t keep = 0;
  for (size_t i = 0; i < h->held.size(); ++i) {
    if (all || h->held[i].tag <= h->pass_absorbed) h->pool.emplace_back(h->held[i].cap, h->held[i].buf);
    else h->held[keep++] = h->held[i];
  }
  h->held.resize(keep);
}

// After a rank-n update alpha is kept current in closed form by the update itself (k_append_finalize / k_rm_alpha);
// z (after a deletion) and the log marginal are recomputed from the current M when somebody asks (fetch_stat).
int mark_updated(mogp_engine* h, Cluster& c, bool z_stale) {
  (void)h;
  if (z_stale) c.z_dirty = true;
  c.stat_on_device = true;
  c.lml_dirty = true;
  c.grad_valid = false;
  return MOGP_OK;
}

int fetch_stat(mogp_engine* h, Cluster& c) {
  if (!c.stat_on_device) return MOGP_OK;
  order_after_side(h, c);
  if (c.z_dirty) {
    TRY(sync_theta(h, c));
    k_z<<<(unsigned)((c.m + 7) / 8), NTHREADS, 0, h->stream>>>(c.M, c.cap, c.x(), c.y(), (int)c.m, c.theta_dev(), c.z());
    LAUNCHED(h);
    c.z_dirty = false;
    c.lml_dirty = true;
  }
  if (c.lml_dirty) {
    TRY(launch_alpha_lml(h, c));
    c.lml_dirty = false;
  }
  TRY(ensure_pin(h, 256));
  CK(h, cudaMemcpyAsync(h->pin, c.stat(), sizeof(double) * 9, cudaMemcpyDeviceToHost, h->stream));
  CK(h, sync_stream(h));
  h->bytes_d2h += 72;
  const double* hs = reinterpret_cast<const double*>(h->pin);
  int hinfo;
  memcpy(&hinfo, hs + 8, sizeof(int));
  c.stat_on_device = false;
  if (hinfo != 0 || !isfinite(hs[0])) {
    c.valid = false;
    FAIL(h, MOGP_ENOTPD, "rank update lost positive definiteness (m=%lld)", (long long)c.m);
  }
  c.lml = hs[0];
  c.quad = hs[1];
  c.logdet = hs[2];
  return MOGP_OK;
}

// Append n <= NP points whose T = M Kx (n columns x mpad) and predictive means are already on the device.
int append_core(mogp_engine* h, Cluster& c, int n, const double* xB_dev, const double* yB_dev, const double* T_dev,
                const double* mu_dev, int64_t t_stride = 0) {
  const int64_t m = c.m, mpad_old = round_up(m, TB), mnew = m + n, mpad_new = round_up(mnew, TB);
  if (t_stride <= 0) t_stride = mpad_old;  // (columns kept current across changes keep the stride they were computed with)
  order_after_side(h, c);
  const bool reserved = mnew > c.cap;
  if (reserved) TRY(cluster_reserve(h, c, mnew, true));
  const int nb = (int)(mpad_old / TB);
  const int nblk = (int)((mpad_old + AP_GRAM_ROWS - 1) / AP_GRAM_ROWS);
  const bool ring = nblk <= AP_RING_NBLK;
  TRY(ensure(h, &h->d_apw, &h->apw_elems, NP * NP + NP + (int64_t)nblk * NG + APPEND_KS * NP * mpad_old));
  ApWork w;
  if (ring) {
    w.Li = h->d_apring + (int64_t)h->apring_next * AP_RING_SLOT;
    h->apring_next = (h->apring_next + 1) % AP_RING;
    w.znew = w.Li + NP * NP;
    w.gpart = w.znew + NP;
    w.part = h->d_apw;
  } else {
    w.Li = h->d_apw;
    w.znew = w.Li + NP * NP;
    w.gpart = w.znew + NP;
    w.part = w.gpart + (int64_t)nblk * NG;
  }
  Theta th = make_theta(c);
  int* info = reinterpret_cast<int*>(c.stat() + 8);
  // Gram of T, Ls and z_B depend only on what the scoring pass left behind, not on the factor: with a join-side score
  // correction waiting for them (look-ahead pipeline) they run on the quick stream, ahead of whatever is still queued
  // on the main stream, which picks up from there
  AlgHook* hk = (h->alg_hook && !h->alg_hook->fired && h->stream == h->main_stream) ? h->alg_hook : nullptr;
  const bool on_quick = hk && ring && !reserved;
  cudaStream_t s1 = on_quick ? h->quick : h->stream;
  if (hk) {
    hk->ntiles = (int)(mpad_old / TB);
    hk->on_quick = on_quick;
    if (hk->npos > 0) {
      TRY(ensure(h, &h->d_algw2, &h->algw2_elems, (int64_t)hk->npos * hk->ntiles * NP * NP));
      TRY(t_wait(h, s1, hk->bi, hk->slot));
      k_alg_join_part<<<dim3((unsigned)hk->ntiles, (unsigned)hk->npos), 256, 0, s1>>>(hk->jobs, h->d_algw2);
      LAUNCHED(h);
    }
  }
  k_append_gram<<<nblk, 256, 0, s1>>>(T_dev, (int)mpad_old, t_stride, n, w);
  LAUNCHED(h);
  k_append_small<<<1, 256, 0, s1>>>(mu_dev, nblk, xB_dev, yB_dev, n, th, w, info);
  LAUNCHED(h);
  if (on_quick) {
    CK(h, cudaEventRecord(h->ev_apq, h->quick));
    CK(h, cudaStreamWaitEvent(h->main_stream, h->ev_apq, 0));
  }
  if (hk) {
    hk->jobs.Li = w.Li;
    hk->jobs.znew = w.znew;
    if (hk->npos > 0) {
      k_alg_join_fin<<<hk->npos, 256, 0, s1>>>(hk->jobs, h->d_algw2, hk->ntiles);
      LAUNCHED(h);
      if (hk->jobs.maintain) TRY(t_mark(h, s1, hk->bi, hk->slot));
    }
    CK(h, cudaEventRecord(h->ev_alg, s1));
    hk->fired = true;
  }
  // alpha is updated in place below: behind the preambles of the passes in flight, which read it (their products read rows
  // of M this append does not touch, and mask the padding it writes into, score.cuh)
  if (h->chain.win.open)
    for (int f : h->chain.win.flight) CK(h, cudaStreamWaitEvent(h->stream, h->bb[f].res.pre_done, 0));
  k_append_rows<<<dim3(nb, APPEND_KS), NTHREADS, AP_ROWS_SMEM, h->stream>>>(c.M, c.cap, nb, (int)mpad_old, T_dev, t_stride, n, w);
  LAUNCHED(h);
  if (mpad_new > mpad_old) {
    k_pad_identity<<<64, 256, 0, h->stream>>>(c.M, c.cap, (int)mpad_old, (int)mpad_new);
    LAUNCHED(h);
  }
  k_append_finalize<<<(unsigned)((mpad_new + 255) / 256), 256, 0, h->stream>>>(c.M, c.cap, (int)m, n, (int)mpad_old,
                                                                               (int)mpad_new, w, xB_dev, yB_dev, c.x(), c.y(),
                                                                               c.z(), c.alpha());
  LAUNCHED(h);
  c.m = mnew;
  CK(h, cudaGetLastError());
  return mark_updated(h, c, false);
}

// mogp_cluster_append: host points, any n (chunks of NP).
int append_points(mogp_engine* h, Cluster& c, int64_t n, const double* x, const double* y) {
  int32_t slot = (int32_t)(&c - h->cl.data());
  for (int64_t q0 = 0; q0 < n; q0 += NP) {
    const int nn = (int)std::min<int64_t>(NP, n - q0);
    int64_t off[2] = {0, nn};
    TRY(stage_query(h, 1, off, x + q0, y + q0, nullptr));
    ScorePlan pl;
    TRY(score_device(h, 1, nn, h->d_qoff, h->d_q, h->d_q + nn, 1, &slot, -1, nullptr, nullptr, nullptr, nullptr, nullptr, true, &pl));
    Cluster& cc = h->cl[slot];
    TRY(append_core(h, cc, nn, h->d_q, h->d_q + nn, pl.T, pl.mu));
    cc.hx.insert(cc.hx.end(), x + q0, x + q0 + nn);
    cc.hy.insert(cc.hy.end(), y + q0, y + q0 + nn);
  }
  return fetch_stat(h, h->cl[slot]);
}

RmGeom make_geom(const Cluster& c, int64_t p, int n) {
  RmGeom g;
  g.m = (int)c.m;
  g.n = n;
  g.p = (int)p;
  g.mnew = (int)(c.m - n);
  g.nbn = (int)std::max<int64_t>(1, round_up(g.mnew, TB) / TB);
  g.I0 = std::min((int)(p / TB), g.nbn - 1);
  g.ld = c.cap;
  return g;
}

int rm_work(mogp_engine* h, const RmGeom& g, RmWork* w) {
  const int64_t nT = g.nbn - g.I0, rows = nT * TB;
  const int64_t ph = nT * g.nbn * (int64_t)(NP * TB);  // P and H tiles of the deletion
  // (the look-ahead streams compute own-cluster scores while the side stream may be deleting rows: own workspace)
  const bool la = h->stream == h->ahead || h->stream == h->ahead2;
  double** buf = la ? &h->d_upd2 : &h->d_upd;
  TRY(ensure(h, buf, la ? &h->upd2_elems : &h->upd_elems, rows * NP * 2 + rows + nT * NG + NG + 8 + NP + nT * TB * TB + 2 * ph));
  w->Bt = *buf;
  w->Gt = w->Bt + rows * NP;
  w->u = w->Gt + rows * NP;
  w->Pgram = w->u + rows;
  w->Gc = w->Pgram + nT * NG;
  w->own = w->Gc + NG;
  w->gsol = w->own + 8;
  w->S = w->gsol + NP;
  return MOGP_OK;
}

// Block Gram partials of the rows [p, p+n) of cluster c (input of both the own score and the deletion).
int rm_gram(mogp_engine* h, Cluster& c, int64_t p, int n) {
  RmGeom g = make_geom(c, p, n);
  RmWork w;
  TRY(rm_work(h, g, &w));
  k_rm_gram<<<g.nbn - g.I0, 256, 0, h->stream>>>(c.M, g, w);
  LAUNCHED(h);
  return MOGP_OK;
}

// Leave-block-out score into w.own[0..2] (device).
int own_score(mogp_engine* h, Cluster& c, int64_t p, int n, const double* yB_dev, int thr_index, double* out_dev = nullptr) {
  RmGeom g = make_geom(c, p, n);
  RmWork w;
  TRY(rm_work(h, g, &w));
  k_own_final<<<1, 256, 0, h->stream>>>(c.alpha(), yB_dev, g, w, g.nbn - g.I0, make_theta(c), thr_index,
                                        out_dev ? out_dev : w.own);
  LAUNCHED(h);
  return MOGP_OK;
}

// Delete rows/cols [p, p+n) (n <= NP) at fixed theta; rm_gram(c, p, n) must have been run on the current M.
int remove_core(mogp_engine* h, Cluster& c, int64_t p, int n) {
  RmGeom g = make_geom(c, p, n);
  RmWork w;
  TRY(rm_work(h, g, &w));
  const int nT = g.nbn - g.I0;
  k_rm_prep<<<nT, 256, RM_PREP_SMEM, h->stream>>>(g, w, c.alpha());
  LAUNCHED(h);
  double* nbuf = nullptr;
  TRY(pool_get(h, c.cap, &nbuf));
  double* nM = nbuf;
  double* nvec = nbuf + c.cap * c.cap;
  static const int legacy = env_int_once("MOGP_RM_LEGACY", 0);  // 1: the three-pass deletion (hpart / hscan / apply)
  if (h->leave_hook) TRY(fire_leave_hook(h, g, w, nT));
  if (!legacy) {
    // one pass over the factor: strip walk with the prefix carried along (k_rm_strip), alpha / x / y / stat in its epilogue
    static const int strip_w = env_int_once("MOGP_RM_SW", 16);
    if (strip_w == 64) k_rm_strip64<<<g.nbn, NTHREADS, RM_STRIP64_SMEM, h->stream>>>(c.M, nM, g, w, c.vec, nvec, c.cap);
    else k_rm_strip<<<g.nbn * (TB / RM_SW), NTHREADS, RM_STRIP_SMEM, h->stream>>>(c.M, nM, g, w, c.vec, nvec, c.cap);
    LAUNCHED(h);
    if (h->stream == h->side) h->pending_release.emplace_back(c.cap, c.M);  // reusable once the main stream has joined
    else pool_put(h, c.cap, c.M);
    c.M = nM;
    c.vec = nvec;
    c.m = g.mnew;
    c.hx.erase(c.hx.begin() + p, c.hx.begin() + p + n);
    c.hy.erase(c.hy.begin() + p, c.hy.begin() + p + n);
    CK(h, cudaGetLastError());
    return mark_updated(h, c, true);
  }
  double* P = w.S + (int64_t)nT * TB * TB;
  double* H = P + (int64_t)nT * g.nbn * (NP * TB);
  k_rm_hpart<<<dim3(g.nbn, nT), NTHREADS, 0, h->stream>>>(c.M, g, w, P);
  LAUNCHED(h);
  k_rm_hscan<<<g.nbn, NP * TB, 0, h->stream>>>(c.M, g, P, H);
  LAUNCHED(h);
  k_rm_apply<<<dim3(g.nbn, g.nbn), NTHREADS, RM_APPLY_SMEM, h->stream>>>(c.M, nM, g, w, H);
  LAUNCHED(h);
  k_rm_vec<<<(unsigned)((g.mnew + 255) / 256), 256, 0, h->stream>>>(c.x(), c.y(), nvec, nvec + c.cap, g);
  LAUNCHED(h);
  k_rm_alpha<<<(unsigned)((g.mnew + 255) / 256), 256, 0, h->stream>>>(c.alpha(), nvec + 2 * c.cap, H, P, g, w.gsol);
  LAUNCHED(h);
  // stat block travels with the buffer
  CK(h, cudaMemcpyAsync(nvec + 4 * c.cap, c.stat(), sizeof(double) * STAT, cudaMemcpyDeviceToDevice, h->stream));
  if (h->stream == h->side) h->pending_release.emplace_back(c.cap, c.M);  // reusable once the main stream has joined
  else pool_put(h, c.cap, c.M);
  c.M = nM;
  c.vec = nvec;
  c.m = g.mnew;
  c.hx.erase(c.hx.begin() + p, c.hx.begin() + p + n);
  c.hy.erase(c.hy.begin() + p, c.hy.begin() + p + n);
  CK(h, cudaGetLastError());
  return mark_updated(h, c, true);
}

// ---- the draw: np.exp(s - logsumexp(s)) then np.random.choice(active, p=p) (mogp_constrained.py:272-277) ----
// choice(): cdf = cumsum(p); cdf /= cdf[-1]; index = searchsorted(cdf, u, side='right').
int host_choice(const double* score, int n, double u, double* p_out, int* index_out, double* margin_out) {
  double amax = -INFINITY;
  for (int i = 0; i < n; ++i) {
    if (isnan(score[i])) return MOGP_EPROB;
    if (score[i] > amax) amax = score[i];
  }
  if (!isfinite(amax)) return MOGP_EPROB;
  double ssum = 0.0;
  for (int i = 0; i < n; ++i) ssum += exp(score[i] - amax);
  const double lse = log(ssum) + amax;
  std::vector<double> cdf(n);
  double run = 0.0;
  for (int i = 0; i < n; ++i) {
    double p = exp(score[i] - lse);
    if (p_out) p_out[i] = p;
    run += p;
    cdf[i] = run;
  }
  if (!(fabs(run - 1.0) <= 1.4901161193847656e-08)) return MOGP_EPROB;  // choice(): probabilities do not sum to 1
  int idx = n;
  double margin = INFINITY;
  for (int i = 0; i < n; ++i) {
    cdf[i] /= cdf[n - 1];
    if (idx == n && cdf[i] > u) idx = i;
    margin = std::min(margin, fabs(cdf[i] - u));
  }
  if (idx >= n) idx = n - 1;
  if (index_out) *index_out = idx;
  if (margin_out) *margin_out = margin;
  return MOGP_OK;
}

// self.p[n] (mogp_constrained.py:276) of one step, kept for mogp_chain_get_probs
void record_step_probs(mogp_engine* h, const std::vector<int32_t>& ids, const double* p) {
  if (!h->record_probs) return;
  if (h->rec_off.empty()) h->rec_off.push_back(0);
  h->rec_ids.insert(h->rec_ids.end(), ids.begin(), ids.end());
  h->rec_p.insert(h->rec_p.end(), p, p + ids.size());
  h->rec_off.push_back((int64_t)h->rec_ids.size());
}

Theta newcomer_theta(const mogp_chain_config& cfg) {
  Theta th;
  th.A = 0.0;
  th.k0 = cfg.signal_variance;
  th.k1 = cfg.spec.kernel == MOGP_KERNEL_RBF ? cfg.lengthscale : 1.0;  // Bias(1) in the linear model (obsmodel.py:37)
  th.noise = cfg.noise_variance;
  th.kernel = cfg.spec.kernel;
  th.mean_func = cfg.spec.mean_func;
  return th;
}

int chain_check(mogp_engine* h, int64_t n) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!h->chain.active) FAIL(h, MOGP_ESTATE, "no chain attached");
  if (n < 0 || n >= h->chain.N) FAIL(h, MOGP_EINVAL, "patient %lld out of range", (long long)n);
  return MOGP_OK;
}

// position (row offset) of patient n inside cluster c; -1 if absent
int64_t member_offset(mogp_engine* h, const Cluster& c, int64_t n, size_t* index) {
  int64_t off = 0;
  for (size_t i = 0; i < c.members.size(); ++i) {
    if (c.members[i] == n) {
      if (index) *index = i;
      return off;
    }
    off += h->h_off[c.members[i] + 1] - h->h_off[c.members[i]];
  }
  return -1;
}

int rebuild_from_members(mogp_engine* h, Cluster& c) {
  std::vector<double> xs, ys;
  for (int64_t id : c.members) {
    xs.insert(xs.end(), h->h_x.begin() + h->h_off[id], h->h_x.begin() + h->h_off[id + 1]);
    ys.insert(ys.end(), h->h_y.begin() + h->h_off[id], h->h_y.begin() + h->h_off[id + 1]);
  }
  if (xs.empty()) FAIL(h, MOGP_ESTATE, "cluster members have no visits");
  return upload_data(h, c, (int64_t)xs.size(), xs.data(), ys.data());
}

// Step phase 1: take patient n out (mogp_constrained.py:232-244).
int chain_begin(mogp_engine* h, int64_t n) {
  Chain& ch = h->chain;
  if (ch.pending >= 0) FAIL(h, MOGP_ESTATE, "previous step not committed");
  const int32_t prev = ch.z[n];
  if (prev < 0 || prev >= (int32_t)ch.Nk.size() || ch.slot_of_id[prev] < 0) FAIL(h, MOGP_ESTATE, "patient %lld has no cluster", (long long)n);
  ch.z[n] = -1;
  ch.Nk[prev] -= 1;  // allocmodel.calc_score, allocmodel.py:24
  ch.pending = n;
  ch.pending_prev = prev;
  ch.pending_lazy = 0;
  ch.pending_p = -1;
  ch.gram_ready = false;
  Cluster& c = h->cl[ch.slot_of_id[prev]];
  const int64_t ni = h->h_off[n + 1] - h->h_off[n];
  if (ch.Nk[prev] == 0) {  // obsmodel[curr_k] = None (:235-237)
    TRY(cluster_free_i(h, ch.slot_of_id[prev]));
    ch.slot_of_id[prev] = -1;
    return MOGP_OK;
  }
  if (ch.cfg.fast_updates && ni > 0 && ni <= NP && c.valid && member_offset(h, c, n, nullptr) >= 0) {
    ch.pending_lazy = 1;  // rows stay in place; the own score is the leave-block-out form
    return MOGP_OK;
  }
  // reference-exact: re-extract the members in patient-index order and refactorise (:239-242).  The member list
  // is re-derived from z (X[z == curr_k] in the reference), which also covers a cluster whose GP was replaced by a
  // newcomer's model (alpha = 0, see chain_score).
  c.members.clear();
  for (int64_t i = 0; i < ch.N; ++i)
    if (ch.z[i] == prev) c.members.push_back(i);
  return rebuild_from_members(h, c);
}

// Step phase 2: score[active] = log prior weight + log likelihood (:244-271).  Needs a_draw for the newcomer.
int chain_score(mogp_engine* h, double a_draw) {
  Chain& ch = h->chain;
  const int64_t n = ch.pending;
  if (n < 0) FAIL(h, MOGP_ESTATE, "chain_score without chain_begin");
  const int64_t c0 = h->h_off[n], ni = h->h_off[n + 1] - c0;
  ch.act_ids.clear();
  ch.act_score.clear();
  ch.score_slots.clear();
  const int32_t n_ids = (int32_t)ch.Nk.size();
  for (int32_t k = 0; k < n_ids; ++k)
    if (ch.Nk[k] > 0) ch.act_ids.push_back(k);
  // active = np.where(hstack([Nk, alpha]) > 0) (allocmodel.py:25-26): the new-cluster slot is active only when
  // alpha > 0.  The reference then treats the LAST active id as "the new component" whatever it is
  // (mogp_constrained.py:252,270): with alpha = 0 that is an existing cluster, which is scored with the newcomer's
  // marginal and, if drawn, has its GP REPLACED by the newcomer's model (:279-281).  Reproduced as is.
  if (ch.cfg.alpha > 0.0) ch.act_ids.push_back(n_ids);
  const int nact = (int)ch.act_ids.size();
  if (nact == 0) FAIL(h, MOGP_ESTATE, "no active cluster and alpha = 0");
  ch.act_score.assign(nact, 0.0);
  int32_t own = ch.pending_lazy ? ch.pending_prev : -1;
  if (own >= 0 && own == ch.act_ids[nact - 1]) own = -1;  // scored as "new": no own-cluster score needed
  for (int i = 0; i + 1 < nact; ++i)
    if (ch.act_ids[i] != own) ch.score_slots.push_back(ch.slot_of_id[ch.act_ids[i]]);
  const int ns = (int)ch.score_slots.size();
  const int thr_index = ch.cfg.use_threshold ? ch.cfg.thr_index : -1;
  // outputs on the device: score[ns] | mu_thr[ns] | new_ll[1] | own[3]
  TRY(ensure(h, &h->d_out, &h->out_elems, 2 * (int64_t)ns + 8));
  double* d_score = h->d_out;
  double* d_muthr = h->d_out + ns;
  double* d_newll = h->d_out + 2 * ns;
  double* d_own = d_newll + 1;
  ch.plan = ScorePlan();
  if (ni > 0) {
    {
      // side stream: the own-cluster (leave-block-out) score and the newcomer's marginal are independent of the
      // batched scoring pass below and hide under it
      SideScope side(h);
      if (own >= 0) {
        Cluster& c = h->cl[ch.slot_of_id[own]];
        ch.pending_p = member_offset(h, c, n, nullptr);
        if (ch.pending_p < 0) FAIL(h, MOGP_ESTATE, "patient %lld not in its cluster's member list", (long long)n);
        TRY(rm_gram(h, c, ch.pending_p, (int)ni));
        ch.gram_ready = true;
        TRY(own_score(h, c, ch.pending_p, (int)ni, h->d_py + c0, thr_index));
        RmGeom g = make_geom(c, ch.pending_p, (int)ni);
        RmWork w;
        TRY(rm_work(h, g, &w));
        CK(h, cudaMemcpyAsync(d_own, w.own, sizeof(double) * 3, cudaMemcpyDeviceToDevice, h->stream));
      }
      // the newcomer: GPobs(x_n, y_n) at A = |a_draw|, l = 4, noise_0 (:246-250)
      ch.a_dev_val = a_draw;
      CK(h, cudaMemcpyAsync(h->d_adraw, &ch.a_dev_val, sizeof(double), cudaMemcpyHostToDevice, h->stream));
      h->bytes_h2d += 8;
      k_new_ll<<<1, 64, 0, h->stream>>>(h->d_off + n, h->d_px + c0, h->d_py + c0, h->d_adraw, newcomer_theta(ch.cfg), d_newll);
      LAUNCHED(h);
    }
    if (ns > 0)
      TRY(score_device(h, 1, ni, h->d_off + n, h->d_px + c0, h->d_py + c0, ns, ch.score_slots.data(), thr_index, d_score,
                       d_muthr, nullptr, nullptr, nullptr, ch.cfg.fast_updates != 0, &ch.plan));
    join_side(h);
    CK(h, cudaGetLastError());
  }
  std::vector<double> host(2 * ns + 8, 0.0);
  if (ni > 0) TRY(copy_back(h, host.data(), h->d_out, 2 * ns + 4));
  if (ni > 64) {  // a row longer than k_new_ll's shared-memory tile: the general cluster path (any length)
    const double th4[4] = {fabs(a_draw), ch.cfg.signal_variance,
                           ch.cfg.spec.kernel == MOGP_KERNEL_RBF ? ch.cfg.lengthscale : 1.0, ch.cfg.noise_variance};
    double ll_long = NAN;
    int rc = newcomer_ll_general(h, ch.cfg.spec, th4, ni, h->h_x.data() + c0, h->h_y.data() + c0, &ll_long);
    if (rc != MOGP_OK && rc != MOGP_ENOTPD) return rc;
    host[2 * ns] = rc == MOGP_OK ? ll_long : NAN;
  }
  const double ythr = (ch.cfg.use_threshold && !h->h_ythr.empty()) ? h->h_ythr[n] : NAN;
  int si = 0;
  for (int i = 0; i < nact; ++i) {
    const int32_t id = ch.act_ids[i];
    double prior, ll, muthr = NAN;
    if (i == nact - 1) {
      prior = log((id == n_ids ? ch.cfg.alpha : (double)ch.Nk[id]) + 1e-16);  // allocmodel.py:25-26
      ll = host[2 * ns];
      if (ni > 0 && !isfinite(ll)) FAIL(h, MOGP_ENOTPD, "new-cluster covariance not positive definite");
    } else {
      prior = log((double)ch.Nk[id] + 1e-16);
      if (id == own) {
        if (ni > 0 && host[2 * ns + 3] != 1.0) FAIL(h, MOGP_ENOTPD, "leave-out block not positive definite");
        ll = host[2 * ns + 1];
        muthr = host[2 * ns + 2];
      } else {
        ll = host[si];
        muthr = host[ns + si];
        ++si;
      }
      // monotonicity rule (:254-263): Y[n, idx] - y* > threshold  ->  -inf
      if (ch.cfg.use_threshold && ni > 0 && (ythr - muthr > ch.cfg.threshold)) ll = -INFINITY;
    }
    ch.act_score[i] = prior + ll;
  }
  return MOGP_OK;
}

// Where the scoring pass left T = M Kx (and the predictive means) of the pending patient under `slot`.
void locate_T(mogp_engine* h, int32_t slot, const double** T, const double** mu, int64_t* stride) {
  Chain& ch = h->chain;
  *T = *mu = nullptr;
  *stride = 0;  // (0: the slot's padded size, append_core)
  if (ch.win.in_step) {
    const Chain::Window& w = ch.win;
    if (slot >= 0 && slot < (int32_t)w.slot_upto.size() && w.slot_upto[slot] > w.cur && w.slot_T_ok[slot] &&
        slot < (int32_t)w.e[w.cur].size() && !w.e[w.cur][slot].pruned) {
      *T = w.e[w.cur][slot].T;
      *mu = w.e[w.cur][slot].mu;
      *stride = w.e[w.cur][slot].stride;
    }
    return;
  }
  if (!ch.plan.T) return;
  int64_t kx_off = 0;
  for (size_t q = 0; q < ch.score_slots.size(); ++q) {
    if (ch.score_slots[q] == slot) {
      *T = ch.plan.T + kx_off;
      *mu = ch.plan.mu + (int64_t)q * ch.plan.Rpad;
      return;
    }
    kx_off += ch.plan.Rpad * round_up(h->cl[ch.score_slots[q]].m, TB);
  }
}

// Step phase 3: assign (mogp_constrained.py:277-284 minus the refit).
int chain_commit(mogp_engine* h, int32_t chosen_id, double a_draw, int32_t* is_new_out) {
  Chain& ch = h->chain;
  const int64_t n = ch.pending;
  if (n < 0) FAIL(h, MOGP_ESTATE, "chain_commit without chain_begin");
  const int32_t n_ids = (int32_t)ch.Nk.size();
  // "new component" = the last active id of the scoring phase (an existing cluster when alpha = 0)
  const int32_t new_comp = ch.act_ids.empty() ? n_ids : ch.act_ids.back();
  const bool is_new = chosen_id == new_comp;
  const bool replaces_existing = is_new && chosen_id < n_ids;
  if (chosen_id < 0 || chosen_id > n_ids || (chosen_id == n_ids && ch.cfg.alpha <= 0.0) ||
      (chosen_id < n_ids && ch.Nk[chosen_id] <= 0))
    FAIL(h, MOGP_EINVAL, "cluster id %d is not active", chosen_id);
  const int64_t c0 = h->h_off[n], ni = h->h_off[n + 1] - c0;
  const int32_t prev = ch.pending_prev;
  if (ch.pending_lazy && chosen_id == prev && !replaces_existing) {
    // stayed: the factor already contains the patient (the reference re-appends the same rows)
  } else {
    bool side_busy = false;
    if (ch.pending_lazy) {  // really leaves: delete its rows from the previous cluster - on the side stream, so that
      Cluster& c = h->cl[ch.slot_of_id[prev]];  // it overlaps the append to the target cluster below
      size_t idx = 0;
      const int64_t p = member_offset(h, c, n, &idx);
      {
        SideScope side(h);
        if (!ch.gram_ready) TRY(rm_gram(h, c, p, (int)ni));
        TRY(remove_core(h, c, p, (int)ni));
      }
      side_busy = true;
      c.members.erase(c.members.begin() + idx);
    }
    // whatever path the assignment takes, the main stream ends up ordered after the deletion - at once, or (look-ahead
    // pipeline) when main-stream work next touches that cluster (order_after_side), opens a batch or re-scores
    struct Joiner {
      mogp_engine* h;
      bool on;
      int32_t slot;
      ~Joiner() {
        if (!on) return;
        if (h->chain.win.in_step && slot >= 0 && h->cl[slot].used) {
          h->cl[slot].side_busy = true;
          h->side_pending = true;
        } else {
          join_side(h);
        }
      }
    } joiner{h, side_busy, ch.pending_lazy ? ch.slot_of_id[prev] : -1};
    if (is_new) {
      int32_t slot;
      double th[4] = {fabs(a_draw), ch.cfg.signal_variance,
                      ch.cfg.spec.kernel == MOGP_KERNEL_RBF ? ch.cfg.lengthscale : 1.0, ch.cfg.noise_variance};
      if (replaces_existing) {  // obsmodel[z[n]] = new_comp_model over an existing cluster (alpha = 0)
        join_side(h);
        joiner.on = false;
        TRY(cluster_free_i(h, ch.slot_of_id[chosen_id]));
        ch.slot_of_id[chosen_id] = -1;
      }
      TRY(cluster_create_i(h, &ch.cfg.spec, th, &slot));
      Cluster& c = h->cl[slot];
      c.members.assign(1, n);
      TRY(rebuild_from_members(h, c));
      if (ch.cfg.fast_updates) TRY(infer(h, c));
      if (replaces_existing) {
        ch.slot_of_id[chosen_id] = slot;
      } else {
        ch.Nk.push_back(0);
        ch.slot_of_id.push_back(slot);
      }
    } else {
      Cluster& c = h->cl[ch.slot_of_id[chosen_id]];
      bool done = false;
      if (ch.cfg.fast_updates && c.valid && ni > 0 && ni <= NP) {
        // T = M Kx and the predictive means of this cluster are still in the scoring scratch
        const double *Tc = nullptr, *muc = nullptr;
        int64_t t_stride = 0;
        locate_T(h, ch.slot_of_id[chosen_id], &Tc, &muc, &t_stride);
        if (Tc && ch.win.in_step) {  // columns kept current by kernels on other streams: behind the last of them
          TRY(t_wait(h, h->main_stream, ch.win.cur_bi, ch.slot_of_id[chosen_id]));
          TRY(t_wait(h, h->quick, ch.win.cur_bi, ch.slot_of_id[chosen_id]));
        }
        if (!Tc && ch.win.in_step) {
          // the cached T under this cluster is stale (its scores were corrected by a rank-n formula): one product
          // against this cluster alone, queued behind the update that made it stale
          int32_t sl = ch.slot_of_id[chosen_id];
          ScorePlan pl1;
          TRY(score_device(h, 1, ni, h->d_off + n, h->d_px + c0, h->d_py + c0, 1, &sl, -1, nullptr, nullptr, nullptr, nullptr,
                           nullptr, true, &pl1));
          Tc = pl1.T;
          muc = pl1.mu;
          t_stride = 0;
        }
        if (Tc) {
          TRY(append_core(h, c, (int)ni, h->d_px + c0, h->d_py + c0, Tc, muc, t_stride));
          c.hx.insert(c.hx.end(), h->h_x.begin() + c0, h->h_x.begin() + c0 + ni);
          c.hy.insert(c.hy.end(), h->h_y.begin() + c0, h->h_y.begin() + c0 + ni);
          c.members.push_back(n);
          done = true;
        }
      }
      if (!done) {  // reference order: existing rows, then the patient's (obsmodel.py:77-78, 89-92)
        c.members.push_back(n);
        TRY(rebuild_from_members(h, c));
      }
    }
  }
  ch.z[n] = chosen_id;
  ch.Nk[chosen_id] += 1;  // allocmodel.update_suffstats, allocmodel.py:29-34
  ch.pending = -1;
  ch.plan = ScorePlan();
  if (is_new_out) *is_new_out = is_new ? 1 : 0;
  return MOGP_OK;
}


// =================================================================================================
// Look-ahead pipeline (fast mode, mogp_chain_sweep).  One patient's columns use the factors at 2 n flops per 8 bytes:
// patient by patient the scoring pass is bound by HBM and every step pays a device round trip.  The sweep order and
// the draws are known in advance, so
//   * consecutive patients totalling <= `lookahead` query columns form a BATCH that is scored against EVERY active
//     cluster in one pass over the factors (FP64-tensor bound instead of HBM bound);
//   * while the host draws and moves the patients of batch i, the pass of batch i+1 already runs on a low-priority
//     stream and fills the SMs the short dependent kernels of the moves leave idle.
// The chain stays strictly sequential: a cached score is used only if its cluster has not changed since it was
// computed (per-slot change counters); when a patient moves a -> b (or opens / empties a cluster) every launched
// position still ahead is re-scored against exactly the clusters that changed, before the next draw.  Same
// arithmetic per (patient, cluster) as the patient-by-patient path.
struct PhaseTimer {  // host wall time of a phase, accumulated into the chain's counters
  double* acc;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(double* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~PhaseTimer() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int window_cols(const Chain& ch) {
  const int env = env_int("MOGP_LOOKAHEAD", -1);  // (read per sweep: tests and tuning runs toggle it)
  int v = env >= 0 ? env : (ch.cfg.lookahead == 0 ? 64 : ch.cfg.lookahead);
  return v <= 0 ? 0 : std::min(v, 64);
}
int pipeline_depth(const mogp_engine* h) {  // 1: one batch at a time; 2: the next batch's pass overlaps the current batch's moves; 3: the next two
  // (a pass launched two batches ahead is stale under every cluster that changes meanwhile: worth it only when those are
  // brought up to date by the corrections chained behind it, MOGP_TMAINT)
  const int env = env_int("MOGP_PIPELINE", tmaint_on(h) ? 3 : 2);
  return std::max(1, std::min(env, NTGT));
}

void win_size_slots(Chain::Window& w, size_t n) {
  if (w.slot_upto.size() >= n) return;
  w.slot_upto.resize(n, 0);
  w.slot_T_ok.resize(n, 0);
  w.slot_alg.resize(n, 0);
  w.slot_epoch.resize(n, 0);
  for (int i = 0; i < NBUF; ++i) {
    w.slot_tev[i].resize(n, -1);
    w.rp_epoch[i].resize(n, -1);
    w.rp_last[i].resize(n, -1);
  }
}

// The cluster in `slot` changed (rows appended / deleted, freed, created): cached scores under it are stale.
void win_invalidate(Chain& ch, int32_t slot) {
  Chain::Window& w = ch.win;
  if (!w.open || slot < 0) return;
  win_size_slots(w, (size_t)slot + 1);
  w.slot_upto[slot] = 0;
  w.slot_T_ok[slot] = 0;
  w.slot_epoch[slot] += 1;
  for (int64_t v = w.cur; v < w.end_launched; ++v) {
    const int32_t id = (w.pat[v] == ch.pending) ? ch.pending_prev : ch.z[w.pat[v]];
    if (id >= 0 && id < (int32_t)ch.slot_of_id.size() && ch.slot_of_id[id] == slot) w.own_ok[v] = 0;
  }
}
void win_invalidate_all(Chain& ch) {
  Chain::Window& w = ch.win;
  for (size_t s = 0; s < w.slot_upto.size(); ++s) {
    w.slot_upto[s] = 0;
    w.slot_T_ok[s] = 0;
    w.slot_epoch[s] += 1;
  }
  for (int64_t v = w.cur; v < w.end_launched; ++v) w.own_ok[v] = 0;
}

// Leave-block-out scores of positions [v0, v1) that need one (own_ok clear when `only_stale`), in ONE launch pair on the
// current stream; out3 + 3 (v - v0) receives (score, mean at the threshold visit, ok flag); slot_out[v - v0] = the
// cluster slot scored or -1.
int win_launch_own_batch(mogp_engine* h, int64_t v0, int64_t v1, bool only_stale, int thr_index, double* out3, int32_t* slot_out) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  const bool la = h->stream == h->ahead || h->stream == h->ahead2;
  double** buf = la ? &h->d_own2 : &h->d_own1;
  int64_t* have = la ? &h->own2_elems : &h->own1_elems;
  for (int64_t base = v0; base < v1; base += OWN_BATCH) {
    OwnJobs jobs;
    int nj = 0, max_nT = 0;
    int64_t need = 0;
    const int64_t top = std::min<int64_t>(v1, base + OWN_BATCH);
    for (int64_t v = base; v < top; ++v) {
      slot_out[v - v0] = -1;
      if (only_stale && w.own_ok[v]) continue;
      const int64_t n = w.pat[v];
      const bool is_pending = n == ch.pending;
      const int32_t id = is_pending ? ch.pending_prev : ch.z[n];
      if (id < 0 || ch.slot_of_id[id] < 0) continue;
      if (ch.Nk[id] < (is_pending ? 1 : 2)) continue;  // alone in its cluster: the cluster goes when it is taken out
      if (is_pending && !ch.pending_lazy) continue;
      Cluster& c = h->cl[ch.slot_of_id[id]];
      if (!c.valid) continue;
      const int64_t c0 = h->h_off[n], ni = h->h_off[n + 1] - c0;
      const int64_t p = member_offset(h, c, n, nullptr);
      if (p < 0) continue;
      OwnJob& jb = jobs.j[nj++];
      jb.M = c.M;
      jb.alpha = c.alpha();
      jb.yB = h->d_py + c0;
      jb.out = out3 + 3 * (v - v0);
      jb.g = make_geom(c, p, (int)ni);
      jb.th = make_theta(c);
      jb.nT = jb.g.nbn - jb.g.I0;
      jb.w = RmWork{nullptr, nullptr, nullptr, reinterpret_cast<double*>((intptr_t)need), nullptr, nullptr};  // offsets first
      need += (int64_t)jb.nT * NG + NG;
      max_nT = std::max(max_nT, jb.nT);
      slot_out[v - v0] = ch.slot_of_id[id];
    }
    if (nj == 0) continue;
    TRY(ensure(h, buf, have, need));
    for (int q = 0; q < nj; ++q) {
      const int64_t off = (int64_t)(intptr_t)jobs.j[q].w.Pgram;
      jobs.j[q].w.Pgram = *buf + off;
      jobs.j[q].w.Gc = jobs.j[q].w.Pgram + (int64_t)jobs.j[q].nT * NG;
    }
    k_own_gram_batch<<<dim3((unsigned)max_nT, (unsigned)nj), 256, 0, h->stream>>>(jobs);
    LAUNCHED(h);
    k_own_final_batch<<<(unsigned)nj, 256, 0, h->stream>>>(jobs, thr_index);
    LAUNCHED(h);
  }
  return MOGP_OK;
}

// Pack the whole sweep's query (patients in sweep order) on the device: x | y | a_draw, offsets per position.
int win_prepare(mogp_engine* h, int64_t n_steps, const int64_t* order, const double* a_draw) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  w.open = false;
  w.in_step = false;
  w.n = n_steps;
  w.pat.assign(order, order + n_steps);
  w.col0.assign(n_steps + 1, 0);
  for (int64_t v = 0; v < n_steps; ++v) {
    const int64_t n = order[v];
    const int64_t ni = (n >= 0 && n < ch.N) ? h->h_off[n + 1] - h->h_off[n] : 0;
    w.col0[v + 1] = w.col0[v] + ni;
  }
  const int64_t R = w.col0[n_steps];
  w.cur = w.end_launched = 0;
  w.e.assign(n_steps, std::vector<Chain::WinEntry>());
  w.slot_upto.assign(h->cl.size(), 0);
  w.slot_T_ok.assign(h->cl.size(), 0);
  w.slot_alg.assign(h->cl.size(), 0);
  w.slot_epoch.assign(h->cl.size(), 0);
  for (int i = 0; i < NBUF; ++i) {
    w.slot_tev[i].assign(h->cl.size(), -1);
    w.rp_epoch[i].assign(h->cl.size(), -1);
    w.rp_last[i].assign(h->cl.size(), -1);
    w.async_items[i].clear();
    w.rpw_used[i] = 0;
  }
  w.flight.clear();
  w.batch_end = 0;
  w.cur_bi = 0;
  {  // partial sums of the corrections queued ahead: room for REPLAY_MAX of them against the largest cluster
    int64_t maxcap = TB;
    for (const Cluster& c : h->cl)
      if (c.used) maxcap = std::max(maxcap, c.cap);
    {  // (a share per batch buffer; a correction that does not fit breaks its chain and the cluster is re-scored)
      const int64_t share = (int64_t)12 * OWN_BATCH * (maxcap / TB + 8) * ALG_ENTRIES;
      if (share > h->rpw_share) {
        TRY(ensure(h, &h->d_rpw, &h->rpw_elems, share * NBUF));
        h->rpw_share = h->rpw_elems / NBUF;
      }
    }
    if (tmaint_on(h)) {  // private copies of the deletions' operands for k_T_leave (fire_leave_hook)
      const int64_t nT = maxcap / TB + 2, rows = nT * TB;
      const int64_t per = rows * NP * 2 + rows + nT * NG + NG + 8 + NP + nT * TB * TB;
      if (per > h->tsw_slot) {
        int64_t have = h->tsw_slot * TS_RING;
        TRY(ensure(h, &h->d_tsw, &have, per * TS_RING));
        h->tsw_slot = have / TS_RING;
      }
    }
  }
  w.new_ll.assign(n_steps, 0.0);
  w.own_ll.assign(n_steps, 0.0);
  w.own_mu.assign(n_steps, 0.0);
  w.own_flag.assign(n_steps, 0.0);
  w.own_ok.assign(n_steps, 0);
  TRY(ensure(h, &h->d_woff, &h->woff_elems, n_steps + 1));
  TRY(ensure(h, &h->d_wq, &h->wq_elems, 2 * R + 2 * n_steps + 1));
  char* up = nullptr;
  const size_t bytes = sizeof(int64_t) * (n_steps + 1) + sizeof(double) * (2 * R + 2 * n_steps);
  TRY(stage_alloc(h, bytes, &up));
  int64_t* so = reinterpret_cast<int64_t*>(up);
  double* sd = reinterpret_cast<double*>(so + n_steps + 1);
  memcpy(so, w.col0.data(), sizeof(int64_t) * (n_steps + 1));
  for (int64_t v = 0; v < n_steps; ++v) {
    const int64_t ni = w.col0[v + 1] - w.col0[v];
    if (ni > 0) {
      const int64_t c0 = h->h_off[order[v]];
      memcpy(sd + w.col0[v], h->h_x.data() + c0, sizeof(double) * ni);
      memcpy(sd + R + w.col0[v], h->h_y.data() + c0, sizeof(double) * ni);
    }
    sd[2 * R + v] = a_draw ? a_draw[v] : 0.0;
    // the raw cell Y[n, index_to_eval] of the monotonicity rule (NaN: the rule never fires), for the device-side pruning
    const int64_t pn = order[v];
    sd[2 * R + n_steps + v] = (!h->h_ythr.empty() && pn >= 0 && pn < ch.N) ? h->h_ythr[pn] : NAN;
  }
  CK(h, cudaMemcpyAsync(h->d_woff, so, sizeof(int64_t) * (n_steps + 1), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_wq, sd, sizeof(double) * (2 * R + 2 * n_steps), cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)bytes;
  w.open = true;
  return MOGP_OK;
}

// How many consecutive positions from `pos` form a batch (0: the patient there cannot be served from a batch).
// The pass computes columns in groups of 8 (the DMMA n-dimension): among the cuts between 70 % of the budget and the
// budget + 8 columns the one that wastes the smallest share of its last group wins (ties: the longer batch).  On
// ALSFRS-like rows (4..11 visits) that fills 97 % of the computed columns instead of 94 % and needs 8 % fewer batches.
int64_t batch_form(mogp_engine* h, int64_t pos, int max_cols) {
  const Chain::Window& w = h->chain.win;
  const int64_t lo = (int64_t)max_cols * 7 / 10, hi = std::min<int64_t>(64, (int64_t)max_cols + 8);
  int64_t cnt = 0, cols = 0, best_cnt = 0, best_cols = 0;
  double best_waste = 2.0;
  while (pos + cnt < w.n && cnt < 16) {
    const int64_t n = w.pat[pos + cnt];
    if (n < 0 || n >= h->chain.N) break;
    const int64_t ni = w.col0[pos + cnt + 1] - w.col0[pos + cnt];
    if (ni <= 0 || ni > NP || (cnt > 0 && cols + ni > hi)) break;
    cols += ni;
    ++cnt;
    if (cols >= lo) {
      const int64_t padded = round_up(cols, 8);
      const double waste = (double)(padded - cols) / (double)padded;
      if (waste < best_waste || (waste == best_waste && cols > best_cols)) {
        best_waste = waste;
        best_cnt = cnt;
        best_cols = cols;
      }
    }
  }
  return best_cnt > 0 ? best_cnt : cnt;  // (fewer than `lo` columns left, or a single long row: take what there is)
}

template <typename T>
int ensure_pinned(mogp_engine* h, T** p, int64_t* have, int64_t need) {
  if (need <= *have) return MOGP_OK;
  if (*p) CK(h, cudaFreeHost(*p));
  *p = nullptr;
  *have = 0;
  const int64_t want = need + need / 2 + 64;
  CK(h, cudaMallocHost((void**)p, sizeof(T) * (size_t)want));
  *have = want;
  return MOGP_OK;
}

// Launch the pass of batch `bi` = positions [s0, s0 + cnt) on the look-ahead streams: every active cluster x the
// batch's columns, the patients' own-cluster scores and new-cluster marginals; results travel to pinned memory.
// Ordered after everything enqueued so far on the main stream; no host wait.
int batch_launch(mogp_engine* h, int bi, int64_t s0, int64_t cnt) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  Chain::Batch& b = w.b[bi];
  BatchBuf& bf = h->bb[bi];
  b.s0 = s0;
  b.cnt = cnt;
  b.slots.clear();
  b.epoch.clear();
  b.mpad.clear();
  b.msz.clear();
  b.prof = h->prof_on;
  win_size_slots(w, h->cl.size());
  std::fill(w.rp_epoch[bi].begin(), w.rp_epoch[bi].end(), -1);  // (what the buffer's previous pass left)
  std::fill(w.rp_last[bi].begin(), w.rp_last[bi].end(), -1);
  w.async_items[bi].clear();
  w.rpw_used[bi] = 0;
  for (size_t k = 0; k < ch.Nk.size(); ++k)
    if (ch.Nk[k] > 0 && ch.slot_of_id[k] >= 0) {
      const int32_t slot = ch.slot_of_id[k];
      if (!h->cl[slot].valid) {  // (synchronises: before anything is queued on the look-ahead streams)
        TRY(infer(h, h->cl[slot]));
        w.slot_epoch[slot] += 1;
      }
      b.slots.push_back(slot);
      b.epoch.push_back(w.slot_epoch[slot]);
      w.rp_epoch[bi][slot] = w.slot_epoch[slot];  // the pass sees the cluster as it is now; corrections may be chained behind it
      w.rp_last[bi][slot] = -1;
      w.slot_tev[bi][slot] = -1;
      b.mpad.push_back(round_up(h->cl[slot].m, TB));
      b.msz.push_back(h->cl[slot].m);
    }
  const int ns = (int)b.slots.size();
  const int thr_index = ch.cfg.use_threshold ? ch.cfg.thr_index : -1;
  const int64_t total = 2 * cnt * (int64_t)ns + 4 * cnt;
  TRY(ensure(h, &bf.d_out, &bf.out_elems, total));
  TRY(ensure_pinned(h, &bf.h_out, &bf.hout_elems, total));
  double* d_score = bf.d_out;
  double* d_mu = d_score + cnt * ns;
  double* d_new = d_mu + cnt * ns;
  double* d_own = d_new + cnt;
  for (int64_t v = s0; v < s0 + cnt; ++v) w.e[v].assign(h->cl.size(), Chain::WinEntry());
  const int64_t Rtot = w.col0[w.n], cb = w.col0[s0], R = w.col0[s0 + cnt] - cb;
  if (h->side_pending) join_side(h);
  h->ahead = h->ahead_ring[bi];
  // Passes in flight are independent streams: the later pass's CTAs fill the SMs the earlier one's tail leaves idle (same
  // priority: the block scheduler serves the older grid first).  MOGP_WIDE_CHAIN=1: the product of this pass waits for that of
  // the pass launched before it instead (measured 2 % slower: 1.031 vs 1.050 sweeps/s).
  bf.res.wide_after = env_int("MOGP_WIDE_CHAIN", 0) != 0 ? h->last_wide : nullptr;
  h->last_wide = bf.res.wide_done;
  CK(h, cudaEventRecord(h->ev_fork, h->main_stream));
  CK(h, cudaStreamWaitEvent(h->ahead, h->ev_fork, 0));
  CK(h, cudaStreamWaitEvent(h->ahead2, h->ev_fork, 0));
  if (tmaint_on(h)) {
    // This pass overwrites columns that maintenance kernels worked on when the buffer last held a batch, and that the
    // corrections chained behind later passes read (the joiner's T): behind the correction streams of every buffer whose
    // pass is not in flight (those of the passes in flight wait for passes this one is queued behind anyway) ...
    for (int q = 0; q < NBUF; ++q) {
      if (std::find(w.flight.begin(), w.flight.end(), q) != w.flight.end()) continue;
      cudaStream_t ms[2] = {h->ts1[q], h->tq[q]};
      for (int e = 0; e < 2; ++e) {
        cudaEvent_t& ev = h->ev_ts[2 * q + e];
        if (!ev) CK(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(h, cudaEventRecord(ev, ms[e]));
        CK(h, cudaStreamWaitEvent(h->ahead, ev, 0));
      }
    }
    // ... and behind the in-batch maintenance of the batch the buffer last held
    if (h->ev_ts_end[bi]) CK(h, cudaStreamWaitEvent(h->ahead, h->ev_ts_end[bi], 0));
  }
  b.own_slot.assign(cnt, -1);
  b.own_epoch.assign(cnt, 0);
  struct StreamGuard {
    mogp_engine* h;
    ~StreamGuard() { h->stream = h->main_stream; }
  } guard{h};
  h->stream = h->ahead2;
  TRY(win_launch_own_batch(h, s0, s0 + cnt, false, thr_index, d_own, b.own_slot.data()));
  for (int64_t v = 0; v < cnt; ++v)
    if (b.own_slot[v] >= 0) b.own_epoch[v] = w.slot_epoch[b.own_slot[v]];
  k_new_ll<<<(unsigned)cnt, 64, 0, h->stream>>>(h->d_woff + s0, h->d_wq + cb, h->d_wq + Rtot + cb, h->d_wq + 2 * Rtot + s0,
                                                newcomer_theta(ch.cfg), d_new);
  LAUNCHED(h);
  CK(h, cudaEventRecord(h->ev_ahead2, h->ahead2));
  h->stream = h->ahead;
  b.pl = ScorePlan();
  bf.res.zigzag = pipeline_depth(h) >= 2 ? env_int("MOGP_ZIGZAG", 0) : 0;
  // the monotonicity rule applied on the device before the O(m^2) product (MOGP_PRUNE=0: after it, on the host only)
  PruneSpec prune{h->d_wq + 2 * Rtot + w.n + s0, ch.cfg.threshold};
  const bool do_prune = ch.cfg.use_threshold && env_int("MOGP_PRUNE", 1) != 0;
  if (ns > 0)
    TRY(score_device(h, cnt, R, h->d_woff + s0, h->d_wq + cb, h->d_wq + Rtot + cb, ns, b.slots.data(), thr_index, d_score, d_mu,
                     nullptr, nullptr, nullptr, true, &b.pl, &bf.res, do_prune ? &prune : nullptr));
  CK(h, cudaEventRecord(bf.scored, h->ahead));  // (corrections queued ahead for this batch start from here, replay_launch)
  CK(h, cudaStreamWaitEvent(h->ahead, h->ev_ahead2, 0));
  CK(h, cudaMemcpyAsync(bf.h_out, bf.d_out, sizeof(double) * total, cudaMemcpyDeviceToHost, h->ahead));
  CK(h, cudaEventRecord(h->blocking_sync ? bf.done_blk : bf.done, h->ahead));
  CK(h, cudaGetLastError());
  h->bytes_d2h += (int64_t)sizeof(double) * total;
  w.end_launched = s0 + cnt;
  h->pass_launched += 1;
  ch.win_passes += 1;
  ch.win_patients += cnt;
  return MOGP_OK;
}

// Wait for the pass of batch `bi` and take over what is still current: scores under clusters that have not changed
// since the pass was launched (the others have been, or will be, re-scored), own-cluster scores likewise.
int batch_absorb(mogp_engine* h, int bi) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  Chain::Batch& b = w.b[bi];
  BatchBuf& bf = h->bb[bi];
  CK(h, cudaEventSynchronize(h->blocking_sync ? bf.done_blk : bf.done));
  h->pass_absorbed += 1;
  pool_unhold(h, false);
  const int ns = (int)b.slots.size();
  const int64_t cnt = b.cnt, cb = w.col0[b.s0];
  const double* hs = bf.h_out;
  const double* hm = hs + cnt * ns;
  const double* hn = hm + cnt * ns;
  const double* ho = hn + cnt;
  int64_t kx_off = 0;
  for (int si = 0; si < ns; ++si) {
    const int32_t slot = b.slots[si];
    const int64_t mpad = b.mpad[si];
    const bool clean = w.slot_epoch[slot] == b.epoch[si];
    // (a slot that changed since the pass was launched keeps the pass's T / ss / mu for win_replay, not its scores)
    if (clean || w.slot_upto[slot] <= b.s0) {
      for (int64_t v = 0; v < cnt; ++v) {
        Chain::WinEntry& e = w.e[b.s0 + v][slot];
        if (clean) {
          e.score = hs[v * ns + si];
          e.muthr = hm[v * ns + si];
        }
        e.pruned = hs[v * ns + si] == -INFINITY;  // (a computed score is finite)
        e.stride = mpad;
        e.T = b.pl.T + kx_off + (w.col0[b.s0 + v] - cb) * mpad;
        e.mu = b.pl.mu + (int64_t)si * b.pl.Rpad + (w.col0[b.s0 + v] - cb);
        e.ss = b.pl.ss + (int64_t)si * b.pl.Rpad + (w.col0[b.s0 + v] - cb);
      }
    }
    if (clean) {
      w.slot_upto[slot] = b.s0 + cnt;
      w.slot_T_ok[slot] = 1;
    }
    kx_off += b.pl.Rpad * mpad;
  }
  // accounting: pairs the rule excluded before the product, and the columns (flops) the pass did not compute for them
  for (int si = 0; si < ns; ++si)
    for (int64_t v = 0; v < cnt; ++v) {
      ch.win_pairs += 1;
      if (hs[v * ns + si] != -INFINITY) continue;
      ch.win_pruned += 1;
      const double mm = (double)b.msz[si], ni = (double)(w.col0[b.s0 + v + 1] - w.col0[b.s0 + v]);
      if (b.prof && h->prof_on) h->prof_flops_c[0] -= (ni - 1.0) * mm * (mm + 10.0);
    }
  for (int64_t v = 0; v < cnt; ++v) {
    w.new_ll[b.s0 + v] = hn[v];
    if (b.own_slot[v] >= 0 && w.slot_epoch[b.own_slot[v]] == b.own_epoch[v]) {
      w.own_ll[b.s0 + v] = ho[3 * v];
      w.own_mu[b.s0 + v] = ho[3 * v + 1];
      w.own_flag[b.s0 + v] = ho[3 * v + 2];
      w.own_ok[b.s0 + v] = 1;
    }
  }
  return MOGP_OK;
}

// The pass of the NEXT batch is launched before the current batch makes its moves, so every cluster a move changes is
// stale in it.  Re-scoring those clusters at the next batch's start costs a wait for the factor updates plus O(m^2)
// products that compete with the pass after next.  Instead, the rank-n correction a move applies to the positions still
// ahead in ITS batch (k_alg_leave / k_alg_join) is queued a second time, for the positions of the next batch, on a
// high-priority stream right behind that batch's pass: it works from the T / ss / mu the pass leaves on the device (the
// pass saw the cluster before the change, which is what the formulas start from) and from copies of what the move
// itself contributes (the block columns and alpha_B of a leave, Ls^-1 and z_B of a join), and it leaves T / ss / mu as a
// pass over the changed cluster would (maintain), so that the next change of the same cluster can be chained behind it
// and the batch, once absorbed, finds current columns.  By the time the batch is absorbed the corrected scores are in
// pinned memory: no launch, no wait.  A cluster whose chain breaks (new / emptied, no room, a change that could not be
// expressed) is re-scored the long way as before.
int replay_launch(mogp_engine* h, const ChangeLog& L, cudaEvent_t dep, int nbi, AlgJobs* Jout, int* npos_out, bool* queued) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  *queued = false;
  const int32_t slot = L.slot;
  if (slot < 0 || slot >= (int32_t)w.rp_epoch[nbi].size()) return MOGP_OK;
  auto broken = [&]() {
    w.rp_epoch[nbi][slot] = -1;
    return MOGP_OK;
  };
  if (w.rp_epoch[nbi][slot] < 0 || w.rp_epoch[nbi][slot] != L.epoch_before) return broken();
  const Chain::Batch& nb = w.b[nbi];
  const BatchBuf& nbf = h->bb[nbi];
  if (nb.cnt <= 0 || nb.cnt > OWN_BATCH || (int)w.async_items[nbi].size() >= REPLAY_MAX || L.ntiles <= 0) return broken();
  int si = -1;
  for (size_t q = 0; q < nb.slots.size(); ++q)
    if (nb.slots[q] == slot) si = (int)q;
  if (si < 0) return broken();  // not in that pass
  const int64_t wneed = nb.cnt * (int64_t)L.ntiles * (L.kind == 0 ? ALG_ENTRIES : NP * NP);
  if (w.rpw_used[nbi] + wneed > h->rpw_share) return broken();
  if (L.kind == 1 && (int64_t)L.J.m + L.J.n > nb.mpad[si]) return broken();  // no room for the rows the join adds
  AlgJobs J = L.J;
  J.maintain = 1;
  int64_t kx_off = 0;
  for (int q = 0; q < si; ++q) kx_off += nb.pl.Rpad * nb.mpad[q];
  const int64_t cb = w.col0[nb.s0];
  const int ns = (int)nb.slots.size();
  for (int64_t q = 0; q < nb.cnt; ++q) {
    const int64_t v = nb.s0 + q, c = w.col0[v] - cb;
    AlgPos& P = J.pos[q];
    P.stride = nb.mpad[si];
    P.T = nb.pl.T + kx_off + c * P.stride;
    P.mu = nb.pl.mu + (int64_t)si * nb.pl.Rpad + c;
    P.ss = nb.pl.ss + (int64_t)si * nb.pl.Rpad + c;
    P.x = h->d_wq + w.col0[v];
    P.y = h->d_wq + w.col0[w.n] + w.col0[v];
    P.ncols = (int)(w.col0[v + 1] - w.col0[v]);
    P.thr = J.thr_index;
    P.flag = nbf.d_out + q * ns + si;  // the pair's score as the pass leaves it: -inf = excluded before the product
  }
  const int k = (int)w.async_items[nbi].size();
  const int64_t out_off = ((int64_t)nbi * REPLAY_MAX + k) * 2 * OWN_BATCH;
  J.out = h->d_replay + out_off;
  J.logB = J.logAlpha = nullptr;
  double* part = h->d_rpw + (int64_t)nbi * h->rpw_share + w.rpw_used[nbi];
  w.rpw_used[nbi] += wneed;
  cudaStream_t tq = h->tq[nbi];
  // behind the pass they correct (its kernels, not its read-back), behind what the move contributes (logged copies /
  // Ls^-1, z_B) and behind whatever last touched these columns
  CK(h, cudaStreamWaitEvent(tq, nbf.scored, 0));
  CK(h, cudaStreamWaitEvent(tq, dep, 0));
  TRY(t_wait(h, tq, nbi, slot));
  if (L.kind == 0) {
    k_alg_leave_part<<<dim3((unsigned)L.ntiles, (unsigned)nb.cnt), 256, 0, tq>>>(J, part);
    LAUNCHED(h);
    k_alg_leave_fin<<<(unsigned)nb.cnt, 256, 0, tq>>>(J, part, L.ntiles);
    LAUNCHED(h);
  } else {
    k_alg_join_part<<<dim3((unsigned)L.ntiles, (unsigned)nb.cnt), 256, 0, tq>>>(J, part);
    LAUNCHED(h);
    k_alg_join_fin<<<(unsigned)nb.cnt, 256, 0, tq>>>(J, part, L.ntiles);
    LAUNCHED(h);
  }
  TRY(t_mark(h, tq, nbi, slot));
  CK(h, cudaMemcpyAsync(h->h_replay + out_off, J.out, sizeof(double) * 2 * OWN_BATCH, cudaMemcpyDeviceToHost, tq));
  CK(h, cudaEventRecord(h->ev_replay[nbi], tq));
  h->bytes_d2h += (int64_t)sizeof(double) * 2 * OWN_BATCH;
  Chain::Window::AsyncItem it;
  it.slot = slot;
  it.epoch_before = L.epoch_before;
  it.out_idx = k;
  w.async_items[nbi].push_back(it);
  w.rp_epoch[nbi][slot] = L.epoch_before + 1;
  w.rp_last[nbi][slot] = k;
  if (Jout) *Jout = J;
  if (npos_out) *npos_out = (int)nb.cnt;
  *queued = true;
  return MOGP_OK;
}

// Take over the corrections queued behind the pass of the batch just absorbed (see replay_launch): a cluster that changed
// since the pass was launched and whose every change was chained behind the pass has current scores (the last
// correction's) and current T / ss / mu.
int win_replay(mogp_engine* h, int bi) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  Chain::Batch& b = w.b[bi];
  if (w.async_items[bi].empty()) return MOGP_OK;
  {
    PhaseTimer pt(&ch.t_chain);
    CK(h, cudaEventSynchronize(h->ev_replay[bi]));
  }
  ch.win_chained += (int64_t)w.async_items[bi].size();
  for (size_t si = 0; si < b.slots.size(); ++si) {
    const int32_t slot = b.slots[si];
    if (slot < 0 || slot >= (int32_t)w.slot_epoch.size() || !h->cl[slot].used) continue;
    if (w.slot_epoch[slot] == b.epoch[si]) continue;  // unchanged: batch_absorb took the pass's own scores
    if (w.rp_epoch[bi][slot] != w.slot_epoch[slot] || w.rp_last[bi][slot] < 0 || w.slot_upto[slot] > b.s0) continue;
    const double* out = h->h_replay + ((int64_t)bi * REPLAY_MAX + w.rp_last[bi][slot]) * 2 * OWN_BATCH;
    bool reopened = false;
    for (int64_t q = 0; q < b.cnt; ++q) {
      Chain::WinEntry& e = w.e[b.s0 + q][slot];
      e.muthr = out[2 * q + 1];
      if (!e.pruned) {
        e.score = out[2 * q];
        continue;
      }
      e.score = -INFINITY;  // excluded before the product: stays excluded unless the moves brought its mean back
      const double yt = h->h_ythr.empty() ? NAN : h->h_ythr[w.pat[b.s0 + q]];
      if (!(yt - e.muthr > ch.cfg.threshold)) reopened = true;
    }
    if (reopened) continue;  // (the cluster stays stale and is re-scored at the batch's first step)
    w.slot_upto[slot] = b.s0 + b.cnt;
    w.slot_T_ok[slot] = 1;
    ch.win_replayed += 1;
  }
  w.async_items[bi].clear();
  w.rpw_used[bi] = 0;
  return MOGP_OK;
}

// Cache entries of positions [v0, batch_end) under `slots`, from a re-scoring pass.
void win_store(mogp_engine* h, int64_t v0, const std::vector<int32_t>& slots, const ScorePlan& pl, const double* host_score,
               const double* host_mu) {
  Chain::Window& w = h->chain.win;
  const int ns = (int)slots.size();
  const int64_t cbase = w.col0[v0];
  int64_t kx_off = 0;
  for (int si = 0; si < ns; ++si) {
    const int32_t slot = slots[si];
    const int64_t mpad = round_up(h->cl[slot].m, TB);
    for (int64_t v = v0; v < w.batch_end; ++v) {
      if ((size_t)slot >= w.e[v].size()) w.e[v].resize(h->cl.size());
      Chain::WinEntry& e = w.e[v][slot];
      e.score = host_score[(v - v0) * ns + si];
      e.muthr = host_mu[(v - v0) * ns + si];
      e.pruned = false;
      e.stride = mpad;
      e.T = pl.T + kx_off + (w.col0[v] - cbase) * mpad;
      e.mu = pl.mu + (int64_t)si * pl.Rpad + (w.col0[v] - cbase);
      e.ss = pl.ss + (int64_t)si * pl.Rpad + (w.col0[v] - cbase);
    }
    w.slot_upto[slot] = w.batch_end;
    w.slot_T_ok[slot] = 1;
    w.slot_alg[slot] = 0;
    w.slot_tev[w.cur_bi][slot] = -1;  // fresh columns (the host has waited for them)
    kx_off += pl.Rpad * mpad;
  }
}

// Bring the cache up to date for the current patient and, in the same launches, for the rest of its batch: re-score
// against the clusters whose cached scores are stale, refresh stale own-cluster scores.  (The positions of the next
// batch, whose pass may be in flight, are brought up to date when their batch starts.)
int win_ensure(mogp_engine* h, int32_t own_id, ScoreRes* arena) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  const int64_t cur = w.cur, end = w.batch_end;
  win_size_slots(w, h->cl.size());
  std::vector<int32_t> bad;
  for (size_t k = 0; k < ch.Nk.size(); ++k) {
    if (ch.Nk[k] <= 0 || (int32_t)k == own_id) continue;
    const int32_t slot = ch.slot_of_id[k];
    if (w.slot_upto[slot] <= cur) bad.push_back(slot);
  }
  const bool need_own = own_id >= 0 && !w.own_ok[cur];
  if (bad.empty() && !need_own) return MOGP_OK;
  const int nb = (int)bad.size();
  const int64_t rem = end - cur, Rrem = w.col0[end] - w.col0[cur], Rtot = w.col0[w.n];
  const int thr_index = ch.cfg.use_threshold ? ch.cfg.thr_index : -1;
  const int64_t total = 2 * rem * nb + 3 * rem;
  TRY(ensure(h, &h->d_out, &h->out_elems, total + 8));
  double* d_score = h->d_out;
  double* d_mu = d_score + rem * nb;
  double* d_own = d_mu + rem * nb;
  std::vector<int32_t> own_slot(rem, -1);
  {
    SideScope side(h);
    TRY(win_launch_own_batch(h, cur, end, true, thr_index, d_own, own_slot.data()));
  }
  ScorePlan pl;
  if (nb > 0) {
    TRY(score_device(h, rem, Rrem, h->d_woff + cur, h->d_wq + w.col0[cur], h->d_wq + Rtot + w.col0[cur], nb, bad.data(), thr_index,
                     d_score, d_mu, nullptr, nullptr, nullptr, true, &pl, arena));
    ch.win_rescores += 1;
  }
  join_side(h);
  CK(h, cudaGetLastError());
  std::vector<double> host(total);
  TRY(copy_back(h, host.data(), d_score, total));
  if (nb > 0) win_store(h, cur, bad, pl, host.data(), host.data() + rem * nb);
  for (int64_t v = cur; v < end; ++v)
    if (own_slot[v - cur] >= 0) {
      const double* o = host.data() + 2 * rem * nb + 3 * (v - cur);
      w.own_ll[v] = o[0];
      w.own_mu[v] = o[1];
      w.own_flag[v] = o[2];
      w.own_ok[v] = 1;
    }
  if (need_own && !w.own_ok[cur]) FAIL(h, MOGP_ESTATE, "own-cluster score of patient %lld unavailable", (long long)w.pat[cur]);
  return MOGP_OK;
}

// Test mode (MOGP_CHECK_CORRECTIONS): re-score positions [v0, batch_end) under `slots` the long way (from the UPDATED
// factors, waiting for them) and record the worst disagreement with the cached - corrected - scores.  take_over (= 1): the
// re-scored values and columns replace the cached ones; 2: compare only, so that chains of corrections on maintained
// columns are checked at every link.
int win_check_slots(mogp_engine* h, int64_t v0, const std::vector<int32_t>& chk, ScoreRes* arena, bool take_over) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  const int64_t rem = w.batch_end - v0;
  if (chk.empty() || rem <= 0) return MOGP_OK;
  const int64_t Rrem = w.col0[w.batch_end] - w.col0[v0], Rtot = w.col0[w.n];
  const int nb = (int)chk.size();
  const int thr_index = ch.cfg.use_threshold ? ch.cfg.thr_index : -1;
  TRY(ensure(h, &h->d_out, &h->out_elems, 2 * rem * nb + 8));
  join_side(h);
  ScorePlan pl;
  TRY(score_device(h, rem, Rrem, h->d_woff + v0, h->d_wq + w.col0[v0], h->d_wq + Rtot + w.col0[v0], nb, chk.data(), thr_index,
                   h->d_out, h->d_out + rem * nb, nullptr, nullptr, nullptr, true, &pl, arena));
  std::vector<double> host(2 * rem * nb);
  TRY(copy_back(h, host.data(), h->d_out, 2 * rem * nb));
  for (int si = 0; si < nb; ++si)
    for (int64_t q = 0; q < rem; ++q) {
      const Chain::WinEntry& e = w.e[v0 + q][chk[si]];
      const double s_ref = host[q * nb + si], m_ref = host[rem * nb + q * nb + si];
      double err = e.pruned ? 0.0 : fabs(e.score - s_ref) / std::max(1.0, fabs(s_ref));
      if (!(isnan(m_ref) && isnan(e.muthr))) err = std::max(err, fabs(e.muthr - m_ref) / std::max(1.0, fabs(m_ref)));
      if (!(err <= ch.max_corr_err)) ch.max_corr_err = err;  // (NaN sticks)
    }
  if (take_over) {
    const int64_t keep_cur = w.cur;
    w.cur = v0;  // win_store fills [v0, batch_end)
    win_store(h, v0, chk, pl, host.data(), host.data() + rem * nb);
    w.cur = keep_cur;
  }
  ch.corr_checked += rem * nb;
  return MOGP_OK;
}

// One Gibbs step of the patient at position w.cur from the cache (begin / score / draw / commit).
int win_step(mogp_engine* h, double a_draw, double u, int32_t* chosen_out, ScoreRes* arena) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  const int64_t n = w.pat[w.cur];
  const int32_t zn = ch.z[n];
  const int32_t prev_slot = (zn >= 0 && zn < (int32_t)ch.slot_of_id.size()) ? ch.slot_of_id[zn] : -1;
  TRY(chain_begin(h, n));
  const int32_t prev = ch.pending_prev;
  if (ch.slot_of_id[prev] < 0 || !ch.pending_lazy) win_invalidate(ch, prev_slot);  // freed, or rebuilt without the patient
  const int32_t n_ids = (int32_t)ch.Nk.size();
  ch.act_ids.clear();
  for (int32_t k = 0; k < n_ids; ++k)
    if (ch.Nk[k] > 0) ch.act_ids.push_back(k);
  ch.act_ids.push_back(n_ids);  // alpha > 0 in this mode: the new-cluster slot is always active
  const int nact = (int)ch.act_ids.size();
  ch.act_score.assign(nact, 0.0);
  const int32_t own = ch.pending_lazy ? prev : -1;
  {
    PhaseTimer pt(&ch.t_ensure);
    TRY(win_ensure(h, own, arena));
  }
  const double ythr = (ch.cfg.use_threshold && !h->h_ythr.empty()) ? h->h_ythr[n] : NAN;
  for (int i = 0; i < nact; ++i) {
    const int32_t id = ch.act_ids[i];
    double prior, ll, muthr = NAN;
    if (i == nact - 1) {
      prior = log(ch.cfg.alpha + 1e-16);  // allocmodel.py:25-26
      ll = w.new_ll[w.cur];
      if (!isfinite(ll)) FAIL(h, MOGP_ENOTPD, "new-cluster covariance not positive definite");
    } else {
      prior = log((double)ch.Nk[id] + 1e-16);
      if (id == own) {
        if (w.own_flag[w.cur] != 1.0) FAIL(h, MOGP_ENOTPD, "leave-out block not positive definite");
        ll = w.own_ll[w.cur];
        muthr = w.own_mu[w.cur];
      } else {
        const Chain::WinEntry& e = w.e[w.cur][ch.slot_of_id[id]];
        ll = e.score;
        muthr = e.muthr;
      }
      if (ch.cfg.use_threshold && (ythr - muthr > ch.cfg.threshold)) ll = -INFINITY;  // monotonicity rule (:254-263)
    }
    ch.act_score[i] = prior + ll;
  }
  int idx = 0;
  std::vector<double> pvec(h->record_probs ? nact : 0);
  int rc = host_choice(ch.act_score.data(), nact, u, h->record_probs ? pvec.data() : nullptr, &idx, &ch.last_margin);
  if (rc != MOGP_OK) FAIL(h, rc, "probabilities contain NaN (patient %lld)", (long long)n);
  record_step_probs(h, ch.act_ids, pvec.data());
  const int32_t chosen = ch.act_ids[idx];
  const bool stays = ch.pending_lazy && chosen == prev;
  // A move changes two clusters.  The scores of the patients still ahead in this batch under them are corrected by
  // rank-n formulas from what the pass left on the device (k_alg_leave / k_alg_join) and read back at once; the
  // O(m^2) updates of the two factors are queued behind and nobody waits for them.  With MOGP_TMAINT (default) the
  // corrections also leave the cached T / ss / mu current (the leave's T by k_T_leave, fired from remove_core), for this
  // batch's positions and - queued behind its pass - for the next batch's (replay_launch): further changes of the same
  // cluster are corrected the same way and the next batch starts with current scores.  Without it: one correction per
  // cluster and batch, and everything else - a cluster without current T, a new or emptied cluster - goes the long way
  // (re-scoring against the updated factor).
  const int64_t npos = w.batch_end - (w.cur + 1);
  const bool alg_on = !stays && npos >= 0 && npos <= OWN_BATCH && ch.pending_lazy && env_int("MOGP_RANK_RESCORE", 1) != 0;
  const bool tm = tmaint_on(h);
  const bool log_on = alg_on && tm && !w.flight.empty();  // passes of later batches are in flight
  const int thr_cfg = ch.cfg.use_threshold ? ch.cfg.thr_index : -1;
  auto alg_ok = [&](int32_t slot) {
    return slot >= 0 && slot < (int32_t)w.slot_upto.size() && w.slot_upto[slot] >= w.batch_end && w.slot_T_ok[slot] &&
           (tm || !w.slot_alg[slot]);
  };
  auto chain_ok = [&](int f, int32_t slot) {  // every change since the pass in buffer f was launched is chained behind it
    return slot >= 0 && slot < (int32_t)w.rp_epoch[f].size() && w.rp_epoch[f][slot] >= 0 &&
           w.rp_epoch[f][slot] == w.slot_epoch[slot];
  };
  auto replay_ready = [&](int32_t slot) {
    if (!log_on) return false;
    for (int f : w.flight)
      if (chain_ok(f, slot)) return true;
    return false;
  };
  auto fill_pos = [&](AlgJobs& J, int32_t slot) {
    for (int64_t q = 0; q < npos; ++q) {
      const int64_t v = w.cur + 1 + q;
      const Chain::WinEntry& e = w.e[v][slot];
      AlgPos& P = J.pos[q];
      P.stride = e.stride;
      // a pair excluded before the product has its threshold column only: the correction re-derives that mean
      const int64_t skip = e.pruned ? thr_cfg : 0;
      P.T = e.T + skip * P.stride;
      P.ss = e.ss + skip;
      P.mu = e.mu + skip;
      P.x = h->d_wq + w.col0[v] + skip;
      P.y = h->d_wq + w.col0[w.n] + w.col0[v] + skip;
      P.ncols = e.pruned ? 1 : (int)(w.col0[v + 1] - w.col0[v]);
      P.thr = e.pruned ? 0 : thr_cfg;
      P.flag = nullptr;
    }
  };
  const int64_t c0n = h->h_off[n], nin = h->h_off[n + 1] - c0n;
  int32_t a_slot = -1, b_slot = -1;
  bool a_alg = false, a_maint = false, a_rp = false;
  AlgHook hook{};
  LeaveHook lhook{};
  if (alg_on) {
    const int32_t sl_a = ch.slot_of_id[prev];
    const bool a0 = alg_ok(sl_a), a1 = replay_ready(sl_a);
    if (sl_a >= 0 && (a0 || a1)) {  // leave side, from the factor as it is now
      a_slot = sl_a;
      Cluster& ca = h->cl[a_slot];
      const int64_t p = member_offset(h, ca, n, nullptr);
      if (p >= 0) {
        const int64_t np0 = a0 ? npos : 0;
        AlgJobs J{};
        if (np0 > 0) fill_pos(J, a_slot);
        J.th = make_theta(ca);
        J.thr_index = thr_cfg;
        J.out = h->d_alg;
        J.M = ca.M;
        J.alpha = ca.alpha();
        J.ld = ca.cap;
        J.p = (int)p;
        J.n = (int)nin;
        J.m = (int)ca.m;
        const int ntiles = (int)((ca.m - p + TB - 1) / TB);
        if (tm) {  // keep what the kernels read from the cluster: the buffer moves on, k_T_leave and the replay come later
          const int64_t need = (int64_t)(ca.m - p) * NP + NP;
          if (w.blog_used + need <= w.blog_half) {
            J.logB = h->d_blog + w.blog_base + w.blog_used;
            J.logAlpha = J.logB + (int64_t)(ca.m - p) * NP;
            w.blog_used += need;
          }
        }
        J.maintain = (tm && J.logB) ? 1 : 0;
        if (np0 > 0 || (J.logB && a1)) {
          TRY(ensure(h, &h->d_algw, &h->algw_elems, (int64_t)std::max<int64_t>(np0, 1) * ntiles * ALG_ENTRIES));
          if (np0 == 0) J.pos[0].ncols = 0;  // log only
          if (np0 > 0) TRY(t_wait(h, h->quick2, w.cur_bi, a_slot));
          k_alg_leave_part<<<dim3((unsigned)ntiles, (unsigned)std::max<int64_t>(np0, 1)), 256, 0, h->quick2>>>(J, h->d_algw);
          LAUNCHED(h);
          if (np0 > 0) {
            k_alg_leave_fin<<<(unsigned)np0, 256, 0, h->quick2>>>(J, h->d_algw, ntiles);
            LAUNCHED(h);
            if (J.maintain) TRY(t_mark(h, h->quick2, w.cur_bi, a_slot));
          }
          CK(h, cudaEventRecord(h->ev_q2, h->quick2));
          // the cluster's old buffer must outlive these reads: the streams that may recycle it wait for them
          CK(h, cudaStreamWaitEvent(h->main_stream, h->ev_q2, 0));
          CK(h, cudaStreamWaitEvent(h->side, h->ev_q2, 0));
        }
        a_alg = a0;
        AlgJobs Jlog = J;  // the same job reading the logged copies: J.M[k * ld + p + a] with ld = NP reads logB[(k - p) * NP + a]
        if (J.logB) {
          Jlog.M = J.logB - ((int64_t)p * NP + p);
          Jlog.ld = NP;
          Jlog.alpha = J.logAlpha - p;
          Jlog.logB = Jlog.logAlpha = nullptr;
        }
        if (J.logB && a1) {
          ChangeLog cl;
          cl.slot = a_slot;
          cl.kind = 0;
          cl.J = Jlog;
          cl.ntiles = ntiles;
          cl.epoch_before = w.slot_epoch[a_slot];
          int t = 1;
          for (int f : w.flight) {
            if (!chain_ok(f, a_slot)) continue;
            bool queued = false;
            if (t < NTGT) TRY(replay_launch(h, cl, h->ev_q2, f, &lhook.J[t], &lhook.npos[t], &queued));
            if (!queued) {
              w.rp_epoch[f][a_slot] = -1;
              continue;
            }
            lhook.bi[t++] = f;
            a_rp = true;
          }
        } else if (a1) {
          for (int f : w.flight) w.rp_epoch[f][a_slot] = -1;
        }
        if (J.maintain && (np0 > 0 || a_rp)) {  // the T of the positions ahead follows in remove_core (k_T_leave)
          lhook.slot = a_slot;
          if (np0 > 0) {
            lhook.J[0] = Jlog;
            lhook.npos[0] = (int)np0;
            lhook.bi[0] = w.cur_bi;
          }
          h->leave_hook = &lhook;
        }
        a_maint = J.maintain != 0;
      }
    }
    if (chosen < n_ids && alg_ok(ch.slot_of_id[chosen]) && (npos > 0 || log_on)) {  // join side: launched by append_core once Ls, z_B exist
      b_slot = ch.slot_of_id[chosen];
      const Chain::WinEntry& en = w.e[w.cur][b_slot];
      fill_pos(hook.jobs, b_slot);
      hook.jobs.th = make_theta(h->cl[b_slot]);
      hook.jobs.thr_index = thr_cfg;
      hook.jobs.out = h->d_alg + 2 * OWN_BATCH;
      hook.jobs.n = (int)nin;
      hook.jobs.m = (int)h->cl[b_slot].m;
      hook.jobs.Tn = en.T;
      hook.jobs.stride_n = en.stride;
      hook.jobs.xn = h->d_px + c0n;
      bool room = tm;  // the n rows the append adds to every cached column must fit under its stride
      for (int64_t q = 0; q < npos; ++q) room = room && (h->cl[b_slot].m + nin <= hook.jobs.pos[q].stride);
      hook.jobs.maintain = room ? 1 : 0;
      hook.npos = (int)npos;
      hook.slot = b_slot;
      hook.bi = w.cur_bi;
      h->alg_hook = &hook;
    }
  }
  int32_t is_new = 0;
  {
    PhaseTimer pt(&ch.t_commit);
    w.in_step = true;
    rc = chain_commit(h, chosen, a_draw, &is_new);
    w.in_step = false;
    h->alg_hook = nullptr;
    h->leave_hook = nullptr;
    if (rc != MOGP_OK) return rc;
  }
  if (!stays) {
    const bool b_alg = hook.fired;
    for (int t = 1; t < NTGT; ++t)  // a correction was queued but T did not follow
      if (lhook.npos[t] > 0 && !lhook.fired[t]) w.rp_epoch[lhook.bi[t]][a_slot] = -1;
    if (b_slot >= 0 && log_on)
      for (int f : w.flight) {
        if (!chain_ok(f, b_slot)) continue;
        bool queued = false;
        if (b_alg && hook.on_quick) {
          ChangeLog cl;
          cl.slot = b_slot;
          cl.kind = 1;
          cl.J = hook.jobs;  // Tn, Li, znew, xn stay valid until the passes in flight have been absorbed
          cl.ntiles = hook.ntiles;
          cl.epoch_before = w.slot_epoch[b_slot];
          TRY(replay_launch(h, cl, h->ev_alg, f, nullptr, nullptr, &queued));
        }
        if (!queued) w.rp_epoch[f][b_slot] = -1;
      }
    if ((a_alg || b_alg) && npos > 0) {
      PhaseTimer pt(&ch.t_alg);
      if (b_alg) CK(h, cudaStreamWaitEvent(h->quick, h->ev_alg, 0));
      if (a_alg) CK(h, cudaStreamWaitEvent(h->quick, h->ev_q2, 0));
      CK(h, cudaMemcpyAsync(h->h_alg, h->d_alg, sizeof(double) * 4 * OWN_BATCH, cudaMemcpyDeviceToHost, h->quick));
      CK(h, sync_helper_stream(h, h->quick));
      h->bytes_d2h += (int64_t)sizeof(double) * 4 * OWN_BATCH;
      ch.win_algebra += 1;
    }
    // scores current; T / ss / mu current too when the change was carried through to them (maintained), else stale
    auto corrected = [&](int32_t slot, const double* out, bool maintained) {
      bool reopened = false;
      for (int64_t q = 0; q < npos; ++q) {
        Chain::WinEntry& e = w.e[w.cur + 1 + q][slot];
        e.muthr = out[2 * q + 1];
        if (!e.pruned) {
          e.score = out[2 * q];
          continue;
        }
        // excluded before the product: only the mean was corrected.  Still excluded -> still -inf; if the move has
        // brought the mean back within the threshold the pair needs a real score: the slot is re-scored the long way
        const int64_t pn = w.pat[w.cur + 1 + q];
        const double yt = h->h_ythr.empty() ? NAN : h->h_ythr[pn];
        if (!(yt - e.muthr > ch.cfg.threshold)) reopened = true;
      }
      if (!maintained) {
        w.slot_T_ok[slot] = 0;
        w.slot_alg[slot] = 1;
      }
      w.slot_epoch[slot] += 1;
      if (reopened) {
        w.slot_upto[slot] = 0;
        w.slot_T_ok[slot] = 0;
      }
      for (int64_t v = w.cur + 1; v < w.end_launched; ++v) {
        const int32_t id = ch.z[w.pat[v]];
        if (id >= 0 && ch.slot_of_id[id] == slot) w.own_ok[v] = 0;
      }
    };
    const int32_t slot_prev = ch.slot_of_id[prev];
    if (a_alg && slot_prev == a_slot) corrected(a_slot, h->h_alg, a_maint && (npos == 0 || lhook.fired[0]));
    else if (slot_prev >= 0) win_invalidate(ch, slot_prev);
    if (b_alg && ch.slot_of_id[chosen] == b_slot) corrected(b_slot, h->h_alg + 2 * OWN_BATCH, hook.jobs.maintain != 0);
    else win_invalidate(ch, ch.slot_of_id[chosen]);
    if (npos > 0 && env_int("MOGP_CHECK_CORRECTIONS", 0) != 0) {
      std::vector<int32_t> chk;
      if (a_alg && slot_prev == a_slot) chk.push_back(a_slot);
      if (b_alg && ch.slot_of_id[chosen] == b_slot) chk.push_back(b_slot);
      TRY(win_check_slots(h, w.cur + 1, chk, arena, env_int("MOGP_CHECK_CORRECTIONS", 0) == 1));
    }
  }
  if (chosen_out) *chosen_out = chosen;
  std::vector<Chain::WinEntry>().swap(w.e[w.cur]);
  w.cur += 1;
  return MOGP_OK;
}

// The whole sweep through the look-ahead pipeline; patients a batch cannot take go through the single-step path.
int sweep_pipelined(mogp_engine* h, int64_t n_steps, const int64_t* order, const double* a_draw, const double* u,
                    int32_t* chosen_ids, double* mm, int max_cols) {
  Chain& ch = h->chain;
  Chain::Window& w = ch.win;
  TRY(win_prepare(h, n_steps, order, a_draw));
  const int depth = pipeline_depth(h);
  int64_t pos = 0;          // position being processed
  int64_t launch_pos = 0;   // first position without a pass
  int64_t jl = 0;           // passes launched so far: pass j lives in buffer j % NBUF
  std::deque<int> pending;  // buffers of the passes launched and not yet absorbed, oldest first
  while (pos < n_steps) {
    if (pending.empty()) {
      const int64_t cnt = batch_form(h, pos, max_cols);
      if (cnt == 0) {  // no visits / too many for a rank update: patient-by-patient step; the cache starts over
        int32_t chosen = -1;
        w.cur = w.end_launched = pos + 1;
        w.flight.clear();
        TRY(mogp_chain_step(h, order[pos], a_draw ? a_draw[pos] : 0.0, u[pos], &chosen, nullptr, nullptr, nullptr, nullptr, 0));
        win_invalidate_all(ch);
        if (chosen_ids) chosen_ids[pos] = chosen;
        *mm = std::min(*mm, ch.last_margin);
        ++pos;
        launch_pos = pos;
        continue;
      }
      w.cur = pos;
      w.flight.clear();
      TRY(batch_launch(h, (int)(jl % NBUF), pos, cnt));
      pending.push_back((int)(jl % NBUF));
      ++jl;
      launch_pos = pos + cnt;
    }
    const int bi = pending.front();
    pending.pop_front();
    Chain::Batch& B = w.b[bi];
    {
      PhaseTimer pt(&ch.t_open);
      TRY(batch_absorb(h, bi));
    }
    w.cur = pos;
    w.batch_end = pos + B.cnt;
    w.cur_bi = bi;
    w.flight.assign(pending.begin(), pending.end());
    // The passes of the next batches go out now, before this batch makes its moves: every cluster a move changes is stale
    // in them and is brought up to date by corrections chained behind them (or re-scored when its batch starts).
    while ((int)pending.size() < depth - 1 && launch_pos < n_steps) {
      const int64_t cnt2 = batch_form(h, launch_pos, max_cols);
      if (cnt2 <= 0) break;
      TRY(batch_launch(h, (int)(jl % NBUF), launch_pos, cnt2));
      pending.push_back((int)(jl % NBUF));
      w.flight.push_back((int)(jl % NBUF));
      ++jl;
      launch_pos += cnt2;
    }
    TRY(win_replay(h, bi));
    if (env_int("MOGP_CHECK_CORRECTIONS", 0) == 2) {  // the scores taken from the corrections chained behind the pass
      std::vector<int32_t> chk;
      for (size_t si = 0; si < B.slots.size(); ++si) {
        const int32_t sl = B.slots[si];
        if (h->cl[sl].used && w.slot_epoch[sl] != B.epoch[si] && w.slot_upto[sl] >= w.batch_end) chk.push_back(sl);
      }
      TRY(win_check_slots(h, pos, chk, &h->res_bump[bi], false));
    }
    w.blog_used = 0;
    {  // room for the logged block columns of this batch's leave-side changes
      int64_t maxm = 0;
      for (const Cluster& c : h->cl)
        if (c.used) maxm = std::max(maxm, c.m);
      // (one share per batch buffer: the kernels queued behind the passes in flight read this batch's log later)
      TRY(ensure(h, &h->d_blog, &h->blog_elems, NBUF * (int64_t)(OWN_BATCH + 1) * ((maxm + (int64_t)OWN_BATCH * NP) * NP + NP)));
      w.blog_half = h->blog_elems / NBUF;
      w.blog_base = bi * w.blog_half;
    }
    std::fill(w.slot_alg.begin(), w.slot_alg.end(), 0);
    // re-scoring blocks written NBUF batches ago are no longer referenced
    ScoreRes& arena = h->res_bump[bi];
    if (h->ev_ts_end[bi]) CK(h, cudaStreamWaitEvent(h->main_stream, h->ev_ts_end[bi], 0));  // (k_T_leave on them: long done)
    arena.scr_used = 0;
    for (double* p : arena.retired) cudaFree(p);
    arena.retired.clear();
    const int64_t bcnt = B.cnt;
    for (int64_t q = 0; q < bcnt; ++q) {
      int32_t chosen = -1;
      TRY(win_step(h, a_draw ? a_draw[pos + q] : 0.0, u[pos + q], &chosen, &arena));
      if (chosen_ids) chosen_ids[pos + q] = chosen;
      *mm = std::min(*mm, ch.last_margin);
    }
    if (tmaint_on(h)) {
      if (!h->ev_ts_end[bi]) CK(h, cudaEventCreateWithFlags(&h->ev_ts_end[bi], cudaEventDisableTiming));
      CK(h, cudaEventRecord(h->ev_ts_end[bi], h->ts0));
    }
    pos += bcnt;
  }
  w.flight.clear();
  return MOGP_OK;
}

// True when order[] visits every patient at most once (the reference sweeps over a permutation).
bool order_is_duplicate_free(const Chain& ch, int64_t n_steps, const int64_t* order) {
  std::vector<char> seen(ch.N, 0);
  for (int64_t s = 0; s < n_steps; ++s) {
    const int64_t n = order[s];
    if (n < 0 || n >= ch.N) return false;
    if (seen[n]) return false;
    seen[n] = 1;
  }
  return true;
}

}  // namespace

extern "C" {

int mogp_host_choice(const double* score, int32_t n, double u, double* p_out, int32_t* index_out, double* margin_out) {
  if (!score || n <= 0) return MOGP_EINVAL;
  int idx = 0;
  int rc = host_choice(score, n, u, p_out, &idx, margin_out);
  if (index_out) *index_out = idx;
  return rc;
}

int mogp_chain_init(mogp_engine* h, const mogp_chain_config* cfg, int64_t n_patients, const int32_t* z, int32_t n_ids,
                    const double* theta_of_id) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!cfg || !z || n_ids <= 0 || !theta_of_id) FAIL(h, MOGP_EINVAL, "null argument");
  if (h->P <= 0 || n_patients != h->P) FAIL(h, MOGP_EINVAL, "chain needs the patient table (%lld patients set)", (long long)h->P);
  if (cfg->use_threshold && h->h_ythr.empty()) FAIL(h, MOGP_EINVAL, "threshold rule needs y_thr in mogp_set_patients");
  Chain& ch = h->chain;
  // release the clusters of a previous chain
  if (ch.active)
    for (int32_t s : ch.slot_of_id)
      if (s >= 0 && s < (int32_t)h->cl.size() && h->cl[s].used) cluster_free_i(h, s);
  ch = Chain();
  ch.cfg = *cfg;
  ch.N = n_patients;
  ch.z.assign(z, z + n_patients);
  ch.Nk.assign(n_ids, 0);
  ch.slot_of_id.assign(n_ids, -1);
  std::vector<std::vector<int64_t>> mem(n_ids);
  for (int64_t i = 0; i < n_patients; ++i) {
    if (z[i] < 0 || z[i] >= n_ids) FAIL(h, MOGP_EINVAL, "z[%lld] = %d out of range", (long long)i, z[i]);
    ch.Nk[z[i]] += 1;  // allocmodel.calc_suffstats: bincount (allocmodel.py:18-20)
    mem[z[i]].push_back(i);
  }
  for (int32_t k = 0; k < n_ids; ++k) {
    if (ch.Nk[k] == 0) continue;
    int32_t slot;
    TRY(cluster_create_i(h, &cfg->spec, theta_of_id + 4 * k, &slot));
    ch.slot_of_id[k] = slot;
    TRY(cluster_set_members_i(h, slot, (int64_t)mem[k].size(), mem[k].data()));
  }
  ch.active = true;
  return MOGP_OK;
}

int mogp_chain_step_begin(mogp_engine* h, int64_t n) {
  TRY(chain_check(h, n));
  return chain_begin(h, n);
}

int mogp_chain_step_score(mogp_engine* h, double a_draw, int32_t* n_active, int32_t* ids_out, double* score_out, int32_t cap_out) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!h->chain.active) FAIL(h, MOGP_ESTATE, "no chain attached");
  TRY(chain_score(h, a_draw));
  Chain& ch = h->chain;
  const int nact = (int)ch.act_ids.size();
  if (n_active) *n_active = nact;
  if (cap_out < nact && (ids_out || score_out)) FAIL(h, MOGP_EINVAL, "output capacity %d < %d active ids", cap_out, nact);
  for (int i = 0; i < nact; ++i) {
    if (ids_out) ids_out[i] = ch.act_ids[i];
    if (score_out) score_out[i] = ch.act_score[i];
  }
  return MOGP_OK;
}

int mogp_chain_step_commit(mogp_engine* h, int32_t chosen_id, double a_draw, int32_t* is_new) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!h->chain.active) FAIL(h, MOGP_ESTATE, "no chain attached");
  return chain_commit(h, chosen_id, a_draw, is_new);
}

int mogp_chain_step(mogp_engine* h, int64_t n, double a_draw, double u, int32_t* chosen_id, int32_t* is_new, int32_t* n_active,
                    int32_t* ids_out, double* p_out, int32_t cap_out) {
  TRY(chain_check(h, n));
  TRY(chain_begin(h, n));
  TRY(chain_score(h, a_draw));
  Chain& ch = h->chain;
  const int nact = (int)ch.act_ids.size();
  std::vector<double> p(nact);
  int idx = 0;
  int rc = host_choice(ch.act_score.data(), nact, u, p.data(), &idx, &ch.last_margin);
  if (rc != MOGP_OK) FAIL(h, rc, "probabilities contain NaN (patient %lld)", (long long)n);
  record_step_probs(h, ch.act_ids, p.data());
  if (n_active) *n_active = nact;
  if (ids_out || p_out) {
    if (cap_out < nact) FAIL(h, MOGP_EINVAL, "output capacity %d < %d active ids", cap_out, nact);
    for (int i = 0; i < nact; ++i) {
      if (ids_out) ids_out[i] = ch.act_ids[i];
      if (p_out) p_out[i] = p[i];
    }
  }
  const int32_t chosen = ch.act_ids[idx];
  if (chosen_id) *chosen_id = chosen;
  return chain_commit(h, chosen, a_draw, is_new);
}

int mogp_chain_sweep(mogp_engine* h, int64_t n_steps, const int64_t* order, const double* a_draw, const double* u,
                     int32_t* chosen_ids, double* min_margin) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!order || !u || n_steps < 0) FAIL(h, MOGP_EINVAL, "null argument");
  if (h->chain.active && h->chain.cfg.spec.mean_func && !a_draw) FAIL(h, MOGP_EINVAL, "mean_func needs a_draw");
  double mm = INFINITY;
  h->rec_off.clear();
  h->rec_ids.clear();
  h->rec_p.clear();
  h->chain.win.open = h->chain.win.in_step = false;
  const int max_cols = h->chain.active && h->chain.cfg.fast_updates && h->chain.cfg.alpha > 0.0 ? window_cols(h->chain) : 0;
  if (max_cols >= 2 && n_steps > 0 && order_is_duplicate_free(h->chain, n_steps, order)) {
    int rc = sweep_pipelined(h, n_steps, order, a_draw, u, chosen_ids, &mm, max_cols);
    // whatever happened, nothing of the pipeline stays in flight or cached
    join_side(h);
    for (int i = 0; i < NBUF; ++i) {
      cudaStreamSynchronize(h->ahead_ring[i]);
      cudaStreamSynchronize(h->bb[i].res.pre);
      cudaStreamSynchronize(h->tq[i]);
      cudaStreamSynchronize(h->ts1[i]);
    }
    cudaStreamSynchronize(h->ahead2);
    cudaStreamSynchronize(h->ts0);
    h->last_wide = nullptr;
    h->pass_absorbed = h->pass_launched;
    pool_unhold(h, true);
    h->chain.win.open = h->chain.win.in_step = false;
    if (rc != MOGP_OK) return rc;
    if (min_margin) *min_margin = mm;
    return MOGP_OK;
  }
  for (int64_t s = 0; s < n_steps; ++s) {
    int32_t chosen = -1;
    TRY(mogp_chain_step(h, order[s], a_draw ? a_draw[s] : 0.0, u[s], &chosen, nullptr, nullptr, nullptr, nullptr, 0));
    if (chosen_ids) chosen_ids[s] = chosen;
    mm = std::min(mm, h->chain.last_margin);
  }
  if (min_margin) *min_margin = mm;
  return MOGP_OK;
}

int mogp_chain_get_state(mogp_engine* h, int32_t* z, int64_t* Nk, int32_t* slot_of_id, int32_t cap_ids, int32_t* n_ids) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (!h->chain.active) FAIL(h, MOGP_ESTATE, "no chain attached");
  Chain& ch = h->chain;
  const int32_t k = (int32_t)ch.Nk.size();
  if (n_ids) *n_ids = k;
  if (z) memcpy(z, ch.z.data(), sizeof(int32_t) * ch.N);
  if (Nk || slot_of_id) {
    if (cap_ids < k) FAIL(h, MOGP_EINVAL, "capacity %d < %d cluster ids", cap_ids, k);
    for (int32_t i = 0; i < k; ++i) {
      if (Nk) Nk[i] = ch.Nk[i];
      if (slot_of_id) slot_of_id[i] = slot_handle(h, ch.slot_of_id[i]);
    }
  }
  return MOGP_OK;
}

double mogp_chain_last_margin(const mogp_engine* h) { return h ? h->chain.last_margin : NAN; }

int mogp_chain_record_probs(mogp_engine* h, int32_t on) {
  if (!h) return MOGP_EINVAL;
  h->record_probs = on != 0;
  if (!h->record_probs) {
    h->rec_off.clear();
    h->rec_ids.clear();
    h->rec_p.clear();
  }
  return MOGP_OK;
}

int mogp_chain_get_probs(mogp_engine* h, int64_t* n_steps, int64_t* total, int64_t* off, int32_t* ids, double* p,
                         int64_t cap_steps, int64_t cap_total) {
  if (!h) return MOGP_EINVAL;
  const int64_t ns = h->rec_off.empty() ? 0 : (int64_t)h->rec_off.size() - 1, tot = (int64_t)h->rec_ids.size();
  if (n_steps) *n_steps = ns;
  if (total) *total = tot;
  if (!off && !ids && !p) return MOGP_OK;
  if (cap_steps < ns || cap_total < tot) FAIL(h, MOGP_EINVAL, "capacity (%lld steps, %lld entries) < (%lld, %lld)",
                                              (long long)cap_steps, (long long)cap_total, (long long)ns, (long long)tot);
  if (off) {
    if (ns == 0) off[0] = 0;
    else memcpy(off, h->rec_off.data(), sizeof(int64_t) * (ns + 1));
  }
  if (ids && tot) memcpy(ids, h->rec_ids.data(), sizeof(int32_t) * tot);
  if (p && tot) memcpy(p, h->rec_p.data(), sizeof(double) * tot);
  return MOGP_OK;
}

}  // extern "C"
