// refit.inl — the hyper-parameter refit of many clusters, driven by ONE host thread (included by engine.cu).
//
// Reference: GPobs.update_suffstats -> model.optimize() (mogp/obsmodel.py:87-94): paramz runs SciPy's L-BFGS-B on
// the objective -(LML + log prior + log Jacobian), warm-started at the cluster's current hyper-parameters, then
// evaluates f(x_opt) once more and sets the optimiser array to x_opt.  One refit is a chain of ~10-25 dependent
// objective evaluations, each a captured graph of ~130 kernels whose diagonal-tile steps run on a single SM: one
// cluster cannot fill the GPU, thirty can.  The refits of different clusters are independent, so every cluster gets
// a LANE (a light worker engine: stream + factorisation workspace + graph) and its own L-BFGS-B state machine
// (lbfgsb.hpp, same evaluation points as SciPy); the caller's thread polls the lanes' events, feeds (f, g) of a
// finished evaluation to that cluster's state machine and queues its next evaluation at once.  No host thread per
// cluster, no Python between evaluations; every cluster sees exactly the evaluations of its own sequential optimize().
struct RefitJob {
  int32_t slot = -1;
  Lbfgsb opt;
  int nfree = 0;
  int evals = 0, fails = 0;
  double last_g[4] = {0, 0, 0, 0};
  bool have_g = false;
  int phase = 0;          // 0: optimiser running, 1: f(x_opt) re-evaluation requested, 2: done
  bool launched = false;  // an evaluation is in flight on the lane
  int lane = -1;
  int64_t seq = 0;        // launch order (the oldest launch is waited for when nothing has finished)
  int rc = MOGP_OK;
  bool collapsed = false; // the optimiser's evaluations run on the lane's shadow of the cluster (repeated inputs merged)
  double t_start = 0.0, t_end = 0.0;  // host seconds since the refit began (MOGP_REFIT_TRACE)
  bool ws = false;        // the optimiser's evaluations run in weight space (wspace.cuh) on the lane ...
  double span = 0.0;      // ... while (hi - lo) / lengthscale <= WS_MAX_SPAN (hi - lo = this); else on the cluster itself
  int target = 0;         // what the evaluation in flight runs on: 0 the lane's shadow (merged inputs), 1 the cluster, 2 weight space
  int ws_evals = 0;
};

// Weight-space evaluations (wspace.cuh): eligible clusters (RBF kernel, m >= 4 r) get the theta-independent part - the
// Lagrange basis values of their inputs, S = C^T C, C^T y, C^T x and three scalar sums - once per refit on the lane's stream;
// the lane's cluster 0 serves as the holder of theta / priors for objective_begin / objective_end.
static bool ws_enabled() {  // on by default; MOGP_WS_REFIT=0: every evaluation the dense way (read per call: tests toggle it)
  const char* e = getenv("MOGP_WS_REFIT");
  return e ? e[0] != '0' : true;
}
static int ws_setup(mogp_engine* lane, const Cluster& c, bool* made, double* span) {
  *made = false;
  if (!ws_enabled() || c.spec.kernel != MOGP_KERNEL_RBF || c.m < 4 * WS_R || (int64_t)c.hx.size() != c.m || !c.M) return MOGP_OK;
  mogp_engine* h = lane;
  double lo = c.hx[0], hi = c.hx[0];
  for (double v : c.hx) {
    lo = std::min(lo, v);
    hi = std::max(hi, v);
  }
  if (!(hi > lo)) return MOGP_OK;
  if (lane->cl.empty()) {
    int32_t slot = -1;
    TRY(cluster_create_i(lane, &c.spec, c.theta, &slot));
  }
  Cluster& sh = lane->cl[0];
  sh.spec = c.spec;
  memcpy(sh.theta, c.theta, sizeof(double) * 4);
  sh.theta_dirty = true;
  sh.m = c.m;
  sh.valid = sh.grad_valid = false;
  sh.stat_on_device = false;
  sh.wt = nullptr;
  sh.corr_n1 = sh.corr_logn = sh.corr_ss = 0.0;
  TRY(ensure(lane, &lane->d_ws, &lane->ws_elems, WS_WORK_ELEMS));
  TRY(ensure(lane, &lane->d_wsC, &lane->wsC_elems, c.m * (int64_t)WS_R));
  TRY(ensure_pin(lane, 256));
  WsNodes nd;
  ws_make_nodes(lo, hi, &nd);
  WsWork w{lane->d_ws};
  k_ws_basis<<<(unsigned)((c.m + 127) / 128), 128, 0, lane->stream>>>(nd, c.x(), (int)c.m, lane->d_wsC, w);
  LAUNCHED(lane);
  k_ws_gram<<<(WS_GRAM_ENTRIES + 127) / 128, 128, 0, lane->stream>>>(lane->d_wsC, c.x(), c.y(), (int)c.m, w);
  LAUNCHED(lane);
  CK(h, cudaGetLastError());
  *made = true;
  *span = hi - lo;
  return MOGP_OK;
}

// Repeated inputs.  Every patient contributes the same onset anchor (x = 0), so ~1/8 of a cluster's points share one input;
// n observations at the same input are ONE observation of their mean with noise s / n, times a factor that does not involve
// the GP: prod_i N(y_i | f, s) = N(ybar | f, s / n) (2 pi s)^-(n-1)/2 n^-1/2 exp(-SS / 2s), s = noise + 1e-8.  The optimiser's
// evaluations - full factorisations, O(m^3) - therefore run on a SHADOW of the cluster with every group of identical inputs
// merged (m' ~ 0.87 m: a third fewer flops), owned by the lane: weights in k_build / k_grad, the closed-form factor and its
// noise derivative added on the host (objective_end).  Same objective and gradient as the full cluster to rounding (~1e-12);
// the last evaluation, f(x_opt), runs on the cluster itself and leaves it with its factor at the optimum.
static int refit_make_shadow(mogp_engine* lane, const Cluster& c, bool* made) {
  *made = false;
  static const bool on = !(getenv("MOGP_COLLAPSE") && getenv("MOGP_COLLAPSE")[0] == '0');
  const int64_t m = c.m;
  if (!on || m < 2 * TB || (int64_t)c.hx.size() != m || (int64_t)c.hy.size() != m) return MOGP_OK;
  std::vector<int64_t> idx(m);
  for (int64_t i = 0; i < m; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return c.hx[a] < c.hx[b]; });
  std::vector<double> xg, yg, wg;
  double n1 = 0.0, logn = 0.0, ss = 0.0;
  for (int64_t i = 0; i < m;) {
    int64_t j = i;
    double sum = 0.0;
    while (j < m && c.hx[idx[j]] == c.hx[idx[i]]) sum += c.hy[idx[j++]];
    const double n = (double)(j - i), mean = sum / n;
    double dev = 0.0;
    for (int64_t q = i; q < j; ++q) dev += (c.hy[idx[q]] - mean) * (c.hy[idx[q]] - mean);
    xg.push_back(c.hx[idx[i]]);
    yg.push_back(mean);
    wg.push_back(n);
    n1 += n - 1.0;
    logn += 0.5 * log(n);
    ss += dev;
    i = j;
  }
  const int64_t m2 = (int64_t)xg.size();
  if (m2 * 20 > m * 19) return MOGP_OK;  // fewer than 5 % of the points merge: not worth a second factor at the end
  mogp_engine* h = lane;
  if (lane->cl.empty()) {
    int32_t slot = -1;
    TRY(cluster_create_i(lane, &c.spec, c.theta, &slot));
  }
  Cluster& sh = lane->cl[0];
  sh.spec = c.spec;
  memcpy(sh.theta, c.theta, sizeof(double) * 4);
  sh.theta_dirty = true;
  TRY(upload_data(lane, sh, m2, xg.data(), yg.data()));
  TRY(ensure(lane, &lane->d_wt, &lane->wt_elems, m2));
  char* up = nullptr;
  TRY(stage_alloc(lane, sizeof(double) * (size_t)m2, &up));
  memcpy(up, wg.data(), sizeof(double) * (size_t)m2);
  CK(h, cudaMemcpyAsync(lane->d_wt, up, sizeof(double) * (size_t)m2, cudaMemcpyHostToDevice, lane->stream));
  lane->bytes_h2d += (int64_t)sizeof(double) * m2;
  sh.wt = lane->d_wt;
  sh.corr_n1 = n1;
  sh.corr_logn = logn;
  sh.corr_ss = ss;
  *made = true;
  return MOGP_OK;
}

static int refit_lanes(mogp_engine* h, int want) {
  while ((int)h->lanes.size() < want) {
    mogp_engine* l = nullptr;
    int rc = engine_create_impl(h->device, true, &l);
    if (rc != MOGP_OK) FAIL(h, rc, "refit lane: %s", g_create_error.c_str());
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) {
      mogp_engine_destroy(l);
      FAIL(h, MOGP_ECUDA, "refit lane: event");
    }
    h->lanes.push_back(l);
    h->lane_ev.push_back(e);
  }
  for (mogp_engine* l : h->lanes) l->blocking_sync = h->blocking_sync;
  return MOGP_OK;
}

// Run job j forward until it has an evaluation in flight (returns with j.launched) or is done.
static int refit_advance(mogp_engine* h, RefitJob& j, int64_t* seq) {
  mogp_engine* lane = h->lanes[j.lane];
  for (;;) {
    if (j.phase == 2) return MOGP_OK;
    // the optimiser's evaluations in weight space or on the lane's collapsed shadow, the closing f(x_opt) on the cluster itself
    if (!j.launched) {
      j.target = j.phase != 0 ? 1 : j.ws ? 2 : j.collapsed ? 0 : 1;
      if (j.target == 2) {  // a length scale too short for the r nodes: this point the dense way, on the cluster
        Gamma pr[4];
        bool fixed[4];
        priors_of(h->cl[j.slot].spec, pr, fixed);
        int q = 0;
        double ls = h->cl[j.slot].theta[2];
        for (int i = 0; i < 4; ++i)
          if (!fixed[i]) {
            if (i == 2) ls = logexp_f(j.opt.x[q]);
            ++q;
          }
        if (!(j.span <= WS_MAX_SPAN * ls)) j.target = 1;
      }
    }
    Cluster& c = j.target == 1 ? h->cl[j.slot] : lane->cl[0];
    double f = 0.0, g[4] = {0, 0, 0, 0};
    int rc;
    if (!j.launched) {  // queue the evaluation at the optimiser's current point
      bool launched = false;
      // (the closing f(x_opt) needs the factor and the objective only: its graph leaves the gradient kernels out)
      rc = objective_begin(lane, c, j.opt.x, j.nfree, j.target != 2, &launched, j.phase == 0);
      j.evals += 1;
      h->refit_evals += 1;
      if (rc == MOGP_OK && j.target == 2 && !c.valid) {  // (valid: the point of the last evaluation again - cached)
        WsTheta th{c.theta[0], c.theta[1], c.theta[2], c.theta[3], c.spec.mean_func ? 1 : 0};
        WsWork w{lane->d_ws};
        k_ws_eval<<<1, 256, 0, lane->stream>>>(w, th);
        lane->launches += 1;
        if (cudaMemcpyAsync(lane->pin, w.out(), sizeof(double) * 8, cudaMemcpyDeviceToHost, lane->stream) != cudaSuccess)
          FAIL(h, MOGP_ECUDA, "refit: read-back of a weight-space evaluation");
        lane->bytes_d2h += 64;
        launched = true;
        j.ws_evals += 1;
      }
      if (rc == MOGP_OK && launched) {
        if (cudaEventRecord(h->lane_ev[j.lane], lane->stream) != cudaSuccess) FAIL(h, MOGP_ECUDA, "refit: event record");
        j.launched = true;
        j.seq = (*seq)++;
        h->refit_launched += 1;
        return MOGP_OK;
      }
      if (rc == MOGP_OK) rc = objective_end(lane, c, false, &f, j.phase == 0 ? g : nullptr);  // cached (same theta) or eager
    } else if (j.target == 2) {  // a weight-space evaluation has arrived
      j.launched = false;
      const double* o = reinterpret_cast<const double*>(lane->pin);
      if (o[5] != 0.0 || !isfinite(o[0])) {
        c.valid = c.grad_valid = false;
        lane->err = "not positive definite (weight-space evaluation)";
        rc = MOGP_ENOTPD;
      } else {
        c.lml = o[0];
        memcpy(c.grad, o + 1, sizeof(double) * 4);
        c.valid = c.grad_valid = true;
        c.stat_on_device = false;
        c.z_dirty = c.lml_dirty = false;
        rc = objective_end(lane, c, false, &f, g);
      }
    } else {  // the caller has seen the lane's event complete
      j.launched = false;
      lane->up_off = 0;  // the lane's stream is idle: its upload staging can be reused from the start
      rc = objective_end(lane, c, true, &f, j.phase == 0 ? g : nullptr);
    }
    if (rc == MOGP_ENOTPD) {
      // paramz Model._objective_grads: a failed evaluation returns a huge objective and the clipped last gradient,
      // up to 10 times in a row; then the LinAlgError goes up
      if (j.fails >= 10) {
        h->err = lane->err;
        return j.rc = rc;
      }
      j.fails += 1;
      f = 1.7976931348623157e308;
      for (int i = 0; i < j.nfree; ++i) g[i] = j.have_g ? std::min(1e10, std::max(-1e10, j.last_g[i])) : 0.0;
    } else if (rc != MOGP_OK) {
      h->err = lane->err;
      return j.rc = rc;
    } else {
      j.fails = 0;
      memcpy(j.last_g, g, sizeof(double) * 4);
      j.have_g = true;
    }
    if (j.phase == 1) {  // that was f(x_opt); the final `optimizer_array = x_opt` finds the cluster at theta(x_opt)
      j.phase = 2;
      return MOGP_OK;
    }
    j.opt.f = f;
    memcpy(j.opt.g, g, sizeof(double) * j.nfree);
    Lbfgsb::Task task;
    do task = j.opt.step();
    while (task == Lbfgsb::T_NEW_X);
    if (task != Lbfgsb::T_FG) j.phase = 1;  // finished: opt.x = x_opt, evaluated once more (usually the cached point)
  }
}

/* see include/mogp_b200.h */
int mogp_refit_clusters(mogp_engine* h, int32_t n_slots, const int32_t* slots, int32_t max_concurrent, int32_t* n_evals_out) {
  if (!h) return MOGP_EINVAL;
  cudaSetDevice(h->device);  // entry points may be called from any host thread
  if (n_slots <= 0 || !slots) FAIL(h, MOGP_EINVAL, "no clusters to refit");
  std::vector<RefitJob> jobs(n_slots);
  for (int i = 0; i < n_slots; ++i) {
    int32_t idx = slots[i];
    TRY(slot_in(h, &idx));
    for (int q = 0; q < i; ++q)
      if (jobs[q].slot == idx) FAIL(h, MOGP_EINVAL, "cluster %d listed twice", slots[i]);
    jobs[i].slot = idx;
    if (h->cl[idx].m <= 0) FAIL(h, MOGP_ESTATE, "cluster has no data");
  }
  // the lanes read the owner's clusters: everything queued on the owner's streams must have finished
  join_side(h);
  CK(h, sync_stream(h));
  static const int env_lanes = [] {
    const char* e = getenv("MOGP_REFIT_LANES");
    return e ? atoi(e) : 0;
  }();
  int n_lanes = max_concurrent > 0 ? max_concurrent : (env_lanes > 0 ? env_lanes : 32);
  n_lanes = std::max(1, std::min(n_lanes, (int)n_slots));
  TRY(refit_lanes(h, n_lanes));
  // largest clusters first: the longest chains of evaluations start earliest
  std::vector<int> order(n_slots);
  for (int i = 0; i < n_slots; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->cl[jobs[a].slot].m > h->cl[jobs[b].slot].m; });
  std::vector<int> lane_job(n_lanes, -1);
  static const bool trace = getenv("MOGP_REFIT_TRACE") && getenv("MOGP_REFIT_TRACE")[0] == '1';
  const auto t_zero = std::chrono::steady_clock::now();
  auto now_s = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_zero).count(); };
  int next = 0, done = 0, rc_all = MOGP_OK;
  int64_t seq = 0;
  auto start_job = [&](int lane) -> int {
    while (next < n_slots) {
      RefitJob& j = jobs[order[next++]];
      Cluster& c = h->cl[j.slot];
      Gamma pr[4];
      bool fixed[4];
      priors_of(c.spec, pr, fixed);
      double t0[4];
      j.nfree = 0;
      for (int i = 0; i < 4; ++i)
        if (!fixed[i]) t0[j.nfree++] = logexp_finv(c.theta[i]);
      if (j.nfree == 0) {  // nothing to optimise (paramz returns at once)
        j.phase = 2;
        ++done;
        continue;
      }
      j.opt.init(j.nfree, t0);
      j.opt.step();  // -> T_FG at t0
      j.lane = lane;
      j.t_start = now_s();
      {
        int rcs = ws_setup(h->lanes[lane], c, &j.ws, &j.span);
        if (rcs == MOGP_OK && !j.ws) rcs = refit_make_shadow(h->lanes[lane], c, &j.collapsed);
        if (rcs != MOGP_OK) {
          h->err = h->lanes[lane]->err;
          return rcs;
        }
      }
      lane_job[lane] = (int)(&j - jobs.data());
      int rc = refit_advance(h, j, &seq);
      if (rc != MOGP_OK) return rc;
      if (j.phase == 2) {  // finished without ever waiting (every evaluation cached or eager)
        lane_job[lane] = -1;
        ++done;
        continue;
      }
      return MOGP_OK;
    }
    lane_job[lane] = -1;
    return MOGP_OK;
  };
  for (int l = 0; l < n_lanes && rc_all == MOGP_OK; ++l) rc_all = start_job(l);
  while (rc_all == MOGP_OK && done < n_slots) {
    bool progressed = false;
    int oldest = -1;
    for (int l = 0; l < n_lanes && rc_all == MOGP_OK; ++l) {
      const int ji = lane_job[l];
      if (ji < 0) continue;
      RefitJob& j = jobs[ji];
      const cudaError_t q = cudaEventQuery(h->lane_ev[l]);
      if (q == cudaErrorNotReady) {
        if (oldest < 0 || j.seq < jobs[lane_job[oldest]].seq) oldest = l;
        continue;
      }
      if (q != cudaSuccess) FAIL(h, MOGP_ECUDA, "refit: %s", cudaGetErrorString(q));
      progressed = true;
      rc_all = refit_advance(h, j, &seq);
      if (rc_all == MOGP_OK && j.phase == 2) {
        ++done;
        j.t_end = now_s();
        rc_all = start_job(l);
      }
    }
    if (rc_all != MOGP_OK) break;
    if (!progressed && oldest >= 0 && h->blocking_sync) cudaEventSynchronize(h->lane_ev[oldest]);  // sleep, not spin
  }
  // whatever happened, nothing stays in flight on the lanes
  for (int l = 0; l < n_lanes; ++l) {
    cudaStreamSynchronize(h->lanes[l]->stream);
    h->lanes[l]->up_off = 0;
    h->launches += h->lanes[l]->launches;
    h->lanes[l]->launches = 0;
    h->bytes_h2d += h->lanes[l]->bytes_h2d;
    h->bytes_d2h += h->lanes[l]->bytes_d2h;
    h->lanes[l]->bytes_h2d = h->lanes[l]->bytes_d2h = 0;
  }
  if (trace) {  // one line per cluster: size, evaluations, when its chain of evaluations started and ended
    fprintf(stderr, "refit trace: %d clusters, %d lanes, %.1f ms\n", (int)n_slots, n_lanes, now_s() * 1e3);
    for (int i : order)
      fprintf(stderr, "  m %5lld evals %3d  %7.1f .. %7.1f ms%s\n", (long long)h->cl[jobs[i].slot].m, jobs[i].evals,
              jobs[i].t_start * 1e3, jobs[i].t_end * 1e3,
              jobs[i].ws ? "  (weight space)" : jobs[i].collapsed ? "  (merged)" : "");
  }
  if (n_evals_out)
    for (int i = 0; i < n_slots; ++i) n_evals_out[i] = jobs[i].evals;
  if (rc_all != MOGP_OK) {  // a cluster whose evaluation is abandoned mid-flight has no valid factor
    for (RefitJob& j : jobs)
      if (j.launched) h->cl[j.slot].valid = h->cl[j.slot].grad_valid = false;
  }
  return rc_all;
}

int mogp_refit_counters(mogp_engine* h, int64_t* evaluations, int64_t* launched) {
  if (!h) return MOGP_EINVAL;
  if (evaluations) *evaluations = h->refit_evals;
  if (launched) *launched = h->refit_launched;
  return MOGP_OK;
}
