#!/usr/bin/env python
"""bench.py — Gibbs sweeps/sec at 10k patients (BASELINE.json `metric`, configs[2]) on N B200s.

One "step" = one collapsed Gibbs sweep of MoGP_constrained over all patients (mogp_constrained.py:231-284)
followed by the hyper-parameter refit of every active cluster (refit='sweep', the schedule BASELINE.json names
for the 10k-patient configuration).  N > 1 runs N independent chains (seed = rank), one per GPU — the
alpha/seed-sweep sharding of SURVEY.md §8(e); a single chain is strictly sequential ("replicas only").

  value     whole-job sweeps/s, patient table resident in HBM (C-ABI chain calls + host L-BFGS-B driver)
  e2e       the same through the public API `MoGP_constrained.sweep()` with the patient table re-sent from
            host memory every sweep and the results (z, Nk, log-likelihoods) read back
  roofline  the scoring product T = M Kx (k_score_gemm), timed with CUDA events on the engine's stream
  cpu_baseline / --impl reference   the CPU oracle (NumPy/SciPy restatement of the reference + GPy), on a
            bounded sample of the same sweep, all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patients", type=int, default=10000)
    ap.add_argument("--clusters", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=12, help="patient-steps of the bounded CPU sample")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "predict", "chains"],
                    help="sweep: the headline (BASELINE configs[2]); predict: batch prediction, patients sharded "
                         "(configs[4]); chains: independent 2k-patient chains dealt to the GPUs (configs[3])")
    ap.add_argument("--query-patients", type=int, default=1000000)
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--chain-patients", type=int, default=2000)
    ap.add_argument("--chain-refit", default="sweep", choices=["sweep", "move"],
                    help="chains workload: 'move' = the reference's schedule (L-BFGS-B after every move)")
    return ap.parse_args()


def workload(args, seed):
    from mogp_b200.synth import alsfrs_like
    # patients drawn from `clusters` latent progression classes; the chain starts from those classes (the
    # reference starts from k-means; either way BASELINE's configuration is "~30 active clusters")
    X, Y, z0 = alsfrs_like(args.patients, seed=seed, n_classes=args.clusters, return_labels=True)
    z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
    kw = dict(alpha=args.clusters / np.log10(args.patients), num_iter=2, rand_seed=seed, normalize=True, kernel='rbf',
              mean_func=True, threshold=0.5, signal_variance=1., signal_variance_fix=True, noise_variance=0.5,
              noise_variance_fix=False, onset_anchor=True, refit='sweep', init_z=z0)
    return X, Y, kw


def config(args):
    return {"workload": "MoGP_constrained Gibbs sweep, {} synthetic ALSFRS-R-like patients (3-10 visits + onset anchor), "
                        "{} initial clusters, rbf + neg_linear mean, threshold 0.5, hyper-parameter refit of every active "
                        "cluster each sweep (BASELINE configs[2])".format(args.patients, args.clusters),
            "patients": args.patients, "init_clusters": args.clusters, "refit": "sweep",
            "cache": "inputs larger than L2: the resident cluster factors total > 700 MB vs 126 MB of L2",
            "parallelism": "independent chains, one per GPU (seed = rank)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def oracle_sample(args, n_steps, refit_clusters=1, seed=0):
    """Time the CPU oracle on a bounded sample of the sweep: n_steps consecutive patient-steps from the same
    initial state + the L-BFGS-B refit of `refit_clusters` clusters; extrapolate to sweeps/s."""
    import oracle
    X, Y, kw = workload(args, seed)
    t0 = time.perf_counter()
    orc = oracle.MoGP_constrained(X=X, Y=Y, **kw)
    orc.initialize_sampler(orc.num_init_clusters, True)
    orc.allocmodel.calc_suffstats(orc.z)
    t_init = time.perf_counter() - t0
    return orc, t_init


def oracle_step_times(orc, n_steps, refit_clusters):
    perm = np.random.permutation(np.arange(orc.allocmodel.N))[:n_steps]
    t0 = time.perf_counter()
    for n in perm:
        orc.gibbs_step(n)
    t_step = (time.perf_counter() - t0) / max(n_steps, 1)
    act = np.where(orc.allocmodel.Nk > 0)[0]
    t0 = time.perf_counter()
    evals = 0
    for k in act[:refit_clusters]:
        m = orc.obsmodel[k].model
        e0 = m.n_evals
        m.optimize()
        evals += m.n_evals - e0
    t_refit = (time.perf_counter() - t0) / max(refit_clusters, 1)
    return t_step, t_refit, len(act), evals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count()
    orc, t_init = oracle_sample(args, 0)
    N = args.patients
    per_step = max(2, args.cpu_steps // 3)
    times = []
    for it in range(args.warmup + args.steps):
        ts, tr, nact, evals = oracle_step_times(orc, per_step, 1)
        if it >= args.warmup:
            times.append(N * ts + nact * tr)
    sweep_s = float(np.mean(times))
    value = 1.0 / sweep_s
    sample = ("each step = {} consecutive patient-steps + the L-BFGS-B refit of 1 cluster, from the same initial state; "
              "sweep time = N*mean(step) + active*refit".format(per_step))
    line = {"impl": "reference", "metric": "Gibbs sweeps/sec @10k patients", "value": value, "unit": "sweeps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sweep_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args),
            "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": ncores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mogp_b200 import _capi, MoGP_constrained
    from mogp_b200.utils import pack_patients

    eng = _capi.Engine(local)
    X, Y, kw = workload(args, seed=rank)
    mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
    mix.initialize_sampler(args.clusters, True)
    off, px, py = pack_patients(mix.X, mix.Y)
    N = args.patients

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    split = [0.0, 0.0]   # host wall time in the patient steps / in the refits

    def device_sweep():
        """value path: C-ABI chain sweep over the resident patient table + refit of every active cluster."""
        perm, a, u = mix._draws_for_sweep(N)
        t0 = time.perf_counter()
        eng.chain_sweep(perm, a, u)
        mix._sync_state()
        t1 = time.perf_counter()
        mix.refit_all()
        split[0] += t1 - t0
        split[1] += time.perf_counter() - t1

    def e2e_sweep():
        """public API with host buffers: patient table host->device, sweep, results device->host."""
        eng.refresh_patients(off, px, py)
        mix.sweep()
        lobs = sum(mix.obsmodel[k].calc_ll() for k in mix._active())
        return float(lobs) + float(mix.allocmodel.calc_ll())

    for _ in range(args.warmup):
        device_sweep()
    moved_probe = mix.z.copy()

    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def all_launches():   # the chain's engine + the worker engines of the concurrent refit
        return eng.launch_count() + sum(w.launch_count() for w in (mix._workers or []))

    l0 = all_launches()
    split[0] = split[1] = 0.0
    eng.profile(True)
    eng.profile_read(reset=True)
    cnt0 = eng.chain_counters()
    eng.timer_start()
    for _ in range(args.steps):
        device_sweep()
    ms = eng.timer_stop()
    cnt1 = eng.chain_counters()
    counters = {k: cnt1[k] - cnt0[k] for k in cnt1}
    prof = eng.profile_read(reset=True)
    eng.profile(False)
    launches = all_launches() - l0
    split_timed = (split[0] / args.steps, split[1] / args.steps)
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    moved = int(np.sum(mix.z != moved_probe))

    # e2e
    barrier()
    b0 = eng.transfer_bytes()
    eng.timer_start()
    for _ in range(args.steps):
        ll = e2e_sweep()
    ms_e2e = eng.timer_stop()
    b1 = eng.transfer_bytes()
    barrier()
    t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # the dominant kernel on its own (no concurrent work): one batch of the look-ahead pipeline = the first patients of a
    # permutation up to the pipeline's column budget, against every active cluster of the final state
    iso = None
    if rank == 0:
        cols = int(os.environ.get("MOGP_LOOKAHEAD", "48"))
        take, tot = [], 0
        for n in np.random.RandomState(1).permutation(N):
            ni = int(off[n + 1] - off[n])
            if tot + ni > cols:
                break
            take.append(n); tot += ni
        qoff = np.zeros(len(take) + 1, dtype=np.int64)
        qx, qy = [], []
        for i, n in enumerate(take):
            qx.append(px[off[n]:off[n + 1]]); qy.append(py[off[n]:off[n + 1]])
            qoff[i + 1] = qoff[i] + (off[n + 1] - off[n])
        qx, qy = np.concatenate(qx), np.concatenate(qy)
        _z, _Nk, slots_now = eng.chain_state(N)
        act = np.array([s for s in slots_now if s >= 0], dtype=np.int32)
        eng.profile(True); eng.profile_read(reset=True)
        for rep in range(10):
            eng.score_pairs(qoff, qx, qy, act)
            if rep == 1:
                eng.profile_read(reset=True)
        pi = eng.profile_read(reset=True); eng.profile(False)
        iso = {"columns": int(tot), "patients": len(take), "avg_launch_us": pi["ms"] / max(pi["launches"], 1) * 1e3,
               "fp64_tflops": pi["alg_flops"] / (pi["ms"] * 1e-3) / 1e12, "hbm_gbs": pi["alg_bytes"] / (pi["ms"] * 1e-3) / 1e9}

    # gather the chains' results (the only collective of this mode)
    occupancy = [int(v) for v in mix.allocmodel.Nk[mix.allocmodel.Nk > 0]]
    if world > 1:
        zt = torch.from_numpy(mix.z.astype(np.int32)).cuda()
        out = [torch.empty_like(zt) for _ in range(world)]
        dist.all_gather(out, zt)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else \
            (6650.0, "fallback (B200_PROFILING.md)")
        # FP64 tensor peak: MEASURED_PEAKS.json holds only HBM and bf16; tcgen05 has no FP64, the FP64 tensor instruction
        # is DMMA.8x8x4, measured on this pool's B200 by scripts/peaks/fp64_peak.cu (profiles/r1_fp64_peaks.txt)
        fp64_peak, fp64_src = 37.2, "measured DMMA.8x8x4 peak (profiles/r1_fp64_peaks.txt)"
        sec = prof["ms"] * 1e-3
        gbs = prof["alg_bytes"] / sec / 1e9 if sec > 0 else None
        tfs = prof["alg_flops"] / sec / 1e12 if sec > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "score_gemm_traffic.json"))).get("traffic_bytes_per_launch")
        except Exception:
            pass
        # one pass over the factors serves several patients (look-ahead window): whichever roof is closer binds
        hbm_frac = gbs / hbm_peak if gbs else None
        fp_frac = tfs / fp64_peak if tfs else None
        tensor_bound = fp_frac is not None and hbm_frac is not None and fp_frac >= hbm_frac
        if tensor_bound:
            roof = {"bound": "tensor", "achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp_frac,
                    "peak_source": fp64_src}
        else:
            roof = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac,
                    "peak_source": hbm_src}
        roof.update({"kernel": "k_score_gemm (T = M Kx, FP64 DMMA)", "traffic": traffic, "launches": prof["launches"],
                     "avg_launch_us": prof["ms"] / max(prof["launches"], 1) * 1e3,
                     "alg_bytes_per_launch": prof["alg_bytes"] / max(prof["launches"], 1),
                     "alg_flops_per_launch": prof["alg_flops"] / max(prof["launches"], 1),
                     "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac, "peak_source": hbm_src},
                     "fp64": {"achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp_frac, "peak_source": fp64_src},
                     "pairs_per_launch": counters["pairs"] / max(prof["launches"], 1),
                     "share_of_step": prof["ms"] / ms,
                     "note": "achieved = whole passes (every active cluster) timed live in the sweeps, where they share the SMs "
                             "with the moves of the previous batch; 'isolated' = the same kernel alone; 'rescoring' = the short "
                             "passes against the one or two clusters a move changed (latency bound, accounted apart)",
                     "isolated": iso,
                     "rescoring": {"launches": counters["rescoring_launches"],
                                   "avg_launch_us": counters["rescoring_ms"] / max(counters["rescoring_launches"], 1) * 1e3,
                                   "share_of_step": counters["rescoring_ms"] / ms}})
        if iso:
            roof["isolated"]["frac_fp64"] = iso["fp64_tflops"] / fp64_peak
            roof["isolated"]["frac_hbm"] = iso["hbm_gbs"] / hbm_peak
        line = {"metric": "Gibbs sweeps/sec @10k patients", "value": world * args.steps / (ms_max * 1e-3), "unit": "sweeps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config(args),
                "e2e": {"value": world * args.steps / (ms_e2e_max * 1e-3), "unit": "sweeps/s",
                        "h2d_bytes_per_step": (b1[0] - b0[0]) // args.steps, "d2h_bytes_per_step": (b1[1] - b0[1]) // args.steps},
                "gpu_launches": int(launches),
                "roofline": roof,
                "clocks": clocks,
                "chain": {"active_clusters": len(occupancy), "patients_moved_in_timed_sweeps": moved,
                          "pair_logliks_per_s": world * args.steps * N * len(occupancy) / (ms_max * 1e-3),
                          "refit_evals": int(mix.n_refit_evals), "total_loglik": ll,
                          "steps_s_per_sweep": split_timed[0], "refit_s_per_sweep": split_timed[1],
                          "lookahead": {k: counters[k] for k in ("window_passes", "rescoring_passes", "window_patients",
                                                                 "open_s", "rescore_s", "commit_s", "rank_corrections", "correct_s", "replayed")}}}
        if world == 1 and not args.no_cpu_baseline:
            orc, _t_init = oracle_sample(args, 0)
            ts, tr, nact, evals = oracle_step_times(orc, args.cpu_steps, 1)
            sweep_s = N * ts + nact * tr
            line["cpu_baseline"] = {"value": 1.0 / sweep_s, "unit": "sweeps/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "{} consecutive patient-steps ({:.3f} s each) + L-BFGS-B refit of 1 cluster ({:.2f} s, "
                                              "{} evaluations) from the same initial state; sweep = N*step + active*refit"
                                              .format(args.cpu_steps, ts, tr, evals)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
def run_predict(args):
    """configs[4]: P synthetic patients scored against a random-init 40-cluster model, patients sharded across the
    ranks (no collective on the data path), one all_gather of the probability matrix at the end."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mogp_b200 import _capi, MoGP_constrained, parallel
    from mogp_b200.synth import alsfrs_like, alsfrs_like_fast
    from mogp_b200.utils import pack_patients
    K = 40
    eng = _capi.Engine(local)
    Xt, Yt, z0 = alsfrs_like(3000, seed=0, n_classes=K, return_labels=True)
    z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
    mix = MoGP_constrained(X=Xt, Y=Yt, alpha=K / np.log10(3000), rand_seed=0, normalize=True, noise_variance=0.5,
                           init_z=z0, engine=eng)
    mix.initialize_sampler(K, True)
    rs = np.random.RandomState(1)
    _z, Nk, slots = eng.chain_state(3000)
    for k in np.where(Nk > 0)[0]:          # theta_k around the priors' means
        eng.cluster_set_theta(int(slots[k]), np.array([rs.gamma(2.22, 1 / 3.33), 1.0, rs.gamma(1.78, 1 / 0.444) + 0.3,
                                                       rs.gamma(9., 1 / 12.)]))
    P = args.query_patients
    Xq, Yq = alsfrs_like_fast(P, seed=2, n_classes=K)
    a = np.random.RandomState(3).randn(P)
    lo, hi = parallel.shard_bounds(P, world, rank)

    def barrier():
        eng.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nq = hi - lo
    for _ in range(max(1, args.warmup)):
        mix.predict_batch(Xq[lo:lo + min(nq, 20000)], Yq[lo:lo + min(nq, 20000)], a[lo:lo + min(nq, 20000)])
    barrier()
    l0 = eng.launch_count(); b0 = eng.transfer_bytes()
    eng.profile(True); eng.profile_read(reset=True)
    eng.timer_start()
    for _ in range(args.steps):
        probs = mix.predict_batch(Xq[lo:hi], Yq[lo:hi], a[lo:hi])
    ms = eng.timer_stop()
    prof = eng.profile_read(reset=True); eng.profile(False)
    b1 = eng.transfer_bytes()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    full = parallel.all_gather_rows(probs, P) if world > 1 else probs
    barrier()
    if rank == 0:
        n_act = int((Nk > 0).sum())
        pairs = float(P) * n_act * args.steps
        peak = 37.2   # measured DMMA.8x8x4 peak on this pool's B200 (profiles/r1_fp64_peaks.txt)
        line = {"metric": "patient x cluster logliks/sec (batch prediction)", "value": pairs / (float(t.item()) * 1e-3),
                "unit": "logliks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(t.item()) / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "batch prediction: {} synthetic patients x {} clusters (m_k ~ {}), patients sharded"
                                       .format(P, n_act, int(np.mean([eng.cluster_size(int(s)) for s in slots if s >= 0])))},
                "e2e": {"value": pairs / (float(t.item()) * 1e-3), "unit": "logliks/s",
                        "h2d_bytes_per_step": (b1[0] - b0[0]) // args.steps, "d2h_bytes_per_step": (b1[1] - b0[1]) // args.steps},
                "gpu_launches": int(eng.launch_count() - l0),
                "roofline": {"kernel": "k_score_wide<8> (64-column tiles)", "bound": "fp64", "achieved": prof["alg_flops"] / (prof["ms"] * 1e-3) / 1e12,
                             "peak": peak, "unit": "TFLOP/s", "frac": prof["alg_flops"] / (prof["ms"] * 1e-3) / 1e12 / peak,
                             "traffic": None, "peak_source": "measured DMMA.8x8x4 register-loop peak (profiles/r1_fp64_peaks.txt)",
                             "share_of_step": prof["ms"] / ms},
                "check": {"rows_sum_to_one": bool(np.allclose(full.sum(1), 1.0, atol=1e-9)), "shape": list(full.shape)}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_chains(args):
    """configs[3]: independent chains (seed = chain id) of `chain-patients` patients dealt round-robin to the
    ranks; within a rank the chains run concurrently, one engine (stream) and one host thread each."""
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mogp_b200 import _capi, MoGP, parallel
    from mogp_b200.synth import alsfrs_like
    mine = parallel.chains_for_rank(args.chains, world, rank)
    N = args.chain_patients
    K0 = max(1, round(N / 50))

    def make(cid):
        eng = _capi.Engine(local, blocking_sync=len(mine) > 2)   # many chain threads share the host cores
        X, Y, z0 = alsfrs_like(N, seed=100 + cid, n_classes=K0, return_labels=True)
        z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
        mix = MoGP(X=X, Y=Y, alpha=K0 / np.log10(N), rand_seed=cid, normalize=True, noise_variance=0.5,
                   refit=args.chain_refit, init_z=z0, engine=eng)
        mix.refit_concurrency = 1          # the chains themselves already run concurrently
        np.random.seed(cid)
        mix.initialize_sampler(K0, True)
        rs = np.random.RandomState(cid)
        return eng, mix, rs

    chains = [make(c) for c in mine]

    def sweep(item):
        eng, mix, rs = item
        perm = rs.permutation(N)
        u = rs.random_sample(N)
        mix.sweep(perm, np.zeros(N), u)
        return eng.launch_count()

    # every chain has its own engine and streams; the host threads that drive them are bounded by the cores
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    pool = ThreadPoolExecutor(max_workers=max(1, min(len(chains), max(4, 2 * cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))))
    for _ in range(args.warmup):
        list(pool.map(sweep, chains))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = sum(c[0].launch_count() for c in chains)
    t0 = time.perf_counter()
    start = torch.cuda.Event(enable_timing=True); stop = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        list(pool.map(sweep, chains))
    for c in chains:
        c[0].sync()
    stop.record(); torch.cuda.synchronize()
    ms = max(start.elapsed_time(stop), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = parallel.gather_chain_results({c: dict(z=chains[i][1].z, Nk=chains[i][1].allocmodel.Nk) for i, c in enumerate(mine)},
                                        args.chains)
    if rank == 0:
        line = {"metric": "Gibbs sweeps/sec over independent chains", "value": args.chains * args.steps / (float(t.item()) * 1e-3),
                "unit": "chain-sweeps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(t.item()) / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "{} independent MoGP (unconstrained) chains x {} patients, refit='{}', "
                                       "dealt round-robin to the GPUs".format(args.chains, N, args.chain_refit)},
                "gpu_launches": int(sum(c[0].launch_count() for c in chains) - l0),
                "check": {"chains_gathered": len(res), "active_clusters_chain0": int((res[0]['Nk'] > 0).sum())}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """Print THE json line on the process's real stdout (see main: fd 1 is pointed at stderr while the run lasts, so
    that banners printed from native code - NCCL's version line, for one - cannot add lines to stdout)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "predict":
        run_predict(args)
    elif args.workload == "chains":
        run_chains(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
