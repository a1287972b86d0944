#!/usr/bin/env python
"""bench.py — Gibbs sweeps/sec at 10k patients (BASELINE.json `metric`, configs[2]) on N B200s.

One "step" = one collapsed Gibbs sweep of MoGP_constrained over all patients (mogp_constrained.py:231-284)
followed by the hyper-parameter refit of every active cluster (refit='sweep', the schedule BASELINE.json names
for the 10k-patient configuration).  N > 1 runs N independent chains (seed = rank), one per GPU — the
alpha/seed-sweep sharding of SURVEY.md §8(e); a single chain is strictly sequential ("replicas only").

  value     whole-job sweeps/s, patient table resident in HBM (C-ABI chain calls + the engine's refit driver)
  e2e       the same sweeps (the chain is restored to the same snapshot first) through the public API
            `MoGP_constrained.sweep()` with the patient table re-sent from host memory every sweep and the results
            (z, Nk, log-likelihoods) read back
  roofline  the scoring product T = M Kx (k_score_wide), timed with CUDA events on the engine's stream
  cpu_baseline / --impl reference   the CPU oracle (NumPy/SciPy restatement of the reference + GPy), on a
            bounded sample of the same sweep, all host cores; the same steps double as a parity check of the engine
            (`parity`: teacher-forced score vectors at the headline size)
  secondary the other configurations BASELINE.json names, in the same run: the reference's refit-after-every-move
            schedule at 10k patients (configs[2], 'move'), the first sweep from a k-means start, batch prediction with
            the patients sharded over the ranks (configs[4]) and independent 2k-patient chains dealt to the ranks
            (configs[3]); `--workload predict|chains` run those two alone at any size.
"""
import argparse
import json
import os

import sys

# Before anything creates the CUDA context (mogp_b200/__init__.py explains): one hardware queue per stream of the look-ahead
# pipeline of a chain that has the GPU to itself.  Many chains sharing a GPU (`--workload chains` alone) do better with the
# default 8 (9.6 vs 8.1 chain-sweeps/s); inside the default run that leg inherits the 32 of the headline leg.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS",
                      "8" if "chains" in [a for i, a in enumerate(sys.argv) if i and sys.argv[i - 1] == "--workload"] else "32")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_FALLBACK = 37.2   # TFLOP/s, DMMA.8x8x4 measured on this pool in round 1 (profiles/r1_fp64_peaks.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patients", type=int, default=10000)
    ap.add_argument("--clusters", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--cpu-steps", type=int, default=12, help="patient-steps of the bounded CPU sample")
    ap.add_argument("--cpu-move-steps", type=int, default=2,
                    help="patient-steps of the CPU sample under the reference's refit-after-every-move schedule (~20 s each)")
    ap.add_argument("--move-steps", type=int, default=64, help="patient-steps of the engine's 'move'-schedule leg")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "predict", "chains"],
                    help="sweep: the headline (BASELINE configs[2]) + secondary legs; predict: batch prediction, patients "
                         "sharded (configs[4]); chains: independent 2k-patient chains dealt to the GPUs (configs[3])")
    ap.add_argument("--query-patients", type=int, default=1000000)
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--chain-patients", type=int, default=2000)
    ap.add_argument("--chain-refit", default="sweep", choices=["sweep", "move"],
                    help="chains workload: 'move' = the reference's schedule (L-BFGS-B after every move)")
    return ap.parse_args()


def workload(args, seed):
    from mogp_b200.synth import alsfrs_like
    # patients drawn from `clusters` latent progression classes; the chain starts from those classes (the
    # reference starts from k-means - the `first_sweep_from_kmeans` leg; either way BASELINE's configuration is
    # "~30 active clusters")
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


def blas_threads(n=None):
    """Set (n given) and report the BLAS thread count the CPU legs run with: torchrun exports OMP_NUM_THREADS=1, which
    would silently halve the CPU baseline."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        if n:
            threadpool_limits(limits=int(n))
        return max([int(d.get("num_threads", 1)) for d in threadpool_info() if d.get("user_api") == "blas"] or [1])
    except Exception:
        return None


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


def measure_fp64_peak(device):
    """The FP64 roofline denominator, measured at bench start on the GPU the bench runs on: scripts/peaks/fp64_peak.cu
    (register-resident DMMA.8x8x4 / DFMA loops; tcgen05 has no FP64 kind, DMMA is the FP64 tensor instruction) compiled
    for sm_100a and run once.  MEASURED_PEAKS.json carries no FP64 entry.  Falls back to the round-1 figure."""
    src = os.path.join(ROOT, "scripts", "peaks", "fp64_peak.cu")
    exe = os.path.join(ROOT, "scripts", "peaks", "fp64_peak")
    try:
        if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
            subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-o", exe, src], check=True, capture_output=True, timeout=300)
        env = dict(os.environ)
        env.setdefault("CUDA_VISIBLE_DEVICES", str(device))       # the probe runs on device 0 of what it sees
        out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120, env=env).stdout
        vals = [float(line.split("TFLOP/s")[0].split(":")[-1]) for line in out.splitlines() if line.startswith("DMMA.8x8x4")]
        clk = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=30).stdout.strip()
        if vals:
            return max(vals), "measured at bench start: scripts/peaks/fp64_peak.cu DMMA.8x8x4 register loop (SM clock after the probe / max: {} MHz)".format(clk)
    except Exception as e:  # noqa: BLE001
        return FP64_FALLBACK, "fallback: round-1 measurement 37.2 TFLOP/s (profiles/r1_fp64_peaks.txt); probe failed: {}".format(type(e).__name__)
    return FP64_FALLBACK, "fallback: round-1 measurement 37.2 TFLOP/s (profiles/r1_fp64_peaks.txt)"


# --------------------------------------------------------------------------------------------------
def oracle_at_start(args, seed=0, **over):
    """The CPU oracle at the chain's initial state."""
    import oracle
    X, Y, kw = workload(args, seed)
    kw.update(over)
    orc = oracle.MoGP_constrained(X=X, Y=Y, **kw)
    orc.initialize_sampler(orc.num_init_clusters, True)
    orc.allocmodel.calc_suffstats(orc.z)
    return orc


def oracle_refit_time(orc, refit_clusters):
    act = np.where(orc.allocmodel.Nk > 0)[0]
    t0 = time.perf_counter()
    evals = 0
    for k in act[:refit_clusters]:
        m = orc.obsmodel[k].model
        e0 = m.n_evals
        m.optimize()
        evals += m.n_evals - e0
    return (time.perf_counter() - t0) / max(refit_clusters, 1), len(act), evals


def oracle_step_times(orc, n_steps, refit_clusters):
    perm = np.random.permutation(np.arange(orc.allocmodel.N))[:n_steps]
    t0 = time.perf_counter()
    for n in perm:
        orc.gibbs_step(n)
    t_step = (time.perf_counter() - t0) / max(n_steps, 1)
    t_refit, nact, evals = oracle_refit_time(orc, refit_clusters)
    return t_step, t_refit, nact, evals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count()
    nthreads = blas_threads(ncores)
    orc = oracle_at_start(args)
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
            "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": ncores, "blas_threads": nthreads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing shared by the legs of one run."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self, eng=None):
        if eng is not None:
            eng.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_b200(args):
    D = Dist()
    torch, rank, local, world = D.torch, D.rank, D.local, D.world
    from mogp_b200 import _capi, MoGP_constrained
    from mogp_b200.utils import pack_patients

    fp64_peak, fp64_src = measure_fp64_peak(local) if rank == 0 else (FP64_FALLBACK, "")
    D.barrier()

    eng = _capi.Engine(local)
    X, Y, kw = workload(args, seed=rank)
    mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
    mix.initialize_sampler(args.clusters, True)
    off, px, py = pack_patients(mix.X, mix.Y)
    N = args.patients

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

    def all_launches():   # the chain's engine (its refit lanes report through it) + worker engines of the SciPy refit mode
        return eng.launch_count() + sum(w.launch_count() for w in (mix._workers or []))

    for _ in range(args.warmup):
        device_sweep()
    # both timed legs start from THIS state: the chain is restored to the snapshot before each of them, so `value` and
    # `e2e` time the same sweeps (same moves, same refits)
    snap = mix.snapshot()

    mix.restore(snap)
    D.barrier(eng)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = all_launches()
    split[0] = split[1] = 0.0
    evals0 = mix.n_refit_evals
    eng.profile(True)
    eng.profile_read(reset=True)
    cnt0 = eng.chain_counters()
    rc0 = eng.refit_counters()
    prof_range = os.environ.get("MOGP_PROFILE_RANGE") == "1"   # ncu --profile-from-start off: capture the timed sweeps only
    if prof_range:
        torch.cuda.profiler.start()
    eng.timer_start()
    for _ in range(args.steps):
        device_sweep()
    ms = eng.timer_stop()
    if prof_range:
        torch.cuda.profiler.stop()
    cnt1 = eng.chain_counters()
    rc1 = eng.refit_counters()
    counters = {k: cnt1[k] - cnt0[k] for k in cnt1}
    prof = eng.profile_read(reset=True)
    eng.profile(False)
    launches = all_launches() - l0
    split_timed = (split[0] / args.steps, split[1] / args.steps)
    evals_timed = mix.n_refit_evals - evals0
    D.barrier(eng)
    ms_max = D.max_over_ranks(ms)
    z_value = mix.z.copy()
    moved = int(np.sum(mix.z != snap['z']))

    # e2e: the same sweeps through the public API with host buffers
    mix.restore(snap)
    D.barrier(eng)
    b0 = eng.transfer_bytes()
    eng.timer_start()
    for _ in range(args.steps):
        ll = e2e_sweep()
    ms_e2e = eng.timer_stop()
    b1 = eng.transfer_bytes()
    D.barrier(eng)
    ms_e2e_max = D.max_over_ranks(ms_e2e)
    clocks = sampler.stop() if rank == 0 else None
    same_sweeps = bool(np.array_equal(mix.z, z_value))

    # the dominant kernel on its own (no concurrent work): one batch of the look-ahead pipeline = the first patients of a
    # permutation up to the pipeline's column budget, against every active cluster of the final state
    iso = None
    if rank == 0:
        cols = int(os.environ.get("MOGP_LOOKAHEAD", "64"))
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
        D.dist.all_gather(out, zt)

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else \
            (6650.0, "fallback (B200_PROFILING.md)")
        # two passes can be in flight, the later one queued on the SMs behind the earlier one: the device time the passes took
        # is the UNION of their (start, stop) intervals, not the sum of the bracketed launches (reported beside it)
        union_ms = counters.get("passes_union_ms", 0.0) or prof["ms"]
        sec = union_ms * 1e-3
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
        roof.update({"kernel": "k_score_wide<NS> (T = M Kx for a batch of patients x every active cluster, FP64 DMMA.8x8x4)",
                     "traffic": traffic,
                     "traffic_note": "dram bytes per launch from one ncu --set full capture of this kernel in a sweep "
                                     "(profiles/score_gemm_traffic.json), not from this run's chain state",
                     "launches": prof["launches"],
                     "avg_launch_us": union_ms / max(prof["launches"], 1) * 1e3,
                     "avg_launch_us_bracketed": prof["ms"] / max(prof["launches"], 1) * 1e3,
                     "alg_bytes_per_launch": prof["alg_bytes"] / max(prof["launches"], 1),
                     "alg_flops_per_launch": prof["alg_flops"] / max(prof["launches"], 1),
                     "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac, "peak_source": hbm_src},
                     "fp64": {"achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp_frac, "peak_source": fp64_src},
                     "pairs_per_launch": counters["pairs"] / max(prof["launches"], 1),
                     "share_of_step": union_ms / ms,
                     "note": "achieved = whole passes (every active cluster) timed live in the sweeps, where they share the SMs "
                             "with the kernels of the moves; two passes are in flight (the later one fills the earlier one's "
                             "tail), so the launch duration is the union of the passes' CUDA-event intervals divided by their "
                             "number ('avg_launch_us_bracketed' = start to stop of one launch, its wait behind the earlier "
                             "pass included); 'isolated' = the same kernel alone; 'rescoring' = the short passes against the "
                             "clusters whose chain of corrections broke (latency bound, accounted apart)",
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
                        "h2d_bytes_per_step": (b1[0] - b0[0]) // args.steps, "d2h_bytes_per_step": (b1[1] - b0[1]) // args.steps,
                        "same_sweeps_as_value": same_sweeps},
                "gpu_launches": int(launches),
                "roofline": roof,
                "clocks": clocks,
                "chain": {"active_clusters": len(occupancy), "patients_moved_in_timed_sweeps": moved,
                          "pair_logliks_per_s": world * args.steps * N * len(occupancy) / (ms_max * 1e-3),
                          "refit_evals_per_sweep": evals_timed / args.steps,
                          "refit_evals_on_device_per_sweep": (rc1["launched"] - rc0["launched"]) / args.steps,
                          "total_loglik": ll,
                          "steps_s_per_sweep": split_timed[0], "refit_s_per_sweep": split_timed[1],
                          "optimizer": mix.optimizer or os.environ.get("MOGP_OPTIMIZER", "native"),
                          "start": "sweeps {}..{} of a chain started from the generator's classes".format(args.warmup + 1, args.warmup + args.steps),
                          "lookahead": {k: counters[k] for k in ("window_passes", "rescoring_passes", "window_patients",
                                                                 "open_s", "rescore_s", "commit_s", "rank_corrections", "correct_s", "replayed",
                                                                 "window_pairs", "pruned_pairs", "chained_wait_s",
                                                                 "chained_corrections")}}}
        if world == 1 and not args.no_cpu_baseline:
            line.update(cpu_and_parity(args, local))
    secondary = {}
    if not args.no_secondary:
        D.barrier(eng)
        secondary["first_sweep_from_kmeans"] = leg_first_sweep(args, D, eng)
        secondary["move_schedule_10k"] = leg_move_schedule(args, D, eng)
        eng.close()
        eng = None
        secondary["batch_prediction"] = leg_predict(args, D, min(args.query_patients, 1000000), steps=1, warmup=1)
        secondary["independent_chains"] = leg_chains(args, D, steps=1, warmup=1)
    if rank == 0:
        if secondary:
            line["secondary"] = secondary
        emit(line)
    D.close()


def cpu_and_parity(args, device):
    """The CPU baseline (oracle port on the box's host cores, bounded sample) and, on the very same patient-steps, the
    parity of the engine at the headline size: a second engine at the initial state is fed the oracle's choices and its
    score vectors / probabilities are compared step by step."""
    from mogp_b200 import _capi, MoGP_constrained
    from mogp_b200._capi import host_choice
    ncores = os.cpu_count()
    nthreads = blas_threads(ncores)
    orc = oracle_at_start(args)
    orc.step_log = []
    X, Y, kw = workload(args, 0)
    eng2 = _capi.Engine(device)
    mix2 = MoGP_constrained(X=X, Y=Y, engine=eng2, **kw)
    mix2.initialize_sampler(args.clusters, True)
    N = args.patients
    perm = np.random.permutation(np.arange(N))[:args.cpu_steps]
    t_cpu, worst_s, worst_p, same_draws = 0.0, 0.0, 0.0, True
    for n in perm:
        st = np.random.get_state()
        a = np.random.randn(1, 1)[0, 0]
        u = np.random.random_sample()
        np.random.set_state(st)
        t0 = time.perf_counter()
        orc.gibbs_step(n)
        t_cpu += time.perf_counter() - t0
        rec = orc.step_log[-1]
        eng2.chain_step_begin(n)
        ids, score = eng2.chain_step_score(a)
        idx, p, _m = host_choice(score, u)
        if list(ids) != list(rec['active']) or not np.array_equal(np.isinf(score), np.isinf(rec['score'])):
            worst_s = float('inf')
        else:
            fin = np.isfinite(score)
            worst_s = max(worst_s, float(np.max(np.abs(score[fin] - rec['score'][fin]) / np.maximum(np.abs(rec['score'][fin]), 1e-300))))
            worst_p = max(worst_p, float(np.max(np.abs(p - rec['p']))))
            same_draws = same_draws and int(ids[idx]) == int(rec['z'])
        eng2.chain_step_commit(rec['z'], a)
    ts = t_cpu / max(len(perm), 1)
    tr, nact, evals = oracle_refit_time(orc, 1)
    # the same cluster refitted by the engine from the same state: same optimum, and (same optimiser logic) the same
    # number of evaluations
    _z, Nk2, slots2 = eng2.chain_state(N)
    k0 = int(np.where(orc.allocmodel.Nk > 0)[0][0])
    n_eng = int(eng2.refit_clusters([int(slots2[k0])])[0])
    th_e = eng2.cluster_get_theta(int(slots2[k0]))
    th_o = np.asarray(orc.obsmodel[k0].model.param_array, dtype=float)
    eng2.close()
    sweep_s = N * ts + nact * tr
    ok = bool(worst_s <= 1e-9 and worst_p <= 1e-9 and same_draws)
    if not ok:
        sys.stderr.write("bench: PARITY FAILURE at the headline size: score {:.3e} p {:.3e} draws {}\n".format(worst_s, worst_p, same_draws))
    return {"cpu_baseline": {"value": 1.0 / sweep_s, "unit": "sweeps/s", "cores": ncores, "blas_threads": nthreads, "kind": "port",
                             "sample": "{} consecutive patient-steps ({:.3f} s each) + L-BFGS-B refit of 1 cluster ({:.2f} s, "
                                       "{} evaluations) from the same initial state; sweep = N*step + active*refit"
                                       .format(len(perm), ts, tr, evals)},
            "parity": {"ok": ok, "parity_max_rel": worst_s, "max_abs_dp": worst_p, "same_draws": same_draws,
                       "steps": int(len(perm)),
                       "what": "engine fed the oracle's choices for the CPU baseline's own patient-steps at 10k patients: "
                               "relative error of the score vectors (log prior + log likelihood per active cluster), |dp| of the "
                               "membership probabilities, draws reproduced; bound 1e-9",
                       "refit_evaluations": {"oracle_scipy": int(evals), "engine_native": n_eng},
                       "refit_theta_rel_diff": float(np.max(np.abs(th_e[-len(th_o):] - th_o) / np.abs(th_o)))}}


# ---- secondary legs -----------------------------------------------------------------------------------------------
def leg_first_sweep(args, D, eng):
    """The reference's own start (mogp_constrained.py:198-207): k-means on the first-visit time, K0 = `clusters`; ONE sweep
    from there (most patients move: the regime where the look-ahead batches see the most invalidations)."""
    from mogp_b200 import MoGP_constrained
    X, Y, kw = workload(args, seed=D.rank)
    kw = dict(kw, init_z=None, num_init_clusters=args.clusters)
    mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
    t0 = time.perf_counter()
    mix.initialize_sampler(args.clusters, True)
    t_init = time.perf_counter() - t0
    z0 = mix.z.copy()
    D.barrier(eng)
    eng.timer_start()
    mix.sweep()
    ms = eng.timer_stop()
    ms_max = D.max_over_ranks(ms)
    return {"metric": "Gibbs sweeps/sec @10k patients, first sweep from the k-means start", "value": D.world / (ms_max * 1e-3),
            "unit": "sweeps/s", "ms_per_step": ms_max, "patients_moved": int(np.sum(mix.z != z0)),
            "active_clusters_after": int(np.sum(mix.allocmodel.Nk > 0)), "init_s": t_init, "steps": 1, "warmup": 0,
            "note": "k-means (sklearn, n_init=10) + initial factorisations are init_s, outside the timed sweep"}


def leg_move_schedule(args, D, eng):
    """The reference's schedule (mogp_constrained.py:229,284): L-BFGS-B refit of the receiving cluster after EVERY move,
    at 10k patients, S consecutive patient-steps from the initial state; sweeps/s = 1 / (N * mean step time).  Rank 0."""
    if D.rank != 0:
        return None
    from mogp_b200 import MoGP_constrained
    X, Y, kw = workload(args, seed=0)
    kw = dict(kw, refit='move')
    mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
    mix.initialize_sampler(args.clusters, True)
    N, S = args.patients, args.move_steps
    perm = np.random.permutation(np.arange(N))
    mix.sweep(perm=perm[:4])                                        # warm-up steps
    e0 = mix.n_refit_evals
    eng.sync()
    eng.timer_start()
    mix.sweep(perm=perm[4:4 + S])
    ms = eng.timer_stop()
    out = {"metric": "Gibbs sweeps/sec @10k patients, refit after every move (the reference's schedule)",
           "value": 1.0 / (N * ms * 1e-3 / S), "unit": "sweeps/s", "ms_per_patient_step": ms / S, "patient_steps": S,
           "refit_evals_per_step": (mix.n_refit_evals - e0) / S, "n_gpus": 1,
           "note": "S consecutive patient-steps from the initial state, extrapolated to a sweep of N steps"}
    if D.world == 1 and not args.no_cpu_baseline and args.cpu_move_steps > 0:
        blas_threads(os.cpu_count())
        orc = oracle_at_start(args, refit='move')
        e0 = sum(m.model.n_evals for m in orc.obsmodel.values() if m is not None)
        t0 = time.perf_counter()
        for n in perm[4:4 + args.cpu_move_steps]:
            orc.gibbs_step(n)
        dt = (time.perf_counter() - t0) / args.cpu_move_steps
        e1 = sum(m.model.n_evals for m in orc.obsmodel.values() if m is not None)
        out["cpu_baseline"] = {"value": 1.0 / (N * dt), "unit": "sweeps/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "{} patient-steps of the same schedule from the same state ({:.1f} s each, {:.1f} "
                                         "evaluations per step)".format(args.cpu_move_steps, dt, (e1 - e0) / args.cpu_move_steps)}
    return out


def predict_model(local, K=40):
    from mogp_b200 import _capi, MoGP_constrained
    from mogp_b200.synth import alsfrs_like
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
    return eng, mix, Nk, slots


def leg_predict(args, D, P, steps, warmup, fp64_peak=FP64_FALLBACK):
    """configs[4]: P synthetic patients scored against a random-init 40-cluster model, patients sharded across the
    ranks (no collective on the data path), one all_gather of the probability matrix at the end."""
    from mogp_b200 import parallel
    from mogp_b200.synth import alsfrs_like_fast
    eng, mix, Nk, slots = predict_model(D.local)
    lo, hi = parallel.shard_bounds(P, D.world, D.rank)
    # every rank generates only its own shard (same generator, seed = shard): synthetic queries, nothing to broadcast
    Xq, Yq = alsfrs_like_fast(hi - lo, seed=2 + D.rank, n_classes=40)
    a = np.random.RandomState(3 + D.rank).randn(hi - lo)
    nq = hi - lo
    for _ in range(max(1, warmup)):
        mix.predict_batch(Xq[:min(nq, 20000)], Yq[:min(nq, 20000)], a[:min(nq, 20000)])
    D.barrier(eng)
    l0 = eng.launch_count(); b0 = eng.transfer_bytes()
    eng.profile(True); eng.profile_read(reset=True)
    eng.timer_start()
    for _ in range(steps):
        probs = mix.predict_batch(Xq, Yq, a)
    ms = eng.timer_stop()
    prof = eng.profile_read(reset=True); eng.profile(False)
    b1 = eng.transfer_bytes()
    t = D.max_over_ranks(ms)
    full = parallel.all_gather_rows(probs, P) if D.world > 1 else probs
    D.barrier(eng)
    out = None
    if D.rank == 0:
        n_act = int((Nk > 0).sum())
        pairs = float(P) * n_act * steps
        tf = prof["alg_flops"] / (prof["ms"] * 1e-3) / 1e12
        out = {"metric": "patient x cluster logliks/sec (batch prediction)", "value": pairs / (t * 1e-3),
               "unit": "logliks/s", "n_gpus": D.world, "steps": steps, "warmup": warmup,
               "ms_per_step": t / steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "batch prediction: {} synthetic patients x {} clusters (m_k ~ {}), patients sharded"
                                      .format(P, n_act, int(np.mean([eng.cluster_size(int(s)) for s in slots if s >= 0])))},
               "e2e": {"value": pairs / (t * 1e-3), "unit": "logliks/s",
                       "h2d_bytes_per_step": (b1[0] - b0[0]) // steps, "d2h_bytes_per_step": (b1[1] - b0[1]) // steps},
               "gpu_launches": int(eng.launch_count() - l0),
               "roofline": {"kernel": "k_score_wide<8> (64-column tiles)", "bound": "tensor", "achieved": tf,
                            "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "traffic": None,
                            "share_of_step": prof["ms"] / ms},
               "check": {"rows_sum_to_one": bool(np.allclose(full.sum(1), 1.0, atol=1e-9)), "shape": list(full.shape)}}
    eng.close()
    return out


def leg_chains(args, D, steps, warmup, n_chains=None):
    """configs[3]: independent chains (seed = chain id) of `chain-patients` patients dealt round-robin to the
    ranks; within a rank the chains run concurrently, one engine (stream) and one host thread each."""
    torch = D.torch
    from concurrent.futures import ThreadPoolExecutor
    from mogp_b200 import _capi, MoGP, parallel
    from mogp_b200.synth import alsfrs_like
    n_chains = n_chains or args.chains
    mine = parallel.chains_for_rank(n_chains, D.world, D.rank)
    N = args.chain_patients
    K0 = max(1, round(N / 50))

    def make(cid):
        eng = _capi.Engine(D.local, blocking_sync=len(mine) > 2)   # many chain threads share the host cores
        X, Y, z0 = alsfrs_like(N, seed=100 + cid, n_classes=K0, return_labels=True)
        z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
        mix = MoGP(X=X, Y=Y, alpha=K0 / np.log10(N), rand_seed=cid, normalize=True, noise_variance=0.5,
                   refit=args.chain_refit, init_z=z0, engine=eng)
        mix.refit_concurrency = 1          # the chains themselves already run concurrently
        mix.refit_lanes = 4
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
    nthr = max(1, min(len(chains), max(4, 2 * cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", D.world))))))
    pool = ThreadPoolExecutor(max_workers=nthr)
    for _ in range(warmup):
        list(pool.map(sweep, chains))
    D.barrier()
    l0 = sum(c[0].launch_count() for c in chains)
    t0 = time.perf_counter()
    start = torch.cuda.Event(enable_timing=True); stop = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        list(pool.map(sweep, chains))
    for c in chains:
        c[0].sync()
    stop.record(); torch.cuda.synchronize()
    ms = max(start.elapsed_time(stop), (time.perf_counter() - t0) * 1e3)
    t = D.max_over_ranks(ms)
    res = parallel.gather_chain_results({c: dict(z=chains[i][1].z, Nk=chains[i][1].allocmodel.Nk) for i, c in enumerate(mine)},
                                        n_chains)
    out = None
    if D.rank == 0:
        out = {"metric": "Gibbs sweeps/sec over independent chains", "value": n_chains * steps / (t * 1e-3),
               "unit": "chain-sweeps/s", "n_gpus": D.world, "steps": steps, "warmup": warmup,
               "ms_per_step": t / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": "{} independent MoGP (unconstrained) chains x {} patients, refit='{}', "
                                      "dealt round-robin to the GPUs".format(n_chains, N, args.chain_refit)},
               "gpu_launches": int(sum(c[0].launch_count() for c in chains) - l0),
               "check": {"chains_gathered": len(res), "active_clusters_chain0": int((res[0]['Nk'] > 0).sum())}}
    pool.shutdown()
    for c in chains:
        c[0].close()
    return out


def run_single_leg(args):
    D = Dist()
    if args.workload == "predict":
        line = leg_predict(args, D, args.query_patients, args.steps, args.warmup)
    else:
        line = leg_chains(args, D, args.steps, args.warmup)
    if D.rank == 0:
        emit(line)
    D.close()


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
    elif args.workload in ("predict", "chains"):
        run_single_leg(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
