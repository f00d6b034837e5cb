#!/usr/bin/env python
"""bench.py -- LM iterations/s of the pose-graph optimisation hot path (BASELINE.json metric).

One "step" = one `optimize(15)` of the workload graph, the call the reference makes per key-frame
(reference src/sparse_gslam/src/drone.cpp:150). Default workload = C5, the 1M-pose grid world of BASELINE.json
(the configuration the metric "LM iterations/sec at 1/2/4/8 B200" is quoted on; it fits one GPU), LM with the
analytic pose-line Jacobian. `--workload c1|c2|c3|c4` select the other BASELINE configs.

  value  : LM iterations/s with the graph resident in HBM (sgb_optimize_resident), CUDA-event timed, max over ranks
  e2e    : the same metric through the C ABI with HOST buffers: sgb_set_graph (host symbolic phase + H2D of the
           whole graph) + sgb_optimize + sgb_get_estimates (D2H) inside the timed region
  roofline: dominant kernel k_pcg; achieved = algorithmic bytes / CUDA-event time of the launches
  cpu_baseline / --impl reference: the CPU oracle (g2o-equivalent restatement, 1 thread) on a bounded sample

Launch: `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun, one rank per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LM_ITERS = 15  # reference: drone.cpp:150 optimize(15, ...)
GN_ITERS = 20  # reference: submap_loop_closer.cpp:287 optimize(20)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.t_mark = None

    def mark(self):
        """Start of the timed region: nvidia-smi needs ~0.5 s to come up, so the sampler is started during the warm-up
        and only the samples that arrive after mark() are reported (if a very short timed region saw none, the last
        warm-up samples -- same kernels, same load -- are used and the fact is recorded)."""
        self.t_mark = time.perf_counter()

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        lines = [ln for (t, ln) in self.lines if self.t_mark is None or t >= self.t_mark]
        if len(lines) < 2 and self.t_mark is not None:
            lines = [ln for (_, ln) in self.lines[-5:]]
            window = "timed region + end of warm-up (timed region shorter than the sampling period)"
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


C5_DESC = "C5 synthetic 1M-pose grid world (P=1e6, L=2e5, E_o~1.5e6, E_l=4e6), LM-15"


def make_workload(name, rank=0, world=1):
    from sparse_gslam_b200 import capi
    from sparse_gslam_b200 import graphgen as gg
    name = name.lower()
    if name == "c5":
        g = gg.make_c5()
        return g, capi.ALGO_LM, LM_ITERS, C5_DESC
    if name == "c5s":  # small C5 used by CI-style quick runs
        g = gg.make_c5(rows=200, cols=200)
        return g, capi.ALGO_LM, LM_ITERS, "C5-shaped 200x200 grid world (P=4e4), LM-15"
    if name in ("c1", "c2", "c3", "c4"):
        g = gg.make(name)
        return g, capi.ALGO_LM, LM_ITERS, f"{name.upper()} {g.name} (P={g.P}, L={g.L}, E_o={g.n_pp}, E_l={g.n_pl}), LM-15"
    if name in ("c1gn", "c2gn", "c3gn"):
        g0 = gg.make(name[:2])
        g = g0.pose_only(phi=g0.meta.get("dcs_phi", 1.0))
        return g, capi.ALGO_GN, GN_ITERS, f"{name[:2].upper()} pose graph + DCS closures (P={g.P}, E_o={g.n_pp}), GN-20"
    raise SystemExit(f"unknown workload {name}")


def graph_h2d_bytes(g):
    tot = 0
    for k in ("pose_id", "pose_est", "pose_fixed", "lm_id", "lm_est", "lm_fixed", "pp_i", "pp_j", "pp_z", "pp_info",
              "pp_phi", "pp_seq", "pl_pose", "pl_lm", "pl_z", "pl_info", "pl_seq"):
        tot += getattr(g, k).nbytes
    return int(tot)


def pcg_bytes_per_iteration(st):
    """Algorithmic bytes of one PCG iteration on the implicit Schur complement:
    Hpp blocks 76 B (72 values + 4 index), Hpl blocks 52 B read twice (pose-major and landmark-major pass),
    (Hll+lambda I)^-1 24 B and t 2x16 B per landmark, vectors 312 B per pose and the preconditioner 144 B per pose
    (three rows of the inverse 12x12 block of the pose's 4-row chunk, single precision; 72 B with the former 3x3 blocks)."""
    P, L = st["n_free_poses"], st["n_free_landmarks"]
    nnzb_pp = P + 2 * st["n_pairs_pp"]
    n_pl = st["n_pairs_pl"]
    return 76 * nnzb_pp + 2 * 52 * n_pl + 56 * L + (312 + 144) * P


def cpu_reference_run(workload, steps, warmup, quiet=False, budget_s=25.0):
    """The reference arm / cpu_baseline: the CPU oracle (g2o-equivalent restatement: serial edge loop with g2o-numeric
    Jacobians, exact sparse LDLt with a minimum-degree ordering, one thread like the reference build).

    C1-C3: the full graph. C5 (1M poses) cannot be finished on one core in bench time (hours: the direct solver's
    cost grows super-linearly with the graph), so the arm runs C5-SHAPED samples of growing size -- the same generator at
    100x100, 150x150, 200x200, ... cells -- until `budget_s` is spent, fits t = a * P^b to the measured LM-15 times and
    extrapolates to P = 1e6. The line says so: `sample` names the sizes, `measured` lists them, `config.workload` of the
    reference arm names the largest sample actually run."""
    from oracle.cpu_oracle import ALGO_GN, ALGO_LM, JAC_G2O_NUMERIC, Oracle
    from sparse_gslam_b200 import capi
    from sparse_gslam_b200 import graphgen as gg

    def timed(g, algo, iters, reps, warm):
        times, its, prof = [], 0, None
        for s in range(warm + reps):
            o = Oracle(g)
            o.initialize_optimization()
            t0 = time.perf_counter()
            n, _ = o.optimize(iters, algo, JAC_G2O_NUMERIC)
            dt = time.perf_counter() - t0
            if s >= warm:
                times.append(dt)
                its += max(n, 0)
                prof = o.profile()
            if sum(times) > 60:
                break
        return times, its, prof

    if workload == "c4":
        return cpu_reference_c4(steps, warmup)
    if workload == "c5":
        full_P = 1_000_000
        measured, spent = [], 0.0
        for side in (100, 150, 200, 250, 300):
            if measured and spent + 3.5 * measured[-1]["seconds"] > budget_s:   # the next size costs ~3x the last one
                break
            g = gg.make_c5(rows=side, cols=side)
            times, its, prof = timed(g, ALGO_LM, LM_ITERS, 1, 0)
            measured.append({"P": int(g.P), "L": int(g.L), "E_l": int(g.n_pl), "seconds": times[0], "lm_iterations": its,
                             "factor_nnz": float(prof["nnzL"])})
            spent += times[0]
        big = measured[-1]
        if len(measured) >= 2:
            xs = np.log([m["P"] for m in measured])
            ys = np.log([m["seconds"] / max(1, m["lm_iterations"]) for m in measured])
            b, a = np.polyfit(xs, ys, 1)
            sec_per_it_full = float(np.exp(a + b * np.log(full_P)))
        else:
            b = 1.0
            sec_per_it_full = big["seconds"] / max(1, big["lm_iterations"]) * full_P / big["P"]
        sample = ("C5-SHAPED SAMPLES, not the 1M-pose graph: LM-15 measured at P = " + ", ".join(str(m["P"]) for m in measured) +
                  f" (same generator, smaller grid); seconds per LM iteration fitted as a*P^b, b = {b:.2f}, and extrapolated "
                  f"to P = 1e6 (a linear scaling of the largest sample would give "
                  f"{big['lm_iterations'] / big['seconds'] * big['P'] / full_P:.4f} LM it/s)")
        return dict(value=1.0 / sec_per_it_full, unit="LM iterations/s", cores=1, kind="port", sample=sample,
                    sample_ms_per_step=1e3 * big["seconds"], sample_steps=1, host_cores=os.cpu_count(), measured=measured,
                    fit_exponent=float(b), largest_sample=f"C5-shaped {int(round(big['P'] ** 0.5))}x{int(round(big['P'] ** 0.5))} "
                                                          f"grid world (P={big['P']}, L={big['L']}, E_l={big['E_l']})",
                    extrapolated=True)
    g, a, iters, _ = make_workload(workload)
    algo = ALGO_LM if a == capi.ALGO_LM else ALGO_GN
    times, its, prof = timed(g, algo, iters, steps, warmup)
    total = sum(times)
    sample = (f"full {workload} graph, optimize({iters}) with g2o-numeric Jacobians and exact sparse LDLt "
              f"(nnz(L) = {int(prof['nnzL'])}, {1e3 * prof['t_solve'] / max(1, its):.1f} ms of factorise+solve per LM iteration)")
    return dict(value=its / total if total > 0 else 0.0, unit="LM iterations/s", cores=1, kind="port", sample=sample,
                sample_ms_per_step=1e3 * total / max(1, len(times)), sample_steps=len(times), host_cores=os.cpu_count(),
                factor_nnz=float(prof["nnzL"]), extrapolated=False)


C4_GRAPHS_PER_GPU = 128  # BASELINE config 4: 1 024 independent windows over 8 GPUs


def lm_bytes_per_iteration(g, st_info, pcg_iters, trials):
    """Algorithmic bytes of LM iterations: B_lin + per trial (n_pcg * B_pcg + B_chi + update)."""
    P, L = st_info["n_free_poses"], st_info["n_free_landmarks"]
    e_o, e_l = g.n_pp, g.n_pl
    b_lin = 80 * e_o + 48 * e_l + 120 * P + 64 * L + 72 * e_o + 48 * e_l
    b_chi = 80 * e_o + 48 * e_l + 24 * P + 16 * L
    b_pcg = 76 * (P + 2 * e_o) + 2 * 52 * e_l + 56 * L + 384 * P
    return b_lin, b_chi + 2 * (24 * P + 16 * L), b_pcg


def run_c4(args, rank, world, local_rank, warmup):
    """BASELINE config 4: batched independent aces-shaped windows, C4_GRAPHS_PER_GPU per GPU, no exchange between
    ranks (weak scaling). One step = sgb_optimize_batch(LM-15) over this rank's graphs."""
    import torch
    import torch.distributed as dist
    from sparse_gslam_b200 import SparseOptimizerB200, capi
    from sparse_gslam_b200 import graphgen as gg
    from sparse_gslam_b200.optimizer import optimize_batch
    n = C4_GRAPHS_PER_GPU
    graphs = [gg.make_c4_window(seed=1000 + rank * n + i) for i in range(n)]
    opts = [SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC, pcg_tolerance=args.pcg_tol, device=local_rank)
            for _ in range(n)]

    def init_all():
        for o, g in zip(opts, graphs):
            assert o.initialize_optimization(g)

    t0 = time.perf_counter()
    init_all()
    t_setgraph = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        optimize_batch(opts, LM_ITERS, resident=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, tot_iters, pcg_iters, trials = 0.0, 0, 0, 0
    for _ in range(args.steps):
        done, stats = optimize_batch(opts, LM_ITERS, resident=True)
        tot_iters += sum(max(d, 0) for d in done)
        dev_ms += opts[0].timings()["total_ms"]
        pcg_iters += sum(o.timings()["pcg_iters"] for o in opts)
    barrier()
    clocks = sampler.stop()
    e2e = None
    if not args.no_e2e:
        h2d = sum(graph_h2d_bytes(g) for g in graphs)
        d2h = sum(int(g.pose_est.nbytes + g.lm_est.nbytes) for g in graphs)
        e_steps, e_iters = max(1, min(args.steps, 3)), 0
        barrier()
        te0 = time.perf_counter()
        for _ in range(e_steps):
            init_all()
            done, _ = optimize_batch(opts, LM_ITERS)
            for o in opts:
                o.estimates()
            e_iters += sum(max(d, 0) for d in done)
        barrier()
        te = time.perf_counter() - te0
        e2e = [e_iters, te, h2d, d2h, e_steps]
    if world > 1:
        tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ss = torch.tensor([float(tot_iters), float(e2e[0]) if e2e else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(ss, op=dist.ReduceOp.SUM)
        dev_ms_max, all_iters = float(tt[0]), float(ss[0])
        if e2e:
            te_t = torch.tensor([e2e[1]], dtype=torch.float64, device="cuda")
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
            e2e[0], e2e[1] = float(ss[1]), float(te_t[0])
    else:
        dev_ms_max, all_iters = dev_ms, float(tot_iters)
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    value = all_iters / (dev_ms_max * 1e-3)
    line = {
        "metric": "LM iterations/s", "value": value, "unit": "LM iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": dev_ms_max / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 batched aces-shaped windows: {n} independent graphs per GPU x {world} GPU(s) "
                               f"(P=128, L=64, E_o=127, E_l=400 each), LM-15, one thread block per graph",
                   "algorithm": "LM", "iterations_per_step": LM_ITERS, "jacobian": "analytic",
                   "pcg_tolerance": args.pcg_tol, "l2": "graphs fit L2 (latency-bound config; see DESIGN.md)",
                   "parallelism": f"{world} GPU(s), independent graphs, no data-path exchange"},
        "optimize_ms": dev_ms_max / max(1, args.steps), "lm_iterations": tot_iters, "graphs_per_gpu": n,
        "set_graph_s": t_setgraph, "gpu_launches": args.steps, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_lm_block", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                     "traffic": None, "peak_source": peak_src, "pcg_iterations": pcg_iters,
                     "note": "one CTA per graph, working set L1/L2-resident: latency-bound, not an HBM roofline case"},
    }
    if e2e:
        line["e2e"] = {"value": e2e[0] / e2e[1], "unit": "LM iterations/s", "h2d_bytes_per_step": e2e[2] * world,
                       "d2h_bytes_per_step": e2e[3] * world, "ms_per_step": 1e3 * e2e[1] / e2e[4], "steps": e2e[4]}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference_run("c4", 1, 0)
    print(json.dumps(line))


def cpu_reference_c4(steps, warmup, sample=16):
    """CPU oracle on a sample of the C4 windows, one after the other on one core (the reference is single-threaded)."""
    from oracle.cpu_oracle import ALGO_LM, JAC_G2O_NUMERIC, Oracle
    from sparse_gslam_b200 import graphgen as gg
    graphs = [gg.make_c4_window(seed=1000 + i) for i in range(sample)]
    total, its, n = 0.0, 0, 0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        k = 0
        for g in graphs:
            o = Oracle(g)
            o.initialize_optimization()
            m, _ = o.optimize(LM_ITERS, ALGO_LM, JAC_G2O_NUMERIC)
            k += max(m, 0)
        dt = time.perf_counter() - t0
        if s >= warmup:
            total += dt
            its += k
            n += 1
    # the fair multi-core companion of the batched GPU number -- one graph per host thread on all cores
    # (the oracle is handle-based C++ behind ctypes, which releases the GIL)
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    many = [gg.make_c4_window(seed=1000 + i) for i in range(max(sample, 4 * cores))]

    oracles = [Oracle(g) for g in many]  # built outside the timed region: packing the graph is Python (holds the GIL)
    for o in oracles:
        o.initialize_optimization()

    def one(o):
        o.set_estimates(o.g.pose_est, o.g.lm_est)
        m, _ = o.optimize(LM_ITERS, ALGO_LM, JAC_G2O_NUMERIC)
        return max(m, 0)

    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(one, oracles))  # warm-up
        its_all, dt_all = 0, 0.0
        for _ in range(3):  # best of three: the host threads of a shared box are not always spread over the cores
            t0 = time.perf_counter()
            k = sum(ex.map(one, oracles))
            dt = time.perf_counter() - t0
            if dt_all == 0.0 or k / dt > its_all / dt_all:
                its_all, dt_all = k, dt
    return dict(value=its / total if total > 0 else 0.0, unit="LM iterations/s", cores=1, kind="port",
                sample=f"{sample} of the C4 windows (seeds 1000..), optimize(15) each, g2o-numeric Jacobians + exact sparse "
                       "LDLt, sequential on one core; iterations/s is per core and independent of the batch size",
                sample_ms_per_step=1e3 * total / max(1, n), sample_steps=n, host_cores=os.cpu_count(),
                all_cores=dict(value=its_all / dt_all if dt_all > 0 else 0.0, unit="LM iterations/s", cores=cores,
                               sample=f"{len(many)} windows, one graph per host thread on {cores} threads"))


class _OracleBackend:
    """CPU arm of the key-frame stream: the oracle behind the same session protocol (host push/pop like g2o)."""

    def __init__(self):
        self.o, self.stack = None, []

    def initialize(self, g):
        from oracle.cpu_oracle import Oracle
        self.o = Oracle(g)
        return self.o.initialize_optimization()

    def push(self):
        self.stack.append(self.o.estimates())

    def pop(self):
        self.o.set_estimates(*self.stack.pop())

    def discard_top(self):
        self.stack.pop()

    def optimize(self, iters, online):
        from oracle.cpu_oracle import ALGO_LM, JAC_G2O_NUMERIC
        return self.o.optimize(iters, ALGO_LM, JAC_G2O_NUMERIC)[0]

    def active_chi2(self):
        return self.o.chi2()[0]

    def estimates(self):
        return self.o.estimates()


def run_stream(args, warmup):
    """the reference's per-key-frame protocol (drone.cpp:146-190) on a synthetic key-frame stream
    replayed from the C1 intel-lab-shaped graph: every key-frame re-initialises the (growing) landmark graph from HOST
    buffers, push, optimize(15), chi2 gate, pop / discardTop, estimates back to the host. End to end by construction."""
    from sparse_gslam_b200 import capi
    from sparse_gslam_b200 import graphgen as gg
    from sparse_gslam_b200.session import GpuBackend, LandmarkGraphSession, stream_from_graph
    g = gg.make("c1")
    frames = stream_from_graph(g)[:args.stream_frames]

    def one_pass(backend):
        s = LandmarkGraphSession(backend)
        t0 = time.perf_counter()
        s.run(frames)
        return time.perf_counter() - t0, s

    be = GpuBackend(jacobian_mode=capi.JAC_ANALYTIC)
    for _ in range(min(warmup, 1)):
        one_pass(be)
    sampler = ClockSampler(0)
    sampler.start()
    for k in be.prof:
        be.prof[k] = 0
    tot, its, kfs, h2d, d2h = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        dt, s = one_pass(be)
        tot += dt
        its += sum(r.iterations for r in s.log)
        kfs += len(s.log)
        h2d += sum(48 * r.n_edges + 24 * r.n_poses + 16 * r.n_landmarks for r in s.log)
        d2h += sum(24 * r.n_poses + 16 * r.n_landmarks for r in s.log)
    clocks = sampler.stop()
    peak, peak_src = measured_peaks()
    line = {"metric": "LM iterations/s", "value": its / tot, "unit": "LM iterations/s", "n_gpus": 1, "steps": args.steps,
            "warmup": warmup, "ms_per_step": 1e3 * tot / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"N1 key-frame stream: first {len(frames)} key-frames of the C1 intel-lab-shaped graph through "
                                   "the reference's per-key-frame protocol (re-initialise, push, optimize(15), 0.99 chi2 gate, "
                                   "pop/discardTop), host buffers every key-frame", "jacobian": "analytic",
                       "l2": "graph fits L2 (latency-bound config)"},
            "keyframes_per_s": kfs / tot, "ms_per_keyframe": 1e3 * tot / max(1, kfs), "clocks": clocks,
            "gpu_launches": None,
            "roofline": {"bound": "hbm", "kernel": "k_pcg", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                         "traffic": None, "peak_source": peak_src, "note": "per-call overhead dominated (small growing graph)"},
            "e2e": {"value": its / tot, "unit": "LM iterations/s", "h2d_bytes_per_step": h2d // max(1, args.steps),
                    "d2h_bytes_per_step": d2h // max(1, args.steps)}}
    calls = max(1, be.prof["calls"])
    line["per_keyframe"] = {k[:-2] + "_ms": 1e3 * v / calls for k, v in be.prof.items() if k.endswith("_s")}
    line["per_keyframe"].update({"device_ms": be.prof["device_ms"] / calls, "pcg_iters": be.prof["pcg_iters"] / calls,
                                 "lm_trials": be.prof["trials"] / calls, "kernel_launches": be.prof["kernel_launches"] / calls})
    line["gpu_launches"] = int(be.prof["kernel_launches"])
    if not args.no_cpu_baseline:
        dt, s = one_pass(_OracleBackend())
        line["cpu_baseline"] = dict(value=sum(r.iterations for r in s.log) / dt, unit="LM iterations/s", cores=1, kind="port",
                                    sample="the same key-frame stream through the same protocol on the CPU oracle (g2o-numeric "
                                           "Jacobians, exact sparse LDLt)", ms_per_keyframe=1e3 * dt / max(1, len(s.log)))
    print(json.dumps(line))


def run_edits(args, warmup):
    """SURVEY.md 8f N3 + N4 at the size of the C5 graph: the producers of its information matrices (1M odometry
    intervals, 1M line segments, 3.6M scan points) and the pose-graph copy of its 1M-pose chain, each through the C ABI
    with HOST buffers (e2e) and with the device time of the kernels alone (resident). Units = information matrices
    (odometry + pose-line edges), point covariances and re-measured poses produced per second. The CPU oracle runs the
    same rows on a 1/20 sample, one thread like the reference."""
    from oracle import cpu_oracle as co
    from sparse_gslam_b200 import frontend as fe
    from sparse_gslam_b200.posegraph import PoseGraphB200
    rng = np.random.default_rng(5)
    n_kf, steps_per_kf = 1_000_000, 10
    deltas = np.stack([0.05 + 0.01 * rng.normal(size=n_kf * steps_per_kf), 0.002 * rng.normal(size=n_kf * steps_per_kf),
                       0.01 * rng.normal(size=n_kf * steps_per_kf)], 1)
    seg = (np.arange(n_kf + 1) * steps_per_kf).astype(np.int32)
    n_seg, ppl = 1_000_000, 16
    t = np.tile(np.linspace(-1.0, 1.0, ppl), n_seg)
    th = np.repeat(rng.uniform(-np.pi, np.pi, n_seg), ppl)
    rho = np.repeat(rng.uniform(0.5, 5.0, n_seg), ppl)
    lpts = np.stack([rho * np.cos(th) - t * np.sin(th), rho * np.sin(th) + t * np.cos(th)], 1)
    lpts = (lpts + 0.01 * rng.normal(size=lpts.shape)).astype(np.float32)
    lcov = np.tile(np.array([1e-4, 1e-5, 1e-5, 2e-4], np.float32), (n_seg * ppl, 1))
    lseg = (np.arange(n_seg + 1) * ppl).astype(np.int32)
    nw, ns, sz = 1000, 10, 360
    ang = np.linspace(-np.pi, np.pi, sz, endpoint=False)
    beam = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)
    sdel = np.stack([0.05 + 0.01 * rng.normal(size=(nw, ns - 1)), 0.002 * rng.normal(size=(nw, ns - 1)),
                     0.01 * rng.normal(size=(nw, ns - 1))], -1)
    r = rng.uniform(0.5, 5.0, size=(nw, ns, sz)).astype(np.float32)
    spts = np.stack([r * beam[:, 0], r * beam[:, 1]], -1).astype(np.float32)
    d = np.stack([0.5 + 0.02 * rng.normal(size=n_kf), 0.02 * rng.normal(size=n_kf), 0.2 * rng.normal(size=n_kf)], 1)
    a = np.concatenate([[0.3], 0.3 + np.cumsum(d[:, 2])])
    chain = np.stack([np.concatenate([[0.0], np.cumsum(d[:, 0] * np.cos(a[:-1]) - d[:, 1] * np.sin(a[:-1]))]),
                      np.concatenate([[0.0], np.cumsum(d[:, 0] * np.sin(a[:-1]) + d[:, 1] * np.cos(a[:-1]))]),
                      a - 2 * np.pi * np.floor((a + np.pi) / (2 * np.pi))], 1)
    cinfo = np.tile(np.array([2500.0, 0, 0, 2500.0, 0, 1e4]), (n_kf, 1))
    pg = PoseGraphB200()

    def chain_copy():
        pg.reset(chain[0], 0)
        pg.append_from_host(chain, cinfo)
        return pg.info()["last_edit_ms"]

    # algorithmic bytes per unit (each array once): see DESIGN.md section 8
    rows = [
        ("k_odom_information", n_kf, 24 * steps_per_kf + 4 + 24 + 72 + 48,
         lambda: fe.odom_information(deltas, seg, 0.02, 0.02, 0.01)[3],
         lambda k: co.odom_information(deltas[:k * steps_per_kf], seg[:k + 1], 0.02, 0.02, 0.01)),
        ("k_line_fit", n_seg, 24 * ppl + 4 + 8 + 16 + 24,
         lambda: fe.line_fit_information(lpts, lcov, lseg)[3],
         lambda k: co.line_fit_information(lpts[:k * ppl], lcov[:k * ppl], lseg[:k + 1])),
        ("k_scan_frames+k_scan_points", nw * ns * sz, 8 + 16 + 8 + 1,
         lambda: fe.scan_point_covariances(sdel, beam, spts, 0.02, 0.02, 0.01, 9e-4)[3],
         lambda k: co.scan_point_covariances(sdel[:max(1, k // (ns * sz))], beam, spts[:max(1, k // (ns * sz))], 0.02, 0.02, 0.01, 9e-4)),
        ("k_pg_remeasure+k_pg_chain", n_kf, 24 + 24 + 24 + 24,
         chain_copy,
         lambda k: co.pg_append(chain[0], chain[:k + 1])),
    ]
    peak, peak_src = measured_peaks()
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(warmup):
        for (_, _, _, f, _) in rows:
            f()
    sampler.mark()
    out, tot_units, tot_wall, tot_kernel = [], 0, 0.0, 0.0
    for (name, units, bpu, f, cpu) in rows:
        kms, wall = 0.0, 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            kms += f()
            wall += time.perf_counter() - t0
        kms /= args.steps
        wall /= args.steps
        gbs = units * bpu / (kms * 1e-3) / 1e9
        rec = {"kernel": name, "units": units, "kernel_ms": kms, "e2e_ms": 1e3 * wall, "units_per_s": units / (kms * 1e-3),
               "e2e_units_per_s": units / wall, "algorithmic_bytes_per_unit": bpu, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        if not args.no_cpu_baseline:
            k = max(1, units // 20)
            t0 = time.perf_counter()
            cpu(k)
            rec["cpu_units_per_s"] = k / (time.perf_counter() - t0)
        out.append(rec)
        tot_units += units
        tot_wall += wall
        tot_kernel += kms * 1e-3
    clocks = sampler.stop()
    top = max(out, key=lambda r: r["kernel_ms"])
    line = {"metric": "producer + edit rows/s", "value": tot_units / tot_kernel, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
            "warmup": warmup, "ms_per_step": 1e3 * tot_kernel, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (odometry, pose chain) / f32 (scan points, line fit: the reference's types)", "data": "synthetic",
            "config": {"workload": "N3+N4 rows at C5 size: 1M odometry intervals x 10 deltas, 1M line segments x 16 points, "
                                   "1000 windows x 10 scans x 360 beams, copy of a 1M-pose chain into the pose graph",
                       "l2": "inputs larger than L2 except the scan points (29 MB)"},
            "kernels": out, "gpu_launches": 6 * args.steps, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_GBps"], "peak": peak, "unit": "GB/s",
                         "frac": top["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src},
            "e2e": {"value": tot_units / tot_wall, "unit": "rows/s",
                    "h2d_bytes_per_step": int(deltas.nbytes + seg.nbytes + lpts.nbytes + lcov.nbytes + lseg.nbytes + sdel.nbytes +
                                              spts.nbytes + beam.nbytes + chain.nbytes + cinfo.nbytes),
                    "d2h_bytes_per_step": int(n_kf * (24 + 72 + 48) + n_seg * (8 + 16 + 24) + nw * ns * sz * 25)}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = {"value": sum(r["units"] for r in out) / sum(r["units"] / r["cpu_units_per_s"] for r in out),
                                "unit": "rows/s", "cores": 1, "kind": "port",
                                "sample": "the first 1/20 of every row's input on the CPU oracle (oracle/sgo_frontend.cpp)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SGB_WORKLOAD", "c5"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 comparison with the one-GPU result")
    ap.add_argument("--pcg-tol", type=float, default=1e-10)
    ap.add_argument("--stream-frames", type=int, default=400, help="key-frames replayed by --workload stream")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(int(os.environ.get("SGB_MIN_WARMUP", "3")), args.warmup)  # SGB_MIN_WARMUP=1 only for ncu captures

    if args.impl == "reference":
        if rank != 0:
            return
        if args.workload.lower() in ("stream", "edits"):
            # these side workloads time their CPU arm inside the GPU line itself (cpu_baseline / per-kernel cpu_units_per_s)
            print(json.dumps({"impl": "reference", "unavailable": f"--workload {args.workload} reports its CPU arm in the "
                              "cpu_baseline of its own line; the reference arm exists for c1..c5"}))
            return
        r = cpu_reference_run(args.workload.lower(), max(1, args.steps), 1, budget_s=120.0)
        # config.workload names what this arm ACTUALLY ran; where that is not the GPU arm's graph it says so
        wl = args.workload.lower()
        if wl == "c5":
            desc = (f"{C5_DESC} -- CPU ARM RAN SAMPLES ONLY: largest = {r['largest_sample']}; value extrapolated to P=1e6 by the "
                    f"fitted law t ~ P^{r['fit_exponent']:.2f} (see cpu_baseline.sample / measured)")
            algo_name, iters = "LM", LM_ITERS
        elif wl == "c4":
            desc = (f"C4 batched aces-shaped windows: {C4_GRAPHS_PER_GPU} independent graphs per GPU x {args.gpus} GPU(s) "
                    "(P=128, L=64, E_l=400 each), LM-15")
            algo_name, iters = "LM", LM_ITERS
        elif wl in ("stream", "edits"):
            desc, algo_name, iters = wl, "LM", LM_ITERS
        else:
            from sparse_gslam_b200 import capi as _capi
            _, a, iters, desc = make_workload(wl)
            algo_name = "LM" if a == _capi.ALGO_LM else "GN+DCS"
        line = {"impl": "reference", "metric": "LM iterations/s", "value": r["value"], "unit": "LM iterations/s",
                "n_gpus": args.gpus, "steps": r["sample_steps"], "warmup": 1, "ms_per_step": r["sample_ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "algorithm": algo_name, "iterations_per_step": iters,
                           "jacobian": "g2o-numeric (the reference's default)", "parallelism": "host CPU, 1 thread (g2o + Eigen Simplicial are serial)"},
                "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "LM iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from sparse_gslam_b200 import SparseOptimizerB200, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the backend has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload.lower() == "stream":
        if rank == 0:
            run_stream(args, warmup)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload.lower() == "edits":
        if rank == 0:
            run_edits(args, warmup)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload.lower() == "c4":
        run_c4(args, rank, world, local_rank, warmup)
        if world > 1:
            dist.destroy_process_group()
        return

    from sparse_gslam_b200 import dist as sdist
    g, algo, iters, desc = make_workload(args.workload, rank, world)
    opt = SparseOptimizerB200(algo, jacobian_mode=capi.JAC_ANALYTIC, pcg_tolerance=args.pcg_tol, device=local_rank)

    def init_graph():
        # world == 1: sgb_set_graph; world > 1: row-block partition + NVLink peer-memory rendezvous
        if world == 1:
            return opt.initialize_optimization(g)
        return opt.initialize_partitioned(g, world, rank, sdist.exchange_blobs)

    t0 = time.perf_counter()
    assert init_graph()
    t_setgraph = time.perf_counter() - t0
    if world == 1:
        st = opt.structure()
        # distinct off-diagonal pairs from the block list
        nP = st["n_free_poses"]
        offd = st["row"] != st["col"]
        st["n_pairs_pp"] = int(np.sum(offd & (st["col"] < nP)))
        st["n_pairs_pl"] = int(np.sum(offd & (st["col"] >= nP)))
    else:
        # every rank only holds its own share of the structure (rank-filtered symbolic phase): count the distinct free
        # vertex pairs of the WHOLE graph from the edge lists, the same quantities the one-GPU block list gives
        pf, lf = g.pose_fixed == 0, g.lm_fixed == 0
        a, b = np.minimum(g.pp_i, g.pp_j).astype(np.int64), np.maximum(g.pp_i, g.pp_j).astype(np.int64)
        m = pf[g.pp_i] & pf[g.pp_j]
        mpl = pf[g.pl_pose] & lf[g.pl_lm]
        st = {"n_free_poses": int(pf.sum()), "n_free_landmarks": int(lf.sum()),
              "n_pairs_pp": int(np.unique(a[m] * g.P + b[m]).size),
              "n_pairs_pl": int(np.unique(g.pl_pose[mpl].astype(np.int64) * g.L + g.pl_lm[mpl]).size)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident (device-timed) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        opt.optimize(iters, resident=True)
    barrier()
    sampler.mark()
    tot_iters = 0
    agg = dict(linearize_ms=0.0, setup_ms=0.0, pcg_ms=0.0, update_ms=0.0, total_ms=0.0, pcg_iters=0, trials=0,
               linearizations=0, kernel_launches=0)
    pcg_phase = [0.0] * 4
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        n, stats = opt.optimize(iters, resident=True)
        last_stats = stats
        tot_iters += max(n, 0)
        t = opt.timings()
        for k in agg:
            agg[k] += t[k]
        pcg_phase = [a + b for a, b in zip(pcg_phase, t["pcg_phase_ms"])]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_ms = agg["total_ms"]
    if world > 1:
        tt = torch.tensor([dev_ms, float(tot_iters)], dtype=torch.float64, device="cuda")
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max = float(mx[0])
        total_iters_all = float(tot_iters)  # one partitioned job: every rank runs the SAME LM iterations
    else:
        dev_ms_max, total_iters_all = dev_ms, float(tot_iters)
    value = total_iters_all / (dev_ms_max * 1e-3) if dev_ms_max > 0 else 0.0

    # ---------------- end to end through the C ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        h2d = graph_h2d_bytes(g)
        d2h = int(g.pose_est.nbytes + g.lm_est.nbytes)
        e_iters, e_steps = 0, max(1, min(args.steps, 3))
        # one untimed end-to-end step first (the resident warm-up does not touch the host path: heap growth and first-touch
        # page faults of the ~1 GB of host-side structure arrays belong to the first call only)
        assert init_graph()
        opt.optimize(iters)
        opt.estimates()
        barrier()
        te0 = time.perf_counter()
        for _ in range(e_steps):
            assert init_graph()                      # host symbolic phase + H2D of the graph (+ peer rendezvous)
            n, _ = opt.optimize(iters)
            opt.estimates()                          # D2H of the result
            e_iters += max(n, 0)
        barrier()
        te = time.perf_counter() - te0
        if world > 1:
            tt = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt[0])
        e2e = {"value": e_iters / te, "unit": "LM iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * te / e_steps, "steps": e_steps}

    # ---------------- multi-GPU correctness inside the scaling run itself ----------------
    # The partitioned job's final state (estimates are replicated on every rank) against the SAME optimize() on ONE GPU,
    # run by rank 0 on its own device after the timed regions: final chi2, largest pose / landmark difference, LM and
    # PCG iteration counts. The scaling record then carries its own correctness evidence.
    parity = None
    if world > 1 and not args.no_parity:
        opt.optimize(iters, resident=True)
        chi_n = opt.active_chi2()          # collective: every rank takes part in the cross-GPU reduction
        p_n, l_n = opt.estimates()
        t_n = opt.timings()
        barrier()
        if rank == 0:
            one = SparseOptimizerB200(algo, jacobian_mode=capi.JAC_ANALYTIC, pcg_tolerance=args.pcg_tol, device=local_rank)
            assert one.initialize_optimization(g)
            n1, _ = one.optimize(iters)
            chi_1 = one.active_chi2()
            p_1, l_1 = one.estimates()
            t_1 = one.timings()
            one.close()
            dth = p_n[:, 2] - p_1[:, 2]
            dth -= 2 * np.pi * np.round(dth / (2 * np.pi))
            parity = {"chi2_n": chi_n[0], "chi2_1": chi_1[0], "chi2_rel": abs(chi_n[0] - chi_1[0]) / max(abs(chi_1[0]), 1e-300),
                      "max_pose_xy_delta": float(np.abs(p_n[:, :2] - p_1[:, :2]).max()),
                      "max_pose_theta_delta": float(np.abs(dth).max()),
                      "max_landmark_delta": float(np.abs(l_n - l_1).max()) if l_1.size else 0.0,
                      "pcg_iterations": [int(t_n["pcg_iters"]), int(t_1["pcg_iters"])],
                      "lm_trials": [int(t_n["trials"]), int(t_1["trials"])],
                      "ok": bool(abs(chi_n[0] - chi_1[0]) <= 1e-6 * abs(chi_1[0]) and
                                 np.abs(p_n[:, :2] - p_1[:, :2]).max() <= 1e-6 * max(1.0, float(np.abs(p_1[:, :2]).max())))}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    b_iter = pcg_bytes_per_iteration(st) // world  # per GPU: rows are split evenly
    pcg_bytes = b_iter * agg["pcg_iters"]
    pcg_s = agg["pcg_ms"] * 1e-3
    achieved = pcg_bytes / pcg_s / 1e9 if pcg_s > 0 else 0.0
    # DRAM bytes (read + write) of ONE ncu-captured k_pcg launch of this workload (profiles/pcg_traffic.json, with the
    # PCG iterations of that launch, so that it can be set against that launch's algorithmic bytes); null if no capture
    traffic, traffic_detail = None, None
    tp = os.path.join(ROOT, "profiles", "pcg_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic_detail = json.load(open(tp)).get(args.workload)
            traffic = traffic_detail["dram_bytes_per_launch"] if traffic_detail else None
        except Exception:
            traffic, traffic_detail = None, None
    roofline = {"bound": "hbm", "kernel": "k_pcg", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peak_src,
                "algorithmic_bytes_per_pcg_iteration": b_iter, "pcg_iterations": agg["pcg_iters"],
                "pcg_launches": agg["trials"], "share_of_step": agg["pcg_ms"] / dev_ms if dev_ms > 0 else None,
                "pcg_phase_us_per_iteration": [1e3 * v / max(1, agg["pcg_iters"]) for v in pcg_phase],
                # PCG iterations of every k_pcg launch of one step, in launch order (maps an ncu capture of launch #k
                # to its algorithmic bytes)
                "pcg_iterations_per_lm_iteration": [int(x["pcg_iters"]) for x in last_stats]}
    # the linearise + assemble stage against the same roofline (B_lin per evaluation), rank 0's share
    b_lin = (80 * g.n_pp + 48 * g.n_pl + 120 * st["n_free_poses"] + 64 * st["n_free_landmarks"] + 72 * st["n_pairs_pp"] +
             48 * st["n_pairs_pl"]) // world
    lin_s = agg["linearize_ms"] * 1e-3
    lin_gbs = b_lin * agg["linearizations"] / lin_s / 1e9 if lin_s > 0 else 0.0
    roofline_lin = {"bound": "hbm", "kernel": "k_lin_pose + k_lin_lm (+ k_finalize_lin)", "achieved": lin_gbs, "peak": peak,
                    "unit": "GB/s", "frac": lin_gbs / peak, "algorithmic_bytes_per_linearization": b_lin,
                    "linearizations": agg["linearizations"], "share_of_step": agg["linearize_ms"] / dev_ms if dev_ms > 0 else None}
    line = {
        "metric": "LM iterations/s", "value": value, "unit": "LM iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": dev_ms_max / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "algorithm": "LM" if algo == capi.ALGO_LM else "GN+DCS",
                   "iterations_per_step": iters, "jacobian": "analytic", "pcg_tolerance": args.pcg_tol,
                   "l2": "inputs larger than L2 (Hessian + edge arrays >> 126 MB)" if g.P >= 200000 else
                         "graph fits L2 (latency-bound config; see DESIGN.md)",
                   "parallelism": "1 GPU" if world == 1 else
                                  f"{world} GPUs, row-block partition of the reduced pose system, halo gathers + PCG "
                                  "reductions through NVLink peer memory"},
        "optimize_ms": dev_ms_max / max(1, args.steps),
        "lm_iterations": tot_iters, "wall_s": wall,
        "phases_ms": {k: agg[k] for k in ("linearize_ms", "setup_ms", "pcg_ms", "update_ms", "total_ms")},
        "lm_trials": agg["trials"], "set_graph_s": t_setgraph,
        "gpu_launches": int(agg["kernel_launches"]),
        "clocks": clocks, "roofline": roofline, "roofline_linearize": roofline_lin,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if parity is not None:
        line["parity_vs_n1"] = parity
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference_run(args.workload, 1, 0)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
