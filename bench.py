#!/usr/bin/env python
"""bench.py — headline benchmark of the batched min-snap hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): a batch of 65,536 random 3-D, 10-segment,
N = 10 min-snap problems per GPU (random vertices in a +-10 m box, Nfabian
segment times v_max = 3, a_max = 5 — the reference's createRandomVertices +
estimateSegmentTimes recipe, vectorised with numpy's RNG). One "step" = one
mtg_solve_argmin_batch over that batch: the solve (Q/A/R construction, the R_pp solve, all
300 coefficients, the cost and the status of every trajectory written to HBM) with the
argmin that folds the batch into the running best candidate fused into the kernel's
epilogue (the sweep of BASELINE configs[4]; --two-launch-step runs it as mtg_solve_batch +
mtg_argmin_batch, the same results in two launches);
the timed region ends with the one collective of the sharded sweep: an all-gather of
one 16-byte {cost, index} pair per rank (torch.distributed / NCCL), also at N = 1.

  value : trajectories solved / s, whole job, inputs resident in HBM (CUDA events)
  e2e   : the same through the C ABI with PINNED HOST buffers (H2D + kernel + D2H
          inside the timed region, staged by the library)
  roofline    : algorithmic bytes (2,752 B / trajectory) / kernel time vs measured HBM peak
  cpu_baseline: the oracle (a port of the reference algorithm; Eigen is not on the image)
                on all host cores over a bounded sample
  sweep       : BASELINE configs[3] beside it (not part of `value`): evaluateRange and the fused
                v/a/tube feasibility sweep, 1000 samples per trajectory, samples/s and fraction of
                the measured HBM peak (rank 0, N = 1 only; --no-sweep skips it)
`--impl reference` times that CPU port alone (rank 0 only).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_SEG, DIM, NCOEF, DERIV = 10, 3, 10, 4
BATCH_PER_GPU = 65536
BYTES_IN = ((K_SEG + 1) * DIM + K_SEG) * 8          # 344
BYTES_OUT = (K_SEG * DIM * NCOEF + 1) * 8           # 2,408
BYTES_PER_TRAJ = BYTES_IN + BYTES_OUT               # 2,752  
FLOP_PER_TRAJ_REF = 8.8e4                           # reference formulation
N_ROTATE = 4                                        # rotating buffer sets: 4 x 180 MB > L2 (126 MB)
METRIC = "min-snap trajectories solved/sec (N=10, 10 seg, 3D)"
# dram__bytes_read.sum + dram__bytes_write.sum of one 65,536-solve launch: a CONSTANT from the named ncu capture
# (ncu cannot run inside the bench), not measured in this run
TRAFFIC_PER_LAUNCH = 123.15e6
TRAFFIC_SOURCE = "constant from profiles/r02_solve_full.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
SOLVE_KERNEL_NAME = "solve_canonical_kernel<5,3,false>"
SWEEP_BATCH = 1_000_000                             # BASELINE configs[3]: 1M trajectories x 1000 samples
SWEEP_SAMPLES = 1000


_REAL_STDOUT = None


def claim_stdout():
    """Only the result line may reach stdout (the driver parses it): everything libraries print
    there (e.g. NCCL's version banner) is sent to stderr; emit() writes to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def make_workload(batch, seed):
    """positions [K+1,D,B], times [K,B] (SoA, batch innermost), float64."""
    rng = np.random.RandomState(seed)
    pos = rng.uniform(-10.0, 10.0, size=(K_SEG + 1, DIM, batch))
    # resample vertices closer than 0.2 m to their predecessor (vertex.cpp:65-72)
    for v in range(1, K_SEG + 1):
        while True:
            d = np.sqrt(((pos[v] - pos[v - 1]) ** 2).sum(axis=0))
            bad = d <= 0.2
            if not bad.any():
                break
            pos[v][:, bad] = rng.uniform(-10.0, 10.0, size=(DIM, int(bad.sum())))
    dist = np.sqrt((np.diff(pos, axis=0) ** 2).sum(axis=1))
    v_max, a_max = 3.0, 5.0
    times = dist / v_max * 2 * (1.0 + 6.5 * v_max / a_max * np.exp(-dist / v_max * 2))
    return np.ascontiguousarray(pos), np.ascontiguousarray(times)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Polls NVML (SM clock + clock-event reasons) from a thread while the timed region runs;
    the timed region of this benchmark is milliseconds long, too short for `nvidia-smi -lms`."""

    def __init__(self, index):
        import threading

        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical order; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                index = int(vis.split(",")[index])
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.0005)

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "how": "NVML polled from a thread during the timed region"}


def cpu_port_rate(batch_sample, threads, reps=1, seed=12345):
    """Oracle (port of the reference algorithm) solves/s over a bounded sample."""
    from oracle import pyoracle as po

    pos, times = make_workload(batch_sample, seed)
    pos_aos = np.ascontiguousarray(np.moveaxis(pos, -1, 0))
    times_aos = np.ascontiguousarray(np.moveaxis(times, -1, 0))
    po.solve_canonical_batch(pos_aos[:256], times_aos[:256], n_threads=threads)  # warm
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        po.solve_canonical_batch(pos_aos, times_aos, N=NCOEF, derivative=DERIV, n_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return batch_sample / best, best


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path. Eigen/glog are not on this image, so the
    reference itself cannot be compiled; the oracle port of its algorithm is timed instead
    (kind = "port"), single process, all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 32768   # >= 0.1 s of all-core work per step: the rate no longer swings with thread start-up
    from oracle import pyoracle as po

    pos, times = make_workload(sample, 777)
    pos_aos = np.ascontiguousarray(np.moveaxis(pos, -1, 0))
    times_aos = np.ascontiguousarray(np.moveaxis(times, -1, 0))
    for _ in range(args.warmup):
        po.solve_canonical_batch(pos_aos, times_aos, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.solve_canonical_batch(pos_aos, times_aos, n_threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} solves per step x {args.steps} steps of the same workload, "
                                   "OpenMP over the batch; oracle restatement of the reference algorithm "
                                   "(dense QR for SparseQR), debug prints excluded; Eigen absent so the "
                                   "reference itself cannot be built here"},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def sweep_section(ctx, peak, batch=None, time_it=True):
    """BASELINE configs[3]: evaluateRange + fused v/a/tube feasibility sweep, S ~ 1000 samples per
    trajectory, trajectory-contiguous outputs (the reference's order), device-resident."""
    import torch

    B = batch or SWEEP_BATCH
    S = SWEEP_SAMPLES
    # candidates generated on the device (Philox keyed by the candidate index): no 344 MB host copy
    p, t = ctx.generate_candidates_batch(B, K_SEG, DIM, seed=0xB200, layout="aos")
    sol = ctx.solve_batch(p, t, layout="aos")
    tmax = ctx.max_time_batch(t, layout="aos")
    dt = tmax / S
    Smax = S + 8
    samples = torch.empty((B, Smax, DIM), dtype=torch.float64, device="cuda")
    flags = torch.empty((B, Smax), dtype=torch.uint8, device="cuda")
    radii = torch.full((B, K_SEG, 2), 0.15, dtype=torch.float64, device="cuda")
    n0 = ctx.launch_count

    def timed(fn, reps=7, warm=3):
        if not time_it:   # the -m gpu parity test of this section: one launch, no statistics
            reps, warm = 1, 0
        return _timed(fn, reps, warm)

    def _timed(fn, reps, warm):
        """`warm` untimed launches (the first launch after a different kernel runs 20-30 % slow while the SM
        clock ramps: profiles/r02_sweep_clock_probe.log), then `reps` launches timed one by one with CUDA events
        on the launch stream; returns (median s, best s, SM MHz median under load)."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        smp = ClockSampler(torch.cuda.current_device())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        clk = smp.stop()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        return ms[len(ms) // 2] * 1e-3, ms[0] * 1e-3, clk["sm_mhz"]

    r = ctx.eval_range_batch(sol["coeffs"], t, 0.0, tmax, dt, 0, Smax, layout="aos", out={"samples": samples})
    nsamp = int(r["n_samples"].sum().item())
    out = {"workload": f"configs[3]: {B} trajectories x ~{S} samples (dt = max_time/{S}), AoS outputs; "
                       f"outputs ({samples.numel() * 8 / 1e9:.1f} GB) exceed L2",
           "samples": nsamp, "unit": "samples/s", "hbm_peak_gbs": peak}
    def entry(secs, best, mhz, bps):
        return {"value": nsamp / secs, "ms": secs * 1e3, "ms_best": best * 1e3, "bytes_per_sample": bps,
                "achieved_gbs": bps * nsamp / secs / 1e9, "frac": bps * nsamp / secs / 1e9 / peak,
                "frac_best": bps * nsamp / best / 1e9 / peak, "sm_mhz": mhz,
                "timing": "median of 7 launches timed one by one after 3 warm-up launches (CUDA events)"}

    secs, best, mhz = timed(lambda: ctx.eval_range_batch(sol["coeffs"], t, 0.0, tmax, dt, 0, Smax, layout="aos",
                                                         out={"samples": samples, "n_samples": r["n_samples"]}))
    out["eval_range"] = entry(secs, best, mhz, 24 + 2.48)
    secs, best, mhz = timed(lambda: ctx.feasibility_batch(sol["coeffs"], t, 0.0, tmax, dt, 3.0, 5.0, positions=p,
                                                          radii=radii, max_samples=Smax, layout="aos",
                                                          want_samples=True, out={"samples": samples, "flags": flags}))
    out["feasibility"] = entry(secs, best, mhz, 25 + 2.48)
    secs, best, mhz = timed(lambda: (ctx.extrema_batch(sol["coeffs"], t, 1, layout="aos"),
                                     ctx.extrema_batch(sol["coeffs"], t, 2, layout="aos")), reps=3, warm=2)
    out["extrema_v_and_a"] = {"value": 2 * K_SEG * B / secs, "unit": "root problems/s (degree 15 and 13)",
                              "ms": secs * 1e3, "ms_best": best * 1e3, "sm_mhz": mhz}
    out["gpu_launches"] = int(ctx.launch_count - n0)
    out["parity_sample"] = sweep_parity_sample(ctx, sol["coeffs"], t, p, radii, tmax, dt, Smax, samples, flags,
                                               r["n_samples"])
    return out


def sweep_parity_sample(ctx, coeffs, t, p, radii, tmax, dt, Smax, samples, flags, n_samples, stride=4096):
    """Every `stride`-th trajectory of the sweep that was just timed against the oracle: sample count,
    sampling times and segment index bit-exact (the reference's serial recurrence), sample values to
    1e-12 of the polynomial's scale, flags identical except within 1e-9 of a limit. `samples` / `flags`
    are the buffers the timed feasibility call filled."""
    import torch

    from oracle import pyoracle as po

    idx = torch.arange(0, coeffs.shape[0], stride, device="cuda")
    sub = lambda x: x.index_select(0, idx).contiguous()  # noqa: E731
    c_s, t_s, p_s, r_s, tm_s, dt_s = sub(coeffs), sub(t), sub(p), sub(radii), sub(tmax), sub(dt)
    # the subset once more with sampling times and segment indices (not outputs of the timed calls)
    rr = ctx.eval_range_batch(c_s, t_s, 0.0, tm_s, dt_s, 0, Smax, layout="aos", want_times=True, want_segments=True)
    full_rows = sub(samples).cpu().numpy()
    full_flags = sub(flags).cpu().numpy()
    full_n = sub(n_samples).cpu().numpy()
    rows, acc, seg, n = (rr[k].cpu().numpy() for k in ("samples", "sampling_times", "segment_idx", "n_samples"))
    ch, th, ph, rh, tmh, dth = (x.cpu().numpy() for x in (c_s, t_s, p_s, r_s, tm_s, dt_s))
    res = {"trajectories": int(idx.numel()), "stride": stride, "count_mismatch": 0, "time_mismatch": 0,
           "segment_mismatch": 0, "value_err_max_rel": 0.0, "flag_mismatch_away_from_limits": 0,
           "subset_rows_bit_identical_to_full_run": True}
    for b in range(idx.numel()):
        ref = po.traj_evaluate_range(ch[b], th[b], 0.0, tmh[b], dth[b], 0)
        k = ref[0].shape[0]
        if n[b] != k or full_n[b] != k:
            res["count_mismatch"] += 1
            continue
        res["time_mismatch"] += int(not np.array_equal(acc[b, :k], ref[1]))
        res["segment_mismatch"] += int(not np.array_equal(seg[b, :k], ref[2]))
        if not np.array_equal(rows[b, :k], full_rows[b, :k]):
            res["subset_rows_bit_identical_to_full_run"] = False
        scale = np.abs(ch[b]).reshape(K_SEG, DIM, NCOEF) * (th[b][:, None, None] ** np.arange(NCOEF))
        res["value_err_max_rel"] = max(res["value_err_max_rel"],
                                       float(np.abs(full_rows[b, :k] - ref[0]).max() / scale.sum(-1).max()))
        f = po.feasibility_sweep(ch[b], th[b], ph[b], rh[b], 3.0, 5.0, 0.0, tmh[b], dth[b])
        v = po.traj_evaluate_range(ch[b], th[b], 0.0, tmh[b], dth[b], 1)[0]
        a = po.traj_evaluate_range(ch[b], th[b], 0.0, tmh[b], dth[b], 2)[0]
        nv, na = np.sqrt((v ** 2).sum(1)), np.sqrt((a ** 2).sum(1))
        g = po.tube_geometry(ph[b], rh[b])[ref[2]]
        x = ref[0]
        y = np.einsum("nij,nj->ni", g[:, :9].reshape(-1, 3, 3), x) + g[:, 9:12]
        q = (y ** 2).sum(1)
        al_s = -np.einsum("ni,ni->n", g[:, 12:15], x - g[:, 15:18])
        al_e = np.einsum("ni,ni->n", g[:, 12:15], x - g[:, 18:21])
        near = ((np.abs(nv - 3.0) < 3e-9) | (np.abs(na - 5.0) < 5e-9) | (np.abs(al_s) < 1e-9) | (np.abs(al_e) < 1e-9)
                | (np.abs(q - g[:, 21] ** 2) < 1e-9 * np.maximum(1.0, q)))
        res["flag_mismatch_away_from_limits"] += int(((full_flags[b, :k] != f[1]) & ~near).sum())
    res["ok"] = (res["count_mismatch"] == 0 and res["time_mismatch"] == 0 and res["segment_mismatch"] == 0
                 and res["value_err_max_rel"] <= 1e-12 and res["flag_mismatch_away_from_limits"] == 0
                 and res["subset_rows_bit_identical_to_full_run"])
    return res


def workload_config(n_gpus):
    return {"workload": "configs[1]: batch of 65,536 random 3-D 10-segment min-snap (N=10) solves per GPU",
            "batch_per_gpu": BATCH_PER_GPU, "segments": K_SEG, "dims": DIM, "N": NCOEF,
            "derivative_to_optimize": DERIV, "global_batch": BATCH_PER_GPU * n_gpus,
            "step": "mtg_solve_argmin_batch: solve (coefficients, cost, status written) + running argmin, one launch; "
                    "consecutive steps overlap by programmatic dependent launch (mtg_set_solve_overlap; --no-overlap: off)",
            "parallelism": f"batch sharded over {n_gpus} GPU(s); one 16-byte argmin all-gather per rank at the "
                           "end of the timed region",
            "l2": f"inputs+outputs rotate over {N_ROTATE} buffer sets of 180 MB (> 126 MB L2)"}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU set NVML reports as local to the GPU (its NUMA node) BEFORE any pinned
    host buffer is allocated (first touch places the pages): with one rank per GPU the host staging of the
    end-to-end path then stays off the inter-socket link. Best effort; returns what was done."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            index = int(vis.split(",")[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to GPU {index}"
    except Exception as exc:  # noqa: BLE001
        return f"unchanged ({type(exc).__name__})"
    return "unchanged"


def time_alloc_section(ctx):
    """BASELINE configs[2]: segment-time allocation driven by batched finite-difference time perturbations,
    4,096 trajectories x 20 iterations. Per iteration: one solve (d_p), the nominal + 2K central perturbed
    costs with d_p held fixed (mtg_cost_time_fd_batch = the loop of NL_I:2495-2584), then the driver-side
    update T <- max(0.1, T - step). Eager, then the whole sweep captured once in a CUDA graph and replayed."""
    import torch

    B, iters, inc = 4096, 20, 0.1
    pos, times = make_workload(B, seed=3)
    p, t0 = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    w_d, w_t, eta = 1.0, 1.0, 0.02

    def run():
        t = t0.clone()
        for _ in range(iters):
            sol = ctx.solve_batch(p, t, want_free=True)
            fd = ctx.cost_time_fd_batch(p, t, sol["free"], inc, central=True)
            g = w_d * fd["grad"] + w_t
            t = torch.clamp(t - eta * g / (g.abs().amax(dim=0, keepdim=True) + 1e-300) * t, min=0.1)
        return t

    run()
    torch.cuda.synchronize()
    n0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        t_final = run()
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / reps
    launches = (ctx.launch_count - n0) // reps
    evals = B * iters * (1 + 2 * K_SEG)
    out = {"workload": f"configs[2]: {B} trajectories x {iters} iterations, central differences, increment {inc}",
           "unit": "cost evaluations/s", "eager": {"value": evals / (eager_ms * 1e-3), "ms_per_iteration": eager_ms / iters},
           "gpu_launches_per_sweep": int(launches)}
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            t_graph = run()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / reps
        out["cuda_graph"] = {"value": evals / (gms * 1e-3), "ms_per_iteration": gms / iters,
                             "bit_identical_to_eager": bool(torch.equal(t_graph, t_final))}
    except Exception as exc:  # noqa: BLE001 - reported, the eager number stands
        out["cuda_graph"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    return out


def extras_section(ctx, peak):
    """The other entry points of the path, each timed on its own (CUDA events, 3 warm-ups, median of 5): the generic
    constraint-pattern solve, the non-linear objective (N1), control points + corridor constraints (N2) and the
    collision potential (N4). Sizes: 65,536 / 4,096 trajectories of the configs[1] shape."""
    import torch

    def med(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2] * 1e-3

    out = {}
    B = BATCH_PER_GPU
    pos, times = make_workload(B, seed=21)
    p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    # generic pattern solve on the canonical mask (what the class API's solveLinear() uses for other patterns)
    h = NCOEF // 2
    mask = np.zeros((K_SEG + 1, h), dtype=np.uint8)
    mask[:, 0] = 1
    mask[0, :] = 1
    mask[-1, :] = 1
    values = torch.zeros((K_SEG + 1, h, DIM, B), dtype=torch.float64, device="cuda")
    values[:, 0] = p
    secs = med(lambda: ctx.solve_generic_batch(mask, values, t))
    out["solve_generic"] = {"value": B / secs, "unit": "trajectories/s", "ms": secs * 1e3,
                            "frac_of_hbm": BYTES_PER_TRAJ * B / secs / 1e9 / peak,
                            "note": "solve_generic_kernel, canonical mask, 65,536 problems"}
    # N1 on 4,096 trajectories (AoS): one objective evaluation = coefficients of d_p + J_d with its analytic
    # gradient + J_sc with its central finite-difference gradient (2 x 108 perturbed 2-segment root problems x 2)
    Bn = 4096
    pa = torch.from_numpy(np.ascontiguousarray(np.moveaxis(pos[..., :Bn], -1, 0))).cuda()
    ta = torch.from_numpy(np.ascontiguousarray(np.moveaxis(times[..., :Bn], -1, 0))).cuda() * 0.7
    sol = ctx.solve_batch(pa, ta, want_free=True, layout="aos")
    ders, lims = [1, 2], [3.0, 5.0]

    def objective():
        c = ctx.set_free_constraints_batch(pa, ta, sol["free"], layout="aos")
        ctx.cost_derivative_batch(pa, ta, sol["free"], layout="aos")
        ctx.soft_constraint_gradient_batch(c["coeffs"], ta, ders, lims, 5.0, 1e12, 0.05)

    secs = med(objective)
    out["nl_objective"] = {"value": Bn / secs, "unit": "objective + gradient evaluations/s", "ms": secs * 1e3,
                           "root_problems_per_evaluation": 2 * (K_SEG + 2 * 2 * DIM * (K_SEG - 1) * (h - 1) * 2),
                           "note": "4,096 trajectories: J_d + analytic gradient, J_sc + central FD gradient (NL_I:1537-1606, 2365-2423)"}
    x = sol["free"].clone()
    iters = 10
    secs = med(lambda: ctx.nl_descent_batch(pa, ta, x, ders, lims, soft_weight=5.0, increment=0.05, step=0.5,
                                            iterations=iters, want_history=False), reps=3, warm=1)
    out["nl_descent"] = {"value": Bn * iters / secs, "unit": "trajectory-iterations/s", "ms_per_iteration": secs * 1e3 / iters}
    # N2 / N4 on the 65,536 batch
    solb = ctx.solve_batch(p, t)
    radii = torch.full((K_SEG, 2, B), 1.0, dtype=torch.float64, device="cuda")
    secs = med(lambda: ctx.control_points_batch(t, coeffs=solb["coeffs"], positions=p, radii=radii))
    out["control_points"] = {"value": B / secs, "unit": "trajectories/s", "ms": secs * 1e3}
    g = torch.full((120, 120, 120), 3.0, dtype=torch.float64, device="cuda")
    secs = med(lambda: ctx.collision_cost_batch(solb["coeffs"], t, g, [-60, -60, -60], 0.2, [-40.0] * 3, [40.0] * 3, 0.1,
                                                epsilon=4.0, robot_radius=0.3))
    out["collision_cost"] = {"value": B / secs, "unit": "trajectories/s", "ms": secs * 1e3,
                             "note": "dt 0.1 s, map resolution 0.2 m, potential active inside the 24 m grid (epsilon 4 m), bounds wide enough "
                                     "that no trajectory ends early in a collision (~620 map checks each), with gradient"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-sweep", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-overlap", action="store_true",
                    help="ordinary launches: consecutive steps do not overlap (mtg_set_solve_overlap off)")
    ap.add_argument("--two-launch-step", action="store_true",
                    help="step = mtg_solve_batch + mtg_argmin_batch (two launches) instead of the fused mtg_solve_argmin_batch")
    ap.add_argument("--sweep-batch", type=int, default=SWEEP_BATCH, help=argparse.SUPPRESS)
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = bind_to_gpu_numa_node(local_rank)   # before torch allocates pinned memory
    import torch
    import torch.distributed as dist

    import mav_tube_trajectory_generation_b200 as m
    from mav_tube_trajectory_generation_b200 import sweep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = m.Context(local_rank)
    comm_ok = None
    if distributed:
        # the product's own communicator (mtg_nccl_init): rank 0's unique id travels over the host channel
        uid = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.nccl_init(uid[0], rank, world)
        comm_ok = True
    B = args.batch

    # ---- synthetic inputs: N_ROTATE distinct sets resident in HBM + one pinned host set
    dev_in, dev_out = [], []
    for i in range(N_ROTATE):
        pos, times = make_workload(B, seed=1000 * rank + i)
        dev_in.append((torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()))
        dev_out.append({"coeffs": torch.empty((K_SEG, DIM, NCOEF, B), dtype=torch.float64, device="cuda"),
                        "cost": torch.empty((B,), dtype=torch.float64, device="cuda"),
                        "status": torch.empty((B,), dtype=torch.int32, device="cuda")})
    host_pos = torch.from_numpy(pos).pin_memory()
    host_times = torch.from_numpy(times).pin_memory()
    host_out = {"coeffs": torch.empty((K_SEG, DIM, NCOEF, B), dtype=torch.float64).pin_memory(),
                "cost": torch.empty((B,), dtype=torch.float64).pin_memory(),
                "status": torch.empty((B,), dtype=torch.int32).pin_memory()}

    best = torch.zeros(2, dtype=torch.int64, device="cuda")          # running device pair {cost, global index}
    best_global = torch.zeros(2, dtype=torch.int64, device="cuda")   # the gathered + folded pair
    start, _ = sweep.shard_range(B * world, rank, world)

    # consecutive steps may overlap (programmatic dependent launch): their inputs are resident, i.e. never written
    # by the stream operation before them, which is what mtg_set_solve_overlap asks for
    ctx.set_solve_overlap(not args.no_overlap)

    # one prepared call per buffer set: a step is then ONE ctypes call (a few microseconds of host time), so that the
    # launch train is paced by the GPU and not by the Python binding (rank 0 also runs the rendezvous store)
    prepared = [ctx.prepare_solve_argmin(dev_in[k][0], dev_in[k][1], N=NCOEF, derivative=DERIV, best=best,
                                         out=dev_out[k]) for k in range(N_ROTATE)]

    def step(i, fresh=False):
        p, t = dev_in[i % N_ROTATE]
        o = dev_out[i % N_ROTATE]
        # candidate index = step * global_batch + position in the global batch
        off = i * B * world + start
        if args.two_launch_step:
            ctx.solve_batch(p, t, N=NCOEF, derivative=DERIV, out=o)
            ctx.argmin_batch(o["cost"], status=o["status"], global_offset=off, best=best, accumulate=not fresh)
        else:
            # the same step in one launch: all outputs written, the argmin folded into the solve kernel's epilogue
            prepared[i % N_ROTATE](off, not fresh)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (CUDA events on the launching stream)
    # Everything with a host cost (NVML, the sampler thread, event objects) is set up on EVERY rank before the
    # barrier, so that no rank enters the timed region late.
    sampler = ClockSampler(local_rank)
    ev0, ev_c, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for i in range(args.warmup):
        step(i, fresh=(i == 0))   # the warm-up also takes the fresh-sweep route once (its reset kernel loads lazily)
    ctx.best_allgather(best, out=best_global)   # warm-up of the collective too (NCCL connects lazily)
    barrier()
    n0 = ctx.launch_count
    # device-side rendezvous: the timed region starts behind a collective on the launch stream, so the GPUs
    # leave it together (the host barrier alone leaves tens of microseconds of rank-arrival skew, which a
    # 20-step region of 1.6 ms would count as collective time)
    ctx.best_allgather(best, out=best_global)
    ev0.record()
    for i in range(args.steps):
        step(i, fresh=(i == 0))
    ev_c.record()
    ctx.best_allgather(best, out=best_global)   # the sweep's only exchange: mtg_best_allgather, 16 B per rank
    ev1.record()
    barrier()
    launches = ctx.launch_count - n0 - 1        # the rendezvous launch is outside the timed region
    ms_total, ms_compute = ev0.elapsed_time(ev1), ev0.elapsed_time(ev_c)
    clocks = sampler.stop()
    best_cost, best_idx = ctx.decode_best(best_global)   # read back AFTER the timed region
    bad = int(sum(int(o["status"].max().item()) for o in dev_out))
    per_rank = torch.tensor([ms_total, ms_compute], dtype=torch.float64, device="cuda")
    if distributed:
        allr = torch.empty((world, 2), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, per_rank)
        # every rank must hold the same global pair
        chk = torch.empty((world, 2), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(chk, best_global)
        comm_ok = bool((chk == chk[0:1]).all().item())
    else:
        allr = per_rank[None, :]
    allr = allr.cpu().numpy()
    ms_total = float(allr[:, 0].max())
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.solve_batch(host_pos, host_times, N=NCOEF, derivative=DERIV, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.solve_batch(host_pos, host_times, N=NCOEF, derivative=DERIV, out=host_out)
        _ = float(host_out["cost"][0])      # the step's result is read on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(e2e_t.item())
    # the candidate-sweep form of the same call: cost-only solves (coeffs = NULL), 344 B in / 12 B out per trajectory
    lean_out = {"cost": host_out["cost"], "status": host_out["status"]}
    ctx.solve_batch(host_pos, host_times, N=NCOEF, derivative=DERIV, out=lean_out, want_coeffs=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.solve_batch(host_pos, host_times, N=NCOEF, derivative=DERIV, out=lean_out, want_coeffs=False)
        _ = float(host_out["cost"][0])
    torch.cuda.synchronize()
    lean_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(lean_t, op=dist.ReduceOp.MAX)
    e2e_lean_value = world * B * e2e_steps / float(lean_t.item())

    if rank == 0:
        peak, peak_src = peaks()
        # the solve kernel's own duration: CUDA events around a train of launches of it alone, same buffers, on the
        # launch stream — as in the timed region (consecutive launches may overlap) and with ordinary launches
        def kernel_train(reps=50):
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for i in range(reps):
                p, t = dev_in[i % N_ROTATE]
                ctx.solve_batch(p, t, N=NCOEF, derivative=DERIV, out=dev_out[i % N_ROTATE])
            k1.record()
            torch.cuda.synchronize()
            return k0.elapsed_time(k1) / reps * 1e-3
        kernel_train(10)
        kernel_s = kernel_train()
        ctx.set_solve_overlap(False)
        kernel_iso_s = kernel_train()
        achieved = BYTES_PER_TRAJ * B / kernel_s / 1e9
        fp64_peak = ctx.probe_fp64_fma(reps=5)       # measured DFMA roof of this GPU (TFLOP/s)
        ref_tflops = FLOP_PER_TRAJ_REF * B / kernel_s / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world), "gpu_launches": int(launches),
            "status_nonzero": bad, "sweep_argmin": {"cost": best_cost, "candidate": best_idx},
            "ranks": {"ms_total": [float(x) for x in allr[:, 0]], "ms_compute": [float(x) for x in allr[:, 1]],
                      "collective_ms": float((allr[:, 0] - allr[:, 1]).max()),
                      "skew_ms": float(allr[:, 1].max() - allr[:, 1].min()),
                      "collective": "mtg_best_allgather (ncclAllGather on the library's communicator + device fold)"
                                    if distributed else "mtg_best_allgather (single rank: device fold only)",
                      "comm_nranks_ok": comm_ok, "host_affinity": numa,
                      "note": "GPUs enter the timed region behind a collective on the launch stream; the pair is "
                              "read back after the region"},
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": BYTES_IN * B,
                    "d2h_bytes_per_step": (BYTES_OUT + 4) * B, "steps": e2e_steps,
                    "note": "mtg_solve_batch(MTG_MEM_HOST) on pinned buffers: chunked H2D/kernel/D2H "
                            "pipeline inside the call; wall clock between device synchronisations"},
            "e2e_cost_only": {"value": e2e_lean_value, "unit": "trajectories/s", "h2d_bytes_per_step": BYTES_IN * B,
                              "d2h_bytes_per_step": 12 * B,
                              "note": "same call with coeffs = NULL (cost + status only): the form a candidate sweep uses"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "kernel": SOLVE_KERNEL_NAME, "kernel_ms": kernel_s * 1e3,
                         "kernel_ms_isolated": kernel_iso_s * 1e3,
                         "frac_isolated": BYTES_PER_TRAJ * B / kernel_iso_s / 1e9 / peak,
                         "kernel_timing": "average launch duration in a train of 50 launches (CUDA events on the launch "
                                          "stream): kernel_ms with programmatic dependent launch as in the timed region "
                                          "(a launch starts on the SMs the previous one's last wave leaves empty), "
                                          "kernel_ms_isolated with ordinary launches (what ncu's serialised replay sees)",
                         "kernel_share_of_step": kernel_s * 1e3 / ms_per_step,
                         "traffic": TRAFFIC_PER_LAUNCH if B == BATCH_PER_GPU else None,
                         "traffic_source": TRAFFIC_SOURCE,
                         "algorithmic_bytes_per_trajectory": BYTES_PER_TRAJ,
                         "fp64": {"peak_tflops": fp64_peak, "peak_source": "measured in this run: mtg_probe_fp64_fma "
                                  "(register-only DFMA chains, CUDA events)",
                                  "achieved_tflops_reference_formulation": ref_tflops,
                                  "frac_reference_formulation": ref_tflops / fp64_peak,
                                  "flop_per_trajectory_reference_formulation": FLOP_PER_TRAJ_REF,
                                  "note": "8.8e4 flop is the reference's dense formulation (SURVEY 8d); the kernel "
                                          "executes ~1.9e4 useful flop per trajectory through the closed forms, so the "
                                          "fp64 pipe utilisation is the ncu figure in profiles/, not this ratio"}},
            "clocks": clocks,
        }
        if world == 1 and not args.no_sweep:
            line["sweep"] = sweep_section(ctx, peak, args.sweep_batch)
            line["time_alloc"] = time_alloc_section(ctx)
            line["extras"] = extras_section(ctx, peak)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = 32768
            rate, secs = cpu_port_rate(sample, threads)
            rate1, secs1 = cpu_port_rate(4096, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "trajectories/s", "cores": threads,
                                    "kind": "port",
                                    "sample": f"{sample} solves of the same workload in {secs:.2f} s wall, "
                                              "OpenMP over the batch (oracle port, dense QR, prints excluded)",
                                    "single_thread": {"value": rate1, "cores": 1,
                                                      "sample": f"4096 solves in {secs1:.2f} s: the reference itself "
                                                                "is single-threaded"}}
        emit(line)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
