#!/usr/bin/env python
"""bench.py -- BioEn optimisation hot path on B200: gradient (f+g) evaluations per second.

Workload (BASELINE.json `metric`): log-weights method, synthetic "generic data" (seed 12345,
sig_exp 0.5, sig_sim 1, uniform w0, G = 0), N = 1e6 structures x M = 1e3 observables per GPU, theta = 10, fp64.
One "step" = one evaluation of the log-posterior AND its gradient at a fresh point (2 passes over yTilde).

  value      steps / device time, yTilde + all vectors resident in HBM (CUDA events, max over ranks)
  e2e        the same through the C ABI with HOST vectors (bioen_b200_eval): every step copies g from pinned
             host memory to the device and reads the objective and the gradient back
  roofline   the dominant kernel (stream_pass_kernel, one pass over yTilde = M*N*8 algorithmic bytes per launch),
             timed live with CUDA events around every launch inside the timed region
  cpu_baseline  the reference's own OpenMP C kernels (oracle/_ref, built from /root/reference by oracle/Makefile)
             on the host cores, on a column sample of the same problem

N > 1 (torchrun): the structure axis is sharded, every rank holds N = 1e6 columns (weak scaling), one exchange
of M+3 doubles (+ tiny scalar reductions) per evaluation, carried by the library's peer-memory kernel over
NVLink (`config.exchange` = "p2p"; "nccl" when peer mapping is unavailable or BIOEN_B200_P2P=0); `value` counts
1e6-structure evaluations per second summed over ranks.

`--impl reference` times the reference CPU implementation alone (no GPU code on that path).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345
SIG_EXP, SIG_SIM = 0.5, 1.0
THETA = 10.0
METRIC = "grad_evals_per_s"


def unit(args):
    return ("evals/s (one evaluation = objective + gradient of the %s method over an N=%d x M=%d block of yTilde; "
            "summed over GPUs)" % (args.method, args.n, args.m))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # long spellings: under torchrun a bare `--m` is swallowed by the launcher's own abbreviation matching
    ap.add_argument("--m", "--observables", dest="m", type=int, default=1000)
    ap.add_argument("--n", "--structures", dest="n", type=int, default=1000000, help="structures PER GPU")
    ap.add_argument("--method", default="logw", choices=["logw", "forces"])
    ap.add_argument("--cpu-cols", type=int, default=200000, help="columns of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --structures is the TOTAL N, split over the GPUs (default: weak, N per GPU)")
    ap.add_argument("--no-optimum", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the host-matrix drop-in call (needs M*N*8 B of host RAM)")
    ap.add_argument("--unfused-forces", action="store_true", help="forces: four tile passes instead of two fused")
    ap.add_argument("--theta-scan", type=int, default=0, metavar="K",
                    help="BASELINE config 4: K theta values (log-spaced 1e3..1e-1) minimised together; prints the "
                         "batched-evaluation line instead of the single-theta one")
    return ap.parse_args()


def observations(M):
    """YTrue / sigma and YTilde of the generic-data recipe (the M-vectors are tiny: host NumPy)."""
    rng = np.random.default_rng(SEED)
    ytrue = rng.standard_normal(M)
    yobs = ytrue + SIG_EXP * rng.standard_normal(M)
    return ytrue / SIG_EXP, yobs / SIG_EXP


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(method, M, N):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed `ncu --set full`
    capture of this same workload (profiles/), or None when the workload differs from the profiled one."""
    if (M, N) != (1000, 1000000):
        return None
    name = "r1_ncu_full_stream_pass.csv" if method == "logw" else "r1_ncu_full_fused_team_pass.csv"
    try:
        import csv
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            rows = list(csv.reader(fh))
        hdr, units = rows[0], rows[1]
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        vals = [float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]] for r in rows[2:] if r]
        return sum(vals) / len(vals)
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons sampled with NVML during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's OpenMP C kernels on the host cores
# ----------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_evals(M, N_full, cols, steps, warmup, method):
    """f+g evaluations/s of the reference C code (oracle/_ref) -- or of the oracle port if _ref is absent --
    on a `cols`-column sample, scaled to N_full columns (cost is linear in N).  Returns (value, info)."""
    from oracle import oracle as O
    from oracle import ref
    cores = host_cores()
    a, YT = observations(M)
    rng = np.random.default_rng(SEED + 1)
    yT = np.empty((M, cols))
    blk = 64
    for i in range(0, M, blk):  # chunked: bounded temporaries
        j = min(M, i + blk)
        yT[i:j] = a[i:j, None] + (SIG_SIM / SIG_EXP) * rng.standard_normal((j - i, cols))
    G = np.zeros(cols)
    w0 = np.full(cols, 1.0 / cols)
    x = 0.1 * rng.standard_normal(cols) if method == "logw" else 1e-3 * rng.standard_normal(M)
    best = None
    if ref.available():
        kind = "reference"
        ref.set_num_threads(cores)
        ref.set_fast_openmp_flag(1)      # the reference's fastest mode
        used = cores
        evs = []
        for caching in (True, False):    # the reference's transposed-cache option: time both, keep the faster
            evs.append((ref.LogwEvaluator(G, yT, YT, THETA, caching=caching) if method == "logw"
                        else ref.ForcesEvaluator(w0, yT, YT, THETA, caching=caching), caching))
    else:
        kind = "port"
        used = 1
        evs = [((lambda v: O.logw_fg(v, G, yT, YT, THETA)) if method == "logw"
                else (lambda v: O.forces_fg(v, w0, yT, YT, THETA)), False)]
    for ev, caching in evs:
        for _ in range(max(1, warmup)):
            ev(x)
        times = []
        for k in range(steps):
            t0 = time.perf_counter()
            ev(x + 1e-9 * k)
            times.append(time.perf_counter() - t0)
        if best is None or min(times) < best[0]:
            best = (min(times), float(np.mean(times)), caching)
    dt_best, dt_mean, caching = best
    per_eval_full = dt_best * (N_full / cols)
    info = {"kind": kind, "cores": used, "ms_per_eval_sample": 1e3 * dt_best,
            "sample": "%d of %d columns x %d rows (%.1f GB), best of %d f+g evaluations (mean %.1f ms, best "
                      "%.1f ms), fast_openmp=1, yTildeT cache %s (faster of on/off); evals/s scaled by %d/%d "
                      "(cost linear in N)" % (cols, N_full, M, M * cols * 8 / 1e9, steps, 1e3 * dt_mean,
                                              1e3 * dt_best, "on" if caching else "off", cols, N_full)}
    return 1.0 / per_eval_full, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    t_wall = time.perf_counter()
    value, info = cpu_reference_evals(args.m, args.n, args.cpu_cols, steps, min(args.warmup, 2), args.method)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": unit(args), "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s f+g evaluation, synthetic generic data N=%d x M=%d, theta=%g"
                               % (args.method, args.n, args.m, THETA), "method": args.method},
        "cpu_baseline": {"value": value, "unit": unit(args), "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": unit(args), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import bioen_b200
    from bioen_b200 import _lib
    from bioen_b200.problem import FORCES, LOGW

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    M = args.m
    if args.strong:
        from bioen_b200.dist import shard_bounds
        lo, hi = shard_bounds(args.n, rank, world)
        N, col0 = hi - lo, lo
    else:
        N, col0 = args.n, rank * args.n
    method = LOGW if args.method == "logw" else FORCES
    nvar = N if method == LOGW else M
    a, YT = observations(M)

    prob = bioen_b200.Problem(shape=(M, N), device=local)
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            raw = ctypes.create_string_buffer(128)
            _lib.check(_lib.load().bioen_b200_nccl_unique_id(raw), "nccl_unique_id")
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        prob.comm_init(bytes(idbuf.cpu().numpy().tobytes()), rank, world, args.n if args.strong else N * world)
    t0 = time.perf_counter()
    prob.generate(SEED, col0, a, SIG_SIM / SIG_EXP)          # this rank's columns of the global matrix
    gen_s = time.perf_counter() - t0
    n_total = args.n if args.strong else N * world
    if method == LOGW:
        prob.set_logw(np.zeros(N), YT, THETA)
    else:
        if args.unfused_forces:
            prob.set_option(1, 0)
        prob.set_forces(np.full(N, 1.0 / n_total), YT, THETA)

    # start point: non-uniform weights (SURVEY 8d parity point); replicated vector for forces
    rng = np.random.default_rng(SEED + 7 + (rank if method == LOGW else 0))
    x_host = torch.empty(nvar, dtype=torch.float64).pin_memory()
    g_host = torch.empty(nvar, dtype=torch.float64).pin_memory()
    x_np, g_np = x_host.numpy(), g_host.numpy()
    x_np[:] = (0.1 if method == LOGW else 1e-3) * rng.standard_normal(nvar)
    x_dev = x_host.to(dev)
    g_dev = torch.zeros_like(x_dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------
    barrier()
    with ClockSampler(local) as clk:
        ms, pass_ms, launches = prob.time_evals(x_dev.data_ptr(), g_dev.data_ptr(), args.warmup, args.steps)
    barrier()
    t = torch.tensor([ms, pass_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, pass_ms = float(t[0]), float(t[1])
    # what a plain read-only stream over the same resident matrix reaches on this device (outside the timed region):
    # the context for a roofline fraction above 1 against the read+write copy figure of MEASURED_PEAKS.json
    read_gbs = ctypes.c_double(0.0)
    _lib.check(_lib.load().bioen_b200_read_stream_peak(prob._h, 10, ctypes.byref(read_gbs)), "read_stream_peak")
    units = 1 if args.strong else world     # weak: every rank adds one N-block per evaluation
    value = units * args.steps / (ms * 1e-3)

    # ---- end to end: host vectors through the C ABI ------------------------------------------------------
    lib = _lib.load()
    f = ctypes.c_double()
    e2e_steps = args.steps
    for k in range(args.warmup):
        _lib.check(lib.bioen_b200_eval(prob._h, method, _lib.ptr(x_np), ctypes.byref(f), _lib.ptr(g_np)), "eval")
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        x_np[0] += 1e-9          # a new point every step
        _lib.check(lib.bioen_b200_eval(prob._h, method, _lib.ptr(x_np), ctypes.byref(f), _lib.ptr(g_np)), "eval")
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units * e2e_steps / float(te[0])

    # ---- time to optimum (device L-BFGS, BioEn defaults) -----------------------------------------------
    optimum = None
    if not args.no_optimum:
        x0 = np.zeros(nvar)
        barrier()
        t0 = time.perf_counter()
        xo, fmin, code, info = prob.opt_lbfgs(x0)
        barrier()
        optimum = {"seconds": time.perf_counter() - t0, "code": code, "fmin": fmin,
                   "iterations": info["iterations"], "evaluations": info["evaluations"],
                   "objective_only_evaluations": info["gradients_skipped"],
                   "minimizer": "device L-BFGS (liblbfgs semantics, BioEn defaults: linesearch=2, past=10, "
                                "delta=1e-6, epsilon=1e-6)", "includes": "x0 H2D + result D2H; yTilde resident"}

    # ---- the reference-facing call with HOST buffers: bioen.optimize.ext.c_bioen.bioen_opt_lbfgs_* ------------
    dropin = None
    if world == 1 and not args.no_dropin and not args.no_optimum:
        from bioen_b200 import optimize
        from bioen_b200.optimize.ext import c_bioen
        yT_host = prob.download()                       # the same matrix, now a pageable NumPy array
        cfg = optimize.minimize.Parameters("lbfgs")
        cfg["verbose"] = False
        t0 = time.perf_counter()
        with bioen_b200.Problem(yT_host, device=local):
            upload_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        if method == LOGW:
            xo2, fmin2 = c_bioen.bioen_opt_lbfgs_logw(np.zeros(N), np.zeros(N), yT_host, YT, THETA, cfg)
        else:
            xo2, fmin2 = c_bioen.bioen_opt_lbfgs_forces(np.zeros(M), np.full(N, 1.0 / N), yT_host, YT, THETA, cfg)
        total_s = time.perf_counter() - t0
        dropin = {"api": "c_bioen.bioen_opt_lbfgs_%s (host NumPy arrays in, result out; = the reference's Cython entry)"
                         % args.method, "seconds": total_s, "ytilde_upload_s": upload_s,
                  "h2d_bytes": M * N * 8 + (N + M) * 8, "fmin": fmin2,
                  "agrees_with_resident_run": bool(optimum and abs(fmin2 - optimum["fmin"]) <= 1e-8 * abs(fmin2))}
        del yT_host

    peak, peak_src = measured_peak()
    alg_bytes = float(M) * N * 8.0
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": unit(args), "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "%s f+g evaluation, synthetic generic data N=%d x M=%d per GPU (yTilde %.1f GB fp64 "
                        "per GPU, N sharded over %d GPU(s)), theta=%g" % (args.method, N, M, alg_bytes / 1e9,
                                                                         world, THETA),
            "method": args.method, "n_per_gpu": N, "m": M, "n_total": n_total,
            "l2": "inputs (%.1f GB) larger than L2 (126 MB); every step evaluates a new point" % (alg_bytes / 1e9),
            "global_evals_per_s": args.steps / (ms * 1e-3),
            "generate_s": gen_s, "exchange": prob.comm_mode(),
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(args.method, M, N),
                     "traffic_source": "profiles/r1_ncu_full_*.csv (ncu --set full of this workload, mean per launch)",
                     "peak_source": peak_src,
                     "kernel": ("stream_pass_kernel" if method == LOGW or args.unfused_forces else "fused_team_pass")
                               + " (one pass over yTilde)", "bytes_per_launch": alg_bytes,
                     "ms_per_launch": pass_ms,
                     "read_only_stream": {"gbs": read_gbs.value, "frac_of_it": achieved / read_gbs.value
                                          if read_gbs.value > 0 else None,
                                          "what": "plain 16-byte-load read kernel over the same yTilde, mean of 10 "
                                                  "launches, measured in this run; `peak` is a copy (read+write) figure"},
                     "step_frac": (2 * M * N * 8.0) / (ms / args.steps * 1e-3) / 1e9 / peak},
        "e2e": {"value": e2e_value, "unit": unit(args), "h2d_bytes_per_step": nvar * 8, "d2h_bytes_per_step": nvar * 8 + 512,
                "api": "bioen_b200_eval (C ABI, pinned host vectors; yTilde resident after one upload)"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if optimum:
        line["time_to_optimum"] = optimum
    if dropin:
        line["dropin_time_to_optimum"] = dropin
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_reference_evals(M, N, args.cpu_cols, 5, 1, args.method)
            line["cpu_baseline"] = {"value": v, "unit": unit(args), "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"]}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": unit(args), "cores": 0, "kind": "unavailable", "sample": str(e)}
    prob.close()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_theta_scan(args):
    """BASELINE.json config 4: the L-curve -- K problems, one yTilde stream per pass for all of them.  With
    N > 1 GPUs the structure axis is sharded (N per GPU, weak scaling) and the per-problem scalars are all-reduced."""
    import torch
    import torch.distributed as dist
    import bioen_b200
    from bioen_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M, N, K = args.m, args.n, args.theta_scan
    n_total = N * world
    a, YT = observations(M)
    prob = bioen_b200.Problem(shape=(M, N), device=local)
    lib = _lib.load()
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            raw = ctypes.create_string_buffer(128)
            _lib.check(lib.bioen_b200_nccl_unique_id(raw), "nccl_unique_id")
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        prob.comm_init(bytes(idbuf.cpu().numpy().tobytes()), rank, world, n_total)
    prob.generate(SEED, rank * N, a, SIG_SIM / SIG_EXP)
    forces = args.method == "forces"
    meth = 1 if forces else 0
    nvar = M if forces else N
    if forces:
        prob.set_forces(np.full(N, 1.0 / n_total), YT, THETA)
    else:
        prob.set_logw(np.zeros(N), YT, THETA)
    thetas = np.geomspace(1e3, 1e-1, K)
    rng = np.random.default_rng(SEED + 7 + (0 if forces else rank))
    X0 = np.ascontiguousarray((1e-3 if forces else 0.1) * rng.standard_normal((K, nvar)))
    peak_tf = ctypes.c_double()
    _lib.check(lib.bioen_b200_dmma_peak(local, ctypes.byref(peak_tf)), "dmma_peak")
    ms, gemm_ms, launches = ctypes.c_float(), ctypes.c_float(), ctypes.c_longlong()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    with ClockSampler(local) as clk:
        _lib.check(lib.bioen_b200_time_scan_evals(prob._h, meth, K, _lib.ptr(thetas), _lib.ptr(X0), args.warmup, args.steps,
                                                  ctypes.byref(ms), ctypes.byref(gemm_ms), ctypes.byref(launches)),
                   "time_scan_evals")
    barrier()
    t = torch.tensor([ms.value, gemm_ms.value], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms, g_ms = float(t[0]), float(t[1])
    KP = (K + 7) // 8 * 8
    flops = 2.0 * M * N * KP
    ach = flops / (g_ms * 1e-3) / 1e12
    hbm_peak, _ = measured_peak()
    line = {
        "metric": "theta_scan_problem_evals_per_s", "value": world * K * args.steps / (tot_ms * 1e-3),
        "unit": "f+g evaluations/s of N=%d x M=%d blocks, summed over K problems and GPUs (%s)" % (N, M, args.method),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "theta L-curve scan (%s method), K=%d theta values batched, N=%d x M=%d per GPU, skinny "
                               "fp64 GEMMs on tensor cores (DMMA)" % (args.method, K, N, M), "K": K,
                   "method": args.method, "n_total": n_total,
                   "l2": "inputs (%.1f GB) larger than L2" % (M * N * 8 / 1e9)},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf.value, "unit": "TFLOP/s",
                     "frac": ach / peak_tf.value, "traffic": None,
                     "peak_source": "measured in this run: register-resident mma.sync.m8n8k4.f64 loop (no fp64 figure "
                                    "in MEASURED_PEAKS.json)",
                     "kernel": "batched_gemm_kernel (one pass over yTilde for all K)", "flops_per_launch": flops,
                     "ms_per_launch": g_ms,
                     "hbm_frac_same_launch": (M * N * 8.0) / (g_ms * 1e-3) / 1e9 / hbm_peak},
        "gpu_launches": int(launches.value), "clocks": clk.summary(),
    }
    if not args.no_optimum:
        barrier()
        t0 = time.perf_counter()
        X, fmin, codes, info = prob.theta_scan(thetas, x0=np.zeros(nvar), method=meth)
        barrier()
        line["time_to_optimum"] = {"seconds": time.perf_counter() - t0, "rounds": info["rounds"],
                                   "codes": [int(c) for c in codes], "iterations": [int(i) for i in info["iterations"]],
                                   "evaluations_total": int(info["evaluations"].sum()),
                                   "fmin_first_last": [float(fmin[0]), float(fmin[-1])],
                                   "includes": "K x n start vectors H2D, results D2H; yTilde resident"}
    prob.close()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, liblbfgs-style progress prints) goes to stderr;
    the one JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    _claim_stdout()
    if args.theta_scan and args.impl != "reference":
        return run_theta_scan(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
