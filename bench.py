#!/usr/bin/env python
"""bench.py -- grid-point*solves/s of the Dirichlet Poisson hot path on B200.

One "step" = one complete Dirichlet Poisson problem with an immersed circle
(test/literate/dirichlet.jl:71-107): the DDF tables of the body, D_s d, L^-1, the
Schur build S = -E L^-1 R (N column solves), LU, the surface-point solve, R s,
L^-1 -- N + 2 inverse Laplacians on an NX x NY grid.
Metric (BASELINE.json): NX*NY*(N+2) / seconds.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 4096] [--impl reference]

* `value`   : boundary data resident in HBM (torch CUDA tensors), one call of
              ilm_dirichlet_poisson per step, CUDA-event timing.
* `e2e`     : the same call with HOST (numpy) buffers: surface points and boundary
              data go H2D, the field and the multiplier come back D2H, every step.
* `roofline`: the dominant kernel of the step timed alone with CUDA events on the
              plan's stream (ilm_profile_conv_probe), beside the other passes, the
              dense (general right-hand side) inverse Laplacian and the stencil /
              regularize / interpolate stages (L2 flushed between repetitions).
* `cpu_baseline` / `--impl reference`: the CPU oracle restatement of the
  reference path (the Julia reference cannot run here) on the host cores.
N > 1 (torchrun): the N Schur columns are sharded over the ranks inside the
library (NCCL communicator bound to the plan, one grouped in-place broadcast);
LU, the solve and the two remaining L^-1 are replicated (strong scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1, which scipy.fft's worker pool honours: the CPU arm would run on ONE
    # core under `torch.distributed.run` (round 1: 0.67 s -> 7.5 s per probe).  The reference arm uses every core
    # the process may run on, however it was launched.
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point-solves/s (Dirichlet Poisson with immersed body, incl. Schur build and solve)"
UNIT = "grid-point*solves/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures at 4096^2 (profiles/r2_*_ncu_*.txt)
NCU_TRAFFIC = {"k_passD": 274.3e6, "passC_probe": 276.0e6,
               # stencils: reads equal the algorithmic reads exactly; the write part is a lower bound (lines still in L2)
               "divergence": 371.9e6, "grad": 343.6e6, "curl_nodes_to_edges": 344.9e6, "curl_edges_to_nodes": 371.7e6,
               "laplacian": 224.5e6, "regularize": 72.9e6, "interpolate": 1.4e6}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Start of the timed region: only samples that arrive after this call are kept.  The process is started
        before the warm-up steps because nvidia-smi's own start-up (NVML initialisation of every GPU of the box) contends
        for the driver while it lasts: started right before the timed region it cost 22 ms per step on two GPUs."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        rows = [r for t_row, r in self.rows if t_row >= t_mark]
        if not rows and self.rows:                       # a timed region shorter than the sampling period: the nearest sample
            rows = [self.rows[-1][1]]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def problem(grid_n):
    import ilm_b200 as ilm
    g = ilm.PhysicalGrid.centered(grid_n)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(grid_n, cache_dir="/tmp/ilm_lgf_cache")
    return g, body, G


def workload_config(args, g, N, world):
    """The `config` of both arms (same workload, same accounting of solves)."""
    return {"workload": f"dirichlet_circle_{args.grid}", "grid": [g.NX, g.NY], "surface_points": int(N),
            "solves_per_step": int(N + 2), "ddf": "yang3"}


# ----------------------------------------------------------------------------- reference arm (CPU oracle)
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_cache(g, body, G):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ilm_oracle as o
    return o, o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body[:5], G, workers=host_threads())


def cpu_probe(oc, col):
    """One column probe of create_RTLinvR on the CPU path (one L^-1)."""
    e = np.zeros(oc.N)
    e[col] = 1.0
    return oc.interpolate(oc.inverse_laplacian(oc.regularize(e)))


def run_reference(args, rank):
    """The reference's own algorithm on the host cores: a step is a bounded SAMPLE of the workload -- `cols` column
    probes (R e_c -> L^-1 -> E, one inverse Laplacian each) of the N + 2 solves of a problem; the rate
    (grid-point*solves/s) does not depend on how many of the identical solves are timed.  At least 32 columns are
    timed over the K steps."""
    if rank != 0:
        return
    g, body, G = problem(args.grid)
    o, oc = oracle_cache(g, body, G)
    steps = max(args.steps, 1)
    cols = max(1, -(-32 // steps))
    c = 0
    for w in range(args.warmup):
        cpu_probe(oc, c % oc.N)
        c += 1
    t0 = time.perf_counter()
    for k in range(steps):
        for _ in range(cols):
            cpu_probe(oc, c % oc.N)
            c += 1
    dt_step = (time.perf_counter() - t0) / steps
    dt = dt_step / cols
    val = g.NX * g.NY * 1.0 / dt
    cores = host_threads()
    sample = (f"{cols} Schur column probes (R e_c -> L^-1 -> E = 1 inverse Laplacian each) per step, {cols * steps} in all, "
              f"{dt:.3f} s each on {cores} threads; scipy.fft on the (2NX-1)^2 pad; value = rate of these solves "
              f"(a full step is {oc.N + 2} of them: {dt * (oc.N + 2):.0f} s)")
    cfg = workload_config(args, g, oc.N, 1)
    cfg["note"] = "CPU oracle restatement of the reference path (Julia reference not runnable here); bounded sample"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt_step * 1e3, "ms_per_full_problem_extrapolated": dt * (oc.N + 2) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-stage roofline section")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # Libraries (NCCL prints its version banner on stdout) must not pollute the one JSON line:
    # everything written to fd 1 goes to stderr until the final print.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import ilm_b200 as ilm
    from ilm_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    g, body, G = problem(args.grid)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
    cache.sync()
    plan_build_ms = (time.perf_counter() - t0) * 1e3            # once per grid: LGF upload, Ghat, Gx, scratch
    if world > 1:
        cache.comm_init()
    N = cache.N
    n_solves = N + 2
    fplus = cache.points()[0].copy()
    fplus_dev = torch.from_numpy(fplus).cuda()
    peak, peak_src = load_peaks()
    lib = cache._lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def launches():
        return cache.launch_count() + int(lib.ilm_dense_launch_count())

    def step_device():
        cache.update_points(body)                        # the problem's DDF tables (what a new / moved body costs)
        return ilm.dirichlet_solve(cache, fplus_dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    lc1 = launches()
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        f, s = step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    lc2 = launches()
    ms = max_over_ranks(e0.elapsed_time(e1) / max(args.steps, 1))
    value = g.NX * g.NY * n_solves / (ms * 1e-3)
    checksum = float(f.data.sum().item())

    def time_dev(fn, reps=3):
        fn()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        a1.synchronize()
        return max_over_ranks(a0.elapsed_time(a1) / reps)

    # ---- where the step goes (each piece timed alone, max over ranks)
    S_dev = ilm.create_schur_sharded(cache, "RTLinvR")
    rhs = torch.randn(N, dtype=torch.float64, device="cuda")
    breakdown = {
        "table_refresh_ms": time_dev(lambda: (cache.update_points(body), cache.sync())),
        "schur_build_ms": time_dev(lambda: ilm.create_schur_sharded(cache, "RTLinvR"), reps=2),
        "lu_factor_ms": time_dev(lambda: ilm.LU(S_dev)),
        "lu_solve_ms": None, "plan_build_ms_once_per_grid": plan_build_ms}
    lu = ilm.LU(S_dev)
    breakdown["lu_solve_ms"] = time_dev(lambda: lu.solve(rhs), reps=5)
    breakdown["note"] = ("schur_build is the sharded part (N column probes / n_gpus in blocks of equal row counts + one in-place broadcast group); "
                         "table refresh, LU, solve and the two dense L^-1 are replicated on every rank")

    # ---- convolution passes timed alone on the plan's stream (CUDA events inside the library)
    msp = (L.C.c_double * 3)()
    L.check(lib.ilm_profile_conv(cache._plan, L.NODES_PRIMAL, 10, L.C.byref(msp)))
    Lx = Ly = 16
    while 2 * Lx < 2 * g.NX - 1:
        Lx *= 2
    while 2 * Ly < 2 * g.NY - 1:
        Ly *= 2
    MYp = (g.NY - 1 + 1) & ~1
    spec = 2 * Lx * MYp * 16
    bytes_pass = {"A_rows_fwd": 2 * (g.NX - 1) * (g.NY - 1) * 8 + spec,
                  "B_columns": 2 * spec + (Lx + 1) * 2 * Ly * 8,
                  "C_rows_inv": spec + 2 * (g.NX - 1) * (g.NY - 1) * 8}
    passes = {k: {"ms": float(msp[i]), "GBps": b / (float(msp[i]) * 1e-3) / 1e9, "bytes": b,
                  "frac_of_hbm_peak": b / (float(msp[i]) * 1e-3) / 1e9 / peak}
              for i, (k, b) in enumerate(bytes_pass.items())}
    conv_ms = sum(float(msp[i]) for i in range(3))
    # the launches the Schur build issues (N of the N + 2 solves of a step): the right-hand side R e_c is a W x W
    # patch.  The band pass (ilm_band.cu) sums the x-spectrum of its <= 8 rows from the DDF windows and forms the
    # y-convolution by direct summation with the x-transformed kernel rows, pass C inverts the rows under the
    # interpolation windows and interpolates them on the fly; one post-sum launch per 32 pairs.
    # (ILM_PROBE_PATCH=0: pre-operator on the grid + pass A on the patch rows, as in the first round-2 builds.)
    msq = (L.C.c_double * 3)()
    L.check(lib.ilm_profile_conv_probe(cache._plan, N // 3, 20, L.C.byref(msq)))
    r0, r1 = L.C.c_int(), L.C.c_int()
    L.check(lib.ilm_probe_output_rows(cache._plan, L.RTLINVR, L.C.byref(r0), L.C.byref(r1)))
    rows_out = r1.value - r0.value
    nrows_in = 8
    W = 5
    patch_mode = float(msq[0]) == 0.0      # the band pass forms the x-spectrum of the DDF windows itself: no pass A
    kernels = {
        "k_passD (band pass: Y_n = sum_r x_r Gx(|n-r|), row-major S2 out" + ("; x_r summed from the DDF windows)" if patch_mode else ")"): {
            "ms": float(msq[1]),
            "bytes": rows_out * 2 * Lx * 16 + (rows_out + nrows_in) * (Lx + 1) * 8 + (0 if patch_mode else nrows_in * 2 * Lx * 16),
            # dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture at 4096^2 (profiles/)
            "traffic": NCU_TRAFFIC.get("k_passD") if args.grid == 4096 else None},
        f"ilm_passC_L{Lx} (probe mode: inverse FFT_x of the window rows + fused interpolation)": {
            "ms": float(msq[2]), "bytes": rows_out * 2 * Lx * 16 + N * W * (16 + 8 * W + 4),
            "traffic": NCU_TRAFFIC.get("passC_probe") if args.grid == 4096 else None},
    }
    if not patch_mode:
        kernels[f"ilm_passA_L{Lx} (probe mode: forward FFT_x of the patch rows)"] = {
            "ms": float(msq[0]), "bytes": nrows_in * ((g.NX - 1) * 16 + 2 * Lx * 16), "traffic": None}
    for k in kernels.values():
        k["GBps"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        k["frac"] = k["GBps"] / peak
    pairs_per_step = (N + 1) // 2 / world
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    dk = kernels[dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dk["GBps"], "peak": peak, "unit": "GB/s", "frac": dk["frac"],
                "traffic": dk["traffic"], "peak_source": peak_src, "algorithmic_bytes_per_launch": dk["bytes"],
                "launch_ms": dk["ms"], "launches_per_step": pairs_per_step,
                # the launch timed here inverts ALL window rows; in the step a pair inverts the rows from its own windows
                # upwards (symmetric build), about half on average: share = S-build share x the kernel's share of a pair
                "share_of_step": (breakdown["schur_build_ms"] / ms) * dk["ms"] / sum(k["ms"] for k in kernels.values()),
                "rows_inverted_per_pair_vs_all_window_rows": breakdown["schur_build_ms"] / (pairs_per_step * sum(k["ms"] for k in kernels.values())),
                "frac_of_nominal_8TBps": dk["GBps"] / 8000.0,
                "probe_kernels": kernels, "probe_pair_ms": sum(float(msq[i]) for i in range(3)),
                "probe_output_rows": [r0.value, r1.value],
                "dense_passes": passes, "dense_solve_pair_ms": conv_ms,
                "dense_solve_frac_of_hbm": sum(bytes_pass.values()) / (conv_ms * 1e-3) / 1e9 / peak,
                "note": ("probe path: band pass and pass C stream the output rows of the x-spectrum once each; the "
                         "FFT passes are bound by FP64 issue + shared-memory wavefronts (DESIGN.md section 4), "
                         "the band pass by HBM")}

    # ---- the other stages of the path (stencils, regularize, interpolate), device resident: one launch per
    #      repetition between CUDA events on the plan's stream (= torch's default stream), L2 flushed before each
    if not args.no_stages:
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")       # 512 MB > 126 MB L2

        def time_op(fn, reps=10):
            fn()
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                fn()
                a1.record()
                a1.synchronize()
                tot += a0.elapsed_time(a1)
            return tot / reps

        Pn = (g.NX - 1) * (g.NY - 1)
        Pd, Pe = g.NX * g.NY, g.NX * (g.NY - 1) + (g.NX - 1) * g.NY
        gen = torch.Generator(device="cuda").manual_seed(0)
        qe, pn_, sd = cache.zeros_gridgrad(), cache.zeros_grid(), cache.zeros_gridcurl()
        qe.data.normal_(generator=gen); pn_.data.normal_(generator=gen); sd.data.normal_(generator=gen)
        oq, op_, os_ = cache.zeros_gridgrad(), cache.zeros_grid(), cache.zeros_gridcurl()
        fs, fo = cache.zeros_surface(), cache.zeros_surface()
        fs.data.normal_(generator=gen)
        stage_defs = {
            "divergence": (lambda: ilm.divergence(op_, qe, cache), 8 * (Pe + Pn)),
            "grad": (lambda: ilm.grad(oq, pn_, cache), 8 * (Pe + Pn)),
            "curl_nodes_to_edges": (lambda: ilm.curl(oq, sd, cache), 8 * (Pe + Pd)),
            "curl_edges_to_nodes": (lambda: ilm.curl(os_, qe, cache), 8 * (Pe + Pd)),
            "laplacian": (lambda: ilm.laplacian(os_, sd, cache), 16 * Pd),
            "regularize": (lambda: ilm.regularize(op_, fs, cache), 8 * Pn + N * (16 * 12 + 8)),
            "interpolate": (lambda: ilm.interpolate(fo, pn_, cache), N * (16 * 20 + 8)),
        }
        stages = {}
        for name, (fn, nbytes) in stage_defs.items():
            tms = time_op(fn)
            stages[name] = {"ms": tms, "bytes": nbytes, "GBps": nbytes / (tms * 1e-3) / 1e9,
                            "frac_of_hbm_peak": nbytes / (tms * 1e-3) / 1e9 / peak,
                            "traffic": NCU_TRAFFIC.get(name) if args.grid == 4096 else None}
        roofline["stages"] = stages
        roofline["stages_note"] = "one launch per repetition, 512 MB written between repetitions (L2 flush), CUDA events"
        del flush

    # ---- end to end: the same call with host buffers (points + boundary data H2D, field + multiplier D2H)
    hcache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=False)
    if world > 1:
        hcache.comm_init()

    my_rows = g.NY - 1
    slab = (rank * my_rows // world, (rank + 1) * my_rows // world)

    def step_host():
        hcache.update_points(body)
        if world == 1:
            return ilm.dirichlet_solve(hcache, fplus)
        # sharded solve: every rank returns its slab of rows of the field (a row-distributed result), the multiplier everywhere
        return ilm.dirichlet_solve(hcache, fplus, field_rows=slab)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fh, sh = step_host()
    barrier()
    dt = max_over_ranks((time.perf_counter() - t0) / max(args.steps, 1))
    P = (g.NX - 1) * (g.NY - 1) * 8
    if world == 1:
        diff = float(np.abs(fh.data - f.numpy()).max())
    else:
        mxp = g.NX - 1
        diff = max_over_ranks(float(np.abs(np.asarray(fh) - f.numpy()[slab[0] * mxp:slab[1] * mxp]).max()))
    e2e = {"value": g.NX * g.NY * n_solves / dt, "unit": UNIT, "h2d_bytes_per_step": int(world * (5 * N * 8 + N * 8)),
           "d2h_bytes_per_step": int(P + world * N * 8), "ms_per_step": dt * 1e3,
           "max_abs_diff_vs_device_path": diff,
           "api": ("ilm_plan_update_points + ilm_dirichlet_poisson with host pointers" if world == 1 else
                   "ilm_plan_update_points + ilm_dirichlet_poisson_rows with host pointers (every rank returns its slab of "
                   "rows of the field and the multiplier; bytes summed over the ranks)")}
    hcache.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        o, oc = oracle_cache(g, body, G)
        cpu_probe(oc, 0)
        t0 = time.perf_counter()
        ns = 0
        while ns < 32 and (ns < 4 or time.perf_counter() - t0 < 12.0):
            cpu_probe(oc, 1 + ns)
            ns += 1
        dtc = (time.perf_counter() - t0) / ns
        cpu = {"value": g.NX * g.NY / dtc, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"{ns} Schur column probes (R e_c -> L^-1 -> E) = {ns} inverse Laplacians of the {N + 2} "
                         f"in a step, {dtc:.2f} s each; scipy.fft on the (2NX-1)^2 pad with workers = all host threads"}

    if rank == 0:
        cfg = workload_config(args, g, N, world)
        cfg.update({"parallelism": f"schur-columns/{world}",
                    "l2": "a column pair over all window rows streams 2 x %.0f MB of x-spectrum rows (> 126 MB L2)" % (rows_out * 2 * Lx * 16 / 1e6),
                    "schur": ("symmetric build: the points are visited by first window row, the probe of a pair inverts the rows from "
                              "its own windows upwards, S[k,c] below comes from S[c,k] wgt_c/wgt_k, two pairs in flight on two streams "
                              "(ILM_SCHUR_SYMM=0: every column over all window rows, 347 ms instead of 176 ms at N = 4593)"),
                    "checksum_f": checksum})
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(lc2 - lc1), "breakdown": breakdown,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        cache.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
