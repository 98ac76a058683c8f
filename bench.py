#!/usr/bin/env python
"""bench.py -- grid-point*solves/s of the Dirichlet Poisson hot path on B200.

One "step" = one complete Dirichlet Poisson problem with an immersed circle
(test/literate/dirichlet.jl:71-107): D_s d, L^-1, the Schur build S = -E L^-1 R
(N column solves), LU, the surface-point solve, R s, L^-1 -- N + 2 inverse
Laplacians on an NX x NY grid.  Metric (BASELINE.json): NX*NY*(N+2) / seconds.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 4096] [--impl reference]

* `value`   : device-resident inputs (torch CUDA tensors), CUDA-event timing.
* `e2e`     : the same step through the public API with HOST (numpy) buffers:
              every operator call stages its inputs H2D and its result D2H.
* `roofline`: the dominant kernel (column pass B of the FFT convolution) timed
              alone with CUDA events on its own stream (ilm_profile_conv).
* `cpu_baseline` / `--impl reference`: the CPU oracle restatement of the
  reference path (the Julia reference cannot run here) on the host cores.
N > 1 (torchrun): the N Schur columns are sharded over the ranks and exchanged
with one NCCL all-gather; everything else is replicated (strong scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point-solves/s (Dirichlet Poisson with immersed body, incl. Schur build and solve)"
UNIT = "grid-point*solves/s"


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
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

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
        for r in self.rows:
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


# ----------------------------------------------------------------------------- reference arm (CPU oracle)
def oracle_cache(g, body, G):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ilm_oracle as o
    return o, o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body[:5], G, workers=os.cpu_count() or 1)


def cpu_probe(oc, col):
    """One column probe of create_RTLinvR on the CPU path (one L^-1)."""
    e = np.zeros(oc.N)
    e[col] = 1.0
    return oc.interpolate(oc.inverse_laplacian(oc.regularize(e)))


def run_reference(args, rank):
    if rank != 0:
        return
    g, body, G = problem(args.grid)
    o, oc = oracle_cache(g, body, G)
    for w in range(args.warmup):
        cpu_probe(oc, w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_probe(oc, args.warmup + k)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = g.NX * g.NY * 1.0 / dt
    cores = os.cpu_count() or 1
    sample = "1 Schur column probe (R e_c -> L^-1 -> E) per step = 1 inverse Laplacian; scipy.fft (2NX-1)^2 pad"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dirichlet_circle_{args.grid}", "grid": [g.NX, g.NY], "surface_points": int(oc.N),
                   "solves_per_step": 1, "note": "CPU oracle restatement of the reference path (Julia reference not runnable here)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from ilm_b200 import shard, _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    g, body, G = problem(args.grid)
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
    N = cache.N
    n_solves = N + 2
    fplus = cache.points()[0].copy()
    fb_dev = torch.from_numpy(0.5 * fplus).cuda()
    d = cache.zeros_surface().set(fplus)
    ranges = shard.column_ranges(N, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        fstar = cache.zeros_grid()
        ilm.surface_divergence(fstar, d, cache)
        ilm.inverse_laplacian(fstar, cache)
        if world > 1:
            S = shard.create_schur_sharded(ilm.create_RTLinvR, cache)
        else:
            S = ilm.create_RTLinvR(cache)
        s = cache.zeros_surface()
        ilm.interpolate(s, fstar, cache)
        rhs = fb_dev - s.data
        sol = ilm.LU(S).solve(rhs)
        s.data.copy_(-sol)
        f = cache.zeros_grid()
        ilm.regularize(f, s, cache)
        ilm.inverse_laplacian(f, cache)
        f.data.add_(fstar.data)
        return f, s

    lc0 = cache.launch_count() + int(L.load().ilm_dense_launch_count())
    for _ in range(args.warmup):
        step_device()
    lc1 = cache.launch_count() + int(L.load().ilm_dense_launch_count())
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        f, s = step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    lc2 = cache.launch_count() + int(L.load().ilm_dense_launch_count())
    ms = e0.elapsed_time(e1) / max(args.steps, 1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = g.NX * g.NY * n_solves / (ms * 1e-3)
    checksum = float(f.data.sum().item())

    # ---- roofline of the dominant kernel (pass B), timed alone on its stream
    msp = (L.C.c_double * 3)()
    L.check(cache._lib.ilm_profile_conv(cache._plan, L.NODES_PRIMAL, 10, L.C.byref(msp)))
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
    peak, peak_src = load_peaks()
    passes = {k: {"ms": float(msp[i]), "GBps": b / (float(msp[i]) * 1e-3) / 1e9, "bytes": b}
              for i, (k, b) in enumerate(bytes_pass.items())}
    conv_ms = sum(float(msp[i]) for i in range(3))
    # the launches the Schur build actually issues (4593 of the 4595 solves of a step): sparse-row probes.
    # Pass B then reads no spectrum (the input rows are summed directly) -> Ghat in, S2 out
    msq = (L.C.c_double * 3)()
    L.check(cache._lib.ilm_profile_conv_probe(cache._plan, N // 3, 10, L.C.byref(msq)))
    # ... and only for the rows under the interpolation windows (ilm_probe_output_rows)
    r0, r1 = L.C.c_int(), L.C.c_int()
    L.check(cache._lib.ilm_probe_output_rows(cache._plan, L.RTLINVR, L.C.byref(r0), L.C.byref(r1)))
    rows_out = r1.value - r0.value
    bytes_probe_B = 2 * Lx * rows_out * 16 + (Lx + 1) * 2 * Ly * 8
    achieved = bytes_probe_B / (float(msq[1]) * 1e-3) / 1e9
    # FP64 pipe: warp instructions per launch (ncu-verified static counts: inverse FFT 782/thread,
    # sparse forward ~300/thread) x 2 issue cycles / (148 SMs x 4 SMSPs x elapsed cycles at 1.965 GHz)
    fp64_winst_B = 2 * Lx * 2 * (782 + 300) * 8
    roofline = {"bound": "hbm",
                "kernel": f"ilm_passB_L{Ly} in Schur-probe mode (column pass: sparse forward DFT_y * Ghat * IFFT_y, 2 columns of S per launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one ncu capture at 4096^2
                # (profiles/r1_passB_probe_dram_v6.txt: 286.6 MB read + 233.7 MB written); null for other grids
                "traffic": 520.3e6 if args.grid == 4096 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_probe_B,
                "launch_ms": float(msq[1]),
                "dense_equivalent_frac": bytes_pass["B_columns"] / (float(msq[1]) * 1e-3) / 1e9 / peak,
                "fp64_pipe_frac_est": fp64_winst_B * 2 / (148 * 4 * float(msq[1]) * 1e-3 * 1.965e9),
                "probe_passes_ms": {"A_rows_fwd": float(msq[0]), "B_columns": float(msq[1]), "C_rows_inv": float(msq[2])},
                "probe_pair_ms": sum(float(msq[i]) for i in range(3)),
                "probe_output_rows": [r0.value, r1.value],
                "dense_passes": passes,
                "dense_solve_pair_ms": conv_ms,
                "dense_solve_frac_of_hbm": sum(bytes_pass.values()) / (conv_ms * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "note": "the convolution is FP64-issue / shared-memory bound on B200 (DESIGN.md section 4), not HBM bound"}

    # ---- the other stages of the path (stencils, regularize, interpolate), device resident,
    #      20 back-to-back launches between CUDA events on the plan's stream (= torch's default stream)
    def time_op(fn, reps=20):
        fn()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        a1.synchronize()
        return a0.elapsed_time(a1) / reps

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
                        "frac_of_hbm_peak": nbytes / (tms * 1e-3) / 1e9 / peak}
    roofline["stages"] = stages
    # transform-free direct-table Schur builder: reported beside, never part of `value` / `e2e`
    t_direct = time_op(lambda: ilm.create_RTLinvR_direct(cache), reps=3)
    S_fft = ilm.create_RTLinvR(cache, cols=(0, 64))
    S_dir = ilm.create_RTLinvR_direct(cache, cols=(0, 64))
    roofline["schur_direct_table"] = {
        "ms": t_direct, "note": "optional O(N^2 W^4) table form of S (SURVEY fact 8); not used by the metric path",
        "rel_diff_vs_column_solves": float(((S_dir - S_fft).abs().max() / S_fft.abs().max()).item())}

    # ---- end to end through the public API with host buffers
    e2e = None
    if rank == 0 or world > 1:
        hcache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=False)

        def step_host():
            if world > 1:
                # host buffers in and out; the shard exchange itself stays on the device, the gathered S comes
                # back to a page-locked host matrix (what a host caller of create_RTLinvR would hold)
                Sd = shard.create_schur_sharded(ilm.create_RTLinvR, cache)
                hb = ilm.api._host_zeros(N * N)
                torch.from_numpy(hb).copy_(Sd.t().reshape(-1))
                return ilm.dirichlet_poisson(hcache, fplus, S=hb.reshape((N, N), order="F"))[:2]
            return ilm.dirichlet_poisson(hcache, fplus)[:2]

        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fh, sh = step_host()
        barrier()
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        P = (g.NX - 1) * (g.NY - 1) * 8
        # per call: surface_divergence (N in, P out), L^-1 (P, P), create_RTLinvR (N^2 out), interpolate (P in, N out),
        # LU (N^2 in; the factors stay on the device), solve (N, N), regularize (N in, P out), L^-1 (P, P)
        h2d = N * 8 + P + P + N * N * 8 + N * 8 + N * 8 + P
        d2h = P + P + N * N * 8 + N * 8 + N * 8 + P + P
        e2e = {"value": g.NX * g.NY * n_solves / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3,
               "max_abs_diff_vs_device_path": float(np.abs(fh.data - f.numpy()).max())}
        hcache.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        o, oc = oracle_cache(g, body, G)
        cpu_probe(oc, 0)
        ns = 3
        t0 = time.perf_counter()
        for k in range(ns):
            cpu_probe(oc, 1 + k)
        dtc = (time.perf_counter() - t0) / ns
        cpu = {"value": g.NX * g.NY / dtc, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{ns} Schur column probes (R e_c -> L^-1 -> E) = {ns} inverse Laplacians of the {N + 2} "
                         f"in a step, {dtc:.2f} s each; scipy.fft on the (2NX-1)^2 pad with workers=all cores"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"dirichlet_circle_{args.grid}", "grid": [g.NX, g.NY], "surface_points": int(N),
                       "solves_per_step": int(n_solves), "ddf": "yang3", "parallelism": f"schur-columns/{world}",
                       "l2": "spectrum buffers (2 x %.0f MB) exceed the 126 MB L2" % (spec / 1e6),
                       "checksum_f": checksum},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(lc2 - lc1), "roofline": roofline, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
