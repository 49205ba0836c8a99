#!/usr/bin/env python
"""bench.py -- headline benchmark of the periodic D2Q9 hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): MLUPS = lattice updates / microsecond (app/main_taylor_green.f90:122) of the
D2Q9 stream+collide step, whole job over all N GPUs, plus achieved HBM GB/s against the
measured peak (144 B per lattice update in fp64 / 72 B in fp32: 9 reads + 9 writes).

Default workload = BASELINE.json configs[4], the configuration the metric is quoted on at
1/2/4/8 GPUs: D2Q9 BGK fp64, slab decomposition along the slow index, WEAK scaling with a
32768 (unit-stride) x 4096 (slow) slab per GPU -- at N = 8 this is the full 32768 x 32768 grid.
One "step" = one fused stream+collide pass over the whole grid.  The other BASELINE configs are
selectable with --workload (C1 64^2 BGK, C2 1024^2 TRT, C3 8192^2 RR fp64/fp32, C4 2048^2 DUGKS).

For N > 1 the driver launches this file with torch.distributed.run, one rank per GPU; ranks
exchange two halo lines per direction per launch (inside libplbm_b200.so: peer stores over NVLink
through CUDA IPC mappings, or NCCL send/recv as fallback), overlapped with the interior update.

`--impl reference` times the reference's own CPU algorithm (the line-faithful C/OpenMP restatement
in oracle/, because the Fortran reference cannot be compiled in this image -- no gfortran) with all
host threads on a bounded SAMPLE of the same workload (stated in cpu_baseline.sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# name: ny (unit stride), nx per GPU (negative: global, split over the ranks), scheme, collision, precision
WORKLOADS = {
    "c5_bgk_f64_slab": dict(ny=32768, nxl=4096, scheme="lbm", collision="bgk", precision="f64",
                            desc="C5 D2Q9 BGK fp64 Taylor-Green, y-slab weak scaling, 32768 x 4096 lines per GPU"),
    # strong scaling of the full C5 grid: nx per GPU = 32768 / N (154.6 GB of PDFs on one GPU)
    "c5_bgk_f64_strong": dict(ny=32768, nxl=-32768, scheme="lbm", collision="bgk", precision="f64",
                              desc="C5 D2Q9 BGK fp64 Taylor-Green 32768 x 32768, y-slab STRONG scaling"),
    "c3_rr_f64_8192": dict(ny=8192, nxl=8192, scheme="lbm", collision="rr", precision="f64",
                           desc="C3 D2Q9 recursive-regularized fp64 Taylor-Green 8192 x 8192 per GPU"),
    "c3_rr_f32_8192": dict(ny=8192, nxl=8192, scheme="lbm", collision="rr", precision="f32",
                           desc="C3 D2Q9 recursive-regularized fp32 Taylor-Green 8192 x 8192 per GPU"),
    "c2_trt_f64_1024": dict(ny=1024, nxl=1024, scheme="lbm", collision="trt", precision="f64",
                            desc="C2 D2Q9 TRT fp64 1024 x 1024 per GPU"),
    "c1_bgk_f64_64": dict(ny=64, nxl=64, scheme="lbm", collision="bgk", precision="f64",
                          desc="C1 D2Q9 BGK fp64 Taylor-Green 64 x 64 (L2-resident, launch-bound)"),
    # C4: DUGKS (src/periodic_dugks.F90:25-38, built with -DDUGKS), dt = min(5 tau, CFL 0.5)
    "c4_dugks_f64_2048": dict(ny=2048, nxl=2048, scheme="dugks", collision="bgk", precision="f64",
                              desc="C4 DUGKS Taylor-Green fp64 2048 x 2048 per GPU (perform_dugks_step, -DDUGKS branch), dt = min(5 tau, 0.5)"),
    "c4_dugks_f32_2048": dict(ny=2048, nxl=2048, scheme="dugks", collision="bgk", precision="f32",
                              desc="C4 DUGKS Taylor-Green fp32 2048 x 2048 per GPU (perform_dugks_step, -DDUGKS branch), dt = min(5 tau, 0.5)"),
    # what app/main_vortex.f90 runs: stream_fvm_bardow + collide_bgk through perform_step
    "c4_fvm_bardow_f64_2048": dict(ny=2048, nxl=2048, scheme="fvm_bardow", collision="bgk", precision="f64",
                                   desc="Bardow FVM + BGK fp64 2048 x 2048 per GPU (perform_step: stream_fvm_bardow + collide_bgk), dt = min(5 tau, 0.5)"),
}
DEFAULT_STEPS = {"c1_bgk_f64_64": 20000, "c2_trt_f64_1024": 10000, "c3_rr_f64_8192": 300, "c3_rr_f32_8192": 400,
                 "c4_dugks_f64_2048": 1000, "c4_dugks_f32_2048": 1000, "c4_fvm_bardow_f64_2048": 1000, "c5_bgk_f64_strong": 100}
BYTES_PER_LUP = {"f64": 144, "f32": 72}
# the reference author's own three-pass accounting for DUGKS (sim/standard_lbm.F90:331): 9 * 8 * 2 * 3 bytes per update
REF_DUGKS_BYTES_PER_LUP = {"f64": 432, "f32": 216}
CPU_SAMPLE_LINES = 512  # lines of the slab timed on the CPU (bounded sample)
FV_CFL_MAX = 0.5        # DUGKS / Bardow FVM time step cap (lattice units: c = dx = 1)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_lup(workload, kernel):
    """dram bytes (read+write) per lattice update of the dominant kernel, from the committed ncu --set full capture
    summarised in profiles/ncu_summary.json (None when that kernel was not captured on this workload).  Keys:
    `<workload>@<kernel>`, or `<workload>` for k_lbm2.  DRAM bytes cannot be counted outside a profiler, so this is a
    constant of the committed capture, not an in-run measurement; `roofline.traffic_source` says so."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as fh:
            d = json.load(fh)
        e = d.get(f"{workload}@{kernel}") or (d.get(workload) if kernel == "k_lbm2" else None)
        return (float(e["dram_bytes_per_lup"]), e.get("source", "profiles/ncu_summary.json")) if e else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, seconds=4.0):
        """nvidia-smi takes a moment to start (longer on a fresh box): do not begin the timed region before it samples"""
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < seconds:
            try:
                if os.path.getsize(self.tmp.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        sm, mx, reasons, pw = [], [], set(), []
        with open(self.tmp.name) as fh:
            for line in fh:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                    pw.append(float(c[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(pw)))
        return out


def config_of(name, world):
    """the `config` object both arms print (same keys and values, so the two lines are comparable)"""
    w = WORKLOADS[name]
    strong = w["nxl"] < 0
    nxl = -w["nxl"] // world if strong else w["nxl"]
    return {"workload": name, "description": w["desc"], "scheme": w["scheme"], "ny_fast": w["ny"], "nx_slow_per_gpu": nxl,
            "nx_slow_global": nxl * world, "collision": w["collision"], "lattice": "D2Q9, two lattices, SoA f(ld,nx,0:8)"}


# ---------------------------------------------------------------------------------------------
def cpu_reference_mlups(workload, seconds_budget, steps=None, warmup=1):
    """Reference CPU path (separate stream and collide sweeps / DUGKS collide + stream passes, OpenMP over x, same layout)
    on a bounded SAMPLE: at most CPU_SAMPLE_LINES lines of the ny-wide grid.  The only place bench.py executes oracle/."""
    from oracle.oracle import Oracle, OracleGrid, taylor_green_setup

    w = WORKLOADS[workload]
    ny, precision, collision, scheme = w["ny"], w["precision"], w["collision"], w["scheme"]
    nx_full = abs(w["nxl"])
    nx = min(CPU_SAMPLE_LINES, nx_full)
    og = OracleGrid(nx, ny, precision, omp=True)
    o = og.o
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it)
    try:
        o.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        o.set_num_threads(os.cpu_count() or 1)
    cores = o.num_threads()
    s = taylor_green_setup(o, ny, dt=1.0)
    if scheme != "lbm":  # DUGKS / Bardow FVM: dt = min(5 tau, 0.5), see run_ours
        s = taylor_green_setup(o, ny, dt=min(5.0 * float(s["tau"]), FV_CFL_MAX))
    og.set_properties(s["nu"], s["dt"], magic=0.25)
    og.rho[:] = 1.0
    og.ux[:] = 0.01
    og.uy[:] = -0.02
    og.set_pdf_to_equilibrium()
    coll = {"bgk": Oracle.BGK, "trt": Oracle.TRT, "rr": Oracle.RR}[collision]
    sch = {"lbm": Oracle.SCHEME_LBM, "dugks": Oracle.SCHEME_DUGKS, "fvm_bardow": Oracle.SCHEME_FVM_BARDOW}[scheme]
    og.run(sch, coll, max(1, warmup))
    if steps is None:
        t0 = time.perf_counter()
        og.run(sch, coll, 1)
        t1 = time.perf_counter() - t0
        steps = max(2, min(200, int(seconds_budget / max(t1, 1e-6))))
    t0 = time.perf_counter()
    og.run(sch, coll, steps)
    dt = time.perf_counter() - t0
    mlups = nx * ny * steps / dt * 1e-6
    passes = {"lbm": "separate stream + collide sweeps", "dugks": "copy + two collision passes + flux pass",
              "fvm_bardow": "separate flux-streaming + collide sweeps"}[scheme]
    sample = (f"SAMPLED: {ny} x {nx} lines (1/{max(1, nx_full // nx)} of one GPU's {ny} x {nx_full} share; uniform initial state), "
              f"{steps} steps, {passes} like the reference, OpenMP {cores} threads")
    return mlups, cores, sample, dt / steps * 1e3, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    K = min(args.steps, 1000) if args.steps > 0 else None
    W = max(1, args.warmup)
    mlups, cores, sample, ms, steps = cpu_reference_mlups(args.workload, 20.0, steps=K, warmup=W)
    cfg = config_of(args.workload, max(1, args.gpus))
    cfg["note"] = "CPU run on rank 0's host cores; does not use the GPUs, value does not scale with n_gpus"
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 2), "unit": "MLUPS (1e6 lattice updates/s)",
        "n_gpus": args.gpus, "steps": steps, "warmup": W, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong" if w["nxl"] < 0 else "weak", "vs_baseline": None, "dtype": w["precision"], "data": "synthetic",
        "config": cfg, "sampled": True,
        "cpu_baseline": {"value": round(mlups, 2), "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample,
                         "why_port": "reference is Fortran; no Fortran compiler in this image (SURVEY F1): timed the line-faithful C/OpenMP restatement (oracle/)"},
        "e2e": {"value": round(mlups, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import periodic_lbm_b200 as p

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    if not torch.cuda.is_available() or p.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- libplbm_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = WORKLOADS[args.workload]
    ny, nxl, scheme, collision, precision = w["ny"], w["nxl"], w["scheme"], w["collision"], w["precision"]
    strong = nxl < 0
    if strong:  # fixed global grid, split over the ranks
        nxl = -nxl // world
    nx_global = nxl * world
    dtype = np.float64 if precision == "f64" else np.float32
    T = dtype
    stream = torch.cuda.Stream()
    tdt = torch.float64 if precision == "f64" else torch.float32

    # Taylor-Green: umax = 0.01/sqrt(3), Re = 100; LBM: dt = 1; DUGKS / Bardow FVM: dt = 5 tau
    umax = T(0.01) / np.sqrt(T(3))
    nu = umax * T(ny) / T(100)
    # finite-volume schemes: dt = 5 tau as in the reference's golden sweep, capped at CFL = dt c / dx = 0.5 -- at 2048^2 with
    # Re = 100 tau = 0.355 and 5 tau = 1.77 would be CFL 1.77: unstable, the lattice turns NaN within tens of steps (seen in r02z)
    dt_step = T(1) if scheme == "lbm" else min(T(5) * (T(3) * nu), T(FV_CFL_MAX))
    ky = T(2) * T(np.pi) / T(ny)

    def make_grid(n_lines, device):
        g = p.alloc_grid(n_lines, ny, nf=2, precision=precision, device=device)
        g.set_stream(stream.cuda_stream)
        g.collision = {"bgk": p.collide_bgk, "trt": p.collide_trt, "rr": p.collide_rr}[collision]
        g.streaming = {"lbm": p.lbm_stream, "fvm_bardow": p.stream_fvm_bardow, "dugks": None}[scheme]
        if scheme == "dugks":
            g.collision = None
        p.set_properties(g, nu, dt_step, magic=0.25)
        if args.variant:
            g.set_variant(args.variant)
        return g

    g = make_grid(nxl, local)
    if scheme == "lbm":
        step = lambda n, gg=None: p.perform_lbm_step(gg or g, n)  # noqa: E731
    elif scheme == "dugks":
        step = lambda n, gg=None: p.perform_dugks_step(gg or g, n)  # noqa: E731
    else:
        step = lambda n, gg=None: p.perform_step(gg or g, n)  # noqa: E731

    # pinned host buffers for the macroscopic fields (the only data that crosses the boundary)
    pin = [torch.empty((nxl, ny), dtype=tdt, pin_memory=True) for _ in range(3)]
    g.rho, g.ux, g.uy = (t.numpy() for t in pin)

    def fill_ic(period_lines, x_offset):
        """Taylor-Green with wavelength `period_lines` along x, evaluated on this rank's lines"""
        tg = p.TaylorGreen(period_lines, ny, T(2) * T(np.pi) / T(period_lines), ky, umax, nu, dtype=dtype)
        tg.eval(0.0, x_offset=x_offset, nx_local=nxl, out=(g.rho, g.ux, g.uy))
        g.rho[:] = g.rho / g.csqr + T(1)  # app/main_taylor_green.f90:145

    fill_ic(nx_global, rank * nxl)

    transport = None
    if world > 1:
        from periodic_lbm_b200.capi import check, lib
        import ctypes as C
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * 128)()
            check(lib.plbm_comm_unique_id(raw), "comm_unique_id")
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        idraw = C.create_string_buffer(idbuf.cpu().numpy().tobytes(), 128)
        check(lib.plbm_comm_init(g._h, idraw, rank, world, nx_global, rank * nxl), "comm_init")
        transport = {1: "CUDA-IPC peer stores over NVLink from a push kernel + cuStreamWaitValue32 epoch flags (NCCL only bootstraps)",
                     0: "grouped ncclSend/ncclRecv"}.get(lib.plbm_comm_transport(g._h), "?")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def timed(fn):
        """device time of fn() on the library's stream, barrier + synchronize on both sides, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        fn()
        e1.record(stream)
        barrier()
        return allmax([e0.elapsed_time(e1)])[0]

    p.set_pdf_to_equilibrium(g)
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident timing: `value` ---------------------------------------------------
    # Warm-up: W steps as asked, preceded by one 8-step call (triple + pair + pair + single under the default schedule) so that
    # every kernel the timed call launches has been loaded: CUDA loads a kernel's code at its first launch, and a 5-step warm-up
    # (pair, pair, single) left the first three-step launch -- 15 ms of module loading -- inside a 20-step timed region (r02z).
    step(8)
    step(4)  # triple + single: with a third lattice buffer (closing dual triple) the 8-step call launches no single step
    step(W)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    barrier()
    l0 = p.launch_count()
    ms_total = timed(lambda: step(K))
    launches = p.launch_count() - l0
    if world > 1:
        t = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = int(t[0])
    ms_step = ms_total / K
    nodes_global = nx_global * ny
    nodes_local = nxl * ny
    mlups = nodes_global / ms_step * 1e-3

    # ---- per-launch duration of the dominant kernel (roofline), events on the launching stream, median of 7 ---------
    def call_ms(nsteps):
        return sorted(timed(lambda: step(nsteps)) for _ in range(7))[3]

    if scheme == "lbm":
        # a 3-step call (one two-step launch + one k_lbm) minus a 1-step call (one k_lbm); under a ring the same
        # difference is the boundary + interior launches of one pair (every rank issues the same sequence).  Where the
        # call advances three steps per pass (k_lbmn_bulk): a 4-step call (one triple + one k_lbm) minus a 1-step call.
        spp = g.steps_per_pass()
        t1 = call_ms(1)
        if spp == 3:
            dom_ms, dom_steps, single_ms, dom_kernel = call_ms(4) - t1, 3, t1, g.triple_kernel()
        else:
            dom_ms, dom_steps, single_ms = call_ms(3) - t1, 2, t1
            dom_kernel = g.pair_kernel()
        if dom_kernel == "k_lbm":  # grids the multi-step kernels do not take: one step per launch
            dom_ms, dom_steps = t1, 1
        if world == 1 and launches == 1 and K >= 4:
            # grids that fit in the shared memory of one thread-block cluster: ALL K steps ran in one launch (csrc/plbm_small.cu)
            dom_kernel, dom_ms, dom_steps = "k_lbm_cluster", ms_total, K
    else:
        # one launch per step: a 41-step call minus a 1-step call, per step (short calls are dominated by launch latency)
        t1, t41 = call_ms(1), call_ms(41)
        dom_ms, dom_steps, single_ms = (t41 - t1) / 40, 1, t1
        dom_kernel = {"dugks": "k_fv_tma<MODE_DUGKS>", "fvm_bardow": "k_fv_tma<MODE_BARDOW>"}[scheme]
        if args.variant == 4:
            dom_kernel = {"dugks": "k_fv_march<MODE_DUGKS>", "fvm_bardow": "k_fv_march<MODE_BARDOW>"}[scheme]

    # the clock samples cover the timed call and the launch-duration calls above (the same kernels on the same lattice: at
    # 50 ms per sample the 20-step call alone is shorter than one sampling period)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the timed call and the per-launch timing calls that follow it (same kernels, same grid)"

    # ---- one step per call, the way the reference drivers call (app/main_taylor_green.f90:98-119) ---------------------
    per_call = None
    if scheme == "lbm":
        ncalls = K
        def calls():
            for _ in range(ncalls):
                step(1)
        ms_eager = timed(calls)
        per_call = {"calls": ncalls, "eager": {"ms_per_step": round(ms_eager / ncalls, 4), "mlups": round(nodes_global * ncalls / ms_eager * 1e-3, 1),
                                               "what": "K x perform_lbm_step(1), default (eager): one k_lbm launch per call"}}
        if world == 1:  # deferral is a single-GPU feature (a ring must issue identical launch sequences on every rank)
            g.set_step_deferral(64)
            def calls_deferred():
                for _ in range(ncalls):
                    step(1)
                g.synchronize()  # any observer runs what is pending; the flush is inside the timed region
            ms_def = timed(calls_deferred)
            g.set_step_deferral(0)
            per_call["deferred"] = {"ms_per_step": round(ms_def / ncalls, 4), "mlups": round(nodes_global * ncalls / ms_def * 1e-3, 1),
                                    "what": "K x perform_lbm_step(1) with plbm_set_step_deferral(64) (on in the Fortran shim): the calls are counted "
                                            "and run batched, two or three steps per pass over HBM, when 64 are pending or anything looks at the grid"}

    # ---- C4: the opt-in, tolerance-gated kernels next to the bit-identical default, same grid, same call ----------------
    fast = None
    if scheme != "lbm" and world == 1 and args.variant == 0:
        fast = {}
        KD = 50  # steps after which the deviation from the default kernel's lattice is taken
        fill_ic(nx_global, 0)
        p.set_pdf_to_equilibrium(g)
        step(KD)
        ref_lat = g.download_f(g.iold)[:, :, :ny]
        for v, what in ((4, "k_fv_march: marching kernel, every cell face reconstructed and relaxed once and shared by its two cells, FMA contraction"),
                        (3, "k_fv_tma compiled with FMA contraction")):
            g.set_variant(v)
            p.set_pdf_to_equilibrium(g)
            step(KD)
            lat = g.download_f(g.iold)[:, :, :ny]
            dev = float(np.max(np.abs(lat.astype(np.float64) - ref_lat.astype(np.float64)) / np.abs(ref_lat.astype(np.float64))))
            step(W)
            ms_v = timed(lambda: step(K))
            fast[f"variant_{v}"] = {"what": what, "mlups": round(nodes_global * K / ms_v * 1e-3, 1), "ms_per_step": round(ms_v / K, 4),
                                    "frac": round(nodes_local * BYTES_PER_LUP[precision] / (ms_v / K * 1e-3) / 1e9 / measured_peak()[0], 4),
                                    "max_rel_dev_from_default": dev, "after_steps": KD,
                                    "tolerance": 1e-12 if precision == "f64" else 1e-5, "bit_identical": bool(dev == 0.0)}
        g.set_variant(0)
        del ref_lat

    # physics sanity inside the bench: mass is conserved to round-off (GLOBAL sum under a ring)
    p.update_macros(g, lagged=False)
    mass = float(g.diagnostics()["sum_rho"])

    # ---- end-to-end through the C ABI with host buffers: `e2e` ------------------------------
    # One driver cycle as the reference apps run it (app/main_taylor_green.f90:133-149, 98-118):
    #   set_pdf_to_equilibrium(host rho,ux,uy) -> K steps -> update_macros(host)
    # H2D = 3 fields, D2H = 3 fields per cycle, from/to pinned host memory, all inside the timing.
    fill_ic(nx_global, rank * nxl)
    def cycle():
        p.set_pdf_to_equilibrium(g)
        step(K)
        p.update_macros(g)  # lagged, like the reference driver
    cycle()  # untimed: like the W warm-up steps of `value` (a single cold cycle moved between 150 and 280 ms from box to box, r02q)
    e2e_all = [timed(cycle) for _ in range(3)]
    e2e_ms = sorted(e2e_all)[1]  # median of three cycles
    # where the cycle goes: the two transfers (PCIe) against the steps (HBM)
    fill_ic(nx_global, rank * nxl)
    ms_up = timed(lambda: p.set_pdf_to_equilibrium(g))
    step(K)
    ms_down = timed(lambda: p.update_macros(g))
    field_bytes = nxl * ny * np.dtype(dtype).itemsize * 3
    e2e_mlups = nodes_global * K / e2e_ms * 1e-3

    # ---- selfcheck: the slabs of the ring hold the single-GPU result, bit for bit ---------------------------------------
    # Initial condition periodic along x with period nx_slow_per_gpu, so every slab of the ring sees exactly what a single
    # GPU sees whose own periodic domain is that slab: after the same call every rank's lattice must have the same 64-bit
    # checksum (plbm_lattice_hash, computed on the device) as that single-GPU run -- which rank 0 performs here, in the same
    # process, on a second grid without a ring.  The hash is therefore also equal across N = 1, 2, 4, 8 of a weak-scaling run.
    # Two calls (4 + 5 steps: triple, single | pair, pair, single), so that every kind of launch also CONSUMES the halo
    # message the launch before it produced, across a call boundary too.
    KS_CALLS = (4, 5) if scheme == "lbm" else (3,)
    KS = sum(KS_CALLS)
    fill_ic(nxl, 0)
    p.set_pdf_to_equilibrium(g)
    for k in KS_CALLS:
        step(k)
    h_ring = g.lattice_hash(g.iold)
    hashes = [h_ring]
    if world > 1:
        t = torch.tensor([h_ring & 0xFFFFFFFF, h_ring >> 32], dtype=torch.int64, device="cuda")
        allh = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        hashes = [int(a[0]) | (int(a[1]) << 32) for a in allh]
    selfcheck = None
    if rank == 0:
        h_single = h_ring
        if world > 1:
            try:
                g1 = make_grid(nxl, local)
            except p.PlbmError:  # no room for a second slab next to the first (strong scaling at N = 2)
                g1, h_single = None, None
            if g1 is not None:
                g1.rho, g1.ux, g1.uy = g.rho, g.ux, g.uy
                p.set_pdf_to_equilibrium(g1)
                for k in KS_CALLS:
                    step(k, g1)
                h_single = g1.lattice_hash(g1.iold)
                p.dealloc_grid(g1)
        selfcheck = {"hash": f"{h_ring:016x}", "ranks_equal": len(set(hashes)) == 1,
                     "single_gpu_hash": None if h_single is None else f"{h_single:016x}",
                     "equals_single_gpu": None if h_single is None else all(h == h_single for h in hashes), "steps": KS,
                     "what": f"Taylor-Green periodic over the {nxl} lines of one slab; {KS} steps in calls of {KS_CALLS}; plbm_lattice_hash(iold) of every rank "
                             "vs the same slab stepped alone on one GPU in this run"}

    if rank == 0:
        peak, peak_src = measured_peak()
        bpl = BYTES_PER_LUP[precision]
        # algorithmic bytes = 9 reads + 9 writes per node PER STEP, so for the two-step LBM kernels frac > 1 means
        # the kernel moves fewer HBM bytes than a one-step-per-pass algorithm can; `traffic` (ncu) and `dram_frac` say how
        # close the bytes it does move are to the HBM roof.
        achieved = dom_steps * nodes_local * bpl / (dom_ms * 1e-3) / 1e9
        tr, tr_src = ncu_traffic_per_lup(args.workload, dom_kernel.split("<")[0])
        traffic = None if tr is None else round(tr * dom_steps * nodes_local)
        cfg = config_of(args.workload, world)
        closing_dual = scheme == "lbm" and g.closing_triple()
        def lbm_schedule(k, depth):
            """launches of one perform_lbm_step(k) call: triples while more than three steps remain (depth 3), pairs while more than two"""
            from periodic_lbm_b200.slab import launch_schedule
            sched = launch_schedule(k, pairs=depth >= 2, triples=depth == 3, dual=closing_dual)
            return sched.count(3), sched.count(2), sched.count(1)
        n3, n2, n1 = lbm_schedule(K, dom_steps if scheme == "lbm" else 1)
        closing = (", the last three-step launch also stores state n-1 into a third lattice buffer (no closing single step)"
                   if closing_dual and n1 == 0 and n3 > 0 else "")
        stepping = {"lbm": f"one perform_lbm_step(K={K}) call: {n3} three-step launches + {n2} two-step launches + {n1} single-step launches{closing}, bit-identical to K single steps",
                    "dugks": f"one perform_dugks_step(K={K}) call: one fused launch per step (collide + face reconstruction + face relaxation + flux update)",
                    "fvm_bardow": f"one perform_step(K={K}) call: one fused launch per step (stream_fvm_bardow + collide_bgk)"}[scheme]
        cfg.update({"stepping": stepping, "variant": args.variant,
                    "warmup_calls": f"an 8-step and a 4-step call (load every kernel the schedule uses) + one {W}-step call, all untimed",
                    "halo": f"3 lines x 9 populations per direction per launch (a launch of one, two or three steps reads one, two or three of them), overlapped with the interior update; transport: {transport}" if world > 1 else "none (periodic index wrap)",
                    "l2": f"inputs vs L2: {2 * 9 * nxl * ny * np.dtype(dtype).itemsize / 1e9:.3f} GB of PDFs per GPU vs 126 MB L2"
                          + (" (larger than L2: no flush needed)" if 2 * 9 * nxl * ny * np.dtype(dtype).itemsize > 4 * 126e6 else " (L2-resident: a launch/latency figure, not a roofline case)"),
                    "mass_sum_rho": mass})
        line = {
            "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS (1e6 lattice updates/s)", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": precision,
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": round(e2e_mlups, 1), "unit": "MLUPS", "h2d_bytes_per_step": int(field_bytes / K), "d2h_bytes_per_step": int(field_bytes / K),
                    "cycle": f"set_pdf_to_equilibrium(host) + {K} steps + update_macros(host) per GPU; {field_bytes} B H2D and {field_bytes} B D2H per cycle (pinned), amortised over the {K} steps",
                    "ms_per_cycle": round(e2e_ms, 3), "cycles_ms": [round(x, 3) for x in e2e_all], "cycles": "one untimed cycle, then the median of three",
                    "breakdown_ms": {"set_pdf_to_equilibrium_h2d_plus_init_kernel": round(ms_up, 3), "steps": round(ms_total, 3),
                                     "update_macros_kernel_plus_d2h": round(ms_down, 3),
                                     "note": f"host link: {field_bytes / 1e9 / max(ms_up, 1e-9) * 1e3:.1f} GB/s up, {field_bytes / 1e9 / max(ms_down, 1e-9) * 1e3:.1f} GB/s down (pinned host memory)"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "traffic_source": tr_src and f"ncu --set full capture of this kernel on this workload, {tr_src} (a profiler constant, not measured in this run)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_steps * nodes_local * bpl,
                         "kernel": f"{dom_kernel}<{ {1: 'one step', 2: 'two fused stream+collide steps', 3: 'three fused stream+collide steps'}.get(dom_steps if dom_kernel != 'k_lbm_cluster' else 0, f'all {dom_steps} steps of the call in one launch, lattices resident in distributed shared memory') }>" if scheme == "lbm" else dom_kernel,
                         "launch_ms": round(dom_ms, 4), "per_gpu": True,
                         "dram_frac": None if traffic is None else round(traffic / (dom_ms * 1e-3) / 1e9 / peak, 4)},
            "selfcheck": selfcheck,
        }
        if scheme == "lbm":
            line["roofline"]["single_step_kernel"] = {"kernel": "k_lbm<fused stream+collide>", "launch_ms": round(single_ms, 4),
                                                      "achieved": round(nodes_local * bpl / (single_ms * 1e-3) / 1e9, 1),
                                                      "frac": round(nodes_local * bpl / (single_ms * 1e-3) / 1e9 / peak, 4)}
            line["per_call"] = per_call
        if scheme == "dugks":
            ref_bpl = REF_DUGKS_BYTES_PER_LUP[precision]
            line["roofline"]["vs_reference_3pass_model"] = {
                "bytes_per_update": ref_bpl, "achieved": round(nodes_local * ref_bpl / (dom_ms * 1e-3) / 1e9, 1),
                "frac": round(nodes_local * ref_bpl / (dom_ms * 1e-3) / 1e9 / peak, 4),
                "what": "the reference author's accounting for DUGKS, 9*8*2*3 B per update (sim/standard_lbm.F90:331); `achieved`/`frac` above use the fused lower bound, state in + state out"}
        if fast is not None:
            line["fast_variants"] = fast
        if world == 1 and not args.no_cpu:
            v, cores, sample, _, _ = cpu_reference_mlups(args.workload, 12.0)
            line["cpu_baseline"] = {"value": round(v, 2), "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        from periodic_lbm_b200.capi import lib
        lib.plbm_comm_finalize(g._h)
        dist.barrier()
        dist.destroy_process_group()
    p.dealloc_grid(g)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_bgk_f64_slab", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=0, help="plbm_set_variant: 0 = default (bit-identical) kernels; see include/plbm.h")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.steps <= 0:
        # long enough that the timed region spans several clock samples (0.1 - 0.5 s of GPU time)
        args.steps = DEFAULT_STEPS.get(args.workload, 300)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
