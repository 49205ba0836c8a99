#!/usr/bin/env python
"""bench.py -- headline benchmark of the periodic D2Q9 hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): MLUPS = lattice updates / microsecond (app/main_taylor_green.f90:122) of the
D2Q9 fp64 stream+collide step, whole job over all N GPUs, plus achieved HBM GB/s against the
measured peak (144 B per lattice update: 9 reads + 9 writes of 8 bytes).

Default workload = BASELINE.json configs[4], the configuration the metric is quoted on at
1/2/4/8 GPUs: D2Q9 BGK fp64, slab decomposition along the slow index, WEAK scaling with a
32768 (unit-stride) x 4096 (slow) slab per GPU -- at N = 8 this is the full 32768 x 32768 grid.
One "step" = one fused stream+collide pass over the whole grid.

For N > 1 the driver launches this file with torch.distributed.run, one rank per GPU; ranks
exchange one halo line of three populations per direction per step (inside libplbm_b200.so: peer
stores over NVLink through CUDA IPC mappings, or NCCL send/recv as fallback), overlapped with the
interior update.

`--impl reference` times the reference's own CPU algorithm (the line-faithful C/OpenMP restatement
in oracle/, because the Fortran reference cannot be compiled in this image -- no gfortran) with all
host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (ny_fast, nx_slow_per_gpu, collision, precision, description)
    "c5_bgk_f64_slab": (32768, 4096, "bgk", "f64", "C5 D2Q9 BGK fp64 Taylor-Green, y-slab weak scaling, 32768 x 4096 lines per GPU"),
    # strong scaling of the full C5 grid: nx_slow_per_gpu = 32768 / N (154.6 GB of PDFs on one GPU)
    "c5_bgk_f64_strong": (32768, -32768, "bgk", "f64", "C5 D2Q9 BGK fp64 Taylor-Green 32768 x 32768, y-slab STRONG scaling"),
    "c3_rr_f64_8192": (8192, 8192, "rr", "f64", "C3 D2Q9 recursive-regularized fp64 Taylor-Green 8192 x 8192 per GPU"),
    "c3_rr_f32_8192": (8192, 8192, "rr", "f32", "C3 D2Q9 recursive-regularized fp32 Taylor-Green 8192 x 8192 per GPU"),
    "c2_trt_f64_1024": (1024, 1024, "trt", "f64", "C2 D2Q9 TRT fp64 1024 x 1024 per GPU"),
    "c1_bgk_f64_64": (64, 64, "bgk", "f64", "C1 D2Q9 BGK fp64 Taylor-Green 64 x 64 (L2-resident, launch-bound)"),
}
BYTES_PER_LUP = {"f64": 144, "f32": 72}
CPU_SAMPLE_LINES = 512  # lines of the slab timed on the CPU (bounded sample)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_lup(workload, kernel="k_lbm2"):
    """dram bytes (read+write) per lattice update of the dominant kernel, from the committed ncu
    --set full capture summarised in profiles/ncu_summary.json (None when that kernel was not captured
    on this workload).  Keys: `<workload>` for k_lbm2, `<workload>@<kernel>` for the others."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as fh:
            return float(json.load(fh)[workload if kernel == "k_lbm2" else f"{workload}@{kernel}"]["dram_bytes_per_lup"])
    except Exception:
        return None


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


# ---------------------------------------------------------------------------------------------
def cpu_reference_mlups(ny, collision, precision, seconds_budget, steps=None, warmup=1):
    """Reference CPU path (stream sweep + collide sweep, OpenMP over x, same layout) on a bounded
    sample: CPU_SAMPLE_LINES lines of the ny-wide slab.  Only place bench.py executes oracle/."""
    from oracle.oracle import Oracle, OracleGrid, taylor_green_setup

    nx = min(CPU_SAMPLE_LINES, ny)
    og = OracleGrid(nx, ny, precision, omp=True)
    o = og.o
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it)
    try:
        o.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        o.set_num_threads(os.cpu_count() or 1)
    cores = o.num_threads()
    s = taylor_green_setup(o, ny, dt=1.0)
    og.set_properties(s["nu"], s["dt"], magic=0.25)
    og.rho[:] = 1.0
    og.ux[:] = 0.01
    og.uy[:] = -0.02
    og.set_pdf_to_equilibrium()
    coll = {"bgk": Oracle.BGK, "trt": Oracle.TRT, "rr": Oracle.RR}[collision]
    og.run(Oracle.SCHEME_LBM, coll, warmup)
    t0 = time.perf_counter()
    og.run(Oracle.SCHEME_LBM, coll, 1)
    t1 = time.perf_counter() - t0
    if steps is None:
        steps = max(2, min(200, int(seconds_budget / max(t1, 1e-6))))
    t0 = time.perf_counter()
    og.run(Oracle.SCHEME_LBM, coll, steps)
    dt = time.perf_counter() - t0
    mlups = nx * ny * steps / dt * 1e-6
    sample = (f"{ny} x {nx} lines of the slab (1/{max(1, 4096 // nx)} of one GPU's share), {steps} steps, "
              f"separate stream + collide sweeps like the reference, OpenMP {cores} threads")
    return mlups, cores, sample, dt / steps * 1e3, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ny, nxl, collision, precision, desc = WORKLOADS[args.workload]
    nxl = abs(nxl)
    steps = args.steps if args.steps else None
    # keep the whole run within a few minutes: cap the number of timed steps
    mlups, cores, sample, ms, steps = cpu_reference_mlups(ny, collision, precision, 20.0, steps=min(steps, 100) if steps else None,
                                                          warmup=max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 2), "unit": "MLUPS (1e6 lattice updates/s)",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong" if WORKLOADS[args.workload][1] < 0 else "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "ny_fast": ny, "nx_slow_per_gpu": nxl, "collision": collision,
                   "note": "CPU run does not use the GPUs; value does not scale with n_gpus"},
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

    ny, nxl, collision, precision, desc = WORKLOADS[args.workload]
    strong = nxl < 0
    if strong:  # fixed global grid, split over the ranks
        nxl = -nxl // world
    nx_global = nxl * world
    dtype = np.float64 if precision == "f64" else np.float32
    stream = torch.cuda.Stream()

    g = p.alloc_grid(nxl, ny, nf=2, precision=precision, device=local)
    g.set_stream(stream.cuda_stream)
    g.collision = {"bgk": p.collide_bgk, "trt": p.collide_trt, "rr": p.collide_rr}[collision]
    g.streaming = p.lbm_stream
    # Taylor-Green on the GLOBAL grid: umax = 0.01/sqrt(3), Re = 100, dt = 1
    T = dtype
    umax = T(0.01) / np.sqrt(T(3))
    nu = umax * T(ny) / T(100)
    kx = T(2) * T(np.pi) / T(nx_global)
    ky = T(2) * T(np.pi) / T(ny)
    tg = p.TaylorGreen(nx_global, ny, kx, ky, umax, nu, dtype=dtype)
    p.set_properties(g, nu, 1.0, magic=0.25)

    # pinned host buffers for the macroscopic fields (the only data that crosses the boundary)
    pin = [torch.empty((nxl, ny), dtype=torch.float64 if precision == "f64" else torch.float32, pin_memory=True) for _ in range(3)]
    g.rho, g.ux, g.uy = (t.numpy() for t in pin)
    tg.eval(0.0, x_offset=rank * nxl, nx_local=nxl, out=(g.rho, g.ux, g.uy))
    g.rho[:] = g.rho / g.csqr + T(1)  # app/main_taylor_green.f90:145

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

    p.set_pdf_to_equilibrium(g)
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident timing: `value` ---------------------------------------------------
    p.perform_lbm_step(g, W)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.12)
    l0 = p.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    p.perform_lbm_step(g, K)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = p.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, launches = float(tmax[0]), int(tsum[1])
    ms_step = ms_total / K
    nodes_global = nx_global * ny
    mlups = nodes_global / ms_step * 1e-3
    # per-launch duration of the two kernels of the path (roofline), events on the launching stream
    def call_ms(nsteps):
        ts = []
        for _ in range(7):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a0.record(stream)
            p.perform_lbm_step(g, nsteps)
            a1.record(stream)
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        return sorted(ts)[len(ts) // 2]
    pair_kernel_ms = None
    if world == 1:
        t1, t3 = call_ms(1), call_ms(3)
        pair_kernel_ms = (t3 - t1, t1)
    else:  # under a ring the launches are split in boundary + interior: use the step average
        pair_kernel_ms = (2 * ms_step, ms_step)
    # physics sanity inside the bench: mass is conserved to round-off
    p.update_macros(g, lagged=False)
    mass = float(g.diagnostics()["sum_rho"])

    # ---- end-to-end through the C ABI with host buffers: `e2e` ------------------------------
    # One driver cycle as the reference apps run it (app/main_taylor_green.f90:133-149, 98-118):
    #   set_pdf_to_equilibrium(host rho,ux,uy) -> K x perform_lbm_step -> update_macros(host)
    # H2D = 3 fields, D2H = 3 fields per cycle, from/to pinned host memory, all inside the timing.
    tg.eval(0.0, x_offset=rank * nxl, nx_local=nxl, out=(g.rho, g.ux, g.uy))
    g.rho[:] = g.rho / g.csqr + T(1)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    c0.record(stream)
    p.set_pdf_to_equilibrium(g)
    p.perform_lbm_step(g, K)
    p.update_macros(g)  # lagged, like the reference driver
    c1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - w0) * 1e3
    e2e_ms = max(c0.elapsed_time(c1), 0.0)
    t = torch.tensor([e2e_ms, e2e_wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    field_bytes = nxl * ny * np.dtype(dtype).itemsize * 3
    e2e_mlups = nodes_global * K / e2e_ms * 1e-3

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        bpl = BYTES_PER_LUP[precision]
        # Dominant kernel: k_lbm2, TWO fused stream+collide steps per launch (temporal blocking through a
        # shared-memory ring); a K-step call runs (K-1)//2 of them and finishes with 1-2 single-step k_lbm
        # launches.  Its launch duration is measured live below (CUDA events on its stream): a 3-step call
        # (one k_lbm2 + one k_lbm) minus a 1-step call (one k_lbm), median of 7.
        # algorithmic bytes = 9 reads + 9 writes per node PER STEP, so frac > 1 means the
        # kernel moves fewer HBM bytes than the one-step-per-pass algorithm can; `traffic` (ncu) and
        # `dram_frac` say how close the bytes it does move are to the HBM roof.
        nodes_local = nxl * ny
        pair_ms, single_ms = pair_kernel_ms
        achieved = 2 * nodes_local * bpl / (pair_ms * 1e-3) / 1e9
        pair_kernel = g.pair_kernel()  # k_lbm2 (raw columns by per-thread loads) or k_lbm2_bulk (by bulk async copies)
        tr = ncu_traffic_per_lup(args.workload, pair_kernel)
        traffic = None if tr is None else round(tr * 2 * nodes_local)
        line = {
            "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS (1e6 lattice updates/s)", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": precision,
            "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "ny_fast": ny, "nx_slow_per_gpu": nxl, "nx_slow_global": nx_global,
                       "collision": collision, "lattice": "D2Q9, two lattices, SoA f(ld,nx,0:8)",
                       "stepping": f"one perform_lbm_step(K={K}) call: {(K - 1) // 2} two-step launches + {K - 2 * ((K - 1) // 2)} single-step launches, bit-identical to K single steps",
                       "halo": f"2 lines x 9 populations per direction per launch, overlapped with the interior update; transport: {transport}" if world > 1 else "none (periodic index wrap)",
                       "l2": f"inputs larger than L2: {2 * 9 * nxl * ny * np.dtype(dtype).itemsize / 1e9:.1f} GB of PDFs per GPU vs 126 MB L2 (no flush needed)",
                       "mass_sum_rho": mass},
            "clocks": clocks,
            "e2e": {"value": round(e2e_mlups, 1), "unit": "MLUPS", "h2d_bytes_per_step": int(field_bytes / K), "d2h_bytes_per_step": int(field_bytes / K),
                    "cycle": f"set_pdf_to_equilibrium(host) + {K} steps + update_macros(host) per GPU; {field_bytes} B H2D and {field_bytes} B D2H per cycle (pinned), amortised over the {K} steps",
                    "ms_per_cycle": round(e2e_ms, 3)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 2 * nodes_local * bpl, "kernel": f"{pair_kernel}<two fused stream+collide steps>",
                         "launch_ms": round(pair_ms, 4), "per_gpu": True,
                         "dram_frac": None if traffic is None else round(traffic / (pair_ms * 1e-3) / 1e9 / peak, 4),
                         "single_step_kernel": {"kernel": "k_lbm<fused stream+collide>", "launch_ms": round(single_ms, 4),
                                                "achieved": round(nodes_local * bpl / (single_ms * 1e-3) / 1e9, 1),
                                                "frac": round(nodes_local * bpl / (single_ms * 1e-3) / 1e9 / peak, 4)}},
        }
        if world == 1 and not args.no_cpu:
            v, cores, sample, _, _ = cpu_reference_mlups(ny, collision, precision, 12.0)
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.steps <= 0:
        args.steps = 300
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
