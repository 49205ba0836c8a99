"""Host-side logic that needs no GPU: slab partitioning, driver recipes, scalar helpers, and the
multi-rank halo protocol (world_size 2 over gloo, oracle as the per-slab compute)."""
import os
import socket

import numpy as np
import pytest

from oracle.oracle import Oracle, OracleGrid, padded_ld, taylor_green_setup, taylor_green_steps_to_tmax


def test_slab_partition_covers_grid():
    from periodic_lbm_b200.slab import Q_FROM_HI, Q_FROM_LO, halo_message_bytes, slab_of
    from periodic_lbm_b200.lattice import cx

    for nxg, n in [(32768, 8), (100, 3), (17, 4), (16, 8)]:
        slabs = [slab_of(r, n, nxg) for r in range(n)]
        assert slabs[0].x_offset == 0 and slabs[-1].x_end == nxg
        assert all(a.x_end == b.x_offset for a, b in zip(slabs, slabs[1:]))
        assert max(s.nx_local for s in slabs) - min(s.nx_local for s in slabs) <= 1
        assert all(s.lo == (s.rank - 1) % n and s.hi == (s.rank + 1) % n for s in slabs)
    with pytest.raises(ValueError):
        slab_of(0, 8, 15)
    assert [q for q in range(9) if cx[q] == 1] == list(Q_FROM_LO)
    assert [q for q in range(9) if cx[q] == -1] == sorted(Q_FROM_HI)
    assert halo_message_bytes(32768, 8) == 3 * 9 * 32768 * 8  # three lines of all nine populations (a fused triple of steps may follow)


def test_driver_recipe_matches_oracle_recipe():
    import periodic_lbm_b200 as p

    for dtype, prec in ((np.float64, "f64"), (np.float32, "f32")):
        o = Oracle(prec)
        s = taylor_green_setup(o, 64, dt_over_tau=5.0)
        t = p.taylor_green_params(64, dt_over_tau=5.0, dtype=dtype)
        for k in ("umax", "nu", "tau", "dt", "kx", "tmax"):
            assert t[k] == s[k], k
        assert t["nsteps"] == s["nsteps_cap"] and t["case"].td == s["td"]
        assert p.steps_until(t["tmax"], t["dt"], t["nsteps"], dtype) == taylor_green_steps_to_tmax(o, s)
    assert p.steps_until(np.float64(9731.422537830847), 1.0, 10704) == (9732, 9732.0)


def test_trt_scalar_helpers():
    import periodic_lbm_b200 as p

    o = Oracle("f64")
    for om, x in [(1.95, 0.25), (0.5, 3.0 / 16.0), (1.0, 1.0 / 12.0)]:
        assert p.lambda_d(om, x) == o.lambda_d(om, x)
        ld = p.lambda_d(om, x)
        assert abs(p.magic_number(om, ld) - x) < 1e-15  # magic_number inverts lambda_d
        assert p.magic_number(om, ld) == o.magic_number(om, ld)


# ---------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _slab_worker(rank, world, port, nxg, ny, steps, coll, out, dual=False):
    """One rank of the ring: owns a slab, exchanges the halo message (three lines of all nine populations per
    direction) with gloo, and advances its slab with the ORACLE kernels on a halo-extended copy -- one step,
    a fused pair or a fused triple per launch, in the library's own schedule.  This exercises exactly the protocol
    libplbm_b200's ring implements (which lines, which neighbour, what a pair needs) without a GPU."""
    import torch
    import torch.distributed as dist

    from periodic_lbm_b200.slab import HALO_LINES, launch_schedule, slab_of

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle("f64")
    sl = slab_of(rank, world, nxg)
    ld = padded_ld(ny)
    rng = np.random.default_rng(5)
    fglob = np.zeros((9, nxg, ld))
    fglob[:, :, :ny] = 0.1 + 0.01 * rng.random((9, nxg, ny))
    f = fglob[:, sl.x_offset:sl.x_end].copy()
    nxl = sl.nx_local
    H = HALO_LINES
    collide = {0: lambda a: o.collide_bgk(a, ny, 1.7), 2: lambda a: o.collide_rr(a, ny, 1.7)}[coll]
    can = torch.tensor([int(nxl >= 4), int(nxl >= 2 * H)])
    dist.all_reduce(can, op=dist.ReduceOp.MIN)  # every rank must issue the same sequence of launches
    sched = launch_schedule(steps, pairs=bool(can[0].item()), triples=bool(can[1].item()), dual=dual)
    f_prev = None  # lattice `inew` of the reference after the call: state steps - 1
    for launch, nfused in enumerate(sched):
        send_lo = torch.from_numpy(f[:, :H].copy())    # my first three lines -> rank lo (its lines nx, nx+1, nx+2)
        send_hi = torch.from_numpy(f[:, -H:].copy())   # my last three lines  -> rank hi (its lines -3, -2, -1)
        halo_lo, halo_hi = torch.empty_like(send_hi), torch.empty_like(send_lo)
        ops = [dist.P2POp(dist.isend, send_lo, sl.lo), dist.P2POp(dist.isend, send_hi, sl.hi),
               dist.P2POp(dist.irecv, halo_hi, sl.hi), dist.P2POp(dist.irecv, halo_lo, sl.lo)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        ext = np.zeros((9, nxl + 2 * H, ld))
        ext[:, H:-H] = f
        ext[:, :H] = halo_lo.numpy()
        ext[:, -H:] = halo_hi.numpy()
        # every step on the extended slab pollutes one more ghost line from each end (its periodic wrap is
        # wrong there): three ghost lines per side keep the owned lines exact for up to three steps
        for k in range(nfused):
            dst = np.zeros_like(ext)
            o.lbm_stream(ext, dst, ny)
            collide(dst)
            ext = dst
            if launch == len(sched) - 1 and k == nfused - 2:
                # the closing triple of a ring whose ranks all hold a third lattice buffer also stores the state after its
                # second step: two ghost lines per side are polluted by then, the owned lines are exact
                f_prev = np.ascontiguousarray(ext[:, H:-H])
        f = np.ascontiguousarray(ext[:, H:-H])
    gathered = [None] * world
    dist.all_gather_object(gathered, (sl.x_offset, f, f_prev))
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda t: t[0])], axis=1)
        # single-domain oracle
        a, b = fglob.copy(), np.zeros_like(fglob)
        for _ in range(steps):
            o.lbm_stream(a, b, ny)
            collide(b)
            a, b = b, a
        ok = bool(np.array_equal(full[:, :, :ny], a[:, :, :ny]))
        if dual and sched[-1] == 3:  # after the last swap `b` holds state steps - 1: what the dual triple leaves in `inew`
            prev = np.concatenate([g[2] for g in sorted(gathered, key=lambda t: t[0])], axis=1)
            ok = ok and bool(np.array_equal(prev[:, :, :ny], b[:, :, :ny]))
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nxg,ny,coll", [(2, 12, 10, 0), (2, 9, 16, 2), (3, 13, 7, 0), (2, 7, 8, 0), (3, 19, 8, 2)])
def test_slab_halo_protocol_world_size_n_gloo(world, nxg, ny, coll):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, nxg, ny, 7, coll, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    ok = q.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert ok, "slab-decomposed run differs from the single-domain run"


@pytest.mark.parametrize("world,nxg,ny,coll,steps", [(2, 12, 10, 0, 8), (2, 14, 16, 2, 9), (3, 19, 8, 0, 11), (2, 13, 8, 2, 6)])
def test_slab_halo_protocol_with_a_closing_dual_triple_gloo(world, nxg, ny, coll, steps):
    """the same protocol when every rank holds a third lattice buffer: the call closes with a triple that also keeps the state
    after its second step -- the slabs must hold states `steps` AND `steps - 1` of the single-domain run"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, nxg, ny, steps, coll, q, True)) for r in range(world)]
    for p_ in procs:
        p_.start()
    ok = q.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert ok, "slab-decomposed run (closing dual triple) differs from the single-domain run"


def test_text_writers_use_the_reference_edit_descriptors(tmp_path):
    """output_vtk / output_gnuplot (src/output/vtk.F90:150-196, gnuplot.F90:21-39): ES24.16E3 reals, the
    reference's record structure and loop order (VTK: y outer, x inner; gnuplot: x outer, blank line per x)."""
    import types

    from periodic_lbm_b200 import output

    assert output._fmt_real(np.array([1.0, -0.5, 0.0, 1.2345678901234567e-105])) == [
        " 1.0000000000000000E+000", "-5.0000000000000000E-001", " 0.0000000000000000E+000", " 1.2345678901234567E-105"]
    assert output._fmt_real(np.array([1.0, -0.5], dtype=np.float32)) == [" 1.00000000E+00", "-5.00000000E-01"]

    nx, ny = 3, 2
    g = types.SimpleNamespace(nx=nx, ny=ny, dtype=np.float64, foldername=str(tmp_path), filename="results",
                              rho=np.arange(6.0).reshape(nx, ny) + 1, ux=np.full((nx, ny), 0.25), uy=np.full((nx, ny), -0.125))
    vtk = open(output.output_vtk(g, step=7)).read().split("\n")
    assert vtk[:5] == ["# vtk DataFile Version 3.0", "fluid", "ASCII", "DATASET STRUCTURED_POINTS", "DIMENSIONS 4 3 2 "]
    assert vtk[5] == "ORIGIN  " + " 0.0000000000000000E+000" * 3 and vtk[6] == "SPACING " + " 1.0000000000000000E+000" * 3
    assert vtk[7] == "" and vtk[8] == "CELL_DATA 6" and vtk[9:11] == ["SCALARS Density float 1", "LOOKUP_TABLE default"]
    # rho(j,i), j = y outer, i = x inner: rho[x][y] = 2x + y + 1
    assert [float(v) for v in vtk[11:17]] == [1.0, 3.0, 5.0, 2.0, 4.0, 6.0]
    assert vtk[17] == "" and vtk[18] == "VECTORS Velocity float"
    assert vtk[19] == " 2.5000000000000000E-001-1.2500000000000000E-001 0.0000000000000000E+000"
    assert output.output_vtk(g, binary=True) is None
    assert os.path.basename(output.output_vtk(g, step=7)) == "results000000007.vtk"

    txt = open(output.output_gnuplot(g)).read().split("\n")
    assert len(txt) == nx * (ny + 1) + 1 and txt[ny] == "" and txt[2 * ny + 1] == ""
    first = [float(v) for v in txt[0].split()]
    assert first == [0.5, 0.5, 1.0, 0.25, -0.125]
    assert [float(v) for v in txt[ny + 1].split()][:3] == [1.5, 0.5, 3.0]


def test_launch_schedule_with_a_closing_dual_triple():
    """periodic_lbm_b200/slab.py launch_schedule(dual=True) mirrors csrc/plbm_internal.h lbm_next_launch: with a third lattice
    buffer a call closes with a triple (no single-step launch) unless one step is left over, and never takes more launches."""
    from periodic_lbm_b200.slab import launch_schedule

    for k in range(1, 60):
        old = launch_schedule(k, pairs=True, triples=True)
        new = launch_schedule(k, pairs=True, triples=True, dual=True)
        assert sum(old) == k and sum(new) == k and old[-1] == 1
        assert len(new) <= len(old)
        if k >= 3 and k % 3 != 1:
            assert new[-1] == 3 and 1 not in new and new.count(2) == (1 if k % 3 == 2 else 0), (k, new)
        elif k >= 4:
            assert new == [3] * (k // 3) + [1], (k, new)
        else:
            assert new == [1] * k
        # no pairs on this grid: a call of 3 m + 2 steps ends triple, single, single
        np_ = launch_schedule(k, pairs=False, triples=True, dual=True)
        assert sum(np_) == k and 2 not in np_
    assert launch_schedule(20, True, True, True) == [3, 3, 3, 3, 3, 2, 3]
    assert launch_schedule(20, True, True, False) == [3, 3, 3, 3, 3, 2, 2, 1]
    assert launch_schedule(7, True, False, True) == launch_schedule(7, True, False, False)  # dual needs triples


def test_python_schedule_mirror_equals_the_library_inline(tmp_path):
    """csrc/plbm_internal.h lbm_next_launch (what step_lbm_t and the slab schedule call) against periodic_lbm_b200/slab.py
    launch_schedule, for every combination of (triples, pairs, dual) and calls of 1 ... 64 steps: the header's inline function is
    compiled into a host-only program (nvcc builds it here without a GPU; nothing of it touches the CUDA runtime)."""
    import shutil
    import subprocess

    from periodic_lbm_b200.slab import launch_schedule

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sched.cu"
    src.write_text('#include <cstdio>\n#include "plbm_internal.h"\n'
                   "int main() {\n"
                   "  for (int flags = 0; flags < 8; ++flags)\n"
                   "    for (int k = 1; k <= 64; ++k) {\n"
                   '      std::printf("%d %d:", flags, k);\n'
                   "      for (int s = 0; s < k;) {\n"
                   "        plbm::LbmLaunch L = plbm::lbm_next_launch(k - s, flags & 1, flags & 2, flags & 4);\n"
                   '        std::printf(" %d%s", L.depth, L.dual ? "d" : "");\n'
                   "        s += L.depth;\n"
                   "      }\n"
                   '      std::printf("\\n");\n'
                   "    }\n"
                   "  return 0;\n}\n")
    exe = tmp_path / "sched"
    r = subprocess.run([nvcc, "-std=c++17", "-I", os.path.join(root, "periodic_lbm_b200", "csrc"), "-o", str(exe), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == 8 * 64
    for line in out:
        head, seq = line.split(":")
        flags, k = (int(x) for x in head.split())
        triples, pairs, dual = bool(flags & 1), bool(flags & 2), bool(flags & 4)
        got = seq.split()
        want = launch_schedule(k, pairs=pairs, triples=triples, dual=dual)
        assert [int(x.rstrip("d")) for x in got] == want, (flags, k, got, want)
        # the dual flag: only on the closing triple of a schedule with a third buffer, and only when three steps are left
        assert [x.endswith("d") for x in got] == [dual and triples and n == 3 and i == len(want) - 1 and sum(want[:i]) == k - 3
                                                  for i, n in enumerate(want)], (flags, k, got)
