"""Executable specification of the schedule of the multi-step kernels (csrc/plbm_lbm2.cu k_lbm2_bulk and its
depth-generic form csrc/plbm_lbmn.cu k_lbmn_bulk), run on the CPU against the oracle.

`emulate()` restates, in numpy, what ONE launch does block by block: the staged raw columns with their periodic
pieces, the stage parity, the 18-slot rings (populations kept 1 / 2 / 3 columns), the per-level active rows,
the warm-up iterations and the gating of the levels -- with the oracle's collision applied to the rows of a
level.  The result must equal NSTEP oracle steps bit for bit.  It pins the index arithmetic of the kernels on
a machine without a GPU (what it cannot see: barriers, the async proxy, mbarrier phases in hardware -- those
are covered by the -m gpu parity tests).  Small NT / segment lengths make every edge (several strips, ragged
last strip, wrap pieces, segments shorter than the warm-up) cheap to hit."""
import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle, padded_ld

CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
DEPTH = [1 if c == -1 else (2 if c == 0 else 3) for c in CX]
BASE = {3: 0, 6: 1, 7: 2, 0: 3, 2: 5, 4: 7, 1: 9, 5: 12, 8: 15}
SLOTS = 18


def emulate(o, collide, src, nx, ny, nstep, nt, v, seg_cols, x_begin=0, x_end=None, va=None, halo_lo=None, halo_hi=None, mid=None):
    """One launch of k_lbmn_bulk<NSTEP = nstep, NT = nt, V = v> over columns [x_begin, x_end).  va = rows per 16 bytes (the
    alignment unit of the bulk copies; va = 2 v is the "wide" shape: half a 16-byte vector per thread).  halo_lo / halo_hi: the
    ring neighbours' three nearest lines in the order of csrc/plbm_internal.h ([3][9][ld]: lines -2, -1, -3 / nx, nx+1, nx+2).
    mid (the DUAL instances, a call's closing launch): an array that receives the state after step nstep - 1 as well."""
    x_end = nx if x_end is None else x_end
    va = v if va is None else va
    dst = np.full_like(src, np.nan)
    nr, w = nstep - 1, nt * v
    hs = -(-(nr * v + 1) // va) * va   # staged rows beyond the strip on each side (level 1 pulls one row beyond its own)
    off = hs - nr * v                  # stage row of thread 0's first row
    ws = w + 2 * off
    ty_max = (nt - 2 * nr) * v
    nstrips = -(-ny // ty_max)
    ty = -(-(-(-ny // nstrips)) // va) * va
    nstrips = -(-ny // ty)
    ncols = x_end - x_begin
    nseg = -(-ncols // seg_cols)
    seglen = -(-ncols // nseg)
    nseg = -(-ncols // seglen)
    tt = np.arange(nt)

    def collide_rows(n, act):
        """n[q, W] rows of all threads; collide the rows of the active threads in place (oracle arithmetic)."""
        rows = (tt[act][:, None] * v + np.arange(v)[None, :]).ravel()
        if rows.size == 0:
            return
        tmp = np.full((9, 1, padded_ld(rows.size)), np.nan, dtype=src.dtype)
        tmp[:, 0, :rows.size] = n[:, rows]
        collide(tmp, rows.size)
        n[:, rows] = tmp[:, 0, :rows.size]

    for seg in range(nseg):
        for strip in range(nstrips):
            y_lo = strip * ty
            y_hi = min(y_lo + ty, ny)
            xs = x_begin + seg * seglen
            xe = min(xs + seglen, x_end)
            yl = y_lo - nr * v + tt * v
            r0, r1 = y_lo - hs, y_hi + hs
            assert r1 - r0 <= ws
            ring = np.full((nr, SLOTS, w), np.nan, dtype=src.dtype)
            stage = np.full((2, 9, ws), np.nan, dtype=src.dtype)
            pending = {}  # stage -> raw column in flight (an mbarrier phase per use)

            def active(level):
                h = (nstep - level) * v
                return (yl >= y_lo - h) & (yl < y_hi + h)

            def issue(xl, s):
                assert s not in pending, "stage refilled before it was consumed"
                for q in range(9):
                    col = xl - CX[q]
                    if halo_lo is not None and col < 0:
                        assert -3 <= col
                        line = halo_lo[2 if col == -3 else col + 2, q]
                    elif halo_hi is not None and col >= nx:
                        assert col - nx < 3
                        line = halo_hi[col - nx, q]
                    else:
                        col = col + nx if col < 0 else (col - nx if col >= nx else col)
                        assert 0 <= col < nx
                        line = src[q, col]
                    d = stage[s, q]
                    d[:] = np.nan
                    nbytes = 0
                    if r0 < 0:
                        assert (ny + r0) % va == 0 and (-r0) % va == 0 and ny + r0 >= 0
                        d[0:-r0] = line[ny + r0:ny]
                        nbytes += -r0
                    m0, m1 = max(r0, 0), min(r1, ny)
                    assert m1 > m0 and m0 % va == 0 and (m1 - m0) % va == 0 and (m0 - r0) % va == 0
                    d[m0 - r0:m1 - r0] = line[m0:m1]
                    nbytes += m1 - m0
                    if r1 > ny:
                        assert (r1 - ny) % va == 0 and r1 - ny <= ny
                        d[ny - r0:r1 - r0] = line[0:r1 - ny]
                        nbytes += r1 - ny
                    assert nbytes == r1 - r0  # expect_tx of the kernel
                pending[s] = xl

            def pull(colof):
                """f[q, W]: rows t V + vv of population q come from colof(q)[t V + vv - cy]."""
                f = np.full((9, w), np.nan, dtype=src.dtype)
                for q in range(9):
                    c, off = colof(q)
                    idx = off + tt[:, None] * v + np.arange(v)[None, :] - CY[q]
                    ok = (idx >= 0) & (idx < c.shape[0])
                    vals = np.full(idx.shape, np.nan, dtype=src.dtype)
                    vals[ok] = c[idx[ok]]
                    f[q] = vals.ravel()
                return f

            x_first = xs - 2 * nr
            c_last = xe - 1 + nr
            issue(x_first + nr, 0)
            if x_first + nr + 1 <= c_last:
                issue(x_first + nr + 1, 1)
            w2 = w3 = 0
            for x in range(x_first, xe):
                k = x - x_first
                r2, r3 = w2 ^ 1, (0 if w3 == 2 else w3 + 1)
                wslot = lambda q: 0 if DEPTH[q] == 1 else (w2 if DEPTH[q] == 2 else w3)  # noqa: E731
                rslot = lambda q: 0 if DEPTH[q] == 1 else (r2 if DEPTH[q] == 2 else r3)  # noqa: E731
                # level 1
                assert pending.pop(k & 1) == x + nr, "wrong raw column in the stage"
                act = active(1)
                n = pull(lambda q: (stage[k & 1, q], off))
                collide_rows(n, act)
                rows = (tt[act][:, None] * v + np.arange(v)[None, :]).ravel()
                for q in range(9):
                    ring[0, BASE[q] + wslot(q), rows] = n[q, rows]
                if x + nr + 2 <= c_last:
                    issue(x + nr + 2, k & 1)
                for level in range(2, nstep + 1):
                    if x >= xs - 2 * (nstep - level):
                        act = active(level)
                        f = pull(lambda q: (ring[level - 2, BASE[q] + rslot(q)], 0))
                        collide_rows(f, act)
                        rows = (tt[act][:, None] * v + np.arange(v)[None, :]).ravel()
                        if level == nstep:
                            ylog = (yl[act][:, None] + np.arange(v)[None, :]).ravel()
                            assert ylog.min(initial=0) >= 0 and ylog.max(initial=0) < ny
                            for q in range(9):
                                assert np.isnan(dst[q, x, ylog]).all(), "a node written twice"
                                dst[q, x, ylog] = f[q, rows]
                        else:
                            for q in range(9):
                                ring[level - 1, BASE[q] + wslot(q), rows] = f[q, rows]
                            if mid is not None and level == nstep - 1 and xs <= x + 1 < xe:  # column x + 1, the strip's own rows
                                own = active(nstep)
                                rows_own = (tt[own][:, None] * v + np.arange(v)[None, :]).ravel()
                                ylog = (yl[own][:, None] + np.arange(v)[None, :]).ravel()
                                for q in range(9):
                                    assert np.isnan(mid[q, x + 1, ylog]).all(), "a node of the second output written twice"
                                    mid[q, x + 1, ylog] = f[q, rows_own]
                w2 ^= 1
                w3 = 0 if w3 == 2 else w3 + 1
            assert not pending, "a bulk copy was still in flight when the block ended"
    return dst


def reference(o, collide, f0, nx, ny, nstep):
    a, b = f0.copy(), np.full_like(f0, np.nan)
    for _ in range(nstep):
        o.lbm_stream(a, b, ny)
        collide(b, ny)
        a, b = b, a
    return a


def collisions(o, p):
    return {"bgk": lambda f, n: o.collide_bgk(f, n, p["omega"]), "trt": lambda f, n: o.collide_trt(f, n, p["omega"], p["trt_magic"]),
            "rr": lambda f, n: o.collide_rr(f, n, p["omega"])}


@pytest.mark.parametrize("nstep", [2, 3])
@pytest.mark.parametrize("nx,ny,nt,v,seg_cols", [
    (7, 16, 16, 2, 64),     # one strip that wraps onto itself, one segment
    (9, 40, 10, 2, 4),      # several strips, ragged last strip, segments of 3 columns
    (5, 24, 8, 2, 1),       # one-column segments: every column pays the whole warm-up
    (6, 32, 8, 4, 3),       # fp32-like vectors of four rows
    (4, 44, 12, 2, 2),      # nx = 4: the raw columns wrap around x in both directions
])
def test_schedule_of_the_multi_step_kernels(nstep, nx, ny, nt, v, seg_cols):
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    f0 = random_state(o, nx, ny)
    for name, collide in collisions(o, p).items():
        got = emulate(o, collide, f0, nx, ny, nstep, nt, v, seg_cols)
        want = reference(o, collide, f0, nx, ny, nstep)
        assert np.array_equal(got[:, :, :ny], want[:, :, :ny]), (name, nstep)


@pytest.mark.parametrize("nstep", [2, 3])
@pytest.mark.parametrize("nx,ny,nt,v,va,seg_cols", [
    (7, 16, 16, 1, 2, 64),    # fp64, one row per thread (the default shape of the three-step kernel): 16-byte pieces of two rows
    (9, 40, 12, 1, 2, 4),     # several strips, ragged last strip
    (6, 32, 10, 2, 4, 3),     # fp32, two rows per thread, 16-byte pieces of four rows
    (5, 24, 8, 1, 2, 1),      # one-column segments
])
def test_schedule_of_the_wide_shape(nstep, nx, ny, nt, v, va, seg_cols):
    """half a 16-byte vector per thread: the staged range is rounded out to 16 bytes and thread 0 starts inside it"""
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    f0 = random_state(o, nx, ny)
    for name, collide in collisions(o, p).items():
        got = emulate(o, collide, f0, nx, ny, nstep, nt, v, seg_cols, va=va)
        want = reference(o, collide, f0, nx, ny, nstep)
        assert np.array_equal(got[:, :, :ny], want[:, :, :ny]), (name, nstep)


@pytest.mark.parametrize("nxl", [6, 7, 12])
def test_three_step_schedule_of_a_slab(nxl):
    """The slab schedule of a triple: lines [0, 3) and [nxl - 3, nxl) read the neighbours' three halo lines (stored in the order
    -2, -1, -3 / nx, nx+1, nx+2), the interior [3, nxl - 3) reads the slab alone; together they cover the slab once and equal
    the middle slab of a three-slab periodic grid stepped as a whole."""
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    ny, nstep = 24, 3
    nxg = 3 * nxl
    f0 = random_state(o, nxg, ny)
    collide = collisions(o, p)["trt"]
    want = reference(o, collide, f0, nxg, ny, nstep)[:, nxl:2 * nxl]
    slab = np.ascontiguousarray(f0[:, nxl:2 * nxl])
    halo_lo = np.stack([f0[:, nxl - 2], f0[:, nxl - 1], f0[:, nxl - 3]])          # lines -2, -1, -3
    halo_hi = np.stack([f0[:, 2 * nxl], f0[:, 2 * nxl + 1], f0[:, 2 * nxl + 2]])  # lines nx, nx+1, nx+2
    got = np.full_like(slab, np.nan)
    ranges = ((0, nxl),) if nxl <= 6 else ((0, 3), (nxl - 3, nxl), (3, nxl - 3))
    for x0, x1 in ranges:
        boundary = x0 == 0 or x1 == nxl
        part = emulate(o, collide, slab, nxl, ny, nstep, 12, 1, 64, x0, x1, va=2,
                       halo_lo=halo_lo if boundary else None, halo_hi=halo_hi if boundary else None)
        assert np.isnan(got[:, x0:x1, :ny]).all()
        got[:, x0:x1] = part[:, x0:x1]
    assert np.array_equal(got[:, :, :ny], want[:, :, :ny])


@pytest.mark.parametrize("nxl", [6, 12])
def test_dual_output_of_the_closing_triple_on_a_slab(nxl):
    """k_lbmn_bulk<HALO, DUAL>, the boundary launches of a call's closing triple under slabs (and the whole-slab launch of a slab of
    six lines): besides the state after step 3 they store the state after step 2 -- lattice `inew` of the reference -- for exactly
    the columns and rows they own."""
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    ny = 24
    nxg = 3 * nxl
    f0 = random_state(o, nxg, ny)
    collide = collisions(o, p)["bgk"]
    want3, want2 = reference(o, collide, f0, nxg, ny, 3)[:, nxl:2 * nxl], reference(o, collide, f0, nxg, ny, 2)[:, nxl:2 * nxl]
    slab = np.ascontiguousarray(f0[:, nxl:2 * nxl])
    halo_lo = np.stack([f0[:, nxl - 2], f0[:, nxl - 1], f0[:, nxl - 3]])
    halo_hi = np.stack([f0[:, 2 * nxl], f0[:, 2 * nxl + 1], f0[:, 2 * nxl + 2]])
    got, mid = np.full_like(slab, np.nan), np.full_like(slab, np.nan)
    ranges = ((0, nxl),) if nxl <= 6 else ((0, 3), (nxl - 3, nxl))
    for x0, x1 in ranges:
        part = emulate(o, collide, slab, nxl, ny, 3, 12, 1, 64, x0, x1, va=2, halo_lo=halo_lo, halo_hi=halo_hi, mid=mid)
        got[:, x0:x1] = part[:, x0:x1]
    for x0, x1 in ranges:
        assert np.array_equal(got[:, x0:x1, :ny], want3[:, x0:x1, :ny]) and np.array_equal(mid[:, x0:x1, :ny], want2[:, x0:x1, :ny])
    if nxl > 6:
        assert np.isnan(mid[:, 3:nxl - 3, :ny]).all()  # the interior is the other kernel's (tests/test_ws_schedule.py)


def test_schedule_on_a_line_sub_range():
    """the slab schedule's three launches (two boundary lines per side, then the interior) cover the grid once"""
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    nx, ny = 12, 24
    f0 = random_state(o, nx, ny)
    collide = collisions(o, p)["bgk"]
    want = reference(o, collide, f0, nx, ny, 2)
    got = np.full_like(f0, np.nan)
    for x0, x1 in ((0, 2), (nx - 2, nx), (2, nx - 2)):
        part = emulate(o, collide, f0, nx, ny, 2, 8, 2, 3, x0, x1)
        assert np.isnan(got[:, x0:x1, :ny]).all()
        got[:, x0:x1] = part[:, x0:x1]
    assert np.array_equal(got[:, :, :ny], want[:, :, :ny])
