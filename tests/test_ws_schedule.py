"""Executable specification of the schedule of the warp-specialised, skewed three-step kernel (csrc/plbm_lbm3w.cu k_lbm3_ws),
run on the CPU against the oracle -- the companion of test_multi_step_schedule.py for the kernel that replaced k_lbmn_bulk in the
launches that read no halo lines.

`emulate_ws()` restates in numpy what ONE launch does block by block: the producer's running line addresses (one population per
lane, wrap around x), the staged periodic pieces, the stage / mbarrier-phase bookkeeping, the 27-slot rings (populations kept
2 / 3 / 4 columns), the skew of the levels (columns j, j - 2, j - 4 in iteration j), the gates of the ramp-up and ramp-down
iterations, the per-level store predicates, the running output pointers and the optional second output (state after step 2).
Every ring slot remembers which column it holds, so a slot that is overwritten before its last reader, or read before it is
written, fails the test; values come out bit for bit equal to three oracle steps.  What it cannot see: barriers, the async proxy,
named-barrier counts -- those are the -m gpu parity tests' job.  It does see what compute-sanitizer racecheck sees inside an
iteration: a word that one thread reads and another writes without a barrier in between (the reason for the pad rows of a slot)."""
import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle, padded_ld
from test_multi_step_schedule import CX, CY, collisions, reference

DEPTH = [2 if c == -1 else (3 if c == 0 else 4) for c in CX]
BASE = {3: 0, 6: 2, 7: 4, 0: 6, 2: 9, 4: 12, 1: 15, 5: 19, 8: 23}
SLOTS = 27


def emulate_ws(o, collide, src, nx, ny, ntc, v, seg_cols, x_begin=0, x_end=None, va=None, dual=False):
    """One launch of k_lbm3_ws<NTC = ntc, V = v> over columns [x_begin, x_end); returns (dst, dst_mid)."""
    x_end = nx if x_end is None else x_end
    va = 2 * v if va is None else va
    dst = np.full_like(src, np.nan)
    mid = np.full_like(src, np.nan)
    nr, w = 2, ntc * v
    hs = -(-(nr * v + 1) // va) * va
    off = hs - nr * v
    ws = w + 2 * off
    ty_max = (ntc - 4) * v
    nstrips = -(-ny // ty_max)
    ty = -(-(-(-ny // nstrips)) // va) * va
    nstrips = -(-ny // ty)
    ncols = x_end - x_begin
    nseg = -(-ncols // seg_cols)
    seglen = -(-ncols // nseg)
    if seglen < 8:
        seglen = ncols if ncols < 8 else 8
    nseg = -(-ncols // seglen)
    tt = np.arange(ntc)

    def collide_rows(n):
        """every thread collides its rows, active or not (the kernel predicates the stores only)"""
        tmp = np.full((9, 1, padded_ld(w)), 1.0, dtype=src.dtype)
        tmp[:, 0, :w] = n
        collide(tmp, w)
        return tmp[:, 0, :w].copy()

    def rows_of(act):
        return (tt[act][:, None] * v + np.arange(v)[None, :]).ravel()

    for seg in range(nseg):
        for strip in range(nstrips):
            y_lo = strip * ty
            y_hi = min(y_lo + ty, ny)
            xs = x_begin + seg * seglen
            xe = min(xs + seglen, x_end)
            yl = y_lo - nr * v + tt * v
            r0, r1 = y_lo - hs, y_hi + hs
            assert r1 - r0 <= ws
            a1 = yl < y_hi + 2 * v
            a2 = (yl >= y_lo - v) & (yl < y_hi + v)
            a3 = (yl >= y_lo) & (yl < y_hi)
            # shared memory starts as ones (finite, positive): [stage | ring 0 | ring 1], flat like in the kernel; a ring slot is
            # v pad rows + the w rows of the threads + v pad rows (wp rows), ring0 = the first thread row of slot 0
            wp = w + 2 * v
            smem = np.full(2 * 9 * ws + 2 * SLOTS * wp, 1.0, dtype=src.dtype)
            stage0, ring0 = 0, 2 * 9 * ws + v
            held = {}     # (ring, slot index) -> column it holds
            pending = {}  # stage -> raw column in flight
            x_first = xs - 2
            n_raw = xe + 2 - x_first
            n_iter = n_raw + 2
            # producer: lane q keeps the column of population q
            ncol = [(x_first - CX[q]) % nx if -nx <= x_first - CX[q] < 2 * nx else None for q in range(9)]
            assert None not in ncol
            for q in range(9):  # the kernel wraps once, by +-nx
                c = x_first - CX[q]
                c = c + nx if c < 0 else (c - nx if c >= nx else c)
                assert c == ncol[q]
            issued = [0]

            def issue_next(s):
                assert s not in pending, "stage refilled before it was consumed"
                xl = x_first + issued[0]
                for q in range(9):
                    assert ncol[q] == (xl - CX[q]) % nx
                    line = src[q, ncol[q]]
                    d0 = stage0 + (s * 9 + q) * ws
                    nbytes = 0
                    if r0 < 0:
                        assert (ny + r0) % va == 0 and (-r0) % va == 0 and ny + r0 >= 0
                        smem[d0:d0 - r0] = line[ny + r0:ny]
                        nbytes += -r0
                    m0, m1 = max(r0, 0), min(r1, ny)
                    assert m1 > m0 and m0 % va == 0 and (m1 - m0) % va == 0 and (m0 - r0) % va == 0
                    smem[d0 + m0 - r0:d0 + m1 - r0] = line[m0:m1]
                    nbytes += m1 - m0
                    if r1 > ny:
                        assert (r1 - ny) % va == 0 and r1 - ny <= ny
                        smem[d0 + ny - r0:d0 + r1 - r0] = line[0:r1 - ny]
                        nbytes += r1 - ny
                    assert nbytes == r1 - r0
                    ncol[q] = 0 if ncol[q] + 1 == nx else ncol[q] + 1
                pending[s] = xl
                issued[0] += 1

            reads = set()  # shared-memory addresses read in the current iteration, by any thread (halo threads included)

            def pull(base_of):
                """f[q, W]: rows t V + vv of population q from flat address base_of(q) + t V + vv - cy (always inside smem)"""
                f = np.empty((9, w), dtype=src.dtype)
                for q in range(9):
                    idx = base_of(q) + tt[:, None] * v + np.arange(v)[None, :] - CY[q]
                    assert idx.min() >= 0 and idx.max() < smem.size, "a halo thread reads outside the block's shared memory"
                    f[q] = smem[idx].ravel()
                    reads.update(idx.ravel().tolist())
                return f

            def store(base, rows, vals):
                """predicated ring store; no thread may have read these words in this iteration (there is no barrier in between:
                compute-sanitizer racecheck flags it even when the reader is a halo thread that throws its result away)"""
                assert reads.isdisjoint((base + rows).tolist()), "a word is read and written in the same iteration"
                smem[base + rows] = vals

            issue_next(0)
            if n_raw > 1:
                issue_next(1)
            w2 = w3 = w4 = 0
            for k in range(n_iter):
                j = x_first + k
                r2, r3, r4 = w2 ^ 1, (0 if w3 == 2 else w3 + 1), (w4 + 1) & 3
                wslot = lambda q: BASE[q] + (w2 if DEPTH[q] == 2 else (w3 if DEPTH[q] == 3 else w4))  # noqa: E731
                rslot = lambda q: BASE[q] + (r2 if DEPTH[q] == 2 else (r3 if DEPTH[q] == 3 else r4))  # noqa: E731
                l1, l2, l3 = k < n_raw, xs + 1 <= j <= xe + 2, j >= xs + 4
                assert l1 or l2 or l3
                # ---- loads of the iteration (all before any store of it: the levels are independent)
                reads.clear()
                n1 = n2 = n3 = None
                if l1:
                    assert pending.pop(k & 1) == j, "wrong raw column in the stage"
                    n1 = pull(lambda q: stage0 + ((k & 1) * 9 + q) * ws + off)
                if l2:
                    for q in range(9):
                        assert held[(0, rslot(q))] == j - 2 - CX[q], "ring 0 slot does not hold the column level 2 pulls"
                    n2 = pull(lambda q: ring0 + rslot(q) * wp)
                if l3:
                    for q in range(9):
                        assert held[(1, rslot(q))] == j - 4 - CX[q], "ring 1 slot does not hold the column level 3 pulls"
                    n3 = pull(lambda q: ring0 + (SLOTS + rslot(q)) * wp)
                # the producer refills the stage once every consumer has read it
                if k + 2 < n_raw:
                    issue_next(k & 1)
                # ---- collisions and predicated stores
                if l1:
                    n1 = collide_rows(n1)
                    rows = rows_of(a1)
                    for q in range(9):
                        store(ring0 + wslot(q) * wp, rows, n1[q, rows])
                        held[(0, wslot(q))] = j
                if l2:
                    n2 = collide_rows(n2)
                    rows = rows_of(a2)
                    for q in range(9):
                        store(ring0 + (SLOTS + wslot(q)) * wp, rows, n2[q, rows])
                        held[(1, wslot(q))] = j - 2
                    if dual and xs <= j - 2 < xe:
                        rows, ylog = rows_of(a3), (yl[a3][:, None] + np.arange(v)[None, :]).ravel()
                        for q in range(9):
                            assert np.isnan(mid[q, j - 2, ylog]).all(), "a node of the second output written twice"
                            mid[q, j - 2, ylog] = n2[q, rows]
                if l3:
                    n3 = collide_rows(n3)
                    rows, ylog = rows_of(a3), (yl[a3][:, None] + np.arange(v)[None, :]).ravel()
                    assert xs <= j - 4 < xe and ylog.min(initial=0) >= 0 and ylog.max(initial=0) < ny
                    for q in range(9):
                        assert np.isnan(dst[q, j - 4, ylog]).all(), "a node written twice"
                        dst[q, j - 4, ylog] = n3[q, rows]
                w2 ^= 1
                w3 = 0 if w3 == 2 else w3 + 1
                w4 = (w4 + 1) & 3
            assert not pending, "a bulk copy was still in flight when the block ended"
            assert issued[0] == n_raw
    return dst, mid


@pytest.mark.parametrize("nx,ny,ntc,v,seg_cols", [
    (7, 16, 16, 1, 64),     # fp64 shape: one row per thread, one strip that wraps onto itself, one segment
    (19, 40, 12, 1, 8),     # several strips, ragged last strip, several segments
    (9, 24, 8, 1, 64),      # nx barely above the ramp
    (17, 32, 10, 2, 8),     # fp32 shape: two rows per thread, 16-byte pieces of four rows
    (4, 44, 12, 1, 64),     # nx = 4: the raw columns wrap around x more than once
    (5, 16, 8, 1, 64),
])
def test_schedule_of_the_warp_specialised_kernel(nx, ny, ntc, v, seg_cols):
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    f0 = random_state(o, nx, ny)
    for name, collide in collisions(o, p).items():
        got, _ = emulate_ws(o, collide, f0, nx, ny, ntc, v, seg_cols)
        want = reference(o, collide, f0, nx, ny, 3)
        assert np.array_equal(got[:, :, :ny], want[:, :, :ny]), name


def test_second_output_is_the_state_after_two_steps():
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    nx, ny = 21, 40
    f0 = random_state(o, nx, ny)
    collide = collisions(o, p)["trt"]
    got, mid = emulate_ws(o, collide, f0, nx, ny, 12, 1, 8, dual=True)
    assert np.array_equal(got[:, :, :ny], reference(o, collide, f0, nx, ny, 3)[:, :, :ny])
    assert np.array_equal(mid[:, :, :ny], reference(o, collide, f0, nx, ny, 2)[:, :, :ny])


@pytest.mark.parametrize("nx,x0,x1", [(8, 3, 5), (7, 3, 4), (12, 3, 9), (30, 3, 27)])
def test_short_interior_ranges(nx, x0, x1):
    """thin slabs: the interior between the two three-line boundaries is a column or two -- ramp iterations only"""
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    ny = 24
    f0 = random_state(o, nx, ny)
    collide = collisions(o, p)["rr"]
    want3, want2 = reference(o, collide, f0, nx, ny, 3), reference(o, collide, f0, nx, ny, 2)
    got, mid = emulate_ws(o, collide, f0, nx, ny, 16, 1, 8, x0, x1, dual=True)
    assert np.array_equal(got[:, x0:x1, :ny], want3[:, x0:x1, :ny]) and np.array_equal(mid[:, x0:x1, :ny], want2[:, x0:x1, :ny])
    assert np.isnan(got[:, :x0, :ny]).all() and np.isnan(got[:, x1:, :ny]).all() and np.isnan(mid[:, :x0, :ny]).all()


def test_interior_range_of_a_slab():
    """the interior launch of a slab (columns [3, nx - 3)) leaves the boundary columns alone and equals the whole-grid result there"""
    o = Oracle("f64")
    p = o.set_properties(0.02, 1.0, 0.25)
    nx, ny = 20, 24
    f0 = random_state(o, nx, ny)
    collide = collisions(o, p)["bgk"]
    want = reference(o, collide, f0, nx, ny, 3)
    got, _ = emulate_ws(o, collide, f0, nx, ny, 16, 1, 8, 3, nx - 3)
    assert np.array_equal(got[:, 3:nx - 3, :ny], want[:, 3:nx - 3, :ny])
    assert np.isnan(got[:, :3, :ny]).all() and np.isnan(got[:, nx - 3:, :ny]).all()
