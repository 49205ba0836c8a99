"""Slab decomposition of the periodic grid along the slow index (Fortran `x`, stride ld) over the
GPUs of one box -- host-side bookkeeping only (the reference is single-process).

Rank r owns lines [x_offset, x_offset + nx_local) of every population and forms a periodic ring
with its neighbours.  One LBM step needs the last line of q = 1,5,8 (cx = +1) from rank r-1
and the first line of q = 3,6,7 (cx = -1) from rank r+1; a fused pair of steps (the library's
two-step kernel) recomputes step 1 on the neighbours' nearest line and therefore needs their TWO
nearest lines of all nine populations, a fused triple (the three-step kernel, the default) their
THREE nearest lines.  The halo message always carries HALO_LINES = 3 lines x 9 populations per
direction, so that any kind of launch can follow; the y shift of the diagonal populations is
applied locally on the received lines, so no corner exchange exists.
"""
from __future__ import annotations

from dataclasses import dataclass

Q_FROM_LO = (1, 5, 8)  # populations moving in +x: received from the low-x neighbour
Q_FROM_HI = (3, 6, 7)  # populations moving in -x: received from the high-x neighbour


@dataclass(frozen=True)
class Slab:
    rank: int
    nranks: int
    nx_global: int
    x_offset: int
    nx_local: int

    @property
    def lo(self) -> int:
        return (self.rank - 1) % self.nranks

    @property
    def hi(self) -> int:
        return (self.rank + 1) % self.nranks

    @property
    def x_end(self) -> int:
        return self.x_offset + self.nx_local


def slab_of(rank: int, nranks: int, nx_global: int) -> Slab:
    """Balanced contiguous partition: the first nx_global % nranks ranks get one extra line."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    if nx_global < 2 * nranks:
        raise ValueError("every slab needs at least 2 lines")
    base, rem = divmod(nx_global, nranks)
    nx_local = base + (1 if rank < rem else 0)
    x_offset = rank * base + min(rank, rem)
    return Slab(rank, nranks, nx_global, x_offset, nx_local)


HALO_LINES = 3  # lines per direction in one halo message


def halo_message_bytes(ny: int, itemsize: int) -> int:
    """bytes per direction per launch: HALO_LINES lines x 9 populations x ld reals (ld = ny padded to 16)."""
    return HALO_LINES * 9 * ((ny + 15) // 16 * 16) * itemsize


def launch_schedule(nsteps: int, pairs: bool = True, triples: bool = False, dual: bool = False) -> list:
    """Steps advanced by each launch of one perform_lbm_step(nsteps) call: fused triples while more than three
    steps remain (where the ring takes them: every slab at least 2 * HALO_LINES lines wide), fused pairs while
    more than two remain (the last step stays single so that lattice `inew` ends up holding state nsteps-1
    exactly like the reference), then single steps.

    dual (with triples): every rank holds a third lattice buffer, so the call may CLOSE with a triple that stores the states
    after its second and third step (csrc/plbm_internal.h lbm_next_launch): 3 left -> that triple; 5 left -> a pair first;
    4 left -> triple + single; 6 or more -> a triple.  The closing dual triple is listed as 3 like any other."""
    out, s = [], 0
    while dual and triples and s < nsteps:
        rem = nsteps - s
        n = 3 if rem == 3 else (2 if rem == 5 and pairs else (3 if rem >= 4 else 1))
        out.append(n)
        s += n
    while s < nsteps:
        rem = nsteps - 1 - s  # steps before the closing single step
        if triples and rem >= 3 and not (pairs and rem == 4):  # four remaining steps go as two pairs, not triple + single
            n = 3
        else:
            n = 2 if pairs and rem >= 2 else 1
        out.append(n)
        s += n
    return out
