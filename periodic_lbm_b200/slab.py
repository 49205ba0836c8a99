"""Slab decomposition of the periodic grid along the slow index (Fortran `x`, stride ld) over the
GPUs of one box -- host-side bookkeeping only (the reference is single-process).

Rank r owns lines [x_offset, x_offset + nx_local) of every population and forms a periodic ring
with its neighbours.  Per LBM step it needs the last line of q = 1,5,8 (cx = +1) from rank r-1
and the first line of q = 3,6,7 (cx = -1) from rank r+1; the y shift of the diagonal populations
is applied locally on the received line, so no corner exchange exists.
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


def halo_message_bytes(ny: int, itemsize: int) -> int:
    """bytes per direction per step: 3 populations x ld reals (ld = ny padded to 16)."""
    return 3 * ((ny + 15) // 16 * 16) * itemsize
