"""Host-side output staging for the device-resident grid: the drivers' field dumps
and a PDF checkpoint the reference lacks.  Pure host I/O on data the C ABI hands back; nothing here
computes.

  set_output_folder(grid, foldername)   <-> lattice_grid%set_output_folder   (src/fvm_bardow.F90:999-1025)
  output_npy(grid, step)                <-> output_npy                       (src/fvm_bardow.F90:929-958,
                                            src/output/npy.f90:14-29): <folder>/<filename><step:09d>.npy holding
                                            mf(ny,nx,3) in Fortran order, byte-compatible with stdlib's save_npy
  save_checkpoint / load_checkpoint     PDFs of every lattice + indices + properties (new)

The gnuplot / VTK text writers of src/output/ stay with the reference (call update_macros first).
"""
from __future__ import annotations

import os

import numpy as np

from .lattice import LatticeGrid, alloc_grid, set_properties


def set_output_folder(grid: LatticeGrid, foldername: str, verbose: bool = False) -> None:
    os.makedirs(foldername, exist_ok=True)
    if verbose:
        print(f"mkdir: created directory '{foldername}'")
    grid.foldername = foldername


def _fullname(grid: LatticeGrid, step, ext: str) -> str:
    istr = "" if step is None else f"{int(step):09d}"
    folder = grid.foldername or ""
    if folder:
        os.makedirs(folder, exist_ok=True)
    return f"{folder}/{grid.filename or 'results'}{istr}{ext}"


def output_npy(grid: LatticeGrid, step=None) -> str:
    """Writes grid.rho/ux/uy (as last filled by update_macros) like the reference's output_npy."""
    mf = np.empty((grid.ny, grid.nx, 3), dtype=grid.dtype, order="F")
    mf[:, :, 0], mf[:, :, 1], mf[:, :, 2] = grid.rho.T, grid.ux.T, grid.uy.T
    name = _fullname(grid, step, ".npy")
    np.save(name, mf)
    return name


def save_checkpoint(grid: LatticeGrid, path: str) -> None:
    """PDF checkpoint: every lattice (f(ld,nx,0:8) each), the lattice indices and the properties."""
    lattices = {f"f{k}": grid.download_f(k) for k in range(1, grid.nf + 1)}  # download_f(inew) materialises fbar+ after DUGKS
    np.savez(path, nx=grid.nx, ny=grid.ny, nf=grid.nf, precision=np.dtype(grid.dtype).name, iold=grid.iold, inew=grid.inew, imid=grid.imid,
             props=np.array([float(x) for x in grid._props()]), **lattices)


def load_checkpoint(path: str, device=None) -> LatticeGrid:
    z = np.load(path if path.endswith(".npz") else path + ".npz")
    prec = "f64" if str(z["precision"]) == "float64" else "f32"
    g = alloc_grid(int(z["nx"]), int(z["ny"]), nf=int(z["nf"]), precision=prec, device=device)
    nu, dt, tau, omega, magic, _ = z["props"]
    set_properties(g, nu, dt, magic)
    g.omega = omega
    # a fresh grid starts with inew=1, iold=2 (alloc_grid); put each saved lattice where the saved
    # indices expect it after the same number of swaps modulo the cycle
    saved = {"iold": int(z["iold"]), "inew": int(z["inew"]), "imid": int(z["imid"])}
    from .capi import check, lib
    guard = 0
    while (g.iold, g.inew) != (saved["iold"], saved["inew"]) and guard < 3:
        check(lib.plbm_swap(g._h), "swap")
        guard += 1
    if g.nf == 2 and (g.iold, g.inew) != (saved["iold"], saved["inew"]):
        raise RuntimeError("load_checkpoint: cannot restore lattice indices")
    for k in range(1, g.nf + 1):
        g.upload_f(k, z[f"f{k}"])
    return g
