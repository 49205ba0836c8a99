"""Host-side output staging for the device-resident grid: the drivers' field dumps
and a PDF checkpoint the reference lacks.  Pure host I/O on data the C ABI hands back; nothing here
computes.

  set_output_folder(grid, foldername)   <-> lattice_grid%set_output_folder   (src/fvm_bardow.F90:999-1025)
  output_npy(grid, step)                <-> output_npy                       (src/fvm_bardow.F90:929-958,
                                            src/output/npy.f90:14-29): <folder>/<filename><step:09d>.npy holding
                                            mf(ny,nx,3) in Fortran order, byte-compatible with stdlib's save_npy
  output_vtk(grid, step)                <-> output_vtk                       (src/fvm_bardow.F90:960-997,
                                            src/output/vtk.F90:150-196): ASCII STRUCTURED_POINTS, cell data
  output_gnuplot(grid, step)            <-> output_gnuplot                   (src/fvm_bardow.F90:895-925,
                                            src/output/gnuplot.F90:21-39): "x y rho ux uy" blocks per line x
  save_checkpoint / load_checkpoint     PDFs of every lattice + indices + properties (new)

Numbers are written with the reference's edit descriptors (ES24.16E3 in double, ES15.8E2 in single precision).
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


def _fmt_real(a) -> list:
    """Fortran ES24.16E3 (float64) / ES15.8E2 (float32) renderings of the values of `a` (1-D)."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        return ["%15.8E" % v for v in a]
    out = []
    for v in a:
        m, e = ("%.16E" % v).split("E")  # C prints at least two exponent digits; the reference's descriptor has three
        out.append(f"{m}E{e[0]}{int(e[1:]):03d}".rjust(24))
    return out


def output_vtk(grid: LatticeGrid, step=None, binary: bool = False):
    """ASCII legacy-VTK STRUCTURED_POINTS file with Density and Velocity as cell data, like the reference's
    output_vtk -> output_vtk_structuredPoints.  binary=True prints the reference's message and returns None."""
    if binary:
        print(" binary output not implemented")  # src/fvm_bardow.F90:984-986
        return None
    nx, ny = grid.nx, grid.ny
    name = _fullname(grid, step, ".vtk")
    zero = _fmt_real(np.zeros(1, grid.dtype))[0]
    one = _fmt_real(np.ones(1, grid.dtype))[0]
    rho = _fmt_real(np.asarray(grid.rho).T.ravel())  # do j = 1, ny; do i = 1, nx: rho(j,i)
    ux = _fmt_real(np.asarray(grid.ux).T.ravel())
    uy = _fmt_real(np.asarray(grid.uy).T.ravel())
    with open(name, "w") as fh:
        fh.write("# vtk DataFile Version 3.0\nfluid\nASCII\nDATASET STRUCTURED_POINTS\n")
        fh.write(f"DIMENSIONS {nx + 1} {ny + 1} 2 \n")
        fh.write("ORIGIN  " + zero * 3 + "\n")
        fh.write("SPACING " + one * 3 + "\n")
        fh.write("\n")
        fh.write(f"CELL_DATA {nx * ny}\n")
        fh.write("SCALARS Density float 1\nLOOKUP_TABLE default\n")
        fh.write("\n".join(rho) + "\n")
        fh.write("\n")
        fh.write("VECTORS Velocity float\n")
        fh.write("\n".join(a + b + zero for a, b in zip(ux, uy)) + "\n")
    return name


def output_gnuplot(grid: LatticeGrid, step=None) -> str:
    """One "x y rho ux uy" record per node, a blank line after every line x (gnuplot's grid format).  The
    reference's writer never assigns its `ry` (src/output/gnuplot.F90:31-33 assigns `rx` twice), so its second
    column is undefined; here the columns are the cell centres the code evidently intends, (x - 1/2, y - 1/2)."""
    nx, ny = grid.nx, grid.ny
    name = _fullname(grid, step, ".txt")
    yy = _fmt_real((np.arange(ny) + 0.5).astype(grid.dtype))
    with open(name, "w") as fh:
        for x in range(nx):
            rx = _fmt_real(np.asarray([x + 0.5], grid.dtype))[0]
            cols = [_fmt_real(np.asarray(f)[x]) for f in (grid.rho, grid.ux, grid.uy)]
            fh.write("".join(f"{rx} {yy[y]} {cols[0][y]} {cols[1][y]} {cols[2][y]}\n" for y in range(ny)))
            fh.write("\n")
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
    # lattices are stored by NUMBER: restore the saved roles first (a fresh grid starts with inew=1, iold=2, imid=3;
    # after perform_triple_step the roles rotate through all three lattices, which plbm_swap alone cannot reach)
    from .capi import check, lib
    check(lib.plbm_set_indices(g._h, int(z["iold"]), int(z["inew"]), int(z["imid"])), "set_indices")
    if (g.iold, g.inew, g.imid) != (int(z["iold"]), int(z["inew"]), int(z["imid"])):
        raise RuntimeError("load_checkpoint: lattice indices were not restored")
    for k in range(1, g.nf + 1):
        g.upload_f(k, z[f"f{k}"])
    return g
