"""ctypes stub of the C ABI in include/plbm.h (the same entry points the Fortran bind(c)
interface block in fortran/plbm_c.f90 declares).

No compute happens in Python and there is no fallback: if libplbm_b200.so is missing the
import of this module raises, and every compute entry point raises PlbmError when no CUDA
device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplbm_b200.so")

F64, F32 = 0, 1
BGK, TRT, RR, BGK_SPLIT, TRT_SPLIT, BGK_IMPROVED = 0, 1, 2, 3, 4, 5
STREAM_LBM, STREAM_FVM_BARDOW, STREAM_FDM_BARDOW, STREAM_FDM_SOFONEA = 0, 1, 2, 3
DIAG_MAX_SPEED, DIAG_MIN_SPEED, DIAG_SUM_RHO, DIAG_KINETIC, DIAG_COUNT = 0, 1, 2, 3, 4

# name -> (restype, argtypes); must list every symbol include/plbm.h declares
_H, _I, _D, _P = C.c_void_p, C.c_int, C.c_double, C.c_void_p
SIGNATURES = {
    "plbm_last_error": (C.c_char_p, []),
    "plbm_version": (_I, []),
    "plbm_device_count": (_I, []),
    "plbm_alloc_grid": (_I, [C.POINTER(_H), _I, _I, _I, _I]),
    "plbm_alloc_grid_on": (_I, [C.POINTER(_H), _I, _I, _I, _I, _I]),
    "plbm_dealloc_grid": (_I, [_H]),
    "plbm_get_dims": (_I, [_H] + [C.POINTER(_I)] * 5),
    "plbm_get_indices": (_I, [_H] + [C.POINTER(_I)] * 3),
    "plbm_set_indices": (_I, [_H, _I, _I, _I]),
    "plbm_set_properties": (_I, [_H, _D, _D, _D, _I]),
    "plbm_get_properties": (_I, [_H, C.POINTER(_D)]),
    "plbm_set_omega": (_I, [_H, _D]),
    "plbm_set_pdf_to_equilibrium": (_I, [_H, _P, _P, _P]),
    "plbm_perform_lbm_step": (_I, [_H, _I, _I]),
    "plbm_perform_step": (_I, [_H, _I, _I, _I]),
    "plbm_perform_triple_step": (_I, [_H, _I, _I, _I]),
    "plbm_perform_dugks_step": (_I, [_H, _I, _I]),
    "plbm_lbm_stream": (_I, [_H]),
    "plbm_stream_fvm_bardow": (_I, [_H]),
    "plbm_stream_fdm_bardow": (_I, [_H]),
    "plbm_stream_fdm_sofonea": (_I, [_H]),
    "plbm_collide": (_I, [_H, _I]),
    "plbm_dugks_collide": (_I, [_H, _I]),
    "plbm_dugks_stream": (_I, [_H, _I]),
    "plbm_swap": (_I, [_H]),
    "plbm_update_macros": (_I, [_H, _P, _P, _P, _I]),
    "plbm_vorticity": (_I, [_H, _I, _P]),
    "plbm_vorticity_host": (_I, [_H, _I, _P, _P, _P]),
    "plbm_diagnostics": (_I, [_H, C.POINTER(_D)]),
    "plbm_l2_sums": (_I, [_H, _P, _P, C.POINTER(_D)]),
    "plbm_lattice_hash": (_I, [_H, _I, C.POINTER(C.c_ulonglong)]),
    "plbm_upload_f": (_I, [_H, _I, _P]),
    "plbm_download_f": (_I, [_H, _I, _P]),
    "plbm_set_stream": (_I, [_H, _P]),
    "plbm_synchronize": (_I, [_H]),
    "plbm_launch_count": (C.c_longlong, []),
    "plbm_set_variant": (_I, [_H, _I]),
    "plbm_set_step_deferral": (_I, [_H, _I]),
    "plbm_lbm_pair_kernel": (_I, [_H]),
    "plbm_lbm_steps_per_pass": (_I, [_H, _I]),
    "plbm_lbm_closing_triple": (_I, [_H, _I]),
    "plbm_lbm_triple_kernel": (_I, [_H, _I]),
    "plbm_set_fdm_stencil": (_I, [_H, _I]),
    "plbm_comm_unique_id": (_I, [_P]),
    "plbm_comm_init": (_I, [_H, _P, _I, _I, _I, _I]),
    "plbm_comm_transport": (_I, [_H]),
    "plbm_comm_finalize": (_I, [_H]),
    "plbm_case_tg_decay_time": (_D, [_I, _D, _D, _D]),
    "plbm_case_taylor_green": (_I, [_I, _I, _I, _D, _D, _D, _D, _D, _P, _P, _P]),
    "plbm_case_taylor_green_slab": (_I, [_I, _I, _I, _I, _D, _D, _D, _D, _D, _P, _P, _P]),
    "plbm_case_vortex": (_I, [_I, _I, _I, _D, _D, _D, _D, _D, _D, _D, _P, _P, _P]),
    # sim/lbm.h plugin seam
    "c_plbm_init": (_P, [_I, _I, _D, _P, _P, _P, _P]),
    "c_plbm_step": (None, [_P, _D]),
    "c_plbm_step_n": (None, [_P, _D, _I]),
    "c_plbm_vars": (None, [_P, _P, _P]),
    "c_plbm_free": (None, [_P]),
    "c_plbm_norm": (_D, [_I, _I, _P, _P]),
    "c_lw_init": (_P, [_I, _I, _D, _P, _P, _P, _P]),
    "c_lw_step": (None, [_P, _D]),
    "c_lw_step_n": (None, [_P, _D, _I]),
    "c_lw_vars": (None, [_P, _P, _P]),
    "c_lw_free": (None, [_P]),
    "c_lw_norm": (_D, [_I, _I, _P, _P]),
    **{f"c_{n}_{e}": sig for n in ("lw4", "lw6", "fvm")
       for e, sig in (("init", (_P, [_I, _I, _D, _P, _P, _P, _P])), ("step", (None, [_P, _D])), ("step_n", (None, [_P, _D, _I])),
                      ("vars", (None, [_P, _P, _P])), ("free", (None, [_P])), ("norm", (_D, [_I, _I, _P, _P])))},
    "c_slbm_init": (_P, [_I, _I, _D, _P, _P, _P, _P]),
    "c_slbm_step": (None, [_P, _D]),
    "c_slbm_vars": (None, [_P, _P, _P]),
    "c_slbm_free": (None, [_P]),
    "c_slbm_norm": (_D, [_I, _I, _P, _P]),
}


class PlbmError(RuntimeError):
    pass


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(periodic_lbm_b200 has no CPU fallback)"
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.plbm_last_error()
        raise PlbmError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
