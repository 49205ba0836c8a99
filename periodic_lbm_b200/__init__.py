"""periodic_lbm_b200 -- B200 (sm_100a) implementation of the periodic D2Q9 hot path of
ivan-pi/periodic-lbm behind the reference's own interfaces.

The compute lives in libplbm_b200.so (hand-written CUDA, C ABI in include/plbm.h).  This package is
the host-side mirror of the reference's Fortran module interfaces used by tests and bench.py.
Importing it requires the built library; there is no CPU fallback.
"""
from . import capi  # noqa: F401  (raises ImportError when libplbm_b200.so is missing)
from .capi import BGK, BGK_IMPROVED, BGK_SPLIT, F32, F64, RR, TRT, TRT_SPLIT, PlbmError  # noqa: F401
from .cases import TaylorGreen, VortexCase, steps_until, taylor_green_params, vortex_params  # noqa: F401
from .lattice import *  # noqa: F401,F403
from .output import load_checkpoint, output_gnuplot, output_npy, output_vtk, save_checkpoint, set_output_folder  # noqa: F401
from .plugin import SimPlugin  # noqa: F401
from .slab import Slab, slab_of  # noqa: F401


def device_count() -> int:
    return int(capi.lib.plbm_device_count())


def launch_count() -> int:
    return int(capi.lib.plbm_launch_count())
