!> Working precision, same switch as the reference (src/precision.F90:11-15):
!! -DPRECISION_SINGLE selects real(4) and the fp32 kernels of libplbm_b200.
module precision
   use plbm_c, only: PLBM_F64, PLBM_F32
   implicit none
   private
   public :: wp, plbm_precision
#if PRECISION_SINGLE
   integer, parameter :: wp = kind(1.0)
   integer, parameter :: plbm_precision = PLBM_F32
#else
   integer, parameter :: wp = kind(1.0d0)
   integer, parameter :: plbm_precision = PLBM_F64
#endif
end module precision
