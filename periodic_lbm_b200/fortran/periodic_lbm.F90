!> Drop-in replacement of the reference module `periodic_lbm` (src/periodic_lbm.f90:9-11).
!! perform_lbm_step looks at the two procedure pointers: for the pairs this library fuses
!! (lbm_stream | stream_fvm_bardow | stream_fdm_*) x (collide_bgk | collide_trt | collide_rr |
!! collide_bgk_improved) ONE kernel does the whole step -- and consecutive calls are batched by the
!! library into two-steps-per-pass launches (see alloc_grid); any other (user-supplied) pair is called
!! one after the other like the reference does.  -DSPLIT and the -DFDM_* stencils are honoured on
!! the fused path exactly as by the stand-alone wrappers (one mapping, plbm_lattice.F90).
module periodic_lbm
   use, intrinsic :: iso_c_binding
   use plbm_lattice, only: lattice_grid, lbm_stream
   use fvm_bardow, only: perform_steps
   implicit none
   private
   public :: perform_lbm_step
   public :: perform_lbm_steps
   public :: lbm_stream
contains

   !> src/periodic_lbm.f90:15-29: streaming(); collision(); swap
   subroutine perform_lbm_step(grid)
      type(lattice_grid), intent(inout) :: grid
      call perform_steps(grid, 1)
   end subroutine

   !> extension: n steps per call (no host round trip between steps)
   subroutine perform_lbm_steps(grid, n)
      type(lattice_grid), intent(inout) :: grid
      integer, intent(in) :: n
      call perform_steps(grid, n)
   end subroutine

end module periodic_lbm
