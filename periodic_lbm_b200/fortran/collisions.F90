!> Drop-in replacements of the reference collision modules.  Each `collide_*` keeps the
!! `subroutine name(grid)` interface so `grid%collision => collide_rr` works unchanged
!! (app/main_taylor_green.f90:39).  The procedures themselves are defined in plbm_lattice.F90
!! (so that the orchestrators of fvm_bardow can recognise them) and re-exported here under the
!! reference's module names.
module collision_bgk
   use plbm_lattice, only: collide_bgk
   implicit none
   private
   public :: collide_bgk        ! src/collision_bgk.F90:11
end module collision_bgk

module collision_trt
   use precision, only: wp
   use plbm_lattice, only: collide_trt
   implicit none
   private
   public :: magic_number, lambda_d   ! src/collision_trt.F90:9-10
   public :: collide_trt
contains
   !> src/collision_trt.F90:24-28
   pure real(wp) function magic_number(le, ld)
      real(wp), intent(in) :: le, ld
      magic_number = (2.0_wp - le)*(2.0_wp - ld)/(4.0_wp*le*ld)
   end function
   !> src/collision_trt.F90:30-34
   pure real(wp) function lambda_d(omega, x)
      real(wp), intent(in) :: omega, x
      lambda_d = (4.0_wp - 2.0_wp*omega)/(4.0_wp*x*omega + 2.0_wp - omega)
   end function
end module collision_trt

module collision_regularized
   use plbm_lattice, only: collide_rr
   implicit none
   private
   public :: collide_rr         ! src/collision_regularized.F90:9
end module collision_regularized

module collision_bgk_improved
   use plbm_lattice, only: collide_bgk_improved
   implicit none
   private
   public :: collide_bgk_improved   ! src/collision_bgk_improved.f90:8
end module collision_bgk_improved
