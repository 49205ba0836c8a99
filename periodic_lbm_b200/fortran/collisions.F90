!> Drop-in replacements of the reference collision modules.  Each `collide_*` keeps the
!! `subroutine name(grid)` interface so `grid%collision => collide_rr` works unchanged
!! (app/main_taylor_green.f90:39).
module collision_bgk
   use, intrinsic :: iso_c_binding
   use fvm_bardow, only: lattice_grid
   use plbm_c
   implicit none
   private
   public :: collide_bgk
contains
   subroutine collide_bgk(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
#if SPLIT
      call plbm_check(plbm_collide(grid%dev, PLBM_BGK_SPLIT), "collide_bgk")
#else
      call plbm_check(plbm_collide(grid%dev, PLBM_BGK), "collide_bgk")
#endif
   end subroutine
end module collision_bgk

module collision_trt
   use, intrinsic :: iso_c_binding
   use precision, only: wp
   use fvm_bardow, only: lattice_grid
   use plbm_c
   implicit none
   private
   public :: magic_number, lambda_d
   public :: collide_trt
contains
   pure real(wp) function magic_number(le, ld)
      real(wp), intent(in) :: le, ld
      magic_number = (2.0_wp - le)*(2.0_wp - ld)/(4.0_wp*le*ld)
   end function
   pure real(wp) function lambda_d(omega, x)
      real(wp), intent(in) :: omega, x
      lambda_d = (4.0_wp - 2.0_wp*omega)/(4.0_wp*x*omega + 2.0_wp - omega)
   end function
   subroutine collide_trt(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
#if SPLIT
      call plbm_check(plbm_collide(grid%dev, 4_c_int), "collide_trt")   ! PLBM_TRT_SPLIT
#else
      call plbm_check(plbm_collide(grid%dev, PLBM_TRT), "collide_trt")
#endif
   end subroutine
end module collision_trt

module collision_regularized
   use, intrinsic :: iso_c_binding
   use fvm_bardow, only: lattice_grid
   use plbm_c
   implicit none
   private
   public :: collide_rr
contains
   subroutine collide_rr(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
      call plbm_check(plbm_collide(grid%dev, PLBM_RR), "collide_rr")
   end subroutine
end module collision_regularized

module collision_bgk_improved
   use, intrinsic :: iso_c_binding
   use fvm_bardow, only: lattice_grid
   use plbm_c
   implicit none
   private
   public :: collide_bgk_improved
contains
   subroutine collide_bgk_improved(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
      call plbm_check(plbm_collide(grid%dev, 5_c_int), "collide_bgk_improved")   ! PLBM_BGK_IMPROVED
   end subroutine
end module collision_bgk_improved
