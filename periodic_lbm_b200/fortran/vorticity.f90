!> Drop-in replacement of the reference module `vorticity` (src/vorticity.f90:8-9): same
!! signatures on host arrays; the stencil runs on the GPU (4th-order weights as shipped, F9).
module vorticity
   use, intrinsic :: iso_c_binding
   use precision, only: wp, plbm_precision
   use plbm_c
   implicit none
   private
   public :: vorticity_2nd
   public :: vorticity_4th
contains
   subroutine vorticity_nth(order, ux, uy, omega)
      integer, intent(in) :: order
      real(wp), intent(in), target, contiguous :: ux(:,:), uy(:,:)
      real(wp), intent(out), target, contiguous :: omega(:,:)
      type(c_ptr) :: tmp
      call plbm_check(plbm_alloc_grid(tmp, int(size(ux,2),c_int), int(size(ux,1),c_int), 2_c_int, plbm_precision), "alloc_grid")
      call plbm_check(plbm_vorticity_host(tmp, int(order,c_int), c_loc(ux), c_loc(uy), c_loc(omega)), "vorticity")
      call plbm_check(plbm_dealloc_grid(tmp), "dealloc_grid")
   end subroutine
   subroutine vorticity_2nd(ux, uy, omega)
      real(wp), intent(in), target, contiguous :: ux(:,:), uy(:,:)
      real(wp), intent(out), target, contiguous :: omega(:,:)
      call vorticity_nth(2, ux, uy, omega)
   end subroutine
   subroutine vorticity_4th(ux, uy, omega)
      real(wp), intent(in), target, contiguous :: ux(:,:), uy(:,:)
      real(wp), intent(out), target, contiguous :: omega(:,:)
      call vorticity_nth(4, ux, uy, omega)
   end subroutine
end module vorticity
