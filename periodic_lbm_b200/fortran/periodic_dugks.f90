!> Drop-in replacement of the reference module `periodic_dugks` (src/periodic_dugks.F90:9-12).
!! grid%dugks selects the -DDUGKS branch (default) or the macro-less default build of the reference.
module periodic_dugks
   use, intrinsic :: iso_c_binding
   use plbm_lattice, only: lattice_grid, dugks_stream, dugks_collide, plbm_sync_indices, plbm_push_omega
   use plbm_c
   implicit none
   private
   public :: perform_dugks_step
   public :: dugks_stream
   public :: dugks_collide
contains

   !> src/periodic_dugks.F90:25-38: collision(); streaming(); swap -- one fused launch when both
   !! pointers are unset or name this module's own passes
   subroutine perform_dugks_step(grid)
      type(lattice_grid), intent(inout) :: grid
      logical :: fused
      call plbm_push_omega(grid)
      fused = .true.
      if (associated(grid%collision)) fused = fused .and. associated(grid%collision, dugks_collide)
      if (associated(grid%streaming)) fused = fused .and. associated(grid%streaming, dugks_stream)
      if (fused) then
         call plbm_check(plbm_perform_dugks_step(grid%dev, merge(1_c_int, 0_c_int, grid%dugks), 1_c_int), "perform_dugks_step")
      else
         call grid%collision()
         call grid%streaming()
         call plbm_check(plbm_swap(grid%dev), "swap")
      end if
      call plbm_sync_indices(grid)
   end subroutine

end module periodic_dugks
