!> Drop-in replacement of the reference module `periodic_lbm` (src/periodic_lbm.f90:9-11).
!! perform_lbm_step inspects the two procedure pointers: for the pairs this library fuses
!! (lbm_stream | stream_fvm_bardow) x (collide_bgk | collide_trt | collide_rr) ONE kernel does the
!! whole step; any other (user-supplied) pair is called one after the other like the reference does.
module periodic_lbm
   use, intrinsic :: iso_c_binding
   use fvm_bardow, only: lattice_grid, stream_fvm_bardow, stream_fdm_bardow, stream_fdm_sofonea, sync_indices
   use collision_bgk, only: collide_bgk
   use collision_trt, only: collide_trt
   use collision_regularized, only: collide_rr
   use plbm_c
   implicit none
   private
   public :: perform_lbm_step
   public :: perform_lbm_steps
   public :: lbm_stream
contains

   subroutine perform_lbm_step(grid)
      type(lattice_grid), intent(inout) :: grid
      call perform_lbm_steps(grid, 1)
   end subroutine

   !> extension: n steps per call (no host round trip between steps)
   subroutine perform_lbm_steps(grid, n)
      type(lattice_grid), intent(inout) :: grid
      integer, intent(in) :: n
      integer(c_int) :: cid, sid
      integer :: i

      cid = -1
      if (associated(grid%collision, collide_bgk)) then
#if SPLIT
         cid = PLBM_BGK_SPLIT
#else
         cid = PLBM_BGK
#endif
      else if (associated(grid%collision, collide_trt)) then
         cid = PLBM_TRT
      else if (associated(grid%collision, collide_rr)) then
         cid = PLBM_RR
      end if
      sid = -1
      if (associated(grid%streaming, lbm_stream)) sid = PLBM_STREAM_LBM
      if (associated(grid%streaming, stream_fvm_bardow)) sid = PLBM_STREAM_FVM_BARDOW
      if (associated(grid%streaming, stream_fdm_bardow)) sid = PLBM_STREAM_FDM_BARDOW
      if (associated(grid%streaming, stream_fdm_sofonea)) sid = PLBM_STREAM_FDM_SOFONEA

      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
      if (cid >= 0 .and. sid >= 0) then
         call plbm_check(plbm_perform_step(grid%dev, sid, cid, int(n,c_int)), "perform_lbm_step")
      else
         do i = 1, n
            call grid%streaming()
            call grid%collision()
            call plbm_check(plbm_swap(grid%dev), "swap")
         end do
      end if
      call sync_indices(grid)
   end subroutine

   subroutine lbm_stream(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_lbm_stream(grid%dev), "lbm_stream")
   end subroutine

end module periodic_lbm
