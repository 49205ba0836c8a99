!> oracle/ref_shim.f90 -- TEST INFRASTRUCTURE ONLY.
!! C-callable wrapper that drives the UNMODIFIED reference kernels (compiled where they lie under
!! /root/reference/src by `make -C oracle ref`) so the C restatement in plbm_oracle_impl.h can be checked
!! against the real thing, bit for bit, on a machine that has a Fortran compiler.  No such compiler exists
!! in the image this repository was developed in (SURVEY.md F1): this file has never been compiled there and
!! tests/test_oracle_vs_ref.py skips when oracle/_ref/libplbm_ref.so is absent.
module ref_shim
   use, intrinsic :: iso_c_binding
   use precision, only: wp
   use fvm_bardow
   use periodic_lbm, only: perform_lbm_step, lbm_stream
   use collision_bgk, only: collide_bgk
   use collision_trt, only: collide_trt
   use collision_regularized, only: collide_rr
   use periodic_dugks, only: perform_dugks_step, dugks_collide, dugks_stream
   implicit none
contains

   !> Run `nsteps` steps of (scheme, collision) from the equilibrium of (rho,ux,uy) and return the PDFs of
   !! lattice `iold` (f(ld,nx,0:8)) and the lagged macros, exactly as the reference computes them.
   !!   scheme: 0 lbm_stream (perform_lbm_step), 1 stream_fvm_bardow (perform_step), 2 perform_dugks_step,
   !!           4 stream_fdm_bardow, 5 stream_fdm_sofonea           collision: 0 bgk, 1 trt, 2 rr
   subroutine ref_run(nx, ny, ld, scheme, collision, nu, dt, magic, nsteps, rho, ux, uy, f_out, iold_out) bind(c)
      integer(c_int), value :: nx, ny, ld, scheme, collision, nsteps
      real(c_double), value :: nu, dt, magic
      real(c_double), intent(inout) :: rho(ny,nx), ux(ny,nx), uy(ny,nx)
      real(c_double), intent(out) :: f_out(ld,nx,0:8)
      integer(c_int), intent(out) :: iold_out
      type(lattice_grid) :: grid
      integer :: step

      call alloc_grid(grid, nx, ny, nf=2, log=.false.)
      select case (collision)
      case (1); grid%collision => collide_trt
      case (2); grid%collision => collide_rr
      case default; grid%collision => collide_bgk
      end select
      select case (scheme)
      case (1); grid%streaming => stream_fvm_bardow
      case (2); grid%collision => dugks_collide; grid%streaming => dugks_stream
      case (4); grid%streaming => stream_fdm_bardow
      case (5); grid%streaming => stream_fdm_sofonea
      case default; grid%streaming => lbm_stream
      end select
      call set_properties(grid, real(nu,wp), real(dt,wp), magic=real(magic,wp))
      grid%rho = real(rho,wp); grid%ux = real(ux,wp); grid%uy = real(uy,wp)
      call set_pdf_to_equilibrium(grid)
      do step = 1, nsteps
         if (scheme == 2) then
            call perform_dugks_step(grid)
         else if (scheme == 0) then
            call perform_lbm_step(grid)
         else
            call perform_step(grid)
         end if
      end do
      call update_macros(grid)          ! reads lattice inew: the reference's one-step lag
      rho = grid%rho; ux = grid%ux; uy = grid%uy
      f_out = 0
      f_out(1:size(grid%f,1),:,:) = grid%f(:,:,:,grid%iold)
      iold_out = grid%iold
      call dealloc_grid(grid)
   end subroutine
end module ref_shim
