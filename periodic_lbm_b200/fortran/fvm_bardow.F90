!> Drop-in replacement of the reference module `fvm_bardow` (src/fvm_bardow.F90:13-33) whose
!! kernels run in libplbm_b200.so.  Public names, the components of `lattice_grid` the drivers
!! touch (app/main_taylor_green.f90:36-42,70,106,140-145,155,192-198; app/main_vortex.f90:38-46,
!! 104-125,150-161), the type-bound `set_output_folder` and the calling sequences are unchanged; the
!! PDF array `grid%f` is gone (it lives on the GPU behind `grid%dev`) -- no driver reads it.
!!
!! Host file output: `output_gnuplot`, `output_vtk`, `output_npy` are thin wrappers over the
!! reference's OWN writer modules (`output_gnuplot`, `vtk`, `output_npy` of src/output/, compiled
!! unchanged next to this shim -- they only need `precision::wp`), fed from the host mirrors
!! grid%rho/ux/uy/mf that `update_macros` fills, exactly like src/fvm_bardow.F90:895-997.
!!
!! The type and the kernel wrappers are defined in plbm_lattice.F90 (see there for why) and
!! re-exported here under the reference's names.
!!
!! SOURCE ONLY: no Fortran compiler exists in the image this library was developed in (SURVEY.md F1).
module fvm_bardow
   use, intrinsic :: iso_c_binding
   use precision, only: wp, plbm_precision
   use plbm_c
   use plbm_lattice, only: lattice_grid, plbm_sync_indices, plbm_push_omega, plbm_collision_id, plbm_streaming_id, &
                           stream_fdm_bardow, stream_fvm_bardow, stream_fdm_sofonea
   ! the reference's writers (src/output/gnuplot.F90, vtk.F90, npy.f90), used as they are
   use vtk, only: output_vtk_structuredPoints
   use output_gnuplot, only: output_gnuplot_grid
   use output_npy, only: output_fluid_npy
   implicit none
   private

   ! --- the reference's public list, src/fvm_bardow.F90:13-33 ---
   public :: wp
   public :: lattice_grid
   public :: alloc_grid, dealloc_grid
   public :: set_properties
   public :: perform_step, perform_triple_step
   public :: update_macros
   public :: output_gnuplot, output_vtk, output_npy
   public :: set_pdf_to_equilibrium
   public :: equilibrium
   public :: stream_fdm_bardow
   public :: stream_fvm_bardow
   public :: stream_fdm_sofonea
   public :: cx, cy, csqr
   ! --- extensions ---
   public :: sync_indices
   public :: perform_steps

   real(wp), parameter :: cx(0:8) = [real(wp) :: 0, 1, 0, -1, 0, 1, -1, -1, 1]
   real(wp), parameter :: cy(0:8) = [real(wp) :: 0, 0, 1, 0, -1, 1, 1, -1, -1]
   real(wp), parameter :: csqr = 1._wp/3._wp

contains

   subroutine alloc_grid(grid, nx, ny, nf, log)
      type(lattice_grid), intent(out), target :: grid
      integer, intent(in) :: nx, ny
      integer, intent(in), optional :: nf
      logical, intent(in), optional :: log
      integer :: nf_
      logical :: log_
      nf_ = 2
      if (present(nf)) nf_ = nf
      grid%nx = nx
      grid%ny = ny
      allocate(grid%mf(ny,nx,3))
      grid%rho => grid%mf(:,:,1)
      grid%ux  => grid%mf(:,:,2)
      grid%uy  => grid%mf(:,:,3)
      call plbm_check(plbm_alloc_grid(grid%dev, int(nx,c_int), int(ny,c_int), int(nf_,c_int), plbm_precision), "alloc_grid")
      ! the drivers call perform_lbm_step once per time step (app/main_taylor_green.f90:98-119): let the library
      ! batch up to 64 of those calls into one launch sequence (two steps per pass over HBM); anything that
      ! looks at the grid (update_macros, ...) runs the pending steps first, so the drivers see no difference
      call plbm_check(plbm_set_step_deferral(grid%dev, 64_c_int), "set_step_deferral")
      call plbm_sync_indices(grid)
      log_ = .true.
      if (present(log)) log_ = log
      if (log_) then
         if (.not. allocated(grid%logfile)) grid%logfile = "lattice_grid_log.txt"
         open(newunit=grid%logunit, file=grid%logfile, status='unknown')
      end if
   end subroutine

   subroutine dealloc_grid(grid)
      type(lattice_grid), intent(inout) :: grid
      logical :: isopen
      inquire(grid%logunit, opened=isopen)
      if (isopen) close(grid%logunit)
      nullify(grid%uy, grid%ux, grid%rho)
      if (allocated(grid%mf)) deallocate(grid%mf)
      if (c_associated(grid%dev)) call plbm_check(plbm_dealloc_grid(grid%dev), "dealloc_grid")
      grid%dev = c_null_ptr
   end subroutine

   subroutine sync_indices(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_sync_indices(grid)
   end subroutine

   subroutine set_properties(grid, nu, dt, magic)
      type(lattice_grid), intent(inout) :: grid
      real(wp), intent(in) :: nu, dt
      real(wp), optional :: magic
      real(c_double) :: props(6)
      if (present(magic)) then
         call plbm_check(plbm_set_properties(grid%dev, real(nu,c_double), real(dt,c_double), real(magic,c_double), 1_c_int), &
                         "set_properties")
      else
         call plbm_check(plbm_set_properties(grid%dev, real(nu,c_double), real(dt,c_double), 0._c_double, 0_c_int), &
                         "set_properties")
      end if
      ! the library derives tau, omega, trt_magic in working precision; mirror them on the host
      call plbm_check(plbm_get_properties(grid%dev, props), "get_properties")
      grid%nu = real(props(1),wp);    grid%dt = real(props(2),wp)
      grid%tau = real(props(3),wp);   grid%omega = real(props(4),wp)
      grid%trt_magic = real(props(5),wp); grid%csqr = real(props(6),wp)
      print *, "trt magic = ", grid%trt_magic
   end subroutine

   pure function equilibrium(rho, ux, uy) result(feq)
      real(wp), intent(in) :: rho, ux, uy
      real(wp) :: feq(0:8)
      real(wp), parameter :: w(0:8) = [4._wp/9._wp, 1._wp/9._wp, 1._wp/9._wp, 1._wp/9._wp, 1._wp/9._wp, &
                                       1._wp/36._wp, 1._wp/36._wp, 1._wp/36._wp, 1._wp/36._wp]
      real(wp) :: indp, cu
      integer :: q
      ! host-side convenience only (the device kernels carry the bit-exact evaluation order)
      indp = 1.0_wp - 1.5_wp*(ux*ux + uy*uy)
      do q = 0, 8
         cu = cx(q)*ux + cy(q)*uy
         feq(q) = w(q)*rho*(indp + 3.0_wp*cu + 4.5_wp*cu*cu)
      end do
   end function

   subroutine set_pdf_to_equilibrium(grid)
      type(lattice_grid), intent(inout), target :: grid
      call plbm_check(plbm_set_pdf_to_equilibrium(grid%dev, c_loc(grid%mf(1,1,1)), c_loc(grid%mf(1,1,2)), &
                                                  c_loc(grid%mf(1,1,3))), "set_pdf_to_equilibrium")
   end subroutine

   !> perform_step (src/fvm_bardow.F90:307-320): streaming(); collision(); swap.  When both procedure
   !! pointers name kernels of this library the whole step is ONE fused launch (app/main_vortex.f90:43-44,114
   !! = stream_fvm_bardow + collide_bgk); a user-supplied procedure on either side is called the way the
   !! reference calls it.
   subroutine perform_step(grid)
      type(lattice_grid), intent(inout) :: grid
      call perform_steps(grid, 1)
   end subroutine

   !> extension: n steps per call (no host round trip between steps)
   subroutine perform_steps(grid, n)
      type(lattice_grid), intent(inout) :: grid
      integer, intent(in) :: n
      integer(c_int) :: cid, sid
      integer :: i
      cid = plbm_collision_id(grid)
      sid = plbm_streaming_id(grid)
      call plbm_push_omega(grid)
      if (cid >= 0 .and. sid >= 0) then
         if (sid == PLBM_STREAM_FDM_BARDOW) call push_fdm_stencil(grid)
         call plbm_check(plbm_perform_step(grid%dev, sid, cid, int(n,c_int)), "perform_step")
      else
         do i = 1, n
            call grid%streaming()
            call grid%collision()
            call plbm_check(plbm_swap(grid%dev), "swap")
         end do
      end if
      call plbm_sync_indices(grid)
   end subroutine

   !> perform_triple_step (src/fvm_bardow.F90:322-340), grids allocated with nf = 3: stream iold -> inew,
   !! keep the pre-collision PDFs in iold, collide inew, rotate (iold, inew, imid) <- (inew, imid, iold).
   !! Needs both pointers to name kernels of this library (the copy of the PDFs happens on the device).
   subroutine perform_triple_step(grid)
      type(lattice_grid), intent(inout) :: grid
      integer(c_int) :: cid, sid
      cid = plbm_collision_id(grid)
      sid = plbm_streaming_id(grid)
      if (cid < 0 .or. sid < 0) then
         write(*,'(A)') "[perform_triple_step] grid%streaming / grid%collision must be kernels of this library"
         error stop
      end if
      call plbm_push_omega(grid)
      if (sid == PLBM_STREAM_FDM_BARDOW) call push_fdm_stencil(grid)
      call plbm_check(plbm_perform_triple_step(grid%dev, sid, cid, 1_c_int), "perform_triple_step")
      call plbm_sync_indices(grid)
   end subroutine

   !> the fused path never runs the stream_fdm_bardow wrapper, so the compile-time stencil choice
   !! (-DFDM_WLS, -DFDM_WLS_GAUSS_V1/_V2, -DFDM_ISO) is pushed here
   subroutine push_fdm_stencil(grid)
      type(lattice_grid), intent(inout) :: grid
#if defined(FDM_WLS)
      call plbm_check(plbm_set_fdm_stencil(grid%dev, 1_c_int), "set_fdm_stencil")
#elif defined(FDM_WLS_GAUSS_V1)
      call plbm_check(plbm_set_fdm_stencil(grid%dev, 2_c_int), "set_fdm_stencil")
#elif defined(FDM_WLS_GAUSS_V2)
      call plbm_check(plbm_set_fdm_stencil(grid%dev, 3_c_int), "set_fdm_stencil")
#elif defined(FDM_ISO)
      call plbm_check(plbm_set_fdm_stencil(grid%dev, 4_c_int), "set_fdm_stencil")
#else
      call plbm_check(plbm_set_fdm_stencil(grid%dev, 0_c_int), "set_fdm_stencil")
#endif
   end subroutine

   !> update_macros: rho, ux, uy of lattice `inew` (the reference's one-step lag, SURVEY F3)
   subroutine update_macros(grid)
      type(lattice_grid), intent(inout), target :: grid
      call plbm_check(plbm_update_macros(grid%dev, c_loc(grid%mf(1,1,1)), c_loc(grid%mf(1,1,2)), &
                                         c_loc(grid%mf(1,1,3)), 1_c_int), "update_macros")
   end subroutine

   ! ---- host file output (src/fvm_bardow.F90:895-997): names, then the reference's writers --------

   !> "<foldername>/<filename><step as I0.9><ext>".  with_mkdir: output_gnuplot and output_npy create the
   !! folder on every call and always put a '/' in front of the file name, output_vtk does neither
   !! (src/fvm_bardow.F90:907-919, 942-954, 981-985) -- kept as the reference has it.
   function output_name(grid, step, ext, with_mkdir) result(fullname)
      type(lattice_grid), intent(in) :: grid
      integer, intent(in), optional :: step
      character(len=*), intent(in) :: ext
      logical, intent(in) :: with_mkdir
      character(len=:), allocatable :: fullname
      character(len=64) :: istr
      integer :: istat
      istr = ''
      if (present(step)) write(istr,'(I0.9)') step
      fullname = ''
      if (allocated(grid%foldername)) then
         if (with_mkdir) then
            call execute_command_line("mkdir -p "//trim(grid%foldername), exitstat=istat, wait=.true.)
            if (istat /= 0) then
               write(*,'(A)') "[output_grid] error making directory "//grid%foldername
               error stop
            end if
            fullname = trim(grid%foldername)
         else
            fullname = trim(grid%foldername)//'/'
         end if
      end if
      if (with_mkdir) fullname = fullname//'/'
      fullname = fullname//grid%filename//trim(istr)//ext
   end function

   subroutine output_gnuplot(grid, step)
      type(lattice_grid), intent(in) :: grid
      integer, intent(in), optional :: step
      call output_gnuplot_grid(output_name(grid, step, '.txt', .true.), grid%nx, grid%ny, grid%rho, grid%ux, grid%uy)
   end subroutine

   subroutine output_npy(grid, step)
      type(lattice_grid), intent(in) :: grid
      integer, intent(in), optional :: step
      call output_fluid_npy(output_name(grid, step, '.npy', .true.), grid%nx, grid%ny, grid%mf)
   end subroutine

   subroutine output_vtk(grid, step, binary)
      type(lattice_grid), intent(in) :: grid
      integer, intent(in), optional :: step
      logical, intent(in), optional :: binary
      if (present(binary)) then
         if (binary) then
            print *, "binary output not implemented"
            return
         end if
      end if
      call output_vtk_structuredPoints(output_name(grid, step, '.vtk', .false.), grid%nx, grid%ny, grid%rho, grid%ux, grid%uy)
   end subroutine

end module fvm_bardow
