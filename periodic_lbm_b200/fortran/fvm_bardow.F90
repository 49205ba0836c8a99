!> Drop-in replacement of the reference module `fvm_bardow` (src/fvm_bardow.F90:13-33) whose
!! kernels run in libplbm_b200.so.  Public names, the components of `lattice_grid` the drivers
!! touch (app/main_taylor_green.f90:36-42,70,106,140-145,155,192-198) and the calling sequences are
!! unchanged; the PDF array `grid%f` is gone (it lives on the GPU behind `grid%dev`) -- no driver
!! reads it.  Host file output (output_gnuplot/vtk/npy) stays with the reference's own
!! src/output/* modules: call update_macros first, then pass grid%rho/ux/uy to them.
module fvm_bardow
   use, intrinsic :: iso_c_binding
   use precision, only: wp, plbm_precision
   use plbm_c
   implicit none
   private

   public :: wp
   public :: lattice_grid
   public :: alloc_grid, dealloc_grid
   public :: set_properties
   public :: perform_step
   public :: update_macros
   public :: set_pdf_to_equilibrium
   public :: equilibrium
   public :: stream_fdm_bardow
   public :: stream_fvm_bardow
   public :: stream_fdm_sofonea
   public :: cx, cy, csqr
   public :: sync_indices

   type :: lattice_grid
      integer :: nx, ny
      !> macroscopic fields (host), contiguous, with the same pointer views as the reference
      real(wp), allocatable :: mf(:,:,:)
      real(wp), pointer :: rho(:,:) => null()
      real(wp), pointer ::  ux(:,:) => null()
      real(wp), pointer ::  uy(:,:) => null()
      real(wp) :: nu, dt, tau
      real(wp) :: omega, trt_magic
      real(wp) :: csqr
      integer :: iold, inew, imid
      procedure(collision_interface), pointer, pass(grid) :: collision => null()
      procedure(streaming_interface), pointer, pass(grid) :: streaming => null()
      character(len=:), allocatable :: filename, foldername
      character(len=:), allocatable :: logfile
      procedure(gridlog_interface), pointer, pass(grid) :: logger => null()
      integer :: logunit
      !> opaque handle of the device-resident lattices
      type(c_ptr) :: dev = c_null_ptr
      !> .true. (default) = periodic_dugks as built with -DDUGKS
      logical :: dugks = .true.
   end type

   abstract interface
      subroutine collision_interface(grid)
         import lattice_grid
         class(lattice_grid), intent(inout) :: grid
      end subroutine
      subroutine streaming_interface(grid)
         import lattice_grid
         class(lattice_grid), intent(inout) :: grid
      end subroutine
      subroutine gridlog_interface(grid, step)
         import lattice_grid
         class(lattice_grid), intent(in) :: grid
         integer, intent(in) :: step
      end subroutine
   end interface

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
      call sync_indices(grid)
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

   !> refresh grid%iold/inew/imid from the device handle (they flip after every step)
   subroutine sync_indices(grid)
      class(lattice_grid), intent(inout) :: grid
      integer(c_int) :: io, in, im
      call plbm_check(plbm_get_indices(grid%dev, io, in, im), "get_indices")
      grid%iold = io; grid%inew = in; grid%imid = im
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

   !> grid%omega is a public component a driver may overwrite: push it before every launch
   subroutine push_omega(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
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

   !> perform_step: streaming(); collision(); swap -- the generic sequence.  The fused fast paths
   !! live in periodic_lbm%perform_lbm_step, which can see the concrete collision procedures.
   subroutine perform_step(grid)
      type(lattice_grid), intent(inout) :: grid
      call push_omega(grid)
      call grid%streaming()
      call grid%collision()
      call plbm_check(plbm_swap(grid%dev), "swap")
      call sync_indices(grid)
   end subroutine

   !> update_macros: rho, ux, uy of lattice `inew` (the reference's one-step lag, SURVEY F3)
   subroutine update_macros(grid)
      type(lattice_grid), intent(inout), target :: grid
      call plbm_check(plbm_update_macros(grid%dev, c_loc(grid%mf(1,1,1)), c_loc(grid%mf(1,1,2)), &
                                         c_loc(grid%mf(1,1,3)), 1_c_int), "update_macros")
   end subroutine

   subroutine stream_fvm_bardow(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_stream_fvm_bardow(grid%dev), "stream_fvm_bardow")
   end subroutine

   !> The reference selects the derivative stencil at compile time (src/fvm_bardow.F90:591-660);
   !! compile this shim with the same macro and the library uses the same stencil.
   subroutine stream_fdm_bardow(grid)
      class(lattice_grid), intent(inout) :: grid
#if defined(FDM_WLS)
      integer(c_int), parameter :: stencil = 1
#elif defined(FDM_WLS_GAUSS_V1)
      integer(c_int), parameter :: stencil = 2
#elif defined(FDM_WLS_GAUSS_V2)
      integer(c_int), parameter :: stencil = 3
#elif defined(FDM_ISO)
      integer(c_int), parameter :: stencil = 4
#else
      integer(c_int), parameter :: stencil = 0
#endif
      call plbm_check(plbm_set_fdm_stencil(grid%dev, stencil), "set_fdm_stencil")
      call plbm_check(plbm_stream_fdm_bardow(grid%dev), "stream_fdm_bardow")
   end subroutine

   subroutine stream_fdm_sofonea(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_stream_fdm_sofonea(grid%dev), "stream_fdm_sofonea")
   end subroutine

end module fvm_bardow
