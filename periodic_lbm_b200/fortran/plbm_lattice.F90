!> plbm_lattice -- the one module of the shim that owns `type lattice_grid` and every kernel wrapper
!! with the reference's `subroutine name(grid)` interface (src/fvm_bardow.F90:71-79).
!!
!! The reference spreads these procedures over several modules that all `use fvm_bardow`
!! (collision_bgk, collision_trt, collision_regularized, periodic_lbm, periodic_dugks).  Its
!! orchestrators call the kernels through two procedure pointers (grid%streaming, grid%collision),
!! whereas the device library runs stream + collide as ONE fused kernel.  To fuse, an orchestrator has to
!! recognise which procedures the two pointers name -- `associated(grid%collision, collide_bgk)` -- and
!! `fvm_bardow::perform_step` could not see `collide_bgk` if that lived in a module that uses fvm_bardow.
!! So all wrappers are defined HERE, below the reference's module names, and those modules
!! (fvm_bardow.F90, collisions.F90, periodic_lbm.F90, periodic_dugks.f90 of this directory) re-export them
!! under the reference's public names.  A driver never names this module.
!!
!! SOURCE ONLY: no Fortran compiler exists in the image this library was developed in (SURVEY.md F1).
module plbm_lattice
   use, intrinsic :: iso_c_binding
   use precision, only: wp, plbm_precision
   use plbm_c
   implicit none
   private

   public :: lattice_grid
   public :: collision_interface, streaming_interface, gridlog_interface
   public :: plbm_sync_indices, plbm_push_omega
   public :: plbm_collision_id, plbm_streaming_id
   ! kernels, re-exported by the modules that carry the reference's names
   public :: collide_bgk, collide_trt, collide_rr, collide_bgk_improved
   public :: lbm_stream, stream_fvm_bardow, stream_fdm_bardow, stream_fdm_sofonea
   public :: dugks_collide, dugks_stream

   !> src/fvm_bardow.F90:37-69 without the PDF array `f` (device-resident behind `dev`; no driver reads it)
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
   contains
      procedure :: set_output_folder
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

   !> the reference selects the derivative stencil of stream_fdm_bardow at compile time
   !! (src/fvm_bardow.F90:591-660); compile this shim with the same macro
#if defined(FDM_WLS)
   integer(c_int), parameter :: fdm_stencil = 1
#elif defined(FDM_WLS_GAUSS_V1)
   integer(c_int), parameter :: fdm_stencil = 2
#elif defined(FDM_WLS_GAUSS_V2)
   integer(c_int), parameter :: fdm_stencil = 3
#elif defined(FDM_ISO)
   integer(c_int), parameter :: fdm_stencil = 4
#else
   integer(c_int), parameter :: fdm_stencil = 0
#endif

contains

   !> refresh grid%iold/inew/imid from the device handle (they flip after every step)
   subroutine plbm_sync_indices(grid)
      class(lattice_grid), intent(inout) :: grid
      integer(c_int) :: io, in, im
      call plbm_check(plbm_get_indices(grid%dev, io, in, im), "get_indices")
      grid%iold = io; grid%inew = in; grid%imid = im
   end subroutine

   !> grid%omega is a public component a driver may overwrite: push it before every launch
   subroutine plbm_push_omega(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_omega(grid%dev, real(grid%omega,c_double)), "set_omega")
   end subroutine

   !> library id of the procedure grid%collision names, -1 for a user-supplied one.
   !! -DSPLIT selects the cache-blocked kernels exactly like the reference
   !! (src/collision_bgk.F90:23-31, src/collision_trt.F90:52-60).
   integer(c_int) function plbm_collision_id(grid) result(cid)
      class(lattice_grid), intent(in) :: grid
      cid = -1
      if (.not. associated(grid%collision)) return
      if (associated(grid%collision, collide_bgk)) then
#if SPLIT
         cid = PLBM_BGK_SPLIT
#else
         cid = PLBM_BGK
#endif
      else if (associated(grid%collision, collide_trt)) then
#if SPLIT
         cid = PLBM_TRT_SPLIT
#else
         cid = PLBM_TRT
#endif
      else if (associated(grid%collision, collide_rr)) then
         cid = PLBM_RR
      else if (associated(grid%collision, collide_bgk_improved)) then
         cid = PLBM_BGK_IMPROVED
      end if
   end function

   !> library id of the procedure grid%streaming names, -1 for a user-supplied one
   integer(c_int) function plbm_streaming_id(grid) result(sid)
      class(lattice_grid), intent(in) :: grid
      sid = -1
      if (.not. associated(grid%streaming)) return
      if (associated(grid%streaming, lbm_stream)) then
         sid = PLBM_STREAM_LBM
      else if (associated(grid%streaming, stream_fvm_bardow)) then
         sid = PLBM_STREAM_FVM_BARDOW
      else if (associated(grid%streaming, stream_fdm_bardow)) then
         sid = PLBM_STREAM_FDM_BARDOW
      else if (associated(grid%streaming, stream_fdm_sofonea)) then
         sid = PLBM_STREAM_FDM_SOFONEA
      end if
   end function

   ! ---- collisions: in place on lattice `inew` ------------------------------------------------

   !> src/collision_bgk.F90:17-33
   subroutine collide_bgk(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_push_omega(grid)
#if SPLIT
      call plbm_check(plbm_collide(grid%dev, PLBM_BGK_SPLIT), "collide_bgk")
#else
      call plbm_check(plbm_collide(grid%dev, PLBM_BGK), "collide_bgk")
#endif
   end subroutine

   !> src/collision_trt.F90:41-62
   subroutine collide_trt(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_push_omega(grid)
#if SPLIT
      call plbm_check(plbm_collide(grid%dev, PLBM_TRT_SPLIT), "collide_trt")
#else
      call plbm_check(plbm_collide(grid%dev, PLBM_TRT), "collide_trt")
#endif
   end subroutine

   !> src/collision_regularized.F90:21-38
   subroutine collide_rr(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_push_omega(grid)
      call plbm_check(plbm_collide(grid%dev, PLBM_RR), "collide_rr")
   end subroutine

   !> src/collision_bgk_improved.f90:16-22
   subroutine collide_bgk_improved(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_push_omega(grid)
      call plbm_check(plbm_collide(grid%dev, PLBM_BGK_IMPROVED), "collide_bgk_improved")
   end subroutine

   ! ---- streaming: lattice `iold` -> lattice `inew` -------------------------------------------

   !> src/periodic_lbm.f90:32-43
   subroutine lbm_stream(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_lbm_stream(grid%dev), "lbm_stream")
   end subroutine

   !> src/fvm_bardow.F90:393-509
   subroutine stream_fvm_bardow(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_stream_fvm_bardow(grid%dev), "stream_fvm_bardow")
   end subroutine

   !> src/fvm_bardow.F90:511-685
   subroutine stream_fdm_bardow(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_set_fdm_stencil(grid%dev, fdm_stencil), "set_fdm_stencil")
      call plbm_check(plbm_stream_fdm_bardow(grid%dev), "stream_fdm_bardow")
   end subroutine

   !> src/fvm_bardow.F90:688-893
   subroutine stream_fdm_sofonea(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_stream_fdm_sofonea(grid%dev), "stream_fdm_sofonea")
   end subroutine

   ! ---- DUGKS passes (src/periodic_dugks.F90:46-77, 172-188) ---------------------------------

   subroutine dugks_collide(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_push_omega(grid)
      call plbm_check(plbm_dugks_collide(grid%dev, merge(1_c_int, 0_c_int, grid%dugks)), "dugks_collide")
   end subroutine

   subroutine dugks_stream(grid)
      class(lattice_grid), intent(inout) :: grid
      call plbm_check(plbm_dugks_stream(grid%dev, merge(1_c_int, 0_c_int, grid%dugks)), "dugks_stream")
   end subroutine

   ! ---- type-bound: grid%set_output_folder (src/fvm_bardow.F90:67-68, 999-1025) -----------------

   !> make the folder (mkdir -p, or -pv when verbose) and remember it for the output_* wrappers
   subroutine set_output_folder(grid, foldername, verbose)
      class(lattice_grid), intent(inout) :: grid
      character(len=*), intent(in) :: foldername
      logical, intent(in), optional :: verbose
      integer :: istat
      logical :: verbose_
      verbose_ = .false.
      if (present(verbose)) verbose_ = verbose
      call execute_command_line('mkdir '//merge('-pv', '-p ', verbose_)//' '//trim(foldername), exitstat=istat, wait=.true.)
      if (istat /= 0) then
         write(*,'(A)') "[set_output_folder] error making directory "//trim(foldername)
         error stop
      end if
      grid%foldername = foldername
   end subroutine

end module plbm_lattice
