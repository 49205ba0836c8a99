!> ISO_C_BINDING interface blocks for the C ABI of libplbm_b200.so (include/plbm.h).
!! One interface per exported entry point; names and argument order are those of the header.
!! SOURCE ONLY: no Fortran compiler exists in the image this library was developed in
!! (SURVEY.md F1), so these modules have not been compiled there.  Build line on a machine with
!! gfortran:  gfortran -c -cpp precision.F90 plbm_c.f90 fvm_bardow.F90 periodic_lbm.F90 ...
!!            gfortran app.f90 *.o -L<repo>/periodic_lbm_b200 -lplbm_b200
!! (the full list, with the reference's own output / flow-case modules, is in INTEGRATION.md section 1)
module plbm_c
   use, intrinsic :: iso_c_binding
   implicit none
   public

   integer(c_int), parameter :: PLBM_F64 = 0, PLBM_F32 = 1
   integer(c_int), parameter :: PLBM_BGK = 0, PLBM_TRT = 1, PLBM_RR = 2, PLBM_BGK_SPLIT = 3, &
                                PLBM_TRT_SPLIT = 4, PLBM_BGK_IMPROVED = 5
   integer(c_int), parameter :: PLBM_STREAM_LBM = 0, PLBM_STREAM_FVM_BARDOW = 1, &
                                PLBM_STREAM_FDM_BARDOW = 2, PLBM_STREAM_FDM_SOFONEA = 3

   interface
      function plbm_last_error() bind(c, name="plbm_last_error") result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      function plbm_alloc_grid(grid, nx, ny, nf, precision) bind(c, name="plbm_alloc_grid") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), intent(out) :: grid
         integer(c_int), value :: nx, ny, nf, precision
         integer(c_int) :: stat
      end function
      function plbm_dealloc_grid(grid) bind(c, name="plbm_dealloc_grid") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
      function plbm_get_indices(grid, iold, inew, imid) bind(c, name="plbm_get_indices") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), intent(out) :: iold, inew, imid
         integer(c_int) :: stat
      end function
      function plbm_set_properties(grid, nu, dt, magic, has_magic) bind(c, name="plbm_set_properties") result(stat)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: grid
         real(c_double), value :: nu, dt, magic
         integer(c_int), value :: has_magic
         integer(c_int) :: stat
      end function
      function plbm_get_properties(grid, props) bind(c, name="plbm_get_properties") result(stat)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: grid
         real(c_double), intent(out) :: props(6)   ! nu, dt, tau, omega, trt_magic, csqr
         integer(c_int) :: stat
      end function
      function plbm_set_step_deferral(grid, max_pending) bind(c, name="plbm_set_step_deferral") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: max_pending
         integer(c_int) :: stat
      end function
      function plbm_set_omega(grid, omega) bind(c, name="plbm_set_omega") result(stat)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: grid
         real(c_double), value :: omega
         integer(c_int) :: stat
      end function
      function plbm_set_pdf_to_equilibrium(grid, rho, ux, uy) bind(c, name="plbm_set_pdf_to_equilibrium") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         type(c_ptr), value :: rho, ux, uy        ! c_loc of (ny,nx) arrays of kind wp
         integer(c_int) :: stat
      end function
      function plbm_perform_lbm_step(grid, collision, nsteps) bind(c, name="plbm_perform_lbm_step") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: collision, nsteps
         integer(c_int) :: stat
      end function
      function plbm_perform_step(grid, streaming, collision, nsteps) bind(c, name="plbm_perform_step") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: streaming, collision, nsteps
         integer(c_int) :: stat
      end function
      function plbm_perform_triple_step(grid, streaming, collision, nsteps) bind(c, name="plbm_perform_triple_step") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: streaming, collision, nsteps
         integer(c_int) :: stat
      end function
      function plbm_perform_dugks_step(grid, dugks, nsteps) bind(c, name="plbm_perform_dugks_step") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: dugks, nsteps
         integer(c_int) :: stat
      end function
      function plbm_lbm_stream(grid) bind(c, name="plbm_lbm_stream") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
      function plbm_stream_fvm_bardow(grid) bind(c, name="plbm_stream_fvm_bardow") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
      function plbm_stream_fdm_bardow(grid) bind(c, name="plbm_stream_fdm_bardow") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
      function plbm_set_fdm_stencil(grid, stencil) bind(c, name="plbm_set_fdm_stencil") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: stencil
         integer(c_int) :: stat
      end function
      function plbm_stream_fdm_sofonea(grid) bind(c, name="plbm_stream_fdm_sofonea") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
      function plbm_collide(grid, collision) bind(c, name="plbm_collide") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: collision
         integer(c_int) :: stat
      end function
      function plbm_dugks_collide(grid, dugks) bind(c, name="plbm_dugks_collide") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: dugks
         integer(c_int) :: stat
      end function
      function plbm_dugks_stream(grid, dugks) bind(c, name="plbm_dugks_stream") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: dugks
         integer(c_int) :: stat
      end function
      function plbm_swap(grid) bind(c, name="plbm_swap") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
      function plbm_update_macros(grid, rho, ux, uy, lagged) bind(c, name="plbm_update_macros") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         type(c_ptr), value :: rho, ux, uy
         integer(c_int), value :: lagged
         integer(c_int) :: stat
      end function
      function plbm_vorticity_host(grid, order, ux, uy, omega) bind(c, name="plbm_vorticity_host") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int), value :: order
         type(c_ptr), value :: ux, uy, omega
         integer(c_int) :: stat
      end function
      function plbm_diagnostics(grid, out) bind(c, name="plbm_diagnostics") result(stat)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: grid
         real(c_double), intent(out) :: out(4)    ! max|u|, min|u|, sum(rho), kinetic energy
         integer(c_int) :: stat
      end function
      function plbm_synchronize(grid) bind(c, name="plbm_synchronize") result(stat)
         import :: c_ptr, c_int
         type(c_ptr), value :: grid
         integer(c_int) :: stat
      end function
   end interface

contains

   !> Turn a non-zero status into the reference's failure mode (`error stop`) with the C message.
   subroutine plbm_check(stat, what)
      integer(c_int), intent(in) :: stat
      character(*), intent(in) :: what
      character(kind=c_char), pointer :: cmsg(:)
      character(len=512) :: msg
      integer :: i
      if (stat == 0) return
      msg = ''
      call c_f_pointer(plbm_last_error(), cmsg, [512])
      do i = 1, 512
         if (cmsg(i) == c_null_char) exit
         msg(i:i) = cmsg(i)
      end do
      write(*,'(a)') "plbm: "//what//" failed: "//trim(msg)
      error stop 1
   end subroutine

end module plbm_c
