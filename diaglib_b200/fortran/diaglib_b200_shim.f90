!
! diaglib_b200_shim.f90 -- drop-in `module diaglib` for Fortran hosts.
!
! Provides lobpcg_driver / davidson_driver / ortho / b_ortho / ortho_cd / ortho_vs_x /
! b_ortho_vs_x with exactly the argument lists of Molecolab-Pisa/diaglib (diaglib.f90:171-172,
! 1483-1484, 3052, 3094, 3185, 3481, 3576)
! and forwards them to the C-ABI of libdiaglib_b200.so (include/diaglib_b200.h) through
! iso_c_binding.  A program that does `use diaglib` and links this module plus
! -ldiaglib_b200 instead of the reference diaglib.o runs the B200 path unchanged.
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler is installed); the C-ABI it
! binds is exercised from C/ctypes in tests/.  Build line for a host that has gfortran:
!   gfortran -c real_precision.f90 diaglib_b200_shim.f90
!   gfortran main.o real_precision.o diaglib_b200_shim.o -L. -ldiaglib_b200 -o main.exe
!
! Contract differences from the CPU reference (see INTEGRATION.md):
!   * matvec / precnd receive DEVICE pointers (x, ax live in HBM).  Pass
!     diaglib_b200_csr_matvec / diaglib_b200_diag_precnd for the built-in CSR callbacks, or
!     device-aware routines of your own (CUDA Fortran, OpenACC host_data, C wrappers).
!   * n is the number of rows owned by the calling rank.
!   * where the reference executes `stop` the library returns ok = .false.; this shim
!     restores the reference behaviour by stopping when diaglib_b200_last_status() /= 0.
!
module diaglib
  use iso_c_binding
  implicit none
  private
  public :: lobpcg_driver, davidson_driver, gen_david_driver, caslr_eff_driver, ortho, b_ortho, ortho_cd, ortho_vs_x, b_ortho_vs_x
  public :: diaglib_b200_init, diaglib_b200_set_csr, diaglib_b200_set_csr_b
  public :: diaglib_b200_csr_matvec, diaglib_b200_csr_bvec, diaglib_b200_diag_precnd
!
  interface
    subroutine c_lobpcg(verbose, gen_eig, n, n_targ, n_max, max_iter, tol, shift, matvec, precnd, bvec, &
                        eig, evec, ok) bind(C, name='diaglib_b200_lobpcg_driver')
      import :: c_int32_t, c_double, c_funptr
      integer(c_int32_t), intent(in)    :: verbose, gen_eig, n, n_targ, n_max, max_iter
      real(c_double),     intent(in)    :: tol, shift
      type(c_funptr),     value         :: matvec, precnd, bvec
      real(c_double),     intent(inout) :: eig(*), evec(*)
      integer(c_int32_t), intent(inout) :: ok
    end subroutine c_lobpcg
    subroutine c_davidson(verbose, n, n_targ, n_max, max_iter, tol, max_dav, shift, matvec, precnd, &
                          eig, evec, ok) bind(C, name='diaglib_b200_davidson_driver')
      import :: c_int32_t, c_double, c_funptr
      integer(c_int32_t), intent(in)    :: verbose, n, n_targ, n_max, max_iter, max_dav
      real(c_double),     intent(in)    :: tol, shift
      type(c_funptr),     value         :: matvec, precnd
      real(c_double),     intent(inout) :: eig(*), evec(*)
      integer(c_int32_t), intent(inout) :: ok
    end subroutine c_davidson
    subroutine c_gen_david(verbose, n, n_targ, n_max, max_iter, tol, max_dav, shift, matvec, precnd, bvec, &
                           eig, evec, ok) bind(C, name='diaglib_b200_gen_david_driver')
      import :: c_int32_t, c_double, c_funptr
      integer(c_int32_t), intent(in)    :: verbose, n, n_targ, n_max, max_iter, max_dav
      real(c_double),     intent(in)    :: tol, shift
      type(c_funptr),     value         :: matvec, precnd, bvec
      real(c_double),     intent(inout) :: eig(*), evec(*)
      integer(c_int32_t), intent(inout) :: ok
    end subroutine c_gen_david
    subroutine c_caslr_eff(verbose, n, n2, n_targ, n_max, max_iter, tol, max_dav, apbmul, ambmul, spdmul, smdmul, &
                           lrprec, eig, evec, ok) bind(C, name='diaglib_b200_caslr_eff_driver')
      import :: c_int32_t, c_double, c_funptr
      integer(c_int32_t), intent(in)    :: verbose, n, n2, n_targ, n_max, max_iter, max_dav
      real(c_double),     intent(in)    :: tol
      type(c_funptr),     value         :: apbmul, ambmul, spdmul, smdmul, lrprec
      real(c_double),     intent(inout) :: eig(*), evec(*)
      integer(c_int32_t), intent(inout) :: ok
    end subroutine c_caslr_eff
    subroutine c_ortho_cd(n, m, u, growth, ok) bind(C, name='diaglib_b200_ortho_cd')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m
      real(c_double),     intent(inout) :: u(*), growth
      integer(c_int32_t), intent(inout) :: ok
    end subroutine c_ortho_cd
    subroutine c_ortho_vs_x(n, m, k, x, u, ax, au) bind(C, name='diaglib_b200_ortho_vs_x')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m, k
      real(c_double),     intent(in)    :: x(*), ax(*)
      real(c_double),     intent(inout) :: u(*), au(*)
    end subroutine c_ortho_vs_x
    subroutine c_b_ortho(n, m, u, bu) bind(C, name='diaglib_b200_b_ortho')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m
      real(c_double),     intent(inout) :: u(*), bu(*)
    end subroutine c_b_ortho
    subroutine c_b_ortho_vs_x(n, m, k, x, bx, u) bind(C, name='diaglib_b200_b_ortho_vs_x')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m, k
      real(c_double),     intent(in)    :: x(*), bx(*)
      real(c_double),     intent(inout) :: u(*)
    end subroutine c_b_ortho_vs_x
    subroutine c_ortho(n, m, u, w) bind(C, name='diaglib_b200_ortho')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m
      real(c_double),     intent(inout) :: u(*), w(*)
    end subroutine c_ortho
    function diaglib_b200_last_status() bind(C, name='diaglib_b200_last_status') result(st)
      import :: c_int32_t
      integer(c_int32_t) :: st
    end function diaglib_b200_last_status
    function diaglib_b200_init(device) bind(C, name='diaglib_b200_init') result(st)
      import :: c_int32_t
      integer(c_int32_t), value :: device
      integer(c_int32_t)        :: st
    end function diaglib_b200_init
    function diaglib_b200_set_csr(n_loc, n_halo, rowptr, col, val, diag) &
             bind(C, name='diaglib_b200_set_csr') result(st)
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), value      :: n_loc, n_halo
      integer(c_int64_t), intent(in) :: rowptr(*)
      integer(c_int32_t), intent(in) :: col(*)
      real(c_double),     intent(in) :: val(*), diag(*)
      integer(c_int32_t)             :: st
    end function diaglib_b200_set_csr
    function diaglib_b200_set_csr_b(n_loc, n_halo, rowptr, col, val) &
             bind(C, name='diaglib_b200_set_csr_b') result(st)
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), value      :: n_loc, n_halo
      integer(c_int64_t), intent(in) :: rowptr(*)
      integer(c_int32_t), intent(in) :: col(*)
      real(c_double),     intent(in) :: val(*)
      integer(c_int32_t)             :: st
    end function diaglib_b200_set_csr_b
    ! built-in conforming callbacks (device pointers); pass them as matvec / precnd / bvec
    subroutine diaglib_b200_csr_bvec(n, m, x, bx) bind(C, name='diaglib_b200_csr_bvec')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m
      real(c_double),     intent(in)    :: x(*)
      real(c_double),     intent(inout) :: bx(*)
    end subroutine diaglib_b200_csr_bvec
    subroutine diaglib_b200_csr_matvec(n, m, x, ax) bind(C, name='diaglib_b200_csr_matvec')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m
      real(c_double),     intent(in)    :: x(*)
      real(c_double),     intent(inout) :: ax(*)
    end subroutine diaglib_b200_csr_matvec
    subroutine diaglib_b200_diag_precnd(n, m, shift, x, px) bind(C, name='diaglib_b200_diag_precnd')
      import :: c_int32_t, c_double
      integer(c_int32_t), intent(in)    :: n, m
      real(c_double),     intent(in)    :: shift, x(*)
      real(c_double),     intent(inout) :: px(*)
    end subroutine diaglib_b200_diag_precnd
  end interface
!
contains
!
  subroutine check_stop(who)
    character(len=*), intent(in) :: who
    integer(c_int32_t) :: st
    st = diaglib_b200_last_status()
    if (st .ne. 0) then
      write(6,'(t3,a,a,i4)') who, ' failed. status = ', st
      stop
    end if
  end subroutine check_stop
!
! diaglib.f90:171-172
  subroutine lobpcg_driver(verbose,gen_eig,n,n_targ,n_max,max_iter,tol,shift,matvec,precnd,bvec,eig,evec,ok)
    logical,  intent(in)    :: verbose, gen_eig
    integer,  intent(in)    :: n, n_targ, n_max, max_iter
    real(8),  intent(in)    :: tol, shift
    real(8),  intent(inout) :: eig(n_max), evec(n,n_max)
    logical,  intent(inout) :: ok
    external                :: matvec, precnd, bvec
    integer(c_int32_t)      :: iok
    iok = 0
    call c_lobpcg(merge(1,0,verbose), merge(1,0,gen_eig), n, n_targ, n_max, max_iter, tol, shift, &
                  c_funloc(matvec), c_funloc(precnd), c_funloc(bvec), eig, evec, iok)
    call check_stop('lobpcg_driver')
    ok = iok .ne. 0
  end subroutine lobpcg_driver
!
! diaglib.f90:1483-1484
  subroutine davidson_driver(verbose,n,n_targ,n_max,max_iter,tol,max_dav,shift,matvec,precnd,eig,evec,ok)
    logical,  intent(in)    :: verbose
    integer,  intent(in)    :: n, n_targ, n_max, max_iter, max_dav
    real(8),  intent(in)    :: tol, shift
    real(8),  intent(inout) :: eig(n_max), evec(n,n_max)
    logical,  intent(inout) :: ok
    external                :: matvec, precnd
    integer(c_int32_t)      :: iok
    iok = 0
    call c_davidson(merge(1,0,verbose), n, n_targ, n_max, max_iter, tol, max_dav, shift, &
                    c_funloc(matvec), c_funloc(precnd), eig, evec, iok)
    call check_stop('davidson_driver')
    ok = iok .ne. 0
  end subroutine davidson_driver
!
! diaglib.f90:1855-1856
  subroutine gen_david_driver(verbose,n,n_targ,n_max,max_iter,tol,max_dav,shift,matvec,precnd,bvec,eig,evec,ok)
    logical,  intent(in)    :: verbose
    integer,  intent(in)    :: n, n_targ, n_max, max_iter, max_dav
    real(8),  intent(in)    :: tol, shift
    real(8),  intent(inout) :: eig(n_max), evec(n,n_max)
    logical,  intent(inout) :: ok
    external                :: matvec, precnd, bvec
    integer(c_int32_t)      :: iok
    iok = 0
    call c_gen_david(merge(1,0,verbose), n, n_targ, n_max, max_iter, tol, max_dav, shift, &
                     c_funloc(matvec), c_funloc(precnd), c_funloc(bvec), eig, evec, iok)
    call check_stop('gen_david_driver')
    ok = iok .ne. 0
  end subroutine gen_david_driver
!
! diaglib.f90:1024-1025
  subroutine caslr_eff_driver(verbose,n,n2,n_targ,n_max,max_iter,tol,max_dav, &
                              apbmul,ambmul,spdmul,smdmul,lrprec,eig,evec,ok)
    logical,  intent(in)    :: verbose
    integer,  intent(in)    :: n, n2, n_targ, n_max, max_iter, max_dav
    real(8),  intent(in)    :: tol
    real(8),  intent(inout) :: eig(n_max), evec(n2,n_max)
    logical,  intent(inout) :: ok
    external                :: apbmul, ambmul, spdmul, smdmul, lrprec
    integer(c_int32_t)      :: iok
    iok = 0
    call c_caslr_eff(merge(1,0,verbose), n, n2, n_targ, n_max, max_iter, tol, max_dav, c_funloc(apbmul), &
                     c_funloc(ambmul), c_funloc(spdmul), c_funloc(smdmul), c_funloc(lrprec), eig, evec, iok)
    call check_stop('caslr_eff_driver')
    ok = iok .ne. 0
  end subroutine caslr_eff_driver
!
! diaglib.f90:3185
  subroutine ortho_cd(n,m,u,growth,ok)
    integer,  intent(in)    :: n, m
    real(8),  intent(inout) :: u(n,m), growth
    logical,  intent(inout) :: ok
    integer(c_int32_t)      :: iok
    iok = 0
    call c_ortho_cd(n, m, u, growth, iok)
    call check_stop('ortho_cd')
    ok = iok .ne. 0
  end subroutine ortho_cd
!
! diaglib.f90:3481
  subroutine ortho_vs_x(n,m,k,x,u,ax,au)
    integer,  intent(in)    :: n, m, k
    real(8),  intent(in)    :: x(n,m), ax(*)
    real(8),  intent(inout) :: u(n,k), au(*)
    call c_ortho_vs_x(n, m, k, x, u, ax, au)
    call check_stop('ortho_vs_x')
  end subroutine ortho_vs_x
!
! diaglib.f90:3094
  subroutine b_ortho(n,m,u,bu)
    integer,  intent(in)    :: n, m
    real(8),  intent(inout) :: u(n,m), bu(n,m)
    call c_b_ortho(n, m, u, bu)
    call check_stop('b_ortho')
  end subroutine b_ortho
!
! diaglib.f90:3576
  subroutine b_ortho_vs_x(n,m,k,x,bx,u)
    integer,  intent(in)    :: n, m, k
    real(8),  intent(in)    :: x(n,m), bx(n,m)
    real(8),  intent(inout) :: u(n,k)
    call c_b_ortho_vs_x(n, m, k, x, bx, u)
    call check_stop('b_ortho_vs_x')
  end subroutine b_ortho_vs_x
!
! diaglib.f90:3052
  subroutine ortho(n,m,u,w)
    integer,  intent(in)    :: n, m
    real(8),  intent(inout) :: u(n,m), w(*)
    call c_ortho(n, m, u, w)
    call check_stop('ortho')
  end subroutine ortho
!
end module diaglib
