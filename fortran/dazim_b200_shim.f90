!> iso_c_binding shim: the reference's forward-modelling subroutines re-declared as thin Fortran
!! wrappers over the C ABI of libdazim_b200.so (include/dazim_b200.h).
!!
!! Two ways to drop the library into the reference (INTEGRATION.md):
!!  (1) link-only: libdazim_b200.so already exports the gfortran-mangled symbols
!!      calsurfg_, calsurfganisojoint_, fwdobstraveltimecps_, depthkernel_, depthkernelti_
!!      with the reference's argument lists -- remove the corresponding .o files from the
!!      Makefile and add -ldazim_b200; this file is not needed.
!!  (2) compiler-independent: compile THIS file instead of CalSurfG.f90 / CalSurfGAniso_Joint.f90 /
!!      FwdTraveltimeCPS.f90 / depthkernelTI.f90 (any Fortran 2003 compiler); it binds the plain C
!!      entry points dazim_create / dazim_gbuild / dazim_depthkernel* by name.
!!
!!  (3) G resident in HBM for the whole outer iteration: subroutine DazimOuterIteration at the end of this file is the
!!      body a maintainer puts inside Main_Jt.f90's `do iter=1,maxiter` (it binds dazim_plan_create / _update_model /
!!      _run / _iterate); python -m dazimsurftomo_b200.invert is the tested twin of that loop.
!!
!! NOTE: no Fortran compiler exists in the build image or on the GPU box, so this file is shipped
!! untested; the tested boundary is the C ABI underneath (tests/test_cabi.py, tests/test_gpu_parity.py).
module dazim_b200
  use iso_c_binding
  implicit none

  !> struct dazim_problem (include/dazim_b200.h)
  type, bind(C) :: dazim_problem
    integer(c_int) :: nx, ny, nz
    type(c_ptr) :: vels
    real(c_float) :: goxd, gozd, dvxd, dvzd
    integer(c_int) :: kmaxRc
    type(c_ptr) :: tRc, depz
    real(c_float) :: minthk
    integer(c_int) :: kmax, nsrc, nrcf
    type(c_ptr) :: periods, nrc1, nsrcsurf1, scxf, sczf, rcxf, rczf
  end type
  !> struct dazim_tables
  type, bind(C) :: dazim_tables
    type(c_ptr) :: pvRc, sen_vs, sen_vp, sen_rho, Lsen_Gsc
  end type
  !> struct dazim_coo
  type, bind(C) :: dazim_coo
    type(c_ptr) :: rw, iw_row, col
    integer(c_long_long) :: maxnar, nar
  end type

  !> struct dazim_lsmr_info / dazim_iter_params / dazim_iter_stats (device-resident outer iteration, Main_Jt.f90:416-727)
  type, bind(C) :: dazim_lsmr_info
    integer(c_int) :: istop, itn
    real(c_float) :: normA, condA, normr, normAr, normx, setup_ms, solve_ms
  end type
  type, bind(C) :: dazim_iter_params
    integer(c_int) :: iso_inv
    real(c_float) :: weightVs, weightGcs, damp, minvel, maxvel
    integer(c_int) :: use_ref_controls
    real(c_float) :: atol, btol, conlim
    integer(c_int) :: itnlim, localSize
  end type
  type, bind(C) :: dazim_iter_stats
    real(c_float) :: before(4), after(4), meandeltaT, mean_weight, meanabs_weighted, norms(6), res2Nm, resW2Nm
    real(c_float) :: meanabs_Taa, meanabs_Tvs
    integer(c_long_long) :: nar1, nar
    integer(c_int) :: count3
    type(dazim_lsmr_info) :: lsmr
    real(c_float) :: step_ms, scale_ms
  end type

  interface
    !> plan API: G stays in HBM between the G build and the solve (include/dazim_b200.h)
    integer(c_int) function dazim_plan_create(h, mode, p, tables, Gctrue, Gstrue, src_begin, src_end, plan) &
        bind(C, name='dazim_plan_create')
      import :: c_ptr, c_int, c_long_long, dazim_problem, dazim_tables
      type(c_ptr), value :: h, Gctrue, Gstrue
      integer(c_int), value :: mode
      type(dazim_problem), intent(in) :: p
      type(dazim_tables), intent(in) :: tables
      integer(c_long_long), value :: src_begin, src_end
      type(c_ptr), intent(out) :: plan
    end function
    integer(c_int) function dazim_plan_update_model(plan, vels, tables) bind(C, name='dazim_plan_update_model')
      import :: c_ptr, c_int, dazim_tables
      type(c_ptr), value :: plan, vels
      type(dazim_tables), intent(in) :: tables
    end function
    integer(c_int) function dazim_plan_run(plan) bind(C, name='dazim_plan_run')
      import :: c_ptr, c_int
      type(c_ptr), value :: plan
    end function
    integer(c_int) function dazim_plan_iterate(plan, obst, prm, vsf, dv, gcf, gsf, dws, sigmaT, resbst, fwdTvs, fwdTaa, &
                                               stats) bind(C, name='dazim_plan_iterate')
      import :: c_ptr, c_int, dazim_iter_params, dazim_iter_stats
      type(c_ptr), value :: plan, obst, vsf, dv, gcf, gsf, dws, sigmaT, resbst, fwdTvs, fwdTaa
      type(dazim_iter_params), intent(in) :: prm
      type(dazim_iter_stats), intent(out) :: stats
    end function
    subroutine dazim_plan_destroy(plan) bind(C, name='dazim_plan_destroy')
      import :: c_ptr
      type(c_ptr), value :: plan
    end subroutine
    integer(c_int) function dazim_create(h, device) bind(C, name='dazim_create')
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: h
      integer(c_int), value :: device
    end function
    integer(c_int) function dazim_gbuild(h, mode, p, tables, tables_precomputed, Gctrue, Gstrue, dsurf, obsTaa, &
                                         tRcV, coo) bind(C, name='dazim_gbuild')
      import :: c_ptr, c_int, dazim_problem, dazim_tables
      type(c_ptr), value :: h
      integer(c_int), value :: mode, tables_precomputed
      type(dazim_problem), intent(in) :: p
      type(dazim_tables), intent(inout) :: tables
      type(c_ptr), value :: Gctrue, Gstrue, dsurf, obsTaa, tRcV, coo
    end function
    integer(c_int) function dazim_depthkernel(h, nx, ny, nz, vel, pvRc, sen_vs, sen_vp, sen_rho, kmaxRc, tRc, depz, &
                                              minthk) bind(C, name='dazim_depthkernel')
      import :: c_ptr, c_int, c_float
      type(c_ptr), value :: h, vel, pvRc, sen_vs, sen_vp, sen_rho, tRc, depz
      integer(c_int), value :: nx, ny, nz, kmaxRc
      real(c_float), value :: minthk
    end function
    integer(c_int) function dazim_depthkernel_ti(h, nx, ny, nz, vel, pvRc, kmaxRc, tRc, depz, minthk, Lsen_Gsc) &
        bind(C, name='dazim_depthkernel_ti')
      import :: c_ptr, c_int, c_float
      type(c_ptr), value :: h, vel, pvRc, tRc, depz, Lsen_Gsc
      integer(c_int), value :: nx, ny, nz, kmaxRc
      real(c_float), value :: minthk
    end function
    function dazim_strerror(code) bind(C, name='dazim_strerror') result(s)
      import :: c_ptr, c_int
      integer(c_int), value :: code
      type(c_ptr) :: s
    end function
  end interface

  type(c_ptr), save :: dz_handle = c_null_ptr

contains

  subroutine dz_init()
    integer(c_int) :: st
    if (.not. c_associated(dz_handle)) then
      st = dazim_create(dz_handle, 0_c_int)
      if (st /= 0) call dz_stop(st, 'dazim_create')
    end if
  end subroutine

  !> the reference STOPs inside the callee (CalSurfG.f90:287-293, Main_Jt.f90:523)
  subroutine dz_stop(st, where)
    integer(c_int), intent(in) :: st
    character(len=*), intent(in) :: where
    write(*,*) 'dazim_b200 error ', st, ' in ', where
    write(*,*) 'TERMINATING PROGRAM!!!'
    stop 1
  end subroutine

  subroutine dz_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                        scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf)
    type(dazim_problem), intent(out) :: p
    integer, intent(in) :: nx, ny, nz, kmaxRc, kmax, nsrcsurf, nrcf
    real, intent(in) :: goxdf, gozdf, dvxdf, dvzdf, minthk
    real, target, intent(in) :: vels(nx, ny, nz), depz(nz)
    real*8, target, intent(in) :: tRc(kmaxRc)
    integer, target, intent(in) :: periods(nsrcsurf, kmax), nrc1(nsrcsurf, kmax), nsrcsurf1(kmax)
    real, target, intent(in) :: scxf(nsrcsurf, kmax), sczf(nsrcsurf, kmax)
    real, target, intent(in) :: rcxf(nrcf, nsrcsurf, kmax), rczf(nrcf, nsrcsurf, kmax)
    p%nx = nx; p%ny = ny; p%nz = nz; p%vels = c_loc(vels)
    p%goxd = goxdf; p%gozd = gozdf; p%dvxd = dvxdf; p%dvzd = dvzdf
    p%kmaxRc = kmaxRc; p%tRc = c_loc(tRc); p%depz = c_loc(depz); p%minthk = minthk
    p%kmax = kmax; p%nsrc = nsrcsurf; p%nrcf = nrcf
    p%periods = c_loc(periods); p%nrc1 = c_loc(nrc1); p%nsrcsurf1 = c_loc(nsrcsurf1)
    p%scxf = c_loc(scxf); p%sczf = c_loc(sczf); p%rcxf = c_loc(rcxf); p%rczf = c_loc(rczf)
  end subroutine
end module dazim_b200

!> CalSurfG (src/src_inv_iso_joint/CalSurfG.f90:909-912), same dummy-argument list.
subroutine CalSurfG(nx, ny, nz, nparpi, vels, iw, rw, col, dsurf, GVs, dall, &
                    goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                    scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf, nar)
  use dazim_b200
  implicit none
  integer :: nx, ny, nz, nparpi, dall, kmaxRc, kmax, nsrcsurf, nrcf, nar
  real, target :: vels(nx, ny, nz), rw(*), dsurf(*), GVs(dall, nparpi), depz(nz)
  integer, target :: iw(*), col(*)
  real :: goxdf, gozdf, dvxdf, dvzdf, minthk
  real*8, target :: tRc(kmaxRc)
  integer, target :: periods(nsrcsurf, kmax), nrc1(nsrcsurf, kmax), nsrcsurf1(kmax)
  real, target :: scxf(nsrcsurf, kmax), sczf(nsrcsurf, kmax), rcxf(nrcf, nsrcsurf, kmax), rczf(nrcf, nsrcsurf, kmax)
  type(dazim_problem) :: p
  type(dazim_tables) :: tb
  type(dazim_coo), target :: c
  integer(c_int) :: st
  integer(c_long_long) :: k
  call dz_init()
  call dz_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                  scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf)
  tb%pvRc = c_null_ptr; tb%sen_vs = c_null_ptr; tb%sen_vp = c_null_ptr; tb%sen_rho = c_null_ptr
  tb%Lsen_Gsc = c_null_ptr
  c%rw = c_loc(rw); c%iw_row = c_loc(iw(2)); c%col = c_loc(col)      ! rows live at iw(2:nar+1) (Main_Jt.f90:529)
  c%maxnar = huge(1_c_long_long); c%nar = 0
  st = dazim_gbuild(dz_handle, 1_c_int, p, tb, 0_c_int, c_null_ptr, c_null_ptr, c_loc(dsurf), c_null_ptr, &
                    c_null_ptr, c_loc(c))
  if (st /= 0) call dz_stop(st, 'CalSurfG')
  nar = int(c%nar)
  GVs = 0.0                                                           ! dense copy only feeds printed statistics
  do k = 1, c%nar
    GVs(iw(1 + k), col(k)) = rw(k)
  end do
end subroutine

!> CalSurfGAnisoJoint (src/src_inv_iso_joint/CalSurfGAniso_Joint.f90:209-212)
subroutine CalSurfGAnisoJoint(nx, ny, nz, nparpi, vels, iw, rw, col, dsurf, GVs, GGc, GGs, Lsen_Gsc, dall, rmax, tRcV, &
                              goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                              scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf, nar, writepath)
  use dazim_b200
  implicit none
  integer :: nx, ny, nz, nparpi, dall, rmax, kmaxRc, kmax, nsrcsurf, nrcf, nar, writepath
  real, target :: vels(nx, ny, nz), rw(*), dsurf(*), depz(nz)
  real, target :: GVs(dall, nparpi), GGc(dall, nparpi), GGs(dall, nparpi), Lsen_Gsc(nx*ny, kmaxRc, nz - 1)
  real*8, target :: tRcV((nx - 2)*(ny - 2), kmaxRc), tRc(kmaxRc)
  integer, target :: iw(*), col(*)
  real :: goxdf, gozdf, dvxdf, dvzdf, minthk
  integer, target :: periods(nsrcsurf, kmax), nrc1(nsrcsurf, kmax), nsrcsurf1(kmax)
  real, target :: scxf(nsrcsurf, kmax), sczf(nsrcsurf, kmax), rcxf(nrcf, nsrcsurf, kmax), rczf(nrcf, nsrcsurf, kmax)
  type(dazim_problem) :: p
  type(dazim_tables) :: tb
  type(dazim_coo), target :: c
  integer(c_int) :: st
  integer(c_long_long) :: k
  integer :: blk, cc
  call dz_init()
  call dz_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                  scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf)
  tb%pvRc = c_null_ptr; tb%sen_vs = c_null_ptr; tb%sen_vp = c_null_ptr; tb%sen_rho = c_null_ptr
  tb%Lsen_Gsc = c_loc(Lsen_Gsc)
  c%rw = c_loc(rw); c%iw_row = c_loc(iw(2)); c%col = c_loc(col)
  c%maxnar = huge(1_c_long_long); c%nar = 0
  st = dazim_gbuild(dz_handle, 2_c_int, p, tb, 0_c_int, c_null_ptr, c_null_ptr, c_loc(dsurf), c_null_ptr, &
                    c_loc(tRcV), c_loc(c))
  if (st /= 0) call dz_stop(st, 'CalSurfGAnisoJoint')
  nar = int(c%nar)
  GVs = 0.0; GGc = 0.0; GGs = 0.0
  do k = 1, c%nar
    blk = (col(k) - 1)/nparpi; cc = col(k) - blk*nparpi
    if (blk == 0) GVs(iw(1 + k), cc) = rw(k)
    if (blk == 1) GGc(iw(1 + k), cc) = rw(k)
    if (blk == 2) GGs(iw(1 + k), cc) = rw(k)
  end do
end subroutine

!> FwdObsTraveltimeCPS (src/src_forward/FwdTraveltimeCPS.f90:208-211); writepath is LOGICAL here.
subroutine FwdObsTraveltimeCPS(nx, ny, nz, nparpi, vels, Gctrue, Gstrue, dsurf, obsTaa, dall, rmax, tRcV, Lsen_Gsc, &
                               goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                               scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf, writepath)
  use dazim_b200
  implicit none
  integer :: nx, ny, nz, nparpi, dall, rmax, kmaxRc, kmax, nsrcsurf, nrcf
  logical :: writepath
  real, target :: vels(nx, ny, nz), Gctrue(nx - 2, ny - 2, nz - 1), Gstrue(nx - 2, ny - 2, nz - 1)
  real, target :: dsurf(*), obsTaa(*), depz(nz), Lsen_Gsc(nx*ny, kmaxRc, nz - 1)
  real*8, target :: tRcV((nx - 2)*(ny - 2), kmaxRc), tRc(kmaxRc)
  real :: goxdf, gozdf, dvxdf, dvzdf, minthk
  integer, target :: periods(nsrcsurf, kmax), nrc1(nsrcsurf, kmax), nsrcsurf1(kmax)
  real, target :: scxf(nsrcsurf, kmax), sczf(nsrcsurf, kmax), rcxf(nrcf, nsrcsurf, kmax), rczf(nrcf, nsrcsurf, kmax)
  type(dazim_problem) :: p
  type(dazim_tables) :: tb
  integer(c_int) :: st
  call dz_init()
  call dz_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, &
                  scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf)
  tb%pvRc = c_null_ptr; tb%sen_vs = c_null_ptr; tb%sen_vp = c_null_ptr; tb%sen_rho = c_null_ptr
  tb%Lsen_Gsc = c_loc(Lsen_Gsc)
  st = dazim_gbuild(dz_handle, 0_c_int, p, tb, 0_c_int, c_loc(Gctrue), c_loc(Gstrue), c_loc(dsurf), c_loc(obsTaa), &
                    c_loc(tRcV), c_null_ptr)
  if (st /= 0) call dz_stop(st, 'FwdObsTraveltimeCPS')
end subroutine

!> depthkernel (src/src_inv_iso_joint/CalSurfG.f90:1-2)
subroutine depthkernel(nx, ny, nz, vel, pvRc, sen_vsRc, sen_vpRc, sen_rhoRc, iwave, igr, kmaxRc, tRc, depz, minthk)
  use dazim_b200
  implicit none
  integer :: nx, ny, nz, iwave, igr, kmaxRc
  real, target :: vel(nx, ny, nz), depz(nz)
  real :: minthk
  real*8, target :: pvRc(nx*ny, kmaxRc), sen_vsRc(nx*ny, kmaxRc, nz), sen_vpRc(nx*ny, kmaxRc, nz)
  real*8, target :: sen_rhoRc(nx*ny, kmaxRc, nz), tRc(kmaxRc)
  integer(c_int) :: st
  call dz_init()
  st = dazim_depthkernel(dz_handle, nx, ny, nz, c_loc(vel), c_loc(pvRc), c_loc(sen_vsRc), c_loc(sen_vpRc), &
                         c_loc(sen_rhoRc), kmaxRc, c_loc(tRc), c_loc(depz), minthk)
  if (st /= 0) call dz_stop(st, 'depthkernel')
end subroutine

!> depthkernelTI (src/src_forward/depthkernelTI.f90:2)
subroutine depthkernelTI(nx, ny, nz, vel, pvRc, iwave, igr, kmaxRc, tRc, depz, minthk, Lsen_Gsc)
  use dazim_b200
  implicit none
  integer :: nx, ny, nz, iwave, igr, kmaxRc
  real, target :: vel(nx, ny, nz), depz(nz), Lsen_Gsc(nx*ny, kmaxRc, nz - 1)
  real :: minthk
  real*8, target :: pvRc(nx*ny, kmaxRc), tRc(kmaxRc)
  integer(c_int) :: st
  call dz_init()
  st = dazim_depthkernel_ti(dz_handle, nx, ny, nz, c_loc(vel), c_loc(pvRc), kmaxRc, c_loc(tRc), c_loc(depz), minthk, &
                            c_loc(Lsen_Gsc))
  if (st /= 0) call dz_stop(st, 'depthkernelTI')
end subroutine


!> Body of the outer iteration (Main_Jt.f90:384-727) with G resident in HBM: what a maintainer puts inside
!! `do iter=1,maxiter` instead of CalSurfG / CalSurfGAnisoJoint + CalDdatSigma + the weighting loops + Tikhonov + LSMR +
!! the model update + the CalSigamNorm diagnostics.  `plan` is kept across iterations (c_null_ptr before the first).
subroutine DazimOuterIteration(plan, iter, iso_inv, nx, ny, nz, vsf, obst, dall, goxd, gozd, dvxd, dvzd, kmaxRc, tRc, &
                               periods, depz, minthk, scxf, sczf, rcxf, rczf, nrc1, nsrc1, kmax, nsrc, nrc, &
                               weightVs, weightGcs, damp, Minvel, Maxvel, dv, gcf, gsf, stats)
  use dazim_b200
  implicit none
  type(c_ptr), intent(inout) :: plan
  integer, intent(in) :: iter, nx, ny, nz, dall, kmaxRc, kmax, nsrc, nrc
  logical, intent(in) :: iso_inv
  real, target, intent(inout) :: vsf(nx, ny, nz)
  real, target, intent(in) :: obst(dall), depz(nz), scxf(nsrc, kmax), sczf(nsrc, kmax), rcxf(nrc, nsrc, kmax), &
                              rczf(nrc, nsrc, kmax)
  real, intent(in) :: goxd, gozd, dvxd, dvzd, minthk, weightVs, weightGcs, damp, Minvel, Maxvel
  real*8, target, intent(in) :: tRc(kmaxRc)
  integer, target, intent(in) :: periods(nsrc, kmax), nrc1(nsrc, kmax), nsrc1(kmax)
  real, target, intent(out) :: dv(*), gcf(nx-2, ny-2, nz-1), gsf(nx-2, ny-2, nz-1)
  type(dazim_iter_stats), intent(out) :: stats
  type(dazim_problem) :: p
  type(dazim_tables) :: tb
  type(dazim_iter_params) :: prm
  real*8, allocatable, target, save :: pvRc(:,:), svs(:,:,:), svp(:,:,:), srho(:,:,:)
  real, allocatable, target, save :: Lsen(:,:,:)
  integer(c_int) :: st, mode
  call dz_init()
  if (.not. allocated(pvRc)) then
    allocate(pvRc(nx*ny, kmaxRc), svs(nx*ny, kmaxRc, nz), svp(nx*ny, kmaxRc, nz), srho(nx*ny, kmaxRc, nz), &
             Lsen(nx*ny, kmaxRc, nz-1))
  end if
  if (.not. iso_inv) then
    st = dazim_depthkernel_ti(dz_handle, nx, ny, nz, c_loc(vsf), c_loc(pvRc), kmaxRc, c_loc(tRc), c_loc(depz), minthk, &
                              c_loc(Lsen))
    if (st /= 0) call dz_stop(st, 'depthkernelTI')
  end if
  st = dazim_depthkernel(dz_handle, nx, ny, nz, c_loc(vsf), c_loc(pvRc), c_loc(svs), c_loc(svp), c_loc(srho), kmaxRc, &
                         c_loc(tRc), c_loc(depz), minthk)
  if (st /= 0) call dz_stop(st, 'depthkernel')
  tb%pvRc = c_loc(pvRc); tb%sen_vs = c_loc(svs); tb%sen_vp = c_loc(svp); tb%sen_rho = c_loc(srho); tb%Lsen_Gsc = c_loc(Lsen)
  mode = 2
  if (iso_inv) mode = 1
  if (iter == 1 .or. .not. c_associated(plan)) then
    call dz_problem(p, nx, ny, nz, vsf, goxd, gozd, dvxd, dvzd, kmaxRc, tRc, periods, depz, minthk, scxf, sczf, rcxf, &
                    rczf, nrc1, nsrc1, kmax, nsrc, nrc)
    st = dazim_plan_create(dz_handle, mode, p, tb, c_null_ptr, c_null_ptr, 0_c_long_long, -1_c_long_long, plan)
  else
    st = dazim_plan_update_model(plan, c_loc(vsf), tb)
  end if
  if (st /= 0) call dz_stop(st, 'plan')
  st = dazim_plan_run(plan)                        ! CalSurfG / CalSurfGAnisoJoint: G stays on the GPU
  if (st /= 0) call dz_stop(st, 'G build')
  prm%iso_inv = 0
  if (iso_inv) prm%iso_inv = 1
  prm%weightVs = weightVs; prm%weightGcs = weightGcs; prm%damp = damp; prm%minvel = Minvel; prm%maxvel = Maxvel
  prm%use_ref_controls = 1                         ! atol/btol/conlim/itnlim/localSize of Main_Jt.f90:542-554
  st = dazim_plan_iterate(plan, c_loc(obst), prm, c_loc(vsf), c_loc(dv), c_loc(gcf), c_loc(gsf), c_null_ptr, c_null_ptr, &
                          c_null_ptr, c_null_ptr, c_null_ptr, stats)
  if (st /= 0) call dz_stop(st, 'outer iteration')
end subroutine
