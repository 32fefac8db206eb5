!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
! bgc_b200_runtime - state shared by the three shim modules (BGC_mod, DMS_mod,
! MACROS_mod): the bgc_ctx handle of this MPI rank's GPU, error handling and the
! c_loc helpers that turn allocatable components into C pointers.
!
! One host thread drives one ctx (the reference itself is not re-entrant:
! co2calc keeps SAVE scratch, co2calc.F90:65-67).  The device is chosen by
! bgc_b200_set_device, else by the environment variable BGC_B200_DEVICE, else
! device 0; with one MPI rank per GPU the host passes its node-local rank.
!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
module bgc_b200_runtime
  use, intrinsic :: iso_c_binding
  use bgc_b200_capi
  implicit none
  private
  public :: bgc_b200_ctx, bgc_b200_set_device, bgc_b200_check, bgc_b200_finalize
  public :: loc1, loc2, loc3, loci1, bgc_b200_ctx_levels, bgc_b200_ctx_columns

  type(c_ptr), save :: ctx = c_null_ptr
  integer(c_int), save :: ctx_levels = 0, ctx_columns = 0, device = -1

contains

  subroutine bgc_b200_set_device(dev)
    integer, intent(in) :: dev
    device = int(dev, c_int)
  end subroutine bgc_b200_set_device

  integer function bgc_b200_ctx_levels()
    bgc_b200_ctx_levels = ctx_levels
  end function
  integer function bgc_b200_ctx_columns()
    bgc_b200_ctx_columns = ctx_columns
  end function

  ! The ctx (persistent device arena) for blocks of up to (nLevelsMax, nColumnsMax);
  ! re-created only when a larger block shows up.  `fresh` tells the caller that the
  ! parameter tables must be uploaded again.
  function bgc_b200_ctx(nLevelsMax, nColumnsMax, fresh) result(h)
    integer, intent(in) :: nLevelsMax, nColumnsMax
    logical, intent(out) :: fresh
    type(c_ptr) :: h
    character(len=32) :: env
    integer :: stat, ios
    fresh = .false.
    if (c_associated(ctx) .and. nLevelsMax <= ctx_levels .and. nColumnsMax <= ctx_columns) then
      h = ctx
      return
    end if
    if (c_associated(ctx)) call bgc_b200_check(bgc_ctx_destroy(ctx), 'bgc_ctx_destroy')
    if (device < 0) then
      device = 0
      call get_environment_variable('BGC_B200_DEVICE', env, status=stat)
      if (stat == 0) then
        read(env, *, iostat=ios) device
        if (ios /= 0) device = 0
      end if
    end if
    ctx_levels = max(ctx_levels, int(nLevelsMax, c_int))
    ctx_columns = max(ctx_columns, int(nColumnsMax, c_int))
    call bgc_b200_check(bgc_ctx_create(device, ctx_levels, ctx_columns, ctx), 'bgc_ctx_create')
    fresh = .true.
    h = ctx
  end function bgc_b200_ctx

  subroutine bgc_b200_finalize()
    if (c_associated(ctx)) call bgc_b200_check(bgc_ctx_destroy(ctx), 'bgc_ctx_destroy')
    ctx = c_null_ptr
    ctx_levels = 0
    ctx_columns = 0
  end subroutine bgc_b200_finalize

  ! The reference has no error reporting at all; a failed GPU call must not be
  ! silent (there is no CPU fallback): print the library's message and stop.
  subroutine bgc_b200_check(rc, what)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    type(c_ptr) :: p
    integer :: n
    if (rc == BGC_OK) return
    p = bgc_last_error()
    write(0, '(a,a,a,i0)') 'bgc_b200: ', what, ' failed with code ', rc
    if (c_associated(p)) then
      call c_f_pointer(p, msg, [512])
      n = 1
      do while (n < 512 .and. msg(n) /= c_null_char)
        n = n + 1
      end do
      write(0, '(512a1)') msg(1:n-1)
    end if
    error stop 'bgc_b200: GPU hot path failed'
  end subroutine bgc_b200_check

  ! c_loc of an allocatable component; NULL when it is not allocated (the C ABI
  ! treats a NULL diagnostic as "do not produce").  Allocatable arrays are
  ! contiguous, so c_loc is legal (Fortran 2008 15.2.3.6).
  function loc1(a) result(p)
    real(c_double), allocatable, target, intent(in) :: a(:)
    type(c_ptr) :: p
    p = c_null_ptr
    if (allocated(a)) then
      if (size(a) > 0) p = c_loc(a)
    end if
  end function loc1
  function loc2(a) result(p)
    real(c_double), allocatable, target, intent(in) :: a(:,:)
    type(c_ptr) :: p
    p = c_null_ptr
    if (allocated(a)) then
      if (size(a) > 0) p = c_loc(a)
    end if
  end function loc2
  function loc3(a) result(p)
    real(c_double), allocatable, target, intent(in) :: a(:,:,:)
    type(c_ptr) :: p
    p = c_null_ptr
    if (allocated(a)) then
      if (size(a) > 0) p = c_loc(a)
    end if
  end function loc3
  function loci1(a) result(p)
    integer(c_int), allocatable, target, intent(in) :: a(:)
    type(c_ptr) :: p
    p = c_null_ptr
    if (allocated(a)) then
      if (size(a) > 0) p = c_loc(a)
    end if
  end function loci1

end module bgc_b200_runtime
