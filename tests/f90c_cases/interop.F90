! Second self-written test module for oracle/f90c.py: the ISO_C_BINDING / internal-procedure /
! character-dummy / pointer-array constructs that the drop-in shim (ocean-bgc_b200/fortran/) uses.
! The C side (twin functions c_scale, c_fill, c_name) is tests/f90c_cases/interop_c.c.
module interop_types
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: cblock
     real(c_double) :: scale
     integer(c_int) :: n
     type(c_ptr) :: data = c_null_ptr
     type(c_ptr) :: spare = c_null_ptr
  end type cblock
  interface
     integer(c_int) function c_scale(blk, factor) bind(C, name="c_scale")
       import :: c_int, c_double, cblock
       type(cblock), intent(inout) :: blk
       real(c_double), value :: factor
     end function
     function c_name() bind(C, name="c_name") result(p)
       import :: c_ptr
       type(c_ptr) :: p
     end function
     integer(c_int) function c_sum3(v) bind(C, name="c_sum3")
       import :: c_int
       integer(c_int), intent(in) :: v(3)
     end function
  end interface
end module interop_types

module interop_mod
  use, intrinsic :: iso_c_binding
  use interop_types
  implicit none
  private
  public :: scale_through_c, label_all, read_c_string, count_to
  integer, save :: calls = 0
  logical :: seen = .false.
  type holder
     real(c_double), allocatable :: values(:)
     character(8), allocatable :: tags(:)
  end type holder
  public :: holder
contains

  ! allocatable component -> c_loc -> C function working in place; NULL when not allocated
  function scale_through_c(h, factor) result(rc)
    type(holder), intent(inout), target :: h
    real(c_double), intent(in) :: factor
    integer :: rc
    type(cblock) :: blk
    integer(c_int) :: three(3)
    calls = calls + 1
    seen = .true.
    blk%scale = 1.0_c_double
    blk%n = 0
    if (allocated(h%values)) then
       blk%n = int(size(h%values), c_int)
       blk%data = c_loc(h%values)
    end if
    rc = c_scale(blk, factor)
    if (.not. c_associated(blk%spare)) rc = rc + 100
    three = [1_c_int, 2_c_int, 3_c_int]
    rc = rc + 1000 * c_sum3(three)
  end function scale_through_c

  ! internal procedure with host association and assumed-length character dummies
  subroutine label_all(h, stem)
    type(holder), intent(inout) :: h
    character(len=*), intent(in) :: stem
    integer :: i
    do i = 1, size(h%tags)
       call put(i, trim(stem) // '-', merge('odd ', 'even', mod(i, 2) == 1))
    end do
  contains
    subroutine put(k, a, b)
      integer, intent(in) :: k
      character(len=*), intent(in) :: a, b
      h%tags(k) = a // b
    end subroutine put
  end subroutine label_all

  ! c_ptr -> character pointer array -> length of a C string
  function read_c_string() result(n)
    integer :: n
    character(kind=c_char), pointer :: s(:)
    type(c_ptr) :: p
    p = c_name()
    n = 0
    if (.not. c_associated(p)) return
    call c_f_pointer(p, s, [64])
    do while (n < 64)
       if (s(n + 1) == c_null_char) exit
       n = n + 1
    end do
  end function read_c_string

  integer function count_to(limit)
    integer, intent(in) :: limit
    count_to = 0
    do while (count_to < limit)
       count_to = count_to + 3; if (count_to > 100) exit
    end do
    count_to = count_to + calls
  end function count_to

end module interop_mod
