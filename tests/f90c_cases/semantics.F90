! Small, self-written Fortran used ONLY to test oracle/f90c.py: each routine isolates one rule
! of the language (or of gfortran's code generation) on which the bit-exact pinning of the
! oracle relies.  The expected values in tests/test_f90c_translator.py are derived by hand from
! the Fortran standard / gfortran documentation, not from the translator.
module kinds_mod
  implicit none
  integer, parameter :: r8 = selected_real_kind(13), i4 = selected_int_kind(6), i8 = selected_int_kind(13)
  real (r8), parameter, private :: hidden = 42.0_r8      ! must not leak through USE
  real (r8), parameter :: single_lit = 1.00e-8           ! REAL(4) literal widened (quirk Q6)
  real (r8), parameter :: double_lit = 1.00e-8_r8
  real (r8), parameter :: derived_const = 1.0_r8 / (365.0_r8 * 86400.0_r8)
  integer (r8), parameter :: int_half = 0.5_r8           ! integer(kind=8) "constant": 0 (quirk Q5)
  integer (r8), parameter :: int_ten = 10.0_r8
  real (r8) :: module_state = 0.0_r8
  type pair
     real (r8) :: a, b
     integer (i4) :: n
  end type pair
  type bag
     real (r8), allocatable, dimension(:,:) :: grid
     real (r8), allocatable, dimension(:) :: line
     integer (i4), allocatable, dimension(:) :: idx
     character (16), allocatable, dimension(:) :: names
     logical :: flag
  end type bag
end module kinds_mod

module sem_mod
  use kinds_mod
  implicit none
  private
  public :: arith, powers, minmax, sums, loops, selects, bags, strings, saved, by_ref, keyword_caller
  real (r8), parameter :: hidden = 7.0_r8                 ! own constant of the same name
  real (r8), dimension(1) :: one_elem
contains

  subroutine arith(x, y, i, j, out)
    real (r8), intent(in) :: x, y
    integer (i4), intent(in) :: i, j
    real (r8), dimension(12), intent(out) :: out
    out(1) = i / j                    ! integer division truncates toward zero
    out(2) = (-i) / j
    out(3) = x / y * y                ! left to right
    out(4) = x + y - y
    out(5) = -x ** 2                  ! -(x**2)
    out(6) = int_half * x             ! 0 * x
    out(7) = x / int_ten              ! integer(8) promoted to real
    out(8) = single_lit
    out(9) = double_lit
    out(10) = hidden                  ! the module's own, not kinds_mod's private one
    out(11) = derived_const
    out(12) = 2 * 3.0e0 * x           ! REAL(4) product widened afterwards: exact here
  end subroutine arith

  subroutine powers(x, n, out)
    real (r8), intent(in) :: x
    integer (i4), intent(in) :: n
    real (r8), dimension(8), intent(out) :: out
    out(1) = x ** 2
    out(2) = x ** 3
    out(3) = x ** n
    out(4) = x ** 1.5_r8
    out(5) = 10.0_r8 ** (-x)
    out(6) = int_ten ** (-x)          ! integer base, real exponent -> pow(10.0, -x)
    out(7) = 2 ** n                   ! integer power
    out(8) = x ** (-2)
  end subroutine powers

  subroutine minmax(x, y, out)
    real (r8), intent(in) :: x, y
    real (r8), dimension(6), intent(out) :: out
    out(1) = max(x, y)
    out(2) = min(x, y)
    out(3) = max(x, y, 0.5_r8)
    out(4) = merge(x, y, x > y)
    out(5) = merge(-1, 0, x > y)      ! integers assigned to a real
    out(6) = abs(x - y)
  end subroutine minmax

  subroutine sums(v, m, pick, out)
    real (r8), dimension(4), intent(in) :: v
    real (r8), allocatable, dimension(:,:), intent(in) :: m
    type(pair), dimension(4), intent(in) :: pick
    real (r8), dimension(6), intent(out) :: out
    real (r8), dimension(4) :: w
    w = v * 2.0_r8
    out(1) = sum(v)                               ! element order, from zero
    out(2) = sum(w)
    out(3) = sum(m(2,:), dim=1)
    out(4) = sum(m(1, pick(:)%n), dim=1)          ! vector subscript through a component
    out(5) = size(m)
    one_elem = exp(v(1))                          ! whole-array assignment of a 1-element array
    out(6) = one_elem(1)
  end subroutine sums

  subroutine loops(n, out)
    integer (i4), intent(in) :: n
    real (r8), dimension(5), intent(out) :: out
    integer (i4) :: i, j, cnt
    cnt = 0
    outer: do i = 1, n
       if (i == 2) cycle outer
       do j = 1, n
          if (j > i) exit
          if (j == 3 .and. i == 4) cycle outer
          cnt = cnt + 1
       end do
    end do outer
    out(1) = cnt
    out(2) = i                        ! n + 1 after a completed loop
    cnt = 0
    do
       cnt = cnt + 1
       if (cnt >= 7) exit
    end do
    out(3) = cnt
    cnt = 0
    do i = 10, 1, -3
       cnt = cnt + i
    end do
    out(4) = cnt                      ! 10 + 7 + 4 + 1
    i = 0
    do while (i < 5)
       i = i + 2
    end do
    out(5) = i
  end subroutine loops

  subroutine selects(k, out)
    integer (i4), intent(in) :: k
    real (r8), intent(out) :: out
    integer (i4), parameter :: first = 1, second = 2
    select case (k)
       case (first)
          out = 10.0_r8
       case (second, 5)
          out = 20.0_r8
       case (7:9)
          out = 30.0_r8
       case default
          out = -1.0_r8
    end select
  end subroutine selects

  subroutine bags(b, n, m, out)
    type(bag), intent(inout) :: b
    integer (i4), intent(in) :: n, m
    real (r8), dimension(4), intent(out) :: out
    real (r8), allocatable, dimension(:,:) :: tmp
    integer (i4) :: i, j
    allocate(tmp(n, m))
    tmp = 1.5_r8
    do j = 1, m
       do i = 1, n
          b%grid(i, j) = tmp(i, j) + i + 10 * j     ! column-major, 1-based
       end do
    end do
    b%line = 0.25_r8
    b%idx(:) = 3
    b%flag = .not. b%flag
    out(1) = b%grid(n, m)
    out(2) = size(b%grid, 1)
    out(3) = size(b%grid, 2)
    out(4) = merge(1.0_r8, 0.0_r8, allocated(tmp))
    deallocate(tmp)
  end subroutine bags

  subroutine strings(b, p)
    type(bag), intent(inout) :: b
    type(pair), intent(in) :: p
    character (16) :: stem
    stem = 'ab'
    b%names(:) = 'x'
    b%names(1) = trim(stem) // 'Chl'
    b%names(2) = stem // 'Z'                     ! no trim: the blanks stay, 'Z' falls off the end
    if (p%n > 1) b%names(3) = 'this text is longer than sixteen characters'
  end subroutine strings

  function saved(x) result(r)
    real (r8), intent(in) :: x
    real (r8) :: r
    module_state = module_state + x               ! module variable keeps its value between calls
    r = module_state
  end function saved

  subroutine callee(a, b, c, flag)
    real (r8), intent(in) :: a
    real (r8), intent(inout) :: b
    real (r8), intent(out) :: c
    logical, intent(in) :: flag
    b = b + a
    c = merge(b, -b, flag)
  end subroutine callee

  subroutine by_ref(p, out)
    type(pair), intent(inout) :: p
    real (r8), dimension(3), intent(out) :: out
    call callee(p%a * 2.0_r8, p%b, out(1), .true.)    ! expression, component, array element
    out(2) = p%b
    call callee(1.0_r8, out(2), out(3), p%n > 100)
  end subroutine by_ref

  subroutine keyword_caller(out)
    real (r8), dimension(2), intent(out) :: out
    real (r8) :: t
    t = 1.0_r8
    call callee(2.0_r8, t, out(1), flag=.false.)
    out(2) = t
  end subroutine keyword_caller

end module sem_mod
