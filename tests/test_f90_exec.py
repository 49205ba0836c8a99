"""The Fortran-subset interpreter (oracle/f90_exec.py, test infrastructure) on statements whose result can be checked by hand:
the typing and evaluation rules that make its output "what the reference computes" -- integer division and mod, literal kinds,
mixed-kind promotion, operator precedence and left-to-right evaluation, x**2, array bounds and sections, host association,
cpp conditionals.  tests/test_oracle_refsrc.py relies on these rules."""
import numpy as np
import pytest

from oracle.f90_exec import FortranError, Interp

SRC = """
module m
   use precision, only: wp
   implicit none
   real(wp), parameter :: third = 1.0_wp / 3.0_wp
   real(wp), parameter :: w(0:2) = [real(wp) :: 4, 1, 1] ! integer items converted to wp
   integer, parameter :: n0 = 7
contains
   subroutine ints(out)
      integer, intent(out) :: out(8)
      integer :: i
      out(1) = 7 / 2            ! 3: integer division truncates
      out(2) = (-7) / 2         ! -3: towards zero
      out(3) = mod(-7, 3)       ! -1: sign of the first argument
      out(4) = mod(7 + 5 - 2, 5) + 1
      out(5) = 2 ** 5
      out(6) = n0 / 2 * 2       ! left to right: (7/2)*2 = 6
      out(7) = 0
      do i = 1, 10, 3           ! 1, 4, 7, 10
         out(7) = out(7) + i
      end do
      out(8) = -2 ** 2          ! -(2**2) = -4
   end subroutine

   subroutine reals(out)
      real(wp), intent(out) :: out(10)
      real(wp) :: a, b
      a = 0.1_wp
      b = 0.1                   ! default-real literal: float32(0.1), then converted to wp
      out(1) = a
      out(2) = b
      out(3) = 1.0d0 / 3.0d0    ! double precision whatever wp is, converted on assignment
      out(4) = 2 * a            ! integer promoted to wp
      out(5) = a / 3 * 3        ! (a/3)*3, left to right
      out(6) = a - a * a        ! precedence
      out(7) = -a * a           ! -(a*a)
      out(8) = (a + b) ** 2     ! x*x
      out(9) = third + w(0)
      out(10) = 1 / 3           ! integer division first: 0
   end subroutine

   subroutine arrays(f, g, total)
      real(wp), intent(inout) :: f(4, 0:2)
      real(wp), intent(out) :: g(:, :)
      real(wp), intent(out) :: total
      real(wp) :: row(0:2)
      integer :: y
      row = f(2, :)             ! section -> local array with lower bound 0
      f(4, :) = row
      g(1:2, 1) = f(3:4, 0)     ! assumed shape: lower bound 1
      total = 0.0_wp
      do y = 1, size(f, 1)
         total = total + f(y, 2)
      end do
      call inner(f)
   contains
      subroutine inner(h)
         real(wp), intent(inout) :: h(4, 0:2)
         h(1, 0) = total + third   ! host association: the host's local and the module parameter
      end subroutine
   end subroutine

   subroutine switch(out)
      integer, intent(out) :: out
#if ALPHA
      out = 1
#elif BETA
      out = 2
#else
      out = 3
#endif
   end subroutine

   subroutine bad_index(f)
      real(wp), intent(inout) :: f(3)
      f(4) = 1.0_wp
   end subroutine
end module
"""
PRECISION = "module precision\n integer, parameter :: sp = kind(1.0)\nend module\n"


def make(wp, **macros):
    it = Interp(wp, macros)
    it.load_text(PRECISION)
    it.load_text(SRC)
    return it


def test_integer_semantics():
    out = np.zeros(8, dtype=np.int64)
    make("f64").run("m", "ints", out)
    assert out.tolist() == [3, -3, -1, 1, 32, 6, 22, -4]


@pytest.mark.parametrize("wp,T", [("f64", np.float64), ("f32", np.float32)])
def test_real_kinds_and_evaluation_order(wp, T):
    out = np.zeros(10, dtype=T)
    make(wp).run("m", "reals", out)
    a, b = T("0.1"), T(np.float32("0.1"))
    assert out[0] == a and out[1] == b and (wp == "f32" or a != b)
    assert out[2] == T(np.float64(1) / np.float64(3))
    assert out[3] == T(2) * a and out[4] == a / T(3) * T(3) and out[5] == a - a * a and out[6] == -(a * a)
    assert out[7] == (a + b) * (a + b)
    assert out[8] == T(1) / T(3) + T(4) and out[9] == 0


def test_arrays_sections_and_host_association():
    f = np.arange(12, dtype=np.float64).reshape(4, 3)
    g = np.zeros((2, 2))
    it = make("f64")
    total = it.run("m", "arrays", f, g, np.float64(0))  # scalars are passed by value here: check through the array
    assert total is None
    assert f[3].tolist() == [3.0, 4.0, 5.0]             # f(4,:) = f(2,:)
    assert g[:, 0].tolist() == [6.0, 3.0]               # f(3:4, 0) after the copy
    assert f[0, 0] == (2.0 + 5.0 + 8.0 + 5.0) + 1.0 / 3.0


def test_cpp_conditionals_with_elif():
    for macros, want in (({}, 3), ({"ALPHA": 1}, 1), ({"BETA": 1}, 2), ({"ALPHA": 1, "BETA": 1}, 1)):
        out = np.zeros(1, dtype=np.int64)
        it = make("f64", **macros)
        # an integer, intent(out) scalar: read it back through a one-element array by wrapping the call
        it.load_text("module w\n use m\ncontains\n subroutine go(o)\n integer, intent(out) :: o(1)\n integer :: k\n call switch(k)\n o(1) = k\n end subroutine\nend module\n")
        it.run("w", "go", out)
        assert out[0] == want, macros


def test_out_of_bounds_and_unsupported_constructs_raise():
    it = make("f64")
    with pytest.raises(FortranError):
        it.run("m", "bad_index", np.zeros(3))
    it.load_text("module u\ncontains\n subroutine s(a)\n real(8), intent(inout) :: a(2)\n a(1) = bessel_j0(a(2))\n end subroutine\nend module\n")
    with pytest.raises(FortranError):
        it.run("u", "s", np.zeros(2))
