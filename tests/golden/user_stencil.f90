!> USER-SIDE code of the kind a LightKrylov user writes (cf. example/ in the reference): constant-coefficient 5- / 7-point
!> stencil operators -- the operator family of BASELINE configs C2 / C3 / C4 -- as extensions of the reference's abstract_linop_rdp /
!> abstract_sym_linop_rdp, acting on the reference's own dense_vector_rdp.  Executed together with the reference's sources by
!> oracle/f90run.py (tests/golden/make_ref_golden.py) to produce reference-side fixtures on the north-star operator.
!> coef = (center, -x, +x, -y, +y, -z, +z), x fastest, homogeneous Dirichlet boundaries (same convention as include/lkb.h).
module user_stencil
    use LightKrylov_Constants
    use LightKrylov_AbstractVectors
    use LightKrylov_AbstractLinops
    implicit none
    private
    public :: stencil_linop_rdp, stencil_sym_linop_rdp, stencil_apply_rdp

    type, extends(abstract_linop_rdp), public :: stencil_linop_rdp
        integer :: nx = 1, ny = 1, nz = 1
        real(dp) :: coef(7) = 0.0_dp
    contains
        procedure, pass(self), public :: matvec => stencil_matvec_rdp
        procedure, pass(self), public :: rmatvec => stencil_rmatvec_rdp
    end type stencil_linop_rdp

    type, extends(abstract_sym_linop_rdp), public :: stencil_sym_linop_rdp
        integer :: nx = 1, ny = 1, nz = 1
        real(dp) :: coef(7) = 0.0_dp
    contains
        procedure, pass(self), public :: matvec => stencil_sym_matvec_rdp
    end type stencil_sym_linop_rdp

contains

    subroutine stencil_apply_rdp(nx, ny, nz, c, x, y)
        integer, intent(in) :: nx, ny, nz
        real(dp), intent(in) :: c(7)
        real(dp), intent(in) :: x(:)
        real(dp), intent(inout) :: y(:)
        integer :: j, k, b
        y = c(1) * x
        do k = 0, nz - 1
            do j = 0, ny - 1
                b = nx * (j + ny * k)
                y(b+2:b+nx) = y(b+2:b+nx) + c(2) * x(b+1:b+nx-1)
                y(b+1:b+nx-1) = y(b+1:b+nx-1) + c(3) * x(b+2:b+nx)
                if (j > 0) y(b+1:b+nx) = y(b+1:b+nx) + c(4) * x(b+1-nx:b)
                if (j < ny - 1) y(b+1:b+nx) = y(b+1:b+nx) + c(5) * x(b+1+nx:b+2*nx)
                if (k > 0) y(b+1:b+nx) = y(b+1:b+nx) + c(6) * x(b+1-nx*ny:b+nx-nx*ny)
                if (k < nz - 1) y(b+1:b+nx) = y(b+1:b+nx) + c(7) * x(b+1+nx*ny:b+nx+nx*ny)
            end do
        end do
    end subroutine stencil_apply_rdp

    subroutine stencil_matvec_rdp(self, vec_in, vec_out)
        class(stencil_linop_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(in) :: vec_in
        class(abstract_vector_rdp), intent(out) :: vec_out
        select type (vec_in)
        type is (dense_vector_rdp)
            select type (vec_out)
            type is (dense_vector_rdp)
                vec_out = vec_in
                call stencil_apply_rdp(self%nx, self%ny, self%nz, self%coef, vec_in%data, vec_out%data)
            end select
        end select
    end subroutine stencil_matvec_rdp

    subroutine stencil_rmatvec_rdp(self, vec_in, vec_out)
        class(stencil_linop_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(in) :: vec_in
        class(abstract_vector_rdp), intent(out) :: vec_out
        real(dp) :: ct(7)
        ct = self%coef([1, 3, 2, 5, 4, 7, 6])          ! the transpose swaps the -/+ neighbours
        select type (vec_in)
        type is (dense_vector_rdp)
            select type (vec_out)
            type is (dense_vector_rdp)
                vec_out = vec_in
                call stencil_apply_rdp(self%nx, self%ny, self%nz, ct, vec_in%data, vec_out%data)
            end select
        end select
    end subroutine stencil_rmatvec_rdp

    subroutine stencil_sym_matvec_rdp(self, vec_in, vec_out)
        class(stencil_sym_linop_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(in) :: vec_in
        class(abstract_vector_rdp), intent(out) :: vec_out
        select type (vec_in)
        type is (dense_vector_rdp)
            select type (vec_out)
            type is (dense_vector_rdp)
                vec_out = vec_in
                call stencil_apply_rdp(self%nx, self%ny, self%nz, self%coef, vec_in%data, vec_out%data)
            end select
        end select
    end subroutine stencil_sym_matvec_rdp
end module user_stencil
