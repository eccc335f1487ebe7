!> USER-SIDE code: a rectangular CSR operator of kind cdp (the operator of BASELINE config C5, at toy size) as an extension of the
!> reference's abstract_linop_cdp acting on its dense_vector_cdp.  rowptr / col are 0-based like include/lkb.h.  Executed together
!> with the reference's sources by oracle/f90run.py (tests/golden/make_ref_golden.py).
module user_csr
    use LightKrylov_Constants
    use LightKrylov_AbstractVectors
    use LightKrylov_AbstractLinops
    implicit none
    private
    type, extends(abstract_linop_cdp), public :: csr_linop_cdp
        integer :: m = 0, n = 0
        integer, allocatable :: rowptr(:), col(:)
        complex(dp), allocatable :: val(:)
    contains
        procedure, pass(self), public :: matvec => csr_matvec_cdp
        procedure, pass(self), public :: rmatvec => csr_rmatvec_cdp
    end type csr_linop_cdp
contains
    subroutine csr_matvec_cdp(self, vec_in, vec_out)
        class(csr_linop_cdp), intent(inout) :: self
        class(abstract_vector_cdp), intent(in) :: vec_in
        class(abstract_vector_cdp), intent(out) :: vec_out
        integer :: i, q
        complex(dp) :: s
        select type (vec_in)
        type is (dense_vector_cdp)
            select type (vec_out)
            type is (dense_vector_cdp)
                if (allocated(vec_out%data)) deallocate(vec_out%data)
                allocate(vec_out%data(self%m))
                vec_out%n = self%m
                do i = 1, self%m
                    s = zero_cdp
                    do q = self%rowptr(i) + 1, self%rowptr(i + 1)
                        s = s + self%val(q) * vec_in%data(self%col(q) + 1)
                    end do
                    vec_out%data(i) = s
                end do
            end select
        end select
    end subroutine csr_matvec_cdp

    subroutine csr_rmatvec_cdp(self, vec_in, vec_out)
        class(csr_linop_cdp), intent(inout) :: self
        class(abstract_vector_cdp), intent(in) :: vec_in
        class(abstract_vector_cdp), intent(out) :: vec_out
        integer :: i, q, j
        select type (vec_in)
        type is (dense_vector_cdp)
            select type (vec_out)
            type is (dense_vector_cdp)
                if (allocated(vec_out%data)) deallocate(vec_out%data)
                allocate(vec_out%data(self%n))
                vec_out%n = self%n
                vec_out%data = zero_cdp
                do i = 1, self%m
                    do q = self%rowptr(i) + 1, self%rowptr(i + 1)
                        j = self%col(q) + 1
                        vec_out%data(j) = vec_out%data(j) + conjg(self%val(q)) * vec_in%data(i)
                    end do
                end do
            end select
        end select
    end subroutine csr_rmatvec_cdp
end module user_csr
