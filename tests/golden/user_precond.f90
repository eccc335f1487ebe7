!> USER-SIDE preconditioner of the kind a LightKrylov user writes: an extension of the reference's abstract_precond_rdp whose
!> `apply` scales the vector (Jacobi for a constant-diagonal stencil).  Works on ANY abstract_vector_rdp through the type-bound
!> `scal`, so the same object serves the reference's CPU run and, through the shim's C trampoline, the device run.
module user_precond
    use LightKrylov_Constants
    use LightKrylov_AbstractVectors
    use LightKrylov_IterativeSolvers, only: abstract_precond_rdp
    implicit none
    private
    type, extends(abstract_precond_rdp), public :: jacobi_precond_rdp
        real(dp) :: inv_diag = 1.0_dp
        integer :: n_applied = 0
    contains
        procedure, pass(self), public :: apply => jacobi_apply_rdp
    end type jacobi_precond_rdp
contains
    subroutine jacobi_apply_rdp(self, vec, iter, current_residual, target_residual)
        class(jacobi_precond_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(inout) :: vec
        integer, optional, intent(in) :: iter
        real(dp), optional, intent(in) :: current_residual
        real(dp), optional, intent(in) :: target_residual
        call vec%scal(self%inv_diag)
        self%n_applied = self%n_applied + 1
    end subroutine jacobi_apply_rdp
end module user_precond
