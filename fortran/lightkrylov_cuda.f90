!> lightkrylov_cuda.f90 -- ISO_C_BINDING shim: device-resident extensions of LightKrylov's abstract
!> types on top of liblkb.so (include/lkb.h).  SHIPPED AS SOURCE: this image has no Fortran
!> compiler, so the shim is validated by inspection against the interfaces it extends:
!>   abstract_vector_rdp   src/AbstractTypes/AbstractVectors.fypp:295-381
!>   abstract_linop_rdp    src/AbstractTypes/AbstractLinops.fypp:58-87
!>   arnoldi / lanczos     src/Krylov/BaseKrylov.fypp:132-152, 221-234
!> Only the rdp instance is spelled out; rsp/csp/cdp differ by the kind constant (LKB_S/C/Z) and
!> the scalar type, exactly like the fypp expansion of the reference.
!>
!> Object semantics (SURVEY.md section 7, hard parts):
!>  * `assignment(=)` is a deep copy (lkb_vec_clone): `wrk = V(k)`, `allocate(.., source=b)`.
!>  * no `final` procedure on the vector type: intent(out) dummies (matvec's vec_out, copy's out)
!>    would otherwise free live basis columns.  Owning vectors are released with `%destroy()`.
!>  * `zero`/`axpby` allocate lazily when the handle is null, like dense_vector does
!>    (AbstractVectors.fypp:481-486, 522-525); axpby with beta == 0 overwrites without reading self.
!>  * bases handed to arnoldi/lanczos come from `cuda_basis_rdp` (one contiguous column-major
!>    device array); `X(i)` views are non-owning.
module lightkrylov_cuda
    use, intrinsic :: iso_c_binding
    use LightKrylov_Constants, only: dp
    use LightKrylov_AbstractVectors, only: abstract_vector_rdp
    use LightKrylov_AbstractLinops, only: abstract_linop_rdp
    use LightKrylov_Logger, only: stop_error
    implicit none(type, external)
    private
    character(len=*), parameter :: this_module = 'lightkrylov_cuda'
    integer(c_int), parameter :: LKB_D = 1

    type(c_ptr), save, public :: lkb_ctx = c_null_ptr     !! one context per process per GPU

    interface
        integer(c_int) function lkb_init(device, ctx) bind(C, name='lkb_init')
            import; integer(c_int), value :: device; type(c_ptr), intent(out) :: ctx
        end function
        integer(c_int) function lkb_init_dist(device, rank, world, id128, ctx) bind(C, name='lkb_init_dist')
            import; integer(c_int), value :: device, rank, world; type(c_ptr), value :: id128; type(c_ptr), intent(out) :: ctx
        end function
        integer(c_int) function lkb_nccl_unique_id(id128) bind(C, name='lkb_nccl_unique_id')
            import; type(c_ptr), value :: id128
        end function
        integer(c_int) function lkb_vec_create(ctx, kind, n_local, n_global, row0, v) bind(C, name='lkb_vec_create')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: kind
            integer(c_int64_t), value :: n_local, n_global, row0; type(c_ptr), intent(out) :: v
        end function
        integer(c_int) function lkb_vec_clone(src, dst) bind(C, name='lkb_vec_clone')
            import; type(c_ptr), value :: src; type(c_ptr), intent(out) :: dst
        end function
        integer(c_int) function lkb_vec_destroy(v) bind(C, name='lkb_vec_destroy')
            import; type(c_ptr), value :: v
        end function
        integer(c_int) function lkb_vec_zero(v) bind(C, name='lkb_vec_zero')
            import; type(c_ptr), value :: v
        end function
        integer(c_int) function lkb_vec_rand(v, ifnorm) bind(C, name='lkb_vec_rand')
            import; type(c_ptr), value :: v; integer(c_int32_t), value :: ifnorm
        end function
        integer(c_int) function lkb_vec_scal(v, alpha) bind(C, name='lkb_vec_scal')
            import; type(c_ptr), value :: v; real(c_double), intent(in) :: alpha
        end function
        integer(c_int) function lkb_vec_axpby(alpha, x, beta, self) bind(C, name='lkb_vec_axpby')
            import; real(c_double), intent(in) :: alpha, beta; type(c_ptr), value :: x, self
        end function
        integer(c_int) function lkb_vec_dot(self, vec, res) bind(C, name='lkb_vec_dot')
            import; type(c_ptr), value :: self, vec; real(c_double), intent(out) :: res
        end function
        integer(c_int64_t) function lkb_vec_size(v) bind(C, name='lkb_vec_size')
            import; type(c_ptr), value :: v
        end function
        integer(c_int) function lkb_basis_create(ctx, kind, n_local, n_global, row0, ncols, b) bind(C, name='lkb_basis_create')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: kind, ncols
            integer(c_int64_t), value :: n_local, n_global, row0; type(c_ptr), intent(out) :: b
        end function
        integer(c_int) function lkb_basis_col(b, i0, view) bind(C, name='lkb_basis_col')
            import; type(c_ptr), value :: b; integer(c_int), value :: i0; type(c_ptr), intent(out) :: view
        end function
        integer(c_int) function lkb_op_stencil5_create(ctx, kind, nx, ny, coef5, slow0, nslow, A) bind(C, name='lkb_op_stencil5_create')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: kind
            integer(c_int64_t), value :: nx, ny, slow0, nslow; real(c_double), intent(in) :: coef5(5); type(c_ptr), intent(out) :: A
        end function
        integer(c_int) function lkb_op_matvec(A, x, y) bind(C, name='lkb_op_matvec')
            import; type(c_ptr), value :: A, x, y
        end function
        integer(c_int) function lkb_op_rmatvec(A, x, y) bind(C, name='lkb_op_rmatvec')
            import; type(c_ptr), value :: A, x, y
        end function
        integer(c_int) function lkb_arnoldi(A, X, H, ldh, info, kstart, kend, tol, transpose, blksize) bind(C, name='lkb_arnoldi')
            import; type(c_ptr), value :: A, X; real(c_double), intent(inout) :: H(ldh, *)
            integer(c_int), value :: ldh; integer(c_int32_t), intent(out) :: info
            integer(c_int32_t), value :: kstart, kend, transpose, blksize; real(c_double), value :: tol
        end function
        integer(c_int) function lkb_bidiag(A, U, V, B, ldb, info, kstart, kend, tol) bind(C, name='lkb_bidiag')
            import; type(c_ptr), value :: A, U, V; real(c_double), intent(inout) :: B(ldb, *)
            integer(c_int), value :: ldb; integer(c_int32_t), intent(out) :: info
            integer(c_int32_t), value :: kstart, kend; real(c_double), value :: tol
        end function
        integer(c_int) function lkb_dgs_step(X, j, W, wcol0, p, chk, beta, ldbeta, info) bind(C, name='lkb_dgs_step')
            import; type(c_ptr), value :: X, W, beta; integer(c_int), value :: j, wcol0, p, ldbeta
            integer(c_int32_t), value :: chk; integer(c_int32_t), intent(out) :: info
        end function
        integer(c_int) function lkb_gmres(A, b, x, info, rtol, atol, transpose, io) bind(C, name='lkb_gmres')
            import; type(c_ptr), value :: A, b, x, io; integer(c_int32_t), intent(out) :: info
            real(c_double), value :: rtol, atol; integer(c_int32_t), value :: transpose
        end function
        integer(c_int) function lkb_cg(A, b, x, info, rtol, atol, io) bind(C, name='lkb_cg')
            import; type(c_ptr), value :: A, b, x, io; integer(c_int32_t), intent(out) :: info
            real(c_double), value :: rtol, atol
        end function
        integer(c_int) function lkb_lanczos(A, X, T, ldt, info, kstart, kend, tol) bind(C, name='lkb_lanczos')
            import; type(c_ptr), value :: A, X; real(c_double), intent(inout) :: T(ldt, *)
            integer(c_int), value :: ldt; integer(c_int32_t), intent(out) :: info
            integer(c_int32_t), value :: kstart, kend; real(c_double), value :: tol
        end function
    end interface

    !> Device-resident vector: drop-in extension of abstract_vector_rdp.
    type, extends(abstract_vector_rdp), public :: cuda_vector_rdp
        type(c_ptr) :: h = c_null_ptr            !! lkb_vec_t
        integer(c_int64_t) :: n_local = 0, n_global = 0, row0 = 0
        logical :: owns = .false.
    contains
        procedure, pass(self), public :: zero => cuda_zero_rdp
        procedure, pass(self), public :: rand => cuda_rand_rdp
        procedure, pass(self), public :: scal => cuda_scal_rdp
        procedure, pass(self), public :: axpby => cuda_axpby_rdp
        procedure, pass(self), public :: dot => cuda_dot_rdp
        procedure, pass(self), public :: get_size => cuda_get_size_rdp
        procedure, pass(self), public :: destroy => cuda_destroy_rdp
        procedure, pass(lhs), private :: cuda_assign_rdp
        generic, public :: assignment(=) => cuda_assign_rdp
    end type

    !> Contiguous column-major device basis; X(i) are non-owning cuda_vector_rdp views.
    type, public :: cuda_basis_rdp
        type(c_ptr) :: h = c_null_ptr            !! lkb_basis_t
        type(cuda_vector_rdp), allocatable :: X(:)
    end type

    !> 5-point stencil operator: drop-in extension of abstract_linop_rdp.
    type, extends(abstract_linop_rdp), public :: cuda_stencil5_rdp
        type(c_ptr) :: h = c_null_ptr            !! lkb_op_t
    contains
        procedure, pass(self), public :: matvec => stencil_matvec_rdp
        procedure, pass(self), public :: rmatvec => stencil_rmatvec_rdp
    end type

    !> C mirrors of gmres_dp_opts + gmres_dp_metadata / cg_dp_opts + cg_dp_metadata (lkb_gmres_io, lkb_cg_io)
    type, bind(C), public :: lkb_gmres_io
        integer(c_int32_t) :: kdim = 30, maxiter = 10
        integer(c_int32_t) :: n_iter = 0, n_inner = 0, n_outer = 0, converged = 0, info = 0
        type(c_ptr) :: res = c_null_ptr
        integer(c_int32_t) :: res_cap = 0, res_len = 0
    end type
    type, bind(C), public :: lkb_cg_io
        integer(c_int32_t) :: maxiter = 100
        integer(c_int32_t) :: n_iter = 0, converged = 0, info = 0
        type(c_ptr) :: res = c_null_ptr
        integer(c_int32_t) :: res_cap = 0, res_len = 0
    end type

    public :: cuda_basis_create_rdp, arnoldi_cuda_rdp, lanczos_cuda_rdp, bidiagonalization_cuda_rdp
    public :: dgs_cuda_rdp, gmres_cuda_rdp, cg_cuda_rdp

contains

    subroutine chk(rc, what)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: what
        if (rc /= 0) call stop_error('liblkb call failed: '//what, this_module, what)
    end subroutine

    subroutine ensure(self, like)
        class(cuda_vector_rdp), intent(inout) :: self
        class(cuda_vector_rdp), intent(in), optional :: like
        if (c_associated(self%h)) return
        if (present(like)) then
            self%n_local = like%n_local; self%n_global = like%n_global; self%row0 = like%row0
        end if
        call chk(lkb_vec_create(lkb_ctx, LKB_D, self%n_local, self%n_global, self%row0, self%h), 'lkb_vec_create')
        self%owns = .true.
    end subroutine

    subroutine cuda_zero_rdp(self)
        class(cuda_vector_rdp), intent(inout) :: self
        call ensure(self)
        call chk(lkb_vec_zero(self%h), 'lkb_vec_zero')
    end subroutine

    subroutine cuda_rand_rdp(self, ifnorm)
        class(cuda_vector_rdp), intent(inout) :: self
        logical, optional, intent(in) :: ifnorm
        integer(c_int32_t) :: flag
        flag = 0; if (present(ifnorm)) flag = merge(1, 0, ifnorm)
        call ensure(self)
        call chk(lkb_vec_rand(self%h, flag), 'lkb_vec_rand')
    end subroutine

    subroutine cuda_scal_rdp(self, alpha)
        class(cuda_vector_rdp), intent(inout) :: self
        real(dp), intent(in) :: alpha
        call chk(lkb_vec_scal(self%h, alpha), 'lkb_vec_scal')
    end subroutine

    subroutine cuda_axpby_rdp(alpha, vec, beta, self)
        class(cuda_vector_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(in) :: vec
        real(dp), intent(in) :: alpha, beta
        select type (vec)
        type is (cuda_vector_rdp)
            call ensure(self, vec)
            call chk(lkb_vec_axpby(alpha, vec%h, beta, self%h), 'lkb_vec_axpby')
        class default
            call stop_error('axpby: vec must be a cuda_vector_rdp', this_module, 'cuda_axpby_rdp')
        end select
    end subroutine

    real(dp) function cuda_dot_rdp(self, vec) result(alpha)
        class(cuda_vector_rdp), intent(in) :: self
        class(abstract_vector_rdp), intent(in) :: vec
        alpha = 0.0_dp
        select type (vec)
        type is (cuda_vector_rdp)
            call chk(lkb_vec_dot(self%h, vec%h, alpha), 'lkb_vec_dot')   ! globally reduced on every rank
        class default
            call stop_error('dot: vec must be a cuda_vector_rdp', this_module, 'cuda_dot_rdp')
        end select
    end function

    integer function cuda_get_size_rdp(self) result(n)
        class(cuda_vector_rdp), intent(in) :: self
        n = int(lkb_vec_size(self%h))
    end function

    subroutine cuda_destroy_rdp(self)
        class(cuda_vector_rdp), intent(inout) :: self
        if (self%owns .and. c_associated(self%h)) call chk(lkb_vec_destroy(self%h), 'lkb_vec_destroy')
        self%h = c_null_ptr; self%owns = .false.
    end subroutine

    subroutine cuda_assign_rdp(lhs, rhs)        ! deep copy: wrk = V(k), allocate(source=)
        class(cuda_vector_rdp), intent(inout) :: lhs
        type(cuda_vector_rdp), intent(in) :: rhs
        call ensure(lhs, rhs)
        call chk(lkb_vec_axpby(1.0_dp, rhs%h, 0.0_dp, lhs%h), 'lkb_vec_axpby(copy)')
    end subroutine

    subroutine cuda_basis_create_rdp(B, n_local, n_global, row0, ncols)
        type(cuda_basis_rdp), intent(out) :: B
        integer(c_int64_t), intent(in) :: n_local, n_global, row0
        integer, intent(in) :: ncols
        integer :: i
        call chk(lkb_basis_create(lkb_ctx, LKB_D, n_local, n_global, row0, int(ncols, c_int), B%h), 'lkb_basis_create')
        allocate(B%X(ncols))
        do i = 1, ncols
            call chk(lkb_basis_col(B%h, int(i - 1, c_int), B%X(i)%h), 'lkb_basis_col')
            B%X(i)%n_local = n_local; B%X(i)%n_global = n_global; B%X(i)%row0 = row0; B%X(i)%owns = .false.
        end do
    end subroutine

    subroutine stencil_matvec_rdp(self, vec_in, vec_out)
        class(cuda_stencil5_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(in) :: vec_in
        class(abstract_vector_rdp), intent(out) :: vec_out
        select type (vec_in); type is (cuda_vector_rdp)
        select type (vec_out); type is (cuda_vector_rdp)
            call ensure(vec_out, vec_in)
            call chk(lkb_op_matvec(self%h, vec_in%h, vec_out%h), 'lkb_op_matvec')
        end select; end select
    end subroutine

    subroutine stencil_rmatvec_rdp(self, vec_in, vec_out)
        class(cuda_stencil5_rdp), intent(inout) :: self
        class(abstract_vector_rdp), intent(in) :: vec_in
        class(abstract_vector_rdp), intent(out) :: vec_out
        select type (vec_in); type is (cuda_vector_rdp)
        select type (vec_out); type is (cuda_vector_rdp)
            call ensure(vec_out, vec_in)
            call chk(lkb_op_rmatvec(self%h, vec_in%h, vec_out%h), 'lkb_op_rmatvec')
        end select; end select
    end subroutine

    !> Same signature and semantics as arnoldi_rdp (BaseKrylov.fypp:132-152); the whole kstart..kend
    !> loop runs as one CUDA graph inside liblkb.  Register it in the generic with
    !>   interface arnoldi; module procedure arnoldi_cuda_rdp; end interface
    subroutine arnoldi_cuda_rdp(A, X, H, info, kstart, kend, tol, transpose, blksize)
        class(cuda_stencil5_rdp), intent(inout) :: A
        type(cuda_basis_rdp), intent(inout) :: X
        real(dp), intent(inout) :: H(:, :)
        integer, intent(out) :: info
        integer, optional, intent(in) :: kstart, kend, blksize
        real(dp), optional, intent(in) :: tol
        logical, optional, intent(in) :: transpose
        integer(c_int32_t) :: ks, ke, tr, bs, cinfo
        real(c_double) :: ctol
        real(dp), allocatable :: Hc(:, :)
        ks = 0; ke = 0; tr = 0; bs = 1; ctol = -1.0_c_double      ! 0 / <0 = "absent" (optval defaults)
        if (present(kstart)) ks = kstart
        if (present(kend)) ke = kend
        if (present(tol)) ctol = tol
        if (present(transpose)) tr = merge(1, 0, transpose)
        if (present(blksize)) bs = blksize
        Hc = H                                   ! contiguous copy: H may be a non-contiguous section
        call chk(lkb_arnoldi(A%h, X%h, Hc, int(size(Hc, 1), c_int), cinfo, ks, ke, ctol, tr, bs), 'lkb_arnoldi')
        H = Hc; info = cinfo
    end subroutine

    subroutine lanczos_cuda_rdp(A, X, T, info, kstart, kend, tol)
        class(cuda_stencil5_rdp), intent(inout) :: A
        type(cuda_basis_rdp), intent(inout) :: X
        real(dp), intent(inout) :: T(:, :)
        integer, intent(out) :: info
        integer, optional, intent(in) :: kstart, kend
        real(dp), optional, intent(in) :: tol
        integer(c_int32_t) :: ks, ke, cinfo
        real(c_double) :: ctol
        real(dp), allocatable :: Tc(:, :)
        ks = 0; ke = 0; ctol = -1.0_c_double
        if (present(kstart)) ks = kstart
        if (present(kend)) ke = kend
        if (present(tol)) ctol = tol
        Tc = T
        call chk(lkb_lanczos(A%h, X%h, Tc, int(size(Tc, 1), c_int), cinfo, ks, ke, ctol), 'lkb_lanczos')
        T = Tc; info = cinfo
    end subroutine

    !> bidiagonalization(A, U, V, B, info, kstart, kend, tol)   BaseKrylov.fypp:311-330
    subroutine bidiagonalization_cuda_rdp(A, U, V, B, info, kstart, kend, tol)
        class(cuda_stencil5_rdp), intent(inout) :: A
        type(cuda_basis_rdp), intent(inout) :: U, V
        real(dp), intent(inout) :: B(:, :)
        integer, intent(out) :: info
        integer, optional, intent(in) :: kstart, kend
        real(dp), optional, intent(in) :: tol
        integer(c_int32_t) :: ks, ke, cinfo
        real(c_double) :: ctol
        real(dp), allocatable :: Bc(:, :)
        ks = 0; ke = 0; ctol = -1.0_c_double
        if (present(kstart)) ks = kstart
        if (present(kend)) ke = kend
        if (present(tol)) ctol = tol
        Bc = B
        call chk(lkb_bidiag(A%h, U%h, V%h, Bc, int(size(Bc, 1), c_int), cinfo, ks, ke, ctol), 'lkb_bidiag')
        B = Bc; info = cinfo
    end subroutine

    !> double_gram_schmidt_step(y, X(:j), info, if_chk_orthonormal, beta) with y = W%X(iw)   BaseKrylov.fypp:679-709
    subroutine dgs_cuda_rdp(W, iw, X, j, info, if_chk_orthonormal, beta)
        type(cuda_basis_rdp), intent(inout) :: W
        type(cuda_basis_rdp), intent(in) :: X
        integer, intent(in) :: iw, j
        integer, intent(out) :: info
        logical, optional, intent(in) :: if_chk_orthonormal
        real(dp), optional, target, intent(out) :: beta(:)
        integer(c_int32_t) :: cinfo, chkflag
        type(c_ptr) :: pbeta
        chkflag = 1; if (present(if_chk_orthonormal)) chkflag = merge(1, 0, if_chk_orthonormal)   ! default .true.
        pbeta = c_null_ptr
        if (present(beta)) then
            if (size(beta) /= j) call stop_error('beta has the wrong shape', this_module, 'dgs_cuda_rdp')  ! assert_shape
            pbeta = c_loc(beta)
        end if
        call chk(lkb_dgs_step(X%h, int(j, c_int), W%h, int(iw - 1, c_int), 1_c_int, chkflag, pbeta, int(max(j, 1), c_int), cinfo), &
                 'lkb_dgs_step')
        info = cinfo
    end subroutine

    !> gmres(A, b, x, info, rtol, atol, options) -- options%kdim / %maxiter as in gmres_dp_opts
    subroutine gmres_cuda_rdp(A, b, x, info, rtol, atol, kdim, maxiter, transpose, n_iter, converged)
        class(cuda_stencil5_rdp), intent(inout) :: A
        type(cuda_vector_rdp), intent(in) :: b
        type(cuda_vector_rdp), intent(inout) :: x
        integer, intent(out) :: info
        real(dp), optional, intent(in) :: rtol, atol
        integer, optional, intent(in) :: kdim, maxiter
        logical, optional, intent(in) :: transpose
        integer, optional, intent(out) :: n_iter
        logical, optional, intent(out) :: converged
        type(lkb_gmres_io), target :: io
        real(c_double) :: r, a
        integer(c_int32_t) :: tr, cinfo
        r = -1.0_c_double; a = -1.0_c_double; tr = 0
        if (present(rtol)) r = rtol
        if (present(atol)) a = atol
        if (present(kdim)) io%kdim = kdim
        if (present(maxiter)) io%maxiter = maxiter
        if (present(transpose)) tr = merge(1, 0, transpose)
        call chk(lkb_gmres(A%h, b%h, x%h, cinfo, r, a, tr, c_loc(io)), 'lkb_gmres')
        info = cinfo                                  ! +n_iter converged / -n_iter not (gmres.fypp:234-238)
        if (present(n_iter)) n_iter = io%n_iter
        if (present(converged)) converged = io%converged /= 0
    end subroutine

    !> cg(A, b, x, info, rtol, atol, options)
    subroutine cg_cuda_rdp(A, b, x, info, rtol, atol, maxiter)
        class(cuda_stencil5_rdp), intent(inout) :: A
        type(cuda_vector_rdp), intent(in) :: b
        type(cuda_vector_rdp), intent(inout) :: x
        integer, intent(out) :: info
        real(dp), optional, intent(in) :: rtol, atol
        integer, optional, intent(in) :: maxiter
        type(lkb_cg_io), target :: io
        real(c_double) :: r, a
        integer(c_int32_t) :: cinfo
        r = -1.0_c_double; a = -1.0_c_double
        if (present(rtol)) r = rtol
        if (present(atol)) a = atol
        if (present(maxiter)) io%maxiter = maxiter
        call chk(lkb_cg(A%h, b%h, x%h, cinfo, r, a, c_loc(io)), 'lkb_cg')
        info = cinfo
    end subroutine

end module lightkrylov_cuda
