!> nonlin_batch.f90 -- iso_c_binding layer over include/nonlin_batch.h (libnonlin_b200.so).
!!
!! The batch extension keeps nonlin's API surface: the same type names and setters as
!! nonlin_multi_eqn_mult_var / nonlin_least_squares / nonlin_solve, with a `solve_batch`
!! that forwards to hand-written sm_100a CUDA.  Host code stays Fortran; this module contains
!! no numerics.  x is declared x(B, n): column-major storage makes the system index fastest,
!! which is exactly the SoA layout the kernels read.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler.  The
!! bind(C) interfaces below are a 1:1 transcription of include/nonlin_batch.h; the C ABI itself
!! is exercised by tests/ through ctypes.  Build: gfortran -c nonlin_batch.f90 ; link -lnonlin_b200.
module nonlin_batch
    use, intrinsic :: iso_c_binding
    use, intrinsic :: iso_fortran_env, only : int32, int64, real64
    implicit none
    private
    public :: nlb_engine, batch_vecfcn_helper, batch_iteration_behavior
    public :: batch_least_squares_solver, batch_newton_solver, batch_quasi_newton_solver, batch_line_search
    public :: batch_constrained_least_squares_solver, batch_polynomial, batch_solver_1var
    public :: NLB_OK, NL_NO_ERROR, NL_CONVERGENCE_ERROR, NL_DIVERGENT_BEHAVIOR_ERROR, &
        NL_SPURIOUS_CONVERGENCE_ERROR

    integer(c_int), parameter :: NLB_OK = 0
    integer(c_int), parameter :: NL_NO_ERROR = 0
    integer(c_int), parameter :: NL_CONVERGENCE_ERROR = 106
    integer(c_int), parameter :: NL_DIVERGENT_BEHAVIOR_ERROR = 206
    integer(c_int), parameter :: NL_SPURIOUS_CONVERGENCE_ERROR = 207

    !> struct nlb_params
    type, bind(C) :: nlb_params
        integer(c_int32_t) :: max_fcn_evals
        real(c_double) :: fcn_tol, var_tol, grad_tol, lm_factor
        integer(c_int32_t) :: jacobian_interval, use_line_search, ls_max_fcn_evals
        real(c_double) :: ls_alpha, ls_factor
        integer(c_int32_t) :: use_analytic_jacobian, max_iter_guard
    end type

    !> struct nlb_params_1var
    type, bind(C) :: nlb_params_1var
        integer(c_int32_t) :: max_fcn_evals
        real(c_double) :: fcn_tol, var_tol, diff_tol
        integer(c_int32_t) :: use_analytic_diff
    end type

    !> struct nlb_constrained_options (lower / upper: c_loc of n host doubles, or c_null_ptr)
    type, bind(C) :: nlb_constrained_options
        real(c_double) :: trust_region_radius, step_scaling_factor
        type(c_ptr) :: lower, upper
    end type

    !> struct nlb_iteration_behavior == iteration_behavior (nonlin_types.f90:8-29) with C ints for the logicals
    type, bind(C) :: batch_iteration_behavior
        integer(c_int32_t) :: iter_count, fcn_count, jacobian_count, gradient_count
        integer(c_int32_t) :: converge_on_fcn, converge_on_chng, converge_on_zero_diff
    end type

    interface
        integer(c_int) function nlb_create(handle, device) bind(C, name = "nlb_create")
            import :: c_ptr, c_int
            type(c_ptr), intent(out) :: handle
            integer(c_int), value :: device
        end function
        integer(c_int) function nlb_destroy(handle) bind(C, name = "nlb_destroy")
            import :: c_ptr, c_int
            type(c_ptr), value :: handle
        end function
        subroutine nlb_params_default(p) bind(C, name = "nlb_params_default")
            import :: nlb_params
            type(nlb_params), intent(out) :: p
        end subroutine
        integer(c_int) function nlb_vecfcn_lookup(name) bind(C, name = "nlb_vecfcn_lookup")
            import :: c_char, c_int
            character(kind = c_char), dimension(*), intent(in) :: name
        end function
        integer(c_int) function nlb_least_squares_solve_batch(handle, params, fcn_id, b, m, n, x, fvec, sys, &
                shared, ib, status, stream) bind(C, name = "nlb_least_squares_solve_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params
            type(c_ptr), value :: handle
            type(nlb_params), intent(in) :: params
            integer(c_int), value :: fcn_id, m, n
            integer(c_int64_t), value :: b
            type(c_ptr), value :: x, fvec, sys, shared, ib, status, stream
        end function
        integer(c_int) function nlb_constrained_least_squares_solve_batch(handle, params, options, fcn_id, b, m, n, &
                x, fvec, sys, shared, ib, status, stream) bind(C, name = "nlb_constrained_least_squares_solve_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params, nlb_constrained_options
            type(c_ptr), value :: handle
            type(nlb_params), intent(in) :: params
            type(nlb_constrained_options), intent(in) :: options
            integer(c_int), value :: fcn_id, m, n
            integer(c_int64_t), value :: b
            type(c_ptr), value :: x, fvec, sys, shared, ib, status, stream
        end function
        integer(c_int) function nlb_polynomial_fit_batch(handle, b, npts, order, thru_zero, x_is_shared, x, y, &
                coeffs, status, stream) bind(C, name = "nlb_polynomial_fit_batch")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: b
            integer(c_int), value :: npts, order, thru_zero, x_is_shared
            type(c_ptr), value :: x, y, coeffs, status, stream
        end function
        integer(c_int) function nlb_polynomial_evaluate_batch(handle, b, order, npts, x_is_shared, coeffs, x, y, &
                stream) bind(C, name = "nlb_polynomial_evaluate_batch")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: b
            integer(c_int), value :: order, npts, x_is_shared
            type(c_ptr), value :: coeffs, x, y, stream
        end function
        subroutine nlb_params_1var_default(p) bind(C, name = "nlb_params_1var_default")
            import :: nlb_params_1var
            type(nlb_params_1var), intent(out) :: p
        end subroutine
        integer(c_int) function nlb_fcn1var_lookup(name) bind(C, name = "nlb_fcn1var_lookup")
            import :: c_char, c_int
            character(kind = c_char), dimension(*), intent(in) :: name
        end function
        integer(c_int) function nlb_brent_solve_batch(handle, params, fcn_id, b, lim1, lim2, x, f, args, ib, &
                status, stream) bind(C, name = "nlb_brent_solve_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params_1var
            type(c_ptr), value :: handle
            type(nlb_params_1var), intent(in) :: params
            integer(c_int), value :: fcn_id
            integer(c_int64_t), value :: b
            type(c_ptr), value :: lim1, lim2, x, f, args, ib, status, stream
        end function
        integer(c_int) function nlb_newton_1var_solve_batch(handle, params, fcn_id, b, lim1, lim2, x, f, args, ib, &
                status, stream) bind(C, name = "nlb_newton_1var_solve_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params_1var
            type(c_ptr), value :: handle
            type(nlb_params_1var), intent(in) :: params
            integer(c_int), value :: fcn_id
            integer(c_int64_t), value :: b
            type(c_ptr), value :: lim1, lim2, x, f, args, ib, status, stream
        end function
        integer(c_int) function nlb_newton_solve_batch(handle, params, fcn_id, b, m, n, x, fvec, sys, &
                shared, ib, status, stream) bind(C, name = "nlb_newton_solve_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params
            type(c_ptr), value :: handle
            type(nlb_params), intent(in) :: params
            integer(c_int), value :: fcn_id, m, n
            integer(c_int64_t), value :: b
            type(c_ptr), value :: x, fvec, sys, shared, ib, status, stream
        end function
        integer(c_int) function nlb_quasi_newton_solve_batch(handle, params, fcn_id, b, m, n, x, fvec, sys, &
                shared, ib, status, stream) bind(C, name = "nlb_quasi_newton_solve_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params
            type(c_ptr), value :: handle
            type(nlb_params), intent(in) :: params
            integer(c_int), value :: fcn_id, m, n
            integer(c_int64_t), value :: b
            type(c_ptr), value :: x, fvec, sys, shared, ib, status, stream
        end function
        integer(c_int) function nlb_jacobian_batch(handle, params, fcn_id, b, m, n, x, jac, sys, shared, stream) &
                bind(C, name = "nlb_jacobian_batch")
            import :: c_ptr, c_int, c_int64_t, nlb_params
            type(c_ptr), value :: handle
            type(nlb_params), intent(in) :: params
            integer(c_int), value :: fcn_id, m, n
            integer(c_int64_t), value :: b
            type(c_ptr), value :: x, jac, sys, shared, stream
        end function
        integer(c_int) function nlb_reduce_stats(handle, b, ib, status, stats, stream) bind(C, name = "nlb_reduce_stats")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: b
            type(c_ptr), value :: ib, status, stats, stream
        end function
        !> One batch over several GPUs of this process (include/nonlin_batch.h: nlb_solve_sharded).  handles = array of
        !> engine handles (one per device); solver = 0 least squares, 1 Newton, 2 quasi-Newton; all data in host memory.
        integer(c_int) function nlb_solve_sharded(handles, ndev, solver, params, fcn_id, b, m, n, x, fvec, sys, shared, &
                ib, status, stats) bind(C, name = "nlb_solve_sharded")
            import :: c_ptr, c_int, c_int64_t, nlb_params
            type(c_ptr), intent(in) :: handles(*)
            integer(c_int), value :: ndev, solver, fcn_id, m, n
            type(nlb_params), intent(in) :: params
            integer(c_int64_t), value :: b
            type(c_ptr), value :: x, fvec, sys, shared, ib, status, stats
        end function
        !> Residual plug-ins (nonlin_b200/csrc/nlb_plugin.cuh): load a library of residuals compiled outside the engine.
        integer(c_int) function nlb_load_plugin(path) bind(C, name = "nlb_load_plugin")
            import :: c_int, c_char
            character(kind = c_char), intent(in) :: path(*)
        end function
    end interface

    !> One engine handle per GPU (stream + staging workspace).
    type :: nlb_engine
        type(c_ptr) :: handle = c_null_ptr
    contains
        procedure, public :: open => eng_open
        procedure, public :: close => eng_close
    end type

    !> vecfcn_helper for registered device residuals (set_fcn takes the registered name).
    type :: batch_vecfcn_helper
        integer(c_int), private :: m_fcn = -1
        integer(int32), private :: m_nfcn = 0
        integer(int32), private :: m_nvar = 0
        logical, private :: m_jac = .false.
    contains
        procedure, public :: set_fcn => bvh_set_fcn
        procedure, public :: set_jacobian => bvh_set_jac
        procedure, public :: is_fcn_defined => bvh_is_fcn_defined
        procedure, public :: is_jacobian_defined => bvh_is_jac_defined
        procedure, public :: get_equation_count => bvh_get_nfcn
        procedure, public :: get_variable_count => bvh_get_nvar
    end type

    !> line_search settings (nonlin_linesearch.f90:18-65)
    type :: batch_line_search
        integer(int32), private :: m_maxEval = 100
        real(real64), private :: m_alpha = 1.0d-4
        real(real64), private :: m_factor = 0.1d0
    contains
        procedure, public :: set_max_fcn_evals => bls_set_max_eval
        procedure, public :: set_scaling_factor => bls_set_scale
        procedure, public :: set_distance_factor => bls_set_dist
    end type

    !> equation_solver members (nonlin_multi_eqn_mult_var.f90:67-91)
    type, abstract :: batch_equation_solver
        integer(int32), private :: m_maxEval = 100
        real(real64), private :: m_fcnTol = 1.0d-8
        real(real64), private :: m_xtol = 1.0d-12
        real(real64), private :: m_gtol = 1.0d-12
    contains
        procedure, public :: set_max_fcn_evals => bes_set_max_eval
        procedure, public :: set_fcn_tolerance => bes_set_fcn_tol
        procedure, public :: set_var_tolerance => bes_set_var_tol
        procedure, public :: set_gradient_tolerance => bes_set_grad_tol
        procedure, public :: get_max_fcn_evals => bes_get_max_eval         ! nonlin_multi_eqn_mult_var.f90:302
        procedure, public :: get_fcn_tolerance => bes_get_fcn_tol          ! :324
        procedure, public :: get_var_tolerance => bes_get_var_tol          ! :344
        procedure, public :: get_gradient_tolerance => bes_get_grad_tol    ! :364
        procedure, public :: base_params => bes_params
    end type

    type, extends(batch_equation_solver) :: batch_least_squares_solver
        real(real64), private :: m_factor = 100.0d0
    contains
        procedure, public :: set_step_scaling_factor => bls_set_factor
        procedure, public :: solve_batch => lss_solve_batch
        procedure, public :: solve_sharded => lss_solve_sharded
    end type

    !> constrained_least_squares_solver (nonlin_least_squares.f90:34-75): limits, radius, step scaling
    type, extends(batch_equation_solver) :: batch_constrained_least_squares_solver
        real(real64), private, allocatable, dimension(:) :: m_upper
        real(real64), private, allocatable, dimension(:) :: m_lower
        real(real64), private :: m_delta = 1.0d0
        real(real64), private :: m_scaling = 1.0d0
    contains
        procedure, public :: set_upper_limits => bcls_set_upper
        procedure, public :: set_lower_limits => bcls_set_lower
        procedure, public :: set_trust_region_radius => bcls_set_radius
        procedure, public :: set_step_scaling_factor => bcls_set_factor
        procedure, public :: solve_batch => cls_solve_batch
    end type

    !> brent_solver / newton_1var_solver (nonlin_solve.f90:69-85) over B equations: m_newton selects the method
    type :: batch_solver_1var
        integer(int32) :: m_maxEval = 100
        real(real64) :: m_fcnTol = 1.0d-8
        real(real64) :: m_xtol = 1.0d-12
        real(real64) :: m_difftol = 1.0d-12
        logical :: m_newton = .false.
        logical :: m_useDiff = .false.
        integer(c_int) :: m_fcn = -1
    contains
        procedure, public :: set_fcn => bs1_set_fcn
        procedure, public :: solve_batch => bs1_solve_batch
    end type

    !> polynomial (nonlin_polynomials.f90:20-71) for B data sets: coefficients c(B, order + 1), c(:,1) = c0
    type :: batch_polynomial
        real(real64), allocatable, dimension(:,:) :: m_coeffs
    contains
        procedure, public :: order => bp_order
        procedure, public :: fit => bp_fit
        procedure, public :: fit_thru_zero => bp_fit_thru_zero
        procedure, public :: evaluate => bp_evaluate
    end type

    type, abstract, extends(batch_equation_solver) :: batch_line_search_solver
        type(batch_line_search), private :: m_lineSearch
        logical, private :: m_useLineSearch = .true.
    contains
        procedure, public :: set_line_search => blss_set_line_search
        procedure, public :: set_use_line_search => blss_set_use_search
        procedure, public :: ls_params => blss_params
    end type

    type, extends(batch_line_search_solver) :: batch_newton_solver
    contains
        procedure, public :: solve_batch => ns_solve_batch
    end type

    type, extends(batch_line_search_solver) :: batch_quasi_newton_solver
        integer(int32), private :: m_jDelta = 5
    contains
        procedure, public :: set_jacobian_interval => qns_set_jac_interval
        procedure, public :: solve_batch => qns_solve_batch
    end type

contains
    subroutine eng_open(this, device, ierr)
        class(nlb_engine), intent(inout) :: this
        integer, intent(in) :: device
        integer, intent(out) :: ierr
        ierr = nlb_create(this%handle, int(device, c_int))
    end subroutine

    subroutine eng_close(this)
        class(nlb_engine), intent(inout) :: this
        integer(c_int) :: rc
        if (c_associated(this%handle)) rc = nlb_destroy(this%handle)
        this%handle = c_null_ptr
    end subroutine

    subroutine bvh_set_fcn(this, name, nfcn, nvar)
        class(batch_vecfcn_helper), intent(inout) :: this
        character(len = *), intent(in) :: name
        integer(int32), intent(in) :: nfcn, nvar
        this%m_fcn = nlb_vecfcn_lookup(trim(name) // c_null_char)
        this%m_nfcn = nfcn
        this%m_nvar = nvar
    end subroutine

    subroutine bvh_set_jac(this, use_registered_jacobian)
        class(batch_vecfcn_helper), intent(inout) :: this
        logical, intent(in) :: use_registered_jacobian
        this%m_jac = use_registered_jacobian
    end subroutine

    pure logical function bvh_is_fcn_defined(this)
        class(batch_vecfcn_helper), intent(in) :: this
        bvh_is_fcn_defined = this%m_fcn >= 0
    end function

    pure logical function bvh_is_jac_defined(this)
        class(batch_vecfcn_helper), intent(in) :: this
        bvh_is_jac_defined = this%m_jac
    end function

    pure integer(int32) function bvh_get_nfcn(this)
        class(batch_vecfcn_helper), intent(in) :: this
        bvh_get_nfcn = this%m_nfcn
    end function

    pure integer(int32) function bvh_get_nvar(this)
        class(batch_vecfcn_helper), intent(in) :: this
        bvh_get_nvar = this%m_nvar
    end function

    subroutine bls_set_max_eval(this, x)
        class(batch_line_search), intent(inout) :: this
        integer(int32), intent(in) :: x
        this%m_maxEval = x
    end subroutine

    subroutine bls_set_scale(this, x)
        class(batch_line_search), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_alpha = x
    end subroutine

    subroutine bls_set_dist(this, x)
        class(batch_line_search), intent(inout) :: this
        real(real64), intent(in) :: x
        if (x <= 0.0d0) then
            this%m_factor = 0.1d0
        else if (x >= 1.0d0) then
            this%m_factor = 0.99d0
        else
            this%m_factor = x
        end if
    end subroutine

    subroutine bes_set_max_eval(this, n)
        class(batch_equation_solver), intent(inout) :: this
        integer(int32), intent(in) :: n
        this%m_maxEval = n
    end subroutine

    subroutine bes_set_fcn_tol(this, x)
        class(batch_equation_solver), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_fcnTol = x
    end subroutine

    subroutine bes_set_var_tol(this, x)
        class(batch_equation_solver), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_xtol = x
    end subroutine

    subroutine bes_set_grad_tol(this, x)
        class(batch_equation_solver), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_gtol = x
    end subroutine

    pure function bes_get_max_eval(this) result(n)
        class(batch_equation_solver), intent(in) :: this
        integer(int32) :: n
        n = this%m_maxEval
    end function

    pure function bes_get_fcn_tol(this) result(x)
        class(batch_equation_solver), intent(in) :: this
        real(real64) :: x
        x = this%m_fcnTol
    end function

    pure function bes_get_var_tol(this) result(x)
        class(batch_equation_solver), intent(in) :: this
        real(real64) :: x
        x = this%m_xtol
    end function

    pure function bes_get_grad_tol(this) result(x)
        class(batch_equation_solver), intent(in) :: this
        real(real64) :: x
        x = this%m_gtol
    end function

    function bes_params(this, fcn) result(p)
        class(batch_equation_solver), intent(in) :: this
        class(batch_vecfcn_helper), intent(in) :: fcn
        type(nlb_params) :: p
        call nlb_params_default(p)
        p%max_fcn_evals = this%m_maxEval
        p%fcn_tol = this%m_fcnTol
        p%var_tol = this%m_xtol
        p%grad_tol = this%m_gtol
        p%use_analytic_jacobian = merge(1_c_int32_t, 0_c_int32_t, fcn%m_jac)
    end function

    subroutine bls_set_factor(this, x)
        class(batch_least_squares_solver), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_factor = min(max(x, 0.1d0), 1.0d2)
    end subroutine

    subroutine blss_set_line_search(this, ls)
        class(batch_line_search_solver), intent(inout) :: this
        type(batch_line_search), intent(in) :: ls
        this%m_lineSearch = ls
    end subroutine

    subroutine blss_set_use_search(this, x)
        class(batch_line_search_solver), intent(inout) :: this
        logical, intent(in) :: x
        this%m_useLineSearch = x
    end subroutine

    function blss_params(this, fcn) result(p)
        class(batch_line_search_solver), intent(in) :: this
        class(batch_vecfcn_helper), intent(in) :: fcn
        type(nlb_params) :: p
        p = this%base_params(fcn)
        p%use_line_search = merge(1_c_int32_t, 0_c_int32_t, this%m_useLineSearch)
        p%ls_max_fcn_evals = this%m_lineSearch%m_maxEval
        p%ls_alpha = this%m_lineSearch%m_alpha
        p%ls_factor = this%m_lineSearch%m_factor
    end function

    subroutine qns_set_jac_interval(this, n)
        class(batch_quasi_newton_solver), intent(inout) :: this
        integer(int32), intent(in) :: n
        this%m_jDelta = n
    end subroutine

    !> Batch analogue of `call solver%solve(fcn, x, fvec, ib, args)`:
    !! x(B, n) in/out, fvec(B, m) out, ib(B), status(B); args(B, sys_len) optional per-system data,
    !! shared(:) optional batch-shared data.  ierr = API-level error (0 = ok); status(b) = 0 or
    !! the NL_* code the reference would have stopped with.
    subroutine lss_solve_batch(this, eng, fcn, x, fvec, ib, status, ierr, args, shared)
        class(batch_least_squares_solver), intent(in) :: this
        type(nlb_engine), intent(in) :: eng
        class(batch_vecfcn_helper), intent(in) :: fcn
        real(real64), intent(inout), dimension(:,:), contiguous, target :: x
        real(real64), intent(out), dimension(:,:), contiguous, target :: fvec
        type(batch_iteration_behavior), intent(out), dimension(:), target :: ib
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        real(real64), intent(in), dimension(:,:), contiguous, target, optional :: args
        real(real64), intent(in), dimension(:), contiguous, target, optional :: shared
        type(nlb_params) :: p
        type(c_ptr) :: pa, ps
        p = this%base_params(fcn)
        p%lm_factor = this%m_factor
        pa = c_null_ptr; ps = c_null_ptr
        if (present(args)) pa = c_loc(args)
        if (present(shared)) ps = c_loc(shared)
        ierr = nlb_least_squares_solve_batch(eng%handle, p, fcn%m_fcn, int(size(x, 1), c_int64_t), &
            int(fcn%m_nfcn, c_int), int(fcn%m_nvar, c_int), c_loc(x), c_loc(fvec), pa, ps, c_loc(ib), &
            c_loc(status), c_null_ptr)
    end subroutine

    !> The same solve over every engine of `engs` (one per GPU): contiguous system ranges, one host thread per device
    !> inside the engine, statistics (16 x int64, see NLB_STAT_* in nonlin_batch.h) combined with one NCCL all-reduce.
    !> x(B, n), fvec(B, m): host arrays.  The reference-side caller is still one `solve` call (nonlin_solver, :94-119).
    subroutine lss_solve_sharded(this, engs, fcn, x, fvec, ib, status, stats, ierr, args, shared)
        class(batch_least_squares_solver), intent(in) :: this
        type(nlb_engine), intent(in), dimension(:) :: engs
        class(batch_vecfcn_helper), intent(in) :: fcn
        real(real64), intent(inout), dimension(:,:), contiguous, target :: x
        real(real64), intent(out), dimension(:,:), contiguous, target :: fvec
        type(batch_iteration_behavior), intent(out), dimension(:), target :: ib
        integer(int32), intent(out), dimension(:), target :: status
        integer(c_int64_t), intent(out), dimension(16), target :: stats
        integer, intent(out) :: ierr
        real(real64), intent(in), dimension(:,:), contiguous, target, optional :: args
        real(real64), intent(in), dimension(:), contiguous, target, optional :: shared
        type(nlb_params) :: p
        type(c_ptr) :: pa, ps
        type(c_ptr), allocatable :: hs(:)
        integer :: d
        p = this%base_params(fcn)
        p%lm_factor = this%m_factor
        pa = c_null_ptr; ps = c_null_ptr
        if (present(args)) pa = c_loc(args)
        if (present(shared)) ps = c_loc(shared)
        allocate(hs(size(engs)))
        do d = 1, size(engs)
            hs(d) = engs(d)%handle
        end do
        ierr = nlb_solve_sharded(hs, int(size(engs), c_int), 0_c_int, p, fcn%m_fcn, int(size(x, 1), c_int64_t), &
            int(fcn%m_nfcn, c_int), int(fcn%m_nvar, c_int), c_loc(x), c_loc(fvec), pa, ps, c_loc(ib), c_loc(status), &
            c_loc(stats))
    end subroutine

    subroutine bcls_set_upper(this, x)
        class(batch_constrained_least_squares_solver), intent(inout) :: this
        real(real64), intent(in), dimension(:) :: x
        this%m_upper = x
    end subroutine

    subroutine bcls_set_lower(this, x)
        class(batch_constrained_least_squares_solver), intent(inout) :: this
        real(real64), intent(in), dimension(:) :: x
        this%m_lower = x
    end subroutine

    subroutine bcls_set_radius(this, x)
        class(batch_constrained_least_squares_solver), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_delta = merge(1.0d0, x, x <= 0.0d0)
    end subroutine

    subroutine bcls_set_factor(this, x)
        class(batch_constrained_least_squares_solver), intent(inout) :: this
        real(real64), intent(in) :: x
        this%m_scaling = merge(1.0d0, x, x <= 0.0d0)
    end subroutine

    !> Batch analogue of constrained_least_squares_solver%solve; arguments as lss_solve_batch.  The limits
    !! apply to every system of the batch; limit arrays whose length is not n are ignored, as cls_solve
    !! replaces them by -huge / +huge.
    subroutine cls_solve_batch(this, eng, fcn, x, fvec, ib, status, ierr, args, shared)
        class(batch_constrained_least_squares_solver), intent(in), target :: this
        type(nlb_engine), intent(in) :: eng
        class(batch_vecfcn_helper), intent(in) :: fcn
        real(real64), intent(inout), dimension(:,:), contiguous, target :: x
        real(real64), intent(out), dimension(:,:), contiguous, target :: fvec
        type(batch_iteration_behavior), intent(out), dimension(:), target :: ib
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        real(real64), intent(in), dimension(:,:), contiguous, target, optional :: args
        real(real64), intent(in), dimension(:), contiguous, target, optional :: shared
        type(nlb_params) :: p
        type(nlb_constrained_options) :: o
        type(c_ptr) :: pa, ps
        p = this%base_params(fcn)
        o%trust_region_radius = this%m_delta
        o%step_scaling_factor = this%m_scaling
        o%lower = c_null_ptr; o%upper = c_null_ptr
        if (allocated(this%m_lower)) then
            if (size(this%m_lower) == fcn%m_nvar) o%lower = c_loc(this%m_lower)
        end if
        if (allocated(this%m_upper)) then
            if (size(this%m_upper) == fcn%m_nvar) o%upper = c_loc(this%m_upper)
        end if
        pa = c_null_ptr; ps = c_null_ptr
        if (present(args)) pa = c_loc(args)
        if (present(shared)) ps = c_loc(shared)
        ierr = nlb_constrained_least_squares_solve_batch(eng%handle, p, o, fcn%m_fcn, int(size(x, 1), c_int64_t), &
            int(fcn%m_nfcn, c_int), int(fcn%m_nvar, c_int), c_loc(x), c_loc(fvec), pa, ps, c_loc(ib), &
            c_loc(status), c_null_ptr)
    end subroutine

    subroutine bs1_set_fcn(this, name)
        class(batch_solver_1var), intent(inout) :: this
        character(len = *), intent(in) :: name
        this%m_fcn = nlb_fcn1var_lookup(trim(name) // c_null_char)
    end subroutine

    !> `call solver%solve(fcn, x, lim, f, ib, args)` for B equations: lim1(B), lim2(B) = the value_pair of each
    !! equation, x(B) in/out, f(B) out, args(B, args_len) optional.
    subroutine bs1_solve_batch(this, eng, lim1, lim2, x, f, ib, status, ierr, args)
        class(batch_solver_1var), intent(in) :: this
        type(nlb_engine), intent(in) :: eng
        real(real64), intent(in), dimension(:), contiguous, target :: lim1, lim2
        real(real64), intent(inout), dimension(:), contiguous, target :: x
        real(real64), intent(out), dimension(:), contiguous, target :: f
        type(batch_iteration_behavior), intent(out), dimension(:), target :: ib
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        real(real64), intent(in), dimension(:,:), contiguous, target, optional :: args
        type(nlb_params_1var) :: p
        type(c_ptr) :: pa
        call nlb_params_1var_default(p)
        p%max_fcn_evals = this%m_maxEval
        p%fcn_tol = this%m_fcnTol
        p%var_tol = this%m_xtol
        p%diff_tol = this%m_difftol
        p%use_analytic_diff = merge(1_c_int32_t, 0_c_int32_t, this%m_useDiff)
        pa = c_null_ptr
        if (present(args)) pa = c_loc(args)
        if (this%m_newton) then
            ierr = nlb_newton_1var_solve_batch(eng%handle, p, this%m_fcn, int(size(x), c_int64_t), c_loc(lim1), &
                c_loc(lim2), c_loc(x), c_loc(f), pa, c_loc(ib), c_loc(status), c_null_ptr)
        else
            ierr = nlb_brent_solve_batch(eng%handle, p, this%m_fcn, int(size(x), c_int64_t), c_loc(lim1), &
                c_loc(lim2), c_loc(x), c_loc(f), pa, c_loc(ib), c_loc(status), c_null_ptr)
        end if
    end subroutine

    pure function bp_order(this) result(n)
        class(batch_polynomial), intent(in) :: this
        integer(int32) :: n
        n = -1
        if (allocated(this%m_coeffs)) n = size(this%m_coeffs, 2) - 1
    end function

    !> `call p%fit(x, y, order)` for B data sets: x(npts) shared abscissae, y(B, npts); status(B) = 0 or 107.
    subroutine bp_fit(this, eng, x, y, order, status, ierr)
        class(batch_polynomial), intent(inout), target :: this
        type(nlb_engine), intent(in) :: eng
        real(real64), intent(in), dimension(:), contiguous, target :: x
        real(real64), intent(in), dimension(:,:), contiguous, target :: y
        integer(int32), intent(in) :: order
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        if (allocated(this%m_coeffs)) deallocate(this%m_coeffs)
        allocate(this%m_coeffs(size(y, 1), order + 1))
        ierr = nlb_polynomial_fit_batch(eng%handle, int(size(y, 1), c_int64_t), int(size(x), c_int), &
            int(order, c_int), 0_c_int, 1_c_int, c_loc(x), c_loc(y), c_loc(this%m_coeffs), c_loc(status), c_null_ptr)
    end subroutine

    subroutine bp_fit_thru_zero(this, eng, x, y, order, status, ierr)
        class(batch_polynomial), intent(inout), target :: this
        type(nlb_engine), intent(in) :: eng
        real(real64), intent(in), dimension(:), contiguous, target :: x
        real(real64), intent(in), dimension(:,:), contiguous, target :: y
        integer(int32), intent(in) :: order
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        if (allocated(this%m_coeffs)) deallocate(this%m_coeffs)
        allocate(this%m_coeffs(size(y, 1), order + 1))
        ierr = nlb_polynomial_fit_batch(eng%handle, int(size(y, 1), c_int64_t), int(size(x), c_int), &
            int(order, c_int), 1_c_int, 1_c_int, c_loc(x), c_loc(y), c_loc(this%m_coeffs), c_loc(status), c_null_ptr)
    end subroutine

    !> `p%evaluate(x)`: yout(B, npts)
    subroutine bp_evaluate(this, eng, x, yout, ierr)
        class(batch_polynomial), intent(in), target :: this
        type(nlb_engine), intent(in) :: eng
        real(real64), intent(in), dimension(:), contiguous, target :: x
        real(real64), intent(out), dimension(:,:), contiguous, target :: yout
        integer, intent(out) :: ierr
        ierr = nlb_polynomial_evaluate_batch(eng%handle, int(size(this%m_coeffs, 1), c_int64_t), &
            int(this%order(), c_int), int(size(x), c_int), 1_c_int, c_loc(this%m_coeffs), c_loc(x), c_loc(yout), &
            c_null_ptr)
    end subroutine

    subroutine ns_solve_batch(this, eng, fcn, x, fvec, ib, status, ierr, args, shared)
        class(batch_newton_solver), intent(in) :: this
        type(nlb_engine), intent(in) :: eng
        class(batch_vecfcn_helper), intent(in) :: fcn
        real(real64), intent(inout), dimension(:,:), contiguous, target :: x
        real(real64), intent(out), dimension(:,:), contiguous, target :: fvec
        type(batch_iteration_behavior), intent(out), dimension(:), target :: ib
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        real(real64), intent(in), dimension(:,:), contiguous, target, optional :: args
        real(real64), intent(in), dimension(:), contiguous, target, optional :: shared
        type(nlb_params) :: p
        type(c_ptr) :: pa, ps
        p = this%ls_params(fcn)
        pa = c_null_ptr; ps = c_null_ptr
        if (present(args)) pa = c_loc(args)
        if (present(shared)) ps = c_loc(shared)
        ierr = nlb_newton_solve_batch(eng%handle, p, fcn%m_fcn, int(size(x, 1), c_int64_t), &
            int(fcn%m_nfcn, c_int), int(fcn%m_nvar, c_int), c_loc(x), c_loc(fvec), pa, ps, c_loc(ib), &
            c_loc(status), c_null_ptr)
    end subroutine

    subroutine qns_solve_batch(this, eng, fcn, x, fvec, ib, status, ierr, args, shared)
        class(batch_quasi_newton_solver), intent(in) :: this
        type(nlb_engine), intent(in) :: eng
        class(batch_vecfcn_helper), intent(in) :: fcn
        real(real64), intent(inout), dimension(:,:), contiguous, target :: x
        real(real64), intent(out), dimension(:,:), contiguous, target :: fvec
        type(batch_iteration_behavior), intent(out), dimension(:), target :: ib
        integer(int32), intent(out), dimension(:), target :: status
        integer, intent(out) :: ierr
        real(real64), intent(in), dimension(:,:), contiguous, target, optional :: args
        real(real64), intent(in), dimension(:), contiguous, target, optional :: shared
        type(nlb_params) :: p
        type(c_ptr) :: pa, ps
        p = this%ls_params(fcn)
        p%jacobian_interval = this%m_jDelta
        pa = c_null_ptr; ps = c_null_ptr
        if (present(args)) pa = c_loc(args)
        if (present(shared)) ps = c_loc(shared)
        ierr = nlb_quasi_newton_solve_batch(eng%handle, p, fcn%m_fcn, int(size(x, 1), c_int64_t), &
            int(fcn%m_nfcn, c_int), int(fcn%m_nvar, c_int), c_loc(x), c_loc(fvec), pa, ps, c_loc(ib), &
            c_loc(status), c_null_ptr)
    end subroutine
end module
