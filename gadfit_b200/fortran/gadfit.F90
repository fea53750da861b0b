! Fortran surface of the B200-native LM solver: module `gadfit` with the
! procedure names and keywords of the reference module
! (reference fortran/gadfit/gadfit.F90:41-58) whose bodies call the C-ABI of
! include/gadfit_b200.h through ISO_C_BINDING.
!
!   gadf_init(f, num_datasets, ...)          gadfit.F90:133-184
!   gadf_add_dataset(x_data, y_data, weights) gadfit.F90:232-246
!   gadf_set(dataset_i, par, val, active) / gadf_set(par, val, active)
!                                             gadfit.F90:255-342
!   gadf_fit(lambda, lam_up, lam_down, accth, ..., max_iter, damp_max, ...)
!                                             gadfit.F90:502-511
!   gadf_close                                gadfit.F90:1397-1414
!
! Deltas to the reference, all at the model-definition point or explicitly
! out of scope (DESIGN.md, "Out of scope"):
!  * `f` is the NAME of a fitting function compiled into the library
!    (character) instead of a polymorphic class(fitfunc): a host callback
!    cannot run on the device.
!  * Parameters are addressed by 1-based integer index as in the reference;
!    results are read with gadf_get(dataset_i, par) instead of
!    fitfuncs(i)%pars(j)%val.
!  * Weights are the reference's w = 1/sigma (gadfit.F90:683,690): they are
!    converted to the errors the C-ABI expects.
!  * Defaults are those of the FORTRAN reference, not of the C++ one:
!    lambda = 1, lam_incs = 2 (gadfit.F90:568-580).
!  * fit() runs the FORTRAN driver (gadfit_b200_gadf_options, enabled = 1):
!    uphill, nielsen, umnigh and the convergence tests (grad_chi2, cos_phi,
!    rel_error, rel_error_global, chi2_rel, chi2_abs) have the reference's
!    meaning (gadfit.F90:751-909); gadf_set_errors(NONE / SQRT_Y / PROPTO_Y /
!    INVERSE_Y / USER) as in gadfit.F90:445-470.
!  * load_balancing and use_ad are accepted for source compatibility and
!    ignored (the device always differentiates automatically; work is split
!    statically); file input and gadf_print are not provided.
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no gfortran.
! Build where one exists:  gfortran -c gadfit.F90 && gfortran prog.F90 gadfit.o
!                          -L<repo>/gadfit_b200 -lgadfit_b200
module gadfit

  use, intrinsic :: iso_c_binding
  use, intrinsic :: iso_fortran_env, only: real32, real64, error_unit

  implicit none

  private
  public :: gadf_init, gadf_add_dataset, gadf_set, gadf_set_errors, gadf_fit, &
       & gadf_get, gadf_chi2, gadf_iterations, gadf_stop_reason, gadf_close, kp
  public :: NONE, SQRT_Y, PROPTO_Y, INVERSE_Y, USER

  ! gadf_constants.F90: data error types
  enum, bind(c)
     enumerator :: NONE, SQRT_Y, PROPTO_Y, INVERSE_Y, USER
  end enum

  integer, parameter :: kp = real64

  ! POD mirror of gadfit_b200_settings (include/gadfit_b200.h)
  type, bind(c) :: settings_t
     integer(c_int) :: iteration_limit
     integer(c_int) :: lambda_incs
     real(c_double) :: lambda_down
     real(c_double) :: lambda_up
     integer(c_int) :: damp_max
     type(c_ptr) :: DTD_min
     integer(c_int) :: n_DTD_min
     real(c_double) :: acceleration_threshold
     integer(c_int) :: verbosity
     integer(c_int) :: loss
     integer(c_int) :: n_threads
  end type settings_t

  ! POD mirror of gadfit_b200_gadf_options
  type, bind(c) :: gadf_options_t
     integer(c_int) :: enabled
     integer(c_int) :: uphill
     integer(c_int) :: nielsen
     integer(c_int) :: umnigh
     real(c_double) :: chi2_abs
     real(c_double) :: chi2_rel
     real(c_double) :: grad_chi2
     real(c_double) :: cos_phi
     real(c_double) :: rel_error
     real(c_double) :: rel_error_global
  end type gadf_options_t

  interface
     integer(c_int) function c_set_gadf_options(handle, o) &
          & bind(c, name='gadfit_b200_set_gadf_options')
       import :: c_int, c_ptr, gadf_options_t
       type(c_ptr), value :: handle
       type(gadf_options_t), intent(in) :: o
     end function c_set_gadf_options
     integer(c_int) function c_get_stop_reason(handle, reason) &
          & bind(c, name='gadfit_b200_get_stop_reason')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), intent(out) :: reason
     end function c_get_stop_reason
     integer(c_int) function c_create(model_name, handle) &
          & bind(c, name='gadfit_b200_create')
       import :: c_int, c_char, c_ptr
       character(kind=c_char), intent(in) :: model_name(*)
       type(c_ptr), intent(out) :: handle
     end function c_create
     integer(c_int) function c_destroy(handle) bind(c, name='gadfit_b200_destroy')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function c_destroy
     integer(c_int) function c_add_dataset(handle, n, x, y, errors) &
          & bind(c, name='gadfit_b200_add_dataset')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: n
       type(c_ptr), value :: x, y, errors
     end function c_add_dataset
     integer(c_int) function c_set_par(handle, i_par, val, active, i_dataset, name) &
          & bind(c, name='gadfit_b200_set_par')
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: i_par, active, i_dataset
       real(c_double), value :: val
       type(c_ptr), value :: name
     end function c_set_par
     integer(c_int) function c_default_settings(s) &
          & bind(c, name='gadfit_b200_default_settings')
       import :: c_int, settings_t
       type(settings_t), intent(out) :: s
     end function c_default_settings
     integer(c_int) function c_set_settings(handle, s) &
          & bind(c, name='gadfit_b200_set_settings')
       import :: c_int, c_ptr, settings_t
       type(c_ptr), value :: handle
       type(settings_t), intent(in) :: s
     end function c_set_settings
     integer(c_int) function c_fit(handle, lambda) bind(c, name='gadfit_b200_fit')
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: handle
       real(c_double), value :: lambda
     end function c_fit
     integer(c_int) function c_get_par(handle, i_par, i_dataset, val) &
          & bind(c, name='gadfit_b200_get_par')
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: i_par, i_dataset
       real(c_double), intent(out) :: val
     end function c_get_par
     integer(c_int) function c_chi2(handle, val) bind(c, name='gadfit_b200_chi2')
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: handle
       real(c_double), intent(out) :: val
     end function c_chi2
     integer(c_int) function c_n_iterations(handle, n) &
          & bind(c, name='gadfit_b200_n_iterations')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), intent(out) :: n
     end function c_n_iterations
     function c_last_error(handle) bind(c, name='gadfit_b200_last_error')
       import :: c_ptr
       type(c_ptr), value :: handle
       type(c_ptr) :: c_last_error
     end function c_last_error
  end interface

  interface gadf_set
     module procedure set_int_local_real, set_int_local_real32, &
          & set_int_global_real, set_int_global_real32
  end interface gadf_set

  type(c_ptr) :: handle = c_null_ptr
  integer :: num_datasets_expected = 1, num_datasets_added = 0
  integer :: data_error_type = NONE
  ! Errors handed to the library must outlive gadf_fit: one array per dataset.
  ! They are BORROWED by the library until the first fit uploads them, so
  ! gadf_fit can still fill them according to gadf_set_errors.
  type :: errors_t
     real(kp), allocatable :: sigma(:)
     real(kp), pointer :: y(:) => null()
     logical :: user_weights = .false.
  end type errors_t
  type(errors_t), allocatable, target :: errors(:)

contains

  subroutine check(status, where)
    integer(c_int), intent(in) :: status
    character(*), intent(in) :: where
    if (status /= 0) then
       write(error_unit, '(a,i0)') 'gadfit (B200): '//where//' failed, status ', status
       error stop
    end if
  end subroutine check

  ! gadfit.F90:133-184.  Only f and num_datasets have a meaning here; the AD
  ! tape and integration workspace sizes of the reference have no counterpart
  ! (registers / per-thread panel stacks).  integration_rule: the rule is part
  ! of the compiled fitting function (c.template integrate<PTS>(...), PTS = 15,
  ! 21, 31, 41, 51 or 61, include/gadfit_b200/quadrature.h), so the argument is
  ! accepted for source compatibility and not used.
  subroutine gadf_init(f, num_datasets, sweep_size, trace_size, const_size, &
       & ws_size, ws_size_inner, integration_rule, ad_memory, &
       & rel_error_inner, rel_error)
    character(*), intent(in) :: f
    integer, intent(in), optional :: num_datasets, sweep_size, trace_size, &
         & const_size, ws_size, ws_size_inner, integration_rule
    character(*), intent(in), optional :: ad_memory
    real(kp), intent(in), optional :: rel_error_inner, rel_error
    character(kind=c_char), allocatable :: cname(:)
    integer :: i
    allocate(cname(len_trim(f) + 1))
    do i = 1, len_trim(f)
       cname(i) = f(i:i)
    end do
    cname(len_trim(f) + 1) = c_null_char
    call check(c_create(cname, handle), 'gadf_init')
    num_datasets_expected = 1
    if (present(num_datasets)) num_datasets_expected = num_datasets
    num_datasets_added = 0
    if (allocated(errors)) deallocate(errors)
    allocate(errors(num_datasets_expected))
  end subroutine gadf_init

  ! gadfit.F90:232-246 (data version).  x_data / y_data are borrowed until
  ! gadf_fit returns, as in the reference (pointer association).
  subroutine gadf_add_dataset(x_data, y_data, weights)
    real(kp), intent(in), target, contiguous :: x_data(:), y_data(:)
    real(kp), intent(in), target, optional :: weights(:)
    type(c_ptr) :: e
    if (num_datasets_added >= num_datasets_expected) then
       write(error_unit, '(a)') 'gadfit (B200): too many calls to gadf_add_dataset'
       error stop
    end if
    num_datasets_added = num_datasets_added + 1
    ! sigma is filled in gadf_fit (init_weights, gadfit.F90:445-470); with
    ! USER errors the third column holds the uncertainties themselves
    allocate(errors(num_datasets_added)%sigma(size(y_data)))
    errors(num_datasets_added)%sigma = 1.0_kp
    errors(num_datasets_added)%y => y_data
    if (present(weights)) then
       errors(num_datasets_added)%sigma = weights
       errors(num_datasets_added)%user_weights = .true.
    end if
    e = c_loc(errors(num_datasets_added)%sigma)
    call check(c_add_dataset(handle, int(size(x_data), c_int), c_loc(x_data), &
         & c_loc(y_data), e), 'gadf_add_dataset')
  end subroutine gadf_add_dataset

  ! gadfit.F90:255-342: 1-based dataset and parameter indices
  subroutine set_int_local_real(dataset_i, par, val, active)
    integer, intent(in) :: dataset_i, par
    real(kp), intent(in) :: val
    logical, intent(in), optional :: active
    integer(c_int) :: a
    a = 0
    if (present(active)) a = merge(1, 0, active)
    call check(c_set_par(handle, int(par - 1, c_int), val, a, &
         & int(dataset_i - 1, c_int), c_null_ptr), 'gadf_set')
  end subroutine set_int_local_real

  subroutine set_int_local_real32(dataset_i, par, val, active)
    integer, intent(in) :: dataset_i, par
    real(real32), intent(in) :: val
    logical, intent(in), optional :: active
    call set_int_local_real(dataset_i, par, real(val, kp), active)
  end subroutine set_int_local_real32

  subroutine set_int_global_real(par, val, active)
    integer, intent(in) :: par
    real(kp), intent(in) :: val
    logical, intent(in), optional :: active
    integer(c_int) :: a
    a = 0
    if (present(active)) a = merge(1, 0, active)
    call check(c_set_par(handle, int(par - 1, c_int), val, a, -1_c_int, &
         & c_null_ptr), 'gadf_set')
  end subroutine set_int_global_real

  subroutine set_int_global_real32(par, val, active)
    integer, intent(in) :: par
    real(real32), intent(in) :: val
    logical, intent(in), optional :: active
    call set_int_global_real(par, real(val, kp), active)
  end subroutine set_int_global_real32

  ! gadfit.F90:391-394
  subroutine gadf_set_errors(e)
    integer, intent(in) :: e
    data_error_type = e
  end subroutine gadf_set_errors

  ! init_weights, gadfit.F90:445-470, expressed as sigma = 1/weight
  subroutine init_sigma()
    integer :: i
    do i = 1, num_datasets_added
       associate(sigma => errors(i)%sigma, y => errors(i)%y)
         select case(data_error_type)
         case(NONE)
            sigma = 1.0_kp
         case(SQRT_Y)
            where(abs(y) < 1d2*tiny(1.0_kp))
               sigma = huge(1.0_kp) ! weight 0
            elsewhere
               sigma = sqrt(y)
            end where
         case(PROPTO_Y)
            where(abs(y) < 1d2*tiny(1.0_kp))
               sigma = huge(1.0_kp)
            elsewhere
               sigma = y
            end where
         case(INVERSE_Y)
            sigma = 1.0_kp/y
         case(USER)
            if (.not. errors(i)%user_weights) then
               write(error_unit, '(a)') 'gadfit (B200): USER errors need the &
                    &third argument of gadf_add_dataset'
               error stop
            end if
         end select
       end associate
    end do
  end subroutine init_sigma

  ! gadfit.F90:502-511, defaults :568-580
  subroutine gadf_fit(lambda, lam_up, lam_down, accth, grad_chi2, cos_phi, &
       & rel_error, rel_error_global, chi2_rel, chi2_abs, DTD_min, lam_incs, &
       & uphill, max_iter, damp_max, nielsen, umnigh, load_balancing, use_ad)
    real(real32), intent(in), optional :: lambda, lam_up, lam_down, accth, &
         & grad_chi2, cos_phi, rel_error, rel_error_global, chi2_rel, chi2_abs
    real(kp), intent(in), optional, target :: DTD_min(:)
    integer, intent(in), optional :: lam_incs, uphill, max_iter
    logical, intent(in), optional :: damp_max, nielsen, umnigh, use_ad
    logical, value, optional :: load_balancing
    type(settings_t) :: s
    type(gadf_options_t) :: o
    real(kp) :: lambda_loc
    logical, save :: sigma_ready = .false.
    if (.not. sigma_ready) then
       call init_sigma() ! once, like init_weights (gadfit.F90:565-568)
       sigma_ready = .true.
    end if
    call check(c_default_settings(s), 'gadf_fit')
    s%iteration_limit = huge(1_c_int) ! no max_iter: run until a test stops it
    lambda_loc = 1.0_kp
    s%lambda_incs = 2
    if (present(lambda)) lambda_loc = real(lambda, kp)
    if (present(lam_up)) s%lambda_up = real(lam_up, kp)
    if (present(lam_down)) s%lambda_down = real(lam_down, kp)
    if (present(accth)) s%acceleration_threshold = real(accth, kp)
    if (present(lam_incs)) s%lambda_incs = lam_incs
    if (present(max_iter)) s%iteration_limit = max_iter
    if (present(damp_max)) s%damp_max = merge(1, 0, damp_max)
    if (present(DTD_min)) then
       s%DTD_min = c_loc(DTD_min)
       s%n_DTD_min = size(DTD_min)
    end if
    ! the Fortran driver and its own options (gadfit_b200_gadf_options);
    ! a negative tolerance means "argument absent"
    o = gadf_options_t(1, 0, 0, 0, -1.0_kp, -1.0_kp, -1.0_kp, -1.0_kp, -1.0_kp, -1.0_kp)
    if (present(uphill)) o%uphill = uphill
    if (present(nielsen)) o%nielsen = merge(1, 0, nielsen)
    if (present(umnigh)) o%umnigh = merge(1, 0, umnigh)
    if (present(chi2_abs)) o%chi2_abs = real(chi2_abs, kp)
    if (present(chi2_rel)) o%chi2_rel = real(chi2_rel, kp)
    if (present(grad_chi2)) o%grad_chi2 = real(grad_chi2, kp)
    if (present(cos_phi)) o%cos_phi = real(cos_phi, kp)
    if (present(rel_error)) o%rel_error = real(rel_error, kp)
    if (present(rel_error_global)) o%rel_error_global = real(rel_error_global, kp)
    call check(c_set_gadf_options(handle, o), 'gadf_fit')
    call check(c_set_settings(handle, s), 'gadf_fit')
    call check(c_fit(handle, lambda_loc), 'gadf_fit')
  end subroutine gadf_fit

  ! Replaces reading fitfuncs(dataset_i)%pars(par)%val
  real(kp) function gadf_get(dataset_i, par) result(val)
    integer, intent(in) :: dataset_i, par
    call check(c_get_par(handle, int(par - 1, c_int), int(dataset_i - 1, c_int), &
         & val), 'gadf_get')
  end function gadf_get

  real(kp) function gadf_chi2() result(val)
    call check(c_chi2(handle, val), 'gadf_chi2')
  end function gadf_chi2

  integer function gadf_iterations() result(n)
    integer(c_int) :: m
    call check(c_n_iterations(handle, m), 'gadf_iterations')
    n = m
  end function gadf_iterations

  ! 0 max_iter, 1 lam_incs, 2 chi2_abs, 3 chi2_rel, 4 grad_chi2, 5 cos_phi,
  ! 6 rel_error, 7 rel_error_global (the messages of gadfit.F90:812-908)
  integer function gadf_stop_reason() result(reason)
    integer(c_int) :: r
    call check(c_get_stop_reason(handle, r), 'gadf_stop_reason')
    reason = r
  end function gadf_stop_reason

  ! gadfit.F90:1397-1414
  subroutine gadf_close()
    if (c_associated(handle)) call check(c_destroy(handle), 'gadf_close')
    handle = c_null_ptr
    if (allocated(errors)) deallocate(errors)
  end subroutine gadf_close

end module gadfit
