// nonlin_batch.hpp — header-only C++ mirror of the reference's solver objects over the C ABI
// (include/nonlin_batch.h).  Same names and argument meaning as jchristopherson/nonlin
// (SURVEY.md App. C); no numerics here.  Link with -lnonlin_b200.
//
//   nonlin::engine eng(0);
//   nonlin::vecfcn_helper obj;          obj.set_fcn("misc_2fcn", 2, 2);
//   nonlin::quasi_newton_solver solver; solver.set_jacobian_interval(20);
//   solver.solve(eng, obj, B, x, fvec, ib, status);      // x[j*B + b] in/out, host or device pointers
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nonlin_batch.h"

namespace nonlin {

struct error : std::runtime_error {
    int code;
    error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

using iteration_behavior = nlb_iteration_behavior;   // reference src/nonlin_types.f90:8-29

class engine {
public:
    explicit engine(int device = 0) {
        int rc = nlb_create(&h_, device);
        if (rc != NLB_OK) throw error(rc, "nlb_create failed: no usable CUDA device (the engine has no CPU fallback)");
    }
    ~engine() { if (h_) nlb_destroy(h_); }
    engine(const engine&) = delete;
    engine& operator=(const engine&) = delete;
    nlb_handle* get() const { return h_; }
    void check(int rc) const { if (rc != NLB_OK) throw error(rc, nlb_last_error(h_)); }
private:
    nlb_handle* h_ = nullptr;
};

// reference src/nonlin_multi_eqn_mult_var.f90:41-65
class vecfcn_helper {
public:
    void set_fcn(const std::string& registered_name, int nfcn, int nvar) {
        id_ = nlb_vecfcn_lookup(registered_name.c_str());
        if (id_ < 0) throw error(NLB_ERR_UNKNOWN_FCN, "residual '" + registered_name + "' is not registered");
        m_ = nfcn; n_ = nvar; jac_ = false;
    }
    void set_jacobian(bool use_registered = true) { jac_ = use_registered; }
    void set_shared_data(const double* shared) { shared_ = shared; }
    bool is_fcn_defined() const { return id_ >= 0; }
    bool is_jacobian_defined() const { return jac_; }
    int get_equation_count() const { return m_; }
    int get_variable_count() const { return n_; }
    int id() const { return id_; }
    const double* shared() const { return shared_; }
private:
    int id_ = -1, m_ = 0, n_ = 0;
    bool jac_ = false;
    const double* shared_ = nullptr;
};

// reference src/nonlin_linesearch.f90:18-65
class line_search {
public:
    int get_max_fcn_evals() const { return max_eval_; }
    void set_max_fcn_evals(int x) { max_eval_ = x; }
    double get_scaling_factor() const { return alpha_; }
    void set_scaling_factor(double x) { alpha_ = x; }
    double get_distance_factor() const { return factor_; }
    void set_distance_factor(double x) { factor_ = x <= 0.0 ? 0.1 : (x >= 1.0 ? 0.99 : x); }
private:
    int max_eval_ = 100;
    double alpha_ = 1.0e-4, factor_ = 0.1;
};

// reference src/nonlin_multi_eqn_mult_var.f90:67-91
class equation_solver {
public:
    virtual ~equation_solver() = default;
    int get_max_fcn_evals() const { return p_.max_fcn_evals; }
    void set_max_fcn_evals(int n) { p_.max_fcn_evals = n; }
    double get_fcn_tolerance() const { return p_.fcn_tol; }
    void set_fcn_tolerance(double x) { p_.fcn_tol = x; }
    double get_var_tolerance() const { return p_.var_tol; }
    void set_var_tolerance(double x) { p_.var_tol = x; }
    double get_gradient_tolerance() const { return p_.grad_tol; }
    void set_gradient_tolerance(double x) { p_.grad_tol = x; }
    bool get_print_status() const { return print_; }
    void set_print_status(bool x) { print_ = x; }   // stored, not acted on (device-resident iterations)

    // `call solver%solve(fcn, x, fvec, ib, args)` over B systems; returns nothing, throws on API errors,
    // per-system outcome in status[] (0 or the NL_* code of the reference's `error stop`).
    void solve(const engine& eng, const vecfcn_helper& fcn, int64_t B, double* x, double* fvec,
               iteration_behavior* ib = nullptr, int32_t* status = nullptr, const double* args = nullptr,
               void* stream = nullptr) {
        if (!fcn.is_fcn_defined()) throw error(NLB_ERR_UNKNOWN_FCN, "no residual set (NL_UNDEFINED_FUNCTION_ERROR)");
        nlb_params p = p_;
        p.use_analytic_jacobian = fcn.is_jacobian_defined();
        eng.check(launch(eng.get(), p, fcn, B, x, fvec, ib, status, args, stream));
    }
protected:
    equation_solver() { nlb_params_default(&p_); }
    virtual int launch(nlb_handle* h, const nlb_params& p, const vecfcn_helper& f, int64_t B, double* x, double* fvec,
                       iteration_behavior* ib, int32_t* status, const double* args, void* stream) = 0;
    nlb_params p_;
    bool print_ = false;
};

// reference src/nonlin_least_squares.f90:20-31
class least_squares_solver : public equation_solver {
public:
    double get_step_scaling_factor() const { return p_.lm_factor; }
    void set_step_scaling_factor(double x) { p_.lm_factor = x < 0.1 ? 0.1 : (x > 100.0 ? 100.0 : x); }
protected:
    int launch(nlb_handle* h, const nlb_params& p, const vecfcn_helper& f, int64_t B, double* x, double* fvec,
               iteration_behavior* ib, int32_t* status, const double* args, void* stream) override {
        return nlb_least_squares_solve_batch(h, &p, f.id(), B, f.get_equation_count(), f.get_variable_count(), x, fvec,
                                             args, f.shared(), ib, status, stream);
    }
};

// reference src/nonlin_least_squares.f90:34-48
class constrained_equation_solver : public equation_solver {
public:
    std::vector<double> get_upper_limits() const { return upper_; }
    void set_upper_limits(const std::vector<double>& x) { upper_ = x; }
    std::vector<double> get_lower_limits() const { return lower_; }
    void set_lower_limits(const std::vector<double>& x) { lower_ = x; }
    // ces_apply_limits (:858-883) on one host vector: lower limits first, then upper
    void apply_limits(std::vector<double>& x) const {
        for (size_t i = 0; i < x.size() && i < lower_.size(); ++i) if (x[i] < lower_[i]) x[i] = lower_[i];
        for (size_t i = 0; i < x.size() && i < upper_.size(); ++i) if (x[i] > upper_[i]) x[i] = upper_[i];
    }
protected:
    std::vector<double> lower_, upper_;
};

// reference src/nonlin_least_squares.f90:50-75
class constrained_least_squares_solver : public constrained_equation_solver {
public:
    double get_trust_region_radius() const { return delta_; }
    void set_trust_region_radius(double x) { delta_ = x <= 0.0 ? 1.0 : x; }
    double get_step_scaling_factor() const { return scaling_; }
    void set_step_scaling_factor(double x) { scaling_ = x <= 0.0 ? 1.0 : x; }
protected:
    int launch(nlb_handle* h, const nlb_params& p, const vecfcn_helper& f, int64_t B, double* x, double* fvec,
               iteration_behavior* ib, int32_t* status, const double* args, void* stream) override {
        nlb_constrained_options o;
        nlb_constrained_options_default(&o);
        o.trust_region_radius = delta_;
        o.step_scaling_factor = scaling_;
        // limit arrays of the wrong length are replaced by -huge / +huge (cls_solve :1014-1024)
        const size_t n = (size_t)f.get_variable_count();
        if (lower_.size() == n) o.lower = lower_.data();
        if (upper_.size() == n) o.upper = upper_.data();
        return nlb_constrained_least_squares_solve_batch(h, &p, &o, f.id(), B, f.get_equation_count(),
                                                         f.get_variable_count(), x, fvec, args, f.shared(), ib, status,
                                                         stream);
    }
private:
    double delta_ = 1.0, scaling_ = 1.0;
};

// reference src/nonlin_solve.f90:20-41
class line_search_solver : public equation_solver {
public:
    void set_line_search(const line_search& ls) {
        p_.ls_max_fcn_evals = ls.get_max_fcn_evals();
        p_.ls_alpha = ls.get_scaling_factor();
        p_.ls_factor = ls.get_distance_factor();
        defined_ = true;
    }
    void set_default_line_search() { set_line_search(line_search()); }
    bool is_line_search_defined() const { return defined_; }
    bool get_use_line_search() const { return p_.use_line_search != 0; }
    void set_use_line_search(bool x) { p_.use_line_search = x; }
private:
    bool defined_ = false;
};

// reference src/nonlin_solve.f90:43-58
class quasi_newton_solver : public line_search_solver {
public:
    int get_jacobian_interval() const { return p_.jacobian_interval; }
    void set_jacobian_interval(int n) { p_.jacobian_interval = n; }
protected:
    int launch(nlb_handle* h, const nlb_params& p, const vecfcn_helper& f, int64_t B, double* x, double* fvec,
               iteration_behavior* ib, int32_t* status, const double* args, void* stream) override {
        return nlb_quasi_newton_solve_batch(h, &p, f.id(), B, f.get_equation_count(), f.get_variable_count(), x, fvec,
                                            args, f.shared(), ib, status, stream);
    }
};

// reference src/nonlin_solve.f90:60-67
class newton_solver : public line_search_solver {
protected:
    int launch(nlb_handle* h, const nlb_params& p, const vecfcn_helper& f, int64_t B, double* x, double* fvec,
               iteration_behavior* ib, int32_t* status, const double* args, void* stream) override {
        return nlb_newton_solve_batch(h, &p, f.id(), B, f.get_equation_count(), f.get_variable_count(), x, fvec, args,
                                      f.shared(), ib, status, stream);
    }
};

// reference src/nonlin_types.f90:31-37 — search limits of the one-variable solvers, one pair per equation
struct value_pair_batch {
    const double* x1;
    const double* x2;
};

// reference src/nonlin_single_var.f90:25-41
class fcn1var_helper {
public:
    void set_fcn(const std::string& registered_name) {
        id_ = nlb_fcn1var_lookup(registered_name.c_str());
        if (id_ < 0) throw error(NLB_ERR_UNKNOWN_FCN, "one-variable function '" + registered_name + "' is not registered");
        diff_ = false;
    }
    void set_diff(bool use_registered = true) { diff_ = use_registered; }
    bool is_fcn_defined() const { return id_ >= 0; }
    bool is_derivative_defined() const { return diff_; }
    int id() const { return id_; }
private:
    int id_ = -1;
    bool diff_ = false;
};

// reference src/nonlin_single_var.f90:43-69
class equation_solver_1var {
public:
    virtual ~equation_solver_1var() = default;
    int get_max_fcn_evals() const { return p_.max_fcn_evals; }
    void set_max_fcn_evals(int n) { p_.max_fcn_evals = n; }
    double get_fcn_tolerance() const { return p_.fcn_tol; }
    void set_fcn_tolerance(double x) { p_.fcn_tol = x; }
    double get_var_tolerance() const { return p_.var_tol; }
    void set_var_tolerance(double x) { p_.var_tol = x; }
    double get_diff_tolerance() const { return p_.diff_tol; }
    void set_diff_tolerance(double x) { p_.diff_tol = x; }
    // `call solver%solve(fcn, x, lim, f, ib, args)` over B equations; f may be null (optional argument absent)
    void solve(const engine& eng, const fcn1var_helper& fcn, int64_t B, double* x, const value_pair_batch& lim,
               double* f = nullptr, iteration_behavior* ib = nullptr, int32_t* status = nullptr,
               const double* args = nullptr, void* stream = nullptr) {
        if (!fcn.is_fcn_defined()) throw error(NLB_ERR_UNKNOWN_FCN, "no function set (NL_UNDEFINED_FUNCTION_ERROR)");
        nlb_params_1var p = p_;
        p.use_analytic_diff = fcn.is_derivative_defined();
        eng.check(launch(eng.get(), p, fcn.id(), B, lim.x1, lim.x2, x, f, args, ib, status, stream));
    }
protected:
    equation_solver_1var() { nlb_params_1var_default(&p_); }
    virtual int launch(nlb_handle* h, const nlb_params_1var& p, int id, int64_t B, const double* l1, const double* l2,
                       double* x, double* f, const double* args, iteration_behavior* ib, int32_t* status,
                       void* stream) = 0;
    nlb_params_1var p_;
};

// reference src/nonlin_solve.f90:69-76
class brent_solver : public equation_solver_1var {
protected:
    int launch(nlb_handle* h, const nlb_params_1var& p, int id, int64_t B, const double* l1, const double* l2, double* x,
               double* f, const double* args, iteration_behavior* ib, int32_t* status, void* stream) override {
        return nlb_brent_solve_batch(h, &p, id, B, l1, l2, x, f, args, ib, status, stream);
    }
};

// reference src/nonlin_solve.f90:78-85
class newton_1var_solver : public equation_solver_1var {
protected:
    int launch(nlb_handle* h, const nlb_params_1var& p, int id, int64_t B, const double* l1, const double* l2, double* x,
               double* f, const double* args, iteration_behavior* ib, int32_t* status, void* stream) override {
        return nlb_newton_1var_solve_batch(h, &p, id, B, l1, l2, x, f, args, ib, status, stream);
    }
};

// reference src/nonlin_polynomials.f90:20-71 — a batch of B polynomials of one order; coefficients
// c[k*B + b], k = 0..order (get(i) of the reference is row i - 1).  Host buffers.
class polynomial {
public:
    int order() const { return order_; }
    int64_t count() const { return B_; }
    const std::vector<double>& get_all() const { return c_; }
    double get(int i, int64_t b = 0) const { return c_[(size_t)(i - 1) * B_ + b]; }
    void set(int i, double v, int64_t b = 0) { c_[(size_t)(i - 1) * B_ + b] = v; }
    void initialize(int order, int64_t B = 1) {
        if (order < 0) throw error(NLB_ERR_INVALID_ARGUMENT, "order must be >= 0");
        order_ = order; B_ = B; c_.assign((size_t)(order + 1) * B, 0.0);
    }
    // `call p%fit(x, y, order)` / `p%fit_thru_zero`: y[i*B + b]; x[i] (shared) or x[i*B + b]; host or device pointers
    void fit(const engine& eng, int64_t B, int npts, const double* x, bool x_is_shared, const double* y, int order,
             int32_t* status = nullptr, bool thru_zero = false) {
        initialize(order, B);
        eng.check(nlb_polynomial_fit_batch(eng.get(), B, npts, order, thru_zero, x_is_shared, x, y, c_.data(), status,
                                           nullptr));
    }
    void fit_thru_zero(const engine& eng, int64_t B, int npts, const double* x, bool x_is_shared, const double* y,
                       int order, int32_t* status = nullptr) {
        fit(eng, B, npts, x, x_is_shared, y, order, status, true);
    }
    // `p%evaluate(x)`: yout[i*B + b]
    void evaluate(const engine& eng, int npts, const double* x, bool x_is_shared, double* yout) const {
        eng.check(nlb_polynomial_evaluate_batch(eng.get(), B_, order_, npts, x_is_shared, c_.data(), x, yout, nullptr));
    }
private:
    int order_ = -1;
    int64_t B_ = 0;
    std::vector<double> c_;
};

}  // namespace nonlin
