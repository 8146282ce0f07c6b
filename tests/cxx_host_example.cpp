// README Example 1 of the reference (examples/nonlin_quasi_newton_example.f90) through the C++ host
// mirror, B copies.  Built and run by tests/test_gpu_parity.py::test_cxx_host_mirror on the GPU box;
// compiled (not run) by tests/test_abi.py on CPU.
#include <cstdio>
#include <vector>

#include "../nonlin_b200/host/nonlin_batch.hpp"

int main() {
    const int64_t B = 1000;
    try {
        nonlin::engine eng(0);
        nonlin::vecfcn_helper obj;
        obj.set_fcn("misc_2fcn", 2, 2);
        nonlin::quasi_newton_solver solver;
        solver.set_jacobian_interval(20);
        solver.set_fcn_tolerance(1.0e-8);
        solver.set_var_tolerance(1.0e-12);
        solver.set_gradient_tolerance(1.0e-12);
        std::vector<double> x(2 * B, 1.0), f(2 * B);
        std::vector<nonlin::iteration_behavior> ib(B);
        std::vector<int32_t> status(B);
        solver.solve(eng, obj, B, x.data(), f.data(), ib.data(), status.data());
        for (int64_t b = 0; b < B; ++b) {
            if (status[b] != 0 || ib[b].iter_count != 11 || ib[b].fcn_count != 15 || ib[b].jacobian_count != 1) {
                std::printf("FAIL at %lld\n", (long long)b);
                return 1;
            }
        }
        // test_constrained_least_squares_bounds (tests/nonlin_test_solve.f90:1186): start (1, 1) outside the box
        // [4, 5.6] x [2, 3.6]; the solution must be feasible and is the root (5, 3)
        nonlin::constrained_least_squares_solver csolver;
        csolver.set_lower_limits({4.0, 2.0});
        csolver.set_upper_limits({5.6, 3.6});
        std::vector<double> xc(2 * B, 1.0), fc(2 * B);
        std::vector<nonlin::iteration_behavior> ibc(B);
        std::vector<int32_t> statusc(B);
        csolver.solve(eng, obj, B, xc.data(), fc.data(), ibc.data(), statusc.data());
        for (int64_t b = 0; b < B; ++b) {
            const double d0 = xc[b] - 5.0, d1 = xc[B + b] - 3.0;
            if (statusc[b] != 0 || ibc[b].converge_on_fcn == 0 || d0 > 1e-6 || d0 < -1e-6 || d1 > 1e-6 || d1 < -1e-6) {
                std::printf("FAIL (constrained) at %lld\n", (long long)b);
                return 1;
            }
        }
        // README Example 3: cubic through the 21 points of Example 2, c0 = 1.1866141861 ... c3 = 1.0647628218
        const double xp[21] = {0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0,
                               1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2.0};
        const double yp[21] = {1.216737514, 1.250032542, 1.305579195, 1.040182335, 1.751867738, 1.109716707,
                               2.018141531, 1.992418729, 1.807916923, 2.078806005, 2.698801324, 2.644662712,
                               3.412756702, 4.406137221, 4.567156645, 4.999550779, 5.652854194, 6.784320119,
                               8.307936836, 8.395126494, 10.30252404};
        std::vector<double> yb(21 * B);
        for (int i = 0; i < 21; ++i)
            for (int64_t b = 0; b < B; ++b) yb[i * B + b] = yp[i];
        nonlin::polynomial poly;
        poly.fit(eng, B, 21, xp, true, yb.data(), 3, statusc.data());
        const double want[4] = {1.1866141861, 0.4466136311, -0.1223204989, 1.0647628218};
        for (int k = 0; k < 4; ++k) {
            const double d = poly.get(k + 1, B - 1) - want[k];
            if (statusc[B - 1] != 0 || d > 5e-11 || d < -5e-11) {
                std::printf("FAIL (polynomial fit) c%d = %.12f\n", k, poly.get(k + 1, B - 1));
                return 1;
            }
        }
        // test_brent_1 / test_newton_1var_1 (tests/nonlin_test_solve.f90:729, :898): sin(x)/x on [1.5, 5] -> pi
        {
            nonlin::fcn1var_helper f1;
            f1.set_fcn("sinx_div_x");
            std::vector<double> l1(B, 1.5), l2(B, 5.0), xr(B), fr(B);
            nonlin::brent_solver brent;
            nonlin::newton_1var_solver newt;
            for (int which = 0; which < 2; ++which) {
                nonlin::value_pair_batch lim{l1.data(), l2.data()};
                if (which == 0) brent.solve(eng, f1, B, xr.data(), lim, fr.data(), ibc.data(), statusc.data());
                else newt.solve(eng, f1, B, xr.data(), lim, fr.data(), ibc.data(), statusc.data());
                const double d = xr[B / 2] - 3.141592653589793;
                if (statusc[B / 2] != 0 || d > 1e-6 || d < -1e-6) {
                    std::printf("FAIL (one-variable solver %d) x = %.12f\n", which, xr[B / 2]);
                    return 1;
                }
            }
        }
        std::printf("Solution: (%.5f, %.5f)\nResidual: (%.3e, %.3e)\nIterations: %d\nFunction Evaluations: %d\nJacobian Evaluations: %d\n",
                    x[0], x[B], f[0], f[B], ib[0].iter_count, ib[0].fcn_count, ib[0].jacobian_count);
    } catch (const nonlin::error& e) {
        std::printf("nonlin error %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
