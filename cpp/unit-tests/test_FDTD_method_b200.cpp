// The reference's 12 Convergence.* tests (unit-tests/test_FDTD_method.cpp:18-203) against FDTD_b200::FDTD.
// Same scenario, same fixture calls, same assertion (ratio within 4.0 +- 0.1).  googletest is not on the GPU
// box, so a 20-line TEST/ASSERT_NEAR harness stands in for it; the test bodies read like the reference's.
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "FDTD_b200/test_FDTD.h"

using namespace FDTD_b200;

struct Case { const char* suite; const char* name; void (*fn)(bool&); };
static std::vector<Case>& registry() { static std::vector<Case> r; return r; }
struct Registrar { Registrar(const char* s, const char* n, void (*f)(bool&)) { registry().push_back({s, n, f}); } };
#define TEST(suite, name)                                                  \
    static void suite##_##name(bool& failed_);                             \
    static Registrar reg_##suite##_##name(#suite, #name, suite##_##name);  \
    static void suite##_##name(bool& failed_)
#define ASSERT_NEAR(val, ref, tol)                                                                   \
    do {                                                                                             \
        const double v_ = (val);                                                                     \
        std::printf("    value = %.15g\n", v_);                                                      \
        if (!(std::fabs(v_ - (ref)) <= (tol))) { std::printf("    expected %g +- %g\n", (double)(ref), (double)(tol)); failed_ = true; return; } \
    } while (0)

const FP default_time = 5e-13;

std::function<FP(FP, FP[2])> initial_func = [](FP x, FP size[2]) {
    return sin(2.0 * FDTD_const::PI * (x - size[0]) / (size[1] - size[0]));
};
std::function<FP(FP, FP, FP[2])> true_func = [](FP x, FP t, FP size[2]) {
    return sin(2.0 * FDTD_const::PI * (x - size[0] - FDTD_const::C * t) / (size[1] - size[0]));
};

static FP run_once(Component test_field, SelectedFields current_fields, int Ni, int Nj, int Nk, int iters) {
    Parameters params{Ni, Nj, Nk, 0.0, 1.0, 0.0, 2.0, 0.0, 3.0,
                      (1 - 0) / static_cast<FP>(Ni), (2 - 0) / static_cast<FP>(Nj), (3 - 0) / static_cast<FP>(Nk)};
    FP dt = default_time / static_cast<FP>(iters);
    FDTD method(params, dt);
    Test_FDTD test(params);
    test.initial_filling(method, current_fields, iters, initial_func);
    for (int t = 0; t < iters; t++) method.update_fields();
    return test.get_max_abs_error(method.get_field(test_field), test_field, true_func, default_time);
}

FP run_test(Component test_field, SelectedFields current_fields, int Ni, int Nj, int Nk) {
    const FP err_1 = run_once(test_field, current_fields, Ni, Nj, Nk, 16);
    const FP err_2 = run_once(test_field, current_fields, Ni * 2, Nj * 2, Nk * 2, 32);
    std::printf("    err_1 = %.17g  err_2 = %.17g\n", err_1, err_2);
    return err_1 / err_2;
}

#define CONVERGENCE(name, E, B, F, ni, nj, nk)                                      \
    TEST(Convergence_b200, name) {                                                  \
        SelectedFields current_fields{Component::E, Component::B};                  \
        ASSERT_NEAR(run_test(Component::F, current_fields, ni, nj, nk), 4.0, 0.1);  \
    }

CONVERGENCE(x_axis_EY, EY, BZ, EY, 16, 8, 4)
CONVERGENCE(x_axis_BZ, EY, BZ, BZ, 16, 8, 4)
CONVERGENCE(x_axis_EZ, EZ, BY, EZ, 16, 8, 4)
CONVERGENCE(x_axis_BY, EZ, BY, BY, 16, 8, 4)
CONVERGENCE(y_axis_EX, EX, BZ, EX, 8, 16, 4)
CONVERGENCE(y_axis_BZ, EX, BZ, BZ, 8, 16, 4)
CONVERGENCE(y_axis_EZ, EZ, BX, EZ, 8, 16, 4)
CONVERGENCE(y_axis_BX, EZ, BX, BX, 8, 16, 4)
CONVERGENCE(z_axis_EX, EX, BY, EX, 4, 8, 16)
CONVERGENCE(z_axis_BY, EX, BY, BY, 4, 8, 16)
CONVERGENCE(z_axis_EY, EY, BX, EY, 4, 8, 16)
CONVERGENCE(z_axis_BX, EY, BX, BX, 4, 8, 16)

// Not in the reference's suite (exceptions are untested upstream) -- the API edge behaviour A.4.
TEST(Api_b200, invalid_parameters_throw) {
    bool threw = false;
    try { Parameters p{0, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1}; FDTD m(p, 0.1); } catch (const std::invalid_argument&) { threw = true; }
    ASSERT_NEAR(threw ? 1.0 : 0.0, 1.0, 0.0);
}
TEST(Api_b200, invalid_component_throws) {
    Parameters p{4, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1};
    FDTD m(p, 0.1);
    bool threw = false;
    try { m.get_field(static_cast<Component>(42)); } catch (const std::logic_error&) { threw = true; }
    ASSERT_NEAR(threw ? 1.0 : 0.0, 1.0, 0.0);
}
TEST(Api_b200, field_reference_is_coherent) {
    Parameters p{8, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1};
    FDTD m(p, 1e-12);
    Field& ex = m.get_field(Component::EX);
    ex[5] = 2.5;
    ex[6] += 1.0;
    double got = ex[5] + ex[6];
    m.update_fields();                       // pushes the logged writes, steps, invalidates the host mirrors
    (void)static_cast<FP>(m.get_field(Component::BX)[0]);   // lazy download after the step
    ASSERT_NEAR(got, 3.5, 0.0);
}

int main() {
    int failed = 0;
    for (const Case& c : registry()) {
        std::printf("[ RUN      ] %s.%s\n", c.suite, c.name);
        bool f = false;
        try { c.fn(f); } catch (const std::exception& e) { std::printf("    exception: %s\n", e.what()); f = true; }
        std::printf("%s %s.%s\n", f ? "[  FAILED  ]" : "[       OK ]", c.suite, c.name);
        failed += f ? 1 : 0;
    }
    std::printf("%d tests, %d failed\n", (int)registry().size(), failed);
    return failed ? 1 : 0;
}
