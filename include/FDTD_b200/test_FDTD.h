// test_FDTD.h -- plane-wave fixture for the convergence tests, FDTD_b200 backend.
// Same interface and arithmetic as the reference's per-backend fixture (include/FDTD/test_FDTD.h:8-27,
// src/FDTD/test_FDTD.cpp:5-130), written as a header-only template so any solver with the reference's
// class API (get_field(Component)[index]) can be plugged in.
#pragma once

#include <cmath>
#include <functional>
#include <stdexcept>

#include "FDTD.h"

namespace FDTD_b200 {

template <class Solver = FDTD>
class Test_FDTD_T {
public:
    explicit Test_FDTD_T(Parameters p) : parameters(p) {}

    // src/FDTD/test_FDTD.cpp:5-51: E = sign * f(x), B = f(x + d/2) along the axis implied by the (E, B) pair
    void initial_filling(Solver& solver, SelectedFields fields, int /*iters*/, std::function<FP(FP, FP[2])>& init_function) {
        pick_axis(fields.selected_E, fields.selected_B);
        pick_sign(fields.selected_E, fields.selected_B);
        const int N[3] = {parameters.Ni, parameters.Nj, parameters.Nk};
        const FP d[3] = {parameters.dx, parameters.dy, parameters.dz};
        FP box[3][2] = {{parameters.ax, parameters.bx}, {parameters.ay, parameters.by}, {parameters.az, parameters.bz}};
        const int a = static_cast<int>(axis);
        auto& E = solver.get_field(fields.selected_E);
        auto& B = solver.get_field(fields.selected_B);
        const FP half = d[a] / 2.0;
        for (int m = 0; m < N[a]; ++m) {
            const FP x = static_cast<FP>(m) * d[a];
            const FP ve = sign * init_function(x, box[a]);
            const FP vb = init_function(half + x, box[a]);
            int c[3];
            c[a] = m;
            const int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
            for (c[a1] = 0; c[a1] < N[a1]; ++c[a1])
                for (c[a2] = 0; c[a2] < N[a2]; ++c[a2]) {
                    const int index = c[0] + c[1] * parameters.Ni + c[2] * parameters.Ni * parameters.Nj;
                    E[index] = ve;
                    B[index] = vb;
                }
        }
    }

    // src/FDTD/test_FDTD.cpp:89-130: max |sign*F - f_true| along the line through the origin
    template <class FieldT>
    FP get_max_abs_error(FieldT& this_field, Component field, std::function<FP(FP, FP, FP[2])>& true_function, FP time) {
        const int N[3] = {parameters.Ni, parameters.Nj, parameters.Nk};
        const FP d[3] = {parameters.dx, parameters.dy, parameters.dz};
        FP box[3][2] = {{parameters.ax, parameters.bx}, {parameters.ay, parameters.by}, {parameters.az, parameters.bz}};
        const int a = static_cast<int>(axis);
        const int stride[3] = {1, parameters.Ni, parameters.Ni * parameters.Nj};
        FP x = 0.0;
        if (static_cast<int>(field) > static_cast<int>(Component::EZ)) {   // get_shift, test_FDTD.cpp:81-87
            sign = 1.0;
            x = d[a] / 2.0;
        }
        FP worst = 0.0;
        for (int m = 0; m < N[a]; ++m, x += d[a]) {
            const FP err = std::fabs(sign * static_cast<FP>(this_field[m * stride[a]]) - true_function(x, time, box[a]));
            if (err > worst) worst = err;
        }
        return worst;
    }

private:
    Parameters parameters;
    FP sign = 1.0;
    Axis axis = Axis::X;

    static bool is(Component e, Component b, Component E, Component B) { return e == E && b == B; }

    void pick_sign(Component e, Component b) {   // test_FDTD.cpp:53-65
        if (is(e, b, Component::EX, Component::BZ) || is(e, b, Component::EZ, Component::BY) || is(e, b, Component::EY, Component::BX)) sign = -1.0;
        else if (is(e, b, Component::EY, Component::BZ) || is(e, b, Component::EZ, Component::BX) || is(e, b, Component::EX, Component::BY)) sign = 1.0;
        else throw std::logic_error("ERROR: invalid selected fields");
    }
    void pick_axis(Component e, Component b) {   // test_FDTD.cpp:66-80
        if (is(e, b, Component::EY, Component::BZ) || is(e, b, Component::EZ, Component::BY)) axis = Axis::X;
        else if (is(e, b, Component::EX, Component::BZ) || is(e, b, Component::EZ, Component::BX)) axis = Axis::Y;
        else if (is(e, b, Component::EX, Component::BY) || is(e, b, Component::EY, Component::BX)) axis = Axis::Z;
        else throw std::logic_error("ERROR: invalid selected fields");
    }
};

using Test_FDTD = Test_FDTD_T<FDTD>;

}  // namespace FDTD_b200
