// ./kokkos_sample_b200 [N iters] -- the CALLER PATTERN of the reference's Kokkos perf test
// (perf-tests/kokkos_sample/kokkos_sample.cpp) on FDTD_b200, which differs from ./sample in three ways that matter to a
// drop-in backend:
//   * the three current components are fetched ONCE, by value, before the time loop and the handles are held across
//     all steps (kokkos_sample.cpp:82-84: `auto Jx = method.get_field(...)` copies a ref-counted View) -- here Field is
//     the same kind of shallow handle;
//   * elements are addressed with operator() (kokkos_sample.cpp:105-107);
//   * after the source window the box is written with zeros again before EVERY remaining step
//     (kokkos_sample.cpp:114-129), on top of the one zeroed_currents() call, with a fence after each step.
// Output: the "Execution time:" line and the 10x10 Ex slice at fixed i = N/2 over (j, k) that kokkos_sample prints
// (kokkos_sample.cpp:141-149) -- the Kokkos configuration banner has no counterpart here.
// FDTD_B200_GPUS=n runs the same program on n GPUs (z slabs behind the same class).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "FDTD_b200/FDTD.h"

using namespace FDTD_b200;

static void spherical_wave(int n, int it) {
    CurrentParameters cur{8, 4, 0.2};
    cur.iterations = static_cast<int>(static_cast<double>(cur.period) / cur.dt);
    const double T = cur.period, Tx = cur.period_x, Ty = cur.period_y, Tz = cur.period_z;
    const double d = FDTD_const::C, half = static_cast<double>(n) / 2.0 * d;
    Parameters params{n, n, n, -half, half, -half, half, -half, half, d, d, d};

    FDTD method(params, cur.dt);
    const int active = std::min(cur.iterations, it);
    const int lo[3] = {static_cast<int>(floor((-Tx / 4.0 - params.ax) / params.dx)), static_cast<int>(floor((-Ty / 4.0 - params.ay) / params.dy)),
                       static_cast<int>(floor((-Tz / 4.0 - params.az) / params.dz))};
    const int hi[3] = {static_cast<int>(floor((Tx / 4.0 - params.ax) / params.dx)), static_cast<int>(floor((Ty / 4.0 - params.ay) / params.dy)),
                       static_cast<int>(floor((Tz / 4.0 - params.az) / params.dz))};
    const int Ni = params.Ni, Nj = params.Nj;

    // handles taken once, held for the whole run
    auto Jx = method.get_field(Component::JX);
    auto Jy = method.get_field(Component::JY);
    auto Jz = method.get_field(Component::JZ);

    auto for_box = [&](auto&& body) {
        for (int k = lo[2]; k < hi[2]; ++k)
            for (int j = lo[1]; j < hi[1]; ++j)
                for (int i = lo[0]; i < hi[0]; ++i) body(i, j, k, i + j * Ni + k * Ni * Nj);
    };

    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int t = 0; t < active; ++t) {
        const double time = (t + 1) * cur.dt;
        for_box([&](int i, int j, int k, int idx) {
            const double val = sin(2.0 * FDTD_const::PI * time / T) * pow(cos(2.0 * FDTD_const::PI * (i * params.dx) / Tx), 2.0) *
                               pow(cos(2.0 * FDTD_const::PI * (j * params.dy) / Ty), 2.0) * pow(cos(2.0 * FDTD_const::PI * (k * params.dz) / Tz), 2.0);
            Jx(idx) = val;
            Jy(idx) = val;
            Jz(idx) = val;
        });
        method.update_fields();
        method.sync();
    }
    method.zeroed_currents();
    method.sync();
    for (int t = active; t < it; ++t) {
        for_box([&](int, int, int, int idx) { Jx(idx) = 0.0; Jy(idx) = 0.0; Jz(idx) = 0.0; });
        method.update_fields();
        method.sync();
    }
    const std::chrono::duration<double> elapsed = std::chrono::high_resolution_clock::now() - t0;
    std::cout << "Execution time: " << elapsed.count() << " s" << std::endl;

    auto& Ex = method.get_field(Component::EX);
    const std::vector<FP>& Ex_host = Ex.host();   // the create_mirror_view + deep_copy of kokkos_sample.cpp:135-139
    const int i = params.Nk / 2;
    for (int j = params.Nj / 2 - 5; j < params.Nj / 2 + 5; j++) {
        for (int k = params.Ni / 2 - 5; k < params.Ni / 2 + 5; k++)
            std::cout << std::setw(12) << std::fixed << std::setprecision(5) << Ex_host[i + j * params.Ni + k * params.Ni * params.Nj];
        std::cout << std::endl;
    }
    std::cout << std::endl;
}

int main(int argc, char* argv[]) {
    if (argc == 1) spherical_wave(32, 100);
    else if (argc == 3) spherical_wave(std::atoi(argv[1]), std::atoi(argv[2]));
    else {
        std::cout << "ERROR: Incorrect number of parameters" << std::endl;
        return 1;
    }
    return 0;
}
