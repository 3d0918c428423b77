// ./sample_b200 [N iters] -- the reference's perf test (perf-tests/sample/sample.cpp) on FDTD_b200.
// Same argv convention (:150-168), same scenario (spherical wave from a 2x2x2-cell current source for the
// first 40 steps, :14-87), same "Execution time:" line (:88-91) and the same 10x10 Ex slice print at k = N/2
// (:125-134); the PML rerun of the reference's never-defined __PML_TEST__ branch (:93-121) is enabled with a
// third argument "pml".  `--kokkos-slice` prints the YZ slice kokkos_sample.cpp:141-149 prints instead (G10).
// A third positional argument "1" (what python_script_legend/visualization.py:52 passes) or `--dump [dir]` writes
// the k = N/2 slice of all six components after every iteration to <dir>/OutFiles_<1..6>/<iter>.csv, the files
// that script animates (visualization.py:13-23).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <string>

#include "FDTD_b200/FDTD_PML.h"
#include "FDTD_b200/field_dump.h"

using namespace FDTD_b200;

static const char* g_dump_root = nullptr;   // non-null: write the per-iteration slice CSVs

template <class Solver>
static double run_scenario(Solver& method, const Parameters& params, CurrentParameters& cur_param, int it) {
    double T = cur_param.period, Tx = cur_param.period_x, Ty = cur_param.period_y, Tz = cur_param.period_z;
    std::function<double(double, double, double, double)> cur_func = [T, Tx, Ty, Tz](double x, double y, double z, double t) {
        return sin(2.0 * FDTD_const::PI * t / T) * pow(cos(2.0 * FDTD_const::PI * x / Tx), 2.0) *
               pow(cos(2.0 * FDTD_const::PI * y / Ty), 2.0) * pow(cos(2.0 * FDTD_const::PI * z / Tz), 2.0);
    };
    int cur_time = std::min(cur_param.iterations, it);
    int start_i = static_cast<int>(floor((-Tx / 4.0 - params.ax) / params.dx));
    int start_j = static_cast<int>(floor((-Ty / 4.0 - params.ay) / params.dy));
    int start_k = static_cast<int>(floor((-Tz / 4.0 - params.az) / params.dz));
    int max_i = static_cast<int>(floor((Tx / 4.0 - params.ax) / params.dx));
    int max_j = static_cast<int>(floor((Ty / 4.0 - params.ay) / params.dy));
    int max_k = static_cast<int>(floor((Tz / 4.0 - params.az) / params.dz));

    auto start = std::chrono::high_resolution_clock::now();
    for (int t = 0; t < cur_time; t++) {
        for (int k = start_k; k < max_k; ++k)
            for (int j = start_j; j < max_j; ++j)
                for (int i = start_i; i < max_i; ++i) {
                    int index = i + j * params.Ni + k * params.Ni * params.Nj;
                    double value = cur_func(static_cast<double>(i) * params.dx, static_cast<double>(j) * params.dy,
                                            static_cast<double>(k) * params.dz, static_cast<double>(t + 1) * cur_param.dt);
                    method.get_field(Component::JX)[index] = value;
                    method.get_field(Component::JY)[index] = value;
                    method.get_field(Component::JZ)[index] = value;
                }
        method.update_fields();
        if (g_dump_root) write_all_slices_csv(method, t, Axis::Z, -1, g_dump_root);
    }
    method.zeroed_currents();
    for (int t = cur_time; t < it; t++) {
        method.update_fields();
        if (g_dump_root) write_all_slices_csv(method, t, Axis::Z, -1, g_dump_root);
    }
    method.sync();   // the GPU runs asynchronously; the reference's loop is synchronous
    auto end = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(end - start).count();
}

template <class Solver>
static void print_slice(Solver& method, const Parameters& params, bool kokkos_slice) {
    auto& ex = method.get_field(Component::EX);
    if (!kokkos_slice) {
        int k = params.Nk / 2;
        for (int j = params.Nj / 2 - 5; j < params.Nj / 2 + 5; j++) {
            for (int i = params.Ni / 2 - 5; i < params.Ni / 2 + 5; i++) {
                int index = i + j * params.Ni + k * params.Ni * params.Nj;
                std::cout << std::setw(12) << std::fixed << std::setprecision(5) << ex[index];
            }
            std::cout << std::endl;
        }
    } else {
        int i = params.Nk / 2;
        for (int j = params.Nj / 2 - 5; j < params.Nj / 2 + 5; j++) {
            for (int k = params.Ni / 2 - 5; k < params.Ni / 2 + 5; k++) {
                int index = i + j * params.Ni + k * params.Ni * params.Nj;
                std::cout << std::setw(12) << std::fixed << std::setprecision(5) << ex[index];
            }
            std::cout << std::endl;
        }
    }
    std::cout << std::endl;
}

static void spherical_wave(int n, int it, bool with_pml, bool kokkos_slice) {
    CurrentParameters cur_param{8, 4, 0.2};
    cur_param.iterations = static_cast<int>(static_cast<double>(cur_param.period) / cur_param.dt);
    double d = FDTD_const::C;
    double boundary = static_cast<double>(n) / 2.0 * d;
    Parameters params{n, n, n, -boundary, boundary, -boundary, boundary, -boundary, boundary, d, d, d};

    FDTD method(params, cur_param.dt);
    double elapsed = run_scenario(method, params, cur_param, it);
    std::cout << "Execution time: " << elapsed << " s" << std::endl;
    if (with_pml) {
        FDTD_PML pml_method(params, cur_param.dt, 0.2);
        double elapsed_pml = run_scenario(pml_method, params, cur_param, it);
        std::cout << "Execution time (PML): " << elapsed_pml << " s" << std::endl;
        print_slice(method, params, kokkos_slice);
        std::cout << "PML: \n" << std::endl;
        print_slice(pml_method, params, kokkos_slice);
        return;
    }
    print_slice(method, params, kokkos_slice);
}

int main(int argc, char* argv[]) {
    bool with_pml = false, kokkos_slice = false;
    int nargs = 0;
    const char* pos[2] = {nullptr, nullptr};
    for (int a = 1; a < argc; ++a) {
        if (!std::strcmp(argv[a], "pml")) with_pml = true;
        else if (!std::strcmp(argv[a], "--kokkos-slice")) kokkos_slice = true;
        else if (!std::strcmp(argv[a], "--dump")) {
            g_dump_root = (a + 1 < argc && argv[a + 1][0] != '-') ? argv[++a] : ".";
        }
        else if (nargs < 2) pos[nargs++] = argv[a];
        else if (nargs == 2 && (!std::strcmp(argv[a], "1") || !std::strcmp(argv[a], "0"))) {
            if (argv[a][0] == '1' && !g_dump_root) g_dump_root = ".";   // visualization.py:52: [exe, N, iters, "1"]
        }
        else nargs++;
    }
    if (nargs == 0) spherical_wave(32, 100, with_pml, kokkos_slice);
    else if (nargs == 2) spherical_wave(std::atoi(pos[0]), std::atoi(pos[1]), with_pml, kokkos_slice);
    else {
        std::cout << "ERROR: Incorrect number of parameters" << std::endl;
        exit(1);
    }
    return 0;
}
