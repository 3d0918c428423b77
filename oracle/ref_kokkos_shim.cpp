// TEST INFRASTRUCTURE ONLY.
// extern "C" shim around the UNMODIFIED reference classes FDTD_kokkos::FDTD (include/FDTD_kokkos/FDTD_kokkos.h:7-37,
// src/FDTD_kokkos/FDTD_kokkos.cpp) and FDTD_kokkos::FDTD_PML (include/FDTD_kokkos/FDTD_PML_kokkos.h), the reference's
// Kokkos-OpenMP path, so that the tests can pin the oracle's distinct-Jx/Jy/Jz mode (kokkos_functors.h:81-89) against
// the real thing and bench.py can time it beside the plain-C++ path.  Compiled by
// oracle/Makefile (target ref_kokkos) from the reference's own sources where they lie under $(REFERENCE), against the
// vendored Kokkos (3rdparty/kokkos, built out of tree with the OpenMP backend); nothing is copied into this repository.
#include <cstddef>
#include <cstdlib>
#include <exception>

#include <Kokkos_Core.hpp>
#include <omp.h>

#include "FDTD_PML_kokkos.h"

using FDTD_kokkos::Component;
using FDTD_kokkos::Parameters;

namespace {
bool g_init = false;
void ensure_kokkos() {
    // process-wide initialisation, like the reference's test_main.cpp:5-8 / kokkos_sample.cpp:154,183
    if (g_init) return;
    Kokkos::InitializationSettings st;
    st.set_num_threads(omp_get_max_threads());
    st.set_disable_warnings(true);
    Kokkos::initialize(st);
    g_init = true;
    std::atexit([] { if (Kokkos::is_initialized() && !Kokkos::is_finalized()) Kokkos::finalize(); });
}
}  // namespace

// (the reference's classes have no virtual destructor: remember the dynamic type for delete)
struct Handle { FDTD_kokkos::FDTD* s; bool pml; };
static FDTD_kokkos::FDTD* S(void* h) { return static_cast<Handle*>(h)->s; }

extern "C" {

void* refk_create(int Ni, int Nj, int Nk, double ax, double bx, double ay, double by, double az, double bz,
                  double dx, double dy, double dz, double dt, double pml_percent) {
    ensure_kokkos();
    Parameters p{Ni, Nj, Nk, ax, bx, ay, by, az, bz, dx, dy, dz};
    try {
        if (pml_percent >= 0.0) return new Handle{new FDTD_kokkos::FDTD_PML(p, dt, pml_percent), true};
        return new Handle{new FDTD_kokkos::FDTD(p, dt), false};
    } catch (const std::exception&) {
        return nullptr;
    }
}

void refk_destroy(void* h) {
    Handle* hd = static_cast<Handle*>(h);
    if (!hd) return;
    if (hd->pml) delete static_cast<FDTD_kokkos::FDTD_PML*>(hd->s); else delete hd->s;
    delete hd;
}

// Pointer to the reference's own storage (get_field returns the View, a ref-counted handle to host memory).
double* refk_field(void* h, int comp) {
    try {
        return S(h)->get_field(static_cast<Component>(comp)).data();
    } catch (const std::exception&) {
        return nullptr;
    }
}

void refk_update_fields(void* h) {
    S(h)->update_fields();
    Kokkos::fence();   // kokkos_sample.cpp:110-112
}

void refk_step(void* h, int n) {
    auto* s = S(h);
    for (int t = 0; t < n; t++) s->update_fields();
    Kokkos::fence();
}

void refk_zeroed_currents(void* h) {
    S(h)->zeroed_currents();
    Kokkos::fence();
}

int refk_threads(void) {
    ensure_kokkos();
    return Kokkos::DefaultExecutionSpace().concurrency();
}

}  // extern "C"
