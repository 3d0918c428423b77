// TEST INFRASTRUCTURE ONLY.
// extern "C" shim around the UNMODIFIED reference classes FDTD_openmp::FDTD
// (include/FDTD/FDTD.h:8-41) and FDTD_openmp::FDTD_PML (include/FDTD/FDTD_PML.h:10-45)
// so that Python tests / bench.py --impl reference can drive the real reference
// through ctypes.  Compiled by oracle/Makefile together with the reference's own
// src/FDTD/FDTD.cpp and src/FDTD/FDTD_PML.cpp, read where they lie under
// $(REFERENCE); nothing from the reference is copied into this repository.
#include <cstddef>
#include <exception>
#include <omp.h>

#include "FDTD_PML.h"

using FDTD_openmp::Component;
using FDTD_openmp::Parameters;

extern "C" {

void* ref_create(int Ni, int Nj, int Nk, double ax, double bx, double ay, double by, double az,
                 double bz, double dx, double dy, double dz, double dt, double pml_percent) {
    Parameters p{Ni, Nj, Nk, ax, bx, ay, by, az, bz, dx, dy, dz};
    try {
        if (pml_percent >= 0.0) return new FDTD_openmp::FDTD_PML(p, dt, pml_percent);
        return new FDTD_openmp::FDTD(p, dt);
    } catch (const std::exception&) {
        return nullptr;
    }
}

void ref_destroy(void* h) { delete static_cast<FDTD_openmp::FDTD*>(h); }

// Pointer to the reference's own storage (get_field returns a mutable Field&).
FP* ref_field(void* h, int comp) {
    try {
        return static_cast<FDTD_openmp::FDTD*>(h)->get_field(static_cast<Component>(comp)).data();
    } catch (const std::exception&) {
        return nullptr;
    }
}

size_t ref_field_size(void* h, int comp) {
    return static_cast<FDTD_openmp::FDTD*>(h)->get_field(static_cast<Component>(comp)).size();
}

void ref_update_fields(void* h) { static_cast<FDTD_openmp::FDTD*>(h)->update_fields(); }

void ref_step(void* h, int n) {
    auto* s = static_cast<FDTD_openmp::FDTD*>(h);
    for (int t = 0; t < n; t++) s->update_fields();
}

void ref_zeroed_currents(void* h) { static_cast<FDTD_openmp::FDTD*>(h)->zeroed_currents(); }

int ref_sizeof_fp(void) { return (int)sizeof(FP); }
int ref_max_threads(void) { return omp_get_max_threads(); }
void ref_set_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
