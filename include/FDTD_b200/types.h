// types.h -- the input structures of the reference, for builds that do not have the reference tree.
//
// north_star: "the same Field and parameter structures from include/Structures.h and FP.h go in".
// When this header is compiled inside the reference tree, define FDTD_B200_USE_REFERENCE_HEADERS and the
// reference's own include/Structures.h (which pulls Constants.h, Enums.h, FP.h) is used verbatim.  Otherwise
// the declarations below provide the same namespaces, names, member order and values
// (reference include/FP.h:3, include/Enums.h:4-7, include/Constants.h:4-12, include/Structures.h:8-38).
#pragma once

#if defined(FDTD_B200_USE_REFERENCE_HEADERS)
#include "Structures.h"
#else

typedef double FP;

namespace FDTD_enums {
enum class Component { EX, EY, EZ, BX, BY, BZ, JX, JY, JZ };
enum class Axis { X, Y, Z };
}  // namespace FDTD_enums

namespace FDTD_const {
const double C = 3e10;
const double R = 1e-12;
const double EPS0 = 1.0;
const double MU0 = EPS0;
const double N = 4.0;
const double PI = 3.14159265358;   // truncated on purpose: part of the numerical spec
}  // namespace FDTD_const

namespace FDTD_struct {
struct SelectedFields {
    FDTD_enums::Component selected_E;
    FDTD_enums::Component selected_B;
};

struct CurrentParameters {
    int period;
    int m;
    FP dt;
    int iterations;
    FP period_x = static_cast<FP>(m) * FDTD_const::C;
    FP period_y = static_cast<FP>(m) * FDTD_const::C;
    FP period_z = static_cast<FP>(m) * FDTD_const::C;
};

struct Parameters {
    int Ni, Nj, Nk;
    FP ax, bx, ay, by, az, bz;
    FP dx, dy, dz;
};
}  // namespace FDTD_struct

#endif  // FDTD_B200_USE_REFERENCE_HEADERS
