// field_dump.h -- per-iteration 2-D slice CSVs: the visualisation feed of the reference.
//
// python_script_legend/visualization.py:10-23,34 of the reference animates `OutFiles_<n>/<iter>.csv`
// (';'-separated rows, component numbering Ex=1 ... Bz=6); the reference's samples no longer write those files
// (SURVEY.md section 8 row f3).  Here the slice is extracted on the device (fdtd_read_slice: one small kernel +
// one Ni*Nj D2H copy), so dumping every iteration costs 2 MiB of PCIe traffic per component at 512^2 instead of a
// 1 GiB field download.
#pragma once

#include <sys/stat.h>

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "FDTD.h"

namespace FDTD_b200 {

// Writes <root>/OutFiles_<component+1>/<iteration>.csv with the slice `index` along `axis` of `component`.
// Returns the path.
inline std::string write_slice_csv(FDTD& method, Component component, int iteration, Axis axis = Axis::Z, int index = -1,
                                   const std::string& root = ".") {
    const Parameters& p = method.get_parameters();
    const int a = static_cast<int>(axis);
    if (index < 0) index = ((a == 0) ? p.Ni : (a == 1) ? p.Nj : p.Nk) / 2;
    const std::size_t n0 = (a == 0) ? p.Nj : p.Ni;
    const std::vector<FP> s = method.read_slice(component, axis, index);
    const std::string dir = root + "/OutFiles_" + std::to_string(static_cast<int>(component) + 1);
    ::mkdir(dir.c_str(), 0777);
    const std::string path = dir + "/" + std::to_string(iteration) + ".csv";
    std::FILE* fh = std::fopen(path.c_str(), "w");
    if (!fh) throw std::runtime_error("cannot write " + path);
    for (std::size_t t = 0; t < s.size(); ++t)
        std::fprintf(fh, "%.17g%c", static_cast<double>(s[t]), ((t + 1) % n0 == 0) ? '\n' : ';');
    std::fclose(fh);
    return path;
}

// All six field components of one iteration (Ex..Bz -> OutFiles_1..OutFiles_6).
inline void write_all_slices_csv(FDTD& method, int iteration, Axis axis = Axis::Z, int index = -1, const std::string& root = ".") {
    for (int c = 0; c < 6; ++c) write_slice_csv(method, static_cast<Component>(c), iteration, axis, index, root);
}

}  // namespace FDTD_b200
