// FDTD_PML.h -- FDTD_b200::FDTD_PML, drop-in for FDTD_openmp::FDTD_PML (reference include/FDTD/FDTD_PML.h:10-45)
// and FDTD_kokkos::FDTD_PML (include/FDTD_kokkos/FDTD_PML_kokkos.h:10-34).
#pragma once

#include "FDTD.h"

namespace FDTD_b200 {

class FDTD_PML : public FDTD {
public:
    // FDTD_PML(Parameters, FP dt, FP pml_percent) -- FDTD_PML.h:42; thickness_d = int(N_d * pml_percent)
    FDTD_PML(Parameters _parameters, FP _dt, FP pml_percent) : FDTD(_parameters, _dt, DeferCreate{}) {
        fdtd_config_t cfg;
        fdtd_config_init(&cfg);
        fill_config(cfg);
        cfg.pml_mode = FDTD_PML_PERCENT;
        cfg.pml_percent = pml_percent;
        create(cfg, devices_from_env());
    }
    // extension: explicit device list (one z slab per entry)
    FDTD_PML(Parameters _parameters, FP _dt, FP pml_percent, const std::vector<int>& devices) : FDTD(_parameters, _dt, DeferCreate{}) {
        fdtd_config_t cfg;
        fdtd_config_init(&cfg);
        fill_config(cfg);
        cfg.pml_mode = FDTD_PML_PERCENT;
        cfg.pml_percent = pml_percent;
        create(cfg, devices);
    }
    // extension: explicit per-axis thickness in cells (weak scaling keeps the shell 32 cells thick)
    FDTD_PML(Parameters _parameters, FP _dt, int pml_i, int pml_j, int pml_k) : FDTD(_parameters, _dt, DeferCreate{}) {
        fdtd_config_t cfg;
        fdtd_config_init(&cfg);
        fill_config(cfg);
        cfg.pml_mode = FDTD_PML_THICKNESS;
        cfg.pml_thickness[0] = pml_i; cfg.pml_thickness[1] = pml_j; cfg.pml_thickness[2] = pml_k;
        create(cfg, devices_from_env());
    }

    void update_fields() override { FDTD::update_fields(); }   // the PML shell is a region predicate in the same launches
};

}  // namespace FDTD_b200
