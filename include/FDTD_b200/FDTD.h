// FDTD.h -- namespace FDTD_b200: the reference's solver class API on top of libfdtd_b200.so.
//
// Drop-in for FDTD_openmp::FDTD (reference include/FDTD/FDTD.h:8-41) and FDTD_kokkos::FDTD
// (include/FDTD_kokkos/FDTD_kokkos.h:7-36): same constructor, get_field / update_fields / zeroed_currents,
// same exceptions.  The reference follows "one namespace + one fixture copy per backend"; this is the
// third backend.  Header-only: it only needs include/fdtd_b200.h and -lfdtd_b200.
//
// Field coherence (the reference returns a mutable `Field&` into host memory; here the data lives in HBM):
//   * `Field::operator[]` returns a proxy.  Reading it fetches the component from the device once (lazy
//     dense download into a host mirror) and serves further reads from the mirror until the next step.
//   * Writing logs (index, value) pairs; the log is scattered to the device right before the next
//     update_fields() -- the 24-writes-per-step source loop of perf-tests/sample/sample.cpp:66-81 costs one
//     tiny H2D copy, not a 3 GiB upload.  A log that grows past 1/16 of the grid turns into one dense upload
//     (Test_FDTD::initial_filling, src/FDTD/test_FDTD.cpp:5-51, writes every cell).
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../fdtd_b200.h"
#include "types.h"

namespace FDTD_b200 {

using namespace FDTD_enums;
using namespace FDTD_struct;

class FDTD;

class Field {
public:
    class Ref {
    public:
        Ref(Field& f, std::size_t i) : f_(f), i_(i) {}
        operator FP() const { return f_.read(i_); }
        Ref& operator=(FP v) { f_.write(i_, v); return *this; }
        Ref& operator=(const Ref& o) { f_.write(i_, static_cast<FP>(o)); return *this; }
        Ref& operator+=(FP v) { f_.write(i_, f_.read(i_) + v); return *this; }
        Ref& operator-=(FP v) { f_.write(i_, f_.read(i_) - v); return *this; }
        Ref& operator*=(FP v) { f_.write(i_, f_.read(i_) * v); return *this; }
    private:
        Field& f_;
        std::size_t i_;
    };

    Ref operator[](std::size_t i) { return Ref(*this, i); }
    FP operator[](std::size_t i) const { return const_cast<Field*>(this)->read(i); }
    Ref operator()(std::size_t i) { return Ref(*this, i); }          // Kokkos::View style (test_FDTD_kokkos.cpp:24)
    FP operator()(std::size_t i) const { return const_cast<Field*>(this)->read(i); }
    std::size_t size() const { return n_; }

    // Dense host copy, current as of the last completed step (downloads if stale).
    const std::vector<FP>& host() { ensure_host(); return host_; }

private:
    friend class FDTD;
    Field() = default;
    Field(const Field&) = delete;
    Field& operator=(const Field&) = delete;

    void bind(fdtd_solver_t* h, int comp, std::size_t n) { h_ = h; comp_ = comp; n_ = n; }

    static void check(fdtd_status_t st) {
        if (st == FDTD_OK) return;
        const std::string msg = fdtd_last_error();
        if (st == FDTD_ERR_INVALID_PARAMETERS) throw std::invalid_argument(msg);
        if (st == FDTD_ERR_INVALID_COMPONENT) throw std::logic_error(msg);
        throw std::runtime_error(msg);
    }

    void ensure_host() {
        if (host_valid_) return;
        flush();                                     // pending writes must land before the device copy is read
        host_.resize(n_);
        check(fdtd_download(h_, comp_, host_.data(), n_));
        host_valid_ = true;
    }

    FP read(std::size_t i) {
        // A handful of probe reads after a step (the 10x10 slice of sample.cpp:125-134) are served by sparse
        // device gathers; a caller that keeps reading (Test_FDTD::get_max_abs_error walks a whole line or more)
        // gets one dense download instead.
        if (!host_valid_ && sparse_reads_ < kSparseReadLimit) {
            flush();
            ++sparse_reads_;
            const int64_t idx = static_cast<int64_t>(i);
            FP v = 0;
            check(fdtd_gather(h_, comp_, &idx, &v, 1));
            return v;
        }
        ensure_host();
        return host_[i];
    }

    void write(std::size_t i, FP v) {
        if (host_valid_) host_[i] = v;
        if (dense_dirty_) return;
        log_idx_.push_back(static_cast<int64_t>(i));
        log_val_.push_back(v);
        if (log_idx_.size() > n_ / 16 + 1024) {      // too many sparse writes: go dense
            ensure_host_for_dense();
        }
    }

    void ensure_host_for_dense() {
        if (!host_valid_) {
            // fetch the device copy, then replay the log on top of it
            std::vector<int64_t> li; std::vector<FP> lv;
            li.swap(log_idx_); lv.swap(log_val_);
            host_.resize(n_);
            check(fdtd_download(h_, comp_, host_.data(), n_));
            host_valid_ = true;
            for (std::size_t t = 0; t < li.size(); ++t) host_[static_cast<std::size_t>(li[t])] = lv[t];
        } else {
            log_idx_.clear(); log_val_.clear();
        }
        dense_dirty_ = true;
    }

    // Push host-side writes to the device (called before a step / a device-side change of this component).
    void flush() {
        if (dense_dirty_) {
            check(fdtd_upload(h_, comp_, host_.data(), n_));
            dense_dirty_ = false;
        } else if (!log_idx_.empty()) {
            check(fdtd_scatter(h_, comp_, log_idx_.data(), log_val_.data(), log_idx_.size()));
        }
        log_idx_.clear(); log_val_.clear();
    }

    void invalidate() { host_valid_ = false; sparse_reads_ = 0; }

    fdtd_solver_t* h_ = nullptr;
    int comp_ = 0;
    std::size_t n_ = 0;
    std::vector<FP> host_;
    bool host_valid_ = false;
    bool dense_dirty_ = false;
    static constexpr std::size_t kSparseReadLimit = 256;
    std::size_t sparse_reads_ = 0;
    std::vector<int64_t> log_idx_;
    std::vector<FP> log_val_;
};

class FDTD {
public:
    // FDTD(Parameters, double dt) -- include/FDTD/FDTD.h:36; throws std::invalid_argument like FDTD.cpp:5-7
    FDTD(Parameters _parameters, double _dt) : parameters(_parameters), dt(_dt) {
        fdtd_config_t cfg;
        fdtd_config_init(&cfg);
        fill_config(cfg);
        create(cfg);
    }
    virtual ~FDTD() { if (h_) fdtd_destroy(h_); }
    FDTD(const FDTD&) = delete;
    FDTD& operator=(const FDTD&) = delete;

    // Field& get_field(Component) -- FDTD.cpp:138-151; throws std::logic_error on an invalid component
    Field& get_field(Component this_field) {
        const int c = static_cast<int>(this_field);
        if (c < 0 || c > static_cast<int>(Component::JZ)) throw std::logic_error("ERROR: Invalid field component");
        return fields_[c];
    }

    // FDTD.cpp:153-157 / FDTD_PML.cpp:343-365 -- one launch of the fused pass (or two sweeps) on the GPU
    virtual void update_fields() {
        for (int c = 0; c < 9; ++c) fields_[c].flush();
        Field::check(fdtd_update_fields(h_));
        for (int c = 0; c < 6; ++c) fields_[c].invalidate();   // E and B changed on the device; J did not
    }

    // FDTD.cpp:132-136
    void zeroed_currents() {
        for (int c = 6; c < 9; ++c) {
            fields_[c].log_idx_.clear(); fields_[c].log_val_.clear();
            fields_[c].dense_dirty_ = false;
            fields_[c].invalidate();
        }
        Field::check(fdtd_zeroed_currents(h_));
    }

    // extensions
    void step(int n) {
        for (int c = 0; c < 9; ++c) fields_[c].flush();
        Field::check(fdtd_step(h_, n));
        for (int c = 0; c < 6; ++c) fields_[c].invalidate();
    }
    void sync() { Field::check(fdtd_sync(h_)); }          // Kokkos::fence() equivalent
    // Dense 2-D slice at a fixed coordinate along `axis` (0 = i, 1 = j, 2 = k), extracted on the device
    // (row-major [n1][n0]: axis 2 -> Nj x Ni, axis 1 -> Nk x Ni, axis 0 -> Nk x Nj).
    std::vector<FP> read_slice(Component this_field, Axis axis, int index) {
        const int c = static_cast<int>(this_field), a = static_cast<int>(axis);
        fields_[c].flush();
        const std::size_t n0 = (a == 0) ? parameters.Nj : parameters.Ni, n1 = (a == 2) ? parameters.Nj : parameters.Nk;
        std::vector<FP> out(n0 * n1);
        std::size_t got = 0;
        Field::check(fdtd_read_slice(h_, c, a, index, out.data(), out.size(), &got));
        out.resize(got);
        return out;
    }
    const Parameters& get_parameters() const { return parameters; }
    fdtd_solver_t* handle() { return h_; }

protected:
    struct DeferCreate {};
    FDTD(Parameters _parameters, double _dt, DeferCreate) : parameters(_parameters), dt(_dt) {}

    void fill_config(fdtd_config_t& cfg) const {
        cfg.grid.Ni = parameters.Ni; cfg.grid.Nj = parameters.Nj; cfg.grid.Nk = parameters.Nk;
        cfg.grid.ax = parameters.ax; cfg.grid.bx = parameters.bx;
        cfg.grid.ay = parameters.ay; cfg.grid.by = parameters.by;
        cfg.grid.az = parameters.az; cfg.grid.bz = parameters.bz;
        cfg.grid.dx = parameters.dx; cfg.grid.dy = parameters.dy; cfg.grid.dz = parameters.dz;
        cfg.dt = dt;
        cfg.dtype = (sizeof(FP) == 4) ? FDTD_F32 : FDTD_F64;
    }

    void create(const fdtd_config_t& cfg) {
        Field::check(fdtd_create_ex(&cfg, &h_));
        const std::size_t n = static_cast<std::size_t>(parameters.Ni) * parameters.Nj * parameters.Nk;
        for (int c = 0; c < 9; ++c) fields_[c].bind(h_, c, n);
    }

    Parameters parameters;
    double dt;
    fdtd_solver_t* h_ = nullptr;
    Field fields_[9];
};

}  // namespace FDTD_b200
