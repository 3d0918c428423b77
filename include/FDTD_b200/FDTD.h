// FDTD.h -- namespace FDTD_b200: the reference's solver class API on top of libfdtd_b200.so.
//
// Drop-in for FDTD_openmp::FDTD (reference include/FDTD/FDTD.h:8-41) and FDTD_kokkos::FDTD
// (include/FDTD_kokkos/FDTD_kokkos.h:7-36): same constructor, get_field / update_fields / zeroed_currents,
// same exceptions.  The reference follows "one namespace + one fixture copy per backend"; this is the
// third backend.  Header-only: it only needs include/fdtd_b200.h and -lfdtd_b200.
//
// Several GPUs behind the same interface.  The reference's caller has no notion of ranks (FDTD.h:35-40), so neither has
// this class: `FDTD(params, dt)` spreads the grid over FDTD_B200_GPUS devices (environment, default 1; or the
// device-list constructor) as z slabs (coarray/fdtd.F90:149-164), one fdtd_solver_t per GPU in THIS process, linked by
// fdtd_comm_init_local (copy engines push the halo planes into peer-mapped ghost planes; no NCCL, no second process).
// Every call fans out over the slab solvers from the caller's thread; get_field reads gather the slabs into the one
// host Field, writes are routed by plane.  Rule kept inside the class: before any call that waits for one slab, all
// slabs issue their recorded work (fdtd_issue / fdtd_flush) -- a pass only completes once its neighbours have issued theirs.
//
// Field coherence (the reference returns a mutable `Field&` into host memory; here the data lives in HBM):
//   * `Field::operator[]` returns a proxy.  Reading it fetches the component from the device once (lazy
//     dense download into a host mirror) and serves further reads from the mirror until the next step.
//   * Writing logs (index, value) pairs; the log is scattered to the device right before the next
//     update_fields() -- the 24-writes-per-step source loop of perf-tests/sample/sample.cpp:66-81 costs one
//     tiny H2D copy, not a 3 GiB upload.  A log that grows past 1/16 of the grid turns into one dense upload
//     (Test_FDTD::initial_filling, src/FDTD/test_FDTD.cpp:5-51, writes every cell).
//   * `Field` is a shallow, reference-counted handle like the Kokkos flavour's View (kokkos_shared.h:16): both caller
//     styles work -- `Field& f = s.get_field(c)` (sample.cpp:76) and `auto f = s.get_field(c)` held across steps
//     (kokkos_sample.cpp:82-84) address the same storage.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../fdtd_b200.h"
#include "types.h"

namespace FDTD_b200 {

using namespace FDTD_enums;
using namespace FDTD_struct;

class FDTD;

namespace detail {

inline void check(fdtd_status_t st) {
    if (st == FDTD_OK) return;
    const std::string msg = fdtd_last_error();
    if (st == FDTD_ERR_INVALID_PARAMETERS) throw std::invalid_argument(msg);
    if (st == FDTD_ERR_INVALID_COMPONENT) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

// The slab solvers behind one FDTD object (one entry on a single GPU).
struct Ring {
    std::vector<fdtd_solver_t*> h;
    std::vector<int> kb, ke;           // slab r owns global planes [kb[r], ke[r])
    std::size_t plane = 0;             // Ni * Nj
    int Ni = 0, Nj = 0, Nk = 0;

    ~Ring() {
        for (fdtd_solver_t* s : h) if (s) fdtd_flush(s);   // nobody may be left waiting for a neighbour
        for (fdtd_solver_t* s : h) if (s) fdtd_destroy(s);
    }
    bool is_B(int comp) const { return comp >= 3 && comp <= 5; }
    // Before a call that waits for the device: every slab issues its recorded step.  Accesses that make the library apply
    // the deferred B half step -- any access to B, and WRITES of E -- get it issued on every slab here: it needs the Ex, Ey
    // ring exchange, and left to the per-slab call it would be issued on one slab and waited for at once (a deadlock the
    // random call-sequence test found on E writes).
    void prep(int comp, bool write = false) {
        if (h.size() == 1) return;
        const bool collective = is_B(comp) || (write && comp < 6);
        for (fdtd_solver_t* s : h) check(collective ? fdtd_flush(s) : fdtd_issue(s));
    }
    void download(int comp, FP* host) {
        prep(comp);
        for (std::size_t r = 0; r < h.size(); ++r)
            check(fdtd_download(h[r], comp, host + (std::size_t)kb[r] * plane, (std::size_t)(ke[r] - kb[r]) * plane));
    }
    void upload(int comp, const FP* host) {
        prep(comp, true);
        for (std::size_t r = 0; r < h.size(); ++r)
            check(fdtd_upload(h[r], comp, host + (std::size_t)kb[r] * plane, (std::size_t)(ke[r] - kb[r]) * plane));
    }
    // global flat indices: every slab gets the whole list (entries of other slabs are skipped on the device; the
    // bounding box of the non-zero currents must agree on all slabs, fdtd_b200.h "COLLECTIVE CALLS")
    void scatter(int comp, const int64_t* idx, const FP* val, std::size_t n) {
        prep(comp, true);
        for (fdtd_solver_t* s : h) check(fdtd_scatter(s, comp, idx, val, n));
    }
    void gather(int comp, const int64_t* idx, FP* val, std::size_t n) {
        prep(comp);
        for (fdtd_solver_t* s : h) check(fdtd_gather(s, comp, idx, val, n));   // each slab fills the entries it owns
    }
};

struct FieldState {
    std::shared_ptr<Ring> ring;
    int comp = 0;
    std::size_t n = 0;
    std::vector<FP> host;
    bool host_valid = false;
    bool dense_dirty = false;
    static constexpr std::size_t kSparseReadLimit = 256;
    std::size_t sparse_reads = 0;
    std::vector<int64_t> log_idx;
    std::vector<FP> log_val;

    void ensure_host() {
        if (host_valid) return;
        flush();                                     // pending writes must land before the device copy is read
        host.resize(n);
        ring->download(comp, host.data());
        host_valid = true;
    }
    FP read(std::size_t i) {
        // A handful of probe reads after a step (the 10x10 slice of sample.cpp:125-134) are served by sparse
        // device gathers; a caller that keeps reading (Test_FDTD::get_max_abs_error walks a whole line or more)
        // gets one dense download instead.
        if (!host_valid && sparse_reads < kSparseReadLimit) {
            flush();
            ++sparse_reads;
            const int64_t idx = static_cast<int64_t>(i);
            FP v = 0;
            ring->gather(comp, &idx, &v, 1);
            return v;
        }
        ensure_host();
        return host[i];
    }
    void write(std::size_t i, FP v) {
        if (host_valid) host[i] = v;
        if (dense_dirty) return;
        log_idx.push_back(static_cast<int64_t>(i));
        log_val.push_back(v);
        if (log_idx.size() > n / 16 + 1024) ensure_host_for_dense();   // too many sparse writes: go dense
    }
    void ensure_host_for_dense() {
        if (!host_valid) {
            // fetch the device copy, then replay the log on top of it
            std::vector<int64_t> li; std::vector<FP> lv;
            li.swap(log_idx); lv.swap(log_val);
            host.resize(n);
            ring->download(comp, host.data());
            host_valid = true;
            for (std::size_t t = 0; t < li.size(); ++t) host[static_cast<std::size_t>(li[t])] = lv[t];
        } else {
            log_idx.clear(); log_val.clear();
        }
        dense_dirty = true;
    }
    // Push host-side writes to the device (called before a step / a device-side change of this component).
    void flush() {
        if (dense_dirty) {
            ring->upload(comp, host.data());
            dense_dirty = false;
        } else if (!log_idx.empty()) {
            ring->scatter(comp, log_idx.data(), log_val.data(), log_idx.size());
        }
        log_idx.clear(); log_val.clear();
    }
    void invalidate() { host_valid = false; sparse_reads = 0; }
    void forget_writes() { log_idx.clear(); log_val.clear(); dense_dirty = false; invalidate(); }
};

}  // namespace detail

class Field {
public:
    class Ref {
    public:
        Ref(detail::FieldState& f, std::size_t i) : f_(f), i_(i) {}
        operator FP() const { return f_.read(i_); }
        Ref& operator=(FP v) { f_.write(i_, v); return *this; }
        Ref& operator=(const Ref& o) { f_.write(i_, static_cast<FP>(o)); return *this; }
        Ref& operator+=(FP v) { f_.write(i_, f_.read(i_) + v); return *this; }
        Ref& operator-=(FP v) { f_.write(i_, f_.read(i_) - v); return *this; }
        Ref& operator*=(FP v) { f_.write(i_, f_.read(i_) * v); return *this; }
    private:
        detail::FieldState& f_;
        std::size_t i_;
    };

    Field() = default;   // an empty handle, like a default-constructed View

    Ref operator[](std::size_t i) { return Ref(*s_, i); }
    FP operator[](std::size_t i) const { return s_->read(i); }
    Ref operator()(std::size_t i) const { return Ref(*s_, i); }      // Kokkos::View style (test_FDTD_kokkos.cpp:24): a View's
                                                                     // operator() is const and returns a mutable reference
    std::size_t size() const { return s_ ? s_->n : 0; }

    // Dense host copy, current as of the last completed step (downloads if stale).
    const std::vector<FP>& host() { s_->ensure_host(); return s_->host; }

private:
    friend class FDTD;
    std::shared_ptr<detail::FieldState> s_;
};

class FDTD {
public:
    // FDTD(Parameters, double dt) -- include/FDTD/FDTD.h:36; throws std::invalid_argument like FDTD.cpp:5-7.
    // FDTD_B200_GPUS=n in the environment spreads the grid over devices 0..n-1.
    FDTD(Parameters _parameters, double _dt) : parameters(_parameters), dt(_dt) {
        fdtd_config_t cfg;
        fdtd_config_init(&cfg);
        fill_config(cfg);
        create(cfg, devices_from_env());
    }
    // extension: explicit device list (one z slab per entry)
    FDTD(Parameters _parameters, double _dt, const std::vector<int>& devices) : parameters(_parameters), dt(_dt) {
        fdtd_config_t cfg;
        fdtd_config_init(&cfg);
        fill_config(cfg);
        create(cfg, devices);
    }
    virtual ~FDTD() = default;          // the slab solvers die with the last handle that refers to them
    FDTD(const FDTD&) = delete;
    FDTD& operator=(const FDTD&) = delete;

    // Field& get_field(Component) -- FDTD.cpp:138-151; throws std::logic_error on an invalid component
    Field& get_field(Component this_field) {
        const int c = static_cast<int>(this_field);
        if (c < 0 || c > static_cast<int>(Component::JZ)) throw std::logic_error("ERROR: Invalid field component");
        return fields_[c];
    }

    // FDTD.cpp:153-157 / FDTD_PML.cpp:343-365 -- one launch of the fused pass (or two sweeps) per GPU
    virtual void update_fields() {
        for (int c = 0; c < 9; ++c) fields_[c].s_->flush();
        for (fdtd_solver_t* s : ring_->h) detail::check(fdtd_update_fields(s));
        for (int c = 0; c < 6; ++c) fields_[c].s_->invalidate();   // E and B changed on the device; J did not
    }

    // FDTD.cpp:132-136
    void zeroed_currents() {
        for (int c = 6; c < 9; ++c) fields_[c].s_->forget_writes();
        ring_->prep(6);
        for (fdtd_solver_t* s : ring_->h) detail::check(fdtd_zeroed_currents(s));
    }

    // extensions
    void step(int n) {
        for (int c = 0; c < 9; ++c) fields_[c].s_->flush();
        for (fdtd_solver_t* s : ring_->h) detail::check(fdtd_step(s, n));
        for (int c = 0; c < 6; ++c) fields_[c].s_->invalidate();
    }
    void sync() {                                               // Kokkos::fence() equivalent
        for (fdtd_solver_t* s : ring_->h) detail::check(fdtd_flush(s));
        for (fdtd_solver_t* s : ring_->h) detail::check(fdtd_sync(s));
    }
    // Dense 2-D slice at a fixed coordinate along `axis` (0 = i, 1 = j, 2 = k), extracted on the device
    // (row-major [n1][n0]: axis 2 -> Nj x Ni, axis 1 -> Nk x Ni, axis 0 -> Nk x Nj).
    std::vector<FP> read_slice(Component this_field, Axis axis, int index) {
        const int c = static_cast<int>(this_field), a = static_cast<int>(axis);
        fields_[c].s_->flush();
        const std::size_t n0 = (a == 0) ? parameters.Nj : parameters.Ni, n1 = (a == 2) ? parameters.Nj : parameters.Nk;
        std::vector<FP> out(n0 * n1);
        std::size_t total = 0;
        ring_->prep(c);
        for (std::size_t r = 0; r < ring_->h.size(); ++r) {
            // axis 2: the slab that owns the plane returns it, the others take part in the (collective) call and copy
            // nothing; axes 0 / 1: every slab returns its rows k_begin .. k_end
            FP* dst = (a == 2) ? out.data() : out.data() + (std::size_t)ring_->kb[r] * n0;
            std::size_t got = 0;
            detail::check(fdtd_read_slice(ring_->h[r], c, a, index, dst, out.size() - (std::size_t)(dst - out.data()), &got));
            total += got;
        }
        out.resize(total);
        return out;
    }
    const Parameters& get_parameters() const { return parameters; }
    fdtd_solver_t* handle() { return ring_->h[0]; }
    const std::vector<fdtd_solver_t*>& handles() const { return ring_->h; }
    int n_gpus() const { return static_cast<int>(ring_->h.size()); }

protected:
    struct DeferCreate {};
    FDTD(Parameters _parameters, double _dt, DeferCreate) : parameters(_parameters), dt(_dt) {}

    static std::vector<int> devices_from_env() {
        const char* e = std::getenv("FDTD_B200_GPUS");
        const int n = e ? std::atoi(e) : 1;
        std::vector<int> d;
        if (n <= 1) { d.push_back(-1); return d; }      // -1: the current device
        for (int i = 0; i < n; ++i) d.push_back(i);
        return d;
    }

    void fill_config(fdtd_config_t& cfg) const {
        cfg.grid.Ni = parameters.Ni; cfg.grid.Nj = parameters.Nj; cfg.grid.Nk = parameters.Nk;
        cfg.grid.ax = parameters.ax; cfg.grid.bx = parameters.bx;
        cfg.grid.ay = parameters.ay; cfg.grid.by = parameters.by;
        cfg.grid.az = parameters.az; cfg.grid.bz = parameters.bz;
        cfg.grid.dx = parameters.dx; cfg.grid.dy = parameters.dy; cfg.grid.dz = parameters.dz;
        cfg.dt = dt;
        cfg.dtype = (sizeof(FP) == 4) ? FDTD_F32 : FDTD_F64;
    }

    void create(fdtd_config_t cfg, const std::vector<int>& devices) {
        if (devices.empty()) throw std::invalid_argument("ERROR: invalid parameters (empty device list)");
        ring_ = std::make_shared<detail::Ring>();
        const int n = static_cast<int>(devices.size());
        ring_->Ni = parameters.Ni; ring_->Nj = parameters.Nj; ring_->Nk = parameters.Nk;
        ring_->plane = static_cast<std::size_t>(parameters.Ni > 0 ? parameters.Ni : 0) * (parameters.Nj > 0 ? parameters.Nj : 0);
        for (int r = 0; r < n; ++r) {
            cfg.device = devices[r]; cfg.rank = r; cfg.nranks = n;
            fdtd_solver_t* s = nullptr;
            detail::check(fdtd_create_ex(&cfg, &s));
            ring_->h.push_back(s);
            fdtd_info_t info;
            detail::check(fdtd_get_info(s, &info));     // (PML solvers use cost-weighted slab heights)
            ring_->kb.push_back(info.k_begin); ring_->ke.push_back(info.k_end);
        }
        if (n > 1) detail::check(fdtd_comm_init_local(ring_->h.data(), n));
        const std::size_t cells = ring_->plane * static_cast<std::size_t>(parameters.Nk);
        for (int c = 0; c < 9; ++c) {
            fields_[c].s_ = std::make_shared<detail::FieldState>();
            fields_[c].s_->ring = ring_;
            fields_[c].s_->comp = c;
            fields_[c].s_->n = cells;
        }
    }

    Parameters parameters;
    double dt;
    std::shared_ptr<detail::Ring> ring_;
    Field fields_[9];
};

}  // namespace FDTD_b200
