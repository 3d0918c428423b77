// solver.h -- host-side state of one fdtd_solver_t (one GPU, one z-slab).
#pragma once

#include <cuda.h>   // CUtensorMap (driver entry points are fetched at run time, nothing links against libcuda)
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/fdtd_b200.h"
#include "fdtd_common.cuh"

namespace fdtd_b200 {

struct NcclApi;   // dlopen'ed libnccl entry points (nccl_ring.cu)
struct PeerRing;  // peer-mapped ghost planes + sequence flags (peer_ring.cu)

// flag words of one rank (device memory, written by the neighbours' copy engines)
enum { PEER_F_READY_FROM_DOWN = 0, PEER_F_READY_FROM_UP = 1, PEER_F_DATA_FROM_DOWN = 2, PEER_F_DATA_FROM_UP = 3,
       PEER_F_ERR = 4, PEER_F_STAGE = 8, PEER_STAGE_SLOTS = 8, PEER_FLAG_WORDS = 64 };

// Environment switches, read ONCE when the solver is created (tools/sweep.py, tools/mgpu_probe.py and the tests set
// them before constructing a solver).  Defaults are the measured best on B200 (profiles/).
struct Tunables {
    int fused_variant = -1;   // FDTD_B200_FUSED_VARIANT: one-step fused pass variant (-1: per-dtype default)
    int fused_kc = 0;         // FDTD_B200_FUSED_KC: planes per CTA chunk (0: chosen by the host)
    int t2_variant = 0;       // FDTD_B200_T2_VARIANT
    int t2_strip = 0;         // FDTD_B200_T2_STRIP: tile columns per strip (0: row-major tile ids)
    int st_cs = 1;            // FDTD_B200_ST_CS: streaming output stores (profiles/stcs_r01.jsonl)
    int tma_l2 = 256;         // FDTD_B200_TMA_L2: L2 promotion of the tensor maps (0, 64, 128, 256)
    bool no_tma = false;      // FDTD_B200_NO_TMA
    bool no_t2 = false;       // FDTD_B200_NO_T2
    bool no_lazy = false;     // FDTD_B200_NO_LAZY
    int mgpu_debug = 0;       // FDTD_B200_MGPU_DEBUG (timing experiments, tools/mgpu_probe.py)
    bool pml_t2_f32 = false;  // FDTD_B200_PML_T2_F32
    int halo_timeout_s = 30;  // FDTD_B200_HALO_TIMEOUT_S: how long a pass CTA waits for a neighbour's planes before giving up
};

struct Solver {
    fdtd_config_t cfg{};
    Tunables tun{};
    int device = 0;
    int dtype = FDTD_F64;
    size_t esz = 8;
    bool f32_arith = false;        // FDTD_FLAG_F32_ARITH: float storage AND float arithmetic (not a reference mode)
    Geom g{};
    Coefs c{};
    bool has_pml = false;
    int pml[3] = {0, 0, 0};
    int main_lo[3] = {0, 0, 0}, main_hi[3] = {0, 0, 0};
    double* d_decay[3] = {nullptr, nullptr, nullptr};
    double* d_coef2[3] = {nullptr, nullptr, nullptr};

    // Device arrays.  base[] are the cudaMalloc pointers ((nk + 2*GHOST_PLANES) planes); p[] = base + GHOST_PLANES planes.
    // E and B have two generations (ping-pong) when the fused pass is enabled; cur selects the live one.
    void* base[NCOMP][2] = {};
    void* p[NCOMP][2] = {};
    void* split_base[2 * NSPLIT] = {};
    void* split_p[2 * NSPLIT] = {};
    int cur = 0;
    bool fused = false;
    bool t2 = false;               // fdtd_step(n >= 2) pairs steps into the temporally blocked T2 pass
    // PML solver, two steps per pass: the T2 pass advances the store box sb (main box shrunk by 2 on the axes that have
    // a shell), the sweep kernels advance everything outside it twice (csrc/fdtd_capi.cu::advance_pml_pair)
    bool pml_t2 = false;
    int sb_lo[3] = {0, 0, 0}, sb_hi[3] = {0, 0, 0};
    bool j_stale = false;          // a T2 pass evaluated the device source in-kernel: the J arrays lag one step
    int64_t device_bytes = 0;

    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;   // halo exchange (highest priority: NCCL's CTAs take the first SMs that free up)
    cudaStream_t bnd_stream = nullptr;    // boundary slabs of an overlapped pass (fill the interior launch's tail wave)
    cudaStream_t launch_stream = nullptr; // where the pass kernels of the current launch group go (stream or bnd_stream)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;

    // Deferred trailing B half step: after update_fields() the device holds E(n+1) and B(n+1/2);
    // the missing half step is merged into the next step's leading half step (B = (B+h)+h) or applied
    // by flush() before anything reads or overwrites E/B.  Bit-identical either way.
    bool b_pending = false;
    // Lazy pairing of single update_fields() calls: an odd call is only recorded; the next call runs both as one
    // two-step pass, and every other entry point that reads or changes solver state runs the recorded step first.
    int lazy_steps = 0;
    bool ghosts_e_valid = false;   // top ghost planes of Ex, Ey hold the upper neighbour's current plane
    bool ghosts_b_valid = false;   // bottom ghost planes of Bx, By hold the lower neighbour's current plane
    bool ghosts_fused_valid = false;
    bool ghosts_t2_valid = false;

    JBox jbox{};                   // where J may be non-zero
    // device-resident source (fdtd_set_source)
    bool src_active = false;
    int src_lo[3] = {}, src_hi[3] = {};
    double* d_w[3] = {nullptr, nullptr, nullptr};
    std::vector<double> src_amp;
    int src_t = 0;

    // Pending J writes: host writes of J (fdtd_scatter) that arrive while one update_fields() call is recorded but not yet
    // issued belong to the step AFTER the recorded one.  They are kept as a small dense box per component so that the
    // next update_fields() can still issue both steps as one two-step pass (stage A reads the J arrays, stage B the box);
    // the box is written into the arrays right after.  Reference-style loops (`J[idx] = v; update_fields();`,
    // perf-tests/sample/sample.cpp:66-87) therefore run at the speed of fdtd_step(n).
    bool jpend = false;
    int jp_lo[3] = {}, jp_hi[3] = {};
    void* d_jpend = nullptr;       // 3 components x JPEND_MAX_CELLS elements
    static constexpr int JPEND_MAX_CELLS = 4096;

    // staging for scatter / gather
    void* h_stage = nullptr;       // pinned
    void* d_stage = nullptr;
    size_t stage_bytes = 0;

    // NCCL ring (bootstrap + fallback transport) and the copy-engine ring over peer-mapped memory (default transport)
    NcclApi* nccl = nullptr;
    void* comm = nullptr;
    PeerRing* peer = nullptr;
    std::string peer_note;         // why the peer transport is not in use (diagnostics)
    bool halo_in_kernel = true;    // T2 pass: the CTAs that read ghost planes wait for the halo themselves (peer transport only)

    // per-solver launch configuration caches (one cudaFuncSetAttribute per kernel and device context)
    unsigned configured = 0;       // bit per kernel family
    // per-pass timeline (fdtd_timeline_enable): events pass start / exchange start / exchange end / pass end
    std::vector<cudaEvent_t> tl_events;
    int tl_cap = 0, tl_n = 0;
    // tensor maps of the six E/B arrays per generation (T2 pass), encoded once
    CUtensorMap tmaps[2][6];
    int tmaps_by[2] = {0, 0};      // box rows the cached maps were encoded for (0: not encoded, -1: encoding failed)

    int64_t launches = 0;
    int64_t steps_done = 0;
    int64_t passes_t2 = 0;

    void* cur_ptr(int comp) const { return p[comp][(comp < JX) ? cur : 0]; }
};

// thread-local error message (fdtd_last_error)
void set_error(const std::string& msg);
fdtd_status_t fail(fdtd_status_t code, const std::string& msg);
fdtd_status_t cuda_fail(cudaError_t e, const char* what);

#define FDTD_CUDA_TRY(expr)                                                     \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return ::fdtd_b200::cuda_fail(_e, #expr);        \
    } while (0)

// nccl_ring.cu
fdtd_status_t nccl_unique_id(void* out, size_t cap);
fdtd_status_t nccl_init(Solver* s, const void* id, size_t bytes);
void nccl_destroy(Solver* s);
// Grouped ring exchange of whole planes.  Each entry: send `bytes` from `send` to `peer_send`,
// receive `bytes` into `recv` from `peer_recv`.
// (push view of the same entry, used by the peer transport: the planes go to the upper / lower neighbour's array
// (comp, gen), plane `dst_plane` in the RECEIVER's numbering: -2, -1 = its bottom ghosts, 0, 1 = its top ghosts
// nk_peer, nk_peer + 1)
struct PlaneXfer {
    const void* send; int peer_send; void* recv; int peer_recv; size_t bytes;
    int to_upper, comp, gen, dst_plane;
};
fdtd_status_t nccl_exchange(Solver* s, const PlaneXfer* x, int n, cudaStream_t stream);

// peer_ring.cu
struct PeerBootstrap {
    void* ctx;
    // send `bytes` from `send` to ranks up and down; receive from `down` into from_down and from `up` into from_up
    fdtd_status_t (*exchange)(void* ctx, const void* send, int up, int down, void* from_down, int down2, void* from_up, int up2, size_t bytes);
    fdtd_status_t (*all_min)(void* ctx, int* value);   // min over all ranks
};
fdtd_status_t peer_ring_init_ipc(Solver* s, const PeerBootstrap& boot);
fdtd_status_t peer_ring_init_local(Solver** all, int n);
void peer_ring_destroy(Solver* s);
fdtd_status_t peer_exchange(Solver* s, const PlaneXfer* x, int n, cudaStream_t q, bool wait_data, cudaEvent_t ev_start, cudaEvent_t ev_end);
fdtd_status_t peer_wait_data(Solver* s, cudaStream_t q);
const unsigned* peer_data_flags(const Solver* s);   // [0] from the lower neighbour, [1] from the upper neighbour
unsigned* peer_error_word(const Solver* s);
unsigned peer_last_seq(const Solver* s);

}  // namespace fdtd_b200
