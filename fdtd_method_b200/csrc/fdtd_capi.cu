// fdtd_capi.cu -- implementation of the C ABI declared in include/fdtd_b200.h.
//
// Host-side control of the hot path.  Mirrors, call for call, what the reference's classes do:
//   fdtd_create*        FDTD::FDTD / FDTD_PML::FDTD_PML    src/FDTD/FDTD.cpp:3-61, src/FDTD/FDTD_PML.cpp:205-341
//   fdtd_update_fields  FDTD::update_fields                src/FDTD/FDTD.cpp:153-157, src/FDTD/FDTD_PML.cpp:343-365
//   fdtd_zeroed_currents FDTD::zeroed_currents             src/FDTD/FDTD.cpp:132-136
//   fdtd_upload/download/scatter/gather  = reads and writes through `Field& get_field(Component)`, FDTD.cpp:138-151
// There is no CPU implementation behind any of these: without a usable CUDA device they fail.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost a function-pointer test unless a profiler is attached

#include "fused_kernel.cuh"
#include "fused_kernel_v2.cuh"
#include "fused_kernel_t2.cuh"
#include "solver.h"
#include "sweep_kernels.cuh"

namespace fdtd_b200 {

// include/Constants.h:6-11 of the reference -- these literals are part of the numerical spec
static const double kC = 3e10;
static const double kR = 1e-12;
static const double kN = 4.0;
static const double kPI = 3.14159265358;

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
fdtd_status_t fail(fdtd_status_t code, const std::string& msg) { g_err = msg; return code; }
fdtd_status_t cuda_fail(cudaError_t e, const char* what) {
    return fail(FDTD_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}

static inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------------
// PML tables (host, libm) -- src/FDTD/FDTD_PML.cpp:3-65, 98-111, 254-256, 293-338
// ------------------------------------------------------------------------------------------------------
static void pml_profile_host(int N, int p, double d, double dt, double* sigma, double* decay, double* coef2) {
    const double cE = (kC * dt) / d;
    double SGm = 0.0;
    if (p > 0) SGm = -(kN + 1.0) / 2.0 * std::log(kR) / (static_cast<double>(p) * d);
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        if (i < p) s = SGm * std::pow(static_cast<double>(p - i) / static_cast<double>(p), kN);
        else if (i >= N - p) s = SGm * std::pow(static_cast<double>(i + 1 + p - N) / static_cast<double>(p), kN);
        const double dec = std::exp(-s * dt * kC);
        if (sigma) sigma[i] = s;
        decay[i] = dec;
        coef2[i] = (s != 0.0) ? (1.0 - dec) / (s * d) : cE;
    }
}

// ------------------------------------------------------------------------------------------------------
// typed helpers
// ------------------------------------------------------------------------------------------------------
template <typename T>
static Fields<T> make_fields(const Solver* s) {
    Fields<T> f;
    for (int c = 0; c < 3; ++c) {
        f.E[c] = static_cast<T*>(s->p[EX + c][s->cur]);
        f.B[c] = static_cast<T*>(s->p[BX + c][s->cur]);
        f.J[c] = static_cast<T*>(s->p[JX + c][0]);
    }
    for (int c = 0; c < NSPLIT; ++c) {
        f.SE[c] = static_cast<T*>(s->split_p[c]);
        f.SB[c] = static_cast<T*>(s->split_p[NSPLIT + c]);
    }
    return f;
}

static PmlDesc make_pml(const Solver* s) {
    PmlDesc p;
    for (int a = 0; a < 3; ++a) {
        p.lo[a] = s->main_lo[a];
        p.hi[a] = s->main_hi[a];
        p.decay[a] = s->d_decay[a];
        p.coef2[a] = s->d_coef2[a];
    }
    return p;
}

static int pick_kc(const Solver* s, int tiles_ij, int nplanes) {
    // enough CTAs for several waves on 148 SMs, chunks long enough to amortise the k prologue
    int kc = 32;
    while (kc > 4 && (long long)tiles_ij * ((nplanes + kc - 1) / kc) < 148 * 8) kc /= 2;
    if (kc > nplanes) kc = nplanes > 0 ? nplanes : 1;
    (void)s;
    return kc;
}

template <typename T, typename A = double>
static fdtd_status_t launch_sweep(Solver* s, bool is_B, int n_half, int do_pml) {
    constexpr int V = VecOf<T>::V;
    SweepArgs<T> a;
    a.g = s->g; a.c = s->c; a.p = make_pml(s); a.f = make_fields<T>(s); a.jbox = s->jbox;
    a.k_lo = 0; a.k_hi = s->g.nk;
    a.n_half = n_half; a.do_pml = do_pml;
    a.j_quirk = (s->cfg.flags & FDTD_FLAG_J_OPENMP_QUIRK) ? 1 : 0;
    a.mode = 0;
    for (int d = 0; d < 3; ++d) { a.ib_lo[d] = 0; a.ib_hi[d] = 0; a.Bout[d] = a.f.B[d]; a.Eout[d] = a.f.E[d]; }
    dim3 block(SWEEP_BX, SWEEP_BY);
    const int gx = (s->g.Ni + SWEEP_BX * V - 1) / (SWEEP_BX * V);
    const int gy = (s->g.Nj + SWEEP_BY - 1) / SWEEP_BY;
    a.kc = pick_kc(s, gx * gy, s->g.nk);
    const int gz = (s->g.nk + a.kc - 1) / a.kc;
    dim3 grid(gx, gy, gz);
    if (!s->has_pml) {
        if (is_B) sweep_B_kernel<T, false, V, A><<<grid, block, 0, s->stream>>>(a);
        else sweep_E_kernel<T, false, V, A><<<grid, block, 0, s->stream>>>(a);
        FDTD_CUDA_TRY(cudaGetLastError());
        s->launches++;
        return FDTD_OK;
    }
    // PML solver: the interior (main box, i bounds aligned inward to the vector width) runs the lean PML=false
    // instantiation, the shell (and the unaligned fringe of the main box) the PML=true one.  Disjoint cell sets,
    // same arithmetic per cell as the single launch (order inside a phase is irrelevant).
    for (int d = 0; d < 3; ++d) { a.ib_lo[d] = s->main_lo[d]; a.ib_hi[d] = s->main_hi[d]; }
    a.ib_lo[0] = (a.ib_lo[0] + V - 1) / V * V;
    a.ib_hi[0] = a.ib_hi[0] / V * V;
    const long long inner = (long long)(a.ib_hi[0] - a.ib_lo[0]) * (a.ib_hi[1] - a.ib_lo[1]) * (a.ib_hi[2] - a.ib_lo[2]);
    const bool split = a.ib_hi[0] > a.ib_lo[0] && a.ib_hi[1] > a.ib_lo[1] && a.ib_hi[2] > a.ib_lo[2] &&
                       inner >= (1 << 15) && !(s->cfg.flags & FDTD_FLAG_NO_PML_SPLIT);
    if (split) {
        a.mode = 1;
        if (is_B) sweep_B_kernel<T, false, V, A><<<grid, block, 0, s->stream>>>(a);
        else sweep_E_kernel<T, false, V, A><<<grid, block, 0, s->stream>>>(a);
        FDTD_CUDA_TRY(cudaGetLastError());
        s->launches++;
        a.mode = 2;
        const bool fringe = (a.ib_lo[0] != s->main_lo[0]) || (a.ib_hi[0] != s->main_hi[0]);
        if (is_B && !do_pml && !fringe) return FDTD_OK;   // deferred half step: nothing left outside the inner box
    }
    // the PML = true instantiation: 2 cells per thread for float storage (see sweep_kernels.cuh)
    constexpr int VP = (sizeof(T) == 4) ? 2 : V;
    const dim3 grid_p((s->g.Ni + SWEEP_BX * VP - 1) / (SWEEP_BX * VP), gy, gz);
    if (is_B) sweep_B_kernel<T, true, VP, A><<<grid_p, block, 0, s->stream>>>(a);
    else sweep_E_kernel<T, true, VP, A><<<grid_p, block, 0, s->stream>>>(a);
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FDTD_OK;
}

// Rim sweep of the PML solver's two-step pass: the PML = true kernel over every cell OUTSIDE the box ib (vector
// granularity in i), reading generation `cur`; `to_new` = write the other generation instead of updating in place
// (B sweep: B' -> gen cur^1; E sweep: reads B from gen cur^1, old E from gen cur, E' -> gen cur^1).
template <typename T, typename A = double>
static fdtd_status_t launch_rim_sweep(Solver* s, bool is_B, int n_half, const int ib_lo[3], const int ib_hi[3], bool to_new) {
    constexpr int V = VecOf<T>::V;
    SweepArgs<T> a;
    a.g = s->g; a.c = s->c; a.p = make_pml(s); a.f = make_fields<T>(s); a.jbox = s->jbox;
    a.k_lo = 0; a.k_hi = s->g.nk;
    a.n_half = n_half; a.do_pml = 1;
    a.j_quirk = (s->cfg.flags & FDTD_FLAG_J_OPENMP_QUIRK) ? 1 : 0;
    a.mode = 2;
    for (int d = 0; d < 3; ++d) {
        a.ib_lo[d] = ib_lo[d]; a.ib_hi[d] = ib_hi[d];
        a.Bout[d] = a.f.B[d]; a.Eout[d] = a.f.E[d];
    }
    if (to_new) {
        for (int d = 0; d < 3; ++d) {
            T* nb = static_cast<T*>(s->p[BX + d][s->cur ^ 1]);
            T* ne = static_cast<T*>(s->p[EX + d][s->cur ^ 1]);
            if (is_B) a.Bout[d] = nb;
            else { a.f.B[d] = nb; a.Eout[d] = ne; }
        }
    }
    dim3 block(SWEEP_BX, SWEEP_BY);
    constexpr int VP = (sizeof(T) == 4) ? 2 : V;
    const int gx = (s->g.Ni + SWEEP_BX * VP - 1) / (SWEEP_BX * VP);
    const int gy = (s->g.Nj + SWEEP_BY - 1) / SWEEP_BY;
    a.kc = pick_kc(s, gx * gy, s->g.nk);
    dim3 grid(gx, gy, (s->g.nk + a.kc - 1) / a.kc);
    if (is_B) sweep_B_kernel<T, true, VP, A><<<grid, block, 0, s->stream>>>(a);
    else sweep_E_kernel<T, true, VP, A><<<grid, block, 0, s->stream>>>(a);
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FDTD_OK;
}

// Environment switches are read once per solver (Tunables, solver.h).
static int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
static void read_tunables(Tunables& t) {
    t.fused_variant = env_int("FDTD_B200_FUSED_VARIANT", -1);
    t.fused_kc = env_int("FDTD_B200_FUSED_KC", 0);
    t.t2_variant = env_int("FDTD_B200_T2_VARIANT", 0);
    t.t2_strip = env_int("FDTD_B200_T2_STRIP", 0);
    t.st_cs = env_int("FDTD_B200_ST_CS", 1);
    t.tma_l2 = env_int("FDTD_B200_TMA_L2", 256);
    t.no_tma = env_int("FDTD_B200_NO_TMA", 0) != 0;
    t.no_t2 = env_int("FDTD_B200_NO_T2", 0) != 0;
    t.no_lazy = env_int("FDTD_B200_NO_LAZY", 0) != 0;
    t.mgpu_debug = env_int("FDTD_B200_MGPU_DEBUG", 0);
    t.pml_t2_f32 = env_int("FDTD_B200_PML_T2_F32", 0) != 0;
    t.halo_timeout_s = std::max(1, env_int("FDTD_B200_HALO_TIMEOUT_S", 30));
}
// one cudaFuncSetAttribute per kernel family and solver (bit in Solver::configured)
enum { CFG_FUSED2 = 0, CFG_T2 = 8 };

template <typename T, int BY, int RJ, int MINB>
static cudaError_t launch_fused_variant(Solver* s, FusedArgs<T>& a) {
    constexpr int V = VecOf<T>::V;
    constexpr int TIU = FUSED_OUT_LANES * V;
    constexpr int TJU = BY * RJ - 2;
    const int gx = (s->g.Ni + TIU - 1) / TIU;
    const int gy = (s->g.Nj + TJU - 1) / TJU;
    const int np = a.k_hi - a.k_lo;
    int kc = s->tun.fused_kc;
    if (kc <= 0) {
        kc = 64;
        while (kc > 8 && (long long)gx * gy * ((np + kc - 1) / kc) < 148 * 6) kc /= 2;
    }
    if (kc > np) kc = np;
    a.kc = kc;
    const int gz = (np + kc - 1) / kc;
    fused_BE_kernel<T, BY, RJ, MINB><<<dim3(gx, gy, gz), dim3(FUSED_BX, BY), 0, s->launch_stream>>>(a);
    return cudaGetLastError();
}

template <typename T, int BY, int PF, int D, int MINB>
static cudaError_t launch_fused2_variant(Solver* s, FusedArgs<T>& a, int cfg_bit) {
    constexpr int V = VecOf<T>::V;
    constexpr int TIU = FUSED_OUT_LANES * V;
    constexpr int TJU = BY - 2;
    constexpr size_t smem = fused2_smem_bytes<BY, PF, D>();
    if (!(s->configured & (1u << cfg_bit))) {
        cudaError_t e = cudaFuncSetAttribute(fused_BE2_kernel<T, BY, PF, D, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        s->configured |= (1u << cfg_bit);
    }
    const int gx = (s->g.Ni + TIU - 1) / TIU;
    const int gy = (s->g.Nj + TJU - 1) / TJU;
    const int np = a.k_hi - a.k_lo;
    int kc = s->tun.fused_kc;
    if (kc <= 0) {
        kc = 64;
        while (kc > 8 && (long long)gx * gy * ((np + kc - 1) / kc) < 148 * 6) kc /= 2;
    }
    if (kc > np) kc = np;
    a.kc = kc;
    const int gz = (np + kc - 1) / kc;
    fused_BE2_kernel<T, BY, PF, D, MINB><<<dim3(gx, gy, gz), dim3(FUSED_BX, BY), smem, s->launch_stream>>>(a);
    return cudaGetLastError();
}

template <typename T>
static fdtd_status_t launch_fused(Solver* s, int n_half, int k_lo, int k_hi) {
    FusedArgs<T> a;
    a.g = s->g; a.c = s->c; a.jbox = s->jbox;
    for (int c = 0; c < 3; ++c) {
        a.Ein[c] = static_cast<const T*>(s->p[EX + c][s->cur]);
        a.Bin[c] = static_cast<const T*>(s->p[BX + c][s->cur]);
        a.Eout[c] = static_cast<T*>(s->p[EX + c][s->cur ^ 1]);
        a.Bout[c] = static_cast<T*>(s->p[BX + c][s->cur ^ 1]);
        a.J[c] = static_cast<const T*>(s->p[JX + c][0]);
    }
    a.k_lo = k_lo; a.k_hi = k_hi; a.n_half = n_half;
    a.j_quirk = (s->cfg.flags & FDTD_FLAG_J_OPENMP_QUIRK) ? 1 : 0;
    cudaError_t e;
    // Variants kept after the round-1 sweeps (profiles/sweep_t2_r01.md, profiles/t1_variants_r01.jsonl): fp64 default
    // v28 = 16 rows, register prefetch, 1 CTA/SM (2.34 ms at 512^3); fp32 default v0 = 8 rows, 2 CTAs/SM; v5 = 8 rows,
    // 3 CTAs/SM (2.42 ms, the round-1 first default); v21 = its register-prefetch twin.
    int variant = s->tun.fused_variant;
    if (variant < 0) variant = (sizeof(T) == 8) ? 28 : 0;
    switch (variant) {
        default:
        case 0: e = launch_fused_variant<T, 8, 1, 2>(s, a); break;
        case 5: e = launch_fused_variant<T, 8, 1, 3>(s, a); break;
        case 21: e = launch_fused2_variant<T, 8, 1, 1, 3>(s, a, CFG_FUSED2 + 0); break;
        case 28: e = launch_fused2_variant<T, 16, 1, 1, 1>(s, a, CFG_FUSED2 + 1); break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "fused_BE_kernel launch");
    s->launches++;
    return FDTD_OK;
}

// ---- temporally blocked pass: two Yee steps per launch (fused_kernel_t2.cuh) -------------------------------
// Variant table <BY rows per CTA, min CTAs/SM>; FDTD_B200_T2_VARIANT picks one.
struct T2Ranges { int lo, hi, lo2, hi2; };   // local plane ranges a T2 launch produces (second range empty when lo2 == hi2)

// Plane chunks of one T2 launch (host-only, pure: also exported as fdtd_debug_t2_chunk_plan for the CPU test suite).
//   nk            planes of the rank's slab            rg        the plane range(s) to produce
//   wait_in_kernel  the kernel waits for the halo itself (slab rank, peer transport): the planes whose dependency cone
//                 reaches a ghost plane become thin chunks at the END of the list
//   tiles, gx     working tiles per chunk / tile columns (L2 chunk-length bound)      slots = CTAs resident on the GPU
// Chunk length.  Every chunk costs 3 redundant plane iterations plus ~2 of start-up, and the CTAs run in waves of one per
// SM, so the host minimises  waves(m) x (planes per chunk + 5)  over the chunk count m -- long chunks, but a CTA count that
// fills its last wave (profiles/kc_sweep_r01.jsonl: 3 chunks of 171 beat 4 of 128 at 512^3).  Second constraint: CTAs are
// dispatched in id order as SMs free up, so the start-time skew between a tile and its y-neighbour (gx ids away) is about
// (chunk duration) x gx / (CTAs in flight); once it exceeds the time a line survives in L2, the 4 halo rows of every 16-row
// tile are read from DRAM twice.  Measured (profiles/kc_traffic_r01.jsonl): fine up to len * gx ~ 1540 (512^3, 171 planes:
// 1.09x compulsory traffic), 1.23x at 2300, 1.37x at 9200 (1024^3 with 512-plane chunks: 87 instead of 110 Gcell/s).
// Thin boundary chunks (H = the pass's reach, 2 output planes each) are issued LAST, after all interior chunks: by then the
// neighbours' pushes have long landed, and their short CTAs fill the interior's tail wave.  (A first version kept
// full-length boundary chunks and lost 11 % in the strong-scaling regime -- 1024^2 planes, two chunks of 64: the top
// chunk's first wave reached the ghost planes 0.13 ms into the pass, the halo took 0.41 ms, profiles/strong_probe_r02.md.)
// Each extra chunk costs ~5 plane iterations per tile: 2 % at 512 planes.
static int t2_chunk_plan(int nk, const T2Ranges& rg, bool wait_in_kernel, long long tiles, int gx, long long slots, int kc_override,
                         int* chunk_lo, int* chunk_hi) {
    const int H = 2;
    int lo = rg.lo, hi = rg.hi;
    bool thin_top = false, thin_bot = false;
    if (wait_in_kernel && rg.hi2 <= rg.lo2 && hi - lo >= 4 * H + 8) {
        thin_bot = lo - 2 < 0;
        thin_top = hi + 1 >= nk;
        if (thin_bot) lo += H;
        if (thin_top) hi -= H;
    }
    const int np = hi - lo, np2 = rg.hi2 > rg.lo2 ? rg.hi2 - rg.lo2 : 0;
    int kc = kc_override;
    if (kc <= 0) {
        long long best_cost = -1;
        int best_m = 1;
        const int len_cap = std::max(32, 1600 / std::max(gx, 1));
        for (int m = 1; m <= np; ++m) {
            const int len = (np + m - 1) / m;
            if (len > len_cap && len > 32) continue;
            if (len < 16 && m > 1) break;
            const int m_eff = (np + len - 1) / len + (np2 > 0 ? (np2 + len - 1) / len : 0);
            const long long waves = (tiles * m_eff + slots - 1) / slots;
            const long long cost = waves * (len + 5);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_m = m; }
        }
        kc = (np + best_m - 1) / best_m;
    }
    if (kc > np) kc = np;
    if (kc < 1) kc = 1;
    while ((np + kc - 1) / kc + (np2 + kc - 1) / kc + 2 > T2_MAXCH) ++kc;
    int gz = 0;
    for (int k = lo; k < hi; k += kc) { chunk_lo[gz] = k; chunk_hi[gz] = std::min(k + kc, hi); ++gz; }
    for (int k = rg.lo2; k < rg.hi2; k += kc) { chunk_lo[gz] = k; chunk_hi[gz] = std::min(k + kc, rg.hi2); ++gz; }
    if (thin_top) { chunk_lo[gz] = hi; chunk_hi[gz] = rg.hi; ++gz; }
    if (thin_bot) { chunk_lo[gz] = rg.lo; chunk_hi[gz] = lo; ++gz; }
    return gz;
}

template <typename T, typename A, int BY, int MINB, int ABL = 0>
static cudaError_t launch_t2_variant(Solver* s, FusedT2Args<T>& a, const T2Ranges& rg, int cfg_bit) {
    constexpr int V = t2_v<A>();   // cells per lane: 2 with double arithmetic (both storage types), 4 with float arithmetic
    constexpr int TIU = FUSED_OUT_LANES * V;
    constexpr int TJU = BY - 4;
    constexpr size_t smem = fused_t2_smem_bytes<T, A, BY>();
    if (!(s->configured & (1u << cfg_bit))) {
        cudaError_t e = cudaFuncSetAttribute(fused_BE_T2_kernel<T, A, BY, MINB, true, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_BE_T2_kernel<T, A, BY, MINB, false, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        s->configured |= (1u << cfg_bit);
    }
    const int gx = (s->g.Ni + TIU - 1) / TIU;
    const int gy = (s->g.Nj + TJU - 1) / TJU;
    // (CTAs whose output tile misses the store box exit at once: count the ones that work)
    const long long tiles = (long long)std::min(gx, (a.sb_hi[0] - a.sb_lo[0] + TIU - 1) / TIU + 1) *
                            std::min(gy, (a.sb_hi[1] - a.sb_lo[1] + TJU - 1) / TJU + 1);
    const int gz = t2_chunk_plan(s->g.nk, rg, a.halo_flags != nullptr, tiles, gx, 148LL * MINB, s->tun.fused_kc, a.chunk_lo, a.chunk_hi);
    a.nchunks = gz;
    a.gx = gx; a.gy = gy;
    {
        // tile columns per strip; >= gx = row-major ids: measured best (profiles/strip_r01.jsonl) -- a missed x-neighbour
        // costs as much as a missed y-neighbour (whole 128-byte lines at both ends of a 512-byte row)
        const int w = s->tun.t2_strip > 0 ? s->tun.t2_strip : gx;
        a.strip_w = w > gx ? gx : w;
    }
    if (a.n_half == 2) fused_BE_T2_kernel<T, A, BY, MINB, true, ABL><<<dim3(gx * gy, 1, gz), dim3(FUSED_BX, BY), smem, s->launch_stream>>>(a);
    else fused_BE_T2_kernel<T, A, BY, MINB, false, ABL><<<dim3(gx * gy, 1, gz), dim3(FUSED_BX, BY), smem, s->launch_stream>>>(a);
    return cudaGetLastError();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table (no link-time dependency on libcuda).
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
    // (looked up per call: a handful of calls per solver lifetime, the maps themselves are cached in the Solver)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<tmap_encode_fn>(p);
    cudaGetLastError();
    return nullptr;
}

// 3-D tensor map of one component array of generation `gen` (ghost planes included), box = {64 cells, by rows, 1 plane}.
static bool encode_tmap(const Solver* s, tmap_encode_fn enc, CUtensorMap* tm, int comp, int gen, int by) {
    const cuuint64_t dims[3] = {(cuuint64_t)s->g.Ni, (cuuint64_t)s->g.Nj, (cuuint64_t)(s->g.nk + 2 * GHOST_PLANES)};
    const cuuint64_t strides[2] = {(cuuint64_t)s->g.pitch * s->esz, (cuuint64_t)s->g.plane * s->esz};
    // 64 cells = 512 B (fp64) / 68 cells = 272 B (fp32 storage, double arithmetic) / 128 cells = 512 B (fp32 storage and arithmetic)
    const cuuint32_t box[3] = {(cuuint32_t)(s->esz == 8 ? t2_rbox<double, double>() : s->f32_arith ? t2_rbox<float, float>() : t2_rbox<float, double>()), (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (box[0] > (cuuint32_t)s->g.Ni || by > s->g.Nj) return false;   // no wrap-free tile exists anyway
    const int promo = s->tun.tma_l2;   // FDTD_B200_TMA_L2 = 0 none, 64, 128, 256 (default; measured, profiles/tma_l2_r01.jsonl)
    const CUtensorMapL2promotion l2 = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    const CUresult r = enc(tm, s->esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           s->base[comp][gen], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// The six maps of generation `gen` (box rows `by`): encoded on first use, then served from the solver's cache (the
// arrays never move; round 1 re-encoded all six on every launch).
static const CUtensorMap* cached_tmaps(Solver* s, int gen, int by) {
    if (s->tun.no_tma) return nullptr;
    if (s->tmaps_by[gen] != by) {
        tmap_encode_fn enc = tmap_encoder();
        bool ok = enc != nullptr;
        for (int c = 0; c < 6 && ok; ++c) ok = encode_tmap(s, enc, &s->tmaps[gen][c], EX + c, gen, by);
        if (!ok) { s->tmaps_by[gen] = -1; return nullptr; }
        s->tmaps_by[gen] = by;
    }
    return s->tmaps[gen];
}

template <typename T, typename A = double>
static fdtd_status_t launch_t2(Solver* s, int n_half, int k_lo, int k_hi, int k_lo2, int k_hi2, int src2, double amp2, unsigned halo_seq) {
    static const int by_of_variant[] = {16, 8, 12};
    int variant = s->tun.t2_variant;
    if (variant < 0 || (variant > 2 && variant < 11) || variant > 15) variant = 0;
    FusedT2Args<T> a;
    std::memset(&a, 0, sizeof(a));
    a.g = s->g; a.c = s->c; a.jbox = s->jbox;
    for (int c = 0; c < 3; ++c) {
        a.Ein[c] = static_cast<const T*>(s->p[EX + c][s->cur]);
        a.Bin[c] = static_cast<const T*>(s->p[BX + c][s->cur]);
        a.Eout[c] = static_cast<T*>(s->p[EX + c][s->cur ^ 1]);
        a.Bout[c] = static_cast<T*>(s->p[BX + c][s->cur ^ 1]);
        a.J[c] = static_cast<const T*>(s->p[JX + c][0]);
        a.s_lo[c] = s->src_lo[c]; a.s_hi[c] = s->src_hi[c]; a.sw[c] = s->d_w[c];
    }
    const T2Ranges rg{k_lo, k_hi, k_lo2, k_hi2};
    a.n_half = n_half;
    a.j_quirk = (s->cfg.flags & FDTD_FLAG_J_OPENMP_QUIRK) ? 1 : 0;
    a.src2 = src2; a.amp2 = amp2;
    if (src2 == 2) {   // pending J writes: stage B takes J on the box from the dense buffers
        for (int c = 0; c < 3; ++c) {
            a.s_lo[c] = s->jp_lo[c]; a.s_hi[c] = s->jp_hi[c];
            a.jb2[c] = static_cast<const T*>(s->d_jpend) + (size_t)c * Solver::JPEND_MAX_CELLS;
        }
    }
    for (int d = 0; d < 2; ++d) {
        a.sb_lo[d] = s->pml_t2 ? s->sb_lo[d] : 0;
        a.sb_hi[d] = s->pml_t2 ? s->sb_hi[d] : (d == 0 ? s->g.Ni : s->g.Nj);
    }
    a.st_cs = s->tun.st_cs;   // profiles/stcs_r01.jsonl: DRAM reads -3 %, +1.5 %
    if (halo_seq) {            // the CTAs that read ghost planes wait for the neighbours' pushes themselves (peer_ring.cu)
        a.halo_flags = peer_data_flags(s);
        a.halo_err = peer_error_word(s);
        a.halo_seq = halo_seq;
        a.halo_timeout_ns = (unsigned long long)s->tun.halo_timeout_s * 1000000000ull;
    }
    const int by = variant < 3 ? by_of_variant[variant] : 16;
    if (const CUtensorMap* tm = cached_tmaps(s, s->cur, by)) {
        a.use_tma = 1;
        for (int c = 0; c < 3; ++c) { a.tmE[c] = tm[c]; a.tmB[c] = tm[3 + c]; }
    }
    cudaError_t e;
    switch (variant) {
        default:
        case 0: e = launch_t2_variant<T, A, 16, 1>(s, a, rg, CFG_T2 + 0 + (sizeof(A) == 4 ? 8 : 0)); break;
        case 1: e = launch_t2_variant<T, A, 8, 2>(s, a, rg, CFG_T2 + 1 + (sizeof(A) == 4 ? 8 : 0)); break;
        case 2: e = launch_t2_variant<T, A, 12, 1>(s, a, rg, CFG_T2 + 2 + (sizeof(A) == 4 ? 8 : 0)); break;
#ifdef FDTD_T2_ABLATE   /* timing experiments only: results are wrong */
        case 11: e = launch_t2_variant<T, A, 16, 1, 1>(s, a, rg, CFG_T2 + 3); break;
        case 12: e = launch_t2_variant<T, A, 16, 1, 2>(s, a, rg, CFG_T2 + 4); break;
        case 13: e = launch_t2_variant<T, A, 16, 1, 3>(s, a, rg, CFG_T2 + 5); break;
        case 14: e = launch_t2_variant<T, A, 16, 1, 4>(s, a, rg, CFG_T2 + 6); break;
#endif
    }
    if (e != cudaSuccess) return cuda_fail(e, "fused_BE_T2_kernel launch");
    s->launches++;
    return FDTD_OK;
}

template <typename T>
static fdtd_status_t launch_source(Solver* s, double amp, int zero) {
    SourceArgs sa;
    for (int a = 0; a < 3; ++a) { sa.lo[a] = s->src_lo[a]; sa.hi[a] = s->src_hi[a]; sa.w[a] = s->d_w[a]; }
    sa.amp = amp; sa.zero = zero;
    const long long total = (long long)(sa.hi[0] - sa.lo[0]) * (sa.hi[1] - sa.lo[1]) * (sa.hi[2] - sa.lo[2]);
    if (total <= 0) return FDTD_OK;
    const int threads = 128;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    source_kernel<T><<<(int)blocks, threads, 0, s->stream>>>(static_cast<T*>(s->p[JX][0]), static_cast<T*>(s->p[JY][0]),
                                                            static_cast<T*>(s->p[JZ][0]), s->g, sa);
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FDTD_OK;
}

#define DISPATCH(s, fn, ...) ((s)->dtype == FDTD_F32 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))
// (storage, arithmetic) pairs: (double, double), (float, double) and, with FDTD_FLAG_F32_ARITH, (float, float)
#define DISPATCH_A(s, fn, ...) ((s)->dtype == FDTD_F32 ? ((s)->f32_arith ? fn<float, float>(__VA_ARGS__) : fn<float, double>(__VA_ARGS__)) : fn<double, double>(__VA_ARGS__))

// ------------------------------------------------------------------------------------------------------
// halo exchange (z-slab ring).  Plane -1 / plane nk of every array are the ghost planes.
// ------------------------------------------------------------------------------------------------------
static char* plane_ptr(const Solver* s, int comp, int gen, int local_plane) {
    return static_cast<char*>(s->p[comp][gen]) + (long long)local_plane * s->g.plane * (long long)s->esz;
}

// Exchange lists.  One entry = a contiguous range of planes of one array, described both ways: as a send / recv pair for
// the NCCL transport and as a push into the neighbour's ghost planes for the copy-engine transport (peer_ring.cu).
struct XferList {
    PlaneXfer x[24];
    int n = 0;
};
// my planes [src, src + np) -> the upper neighbour's bottom ghost planes [dst, dst + np), dst < 0 (the mirror image arrives
// from the lower neighbour into my own planes [dst, dst + np))
static void push_up(const Solver* s, XferList& l, int comp, int gen, int src, int dst, int np) {
    const int up = (s->cfg.rank + 1) % s->cfg.nranks, down = (s->cfg.rank + s->cfg.nranks - 1) % s->cfg.nranks;
    const size_t bytes = (size_t)s->g.plane * s->esz * np;
    l.x[l.n++] = PlaneXfer{plane_ptr(s, comp, gen, src), up, plane_ptr(s, comp, gen, dst), down, bytes, 1, comp, gen, dst};
}
// my planes [src, src + np) -> the lower neighbour's top ghost planes nk_peer + [d, d + np) (mirror: from the upper
// neighbour into my planes nk + [d, d + np))
static void push_down(const Solver* s, XferList& l, int comp, int gen, int src, int d, int np) {
    const int up = (s->cfg.rank + 1) % s->cfg.nranks, down = (s->cfg.rank + s->cfg.nranks - 1) % s->cfg.nranks;
    const size_t bytes = (size_t)s->g.plane * s->esz * np;
    l.x[l.n++] = PlaneXfer{plane_ptr(s, comp, gen, src), down, plane_ptr(s, comp, gen, s->g.nk + d), up, bytes, 0, comp, gen, d};
}

// Both transports have the same meaning on `stream`: when the call's work has run, the ghost planes hold the neighbours'
// planes -- except wait_data = false on the peer transport, where the consumer (the T2 pass kernel) waits itself.
static fdtd_status_t ring_exchange(Solver* s, const XferList& l, cudaStream_t stream, bool wait_data = true,
                                   cudaEvent_t ev_start = nullptr, cudaEvent_t ev_end = nullptr) {
    if (s->peer) return peer_exchange(s, l.x, l.n, stream, wait_data, ev_start, ev_end);
    if (ev_start) FDTD_CUDA_TRY(cudaEventRecord(ev_start, stream));
    fdtd_status_t st = nccl_exchange(s, l.x, l.n, stream);
    if (st == FDTD_OK && ev_end) FDTD_CUDA_TRY(cudaEventRecord(ev_end, stream));
    return st;
}

// Top ghost of Ex,Ey <- upper neighbour's bottom plane (coarray/fdtd.F90:97-98).
static fdtd_status_t exchange_E_top(Solver* s) {
    if (s->cfg.nranks <= 1 || s->ghosts_e_valid) return FDTD_OK;
    XferList l;
    for (int c = 0; c < 2; ++c) push_down(s, l, EX + c, s->cur, 0, 0, 1);
    fdtd_status_t st = ring_exchange(s, l, s->stream);
    if (st == FDTD_OK) s->ghosts_e_valid = true;
    return st;
}

// Bottom ghost of Bx,By <- lower neighbour's top plane (coarray/fdtd.F90:90-91).
static fdtd_status_t exchange_B_bottom(Solver* s) {
    if (s->cfg.nranks <= 1 || s->ghosts_b_valid) return FDTD_OK;
    XferList l;
    for (int c = 0; c < 2; ++c) push_up(s, l, BX + c, s->cur, s->g.nk - 1, -1, 1);
    fdtd_status_t st = ring_exchange(s, l, s->stream);
    if (st == FDTD_OK) s->ghosts_b_valid = true;
    return st;
}

// Everything the fused pass needs: bottom ghost of Bx,By,Ex,Ey,Ez (to rebuild B'(-1)) and top ghost of Ex,Ey.
static fdtd_status_t exchange_fused(Solver* s, cudaStream_t stream) {
    if (s->cfg.nranks <= 1 || s->ghosts_fused_valid) return FDTD_OK;
    XferList l;
    const int up_comps[5] = {BX, BY, EX, EY, EZ};
    for (int c = 0; c < 5; ++c) push_up(s, l, up_comps[c], s->cur, s->g.nk - 1, -1, 1);
    for (int c = 0; c < 2; ++c) push_down(s, l, EX + c, s->cur, 0, 0, 1);
    fdtd_status_t st = ring_exchange(s, l, stream);
    if (st == FDTD_OK) { s->ghosts_fused_valid = true; s->ghosts_e_valid = true; s->ghosts_b_valid = true; }
    return st;
}

// Everything the T2 pass needs: two bottom ghost planes of E, B (to rebuild B1(-2), B1(-1), E1(-1)) plus J(-1),
// top ghost plane nk of E, B, J (B1(nk), E1(nk)) and top ghost plane nk+1 of Ex, Ey.  J travels because stage A
// re-computes the neighbour's boundary plane, current term included.  26 planes each way, 18 contiguous ranges.
static fdtd_status_t exchange_t2(Solver* s, cudaStream_t stream, bool wait_data = true, cudaEvent_t ev_start = nullptr,
                                 cudaEvent_t ev_end = nullptr) {
    if (s->cfg.nranks <= 1 || s->ghosts_t2_valid) return FDTD_OK;
    const int nk = s->g.nk;
    XferList l;
    // to the upper neighbour: our top planes nk-2, nk-1 become its planes -2, -1
    for (int c = EX; c <= BZ; ++c) push_up(s, l, c, s->cur, nk - 2, -2, 2);
    for (int c = JX; c <= JZ; ++c) push_up(s, l, c, 0, nk - 1, -1, 1);
    // to the lower neighbour: our bottom plane 0 becomes its plane nk (E, B, J), our plane 1 its plane nk+1 (Ex, Ey)
    for (int c = EX; c <= EY; ++c) push_down(s, l, c, s->cur, 0, 0, 2);
    for (int c = EZ; c <= BZ; ++c) push_down(s, l, c, s->cur, 0, 0, 1);
    for (int c = JX; c <= JZ; ++c) push_down(s, l, c, 0, 0, 0, 1);
    fdtd_status_t st = ring_exchange(s, l, stream, wait_data, ev_start, ev_end);
    if (st == FDTD_OK) { s->ghosts_t2_valid = true; s->ghosts_fused_valid = true; s->ghosts_e_valid = true; s->ghosts_b_valid = true; }
    return st;
}

// Mid-pass exchanges of the PML solver's two-step pass on slab ranks: (Bx, By) top plane of generation `gen` -> upper
// neighbour's plane -1, or (Ex, Ey) bottom plane -> lower neighbour's plane nk (only the rim cells of those planes
// are meaningful on both sides, and only those are read).
static fdtd_status_t exchange_pair_planes(Solver* s, bool b_to_upper, int gen) {
    if (s->cfg.nranks <= 1) return FDTD_OK;
    XferList l;
    for (int c = 0; c < 2; ++c) {
        if (b_to_upper) push_up(s, l, BX + c, gen, s->g.nk - 1, -1, 1);
        else push_down(s, l, EX + c, gen, 0, 0, 1);
    }
    return ring_exchange(s, l, s->stream);
}

static void invalidate_ghosts(Solver* s) {
    s->ghosts_e_valid = s->ghosts_b_valid = s->ghosts_fused_valid = s->ghosts_t2_valid = false;
}

// ------------------------------------------------------------------------------------------------------
// the step state machine
// ------------------------------------------------------------------------------------------------------

// Apply the deferred trailing B half step (main cells only; the PML shell advances once per step).
static fdtd_status_t flush_pending(Solver* s) {
    if (!s->b_pending) return FDTD_OK;
    nvtxRangePushA("UpdateBField (deferred trailing half step)");
    struct Pop { ~Pop() { nvtxRangePop(); } } pop;
    fdtd_status_t st = exchange_E_top(s);
    if (st != FDTD_OK) return st;
    st = DISPATCH_A(s, launch_sweep, s, true, 1, 0);
    if (st != FDTD_OK) return st;
    s->b_pending = false;
    s->ghosts_b_valid = false;
    s->ghosts_fused_valid = false;
    return FDTD_OK;
}

static void jbox_union(Solver* s, const int lo[3], const int hi[3]) {
    if (lo[0] >= hi[0] || lo[1] >= hi[1] || lo[2] >= hi[2]) return;
    if (s->jbox.empty()) {
        for (int a = 0; a < 3; ++a) { s->jbox.lo[a] = lo[a]; s->jbox.hi[a] = hi[a]; }
        return;
    }
    for (int a = 0; a < 3; ++a) {
        if (lo[a] < s->jbox.lo[a]) s->jbox.lo[a] = lo[a];
        if (hi[a] > s->jbox.hi[a]) s->jbox.hi[a] = hi[a];
    }
}
static void jbox_clear(Solver* s) {
    for (int a = 0; a < 3; ++a) { s->jbox.lo[a] = 0; s->jbox.hi[a] = 0; }
}
static void jbox_full(Solver* s) {
    s->jbox.lo[0] = s->jbox.lo[1] = s->jbox.lo[2] = 0;
    s->jbox.hi[0] = s->g.Ni; s->jbox.hi[1] = s->g.Nj; s->jbox.hi[2] = s->g.Nk;
}

static fdtd_status_t zero_currents_impl(Solver* s) {
    if (s->jbox.empty()) return FDTD_OK;   // J is already +0.0 everywhere
    const long long vol = (long long)(s->jbox.hi[0] - s->jbox.lo[0]) * (s->jbox.hi[1] - s->jbox.lo[1]) *
                          (s->jbox.hi[2] - s->jbox.lo[2]);
    if (vol <= (1 << 16)) {
        // small box: the source kernel with zero=1 writes +0.0 on the box only (the tables are not read)
        int lo[3], hi[3];
        for (int a = 0; a < 3; ++a) { lo[a] = s->src_lo[a]; hi[a] = s->src_hi[a]; }
        for (int a = 0; a < 3; ++a) { s->src_lo[a] = s->jbox.lo[a]; s->src_hi[a] = s->jbox.hi[a]; }
        fdtd_status_t st = DISPATCH(s, launch_source, s, 0.0, 1);
        for (int a = 0; a < 3; ++a) { s->src_lo[a] = lo[a]; s->src_hi[a] = hi[a]; }
        if (st != FDTD_OK) return st;
    } else {
        const size_t bytes = (size_t)s->g.plane * (size_t)(s->g.nk + 2 * GHOST_PLANES) * s->esz;
        for (int c = JX; c <= JZ; ++c) FDTD_CUDA_TRY(cudaMemsetAsync(s->base[c][0], 0, bytes, s->stream));
    }
    jbox_clear(s);
    return FDTD_OK;
}

// The J arrays lag the device-resident source by one step after a T2 pass that evaluated the second step's
// source in the kernel: write that step's values now (anything that reads J, or the end of fdtd_step()).
static fdtd_status_t materialize_J(Solver* s) {
    if (!s->j_stale) return FDTD_OK;
    s->j_stale = false;
    if (s->src_t < 1 || s->src_t > (int)s->src_amp.size()) return FDTD_OK;
    return DISPATCH(s, launch_source, s, s->src_amp[s->src_t - 1], 0);
}

// ---- pending J writes (see Solver::jpend) ---------------------------------------------------------------------------
template <typename T>
static fdtd_status_t jbox_launch(Solver* s, int comp, int mode, const long long* d_idx, const T* d_vals, int n) {
    JBoxArgs b;
    for (int a = 0; a < 3; ++a) { b.lo[a] = s->jp_lo[a]; b.hi[a] = s->jp_hi[a]; }
    T* box = static_cast<T*>(s->d_jpend) + (size_t)(comp - JX) * Solver::JPEND_MAX_CELLS;
    const long long cells = (long long)(b.hi[0] - b.lo[0]) * (b.hi[1] - b.lo[1]) * (b.hi[2] - b.lo[2]);
    const int threads = 128;
    const int blocks = (int)(((mode == 2 ? (long long)n : cells) + threads - 1) / threads);
    jbox_kernel<T><<<blocks, threads, 0, s->stream>>>(static_cast<T*>(s->p[comp][0]), s->g, b, box, mode, d_idx, d_vals, n);
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FDTD_OK;
}

// Write the pending box into the J arrays (after the step that had to see the old J has been issued).
static fdtd_status_t apply_pending_J(Solver* s) {
    if (!s->jpend) return FDTD_OK;
    s->jpend = false;
    for (int c = JX; c <= JZ; ++c) {
        fdtd_status_t st = (s->dtype == FDTD_F32) ? jbox_launch<float>(s, c, 1, nullptr, nullptr, 0) : jbox_launch<double>(s, c, 1, nullptr, nullptr, 0);
        if (st != FDTD_OK) return st;
    }
    return FDTD_OK;
}

// ---- per-pass timeline (fdtd_timeline_enable): four timing events per overlapped pass -------------------------------
//   [0] pass start (compute stream, before anything of the pass)      [1] halo copies start (after the ready handshake)
//   [2] halo copies issued and done on this rank's side               [3] pass end (compute stream, kernels + halo)
static cudaEvent_t tl_event(Solver* s, int which) {
    if (s->tl_n >= s->tl_cap) return nullptr;
    return s->tl_events[(size_t)s->tl_n * 4 + which];
}

// One launch group over the slab with the halo exchange overlapped.  launch(lo, hi, lo2, hi2, halo_seq): second plane
// range empty when lo2 == hi2; halo_seq != 0 = the kernel waits for exchange number halo_seq itself.
//
// (A) peer transport + T2 pass (the default on a ring): ONE launch over the whole slab.  The copy engines push the 26
//     boundary planes into the neighbours' ghost planes while the pass runs; the planes that read ghost planes form two
//     thin chunks at the end of the chunk list and their CTAs wait on the sequence flags -- no boundary launch, no third
//     stream, no SM ever runs anything but the pass:
//        comm stream    : wait(previous work) -> ready handshake -> plane copies -> data flags -> ev_b
//        compute stream : the pass (whole slab) -> wait(ev_b: our planes have left before the next pass overwrites them)
// (B) NCCL transport, or a kernel that cannot wait: three streams.  The pass kernels own every SM (one 512-thread CTA holds
//     the whole register file), so (1) the comm stream has the highest priority: NCCL's few CTAs are dispatched to the first
//     SMs that free up instead of behind the interior launch's ~1500 pending CTAs; (2) the boundary slabs go to a third,
//     normal-priority stream: they read only the input generation and write planes the interior launch does not, so they are
//     independent of it, and their short CTAs fill its partial last wave:
//        comm stream    : wait(previous work) -> ring exchange into the ghost planes -> ev_b
//        compute stream : interior planes [H, nk-H) ........................................ -> wait(ev_c)
//        boundary stream: wait(previous work, ev_b) -> the two boundary slabs -> ev_c
// FDTD_B200_MGPU_DEBUG (timing experiments only, tools/mgpu_probe.py): bit 0 = no overlap, bit 1 = skip the exchange
// (wrong fields), bit 2 = boundary slabs on the compute stream (after the interior launch), bit 3 = comm stream without
// priority, bit 4 = structure (B) on the peer transport, bits 8.. = boundary depth H.
template <typename LaunchFn, typename ExchangeFn>
static fdtd_status_t overlapped(Solver* s, bool ghosts_valid, bool kernel_can_wait, LaunchFn launch, ExchangeFn exchange_in) {
    const int dbg = (s->cfg.nranks > 1) ? s->tun.mgpu_debug : 0;
    const int H = (dbg >> 8) > 0 ? (dbg >> 8) : 2;   // boundary depth (planes) computed after the halo has landed (>= 2: the T2 pass reads k +- 2)
    auto exchange = [&](cudaStream_t q, bool wait, cudaEvent_t e0, cudaEvent_t e1) -> fdtd_status_t {
        return (dbg & 2) ? FDTD_OK : exchange_in(q, wait, e0, e1);
    };
    fdtd_status_t st;
    s->launch_stream = s->stream;
    const bool ring = s->cfg.nranks > 1 && !ghosts_valid && !(s->cfg.flags & FDTD_FLAG_NO_OVERLAP) && !(dbg & 1);
    const bool timeline = ring && s->tl_n < s->tl_cap;
    if (timeline) FDTD_CUDA_TRY(cudaEventRecord(tl_event(s, 0), s->stream));
    if (ring && s->peer && s->halo_in_kernel && kernel_can_wait && !(dbg & (2 | 16))) {
        FDTD_CUDA_TRY(cudaEventRecord(s->ev_a, s->stream));
        FDTD_CUDA_TRY(cudaStreamWaitEvent(s->comm_stream, s->ev_a, 0));
        if ((st = exchange(s->comm_stream, false, timeline ? tl_event(s, 1) : nullptr, timeline ? tl_event(s, 2) : nullptr)) != FDTD_OK) return st;
        FDTD_CUDA_TRY(cudaEventRecord(s->ev_b, s->comm_stream));
        if ((st = launch(0, s->g.nk, 0, 0, peer_last_seq(s))) != FDTD_OK) return st;
        // whatever follows on the compute stream may read any ghost plane (PML rim sweeps; a rank whose T2 launch was
        // clipped away from a slab boundary never waited in the kernel): satisfied long ago, costs two front-end ops
        if ((st = peer_wait_data(s, s->stream)) != FDTD_OK) return st;
        FDTD_CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_b, 0));
        if (timeline) { FDTD_CUDA_TRY(cudaEventRecord(tl_event(s, 3), s->stream)); s->tl_n++; }
        return FDTD_OK;
    }
    if (ring && H >= 2 && s->g.nk >= 4 * H) {
        FDTD_CUDA_TRY(cudaEventRecord(s->ev_a, s->stream));
        FDTD_CUDA_TRY(cudaStreamWaitEvent(s->comm_stream, s->ev_a, 0));
        if ((st = exchange(s->comm_stream, true, timeline ? tl_event(s, 1) : nullptr, timeline ? tl_event(s, 2) : nullptr)) != FDTD_OK) return st;
        FDTD_CUDA_TRY(cudaEventRecord(s->ev_b, s->comm_stream));
        if ((st = launch(H, s->g.nk - H, 0, 0, 0u)) != FDTD_OK) return st;
        if (dbg & 4) {
            FDTD_CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_b, 0));
            st = launch(0, H, s->g.nk - H, s->g.nk, 0u);
        } else {
            FDTD_CUDA_TRY(cudaStreamWaitEvent(s->bnd_stream, s->ev_a, 0));
            FDTD_CUDA_TRY(cudaStreamWaitEvent(s->bnd_stream, s->ev_b, 0));
            s->launch_stream = s->bnd_stream;
            st = launch(0, H, s->g.nk - H, s->g.nk, 0u);
            s->launch_stream = s->stream;
            if (st != FDTD_OK) return st;
            FDTD_CUDA_TRY(cudaEventRecord(s->ev_c, s->bnd_stream));
            FDTD_CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_c, 0));
        }
        if (st == FDTD_OK && timeline) { FDTD_CUDA_TRY(cudaEventRecord(tl_event(s, 3), s->stream)); s->tl_n++; }
        return st;
    }
    if ((st = exchange(s->stream, true, nullptr, nullptr)) != FDTD_OK) return st;
    return launch(0, s->g.nk, 0, 0, 0u);
}

// PML solver, two steps in one pass.  Reach of the T2 pass is 2 cells, so it is exact on the store box
// SB = main box shrunk by 2 (on the axes that have a shell; i bounds aligned to the vector width) as long as it reads
// generation `cur` untouched.  Everything outside SB -- the shell and a thin rim of main cells -- is advanced by the
// sweep kernels, with the reference's per-cell arithmetic (FDTD_PML.cpp:343-365), on shrinking regions:
//     T2 pass                 gen cur -> gen new on SB                       (S0 -> S2)
//     B sweep, in place       outside SB shrunk by 2 (2V in i)               (B1 needs E0 at +1: any cell)
//     E sweep, in place       outside SB shrunk by 1 (V in i)                (E1 needs B1 at -1: covered by the line above)
//     [device source: J of the second step]
//     B sweep, cur -> new     outside SB                                     (B2 needs E1 at +1: covered)
//     E sweep, cur -> new     outside SB, B read from gen new                (E2 needs B2 at -1: rim or SB, both in gen new)
// Every cell ends up in gen new; the split fields are only touched by the sweeps, twice, in place.
static void shrink_box(const Solver* s, int by, int by_i, int lo[3], int hi[3]) {
    for (int a = 0; a < 3; ++a) {
        const int d = (s->pml[a] > 0) ? (a == 0 ? by_i : by) : 0;   // no shell on this axis: the box spans the periodic axis
        lo[a] = s->sb_lo[a] + d;
        hi[a] = s->sb_hi[a] - d;
    }
}

static fdtd_status_t advance_pml_pair(Solver* s, int n_half, int src2, double amp2) {
    const int V = (int)(16 / s->esz);
    fdtd_status_t st;
    // local planes of the store box on this rank (possibly none: a rank inside the k shell only runs rim sweeps)
    const int klo = std::max(s->sb_lo[2] - s->g.k0, 0), khi = std::min(s->sb_hi[2] - s->g.k0, s->g.nk);
    auto t2_clipped = [&](int lo, int hi, int lo2, int hi2, unsigned halo_seq) -> fdtd_status_t {
        lo = std::max(lo, klo); hi = std::min(hi, khi);
        lo2 = std::max(lo2, klo); hi2 = std::min(hi2, khi);
        if (hi <= lo) { lo = lo2; hi = hi2; lo2 = hi2 = 0; }
        if (hi2 <= lo2) lo2 = hi2 = 0;
        if (hi <= lo) return FDTD_OK;
        return DISPATCH_A(s, launch_t2, s, n_half, lo, hi, lo2, hi2, src2, amp2, halo_seq);
    };
    // slab ranks: ring exchange of the two ghost planes per side (+ J) overlapped with the interior planes, as in the
    // periodic solver; single GPU: one launch
    st = overlapped(s, s->ghosts_t2_valid, true, t2_clipped,
                    [&](cudaStream_t q, bool wait, cudaEvent_t e0, cudaEvent_t e1) { return exchange_t2(s, q, wait, e0, e1); });
    if (st != FDTD_OK) return st;
    int lo[3], hi[3];
    shrink_box(s, 2, 2 * V, lo, hi);
    if ((st = DISPATCH_A(s, launch_rim_sweep, s, true, n_half, lo, hi, false)) != FDTD_OK) return st;
    if ((st = exchange_pair_planes(s, true, s->cur)) != FDTD_OK) return st;          // B1 top plane -> upper rank's plane -1
    shrink_box(s, 1, V, lo, hi);
    if ((st = DISPATCH_A(s, launch_rim_sweep, s, false, 0, lo, hi, false)) != FDTD_OK) return st;
    if ((st = exchange_pair_planes(s, false, s->cur)) != FDTD_OK) return st;         // E1 bottom plane -> lower rank's plane nk
    if (src2 == 1) {
        // the rim may meet the source box: the second step's sweeps read J from the arrays
        if ((st = DISPATCH(s, launch_source, s, amp2, 0)) != FDTD_OK) return st;
    } else if (src2 == 2) {
        if ((st = apply_pending_J(s)) != FDTD_OK) return st;   // pending host writes: the arrays take them now (the T2 core read the box)
    }
    shrink_box(s, 0, 0, lo, hi);
    if ((st = DISPATCH_A(s, launch_rim_sweep, s, true, 2, lo, hi, true)) != FDTD_OK) return st;
    if ((st = exchange_pair_planes(s, true, s->cur ^ 1)) != FDTD_OK) return st;      // B2 top plane (new generation)
    if ((st = DISPATCH_A(s, launch_rim_sweep, s, false, 0, lo, hi, true)) != FDTD_OK) return st;
    return FDTD_OK;
}

// NVTX ranges around every launch group, named after the reference's Kokkos kernel labels (kokkos_functors.h:62,126,265,362:
// "UpdateEField", "UpdateBField", "UpdateEPMLField", "UpdateBPMLField") so that a timeline of this library reads like
// a Kokkos-tools timeline of the reference.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static fdtd_status_t advance_impl(Solver* s, int remaining, int* done);

// Advance by one step, or by two when the temporally blocked pass applies; *done = steps advanced.
static fdtd_status_t advance(Solver* s, int remaining, int* done) {
    const bool pair = remaining >= 2 && !s->tun.no_t2 && (s->t2 || s->pml_t2);
    NvtxRange r(s->has_pml ? (pair ? "update_fields x2: UpdateBField+UpdateBPMLField+UpdateEField+UpdateEPMLField (two-step pass + rim sweeps)"
                                   : "update_fields: UpdateBField+UpdateBPMLField, UpdateEField+UpdateEPMLField (sweeps)")
                           : (pair ? "update_fields x2: UpdateBField+UpdateEField (fused two-step pass)"
                                   : "update_fields: UpdateBField+UpdateEField"));
    return advance_impl(s, remaining, done);
}

static fdtd_status_t advance_impl(Solver* s, int remaining, int* done) {
    fdtd_status_t st;
    *done = 1;
    // device-resident source: write J for this step (kokkos_sample.cpp:91-108), or retire it (sample.cpp:84)
    if (s->src_active) {
        if (s->src_t < (int)s->src_amp.size()) {
            st = DISPATCH(s, launch_source, s, s->src_amp[s->src_t], 0);
            if (st != FDTD_OK) return st;
            jbox_union(s, s->src_lo, s->src_hi);
            s->src_t++;
            s->j_stale = false;
        } else {
            s->src_active = false;
            s->j_stale = false;
            st = zero_currents_impl(s);
            if (st != FDTD_OK) return st;
        }
    }
    const int n_half = s->b_pending ? 2 : 1;
    if (s->pml_t2 && remaining >= 2 && !s->tun.no_t2 &&
        !(s->src_active && s->src_t >= (int)s->src_amp.size())) {   // (the source does not retire between the two steps)
        const int src2 = s->src_active ? 1 : (s->jpend ? 2 : 0);
        const double amp2 = src2 == 1 ? s->src_amp[s->src_t] : 0.0;
        st = advance_pml_pair(s, n_half, src2, amp2);
        if (st != FDTD_OK) return st;
        if (src2 == 1) { s->src_t++; s->j_stale = false; }   // advance_pml_pair has written the second step's J
        s->passes_t2++;
        *done = 2;
        s->cur ^= 1;
    } else if (s->fused) {
        // Pair this step with the next one unless the source retires in between (J would have to change to zero).
        const bool src_ends = s->src_active && s->src_t >= (int)s->src_amp.size();
        if (s->t2 && remaining >= 2 && !src_ends && !s->tun.no_t2) {
            const int src2 = s->src_active ? 1 : (s->jpend ? 2 : 0);
            const double amp2 = src2 == 1 ? s->src_amp[s->src_t] : 0.0;
            st = overlapped(s, s->ghosts_t2_valid, true,
                            [&](int lo, int hi, int lo2, int hi2, unsigned halo_seq) { return DISPATCH_A(s, launch_t2, s, n_half, lo, hi, lo2, hi2, src2, amp2, halo_seq); },
                            [&](cudaStream_t q, bool wait, cudaEvent_t e0, cudaEvent_t e1) { return exchange_t2(s, q, wait, e0, e1); });
            if (st != FDTD_OK) return st;
            if (src2 == 1) { s->src_t++; s->j_stale = true; }
            s->passes_t2++;
            *done = 2;
        } else if (s->f32_arith) {
            // the one-step fused kernels exist in double arithmetic only: a single step runs the two sweeps in place
            if ((st = exchange_E_top(s)) != FDTD_OK) return st;
            if ((st = DISPATCH_A(s, launch_sweep, s, true, n_half, 1)) != FDTD_OK) return st;
            s->ghosts_b_valid = false;
            if ((st = exchange_B_bottom(s)) != FDTD_OK) return st;
            if ((st = DISPATCH_A(s, launch_sweep, s, false, 0, 0)) != FDTD_OK) return st;
            s->cur ^= 1;   // (undone below: the sweeps work in place on the live generation)
        } else {
            st = overlapped(s, s->ghosts_fused_valid, false,
                            [&](int lo, int hi, int lo2, int hi2, unsigned) {
                                fdtd_status_t r = DISPATCH(s, launch_fused, s, n_half, lo, hi);
                                if (r == FDTD_OK && hi2 > lo2) r = DISPATCH(s, launch_fused, s, n_half, lo2, hi2);
                                return r;
                            },
                            [&](cudaStream_t q, bool, cudaEvent_t, cudaEvent_t) { return exchange_fused(s, q); });
            if (st != FDTD_OK) return st;
        }
        s->cur ^= 1;
    } else {
        st = exchange_E_top(s);
        if (st != FDTD_OK) return st;
        st = DISPATCH_A(s, launch_sweep, s, true, n_half, 1);
        if (st != FDTD_OK) return st;
        s->ghosts_b_valid = false;
        st = exchange_B_bottom(s);
        if (st != FDTD_OK) return st;
        st = DISPATCH_A(s, launch_sweep, s, false, 0, 0);
        if (st != FDTD_OK) return st;
    }
    invalidate_ghosts(s);
    s->b_pending = true;
    s->steps_done += *done;
    return FDTD_OK;
}

// ------------------------------------------------------------------------------------------------------
// creation / destruction
// ------------------------------------------------------------------------------------------------------
static void destroy_impl(Solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    nccl_destroy(s);
    for (int c = 0; c < NCOMP; ++c)
        for (int gdx = 0; gdx < 2; ++gdx)
            if (s->base[c][gdx]) cudaFree(s->base[c][gdx]);
    for (int c = 0; c < 2 * NSPLIT; ++c)
        if (s->split_base[c]) cudaFree(s->split_base[c]);
    for (int a = 0; a < 3; ++a) {
        if (s->d_decay[a]) cudaFree(s->d_decay[a]);
        if (s->d_coef2[a]) cudaFree(s->d_coef2[a]);
        if (s->d_w[a]) cudaFree(s->d_w[a]);
    }
    if (s->d_stage) cudaFree(s->d_stage);
    if (s->d_jpend) cudaFree(s->d_jpend);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    if (s->ev_t0) cudaEventDestroy(s->ev_t0);
    if (s->ev_t1) cudaEventDestroy(s->ev_t1);
    if (s->ev_a) cudaEventDestroy(s->ev_a);
    if (s->ev_b) cudaEventDestroy(s->ev_b);
    if (s->ev_c) cudaEventDestroy(s->ev_c);
    for (cudaEvent_t e : s->tl_events) cudaEventDestroy(e);
    if (s->bnd_stream) cudaStreamDestroy(s->bnd_stream);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

static fdtd_status_t alloc_array(Solver* s, void** base, void** p) {
    const size_t bytes = (size_t)s->g.plane * (size_t)(s->g.nk + 2 * GHOST_PLANES) * s->esz;
    cudaError_t e = cudaMalloc(base, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(FDTD_ERR_NOMEM, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    FDTD_CUDA_TRY(cudaMemsetAsync(*base, 0, bytes, s->stream));   // FDTD.cpp:21-32: all fields start at zero
    *p = static_cast<char*>(*base) + (size_t)GHOST_PLANES * s->g.plane * s->esz;
    s->device_bytes += (int64_t)bytes;
    return FDTD_OK;
}

static fdtd_status_t create_impl(const fdtd_config_t* cfg, Solver** out) {
    const fdtd_params_t& P = cfg->grid;
    // FDTD.cpp:5-7
    if (P.Ni <= 0 || P.Nj <= 0 || P.Nk <= 0 || !(cfg->dt > 0)) return fail(FDTD_ERR_INVALID_PARAMETERS, "ERROR: invalid parameters");
    if (cfg->dtype != FDTD_F64 && cfg->dtype != FDTD_F32) return fail(FDTD_ERR_BAD_ARGUMENT, "unknown dtype");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(FDTD_ERR_BAD_ARGUMENT, "bad rank / nranks");
    if (cfg->nranks > P.Nk) return fail(FDTD_ERR_BAD_ARGUMENT, "more ranks than k planes");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(FDTD_ERR_CUDA, "no CUDA device available: libfdtd_b200 has no CPU fallback");
    }
    int dev = cfg->device;
    if (dev < 0) FDTD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(FDTD_ERR_BAD_ARGUMENT, "device ordinal out of range");
    FDTD_CUDA_TRY(cudaSetDevice(dev));

    Solver* s = new (std::nothrow) Solver();
    if (!s) return fail(FDTD_ERR_NOMEM, "out of host memory");
    s->cfg = *cfg;
    read_tunables(s->tun);
    { const char* e = std::getenv("FDTD_B200_HALO_IN_KERNEL"); s->halo_in_kernel = !(e && std::atoi(e) == 0); }
    s->device = dev;
    s->dtype = cfg->dtype;
    s->esz = (cfg->dtype == FDTD_F32) ? 4 : 8;
    if (cfg->flags & FDTD_FLAG_F32_ARITH) {
        if (cfg->dtype != FDTD_F32) { delete s; return fail(FDTD_ERR_BAD_ARGUMENT, "FDTD_FLAG_F32_ARITH needs dtype FDTD_F32"); }
        s->f32_arith = true;
    }

    int k_begin, k_end;
    fdtd_slab_range_cfg(cfg, cfg->rank, &k_begin, &k_end);
    int min_slab = P.Nk;   // the smallest slab of the ring: decides, identically on every rank, which passes are used
    for (int r = 0; r < cfg->nranks; ++r) {
        int b, e;
        fdtd_slab_range_cfg(cfg, r, &b, &e);
        min_slab = std::min(min_slab, e - b);
    }
    if (min_slab < 1) { delete s; return fail(FDTD_ERR_BAD_ARGUMENT, "more ranks than k planes"); }
    s->g.Ni = P.Ni; s->g.Nj = P.Nj; s->g.Nk = P.Nk;
    s->g.nk = k_end - k_begin;
    s->g.k0 = k_begin;
    s->g.wrap_k = (cfg->nranks == 1) ? 1 : 0;
    s->g.pitch = (long long)round_up((size_t)P.Ni, 128 / s->esz);
    s->g.plane = s->g.pitch * P.Nj;

    // FDTD.cpp:43-53
    const double cdt = kC * cfg->dt;
    s->c.cEx = cdt / P.dx; s->c.cEy = cdt / P.dy; s->c.cEz = cdt / P.dz;
    s->c.cBx = cdt / (2.0 * P.dx); s->c.cBy = cdt / (2.0 * P.dy); s->c.cBz = cdt / (2.0 * P.dz);
    s->c.cJ = -4.0 * kPI * cfg->dt;

    const int N[3] = {P.Ni, P.Nj, P.Nk};
    const double d[3] = {P.dx, P.dy, P.dz};
    for (int a = 0; a < 3; ++a) { s->main_lo[a] = 0; s->main_hi[a] = N[a]; }
    if (cfg->pml_mode != FDTD_PML_NONE) {
        s->has_pml = true;
        for (int a = 0; a < 3; ++a) {
            s->pml[a] = (cfg->pml_mode == FDTD_PML_PERCENT) ? fdtd_pml_thickness(N[a], cfg->pml_percent) : cfg->pml_thickness[a];
            if (s->pml[a] < 0 || 2 * s->pml[a] > N[a]) {
                delete s;
                return fail(FDTD_ERR_INVALID_PARAMETERS, "ERROR: invalid parameters (PML thicker than half the grid)");
            }
            s->main_lo[a] = s->pml[a];          // FDTD_PML.cpp:260-269
            s->main_hi[a] = N[a] - s->pml[a];
        }
    }

    fdtd_status_t st = FDTD_OK;
    auto bail = [&](fdtd_status_t code) { std::string keep = g_err; destroy_impl(s); g_err = keep; return code; };

    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaStreamCreate"));
    {
        int prio_lo = 0, prio_hi = 0;   // numerically lower = higher priority
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (s->tun.mgpu_debug & 8) prio_hi = prio_lo;   // bit 3: no priority
        if (cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaStreamCreate"));
    }
    if (cudaStreamCreateWithFlags(&s->bnd_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaStreamCreate"));
    s->launch_stream = s->stream;
    if (cudaEventCreate(&s->ev_t0) != cudaSuccess || cudaEventCreate(&s->ev_t1) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_b, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_c, cudaEventDisableTiming) != cudaSuccess)
        return bail(cuda_fail(cudaGetLastError(), "cudaEventCreate"));

    // The fused pass serves the periodic solver with vector-aligned rows; everything else runs the two sweeps.
    const int V = (int)(16 / s->esz);
    s->fused = !s->has_pml && !(cfg->flags & FDTD_FLAG_NO_FUSION) && (P.Ni % V == 0);
    // (decided from the smallest slab of the ring, not the local one: every rank must take the same path, or the
    // exchange plans of neighbours do not match -- fdtd_slab_range gives the remainder planes to the low ranks)
    s->t2 = s->fused && !(cfg->flags & FDTD_FLAG_NO_TEMPORAL) && min_slab >= 4;
    // fp64 by default: the fp32 T2 pass is conversion-bound and loses to the two lean interior sweeps
    // (measured 5.5 vs 4.8 ms per step at 512^3); FDTD_B200_PML_T2_F32=1 turns it on anyway (parity tests do).
    const bool pair_dtype_ok = (s->esz == 8) || s->tun.pml_t2_f32;
    if (s->has_pml && pair_dtype_ok && !(cfg->flags & (FDTD_FLAG_NO_FUSION | FDTD_FLAG_NO_TEMPORAL)) && (P.Ni % V == 0) && min_slab >= 4) {
        // store box of the two-step pass: main box shrunk by the pass's reach on the axes that have a shell
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            const int d = s->pml[a] > 0 ? 2 : 0;
            s->sb_lo[a] = s->main_lo[a] + d;
            s->sb_hi[a] = s->main_hi[a] - d;
        }
        s->sb_lo[0] = (s->sb_lo[0] + V - 1) / V * V;
        s->sb_hi[0] = s->sb_hi[0] / V * V;
        // worth it only when the core is a real volume (and the rim boxes below stay non-degenerate)
        ok = (s->sb_hi[0] - s->sb_lo[0] >= 8 * V) && (s->sb_hi[1] - s->sb_lo[1] >= 8) && (s->sb_hi[2] - s->sb_lo[2] >= 8);
        s->pml_t2 = ok;
    }

    for (int c = 0; c < NCOMP && st == FDTD_OK; ++c) {
        st = alloc_array(s, &s->base[c][0], &s->p[c][0]);
        if (st == FDTD_OK && (s->fused || s->pml_t2) && c < JX) st = alloc_array(s, &s->base[c][1], &s->p[c][1]);
    }
    if (st != FDTD_OK) return bail(st);
    if (s->has_pml) {
        for (int c = 0; c < 2 * NSPLIT && st == FDTD_OK; ++c) st = alloc_array(s, &s->split_base[c], &s->split_p[c]);   // FDTD_PML.cpp:209-252
        if (st != FDTD_OK) return bail(st);
        for (int a = 0; a < 3; ++a) {
            std::vector<double> dec(N[a]), c2(N[a]);
            pml_profile_host(N[a], s->pml[a], d[a], cfg->dt, nullptr, dec.data(), c2.data());
            if (cudaMalloc(&s->d_decay[a], sizeof(double) * N[a]) != cudaSuccess ||
                cudaMalloc(&s->d_coef2[a], sizeof(double) * N[a]) != cudaSuccess)
                return bail(cuda_fail(cudaGetLastError(), "cudaMalloc(pml tables)"));
            // on the solver's own (non-blocking) stream: the kernels that read the tables are ordered after the copies
            // (pageable source: the call returns once the data is staged, the vectors may die at the end of this scope)
            if (cudaMemcpyAsync(s->d_decay[a], dec.data(), sizeof(double) * N[a], cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
                cudaMemcpyAsync(s->d_coef2[a], c2.data(), sizeof(double) * N[a], cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess)
                return bail(cuda_fail(cudaGetLastError(), "cudaMemcpy(pml tables)"));
        }
    }
    jbox_clear(s);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "initial zero fill"));
    *out = s;
    return FDTD_OK;
}

static fdtd_status_t check_handle(fdtd_solver_t* h, Solver** s) {
    if (!h) return fail(FDTD_ERR_BAD_ARGUMENT, "null solver handle");
    *s = reinterpret_cast<Solver*>(h);
    cudaError_t e = cudaSetDevice((*s)->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return FDTD_OK;
}

static fdtd_status_t run_steps(Solver* s, int nsteps) {
    fdtd_status_t st = FDTD_OK;
    for (int t = 0; t < nsteps;) {
        int done = 1;
        st = advance(s, nsteps - t, &done);
        if (st != FDTD_OK) return st;
        t += done;
        // pending J writes belong to the step after the first one issued here: either a two-step pass has just consumed the
        // box in its second stage, or a single step has run on the old J -- in both cases the arrays take the writes now
        if (s->jpend && (st = apply_pending_J(s)) != FDTD_OK) return st;
    }
    return materialize_J(s);
}

// Run the update_fields() call that fdtd_update_fields() recorded but did not issue (see Solver::lazy_steps).
static fdtd_status_t flush_lazy(Solver* s) {
    if (!s->lazy_steps) return FDTD_OK;
    const int n = s->lazy_steps;
    s->lazy_steps = 0;
    return run_steps(s, n);
}

// Entry of every call that reads or changes solver state: the handle, then the recorded step.
static fdtd_status_t enter(fdtd_solver_t* h, Solver** s) {
    fdtd_status_t st = check_handle(h, s);
    if (st != FDTD_OK) return st;
    return flush_lazy(*s);
}

static fdtd_status_t check_component(int comp) {
    if (comp < EX || comp > JZ) return fail(FDTD_ERR_INVALID_COMPONENT, "ERROR: Invalid field component");   // FDTD.cpp:149
    return FDTD_OK;
}

static fdtd_status_t ensure_stage(Solver* s, size_t bytes) {
    if (bytes <= s->stage_bytes) return FDTD_OK;
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->d_stage) cudaFree(s->d_stage);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    s->d_stage = nullptr; s->h_stage = nullptr; s->stage_bytes = 0;
    size_t cap = round_up(bytes < 4096 ? 4096 : bytes * 2, 256);
    FDTD_CUDA_TRY(cudaMalloc(&s->d_stage, cap));
    FDTD_CUDA_TRY(cudaMallocHost(&s->h_stage, cap));
    s->stage_bytes = cap;
    return FDTD_OK;
}

template <typename T>
static fdtd_status_t scatter_impl(Solver* s, int comp, const int64_t* idx, const void* vals, size_t n) {
    const size_t ib = round_up(n * sizeof(long long), 256), vb = n * sizeof(T);
    fdtd_status_t st = ensure_stage(s, ib + vb);
    if (st != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));   // staging buffer may still be in flight
    std::memcpy(s->h_stage, idx, n * sizeof(long long));
    std::memcpy(static_cast<char*>(s->h_stage) + ib, vals, vb);
    FDTD_CUDA_TRY(cudaMemcpyAsync(s->d_stage, s->h_stage, ib + vb, cudaMemcpyHostToDevice, s->stream));
    const int threads = 128, blocks = (int)((n + threads - 1) / threads);
    scatter_kernel<T><<<blocks, threads, 0, s->stream>>>(static_cast<T*>(s->cur_ptr(comp)), s->g,
                                                        static_cast<const long long*>(s->d_stage),
                                                        reinterpret_cast<const T*>(static_cast<char*>(s->d_stage) + ib), (int)n);
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FDTD_OK;
}

// fdtd_scatter of a J component into the pending box (Solver::jpend) instead of the array.
template <typename T>
static fdtd_status_t scatter_pending_impl(Solver* s, int comp, const int64_t* idx, const void* vals, size_t n) {
    const size_t ib = round_up(n * sizeof(long long), 256), vb = n * sizeof(T);
    fdtd_status_t st = ensure_stage(s, ib + vb);
    if (st != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));   // staging buffer may still be in flight
    std::memcpy(s->h_stage, idx, n * sizeof(long long));
    std::memcpy(static_cast<char*>(s->h_stage) + ib, vals, vb);
    FDTD_CUDA_TRY(cudaMemcpyAsync(s->d_stage, s->h_stage, ib + vb, cudaMemcpyHostToDevice, s->stream));
    return jbox_launch<T>(s, comp, 2, static_cast<const long long*>(s->d_stage),
                          reinterpret_cast<const T*>(static_cast<char*>(s->d_stage) + ib), (int)n);
}

template <typename T>
static fdtd_status_t gather_impl(Solver* s, int comp, const int64_t* idx, void* vals, size_t n) {
    const size_t ib = round_up(n * sizeof(long long), 256), vb = n * sizeof(T);
    fdtd_status_t st = ensure_stage(s, ib + vb);
    if (st != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    std::memcpy(s->h_stage, idx, n * sizeof(long long));
    std::memcpy(static_cast<char*>(s->h_stage) + ib, vals, vb);   // entries not owned by this rank stay untouched
    FDTD_CUDA_TRY(cudaMemcpyAsync(s->d_stage, s->h_stage, ib + vb, cudaMemcpyHostToDevice, s->stream));
    const int threads = 128, blocks = (int)((n + threads - 1) / threads);
    gather_kernel<T><<<blocks, threads, 0, s->stream>>>(static_cast<const T*>(s->cur_ptr(comp)), s->g,
                                                       static_cast<const long long*>(s->d_stage),
                                                       reinterpret_cast<T*>(static_cast<char*>(s->d_stage) + ib), (int)n);
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    FDTD_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(s->h_stage) + ib, static_cast<char*>(s->d_stage) + ib, vb,
                                  cudaMemcpyDeviceToHost, s->stream));
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    std::memcpy(vals, static_cast<char*>(s->h_stage) + ib, vb);
    return FDTD_OK;
}

template <typename T>
static fdtd_status_t slice_impl(Solver* s, int comp, int axis, int local_index, int n0, int n1, void* host) {
    const size_t bytes = (size_t)n0 * n1 * sizeof(T);
    fdtd_status_t st = ensure_stage(s, bytes);
    if (st != FDTD_OK) return st;
    const long long total = (long long)n0 * n1;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    slice_kernel<T><<<(int)blocks, 256, 0, s->stream>>>(static_cast<const T*>(s->cur_ptr(comp)), s->g, axis, local_index, n0, n1,
                                                        static_cast<T*>(s->d_stage));
    FDTD_CUDA_TRY(cudaGetLastError());
    s->launches++;
    FDTD_CUDA_TRY(cudaMemcpyAsync(s->h_stage, s->d_stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    std::memcpy(host, s->h_stage, bytes);
    return FDTD_OK;
}

}  // namespace fdtd_b200

// ======================================================================================================
// C ABI
// ======================================================================================================
using namespace fdtd_b200;

extern "C" {

const char* fdtd_last_error(void) { return g_err.c_str(); }
int fdtd_version(void) { return FDTD_B200_VERSION; }

int fdtd_pml_thickness(int N, double pml_percent) {
    return static_cast<int>(static_cast<double>(N) * pml_percent);   // FDTD_PML.cpp:254-256
}

fdtd_status_t fdtd_pml_profile(int N, int thickness, double d, double dt, double* sigma, double* decay, double* coef2) {
    if (N <= 0 || thickness < 0 || 2 * thickness > N || !decay || !coef2) return fail(FDTD_ERR_BAD_ARGUMENT, "bad PML profile request");
    pml_profile_host(N, thickness, d, dt, sigma, decay, coef2);
    return FDTD_OK;
}

void fdtd_slab_range(int Nk, int rank, int nranks, int* k_begin, int* k_end) {
    const int base = Nk / nranks, rem = Nk % nranks;
    const int b = rank * base + (rank < rem ? rank : rem);
    if (k_begin) *k_begin = b;
    if (k_end) *k_end = b + base + (rank < rem ? 1 : 0);
}

// Cost-weighted z slabs for PML solvers (SURVEY.md 8(e) "PML interaction").  A shell cell moves 36 words per step through
// the rim sweeps, a core cell 6 through the two-step pass (measured 512^3 / 32: 120 vs 21 ps per cell-pair,
// profiles/ncu_pml_r01.md), so a plane inside the k shell costs Ni*Nj*W and a plane of the main k range
// shell_ij*W + core_ij with W = 6; without a k shell (or one rank, or FDTD_FLAG_UNIFORM_SLABS) the split is the uniform one.
// Boundaries are the points where the running cost crosses r/nranks of the total (every rank keeps >= 1 plane).
void fdtd_slab_range_cfg(const fdtd_config_t* cfg, int rank, int* k_begin, int* k_end) {
    const fdtd_params_t& P = cfg->grid;
    const int n = cfg->nranks < 1 ? 1 : cfg->nranks;
    int pml[3] = {0, 0, 0};
    if (cfg->pml_mode == FDTD_PML_PERCENT) {
        pml[0] = fdtd_pml_thickness(P.Ni, cfg->pml_percent); pml[1] = fdtd_pml_thickness(P.Nj, cfg->pml_percent); pml[2] = fdtd_pml_thickness(P.Nk, cfg->pml_percent);
    } else if (cfg->pml_mode == FDTD_PML_THICKNESS) {
        for (int a = 0; a < 3; ++a) pml[a] = cfg->pml_thickness[a];
    }
    const bool weighted = n > 1 && pml[2] > 0 && 2 * pml[2] <= P.Nk && 2 * pml[0] <= P.Ni && 2 * pml[1] <= P.Nj &&
                          !(cfg->flags & FDTD_FLAG_UNIFORM_SLABS) && P.Nk >= 8 * n;
    if (!weighted) { fdtd_slab_range(P.Nk, rank, n, k_begin, k_end); return; }
    const double W = 6.0;
    const double core_ij = (double)(P.Ni - 2 * pml[0]) * (double)(P.Nj - 2 * pml[1]);
    const double shell_plane = (double)P.Ni * P.Nj * W, main_plane = ((double)P.Ni * P.Nj - core_ij) * W + core_ij;
    const double total = 2.0 * pml[2] * shell_plane + (double)(P.Nk - 2 * pml[2]) * main_plane;
    auto boundary = [&](int r) -> int {   // first plane of rank r
        if (r <= 0) return 0;
        if (r >= n) return P.Nk;
        const double target = total * r / n;
        double acc = 0.0;
        int k = 0;
        for (; k < P.Nk; ++k) {
            const double c = (k < pml[2] || k >= P.Nk - pml[2]) ? shell_plane : main_plane;
            if (acc + 0.5 * c >= target) break;
            acc += c;
        }
        // keep at least 4 planes per rank (the two-step pass), whatever the weights say
        const int lo = 4 * r, hi = P.Nk - 4 * (n - r);
        return k < lo ? lo : (k > hi ? hi : k);
    };
    if (k_begin) *k_begin = boundary(rank);
    if (k_end) *k_end = boundary(rank + 1);
}

void fdtd_config_init(fdtd_config_t* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->dtype = FDTD_F64;
    cfg->pml_mode = FDTD_PML_NONE;
    cfg->device = -1;
    cfg->rank = 0;
    cfg->nranks = 1;
}

fdtd_status_t fdtd_create_ex(const fdtd_config_t* cfg, fdtd_solver_t** out) {
    if (!cfg || !out) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    if (cfg->struct_size != sizeof(fdtd_config_t)) return fail(FDTD_ERR_BAD_ARGUMENT, "fdtd_config_t size mismatch (use fdtd_config_init)");
    Solver* s = nullptr;
    fdtd_status_t st = create_impl(cfg, &s);
    if (st != FDTD_OK) return st;
    *out = reinterpret_cast<fdtd_solver_t*>(s);
    return FDTD_OK;
}

fdtd_status_t fdtd_create(const fdtd_params_t* params, double dt, fdtd_solver_t** out) {
    if (!params) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    fdtd_config_t cfg;
    fdtd_config_init(&cfg);
    cfg.grid = *params;
    cfg.dt = dt;
    return fdtd_create_ex(&cfg, out);
}

fdtd_status_t fdtd_create_pml(const fdtd_params_t* params, double dt, double pml_percent, fdtd_solver_t** out) {
    if (!params) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    fdtd_config_t cfg;
    fdtd_config_init(&cfg);
    cfg.grid = *params;
    cfg.dt = dt;
    cfg.pml_mode = FDTD_PML_PERCENT;
    cfg.pml_percent = pml_percent;
    return fdtd_create_ex(&cfg, out);
}

fdtd_status_t fdtd_destroy(fdtd_solver_t* h) {
    if (!h) return FDTD_OK;
    destroy_impl(reinterpret_cast<Solver*>(h));
    return FDTD_OK;
}

fdtd_status_t fdtd_step(fdtd_solver_t* h, int nsteps) {
    Solver* s;
    fdtd_status_t st = check_handle(h, &s);
    if (st != FDTD_OK) return st;
    if (nsteps < 0) return fail(FDTD_ERR_BAD_ARGUMENT, "negative step count");
    const int n = nsteps + s->lazy_steps;   // a recorded update_fields() call joins the batch
    s->lazy_steps = 0;
    return run_steps(s, n);
}

// FDTD::update_fields().  The reference's callers step one call at a time (sample.cpp:57-87, test_FDTD_method.cpp:32-41);
// so that such loops still reach the two-step pass, an odd call is recorded and returns at once, and the next call
// issues both steps as one pass.  Anything that reads or changes state (field access, J writes, sources, sync, timers)
// runs the recorded step first, so the observable sequence is exactly one step per call (errors of a deferred step
// surface at that later call).  FDTD_B200_NO_LAZY=1 issues every call immediately.
fdtd_status_t fdtd_update_fields(fdtd_solver_t* h) {
    Solver* s;
    fdtd_status_t st = check_handle(h, &s);
    if (st != FDTD_OK) return st;
    if (s->lazy_steps) {
        s->lazy_steps = 0;
        return run_steps(s, 2);
    }
    if ((s->t2 || s->pml_t2) && !s->tun.no_t2 && !s->tun.no_lazy) {
        s->lazy_steps = 1;
        return FDTD_OK;
    }
    return run_steps(s, 1);
}

fdtd_status_t fdtd_zeroed_currents(fdtd_solver_t* h) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    s->src_active = false;
    return zero_currents_impl(s);
}

fdtd_status_t fdtd_upload(fdtd_solver_t* h, int comp, const void* host, size_t count) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = check_component(comp)) != FDTD_OK) return st;
    const size_t expect = (size_t)s->g.Ni * s->g.Nj * s->g.nk;
    if (!host || count != expect) return fail(FDTD_ERR_BAD_ARGUMENT, "upload: count must equal Ni*Nj*(k_end-k_begin)");
    if (comp < JX) {
        if ((st = flush_pending(s)) != FDTD_OK) return st;
        invalidate_ghosts(s);
    } else {
        jbox_full(s);
    }
    if (s->g.pitch == s->g.Ni)   // unpadded rows: the slab is one contiguous block, like the reference's flat vector
        FDTD_CUDA_TRY(cudaMemcpyAsync(s->cur_ptr(comp), host, expect * s->esz, cudaMemcpyHostToDevice, s->stream));
    else
        FDTD_CUDA_TRY(cudaMemcpy2DAsync(s->cur_ptr(comp), (size_t)s->g.pitch * s->esz, host, (size_t)s->g.Ni * s->esz,
                                        (size_t)s->g.Ni * s->esz, (size_t)s->g.Nj * s->g.nk, cudaMemcpyHostToDevice, s->stream));
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return FDTD_OK;
}

fdtd_status_t fdtd_download(fdtd_solver_t* h, int comp, void* host, size_t count) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = check_component(comp)) != FDTD_OK) return st;
    const size_t expect = (size_t)s->g.Ni * s->g.Nj * s->g.nk;
    if (!host || count != expect) return fail(FDTD_ERR_BAD_ARGUMENT, "download: count must equal Ni*Nj*(k_end-k_begin)");
    // only B carries a deferred half step; E(n+1) is already final after the pass
    if (comp >= BX && comp < JX && (st = flush_pending(s)) != FDTD_OK) return st;
    if (s->g.pitch == s->g.Ni)
        FDTD_CUDA_TRY(cudaMemcpyAsync(host, s->cur_ptr(comp), expect * s->esz, cudaMemcpyDeviceToHost, s->stream));
    else
        FDTD_CUDA_TRY(cudaMemcpy2DAsync(host, (size_t)s->g.Ni * s->esz, s->cur_ptr(comp), (size_t)s->g.pitch * s->esz,
                                        (size_t)s->g.Ni * s->esz, (size_t)s->g.Nj * s->g.nk, cudaMemcpyDeviceToHost, s->stream));
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return FDTD_OK;
}

fdtd_status_t fdtd_scatter(fdtd_solver_t* h, int comp, const int64_t* idx, const void* values, size_t n) {
    Solver* s;
    fdtd_status_t st = check_handle(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = check_component(comp)) != FDTD_OK) return st;
    if (n == 0) return flush_lazy(s);
    if (!idx || !values) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    const long long total = (long long)s->g.Ni * s->g.Nj * s->g.Nk, ij = (long long)s->g.Ni * s->g.Nj;
    int lo[3] = {s->g.Ni, s->g.Nj, s->g.Nk}, hi[3] = {0, 0, 0};
    for (size_t t = 0; t < n; ++t) {
        if (idx[t] < 0 || idx[t] >= total) return fail(FDTD_ERR_BAD_ARGUMENT, "scatter: index out of range");
        const int k = (int)(idx[t] / ij), j = (int)((idx[t] % ij) / s->g.Ni), i = (int)(idx[t] % s->g.Ni);
        const int c3[3] = {i, j, k};
        for (int a = 0; a < 3; ++a) { if (c3[a] < lo[a]) lo[a] = c3[a]; if (c3[a] + 1 > hi[a]) hi[a] = c3[a] + 1; }
    }
    // J writes while one update_fields() call is recorded: keep them as a pending box so that the next call can still pair
    // (a small box, no device source, every write inside the box the first of them opened)
    if (comp >= JX && s->lazy_steps == 1 && !s->src_active && !s->tun.no_t2 && n <= (size_t)Solver::JPEND_MAX_CELLS) {
        const long long vol = (long long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
        bool inside = s->jpend;
        for (int a = 0; a < 3 && inside; ++a) inside = lo[a] >= s->jp_lo[a] && hi[a] <= s->jp_hi[a];
        if (inside || (!s->jpend && vol <= Solver::JPEND_MAX_CELLS)) {
            if (!s->jpend) {
                if (!s->d_jpend) FDTD_CUDA_TRY(cudaMalloc(&s->d_jpend, (size_t)3 * Solver::JPEND_MAX_CELLS * s->esz));
                for (int a = 0; a < 3; ++a) { s->jp_lo[a] = lo[a]; s->jp_hi[a] = hi[a]; }
                for (int c = JX; c <= JZ; ++c) {   // the boxes start as copies of the arrays
                    st = (s->dtype == FDTD_F32) ? jbox_launch<float>(s, c, 0, nullptr, nullptr, 0) : jbox_launch<double>(s, c, 0, nullptr, nullptr, 0);
                    if (st != FDTD_OK) return st;
                }
                s->jpend = true;
                jbox_union(s, lo, hi);   // J may be non-zero there from the next step on (a larger box only costs skipped skips)
            }
            return (s->dtype == FDTD_F32) ? scatter_pending_impl<float>(s, comp, idx, values, n) : scatter_pending_impl<double>(s, comp, idx, values, n);
        }
    }
    if ((st = flush_lazy(s)) != FDTD_OK) return st;
    if (comp < JX) {
        if ((st = flush_pending(s)) != FDTD_OK) return st;
        invalidate_ghosts(s);
    } else {
        jbox_union(s, lo, hi);
    }
    return (s->dtype == FDTD_F32) ? scatter_impl<float>(s, comp, idx, values, n) : scatter_impl<double>(s, comp, idx, values, n);
}

fdtd_status_t fdtd_gather(fdtd_solver_t* h, int comp, const int64_t* idx, void* values, size_t n) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = check_component(comp)) != FDTD_OK) return st;
    if (n == 0) return FDTD_OK;
    if (!idx || !values) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    const long long total = (long long)s->g.Ni * s->g.Nj * s->g.Nk;
    for (size_t t = 0; t < n; ++t)
        if (idx[t] < 0 || idx[t] >= total) return fail(FDTD_ERR_BAD_ARGUMENT, "gather: index out of range");
    if (comp >= BX && comp < JX && (st = flush_pending(s)) != FDTD_OK) return st;
    return (s->dtype == FDTD_F32) ? gather_impl<float>(s, comp, idx, values, n) : gather_impl<double>(s, comp, idx, values, n);
}

fdtd_status_t fdtd_read_slice(fdtd_solver_t* h, int comp, int axis, int index, void* host, size_t capacity, size_t* count_out) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = check_component(comp)) != FDTD_OK) return st;
    if (count_out) *count_out = 0;
    const int N[3] = {s->g.Ni, s->g.Nj, s->g.Nk};
    if (axis < 0 || axis > 2 || index < 0 || index >= N[axis]) return fail(FDTD_ERR_BAD_ARGUMENT, "read_slice: axis / index out of range");
    // Collective part first (every rank of a multi-rank solver must get here, whoever owns the plane): the deferred B
    // half step needs the Ex, Ey ring exchange.  Only then the ownership test.
    if (comp >= BX && comp < JX && (st = flush_pending(s)) != FDTD_OK) return st;
    if (comp >= JX && (st = materialize_J(s)) != FDTD_OK) return st;
    int n0, n1, local = index;
    if (axis == 2) {
        n0 = s->g.Ni; n1 = s->g.Nj; local = index - s->g.k0;
        if (local < 0 || local >= s->g.nk) return FDTD_OK;   // another rank owns this plane: nothing to copy here
    } else {
        n0 = (axis == 1) ? s->g.Ni : s->g.Nj; n1 = s->g.nk;
    }
    const size_t n = (size_t)n0 * n1;
    if (!host || capacity < n) return fail(FDTD_ERR_BAD_ARGUMENT, "read_slice: host buffer too small");
    st = (s->dtype == FDTD_F32) ? slice_impl<float>(s, comp, axis, local, n0, n1, host) : slice_impl<double>(s, comp, axis, local, n0, n1, host);
    if (st == FDTD_OK && count_out) *count_out = n;
    return st;
}

fdtd_status_t fdtd_set_source(fdtd_solver_t* h, const int lo[3], const int hi[3], const double* wx, const double* wy,
                              const double* wz, const double* amp, int n_amp) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if (!lo || !hi || !wx || !wy || !wz || !amp || n_amp < 0) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    const int N[3] = {s->g.Ni, s->g.Nj, s->g.Nk};
    for (int a = 0; a < 3; ++a)
        if (lo[a] < 0 || hi[a] > N[a] || lo[a] > hi[a]) return fail(FDTD_ERR_BAD_ARGUMENT, "source box outside the grid");
    const double* w[3] = {wx, wy, wz};
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (int a = 0; a < 3; ++a) {
        if (s->d_w[a]) { cudaFree(s->d_w[a]); s->d_w[a] = nullptr; }
        const int n = hi[a] - lo[a];
        if (n > 0) {
            FDTD_CUDA_TRY(cudaMalloc(&s->d_w[a], sizeof(double) * n));
            FDTD_CUDA_TRY(cudaMemcpyAsync(s->d_w[a], w[a], sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        }
        s->src_lo[a] = lo[a]; s->src_hi[a] = hi[a];
    }
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));   // the caller's tables may be freed after this call
    s->src_amp.assign(amp, amp + n_amp);
    s->src_t = 0;
    s->src_active = true;
    return FDTD_OK;
}

fdtd_status_t fdtd_clear_source(fdtd_solver_t* h) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    s->src_active = false;
    return FDTD_OK;
}

fdtd_status_t fdtd_issue(fdtd_solver_t* h) {
    Solver* s;
    return enter(h, &s);
}

fdtd_status_t fdtd_flush(fdtd_solver_t* h) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = flush_pending(s)) != FDTD_OK) return st;
    return materialize_J(s);
}

fdtd_status_t fdtd_sync(fdtd_solver_t* h) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = flush_pending(s)) != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (unsigned* err = peer_error_word(s)) {
        unsigned v = 0;
        FDTD_CUDA_TRY(cudaMemcpy(&v, err, sizeof(v), cudaMemcpyDeviceToHost));
        if (v) return fail(FDTD_ERR_STATE, "halo wait timed out inside the pass kernel (a neighbour rank never pushed its planes): fields are invalid");
    }
    return FDTD_OK;
}

fdtd_status_t fdtd_device_ptr(fdtd_solver_t* h, int comp, void** dptr) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if ((st = check_component(comp)) != FDTD_OK) return st;
    if (!dptr) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    if ((st = flush_pending(s)) != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (comp < JX) invalidate_ghosts(s); else jbox_full(s);   // the caller may write through the pointer
    *dptr = s->cur_ptr(comp);
    return FDTD_OK;
}

fdtd_status_t fdtd_get_info(fdtd_solver_t* h, fdtd_info_t* info) {
    if (!h || !info) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    Solver* s = reinterpret_cast<Solver*>(h);
    std::memset(info, 0, sizeof(*info));
    info->Ni = s->g.Ni; info->Nj = s->g.Nj; info->Nk = s->g.Nk;
    info->k_begin = s->g.k0; info->k_end = s->g.k0 + s->g.nk;
    info->dtype = s->dtype; info->has_pml = s->has_pml ? 1 : 0;
    for (int a = 0; a < 3; ++a) info->pml_thickness[a] = s->pml[a];
    info->pitch = s->g.pitch; info->plane = s->g.plane;
    info->device_bytes = s->device_bytes;
    info->launches = s->launches;
    info->steps_done = s->steps_done + s->lazy_steps;   // a recorded update_fields() call counts: it is observably done
    info->fused = s->fused ? 1 : 0;
    info->rank = s->cfg.rank; info->nranks = s->cfg.nranks; info->device = s->device;
    info->temporal = (s->t2 || s->pml_t2) ? 1 : 0;
    info->passes_t2 = s->passes_t2;
    info->transport = s->cfg.nranks <= 1 ? 0 : (s->peer ? 2 : (s->comm ? 1 : 0));
    info->halo_in_kernel = (s->peer && s->halo_in_kernel) ? 1 : 0;
    info->f32_arith = s->f32_arith ? 1 : 0;
    return FDTD_OK;
}

fdtd_status_t fdtd_timer_start(fdtd_solver_t* h) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaEventRecord(s->ev_t0, s->stream));
    return FDTD_OK;
}

fdtd_status_t fdtd_timer_stop(fdtd_solver_t* h, double* elapsed_ms) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    FDTD_CUDA_TRY(cudaEventRecord(s->ev_t1, s->stream));
    FDTD_CUDA_TRY(cudaEventSynchronize(s->ev_t1));
    float ms = 0.f;
    FDTD_CUDA_TRY(cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1));
    if (elapsed_ms) *elapsed_ms = (double)ms;
    return FDTD_OK;
}

fdtd_status_t fdtd_get_stream(fdtd_solver_t* h, void** stream) {
    if (!h || !stream) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    Solver* s;
    fdtd_status_t st = enter(h, &s);   // the caller is about to order its own work after ours
    if (st != FDTD_OK) return st;
    *stream = s->stream;
    return FDTD_OK;
}

fdtd_status_t fdtd_nccl_unique_id(void* id_out, size_t capacity) { return nccl_unique_id(id_out, capacity); }

fdtd_status_t fdtd_comm_init(fdtd_solver_t* h, const void* id, size_t id_bytes) {
    Solver* s;
    fdtd_status_t st = check_handle(h, &s);
    if (st != FDTD_OK) return st;
    return nccl_init(s, id, id_bytes);
}

int fdtd_debug_t2_chunk_plan(int nk, int lo, int hi, int lo2, int hi2, int wait_in_kernel, int tiles, int gx, int kc_override,
                             int* chunk_lo, int* chunk_hi, int capacity) {
    if (!chunk_lo || !chunk_hi || capacity < T2_MAXCH || hi <= lo) return -1;
    const T2Ranges rg{lo, hi, lo2, hi2};
    return t2_chunk_plan(nk, rg, wait_in_kernel != 0, tiles, gx, 148, kc_override, chunk_lo, chunk_hi);
}

fdtd_status_t fdtd_comm_init_local(fdtd_solver_t** solvers, int n) {
    if (!solvers || n < 1) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    for (int i = 0; i < n; ++i) if (!solvers[i]) return fail(FDTD_ERR_BAD_ARGUMENT, "null solver handle");
    return peer_ring_init_local(reinterpret_cast<Solver**>(solvers), n);
}

fdtd_status_t fdtd_timeline_enable(fdtd_solver_t* h, int max_passes) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if (max_passes < 0) return fail(FDTD_ERR_BAD_ARGUMENT, "negative pass count");
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (cudaEvent_t e : s->tl_events) cudaEventDestroy(e);
    s->tl_events.clear();
    s->tl_cap = s->tl_n = 0;
    for (int i = 0; i < 4 * max_passes; ++i) {
        cudaEvent_t e;
        FDTD_CUDA_TRY(cudaEventCreate(&e));
        s->tl_events.push_back(e);
    }
    s->tl_cap = max_passes;
    return FDTD_OK;
}

fdtd_status_t fdtd_timeline_read(fdtd_solver_t* h, double* ms, int capacity_passes, int* n_passes) {
    Solver* s;
    fdtd_status_t st = enter(h, &s);
    if (st != FDTD_OK) return st;
    if (!n_passes || (capacity_passes > 0 && !ms)) return fail(FDTD_ERR_BAD_ARGUMENT, "null argument");
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
    FDTD_CUDA_TRY(cudaStreamSynchronize(s->comm_stream));
    const int n = s->tl_n < capacity_passes ? s->tl_n : capacity_passes;
    for (int p = 0; p < n; ++p)
        for (int w = 0; w < 4; ++w) {
            float t = 0.f;
            FDTD_CUDA_TRY(cudaEventElapsedTime(&t, s->tl_events[0], s->tl_events[(size_t)p * 4 + w]));
            ms[p * 4 + w] = (double)t;
        }
    *n_passes = n;
    return FDTD_OK;
}

}  // extern "C"
