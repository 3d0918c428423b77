// fused_kernel_t2.cuh -- temporally blocked fused pass: TWO Yee steps per launch ("T2 pass").
//
// Replaces two consecutive calls of FDTD::update_fields() (reference src/FDTD/FDTD.cpp:153-157 =
// update_B, update_E, update_B; Kokkos twin src/FDTD_kokkos/FDTD_kokkos.cpp:91-104) with ONE kernel:
//
//     stage A (step s)    B1 = round(round(B0 + h(E0)) [+ h(E0)])      n_half = 1 or 2, like fused_BE_kernel
//                         E1 = E0 + g(B1, J_s)
//     stage B (step s+1)  B2 = round(round(B1 + h(E1)) + h(E1))        trailing half of s merged with leading half of s+1
//                         E2 = E1 + g(B2, J_{s+1})
//
// Every cell sees exactly the operations of the reference in the reference's order, so the result is
// bit-identical to two single-step passes (tests/test_parity_gpu.py).  What changes is the HBM traffic:
// read E0(3)+B0(3), write E2(3)+B2(3) = 12 words per cell per TWO steps = 6 words (48 B fp64) per
// cell-step, against 12 for the one-step fused pass and 30 for the reference's three sweeps.
//
// Tile roles (V = cells per 16-byte vector; the dependency cone of one output cell reaches 2 cells down
// and 2 cells up in i, j and k):
//   * warp = 32 lanes x V cells of one row; lane 0 / lane 31 are the left / right halo lanes
//     (2 halo cells fit in one lane for V = 2 and V = 4), lanes 1..30 store -> 30*V cells per row;
//     i-neighbours travel by __shfl_up/down (north_star (b));
//   * CTA = BY warps = BY consecutive rows; rows 0,1 and BY-2,BY-1 are halo rows, rows 2..BY-3 store;
//     j-neighbours travel through four single-buffered shared-memory row exchanges (E0, B1, E1, B2),
//     two __syncthreads per plane;
//   * each thread streams a chunk of k planes upward; stage B runs one plane behind stage A.  Carried in
//     registers: E0(k), B1(k-1), E1(k-1), B2(k-2).xy.  A chunk starts two planes early and ends one
//     plane late (3 redundant plane iterations per chunk);
//   * the k loop is unrolled three times with the register sets rotating (E0(k+1) -> E0/E1(k) -> E1/E2(k-1),
//     B0/B1(k) -> B1/B2(k-1) -> B2(k-2)), so nothing is copied between planes;
//   * the six 16-byte vectors a thread needs per plane (E0(k+1) x3, B0(k) x3) arrive through a 3-slot
//     per-thread cp.async ring in shared memory (LDGSTS, L1 bypass), issued two planes ahead: no register
//     cost for the bytes in flight, and no barrier (each thread reads only what it copied itself).  (Bulk / TMA
//     row copies were measured and lost: 512-byte copies are too small, see DESIGN.md.)
//   * periodic wrap: halo lanes / rows / planes read the wrapped address of the INPUT generation (E and B are
//     double-buffered); on a z-slab rank the k halos are the two ghost planes on each side.
//
// J: stage A reads the J arrays (inside the host-tracked bounding box only).  Stage B reads the same arrays
// (static J) unless a device-resident source is active, in which case the box of that source is evaluated in
// the kernel for step s+1 -- ((amp*wx)*wy)*wz, the product order of perf-tests/sample/sample.cpp:26-31 --
// i.e. source injection as a fused epilogue (north_star (c)).
//
// Requires Ni % V == 0 and nk >= 4 (the host falls back to the one-step pass otherwise).
#pragma once

#include <cuda.h>   // CUtensorMap (the driver entry point is fetched at run time, nothing links against libcuda)

#include "fused_kernel_v2.cuh"

namespace fdtd_b200 {

template <typename T>
struct FusedT2Args {
    Geom g;
    Coefs c;
    const T* Ein[3];
    const T* Bin[3];
    const T* J[3];
    T* Eout[3];
    T* Bout[3];
    JBox jbox;
    int k_lo, k_hi;   // local planes [k_lo, k_hi) produced by this launch ...
    int k_lo2, k_hi2; // ... plus a second range (the other boundary slab of a z-slab rank; empty otherwise)
    int nz1;          // blockIdx.z < nz1 serves the first range
    // Tile rasterisation: blockIdx.x is a linear tile id that walks strips of strip_w tile columns, row-major inside a
    // strip (strip_w = gx, the default: plain row-major).  CTAs are dispatched in id order, one at a time as SMs free
    // up, so the start-time skew between two tiles is about (chunk duration) x (id distance) / (CTAs in flight), and a
    // neighbour's halo only hits L2 while that skew is short.  Both neighbours matter equally: y-neighbours share 4 of
    // the 16 rows a tile reads, x-neighbours the two 128-byte lines at the ends of every 512-byte row (640 / 480).
    // Narrow strips were measured and lost (profiles/strip_r01.jsonl); the host bounds the chunk length instead.
    int gx, gy, strip_w;
    int kc;           // planes per CTA chunk
    int n_half;       // stage A: 1 or 2 half steps of B (stage B always applies 2)
    int j_quirk;      // Jx feeds all three components (FDTD_openmp semantics)
    // device-resident source evaluated in-kernel for stage B
    int src2;                  // 1: cells inside [s_lo, s_hi) take J = ((amp2*wx)*wy)*wz in stage B
    int s_lo[3], s_hi[3];      // global box
    const double* sw[3];       // device tables, indexed from s_lo
    double amp2;
    // TMA: 3-D tensor maps {Ni, Nj, nk + 2*GHOST_PLANES} of the six input arrays, box = {32*V cells, BY rows, 1 plane};
    // tiles whose footprint needs no periodic wrap in i / j fill their ring slots with six tensor copies per plane
    int use_tma;
    int st_cs;        // 1: output stores are streaming (evict-first in L2)
    // Store box in i, j (global coordinates, i bounds multiples of V): only cells inside it are written, CTAs whose
    // output tile misses it exit at once.  The whole grid for the periodic solver; the main box shrunk by the pass's
    // dependency reach (2 cells) for the PML solver, whose shell and rim are advanced by the sweep kernels.
    int sb_lo[2], sb_hi[2];
    alignas(64) CUtensorMap tmE[3];
    alignas(64) CUtensorMap tmB[3];
};

// ---- shared-memory plumbing ---------------------------------------------------------------------------------
// Every shared access of a thread is [sa + compile-time offset] (one address register): exchange arrays are
// (BY + 2) rows so that "one row up / down" is +-512 B with no clamping, the ring follows.
constexpr int T2_ROWB = FUSED_BX * 16;   // bytes per exchange / ring row (32 lanes x 16 B)
template <int BY> __host__ __device__ constexpr int t2_xq(int q) { return q * (BY + 2) * T2_ROWB; }          // exchange array q, own slot
template <int BY> __host__ __device__ constexpr int t2_ring0() { return 8 * (BY + 2) * T2_ROWB - T2_ROWB; }   // ring slot 0 comp 0, rel. to sa
constexpr int T2_D = 3;                  // ring depth = unroll factor of the k loop
template <int BY> __host__ __device__ constexpr int t2_mbar0() { return (8 * (BY + 2) + T2_D * 6 * BY) * T2_ROWB; }   // absolute
template <int BY>
constexpr size_t fused_t2_smem_bytes() {
    return (size_t)t2_mbar0<BY>() + 64;
}

// ---- mbarrier / TMA primitives --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "T2_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra T2_DONE;\n"
        "bra T2_WAIT;\n"
        "T2_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* tm, int x, int y, int z, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

template <typename T> struct SmemIO;
template <> struct SmemIO<double> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, double (&v)[2]) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v[0]), "=d"(v[1]) : "r"(a), "n"(OFF) : "memory");
    }
    template <int OFF> static __device__ __forceinline__ void st(unsigned a, const double (&v)[2]) {
        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(v[0]), "d"(v[1]) : "memory");
    }
};
template <> struct SmemIO<float> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, float (&v)[4]) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a), "n"(OFF) : "memory");
    }
    template <int OFF> static __device__ __forceinline__ void st(unsigned a, const float (&v)[4]) {
        asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};
// B += n_half * h(E)   (FDTD.cpp:121-126).  e = E(k), ek = E(k+1), (ezu, exu) = Ez, Ex one row up,
// (ez_nl, ey_nl) = Ez, Ey first element of the next lane.
template <typename T, int V>
__device__ __forceinline__ void t2_update_B(T (&b)[3][V], const T (&e)[3][V], const T (&ek)[3][V], const T (&ezu)[V],
                                            const T (&exu)[V], const T ez_nl, const T ey_nl, const double cBx,
                                            const double cBy, const double cBz, const bool two) {
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const double ex = (double)e[0][q], ey = (double)e[1][q], ez = (double)e[2][q];
        const double ezr = (double)((q == V - 1) ? ez_nl : e[2][(q + 1) % V]);
        const double eyr = (double)((q == V - 1) ? ey_nl : e[1][(q + 1) % V]);
        const double hx = dsub(dmul(cBz, dsub((double)ek[1][q], ey)), dmul(cBy, dsub((double)ezu[q], ez)));
        const double hy = dsub(dmul(cBx, dsub(ezr, ez)), dmul(cBz, dsub((double)ek[0][q], ex)));
        const double hz = dsub(dmul(cBy, dsub((double)exu[q], ex)), dmul(cBx, dsub(eyr, ey)));
        T nbx = (T)dadd((double)b[0][q], hx);
        T nby = (T)dadd((double)b[1][q], hy);
        T nbz = (T)dadd((double)b[2][q], hz);
        if (two) {
            nbx = (T)dadd((double)nbx, hx);
            nby = (T)dadd((double)nby, hy);
            nbz = (T)dadd((double)nbz, hz);
        }
        b[0][q] = nbx; b[1][q] = nby; b[2][q] = nbz;
    }
}

// E += g(B, J)   (FDTD.cpp:85-93 / kokkos_functors.h:81-89), in place.  b = B(k), (bxk, byk) = Bx, By at k-1,
// (bzd, bxd) = Bz, Bx one row down, (bz_pl, by_pl) = Bz, By last element of the previous lane.
template <typename T, int V>
__device__ __forceinline__ void t2_update_E(T (&e)[3][V], const T (&b)[3][V], const T (&bxk)[V], const T (&byk)[V],
                                            const T (&bzd)[V], const T (&bxd)[V], const T bz_pl, const T by_pl,
                                            const double cEx, const double cEy, const double cEz, const double cJ,
                                            const bool use_j, const T (&jv)[3][V]) {
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const double bx = (double)b[0][q], by = (double)b[1][q], bz = (double)b[2][q];
        const double bzl = (double)((q == 0) ? bz_pl : b[2][(q + V - 1) % V]);
        const double byl = (double)((q == 0) ? by_pl : b[1][(q + V - 1) % V]);
        double tx_ = dmul(cEy, dsub(bz, (double)bzd[q]));
        double ty_ = dmul(cEz, dsub(bx, (double)bxk[q]));
        double tz_ = dmul(cEx, dsub(by, byl));
        if (use_j) {
            tx_ = dadd(dmul(cJ, (double)jv[0][q]), tx_);
            ty_ = dadd(dmul(cJ, (double)jv[1][q]), ty_);
            tz_ = dadd(dmul(cJ, (double)jv[2][q]), tz_);
        }
        e[0][q] = (T)dadd((double)e[0][q], dsub(tx_, dmul(cEz, dsub(by, (double)byk[q]))));
        e[1][q] = (T)dadd((double)e[1][q], dsub(ty_, dmul(cEx, dsub(bz, bzl))));
        e[2][q] = (T)dadd((double)e[2][q], dsub(tz_, dmul(cEy, dsub(bx, (double)bxd[q]))));
    }
}

template <typename T>
__device__ __forceinline__ void t2_tile_of(const FusedT2Args<T>& a, int& bx, int& by) {
    const int id = (int)blockIdx.x;
    const int per = a.strip_w * a.gy;
    const int s = id / per, r = id - s * per;
    const int ws = min(a.strip_w, a.gx - s * a.strip_w);   // the last strip may be narrower
    by = r / ws;
    bx = s * a.strip_w + (r - by * ws);
}

// Per-thread constants of the pass (everything the plane iteration needs besides the register sets).
template <typename T>
struct T2Ctx {
    unsigned sa;            // shared address of the thread's slot in exchange array 0
    int tile_x, tile_y;     // first column / row of the tile's footprint (halo included; CTA-uniform)
    int i, jw;
    int kb, ke;
    bool needB1, needE1, needB2, needE2, ldE, out, j_ijA, j_ijB, producer;
    long long roff;         // element offset of (jw, iw) inside a plane
};

template <typename T>
__device__ __forceinline__ long long t2_plane_of(const FusedT2Args<T>& a, int k) {
    // plane (element offset) holding local plane k of the input generation: index wrap on a single GPU,
    // ghost planes -2, -1, nk, nk+1 on a slab rank
    if (a.g.wrap_k) {
        if (k < 0) k += a.g.nk;
        else if (k >= a.g.nk) k -= a.g.nk;
    }
    return (long long)k * a.g.plane;
}

template <int OFF>
__device__ __forceinline__ void cp_async16_at(unsigned a, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0+%1], [%2], 16;\n" ::"r"(a), "n"(OFF), "l"(gmem_src) : "memory");
}

// Asynchronous copies (LDGSTS, L1 bypass) of plane iteration kk into ring slot SLOT: old E(kk+1) and, for the rows
// that produce B1, B0(kk).  Every thread copies the six vectors it will read back itself, so the ring needs no
// barrier, only cp.async.wait_group.
template <typename T, int BY, int SLOT>
__device__ __forceinline__ void t2_issue_slot(const FusedT2Args<T>& a, const T2Ctx<T>& c, const int kk) {
    constexpr int COMPB = BY * T2_ROWB, SLOTB = 6 * COMPB;
    constexpr int RING = t2_ring0<BY>() + SLOT * SLOTB;
    if (c.ldE) {
        const long long pe = t2_plane_of(a, kk + 1) + c.roff;
        cp_async16_at<RING + 0 * COMPB>(c.sa, a.Ein[0] + pe);
        cp_async16_at<RING + 1 * COMPB>(c.sa, a.Ein[1] + pe);
        cp_async16_at<RING + 2 * COMPB>(c.sa, a.Ein[2] + pe);
        if (c.needB1) {
            const long long pb = t2_plane_of(a, kk) + c.roff;
            cp_async16_at<RING + 3 * COMPB>(c.sa, a.Bin[0] + pb);
            cp_async16_at<RING + 4 * COMPB>(c.sa, a.Bin[1] + pb);
            cp_async16_at<RING + 5 * COMPB>(c.sa, a.Bin[2] + pb);
        }
    }
}

// TMA flavour of the ring fill (tiles that need no periodic wrap in i / j): one thread issues six tensor copies, each
// a {32*V cells, BY rows, 1 plane} box that lands in the slot with exactly the ring's [row][lane] layout.
template <typename T, int BY, int SLOT>
__device__ __forceinline__ void t2_issue_slot_tma(const FusedT2Args<T>& a, const unsigned smem0, const int x, const int y, const int kk) {
    constexpr int COMPB = BY * T2_ROWB, SLOTB = 6 * COMPB;
    const unsigned bar = smem0 + (unsigned)(t2_mbar0<BY>() + 8 * SLOT);
    const unsigned dst = smem0 + (unsigned)(8 * (BY + 2) * T2_ROWB + SLOT * SLOTB);
    int ke = kk + 1, kb = kk;
    if (a.g.wrap_k) {
        if (ke < 0) ke += a.g.nk; else if (ke >= a.g.nk) ke -= a.g.nk;
        if (kb < 0) kb += a.g.nk; else if (kb >= a.g.nk) kb -= a.g.nk;
    }
    mbar_arrive_expect_tx(bar, (unsigned)(6 * COMPB));
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        tma_load_3d(dst + q * COMPB, &a.tmE[q], x, y, ke + GHOST_PLANES, bar);
        tma_load_3d(dst + (3 + q) * COMPB, &a.tmB[q], x, y, kb + GHOST_PLANES, bar);
    }
}

// One plane iteration.  Register sets by role on entry:
//   en : free -> old E(k+1)           e0 : old E(k) -> E1(k)          e1 : E1(k-1) -> E2(k-1) (stored)
//   b  : free -> B0(k) -> B1(k)        b1 : B1(k-1) -> B2(k-1) (stored)  b2 : B2(k-2) (x, y used)
template <typename T, int BY, bool TWO_A, bool HAS_J, bool TMA, int ABL, int SLOT>
__device__ __forceinline__ void t2_plane(const FusedT2Args<T>& a, const T2Ctx<T>& c, const int k, const unsigned parity,
                                         T (&en)[3][VecOf<T>::V], T (&e0)[3][VecOf<T>::V], T (&e1)[3][VecOf<T>::V],
                                         T (&b)[3][VecOf<T>::V], T (&b1)[3][VecOf<T>::V], T (&b2)[3][VecOf<T>::V]) {
    constexpr int V = VecOf<T>::V;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int UP = T2_ROWB, DN = -T2_ROWB;
    constexpr int XE0z = t2_xq<BY>(0), XE0x = t2_xq<BY>(1), XB1z = t2_xq<BY>(2), XB1x = t2_xq<BY>(3);
    constexpr int XE1z = t2_xq<BY>(4), XE1x = t2_xq<BY>(5), XB2z = t2_xq<BY>(6), XB2x = t2_xq<BY>(7);
    constexpr int COMPB = BY * T2_ROWB, SLOTB = 6 * COMPB;
    constexpr int RING = t2_ring0<BY>() + SLOT * SLOTB;
    constexpr int NEXT = (SLOT + T2_D - 1) % T2_D;     // slot freed by the previous plane
    using IO = SmemIO<T>;
    const unsigned sa = c.sa;
    const double cBx = a.c.cBx, cBy = a.c.cBy, cBz = a.c.cBz;
    const double cEx = a.c.cEx, cEy = a.c.cEy, cEz = a.c.cEz, cJ = a.c.cJ;

    // ================= phase X: B1(k) =============================================================================
    if (TMA) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
        if (c.producer && k + T2_D - 1 <= c.ke) t2_issue_slot_tma<T, BY, NEXT>(a, smem0, c.tile_x, c.tile_y, k + T2_D - 1);
        mbar_wait(smem0 + (unsigned)(t2_mbar0<BY>() + 8 * SLOT), parity);
    } else {
        if (k + T2_D - 1 <= c.ke && ABL != 3) t2_issue_slot<T, BY, NEXT>(a, c, k + T2_D - 1);
        cp_async_commit();
        cp_async_wait<T2_D - 1>();
    }
    IO::template ld<RING + 1 * COMPB>(sa, en[1]);
    IO::template ld<RING + 0 * COMPB>(sa, en[0]);
    IO::template ld<RING + 2 * COMPB>(sa, en[2]);
    if (c.needB1) {
        IO::template ld<RING + 3 * COMPB>(sa, b[0]);
        IO::template ld<RING + 4 * COMPB>(sa, b[1]);
        IO::template ld<RING + 5 * COMPB>(sa, b[2]);
        T ezu[V], exu[V];
        IO::template ld<XE0z + UP>(sa, ezu);
        IO::template ld<XE0x + UP>(sa, exu);
        const T ez_nl = __shfl_down_sync(FULL, e0[2][0], 1);
        const T ey_nl = __shfl_down_sync(FULL, e0[1][0], 1);
        if (ABL != 4) t2_update_B<T, V>(b, e0, en, ezu, exu, ez_nl, ey_nl, cBx, cBy, cBz, TWO_A);
        IO::template st<XB1z>(sa, b[2]);
        IO::template st<XB1x>(sa, b[0]);
    }
    if (ABL != 1) __syncthreads();   // barrier 1: B1(k) rows visible; everybody is done reading sE0 / sB2 of the previous plane

    // ================= phase Y: E1(k), then B2(k-1) ==================================================================
    IO::template st<XE0z>(sa, en[2]);   // old E(k+1) rows for the next plane's phase X
    IO::template st<XE0x>(sa, en[0]);
    if (c.needE1) {
        T bzd[V], bxd[V], jv[3][V];
        IO::template ld<XB1z + DN>(sa, bzd);
        IO::template ld<XB1x + DN>(sa, bxd);
        const T bz_pl = __shfl_up_sync(FULL, b[2][V - 1], 1);
        const T by_pl = __shfl_up_sync(FULL, b[1][V - 1], 1);
        bool use_j = false;
        if (HAS_J) {
            // J of step s applies to E1(k) on owned planes and on the halo planes that recompute a neighbour's
            // cells (same global coordinates, same J)
            int kgw = a.g.k0 + k;
            if (kgw < 0) kgw += a.g.Nk; else if (kgw >= a.g.Nk) kgw -= a.g.Nk;
            use_j = c.j_ijA && (kgw >= a.jbox.lo[2]) && (kgw < a.jbox.hi[2]);
            if (use_j) {
                const long long pj = t2_plane_of(a, k) + c.roff;
                ldg_vec<T, V>(a.J[0] + pj, jv[0]);
                ldg_vec<T, V>((a.j_quirk ? a.J[0] : a.J[1]) + pj, jv[1]);
                ldg_vec<T, V>((a.j_quirk ? a.J[0] : a.J[2]) + pj, jv[2]);
            }
        }
        // e0 (= old E(k)) becomes E1(k) in place
        if (ABL != 4) t2_update_E<T, V>(e0, b, b1[0], b1[1], bzd, bxd, bz_pl, by_pl, cEx, cEy, cEz, cJ, use_j, jv);
    }
    if (c.needB2) {
        T ezu[V], exu[V];
        IO::template ld<XE1z + UP>(sa, ezu);   // E1(k-1) one row up (written in phase Z of the previous plane)
        IO::template ld<XE1x + UP>(sa, exu);
        const T ez_nl = __shfl_down_sync(FULL, e1[2][0], 1);
        const T ey_nl = __shfl_down_sync(FULL, e1[1][0], 1);
        // b1 (= B1(k-1)) becomes B2(k-1) in place
        if (ABL != 4) t2_update_B<T, V>(b1, e1, e0, ezu, exu, ez_nl, ey_nl, cBx, cBy, cBz, true);
        IO::template st<XB2z>(sa, b1[2]);
        IO::template st<XB2x>(sa, b1[0]);
    }
    if (ABL != 1) __syncthreads();   // barrier 2: B2(k-1) rows and old E(k+1) rows visible; everybody is done reading sB1 / sE1

    // ================= phase Z: E2(k-1), stores ==========================================================================
    if (c.needE1) {
        IO::template st<XE1z>(sa, e0[2]);   // E1(k) rows for the next plane's phase Y
        IO::template st<XE1x>(sa, e0[0]);
    }
    if (c.needE2) {
        T bzd[V], bxd[V], jv[3][V];
        IO::template ld<XB2z + DN>(sa, bzd);
        IO::template ld<XB2x + DN>(sa, bxd);
        const T bz_pl = __shfl_up_sync(FULL, b1[2][V - 1], 1);
        const T by_pl = __shfl_up_sync(FULL, b1[1][V - 1], 1);
        const int kB = k - 1;                          // plane of stage B (an owned plane whenever it is stored)
        const bool stored = (kB >= c.kb);
        bool use_j = false;
        if (HAS_J) {
            const int kgB = a.g.k0 + kB;
            use_j = c.j_ijB && stored && (kgB >= a.jbox.lo[2]) && (kgB < a.jbox.hi[2]);
            if (use_j) {
                const long long pj = (long long)kB * a.g.plane + c.roff;
                ldg_vec<T, V>(a.J[0] + pj, jv[0]);
                ldg_vec<T, V>((a.j_quirk ? a.J[0] : a.J[1]) + pj, jv[1]);
                ldg_vec<T, V>((a.j_quirk ? a.J[0] : a.J[2]) + pj, jv[2]);
                if (a.src2 && kgB >= a.s_lo[2] && kgB < a.s_hi[2] && c.jw >= a.s_lo[1] && c.jw < a.s_hi[1]) {
                    const double wy = a.sw[1][c.jw - a.s_lo[1]], wz = a.sw[2][kgB - a.s_lo[2]];
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        const int ii = c.i + q;
                        if (ii >= a.s_lo[0] && ii < a.s_hi[0]) {
                            const T v = (T)dmul(dmul(dmul(a.amp2, a.sw[0][ii - a.s_lo[0]]), wy), wz);
                            jv[0][q] = v; jv[1][q] = v; jv[2][q] = v;
                        }
                    }
                }
            }
        }
        // e1 (= E1(k-1)) becomes E2(k-1) in place
        if (ABL != 4) t2_update_E<T, V>(e1, b1, b2[0], b2[1], bzd, bxd, bz_pl, by_pl, cEx, cEy, cEz, cJ, use_j, jv);
        if (c.out && stored && ABL != 2) {
            const long long o = (long long)kB * a.g.plane + c.roff;
            if (a.st_cs) {
                stg_vec_cs<T, V>(a.Eout[0] + o, e1[0]);
                stg_vec_cs<T, V>(a.Eout[1] + o, e1[1]);
                stg_vec_cs<T, V>(a.Eout[2] + o, e1[2]);
                stg_vec_cs<T, V>(a.Bout[0] + o, b1[0]);
                stg_vec_cs<T, V>(a.Bout[1] + o, b1[1]);
                stg_vec_cs<T, V>(a.Bout[2] + o, b1[2]);
            } else {
                stg_vec<T, V>(a.Eout[0] + o, e1[0]);
                stg_vec<T, V>(a.Eout[1] + o, e1[1]);
                stg_vec<T, V>(a.Eout[2] + o, e1[2]);
                stg_vec<T, V>(a.Bout[0] + o, b1[0]);
                stg_vec<T, V>(a.Bout[1] + o, b1[1]);
                stg_vec<T, V>(a.Bout[2] + o, b1[2]);
            }
        }
    }
}

template <typename T, int BY, bool TWO_A, bool HAS_J, bool TMA, int ABL>
__device__ __forceinline__ void fused_BE_T2_body(const FusedT2Args<T>& a) {
    constexpr int V = VecOf<T>::V;
    constexpr int TJU = BY - 4;               // output rows per CTA
    constexpr int TIU = FUSED_OUT_LANES * V;  // output cells per CTA row
    using IO = SmemIO<T>;
    static_assert(BY >= 5, "T2 pass needs at least one output row");

    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Ni = a.g.Ni, Nj = a.g.Nj;
    T2Ctx<T> c;
    c.sa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)((ty + 1) * T2_ROWB + tx * 16);   // the one shared-memory address register

    // ---- roles ---------------------------------------------------------------------------------------------
    {
        int bx, by;
        t2_tile_of(a, bx, by);
        c.tile_x = bx * TIU - V;
        c.tile_y = by * TJU - 2;
    }
    c.i = c.tile_x + tx * V;                                // first cell of this lane (may be -V or >= Ni)
    const bool lane_active = (c.i <= Ni);                   // i == Ni: right halo, wrapped to column 0
    const int iw = (c.i < 0) ? c.i + Ni : ((c.i >= Ni) ? c.i - Ni : c.i);
    const int j = c.tile_y + ty;
    const bool row_active = (j <= Nj + 1);
    c.jw = j % Nj;
    if (c.jw < 0) c.jw += Nj;
    // warp-uniform row roles: which stages this row has to produce for the CTA's output rows 2..BY-3
    c.needB1 = row_active && (ty <= BY - 2);
    c.needE1 = c.needB1 && (ty >= 1);
    c.needB2 = c.needE1 && (ty <= BY - 3);
    c.needE2 = c.needB2 && (ty >= 2) && (j < Nj);
    c.ldE = lane_active && row_active;
    c.producer = TMA && (ty == BY - 1) && (tx == 0);   // the top halo row's warp has no arithmetic of its own
    c.out = c.needE2 && lane_active && (tx >= 1) && (tx <= FUSED_OUT_LANES) && (c.i < Ni) &&
            (c.i >= a.sb_lo[0]) && (c.i + V <= a.sb_hi[0]) && (j >= a.sb_lo[1]) && (j < a.sb_hi[1]);
    c.roff = (long long)c.jw * a.g.pitch + (lane_active ? iw : 0);
    {
        const bool second = (int)blockIdx.z >= a.nz1;
        c.kb = (second ? a.k_lo2 : a.k_lo) + ((int)blockIdx.z - (second ? a.nz1 : 0)) * a.kc;
        c.ke = min(c.kb + a.kc, second ? a.k_hi2 : a.k_hi);
    }
    // J may be non-zero only inside jbox (global coordinates).  Stage A needs it on every cell whose E1 feeds an
    // output cell, halo lanes / rows included (their wrapped coordinates are tested); stage B only where it stores.
    c.j_ijA = HAS_J && lane_active && row_active && (iw < a.jbox.hi[0]) && (iw + V > a.jbox.lo[0]) &&
              (c.jw >= a.jbox.lo[1]) && (c.jw < a.jbox.hi[1]);
    c.j_ijB = c.j_ijA && c.out;

    // ---- register sets (rotating roles, see t2_plane) ----------------------------------------------------------
    T eA[3][V], eB[3][V], eC[3][V], bA[3][V], bB[3][V], bC[3][V];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) { eA[q][v] = eB[q][v] = eC[q][v] = (T)0; bA[q][v] = bB[q][v] = bC[q][v] = (T)0; }

    // ---- prologue: the first two ring slots, old E(kb-2) and its rows ------------------------------------------------
    const int k_first = c.kb - 2;
    if (TMA) {
        const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
        if (c.producer) {
#pragma unroll
            for (int d = 0; d < T2_D; ++d) mbar_init(smem0 + (unsigned)(t2_mbar0<BY>() + 8 * d), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t2_issue_slot_tma<T, BY, 0>(a, smem0, c.tile_x, c.tile_y, k_first);
            if (k_first + 1 <= c.ke) t2_issue_slot_tma<T, BY, 1>(a, smem0, c.tile_x, c.tile_y, k_first + 1);
        }
    } else {
        if (ABL != 3) t2_issue_slot<T, BY, 0>(a, c, k_first);
        cp_async_commit();
        if (k_first + 1 <= c.ke && ABL != 3) t2_issue_slot<T, BY, 1>(a, c, k_first + 1);
        cp_async_commit();
    }
    if (c.ldE) {
        const long long p0 = t2_plane_of(a, k_first) + c.roff;
        ldg_vec<T, V>(a.Ein[0] + p0, eB[0]);
        ldg_vec<T, V>(a.Ein[1] + p0, eB[1]);
        ldg_vec<T, V>(a.Ein[2] + p0, eB[2]);
    }
    IO::template st<t2_xq<BY>(0)>(c.sa, eB[2]);
    IO::template st<t2_xq<BY>(1)>(c.sa, eB[0]);
    __syncthreads();   // (also publishes the mbarrier initialisation to the waiting threads)

    unsigned parity = 0;
#pragma unroll 1
    for (int k = k_first; k <= c.ke; k += 3) {
        t2_plane<T, BY, TWO_A, HAS_J, TMA, ABL, 0>(a, c, k, parity, eA, eB, eC, bA, bB, bC);
        if (k + 1 > c.ke) break;
        t2_plane<T, BY, TWO_A, HAS_J, TMA, ABL, 1>(a, c, k + 1, parity, eC, eA, eB, bC, bA, bB);
        if (k + 2 > c.ke) break;
        t2_plane<T, BY, TWO_A, HAS_J, TMA, ABL, 2>(a, c, k + 2, parity, eB, eC, eA, bB, bC, bA);
        parity ^= 1u;
    }
    if (!TMA) cp_async_wait<0>();
}

// Does the footprint [lo, hi) of a tile (unwrapped coordinates, may stick out of [0, N) by the halo) meet the
// box [blo, bhi) or one of its periodic images?
__device__ __forceinline__ bool t2_meets(int lo, int hi, int blo, int bhi, int N) {
    return (lo < bhi && hi > blo) || (lo < bhi - N && hi > blo - N) || (lo < bhi + N && hi > blo + N);
}

template <typename T, int BY, int MINB, bool TWO_A, int ABL = 0>
__global__ void __launch_bounds__(FUSED_BX * BY, MINB) fused_BE_T2_kernel(const __grid_constant__ FusedT2Args<T> a) {
    constexpr int V = VecOf<T>::V;
    // CTA-uniform: only the few tiles whose footprint (halo included) meets the box where J may be non-zero run
    // the loop that knows about currents; tiles that need no periodic wrap in i / j load through TMA.
    int tbx, tby;
    t2_tile_of(a, tbx, tby);
    const int i0 = tbx * (FUSED_OUT_LANES * V) - V, j0 = tby * (BY - 4) - 2;
    // output tile = [i0 + V, i0 + V + 30 V) x [j0 + 2, j0 + BY - 2)
    if (i0 + V >= a.sb_hi[0] || i0 + V + FUSED_OUT_LANES * V <= a.sb_lo[0] || j0 + 2 >= a.sb_hi[1] || j0 + BY - 2 <= a.sb_lo[1]) return;
    const bool second = (int)blockIdx.z >= a.nz1;
    const int kb = (second ? a.k_lo2 : a.k_lo) + ((int)blockIdx.z - (second ? a.nz1 : 0)) * a.kc;
    const int k0 = a.g.k0 + kb - 2, k1 = a.g.k0 + min(kb + a.kc, second ? a.k_hi2 : a.k_hi) + 1;
    const bool has_j = !a.jbox.empty() && t2_meets(i0, i0 + FUSED_BX * V, a.jbox.lo[0], a.jbox.hi[0], a.g.Ni) &&
                       t2_meets(j0, j0 + BY, a.jbox.lo[1], a.jbox.hi[1], a.g.Nj) &&
                       t2_meets(k0, k1, a.jbox.lo[2], a.jbox.hi[2], a.g.Nk);
    const bool tma = a.use_tma && ABL == 0 && i0 >= 0 && i0 + FUSED_BX * V <= a.g.Ni && j0 >= 0 && j0 + BY <= a.g.Nj;
    if (has_j) {
        fused_BE_T2_body<T, BY, TWO_A, true, false, ABL>(a);   // (rare tiles: keep one flavour)
    } else if (tma) {
        fused_BE_T2_body<T, BY, TWO_A, false, true, ABL>(a);
    } else {
        fused_BE_T2_body<T, BY, TWO_A, false, false, ABL>(a);
    }
}

}  // namespace fdtd_b200
