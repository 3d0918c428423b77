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
// Storage vs arithmetic: V = 2 cells per lane for BOTH storage types.  Registers and the shared-memory row exchanges
// always hold doubles; only the ring (raw bytes copied from global memory) and the global loads / stores are in the
// storage type.  fp32 storage ("float storage, double arithmetic") therefore converts each input once
// when it leaves the ring and rounds each produced value once (double -> float -> double), instead of converting
// at every use -- F2F runs at 16 per clock per SM and was the bound of the first fp32 version of this pass.
//
// Tile roles (the dependency cone of one output cell reaches 2 cells down and 2 cells up in i, j and k):
//   * warp = 32 lanes x V cells of one row; lane 0 / lane 31 are the left / right halo lanes
//     (the 2 halo cells fit in one lane), lanes 1..30 store -> 30*V = 60 cells per row;
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

constexpr int T2_MAXCH = 72;   // plane chunks per launch (1024 planes / 16 + the two thin boundary chunks, with room)

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
    // Plane chunks: CTA (tile, z) produces local planes [chunk_lo[z], chunk_hi[z]).  The host builds the list (interior
    // chunks first; on a slab rank whose kernel waits for the halo itself, a thin top and a thin bottom chunk last).
    int nchunks;
    int chunk_lo[T2_MAXCH], chunk_hi[T2_MAXCH];
    // Tile rasterisation: blockIdx.x is a linear tile id that walks strips of strip_w tile columns, row-major inside a
    // strip (strip_w = gx, the default: plain row-major).  CTAs are dispatched in id order, one at a time as SMs free
    // up, so the start-time skew between two tiles is about (chunk duration) x (id distance) / (CTAs in flight), and a
    // neighbour's halo only hits L2 while that skew is short.  Both neighbours matter equally: y-neighbours share 4 of
    // the 16 rows a tile reads, x-neighbours the two 128-byte lines at the ends of every 512-byte row (640 / 480).
    // Narrow strips were measured and lost (profiles/strip_r01.jsonl); the host bounds the chunk length instead.
    int gx, gy, strip_w;
    int n_half;       // stage A: 1 or 2 half steps of B (stage B always applies 2)
    int j_quirk;      // Jx feeds all three components (FDTD_openmp semantics)
    // device-resident source evaluated in-kernel for stage B
    int src2;                  // 1: cells inside [s_lo, s_hi) take J = ((amp2*wx)*wy)*wz in stage B
                               // 2: cells inside [s_lo, s_hi) take J from the dense boxes jb2[] (host J writes issued between
                               //    the two update_fields() calls this launch pairs, csrc/fdtd_capi.cu "pending J writes")
    const T* jb2[3];           // mode 2: Jx, Jy, Jz on the box, [k][j][i] dense
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
    // Halo hand-off inside the kernel (z-slab rank on the peer transport, csrc/peer_ring.cu): the neighbours' copy engines
    // push their boundary planes into this rank's ghost planes while the pass runs and then set halo_flags[0] (lower
    // neighbour) / halo_flags[1] (upper neighbour) to the exchange number.  A CTA whose chunk reads ghost planes waits
    // until the flag has reached halo_seq before it starts; the host puts those (thin) chunks at the end of the list.
    const unsigned* halo_flags;   // nullptr: nothing to wait for
    unsigned* halo_err;           // set to 1 when a wait gives up (halo_timeout_ns): the host reports it at fdtd_sync
    unsigned long long halo_timeout_ns;   // generous (30 s by default): ranks may reach their first pass seconds apart
    unsigned halo_seq;
    alignas(64) CUtensorMap tmE[3];
    alignas(64) CUtensorMap tmB[3];
};

// ---- shared-memory plumbing ---------------------------------------------------------------------------------
// Every shared access of a thread is [sa + compile-time offset] (one address register): exchange arrays are
// (BY + 2) rows so that "one row up / down" is +-512 B with no clamping, the ring follows.
// A = arithmetic type: double for the two reference-faithful modes (fp64; fp32 storage with double arithmetic), float
// for the opt-in FDTD_FLAG_F32_ARITH mode.  A lane always owns 16 bytes of registers per component and plane:
template <typename A> __host__ __device__ constexpr int t2_v() { return 16 / (int)sizeof(A); }   // cells per lane: 2 / 4
constexpr int T2_ROWB = FUSED_BX * 16;   // bytes per exchange row (32 lanes x 16 bytes of A)
// Ring rows hold raw storage.  fp64: 64 cells = 512 B, the tile's footprint starts at byte 480*bx - 16 of a row.  fp32: the
// footprint starts at byte 240*bx - 8, but a TMA box must start on a 16-byte boundary in global memory, so the box
// starts 2 cells earlier and is 68 cells (272 B) wide; lane tx reads bytes [8 + 8 tx, 16 + 8 tx) of its ring row.
// (float storage with float arithmetic: 4 cells = 16 bytes per lane, the fp64 layout again -- 128-cell boxes, no skip)
template <typename T, typename A> __host__ __device__ constexpr int t2_rbox() { return sizeof(T) == sizeof(A) ? FUSED_BX * t2_v<A>() : FUSED_BX * t2_v<A>() + 4; }   // cells per ring row
template <typename T, typename A> __host__ __device__ constexpr int t2_rskip() { return sizeof(T) == sizeof(A) ? 0 : 2; }                                            // cells before lane 0
template <typename T, typename A> __host__ __device__ constexpr int t2_rrowb() { return t2_rbox<T, A>() * (int)sizeof(T); }                                          // bytes per ring row
template <int BY> __host__ __device__ constexpr int t2_xq(int q) { return q * (BY + 2) * T2_ROWB; }          // exchange array q, own slot
template <int BY> __host__ __device__ constexpr int t2_ringbase() { return 8 * (BY + 2) * T2_ROWB; }         // absolute offset of ring slot 0 comp 0
constexpr int T2_D = 3;                  // ring depth = unroll factor of the k loop
template <typename T, typename A, int BY> __host__ __device__ constexpr int t2_mbar0() { return t2_ringbase<BY>() + T2_D * 6 * BY * t2_rrowb<T, A>(); }   // absolute
template <typename T, typename A, int BY>
constexpr size_t fused_t2_smem_bytes() {
    return (size_t)t2_mbar0<T, A, BY>() + 64;
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

// Exchange arrays: 16 bytes of A per lane.
template <typename A> struct XIO;
template <> struct XIO<double> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, double (&v)[2]) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v[0]), "=d"(v[1]) : "r"(a), "n"(OFF) : "memory");
    }
    template <int OFF> static __device__ __forceinline__ void st(unsigned a, const double (&v)[2]) {
        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(v[0]), "d"(v[1]) : "memory");
    }
};
template <> struct XIO<float> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, float (&v)[4]) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a), "n"(OFF) : "memory");
    }
    template <int OFF> static __device__ __forceinline__ void st(unsigned a, const float (&v)[4]) {
        asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};
// Ring: raw storage, converted once on the way into registers.
template <typename T, typename A> struct RingIO;
template <> struct RingIO<double, double> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, double (&v)[2]) { XIO<double>::ld<OFF>(a, v); }
};
template <> struct RingIO<float, float> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, float (&v)[4]) { XIO<float>::ld<OFF>(a, v); }
};
template <> struct RingIO<float, double> {
    template <int OFF> static __device__ __forceinline__ void ld(unsigned a, double (&v)[2]) {
        float f0, f1;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(f0), "=f"(f1) : "r"(a), "n"(OFF) : "memory");
        v[0] = (double)f0; v[1] = (double)f1;
    }
};
// Global memory: one lane's cells of storage type <-> registers of arithmetic type.
__device__ __forceinline__ void t2_ldg(const double* p, double (&o)[2]) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    o[0] = v.x; o[1] = v.y;
}
__device__ __forceinline__ void t2_ldg(const float* p, double (&o)[2]) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = (double)v.x; o[1] = (double)v.y;
}
__device__ __forceinline__ void t2_ldg(const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <bool CS> __device__ __forceinline__ void t2_stg(double* p, const double (&o)[2]) {
    if (CS) __stcs(reinterpret_cast<double2*>(p), make_double2(o[0], o[1]));
    else *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
}
template <bool CS> __device__ __forceinline__ void t2_stg(float* p, const double (&o)[2]) {   // values are already float-representable
    if (CS) __stcs(reinterpret_cast<float2*>(p), make_float2((float)o[0], (float)o[1]));
    else *reinterpret_cast<float2*>(p) = make_float2((float)o[0], (float)o[1]);
}
template <bool CS> __device__ __forceinline__ void t2_stg(float* p, const float (&o)[4]) {
    if (CS) __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
    else *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}

// B += n_half * h(E)   (FDTD.cpp:121-126).  e = E(k), ek = E(k+1), (ezu, exu) = Ez, Ex one row up,
// (ez_nl, ey_nl) = Ez, Ey first element of the next lane.
// Round a double to the nearest float (ties to even), result kept as a double.  MAGIC: for x in the binade 2^E the
// constant M = 1.5 * 2^(E + 29) has ulp(M) = 2^(E - 23) = the float spacing there, so x + M is x rounded to that spacing
// by the adder (ties to even: M / ulp is even) and subtracting M is exact.  Float denormals: E clamped to -126.  Anything
// at or above 2^127 (overflow to inf, NaN) takes the conversion path.  Two DADD + a few integer ops on the high word
// instead of two F2F (which run at 16 per clock per SM).
// FDTD_T2_RINGUP (build-time, default on; only where a lane's ring entry is 16 bytes of the arithmetic type, i.e. fp64 and
// the fp32-arithmetic mode): stage A's "row above" values of old E(k) are read straight from the ring slot the TMA filled
// for the previous plane (the row above is the next 512-byte ring row) instead of travelling through the E0 exchange rows:
// two shared-memory stores fewer per thread and plane (of 22 accesses).  The refill of that slot moves behind barrier 1,
// after everybody's reads of it.  The first plane of a chunk still takes the exchange rows (its old E came by plain loads).
#ifndef FDTD_T2_RINGUP
#define FDTD_T2_RINGUP 1
#endif
#ifndef FDTD_T2_F32_MAGIC
#define FDTD_T2_F32_MAGIC 0   // build-time switch (fdtd_method_b200/build.py: FDTD_T2_F32_MAGIC=1 in the environment)
#endif
template <typename T> __device__ __forceinline__ float t2_round(float x) { return x; }   // float arithmetic: every operation already rounds to float
template <typename T>
__device__ __forceinline__ double t2_round(double x) {
    if (sizeof(T) == 8) return x;
    if (!FDTD_T2_F32_MAGIC) return (double)__double2float_rn(x);
    const int hi = __double2hiint(x);
    int e = (hi >> 20) & 0x7ff;
    if (e >= 1023 + 127) return (double)__double2float_rn(x);
    e = max(e, 1023 - 126);
    const double M = __hiloint2double(((e + 29) << 20) | 0x00080000, 0);
    return __dsub_rn(__dadd_rn(x, M), M);
}

template <typename T, typename A, int V>
__device__ __forceinline__ void t2_update_B(A (&b)[3][V], const A (&e)[3][V], const A (&ek)[3][V],
                                            const A (&ezu)[V], const A (&exu)[V], const A ez_nl,
                                            const A ey_nl, const A cBx, const A cBy, const A cBz,
                                            const bool two) {
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const A ex = e[0][q], ey = e[1][q], ez = e[2][q];
        const A ezr = (q == V - 1) ? ez_nl : e[2][(q + 1) % V];
        const A eyr = (q == V - 1) ? ey_nl : e[1][(q + 1) % V];
        const A hx = curl2(cBz, dsub(ek[1][q], ey), cBy, dsub(ezu[q], ez));
        const A hy = curl2(cBx, dsub(ezr, ez), cBz, dsub(ek[0][q], ex));
        const A hz = curl2(cBy, dsub(exu[q], ex), cBx, dsub(eyr, ey));
        A nbx = t2_round<T>(dadd(b[0][q], hx));
        A nby = t2_round<T>(dadd(b[1][q], hy));
        A nbz = t2_round<T>(dadd(b[2][q], hz));
        if (two) {
            nbx = t2_round<T>(dadd(nbx, hx));
            nby = t2_round<T>(dadd(nby, hy));
            nbz = t2_round<T>(dadd(nbz, hz));
        }
        b[0][q] = nbx; b[1][q] = nby; b[2][q] = nbz;
    }
}

// E += g(B, J)   (FDTD.cpp:85-93 / kokkos_functors.h:81-89), in place.  b = B(k), (bxk, byk) = Bx, By at k-1,
// (bzd, bxd) = Bz, Bx one row down, (bz_pl, by_pl) = Bz, By last element of the previous lane.
template <typename T, typename A, int V>
__device__ __forceinline__ void t2_update_E(A (&e)[3][V], const A (&b)[3][V], const A (&bxk)[V],
                                            const A (&byk)[V], const A (&bzd)[V], const A (&bxd)[V],
                                            const A bz_pl, const A by_pl, const A cEx, const A cEy,
                                            const A cEz, const A cJ, const bool use_j, const A (&jv)[3][V]) {
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const A bx = b[0][q], by = b[1][q], bz = b[2][q];
        const A bzl = (q == 0) ? bz_pl : b[2][(q + V - 1) % V];
        const A byl = (q == 0) ? by_pl : b[1][(q + V - 1) % V];
        const A jx = use_j ? jv[0][q] : (A)0, jy = use_j ? jv[1][q] : (A)0, jz = use_j ? jv[2][q] : (A)0;
        e[0][q] = t2_round<T>(dadd(e[0][q], curl2j(cEy, dsub(bz, bzd[q]), cEz, dsub(by, byk[q]), cJ, jx, use_j)));
        e[1][q] = t2_round<T>(dadd(e[1][q], curl2j(cEz, dsub(bx, bxk[q]), cEx, dsub(bz, bzl), cJ, jy, use_j)));
        e[2][q] = t2_round<T>(dadd(e[2][q], curl2j(cEx, dsub(by, byl), cEy, dsub(bx, bxd[q]), cJ, jz, use_j)));
    }
}

// Plane chunk [kb, ke) of this CTA.
template <typename T>
__device__ __forceinline__ void t2_chunk_of(const FusedT2Args<T>& a, int& kb, int& ke) {
    kb = a.chunk_lo[blockIdx.z];
    ke = a.chunk_hi[blockIdx.z];
}

// Wait (one thread, then the CTA) until the neighbour's planes of exchange `seq` have landed in this rank's ghost
// planes.  The flag is written by the neighbour's copy engine after its plane copies, on the same stream.
__device__ __forceinline__ void t2_halo_wait(const unsigned* flag, unsigned seq, unsigned* err, unsigned long long timeout_ns) {
    unsigned v;
    unsigned long long t0 = 0, now;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - seq) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) { *err = 1u; break; }   // a neighbour died: give up rather than hang the GPU for ever
        __nanosleep(200);
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
    unsigned sr;            // shared address of the thread's entry in ring slot 0, component 0 (= sa + constant for fp64)
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

// One lane's cells of one ring row: 16 bytes (L1 bypass) or 8 bytes (float storage, double arithmetic).
template <typename T, typename A, int OFF>
__device__ __forceinline__ void t2_cp_async_at(unsigned a, const T* gmem_src) {
    if (sizeof(T) == sizeof(A)) asm volatile("cp.async.cg.shared.global [%0+%1], [%2], 16;\n" ::"r"(a), "n"(OFF), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0+%1], [%2], 8;\n" ::"r"(a), "n"(OFF), "l"(gmem_src) : "memory");
}

// Asynchronous copies (LDGSTS) of plane iteration kk into ring slot SLOT: old E(kk+1) and, for the rows
// that produce B1, B0(kk).  Every thread copies the six vectors it will read back itself, so the ring needs no
// barrier, only cp.async.wait_group.
template <typename T, typename A, int BY, int SLOT>
__device__ __forceinline__ void t2_issue_slot(const FusedT2Args<T>& a, const T2Ctx<T>& c, const int kk) {
    constexpr int COMPB = BY * t2_rrowb<T, A>(), SLOTB = 6 * COMPB;
    constexpr int RING = SLOT * SLOTB;
    if (c.ldE) {
        const long long pe = t2_plane_of(a, kk + 1) + c.roff;
        t2_cp_async_at<T, A, RING + 0 * COMPB>(c.sr, a.Ein[0] + pe);
        t2_cp_async_at<T, A, RING + 1 * COMPB>(c.sr, a.Ein[1] + pe);
        t2_cp_async_at<T, A, RING + 2 * COMPB>(c.sr, a.Ein[2] + pe);
        if (c.needB1) {
            const long long pb = t2_plane_of(a, kk) + c.roff;
            t2_cp_async_at<T, A, RING + 3 * COMPB>(c.sr, a.Bin[0] + pb);
            t2_cp_async_at<T, A, RING + 4 * COMPB>(c.sr, a.Bin[1] + pb);
            t2_cp_async_at<T, A, RING + 5 * COMPB>(c.sr, a.Bin[2] + pb);
        }
    }
}

// TMA flavour of the ring fill (tiles that need no periodic wrap in i / j): one thread issues six tensor copies, each
// a {t2_rbox cells, BY rows, 1 plane} box that lands in the slot with exactly the ring's [row][lane] layout.
template <typename T, typename A, int BY, int SLOT>
__device__ __forceinline__ void t2_issue_slot_tma(const FusedT2Args<T>& a, const unsigned smem0, const int x, const int y, const int kk) {
    constexpr int COMPB = BY * t2_rrowb<T, A>(), SLOTB = 6 * COMPB;
    const unsigned bar = smem0 + (unsigned)(t2_mbar0<T, A, BY>() + 8 * SLOT);
    const unsigned dst = smem0 + (unsigned)(t2_ringbase<BY>() + SLOT * SLOTB);
    int ke = kk + 1, kb = kk;
    if (a.g.wrap_k) {
        if (ke < 0) ke += a.g.nk; else if (ke >= a.g.nk) ke -= a.g.nk;
        if (kb < 0) kb += a.g.nk; else if (kb >= a.g.nk) kb -= a.g.nk;
    }
    mbar_arrive_expect_tx(bar, (unsigned)(6 * COMPB));
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        tma_load_3d(dst + q * COMPB, &a.tmE[q], x - t2_rskip<T, A>(), y, ke + GHOST_PLANES, bar);
        tma_load_3d(dst + (3 + q) * COMPB, &a.tmB[q], x - t2_rskip<T, A>(), y, kb + GHOST_PLANES, bar);
    }
}

// One plane iteration.  Register sets by role on entry:
//   en : free -> old E(k+1)           e0 : old E(k) -> E1(k)          e1 : E1(k-1) -> E2(k-1) (stored)
//   b  : free -> B0(k) -> B1(k)        b1 : B1(k-1) -> B2(k-1) (stored)  b2 : B2(k-2) (x, y used)
template <typename T, typename A, int BY, bool TWO_A, bool HAS_J, bool TMA, int ABL, int SLOT>
__device__ __forceinline__ void t2_plane(const FusedT2Args<T>& a, const T2Ctx<T>& c, const int k, const unsigned parity,
                                         A (&en)[3][t2_v<A>()], A (&e0)[3][t2_v<A>()], A (&e1)[3][t2_v<A>()],
                                         A (&b)[3][t2_v<A>()], A (&b1)[3][t2_v<A>()], A (&b2)[3][t2_v<A>()]) {
    constexpr int V = t2_v<A>();
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int UP = T2_ROWB, DN = -T2_ROWB;
    constexpr int XE0z = t2_xq<BY>(0), XE0x = t2_xq<BY>(1), XB1z = t2_xq<BY>(2), XB1x = t2_xq<BY>(3);
    constexpr int XE1z = t2_xq<BY>(4), XE1x = t2_xq<BY>(5), XB2z = t2_xq<BY>(6), XB2x = t2_xq<BY>(7);
    constexpr int COMPB = BY * t2_rrowb<T, A>(), SLOTB = 6 * COMPB;
    constexpr int RING = SLOT * SLOTB;                 // relative to c.sr
    constexpr int NEXT = (SLOT + T2_D - 1) % T2_D;     // slot freed by the previous plane
    using IO = XIO<A>;
    using RIO = RingIO<T, A>;
    constexpr bool RINGUP = FDTD_T2_RINGUP && sizeof(T) == sizeof(A) && ABL == 0;
    constexpr int PREV = NEXT * SLOTB + t2_rrowb<T, A>();   // ring slot of the previous plane (old E(k), B0(k-1)), one row up
    const unsigned sa = c.sa, sr = c.sr;
    const A cBx = (A)a.c.cBx, cBy = (A)a.c.cBy, cBz = (A)a.c.cBz;
    const A cEx = (A)a.c.cEx, cEy = (A)a.c.cEy, cEz = (A)a.c.cEz, cJ = (A)a.c.cJ;

    // ================= phase X: B1(k) =============================================================================
    if (TMA) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
        if (!RINGUP && c.producer && k + T2_D - 1 <= c.ke) t2_issue_slot_tma<T, A, BY, NEXT>(a, smem0, c.tile_x, c.tile_y, k + T2_D - 1);
        mbar_wait(smem0 + (unsigned)(t2_mbar0<T, A, BY>() + 8 * SLOT), parity);
    } else if (!RINGUP) {
        if (k + T2_D - 1 <= c.ke && ABL != 3) t2_issue_slot<T, A, BY, NEXT>(a, c, k + T2_D - 1);
        cp_async_commit();
        cp_async_wait<T2_D - 1>();
    } else {
        cp_async_wait<T2_D - 2>();   // (the refill of slot NEXT is issued behind barrier 1: one group fewer in flight here)
    }
    RIO::template ld<RING + 1 * COMPB>(sr, en[1]);
    RIO::template ld<RING + 0 * COMPB>(sr, en[0]);
    RIO::template ld<RING + 2 * COMPB>(sr, en[2]);
    if (c.needB1) {
        RIO::template ld<RING + 3 * COMPB>(sr, b[0]);
        RIO::template ld<RING + 4 * COMPB>(sr, b[1]);
        RIO::template ld<RING + 5 * COMPB>(sr, b[2]);
        A ezu[V], exu[V];
        if (RINGUP && k != c.kb - 2) {
            RIO::template ld<PREV + 2 * COMPB>(sr, ezu);
            RIO::template ld<PREV + 0 * COMPB>(sr, exu);
        } else {
            IO::template ld<XE0z + UP>(sa, ezu);
            IO::template ld<XE0x + UP>(sa, exu);
        }
        const A ez_nl = __shfl_down_sync(FULL, e0[2][0], 1);
        const A ey_nl = __shfl_down_sync(FULL, e0[1][0], 1);
        if (ABL != 4) t2_update_B<T, A, V>(b, e0, en, ezu, exu, ez_nl, ey_nl, cBx, cBy, cBz, TWO_A);
        IO::template st<XB1z>(sa, b[2]);
        IO::template st<XB1x>(sa, b[0]);
    }
    if (ABL != 1) __syncthreads();   // barrier 1: B1(k) rows visible; everybody is done reading sE0 / sB2 of the previous plane

    // ================= phase Y: E1(k), then B2(k-1) ==================================================================
    if (RINGUP) {
        // everybody has read slot NEXT (own row in the previous plane, the row above in this one): refill it
        if (TMA) {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
            if (c.producer && k + T2_D - 1 <= c.ke) t2_issue_slot_tma<T, A, BY, NEXT>(a, smem0, c.tile_x, c.tile_y, k + T2_D - 1);
        } else {
            if (k + T2_D - 1 <= c.ke) t2_issue_slot<T, A, BY, NEXT>(a, c, k + T2_D - 1);
            cp_async_commit();
        }
    } else {
        IO::template st<XE0z>(sa, en[2]);   // old E(k+1) rows for the next plane's phase X
        IO::template st<XE0x>(sa, en[0]);
    }
    if (c.needE1) {
        A bzd[V], bxd[V], jv[3][V];
        IO::template ld<XB1z + DN>(sa, bzd);
        IO::template ld<XB1x + DN>(sa, bxd);
        const A bz_pl = __shfl_up_sync(FULL, b[2][V - 1], 1);
        const A by_pl = __shfl_up_sync(FULL, b[1][V - 1], 1);
        bool use_j = false;
        if (HAS_J) {
            // J of step s applies to E1(k) on owned planes and on the halo planes that recompute a neighbour's
            // cells (same global coordinates, same J)
            int kgw = a.g.k0 + k;
            if (kgw < 0) kgw += a.g.Nk; else if (kgw >= a.g.Nk) kgw -= a.g.Nk;
            use_j = c.j_ijA && (kgw >= a.jbox.lo[2]) && (kgw < a.jbox.hi[2]);
            if (use_j) {
                const long long pj = t2_plane_of(a, k) + c.roff;
                t2_ldg(a.J[0] + pj, jv[0]);
                t2_ldg((a.j_quirk ? a.J[0] : a.J[1]) + pj, jv[1]);
                t2_ldg((a.j_quirk ? a.J[0] : a.J[2]) + pj, jv[2]);
            }
        }
        // e0 (= old E(k)) becomes E1(k) in place
        if (ABL != 4) t2_update_E<T, A, V>(e0, b, b1[0], b1[1], bzd, bxd, bz_pl, by_pl, cEx, cEy, cEz, cJ, use_j, jv);
    }
    if (c.needB2) {
        A ezu[V], exu[V];
        IO::template ld<XE1z + UP>(sa, ezu);   // E1(k-1) one row up (written in phase Z of the previous plane)
        IO::template ld<XE1x + UP>(sa, exu);
        const A ez_nl = __shfl_down_sync(FULL, e1[2][0], 1);
        const A ey_nl = __shfl_down_sync(FULL, e1[1][0], 1);
        // b1 (= B1(k-1)) becomes B2(k-1) in place
        if (ABL != 4) t2_update_B<T, A, V>(b1, e1, e0, ezu, exu, ez_nl, ey_nl, cBx, cBy, cBz, true);
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
        A bzd[V], bxd[V], jv[3][V];
        IO::template ld<XB2z + DN>(sa, bzd);
        IO::template ld<XB2x + DN>(sa, bxd);
        const A bz_pl = __shfl_up_sync(FULL, b1[2][V - 1], 1);
        const A by_pl = __shfl_up_sync(FULL, b1[1][V - 1], 1);
        const int kB = k - 1;                          // plane of stage B (an owned plane whenever it is stored)
        const bool stored = (kB >= c.kb);
        bool use_j = false;
        if (HAS_J) {
            const int kgB = a.g.k0 + kB;
            use_j = c.j_ijB && stored && (kgB >= a.jbox.lo[2]) && (kgB < a.jbox.hi[2]);
            if (use_j) {
                const long long pj = (long long)kB * a.g.plane + c.roff;
                t2_ldg(a.J[0] + pj, jv[0]);
                t2_ldg((a.j_quirk ? a.J[0] : a.J[1]) + pj, jv[1]);
                t2_ldg((a.j_quirk ? a.J[0] : a.J[2]) + pj, jv[2]);
                if (a.src2 == 2 && kgB >= a.s_lo[2] && kgB < a.s_hi[2] && c.jw >= a.s_lo[1] && c.jw < a.s_hi[1]) {
                    const int bni = a.s_hi[0] - a.s_lo[0];
                    const long long row = ((long long)(kgB - a.s_lo[2]) * (a.s_hi[1] - a.s_lo[1]) + (c.jw - a.s_lo[1])) * bni - a.s_lo[0];
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        const int ii = c.i + q;
                        if (ii >= a.s_lo[0] && ii < a.s_hi[0]) {
                            jv[0][q] = (A)a.jb2[0][row + ii];
                            jv[1][q] = (A)a.jb2[a.j_quirk ? 0 : 1][row + ii];
                            jv[2][q] = (A)a.jb2[a.j_quirk ? 0 : 2][row + ii];
                        }
                    }
                } else if (a.src2 == 1 && kgB >= a.s_lo[2] && kgB < a.s_hi[2] && c.jw >= a.s_lo[1] && c.jw < a.s_hi[1]) {
                    const double wy = a.sw[1][c.jw - a.s_lo[1]], wz = a.sw[2][kgB - a.s_lo[2]];
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        const int ii = c.i + q;
                        if (ii >= a.s_lo[0] && ii < a.s_hi[0]) {
                            const A v = (A)t2_round<T>(dmul(dmul(dmul(a.amp2, a.sw[0][ii - a.s_lo[0]]), wy), wz));   // (double product, one rounding: source_kernel)
                            jv[0][q] = v; jv[1][q] = v; jv[2][q] = v;
                        }
                    }
                }
            }
        }
        // e1 (= E1(k-1)) becomes E2(k-1) in place
        if (ABL != 4) t2_update_E<T, A, V>(e1, b1, b2[0], b2[1], bzd, bxd, bz_pl, by_pl, cEx, cEy, cEz, cJ, use_j, jv);
        if (c.out && stored && ABL != 2) {
            const long long o = (long long)kB * a.g.plane + c.roff;
            if (a.st_cs) {
                t2_stg<true>(a.Eout[0] + o, e1[0]);
                t2_stg<true>(a.Eout[1] + o, e1[1]);
                t2_stg<true>(a.Eout[2] + o, e1[2]);
                t2_stg<true>(a.Bout[0] + o, b1[0]);
                t2_stg<true>(a.Bout[1] + o, b1[1]);
                t2_stg<true>(a.Bout[2] + o, b1[2]);
            } else {
                t2_stg<false>(a.Eout[0] + o, e1[0]);
                t2_stg<false>(a.Eout[1] + o, e1[1]);
                t2_stg<false>(a.Eout[2] + o, e1[2]);
                t2_stg<false>(a.Bout[0] + o, b1[0]);
                t2_stg<false>(a.Bout[1] + o, b1[1]);
                t2_stg<false>(a.Bout[2] + o, b1[2]);
            }
        }
    }
}

template <typename T, typename A, int BY, bool TWO_A, bool HAS_J, bool TMA, int ABL>
__device__ __forceinline__ void fused_BE_T2_body(const FusedT2Args<T>& a) {
    constexpr int V = t2_v<A>();
    constexpr int TJU = BY - 4;               // output rows per CTA
    constexpr int TIU = FUSED_OUT_LANES * V;  // output cells per CTA row
    using IO = XIO<A>;
    static_assert(BY >= 5, "T2 pass needs at least one output row");

    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Ni = a.g.Ni, Nj = a.g.Nj;
    T2Ctx<T> c;
    c.sa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)((ty + 1) * T2_ROWB + tx * 16);   // the one shared-memory address register ...
    if (sizeof(T) == sizeof(A)) c.sr = c.sa + (unsigned)(t2_ringbase<BY>() - T2_ROWB);                  // ... (16 bytes of storage per lane: the ring is sa + constant)
    else c.sr = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)(t2_ringbase<BY>() + ty * t2_rrowb<T, A>() + (t2_rskip<T, A>() + tx * V) * (int)sizeof(T));

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
    t2_chunk_of(a, c.kb, c.ke);
    // J may be non-zero only inside jbox (global coordinates).  Stage A needs it on every cell whose E1 feeds an
    // output cell, halo lanes / rows included (their wrapped coordinates are tested); stage B only where it stores.
    c.j_ijA = HAS_J && lane_active && row_active && (iw < a.jbox.hi[0]) && (iw + V > a.jbox.lo[0]) &&
              (c.jw >= a.jbox.lo[1]) && (c.jw < a.jbox.hi[1]);
    c.j_ijB = c.j_ijA && c.out;

    // ---- register sets (rotating roles, see t2_plane) ----------------------------------------------------------
    A eA[3][V], eB[3][V], eC[3][V], bA[3][V], bB[3][V], bC[3][V];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) { eA[q][v] = eB[q][v] = eC[q][v] = (A)0; bA[q][v] = bB[q][v] = bC[q][v] = (A)0; }

    // ---- prologue: the first two ring slots, old E(kb-2) and its rows ------------------------------------------------
    const int k_first = c.kb - 2;
    if (TMA) {
        const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem_raw);
        if (c.producer) {
#pragma unroll
            for (int d = 0; d < T2_D; ++d) mbar_init(smem0 + (unsigned)(t2_mbar0<T, A, BY>() + 8 * d), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t2_issue_slot_tma<T, A, BY, 0>(a, smem0, c.tile_x, c.tile_y, k_first);
            if (k_first + 1 <= c.ke) t2_issue_slot_tma<T, A, BY, 1>(a, smem0, c.tile_x, c.tile_y, k_first + 1);
        }
    } else {
        if (ABL != 3) t2_issue_slot<T, A, BY, 0>(a, c, k_first);
        cp_async_commit();
        if (k_first + 1 <= c.ke && ABL != 3) t2_issue_slot<T, A, BY, 1>(a, c, k_first + 1);
        cp_async_commit();
    }
    if (c.ldE) {
        const long long p0 = t2_plane_of(a, k_first) + c.roff;
        t2_ldg(a.Ein[0] + p0, eB[0]);
        t2_ldg(a.Ein[1] + p0, eB[1]);
        t2_ldg(a.Ein[2] + p0, eB[2]);
    }
    IO::template st<t2_xq<BY>(0)>(c.sa, eB[2]);
    IO::template st<t2_xq<BY>(1)>(c.sa, eB[0]);
    __syncthreads();   // (also publishes the mbarrier initialisation to the waiting threads)

    unsigned parity = 0;
#pragma unroll 1
    for (int k = k_first; k <= c.ke; k += 3) {
        t2_plane<T, A, BY, TWO_A, HAS_J, TMA, ABL, 0>(a, c, k, parity, eA, eB, eC, bA, bB, bC);
        if (k + 1 > c.ke) break;
        t2_plane<T, A, BY, TWO_A, HAS_J, TMA, ABL, 1>(a, c, k + 1, parity, eC, eA, eB, bC, bA, bB);
        if (k + 2 > c.ke) break;
        t2_plane<T, A, BY, TWO_A, HAS_J, TMA, ABL, 2>(a, c, k + 2, parity, eB, eC, eA, bB, bC, bA);
        parity ^= 1u;
    }
    if (!TMA) cp_async_wait<0>();
}

// Does the footprint [lo, hi) of a tile (unwrapped coordinates, may stick out of [0, N) by the halo) meet the
// box [blo, bhi) or one of its periodic images?
__device__ __forceinline__ bool t2_meets(int lo, int hi, int blo, int bhi, int N) {
    return (lo < bhi && hi > blo) || (lo < bhi - N && hi > blo - N) || (lo < bhi + N && hi > blo + N);
}

template <typename T, typename A, int BY, int MINB, bool TWO_A, int ABL = 0>
__global__ void __launch_bounds__(FUSED_BX * BY, MINB) fused_BE_T2_kernel(const __grid_constant__ FusedT2Args<T> a) {
    constexpr int V = t2_v<A>();
    // CTA-uniform: only the few tiles whose footprint (halo included) meets the box where J may be non-zero run
    // the loop that knows about currents; tiles that need no periodic wrap in i / j load through TMA.
    int tbx, tby;
    t2_tile_of(a, tbx, tby);
    const int i0 = tbx * (FUSED_OUT_LANES * V) - V, j0 = tby * (BY - 4) - 2;
    // output tile = [i0 + V, i0 + V + 30 V) x [j0 + 2, j0 + BY - 2)
    if (i0 + V >= a.sb_hi[0] || i0 + V + FUSED_OUT_LANES * V <= a.sb_lo[0] || j0 + 2 >= a.sb_hi[1] || j0 + BY - 2 <= a.sb_lo[1]) return;
    int kb, ke;
    t2_chunk_of(a, kb, ke);
    const int k0 = a.g.k0 + kb - 2, k1 = a.g.k0 + ke + 1;
    if (a.halo_flags) {
        // the chunk reads planes kb-2 .. ke+1: ghost planes below 0 come from the lower neighbour, at or above nk from the upper
        const bool need_dn = kb - 2 < 0, need_up = ke + 1 >= a.g.nk;
        if (need_dn || need_up) {
            if (threadIdx.x == 0 && threadIdx.y == 0) {
                if (need_dn) t2_halo_wait(a.halo_flags + 0, a.halo_seq, a.halo_err, a.halo_timeout_ns);
                if (need_up) t2_halo_wait(a.halo_flags + 1, a.halo_seq, a.halo_err, a.halo_timeout_ns);
            }
            __syncthreads();
            asm volatile("fence.proxy.async;" ::: "memory");   // the TMA loads of the ghost planes are ordered after the acquire
        }
    }
    const bool has_j = !a.jbox.empty() && t2_meets(i0, i0 + FUSED_BX * V, a.jbox.lo[0], a.jbox.hi[0], a.g.Ni) &&
                       t2_meets(j0, j0 + BY, a.jbox.lo[1], a.jbox.hi[1], a.g.Nj) &&
                       t2_meets(k0, k1, a.jbox.lo[2], a.jbox.hi[2], a.g.Nk);
    const bool tma = a.use_tma && ABL == 0 && i0 >= 0 && i0 + FUSED_BX * V <= a.g.Ni && j0 >= 0 && j0 + BY <= a.g.Nj;
    if (has_j) {
        fused_BE_T2_body<T, A, BY, TWO_A, true, false, ABL>(a);   // (rare tiles: keep one flavour)
    } else if (tma) {
        fused_BE_T2_body<T, A, BY, TWO_A, false, true, ABL>(a);
    } else {
        fused_BE_T2_body<T, A, BY, TWO_A, false, false, ABL>(a);
    }
}

}  // namespace fdtd_b200
