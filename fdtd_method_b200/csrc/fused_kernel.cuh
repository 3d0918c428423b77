// fused_kernel.cuh -- one-pass fused B+E update ("BE pass"), the fast path of fdtd_step().
//
// Replaces, for the periodic solver, the whole body of FDTD::update_fields()
// (reference src/FDTD/FDTD.cpp:153-157 = update_B, update_E, update_B; Kokkos twin
// src/FDTD_kokkos/FDTD_kokkos.cpp:91-104) with ONE kernel per time step:
//
//     B' = round(round(B + h(E)) + h(E))      (n_half = 2: trailing half step of the previous call merged with
//                                              the leading half step of this one; n_half = 1 on the first call)
//     E' = E + g(B', J)
//
// followed (lazily, only when somebody reads the fields) by the trailing half step B'' = B' + h(E') of
// sweep_B_kernel.  Same operations in the same order per cell as the reference, hence bit-identical.
//
// HBM traffic per cell-step: read E(3) + B(3), write E'(3) + B'(3) = 12 words (96 B fp64), against
// 30 words for the reference's three sweeps and 18-21 words for two unfused sweeps.
//
// E'(i,j,k) needs B' at (i,j,k), (i-1,j,k), (i,j-1,k), (i,j,k-1), and B' needs the OLD E at +1 in i, j, k,
// so a CTA that produces a tile of E' must also compute B' one cell to the left / below and see the old E one
// cell to the right / above.  In-place update would race between CTAs, so E and B are double-buffered
// (generation `cur` is read, `cur^1` written) and every tile is independent:
//
//   * warp = 32 lanes x V cells of one row; lane 0 is the left-halo lane (computes B' only), lane 31 the
//     right-halo lane (loads old E only), lanes 1..30 produce output -> 30*V*sizeof(T) = 480 B per row,
//     a whole number of 32 B sectors;  i-neighbours travel by __shfl_up/down (north_star (b));
//   * CTA = BY warps, each owning RJ consecutive rows; the first row slot is the bottom halo (B' only), the
//     last the top halo (old E only); j-neighbours between warps travel through a double-buffered
//     shared-memory row exchange, one __syncthreads per plane;
//   * each thread streams a chunk of k planes upward keeping E(k) and B'(k-1) in registers; the chunk starts
//     one plane early to rebuild B'(k_begin - 1) (redundant work 1/kc).
//   * periodic wrap: halo lanes / rows / planes simply read the wrapped address; on a z-slab rank the k
//     halos are the ghost planes filled by the NCCL ring exchange.
//
// Requires Ni % V == 0 (the host falls back to the two-sweep kernels otherwise).
#pragma once

#include "fdtd_common.cuh"

namespace fdtd_b200 {

template <typename T>
struct FusedArgs {
    Geom g;
    Coefs c;
    const T* Ein[3];
    const T* Bin[3];
    const T* J[3];
    T* Eout[3];
    T* Bout[3];
    JBox jbox;
    int k_lo, k_hi;   // local planes [k_lo, k_hi) produced by this launch
    int kc;           // planes per CTA chunk
    int n_half;       // 1 or 2 half steps of B
    int j_quirk;      // Jx feeds all three components (FDTD_openmp semantics)
};

constexpr int FUSED_BX = 32;
constexpr int FUSED_OUT_LANES = 30;

template <typename T> struct FusedVec;
template <> struct FusedVec<double> {
    using type = double2;
    static __device__ __forceinline__ double2 pack(const double (&v)[2]) { return make_double2(v[0], v[1]); }
    static __device__ __forceinline__ void unpack(const double2& p, double (&v)[2]) { v[0] = p.x; v[1] = p.y; }
};
template <> struct FusedVec<float> {
    using type = float4;
    static __device__ __forceinline__ float4 pack(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void unpack(const float4& p, float (&v)[4]) { v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; }
};

template <typename T, int V>
__device__ __forceinline__ void ldg_vec(const T* p, T (&o)[V]) {
    typename FusedVec<T>::type v = *reinterpret_cast<const typename FusedVec<T>::type*>(p);
    FusedVec<T>::unpack(v, o);
}
template <typename T, int V>
__device__ __forceinline__ void stg_vec(T* p, const T (&o)[V]) {
    *reinterpret_cast<typename FusedVec<T>::type*>(p) = FusedVec<T>::pack(o);
}

// Streaming store (st.global.cs): the line is written once and not read again by this pass, so it should be the
// first to leave L2 -- keeps the input lines that neighbouring tiles re-read (tile halos) resident longer.
template <typename T, int V>
__device__ __forceinline__ void stg_vec_cs(T* p, const T (&o)[V]) {
    __stcs(reinterpret_cast<typename FusedVec<T>::type*>(p), FusedVec<T>::pack(o));
}

// Shared-memory row exchange: [buffer][component][warp row][lane] of 16-byte vectors.
template <typename T, int BY>
struct FusedSmem {
    typename FusedVec<T>::type e[2][2][BY][FUSED_BX];   // old E(k) of each warp's first row: comp 0 = Ez, 1 = Ex
    typename FusedVec<T>::type b[2][2][BY][FUSED_BX];   // new B'(k) of each warp's last row: comp 0 = Bz, 1 = Bx
};

template <typename T, int BY, int RJ, int MINB>
__global__ void __launch_bounds__(FUSED_BX * BY, MINB) fused_BE_kernel(const FusedArgs<T> a) {
    constexpr int V = VecOf<T>::V;
    constexpr int SLOTS = BY * RJ;            // row slots per CTA
    constexpr int TJU = SLOTS - 2;            // output rows per CTA
    constexpr int TIU = FUSED_OUT_LANES * V;  // output cells per CTA row
    constexpr unsigned FULL = 0xffffffffu;
    using VT = typename FusedVec<T>::type;

    __shared__ FusedSmem<T, BY> sm;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Ni = a.g.Ni, Nj = a.g.Nj;

    // ---- lane role ---------------------------------------------------------------------------
    const int i = blockIdx.x * TIU - V + tx * V;                 // first cell of this lane (may be -V or >= Ni)
    const bool lane_active = (i <= Ni);                          // i == Ni: right halo, wrapped to column 0
    const int iw = (i < 0) ? i + Ni : ((i == Ni) ? 0 : i);
    const bool lane_B = lane_active && (tx <= FUSED_OUT_LANES) && (i < Ni);   // computes B'
    const bool lane_out = lane_B && (tx >= 1);                                // stores E', B'

    // ---- row roles ---------------------------------------------------------------------------
    bool row_active[RJ], row_B[RJ], row_out[RJ];
    long long roff[RJ];
    int jrow[RJ];
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        const int slot = ty * RJ + r;
        const int j = blockIdx.y * TJU - 1 + slot;
        row_active[r] = (j <= Nj);
        const int jw = (j < 0) ? Nj - 1 : ((j == Nj) ? 0 : j);
        row_B[r] = row_active[r] && (slot <= SLOTS - 2) && (j < Nj);
        row_out[r] = row_B[r] && (slot >= 1);
        jrow[r] = jw;
        roff[r] = (long long)jw * a.g.pitch + iw;
    }

    const int kb = a.k_lo + blockIdx.z * a.kc;
    const int ke = min(kb + a.kc, a.k_hi);

    const double cBx = a.c.cBx, cBy = a.c.cBy, cBz = a.c.cBz;
    const double cEx = a.c.cEx, cEy = a.c.cEy, cEz = a.c.cEz, cJ = a.c.cJ;

    const T* __restrict__ Ex = a.Ein[0];
    const T* __restrict__ Ey = a.Ein[1];
    const T* __restrict__ Ez = a.Ein[2];
    const T* __restrict__ Bx = a.Bin[0];
    const T* __restrict__ By = a.Bin[1];
    const T* __restrict__ Bz = a.Bin[2];
    const T* __restrict__ Jx = a.J[0];
    const T* __restrict__ Jy = a.j_quirk ? a.J[0] : a.J[1];
    const T* __restrict__ Jz = a.j_quirk ? a.J[0] : a.J[2];

    // Does this CTA's output footprint touch the box where J may be non-zero? (CTA-uniform in i/j.)
    bool j_ij[RJ];
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        j_ij[r] = !a.jbox.empty() && lane_out && row_out[r] && (i < a.jbox.hi[0]) && (i + V > a.jbox.lo[0]) &&
                  (jrow[r] >= a.jbox.lo[1]) && (jrow[r] < a.jbox.hi[1]);
    }

    // ---- carried state -------------------------------------------------------------------------
    T eo[RJ][3][V];      // old E at plane k
    T bp[RJ][2][V];      // new B' (x, y) at plane k-1
#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int e = 0; e < V; ++e) {
            eo[r][0][e] = eo[r][1][e] = eo[r][2][e] = (T)0;
            bp[r][0][e] = bp[r][1][e] = (T)0;
        }

    // Prologue: old E at plane kb-1 (wrapped or ghost plane -1).
    {
        int ks = kb - 1;
        if (ks < 0 && a.g.wrap_k) ks = a.g.nk - 1;
        const long long pk = (long long)ks * a.g.plane;
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
            if (lane_active && row_active[r]) {
                ldg_vec<T, V>(Ex + pk + roff[r], eo[r][0]);
                ldg_vec<T, V>(Ey + pk + roff[r], eo[r][1]);
                ldg_vec<T, V>(Ez + pk + roff[r], eo[r][2]);
            }
        }
        sm.e[(kb - 1) & 1][0][ty][tx] = FusedVec<T>::pack(eo[0][2]);
        sm.e[(kb - 1) & 1][1][ty][tx] = FusedVec<T>::pack(eo[0][0]);
    }
    __syncthreads();

    for (int k = kb - 1; k < ke; ++k) {
        const bool prologue = (k == kb - 1);
        const int par = k & 1;
        int kin = k;                       // plane holding B(k) and old E(k)
        if (kin < 0 && a.g.wrap_k) kin = a.g.nk - 1;
        int kn = k + 1;                    // plane holding old E(k+1)
        if (kn == a.g.nk && a.g.wrap_k) kn = 0;
        const long long pk = (long long)kin * a.g.plane;
        const long long pkn = (long long)kn * a.g.plane;
        const int kg = a.g.k0 + k;
        const bool j_k = !prologue && (kg >= a.jbox.lo[2]) && (kg < a.jbox.hi[2]);

        // ---- loads -------------------------------------------------------------------------------
        T en[RJ][3][V];    // old E at plane k+1
        T b[RJ][3][V];     // B at plane k, becomes B'
        T jv[RJ][3][V];
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                en[r][0][e] = en[r][1][e] = en[r][2][e] = (T)0;
                b[r][0][e] = b[r][1][e] = b[r][2][e] = (T)0;
                jv[r][0][e] = jv[r][1][e] = jv[r][2][e] = (T)0;
            }
            if (lane_active && row_active[r]) {
                ldg_vec<T, V>(Ex + pkn + roff[r], en[r][0]);
                ldg_vec<T, V>(Ey + pkn + roff[r], en[r][1]);
                ldg_vec<T, V>(Ez + pkn + roff[r], en[r][2]);
            }
            if (lane_B && row_B[r]) {
                ldg_vec<T, V>(Bx + pk + roff[r], b[r][0]);
                ldg_vec<T, V>(By + pk + roff[r], b[r][1]);
                ldg_vec<T, V>(Bz + pk + roff[r], b[r][2]);
            }
            if (j_ij[r] && j_k) {
                ldg_vec<T, V>(Jx + pk + roff[r], jv[r][0]);
                ldg_vec<T, V>(Jy + pk + roff[r], jv[r][1]);
                ldg_vec<T, V>(Jz + pk + roff[r], jv[r][2]);
            }
        }

        // ---- B'(k) = B(k) + n_half * h(E_old) ----------------------------------------------------
        // old E(k) of the row above the last row of this warp comes from the next warp's first row
        T ezu[V], exu[V];
        {
            const int tyn = (ty + 1 < BY) ? ty + 1 : ty;
            FusedVec<T>::unpack(sm.e[par][0][tyn][tx], ezu);
            FusedVec<T>::unpack(sm.e[par][1][tyn][tx], exu);
        }
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
            // i+1 neighbours of the last element come from the next lane's first element
            const T ez_nl = __shfl_down_sync(FULL, eo[r][2][0], 1);
            const T ey_nl = __shfl_down_sync(FULL, eo[r][1][0], 1);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const double ex = (double)eo[r][0][e], ey = (double)eo[r][1][e], ez = (double)eo[r][2][e];
                const double ezr = (double)((e == V - 1) ? ez_nl : eo[r][2][(e + 1) % V]);
                const double eyr = (double)((e == V - 1) ? ey_nl : eo[r][1][(e + 1) % V]);
                const double ezj = (double)((r == RJ - 1) ? ezu[e] : eo[(r + 1) % RJ][2][e]);
                const double exj = (double)((r == RJ - 1) ? exu[e] : eo[(r + 1) % RJ][0][e]);
                const double eyk = (double)en[r][1][e], exk = (double)en[r][0][e];
                // FDTD.cpp:121-126
                const double hx = dsub(dmul(cBz, dsub(eyk, ey)), dmul(cBy, dsub(ezj, ez)));
                const double hy = dsub(dmul(cBx, dsub(ezr, ez)), dmul(cBz, dsub(exk, ex)));
                const double hz = dsub(dmul(cBy, dsub(exj, ex)), dmul(cBx, dsub(eyr, ey)));
                T nbx = (T)dadd((double)b[r][0][e], hx);
                T nby = (T)dadd((double)b[r][1][e], hy);
                T nbz = (T)dadd((double)b[r][2][e], hz);
                if (a.n_half == 2) {
                    nbx = (T)dadd((double)nbx, hx);
                    nby = (T)dadd((double)nby, hy);
                    nbz = (T)dadd((double)nbz, hz);
                }
                b[r][0][e] = nbx; b[r][1][e] = nby; b[r][2][e] = nbz;
            }
        }

        // ---- publish rows for the neighbouring warps, one barrier per plane ---------------------------
        sm.b[par][0][ty][tx] = FusedVec<T>::pack(b[RJ - 1][2]);
        sm.b[par][1][ty][tx] = FusedVec<T>::pack(b[RJ - 1][0]);
        sm.e[par ^ 1][0][ty][tx] = FusedVec<T>::pack(en[0][2]);
        sm.e[par ^ 1][1][ty][tx] = FusedVec<T>::pack(en[0][0]);
        __syncthreads();

        // ---- E'(k) = E(k) + g(B'(k), B'(k-1), J) and the stores --------------------------------------
        if (!prologue) {
            T bzd[V], bxd[V];     // B' of the row below the first row of this warp
            {
                const int typ = (ty > 0) ? ty - 1 : ty;
                FusedVec<T>::unpack(sm.b[par][0][typ][tx], bzd);
                FusedVec<T>::unpack(sm.b[par][1][typ][tx], bxd);
            }
#pragma unroll
            for (int r = 0; r < RJ; ++r) {
                const T bz_pl = __shfl_up_sync(FULL, b[r][2][V - 1], 1);
                const T by_pl = __shfl_up_sync(FULL, b[r][1][V - 1], 1);
                T ne[3][V];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const double bx = (double)b[r][0][e], by = (double)b[r][1][e], bz = (double)b[r][2][e];
                    const double bzl = (double)((e == 0) ? bz_pl : b[r][2][(e + V - 1) % V]);
                    const double byl = (double)((e == 0) ? by_pl : b[r][1][(e + V - 1) % V]);
                    const double bzj = (double)((r == 0) ? bzd[e] : b[(r + RJ - 1) % RJ][2][e]);
                    const double bxj = (double)((r == 0) ? bxd[e] : b[(r + RJ - 1) % RJ][0][e]);
                    const double byk = (double)bp[r][1][e], bxk = (double)bp[r][0][e];
                    // FDTD.cpp:85-93 / kokkos_functors.h:81-89
                    double tx_ = dmul(cEy, dsub(bz, bzj));
                    double ty_ = dmul(cEz, dsub(bx, bxk));
                    double tz_ = dmul(cEx, dsub(by, byl));
                    if (j_ij[r] && j_k) {
                        tx_ = dadd(dmul(cJ, (double)jv[r][0][e]), tx_);
                        ty_ = dadd(dmul(cJ, (double)jv[r][1][e]), ty_);
                        tz_ = dadd(dmul(cJ, (double)jv[r][2][e]), tz_);
                    }
                    ne[0][e] = (T)dadd((double)eo[r][0][e], dsub(tx_, dmul(cEz, dsub(by, byk))));
                    ne[1][e] = (T)dadd((double)eo[r][1][e], dsub(ty_, dmul(cEx, dsub(bz, bzl))));
                    ne[2][e] = (T)dadd((double)eo[r][2][e], dsub(tz_, dmul(cEy, dsub(bx, bxj))));
                }
                if (lane_out && row_out[r]) {
                    const long long o = pk + roff[r];
                    stg_vec<T, V>(a.Eout[0] + o, ne[0]);
                    stg_vec<T, V>(a.Eout[1] + o, ne[1]);
                    stg_vec<T, V>(a.Eout[2] + o, ne[2]);
                    stg_vec<T, V>(a.Bout[0] + o, b[r][0]);
                    stg_vec<T, V>(a.Bout[1] + o, b[r][1]);
                    stg_vec<T, V>(a.Bout[2] + o, b[r][2]);
                }
            }
        }

        // ---- carry ----------------------------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int e = 0; e < V; ++e) {
                eo[r][0][e] = en[r][0][e]; eo[r][1][e] = en[r][1][e]; eo[r][2][e] = en[r][2][e];
                bp[r][0][e] = b[r][0][e]; bp[r][1][e] = b[r][1][e];
            }
    }
}

}  // namespace fdtd_b200
