// fused_kernel_v2.cuh -- the fused B+E pass with the plane loads taken out of the critical path.
//
// Same algorithm, tile roles and bit-exact arithmetic as fused_BE_kernel (fused_kernel.cuh, one row per
// warp); what changes is HOW the six 16-byte vectors a thread needs per plane (old E(k+1) x3, B(k) x3)
// reach it.  ncu on v1 (profiles/ncu_summary_r01.md) shows the pass is latency-bound (long_scoreboard 8.9 and
// barrier 4.3 stall cycles per issue at 37 % occupancy, DRAM at 0.83 of the measured peak), so:
//
//   PF = 1  register prefetch: the loads of plane k+1 are issued before the arithmetic of plane k.
//   PF = 2  cp.async ring (LDGSTS, L1-bypassing .cg): every thread copies its own six vectors for plane
//           k+D-1 into a private slot of a D-deep shared-memory ring while it computes plane k; it only ever
//           reads what it copied itself, so the ring needs no barrier -- just cp.async.wait_group.
//           In-flight bytes per SM = CTAs/SM x 256 threads x 96 B x (D-1), with no register cost.
//
// Replaces reference src/FDTD/FDTD.cpp:153-157 (update_B, update_E, update_B) like v1.
#pragma once

#include "fused_kernel.cuh"

namespace fdtd_b200 {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int BY, int PF, int D>
constexpr size_t fused2_smem_bytes() {
    return (size_t)(2 * 2 * BY * FUSED_BX * 2 + (PF == 2 ? D * 6 * BY * FUSED_BX : 0)) * 16;
}

template <typename T, int BY, int PF, int D, int MINB>
__global__ void __launch_bounds__(FUSED_BX * BY, MINB) fused_BE2_kernel(const FusedArgs<T> a) {
    constexpr int V = VecOf<T>::V;
    constexpr int TJU = BY - 2;
    constexpr int TIU = FUSED_OUT_LANES * V;
    constexpr unsigned FULL = 0xffffffffu;
    using VT = typename FusedVec<T>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    VT* const sE = reinterpret_cast<VT*>(smem_raw);            // [2 buffers][2 comps][BY][32]
    VT* const sB = sE + 2 * 2 * BY * FUSED_BX;                 // [2 buffers][2 comps][BY][32]
    VT* const ring = sB + 2 * 2 * BY * FUSED_BX;               // [D][6][BY][32]   (PF == 2)

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int Ni = a.g.Ni, Nj = a.g.Nj;
    const int tid = ty * FUSED_BX + tx;
    auto xs = [&](VT* base, int buf, int comp, int row) -> VT& { return base[((buf * 2 + comp) * BY + row) * FUSED_BX + tx]; };

    // ---- roles (identical to v1 with RJ = 1) -----------------------------------------------------------
    const int i = blockIdx.x * TIU - V + tx * V;
    const bool lane_active = (i <= Ni);
    const int iw = (i < 0) ? i + Ni : ((i == Ni) ? 0 : i);
    const bool lane_B = lane_active && (tx <= FUSED_OUT_LANES) && (i < Ni);
    const bool lane_out = lane_B && (tx >= 1);
    const int j = blockIdx.y * TJU - 1 + ty;
    const bool row_active = (j <= Nj);
    const int jw = (j < 0) ? Nj - 1 : ((j == Nj) ? 0 : j);
    const bool row_B = row_active && (ty <= BY - 2) && (j < Nj);
    const bool row_out = row_B && (ty >= 1);
    const bool ldE = lane_active && row_active;
    const bool ldB = lane_B && row_B;
    const bool out = lane_out && row_out;
    const long long roff = (long long)jw * a.g.pitch + iw;

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
    const bool j_ij = !a.jbox.empty() && out && (i < a.jbox.hi[0]) && (i + V > a.jbox.lo[0]) &&
                      (jw >= a.jbox.lo[1]) && (jw < a.jbox.hi[1]);

    // plane offsets of iteration k: B(k) / old E(k) live in plane kin, old E(k+1) in plane kn
    auto plane_in = [&](int k) -> long long {
        int kin = k;
        if (kin < 0 && a.g.wrap_k) kin = a.g.nk - 1;
        return (long long)kin * a.g.plane;
    };
    auto plane_next = [&](int k) -> long long {
        int kn = k + 1;
        if (kn == a.g.nk && a.g.wrap_k) kn = 0;
        return (long long)kn * a.g.plane;
    };
    auto zero3 = [&](T (&x)[3][V]) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int e = 0; e < V; ++e) x[c][e] = (T)0;
    };
    // direct (register) loads of iteration k
    auto load_regs = [&](int k, T (&en)[3][V], T (&b)[3][V]) {
        const long long pk = plane_in(k) + roff, pkn = plane_next(k) + roff;
        if (ldE) {
            ldg_vec<T, V>(Ex + pkn, en[0]);
            ldg_vec<T, V>(Ey + pkn, en[1]);
            ldg_vec<T, V>(Ez + pkn, en[2]);
        }
        if (ldB) {
            ldg_vec<T, V>(Bx + pk, b[0]);
            ldg_vec<T, V>(By + pk, b[1]);
            ldg_vec<T, V>(Bz + pk, b[2]);
        }
    };
    // asynchronous copies of iteration k into ring slot `slot`
    auto issue_ring = [&](int k, int slot) {
        const long long pk = plane_in(k) + roff, pkn = plane_next(k) + roff;
        VT* s = ring + (size_t)slot * 6 * BY * FUSED_BX + tid;
        if (ldE) {
            cp_async16(s + 0 * BY * FUSED_BX, Ex + pkn);
            cp_async16(s + 1 * BY * FUSED_BX, Ey + pkn);
            cp_async16(s + 2 * BY * FUSED_BX, Ez + pkn);
        }
        if (ldB) {
            cp_async16(s + 3 * BY * FUSED_BX, Bx + pk);
            cp_async16(s + 4 * BY * FUSED_BX, By + pk);
            cp_async16(s + 5 * BY * FUSED_BX, Bz + pk);
        }
    };
    auto read_ring = [&](int slot, T (&en)[3][V], T (&b)[3][V]) {
        const VT* s = ring + (size_t)slot * 6 * BY * FUSED_BX + tid;
        if (ldE) {
            FusedVec<T>::unpack(s[0 * BY * FUSED_BX], en[0]);
            FusedVec<T>::unpack(s[1 * BY * FUSED_BX], en[1]);
            FusedVec<T>::unpack(s[2 * BY * FUSED_BX], en[2]);
        }
        if (ldB) {
            FusedVec<T>::unpack(s[3 * BY * FUSED_BX], b[0]);
            FusedVec<T>::unpack(s[4 * BY * FUSED_BX], b[1]);
            FusedVec<T>::unpack(s[5 * BY * FUSED_BX], b[2]);
        }
    };

    // ---- carried state -----------------------------------------------------------------------------------
    T eo[3][V];      // old E at plane k
    T bp[2][V];      // new B' (x, y) at plane k-1
    zero3(eo);
#pragma unroll
    for (int e = 0; e < V; ++e) bp[0][e] = bp[1][e] = (T)0;

    // prologue: old E(kb-1); first stages / prefetch
    T pen[3][V], pb[3][V];   // PF == 1: registers holding the NEXT iteration's loads
    zero3(pen); zero3(pb);
    if (PF == 2) {
#pragma unroll
        for (int d = 0; d < D - 1; ++d) {
            if (kb - 1 + d < ke) issue_ring(kb - 1 + d, d);
            cp_async_commit();
        }
    }
    if (ldE) {
        const long long pk = plane_in(kb - 1) + roff;
        ldg_vec<T, V>(Ex + pk, eo[0]);
        ldg_vec<T, V>(Ey + pk, eo[1]);
        ldg_vec<T, V>(Ez + pk, eo[2]);
    }
    if (PF == 1) load_regs(kb - 1, pen, pb);
    xs(sE, (kb - 1) & 1, 0, ty) = FusedVec<T>::pack(eo[2]);
    xs(sE, (kb - 1) & 1, 1, ty) = FusedVec<T>::pack(eo[0]);
    __syncthreads();

    int slot = 0;   // ring slot of iteration k (PF == 2)
    for (int k = kb - 1; k < ke; ++k) {
        const bool prologue = (k == kb - 1);
        const int par = k & 1;
        const int kg = a.g.k0 + k;
        const bool use_j = j_ij && !prologue && (kg >= a.jbox.lo[2]) && (kg < a.jbox.hi[2]);

        // ---- this plane's inputs --------------------------------------------------------------------------
        T en[3][V], b[3][V], jv[3][V];
        zero3(en); zero3(b); zero3(jv);
        if (PF == 0) {
            load_regs(k, en, b);
        } else if (PF == 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int e = 0; e < V; ++e) { en[c][e] = pen[c][e]; b[c][e] = pb[c][e]; }
            if (k + 1 < ke) load_regs(k + 1, pen, pb);
        } else {
            int nslot = slot + (D - 1);
            if (nslot >= D) nslot -= D;
            if (k + D - 1 < ke) issue_ring(k + D - 1, nslot);
            cp_async_commit();
            cp_async_wait<D - 1>();
            read_ring(slot, en, b);
            slot = (slot + 1 == D) ? 0 : slot + 1;
        }
        if (use_j) {
            const long long pk = plane_in(k) + roff;
            ldg_vec<T, V>(Jx + pk, jv[0]);
            ldg_vec<T, V>(Jy + pk, jv[1]);
            ldg_vec<T, V>(Jz + pk, jv[2]);
        }

        // ---- B'(k) = B(k) + n_half * h(E_old)  (FDTD.cpp:121-126) ---------------------------------------
        T ezu[V], exu[V];   // old E(k) one row up: the next warp's row
        {
            const int tyn = (ty + 1 < BY) ? ty + 1 : ty;
            FusedVec<T>::unpack(xs(sE, par, 0, tyn), ezu);
            FusedVec<T>::unpack(xs(sE, par, 1, tyn), exu);
        }
        const T ez_nl = __shfl_down_sync(FULL, eo[2][0], 1);
        const T ey_nl = __shfl_down_sync(FULL, eo[1][0], 1);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const double ex = (double)eo[0][e], ey = (double)eo[1][e], ez = (double)eo[2][e];
            const double ezr = (double)((e == V - 1) ? ez_nl : eo[2][(e + 1) % V]);
            const double eyr = (double)((e == V - 1) ? ey_nl : eo[1][(e + 1) % V]);
            const double hx = dsub(dmul(cBz, dsub((double)en[1][e], ey)), dmul(cBy, dsub((double)ezu[e], ez)));
            const double hy = dsub(dmul(cBx, dsub(ezr, ez)), dmul(cBz, dsub((double)en[0][e], ex)));
            const double hz = dsub(dmul(cBy, dsub((double)exu[e], ex)), dmul(cBx, dsub(eyr, ey)));
            T nbx = (T)dadd((double)b[0][e], hx);
            T nby = (T)dadd((double)b[1][e], hy);
            T nbz = (T)dadd((double)b[2][e], hz);
            if (a.n_half == 2) {
                nbx = (T)dadd((double)nbx, hx);
                nby = (T)dadd((double)nby, hy);
                nbz = (T)dadd((double)nbz, hz);
            }
            b[0][e] = nbx; b[1][e] = nby; b[2][e] = nbz;
        }

        // ---- publish rows for the neighbouring warps, one barrier per plane -----------------------------
        xs(sB, par, 0, ty) = FusedVec<T>::pack(b[2]);
        xs(sB, par, 1, ty) = FusedVec<T>::pack(b[0]);
        xs(sE, par ^ 1, 0, ty) = FusedVec<T>::pack(en[2]);
        xs(sE, par ^ 1, 1, ty) = FusedVec<T>::pack(en[0]);
        __syncthreads();

        // ---- E'(k) = E(k) + g(B'(k), B'(k-1), J)  (FDTD.cpp:85-93 / kokkos_functors.h:81-89) --------------
        if (!prologue) {
            T bzd[V], bxd[V];   // B' one row down: the previous warp's row
            {
                const int typ = (ty > 0) ? ty - 1 : ty;
                FusedVec<T>::unpack(xs(sB, par, 0, typ), bzd);
                FusedVec<T>::unpack(xs(sB, par, 1, typ), bxd);
            }
            const T bz_pl = __shfl_up_sync(FULL, b[2][V - 1], 1);
            const T by_pl = __shfl_up_sync(FULL, b[1][V - 1], 1);
            T ne[3][V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const double bx = (double)b[0][e], by = (double)b[1][e], bz = (double)b[2][e];
                const double bzl = (double)((e == 0) ? bz_pl : b[2][(e + V - 1) % V]);
                const double byl = (double)((e == 0) ? by_pl : b[1][(e + V - 1) % V]);
                double tx_ = dmul(cEy, dsub(bz, (double)bzd[e]));
                double ty_ = dmul(cEz, dsub(bx, (double)bp[0][e]));
                double tz_ = dmul(cEx, dsub(by, byl));
                if (use_j) {
                    tx_ = dadd(dmul(cJ, (double)jv[0][e]), tx_);
                    ty_ = dadd(dmul(cJ, (double)jv[1][e]), ty_);
                    tz_ = dadd(dmul(cJ, (double)jv[2][e]), tz_);
                }
                ne[0][e] = (T)dadd((double)eo[0][e], dsub(tx_, dmul(cEz, dsub(by, (double)bp[1][e]))));
                ne[1][e] = (T)dadd((double)eo[1][e], dsub(ty_, dmul(cEx, dsub(bz, bzl))));
                ne[2][e] = (T)dadd((double)eo[2][e], dsub(tz_, dmul(cEy, dsub(bx, (double)bxd[e]))));
            }
            if (out) {
                const long long o = plane_in(k) + roff;
                stg_vec<T, V>(a.Eout[0] + o, ne[0]);
                stg_vec<T, V>(a.Eout[1] + o, ne[1]);
                stg_vec<T, V>(a.Eout[2] + o, ne[2]);
                stg_vec<T, V>(a.Bout[0] + o, b[0]);
                stg_vec<T, V>(a.Bout[1] + o, b[1]);
                stg_vec<T, V>(a.Bout[2] + o, b[2]);
            }
        }

        // ---- carry -------------------------------------------------------------------------------------------
#pragma unroll
        for (int e = 0; e < V; ++e) {
            eo[0][e] = en[0][e]; eo[1][e] = en[1][e]; eo[2][e] = en[2][e];
            bp[0][e] = b[0][e]; bp[1][e] = b[1][e];
        }
    }
    if (PF == 2) cp_async_wait<0>();
}

}  // namespace fdtd_b200
