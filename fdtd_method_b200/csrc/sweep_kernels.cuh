// sweep_kernels.cuh -- the general two-sweep path: one B sweep and one E sweep per Yee step.
//
// These kernels serve every configuration (periodic / PML, fp64 / fp32, single GPU / z-slab rank);
// the fused E+B pass in fused_kernel.cuh is the fast path for the periodic solver.
//
//   sweep_B_kernel  replaces FDTD::update_B           (reference src/FDTD/FDTD.cpp:99-130,
//                                                       include/FDTD_kokkos/kokkos_functors.h:128-151)
//                   and     FDTD_PML::update_B_PML x6 (src/FDTD/FDTD_PML.cpp:136-203, call sites :346-352)
//   sweep_E_kernel  replaces FDTD::update_E           (src/FDTD/FDTD.cpp:63-97, kokkos_functors.h:64-90)
//                   and     FDTD_PML::update_E_PML x6 (src/FDTD/FDTD_PML.cpp:67-134, call sites :356-362)
//
// Differences from the reference's structure, none of which change a single bit of the result:
//   * the reference's trailing B half step of step s and leading B half step of step s+1 see the same
//     E, so they are applied together as B = (B + h) + h (n_half = 2) -- two sweeps per step, not three;
//   * the six PML shell boxes and the main box are one launch with a per-cell region predicate
//     (each phase only reads the other field family);
//   * sigma is a function of one coordinate, so exp()/division live in host-computed 1-D tables;
//   * the current term is skipped outside the bounding box where J may be non-zero
//     (cJ * (+0.0) = -0.0 and x + (-0.0) == x, so the skip is exact).
//
// Mapping: one thread owns V = 16 B / sizeof(T) consecutive cells in i (one 128-bit load/store per
// array), a warp covers 32*V contiguous cells of one row, a CTA 8 rows, and each thread streams a
// chunk of k planes keeping the k-neighbour in registers.  j/k neighbours are whole-row offsets;
// the i neighbour of the last lane is one extra (L1-resident) scalar load.
#pragma once

#include "fdtd_common.cuh"

namespace fdtd_b200 {

template <typename T>
struct SweepArgs {
    Geom g;
    Coefs c;
    PmlDesc p;
    Fields<T> f;
    JBox jbox;
    int k_lo, k_hi;     // local plane range [k_lo, k_hi) this launch covers
    int kc;             // planes per thread (k chunk)
    int n_half;         // B sweep: 1 or 2 half steps on main cells (0: leave main cells untouched)
    int do_pml;         // B sweep: also advance the PML shell (full step)
    int j_quirk;        // E sweep: FDTD_openmp semantics, Jx feeds all three components
    // PML solvers split every sweep into two launches so that the (large) interior runs the lean PML=false
    // instantiation at full occupancy: mode 1 = only cells inside the inner box `ib` (the main box with its
    // i bounds aligned inward to the vector width), mode 2 = only cells outside it, mode 0 = every cell.
    int mode;
    int ib_lo[3], ib_hi[3];   // inner box, GLOBAL coordinates
    // Where the sweep writes: f.B / f.E themselves (in place, the default) or the other generation (the second
    // step of the PML solver's two-step pass, whose core cells the T2 pass has already written there).
    T* Bout[3];
    T* Eout[3];
};

constexpr int SWEEP_BX = 32;  // lanes along i
constexpr int SWEEP_BY = 8;   // rows along j

// V = cells per thread: 16 bytes' worth by default; the float PML instantiation takes 2 (8-byte accesses, still whole
// 32-byte sectors per 4 lanes) -- with 4 cells of double arithmetic plus split fields it needs 194-224 registers and
// runs one CTA per SM.
// A = arithmetic type (double; float only with FDTD_FLAG_F32_ARITH, where T = float too).
template <typename T, bool PML, int V = VecOf<T>::V, typename A = double>
__global__ void __launch_bounds__(SWEEP_BX * SWEEP_BY) sweep_B_kernel(const SweepArgs<T> a) {
    const int Ni = a.g.Ni, Nj = a.g.Nj;
    const int i0 = (blockIdx.x * SWEEP_BX + threadIdx.x) * V;
    const int j = blockIdx.y * SWEEP_BY + threadIdx.y;
    if (i0 >= Ni || j >= Nj) return;
    int kb = a.k_lo + blockIdx.z * a.kc;
    int ke = min(kb + a.kc, a.k_hi);
    const bool in_ij = (a.mode != 0) && (i0 >= a.ib_lo[0]) && (i0 + V <= a.ib_hi[0]) && (j >= a.ib_lo[1]) && (j < a.ib_hi[1]);
    if (a.mode == 1) {
        if (!in_ij) return;
        kb = max(kb, a.ib_lo[2] - a.g.k0);
        ke = min(ke, a.ib_hi[2] - a.g.k0);
    }
    if (kb >= ke) return;

    const int nvalid = min(V, Ni - i0);
    const int jn = (j + 1 == Nj) ? 0 : j + 1;
    const int cn = (i0 + V < Ni) ? i0 + V : 0;   // column right of this thread's last cell (wrapped)
    const long long row = (long long)j * a.g.pitch;
    const long long rown = (long long)jn * a.g.pitch;

    const T* __restrict__ Ex = a.f.E[0];
    const T* __restrict__ Ey = a.f.E[1];
    const T* __restrict__ Ez = a.f.E[2];
    const T* Bx = a.f.B[0];   // (no __restrict__: Bout aliases f.B when the sweep is in place)
    const T* By = a.f.B[1];
    const T* Bz = a.f.B[2];
    T* BxO = a.Bout[0];
    T* ByO = a.Bout[1];
    T* BzO = a.Bout[2];
    const A cx = a.c.cBx, cy = a.c.cBy, cz = a.c.cBz;

    bool col_main[V];
    A dcx[V], c2x[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        col_main[e] = true;
        dcx[e] = 1.0; c2x[e] = 0.0;
        if (PML) {
            const int i = min(i0 + e, Ni - 1);
            col_main[e] = (i >= a.p.lo[0] && i < a.p.hi[0]);
            dcx[e] = a.p.decay[0][i];
            c2x[e] = a.p.coef2[0][i];
        }
    }
    bool jrow_main = true;
    A dcy = 1.0, c2y = 0.0;
    if (PML) {
        jrow_main = (j >= a.p.lo[1] && j < a.p.hi[1]);
        dcy = a.p.decay[1][j];
        c2y = a.p.coef2[1][j];
    }

    A ex[V], ey[V];
    {
        const long long o = (long long)kb * a.g.plane + row + i0;
        ldv(Ex + o, ex);
        ldv(Ey + o, ey);
    }
    bool carry_ok = true;
    for (int k = kb; k < ke; ++k) {
        int kn = k + 1;
        if (kn == a.g.nk && a.g.wrap_k) kn = 0;
        const long long pk = (long long)k * a.g.plane;
        const long long pkn = (long long)kn * a.g.plane;
        const long long o = pk + row + i0;
        if (a.mode == 2 && in_ij && (a.g.k0 + k >= a.ib_lo[2]) && (a.g.k0 + k < a.ib_hi[2])) {
            carry_ok = false;      // this cell belongs to the mode-1 launch
            continue;
        }

        bool row_main = jrow_main;
        A dcz = 1.0, c2z = 0.0;
        if (PML) {
            const int kg = a.g.k0 + k;
            row_main = row_main && (kg >= a.p.lo[2] && kg < a.p.hi[2]);
            dcz = a.p.decay[2][kg];
            c2z = a.p.coef2[2][kg];
        }
        bool any_pml = false, all_pml = PML;
        if (PML) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool cell_pml = !(row_main && col_main[e]);
                any_pml = any_pml || (e < nvalid && cell_pml);
                all_pml = all_pml && (e >= nvalid || cell_pml);
            }
            if (all_pml && !a.do_pml) {   // deferred half step: shell cells stay as they are
                carry_ok = false;
                continue;
            }
            any_pml = any_pml && a.do_pml;
        }
        if (!carry_ok) {
            ldv(Ex + o, ex);
            ldv(Ey + o, ey);
            carry_ok = true;
        }

        A exn[V], eyn[V], ez[V], ezj[V], exj[V], bx[V], by[V], bz[V];
        ldv(Ex + pkn + row + i0, exn);
        ldv(Ey + pkn + row + i0, eyn);
        ldv(Ez + o, ez);
        ldv(Ez + pk + rown + i0, ezj);
        ldv(Ex + pk + rown + i0, exj);
        const A ez_r = lds1<A>(Ez + pk + row + cn);
        const A ey_r = lds1<A>(Ey + pk + row + cn);
        if (!all_pml) {   // a shell cell's B is the sum of its split fields (FDTD_PML.cpp:197-199): the old value is never read
            ldv(Bx + o, bx);
            ldv(By + o, by);
            ldv(Bz + o, bz);
        }
        A sxy[V], sxz[V], syx[V], syz[V], szx[V], szy[V];
        if (PML && any_pml) {
            ldv(a.f.SB[S_XY] + o, sxy); ldv(a.f.SB[S_XZ] + o, sxz);
            ldv(a.f.SB[S_YX] + o, syx); ldv(a.f.SB[S_YZ] + o, syz);
            ldv(a.f.SB[S_ZX] + o, szx); ldv(a.f.SB[S_ZY] + o, szy);
        }

#pragma unroll
        for (int e = 0; e < V; ++e) {
            // right (i+1) neighbours: own next element, or the wrapped/next-thread scalar
            const bool use_scalar = (e == V - 1) || (i0 + e + 1 == Ni);
            const A ezr = use_scalar ? ez_r : ez[(e + 1) % V];
            const A eyr = use_scalar ? ey_r : ey[(e + 1) % V];
            const A dEy_k = dsub(eyn[e], ey[e]);
            const A dEz_j = dsub(ezj[e], ez[e]);
            const A dEz_i = dsub(ezr, ez[e]);
            const A dEx_k = dsub(exn[e], ex[e]);
            const A dEx_j = dsub(exj[e], ex[e]);
            const A dEy_i = dsub(eyr, ey[e]);
            if (!PML || (row_main && col_main[e])) {
                // FDTD.cpp:121-126
                const A hx = curl2(cz, dEy_k, cy, dEz_j);
                const A hy = curl2(cx, dEz_i, cz, dEx_k);
                const A hz = curl2(cy, dEx_j, cx, dEy_i);
                if (a.n_half >= 1) {
                    bx[e] = round_store<T>(dadd(bx[e], hx));
                    by[e] = round_store<T>(dadd(by[e], hy));
                    bz[e] = round_store<T>(dadd(bz[e], hz));
                }
                if (a.n_half >= 2) {
                    bx[e] = round_store<T>(dadd(bx[e], hx));
                    by[e] = round_store<T>(dadd(by[e], hy));
                    bz[e] = round_store<T>(dadd(bz[e], hz));
                }
            } else if (PML && a.do_pml) {
                // FDTD_PML.cpp:182-199
                syx[e] = round_store<T>(dadd(dmul(syx[e], dcx[e]), dmul(c2x[e], dEz_i)));
                szx[e] = round_store<T>(dsub(dmul(szx[e], dcx[e]), dmul(c2x[e], dEy_i)));
                sxy[e] = round_store<T>(dsub(dmul(sxy[e], dcy), dmul(c2y, dEz_j)));
                szy[e] = round_store<T>(dadd(dmul(szy[e], dcy), dmul(c2y, dEx_j)));
                sxz[e] = round_store<T>(dadd(dmul(sxz[e], dcz), dmul(c2z, dEy_k)));
                syz[e] = round_store<T>(dsub(dmul(syz[e], dcz), dmul(c2z, dEx_k)));
                bx[e] = round_store<T>(dadd(sxy[e], sxz[e]));
                by[e] = round_store<T>(dadd(syz[e], syx[e]));
                bz[e] = round_store<T>(dadd(szx[e], szy[e]));
            }
        }
        stv(BxO + o, bx, nvalid);
        stv(ByO + o, by, nvalid);
        stv(BzO + o, bz, nvalid);
        if (PML && any_pml) {
            stv(a.f.SB[S_XY] + o, sxy, nvalid); stv(a.f.SB[S_XZ] + o, sxz, nvalid);
            stv(a.f.SB[S_YX] + o, syx, nvalid); stv(a.f.SB[S_YZ] + o, syz, nvalid);
            stv(a.f.SB[S_ZX] + o, szx, nvalid); stv(a.f.SB[S_ZY] + o, szy, nvalid);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) { ex[e] = exn[e]; ey[e] = eyn[e]; }
    }
}

template <typename T, bool PML, int V = VecOf<T>::V, typename A = double>
__global__ void __launch_bounds__(SWEEP_BX * SWEEP_BY) sweep_E_kernel(const SweepArgs<T> a) {
    const int Ni = a.g.Ni, Nj = a.g.Nj;
    const int i0 = (blockIdx.x * SWEEP_BX + threadIdx.x) * V;
    const int j = blockIdx.y * SWEEP_BY + threadIdx.y;
    if (i0 >= Ni || j >= Nj) return;
    int kb = a.k_lo + blockIdx.z * a.kc;
    int ke = min(kb + a.kc, a.k_hi);
    const bool in_ij = (a.mode != 0) && (i0 >= a.ib_lo[0]) && (i0 + V <= a.ib_hi[0]) && (j >= a.ib_lo[1]) && (j < a.ib_hi[1]);
    if (a.mode == 1) {
        if (!in_ij) return;
        kb = max(kb, a.ib_lo[2] - a.g.k0);
        ke = min(ke, a.ib_hi[2] - a.g.k0);
    }
    if (kb >= ke) return;

    const int nvalid = min(V, Ni - i0);
    const int jp = (j == 0) ? Nj - 1 : j - 1;
    const int cp = (i0 == 0) ? Ni - 1 : i0 - 1;   // column left of this thread's first cell (wrapped)
    const long long row = (long long)j * a.g.pitch;
    const long long rowp = (long long)jp * a.g.pitch;

    const T* Ex = a.f.E[0];   // (no __restrict__: Eout aliases f.E when the sweep is in place)
    const T* Ey = a.f.E[1];
    const T* Ez = a.f.E[2];
    T* ExO = a.Eout[0];
    T* EyO = a.Eout[1];
    T* EzO = a.Eout[2];
    const T* __restrict__ Bx = a.f.B[0];
    const T* __restrict__ By = a.f.B[1];
    const T* __restrict__ Bz = a.f.B[2];
    const T* __restrict__ Jx = a.f.J[0];
    const T* __restrict__ Jy = a.j_quirk ? a.f.J[0] : a.f.J[1];
    const T* __restrict__ Jz = a.j_quirk ? a.f.J[0] : a.f.J[2];
    const A cx = a.c.cEx, cy = a.c.cEy, cz = a.c.cEz, cj = a.c.cJ;

    bool col_main[V];
    A dcx[V], c2x[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        col_main[e] = true;
        dcx[e] = 1.0; c2x[e] = 0.0;
        if (PML) {
            const int i = min(i0 + e, Ni - 1);
            col_main[e] = (i >= a.p.lo[0] && i < a.p.hi[0]);
            dcx[e] = a.p.decay[0][i];
            c2x[e] = a.p.coef2[0][i];
        }
    }
    bool jrow_main = true;
    A dcy = 1.0, c2y = 0.0;
    if (PML) {
        jrow_main = (j >= a.p.lo[1] && j < a.p.hi[1]);
        dcy = a.p.decay[1][j];
        c2y = a.p.coef2[1][j];
    }
    // Does this thread's (i, j) footprint touch the box where J may be non-zero?
    const bool j_ij = !a.jbox.empty() && (i0 < a.jbox.hi[0] && i0 + V > a.jbox.lo[0]) &&
                      (j >= a.jbox.lo[1] && j < a.jbox.hi[1]);

    A bxm[V], bym[V];   // B at plane k-1
    {
        int km = kb - 1;
        if (km < 0 && a.g.wrap_k) km = a.g.nk - 1;
        const long long o = (long long)km * a.g.plane + row + i0;
        ldv(Bx + o, bxm);
        ldv(By + o, bym);
    }
    bool carry_ok = true;
    for (int k = kb; k < ke; ++k) {
        const long long pk = (long long)k * a.g.plane;
        const long long o = pk + row + i0;
        const int kg = a.g.k0 + k;
        if (a.mode == 2 && in_ij && (kg >= a.ib_lo[2]) && (kg < a.ib_hi[2])) {
            carry_ok = false;      // this cell belongs to the mode-1 launch
            continue;
        }
        if (!carry_ok) {
            int km = k - 1;
            if (km < 0 && a.g.wrap_k) km = a.g.nk - 1;
            const long long om = (long long)km * a.g.plane + row + i0;
            ldv(Bx + om, bxm);
            ldv(By + om, bym);
            carry_ok = true;
        }

        A bx[V], by[V], bz[V], bzj[V], bxj[V], e_x[V], e_y[V], e_z[V];
        ldv(Bx + o, bx);
        ldv(By + o, by);
        ldv(Bz + o, bz);
        ldv(Bz + pk + rowp + i0, bzj);
        ldv(Bx + pk + rowp + i0, bxj);
        const A bz_l = lds1<A>(Bz + pk + row + cp);
        const A by_l = lds1<A>(By + pk + row + cp);
        bool row_main = jrow_main;
        A dcz = 1.0, c2z = 0.0;
        if (PML) {
            row_main = row_main && (kg >= a.p.lo[2] && kg < a.p.hi[2]);
            dcz = a.p.decay[2][kg];
            c2z = a.p.coef2[2][kg];
        }
        bool any_pml = false, all_pml = PML;
        if (PML) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool cell_pml = !(row_main && col_main[e]);
                any_pml = any_pml || (e < nvalid && cell_pml);
                all_pml = all_pml && (e >= nvalid || cell_pml);
            }
        }
        if (!all_pml) {   // a shell cell's E is the sum of its split fields (FDTD_PML.cpp:128-130): the old value is never read
            ldv(Ex + o, e_x);
            ldv(Ey + o, e_y);
            ldv(Ez + o, e_z);
        }

        const bool use_j = !all_pml && j_ij && (kg >= a.jbox.lo[2] && kg < a.jbox.hi[2]);
        A jx[V], jy[V], jz[V];
        if (use_j) {
            ldv(Jx + o, jx);
            ldv(Jy + o, jy);
            ldv(Jz + o, jz);
        }
        A sxy[V], sxz[V], syx[V], syz[V], szx[V], szy[V];
        if (PML && any_pml) {
            ldv(a.f.SE[S_XY] + o, sxy); ldv(a.f.SE[S_XZ] + o, sxz);
            ldv(a.f.SE[S_YX] + o, syx); ldv(a.f.SE[S_YZ] + o, syz);
            ldv(a.f.SE[S_ZX] + o, szx); ldv(a.f.SE[S_ZY] + o, szy);
        }

#pragma unroll
        for (int e = 0; e < V; ++e) {
            const A bzl = (e == 0) ? bz_l : bz[(e + V - 1) % V];
            const A byl = (e == 0) ? by_l : by[(e + V - 1) % V];
            const A dBz_j = dsub(bz[e], bzj[e]);
            const A dBy_k = dsub(by[e], bym[e]);
            const A dBx_k = dsub(bx[e], bxm[e]);
            const A dBz_i = dsub(bz[e], bzl);
            const A dBy_i = dsub(by[e], byl);
            const A dBx_j = dsub(bx[e], bxj[e]);
            if (!PML || (row_main && col_main[e])) {
                // FDTD.cpp:85-93 / kokkos_functors.h:81-89
                const A vx = use_j ? jx[e] : (A)0, vy = use_j ? jy[e] : (A)0, vz = use_j ? jz[e] : (A)0;
                e_x[e] = dadd(e_x[e], curl2j(cy, dBz_j, cz, dBy_k, cj, vx, use_j));
                e_y[e] = dadd(e_y[e], curl2j(cz, dBx_k, cx, dBz_i, cj, vy, use_j));
                e_z[e] = dadd(e_z[e], curl2j(cx, dBy_i, cy, dBx_j, cj, vz, use_j));
            } else {
                // FDTD_PML.cpp:113-130
                syx[e] = round_store<T>(dsub(dmul(syx[e], dcx[e]), dmul(c2x[e], dBz_i)));
                szx[e] = round_store<T>(dadd(dmul(szx[e], dcx[e]), dmul(c2x[e], dBy_i)));
                sxy[e] = round_store<T>(dadd(dmul(sxy[e], dcy), dmul(c2y, dBz_j)));
                szy[e] = round_store<T>(dsub(dmul(szy[e], dcy), dmul(c2y, dBx_j)));
                sxz[e] = round_store<T>(dsub(dmul(sxz[e], dcz), dmul(c2z, dBy_k)));
                syz[e] = round_store<T>(dadd(dmul(syz[e], dcz), dmul(c2z, dBx_k)));
                e_x[e] = dadd(sxz[e], sxy[e]);
                e_y[e] = dadd(syx[e], syz[e]);
                e_z[e] = dadd(szy[e], szx[e]);
            }
        }
        stv(ExO + o, e_x, nvalid);
        stv(EyO + o, e_y, nvalid);
        stv(EzO + o, e_z, nvalid);
        if (PML && any_pml) {
            stv(a.f.SE[S_XY] + o, sxy, nvalid); stv(a.f.SE[S_XZ] + o, sxz, nvalid);
            stv(a.f.SE[S_YX] + o, syx, nvalid); stv(a.f.SE[S_YZ] + o, syz, nvalid);
            stv(a.f.SE[S_ZX] + o, szx, nvalid); stv(a.f.SE[S_ZY] + o, szy, nvalid);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) { bxm[e] = bx[e]; bym[e] = by[e]; }
    }
}

// ---- small utility kernels ------------------------------------------------------------------------

// Sparse host writes / reads (fdtd_scatter / fdtd_gather): the per-step `get_field(JX)[index] = v`
// pattern of perf-tests/sample/sample.cpp:66-81.
template <typename T>
__global__ void scatter_kernel(T* __restrict__ f, Geom g, const long long* __restrict__ idx,
                               const T* __restrict__ vals, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long ij = (long long)g.Ni * g.Nj;
    const long long id = idx[t];
    const int k = (int)(id / ij) - g.k0;
    if (k < 0 || k >= g.nk) return;
    const long long r = id % ij;
    f[(long long)k * g.plane + (r / g.Ni) * g.pitch + (r % g.Ni)] = vals[t];
}

template <typename T>
__global__ void gather_kernel(const T* __restrict__ f, Geom g, const long long* __restrict__ idx,
                              T* __restrict__ vals, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long ij = (long long)g.Ni * g.Nj;
    const long long id = idx[t];
    const int k = (int)(id / ij) - g.k0;
    if (k < 0 || k >= g.nk) return;
    const long long r = id % ij;
    vals[t] = f[(long long)k * g.plane + (r / g.Ni) * g.pitch + (r % g.Ni)];
}

// Dense 2-D slice at a fixed coordinate along `axis` (fdtd_read_slice): the per-step field dump that feeds
// python_script_legend/visualization.py:13-23 (OutFiles_<n>/<iter>.csv).  out is row-major [n1][n0] with
// (n0, n1) = (Ni, Nj) for axis 2 (local plane `index`), (Ni, nk) for axis 1, (Nj, nk) for axis 0.
template <typename T>
__global__ void slice_kernel(const T* __restrict__ f, Geom g, int axis, int index, int n0, int n1, T* __restrict__ out) {
    const long long total = (long long)n0 * n1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t % n0), b = (int)(t / n0);
        long long o;
        if (axis == 2) o = (long long)index * g.plane + (long long)b * g.pitch + a;
        else if (axis == 1) o = (long long)b * g.plane + (long long)index * g.pitch + a;
        else o = (long long)b * g.plane + (long long)a * g.pitch + index;
        out[t] = f[o];
    }
}

// Pending J writes (csrc/fdtd_capi.cu): a dense box [lo, hi) of one J component, [k][j][i].
//   mode 0: box <- array (planes this rank does not own: +0.0)      mode 1: array <- box (owned planes only)
//   mode 2: box[idx[t]] <- vals[t] for the global flat indices that fall inside the box
struct JBoxArgs { int lo[3], hi[3]; };
template <typename T>
__global__ void jbox_kernel(T* __restrict__ f, Geom g, JBoxArgs b, T* __restrict__ box, int mode,
                            const long long* __restrict__ idx, const T* __restrict__ vals, int n) {
    const int ni = b.hi[0] - b.lo[0], nj = b.hi[1] - b.lo[1], nk = b.hi[2] - b.lo[2];
    if (mode == 2) {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n) return;
        const long long ij = (long long)g.Ni * g.Nj, id = idx[t], r = id % ij;
        const int k = (int)(id / ij), j = (int)(r / g.Ni), i = (int)(r % g.Ni);
        if (i < b.lo[0] || i >= b.hi[0] || j < b.lo[1] || j >= b.hi[1] || k < b.lo[2] || k >= b.hi[2]) return;
        box[((long long)(k - b.lo[2]) * nj + (j - b.lo[1])) * ni + (i - b.lo[0])] = vals[t];
        return;
    }
    const long long total = (long long)ni * nj * nk;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int di = (int)(t % ni), dj = (int)((t / ni) % nj), dk = (int)(t / ((long long)ni * nj));
        const int k = b.lo[2] + dk - g.k0;
        const bool owned = k >= 0 && k < g.nk;
        const long long o = (long long)k * g.plane + (long long)(b.lo[1] + dj) * g.pitch + (b.lo[0] + di);
        if (mode == 0) box[t] = owned ? f[o] : (T)0;
        else if (owned) f[o] = box[t];
    }
}

// Device-resident current source (fdtd_set_source): J = ((amp*wx)*wy)*wz on a box, the product order of
// perf-tests/sample/sample.cpp:26-31; `zero` writes +0.0 instead (source expired / zeroed_currents on a box).
struct SourceArgs {
    int lo[3], hi[3];          // global box
    const double* w[3];        // device tables, indexed from lo
    double amp;
    int zero;
};

template <typename T>
__global__ void source_kernel(T* __restrict__ jx, T* __restrict__ jy, T* __restrict__ jz, Geom g, SourceArgs s) {
    const int ni = s.hi[0] - s.lo[0], nj = s.hi[1] - s.lo[1], nk = s.hi[2] - s.lo[2];
    const long long total = (long long)ni * nj * nk;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int di = (int)(t % ni), dj = (int)((t / ni) % nj), dk = (int)(t / ((long long)ni * nj));
        const int k = s.lo[2] + dk - g.k0;
        if (k < 0 || k >= g.nk) continue;
        double v = 0.0;
        if (!s.zero) v = dmul(dmul(dmul(s.amp, s.w[0][di]), s.w[1][dj]), s.w[2][dk]);
        const long long o = (long long)k * g.plane + (long long)(s.lo[1] + dj) * g.pitch + (s.lo[0] + di);
        jx[o] = (T)v; jy[o] = (T)v; jz[o] = (T)v;
    }
}

}  // namespace fdtd_b200
