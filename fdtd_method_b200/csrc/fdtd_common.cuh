// fdtd_common.cuh -- shared device-side types and bit-exact arithmetic helpers.
//
// Storage layout (DESIGN.md "Data layout in HBM"): every field component is one SoA device
// array of (nk + 2*GHOST_PLANES) planes x Nj rows x pitch elements; `pitch` is Ni rounded up to a multiple of
// 128 bytes so that every row starts on a 128-byte line and 16-byte vector accesses never split.
// The pointer held in Fields<T> addresses element (i=0, j=0, local plane 0); planes -2, -1 and nk, nk+1
// are the k ghost planes used by the z-slab halo exchange (one per side for the one-step kernels, two for
// the temporally blocked T2 pass).  Replaces the flat
// std::vector / Kokkos::View of reference include/FDTD/shared.h:15, include/FDTD_kokkos/kokkos_shared.h:16.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fdtd_b200 {

constexpr int GHOST_PLANES = 2;

enum { EX = 0, EY, EZ, BX, BY, BZ, JX, JY, JZ, NCOMP };
// split-field order inside Fields::SE / Fields::SB (reference include/FDTD/FDTD_PML.h:12-13)
enum { S_XY = 0, S_XZ, S_YX, S_YZ, S_ZX, S_ZY, NSPLIT };

template <typename T> struct VecOf;
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };
template <> struct VecOf<float>  { using type = float4;  static constexpr int V = 4; };

struct Geom {
    int Ni, Nj, Nk;      // global extents
    int nk;              // planes owned by this rank
    int k0;              // global index of local plane 0
    int wrap_k;          // 1: periodic index wrap in k inside this rank (single GPU); 0: ghost planes
    long long pitch;     // elements per row
    long long plane;     // elements per plane (pitch * Nj)
};

// src/FDTD/FDTD.cpp:43-53
struct Coefs {
    double cEx, cEy, cEz;   // C*dt/d
    double cBx, cBy, cBz;   // C*dt/(2 d)
    double cJ;              // -4 PI dt
};

// Main box [lo, hi) in GLOBAL coordinates (whole grid when there is no PML) and the 1-D PML tables
// indexed by global coordinate (host-computed with libm in pml_tables.cpp).
struct PmlDesc {
    int lo[3], hi[3];
    const double* decay[3];
    const double* coef2[3];
};

template <typename T>
struct Fields {
    T* E[3];
    T* B[3];
    T* J[3];
    T* SE[NSPLIT];   // Exy Exz Eyx Eyz Ezx Ezy
    T* SB[NSPLIT];   // Bxy Bxz Byx Byz Bzx Bzy
};

// Bounding box (global coords, [lo,hi)) outside of which J is known to be +0.0.
struct JBox {
    int lo[3], hi[3];
    __host__ __device__ bool empty() const { return lo[0] >= hi[0] || lo[1] >= hi[1] || lo[2] >= hi[2]; }
};

// ---- IEEE double arithmetic with the contraction the reference binary does not have -------------
// The reference is built without -march (top-level CMakeLists.txt:20) so it contains no FMA.  Using
// the explicit round-to-nearest intrinsics keeps nvcc from fusing a*b+c whatever -fmad says, which
// is what makes the results bit-identical to the reference (tests/test_parity_gpu.py).
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// float arithmetic (FDTD_FLAG_F32_ARITH): same association, every operation rounded to float, never contracted
__device__ __forceinline__ float dadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float dsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dmul(float a, float b) { return __fmul_rn(a, b); }

// The two curl terms of the main-cell updates.
//   curl2  : c1*d1 - c2*d2                      (B update, FDTD.cpp:121-126)
//   curl2j : (cJ*j + c1*d1) - c2*d2             (E update, FDTD.cpp:85-93; use_j false: the current term is skipped,
//                                                exact because J = +0.0 there)
// double: the reference's operations, one rounding each.  float (FDTD_FLAG_F32_ARITH): explicit fused multiply-adds,
// fma(c1, d1, -(c2*d2)) and fma(-c2, d2, fma(c1, d1, cJ*j)) -- two roundings fewer per component, which keeps the mode
// inside 1e-5 relative L-inf of the fp64 reference (the CPU checker restates exactly this form for the mode).
__device__ __forceinline__ double curl2(double c1, double d1, double c2, double d2) { return dsub(dmul(c1, d1), dmul(c2, d2)); }
__device__ __forceinline__ float curl2(float c1, float d1, float c2, float d2) { return __fmaf_rn(c1, d1, -__fmul_rn(c2, d2)); }
__device__ __forceinline__ double curl2j(double c1, double d1, double c2, double d2, double cj, double j, bool use_j) {
    double t = dmul(c1, d1);
    if (use_j) t = dadd(dmul(cj, j), t);
    return dsub(t, dmul(c2, d2));
}
__device__ __forceinline__ float curl2j(float c1, float d1, float c2, float d2, float cj, float j, bool use_j) {
    const float t = use_j ? __fmaf_rn(c1, d1, __fmul_rn(cj, j)) : __fmul_rn(c1, d1);
    return __fmaf_rn(-c2, d2, t);
}

// One rounding to the storage type (fp32 mode: "float storage, double arithmetic").
template <typename T> __device__ __forceinline__ double round_store(double x);
template <> __device__ __forceinline__ double round_store<double>(double x) { return x; }
template <> __device__ __forceinline__ double round_store<float>(double x) { return (double)__double2float_rn(x); }
template <typename T> __device__ __forceinline__ float round_store(float x) { return x; }   // float arithmetic: already rounded

// ---- 128-bit vector access -----------------------------------------------------------------------
__device__ __forceinline__ void ldv(const double* p, double (&o)[2]) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    o[0] = v.x; o[1] = v.y;
}
__device__ __forceinline__ void ldv(const float* p, double (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = (double)v.x; o[1] = (double)v.y; o[2] = (double)v.z; o[3] = (double)v.w;
}
__device__ __forceinline__ void ldv(const float* p, double (&o)[2]) {   // 2 cells per thread (float PML sweeps)
    const float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = (double)v.x; o[1] = (double)v.y;
}
template <typename A, typename T> __device__ __forceinline__ A lds1(const T* p) { return (A)*p; }
__device__ __forceinline__ void ldv(const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void ldv(const float* p, float (&o)[2]) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = v.x; o[1] = v.y;
}

// Store V values; `nvalid` < V only in the last vector of a row whose Ni is not a multiple of V.
__device__ __forceinline__ void stv(double* p, const double (&v)[2], int nvalid) {
    if (nvalid >= 2) *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    else if (nvalid == 1) p[0] = v[0];
}
__device__ __forceinline__ void stv(float* p, const double (&v)[2], int nvalid) {
    if (nvalid >= 2) *reinterpret_cast<float2*>(p) = make_float2((float)v[0], (float)v[1]);
    else if (nvalid == 1) p[0] = (float)v[0];
}
__device__ __forceinline__ void stv(float* p, const double (&v)[4], int nvalid) {
    if (nvalid >= 4) {
        *reinterpret_cast<float4*>(p) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    } else {
        for (int e = 0; e < nvalid; ++e) p[e] = (float)v[e];
    }
}

__device__ __forceinline__ void stv(float* p, const float (&v)[2], int nvalid) {
    if (nvalid >= 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else if (nvalid == 1) p[0] = v[0];
}
__device__ __forceinline__ void stv(float* p, const float (&v)[4], int nvalid) {
    if (nvalid >= 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        for (int e = 0; e < nvalid; ++e) p[e] = v[e];
    }
}

}  // namespace fdtd_b200
