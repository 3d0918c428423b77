/* ===========================================================================
 * fdtd_oracle.c -- CPU restatement of the reference's Yee leapfrog time step.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product (libfdtd_b200.so) never links, loads or
 * calls anything in this directory and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 *   (a) SURVEY.md Appendix B values extracted from the real reference,
 *   (b) tests/golden/ fixtures written by oracle/make_golden.py from
 *       oracle/_ref/libfdtd_ref.so (the unmodified reference sources compiled
 *       by oracle/Makefile), and
 *   (c) when oracle/_ref/libfdtd_ref.so is present, the real reference run
 *       side by side on random inputs (bit-for-bit).
 *
 * What it restates (reference = Amazingkivas/FDTD_Method @ df33663):
 *   constructor / coefficients  src/FDTD/FDTD.cpp:3-61
 *   periodic wrap               include/FDTD/FDTD.h:24-31
 *   update_B / update_E         src/FDTD/FDTD.cpp:99-130 / 63-97
 *   Kokkos J semantics          include/FDTD_kokkos/kokkos_functors.h:64-90
 *   update_fields               src/FDTD/FDTD.cpp:153-157
 *   zeroed_currents             src/FDTD/FDTD.cpp:132-136
 *   PML constructor, sigma      src/FDTD/FDTD_PML.cpp:3-65, 205-341
 *   update_{E,B}_PML            src/FDTD/FDTD_PML.cpp:67-203
 *   PML update_fields           src/FDTD/FDTD_PML.cpp:343-365
 *   constants                   include/Constants.h:6-11
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off because the reference binary has no FMA (top-level
 * CMakeLists.txt: -O3 -fopenmp -DNDEBUG, no -march).
 * ===========================================================================*/
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* include/Constants.h:6-11 -- these exact literals are part of the spec. */
#define ORACLE_C 3e10
#define ORACLE_R 1e-12
#define ORACLE_N 4.0
#define ORACLE_PI 3.14159265358

enum { EX, EY, EZ, BX, BY, BZ, JX, JY, JZ, NCOMP };  /* include/Enums.h:5 */
enum { S_EXY, S_EXZ, S_EYX, S_EYZ, S_EZX, S_EZY, S_BXY, S_BXZ, S_BYX, S_BYZ, S_BZX, S_BZY, NSPLIT };
enum { ORACLE_J_KOKKOS = 0, ORACLE_J_OPENMP = 1 };

typedef struct oracle {
    int Ni, Nj, Nk;
    int is_f32, j_mode, has_pml;
    double dx, dy, dz, dt;
    double cE[3], cB[3], cJ;
    int mb[3], me[3];   /* main box [mb, me) per axis */
    int pml[3];         /* PML thickness per axis */
    double *sigma[3], *decay[3], *coef2[3];
    void *f[NCOMP];
    void *s[NSPLIT];
} oracle_t_untyped;

/* Two typed views of the same struct so the body can use REAL* directly. */
#define DECL_TYPED(NAME, REAL_T)                                            \
    typedef struct {                                                        \
        int Ni, Nj, Nk;                                                     \
        int is_f32, j_mode, has_pml;                                        \
        double dx, dy, dz, dt;                                              \
        double cE[3], cB[3], cJ;                                            \
        int mb[3], me[3];                                                   \
        int pml[3];                                                         \
        double *sigma[3], *decay[3], *coef2[3];                             \
        REAL_T *f[NCOMP];                                                   \
        REAL_T *s[NSPLIT];                                                  \
    } NAME;
DECL_TYPED(oracle_f64_t, double)
DECL_TYPED(oracle_f32_t, float)

#define REAL double
#define SUFFIX _f64
#define oracle_t oracle_f64_t
#include "fdtd_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef oracle_t

#define REAL float
#define SUFFIX _f32
#define oracle_t oracle_f32_t
#include "fdtd_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef oracle_t

/* float storage AND float arithmetic (the GPU library's opt-in FDTD_FLAG_F32_ARITH mode; is_f32 == 2) */
#define REAL float
#define ARITH float
#define ARITH_FMA 1
#define SUFFIX _f32a
#define oracle_t oracle_f32_t
#include "fdtd_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef oracle_t

typedef oracle_t_untyped oracle_t;

/* PML profile for one axis.  Follows src/FDTD/FDTD_PML.cpp:254-256 (thickness),
 * :316-321 (SGm), :293-311 (distance functions), :3-61 (sigma = SGm*pow(d/p, N)),
 * :63-65 (decay = exp(-sigma*dt*C)) and :98-111 (coef2 with the sigma==0
 * fallback to the FULL-step coefficient coef_E_d). */
static void pml_profile_axis(int N, int p, double d, double dt, double cE,
                             double *sigma, double *decay, double *coef2) {
    double SGm = 0.0;
    if (p > 0) SGm = -(ORACLE_N + 1.0) / 2.0 * log(ORACLE_R) / ((double)p * d);
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        if (i < p) s = SGm * pow((double)(p - i) / (double)p, ORACLE_N);
        else if (i >= N - p) s = SGm * pow((double)(i + 1 + p - N) / (double)p, ORACLE_N);
        sigma[i] = s;
        decay[i] = exp(-s * dt * ORACLE_C);
        coef2[i] = (s != 0.0) ? (1.0 - decay[i]) / (s * d) : cE;
    }
}

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    for (int c = 0; c < NCOMP; c++) free(o->f[c]);
    for (int c = 0; c < NSPLIT; c++) free(o->s[c]);
    for (int a = 0; a < 3; a++) { free(o->sigma[a]); free(o->decay[a]); free(o->coef2[a]); }
    free(o);
}

/* pml_percent < 0 -> periodic solver (FDTD); >= 0 -> FDTD_PML with that percent.
 * Returns NULL on the reference's invalid_argument condition (FDTD.cpp:5-7). */
oracle_t *oracle_create(int Ni, int Nj, int Nk, double dx, double dy, double dz, double dt,
                        int is_f32, int j_mode, double pml_percent) {
    if (Ni <= 0 || Nj <= 0 || Nk <= 0 || !(dt > 0)) return NULL;
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    if (!o) return NULL;
    o->Ni = Ni; o->Nj = Nj; o->Nk = Nk;
    o->is_f32 = is_f32; o->j_mode = j_mode;
    o->dx = dx; o->dy = dy; o->dz = dz; o->dt = dt;
    const double cdt = ORACLE_C * dt;              /* FDTD.cpp:43 */
    o->cE[0] = cdt / dx; o->cE[1] = cdt / dy; o->cE[2] = cdt / dz;               /* :45-47 */
    o->cB[0] = cdt / (2.0 * dx); o->cB[1] = cdt / (2.0 * dy); o->cB[2] = cdt / (2.0 * dz); /* :49-51 */
    o->cJ = -4.0 * ORACLE_PI * dt;                 /* :53 */
    o->mb[0] = o->mb[1] = o->mb[2] = 0;
    o->me[0] = Ni; o->me[1] = Nj; o->me[2] = Nk;
    const size_t n = (size_t)Ni * (size_t)Nj * (size_t)Nk;
    const size_t esz = is_f32 ? sizeof(float) : sizeof(double);
    for (int c = 0; c < NCOMP; c++) {
        o->f[c] = calloc(n, esz);
        if (!o->f[c]) { oracle_destroy(o); return NULL; }
    }
    if (pml_percent >= 0.0) {
        o->has_pml = 1;
        const int N[3] = {Ni, Nj, Nk};
        const double d[3] = {dx, dy, dz};
        for (int a = 0; a < 3; a++) {
            o->pml[a] = (int)((double)N[a] * pml_percent);   /* FDTD_PML.cpp:254-256 */
            o->mb[a] = o->pml[a];
            o->me[a] = N[a] - o->pml[a];
            o->sigma[a] = (double *)calloc((size_t)N[a], sizeof(double));
            o->decay[a] = (double *)calloc((size_t)N[a], sizeof(double));
            o->coef2[a] = (double *)calloc((size_t)N[a], sizeof(double));
            pml_profile_axis(N[a], o->pml[a], d[a], dt, o->cE[a], o->sigma[a], o->decay[a], o->coef2[a]);
        }
        for (int c = 0; c < NSPLIT; c++) {
            o->s[c] = calloc(n, esz);
            if (!o->s[c]) { oracle_destroy(o); return NULL; }
        }
    }
    return o;
}

void *oracle_field(oracle_t *o, int comp) { return (comp >= 0 && comp < NCOMP) ? o->f[comp] : NULL; }
void *oracle_split(oracle_t *o, int which) { return (o->has_pml && which >= 0 && which < NSPLIT) ? o->s[which] : NULL; }
const double *oracle_pml_sigma(oracle_t *o, int axis) { return o->has_pml ? o->sigma[axis] : NULL; }
const double *oracle_pml_decay(oracle_t *o, int axis) { return o->has_pml ? o->decay[axis] : NULL; }
const double *oracle_pml_coef2(oracle_t *o, int axis) { return o->has_pml ? o->coef2[axis] : NULL; }
int oracle_pml_size(oracle_t *o, int axis) { return o->pml[axis]; }
double oracle_coef(oracle_t *o, int which) {   /* 0-2 cE, 3-5 cB, 6 cJ */
    if (which < 3) return o->cE[which];
    if (which < 6) return o->cB[which - 3];
    return o->cJ;
}

void oracle_update_B(oracle_t *o) {
    if (o->is_f32 == 2) update_B_f32a((oracle_f32_t *)o); else if (o->is_f32) update_B_f32((oracle_f32_t *)o); else update_B_f64((oracle_f64_t *)o);
}
void oracle_update_E(oracle_t *o) {
    if (o->is_f32 == 2) update_E_f32a((oracle_f32_t *)o); else if (o->is_f32) update_E_f32((oracle_f32_t *)o); else update_E_f64((oracle_f64_t *)o);
}
void oracle_update_fields(oracle_t *o) {
    if (o->is_f32 == 2) update_fields_f32a((oracle_f32_t *)o); else if (o->is_f32) update_fields_f32((oracle_f32_t *)o); else update_fields_f64((oracle_f64_t *)o);
}
void oracle_step(oracle_t *o, int nsteps) {
    for (int s = 0; s < nsteps; s++) oracle_update_fields(o);
}
/* src/FDTD/FDTD.cpp:132-136 */
void oracle_zeroed_currents(oracle_t *o) {
    const size_t n = (size_t)o->Ni * (size_t)o->Nj * (size_t)o->Nk;
    const size_t esz = o->is_f32 ? sizeof(float) : sizeof(double);
    memset(o->f[JX], 0, n * esz);
    memset(o->f[JY], 0, n * esz);
    memset(o->f[JZ], 0, n * esz);
}
