/* ===========================================================================
 * fdtd_b200.h -- C ABI of libfdtd_b200.so, the B200 (sm_100a) implementation of
 * the reference's 3-D FDTD Yee leapfrog time step.
 *
 * This is the drop-in boundary for the path
 *     FDTD::update_fields()      (reference src/FDTD/FDTD.cpp:153-157,
 *                                 src/FDTD_kokkos/FDTD_kokkos.cpp:91-104)
 *     FDTD_PML::update_fields()  (reference src/FDTD/FDTD_PML.cpp:343-365)
 * and the calls either side of it (constructor, get_field, zeroed_currents).
 * The reference has no FFI layer -- its boundary is the C++ class interface
 * (include/FDTD/FDTD.h:35-40, include/FDTD/FDTD_PML.h:41-44).  The C++ classes
 * in include/FDTD_b200/ re-expose exactly that interface on top of the entry
 * points below; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - extern "C", opaque handle, plain pointers and sizes, no C++/torch types.
 *   - every call returns an fdtd_status_t; fdtd_last_error() gives the
 *     thread-local message of the last failing call.
 *   - host buffers use the reference's dense layout: element (i,j,k) at flat
 *     index i + j*Ni + k*Ni*Nj (src/FDTD/FDTD.cpp:66-80), element type = the
 *     solver's dtype (double for FDTD_F64, float for FDTD_F32).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     FDTD_ERR_CUDA when no sm_100 device is usable.
 * ===========================================================================*/
#ifndef FDTD_B200_H_
#define FDTD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDTD_B200_VERSION 100 /* 0.1.0 */

typedef struct fdtd_solver fdtd_solver_t;

/* Same members, order and types as FDTD_struct::Parameters with FP = double
 * (reference include/Structures.h:24-38, include/FP.h:3); a Parameters object
 * may be passed by pointer cast. */
typedef struct fdtd_params {
    int Ni, Nj, Nk;
    double ax, bx, ay, by, az, bz;
    double dx, dy, dz;
} fdtd_params_t;

/* Same enumerators and values as FDTD_enums::Component (reference include/Enums.h:5). */
typedef enum fdtd_component {
    FDTD_EX = 0, FDTD_EY = 1, FDTD_EZ = 2,
    FDTD_BX = 3, FDTD_BY = 4, FDTD_BZ = 5,
    FDTD_JX = 6, FDTD_JY = 7, FDTD_JZ = 8
} fdtd_component_t;

typedef enum fdtd_dtype {
    FDTD_F64 = 0, /* FP = double, the reference's only compiling configuration */
    FDTD_F32 = 1  /* float storage, double arithmetic, one rounding on store */
} fdtd_dtype_t;

typedef enum fdtd_status {
    FDTD_OK = 0,
    FDTD_ERR_INVALID_PARAMETERS = 1, /* std::invalid_argument("ERROR: invalid parameters"), FDTD.cpp:5-7 */
    FDTD_ERR_INVALID_COMPONENT = 2,  /* std::logic_error("ERROR: Invalid field component"), FDTD.cpp:149 */
    FDTD_ERR_CUDA = 3,
    FDTD_ERR_NCCL = 4,
    FDTD_ERR_STATE = 5,
    FDTD_ERR_NOMEM = 6,
    FDTD_ERR_BAD_ARGUMENT = 7
} fdtd_status_t;

/* fdtd_config_t.flags */
#define FDTD_FLAG_J_OPENMP_QUIRK 0x1u /* Jx feeds Ex, Ey AND Ez like FDTD_openmp (FDTD.cpp:85,88,91).
                                         Default is the FDTD_kokkos behaviour (kokkos_functors.h:81-89). */
#define FDTD_FLAG_NO_FUSION 0x2u      /* force the two-sweep kernels (B sweep, E sweep) instead of the fused pass */
#define FDTD_FLAG_NO_GRAPH 0x4u       /* reserved, ignored: fdtd_step(n) issues one pass launch per two steps and is not graph-captured */
#define FDTD_FLAG_NO_OVERLAP 0x8u     /* multi-GPU: issue the halo exchange on the compute stream (no overlap) */
#define FDTD_FLAG_NO_PML_SPLIT 0x10u  /* PML: one launch per sweep with a per-cell predicate instead of interior + shell launches */
#define FDTD_FLAG_NO_TEMPORAL 0x20u   /* fdtd_step(n): never pair steps into the temporally blocked two-step pass */
#define FDTD_FLAG_F32_ARITH 0x80u     /* FDTD_F32 only, opt-in, NOT a reference mode: float storage AND float arithmetic (every operation
                                         rounded to float, same association, no FMA contraction; 4 cells per lane).  The default FDTD_F32 mode
                                         (float storage, double arithmetic) is what a float build of the reference computes;
                                         this one trades bits for speed and is validated against the fp64 reference at north_star's fp32
                                         tolerance (<= 1e-5 relative L-inf) and bit for bit against the oracle's float-arithmetic restatement. */
#define FDTD_FLAG_UNIFORM_SLABS 0x40u /* multi-GPU PML: equal slab heights instead of cost-weighted ones (fdtd_slab_range_cfg) */

typedef enum fdtd_pml_mode {
    FDTD_PML_NONE = 0,     /* class FDTD: periodic everywhere */
    FDTD_PML_PERCENT = 1,  /* class FDTD_PML: thickness_d = int(N_d * pml_percent), FDTD_PML.cpp:254-256 */
    FDTD_PML_THICKNESS = 2 /* extension: explicit per-axis thickness in cells */
} fdtd_pml_mode_t;

typedef struct fdtd_config {
    uint32_t struct_size;   /* = sizeof(fdtd_config_t) */
    fdtd_params_t grid;     /* GLOBAL grid (all ranks pass the same) */
    double dt;
    int32_t dtype;          /* fdtd_dtype_t */
    uint32_t flags;
    int32_t pml_mode;       /* fdtd_pml_mode_t */
    double pml_percent;
    int32_t pml_thickness[3];
    int32_t device;         /* CUDA device ordinal; -1 = current device */
    int32_t rank, nranks;   /* z-slab decomposition, one process per GPU; 0,1 = single GPU */
} fdtd_config_t;

typedef struct fdtd_info {
    int32_t Ni, Nj, Nk;        /* global grid */
    int32_t k_begin, k_end;    /* this rank owns global planes [k_begin, k_end) */
    int32_t dtype, has_pml;
    int32_t pml_thickness[3];
    int64_t pitch;             /* elements per device row (>= Ni, multiple of 128 B) */
    int64_t plane;             /* elements per device plane = pitch * Nj */
    int64_t device_bytes;      /* HBM allocated by this solver */
    int64_t launches;          /* kernels launched by this solver so far */
    int64_t steps_done;
    int32_t fused;             /* 1 if fdtd_step uses the fused E+B pass */
    int32_t rank, nranks, device;
    int32_t temporal;          /* 1 if fdtd_step(n >= 2) pairs steps into the temporally blocked two-step pass */
    int64_t passes_t2;         /* two-step passes run so far (each advances 2 steps) */
    int64_t kernel_ns;         /* reserved */
    int32_t transport;         /* halo transport of a multi-rank solver: 0 none (single GPU / not initialised), 1 NCCL send/recv,
                                  2 copy engines into peer-mapped ghost planes (CUDA IPC or in-process peer access) */
    int32_t halo_in_kernel;    /* 1: the two-step pass waits for the halo inside the kernel (one launch per pass per rank) */
    int32_t f32_arith;         /* 1: FDTD_FLAG_F32_ARITH in effect */
    int32_t reserved0;
} fdtd_info_t;

/* ---- life cycle --------------------------------------------------------- */

/* FDTD(Parameters, double dt)  -- reference include/FDTD/FDTD.h:36, src/FDTD/FDTD.cpp:3-61.
 * fp64, periodic, single GPU (current device), Kokkos J semantics. */
fdtd_status_t fdtd_create(const fdtd_params_t* params, double dt, fdtd_solver_t** out);

/* FDTD_PML(Parameters, FP dt, FP pml_percent) -- include/FDTD/FDTD_PML.h:42, src/FDTD/FDTD_PML.cpp:205-341. */
fdtd_status_t fdtd_create_pml(const fdtd_params_t* params, double dt, double pml_percent, fdtd_solver_t** out);

/* Full-control constructor (dtype, J semantics, explicit PML thickness, device, z-slab rank). */
fdtd_status_t fdtd_create_ex(const fdtd_config_t* cfg, fdtd_solver_t** out);

void fdtd_config_init(fdtd_config_t* cfg); /* zero + defaults (struct_size, device=-1, nranks=1) */

/* ~FDTD() */
fdtd_status_t fdtd_destroy(fdtd_solver_t* s);

/* ---- the hot path ------------------------------------------------------- */

/* FDTD::update_fields() / FDTD_PML::update_fields(): one Yee step (B half, E, B half).
 * Asynchronous on the solver's stream; field reads below synchronise as needed.
 * Caller loops that step one call at a time still reach the two-step pass: an odd call is recorded and returns at
 * once, the next call issues both steps together; every other call of this API that reads or changes solver state
 * runs the recorded step first, so the observable sequence is exactly one step per call (exception, invisible to the caller:
 * fdtd_scatter of Jx / Jy / Jz on a small box keeps the writes pending for the second step of the pair).  A CUDA error of a
 * recorded step is reported by the call that runs it.  (FDTD_B200_NO_LAZY=1 in the environment issues every call at once.) */
fdtd_status_t fdtd_update_fields(fdtd_solver_t* s);

/* nsteps x update_fields().  Bit-identical to nsteps separate calls. */
fdtd_status_t fdtd_step(fdtd_solver_t* s, int nsteps);

/* FDTD::zeroed_currents() -- src/FDTD/FDTD.cpp:132-136. */
fdtd_status_t fdtd_zeroed_currents(fdtd_solver_t* s);

/* ---- field access: what callers do through `Field& get_field(Component)` -- FDTD.cpp:138-151 */

/* Dense copy of this rank's slab (Ni*Nj*(k_end-k_begin) elements of the solver dtype). `count` must
 * equal that element count.  `host` may be pageable or pinned. */
fdtd_status_t fdtd_upload(fdtd_solver_t* s, int component, const void* host, size_t count);
fdtd_status_t fdtd_download(fdtd_solver_t* s, int component, void* host, size_t count);

/* Sparse writes/reads: idx[] are GLOBAL flat indices i + j*Ni + k*Ni*Nj; entries whose plane is not
 * owned by this rank are skipped on write and left untouched on read.  This is the per-step
 * `get_field(JX)[index] = value` pattern of perf-tests/sample/sample.cpp:66-81. */
fdtd_status_t fdtd_scatter(fdtd_solver_t* s, int component, const int64_t* idx, const void* values, size_t n);
fdtd_status_t fdtd_gather(fdtd_solver_t* s, int component, const int64_t* idx, void* values, size_t n);

/* Dense 2-D slice at a fixed GLOBAL coordinate `index` along `axis` (0 = i, 1 = j, 2 = k), extracted on the
 * device: the per-iteration field dump behind python_script_legend/visualization.py:13-23 (one CSV per
 * iteration and component).  Row-major [n1][n0] elements of the solver dtype with
 *   axis 2: (n0, n1) = (Ni, Nj) when this rank owns plane `index`, otherwise nothing is copied;
 *   axis 1: (n0, n1) = (Ni, k_end - k_begin);   axis 0: (n0, n1) = (Nj, k_end - k_begin).
 * *count_out = elements written (0 on a rank that does not own the plane). */
fdtd_status_t fdtd_read_slice(fdtd_solver_t* s, int component, int axis, int index, void* host, size_t capacity,
                              size_t* count_out);

/* Device-resident current source (so the sample scenario never crosses PCIe):
 *   J{x,y,z}(i,j,k) = ((amp[t] * wx[i-lo_i]) * wy[j-lo_j]) * wz[k-lo_k]   on the box [lo, hi)
 * written before step t (t counted from this call), for t < n_amp.  The product order is the
 * reference's (sample.cpp:26-31), so host-computed tables reproduce its values bit for bit.
 * After the last amplitude the solver behaves as if zeroed_currents() had been called
 * (sample.cpp:84).  Replaces the host loop sample.cpp:66-83 / the "SetCurrent" kernel of
 * perf-tests/kokkos_sample/kokkos_sample.cpp:91-108. */
fdtd_status_t fdtd_set_source(fdtd_solver_t* s, const int lo[3], const int hi[3], const double* wx,
                              const double* wy, const double* wz, const double* amp, int n_amp);
fdtd_status_t fdtd_clear_source(fdtd_solver_t* s);

/* Issue a recorded update_fields() call (see fdtd_update_fields) on the solver's stream; never waits for the device.
 * A host thread that drives several slab solvers (fdtd_comm_init_local) calls this -- or fdtd_flush before B accesses --
 * on ALL of them before any call that waits for one of them (uploads, downloads, scatter / gather, sync, destroy):
 * a pass of one slab only completes once its neighbours have issued theirs. */
fdtd_status_t fdtd_issue(fdtd_solver_t* s);
/* Issue any deferred work (a recorded update_fields() call, the trailing B half step) on the solver's stream without
 * waiting for it: afterwards the device arrays hold E(n), B(n) exactly as the reference's do when update_fields()
 * returns.  (bench.py calls this before fdtd_timer_stop so that the closing half step is inside the timed region.) */
fdtd_status_t fdtd_flush(fdtd_solver_t* s);
/* Apply any deferred work and wait for the device (Kokkos::fence() in kokkos_sample.cpp:110-112). */
fdtd_status_t fdtd_sync(fdtd_solver_t* s);

/* Zero-copy access for CUDA-aware callers: device pointer to element (0,0,k_begin) of `component`
 * (row pitch and plane stride in fdtd_info_t).  Flushes deferred work first.
 * VALIDITY: E and B are double-buffered by the fused passes, so the pointer addresses the live generation only until
 * the next fdtd_step / fdtd_update_fields / fdtd_zeroed_currents call on this solver; fetch it again after stepping.
 * The solver assumes the caller may write through it (ghost planes and the J bounding box are invalidated at call
 * time only). */
fdtd_status_t fdtd_device_ptr(fdtd_solver_t* s, int component, void** dptr);

/* ---- introspection / measurement --------------------------------------- */
fdtd_status_t fdtd_get_info(fdtd_solver_t* s, fdtd_info_t* info);
/* CUDA-event stopwatch on the solver's own stream. */
fdtd_status_t fdtd_timer_start(fdtd_solver_t* s);
fdtd_status_t fdtd_timer_stop(fdtd_solver_t* s, double* elapsed_ms);
/* The solver's cudaStream_t (as void*), for callers that want to record their own events. */
fdtd_status_t fdtd_get_stream(fdtd_solver_t* s, void** stream);

/* ---- multi-GPU: z-slab ring (the reference's only distributed design is coarray/fdtd.F90:85-102,149-164) ----
 *
 * COLLECTIVE CALLS.  On a multi-rank solver every call that steps, or that reads or writes a B component
 * (fdtd_step, fdtd_update_fields, fdtd_sync, fdtd_download / fdtd_gather / fdtd_read_slice / fdtd_upload / fdtd_scatter /
 * fdtd_device_ptr of Bx, By, Bz), may run the deferred B half step, which needs the Ex, Ey ring exchange: ALL ranks
 * must make the same sequence of such calls (fdtd_read_slice with axis 2 included -- ranks that do not own the plane
 * take part in the exchange and copy nothing).  J writes (fdtd_upload / fdtd_scatter of Jx..Jz, fdtd_set_source) must
 * also be issued identically on all ranks, with GLOBAL index lists / boxes: each rank tracks the bounding box of the
 * non-zero currents on the host, and the two-step pass re-computes its neighbours' boundary planes from the exchanged
 * J planes under that same box.
 *
 * TRANSPORT.  Default: every rank pushes its boundary planes into the neighbours' ghost planes with the copy engines
 * over NVLink (cudaMemcpyAsync into peer-mapped memory, 32-bit sequence flags for the hand-off; csrc/peer_ring.cu) --
 * no kernel of anybody's runs on an SM, and the two-step pass waits for the halo inside the kernel.  One process per
 * GPU: fdtd_comm_init() (NCCL carries the CUDA IPC handles and stays the fallback; FDTD_B200_TRANSPORT=nccl forces it).
 * One process driving several GPUs: fdtd_comm_init_local(). */
#define FDTD_NCCL_UNIQUE_ID_BYTES 128
/* Rank 0 calls this and ships the bytes to the other ranks (torch.distributed broadcast, a file, ...). */
fdtd_status_t fdtd_nccl_unique_id(void* id_out, size_t capacity);
/* All ranks call this with the same id; uses cfg.rank / cfg.nranks given at creation. */
fdtd_status_t fdtd_comm_init(fdtd_solver_t* s, const void* id, size_t id_bytes);
/* One process, n GPUs: solvers[r] was created with cfg.rank = r, cfg.nranks = n, cfg.device = its GPU.  Links the
 * ring through plain peer access (no NCCL).  The caller then issues every collective call on all n handles from one
 * host thread, in any order (all calls are asynchronous; nothing blocks on a neighbour on the host). */
fdtd_status_t fdtd_comm_init_local(fdtd_solver_t** solvers, int n);
/* Per-pass timeline of the overlapped passes: after enable, the next max_passes passes record four CUDA events each
 * (pass start, halo copies start, halo copies issued/done on this rank's side, pass end); read returns their times in
 * ms relative to the first pass start, 4 doubles per pass. */
fdtd_status_t fdtd_timeline_enable(fdtd_solver_t* s, int max_passes);
fdtd_status_t fdtd_timeline_read(fdtd_solver_t* s, double* ms, int capacity_passes, int* n_passes);
/* Plane range owned by `rank` of `nranks` for a grid with Nk planes (remainder spread over low ranks). */
void fdtd_slab_range(int Nk, int rank, int nranks, int* k_begin, int* k_end);

/* Plane range of `rank` for the solver described by cfg (cfg->nranks ranks).  Equal to fdtd_slab_range() unless the
 * solver has a PML shell in k: then slab heights are cost-weighted (a shell cell moves 36 words per step, a core cell 6),
 * so the ranks that own k-shell planes get fewer planes.  fdtd_get_info() reports the range in use. */
void fdtd_slab_range_cfg(const fdtd_config_t* cfg, int rank, int* k_begin, int* k_end);

/* ---- host-only helpers (no GPU needed) ---------------------------------- */
/* 1-D PML tables for one axis, from src/FDTD/FDTD_PML.cpp:3-65,98-111,316-338:
 * sigma[i], decay[i] = exp(-sigma*dt*C), coef2[i] = sigma ? (1-decay)/(sigma*d) : C*dt/d. */
fdtd_status_t fdtd_pml_profile(int N, int thickness, double d, double dt, double* sigma, double* decay,
                               double* coef2);
/* int(N * pml_percent), FDTD_PML.cpp:254-256 */
int fdtd_pml_thickness(int N, double pml_percent);

/* Test hook (host-only): the plane-chunk list a two-step-pass launch would use for a slab of nk planes producing [lo, hi)
 * (and [lo2, hi2)), see csrc/fdtd_capi.cu::t2_chunk_plan.  capacity >= 72.  Returns the number of chunks. */
int fdtd_debug_t2_chunk_plan(int nk, int lo, int hi, int lo2, int hi2, int wait_in_kernel, int tiles, int gx, int kc_override,
                             int* chunk_lo, int* chunk_hi, int capacity);

const char* fdtd_last_error(void);
int fdtd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FDTD_B200_H_ */
