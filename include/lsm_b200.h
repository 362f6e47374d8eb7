/*
 * lsm_b200.h — C ABI of liblsm_b200.so, a B200 (sm_100a) engine for the dense-grid explicit
 * time-integration path of LevelSetMethods.jl (reference v0.2.0).
 *
 * The reference has no FFI: its seam is Julia dispatch on four internal generics plus the
 * AbstractMeshField interface.  Each entry point below names the reference
 * function(s) it replaces (paths relative to the reference repo root).  Julia binds these with
 * `ccall` (julia/LSMB200.jl, shown in INTEGRATION.md); Python binds them with ctypes
 * (levelsetmethods.jl_b200/_lib.py).
 *
 * Conventions
 *  - plain C types only; every function returns an int32 status (LSM_OK == 0); C++ exceptions
 *    never cross the boundary; lsm_last_error() gives the message of the last failure.
 *  - host arrays are column-major (dim 1 contiguous) exactly like a Julia Array{V,N}; vector
 *    fields are AoS like Array{SVector{N,T},N} (shape (N, n1, .., nN)).  Host pointers are
 *    never retained after a call returns.
 *  - one context == one process == one GPU.  With nranks > 1 every field is slab-partitioned
 *    along its LAST dimension and each rank uploads / downloads only its own slab.
 *  - handles are not thread-safe; distinct contexts are independent.
 *  - there is no CPU fallback: every compute entry point fails with LSM_ERR_CUDA when no
 *    sm_100 device is usable.
 */
#ifndef LSM_B200_H
#define LSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSM_ABI_VERSION 1

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    LSM_OK = 0,
    LSM_ERR_ARG = 1,          /* bad argument (null handle, bad enum, shape mismatch)                     */
    LSM_ERR_CFL = 2,          /* levelsetterms.jl:26  ArgumentError("invalid time-step based on CFL ...") */
    LSM_ERR_TIME = 3,         /* levelsetequation.jl:196  ArgumentError("final time ... must be >= ...")  */
    LSM_ERR_BC = 4,           /* boundaryconditions.jl:184-186 periodic mixed; levelsetequation.jl:69-70  */
    LSM_ERR_CUDA = 5,
    LSM_ERR_NCCL = 6,
    LSM_ERR_OOM = 7,
    LSM_ERR_UNSUPPORTED = 8
};

enum { LSM_F32 = 0, LSM_F64 = 1 };

/* boundaryconditions.jl:27-74.  NeumannBC == EXTRAP with P = 0, LinearExtrapolationBC == P = 1. */
enum { LSM_BC_NONE = -1, LSM_BC_PERIODIC = 0, LSM_BC_EXTRAP = 1, LSM_BC_SYMMETRY = 2 };
#define LSM_MAX_EXTRAP_P 5
typedef struct { int32_t kind; int32_t P; } lsm_bc;

/* levelsetterms.jl:45-49,104-106,139-142,211-213 */
enum { LSM_TERM_ADVECTION = 0, LSM_TERM_NORMAL = 1, LSM_TERM_CURVATURE = 2, LSM_TERM_EIKONAL = 3 };
/* derivatives.jl:11,20 */
enum { LSM_UPWIND = 0, LSM_WENO5 = 1 };
/* how a term's coefficient (velocity / speed / b / S0) is given; replaces _eval_field (levelsetterms.jl:42-43) */
enum {
    LSM_COEF_CONST = 0,       /* cval[0..N-1] (velocity) or cval[0]                                        */
    LSM_COEF_FIELD = 1,       /* device field (ncomp = N for a velocity, 1 otherwise)                      */
    LSM_COEF_SEPARABLE = 2,   /* field made by lsm_field_create_separable: u_d = s_d X_d[i1] Y_d[i2] Z_d[i3] */
    LSM_COEF_NONE = 3         /* EikonalReinitializationTerm() — live sign (levelsetterms.jl:222)           */
};
/* time factor g(t) multiplying the coefficient (device-side stand-in for update_func / f(x,t)) */
enum { LSM_TS_NONE = 0, LSM_TS_COS = 1 /* cos(pi t / tparam) */, LSM_TS_HOST = 2 /* caller passes g per call */ };

/* timestepping.jl:26-28,46-48,65-67 */
enum { LSM_FORWARD_EULER = 0, LSM_RK2 = 1, LSM_RK3 = 2 };

typedef struct lsm_ctx lsm_ctx;
typedef struct lsm_field lsm_field;

#define LSM_MAX_TERMS 4
typedef struct {
    int32_t kind;             /* LSM_TERM_*   */
    int32_t scheme;           /* LSM_UPWIND / LSM_WENO5 (advection only) */
    int32_t coef_kind;        /* LSM_COEF_*   */
    int32_t tscale_kind;      /* LSM_TS_*     */
    double  cval[3];
    double  tparam;
    lsm_field* field;         /* coefficient field for FIELD / SEPARABLE, else NULL */
} lsm_term;

typedef struct {
    int64_t kernel_launches;  /* kernels of this library launched since ctx creation */
    int64_t stage_launches;   /* of which fused stage kernels                         */
    int64_t cfl_passes;       /* CFL reduction passes actually run (cache misses)     */
    int64_t h2d_bytes, d2h_bytes;
    int64_t halo_bytes_sent;  /* NCCL send bytes (multi-rank)                          */
    double  last_stage_ms;    /* CUDA-event time of the last timed stage (see LSM_OPT_TIME_STAGES) */
    double  sum_stage_ms;     /* accumulated over timed stages since lsm_reset_counters */
    int64_t timed_stages;
    int64_t pair_launches;    /* stage launches that took the x-pair kernel (3-D single-term WENO5 advection)  */
    int64_t resident_steps;   /* time steps taken inside the resident cluster kernel (small 2-D grids; one launch per lsm_integrate) */
} lsm_counters;

/* options for lsm_set_option */
enum {
    LSM_OPT_KERNEL = 0,       /* 0 auto (x-pair / tiled where available), 1 force generic strict kernel, 2 force tiled / x-pair, 3 tiled without the x-pair kernel,
                                 4 x-pair kernel with the exact WENO epsilon maximum (bit-identical to 3; default is a 20-bit maximum) */
    LSM_OPT_TIME_STAGES = 1,  /* 1: bracket every stage launch with CUDA events, resolved at the next sync  */
    LSM_OPT_CFL_CACHE = 2,    /* 1 (default): reuse the CFL reduction while coefficient data/scale unchanged  */
    LSM_OPT_OVERLAP = 3,      /* 1 (default): overlap halo exchange with interior compute (multi-rank)        */
    LSM_OPT_FUSE_CFL = 4,     /* 1 (default): lsm_integrate lets the last RK stage reduce the next step's CFL maximum (time-scaled stored velocity) */
    LSM_OPT_GRAPH = 5,        /* 1 (default): lsm_integrate replays a captured CUDA graph of one step's stage launches on small grids (launch-bound regime) */
    LSM_OPT_CFL_CANDIDATES = 6, /* 1 (default): the CFL maximum of a TIME-SCALED static coefficient field is evaluated on the host over the few nodes that
                                  can attain it for any scale (exact, see lsm_api.cu) - no reduction pass, D2H or host sync per step */
    LSM_OPT_RESIDENT = 7      /* 1 (default): lsm_integrate runs the whole time loop of a small 2-D grid (<= 2^15 nodes, one stored-velocity WENO5
                                 advection term without a time factor, index-map BCs) in ONE cluster kernel that keeps the state in distributed
                                 shared memory (lsm_resident2d.cu); bit-identical to the per-stage kernels */
};

/* ---- lifecycle ------------------------------------------------------------------------------ */
int32_t lsm_abi_version(void);
/* Thread-local message of the last failing call (ctx may be NULL). */
const char* lsm_last_error(void);
int32_t lsm_device_count(int32_t* n_out);

/* Single-GPU context on `device`. */
int32_t lsm_ctx_create(int32_t device, lsm_ctx** out);
/* Multi-GPU: rank 0 calls lsm_nccl_unique_id and the host broadcasts the 128 bytes (MPI.jl,
 * torch.distributed, a file ...); then every rank creates its context. */
int32_t lsm_nccl_unique_id(void* id128);
int32_t lsm_ctx_create_rank(int32_t device, int32_t rank, int32_t nranks, const void* id128, lsm_ctx** out);
/* Multi-GPU from ONE process (one Julia task / one Python thread driving a whole box): n_gpus contexts, rank r on
 * device_ids[r], whose NCCL communicators come from ncclCommInitAll.  `out` receives n_gpus handles.  Per-rank calls without
 * an exchange (lsm_field_create / upload / download / set_bc / fill / destroy) are made one context after the other; the
 * collective ones (everything that exchanges halos or all-reduces: compute_cfl, stage, advance, integrate) must run on all
 * ranks concurrently — the lsm_multi_* entry points below do that on one host thread per GPU and return the first error. */
int32_t lsm_ctx_create_multi(int32_t n_gpus, const int32_t* device_ids, lsm_ctx** out);
int32_t lsm_multi_compute_cfl(int32_t n_gpus, lsm_ctx* const* ctx, lsm_field* const* phi, const lsm_term* const* terms, int32_t nterms,
                              double t, const double* gscale, double* dt_out);
int32_t lsm_multi_integrate(int32_t n_gpus, lsm_ctx* const* ctx, int32_t integrator, double cfl, lsm_field* const* phi,
                            const lsm_term* const* terms, int32_t nterms, double t0, double tf, double dt_max, int64_t max_steps,
                            double* t_out, int64_t* steps_out);
int32_t lsm_ctx_destroy(lsm_ctx* ctx);
int32_t lsm_sync(lsm_ctx* ctx);
int32_t lsm_set_option(lsm_ctx* ctx, int32_t option, int32_t value);
int32_t lsm_get_counters(lsm_ctx* ctx, lsm_counters* out);
int32_t lsm_reset_counters(lsm_ctx* ctx);

/* CUDA-event timing on the context's compute stream (the stream every kernel of this library is
 * launched on): record into slot 0..7, then read the elapsed milliseconds between two slots
 * (synchronises on the later event). */
int32_t lsm_event_record(lsm_ctx* ctx, int32_t slot);
int32_t lsm_event_elapsed_ms(lsm_ctx* ctx, int32_t slot_start, int32_t slot_stop, double* ms_out);

/* Pin / unpin a host array (e.g. the memory of a Julia Array) so uploads/downloads run at full
 * PCIe speed.  Optional. */
int32_t lsm_host_register(void* ptr, int64_t bytes);
int32_t lsm_host_unregister(void* ptr);

/* Pure host function (no GPU needed): which planes of the last dimension rank `rank` owns.
 * first_out is 0-based. */
int32_t lsm_slab_plan(int32_t n_last, int32_t nranks, int32_t rank, int32_t* first_out, int32_t* count_out);

/* Pure host function (no GPU needed): the step sizes the time loop takes when the CFL step dt_cfl is CONSTANT (static coefficients), i.e.
 * the loop of _integrate! (timestepping.jl:104-118) replayed with its own arithmetic: while t <= tf - eps(t): dt = min(dt_max, cfl * dt_cfl,
 * tf - t); t += dt — at most max_steps steps (< 0: no limit).  The sequence is run-length encoded into dt_out / count_out (capacity cap;
 * cap = 0 only counts): normally two runs, n full steps and one shorter final step.  This is what lsm_integrate hands to the resident
 * cluster kernel of small 2-D grids (LSM_OPT_RESIDENT); t_out / steps_out are the time and step count that loop reaches. */
int32_t lsm_step_plan(double t0, double tf, double dt_max, double cfl, double dt_cfl, int64_t max_steps, int32_t cap,
                      double* dt_out, int64_t* count_out, int32_t* nruns_out, int64_t* steps_out, double* t_out);

/* ---- grid + field : CartesianGrid (meshes.jl:1-5,34-42) + MeshField (meshfield.jl:51-55) ------ */
/* n = GLOBAL node counts.  ncomp = 1 (scalar) or ndim (velocity).  The field owns its device
 * memory (and, for a state field, the RK stage buffers of _alloc_buffers, timestepping.jl:126,141,168). */
int32_t lsm_field_create(lsm_ctx* ctx, int32_t ndim, const int32_t* n, int32_t dtype, int32_t ncomp,
                         const double* lc, const double* hc, lsm_field** out);
/* Rank-1 separable vector coefficient: component d at node (i1,i2,i3) is
 * scale[d] * tab[d][0][i1] * tab[d][1][i2] * tab[d][2][i3].  `tabs` is the concatenation, for
 * d = 0..ndim-1 and axis a = 0..ndim-1, of n[a] doubles (GLOBAL extents, host memory, copied). */
int32_t lsm_field_create_separable(lsm_ctx* ctx, int32_t ndim, const int32_t* n, const double* lc, const double* hc,
                                   const double* scale, const double* tabs, lsm_field** out);
int32_t lsm_field_destroy(lsm_field* f);
/* _normalize_bc'd boundary conditions: bc[2*d + side], side 0 = left, 1 = right
 * (boundaryconditions.jl:166-188; periodic on one side only -> LSM_ERR_BC). */
int32_t lsm_field_set_bc(lsm_field* f, const lsm_bc* bc);
/* Local (this rank's) extents: n_local[ndim], and the 0-based global index of the first owned
 * plane of the last dimension. */
int32_t lsm_field_local_extent(const lsm_field* f, int32_t* n_local, int32_t* first_last);
/* values(phi) <- host / host <- values(phi): this rank's slab, column-major, AoS for ncomp > 1. Blocking. */
int32_t lsm_field_upload(lsm_field* f, const void* host);
int32_t lsm_field_download(lsm_field* f, void* host);
/* copy!(dst, src) (meshfield.jl:275-278) */
int32_t lsm_field_copy(lsm_field* dst, const lsm_field* src);
/* meshsize(phi) (meshes.jl:109-110) */
int32_t lsm_field_meshsize(const lsm_field* f, double* h_out);
/* phi[I] (meshfield.jl:213-260): BC-aware read evaluated ON THE DEVICE by the same ghost-cell
 * code the stencil kernels use.  I is 1-based like the reference and may lie outside the grid
 * (single-rank contexts only).  count indices of ndim int32 each -> count doubles. */
int32_t lsm_field_getindex(lsm_field* f, const int32_t* I, int32_t count, double* out);
/* Borrowed handle to RK stage buffer `which` (1 or 2) of a state field (the tuple returned by
 * _alloc_buffers, timestepping.jl:126,141,168), so host update_func callbacks can be handed the
 * stage field like the reference does.  Owned by `phi`; do not destroy. */
int32_t lsm_field_stage_buffer(lsm_field* phi, int32_t which, lsm_field** out);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* compute_cfl(terms, phi, t) (levelsetterms.jl:22-38): minimum over terms and nodes, all-reduced
 * over ranks.  gscale: per-term g for LSM_TS_HOST terms (may be NULL).  LSM_ERR_CFL unless dt > 0. */
int32_t lsm_compute_cfl(lsm_ctx* ctx, lsm_field* phi, const lsm_term* terms, int32_t nterms, double t,
                        const double* gscale, double* dt_out);
/* Number of stages of an integrator (1, 2, 3). */
int32_t lsm_nstages(int32_t integrator);
/* One stage (1-based) of _advance! (timestepping.jl:128-137,143-164,170-202), so that a host can
 * run update_term! callbacks between stages like the reference does.  Asynchronous. */
int32_t lsm_stage(lsm_ctx* ctx, int32_t integrator, int32_t stage, lsm_field* phi, const lsm_term* terms,
                  int32_t nterms, double tc, double dt, const double* gscale);
/* _advance!(integrator, phi, buffers, terms, tc, dt): all stages.  Asynchronous. */
int32_t lsm_advance(lsm_ctx* ctx, int32_t integrator, lsm_field* phi, const lsm_term* terms, int32_t nterms,
                    double tc, double dt);
/* integrate!(eq, tf, dt_max) with default hooks: the whole step loop of _integrate!
 * (timestepping.jl:101-122) — dt = min(dt_max, cfl*compute_cfl, tf - tc), loop while
 * tc <= tf - eps(tc), land exactly on tf.  max_steps < 0 = unlimited.  Blocking. */
int32_t lsm_integrate(lsm_ctx* ctx, int32_t integrator, double cfl, lsm_field* phi, const lsm_term* terms,
                      int32_t nterms, double t0, double tf, double dt_max, int64_t max_steps,
                      double* t_out, int64_t* steps_out);

/* EikonalReinitializationTerm(phi0) constructor (levelsetterms.jl:217-221): dst = phi0/sqrt(phi0^2+min(h)^2). */
int32_t lsm_eikonal_s0(lsm_field* dst, const lsm_field* phi0);

/* ---- level-set measures (SURVEY.md §8f "next" row 1): the usual posthook / update_func payload, reduced on the device ---- */
/* volume(phi) (levelsetops.jl:27-33): prod(h) * sum smooth_heaviside(-phi, min h), all-reduced over ranks. */
int32_t lsm_volume(lsm_ctx* ctx, lsm_field* phi, double* out);
/* perimeter(phi) (levelsetops.jl:139-149): prod(h) * sum smooth_delta(phi, min h) * |grad phi| with centred differences; a field
 * without boundary conditions is given LinearExtrapolationBC like the reference does.  Single-rank contexts, or multi-rank
 * fields whose ghost planes are current. */
int32_t lsm_perimeter(lsm_ctx* ctx, lsm_field* phi, double* out);

/* extend_along_normals!(F, phi; nb_iters, cfl, frozen, interface_band, min_norm) (velocityextension.jl:20-116, "next" row 2):
 * nb_iters first-order upwind pseudo-time steps of F_tau + sign(phi) n.grad F = 0 with tau = cfl * min(h).  Runs as ForwardEuler
 * stages of an upwind AdvectionTerm whose velocity is the signed normal S grad(phi)/|grad(phi)|, zeroed on frozen nodes (which
 * reproduces the reference's Dirichlet constraint exactly).  frozen: host uint8 mask of this rank's slab (1 = frozen) or NULL
 * for the band |phi| <= interface_band * min(h).  F and phi must share grid and dtype. */
int32_t lsm_extend_along_normals(lsm_ctx* ctx, lsm_field* F, lsm_field* phi, int32_t nb_iters, double cfl, const uint8_t* frozen,
                                 double interface_band, double min_norm);

/* Set operations on level sets, in place on the device (levelsetops.jl:253-325, "next" row 3):
 *   LSM_CSG_UNION      dst = min(dst, src)   union!      LSM_CSG_INTERSECT  dst = max(dst, src)   intersect!
 *   LSM_CSG_SETDIFF    dst = max(dst, -src)  setdiff!    LSM_CSG_COMPLEMENT dst = -dst (src NULL) complement!
 * min/max follow Julia: NaN if either operand is NaN, min(0.0,-0.0) = -0.0, max(0.0,-0.0) = 0.0. */
enum { LSM_CSG_UNION = 0, LSM_CSG_INTERSECT = 1, LSM_CSG_SETDIFF = 2, LSM_CSG_COMPLEMENT = 3 };
int32_t lsm_field_csg(lsm_ctx* ctx, lsm_field* dst, const lsm_field* src, int32_t op);

/* Analytic initial conditions and coefficients generated ON THE DEVICE ("next" row 3): the reference builds them with
 * MeshField(f, grid) (meshfield.jl:208-211: f evaluated at x = lc + (I-1) h, meshes.jl:115-117) and combines them with the set
 * operations above (docs/src/example-zalesak.md:21-40).  A 1024^3 configuration is 60 GB of fields: generating them in HBM avoids
 * building and uploading them from the host.  Operation order (no FMA), so that a NumPy restatement is bit-identical:
 *   LSM_SHAPE_SPHERE  params = c[ndim], r        sqrt(((x1-c1)^2 + (x2-c2)^2) + (x3-c3)^2) - r
 *   LSM_SHAPE_BOX     params = c[ndim], w[ndim]  max_d (|x_d - c_d| - w_d / 2)           (the notch of the Zalesak disk)
 *   LSM_SHAPE_PLANE   params = n[ndim], offset   ((n1 x1 + n2 x2) + n3 x3) - offset
 *   LSM_SHAPE_CONST   params = v[ncomp]          every node of component c = v[c]        (scalar or vector fields)
 * with x_d = lc_d + i_d * h_d, i_d the 0-based GLOBAL node index.  Values are rounded to the field's dtype at the end. */
enum { LSM_SHAPE_SPHERE = 0, LSM_SHAPE_BOX = 1, LSM_SHAPE_PLANE = 2, LSM_SHAPE_CONST = 3 };
int32_t lsm_field_fill_shape(lsm_field* f, int32_t shape, const double* params, int32_t nparams);
/* Materialise a separable velocity (lsm_field_create_separable) into a stored vector field:
 * u_d[i,j,k] = ((scale_d * X_d[i]) * Y_d[j]) * Z_d[k] — rigid rotation, shear, the Enright field ... from 3 x ndim small tables. */
int32_t lsm_field_fill_separable(lsm_field* dst, const lsm_field* sep);

/* ---- diagnostics (test harness) -------------------------------------------- */
/* max |a - b| over all owned nodes, all-reduced over ranks. */
int32_t lsm_max_abs_diff(lsm_ctx* ctx, const lsm_field* a, const lsm_field* b, double* out);

#ifdef __cplusplus
}
#endif
#endif /* LSM_B200_H */
