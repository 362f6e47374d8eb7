/*
 * lsm_oracle.h — C API of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the dense-grid time-integration
 * path of maltezfaria/LevelSetMethods.jl (v0.2.0).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (liblsm_b200.so)
 * never links, loads or calls anything in this directory.
 *
 * Pinning status: the reference ships NO stored golden arrays and Julia is not installed in
 * this image, so array-level parity with the real Julia package is UNPINNED ("parity
 * unpinned" at the bit level).  What IS pinned (tests/test_oracle_*.py): every exact-value
 * and known-answer check the reference's own test-suite holds for this path
 * (test/test-meshfield.jl, test-derivatives.jl, test-levelsetterms.jl, test-timestepping.jl,
 * test-levelsetequation.jl, test-meshes.jl, test-boundaryconditions.jl) and the two doctest
 * scalars of src/levelsetops.jl:14-25,126-137 (volume / perimeter of a 200x200 circle).
 * In addition tests/test_oracle_numpy_crosscheck.py holds an INDEPENDENT whole-array NumPy restatement of every term, the
 * RK2/RK3 combinations, the ghost composition and the Float32 promotion rules; it agrees with this file to <= 5e-15
 * (bit for bit for Float32, curvature, S0 and the CFL step).
 *
 * Conventions: node indices are 1-based like the reference; arrays are column-major
 * (dim 1 contiguous) exactly like a Julia Array{V,N}; vector-valued fields are AoS
 * (component fastest) like Array{SVector{N,T},N}.
 */
#ifndef LSM_ORACLE_H
#define LSM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_F32 = 0, ORC_F64 = 1 };

/* boundary-condition kinds (src/boundaryconditions.jl:27-74) */
enum { ORC_BC_NONE = -1, ORC_BC_PERIODIC = 0, ORC_BC_EXTRAP = 1, ORC_BC_SYMMETRY = 2,
       /* not in the reference: ghost planes are STORED (slab decomposition test harness) */
       ORC_BC_HALO = 3 };

typedef struct { int32_t kind; int32_t P; } orc_bc;

/* A dense node field.  n[d] = owned nodes per dim; gl/gr = stored ghost planes below/above
 * (non-zero only for ORC_BC_HALO sides).  Storage extent per dim is gl+n+gr. */
typedef struct {
    int32_t ndim;
    int32_t dtype;          /* ORC_F32 / ORC_F64 : the Julia valtype V */
    int32_t n[3];
    int32_t gl[3], gr[3];
    double  lc[3], hc[3];   /* grid corners (CartesianGrid, src/meshes.jl:1-5) */
    /* global node count along each dim and offset of owned node 1 (slabs); for a
     * non-decomposed field nglob == n and off == 0.  Used only for meshsize/getnode. */
    int32_t nglob[3], off[3];
    orc_bc  bc[3][2];       /* (left,right) per dim, as _normalize_bc returns */
    void*   vals;           /* column-major, dtype */
} orc_field;

enum { ORC_TERM_ADVECTION = 0, ORC_TERM_NORMAL = 1, ORC_TERM_CURVATURE = 2, ORC_TERM_EIKONAL = 3 };
enum { ORC_UPWIND = 0, ORC_WENO5 = 1 };
/* coefficient kinds */
enum { ORC_COEF_CONST = 0,      /* cval[0..N-1] (velocity) or cval[0] (speed, b)               */
       ORC_COEF_FIELD = 1,      /* stored field (AoS, ncomp = N for velocity, 1 otherwise)      */
       ORC_COEF_SEPARABLE = 2,  /* u_d(I) = cval[d] * X_d[i1] * Y_d[i2] * Z_d[i3]  (tables)     */
       ORC_COEF_NONE = 3 };     /* eikonal live-sign form (S0 === nothing)                      */
/* time factor g(t) multiplying the coefficient */
enum { ORC_TS_NONE = 0, ORC_TS_COS = 1 /* cos(pi*t/tparam) */, ORC_TS_HOST = 2 /* host passes g */ };

typedef struct {
    int32_t kind, scheme, coef_kind, tscale_kind;
    double  cval[3];
    double  tparam;
    const void* field;        /* coefficient values, same dtype as phi unless field_dtype set */
    int32_t field_dtype;      /* ORC_F32/ORC_F64 */
    int32_t _pad;
    const double* tab[3][3];  /* SEPARABLE: tab[d][axis] -> n[axis] doubles (owned range)    */
} orc_term;

enum { ORC_FE = 0, ORC_RK2 = 1, ORC_RK3 = 2 };

void   orc_set_threads(int nthreads);          /* OpenMP threads for the node loops; 1 = serial like the reference */
int    orc_get_max_threads(void);

/* meshes.jl:109-117 */
double orc_meshsize(const orc_field* f, int dim /*1-based*/);
void   orc_getnode(const orc_field* f, const int32_t* I, double* x);

/* meshfield.jl:213-260 : BC-aware read, I 1-based, may be out of grid */
double orc_getindex(const orc_field* f, const int32_t* I);

/* derivatives.jl:28-175.  op: 0 D0, 1 D+, 2 D-, 3 weno5-, 4 weno5+, 5 D2_0, 6 D2++, 7 D2--,
 * 8 D2 mixed (dim,dim2).  dims 1-based. */
double orc_deriv(const orc_field* f, int op, const int32_t* I, int dim, int dim2);
double orc_weno5(double v1, double v2, double v3, double v4, double v5);
double orc_limiter(double x, double y);

/* levelsetops.jl:197-244, 27-33, 139-183 */
double orc_curvature(const orc_field* f, const int32_t* I);
double orc_volume(const orc_field* f);
double orc_perimeter(const orc_field* f);

/* levelsetterms.jl: per-node Hamiltonian, per-node CFL, global CFL (returns dt, NaN/<=0 flagged by return code) */
double orc_compute_term(const orc_field* phi, const orc_term* term, const int32_t* I, double t, double gscale);
double orc_compute_cfl_term(const orc_field* phi, const orc_term* term, double t, double gscale);
/* returns 0 ok, 1 = would throw ArgumentError (dt !> 0) */
int    orc_compute_cfl(const orc_field* phi, const orc_term* terms, int nterms, double t,
                       const double* gscale /*nullable, per term*/, double* dt_out);
double orc_tscale(const orc_term* term, double t);

/* EikonalReinitializationTerm(phi0) ctor: S0 = v/sqrt(v^2+dx^2) in Float64 (levelsetterms.jl:217-221) */
void   orc_eikonal_s0(const orc_field* phi0, double* s0_out);

/* timestepping.jl:126-202.  One stage of one step.  phi/buf1/buf2 share shape and BCs with
 * `desc` (only desc->vals is ignored).  Stage numbering 1-based.  gscale: per-term host-supplied
 * factor for ORC_TS_HOST terms (else ignored, may be NULL).  The stage input's ghost planes
 * (ORC_BC_HALO) must be current. */
int    orc_stage(const orc_field* desc, int integrator, int stage, void* phi, void* buf1, void* buf2,
                 const orc_term* terms, int nterms, double tc, double dt, const double* gscale);
int    orc_nstages(int integrator);
int    orc_advance(const orc_field* desc, int integrator, void* phi, void* buf1, void* buf2,
                   const orc_term* terms, int nterms, double tc, double dt);
/* timestepping.jl:101-122 + levelsetequation.jl:194-203.  Returns 0 ok, 1 CFL error, 2 tf<t. */
int    orc_integrate(const orc_field* desc, int integrator, double cfl, void* phi,
                     const orc_term* terms, int nterms, double t0, double tf, double dt_max,
                     int64_t max_steps, double* t_out, int64_t* steps_out);

/* velocityextension.jl:20-116 : extend_along_normals!(F, phi; nb_iters, cfl, frozen, interface_band, min_norm).
 * F: array of phi's dtype and owned shape; frozen: nullable uint8 mask (1 = frozen). */
int    orc_extend_along_normals(const orc_field* phi, void* F, int nb_iters, double cfl, const uint8_t* frozen,
                                double interface_band, double min_norm);

/* levelsetops.jl:253-325 : union! (0) / intersect! (1) / setdiff! (2) / complement! (3) on raw value arrays of `dtype`;
 * Julia min/max semantics (NaN propagates, min(0.0,-0.0) = -0.0). */
void   orc_csg(int dtype, void* dst, const void* src, int64_t n, int op);

#ifdef __cplusplus
}
#endif
#endif
