// lsm_kernels.h — launchers implemented by the .cu translation units, called by lsm_api.cu.
#pragma once
#include "lsm_dev.cuh"

namespace lsm {

struct CflParams {
    TermDev term;
    int n[3];                  // owned nodes (1 for unused dims)
    double h[3];
    unsigned long long* out;   // device scalar, zeroed by the caller
};

// CFL candidate extraction (see lsm_api.cu, cfl_candidates): every node whose unscaled CFL quantity reaches `thr` appends its
// raw coefficient tuple (ncomp doubles) to `out`; `count` keeps counting past `cap` so that the host can tell an overflow.
struct CandParams {
    TermDev term;              // scaled == 0
    int n[3];
    double h[3];
    double thr;
    int cap, ncomp;
    unsigned* count;           // device counter, zeroed by the caller
    double* out;               // cap x ncomp
};

// lsm_generic.cu (strict arithmetic, -fmad=false)
template <class T> cudaError_t launch_stage_generic(int ndim, const StageParams<T>& P, cudaStream_t s);
cudaError_t launch_cfl(int ndim, int dtype_f64, const CflParams& P, int sm_count, cudaStream_t s);
cudaError_t launch_cfl_candidates(int ndim, int dtype_f64, const CandParams& P, int sm_count, cudaStream_t s);
cudaError_t launch_eikonal_s0(int src_f64, int dst_f64, const void* phi, void* out, long n, double dx, cudaStream_t s);
cudaError_t launch_transpose(int f64, bool to_soa, const void* src, void* dst, long n, int ncomp, long cstride, cudaStream_t s);
template <class T> cudaError_t launch_getindex(int ndim, const View<T>& v, const int* d_idx, int count, double* d_out, cudaStream_t s);
template <class T> cudaError_t launch_measure(int ndim, bool perimeter, const View<T>& v, const double* h, double* d_partials, int nblocks, double* d_out, cudaStream_t s);
template <class T> cudaError_t launch_signed_normals(int ndim, const View<T>& v, const double* h, double min_norm, const unsigned char* d_frozen, double band, T* a, long cstride, cudaStream_t s);
// analytic generators (meshfield.jl:208-211 evaluated on the device): see lsm_field_fill_shape / lsm_field_fill_separable
struct ShapeParams {
    int shape, ndim, ncomp;
    int n[3];                  // owned nodes
    int first_last;            // global index of the first owned plane of the last dimension
    double lc[3], h[3];
    double p[8];               // shape parameters
    long cstride;
};
cudaError_t launch_fill_shape(int f64, void* dst, const ShapeParams& P, cudaStream_t s);
cudaError_t launch_fill_separable(int f64, void* dst, long cstride, const int* n, int ndim, const double* scale, const double* const (*tab)[3], cudaStream_t s);
cudaError_t launch_csg(int f64, void* dst, const void* src, long n, int op, cudaStream_t s);
cudaError_t launch_max_abs_diff(int f64, const void* a, const void* b, long n, unsigned long long* out, cudaStream_t s);

// The five Float64 constants of the WENO5 evaluation that do not fit an instruction immediate travel in the kernel parameters
// (constant bank): as literals the compiler re-materialises them with 10 UMOV per node, from the constant bank it takes 3 uniform loads.
struct WenoK { double c133, c56, cm13, e6, fl, pad; };
inline WenoK weno_constants() { return {13.0 / 3.0, 5.0 / 6.0, -1.0 / 3.0, 4.0e-6, 1.0e-70, 0.0}; }

// stored coefficient components and phi^n staged in shared memory next to the phi ring
struct AuxList {
    int n;                 // number of staged scalar tiles per plane
    int first[4];          // first aux index of term k (-1: not staged)
    int p0;                // aux index of phi^n / corr (-1: none)
    const void* src[8];    // box pointers (no ghost planes before the first owned node; same strides as the state)
    WenoK wk;              // see WenoK
};

// lsm_pair3d.cu (3-D single-term WENO5 advection, x-pair threads).  cudaErrorNotSupported -> use the general tiled kernel.
template <class T> cudaError_t launch_stage_pair3d(const StageParams<T>& P, const AuxList& A, cudaStream_t s, bool exact_eps);
// lsm_pair2d.cu (2-D WENO5 advection [+ constant-b curvature], x-pair threads marching along y)
template <class T> cudaError_t launch_stage_pair2d(const StageParams<T>& P, const AuxList& A, cudaStream_t s, bool force, int sm_count);

// lsm_resident2d.cu (small 2-D grids: the whole time loop of one stored-velocity WENO5 advection term in one cluster kernel)
template <class T>
struct ResidentArgs {
    const T* phi;          // state, read once
    T* out;                // state, written once (may alias phi)
    const T* u[2];         // velocity components (same box as the state)
    int n[2];
    long s1;               // row stride (elements)
    int bc[2][2];          // BC_* kinds (index maps only)
    double h[2];
    int nstages;           // 1 ForwardEuler, 2 RK2, 3 TVD-RK3
    int nruns;             // run-length encoded step sizes: count[r] steps of dt[r]
    double dt[4];
    long count[4];
    WenoK wk;
    unsigned long long* status;   // device word, zeroed by the caller: != 0 after the run = internal protocol failure (never expected)
};
template <class T> bool resident2d_supported(int n0, int n1);
template <class T> cudaError_t launch_resident2d(const ResidentArgs<T>& R, cudaStream_t s);

// lsm_tiled.cu (performance kernels).  Returns cudaErrorNotSupported when the configuration is
// not covered, in which case the caller uses the generic kernel.
template <class T> cudaError_t launch_stage_tiled(int ndim, const StageParams<T>& P, int sm_count, cudaStream_t s, int pair_mode = 1, int* used_pair = nullptr);   // pair_mode: 0 never, 1 x-pair kernels (3-D: 20-bit eps max), 2 3-D x-pair kernel with the exact eps max, 3 x-pair kernels forced on small 2-D grids too
template <class T> bool stage_tiled_supported(int ndim, const StageParams<T>& P);

}  // namespace lsm
