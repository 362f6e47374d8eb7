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

// lsm_generic.cu (strict arithmetic, -fmad=false)
template <class T> cudaError_t launch_stage_generic(int ndim, const StageParams<T>& P, cudaStream_t s);
cudaError_t launch_cfl(int ndim, int dtype_f64, const CflParams& P, int sm_count, cudaStream_t s);
cudaError_t launch_eikonal_s0(int src_f64, int dst_f64, const void* phi, void* out, long n, double dx, cudaStream_t s);
cudaError_t launch_transpose(int f64, bool to_soa, const void* src, void* dst, long n, int ncomp, long cstride, cudaStream_t s);
template <class T> cudaError_t launch_getindex(int ndim, const View<T>& v, const int* d_idx, int count, double* d_out, cudaStream_t s);
template <class T> cudaError_t launch_measure(int ndim, bool perimeter, const View<T>& v, const double* h, double* d_partials, int nblocks, double* d_out, cudaStream_t s);
template <class T> cudaError_t launch_signed_normals(int ndim, const View<T>& v, const double* h, double min_norm, const unsigned char* d_frozen, double band, T* a, long cstride, cudaStream_t s);
cudaError_t launch_csg(int f64, void* dst, const void* src, long n, int op, cudaStream_t s);
cudaError_t launch_max_abs_diff(int f64, const void* a, const void* b, long n, unsigned long long* out, cudaStream_t s);

// lsm_tiled.cu (performance kernels).  Returns cudaErrorNotSupported when the configuration is
// not covered, in which case the caller uses the generic kernel.
template <class T> cudaError_t launch_stage_tiled(int ndim, const StageParams<T>& P, int sm_count, cudaStream_t s);
template <class T> bool stage_tiled_supported(int ndim, const StageParams<T>& P);

}  // namespace lsm
