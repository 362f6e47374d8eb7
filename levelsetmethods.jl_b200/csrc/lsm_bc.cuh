// lsm_bc.cuh — ghost-cell resolution on the device, shared by the strict and the tiled kernels.
//
// Restates meshfield.jl:213-260 (getindex / _getindexbc: resolve dimensions N -> 1, so corner
// ghosts compose) with bc_stencil (boundaryconditions.jl:90-153) inlined.  0-based indices.
// The accumulation uses __*_rn intrinsics, which the compiler never contracts into an FMA, so the
// ghost values are bit-identical in every translation unit regardless of -fmad.
#pragma once
#include "lsm_dev.cuh"

namespace lsm {

__device__ __forceinline__ double bc_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double bc_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float bc_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float bc_add(float a, float b) { return __fadd_rn(a, b); }

// boundaryconditions.jl:90-97 : w_j = prod_{m != j} (-k - m) / (j - m)
__device__ inline double lagrange_w(int j, int k, int P) {
    double w = 1.0;
    for (int m = 0; m <= P; ++m) {
        if (m == j) continue;
        w = __dmul_rn(w, __ddiv_rn(double(-k - m), double(j - m)));
    }
    return w;
}

template <class T> __device__ __forceinline__ T quiet_nan() { return T(__longlong_as_double(0x7FF8000000000000LL)); }

template <int N, class T, int DIM>
__device__ T read_bc(const View<T>& v, int i0, int i1, int i2) {
    if constexpr (DIM == 0) {
        return v.p[(long)i0 + (long)i1 * v.s1 + (long)i2 * v.s2];
    } else {
        constexpr int d = DIM - 1;
        const int i = d == 0 ? i0 : (d == 1 ? i1 : i2);
        const int n = v.n[d];
        if (i >= 0 && i < n) return read_bc<N, T, DIM - 1>(v, i0, i1, i2);
        const BCDev bc = i < 0 ? v.bc[d][0] : v.bc[d][1];
        if (bc.kind == BC_HALO) return read_bc<N, T, DIM - 1>(v, i0, i1, i2);   // stored ghost plane
        auto rd = [&](int j) -> T {
            return read_bc<N, T, DIM - 1>(v, d == 0 ? j : i0, d == 1 ? j : i1, d == 2 ? j : i2);
        };
        T acc = T(0);
        if (bc.kind == BC_PERIODIC) {
            // boundaryconditions.jl:107-119 (1-based: i<1 -> n-(1-i); i>n -> 1+(i-n)): node n duplicates
            // node 1.  The reference re-enters getindex when one wrap is not enough (tiny grids), which
            // is the same as wrapping again.
            int j = i;
            for (int it = 0; it < 64 && (j < 0 || j >= n); ++it) j = j < 0 ? (n - 1) + j : 1 + j - n;
            if (j < 0 || j >= n) return quiet_nan<T>();
            acc = bc_add(acc, bc_mul(T(1.0), rd(j)));
        } else if (bc.kind == BC_EXTRAP) {
            // boundaryconditions.jl:134-144 : nodes b, b+d, .., b+d*P with Lagrange weights
            const int k = i < 0 ? -i : i - (n - 1);
            const int b = i < 0 ? 0 : n - 1;
            const int dd = i < 0 ? 1 : -1;
            for (int j = 0; j <= bc.P; ++j) acc = bc_add(acc, bc_mul(T(lagrange_w(j, k, bc.P)), rd(b + dd * j)));
        } else if (bc.kind == BC_SYMMETRY) {
            // boundaryconditions.jl:146-153 : ghost(b -+ k) = node(b +- k)
            int j = i;
            for (int it = 0; it < 64 && (j < 0 || j >= n); ++it) j = j < 0 ? -j : 2 * (n - 1) - j;
            if (j < 0 || j >= n) return quiet_nan<T>();
            acc = bc_add(acc, bc_mul(T(1.0), rd(j)));
        } else {
            return quiet_nan<T>();   // no BC: the reference throws (meshfield.jl:222-232); the host refuses earlier
        }
        return acc;
    }
}

// phi[I] for a possibly out-of-grid index (meshfield.jl:213-217)
template <int N, class T>
__device__ __noinline__ T getindex_slow(const View<T>& v, int i0, int i1, int i2) {
    return read_bc<N, T, N>(v, i0, i1, i2);
}

}  // namespace lsm
