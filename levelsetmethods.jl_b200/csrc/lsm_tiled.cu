// lsm_tiled.cu — performance kernels: fused RK-stage stencil kernels for 2-D and 3-D grids.
//
// K1: one kernel per RK stage = ghost resolution + the finite differences of
// every term + each term's Hamiltonian + the stage combination, computed from shared memory.
//
//   * thread block = TX x TY threads owning an x-y tile of TX x (TY*NY) nodes.  In 3-D it marches along
//     z over a chunk of planes on a ring of 2*3+2 planes; in 2-D the ring is a single plane.  Every plane
//     (tile + 3-cell halo, corners included, so curvature's mixed differences are covered too) is
//     brought into the ring ONCE with cp.async (LDGSTS, no register staging) one iteration ahead of its
//     first use.  Stored coefficient fields (velocity, speed, b, S0) and phi^n are staged the same way
//     in double-buffered "aux" tiles, so the compute phase reads shared memory only.
//   * ghost cells: periodic / Neumann / symmetry / stored-halo boundaries are index maps with weight 1
//     (boundaryconditions.jl:107-153), so boundary tiles fill their ghosts by copying from the REMAPPED
//     address — no arithmetic, no divergence in the compute phase.  ExtrapolationBC{P>=1} (a weighted
//     stencil) takes the REMAP=false instantiation, whose boundary tiles call lsm_bc.cuh.
//   * the B200 FP64 pipe (measured 63 lane-ops/clk/SM, tools/fp64_peak.cu) is a co-bound of these
//     stencils, so every Hamiltonian is restructured to minimise issue slots while staying within
//     1e-10 of the reference:
//       - WENO5: samples are read in UPWIND ORDER q_k = phi[i - s*(3-k)], s = sign(u), which turns
//         u * (u > 0 ? weno5- : weno5+) into (|u|/h) * W(q) with no selects; W works on undivided
//         differences (WENO5 is homogeneous of degree 1), smoothness indicators and candidates are
//         written on second differences, the three weight divisions + three normalisations become
//         ONE reciprocal (MUFU.RCP64H + 2 Newton steps), and max(v^2) runs on the integer pipe:
//         ~47 DP instructions per evaluation instead of ~65 + 6 divisions (13.6 slots each);
//       - Godunov/ENO2 terms use undivided differences and the positive homogeneity of minmod;
//       - curvature uses kappa*|grad phi| = (tr(H) q - g'Hg)/q, i.e. no pow() and no sqrt().
//
// Compiled with FMA contraction ON.  Parity with the CPU oracle is checked in tests/ (<= 1e-10 after
// 100 RK3 steps in Float64, <= 1e-4 in Float32).
#include "lsm_dev.cuh"
#include "lsm_bc.cuh"
#include "lsm_kernels.h"

namespace lsm {

namespace {

constexpr int HAL = 3;           // WENO5 reach; every term's stencil fits in it

enum : int { M_ADV_WENO = 1, M_ADV_UPWIND = 2, M_NORMAL = 4, M_CURV = 8, M_EIK = 16, M_ALL = 31 };

__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc, int bytes8) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (bytes8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 1/x for a positive normal x: MUFU.RCP64H seed + two Newton steps (relative error ~1e-16)
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = fma(-x, y, 1.0);
    y = fma(y, t, y);
    t = fma(-x, y, 1.0);
    y = fma(y, t, y);
    return y;
}

// max(|a|..|e|) exactly, on the integer pipe (non-negative doubles order like unsigned integers)
__device__ __forceinline__ double absmax5(double a, double b, double c, double d, double e) {
    auto key = [](double x) -> unsigned long long {
        return ((unsigned long long)((unsigned)__double2hiint(x) & 0x7fffffffu) << 32) | (unsigned)__double2loint(x);
    };
    unsigned long long m = key(a), k;
    k = key(b); m = k > m ? k : m;
    k = key(c); m = k > m ? k : m;
    k = key(d); m = k > m ? k : m;
    k = key(e); m = k > m ? k : m;
    return __longlong_as_double((long long)m);
}

// Undivided upwind WENO5: given six samples in upwind order (q3 is the node, q0 the far upwind
// end), returns h * weno5 of the reference (derivatives.jl:61-121), i.e. the reference value is
// this divided by h and multiplied by s = sign of the sampling direction.
//   d_k  = q_{k+1} - q_k            (v_k * h of the reference, up to the global sign s)
//   e_k  = d_k - d_{k-1}
//   4*S1 = 13/3 (e2-e1)^2 + (3e2-e1)^2 ; 4*S2 = 13/3 (e3-e2)^2 + (e2+e3)^2 ; 4*S3 = 13/3 (e4-e3)^2 + (e4-3e3)^2
//   b_k  = 4 h^2 (S_k + eps) = 4*S_k(d) + 4e-6 max(d^2) + floor
//   w_k  ~ {1,6,3} / b_k^2  ->  result = d2 + (q1 G1 + q2 G2 + q3 G3) / (q1 + 6 q2 + 3 q3),  q1 = (b2 b3)^2 ...
//   G1 = 5/6 e2 - 1/3 e1 ; G2 = 2 e3 + e2 ; G3 = 2 e3 - 1/2 e4     (6 and 3 folded in)
// The floor 1e-70 replaces 4e-99*h^2 (which would underflow in the product form).  It only matters
// where every |d| < ~1e-26, i.e. on numerically flat data, where the result is O(|d|) either way.
template <class T>
__device__ __forceinline__ double weno5_up(T q0, T q1, T q2, T q3, T q4, T q5) {
    const double d0 = double(T(q1 - q0)), d1 = double(T(q2 - q1)), d2 = double(T(q3 - q2)),
                 d3 = double(T(q4 - q3)), d4 = double(T(q5 - q4));
    const double e1 = d1 - d0, e2 = d2 - d1, e3 = d3 - d2, e4 = d4 - d3;
    const double m = absmax5(d0, d1, d2, d3, d4);
    const double eps = fma(4.0e-6, m * m, 1.0e-70);
    const double c133 = 13.0 / 3.0;
    const double t1a = e2 - e1, t1b = e3 - e2, t1c = e4 - e3;
    const double t2a = fma(3.0, e2, -e1), t2b = e2 + e3, t2c = fma(-3.0, e3, e4);
    const double b1 = fma(t2a, t2a, fma(c133, t1a * t1a, eps));
    const double b2 = fma(t2b, t2b, fma(c133, t1b * t1b, eps));
    const double b3 = fma(t2c, t2c, fma(c133, t1c * t1c, eps));
    const double p12 = b1 * b2, p13 = b1 * b3, p23 = b2 * b3;
    const double w1 = p23 * p23, w2 = p13 * p13, w3 = p12 * p12;
    const double den = fma(3.0, w3, fma(6.0, w2, w1));
    const double G1 = fma(5.0 / 6.0, e2, (-1.0 / 3.0) * e1);
    const double G2 = fma(2.0, e3, e2);
    const double G3 = fma(2.0, e3, -0.5 * e4);
    const double num = fma(w3, G3, fma(w2, G2, w1 * G1));
    return fma(num, fast_rcp(den), d2);
}

// levelsetterms.jl:184-187
__device__ __forceinline__ double minmod(double x, double y) {
    if (!(x * y > 0.0)) return 0.0;
    return fabs(x) <= fabs(y) ? x : y;
}

template <class T, int NDIM, int TX, int TY, int NY>
struct TileGeom {
    static constexpr int W = TX + 2 * HAL;
    static constexpr int HH = TY * NY + 2 * HAL;
    static constexpr int PLANE = W * HH;          // phi plane: tile + halo
    static constexpr int TILE = TX * TY * NY;     // owned nodes of one plane of the tile
    static constexpr int NT = TX * TY;
    static constexpr int NW = NT / 32;
    static constexpr int RING = NDIM == 3 ? 2 * HAL + 2 : 1;
    static constexpr int NBUF = NDIM == 3 ? 2 : 1;        // aux double buffering along z
    static size_t smem_bytes(int naux) { return ((size_t)RING * PLANE + (size_t)NBUF * naux * TILE) * sizeof(T); }
};

// stored coefficient components and phi^n staged in shared memory next to the phi ring
struct AuxList {
    int n;                 // number of staged scalar tiles per plane
    int first[4];          // first aux index of term k (-1: not staged)
    int p0;                // aux index of phi^n / corr (-1: none)
    const void* src[8];    // box pointers (no ghost planes before the first owned node; same strides as the state)
};

// Ghost index -> stored index for the boundary conditions that are pure index maps with weight 1
// (boundaryconditions.jl:107-119 periodic wrap, :134-144 with P = 0 i.e. NeumannBC, :146-153 symmetry).
// Applied independently per dimension this equals the reference's dimension-by-dimension recursion
// (meshfield.jl:248-260).  BC_HALO sides keep the index (stored ghost plane).  Needs n >= 4.
__device__ __forceinline__ int remap_index(int i, int n, int kind_lo, int kind_hi) {
    if (i < 0) {
        if (kind_lo == BC_PERIODIC) return i + n - 1;
        if (kind_lo == BC_EXTRAP) return 0;
        if (kind_lo == BC_SYMMETRY) return -i;
        return i;
    }
    if (i >= n) {
        if (kind_hi == BC_PERIODIC) return i - n + 1;
        if (kind_hi == BC_EXTRAP) return n - 1;
        if (kind_hi == BC_SYMMETRY) return 2 * (n - 1) - i;
        return i;
    }
    return i;
}

// Fused stage kernel.  MASK = which term kinds the instantiation carries code for; the terms themselves
// (order, coefficients) are runtime data, applied one after the other like the reference
// (x = base; x -= c*H_1; x -= c*H_2; ..., timestepping.jl:128-202).
template <class T, int NDIM, int MASK, int NTS, int COEFK, bool REMAP, int TX, int TY, int NY, int MINB>
__global__ void __launch_bounds__(TX * TY, MINB)
stage_tiled_kernel(const __grid_constant__ StageParams<T> P, const __grid_constant__ AuxList A, const int cz) {
    using G = TileGeom<T, NDIM, TX, TY, NY>;
    constexpr int RING = G::RING;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const ring = reinterpret_cast<T*>(smem_raw);
    T* const aux = ring + (size_t)RING * G::PLANE;                   // [NBUF][naux][TILE]

    const int n0 = P.in.n[0], n1 = P.in.n[1], n2 = NDIM == 3 ? P.in.n[2] : 1;
    const long vs1 = P.in.s1, vs2 = NDIM == 3 ? P.in.s2 : 0;
    const T* __restrict__ const vp = P.in.p;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int lane = tx & 31, warp = (ty * TX + tx) >> 5;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * (TY * NY);
    // range of the LAST dimension to update: z planes in 3-D, rows in 2-D (slab decomposition splits it)
    const int zbeg = NDIM == 3 ? P.r0 + blockIdx.z * cz : 0;
    const int zend = NDIM == 3 ? min(P.r1, zbeg + cz) : 1;
    if (zbeg >= zend) return;
    const int ylo = NDIM == 2 ? P.r0 : 0, yhi = NDIM == 2 ? P.r1 : n1;
    if (NDIM == 2 && (y0 >= yhi || y0 + TY * NY <= ylo)) return;

    // whole tile + halo inside the stored x-y extent -> plain copies; otherwise resolve ghosts
    const bool y_lo_ok = (y0 - HAL >= 0) || (NDIM == 2 && P.in.bc[1][0].kind == BC_HALO);
    const bool y_hi_ok = (y0 + TY * NY + HAL <= n1) || (NDIM == 2 && P.in.bc[1][1].kind == BC_HALO);
    const bool xy_in = (x0 - HAL >= 0) && (x0 + TX + HAL <= n0) && y_lo_ok && y_hi_ok;
    const int kzl = P.in.bc[2][0].kind, kzh = P.in.bc[2][1].kind;

    // one warp per row of the (tile + halo) plane: lanes 0..31 copy columns 0..31, lanes 0..W-33 also 32..W-1
    auto load_phi = [&](int z) {
        T* dst = ring + ((z + 1024) & (RING - 1)) * G::PLANE;
        const bool z_plain = NDIM == 2 || ((z >= 0 || kzl == BC_HALO) && (z < n2 || kzh == BC_HALO));
        if (xy_in && z_plain) {
            const T* src = vp + (long)(x0 - HAL) + (long)(y0 - HAL) * vs1 + (long)z * vs2;
            for (int r = warp; r < G::HH; r += G::NW) {
                const T* s = src + (long)r * vs1;
                T* d = dst + r * G::W;
                cp_async(d + lane, s + lane, sizeof(T) == 8);
                if (lane < G::W - 32) cp_async(d + 32 + lane, s + 32 + lane, sizeof(T) == 8);
            }
        } else if (REMAP) {
            const int zz = NDIM == 3 ? remap_index(z, n2, kzl, kzh) : 0;
            // columns / rows beyond the grid + halo of a partial tile are never read: clamp them into range
            const int gxa = min(max(remap_index(x0 - HAL + lane, n0, P.in.bc[0][0].kind, P.in.bc[0][1].kind), 0), n0 - 1);
            const int gxb = min(max(remap_index(x0 - HAL + 32 + lane, n0, P.in.bc[0][0].kind, P.in.bc[0][1].kind), 0), n0 - 1);
            for (int r = warp; r < G::HH; r += G::NW) {
                int gy = remap_index(y0 - HAL + r, n1, P.in.bc[1][0].kind, P.in.bc[1][1].kind);
                if (NDIM == 2) gy = min(max(gy, P.in.bc[1][0].kind == BC_HALO ? -HAL : 0), P.in.bc[1][1].kind == BC_HALO ? n1 - 1 + HAL : n1 - 1);
                else gy = min(max(gy, 0), n1 - 1);
                const T* s = vp + (long)gy * vs1 + (long)zz * vs2;
                T* d = dst + r * G::W;
                cp_async(d + lane, s + gxa, sizeof(T) == 8);
                if (lane < G::W - 32) cp_async(d + 32 + lane, s + gxb, sizeof(T) == 8);
            }
        } else {
            for (int r = warp; r < G::HH; r += G::NW) {
                T* d = dst + r * G::W;
                d[lane] = getindex_slow<NDIM, T>(P.in, x0 - HAL + lane, y0 - HAL + r, z);
                if (lane < G::W - 32) d[32 + lane] = getindex_slow<NDIM, T>(P.in, x0 - HAL + 32 + lane, y0 - HAL + r, z);
            }
        }
    };
    // stored coefficients and phi^n of the owned nodes of plane z
    auto load_aux = [&](int z) {
        T* dst = aux + (size_t)(z & (G::NBUF - 1)) * A.n * G::TILE;
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int r = ty + k * TY;
            const int ii = x0 + tx, jj = y0 + r;
            if (ii < n0 && jj < n1) {
                const long node = (long)ii + (long)jj * vs1 + (long)z * vs2;
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (a < A.n) cp_async(dst + a * G::TILE + r * TX + tx, static_cast<const T*>(A.src[a]) + node, sizeof(T) == 8);
            }
        }
    };

    if (NDIM == 3) for (int z = zbeg - HAL; z <= zbeg + HAL; ++z) load_phi(z);
    else load_phi(0);
    load_aux(zbeg);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    const double ih[3] = {1.0 / P.h[0], 1.0 / P.h[1], NDIM == 3 ? 1.0 / P.h[2] : 0.0};
    const int i = x0 + tx;

    for (int z = zbeg; z < zend; ++z) {
        if (NDIM == 3) {
            if (z + HAL + 1 <= zend - 1 + HAL) load_phi(z + HAL + 1);
            if (z + 1 < zend) load_aux(z + 1);
            cp_async_commit();
        }
        const T* cur = ring + ((z + 1024) & (RING - 1)) * G::PLANE;
        const T* auxz = aux + (size_t)(z & (G::NBUF - 1)) * A.n * G::TILE;
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int r = ty + k * TY;
            const int j = y0 + r;
            if (i < n0 && j < n1 && (NDIM == 3 || (j >= ylo && j < yhi))) {
                const int sc = (r + HAL) * G::W + tx + HAL;
                const int st = r * TX + tx;
                const T* c0 = cur + sc;
                const T qc = c0[0];
                // sample at offset m along dimension d (all inside the tile + halo)
                auto at = [&](int d, int m) -> T {
                    if (d == 0) return c0[m];
                    if (d == 1) return c0[m * G::W];
                    return ring[((z + m + 1024) & (RING - 1)) * G::PLANE + sc];
                };
                auto at2 = [&](int d1, int m1, int d2, int m2) -> T {      // d1 < d2
                    const int off = (d1 == 0 ? m1 : m1 * G::W) + (d2 == 1 ? m2 * G::W : 0);
                    if (d2 == 2) return ring[((z + m2 + 1024) & (RING - 1)) * G::PLANE + sc + off];
                    return c0[off];
                };
                // coefficient component d of term k (times g(t))
                auto coef = [&](const TermDev& tm, int kk, int d) -> double {
                    double v;
                    const int ck = COEFK >= 0 ? COEFK : tm.coef_kind;      // compile-time for the single-term advection kernels
                    if (ck == COEF_FIELD) {
                        if (COEFK >= 0 || A.first[kk] >= 0) v = double(auxz[((COEFK >= 0 ? 0 : A.first[kk]) + d) * G::TILE + st]);
                        else {   // Float64 coefficient with a Float32 state (S0): read directly
                            const long node = (long)i + (long)j * vs1 + (long)z * vs2;
                            v = static_cast<const double*>(tm.coef)[(long)d * tm.cstride + node];
                        }
                    } else if (ck == COEF_SEPARABLE) {
                        v = (tm.cval[d] * __ldg(tm.tab[d][0] + i)) * __ldg(tm.tab[d][1] + j);
                        if (NDIM == 3) v = v * __ldg(tm.tab[d][2] + z);
                    } else v = tm.cval[d];
                    if (tm.scaled) v = v * tm.g;
                    return v;
                };
                // second-order ENO pair along d, undivided: returns h*neg, h*pos (levelsetterms.jl:156-170, 252-265)
                auto eno2 = [&](int d, double& ng, double& ps) {
                    const T pm2 = at(d, -2), pm1 = at(d, -1), pp1 = at(d, 1), pp2 = at(d, 2);
                    const double dm = double(T(qc - pm1)), dp = double(T(pp1 - qc));
                    const double cc = double(T(pp1 - T(2) * qc + pm1));
                    const double cm = double(T(pm2 - T(2) * pm1 + qc)), cp = double(T(qc - T(2) * pp1 + pp2));
                    ng = fma(0.5, minmod(cm, cc), dm);
                    ps = fma(-0.5, minmod(cp, cc), dp);
                };

                T x = qc;
                if (A.p0 >= 0) {
                    const T pn = auxz[A.p0 * G::TILE + st];
                    if (P.base == BASE_RK3_S2) x = T(fma(0.75, double(pn), 0.25 * double(qc)));       // timestepping.jl:183
                    else if (P.base == BASE_RK3_S3) x = T((pn + T(2) * qc) / T(3));                   // timestepping.jl:194
                    else x = pn;                                                                       // RK2 S2 (corr)
                }
                T x2 = qc;

                constexpr bool ONE = (MASK & (MASK - 1)) == 0;      // single kind: no runtime kind tests
                auto one_term = [&](const TermDev& tm, const int kk) {
                    double H = 0.0;
                    if ((MASK & M_ADV_WENO) && (ONE || (tm.kind == TERM_ADVECTION && tm.scheme == SCHEME_WENO5))) {
                        // levelsetterms.jl:73-82 : H = sum_d u_d * weno(d) = sum_d (|u_d| / h_d) * W_d, left to right
#pragma unroll
                        for (int d = 0; d < NDIM; ++d) {
                            const double u = coef(tm, kk, d);
                            const int s = u > 0 ? 1 : -1;       // upwind-ordered sampling: q_k = phi[i - s*(3-k)]
                            const double w = weno5_up<T>(at(d, -3 * s), at(d, -2 * s), at(d, -s), qc, at(d, s), at(d, 2 * s));
                            const double a = fabs(u) * ih[d];
                            H = d == 0 ? a * w : fma(a, w, H);
                        }
                    } else if ((MASK & M_ADV_UPWIND) && (ONE || tm.kind == TERM_ADVECTION)) {
#pragma unroll
                        for (int d = 0; d < NDIM; ++d) {
                            const double u = coef(tm, kk, d);
                            const double der = u > 0 ? double(T(qc - at(d, -1))) : double(T(at(d, 1) - qc));
                            const double a = u * ih[d];
                            H = d == 0 ? a * der : fma(a, der, H);
                        }
                    } else if ((MASK & (M_NORMAL | M_EIK)) && (ONE || tm.kind == TERM_NORMAL || tm.kind == TERM_EIKONAL)) {
                        // Godunov |grad phi| from the ENO2 pair, both upwind selections (levelsetterms.jl:156-170, 252-265)
                        double gp = 0.0, gm = 0.0;
#pragma unroll
                        for (int d = 0; d < NDIM; ++d) {
                            double ng, ps;
                            eno2(d, ng, ps);
                            const double i2 = ih[d] * ih[d];
                            const double a = fmax(ng, 0.0), b = fmin(ps, 0.0), c = fmin(ng, 0.0), e = fmax(ps, 0.0);
                            gp = fma(fma(a, a, b * b), i2, gp);
                            gm = fma(fma(c, c, e * e), i2, gm);
                        }
                        if (tm.kind == TERM_NORMAL) {
                            const double v = coef(tm, kk, 0);
                            H = fmax(v, 0.0) * sqrt(gp) + fmin(v, 0.0) * sqrt(gm);
                        } else if (tm.coef_kind == COEF_NONE) {          // live sign, O&F 7.6 (levelsetterms.jl:237-242)
                            const double nrm = sqrt(qc > T(0) ? gp : gm);
                            const double den = sqrt(double(T(qc * qc)) + (nrm * nrm) * (P.dxmin * P.dxmin));
                            const double S = den == 0.0 ? 0.0 : double(qc) / den;
                            H = S * (nrm - 1.0);
                        } else {                                         // frozen sign, O&F 7.5 (levelsetterms.jl:243-247)
                            const double S0 = coef(tm, kk, 0);
                            H = S0 * (sqrt(S0 > 0 ? gp : gm) - 1.0);
                        }
                    } else if ((MASK & M_CURV) && (ONE || tm.kind == TERM_CURVATURE)) {
                        // levelsetterms.jl:111-121 + levelsetops.jl:197-244:  b * kappa * |grad phi| = b * (tr(H) q - g'Hg) / q
                        double g[3] = {0, 0, 0}, Hd[3] = {0, 0, 0};
#pragma unroll
                        for (int d = 0; d < NDIM; ++d) {
                            const T pp = at(d, 1), pm = at(d, -1);
                            g[d] = double(T(pp - pm)) * (0.5 * ih[d]);
                            Hd[d] = double(T(pp - T(2) * qc + pm)) * (ih[d] * ih[d]);
                        }
                        auto mixed = [&](int d1, int d2) -> double {
                            const double a = double(T(at2(d1, 1, d2, 1) - at2(d1, 1, d2, -1)));
                            const double b = double(T(at2(d1, -1, d2, 1) - at2(d1, -1, d2, -1)));
                            return (a - b) * (0.25 * ih[d1] * ih[d2]);
                        };
                        const double h01 = mixed(0, 1);
                        double q = fma(g[1], g[1], g[0] * g[0]);
                        double tr = Hd[0] + Hd[1];
                        double quad = fma(Hd[1] * g[1], g[1], fma(Hd[0] * g[0], g[0], 2.0 * (h01 * g[0] * g[1])));
                        if (NDIM == 3) {
                            const double h02 = mixed(0, 2), h12 = mixed(1, 2);
                            q = fma(g[2], g[2], q);
                            tr += Hd[2];
                            quad = fma(Hd[2] * g[2], g[2], quad) + 2.0 * (h02 * g[0] * g[2] + h12 * g[1] * g[2]);
                        }
                        const double eps = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
                        const double b = coef(tm, kk, 0);
                        H = q < eps ? b * 0.0 : b * (fma(tr, q, -quad) * fast_rcp(q));
                    }
                    x = T(fma(-P.c, H, double(x)));
                    if (P.out2) x2 = T(fma(-P.c2, H, double(x2)));
                };
                if (NTS > 0) {
#pragma unroll
                    for (int kk = 0; kk < NTS; ++kk) one_term(P.terms[kk], kk);
                } else {
                    for (int kk = 0; kk < P.nterms; ++kk) one_term(P.terms[kk], kk);
                }
                const long lin = (long)i + (long)j * vs1 + (long)z * vs2;
                P.out[lin] = x;
                if (P.out2) P.out2[lin] = x2;
            }
        }
        if (NDIM == 3) {
            cp_async_wait_all();
            __syncthreads();
        }
    }
}

#ifndef LSM_TX
#define LSM_TX 32
#define LSM_TY 8
#define LSM_NY 2
#define LSM_MINB 2
#endif

template <class T, int NDIM, int MASK, int NTS, int COEFK, bool REMAP>
cudaError_t launch_tiled(const StageParams<T>& P, const AuxList& A, cudaStream_t s) {
    constexpr int TX = LSM_TX, TY = LSM_TY, NY = LSM_NY;
    using G = TileGeom<T, NDIM, TX, TY, NY>;
    auto kern = stage_tiled_kernel<T, NDIM, MASK, NTS, COEFK, REMAP, TX, TY, NY, LSM_MINB>;
    const size_t smem = G::smem_bytes(A.n);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    const View<T>& v = P.in;
    dim3 block(TX, TY), grid;
    int cz = 1;
    if (NDIM == 3) {
        const int nr = P.r1 - P.r0;
        cz = nr >= 128 ? 64 : (nr >= 32 ? 32 : nr);
        grid = dim3((v.n[0] + TX - 1) / TX, (v.n[1] + TY * NY - 1) / (TY * NY), (nr + cz - 1) / cz);
    } else {
        grid = dim3((v.n[0] + TX - 1) / TX, (v.n[1] + TY * NY - 1) / (TY * NY), 1);
    }
    kern<<<grid, block, smem, s>>>(P, A, cz);
    return cudaGetLastError();
}

template <class T, int NDIM, bool REMAP>
cudaError_t launch_by_mask(int mask, const StageParams<T>& P, const AuxList& A, cudaStream_t s) {
    if constexpr (REMAP) {
        switch (mask) {
            case M_ADV_WENO:
                if (P.nterms == 1) {
                    const TermDev& t0 = P.terms[0];
                    if (t0.coef_kind == COEF_FIELD && A.first[0] == 0) return launch_tiled<T, NDIM, M_ADV_WENO, 1, COEF_FIELD, REMAP>(P, A, s);
                    if (t0.coef_kind == COEF_SEPARABLE) return launch_tiled<T, NDIM, M_ADV_WENO, 1, COEF_SEPARABLE, REMAP>(P, A, s);
                    if (t0.coef_kind == COEF_CONST) return launch_tiled<T, NDIM, M_ADV_WENO, 1, COEF_CONST, REMAP>(P, A, s);
                }
                break;
            case M_EIK:                 if (P.nterms == 1) return launch_tiled<T, NDIM, M_EIK, 1, -1, REMAP>(P, A, s); break;
            case M_NORMAL | M_ADV_WENO: if (P.nterms == 2) return launch_tiled<T, NDIM, M_NORMAL | M_ADV_WENO, 2, -1, REMAP>(P, A, s); break;
            case M_ADV_WENO | M_CURV:   if (P.nterms == 2) return launch_tiled<T, NDIM, M_ADV_WENO | M_CURV, 2, -1, REMAP>(P, A, s); break;
            default: break;
        }
    }
    return launch_tiled<T, NDIM, M_ALL, 0, -1, REMAP>(P, A, s);
}

int term_mask(const TermDev& t) {
    switch (t.kind) {
        case TERM_ADVECTION: return t.scheme == SCHEME_WENO5 ? M_ADV_WENO : M_ADV_UPWIND;
        case TERM_NORMAL:    return M_NORMAL;
        case TERM_CURVATURE: return M_CURV;
        default:             return M_EIK;
    }
}

}  // namespace

template <class T>
bool stage_tiled_supported(int ndim, const StageParams<T>& P) {
    if (ndim != 2 && ndim != 3) return false;
    if (P.nterms < 1 || P.nterms > 4) return false;
    for (int d = 0; d < ndim; ++d) if (P.in.n[d] < 8) return false;      // tiny grids: strict kernel
    int naux = P.p0 ? 1 : 0;
    for (int k = 0; k < P.nterms; ++k) {
        const TermDev& t = P.terms[k];
        if (t.coef_kind == COEF_FIELD && !(t.coef_f64 && sizeof(T) == 4)) naux += t.kind == TERM_ADVECTION ? ndim : 1;
        if (t.coef_kind == COEF_SEPARABLE && t.kind != TERM_ADVECTION) return false;
    }
    return naux <= 8;
}

template <class T>
cudaError_t launch_stage_tiled(int ndim, const StageParams<T>& P, int sm_count, cudaStream_t s) {
    if (!stage_tiled_supported<T>(ndim, P)) return cudaErrorNotSupported;
    if (P.r1 <= P.r0) return cudaSuccess;
    AuxList A{};
    A.n = 0; A.p0 = -1;
    int mask = 0;
    for (int k = 0; k < 4; ++k) A.first[k] = -1;
    for (int k = 0; k < P.nterms; ++k) {
        const TermDev& t = P.terms[k];
        mask |= term_mask(t);
        if (t.coef_kind == COEF_FIELD && !(t.coef_f64 && sizeof(T) == 4)) {
            A.first[k] = A.n;
            const int nc = t.kind == TERM_ADVECTION ? ndim : 1;
            for (int d = 0; d < nc; ++d) A.src[A.n++] = static_cast<const T*>(t.coef) + (long)d * t.cstride;
        }
    }
    if (P.p0) { A.p0 = A.n; A.src[A.n++] = P.p0; }
    bool remap = true;     // every BC an index map?  (ExtrapolationBC{P>=1} is a weighted stencil)
    for (int d = 0; d < ndim; ++d)
        for (int sd = 0; sd < 2; ++sd)
            if (P.in.bc[d][sd].kind == BC_EXTRAP && P.in.bc[d][sd].P > 0) remap = false;
    if (ndim == 3) return remap ? launch_by_mask<T, 3, true>(mask, P, A, s) : launch_by_mask<T, 3, false>(mask, P, A, s);
    return remap ? launch_by_mask<T, 2, true>(mask, P, A, s) : launch_by_mask<T, 2, false>(mask, P, A, s);
}

template bool stage_tiled_supported<float>(int, const StageParams<float>&);
template bool stage_tiled_supported<double>(int, const StageParams<double>&);
template cudaError_t launch_stage_tiled<float>(int, const StageParams<float>&, int, cudaStream_t);
template cudaError_t launch_stage_tiled<double>(int, const StageParams<double>&, int, cudaStream_t);

}  // namespace lsm
