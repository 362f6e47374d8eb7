// lsm_tiled.cu — performance kernels: fused RK-stage stencil kernels for 3-D grids.
//
// K1: one kernel per RK stage = ghost resolution + upwind WENO5 differences +
// the term's Hamiltonian + the stage combination, on a shared-memory ring of 2*3+2 z-planes.
//
//   * thread block = TX x TY threads owning an x-y tile of TX x (TY*NY) nodes; it marches along z
//     over a chunk of planes.  Every plane (tile + 3-cell halo, corners included) is brought into
//     the ring ONCE with cp.async (LDGSTS, no register staging) one iteration ahead of its first
//     use; all 19 stencil reads of a node then come from shared memory.  Tiles that touch a
//     physical boundary resolve their ghost cells while filling the ring, with the same code as the
//     strict kernel (lsm_bc.cuh), so every BC kind is supported.
//   * the B200 FP64 pipe (measured 63 lane-ops/clk/SM, tools/fp64_peak.cu) is a co-bound of this
//     stencil, so the WENO5 evaluation is restructured to minimise DP issue slots while staying
//     within 1e-10 of the reference:
//       - samples are read in UPWIND ORDER q_k = phi[i - s*(3-k)], s = sign(u), which turns
//         u * (u > 0 ? weno5- : weno5+) into (|u|/h) * W(q) with no selects and no sign logic;
//       - W works on undivided differences (WENO5 is homogeneous of degree 1; the epsilon floor is
//         rescaled), smoothness indicators and candidates are written on second differences,
//         the three weight divisions + three normalisations become ONE reciprocal
//         (MUFU.RCP64H + 2 Newton steps), and max(v^2) runs on the integer pipe.
//     ~47 DP instructions per WENO5 evaluation instead of ~65 + 6 divisions (13.6 slots each).
//
// Compiled with FMA contraction ON.  Parity with the CPU oracle is checked in tests/ (<= 1e-10 after
// 100 RK3 steps in Float64).
#include "lsm_dev.cuh"
#include "lsm_bc.cuh"
#include "lsm_kernels.h"

namespace lsm {

namespace {

constexpr int HAL = 3;           // WENO5 reach
constexpr int RING = 2 * HAL + 2;

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
// eps_floor replaces 4e-99*h^2 (which would underflow in the product form): 1e-70.  It only matters
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

template <class T, int TX, int TY, int NY>
struct TileGeom {
    static constexpr int W = TX + 2 * HAL;
    static constexpr int HH = TY * NY + 2 * HAL;
    static constexpr int PLANE = W * HH;          // phi plane: tile + halo
    static constexpr int TILE = TX * TY * NY;     // owned nodes of one plane of the tile
    static constexpr int NT = TX * TY;
    static constexpr int NW = NT / 32;
    static constexpr size_t smem_bytes(bool field_coef, bool has_p0) {
        return ((size_t)RING * PLANE + (field_coef ? 2 * 3 * TILE : 0) + (has_p0 ? 2 * TILE : 0)) * sizeof(T);
    }
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

// Fused stage kernel, 3-D, single AdvectionTerm with WENO5 (levelsetterms.jl:73-82 + derivatives.jl:61-121
// + timestepping.jl:128-202).  Shared memory: [ring of RING phi planes][2 x 3 velocity tiles][2 x phi^n tiles];
// everything a plane needs is issued with cp.async while the previous plane is being computed.
// REMAP = every BC of the field is an index map (periodic / Neumann / symmetry / halo): ghosts are filled by
// copying from the remapped address, no arithmetic.  Otherwise (ExtrapolationBC{P>=1}) boundary tiles go through
// getindex_slow (lsm_bc.cuh).
template <class T, int COEF, bool HAS_P0, bool REMAP, int TX, int TY, int NY, int MINB>
__global__ void __launch_bounds__(TX * TY, MINB)
adv_weno5_3d_kernel(const __grid_constant__ StageParams<T> P, const int cz) {
    using G = TileGeom<T, TX, TY, NY>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const ring = reinterpret_cast<T*>(smem_raw);
    T* const ubuf = ring + (size_t)RING * G::PLANE;                                  // [2][3][TILE]   (COEF_FIELD)
    T* const pbuf = ubuf + (COEF == COEF_FIELD ? 2 * 3 * G::TILE : 0);               // [2][TILE]      (HAS_P0)

    const int n0 = P.in.n[0], n1 = P.in.n[1], n2 = P.in.n[2];
    const long vs1 = P.in.s1, vs2 = P.in.s2;
    const T* __restrict__ const vp = P.in.p;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int lane = tx & 31, warp = (ty * TX + tx) >> 5;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * (TY * NY);
    const int zbeg = P.r0 + blockIdx.z * cz;
    const int zend = min(P.r1, zbeg + cz);
    if (zbeg >= zend) return;

    // whole tile + halo inside the stored x-y extent -> plain copies; otherwise resolve ghosts
    const bool xy_in = (x0 - HAL >= 0) && (x0 + TX + HAL <= n0) && (y0 - HAL >= 0) && (y0 + TY * NY + HAL <= n1);
    const int kzl = P.in.bc[2][0].kind, kzh = P.in.bc[2][1].kind;

    // one warp per row of the (tile + halo) plane: lanes 0..31 copy columns 0..31, lanes 0..W-33 also 32..W-1
    auto load_phi = [&](int z) {
        T* dst = ring + ((z + 1024) & (RING - 1)) * G::PLANE;
        const bool z_plain = (z >= 0 || kzl == BC_HALO) && (z < n2 || kzh == BC_HALO);
        if (xy_in && z_plain) {
            const T* src = vp + (long)(x0 - HAL) + (long)(y0 - HAL) * vs1 + (long)z * vs2;
            for (int r = warp; r < G::HH; r += G::NW) {
                const T* s = src + (long)r * vs1;
                T* d = dst + r * G::W;
                cp_async(d + lane, s + lane, sizeof(T) == 8);
                if (lane < G::W - 32) cp_async(d + 32 + lane, s + 32 + lane, sizeof(T) == 8);
            }
        } else if (REMAP) {
            const int zz = remap_index(z, n2, kzl, kzh);
            const int gxa = remap_index(x0 - HAL + lane, n0, P.in.bc[0][0].kind, P.in.bc[0][1].kind);
            const int gxb = remap_index(x0 - HAL + 32 + lane, n0, P.in.bc[0][0].kind, P.in.bc[0][1].kind);
            for (int r = warp; r < G::HH; r += G::NW) {
                const int gy = min(max(remap_index(y0 - HAL + r, n1, P.in.bc[1][0].kind, P.in.bc[1][1].kind), 0), n1 - 1);
                const T* s = vp + (long)gy * vs1 + (long)zz * vs2;
                T* d = dst + r * G::W;
                // columns beyond the grid + halo of a partial tile are never read: clamp them into range
                cp_async(d + lane, s + min(max(gxa, 0), n0 - 1), sizeof(T) == 8);
                if (lane < G::W - 32) cp_async(d + 32 + lane, s + min(max(gxb, 0), n0 - 1), sizeof(T) == 8);
            }
        } else {
            for (int r = warp; r < G::HH; r += G::NW) {
                T* d = dst + r * G::W;
                d[lane] = getindex_slow<3, T>(P.in, x0 - HAL + lane, y0 - HAL + r, z);
                if (lane < G::W - 32) d[32 + lane] = getindex_slow<3, T>(P.in, x0 - HAL + 32 + lane, y0 - HAL + r, z);
            }
        }
    };
    // velocity components and phi^n of the owned nodes of plane z (coefficient boxes carry no ghosts)
    auto load_aux = [&](int z) {
        const int slot = z & 1;
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int r = ty + k * TY;
            const int ii = x0 + tx, jj = y0 + r;
            if (ii < n0 && jj < n1) {
                if (COEF == COEF_FIELD) {
                    const T* src = static_cast<const T*>(P.terms[0].coef) + (long)ii + (long)n0 * jj + (long)n0 * n1 * z;
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        cp_async(ubuf + (slot * 3 + d) * G::TILE + r * TX + tx, src + (long)d * P.terms[0].cstride, sizeof(T) == 8);
                }
                if (HAS_P0)
                    cp_async(pbuf + slot * G::TILE + r * TX + tx, P.p0 + (long)ii + (long)jj * vs1 + (long)z * vs2, sizeof(T) == 8);
            }
        }
    };

    for (int z = zbeg - HAL; z <= zbeg + HAL; ++z) load_phi(z);
    load_aux(zbeg);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    const double ih0 = 1.0 / P.h[0], ih1 = 1.0 / P.h[1], ih2 = 1.0 / P.h[2];
    const int i = x0 + tx;

    for (int z = zbeg; z < zend; ++z) {
        if (z + HAL + 1 <= zend - 1 + HAL) load_phi(z + HAL + 1);
        if (z + 1 < zend) load_aux(z + 1);
        cp_async_commit();

        const T* cur = ring + ((z + 1024) & (RING - 1)) * G::PLANE;
        const int aslot = z & 1;
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int r = ty + k * TY;
            const int j = y0 + r;
            if (i < n0 && j < n1) {
                const int sc = (r + HAL) * G::W + tx + HAL;
                const int st = r * TX + tx;
                double u0, u1, u2;
                if (COEF == COEF_FIELD) {
                    u0 = double(ubuf[(aslot * 3 + 0) * G::TILE + st]);
                    u1 = double(ubuf[(aslot * 3 + 1) * G::TILE + st]);
                    u2 = double(ubuf[(aslot * 3 + 2) * G::TILE + st]);
                } else if (COEF == COEF_SEPARABLE) {
                    const TermDev& tm = P.terms[0];
                    u0 = ((tm.cval[0] * __ldg(tm.tab[0][0] + i)) * __ldg(tm.tab[0][1] + j)) * __ldg(tm.tab[0][2] + z);
                    u1 = ((tm.cval[1] * __ldg(tm.tab[1][0] + i)) * __ldg(tm.tab[1][1] + j)) * __ldg(tm.tab[1][2] + z);
                    u2 = ((tm.cval[2] * __ldg(tm.tab[2][0] + i)) * __ldg(tm.tab[2][1] + j)) * __ldg(tm.tab[2][2] + z);
                } else {
                    u0 = P.terms[0].cval[0]; u1 = P.terms[0].cval[1]; u2 = P.terms[0].cval[2];
                }
                if (P.terms[0].scaled) { const double g = P.terms[0].g; u0 = u0 * g; u1 = u1 * g; u2 = u2 * g; }
                const T qc = cur[sc];
                // upwind-ordered sampling: q_k = phi[i - s*(3-k)]; s = +1 when the velocity is > 0
                const int s0 = u0 > 0 ? 1 : -1, s1 = u1 > 0 ? G::W : -G::W, s2 = u2 > 0 ? 1 : -1;
                const T* c0 = cur + sc;
                const double w0 = weno5_up<T>(c0[-3 * s0], c0[-2 * s0], c0[-s0], qc, c0[s0], c0[2 * s0]);
                const double w1 = weno5_up<T>(c0[-3 * s1], c0[-2 * s1], c0[-s1], qc, c0[s1], c0[2 * s1]);
                auto zp = [&](int m) -> T { return ring[((z + m * s2 + 1024) & (RING - 1)) * G::PLANE + sc]; };
                const double w2 = weno5_up<T>(zp(-3), zp(-2), zp(-1), qc, zp(1), zp(2));
                // H = sum_d u_d * weno(d) = sum_d (|u_d| / h_d) * W_d        (left-to-right like the reference)
                double H = (fabs(u0) * ih0) * w0;
                H = fma(fabs(u1) * ih1, w1, H);
                H = fma(fabs(u2) * ih2, w2, H);

                T x = qc;
                if (HAS_P0) {
                    const T pn = pbuf[aslot * G::TILE + st];
                    if (P.base == BASE_RK3_S2) x = T(fma(0.75, double(pn), 0.25 * double(qc)));       // timestepping.jl:183
                    else if (P.base == BASE_RK3_S3) x = T((pn + T(2) * qc) / T(3));                   // timestepping.jl:194
                    else x = pn;                                                                       // RK2 S2 (corr)
                }
                const long lin = (long)i + (long)j * vs1 + (long)z * vs2;
                P.out[lin] = T(fma(-P.c, H, double(x)));
                if (!HAS_P0 && P.out2) P.out2[lin] = T(fma(-P.c2, H, double(qc)));
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }
}

#ifndef LSM_TX
#define LSM_TX 32
#define LSM_TY 8
#define LSM_NY 2
#define LSM_MINB 2
#endif

template <class T, int COEF, bool HAS_P0, bool REMAP>
cudaError_t launch_adv3d(const StageParams<T>& P, int sm_count, cudaStream_t s) {
    constexpr int TX = LSM_TX, TY = LSM_TY, NY = LSM_NY;
    using G = TileGeom<T, TX, TY, NY>;
    auto kern = adv_weno5_3d_kernel<T, COEF, HAS_P0, REMAP, TX, TY, NY, LSM_MINB>;
    const size_t smem = G::smem_bytes(COEF == COEF_FIELD, HAS_P0);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const View<T>& v = P.in;
    const int nr = P.r1 - P.r0;
    const int cz = nr >= 128 ? 64 : (nr >= 32 ? 32 : nr);
    dim3 grid((v.n[0] + TX - 1) / TX, (v.n[1] + TY * NY - 1) / (TY * NY), (nr + cz - 1) / cz);
    dim3 block(TX, TY);
    kern<<<grid, block, smem, s>>>(P, cz);
    return cudaGetLastError();
}

template <class T, int COEF>
cudaError_t launch_adv3d_base(const StageParams<T>& P, int sm_count, cudaStream_t s) {
    bool remap = true;     // every BC an index map?  (ExtrapolationBC{P>=1} is a weighted stencil)
    for (int d = 0; d < 3; ++d)
        for (int sd = 0; sd < 2; ++sd)
            if (P.in.bc[d][sd].kind == BC_EXTRAP && P.in.bc[d][sd].P > 0) remap = false;
    if (P.base == BASE_IN) return remap ? launch_adv3d<T, COEF, false, true>(P, sm_count, s) : launch_adv3d<T, COEF, false, false>(P, sm_count, s);
    return remap ? launch_adv3d<T, COEF, true, true>(P, sm_count, s) : launch_adv3d<T, COEF, true, false>(P, sm_count, s);
}

}  // namespace

template <class T>
bool stage_tiled_supported(int ndim, const StageParams<T>& P) {
    if (ndim != 3 || P.nterms != 1) return false;
    const TermDev& t = P.terms[0];
    if (t.kind != TERM_ADVECTION || t.scheme != SCHEME_WENO5) return false;
    if (t.coef_kind == COEF_FIELD && t.coef_f64 && sizeof(T) == 4) return false;
    if (P.in.n[0] < 8 || P.in.n[1] < 8 || P.in.n[2] < 4) return false;     // tiny grids: strict kernel
    return true;
}

template <class T>
cudaError_t launch_stage_tiled(int ndim, const StageParams<T>& P, int sm_count, cudaStream_t s) {
    if (!stage_tiled_supported<T>(ndim, P)) return cudaErrorNotSupported;
    if (P.r1 <= P.r0) return cudaSuccess;
    switch (P.terms[0].coef_kind) {
        case COEF_FIELD:     return launch_adv3d_base<T, COEF_FIELD>(P, sm_count, s);
        case COEF_SEPARABLE: return launch_adv3d_base<T, COEF_SEPARABLE>(P, sm_count, s);
        default:             return launch_adv3d_base<T, COEF_CONST>(P, sm_count, s);
    }
}

template bool stage_tiled_supported<float>(int, const StageParams<float>&);
template bool stage_tiled_supported<double>(int, const StageParams<double>&);
template cudaError_t launch_stage_tiled<float>(int, const StageParams<float>&, int, cudaStream_t);
template cudaError_t launch_stage_tiled<double>(int, const StageParams<double>&, int, cudaStream_t);

}  // namespace lsm
