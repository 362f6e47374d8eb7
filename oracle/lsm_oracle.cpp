// lsm_oracle.cpp — CPU ORACLE (test infrastructure; see lsm_oracle.h for scope and pinning status).
//
// A restatement, function by function and in the reference's own operation order, of the
// dense-grid explicit time-integration path of LevelSetMethods.jl v0.2.0.  Every function
// cites the reference file:line it follows (paths relative to the reference repo root).
// Build with -O2 -ffp-contract=off (Julia never contracts a*b+c into an FMA).
//
// Precision model (SURVEY.md §8a "precision notes"): the field valtype V is float or double,
// the grid type is Float64.  All numeric literals on the path are Float64, so with V=Float32
// the first difference is taken in Float32 and everything after it is Float64, rounding back
// to V only on store.  Ghost accumulation stays in V (meshfield.jl:254-257).
//
// Known places where the exact Julia/StaticArrays summation order cannot be confirmed here
// (Julia is not installed): dot(g,g) and g'*H*g in curvature (levelsetops.jl:199-205).  They
// are restated left-to-right; a different association changes results by O(1 ulp).

#include "lsm_oracle.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

int g_threads = 1;

template <class T> struct Eps;
template <> struct Eps<float>  { static constexpr double v = 1.1920928955078125e-07; };   // eps(Float32)
template <> struct Eps<double> { static constexpr double v = 2.220446049250313e-16; };    // eps(Float64)

// Julia's min/max propagate NaN (Base.min/max for floats).
inline double jl_min(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? std::numeric_limits<double>::quiet_NaN() : (b < a ? b : a); }
inline double jl_max(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? std::numeric_limits<double>::quiet_NaN() : (b > a ? b : a); }

// levelsetterms.jl:180-181
inline double positive(double x) { return x > 0.0 ? x : 0.0; }
inline double negative(double x) { return x < 0.0 ? x : 0.0; }
// levelsetterms.jl:184-187 (minmod)
inline double limiter(double x, double y) {
    if (!(x * y > 0.0)) return 0.0;
    return std::fabs(x) <= std::fabs(y) ? x : y;
}

// boundaryconditions.jl:90-97
inline double lagrange_extrap_weight(int j, int k, int P) {
    double w = 1.0;
    for (int m = 0; m <= P; ++m) {
        if (m == j) continue;
        w *= double(-k - m) / double(j - m);
    }
    return w;
}

// derivatives.jl:61-81
inline double weno5(double v1, double v2, double v3, double v4, double v5) {
    const double c13 = 1.0 / 3.0, c76 = 7.0 / 6.0, c116 = 11.0 / 6.0, c16 = 1.0 / 6.0, c56 = 5.0 / 6.0;
    const double c1312 = 13.0 / 12.0, c14 = 1.0 / 4.0;
    double d1 = c13 * v1 - c76 * v2 + c116 * v3;
    double d2 = -c16 * v2 + c56 * v3 + c13 * v4;
    double d3 = c13 * v3 + c56 * v4 - c16 * v5;
    double a, b;
    a = v1 - 2 * v2 + v3;  b = v1 - 4 * v2 + 3 * v3;
    double S1 = c1312 * (a * a) + c14 * (b * b);
    a = v2 - 2 * v3 + v4;  b = v2 - v4;
    double S2 = c1312 * (a * a) + c14 * (b * b);
    a = v3 - 2 * v4 + v5;  b = 3 * v3 - 4 * v4 + v5;
    double S3 = c1312 * (a * a) + c14 * (b * b);
    double m = jl_max(jl_max(jl_max(jl_max(v1 * v1, v2 * v2), v3 * v3), v4 * v4), v5 * v5);
    double eps = 1.0e-6 * m + 1.0e-99;
    double t;
    t = S1 + eps; double a1 = 0.1 / (t * t);
    t = S2 + eps; double a2 = 0.6 / (t * t);
    t = S3 + eps; double a3 = 0.3 / (t * t);
    double w1 = a1 / (a1 + a2 + a3);
    double w2 = a2 / (a1 + a2 + a3);
    double w3 = a3 / (a1 + a2 + a3);
    return w1 * d1 + w2 * d2 + w3 * d3;
}

struct Idx { int i[3]; };
inline Idx shift(Idx I, int dim /*0-based*/, int nb) { I.i[dim] += nb; return I; }

// ---------------------------------------------------------------------------------------------
// Field view: meshfield.jl:51-55 (MeshField = vals + mesh + bcs)
// ---------------------------------------------------------------------------------------------
template <int N, class T>
struct Field {
    const T* v;
    int n[3], gl[3], gr[3];
    long stride[3];
    orc_bc bc[3][2];
    double h[3];      // meshes.jl:109-110 meshsize = (hc - lc) / (n - 1)   (global n for slabs)
    double lc[3];
    int off[3];

    explicit Field(const orc_field& f, const void* vals = nullptr) {
        v = static_cast<const T*>(vals ? vals : f.vals);
        long s = 1;
        for (int d = 0; d < 3; ++d) {
            n[d] = d < N ? f.n[d] : 1;
            gl[d] = d < N ? f.gl[d] : 0;
            gr[d] = d < N ? f.gr[d] : 0;
            stride[d] = s;
            s *= (long)(gl[d] + n[d] + gr[d]);
            bc[d][0] = f.bc[d][0]; bc[d][1] = f.bc[d][1];
            int ng = (d < N && f.nglob[d] > 0) ? f.nglob[d] : n[d];
            h[d] = d < N ? (f.hc[d] - f.lc[d]) / double(ng - 1) : 1.0;
            lc[d] = f.lc[d];
            off[d] = d < N ? f.off[d] : 0;
        }
    }
    inline bool stored(const Idx& I) const {
        for (int d = 0; d < N; ++d)
            if (I.i[d] < 1 - gl[d] || I.i[d] > n[d] + gr[d]) return false;
        return true;
    }
    inline bool ingrid(const Idx& I) const {
        for (int d = 0; d < N; ++d)
            if (I.i[d] < 1 || I.i[d] > n[d]) return false;
        return true;
    }
    inline long lin(const Idx& I) const {
        long l = 0;
        for (int d = 0; d < N; ++d) l += (long)(I.i[d] - 1 + gl[d]) * stride[d];
        return l;
    }
    inline T raw(const Idx& I) const { return v[lin(I)]; }
};

template <int N, class T> T getindex(const Field<N, T>& f, const Idx& I, int depth = 0);

// meshfield.jl:248-260 (_getindexbc), boundaryconditions.jl:107-153 (bc_stencil)
template <int N, class T, int DIM>
T getindexbc(const Field<N, T>& f, const Idx& I, int depth) {
    if constexpr (DIM == 0) {
        // "At level 0 every component is back in range, so the value is read through the field's
        //  own in-grid getindex" — which re-enters the BC path if it is not (tiny grids).
        if (f.stored(I)) {
            bool ok = true;   // inside the owned box or inside a stored HALO plane
            for (int d = 0; d < N; ++d) {
                if (I.i[d] < 1 && f.bc[d][0].kind != ORC_BC_HALO) ok = false;
                if (I.i[d] > f.n[d] && f.bc[d][1].kind != ORC_BC_HALO) ok = false;
            }
            if (ok) return f.raw(I);
        }
        if (depth > 64) return std::numeric_limits<T>::quiet_NaN();
        return getindex<N, T>(f, I, depth + 1);
    } else {
        constexpr int d = DIM - 1;
        const int i = I.i[d], n = f.n[d];
        if (i >= 1 && i <= n) return getindexbc<N, T, DIM - 1>(f, I, depth);
        const orc_bc bc = i < 1 ? f.bc[d][0] : f.bc[d][1];
        if (bc.kind == ORC_BC_HALO) return getindexbc<N, T, DIM - 1>(f, I, depth);  // stored ghost plane
        T acc = T(0);
        if (bc.kind == ORC_BC_PERIODIC) {
            // boundaryconditions.jl:107-119: i<1 -> n-(1-i) ; i>n -> 1+(i-n)   (node n duplicates node 1)
            Idx J = I; J.i[d] = i < 1 ? (n - (1 - i)) : (1 + (i - n));
            acc += T(1.0) * getindexbc<N, T, DIM - 1>(f, J, depth);
        } else if (bc.kind == ORC_BC_EXTRAP) {
            // boundaryconditions.jl:134-144
            const int k = i < 1 ? (1 - i) : (i - n);
            const int b = i < 1 ? 1 : n;
            const int dd = i < 1 ? 1 : -1;
            for (int j = 0; j <= bc.P; ++j) {
                Idx J = I; J.i[d] = b + dd * j;
                acc += T(lagrange_extrap_weight(j, k, bc.P)) * getindexbc<N, T, DIM - 1>(f, J, depth);
            }
        } else if (bc.kind == ORC_BC_SYMMETRY) {
            // boundaryconditions.jl:146-153
            const int k = i < 1 ? (1 - i) : (i - n);
            const int b = i < 1 ? 1 : n;
            const int dd = i < 1 ? 1 : -1;
            Idx J = I; J.i[d] = b + dd * k;
            acc += T(1.0) * getindexbc<N, T, DIM - 1>(f, J, depth);
        } else {
            return std::numeric_limits<T>::quiet_NaN();   // reference throws (meshfield.jl:222-232)
        }
        return acc;
    }
}

// meshfield.jl:213-217
template <int N, class T>
T getindex(const Field<N, T>& f, const Idx& I, int depth) {
    if (f.ingrid(I)) return f.raw(I);
    return getindexbc<N, T, N>(f, I, depth);
}

// ---------------------------------------------------------------------------------------------
// derivatives.jl:28-175.  dim is 0-based here.
// ---------------------------------------------------------------------------------------------
template <int N, class T> inline double D0(const Field<N, T>& f, const Idx& I, int dim) {
    const double h = f.h[dim];
    T d = getindex(f, shift(I, dim, 1)) - getindex(f, shift(I, dim, -1));
    return double(d) / (2 * h);
}
template <int N, class T> inline double Dp(const Field<N, T>& f, const Idx& I, int dim) {
    const double h = f.h[dim];
    T d = getindex(f, shift(I, dim, 1)) - getindex(f, I);
    return double(d) / h;
}
template <int N, class T> inline double Dm(const Field<N, T>& f, const Idx& I, int dim) {
    const double h = f.h[dim];
    T d = getindex(f, I) - getindex(f, shift(I, dim, -1));
    return double(d) / h;
}
// derivatives.jl:89-101
template <int N, class T> inline double weno5m(const Field<N, T>& f, const Idx& I, int dim) {
    return weno5(Dm(f, shift(I, dim, -2), dim), Dm(f, shift(I, dim, -1), dim), Dm(f, I, dim),
                 Dm(f, shift(I, dim, 1), dim), Dm(f, shift(I, dim, 2), dim));
}
// derivatives.jl:109-121
template <int N, class T> inline double weno5p(const Field<N, T>& f, const Idx& I, int dim) {
    return weno5(Dp(f, shift(I, dim, 2), dim), Dp(f, shift(I, dim, 1), dim), Dp(f, I, dim),
                 Dp(f, shift(I, dim, -1), dim), Dp(f, shift(I, dim, -2), dim));
}
// derivatives.jl:129-134
template <int N, class T> inline double D20(const Field<N, T>& f, const Idx& I, int dim) {
    const double h = f.h[dim];
    T d = getindex(f, shift(I, dim, 1)) - T(2) * getindex(f, I) + getindex(f, shift(I, dim, -1));
    return double(d) / (h * h);
}
// derivatives.jl:144-149
template <int N, class T> inline double D2mixed(const Field<N, T>& f, const Idx& I, int d1, int d2) {
    const double h = f.h[d1];
    return (D0(f, shift(I, d1, 1), d2) - D0(f, shift(I, d1, -1), d2)) / (2 * h);
}
// derivatives.jl:157-162
template <int N, class T> inline double D2pp(const Field<N, T>& f, const Idx& I, int dim) {
    const double h = f.h[dim];
    T d = getindex(f, I) - T(2) * getindex(f, shift(I, dim, 1)) + getindex(f, shift(I, dim, 2));
    return double(d) / (h * h);
}
// derivatives.jl:170-175
template <int N, class T> inline double D2mm(const Field<N, T>& f, const Idx& I, int dim) {
    const double h = f.h[dim];
    T d = getindex(f, shift(I, dim, -2)) - T(2) * getindex(f, shift(I, dim, -1)) + getindex(f, I);
    return double(d) / (h * h);
}

// levelsetops.jl:197-205, 212-215, 234-244
template <int N, class T> inline double curvature(const Field<N, T>& f, const Idx& I) {
    double g[3];
    for (int d = 0; d < N; ++d) g[d] = D0(f, I, d);
    double nrmsq = 0.0 * 0.0 + 0.0 * 0.0;            // StaticArrays _vecdot seed
    for (int d = 0; d < N; ++d) nrmsq += g[d] * g[d];
    if (nrmsq < Eps<T>::v) return 0.0;
    double H[3][3];
    for (int j = 0; j < N; ++j)
        for (int i = 0; i <= j; ++i) {
            // Symmetric(H) reads the upper triangle: H[i,j], i<j, is D2(phi, I, (i, j))
            H[i][j] = (i == j) ? D20(f, I, i) : D2mixed(f, I, i, j);
            H[j][i] = H[i][j];
        }
    double tr = H[0][0];
    for (int d = 1; d < N; ++d) tr += H[d][d];
    double quad = 0.0;
    bool first = true;
    for (int j = 0; j < N; ++j) {
        double w = g[0] * H[0][j];
        for (int i = 1; i < N; ++i) w += g[i] * H[i][j];
        if (first) { quad = w * g[j]; first = false; } else quad += w * g[j];
    }
    return (tr * nrmsq - quad) / std::pow(nrmsq, 3.0 / 2.0);
}

// ---------------------------------------------------------------------------------------------
// coefficients: levelsetterms.jl:42-43 (_eval_field) + this engine's "field x g(t)" and
// separable-table kinds (SURVEY.md §8b coefficient kinds)
// ---------------------------------------------------------------------------------------------
template <int N, class T>
struct Coef {
    const orc_term& tm;
    const Field<N, T>& phi;
    Coef(const orc_term& t, const Field<N, T>& p) : tm(t), phi(p) {}
    // linear index of an in-grid node in a coefficient array WITHOUT ghost planes
    inline long node(const Idx& I) const {
        long l = 0, s = 1;
        for (int d = 0; d < N; ++d) { l += (long)(I.i[d] - 1) * s; s *= phi.n[d]; }
        return l;
    }
    inline double comp(const Idx& I, int d, int ncomp) const {
        switch (tm.coef_kind) {
            case ORC_COEF_CONST: return tm.cval[d];
            case ORC_COEF_FIELD: {
                long l = node(I) * ncomp + d;
                return tm.field_dtype == ORC_F32 ? double(static_cast<const float*>(tm.field)[l])
                                                 : static_cast<const double*>(tm.field)[l];
            }
            case ORC_COEF_SEPARABLE: {
                double r = tm.cval[d];
                for (int a = 0; a < N; ++a) r = r * tm.tab[d][a][I.i[a] - 1];
                return r;
            }
            default: return 0.0;
        }
    }
};

// levelsetterms.jl:73-82
template <int N, class T>
inline double term_advection(const Field<N, T>& f, const orc_term& tm, const Idx& I, double g) {
    Coef<N, T> c(tm, f);
    double s = 0.0;
    for (int d = 0; d < N; ++d) {
        double v = c.comp(I, d, N);
        if (tm.tscale_kind != ORC_TS_NONE) v = v * g;
        double der;
        if (tm.scheme == ORC_WENO5) der = (v > 0) ? weno5m(f, I, d) : weno5p(f, I, d);
        else                        der = (v > 0) ? Dm(f, I, d) : Dp(f, I, d);
        double p = v * der;
        s = (d == 0) ? p : s + p;
    }
    return s;
}

// levelsetterms.jl:156-170 / 252-265 : second-order ENO one-sided pair for one dim
template <int N, class T>
inline void eno2_pair(const Field<N, T>& f, const Idx& I, int d, double& neg, double& pos) {
    const double h = f.h[d];
    const double d20 = D20(f, I, d);
    neg = Dm(f, I, d) + (0.5 * h) * limiter(D2mm(f, I, d), d20);
    pos = Dp(f, I, d) - (0.5 * h) * limiter(D2pp(f, I, d), d20);   // D2⁰ is recomputed in the reference; same value
}

// levelsetterms.jl:156-170
template <int N, class T>
inline double term_normal(const Field<N, T>& f, const orc_term& tm, const Idx& I, double g) {
    Coef<N, T> c(tm, f);
    double v = c.comp(I, 0, 1);
    if (tm.tscale_kind != ORC_TS_NONE) v = v * g;
    double gp = 0, gm = 0;
    for (int d = 0; d < N; ++d) {
        double neg, pos;
        eno2_pair(f, I, d, neg, pos);
        double a = positive(neg) * positive(neg) + negative(pos) * negative(pos);
        double b = negative(neg) * negative(neg) + positive(pos) * positive(pos);
        gp = (d == 0) ? a : gp + a;
        gm = (d == 0) ? b : gm + b;
    }
    return positive(v) * std::sqrt(gp) + negative(v) * std::sqrt(gm);
}

// levelsetterms.jl:111-121
template <int N, class T>
inline double term_curvature(const Field<N, T>& f, const orc_term& tm, const Idx& I, double g) {
    Coef<N, T> c(tm, f);
    double kappa = curvature(f, I);
    double b = c.comp(I, 0, 1);
    if (tm.tscale_kind != ORC_TS_NONE) b = b * g;
    double p2 = 0;
    for (int d = 0; d < N; ++d) {
        double q = D0(f, I, d);
        p2 = (d == 0) ? q * q : p2 + q * q;
    }
    return b * kappa * std::sqrt(p2);
}

// levelsetterms.jl:252-265
template <int N, class T>
inline double grad_norm_godunov(const Field<N, T>& f, const Idx& I, bool vpos) {
    double sa = 0, sb = 0;
    for (int d = 0; d < N; ++d) {
        double A, B;
        eno2_pair(f, I, d, A, B);
        double a, b;
        if (vpos) { a = positive(A) * positive(A); b = negative(B) * negative(B); }
        else      { a = negative(A) * negative(A); b = positive(B) * positive(B); }
        sa = (d == 0) ? a : sa + a;
        sb = (d == 0) ? b : sb + b;
    }
    return std::sqrt(sa + sb);
}

// levelsetterms.jl:234-248
template <int N, class T>
inline double term_eikonal(const Field<N, T>& f, const orc_term& tm, const Idx& I) {
    double dxmin = f.h[0];
    for (int d = 1; d < N; ++d) dxmin = std::min(dxmin, f.h[d]);
    if (tm.coef_kind == ORC_COEF_NONE) {
        T p = getindex(f, I);
        double nrm = grad_norm_godunov(f, I, p > T(0));
        double den = std::sqrt(double(T(p * p)) + (nrm * nrm) * (dxmin * dxmin));
        double S = (den == 0.0) ? 0.0 : double(p) / den;
        return S * (nrm - 1);
    } else {
        Coef<N, T> c(tm, f);
        double S0 = c.comp(I, 0, 1);
        double nrm = grad_norm_godunov(f, I, S0 > 0);
        return S0 * (nrm - 1);
    }
}

template <int N, class T>
inline double compute_term(const Field<N, T>& f, const orc_term& tm, const Idx& I, double g) {
    switch (tm.kind) {
        case ORC_TERM_ADVECTION: return term_advection(f, tm, I, g);
        case ORC_TERM_NORMAL:    return term_normal(f, tm, I, g);
        case ORC_TERM_CURVATURE: return term_curvature(f, tm, I, g);
        default:                 return term_eikonal(f, tm, I);
    }
}

// per-node CFL: levelsetterms.jl:90-96, 123-127, 172-178, 250
template <int N, class T>
inline double cfl_node(const Field<N, T>& f, const orc_term& tm, const Idx& I, double g) {
    Coef<N, T> c(tm, f);
    const bool sc = tm.tscale_kind != ORC_TS_NONE;
    switch (tm.kind) {
        case ORC_TERM_ADVECTION: {
            double s = 0;
            for (int d = 0; d < N; ++d) {
                double v = c.comp(I, d, N); if (sc) v = v * g;
                double q = std::fabs(v) / f.h[d];
                s = (d == 0) ? q : s + q;
            }
            return 1 / s;
        }
        case ORC_TERM_NORMAL: {
            double v = c.comp(I, 0, 1); if (sc) v = v * g;
            double s = 0;
            for (int d = 0; d < N; ++d) { double q = std::fabs(v) / f.h[d]; s = (d == 0) ? q : s + q; }
            return 1 / s;
        }
        case ORC_TERM_CURVATURE: {
            double b = c.comp(I, 0, 1); if (sc) b = b * g;
            double dx = f.h[0];
            for (int d = 1; d < N; ++d) dx = std::min(dx, f.h[d]);
            return (dx * dx) / (2 * std::fabs(b));
        }
        default: {
            double dx = f.h[0];
            for (int d = 1; d < N; ++d) dx = std::min(dx, f.h[d]);
            return dx;
        }
    }
}

// iterate all owned nodes, column-major, optionally with OpenMP over the slowest axis
template <int N, class F> inline void for_nodes(const int* n, F&& fn) {
    const int n1 = n[0], n2 = N > 1 ? n[1] : 1, n3 = N > 2 ? n[2] : 1;
    const long outer = (long)n2 * n3;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (long o = 0; o < outer; ++o) {
        Idx I; I.i[2] = int(o / n2) + 1; I.i[1] = int(o % n2) + 1;
        for (int i = 1; i <= n1; ++i) { I.i[0] = i; fn(I); }
    }
}

// levelsetterms.jl:30-37 (generic _compute_cfl: min over active nodes, starting from Inf)
template <int N, class T>
double cfl_term(const Field<N, T>& f, const orc_term& tm, double g) {
    const int n1 = f.n[0], n2 = N > 1 ? f.n[1] : 1, n3 = N > 2 ? f.n[2] : 1;
    const long outer = (long)n2 * n3;
    double dt = std::numeric_limits<double>::infinity();
    if (tm.kind == ORC_TERM_EIKONAL || tm.coef_kind == ORC_COEF_CONST) {
        // same value at every node; min(Inf, x, x, ...) == min(Inf, x)
        Idx I{{1, 1, 1}};
        return jl_min(dt, cfl_node(f, tm, I, g));
    }
    bool isnan = false;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1) reduction(min : dt) reduction(|| : isnan)
    for (long o = 0; o < outer; ++o) {
        Idx I; I.i[2] = int(o / n2) + 1; I.i[1] = int(o % n2) + 1;
        for (int i = 1; i <= n1; ++i) {
            I.i[0] = i;
            double c = cfl_node(f, tm, I, g);
            if (std::isnan(c)) isnan = true;
            else if (c < dt) dt = c;
        }
    }
    return isnan ? std::numeric_limits<double>::quiet_NaN() : dt;
}

double tscale(const orc_term& tm, double t, const double* gscale, int k) {
    switch (tm.tscale_kind) {
        case ORC_TS_COS:  return std::cos(M_PI * t / tm.tparam);
        case ORC_TS_HOST: return gscale ? gscale[k] : 1.0;
        default:          return 1.0;
    }
}

// One "for k in terms; for I: dst[I] -= c * H_k(src, I, t)" sweep set (timestepping.jl:131-136 etc.)
// dst is pre-seeded with the stage base.  dst2/c2 serve RK2's second accumulator (:149-155).
template <int N, class T>
void term_sweeps(const orc_field& desc, const void* src_vals, T* dst, double c, T* dst2, double c2,
                 const orc_term* terms, int nterms, double t, const double* gscale) {
    Field<N, T> src(desc, src_vals);
    for (int k = 0; k < nterms; ++k) {
        const orc_term& tm = terms[k];
        const double g = tscale(tm, t, gscale, k);
        for_nodes<N>(src.n, [&](const Idx& I) {
            const double H = compute_term(src, tm, I, g);
            const long l = src.lin(I);
            dst[l] = T(double(dst[l]) - c * H);
            if (dst2) dst2[l] = T(double(dst2[l]) - c2 * H);
        });
    }
}

template <int N, class T>
long total_elems(const orc_field& d) {
    long s = 1;
    for (int k = 0; k < N; ++k) s *= (long)(d.gl[k] + d.n[k] + d.gr[k]);
    return s;
}

// timestepping.jl:126-202
template <int N, class T>
int stage_impl(const orc_field& desc, int integ, int stage, T* phi, T* buf1, T* buf2,
               const orc_term* terms, int nterms, double tc, double dt, const double* gs) {
    Field<N, T> geo(desc, phi);
    const long tot = total_elems<N, T>(desc);
    if (integ == ORC_FE) {
        if (stage != 1) return -1;
        std::memcpy(buf1, phi, tot * sizeof(T));                                   // copy!(dst, phi)
        term_sweeps<N, T>(desc, phi, buf1, dt, nullptr, 0, terms, nterms, tc, gs);  // dst[I] -= dt*H
        std::memcpy(phi, buf1, tot * sizeof(T));                                   // copy!(phi, dst)
        return 0;
    }
    if (integ == ORC_RK2) {
        T* pred = buf1; T* corr = buf2;
        if (stage == 1) {
            std::memcpy(pred, phi, tot * sizeof(T));
            std::memcpy(corr, phi, tot * sizeof(T));
            term_sweeps<N, T>(desc, phi, pred, dt, corr, 0.5 * dt, terms, nterms, tc, gs);
            return 0;
        }
        if (stage == 2) {
            term_sweeps<N, T>(desc, pred, corr, 0.5 * dt, nullptr, 0, terms, nterms, tc + dt, gs);
            std::memcpy(phi, corr, tot * sizeof(T));
            return 0;
        }
        return -1;
    }
    if (integ == ORC_RK3) {
        if (stage == 1) {
            std::memcpy(buf1, phi, tot * sizeof(T));
            term_sweeps<N, T>(desc, phi, buf1, dt, nullptr, 0, terms, nterms, tc, gs);
            return 0;
        }
        if (stage == 2) {
            std::memcpy(buf2, phi, tot * sizeof(T));
            for_nodes<N>(geo.n, [&](const Idx& I) {
                const long l = geo.lin(I);
                buf2[l] = T(0.75 * double(phi[l]) + 0.25 * double(buf1[l]));
            });
            term_sweeps<N, T>(desc, buf1, buf2, 0.25 * dt, nullptr, 0, terms, nterms, tc + dt, gs);
            return 0;
        }
        if (stage == 3) {
            std::memcpy(buf1, phi, tot * sizeof(T));
            for_nodes<N>(geo.n, [&](const Idx& I) {
                const long l = geo.lin(I);
                buf1[l] = T((phi[l] + T(2) * buf2[l]) / T(3));      // all in V: (phi + 2*buf2)/3
            });
            term_sweeps<N, T>(desc, buf2, buf1, (2.0 / 3.0) * dt, nullptr, 0, terms, nterms, tc + 0.5 * dt, gs);
            std::memcpy(phi, buf1, tot * sizeof(T));
            return 0;
        }
        return -1;
    }
    return -1;
}

template <int N, class T>
int cfl_impl(const orc_field& f, const orc_term* terms, int nterms, double t, const double* gs, double* dt_out) {
    Field<N, T> F(f);
    double dt = std::numeric_limits<double>::infinity();
    for (int k = 0; k < nterms; ++k) {
        double d = cfl_term<N, T>(F, terms[k], tscale(terms[k], t, gs, k));
        dt = (k == 0) ? d : jl_min(dt, d);
    }
    *dt_out = dt;
    return (dt > 0) ? 0 : 1;    // levelsetterms.jl:26 : Δt > 0 || throw(ArgumentError)
}

inline double jl_eps(double x) {   // Base.eps(::Float64)
    x = std::fabs(x);
    return std::nextafter(x, std::numeric_limits<double>::infinity()) - x;
}

int nstages(int integ) { return integ == ORC_FE ? 1 : integ == ORC_RK2 ? 2 : 3; }

// timestepping.jl:101-122 with default (no-op) hooks
template <int N, class T>
int integrate_impl(const orc_field& desc, int integ, double alpha, T* phi, const orc_term* terms, int nterms,
                   double t0, double tf, double dt_max, int64_t max_steps, double* t_out, int64_t* steps_out) {
    if (!(tf >= t0)) return 2;   // levelsetequation.jl:196
    const long tot = total_elems<N, T>(desc);
    std::vector<T> b1(phi, phi + tot), b2(phi, phi + tot);     // _alloc_buffers: copies of phi
    double tc = t0;
    int64_t steps = 0;
    orc_field d = desc;
    while (tc <= tf - jl_eps(tc)) {
        if (max_steps >= 0 && steps >= max_steps) { *t_out = tc; *steps_out = steps; return 0; }
        d.vals = phi;
        double cfl;
        if (cfl_impl<N, T>(d, terms, nterms, tc, nullptr, &cfl)) { *t_out = tc; *steps_out = steps; return 1; }
        double dt = jl_min(jl_min(dt_max, alpha * cfl), tf - tc);
        for (int s = 1; s <= nstages(integ); ++s)
            stage_impl<N, T>(desc, integ, s, phi, b1.data(), b2.data(), terms, nterms, tc, dt, nullptr);
        tc += dt;
        ++steps;
    }
    *t_out = tf;                 // "land on tf exactly" (:120)
    *steps_out = steps;
    return 0;
}

// levelsetops.jl smooth_heaviside / smooth_delta (:186-195 region)
inline double smooth_heaviside(double x, double a) {
    if (x > a) return 1.0;
    if (x < -a) return 0.0;
    return 0.5 * (1.0 + x / a + 1.0 / M_PI * std::sin(M_PI * x / a));
}
inline double smooth_delta(double x, double a) {
    return std::fabs(x) > a ? 0.0 : 0.5 / a * (1.0 + std::cos(M_PI * x / a));
}

// Base.mapreduce_impl pairwise summation (block 1024) — what sum(f, ::Array) does in Julia.
template <class F> double pairwise_sum(F&& f, long first, long last) {
    if (first == last) return f(first);
    if (last - first < 1024) {
        double v = f(first) + f(first + 1);
        for (long i = first + 2; i <= last; ++i) v += f(i);
        return v;
    }
    long mid = first + ((last - first) >> 1);
    double v1 = pairwise_sum(f, first, mid);
    double v2 = pairwise_sum(f, mid + 1, last);
    return v1 + v2;
}

// velocityextension.jl:20-69, 95-116
template <int N, class T>
void extend_impl(const orc_field& pf, T* Fv, int nb_iters, double cfl, const uint8_t* frozen, double band, double min_norm) {
    orc_field d = pf;
    bool any = false;
    for (int k = 0; k < N; ++k) if (d.bc[k][0].kind != ORC_BC_NONE) any = true;
    if (!any) for (int k = 0; k < N; ++k) { d.bc[k][0] = {ORC_BC_EXTRAP, 1}; d.bc[k][1] = {ORC_BC_EXTRAP, 1}; }
    Field<N, T> phi(d);
    long tot = 1; for (int k = 0; k < N; ++k) tot *= phi.n[k];
    double dx = phi.h[0]; for (int k = 1; k < N; ++k) dx = std::min(dx, phi.h[k]);
    const double tau = cfl * dx;
    std::vector<uint8_t> mask(tot);
    std::vector<T> comp[3];
    for (int k = 0; k < N; ++k) comp[k].assign(tot, T(0));
    const double mn2 = min_norm * min_norm;
    const int n1 = phi.n[0], n2 = N > 1 ? phi.n[1] : 1, n3 = N > 2 ? phi.n[2] : 1;
    long l = 0;
    for (int k3 = 1; k3 <= n3; ++k3) for (int k2 = 1; k2 <= n2; ++k2) for (int k1 = 1; k1 <= n1; ++k1, ++l) {
        Idx I{{k1, k2, k3}};
        const T p = phi.raw(I);
        mask[l] = frozen ? frozen[l] : (std::fabs(double(p)) <= band * dx);
        T g[3]; T nrm2 = T(0);
        for (int dd = 0; dd < N; ++dd) { g[dd] = T(D0(phi, I, dd)); nrm2 = dd == 0 ? T(g[dd] * g[dd]) : T(nrm2 + T(g[dd] * g[dd])); }
        if (double(nrm2) <= mn2) continue;
        const T invn = T(1) / T(std::sqrt(nrm2));
        const double S = double(p) / std::sqrt(double(T(p * p)) + dx * dx);
        for (int dd = 0; dd < N; ++dd) comp[dd][l] = T(S * double(g[dd]) * double(invn));
    }
    orc_field fd = d; fd.vals = Fv;
    std::vector<T> Fnew(tot);
    for (int it = 0; it < nb_iters; ++it) {
        Field<N, T> Fw(fd);
        l = 0;
        for (int k3 = 1; k3 <= n3; ++k3) for (int k2 = 1; k2 <= n2; ++k2) for (int k1 = 1; k1 <= n1; ++k1, ++l) {
            Idx I{{k1, k2, k3}};
            if (mask[l]) { Fnew[l] = Fw.raw(I); continue; }
            double adv = 0.0;      // zero(eltype(F_new)) + a*dF promotes to Float64 at the first += (dF is Float64)
            for (int dd = 0; dd < N; ++dd) {
                const T a = comp[dd][l];
                const double dF = a > T(0) ? Dm(Fw, I, dd) : Dp(Fw, I, dd);
                adv = adv + double(a) * dF;
            }
            Fnew[l] = T(double(Fw.raw(I)) - tau * adv);
        }
        std::memcpy(Fv, Fnew.data(), tot * sizeof(T));
    }
}

#define DISPATCH(f, ...)                                                      \
    do {                                                                      \
        const int nd_ = (f)->ndim; const bool f32_ = (f)->dtype == ORC_F32;   \
        if (nd_ == 1) { if (f32_) { constexpr int N = 1; using T = float; __VA_ARGS__; } else { constexpr int N = 1; using T = double; __VA_ARGS__; } } \
        else if (nd_ == 2) { if (f32_) { constexpr int N = 2; using T = float; __VA_ARGS__; } else { constexpr int N = 2; using T = double; __VA_ARGS__; } } \
        else { if (f32_) { constexpr int N = 3; using T = float; __VA_ARGS__; } else { constexpr int N = 3; using T = double; __VA_ARGS__; } } \
    } while (0)

inline Idx mkidx(const int32_t* I, int nd) {
    Idx J{{1, 1, 1}};
    for (int d = 0; d < nd; ++d) J.i[d] = I[d];
    return J;
}

}  // namespace

namespace {
template <class T> T jl_min(T a, T b) { if (a != a || b != b) return a + b; if (a == b) return std::signbit(a) ? a : b; return a < b ? a : b; }   // Base.min
template <class T> T jl_max(T a, T b) { if (a != a || b != b) return a + b; if (a == b) return std::signbit(a) ? b : a; return a > b ? a : b; }   // Base.max
template <class T> void csg_impl(T* d, const T* s, int64_t n, int op) {
    for (int64_t i = 0; i < n; ++i) {
        if (op == 0) d[i] = jl_min<T>(d[i], s[i]);            // levelsetops.jl:257
        else if (op == 1) d[i] = jl_max<T>(d[i], s[i]);       // :277
        else if (op == 2) d[i] = jl_max<T>(d[i], -s[i]);      // :315
        else d[i] = -d[i];                                    // :296
    }
}
}  // namespace

extern "C" {

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

double orc_meshsize(const orc_field* f, int dim) {
    int d = dim - 1;
    int ng = f->nglob[d] > 0 ? f->nglob[d] : f->n[d];
    return (f->hc[d] - f->lc[d]) / double(ng - 1);
}

// meshes.jl:114-117 : lc .+ (I .- 1) .* h
void orc_getnode(const orc_field* f, const int32_t* I, double* x) {
    for (int d = 0; d < f->ndim; ++d) {
        double h = orc_meshsize(f, d + 1);
        x[d] = f->lc[d] + double(I[d] + f->off[d] - 1) * h;
    }
}

double orc_getindex(const orc_field* f, const int32_t* I) {
    double r = 0;
    DISPATCH(f, { Field<N, T> F(*f); r = double(getindex<N, T>(F, mkidx(I, N))); });
    return r;
}

double orc_weno5(double a, double b, double c, double d, double e) { return weno5(a, b, c, d, e); }
double orc_limiter(double x, double y) { return limiter(x, y); }

double orc_deriv(const orc_field* f, int op, const int32_t* I, int dim, int dim2) {
    double r = 0;
    const int d = dim - 1, d2 = dim2 - 1;
    DISPATCH(f, {
        Field<N, T> F(*f); Idx J = mkidx(I, N);
        switch (op) {
            case 0: r = D0(F, J, d); break;
            case 1: r = Dp(F, J, d); break;
            case 2: r = Dm(F, J, d); break;
            case 3: r = weno5m(F, J, d); break;
            case 4: r = weno5p(F, J, d); break;
            case 5: r = D20(F, J, d); break;
            case 6: r = D2pp(F, J, d); break;
            case 7: r = D2mm(F, J, d); break;
            case 8: r = D2mixed(F, J, d, d2); break;
            default: r = std::numeric_limits<double>::quiet_NaN();
        }
    });
    return r;
}

double orc_curvature(const orc_field* f, const int32_t* I) {
    double r = 0;
    DISPATCH(f, { Field<N, T> F(*f); r = curvature(F, mkidx(I, N)); });
    return r;
}

// levelsetops.jl:27-33
double orc_volume(const orc_field* f) {
    double r = 0;
    DISPATCH(f, {
        Field<N, T> F(*f);
        double dmin = F.h[0], vol = F.h[0];
        for (int d = 1; d < N; ++d) { dmin = std::min(dmin, F.h[d]); vol *= F.h[d]; }
        long tot = 1; for (int d = 0; d < N; ++d) tot *= F.n[d];
        const T* v = F.v;
        r = vol * pairwise_sum([&](long i) { return smooth_heaviside(-double(v[i]), dmin); }, 0, tot - 1);
    });
    return r;
}

// levelsetops.jl:139-149 : sequential (mapfoldl over CartesianIndices), default LinearExtrapolationBC if none
double orc_perimeter(const orc_field* f0) {
    orc_field f = *f0;
    bool any = false;
    for (int d = 0; d < f.ndim; ++d) if (f.bc[d][0].kind != ORC_BC_NONE) any = true;
    if (!any) for (int d = 0; d < f.ndim; ++d) { f.bc[d][0] = {ORC_BC_EXTRAP, 1}; f.bc[d][1] = {ORC_BC_EXTRAP, 1}; }
    double r = 0;
    DISPATCH(&f, {
        Field<N, T> F(f);
        double dmin = F.h[0], vol = F.h[0];
        for (int d = 1; d < N; ++d) { dmin = std::min(dmin, F.h[d]); vol *= F.h[d]; }
        double s = 0; bool first = true;
        const int n1 = F.n[0], n2 = N > 1 ? F.n[1] : 1, n3 = N > 2 ? F.n[2] : 1;
        for (int k = 1; k <= n3; ++k) for (int j = 1; j <= n2; ++j) for (int i = 1; i <= n1; ++i) {
            Idx I{{i, j, k}};
            double g2 = 0;     // norm(gradient): sqrt(sum of squares) for a short SVector
            for (int d = 0; d < N; ++d) { double q = D0(F, I, d); g2 = (d == 0) ? q * q : g2 + q * q; }
            double term = smooth_delta(double(F.raw(I)), dmin) * std::sqrt(g2);
            if (first) { s = term; first = false; } else s += term;
        }
        r = vol * s;
    });
    return r;
}

double orc_tscale(const orc_term* term, double t) { return tscale(*term, t, nullptr, 0); }

double orc_compute_term(const orc_field* phi, const orc_term* term, const int32_t* I, double t, double gscale) {
    double r = 0;
    double g = term->tscale_kind == ORC_TS_HOST ? gscale : tscale(*term, t, nullptr, 0);
    DISPATCH(phi, { Field<N, T> F(*phi); r = compute_term(F, *term, mkidx(I, N), g); });
    return r;
}

double orc_compute_cfl_term(const orc_field* phi, const orc_term* term, double t, double gscale) {
    double r = 0;
    double g = term->tscale_kind == ORC_TS_HOST ? gscale : tscale(*term, t, nullptr, 0);
    DISPATCH(phi, { Field<N, T> F(*phi); r = cfl_term<N, T>(F, *term, g); });
    return r;
}

int orc_compute_cfl(const orc_field* phi, const orc_term* terms, int nterms, double t, const double* gscale, double* dt_out) {
    int rc = 0;
    DISPATCH(phi, { rc = cfl_impl<N, T>(*phi, terms, nterms, t, gscale, dt_out); });
    return rc;
}

void orc_eikonal_s0(const orc_field* phi0, double* out) {
    DISPATCH(phi0, {
        Field<N, T> F(*phi0);
        double dx = F.h[0];
        for (int d = 1; d < N; ++d) dx = std::min(dx, F.h[d]);
        long l = 0;
        const int n1 = F.n[0], n2 = N > 1 ? F.n[1] : 1, n3 = N > 2 ? F.n[2] : 1;
        for (int k = 1; k <= n3; ++k) for (int j = 1; j <= n2; ++j) for (int i = 1; i <= n1; ++i) {
            Idx I{{i, j, k}};
            T v = F.raw(I);
            out[l++] = double(v) / std::sqrt(double(T(v * v)) + dx * dx);     // v / sqrt(v^2 + Δx^2)
        }
    });
}

void orc_csg(int dtype, void* dst, const void* src, int64_t n, int op) {
    if (dtype == ORC_F64) csg_impl<double>(static_cast<double*>(dst), static_cast<const double*>(src), n, op);
    else csg_impl<float>(static_cast<float*>(dst), static_cast<const float*>(src), n, op);
}

int orc_extend_along_normals(const orc_field* phi, void* F, int nb_iters, double cfl, const uint8_t* frozen, double band, double min_norm) {
    if (nb_iters < 0 || !(cfl > 0) || band < 0 || min_norm < 0) return 1;
    DISPATCH(phi, { extend_impl<N, T>(*phi, (T*)F, nb_iters, cfl, frozen, band, min_norm); });
    return 0;
}

int orc_nstages(int integ) { return nstages(integ); }

int orc_stage(const orc_field* desc, int integ, int stage, void* phi, void* b1, void* b2,
              const orc_term* terms, int nterms, double tc, double dt, const double* gs) {
    int rc = 0;
    DISPATCH(desc, { rc = stage_impl<N, T>(*desc, integ, stage, (T*)phi, (T*)b1, (T*)b2, terms, nterms, tc, dt, gs); });
    return rc;
}

int orc_advance(const orc_field* desc, int integ, void* phi, void* b1, void* b2,
                const orc_term* terms, int nterms, double tc, double dt) {
    for (int s = 1; s <= nstages(integ); ++s) {
        int rc = orc_stage(desc, integ, s, phi, b1, b2, terms, nterms, tc, dt, nullptr);
        if (rc) return rc;
    }
    return 0;
}

int orc_integrate(const orc_field* desc, int integ, double cfl, void* phi, const orc_term* terms, int nterms,
                  double t0, double tf, double dt_max, int64_t max_steps, double* t_out, int64_t* steps_out) {
    int rc = 0;
    DISPATCH(desc, { rc = integrate_impl<N, T>(*desc, integ, cfl, (T*)phi, terms, nterms, t0, tf, dt_max, max_steps, t_out, steps_out); });
    return rc;
}

}  // extern "C"
