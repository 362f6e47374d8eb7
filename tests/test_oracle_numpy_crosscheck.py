"""An independent, vectorised NumPy restatement of the reference's periodic WENO5 advection + TVD-RK3 step, checked against
the C++ oracle.  The two restatements share no code: this one is written from the reference text (src/derivatives.jl:28-121,
src/levelsetterms.jl:73-96, src/timestepping.jl:170-202, src/boundaryconditions.jl:107-119) with whole-array operations, the
oracle with per-node loops — a transcription slip in either shows up as a difference far above rounding."""
import numpy as np
import pytest


def _pad_periodic(a, axis, g=3):
    """Ghosts of a periodic field whose node n duplicates node 1 (boundaryconditions.jl:107-119): i<1 -> n-(1-i), i>n -> 1+(i-n)."""
    n = a.shape[axis]
    idx = np.arange(-g, n + g)                       # 0-based node index incl. ghosts
    src = np.where(idx < 0, idx + n - 1, np.where(idx >= n, idx - n + 1, idx))
    return np.take(a, src, axis=axis)


def _weno5(v1, v2, v3, v4, v5):                       # derivatives.jl:61-81, operation for operation
    d1 = (1 / 3) * v1 - (7 / 6) * v2 + (11 / 6) * v3
    d2 = -(1 / 6) * v2 + (5 / 6) * v3 + (1 / 3) * v4
    d3 = (1 / 3) * v3 + (5 / 6) * v4 - (1 / 6) * v5
    S1 = (13 / 12) * (v1 - 2 * v2 + v3) ** 2 + (1 / 4) * (v1 - 4 * v2 + 3 * v3) ** 2
    S2 = (13 / 12) * (v2 - 2 * v3 + v4) ** 2 + (1 / 4) * (v2 - v4) ** 2
    S3 = (13 / 12) * (v3 - 2 * v4 + v5) ** 2 + (1 / 4) * (3 * v3 - 4 * v4 + v5) ** 2
    eps = 1.0e-6 * np.maximum.reduce([v1 ** 2, v2 ** 2, v3 ** 2, v4 ** 2, v5 ** 2]) + 1.0e-99
    a1, a2, a3 = 0.1 / (S1 + eps) ** 2, 0.6 / (S2 + eps) ** 2, 0.3 / (S3 + eps) ** 2
    s = a1 + a2 + a3
    return (a1 / s) * d1 + (a2 / s) * d2 + (a3 / s) * d3


def _advection(phi, u, h):
    """sum_d u_d * (u_d > 0 ? weno5-(d) : weno5+(d)) on a periodic grid (levelsetterms.jl:73-82)."""
    H = np.zeros_like(phi)
    for d in range(phi.ndim):
        n = phi.shape[d]
        P = _pad_periodic(phi, d)

        def at(k):                                    # phi[I + k e_d] for every node
            return np.take(P, np.arange(3 + k, 3 + k + n), axis=d)

        Dm = lambda k: (at(k) - at(k - 1)) / h[d]     # D-(phi, I + k e_d)   (derivatives.jl:52-57)
        Dp = lambda k: (at(k + 1) - at(k)) / h[d]     # D+(phi, I + k e_d)   (derivatives.jl:39-44)
        wm = _weno5(Dm(-2), Dm(-1), Dm(0), Dm(1), Dm(2))            # weno5-  (:89-101)
        wp = _weno5(Dp(2), Dp(1), Dp(0), Dp(-1), Dp(-2))            # weno5+  (:109-121)
        term = u[d] * np.where(u[d] > 0, wm, wp)
        H = term if d == 0 else H + term
    return H


def _rk3(phi, u, h, dt):                              # timestepping.jl:170-202
    b1 = phi - dt * _advection(phi, u, h)
    b2 = 0.75 * phi + 0.25 * b1
    b2 = b2 - 0.25 * dt * _advection(b1, u, h)
    b3 = (phi + 2 * b2) / 3
    return b3 - (2 / 3) * dt * _advection(b2, u, h)


@pytest.mark.parametrize("n", [(48, 40), (20, 18, 16)])
def test_numpy_restatement_agrees_with_oracle(O, n):
    N = len(n)
    lc, hc = (-1.0,) * N, (1.0,) * N
    f = O.Field(np.zeros(n, order="F"), lc, hc, bc=O.PERIODIC)
    X = f.nodes()
    r = np.sqrt(sum((x - 0.1 * (i + 1)) ** 2 for i, x in enumerate(X)))
    f.vals[...] = np.broadcast_to(r - 0.45, n)
    if N == 2:
        u = np.stack([np.broadcast_to(-X[1], n), np.broadcast_to(X[0], n)])
    else:
        u = np.stack([np.broadcast_to(-X[1] + 0.3, n), np.broadcast_to(X[0] * X[2], n), np.broadcast_to(0.5 - X[0], n)])
    u = np.asfortranarray(u)
    h = [f.meshsize(d + 1) for d in range(N)]
    terms = [O.advection(u)]
    dt = 0.5 * O.compute_cfl(f, terms, 0.0)
    assert dt == 0.5 / np.max(sum(np.abs(u[d]) / h[d] for d in range(N)))          # levelsetterms.jl:90-96, bit for bit
    phi = f.vals.copy()
    for _ in range(5):
        O.advance(f, O.RK3, terms, 0.0, dt)
        phi = _rk3(phi, u, h, dt)
    assert np.abs(phi - f.vals).max() <= 5e-15, np.abs(phi - f.vals).max()


# ---- Godunov / ENO2 terms and curvature (levelsetterms.jl:111-121,156-187,234-265; derivatives.jl:28-57,129-175;
# ---- levelsetops.jl:197-244), again with whole-array operations on a periodic grid ----------------------------------
class _Periodic:
    def __init__(self, phi, h):
        self.n, self.h, self.N = phi.shape, h, phi.ndim
        P = phi
        for d in range(phi.ndim):
            P = _pad_periodic(P, d)
        self.P = P

    def at(self, *off):
        """phi[I + off] for every node I (ghosts resolved dimension by dimension, so corners compose)."""
        off = tuple(off) + (0,) * (self.N - len(off))
        return self.P[tuple(slice(3 + o, 3 + o + n) for o, n in zip(off, self.n))]

    def e(self, d, k=1):
        return tuple(k if i == d else 0 for i in range(self.N))

    def D0(self, d, base=None):
        b = base or (0,) * self.N
        p = tuple(x + y for x, y in zip(b, self.e(d)))
        m = tuple(x - y for x, y in zip(b, self.e(d)))
        return (self.at(*p) - self.at(*m)) / (2 * self.h[d])


def _limiter(x, y):                                   # levelsetterms.jl:184-187
    return np.where(x * y > 0, np.where(np.abs(x) <= np.abs(y), x, y), 0.0)


def _pos(x):
    return np.where(x > 0, x, 0.0)


def _neg(x):
    return np.where(x < 0, x, 0.0)


def _eno_pair(G, d):
    h = G.h[d]
    c, p1, p2, m1, m2 = G.at(), G.at(*G.e(d)), G.at(*G.e(d, 2)), G.at(*G.e(d, -1)), G.at(*G.e(d, -2))
    D2c = (p1 - 2 * c + m1) / h ** 2
    D2mm = (m2 - 2 * m1 + c) / h ** 2
    D2pp = (c - 2 * p1 + p2) / h ** 2
    neg = (c - m1) / h + (0.5 * h) * _limiter(D2mm, D2c)
    pos = (p1 - c) / h - (0.5 * h) * _limiter(D2pp, D2c)
    return neg, pos


def _normal_motion(phi, v, h):
    G = _Periodic(phi, h)
    gp = gm = None
    for d in range(phi.ndim):
        neg, pos = _eno_pair(G, d)
        a, b = _pos(neg) ** 2 + _neg(pos) ** 2, _neg(neg) ** 2 + _pos(pos) ** 2
        gp, gm = (a, b) if d == 0 else (gp + a, gm + b)
    return _pos(v) * np.sqrt(gp) + _neg(v) * np.sqrt(gm)


def _eikonal_frozen(phi, S0, h):
    G = _Periodic(phi, h)
    A2 = B2 = None
    up = np.sign(S0) > 0
    for d in range(phi.ndim):
        A, B = _eno_pair(G, d)
        a = np.where(up, _pos(A) ** 2, _neg(A) ** 2)
        b = np.where(up, _neg(B) ** 2, _pos(B) ** 2)
        A2, B2 = (a, b) if d == 0 else (A2 + a, B2 + b)
    return S0 * (np.sqrt(A2 + B2) - 1)


def _curvature_term(phi, b, h):
    return _curvature_term_padded(_Periodic(phi, h), b, h)


@pytest.mark.parametrize("n", [(40, 36), (18, 16, 14)])
def test_numpy_godunov_and_curvature_agree_with_oracle(O, n):
    N = len(n)
    lc, hc = (-1.0,) * N, (1.0,) * N
    f = O.Field(np.zeros(n, order="F"), lc, hc, bc=O.PERIODIC)
    X = f.nodes()
    r = np.sqrt(sum((x - 0.05 * (i + 1)) ** 2 for i, x in enumerate(X)))
    phi0 = np.asfortranarray(np.broadcast_to((r - 0.5) * (1 + 0.3 * np.sin(3 * X[0])), n).copy())
    h = [f.meshsize(d + 1) for d in range(N)]
    v = np.asfortranarray(np.broadcast_to(0.3 + 0.5 * np.cos(2 * X[0]) - 0.4 * (X[1] > 0), n).copy())      # both signs
    dt = 0.2 * min(h)
    # forward-Euler steps of each term: oracle (per-node loops) vs NumPy (whole arrays)
    for name, term, fn, tol in (("normal", O.normal_motion(v), lambda p: _normal_motion(p, v, h), 2e-15),
                                ("curvature", O.curvature(-0.05), lambda p: _curvature_term(p, -0.05, h), 2e-15)):
        f.vals[...] = phi0
        phi = phi0.copy()
        for _ in range(3):
            O.advance(f, O.FE, [term], 0.0, dt if name == "normal" else 0.1 * min(h) ** 2)
            phi = phi - (dt if name == "normal" else 0.1 * min(h) ** 2) * fn(phi)
        assert np.abs(phi - f.vals).max() <= tol, (name, np.abs(phi - f.vals).max())
    # Eikonal, frozen sign (O&F 7.5)
    f.vals[...] = phi0
    S0 = phi0 / np.sqrt(phi0 ** 2 + min(h) ** 2)
    assert np.array_equal(O.eikonal_s0(f).reshape(n, order="F"), S0)                 # levelsetterms.jl:217-221, bit for bit
    phi = phi0.copy()
    term = O.eikonal(O.eikonal_s0(f))
    for _ in range(3):
        O.advance(f, O.FE, [term], 0.0, 0.5 * min(h))
        phi = phi - (0.5 * min(h)) * _eikonal_frozen(phi, S0, h)
    assert np.abs(phi - f.vals).max() <= 2e-15


# ---- ghost cells of ExtrapolationBC{P} / SymmetryBC with corner composition (boundaryconditions.jl:90-97,132-153;
# ---- meshfield.jl:248-260): padding axis by axis, lowest dimension first, equals the reference's dim N -> 1 recursion ----
def _w(j, k, P):
    w = 1.0
    for mm in range(P + 1):
        if mm != j:
            w *= (-k - mm) / (j - mm)
    return w


def _pad_bc(a, axis, kind, P=0, g=3):
    n = a.shape[axis]
    a = np.moveaxis(a, axis, 0)
    lo, hi = [], []
    for k in range(g, 0, -1):                         # ghosts at distance k below node 1 (stored from far to near)
        if kind == "sym":
            lo.append(a[k])
        else:
            acc = np.zeros_like(a[0])
            for j in range(P + 1):
                acc = acc + _w(j, k, P) * a[j]
            lo.append(acc)
    for k in range(1, g + 1):
        if kind == "sym":
            hi.append(a[n - 1 - k])
        else:
            acc = np.zeros_like(a[0])
            for j in range(P + 1):
                acc = acc + _w(j, k, P) * a[n - 1 - j]
            hi.append(acc)
    return np.moveaxis(np.concatenate([np.stack(lo), a, np.stack(hi)], axis=0), 0, axis)


class _Padded(_Periodic):
    def __init__(self, phi, h, bcs):                  # bcs[d] = (kind, P)
        self.n, self.h, self.N = phi.shape, h, phi.ndim
        P = phi
        for d in range(phi.ndim):                     # dimension 1 first: the recursion resolves dim N outermost, dim 1 innermost
            P = _pad_bc(P, d, *bcs[d])
        self.P = P


@pytest.mark.parametrize("n", [(30, 26), (14, 13, 12)])
def test_numpy_ghost_composition_agrees_with_oracle(O, n):
    """Every ghost the stencils can touch (3 deep, corners and edges included) for Neumann, linear / quadratic extrapolation and
    symmetry, then one curvature + upwind step that reads them."""
    N = len(n)
    rng = np.random.default_rng(5)
    phi0 = np.asfortranarray(rng.standard_normal(n))
    kinds = [("extrap", 0), ("extrap", 2), ("sym", 0)][:N] if N == 3 else [("extrap", 1), ("sym", 0)]
    obc = [O.EXTRAP(p) if k == "extrap" else O.SYMMETRY for k, p in kinds]
    f = O.Field(phi0.copy(order="F"), (-1.0,) * N, (1.0,) * N, bc=obc)
    h = [f.meshsize(d + 1) for d in range(N)]
    G = _Padded(phi0, h, kinds)
    worst = 0.0
    import itertools
    for off in itertools.product(*[range(-3, m + 3) for m in n]):
        if all(0 <= o < m for o, m in zip(off, n)):
            continue
        if sum(1 for o, m in zip(off, n) if not 0 <= o < m) < 2 and rng.random() > 0.1:
            continue                                   # all corners / edges, a sample of the faces
        ref = f[tuple(o + 1 for o in off)]
        worst = max(worst, abs(ref - G.P[tuple(o + 3 for o in off)]))
    assert worst <= 1e-12, worst
    phi = phi0 - 1e-4 * _curvature_term_padded(G, -0.05, h)
    O.advance(f, O.FE, [O.curvature(-0.05)], 0.0, 1e-4)
    assert np.abs(phi - f.vals).max() <= 1e-12


def _curvature_term_padded(G, b, h):
    N = G.N
    g = [G.D0(d) for d in range(N)]
    nrmsq = g[0] * g[0]
    for d in range(1, N):
        nrmsq = nrmsq + g[d] * g[d]
    Hm = [[None] * N for _ in range(N)]
    for i in range(N):
        Hm[i][i] = (G.at(*G.e(i)) - 2 * G.at() + G.at(*G.e(i, -1))) / h[i] ** 2
        for j in range(i + 1, N):
            Hm[i][j] = Hm[j][i] = (G.D0(j, G.e(i)) - G.D0(j, G.e(i, -1))) / (2 * h[i])
    tr = Hm[0][0]
    for d in range(1, N):
        tr = tr + Hm[d][d]
    quad = None
    for j in range(N):
        col = g[0] * Hm[0][j]
        for i in range(1, N):
            col = col + g[i] * Hm[i][j]
        quad = col * g[j] if j == 0 else quad + col * g[j]
    with np.errstate(divide="ignore", invalid="ignore"):
        kappa = np.where(nrmsq < np.finfo(np.float64).eps, 0.0, (tr * nrmsq - quad) / nrmsq ** 1.5)
    return b * kappa * np.sqrt(nrmsq)


# ---- Float32 fields: Julia's promotion rules (the reference never tests them; SURVEY.md §8a "precision notes") ----------
def _advection_f32(phi, u, h):
    """phi, u Float32.  The first difference is Float32 (V - V), the division by the Float64 meshsize promotes, everything
    after is Float64 (all literals are Float64); u_d * weno is Float32 * Float64 -> Float64."""
    H = None
    for d in range(phi.ndim):
        n = phi.shape[d]
        P = _pad_periodic(phi, d)
        at = lambda k: np.take(P, np.arange(3 + k, 3 + k + n), axis=d)
        Dm = lambda k: (at(k) - at(k - 1)).astype(np.float64) / h[d]
        Dp = lambda k: (at(k + 1) - at(k)).astype(np.float64) / h[d]
        wm = _weno5(Dm(-2), Dm(-1), Dm(0), Dm(1), Dm(2))
        wp = _weno5(Dp(2), Dp(1), Dp(0), Dp(-1), Dp(-2))
        term = u[d].astype(np.float64) * np.where(u[d] > 0, wm, wp)
        H = term if d == 0 else H + term
    return H


def _rk3_f32(phi, u, h, dt):
    f32, f64 = np.float32, np.float64
    b1 = (phi.astype(f64) - dt * _advection_f32(phi, u, h)).astype(f32)                  # buf1[I] -= dt * H   (store rounds)
    b2 = (0.75 * phi.astype(f64) + 0.25 * b1.astype(f64)).astype(f32)                    # Float64 literals promote
    b2 = (b2.astype(f64) - 0.25 * dt * _advection_f32(b1, u, h)).astype(f32)
    b3 = ((phi + f32(2) * b2) / f32(3)).astype(f32)                                      # (phi + 2 buf2) / 3 stays in Float32
    return (b3.astype(f64) - (2 / 3) * dt * _advection_f32(b2, u, h)).astype(f32)


def test_numpy_float32_promotion_agrees_with_oracle(O):
    n = (40, 34)
    f = O.Field(np.zeros(n, dtype=np.float32, order="F"), (-1.0, -1.0), (1.0, 1.0), bc=O.PERIODIC)
    X = f.nodes()
    f.vals[...] = np.broadcast_to(np.hypot(X[0] - 0.3, X[1]) - 0.4, n).astype(np.float32)
    u = np.asfortranarray(np.stack([np.broadcast_to(-X[1], n), np.broadcast_to(X[0], n)]).astype(np.float32))
    h = [f.meshsize(1), f.meshsize(2)]
    terms = [O.advection(u)]
    dt = 0.5 * O.compute_cfl(f, terms, 0.0)
    phi = f.vals.copy()
    for _ in range(5):
        O.advance(f, O.RK3, terms, 0.0, dt)
        phi = _rk3_f32(phi, u, h, dt)
    assert phi.dtype == np.float32 and f.vals.dtype == np.float32
    assert np.array_equal(phi, f.vals)          # same IEEE operations in the same order: bit-identical


# ---- extend_along_normals! (velocityextension.jl:20-116), whole-array restatement -------------------------------------
def _extend(F, phi, h, bcs, nb_iters, cfl=0.45, frozen=None, band=1.5, min_norm=1e-14):
    N = phi.ndim
    D = min(h)
    tau = cfl * D
    Gp = _Padded(phi, h, bcs)
    frozen = np.abs(phi) <= band * D if frozen is None else frozen
    g = [Gp.D0(d) for d in range(N)]
    n2 = g[0] ** 2
    for d in range(1, N):
        n2 = n2 + g[d] ** 2
    ok = n2 > min_norm ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / np.sqrt(n2)
        S = phi / np.sqrt(phi ** 2 + D ** 2)
        a = [np.where(ok, (S * g[d]) * inv, 0.0) for d in range(N)]
    F = F.copy()
    for _ in range(nb_iters):
        G = _Padded(F, h, bcs)
        adv = np.zeros_like(F)
        for d in range(N):
            dm = (G.at() - G.at(*G.e(d, -1))) / h[d]
            dp = (G.at(*G.e(d)) - G.at()) / h[d]
            adv = adv + a[d] * np.where(a[d] > 0, dm, dp)
        F = np.where(frozen, F, F - tau * adv)
    return F


@pytest.mark.parametrize("n,kinds", [((41, 37), [("extrap", 1), ("extrap", 1)]), ((41, 37), [("extrap", 0), ("sym", 0)]),
                                     ((18, 17, 16), [("extrap", 1)] * 3)])
def test_numpy_velocity_extension_agrees_with_oracle(O, n, kinds):
    N = len(n)
    default_bc = all(k == ("extrap", 1) for k in kinds)          # a field without BCs gets LinearExtrapolationBC (:38-43)
    obc = None if default_bc else [O.EXTRAP(p) if k == "extrap" else O.SYMMETRY for k, p in kinds]
    f = O.Field(np.zeros(n, order="F"), (-1.0,) * N, (1.0,) * N, bc=obc)
    X = f.nodes()
    r = np.sqrt(sum(x * x for x in X))
    f.vals[...] = np.broadcast_to((r - 0.5) * (1 + 0.2 * X[0]), n)
    F0 = np.asfortranarray(np.broadcast_to(np.sin(3 * X[0]) + X[1] ** 2, n).copy())
    h = [f.meshsize(d + 1) for d in range(N)]
    for frozen in (None, np.asfortranarray(np.abs(f.vals) < 0.07)):
        ref = O.extend_along_normals(F0.copy(order="F"), f, nb_iters=9, frozen=frozen)
        mine = _extend(F0, f.vals, h, kinds, 9, frozen=frozen)
        assert np.abs(ref - mine).max() <= 1e-14, np.abs(ref - mine).max()


# ---- RK2 (Heun, timestepping.jl:143-164) and a two-term equation with the per-term sequential subtraction ---------------
def test_numpy_rk2_two_terms_agrees_with_oracle(O):
    n = (36, 30)
    f = O.Field(np.zeros(n, order="F"), (-1.0, -1.0), (1.0, 1.0), bc=O.PERIODIC)
    X = f.nodes()
    f.vals[...] = np.broadcast_to((np.hypot(X[0] - 0.1, X[1]) - 0.5) * (1 + 0.3 * np.sin(3 * X[0])), n)
    h = [f.meshsize(1), f.meshsize(2)]
    u = np.asfortranarray(np.stack([np.broadcast_to(-X[1], n), np.broadcast_to(X[0], n)]))
    v = np.asfortranarray(np.broadcast_to(0.2 + 0.3 * np.cos(2 * X[1]), n).copy())
    terms = [O.normal_motion(v), O.advection(u)]
    L = [lambda p: _normal_motion(p, v, h), lambda p: _advection(p, u, h)]
    dt = 0.5 * O.compute_cfl(f, terms, 0.0)
    # compute_cfl: minimum over terms of the per-term node minimum (levelsetterms.jl:22-38)
    assert dt == 0.5 * min(1 / np.max(np.abs(v) / h[0] + np.abs(v) / h[1]), 1 / np.max(np.abs(u[0]) / h[0] + np.abs(u[1]) / h[1]))
    phi = f.vals.copy()
    for _ in range(4):
        O.advance(f, O.RK2, terms, 0.0, dt)
        pred, corr = phi.copy(), phi.copy()
        for Lk in L:                                   # pred -= dt*v ; corr -= 0.5*dt*v, term after term
            val = Lk(phi)
            pred = pred - dt * val
            corr = corr - 0.5 * dt * val
        for Lk in L:
            corr = corr - 0.5 * dt * Lk(pred)
        phi = corr
    assert np.abs(phi - f.vals).max() <= 2e-15, np.abs(phi - f.vals).max()
