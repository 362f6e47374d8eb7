"""Shared builders for the parity tests: the same synthetic, seed-free configuration is materialised
once for the CPU oracle (``oracle.Field`` / ``oracle.Term``) and once for the engine
(``lsm_b200.MeshField`` / terms).  Configurations follow SURVEY.md §8(d) / BASELINE.json ``configs``.
"""
from __future__ import annotations

import numpy as np

import oracle as O


def coords(lc, hc, n):
    out = []
    N = len(n)
    for d in range(N):
        h = (hc[d] - lc[d]) / (n[d] - 1)
        x = lc[d] + np.arange(n[d], dtype=np.float64) * h
        shape = [1] * N
        shape[d] = n[d]
        out.append(x.reshape(shape))
    return out


def bcast(a, n):
    return np.asfortranarray(np.broadcast_to(a, n).copy())


class Case:
    """A named configuration: grid, phi0, list of term specs, BC spec, integrator, dtype."""

    def __init__(self, name, lc, hc, n, phi0, terms, bc, dtype=np.float64):
        self.name, self.lc, self.hc, self.n = name, tuple(lc), tuple(hc), tuple(n)
        self.phi0 = np.asfortranarray(phi0.astype(dtype))
        self.terms, self.bc, self.dtype = terms, bc, dtype

    # ---- oracle side ----
    def oracle_bc(self):
        def one(b):
            k = b[0]
            return {"periodic": O.PERIODIC, "neumann": O.NEUMANN, "symmetry": O.SYMMETRY}.get(k) or O.EXTRAP(b[1])
        if isinstance(self.bc, tuple) and isinstance(self.bc[0], str):
            return one(self.bc)
        return [one(b) if isinstance(b[0], str) else (one(b[0]), one(b[1])) for b in self.bc]

    def oracle_field(self):
        return O.Field(self.phi0.copy(order="F"), self.lc, self.hc, bc=self.oracle_bc())

    def oracle_terms(self, s0=None):
        out = []
        for t in self.terms:
            k = t["kind"]
            if k == "advection":
                sch = O.UPWIND if t.get("scheme") == "upwind" else O.WENO5
                ts = O.TS_COS if "cos_period" in t else O.TS_NONE
                tp = t.get("cos_period", 1.0)
                if "separable" in t:
                    sc, tabs = t["separable"]
                    out.append(O.advection_separable(sc, tabs, scheme=sch, tscale=ts, tparam=tp))
                elif "field" in t:
                    out.append(O.advection(np.asfortranarray(t["field"].astype(self.dtype)), scheme=sch, tscale=ts, tparam=tp))
                else:
                    out.append(O.advection(t["const"], scheme=sch, tscale=ts, tparam=tp))
            elif k == "normal":
                v = t["field"].astype(self.dtype) if "field" in t else t["const"]
                out.append(O.normal_motion(np.asfortranarray(v) if "field" in t else v))
            elif k == "curvature":
                b = t["field"].astype(self.dtype) if "field" in t else t["const"]
                out.append(O.curvature(np.asfortranarray(b) if "field" in t else b))
            elif k == "eikonal":
                if t.get("frozen", True):
                    out.append(O.eikonal(O.eikonal_s0(self.oracle_field()) if s0 is None else s0))
                else:
                    out.append(O.eikonal())
        return out

    # ---- engine side (lsm_b200 host mirror) ----
    def engine_bc(self, m):
        def one(b):
            k = b[0]
            if k == "periodic":
                return m.PeriodicBC()
            if k == "neumann":
                return m.NeumannBC()
            if k == "symmetry":
                return m.SymmetryBC()
            return m.ExtrapolationBC(b[1])
        if isinstance(self.bc, tuple) and isinstance(self.bc[0], str):
            return one(self.bc)
        return tuple(one(b) if isinstance(b[0], str) else (one(b[0]), one(b[1])) for b in self.bc)

    def engine_grid(self, m):
        return m.CartesianGrid(self.lc, self.hc, self.n)

    def engine_field(self, m, ctx=None):
        return m.MeshField(self.phi0.copy(order="F"), self.engine_grid(m), bc=self.engine_bc(m), ctx=ctx)

    def engine_terms(self, m, phi, ctx=None):
        out = []
        g = phi.mesh
        for t in self.terms:
            k = t["kind"]
            if k == "advection":
                sch = m.Upwind() if t.get("scheme") == "upwind" else m.WENO5()
                if "separable" in t:
                    sc, tabs = t["separable"]
                    base = m.SeparableVelocity(g, sc, tabs, ctx=ctx)
                elif "field" in t:
                    base = m.MeshField(np.asfortranarray(t["field"].astype(self.dtype)), g, ctx=ctx)
                else:
                    base = tuple(t["const"])
                if "cos_period" in t:
                    base = m.TimeScaled(base, ("cos", t["cos_period"]))
                out.append(m.AdvectionTerm(base, sch))
            elif k == "normal":
                v = m.MeshField(np.asfortranarray(t["field"].astype(self.dtype)), g, ctx=ctx) if "field" in t else t["const"]
                out.append(m.NormalMotionTerm(v))
            elif k == "curvature":
                b = m.MeshField(np.asfortranarray(t["field"].astype(self.dtype)), g, ctx=ctx) if "field" in t else t["const"]
                out.append(m.CurvatureTerm(b))
            elif k == "eikonal":
                out.append(m.EikonalReinitializationTerm(phi) if t.get("frozen", True) else m.EikonalReinitializationTerm())
        return tuple(out)


# ---- BASELINE.json configs, scaled ----------------------------------------------
def c1_circle_rotation(n=128, dtype=np.float64, bc=("periodic",)):
    """C1: 2-D circle SDF under rigid rotation, WENO5, periodic."""
    lc, hc, nn = (-1, -1), (1, 1), (n, n)
    x, y = coords(lc, hc, nn)
    phi = np.hypot(x - 0.3, y) - 0.4
    u = np.stack([bcast(-y, nn), bcast(x, nn)], axis=0)
    return Case("C1", lc, hc, nn, phi, [dict(kind="advection", field=u)], bc, dtype)


def c2_zalesak_curvature(n=256, dtype=np.float64):
    """C2: Zalesak disk rotation + curvature (b = -0.01), Neumann (docs/src/example-zalesak.md:21-40)."""
    lc, hc, nn = (-1.5, -1.5), (1.5, 1.5), (n, n)
    x, y = coords(lc, hc, nn)
    cx, cy, r, w, hgt = -0.75, 0.0, 0.5, 0.2, 1.0
    disk = np.hypot(x - cx, y - cy) - r
    rec = np.maximum(np.abs(x - cx) - w / 2, np.abs(y - (cy - r)) - hgt / 2)
    phi = np.maximum(disk, -rec)
    u = np.stack([bcast(-y, nn), bcast(x, nn)], axis=0)
    return Case("C2", lc, hc, nn, bcast(phi, nn), [dict(kind="advection", field=u), dict(kind="curvature", const=-0.01)],
                ("neumann",), dtype)


def enright_tables(lc, hc, n):
    x, y, z = [c.ravel() for c in coords(lc, hc, n)]
    s2 = lambda a: np.sin(np.pi * a) ** 2
    s = lambda a: np.sin(2 * np.pi * a)
    tabs = [[s2(x), s(y), s(z)], [s(x), s2(y), s(z)], [s(x), s(y), s2(z)]]
    return (2.0, -1.0, -1.0), tabs


def c3_enright(n=64, dtype=np.float64, separable=False, period=3.0):
    """C3: 3-D sphere in the Enright/LeVeque deformation field x cos(pi t / T), WENO5, Neumann."""
    lc, hc, nn = (0, 0, 0), (1, 1, 1), (n, n, n)
    x, y, z = coords(lc, hc, nn)
    phi = np.sqrt((x - 0.35) ** 2 + (y - 0.35) ** 2 + (z - 0.35) ** 2) - 0.15
    sc, tabs = enright_tables(lc, hc, nn)
    if separable:
        term = dict(kind="advection", separable=(sc, tabs), cos_period=period)
    else:
        X, Y, Z = np.meshgrid(*[np.arange(k) for k in nn], indexing="ij", sparse=True)
        u = np.stack([((sc[d] * tabs[d][0][X]) * tabs[d][1][Y]) * tabs[d][2][Z] for d in range(3)], axis=0)
        term = dict(kind="advection", field=u, cos_period=period)
    return Case("C3", lc, hc, nn, phi, [term], ("neumann",), dtype)


def c4_eikonal(n=64, dtype=np.float64, frozen=True):
    """C4: Eikonal reinitialisation of a perturbed sphere SDF, Neumann."""
    lc, hc, nn = (-1, -1, -1), (1, 1, 1), (n, n, n)
    x, y, z = coords(lc, hc, nn)
    r = np.sqrt(x * x + y * y + z * z)
    phi = (r - 0.5) * (1 + 0.4 * np.sin(3 * np.pi * x) * np.sin(3 * np.pi * y) * np.sin(3 * np.pi * z))
    return Case("C4", lc, hc, nn, phi, [dict(kind="eikonal", frozen=frozen)], ("neumann",), dtype)


def c5_normal_advection(n=64, dtype=np.float64):
    """C5: NormalMotionTerm(v field = 0.2) + AdvectionTerm(u = (-y, x, 0) field), Neumann."""
    lc, hc, nn = (-1, -1, -1), (1, 1, 1), (n, n, n)
    x, y, z = coords(lc, hc, nn)
    phi = np.sqrt((x - 0.3) ** 2 + y * y + z * z) - 0.4
    v = np.full(nn, 0.2)
    u = np.stack([bcast(-y, nn), bcast(x, nn), np.zeros(nn)], axis=0)
    return Case("C5", lc, hc, nn, phi, [dict(kind="normal", field=v), dict(kind="advection", field=u)], ("neumann",), dtype)


def cut_cells(phi):
    """Per-cell cut flag over all 2^N corners: vmin <= 0 <= vmax (meshfield.jl:566-575)."""
    N = phi.ndim
    vmin = vmax = None
    for corner in np.ndindex(*([2] * N)):
        sl = tuple(slice(c, phi.shape[d] - 1 + c) for d, c in enumerate(corner))
        v = phi[sl]
        vmin = v if vmin is None else np.minimum(vmin, v)
        vmax = v if vmax is None else np.maximum(vmax, v)
    return (vmin <= 0) & (vmax >= 0)
