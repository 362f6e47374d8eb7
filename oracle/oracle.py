"""ctypes loader for the CPU ORACLE (``liblsm_oracle.so``).

TEST INFRASTRUCTURE ONLY — see ``lsm_oracle.h``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
package never does.

Arrays are NumPy, Fortran-ordered (column-major, like a Julia ``Array{V,N}``); vector-valued
coefficient fields have shape ``(N, n1, ..., nN)`` (AoS, like ``Array{SVector{N,T},N}``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblsm_oracle.so")

F32, F64 = 0, 1
BC_NONE, BC_PERIODIC, BC_EXTRAP, BC_SYMMETRY, BC_HALO = -1, 0, 1, 2, 3
TERM_ADVECTION, TERM_NORMAL, TERM_CURVATURE, TERM_EIKONAL = 0, 1, 2, 3
UPWIND, WENO5 = 0, 1
COEF_CONST, COEF_FIELD, COEF_SEPARABLE, COEF_NONE = 0, 1, 2, 3
TS_NONE, TS_COS, TS_HOST = 0, 1, 2
FE, RK2, RK3 = 0, 1, 2
OPS = {"D0": 0, "D+": 1, "D-": 2, "weno5-": 3, "weno5+": 4, "D2_0": 5, "D2++": 6, "D2--": 7, "D2": 8}


class _BC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("P", C.c_int32)]


class _Field(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("dtype", C.c_int32),
        ("n", C.c_int32 * 3), ("gl", C.c_int32 * 3), ("gr", C.c_int32 * 3),
        ("lc", C.c_double * 3), ("hc", C.c_double * 3),
        ("nglob", C.c_int32 * 3), ("off", C.c_int32 * 3),
        ("bc", (_BC * 2) * 3),
        ("vals", C.c_void_p),
    ]


class _Term(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("scheme", C.c_int32), ("coef_kind", C.c_int32), ("tscale_kind", C.c_int32),
        ("cval", C.c_double * 3), ("tparam", C.c_double),
        ("field", C.c_void_p), ("field_dtype", C.c_int32), ("_pad", C.c_int32),
        ("tab", (C.POINTER(C.c_double) * 3) * 3),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its committed Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "lsm_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "lsm_oracle.h"))):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dbl, i32p, fp, tp = C.c_double, C.POINTER(C.c_int32), C.POINTER(_Field), C.POINTER(_Term)
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        L.orc_meshsize.argtypes = [fp, C.c_int]; L.orc_meshsize.restype = dbl
        L.orc_getnode.argtypes = [fp, i32p, C.POINTER(dbl)]
        L.orc_getindex.argtypes = [fp, i32p]; L.orc_getindex.restype = dbl
        L.orc_deriv.argtypes = [fp, C.c_int, i32p, C.c_int, C.c_int]; L.orc_deriv.restype = dbl
        L.orc_weno5.argtypes = [dbl] * 5; L.orc_weno5.restype = dbl
        L.orc_limiter.argtypes = [dbl] * 2; L.orc_limiter.restype = dbl
        L.orc_curvature.argtypes = [fp, i32p]; L.orc_curvature.restype = dbl
        L.orc_volume.argtypes = [fp]; L.orc_volume.restype = dbl
        L.orc_perimeter.argtypes = [fp]; L.orc_perimeter.restype = dbl
        L.orc_compute_term.argtypes = [fp, tp, i32p, dbl, dbl]; L.orc_compute_term.restype = dbl
        L.orc_compute_cfl_term.argtypes = [fp, tp, dbl, dbl]; L.orc_compute_cfl_term.restype = dbl
        L.orc_compute_cfl.argtypes = [fp, tp, C.c_int, dbl, C.POINTER(dbl), C.POINTER(dbl)]
        L.orc_compute_cfl.restype = C.c_int
        L.orc_tscale.argtypes = [tp, dbl]; L.orc_tscale.restype = dbl
        L.orc_eikonal_s0.argtypes = [fp, C.POINTER(dbl)]
        L.orc_stage.argtypes = [fp, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, tp, C.c_int, dbl, dbl,
                                C.POINTER(dbl)]
        L.orc_stage.restype = C.c_int
        L.orc_csg.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]; L.orc_csg.restype = None
        L.orc_extend_along_normals.argtypes = [fp, C.c_void_p, C.c_int, dbl, C.c_void_p, dbl, dbl]; L.orc_extend_along_normals.restype = C.c_int
        L.orc_nstages.argtypes = [C.c_int]; L.orc_nstages.restype = C.c_int
        L.orc_advance.argtypes = [fp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, tp, C.c_int, dbl, dbl]
        L.orc_advance.restype = C.c_int
        L.orc_integrate.argtypes = [fp, C.c_int, dbl, C.c_void_p, tp, C.c_int, dbl, dbl, dbl, C.c_int64,
                                    C.POINTER(dbl), C.POINTER(C.c_int64)]
        L.orc_integrate.restype = C.c_int
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_get_max_threads())


def _norm_bc(bc, ndim):
    """Mirror of ``_normalize_bc`` (boundaryconditions.jl:166-188) on (kind, P) tuples."""
    if bc is None:
        return [((BC_NONE, 0), (BC_NONE, 0))] * ndim
    if isinstance(bc, tuple) and len(bc) == 2 and all(isinstance(x, int) for x in bc):
        return [(bc, bc)] * ndim
    if len(bc) != ndim:
        raise ValueError("invalid number of boundary conditions")
    out = []
    for b in bc:
        if isinstance(b, tuple) and len(b) == 2 and all(isinstance(x, int) for x in b):
            out.append((b, b))
        else:
            l, r = b
            if (l[0] == BC_PERIODIC) != (r[0] == BC_PERIODIC):
                raise ValueError("periodic boundary conditions cannot be mixed with others")
            out.append((tuple(l), tuple(r)))
    return out


PERIODIC = (BC_PERIODIC, 0)
NEUMANN = (BC_EXTRAP, 0)
SYMMETRY = (BC_SYMMETRY, 0)
HALO = (BC_HALO, 0)


def EXTRAP(p):
    return (BC_EXTRAP, int(p))


class Field:
    """A dense node field + grid + BCs (the oracle's ``MeshField``)."""

    def __init__(self, vals, lc, hc, bc=None, gl=None, gr=None, nglob=None, off=None):
        vals = np.asarray(vals)
        if vals.dtype not in (np.float32, np.float64):
            vals = vals.astype(np.float64)
        self.vals = np.asfortranarray(vals)
        self.ndim = self.vals.ndim
        nd = self.ndim
        self.gl = list(gl) if gl is not None else [0] * nd
        self.gr = list(gr) if gr is not None else [0] * nd
        self.n = [self.vals.shape[d] - self.gl[d] - self.gr[d] for d in range(nd)]
        self.lc = [float(x) for x in lc]
        self.hc = [float(x) for x in hc]
        self.bc = _norm_bc(bc, nd)
        self.nglob = list(nglob) if nglob is not None else list(self.n)
        self.off = list(off) if off is not None else [0] * nd

    @property
    def dtype(self):
        return self.vals.dtype

    def c(self, vals=None):
        f = _Field()
        f.ndim = self.ndim
        f.dtype = F32 if self.vals.dtype == np.float32 else F64
        for d in range(3):
            f.n[d] = self.n[d] if d < self.ndim else 1
            f.gl[d] = self.gl[d] if d < self.ndim else 0
            f.gr[d] = self.gr[d] if d < self.ndim else 0
            f.lc[d] = self.lc[d] if d < self.ndim else 0.0
            f.hc[d] = self.hc[d] if d < self.ndim else 1.0
            f.nglob[d] = self.nglob[d] if d < self.ndim else 1
            f.off[d] = self.off[d] if d < self.ndim else 0
            for s in range(2):
                k, p = self.bc[d][s] if d < self.ndim else (BC_NONE, 0)
                f.bc[d][s].kind = k
                f.bc[d][s].P = p
        arr = self.vals if vals is None else vals
        f.vals = arr.ctypes.data
        return f

    def _I(self, I):
        return (C.c_int32 * 3)(*([int(i) for i in I] + [1] * (3 - len(I))))

    def meshsize(self, dim=None):
        if dim is None:
            return [lib().orc_meshsize(C.byref(self.c()), d + 1) for d in range(self.ndim)]
        return lib().orc_meshsize(C.byref(self.c()), dim)

    def getnode(self, I):
        x = (C.c_double * 3)()
        lib().orc_getnode(C.byref(self.c()), self._I(I), x)
        return [x[d] for d in range(self.ndim)]

    def __getitem__(self, I):
        """BC-aware read with a 1-based (possibly out-of-grid) multi-index."""
        if not isinstance(I, tuple):
            I = (I,)
        return lib().orc_getindex(C.byref(self.c()), self._I(I))

    def deriv(self, op, I, dim, dim2=0):
        return lib().orc_deriv(C.byref(self.c()), OPS[op], self._I(I), dim, dim2)

    def curvature(self, I):
        return lib().orc_curvature(C.byref(self.c()), self._I(I))

    def volume(self):
        return lib().orc_volume(C.byref(self.c()))

    def perimeter(self):
        return lib().orc_perimeter(C.byref(self.c()))

    def nodes(self):
        """Node coordinates as broadcastable arrays, computed as lc + (I-1)*h (meshes.jl:114-117)."""
        out = []
        for d in range(self.ndim):
            h = (self.hc[d] - self.lc[d]) / (self.nglob[d] - 1)
            x = self.lc[d] + (np.arange(self.n[d], dtype=np.float64) + self.off[d]) * h
            shape = [1] * self.ndim
            shape[d] = self.n[d]
            out.append(x.reshape(shape))
        return out


class Term:
    """One level-set term descriptor.  Keeps coefficient arrays alive."""

    def __init__(self, kind, scheme=WENO5, coef_kind=COEF_CONST, cval=(0, 0, 0), field=None,
                 tscale=TS_NONE, tparam=1.0, tabs=None):
        self.kind, self.scheme, self.coef_kind = kind, scheme, coef_kind
        self.cval = [float(x) for x in cval] + [0.0] * (3 - len(cval))
        self.tscale, self.tparam = tscale, float(tparam)
        self.field = None if field is None else np.asfortranarray(field)
        self.tabs = None
        if tabs is not None:
            self.tabs = [[np.ascontiguousarray(t, dtype=np.float64) for t in row] for row in tabs]

    def c(self):
        t = _Term()
        t.kind, t.scheme, t.coef_kind, t.tscale_kind = self.kind, self.scheme, self.coef_kind, self.tscale
        for d in range(3):
            t.cval[d] = self.cval[d]
        t.tparam = self.tparam
        if self.field is not None:
            t.field = self.field.ctypes.data
            t.field_dtype = F32 if self.field.dtype == np.float32 else F64
        if self.tabs is not None:
            for d, row in enumerate(self.tabs):
                for a, tab in enumerate(row):
                    t.tab[d][a] = tab.ctypes.data_as(C.POINTER(C.c_double))
        return t


def advection(velocity, scheme=WENO5, tscale=TS_NONE, tparam=1.0):
    if isinstance(velocity, np.ndarray):
        return Term(TERM_ADVECTION, scheme, COEF_FIELD, field=velocity, tscale=tscale, tparam=tparam)
    return Term(TERM_ADVECTION, scheme, COEF_CONST, cval=tuple(velocity), tscale=tscale, tparam=tparam)


def advection_separable(scales, tabs, scheme=WENO5, tscale=TS_NONE, tparam=1.0):
    return Term(TERM_ADVECTION, scheme, COEF_SEPARABLE, cval=tuple(scales), tabs=tabs, tscale=tscale, tparam=tparam)


def normal_motion(speed):
    if isinstance(speed, np.ndarray):
        return Term(TERM_NORMAL, coef_kind=COEF_FIELD, field=speed)
    return Term(TERM_NORMAL, coef_kind=COEF_CONST, cval=(speed,))


def curvature(b):
    if isinstance(b, np.ndarray):
        return Term(TERM_CURVATURE, coef_kind=COEF_FIELD, field=b)
    return Term(TERM_CURVATURE, coef_kind=COEF_CONST, cval=(b,))


def eikonal(s0=None):
    if s0 is None:
        return Term(TERM_EIKONAL, coef_kind=COEF_NONE)
    return Term(TERM_EIKONAL, coef_kind=COEF_FIELD, field=s0)


def eikonal_s0(phi0: Field) -> np.ndarray:
    out = np.zeros(phi0.n, dtype=np.float64, order="F")
    lib().orc_eikonal_s0(C.byref(phi0.c()), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _terms_array(terms):
    arr = (_Term * len(terms))()
    for i, t in enumerate(terms):
        arr[i] = t.c()
    return arr


def compute_term(phi: Field, term: Term, I, t=0.0, g=1.0):
    tc = term.c()
    return lib().orc_compute_term(C.byref(phi.c()), C.byref(tc), phi._I(I), float(t), float(g))


class CFLError(ArithmeticError):
    """Mirror of the ArgumentError thrown at levelsetterms.jl:26."""


def compute_cfl(phi: Field, terms, t=0.0, gscale=None):
    arr = _terms_array(terms)
    dt = C.c_double()
    gs = None if gscale is None else (C.c_double * len(terms))(*gscale)
    rc = lib().orc_compute_cfl(C.byref(phi.c()), arr, len(terms), float(t), gs, C.byref(dt))
    if rc:
        raise CFLError(f"invalid time-step based on CFL condition: dt = {dt.value}")
    return dt.value


def nstages(integrator):
    return lib().orc_nstages(integrator)


def stage(phi: Field, integrator, s, buf1, buf2, terms, tc, dt, gscale=None):
    """Run stage ``s`` (1-based) in place on ``phi.vals`` / ``buf1`` / ``buf2`` (same shape, F-order)."""
    arr = _terms_array(terms)
    gs = None if gscale is None else (C.c_double * len(terms))(*gscale)
    rc = lib().orc_stage(C.byref(phi.c()), integrator, s, phi.vals.ctypes.data, buf1.ctypes.data, buf2.ctypes.data,
                         arr, len(terms), float(tc), float(dt), gs)
    assert rc == 0, rc


def advance(phi: Field, integrator, terms, tc, dt, bufs=None):
    b1, b2 = bufs if bufs is not None else (phi.vals.copy(order="F"), phi.vals.copy(order="F"))
    arr = _terms_array(terms)
    rc = lib().orc_advance(C.byref(phi.c()), integrator, phi.vals.ctypes.data, b1.ctypes.data, b2.ctypes.data,
                           arr, len(terms), float(tc), float(dt))
    assert rc == 0, rc


def integrate(phi: Field, integrator, terms, tf, t0=0.0, cfl=0.5, dt_max=float("inf"), max_steps=-1):
    """``integrate!`` with default hooks.  Mutates ``phi.vals``; returns (t, steps)."""
    arr = _terms_array(terms)
    t_out, steps = C.c_double(), C.c_int64()
    rc = lib().orc_integrate(C.byref(phi.c()), integrator, float(cfl), phi.vals.ctypes.data, arr, len(terms),
                             float(t0), float(tf), float(dt_max), int(max_steps), C.byref(t_out), C.byref(steps))
    if rc == 1:
        raise CFLError("invalid time-step based on CFL condition")
    if rc == 2:
        raise ValueError(f"final time {tf} must be >= initial time {t0}")
    return t_out.value, steps.value


def extend_along_normals(F: np.ndarray, phi: Field, nb_iters=50, cfl=0.45, frozen=None, interface_band=1.5, min_norm=1e-14):
    """``extend_along_normals!`` (velocityextension.jl:20-69).  Mutates and returns ``F`` (same dtype/shape as phi, F-order)."""
    assert F.flags.f_contiguous and F.dtype == phi.vals.dtype and F.shape == phi.vals.shape
    fr = None
    if frozen is not None:
        fr = np.asfortranarray(frozen.astype(np.uint8))
    rc = lib().orc_extend_along_normals(C.byref(phi.c()), F.ctypes.data, int(nb_iters), float(cfl),
                                        None if fr is None else fr.ctypes.data, float(interface_band), float(min_norm))
    if rc:
        raise ValueError("invalid extend_along_normals arguments")
    return F


def csg(op: str, a: np.ndarray, b=None) -> np.ndarray:
    """``union!`` / ``intersect!`` / ``setdiff!`` / ``complement!`` (levelsetops.jl:253-325) on value arrays; returns a new array."""
    code = {"union": 0, "intersect": 1, "setdiff": 2, "complement": 3}[op]
    out = np.array(a, order="F", copy=True)
    src = None if b is None else np.asfortranarray(b, dtype=out.dtype)
    lib().orc_csg(F64 if out.dtype == np.float64 else F32, out.ctypes.data, None if src is None else src.ctypes.data, out.size, code)
    return out
