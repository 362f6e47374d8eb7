#!/usr/bin/env python
"""pair_check.py — development check for the x-pair kernel (csrc/lsm_pair3d.cu): run the same 3-D WENO5 advection problem
with LSM_OPT_KERNEL = 4 (x-pair kernel, exact epsilon maximum) and 3 (general tiled kernel) and require BIT-IDENTICAL states, over grid shapes
that are not multiples of the tile, every index-map boundary condition, the three integrators, stored / separable velocity
and both dtypes.  (The pytest version of this lives in tests/test_gpu_parity.py.)

    python tools/pair_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import lsm_b200 as m          # noqa: E402
import helpers as H           # noqa: E402

OPT_KERNEL = 0


def case_3d(n, bc, dtype, separable, period=3.0):
    lc, hc = (0, 0, 0), (1, 1, 1)
    x, y, z = H.coords(lc, hc, n)
    phi = np.sqrt((x - 0.35) ** 2 + (y - 0.4) ** 2 + (z - 0.45) ** 2) - 0.15 + 0.02 * np.sin(9 * x) * np.cos(7 * y) * np.sin(5 * z)
    sc, tabs = H.enright_tables(lc, hc, n)
    tabs = [[t + 0.05 * (a + 1) for a, t in enumerate(row)] for row in tabs]       # no exact zeros, sign changes in every direction
    if separable:
        term = dict(kind="advection", separable=(sc, tabs), cos_period=period)
    else:
        X, Y, Z = np.meshgrid(*[np.arange(k) for k in n], indexing="ij", sparse=True)
        u = np.stack([((sc[d] * tabs[d][0][X]) * tabs[d][1][Y]) * tabs[d][2][Z] for d in range(3)], axis=0)
        term = dict(kind="advection", field=u, cos_period=period)
    return H.Case("P", lc, hc, n, phi, [term], bc, dtype)


def run(case, integ, kernel, steps):
    ctx = m.default_context()
    ctx.set_option(OPT_KERNEL, kernel)
    phi = case.engine_field(m)
    eq = m.LevelSetEquation(terms=case.engine_terms(m, phi), ic=phi, integrator=integ)
    dt = 0.5 * m.compute_cfl(eq.terms, eq.state, 0.0)
    m.integrate(eq, steps * dt * (1 - 1e-12))
    out = eq.state.peek().copy()
    ctx.set_option(OPT_KERNEL, 0)
    return eq.t, eq.steps_taken, out


def main():
    bad = 0
    shapes = [(128, 64, 40), (72, 52, 37), (64, 8, 8), (200, 30, 70), (16, 24, 20), (136, 9, 130)]
    bcs = [("neumann",), ("periodic",), ("symmetry",), (("neumann",), ("periodic",), ("symmetry",)),
           ((("neumann",), ("symmetry",)), ("periodic",), (("symmetry",), ("neumann",)))]
    k = 0
    for n in shapes:
        for bc in bcs:
            for dtype in (np.float64, np.float32):
                if dtype == np.float32 and n[0] % 4:
                    continue
                k += 1
                separable = k % 3 == 0
                integ = (m.RK3(), m.RK2(), m.ForwardEuler())[k % 3] if k % 5 else m.RK3()
                case = case_3d(n, bc, dtype, separable)
                a = run(case, integ, 4, 3)
                b = run(case, integ, 3, 3)
                same = a[:2] == b[:2] and np.array_equal(a[2], b[2])
                d = float(np.abs(a[2].astype(np.float64) - b[2].astype(np.float64)).max())
                print(f"n={n} bc={bc} {np.dtype(dtype).name} sep={separable} {type(integ).__name__}: steps {a[1]} "
                      f"{'BITWISE' if same else 'DIFF'} maxabs {d:.3e}", flush=True)
                bad += 0 if same else 1
    cnt = m.default_context().counters()
    print("counters", cnt)
    print("FAILED" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
