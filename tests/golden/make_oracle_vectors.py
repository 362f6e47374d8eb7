#!/usr/bin/env python
"""Writes tests/golden/oracle_vectors.npz: outputs of the CPU ORACLE (oracle/lsm_oracle.cpp, the C++ restatement of the
reference — NOT of Julia, which is not installed here) for small instances of the BASELINE.json configurations.

    python tests/golden/make_oracle_vectors.py

They serve two purposes: (1) regression-pin the oracle itself across rounds (tests/test_oracle_pins.py::test_oracle_golden_vectors
recomputes them), (2) give the GPU suite committed vectors to compare with (tests/test_gpu_parity.py::test_against_committed_vectors),
so that a change that moves oracle AND engine together is still caught.  Inputs are the seed-free analytic configurations of
tests/helpers.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as O      # noqa: E402
import helpers as H     # noqa: E402

CASES = {           # name -> (builder, integrator, steps)
    "C1_32": (lambda dt: H.c1_circle_rotation(32, dt), "RK3", 20),
    "C2_48": (lambda dt: H.c2_zalesak_curvature(48, dt), "RK3", 20),
    "C3_20": (lambda dt: H.c3_enright(20, dt), "RK3", 12),
    "C4_20": (lambda dt: H.c4_eikonal(20, dt), "RK2", 12),
    "C5_20": (lambda dt: H.c5_normal_advection(20, dt), "RK3", 12),
}
INTEG = {"FE": O.FE, "RK2": O.RK2, "RK3": O.RK3}


def run(name, dtype):
    mk, integ, steps = CASES[name]
    case = mk(dtype)
    fo, to = case.oracle_field(), case.oracle_terms()
    tf = 0.5 * O.compute_cfl(fo, to, 0.0) * steps * (1 - 1e-12)
    t, n = O.integrate(fo, INTEG[integ], to, tf)
    return fo.vals, tf, n


def main():
    out = {}
    for name in CASES:
        for dtype, tag in ((np.float64, "f64"), (np.float32, "f32")):
            v, tf, n = run(name, dtype)
            out[f"{name}_{tag}"] = v
            out[f"{name}_{tag}_tf_steps"] = np.array([tf, n], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
