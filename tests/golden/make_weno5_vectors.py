#!/usr/bin/env python
"""make_weno5_vectors.py — exact-arithmetic known answers for `_weno5` (src/derivatives.jl:61-81).

The reference holds no stored vectors for `_weno5`, and its constants (1/3, 7/6, 11/6, 13/12, 0.1, 0.6, 0.3, 1e-6, 1e-99) are
not dyadic, so no non-trivial input has an exactly representable result.  What CAN be pinned independently of any rounding
order is the value of the reference's formula evaluated in EXACT rational arithmetic on (a) dyadic inputs and (b) the exact
rational values of the Float64 literals Julia parses (`Fraction(1/3)` is the double Julia computes for `1 / 3`, and so on).
The correctly rounded double of that rational is stored; a faithful Float64 implementation (about 40 roundings, no
cancellation in the final sum for these inputs) must land within a few ulp of it.  tests/test_oracle_pins.py checks the
oracle against these numbers (rel 1e-14); the GPU parity tests inherit the pin through the oracle.

    python tests/golden/make_weno5_vectors.py          # rewrites the "weno5_exact" entry of reference_known_answers.json
"""
import json
import os
from fractions import Fraction as Fr

HERE = os.path.dirname(os.path.abspath(__file__))


def weno5_exact(v):
    v1, v2, v3, v4, v5 = [Fr(x) for x in v]
    c13, c76, c116, c16, c56 = Fr(1 / 3), Fr(7 / 6), Fr(11 / 6), Fr(1 / 6), Fr(5 / 6)
    d1 = c13 * v1 - c76 * v2 + c116 * v3
    d2 = -c16 * v2 + c56 * v3 + c13 * v4
    d3 = c13 * v3 + c56 * v4 - c16 * v5
    c1312, c14 = Fr(13 / 12), Fr(1 / 4)
    S1 = c1312 * (v1 - 2 * v2 + v3) ** 2 + c14 * (v1 - 4 * v2 + 3 * v3) ** 2
    S2 = c1312 * (v2 - 2 * v3 + v4) ** 2 + c14 * (v2 - v4) ** 2
    S3 = c1312 * (v3 - 2 * v4 + v5) ** 2 + c14 * (3 * v3 - 4 * v4 + v5) ** 2
    eps = Fr(1.0e-6) * max(x * x for x in (v1, v2, v3, v4, v5)) + Fr(1.0e-99)
    a1, a2, a3 = Fr(0.1) / (S1 + eps) ** 2, Fr(0.6) / (S2 + eps) ** 2, Fr(0.3) / (S3 + eps) ** 2
    s = a1 + a2 + a3
    return float((a1 / s) * d1 + (a2 / s) * d2 + (a3 / s) * d3)       # float(Fraction) rounds correctly


INPUTS = [
    [1.0, 0.5, 2.0, -1.5, 0.25],            # rough data: all three stencils active
    [0.125, 0.25, 0.5, 1.0, 2.0],           # monotone, growing: weight on the smooth upwind side
    [3.0, 3.0, 3.0, 3.0, -5.0],             # a kink at the downwind end: stencil 3 switched off
    [-5.0, 3.0, 3.0, 3.0, 3.0],             # a kink at the upwind end: stencil 1 switched off
    [1.0, 2.0, 3.0, 4.0, 5.0],              # linear data: every candidate equals 3.5
    [0.75, -0.75, 0.75, -0.75, 0.75],       # sawtooth
    [2.0 ** -20, 2.0 ** -21, 2.0 ** -19, 2.0 ** -20, 2.0 ** -22],   # small magnitudes (epsilon term relatively the same)
]


def main():
    path = os.path.join(HERE, "reference_known_answers.json")
    d = json.load(open(path))
    d["weno5_exact"] = {
        "source": "src/derivatives.jl:61-81 (_weno5) evaluated in exact rational arithmetic on dyadic inputs and on the exact values of "
                  "its Float64 literals; stored value = correctly rounded double (tests/golden/make_weno5_vectors.py)",
        "rel_tol": 1e-14,
        "cases": [{"v": v, "expected": weno5_exact(v)} for v in INPUTS],
    }
    json.dump(d, open(path, "w"), indent=1)
    for c in d["weno5_exact"]["cases"]:
        print(c)


if __name__ == "__main__":
    main()
