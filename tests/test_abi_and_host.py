"""CPU-only checks: the C-ABI library loads and exports every symbol include/lsm_b200.h declares,
host-side logic (grid, BC normalisation, slab plan, step plan) and loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def m():
    import lsm_b200
    return lsm_b200


def test_header_symbols_all_exported(m):
    hdr = open(os.path.join(ROOT, "include", "lsm_b200.h")).read()
    declared = set(re.findall(r"\b(lsm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(m._lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(m._lib.SYMBOLS), "ctypes binding and header disagree"
    assert lib.lsm_abi_version() == 1


def test_integration_doc_covers_every_entry_point():
    hdr = open(os.path.join(ROOT, "include", "lsm_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [f for f in set(re.findall(r"\b(lsm_[a-z0-9_]+)\s*\(", hdr)) if f not in doc]
    assert not missing, f"INTEGRATION.md does not mention {missing}"


def test_header_cites_reference(m):
    hdr = open(os.path.join(ROOT, "include", "lsm_b200.h")).read()
    for cite in ("timestepping.jl:101-122", "levelsetterms.jl:22-38", "meshfield.jl:213-260", "boundaryconditions.jl:166-188"):
        assert cite in hdr


def test_product_never_touches_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "levelsetmethods.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "lsm_oracle" not in txt and "import oracle" not in txt and "orc_" not in txt, f


def test_no_gpu_fails_loudly(m):
    lib = m._lib.lib()
    n = C.c_int32(-1)
    rc = lib.lsm_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(m.LSMError) as ei:
        m.Context(0)
    assert ei.value.code == m._lib.ERR_CUDA


def test_slab_plan(m):
    lib = m._lib.lib()
    for n, g in ((1024, 8), (512, 3), (130, 4), (7, 7)):
        covered = []
        for r in range(g):
            f, c = C.c_int32(), C.c_int32()
            assert lib.lsm_slab_plan(n, g, r, C.byref(f), C.byref(c)) == 0
            covered += list(range(f.value, f.value + c.value))
            assert abs(c.value - n / g) < 1
        assert covered == list(range(n))
    f, c = C.c_int32(), C.c_int32()
    assert lib.lsm_slab_plan(10, 2, 2, C.byref(f), C.byref(c)) == m._lib.ERR_ARG


def test_step_plan_replays_the_reference_loop(m):
    """lsm_step_plan (pure host; what lsm_integrate hands to the resident cluster kernel of small 2-D grids) against a literal
    transcription of _integrate!'s loop (timestepping.jl:104-118): same step sizes bit for bit, same step count, same final t —
    for final times that are / are not multiples of the step, dt_max below and above the CFL step, a start time != 0, step limits."""
    import math
    import random

    def reference(t0, tf, dt_max, cfl, dt_cfl, max_steps):
        t, dts = t0, []
        while t <= tf - math.ulp(t):                       # Base.eps(t) == ulp(t), eps(0.0) == 5e-324
            if 0 <= max_steps <= len(dts):
                break
            dt = min(dt_max, cfl * dt_cfl, tf - t)
            dts.append(dt)
            t += dt
        return dts, t

    rng = random.Random(7)
    cases = [(0.0, 1.0, math.inf, 0.5, 0.015625, -1), (0.0, 1.0, math.inf, 0.5, 0.0123, -1), (0.25, 0.25, math.inf, 0.5, 0.01, -1),
             (0.0, 0.3, 0.004, 0.5, 0.0123, -1), (1.7, 2.9, math.inf, 0.9, 0.0371, 10), (0.0, 1e-3, math.inf, 0.5, 1.0, -1),
             (0.0, 1.0, math.inf, 0.5, 0.0123, 0)]
    for _ in range(40):
        t0 = rng.choice([0.0, rng.uniform(0, 3)])
        cases.append((t0, t0 + rng.uniform(0, 2), rng.choice([math.inf, rng.uniform(1e-3, 1e-1)]), rng.choice([0.5, 0.9, 1.0]),
                      rng.uniform(1e-3, 0.2), rng.choice([-1, -1, 5, 200])))
    for t0, tf, dt_max, cfl, dt_cfl, ms in cases:
        runs, steps, t_end = m.step_plan(t0, tf, dt_max, cfl, dt_cfl, ms)
        dts, t_ref = reference(t0, tf, dt_max, cfl, dt_cfl, ms)
        flat = [dt for dt, c in runs for _ in range(c)]
        assert flat == dts and steps == len(dts) and t_end == t_ref, (t0, tf, dt_max, cfl, dt_cfl, ms)
        assert all(runs[i][0] != runs[i + 1][0] for i in range(len(runs) - 1))          # maximal runs
        if ms < 0 and dts:
            assert len(runs) <= 3 and abs(t_end - tf) <= 4 * math.ulp(tf)
    with pytest.raises(m.TimeError):
        m.step_plan(1.0, 0.5, math.inf, 0.5, 0.01)
    with pytest.raises(m.LSMError):
        m.step_plan(0.0, 1.0, math.inf, 0.5, float("nan"))


# ---- test/test-meshes.jl ----
def test_grid(m):
    g = m.CartesianGrid((-1, 0), (1, 3), (100, 50))
    assert g.size == (100, 50) and len(g) == 5000
    assert g.getnode(1, 1) == (-1.0, 0.0) and g.getnode(100, 50) == (1.0, 3.0)
    g = m.CartesianGrid((-1, -1), (1, 1), meshsize=0.5)
    assert g.size == (5, 5) and g.meshsize() == pytest.approx((0.5, 0.5))
    assert g.getnode(5, 5) == (1.0, 1.0)
    g = m.CartesianGrid((0, 0), (1, 1), meshsize=0.3)
    assert g.size == (5, 5) and all(h <= 0.3 for h in g.meshsize())
    g = m.CartesianGrid((0, 0), (2, 1), meshsize=(0.4, 0.3))
    assert g.size == (6, 5)
    for bad in (dict(meshsize=-0.1), dict(meshsize=(0.1,))):
        with pytest.raises(ValueError):
            m.CartesianGrid((0, 0), (1, 1), **bad)
    with pytest.raises(ValueError):
        m.CartesianGrid((1, 1), (0, 0), meshsize=0.1)


# ---- test/test-boundaryconditions.jl ----
def test_normalize_bc(m):
    P, N, E = m.PeriodicBC(), m.NeumannBC(), m.ExtrapolationBC(2)
    assert m._normalize_bc(P, 2) == ((P, P), (P, P))
    assert m._normalize_bc((P, P), 2) == ((P, P), (P, P))
    assert m._normalize_bc((P, N), 2) == ((P, P), (N, N))
    r = m._normalize_bc([P, (E, N)], 2)
    assert r[0] == (P, P) and r[1] == (E, N)
    with pytest.raises(ValueError):
        m._normalize_bc([(P, E), (E, N)], 2)
    with pytest.raises(ValueError):
        m.ExtrapolationBC(-1)


def test_equation_construction_errors(m):
    g = m.CartesianGrid((-1, -1), (1, 1), (8, 8))
    phi = m.MeshField(lambda x: x[0] ** 2 + x[1] ** 2 - 0.25, g)
    assert phi.vals.shape == (8, 8) and not phi.has_boundary_conditions()
    with pytest.raises(m.BCError):
        m.LevelSetEquation(terms=(m.CurvatureTerm(-0.1),), ic=phi)                 # levelsetequation.jl:69-70
    with pytest.raises(ValueError):
        m.LevelSetEquation(terms=[m.CurvatureTerm(-0.1)], ic=phi, bc=m.NeumannBC())  # _normalize_terms
    eq = m.LevelSetEquation(terms=m.CurvatureTerm(-0.1), ic=phi, bc=m.NeumannBC())
    assert isinstance(eq.integrator, m.RK2) and eq.integrator.cfl == 0.5           # default RK2 (:61)
    assert eq.state is not phi and np.array_equal(eq.state.peek(), phi.peek())     # ic is copied
    with pytest.raises(m.TimeError):
        m.integrate(eq, -1.0)                                                      # levelsetequation.jl:196


def test_next_row_argument_checks_need_no_gpu(m):
    """The argument checks of the "next" rows (velocityextension.jl:25-37, test-velocityextension.jl:87-101; set operations on
    equal-size real-valued fields) fire on the host, before any device work."""
    g = m.CartesianGrid((-1, -1), (1, 1), (41, 41))
    phi = m.MeshField(lambda x: x[0] + x[1], g)
    F = m.MeshField(np.zeros((41, 41)), g)
    with pytest.raises(ValueError):
        m.extend_along_normals(F, phi, nb_iters=-1)
    with pytest.raises(ValueError):
        m.extend_along_normals(F, phi, cfl=0.0)
    with pytest.raises(ValueError):
        m.extend_along_normals(m.MeshField(np.zeros((2, 2)), m.CartesianGrid((-1, -1), (1, 1), (2, 2))), phi)
    with pytest.raises(ValueError):
        m.extend_along_normals(F, phi, frozen=np.zeros((40, 41), dtype=bool))
    with pytest.raises(ValueError):
        m.extend_along_normals(F, phi, frozen=np.zeros((41, 41), dtype=np.int32))      # mask must contain Bool values
    with pytest.raises(ValueError):
        m.union_(phi, m.MeshField(np.zeros((4, 4)), m.CartesianGrid((-1, -1), (1, 1), (4, 4))))
    u = m.MeshField(lambda x: (-x[1], x[0]), g)
    with pytest.raises(ValueError):
        m.complement_(u)                                                                # check_real_valued (levelsetops.jl:254)


def test_fails_loudly_without_the_extension(tmp_path):
    """The product has no CPU fallback: with the shared library missing, the first call raises ImportError (checked in a
    fresh interpreter so that this process's loaded library is untouched)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import lsm_b200 as m\n"
            "try:\n    m._lib.lib()\nexcept ImportError as e:\n    print('LOUD', 'no CPU fallback' in str(e))\n" % ROOT)
    env = dict(os.environ, LSM_B200_SO=str(tmp_path / "does_not_exist.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert "LOUD True" in out.stdout, out.stdout + out.stderr


def test_julia_glue_binds_declared_entry_points_with_matching_arity():
    """julia/LSMB200.jl cannot run here (no Julia in the image): at least every `ccall((:name, LIB), ret, (argtypes...), ...)`
    must name a function include/lsm_b200.h declares, with as many argument types as the C prototype has parameters, and the
    module must define methods on the reference's seam functions (timestepping.jl:101,126-202; levelsetterms.jl:22)."""
    hdr = open(os.path.join(ROOT, "include", "lsm_b200.h")).read()
    src = open(os.path.join(ROOT, "levelsetmethods.jl_b200", "julia", "LSMB200.jl")).read()
    protos = {}
    for mm in re.finditer(r"\b(?:int32_t|const char\*)\s+(lsm_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        args = mm.group(2).strip()
        protos[mm.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    calls = re.findall(r"ccall\(\(:(lsm_[a-z0-9_]+), LIB\),\s*\w+,\s*\(([^()]*)\)", src, re.S)
    assert len(calls) >= 18
    for name, argt in calls:
        assert name in protos, f"{name} is not declared in the header"
        n = 0 if argt.strip() == "" else len([a for a in re.split(r",(?![^{]*})", argt) if a.strip()])
        assert n == protos[name], f"{name}: Julia passes {n} argument types, the C prototype has {protos[name]}"
    for seam in ("LSM._integrate!(ls, ϕ::DeviceMeshField", "LSM._alloc_buffers(", "LSM._advance!(", "LSM.compute_cfl(terms, ϕ::DeviceMeshField",
                 "<: LSM.AbstractMeshField{N, T, V}", "LSM.update_band!(ϕ::DeviceMeshField"):
        assert seam in src, seam
