"""Multi-GPU tests of the `-m gpu` tier: run on a box with >= 2 GPUs, skipped otherwise.

tests/mgpu_check.py (slab-decomposed run == single-GPU run BIT FOR BIT, == oracle within 1e-10; NCCL halo exchange with and
without interior/boundary overlap, periodic wrap across ranks, uneven slabs, set operations / volume / perimeter / velocity
extension on decomposed fields) is launched under torchrun with one process per GPU.  Its output is kept under
gpurun_out/mgpu_check_<N>gpu.log (a copy of a 2/4/8-rank run is committed under profiles/).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_slab_decomposition_is_bitwise_invariant(nproc):
    have = _gpu_count()
    if have < nproc:
        pytest.skip(f"needs {nproc} GPUs, {have} visible")
    if nproc != 2 and nproc != have:
        pytest.skip("larger rank counts run only when they use the whole box")
    port = 29600 + nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    env = dict(os.environ, NCCL_DEBUG_FILE="/dev/stderr")
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"mgpu_check_{nproc}gpu.log"), "w") as f:
        f.write(p.stdout)
    assert p.returncode == 0, p.stdout[-4000:]
    assert "ALL OK" in p.stdout and "[FAIL]" not in p.stdout, p.stdout[-2000:]


def test_single_process_multi_gpu_context_matches_single_gpu():
    """lsm_ctx_create_multi: ONE process (this one) drives 2 GPUs through ncclCommInitAll contexts and the lsm_multi_* fan-out.
    The slab-decomposed results must equal the single-GPU run bit for bit (C3: time-scaled advection with the host-side CFL
    candidates gathered over ranks; C5: two-term x-pair kernel), and compute_cfl must agree exactly."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import lsm_b200 as m
    import helpers as H
    mc = m.MultiContext([0, 1])
    solo = m.Context(0)
    try:
        for case, integ, steps in ((H.c3_enright(48), m.RK3, 8), (H.c5_normal_advection(40), m.RK3, 6), (H.c4_eikonal(40), m.RK2, 5)):
            def build(ctx):
                phi = case.engine_field(m, ctx=ctx)
                return m.LevelSetEquation(terms=case.engine_terms(m, phi, ctx=ctx), ic=phi, integrator=integ())
            ref = build(solo)
            dt0 = 0.5 * m.compute_cfl(ref.terms, ref.state, 0.0)
            tf = dt0 * steps * (1 - 1e-12)
            m.integrate(ref, tf)
            eqs = [build(c) for c in mc.ranks]
            assert 0.5 * m.compute_cfl_multi(mc, [e.terms for e in eqs], [e.state for e in eqs], 0.0) == dt0
            m.integrate_multi(mc, eqs, tf)
            got = m.MultiContext.gather([e.state for e in eqs])
            assert eqs[0].t == ref.t and eqs[0].steps_taken == ref.steps_taken
            assert np.array_equal(got, ref.state.peek()), (case.name, float(np.abs(got - ref.state.peek()).max()))
    finally:
        mc.close(); solo.close()
