import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/oracle"]
import numpy as np, lsm_b200 as m, helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
case = H.c3_enright(n)
phi = case.engine_field(m)
eq = m.LevelSetEquation(terms=case.engine_terms(m, phi), ic=phi, integrator=m.RK3())
m.integrate(eq, 3 * 0.5 * m.compute_cfl(eq.terms, eq.state, 0.0))
out = eq.state.peek()
print("mode", os.environ.get("LSM_B200_TMA_MODE"), "NO_TMA", os.environ.get("LSM_B200_NO_TMA"), "steps", eq.steps_taken, "checksum", float(np.abs(out).sum()), flush=True)
