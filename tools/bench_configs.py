#!/usr/bin/env python
"""bench_configs.py — cell-updates/s and HBM-roofline fraction of every BASELINE.json configuration at full size on
one B200 (the headline config C3 is what bench.py reports; this script covers the others with the same method:
CUDA events on the library stream, >= 3 warm-up steps, fields larger than L2 except C1).

    python tools/bench_configs.py [--steps 20] [--only C4] > profiles/configs_rNN.jsonl
Algorithmic bytes per cell-update are SURVEY.md §8(d)'s figures.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import lsm_b200 as m          # noqa: E402
import helpers as H           # noqa: E402

L = m._lib


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(name, case, integ, balg, steps, warmup=3, reps=5):
    ctx = m.default_context()
    phi = case.engine_field(m)
    terms = case.engine_terms(m, phi)
    eq = m.LevelSetEquation(terms=terms, ic=phi, integrator=integ)
    st = eq.state
    low = m.api._Lowered(eq.terms, st, 0.0)
    dev = st.device()
    lib = L.lib()

    def go(k, t0):
        t_out, n = C.c_double(), C.c_int64()
        L.check(lib.lsm_integrate(ctx.handle, integ.code, integ.cfl, dev, low.arr, len(terms), t0, 1e9, float("inf"), k,
                                  C.byref(t_out), C.byref(n)))
        return t_out.value

    t = go(warmup, 0.0) if warmup > 0 else 0.0
    # median of `reps` timed runs of `steps` steps each (device time, CUDA events on the library stream)
    runs = []
    for _ in range(max(reps, 1)):
        ctx.sync(); ctx.reset_counters()
        ctx.event_record(0)
        t = go(steps, t)
        ctx.event_record(1)
        runs.append((ctx.event_elapsed_ms(0, 1), ctx.counters()))
    runs.sort(key=lambda r: r[0])
    ms, cnt0 = runs[len(runs) // 2]
    # per-launch times of the stage kernels: a separate pass with CUDA events around every stage launch
    ctx.set_option(L.OPT_TIME_STAGES, 1)
    ctx.sync(); ctx.reset_counters()
    t = go(steps, t)
    ctx.sync()
    cnt = ctx.counters()
    ctx.set_option(L.OPT_TIME_STAGES, 0)
    cnt["kernel_launches"], cnt["cfl_passes"] = cnt0["kernel_launches"], cnt0["cfl_passes"]
    nodes = int(np.prod(case.n))
    ups = nodes * steps / (ms * 1e-3)
    stage_ms = cnt["sum_stage_ms"] / max(cnt["timed_stages"], 1)
    nst = integ.nstages
    ach = balg / nst * nodes / (stage_ms * 1e-3) / 1e9
    out = {"config": name, "grid": list(case.n), "dtype": np.dtype(case.dtype).name, "integrator": type(integ).__name__,
           "steps": steps, "reps": len(runs), "ms_per_step_min_max": [runs[0][0] / steps, runs[-1][0] / steps], "ms_per_step": ms / steps, "cell_updates_per_s": ups, "avg_stage_launch_ms": stage_ms,
           "algorithmic_bytes_per_update": balg, "achieved_gbs": ach, "peak_gbs": peak(), "frac": ach / peak(),
           "step_frac_of_roofline": ups * balg / 1e9 / peak(), "kernel_launches": cnt["kernel_launches"], "cfl_passes": cnt["cfl_passes"],
           "resident_steps": cnt0.get("resident_steps", 0)}
    print(json.dumps(out), flush=True)
    del eq, phi, terms, low


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--skip", default="")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kernel", type=int, default=0)
    a = ap.parse_args()
    m.default_context().set_option(L.OPT_KERNEL, a.kernel)
    f64, f32 = np.float64, np.float32
    cfgs = [
        ("C1 2-D circle rotation 128^2 periodic f64 (advection, stored u)", lambda: H.c1_circle_rotation(128), m.RK3(), 112.0),
        ("C2 2-D Zalesak + curvature 2048^2 Neumann f64", lambda: H.c2_zalesak_curvature(2048, f64), m.RK3(), 112.0),
        ("C2 2-D Zalesak + curvature 2048^2 Neumann f32", lambda: H.c2_zalesak_curvature(2048, f32), m.RK3(), 56.0),
        ("C3 3-D Enright 512^3 f64 stored velocity x cos", lambda: H.c3_enright(512, f64), m.RK3(), 136.0),
        ("C3 3-D Enright 512^3 f64 separable in-kernel velocity", lambda: H.c3_enright(512, f64, separable=True), m.RK3(), 64.0),
        ("C3 3-D Enright 512^3 f32 stored velocity x cos", lambda: H.c3_enright(512, f32), m.RK3(), 68.0),
        ("C4 3-D Eikonal reinit 512^3 f64 RK3 frozen S0", lambda: H.c4_eikonal(512, f64), m.RK3(), 88.0),
        ("C4 3-D Eikonal reinit 512^3 f64 RK2 (reference default)", lambda: H.c4_eikonal(512, f64), m.RK2(), 64.0),
        ("C5 3-D normal motion + advection 512^3 f64 (1-GPU slice of the 1024^3 config)", lambda: H.c5_normal_advection(512, f64), m.RK3(), 160.0),
    ]
    for name, mk, integ, balg in cfgs:
        if (a.only and not any(name.startswith(o) for o in a.only.split(","))) or (a.skip and name.startswith(a.skip)):
            continue
        run(name, mk(), integ, balg, a.steps if "128^2" not in name else 10 * a.steps, a.warmup, a.reps)


if __name__ == "__main__":
    main()
