#!/usr/bin/env python
"""bench.py — cell-updates/s per RK3 step of the dense-grid hot path (BASELINE.json metric).

Workload (N = 1): BASELINE.json configs[2] "C3": 3-D sphere SDF in the Enright/LeVeque deformation
velocity (stored Float64 velocity field x cos(pi t / 3)), 512^3 nodes, WENO5 + TVD-RK3, NeumannBC.
One bench "step" = one RK3 step of integrate! = CFL reduction + 3 fused stage kernels.
N > 1: the same workload weak-scaled — every rank owns a 512 x 512 x 512 slab of a 512 x 512 x (512 N)
grid, 3-plane halos exchanged over NCCL each stage ("scaling": "weak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 512] [--impl reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-timed (CUDA events on the library's compute stream,
max over ranks) with inputs resident in HBM; `e2e` = same metric through the public API with pinned
HOST buffers (H2D of phi, K steps, D2H of phi inside the timed region); `roofline` = the fused stage
kernel against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle (a C++ restatement of
the reference — Julia is not installed) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "cell-updates/s per RK3 step (3D WENO5 advection, Float64)"
UNIT = "cell-updates/s"
PERIOD = 3.0
B_ALG = 136.0          # bytes per cell-update, 3-D f64 RK3 advection with a stored velocity: (8 + 3N) * s
CPU_SAMPLE_N = 96      # the CPU legs run the same configuration on a 96^3 grid (bounded sample)


def c5_slab(n, z_first, nz_loc):
    """BASELINE.json configs[4] "C5" on a slab of the n^3 grid on (-1,-1,-1)..(1,1,1): phi0 = |x - (0.3,0,0)| - 0.4,
    NormalMotionTerm(v = 0.2 stored scalar field) + AdvectionTerm(u = (-y, x, 0) stored field)."""
    h = 2.0 / (n - 1)
    x = (-1.0 + np.arange(n) * h).reshape(n, 1, 1)
    y = (-1.0 + np.arange(n) * h).reshape(1, n, 1)
    z = (-1.0 + (np.arange(nz_loc) + z_first) * h).reshape(1, 1, nz_loc)
    phi = np.empty((n, n, nz_loc), order="F")
    np.sqrt((x - 0.3) ** 2 + y * y + z * z, out=phi)
    phi -= 0.4
    u = np.zeros((3, n, n, nz_loc), order="F")
    u[0] = -y
    u[1] = x
    v = np.full((n, n, nz_loc), 0.2, order="F")
    return phi, u, v


def enright_slab(n, nz_glob, z_first, nz_loc, lz):
    """phi0 and the stored velocity of C3 on a slab [z_first, z_first+nz_loc) of an n x n x nz_glob grid on
    (0,0,0)..(1,1,lz).  Velocity components are rank-1 products ((s*X)*Y)*Z (tests/helpers.py)."""
    hx = 1.0 / (n - 1)
    hz = lz / (nz_glob - 1)
    x = (np.arange(n) * hx).reshape(n, 1, 1)
    y = (np.arange(n) * hx).reshape(1, n, 1)
    z = ((np.arange(nz_loc) + z_first) * hz).reshape(1, 1, nz_loc)
    phi = np.sqrt((x - 0.35) ** 2 + (y - 0.35) ** 2 + (z - 0.35) ** 2) - 0.15
    s2 = lambda a: np.sin(np.pi * a) ** 2
    s1 = lambda a: np.sin(2 * np.pi * a)
    u = np.empty((3, n, n, nz_loc), order="F")
    u[0] = ((2.0 * s2(x)) * s1(y)) * s1(z)
    u[1] = ((-1.0 * s1(x)) * s2(y)) * s1(z)
    u[2] = ((-1.0 * s1(x)) * s1(y)) * s2(z)
    return np.asfortranarray(phi), u


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        out["samples"] = len(rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any(r[5 + i].strip().lower().startswith("active") for r in rows):
                out["reasons"].append(nm)
        return out


def hbm_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per stage-kernel launch from the committed ncu capture, if any (profiles/traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["c3_stage_dram_bytes_per_launch"]
    except Exception:
        return None


def cpu_leg(steps, warmup, threads):
    """The CPU oracle on a CPU_SAMPLE_N^3 grid of the same configuration; returns (updates/s, seconds/step)."""
    import oracle as O
    import helpers as H
    O.set_threads(threads)
    case = H.c3_enright(CPU_SAMPLE_N, period=PERIOD)
    f, terms = case.oracle_field(), case.oracle_terms()
    nodes = CPU_SAMPLE_N ** 3
    t = 0.0
    for _ in range(warmup):
        dt = 0.5 * O.compute_cfl(f, terms, t)
        O.advance(f, O.RK3, terms, t, dt); t += dt
    t0 = time.perf_counter()
    for _ in range(steps):
        dt = 0.5 * O.compute_cfl(f, terms, t)
        O.advance(f, O.RK3, terms, t, dt); t += dt
    sec = time.perf_counter() - t0
    return nodes * steps / sec, sec / steps


def host_threads():
    """Threads the CPU legs use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which would
    otherwise silently serialise the reference arm under N > 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is 100 % Julia
    and Julia is not in this image, so this is the oracle port (kind "port") with all host threads."""
    if rank != 0:
        return
    threads = host_threads()
    v, sps = cpu_leg(args.steps, max(args.warmup, 1), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 Enright sphere, {CPU_SAMPLE_N}^3 sample of the 512^3 config, WENO5+RK3, NeumannBC, stored velocity x cos(pi t/3)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{CPU_SAMPLE_N}^3 grid, {args.steps} RK3 steps, OpenMP over the slowest axis; C++ restatement of the "
                                   "reference (oracle/), not Julia — the reference's hot loop itself is serial"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=0, help="nodes per axis (c3: per GPU, default 512; c5: global, default 1024)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3 = headline (weak-scaled Enright advection); c5 = BASELINE configs[4], strong-scaled 1024^3 normal motion + advection")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 strict generic, 2 tiled")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    # stdout carries exactly ONE JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG >= VERSION (WARN included),
    # so send NCCL's debug output to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

    import lsm_b200 as m
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = m.Context.from_torch_distributed(local)
    else:
        ctx = m.Context(local)
    m.set_default_context(ctx)
    ctx.set_option(m._lib.OPT_KERNEL, args.kernel)

    G = world
    c5 = args.workload == "c5"
    n = args.n or (1024 if c5 else 512)
    if c5:
        nz = n
        grid = m.CartesianGrid((-1, -1, -1), (1, 1, 1), (n, n, nz))
        z_first, nz_loc = ctx.slab(nz) if G > 1 else (0, nz)
        phi0, u, v = c5_slab(n, z_first, nz_loc)
        phi = m.MeshField(phi0, grid, bc=m.NeumannBC(), ctx=ctx)
        vel = m.MeshField(u, grid, ctx=ctx)
        spd = m.MeshField(v, grid, ctx=ctx)
        del u, v
        terms = (m.NormalMotionTerm(spd), m.AdvectionTerm(vel, m.WENO5()))
        args.no_e2e = True
        args.no_cpu = True
    else:
        nz = n * G
        lz = float(G)
        grid = m.CartesianGrid((0, 0, 0), (1, 1, lz), (n, n, nz))
        z_first, nz_loc = ctx.slab(nz) if G > 1 else (0, nz)
        phi0, u = enright_slab(n, nz, z_first, nz_loc, lz)
        phi = m.MeshField(phi0, grid, bc=m.NeumannBC(), ctx=ctx)
        vel = m.MeshField(u, grid, ctx=ctx)
        del u
        terms = (m.AdvectionTerm(m.TimeScaled(vel, ("cos", PERIOD)), m.WENO5()),)
    nodes_total = n * n * nz
    balg = 160.0 if c5 else B_ALG
    eq = m.LevelSetEquation(terms=terms, ic=phi, integrator=m.RK3())
    state = eq.state
    lib, L = m._lib.lib(), m._lib
    import ctypes as C
    low = m.api._Lowered(eq.terms, state, 0.0)
    dev = state.device()

    def steps_on_device(k, t0):
        t_out, st = C.c_double(), C.c_int64()
        L.check(lib.lsm_integrate(ctx.handle, L.RK3, 0.5, dev, low.arr, len(eq.terms), t0, 1e9, float("inf"), k, C.byref(t_out), C.byref(st)))
        assert st.value == k
        return t_out.value

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    # ---- warm-up, then the device-timed region -------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None      # samples clocks from the warm-up on (all of it is under load)
    if sampler:
        time.sleep(0.4)                                       # let nvidia-smi start sampling
    t = steps_on_device(args.warmup, 0.0)
    ctx.set_option(L.OPT_TIME_STAGES, 1)
    barrier()
    ctx.reset_counters()
    ctx.event_record(0)
    t = steps_on_device(args.steps, t)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    clocks = sampler.stop() if sampler else None
    cnt = ctx.counters()
    ctx.set_option(L.OPT_TIME_STAGES, 0)
    if dist is not None:
        import torch
        tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = nodes_total * args.steps / (ms * 1e-3)

    # ---- e2e: public API, pinned host buffers, H2D + K steps + D2H inside the timed region ---------
    e2e = None
    if not args.no_e2e:
        host = state.vals                        # host array of the state (device copy becomes stale)
        host[...] = phi0
        L.check(lib.lsm_host_register(host.ctypes.data, host.nbytes))
        eq.t = 0.0
        tf = None
        barrier()
        w0 = time.perf_counter()
        state.vals                               # mark host as the fresh copy -> integrate! uploads it
        dev2 = state.device()                    # H2D of phi (pinned)
        t_out, st = C.c_double(), C.c_int64()
        L.check(lib.lsm_integrate(ctx.handle, L.RK3, 0.5, dev2, low.arr, len(eq.terms), 0.0, 1e9, float("inf"), args.steps, C.byref(t_out), C.byref(st)))
        state._mark_device_advanced()
        res = state.peek()                       # D2H of phi (pinned)
        chk = float(res[0, 0, 0])
        barrier()
        w = time.perf_counter() - w0
        if dist is not None:
            tt = torch.tensor([w], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            w = float(tt.item())
        L.check(lib.lsm_host_unregister(host.ctypes.data))
        e2e = {"value": nodes_total * args.steps / w, "unit": UNIT,
               "h2d_bytes_per_step": host.nbytes * G / args.steps, "d2h_bytes_per_step": host.nbytes * G / args.steps,
               "call": f"one integrate! of {args.steps} RK3 steps: H2D phi from pinned host, steps, D2H phi; velocity field resident",
               "check": chk}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    stage_ms = cnt["sum_stage_ms"] / max(cnt["timed_stages"], 1)
    bytes_per_launch = (balg / 3.0) * (n * n * nz_loc)           # average over the 3 stage launches of a step (c3: 40+48+48 B/node)
    achieved = bytes_per_launch / (stage_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if c5 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"C5 NormalMotionTerm(v field) + AdvectionTerm(u field), {n}^3 global (BASELINE.json configs[4]), WENO5 + TVD-RK3, "
                                "NeumannBC, slab-decomposed with NCCL halo exchange" if c5 else
                                f"C3 Enright sphere {n}x{n}x{nz} (BASELINE.json configs[2]; {n}^3 per GPU), WENO5 + TVD-RK3, NeumannBC, "
                                "stored Float64 velocity field x cos(pi t/3), CFL reduction every step (fused into the last RK stage)"),
                   "grid": [n, n, nz], "parallelism": f"slab{G}" if G > 1 else "single",
                   "l2": "inputs larger than L2 (each field >= 1 GB per GPU)", "kernel": ["auto", "strict-generic", "tiled"][args.kernel]},
        "clocks": clocks,
        "gpu_launches": int(cnt["kernel_launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(), "peak_source": peak_src,
                     "kernel": "fused RK stage (stencil + Hamiltonian + RK combination)",
                     "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": stage_ms,
                     "launches_timed": int(cnt["timed_stages"]),
                     "frac_of_nominal_8TBs": achieved / 8000.0,      # reported against both the measured and the nominal roof
                     "co_bound": "FP64 issue: a DP instruction holds the SMSP dispatch port 2 cycles (tools/issue_model.cu); "
                                 "18.3 T lane-ops/s measured (tools/fp64_peak.cu); see DESIGN.md §4.1"},
        "e2e": e2e,
    }
    if G == 1 and not args.no_cpu:
        thr = host_threads()
        v_all, _ = cpu_leg(3, 1, thr)
        v_one, _ = cpu_leg(1, 0, 1)
        line["cpu_baseline"] = {"value": v_all, "unit": UNIT, "cores": thr, "kind": "port",
                                "sample": f"C3 on a {CPU_SAMPLE_N}^3 grid, 3 RK3 steps, OpenMP x{thr}; 1 thread (the reference's hot loop "
                                          f"is serial): {v_one:.4g} updates/s. C++ restatement of the reference, not Julia (not installed)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
