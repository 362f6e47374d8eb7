"""N > 1 path on CPU (gloo, world_size 2 and 3): the slab plan exported by the C ABI (lsm_slab_plan) and the
halo-exchange rule of liblsm_b200 (lsm_api.cu:exchange_halo — 3 whole planes of the last axis per side, with
the reference's periodic rule "node n duplicates node 1" for the wrap-around exchange) are replayed with the CPU
oracle in stored-ghost mode: the assembled multi-rank result must equal the single-domain oracle bit for bit.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HALO = 3


def exchange(buf, nl, rank, world, periodic):
    """Mirror of exchange_halo: buf has HALO ghost planes on both sides of the last axis."""
    import torch

    def plane(k0):   # planes [k0, k0+HALO) in owned numbering (k0 may be negative / >= nl)
        return buf[..., HALO + k0:HALO + k0 + HALO]

    ops, recvs = [], []

    up = rank + 1 if rank + 1 < world else (0 if periodic else -1)
    down = rank - 1 if rank > 0 else (world - 1 if periodic else -1)
    top_first = nl - HALO if rank + 1 < world else nl - 1 - HALO     # wrap: planes nl-4..nl-2 are rank 0's low ghosts
    bottom_first = 0 if rank > 0 else 1                               # wrap: ghost(n+k) = node(1+k)
    # upward traffic first, then downward (a pair of ranks matches messages in issue order)
    if up >= 0:
        ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(plane(top_first))), up))
    if down >= 0:
        r = torch.empty(plane(-HALO).shape, dtype=torch.float64)
        ops.append(dist.P2POp(dist.irecv, r, down)); recvs.append((-HALO, r))
    if down >= 0:
        ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(plane(bottom_first))), down))
    if up >= 0:
        r = torch.empty(plane(nl).shape, dtype=torch.float64)
        ops.append(dist.P2POp(dist.irecv, r, up)); recvs.append((nl, r))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for k0, r in recvs:
        plane(k0)[...] = r.numpy()


def worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import helpers as H
    import lsm_b200 as m
    lib = m._lib.lib()
    ok = True
    for periodic in (False, True):
        n = (14, 12, 9 * world + 2)
        lc, hc = (-1, -1, -1), (1, 1, 1)
        X = H.coords(lc, hc, n)
        phi0 = H.bcast(np.sqrt((X[0] - 0.2) ** 2 + X[1] ** 2 + (X[2] + 0.1) ** 2) - 0.5, n)
        u = np.stack([H.bcast(0.3 + 0 * X[0], n), H.bcast(-0.2 + 0.1 * X[1], n), H.bcast(1.0 + 0.2 * np.sin(X[0]), n)], axis=0)
        v = H.bcast(0.1 + 0.05 * np.cos(X[2]), n)
        zbc = O.PERIODIC if periodic else O.EXTRAP(1)
        bc = [O.NEUMANN, O.EXTRAP(2), zbc]
        # ---- single-domain oracle
        fo = O.Field(phi0.copy(order="F"), lc, hc, bc=bc)
        terms = [O.normal_motion(np.asfortranarray(v)), O.advection(np.asfortranarray(u))]
        dt = 0.5 * O.compute_cfl(fo, terms, 0.0)
        for s in range(3):
            O.advance(fo, O.RK3, terms, s * dt, dt)
        # ---- this rank's slab, ghost planes stored
        f, c = C.c_int32(), C.c_int32()
        assert lib.lsm_slab_plan(n[2], world, rank, C.byref(f), C.byref(c)) == 0
        z0, nl = f.value, c.value
        lo_halo = rank > 0 or periodic
        hi_halo = rank < world - 1 or periodic
        slab_bc = [O.NEUMANN, O.EXTRAP(2), (O.HALO if lo_halo else zbc, O.HALO if hi_halo else zbc)]

        def mk():
            a = np.zeros((n[0], n[1], nl + 2 * HALO), order="F")
            a[..., HALO:HALO + nl] = phi0[..., z0:z0 + nl]
            return a

        phi, b1, b2 = mk(), mk(), mk()
        fs = O.Field(phi, lc, hc, bc=slab_bc, gl=[0, 0, HALO], gr=[0, 0, HALO], nglob=list(n), off=[0, 0, z0])
        lterms = [O.normal_motion(np.asfortranarray(v[..., z0:z0 + nl])), O.advection(np.asfortranarray(u[..., z0:z0 + nl]))]
        exchange(fs.vals, nl, rank, world, periodic)
        for s in range(3):
            # stage inputs: phi (S1), buf1 (S2), buf2 (S3); each output's ghosts are exchanged before it is differentiated
            O.stage(fs, O.RK3, 1, b1, b2, lterms, s * dt, dt); exchange(b1, nl, rank, world, periodic)
            O.stage(fs, O.RK3, 2, b1, b2, lterms, s * dt, dt); exchange(b2, nl, rank, world, periodic)
            O.stage(fs, O.RK3, 3, b1, b2, lterms, s * dt, dt); exchange(fs.vals, nl, rank, world, periodic)
        mine = fs.vals[..., HALO:HALO + nl]
        ok &= bool(np.array_equal(mine, fo.vals[..., z0:z0 + nl]))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_invariance_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
