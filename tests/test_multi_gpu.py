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
