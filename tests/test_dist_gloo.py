"""CPU, world_size 2 and 3 over gloo: the host logic of the multi-GPU path (partition, ghost numbering, halo send
lists exchanged through the allgather callback) reproduces the global SpMV exactly.  See tests/dist_worker.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch(world, mode, name, timeout=240):
    port = 29500 + (os.getpid() * 7 + world * 13 + hash(name) % 50) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           mode, name]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,name", [(2, "poisson3d"), (2, "powerlaw"), (3, "convdiff3d")])
def test_partition_and_halo_plan(world, name):
    res = launch(world, "plan", name)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("plan ok") == world


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,name", [(2, "poisson3d"), (3, "powerlaw"), (2, "varcoef3d")])
def test_one_triangle_row_partitioned(world, name):
    """ConjugateGradient<_, Lower> / <_, Upper> on a row-partitioned matrix: each rank stores one triangle of its rows,
    the plan exchanges the mirror images (SparseSelfAdjointView.h:279-337 applied across ranks)."""
    res = launch(world, "tri", name)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("tri ok") == world
