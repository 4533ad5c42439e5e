"""GPU, one process per GPU (needs >= 2 B200s: `gpurun --gpus 2`): row-partitioned SpMV / CG / BiCGSTAB through the
C ABI with NVLink peer-memory halo exchange and all-reduced scalars, against the oracle on the whole problem."""
import pytest

from test_dist_gloo import launch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _gpus():
    import eigen_git_mirror_b200 as egm
    return egm.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("name", ["poisson3d", "convdiff3d", "powerlaw"])
@pytest.mark.parametrize("loop_mode", [1, 4])
def test_row_partitioned_solvers(world, name, loop_mode, monkeypatch):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("B200S_LOOP_MODE", str(loop_mode))  # 1 = WHILE graph, 4 = persistent cooperative CG kernel
    res = launch(world, "gpu", name, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert res.stdout.count("gpu ok") == world
