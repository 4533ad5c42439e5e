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


@pytest.mark.parametrize("world,name", [(2, "cg_256"), (8, "cg_256"), (8, "bicg_256"), (8, "cg_512"), (4, "cg_256")])
def test_baseline_configs_row_partitioned_vs_reference(world, name):
    """configs[1], [2] (256^3, full convergence + trajectories) and configs[4] (512^3, trajectories k <= 50) on N
    GPUs against the unmodified reference's outputs (tests/golden/fullsize_v1.npz)."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    res = launch(world, "fullsize", name, timeout=850)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert res.stdout.count("fullsize ok") == world


@pytest.mark.parametrize("world", [2])
def test_a_rank_that_fails_before_launch_does_not_hang_the_others(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    res = launch(world, "deadpeer", "poisson3d", timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert res.stdout.count("deadpeer ok") == world
