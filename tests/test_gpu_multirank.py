"""Multi-GPU generation loop on NCCL (needs >= 2 GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`):
population sharded over ranks (uneven shards included), per-generation record all-gather issued asynchronously and consumed one
generation later, merged elites fed to every rank's hall of fame.  Checked: the hall of fame and the logbook are identical on
every rank, the logbook describes the whole population, the hall holds the best distinct individuals of the merged elites."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_total", [32, 33])
def test_sharded_generation_loop_is_rank_invariant(tmp_path, n_total):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    world, ngen = 2, 3
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + n_total), os.path.join(ROOT, "tests", "multirank", "worker.py"), str(tmp_path), str(n_total), str(ngen)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    logs = [json.loads(str(r["log"])) for r in res]
    # identical hall of fame and logbook on every rank
    for r in res[1:]:
        assert np.array_equal(r["hof_genomes"], res[0]["hof_genomes"]) and np.array_equal(r["hof_fitness"], res[0]["hof_fitness"])
    for row0, row1 in zip(*logs):
        for key in ("gen", "avg", "std", "min", "max"):
            assert row0[key] == row1[key], key
    assert [row["gen"] for row in logs[0]] == list(range(ngen + 1))
    # shards tile the population; the last logbook row is the statistics of the final global fitness vector
    assert res[0]["lo"] == 0 and res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == n_total
    fit = np.concatenate([r["fitness"] for r in res])
    last = logs[0][-1]
    np.testing.assert_allclose([last["avg"], last["std"], last["min"], last["max"]], [fit.mean(), fit.std(), fit.min(), fit.max()], rtol=1e-12, atol=1e-15)
    assert sum(r[-1]["nevals"] for r in logs) <= 2 * n_total
    # hall of fame: sorted best first, distinct, and it contains the best individual of the final population
    hf, hg = res[0]["hof_fitness"], res[0]["hof_genomes"]
    assert np.all(np.diff(hf) <= 0) and len({g.tobytes() for g in hg}) == len(hg)
    assert hf[0] >= fit.max()
