"""N>1 multi-patch path on CPU: world_size-2 gloo processes run the patch partition + interface exchange of
gsstructuralanalysis_b200/parallel.py.  Each rank's partial matrix comes from the multi-patch oracle restricted to the rank's
patches (the oracle is the checker; the product kernels run in the -m gpu twin, tests/multigpu_patches.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(case):
    from gsstructuralanalysis_b200 import workloads as W
    from tests.mp_problems import cut
    from oracle.multipatch import build_dofmap_mp
    base, cuts, ranks = {"roof_2x2": (lambda: W.roof(6), ([0.5], [0.5]), [0, 1, 1, 0]),
                         "balloon_3x1": (lambda: W.balloon(6), ([1.0 / 3, 2.0 / 3], []), [0, 1, 0]),
                         "roof_chain4": (lambda: W.roof(8), ([], [0.25, 0.5, 0.75]), [0, 1, 2, 3]),
                         "tension_chain4": (lambda: W.tension_sheet(8), ([], [0.25, 0.5, 0.75]), [0, 1, 2, 3])}[case]
    _, multi, _ = cut(base(), *cuts)
    multi.number_dofs(build_dofmap_mp)
    return multi, ranks


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsstructuralanalysis_b200.parallel import plan_patches, exchange_patches, value_ranges
    from oracle.multipatch import MultiPatchOracle
    multi, ranks = _problem(case)
    orc = MultiPatchOracle(multi)
    x = (1e-6 if case.startswith("tension") else 1e-3) * np.random.default_rng(4).uniform(-1, 1, orc.n_dofs)
    Kfull, Rfull = orc.jacobian_values(x), orc.residual(x)
    plan = plan_patches([p.dof_map for p in multi.patches], multi.n_free, ranks, world, rank)
    mine = [k for k, a in enumerate(plan.active) if a]
    Kp = torch.from_numpy(orc.jacobian_values(x, parts=mine))
    Rp = torch.from_numpy(orc.residual(x, parts=mine))
    moved = exchange_patches(plan, orc.outer, Kp, Rp, dist)
    ok = True
    for a, b in value_ranges(plan.owned_cols, orc.outer):
        ok &= bool(np.abs(Kp.numpy()[a:b] - Kfull[a:b]).max() <= 1e-13 * np.abs(Kfull).max())
    for c0, c1 in plan.owned_cols:
        ok &= bool(np.abs(Rp.numpy()[c0:c1] - Rfull[c0:c1]).max() <= 1e-13 * max(np.abs(Rfull).max(), 1e-300))
    q.put((rank, ok, sum(c1 - c0 for c0, c1 in plan.owned_cols), moved))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,world", [("roof_2x2", 2), ("balloon_3x1", 2), ("roof_chain4", 4), ("tension_chain4", 4)])
def test_patch_partition(case, world):
    """roof_chain4: a chain of patches, one per rank — every inner rank both sends (to the owner below) and receives (from above)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 500) + {"roof_2x2": 0, "balloon_3x1": 1, "roof_chain4": 2, "tension_chain4": 3}[case]
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    multi, _ = _problem(case)
    assert sum(r[2] for r in res) == multi.n_free          # every column owned exactly once
    assert max(r[3] for r in res) > 0                      # something crossed an interface
