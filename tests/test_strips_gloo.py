"""N>1 path on CPU: world_size-2 gloo processes run the strip partition + halo exchange of
gsstructuralanalysis_b200/parallel.py.  Each rank's partial matrix comes from the oracle restricted to its element
rows (the oracle is the checker here, the product kernels are exercised by the -m gpu twin of this test)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    from gsstructuralanalysis_b200 import workloads as W
    return {"roof": lambda: W.roof(9), "paraboloid": lambda: W.tutorial_paraboloid(8), "balloon": lambda: W.balloon(8),
            "tension": lambda: W.tension_sheet(8)}        # tension: the collapsed east side is ONE DoF shared by every strip


def _plan(pr, world, rank):
    from gsstructuralanalysis_b200.parallel import plan_strips, function_supports
    n1, n2 = pr.surface.n
    nel2 = function_supports(pr.surface.U[1], 3)[2]
    return plan_strips(n1, n2, 3, nel2, pr.dof_map, pr.n_free, world, rank, knots2=pr.surface.U[1])


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsstructuralanalysis_b200 import workloads as W
    from gsstructuralanalysis_b200.parallel import plan_strips, exchange_halo, value_ranges
    from oracle.binding import Oracle
    CASES = _cases()
    pr = CASES[case]()
    full = Oracle(pr)
    x = W.displacement_state(full.n_dofs, 1e-5 if case == "tension" else 1e-3)
    Kfull, Rint_full = full.jacobian_values(x), full.internal_force(x)
    n1, n2 = pr.surface.n
    plan = _plan(pr, world, rank)
    part = Oracle(pr)
    part.set_strip(plan.e2_begin, plan.e2_end)
    Kp = torch.from_numpy(part.jacobian_values(x))
    Rp = torch.from_numpy(part.internal_force(x))      # partial internal force
    moved = exchange_halo(plan, full.outer, Kp, Rp, dist)
    ok = True
    for (a, b) in value_ranges(plan.owned_cols, full.outer):
        ok &= bool(np.abs(Kp.numpy()[a:b] - Kfull[a:b]).max() <= 1e-13 * np.abs(Kfull).max())
    for (c0, c1) in plan.owned_cols:
        ok &= bool(np.abs(Rp.numpy()[c0:c1] - Rint_full[c0:c1]).max() <= 1e-13 * max(np.abs(Rint_full).max(), 1e-300))
    # halo-compute: the owner integrates the overlapping element rows itself and nothing is exchanged
    if plan.compute_begin >= 0:
        part.set_strip(plan.compute_begin, plan.e2_end)
        Kh, Rh = part.jacobian_values(x), part.internal_force(x)
        for (a, b) in value_ranges(plan.owned_cols, full.outer):
            ok &= bool(np.abs(Kh[a:b] - Kfull[a:b]).max() <= 1e-13 * np.abs(Kfull).max())
        for (c0, c1) in plan.owned_cols:
            ok &= bool(np.abs(Rh[c0:c1] - Rint_full[c0:c1]).max() <= 1e-13 * max(np.abs(Rint_full).max(), 1e-300))
    else:
        ok &= case == "tension"          # only the collapsed side forbids it
    owned = sum(c1 - c0 for c0, c1 in plan.owned_cols)
    q.put((rank, ok, owned, moved))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["roof", "paraboloid", "balloon", "tension"])
def test_strip_partition_world2(case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + {"roof": 0, "paraboloid": 1, "balloon": 2, "tension": 3}[case]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    from oracle.binding import Oracle
    pr = _cases()[case]()
    n = Oracle(pr).n_dofs
    nshared = sum(c1 - c0 for c0, c1 in _plan(pr, 2, 0).shared_cols)
    assert sum(r[2] for r in res) == n + nshared          # every column owned exactly once (shared ones complete on every rank)
    assert (nshared > 0) == (case == "tension")
    assert max(r[3] for r in res) > 0           # something crossed the interface


def test_plan_ranges_cover_all_columns():
    from gsstructuralanalysis_b200 import workloads as W
    from gsstructuralanalysis_b200.parallel import plan_strips
    from oracle.binding import lib
    pr = W.roof(16)
    pr.number_dofs(lib().klo_build_dofmap)
    n1, n2 = pr.surface.n
    for world in (2, 4, 5):
        seen = np.zeros(pr.n_free, dtype=int)
        for r in range(world):
            pl = plan_strips(n1, n2, 3, n2 - 3, pr.dof_map, pr.n_free, world, r)
            for c0, c1 in pl.owned_cols:
                seen[c0:c1] += 1
            if r > 0:
                assert pl.recv_cols == plan_strips(n1, n2, 3, n2 - 3, pr.dof_map, pr.n_free, world, r - 1).send_cols
        assert (seen == 1).all()
    with pytest.raises(ValueError):         # strips thinner than p element rows would need non-neighbour exchanges
        plan_strips(n1, n2, 3, n2 - 3, pr.dof_map, pr.n_free, 8, 0)


def test_function_supports_with_repeated_knots():
    from gsstructuralanalysis_b200.parallel import function_supports
    U = [0, 0, 0, 0, 0.25, 0.5, 0.5, 0.75, 1, 1, 1, 1]      # degree 3, double knot at 0.5: 4 elements, 8 functions
    flo, fhi, ne = function_supports(U, 3)
    assert ne == 4 and list(flo) == [0, 0, 0, 0, 1, 2, 2, 3] and list(fhi) == [0, 1, 1, 2, 3, 3, 3, 3]
