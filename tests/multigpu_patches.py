"""torchrun entry (one process per GPU): a multi-patch assembled by PATCH partition — every rank assembles the patches assigned to
it (kl_mp_set_active), the interface columns are completed on their owners by the NCCL point-to-point exchange of
gsstructuralanalysis_b200/parallel.py — checked against the multi-patch oracle.  Launched by tests/test_gpu_multigpu.py; by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu_patches.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    shared_gpu = torch.cuda.device_count() < world      # the driver's 1-GPU box: both ranks on GPU 0, interface over gloo
    if shared_gpu:
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gsstructuralanalysis_b200 import workloads as W
    from gsstructuralanalysis_b200.ops import MultiPatchAssembler
    from gsstructuralanalysis_b200.parallel import plan_patches, exchange_patches, value_ranges, DevicePointerView
    from tests.mp_problems import cut
    nel = int(os.environ.get("KL_NEL", "24"))
    g1, g2 = [int(v) for v in os.environ.get("KL_GRID", "2x2").split("x")]
    case = os.environ.get("KL_CASE", "roof")
    base = {"roof": lambda: W.roof(nel), "tension": lambda: W.tension_sheet(nel)}[case]()     # tension: ONE collapsed DoF touches every patch
    _, multi, _ = cut(base, [k / g1 for k in range(1, g1)], [k / g2 for k in range(1, g2)])
    asm = MultiPatchAssembler(multi, device=local)
    npatch = len(multi.patches)
    patch_rank = [q * world // npatch for q in range(npatch)]           # contiguous blocks of patches per rank
    plan = plan_patches([p.dof_map for p in multi.patches], multi.n_free, patch_rank, world, rank)
    asm.set_active(plan.active)
    x = W.displacement_state(asm.n_dofs, 0.05 if case == "roof" else 1e-6)
    xd = torch.from_numpy(x).cuda()
    rd = torch.zeros(asm.n_dofs, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    vals = DevicePointerView(asm.values_device_ptr(), asm.nnz).tensor()
    outer, _ = asm.pattern()
    reps = int(os.environ.get("KL_REPS", "2"))
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_asm, t_x, moved = [], [], 0
    for it in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        asm.residual_device(xd.data_ptr(), rd.data_ptr(), 1.0, -1.0, stream)      # partial F_ext - F_int of the rank's patches
        asm.jacobian_device(xd.data_ptr(), stream)
        e1.record()
        if shared_gpu:
            vh, rh = vals.cpu(), rd.cpu()
            moved = exchange_patches(plan, outer, vh, rh, dist)
        else:
            moved = exchange_patches(plan, outer, vals, rd, dist)
        e2.record()
        torch.cuda.synchronize()
        assert asm.check(stream) == 0
        if it > 0:
            t_asm.append(e0.elapsed_time(e1)); t_x.append(e1.elapsed_time(e2))
    if t_asm and not shared_gpu:
        tt = torch.tensor([min(t_asm), min(t_x)], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"PATCHES-TIMING world={world} patches={npatch} n_dofs={asm.n_dofs} assemble_ms={tt[0].item():.3f} exchange_ms={tt[1].item():.3f} "
                  f"quad_pts_per_s={asm.n_qp / ((tt[0].item() + tt[1].item()) * 1e-3):.4e}")
    ok = True
    if os.environ.get("KL_CHECK", "1") == "1":
        from oracle.multipatch import MultiPatchOracle
        orc = MultiPatchOracle(multi)
        Kf, Rf = orc.jacobian_values(x), orc.residual(x)
        v, r = (vh.numpy(), rh.numpy()) if shared_gpu else (vals.cpu().numpy(), rd.cpu().numpy())
        eK = max([np.abs(v[a:b] - Kf[a:b]).max() for a, b in value_ranges(plan.owned_cols, outer) if b > a] + [0.0]) / np.abs(Kf).max()
        eR = max([np.abs(r[c0:c1] - Rf[c0:c1]).max() for c0, c1 in plan.owned_cols if c1 > c0] + [0.0]) / np.abs(Rf).max()
        ok = bool(eK <= 1e-12 and eR <= 1e-12)
        if not ok or os.environ.get("KL_VERBOSE"):
            iface = set(asm.interface_dofs().tolist())
            cols = np.repeat(np.arange(asm.n_dofs), np.diff(outer))
            own = np.zeros(asm.n_dofs, dtype=bool)
            for c0, c1 in plan.owned_cols:
                own[c0:c1] = True
            bad = own[cols] & (np.abs(v - Kf) > 1e-12 * np.abs(Kf).max())
            badcols = np.unique(cols[bad])
            print(f"PATCHES-DEBUG rank={rank} errK={eK:.2e} errR={eR:.2e} bad_cols={len(badcols)} of which interface={sum(int(c) in iface for c in badcols)} "
                  f"send={ {k: len(v_) for k, v_ in plan.send.items()} } recv={ {k: len(v_) for k, v_ in plan.recv.items()} }", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device="cpu" if shared_gpu else "cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"PATCHES world={world} patches={npatch} backend={'gloo, one shared GPU' if shared_gpu else 'nccl'} n_dofs={asm.n_dofs} "
              f"ok={bool(t.item())} interface_bytes_rank0={moved}")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
