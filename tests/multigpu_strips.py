"""torchrun entry (one process per GPU): strip-partitioned assembly of ONE matrix with the NCCL halo exchange,
checked against the full oracle matrix.  Launched by tests/test_gpu_multigpu.py and usable by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_strips.py"""
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
    # fewer GPUs than ranks (the driver's 1-GPU test box): every rank assembles its strip on GPU 0 and the halo travels through
    # host tensors over gloo — same plan, same kl_set_strip kernels, same exchange code; NCCL needs one device per rank
    shared_gpu = torch.cuda.device_count() < world
    if shared_gpu:
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gsstructuralanalysis_b200 import workloads as W
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from gsstructuralanalysis_b200.parallel import plan_strips, exchange_halo, value_ranges, DevicePointerView, function_supports
    from oracle.binding import Oracle
    nel = int(os.environ.get("KL_NEL", "24"))
    case = os.environ.get("KL_CASE", "roof")
    def roof_pressure():     # follower pressure: the unsymmetric tangent of k_jacobian_sw<false,true> in strips (both phases of the overlap)
        p = W.roof(nel)
        p.pressure = 0.3
        return p
    pr = {"roof": lambda: W.roof(nel), "tension": lambda: W.tension_sheet(nel), "roof_pressure": roof_pressure}[case]()
    asm = ShellAssembler(pr, device=local)
    n1, n2 = pr.surface.n
    plan = plan_strips(n1, n2, 3, function_supports(pr.surface.U[1], 3)[2], pr.dof_map, pr.n_free, world, rank, knots2=pr.surface.U[1])
    asm.set_strip(plan.e2_begin, plan.e2_end)
    x = W.displacement_state(asm.n_dofs, 0.05 if case.startswith("roof") else 1e-5)
    xd = torch.from_numpy(x).cuda()
    rd = torch.zeros(asm.n_dofs, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    asm.jacobian_device(xd.data_ptr(), stream)
    asm.residual_device(xd.data_ptr(), rd.data_ptr(), 0.0, 1.0, stream)     # partial F_int of the strip
    assert asm.check(stream) == 0
    overlapped = os.environ.get("KL_OVERLAP", "1") == "1" and torch.cuda.device_count() >= world
    vals = DevicePointerView(asm.values_device_ptr(), asm.nnz).tensor()
    outer, _ = asm.pattern()
    torch.cuda.synchronize()
    reps = int(os.environ.get("KL_REPS", "1"))
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    if shared_gpu:
        vals_dev, rd_dev = vals, rd
        vals, rd = vals_dev.cpu(), rd_dev.cpu()
    if overlapped:        # two-phase assembly, exchange in flight while the bulk of the strip is assembled
        from gsstructuralanalysis_b200.parallel import assemble_strip_overlapped
        moved = assemble_strip_overlapped(asm, plan, outer, vals, rd, xd.data_ptr(), dist, stream)
        assert asm.check(stream) == 0
    else:
        moved = exchange_halo(plan, outer, vals, rd, dist)      # first call also warms NCCL up
    t_asm, t_x = [], []
    for _ in range(0 if shared_gpu else reps - 1):
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        asm.jacobian_device(xd.data_ptr(), stream)
        asm.residual_device(xd.data_ptr(), rd.data_ptr(), 0.0, 1.0, stream)
        e1.record()
        moved = exchange_halo(plan, outer, vals, rd, dist)
        e2.record()
        torch.cuda.synchronize()
        t_asm.append(e0.elapsed_time(e1)); t_x.append(e1.elapsed_time(e2))
    if t_asm:
        tt = torch.tensor([min(t_asm), min(t_x)], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"STRIPS-TIMING world={world} n_dofs={asm.n_dofs} assemble_ms={tt[0].item():.3f} exchange_ms={tt[1].item():.3f} "
                  f"quad_pts_per_s={asm.n_qp / ((tt[0].item() + tt[1].item()) * 1e-3):.4e}")
    ok = True
    if os.environ.get("KL_CHECK", "1") == "1":
        orc = Oracle(pr)
        Kf, Rf = orc.jacobian_values(x), orc.internal_force(x)
        v, r = vals.cpu().numpy(), rd.cpu().numpy()
        for (a, b) in value_ranges(plan.owned_cols, outer):
            ok &= bool(np.abs(v[a:b] - Kf[a:b]).max() <= 1e-12 * np.abs(Kf).max())
        for (c0, c1) in plan.owned_cols:
            ok &= bool(np.abs(r[c0:c1] - Rf[c0:c1]).max() <= 1e-12 * np.abs(Rf).max())
    t = torch.tensor([1.0 if ok else 0.0], device="cpu" if shared_gpu else "cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"STRIPS world={world} case={case} backend={'gloo, one shared GPU' if shared_gpu else 'nccl'} n_dofs={asm.n_dofs} ok={bool(t.item())} halo_bytes_rank0={moved}")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
