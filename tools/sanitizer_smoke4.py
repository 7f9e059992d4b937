"""compute-sanitizer workload for the sum-factorised internal force (k_points<P,true> through kl_assemble_device, k_residual,
k_residual<P,FULL>): the partial sums alias the dead part of the element staging area, so racecheck is the point."""
import numpy as np, sys
sys.path.insert(0, '.')
import torch
from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.ops import ShellAssembler
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR
for degree, mat, comp in ((3, KL_MAT_NH, False), (2, KL_MAT_SVK, False), (4, KL_MAT_MR, True)):
    a = ShellAssembler(W.tutorial_paraboloid(5, degree, mat, comp))
    x = W.displacement_state(a.n_dofs, 1e-4)
    xd = torch.from_numpy(x).cuda()
    rd = torch.zeros(a.n_dofs, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    a.assemble_device(xd.data_ptr(), rd.data_ptr(), 1.0, -1.0, s)
    assert a.check(s) == 0
    torch.cuda.synchronize()
    ok, r = a.residual(x)
    f = a.boundaryForce(x, 0)
    print("assemble_device", degree, mat, comp, float(np.abs(rd.cpu().numpy() - r).max()), np.abs(f).max())
    a.close()
