"""One shell workload at benchmark size for profiling (ncu -k regex:...) and A/B timing of single entry points.
usage: prof_shell.py WORKLOAD [NEL] [STEPS]   WORKLOAD in roof|plate_nh|plate_svk|balloon|cylinder|tension|frustrum"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsstructuralanalysis_b200 import capi, workloads as W                      # noqa: E402
from gsstructuralanalysis_b200.ops import ShellAssembler                        # noqa: E402
from gsstructuralanalysis_b200.problem import KL_MAT_NH, KL_MAT_SVK             # noqa: E402

MK = {"roof": lambda n: W.roof(n, 3), "plate_nh": lambda n: W.tutorial_paraboloid(n, 3, KL_MAT_NH, False),
      "plate_svk": lambda n: W.tutorial_paraboloid(n, 3, KL_MAT_SVK, False), "balloon": W.balloon, "cylinder": W.cylinder,
      "tension": W.tension_sheet, "frustrum": W.frustrum}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "roof"
    nel = int(sys.argv[2]) if len(sys.argv) > 2 else 576
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    pr = MK[name](nel)
    asm = ShellAssembler(pr, device=0)
    L = max(np.ptp(pr.surface.cp[:, k]) for k in range(3))
    stream = torch.cuda.current_stream().cuda_stream
    r = torch.empty(asm.n_dofs, dtype=torch.float64, device="cuda")
    for rel in (1e-3, 1e-4, 1e-5, 0.0):
        x = torch.from_numpy(W.smooth_state(pr, rel * L, noise=0.0)).cuda()
        asm.residual_device(x.data_ptr(), r.data_ptr(), 1.0, -1.0, stream)
        asm.jacobian_device(x.data_ptr(), stream)
        if asm.check(stream) == 0:
            break
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tr = tj = 0.0
    for _ in range(steps):
        ev[0].record()
        asm.residual_device(x.data_ptr(), r.data_ptr(), 1.0, -1.0, stream)
        ev[1].record()
        asm.jacobian_device(x.data_ptr(), stream)
        ev[2].record()
        torch.cuda.synchronize()
        tr += ev[0].elapsed_time(ev[1]); tj += ev[1].elapsed_time(ev[2])
    jm = C.c_float()
    capi.check(asm.L.kl_jacobian_kernel_ms(asm.h, C.byref(jm)))
    print(f"{name} nel={nel} n_dofs={asm.n_dofs} residual_call_ms={tr / steps:.3f} jacobian_call_ms={tj / steps:.3f} main_kernel_ms={jm.value:.3f}")


if __name__ == "__main__":
    main()
