"""timing of kl_stability (device banded LDL^T) on the roof at several sizes: python tools/stab_bench.py 96 192 384 576"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.ops import ShellAssembler
for nel in [int(a) for a in sys.argv[1:]] or [96]:
    pr = W.roof(nel)
    asm = ShellAssembler(pr)
    x = W.displacement_state(asm.n_dofs, 0.002 * 508.0 / nel)
    ok, _ = asm.jacobian(x, fetch=False)
    assert ok
    for rep in range(2):
        t0 = time.perf_counter()
        ind, neg = asm.stability()
        dt = time.perf_counter() - t0
    n1 = pr.surface.n[0]
    bw = 3 * (3 * n1 + 4) - 1
    flops = asm.n_dofs * float(bw) ** 2
    print(f"STAB nel={nel} n_dofs={asm.n_dofs} half_bandwidth~{bw} band_GiB={asm.n_dofs * (bw + 33) * 8 / 2**30:.1f} "
          f"time_s={dt:.3f} indicator={ind:.6e} negatives={neg} TFLOPs={flops / dt / 1e12:.2f}", flush=True)
    asm.close()
