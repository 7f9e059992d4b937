#!/usr/bin/env python
"""bench.py — KL-shell Jacobian + residual assembly on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nel 576] [--material svk|nh|mr|nh_c|mr_c]

A "step" is one pass of the hot path over one displacement state: one Jacobian assembly K(x) plus one
residual assembly R(x) of the named workload (the two closures every Newton / arc-length iteration calls,
reference: src/gsStaticSolvers/gsStaticNewton.hpp:160-191).  value = quadrature points processed per second
by all ranks (each step integrates every quadrature point of the mesh in both assemblies; the unit counts
a point once per step).

  value : inputs resident in HBM, device-side calls (kl_jacobian_device + kl_residual_device), CUDA events
  e2e   : the reference-facing host-pointer calls (kl_jacobian + kl_residual) with pinned HOST buffers; the
          H2D copy of x and the D2H copy of all matrix values and the residual are inside the timed region
  N>1   : one process per GPU; every rank assembles its own replica at its own displacement state (the way
          gsAPALM workers own one arc-length interval each, benchmarks/benchmark_Frustrum_APALM.cpp:391-458);
          no data-path collective; weak scaling.
  --impl reference : the CPU path (oracle port, OpenMP over all host cores) on a bounded sample of the same
          workload — the real gismo/gsKLShell assembler cannot be built in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gsstructuralanalysis_b200 import workloads as W  # noqa: E402
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR  # noqa: E402

MATS = {"svk": (KL_MAT_SVK, False), "nh": (KL_MAT_NH, False), "mr": (KL_MAT_MR, False), "nh_c": (KL_MAT_NH, True),
        "mr_c": (KL_MAT_MR, True)}
METRIC = "KL-shell Jacobian+residual assembly throughput (quadrature points per second per J+R step) at 1M DOF"
UNIT = "quad-pts/s"


def make_problem(nel, material):
    mat, comp = MATS[material]
    pr = W.roof(nel, 3)
    pr.material, pr.compressible = mat, comp
    if mat != KL_MAT_SVK:
        pr.nu = 0.45 if comp else 0.5
    return pr


_REAL_STDOUT = None


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def flops_per_qp(p, material):
    """FP64 operations of the Jacobian kernel per quadrature point for the algorithm of DESIGN.md §5 (FMA = 2 flops):
    phase 3 (upper-triangle tiles, sum-factorised): tiles * [ (p+1)*45 + 27*(p+1) ] FMA per fixed-q1 column
    phase 2 (Z_j = T.d_j):  478 flops per (basis function, point) for the linear law (B = 0), 555 with the
                             membrane-bending coupling block (hyperelastic laws)
    p = 3 linear law: 5760 + 16*478 = 13408, which is what ncu counts as executed
    (sm__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on, profiles/r1_r1i_summary.txt)."""
    nloc = (p + 1) ** 2
    tiles = (p + 1) * (p + 1) * (p + 2) // 2
    ph3 = 2 * tiles * ((p + 1) * 45 + 27 * (p + 1)) / (p + 1)     # per point: one column has p+1 points
    ph2 = (478 if material == "svk" else 555) * nloc
    return int(ph3 + ph2)


def bind_to_gpu_numa_node(index):
    """Restrict this rank to the CPUs NVML reports as local to the GPU, so that pinned allocations are NUMA-local."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({cpus[0]}..{cpus[-1]})"
    except Exception as e:      # affinity is an optimisation only
        return f"unbound ({type(e).__name__})"
    return "unbound"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _nvml(self):
        """In-process NVML polling (about 200 samples per second): several samples fall inside a 70 ms timed region."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = (0x8, 0x40, 0x20, 0x4)      # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self._stop.is_set():
            sm, r = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), reasons(h)
            self.samples.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits])
            self._stop.wait(0.005)

    def _run(self):
        try:
            self._nvml()
            return
        except Exception:       # no NVML binding: fall back to polling nvidia-smi
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 3 + k and s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def run_reference(args):
    """CPU arm: oracle port with OpenMP on all host cores, bounded sample (coarser mesh of the same workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import Oracle
    nel = args.ref_nel
    pr = make_problem(nel, args.material)
    orc = Oracle(pr, threads=os.cpu_count())     # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    cores = orc.threads
    x = W.displacement_state(orc.n_dofs, args.scale * 508.0 / nel)
    vals, r = np.zeros(orc.nnz), np.zeros(orc.n_dofs)
    for _ in range(min(args.warmup, 1)):
        orc.jacobian_residual(x, vals, r)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.jacobian_residual(x, vals, r)
    dt = (time.perf_counter() - t0) / args.steps
    v = orc.n_qp / dt
    sample = f"roof {nel}x{nel} elements ({orc.n_dofs} DOFs, {orc.n_qp} quadrature points) per step, same material/BCs"
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "reference_kind": "oracle port of the gsKLShell algorithm (OpenMP); "
                   "the real gismo+gsKLShell assembler is not buildable here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    return (f"benchmark_Roof shallow Scordelis-Lo roof (configs[1]), degree 3, {args.nel}x{args.nel} elements, "
            f"material={args.material}, t=6.35, N/S edges fixed, x = {args.scale}*h*U(-1,1)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nel", type=int, default=576)
    ap.add_argument("--ref-nel", type=int, default=96)
    ap.add_argument("--material", default="svk", choices=list(MATS))
    ap.add_argument("--scale", type=float, default=0.002, help="displacement amplitude as a fraction of the element size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--separate-calls", action="store_true", help="device leg: kl_jacobian_device + kl_residual_device back to back instead of kl_assemble_device")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "strips"],
                    help="N>1: independent replicas (weak scaling, default) or ONE matrix split into element-row strips with the halo exchange (strong scaling)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything libraries print (NCCL banner, ...) is diverted to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)     # pinned host buffers of the e2e leg end up next to this GPU's PCIe root
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gsstructuralanalysis_b200 import build as kbuild, capi
    if rank == 0:
        kbuild.build()
    if world > 1:
        dist.barrier()
    from gsstructuralanalysis_b200.ops import ShellAssembler

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pr = make_problem(args.nel, args.material)
    t_setup = time.perf_counter()
    asm = ShellAssembler(pr, device=local)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    n, nnz, nqp = asm.n_dofs, asm.nnz, asm.n_qp
    h = 508.0 / args.nel
    x_host = W.displacement_state(n, args.scale * h, seed=20240607 + rank)
    x_dev = torch.from_numpy(x_host).cuda()
    r_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    strips = args.mode == "strips" and world > 1
    if strips:
        from gsstructuralanalysis_b200.parallel import plan_strips, exchange_halo, DevicePointerView
        n1_, n2_ = pr.surface.n
        plan = plan_strips(n1_, n2_, 3, n2_ - 3, pr.dof_map, pr.n_free, world, rank)
        asm.set_strip(plan.e2_begin, plan.e2_end)
        vals_view = DevicePointerView(asm.values_device_ptr(), nnz).tensor()
        outer_h, _ = asm.pattern()
        x_host = W.displacement_state(n, args.scale * h, seed=20240607)      # one state, one matrix
        x_dev = torch.from_numpy(x_host).cuda()

    def step_device():
        if not strips and not args.separate_calls:
            # one Jacobian + one residual at the same state through the fused entry (internal force integrated by the point kernel)
            asm.assemble_device(x_dev.data_ptr(), r_dev.data_ptr(), 1.0, -1.0, stream)
            return
        asm.jacobian_device(x_dev.data_ptr(), stream)
        if strips:
            # partial internal force of the strip; the owner adds F_ext after the exchange (not timed: one axpy)
            asm.residual_device(x_dev.data_ptr(), r_dev.data_ptr(), 0.0, 1.0, stream)
            exchange_halo(plan, outer_h, vals_view, r_dev, dist)
        else:
            asm.residual_device(x_dev.data_ptr(), r_dev.data_ptr(), 1.0, -1.0, stream)

    launches0 = asm.kernel_launches()
    for _ in range(max(args.warmup, 3)):
        step_device()
    if asm.check(stream) != 0:
        raise SystemExit("assembly failed: " + capi.lib().kl_last_error().decode())
    launches_per_step = (asm.kernel_launches() - launches0) // max(args.warmup, 3)

    # ---- timed region: K steps, device resident, CUDA events, max over ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    jac_ms = []
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_device()
        e1.record()
        barrier()
        total_ms = e0.elapsed_time(e1)
        # dominant kernel alone (events recorded around the launch inside the library), same stream
        for _ in range(min(args.steps, 5)):
            asm.jacobian_device(x_dev.data_ptr(), stream)
            ms = C.c_float()
            capi.check(asm.L.kl_jacobian_kernel_ms(asm.h, C.byref(ms)))
            jac_ms.append(ms.value)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = (1 if strips else world) * nqp / (ms_per_step * 1e-3)
    jac_kernel_ms = float(np.mean(jac_ms))

    # ---- e2e: host-pointer closures with pinned host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e and not strips:
        vals_pinned = torch.empty(nnz, dtype=torch.float64).pin_memory()
        vals_np = vals_pinned.numpy()
        asm._values = vals_np
        xin = torch.from_numpy(x_host).pin_memory().numpy()
        r_pinned = torch.empty(n, dtype=torch.float64).pin_memory().numpy()     # the solver's result vector, reused every call
        for _ in range(2):
            ok, _ = asm.jacobian(xin)
            ok2, _ = asm.residual(xin, out=r_pinned)
            assert ok and ok2
        barrier()
        t0 = time.perf_counter()
        ksteps = max(2, min(args.steps, 5))
        for _ in range(ksteps):
            ok, K = asm.jacobian(xin)
            tj = asm.last_timing()
            ok2, r = asm.residual(xin, out=r_pinned)
            assert ok and ok2
        barrier()
        dt = (time.perf_counter() - t0) / ksteps
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": world * nqp / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * nnz + 8 * n,
               "ms_per_step": dt * 1e3, "jacobian_breakdown_ms": tj, "steps": ksteps}

    # ---- device-resident linear solve on the matrix just assembled (not part of `value`):
    #      a bounded number of Jacobi-PCG iterations, timed by CUDA events inside the library
    solver = None
    if rank == 0 and not strips:
        try:
            asm.jacobian_device(x_dev.data_ptr(), stream)
            torch.cuda.synchronize()
            xs_, its_, err_ = asm.cg_solve(asm.force(), tol=1e-30, max_iter=96)
            tcg = asm.cg_last_timing()
            regular_bytes = 8 * nnz + 3 * 8 * n          # values + x gather + y (row indices are arithmetic for regular columns)
            solver = {"kind": "Jacobi-PCG = gsSparseSolver CGDiagonal, device resident", "iterations_timed": its_,
                      "ms_per_iteration": tcg["iter_ms"], "algorithmic_GBps": regular_bytes / (tcg["iter_ms"] * 1e-3) / 1e9,
                      "pcie_bytes_per_solve": 2 * 8 * n}
        except Exception as exc:      # the follower-pressure tangent is unsymmetric: CG is refused
            solver = {"unavailable": str(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (Jacobian)
    peak = C.c_double()
    pms = C.c_float()
    capi.check(asm.L.kl_measure_fp64_peak(local, C.byref(peak), C.byref(pms)))
    fpq = flops_per_qp(3, args.material)
    ncp = pr.surface.n[0] * pr.surface.n[1]
    bytes_alg = 8 * nnz + 2 * 24 * ncp
    hbm_peak = 6453.1
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json hbm_gbs; FP64 peak measured live by kl_measure_fp64_peak (DFMA chain kernel)"
    except Exception:
        peak_src = "fallback 6453.1 GB/s (MEASURED_PEAKS.json absent on this box); FP64 peak measured live"
    achieved_tf = fpq * nqp / (jac_kernel_ms * 1e-3) / 1e12
    traffic = None      # dram__bytes_read+write of the kernel from the committed ncu --set full capture of this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if args.nel == 576 and args.material == "svk":
            traffic = tj["traffic"]
    except Exception:
        pass
    roofline = {"kernel": "k_jacobian<3>", "bound": "fp64", "achieved": achieved_tf, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved_tf / peak.value, "traffic": traffic, "kernel_ms": jac_kernel_ms, "flops_per_qp": fpq,
                "note": "FP64 flops are the binding roofline of the fused assembly (13.4 kflop vs 224 B per point); ncu shows the "
                        "kernel limited by the L1/LSU pipe (87 % of peak) with the FP64 pipe 38 % busy, and its time follows the number "
                        "of resident CTAs (profiles/r1_s2_jacobian_summary.txt, profiles/r1_ablation.txt) - DESIGN.md sections 5 and 8",
                "hbm": {"algorithmic_bytes": bytes_alg, "bytes_per_qp": bytes_alg / nqp,
                        "achieved": bytes_alg / (jac_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_alg / (jac_kernel_ms * 1e-3) / 1e9 / hbm_peak},
                "peak_source": peak_src}

    # ---- CPU baseline on a bounded sample (oracle port; the checker timed, never shipped)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle.binding import Oracle
        os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every host core back
        nel_s = args.ref_nel
        prs = make_problem(nel_s, args.material)
        orc = Oracle(prs, threads=os.cpu_count())
        xs = W.displacement_state(orc.n_dofs, args.scale * 508.0 / nel_s)
        vals, rr = np.zeros(orc.nnz), np.zeros(orc.n_dofs)
        orc.jacobian_residual(xs, vals, rr)
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or time.perf_counter() - t0 < 10.0:
            orc.jacobian_residual(xs, vals, rr)
            reps += 1
            if time.perf_counter() - t0 > 30.0:
                break
        dtc = (time.perf_counter() - t0) / reps
        cpu = {"value": orc.n_qp / dtc, "unit": UNIT, "cores": orc.threads, "kind": "port",
               "sample": f"roof {nel_s}x{nel_s} elements ({orc.n_qp} quadrature points) x {reps} J+R steps, OpenMP oracle"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strips else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "n_dofs": n, "nnz": nnz, "elements": asm.n_elements, "quad_points": nqp,
                   "l2": "matrix values (8*nnz bytes = %.2f GB) exceed the 126 MB L2 every step" % (8 * nnz / 1e9),
                   "multi_gpu": ("one matrix in element-row strips, point-to-point halo exchange of the interface columns" if strips else
                                 "one replica per GPU at its own displacement state (APALM interval style), no collective"),
                   "setup_s": t_setup, "cpu_affinity": numa,
                   "step": ("kl_jacobian_device + kl_residual_device" if (strips or args.separate_calls) else
                            "kl_assemble_device: one Jacobian + one residual at the same state, internal force integrated by the point kernel")},
        "clocks": clk.summary(),
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "jacobian_ms": jac_kernel_ms,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "linear_solve": solver,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
