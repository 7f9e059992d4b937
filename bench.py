#!/usr/bin/env python
"""bench.py — KL-shell Jacobian + residual assembly on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nel 576] [--material svk|nh|mr|nh_c|mr_c]

A "step" is one pass of the hot path over one displacement state: one residual assembly R(x) followed by one Jacobian
assembly K(x) at the same state — the two closures every Newton / arc-length iteration calls, in the order the reference
calls them (src/gsStaticSolvers/gsStaticNewton.hpp:160-191).  value = quadrature points processed per second by all
ranks (each step integrates every quadrature point of the mesh in both assemblies; the unit counts a point once per step).

  value : inputs resident in HBM; the two SEPARATE device calls kl_residual_device + kl_jacobian_device that the
          Residual_t / Jacobian_t closures map to (the library detects the repeated state on the device and reuses the
          per-point records; `--fused-call` times the single-call entry kl_assemble_device instead); CUDA events
  e2e   : the reference-facing host-pointer calls (kl_residual + kl_jacobian) with pinned HOST buffers; the H2D copy of x
          and the D2H copy of ALL matrix values and the residual are inside the timed region.  Two more host-buffer legs
          are reported next to it: e2e_lower (lower-triangular view for LDLT consumers, half the bytes) and
          e2e_device_solve (values stay in HBM, the CGDiagonal solve runs there; only vectors cross PCIe)
  N>1   : one process per GPU; every rank assembles its own replica at its own displacement state (the way gsAPALM
          workers own one arc-length interval each, benchmarks/benchmark_Frustrum_APALM.cpp:391-458); no data-path
          collective; weak scaling.  The same run also times ONE matrix split into element-row strips with the NCCL halo
          exchange (strong scaling) and reports it as the `strong` sub-record, with its parity check against the
          single-GPU assembly of the same state.
  --impl reference : the CPU path (oracle port, OpenMP over all host cores) on the SAME workload — the real
          gismo/gsKLShell assembler cannot be built in this image.
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


def flops_per_qp(p, hasB):
    """FP64 operations the degree-3 Jacobian kernel k_jacobian_sw executes per quadrature point: per (basis
    function, point) thread 286 DFMA + 124 DMUL + 25 DADD = 721 flops for the linear law (membrane-bending block B = 0) and
    324 + 135 + 28 -> 811 flops with B (hyperelastic laws), counted in the SASS of the kernel and equal to what ncu reports
    as executed (sm__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on, profiles/r2_sw1_jacobian_summary.txt); times the
    16 basis functions of an element.  This is the kernel's own operation count (FMA = 2), not an estimate of another
    algorithm: the shared-memory kernel of round 1 executed 13 408 flops per point for the same result."""
    assert p == 3
    return 16 * (811 if hasB else 721)


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
        """In-process NVML polling (about 200 samples per second): several samples fall inside a 50 ms timed region."""
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


def workload_name(args):
    return (f"benchmark_Roof shallow Scordelis-Lo roof (configs[1]), degree 3, {args.nel}x{args.nel} elements, "
            f"material={args.material}, t=6.35, N/S edges fixed, x = {args.scale}*h*U(-1,1)")


def run_reference(args):
    """CPU arm: oracle port with OpenMP on all host cores on the SAME mesh, state and warm-up count as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import Oracle
    nel = args.ref_nel or args.nel
    pr = make_problem(nel, args.material)
    orc = Oracle(pr, threads=os.cpu_count())     # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    cores = orc.threads
    x = W.displacement_state(orc.n_dofs, args.scale * 508.0 / nel)
    vals, r = np.zeros(orc.nnz), np.zeros(orc.n_dofs)
    warm = max(args.warmup, 0)
    for _ in range(warm):
        orc.jacobian_residual(x, vals, r)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.jacobian_residual(x, vals, r)
    dt = (time.perf_counter() - t0) / args.steps
    v = orc.n_qp / dt
    sample = (f"roof {nel}x{nel} elements ({orc.n_dofs} DOFs, {orc.n_qp} quadrature points) per step, same material/BCs/state as "
              f"the GPU arm" + ("" if nel == args.nel else f" (coarser than the GPU arm's {args.nel}x{args.nel}: --ref-nel)"))
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args) if nel == args.nel else workload_name(args).replace(f"{args.nel}x{args.nel}", f"{nel}x{nel}"),
                   "n_dofs": orc.n_dofs, "nnz": orc.nnz, "quad_points": orc.n_qp,
                   "reference_kind": "oracle port of the gsKLShell algorithm (OpenMP); the real gismo+gsKLShell assembler is not "
                                     "buildable here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def time_device_steps(torch, step, steps, barrier=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if barrier:
        barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_apalm(n_gpus, args):
    """benchmark_Frustrum_APALM on this box: serial chain of coarse Crisfield steps, then the correction jobs on n_gpus workers."""
    from gsstructuralanalysis_b200 import build as kbuild, capi
    exe = kbuild.build_examples()
    pr = W.frustrum(args.apalm_nel)
    pr.number_dofs(capi.lib().kl_build_dofmap)
    path = os.path.join("/tmp", "kl_apalm_%d.klp" % os.getpid())
    pr.save(path)
    cmd = [exe, path, str(n_gpus), str(args.apalm_steps), "0.05", "2", "1e-3", "2", "1e-12"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    wall = time.perf_counter() - t0
    os.remove(path)
    line = [l for l in r.stdout.splitlines() if l.startswith("APALM ")]
    if r.returncode != 0 or not line:
        return {"error": (r.stdout + r.stderr)[-400:]}
    d = dict(tok.split("=") for tok in line[-1].split() if "=" in tok and not tok.startswith("per_worker"))
    out = {k: (float(v) if "." in v else int(v)) for k, v in d.items()}
    out["jobs_per_worker"] = [int(v) for v in line[-1].split("per_worker=")[1].split()]
    out["process_wall_s"] = wall
    out["workload"] = (f"benchmark_Frustrum_APALM testCase 0 (configs[4]): frustrum {args.apalm_nel}x{args.apalm_nel} elements, degree 3, Mooney-Rivlin, "
                       f"Neumann top edge, gsALMCrisfield (CGDiagonal on the device, Scaling 0), {args.apalm_steps} level-0 steps of dL = 0.05, "
                       "SubIntervals 2, tolerance 1e-3, MaxLevel 2")
    out["what"] = ("t_chain_s: sequential level-0 chain on one GPU; sum_job_s: the correction jobs one after the other; t_parallel_s: the same jobs on "
                   "n_gpus worker threads (one GPU + one assembler replica each); speedup_total = (t_chain + sum_job) / (t_chain + t_parallel)")
    return out


def config_records(torch, capi, args, local, hbm_peak, fp64_peak):
    """one sub-record per BASELINE.json config at benchmark size (Jacobian ms, step ms, roofline fractions),
    each with its parity against the oracle on a coarse twin of the same problem."""
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from oracle.binding import Oracle
    cases = [
        ("S1 plate (configs[0] scaled): tutorial paraboloid, corner-pinned, NH incompressible", lambda n: W.tutorial_paraboloid(n, 3, KL_MAT_NH, False), 576, 2e-3),
        ("S1 plate, St.Venant-Kirchhoff", lambda n: W.tutorial_paraboloid(n, 3, KL_MAT_SVK, False), 576, 2e-3),
        ("S2 roof r=10 (configs[1], 3.16M DOFs), SvK", lambda n: W.roof(n, 3), 1024, 2e-3),
        ("S3 balloon (configs[2]): NURBS eighth sphere, NH incompressible, 4 thickness points, follower pressure", lambda n: W.balloon(n), 576, 2e-3),
        ("S3 cylinder (configs[2]): NH incompressible, Neumann edge traction", lambda n: W.cylinder(n), 576, 2e-3),
        ("S4 tension sheet (configs[3]): Mooney-Rivlin, non-uniform knots", lambda n: W.tension_sheet(n), 576, 1e-5),
        ("S5 frustrum (configs[4]): Mooney-Rivlin, Neumann edge load", lambda n: W.frustrum(n), 576, 2e-3),
    ]
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    for name, mk, nel, amp in cases:
        try:
            pr = mk(nel)
            asm = ShellAssembler(pr, device=local)
            L = max(np.ptp(pr.surface.cp[:, 0]), np.ptp(pr.surface.cp[:, 1]), np.ptp(pr.surface.cp[:, 2]))
            # smooth state + seeded noise small enough for thick / degenerate parametrisations (see workloads.smooth_state)
            hloc = L / nel
            r = torch.empty(asm.n_dofs, dtype=torch.float64, device="cuda")
            x = None

            def step():
                asm.residual_device(x.data_ptr(), r.data_ptr(), 1.0, -1.0, stream)
                asm.jacobian_device(x.data_ptr(), stream)
            # the largest smooth amplitude that is a valid configuration on this mesh (the degenerate pole of the balloon limits it)
            for rel in (1e-3, 1e-4, 1e-5, 0.0):
                x = torch.from_numpy(W.smooth_state(pr, rel * L, noise=amp * min(hloc, hloc * hloc / pr.thickness) * 0.1 * (rel > 0))).cuda()
                for _ in range(3):
                    step()
                if asm.check(stream) == 0:
                    break
            else:
                raise RuntimeError(capi.lib().kl_last_error().decode())
            ms_step = time_device_steps(torch, step, 5)
            jm, pm = C.c_float(), C.c_float()
            asm.residual_device(x.data_ptr(), r.data_ptr(), 1.0, -1.0, stream)
            capi.check(asm.L.kl_points_kernel_ms(asm.h, C.byref(pm)))
            asm.jacobian_device(x.data_ptr(), stream)
            capi.check(asm.L.kl_jacobian_kernel_ms(asm.h, C.byref(jm)))
            hasB = pr.material != KL_MAT_SVK and pr.bending
            fl = flops_per_qp(3, hasB) * asm.n_qp
            ncp = pr.surface.n[0] * pr.surface.n[1]
            by = 8 * asm.nnz + 48 * ncp
            rec = {"workload": name, "elements": f"{nel}x{nel}", "state": f"smooth, amplitude {rel:g}*L + seeded noise", "n_dofs": asm.n_dofs, "nnz": asm.nnz, "quad_points": asm.n_qp,
                   "jacobian_ms": jm.value, "points_residual_ms": pm.value, "step_ms": ms_step, "quad_pts_per_s": asm.n_qp / (ms_step * 1e-3),
                   "fp64_frac": fl / (jm.value * 1e-3) / 1e12 / fp64_peak, "hbm_frac": by / (jm.value * 1e-3) / 1e9 / hbm_peak}
            asm.close()
            # parity on a coarse twin (the oracle finishes in a fraction of a second there)
            prs = mk(12)
            a2, o2 = ShellAssembler(prs, device=local), Oracle(prs)
            Ls = max(np.ptp(prs.surface.cp[:, 0]), np.ptp(prs.surface.cp[:, 1]), np.ptp(prs.surface.cp[:, 2]))
            hs = Ls / 12
            xs = W.smooth_state(prs, 1e-3 * Ls, noise=amp * min(hs, hs * hs / prs.thickness) * 0.1)
            ok, K = a2.jacobian(xs)
            ok2, rr = a2.residual(xs)
            Ko, ro = o2.jacobian_values(xs), o2.residual(xs)
            rec["max_rel_diff_vs_oracle_12x12"] = {"K": float(np.abs(K.values - Ko).max() / np.abs(Ko).max()),
                                                   "R": float(np.abs(rr - ro).max() / max(np.abs(ro).max(), np.abs(o2.force()).max(), 1e-300))}
            a2.close(); o2.close()
            out.append(rec)
        except Exception as exc:       # a config that cannot run is reported, never silently dropped
            out.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})
        torch.cuda.empty_cache()
    return out


def solid_record(torch, args, local, fp64_peak):
    """SURVEY 8a row a9 / 8f rank 3 (north_star: "and gsElasticity solids"): K and rhs of a tri-cubic neo-Hookean block at ~1M DOF in ONE
    device-resident pass (the reference's closures run the full assemble twice), per-kernel times, and the parity of a coarse twin."""
    from gsstructuralanalysis_b200 import solid as S
    from oracle.binding_solid import SolidOracle
    nel = args.solid_nel

    def mk(n):
        v = S.brick(1.0, 1.0, 1.0, degrees=(3, 3, 3), nels=(n, n, n))
        return S.SolidProblem(v, S.SolidBC().add_condition(S.KS_WEST), law=S.KS_LAW_NEO_HOOKE_LN, E=5.0, nu=0.3, tractions=[(S.KS_EAST, (0.0, 0.0, 0.01))])
    asm = S.SolidAssembler(mk(nel), device=local)
    n = asm.n_dofs
    x = torch.from_numpy(1e-3 / nel * np.random.default_rng(20240607).uniform(-1, 1, n)).cuda()
    r = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        asm.assemble_device(x.data_ptr(), r.data_ptr(), True, stream)
    for _ in range(3):
        step()
    if asm.check(stream) != 0:
        raise RuntimeError("solid assembly flagged an invalid state")
    ms = time_device_steps(torch, step, 3)
    ok, K, rr = asm.assemble(x.cpu().numpy())          # host entry: fills the per-kernel event times
    t = asm.last_timing()
    fma_per_elem = 8 * 4 * (128 * 81 + 3456 * 4 + 1152 * 12 + 1152 * 8)      # Z, U, W, acc per (column block, slab), every node pair
    rec = {"workload": f"gsElasticity solid: unit cube, tri-cubic, {nel}^3 elements, neo_hooke_ln, west face fixed, dead traction on the east face",
           "n_dofs": n, "nnz": asm.nnz, "elements": asm.n_elements, "quad_points": asm.n_qp,
           "ms_per_assembly": ms, "quad_pts_per_s": asm.n_qp / (ms * 1e-3), "kernels_ms": t,
           "jacobian_kernel": "k3_jacobian_sw (sliding window along direction 1) + k3_mirror",
           "fp64_frac_full_contraction": 2.0 * fma_per_elem * asm.n_elements / (t["jacobian_ms"] * 1e-3) / 1e12 / fp64_peak,
           "values_GB": 8 * asm.nnz / 1e9, "records_GB": 8 * 100 * asm.n_qp / 1e9,
           "what": "K and rhs in one pass, device resident (ks_assemble_device); kernels_ms from the library's events through the host entry"}
    del K, rr
    asm.close()
    torch.cuda.empty_cache()
    p2 = mk(5)
    a2, o2 = S.SolidAssembler(p2, device=local), SolidOracle(p2)
    xs = 1e-3 * np.random.default_rng(3).standard_normal(a2.n_dofs)
    ok, K2, r2 = a2.assemble(xs)
    Ko, ro = o2.assemble(xs)
    rec["max_rel_diff_vs_oracle_5x5x5"] = {"K": float(np.abs(K2.values - Ko).max() / np.abs(Ko).max()),
                                           "R": float(np.abs(r2 - ro).max() / max(np.abs(ro).max(), np.abs(o2.force()).max(), 1e-300))}
    a2.close()
    return rec


def multipatch_record(torch, dist, args, local, rank, world, barrier, allmax):
    """configs[3] as BASELINE.json frames it ("multipatch row-partitioned assembly at 1/2/4/8 GPUs"): the tension sheet cut into 8 conforming
    patches along the second direction, glued C0 by one DoF mapper (kl_mp_*).  N = 1: all patches on one GPU next to the uncut single
    patch; N > 1: patch -> GPU partition (kl_mp_set_active) + the interface exchange of parallel.exchange_patches."""
    from gsstructuralanalysis_b200.ops import ShellAssembler, MultiPatchAssembler
    from gsstructuralanalysis_b200.parallel import plan_patches, exchange_patches, DevicePointerView, value_ranges
    npatch, nel = 8, args.nel
    base = W.tension_sheet(nel)
    single, multi, cps = W.cut(base, [], [k / npatch for k in range(1, npatch)])
    stream = torch.cuda.current_stream().cuda_stream
    asm = MultiPatchAssembler(multi, device=local)
    n, nnz, nqp = asm.n_dofs, asm.nnz, asm.n_qp
    Lx = float(np.ptp(base.surface.cp[:, 0]))
    hloc = Lx / nel
    # one smooth state for every rank, expressed through the uncut patch and permuted to the multi-patch numbering
    single.number_dofs(asm.L.kl_build_dofmap)
    perm = W.dof_permutation(single, multi, cps)
    xs = W.smooth_state(single, 1e-3 * Lx, noise=1e-5 * min(hloc, hloc * hloc / base.thickness) * 0.1)
    x = torch.from_numpy(np.ascontiguousarray(xs[perm])).cuda()
    r = torch.zeros(n, dtype=torch.float64, device="cuda")
    vals = DevicePointerView(asm.values_device_ptr(), nnz).tensor()
    outer_h, _ = asm.pattern()

    def whole():
        asm.residual_device(x.data_ptr(), r.data_ptr(), 1.0, -1.0, stream)
        asm.jacobian_device(x.data_ptr(), stream)
    for _ in range(3):
        whole()
    if asm.check(stream) != 0:
        raise RuntimeError("multipatch: " + asm.L.kl_last_error().decode())
    ms_whole = time_device_steps(torch, whole, args.steps)
    rec = {"workload": f"S4 tension sheet (configs[3]) {nel}x{nel} elements cut into {npatch} conforming patches of {nel}x{nel // npatch} elements, Mooney-Rivlin",
           "n_dofs": n, "nnz": nnz, "quad_points": nqp, "patches": npatch, "interface_dofs": int(len(asm.interface_dofs())),
           "one_gpu_all_patches_ms": ms_whole, "one_gpu_quad_pts_per_s": nqp / (ms_whole * 1e-3)}
    K_full, R_full = vals.clone(), r.clone()
    if rank == 0:
        # the uncut patch (same function space) on this GPU: what gluing costs, and the identity that pins it
        one = ShellAssembler(single, device=local)
        x1 = torch.from_numpy(np.ascontiguousarray(xs)).cuda()
        r1 = torch.zeros(n, dtype=torch.float64, device="cuda")

        def uncut():
            one.residual_device(x1.data_ptr(), r1.data_ptr(), 1.0, -1.0, stream)
            one.jacobian_device(x1.data_ptr(), stream)
        for _ in range(3):
            uncut()
        rec["uncut_single_patch_ms"] = time_device_steps(torch, uncut, args.steps)
        torch.cuda.synchronize()
        pt = torch.from_numpy(perm).cuda()
        rec["max_rel_diff_vs_uncut_patch"] = {"R": float((R_full - r1[pt]).abs().max() / max(float(r1.abs().max()), 1e-300))}
        one.close()
        del x1, r1
    if world > 1:
        patch_rank = [q * world // npatch for q in range(npatch)]
        plan = plan_patches([p.dof_map for p in multi.patches], multi.n_free, patch_rank, world, rank)
        asm.set_active(plan.active)

        def part():
            asm.residual_device(x.data_ptr(), r.data_ptr(), 1.0, -1.0, stream)
            asm.jacobian_device(x.data_ptr(), stream)
            return exchange_patches(plan, outer_h, vals, r, dist)
        for _ in range(3):
            moved = part()
        ms_part = allmax(time_device_steps(torch, part, args.steps, barrier))
        sK, sR = float(K_full.abs().max()), float(R_full.abs().max())
        eK = max([float((vals[a:b] - K_full[a:b]).abs().max()) for a, b in value_ranges(plan.owned_cols, outer_h) if b > a] + [0.0]) / sK
        eR = max([float((r[c0:c1] - R_full[c0:c1]).abs().max()) for c0, c1 in plan.owned_cols if c1 > c0] + [0.0]) / sR
        eK, eR = allmax(eK), allmax(eR)
        rec["strong"] = {"scaling": "strong", "ms_per_step": ms_part, "value": nqp / (ms_part * 1e-3), "unit": UNIT, "patch_rank": patch_rank,
                         "interface_bytes_received_max": int(allmax(float(moved))),
                         "nccl_op": "batched ncclSend/ncclRecv of the interface columns to their owner rank (batch_isend_irecv), one fused add",
                         "parity_vs_all_patches_on_one_gpu": {"max_rel_K": eK, "max_rel_R": eR, "ok": bool(eK <= 1e-12 and eR <= 1e-12)}}
    asm.close()
    del K_full, R_full, vals
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nel", type=int, default=576)
    ap.add_argument("--ref-nel", type=int, default=0, help="CPU arms: mesh of the sample (0 = the GPU arm's mesh, the default)")
    ap.add_argument("--material", default="svk", choices=list(MATS))
    ap.add_argument("--scale", type=float, default=0.002, help="displacement amplitude as a fraction of the element size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config sub-records (BASELINE.md table)")
    ap.add_argument("--no-strips", action="store_true", help="N>1: skip the strong-scaling strips sub-record")
    ap.add_argument("--no-apalm", action="store_true", help="skip the gsAPALM traversal sub-record")
    ap.add_argument("--no-multipatch", action="store_true", help="skip the multi-patch sub-record")
    ap.add_argument("--no-solid", action="store_true", help="skip the gsElasticity solid sub-record")
    ap.add_argument("--solid-nel", type=int, default=67)
    ap.add_argument("--apalm-nel", type=int, default=64)
    ap.add_argument("--apalm-steps", type=int, default=16)
    ap.add_argument("--fused-call", action="store_true", help="device leg: the single-call entry kl_assemble_device instead of the two closure calls")
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

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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

    def step_device():
        if args.fused_call:
            asm.assemble_device(x_dev.data_ptr(), r_dev.data_ptr(), 1.0, -1.0, stream)
            return
        # the Newton order of the reference: residual first, then the Jacobian at the same state (two closure calls)
        asm.residual_device(x_dev.data_ptr(), r_dev.data_ptr(), 1.0, -1.0, stream)
        asm.jacobian_device(x_dev.data_ptr(), stream)

    launches0 = asm.kernel_launches()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    if asm.check(stream) != 0:
        raise SystemExit("assembly failed: " + capi.lib().kl_last_error().decode())
    launches_per_step = (asm.kernel_launches() - launches0) // warm

    # ---- timed region: K steps, device resident, CUDA events, max over ranks
    jac_ms, pts_ms = [], []
    with ClockSampler(local) as clk:
        ms_per_step = allmax(time_device_steps(torch, step_device, args.steps, barrier))
        # dominant kernel alone (events recorded around the launch inside the library), same stream
        for _ in range(min(args.steps, 5)):
            ms = C.c_float()
            asm.residual_device(x_dev.data_ptr(), r_dev.data_ptr(), 1.0, -1.0, stream)
            capi.check(asm.L.kl_points_kernel_ms(asm.h, C.byref(ms)))
            pts_ms.append(ms.value)
            asm.jacobian_device(x_dev.data_ptr(), stream)
            capi.check(asm.L.kl_jacobian_kernel_ms(asm.h, C.byref(ms)))
            jac_ms.append(ms.value)
    value = world * nqp / (ms_per_step * 1e-3)
    jac_kernel_ms = float(np.mean(jac_ms))

    # ---- e2e: host-pointer closures with pinned host buffers, copies inside the timed region
    e2e = e2e_lower = e2e_dev = None
    if not args.no_e2e:
        xin = torch.from_numpy(x_host).pin_memory().numpy()
        r_pinned = torch.empty(n, dtype=torch.float64).pin_memory().numpy()     # the solver's result vector, reused every call
        ksteps = max(2, min(args.steps, 5))

        def timed_host(fn):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                fn()
            barrier()
            return allmax((time.perf_counter() - t0) / ksteps)

        vals_pinned = torch.empty(nnz, dtype=torch.float64).pin_memory()
        asm._values = vals_pinned.numpy()
        tj = {}

        def host_step():
            ok2, _ = asm.residual(xin, out=r_pinned)
            ok, _ = asm.jacobian(xin)
            tj.update(asm.last_timing())
            assert ok and ok2
        dt = timed_host(host_step)
        e2e = {"value": world * nqp / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * nnz + 8 * n,
               "ms_per_step": dt * 1e3, "jacobian_breakdown_ms": dict(tj), "steps": ksteps,
               "what": "kl_residual + kl_jacobian, all matrix values copied to the caller's pinned array every step"}
        asm._values = None
        del vals_pinned
        # lower-triangular view: what a SimplicialLDLT consumer reads (benchmark_Roof.cpp:359-360)
        asm.pattern_lower()
        low_pinned = torch.empty(asm.nnz_lower, dtype=torch.float64).pin_memory().numpy()

        def host_step_lower():
            ok2, _ = asm.residual(xin, out=r_pinned)
            ok, _ = asm.jacobian_lower(xin, out=low_pinned)
            assert ok and ok2
        dt = timed_host(host_step_lower)
        e2e_lower = {"value": world * nqp / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * asm.nnz_lower + 8 * n,
                     "ms_per_step": dt * 1e3, "steps": ksteps, "what": "kl_residual + kl_jacobian_lower (row >= col entries only)"}
        del low_pinned
        # values stay in HBM, the linear solve runs there (gsSparseSolver-shaped adapter): only vectors cross PCIe
        f_host = torch.from_numpy(asm.force()).pin_memory().numpy()
        cg_iters = 50

        def host_step_dev():
            ok2, _ = asm.residual(xin, out=r_pinned)
            ok, _ = asm.jacobian(xin, fetch=False)
            assert ok and ok2
        dt_asm = timed_host(host_step_dev)

        def host_step_solve():
            host_step_dev()
            asm.cg_solve(f_host, tol=1e-30, max_iter=cg_iters)
        dt = timed_host(host_step_solve)
        e2e_dev = {"value": world * nqp / dt_asm, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * n,
                   "ms_per_step": dt_asm * 1e3, "steps": ksteps,
                   "what": "kl_residual + kl_jacobian(values_host = NULL): the matrix stays on the device for the device CG",
                   "with_solve": {"ms_per_step": dt * 1e3, "cg_iterations_per_step": cg_iters, "h2d_bytes_per_step": 3 * 8 * n,
                                  "d2h_bytes_per_step": 2 * 8 * n, "value": world * nqp / dt,
                                  "what": "the same plus one kl_cg_solve (CGDiagonal, capped at 50 iterations) per step"}}

    # ---- strong scaling: ONE matrix in element-row strips + halo exchange (the path with a real exchange step)
    strong = None
    if world > 1 and not args.no_strips:
        from gsstructuralanalysis_b200.parallel import (plan_strips, exchange_halo, assemble_strip_overlapped, DevicePointerView,
                                                        function_supports, value_ranges)
        x1 = torch.from_numpy(W.displacement_state(n, args.scale * h, seed=20240607)).cuda()      # one state, one matrix
        vals_view = DevicePointerView(asm.values_device_ptr(), nnz).tensor()
        outer_h, _ = asm.pattern()
        # single-GPU assembly of the same state on this rank: the parity reference of the strips path
        asm.jacobian_device(x1.data_ptr(), stream)
        asm.residual_device(x1.data_ptr(), r_dev.data_ptr(), 0.0, 1.0, stream)
        torch.cuda.synchronize()
        K_full, R_full = vals_view.clone(), r_dev.clone()
        n1_, n2_ = pr.surface.n
        plan = plan_strips(n1_, n2_, 3, function_supports(pr.surface.U[1], 3)[2], pr.dof_map, pr.n_free, world, rank, knots2=pr.surface.U[1])
        asm.set_strip(plan.e2_begin, plan.e2_end)
        ex_ms = []

        def strip_step(timed=False):
            if not timed:
                # interface rows first, halo exchange in flight while the rest of the strip is assembled
                return assemble_strip_overlapped(asm, plan, outer_h, vals_view, r_dev, x1.data_ptr(), dist, stream)
            asm.residual_device(x1.data_ptr(), r_dev.data_ptr(), 0.0, 1.0, stream)     # partial internal force of the strip
            asm.jacobian_device(x1.data_ptr(), stream)
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            nb = exchange_halo(plan, outer_h, vals_view, r_dev, dist)
            if timed:
                b.record(); torch.cuda.synchronize(); ex_ms.append(a.elapsed_time(b))
            return nb
        for _ in range(3):
            halo_bytes = strip_step()
        ms_strong = allmax(time_device_steps(torch, strip_step, args.steps, barrier))
        for _ in range(3):
            strip_step(True)
        # parity of the owned columns against the single-GPU assembly (1e-12 of the largest entry)
        sK, sR = float(K_full.abs().max()), float(R_full.abs().max())
        errK = max([float((vals_view[a:b] - K_full[a:b]).abs().max()) for a, b in value_ranges(plan.owned_cols, outer_h) if b > a] + [0.0]) / sK
        errR = max([float((r_dev[c0:c1] - R_full[c0:c1]).abs().max()) for c0, c1 in plan.owned_cols if c1 > c0] + [0.0]) / sR
        errK, errR = allmax(errK), allmax(errR)
        # halo-compute variant: every rank also integrates the p element rows of the previous strip its own control-point
        # rows reach into; its owned columns are then complete without any exchange (about p/rows_per_strip redundant work)
        halo_compute = None
        if plan.compute_begin >= 0:
            asm.set_strip(plan.compute_begin, plan.e2_end)

            def hc_step():
                asm.residual_device(x1.data_ptr(), r_dev.data_ptr(), 0.0, 1.0, stream)
                asm.jacobian_device(x1.data_ptr(), stream)
            for _ in range(3):
                hc_step()
            ms_hc = allmax(time_device_steps(torch, hc_step, args.steps, barrier))
            eK = max([float((vals_view[a:b] - K_full[a:b]).abs().max()) for a, b in value_ranges(plan.owned_cols, outer_h) if b > a] + [0.0]) / sK
            eR = max([float((r_dev[c0:c1] - R_full[c0:c1]).abs().max()) for c0, c1 in plan.owned_cols if c1 > c0] + [0.0]) / sR
            eK, eR = allmax(eK), allmax(eR)
            halo_compute = {"ms_per_step": ms_hc, "value": nqp / (ms_hc * 1e-3), "redundant_element_rows_per_interface": int(plan.e2_begin - plan.compute_begin) if rank > 0 else 3,
                            "parity_vs_single_gpu": {"max_rel_K": eK, "max_rel_R": eR, "ok": bool(eK <= 1e-12 and eR <= 1e-12)},
                            "what": "no exchange at all: each rank assembles its strip plus the element rows of the previous strip that its owned control-point rows reach into"}
        # per-rank fixed cost: an empty strip (launch overheads, memsets of nothing, state upload)
        asm.set_strip(0, 0)
        ms_fixed = allmax(time_device_steps(torch, lambda: (asm.residual_device(x1.data_ptr(), r_dev.data_ptr(), 0.0, 1.0, stream),
                                                            asm.jacobian_device(x1.data_ptr(), stream)), 5, barrier))
        asm.set_strip(0, asm.n_elements // (pr.surface.n[0] - 3))
        reduce_rec = {"ms_per_step": ms_strong, "value": nqp / (ms_strong * 1e-3), "exchange_ms": allmax(float(np.mean(ex_ms))),
                      "halo_bytes_received_max": int(allmax(float(halo_bytes))),
                      "nccl_op": "batched ncclSend/ncclRecv to the neighbour strip (batch_isend_irecv) posted after the interface rows and overlapped with the "
                                 "rest of the strip, one fused add of the received ranges; exchange_ms = the same exchange timed without overlap",
                      "parity_vs_single_gpu": {"max_rel_K": errK, "max_rel_R": errR, "ok": bool(errK <= 1e-12 and errR <= 1e-12)}}
        best_hc = halo_compute is not None and halo_compute["ms_per_step"] < ms_strong
        best = halo_compute if best_hc else reduce_rec
        strong = {"scaling": "strong", "mode": "halo_compute" if best_hc else "halo_reduce", "ms_per_step": best["ms_per_step"],
                  "value": best["value"], "unit": UNIT, "per_rank_fixed_ms": ms_fixed, "speedup_vs_1gpu_step": None,
                  "parity_vs_single_gpu": best["parity_vs_single_gpu"], "halo_reduce": reduce_rec, "halo_compute": halo_compute,
                  "what": "ONE matrix of the same workload in element-row strips, one per rank; owned columns complete on their owner.  halo_reduce: "
                          "interface columns travel to their owner over NCCL, overlapped with the bulk of the strip; halo_compute: every rank also "
                          "integrates the p element rows of the previous strip that its owned control-point rows reach into, no exchange.  The headline "
                          "of this record is the faster of the two (`mode`)"}
        del K_full, R_full

    # ---- multi-patch (configs[3] as framed): 8 glued patches, patch -> GPU partition with the interface exchange when N > 1
    multipatch = None
    if not args.no_multipatch:
        try:
            multipatch = multipatch_record(torch, dist if world > 1 else None, args, local, rank, world, barrier, allmax)
        except Exception as exc:
            if world > 1:
                raise          # a rank that drops out of a collective would hang the others
            multipatch = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- gsAPALM traversal of the frustrum (configs[4]): level-0 chain + correction jobs, one worker thread per GPU of this box
    #      (examples/apalm_dispatch.cpp over include/gsAPALM_b200.h); the other ranks keep their GPUs idle meanwhile
    apalm = None
    if not args.no_apalm:
        flag = os.path.join("/tmp", "kl_apalm_done_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()))
        if rank == 0:
            try:
                apalm = run_apalm(world, args)
            except Exception as exc:
                apalm = {"error": f"{type(exc).__name__}: {exc}"}
            if world > 1:
                open(flag, "w").close()
        else:
            t_wait = time.time()
            while not os.path.exists(flag) and time.time() - t_wait < 900:
                time.sleep(0.2)
        if world > 1:
            dist.barrier()
            if rank == 0 and os.path.exists(flag):
                os.remove(flag)

    # ---- device-resident linear solve on the matrix just assembled (not part of `value`)
    solver = None
    if rank == 0:
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
    hasB = pr.material != KL_MAT_SVK and pr.bending
    fpq = flops_per_qp(3, hasB)
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
        tj_ = json.load(open(os.path.join(ROOT, "profiles", "r3_traffic.json")))
        if args.nel == 576 and args.material == "svk":
            traffic = tj_["traffic"]
    except Exception:
        pass
    roofline = {"kernel": "k_jacobian_sw", "bound": "fp64", "achieved": achieved_tf, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved_tf / peak.value, "traffic": traffic, "kernel_ms": jac_kernel_ms, "flops_per_qp": fpq,
                "share_of_step": jac_kernel_ms / ms_per_step,
                "note": "FP64 flops are the binding roofline of the fused assembly (11.5 kflop executed vs 224 B per point); flops_per_qp is "
                        "the kernel's own executed count (ncu: dfma/dmul/dadd), so frac equals the share of peak DFMA-equivalent issue; ncu: "
                        "FP64 pipe 58 % busy, L1/LSU data pipe 69 % (shuffles, RED scatter, per-point record loads) - profiles/r3_head_jacobian_summary.txt, DESIGN.md section 5",
                "hbm": {"algorithmic_bytes": bytes_alg, "bytes_per_qp": bytes_alg / nqp,
                        "achieved": bytes_alg / (jac_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_alg / (jac_kernel_ms * 1e-3) / 1e9 / hbm_peak},
                "peak_source": peak_src}
    if strong:
        strong["speedup_vs_1gpu_step"] = None      # the driver computes efficiencies from the per-N lines

    configs = None
    if world == 1 and not args.no_configs:
        asm.close()
        torch.cuda.empty_cache()
        configs = config_records(torch, capi, args, local, hbm_peak, peak.value)

    solid = None
    if world == 1 and not args.no_solid:
        try:
            asm.close()
            torch.cuda.empty_cache()
            solid = solid_record(torch, args, local, peak.value)
        except Exception as exc:
            solid = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    # ---- CPU baseline on a bounded sample of the SAME workload (oracle port; the checker timed, never shipped)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle.binding import Oracle
        os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every host core back
        nel_s = args.ref_nel or args.nel
        prs = make_problem(nel_s, args.material)
        orc = Oracle(prs, threads=os.cpu_count())
        xs = W.displacement_state(orc.n_dofs, args.scale * 508.0 / nel_s)
        vals, rr = np.zeros(orc.nnz), np.zeros(orc.n_dofs)
        orc.jacobian_residual(xs, vals, rr)
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or time.perf_counter() - t0 < 10.0:
            orc.jacobian_residual(xs, vals, rr)
            reps += 1
            if time.perf_counter() - t0 > 30.0:
                break
        dtc = (time.perf_counter() - t0) / reps
        cpu = {"value": orc.n_qp / dtc, "unit": UNIT, "cores": orc.threads, "kind": "port", "ms_per_step": dtc * 1e3,
               "sample": f"roof {nel_s}x{nel_s} elements ({orc.n_qp} quadrature points, the GPU arm's mesh and state) x {reps} J+R steps "
                         f"after one warm-up, OpenMP oracle"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "n_dofs": n, "nnz": nnz, "elements": nqp // 16, "quad_points": nqp,
                   "l2": "matrix values (8*nnz bytes = %.2f GB) exceed the 126 MB L2 every step" % (8 * nnz / 1e9),
                   "multi_gpu": "one replica per GPU at its own displacement state (APALM interval style), no collective; `strong` = one matrix in strips",
                   "setup_s": t_setup, "cpu_affinity": numa,
                   "step": ("kl_assemble_device: one Jacobian + one residual at the same state in one call" if args.fused_call else
                            "kl_residual_device then kl_jacobian_device at the same state: the two calls behind the Residual_t / Jacobian_t closures"),
                   "state_amplitude": f"{args.scale}*h; BASELINE.md's 1e-2*L amplitude is covered by the parity tests (tests/test_gpu_parity.py)"},
        "clocks": clk.summary(),
        "e2e": e2e, "e2e_lower": e2e_lower, "e2e_device_solve": e2e_dev,
        "gpu_launches": launches_per_step * args.steps,
        "jacobian_ms": jac_kernel_ms, "points_residual_ms": float(np.mean(pts_ms)),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "linear_solve": solver,
        "strong": strong,
        "multipatch": multipatch,
        "apalm": apalm,
        "configs": configs,
        "solid": solid,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
