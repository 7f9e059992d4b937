"""The header-only C++ adapter (include/gsStructuralAnalysisOps_b200.h) builds with g++ against libkl_shell.so and
drives a Newton solve through Jacobian_t / Residual_t closures (examples/newton_shell.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "newton_shell")


def _build():
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    cmd = ["g++", "-std=c++17", "-O2", "-o", EXE, os.path.join(ROOT, "examples", "newton_shell.cpp"),
           "-L" + os.path.join(ROOT, "gsstructuralanalysis_b200"), "-l:libkl_shell.so",
           "-Wl,-rpath," + os.path.join(ROOT, "gsstructuralanalysis_b200")]
    subprocess.check_call(cmd)


def _problem(tmp_path):
    from gsstructuralanalysis_b200 import workloads as W, capi
    pr = W.tutorial_paraboloid(4)
    pr.point_loads = [((0.5, 0.5), (0.0, 0.0, -1e2))]
    pr.number_dofs(capi.lib().kl_build_dofmap)
    path = os.path.join(tmp_path, "p.klp")
    pr.save(path)
    return path


def test_cpp_example_builds_and_refuses_cpu(tmp_path):
    import torch
    _build()
    r = subprocess.run([EXE, _problem(str(tmp_path))], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    if not torch.cuda.is_available():
        assert "NO_GPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_newton_converges_on_gpu(tmp_path):
    _build()
    r = subprocess.run([EXE, _problem(str(tmp_path)), "25"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "STATUS Success" in r.stdout, r.stdout
    assert "DEVICE_NEWTON Success" in r.stdout, r.stdout
    assert "SPARSE_SOLVER_B200 Success" in r.stdout, r.stdout    # gsSparseSolver-shaped device CG, Device / Lower copy-out modes
    assert "STRETCHES" in r.stdout, r.stdout      # computePrincipalStretches / boundaryForce / evalStress of the adapter


APALM = os.path.join(ROOT, "examples", "apalm_dispatch")


def _build_apalm():
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", APALM, os.path.join(ROOT, "examples", "apalm_dispatch.cpp"),
                           "-L" + os.path.join(ROOT, "gsstructuralanalysis_b200"), "-l:libkl_shell.so",
                           "-Wl,-rpath," + os.path.join(ROOT, "gsstructuralanalysis_b200")])


def _parse(line):
    return dict(tok.split("=") for tok in line.split() if "=" in tok and not tok.startswith("per_worker"))


def test_apalm_queue_semantics_cpu():
    """gsAPALMData pop / submit / storage rules (src/gsALMSolvers/gsAPALMData.hpp:215-252,291-433) with fake workers: on level 1
    the lower error of an interval is 25 % and its two computed sub-intervals are queued on level 2, which is exact."""
    _build_apalm()
    for workers in (1, 3, 8):
        r = subprocess.run([APALM, "--fake", str(workers), "8"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        d = _parse(r.stdout.strip().splitlines()[-1])
        assert int(d["jobs"]) == 8 + 16 and int(d["points"]) == 48 and int(d["maxLevel"]) == 2 and int(d["failed"]) == 0
        per = [int(v) for v in r.stdout.strip().split("per_worker=")[1].split()]
        assert len(per) == workers and sum(per) == 24 and min(per) >= 1


@pytest.mark.gpu
def test_apalm_traversal_on_the_gpus(tmp_path):
    """benchmark_Frustrum_APALM in small: serial level-0 chain of Crisfield steps, then the correction jobs on every GPU of the box."""
    import torch
    _build_apalm()
    from gsstructuralanalysis_b200 import workloads as W, capi
    pr = W.frustrum(12)
    pr.number_dofs(capi.lib().kl_build_dofmap)
    path = os.path.join(str(tmp_path), "f.klp")
    pr.save(path)
    n = torch.cuda.device_count()
    r = subprocess.run([APALM, path, str(n), "6", "0.05", "2", "1e-3", "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    d = _parse(r.stdout.strip().splitlines()[-1])
    assert int(d["jobs"]) >= 6 and int(d["failed"]) == 0 and int(d["points"]) >= 7 + 2 * 6
    assert float(d["lambda_end"]) > 0 and float(d["t_chain_s"]) > 0 and float(d["sum_job_s"]) > 0


SOLID_EXE = os.path.join(ROOT, "examples", "solid_newton")


def _build_solid():
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", SOLID_EXE, os.path.join(ROOT, "examples", "solid_newton.cpp"),
                           "-L" + os.path.join(ROOT, "gsstructuralanalysis_b200"), "-l:libkl_shell.so",
                           "-Wl,-rpath," + os.path.join(ROOT, "gsstructuralanalysis_b200")])


def test_cpp_solid_example_builds_and_refuses_cpu():
    import torch
    _build_solid()
    r = subprocess.run([SOLID_EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "NO_GPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_solid_newton_converges_on_gpu():
    _build_solid()
    r = subprocess.run([SOLID_EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "STATUS Success" in r.stdout, r.stdout
