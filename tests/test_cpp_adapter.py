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
