"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/kl_shell.h declares, numbers DoFs like the oracle, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gsstructuralanalysis_b200 import capi, build as kbuild, workloads as W
from gsstructuralanalysis_b200.problem import (BoundaryConditions, KL_BC_DIRICHLET, KL_BC_CLAMPED, KL_BC_COLLAPSED,
                                                WEST, EAST, SOUTH, NORTH, SW, NE, c_int_p)
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    kbuild.build()
    return capi.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "kl_shell.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(kl_[a-z_0-9]+)\s*\(", hdr))
    assert names == set(capi.SYMBOLS), names ^ set(capi.SYMBOLS)
    for n in names:
        assert hasattr(L, n), n


def _bcs():
    out = []
    b = BoundaryConditions()
    out.append(("free", b))
    b = BoundaryConditions()
    for c in range(4):
        b.add_corner_value(c)
    out.append(("corners", b))
    b = BoundaryConditions().add_condition(NORTH, KL_BC_DIRICHLET).add_condition(SOUTH, KL_BC_DIRICHLET)
    out.append(("roof", b))
    out.append(("balloon", W.balloon(4).bc))
    out.append(("tension", W.tension_sheet(4).bc))
    out.append(("frustrum", W.frustrum(4).bc))
    b = BoundaryConditions().add_condition(WEST, KL_BC_CLAMPED).add_condition(SOUTH, KL_BC_COLLAPSED, 1).add_corner_value(NE, 2)
    out.append(("mixed", b))
    return out


@pytest.mark.parametrize("name,bc", _bcs())
@pytest.mark.parametrize("n1,n2", [(5, 5), (7, 4), (11, 9)])
def test_dofmap_matches_oracle(L, name, bc, n1, n2):
    O = ob.lib()
    cb = bc.to_c()
    m1 = np.zeros(3 * n1 * n2, dtype=np.int32); m2 = np.zeros_like(m1)
    f1, x1, f2, x2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    assert L.kl_build_dofmap(n1, n2, C.byref(cb), m1.ctypes.data_as(c_int_p), C.byref(f1), C.byref(x1)) == 0
    assert O.klo_build_dofmap(n1, n2, C.byref(cb), m2.ctypes.data_as(c_int_p), C.byref(f2), C.byref(x2)) == 0
    assert (f1.value, x1.value) == (f2.value, x2.value)
    assert np.array_equal(m1, m2)
    # every free index used, component-major monotone for plain DoFs
    assert set(m1[m1 < f1.value]) == set(range(f1.value))


def test_dofmap_numbering_rules(L):
    """component-major; eliminated after all free."""
    n1 = n2 = 5
    bc = BoundaryConditions().add_condition(WEST, KL_BC_DIRICHLET, 0)
    cb = bc.to_c()
    m = np.zeros(3 * n1 * n2, dtype=np.int32)
    f, x = C.c_int32(), C.c_int32()
    L.kl_build_dofmap(n1, n2, C.byref(cb), m.ctypes.data_as(c_int_p), C.byref(f), C.byref(x))
    assert f.value == 75 - 5 and x.value == 5
    m = m.reshape(3, n2, n1)
    assert (m[0, :, 0] >= f.value).all()
    assert m[0, 0, 1] == 0 and m[1, 0, 0] == 20 and m[2, 0, 0] == 45


@pytest.mark.skipif(os.environ.get("KL_HAVE_GPU") == "1", reason="GPU box")
def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pr = W.tutorial_paraboloid(2)
    pr.number_dofs(L.kl_build_dofmap)
    P, keep = pr.to_c()
    h = C.c_void_p()
    rc = L.kl_create(C.byref(P), -1, C.byref(h))
    assert rc == -6 and not h.value
    assert b"no CPU fallback" in L.kl_last_error()


# ---- solid path (include/ks_solid.h) ------------------------------------------------------------------------------------
def test_solid_header_symbols_exported(L):
    from gsstructuralanalysis_b200 import solid as S
    hdr = open(os.path.join(ROOT, "include", "ks_solid.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ks_[a-z_0-9]+)\s*\(", hdr))
    assert names == set(S.SYMBOLS), names ^ set(S.SYMBOLS)
    for n in names:
        assert hasattr(L, n), n


def test_solid_dofmap_matches_oracle(L):
    from gsstructuralanalysis_b200 import solid as S
    from oracle.binding_solid import lib as olib
    S._bind(L)
    for bc in (S.SolidBC(), S.SolidBC().add_condition(S.KS_WEST).add_condition(S.KS_BACK, 1),
               S.SolidBC().add_corner_value(0).add_corner_value(7, 2).add_condition(S.KS_SOUTH, 0)):
        n1, n2, n3 = 4, 3, 5
        m1, m2 = np.zeros(3 * n1 * n2 * n3, dtype=np.int32), np.zeros(3 * n1 * n2 * n3, dtype=np.int32)
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int(), C.c_int()
        cb = bc.to_c()
        assert L.ks_build_dofmap(n1, n2, n3, C.byref(cb), m1.ctypes.data_as(c_int_p), C.byref(a), C.byref(b)) == 0
        assert olib().kso_build_dofmap(n1, n2, n3, C.byref(cb), m2.ctypes.data_as(c_int_p), C.byref(c), C.byref(d)) == 0
        assert np.array_equal(m1, m2) and a.value == c.value and b.value == d.value
        assert sorted(m1.tolist()) == list(range(3 * n1 * n2 * n3))


def test_solid_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gsstructuralanalysis_b200 import solid as S
    from gsstructuralanalysis_b200.capi import KLError
    with pytest.raises(KLError) as ei:
        S.SolidAssembler(S.SolidProblem(S.brick(nels=(2, 1, 1)), S.SolidBC()))
    assert ei.value.rc == -6


def test_headers_are_plain_c_and_link(tmp_path):
    """The boundary is a C ABI: both headers compile as C99 (-pedantic) and a C program links against libkl_shell.so and
    calls the entry points that need no GPU (kl_stress_dim, kl_build_dofmap, kl_create refusing without a device)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "kl_shell.h"
#include "ks_solid.h"
int main(void) {
    kl_bc bc;
    int32_t map[3 * 4 * 4], nf = 0, nx = 0;
    memset(&bc, 0, sizeof bc);
    bc.side[KL_WEST][0] = KL_BC_DIRICHLET;
    if (kl_build_dofmap(4, 4, &bc, map, &nf, &nx) != KL_OK) return 2;
    if (nf + nx != 48 || nx != 4) return 3;
    if (kl_stress_dim(KL_STRESS_PRINCIPAL_STRETCH_DIR) != 9 || kl_stress_dim(KL_STRESS_NTYPES) != 0) return 4;
    printf("C ABI ok: %d free, %d eliminated\n", (int)nf, (int)nx);
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "gsstructuralanalysis_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L" + libdir, "-l:libkl_shell.so", "-Wl,-rpath," + libdir])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "C ABI ok" in r.stdout
