"""Known-answer problems the reference itself holds for this path:
   - uniaxial tension, unittests/gsStaticSolver_test.cpp:108-418 (lateral stretch = sqrt(J/lambda), J tabulated at
     :358-378 for the compressible laws, J = 1 for the incompressible ones; tolerance 1e-7 at :415)
   - Scordelis-Lo roof, filedata/pde/kirchhoff_shell_scordelis.xml:6-12,80-85,104-107 (reference deflection 0.30024)
Both are solved with a plain Newton iteration on top of ANY assembler exposing jacobian(x)/residual(x) in the
closure shapes of gsStructuralAnalysisOps (oracle or GPU path)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from gsstructuralanalysis_b200 import geometry as G
from gsstructuralanalysis_b200.problem import (ShellProblem, BoundaryConditions, KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR,
                                                KL_BC_DIRICHLET, WEST, EAST, SOUTH, NORTH)

UAT_J = {(KL_MAT_NH, False): 1.0, (KL_MAT_MR, False): 1.0, (KL_MAT_NH, True): 1.105598565, (KL_MAT_MR, True): 1.099905842}


def uat_problem(material, compressible, build_dofmap, stretch=2.0):
    """Unit square, degree 1 elevated once, refined once; x fixed on west, x = stretch-1 on east, y fixed on south
    (unittests/gsStaticSolver_test.cpp:141-166).  The reference uses the planar membrane assembler <2,real_t,false>;
    here the same sheet lives in 3-D with z held on the boundary and a negligible bending stiffness (t = 1e-3)."""
    mu = 1.5e6
    nu = 0.45 if compressible else 0.5
    s = G.plate(1.0, 1.0).degree_elevate(1).uniform_refine(1)
    bc = BoundaryConditions()
    bc.add_condition(WEST, KL_BC_DIRICHLET, 0).add_condition(EAST, KL_BC_DIRICHLET, 0).add_condition(SOUTH, KL_BC_DIRICHLET, 1)
    for side in (WEST, EAST, SOUTH, NORTH):
        bc.add_condition(side, KL_BC_DIRICHLET, 2)
    pr = ShellProblem(s, bc, material=material, compressible=compressible, E=2 * mu * (1 + nu), nu=nu, thickness=1e-3, mr_ratio=7.0)
    pr.number_dofs(build_dofmap)
    # Dirichlet values: displacement (stretch-1) in x on the east boundary
    n1, n2 = s.n
    ncp = n1 * n2
    fv = np.zeros(pr.n_fixed)
    for i2 in range(n2):
        g = pr.dof_map[0 * ncp + (n1 - 1) + n1 * i2]
        fv[g - pr.n_free] = 1.0
    pr.fixed_values = fv
    return pr, stretch - 1.0


def newton(make_assembler, pr, load_steps, tol=1e-11, max_it=30, scale_fixed=None):
    """Displacement- or load-controlled Newton: K du = R (gsStaticNewton.hpp:160-191)."""
    x = None
    base_fixed = None if pr.fixed_values is None else pr.fixed_values.copy()
    asm = None
    for s in load_steps:
        if scale_fixed is not None:
            pr.fixed_values = base_fixed * s * scale_fixed
            if asm is not None and hasattr(asm, "close"):
                asm.close()
            asm = make_assembler(pr)
        elif asm is None:
            asm = make_assembler(pr)
        if x is None:
            x = np.zeros(asm.n_dofs)
        for it in range(max_it):
            ok, r = asm.residual(x)
            assert ok
            ok, K = asm.jacobian(x)
            assert ok
            K = K.to_scipy() if hasattr(K, "to_scipy") else K
            dx = spla.spsolve(sp.csc_matrix(K), r)
            x = x + dx
            if np.linalg.norm(dx) <= tol * max(np.linalg.norm(x), 1e-30):
                break
        else:
            raise AssertionError("Newton did not converge")
    return asm, x


def uat_lateral_stretch(pr, x):
    """lambda_2 = 1 + u_y on the north edge (homogeneous deformation)."""
    n1, n2 = pr.surface.n
    ncp = n1 * n2
    g = pr.dof_map[1 * ncp + 0 + n1 * (n2 - 1)]
    return 1.0 + x[g]


def uat_analytical_cauchy_stress(material, compressible, lam=2.0):
    """San of UAT_analytical (unittests/gsStaticSolver_test.cpp:355-385; mu = 1.5e6, Ratio = 7, nu = 0.45 when compressible)."""
    mu, ratio = 1.5e6, 7.0
    c2 = 1.0 / (ratio + 1.0)
    c1 = 1.0 - c2
    if not compressible:
        if material == KL_MAT_NH:
            return mu * (lam * lam - 1.0 / lam)
        return -mu * (c2 * lam * lam + c2 / lam + c1) / lam + lam * (c1 * lam * mu + 2 * c2 * mu)
    nu = 0.45
    K = 2 * mu * (1 + nu) / (3 - 6 * nu)
    J = UAT_J[(material, True)]
    if material == KL_MAT_NH:
        return lam * (0.5 * mu * (-(2 * (lam ** 2 + 2 * J / lam)) / (3 * J ** (2. / 3.) * lam) + 2 * lam / J ** (2. / 3.))
                      + 0.25 * K * (2 * J ** 2 / lam - 2. / lam)) / J
    return lam * (0.5 * c1 * mu * (-(2 * (lam ** 2 + 2 * J / lam)) / (3 * J ** (2. / 3.) * lam) + 2 * lam / J ** (2. / 3.))
                  + 0.5 * c2 * mu * (-(4 * (2 * lam * J + J ** 2 / lam ** 2)) / (3 * J ** (4. / 3.) * lam) + 4 / J ** (1. / 3.))
                  + 0.25 * K * (2 * J ** 2 / lam - 2 / lam)) / J


def uat_cauchy_stress(make_assembler, pr, x, build_dofmap):
    """S = sideForce / (thickness lambda(0) lambda(2)) as in UAT_numerical (unittests/gsStaticSolver_test.cpp:317-324): the
    force on the tension boundary is the sum of the x-reactions on the east edge.  They are read from the residual of the same
    sheet with the east x-DoFs left free (no external load, so residual = -F_int), evaluated at the converged state; by
    isotropy the thickness stretch lambda(2) equals the lateral in-plane stretch lambda(0)."""
    s = pr.surface
    n1, n2 = s.n
    ncp = n1 * n2
    bc = BoundaryConditions()
    bc.add_condition(WEST, KL_BC_DIRICHLET, 0).add_condition(SOUTH, KL_BC_DIRICHLET, 1)
    for side in (WEST, EAST, SOUTH, NORTH):
        bc.add_condition(side, KL_BC_DIRICHLET, 2)
    pb = ShellProblem(s, bc, material=pr.material, compressible=pr.compressible, E=pr.E, nu=pr.nu, thickness=pr.thickness,
                      mr_ratio=pr.mr_ratio)
    pb.number_dofs(build_dofmap)
    xb = np.zeros(pb.n_free)
    for c in range(3):
        for i in range(ncp):
            gb = pb.dof_map[c * ncp + i]
            if gb >= pb.n_free:
                continue
            ga = pr.dof_map[c * ncp + i]
            xb[gb] = x[ga] if ga < pr.n_free else pr.fixed_values[ga - pr.n_free]
    asm = make_assembler(pb)
    ok, r = asm.residual(xb)
    assert ok
    east = [pb.dof_map[0 * ncp + (n1 - 1) + n1 * i2] for i2 in range(n2)]
    side_force = -sum(r[g] for g in east)
    lam0 = uat_lateral_stretch(pr, x)
    return side_force / (pr.thickness * lam0 * lam0)


def scordelis_lo_problem(nel, build_dofmap):
    """Classic roof R=25, L=50, 40 deg, E=4.32e8, nu=0, t=0.25, gravity load 90 per unit area, rigid diaphragms at the
    curved ends (filedata/pde/kirchhoff_shell_scordelis.xml:6-12,80-85)."""
    s0 = G.scordelis_lo_roof_classic()
    s = s0.respace((3, 3), (G.open_uniform_knots(3, nel), G.open_uniform_knots(3, nel)))
    bc = BoundaryConditions()
    # u: along the length (x); rigid diaphragms at u=0 and u=1 hold y and z
    for side in (WEST, EAST):
        bc.add_condition(side, KL_BC_DIRICHLET, 1).add_condition(side, KL_BC_DIRICHLET, 2)
    bc.add_corner_value(0, 0)   # remove the rigid translation along x
    pr = ShellProblem(s, bc, material=KL_MAT_SVK, E=4.32e8, nu=0.0, thickness=0.25, body_force=(0.0, 0.0, -90.0))
    pr.number_dofs(build_dofmap)
    return pr


def scordelis_lo_deflection(pr, x):
    """vertical displacement at the middle of the free edge (u=0.5, v=0)."""
    from gsstructuralanalysis_b200.geometry import basis_matrix
    s = pr.surface
    n1, n2 = s.n
    ncp = n1 * n2
    B1 = basis_matrix(s.p[0], s.U[0], np.array([0.5]))[0]
    uz = 0.0
    for i1 in range(n1):
        g = pr.dof_map[2 * ncp + i1 + n1 * 0]
        if g < pr.n_free:
            uz += B1[i1] * x[g]
    return uz
