"""Assembly invariants (SURVEY.md 8(c) iii) for the DEVICE element algebra compiled for the host — no oracle involved:
  * a uniform velocity field with zero pressure, acceleration and body force leaves no residual (VMS fluid, TET4);
  * a linear pressure with the matching body force (hydrostatic state) leaves no momentum residual;
  * the continuity residual sums to the integral of div u over the mesh (the shape functions sum to one);
  * a rigid translation of a solid leaves no internal force, and its tangent annihilates translations;
  * the heat residual vanishes for a uniform temperature without source, and its tangent rows sum to the mass term."""
import ctypes as C

import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements, meshgen
from tests import common
from tests.test_hostmath_cpu import FluidArgs, HostStructArgs, HostHeatArgs, hostmath, _fill_tables  # noqa: F401


def _csr(m):
    from oracle import refbind
    c = refbind.OracleCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
    return c.build_graph(0)


def _fluid(hostmath, m, Ag, Yg, Bf, d, rowPtr, colPtr, dt=0.005):
    eq = abi.fluid_eq(dt)
    w, N, Nx = elements.tables(4)
    A = FluidArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Bf = (k.ctypes.data for k in keep)
    A.e0, A.e1, A.tDof, A.mvMsh, A.nDmn = 0, m.nEl, 4, 0, 1
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    for g in range(4):
        A.w[g] = w[g]
        for a in range(4):
            A.N[g][a] = N[a, g]
            for k in range(3):
                A.Nxi[g][a][k] = Nx[k, a, g]
    A.dmn[0].rho, A.dmn[0].Kd = d.rho, d.K_darcy
    for i in range(3):
        A.dmn[0].f[i] = d.f[i]
    A.dmn[0].mu_i, A.dmn[0].viscType, A.dmn[0].Id, A.dmn[0].isFluid = d.mu_i, d.viscType, -1, 1
    R = np.zeros((m.nNo, 4))
    V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_tet4(C.byref(A), m.nNo, rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                      R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return R.T, V.T


def test_fluid_uniform_flow_and_hydrostatic_state_leave_no_residual(hostmath):
    m = meshgen.cylinder_tet4(3, 4)
    rowPtr, colPtr = _csr(m)
    z = np.zeros((4, m.nNo), order="F")
    Y = z.copy(order="F"); Y[0], Y[1], Y[2] = 1.5, -0.7, 3.0                        # uniform velocity, p = 0
    R, _ = _fluid(hostmath, m, z, Y, np.zeros((3, m.nNo), order="F"), abi.fluid_domain(), rowPtr, colPtr)
    interior = np.ones(m.nNo, bool)
    for k in ("wall", "inlet", "outlet_all"):
        interior[m.faces[k]] = False
    # boundary nodes keep the flux terms of the weak form; interior rows vanish
    assert np.abs(R[:, interior]).max() < 1e-11
    # hydrostatic: u = 0, p = rho g . x, body force g  ->  momentum residual zero in the interior
    d = abi.fluid_domain(rho=1.06, f=(0.3, -0.2, 0.5))
    Y = z.copy(order="F"); Y[3] = d.rho * (0.3 * m.x[0] - 0.2 * m.x[1] + 0.5 * m.x[2])
    R, _ = _fluid(hostmath, m, z, Y, np.zeros((3, m.nNo), order="F"), d, rowPtr, colPtr)
    assert np.abs(R[:, interior]).max() < 1e-10


def test_continuity_residual_sums_to_the_divergence_integral(hostmath):
    """sum_a lR_c(a) = int div u: the stabilisation term int tau_M r_M . grad N_a sums to zero over a because sum_a N_a = 1."""
    m = meshgen.cylinder_tet4(3, 4)
    rowPtr, colPtr = _csr(m)
    z = np.zeros((4, m.nNo), order="F")
    Y = z.copy(order="F")
    Y[0], Y[1], Y[2] = 0.3 * m.x[0], -0.1 * m.x[1] + 0.2 * m.x[2], 0.5 * m.x[2]     # div u = 0.3 - 0.1 + 0.5
    R, _ = _fluid(hostmath, m, z, Y, np.zeros((3, m.nNo), order="F"), abi.fluid_domain(), rowPtr, colPtr)
    w, N, Nx = elements.tables(4)
    vol = 0.0
    for e in range(m.nEl):
        x = m.x[:, m.IEN[:, e]]
        vol += abs(np.linalg.det(x[:, :3] - x[:, [3]])) / 6.0
    assert abs(R[3].sum() - 0.7 * vol) < 1e-10 * vol


def _struct(hostmath, m, Ag, Yg, Dg, Bf, d, rowPtr, colPtr):
    eq = abi.struct_eq(1e-4)
    A = HostStructArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Dg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Dg, A.Bf = (k.ctypes.data for k in keep)
    A.eNoN, A.nEl, A.tDof, A.dof, A.s, A.nFn = m.eNoN, m.nEl, 3, 3, 0, 0
    A.nG = _fill_tables(A, m.eNoN)
    A.dt, A.af, A.am, A.gam, A.beta = eq.dt, eq.af, eq.am, eq.gam, eq.beta
    dm = A.dm
    dm.rho, dm.Kpen, dm.C10, dm.C01 = d.rho, d.Kpen, d.C10, d.C01
    dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
    R = np.zeros((m.nNo, 3))
    V = np.zeros((len(colPtr), 9))
    rc = hostmath.hostmath_struct(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                  R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return R.T, V.T, eq


@pytest.mark.parametrize("kind", ["hex8", "tet4"])
def test_solid_rigid_translation(hostmath, kind):
    m = meshgen.box_hex8(3, 2, 2, (1.0, 1.0, 1.0)) if kind == "hex8" else meshgen.box_tet4(2, 2, 2, (1.0, 1.0, 1.0))
    rowPtr, colPtr = _csr(m)
    z = np.zeros((3, m.nNo), order="F")
    D = z.copy(order="F"); D[0], D[1], D[2] = 0.3, -0.2, 0.1                     # rigid translation: F = I, S = 0
    d = abi.struct_domain(E=1e6, nu=0.4, Kpen=1e6, rho=0.0)                          # rho = 0: no inertia in the tangent
    R, V, eq = _struct(hostmath, m, z, z, D, z, d, rowPtr, colPtr)
    assert np.abs(R).max() < 1e-9
    # K t = 0 for a translation t: the 3x3 blocks of every row sum to zero over the columns
    rows = np.repeat(np.arange(m.nNo), np.diff(rowPtr))
    S = np.zeros((m.nNo, 9))
    np.add.at(S, rows, V.T)
    assert np.abs(S).max() < 1e-9 * np.abs(V).max()


def test_heat_uniform_temperature(hostmath):
    m = meshgen.box_hex8(3, 2, 2, (1.0, 1.0, 1.0))
    rowPtr, colPtr = _csr(m)
    eq, d = abi.heat_eq(0.01, False), abi.heat_domain(False, conductivity=0.7, source=0.0, rho=2.5)
    A = HostHeatArgs()
    Ag = np.zeros((1, m.nNo), order="F"); Yg = np.full((1, m.nNo), 300.0, order="F")
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T), np.ascontiguousarray(Yg.T)]
    A.IEN, A.x, A.Ag, A.Yg = (k.ctypes.data for k in keep)
    A.eNoN, A.nEl, A.tDof, A.s, A.mvMsh, A.fluid = 8, m.nEl, 1, 0, 0, 0
    A.nG = _fill_tables(A, 8)
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    A.dm.rho, A.dm.nu, A.dm.s, A.dm.Id, A.dm.active = d.rho, d.conductivity, 0.0, -1, 1
    R = np.zeros(m.nNo); V = np.zeros(len(colPtr))
    assert hostmath.hostmath_heat(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                  R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p)) == 0
    assert np.abs(R).max() < 1e-12
    # the conductivity part of a row sums to zero (constants are in its kernel); what is left is the lumped mass term:
    # sum over all entries = am rho |Omega|
    assert abs(V.sum() - eq.am * d.rho * 1.0) < 1e-12


def _csr_matvec(rowPtr, colPtr, V, x, dof=3):
    """y = K x for the block-CSR Val (dof*dof, nnz) with blocks row-major; x, y are (dof, nNo)."""
    y = np.zeros_like(x)
    rows = np.repeat(np.arange(len(rowPtr) - 1), np.diff(rowPtr))
    B = V.T.reshape(-1, dof, dof)
    np.add.at(y.T, rows, np.einsum("kij,kj->ki", B, x.T[colPtr]))
    return y


@pytest.mark.parametrize("kind,iso", [("hex8", abi.ISO_NHK), ("hex8", abi.ISO_MR), ("tet4", abi.ISO_NHK), ("tet4_closed_form", abi.ISO_NHK),
                                      ("tet4_closed_form", abi.ISO_STVK)])
def test_solid_tangent_is_the_derivative_of_the_residual(hostmath, kind, iso):
    """For a hyperelastic solid without inertia the assembled tangent is exact: Val = af beta dt^2 dR/dd (sv_struct.cpp:736-825), so
    Val . delta = afu (R(d + e delta) - R(d - e delta)) / (2 e) + O(e^2) for any nodal field delta — the general two-phase algebra
    (hostmath_struct) and the closed-form TET4 algebra (hostmath_tet4) both have to satisfy it."""
    from tests.test_hostmath_cpu import _run_tet4
    m = meshgen.box_hex8(3, 2, 2, (1.0, 1.0, 1.0)) if kind == "hex8" else meshgen.box_tet4(2, 2, 2, (1.0, 1.0, 1.0))
    rowPtr, colPtr = _csr(m)
    rng = np.random.default_rng(8)
    z = np.zeros((3, m.nNo), order="F")
    D = np.asfortranarray(0.02 * rng.standard_normal((3, m.nNo)))
    delta = np.asfortranarray(rng.standard_normal((3, m.nNo)))
    kw = dict(E=1e6, nu=0.4, Kpen=1e6, rho=0.0)
    if iso == abi.ISO_MR:
        kw.update(isoType=iso, C10=1e5, C01=3e4)
    if iso == abi.ISO_STVK:
        kw.update(isoType=iso, C10=2e5, C01=1e5, Kpen=0.0)
    d = abi.struct_domain(**kw)

    def run(Dg):
        if kind == "tet4_closed_form":
            eq = abi.struct_eq(1e-4)

            def fill(dm):
                dm.rho, dm.Kpen, dm.C10, dm.C01 = d.rho, d.Kpen, d.C10, d.C01
                dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
            R, V = _run_tet4(hostmath, m, z, z, Dg, z, None, eq, 0, fill, 0, None, rowPtr, colPtr)
            return R, V, eq
        return _struct(hostmath, m, z, z, Dg, z, d, rowPtr, colPtr)

    R0, V, eq = run(D)
    e = 1e-6
    Rp, _, _ = run(np.asfortranarray(D + e * delta))
    Rm, _, _ = run(np.asfortranarray(D - e * delta))
    afu = eq.af * eq.beta * eq.dt * eq.dt
    lhs = _csr_matvec(rowPtr, colPtr, V, delta)
    rhs = afu * (Rp - Rm) / (2 * e)
    assert np.abs(lhs - rhs).max() < 1e-6 * np.abs(rhs).max()


def test_linear_elasticity_tangent_is_the_derivative_of_the_residual(hostmath):
    """Same identity for the closed-form TET4 routines of the mesh / lElas equation (l_elas_3d is linear in d: the difference
    quotient is exact up to round-off)."""
    from tests.test_hostmath_cpu import _run_tet4
    m = meshgen.box_tet4(2, 2, 2, (1.0, 1.0, 1.0))
    rowPtr, colPtr = _csr(m)
    rng = np.random.default_rng(9)
    z = np.zeros((3, m.nNo), order="F")
    D = np.asfortranarray(0.02 * rng.standard_normal((3, m.nNo)))
    delta = np.asfortranarray(rng.standard_normal((3, m.nNo)))
    eq = abi.lelas_eq(1e-3)

    def fill(dm):
        dm.rho, dm.C10, dm.C01 = 0.0, 1.0e6, 0.3
        dm.Id, dm.isStruct = -1, 1
    run = lambda Dg: _run_tet4(hostmath, m, z, z, Dg, z, None, eq, 1, fill, 0, None, rowPtr, colPtr)   # noqa: E731
    _, V = run(D)
    Rp, _ = run(np.asfortranarray(D + delta))
    Rm, _ = run(np.asfortranarray(D - delta))
    afu = eq.af * eq.beta * eq.dt * eq.dt
    lhs = _csr_matvec(rowPtr, colPtr, V, delta)
    rhs = afu * (Rp - Rm) / 2.0
    assert np.abs(lhs - rhs).max() < 1e-11 * np.abs(rhs).max()


@pytest.mark.parametrize("kind,entry", [("hex8", "hostmath_ustruct"), ("tet4", "hostmath_ustruct"), ("tet4", "hostmath_ustruct_tet4")],
                         ids=["hex8", "tet4", "tet4_closed_form"])
def test_ustruct_displacement_tangent_is_the_derivative_of_the_residual(hostmath, kind, entry):
    """com_mod.Kd of the mixed solid: with the VMS constants switched off (ctau_M = ctau_C = 0) lKd = af dR/dd EXACTLY, momentum and
    continuity rows alike (ustruct.cpp:1455-1572, 820-856): Kd . delta = af (R(d + e delta) - R(d - e delta)) / (2 e).  With the
    stabilisation on, the reference does not linearise tauM(J), tauC(J) (compute_tau is evaluated at the current J), so only the
    momentum rows, where those terms are small, still agree to 1e-5 — measured, and asserted as such."""
    from tests.test_hostmath_cpu import HostUstructArgs
    m = meshgen.box_hex8(3, 2, 2, (1.0, 1.0, 1.0)) if kind == "hex8" else meshgen.box_tet4(2, 2, 2, (1.0, 1.0, 1.0))
    rowPtr, colPtr = _csr(m)
    Ag, Yg, Dg, Bf, _ = common.ustruct_state(m)
    eq = abi.ustruct_eq(1e-3)
    rng = np.random.default_rng(1)
    delta = np.zeros((4, m.nNo), order="F"); delta[:3] = rng.standard_normal((3, m.nNo))

    def run(d, D):
        A = HostUstructArgs()
        keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
                np.ascontiguousarray(Yg.T), np.ascontiguousarray(D.T), np.ascontiguousarray(Bf.T)]
        A.IEN, A.x, A.Ag, A.Yg, A.Dg, A.Bf = (k.ctypes.data for k in keep)
        A.eNoN, A.nEl, A.tDof, A.s, A.nFn = m.eNoN, m.nEl, 4, 0, 0
        A.nG = _fill_tables(A, m.eNoN)
        A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
        dm = A.dm.st
        dm.rho, dm.Kpen, dm.C10, dm.C01 = d.rho, d.Kpen, d.C10, d.C01
        for i in range(3):
            dm.f[i] = d.f[i]
        dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
        A.dm.E, A.dm.nu, A.dm.ctM, A.dm.ctC = d.E, d.nu, d.ctau_M, d.ctau_C
        R = np.zeros((m.nNo, 4)); V = np.zeros((len(colPtr), 16)); Kd = np.zeros((len(colPtr), 12))
        rc = getattr(hostmath, entry)(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                      R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), Kd.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return R.T, Kd.T

    af = eq.af * eq.gam * eq.dt
    rows = np.repeat(np.arange(m.nNo), np.diff(rowPtr))
    for ct, tol_m, tol_c in ((0.0, 1e-8, 1e-8), (1e-3, 1e-5, None)):
        d = abi.ustruct_domain(E=1e6, nu=0.45, Kpen=2e6, rho=1.2, ctau_M=ct, ctau_C=ct, f=(0.1, -0.2, 0.3))
        _, Kd = run(d, Dg)
        e = 1e-6
        Rp, _ = run(d, np.asfortranarray(Dg + e * delta))
        Rm, _ = run(d, np.asfortranarray(Dg - e * delta))
        y = np.zeros((m.nNo, 4))
        np.add.at(y, rows, np.einsum("kij,kj->ki", Kd.T.reshape(-1, 4, 3), delta[:3].T[colPtr]))
        rhs = af * (Rp - Rm) / (2 * e)
        assert np.abs(y.T[:3] - rhs[:3]).max() < tol_m * np.abs(rhs[:3]).max()
        if tol_c is not None:
            assert np.abs(y.T[3] - rhs[3]).max() < tol_c * np.abs(rhs[3]).max()


def test_fluid_tangent_against_the_derivative_of_the_residual(hostmath):
    """lK of fluid_3d_m / fluid_3d_c against dR/d(delta) for the generalised-alpha increment (Ag += am delta, Yg += af gam dt delta).
    At rest (u = 0: no convection, tau_M constant) the tangent is the exact derivative (1e-6); in a flow the reference's tangent
    leaves parts of the VMS terms unlinearised (fluid.cpp:2146-2224 freezes tau_M, tau_C and the fine-scale velocity), which
    shows as a 1e-3 .. 1e-2 relative difference — measured, bounded here, and the reason Newton needs a few iterations."""
    m = meshgen.cylinder_tet4(3, 4)
    rowPtr, colPtr = _csr(m)
    rng = np.random.default_rng(2)
    eq = abi.fluid_eq(0.005)
    Bf = np.zeros((3, m.nNo), order="F")
    d = abi.fluid_domain()
    c_a, c_y = eq.am, eq.af * eq.gam * eq.dt
    for scale, tol in ((0.0, 1e-6), (1.0, 2e-2)):
        A0 = np.asfortranarray(0.1 * rng.standard_normal((4, m.nNo)))
        Y0 = np.asfortranarray(scale * rng.standard_normal((4, m.nNo))); Y0[3] = rng.standard_normal(m.nNo)
        delta = np.asfortranarray(rng.standard_normal((4, m.nNo)))
        _, V = _fluid(hostmath, m, A0, Y0, Bf, d, rowPtr, colPtr)
        e = 1e-6
        Rp, _ = _fluid(hostmath, m, np.asfortranarray(A0 + e * c_a * delta), np.asfortranarray(Y0 + e * c_y * delta), Bf, d, rowPtr, colPtr)
        Rm, _ = _fluid(hostmath, m, np.asfortranarray(A0 - e * c_a * delta), np.asfortranarray(Y0 - e * c_y * delta), Bf, d, rowPtr, colPtr)
        lhs, rhs = _csr_matvec(rowPtr, colPtr, V, delta, dof=4), (Rp - Rm) / (2 * e)
        assert np.abs(lhs[:3] - rhs[:3]).max() < tol * np.abs(rhs[:3]).max()
        assert np.abs(lhs[3] - rhs[3]).max() < tol * np.abs(rhs[3]).max()


def _fluid_gen(hostmath, m, Ag, Yg, Bf, d, rowPtr, colPtr, dt=0.005):
    from tests.test_hostmath_cpu import HostFluidGenArgs
    eq = abi.fluid_eq(dt)
    w, N, Nx = elements.tables(m.eNoN)
    Nxx = elements.nxx_tables(m.eNoN)
    A = HostFluidGenArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Bf = (k.ctypes.data for k in keep)
    A.eNoN, A.nEl, A.nG, A.tDof, A.mvMsh, A.factored = m.eNoN, m.nEl, len(w), 4, 0, 1
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    for g in range(len(w)):
        A.w[g] = w[g]
        for a in range(m.eNoN):
            A.N[g][a] = N[a, g]
            for k in range(3):
                A.Nxi[g][a][k] = Nx[k, a, g]
            for k in range(6):
                A.Nxi2[g][a][k] = Nxx[k, a, g]
    A.dm.rho, A.dm.Kd = d.rho, d.K_darcy
    for i in range(3):
        A.dm.f[i] = d.f[i]
    A.dm.mu_i, A.dm.viscType, A.dm.Id, A.dm.isFluid = d.mu_i, d.viscType, -1, 1
    R = np.zeros((m.nNo, 4))
    V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_gen(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                     R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return R.T, V.T, eq


def test_general_fluid_kernel_algebra_on_skewed_hex8(hostmath):
    """The per-Gauss-point VMS algebra (gnn + gn_nxx, second-derivative terms) on non-parallelepiped HEX8: uniform flow and the
    hydrostatic state leave no interior residual; at rest the tangent is the derivative of the residual."""
    m = common._hex_skewed()
    rowPtr, colPtr = _csr(m)
    z = np.zeros((4, m.nNo), order="F")
    Bf = np.zeros((3, m.nNo), order="F")
    interior = np.ones(m.nNo, bool)
    for k in ("X0", "X1", "Y0", "Y1", "Z0", "Z1"):
        interior[m.faces[k]] = False
    assert interior.any()
    Y = z.copy(order="F"); Y[0], Y[1], Y[2] = 1.5, -0.7, 3.0
    R, _, _ = _fluid_gen(hostmath, m, z, Y, Bf, abi.fluid_domain(), rowPtr, colPtr)
    assert np.abs(R[:, interior]).max() < 1e-11
    d = abi.fluid_domain(rho=1.06, f=(0.3, -0.2, 0.5))
    Y = z.copy(order="F"); Y[3] = d.rho * (0.3 * m.x[0] - 0.2 * m.x[1] + 0.5 * m.x[2])
    R, _, _ = _fluid_gen(hostmath, m, z, Y, Bf, d, rowPtr, colPtr)
    assert np.abs(R[:, interior]).max() < 1e-10
    # tangent at rest
    rng = np.random.default_rng(6)
    A0 = np.asfortranarray(0.1 * rng.standard_normal((4, m.nNo)))
    Y0 = z.copy(order="F"); Y0[3] = rng.standard_normal(m.nNo)
    delta = np.asfortranarray(rng.standard_normal((4, m.nNo)))
    _, V, eq = _fluid_gen(hostmath, m, A0, Y0, Bf, abi.fluid_domain(), rowPtr, colPtr)
    e, c_a, c_y = 1e-6, eq.am, eq.af * eq.gam * eq.dt
    Rp, _, _ = _fluid_gen(hostmath, m, np.asfortranarray(A0 + e * c_a * delta), np.asfortranarray(Y0 + e * c_y * delta), Bf, abi.fluid_domain(), rowPtr, colPtr)
    Rm, _, _ = _fluid_gen(hostmath, m, np.asfortranarray(A0 - e * c_a * delta), np.asfortranarray(Y0 - e * c_y * delta), Bf, abi.fluid_domain(), rowPtr, colPtr)
    lhs, rhs = _csr_matvec(rowPtr, colPtr, V, delta, dof=4), (Rp - Rm) / (2 * e)
    print("hex8 tangent-vs-FD at rest:", np.abs(lhs[:3] - rhs[:3]).max() / np.abs(rhs[:3]).max(), np.abs(lhs[3] - rhs[3]).max() / np.abs(rhs[3]).max())
    assert np.abs(lhs[:3] - rhs[:3]).max() < 1e-4 * np.abs(rhs[:3]).max()
    assert np.abs(lhs[3] - rhs[3]).max() < 1e-4 * np.abs(rhs[3]).max()


# ---- Taylor-Hood fluid (fluid_thood.cuh): oracle-independent identities -----------------------------------------------------------------
def _csr_any(m):
    """CSR graph of any mesh: all node pairs of an element, columns sorted."""
    E = m.IEN.shape[0]
    r = np.repeat(m.IEN, E, axis=0).ravel(order="F").astype(np.int64)
    c = np.tile(m.IEN, (E, 1)).ravel(order="F").astype(np.int64)
    key = np.unique(r * m.nNo + c)
    rows, cols = (key // m.nNo).astype(np.int32), (key % m.nNo).astype(np.int32)
    rowPtr = np.zeros(m.nNo + 1, np.int32)
    np.add.at(rowPtr, rows + 1, 1)
    return np.cumsum(rowPtr).astype(np.int32), np.ascontiguousarray(cols)


def _fluid_thood(hostmath, m, Ag, Yg, Bf, d, rowPtr, colPtr, dt=0.005):
    from tests.test_hostmath_cpu import HostThoodArgs
    golden, tabs = common.load_golden("fluid_thood.npz"), common.load_golden("fluid_hi.npz")
    w, N, Nx, Nxx = (tabs[f"tables/tet10/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    t = {k: golden[f"tables/tet10/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
    eq = common.fluid_thood_eq(dt)
    tr = lambda a: np.ascontiguousarray(np.asarray(a).T)
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), tr(m.x), tr(Ag), tr(Yg), tr(Bf), np.ascontiguousarray(w), tr(N), tr(Nx), tr(Nxx),
            tr(t["Nq1"]), tr(t["Nqxi1"]), np.ascontiguousarray(t["w2"]), tr(t["Nw2"]), tr(t["Nwxi2"]), tr(t["Nq2"]), tr(t["Nqxi2"])]
    A = HostThoodArgs()
    (A.IEN, A.x, A.Ag, A.Yg, A.Bf, A.w, A.N, A.Nxi, A.Nxi2, A.Nq1, A.Nqxi1, A.w2, A.Nw2, A.Nwxi2, A.Nq2, A.Nqxi2) = (k.ctypes.data for k in keep)
    A.eNoN, A.eNoNq, A.nEl, A.nG, A.nG2, A.tDof, A.mvMsh, A.lShpFq = 10, 4, m.nEl, len(w), int(t["nG2"]), 4, 0, int(t["lShpF_q"])
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    A.dm.rho, A.dm.Kd = d.rho, d.K_darcy
    for i in range(3):
        A.dm.f[i] = d.f[i]
    A.dm.mu_i, A.dm.mu_o, A.dm.lam, A.dm.a, A.dm.n = d.mu_i, d.mu_o, d.lam, d.a, d.n
    A.dm.viscType, A.dm.Id, A.dm.isFluid = d.viscType, -1, 1
    R = np.zeros((m.nNo, 4)); V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_thood(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                       R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return R.T.copy(), V.T.copy()


def test_taylor_hood_tangent_and_continuity_invariants(hostmath):
    """P2-P1 algebra, no oracle: (i) the continuity rows are linear in the velocity and independent of the pressure, so their tangent is
    the exact derivative in any state; (ii) at rest the momentum tangent is the derivative of the momentum residual up to the one term
    the reference keeps from the VMS form — rho N_a (up . grad N_b) in T1 (fluid.cpp:2147), although with vmsFlag false the residual
    convects with u, not u + up (fluid.cpp:2070-2077, 2097-2099): measured 2 % of the largest entry, bounded at 5 % here; (iii) the continuity residual sums to the integral of div u over the mesh (the pressure shape functions sum to
    one); (iv) rows 3 of the edge nodes and the whole pressure-pressure block are empty before fs::thood_val_rc."""
    m = meshgen.elevate(meshgen.box_tet4(2, 2, 2, (1.0, 1.0, 1.0)), "tet10", bend=0.0)       # straight edges: exact quadrature of div u
    rowPtr, colPtr = _csr_any(m)
    rng = np.random.default_rng(4)
    eq = common.fluid_thood_eq(0.005)
    Bf = np.zeros((3, m.nNo), order="F")
    d = abi.fluid_domain()
    c_a, c_y = eq.am, eq.af * eq.gam * eq.dt
    edge = np.setdiff1d(np.arange(m.nNo), np.unique(m.IEN[:4]))
    for scale, tol_m in ((0.0, 5e-2), (1.0, None)):
        A0 = np.asfortranarray(0.1 * rng.standard_normal((4, m.nNo)))
        Y0 = np.asfortranarray(scale * rng.standard_normal((4, m.nNo))); Y0[3] = rng.standard_normal(m.nNo)
        delta = np.asfortranarray(rng.standard_normal((4, m.nNo)))
        R0, V = _fluid_thood(hostmath, m, A0, Y0, Bf, d, rowPtr, colPtr)
        e = 1e-6
        Rp, _ = _fluid_thood(hostmath, m, np.asfortranarray(A0 + e * c_a * delta), np.asfortranarray(Y0 + e * c_y * delta), Bf, d, rowPtr, colPtr)
        Rm, _ = _fluid_thood(hostmath, m, np.asfortranarray(A0 - e * c_a * delta), np.asfortranarray(Y0 - e * c_y * delta), Bf, d, rowPtr, colPtr)
        lhs, rhs = _csr_matvec(rowPtr, colPtr, V, delta, dof=4), (Rp - Rm) / (2 * e)
        assert np.abs(lhs[3] - rhs[3]).max() < 1e-7 * np.abs(rhs[3]).max()                   # (i)
        if tol_m is not None:
            assert np.abs(lhs[:3] - rhs[:3]).max() < tol_m * np.abs(rhs[:3]).max()           # (ii)
        assert not R0[3, edge].any() and not V[15].any() and not V[12:15][:, np.isin(np.repeat(np.arange(m.nNo), np.diff(rowPtr)), edge)].any()   # (iv)
    # (iii) u = (x, 2y, -0.5z): div u = 2.5, volume 1
    Y = np.zeros((4, m.nNo), order="F"); Y[0], Y[1], Y[2] = m.x[0], 2.0 * m.x[1], -0.5 * m.x[2]
    R, _ = _fluid_thood(hostmath, m, np.zeros((4, m.nNo), order="F"), Y, Bf, d, rowPtr, colPtr)
    assert abs(R[3].sum() - 2.5) < 1e-12
