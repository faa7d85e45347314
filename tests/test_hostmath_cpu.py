"""CPU test of the DEVICE element algebra: svmultiphysics_b200/csrc/fluid_elem.cuh compiled for the host
(tests/hostmath, test-only) must reproduce the reference's assembled R / Val to 1e-12."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements
from tests import common

HERE = os.path.dirname(os.path.abspath(__file__))


class FluidDmn(C.Structure):
    _fields_ = [("rho", C.c_double), ("f", C.c_double * 3), ("Kd", C.c_double), ("mu_i", C.c_double), ("mu_o", C.c_double),
                ("lam", C.c_double), ("a", C.c_double), ("n", C.c_double),
                ("viscType", C.c_int), ("Id", C.c_int), ("isFluid", C.c_int), ("pad", C.c_int)]


class FluidArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "eId", "slot", "perm", "kU_ptr", "kU_ent", "kU_partner", "kContrib", "rU_ptr", "rU_ent", "rContrib", "x", "Ag", "Yg", "Bf", "Dg", "R", "Val")] + \
               [(k, C.c_int) for k in ("e0", "e1", "tDof", "mvMsh", "nDmn", "atomic", "ale", "pad0")] + [("err", C.c_void_p), ("gperm", C.c_void_p), ("g0", C.c_int), ("nGrpLaunch", C.c_int)] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam")] + \
               [("w", C.c_double * 8), ("N", (C.c_double * 8) * 8), ("Nxi", ((C.c_double * 3) * 8) * 8), ("dmn", FluidDmn * 8),
                ("uris", C.c_void_p), ("nUris", C.c_int), ("urisP", abi.Uris * abi.MAX_URIS), ("emask", C.c_void_p), ("emask_val", C.c_int)]


@pytest.fixture(scope="module")
def hostmath():
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    subprocess.check_call(["make"], cwd=os.path.join(HERE, "hostmath"))
    lib = C.CDLL(so)
    assert lib.hostmath_sizeof_fluidargs() == C.sizeof(FluidArgs)
    return lib


@pytest.mark.parametrize("variant", ["hostmath_fluid_tet4", "hostmath_fluid_tet4_staged", "hostmath_fluid_tet4_acc"],
                         ids=["hoisted", "staged", "accumulator_form"])
@pytest.mark.parametrize("case", common.FLUID_CASES, ids=[c[0] for c in common.FLUID_CASES])
def test_device_element_algebra_matches_golden(hostmath, case, variant):
    golden = common.load_golden()
    name, visc, Kd, f, tDof, mv = case
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=common.GOLDEN_N, nz=common.GOLDEN_NZ, tDof=tDof)
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv)
    d = abi.fluid_domain(K_darcy=Kd, f=f, **visc)
    w, N, Nx = elements.tables(4)
    A = FluidArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Bf = (k.ctypes.data for k in keep)
    A.e0, A.e1, A.tDof, A.mvMsh, A.nDmn = 0, m.nEl, tDof, mv, 1
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
    A.dmn[0].mu_i, A.dmn[0].mu_o, A.dmn[0].lam, A.dmn[0].a, A.dmn[0].n = d.mu_i, d.mu_o, d.lam, d.a, d.n
    A.dmn[0].viscType, A.dmn[0].Id, A.dmn[0].isFluid = d.viscType, -1, 1
    rowPtr, colPtr = golden["rowPtr"], golden["colPtr"]
    R = np.zeros((m.nNo, 4))
    V = np.zeros((len(colPtr), 16))
    rc = getattr(hostmath, variant)(C.byref(A), m.nNo, rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                      R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(R.T, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V.T, golden[f"{name}/Val"]) < 1e-12


class StructDmn(C.Structure):
    _fields_ = [("rho", C.c_double), ("f", C.c_double * 3), ("dmp", C.c_double),
                ("Kpen", C.c_double), ("C10", C.c_double), ("C01", C.c_double), ("bff", C.c_double), ("bss", C.c_double),
                ("bfs", C.c_double), ("st_a", C.c_double), ("st_b", C.c_double), ("aff", C.c_double), ("ass", C.c_double),
                ("afs", C.c_double), ("kap", C.c_double), ("khs", C.c_double), ("visc_mu", C.c_double),
                ("isoType", C.c_int), ("volType", C.c_int), ("Id", C.c_int), ("isStruct", C.c_int), ("viscType", C.c_int), ("active", C.c_int),
                ("cann_off", C.c_int), ("cann_rows", C.c_int)]


class CannRow(C.Structure):
    _fields_ = [("inv", C.c_int), ("a0", C.c_int), ("a1", C.c_int), ("a2", C.c_int), ("w0", C.c_double), ("w1", C.c_double), ("w2", C.c_double)]


_EXTRAS = [("Ya", C.c_void_p), ("cann", CannRow * 16)]


def fill_cann(dm, table, d):
    """CANN parameter table of the domain d -> the host argument block (rows 0.. of `table`)."""
    dm.cann_off, dm.cann_rows = 0, d.cann_rows
    for r in range(d.cann_rows):
        table[r].inv = d.cann_inv[r]
        table[r].a0, table[r].a1, table[r].a2 = d.cann_act[r][0], d.cann_act[r][1], d.cann_act[r][2]
        table[r].w0, table[r].w1, table[r].w2 = d.cann_w[r][0], d.cann_w[r][1], d.cann_w[r][2]


def _fill_extras(A, dm, d, m, keep):
    """Active tensions and the CANN table of a solid case (tests/common.py: active_tension)."""
    fill_cann(dm, A.cann, d)
    dm.active = d.active_stress
    if d.active_stress:
        ya = np.ascontiguousarray(np.stack(common.active_tension(m, d.isoType), axis=1))
        keep.append(ya)
        A.Ya = ya.ctypes.data


class HostStructArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "fN", "x", "Ag", "Yg", "Dg", "Bf")] + \
               [(k, C.c_int) for k in ("eNoN", "nEl", "nG", "tDof", "dof", "s", "nFn")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam", "beta")] + \
               [("w", C.c_double * 8), ("N", (C.c_double * 8) * 8), ("Nxi", ((C.c_double * 3) * 8) * 8), ("dm", StructDmn)] + _EXTRAS


@pytest.mark.parametrize("name,mk,dkw,nFn", common.STRUCT_CASES, ids=[c[0] for c in common.STRUCT_CASES])
def test_device_solid_algebra_matches_golden(hostmath, name, mk, dkw, nFn):
    """svmultiphysics_b200/csrc/struct_elem.cuh (pk2cc_voigt, solid viscosity, Bm/DBm blocks) compiled for the host
    against the R / Val the unmodified reference assembled (tests/golden/struct.npz), tolerance 1e-12."""
    golden = common.load_golden("struct.npz")
    assert hostmath.hostmath_sizeof_structargs() == C.sizeof(HostStructArgs)
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
    eq, d = abi.struct_eq(1e-4), abi.struct_domain(**dkw)
    w, N, Nx = elements.tables(m.eNoN)
    A = HostStructArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Dg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Dg, A.Bf = (k.ctypes.data for k in keep)
    if nFn:
        fk = np.ascontiguousarray(fN.T)
        A.fN = fk.ctypes.data
    A.eNoN, A.nEl, A.nG, A.tDof, A.dof, A.s, A.nFn = m.eNoN, m.nEl, len(w), 3, 3, 0, nFn
    A.dt, A.af, A.am, A.gam, A.beta = eq.dt, eq.af, eq.am, eq.gam, eq.beta
    for g in range(len(w)):
        A.w[g] = w[g]
        for a in range(m.eNoN):
            A.N[g][a] = N[a, g]
            for k in range(3):
                A.Nxi[g][a][k] = Nx[k, a, g]
    dm = A.dm
    dm.rho, dm.dmp, dm.Kpen, dm.C10, dm.C01, dm.bff, dm.bss, dm.bfs = d.rho, d.dmp, d.Kpen, d.C10, d.C01, d.bff, d.bss, d.bfs
    for i in range(3):
        dm.f[i] = d.f[i]
    dm.visc_mu, dm.viscType = d.solid_visc_mu, d.solidViscType
    dm.st_a, dm.st_b, dm.aff, dm.ass, dm.afs, dm.kap, dm.khs = d.st_a, d.st_b, d.aff, d.ass, d.afs, d.kap, d.khs
    dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
    _fill_extras(A, dm, d, m, keep)
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 3))
    V = np.zeros((len(colPtr), 9))
    rc = hostmath.hostmath_struct(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                  R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(R.T, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V.T, golden[f"{name}/Val"]) < 1e-12


class HostFluidGenArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "x", "Ag", "Yg", "Bf")] + \
               [(k, C.c_int) for k in ("eNoN", "nEl", "nG", "tDof", "mvMsh", "factored")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam")] + \
               [("w", C.c_double * 8), ("N", (C.c_double * 8) * 8), ("Nxi", ((C.c_double * 3) * 8) * 8),
                ("Nxi2", ((C.c_double * 6) * 8) * 8), ("dm", FluidDmn),
                ("uris", C.c_void_p), ("nUris", C.c_int), ("urisP", abi.Uris * abi.MAX_URIS)]


@pytest.mark.parametrize("factored", [0, 1], ids=["reference_form", "factored_rows"])
@pytest.mark.parametrize("case", common.FLUID_GEN_CASES, ids=[c[0] for c in common.FLUID_GEN_CASES])
def test_device_general_fluid_algebra_matches_golden(hostmath, case, factored):
    """svmultiphysics_b200/csrc/fluid_gen.cuh (gnn + gn_nxx per Gauss point, fluid_3d_m/c for any element) compiled for
    the host against what the unmodified reference assembled for skewed HEX8 meshes (tests/golden/fluid_gen.npz)."""
    golden = common.load_golden("fluid_gen.npz")
    assert hostmath.hostmath_sizeof_fluidgenargs() == C.sizeof(HostFluidGenArgs)
    name, mk, visc, Kd, f, tDof, mv = case
    m = mk()
    Ag, Yg, _, Bf = common.fluid_gen_state(m, tDof)
    eq, d = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv), abi.fluid_domain(K_darcy=Kd, f=f, **visc)
    w, N, Nx = elements.tables(m.eNoN)
    Nxx = elements.nxx_tables(m.eNoN)
    assert np.array_equal(Nxx, golden[f"{name}/Nxx"])
    A = HostFluidGenArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Bf = (k.ctypes.data for k in keep)
    A.eNoN, A.nEl, A.nG, A.tDof, A.mvMsh, A.factored = m.eNoN, m.nEl, len(w), tDof, mv, factored
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
    A.dm.mu_i, A.dm.mu_o, A.dm.lam, A.dm.a, A.dm.n = d.mu_i, d.mu_o, d.lam, d.a, d.n
    A.dm.viscType, A.dm.Id, A.dm.isFluid = d.viscType, -1, 1
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 4))
    V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_gen(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                     R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(R.T, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V.T, golden[f"{name}/Val"]) < 1e-12


@pytest.mark.parametrize("factored", [0, 1], ids=["reference_form", "factored_rows"])
@pytest.mark.parametrize("name", [c[0] for c in common.URIS_CASES if not c[4]])
def test_device_fluid_algebra_with_uris_valves_matches_golden(hostmath, name, factored):
    """URIS penalty terms of fluid_3d_m / fluid_3d_c (fluid.cpp:2006-2008, 2042-2047, 2126-2129, 2166-2204, 2228-2234, 1660-1703) and
    the per-Gauss-point valve factor (uris.cpp:1577-1673) in fluid_gen.cuh, compiled for the host, against tests/golden/fluid_uris.npz
    (compiled reference; two valves: ramped thickness + valve velocity, scaffold)."""
    golden = common.load_golden("fluid_uris.npz")
    assert hostmath.hostmath_sizeof_fluidgenargs() == C.sizeof(HostFluidGenArgs)
    m, Ag, Yg, Dg, Bf, eq, dmn = common.uris_case(name)
    raw, dev, sdf, udf, vel = common.uris_valves(m)
    d = dmn[0]
    w, N, Nx = elements.tables(m.eNoN)
    Nxx = elements.nxx_tables(m.eNoN)
    A = HostFluidGenArgs()
    nodal = np.zeros((m.nNo, len(dev), 5))
    nodal[:, :, 0] = np.abs(sdf).T
    nodal[:, :, 1] = np.abs(udf).T * np.array([v.scaffold for v in dev])[None, :]
    nodal[:, :, 2:5] = vel.transpose(1, 0, 2) * np.array([v.include_velocity for v in dev])[None, :, None]
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Bf.T), np.ascontiguousarray(nodal)]
    A.IEN, A.x, A.Ag, A.Yg, A.Bf, A.uris = (k.ctypes.data for k in keep)
    A.nUris = len(dev)
    for v, u in enumerate(dev):
        A.urisP[v] = u
    A.eNoN, A.nEl, A.nG, A.tDof, A.mvMsh, A.factored = m.eNoN, m.nEl, len(w), eq.tDof, eq.mvMsh, factored
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
    A.dm.mu_i, A.dm.mu_o, A.dm.lam, A.dm.a, A.dm.n = d.mu_i, d.mu_o, d.lam, d.a, d.n
    A.dm.viscType, A.dm.Id, A.dm.isFluid = d.viscType, -1, 1
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 4)); V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_gen(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                     R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(R.T, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V.T, golden[f"{name}/Val"]) < 1e-12


class HostThoodArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "x", "Ag", "Yg", "Bf", "w", "N", "Nxi", "Nxi2", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")] + \
               [(k, C.c_int) for k in ("eNoN", "eNoNq", "nEl", "nG", "nG2", "tDof", "mvMsh", "lShpFq")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam")] + [("dm", FluidDmn)] + \
               [("uris", C.c_void_p), ("nUris", C.c_int), ("urisP", abi.Uris * abi.MAX_URIS)]


@pytest.mark.parametrize("case", common.FLUID_THOOD_CASES + [common.FLUID_THOOD_URIS_CASE], ids=[c[0] for c in common.FLUID_THOOD_CASES] + ["tet10_uris"])
def test_device_taylor_hood_fluid_algebra_matches_golden(hostmath, case):
    """svmultiphysics_b200/csrc/fluid_thood.cuh (fluid_3d_m / fluid_3d_c with vmsFlag false on P2-P1 / Q2-Q1 function spaces, the
    momentum loop on the velocity rule and the continuity loop on the pressure rule with the reference's choice of Jacobian) compiled for
    the host against what the unmodified reference assembled (tests/golden/fluid_thood.npz), entry type by entry type."""
    name, mk, visc, Kd, f, tDof, mv = case
    golden, tabs = common.load_golden("fluid_thood.npz"), common.load_golden("fluid_hi.npz")
    assert hostmath.hostmath_sizeof_thoodargs() == C.sizeof(HostThoodArgs)
    m = mk()
    et = name.split("_")[0]
    Ag, Yg, _, Bf = common.fluid_gen_state(m, tDof)
    eq, d = common.fluid_thood_eq(0.005, tDof=tDof, mvMsh=mv), abi.fluid_domain(K_darcy=Kd, f=f, **visc)
    w, N, Nx, Nxx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    t = {k: golden[f"tables/{et}/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
    tr = lambda a: np.ascontiguousarray(np.asarray(a).T)          # (k, a, g) column-major -> [g][a][k] row-major
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), tr(m.x), tr(Ag), tr(Yg), tr(Bf), np.ascontiguousarray(w), tr(N), tr(Nx), tr(Nxx),
            tr(t["Nq1"]), tr(t["Nqxi1"]), np.ascontiguousarray(t["w2"]), tr(t["Nw2"]), tr(t["Nwxi2"]), tr(t["Nq2"]), tr(t["Nqxi2"])]
    A = HostThoodArgs()
    (A.IEN, A.x, A.Ag, A.Yg, A.Bf, A.w, A.N, A.Nxi, A.Nxi2, A.Nq1, A.Nqxi1, A.w2, A.Nw2, A.Nwxi2, A.Nq2, A.Nqxi2) = (k.ctypes.data for k in keep)
    A.eNoN, A.eNoNq, A.nEl, A.nG, A.nG2, A.tDof, A.mvMsh, A.lShpFq = m.eNoN, int(t["eNoNq"]), m.nEl, len(w), int(t["nG2"]), tDof, mv, int(t["lShpF_q"])
    assert int(t["nG1"]) == len(w)
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    A.dm.rho, A.dm.Kd = d.rho, d.K_darcy
    for i in range(3):
        A.dm.f[i] = d.f[i]
    A.dm.mu_i, A.dm.mu_o, A.dm.lam, A.dm.a, A.dm.n = d.mu_i, d.mu_o, d.lam, d.a, d.n
    A.dm.viscType, A.dm.Id, A.dm.isFluid = d.viscType, -1, 1
    if name.endswith("uris"):
        raw, dev, sdf, udf, vel = common.uris_valves(m)
        nodal = np.zeros((m.nNo, len(dev), 5))
        nodal[:, :, 0] = np.abs(sdf).T
        nodal[:, :, 1] = np.abs(udf).T * np.array([v.scaffold for v in dev])[None, :]
        nodal[:, :, 2:5] = vel.transpose(1, 0, 2) * np.array([v.include_velocity for v in dev])[None, :, None]
        keep.append(np.ascontiguousarray(nodal))
        A.uris, A.nUris = keep[-1].ctypes.data, len(dev)
        for v, u in enumerate(dev):
            A.urisP[v] = u
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 4)); V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_thood(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                       R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    GR, GV = golden[f"{name}/R"], golden[f"{name}/Val"]
    assert common.rel_err(R.T[:3], GR[:3]) < 1e-12 and common.rel_err(R.T[3], GR[3]) < 1e-12
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14]):
        assert common.rel_err(V.T[rows], GV[rows]) < 1e-12
    assert not V.T[15].any() and not GV[15].any()


class HostFluidAnyArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "x", "Ag", "Yg", "Bf", "w", "N", "Nxi", "Nxi2")] + \
               [(k, C.c_int) for k in ("eNoN", "nEl", "nG", "tDof", "mvMsh", "lShpF")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam")] + [("dm", FluidDmn)]


@pytest.mark.parametrize("case", common.FLUID_HI_CASES, ids=[c[0] for c in common.FLUID_HI_CASES])
def test_device_fluid_algebra_on_quadratic_and_wedge_elements(hostmath, case):
    """fluid_gen.cuh on curved TET10 (15 Gauss points), HEX20 / HEX27 (27) and WDG (6) elements against what the unmodified reference
    assembled (tests/golden/fluid_hi.npz), with the reference's own shape-function tables — the general path of
    construct_fluid: nn::gnn + nn::gn_nxx per Gauss point, fluid_3d_m / fluid_3d_c (fluid.cpp:620-745)."""
    golden = common.load_golden("fluid_hi.npz")
    assert hostmath.hostmath_sizeof_fluidanyargs() == C.sizeof(HostFluidAnyArgs)
    name, mk, visc, Kd, f, tDof, mv = case
    m = mk()
    Ag, Yg, _, Bf = common.fluid_gen_state(m, tDof)
    eq, d = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv), abi.fluid_domain(K_darcy=Kd, f=f, **visc)
    et = name.split("_")[0]
    w, N, Nx, Nxx = (golden[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    nG = len(w)
    assert N.shape == (m.eNoN, nG) and Nx.shape == (3, m.eNoN, nG) and Nxx.shape == (6, m.eNoN, nG)
    A = HostFluidAnyArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Bf.T), np.ascontiguousarray(w), np.ascontiguousarray(N.T),
            np.ascontiguousarray(Nx.transpose(2, 1, 0)), np.ascontiguousarray(Nxx.transpose(2, 1, 0))]
    A.IEN, A.x, A.Ag, A.Yg, A.Bf, A.w, A.N, A.Nxi, A.Nxi2 = (k.ctypes.data for k in keep)
    A.eNoN, A.nEl, A.nG, A.tDof, A.mvMsh = m.eNoN, m.nEl, nG, tDof, mv
    # mshType::lShpF is set for the wedge (nn_elem_props.h:23-30): construct_fluid evaluates gnn / gn_nxx at Gauss point 0 only,
    # although the wedge's gradients vary — the device path reproduces that
    A.lShpF = 1 if m.eNoN in (4, 6) else 0
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    A.dm.rho, A.dm.Kd = d.rho, d.K_darcy
    for i in range(3):
        A.dm.f[i] = d.f[i]
    A.dm.mu_i, A.dm.mu_o, A.dm.lam, A.dm.a, A.dm.n = d.mu_i, d.mu_o, d.lam, d.a, d.n
    A.dm.viscType, A.dm.Id, A.dm.isFluid = d.viscType, -1, 1
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 4))
    V = np.zeros((len(colPtr), 16))
    rc = hostmath.hostmath_fluid_any(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                     R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(R.T, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V.T, golden[f"{name}/Val"]) < 1e-12


class HeatDmn(C.Structure):
    _fields_ = [("rho", C.c_double), ("nu", C.c_double), ("s", C.c_double),
                ("Id", C.c_int), ("active", C.c_int), ("pad0", C.c_int), ("pad1", C.c_int)]


class HostHeatArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "x", "Ag", "Yg")] + \
               [(k, C.c_int) for k in ("eNoN", "nEl", "nG", "tDof", "s", "mvMsh", "fluid", "pad")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam")] + \
               [("w", C.c_double * 8), ("N", (C.c_double * 8) * 8), ("Nxi", ((C.c_double * 3) * 8) * 8), ("dm", HeatDmn)]


def _fill_tables(A, eNoN):
    w, N, Nx = elements.tables(eNoN)
    for g in range(len(w)):
        A.w[g] = w[g]
        for a in range(eNoN):
            A.N[g][a] = N[a, g]
            for k in range(3):
                A.Nxi[g][a][k] = Nx[k, a, g]
    return len(w)


@pytest.mark.parametrize("name,mk,fluid,tDof,s,mv,dkw", common.HEAT_CASES, ids=[c[0] for c in common.HEAT_CASES])
def test_device_heat_algebra_matches_golden(hostmath, name, mk, fluid, tDof, s, mv, dkw):
    """svmultiphysics_b200/csrc/heat_elem.cuh compiled for the host against the R / Val that the unmodified reference
    (heats_3d / heatf_3d) assembled (tests/golden/heat.npz), tolerance 1e-12."""
    golden = common.load_golden("heat.npz")
    assert hostmath.hostmath_sizeof_heatargs() == C.sizeof(HostHeatArgs)
    m = mk()
    Ag, Yg, _, _ = common.heat_state(m, tDof, s)
    eq, d = abi.heat_eq(0.01, fluid, tDof=tDof, s=s, mvMsh=mv), abi.heat_domain(fluid, **dkw)
    A = HostHeatArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T)]
    A.IEN, A.x, A.Ag, A.Yg = (k.ctypes.data for k in keep)
    A.eNoN, A.nEl, A.tDof, A.s, A.mvMsh, A.fluid = m.eNoN, m.nEl, tDof, s, mv, int(fluid)
    A.nG = _fill_tables(A, m.eNoN)
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    A.dm.rho, A.dm.nu, A.dm.s, A.dm.Id, A.dm.active = d.rho, d.conductivity, d.source_term, -1, 1
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros(m.nNo)
    V = np.zeros(len(colPtr))
    rc = hostmath.hostmath_heat(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.abs(golden[f"{name}/R"]).max() > 0
    assert common.rel_err(R, golden[f"{name}/R"].ravel()) < 1e-12
    assert common.rel_err(V, golden[f"{name}/Val"].ravel()) < 1e-12


class UstructDmn(C.Structure):
    _fields_ = [("st", StructDmn), ("E", C.c_double), ("nu", C.c_double), ("ctM", C.c_double), ("ctC", C.c_double)]


class HostUstructArgs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "fN", "x", "Ag", "Yg", "Dg", "Bf")] + \
               [(k, C.c_int) for k in ("eNoN", "nEl", "nG", "tDof", "s", "nFn")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam")] + \
               [("w", C.c_double * 8), ("N", (C.c_double * 8) * 8), ("Nxi", ((C.c_double * 3) * 8) * 8), ("dm", UstructDmn)] + _EXTRAS


USTRUCT_VARIANTS = [(c, "hostmath_ustruct") for c in common.USTRUCT_CASES] + \
                   [(c, "hostmath_ustruct_tet4") for c in common.USTRUCT_CASES if c[0].startswith("tet4") and "visc" not in c[0]]


@pytest.mark.parametrize("case,entry", USTRUCT_VARIANTS, ids=[c[0] + ("_closed_form" if e.endswith("tet4") else "") for c, e in USTRUCT_VARIANTS])
def test_device_ustruct_algebra_matches_golden(hostmath, case, entry):
    """svmultiphysics_b200/csrc/ustruct_elem.cuh compiled for the host against R / Val / Kd of the unmodified reference
    (ustruct_3d_m, ustruct_3d_c, ustruct_do_assem; tests/golden/ustruct.npz), tolerance 1e-12."""
    name, mk, dkw, nFn = case
    golden = common.load_golden("ustruct.npz")
    assert hostmath.hostmath_sizeof_ustructargs() == C.sizeof(HostUstructArgs)
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.ustruct_state(m, nFn)
    eq, d = abi.ustruct_eq(1e-3), abi.ustruct_domain(**dkw)
    A = HostUstructArgs()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Dg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Dg, A.Bf = (k.ctypes.data for k in keep)
    if nFn:
        fk = np.ascontiguousarray(fN.T)
        A.fN = fk.ctypes.data
    A.eNoN, A.nEl, A.tDof, A.s, A.nFn = m.eNoN, m.nEl, 4, 0, nFn
    A.nG = _fill_tables(A, m.eNoN)
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    dm = A.dm.st
    dm.rho, dm.dmp, dm.Kpen, dm.C10, dm.C01, dm.bff, dm.bss, dm.bfs = d.rho, d.dmp, d.Kpen, d.C10, d.C01, d.bff, d.bss, d.bfs
    for i in range(3):
        dm.f[i] = d.f[i]
    dm.st_a, dm.st_b, dm.aff, dm.ass, dm.afs, dm.kap, dm.khs = d.st_a, d.st_b, d.aff, d.ass, d.afs, d.kap, d.khs
    dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
    dm.visc_mu, dm.viscType = d.solid_visc_mu, d.solidViscType
    _fill_extras(A, dm, d, m, keep)
    A.dm.E, A.dm.nu, A.dm.ctM, A.dm.ctC = d.E, d.nu, d.ctau_M, d.ctau_C
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 4))
    V = np.zeros((len(colPtr), 16))
    Kd = np.zeros((len(colPtr), 12))
    rc = getattr(hostmath, entry)(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                  R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), Kd.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(R.T, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(Kd.T, golden[f"{name}/Kd"]) < 1e-12
    # entry-type-wise: the (v,v), (v,p), (p,v) and (p,p) blocks differ by orders of magnitude
    G = golden[f"{name}/Val"]
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14], [15]):
        assert common.rel_err(V.T[rows], G[rows]) < 1e-12


@pytest.mark.parametrize("entry", ["hostmath_ustruct", "hostmath_ustruct_tet4"])
@pytest.mark.parametrize("name", list(common.FSI_USTRUCT_CASES))
def test_device_ustruct_algebra_inside_fsi_matches_golden(hostmath, name, entry):
    """construct_fsi with a ustruct wall (fsi.cpp:243-262; tests/golden/fsi_ustruct.npz from the compiled reference): the device
    ustruct algebra run over the solid elements with the tDof = 7 state of the FSI equation reproduces Kd everywhere (only the
    solid writes it) and R / Val on the rows of nodes no fluid element touches."""
    dkw, nFn = common.FSI_USTRUCT_CASES[name]
    if entry.endswith("tet4") and "visc" in name:
        pytest.skip("the closed-form TET4 path has no solid viscosity (the kernel dispatch sends such domains to the general path)")
    golden = common.load_golden("fsi_ustruct.npz")
    m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, Ad, flags = common.fsi_ustruct_case(name)
    d = dmn[1]
    so = np.where((m.eId & 2) != 0)[0]
    A = HostUstructArgs()
    keep = [np.ascontiguousarray(m.IEN[:, so].T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Dg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Dg, A.Bf = (k.ctypes.data for k in keep)
    if nFn:
        fk = np.ascontiguousarray(fN[:, so].T); keep.append(fk)
        A.fN = fk.ctypes.data
    A.eNoN, A.nEl, A.tDof, A.s, A.nFn = 4, len(so), 7, 0, nFn
    A.nG = _fill_tables(A, 4)
    A.dt, A.af, A.am, A.gam = eq.dt, eq.af, eq.am, eq.gam
    dm = A.dm.st
    dm.rho, dm.dmp, dm.Kpen, dm.C10, dm.C01, dm.bff, dm.bss, dm.bfs = d.rho, d.dmp, d.Kpen, d.C10, d.C01, d.bff, d.bss, d.bfs
    for i in range(3):
        dm.f[i] = d.f[i]
    dm.st_a, dm.st_b, dm.aff, dm.ass, dm.afs, dm.kap, dm.khs = d.st_a, d.st_b, d.aff, d.ass, d.afs, d.kap, d.khs
    dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
    dm.visc_mu, dm.viscType = d.solid_visc_mu, d.solidViscType
    _fill_extras(A, dm, d, m, keep)
    A.dm.E, A.dm.nu, A.dm.ctM, A.dm.ctC = d.E, d.nu, d.ctau_M, d.ctau_C
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    R = np.zeros((m.nNo, 4)); V = np.zeros((len(colPtr), 16)); Kd = np.zeros((len(colPtr), 12))
    rc = getattr(hostmath, entry)(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                  R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), Kd.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert common.rel_err(Kd.T, golden[f"{name}/Kd"]) < 1e-12
    fluid_nodes = np.unique(m.IEN[:, (m.eId & 2) == 0])
    pure = np.setdiff1d(np.arange(m.nNo), fluid_nodes)
    assert len(pure) > 50
    assert common.rel_err(R[pure].T, golden[f"{name}/R"][:, pure]) < 1e-12
    slots = np.concatenate([np.arange(rowPtr[a], rowPtr[a + 1]) for a in pure])
    G = golden[f"{name}/Val"]
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14], [15]):
        assert common.rel_err(V[slots].T[rows], G[rows][:, slots]) < 1e-12
    # ustruct_r (ustruct.cpp:1742-1845) restated with numpy on the golden Kd: only the nodes of the ustruct domain take part
    amg, ami = (eq.gam - eq.am) / (eq.gam - 1.0), 1.0 / eq.am
    Rd = np.where(flags[None, :] != 0, amg * Ad - Yg[0:3], 0.0)
    KU = np.zeros((4, m.nNo))
    rows_of = np.repeat(np.arange(m.nNo), np.diff(rowPtr))
    Kg = golden[f"{name}/Kd"].reshape(4, 3, -1)
    np.add.at(KU.T, rows_of, np.einsum("ijk,jk->ki", Kg, Rd[:, colPtr]))
    KU[:, flags == 0] = 0.0
    assert common.rel_err(golden[f"{name}/R"] - ami * KU, golden[f"{name}/R_after_ustruct_r"]) < 1e-12


class HostTet4Args(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("IEN", "fN", "x", "Ag", "Yg", "Dg", "Bf", "Do")] + \
               [(k, C.c_int) for k in ("nEl", "tDof", "dof", "s", "nFn", "kind")] + \
               [(k, C.c_double) for k in ("dt", "af", "am", "gam", "beta")] + \
               [("w", C.c_double * 8), ("N", (C.c_double * 8) * 8), ("Nxi", ((C.c_double * 3) * 8) * 8), ("dm", StructDmn)] + _EXTRAS


def _run_tet4(hostmath, m, Ag, Yg, Dg, Bf, Do, eq, kind, fill_dm, nFn, fN, rowPtr, colPtr, d=None):
    assert hostmath.hostmath_sizeof_tet4args() == C.sizeof(HostTet4Args)
    A = HostTet4Args()
    keep = [np.ascontiguousarray(m.IEN.T.astype(np.int32)), np.ascontiguousarray(m.x.T), np.ascontiguousarray(Ag.T),
            np.ascontiguousarray(Yg.T), np.ascontiguousarray(Dg.T), np.ascontiguousarray(Bf.T)]
    A.IEN, A.x, A.Ag, A.Yg, A.Dg, A.Bf = (k.ctypes.data for k in keep)
    if Do is not None:
        dk = np.ascontiguousarray(Do.T); keep.append(dk)
        A.Do = dk.ctypes.data
    if nFn:
        fk = np.ascontiguousarray(fN.T); keep.append(fk)
        A.fN = fk.ctypes.data
    A.nEl, A.tDof, A.dof, A.s, A.nFn, A.kind = m.nEl, eq.tDof, eq.dof, eq.s, nFn, kind
    assert _fill_tables(A, 4) == 4
    A.dt, A.af, A.am, A.gam, A.beta = eq.dt, eq.af, eq.am, eq.gam, eq.beta
    fill_dm(A.dm)
    if d is not None:
        _fill_extras(A, A.dm, d, m, keep)
    R = np.zeros((m.nNo, eq.dof))
    V = np.zeros((len(colPtr), eq.dof * eq.dof))
    rc = hostmath.hostmath_tet4(C.byref(A), rowPtr.ctypes.data_as(C.c_void_p), colPtr.ctypes.data_as(C.c_void_p),
                                R.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return R.T, V.T


@pytest.mark.parametrize("name,mk,dkw,nFn", [c for c in common.STRUCT_CASES if c[0].startswith("tet4") and "visc" not in c[0]],
                         ids=[c[0] for c in common.STRUCT_CASES if c[0].startswith("tet4") and "visc" not in c[0]])
def test_tet4_closed_form_solid_matches_golden(hostmath, name, mk, dkw, nFn):
    """The closed-form Gauss sums of assemble_struct_tet4_kernel (tet4_moments, struct_tet4_residual, struct_tet4_block) against
    the R / Val that struct_3d produced with its four-point Gauss loop in the compiled reference."""
    golden = common.load_golden("struct.npz")
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
    eq, d = abi.struct_eq(1e-4), abi.struct_domain(**dkw)

    def fill(dm):
        dm.rho, dm.dmp, dm.Kpen, dm.C10, dm.C01, dm.bff, dm.bss, dm.bfs = d.rho, d.dmp, d.Kpen, d.C10, d.C01, d.bff, d.bss, d.bfs
        for i in range(3):
            dm.f[i] = d.f[i]
        dm.st_a, dm.st_b, dm.aff, dm.ass, dm.afs, dm.kap, dm.khs = d.st_a, d.st_b, d.aff, d.ass, d.afs, d.kap, d.khs
        dm.isoType, dm.volType, dm.Id, dm.isStruct = d.isoType, d.volType, -1, 1
    R, V = _run_tet4(hostmath, m, Ag, Yg, Dg, Bf, None, eq, 0, fill, nFn, fN, golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"], d=d)
    assert common.rel_err(R, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V, golden[f"{name}/Val"]) < 1e-12


@pytest.mark.parametrize("name", common.LELAS_CASES)
def test_tet4_closed_form_lelas_matches_golden(hostmath, name):
    """lelas_tet4_stress / _residual / _block (assemble_mesh_tet4_kernel) against l_elas_3d of the compiled reference: the
    linear-elasticity equation and the mesh-motion equation with its Jacobian-free weight and old displacement."""
    golden = common.load_golden("lelas.npz")
    m, Ag, Yg, Dg, Bf, Do, eq, dmn = common.lelas_case(name)
    d = dmn[0]

    def fill(dm):
        dm.rho, dm.C10, dm.C01 = d.rho, d.E, d.nu
        for i in range(3):
            dm.f[i] = d.f[i]
        dm.Id, dm.isStruct = -1, 1
    R, V = _run_tet4(hostmath, m, Ag, Yg, Dg, Bf, Do, eq, 1 if Do is None else 2, fill, 0, None,
                     golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"])
    assert np.abs(golden[f"{name}/R"]).max() > 0
    assert common.rel_err(R, golden[f"{name}/R"]) < 1e-12
    assert common.rel_err(V, golden[f"{name}/Val"]) < 1e-12
