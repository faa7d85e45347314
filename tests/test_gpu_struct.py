"""GPU parity tests for the solid (struct_3d) and FSI assembly and the dof = 3 solvers."""
import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements
from tests import common

pytestmark = pytest.mark.gpu
ASM_TOL = 1e-12


def _oracle():
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("struct / FSI parity needs oracle/_ref/libsvref.so (the C restatement covers the fluid path)")
    return refbind.RefCase


def _engine(m, rowPtr, colPtr, nFn=0, fN=None):
    from svmultiphysics_b200.engine import Engine
    e = Engine(0)
    e.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(m.eNoN)
    e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, nFn=nFn, fN=fN)
    e.set_coords(m.x)
    return e


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name,mk,dkw,nFn", common.STRUCT_CASES, ids=[c[0] for c in common.STRUCT_CASES])
def test_struct_assembly_parity(name, mk, dkw, nFn, scatter):
    cls = _oracle()
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN, nFn=nFn, fN=fN)
    rowPtr, colPtr = orc.build_graph(0)
    eq, dmn = abi.struct_eq(1e-4, scatter=scatter), [abi.struct_domain(**dkw)]
    orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf)
    eng = _engine(m, rowPtr, colPtr, nFn, fN)
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf)
    if dmn[0].active_stress:
        ya = common.active_tension(m, dmn[0].isoType)
        orc.set_active_tension(*ya)
        eng.set_active_tension(*ya)
    orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, R0) < ASM_TOL
    assert common.rel_err(V1, V0) < ASM_TOL
    golden = common.load_golden("struct.npz")          # and the committed vectors of the same cases
    assert common.rel_err(R1, golden[f"{name}/R"]) < ASM_TOL
    assert common.rel_err(V1, golden[f"{name}/Val"]) < ASM_TOL
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(3); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1)
    eng.close()


@pytest.mark.parametrize("ls_type,kw", [(abi.LS_BICGS, dict(mItr=600, relTol=1e-10)), (abi.LS_GMRES, dict(mItr=50, sD=60, relTol=1e-10)),
                                        (abi.LS_CG, dict(mItr=2000, relTol=1e-10))], ids=["bicgs", "gmres", "cg"])
def test_struct_solve_parity(ls_type, kw):
    """block_compression-like: symmetric Dirichlet planes X0/Y0/Z0 (one direction each), dof = 3."""
    cls = _oracle()
    m = common.STRUCT_CASES[0][1]()
    Ag, Yg, Dg, Bf, _ = common.struct_state(m)
    faces = []
    for k, name in enumerate(("X0", "Y0", "Z0")):
        val = np.ones((3, len(m.faces[name])), order="F"); val[k] = 0.0       # Effective_direction: only k is constrained
        faces.append((abi.BC_DIR, m.faces[name], val))
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(len(faces))
    eq, dmn = abi.struct_eq(1e-4), [abi.struct_domain()]
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val)
    orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    ls = abi.ls_params(ls_type, **kw)
    incL, res = np.ones(3, np.int32), np.zeros(3)
    X0, o0, _ = orc.solve(3, ls_type, ls, incL, res)
    eng = _engine(m, rowPtr, colPtr)
    eng.set_num_faces(3)
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    X1, o1, _ = eng.solve(3, ls_type, ls, incL, res)
    assert o1.RI.success == o0.RI.success
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    # BiCGStab on the nearly incompressible block (kappa/mu ~ 50) is erratic: its iteration count moves by
    # ~10 % under last-bit perturbations; GMRES/CG are held to 5 %
    slack = 6 if ls_type == abi.LS_BICGS else 20
    assert abs(o1.RI.itr - o0.RI.itr) <= max(3, o0.RI.itr // slack)
    assert common.rel_err(X1, X0) < 1e-6
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
def test_fsi_assembly_parity(scatter):
    cls = _oracle()
    m, Ag, Yg, Dg, Bf = common.fsi_case()
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN, eId=m.eId)
    rowPtr, colPtr = orc.build_graph(0)
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=scatter, reserved=0)
    dfl = abi.fluid_domain(rho=1.0, mu=0.04, Id=0)
    dso = abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, [dfl, dso])
    R0, V0 = orc.get_R(), orc.get_Val()
    eng = _engine(m, rowPtr, colPtr)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, [dfl, dso])
    R1, V1 = eng.get_R(), eng.get_Val()
    # fluid and solid blocks differ by many orders of magnitude: compare each against its own scale
    assert common.rel_err(R1, R0) < ASM_TOL
    assert common.rel_err(V1, V0) < ASM_TOL
    fl_rows = np.unique(m.IEN[:, m.eId == 1])
    assert common.rel_err(R1[:, fl_rows], R0[:, fl_rows]) < 1e-11
    eng.close()


def test_fsi_two_meshes_parity():
    """The layout of tests/cases/fsi/pipe_3d: TWO meshes over one node set (lumen = fluid domain, wall = solid domain), assembled
    mesh by mesh into the same R / Val like the loop over msh[] around global_eq_assem (eq_assem.cpp:377).  Each launch then
    sees only elements of its own physics (the all-active fast path of the grouped fluid kernel, the TET4 solid kernel)."""
    cls = _oracle()
    from svmultiphysics_b200.engine import Engine
    m, Ag, Yg, Dg, Bf = common.fsi_case()
    fl, so = np.where(m.eId == 1)[0], np.where(m.eId == 2)[0]
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=abi.SCATTER_ATOMIC, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0), abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
    orc = cls(); orc.set_coords(m.x)
    orc.add_mesh(np.asfortranarray(m.IEN[:, fl]), eId=m.eId[fl]); orc.add_mesh(np.asfortranarray(m.IEN[:, so]), eId=m.eId[so])
    rowPtr, colPtr = orc.build_graph(0)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn); orc.assemble(1, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    # the one-mesh layout of the other tests gives the same system (the CSR columns are sorted, lhsa.cpp:13-54)
    one = cls(); one.set_coords(m.x); one.add_mesh(m.IEN, eId=m.eId)
    rp1, cp1 = one.build_graph(0)
    assert np.array_equal(rp1, rowPtr) and np.array_equal(cp1, colPtr)
    one.alloc(4); one.set_state(Ag, Yg, Dg, Bf); one.assemble(0, eq, dmn)
    assert common.rel_err(one.get_R(), R0) < 1e-13
    eng = Engine(0)
    eng.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(4)
    eng.set_mesh(0, np.asfortranarray(m.IEN[:, fl]), w, N, Nx, eId=m.eId[fl])
    eng.set_mesh(1, np.asfortranarray(m.IEN[:, so]), w, N, Nx, eId=m.eId[so])
    eng.set_coords(m.x)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn); eng.assemble(1, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, R0) < ASM_TOL and common.rel_err(V1, V0) < ASM_TOL
    fl_rows = np.unique(m.IEN[:, fl])
    assert common.rel_err(R1[:, fl_rows], R0[:, fl_rows]) < 1e-11
    eng.close()


def test_fsi_solve_history():
    """GMRES(50) on the coupled FSI system (config C5).  The system is ill-conditioned (solid blocks ~1e7 x the fluid
    ones): the residual histories of the device and of the reference agree to round-off over the first iterations and
    then separate as classical Gram-Schmidt loses orthogonality (the reference stagnates until its restart, 53
    iterations; the device needs ~33).  Checked: early history, the set tolerance, and the answer."""
    cls = _oracle()
    from oracle import refbind
    m, Ag, Yg, Dg, Bf = common.fsi_case()
    wall = m.faces["wall"]
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=abi.SCATTER_ATOMIC, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0),
           abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
    ls = abi.ls_params(abi.LS_GMRES, mItr=100, sD=50, relTol=1e-8)
    incL, res, val = np.ones(1, np.int32), np.zeros(1), np.zeros((3, len(wall)), order="F")
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN, eId=m.eId)
    rowPtr, colPtr = orc.build_graph(1)
    orc.set_face(0, abi.BC_DIR, wall, val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
    # the C restatement fed with the reference's system reproduces the reference and records the history
    oc = refbind.OracleCase(); oc.set_coords(m.x); oc.add_mesh(m.IEN, eId=m.eId); oc.build_graph(1)
    oc.set_face(0, abi.BC_DIR, wall, val)
    oc.alloc(4); oc.put_Val(V0, 4); oc.put_R(R0)
    _, occ, h0 = oc.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=512)
    assert occ.RI.itr == o0.RI.itr and occ.RI.fNorm == o0.RI.fNorm
    eng = _engine(m, rowPtr, colPtr)
    eng.set_num_faces(1); eng.set_face(0, abi.BC_DIR, wall, val)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    X1, o1, h1 = eng.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=512)
    drift = np.abs(h1[:10] - h0[:10]) / h0[:10]
    assert drift.max() < 1e-9
    assert o1.RI.success and abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert o1.RI.fNorm <= ls.RI.relTol * o1.RI.iNorm
    assert o1.RI.itr <= o0.RI.itr + 3
    assert common.rel_err(X1, X0) < 1e-6
    eng.close()


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_mesh_equation_parity(kind):
    """Mesh-motion equation of an FSI run (mesh::construct_mesh + l_elas_3d): dof 3, state dofs 4..6."""
    cls = _oracle()
    if kind == "tet4":
        m, Ag, Yg, Dg, Bf = common.fsi_case()
        m.eId = None
    else:
        from svmultiphysics_b200 import meshgen
        m = meshgen.box_hex8(3, 3, 2, (1.0, 2.0, 1.0))
        rng = np.random.default_rng(5)
        Ag = np.asfortranarray(rng.standard_normal((7, m.nNo))); Yg = np.asfortranarray(rng.standard_normal((7, m.nNo)))
        Dg = np.asfortranarray(1e-2 * rng.standard_normal((7, m.nNo))); Bf = None
    Do = np.asfortranarray(Dg + 3e-3 * np.random.default_rng(9).standard_normal(Dg.shape))
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(0)
    eq, dmn = abi.mesh_eq(1e-3), [abi.mesh_domain(E=1.0, nu=0.3)]
    orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf); orc.set_old_disp(Do); orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    eng = _engine(m, rowPtr, colPtr)
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf); eng.set_old_disp(Do); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), R0) < ASM_TOL
    assert common.rel_err(eng.get_Val(), V0) < ASM_TOL
    # CG as in tests/cases/fsi/pipe_3d/solver.xml (mesh equation: LS type CG, tolerance 1e-12)
    faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in (("wall",) if kind == "tet4" else ("X0", "X1"))]
    orc2 = cls(); orc2.set_coords(m.x); orc2.add_mesh(m.IEN); orc2.build_graph(len(faces))
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc2.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    orc2.alloc(3); orc2.set_state(Ag, Yg, Dg, Bf); orc2.set_old_disp(Do); orc2.assemble(0, eq, dmn)
    ls = abi.ls_params(abi.LS_CG, mItr=1000, relTol=1e-12)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    X0, o0, _ = orc2.solve(3, abi.LS_CG, ls, incL, res)
    X1, o1, _ = eng.solve(3, abi.LS_CG, ls, incL, res)
    assert o1.RI.success == o0.RI.success and abs(o1.RI.itr - o0.RI.itr) <= max(2, o0.RI.itr // 20)
    assert common.rel_err(X1, X0) < 1e-8
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_linear_elasticity_parity(kind, scatter):
    """Linear-elasticity equation (l_elas::construct_l_elas + l_elas_3d, tests/cases/linear-elasticity): assembly to 1e-12
    and a CG solve against the compiled reference."""
    cls = _oracle()
    from svmultiphysics_b200 import meshgen
    m = meshgen.box_tet4(3, 3, 2, (1.0, 1.0, 1.0)) if kind == "tet4" else meshgen.box_hex8(4, 3, 3, (1.0, 2.0, 1.0))
    Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
    eq, dmn = abi.lelas_eq(1e-3, scatter=scatter), [abi.lelas_domain(E=1.0e6, nu=0.3, rho=2.0, f=(0.1, -0.2, 0.3))]
    faces = [(abi.BC_DIR, m.faces["X0"], np.zeros((3, len(m.faces["X0"])), order="F"))]
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(len(faces))
    eng = _engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), R0) < ASM_TOL
    assert common.rel_err(eng.get_Val(), V0) < ASM_TOL
    ls = abi.ls_params(abi.LS_CG, mItr=1000, relTol=1e-10)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    X0, o0, _ = orc.solve(3, abi.LS_CG, ls, incL, res)
    X1, o1, _ = eng.solve(3, abi.LS_CG, ls, incL, res)
    assert o1.RI.success == o0.RI.success and abs(o1.RI.itr - o0.RI.itr) <= max(2, o0.RI.itr // 20)
    assert common.rel_err(X1, X0) < 1e-7
    eng.close()


def test_active_stress_errors_like_the_reference():
    """mat_models.cpp:334-340: sheet / sheet-normal active tensions are rejected for models without directional distribution;
    an active-stress domain without nodal tensions is an error, not a silent passive run."""
    m = common.STRUCT_CASES[0][1]()
    Ag, Yg, Dg, Bf, fN = common.struct_state(m, 2)
    from svmultiphysics_b200.engine import Engine, Svb200Error
    e0 = Engine(0)
    rowPtr, colPtr = e0.lhsa(m.nNo, [m.IEN]); e0.close()
    eng = _engine(m, rowPtr, colPtr, 2, fN)
    eq, dmn = abi.struct_eq(1e-4), [abi.struct_domain(active_stress=True)]
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf)
    with pytest.raises(Svb200Error, match="svb200_set_active_tension"):
        eng.assemble(0, eq, dmn)
    eng.set_active_tension(np.ones(m.nNo), 0.5 * np.ones(m.nNo), None)
    with pytest.raises(Svb200Error, match="Directional distribution of active stress"):
        eng.assemble(0, eq, dmn)
    eng.set_active_tension(np.ones(m.nNo))
    eng.assemble(0, eq, dmn)
    eng.close()


@pytest.mark.parametrize("name", [c[0] for c in common.PRESTRESS_CASES])
def test_prestress_parity(name):
    """Nodal prestress pS0 in struct_3d / l_elas_3d and the pSn / pSa accumulators of a prestress equation (com_mod.pstEq):
    sv_struct.cpp:271-274, 635-680, 327-336; l_elas.cpp:321-338, 130-140 — against the committed vectors of the compiled
    reference (tests/golden/prestress.npz) and a live run of it."""
    cls = _oracle()
    golden = common.load_golden("prestress.npz")
    m, Ag, Yg, Dg, Bf, pS0, eq, dmn = common.prestress_case(name)
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(0)
    orc.set_prestress(pS0)
    orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    eng = _engine(m, rowPtr, colPtr)
    eng.set_prestress(pS0)
    for rep in range(2):          # twice: svb200_alloc must zero the accumulators like Integrator::initiator
        eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    for R0, V0 in ((orc.get_R(), orc.get_Val()), (golden[f"{name}/R"], golden[f"{name}/Val"])):
        assert common.rel_err(R1, R0) < ASM_TOL
        assert common.rel_err(V1, V0) < ASM_TOL
    if eq.reserved & abi.EQ_PRESTRESS:
        pSn1, pSa1 = eng.get_prestress()
        pSn0, pSa0 = orc.get_prestress()
        assert common.rel_err(pSn1, pSn0) < ASM_TOL and common.rel_err(pSa1, pSa0) < ASM_TOL
        assert common.rel_err(pSn1, golden[f"{name}/pSn"]) < ASM_TOL and common.rel_err(pSa1, golden[f"{name}/pSa"]) < ASM_TOL
    # removing the prestress restores the plain equation
    eng.set_prestress(None)
    orc.set_prestress(None)
    orc.alloc(3); orc.assemble(0, eq, dmn)
    eng.alloc(3); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), orc.get_R()) < ASM_TOL
    assert common.rel_err(eng.get_R(), R1) > 1e-3
    eng.close()


def test_fsi_solid_prestress_parity():
    """construct_fsi hands pS0 to struct_3d for the solid elements (fsi.cpp:127-129, 222)."""
    cls = _oracle()
    m, Ag, Yg, Dg, Bf = common.fsi_case()
    pS0 = np.asfortranarray(5.0e5 * np.random.default_rng(43).standard_normal((6, m.nNo)))
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN, eId=m.eId)
    rowPtr, colPtr = orc.build_graph(0)
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=abi.SCATTER_ATOMIC, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0), abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    Rplain = orc.get_R()
    orc.set_prestress(pS0)
    orc.alloc(4); orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    assert common.rel_err(Rplain, R0) > 1e-3
    eng = _engine(m, rowPtr, colPtr)
    eng.set_prestress(pS0)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), R0) < ASM_TOL
    assert common.rel_err(eng.get_Val(), V0) < ASM_TOL
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name,mk,dkw,nFn", common.STRUCT_HI_CASES, ids=[c[0] for c in common.STRUCT_HI_CASES])
def test_struct_on_quadratic_and_wedge_elements(name, mk, dkw, nFn, scatter):
    """struct_3d on curved TET10 (15 Gauss points), HEX20 / HEX27 (27; odd node count: another half-block rule) and WDG (6, lShpF)
    elements against the committed vectors of the compiled reference and a live run of it."""
    from svmultiphysics_b200.engine import Engine
    golden, tabs = common.load_golden("struct_hi.npz"), common.load_golden("fluid_hi.npz")
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
    et = name.split("_")[0]
    w, N, Nx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx"))
    eq, dmn = abi.struct_eq(1e-4, scatter=scatter), [abi.struct_domain(**dkw)]
    eng = Engine(0)
    eng.set_graph(golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"])
    eng.set_mesh(0, m.IEN, w, N, Nx, nFn=nFn, fN=fN)
    eng.set_coords(m.x)
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf)
    ya = common.active_tension(m, dmn[0].isoType) if dmn[0].active_stress else None
    if ya is not None:
        eng.set_active_tension(*ya)
    eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, golden[f"{name}/R"]) < ASM_TOL
    assert common.rel_err(V1, golden[f"{name}/Val"]) < ASM_TOL
    from oracle import refbind
    if refbind.have_ref():
        orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN, nFn=nFn, fN=fN); orc.build_graph(0)
        orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf)
        if ya is not None:
            orc.set_active_tension(*ya)
        orc.assemble(0, eq, dmn)
        assert common.rel_err(R1, orc.get_R()) < ASM_TOL and common.rel_err(V1, orc.get_Val()) < ASM_TOL
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(3); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1)
    eng.close()
