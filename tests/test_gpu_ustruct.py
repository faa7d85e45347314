"""GPU parity tests for the mixed velocity-pressure solid (ustruct): R / Val / Kd against the golden
vectors of the compiled reference (ustruct_3d_m/c + ustruct_do_assem), ustruct_r, and a GMRES solve."""
import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements
from tests import common

pytestmark = pytest.mark.gpu
ASM_TOL = 1e-12
VAL_GROUPS = ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14], [15])   # (v,v) (v,p) (p,v) (p,p): orders of magnitude apart


def _engine(m, rowPtr, colPtr, nFn=0, fN=None):
    from svmultiphysics_b200.engine import Engine
    e = Engine(0)
    e.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(m.eNoN)
    e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, nFn=nFn, fN=fN)
    e.set_coords(m.x)
    return e


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name,mk,dkw,nFn", common.USTRUCT_CASES, ids=[c[0] for c in common.USTRUCT_CASES])
def test_ustruct_assembly_matches_golden(name, mk, dkw, nFn, scatter):
    golden = common.load_golden("ustruct.npz")
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.ustruct_state(m, nFn)
    eq, dmn = abi.ustruct_eq(1e-3, scatter=scatter), [abi.ustruct_domain(**dkw)]
    eng = _engine(m, golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"], nFn, fN)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf)
    if dmn[0].active_stress:
        eng.set_active_tension(*common.active_tension(m, dmn[0].isoType))
    eng.assemble(0, eq, dmn)
    R1, V1, K1 = eng.get_R(), eng.get_Val(), eng.get_Kd()
    assert common.rel_err(R1, golden[f"{name}/R"]) < ASM_TOL
    assert common.rel_err(K1, golden[f"{name}/Kd"]) < ASM_TOL
    for rows in VAL_GROUPS:
        assert common.rel_err(V1[rows], golden[f"{name}/Val"][rows]) < ASM_TOL
    # ustruct_r, first Newton iteration: R -= Kd (amg Ad - Yg) / am; later iterations: unchanged
    eng.ustruct_r(eq, 1, common.ustruct_Ad(m))
    R2 = eng.get_R()
    assert common.rel_err(R2, golden[f"{name}/R_after_ustruct_r"]) < ASM_TOL
    eng.ustruct_r(eq, 2, common.ustruct_Ad(m))
    assert np.array_equal(eng.get_R(), R2)
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)            # alloc zeroes Kd together with R and Val
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1) and np.array_equal(eng.get_Kd(), K1)
    eng.close()


def test_ustruct_solve_parity():
    """ustruct/block_compression-like Newton iteration: assembly, ustruct_r, GMRES (dof = 4) against the compiled reference."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    from svmultiphysics_b200 import meshgen
    m = meshgen.box_hex8(5, 4, 4, (1.0, 1.0, 1.0))
    Ag, Yg, Dg, Bf, _ = common.ustruct_state(m)
    eq, dmn = abi.ustruct_eq(1e-3), [abi.ustruct_domain(E=1.0e6, nu=0.45, Kpen=1.0e6 / (3 * (1 - 0.9)), rho=1.2)]
    faces = []
    for k, name in enumerate(("X0", "Y0", "Z0")):
        val = np.ones((4, len(m.faces[name])), order="F"); val[k] = 0.0
        faces.append((abi.BC_DIR, m.faces[name], val))
    orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(len(faces))
    eng = _engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    Ad = common.ustruct_Ad(m)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn); orc.ustruct_r(1, Ad)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn); eng.ustruct_r(eq, 1, Ad)
    assert common.rel_err(eng.get_R(), orc.get_R()) < ASM_TOL
    V0, V1 = orc.get_Val(), eng.get_Val()
    for rows in VAL_GROUPS:
        assert common.rel_err(V1[rows], V0[rows]) < ASM_TOL
    ls = abi.ls_params(abi.LS_GMRES, mItr=4, sD=200, relTol=1e-6)      # 117 iterations inside one Krylov cycle in the reference
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
    X1, o1, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res)
    assert o1.RI.success == o0.RI.success and abs(o1.RI.itr - o0.RI.itr) <= max(2, o0.RI.itr // 20)
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-4                                 # solver tolerance 1e-6 on an ill-conditioned saddle-point system
    eng.close()


def test_ustruct_unsupported_options_fail_loudly():
    from svmultiphysics_b200 import meshgen
    from svmultiphysics_b200.engine import Svb200Error
    from oracle import refbind
    m = meshgen.box_tet4(2, 2, 2, (1.0, 1.0, 1.0))
    _, rowPtr, colPtr = common.make_oracle(refbind.OracleCase, m)
    eng = _engine(m, rowPtr, colPtr)
    Ag, Yg, Dg, Bf, _ = common.ustruct_state(m)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf)
    with pytest.raises(Svb200Error, match="dof = 4"):
        bad = abi.ustruct_eq(1e-3)
        bad.dof = 3
        eng.assemble(0, bad, [abi.ustruct_domain()])
    with pytest.raises(Svb200Error, match="Min fiber directions"):
        eng.assemble(0, abi.ustruct_eq(1e-3), [abi.ustruct_domain(isoType=abi.ISO_HGO)])
    eng.close()


def test_ustruct_device_resident_newton_loop():
    """Two time steps x three Newton iterations of a ustruct equation: (a) everything on the device (predictor incl. the Ad
    scaling, initiator, assemble, ustruct_r on the device-resident Ad, GMRES, the ustruct corrector that updates An, Yn, Dn
    and Ad) against (b) the same loop with the compiled reference's assembly + ustruct_r + GMRES and the numpy restatement
    of Integrator::predictor / initiator / corrector (sstEq branches, Integrator.cpp:626-630, 826-846)."""
    from oracle import genalpha_oracle as go, refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    from svmultiphysics_b200 import meshgen
    m = meshgen.box_hex8(4, 4, 3, (1.0, 1.0, 1.0))
    rng = np.random.default_rng(5)
    dt = 1e-3
    eq, dmn = abi.ustruct_eq(dt), [abi.ustruct_domain(E=1.0e6, nu=0.45, Kpen=1.0e6 / (3 * (1 - 0.9)), rho=1.2, f=(0.0, 0.0, -50.0))]
    qt = [abi.eq_time(0, 3, abi.PHYS_USTRUCT, 0.5)]
    assert abs(qt[0].af - eq.af) < 1e-15 and abs(qt[0].am - eq.am) < 1e-15
    faces = []
    for k, name in enumerate(("X0", "Y0", "Z0")):
        val = np.ones((4, len(m.faces[name])), order="F"); val[k] = 0.0
        faces.append((abi.BC_DIR, m.faces[name], val))
    orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(len(faces))
    eng = _engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    ls = abi.ls_params(abi.LS_GMRES, mItr=4, sD=200, relTol=1e-6)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    Ao = np.zeros((4, m.nNo), order="F")
    Yo = np.asfortranarray(np.vstack([1e-2 * rng.standard_normal((3, m.nNo)), np.zeros((1, m.nNo))]))
    Do = np.zeros((4, m.nNo), order="F"); Do[:3] = 1e-3 * rng.standard_normal((3, m.nNo))
    Ad = np.asfortranarray(1e-2 * rng.standard_normal((3, m.nNo)))
    An, Yn, Dn = Ao.copy(order="F"), Yo.copy(order="F"), Do.copy(order="F")
    Ag, Yg, Dg = (np.zeros_like(Ao) for _ in range(3))
    Bf = np.zeros((3, m.nNo), order="F")
    eng.set_state(Ag, Yg, Dg, Bf)
    eng.set_solution(abi.SOL_OLD, Ao, Yo, Do)
    eng.set_solution(abi.SOL_CURRENT, An, Yn, Dn)
    eng.set_ad(Ad)
    amg = (eq.gam - eq.am) / (eq.gam - 1.0)
    norms_dev, norms_ref = [], []
    for step in range(2):
        eng.predictor(qt, dt, 1)
        go.predictor(qt, dt, 1, Ao, Yo, Do, An, Yn, Dn, Ad)
        for it in range(3):
            eng.initiator(qt)
            eng.alloc(4); eng.assemble(0, eq, dmn); eng.ustruct_r(eq, it + 1)
            _, out1, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res, want_solution=False)
            eng.corrector(qt[0], dt)
            go.initiator(qt, Ao, Yo, Do, An, Yn, Dn, Ag, Yg, Dg)
            orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn); orc.ustruct_r(it + 1, Ad)
            Rd = (amg * Ad - Yg[0:3]) if it == 0 else np.zeros_like(Ad)          # ustruct.cpp:1773-1783
            X0, out0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
            go.corrector_ustruct(qt[0], dt, X0, Rd, An, Yn, Dn, Ad)
            norms_dev.append(out1.RI.iNorm); norms_ref.append(out0.RI.iNorm)
        eng.advance_time_step()
        Ao, Yo, Do = An.copy(order="F"), Yn.copy(order="F"), Dn.copy(order="F")
    # Newton history: 1.3e4 -> 98 -> 6e-3 in the reference.  The first two norms of a step are determined by the state to
    # ~1e-6 (the linear-solver tolerance); the third one IS the linear-solver error of the first two and only has to be as small.
    nd, nr = np.array(norms_dev).reshape(2, 3), np.array(norms_ref).reshape(2, 3)
    assert np.allclose(nd[:, 0], nr[:, 0], rtol=1e-5) and np.allclose(nd[:, 1], nr[:, 1], rtol=2e-3)
    assert np.all(nr[:, 2] < 1e-5 * nr[:, 0]) and np.all(nd[:, 2] < 1e-5 * nd[:, 0])
    A1, Y1, D1 = eng.get_solution(abi.SOL_CURRENT)
    assert common.rel_err(Y1[:3], Yn[:3]) < 1e-4 and common.rel_err(D1[:3], Dn[:3]) < 1e-4 and common.rel_err(eng.get_ad(), Ad) < 1e-4
    eng.close()


# ---- FSI with a ustruct wall: construct_fsi's ustruct_3d_m/c branch (fsi.cpp:243-262), tests/cases/fsi_ustruct ------------------
@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name", list(common.FSI_USTRUCT_CASES))
def test_fsi_ustruct_assembly_matches_golden(name, scatter):
    """One mesh with element domain ids: fluid elements through the ALE fluid kernel, solid elements through the ustruct kernel with
    the tDof = 7 state; R / Val / Kd and R after ustruct_r (ustruct-domain nodes only) against the compiled reference."""
    golden = common.load_golden("fsi_ustruct.npz")
    m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, Ad, flags = common.fsi_ustruct_case(name, scatter)
    eng = _engine(m, golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"], nFn, fN)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    R1, V1, K1 = eng.get_R(), eng.get_Val(), eng.get_Kd()
    G = golden[f"{name}/Val"]
    assert common.rel_err(K1, golden[f"{name}/Kd"]) < ASM_TOL
    # fluid and solid entries differ by many orders of magnitude: rows of fluid-only nodes, of solid-only nodes and of the interface
    # are compared separately, and Val by entry type inside each set
    fluid_nodes = np.unique(m.IEN[:, (m.eId & 2) == 0]); solid_nodes = np.where(flags != 0)[0]
    rowPtr = golden[f"{name}/rowPtr"]
    sets = (np.setdiff1d(fluid_nodes, solid_nodes), np.setdiff1d(solid_nodes, fluid_nodes), np.intersect1d(fluid_nodes, solid_nodes))
    assert all(len(x) > 0 for x in sets)
    for nodes in sets:
        assert common.rel_err(R1[:, nodes], golden[f"{name}/R"][:, nodes]) < ASM_TOL
        slots = np.concatenate([np.arange(rowPtr[a], rowPtr[a + 1]) for a in nodes])
        for rows in VAL_GROUPS:
            assert common.rel_err(V1[rows][:, slots], G[rows][:, slots]) < ASM_TOL
    with pytest.raises(RuntimeError):
        eng.ustruct_r(eq, 1, Ad)                  # an FSI equation needs the membership in the ustruct domains
    eng.set_node_flags(flags)
    eng.ustruct_r(eq, 1, Ad)
    R2 = eng.get_R()
    for nodes in sets:
        assert common.rel_err(R2[:, nodes], golden[f"{name}/R_after_ustruct_r"][:, nodes]) < ASM_TOL
    assert np.array_equal(R2[:, sets[0]], R1[:, sets[0]])          # fluid-only rows untouched
    assert not np.array_equal(R2[:, sets[1]], R1[:, sets[1]])
    Rd = eng.get_Rd()
    assert np.all(Rd[:, flags == 0] == 0.0) and np.abs(Rd[:, flags != 0]).min() > 0
    eng.ustruct_r(eq, 2, Ad)
    assert np.array_equal(eng.get_R(), R2) and np.all(eng.get_Rd() == 0.0)
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1) and np.array_equal(eng.get_Kd(), K1)
    eng.close()


def test_fsi_ustruct_two_meshes_and_solve_parity():
    """The layout of tests/cases/fsi_ustruct/pipe_3d: lumen mesh (fluid domain) + wall mesh (ustruct domain) over one node set,
    assembled mesh by mesh; then ustruct_r and a GMRES solve of the coupled system, against the compiled reference live."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    from svmultiphysics_b200.engine import Engine
    name = "nHK_ST91"
    m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, Ad, flags = common.fsi_ustruct_case(name)
    fl, so = np.where(m.eId == 1)[0], np.where(m.eId == 2)[0]
    faces = common.dirichlet_faces(m)
    orc = refbind.RefCase(); orc.set_coords(m.x)
    orc.add_mesh(np.asfortranarray(m.IEN[:, fl]), eId=m.eId[fl]); orc.add_mesh(np.asfortranarray(m.IEN[:, so]), eId=m.eId[so])
    rowPtr, colPtr = orc.build_graph(len(faces))
    eng = Engine(0)
    eng.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(4)
    eng.set_mesh(0, np.asfortranarray(m.IEN[:, fl]), w, N, Nx, eId=m.eId[fl])
    eng.set_mesh(1, np.asfortranarray(m.IEN[:, so]), w, N, Nx, eId=m.eId[so])
    eng.set_coords(m.x)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn); orc.assemble(1, eq, dmn); orc.ustruct_r(1, Ad)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn); eng.assemble(1, eq, dmn)
    eng.set_node_flags(flags); eng.ustruct_r(eq, 1, Ad)
    golden = common.load_golden("fsi_ustruct.npz")
    assert common.rel_err(orc.get_Kd(), golden[f"{name}/Kd"]) < 1e-13          # two meshes = one mesh with domain ids
    assert common.rel_err(eng.get_Kd(), orc.get_Kd()) < ASM_TOL
    R0, R1, V0, V1 = orc.get_R(), eng.get_R(), orc.get_Val(), eng.get_Val()
    solid_nodes = np.where(flags != 0)[0]; fluid_only = np.where(flags == 0)[0]
    for nodes in (solid_nodes, fluid_only):
        assert common.rel_err(R1[:, nodes], R0[:, nodes]) < ASM_TOL
        slots = np.concatenate([np.arange(rowPtr[a], rowPtr[a + 1]) for a in nodes])
        for rows in VAL_GROUPS:
            assert common.rel_err(V1[rows][:, slots], V0[rows][:, slots]) < 1e-11
    ls = abi.ls_params(abi.LS_GMRES, mItr=4, sD=150, relTol=1e-6)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
    X1, o1, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res)
    assert o1.RI.success == o0.RI.success and abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert abs(o1.RI.itr - o0.RI.itr) <= max(2, o0.RI.itr // 20)
    assert common.rel_err(X1, X0) < 1e-4
    eng.close()
