"""GPU parity tests (run on a B200 with -m gpu): CUDA path through the C ABI vs the oracle."""
import ctypes as C

import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements, meshgen
from tests import common

pytestmark = pytest.mark.gpu

ASM_TOL = 1e-12   # north_star: assembled R / Val within 1e-12 relative


def _oracle():
    from oracle import refbind
    return refbind.RefCase if refbind.have_ref() else refbind.OracleCase


CASES = common.FLUID_CASES


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name,visc,Kd,f,tDof,mv", CASES, ids=[c[0] for c in CASES])
def test_fluid_assembly_parity(name, visc, Kd, f, tDof, mv, scatter):
    m, Ag, Yg, Dg, Bf = common.fluid_case(tDof=tDof)
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m)
    dmn = [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter)
    orc.alloc(4)
    orc.set_state(Ag, Yg, Dg, Bf)
    orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()

    eng = common.make_engine(m, rowPtr, colPtr)
    eng.alloc(4)
    eng.set_state(Ag, Yg, Dg, Bf)
    eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, R0) < ASM_TOL
    assert common.rel_err(V1, V0) < ASM_TOL
    if scatter == abi.SCATTER_COLORED:
        # deterministic mode: bitwise reproducible
        eng.alloc(4)
        eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1)
    assert eng.launch_count > 0
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
def test_grouped_scatter_on_a_mesh_without_numbering_locality(scatter):
    """The grouped TET4 scatter pre-reduces the contributions of 128 CONSECUTIVE elements; its plan must stay correct when the element
    and node numbering has no locality at all (a group then has ~1280 distinct tangent targets and ~510 distinct nodes instead of ~710
    and ~90, the plan's worst case): elements and nodes of the cylinder randomly renumbered, R / Val against the oracle on the SAME
    renumbered mesh and against the naturally numbered run mapped through the permutations."""
    import copy
    m0 = meshgen.cylinder_tet4(10, 12)                       # 7,200 tets = 57 groups
    Ag0, Yg0, Dg0 = meshgen.poiseuille_state(m0)
    rng = np.random.default_rng(5)
    pn = rng.permutation(m0.nNo)                             # old node -> new node
    pe = rng.permutation(m0.nEl)
    m = copy.copy(m0)
    m.x = np.asfortranarray(np.empty_like(m0.x)); m.x[:, pn] = m0.x
    m.IEN = np.asfortranarray(pn[m0.IEN][:, pe].astype(np.int32))
    m.faces = {k: pn[v].astype(np.int32) for k, v in m0.faces.items()}
    Ag, Yg = np.asfortranarray(np.empty_like(Ag0)), np.asfortranarray(np.empty_like(Yg0))
    Ag[:, pn], Yg[:, pn] = Ag0, Yg0
    eq, dmn = abi.fluid_eq(0.005, scatter=scatter), [abi.fluid_domain(K_darcy=0.7, f=(0.1, 0.2, -0.3))]
    res = []
    for mm, A, Y in ((m0, Ag0, Yg0), (m, Ag, Yg)):
        orc, rowPtr, colPtr = common.make_oracle(_oracle(), mm)
        orc.alloc(4); orc.set_state(A, Y, None, None); orc.assemble(0, eq, dmn)
        eng = common.make_engine(mm, rowPtr, colPtr)
        eng.alloc(4); eng.set_state(A, Y, None, None); eng.assemble(0, eq, dmn)
        R1, V1 = eng.get_R(), eng.get_Val()
        assert common.rel_err(R1, orc.get_R()) < ASM_TOL and common.rel_err(V1, orc.get_Val()) < ASM_TOL
        if scatter == abi.SCATTER_COLORED:
            eng.alloc(4); eng.assemble(0, eq, dmn)
            assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1)
        res.append((R1, V1, rowPtr, colPtr))
        eng.close()
    (Rn, Vn, rp0, cp0), (Rs, Vs, rp1, cp1) = res
    assert common.rel_err(Rs[:, pn], Rn) < ASM_TOL           # same physics, renumbered
    # block (a, b) of the natural run = block (pn[a], pn[b]) of the shuffled run
    rows0 = np.repeat(np.arange(m0.nNo), np.diff(rp0))
    rows1 = np.repeat(np.arange(m.nNo), np.diff(rp1))
    lut = dict(zip(zip(rows1.tolist(), cp1.tolist()), range(len(cp1))))
    idx = np.array([lut[(int(pn[a]), int(pn[b]))] for a, b in zip(rows0, cp0)])
    assert common.rel_err(Vs[:, idx], Vn) < ASM_TOL


def test_assemble_host_pipelined_matches_plain_sequence():
    """svb200_assemble_host (state uploaded in node chunks on a copy stream, element groups launched chunk by chunk behind it,
    finished residual rows streamed back) leaves the same R and Val as set_state + alloc + assemble + download."""
    m = meshgen.cylinder_tet4(16, 24)                      # 36,864 tets = 288 groups: enough for the pipelined path
    Ag, Yg, Dg = meshgen.poiseuille_state(m)
    from svmultiphysics_b200.engine import Engine
    eng = Engine(0)
    rowPtr, colPtr = eng.lhsa(m.nNo, [m.IEN])
    eng.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(4)
    eng.set_mesh(0, m.IEN, w, N, Nx)
    eng.set_coords(m.x)
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    eng.alloc(4); eng.set_state(Ag, Yg); eng.assemble(0, eq, dmn)
    R0, V0 = eng.get_R(), eng.get_Val()
    # other state on the device, then the pipelined call with the right one: it must upload everything it reads
    eng.set_state(np.asfortranarray(Ag * 0.0 + 7.0), np.asfortranarray(Yg * 0.0 - 3.0))
    Ah, Yh = np.asfortranarray(Ag.copy()), np.asfortranarray(Yg.copy())
    Rh = np.full((4, m.nNo), np.nan, order="F")
    for a in (Ah, Yh, Rh):
        eng.pin(a)
    for rep in range(3):
        Rh[:] = np.nan
        eng.assemble_host(0, eq, dmn, Ah, Yh, Rh)
        assert common.rel_err(Rh, R0) < 1e-13
        assert common.rel_err(eng.get_R(), R0) < 1e-13 and common.rel_err(eng.get_Val(), V0) < 1e-13
    for a in (Ah, Yh, Rh):
        eng.unpin(a)
    eng.close()


def test_lhsa_matches_reference():
    m, *_ = common.fluid_case()
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m)
    from svmultiphysics_b200.engine import Engine
    eng = Engine(0)
    rp, cp = eng.lhsa(m.nNo, [m.IEN])
    assert np.array_equal(rp, rowPtr) and np.array_equal(cp, colPtr)
    eng.close()


def test_spmv_parity():
    m, Ag, Yg, Dg, Bf = common.fluid_case()
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m)
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    V0 = orc.get_Val()
    U = np.asfortranarray(np.random.default_rng(3).standard_normal((4, m.nNo)))
    KU0 = orc.spmv(4, U)
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.alloc(4)
    eng.put_Val(V0, 4)
    KU1 = eng.spmv(4, U)
    assert common.rel_err(KU1, KU0) < 1e-13
    eng.close()


@pytest.mark.parametrize("ls_type,kw", [
    (abi.LS_GMRES, dict(mItr=100, sD=50, relTol=1e-8)),
    (abi.LS_GMRES, dict(mItr=20, sD=10, relTol=1e-10)),    # restarts
    (abi.LS_GMRES, dict(mItr=5, sD=250, relTol=1e-5)),     # one cycle: iteration count and final residual must match exactly
    (abi.LS_BICGS, dict(mItr=400, relTol=1e-8)),
], ids=["gmres50", "gmres10_restart", "gmres250_one_cycle", "bicgs"])
def test_fluid_solve_parity(ls_type, kw):
    m, Ag, Yg, Dg, Bf = common.fluid_case()
    faces = common.dirichlet_faces(m)
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m, nFaces=len(faces))
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    ls = abi.ls_params(ls_type, **kw)
    incL = np.ones(len(faces), dtype=np.int32)
    res = np.zeros(len(faces))
    X0, out0, _ = orc.solve(4, ls_type, ls, incL, res)

    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    X1, out1, hist = eng.solve(4, ls_type, ls, incL, res, hist_cap=512)
    assert out1.RI.success == out0.RI.success
    assert abs(out1.RI.iNorm - out0.RI.iNorm) <= 1e-10 * out0.RI.iNorm
    if ls_type == abi.LS_BICGS:
        # BiCGStab's residual is non-monotone and amplifies last-bit differences of the dot products: the
        # iteration at which it dips under the tolerance may move by a few steps; the answer may not.
        assert abs(out1.RI.itr - out0.RI.itr) <= max(3, out0.RI.itr // 20)
        assert out1.RI.fNorm < ls.RI.relTol * out1.RI.iNorm
        assert common.rel_err(X1, X0) < 1e-6
        eng.close()
        return
    if not out0.RI.success:
        # did not converge within mItr cycles: both sides ran the same number of iterations
        assert out1.RI.itr == out0.RI.itr
        assert abs(out1.RI.fNorm - out0.RI.fNorm) <= 0.15 * out0.RI.fNorm
    else:
        # classical Gram-Schmidt (the reference's choice) amplifies last-bit differences of the dots and of the
        # atomically scattered Val over hundreds of iterations: the stopping test may fire one step earlier or later
        # per cycle (observed 966/967 over 20 restarts, 242/241 inside one 250-dimensional cycle, final residual
        # within 1.5 %); the contract is the set tolerance
        assert abs(out1.RI.itr - out0.RI.itr) <= max(1, out0.RI.itr // 25)
        assert out1.RI.fNorm <= ls.RI.relTol * out1.RI.iNorm
        assert abs(out1.RI.fNorm - out0.RI.fNorm) <= 0.15 * out0.RI.fNorm
    # residual HISTORY against the C restatement (bit-identical to the reference, and it records |err(i+1)| per
    # iteration like gmres.cpp under debug_gmres_v): the early iterations agree to round-off, the drift grows slowly
    from oracle import refbind
    if ls_type == abi.LS_GMRES and len(hist) > 0:
        oc, _, _ = common.make_oracle(refbind.OracleCase, m, nFaces=len(faces))
        for i, (g, nodes, val) in enumerate(faces):
            oc.set_face(i, g, nodes, val)
        oc.alloc(4); oc.set_state(Ag, Yg, Dg, Bf); oc.assemble(0, eq, dmn)
        _, _, hist0 = oc.solve(4, ls_type, ls, incL, res, hist_cap=512)
        n = min(len(hist), len(hist0))
        drift = np.abs(hist[:n] - hist0[:n]) / hist0[:n]
        print(f"history drift: first 20 {drift[:20].max():.2e}, first 50 {drift[:50].max():.2e}, all {n}: {drift.max():.2e}")
        assert drift[:20].max() < 1e-9
        assert drift[:min(n, ls.RI.sD)].max() < 0.05
    # solution agrees to the solver tolerance (both are relTol-accurate solutions of the same system)
    assert common.rel_err(X1, X0) < max(1e-6, 20 * ls.RI.relTol)
    eng.close()


def _pipe_faces(m):
    """pipe_RCR_3d-like linear-solver faces: Dirichlet wall + inlet, coupled Neumann (resistance) outlet whose
    val holds the nodal normal integrals (here: unit z normal times a lumped area weight)."""
    faces = common.dirichlet_faces(m)
    out = m.faces["outlet"]
    val = np.zeros((3, len(out)), order="F")
    val[2] = 4.0 * np.pi / len(out)
    faces.append((abi.BC_NEU, out, val))
    return faces


@pytest.mark.parametrize("ls_type,kw,res_out", [
    (abi.LS_NS, dict(mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17)), 0.0),
    (abi.LS_NS, dict(mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17)), 0.8),
    (abi.LS_NS, dict(mItr=10, sD=100, relTol=0.4, absTol=1e-10), 0.8),       # FSILS defaults
    (abi.LS_GMRES, dict(mItr=100, sD=50, relTol=1e-8), 0.8),                  # coupled face inside gmres_v
], ids=["ns_pipe", "ns_pipe_resistance", "ns_defaults_resistance", "gmres_resistance"])
def test_fluid_ns_and_coupled_face_parity(ls_type, kw, res_out):
    m, Ag, Yg, Dg, Bf = common.fluid_case()
    faces = _pipe_faces(m)
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m, nFaces=len(faces))
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    ls = abi.ls_params(ls_type, **kw)
    incL = np.ones(len(faces), dtype=np.int32)
    res = np.array([0.0, 0.0, res_out])
    try:
        X0, out0, _ = orc.solve(4, ls_type, ls, incL, res)
    except RuntimeError as ex:
        if "not restated" in str(ex):
            pytest.skip("oracle restatement lacks this solver and libsvref.so is absent")
        raise
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    X1, out1, _ = eng.solve(4, ls_type, ls, incL, res)
    assert out1.RI.success == out0.RI.success
    if ls_type == abi.LS_GMRES and out0.RI.itr > ls.RI.sD + 1:
        # dozens of restarts: summation-order differences may move the final stopping test by a few steps
        assert abs(out1.RI.itr - out0.RI.itr) <= max(2, out0.RI.itr // 25)   # atomic scatter: last-bit run-to-run differences in Val
    else:
        assert out1.RI.itr == out0.RI.itr
    assert abs(out1.RI.iNorm - out0.RI.iNorm) <= 1e-10 * out0.RI.iNorm
    if ls_type == abi.LS_GMRES and out0.RI.itr > ls.RI.sD + 1:
        # after ~20 restarts the residual at which the stopping test first fires differs by up to one
        # iteration's reduction factor; the contract is the set tolerance itself
        assert out1.RI.fNorm <= ls.RI.relTol * out1.RI.iNorm
        assert abs(out1.RI.fNorm - out0.RI.fNorm) <= 0.15 * out0.RI.fNorm
    else:
        assert abs(out1.RI.fNorm - out0.RI.fNorm) <= 2e-2 * out0.RI.fNorm
    if ls_type == abi.LS_NS:
        # accumulated inner iteration counts: a CG/GMRES stopping test that lands within round-off of its
        # tolerance may fire one step earlier or later
        assert abs(out1.GM.itr - out0.GM.itr) <= max(2, out0.GM.itr // 50)
        assert abs(out1.CG.itr - out0.CG.itr) <= max(2, out0.CG.itr // 50)
        assert abs(out1.Resm - out0.Resm) <= 1 and abs(out1.Resc - out0.Resc) <= 1
    # both are solutions of the same system to the solver tolerance; compare to that tolerance
    tol = 50 * ls.RI.relTol if ls_type == abi.LS_NS else 1e-6
    assert common.rel_err(X1, X0) < tol
    eng.close()


@pytest.mark.parametrize("ls_type,kw", [
    (abi.LS_GMRES, dict(mItr=100, sD=50, relTol=1e-8)),
    (abi.LS_NS, dict(mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17))),
], ids=["gmres_cap", "ns_cap"])
def test_capped_coupled_face_parity(ls_type, kw):
    """A coupled (resistance) outlet WITH a capping surface (FSILS_faceType::has_cap, fils_struct.hpp:131-143): the cap's normal
    integrals enter the flow-rate sum of add_bc_mul (add_bc_mul.cpp:62-81, 102-111) and are scaled by precond_diag
    (precond.cpp:229-237).  The cap here is the free interior nodes of the plane one layer upstream of the outlet."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so (the restatement has no coupled faces)")
    m, Ag, Yg, Dg, Bf = common.fluid_case()
    n, _, nz = m.lattice
    faces = _pipe_faces(m)
    plane = np.arange((n + 1) ** 2) + (n + 1) ** 2 * (nz - 1)
    cap = np.setdiff1d(plane, m.faces["wall"]).astype(np.int32)
    cap_val = np.zeros((3, len(cap)), order="F")
    cap_val[2] = 2.0 * np.pi / len(cap)
    cap_val[0] = 0.1 * np.pi / len(cap)
    orc, rowPtr, colPtr = common.make_oracle(refbind.RefCase, m, nFaces=len(faces))
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val)
    orc.set_face_cap(2, cap, cap_val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    ls = abi.ls_params(ls_type, **kw)
    incL, res = np.ones(len(faces), dtype=np.int32), np.array([0.0, 0.0, 0.8])
    X0, out0, _ = orc.solve(4, ls_type, ls, incL, res)
    # the same system WITHOUT the cap: the cap must change the answer, or the test proves nothing
    orc2, _, _ = common.make_oracle(refbind.RefCase, m, nFaces=len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc2.set_face(i, g, nodes, val)
    orc2.alloc(4); orc2.set_state(Ag, Yg, Dg, Bf); orc2.assemble(0, eq, dmn)
    Xn, _, _ = orc2.solve(4, ls_type, ls, incL, res)
    assert common.rel_err(Xn, X0) > 1e-3

    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.set_face_cap(2, cap, cap_val)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    X1, out1, _ = eng.solve(4, ls_type, ls, incL, res)
    eng.close()
    print(f"cap: itr {out1.RI.itr} vs {out0.RI.itr}, fNorm {out1.RI.fNorm:.6e} vs {out0.RI.fNorm:.6e}, relerr X {common.rel_err(X1, X0):.2e}; "
          f"without the cap the answer differs by {common.rel_err(Xn, X0):.2e}")
    assert out1.RI.success == out0.RI.success
    assert abs(out1.RI.iNorm - out0.RI.iNorm) <= 1e-10 * out0.RI.iNorm
    if ls_type == abi.LS_NS:
        assert out1.RI.itr == out0.RI.itr
        assert abs(out1.RI.fNorm - out0.RI.fNorm) <= 2e-2 * out0.RI.fNorm
        assert common.rel_err(X1, X0) < 50 * ls.RI.relTol
    else:
        # 2100 iterations over 42 stagnating restarts: the atomically scattered Val changes the count from run to run
        # (measured 2069 / 2091 / 2104 / 2175 on the same box, profiles/r2q: gpurun_out log), hence the wide band here; the
        # equality-grade statement for GMRES lives in test_gpu_solver_equality.py
        assert abs(out1.RI.itr - out0.RI.itr) <= out0.RI.itr // 12
        assert out1.RI.fNorm <= ls.RI.relTol * out1.RI.iNorm
        assert common.rel_err(X1, X0) < 1e-6


@pytest.mark.parametrize("ls_type,kw", [
    (abi.LS_GMRES, dict(mItr=10, sD=250, relTol=1e-4)),     # converges inside one cycle (241 iterations in the reference)
    (abi.LS_GMRES, dict(mItr=100, sD=50, relTol=1e-8)),     # stagnating restarts (4328 iterations in the reference)
    (abi.LS_BICGS, dict(mItr=400, relTol=1e-8)),
], ids=["rcs_gmres250", "rcs_gmres50_restarts", "rcs_bicgs"])
def test_precond_rcs_parity(ls_type, kw):
    """precond_rcs (linear_solver/precond.cpp:251-523) on the device against the compiled reference: the equilibrated
    matrix fsils_solve leaves in Val, the preconditioned initial norm, the iteration count and the solution."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("precond_rcs is checked against libsvref.so only")
    m, Ag, Yg, Dg, Bf = common.fluid_case()
    faces = common.dirichlet_faces(m)
    orc, rowPtr, colPtr = common.make_oracle(refbind.RefCase, m, nFaces=len(faces))
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    R0, V0 = orc.get_R(), orc.get_Val()
    ls = abi.ls_params(ls_type, **kw)
    incL, res = np.ones(len(faces), dtype=np.int32), np.zeros(len(faces))
    X0, out0, _ = orc.solve(4, ls_type, ls, incL, res, prec=abi.PREC_RCS)
    Vs0 = orc.get_Val()

    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.alloc(4)
    eng.put_Val(V0, 4); eng.put_R(R0)        # identical input bits: the comparison isolates the preconditioner + solver
    X1, out1, _ = eng.solve(4, ls_type, ls, incL, res, prec=abi.PREC_RCS)
    Vs1 = eng.get_Val()
    assert common.rel_err(Vs1, Vs0) < 1e-14
    assert abs(np.abs(Vs1).max() - 1.0) < 1.0            # equilibrated: max norms O(1)
    assert abs(out1.RI.iNorm - out0.RI.iNorm) <= 1e-12 * out0.RI.iNorm
    assert out1.RI.success == out0.RI.success
    if ls_type == abi.LS_BICGS:
        assert abs(out1.RI.itr - out0.RI.itr) <= max(3, out0.RI.itr // 20)
    elif out0.RI.itr > ls.RI.sD + 1:
        # ~85 stagnating restarts of GMRES(50): the number of cycles is sensitive to the summation order of the dots
        # (measured 4019 vs 4328); the contract is the set tolerance and the answer
        assert abs(out1.RI.itr - out0.RI.itr) <= out0.RI.itr // 8
        assert out1.RI.fNorm <= ls.RI.relTol * out1.RI.iNorm
    else:
        assert abs(out1.RI.itr - out0.RI.itr) <= 1           # one 250-dimensional cycle of classical Gram-Schmidt
        assert out1.RI.fNorm <= ls.RI.relTol * out1.RI.iNorm
        assert abs(out1.RI.fNorm - out0.RI.fNorm) <= 0.05 * out0.RI.fNorm
    assert common.rel_err(X1, X0) < max(1e-6, 20 * ls.RI.relTol)
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED], ids=["atomic", "colored"])
@pytest.mark.parametrize("case", common.FLUID_GEN_CASES, ids=[c[0] for c in common.FLUID_GEN_CASES])
def test_general_element_fluid_assembly_parity(case, scatter):
    """HEX8 fluid (gnn + gn_nxx per Gauss point, second-derivative terms, stale-Nwxx continuity loop) and TET4 through
    the same general kernel, against what the unmodified reference assembled (tests/golden/fluid_gen.npz)."""
    golden = common.load_golden("fluid_gen.npz")
    name, mk, visc, Kd, f, tDof, mv = case
    m = mk()
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter, general=True)
    dmn = [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, golden[f"{name}/R"]) < ASM_TOL
    assert common.rel_err(V1, golden[f"{name}/Val"]) < ASM_TOL
    if m.eNoN == 4:
        # the specialised TET4 kernel and the general kernel are two independent derivations of the same element
        eng.alloc(4); eng.assemble(0, abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter), dmn)
        assert common.rel_err(eng.get_Val(), V1) < ASM_TOL and common.rel_err(eng.get_R(), R1) < ASM_TOL
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1)
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED], ids=["atomic", "colored"])
@pytest.mark.parametrize("case", common.FLUID_HI_CASES, ids=[c[0] for c in common.FLUID_HI_CASES])
def test_quadratic_and_wedge_fluid_assembly_parity(case, scatter):
    """VMS fluid on curved TET10 (15 Gauss points; tests/cases/fluid/quadratic_tet10), HEX20 / HEX27 (27) and WDG (6, with the
    reference's lShpF behaviour) elements: the general kernel with nG != eNoN against the committed vectors of the compiled
    reference (tests/golden/fluid_hi.npz) and, when it is present, a live run of it."""
    from svmultiphysics_b200.engine import Engine
    golden = common.load_golden("fluid_hi.npz")
    name, mk, visc, Kd, f, tDof, mv = case
    m = mk()
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
    et = name.split("_")[0]
    w, N, Nx, Nxx = (golden[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter)
    dmn = [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    eng = Engine(0)
    rp, cp = eng.lhsa(m.nNo, [m.IEN])                     # lhsa on the device for eNoN = 6 / 10 / 20 / 27 as well
    assert np.array_equal(rp, rowPtr) and np.array_equal(cp, colPtr)
    eng.set_graph(rowPtr, colPtr)
    eng.set_mesh(0, m.IEN, w, N, Nx, Nxx=Nxx)
    eng.set_coords(m.x)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, golden[f"{name}/R"]) < ASM_TOL
    assert common.rel_err(V1, golden[f"{name}/Val"]) < ASM_TOL
    from oracle import refbind
    if refbind.have_ref():
        orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN); orc.build_graph(0)
        orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
        assert common.rel_err(R1, orc.get_R()) < ASM_TOL and common.rel_err(V1, orc.get_Val()) < ASM_TOL
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_Val(), V1) and np.array_equal(eng.get_R(), R1)
    eng.close()


def test_hex8_fluid_newton_iteration_parity():
    """Assembly + GMRES on a HEX8 fluid mesh against the compiled reference."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs libsvref.so")
    name, mk, visc, Kd, f, tDof, mv = common.FLUID_GEN_CASES[0]
    m = mk()
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("X0", "Y0", "Y1", "Z0", "Z1")]
    orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(len(faces))
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    # relTol 1e-6: on this 48-node mesh the GMRES residual stagnates at 1.6e-8 |r0| from iteration 23 on, so a tighter
    # tolerance sits on the plateau and the iteration count then depends on the last bits of the atomically scattered Val
    # (tools/stress_hex8.py: 23 iterations in 200/200 runs with the coloured scatter, 24 / 85 / 115 with atomics)
    ls = abi.ls_params(abi.LS_GMRES, mItr=10, sD=100, relTol=1e-6)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    R0 = orc.get_R()
    X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), R0) < ASM_TOL
    X1, o1, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res)
    assert o1.RI.success == o0.RI.success and abs(o1.RI.itr - o0.RI.itr) <= 1
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-6
    eng.close()


def test_edge_cases_and_error_behaviour():
    """Edge cases of the reference path: a degenerate element makes construct_fluid throw "Jacobian for element e is < 0."
    (fluid.cpp:637-639) -> SVB200_ERR_NUMERIC with the same text; a zero right-hand side returns immediately with R
    untouched (gmres.cpp:470-475); a zero-node face and a domain without elements are accepted; a missing res for a
    Neumann face is the reference's "res is required" error (solve.cpp:69-71)."""
    from svmultiphysics_b200.engine import Svb200Error
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=3, nz=3)
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m)
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    # (1) degenerate element: nodes 0 and 1 of element 5 moved onto node 3, so that a column of dx/dxi is exactly zero
    # (utils::is_zero only fires for |Jac| < 10 eps^2: a merely flat element leaves round-off in the determinant)
    x = m.x.copy()
    e_bad = 5
    x[:, m.IEN[0, e_bad]] = x[:, m.IEN[3, e_bad]]
    x[:, m.IEN[1, e_bad]] = x[:, m.IEN[3, e_bad]]
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_coords(np.asfortranarray(x))
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf)
    with pytest.raises(Svb200Error, match=r"\[construct_fluid\] Jacobian for element \d+ is < 0\."):
        eng.assemble(0, eq, dmn)
    with pytest.raises(Svb200Error, match=r"\[construct_fluid\] Jacobian for element \d+ is < 0\."):
        eng.alloc(4); eng.assemble(0, abi.fluid_eq(0.005, general=True), dmn)
    # the context stays usable: good coordinates assemble and match the oracle
    eng.set_coords(m.x)
    eng.alloc(4); eng.assemble(0, eq, dmn)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_Val(), orc.get_Val()) < ASM_TOL
    # (2) zero right-hand side: success, zero iterations, R untouched
    eng.put_R(np.zeros((4, m.nNo), order="F"))
    ls = abi.ls_params(abi.LS_GMRES, mItr=10, sD=20, relTol=1e-8, absTol=1e-10)
    X, o, _ = eng.solve(4, abi.LS_GMRES, ls)
    assert o.RI.success == 1 and o.RI.itr == 0 and o.RI.iNorm == 0.0 and not X.any()
    # (3) faces: an empty Dirichlet face is legal; a Neumann face without res is the reference's error
    eng.set_num_faces(2)
    eng.set_face(0, abi.BC_DIR, np.zeros(0, np.int32), np.zeros((3, 0), order="F"))
    out = m.faces["outlet"]
    eng.set_face(1, abi.BC_NEU, out, np.ones((3, len(out)), order="F"))
    eng.alloc(4); eng.assemble(0, eq, dmn)
    with pytest.raises(Svb200Error, match="res is required for Neu surfaces"):
        eng._call("svb200_solve", C.c_int32(4), C.c_int32(abi.LS_GMRES), C.c_int32(abi.PREC_FSILS), C.byref(ls), C.c_int32(2),
                  None, None, None, None)
    # (4) a domain list whose fluid domain owns no element: nothing is assembled, nothing fails
    eId = np.full(m.nEl, 2, np.int32)                       # every element in domain bit 1
    eng2 = common.make_engine(m, rowPtr, colPtr)
    w, N, Nx = elements.tables(4)
    eng2.set_mesh(0, m.IEN, w, N, Nx, eId=eId)
    eng2.alloc(4); eng2.set_state(Ag, Yg, Dg, Bf)
    eng2.assemble(0, eq, [abi.fluid_domain(Id=0), abi.struct_domain(Id=1)])   # fluid equation: struct elements are skipped
    assert not eng2.get_Val().any() and not eng2.get_R().any()
    eng.close(); eng2.close()


# ---- URIS valves: penalty terms of fluid_3d_m / fluid_3d_c with the per-Gauss-point valve factor (tests/cases/uris) --------------
@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name", [c[0] for c in common.URIS_CASES])
def test_uris_valves_assembly_matches_golden(name, scatter):
    """construct_fluid (TET4, HEX8 with a moving mesh) and the fluid elements of construct_fsi with two URIS valves (ramped thickness +
    valve velocity; scaffold): R / Val against the compiled reference (tests/golden/fluid_uris.npz).  With valves set the TET4 mesh
    runs through the per-Gauss-point kernel; removing them restores the closed-form kernel's result."""
    from svmultiphysics_b200 import elements
    from svmultiphysics_b200.engine import Engine
    golden = common.load_golden("fluid_uris.npz")
    m, Ag, Yg, Dg, Bf, eq, dmn = common.uris_case(name, scatter)
    raw, dev, sdf, udf, vel = common.uris_valves(m)
    eng = Engine(0)
    eng.set_graph(golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"])
    w, N, Nx = elements.tables(m.eNoN)
    eng.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, Nxx=elements.nxx_tables(m.eNoN) if m.eNoN != 4 else None)
    eng.set_coords(m.x)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    R0, V0 = eng.get_R(), eng.get_Val()
    eng.set_uris(dev, sdf, udf, vel)
    eng.alloc(4); eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    fl = np.arange(m.nNo) if m.eId is None else np.unique(m.IEN[:, (m.eId & 1) != 0])
    rowPtr = golden[f"{name}/rowPtr"]
    slots = np.concatenate([np.arange(rowPtr[a], rowPtr[a + 1]) for a in fl])
    assert common.rel_err(R1[:, fl], golden[f"{name}/R"][:, fl]) < 1e-12
    assert common.rel_err(V1[:, slots], golden[f"{name}/Val"][:, slots]) < 1e-12
    assert common.rel_err(R1, golden[f"{name}/R"]) < 1e-12 and common.rel_err(V1, golden[f"{name}/Val"]) < 1e-12
    assert common.rel_err(R1[:, fl], R0[:, fl]) > 1e-3                  # the valves matter
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_R(), R1) and np.array_equal(eng.get_Val(), V1)
    eng.set_uris([])
    eng.alloc(4); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), R0) < 1e-13 and common.rel_err(eng.get_Val(), V0) < 1e-13
    eng.close()


# ---- fitted RIS: an open resistive immersed surface couples the twin nodes of two lumen meshes (tests/cases/ris) ---------------------
@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
def test_open_ris_surface_matches_the_reference(scatter):
    """construct_fluid on two lumen meshes with ris::doassem_ris (the reference's own ris.cpp, tests/golden/ris.npz): surface open
    (twin rows receive each other's element rows, mapped columns replaced by twins), closed (plain assembly), and opened again after a
    closed step (svb200_set_ris is called whenever the surface changes state)."""
    from svmultiphysics_b200 import elements
    from svmultiphysics_b200.engine import Engine
    g = common.load_golden("ris.npz")
    x, IENs, mp, Ag, Yg, Bf, eq, dmn = common.ris_case(scatter)
    eng = Engine(0)
    eng.set_graph(g["rowPtr"], g["colPtr"])
    w, N, Nx = elements.tables(4)
    for iM, I in enumerate(IENs):
        eng.set_mesh(iM, I, w, N, Nx)
    eng.set_coords(x)

    def run():
        eng.alloc(4); eng.set_state(Ag, Yg, None, Bf)
        for iM in range(len(IENs)):
            eng.assemble(iM, eq, dmn)
        return eng.get_R(), eng.get_Val()

    R0, V0 = run()                                   # no plan: the two lumens uncoupled
    assert common.rel_err(R0, g["closed/R"]) < 1e-12 and common.rel_err(V0, g["closed/Val"]) < 1e-12
    eng.set_ris([mp], [0])
    R1, V1 = run()
    assert common.rel_err(R1, g["open/R"]) < 1e-12 and common.rel_err(V1, g["open/Val"]) < 1e-12
    assert common.rel_err(R1[:, mp[0]], R1[:, mp[1]]) < 1e-13          # twins carry the same residual
    eng.set_ris([mp], [1])
    R2, V2 = run()
    assert common.rel_err(R2, g["closed/R"]) < 1e-12 and common.rel_err(V2, g["closed/Val"]) < 1e-12
    eng.set_ris([mp], [0])
    R3, V3 = run()
    assert common.rel_err(R3, g["open/R"]) < 1e-12 and common.rel_err(V3, g["open/Val"]) < 1e-12
    if scatter == abi.SCATTER_COLORED:
        assert np.array_equal(R3, R1) and np.array_equal(V3, V1)
    bad = mp.copy(); bad[1, 0] = -1
    with pytest.raises(RuntimeError):
        eng.set_ris([bad], [0])                      # a mapped node without a twin (the reference would reuse a stale row)
    eng.set_ris([], [])
    R4, _ = run()
    assert common.rel_err(R4, g["closed/R"]) < 1e-12
    eng.close()


# ---- Taylor-Hood function spaces (mshType::nFs = 2): construct_fluid with vmsStab = false ---------------------------------------
@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("case", common.FLUID_THOOD_CASES + [common.FLUID_THOOD_URIS_CASE],
                         ids=[c[0] for c in common.FLUID_THOOD_CASES] + ["tet10_uris"])
def test_taylor_hood_fluid_matches_golden(case, scatter):
    """P2-P1 tetrahedra (TET10 / TET4) and Q2-Q1 hexahedra (HEX27, HEX20 / HEX8): momentum loop on the velocity rule, continuity loop on
    the pressure rule, then fs::thood_val_rc — R / Val against the compiled reference (tests/golden/fluid_thood.npz), entry type by
    entry type; an equal-order assembly on the same mesh afterwards (svb200_set_mesh_thood(eNoNq = 0)) still matches fluid_hi.npz."""
    from svmultiphysics_b200.engine import Engine
    name, mk, visc, Kd, f, tDof, mv = case
    golden, tabs = common.load_golden("fluid_thood.npz"), common.load_golden("fluid_hi.npz")
    m = mk()
    et = name.split("_")[0]
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    w, N, Nx, Nxx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    t = {k: golden[f"tables/{et}/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
    eq, dmn = common.fluid_thood_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    eng = Engine(0)
    eng.set_graph(golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"])
    eng.set_mesh(0, m.IEN, w, N, Nx, Nxx=Nxx)
    eng.set_coords(m.x)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf)
    with pytest.raises(RuntimeError):
        eng.assemble(0, eq, dmn)                        # vmsStab = 0 without Taylor-Hood tables
    eng.set_mesh_thood(0, t)
    if name.endswith("uris"):                           # two URIS valves: the momentum loop sees the factor at the velocity rule's points
        raw, dev, sdf, udf, vel = common.uris_valves(m)
        eng.set_uris(dev, sdf, udf, vel)
    eng.alloc(4); eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    GR, GV = golden[f"{name}/R"], golden[f"{name}/Val"]
    assert common.rel_err(R1[:3], GR[:3]) < 1e-12 and common.rel_err(R1[3], GR[3]) < 1e-12
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14]):
        assert common.rel_err(V1[rows], GV[rows]) < 1e-12
    assert not V1[15].any()
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_R(), R1) and np.array_equal(eng.get_Val(), V1)
    eng.thood_val_rc()
    R2, V2 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R2[:3], golden[f"{name}/R_rc"][:3]) < 1e-12 and common.rel_err(R2[3], golden[f"{name}/R_rc"][3]) < 1e-12
    assert np.array_equal(V2[15], golden[f"{name}/Val_rc"][15])          # zeros and ones
    assert (V2[15] == 1.0).sum() == (R2[3] == 0.0).sum() > 0
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14]):
        assert np.array_equal(V2[rows], V1[rows])
    # back to equal-order VMS spaces on the same mesh object
    eng.set_mesh_thood(0, None)
    eng.set_uris([])
    hi = next((c for c in common.FLUID_HI_CASES if c[0].startswith(et) and c[5] == tDof and c[6] == mv and c[3] == Kd and c[4] == f
               and c[2] == visc), None)
    if hi is not None:
        eng.alloc(4); eng.assemble(0, abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter), dmn)
        assert common.rel_err(eng.get_R(), tabs[f"{hi[0]}/R"]) < 1e-12
    eng.close()


def test_taylor_hood_newton_iteration_matches_the_reference():
    """One Newton iteration on a P2-P1 mesh: assembly, fs::thood_val_rc, Dirichlet faces and GMRES on the saddle-point system (no
    pressure-pressure block besides the identity rows of the edge nodes) — iteration count, norms and increment against the compiled
    reference run live."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    from svmultiphysics_b200.engine import Engine
    golden, tabs = common.load_golden("fluid_thood.npz"), common.load_golden("fluid_hi.npz")
    m = meshgen.elevate(meshgen.box_tet4(3, 3, 3, (1.0, 1.0, 1.0)), "tet10", bend=0.0)      # straight edges: the box faces stay planes
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, 4)
    w, N, Nx, Nxx = (tabs[f"tables/tet10/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    t = {k: golden[f"tables/tet10/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
    lo, hi = m.x.min(axis=1), m.x.max(axis=1)
    wall = np.where((np.abs(m.x[0] - lo[0]) < 1e-9) | (np.abs(m.x[0] - hi[0]) < 1e-9) | (np.abs(m.x[1] - lo[1]) < 1e-9) |
                    (np.abs(m.x[1] - hi[1]) < 1e-9) | (np.abs(m.x[2] - lo[2]) < 1e-9))[0].astype(np.int32)
    faces = [(abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))]
    eq, dmn = common.fluid_thood_eq(0.005), [abi.fluid_domain(f=(0.0, 0.0, 1.0))]
    orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN); orc.set_mesh_thood(0)
    rowPtr, colPtr = orc.build_graph(len(faces))
    eng = Engine(0); eng.set_graph(rowPtr, colPtr)
    eng.set_mesh(0, m.IEN, w, N, Nx, Nxx=Nxx); eng.set_mesh_thood(0, t); eng.set_coords(m.x)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn); orc.thood_val_rc()
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn); eng.thood_val_rc()
    R0, R1 = orc.get_R(), eng.get_R()
    assert common.rel_err(R1[:3], R0[:3]) < 1e-12 and common.rel_err(R1[3], R0[3]) < 1e-12
    # ONE Krylov cycle of 60 vectors: restarted GMRES on this saddle-point system needs ~500-900 iterations, its count moves by tens of
    # iterations under last-bit differences (measured 517 in the reference) and classical Gram-Schmidt over 250 vectors separates the two
    # histories further.  Measured after 60 vectors: residual reduction 370x, final norms equal to 1.1e-4 — the classical Gram-Schmidt of
    # gmres.cpp (h(i+1,i) = sqrt|h(i+1,i) - sum h(j,i)^2|) amplifies last-bit differences on this indefinite system; R, Val and the
    # preconditioned initial norm (1e-10) are the parity statements, the solve shows the solver consumes the system
    ls = abi.ls_params(abi.LS_GMRES, mItr=1, sD=60, relTol=1e-6)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
    X1, o1, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res)
    assert o1.RI.success == o0.RI.success
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert o1.RI.itr == o0.RI.itr == 61            # 1 per cycle + 1 per Krylov vector (gmres.cpp:483, 513)
    assert abs(o1.RI.fNorm - o0.RI.fNorm) <= 1e-3 * o0.RI.fNorm, (o1.RI.fNorm, o0.RI.fNorm)
    assert o0.RI.fNorm < 1e-2 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-2, common.rel_err(X1, X0)
    eng.close()


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
def test_fsi_on_a_taylor_hood_mesh_matches_golden(scatter):
    """fsi::construct_fsi on a curved TET10 mesh with Taylor-Hood function spaces: the fluid core through the Taylor-Hood kernel on the
    moved geometry (ale), the struct wall through the general solid kernel on the velocity space, then thood_val_rc — against the
    compiled reference (tests/golden/fluid_thood.npz: fsi_tet10), rows compared per node class and Val per entry type."""
    from svmultiphysics_b200.engine import Engine
    golden, tabs = common.load_golden("fluid_thood.npz"), common.load_golden("fluid_hi.npz")
    m, Ag, Yg, Dg, Bf, eq, dmn = common.fsi_thood_case(scatter)
    w, N, Nx, Nxx = (tabs[f"tables/tet10/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    t = {k: golden[f"tables/tet10/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
    eng = Engine(0)
    rowPtr = golden["fsi_tet10/rowPtr"]
    eng.set_graph(rowPtr, golden["fsi_tet10/colPtr"])
    eng.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, Nxx=Nxx); eng.set_mesh_thood(0, t)
    eng.set_coords(m.x)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    GR, GV = golden["fsi_tet10/R"], golden["fsi_tet10/Val"]
    fluid_nodes = np.unique(m.IEN[:, (m.eId & 1) != 0]); solid_nodes = np.unique(m.IEN[:, (m.eId & 2) != 0])
    sets = (np.setdiff1d(fluid_nodes, solid_nodes), np.setdiff1d(solid_nodes, fluid_nodes), np.intersect1d(fluid_nodes, solid_nodes))
    assert all(len(x) > 0 for x in sets)
    for nodes in sets:
        assert common.rel_err(R1[:3, nodes], GR[:3, nodes]) < 1e-12
        if np.abs(GR[3, nodes]).max() > 0:
            assert common.rel_err(R1[3, nodes], GR[3, nodes]) < 1e-12
        slots = np.concatenate([np.arange(rowPtr[a], rowPtr[a + 1]) for a in nodes])
        for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14]):
            if np.abs(GV[rows][:, slots]).max() > 0:
                assert common.rel_err(V1[rows][:, slots], GV[rows][:, slots]) < 1e-12
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(4); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_R(), R1) and np.array_equal(eng.get_Val(), V1)
    eng.thood_val_rc()
    assert np.array_equal(eng.get_Val()[15], golden["fsi_tet10/Val_rc"][15])
    assert np.array_equal(eng.get_R()[3] == 0.0, golden["fsi_tet10/R_rc"][3] == 0.0)
    eng.close()
