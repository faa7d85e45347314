"""GPU parity of the device-side generalised-alpha updates against oracle/genalpha_oracle.py (bit-exact: the
kernels use separate roundings in the reference's evaluation order), and a device-resident Newton loop."""
import numpy as np
import pytest

from svmultiphysics_b200 import abi
from tests import common

pytestmark = pytest.mark.gpu


def _states(tDof, nNo, seed=5):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray(rng.standard_normal((tDof, nNo))) for _ in range(6)]


def test_genalpha_kernels_bit_exact():
    from oracle import genalpha_oracle as go
    m, *_ = common.fsi_case()
    orc, rowPtr, colPtr = common.make_oracle(_any_oracle(), m)
    eng = common.make_engine(m, rowPtr, colPtr)
    tDof, dt = 7, 0.0125
    eqs = [abi.eq_time(0, 3, abi.PHYS_FSI, 0.5), abi.eq_time(4, 6, abi.PHYS_MESH, 0.2)]
    Ao, Yo, Do, An, Yn, Dn = _states(tDof, m.nNo)
    eng.set_solution(abi.SOL_OLD, Ao, Yo, Do)
    eng.set_solution(abi.SOL_CURRENT, An, Yn, Dn)
    for dFlag in (1, 0):
        eng.predictor(eqs, dt, dFlag)
        go.predictor(eqs, dt, dFlag, Ao, Yo, Do, An, Yn, Dn)
        for got, want in zip(eng.get_solution(abi.SOL_CURRENT), (An, Yn, Dn)):
            assert np.array_equal(got, want)
    # strong Dirichlet rows (set_bc_dir): velocity rows of the wall nodes
    wall = m.faces["wall"]
    vA = np.asfortranarray(np.random.default_rng(1).standard_normal((3, len(wall))))
    vY = np.asfortranarray(np.random.default_rng(2).standard_normal((3, len(wall))))
    eng.set_dirichlet_rows(0, wall, valA=vA, valY=vY)
    An[0:3, wall] = vA; Yn[0:3, wall] = vY
    Ag, Yg, Dg = (np.zeros_like(An) for _ in range(3))
    eng.initiator(eqs)
    go.initiator(eqs, Ao, Yo, Do, An, Yn, Dn, Ag, Yg, Dg)
    for got, want in zip(eng.get_solution(abi.SOL_INTERMEDIATE), (Ag, Yg, Dg)):
        assert np.array_equal(got, want)
    # corrector with the FSI copy on solid nodes; the increment is uploaded as R of a dof-4 system
    R = np.asfortranarray(np.random.default_rng(3).standard_normal((4, m.nNo)))
    solid = np.zeros(m.nNo, np.int32)
    solid[np.unique(m.IEN[:, (m.eId & 2) != 0])] = 1
    eng.alloc(4); eng.put_R(R); eng.set_node_flags(solid)
    eng.corrector(eqs[0], dt, mesh_s=4)
    go.corrector(eqs[0], dt, R, An, Yn, Dn, mesh_s=4, solid=solid)
    for got, want in zip(eng.get_solution(abi.SOL_CURRENT), (An, Yn, Dn)):
        assert np.array_equal(got, want)
    eng.advance_time_step()
    for got, want in zip(eng.get_solution(abi.SOL_OLD), (An, Yn, Dn)):
        assert np.array_equal(got, want)
    eng.close()


def _any_oracle():
    from oracle import refbind
    return refbind.RefCase if refbind.have_ref() else refbind.OracleCase


def test_device_resident_newton_loop_matches_host_loop():
    """Two time steps x three Newton iterations of the fluid equation: (a) everything on the device (predictor,
    initiator, assemble, solve, corrector; no nodal array crosses PCIe inside the loop) against (b) the same loop with
    the oracle's assembly + GMRES and the numpy gen-alpha restatement."""
    from oracle import genalpha_oracle as go
    cls = _any_oracle()
    from svmultiphysics_b200 import meshgen
    m = meshgen.cylinder_tet4(5, 6)
    _, Y0, _ = meshgen.poiseuille_state(m, noise=0.0, dpdz=-0.4)      # Poiseuille flow with its own pressure gradient
    Y0[0:3] += 0.05 * np.random.default_rng(3).standard_normal((3, m.nNo))   # + a perturbation for Newton to remove
    Bf = None
    faces = common.dirichlet_faces(m)
    orc, rowPtr, colPtr = common.make_oracle(cls, m, nFaces=len(faces))
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    dt = 0.01
    eq, dmn = abi.fluid_eq(dt), [abi.fluid_domain()]
    qt = [abi.eq_time(0, 3, abi.PHYS_FLUID, 0.5)]
    ls = abi.ls_params(abi.LS_GMRES, mItr=5, sD=120, relTol=1e-8)
    incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
    Ao, Yo, Do = np.zeros_like(Y0), Y0.copy(order="F"), np.zeros_like(Y0)
    An, Yn, Dn = Ao.copy(order="F"), Yo.copy(order="F"), Do.copy(order="F")
    Ag, Yg, Dg = (np.zeros_like(Ao) for _ in range(3))
    eng.set_state(Ag, Yg, Dg, Bf)
    eng.set_solution(abi.SOL_OLD, Ao, Yo, Do)
    eng.set_solution(abi.SOL_CURRENT, An, Yn, Dn)
    norms_dev, norms_ref = [], []
    for step in range(2):
        eng.predictor(qt, dt, 0)
        go.predictor(qt, dt, 0, Ao, Yo, Do, An, Yn, Dn)
        for it in range(3):
            eng.initiator(qt)
            eng.alloc(4); eng.assemble(0, eq, dmn)
            _, out1, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res, want_solution=False)
            eng.corrector(qt[0], dt)
            go.initiator(qt, Ao, Yo, Do, An, Yn, Dn, Ag, Yg, Dg)
            orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
            X0, out0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
            go.corrector(qt[0], dt, X0, An, Yn, Dn)
            norms_dev.append(out1.RI.iNorm); norms_ref.append(out0.RI.iNorm)
        eng.advance_time_step()
        Ao, Yo, Do = An.copy(order="F"), Yn.copy(order="F"), Dn.copy(order="F")
    # Newton convergence history (the preconditioned residual norm that drives Integrator::corrector's test)
    # the k-th norm of a step carries the linear-solver error (relTol 1e-8) of the previous iterations relative to its own,
    # shrinking, size: 1e-10, ~2e-6 and ~3e-4 here
    nd, nr = np.array(norms_dev).reshape(2, 3), np.array(norms_ref).reshape(2, 3)
    assert np.allclose(nd[:, 0], nr[:, 0], rtol=1e-8) and np.allclose(nd[:, 1], nr[:, 1], rtol=1e-4) and np.allclose(nd[:, 2], nr[:, 2], rtol=2e-2)
    assert norms_ref[2] < 0.5 * norms_ref[0]           # the Newton iteration contracts
    A1, Y1, D1 = eng.get_solution(abi.SOL_CURRENT)
    assert common.rel_err(Y1, Yn) < 1e-6 and common.rel_err(A1, An) < 1e-6
    eng.close()
