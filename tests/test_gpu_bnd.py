"""GPU parity of the boundary-face (Neumann / backflow) assembly against the compiled reference's nn::gnnb +
fluid::b_fluid / l_elas::b_l_elas driven by the loop of eq_assem::b_assem_neu_bc."""
import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements, meshgen
from tests import common

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("boundary-face parity needs oracle/_ref/libsvref.so")
    return refbind.RefCase


@pytest.mark.parametrize("mv", [0, 1], ids=["fixed_mesh", "moving_mesh"])
def test_fluid_neumann_backflow_parity(mv):
    tDof = 7 if mv else 4
    m, Ag, Yg, Dg, Bf = common.fluid_case(tDof=tDof)
    rng = np.random.default_rng(8)
    Yg[2] *= np.where(rng.random(m.nNo) < 0.5, -1.0, 1.0)      # mixed in/outflow so that backflow terms fire
    Do = np.zeros((tDof, m.nNo), order="F")
    if mv:
        Do[4:7] = 5e-3 * rng.standard_normal((3, m.nNo))
    orc, rowPtr, colPtr = common.make_oracle(_ref(), m)
    eng = common.make_engine(m, rowPtr, colPtr)
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv)
    dmn = [abi.fluid_domain(backflow_stab=0.2)]
    orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.set_old_disp(Do) if mv else None
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.set_old_disp(Do) if mv else None
    for k, name in enumerate(("outlet_all", "inlet")):
        IENb, gE = meshgen.boundary_face_elements(m, m.faces[name])
        assert IENb.shape[1] > 0
        iFa = orc.add_face(0, IENb, gE)
        w, N, Nx = orc.face_tables(0, iFa)
        w2, N2, Nx2 = elements.face_tables(3)
        assert np.allclose(w, w2, rtol=0, atol=1e-15) and np.allclose(N, N2, rtol=0, atol=1e-15) and np.allclose(Nx, Nx2, rtol=0, atol=1e-15)
        eng.set_bface(iFa, 0, IENb, gE, w2, N2, Nx2)
        hg = np.zeros(m.nNo)
        hg[m.faces[name]] = -(100.0 + 10.0 * k) * (1.0 + 0.1 * rng.standard_normal(len(m.faces[name])))
        orc.assemble_neu(0, iFa, eq, dmn, hg)
        eng.assemble_neu(iFa, eq, dmn, hg)
    R0, V0 = orc.get_R(), orc.get_Val()
    assert np.abs(R0).max() > 0 and np.abs(V0).max() > 0
    assert common.rel_err(eng.get_R(), R0) < 1e-12
    assert common.rel_err(eng.get_Val(), V0) < 1e-12
    eng.close()


def test_struct_traction_face_parity():
    """Pressure load on the Z1 face of the hex8 block (struct/block_compression): b_l_elas through QUD4 faces."""
    m = common.STRUCT_CASES[0][1]()
    Ag, Yg, Dg, Bf, _ = common.struct_state(m)
    cls = _ref()
    orc = cls(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
    rowPtr, colPtr = orc.build_graph(0)
    from svmultiphysics_b200.engine import Engine
    eng = Engine(0); eng.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(8); eng.set_mesh(0, m.IEN, w, N, Nx); eng.set_coords(m.x)
    eq, dmn = abi.struct_eq(1e-4), [abi.struct_domain()]
    orc.alloc(3); orc.set_state(Ag, Yg, Dg, Bf)
    eng.alloc(3); eng.set_state(Ag, Yg, Dg, Bf)
    IENb, gE = meshgen.boundary_face_elements(m, m.faces["Z1"])
    iFa = orc.add_face(0, IENb, gE)
    wf, Nf, Nxf = orc.face_tables(0, iFa)
    w2, N2, Nx2 = elements.face_tables(4)
    assert np.allclose(wf, w2, atol=1e-15) and np.allclose(Nf, N2, atol=1e-15) and np.allclose(Nxf, Nx2, atol=1e-15)
    eng.set_bface(iFa, 0, IENb, gE, w2, N2, Nx2)
    hg = np.zeros(m.nNo); hg[m.faces["Z1"]] = -5.0e4
    orc.assemble_neu(0, iFa, eq, dmn, hg)
    eng.assemble_neu(iFa, eq, dmn, hg)
    R0 = orc.get_R()
    assert np.abs(R0).max() > 0
    assert common.rel_err(eng.get_R(), R0) < 1e-12
    eng.close()
