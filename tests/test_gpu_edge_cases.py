"""GPU parity on degenerate sizes: a single element, element counts that do not fill the 128-element groups / 32-lane warps of the
kernels, a mesh none of whose elements belongs to the equation, an empty mesh, and a Krylov solve on a 4-node system — against the
oracle (the compiled reference when oracle/_ref/libsvref.so is present, its C restatement otherwise) on the same input."""
import copy

import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements, meshgen
from tests import common

pytestmark = pytest.mark.gpu
ASM_TOL = 1e-12


def _oracle():
    from oracle import refbind
    return refbind.RefCase if refbind.have_ref() else refbind.OracleCase


def _ref_only():
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    return refbind.RefCase


def _submesh(m, elems):
    """The elements `elems` of m with compact node numbering."""
    I = m.IEN[:, elems]
    nodes, inv = np.unique(I, return_inverse=True)
    s = copy.copy(m)
    s.x = np.asfortranarray(m.x[:, nodes])
    s.IEN = np.asfortranarray(inv.reshape(I.shape).astype(np.int32))
    s.faces = {}
    s.eId = None if m.eId is None else m.eId[elems]
    return s


def _fluid_state(m, seed=9):
    rng = np.random.default_rng(seed)
    Yg = np.asfortranarray(rng.standard_normal((4, m.nNo)))
    Ag = np.asfortranarray(rng.standard_normal((4, m.nNo)))
    return Ag, Yg


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("nEl", [1, 2, 127, 128, 129, 257])
def test_fluid_tet4_with_few_elements(nEl, scatter):
    """1 element (4 nodes, a 4 x 4 block matrix) up to two groups and one element: the grouped TET4 kernel, its plan and the
    colour-by-group mode on partial groups."""
    big = meshgen.cylinder_tet4(4, 3)
    assert big.nEl >= 257
    m = _submesh(big, np.arange(nEl))
    Ag, Yg = _fluid_state(m)
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m)
    eq, dmn = abi.fluid_eq(0.005, scatter=scatter), [abi.fluid_domain(K_darcy=0.3, f=(0.1, 0.2, 0.3))]
    orc.alloc(4); orc.set_state(Ag, Yg, None, None); orc.assemble(0, eq, dmn)
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.alloc(4); eng.set_state(Ag, Yg, None, None); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), orc.get_R()) < ASM_TOL
    assert common.rel_err(eng.get_Val(), orc.get_Val()) < ASM_TOL
    eng.close()


@pytest.mark.parametrize("kind", ["struct_hex8", "struct_tet4", "heats_hex8", "lelas_hex8", "fluid_hex8"])
@pytest.mark.parametrize("nEl", [1, 3, 5])
def test_lane_group_kernels_with_few_elements(kind, nEl):
    """Kernels that pack 32 / eNoN elements into a warp, with fewer elements than one warp holds (idle lane groups)."""
    cls = _ref_only()
    big = meshgen.box_hex8(3, 2, 1, (1.0, 1.1, 0.9)) if kind.endswith("hex8") else meshgen.box_tet4(1, 1, 1, (1.0, 1.0, 1.0))
    rng = np.random.default_rng(3)
    big.x = np.asfortranarray(big.x + 0.03 * rng.standard_normal(big.x.shape))
    m = _submesh(big, np.arange(nEl))
    if kind.startswith("struct"):
        Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
        dof, eq, dmn = 3, abi.struct_eq(1e-4), [abi.struct_domain(E=1e6, nu=0.4, Kpen=1e6, rho=1.0, volType=abi.VOL_QUAD)]
    elif kind.startswith("heats"):
        Ag, Yg, Dg, Bf = common.heat_state(m, 1, 0)
        dof, eq, dmn = 1, abi.heat_eq(0.01, False), [abi.heat_domain(False, conductivity=0.7, source=0.3, rho=2.5)]
    elif kind.startswith("lelas"):
        Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
        dof, eq, dmn = 3, abi.lelas_eq(1e-3), [abi.lelas_domain(E=1.0e6, nu=0.3, rho=2.0, f=(0.1, -0.2, 0.3))]
    else:
        Ag, Yg, Dg, Bf = common.fluid_gen_state(m, 4)
        dof, eq, dmn = 4, abi.fluid_eq(0.005), [abi.fluid_domain()]
    orc, rowPtr, colPtr = common.make_oracle(cls, m)
    orc.alloc(dof); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.alloc(dof); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    assert common.rel_err(eng.get_R(), orc.get_R()) < ASM_TOL
    assert common.rel_err(eng.get_Val(), orc.get_Val()) < ASM_TOL
    eng.close()


def test_mesh_without_an_element_of_the_equation_and_empty_mesh():
    """(i) every element of the mesh belongs to a domain of another physics (a solid mesh handed to a fluid-only assembly of an FSI
    equation): nothing is added; (ii) a mesh with zero elements next to a normal one: a no-op, like an empty element loop."""
    from svmultiphysics_b200.engine import Engine
    m = meshgen.cylinder_tet4(3, 3)
    m.eId = np.full(m.nEl, 2, np.int32)                       # all elements in domain Id 1
    Ag, Yg = _fluid_state(m)
    eng = Engine(0)
    rowPtr, colPtr = eng.lhsa(m.nNo, [m.IEN]); eng.set_graph(rowPtr, colPtr)
    w, N, Nx = elements.tables(4)
    eng.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId)
    eng.set_mesh(1, np.zeros((4, 0), np.int32, order="F"), w, N, Nx)
    eng.set_coords(m.x)
    eq = abi.fluid_eq(0.005)
    dmn = [abi.fluid_domain(Id=0), abi.struct_domain(Id=1)]   # the fluid domain covers no element of this mesh
    eng.alloc(4); eng.set_state(Ag, Yg, None, None)
    eng.assemble(0, eq, dmn)
    eng.assemble(1, eq, [abi.fluid_domain()])
    assert not eng.get_R().any() and not eng.get_Val().any()
    eng.close()


@pytest.mark.parametrize("ls_type", [abi.LS_GMRES, abi.LS_BICGS])
def test_krylov_solve_on_a_single_element(ls_type):
    """fsils_solve on the 4-node system of one tetrahedron (one Dirichlet node): the solvers' block sizes, reductions and Krylov
    dimension exceed the problem size."""
    big = meshgen.cylinder_tet4(2, 2)
    m = _submesh(big, np.arange(1))
    Ag, Yg = _fluid_state(m)
    faces = [(abi.BC_DIR, np.array([0], np.int32), np.zeros((3, 1), order="F"))]
    orc, rowPtr, colPtr = common.make_oracle(_oracle(), m, nFaces=1)
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(1)
    for i, (g, nodes, val) in enumerate(faces):
        orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    orc.alloc(4); orc.set_state(Ag, Yg, None, None); orc.assemble(0, eq, dmn)
    eng.alloc(4); eng.set_state(Ag, Yg, None, None); eng.assemble(0, eq, dmn)
    ls = abi.ls_params(ls_type, mItr=3 if ls_type == abi.LS_GMRES else 60, sD=20, relTol=1e-10)
    incL, res = np.ones(1, np.int32), np.zeros(1)
    X0, o0, _ = orc.solve(4, ls_type, ls, incL, res)
    X1, o1, _ = eng.solve(4, ls_type, ls, incL, res)
    assert o1.RI.success == o0.RI.success
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-6
    eng.close()
