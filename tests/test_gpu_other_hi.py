"""GPU parity of the scalar heat equations (heats_3d / heatf_3d) and of l_elas_3d (linear-elasticity equation and mesh-motion equation)
on curved TET10 (15 Gauss points), HEX20 / HEX27 (27) and WDG (6, the reference's lShpF behaviour: one gnn per element) elements, against
the golden vectors of the compiled reference (tests/golden/other_hi.npz; element tables tests/golden/fluid_hi.npz)."""
import numpy as np
import pytest

from svmultiphysics_b200 import abi
from tests import common

pytestmark = pytest.mark.gpu
ASM_TOL = 1e-12


@pytest.mark.parametrize("scatter", [abi.SCATTER_ATOMIC, abi.SCATTER_COLORED])
@pytest.mark.parametrize("name", [c[0] for c in common.OTHER_HI_CASES])
def test_heat_and_linear_elasticity_on_quadratic_and_wedge_elements(name, scatter):
    from svmultiphysics_b200.engine import Engine
    golden, tabs = common.load_golden("other_hi.npz"), common.load_golden("fluid_hi.npz")
    m, et, dof, Ag, Yg, Dg, Bf, Do, eq, dmn = common.other_hi_case(name, scatter)
    w, N, Nx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx"))
    eng = Engine(0)
    eng.set_graph(golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"])
    eng.set_mesh(0, m.IEN, w, N, Nx)
    eng.set_coords(m.x)
    eng.alloc(dof); eng.set_state(Ag, Yg, Dg, Bf)
    if Do is not None:
        eng.set_old_disp(Do)
    eng.assemble(0, eq, dmn)
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, golden[f"{name}/R"]) < ASM_TOL
    assert common.rel_err(V1, golden[f"{name}/Val"]) < ASM_TOL
    if scatter == abi.SCATTER_COLORED:
        eng.alloc(dof); eng.assemble(0, eq, dmn)
        assert np.array_equal(eng.get_R(), R1) and np.array_equal(eng.get_Val(), V1)
    eng.close()
