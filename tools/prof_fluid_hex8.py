"""ncu driver: HEX8 fluid through the general kernel, one warm-up + one launch (n^3 skewed hexahedra)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
m = meshgen.box_hex8(n, n, n, (1.0, 1.0, 1.0))
rng = np.random.default_rng(17)
m.x = np.asfortranarray(m.x + (0.1 / n) * rng.standard_normal(m.x.shape))
Ag, Yg, Dg, Bf = common.fluid_gen_state(m, 4)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(8)
e.set_mesh(0, m.IEN, w, N, Nx, Nxx=elements.nxx_tables(8)); e.set_coords(m.x)
e.alloc(4); e.set_state(Ag, Yg, None, Bf)
for _ in range(2):
    e.alloc(4); e.assemble(0, abi.fluid_eq(1e-3), [abi.fluid_domain()])
print("done", m.nEl)
