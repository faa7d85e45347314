import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
m = meshgen.box_hex8(n, n, n, (1.0, 1.0, 1.0))
m.x = np.asfortranarray(m.x + (0.1 / n) * np.random.default_rng(17).standard_normal(m.x.shape))
Ag, Yg, Dg, Bf, _ = common.ustruct_state(m)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(8)
e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
e.alloc(4); e.set_state(Ag, Yg, Dg, Bf)
eq, dm = abi.ustruct_eq(1e-3), [abi.ustruct_domain(E=1.0e6, nu=0.45, Kpen=1.0e6 / (3 * (1 - 0.9)), rho=1.2)]
for _ in range(2):
    e.alloc(4); e.assemble(0, eq, dm)
print("done", m.nEl)
