"""ncu driver: TET10 fluid through the general kernel, one warm-up + one launch."""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
tabs = common.load_golden("fluid_hi.npz")
m = meshgen.elevate(meshgen.box_tet4(n, n, n), "tet10", bend=0.02)
w, N, Nx, Nxx = (tabs[f"tables/tet10/{k}"] for k in ("w", "N", "Nx", "Nxx"))
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
e.set_mesh(0, m.IEN, w, N, Nx, Nxx=Nxx); e.set_coords(m.x)
A, Y, D, B = common.fluid_gen_state(m, 4)
e.alloc(4); e.set_state(A, Y, D, B)
for _ in range(2):
    e.alloc(4); e.assemble(0, abi.fluid_eq(0.005), [abi.fluid_domain()])
print("done", m.nEl)
