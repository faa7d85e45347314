"""Small driver for ncu: HEX8 solid (struct_3d) and HEX8 fluid (general kernel) assembly, one warm-up + one launch each."""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
m = meshgen.box_hex8(n, n, n, (1e-3, 1e-3, 1e-3))
w, N, Nx = elements.tables(8)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
e.set_mesh(0, m.IEN, w, N, Nx, Nxx=elements.nxx_tables(8)); e.set_coords(m.x)
Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
Dg *= 0.1
e.alloc(3); e.set_state(Ag, Yg, Dg, Bf)
for _ in range(2):
    e.alloc(3); e.assemble(0, abi.struct_eq(1e-4), [abi.struct_domain()])
e.bench_spmv(3, 2)
Ag, Yg, Dg, Bf = common.fluid_gen_state(m, 4)
e.alloc(4); e.set_state(Ag, Yg, None, Bf)
for _ in range(2):
    e.alloc(4); e.assemble(0, abi.fluid_eq(1e-3), [abi.fluid_domain()])
print("done", m.nEl)
