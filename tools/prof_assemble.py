"""Small driver for ncu: a few assembly launches + SpMV + a short GMRES on a configurable cylinder."""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 120
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
do_solve = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m = meshgen.cylinder_tet4(n, nz)
Ag, Yg, Dg = meshgen.poiseuille_state(m)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(4); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
e.alloc(4); e.set_state(Ag, Yg, Dg)
eq = abi.fluid_eq(1e-3); dm = [abi.fluid_domain()]
for _ in range(reps):
    e.alloc(4); e.assemble(0, eq, dm)
e.bench_spmv(4, 2)
if do_solve:
    wall = m.faces["wall"]; e.set_num_faces(1); e.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order='F'))
    ls = abi.ls_params(abi.LS_GMRES, mItr=1, sD=12, relTol=1e-12)
    e.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1), want_solution=False)
print("done", m.nEl)
