"""NS (Schur complement) solver on C2 (10 M tet4): the reference's default linear solver for fluid cases, with the
parameters of tests/cases/fluid/pipe_RCR_3d/solver.xml.  Usage: python tools/bench_ns.py [n=118] [nz=120]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 120
m = meshgen.cylinder_tet4(n, nz)
Ag, Yg, Dg = meshgen.poiseuille_state(m)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(4); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("wall", "inlet")]
out = m.faces["outlet"]
val = np.zeros((3, len(out)), order="F"); val[2] = 4.0 * np.pi / len(out)
faces.append((abi.BC_NEU, out, val))
e.set_num_faces(len(faces))
for i, (g, nodes, v) in enumerate(faces):
    e.set_face(i, g, nodes, v)
eq, dm = abi.fluid_eq(1e-3), [abi.fluid_domain()]
ls = abi.ls_params(abi.LS_NS, mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17))
incL, res = np.ones(3, np.int32), np.array([0.0, 0.0, 0.8])
for rep in range(3):
    e.alloc(4); e.set_state(Ag, Yg, Dg); e.assemble(0, eq, dm)
    e.timer_mark(0)
    _, o, _ = e.solve(4, abi.LS_NS, ls, incL, res, want_solution=False)
    e.timer_mark(1)
    print(f"NS solve {e.timer_elapsed():.1f} ms: outer {o.RI.itr} (success {o.RI.success}, {o.RI.fNorm/o.RI.iNorm:.2e}), "
          f"GMRES {o.GM.itr} its, CG {o.CG.itr} its, Resm {o.Resm} Resc {o.Resc}")
