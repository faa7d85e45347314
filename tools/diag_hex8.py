import os, sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi
from oracle import refbind
from tests import common
name, mk, visc, Kd, f, tDof, mv = common.FLUID_GEN_CASES[0]
m = mk()
Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("X0", "Y0", "Y1", "Z0", "Z1")]
orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN)
rowPtr, colPtr = orc.build_graph(len(faces))
eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
eng = common.make_engine(m, rowPtr, colPtr)
eng.set_num_faces(len(faces))
for i, (g, nodes, val) in enumerate(faces):
    orc.set_face(i, g, nodes, val); eng.set_face(i, g, nodes, val)
ls = abi.ls_params(abi.LS_GMRES, mItr=10, sD=100, relTol=1e-8)
incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
R0, V0 = orc.get_R(), orc.get_Val()
X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
Vs0 = orc.get_Val()
print("ref itr", o0.RI.itr, o0.RI.iNorm, o0.RI.fNorm)
oc = refbind.OracleCase(); oc.set_coords(m.x); oc.add_mesh(m.IEN); oc.build_graph(len(faces))
for i, (g, nodes, val) in enumerate(faces): oc.set_face(i, g, nodes, val)
oc.alloc(4); oc.put_Val(V0, 4); oc.put_R(R0)
Xc, occ, hc = oc.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=512)
print("restatement itr", occ.RI.itr, occ.RI.iNorm, occ.RI.fNorm)
for mode in ("own", "uploaded"):
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf)
    if mode == "own":
        eng.assemble(0, eq, dmn)
        print("asm relerr", common.rel_err(eng.get_R(), R0), common.rel_err(eng.get_Val(), V0))
    else:
        eng.put_Val(V0, 4); eng.put_R(R0)
    X1, o1, h1 = eng.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=512)
    print(mode, "itr", o1.RI.itr, o1.RI.iNorm, o1.RI.fNorm, "X relerr", common.rel_err(X1, X0), "scaled Val relerr", common.rel_err(eng.get_Val(), Vs0))
    n = min(len(h1), len(hc))
    for i in list(range(0, n, 4)):
        print(f"   it {i:3d}  gpu {h1[i]:.6e}  cpu {hc[i]:.6e}")
