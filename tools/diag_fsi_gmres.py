"""Diagnostic (1 GPU): GMRES residual history of the FSI equation (config C5) on the device vs the reference / the
C restatement fed with the reference's assembled system."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svmultiphysics_b200 import abi, elements
from svmultiphysics_b200.engine import Engine
from oracle import refbind
from tests import common

m, Ag, Yg, Dg, Bf = common.fsi_case()
wall = m.faces["wall"]
af, am, gam, beta = abi.gen_alpha(0.5)
eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                  scatter=abi.SCATTER_ATOMIC, reserved=0)
dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0),
       abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
ls = abi.ls_params(abi.LS_GMRES, mItr=100, sD=50, relTol=1e-8)
incL, res = np.ones(1, np.int32), np.zeros(1)
val = np.zeros((3, len(wall)), order="F")

orc = refbind.RefCase(); orc.set_coords(m.x); orc.add_mesh(m.IEN, eId=m.eId)
rowPtr, colPtr = orc.build_graph(1)
orc.set_face(0, abi.BC_DIR, wall, val)
orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
R0, V0 = orc.get_R(), orc.get_Val()
X0, o0, _ = orc.solve(4, abi.LS_GMRES, ls, incL, res)
print("reference: itr", o0.RI.itr, "iNorm", o0.RI.iNorm, "fNorm", o0.RI.fNorm)

oc = refbind.OracleCase(); oc.set_coords(m.x); oc.add_mesh(m.IEN, eId=m.eId); oc.build_graph(1)
oc.set_face(0, abi.BC_DIR, wall, val)
oc.alloc(4); oc.put_Val(V0, 4); oc.put_R(R0)
Xc, occ, hc = oc.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=512)
print("restatement on the reference's system: itr", occ.RI.itr, "fNorm", occ.RI.fNorm)

eng = Engine(0)
eng.set_graph(rowPtr, colPtr)
w, N, Nx = elements.tables(4)
eng.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId)
eng.set_coords(m.x)
eng.set_num_faces(1); eng.set_face(0, abi.BC_DIR, wall, val)
for mode in ("own", "uploaded"):
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf)
    if mode == "own":
        eng.assemble(0, eq, dmn)
        print("assembly relerr R", common.rel_err(eng.get_R(), R0), "Val", common.rel_err(eng.get_Val(), V0))
    else:
        eng.put_Val(V0, 4); eng.put_R(R0)
    X1, o1, h1 = eng.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=512)
    print(mode, ": itr", o1.RI.itr, "fNorm", o1.RI.fNorm, "relerr X", common.rel_err(X1, X0))
    n = min(len(h1), len(hc))
    for i in list(range(0, n, 5)) + [n - 1]:
        print(f"   it {i:3d}  gpu {h1[i]:.6e}  cpu {hc[i]:.6e}  rel {abs(h1[i]-hc[i])/hc[i]:.1e}")
print("tolerance:", ls.RI.relTol * o0.RI.iNorm)
