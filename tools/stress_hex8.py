import os, sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi
from tests import common
golden = common.load_golden("fluid_gen.npz")
name, mk, visc, Kd, f, tDof, mv = common.FLUID_GEN_CASES[0]
m = mk()
Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("X0", "Y0", "Y1", "Z0", "Z1")]
rowPtr, colPtr = golden[f"{name}/rowPtr"], golden[f"{name}/colPtr"]
V0, R0 = golden[f"{name}/Val"], golden[f"{name}/R"]
scatter = abi.SCATTER_COLORED if len(sys.argv) > 2 and sys.argv[2] == "colored" else abi.SCATTER_ATOMIC
eq, dmn = abi.fluid_eq(0.005, scatter=scatter), [abi.fluid_domain()]
h_first = None
ls = abi.ls_params(abi.LS_GMRES, mItr=10, sD=100, relTol=1e-8)
incL, res = np.ones(len(faces), np.int32), np.zeros(len(faces))
counts = {}
bad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 100):
    eng = common.make_engine(m, rowPtr, colPtr)
    eng.set_num_faces(len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg, Bf); eng.assemble(0, eq, dmn)
    eR = common.rel_err(eng.get_R(), R0)
    eV = common.rel_err(eng.get_Val(), V0) if rep % 2 == 0 else -1.0
    X1, o1, h1 = eng.solve(4, abi.LS_GMRES, ls, incL, res, hist_cap=256)
    W = eng.get_W()
    counts[o1.RI.itr] = counts.get(o1.RI.itr, 0) + 1
    if h_first is None:
        h_first = h1.copy()
    if o1.RI.itr > 30 or eR > 1e-12 or eV > 1e-12:
        bad += 1
        n = min(len(h1), len(h_first))
        d = np.abs(h1[:n] - h_first[:n]) / h_first[:n]
        print("   drift vs first run at it 5,10,15,20,23:", [f"{d[k]:.1e}" for k in (5, 10, 15, 20, 23) if k < n], "hist[20:30]", h1[20:30])
        print(f"rep {rep}: itr {o1.RI.itr} iNorm {o1.RI.iNorm!r} eR {eR:.2e} eV {eV:.2e} W zeros {(W == 0).sum()} hist {h1[:6]}")
    eng.close()
print("iteration histogram:", counts, "bad", bad)
