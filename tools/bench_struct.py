"""C4 of SURVEY.md §8(d): hex8 neo-Hookean (ST91) block, n^3 elements — struct_3d assembly time, dof-3 SpMV and a
short BiCGStab solve on one B200.  Usage: python tools/bench_struct.py [n=171] [reps=3] [iso=nhk|guccione]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 171
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
iso = sys.argv[3] if len(sys.argv) > 3 else "nhk"
t0 = time.time()
m = meshgen.box_hex8(n, n, n, (1e-3, 1e-3, 1e-3))
nFn = 2 if iso == "guccione" else 0
Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
Dg *= 0.1   # keep the 1e-3-sized noise of struct_state well inside det F > 0 on the fine mesh
dkw = dict(isoType=abi.ISO_GUCCIONE, C10=440.0, bff=8.0, bss=6.0, bfs=12.0, Kpen=1e6, rho=1e-3) if nFn else {}
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(8); e.set_mesh(0, m.IEN, w, N, Nx, nFn=nFn, fN=fN); e.set_coords(m.x)
e.alloc(3); e.set_state(Ag, Yg, Dg, Bf)
eq, dm = abi.struct_eq(1e-4), [abi.struct_domain(**dkw)]
print(f"setup {time.time()-t0:.1f} s: {m.nEl} hex8, {m.nNo} nodes, nnz {len(cp)} blocks ({len(cp)*72/1e9:.2f} GB Val)")
e.alloc(3); e.assemble(0, eq, dm)
for _ in range(reps):
    e.alloc(3)
    e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1)
    ms = e.timer_elapsed()
    print(f"struct assemble ({iso}) {ms:.3f} ms  {m.nEl/ms*1e-6:.4f} G el/s  ({m.nEl*130e3/ms*1e-9:.2f} TFLOP/s by the 130 kflop/hex8 model)")
ms = e.bench_spmv(3, 10)
print(f"SpMV dof=3 {ms:.3f} ms  {(len(cp)*76 + m.nNo*56)/ms*1e-6:.1f} GB/s algorithmic (nnz*76 + nNo*56)")
faces = []
for k, name in enumerate(("X0", "Y0", "Z0")):
    val = np.ones((3, len(m.faces[name])), order="F"); val[k] = 0.0
    faces.append((abi.BC_DIR, m.faces[name], val))
e.set_num_faces(len(faces))
for i, (g, nodes, val) in enumerate(faces):
    e.set_face(i, g, nodes, val)
ls = abi.ls_params(abi.LS_BICGS, mItr=50, relTol=1e-12)
e.timer_mark(0)
_, out, _ = e.solve(3, abi.LS_BICGS, ls, np.ones(3, np.int32), np.zeros(3), want_solution=False)
e.timer_mark(1)
print(f"BiCGStab {out.RI.itr} its in {e.timer_elapsed():.2f} ms  iNorm {out.RI.iNorm:.3e} fNorm {out.RI.fNorm:.3e}")
