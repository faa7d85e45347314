"""Quick performance probe on a B200 (not a test, not the bench): sizes via argv."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 120
t = time.time(); m = meshgen.cylinder_tet4(n, nz); print("mesh", m.nEl, m.nNo, "%.1fs" % (time.time()-t), flush=True)
Ag, Yg, Dg = meshgen.poiseuille_state(m)
e = Engine(0)
t = time.time(); rp, cp = e.lhsa(m.nNo, [m.IEN]); print("lhsa nnz", len(cp), "%.2fs" % (time.time()-t), flush=True)
t = time.time(); e.set_graph(rp, cp); print("set_graph %.2fs" % (time.time()-t), flush=True)
w, N, Nx = elements.tables(4)
t = time.time(); e.set_mesh(0, m.IEN, w, N, Nx); print("set_mesh %.2fs" % (time.time()-t), flush=True)
e.set_coords(m.x)
e.alloc(4); e.set_state(Ag, Yg, Dg)
print("fp64 peak %.2f TF" % e.fp64_peak())
for sc, name in ((abi.SCATTER_ATOMIC, "atomic"), (abi.SCATTER_COLORED, "colored")):
    eq = abi.fluid_eq(1e-3, scatter=sc); dm = [abi.fluid_domain()]
    e.alloc(4); e.assemble(0, eq, dm)
    ms = e.bench_assemble(0, eq, dm, 5)
    print("assemble %s: %.3f ms  %.3f Gel/s  alg %.2f TF" % (name, ms, m.nEl/ms*1e-6, m.nEl*11.6e3/ms*1e-9), flush=True)
eq = abi.fluid_eq(1e-3); dm = [abi.fluid_domain()]
e.alloc(4); e.assemble(0, eq, dm)
ms = e.bench_spmv(4, 10)
by = len(cp)*132 + m.nNo*72
print("spmv: %.3f ms  %.1f GB/s" % (ms, by/ms*1e-6), flush=True)
faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order='F')) for k in ("wall", "inlet")]
e.set_num_faces(2)
for i, (g, nodes, val) in enumerate(faces): e.set_face(i, g, nodes, val)
ls = abi.ls_params(abi.LS_GMRES, mItr=100, sD=50, relTol=1e-6)
t = time.time()
X, out, hist = e.solve(4, abi.LS_GMRES, ls, np.ones(2, np.int32), np.zeros(2), hist_cap=512, want_solution=False)
print("gmres: itr %d success %d iNorm %.3e fNorm %.3e  wall %.1f ms  dev %.1f ms" % (out.RI.itr, out.RI.success, out.RI.iNorm, out.RI.fNorm, (time.time()-t)*1e3, e.last_timing()[1]), flush=True)
print("launches", e.launch_count)
