"""C5 of SURVEY.md 8(d) on one B200: FSI pipe (fluid core + solid wall, tDof = 7) — coupled FSI assembly (construct_fsi: ALE
fluid tets + struct_3d tets, dof 4), the mesh-motion assembly (construct_mesh, dof 3), GMRES on the FSI system and CG on the
mesh system.  Usage: python tools/bench_fsi.py [n=90] [nz=120] [reps=3]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements
from svmultiphysics_b200.engine import Engine
from tests import common

n = int(sys.argv[1]) if len(sys.argv) > 1 else 90
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 120
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.time()
m, Ag, Yg, Dg, Bf = common.fsi_case(n=n, nz=nz)
Dg *= 0.05 * (6.0 / n)          # keep the random displacements well inside the (finer) elements
nSolid = int((m.eId == 2).sum())
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(4); e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId); e.set_coords(m.x)
af, am, gam, beta = abi.gen_alpha(0.5)
eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                  scatter=abi.SCATTER_ATOMIC, reserved=0)
dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0), abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
print(f"setup {time.time()-t0:.1f} s: {m.nEl} tet4 ({m.nEl - nSolid} fluid + {nSolid} solid), {m.nNo} nodes, nnz {len(cp)}")
e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, eq, dmn)
for _ in range(reps):
    e.alloc(4)
    e.timer_mark(0); e.assemble(0, eq, dmn); e.timer_mark(1)
    ms = e.timer_elapsed()
    print(f"FSI assemble (construct_fsi) {ms:.3f} ms  {m.nEl/ms*1e-6:.3f} G el/s")
# the layout of tests/cases/fsi/pipe_3d: lumen and wall as two meshes over the same nodes, assembled mesh by mesh
fl, so = np.where(m.eId == 1)[0], np.where(m.eId == 2)[0]
e2 = Engine(0)
e2.set_graph(rp, cp)
e2.set_mesh(0, np.asfortranarray(m.IEN[:, fl]), w, N, Nx, eId=m.eId[fl]); e2.set_mesh(1, np.asfortranarray(m.IEN[:, so]), w, N, Nx, eId=m.eId[so])
e2.set_coords(m.x)
e2.alloc(4); e2.set_state(Ag, Yg, Dg, Bf); e2.assemble(0, eq, dmn); e2.assemble(1, eq, dmn)
for _ in range(reps):
    e2.alloc(4)
    e2.timer_mark(0); e2.assemble(0, eq, dmn); e2.assemble(1, eq, dmn); e2.timer_mark(1)
    ms = e2.timer_elapsed()
    print(f"FSI assemble, lumen + wall as two meshes {ms:.3f} ms  {m.nEl/ms*1e-6:.3f} G el/s")
e2.close()
wall = m.faces["wall"]
e.set_num_faces(1); e.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))
ls = abi.ls_params(abi.LS_GMRES, mItr=2, sD=50, relTol=1e-8)
e.timer_mark(0)
_, out, _ = e.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1), want_solution=False)
e.timer_mark(1)
print(f"FSI GMRES(50): {out.RI.itr} its in {e.timer_elapsed():.2f} ms ({e.timer_elapsed()/max(out.RI.itr,1):.3f} ms/it)  iNorm {out.RI.iNorm:.3e} fNorm {out.RI.fNorm:.3e}")
# mesh-motion equation on the same mesh (every element), dof 3, state dofs 4..6
eqm, dmm = abi.mesh_eq(1e-3), [abi.mesh_domain(E=1.0, nu=0.3)]
Do = np.asfortranarray(0.9 * Dg)
e.alloc(3); e.set_old_disp(Do); e.assemble(0, eqm, dmm)
for _ in range(reps):
    e.alloc(3)
    e.timer_mark(0); e.assemble(0, eqm, dmm); e.timer_mark(1)
    ms = e.timer_elapsed()
    print(f"mesh assemble (construct_mesh) {ms:.3f} ms  {m.nEl/ms*1e-6:.3f} G el/s")
lsc = abi.ls_params(abi.LS_CG, mItr=100, relTol=1e-10)
e.timer_mark(0)
_, out, _ = e.solve(3, abi.LS_CG, lsc, np.ones(1, np.int32), np.zeros(1), want_solution=False)
e.timer_mark(1)
print(f"mesh CG: {out.RI.itr} its in {e.timer_elapsed():.2f} ms ({e.timer_elapsed()/max(out.RI.itr,1):.3f} ms/it)")
