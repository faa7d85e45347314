"""Assembly time of the linear-elasticity equation and the mesh-motion equation (l_elas_3d) on a HEX8 block.
Usage: python tools/bench_lelas_hex8.py [n=100] [reps=3]   (SVB200_MESH_HEX8_LEGACY=1: the lane-per-row kernel)"""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = meshgen.box_hex8(n, n, n, (1.0, 1.0, 1.0))
m.x = np.asfortranarray(m.x + (0.1 / n) * np.random.default_rng(17).standard_normal(m.x.shape))
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(8); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
for label, tDof, eq, dm in (("lElas", 3, abi.lelas_eq(1e-3), [abi.lelas_domain(E=1.0e6, nu=0.3, rho=2.0, f=(0.1, -0.2, 0.3))]),
                            ("mesh ", 7, abi.mesh_eq(1e-3), [abi.mesh_domain(E=1.0, nu=0.3)])):
    Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0, tDof=tDof)
    if tDof == 7:
        Dg[4:7] = Dg[0:3]
    e.alloc(3); e.set_state(Ag, Yg, Dg, Bf)
    if tDof == 7:
        e.set_old_disp(np.asfortranarray(0.9 * Dg))
    e.assemble(0, eq, dm)
    best = 1e30
    for _ in range(reps):
        e.alloc(3); e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1); best = min(best, e.timer_elapsed())
    print(f"{label} hex8: {m.nEl} el, nnz {len(cp)}: {best:.3f} ms  {m.nEl / best * 1e-6:.3f} G el/s")
e.close()
