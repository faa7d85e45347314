"""URIS valves on a TET4 cylinder: assembly time with the split launch (band -> per-Gauss-point kernel, rest -> closed form) and, with
SVB200_URIS_NO_SPLIT=1, with every element through the per-Gauss-point kernel.  Usage: python tools/bench_uris.py [n=60] [nz=80]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 80
m = meshgen.cylinder_tet4(n, nz, R=1.0, L=4.0)
Ag, Yg, _ = meshgen.poiseuille_state(m, R=1.0, U=5.0)
raw, dev, sdf, udf, vel = common.uris_valves(m)
for v in dev:                       # thin valves: two element layers
    v.sdf_deps = 2.0 * 4.0 / nz
    v.scaffold = 0
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(4); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
eq, dm = abi.fluid_eq(0.005), [abi.fluid_domain()]
e.alloc(4); e.set_state(Ag, Yg, None, None)
for label in ("no valves", "two valves"):
    if label == "two valves":
        e.set_uris(dev, sdf, None, vel)
    e.assemble(0, eq, dm)
    best = 1e30
    for _ in range(3):
        e.alloc(4); e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1); best = min(best, e.timer_elapsed())
    near = float((np.abs(sdf) < dev[0].sdf_deps).any(axis=0).mean())
    print(f"{label}: {m.nEl} tet4: {best:.3f} ms  {m.nEl / best * 1e-6:.3f} G el/s  (nodes inside a valve thickness: {100 * near:.1f} %)")
e.close()
