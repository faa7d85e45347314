"""A/B timing of the fluid tet4 assembly kernel (set SVB200_ASM_LEGACY=1 for the per-entry RED scatter / per-element colours).
Usage: python tools/ab_assemble.py [n=118] [nz=120] [reps=10] [atomic|colored]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 120
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
scatter = abi.SCATTER_COLORED if (len(sys.argv) > 4 and sys.argv[4] == 'colored') else abi.SCATTER_ATOMIC
m = meshgen.cylinder_tet4(n, nz)
Ag, Yg, Dg = meshgen.poiseuille_state(m)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
t0 = time.time()
w, N, Nx = elements.tables(4); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
t1 = time.time()
e.alloc(4); e.set_state(Ag, Yg, Dg)
eq = abi.fluid_eq(1e-3, scatter=scatter); dm = [abi.fluid_domain()]
e.bench_assemble(0, eq, dm, 2)
for _ in range(3):
    ms = e.bench_assemble(0, eq, dm, reps)
    print(f"{'colored' if scatter else 'atomic'}: nEl {m.nEl}  assemble {ms:.3f} ms  {m.nEl/ms*1e-6:.3f} G el/s  (set_mesh {t1-t0:.2f} s)")
