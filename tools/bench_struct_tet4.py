"""struct_3d on TET4 (the solid part of the FSI pipe, config C5): assembly time of an n^3 x 6 tet block on one B200.
Usage: python tools/bench_struct_tet4.py [n=90] [reps=3]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 90
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = meshgen.box_tet4(n, n, n, (1.0, 1.0, 1.0))
Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
Dg *= 0.02
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(4); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
e.alloc(3); e.set_state(Ag, Yg, Dg, Bf)
for label, dkw in (("nHK/M94", dict(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)))),
                   ("MR/ST91", dict(isoType=abi.ISO_MR, C10=1e5, C01=3e4, Kpen=1e7, rho=1.0))):
    eq, dm = abi.struct_eq(1e-4), [abi.struct_domain(**dkw)]
    e.alloc(3); e.assemble(0, eq, dm)
    for _ in range(reps):
        e.alloc(3)
        e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1)
        ms = e.timer_elapsed()
    print(f"struct tet4 {label}: {m.nEl} el, nnz {len(cp)}: {ms:.3f} ms  {m.nEl/ms*1e-6:.3f} G el/s")
