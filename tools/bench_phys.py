"""Assembly time of the SURVEY 8f rank-4 physics on one B200: heatS / heatF on a TET4 cylinder and a HEX8 block, ustruct on
TET4 and HEX8 blocks.  Usage: python tools/bench_phys.py [n_tet=80] [n_hex=100] [reps=3] [heat|ustruct]
Prints elements/s and the compulsory-traffic figure (RMW of the CSR values + nodal gather) next to the HBM peak."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common

n_tet = int(sys.argv[1]) if len(sys.argv) > 1 else 80
n_hex = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
only = sys.argv[4] if len(sys.argv) > 4 else ""      # "heat" or "ustruct"


def run(label, m, dof, state, eq, dm, val_bytes_per_blk, extra=None):
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
    w, N, Nx = elements.tables(m.eNoN); e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
    Ag, Yg, Dg, Bf = state
    e.alloc(dof); e.set_state(Ag, Yg, Dg, Bf)
    e.assemble(0, eq, dm)
    best = 1e30
    for _ in range(reps):
        e.alloc(dof)
        e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1)
        best = min(best, e.timer_elapsed())
    nnz = len(cp)
    traffic = nnz * val_bytes_per_blk * 2 + m.nEl * m.eNoN * 4 + m.nNo * (24 + 16 * eq.tDof)
    print(f"{label}: {m.nEl} el, {m.nNo} nodes, nnz {nnz}: {best:.3f} ms  {m.nEl/best*1e-6:.3f} G el/s  "
          f"compulsory {traffic/1e9:.2f} GB -> {traffic/best*1e-6:.0f} GB/s")
    e.close()


mt = meshgen.cylinder_tet4(n_tet, n_tet) if only in ("", "heat") else None
mh = meshgen.box_hex8(n_hex, n_hex, n_hex, (1.0, 1.0, 1.0))
for fluid, tDof, s in (((False, 1, 0), (True, 5, 4)) if only in ("", "heat") else ()):
    name = "heatF" if fluid else "heatS"
    for m in (mt, mh):
        Ag, Yg, Dg, Bf = common.heat_state(m, tDof, s)
        run(f"{name} {'tet4' if m.eNoN == 4 else 'hex8'}", m, 1, (Ag, Yg, Dg, Bf), abi.heat_eq(0.01, fluid, tDof=tDof, s=s),
            [abi.heat_domain(fluid, conductivity=0.5, source=1.0, rho=2.0)], 8)
for m in ((meshgen.box_tet4(n_hex // 2, n_hex // 2, n_hex // 2, (1.0, 1.0, 1.0)), mh) if only in ("", "ustruct") else ()):
    Ag, Yg, Dg, Bf, _ = common.ustruct_state(m)
    Dg *= 0.05
    run(f"ustruct {'tet4' if m.eNoN == 4 else 'hex8'}", m, 4, (Ag, Yg, Dg, Bf), abi.ustruct_eq(1e-3),
        [abi.ustruct_domain(E=1.0e6, nu=0.45, Kpen=1.0e6 / (3 * (1 - 0.9)), rho=1.2)], 128 + 96)
