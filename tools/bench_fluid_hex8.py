"""HEX8 VMS fluid through the general-element kernel (assemble_fluid_gen.cu): n^3 skewed hexahedra on one B200.
Usage: python tools/bench_fluid_hex8.py [n=128] [reps=3]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.time()
m = meshgen.box_hex8(n, n, n, (1.0, 1.0, 1.0))
rng = np.random.default_rng(17)
m.x = np.asfortranarray(m.x + (0.1 / n) * rng.standard_normal(m.x.shape))
Ag, Yg, Dg, Bf = common.fluid_gen_state(m, 4)
e = Engine(0)
rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(8)
e.set_mesh(0, m.IEN, w, N, Nx, Nxx=elements.nxx_tables(8)); e.set_coords(m.x)
e.alloc(4); e.set_state(Ag, Yg, Dg, Bf)
eq, dm = abi.fluid_eq(1e-3), [abi.fluid_domain()]
print(f"setup {time.time()-t0:.1f} s: {m.nEl} hex8, {m.nNo} nodes, nnz {len(cp)} blocks ({len(cp)*128/1e9:.2f} GB Val)")
e.alloc(4); e.assemble(0, eq, dm)
for _ in range(reps):
    e.alloc(4)
    e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1)
    ms = e.timer_elapsed()
    print(f"hex8 fluid assemble {ms:.3f} ms  {m.nEl/ms*1e-6:.4f} G el/s")
ms = e.bench_spmv(4, 10)
print(f"SpMV dof=4 {ms:.3f} ms  {(len(cp)*132 + m.nNo*72)/ms*1e-6:.1f} GB/s algorithmic")
