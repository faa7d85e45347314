"""Throughput of the general fluid / solid kernels on quadratic elements (TET10: 15 Gauss points, HEX27: 27).
Usage: python tools/bench_quadratic.py [n_tet=36] [n_hex=28]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common
n_tet = int(sys.argv[1]) if len(sys.argv) > 1 else 36
n_hex = int(sys.argv[2]) if len(sys.argv) > 2 else 28
tabs = common.load_golden("fluid_hi.npz")


def run(name, m, et):
    w, N, Nx, Nxx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
    e.set_mesh(0, m.IEN, w, N, Nx, Nxx=Nxx); e.set_coords(m.x)
    A, Y, D, B = common.fluid_gen_state(m, 4)
    eq, dm = abi.fluid_eq(0.005), [abi.fluid_domain()]
    e.alloc(4); e.set_state(A, Y, D, B); e.assemble(0, eq, dm)
    ms = []
    for _ in range(3):
        e.alloc(4); e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1); ms.append(e.timer_elapsed())
    print(f"{name} fluid : {m.nEl} el, {m.nNo} nodes, nnz {len(cp)}: {min(ms):.3f} ms  {m.nEl / min(ms) * 1e-3:.2f} M el/s  {m.nNo / min(ms) * 1e-3:.2f} M nodes/s")
    # the same mesh with Taylor-Hood function spaces (P2-P1 / Q2-Q1, vmsStab = false)
    th = common.load_golden("fluid_thood.npz")
    t = {k: th[f"tables/{et}/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
    e.set_mesh_thood(0, t)
    eqt = common.fluid_thood_eq(0.005)
    e.alloc(4); e.assemble(0, eqt, dm)
    ms = []
    for _ in range(3):
        e.alloc(4); e.timer_mark(0); e.assemble(0, eqt, dm); e.timer_mark(1); ms.append(e.timer_elapsed())
    print(f"{name} Taylor-Hood fluid: {min(ms):.3f} ms  {m.nEl / min(ms) * 1e-3:.2f} M el/s  {m.nNo / min(ms) * 1e-3:.2f} M nodes/s")
    e.set_mesh_thood(0, None)
    A, Y, D, B, _ = common.struct_state(m, 0)
    eqs, dms = abi.struct_eq(1e-4), [abi.struct_domain(E=1e6, nu=0.4, Kpen=1e6, rho=1.0)]
    e.alloc(3); e.set_state(A, Y, D, B); e.assemble(0, eqs, dms)
    ms = []
    for _ in range(3):
        e.alloc(3); e.timer_mark(0); e.assemble(0, eqs, dms); e.timer_mark(1); ms.append(e.timer_elapsed())
    print(f"{name} struct: {min(ms):.3f} ms  {m.nEl / min(ms) * 1e-3:.2f} M el/s  {m.nNo / min(ms) * 1e-3:.2f} M nodes/s")
    e.close()


t0 = time.time()
m = meshgen.elevate(meshgen.box_tet4(n_tet, n_tet, n_tet), "tet10", bend=0.02)
print(f"tet10 mesh {time.time() - t0:.1f} s")
run("TET10", m, "tet10")
t0 = time.time()
m = meshgen.elevate(meshgen.box_hex8(n_hex, n_hex, n_hex), "hex27", bend=0.02)
print(f"hex27 mesh {time.time() - t0:.1f} s")
run("HEX27", m, "hex27")
m = meshgen.box_hex8(60, 60, 60)
m.x = np.asfortranarray(m.x + 0.002 * np.random.default_rng(1).standard_normal(m.x.shape))
from svmultiphysics_b200 import elements
e = Engine(0); rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
w, N, Nx = elements.tables(8); e.set_mesh(0, m.IEN, w, N, Nx, Nxx=elements.nxx_tables(8)); e.set_coords(m.x)
A, Y, D, B = common.fluid_gen_state(m, 4)
eq, dm = abi.fluid_eq(0.005), [abi.fluid_domain()]
e.alloc(4); e.set_state(A, Y, D, B); e.assemble(0, eq, dm)
ms = []
for _ in range(3):
    e.alloc(4); e.timer_mark(0); e.assemble(0, eq, dm); e.timer_mark(1); ms.append(e.timer_elapsed())
print(f"HEX8 fluid : {m.nEl} el: {min(ms):.3f} ms  {m.nEl / min(ms) * 1e-3:.2f} M el/s  {m.nNo / min(ms) * 1e-3:.2f} M nodes/s")
e.close()
