"""Experiment: does an element order with COMPACT 128-element groups speed up the grouped TET4 fluid assembly?  The mesh generator's
lattice order (strips of 21 hexes per group) against a windowed Morton order of the element centroids (coarse order kept, bricks inside
windows of W elements).  Usage: python tools/ab_reorder.py [n=118] [nz=120] [reps=10] [W=262144,...]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 120
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
Ws = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 16384, 262144, 1 << 30]
m = meshgen.cylinder_tet4(n, nz)
Ag, Yg, Dg = meshgen.poiseuille_state(m)


def part1by2(v):
    v = v.astype(np.uint64) & 0x3FF
    v = (v | (v << 16)) & 0x30000FF
    v = (v | (v << 8)) & 0x300F00F
    v = (v | (v << 4)) & 0x30C30C3
    v = (v | (v << 2)) & 0x9249249
    return v


def windowed_morton(m, W):
    cen = m.x[:, m.IEN].mean(axis=1)          # (3, nEl)
    lo, hi = cen.min(axis=1, keepdims=True), cen.max(axis=1, keepdims=True)
    # quantise with the SAME cell size on all axes so that Morton cells are cubes
    h = ((hi - lo).max()) / 1023.0
    q = np.minimum(((cen - lo) / h).astype(np.int64), 1023)
    code = part1by2(q[0]) | (part1by2(q[1]) << 1) | (part1by2(q[2]) << 2)
    win = (np.arange(m.nEl) // W).astype(np.uint64)
    return np.lexsort((code, win))


for W in Ws:
    IEN = m.IEN if W == 0 else np.asfortranarray(m.IEN[:, windowed_morton(m, W)])
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [IEN]); e.set_graph(rp, cp)
    w, N, Nx = elements.tables(4); e.set_mesh(0, IEN, w, N, Nx); e.set_coords(m.x)
    e.alloc(4); e.set_state(Ag, Yg, Dg)
    eq = abi.fluid_eq(1e-3); dm = [abi.fluid_domain()]
    e.bench_assemble(0, eq, dm, 2)
    ms = min(e.bench_assemble(0, eq, dm, reps) for _ in range(3))
    print(f"W {W:>10d}: assemble {ms:.3f} ms  {m.nEl / ms * 1e-6:.3f} G el/s", flush=True)
    e.close()
