"""A/B of every lane mapping of the rectangular-block SpMV kernels and of the Schur operator on the C2 tet graph (and, with
`hex`, the C4 hex8 graph).  Usage: python tools/spmv_variants.py [tet|hex|both] [reps=5]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import meshgen
from svmultiphysics_b200.engine import Engine
which = sys.argv[1] if len(sys.argv) > 1 else "tet"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def run(name, nNo, IEN, shapes):
    e = Engine(0)
    rp, cp = e.lhsa(nNo, [IEN]); e.set_graph(rp, cp)
    nnz = len(cp)
    print(f"== {name}: nNo {nNo} nnz {nnz} ({nnz / nNo:.1f} blocks per row)")
    for R, Cc in shapes:
        algo = nnz * (R * Cc * 8 + 4) + nNo * (8 * R + 8 * Cc + 4)
        for v in range(e.spmv_rc_variants(R, Cc)):
            ms = min(e.bench_spmv_rc(R, Cc, v, reps) for _ in range(2))
            print(f"spmv {R}x{Cc} variant {v}: {ms * 1e3:8.1f} us  {algo / ms / 1e6:7.0f} GB/s algorithmic")
    if (3, 1) in shapes:
        algo = nnz * 36 + nNo * (8 + 32 + 4)
        for v in [-2] + list(range(e.schur_sp_variants())):
            ms = min(e.bench_schur_sp(v, reps) for _ in range(2))
            print(f"schur_sp variant {v}: {ms * 1e3:8.1f} us  {algo / ms / 1e6:7.0f} GB/s algorithmic (nnz*36)")
    e.close()


if which == "prof":
    # one launch (+ warm-up) of selected kernels for an `ncu --set full` capture
    m = meshgen.cylinder_tet4(118, 120)
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
    sel = sys.argv[2] if len(sys.argv) > 2 else "33:0,33:2,33:4,31:0,31:2,31:4,13:0,13:2,11:0,11:1,s:-2,s:0,s:2"
    for item in sel.split(","):
        a, v = item.split(":")
        if a == "s":
            e.bench_schur_sp(int(v), 1)
        else:
            e.bench_spmv_rc(int(a[0]), int(a[1]), int(v), 1)
    e.close()
    sys.exit(0)
if which in ("tet", "both"):
    m = meshgen.cylinder_tet4(118, 120)
    run("C2 tet4", m.nNo, m.IEN, [(3, 3), (3, 1), (1, 3), (1, 1)])
if which in ("hex", "both"):
    m = meshgen.box_hex8(171, 171, 171)
    run("C4 hex8", m.nNo, m.IEN, [(3, 3), (1, 1)])
