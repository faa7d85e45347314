"""Small run of this session's kernels for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
heatS/heatF, ustruct (TET4 + HEX8, with and without solid viscosity) incl. ustruct_r, the TET4 solid and mesh kernels, the
HEX8 solid kernel (tile scatter) and the group-coloured deterministic fluid assembly, both scatter modes, on tiny meshes."""
import sys
import numpy as np
sys.path.insert(0, '.')
from svmultiphysics_b200 import abi, elements, meshgen
from svmultiphysics_b200.engine import Engine
from tests import common


only = sys.argv[1] if len(sys.argv) > 1 else ""     # "", "heat", "ustruct", "struct", "fluid"


def engine(m, nFn=0, fN=None):
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
    w, N, Nx = elements.tables(m.eNoN); e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, nFn=nFn, fN=fN); e.set_coords(m.x)
    return e


for sc in (abi.SCATTER_ATOMIC, abi.SCATTER_COLORED):
    for name, mk, fluid, tDof, s, mv, dkw in (common.HEAT_CASES if only in ("", "heat") else []):
        m = mk(); e = engine(m)
        Ag, Yg, Dg, Bf = common.heat_state(m, tDof, s)
        e.alloc(1); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.heat_eq(0.01, fluid, tDof=tDof, s=s, mvMsh=mv, scatter=sc), [abi.heat_domain(fluid, **dkw)])
        e.get_R(); e.close()
    for name, mk, dkw, nFn in (common.USTRUCT_CASES if only in ("", "ustruct") else []):
        m = mk()
        Ag, Yg, Dg, Bf, fN = common.ustruct_state(m, nFn)
        e = engine(m, nFn, fN)
        eq = abi.ustruct_eq(1e-3, scatter=sc)
        e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, eq, [abi.ustruct_domain(**dkw)]); e.ustruct_r(eq, 1, common.ustruct_Ad(m))
        e.get_Kd(); e.close()
    for name, mk, dkw, nFn in (common.STRUCT_CASES if only in ("", "struct") else []):
        m = mk()
        Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
        e = engine(m, nFn, fN)
        e.alloc(3); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.struct_eq(1e-4, scatter=sc), [abi.struct_domain(**dkw)])
        e.get_Val(); e.close()
    if only not in ("", "struct", "fluid"):
        continue
    m = meshgen.box_tet4(3, 3, 2, (1.0, 1.0, 1.0)); e = engine(m)
    Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
    e.alloc(3); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.lelas_eq(1e-3, scatter=sc), [abi.lelas_domain()]); e.get_Val(); e.close()
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=4, nz=5); e = engine(m)
    e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.fluid_eq(0.005, scatter=sc), [abi.fluid_domain()]); e.get_Val(); e.close()
print("sanitize_small: done")
