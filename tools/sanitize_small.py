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


only = sys.argv[1] if len(sys.argv) > 1 else ""     # "", "heat", "ustruct", "struct", "fluid", "new" (round-2 session-2 kernels only)


def engine(m, nFn=0, fN=None):
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
    w, N, Nx = elements.tables(m.eNoN); e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, nFn=nFn, fN=fN); e.set_coords(m.x)
    return e


for sc in ((abi.SCATTER_ATOMIC, abi.SCATTER_COLORED) if only != "new" else ()):
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
        d = abi.ustruct_domain(**dkw)
        e.alloc(4); e.set_state(Ag, Yg, Dg, Bf)
        if d.active_stress:
            e.set_active_tension(*common.active_tension(m, d.isoType))
        e.assemble(0, eq, [d]); e.ustruct_r(eq, 1, common.ustruct_Ad(m))
        e.get_Kd(); e.close()
    for name, mk, dkw, nFn in (common.STRUCT_CASES if only in ("", "struct") else []):
        m = mk()
        Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
        e = engine(m, nFn, fN)
        d = abi.struct_domain(**dkw)
        e.alloc(3); e.set_state(Ag, Yg, Dg, Bf)
        if d.active_stress:
            e.set_active_tension(*common.active_tension(m, d.isoType))
        e.assemble(0, abi.struct_eq(1e-4, scatter=sc), [d])
        e.get_Val(); e.close()
    if only not in ("", "struct", "fluid"):
        continue
    m = meshgen.box_tet4(3, 3, 2, (1.0, 1.0, 1.0)); e = engine(m)
    Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
    e.alloc(3); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.lelas_eq(1e-3, scatter=sc), [abi.lelas_domain()]); e.get_Val(); e.close()
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=4, nz=5); e = engine(m)
    e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.fluid_eq(0.005, scatter=sc), [abi.fluid_domain()]); e.get_Val(); e.close()


def new_kernels():
    """Round 2, session 2: lane-group SpMVs (every variant), the interleaved Schur operator, the fused CG tail inside an NS solve,
    quadratic / wedge fluid and solid elements, prestress, active stress, CANN."""
    golden = common.load_golden("fluid_hi.npz")
    rng = np.random.default_rng(5)
    m = meshgen.cylinder_tet4(4, 3)
    e = Engine(0)
    rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
    for R, Cc in ((3, 3), (3, 1), (1, 3), (1, 1)):
        K = np.asfortranarray(rng.standard_normal((R * Cc, len(cp)))); U = np.asfortranarray(rng.standard_normal((Cc, m.nNo)))
        for v in range(e.spmv_rc_variants(R, Cc)):
            e.spmv_rc(R, Cc, K, U, variant=v)
    for v in [-2] + list(range(e.schur_sp_variants())):
        e.schur_sp(rng.standard_normal(len(cp)), rng.standard_normal((3, len(cp))), rng.standard_normal(m.nNo),
                   rng.standard_normal((3, m.nNo)), variant=v)
    e.close()
    # NS solve (depart with DL, schur_sp4 + fused tail, coupled resistance outlet)
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=4, nz=5)
    e = engine(m)
    faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("wall", "inlet")]
    out = m.faces["outlet"]; val = np.zeros((3, len(out)), order="F"); val[2] = 4.0 * np.pi / len(out)
    faces.append((abi.BC_NEU, out, val))
    e.set_num_faces(3)
    for i, (g, nodes, v) in enumerate(faces):
        e.set_face(i, g, nodes, v)
    e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.fluid_eq(0.005), [abi.fluid_domain()])
    ls = abi.ls_params(abi.LS_NS, mItr=5, sD=50, relTol=1e-3, absTol=1e-17, gm=(5, 50, 1e-3, 1e-17), cg=(50, 0, 1e-3, 1e-17))
    e.solve(4, abi.LS_NS, ls, np.ones(3, np.int32), np.array([0.0, 0.0, 0.8]))
    e.close()
    for sc in (abi.SCATTER_ATOMIC, abi.SCATTER_COLORED):
        for name, mk, visc, Kd, f, tDof, mv in common.FLUID_HI_CASES:
            mm = mk(); et = name.split("_")[0]
            w, N, Nx, Nxx = (golden[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
            A, Y, D, B = common.fluid_gen_state(mm, tDof)
            e = Engine(0); rp, cp = e.lhsa(mm.nNo, [mm.IEN]); e.set_graph(rp, cp)
            e.set_mesh(0, mm.IEN, w, N, Nx, Nxx=Nxx); e.set_coords(mm.x)
            e.alloc(4); e.set_state(A, Y, D, B); e.assemble(0, abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=sc), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
            e.get_Val(); e.close()
        for name, mk, dkw, nFn in common.STRUCT_HI_CASES:
            mm = mk(); et = name.split("_")[0]
            w, N, Nx = (golden[f"tables/{et}/{k}"] for k in ("w", "N", "Nx"))
            A, Y, D, B, fN = common.struct_state(mm, nFn)
            e = Engine(0); rp, cp = e.lhsa(mm.nNo, [mm.IEN]); e.set_graph(rp, cp)
            e.set_mesh(0, mm.IEN, w, N, Nx, nFn=nFn, fN=fN); e.set_coords(mm.x)
            d = abi.struct_domain(**dkw)
            e.alloc(3); e.set_state(A, Y, D, B)
            if d.active_stress:
                e.set_active_tension(*common.active_tension(mm, d.isoType))
            e.assemble(0, abi.struct_eq(1e-4, scatter=sc), [d]); e.get_Val(); e.close()
    for name, *_ in common.PRESTRESS_CASES:
        mm, A, Y, D, B, pS0, eq, dmn = common.prestress_case(name)
        e = engine(mm); e.set_prestress(pS0)
        e.alloc(3); e.set_state(A, Y, D, B); e.assemble(0, eq, dmn)
        if eq.reserved & abi.EQ_PRESTRESS:
            e.get_prestress()
        e.close()


def session3_kernels():
    """Round 2, session 3: FSI with a ustruct wall + masked ustruct_r, URIS valves (split launch and general kernel), fitted RIS
    (take / put / apply), heat / lElas / mesh on quadratic and wedge elements, the HEX8 heat and lElas kernels with one lane per Gauss
    point, the rolled ustruct TET4 kernel, Taylor-Hood fluid + thood_val_rc, degenerate sizes."""
    tabs, th = common.load_golden("fluid_hi.npz"), common.load_golden("fluid_thood.npz")
    for sc in (abi.SCATTER_ATOMIC, abi.SCATTER_COLORED):
        for name in common.FSI_USTRUCT_CASES:
            m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, Ad, flags = common.fsi_ustruct_case(name, sc)
            e = engine(m, nFn, fN)
            e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, eq, dmn)
            e.set_node_flags(flags); e.ustruct_r(eq, 1, Ad); e.get_Kd(); e.get_Rd(); e.close()
        for name, *_ in common.URIS_CASES:
            m, Ag, Yg, Dg, Bf, eq, dmn = common.uris_case(name, sc)
            raw, dev, sdf, udf, vel = common.uris_valves(m)
            e = Engine(0); rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
            w, N, Nx = elements.tables(m.eNoN)
            e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, Nxx=elements.nxx_tables(m.eNoN) if m.eNoN != 4 else None); e.set_coords(m.x)
            e.set_uris(dev, sdf, udf, vel)
            e.alloc(4); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, eq, dmn); e.get_Val()
            e.set_uris([]); e.alloc(4); e.assemble(0, eq, dmn); e.close()
        g = common.load_golden("ris.npz")
        x, IENs, mp, Ag, Yg, Bf, eq, dmn = common.ris_case(sc)
        e = Engine(0); e.set_graph(g["rowPtr"], g["colPtr"])
        w, N, Nx = elements.tables(4)
        for iM, I in enumerate(IENs):
            e.set_mesh(iM, I, w, N, Nx)
        e.set_coords(x); e.set_ris([mp], [0])
        e.alloc(4); e.set_state(Ag, Yg, None, Bf)
        for iM in range(len(IENs)):
            e.assemble(iM, eq, dmn)
        e.get_Val(); e.close()
        for name, *_ in common.OTHER_HI_CASES:
            m, et, dof, Ag, Yg, Dg, Bf, Do, eq, dmn = common.other_hi_case(name, sc)
            w, N, Nx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx"))
            e = Engine(0); rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
            e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
            e.alloc(dof); e.set_state(Ag, Yg, Dg, Bf)
            if Do is not None:
                e.set_old_disp(Do)
            e.assemble(0, eq, dmn); e.get_Val(); e.close()
        for name, mk, fluid, tDof, s_, mv, dkw in common.HEAT_CASES:            # HEX8: the lane-per-Gauss-point kernel
            m = mk(); e = engine(m)
            Ag, Yg, Dg, Bf = common.heat_state(m, tDof, s_)
            e.alloc(1); e.set_state(Ag, Yg, Dg, Bf)
            e.assemble(0, abi.heat_eq(0.01, fluid, tDof=tDof, s=s_, mvMsh=mv, scatter=sc), [abi.heat_domain(fluid, **dkw)]); e.get_Val(); e.close()
        m = meshgen.box_hex8(3, 2, 2, (1.0, 1.0, 1.0)); e = engine(m)
        Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
        e.alloc(3); e.set_state(Ag, Yg, Dg, Bf); e.assemble(0, abi.lelas_eq(1e-3, scatter=sc), [abi.lelas_domain()]); e.get_Val(); e.close()
        for name, mk, dkw, nFn in [c for c in common.USTRUCT_CASES if c[0].startswith("tet4")]:
            m = mk(); Ag, Yg, Dg, Bf, fN = common.ustruct_state(m, nFn); e = engine(m, nFn, fN)
            d = abi.ustruct_domain(**dkw)
            e.alloc(4); e.set_state(Ag, Yg, Dg, Bf)
            if d.active_stress:
                e.set_active_tension(*common.active_tension(m, d.isoType))
            e.assemble(0, abi.ustruct_eq(1e-3, scatter=sc), [d]); e.get_Kd(); e.close()
        for name, mk, visc, Kd, f, tDof, mv in common.FLUID_THOOD_CASES:
            m = mk(); et = name.split("_")[0]
            w, N, Nx, Nxx = (tabs[f"tables/{et}/{k}"] for k in ("w", "N", "Nx", "Nxx"))
            t = {k: th[f"tables/{et}/{k}"] for k in ("eNoNq", "nG1", "nG2", "lShpF_q", "Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
            A, Y, D, B = common.fluid_gen_state(m, tDof)
            e = Engine(0); rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
            e.set_mesh(0, m.IEN, w, N, Nx, Nxx=Nxx); e.set_mesh_thood(0, t); e.set_coords(m.x)
            e.alloc(4); e.set_state(A, Y, D, B)
            e.assemble(0, common.fluid_thood_eq(0.005, tDof=tDof, mvMsh=mv, scatter=sc), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
            e.thood_val_rc(); e.get_Val(); e.close()
    # degenerate sizes: one element, a partial group
    big = meshgen.cylinder_tet4(4, 3)
    for nEl in (1, 129):
        import copy
        I = big.IEN[:, :nEl]; nodes, inv = np.unique(I, return_inverse=True)
        m = copy.copy(big); m.x = np.asfortranarray(big.x[:, nodes]); m.IEN = np.asfortranarray(inv.reshape(I.shape).astype(np.int32)); m.eId = None
        e = engine(m)
        rng = np.random.default_rng(1)
        e.alloc(4); e.set_state(np.asfortranarray(rng.standard_normal((4, m.nNo))), np.asfortranarray(rng.standard_normal((4, m.nNo))), None, None)
        e.assemble(0, abi.fluid_eq(0.005), [abi.fluid_domain()]); e.get_Val(); e.close()


if only in ("", "new"):
    new_kernels()
if only in ("", "s3"):
    session3_kernels()
print("sanitize_small: done")
