"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref/libsvref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
Inputs are regenerated deterministically by tests/common.py (seeded), only the reference OUTPUTS are
stored: CSR structure, assembled R / Val, GMRES / BiCGStab / CG results, SpMV product, element tables.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.refbind import RefCase, have_ref  # noqa: E402
from svmultiphysics_b200 import abi, meshgen  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fluid_golden():
    out = {}
    for name, visc, Kd, f, tDof, mv in common.FLUID_CASES:
        m, Ag, Yg, Dg, Bf = common.fluid_case(n=common.GOLDEN_N, nz=common.GOLDEN_NZ, tDof=tDof)
        faces = common.dirichlet_faces(m)
        c, rowPtr, colPtr = common.make_oracle(RefCase, m, nFaces=len(faces))
        for i, (g, nodes, val) in enumerate(faces):
            c.set_face(i, g, nodes, val)
        eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv)
        dmn = [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        U = common.spmv_vector(m.nNo)
        out[f"{name}/KU"] = c.spmv(4, U)
        for ls_name, ls_type, kw in common.LS_CASES:
            c.put_R(out[f"{name}/R"]); c.put_Val(out[f"{name}/Val"], 4)
            ls = abi.ls_params(ls_type, **kw)
            X, o, _ = c.solve(4, ls_type, ls, np.ones(len(faces), np.int32), np.zeros(len(faces)))
            out[f"{name}/{ls_name}/X"] = X
            out[f"{name}/{ls_name}/stats"] = np.array([o.RI.itr, o.RI.success, o.RI.iNorm, o.RI.fNorm, o.RI.dB])
        out["rowPtr"], out["colPtr"] = rowPtr, colPtr
    for eNoN, mk in ((4, lambda: meshgen.cylinder_tet4(2, 2)), (8, lambda: meshgen.box_hex8(2, 2, 2))):
        m = mk()
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        w, N, Nx = c.mesh_tables(0)
        out[f"tables{eNoN}/w"], out[f"tables{eNoN}/N"], out[f"tables{eNoN}/Nx"] = w, N, Nx
    np.savez_compressed(os.path.join(HERE, "fluid_tet4.npz"), **out)
    print("wrote fluid_tet4.npz with", len(out), "arrays")


def struct_golden():
    """Assembled R / Val of every solid case of tests/common.py (struct_3d + compute_pk2cc + solid viscosity)."""
    out = {}
    for name, mk, dkw, nFn in common.STRUCT_CASES:
        m = mk()
        Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN, nFn=nFn, fN=fN)
        rowPtr, colPtr = c.build_graph(0)
        d = abi.struct_domain(**dkw)
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf)
        if d.active_stress:
            c.set_active_tension(*common.active_tension(m, d.isoType))
        c.assemble(0, abi.struct_eq(1e-4), [d])
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
    np.savez_compressed(os.path.join(HERE, "struct.npz"), **out)
    print("wrote struct.npz with", len(out), "arrays")


def fluid_gen_golden():
    """HEX8 (and TET4 through the same general path) VMS fluid: gnn + gn_nxx per Gauss point, fluid_3d_m / fluid_3d_c."""
    out = {}
    for name, mk, visc, Kd, f, tDof, mv in common.FLUID_GEN_CASES:
        m = mk()
        Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf)
        c.assemble(0, abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        out[f"{name}/Nxx"] = c.mesh_nxx(0)
    np.savez_compressed(os.path.join(HERE, "fluid_gen.npz"), **out)
    print("wrote fluid_gen.npz with", len(out), "arrays")


def fluid_hi_golden():
    """VMS fluid on TET10 / HEX20 / HEX27 / WDG (curved elements), with the reference's own element tables: the tests feed them
    to svb200_set_mesh / svb200_set_mesh_nxx, as the plug-in does from mshType."""
    out = {}
    for name, mk, visc, Kd, f, tDof, mv in common.FLUID_HI_CASES:
        m = mk()
        Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf)
        c.assemble(0, abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        et = name.split("_")[0]
        out[f"tables/{et}/w"], out[f"tables/{et}/N"], out[f"tables/{et}/Nx"] = c.mesh_tables(0)
        out[f"tables/{et}/Nxx"] = c.mesh_nxx(0)
    np.savez_compressed(os.path.join(HERE, "fluid_hi.npz"), **out)
    print("wrote fluid_hi.npz with", len(out), "arrays")


def struct_hi_golden():
    """struct_3d on curved TET10 / HEX20 / HEX27 / WDG elements (element tables: tests/golden/fluid_hi.npz)."""
    out = {}
    for name, mk, dkw, nFn in common.STRUCT_HI_CASES:
        m = mk()
        Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN, nFn=nFn, fN=fN)
        rowPtr, colPtr = c.build_graph(0)
        d = abi.struct_domain(**dkw)
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf)
        if d.active_stress:
            c.set_active_tension(*common.active_tension(m, d.isoType))
        c.assemble(0, abi.struct_eq(1e-4), [d])
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
    np.savez_compressed(os.path.join(HERE, "struct_hi.npz"), **out)
    print("wrote struct_hi.npz with", len(out), "arrays")


def heat_golden():
    """Assembled R / Val of the scalar heat equations (heats_3d / heatf_3d) for tests/common.py:HEAT_CASES."""
    out = {}
    for name, mk, fluid, tDof, s, mv, dkw in common.HEAT_CASES:
        m = mk()
        Ag, Yg, Dg, Bf = common.heat_state(m, tDof, s)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        rowPtr, colPtr = c.build_graph(0)
        eq, dmn = abi.heat_eq(0.01, fluid, tDof=tDof, s=s, mvMsh=mv), [abi.heat_domain(fluid, **dkw)]
        c.alloc(1); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
    np.savez_compressed(os.path.join(HERE, "heat.npz"), **out)
    print("wrote heat.npz with", len(out), "arrays")


def ustruct_golden():
    """Assembled R / Val / Kd of the mixed solid (ustruct_3d_m/c + ustruct_do_assem) and R after ustruct_r."""
    out = {}
    for name, mk, dkw, nFn in common.USTRUCT_CASES:
        m = mk()
        Ag, Yg, Dg, Bf, fN = common.ustruct_state(m, nFn)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN, nFn=nFn, fN=fN)
        rowPtr, colPtr = c.build_graph(0)
        eq, dmn = abi.ustruct_eq(1e-3), [abi.ustruct_domain(**dkw)]
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf)
        if dmn[0].active_stress:
            c.set_active_tension(*common.active_tension(m, dmn[0].isoType))
        c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"], out[f"{name}/Kd"] = c.get_R(), c.get_Val(), c.get_Kd()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        Ad = common.ustruct_Ad(m)
        c.ustruct_r(1, Ad)
        out[f"{name}/R_after_ustruct_r"] = c.get_R()
    np.savez_compressed(os.path.join(HERE, "ustruct.npz"), **out)
    print("wrote ustruct.npz with", len(out), "arrays")


def fsi_ustruct_golden():
    """construct_fsi with a ustruct wall (fsi.cpp:243-262): R / Val / Kd, and R after ustruct_r (nodes of the ustruct domain only)."""
    out = {}
    for name in common.FSI_USTRUCT_CASES:
        m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, Ad, flags = common.fsi_ustruct_case(name)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN, eId=m.eId, nFn=nFn, fN=fN)
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"], out[f"{name}/Kd"] = c.get_R(), c.get_Val(), c.get_Kd()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        c.ustruct_r(1, Ad)
        out[f"{name}/R_after_ustruct_r"] = c.get_R()
        fl = np.where(flags == 0)[0]
        assert np.abs(out[f"{name}/Kd"]).max() > 0 and common.rel_err(out[f"{name}/R_after_ustruct_r"], out[f"{name}/R"]) > 1e-6
        assert np.array_equal(out[f"{name}/R_after_ustruct_r"][:, fl], out[f"{name}/R"][:, fl])     # fluid-only rows untouched
    np.savez_compressed(os.path.join(HERE, "fsi_ustruct.npz"), **out)
    print("wrote fsi_ustruct.npz with", len(out), "arrays")


def fluid_uris_golden():
    """fluid_3d_m / fluid_3d_c with the URIS penalty terms (construct_fluid and the fluid elements of construct_fsi); the valve factor
    itself comes from the restatement of uris::eval_uris_ris_factors_quadrature in oracle/ref_build/ref_stubs.cpp."""
    out = {}
    for name, *_ in common.URIS_CASES:
        m, Ag, Yg, Dg, Bf, eq, dmn = common.uris_case(name)
        raw, dev, sdf, udf, vel = common.uris_valves(m)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN, eId=m.eId)
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        R0, V0 = c.get_R(), c.get_Val()
        c.set_uris(raw, sdf, udf, vel)
        c.alloc(4); c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        # the valves must matter, and not everywhere: the fixture proves nothing otherwise
        assert common.rel_err(out[f"{name}/R"], R0) > 1e-3 and common.rel_err(out[f"{name}/Val"], V0) > 1e-6, name
        for v in range(len(dev)):          # nodes inside and outside every smeared surface
            assert 0.05 < (np.abs(sdf[v]) < dev[v].sdf_deps).mean() < 0.95, name
        c.set_uris([]); c.alloc(4); c.assemble(0, eq, dmn)
        assert np.array_equal(c.get_R(), R0)
    np.savez_compressed(os.path.join(HERE, "fluid_uris.npz"), **out)
    print("wrote fluid_uris.npz with", len(out), "arrays")


def other_hi_golden():
    """heats_3d / heatf_3d / l_elas_3d (linear-elasticity and mesh-motion equations) on curved TET10 / HEX20 / HEX27 / WDG elements
    (element tables: tests/golden/fluid_hi.npz)."""
    out = {}
    for name, *_ in common.OTHER_HI_CASES:
        m, et, dof, Ag, Yg, Dg, Bf, Do, eq, dmn = common.other_hi_case(name)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(dof); c.set_state(Ag, Yg, Dg, Bf)
        if Do is not None:
            c.set_old_disp(Do)
        c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        assert np.abs(out[f"{name}/R"]).max() > 0 and np.abs(out[f"{name}/Val"]).max() > 0
    np.savez_compressed(os.path.join(HERE, "other_hi.npz"), **out)
    print("wrote other_hi.npz with", len(out), "arrays")


def ris_golden():
    """construct_fluid on two lumen meshes separated by an OPEN resistive immersed surface: the reference's own ris::doassem_ris
    (compiled ris.cpp) adds every element row of a mapped node into its twin's row; CSR graph with the RIS connections of lhsa."""
    out = {}
    x, IENs, mp, Ag, Yg, Bf, eq, dmn = common.ris_case()
    res = {}
    for label, closed in (("open", [0]), ("closed", [1])):
        c = RefCase(); c.set_coords(x)
        for I in IENs:
            c.add_mesh(I)
        c.set_ris([mp], closed, [(0, 1)])
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(4); c.set_state(Ag, Yg, None, Bf)
        for iM in range(len(IENs)):
            c.assemble(iM, eq, dmn)
        res[label] = (c.get_R(), c.get_Val())
        out["rowPtr"], out["colPtr"] = rowPtr, colPtr
    out["open/R"], out["open/Val"] = res["open"]
    out["closed/R"], out["closed/Val"] = res["closed"]
    out["map"] = mp
    # open: rows of twins carry the sum of both sides; closed: the two lumens are uncoupled
    Ro, Rc = res["open"][0], res["closed"][0]
    assert common.rel_err(Ro[:, mp[0]], Rc[:, mp[0]] + Rc[:, mp[1]]) < 1e-13 and common.rel_err(Ro[:, mp[1]], Ro[:, mp[0]]) < 1e-13
    rest = np.setdiff1d(np.arange(x.shape[1]), mp.ravel())
    assert np.array_equal(Ro[:, rest], Rc[:, rest])
    np.savez_compressed(os.path.join(HERE, "ris.npz"), **out)
    print("wrote ris.npz with", len(out), "arrays")


def fluid_thood_golden():
    """Navier-Stokes on Taylor-Hood function spaces (construct_fluid with vmsStab = false): R / Val after the element loop and after
    fs::thood_val_rc, and the four tables of fs::get_thood_fs per element type."""
    out = {}
    for name, mk, visc, Kd, f, tDof, mv in common.FLUID_THOOD_CASES + [common.FLUID_THOOD_URIS_CASE]:
        m = mk()
        Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN); c.set_mesh_thood(0)
        rowPtr, colPtr = c.build_graph(0)
        if name.endswith("uris"):
            raw, dev, sdf, udf, vel = common.uris_valves(m)
            c.set_uris(raw, sdf, udf, vel)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf)
        c.assemble(0, common.fluid_thood_eq(0.005, tDof=tDof, mvMsh=mv), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        c.thood_val_rc()
        out[f"{name}/R_rc"], out[f"{name}/Val_rc"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        et = name.split("_")[0]
        for k, v in c.thood_tables(0).items():
            out[f"tables/{et}/{k}"] = np.asarray(v)
        assert np.abs(out[f"{name}/Val"][12:15]).max() > 0 and not out[f"{name}/Val"][15].any()
    # construct_fsi on a Taylor-Hood mesh: fluid core (ALE geometry) + struct wall
    m, Ag, Yg, Dg, Bf, eq, dmn = common.fsi_thood_case()
    c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN, eId=m.eId); c.set_mesh_thood(0)
    rowPtr, colPtr = c.build_graph(0)
    c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
    out["fsi_tet10/R"], out["fsi_tet10/Val"] = c.get_R(), c.get_Val()
    c.thood_val_rc()
    out["fsi_tet10/R_rc"], out["fsi_tet10/Val_rc"] = c.get_R(), c.get_Val()
    out["fsi_tet10/rowPtr"], out["fsi_tet10/colPtr"] = rowPtr, colPtr
    np.savez_compressed(os.path.join(HERE, "fluid_thood.npz"), **out)
    print("wrote fluid_thood.npz with", len(out), "arrays")


def lelas_golden():
    """R / Val of l_elas_3d on TET4: the linear-elasticity equation and the mesh-motion equation (tDof = 7, old displacement)."""
    out = {}
    for name in common.LELAS_CASES:
        m, Ag, Yg, Dg, Bf, Do, eq, dmn = common.lelas_case(name)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        rowPtr, colPtr = c.build_graph(0)
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf)
        if Do is not None:
            c.set_old_disp(Do)
        c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
    np.savez_compressed(os.path.join(HERE, "lelas.npz"), **out)
    print("wrote lelas.npz with", len(out), "arrays")


def prestress_golden():
    """struct_3d / l_elas_3d with a nodal prestress pS0, and the pSn / pSa accumulators of a prestress equation (pstEq)."""
    out = {}
    for name, *_ in common.PRESTRESS_CASES:
        m, Ag, Yg, Dg, Bf, pS0, eq, dmn = common.prestress_case(name)
        c = RefCase(); c.set_coords(m.x); c.add_mesh(m.IEN)
        rowPtr, colPtr = c.build_graph(0)
        c.set_prestress(pS0)
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        out[f"{name}/R"], out[f"{name}/Val"] = c.get_R(), c.get_Val()
        out[f"{name}/rowPtr"], out[f"{name}/colPtr"] = rowPtr, colPtr
        if eq.reserved & abi.EQ_PRESTRESS:
            out[f"{name}/pSn"], out[f"{name}/pSa"] = c.get_prestress()
        # the same case without pS0: the prestress must change the residual, or the fixture proves nothing
        c2 = RefCase(); c2.set_coords(m.x); c2.add_mesh(m.IEN); c2.build_graph(0)
        c2.alloc(3); c2.set_state(Ag, Yg, Dg, Bf); c2.assemble(0, eq, dmn)
        assert common.rel_err(c2.get_R(), out[f"{name}/R"]) > 1e-3, name
    np.savez_compressed(os.path.join(HERE, "prestress.npz"), **out)
    print("wrote prestress.npz with", len(out), "arrays")


if __name__ == "__main__":
    if not have_ref():
        raise SystemExit("oracle/_ref/libsvref.so is missing: run `make -C oracle ref` first")
    only = sys.argv[1:]
    for fn in (fluid_golden, struct_golden, fluid_gen_golden, fluid_hi_golden, struct_hi_golden, heat_golden, ustruct_golden, lelas_golden, prestress_golden, fsi_ustruct_golden, fluid_uris_golden, other_hi_golden, ris_golden, fluid_thood_golden):
        if not only or fn.__name__.replace("_golden", "") in only:
            fn()
