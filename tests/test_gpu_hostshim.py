"""The reference's own host objects driving the GPU through the C++ plug-in layer.

`svmultiphysics_b200/host/B200LinearAlgebra.cpp` implements the reference's `class LinearAlgebra`
(Code/Source/solver/LinearAlgebra.h:13-37) over the C ABI and carries the early-out of
`eq_assem::global_eq_assem` (Code/Source/solver/eq_assem.cpp:397).  Here the compiled reference fills its ComMod /
eqType / mshType / FSILS_lhsType exactly as for the CPU runs and then executes

    ls_alloc -> global_eq_assem -> [b_assem_neu_bc on the host -> LinearAlgebra::assemble] -> ls_solve

twice: with FsilsLinearAlgebra (CPU, the oracle) and with B200LinearAlgebra (device).  Same bytes in, the
assembled R / Val must agree to 1e-12 and the solver results to the set tolerance.
"""
import numpy as np
import pytest

from svmultiphysics_b200 import abi, meshgen
from tests import common

pytestmark = pytest.mark.gpu


def _pair(m, nFaces=0, **mesh_kw):
    from oracle import refbind
    if not refbind.have_host():
        pytest.skip("needs oracle/_ref/libsvref.so and svmultiphysics_b200/lib/libsvb200_host.so (built where the reference tree is present)")
    out = []
    for gpu in (False, True):
        c = refbind.RefCase()
        c.set_coords(m.x)
        c.add_mesh(m.IEN, **mesh_kw)
        c.build_graph(nFaces)
        if gpu:
            c.use_b200_backend(device=0)
        out.append(c)
    return out


def _close(*cs):
    for c in cs:
        c.close()


@pytest.mark.parametrize("ls_type,kw", [
    (abi.LS_GMRES, dict(mItr=100, sD=50, relTol=1e-8)),
    (abi.LS_NS, dict(mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17))),
], ids=["gmres", "ns_resistance"])
def test_fluid_newton_iteration_through_cpp_plugin(ls_type, kw):
    m, Ag, Yg, Dg, Bf = common.fluid_case()
    faces = common.dirichlet_faces(m)
    out = m.faces["outlet"]
    val = np.zeros((3, len(out)), order="F"); val[2] = 4.0 * np.pi / len(out)
    faces.append((abi.BC_NEU, out, val))
    cpu, gpu = _pair(m, nFaces=len(faces), eId=m.eId)
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain(backflow_stab=0.2)]
    # a Neumann face integrated on the HOST by the reference's b_assem_neu_bc: reaches the device through the
    # per-element LinearAlgebra::assemble() of the plug-in
    IENb, gE = meshgen.boundary_face_elements(m, m.faces["outlet_all"])
    hg = np.zeros(m.nNo); hg[m.faces["outlet_all"]] = -120.0
    res = []
    for c in (cpu, gpu):
        for i, (g, nodes, v) in enumerate(faces):
            c.set_face(i, g, nodes, v)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        iFa = c.add_face(0, IENb, gE)
        c.assemble_neu(0, iFa, eq, dmn, hg)
        R, V = c.get_R(), c.get_Val()
        ls = abi.ls_params(ls_type, **kw)
        X, o, _ = c.solve(4, ls_type, ls, np.ones(len(faces), np.int32), np.array([0.0, 0.0, 0.8]))
        res.append((R, V, X, o))
    (R0, V0, X0, o0), (R1, V1, X1, o1) = res
    assert gpu.backend_launch_count() > 10, "the plug-in did not launch device kernels"
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(V1, V0) < 1e-12
    assert o1.RI.success == o0.RI.success
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    if ls_type == abi.LS_GMRES and o0.RI.itr > ls.RI.sD + 1:
        # dozens of restarts (same contract as tests/test_gpu_fluid.py): the stopping test may fire a few steps apart
        assert abs(o1.RI.itr - o0.RI.itr) <= max(2, o0.RI.itr // 25)
        assert o1.RI.fNorm <= ls.RI.relTol * o1.RI.iNorm
    else:
        assert o1.RI.itr == o0.RI.itr
        assert abs(o1.RI.fNorm - o0.RI.fNorm) <= 2e-2 * o0.RI.fNorm
    assert common.rel_err(X1, X0) < (50 * ls.RI.relTol if ls_type == abi.LS_NS else 1e-6)
    _close(cpu, gpu)


@pytest.mark.parametrize("case", ["hex8_nHK_ST91", "hex8_Guccione", "hex8_HO_active_fsn", "hex8_CANN_HO"])
def test_struct_newton_iteration_through_cpp_plugin(case):
    """The last two cases also check that the plug-in hands cep_mod.cem.Ya_f / Ya_s / Ya_n and the CANN parameter table of
    stM.paramTable to the device (B200LinearAlgebra.cpp: domain_params, set_active_tension)."""
    name, mk, dkw, nFn = next(c for c in common.STRUCT_CASES if c[0] == case)
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.struct_state(m, nFn=nFn)
    faces = []
    for k, fname in enumerate(("X0", "Y0", "Z0")):          # symmetric Dirichlet planes, one direction each
        val = np.ones((3, len(m.faces[fname])), order="F"); val[k] = 0.0
        faces.append((abi.BC_DIR, m.faces[fname], val))
    cpu, gpu = _pair(m, nFaces=3, nFn=nFn, fN=fN)
    eq, dmn = abi.struct_eq(1e-4), [abi.struct_domain(**dkw)]
    ls = abi.ls_params(abi.LS_GMRES, mItr=50, sD=60, relTol=1e-10)
    res = []
    for c in (cpu, gpu):
        for i, (g, nodes, v) in enumerate(faces):
            c.set_face(i, g, nodes, v)
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf)
        if dmn[0].active_stress:
            c.set_active_tension(*common.active_tension(m, dmn[0].isoType))      # fills cep_mod.cem of the reference
        c.assemble(0, eq, dmn)
        R, V = c.get_R(), c.get_Val()
        X, o, _ = c.solve(3, abi.LS_GMRES, ls, np.ones(3, np.int32), np.zeros(3))
        res.append((R, V, X, o))
    (R0, V0, X0, o0), (R1, V1, X1, o1) = res
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(V1, V0) < 1e-12
    assert o1.RI.success == o0.RI.success
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert abs(o1.RI.itr - o0.RI.itr) <= max(3, o0.RI.itr // 20)
    if o0.RI.success:
        assert common.rel_err(X1, X0) < 1e-6
    _close(cpu, gpu)


def test_plugin_error_behaviour():
    """Failures surface as std::runtime_error like everywhere on the reference's path (the harness turns them into a
    status + message): element assembly before ls_alloc is a call-order error, not a silent host fallback."""
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=3, nz=3)
    cpu, gpu = _pair(m, nFaces=0, eId=m.eId)
    gpu.set_state(Ag, Yg, Dg, Bf)
    with pytest.raises(RuntimeError, match="before ls_alloc"):
        gpu.assemble(0, abi.fluid_eq(0.005), [abi.fluid_domain()])
    _close(cpu, gpu)


def test_hex8_fluid_through_cpp_plugin():
    """HEX8 fluid: the plug-in hands fs[0].Nxx of the reference's own mshType to svb200_set_mesh_nxx; the general-element
    kernel must reproduce construct_fluid (gnn + gn_nxx per Gauss point) to 1e-12, and the RCS preconditioner selected
    through eq.linear_algebra_preconditioner must reach the device."""
    name, mk, visc, Kd, f, tDof, mv = common.FLUID_GEN_CASES[1]
    m = mk()
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("X0", "Y0", "Y1", "Z0", "Z1")]
    cpu, gpu = _pair(m, nFaces=len(faces))
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    ls = abi.ls_params(abi.LS_GMRES, mItr=10, sD=100, relTol=1e-6)
    res = []
    for c in (cpu, gpu):
        for i, (g, nodes, v) in enumerate(faces):
            c.set_face(i, g, nodes, v)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        R, V = c.get_R(), c.get_Val()
        X, o, _ = c.solve(4, abi.LS_GMRES, ls, np.ones(len(faces), np.int32), np.zeros(len(faces)), prec=abi.PREC_RCS)
        res.append((R, V, X, o))
    (R0, V0, X0, o0), (R1, V1, X1, o1) = res
    assert gpu.backend_launch_count() > 10
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(V1, V0) < 1e-12
    assert o1.RI.success == o0.RI.success and abs(o1.RI.itr - o0.RI.itr) <= 2
    assert abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 2e-5
    _close(cpu, gpu)


@pytest.mark.parametrize("case", [1, 2], ids=["heatS_hex8", "heatF_tet4"])
def test_heat_through_cpp_plugin(case):
    """heatS / heatF (dof = 1): conductivity, source term and solid density travel from dmnType.prop through
    b200::domain_params; assembly to 1e-12 and a BiCGStab solve against the reference's own run."""
    name, mk, fluid, tDof, s, mv, dkw = common.HEAT_CASES[case]
    m = mk()
    Ag, Yg, Dg, Bf = common.heat_state(m, tDof, s)
    fname = "X0" if "X0" in m.faces else "inlet"
    faces = [(abi.BC_DIR, m.faces[fname], np.zeros((1, len(m.faces[fname])), order="F"))]
    cpu, gpu = _pair(m, nFaces=1)
    eq, dmn = abi.heat_eq(0.01, fluid, tDof=tDof, s=s, mvMsh=mv), [abi.heat_domain(fluid, **dkw)]
    ls = abi.ls_params(abi.LS_BICGS, mItr=400, relTol=1e-10)
    res = []
    for c in (cpu, gpu):
        for i, (g, nodes, v) in enumerate(faces):
            c.set_face(i, g, nodes, v)
        c.alloc(1); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        R, V = c.get_R(), c.get_Val()
        X, o, _ = c.solve(1, abi.LS_BICGS, ls, np.ones(1, np.int32), np.zeros(1))
        res.append((R, V, X, o))
    (R0, V0, X0, o0), (R1, V1, X1, o1) = res
    assert gpu.backend_launch_count() > 0
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(V1, V0) < 1e-12
    assert o1.RI.success == o0.RI.success and abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-6
    _close(cpu, gpu)


def test_ustruct_through_cpp_plugin():
    """ustruct: stM + ctau_M / ctau_C / E / nu through b200::domain_params, Kd stays on the device, and the Integrator::step patch
    calls B200LinearAlgebra::ustruct_r where the reference calls ustruct::ustruct_r."""
    name, mk, dkw, nFn = common.USTRUCT_CASES[1]
    m = mk()
    Ag, Yg, Dg, Bf, fN = common.ustruct_state(m, nFn)
    faces = []
    for k, fname in enumerate(("X0", "Y0", "Z0")):
        val = np.ones((4, len(m.faces[fname])), order="F"); val[k] = 0.0
        faces.append((abi.BC_DIR, m.faces[fname], val))
    cpu, gpu = _pair(m, nFaces=3)
    eq, dmn = abi.ustruct_eq(1e-3), [abi.ustruct_domain(**dkw)]
    Ad = common.ustruct_Ad(m)
    ls = abi.ls_params(abi.LS_GMRES, mItr=4, sD=200, relTol=1e-6)
    res = []
    for c in (cpu, gpu):
        for i, (g, nodes, v) in enumerate(faces):
            c.set_face(i, g, nodes, v)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        R, V, K = c.get_R(), c.get_Val(), c.get_Kd()
        c.ustruct_r(1, Ad)
        R2 = c.get_R()
        X, o, _ = c.solve(4, abi.LS_GMRES, ls, np.ones(3, np.int32), np.zeros(3))
        res.append((R, V, K, R2, X, o))
    (R0, V0, K0, R20, X0, o0), (R1, V1, K1, R21, X1, o1) = res
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(K1, K0) < 1e-12 and common.rel_err(R21, R20) < 1e-12
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14], [15]):
        assert common.rel_err(V1[rows], V0[rows]) < 1e-12
    assert o1.RI.success == o0.RI.success and abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-4
    _close(cpu, gpu)


def test_fsi_ustruct_through_cpp_plugin():
    """FSI with a ustruct wall (tests/cases/fsi_ustruct): b200::global_eq_assem hands construct_fsi's fluid + ustruct domains to the
    device, and B200LinearAlgebra::ustruct_r restricts the displacement-residual update to the nodes all_fun::is_domain puts in the
    ustruct domain (com_mod.dmnId)."""
    name = "HO_ma_fibres_visc"
    m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, Ad, flags = common.fsi_ustruct_case(name)
    faces = common.dirichlet_faces(m)
    cpu, gpu = _pair(m, nFaces=len(faces), eId=m.eId, nFn=nFn, fN=fN)
    ls = abi.ls_params(abi.LS_GMRES, mItr=4, sD=150, relTol=1e-6)
    res = []
    for c in (cpu, gpu):
        for i, (g, nodes, v) in enumerate(faces):
            c.set_face(i, g, nodes, v)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        R, V, K = c.get_R(), c.get_Val(), c.get_Kd()
        c.ustruct_r(1, Ad)
        R2 = c.get_R()
        X, o, _ = c.solve(4, abi.LS_GMRES, ls, np.ones(len(faces), np.int32), np.zeros(len(faces)))
        res.append((R, V, K, R2, X, o))
    (R0, V0, K0, R20, X0, o0), (R1, V1, K1, R21, X1, o1) = res
    assert common.rel_err(K1, K0) < 1e-12
    solid, fluid_only = np.where(flags != 0)[0], np.where(flags == 0)[0]
    for nodes in (solid, fluid_only):
        assert common.rel_err(R1[:, nodes], R0[:, nodes]) < 1e-12 and common.rel_err(R21[:, nodes], R20[:, nodes]) < 1e-12
    assert np.array_equal(R21[:, fluid_only], R1[:, fluid_only]) and not np.array_equal(R21[:, solid], R1[:, solid])
    assert common.rel_err(V1, V0) < 1e-12
    assert o1.RI.success == o0.RI.success and abs(o1.RI.iNorm - o0.RI.iNorm) <= 1e-10 * o0.RI.iNorm
    assert common.rel_err(X1, X0) < 1e-4
    _close(cpu, gpu)


def test_uris_valves_through_cpp_plugin():
    """com_mod.uris[] through B200LinearAlgebra::set_uris (the open/close thickness ramp is computed by the plug-in from clsFlg / cnt /
    DxClose like uris.cpp:1625-1649) against the reference's own construct_fluid with the same valves."""
    name = "tet4_two_valves"
    m, Ag, Yg, Dg, Bf, eq, dmn = common.uris_case(name)
    raw, dev, sdf, udf, vel = common.uris_valves(m)
    cpu, gpu = _pair(m)
    res = []
    for c in (cpu, gpu):
        c.set_uris(raw, sdf, udf, vel)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        res.append((c.get_R(), c.get_Val()))
    golden = common.load_golden("fluid_uris.npz")
    assert np.array_equal(res[0][0], golden[f"{name}/R"])
    assert common.rel_err(res[1][0], res[0][0]) < 1e-12 and common.rel_err(res[1][1], res[0][1]) < 1e-12
    _close(cpu, gpu)


def test_open_ris_surface_through_cpp_plugin():
    """com_mod.risFlag / RIS.clsFlg / grisMapList through B200LinearAlgebra::set_ris: the reference's construct_fluid + doassem_ris on
    the CPU against the device's row operation, surface open and then closed (the plug-in rebuilds the plan when clsFlg changes)."""
    from oracle import refbind
    if not refbind.have_host():
        pytest.skip("needs oracle/_ref/libsvref.so and svmultiphysics_b200/lib/libsvb200_host.so")
    x, IENs, mp, Ag, Yg, Bf, eq, dmn = common.ris_case()
    golden = common.load_golden("ris.npz")
    gpu = refbind.RefCase(); gpu.set_coords(x)
    for I in IENs:
        gpu.add_mesh(I)
    gpu.set_ris([mp], [0], [(0, 1)])
    gpu.build_graph(0)
    gpu.use_b200_backend(device=0)
    for label, closed in (("open", [0]), ("closed", [1]), ("open", [0])):
        gpu.set_ris([mp], closed, [(0, 1)])
        gpu.alloc(4); gpu.set_state(Ag, Yg, None, Bf)
        for iM in range(len(IENs)):
            gpu.assemble(iM, eq, dmn)
        assert common.rel_err(gpu.get_R(), golden[f"{label}/R"]) < 1e-12
        assert common.rel_err(gpu.get_Val(), golden[f"{label}/Val"]) < 1e-12
    gpu.close()


@pytest.mark.parametrize("name", ["lelas_tet4", "mesh_tet4", "hex27_lelas", "hex20_mesh"])
def test_linear_elasticity_and_mesh_motion_through_cpp_plugin(name):
    """phys_lElas and phys_mesh through b200::global_eq_assem: the plug-in passes solutions.old's displacement for the mesh equation
    (mesh.cpp:60-75) and any mshType's tables (here also curved HEX27 / HEX20 elements)."""
    if name in common.LELAS_CASES:
        m, Ag, Yg, Dg, Bf, Do, eq, dmn = common.lelas_case(name)
    else:
        m, et, dof, Ag, Yg, Dg, Bf, Do, eq, dmn = common.other_hi_case(name)
    cpu, gpu = _pair(m)
    res = []
    for c in (cpu, gpu):
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf)
        if Do is not None:
            c.set_old_disp(Do)
        c.assemble(0, eq, dmn)
        res.append((c.get_R(), c.get_Val()))
    assert np.abs(res[0][0]).max() > 0
    assert common.rel_err(res[1][0], res[0][0]) < 1e-12 and common.rel_err(res[1][1], res[0][1]) < 1e-12
    _close(cpu, gpu)


@pytest.mark.parametrize("name", ["tet10_carreau_yasuda_darcy_moving_mesh", "hex27_casson"])
def test_taylor_hood_fluid_through_cpp_plugin(name):
    """A Taylor-Hood mesh (mshType::nFs = 2) through the plug-in: upload_structure hands the tables of fs::get_thood_fs to
    svb200_set_mesh_thood, eq_params sets vmsStab = 0, and thood_val_rc runs on the device where Integrator::step calls fs::thood_val_rc."""
    from oracle import refbind
    if not refbind.have_host():
        pytest.skip("needs oracle/_ref/libsvref.so and svmultiphysics_b200/lib/libsvb200_host.so")
    _, mk, visc, Kd, f, tDof, mv = next(c for c in common.FLUID_THOOD_CASES if c[0] == name)
    golden = common.load_golden("fluid_thood.npz")
    m = mk()
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    gpu = refbind.RefCase(); gpu.set_coords(m.x); gpu.add_mesh(m.IEN); gpu.set_mesh_thood(0)
    gpu.build_graph(0)
    gpu.use_b200_backend(device=0)
    gpu.alloc(4); gpu.set_state(Ag, Yg, Dg, Bf)
    gpu.assemble(0, common.fluid_thood_eq(0.005, tDof=tDof, mvMsh=mv), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
    R1, V1 = gpu.get_R(), gpu.get_Val()
    GR, GV = golden[f"{name}/R"], golden[f"{name}/Val"]
    assert common.rel_err(R1[:3], GR[:3]) < 1e-12 and common.rel_err(R1[3], GR[3]) < 1e-12
    for rows in ([0, 1, 2, 4, 5, 6, 8, 9, 10], [3, 7, 11], [12, 13, 14]):
        assert common.rel_err(V1[rows], GV[rows]) < 1e-12
    gpu.thood_val_rc()
    assert np.array_equal(gpu.get_Val()[15], golden[f"{name}/Val_rc"][15])
    assert common.rel_err(gpu.get_R()[3], golden[f"{name}/R_rc"][3]) < 1e-12
    gpu.close()


def test_prestress_equation_through_cpp_plugin():
    """com_mod.pS0 and pstEq through B200LinearAlgebra: the plug-in uploads pS0, flags the prestress equation and writes the
    device accumulators back into com_mod.pSn / pSa (what Integrator::corrector then communicates and divides)."""
    m, Ag, Yg, Dg, Bf, pS0, eq, dmn = common.prestress_case("hex8_struct_pstEq")
    cpu, gpu = _pair(m, nFaces=0)
    res = []
    for c in (cpu, gpu):
        c.set_prestress(pS0)
        c.alloc(3); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        res.append((c.get_R(), c.get_Val()) + c.get_prestress())
    (R0, V0, pSn0, pSa0), (R1, V1, pSn1, pSa1) = res
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(V1, V0) < 1e-12
    assert common.rel_err(pSn1, pSn0) < 1e-12 and common.rel_err(pSa1, pSa0) < 1e-12
    _close(cpu, gpu)


@pytest.mark.parametrize("case", ["tet10_newtonian", "hex27_casson", "wdg6_newtonian"])
def test_quadratic_fluid_through_cpp_plugin(case):
    """TET10 / HEX27 / WDG fluid meshes of the reference's own mshType (w, N, Nx, fs[0].Nxx, nG != eNoN) through the plug-in."""
    name, mk, visc, Kd, f, tDof, mv = next(c for c in common.FLUID_HI_CASES if c[0] == case)
    m = mk()
    Ag, Yg, Dg, Bf = common.fluid_gen_state(m, tDof)
    cpu, gpu = _pair(m, nFaces=0)
    eq, dmn = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    res = []
    for c in (cpu, gpu):
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf); c.assemble(0, eq, dmn)
        res.append((c.get_R(), c.get_Val()))
    (R0, V0), (R1, V1) = res
    assert gpu.backend_launch_count() > 3
    assert common.rel_err(R1, R0) < 1e-12 and common.rel_err(V1, V0) < 1e-12
    _close(cpu, gpu)
