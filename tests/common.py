"""Shared helpers of the parity tests: build the same synthetic case on the oracle and on the engine."""
import numpy as np

from svmultiphysics_b200 import abi, elements, meshgen


GOLDEN_N, GOLDEN_NZ = 3, 2      # mesh of the committed golden vectors (tests/golden/make_golden.py)

# (name, viscosity kwargs, K_darcy, body force, tDof, mvMsh)
FLUID_CASES = [
    ("newtonian", {}, 0.0, (0.0, 0.0, 0.0), 4, 0),
    ("darcy_bodyforce", {}, 3.0, (0.1, -0.2, 0.3), 4, 0),
    ("carreau_yasuda", dict(viscType=abi.VISC_CY, mu=0.035, mu_o=0.16, lam=8.2, a=0.64, n=0.2128), 0.0, (0.0, 0.0, 0.0), 4, 0),
    ("casson", dict(viscType=abi.VISC_CASSON, mu=0.3, mu_o=0.1, lam=0.5), 0.5, (0.0, 0.0, 1.0), 4, 0),
    ("moving_mesh", {}, 0.0, (0.0, 0.0, 0.0), 7, 1),
]

# (name, LS type, parameter overrides)
LS_CASES = [
    ("gmres", abi.LS_GMRES, dict(mItr=100, sD=50, relTol=1e-8)),
    ("gmres_restart", abi.LS_GMRES, dict(mItr=20, sD=10, relTol=1e-10)),
    ("bicgs", abi.LS_BICGS, dict(mItr=400, relTol=1e-8)),
]


def spmv_vector(nNo, dof=4):
    return np.asfortranarray(np.random.default_rng(3).standard_normal((dof, nNo)))


def load_golden(name="fluid_tet4.npz"):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def rel_err(a, b):
    """max |a-b| / max |b| : the FP64 parity measure used for assembled R / Val (tolerance 1e-12)."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def fluid_case(n=6, nz=8, tDof=4, seed=1234, with_bf=True):
    m = meshgen.cylinder_tet4(n, nz)
    Ag, Yg, Dg = meshgen.poiseuille_state(m, tDof=tDof, seed=seed)
    rng = np.random.default_rng(seed + 7)
    if tDof >= 7:
        Yg[4:7] = 0.3 * rng.standard_normal((3, m.nNo))
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo))) if with_bf else None
    return m, Ag, Yg, Dg, Bf


def make_oracle(cls, m, nFaces=0):
    c = cls()
    c.set_coords(m.x)
    c.add_mesh(m.IEN, eId=m.eId)
    rowPtr, colPtr = c.build_graph(nFaces)
    return c, rowPtr, colPtr


def make_engine(m, rowPtr, colPtr, device=0, **graph_kw):
    from svmultiphysics_b200.engine import Engine
    e = Engine(device)
    e.set_graph(rowPtr, colPtr, **graph_kw)
    w, N, Nx = elements.tables(m.eNoN)
    e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId, Nxx=elements.nxx_tables(m.eNoN) if m.eNoN == 8 else None)
    e.set_coords(m.x)
    return e


def dirichlet_faces(m):
    """Wall + inlet: all three velocity components constrained (val = 0), as fsi_ls_ini registers them
    (Code/Source/solver/baf_ini.cpp:749-773)."""
    faces = []
    for name in ("wall", "inlet"):
        g = m.faces[name]
        faces.append((abi.BC_DIR, g, np.zeros((3, len(g)), order="F")))
    return faces


# ---- solid / FSI cases -------------------------------------------------------------------------------
# (name, mesh factory, struct_domain kwargs, number of fibre families)
def _hex():
    return meshgen.box_hex8(4, 3, 3, (1e-3, 1e-3, 1e-3))


def _tet():
    return meshgen.box_tet4(3, 3, 2, (1.0, 1.0, 1.0))


# CANN parameter tables: rows (invariant, (kf0, kf1, kf2), (W0, W1, W2)).
CANN_HO = [(1, (1, 1, 2), (1.0, 8.023, 36.769)), (4, (1, 2, 2), (1.0, 16.026, 2881.6)), (8, (1, 2, 2), (1.0, 11.12, 557.7788)),
           (6, (1, 2, 2), (1.0, 11.436, 94.438))]
CANN_NHK = [(1, (1, 1, 1), (1.0, 1.0, 40.0942e4))]
CANN_ARTERY = [(1, (1, 1, 1), (1.0, 1.0, 4.0e4)), (1, (1, 2, 2), (1.0, 1.2, 3.0e3)), (4, (2, 2, 2), (1.0, 2.5, 6.0e4)),
               (8, (2, 2, 2), (1.0, 2.5, 6.0e4))]
CANN_ALL = [(1, (1, 1, 2), (1.0, 2.0, 5.0e4)), (2, (1, 2, 1), (0.7, 1.3, 4.0e4)), (3, (3, 1, 3), (0.5, 0.4, 2.0e4)),
            (4, (2, 2, 2), (1.0, 3.0, 3.0e4)), (5, (1, 1, 2), (1.0, 1.5, 1.0e4)), (6, (3, 2, 1), (1.0, 1.0, 2.0e4)),
            (7, (1, 2, 2), (0.8, 1.1, 1.5e4)), (8, (2, 1, 3), (0.6, 0.5, 2.5e4)), (9, (1, 2, 1), (1.0, 1.0, 1.0e4))]

STRUCT_CASES = [
    ("hex8_nHK_ST91", _hex, dict(), 0),                                         # struct/block_compression
    ("hex8_nHK_M94_damped", _hex, dict(volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), dmp=5.0, f=(0.1, 0.2, 0.3)), 0),
    ("tet4_nHK_Quad", _tet, dict(volType=abi.VOL_QUAD, E=1e6, nu=0.4, Kpen=1e6, rho=1.0), 0),
    ("hex8_MR", _hex, dict(isoType=abi.ISO_MR, C10=1e5, C01=3e4, Kpen=1e7, rho=1.0), 0),
    ("hex8_StVK", _hex, dict(isoType=abi.ISO_STVK, C10=2e5, C01=1e5, Kpen=0.0, rho=1.0), 0),
    # solid viscosity (mat_models.cpp:1583-1762): pseudo-potential and Newtonian models, element matrix no longer symmetric
    ("hex8_nHK_visc_potential", _hex, dict(solid_visc=abi.SOLID_VISC_POTENTIAL, solid_visc_mu=2.0e4, dmp=3.0), 0),
    ("tet4_nHK_visc_newtonian", _tet, dict(volType=abi.VOL_QUAD, E=1e6, nu=0.4, Kpen=1e6, rho=1.0, solid_visc=abi.SOLID_VISC_NEWTONIAN,
                                           solid_visc_mu=3.0e5), 0),
    ("hex8_MR_visc_newtonian", _hex, dict(isoType=abi.ISO_MR, C10=1e5, C01=3e4, Kpen=1e7, rho=1.0, solid_visc=abi.SOLID_VISC_NEWTONIAN,
                                          solid_visc_mu=50.0), 0),
    # fibre-reinforced models with the parameters of tests/cases/struct (LV_Holzapfel*, HGO arteries), cgs-like units
    ("hex8_HGO", _hex, dict(isoType=abi.ISO_HGO, C10=3.0e4, aff=2.4e4, bff=0.84, ass=2.4e4, bss=0.84, kap=0.226, Kpen=1e7, rho=1.0), 2),
    ("hex8_HO", _hex, dict(isoType=abi.ISO_HO, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12, afs=2160.0,
                           bfs=11.436, khs=100.0, Kpen=1e7, rho=1.0), 2),
    ("tet4_HO_ma", _tet, dict(isoType=abi.ISO_HO_MA, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                              afs=2160.0, bfs=11.436, khs=100.0, volType=abi.VOL_QUAD, Kpen=1e6, rho=1.0), 2),
    ("hex8_Guccione", _hex, dict(isoType=abi.ISO_GUCCIONE, C10=440.0, bff=8.0, bss=6.0, bfs=12.0, Kpen=1e6, rho=1e-3), 2),   # struct/LV_Guccione_passive
    # active stress along the fibre / sheet / sheet-normal directions (sv_struct.cpp:277-281, mat_models.cpp:321-340;
    # struct/LV_HolzapfelOgden_active, directionally_distributed_active_stress, tensile_adventitia_Guccione_active)
    ("hex8_nHK_active", _hex, dict(active_stress=True), 2),
    ("hex8_MR_active", _hex, dict(isoType=abi.ISO_MR, C10=1e5, C01=3e4, Kpen=1e7, rho=1.0, active_stress=True), 2),
    ("hex8_HGO_active", _hex, dict(isoType=abi.ISO_HGO, C10=3.0e4, aff=2.4e4, bff=0.84, ass=2.4e4, bss=0.84, kap=0.226, Kpen=1e7, rho=1.0,
                                   active_stress=True), 2),
    ("tet4_Guccione_active_fsn", _tet, dict(isoType=abi.ISO_GUCCIONE, C10=440.0, bff=8.0, bss=6.0, bfs=12.0, Kpen=1e6, rho=1e-3,
                                            volType=abi.VOL_QUAD, active_stress=True), 2),
    ("hex8_HO_active_fsn", _hex, dict(isoType=abi.ISO_HO, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                                      afs=2160.0, bfs=11.436, khs=100.0, Kpen=1e7, rho=1.0, active_stress=True), 2),
    ("tet4_HO_ma_active_fsn", _tet, dict(isoType=abi.ISO_HO_MA, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                                         afs=2160.0, bfs=11.436, khs=100.0, volType=abi.VOL_QUAD, Kpen=1e6, rho=1.0, active_stress=True), 2),
    # CANN: the parameter tables of struct/LV_HolzapfelOgden_passive_CANN, block_compression_CANN and
    # LV_CANN_artery_material_model (solver.xml <Add_row>), and a synthetic table touching every invariant and activation function
    ("hex8_CANN_HO", _hex, dict(cann=CANN_HO, Kpen=1e7, rho=1.0), 2),
    ("tet4_CANN_nHK", _tet, dict(cann=CANN_NHK, volType=abi.VOL_QUAD, Kpen=1e6, rho=1.0), 0),
    ("hex8_CANN_artery", _hex, dict(cann=CANN_ARTERY, Kpen=1e6, rho=1.0), 2),
    ("hex8_CANN_all_terms", _hex, dict(cann=CANN_ALL, Kpen=1e6, rho=1.0), 2),
    ("tet4_CANN_HO_active", _tet, dict(cann=CANN_HO, volType=abi.VOL_QUAD, Kpen=1e6, rho=1.0, active_stress=True), 2),
]


def active_tension(m, isoType, seed=31):
    """Nodal active tensions (cep_mod.cem.Ya_f, Ya_s, Ya_n): sheet / sheet-normal parts only for the models that accept them."""
    rng = np.random.default_rng(seed)
    scale = 2.0e4
    Ya_f = scale * (0.5 + rng.random(m.nNo))
    dirs = isoType in (abi.ISO_GUCCIONE, abi.ISO_HO, abi.ISO_HO_MA)
    Ya_s = 0.4 * scale * rng.random(m.nNo) if dirs else np.zeros(m.nNo)
    Ya_n = 0.2 * scale * rng.random(m.nNo) if dirs else np.zeros(m.nNo)
    return Ya_f, Ya_s, Ya_n


def struct_state(m, nFn=0, seed=11, tDof=3):
    rng = np.random.default_rng(seed)
    L = m.x.max()
    Dg = np.zeros((tDof, m.nNo), order="F")
    Dg[:3] = 0.01 * L * (m.x / L) * np.array([[1.0], [-0.5], [0.3]]) + 1e-3 * L * rng.standard_normal((3, m.nNo))
    Yg = np.asfortranarray(0.1 * rng.standard_normal((tDof, m.nNo)))
    Ag = np.asfortranarray(rng.standard_normal((tDof, m.nNo)))
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo)))
    fN = None
    if nFn:
        f = rng.standard_normal((3, m.nEl)); f /= np.linalg.norm(f, axis=0)
        s = rng.standard_normal((3, m.nEl)); s -= (s * f).sum(0) * f; s /= np.linalg.norm(s, axis=0)
        fN = np.asfortranarray(np.vstack([f, s]))
    return Ag, Yg, Dg, Bf, fN


def fsi_case(n=4, nz=6):
    """Pipe with a fluid core (domain Id 0) and a solid outer ring (domain Id 1) sharing interface nodes, like
    tests/cases/fsi/pipe_3d: tDof = 7 (u,v,w,p + mesh displacement/velocity in dofs 4..6)."""
    m = meshgen.cylinder_tet4(n, nz, R=1.0, L=3.0)
    c = m.x[:, m.IEN].mean(axis=1)                      # element centroids
    solid = (c[0] ** 2 + c[1] ** 2) > 0.55 ** 2
    m.eId = np.where(solid, 2, 1).astype(np.int32)      # bit 1 = solid, bit 0 = fluid
    rng = np.random.default_rng(21)
    Ag, Yg, _ = meshgen.poiseuille_state(m, R=1.0, U=5.0, tDof=7)
    Yg[4:7] = 0.2 * rng.standard_normal((3, m.nNo))
    Dg = np.zeros((7, m.nNo), order="F")
    Dg[:3] = 2e-3 * rng.standard_normal((3, m.nNo))
    Dg[4:7] = 5e-3 * rng.standard_normal((3, m.nNo))
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo)))
    return m, Ag, Yg, Dg, Bf


# ---- general (non-TET4) fluid elements: gnn + gn_nxx per Gauss point -----------------------------------------------
def _hex_skewed():
    """HEX8 box whose nodes are displaced pseudo-randomly: no element is a parallelepiped, so the second derivatives
    of the isoparametric map (xXi2 in nn::gn_nxx) do not vanish."""
    m = meshgen.box_hex8(3, 3, 2, (1.0, 1.2, 0.8))
    rng = np.random.default_rng(17)
    m.x = np.asfortranarray(m.x + 0.035 * rng.standard_normal(m.x.shape))
    return m


def _tet_cyl():
    return meshgen.cylinder_tet4(GOLDEN_N, GOLDEN_NZ)


# (name, mesh factory, viscosity kwargs, K_darcy, body force, tDof, mvMsh)
FLUID_GEN_CASES = [
    ("hex8_newtonian", _hex_skewed, {}, 0.0, (0.0, 0.0, 0.0), 4, 0),
    ("hex8_carreau_yasuda_darcy", _hex_skewed, dict(viscType=abi.VISC_CY, mu=0.035, mu_o=0.16, lam=8.2, a=0.64, n=0.2128), 2.0,
     (0.1, -0.2, 0.3), 4, 0),
    ("hex8_casson_moving_mesh", _hex_skewed, dict(viscType=abi.VISC_CASSON, mu=0.3, mu_o=0.1, lam=0.5), 0.0, (0.0, 0.0, 1.0), 7, 1),
    ("tet4_newtonian_general_path", _tet_cyl, {}, 0.5, (0.0, 0.1, 0.0), 4, 0),
]


# Quadratic and wedge elements of nn_elem_props.h through the same general path (nG != eNoN: TET10 15, HEX20 / HEX27 27 Gauss
# points; tests/cases/fluid/quadratic_tet10 is a TET10 VMS case).  Curved elements: the geometry's second derivatives enter gn_nxx.
def _tet10():
    return meshgen.elevate(meshgen.box_tet4(2, 2, 2, (1.0, 1.2, 0.8)), "tet10", bend=0.05)


def _hex20():
    return meshgen.elevate(meshgen.box_hex8(2, 2, 2, (1.0, 1.2, 0.8)), "hex20", bend=0.05)


def _hex27():
    return meshgen.elevate(meshgen.box_hex8(2, 2, 2, (1.0, 1.2, 0.8)), "hex27", bend=0.05)


def _wdg6():
    return meshgen.box_wdg6(2, 2, 2, (1.0, 1.2, 0.8), bend=0.05)


FLUID_HI_CASES = [
    ("tet10_newtonian", _tet10, {}, 0.0, (0.0, 0.0, 0.0), 4, 0),
    ("tet10_carreau_yasuda_moving_mesh", _tet10, dict(viscType=abi.VISC_CY, mu=0.035, mu_o=0.16, lam=8.2, a=0.64, n=0.2128), 2.0,
     (0.1, -0.2, 0.3), 7, 1),
    ("hex20_newtonian", _hex20, {}, 0.5, (0.0, 0.1, 0.0), 4, 0),
    ("hex27_casson", _hex27, dict(viscType=abi.VISC_CASSON, mu=0.3, mu_o=0.1, lam=0.5), 0.0, (0.0, 0.0, 1.0), 4, 0),
    ("wdg6_newtonian", _wdg6, {}, 0.0, (0.2, 0.0, 0.0), 4, 0),
]


# displacement-based solid on the same curved quadratic / wedge elements (struct_3d through the general kernel with nG != eNoN)
STRUCT_HI_CASES = [
    ("tet10_nHK_ST91", _tet10, dict(E=1e6, nu=0.4, Kpen=1e6, rho=1.0, dmp=2.0, f=(0.1, 0.2, 0.3)), 0),
    ("hex20_MR", _hex20, dict(isoType=abi.ISO_MR, C10=1e5, C01=3e4, Kpen=1e6, rho=1.0), 0),
    ("hex27_HO_active_fsn", _hex27, dict(isoType=abi.ISO_HO, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                                        afs=2160.0, bfs=11.436, khs=100.0, Kpen=1e6, rho=1.0, active_stress=True), 2),
    ("wdg6_nHK_M94", _wdg6, dict(volType=abi.VOL_M94, E=1e6, nu=0.3, Kpen=1e6, rho=1.0), 0),
    ("tet10_CANN_HO", _tet10, dict(cann=CANN_HO, Kpen=1e6, rho=1.0), 2),
]


def fluid_gen_state(m, tDof, seed=31):
    rng = np.random.default_rng(seed)
    Yg = np.zeros((tDof, m.nNo), order="F")
    Yg[0] = 1.0 + 0.5 * np.sin(2.0 * m.x[1]) + 0.1 * rng.standard_normal(m.nNo)
    Yg[1] = 0.3 * np.cos(1.5 * m.x[0] + m.x[2]) + 0.1 * rng.standard_normal(m.nNo)
    Yg[2] = 0.7 * m.x[0] * m.x[1] + 0.1 * rng.standard_normal(m.nNo)
    Yg[3] = 2.0 - m.x[2] + 0.05 * rng.standard_normal(m.nNo)
    if tDof >= 7:
        Yg[4:7] = 0.2 * rng.standard_normal((3, m.nNo))
    Ag = np.asfortranarray(0.5 * rng.standard_normal((tDof, m.nNo)))
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo)))
    return Ag, Yg, None, Bf


# ---- URIS valves (unfitted resistive immersed surfaces; tests/cases/uris/pipe_uris_cfd, pipe_uris_fsi) ------------------------------
# (name, mesh factory, tDof, mvMsh, FSI?)
def _tet_pipe():
    return meshgen.cylinder_tet4(4, 8, R=1.0, L=4.0)


def _hex_skewed_fine():
    m = meshgen.box_hex8(5, 5, 6, (1.0, 1.2, 1.5))
    rng = np.random.default_rng(19)
    m.x = np.asfortranarray(m.x + 0.02 * rng.standard_normal(m.x.shape))
    return m


URIS_CASES = [
    ("tet4_two_valves", _tet_pipe, 4, 0, False),
    ("hex8_two_valves_ale", _hex_skewed_fine, 7, 1, False),
    ("tet4_fsi_pipe", None, 7, 1, True),
]


def uris_valves(m, seed=41):
    """Two valves on mesh m: (raw urisType members for the reference harness, abi.Uris list for the device, sdf, scaffold_udf,
    valve_vel).  Valve 0: a tilted plane through the middle of the mesh, closing (ramped thickness), with its velocity; valve 1: a
    plane further downstream, fully open, with a cylindrical scaffold.  The thicknesses span a few elements, so that Gauss points
    inside, at the edge of and outside the smeared surfaces all occur."""
    rng = np.random.default_rng(seed)
    lo, hi = m.x.min(axis=1), m.x.max(axis=1)
    L = hi - lo
    h = 0.1 * L[2]                                       # thickness unit: about one element layer along the axis
    hs = 0.08 * min(L[0], L[1])                          # scaffold thickness unit
    ctr = 0.5 * (lo + hi)
    n0 = np.array([0.2, -0.1, 1.0]); n0 /= np.linalg.norm(n0)
    sdf0 = n0 @ (m.x - (ctr - 0.15 * L * np.array([0, 0, 1.0]))[:, None])
    sdf1 = m.x[2] - (ctr[2] + 0.25 * L[2])
    r = np.hypot(m.x[0] - ctr[0], m.x[1] - ctr[1])
    udf1 = np.abs(r - 0.3 * min(L[0], L[1]))
    sdf = np.ascontiguousarray(np.stack([sdf0, sdf1]))
    udf = np.ascontiguousarray(np.stack([np.zeros(m.nNo), udf1]))
    vel = np.zeros((2, m.nNo, 3))
    vel[0] = 0.5 * np.stack([np.sin(m.x[1]), np.cos(m.x[0]), 1.0 + 0.2 * m.x[2]], axis=1) + 0.05 * rng.standard_normal((m.nNo, 3))
    raw = [dict(resistance=3.0e3, sdf_deps=1.2 * h, sdf_deps_close=2.4 * h, clsFlg=True, cnt=3, n_open=8, n_close=10, scaffold=False,
                include_velocity=True),
           dict(resistance=1.5e3, sdf_deps=1.6 * h, sdf_deps_close=2.0 * hs, clsFlg=False, cnt=50, n_open=8, n_close=10, scaffold=True,
                include_velocity=False)]
    dev = [abi.Uris(resistance=u["resistance"],
                    sdf_deps=abi.uris_effective_deps(u["sdf_deps"], u["sdf_deps_close"], u["clsFlg"], u["cnt"], u["n_open"], u["n_close"]),
                    scaffold_deps=u["sdf_deps_close"], scaffold=int(u["scaffold"]), include_velocity=int(u["include_velocity"]))
           for u in raw]
    return raw, dev, sdf, udf, vel


def uris_case(name, scatter=abi.SCATTER_ATOMIC):
    """(mesh, Ag, Yg, Dg, Bf, eq, domains)"""
    _, mk, tDof, mv, fsi = next(c for c in URIS_CASES if c[0] == name)
    if fsi:
        m, Ag, Yg, Dg, Bf = fsi_case()
        af, am, gam, beta = abi.gen_alpha(0.5)
        eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                          scatter=scatter, reserved=0)
        dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0),
               abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
        return m, Ag, Yg, Dg, Bf, eq, dmn
    m = mk()
    Ag, Yg, Dg, Bf = fluid_gen_state(m, tDof)
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv, scatter=scatter)
    return m, Ag, Yg, Dg, Bf, eq, [abi.fluid_domain(K_darcy=1.5, f=(0.1, -0.2, 0.3))]


# ---- fitted RIS: two lumen meshes separated by a resistive surface with node-to-node twins (tests/cases/ris/pipe_ris_3d) ------------
def ris_case(scatter=abi.SCATTER_ATOMIC, seed=43):
    """(x, [IEN upstream, IEN downstream], map(2, n), Ag, Yg, Bf, eq, domains): the cylinder cut at a cell layer; the nodes of the cut
    plane exist twice (the upstream copy keeps its id, the downstream copy is appended), map(0, j) / map(1, j) are the twins."""
    m = meshgen.cylinder_tet4(4, 8, R=1.0, L=4.0)
    zc = m.x[2, m.IEN].mean(axis=0)
    zcut = np.unique(np.round(m.x[2], 9))[4]                  # a node plane in the middle
    up = zc < zcut
    plane = np.where(np.abs(m.x[2] - zcut) < 1e-9)[0]
    twin = np.full(m.nNo, -1, np.int64)
    twin[plane] = m.nNo + np.arange(len(plane))
    x = np.asfortranarray(np.hstack([m.x, m.x[:, plane]]))
    IEN0 = np.asfortranarray(m.IEN[:, up].astype(np.int32))
    I1 = m.IEN[:, ~up].astype(np.int64)
    I1 = np.where(twin[I1] >= 0, twin[I1], I1)
    IEN1 = np.asfortranarray(I1.astype(np.int32))
    mp = np.asfortranarray(np.stack([plane, twin[plane]]).astype(np.int32))
    nNo = x.shape[1]
    rng = np.random.default_rng(seed)
    Yg = np.zeros((4, nNo), order="F")
    r2 = x[0] ** 2 + x[1] ** 2
    Yg[2] = 5.0 * (1.0 - r2) + 0.2 * rng.standard_normal(nNo)
    Yg[0:2] = 0.2 * rng.standard_normal((2, nNo))
    Yg[3] = 10.0 - x[2] + 0.1 * rng.standard_normal(nNo)
    Ag = np.asfortranarray(0.5 * rng.standard_normal((4, nNo)))
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, nNo)))
    eq = abi.fluid_eq(0.005, scatter=scatter)
    return x, [IEN0, IEN1], mp, Ag, Yg, Bf, eq, [abi.fluid_domain(K_darcy=0.5, f=(0.1, 0.0, -0.2))]


# ---- Taylor-Hood fluid (mshType::nFs = 2: P2-P1 / Q2-Q1, vmsStab = false; fluid.cpp:494-500, fs.cpp:73-178) --------------------------
# (name, mesh factory, viscosity kwargs, K_darcy, body force, tDof, mvMsh)
FLUID_THOOD_CASES = [
    ("tet10_newtonian", _tet10, {}, 0.0, (0.0, 0.0, 0.0), 4, 0),
    ("tet10_carreau_yasuda_darcy_moving_mesh", _tet10, dict(viscType=abi.VISC_CY, mu=0.035, mu_o=0.16, lam=8.2, a=0.64, n=0.2128), 2.0,
     (0.1, -0.2, 0.3), 7, 1),
    ("hex27_casson", _hex27, dict(viscType=abi.VISC_CASSON, mu=0.3, mu_o=0.1, lam=0.5), 0.5, (0.0, 0.0, 1.0), 4, 0),
    ("hex20_newtonian", _hex20, {}, 0.0, (0.0, 0.1, 0.0), 4, 0),
]


# Taylor-Hood + two URIS valves (the momentum loop sees the valve factor at the velocity rule's Gauss points)
FLUID_THOOD_URIS_CASE = ("tet10_uris", _tet10, {}, 0.5, (0.1, 0.0, -0.2), 4, 0)


def fsi_thood_case(scatter=abi.SCATTER_ATOMIC, seed=53):
    """(mesh, Ag, Yg, Dg, Bf, eq, domains): the FSI pipe of fsi_case elevated to curved TET10 elements — Taylor-Hood fluid core on the
    moved (ALE) geometry, struct_3d wall on the velocity space (fsi.cpp:84-88, 170-262)."""
    m0, *_ = fsi_case(n=3, nz=4)
    m = meshgen.elevate(m0, "tet10", bend=0.01)
    m.eId = m0.eId
    rng = np.random.default_rng(seed)
    Ag, Yg, _, Bf = fluid_gen_state(m, 7)
    Dg = np.zeros((7, m.nNo), order="F")
    Dg[:3] = 2e-3 * rng.standard_normal((3, m.nNo))
    Dg[4:7] = 3e-3 * rng.standard_normal((3, m.nNo))
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=0,
                      scatter=scatter, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0),
           abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
    return m, Ag, Yg, Dg, Bf, eq, dmn


def fluid_thood_eq(dt, tDof=4, mvMsh=0, scatter=abi.SCATTER_ATOMIC):
    eq = abi.fluid_eq(dt, tDof=tDof, mvMsh=mvMsh, scatter=scatter)
    eq.vmsStab = 0
    return eq


# ---- scalar heat equations (heatS / heatF) -------------------------------------------------------
# (name, mesh factory, fluid?, tDof, eq.s, mvMsh, heat_domain kwargs)
HEAT_CASES = [
    ("heats_tet4", _tet, False, 1, 0, 0, dict(conductivity=0.7, source=0.3, rho=2.5)),
    ("heats_hex8", _hex_skewed, False, 3, 2, 0, dict(conductivity=12.0, source=-1.5, rho=0.8)),
    ("heatf_tet4", _tet_cyl, True, 5, 4, 0, dict(conductivity=0.01, source=0.2)),
    ("heatf_hex8_moving_mesh", _hex_skewed, True, 8, 7, 1, dict(conductivity=0.05, source=0.0)),
]


def heat_state(m, tDof, s, seed=47):
    """Velocity-like fields in dofs 0..2 (and 4..6), a smooth temperature + noise in dof s."""
    rng = np.random.default_rng(seed)
    Yg = np.asfortranarray(0.3 * rng.standard_normal((tDof, m.nNo)))
    if tDof >= 4:
        Yg[0] += 1.0 + 0.5 * np.sin(2.0 * m.x[1])
        Yg[2] += 0.7 * m.x[0]
    Yg[s] = 300.0 + 20.0 * np.sin(m.x[0] + 0.5 * m.x[2]) + 5.0 * m.x[1] + 0.5 * rng.standard_normal(m.nNo)
    Ag = np.asfortranarray(rng.standard_normal((tDof, m.nNo)))
    Dg = np.zeros((tDof, m.nNo), order="F")
    Bf = np.zeros((3, m.nNo), order="F")
    return Ag, Yg, Dg, Bf


# ---- heat / linear elasticity / mesh motion on curved TET10 / HEX20 / HEX27 / WDG elements (nG != eNoN; lShpF wedge behaviour) -------
# (name, kind, mesh factory, parameters).  kind: heats | heatf | lelas | mesh
OTHER_HI_CASES = [
    ("tet10_heats", "heats", _tet10, dict(tDof=1, s=0, mv=0, dkw=dict(conductivity=0.7, source=0.3, rho=2.5))),
    ("hex27_heatf_moving_mesh", "heatf", _hex27, dict(tDof=8, s=7, mv=1, dkw=dict(conductivity=0.05, source=0.1))),
    ("hex20_heatf", "heatf", _hex20, dict(tDof=5, s=4, mv=0, dkw=dict(conductivity=0.01, source=0.2))),
    ("wdg6_heats", "heats", _wdg6, dict(tDof=3, s=2, mv=0, dkw=dict(conductivity=12.0, source=-1.5, rho=0.8))),
    ("tet10_lelas", "lelas", _tet10, dict(dkw=dict(E=1.0e6, nu=0.3, rho=2.0, f=(0.1, -0.2, 0.3)))),
    ("hex27_lelas", "lelas", _hex27, dict(dkw=dict(E=2.0e6, nu=0.25, rho=1.0))),
    ("wdg6_lelas", "lelas", _wdg6, dict(dkw=dict(E=1.0e6, nu=0.35, rho=1.5, f=(0.0, 0.0, -1.0)))),
    ("hex20_mesh", "mesh", _hex20, dict(dkw=dict(E=1.0, nu=0.3))),
    ("tet10_mesh", "mesh", _tet10, dict(dkw=dict(E=1.0, nu=0.3))),
]


def other_hi_case(name, scatter=abi.SCATTER_ATOMIC):
    """(mesh, element type, dof, Ag, Yg, Dg, Bf, Do or None, eq, domains)"""
    _, kind, mk, kw = next(c for c in OTHER_HI_CASES if c[0] == name)
    m = mk()
    et = name.split("_")[0]
    if kind in ("heats", "heatf"):
        Ag, Yg, Dg, Bf = heat_state(m, kw["tDof"], kw["s"])
        eq = abi.heat_eq(0.01, kind == "heatf", tDof=kw["tDof"], s=kw["s"], mvMsh=kw["mv"], scatter=scatter)
        return m, et, 1, Ag, Yg, Dg, Bf, None, eq, [abi.heat_domain(kind == "heatf", **kw["dkw"])]
    if kind == "lelas":
        Ag, Yg, Dg, Bf, _ = struct_state(m, 0)
        return m, et, 3, Ag, Yg, Dg, Bf, None, abi.lelas_eq(1e-3, scatter=scatter), [abi.lelas_domain(**kw["dkw"])]
    Ag, Yg, Dg, Bf, _ = struct_state(m, 0, tDof=7)
    Dg[4:7] = Dg[0:3]; Ag[4:7] *= 0.5
    Do = np.asfortranarray(0.9 * Dg)
    return m, et, 3, Ag, Yg, Dg, Bf, Do, abi.mesh_eq(1e-3, scatter=scatter), [abi.mesh_domain(**kw["dkw"])]


# ---- mixed velocity-pressure solid (ustruct) ------------------------------------------------------
# (name, mesh factory, ustruct_domain kwargs, number of fibre families)
USTRUCT_CASES = [
    ("tet4_nHK_ST91", _tet, dict(E=1.0e6, nu=0.45, Kpen=1.0e6 / (3 * (1 - 0.9)), rho=1.2, ctau_M=1e-3, ctau_C=1e-3, f=(0.1, -0.2, 0.3)), 0),  # ustruct/block_compression
    ("hex8_nHK_M94", _hex_skewed, dict(volType=abi.VOL_M94, E=2.0e5, nu=0.3, Kpen=1.0e6, rho=1.0, ctau_M=1e-2, ctau_C=1e-2), 0),
    ("hex8_MR_Quad_incompressible", _hex_skewed, dict(isoType=abi.ISO_MR, volType=abi.VOL_QUAD, C10=1e5, C01=3e4, E=8e5, nu=0.5, Kpen=5.0e6,
                                                      rho=1.0, ctau_M=1e-3, ctau_C=1e-5), 0),
    ("tet4_HO_ma_fibres", _tet, dict(isoType=abi.ISO_HO_MA, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                                     afs=2160.0, bfs=11.436, khs=100.0, E=1.0e5, nu=0.483333, Kpen=1e6, rho=1.0, ctau_M=1e-5, ctau_C=1e-5), 2),
    ("hex8_nHK_no_penalty", _hex_skewed, dict(E=1.0e6, nu=0.4, Kpen=0.0, rho=2.0, ctau_M=1e-3, ctau_C=1e-3), 0),
    # solid viscosity (ustruct/tensile_adventitia_{Newtonian,Potential}_viscosity)
    ("hex8_nHK_visc_potential", _hex_skewed, dict(E=1.0e6, nu=0.45, Kpen=2.0e6, rho=1.0, ctau_M=1e-3, ctau_C=1e-3,
                                                   solid_visc=abi.SOLID_VISC_POTENTIAL, solid_visc_mu=2.0e4), 0),
    ("tet4_nHK_visc_newtonian", _tet, dict(E=1.0e6, nu=0.45, Kpen=2.0e6, rho=1.0, ctau_M=1e-3, ctau_C=1e-3,
                                           solid_visc=abi.SOLID_VISC_NEWTONIAN, solid_visc_mu=3.0e4), 0),
    # active stress (ustruct/LV_Guccione_active, LV_HolzapfelOgden_active) and CANN
    ("tet4_Guccione_active_fsn", _tet, dict(isoType=abi.ISO_GUCCIONE, C10=440.0, bff=8.0, bss=6.0, bfs=12.0, E=1.0e5, nu=0.45, Kpen=1e6,
                                            rho=1.0, ctau_M=1e-4, ctau_C=1e-4, active_stress=True), 2),
    ("hex8_HO_active_fsn", _hex_skewed, dict(isoType=abi.ISO_HO, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                                             afs=2160.0, bfs=11.436, khs=100.0, E=1.0e5, nu=0.483333, Kpen=1e6, rho=1.0, ctau_M=1e-5,
                                             ctau_C=1e-5, active_stress=True), 2),
    ("tet4_CANN_HO", _tet, dict(cann=CANN_HO, E=1.0e5, nu=0.483333, Kpen=1e6, rho=1.0, ctau_M=1e-5, ctau_C=1e-5), 2),
]


def ustruct_state(m, nFn=0, seed=23):
    """tDof = 4: velocity + pressure in Yg, their rates in Ag, displacement in Dg(0..2)."""
    rng = np.random.default_rng(seed)
    L = m.x.max()
    Dg = np.zeros((4, m.nNo), order="F")
    Dg[:3] = 0.02 * L * (m.x / L) * np.array([[1.0], [-0.5], [0.3]]) + 2e-3 * L * rng.standard_normal((3, m.nNo))
    Yg = np.asfortranarray(0.1 * rng.standard_normal((4, m.nNo)))
    Yg[3] = 1.0e3 * (1.0 + 0.3 * rng.standard_normal(m.nNo))
    Ag = np.asfortranarray(rng.standard_normal((4, m.nNo)))
    Ag[3] *= 50.0
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo)))
    fN = None
    if nFn:
        f = rng.standard_normal((3, m.nEl)); f /= np.linalg.norm(f, axis=0)
        s = rng.standard_normal((3, m.nEl)); s -= (s * f).sum(0) * f; s /= np.linalg.norm(s, axis=0)
        fN = np.asfortranarray(np.vstack([f, s]))
    return Ag, Yg, Dg, Bf, fN


def ustruct_Ad(m, seed=29):
    """com_mod.Ad(3, tnNo): time derivative of the displacement (Integrator.cpp:412, ustruct_r)."""
    return np.asfortranarray(0.05 * np.random.default_rng(seed).standard_normal((3, m.nNo)))


# ---- FSI with velocity-pressure (ustruct) solids: construct_fsi's ustruct_3d_m/c branch (fsi.cpp:243-262, tests/cases/fsi_ustruct) ----
FSI_USTRUCT_CASES = {
    "nHK_ST91": (dict(E=1.0e6, nu=0.45, Kpen=1.0e6 / (3 * (1 - 0.9)), rho=1.2, ctau_M=1e-3, ctau_C=1e-3, f=(0.1, -0.2, 0.3)), 0),
    "HO_ma_fibres_visc": (dict(isoType=abi.ISO_HO_MA, st_a=590.0, st_b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12,
                               afs=2160.0, bfs=11.436, khs=100.0, E=1.0e5, nu=0.483333, Kpen=1e6, rho=1.0, ctau_M=1e-5, ctau_C=1e-5,
                               solid_visc=abi.SOLID_VISC_NEWTONIAN, solid_visc_mu=3.0e4), 2),
}


def fsi_ustruct_case(name, scatter=abi.SCATTER_ATOMIC):
    """(mesh, Ag, Yg, Dg, Bf, fN, nFn, eq, domains, Ad, ustruct-node flags): the pipe of fsi_case with a ustruct wall; the solid
    nodes carry a pressure of the size the ustruct cases use."""
    dkw, nFn = FSI_USTRUCT_CASES[name]
    m, Ag, Yg, Dg, Bf = fsi_case()
    rng = np.random.default_rng(37)
    solid_nodes = np.unique(m.IEN[:, (m.eId & 2) != 0])
    flags = np.zeros(m.nNo, np.int32); flags[solid_nodes] = 1
    Yg[3, solid_nodes] = 1.0e3 * (1.0 + 0.3 * rng.standard_normal(len(solid_nodes)))
    fN = None
    if nFn:
        f = rng.standard_normal((3, m.nEl)); f /= np.linalg.norm(f, axis=0)
        t = rng.standard_normal((3, m.nEl)); t -= (t * f).sum(0) * f; t /= np.linalg.norm(t, axis=0)
        fN = np.asfortranarray(np.vstack([f, t]))
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=scatter, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0), abi.ustruct_domain(Id=1, **dkw)]
    return m, Ag, Yg, Dg, Bf, fN, nFn, eq, dmn, ustruct_Ad(m), flags


# ---- l_elas_3d on TET4: linear-elasticity equation and mesh-motion equation ------------------------------------------
LELAS_CASES = ["lelas_tet4", "mesh_tet4"]


def lelas_case(name):
    """(mesh, Ag, Yg, Dg, Bf, Do or None, eq, domains)"""
    if name == "lelas_tet4":
        m = _tet()
        Ag, Yg, Dg, Bf, _ = struct_state(m, 0)
        return m, Ag, Yg, Dg, Bf, None, abi.lelas_eq(1e-3), [abi.lelas_domain(E=1.0e6, nu=0.3, rho=2.0, f=(0.1, -0.2, 0.3))]
    m, Ag, Yg, Dg, Bf = fsi_case()
    m.eId = None
    Do = np.asfortranarray(0.9 * Dg)
    return m, Ag, Yg, Dg, Bf, Do, abi.mesh_eq(1e-3), [abi.mesh_domain(E=1.0, nu=0.3)]


# ---- prestress (com_mod.pS0, pstEq: sv_struct.cpp:271-274, 635-680, 327-336; l_elas.cpp:321-338, 130-140) -----------------------
# (name, mesh, physics, domain kwargs, pstEq)
PRESTRESS_CASES = [
    ("hex8_struct_pS0", _hex, "struct", dict(), False),
    ("hex8_struct_pstEq", _hex, "struct", dict(isoType=abi.ISO_MR, C10=1e5, C01=3e4, Kpen=1e7, rho=1.0), True),
    ("tet4_struct_pstEq_visc", _tet, "struct", dict(volType=abi.VOL_QUAD, E=1e6, nu=0.4, Kpen=1e6, rho=1.0,
                                                   solid_visc=abi.SOLID_VISC_POTENTIAL, solid_visc_mu=2.0e4), True),
    ("tet4_struct_pstEq", _tet, "struct", dict(volType=abi.VOL_QUAD, E=1e6, nu=0.4, Kpen=1e6, rho=1.0), True),
    ("tet4_lelas_pstEq", _tet, "lelas", dict(E=1.0e6, nu=0.3, rho=2.0, f=(0.1, -0.2, 0.3)), True),
    ("hex8_lelas_pS0", _hex, "lelas", dict(E=2.0e6, nu=0.25, rho=1.0), False),
]


def prestress_case(name):
    """(mesh, Ag, Yg, Dg, Bf, pS0, eq, domains) of a PRESTRESS_CASES entry."""
    _, mk, phys, dkw, pst = next(c for c in PRESTRESS_CASES if c[0] == name)
    m = mk()
    Ag, Yg, Dg, Bf, _ = struct_state(m, 0)
    pS0 = np.asfortranarray(2.0e6 * np.random.default_rng(41).standard_normal((6, m.nNo)))
    if phys == "struct":
        eq, dmn = abi.struct_eq(1e-4), [abi.struct_domain(**dkw)]
    else:
        eq, dmn = abi.lelas_eq(1e-3), [abi.lelas_domain(**dkw)]
    if pst:
        eq.reserved |= abi.EQ_PRESTRESS
    return m, Ag, Yg, Dg, Bf, pS0, eq, dmn


# ---- multi-rank reference runs (tests/test_multirank_reference_cpu.py, tests/mrank_ref_worker.py) -------------------------
def mrank_faces(m, ls_name):
    """Global face list and resistances: Dirichlet wall + inlet; for "ns" also the coupled resistance outlet of pipe_RCR_3d."""
    faces = dirichlet_faces(m)
    res = np.zeros(len(faces))
    if ls_name == "ns":
        out = m.faces["outlet"]
        val = np.zeros((3, len(out)), order="F"); val[2] = 4.0 * np.pi / len(out)
        faces.append((abi.BC_NEU, out, val))
        res = np.append(res, 0.8)
    return faces, res


def mrank_ls(ls_name):
    if ls_name == "ns":
        return abi.LS_NS, abi.ls_params(abi.LS_NS, mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17),
                                        cg=(300, 0, 1e-3, 1e-17))
    return abi.LS_GMRES, abi.ls_params(abi.LS_GMRES, mItr=10, sD=80, relTol=1e-9)


def mrank_struct_case():
    """HEX8 nHK/ST91 block with the symmetric Dirichlet planes of struct/block_compression; Dirichlet values are not partial sums."""
    m = meshgen.box_hex8(5, 4, 6, (1e-3, 1e-3, 1e-3))
    Ag, Yg, Dg, Bf, _ = struct_state(m, 0)
    faces = []
    for k, name in enumerate(("X0", "Y0", "Z0")):
        val = np.ones((3, len(m.faces[name])), order="F"); val[k] = 0.0
        faces.append((abi.BC_DIR, m.faces[name], val))
    return m, Ag, Yg, Dg, Bf, faces, abi.struct_eq(1e-4), [abi.struct_domain()], abi.ls_params(abi.LS_BICGS, mItr=600, relTol=1e-10)
