"""ctypes mirror of include/svb200.h (structs and enums only; no library loading here)."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 6

# svb200_phys
PHYS_FLUID, PHYS_STRUCT, PHYS_FSI, PHYS_MESH, PHYS_LELAS, PHYS_HEATS, PHYS_HEATF, PHYS_USTRUCT = 0, 1, 2, 3, 4, 5, 6, 7
# svb200_visc
VISC_CONST, VISC_CY, VISC_CASSON = 0, 1, 2
# svb200_iso / svb200_vol
ISO_NHK, ISO_MR, ISO_GUCCIONE, ISO_STVK, ISO_HGO, ISO_HO, ISO_HO_MA, ISO_CANN = 0, 1, 2, 3, 4, 5, 6, 7
CANN_MAX_ROWS = 16
VOL_NONE, VOL_QUAD, VOL_ST91, VOL_M94 = 0, 1, 2, 3
# svb200_ls_type
LS_NS, LS_GMRES, LS_CG, LS_BICGS = 0, 1, 2, 3
PREC_FSILS = 0
PREC_RCS = 1
SOLID_VISC_NONE, SOLID_VISC_NEWTONIAN, SOLID_VISC_POTENTIAL = 0, 1, 2
BC_DIR, BC_NEU = 0, 1
SCATTER_ATOMIC, SCATTER_COLORED = 0, 1
ARRAY_R, ARRAY_VAL, ARRAY_W, ARRAY_KD, ARRAY_RD = 0, 1, 2, 3, 4


class EqParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("af", C.c_double), ("am", C.c_double), ("gam", C.c_double), ("beta", C.c_double),
        ("phys", C.c_int32), ("dof", C.c_int32), ("tDof", C.c_int32), ("s", C.c_int32),
        ("mvMsh", C.c_int32), ("vmsStab", C.c_int32), ("scatter", C.c_int32), ("reserved", C.c_int32),
    ]


class EqTime(C.Structure):
    """svb200_eqtime: rows [s,e] and generalised-alpha coefficients of one equation."""
    _fields_ = [("s", C.c_int32), ("e", C.c_int32), ("phys", C.c_int32), ("reserved", C.c_int32),
                ("af", C.c_double), ("am", C.c_double), ("gam", C.c_double), ("beta", C.c_double)]


SOL_OLD, SOL_CURRENT, SOL_INTERMEDIATE = 0, 1, 2

MAX_URIS = 4


class Uris(C.Structure):
    """svb200_uris: one unfitted-RIS valve (com_mod.uris[i]) as the element kernels need it."""
    _fields_ = [("resistance", C.c_double), ("sdf_deps", C.c_double), ("scaffold_deps", C.c_double),
                ("scaffold", C.c_int32), ("include_velocity", C.c_int32)]


def uris_effective_deps(sdf_deps: float, sdf_deps_close: float, clsFlg: bool, cnt: int, n_open: int, n_close: int) -> float:
    """Half-thickness of a valve surface at this step: the linear ramp between the open and the closed value over the DxOpen /
    DxClose steps (Code/Source/solver/uris.cpp:1625-1649).  Scalar host logic: the device gets the result (svb200_uris.sdf_deps)."""
    start, end, n = (sdf_deps, sdf_deps_close, n_close) if clsFlg else (sdf_deps_close, sdf_deps, n_open)
    if n <= 0 or cnt >= n:
        return end
    if cnt <= 0:
        return start
    return start + (float(cnt) / float(n)) * (end - start)


class DmnParams(C.Structure):
    _fields_ = [
        ("Id", C.c_int32), ("phys", C.c_int32),
        ("rho", C.c_double),
        ("f", C.c_double * 3),
        ("K_darcy", C.c_double),
        ("viscType", C.c_int32), ("isoType", C.c_int32),
        ("mu_i", C.c_double), ("mu_o", C.c_double), ("lam", C.c_double), ("a", C.c_double), ("n", C.c_double),
        ("volType", C.c_int32), ("solidViscType", C.c_int32),
        ("Kpen", C.c_double),
        ("C10", C.c_double), ("C01", C.c_double),
        ("bff", C.c_double), ("bss", C.c_double), ("bfs", C.c_double),
        ("dmp", C.c_double),
        ("E", C.c_double), ("nu", C.c_double),
        ("solid_visc_mu", C.c_double),
        ("backflow_stab", C.c_double),
        ("st_a", C.c_double), ("st_b", C.c_double), ("aff", C.c_double), ("ass", C.c_double), ("afs", C.c_double),
        ("kap", C.c_double), ("khs", C.c_double),
        ("conductivity", C.c_double), ("source_term", C.c_double),
        ("ctau_M", C.c_double), ("ctau_C", C.c_double),
        ("active_stress", C.c_int32), ("cann_rows", C.c_int32),
        ("cann_inv", C.c_int32 * CANN_MAX_ROWS),
        ("cann_act", (C.c_int32 * 3) * CANN_MAX_ROWS),
        ("cann_w", (C.c_double * 3) * CANN_MAX_ROWS),
    ]


class SubLsParams(C.Structure):
    _fields_ = [("mItr", C.c_int32), ("sD", C.c_int32), ("relTol", C.c_double), ("absTol", C.c_double)]


class LsParams(C.Structure):
    _fields_ = [("RI", SubLsParams), ("GM", SubLsParams), ("CG", SubLsParams)]


class SubLsResult(C.Structure):
    _fields_ = [("success", C.c_int32), ("itr", C.c_int32),
                ("iNorm", C.c_double), ("fNorm", C.c_double), ("dB", C.c_double), ("callD", C.c_double)]


class LsResult(C.Structure):
    _fields_ = [("RI", SubLsResult), ("GM", SubLsResult), ("CG", SubLsResult),
                ("Resm", C.c_int32), ("Resc", C.c_int32),
                ("hist_n", C.c_int32), ("hist_cap", C.c_int32),
                ("hist", C.POINTER(C.c_double))]


def gen_alpha(rho_inf: float):
    """Generalised-alpha coefficients from the spectral radius (Code/Source/solver/initialize.cpp:484-486)."""
    am = 0.5 * (3.0 - rho_inf) / (1.0 + rho_inf)
    af = 1.0 / (1.0 + rho_inf)
    gam = 0.5 + am - af
    beta = 0.25 * (1.0 + am - af) ** 2
    return af, am, gam, beta


EQTIME_SSTEQ = 1     # com_mod.sstEq for an FSI equation (its solids are ustruct domains)


def eq_time(s: int, e: int, phys: int, rho_inf: float = 0.5, sstEq: bool = False) -> EqTime:
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqTime(s=s, e=e, phys=phys, reserved=EQTIME_SSTEQ if sstEq else 0, af=af, am=am, gam=gam, beta=beta)


EQ_GENERAL_KERNEL = 1
EQ_PRESTRESS = 2     # com_mod.pstEq


def fluid_eq(dt: float, rho_inf: float = 0.5, tDof: int = 4, scatter: int = SCATTER_ATOMIC, mvMsh: int = 0,
             general: bool = False) -> EqParams:
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqParams(dt=dt, af=af, am=am, gam=gam, beta=beta, phys=PHYS_FLUID, dof=4, tDof=tDof, s=0,
                    mvMsh=mvMsh, vmsStab=1, scatter=scatter, reserved=EQ_GENERAL_KERNEL if general else 0)


def fluid_domain(rho: float = 1.06, mu: float = 0.04, f=(0.0, 0.0, 0.0), K_darcy: float = 0.0,
                 viscType: int = VISC_CONST, mu_o: float = 0.0, lam: float = 0.0, a: float = 0.0, n: float = 0.0,
                 Id: int = -1, backflow_stab: float = 0.0) -> DmnParams:
    d = DmnParams()
    d.Id = Id
    d.phys = PHYS_FLUID
    d.backflow_stab = backflow_stab
    d.rho = rho
    d.f[0], d.f[1], d.f[2] = f
    d.K_darcy = K_darcy
    d.viscType = viscType
    d.mu_i, d.mu_o, d.lam, d.a, d.n = mu, mu_o, lam, a, n
    return d


def struct_eq(dt: float, rho_inf: float = 0.5, tDof: int = 3, dof: int = 3, s: int = 0, scatter: int = SCATTER_ATOMIC) -> EqParams:
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqParams(dt=dt, af=af, am=am, gam=gam, beta=beta, phys=PHYS_STRUCT, dof=dof, tDof=tDof, s=s,
                    mvMsh=0, vmsStab=1, scatter=scatter, reserved=0)


def struct_domain(rho: float = 1000.0, isoType: int = ISO_NHK, volType: int = VOL_ST91, E: float = 240.56596e6, nu: float = 0.5,
                  Kpen: float = 4.0e9, C10=None, C01: float = 0.0, bff: float = 0.0, bss: float = 0.0, bfs: float = 0.0,
                  dmp: float = 0.0, f=(0.0, 0.0, 0.0), Id: int = -1, solid_visc: int = 0, solid_visc_mu: float = 0.0,
                  st_a: float = 0.0, st_b: float = 0.0, aff: float = 0.0, ass: float = 0.0, afs: float = 0.0, kap: float = 0.0,
                  khs: float = 100.0, active_stress: bool = False, cann=None) -> DmnParams:
    """Solid domain; C10 defaults to mu/2 with mu = E/(2(1+nu)) as set_material_props does for nHK
    (Code/Source/solver/set_material_props.h).  cann: rows (invariant, (kf0, kf1, kf2), (W0, W1, W2)) of the CANN parameter
    table (the <Add_row> entries of a Constitutive_model type="CANN", ArtificialNeuralNetMaterial.h); sets isoType."""
    d = DmnParams()
    d.Id = Id
    d.phys = PHYS_STRUCT
    d.rho = rho
    d.f[0], d.f[1], d.f[2] = f
    d.isoType, d.volType = isoType, volType
    d.Kpen = Kpen
    mu = E / (2.0 * (1.0 + nu))
    d.C10 = 0.5 * mu if C10 is None else C10
    d.C01 = C01
    d.bff, d.bss, d.bfs = bff, bss, bfs
    d.dmp = dmp
    d.E, d.nu = E, nu
    d.solidViscType, d.solid_visc_mu = solid_visc, solid_visc_mu
    d.st_a, d.st_b, d.aff, d.ass, d.afs, d.kap, d.khs = st_a, st_b, aff, ass, afs, kap, khs
    d.active_stress = 1 if active_stress else 0
    if cann is not None:
        assert 1 <= len(cann) <= CANN_MAX_ROWS
        d.isoType = ISO_CANN
        d.cann_rows = len(cann)
        for r, (inv, act, w) in enumerate(cann):
            d.cann_inv[r] = inv
            for k in range(3):
                d.cann_act[r][k] = act[k]
                d.cann_w[r][k] = w[k]
    return d


def mesh_eq(dt: float, rho_inf: float = 0.5, tDof: int = 7, s: int = 4, scatter: int = SCATTER_ATOMIC) -> EqParams:
    """Mesh-motion equation of an FSI run: dof = 3, state dofs nsd+1..2nsd (Code/Source/solver/mesh.cpp:47-48)."""
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqParams(dt=dt, af=af, am=am, gam=gam, beta=beta, phys=PHYS_MESH, dof=3, tDof=tDof, s=s,
                    mvMsh=1, vmsStab=1, scatter=scatter, reserved=0)


def mesh_domain(E: float = 1.0, nu: float = 0.3, rho: float = 0.0, f=(0.0, 0.0, 0.0), Id: int = -1) -> DmnParams:
    d = DmnParams()
    d.Id = Id
    d.phys = PHYS_MESH
    d.rho = rho
    d.f[0], d.f[1], d.f[2] = f
    d.E, d.nu = E, nu
    return d


def lelas_eq(dt: float, rho_inf: float = 0.5, tDof: int = 3, scatter: int = SCATTER_ATOMIC) -> EqParams:
    """Linear-elasticity equation (tests/cases/linear-elasticity): dof = 3, displacement-based."""
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqParams(dt=dt, af=af, am=am, gam=gam, beta=beta, phys=PHYS_LELAS, dof=3, tDof=tDof, s=0,
                    mvMsh=0, vmsStab=1, scatter=scatter, reserved=0)


def lelas_domain(E: float = 1.0e6, nu: float = 0.3, rho: float = 1.0, f=(0.0, 0.0, 0.0), Id: int = -1) -> DmnParams:
    d = mesh_domain(E=E, nu=nu, rho=rho, f=f, Id=Id)
    d.phys = PHYS_LELAS
    return d


def heat_eq(dt: float, fluid: bool, rho_inf: float = 0.5, tDof: int = 1, s: int = 0, mvMsh: int = 0,
            scatter: int = SCATTER_ATOMIC) -> EqParams:
    """Heat equation in a solid (heatS) or advection-diffusion in a fluid (heatF): dof = 1, temperature in state dof s
    (Code/Source/solver/heats.cpp:196, heatf.cpp:253)."""
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqParams(dt=dt, af=af, am=am, gam=gam, beta=beta, phys=PHYS_HEATF if fluid else PHYS_HEATS, dof=1, tDof=tDof,
                    s=s, mvMsh=mvMsh, vmsStab=1, scatter=scatter, reserved=0)


def heat_domain(fluid: bool, conductivity: float = 1.0, source: float = 0.0, rho: float = 1.0, Id: int = -1) -> DmnParams:
    d = DmnParams()
    d.Id = Id
    d.phys = PHYS_HEATF if fluid else PHYS_HEATS
    d.rho = rho
    d.conductivity, d.source_term = conductivity, source
    return d


def ustruct_eq(dt: float, rho_inf: float = 0.5, tDof: int = 4, scatter: int = SCATTER_ATOMIC) -> EqParams:
    """Mixed velocity-pressure solid (ustruct): dof = 4 = (v, p), equal-order VMS (Code/Source/solver/ustruct.cpp:203-400)."""
    af, am, gam, beta = gen_alpha(rho_inf)
    return EqParams(dt=dt, af=af, am=am, gam=gam, beta=beta, phys=PHYS_USTRUCT, dof=4, tDof=tDof, s=0,
                    mvMsh=0, vmsStab=1, scatter=scatter, reserved=0)


def ustruct_domain(isoType: int = ISO_NHK, volType: int = VOL_ST91, E: float = 1.0e6, nu: float = 0.45, rho: float = 1.0,
                   ctau_M: float = 1.0e-3, ctau_C: float = 1.0e-3, **kw) -> DmnParams:
    """Like struct_domain, with the VMS constants ctau_M / ctau_C and E / nu kept for compute_tau
    (Code/Source/solver/mat_models.cpp:1470-1493)."""
    d = struct_domain(isoType=isoType, volType=volType, E=E, nu=nu, rho=rho, **kw)
    d.phys = PHYS_USTRUCT
    d.E, d.nu = E, nu
    d.ctau_M, d.ctau_C = ctau_M, ctau_C
    return d


def ls_params(ls_type: int, mItr=None, sD=None, relTol=None, absTol=1e-10, gm=None, cg=None) -> LsParams:
    """Defaults of fsils_ls_create (Code/Source/linear_solver/ls.cpp:22-59), overridable like read_ls does."""
    p = LsParams()
    if ls_type == LS_NS:
        p.RI = SubLsParams(10, 100, 0.4, 1e-10)
        p.GM = SubLsParams(2, 100, 1e-2, 1e-10)
        p.CG = SubLsParams(500, 0, 0.2, 1e-10)
    elif ls_type == LS_GMRES:
        p.RI = SubLsParams(1000, 250, 0.1, 1e-10)
    else:
        p.RI = SubLsParams(1000, 250, 1e-2, 1e-10)
    if mItr is not None:
        p.RI.mItr = mItr
    if sD is not None:
        p.RI.sD = sD
    if relTol is not None:
        p.RI.relTol = relTol
    p.RI.absTol = absTol
    if gm:
        p.GM = SubLsParams(*gm)
    if cg:
        p.CG = SubLsParams(*cg)
    return p
