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
    e.set_mesh(0, m.IEN, w, N, Nx, eId=m.eId)
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
