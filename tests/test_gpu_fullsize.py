"""Parity at BASELINE.json's full size (config C2: 10,025,280 tet4 on one B200).

The oracle cannot assemble 10 M elements in seconds, so the full-size run is checked through
  * z sub-slabs of the SAME mesh and state assembled by the oracle (the compiled reference when it is
    present): every row whose elements all lie inside the sub-slab must agree to 1e-12,
  * the two independent scatter paths (grouped REDs vs coloured read-modify-write) agreeing to 1e-12,
  * SpMV against scipy's BSR product of the downloaded matrix and linearity of the device SpMV,
  * the GMRES answer: the TRUE preconditioned residual ||W(R - K x)|| recomputed on the host from the
    unscaled system is below relTol * ||W R||, and equals the solver's reported fNorm.
"""
import numpy as np
import pytest

from oracle import refbind
from svmultiphysics_b200 import abi, elements, meshgen
from tests import common

pytestmark = pytest.mark.gpu

N, NZ = 118, 120            # bench.py's default C2 lattice: 6*118^2*120 tets
DT = 1e-3


def _state(m):
    import bench
    return bench.lattice_state(m, 0, NZ)


@pytest.fixture(scope="module")
def full():
    from svmultiphysics_b200.engine import Engine
    m, _, _, _ = meshgen.cylinder_slab(N, NZ, 0, 1)
    assert m.nEl >= 10_000_000
    Ag, Yg = _state(m)
    eng = Engine(0)
    rowPtr, colPtr = eng.lhsa(m.nNo, [m.IEN])
    eng.set_graph(rowPtr, colPtr)
    w, Nt, Nx = elements.tables(4)
    eng.set_mesh(0, m.IEN, w, Nt, Nx)
    eng.set_coords(m.x)
    wall = m.faces["wall"]
    eng.set_num_faces(1)
    eng.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))
    eng.alloc(4)
    eng.set_state(Ag, Yg)
    eng.assemble(0, abi.fluid_eq(DT), [abi.fluid_domain()])
    d = dict(m=m, Ag=Ag, Yg=Yg, eng=eng, rowPtr=rowPtr, colPtr=colPtr, R=eng.get_R(), Val=eng.get_Val(), wall=wall)
    yield d
    eng.close()


def _subslab_rows(full, k0, k1):
    """Oracle assembly of cell layers [k0,k1) of the full mesh; returns (err_R, err_Val, rows compared)."""
    m = full["m"]
    n1 = N + 1
    lo, hi = n1 * n1 * k0, n1 * n1 * (k1 + 1)
    layer = m.IEN.min(axis=0) // (n1 * n1)
    els = np.nonzero((layer >= k0) & (layer < k1))[0]
    assert len(els) == 6 * N * N * (k1 - k0)
    IEN = np.asfortranarray(m.IEN[:, els] - lo)
    assert IEN.min() >= 0 and IEN.max() < hi - lo
    cls = refbind.RefCase if refbind.have_ref() else refbind.OracleCase
    c = cls()
    c.set_coords(np.asfortranarray(m.x[:, lo:hi]))
    c.add_mesh(IEN)
    rp, cp = c.build_graph(0)
    c.alloc(4)
    c.set_state(np.asfortranarray(full["Ag"][:, lo:hi]), np.asfortranarray(full["Yg"][:, lo:hi]))
    c.assemble(0, abi.fluid_eq(DT), [abi.fluid_domain()])
    R0, V0 = c.get_R(), c.get_Val()
    # rows strictly inside the sub-slab see all of their elements (plane k0 too when it is the inlet plane)
    r0 = 0 if k0 == 0 else n1 * n1
    r1 = (hi - lo) - n1 * n1
    g0, g1 = r0 + lo, r1 + lo
    RP, CP = full["rowPtr"], full["colPtr"]
    assert np.array_equal(np.diff(rp[r0:r1 + 1]), np.diff(RP[g0:g1 + 1]))
    s0, s1, t0, t1 = rp[r0], rp[r1], RP[g0], RP[g1]
    assert np.array_equal(cp[s0:s1] + lo, CP[t0:t1])
    eR = common.rel_err(full["R"][:, g0:g1], R0[:, r0:r1])
    eV = common.rel_err(full["Val"][:, t0:t1], V0[:, s0:s1])
    return eR, eV, r1 - r0


@pytest.mark.parametrize("k0,k1", [(0, 3), (58, 62), (NZ - 4, NZ)], ids=["inlet", "middle", "outlet_side"])
def test_fullsize_assembly_matches_oracle_on_subslab(full, k0, k1):
    eR, eV, rows = _subslab_rows(full, k0, k1)
    assert rows > 2 * (N + 1) ** 2 - 1
    assert eR < 1e-12 and eV < 1e-12, (eR, eV)


def test_fullsize_atomic_equals_colored(full):
    eng = full["eng"]
    eng.alloc(4)
    eng.assemble(0, abi.fluid_eq(DT, scatter=abi.SCATTER_COLORED), [abi.fluid_domain()])
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, full["R"]) < 1e-12
    assert common.rel_err(V1, full["Val"]) < 1e-12
    # conservation check on the assembled residual: sum over nodes of the continuity rows = sum over elements of
    # the element integrals, independent of the scatter order
    assert abs(R1[3].sum() - full["R"][3].sum()) <= 1e-9 * np.abs(full["R"][3]).sum()


def _bsr(full):
    import scipy.sparse as sp
    V = full["Val"]
    data = np.ascontiguousarray(V.T).reshape(-1, 4, 4)          # K(4i+j, k) -> data[k, i, j]
    nNo = full["m"].nNo
    return sp.bsr_matrix((data, full["colPtr"], full["rowPtr"]), shape=(4 * nNo, 4 * nNo))


def test_fullsize_spmv_matches_scipy_and_is_linear(full):
    eng, nNo = full["eng"], full["m"].nNo
    eng.put_Val(full["Val"], 4)
    K = _bsr(full)
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.standard_normal((4, nNo)))
    Y = np.asfortranarray(rng.standard_normal((4, nNo)))
    KX = eng.spmv(4, X)
    ref = (K @ X.T.reshape(-1)).reshape(nNo, 4).T
    assert common.rel_err(KX, ref) < 1e-13
    KY = eng.spmv(4, Y)
    KZ = eng.spmv(4, np.asfortranarray(2.5 * X - 0.75 * Y))
    assert common.rel_err(KZ, 2.5 * KX - 0.75 * KY) < 1e-13


def test_fullsize_gmres_true_residual(full):
    eng, m = full["eng"], full["m"]
    nNo = m.nNo
    eng.put_Val(full["Val"], 4)
    eng.put_R(full["R"])
    relTol = 1e-3
    ls = abi.ls_params(abi.LS_GMRES, mItr=10, sD=50, relTol=relTol)
    X, out, _ = eng.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1))
    assert out.RI.success and out.RI.itr > 10
    # FSILS diagonal preconditioner restated on the host (Code/Source/linear_solver/precond.cpp:141-216)
    rows = np.repeat(np.arange(nNo), np.diff(full["rowPtr"]))
    dpos = np.nonzero(full["colPtr"] == rows)[0]
    assert len(dpos) == nNo
    W = np.abs(full["Val"][[0, 5, 10, 15]][:, dpos])
    W = np.where(W == 0.0, 1.0, 1.0 / np.sqrt(np.where(W == 0.0, 1.0, W)))
    W[:3, full["wall"]] = 0.0                                    # Dirichlet face, val = 0 on the velocity dofs
    assert common.rel_err(eng.get_W(), W) < 1e-14
    K = _bsr(full)
    r = W * (full["R"] - (K @ X.T.reshape(-1)).reshape(nNo, 4).T)
    nr, n0 = np.linalg.norm(r), np.linalg.norm(W * full["R"])
    assert abs(n0 - out.RI.iNorm) < 1e-10 * n0
    assert nr <= relTol * n0 * 1.02, (nr, n0)
    assert abs(nr - out.RI.fNorm) < 0.05 * nr                    # recurrence residual vs true residual
    assert np.all(X[:3, full["wall"]] == 0.0)
