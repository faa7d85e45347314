"""Parity at BASELINE.json's full size for config C4: 5,000,211 hex8 neo-Hookean/ST91 (struct_3d) on one B200.

As for C2 (tests/test_gpu_fullsize.py) the oracle cannot assemble the full mesh in seconds, so the full-size run is checked through
  * z sub-slabs of the SAME mesh and state assembled by the compiled reference: every row whose elements all lie inside
    the sub-slab must agree to 1e-12,
  * the atomic and the coloured (bitwise reproducible) scatter agreeing to 1e-12, and the coloured one with itself bit for bit,
  * symmetry of the assembled tangent of a hyperelastic solid (K_ab = K_ba^T block by block) and linearity of the dof-3 SpMV,
  * BiCGStab reaching the set tolerance with the TRUE preconditioned residual recomputed on the host.
"""
import numpy as np
import pytest

from oracle import refbind
from svmultiphysics_b200 import abi, elements, meshgen
from tests import common

pytestmark = pytest.mark.gpu

NH = 171                  # 171^3 = 5,000,211 hex8
DT = 1e-4


@pytest.fixture(scope="module")
def full():
    from svmultiphysics_b200.engine import Engine
    m = meshgen.box_hex8(NH, NH, NH, (1e-3, 1e-3, 1e-3))
    assert m.nEl >= 5_000_000
    Ag, Yg, Dg, Bf, _ = common.struct_state(m, 0)
    Dg *= 0.1                # keep det F > 0 on the fine mesh (tools/bench_struct.py)
    eng = Engine(0)
    rowPtr, colPtr = eng.lhsa(m.nNo, [m.IEN])
    eng.set_graph(rowPtr, colPtr)
    w, Nt, Nx = elements.tables(8)
    eng.set_mesh(0, m.IEN, w, Nt, Nx)
    eng.set_coords(m.x)
    eng.alloc(3)
    eng.set_state(Ag, Yg, Dg, Bf)
    eq, dmn = abi.struct_eq(DT), [abi.struct_domain()]
    eng.assemble(0, eq, dmn)
    d = dict(m=m, Ag=Ag, Yg=Yg, Dg=Dg, Bf=Bf, eng=eng, rowPtr=rowPtr, colPtr=colPtr, R=eng.get_R(), Val=eng.get_Val(), dmn=dmn)
    yield d
    eng.close()


@pytest.mark.parametrize("k0,k1", [(0, 2), (85, 87), (NH - 2, NH)], ids=["bottom", "middle", "top"])
def test_fullsize_struct_matches_reference_on_subslab(full, k0, k1):
    if not refbind.have_ref():
        pytest.skip("struct_3d is checked against libsvref.so")
    m = full["m"]
    n1 = NH + 1
    lo, hi = n1 * n1 * k0, n1 * n1 * (k1 + 1)
    layer = m.IEN.min(axis=0) // (n1 * n1)
    els = np.nonzero((layer >= k0) & (layer < k1))[0]
    assert len(els) == NH * NH * (k1 - k0)
    IEN = np.asfortranarray(m.IEN[:, els] - lo)
    c = refbind.RefCase()
    c.set_coords(np.asfortranarray(m.x[:, lo:hi]))
    c.add_mesh(IEN)
    rp, cp = c.build_graph(0)
    c.alloc(3)
    c.set_state(*(np.asfortranarray(full[k][:, lo:hi]) for k in ("Ag", "Yg", "Dg", "Bf")))
    c.assemble(0, abi.struct_eq(DT), full["dmn"])
    R0, V0 = c.get_R(), c.get_Val()
    # rows that see all of their elements inside the sub-slab: interior planes, plus the outer plane at the mesh boundary
    r0 = 0 if k0 == 0 else n1 * n1
    r1 = (hi - lo) if k1 == NH else (hi - lo) - n1 * n1
    g0, g1 = r0 + lo, r1 + lo
    RP, CP = full["rowPtr"], full["colPtr"]
    assert np.array_equal(np.diff(rp[r0:r1 + 1]), np.diff(RP[g0:g1 + 1]))
    s0, s1, t0, t1 = rp[r0], rp[r1], RP[g0], RP[g1]
    assert np.array_equal(cp[s0:s1] + lo, CP[t0:t1])
    assert r1 - r0 >= n1 * n1
    assert common.rel_err(full["R"][:, g0:g1], R0[:, r0:r1]) < 1e-12
    assert common.rel_err(full["Val"][:, t0:t1], V0[:, s0:s1]) < 1e-12


def test_fullsize_struct_atomic_equals_colored_and_is_symmetric(full):
    eng = full["eng"]
    eq = abi.struct_eq(DT, scatter=abi.SCATTER_COLORED)
    eng.alloc(3); eng.assemble(0, eq, full["dmn"])
    R1, V1 = eng.get_R(), eng.get_Val()
    assert common.rel_err(R1, full["R"]) < 1e-12
    assert common.rel_err(V1, full["Val"]) < 1e-12
    eng.alloc(3); eng.assemble(0, eq, full["dmn"])
    assert np.array_equal(eng.get_Val(), V1)               # coloured scatter: bitwise reproducible
    # K_ab = K_ba^T for a hyperelastic solid without viscosity: check on a sample of rows through the CSR structure
    RP, CP = full["rowPtr"], full["colPtr"]
    rng = np.random.default_rng(3)
    worst = 0.0
    scale = np.abs(V1).max()
    for a in rng.integers(0, full["m"].nNo, size=2000):
        for s in range(RP[a], RP[a + 1]):
            b = CP[s]
            t = RP[b] + np.searchsorted(CP[RP[b]:RP[b + 1]], a)
            assert CP[t] == a
            Kab = V1[:, s].reshape(3, 3)
            Kba = V1[:, t].reshape(3, 3)
            worst = max(worst, np.abs(Kab - Kba.T).max() / scale)
    assert worst < 1e-13


def test_fullsize_struct_spmv_linear_and_bicgs_true_residual(full):
    import scipy.sparse as sp
    eng, m = full["eng"], full["m"]
    nNo = m.nNo
    eng.put_Val(full["Val"], 3)
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.standard_normal((3, nNo)))
    Y = np.asfortranarray(rng.standard_normal((3, nNo)))
    KX, KY = eng.spmv(3, X), eng.spmv(3, Y)
    KZ = eng.spmv(3, np.asfortranarray(2.5 * X - 0.75 * Y))
    assert common.rel_err(KZ, 2.5 * KX - 0.75 * KY) < 1e-13
    data = np.ascontiguousarray(full["Val"].T).reshape(-1, 3, 3)
    K = sp.bsr_matrix((data, full["colPtr"], full["rowPtr"]), shape=(3 * nNo, 3 * nNo))
    assert common.rel_err(KX, (K @ X.T.reshape(-1)).reshape(nNo, 3).T) < 1e-13
    # block_compression-like Dirichlet planes, BiCGStab as in tests/cases/struct/block_compression/solver.xml
    faces = []
    for k, name in enumerate(("X0", "Y0", "Z0")):
        val = np.ones((3, len(m.faces[name])), order="F"); val[k] = 0.0
        faces.append((abi.BC_DIR, m.faces[name], val))
    eng.set_num_faces(3)
    for i, (g, nodes, val) in enumerate(faces):
        eng.set_face(i, g, nodes, val)
    eng.put_Val(full["Val"], 3); eng.put_R(full["R"])
    relTol = 1e-2
    ls = abi.ls_params(abi.LS_BICGS, mItr=400, relTol=relTol)
    Xs, out, _ = eng.solve(3, abi.LS_BICGS, ls, np.ones(3, np.int32), np.zeros(3))
    assert out.RI.success
    W = eng.get_W()
    # Xs is the un-scaled increment (solve.cpp:157-159 multiplies by W), so the preconditioned residual is W (R - K Xs)
    assert not Xs[W == 0.0].any()                              # constrained dofs keep a zero increment
    r = W * (full["R"] - (K @ Xs.T.reshape(-1)).reshape(nNo, 3).T)
    nr, n0 = np.linalg.norm(r), np.linalg.norm(W * full["R"])
    assert abs(n0 - out.RI.iNorm) < 1e-10 * n0
    assert nr <= relTol * n0 * 1.05, (nr, n0)
