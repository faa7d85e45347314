"""GPU tests of the rectangular-block SpMV kernels (fsils_spar_mul_vv/sv/vs/ss, linear_solver/spar_mul.cpp:19-231) and of the
interleaved Schur-complement operator (cgrad::schur, linear_solver/cgrad.cpp:77-84): EVERY lane mapping the library can select
against a plain numpy evaluation of the same sums, on a tet graph (short ragged rows) and a hex graph (27 blocks per row)."""
import numpy as np
import pytest

from svmultiphysics_b200 import meshgen
from svmultiphysics_b200.engine import Engine

pytestmark = pytest.mark.gpu

SPMV_TOL = 1e-13   # relative to max |KU|: the products differ from the sequential sum by the order of additions only


def _graphs():
    mt = meshgen.cylinder_tet4(7, 5)
    mh = meshgen.box_hex8(6, 5, 7)
    return [("tet4", mt.nNo, mt.IEN), ("hex8", mh.nNo, mh.IEN)]


def _engine(nNo, IEN):
    e = Engine(0)
    rp, cp = e.lhsa(nNo, [IEN])
    e.set_graph(rp, cp)
    return e, rp, cp


def _ref_spmv(R, Cc, rp, cp, K, U):
    nNo = len(rp) - 1
    rows = np.repeat(np.arange(nNo), np.diff(rp))
    KU = np.zeros((R, nNo))
    for i in range(R):
        contrib = np.zeros(len(cp))
        for j in range(Cc):
            contrib += K[i * Cc + j] * U[j, cp]
        KU[i] = np.bincount(rows, weights=contrib, minlength=nNo)
    return KU


@pytest.mark.parametrize("R,Cc", [(3, 3), (3, 1), (1, 3), (1, 1)])
def test_spmv_rc_all_variants(R, Cc):
    rng = np.random.default_rng(11)
    for name, nNo, IEN in _graphs():
        e, rp, cp = _engine(nNo, IEN)
        K = np.asfortranarray(rng.standard_normal((R * Cc, len(cp))))
        U = np.asfortranarray(rng.standard_normal((Cc, nNo)))
        ref = _ref_spmv(R, Cc, rp, cp, K, U)
        nv = e.spmv_rc_variants(R, Cc)
        assert nv >= 2
        for v in [-1] + list(range(nv)):
            KU = e.spmv_rc(R, Cc, K, U, variant=v)
            err = np.abs(KU - ref).max() / np.abs(ref).max()
            assert err < SPMV_TOL, (name, R, Cc, v, err)
        e.close()


def test_schur_operator_all_variants():
    rng = np.random.default_rng(12)
    for name, nNo, IEN in _graphs():
        e, rp, cp = _engine(nNo, IEN)
        nnz = len(cp)
        L = rng.standard_normal(nnz)
        Gt = np.asfortranarray(rng.standard_normal((3, nnz)))
        P = rng.standard_normal(nNo)
        GP = np.asfortranarray(rng.standard_normal((3, nNo)))
        ref = _ref_spmv(1, 1, rp, cp, L[None, :], P[None, :])[0] - _ref_spmv(1, 3, rp, cp, Gt, GP)[0]
        ref_dot = float(P @ ref)
        for v in [-2, -1] + list(range(e.schur_sp_variants())):
            SP, dot = e.schur_sp(L, Gt, P, GP, variant=v)
            err = np.abs(SP - ref).max() / np.abs(ref).max()
            assert err < SPMV_TOL, (name, v, err)
            assert abs(dot - ref_dot) <= 1e-12 * np.abs(P * ref).sum(), (name, v, dot, ref_dot)
        e.close()
