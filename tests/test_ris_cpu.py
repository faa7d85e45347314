"""CPU: the device's treatment of an open resistive immersed surface (svmultiphysics_b200/csrc/ris.cu) rests on one claim — what
ris::doassem_ris (Code/Source/solver/ris.cpp:269-349) does element by element equals an operation on ASSEMBLED rows: the rows of the
mapped nodes, as one mesh's elements filled them, added into the twin rows with mapped columns replaced by their twins.  Checked here
with numpy on the golden vectors of the compiled reference (tests/golden/ris.npz: the same two-lumen case with the surface closed =
plain assembly, and open = the reference's own doassem_ris)."""
import numpy as np

from tests import common


def test_open_ris_surface_is_a_row_operation_on_the_assembled_system():
    g = common.load_golden("ris.npz")
    rowPtr, colPtr, mp = g["rowPtr"], g["colPtr"], g["map"]
    Rc, Vc, Ro, Vo = g["closed/R"], g["closed/Val"], g["open/R"], g["open/Val"]
    nNo = len(rowPtr) - 1
    twin = np.full(nNo, -1)
    twin[mp[0]], twin[mp[1]] = mp[1], mp[0]
    R, V = Rc.copy(), Vc.copy()
    n_added = 0
    for a in mp.ravel():
        at = twin[a]
        R[:, at] += Rc[:, a]
        for q in range(rowPtr[a], rowPtr[a + 1]):
            b = colPtr[q]
            c = twin[b] if twin[b] >= 0 else b
            row = colPtr[rowPtr[at]:rowPtr[at + 1]]
            hit = np.where(row == c)[0]
            if len(hit) == 0:
                assert not Vc[:, q].any()           # only a connection lhsa added for the RIS (lhsa.cpp:168-193): no element wrote it
                continue
            V[:, rowPtr[at] + hit[0]] += Vc[:, q]
            n_added += bool(Vc[:, q].any())
    assert n_added > 100
    assert common.rel_err(R, Ro) < 1e-13
    assert common.rel_err(V, Vo) < 1e-13
    assert common.rel_err(Vo, Vc) > 1e-3            # the coupling matters
