"""CPU tests: the C restatement (oracle/sv_oracle.c) against the golden vectors produced by the
unmodified reference, and against the compiled reference itself when it is present."""
import os
import subprocess

import numpy as np
import pytest

from svmultiphysics_b200 import abi, elements
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def oracle_cls():
    from oracle import refbind
    if not os.path.exists(refbind.ORACLE_SO):
        subprocess.check_call(["make", "oracle"], cwd=os.path.join(ROOT, "oracle"))
    return refbind.OracleCase


@pytest.fixture(scope="module")
def golden():
    return common.load_golden()


def _setup(cls, case):
    name, visc, Kd, f, tDof, mv = case
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=common.GOLDEN_N, nz=common.GOLDEN_NZ, tDof=tDof)
    faces = common.dirichlet_faces(m)
    c, rowPtr, colPtr = common.make_oracle(cls, m, nFaces=len(faces))
    for i, (g, nodes, val) in enumerate(faces):
        c.set_face(i, g, nodes, val)
    eq = abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv)
    dmn = [abi.fluid_domain(K_darcy=Kd, f=f, **visc)]
    c.alloc(4); c.set_state(Ag, Yg, Dg, Bf)
    return c, m, faces, eq, dmn, rowPtr, colPtr


@pytest.mark.parametrize("case", common.FLUID_CASES, ids=[c[0] for c in common.FLUID_CASES])
def test_restatement_matches_golden_assembly(oracle_cls, golden, case):
    c, m, faces, eq, dmn, rowPtr, colPtr = _setup(oracle_cls, case)
    assert np.array_equal(rowPtr, golden["rowPtr"]) and np.array_equal(colPtr, golden["colPtr"])
    c.assemble(0, eq, dmn)
    # same operation order as the reference, no FMA contraction: agreement is to the last bits
    assert common.rel_err(c.get_R(), golden[f"{case[0]}/R"]) < 1e-14
    assert common.rel_err(c.get_Val(), golden[f"{case[0]}/Val"]) < 1e-14
    KU = c.spmv(4, common.spmv_vector(m.nNo))
    assert common.rel_err(KU, golden[f"{case[0]}/KU"]) < 1e-14


@pytest.mark.parametrize("ls_case", common.LS_CASES, ids=[c[0] for c in common.LS_CASES])
@pytest.mark.parametrize("case", [common.FLUID_CASES[0], common.FLUID_CASES[3]], ids=["newtonian", "casson"])
def test_restatement_matches_golden_solve(oracle_cls, golden, case, ls_case):
    c, m, faces, eq, dmn, *_ = _setup(oracle_cls, case)
    c.put_R(golden[f"{case[0]}/R"]); c.put_Val(golden[f"{case[0]}/Val"], 4)
    ls_name, ls_type, kw = ls_case
    X, o, hist = c.solve(4, ls_type, abi.ls_params(ls_type, **kw), np.ones(len(faces), np.int32), np.zeros(len(faces)), hist_cap=600)
    stats = golden[f"{case[0]}/{ls_name}/stats"]
    assert o.RI.itr == int(stats[0]) and o.RI.success == int(stats[1])
    assert abs(o.RI.iNorm - stats[2]) <= 1e-13 * stats[2]
    assert abs(o.RI.fNorm - stats[3]) <= 1e-6 * stats[3]
    assert common.rel_err(X, golden[f"{case[0]}/{ls_name}/X"]) < 1e-9
    if ls_type == abi.LS_GMRES:
        assert len(hist) > 0 and abs(hist[-1] - o.RI.fNorm) <= 1e-12 * o.RI.fNorm   # history ends at fNorm


def test_restatement_matches_compiled_reference(oracle_cls, ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref/libsvref.so not built (needs /root/reference)")
    from oracle.refbind import RefCase
    case = ("cy_mv", dict(viscType=abi.VISC_CY, mu=0.035, mu_o=0.16, lam=8.2, a=0.64, n=0.2128), 2.0, (0.1, 0.2, 0.3), 7, 1)
    res = []
    for cls in (RefCase, oracle_cls):
        name, visc, Kd, f, tDof, mv = case
        m, Ag, Yg, Dg, Bf = common.fluid_case(n=5, nz=6, tDof=tDof)
        faces = common.dirichlet_faces(m)
        c, rowPtr, colPtr = common.make_oracle(cls, m, nFaces=len(faces))
        for i, (g, nodes, val) in enumerate(faces):
            c.set_face(i, g, nodes, val)
        c.alloc(4); c.set_state(Ag, Yg, Dg, Bf)
        c.assemble(0, abi.fluid_eq(0.005, tDof=tDof, mvMsh=mv), [abi.fluid_domain(K_darcy=Kd, f=f, **visc)])
        R, V = c.get_R(), c.get_Val()
        X, o, _ = c.solve(4, abi.LS_GMRES, abi.ls_params(abi.LS_GMRES, mItr=30, sD=20, relTol=1e-9), np.ones(2, np.int32), np.zeros(2))
        res.append((rowPtr, colPtr, R, V, X, o.RI.itr, o.RI.fNorm))
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])     # bit-identical assembly
    assert a[5] == b[5] and common.rel_err(b[4], a[4]) < 1e-12


@pytest.mark.parametrize("eNoN", [4, 8])
def test_element_tables_match_golden(golden, oracle_cls, eNoN):
    w, N, Nx = elements.tables(eNoN)
    assert np.array_equal(w, golden[f"tables{eNoN}/w"])
    assert np.abs(N - golden[f"tables{eNoN}/N"]).max() < 1e-15
    assert np.abs(Nx - golden[f"tables{eNoN}/Nx"]).max() < 1e-15
    # FE basis identities the reference's unit tests pin (tests/unitTests/FE/Basis/test_LagrangeBasis.cpp:251,469)
    assert np.allclose(N.sum(axis=0), 1.0, atol=1e-14)          # partition of unity
    assert np.allclose(Nx.sum(axis=1), 0.0, atol=1e-14)         # gradients sum to zero


def test_oracle_error_paths(oracle_cls):
    from svmultiphysics_b200 import meshgen
    m = meshgen.cylinder_tet4(2, 2)
    c, *_ = common.make_oracle(oracle_cls, m, nFaces=1)
    with pytest.raises(RuntimeError, match="faIn is exceeding"):
        c.set_face(3, abi.BC_DIR, m.faces["wall"], np.zeros((3, len(m.faces["wall"])), order="F"))
    # zero right-hand side: GMRES returns immediately with success and leaves R untouched (gmres.cpp:470-475)
    c.alloc(4)
    X, o, _ = c.solve(4, abi.LS_GMRES, abi.ls_params(abi.LS_GMRES), np.ones(1, np.int32), np.zeros(1))
    assert o.RI.success == 1 and o.RI.itr == 0 and np.all(X == 0.0)


def test_face_tables_and_face_extraction_match_reference():
    """TRI3 / QUD4 boundary-face Gauss points and shape functions of svmultiphysics_b200.elements against the compiled
    reference's nn::get_gip / nn::get_gnn (solver/nn.cpp:455,500), and the face-element extraction used by the tests."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    from svmultiphysics_b200 import elements, meshgen
    m, *_ = common.fluid_case()
    orc, _, _ = common.make_oracle(refbind.RefCase, m)
    IENb, gE = meshgen.boundary_face_elements(m, m.faces["outlet_all"])
    assert IENb.shape == (3, 2 * m.lattice[0] * m.lattice[1])
    for a in range(3):      # every face node belongs to its parent element
        assert (m.IEN[:, gE] == IENb[a]).any(axis=0).all()
    i = orc.add_face(0, IENb, gE)
    for got, want in zip(orc.face_tables(0, i), elements.face_tables(3)):
        assert np.array_equal(got, want)
    mh = common.STRUCT_CASES[0][1]()
    o2 = refbind.RefCase(); o2.set_coords(mh.x); o2.add_mesh(mh.IEN); o2.build_graph(0)
    Ib, gEb = meshgen.boundary_face_elements(mh, mh.faces["Z1"])
    assert Ib.shape == (4, mh.lattice[0] * mh.lattice[1])
    j = o2.add_face(0, Ib, gEb)
    for got, want in zip(o2.face_tables(0, j), elements.face_tables(4)):
        assert np.allclose(got, want, rtol=0, atol=1e-15)


def test_genalpha_restatement_identities():
    """Known answers of the generalised-alpha restatement (oracle/genalpha_oracle.py): with rho_inf = 1 the scheme is the
    mid-point rule (am = af = 1/2, gam = 1/2 ... ), a constant-acceleration field is reproduced exactly by
    predictor + corrector, and initiator(am = af = 1) returns the current state."""
    from oracle import genalpha_oracle as go
    from svmultiphysics_b200 import abi
    rng = np.random.default_rng(0)
    q = abi.eq_time(0, 2, abi.PHYS_STRUCT, 0.3)
    dt = 0.01
    Ao, Yo, Do = (rng.standard_normal((3, 5)) for _ in range(3))
    An, Yn, Dn = (np.zeros((3, 5)) for _ in range(3))
    go.predictor([q], dt, 1, Ao, Yo, Do, An, Yn, Dn)
    assert np.array_equal(Yn, Yo) and np.allclose(An, Ao * (q.gam - 1) / q.gam)
    # Newmark consistency: correct An back to Ao (R = An - Ao) => Yn = Yo + dt Ao, Dn = Do + dt Yo + dt^2/2 Ao
    R = An - Ao
    go.corrector(q, dt, R, An, Yn, Dn)
    assert np.allclose(An, Ao) and np.allclose(Yn, Yo + dt * Ao) and np.allclose(Dn, Do + dt * Yo + 0.5 * dt * dt * Ao)
    q1 = abi.EqTime(s=0, e=2, phys=0, af=1.0, am=1.0, gam=0.5, beta=0.25)
    Ag, Yg, Dg = (np.zeros((3, 5)) for _ in range(3))
    go.initiator([q1], Ao, Yo, Do, An, Yn, Dn, Ag, Yg, Dg)
    assert np.array_equal(Ag, An) and np.array_equal(Yg, Yn) and np.array_equal(Dg, Dn)


def test_algorithmic_flop_count_pins_the_roofline_numerator():
    """oracle/flop_count.cpp instantiates the restatement's Gauss-point arithmetic (oracle/fluid_gp.inc, the text that
    reproduces the reference bit for bit) over a counting scalar.  bench.py's FP64 roofline uses SURVEY.md 8(d)'s hand count
    of 11.6 kflop per TET4 element; it must not exceed the counted figure (structural zeros removed, shared front part
    evaluated once) and must agree with it within the SURVEY's +-10 %."""
    import json
    out = subprocess.check_output(["make", "-s", "flops"], cwd=os.path.join(ROOT, "oracle"), text=True)
    d = json.loads(out)
    counted = d["structural_zeros_removed"]["per_element_front_part_evaluated_once"]
    executed = d["executed_by_restatement"]["per_element_2gnn_plus_nG_times_m_plus_c"]
    import bench
    assert bench.FLOP_PER_ELEMENT <= counted <= executed
    assert abs(counted - bench.FLOP_PER_ELEMENT) <= 0.10 * bench.FLOP_PER_ELEMENT
    assert d["survey_hand_count"]["per_element"] == bench.FLOP_PER_ELEMENT
