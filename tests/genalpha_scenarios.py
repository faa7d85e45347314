"""Generalised-alpha parity scenarios: the same deterministic inputs run through (a) the reference's own Integrator and
set_bc::set_bc_dir (oracle.refbind.GenAlphaRef, compiled from Code/Source/solver/Integrator.cpp / set_bc.cpp), whose outputs are
committed as tests/golden/genalpha.npz by tests/golden/make_genalpha_golden.py, and (b) the device kernels through the C ABI.
Every stage's arrays must agree BIT FOR BIT."""
import numpy as np

from svmultiphysics_b200 import abi
from tests import common

DT = 0.0125


def _bc_values(g, gx, nV, eDrn, lDof):
    """What set_bc::set_bc_dir_l hands back for a steady BC (set_bc.cpp:1159-1181): (lA, lY), same rounding order."""
    n = len(gx)
    if lDof == 3:
        lY = np.asfortranarray((g * gx)[None, :] * nV)          # dirY * gx(a) * nV(i,a)
        lA = np.asfortranarray((0.0 * gx)[None, :] * nV)
    else:
        lY = np.asfortranarray(np.repeat((g * gx)[None, :], lDof, axis=0))
        lA = np.asfortranarray(np.repeat((0.0 * gx)[None, :], lDof, axis=0))
    return lA, lY


def scenarios():
    m, *_ = common.fsi_case()
    rng = np.random.default_rng(2024)
    nNo = m.nNo
    wall, inlet = m.faces["wall"], m.faces["inlet"]
    solid = np.zeros(nNo, np.int32)
    solid[np.unique(m.IEN[:, (m.eId & 2) != 0])] = 1

    def states(tDof):
        return [np.asfortranarray(rng.standard_normal((tDof, nNo))) for _ in range(6)]

    def unit(n):
        v = rng.standard_normal((3, n))
        return np.asfortranarray(v / np.linalg.norm(v, axis=0))

    out = {}
    out["fsi_mesh"] = dict(
        mesh=m, tDof=7, dFlag=1, sstEq=0, eqs=[abi.eq_time(0, 3, abi.PHYS_FSI, 0.5), abi.eq_time(4, 6, abi.PHYS_MESH, 0.2)],
        states=states(7), Ad=None, solid=solid,
        bcs=[dict(iEq=0, nodes=wall, eDrn=(0, 0, 0), impD=False, g=1.7, gx=rng.standard_normal(len(wall)), nV=unit(len(wall))),
             dict(iEq=0, nodes=inlet, eDrn=(1, 0, 1), impD=True, g=0.3, gx=rng.standard_normal(len(inlet)), nV=unit(len(inlet)))],
        R=np.asfortranarray(rng.standard_normal((4, nNo))), Rd=None)
    out["fluid"] = dict(
        mesh=m, tDof=4, dFlag=0, sstEq=0, eqs=[abi.eq_time(0, 3, abi.PHYS_FLUID, 0.5)], states=states(4), Ad=None, solid=None,
        bcs=[dict(iEq=0, nodes=wall, eDrn=(0, 0, 0), impD=False, g=-2.5, gx=rng.standard_normal(len(wall)), nV=unit(len(wall)))],
        R=np.asfortranarray(rng.standard_normal((4, nNo))), Rd=None)
    out["ustruct"] = dict(
        mesh=m, tDof=4, dFlag=1, sstEq=1, eqs=[abi.eq_time(0, 3, abi.PHYS_USTRUCT, 0.5)], states=states(4),
        Ad=np.asfortranarray(rng.standard_normal((3, nNo))), solid=None,
        bcs=[dict(iEq=0, nodes=wall, eDrn=(0, 0, 0), impD=False, g=0.8, gx=rng.standard_normal(len(wall)), nV=unit(len(wall))),
             dict(iEq=0, nodes=inlet, eDrn=(0, 1, 0), impD=True, g=-0.6, gx=rng.standard_normal(len(inlet)), nV=unit(len(inlet)))],
        R=np.asfortranarray(rng.standard_normal((4, nNo))), Rd=np.asfortranarray(rng.standard_normal((3, nNo))))
    # FSI with ustruct solids (tests/cases/fsi_ustruct): com_mod.sstEq routes the FSI rows through the velocity-pressure branches of
    # the predictor / set_bc_dir / corrector (Integrator.cpp:626-630, 828-846, set_bc.cpp:1062), the mesh equation keeps its own
    out["fsi_ustruct"] = dict(
        mesh=m, tDof=7, dFlag=1, sstEq=1, solid_phys=abi.PHYS_USTRUCT,
        eqs=[abi.eq_time(0, 3, abi.PHYS_FSI, 0.5, sstEq=True), abi.eq_time(4, 6, abi.PHYS_MESH, 0.2)],
        states=states(7), Ad=np.asfortranarray(rng.standard_normal((3, nNo))), solid=solid,
        bcs=[dict(iEq=0, nodes=wall, eDrn=(0, 0, 0), impD=False, g=1.1, gx=rng.standard_normal(len(wall)), nV=unit(len(wall))),
             dict(iEq=0, nodes=inlet, eDrn=(1, 1, 0), impD=True, g=-0.4, gx=rng.standard_normal(len(inlet)), nV=unit(len(inlet)))],
        R=np.asfortranarray(rng.standard_normal((4, nNo))), Rd=np.asfortranarray(rng.standard_normal((3, nNo))))
    return out


def run_reference(s):
    """-> {stage/array: values} from the compiled reference."""
    from oracle import refbind
    Ao, Yo, Do, An, Yn, Dn = [a.copy(order="F") for a in s["states"]]
    g = refbind.GenAlphaRef(s["eqs"], DT, s["dFlag"], s["sstEq"], Ao, Yo, Do, maxBc=4)
    g.set(abi.SOL_CURRENT, An, Yn, Dn)
    if s["Ad"] is not None:
        g.set_ad(s["Ad"])
    out = {}

    def snap(stage, which=abi.SOL_CURRENT):
        for name, a in zip("AYD", g.get(which)):
            out[f"{stage}/{name}"] = a
        if s["Ad"] is not None:
            out[f"{stage}/Ad"] = g.get_ad()

    g.predictor(); snap("predictor")
    for b in s["bcs"]:
        g.add_dir_bc(b["iEq"], b["nodes"], b["eDrn"], b["impD"], b["g"], b["gx"], b["nV"])
    g.set_bc_dir(); snap("set_bc_dir")
    g.initiator(0); snap("initiator", abi.SOL_INTERMEDIATE)
    if s["solid"] is not None:
        g.set_solid_nodes(0, s.get("solid_phys", abi.PHYS_STRUCT), s["solid"])
    g.corrector(0, s["R"], s["Rd"]); snap("corrector")
    g.close()
    return out


def run_engine(s):
    """-> the same dictionary from the device kernels (svb200_predictor / set_dirichlet_rows / dirichlet_ustruct / initiator /
    corrector); the prescribed Dirichlet values are computed by the host, as in the reference."""
    from oracle import refbind
    m = s["mesh"]
    cls = refbind.RefCase if refbind.have_ref() else refbind.OracleCase
    orc, rowPtr, colPtr = common.make_oracle(cls, m)
    eng = common.make_engine(m, rowPtr, colPtr)
    Ao, Yo, Do, An, Yn, Dn = s["states"]
    eng.set_solution(abi.SOL_OLD, Ao, Yo, Do)
    eng.set_solution(abi.SOL_CURRENT, An, Yn, Dn)
    if s["Ad"] is not None:
        eng.set_ad(s["Ad"])
    eqs = s["eqs"]
    out = {}

    def snap(stage, which=abi.SOL_CURRENT):
        for name, a in zip("AYD", eng.get_solution(which)):
            out[f"{stage}/{name}"] = a
        if s["Ad"] is not None:
            out[f"{stage}/Ad"] = eng.get_ad()

    eng.predictor(eqs, DT, s["dFlag"]); snap("predictor")
    for b in s["bcs"]:
        q = eqs[b["iEq"]]
        dirs = [i for i in range(3) if b["eDrn"][i]]
        lDof = len(dirs) if dirs else 3                      # e - s + 1 with e reduced by one for dof = nsd + 1 (set_bc.cpp:971-989)
        lA, lY = _bc_values(b["g"], b["gx"], b["nV"], b["eDrn"], lDof)
        rows = dirs if dirs else [0, 1, 2]
        for k, i in enumerate(rows):
            va, vy = np.asfortranarray(lA[k:k + 1]), np.asfortranarray(lY[k:k + 1])
            if b["impD"]:
                eng.set_dirichlet_rows(q.s + i, b["nodes"], valY=va, valD=vy)      # Yn = tmpA, Dn = tmpY (set_bc.cpp:1004-1016)
            else:
                eng.set_dirichlet_rows(q.s + i, b["nodes"], valA=va, valY=vy)
        if q.phys == abi.PHYS_USTRUCT or (q.phys == abi.PHYS_FSI and s["sstEq"]):      # set_bc.cpp:1062
            eng.dirichlet_ustruct(q, DT, b["nodes"], dir_mask=sum(1 << i for i in rows), impD=b["impD"])
    snap("set_bc_dir")
    eng.initiator(eqs); snap("initiator", abi.SOL_INTERMEDIATE)
    eng.alloc(4); eng.put_R(s["R"])
    if s["Rd"] is not None:
        eng.put_Rd(s["Rd"])
    if s["solid"] is not None:
        eng.set_node_flags(s["solid"])
    eng.corrector(eqs[0], DT, mesh_s=4 if s["solid"] is not None else -1); snap("corrector")
    eng.close()
    return out
