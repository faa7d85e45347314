"""CPU: the committed generalised-alpha fixtures (tests/golden/genalpha.npz) are what the COMPILED reference's Integrator and
set_bc::set_bc_dir produce (when oracle/_ref/libsvref.so is present), and the numpy restatement oracle/genalpha_oracle.py — still used
as the host side of some device-resident Newton-loop tests — reproduces them bit for bit."""
import numpy as np
import pytest

from svmultiphysics_b200 import abi
from tests import common
from tests import genalpha_scenarios as gs


@pytest.mark.parametrize("name", ["fsi_mesh", "fluid", "ustruct", "fsi_ustruct"])
def test_golden_is_the_compiled_reference(name):
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs oracle/_ref/libsvref.so")
    golden = common.load_golden("genalpha.npz")
    live = gs.run_reference(gs.scenarios()[name])
    keys = [k for k in golden.files if k.startswith(name + "/")]
    assert len(keys) == len(live)
    for k in keys:
        assert np.array_equal(live[k[len(name) + 1:]], golden[k]), k


@pytest.mark.parametrize("name", ["fsi_mesh", "fluid", "ustruct", "fsi_ustruct"])
def test_numpy_restatement_matches_golden(name):
    from oracle import genalpha_oracle as go
    golden = common.load_golden("genalpha.npz")
    s = gs.scenarios()[name]
    Ao, Yo, Do, An, Yn, Dn = [a.copy(order="F") for a in s["states"]]
    Ad = None if s["Ad"] is None else s["Ad"].copy(order="F")
    eqs = s["eqs"]
    go.predictor(eqs, gs.DT, s["dFlag"], Ao, Yo, Do, An, Yn, Dn, Ad)
    for nm, a in zip("AYD", (An, Yn, Dn)):
        assert np.array_equal(a, golden[f"{name}/predictor/{nm}"])
    if Ad is not None:
        assert np.array_equal(Ad, golden[f"{name}/predictor/Ad"])
    # continue from the reference's state after set_bc_dir (the restatement has no BC routine)
    An, Yn, Dn = [golden[f"{name}/set_bc_dir/{nm}"].copy(order="F") for nm in "AYD"]
    if Ad is not None:
        Ad = golden[f"{name}/set_bc_dir/Ad"].copy(order="F")
    Ag, Yg, Dg = (np.zeros_like(An) for _ in range(3))
    go.initiator(eqs, Ao, Yo, Do, An, Yn, Dn, Ag, Yg, Dg)
    for nm, a in zip("AYD", (Ag, Yg, Dg)):
        r = slice(eqs[0].s, eqs[-1].e + 1)
        assert np.array_equal(a[r], golden[f"{name}/initiator/{nm}"][r])
    if go.is_sst(eqs[0]):
        go.corrector_ustruct(eqs[0], gs.DT, s["R"], s["Rd"], An, Yn, Dn, Ad, mesh_s=4 if s["solid"] is not None else -1, solid=s["solid"])
        assert np.array_equal(Ad, golden[f"{name}/corrector/Ad"])
    else:
        go.corrector(eqs[0], gs.DT, s["R"], An, Yn, Dn, mesh_s=4 if s["solid"] is not None else -1, solid=s["solid"])
    for nm, a in zip("AYD", (An, Yn, Dn)):
        assert np.array_equal(a, golden[f"{name}/corrector/{nm}"])
