"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/svb200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

from svmultiphysics_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(engine.LIB_PATH):
        subprocess.check_call(["make", "-j", "8"], cwd=os.path.join(ROOT, "svmultiphysics_b200", "csrc"))
    return engine.load_library()


def _gpu_visible():
    # nvidia-smi rather than torch.cuda: importing torch after the reference library was loaded RTLD_GLOBAL crashes
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
    except Exception:
        return False
    return any(line.startswith("GPU ") for line in out.splitlines())


def header_symbols():
    src = open(os.path.join(ROOT, "include", "svb200.h")).read()
    return sorted(set(re.findall(r"SVB200_API\s+[\w\s\*]+?\b(svb200_\w+)\s*\(", src)))


def test_header_and_library_agree(lib):
    declared = header_symbols()
    assert len(declared) >= 30
    assert sorted(engine.ABI_SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), f"libsvb200.so does not export {s}"
    lib.svb200_abi_version.restype = C.c_int
    assert lib.svb200_abi_version() == abi.ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """Sizes and field offsets of the ctypes mirrors against what a C compiler makes of include/svb200.h."""
    src = tmp_path / "layout.c"
    fields = {
        "svb200_eqparams": (abi.EqParams, ["dt", "phys", "reserved"]),
        "svb200_dmnparams": (abi.DmnParams, ["Id", "rho", "viscType", "volType", "Kpen", "st_a", "ctau_C", "active_stress",
                                              "cann_rows", "cann_inv", "cann_act", "cann_w"]),
        "svb200_sublsparams": (abi.SubLsParams, ["mItr", "absTol"]),
        "svb200_lsparams": (abi.LsParams, ["CG"]),
        "svb200_sublsresult": (abi.SubLsResult, ["success", "callD"]),
        "svb200_lsresult": (abi.LsResult, ["Resm", "hist"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "svb200.h"', 'int main(void) {']
    for cname, (_, names) in fields.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for n in names:
            lines.append(f'  printf("{cname}.{n} %zu\\n", offsetof({cname}, {n}));')
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, (cls, names) in fields.items():
        assert C.sizeof(cls) == int(out[cname]), cname
        for n in names:
            assert getattr(cls, n).offset == int(out[f"{cname}.{n}"]), f"{cname}.{n}"


def test_no_cpu_fallback(lib):
    if _gpu_visible():
        pytest.skip("a GPU is visible; the no-GPU failure path cannot be exercised")
    with pytest.raises(engine.Svb200Error, match="no usable CUDA device"):
        engine.Engine(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under svmultiphysics_b200/ may reference it."""
    pkg = os.path.join(ROOT, "svmultiphysics_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, fn)).read()
                assert "oracle" not in text.replace("the oracle", "").lower() or fn == "meshgen.py", f"{fn} mentions the oracle"


def test_gen_alpha_matches_reference_formula():
    # rho_inf = 0.5 (pipe_RCR_3d): am = 5/6, af = 2/3, gam = 2/3 (Code/Source/solver/initialize.cpp:484-486)
    af, am, gam, beta = abi.gen_alpha(0.5)
    assert abs(am - 5.0 / 6.0) < 1e-15 and abs(af - 2.0 / 3.0) < 1e-15 and abs(gam - 2.0 / 3.0) < 1e-15


def test_cpp_host_plugin_loads_and_fails_loudly_without_gpu():
    """svmultiphysics_b200/host/B200LinearAlgebra (the reference-side C++ plug-in) links against the C ABI and, like the
    library itself, has no CPU path: on a box without a B200 ls_alloc through it raises instead of falling back."""
    import numpy as np
    from oracle import refbind
    from svmultiphysics_b200 import meshgen
    if not refbind.have_host():
        pytest.skip("host plug-in is built only where the reference tree is present")
    if _gpu_visible():
        pytest.skip("GPU present: covered by tests/test_gpu_hostshim.py")
    m = meshgen.cylinder_tet4(3, 3)
    c = refbind.RefCase()
    c.set_coords(m.x); c.add_mesh(m.IEN); c.build_graph(0)
    c.use_b200_backend(device=0)
    with pytest.raises(RuntimeError):
        c.alloc(4)
    c.close()
