"""Regenerates tests/golden/genalpha.npz from the COMPILED REFERENCE (oracle/_ref/libsvref.so: Integrator.cpp, set_bc.cpp built
unmodified by oracle/Makefile).  Run from the repository root in the container that has /root/reference:
    make -C oracle ref && python tests/golden/make_genalpha_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import genalpha_scenarios as gs  # noqa: E402

out = {}
for name, s in gs.scenarios().items():
    for k, v in gs.run_reference(s).items():
        out[f"{name}/{k}"] = v
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "genalpha.npz"), **out)
print(f"wrote {len(out)} arrays")
