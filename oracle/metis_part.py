"""TEST INFRASTRUCTURE ONLY: k-way element partition by the serial METIS the reference vendors
(oracle/_ref/libsvmetis.so, built by `make -C oracle metis`; see oracle/metis_shim.c for what it stands in for)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsvmetis.so")
_lib = None


def have_metis() -> bool:
    return os.path.exists(SO)


def part_mesh_dual(IEN: np.ndarray, nNo: int, nparts: int, ncommon: int | None = None, seed: int = 10):
    """IEN (eNoN, nEl) 0-based -> (part[nEl] int32, edge cut).  ncommon defaults to the face size eNoNb (3 for TET4, 4 for
    HEX8) like SPLIT.c:81."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
    eNoN, nEl = IEN.shape
    if ncommon is None:
        ncommon = {4: 3, 8: 4, 10: 6}.get(eNoN, 3)
    if nparts == 1:
        return np.zeros(nEl, np.int32), 0
    ien = np.ascontiguousarray(IEN.T.astype(np.int32))
    epart = np.zeros(nEl, np.int32)
    npart = np.zeros(nNo, np.int32)
    cut = _lib.svmetis_part_mesh_dual(C.c_int(nEl), C.c_int(nNo), C.c_int(eNoN), ien.ctypes.data_as(C.c_void_p), C.c_int(ncommon),
                                      C.c_int(nparts), C.c_int(seed), epart.ctypes.data_as(C.c_void_p), npart.ctypes.data_as(C.c_void_p))
    if cut < 0:
        raise RuntimeError(f"METIS_PartMeshDual failed ({cut})")
    return epart, int(cut)
