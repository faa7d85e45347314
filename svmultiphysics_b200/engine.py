"""ctypes host layer over the C ABI (include/svb200.h -> lib/libsvb200.so).

``Engine`` mirrors, call for call, what the reference host does around its linear-algebra plugin
(Code/Source/solver/LinearAlgebra.h:13-37, FsilsLinearAlgebra.cpp:26-128) and its element loops
(eq_assem::global_eq_assem, Code/Source/solver/eq_assem.cpp:377-455):

    lhsa            -> Engine.lhsa(meshes)                     (solver/lhsa.cpp:126)
    fsils_lhs_create-> Engine.set_graph(rowPtr, colPtr, ...)   (linear_solver/lhs.cpp:30)
    fsils_bc_create -> Engine.set_face(...)                    (linear_solver/bc.cpp:18)
    ls_alloc        -> Engine.alloc(dof)                       (solver/ls.cpp:24)
    global_eq_assem -> Engine.assemble(iM, eq, domains)
    all_fun::commu  -> Engine.commu_R()                        (solver/all_fun.cpp:95)
    ls_solve        -> Engine.solve(...)                       (solver/ls.cpp:45 -> fsils_solve)

There is no CPU implementation behind this class: if the CUDA library is missing or no B200 is
visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsvb200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# every symbol include/svb200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "svb200_abi_version", "svb200_last_error", "svb200_create", "svb200_destroy",
    "svb200_comm_unique_id", "svb200_comm_init", "svb200_comm_transport",
    "svb200_set_graph", "svb200_lhsa_begin", "svb200_lhsa_add_mesh", "svb200_lhsa_finish", "svb200_lhsa_get",
    "svb200_set_mesh", "svb200_set_mesh_nxx", "svb200_set_coords", "svb200_set_num_faces", "svb200_set_face", "svb200_set_face_cap",
    "svb200_alloc", "svb200_set_state", "svb200_set_old_disp", "svb200_assemble", "svb200_assemble_host", "svb200_add_host_contrib", "svb200_commu_R", "svb200_ustruct_r", "svb200_set_ad", "svb200_get_ad",
    "svb200_solve", "svb200_download", "svb200_download_rows", "svb200_upload", "svb200_spmv", "svb200_last_timing",
    "svb200_host_register", "svb200_host_unregister", "svb200_timer_mark", "svb200_timer_elapsed",
    "svb200_bench_assemble", "svb200_bench_spmv", "svb200_measure_fp64_peak", "svb200_launch_count",
    "svb200_set_solution", "svb200_get_solution", "svb200_predictor", "svb200_initiator", "svb200_corrector",
    "svb200_set_node_flags", "svb200_set_dirichlet_rows", "svb200_dirichlet_ustruct", "svb200_advance_time_step",
    "svb200_set_bface", "svb200_assemble_neu",
    "svb200_last_host_stage", "svb200_set_uris", "svb200_set_ris", "svb200_set_mesh_thood", "svb200_thood_val_rc", "svb200_set_active_tension", "svb200_set_prestress", "svb200_get_prestress",
    "svb200_spmv_rc", "svb200_spmv_rc_variants", "svb200_bench_spmv_rc",
    "svb200_schur_sp", "svb200_schur_sp_variants", "svb200_bench_schur_sp",
]

_lib = None


class Svb200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def load_library():
    """Load libsvb200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Svb200Error(-1, f"{LIB_PATH} is missing: build it with `make -C svmultiphysics_b200/csrc` "
                                  "(or python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.svb200_last_error.restype = C.c_char_p
        lib.svb200_launch_count.restype = C.c_int64
        lib.svb200_comm_transport.restype = C.c_char_p
        lib.svb200_launch_count.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.asfortranarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.asfortranarray(a, dtype=np.int32)


class Engine:
    """One mesh partition on one B200 (one svb200_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        self.nNo = 0
        self.nnz = 0
        self.dof = 0
        self.meshes = []
        rc = self.lib.svb200_create(C.byref(self.h), C.c_int(device))
        if rc != 0:
            raise Svb200Error(rc, self.lib.svb200_last_error().decode())

    # ---- plumbing ------------------------------------------------------------------------------
    def _call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise Svb200Error(rc, self.lib.svb200_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.svb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launch_count(self) -> int:
        return int(self.lib.svb200_launch_count(self.h))

    # ---- multi-GPU -----------------------------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.svb200_comm_unique_id(buf)
        if rc != 0:
            raise Svb200Error(rc, lib.svb200_last_error().decode())
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        self._call("svb200_comm_init", C.c_int(nranks), C.c_int(rank), C.c_char_p(uid))

    def comm_transport(self) -> str:
        """"p2p" (own kernels over peer memory), "nccl" or "none"; valid after set_graph."""
        return self.lib.svb200_comm_transport(self.h).decode()

    # ---- structure -----------------------------------------------------------------------------
    def lhsa(self, nNo: int, IENs):
        """Sparse structure of all meshes (lhsa_ns::lhsa): returns (rowPtr, colPtr)."""
        self._call("svb200_lhsa_begin", C.c_int32(nNo))
        for IEN in IENs:
            IEN = _i32(IEN)
            self._call("svb200_lhsa_add_mesh", C.c_int32(IEN.shape[0]), C.c_int32(IEN.shape[1]), _i(IEN))
        nnz = C.c_int32(0)
        self._call("svb200_lhsa_finish", C.byref(nnz))
        rowPtr = np.zeros(nNo + 1, dtype=np.int32)
        colPtr = np.zeros(max(nnz.value, 1), dtype=np.int32)
        self._call("svb200_lhsa_get", _i(rowPtr), _i(colPtr))
        return rowPtr, colPtr[:nnz.value]

    def set_graph(self, rowPtr, colPtr, mynNo=None, node_map=None, neighbours=None):
        """neighbours: list of (rank, ptr) with ptr = FSILS-order local ids shared with that rank."""
        rowPtr, colPtr = _i32(rowPtr), _i32(colPtr)
        self.nNo = len(rowPtr) - 1
        self.nnz = len(colPtr)
        node_map = _i32(node_map)
        neighbours = neighbours or []
        ranks = np.array([r for r, _ in neighbours], dtype=np.int32)
        counts = np.array([len(p) for _, p in neighbours], dtype=np.int32)
        ptrs = np.concatenate([np.asarray(p, dtype=np.int32) for _, p in neighbours]) if neighbours else np.zeros(0, np.int32)
        self._call("svb200_set_graph", C.c_int32(self.nNo), C.c_int32(self.nnz), _i(rowPtr), _i(colPtr),
                   C.c_int32(self.nNo if mynNo is None else mynNo), _i(node_map),
                   C.c_int32(len(neighbours)), _i(ranks), _i(counts), _i(np.ascontiguousarray(ptrs)))

    def set_mesh(self, iM, IEN, w, N, Nx, eId=None, nFn=0, fN=None, Nxx=None):
        IEN, eId, fN = _i32(IEN), _i32(eId), _f64(fN)
        w, N, Nx = _f64(w), _f64(N), _f64(Nx)
        self._call("svb200_set_mesh", C.c_int32(iM), C.c_int32(IEN.shape[0]), C.c_int32(IEN.shape[1]), _i(IEN), _i(eId),
                   C.c_int32(nFn), _d(fN), C.c_int32(len(w)), _d(w), _d(N), _d(Nx))
        if Nxx is not None:      # fs[0].Nxx(6,eNoN,nG): second derivatives for nn::gn_nxx (fluid on non-linear elements)
            Nxx = _f64(Nxx)
            assert Nxx.shape == (6, IEN.shape[0], len(w))
            self._call("svb200_set_mesh_nxx", C.c_int32(iM), _d(Nxx))
        while len(self.meshes) <= iM:
            self.meshes.append(None)
        self.meshes[iM] = (IEN.shape[0], IEN.shape[1])

    def set_coords(self, x):
        x = _f64(x)
        assert x.shape == (3, self.nNo)
        self._call("svb200_set_coords", _d(x))

    def set_num_faces(self, n):
        self._call("svb200_set_num_faces", C.c_int32(n))

    def set_face(self, faIn, bGrp, glob, val, shared=0):
        glob, val = _i32(glob), _f64(val)
        self._call("svb200_set_face", C.c_int32(faIn), C.c_int32(bGrp), C.c_int32(val.shape[0]), C.c_int32(len(glob)),
                   _i(glob), _d(val), C.c_int32(shared))

    def set_face_cap(self, faIn, cap_glob, cap_val):
        """Capping surface of a coupled face: cap node ids (input order, negative = not on this partition), cap_val(face_dof, n)."""
        cap_glob, cap_val = _i32(cap_glob), _f64(cap_val)
        self._call("svb200_set_face_cap", C.c_int32(faIn), C.c_int32(len(cap_glob)), _i(cap_glob), _d(cap_val))

    # ---- per Newton iteration ------------------------------------------------------------------
    def alloc(self, dof):
        self.dof = dof
        self._call("svb200_alloc", C.c_int32(dof))

    def set_state(self, Ag, Yg, Dg=None, Bf=None):
        Ag, Yg, Dg, Bf = _f64(Ag), _f64(Yg), _f64(Dg), _f64(Bf)
        tDof = (Ag if Ag is not None else Yg).shape[0]
        self._call("svb200_set_state", C.c_int32(tDof), _d(Ag), _d(Yg), _d(Dg), _d(Bf))

    def set_prestress(self, pS0):
        """Nodal prestress com_mod.pS0 (6, nNo), Voigt 11,22,33,12,23,31; None removes it."""
        self._call("svb200_set_prestress", _d(_f64(pS0)))

    def get_prestress(self):
        """(pSn (6, nNo), pSa (nNo)) accumulated by the last assembly of a prestress equation (abi.EQ_PRESTRESS)."""
        pSn, pSa = np.zeros((6, self.nNo), order="F"), np.zeros(self.nNo)
        self._call("svb200_get_prestress", _d(pSn), _d(pSa))
        return pSn, pSa

    def set_active_tension(self, Ya_f, Ya_s=None, Ya_n=None):
        """Nodal active tensions cep_mod.cem.Ya_f / Ya_s / Ya_n (nNo each) for domains with an active-stress model."""
        f = np.ascontiguousarray(Ya_f, dtype=np.float64)
        s_ = None if Ya_s is None else np.ascontiguousarray(Ya_s, dtype=np.float64)
        n_ = None if Ya_n is None else np.ascontiguousarray(Ya_n, dtype=np.float64)
        self._call("svb200_set_active_tension", _d(f), _d(s_), _d(n_))

    def set_mesh_thood(self, iM, t):
        """Taylor-Hood function spaces for mesh iM (svb200_set_mesh_thood); t: the dict of fs::get_thood_fs tables (eNoNq, nG2, lShpF_q,
        Nq1, Nqxi1, w2, Nw2, Nwxi2, Nq2, Nqxi2, column-major).  t = None returns the mesh to equal-order spaces."""
        if t is None:
            self._call("svb200_set_mesh_thood", C.c_int32(iM), C.c_int32(0), C.c_int32(0), C.c_int32(0), None, None, None, None, None, None, None)
            return
        a = {k: _f64(t[k]) for k in ("Nq1", "Nqxi1", "w2", "Nw2", "Nwxi2", "Nq2", "Nqxi2")}
        self._call("svb200_set_mesh_thood", C.c_int32(iM), C.c_int32(int(t["eNoNq"])), C.c_int32(int(t["nG2"])), C.c_int32(int(t["lShpF_q"])),
                   _d(a["Nq1"]), _d(a["Nqxi1"]), _d(a["w2"]), _d(a["Nw2"]), _d(a["Nwxi2"]), _d(a["Nq2"]), _d(a["Nqxi2"]))

    def thood_val_rc(self):
        """fs::thood_val_rc: pressure rows of the nodes that carry no pressure dof (call after the assembly and the boundary terms)."""
        self._call("svb200_thood_val_rc")

    def set_ris(self, maps, closed):
        """Fitted RIS surfaces (svb200_set_ris): maps = list of (2, n) int arrays (grisMapList[p].map), closed = RIS.clsFlg.  An empty list
        removes the plan."""
        n = len(maps)
        nMap = np.array([mp.shape[1] for mp in maps], dtype=np.int32)
        flat = np.concatenate([np.asfortranarray(mp, dtype=np.int32).ravel(order="F") for mp in maps]) if n else np.zeros(0, np.int32)
        flat = np.ascontiguousarray(flat, dtype=np.int32)
        cl = np.array([int(x) for x in closed], dtype=np.int32)
        self._call("svb200_set_ris", C.c_int32(n), _i(nMap) if n else None, _i(flat) if n else None, _i(cl) if n else None)

    def set_uris(self, valves, sdf=None, scaffold_udf=None, valve_vel=None):
        """URIS valves (svb200_set_uris): valves = list of abi.Uris; sdf / scaffold_udf: (nUris, nNo); valve_vel: (nUris, nNo, 3).
        An empty list removes them."""
        n = len(valves)
        arr = (abi.Uris * max(n, 1))(*valves)
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        sdf, scaffold_udf, valve_vel = c(sdf), c(scaffold_udf), c(valve_vel)
        self._keep_uris = (arr, sdf, scaffold_udf, valve_vel)
        self._call("svb200_set_uris", C.c_int32(n), arr, _d(sdf), _d(scaffold_udf), _d(valve_vel))

    def set_old_disp(self, Do):
        Do = _f64(Do)
        self._call("svb200_set_old_disp", C.c_int32(Do.shape[0]), _d(Do))

    # ---- generalised-alpha state on the device (Integrator::predictor / initiator / corrector) -----------
    def set_solution(self, which, A=None, Y=None, D=None):
        A, Y, D = _f64(A), _f64(Y), _f64(D)
        tDof = next(a for a in (A, Y, D) if a is not None).shape[0]
        self._call("svb200_set_solution", C.c_int32(tDof), C.c_int32(which), _d(A), _d(Y), _d(D))
        self.tDof = tDof

    def get_solution(self, which):
        out = [np.zeros((self.tDof, self.nNo), order="F") for _ in range(3)]
        self._call("svb200_get_solution", C.c_int32(which), _d(out[0]), _d(out[1]), _d(out[2]))
        return out

    def predictor(self, eqs, dt, dFlag):
        arr = (abi.EqTime * len(eqs))(*eqs)
        self._call("svb200_predictor", C.c_int32(len(eqs)), arr, C.c_double(dt), C.c_int32(int(dFlag)))

    def initiator(self, eqs):
        arr = (abi.EqTime * len(eqs))(*eqs)
        self._call("svb200_initiator", C.c_int32(len(eqs)), arr)

    def corrector(self, eq, dt, mesh_s=-1):
        self._call("svb200_corrector", C.byref(eq), C.c_double(dt), C.c_int32(mesh_s))

    def set_node_flags(self, flags):
        self._call("svb200_set_node_flags", _i(_i32(flags)))

    def set_dirichlet_rows(self, row0, nodes, valA=None, valY=None, valD=None):
        nodes = _i32(nodes)
        valA, valY, valD = _f64(valA), _f64(valY), _f64(valD)
        nrow = next(v for v in (valA, valY, valD) if v is not None).shape[0]
        self._call("svb200_set_dirichlet_rows", C.c_int32(row0), C.c_int32(nrow), C.c_int32(len(nodes)), _i(nodes),
                   _d(valA), _d(valY), _d(valD))

    def dirichlet_ustruct(self, eq, dt, nodes, dir_mask=7, impD=False):
        """set_bc_dir's update of (Dn, Ad) / (An, Ad) on the Dirichlet nodes of a ustruct equation (set_bc.cpp:1046-1117)."""
        nodes = _i32(nodes)
        self._call("svb200_dirichlet_ustruct", C.byref(eq), C.c_double(dt), C.c_int32(len(nodes)), _i(nodes), C.c_int32(dir_mask),
                   C.c_int32(int(impD)))

    def advance_time_step(self):
        self._call("svb200_advance_time_step")

    def assemble(self, iM, eq: abi.EqParams, dmns):
        arr = (abi.DmnParams * len(dmns))(*dmns)
        self._call("svb200_assemble", C.c_int32(iM), C.byref(eq), arr, C.c_int32(len(dmns)))

    def assemble_host(self, iM, eq: abi.EqParams, dmns, Ag, Yg, R_out=None):
        """set_state + alloc + assemble + commu_R + download(R) in one pipelined call (host-resident state)."""
        arr = (abi.DmnParams * len(dmns))(*dmns)
        assert Ag.flags.f_contiguous and Yg.flags.f_contiguous and (R_out is None or R_out.flags.f_contiguous)
        self._call("svb200_assemble_host", C.c_int32(iM), C.byref(eq), arr, C.c_int32(len(dmns)), _d(Ag), _d(Yg), _d(R_out))

    def set_bface(self, iFa, iM, IENb, gE, w, N, Nx):
        IENb, gE = _i32(np.asfortranarray(IENb)), _i32(gE)
        w, N, Nx = _f64(w), _f64(N), _f64(Nx)
        self._call("svb200_set_bface", C.c_int32(iFa), C.c_int32(iM), C.c_int32(IENb.shape[0]), C.c_int32(IENb.shape[1]),
                   _i(IENb), _i(gE), C.c_int32(len(w)), _d(w), _d(N), _d(Nx))

    def assemble_neu(self, iFa, eq: abi.EqParams, dmns, hg):
        arr = (abi.DmnParams * len(dmns))(*dmns)
        self._call("svb200_assemble_neu", C.c_int32(iFa), C.byref(eq), arr, C.c_int32(len(dmns)), _d(_f64(hg)))

    def add_host_contrib(self, dof, rows=None, R_add=None, krows=None, kcols=None, K_add=None):
        rows, krows, kcols = _i32(rows), _i32(krows), _i32(kcols)
        R_add, K_add = _f64(R_add), _f64(K_add)
        self._call("svb200_add_host_contrib", C.c_int32(dof), C.c_int32(0 if rows is None else len(rows)), _i(rows),
                   _d(R_add), C.c_int32(0 if krows is None else len(krows)), _i(krows), _i(kcols), _d(K_add))

    def commu_R(self):
        self._call("svb200_commu_R")

    def solve(self, dof, ls_type, ls: abi.LsParams, incL=None, res=None, hist_cap=0, want_solution=True,
              prec=abi.PREC_FSILS):
        nFaces = 0 if incL is None else len(incL)
        incL, res = _i32(incL), _f64(res)
        out = abi.LsResult()
        hist = np.zeros(max(hist_cap, 1))
        out.hist = hist.ctypes.data_as(_dp)
        out.hist_cap = hist_cap
        X = np.zeros((dof, self.nNo), order="F") if want_solution else None
        self._call("svb200_solve", C.c_int32(dof), C.c_int32(ls_type), C.c_int32(prec), C.byref(ls),
                   C.c_int32(nFaces), _i(incL), _d(res), _d(X), C.byref(out))
        return X, out, hist[:out.hist_n].copy()

    # ---- debug / parity / bench ----------------------------------------------------------------
    def get_R(self):
        R = np.zeros((self.dof, self.nNo), order="F")
        self._call("svb200_download", C.c_int32(abi.ARRAY_R), _d(R))
        return R

    def last_host_stage(self):
        """Timeline (ms) of the last pipelined assemble_host: uploads, kernels, shared-node sum, streamed D2H, end."""
        t = (C.c_double * 8)()
        self._call("svb200_last_host_stage", t)
        return dict(zip(("uploads_done", "kernels_done", "halo_done", "streamed_rows_on_host", "end", "host_enqueued", "host_drained",
                         "host_return"), (round(float(v), 4) for v in t)))

    def download_into(self, what, dst):
        """svb200_download into a caller-owned (e.g. page-locked) array."""
        self._call("svb200_download", C.c_int32(what), _d(dst))

    def get_Val(self):
        V = np.zeros((self.dof * self.dof, self.nnz), order="F")
        self._call("svb200_download", C.c_int32(abi.ARRAY_VAL), _d(V))
        return V

    def get_Kd(self):
        """com_mod.Kd(12, nnz) left by a ustruct assembly."""
        K = np.zeros((12, self.nnz), order="F")
        self._call("svb200_download", C.c_int32(abi.ARRAY_KD), _d(K))
        return K

    def get_Rd(self):
        """com_mod.Rd(3, nNo) as the last ustruct_r left it."""
        Rd = np.zeros((3, self.nNo), order="F")
        self._call("svb200_download", C.c_int32(abi.ARRAY_RD), _d(Rd))
        return Rd

    def ustruct_r(self, eq: abi.EqParams, itr, Ad=None):
        """Ad = None: use the device-resident Ad (set_ad / predictor / corrector keep it)."""
        self._call("svb200_ustruct_r", C.byref(eq), C.c_int32(itr), _d(_f64(Ad)) if Ad is not None else None)

    def set_ad(self, Ad):
        self._call("svb200_set_ad", _d(_f64(Ad)))

    def get_ad(self):
        Ad = np.zeros((3, self.nNo), order="F")
        self._call("svb200_get_ad", _d(Ad))
        return Ad

    def get_rows(self, what, nodes, rowPtr=None):
        """Rows of R / W (-> (dof, n)) or CSR rows of Val (-> (dof*dof, sum of row lengths); needs the caller's rowPtr) of the
        listed input-order nodes."""
        nodes = _i32(nodes)
        if what in (abi.ARRAY_VAL, abi.ARRAY_KD):
            d2 = self.dof * self.dof if what == abi.ARRAY_VAL else 12
            cnt = int((np.asarray(rowPtr)[nodes + 1] - np.asarray(rowPtr)[nodes]).sum())
        else:
            d2, cnt = self.dof, len(nodes)
        out = np.zeros((d2, cnt), order="F")
        self._call("svb200_download_rows", C.c_int32(what), C.c_int32(len(nodes)), _i(nodes), _d(out))
        return out

    def get_W(self):
        W = np.zeros((self.dof, self.nNo), order="F")
        self._call("svb200_download", C.c_int32(abi.ARRAY_W), _d(W))
        return W

    def put_R(self, R):
        R = _f64(R)
        self._call("svb200_upload", C.c_int32(abi.ARRAY_R), C.c_int32(R.shape[0]), _d(R))

    def put_Rd(self, Rd):
        """com_mod.Rd(3, nNo) (normally left on the device by ustruct_r)."""
        self._call("svb200_upload", C.c_int32(abi.ARRAY_RD), C.c_int32(self.dof), _d(_f64(Rd)))

    def put_Val(self, V, dof):
        V = _f64(V)
        self._call("svb200_upload", C.c_int32(abi.ARRAY_VAL), C.c_int32(dof), _d(V))

    def spmv(self, dof, U):
        U = _f64(U)
        KU = np.zeros_like(U, order="F")
        self._call("svb200_spmv", C.c_int32(dof), _d(U), _d(KU))
        return KU

    def last_timing(self):
        a, s = C.c_double(0), C.c_double(0)
        self._call("svb200_last_timing", C.byref(a), C.byref(s))
        return a.value, s.value

    def bench_assemble(self, iM, eq, dmns, reps):
        arr = (abi.DmnParams * len(dmns))(*dmns)
        ms = C.c_double(0)
        self._call("svb200_bench_assemble", C.c_int32(iM), C.byref(eq), arr, C.c_int32(len(dmns)), C.c_int32(reps), C.byref(ms))
        return ms.value

    def bench_spmv(self, dof, reps):
        ms = C.c_double(0)
        self._call("svb200_bench_spmv", C.c_int32(dof), C.c_int32(reps), C.byref(ms))
        return ms.value

    # ---- rectangular-block products / Schur operator with caller-supplied matrices (tests, A/B) ----
    def spmv_rc(self, R, Cc, K, U, variant=-1):
        """KU = K U with (R x Cc) blocks on the context's graph: K (R*Cc, nnz), U (Cc, nNo) -> KU (R, nNo)."""
        K, U = _f64(K), _f64(U)
        KU = np.zeros((R, self.nNo), order="F")
        self._call("svb200_spmv_rc", C.c_int32(R), C.c_int32(Cc), C.c_int32(variant), _d(K), _d(U), _d(KU))
        return KU

    def spmv_rc_variants(self, R, Cc) -> int:
        return int(self.lib.svb200_spmv_rc_variants(C.c_int32(R), C.c_int32(Cc)))

    def bench_spmv_rc(self, R, Cc, variant, reps):
        ms = C.c_double(0)
        self._call("svb200_bench_spmv_rc", C.c_int32(R), C.c_int32(Cc), C.c_int32(variant), C.c_int32(reps), C.byref(ms))
        return ms.value

    def schur_sp(self, L, Gt, P, GP, variant=-1):
        """SP = L p - Gt (G p) (cgrad::schur) and <p, SP>: L (nnz), Gt (3, nnz), P (nNo), GP (3, nNo)."""
        L, Gt, P, GP = _f64(L), _f64(Gt), _f64(P), _f64(GP)
        SP = np.zeros(self.nNo)
        dot = C.c_double(0)
        self._call("svb200_schur_sp", C.c_int32(variant), _d(L), _d(Gt), _d(P), _d(GP), _d(SP), C.byref(dot))
        return SP, dot.value

    def schur_sp_variants(self) -> int:
        return int(self.lib.svb200_schur_sp_variants())

    def bench_schur_sp(self, variant, reps):
        ms = C.c_double(0)
        self._call("svb200_bench_schur_sp", C.c_int32(variant), C.c_int32(reps), C.byref(ms))
        return ms.value

    def pin(self, arr: np.ndarray):
        """Page-lock a numpy buffer that is repeatedly copied to / from the device."""
        self._call("svb200_host_register", C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes))

    def unpin(self, arr: np.ndarray):
        self._call("svb200_host_unregister", C.c_void_p(arr.ctypes.data))

    def timer_mark(self, which: int):
        self._call("svb200_timer_mark", C.c_int32(which))

    def timer_elapsed(self) -> float:
        ms = C.c_double(0)
        self._call("svb200_timer_elapsed", C.byref(ms))
        return ms.value

    def fp64_peak(self):
        t = C.c_double(0)
        self._call("svb200_measure_fp64_peak", C.byref(t))
        return t.value
