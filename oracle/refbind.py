"""ctypes bindings of the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``RefCase``    -> oracle/_ref/libsvref.so : the unmodified reference hot path (oracle/ref_harness.cpp)
* ``OracleCase`` -> oracle/libsvoracle.so  : our C restatement of the same algorithms (oracle/sv_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (svmultiphysics_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from svmultiphysics_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libsvref.so")
ORACLE_SO = os.path.join(_HERE, "libsvoracle.so")
HOST_SO = os.path.join(os.path.dirname(_HERE), "svmultiphysics_b200", "lib", "libsvb200_host.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.asfortranarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.asfortranarray(a, dtype=np.int32)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_host() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(HOST_SO)


class _FlatCase:
    """Shared driver for both checkers: they export the same flat entry points under a prefix."""
    prefix = ""
    so_path = ""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = C.CDLL(cls.so_path)
            p = cls.prefix
            getattr(lib, p + "create").restype = C.c_void_p
            getattr(lib, p + "last_error").restype = C.c_char_p
            for name in ("destroy", "set_coords", "add_mesh", "get_mesh_tables", "build_graph", "get_graph",
                         "set_face", "alloc", "set_state", "set_old_disp", "assemble", "get", "put", "solve", "spmv", "last_timing"):
                getattr(lib, p + name).argtypes = None
            cls._lib = lib
        return cls._lib

    def _call(self, name, *args):
        fn = getattr(self.lib(), self.prefix + name)
        rc = fn(C.c_void_p(self.h), *args)
        if rc != 0:
            raise RuntimeError(f"{self.prefix}{name}: {getattr(self.lib(), self.prefix + 'last_error')().decode()}")

    def __init__(self):
        self.h = getattr(self.lib(), self.prefix + "create")()
        self.nNo = 0
        self.nnz = 0
        self.meshes = []

    def close(self):
        if self.h:
            getattr(self.lib(), self.prefix + "destroy")(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_coords(self, x):
        x = _f64(x)
        self.nNo = x.shape[1]
        self._call("set_coords", C.c_int(self.nNo), _d(x))

    def add_mesh(self, IEN, eId=None, nFn=0, fN=None):
        IEN = _i32(IEN)
        eId = _i32(eId)
        fN = _f64(fN)
        self._call("add_mesh", C.c_int(IEN.shape[0]), C.c_int(IEN.shape[1]), _i(IEN), _i(eId), C.c_int(nFn), _d(fN))
        self.meshes.append((IEN.shape[0], IEN.shape[1]))
        return len(self.meshes) - 1

    def mesh_tables(self, iM):
        eNoN = self.meshes[iM][0]
        nG = C.c_int(0)
        self._call("get_mesh_tables", C.c_int(iM), C.byref(nG), None, None, None)
        w = np.zeros(nG.value)
        N = np.zeros((eNoN, nG.value), order="F")
        Nx = np.zeros((3, eNoN, nG.value), order="F")
        self._call("get_mesh_tables", C.c_int(iM), C.byref(nG), _d(w), _d(N), _d(Nx))
        return w, N, Nx

    def mesh_nxx(self, iM):
        """fs[0].Nxx(6,eNoN,nG) of the reference (compiled reference only)."""
        eNoN = self.meshes[iM][0]
        w, _, _ = self.mesh_tables(iM)
        Nxx = np.zeros((6, eNoN, len(w)), order="F")
        self._call("get_mesh_nxx", C.c_int(iM), _d(Nxx))
        return Nxx

    def build_graph(self, nFaces=0):
        nnz = C.c_int(0)
        self._call("build_graph", C.c_int(nFaces), C.byref(nnz))
        self.nnz = nnz.value
        rowPtr = np.zeros(self.nNo + 1, dtype=np.int32)
        colPtr = np.zeros(self.nnz, dtype=np.int32)
        self._call("get_graph", _i(rowPtr), _i(colPtr))
        return rowPtr, colPtr

    def set_face(self, faIn, bGrp, glob, val):
        glob = _i32(glob)
        val = _f64(val)
        self._call("set_face", C.c_int(faIn), C.c_int(bGrp), C.c_int(val.shape[0]), C.c_int(len(glob)), _i(glob), _d(val))

    def alloc(self, dof):
        self.dof = dof
        self._call("alloc", C.c_int(dof))

    def set_state(self, Ag, Yg, Dg=None, Bf=None):
        Ag, Yg, Dg, Bf = _f64(Ag), _f64(Yg), _f64(Dg), _f64(Bf)
        self._call("set_state", C.c_int(Ag.shape[0]), _d(Ag), _d(Yg), _d(Dg), _d(Bf))

    def set_prestress(self, pS0):
        self._call("set_prestress", _d(_f64(pS0)))

    def get_prestress(self):
        pSn, pSa = np.zeros((6, self.nNo), order="F"), np.zeros(self.nNo)
        self._call("get_prestress", _d(pSn), _d(pSa))
        return pSn, pSa

    def set_active_tension(self, Ya_f, Ya_s=None, Ya_n=None):
        f = np.ascontiguousarray(Ya_f, dtype=np.float64)
        s_ = None if Ya_s is None else np.ascontiguousarray(Ya_s, dtype=np.float64)
        n_ = None if Ya_n is None else np.ascontiguousarray(Ya_n, dtype=np.float64)
        self._call("set_active_tension", _d(f), _d(s_), _d(n_))

    def set_mesh_thood(self, iM):
        """Taylor-Hood function spaces for mesh iM (mshType::nFs = 2)."""
        self._call("set_mesh_thood", C.c_int(iM))

    def thood_tables(self, iM):
        """dict with eNoNq, nG1, nG2, lShpF_w, lShpF_q and the tables Nq1(eNoNq,nG1), Nqxi1(3,eNoNq,nG1), w2, Nw2(eNoN,nG2), Nwxi2(3,eNoN,nG2),
        Nq2(eNoNq,nG2), Nqxi2(3,eNoNq,nG2) of fs::get_thood_fs."""
        dims = np.zeros(5, np.int32)
        z = None
        self._call("get_thood_tables", C.c_int(iM), _i(dims), z, z, z, z, z, z, z)
        q, g1, g2 = int(dims[0]), int(dims[1]), int(dims[2])
        eNoN = self.meshes[iM][0]
        t = dict(eNoNq=q, nG1=g1, nG2=g2, lShpF_w=int(dims[3]), lShpF_q=int(dims[4]),
                 Nq1=np.zeros((q, g1), order="F"), Nqxi1=np.zeros((3, q, g1), order="F"), w2=np.zeros(g2),
                 Nw2=np.zeros((eNoN, g2), order="F"), Nwxi2=np.zeros((3, eNoN, g2), order="F"),
                 Nq2=np.zeros((q, g2), order="F"), Nqxi2=np.zeros((3, q, g2), order="F"))
        self._call("get_thood_tables", C.c_int(iM), _i(dims), _d(t["Nq1"]), _d(t["Nqxi1"]), _d(t["w2"]), _d(t["Nw2"]), _d(t["Nwxi2"]),
                   _d(t["Nq2"]), _d(t["Nqxi2"]))
        return t

    def thood_val_rc(self):
        self._call("thood_val_rc")

    def set_ris(self, maps, closed, meshes):
        """Fitted RIS (before build_graph): maps = list of (2, n) int arrays (grisMapList[p].map), closed = RIS.clsFlg, meshes = list of
        (mesh of face 0, mesh of face 1)."""
        n = len(maps)
        nMap = np.array([mp.shape[1] for mp in maps], dtype=np.int32)
        flat = np.concatenate([np.asfortranarray(mp, dtype=np.int32).ravel(order="F") for mp in maps]) if n else np.zeros(0, np.int32)
        flat = np.ascontiguousarray(flat, dtype=np.int32)
        cl = np.array([int(x) for x in closed], dtype=np.int32)
        ms = np.ascontiguousarray(np.array(meshes, dtype=np.int32).reshape(-1))
        self._call("set_ris", C.c_int(n), _i(nMap), _i(flat), _i(cl), _i(ms))

    def set_uris(self, raw, sdf=None, scaffold_udf=None, valve_vel=None):
        """com_mod.uris[] for uris::eval_uris_ris_factors_quadrature.  raw = list of dicts with resistance, sdf_deps, sdf_deps_close,
        clsFlg, cnt, n_open, n_close, scaffold, include_velocity (the urisType members, not the effective thickness);
        sdf / scaffold_udf: (nUris, nNo); valve_vel: (nUris, nNo, 3)."""
        n = len(raw)
        scal = np.array([[u["resistance"], u["sdf_deps"], u["sdf_deps_close"]] for u in raw], dtype=np.float64).reshape(-1)
        flags = np.array([[int(u["clsFlg"]), u["cnt"], u["n_open"], u["n_close"], int(u["scaffold"]), int(u["include_velocity"])]
                          for u in raw], dtype=np.int32).reshape(-1)
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self._call("set_uris", C.c_int(n), _d(c(scal)) if n else None, _i(flags) if n else None, _d(c(sdf)), _d(c(scaffold_udf)),
                   _d(c(valve_vel)))

    def set_old_disp(self, Do):
        Do = _f64(Do)
        self._call("set_old_disp", C.c_int(Do.shape[0]), _d(Do))

    def assemble(self, iM, eq: abi.EqParams, dmns):
        arr = (abi.DmnParams * len(dmns))(*dmns)
        self._call("assemble", C.c_int(iM), C.byref(eq), arr, C.c_int(len(dmns)))

    # ---- boundary faces (compiled reference only) -------------------------------------------------------
    def add_face(self, iM, IENb, gE):
        IENb, gE = _i32(np.asfortranarray(IENb)), _i32(gE)
        self._call("add_face", C.c_int(iM), C.c_int(IENb.shape[0]), C.c_int(IENb.shape[1]), _i(IENb), _i(gE))
        self.faces = getattr(self, "faces", []) + [IENb.shape[0]]
        return len(self.faces) - 1

    def face_tables(self, iM, iFa):
        eNoNb = self.faces[iFa]
        nG = C.c_int(0)
        self._call("get_face_tables", C.c_int(iM), C.c_int(iFa), C.byref(nG), None, None, None)
        w = np.zeros(nG.value)
        N = np.zeros((eNoNb, nG.value), order="F")
        Nx = np.zeros((2, eNoNb, nG.value), order="F")
        self._call("get_face_tables", C.c_int(iM), C.c_int(iFa), C.byref(nG), _d(w), _d(N), _d(Nx))
        return w, N, Nx

    def assemble_neu(self, iM, iFa, eq: abi.EqParams, dmns, hg):
        arr = (abi.DmnParams * len(dmns))(*dmns)
        hg = _f64(hg)
        self._call("assemble_neu", C.c_int(iM), C.c_int(iFa), C.byref(eq), arr, C.c_int(len(dmns)), _d(hg))

    def get_R(self):
        R = np.zeros((self.dof, self.nNo), order="F")
        self._call("get", C.c_int(abi.ARRAY_R), _d(R))
        return R

    def get_Val(self):
        V = np.zeros((self.dof * self.dof, self.nnz), order="F")
        self._call("get", C.c_int(abi.ARRAY_VAL), _d(V))
        return V

    def put_R(self, R):
        R = _f64(R)
        self.dof = R.shape[0]
        self._call("put", C.c_int(abi.ARRAY_R), C.c_int(self.dof), _d(R))

    def put_Val(self, V, dof):
        V = _f64(V)
        self.dof = dof
        self._call("put", C.c_int(abi.ARRAY_VAL), C.c_int(dof), _d(V))

    def solve(self, dof, ls_type, ls: abi.LsParams, incL=None, res=None, hist_cap=0, prec=abi.PREC_FSILS):
        nFaces = 0 if incL is None else len(incL)
        incL = _i32(incL)
        res = _f64(res)
        out = abi.LsResult()
        hist = np.zeros(max(hist_cap, 1))
        out.hist = hist.ctypes.data_as(_dp)
        out.hist_cap = hist_cap
        X = np.zeros((dof, self.nNo), order="F")
        self._call("solve", C.c_int(dof), C.c_int(ls_type), C.c_int(prec), C.byref(ls), C.c_int(nFaces),
                   _i(incL), _d(res), _d(X), C.byref(out))
        return X, out, hist[:out.hist_n].copy()

    def spmv(self, dof, U):
        U = _f64(U)
        KU = np.zeros_like(U, order="F")
        self._call("spmv", C.c_int(dof), _d(U), _d(KU))
        return KU

    def last_timing(self):
        a, s = C.c_double(0), C.c_double(0)
        self._call("last_timing", C.byref(a), C.byref(s))
        return a.value, s.value


class RefCase(_FlatCase):
    prefix = "svref_"
    so_path = REF_SO
    _lib = None
    _host = None

    def use_b200_backend(self, device=0, scatter=0):
        """Swap FsilsLinearAlgebra for the product's C++ host layer B200LinearAlgebra (svmultiphysics_b200/host/):
        the reference's own objects (ComMod, eqType, mshType, FSILS_lhsType) then drive libsvb200.so through the
        LinearAlgebra plug-in interface and the global_eq_assem early-out, exactly as in INTEGRATION.md."""
        cls = type(self)
        if cls._host is None:
            C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)          # the host application's symbols (ComMod, Array, LinearAlgebra ...)
            host = C.CDLL(HOST_SO)
            host.b200host_new.restype = C.c_void_p
            host.b200host_launch_count.restype = C.c_longlong
            cls._host = host
        host = cls._host
        self.backend = host.b200host_new(C.c_int(device), C.c_int(scatter))
        self._call("set_backend", C.c_void_p(self.backend), C.cast(host.b200host_global_eq_assem, C.c_void_p),
                   C.cast(host.b200host_download, C.c_void_p))
        self._call("set_backend_ustruct_r", C.cast(host.b200host_ustruct_r, C.c_void_p))
        self._call("set_backend_thood_val_rc", C.cast(host.b200host_thood_val_rc, C.c_void_p))

    def set_partition(self, gnNo, ltg):
        """Multi-rank runs (SVREF_MPI_SIZE > 1 in the environment before this library is loaded): local -> global node map."""
        ltg = _i32(ltg)
        self._call("set_partition", C.c_int(gnNo), C.c_int(len(ltg)), _i(ltg))

    def get_lhs(self):
        """mynNo, lhs.map(nNo) and [(iP, ptr)] as fsils_lhs_create built them (linear_solver/lhs.cpp:30-348)."""
        mynNo, nReq = C.c_int(0), C.c_int(0)
        mp = np.zeros(self.nNo, dtype=np.int32)
        self._call("get_lhs", C.byref(mynNo), C.byref(nReq), _i(mp))
        reqs = []
        for i in range(nReq.value):
            iP, n = C.c_int(0), C.c_int(0)
            self._call("get_lhs_req", C.c_int(i), C.byref(iP), C.byref(n), None)
            ptr = np.zeros(n.value, dtype=np.int32)
            self._call("get_lhs_req", C.c_int(i), C.byref(iP), C.byref(n), _i(ptr))
            reqs.append((iP.value, ptr))
        return mynNo.value, mp, reqs

    def set_face_cap(self, faIn, cap_glob, cap_val):
        """lhs.face[faIn].{has_cap, cap_glob, cap_val}: the capping surface of a coupled face (fils_struct.hpp:131-143)."""
        cap_glob, cap_val = _i32(cap_glob), _f64(cap_val)
        self._call("set_face_cap", C.c_int(faIn), C.c_int(len(cap_glob)), _i(cap_glob), _d(cap_val))

    def barrier(self):
        """MPI_Barrier of the multi-rank shim (no-op for one rank)."""
        self._call("barrier")

    def commu_R(self):
        """all_fun::commu(com_mod, R): shared-node sum of the residual (solver/Integrator.cpp:124-129)."""
        self._call("commu_R")

    def get_Kd(self):
        """com_mod.Kd(12, nnz): displacement tangent of the ustruct equation (solver/ustruct.cpp:1621)."""
        K = np.zeros((12, self.nnz), order="F")
        self._call("get", C.c_int(abi.ARRAY_KD), _d(K))
        return K

    def ustruct_r(self, itr, Ad):
        """ustruct::ustruct_r (solver/ustruct.cpp:1742-1845) on the assembled R / Kd."""
        self._call("ustruct_r", C.c_int(itr), _d(_f64(Ad)))

    def backend_launch_count(self):
        return int(type(self)._host.b200host_launch_count(C.c_void_p(self.backend)))

    def close(self):
        super().close()
        if getattr(self, "backend", None):
            type(self)._host.b200host_delete(C.c_void_p(self.backend))
            self.backend = None


class OracleCase(_FlatCase):
    prefix = "svorc_"
    so_path = ORACLE_SO
    _lib = None


class GenAlphaRef:
    """The reference's own Integrator (predictor / initiator / corrector) and set_bc::set_bc_dir on a hand-filled Simulation
    (oracle/ref_harness_genalpha.cpp; Code/Source/solver/Integrator.cpp:393-1070, set_bc.cpp:901-1182).  Arrays are (tDof, nNo)."""

    def __init__(self, eqs, dt, dFlag, sstEq, Ao, Yo, Do, maxBc=8):
        lib = RefCase.lib()
        lib.svref_ga_create.restype = C.c_void_p
        lib.svref_ga_last_error.restype = C.c_char_p
        self.lib = lib
        Ao, Yo, Do = _f64(Ao), _f64(Yo), _f64(Do)
        self.tDof, self.nNo = Ao.shape
        self.eqs = list(eqs)
        arr = (abi.EqTime * len(eqs))(*eqs)
        self.h = lib.svref_ga_create(C.c_int(self.tDof), C.c_int(self.nNo), C.c_int(len(eqs)), arr, C.c_double(dt), C.c_int(int(dFlag)),
                                     C.c_int(int(sstEq)), _d(Ao), _d(Yo), _d(Do), C.c_int(maxBc))
        if not self.h:
            raise RuntimeError("svref_ga_create: " + lib.svref_ga_last_error().decode())

    def _call(self, name, *args):
        rc = getattr(self.lib, "svref_ga_" + name)(C.c_void_p(self.h), *args)
        if rc != 0:
            raise RuntimeError(f"svref_ga_{name}: {self.lib.svref_ga_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.svref_ga_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, which, A=None, Y=None, D=None):
        self._call("set", C.c_int(which), _d(_f64(A)), _d(_f64(Y)), _d(_f64(D)))

    def get(self, which):
        out = [np.zeros((self.tDof, self.nNo), order="F") for _ in range(3)]
        self._call("get", C.c_int(which), _d(out[0]), _d(out[1]), _d(out[2]))
        return out

    def set_ad(self, Ad):
        self._call("set_ad", _d(_f64(Ad)))

    def get_ad(self):
        Ad = np.zeros((3, self.nNo), order="F")
        self._call("get_ad", _d(Ad))
        return Ad

    def set_solid_nodes(self, iEq, solid_phys, flags):
        self._call("set_solid_nodes", C.c_int(iEq), C.c_int(solid_phys), _i(_i32(flags)))

    def predictor(self):
        self._call("predictor")

    def initiator(self, cEq=0):
        self._call("initiator", C.c_int(cEq))

    def corrector(self, cEq, R, Rd=None):
        self._call("corrector", C.c_int(cEq), _d(_f64(R)), _d(_f64(Rd)))

    def add_dir_bc(self, iEq, nodes, eDrn=(0, 0, 0), impD=False, g=1.0, gx=None, nV=None):
        nodes = _i32(nodes)
        e = _i32(np.asarray(eDrn, dtype=np.int32))
        self._call("add_dir_bc", C.c_int(iEq), C.c_int(len(nodes)), _i(nodes), _i(e), C.c_int(int(impD)), C.c_double(g),
                   _d(_f64(gx)), _d(_f64(nV)))

    def set_bc_dir(self):
        self._call("set_bc_dir")
