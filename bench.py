#!/usr/bin/env python
"""bench.py — element assemblies/s (FP64 tet4 fluid) and Newton-step time on N B200s.

Contract:
  python bench.py --gpus N --steps K --warmup W            our arm (libsvb200.so through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's own CPU code on the host cores

One "step" is one Newton iteration of the hot path on the synthetic 10 M-tet4 cylinder (config C2 of
SURVEY.md §8d), one 10 M-element partition per GPU (weak scaling; the global mesh is the N-times longer cylinder):
    predictor/initiator -> ls_alloc (zero R, Val) -> element assembly + scatter -> shared-node sum of R ->
    fsils_solve (GMRES) -> corrector, all on the device (no nodal array crosses PCIe inside the step).
`value` is elements assembled per second over the ASSEMBLY stage of the timed steps (zero + kernel +
halo, device-resident inputs, CUDA events on the library's stream, max over ranks); `ms_per_step` is
the whole Newton step; `e2e` is the assembly stage driven with HOST buffers through the C ABI (H2D
of Ag/Yg from pinned memory and D2H of the residual inside the timed region).

After the timed region every rank runs the PARITY GATE (never timed): its assembled R / Val against an oracle
assembly (oracle/_ref/libsvref.so = the compiled reference, else the C restatement) of sub-blocks of the same mesh and
state — one inside the partition and one straddling the partition interface where the most ranks meet (rows completed
by the shared-node sum against the single-partition oracle) — and the TRUE preconditioned residual of the GMRES
answer, ||W (R - K x)|| <= relTol ||W R||, reduced over the ranks.  The line carries "parity": {...}; a failed gate
makes the process exit non-zero.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout of this script is exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints
# "NCCL version ..." there at every NCCL_DEBUG level from VERSION up, WARN included), so descriptor 1 is pointed at stderr
# for the whole run and the JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def claim_stdout():
    """Called by main() only (importing this module, as tests/test_gpu_fullsize.py does, must not touch the descriptors)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off (sysfs: numa_node / local_cpulist of the PCI device), so that
    the page-locked host buffers of the e2e stage are first-touched next to the GPU's root complex.  What `mpirun --bind-to` /
    `numactl` does for the reference's MPI ranks; a no-op on single-node virtual machines (numa_node = -1)."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return {"bound": False, "node": node, "pci": bdf, "cpus_allowed": len(allowed)}
        os.sched_setaffinity(0, use)
        return {"bound": True, "node": node, "pci": bdf, "cpus": len(use), "cpus_allowed": len(allowed)}
    except Exception as ex:        # no sysfs entry, old torch: stay unbound
        return {"bound": False, "why": str(ex)[:80]}


def emit_line(obj):
    if _REAL_STDOUT is None:
        print(json.dumps(obj), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


from svmultiphysics_b200 import abi, elements, meshgen, partition  # noqa: E402

FLOP_PER_ELEMENT = 11.6e3     # algorithmic FP64 flop per tet4 VMS element (Newtonian)
SPMV_BYTES = lambda nnz, nNo: nnz * 132 + nNo * 72   # noqa: E731  SURVEY.md §8(d), dof = 4
LSEG = 30.0                   # cylinder length per GPU share


def lattice_state(m, rank=0, nz=0, U=10.0, R=2.0, tDof=4, noise=0.01):
    """Poiseuille flow + 1 % deterministic pseudo-noise keyed on the GLOBAL lattice index, so that the nodes two
    partitions share carry identical state on both ranks and any sub-block of the mesh can be regenerated on its own
    (seeds replaced by a hash).  Meshes of meshgen.cylinder_box carry their global indices (m.gijk); for
    the slabs of meshgen.cylinder_slab the index is rebuilt from (rank, nz)."""
    if m.gijk is not None:
        i, j, k = m.gijk
    else:
        n1 = m.lattice[0] + 1
        ids = np.arange(m.nNo, dtype=np.int64)
        i, j, k = ids % n1, (ids // n1) % n1, ids // (n1 * n1) + rank * nz
    def h(c):
        v = (i * 73856093) ^ (j * 19349663) ^ (k * 83492791) ^ (c * 2654435761)
        v = (v ^ (v >> 13)) * 1274126177 & 0xFFFFFFFF
        return (v / 2147483648.0) - 1.0
    r2 = (m.x[0] ** 2 + m.x[1] ** 2) / (R * R)
    Yg = np.zeros((tDof, m.nNo), order="F")
    Yg[2] = U * (1.0 - r2)
    for c in range(3):
        Yg[c] += noise * U * h(c)
    Yg[3] = -1.0 * m.x[2] + noise * h(3)
    Ag = np.zeros((tDof, m.nNo), order="F")
    for c in range(4):
        Ag[c] = 1e-2 * h(4 + c)
    return Ag, Yg


def ls_config(args):
    return abi.ls_params(abi.LS_GMRES, mItr=args.ls_mitr, sD=args.ls_sd, relTol=args.ls_reltol)


def workload_config(args):
    """The workload both arms are run on (identical in the two JSON lines); everything arm-specific (partition,
    transport, scatter mode, cores) lives under "run" / "cpu_baseline"."""
    nel = 6 * args.n * args.n * args.nz
    return {"workload": f"C2 synthetic cylinder, 6*{args.n}^2*{args.nz} = {nel} tet4 per GPU, Newtonian VMS fluid",
            "elements_per_gpu": nel, "dt": 1e-3,
            "linear_solver": f"GMRES sD={args.ls_sd} mItr={args.ls_mitr} relTol={args.ls_reltol} + FSILS diagonal preconditioner",
            "newton_step": "ls_alloc -> construct_fluid (+ do_assem) -> commu(R) -> fsils_solve",
            "l2": "inputs larger than L2 (Val = 3.2 GB/GPU is rewritten every step)"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.samples:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self, t0=None, t1=None):
        if self.proc:
            self.proc.terminate()
        return self.window(t0, t1) if t0 is not None else None


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref/libsvref.so) as an MPI run on the host cores
# ------------------------------------------------------------------------------------------------
def best_blocks(ncells, P):
    """(bx, by, bz) with bx*by*bz = P minimising the largest block, then the interface area — what a graph partitioner
    balances (ParMETIS_V3_PartMeshKway, Code/Source/solver/SPLIT.c:76-89)."""
    best = None
    for bx in range(1, P + 1):
        if P % bx:
            continue
        for by in range(1, P // bx + 1):
            if (P // bx) % by:
                continue
            bz = P // (bx * by)
            if bx > ncells[0] or by > ncells[1] or bz > ncells[2]:
                continue
            big = [-(-ncells[d] // b) for d, b in enumerate((bx, by, bz))]
            vol = big[0] * big[1] * big[2]
            cut = (bx - 1) * ncells[1] * ncells[2] + (by - 1) * ncells[0] * ncells[2] + (bz - 1) * ncells[0] * ncells[1]
            key = (vol, cut)
            if best is None or key < best[0]:
                best = (key, (bx, by, bz))
    return best[1]


def _ref_rank_worker(q, n, nzg, L, blocks, r, P, shm, steps, warmup, solve_steps, ls_tuple):
    """Rank r of a P-rank run of the compiled reference over the shared-memory MPI shim (oracle/ref_build/mpi_stub.cpp):
    the reference's own partition set-up (fsils_lhs_create), then per step ls_alloc -> construct_fluid -> commu(R) ->
    fsils_solve (Code/Source/solver/Integrator.cpp:104-160)."""
    try:
        os.environ.update(SVREF_MPI_SIZE=str(P), SVREF_MPI_RANK=str(r), SVREF_MPI_SHM=shm)
        from oracle import refbind
        m = meshgen.cylinder_block(n, nzg, blocks, r, L=L)
        Ag, Yg = lattice_state(m)
        gid = (m.gijk[0] + (n + 1) * (m.gijk[1] + (n + 1) * m.gijk[2])).astype(np.int32)
        c = refbind.RefCase()
        c.set_coords(m.x)
        if P > 1:
            c.set_partition((n + 1) * (n + 1) * (nzg + 1), gid)
        c.add_mesh(m.IEN)
        c.build_graph(1)
        wall = m.faces["wall"]
        c.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))
        eq, dmn = abi.fluid_eq(1e-3), [abi.fluid_domain()]
        ls = abi.ls_params(abi.LS_GMRES, mItr=ls_tuple[0], sD=ls_tuple[1], relTol=ls_tuple[2])
        c.set_state(Ag, Yg)
        t_asm, t_solve, info = [], [], None
        for s in range(warmup + steps):
            c.barrier()
            t0 = time.perf_counter()
            c.alloc(4)
            c.assemble(0, eq, dmn)
            c.commu_R()
            t1 = time.perf_counter()
            timed = s >= warmup
            if timed:
                t_asm.append(t1 - t0)
            if timed and len(t_solve) < solve_steps:
                c.barrier()
                t2 = time.perf_counter()
                _, out, _ = c.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1))
                t_solve.append(time.perf_counter() - t2)
                info = (out.RI.itr, int(out.RI.success), out.RI.iNorm, out.RI.fNorm)
        q.put(("ok", r, m.nEl, t_asm, t_solve, info))
    except Exception as ex:      # a dead rank would leave the others in a barrier: report, the parent tears the run down
        q.put(("error", r, repr(ex)))


def reference_newton(args, steps, warmup, solve_steps, procs=None, timeout_s=1500):
    """The reference arm / CPU baseline: the C2 mesh (one GPU's share of the workload: 6*n^2*nz tet4) on all host cores as
    an MPI-style run of the unmodified reference, element-partitioned into balanced blocks.  Returns a dict."""
    from oracle import refbind
    if not refbind.have_ref():
        return reference_port_sample(args)
    procs = int(procs or os.environ.get("SVB200_REF_PROCS", 0) or os.cpu_count() or 1)
    procs = max(1, min(procs, 64))
    blocks = best_blocks((args.n, args.n, args.nz), procs)
    shm = "/svref_bench_" + uuid.uuid4().hex[:10]
    ctx = mp.get_context("spawn")     # fresh interpreters: the MPI shim reads its environment once per process
    q = ctx.Queue()
    ls_tuple = (args.ls_mitr, args.ls_sd, args.ls_reltol)
    ps = [ctx.Process(target=_ref_rank_worker, args=(q, args.n, args.nz, LSEG, blocks, r, procs, shm, steps, warmup,
                                                     solve_steps, ls_tuple)) for r in range(procs)]
    for p in ps:
        p.start()
    res, err = [], None
    deadline = time.time() + timeout_s
    try:
        while len(res) < procs and err is None:
            try:
                item = q.get(timeout=max(1.0, min(5.0, deadline - time.time())))
            except Exception:
                if time.time() > deadline:
                    err = "timeout"
                elif any(p.exitcode not in (None, 0) for p in ps):
                    err = "a reference rank died"
                continue
            if item[0] == "error":
                err = f"rank {item[1]}: {item[2]}"
            else:
                res.append(item)
    finally:
        for p in ps:
            if err is not None and p.is_alive():
                p.terminate()
            p.join(timeout=30)
        try:
            os.unlink("/dev/shm" + shm)
        except OSError:
            pass
    if err is not None:
        raise RuntimeError("reference run failed: " + err)
    nEl = sum(r[2] for r in res)
    nst = len(res[0][3])
    asm = [max(r[3][s] for r in res) for s in range(nst)]                 # max over ranks per step
    nso = len(res[0][4])
    sol = [max(r[4][s] for r in res) for s in range(nso)]
    t_asm, t_sol = float(np.mean(asm)), (float(np.mean(sol)) if sol else None)
    info = res[0][5]
    out = {"value": nEl / t_asm, "unit": "element assemblies/s", "cores": procs, "kind": "reference", "elements": nEl,
           "sample": f"the C2 mesh itself (6*{args.n}^2*{args.nz} = {nEl} tet4, one GPU's share of the workload), element-partitioned "
                     f"into {blocks[0]}x{blocks[1]}x{blocks[2]} blocks over {procs} MPI ranks of the compiled reference (shared-memory MPI shim); "
                     f"ls_alloc + construct_fluid + commu(R) in each of {nst} timed steps, fsils_solve in the first {nso}; max over ranks",
           "assemble_ms": t_asm * 1e3, "solve_ms": None if t_sol is None else t_sol * 1e3,
           "newton_step_ms": None if t_sol is None else (t_asm + t_sol) * 1e3, "partition_blocks": list(blocks)}
    if info:
        out["gmres"] = {"itr": info[0], "success": info[1], "iNorm": info[2], "fNorm": info[3]}
    return out


def reference_port_sample(args, n=24, nz=12):
    """No compiled reference on this box: the C restatement (oracle/libsvoracle.so) on one core, bounded sample."""
    from oracle import refbind
    m = meshgen.cylinder_box(n, nz, [(0, n), (0, n), (0, nz)])
    Ag, Yg = lattice_state(m)
    c = refbind.OracleCase()
    c.set_coords(m.x); c.add_mesh(m.IEN); c.build_graph(1)
    wall = m.faces["wall"]
    c.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))
    eq, dmn = abi.fluid_eq(1e-3), [abi.fluid_domain()]
    c.set_state(Ag, Yg)
    ts = []
    for s in range(3):
        t0 = time.perf_counter(); c.alloc(4); c.assemble(0, eq, dmn); ts.append(time.perf_counter() - t0)
    t2 = time.perf_counter()
    _, o, _ = c.solve(4, abi.LS_GMRES, ls_config(args), np.ones(1, np.int32), np.zeros(1))
    ts_sol = time.perf_counter() - t2
    t = float(np.mean(ts[1:]))
    return {"value": m.nEl / t, "unit": "element assemblies/s", "cores": 1, "kind": "port", "elements": m.nEl,
            "sample": f"6*{n}^2*{nz} = {m.nEl} tet4 cylinder, C restatement on one core (oracle/_ref/libsvref.so absent)",
            "assemble_ms": t * 1e3, "solve_ms": ts_sol * 1e3, "newton_step_ms": (t + ts_sol) * 1e3,
            "gmres": {"itr": o.RI.itr, "success": int(o.RI.success), "iNorm": o.RI.iNorm, "fNorm": o.RI.fNorm}}


# ------------------------------------------------------------------------------------------------
# parity gate
# ------------------------------------------------------------------------------------------------
def _oracle_cls():
    from oracle import refbind
    return (refbind.RefCase, "reference") if refbind.have_ref() else (refbind.OracleCase, "port")


def _clip_box(centre, half, lo3, hi3):
    return [(max(lo3[d], centre[d] - half[d]), min(hi3[d], centre[d] + half[d])) for d in range(3)]


def parity_boxes(lb, rank, half=(4, 4, 4)):
    """Two sub-boxes of GLOBAL cells per rank: one centred inside the rank's block, one centred on the corner of the block
    where the most partitions meet (straddling the interfaces; at the global boundary when there is no neighbour)."""
    nc = lb.nc
    rng = lb.cell_ranges(rank)
    centre = [(lo + hi) // 2 for lo, hi in rng]
    inner = _clip_box(centre, half, [lo for lo, _ in rng], [hi for _, hi in rng])
    corner = []
    for d in range(3):
        lo, hi = rng[d]
        if hi < nc[d]:
            corner.append(hi)              # interface with a higher block
        elif lo > 0:
            corner.append(lo)              # interface with a lower block
        else:
            corner.append(0 if d < 2 else nc[d])   # no neighbour: the lateral wall / the outlet plane
    iface = _clip_box(corner, half, [0, 0, 0], list(nc))
    return {"interior": inner, "interface": iface}


def parity_assembly(eng, m, lb, rank, rowPtr, colPtr, eq, dmn, n, nzg, L):
    """max relative error of this rank's R rows (after the shared-node sum) and Val rows against oracle assemblies of
    sub-boxes of the same mesh and state.  R is compared with the oracle on ALL cells of the box (the single-partition
    answer); Val, which is never communicated, with the oracle on the cells of the box this rank owns."""
    cls, kind = _oracle_cls()
    rng = lb.cell_ranges(rank)
    gid_mine = m.gijk[0] + (n + 1) * (m.gijk[1] + (n + 1) * m.gijk[2])
    out = {"oracle": kind, "boxes": {}}
    worst = 0.0
    for name, F in parity_boxes(lb, rank).items():
        M = [(max(F[d][0], rng[d][0]), min(F[d][1], rng[d][1])) for d in range(3)]     # cells of the box this rank owns
        if any(lo >= hi for lo, hi in M):
            continue
        res = {}
        for which, box in (("full", F), ("mine", M)):
            if which == "mine" and box == F:
                res["mine"] = res["full"]
                continue
            b = meshgen.cylinder_box(n, nzg, box, L=L)
            Ab, Yb = lattice_state(b)
            c = cls()
            c.set_coords(b.x); c.add_mesh(b.IEN)
            rp, cp = c.build_graph(0)
            c.alloc(4); c.set_state(Ab, Yb); c.assemble(0, eq, dmn)
            res[which] = (b, rp, cp, c.get_R(), c.get_Val())
            c.close()
        # rows that see all of their elements inside F: strictly inside the box or on the global lattice boundary
        bM = res["mine"][0]
        gi, gj, gk = bM.gijk
        nc = lb.nc
        comp = np.ones(bM.nNo, bool)
        for d, g in enumerate((gi, gj, gk)):
            comp &= ((g > F[d][0]) | (g == 0)) & ((g < F[d][1]) | (g == nc[d]))
        rowsM = np.flatnonzero(comp)
        gidM = (gi + (n + 1) * (gj + (n + 1) * gk))[rowsM]
        loc = lb.local_of(rank, gidM).astype(np.int32)
        assert np.array_equal(gid_mine[loc], gidM)
        # R: the full-box oracle (all ranks' elements) against this rank's summed rows
        bF, _, _, RF, _ = res["full"]
        gidF = bF.gijk[0] + (n + 1) * (bF.gijk[1] + (n + 1) * bF.gijk[2])
        posF = np.searchsorted(gidF, gidM)
        assert np.array_equal(gidF[posF], gidM)
        R_loc = eng.get_rows(abi.ARRAY_R, loc)
        scaleR = max(np.abs(RF).max(), 1e-300)
        eR = float(np.abs(R_loc - RF[:, posF]).max() / scaleR)
        # Val: this rank's own elements only
        _, rpM, cpM, _, VM = res["mine"]
        V_loc = eng.get_rows(abi.ARRAY_VAL, loc, rowPtr)
        cols_loc = np.concatenate([colPtr[rowPtr[a]:rowPtr[a + 1]] for a in loc])
        cols_or = np.concatenate([cpM[rpM[a]:rpM[a + 1]] for a in rowsM])
        gidMall = gi + (n + 1) * (gj + (n + 1) * gk)
        same_graph = len(cols_loc) == len(cols_or) and bool(np.array_equal(gid_mine[cols_loc], gidMall[cols_or]))
        eV = float("inf")
        if same_graph:
            V_or = np.concatenate([VM[:, rpM[a]:rpM[a + 1]] for a in rowsM], axis=1)
            eV = float(np.abs(V_loc - V_or).max() / max(np.abs(VM).max(), 1e-300))
        nshared = int((lb.multiplicity(rank)[loc] > 1).sum())
        out["boxes"][name] = {"cells": [list(x) for x in F], "rows": int(len(rowsM)), "rows_shared": nshared,
                              "max_ranks_on_a_row": int(lb.multiplicity(rank)[loc].max()), "R_max_rel": eR, "Val_max_rel": eV,
                              "same_graph": same_graph}
        worst = max(worst, eR, eV)
    out["max_rel"] = worst
    return out


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configs (C1, C4, C5) at N = 1: driver-visible numbers next to the headline
# ------------------------------------------------------------------------------------------------
def _timed(eng, fn, reps):
    ms = []
    for _ in range(reps):
        eng.timer_mark(0); fn(); eng.timer_mark(1)
        ms.append(eng.timer_elapsed())
    return float(np.mean(ms))


def config_c1_ns(eng, m, Ag, Yg, eq, dmn, hbm_peak, nnz):
    """C1: the NS (Schur-complement) solver — the default of every fluid case, with the settings and the face layout of
    tests/cases/fluid/pipe_RCR_3d/solver.xml:75-87 (Dirichlet wall + inlet, coupled resistance outlet) — on the C2 mesh."""
    faces = [(abi.BC_DIR, m.faces[k], np.zeros((3, len(m.faces[k])), order="F")) for k in ("wall", "inlet")]
    out = m.faces["outlet"]
    val = np.zeros((3, len(out)), order="F"); val[2] = 4.0 * np.pi / len(out)
    faces.append((abi.BC_NEU, out, val))
    eng.set_num_faces(len(faces))
    for i, (g, nodes, v) in enumerate(faces):
        eng.set_face(i, g, nodes, v)
    ls = abi.ls_params(abi.LS_NS, mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17))
    incL, res = np.ones(3, np.int32), np.array([0.0, 0.0, 0.8])
    best = None
    for rep in range(3):
        eng.set_state(Ag, Yg); eng.alloc(4); eng.assemble(0, eq, dmn)
        eng.timer_mark(0)
        _, o, _ = eng.solve(4, abi.LS_NS, ls, incL, res, want_solution=False)
        eng.timer_mark(1)
        ms = eng.timer_elapsed()
        if rep > 0 and (best is None or ms < best[0]):
            best = (ms, o)
    ms, o = best
    nNo = m.nNo
    cg_bytes = nnz * (28 + 28 + 12) + nNo * 8 * 14         # G p, D (G p), L p (values + column index) + the CG vectors
    cg_ms = o.CG.callD * 1e3 / max(o.CG.itr, 1)
    return {"workload": f"pipe_RCR_3d solver settings (NS: RI 15/1e-3, GM 10/250/1e-3, CG 300/1e-3; resistance outlet) on the C2 mesh, {m.nEl} tet4",
            "ns_solve_ms": ms, "outer_itr": o.RI.itr, "success": int(o.RI.success), "gmres_itr": o.GM.itr, "cg_itr": o.CG.itr,
            "gmres_ms_per_itr": o.GM.callD * 1e3 / max(o.GM.itr, 1), "cg_ms_per_itr": cg_ms,
            "Resm": o.Resm, "Resc": o.Resc, "fNorm_over_iNorm": o.RI.fNorm / o.RI.iNorm,
            "roofline": {"bound": "hbm", "kernel": "Schur-complement CG iteration (bsr_spmv_lg_kernel<3,1,..> G p, schur_sp4_kernel with the <p,Sp> partials, fused X/R/P updates)",
                         "achieved": cg_bytes / (cg_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": cg_bytes / (cg_ms * 1e-3) * 1e-9 / hbm_peak, "traffic": None,
                         "algorithmic": "nnz*(28+28+12) + 14 nodal scalars per CG iteration (SURVEY 8d NS sub-blocks)"}}


def config_c4_struct(device, fp64_peak, hbm_peak, n=171):
    """C4: hex8 neo-Hookean (ST91) block compression, n^3 = 5.0 M elements (tests/cases/struct/block_compression): struct_3d
    assembly, the dof-3 SpMV and BiCGStab."""
    from svmultiphysics_b200.engine import Engine
    m = meshgen.box_hex8(n, n, n, (1e-3, 1e-3, 1e-3))
    rng = np.random.default_rng(1236)
    L = 1e-3
    Dg = np.zeros((3, m.nNo), order="F")
    Dg[:3] = 0.01 * m.x * np.array([[1.0], [-0.5], [0.3]]) + 1e-4 * L * rng.standard_normal((3, m.nNo))
    Yg = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo)))
    Ag = np.asfortranarray(rng.standard_normal((3, m.nNo)))
    e = Engine(device)
    try:
        rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
        w, N, Nx = elements.tables(8)
        e.set_mesh(0, m.IEN, w, N, Nx); e.set_coords(m.x)
        e.alloc(3); e.set_state(Ag, Yg, Dg)
        eq, dm = abi.struct_eq(1e-4), [abi.struct_domain()]
        e.assemble(0, eq, dm)
        stage = _timed(e, lambda: (e.alloc(3), e.assemble(0, eq, dm)), 3)
        kern = e.last_timing()[0]
        spmv = e.bench_spmv(3, 10)
        faces = []
        for k, name in enumerate(("X0", "Y0", "Z0")):
            val = np.ones((3, len(m.faces[name])), order="F"); val[k] = 0.0
            faces.append((abi.BC_DIR, m.faces[name], val))
        e.set_num_faces(len(faces))
        for i, (g, nodes, val) in enumerate(faces):
            e.set_face(i, g, nodes, val)
        ls = abi.ls_params(abi.LS_BICGS, mItr=50, relTol=1e-12)
        for rep in range(2):      # the first solve allocates the Krylov workspace (host-side cudaMalloc inside the timed window)
            e.alloc(3); e.assemble(0, eq, dm)
            e.timer_mark(0)
            _, o, _ = e.solve(3, abi.LS_BICGS, ls, np.ones(3, np.int32), np.zeros(3), want_solution=False)
            e.timer_mark(1)
            sol = e.timer_elapsed()
    finally:
        e.close()
    tf = m.nEl * 130e3 / (kern * 1e-3) * 1e-12
    sp = (len(cp) * 76 + m.nNo * 56) / (spmv * 1e-3) * 1e-9
    return {"workload": f"hex8 neo-Hookean + ST91 block, {n}^3 = {m.nEl} elements, dt 1e-4 (block_compression analogue)",
            "value": m.nEl / (stage * 1e-3), "unit": "element assemblies/s", "assembly_stage_ms": stage, "assembly_kernel_ms": kern,
            "bicgstab_itr": o.RI.itr, "bicgstab_ms_per_itr": sol / max(o.RI.itr, 1),
            "roofline": {"bound": "fp64", "kernel": "assemble_struct_kernel<8>", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": tf / fp64_peak, "traffic": None, "algorithmic": f"130000 flop/hex8 (SURVEY 8d) x {m.nEl} elements"},
            "roofline_spmv": {"bound": "hbm", "kernel": "bsr_spmv_lg_kernel<3,3,1,2,4,0> (two blocks per 6-lane row group and step, 4 steps in flight)", "achieved": sp, "peak": hbm_peak, "unit": "GB/s",
                              "frac": sp / hbm_peak, "ms": spmv, "traffic": None, "algorithmic": "nnz*76 + nNo*56 bytes per launch"}}


def config_c5_fsi(device, hbm_peak, n=90, nz=120):
    """C5: FSI pipe (fluid lumen + solid wall sharing interface nodes, tDof = 7, tests/cases/fsi/pipe_3d): construct_fsi with the
    lumen and the wall as two meshes, construct_mesh, GMRES on the coupled system and CG on the mesh equation."""
    from svmultiphysics_b200.engine import Engine
    m = meshgen.cylinder_tet4(n, nz, R=1.0, L=3.0)
    c = m.x[:, m.IEN].mean(axis=1)
    solid = (c[0] ** 2 + c[1] ** 2) > 0.55 ** 2
    eId = np.where(solid, 2, 1).astype(np.int32)
    rng = np.random.default_rng(21)
    Ag, Yg, _ = meshgen.poiseuille_state(m, R=1.0, U=5.0, tDof=7)
    Yg[4:7] = 0.2 * rng.standard_normal((3, m.nNo))
    sc = 0.05 * (6.0 / n)                      # keep the random displacements well inside the elements
    Dg = np.zeros((7, m.nNo), order="F")
    Dg[:3] = sc * 2e-3 * rng.standard_normal((3, m.nNo))
    Dg[4:7] = sc * 5e-3 * rng.standard_normal((3, m.nNo))
    Bf = np.asfortranarray(0.1 * rng.standard_normal((3, m.nNo)))
    fl, so = np.where(eId == 1)[0], np.where(eId == 2)[0]
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=abi.SCATTER_ATOMIC, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0), abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
    e = Engine(device)
    try:
        rp, cp = e.lhsa(m.nNo, [m.IEN]); e.set_graph(rp, cp)
        w, N, Nx = elements.tables(4)
        e.set_mesh(0, np.asfortranarray(m.IEN[:, fl]), w, N, Nx, eId=eId[fl])
        e.set_mesh(1, np.asfortranarray(m.IEN[:, so]), w, N, Nx, eId=eId[so])
        e.set_mesh(2, m.IEN, w, N, Nx)              # the mesh-motion equation runs over every element
        e.set_coords(m.x)
        e.alloc(4); e.set_state(Ag, Yg, Dg, Bf)
        e.assemble(0, eq, dmn); e.assemble(1, eq, dmn)
        fsi_ms = _timed(e, lambda: (e.alloc(4), e.assemble(0, eq, dmn), e.assemble(1, eq, dmn)), 3)
        wall = m.faces["wall"]
        e.set_num_faces(1); e.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))
        ls = abi.ls_params(abi.LS_GMRES, mItr=2, sD=50, relTol=1e-8)
        for rep in range(2):      # first solve = workspace allocation
            e.alloc(4); e.assemble(0, eq, dmn); e.assemble(1, eq, dmn)
            e.timer_mark(0)
            _, o, _ = e.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1), want_solution=False)
            e.timer_mark(1)
            gm_ms = e.timer_elapsed() / max(o.RI.itr, 1)
        eqm, dmm = abi.mesh_eq(1e-3), [abi.mesh_domain(E=1.0, nu=0.3)]
        e.alloc(3); e.set_old_disp(np.asfortranarray(0.9 * Dg)); e.assemble(2, eqm, dmm)
        msh_ms = _timed(e, lambda: (e.alloc(3), e.assemble(2, eqm, dmm)), 3)
        lsc = abi.ls_params(abi.LS_CG, mItr=100, relTol=1e-10)
        for rep in range(2):
            e.alloc(3); e.assemble(2, eqm, dmm)
            e.timer_mark(0)
            _, oc, _ = e.solve(3, abi.LS_CG, lsc, np.ones(1, np.int32), np.zeros(1), want_solution=False)
            e.timer_mark(1)
            cg_ms = e.timer_elapsed() / max(oc.RI.itr, 1)
    finally:
        e.close()
    nnz = len(cp)
    bytes_fsi = nnz * 128 + m.nNo * 32 + m.nEl * 464           # Val written once + R + the nodal gather
    return {"workload": f"FSI pipe, {m.nEl} tet4 = {len(fl)} fluid (ALE VMS) + {len(so)} solid (nHK M94), tDof 7, dt 1e-3 (pipe_3d analogue)",
            "value": m.nEl / (fsi_ms * 1e-3), "unit": "element assemblies/s (construct_fsi, zero + both meshes)",
            "construct_fsi_stage_ms": fsi_ms, "construct_mesh_stage_ms": msh_ms, "mesh_value": m.nEl / (msh_ms * 1e-3),
            "fsi_gmres_ms_per_itr": gm_ms, "fsi_gmres_itr": o.RI.itr, "mesh_cg_ms_per_itr": cg_ms, "mesh_cg_itr": oc.RI.itr,
            "roofline": {"bound": "hbm", "kernel": "construct_fsi (fluid TET4 grouped kernel + struct TET4 kernel)",
                         "achieved": bytes_fsi / (fsi_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_fsi / (fsi_ms * 1e-3) * 1e-9 / hbm_peak, "traffic": None,
                         "algorithmic": "nnz*128 (Val written once) + nNo*32 (R) + nEl*464 (nodal gather) bytes per assembly",
                         "note": "the stage is compute bound, not HBM bound: see fp64_model"},
            "fp64_model": {"flop": len(fl) * FLOP_PER_ELEMENT + len(so) * 24000.0,
                           "TFLOP/s": (len(fl) * FLOP_PER_ELEMENT + len(so) * 24000.0) / (fsi_ms * 1e-3) * 1e-12,
                           "derivation": "fluid tets x 11.6 kflop (SURVEY 8d) + solid tets x 24 kflop = 4 Gauss points x (compute_pk2cc ~2 k + "
                                         "4 nodes x 0.3 k Bm/DBm + 16 node pairs x 0.18 k), the SURVEY's 16 kflop per HEX8 Gauss point rescaled "
                                         "from 8 nodes / 64 pairs to 4 / 16; the closed-form TET4 kernels execute about a quarter of it"}}


def _lattice_hash(m, c):
    """Deterministic pseudo-noise in [-1, 1) keyed on the GLOBAL lattice index of the nodes of a lattice block (m.gijk)."""
    i, j, k = m.gijk
    v = (i * 73856093) ^ (j * 19349663) ^ (k * 83492791) ^ (c * 2654435761)
    v = (v ^ (v >> 13)) * 1274126177 & 0xFFFFFFFF
    return (v / 2147483648.0) - 1.0


def lattice_fsi_case(b, n, R=2.0):
    """FSI set-up of a lattice block of the cylinder (any sub-box regenerates the same numbers): solid wall = elements whose
    centroid lies outside 0.55 R (domain bit 1), lumen = fluid (bit 0); tDof = 7 state with mesh velocity / displacement."""
    c = b.x[:, b.IEN].mean(axis=1)
    eId = np.where(c[0] ** 2 + c[1] ** 2 > (0.55 * R) ** 2, 2, 1).astype(np.int32)
    Ag, Yg = lattice_state(b, tDof=7)
    h = 2.0 * R / n
    Dg = np.zeros((7, b.nNo), order="F")
    for k in range(3):
        Yg[4 + k] = 0.2 * _lattice_hash(b, 20 + k)
        Dg[k] = 1e-3 * h * _lattice_hash(b, 30 + k)
        Dg[4 + k] = 2e-3 * h * _lattice_hash(b, 40 + k)
        Ag[4 + k] = 1e-2 * _lattice_hash(b, 50 + k)
    return eId, Ag, Yg, Dg


def config_c5_fsi_multi(eng, m, lb, rank, world, rowPtr, colPtr, n, nzg, L, reduce_ranks, with_parity):
    """C5 on N GPUs (BASELINE.json: "FSI pipe ... coupled assembly at 2/4/8 GPUs"): construct_fsi (ALE fluid lumen + nHK wall) and the
    mesh-motion equation on this rank's block of the N-times-longer cylinder, shared-node sums included, on the engine and the
    partition of the headline run; parity of the coupled residual on the block corner where the most ranks meet against a
    single-partition oracle assembly of the same sub-box."""
    eId, Ag, Yg, Dg = lattice_fsi_case(m, n)
    fl, so = np.flatnonzero(eId == 1), np.flatnonzero(eId == 2)
    af, am, gam, beta = abi.gen_alpha(0.5)
    eq = abi.EqParams(dt=1e-3, af=af, am=am, gam=gam, beta=beta, phys=abi.PHYS_FSI, dof=4, tDof=7, s=0, mvMsh=1, vmsStab=1,
                      scatter=abi.SCATTER_ATOMIC, reserved=0)
    dmn = [abi.fluid_domain(rho=1.0, mu=0.04, Id=0), abi.struct_domain(rho=1.0, volType=abi.VOL_M94, E=1e7, nu=0.3, Kpen=1e7 / (3 * (1 - 0.6)), Id=1)]
    w, N, Nx = elements.tables(4)
    eng.set_mesh(1, np.asfortranarray(m.IEN[:, fl]), w, N, Nx, eId=eId[fl])
    eng.set_mesh(2, np.asfortranarray(m.IEN[:, so]), w, N, Nx, eId=eId[so])
    eng.set_mesh(3, m.IEN, w, N, Nx)
    eng.alloc(4); eng.set_state(Ag, Yg, Dg)

    def fsi():
        eng.alloc(4); eng.assemble(1, eq, dmn); eng.assemble(2, eq, dmn); eng.commu_R()
    fsi()
    fsi_ms = reduce_ranks(_timed(eng, fsi, 3), "max")
    out = {}
    if with_parity:
        cls, kind = _oracle_cls()
        err, nrows, nwall, mult = 0.0, 0, 0, 0
        gid_mine = m.gijk[0] + (n + 1) * (m.gijk[1] + (n + 1) * m.gijk[2])
        for F in parity_boxes(lb, rank).values():      # the box inside the block and the one on its most-shared corner
            b = meshgen.cylinder_box(n, nzg, F, L=L)
            eIdb, Ab, Yb, Db = lattice_fsi_case(b, n)
            c = cls(); c.set_coords(b.x); c.add_mesh(b.IEN, eId=eIdb); c.build_graph(0)
            c.alloc(4); c.set_state(Ab, Yb, Db); c.assemble(0, eq, dmn)
            RF = c.get_R(); c.close()
            gi, gj, gk = b.gijk
            comp = np.ones(b.nNo, bool)
            for d, g in enumerate((gi, gj, gk)):
                comp &= ((g > F[d][0]) | (g == 0)) & ((g < F[d][1]) | (g == lb.nc[d]))
            gid = gi + (n + 1) * (gj + (n + 1) * gk)
            rows = np.flatnonzero(comp & np.isin(gid, gid_mine))
            if len(rows) == 0:
                continue
            loc = lb.local_of(rank, gid[rows]).astype(np.int32)
            R_loc = eng.get_rows(abi.ARRAY_R, loc)
            # momentum rows of the wall are ~10^7 x the fluid rows: compare every row against its own scale class
            sol_nodes = np.isin(np.arange(b.nNo), np.unique(b.IEN[:, eIdb == 2]))[rows]
            for sel in (sol_nodes, ~sol_nodes):
                if sel.any():
                    ref = RF[:, rows[sel]]
                    err = max(err, float(np.abs(R_loc[:, sel] - ref).max() / max(np.abs(ref).max(), 1e-300)))
            nrows += len(rows); nwall += int(sol_nodes.sum()); mult = max(mult, int(lb.multiplicity(rank)[loc].max()))
        out["parity"] = {"oracle": kind, "rows_rank0": nrows, "rows_on_wall_rank0": nwall, "max_ranks_on_a_row_rank0": mult,
                         "R_max_rel": reduce_ranks(err, "max"), "tol": 1e-12}
        out["parity"]["ok"] = out["parity"]["R_max_rel"] < 1e-12
    ls = abi.ls_params(abi.LS_GMRES, mItr=1, sD=50, relTol=1e-12)
    gm_ms = 0.0
    for rep in range(2):
        fsi()
        eng.timer_mark(0)
        _, o, _ = eng.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1), want_solution=False)
        eng.timer_mark(1)
        gm_ms = reduce_ranks(eng.timer_elapsed() / max(o.RI.itr, 1), "max")
    eqm, dmm = abi.mesh_eq(1e-3), [abi.mesh_domain(E=1.0, nu=0.3)]

    def msh():
        eng.alloc(3); eng.assemble(3, eqm, dmm); eng.commu_R()
    eng.alloc(3); eng.set_old_disp(np.asfortranarray(0.9 * Dg)); msh()
    msh_ms = reduce_ranks(_timed(eng, msh, 3), "max")
    nEl = int(reduce_ranks(float(m.nEl), "sum"))
    out.update({"workload": f"FSI pipe on {world} GPUs, {nEl} tet4 (lumen ALE VMS fluid + nHK M94 wall outside 0.55 R), tDof 7, on the "
                            "partition of the headline run",
                "value": nEl / (fsi_ms * 1e-3), "unit": "element assemblies/s (construct_fsi: zero + lumen + wall + shared-node sum)",
                "construct_fsi_stage_ms": fsi_ms, "construct_mesh_stage_ms": msh_ms, "mesh_value": nEl / (msh_ms * 1e-3),
                "fsi_gmres_ms_per_itr": gm_ms, "fluid_elements_rank0": int(len(fl)), "solid_elements_rank0": int(len(so))})
    return out


def extra_configs(args, device, sampler, fp64_peak, hbm_peak):
    out = {}
    for name, fn in (("C4_struct_hex8", lambda: config_c4_struct(device, fp64_peak, hbm_peak)),
                     ("C5_fsi_pipe", lambda: config_c5_fsi(device, hbm_peak))):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as ex:       # an extra; never hide the headline
            out[name] = {"error": repr(ex)}
        out[name]["clocks"] = sampler.window(t0, time.perf_counter())
        out[name]["wall_s"] = time.perf_counter() - t0
    return out


# ------------------------------------------------------------------------------------------------
def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=118, help="cross-section lattice (hexes per side)")
    ap.add_argument("--nz", type=int, default=120, help="cell layers per GPU share")
    ap.add_argument("--ls-reltol", type=float, default=1e-3)
    ap.add_argument("--ls-sd", type=int, default=50)
    ap.add_argument("--ls-mitr", type=int, default=4)
    ap.add_argument("--partition", default="auto", choices=["auto", "slab", "blocks"],
                    help="element partition over the GPUs: z-slabs, or x/y/z blocks (2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2: up to 7 "
                         "neighbours per rank, nodes shared by up to 8 ranks); auto = blocks from 4 GPUs on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to the CPUs / memory of its GPU's NUMA node")
    ap.add_argument("--scatter", default="atomic", choices=["atomic", "colored"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = workload_config(args)
    metric = "element assemblies/s (FP64 tet4 fluid)"

    if args.impl == "reference":
        if rank != 0:
            return
        # the driver's --steps K --warmup W are honoured: every step runs ls_alloc + construct_fluid + commu(R) on the whole
        # C2 mesh over all host cores (about 3 s); the fsils_solve (about 5-10 s) only in the first few, so K + W steps end
        # within a few minutes
        steps, warm = max(1, min(args.steps, 50)), max(0, min(args.warmup, 10))
        r = reference_newton(args, steps, warm, solve_steps=min(steps, 3))
        line = {"impl": "reference", "metric": metric, "value": r["value"],
                "unit": "element assemblies/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": r["newton_step_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "newton_step_ms": r["newton_step_ms"], "assembly_stage_ms": r["assemble_ms"], "solve_ms": r["solve_ms"],
                "gmres": r.get("gmres"),
                "cpu_baseline": {k: r.get(k) for k in ("value", "unit", "cores", "kind", "sample", "assemble_ms", "solve_ms",
                                                       "newton_step_ms", "gmres", "partition_blocks")},
                "e2e": {"value": r["value"], "unit": "element assemblies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit_line(line)
        return

    import torch
    import torch.distributed as dist
    from svmultiphysics_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if not args.no_numa_bind else {"bound": False, "why": "--no-numa-bind"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(v, op="max"):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    max_over_ranks = reduce_ranks

    # ---- build this rank's partition ------------------------------------------------------------
    pmode = args.partition if args.partition != "auto" else ("blocks" if world >= 4 else "slab")
    blocks = meshgen.default_blocks(world, pmode)
    n, nzg, L = args.n, args.nz * world, LSEG * world
    lb = partition.LatticeBlocks((n, n, nzg), blocks)
    m = meshgen.cylinder_block(n, nzg, blocks, rank, L=L)
    Ag, Yg = lattice_state(m)
    eng = Engine(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(Engine.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().tolist()))
    rowPtr, colPtr = eng.lhsa(m.nNo, [m.IEN])
    node_map, mynNo = lb.order(rank)
    neigh = lb.neighbours(rank)
    if world > 1:
        eng.set_graph(rowPtr, colPtr, mynNo=mynNo, node_map=node_map, neighbours=neigh)
    else:
        eng.set_graph(rowPtr, colPtr)
    mult = lb.multiplicity(rank)
    run = {"partition": {"slab": "z-slabs", "blocks": "x/y/z blocks"}[pmode] + f" {blocks[0]}x{blocks[1]}x{blocks[2]}, one per GPU "
                        "(element partition with duplicated interface nodes, FSILS node order; stand-in for ParMETIS, which needs MPI)",
           "blocks": list(blocks), "transport": eng.comm_transport(), "scatter": args.scatter,
           "neighbours_per_rank": int(reduce_ranks(float(len(neigh)), "max")),
           "shared_nodes_per_rank": int(reduce_ranks(float((mult > 1).sum()), "max")),
           "max_ranks_sharing_a_node": int(reduce_ranks(float(mult.max()), "max")), "numa": numa}
    w, N, Nx = elements.tables(4)
    eng.set_mesh(0, m.IEN, w, N, Nx)
    eng.set_coords(m.x)
    wall = m.faces["wall"]
    eng.set_num_faces(1)
    eng.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"), shared=int(world > 1))
    eng.alloc(4)
    eng.set_state(Ag, Yg)
    eq = abi.fluid_eq(1e-3, scatter=abi.SCATTER_ATOMIC if args.scatter == "atomic" else abi.SCATTER_COLORED)
    dmn = [abi.fluid_domain()]
    ls = ls_config(args)
    incL, res = np.ones(1, np.int32), np.zeros(1)
    nEl_total = int(reduce_ranks(float(m.nEl), "sum"))

    # device-resident generalised-alpha state: old = (Ao, Yo) chosen so that predictor + initiator reproduce exactly
    # the (Ag, Yg) the reference arm assembles with; every bench step is then the first Newton iteration of a time
    # step: predictor -> initiator -> ls_alloc -> assembly -> halo -> fsils_solve -> corrector.
    qt = [abi.eq_time(0, 3, abi.PHYS_FLUID, 0.5)]
    cA = (1.0 - qt[0].am) + qt[0].am * (qt[0].gam - 1.0) / qt[0].gam
    eng.set_solution(abi.SOL_OLD, np.asfortranarray(Ag / cA), Yg, np.zeros_like(Yg))

    def newton_step(stats=None):
        eng.predictor(qt, 1e-3, 0)
        eng.initiator(qt)
        eng.timer_mark(0)
        eng.alloc(4)
        eng.assemble(0, eq, dmn)
        eng.commu_R()
        eng.timer_mark(1)
        t_asm = eng.timer_elapsed()
        k_asm = eng.last_timing()[0]
        _, out, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res, want_solution=False)
        eng.corrector(qt[0], 1e-3)
        if stats is not None:
            stats.append((t_asm, k_asm, eng.last_timing()[1], out.RI.itr, out.RI.success, out.RI.iNorm, out.RI.fNorm))

    for _ in range(args.warmup):
        newton_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    stats = []
    launches0 = eng.launch_count
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        newton_step(stats)
    barrier()
    t1 = time.perf_counter()
    launches = eng.launch_count - launches0
    clocks = sampler.window(t0, t1) if rank == 0 else None
    step_ms = max_over_ranks((t1 - t0) * 1e3 / args.steps)
    asm_ms = max_over_ranks(float(np.mean([s[0] for s in stats])))
    kern_ms = max_over_ranks(float(np.mean([s[1] for s in stats])))
    solve_ms = max_over_ranks(float(np.mean([s[2] for s in stats])))

    # ---- e2e: assembly stage with HOST buffers through the C ABI ------------------------------------
    Ah, Yh = np.asfortranarray(Ag.copy()), np.asfortranarray(Yg.copy())
    Rh = np.zeros((4, m.nNo), order="F")
    for a in (Ah, Yh, Rh):
        eng.pin(a)

    def e2e_step():
        # svb200_assemble_host = set_state + alloc + assemble + commu_R + download(R), pipelined (H2D in node chunks on a copy
        # stream behind which the element groups start, finished residual rows streamed back on a single partition)
        eng.assemble_host(0, eq, dmn, Ah, Yh, Rh)

    for _ in range(max(args.warmup, 1)):
        e2e_step()
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    te1 = time.perf_counter()
    e2e_ms = max_over_ranks((te1 - te0) * 1e3 / args.steps)
    e2e_timeline = eng.last_host_stage()          # rank 0, last step
    # the PCIe floor of the same call: only its copies (H2D of Ag, Yg; D2H of R), all ranks at once — H2D and D2H are full duplex,
    # so the stage cannot be shorter than max(h2d, d2h, kernel)
    def timed(fn, reps=3):
        fn()
        barrier()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e3 / reps
        barrier()
        return max_over_ranks(dt)
    h2d_ms = timed(lambda: eng.set_state(Ah, Yh))
    d2h_ms = timed(lambda: eng.download_into(abi.ARRAY_R, Rh))
    # what the pipelined call returned against the plain sequence of calls (device-resident R after the shared-node sum)
    eng.set_state(Ag, Yg); eng.alloc(4); eng.assemble(0, eq, dmn); eng.commu_R()
    R_plain = eng.get_R()
    e2e_err = reduce_ranks(float(np.abs(Rh - R_plain).max() / max(np.abs(R_plain).max(), 1e-300)), "max")
    del R_plain
    for a in (Ah, Yh, Rh):
        eng.unpin(a)

    # ---- rooflines --------------------------------------------------------------------------------
    fp64_peak = eng.fp64_peak()
    spmv_ms = eng.bench_spmv(4, 20)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    asm_tf = m.nEl * FLOP_PER_ELEMENT / (kern_ms * 1e-3) * 1e-12
    # DRAM traffic per launch of the two kernels from the committed `ncu --set full` captures (profiles/), valid
    # for the default C2 size only
    traffic = {}
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if t.get("elements") == m.nEl:
            traffic = t
    except Exception:
        pass
    spmv_gbs = SPMV_BYTES(len(colPtr), m.nNo) / (spmv_ms * 1e-3) * 1e-9
    # the other scatter mode of the same assembly, kernel only (deterministic = graph-coloured groups, bitwise reproducible)
    other = abi.SCATTER_COLORED if args.scatter == "atomic" else abi.SCATTER_ATOMIC
    eq_other = abi.fluid_eq(1e-3, scatter=other)
    eng.alloc(4)
    eng.bench_assemble(0, eq_other, dmn, 1)
    other_ms = max_over_ranks(eng.bench_assemble(0, eq_other, dmn, 5))

    # ---- parity gate (never timed) ------------------------------------------------------------------
    parity = None
    if not args.no_parity:
        eng.set_state(Ag, Yg)
        eng.alloc(4)
        eng.assemble(0, eq, dmn)
        eng.commu_R()
        pa = parity_assembly(eng, m, lb, rank, rowPtr, colPtr, eq, dmn, n, nzg, L)
        R_loc = eng.get_R()
        X_loc, o, _ = eng.solve(4, abi.LS_GMRES, ls, incL, res)
        W_loc = eng.get_W()
        eng.alloc(4)
        eng.assemble(0, eq, dmn)                     # the solve preconditioned Val in place: the unscaled matrix again
        KX = eng.spmv(4, X_loc)                      # K x incl. the shared-node sum
        owned = node_map < mynNo
        rr = (W_loc * (R_loc - KX))[:, owned]
        r0 = (W_loc * R_loc)[:, owned]
        nr = float(np.sqrt(reduce_ranks(float((rr * rr).sum()), "sum")))
        n0 = float(np.sqrt(reduce_ranks(float((r0 * r0).sum()), "sum")))
        asm_err = reduce_ranks(pa["max_rel"], "max")
        ratio = nr / (args.ls_reltol * n0)
        # per-box detail of the rank with the largest number of ranks on a compared row (rank 0 otherwise)
        solve_ok = bool(o.RI.success) and ratio <= 1.05 and abs(n0 - o.RI.iNorm) <= 1e-10 * n0 and abs(nr - o.RI.fNorm) <= 0.05 * nr
        ok_local = float(pa["max_rel"] < 1e-12 and all(b["same_graph"] for b in pa["boxes"].values()) and solve_ok and e2e_err < 1e-12)
        ok = reduce_ranks(ok_local, "min") == 1.0
        parity = {"ok": ok, "assembly_max_rel": asm_err, "assembly_tol": 1e-12, "oracle": pa["oracle"],
                  "solve_true_res_ratio": ratio, "solve_true_res": nr, "solve_reported_fNorm": o.RI.fNorm,
                  "solve_iNorm_rel_diff": abs(n0 - o.RI.iNorm) / n0, "gmres_itr": o.RI.itr,
                  "what": "R (after the shared-node sum) and Val rows of two sub-boxes per rank vs the oracle on the same mesh/state, "
                          "max over ranks; ||W(R-Kx)|| / (relTol ||WR||) over all ranks",
                  "rank0_boxes": pa["boxes"]}

    if rank == 0:
        line = {
            "metric": metric, "value": nEl_total / (asm_ms * 1e-3),
            "unit": "element assemblies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "run": run,
            "newton_step_ms": step_ms, "assembly_stage_ms": asm_ms, "assembly_kernel_ms": kern_ms, "solve_ms": solve_ms,
            "gmres": {"itr": stats[-1][3], "success": int(stats[-1][4]), "iNorm": stats[-1][5], "fNorm": stats[-1][6],
                      "ms_per_iteration": solve_ms / max(stats[-1][3], 1)},
            "parity": parity,
            "roofline": {"bound": "fp64", "achieved": asm_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": asm_tf / fp64_peak,
                         "traffic": traffic.get("assemble_bytes"), "traffic_unit": "bytes/launch (ncu dram read+write)",
                         "kernel": "assemble_fluid_tet4_grouped_kernel" if args.scatter == "atomic" else "assemble_fluid_tet4_kernel",
                         "hbm_check": {"algorithmic_bytes_per_element": 730, "note": "compulsory DRAM bytes (Val RMW + plan + gather) "
                                       "x elements / kernel time vs HBM peak", "GB/s": m.nEl * 730 / (kern_ms * 1e-3) * 1e-9,
                                       "frac_of_hbm_peak": m.nEl * 730 / (kern_ms * 1e-3) * 1e-9 / hbm_peak},
                         "peak_source": "FP64 FMA peak measured in this run by svb200_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 figure)",
                         "algorithmic": f"{FLOP_PER_ELEMENT:.0f} flop/element x {m.nEl} elements per launch"},
            "roofline_spmv": {"bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                              "traffic": traffic.get("spmv_bytes"), "kernel": "bsr_spmv4_kernel", "peak_source": hbm_src,
                              "algorithmic": "nnz*132 + nNo*72 bytes per launch", "ms": spmv_ms},
            "e2e": {"value": nEl_total / (e2e_ms * 1e-3), "unit": "element assemblies/s",
                    "h2d_bytes_per_step": int(Ah.nbytes + Yh.nbytes), "d2h_bytes_per_step": int(Rh.nbytes), "ms_per_step": e2e_ms,
                    "call": "svb200_assemble_host (pinned host Ag, Yg in; R out; copies pipelined behind the element kernel)",
                    "R_max_rel_vs_plain_sequence": e2e_err, "device_timeline_ms_rank0": e2e_timeline,
                    "copies_alone_ms": {"h2d": h2d_ms, "d2h": d2h_ms, "GB/s_h2d_per_gpu": (Ah.nbytes + Yh.nbytes) / h2d_ms * 1e-6,
                                        "GB/s_d2h_per_gpu": Rh.nbytes / d2h_ms * 1e-6,
                                        "note": "the same pinned buffers copied with no kernel running, all ranks at once (max over "
                                                "ranks): the PCIe / host-memory floor of the host-buffer stage on this box"}},
            "gpu_launches": int(launches), "clocks": clocks,
            "other_scatter_mode": {"scatter": "colored (deterministic)" if args.scatter == "atomic" else "atomic",
                                   "assembly_kernel_ms": other_ms, "value": nEl_total / (other_ms * 1e-3),
                                   "unit": "element assemblies/s (kernel only, outside the timed steps)"},
        }
    configs = {}
    if world == 1 and not args.no_extra_configs:
        t0c = time.perf_counter()
        try:
            configs["C1_ns_solver"] = config_c1_ns(eng, m, Ag, Yg, eq, dmn, hbm_peak, len(colPtr))
        except Exception as ex:
            configs["C1_ns_solver"] = {"error": repr(ex)}
        configs["C1_ns_solver"]["clocks"] = sampler.window(t0c, time.perf_counter())
        configs["C1_ns_solver"]["wall_s"] = time.perf_counter() - t0c
    c5m = None
    if world > 1 and not args.no_extra_configs:
        t0c = time.perf_counter()
        c5m = config_c5_fsi_multi(eng, m, lb, rank, world, rowPtr, colPtr, n, nzg, L, reduce_ranks, not args.no_parity)
        c5m["wall_s"] = time.perf_counter() - t0c
        if rank == 0:
            c5m["clocks"] = sampler.window(t0c, time.perf_counter())
            line["configs"] = {"C5_fsi_pipe_multi_gpu": c5m}
        if parity is not None and "parity" in c5m:
            parity["c5_fsi_R_max_rel"] = c5m["parity"]["R_max_rel"]
            parity["ok"] = bool(parity["ok"] and c5m["parity"]["ok"])
    eng.close()
    if rank == 0:
        if world == 1 and not args.no_extra_configs:
            configs.update(extra_configs(args, local_rank, sampler, fp64_peak, hbm_peak))
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_newton(args, steps=2, warmup=1, solve_steps=1)
                line["cpu_baseline"] = {k: r.get(k) for k in ("value", "unit", "cores", "kind", "sample", "assemble_ms", "solve_ms",
                                                              "newton_step_ms", "gmres", "partition_blocks")}
            except Exception as ex:   # the baseline is a reported extra; never hide the GPU result
                line["cpu_baseline"] = {"error": repr(ex)}
        sampler.stop()
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY GATE FAILED: " + json.dumps(parity) + "\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
