#!/usr/bin/env python
"""bench.py — element assemblies/s (FP64 tet4 fluid) and Newton-step time on N B200s.

Contract:
  python bench.py --gpus N --steps K --warmup W            our arm (libsvb200.so through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's own CPU code on the host cores

One "step" is one Newton iteration of the hot path on the synthetic 10 M-tet4 cylinder (config C2 of
SURVEY.md §8d), one such cylinder slab per GPU (weak scaling):
    predictor/initiator -> ls_alloc (zero R, Val) -> element assembly + scatter -> shared-node sum of R ->
    fsils_solve (GMRES) -> corrector, all on the device (no nodal array crosses PCIe inside the step).
`value` is elements assembled per second over the ASSEMBLY stage of the timed steps (zero + kernel +
halo, device-resident inputs, CUDA events on the library's stream, max over ranks); `ms_per_step` is
the whole Newton step; `e2e` is the assembly stage driven with HOST buffers through the C ABI (H2D
of Ag/Yg from pinned memory and D2H of the residual inside the timed region).
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


def emit_line(obj):
    if _REAL_STDOUT is None:
        print(json.dumps(obj), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


from svmultiphysics_b200 import abi, elements, meshgen, partition  # noqa: E402

FLOP_PER_ELEMENT = 11.6e3     # algorithmic FP64 flop per tet4 VMS element (Newtonian)
SPMV_BYTES = lambda nnz, nNo: nnz * 132 + nNo * 72   # noqa: E731  SURVEY.md §8(d), dof = 4


def lattice_state(m, rank, nz, U=10.0, R=2.0, tDof=4, noise=0.01):
    """Poiseuille flow + 1 % deterministic pseudo-noise keyed on the GLOBAL lattice index, so that the
    nodes two slabs share carry identical state on both ranks (seeds replaced by a hash)."""
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

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
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


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref/libsvref.so) on the host cores
# ------------------------------------------------------------------------------------------------
def _ref_worker(q, n, nz, rank, nranks, steps, warmup, ls_tuple, with_solve):
    from oracle import refbind
    cls, kind = (refbind.RefCase, "reference") if refbind.have_ref() else (refbind.OracleCase, "port")
    m, other, plo, phi = meshgen.cylinder_slab(n, nz, rank, nranks)
    Ag, Yg = lattice_state(m, rank, nz)
    c = cls()
    c.set_coords(m.x)
    c.add_mesh(m.IEN)
    wall = m.faces["wall"]
    c.build_graph(1)
    c.set_face(0, abi.BC_DIR, wall, np.zeros((3, len(wall)), order="F"))
    eq, dmn = abi.fluid_eq(1e-3), [abi.fluid_domain()]
    ts, tsolve, itr = [], [], 0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        c.alloc(4)
        c.set_state(Ag, Yg) if s == 0 else None
        c.assemble(0, eq, dmn)
        t1 = time.perf_counter()
        if s >= warmup:
            ts.append(t1 - t0)
        if with_solve and s >= warmup:
            ls = abi.ls_params(abi.LS_GMRES, mItr=ls_tuple[0], sD=ls_tuple[1], relTol=ls_tuple[2])
            t2 = time.perf_counter()
            _, out, _ = c.solve(4, abi.LS_GMRES, ls, np.ones(1, np.int32), np.zeros(1))
            tsolve.append(time.perf_counter() - t2)
            itr = out.RI.itr
    q.put((m.nEl, ts, tsolve, itr, kind))


def run_reference_arm(args, sample_n=36, sample_nz=24, procs=None, with_solve=True):
    """All host cores, one independent mesh partition per process (what the reference's MPI ranks do in
    construct_fluid, which has no inter-rank communication).  Bounded sample: 6*n*n*nz tets per core."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ls_tuple = (args.ls_mitr, args.ls_sd, args.ls_reltol)
    steps = getattr(args, "ref_steps_eff", args.steps_ref)
    warm = getattr(args, "ref_warmup_eff", 1)
    ps = [ctx.Process(target=_ref_worker, args=(q, sample_n, sample_nz, r, procs, steps, warm, ls_tuple, with_solve and r == 0))
          for r in range(procs)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    nEl = sum(r[0] for r in res)
    steps = len(res[0][1])
    t_step = [max(r[1][s] for r in res) for s in range(steps)]        # max over "ranks" per step
    t_asm = float(np.mean(t_step))
    solve = [r for r in res if r[2]]
    out = {"value": nEl / t_asm, "unit": "element assemblies/s", "cores": procs, "kind": res[0][4],
           "sample": f"{procs} independent cylinder slabs of {res[0][0]} tet4 each (6*{sample_n}^2*{sample_nz}), "
                     f"ls_alloc + construct_fluid per step, mean of {steps} steps, max over processes",
           "ms_per_step": t_asm * 1e3, "elements": nEl}
    if solve:
        out["newton_step_1core"] = {"elements": solve[0][0], "assemble_ms": float(np.mean(solve[0][1])) * 1e3,
                                    "fsils_solve_ms": float(np.mean(solve[0][2])) * 1e3, "gmres_itr": solve[0][3]}
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
    ap.add_argument("--nz", type=int, default=120, help="cell layers per GPU slab")
    ap.add_argument("--ls-reltol", type=float, default=1e-3)
    ap.add_argument("--ls-sd", type=int, default=50)
    ap.add_argument("--ls-mitr", type=int, default=4)
    ap.add_argument("--steps-ref", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scatter", default="atomic", choices=["atomic", "colored"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"C2 synthetic cylinder, 6*{args.n}^2*{args.nz} = {6*args.n*args.n*args.nz} tet4 per GPU, Newtonian VMS fluid"
    cfg = {"workload": workload, "elements_per_gpu": 6 * args.n * args.n * args.nz, "dt": 1e-3,
           "linear_solver": f"GMRES sD={args.ls_sd} mItr={args.ls_mitr} relTol={args.ls_reltol} + FSILS diagonal preconditioner",
           "partition": "z-slabs, one per GPU, shared interface planes (stand-in for ParMETIS, which needs MPI)",
           "scatter": args.scatter,
           "l2": "inputs larger than L2 (Val = 3.2 GB/GPU is rewritten every step)"}

    if args.impl == "reference":
        if rank != 0:
            return
        # the driver's --steps K --warmup W are honoured; each step is a bounded sample (186,624 tet4 per core,
        # about 1 s of construct_fluid), so K + W steps end within a few minutes
        args.ref_steps_eff, args.ref_warmup_eff = max(1, min(args.steps, 50)), max(0, min(args.warmup, 10))
        r = run_reference_arm(args)
        line = {"impl": "reference", "metric": "element assemblies/s (FP64 tet4 fluid)", "value": r["value"],
                "unit": "element assemblies/s", "n_gpus": args.gpus, "steps": args.ref_steps_eff, "warmup": args.ref_warmup_eff,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "element assemblies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "newton_step_1core": r.get("newton_step_1core")}
        emit_line(line)
        return

    import torch
    import torch.distributed as dist
    from svmultiphysics_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- build this rank's partition ------------------------------------------------------------
    m, other, plane_lo, plane_hi = meshgen.cylinder_slab(args.n, args.nz, rank, world)
    Ag, Yg = lattice_state(m, rank, args.nz)
    eng = Engine(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(Engine.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().tolist()))
    rowPtr, colPtr = eng.lhsa(m.nNo, [m.IEN])
    if world > 1:
        node_map, mynNo = partition.fsils_order(other, rank)
        neigh = []
        if rank > 0:
            neigh.append((rank - 1, node_map[plane_lo]))
        if rank < world - 1:
            neigh.append((rank + 1, node_map[plane_hi]))
        eng.set_graph(rowPtr, colPtr, mynNo=mynNo, node_map=node_map, neighbours=neigh)
    else:
        eng.set_graph(rowPtr, colPtr)
    cfg["transport"] = eng.comm_transport()     # "p2p": halo sums / scalar all-reduces by the library's own peer-memory kernels
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
    nEl_total = m.nEl * world

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
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    step_ms = max_over_ranks((t1 - t0) * 1e3 / args.steps)
    asm_ms = max_over_ranks(float(np.mean([s[0] for s in stats])))
    kern_ms = max_over_ranks(float(np.mean([s[1] for s in stats])))
    solve_ms = max_over_ranks(float(np.mean([s[2] for s in stats])))

    # ---- e2e: assembly stage with HOST buffers through the C ABI ------------------------------------
    Ah, Yh = np.asfortranarray(Ag.copy()), np.asfortranarray(Yg.copy())
    Rh = np.zeros((4, m.nNo), order="F")
    for a in (Ah, Yh, Rh):
        eng.pin(a)
    import ctypes as C

    def e2e_step():
        eng.set_state(Ah, Yh)
        eng.alloc(4)
        eng.assemble(0, eq, dmn)
        eng.commu_R()
        eng._call("svb200_download", C.c_int32(abi.ARRAY_R), Rh.ctypes.data_as(C.POINTER(C.c_double)))

    for _ in range(max(args.warmup, 1)):
        e2e_step()
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    te1 = time.perf_counter()
    e2e_ms = max_over_ranks((te1 - te0) * 1e3 / args.steps)
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

    if rank == 0:
        line = {
            "metric": "element assemblies/s (FP64 tet4 fluid)", "value": nEl_total / (asm_ms * 1e-3),
            "unit": "element assemblies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "newton_step_ms": step_ms, "assembly_stage_ms": asm_ms, "assembly_kernel_ms": kern_ms, "solve_ms": solve_ms,
            "gmres": {"itr": stats[-1][3], "success": int(stats[-1][4]), "iNorm": stats[-1][5], "fNorm": stats[-1][6]},
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
                    "h2d_bytes_per_step": int(Ah.nbytes + Yh.nbytes), "d2h_bytes_per_step": int(Rh.nbytes), "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks,
            "other_scatter_mode": {"scatter": "colored (deterministic)" if args.scatter == "atomic" else "atomic",
                                   "assembly_kernel_ms": other_ms, "value": nEl_total / (other_ms * 1e-3),
                                   "unit": "element assemblies/s (kernel only, outside the timed steps)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = run_reference_arm(args, sample_n=30, sample_nz=16, procs=1)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["newton_step_1core"] = r.get("newton_step_1core")
            except Exception as ex:   # the baseline is a reported extra; never hide the GPU result
                line["cpu_baseline"] = {"error": repr(ex)}
        emit_line(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
