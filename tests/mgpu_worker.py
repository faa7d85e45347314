"""Multi-GPU parity worker (launched by torchrun, one rank per GPU): the global mesh is split by a general
element partition, every rank assembles + halo-sums + solves its part through the C ABI, and rank 0 compares
the glued result with the single-partition oracle.  Exit code 0 = parity holds."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svmultiphysics_b200 import abi, elements, partition  # noqa: E402
from svmultiphysics_b200.engine import Engine  # noqa: E402
from tests import common  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "slab"
    ls_name = sys.argv[2] if len(sys.argv) > 2 else "gmres"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n, nz = 6, 8
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=n, nz=nz)
    if mode == "slab":
        part = (((np.arange(m.nEl) // 6) // (n * n)) * world // nz).astype(np.int32)
    else:   # scattered blocks: many neighbours, nodes shared by more than two ranks
        part = (((np.arange(m.nEl) // 6) * 7919) % world).astype(np.int32)
    parts = partition.partition_mesh(m.IEN, m.nNo, part, world)
    p = parts[rank]
    eng = Engine(lr)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(Engine.unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    eng.comm_init(world, rank, bytes(uid.cpu().tolist()))
    rowPtr, colPtr = eng.lhsa(p.nNo, [p.IEN])
    eng.set_graph(rowPtr, colPtr, mynNo=p.mynNo, node_map=p.node_map, neighbours=p.neighbours)
    w, N, Nx = elements.tables(4)
    eng.set_mesh(0, p.IEN, w, N, Nx)
    eng.set_coords(m.x[:, p.ltg])
    # faces: Dirichlet wall + inlet, coupled resistance outlet; local node lists
    gfaces = common.dirichlet_faces(m)
    out = m.faces["outlet"]
    val = np.zeros((3, len(out)), order="F"); val[2] = 4.0 * np.pi / len(out)
    gfaces.append((abi.BC_NEU, out, val))
    gtl = -np.ones(m.nNo, dtype=np.int64); gtl[p.ltg] = np.arange(p.nNo)
    eng.set_num_faces(len(gfaces))
    count = np.zeros(m.nNo, dtype=np.int32)
    for q in parts:
        count[q.ltg] += 1
    for i, (g, nodes, v) in enumerate(gfaces):
        mine = gtl[nodes] >= 0
        # nodal normal integrals of a shared face node are PARTIAL on each rank (summed by fsils_bc_create)
        vloc = v[:, mine] / count[nodes[mine]]
        eng.set_face(i, g, gtl[nodes[mine]].astype(np.int32), np.asfortranarray(vloc), shared=1)
    eng.alloc(4)
    eng.set_state(Ag[:, p.ltg], Yg[:, p.ltg], None, Bf[:, p.ltg])
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    eng.assemble(0, eq, dmn)
    eng.commu_R()
    R_loc = eng.get_R()
    if ls_name == "ns":
        ls_type = abi.LS_NS
        ls = abi.ls_params(abi.LS_NS, mItr=15, sD=250, relTol=1e-3, absTol=1e-17, gm=(10, 250, 1e-3, 1e-17), cg=(300, 0, 1e-3, 1e-17))
    else:
        ls_type = abi.LS_GMRES
        ls = abi.ls_params(abi.LS_GMRES, mItr=100, sD=50, relTol=1e-8)
    incL, res = np.ones(3, np.int32), np.array([0.0, 0.0, 0.8])
    X_loc, o, _ = eng.solve(4, ls_type, ls, incL, res)
    # gather on rank 0
    def gather(a):
        t = torch.zeros((4, m.nNo), dtype=torch.float64, device="cuda")
        t[:, torch.from_numpy(p.ltg).cuda()] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        owner = torch.zeros(m.nNo, dtype=torch.float64, device="cuda")
        owner[torch.from_numpy(p.ltg).cuda()] = 1.0
        dist.all_reduce(t); dist.all_reduce(owner)
        return (t / owner).cpu().numpy()
    Rg, Xg = gather(R_loc), gather(X_loc)
    ok = 1
    if rank == 0:
        from oracle import refbind
        cls = refbind.RefCase if refbind.have_ref() else refbind.OracleCase
        orc, rp, cp = common.make_oracle(cls, m, nFaces=len(gfaces))
        for i, (g, nodes, v) in enumerate(gfaces):
            orc.set_face(i, g, nodes, v)
        orc.alloc(4); orc.set_state(Ag, Yg, Dg, Bf); orc.assemble(0, eq, dmn)
        R0 = orc.get_R()
        X0, o0, _ = orc.solve(4, ls_type, ls, incL, res)
        eR, eX = common.rel_err(Rg, R0), common.rel_err(Xg, X0)
        tolX = 0.05 if ls_type == abi.LS_NS else 1e-6
        print(f"[mgpu {mode}/{ls_name} x{world}] relerr R {eR:.2e} X {eX:.2e}; itr {o.RI.itr} vs {o0.RI.itr}; iNorm {o.RI.iNorm:.6e} vs {o0.RI.iNorm:.6e}")
        ok = int(eR < 1e-12 and eX < tolX and abs(o.RI.iNorm - o0.RI.iNorm) < 1e-10 * o0.RI.iNorm
                 and abs(o.RI.itr - o0.RI.itr) <= max(2, o0.RI.itr // 25))
    flag = torch.tensor([ok], device="cuda")
    dist.broadcast(flag, 0)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
