"""One rank of the multi-rank REFERENCE run (TEST INFRASTRUCTURE): the compiled reference (oracle/_ref/libsvref.so) with the
shared-memory MPI shim of oracle/ref_build/mpi_stub.cpp, SVREF_MPI_{SIZE,RANK,SHM} set by the parent test.  The rank builds its
local mesh with the product's host logic (svmultiphysics_b200/partition.py), lets the reference's own fsils_lhs_create derive
lhs.map / mynNo / cS[] from the local -> global map, compares them with partition.py's, then assembles, sums shared nodes and
solves with the reference's multi-rank FSILS; results go to <outdir>/rank<r>.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from svmultiphysics_b200 import abi, partition  # noqa: E402
from tests import common  # noqa: E402


def struct_main(outdir, mode, rank, world):
    """HEX8 solid (struct_3d, dof 3, BiCGStab like tests/cases/struct/block_compression) on a METIS or slab partition."""
    m, Ag, Yg, Dg, Bf, faces_g, eq, dmn, ls = common.mrank_struct_case()
    if mode == "metis":
        from oracle import metis_part
        part, _ = metis_part.part_mesh_dual(m.IEN, m.nNo, world)
    else:
        cz = m.x[2, m.IEN].mean(axis=0)
        part = np.minimum((cz / cz.max() * world * 0.999).astype(np.int32), world - 1)
    parts = partition.partition_mesh(m.IEN, m.nNo, part, world)
    p = parts[rank]
    count = np.zeros(m.nNo, dtype=np.int32)
    for q in parts:
        count[q.ltg] += 1
    faces = []
    for (g, nodes, val) in faces_g:
        loc = np.searchsorted(p.ltg, nodes)
        ok = (loc < p.nNo) & (p.ltg[np.minimum(loc, p.nNo - 1)] == nodes)
        # fsils_bc_create SUMS the face values of a node over the ranks that share it (bc.cpp:68-88), Dirichlet faces included:
        # the 1.0 that fsi_ls_ini passes for the free directions of an effective-direction Dirichlet face (baf_ini.cpp:764-772)
        # becomes 2.0 on a node shared by two ranks in a real MPI run, which rescales W there (same solution, different iNorm /
        # iteration count than a single-rank run).  To compare with the single-rank run the values are passed as partial sums.
        faces.append((g, loc[ok].astype(np.int32), np.asfortranarray(val[:, ok] / count[nodes[ok]])))
    c = refbind.RefCase()
    c.set_coords(m.x[:, p.ltg]); c.set_partition(m.nNo, p.ltg); c.add_mesh(p.IEN)
    c.build_graph(len(faces))
    mynNo, lmap, reqs = c.get_lhs()
    same_map = bool(np.array_equal(lmap, p.node_map)) and mynNo == p.mynNo
    same_nb = [q for q, _ in reqs] == [q for q, _ in p.neighbours]
    same_ptr = same_nb and all(np.array_equal(a[1], b[1]) for a, b in zip(reqs, p.neighbours))
    for i, (g, nodes, val) in enumerate(faces):
        c.set_face(i, g, nodes, val)
    c.alloc(3); c.set_state(Ag[:, p.ltg], Yg[:, p.ltg], Dg[:, p.ltg], Bf[:, p.ltg]); c.assemble(0, eq, dmn)
    c.commu_R()
    R = c.get_R()
    X, o, _ = c.solve(3, abi.LS_BICGS, ls, np.ones(len(faces), np.int32), np.zeros(len(faces)))
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), ltg=p.ltg, R=R, X=X, itr=o.RI.itr, iNorm=o.RI.iNorm, fNorm=o.RI.fNorm,
             success=o.RI.success, same_map=same_map, same_nb=same_nb, same_ptr=same_ptr, same_sets=same_ptr, mynNo=mynNo)


def main():
    outdir, mode = sys.argv[1], sys.argv[2]
    ls_name = sys.argv[3] if len(sys.argv) > 3 else "gmres"
    rank, world = int(os.environ["SVREF_MPI_RANK"]), int(os.environ["SVREF_MPI_SIZE"])
    if ls_name == "struct":
        return struct_main(outdir, mode, rank, world)
    n, nz = 4, 6
    m, Ag, Yg, Dg, Bf = common.fluid_case(n=n, nz=nz)
    if mode == "metis":
        from oracle import metis_part
        part, _ = metis_part.part_mesh_dual(m.IEN, m.nNo, world)
    elif mode == "scattered":      # hexes dealt round-robin: many neighbours, nodes shared by three and more ranks
        part = (((np.arange(m.nEl) // 6) * 7919) % world).astype(np.int32)
    elif mode == "random":
        part = np.random.default_rng(77).integers(0, world, m.nEl).astype(np.int32)
    else:
        part = (((np.arange(m.nEl) // 6) // (n * n)) * world // nz).astype(np.int32)
    parts = partition.partition_mesh(m.IEN, m.nNo, part, world)
    p = parts[rank]
    gfaces, res = common.mrank_faces(m, ls_name)
    count = np.zeros(m.nNo, dtype=np.int32)
    for q in parts:
        count[q.ltg] += 1
    faces = []
    for (g, nodes, val) in gfaces:
        loc = np.searchsorted(p.ltg, nodes)
        ok = (loc < p.nNo) & (p.ltg[np.minimum(loc, p.nNo - 1)] == nodes)
        # nodal normal integrals of a shared face node are PARTIAL on each rank (fsils_bc_create sums them, bc.cpp:60-90)
        faces.append((g, loc[ok].astype(np.int32), np.asfortranarray(val[:, ok] / count[nodes[ok]])))
    c = refbind.RefCase()
    c.set_coords(m.x[:, p.ltg])
    c.set_partition(m.nNo, p.ltg)
    c.add_mesh(p.IEN)
    c.build_graph(len(faces))
    mynNo, lmap, reqs = c.get_lhs()
    # ---- the product's host logic against fsils_lhs_create ------------------------------------------------------
    same_map = bool(np.array_equal(lmap, p.node_map)) and mynNo == p.mynNo
    same_nb = [q for q, _ in reqs] == [q for q, _ in p.neighbours]
    same_ptr = same_nb and all(np.array_equal(a[1], b[1]) for a, b in zip(reqs, p.neighbours))
    same_sets = same_nb and all(np.array_equal(np.sort(a[1]), np.sort(b[1])) for a, b in zip(reqs, p.neighbours))
    for i, (g, nodes, val) in enumerate(faces):
        c.set_face(i, g, nodes, val)
    eq, dmn = abi.fluid_eq(0.005), [abi.fluid_domain()]
    c.alloc(4)
    c.set_state(Ag[:, p.ltg], Yg[:, p.ltg], Dg[:, p.ltg], Bf[:, p.ltg])
    c.assemble(0, eq, dmn)
    c.commu_R()
    R = c.get_R()
    ls_type, ls = common.mrank_ls(ls_name)
    X, o, _ = c.solve(4, ls_type, ls, np.ones(len(faces), np.int32), res)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), ltg=p.ltg, R=R, X=X, itr=o.RI.itr, iNorm=o.RI.iNorm, fNorm=o.RI.fNorm,
             success=o.RI.success, same_map=same_map, same_nb=same_nb, same_ptr=same_ptr, same_sets=same_sets, mynNo=mynNo)


if __name__ == "__main__":
    main()
