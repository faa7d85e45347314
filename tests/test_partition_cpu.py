"""CPU tests of the multi-GPU host logic: element partition -> local meshes, FSILS node order, shared-node
lists (svmultiphysics_b200/partition.py), incl. a world_size-2 gloo run of the halo-sum exchange pattern."""
import os
import sys

import numpy as np
import pytest

from svmultiphysics_b200 import meshgen, partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _parts(nranks, n=4, nz=6, mode="slab"):
    m = meshgen.cylinder_tet4(n, nz)
    if mode == "slab":
        hexid = np.arange(m.nEl) // 6
        k = hexid // (n * n)
        part = (k * nranks // nz).astype(np.int32)
    elif mode == "metis":
        from oracle import metis_part
        if not metis_part.have_metis():
            pytest.skip("needs oracle/_ref/libsvmetis.so (make -C oracle metis)")
        part, cut = metis_part.part_mesh_dual(m.IEN, m.nNo, nranks)
        counts = np.bincount(part, minlength=nranks)
        assert cut > 0 and counts.min() > 0 and counts.max() <= 1.10 * m.nEl / nranks     # balanced k-way partition
    else:
        part = np.random.default_rng(7).integers(0, nranks, m.nEl).astype(np.int32)
    return m, part, partition.partition_mesh(m.IEN, m.nNo, part, nranks)


@pytest.mark.parametrize("nranks,mode", [(2, "slab"), (3, "slab"), (4, "random"), (2, "metis"), (8, "metis")])
def test_partition_invariants(nranks, mode):
    m, part, parts = _parts(nranks, mode=mode)
    owned = np.zeros(m.nNo, dtype=np.int32)
    for p in parts:
        assert sorted(p.node_map.tolist()) == list(range(p.nNo))           # a permutation
        inv = np.empty(p.nNo, dtype=np.int64); inv[p.node_map] = np.arange(p.nNo)
        owned[p.ltg[inv[:p.mynNo]]] += 1                                    # FSILS positions [0,mynNo) are "owned"
        assert np.array_equal(p.ltg[p.IEN], m.IEN[:, p.elems])              # local connectivity maps back
    assert np.all(owned == 1), "every global node must be owned by exactly one rank (dots count it once)"
    # neighbour lists: same global nodes in the same order on both sides
    for p in parts:
        inv = np.empty(p.nNo, dtype=np.int64); inv[p.node_map] = np.arange(p.nNo)
        for (q, ptr) in p.neighbours:
            other = parts[q]
            inv_o = np.empty(other.nNo, dtype=np.int64); inv_o[other.node_map] = np.arange(other.nNo)
            ptr_o = dict(other.neighbours)[p.rank]
            assert np.array_equal(p.ltg[inv[ptr]], other.ltg[inv_o[ptr_o]])


def test_slab_generator_matches_general_partition():
    n, nz, nranks = 3, 2, 3
    for r in range(nranks):
        m, other, plo, phi = meshgen.cylinder_slab(n, nz, r, nranks)
        node_map, mynNo = partition.fsils_order(other, r)
        nshared_hi = len(phi) if r < nranks - 1 else 0
        assert mynNo == m.nNo - nshared_hi
        if r > 0:
            assert np.array_equal(np.sort(node_map[plo]), np.arange(len(plo)))      # low group first
        if r < nranks - 1:
            assert node_map[phi].min() == mynNo                                       # high group last


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from svmultiphysics_b200 import meshgen, partition
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n, nz = 3, 4
m = meshgen.cylinder_tet4(n, nz)
if len(sys.argv) > 2 and sys.argv[2] == "metis":
    from oracle import metis_part
    part, _ = metis_part.part_mesh_dual(m.IEN, m.nNo, world)
else:
    part = ((np.arange(m.nEl) // 6) // (n * n) * world // nz).astype(np.int32)
parts = partition.partition_mesh(m.IEN, m.nNo, part, world)
p = parts[rank]
# a nodal field that is a partial sum on every rank: value = number of local elements touching the node
inv = np.empty(p.nNo, dtype=np.int64); inv[p.node_map] = np.arange(p.nNo)
v = np.zeros(p.nNo)                         # FSILS order
np.add.at(v, p.node_map[p.IEN.ravel()], 1.0)
# halo sum exactly as fsils_commuv: send my partial values of the shared nodes, add what I receive
reqs, recv = [], {}
for (q, ptr) in p.neighbours:
    recv[q] = torch.zeros(len(ptr), dtype=torch.float64)
    reqs.append(dist.isend(torch.from_numpy(v[ptr].copy()), q))
    reqs.append(dist.irecv(recv[q], q))
for r in reqs: r.wait()
for (q, ptr) in p.neighbours:
    v[ptr] += recv[q].numpy()
# compare with the global count
g = np.zeros(m.nNo); np.add.at(g, m.IEN.ravel(), 1.0)
ok = np.array_equal(v[p.node_map], g[p.ltg])
# owned-node dot: sum over ranks of sum_{a<mynNo} v_a must equal the global sum
loc = torch.tensor([v[:p.mynNo].sum()], dtype=torch.float64)
dist.all_reduce(loc)
ok = ok and abs(loc.item() - g.sum()) < 1e-9
flag = torch.tensor([1 if ok else 0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
'''


@pytest.mark.parametrize("world,mode", [(2, "slab"), (3, "metis")])
def test_halo_sum_pattern_gloo_world2(tmp_path, world, mode):
    import subprocess
    if mode == "metis":
        from oracle import metis_part
        if not metis_part.have_metis():
            pytest.skip("needs oracle/_ref/libsvmetis.so (make -C oracle metis)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if world == 2 else "29534", str(script), ROOT, mode]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.parametrize("n,nzg,blocks", [(4, 6, (1, 1, 3)), (4, 4, (2, 1, 1)), (4, 6, (2, 2, 1)), (4, 4, (2, 2, 2)), (6, 5, (3, 2, 2))])
def test_lattice_blocks_equal_general_partition(n, nzg, blocks):
    """The index-arithmetic block partitioner bench.py uses at 10 M elements per GPU produces exactly what partition_mesh
    (itself identical to the reference's fsils_lhs_create, test_multirank_reference_cpu.py) gives for the block part[] array:
    local meshes, lhs.map, mynNo, neighbour ranks and the order of every shared-node list."""
    m = meshgen.cylinder_tet4(n, nzg)
    lb = partition.LatticeBlocks((n, n, nzg), blocks)
    part = np.repeat(lb.part_array(), 6)
    parts = partition.partition_mesh(m.IEN, m.nNo, part, lb.nranks)
    maxshare = 0
    for r, p in enumerate(parts):
        mb = meshgen.cylinder_block(n, nzg, blocks, r)
        gid = mb.gijk[0] + (n + 1) * (mb.gijk[1] + (n + 1) * mb.gijk[2])
        assert np.array_equal(gid, p.ltg)
        assert np.array_equal(mb.IEN, p.IEN)
        assert np.array_equal(mb.x, m.x[:, p.ltg])
        node_map, mynNo = lb.order(r)
        assert mynNo == p.mynNo and np.array_equal(node_map, p.node_map)
        nb = lb.neighbours(r)
        assert [q for q, _ in nb] == [q for q, _ in p.neighbours]
        for (q, ptr), (_, ptr0) in zip(nb, p.neighbours):
            assert np.array_equal(ptr, ptr0)
        for name in ("wall", "inlet", "outlet"):
            assert np.array_equal(np.sort(p.ltg[mb.faces[name]]), np.intersect1d(m.faces[name], p.ltg))
        maxshare = max(maxshare, int(lb.multiplicity(r).max()))
    assert maxshare == min(lb.nranks, 2 ** sum(b > 1 for b in blocks))


def test_default_blocks():
    assert meshgen.default_blocks(8) == (2, 2, 2) and meshgen.default_blocks(4) == (2, 2, 1) and meshgen.default_blocks(2) == (2, 1, 1)
    assert meshgen.default_blocks(8, "slab") == (1, 1, 8) and meshgen.default_blocks(6) == (1, 1, 6) and meshgen.default_blocks(1) == (1, 1, 1)
