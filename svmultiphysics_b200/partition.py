"""Element-partition -> per-rank local meshes and FSILS communication structure (host logic).

Mirrors what the reference does once at start-up:

* ``part_msh`` (Code/Source/solver/distribute.cpp:1972) splits the ELEMENTS among ranks (ParMETIS there;
  any ``part[e]`` array here) and duplicates interface nodes on every rank that touches them;
* ``fsils_lhs_create`` (Code/Source/linear_solver/lhs.cpp:30-348) reorders the local nodes of a rank as
  ``[shared only with lower ranks | interior | shared with a higher rank]``, sets ``mynNo`` so that the
  last group is excluded (a node is "owned" by the highest rank holding it, lhs.cpp:139-160), and
  builds per-neighbour lists of shared nodes in the same order on both sides (the order of the
  higher rank's FSILS numbering, lhs.cpp:281-347).

Everything here is integer bookkeeping on the host (numpy); the numbers it produces are fed to
``svb200_set_graph``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class LocalPart:
    rank: int
    ltg: np.ndarray                 # (nNo,) local (input order) -> global node id
    IEN: np.ndarray                 # (eNoN, nEl) local node ids (input order)
    elems: np.ndarray               # global element ids of this rank
    node_map: np.ndarray            # (nNo,) input order -> FSILS order (lhs.map)
    mynNo: int
    neighbours: list = field(default_factory=list)   # [(rank, ptr)] ptr = FSILS-order local ids

    @property
    def nNo(self):
        return len(self.ltg)


def fsils_order(max_other_rank: np.ndarray, rank: int):
    """lhs.map and mynNo for one rank.  ``max_other_rank[a]`` = highest OTHER rank that also holds
    local node a, or -1 when the node is interior to this rank."""
    n = len(max_other_rank)
    group = np.where(max_other_rank < 0, 1, np.where(max_other_rank > rank, 2, 0)).astype(np.int8)
    order = np.argsort(group, kind="stable")     # FSILS position -> input id
    node_map = np.empty(n, dtype=np.int32)
    node_map[order] = np.arange(n, dtype=np.int32)
    mynNo = int(np.count_nonzero(group < 2))
    return node_map, mynNo


def fsils_order_exact(node_sets, rank: int):
    """lhs.map and mynNo for one rank in EXACTLY the order fsils_lhs_create produces (linear_solver/lhs.cpp:118-175): the
    other ranks are visited from the highest to the lowest, each one's nodes in that rank's local order (ascending global
    id here); a node this rank also holds and that is not placed yet goes, for a LOWER rank, to the next free position at
    the front, for a HIGHER rank to the last free position at the back (so the back group is in reverse order of encounter);
    the interior nodes fill the middle in local order.  ``node_sets[r]`` = sorted global ids of rank r's nodes."""
    mine = node_sets[rank]
    n = len(mine)
    placed = np.zeros(n, dtype=bool)
    low, high = [], []
    for i in range(len(node_sets) - 1, -1, -1):
        if i == rank:
            continue
        common = np.intersect1d(node_sets[i], mine, assume_unique=True)      # ascending global id = rank i's local order
        loc = np.searchsorted(mine, common)
        sel = loc[~placed[loc]]
        (low if i < rank else high).append(sel)
        placed[sel] = True
    low = np.concatenate(low) if low else np.zeros(0, dtype=np.int64)
    high = np.concatenate(high) if high else np.zeros(0, dtype=np.int64)
    interior = np.flatnonzero(~placed)
    order = np.concatenate([low, interior, high[::-1]]).astype(np.int64)          # FSILS position -> local id
    node_map = np.empty(n, dtype=np.int32)
    node_map[order] = np.arange(n, dtype=np.int32)
    return node_map, int(len(low) + len(interior))


def partition_mesh(IEN: np.ndarray, nNo: int, part: np.ndarray, nranks: int):
    """Split a global mesh by the element partition ``part`` (values in [0,nranks))."""
    node_sets = []
    for r in range(nranks):
        el = np.flatnonzero(part == r)
        node_sets.append(np.unique(IEN[:, el]))
    # highest and second-highest rank holding each global node
    max1 = -np.ones(nNo, dtype=np.int32)
    max2 = -np.ones(nNo, dtype=np.int32)
    for r, nodes in enumerate(node_sets):          # ascending r: the new rank is always the largest so far
        max2[nodes] = max1[nodes]
        max1[nodes] = r
    parts = []
    for r in range(nranks):
        nodes = node_sets[r]
        gtl = -np.ones(nNo, dtype=np.int64)
        gtl[nodes] = np.arange(len(nodes))
        el = np.flatnonzero(part == r).astype(np.int64)
        lIEN = np.asfortranarray(gtl[IEN[:, el]].astype(np.int32))
        other = np.where(max1[nodes] != r, max1[nodes], max2[nodes])
        # a node held by lower ranks only besides r: max1 == r and max2 = highest lower rank (or -1)
        node_map, mynNo = fsils_order_exact(node_sets, r)
        assert mynNo == fsils_order(other, r)[1]          # same three groups as the closed-form rule above
        parts.append(LocalPart(rank=r, ltg=nodes.astype(np.int64), IEN=lIEN, elems=el, node_map=node_map, mynNo=mynNo))
    # neighbour lists: shared nodes of (lo,hi) in the order of hi's FSILS numbering
    for hi in range(nranks):
        for lo in range(hi):
            common = np.intersect1d(node_sets[lo], node_sets[hi], assume_unique=True)
            if len(common) == 0:
                continue
            ph, pl = parts[hi], parts[lo]
            fs_hi = ph.node_map[np.searchsorted(ph.ltg, common)]
            o = np.argsort(fs_hi, kind="stable")
            common = common[o]
            ptr_hi = fs_hi[o].astype(np.int32)
            ptr_lo = pl.node_map[np.searchsorted(pl.ltg, common)].astype(np.int32)
            ph.neighbours.append((lo, ptr_hi))
            pl.neighbours.append((hi, ptr_lo))
    for p in parts:
        p.neighbours.sort(key=lambda t: t[0])
    return parts


def glue_nodal(parts, arrays, nNo_global):
    """Assemble a global (rows, nNo) array from per-rank local arrays (input order); shared nodes
    must agree across ranks after a halo sum, the highest rank's copy is kept."""
    rows = arrays[0].shape[0]
    out = np.zeros((rows, nNo_global), order="F")
    for p, a in zip(parts, arrays):
        out[:, p.ltg] = a
    return out


class LatticeBlocks:
    """The element partition of a structured (nx, ny, nz)-cell lattice into (bx, by, bz) blocks, rank = ri + bx (rj + by rk), with
    the FSILS node order and shared-node lists of every rank computed from index arithmetic instead of from global node sets:
    the same numbers ``partition_mesh`` produces for the block ``part[]`` array (and hence the reference's ``fsils_lhs_create``,
    linear_solver/lhs.cpp:30-348; tests/test_partition_cpu.py checks the identity), without ever holding the global mesh —
    bench.py builds one 10 M-element block per GPU.  A rank's local node order is its local lattice order (i fastest), which is
    the ascending-global-id order ``partition_mesh`` uses."""

    def __init__(self, ncells, blocks):
        self.nc = tuple(int(c) for c in ncells)
        self.blocks = tuple(int(b) for b in blocks)
        self.nranks = self.blocks[0] * self.blocks[1] * self.blocks[2]
        self._order = {}

    def cell_ranges(self, r):
        bx, by, _ = self.blocks
        r3 = (r % bx, (r // bx) % by, r // (bx * by))
        return [((self.nc[d] * r3[d]) // self.blocks[d], (self.nc[d] * (r3[d] + 1)) // self.blocks[d]) for d in range(3)]

    def nNo(self, r):
        return int(np.prod([hi - lo + 1 for lo, hi in self.cell_ranges(r)]))

    def _gids_box(self, box):
        """Ascending global node ids of the closed index box [(lo, hi)] x 3."""
        n1x, n1y = self.nc[0] + 1, self.nc[1] + 1
        i = np.arange(box[0][0], box[0][1] + 1, dtype=np.int64)
        j = np.arange(box[1][0], box[1][1] + 1, dtype=np.int64)
        k = np.arange(box[2][0], box[2][1] + 1, dtype=np.int64)
        return (i[None, None, :] + n1x * (j[None, :, None] + n1y * k[:, None, None])).reshape(-1)

    def common(self, r, s):
        """Global ids (ascending) of the nodes ranks r and s both hold."""
        a, b = self.cell_ranges(r), self.cell_ranges(s)
        box = [(max(a[d][0], b[d][0]), min(a[d][1], b[d][1])) for d in range(3)]
        if any(lo > hi for lo, hi in box):
            return np.zeros(0, dtype=np.int64)
        return self._gids_box(box)

    def local_of(self, r, gids):
        """Local (input-order) node ids on rank r of global node ids that lie in its box."""
        n1x, n1y = self.nc[0] + 1, self.nc[1] + 1
        (i0, i1), (j0, j1), (k0, _) = self.cell_ranges(r)
        i, j, k = gids % n1x, (gids // n1x) % n1y, gids // (n1x * n1y)
        return (i - i0) + (i1 - i0 + 1) * ((j - j0) + (j1 - j0 + 1) * (k - k0))

    def order(self, r):
        """(lhs.map, mynNo) of rank r, visit by visit like lhs.cpp:118-175 (see fsils_order_exact)."""
        if r in self._order:
            return self._order[r]
        n = self.nNo(r)
        placed = np.zeros(n, dtype=bool)
        low, high = [], []
        for s in range(self.nranks - 1, -1, -1):
            if s == r:
                continue
            c = self.common(r, s)
            if len(c) == 0:
                continue
            loc = self.local_of(r, c)
            sel = loc[~placed[loc]]
            (low if s < r else high).append(sel)
            placed[sel] = True
        low = np.concatenate(low) if low else np.zeros(0, dtype=np.int64)
        high = np.concatenate(high) if high else np.zeros(0, dtype=np.int64)
        interior = np.flatnonzero(~placed)
        order = np.concatenate([low, interior, high[::-1]]).astype(np.int64)
        node_map = np.empty(n, dtype=np.int32)
        node_map[order] = np.arange(n, dtype=np.int32)
        self._order[r] = (node_map, int(len(low) + len(interior)))
        return self._order[r]

    def neighbours(self, r):
        """[(rank, ptr)] ascending in rank; ptr = FSILS-order local ids in the order of the HIGHER rank's numbering (lhs.cpp:281-347)."""
        out = []
        for s in range(self.nranks):
            if s == r:
                continue
            c = self.common(r, s)
            if len(c) == 0:
                continue
            hi, lo = max(r, s), min(r, s)
            fs_hi = self.order(hi)[0][self.local_of(hi, c)]
            o = np.argsort(fs_hi, kind="stable")
            if r == hi:
                ptr = fs_hi[o].astype(np.int32)
            else:
                ptr = self.order(lo)[0][self.local_of(lo, c[o])].astype(np.int32)
            out.append((s, ptr))
        return out

    def multiplicity(self, r):
        """Number of ranks holding each local node of rank r (1 = interior)."""
        cnt = np.ones(self.nNo(r), dtype=np.int32)
        for s in range(self.nranks):
            if s != r:
                c = self.common(r, s)
                if len(c):
                    cnt[self.local_of(r, c)] += 1
        return cnt

    def part_array(self):
        """part[e] of the GLOBAL tet mesh (6 tets per cell, cells i fastest) — small lattices / tests only."""
        nx, ny, nz = self.nc
        ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        cid = (ci + nx * (cj + ny * ck)).ravel()
        o = np.argsort(cid, kind="stable")
        ci, cj, ck = ci.ravel()[o], cj.ravel()[o], ck.ravel()[o]
        part = np.zeros(len(ci), dtype=np.int32)
        for r in range(self.nranks):
            (i0, i1), (j0, j1), (k0, k1) = self.cell_ranges(r)
            part[(ci >= i0) & (ci < i1) & (cj >= j0) & (cj < j1) & (ck >= k0) & (ck < k1)] = r
        return part
