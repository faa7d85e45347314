"""Deterministic synthetic meshes for the hot-path configs.

Nothing here reads external files: the reference's own test meshes are Git-LFS pointers that are
not available offline, so every config is generated from a structured hex lattice.

* ``cylinder_tet4``  — the ``pipe_RCR_3d`` analogue (C1) and the 10 M / 80 M tet4 cylinders (C2/C3):
  a (n x n) square lattice mapped smoothly onto a disc, extruded along z, every hex Kuhn-split into
  6 tet4 sharing the (0,0,0)-(1,1,1) diagonal so that face diagonals match between neighbours.
* ``box_hex8``       — the ``block_compression`` analogue (C4).

Node order per tet obeys the solver's convention det[x0-x3, x1-x3, x2-x3] > 0
(tet4 shape functions are "origin-last", Code/Source/solver/nn.cpp:174; the Jacobian is the
determinant computed in nn::gnn, Code/Source/solver/nn.cpp:871).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    """Column-major (Fortran-order) arrays exactly as the reference holds them."""
    x: np.ndarray            # (3, nNo) float64
    IEN: np.ndarray          # (eNoN, nEl) int32
    eNoN: int
    faces: dict = field(default_factory=dict)   # name -> int32 node ids
    eId: np.ndarray | None = None               # (nEl,) int32 domain bitmask
    lattice: tuple | None = None                # (nx, ny, nz) cells
    gijk: tuple | None = None                   # lattice blocks: GLOBAL lattice indices (i, j, k) of the local nodes
    glattice: tuple | None = None               # lattice blocks: cells of the GLOBAL lattice
    origin: tuple | None = None                 # lattice blocks: first global cell (i0, j0, k0) of this block

    @property
    def nNo(self) -> int:
        return self.x.shape[1]

    @property
    def nEl(self) -> int:
        return self.IEN.shape[1]


def _lattice_nodes(nx, ny, nz):
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    # node id = i + (nx+1)*(j + (ny+1)*k): i fastest
    order = np.argsort((i + (nx + 1) * (j + (ny + 1) * k)).ravel(), kind="stable")
    return i.ravel()[order], j.ravel()[order], k.ravel()[order]


def _hex_cells(nx, ny, nz):
    """(8, nCells) node ids in VTK hexahedron order, cells ordered i fastest."""
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cid = (ci + nx * (cj + ny * ck)).ravel()
    order = np.argsort(cid, kind="stable")
    ci, cj, ck = ci.ravel()[order], cj.ravel()[order], ck.ravel()[order]

    def nid(di, dj, dk):
        return (ci + di) + (nx + 1) * ((cj + dj) + (ny + 1) * (ck + dk))

    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    return np.stack([nid(*c) for c in corners]).astype(np.int64)


def _kuhn_tets(hexes):
    """Split VTK-ordered hexes (8,n) into 6 tets each: (4, 6n), tets of one hex adjacent."""
    corner_of = {(0, 0, 0): 0, (1, 0, 0): 1, (1, 1, 0): 2, (0, 1, 0): 3,
                 (0, 0, 1): 4, (1, 0, 1): 5, (1, 1, 1): 6, (0, 1, 1): 7}
    local = []
    for perm in itertools.permutations(range(3)):
        v = [0, 0, 0]
        tet = [corner_of[tuple(v)]]
        for ax in perm:
            v[ax] = 1
            tet.append(corner_of[tuple(v)])
        local.append(tet)
    local = np.array(local)                       # (6,4)
    tets = hexes[local.T.reshape(4, 6), :]        # (4,6,n)
    n = hexes.shape[1]
    return tets.transpose(0, 2, 1).reshape(4, 6 * n)


def _fix_orientation(x, IEN):
    p = x[:, IEN]                                 # (3,4,nEl)
    d = p[:, :3, :] - p[:, 3:4, :]
    det = np.einsum("ie,ie->e", d[:, 0, :], np.cross(d[:, 1, :], d[:, 2, :], axis=0))
    neg = det < 0
    IEN = IEN.copy()
    IEN[0, neg], IEN[1, neg] = IEN[1, neg].copy(), IEN[0, neg].copy()
    return IEN


def cylinder_tet4(n: int, nz: int, R: float = 2.0, L: float = 30.0) -> Mesh:
    """Cylinder of radius R, length L along z: (n x n x nz) hexes -> 6*n*n*nz tet4."""
    i, j, k = _lattice_nodes(n, n, nz)
    u = 2.0 * i / n - 1.0
    v = 2.0 * j / n - 1.0
    # elliptical square -> disc map (smooth, bijective, boundary -> circle)
    xs = R * u * np.sqrt(1.0 - 0.5 * v * v)
    ys = R * v * np.sqrt(1.0 - 0.5 * u * u)
    zs = L * k / nz
    x = np.asfortranarray(np.stack([xs, ys, zs]).astype(np.float64))
    IEN = _fix_orientation(x, _kuhn_tets(_hex_cells(n, n, nz)))
    ids = np.arange(x.shape[1], dtype=np.int32)
    wall = (i == 0) | (i == n) | (j == 0) | (j == n)
    faces = {
        "wall": ids[wall],
        "inlet": ids[(k == 0) & ~wall],
        "outlet": ids[(k == nz) & ~wall],
        "outlet_all": ids[k == nz],
    }
    return Mesh(x=x, IEN=np.asfortranarray(IEN.astype(np.int32)), eNoN=4, faces=faces, lattice=(n, n, nz))


def box_tet4(nx: int, ny: int, nz: int, lengths=(1.0, 1.0, 1.0)) -> Mesh:
    i, j, k = _lattice_nodes(nx, ny, nz)
    x = np.asfortranarray(np.stack([lengths[0] * i / nx, lengths[1] * j / ny, lengths[2] * k / nz]).astype(np.float64))
    IEN = _fix_orientation(x, _kuhn_tets(_hex_cells(nx, ny, nz)))
    ids = np.arange(x.shape[1], dtype=np.int32)
    faces = {"X0": ids[i == 0], "X1": ids[i == nx], "Y0": ids[j == 0], "Y1": ids[j == ny],
             "Z0": ids[k == 0], "Z1": ids[k == nz]}
    return Mesh(x=x, IEN=np.asfortranarray(IEN.astype(np.int32)), eNoN=4, faces=faces, lattice=(nx, ny, nz))


def box_hex8(nx: int, ny: int, nz: int, lengths=(1.0, 1.0, 1.0)) -> Mesh:
    i, j, k = _lattice_nodes(nx, ny, nz)
    x = np.asfortranarray(np.stack([lengths[0] * i / nx, lengths[1] * j / ny, lengths[2] * k / nz]).astype(np.float64))
    IEN = _hex_cells(nx, ny, nz)
    ids = np.arange(x.shape[1], dtype=np.int32)
    faces = {"X0": ids[i == 0], "X1": ids[i == nx], "Y0": ids[j == 0], "Y1": ids[j == ny],
             "Z0": ids[k == 0], "Z1": ids[k == nz]}
    return Mesh(x=x, IEN=np.asfortranarray(IEN.astype(np.int32)), eNoN=8, faces=faces, lattice=(nx, ny, nz))


def poiseuille_state(mesh: Mesh, R: float = 2.0, U: float = 10.0, seed: int = 1234, tDof: int = 4,
                     noise: float = 0.01, dpdz: float = -1.0):
    """Seeded synthetic state of SURVEY §8(d) C2: Poiseuille u_z + 1 % noise, linear p, Ag ~ N(0,1)e-2."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = mesh.nNo
    r2 = (mesh.x[0] ** 2 + mesh.x[1] ** 2) / (R * R)
    Yg = np.zeros((tDof, n), order="F")
    Yg[2] = U * (1.0 - r2)
    Yg[:3] += noise * U * rng.uniform(-1.0, 1.0, size=(3, n))
    Yg[3] = dpdz * mesh.x[2] + noise * rng.uniform(-1.0, 1.0, size=n)
    rng2 = np.random.Generator(np.random.PCG64(seed + 1))
    Ag = np.asfortranarray(1e-2 * rng2.standard_normal((tDof, n)))
    Dg = np.zeros((tDof, n), order="F")
    return np.asfortranarray(Ag), np.asfortranarray(Yg), Dg


def cylinder_slab(n: int, nz: int, rank: int, nranks: int, R: float = 2.0, Lseg: float = 30.0):
    """Weak-scaling partition used by bench.py: rank r owns the z-slab [r*Lseg,(r+1)*Lseg] of a cylinder of
    nranks*nz cell layers (one 6*n*n*nz-tet slab per GPU).  Returns (mesh, max_other_rank, plane_lo,
    plane_hi): the local mesh in local node ids, for every local node the highest other rank holding
    it (-1 = interior), and the local ids of the two interface planes in (i,j) order.  This geometric
    slab partition stands in for the reference's ParMETIS element partition (Code/Source/solver/SPLIT.c),
    which needs MPI; any other partition can be fed through partition.partition_mesh."""
    m = cylinder_tet4(n, nz, R=R, L=Lseg)
    m.x[2] += rank * Lseg
    P = (n + 1) * (n + 1)
    plane_lo = np.arange(P, dtype=np.int32)                      # k = 0
    plane_hi = (np.arange(P) + P * nz).astype(np.int32)          # k = nz
    other = -np.ones(m.nNo, dtype=np.int32)
    if rank > 0:
        other[plane_lo] = rank - 1
    if rank < nranks - 1:
        other[plane_hi] = rank + 1
    return m, other, plane_lo, plane_hi


def block_ranges(ncells, blocks, rank):
    """Cell ranges [(lo, hi)] x 3 of block `rank` of a (bx, by, bz) split of an (nx, ny, nz)-cell lattice; rank = ri + bx (rj + by rk)."""
    bx, by, bz = blocks
    r3 = (rank % bx, (rank // bx) % by, rank // (bx * by))
    return [((ncells[d] * r3[d]) // blocks[d], (ncells[d] * (r3[d] + 1)) // blocks[d]) for d in range(3)]


def cylinder_box(n: int, nzg: int, ranges, R: float = 2.0, L: float = 30.0) -> Mesh:
    """The cells [i0,i1) x [j0,j1) x [k0,k1) (``ranges``) of the (n x n x nzg)-hex cylinder lattice of ``cylinder_tet4(n, nzg, R, L)``
    as a mesh of its own in local node ids (local lattice order, i fastest = ascending global node id); elements keep the global
    order.  faces / gijk refer to the GLOBAL lattice (wall = lateral surface, inlet k = 0, outlet k = nzg)."""
    (i0, i1), (j0, j1), (k0, k1) = ranges
    nx, ny, nz = i1 - i0, j1 - j0, k1 - k0
    i, j, k = _lattice_nodes(nx, ny, nz)
    gi, gj, gk = i + i0, j + j0, k + k0
    u = 2.0 * gi / n - 1.0
    v = 2.0 * gj / n - 1.0
    xs = R * u * np.sqrt(1.0 - 0.5 * v * v)
    ys = R * v * np.sqrt(1.0 - 0.5 * u * u)
    zs = L * gk / nzg
    x = np.asfortranarray(np.stack([xs, ys, zs]).astype(np.float64))
    IEN = _fix_orientation(x, _kuhn_tets(_hex_cells(nx, ny, nz)))
    ids = np.arange(x.shape[1], dtype=np.int32)
    wall = (gi == 0) | (gi == n) | (gj == 0) | (gj == n)
    faces = {"wall": ids[wall], "inlet": ids[(gk == 0) & ~wall], "outlet": ids[(gk == nzg) & ~wall], "outlet_all": ids[gk == nzg]}
    return Mesh(x=x, IEN=np.asfortranarray(IEN.astype(np.int32)), eNoN=4, faces=faces, lattice=(nx, ny, nz),
                gijk=(gi.astype(np.int64), gj.astype(np.int64), gk.astype(np.int64)), glattice=(n, n, nzg), origin=(i0, j0, k0))


def cylinder_block(n: int, nzg: int, blocks, rank: int, R: float = 2.0, L: float = 30.0) -> Mesh:
    """Block `rank` of a (bx, by, bz) split of the cylinder lattice as a local mesh.  Nodes on block interfaces are
    duplicated on every block that touches them, exactly what an ELEMENT partition does (Code/Source/solver/distribute.cpp:1972);
    the local mesh equals ``partition.partition_mesh`` applied to the global mesh with the block ``part[]`` array
    (tests/test_partition_cpu.py).  ``blocks = (1, 1, N)`` is the z-slab partition; (2, 2, 2) gives every rank 7 neighbours and
    nodes shared by up to 8 ranks."""
    return cylinder_box(n, nzg, block_ranges((n, n, nzg), blocks, rank), R=R, L=L)


def default_blocks(nranks: int, mode: str = "blocks"):
    """(bx, by, bz) for `nranks` partitions: "slab" = (1, 1, N); "blocks" = split x, then y, then z by powers of two
    (2 -> (2,1,1), 4 -> (2,2,1), 8 -> (2,2,2)); other counts fall back to slabs."""
    if mode == "slab" or nranks & (nranks - 1):
        return (1, 1, nranks)
    b = [1, 1, 1]
    d = 0
    r = nranks
    while r > 1:
        b[d % 3] *= 2
        r //= 2
        d += 1
    return tuple(b)


_TET_FACES = ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))
_HEX_FACES = ((0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7))   # VTK hexahedron


def boundary_face_elements(mesh: Mesh, nodes):
    """Face elements (faceType.IEN / gE of the reference) of the boundary patch spanned by `nodes`: every element face
    whose nodes all lie in the set and that belongs to exactly one element.  Returns IENb(eNoNb, nElb), gE(nElb)."""
    inset = np.zeros(mesh.nNo, bool)
    inset[np.asarray(nodes)] = True
    loc = _TET_FACES if mesh.eNoN == 4 else _HEX_FACES
    cand_f, cand_e = [], []
    for f in loc:
        fn = mesh.IEN[list(f), :]                     # (eNoNb, nEl)
        sel = np.nonzero(inset[fn].all(axis=0))[0]
        cand_f.append(fn[:, sel]); cand_e.append(sel)
    F = np.concatenate(cand_f, axis=1)
    E = np.concatenate(cand_e)
    key = np.sort(F, axis=0)
    _, inv, cnt = np.unique(key, axis=1, return_inverse=True, return_counts=True)
    keep = cnt[inv.ravel()] == 1
    order = np.argsort(E[keep], kind="stable")
    return np.asfortranarray(F[:, keep][:, order].astype(np.int32)), E[keep][order].astype(np.int32)


# ---- quadratic and wedge elements (nn_elem_props.h: TET10, HEX20, HEX27, WDG; VTK node order, checked against the reference's
# shape-function tables in tests/test_quadratic_elements_cpu.py) ---------------------------------------------------------------
_TET_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
_HEX_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
_HEX_FACES = [(0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (3, 2, 6, 7), (0, 1, 2, 3), (4, 5, 6, 7)]


def elevate(m: Mesh, kind: str, bend: float = 0.0) -> Mesh:
    """TET4 -> "tet10", HEX8 -> "hex20" / "hex27": one new node per edge (and per face and cell for hex27), shared between the
    elements, placed at the mean of its corners; `bend` then applies a smooth non-linear map to ALL nodes so that the elements are
    curved (non-constant Jacobian, non-zero second derivatives of the geometry)."""
    if kind == "tet10":
        assert m.eNoN == 4
        groups = _TET_EDGES
    elif kind == "hex20":
        assert m.eNoN == 8
        groups = _HEX_EDGES
    elif kind == "hex27":
        assert m.eNoN == 8
        groups = _HEX_EDGES + _HEX_FACES + [tuple(range(8))]
    else:
        raise ValueError(kind)
    ids = {}
    pts = [m.x[:, a] for a in range(m.nNo)]
    IEN = np.zeros((m.eNoN + len(groups), m.nEl), dtype=np.int32, order="F")
    IEN[:m.eNoN] = m.IEN
    for e in range(m.nEl):
        for k, grp in enumerate(groups):
            key = tuple(sorted(int(m.IEN[c, e]) for c in grp))
            if key not in ids:
                ids[key] = len(pts)
                pts.append(m.x[:, list(key)].mean(axis=1))
            IEN[m.eNoN + k, e] = ids[key]
    x = np.asfortranarray(np.stack(pts, axis=1))
    if bend:
        L = np.abs(x).max()
        s = x / L
        x = np.asfortranarray(x + bend * L * np.stack([np.sin(2.1 * s[1]) * s[2], s[0] * s[0] - 0.5 * s[2], np.cos(1.7 * s[0]) * s[1]]))
    return Mesh(x=x, IEN=IEN, eNoN=IEN.shape[0], faces={}, lattice=m.lattice)


def box_wdg6(nx: int, ny: int, nz: int, lengths=(1.0, 1.0, 1.0), bend: float = 0.0) -> Mesh:
    """Every hex of the box split into two 6-node wedges (triangles in the x-y plane, "origin-last" like the solver's TRI3 / TET4)."""
    h = box_hex8(nx, ny, nz, lengths)
    H = h.IEN
    IEN = np.concatenate([np.stack([H[1], H[3], H[0], H[5], H[7], H[4]]), np.stack([H[3], H[1], H[2], H[7], H[5], H[6]])], axis=1)
    x = h.x
    if bend:
        L = np.abs(x).max()
        s = x / L
        x = np.asfortranarray(x + bend * L * np.stack([np.sin(2.1 * s[1]) * s[2], s[0] * s[0] - 0.5 * s[2], np.cos(1.7 * s[0]) * s[1]]))
    return Mesh(x=x, IEN=np.asfortranarray(IEN.astype(np.int32)), eNoN=6, faces=h.faces, lattice=h.lattice)
