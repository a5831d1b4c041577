"""Structured hexahedral meshes and their box partitions: the host-side setup for the multi-GPU path.

The reference gets its partitions from METIS through mfem::ParMesh (src/serac/mesh/mesh_utils.cpp:461) and its shared-dof
bookkeeping from mfem's GroupCommunicator / ConformingProlongationOperator (mfem/fem/pfespace.cpp:3527-3663).  Neither
MPI nor METIS exists in this image, and METIS partitions are parity-unpinned anyway, so for the
synthetic structured cubes / beams of BASELINE.json configs[1] and [4] this module builds the same objects directly:

  * the global H1 order-p node grid and the element -> node table in the lexicographic order the kernels use
    (what ElementRestriction produces from mfem's native dofs, element_restriction.cpp:216-277);
  * a px x py x pz box partition, one box per rank, with local numbering;
  * ownership of shared nodes by the lowest sharing rank (mfem GroupTopology master rule, pfespace.cpp:3527-3545);
  * per-neighbour send / receive lists for P (owner -> sharers) and P^T (sharers -> owner), ordered by global node id so
    both sides agree without communicating.

Everything here runs once at setup on the host (numpy); nothing is on the evaluation path."""
import numpy as np


class StructuredHexMesh:
    """nx x ny x nz hexes on [0,lx] x [0,ly] x [0,lz]; element e = (ez*ny + ey)*nx + ex."""

    def __init__(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, origin=(0.0, 0.0, 0.0)):
        self.n = (int(nx), int(ny), int(nz))
        self.length = (float(lx), float(ly), float(lz))
        self.origin = tuple(float(o) for o in origin)
        self.ne = int(nx) * int(ny) * int(nz)
        self.dim = 3

    def node_grid(self, p):
        return tuple(p * n + 1 for n in self.n)

    def num_nodes(self, p):
        gx, gy, gz = self.node_grid(p)
        return gx * gy * gz

    def element_table(self, p, box=None, node_box=None):
        """[elements of `box`][(p+1)^3] node ids, lexicographic inside the element (x fastest).  box = ((ex0,ex1),(ey0,
        ey1),(ez0,ez1)) element ranges (default: all); node ids are numbered inside node_box = ((i0,i1),...) inclusive
        node ranges (default: the global grid)."""
        nx, ny, nz = self.n
        box = box or ((0, nx), (0, ny), (0, nz))
        gx, gy, gz = self.node_grid(p)
        node_box = node_box or ((0, gx - 1), (0, gy - 1), (0, gz - 1))
        (i0, i1), (j0, j1), (k0, k1) = node_box
        NX, NY = i1 - i0 + 1, j1 - j0 + 1
        ex = np.arange(box[0][0], box[0][1])
        ey = np.arange(box[1][0], box[1][1])
        ez = np.arange(box[2][0], box[2][1])
        EZ, EY, EX = np.meshgrid(ez, ey, ex, indexing="ij")  # element order: x fastest
        a = np.arange(p + 1)
        C, B, A = np.meshgrid(a, a, a, indexing="ij")  # local node order: x fastest
        I = (p * EX.reshape(-1, 1) + A.reshape(1, -1)) - i0
        J = (p * EY.reshape(-1, 1) + B.reshape(1, -1)) - j0
        K = (p * EZ.reshape(-1, 1) + C.reshape(1, -1)) - k0
        return ((K * NY + J) * NX + I).astype(np.uint64)

    def vertex_coordinates(self, node_box=None):
        """p = 1 mesh nodes as an L-vector ordered byVDIM ([node][3]), numbered inside node_box (inclusive ranges)."""
        gx, gy, gz = self.node_grid(1)
        node_box = node_box or ((0, gx - 1), (0, gy - 1), (0, gz - 1))
        h = [self.length[d] / self.n[d] for d in range(3)]
        ax = [self.origin[d] + h[d] * np.arange(node_box[d][0], node_box[d][1] + 1) for d in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).ravel()


def _split(n, parts):
    """element ranges of `parts` nearly equal chunks of n"""
    base, rem = divmod(n, parts)
    out, s = [], 0
    for i in range(parts):
        e = s + base + (1 if i < rem else 0)
        out.append((s, e))
        s = e
    return out


def factor_ranks(world, n=None):
    """px, py, pz with px*py*pz = world, as cubic as possible: 2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2;
    with n = (nx, ny, nz) given, longer mesh directions get the larger factors."""
    f = [1, 1, 1]
    w, d = world, 2
    primes = []
    while w > 1:
        while w % d == 0:
            primes.append(d)
            w //= d
        d += 1
    for q in sorted(primes, reverse=True):
        if n is None:
            i = int(np.argmin(f))
        else:
            i = int(np.argmax([n[k] / f[k] for k in range(3)]))
        f[i] *= q
    return tuple(f)


class Partition:
    """One rank's view of a box partition of a StructuredHexMesh for an H1 space of order p with `components` per node."""

    def __init__(self, mesh, ranks, rank, p, components=1, owned_first=False):
        """owned_first: number the nodes this rank owns first (in their lexicographic order), the nodes owned by other
        ranks after them.  The T-vector is then the prefix of the L-vector and P / P^T can run in place
        (sb_halo_*_inplace).  Default: lexicographic local numbering."""
        self.mesh, self.ranks, self.rank, self.p, self.components = mesh, tuple(ranks), int(rank), int(p), int(components)
        self.owned_first = bool(owned_first)
        self.new_of_lex = None  # lexicographic local id -> local id
        px, py, pz = self.ranks
        self.world = px * py * pz
        self.splits = [_split(mesh.n[d], self.ranks[d]) for d in range(3)]
        self.coords = self.rank_coords(self.rank)
        self.box = tuple(self.splits[d][self.coords[d]] for d in range(3))
        self.node_box = tuple((p * b[0], p * b[1]) for b in self.box)  # inclusive
        self.local_grid = tuple(nb[1] - nb[0] + 1 for nb in self.node_box)
        self.num_local_nodes = int(np.prod(self.local_grid))
        self.num_elements = int(np.prod([b[1] - b[0] for b in self.box]))
        self._build_sharing()

    def rank_coords(self, r):
        px, py, _ = self.ranks
        return (r % px, (r // px) % py, r // (px * py))

    def rank_of(self, c):
        px, py, _ = self.ranks
        return (c[2] * py + c[1]) * px + c[0]

    def element_table(self, p=None, element_order=None):
        """element_order: permutation of the local elements (see halo_independent_first)"""
        p = self.p if p is None else p
        nb = tuple((p * b[0], p * b[1]) for b in self.box)
        tab = self.mesh.element_table(p, self.box, nb)
        if self.new_of_lex is not None:
            assert p == self.p, "an owned-first partition numbers the nodes of its own order only"
            tab = self.new_of_lex[tab.astype(np.int64)].astype(np.uint64)
        if element_order is not None:
            tab = tab[np.asarray(element_order)]
        return tab

    def halo_independent_first(self):
        """-> (element_order, n_independent): a permutation of the local elements that puts first the elements all of
        whose nodes this rank owns -- they can be evaluated before P has delivered the values owned elsewhere -- and the
        halo-dependent ones last.  Use the permutation of the highest-order space for every space of the mesh (a vertex
        owned elsewhere is also a node of the higher-order space, so its mask contains the lower-order ones)."""
        tab = self.element_table().astype(np.int64)
        dependent = (self.owner[tab] != self.rank).any(axis=1)
        order = np.concatenate([np.nonzero(~dependent)[0], np.nonzero(dependent)[0]])
        return order, int((~dependent).sum())

    def vertex_coordinates(self):
        xyz = self.mesh.vertex_coordinates(tuple((b[0], b[1]) for b in self.box))
        if self.new_of_lex is not None:
            assert self.p == 1, "vertex coordinates in owned-first numbering come from the p = 1 partition"
            out = np.empty_like(xyz.reshape(-1, 3))
            out[self.new_of_lex] = xyz.reshape(-1, 3)
            xyz = out.ravel()
        return xyz

    def _lex_global_node_ids(self):
        G = self.mesh.node_grid(self.p)
        (i0, i1), (j0, j1), (k0, k1) = self.node_box
        K, J, I = np.meshgrid(np.arange(k0, k1 + 1), np.arange(j0, j1 + 1), np.arange(i0, i1 + 1), indexing="ij")
        return ((K * G[1] + J) * G[0] + I).ravel().astype(np.int64)

    def global_node_ids(self):
        """global lexicographic id of every local node (local order)"""
        gid = self._lex_global_node_ids()
        if self.new_of_lex is not None:
            out = np.empty_like(gid)
            out[self.new_of_lex] = gid
            gid = out
        return gid

    def _build_sharing(self):
        """which ranks hold each local node; owner = lowest (pfespace.cpp:3527-3545)"""
        p = self.p
        # per direction: for each local node index, the rank coordinates whose node range contains it
        holders = []
        for d in range(3):
            lo, hi = self.node_box[d]
            idx = np.arange(lo, hi + 1)
            hd = []
            for c, (s, e) in enumerate(self.splits[d]):
                if s == e:
                    continue
                hd.append((c, (idx >= p * s) & (idx <= p * e)))
            holders.append(hd)
        gx, gy, gz = self.local_grid
        owner = np.full((gz, gy, gx), self.rank, dtype=np.int64)
        self.shared_with = {}  # neighbour rank -> bool mask over local nodes
        for cz, mz in holders[2]:
            for cy, my in holders[1]:
                for cx, mx in holders[0]:
                    r = self.rank_of((cx, cy, cz))
                    if r == self.rank:
                        continue
                    mask = mz[:, None, None] & my[None, :, None] & mx[None, None, :]
                    if not mask.any():
                        continue
                    self.shared_with[r] = mask.ravel()
                    owner = np.where(mask & (r < owner), r, owner)
        self.owner = owner.ravel()
        self.neighbors = sorted(self.shared_with)
        if self.owned_first:
            mine = self.owner == self.rank
            order = np.concatenate([np.nonzero(mine)[0], np.nonzero(~mine)[0]])  # new id -> lexicographic id
            self.new_of_lex = np.empty(self.num_local_nodes, dtype=np.int64)
            self.new_of_lex[order] = np.arange(self.num_local_nodes)
            self.owner = self.owner[order]
            self.shared_with = {r: m[order] for r, m in self.shared_with.items()}
        owned = self.owner == self.rank
        self.num_true_nodes = int(owned.sum())
        self.true_of_local = np.full(self.num_local_nodes, -1, dtype=np.int64)
        self.true_of_local[owned] = np.arange(self.num_true_nodes)

    def halo_lists(self):
        """-> dict for sb_halo_create (vdofs, byVDIM): ltdof, neighbours, send_off/send_idx (true vdofs this rank owns and
        the neighbour holds), recv_off/recv_idx (local vdofs owned by the neighbour).  Within a neighbour, entries are
        ordered by global node id, then component -- the same order on both sides."""
        c = self.components
        gid = self.global_node_ids()
        send_idx, recv_idx, send_off, recv_off = [], [], [0], [0]
        for r in self.neighbors:
            mask = self.shared_with[r]
            mine = np.nonzero(mask & (self.owner == self.rank))[0]
            theirs = np.nonzero(mask & (self.owner == r))[0]
            mine = mine[np.argsort(gid[mine], kind="stable")]
            theirs = theirs[np.argsort(gid[theirs], kind="stable")]
            s = (self.true_of_local[mine][:, None] * c + np.arange(c)[None, :]).ravel()
            t = (theirs[:, None] * c + np.arange(c)[None, :]).ravel()
            send_idx.append(s)
            recv_idx.append(t)
            send_off.append(send_off[-1] + len(s))
            recv_off.append(recv_off[-1] + len(t))
        ltdof = np.where(self.true_of_local[:, None] >= 0, self.true_of_local[:, None] * c + np.arange(c)[None, :], -1)
        cat = lambda parts: (np.concatenate(parts) if parts else np.zeros(0)).astype(np.int32)
        return dict(lsize=self.num_local_nodes * c, tsize=self.num_true_nodes * c, ltdof=ltdof.ravel().astype(np.int32),
                    neighbors=list(self.neighbors), send_off=np.asarray(send_off, dtype=np.int64), send_idx=cat(send_idx),
                    recv_off=np.asarray(recv_off, dtype=np.int64), recv_idx=cat(recv_idx))

    def global_true_vdofs(self):
        """global vdof id (byVDIM on the global node grid) of every true vdof of this rank, in true-dof order"""
        c = self.components
        gid = self.global_node_ids()[self.owner == self.rank]
        return (gid[:, None] * c + np.arange(c)[None, :]).ravel()

    def global_local_vdofs(self):
        c = self.components
        gid = self.global_node_ids()
        return (gid[:, None] * c + np.arange(c)[None, :]).ravel()
