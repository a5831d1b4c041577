"""CPU: the structured-mesh partitioner and the shared-dof exchange logic (serac_b200/structured.py, parallel.py), incl.
world_size-2 and -4 runs over the gloo backend.  The distributed residual is checked against the numpy oracle evaluated
on the global mesh (oracle as checker only)."""
import os
import socket

import numpy as np
import pytest

from harness import rel_l2
from serac_b200.structured import Partition, StructuredHexMesh, factor_ranks


def test_factor_ranks():
    assert factor_ranks(1) == (1, 1, 1)
    assert sorted(factor_ranks(2)) == [1, 1, 2] and sorted(factor_ranks(4)) == [1, 2, 2] and factor_ranks(8) == (2, 2, 2)
    assert factor_ranks(4, (128, 16, 16)) == (4, 1, 1)  # beams are cut along their length


def test_element_table_matches_lexicographic_convention():
    m = StructuredHexMesh(2, 1, 1)
    t = m.element_table(1)
    # nodes of a 3 x 2 x 2 grid; element 0: x fastest, then y, then z
    assert t.tolist() == [[0, 1, 3, 4, 6, 7, 9, 10], [1, 2, 4, 5, 7, 8, 10, 11]]
    assert m.element_table(2).shape == (2, 27) and m.num_nodes(2) == 5 * 3 * 3


@pytest.mark.parametrize("ranks", [(2, 2, 1), (2, 2, 2), (3, 1, 2)])
@pytest.mark.parametrize("p", [1, 2])
def test_owned_first_numbering(ranks, p):
    """owned_first=True: the same partition, renumbered so that the T-vector is the prefix of the L-vector (what the
    in-place exchanges sb_halo_*_inplace require); element tables and node ids follow the renumbering"""
    mesh = StructuredHexMesh(5, 4, 3)
    world = int(np.prod(ranks))
    for r in range(world):
        a, b = Partition(mesh, ranks, r, p, components=3), Partition(mesh, ranks, r, p, components=3, owned_first=True)
        l = b.halo_lists()
        assert np.array_equal(l["ltdof"][:l["tsize"]], np.arange(l["tsize"])) and (l["ltdof"][l["tsize"]:] == -1).all()
        assert np.array_equal(a.global_true_vdofs(), b.global_true_vdofs())  # owned nodes keep their relative order
        # same elements, same global nodes
        assert np.array_equal(a.global_node_ids()[a.element_table().astype(np.int64)],
                              b.global_node_ids()[b.element_table().astype(np.int64)])
        assert sorted(a.global_node_ids()) == sorted(b.global_node_ids())
    if p == 1:
        a, b = Partition(mesh, ranks, world - 1, 1, 3), Partition(mesh, ranks, world - 1, 1, 3, owned_first=True)
        xa, xb = a.vertex_coordinates().reshape(-1, 3), b.vertex_coordinates().reshape(-1, 3)
        assert np.array_equal(xa[np.argsort(a.global_node_ids())], xb[np.argsort(b.global_node_ids())])


@pytest.mark.parametrize("ranks", [(2, 1, 1), (2, 2, 2)])
def test_halo_independent_elements_first(ranks):
    """the first n_independent elements of the reordered table touch only nodes this rank owns, the rest at least one
    node owned elsewhere; the permutation carries over to the lower-order space of the same mesh"""
    mesh = StructuredHexMesh(5, 4, 3)
    world = int(np.prod(ranks))
    for r in range(world):
        part = Partition(mesh, ranks, r, 2, 3, owned_first=True)
        part1 = Partition(mesh, ranks, r, 1, 3, owned_first=True)
        order, n_ind = part.halo_independent_first()
        assert sorted(order.tolist()) == list(range(part.num_elements))
        tab = part.element_table(element_order=order).astype(np.int64)
        owned = part.owner[tab] == r
        assert owned[:n_ind].all() and (~owned[n_ind:]).any(axis=1).all()
        tab1 = part1.element_table(element_order=order).astype(np.int64)
        assert (part1.owner[tab1[:n_ind]] == r).all()
        # same elements in the same order in both spaces: the vertices of an element are among its p = 2 nodes
        g2, g1 = part.global_node_ids()[tab], part1.global_node_ids()[tab1]
        G2, G1 = mesh.node_grid(2), mesh.node_grid(1)
        corner2 = g2[:, 0]
        corner1 = g1[:, 0]
        i2, j2, k2 = corner2 % G2[0], (corner2 // G2[0]) % G2[1], corner2 // (G2[0] * G2[1])
        i1, j1, k1 = corner1 % G1[0], (corner1 // G1[0]) % G1[1], corner1 // (G1[0] * G1[1])
        assert np.array_equal(i2, 2 * i1) and np.array_equal(j2, 2 * j1) and np.array_equal(k2, 2 * k1)
    if world > 1:
        assert Partition(mesh, ranks, 0, 2, 3, owned_first=True).halo_independent_first()[1] == \
            Partition(mesh, ranks, 0, 2, 3, owned_first=True).num_elements  # the lowest rank owns all its nodes


@pytest.mark.parametrize("owned_first", [False, True])
@pytest.mark.parametrize("ranks", [(2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 2)])
@pytest.mark.parametrize("p", [1, 2])
def test_ownership_and_lists_are_consistent(ranks, p, owned_first):
    mesh = StructuredHexMesh(5, 4, 3)
    world = int(np.prod(ranks))
    parts = [Partition(mesh, ranks, r, p, components=3, owned_first=owned_first) for r in range(world)]
    # every global vdof is owned exactly once
    owned = np.concatenate([pt.global_true_vdofs() for pt in parts])
    assert len(owned) == mesh.num_nodes(p) * 3 and len(np.unique(owned)) == len(owned)
    assert sum(pt.num_elements for pt in parts) == mesh.ne
    lists = [pt.halo_lists() for pt in parts]
    for r, (pt, l) in enumerate(zip(parts, lists)):
        gl, gt = pt.global_local_vdofs(), pt.global_true_vdofs()
        for i, s in enumerate(l["neighbors"]):
            j = lists[s]["neighbors"].index(r)
            mine = gt[l["send_idx"][l["send_off"][i]:l["send_off"][i + 1]]]
            theirs = parts[s].global_local_vdofs()[lists[s]["recv_idx"][lists[s]["recv_off"][j]:lists[s]["recv_off"][j + 1]]]
            assert np.array_equal(mine, theirs)  # same dofs in the same order on both sides
        # the owner of a shared node is the lowest rank that holds it
        for s in l["neighbors"]:
            shared = pt.shared_with[s]
            assert (pt.owner[shared] <= min(r, s)).all()
        assert (l["ltdof"] >= 0).sum() == l["tsize"] and len(gl) == l["lsize"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ranks, p, out):
    import torch.distributed as dist
    from oracle import functional_oracle as fo
    from serac_b200.parallel import HostHalo
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = StructuredHexMesh(4, 3, 2, 2.0, 1.5, 1.0)
        c = 3
        pt = Partition(mesh, ranks, rank, p, components=c)
        halo = HostHalo(pt.halo_lists())
        f = lambda g: np.sin(0.37 * g) + 0.01 * g  # a globally consistent vector
        # --- P: T -> L reproduces the global vector on every local dof ---
        L = halo.P(f(pt.global_true_vdofs().astype(np.float64)))
        assert np.array_equal(L, f(pt.global_local_vdofs().astype(np.float64)))
        # --- local residual with the numpy oracle, then P^T; compare with the oracle on the global mesh ---
        U_glob = 0.01 * f(np.arange(mesh.num_nodes(p) * c, dtype=np.float64))
        qf = fo.make_qf("solid_neohookean", [1.0, 100.0, 50.0, 1.0])

        def residual(table, shape_table, coords, U, nnodes):
            q = p + 1
            X_q, J_q = fo.geometric_factors(fo.gather(shape_table, coords, 3), 1, q, 3)
            trials = [fo.gather(shape_table, np.zeros_like(coords), 3), fo.gather(table, U, c),
                      fo.gather(table, np.zeros_like(U), c)]
            r_E, _ = fo.evaluation_kernel(qf, trials, [1, p, p], X_q, J_q, p, q, 3, wrt=-1)
            return fo.scatter_add(table, r_E, np.zeros(nnodes * c), c)

        r_loc = residual(pt.element_table(p), pt.element_table(1), pt.vertex_coordinates(),
                         U_glob[pt.global_local_vdofs()], pt.num_local_nodes)
        r_T = halo.Pt(r_loc)
        r_glob = residual(mesh.element_table(p), mesh.element_table(1), mesh.vertex_coordinates(), U_glob,
                          mesh.num_nodes(p))
        err = rel_l2(r_T, r_glob[pt.global_true_vdofs()])
        assert err < 1e-13, err
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ranks", [(2, (2, 1, 1)), (4, (2, 2, 1))])
def test_distributed_residual_over_gloo(world, ranks):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ranks, 2, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [out.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res


@pytest.mark.parametrize("mesh_n,p", [((7, 7, 7), 3), ((5, 4, 3), 3), ((7, 6, 5), 2)])
def test_exchange_lists_of_uneven_2x2x2_partitions(mesh_n, p):
    """The shape of the strong config-B series at 8 GPUs (BASELINE configs[4]: 135 = 68 + 67 hexes per direction, p = 3):
    boxes of unequal size, seven neighbours per rank incl. edge and corner neighbours.  All eight ranks in one process:
    P and P^T are played through the exchange lists exactly as the device kernels use them (owner -> sharers by send /
    recv segments, sharers -> owner in ascending slot order) and compared with the global vectors."""
    ranks = (2, 2, 2)
    mesh = StructuredHexMesh(*mesh_n)
    parts = [Partition(mesh, ranks, r, p, components=3, owned_first=True) for r in range(8)]
    lists = [pt.halo_lists() for pt in parts]
    nglob = mesh.num_nodes(p) * 3
    assert {pt.num_elements for pt in parts} != {mesh.ne // 8}, "the boxes are meant to differ in size"
    rng = np.random.default_rng(1)
    g = rng.uniform(-1, 1, nglob)

    def segment(l, key_off, i):
        return slice(int(l[key_off][i]), int(l[key_off][i + 1]))

    # ---- P: every local copy equals the global value; owned-first numbering: the T-vector is the prefix ----
    Ls = []
    for r, (pt, l) in enumerate(zip(parts, lists)):
        L = np.full(l["lsize"], np.nan)
        L[:l["tsize"]] = g[pt.global_true_vdofs()]
        assert np.array_equal(l["ltdof"][:l["tsize"]], np.arange(l["tsize"]))  # in place: T is the prefix of L
        Ls.append(L)
    for r, l in enumerate(lists):  # owner r -> sharer s
        for i, s in enumerate(l["neighbors"]):
            j = list(lists[s]["neighbors"]).index(r)
            sent = Ls[r][:l["tsize"]][l["send_idx"][segment(l, "send_off", i)]]
            Ls[s][lists[s]["recv_idx"][segment(lists[s], "recv_off", j)]] = sent
    for pt, L in zip(parts, Ls):
        assert np.array_equal(L, g[pt.global_local_vdofs()])
    # ---- P^T: the owner adds every sharer's partial sum; the result is the sum over all copies ----
    parts_L = [rng.uniform(-1, 1, l["lsize"]) for l in lists]
    tot = np.zeros(nglob)
    for pt, L in zip(parts, parts_L):
        np.add.at(tot, pt.global_local_vdofs(), L)
    for r, (pt, l) in enumerate(zip(parts, lists)):
        T = parts_L[r][:l["tsize"]].copy()
        for i, s in enumerate(l["neighbors"]):  # fixed neighbour order
            j = list(lists[s]["neighbors"]).index(r)
            theirs = parts_L[s][lists[s]["recv_idx"][segment(lists[s], "recv_off", j)]]
            np.add.at(T, l["send_idx"][segment(l, "send_off", i)], theirs)
        assert np.allclose(T, tot[pt.global_true_vdofs()], rtol=0, atol=1e-13)
