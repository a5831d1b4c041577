"""Shared-dof exchange P / P^T across ranks and the rank-parallel Functional built on it.

Reference: mfem::ConformingProlongationOperator::Mult / MultTranspose (mfem/fem/pfespace.cpp:3599-3663), which
serac::Functional applies around every evaluation (src/serac/numerics/functional/functional.hpp:446-498, 406-432).
One process per GPU; the pack / ordered-unpack kernels are the C ABI's sb_halo_* (no atomics: P^T adds the received
partial sums in a fixed neighbour order); the transport is torch.distributed point-to-point (NCCL over NVLink on GPUs).

`HostHalo` is the same index logic in numpy over any torch.distributed backend (gloo): it exists so the partition /
ownership / list construction can be tested on CPU; it is never used on the GPU path."""
import ctypes as C

import numpy as np

from . import capi


def _p2p(dist, sends, recvs, group=None):
    """sends / recvs: lists of (tensor view, peer rank); one batched round"""
    ops = [dist.P2POp(dist.irecv, t, peer, group) for t, peer in recvs if t.numel()]
    ops += [dist.P2POp(dist.isend, t, peer, group) for t, peer in sends if t.numel()]
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class HaloExchange:
    """P: T -> L and P^T: L -> T for one space on one rank (device buffers)."""

    def __init__(self, lists, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, device
        self.lists = lists
        self.lsize, self.tsize = int(lists["lsize"]), int(lists["tsize"])
        self.neighbors = list(lists["neighbors"])
        self.send_off = np.asarray(lists["send_off"], dtype=np.int64)
        self.recv_off = np.asarray(lists["recv_off"], dtype=np.int64)
        h = C.c_void_p()
        L = capi.lib()
        capi.check(L.sb_halo_create(self.lsize, self.tsize, capi._np_ptr(lists["ltdof"]), len(self.neighbors),
                                    capi._np_ptr(self.send_off), capi._np_ptr(lists["send_idx"]),
                                    capi._np_ptr(self.recv_off), capi._np_ptr(lists["recv_idx"]), C.byref(h)))
        self.h = h
        n = max(int(L.sb_halo_buffer_doubles(h)), 1)
        self.send_buf = torch.zeros(n, dtype=torch.float64, device=device)
        self.recv_buf = torch.zeros(n, dtype=torch.float64, device=device)
        capi.check(L.sb_halo_set_buffers(h, capi._dev_ptr(self.send_buf), capi._dev_ptr(self.recv_buf)))

        self.peer = False

    def __del__(self):
        try:
            capi.lib().sb_halo_destroy(self.h)
        except Exception:
            pass

    def enable_peer(self):
        """Switch the transport to the peer-memory kernels (sb_halo_*_peer): every rank's receive ring lives in
        torch symmetric memory (CUDA IPC-mapped over NVLink), neighbours write into it directly.  Collective: every
        rank of the group must call it.  Returns False (and keeps the NCCL transport) if symmetric memory is
        unavailable."""
        torch, dist = self.torch, self.dist
        import torch.distributed._symmetric_memory as symm_mem
        lib = capi.lib()
        MAXN = 32
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        slot = torch.tensor([max(int(lib.sb_halo_buffer_doubles(self.h)), 1)], dtype=torch.int64, device=self.device)
        dist.all_reduce(slot, op=dist.ReduceOp.MAX, group=self.group)
        slot = int(slot.item())
        meta = [None] * world
        dist.all_gather_object(meta, dict(neighbors=self.neighbors, send_off=self.send_off.tolist(),
                                          recv_off=self.recv_off.tolist()), group=self.group)
        try:
            grp = self.group if self.group is not None else dist.group.WORLD
            # two slots of (2 * slot + MAXN) 8-byte words: enough for the flag protocol and for the LL one (SB_HALO_LL=1)
            ring = symm_mem.empty(self.ring_doubles(slot), dtype=torch.float64, device=self.device)
            ring.zero_()
            hdl = symm_mem.rendezvous(ring, grp.group_name if hasattr(grp, "group_name") else grp)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[serac_b200] symmetric memory unavailable ({e!r}); keeping the NCCL transport", flush=True)
            return False
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # every ring is zeroed before anyone writes
        self.attach_rings(ring, slot, rank, ptrs, meta)
        self._hdl = hdl
        return True

    @staticmethod
    def ring_doubles(slot):
        """8-byte words of a receive ring for slots of `slot` doubles: two slots of (2 * slot + MAXN) words -- enough for
        the flag protocol and for the LL one"""
        return 4 * slot + 2 * 32

    def attach_rings(self, ring, slot, rank, ring_ptrs, meta):
        """ring: this rank's zeroed receive ring (ring_doubles(slot) doubles); ring_ptrs[r]: rank r's ring as addressable
        from this process (symmetric-memory mapping, CUDA IPC, or simply another tensor of this process when several
        ranks of a partition live on one device -- the loop-back tests); meta[r]: dict(neighbors, send_off, recv_off) of
        rank r's halo lists."""
        lib = capi.lib()
        nn = len(self.neighbors)
        peer_ring = (C.c_void_p * max(nn, 1))()
        P_off, Pt_off = np.zeros(max(nn, 1), dtype=np.int64), np.zeros(max(nn, 1), dtype=np.int64)
        my_idx = np.zeros(max(nn, 1), dtype=np.int32)
        for i, s in enumerate(self.neighbors):
            j = list(meta[s]["neighbors"]).index(rank)
            peer_ring[i] = int(ring_ptrs[s])
            P_off[i], Pt_off[i], my_idx[i] = meta[s]["recv_off"][j], meta[s]["send_off"][j], j
        capi.check(lib.sb_halo_enable_peer(self.h, capi._dev_ptr(ring), slot, peer_ring, capi._np_ptr(P_off),
                                           capi._np_ptr(Pt_off), capi._np_ptr(my_idx)))
        self._ring, self.peer = ring, True

    def peer_status(self):
        st = C.c_int(0)
        capi.check(capi.lib().sb_halo_peer_status(self.h, C.byref(st)))
        return st.value

    def _round(self, out_off, in_off):
        sends = [(self.send_buf[out_off[i]:out_off[i + 1]], r) for i, r in enumerate(self.neighbors)]
        recvs = [(self.recv_buf[in_off[i]:in_off[i + 1]], r) for i, r in enumerate(self.neighbors)]
        _p2p(self.dist, sends, recvs, self.group)

    def P(self, T, L=None):
        """T-vector -> L-vector: owned entries copied, entries owned elsewhere received from their owner"""
        lib = capi.lib()
        if L is None:
            L = self.torch.empty(self.lsize, dtype=self.torch.float64, device=self.device)
        s = capi._stream_ptr()
        if self.peer:
            capi.check(lib.sb_halo_P_peer(self.h, capi._dev_ptr(T), capi._dev_ptr(L), s))
            return L
        capi.check(lib.sb_halo_P_pack(self.h, capi._dev_ptr(T), s))
        self._round(self.send_off, self.recv_off)
        capi.check(lib.sb_halo_P_finish(self.h, capi._dev_ptr(T), capi._dev_ptr(L), s))
        return L

    # --- the two halves of the peer exchanges (put now, wait + unpack later) ---
    def P_put(self, T):
        assert self.peer, "split exchanges need the peer-memory transport"
        capi.check(capi.lib().sb_halo_P_peer_put(self.h, capi._dev_ptr(T), capi._stream_ptr()))

    def P_finish(self, T, L=None):
        if L is None:
            L = self.torch.empty(self.lsize, dtype=self.torch.float64, device=self.device)
        capi.check(capi.lib().sb_halo_P_peer_finish(self.h, capi._dev_ptr(T), capi._dev_ptr(L), capi._stream_ptr()))
        return L

    # --- in-place exchanges (owned-first numbering: the T-vector is the first tsize entries of the L-vector TL) ---
    def owned_first(self):
        return bool(capi.lib().sb_halo_owned_first(self.h))

    def P_finish_inplace(self, TL):
        capi.check(capi.lib().sb_halo_P_peer_finish_inplace(self.h, capi._dev_ptr(TL), capi._stream_ptr()))
        return TL

    def P_inplace(self, TL):
        """fill the entries of TL owned by other ranks from their owners; the owned prefix is the T-vector, untouched
        (one fused put + wait + unpack kernel)"""
        hs, vs = (C.c_void_p * 1)(self.h), (C.c_void_p * 1)(TL.data_ptr())
        capi.check(capi.lib().sb_halo_P_peer_inplace_multi(1, hs, vs, capi._stream_ptr()))
        return TL

    def Pt_inplace(self, TL):
        """add every sharer's partial sums into the owned shared entries of TL; TL[:tsize] is then the T-vector"""
        capi.check(capi.lib().sb_halo_Pt_peer_inplace(self.h, capi._dev_ptr(TL), capi._stream_ptr()))
        return TL[:self.tsize]

    def Pt(self, L, T=None):
        """L-vector -> T-vector: owned entries plus the partial sums every sharer computed, in fixed neighbour order"""
        lib = capi.lib()
        if T is None:
            T = self.torch.empty(self.tsize, dtype=self.torch.float64, device=self.device)
        s = capi._stream_ptr()
        if self.peer:
            capi.check(lib.sb_halo_Pt_peer(self.h, capi._dev_ptr(L), capi._dev_ptr(T), s))
            return T
        capi.check(lib.sb_halo_Pt_pack(self.h, capi._dev_ptr(L), s))
        self._round(self.recv_off, self.send_off)  # P^T sends what P received and vice versa
        capi.check(lib.sb_halo_Pt_finish(self.h, capi._dev_ptr(L), capi._dev_ptr(T), s))
        return T


class HostHalo:
    """numpy restatement of HaloExchange's semantics (CPU tests of the partition logic only)."""

    def __init__(self, lists, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.l = torch, dist, group, lists
        self.neighbors = list(lists["neighbors"])

    def _round(self, send, out_off, in_off):
        t = self.torch
        sbuf = t.as_tensor(np.ascontiguousarray(send))
        rbuf = t.zeros(int(in_off[-1]), dtype=t.float64)
        sends = [(sbuf[out_off[i]:out_off[i + 1]], r) for i, r in enumerate(self.neighbors)]
        recvs = [(rbuf[in_off[i]:in_off[i + 1]], r) for i, r in enumerate(self.neighbors)]
        _p2p(self.dist, sends, recvs, self.group)
        return rbuf.numpy()

    def P(self, T):
        l = self.l
        recv = self._round(T[l["send_idx"]], l["send_off"], l["recv_off"])
        L = np.zeros(l["lsize"])
        own = l["ltdof"] >= 0
        L[own] = T[l["ltdof"][own]]
        L[l["recv_idx"]] = recv
        return L

    def Pt(self, L):
        l = self.l
        recv = self._round(L[l["recv_idx"]], l["recv_off"], l["send_off"])
        T = np.zeros(l["tsize"])
        own = l["ltdof"] >= 0
        T[l["ltdof"][own]] = L[own]
        for k in range(len(recv)):  # ascending slot order = neighbour order, then list order
            T[l["send_idx"][k]] += recv[k]
        return T


class ParFunctional:
    """serac::Functional on a partitioned mesh (functional.hpp:446-498): T-vectors in, T-vector out.

    Exchange scheduling (peer transport; every field in flight has its own halo object):
      * the puts of all inputs are issued before the first flag wait, so the waits overlap;
      * with `overlap_halos` (one more halo object per space: du_halos[wrt] for Jacobian-action directions, out_halo for
        the residual's P^T) the exchange of the direction can be started early with prefetch_direction() and the
        residual's P^T runs on a side stream next to the following kernels; join() makes the calling stream wait for
        everything the side stream still owes."""

    def __init__(self, functional, trial_halos, test_halo, du_halos=None, out_halo=None, inplace=False,
                 independent_elements=0):
        """inplace: the partitions are numbered owned-first and every vector handed in is an L-sized buffer whose first
        tsize entries are the T-vector; results are returned as views of the Functional's L-vectors.  P / P^T then touch
        only the shared entries (no full-vector copies)."""
        self.F, self.trial_halos, self.test_halo = functional, list(trial_halos), test_halo
        self.du_halos, self.out_halo = dict(du_halos or {}), out_halo
        # elements [0, independent_elements) of the (single, whole-mesh) domain integral touch no node owned elsewhere
        # (Partition.halo_independent_first): they are evaluated while the input exchange is in flight
        self.independent_elements = int(independent_elements)
        self.inplace = bool(inplace)
        if self.independent_elements > 0:
            import os
            # the two halves of the input exchange exist for the flag protocol only (sb_halo_P_peer_put_multi)
            if os.environ.get("SB_HALO_LL", "1") != "0":
                raise ValueError("independent_elements > 0 needs the flag protocol: set SB_HALO_LL=0")
            ids = [id(h) for h in trial_halos]
            if len(set(ids)) != len(ids) or len(ids) > 4:
                raise ValueError("independent_elements > 0 needs one halo object per trial field, at most 4 fields")
        if self.inplace:
            hs = list(trial_halos) + [test_halo] + list(self.du_halos.values()) + ([out_halo] if out_halo else [])
            assert all(h.peer and h.owned_first() for h in hs), "in-place exchanges need owned-first peer halos"
        self._L = [None] * len(self.trial_halos)
        self._rL = None
        self._drL = None
        self._duL = {}
        self._side = None
        self._du_ready = {}
        # distinct objects are needed for fields whose exchanges are in flight together
        ids = [id(h) for h in self.trial_halos]
        self._split = all(h.peer for h in self.trial_halos) and len(set(ids)) == len(ids)

    def _side_stream(self):
        if self._side is None:
            self._side = capi_torch().cuda.Stream()
        return self._side

    def prefetch_direction(self, wrt, dT):
        """start P(dT) for a later GradientMult(wrt, dT) on the side stream (needs du_halos[wrt])"""
        torch = capi_torch()
        h = self.du_halos[wrt]
        side, main = self._side_stream(), torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._duL[wrt] = h.P_inplace(dT) if self.inplace else h.P(dT, self._duL.get(wrt))
            ev = torch.cuda.Event()
            ev.record(side)
        self._du_ready[wrt] = (ev, dT.data_ptr())

    def __call__(self, t, T_inputs, wrt=capi.NO_DIFFERENTIATION, out=None, update_state=False):
        torch = capi_torch()
        if self.inplace:
            assert out is None, "in-place mode returns a view of the Functional's own L-vector"
            if self._split and len(self.trial_halos) <= 4:  # one put launch + one wait/unpack launch for every input
                hs = (C.c_void_p * len(self.trial_halos))(*[h.h for h in self.trial_halos])
                vs = (C.c_void_p * len(T_inputs))(*[TL.data_ptr() for TL in T_inputs])
                n, lib = len(self.trial_halos), capi.lib()
                self._L = list(T_inputs)
                if self.independent_elements > 0:
                    # puts -> elements that need no halo data -> wait + unpack -> the halo-dependent elements + scatter
                    capi.check(lib.sb_halo_P_peer_put_multi(n, hs, vs, capi._stream_ptr()))
                    self.F.evaluate_begin(t, self._L, wrt=wrt, update_state=update_state,
                                          split_element=self.independent_elements)
                    capi.check(lib.sb_halo_P_peer_finish_inplace_multi(n, hs, vs, capi._stream_ptr()))
                    self._rL = self.F.evaluate_end(out=self._rL)
                    return self._finish_output(out)
                capi.check(lib.sb_halo_P_peer_inplace_multi(n, hs, vs, capi._stream_ptr()))
            else:
                # fields that share a halo object (or more than 4 of them): one whole exchange per field through the same
                # entry point as everywhere else -- a halo object serves one field in flight and stays with one protocol
                for i, (h, TL) in enumerate(zip(self.trial_halos, T_inputs)):
                    self._L[i] = h.P_inplace(TL)
        elif self._split:
            for h, T in zip(self.trial_halos, T_inputs):
                h.P_put(T)
            for i, (h, T) in enumerate(zip(self.trial_halos, T_inputs)):
                self._L[i] = h.P_finish(T, self._L[i])
        else:
            for i, (h, T) in enumerate(zip(self.trial_halos, T_inputs)):
                self._L[i] = h.P(T, self._L[i])
        self._rL = self.F(t, self._L, wrt=wrt, out=self._rL, update_state=update_state)
        return self._finish_output(out)

    def _finish_output(self, out):
        """P^T of the residual L-vector (on the side stream when an out_halo was given)"""
        torch = capi_torch()
        if self.out_halo is not None:
            side, main = self._side_stream(), torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = self.out_halo.Pt_inplace(self._rL) if self.inplace else self.out_halo.Pt(self._rL, out)
            return out  # valid on the calling stream after join()
        return self.test_halo.Pt_inplace(self._rL) if self.inplace else self.test_halo.Pt(self._rL, out)

    def GradientMult(self, wrt, dT, out=None):
        torch = capi_torch()
        pre = self._du_ready.pop(wrt, None)
        if pre is not None and pre[1] == dT.data_ptr():
            torch.cuda.current_stream().wait_event(pre[0])
        elif self.inplace:
            self._duL[wrt] = self.trial_halos[wrt].P_inplace(dT)
        else:
            self._duL[wrt] = self.trial_halos[wrt].P(dT, self._duL.get(wrt))
        # its own L-vector: the residual's may still be read by the side stream's P^T
        self._drL = self.F.GradientMult(wrt, self._duL[wrt], out=self._drL)
        if self.inplace:
            assert out is None
            return self.test_halo.Pt_inplace(self._drL)
        return self.test_halo.Pt(self._drL, out)

    def join(self):
        """the calling stream waits for the side stream (results of __call__ with out_halo become valid)"""
        if self._side is not None:
            capi_torch().cuda.current_stream().wait_stream(self._side)

    def check_peers(self):
        """synchronises and raises if a wait for a neighbour gave up (~2 s): the result of that step is not valid.
        Call it after a step (or every few steps) -- a lost or stalled peer must not turn into silently wrong numbers."""
        capi_torch().cuda.synchronize()
        hs = list(self.trial_halos) + [self.test_halo] + list(self.du_halos.values())
        hs += [self.out_halo] if self.out_halo is not None else []
        bad = [i for i, h in enumerate(hs) if h.peer and h.peer_status() != 0]
        if bad:
            raise RuntimeError(f"shared-dof exchange: wait for a neighbour timed out on halo object(s) {bad}")


def capi_torch():
    import torch
    return torch


# ---- the Newton step's linear algebra at N ranks ----------------------------------------------------------------------
class ParCsrOperator:
    """The assembled Jacobian of a partitioned Functional as an operator on T-vectors.

    The reference turns the rank-local matrix A_L into K = P^T A_L P with hypre's RAP (functional.hpp:654-662) and hands
    K to hypre's solvers.  Without hypre the product is kept in its factors -- the *sub-assembled* form: every rank holds
    the CSR matrix of its own elements over its L-dofs (what sb_functional_assemble writes, no communication), and

        y_T = P^T ( A_L ( P x_T ) )

    is the action of the same K: P and P^T are the shared-dof exchanges that already surround every Functional
    evaluation.  diag(K) = P^T diag(A_L) gives the Jacobi preconditioner.  Essential dofs are eliminated rank-locally
    (rows and columns of every local copy); the unit diagonal is kept on the owner's copy only, so that P^T adds up to
    exactly 1 -- K is then [K_ii 0; 0 I] as eliminateAllEssentialDofsFromMatrix leaves it
    (boundary_condition_manager.hpp:91-94)."""

    def __init__(self, pattern, values, halo, ess_local=None):
        """pattern / values: the local CSR triple (L x L); halo: HaloExchange of the space (separate T / L vectors);
        ess_local: local L-dof ids of the essential dofs (every local copy, owned or not)"""
        import torch
        self.pat, self.vals, self.halo, self.torch = pattern, values, halo, torch
        self._xL = self._yL = None
        if ess_local is not None and len(ess_local):
            ess_local = np.ascontiguousarray(ess_local, dtype=np.int32)
            capi.eliminate_rows_cols(pattern, values, ess_local, want_Ae=False)
            ltdof = np.asarray(halo.lists["ltdof"])
            ghost = ess_local[ltdof[ess_local] < 0]  # copies owned by another rank: their diagonal must not add up
            if len(ghost):
                rp, ci = pattern.to_host()
                at = np.array([rp[i] + np.searchsorted(ci[rp[i]:rp[i + 1]], i) for i in ghost], dtype=np.int64)
                values[torch.as_tensor(at, device=values.device)] = 0.0

    def Mult(self, x_T, out=None):
        self._xL = self.halo.P(x_T, self._xL)
        self._yL = capi.csr_spmv(self.pat, self.vals, self._xL, out=self._yL)
        return self.halo.Pt(self._yL, out)

    def diagonal(self):
        """diag(P^T A_L P) as a T-vector"""
        return self.halo.Pt(capi.csr_diagonal(self.pat, self.vals))


def par_dot(x_T, y_T, group=None):
    """(x, y) over the true dofs of all ranks: every true dof lives on exactly one rank"""
    import torch
    import torch.distributed as dist
    d = torch.dot(x_T, y_T).reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.SUM, group=group)
    return float(d.item())


def par_cg(A, b_T, dinv_T=None, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, group=None):
    """mfem::CGSolver (mfem/linalg/solvers.cpp:719-897) on T-vectors across the ranks: the same recurrences and stopping
    test as the rank-local sb_cg_*, with every inner product reduced over the ranks.  A: object with Mult(x_T) -> y_T;
    dinv_T: inverse diagonal (Jacobi) or None.  -> x_T, info"""
    torch = capi_torch()
    x = torch.zeros_like(b_T)
    r = b_T.clone()
    z = r * dinv_T if dinv_T is not None else r.clone()
    d = z.clone()
    nom0 = nom = par_dot(d, r, group)
    info = dict(converged=False, final_iter=0, initial_norm=float(np.sqrt(max(nom0, 0.0))))
    if nom < 0:
        return x, info
    r0 = max(nom * rel_tol * rel_tol, abs_tol * abs_tol)
    if nom <= r0:
        info.update(converged=True, final_norm=float(np.sqrt(nom)))
        return x, info
    z = A.Mult(d)
    den = par_dot(z, d, group)
    if den <= 0.0:
        return x, info
    i = 1
    while True:
        alpha = nom / den
        x.add_(d, alpha=alpha)
        r.add_(z, alpha=-alpha)
        z = r * dinv_T if dinv_T is not None else r.clone()
        betanom = par_dot(r, z, group)
        if betanom <= r0:
            info.update(converged=True, final_iter=i, final_norm=float(np.sqrt(max(betanom, 0.0))))
            return x, info
        if i + 1 > max_iter:
            info.update(final_iter=i, final_norm=float(np.sqrt(max(betanom, 0.0))))
            return x, info
        i += 1
        beta = betanom / nom
        d.mul_(beta).add_(z)
        z = A.Mult(d)
        den = par_dot(d, z, group)
        if den <= 0.0:
            info.update(final_iter=i, final_norm=float(np.sqrt(max(betanom, 0.0))))
            return x, info
        nom = betanom
