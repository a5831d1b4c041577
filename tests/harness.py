"""Shared test plumbing: load golden fixtures / live reference problems and drive the product (C ABI through
serac_b200.capi) on them.  The oracle (oracle/) is only ever the checker here."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

KIND_TO_NAME = {5: "hex", 3: "quad", 4: "tet", 2: "tri"}


def golden_names():
    """small fixtures with full reference outputs (the config_* files are full-size inputs + sampled outputs)"""
    names = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if not n.startswith(("config_", "bdr_", "multi_", "qoi_", "bqoi_", "l2param_", "krylov_", "newton_", "warmstart_", "l2faces_"))]


def multifield_golden_names():
    """parameterized-physics fixtures (oracle/make_golden.py MULTIFIELD_FIXTURES)"""
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "multi_*.npz")))


def boundary_golden_names():
    """boundary-integral fixtures (oracle/make_golden.py BOUNDARY_FIXTURES)"""
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "bdr_*.npz")))


def krylov_golden_names():
    """Newton-step linear-algebra fixtures (oracle/make_golden_krylov.py): the reference's mfem on reference Jacobians"""
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "krylov_*.npz")))


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (d[k].item() if d[k].ndim == 0 else d[k]) for k in d.files}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def make_inputs(seed, vsize, shape_vsize, shape_amp=0.0, u_amp=0.01):
    """the same draws, in the same order, as oracle.make_golden.reference_problem"""
    rng = np.random.default_rng(seed)
    U = u_amp * rng.uniform(-1.0, 1.0, vsize)
    A = rng.uniform(-1.0, 1.0, vsize)
    S = shape_amp * rng.uniform(-1.0, 1.0, shape_vsize)
    dU = rng.uniform(0.0, 1.0, vsize)
    return dict(U=U, A=A, S=S, dU=dU)


class Product:
    """The product objects for one parity problem (same signature as oracle.make_golden.reference_problem)."""

    def __init__(self, prob, device, name_suffix=""):
        import torch
        from serac_b200 import capi
        self.capi, self.torch, self.prob, self.dev = capi, torch, prob, device
        geom, p, dim, vdim = int(prob["geom"]), int(prob["p"]), int(prob["dim"]), int(prob["vdim"])
        self.geom, self.p, self.dim, self.vdim = geom, p, dim, vdim
        # restriction tables built by the PRODUCT from the host library's native dofs
        self.dof_info = capi.h1_dof_info(geom, p, prob["native_dofs"])
        self.shape_dof_info = capi.h1_dof_info(geom, 1, prob["shape_native_dofs"])
        self.G = capi.ElementRestriction(self.dof_info, vdim, int(prob["num_nodes"]))
        self.G_shape = capi.ElementRestriction(self.shape_dof_info, dim, int(prob["shape_num_nodes"]))
        # the mesh-nodes space is H1 p=1 with the same scalar numbering as the shape space
        mesh_info = prob["mesh_dof_info"] if "mesh_dof_info" in prob else self.shape_dof_info
        self.G_mesh = capi.ElementRestriction(mesh_info, dim, int(prob["mesh_num_nodes"]),
                                              int(prob["mesh_ordering"]))
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
        self.t = t
        self.nodes = t(prob["mesh_nodes"])
        # integrals on part of the mesh (Domain::ofElements): geometric factors, stored derivatives and state are indexed
        # by position in the domain, E-vectors by element id (domain_integral_kernels.hpp:183-216)
        self.domain = prob["domain_elements"].astype(np.int32) if "domain_elements" in prob else None
        if self.domain is None:
            self.X_q, self.J_q = capi.geometric_factors(geom, int(prob["mesh_order"]), p + 1, self.G_mesh, self.nodes)
            ne_dom = int(prob["ne"])
        else:
            self._domain_dev = torch.as_tensor(self.domain, dtype=torch.int32, device=device)
            ne_dom = len(self.domain)
            self.X_q, self.J_q = capi.geometric_factors(geom, int(prob["mesh_order"]), p + 1, self.G_mesh, self.nodes,
                                                        num_elements=ne_dom, elements=self._domain_dev)
        name = f"{prob['kind']}_{KIND_TO_NAME[geom]}_p{p}{name_suffix}"
        self.integral = capi.Integral(name, prob["params"], self.X_q, self.J_q, ne_dom, elements=self.domain)
        if "U" not in prob:  # full-size fixtures regenerate the inputs from the seed (oracle/make_golden.py)
            prob.update(make_inputs(int(prob["seed"]), self.G.LSize(), self.G_shape.LSize(), u_amp=float(prob["u_amp"])))
        self.U, self.A, self.S, self.dU = t(prob["U"]), t(prob["A"]), t(prob["S"]), t(prob["dU"])
        self.F = capi.Functional(self.G, [self.G_shape, self.G, self.G])
        self.F.AddDomainIntegral(self.integral, [0, 1, 2])
        if self.integral.table.state_bytes:
            self.qdata = t(prob["qdata0"]).clone()
            self.integral.set_qdata(self.qdata)

    def fields(self):
        return [self.G_shape.field(self.S), self.G.field(self.U), self.G.field(self.A)]


class BoundaryProduct:
    """The product objects for one boundary-integral parity problem (oracle.make_golden.reference_boundary_problem):
    Functional H1<p,dim>(shape, u, a) with ONE boundary integral on EntireBoundary (functional.hpp:328-354)."""

    def __init__(self, prob, device):
        import torch
        from serac_b200 import capi
        self.capi, self.torch, self.prob, self.dev = capi, torch, prob, device
        geom, fgeom, p, dim = int(prob["geom"]), int(prob["face_geom"]), int(prob["p"]), int(prob["dim"])
        self.p, self.dim = p, dim
        # domain restrictions: the Functional's spaces and the sparsity pattern (functional.hpp:582-583)
        self.G = capi.ElementRestriction(capi.h1_dof_info(geom, p, prob["native_dofs"]), dim, int(prob["num_nodes"]))
        # boundary-face restrictions built by the PRODUCT from mfem's native face dofs (fes->GetFaceDofs)
        self.face_dof_info = capi.h1_dof_info(fgeom, p, prob["face_native_dofs"])
        self.shape_face_dof_info = capi.h1_dof_info(fgeom, 1, prob["shape_face_native_dofs"])
        self.Gf = capi.ElementRestriction(self.face_dof_info, dim, int(prob["num_nodes"]))
        self.Gf_shape = capi.ElementRestriction(self.shape_face_dof_info, dim, int(prob["shape_num_nodes"]))
        mesh_face_info = capi.h1_dof_info(fgeom, int(prob["mesh_order"]), prob["mesh_face_native_dofs"])
        self.Gf_mesh = capi.ElementRestriction(mesh_face_info, dim, int(prob["mesh_num_nodes"]),
                                               int(prob["mesh_ordering"]))
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
        self.t = t
        self.nodes = t(prob["mesh_nodes"])
        self.X_q, self.J_q = capi.geometric_factors(fgeom, int(prob["mesh_order"]), p + 1, self.Gf_mesh, self.nodes)
        name = f"{prob['kind']}_{KIND_TO_NAME[geom]}_p{p}"
        self.integral = capi.Integral(name, prob["params"], self.X_q, self.J_q, int(prob["nf"]))
        self.U, self.A, self.S, self.dU, self.dS = (t(prob[k]) for k in ("U", "A", "S", "dU", "dS"))
        # shape space domain restriction is not needed by a boundary-only Functional: pass the face one
        self.F = capi.Functional(self.G, [self.Gf_shape, self.G, self.G])
        self.F.AddBoundaryIntegral(self.integral, [0, 1, 2], self.Gf, [self.Gf_shape, self.Gf, self.Gf])

    def fields(self):
        return [self.Gf_shape.field(self.S), self.Gf.field(self.U), self.Gf.field(self.A)]


class MultiFieldProduct:
    """The product objects for one multi-field parity problem (oracle.make_golden.reference_multifield_problem):
    Functional H1<p,test_vdim>(shape H1<1,dim>, H1<p,v1>, H1<p,v2>, H1<p,v3>) with one domain integral -- the
    signatures of the parameterized physics modules (thermomechanics.hpp:534-537)."""

    def __init__(self, prob, device):
        import torch
        from serac_b200 import capi
        self.capi, self.torch, self.prob, self.dev = capi, torch, prob, device
        geom, p, dim = int(prob["geom"]), int(prob["p"]), int(prob["dim"])
        self.geom, self.p, self.dim = geom, p, dim
        self.dof_info = capi.h1_dof_info(geom, p, prob["native_dofs"])
        self.shape_dof_info = capi.h1_dof_info(geom, 1, prob["shape_native_dofs"])
        nn = int(prob["num_nodes"])
        vdims = [int(v) for v in prob["trial_vdims"]]
        self.G_shape = capi.ElementRestriction(self.shape_dof_info, dim, int(prob["shape_num_nodes"]))
        by_vdim = {}
        for v in set(vdims[1:]) | {int(prob["test_vdim"])}:
            by_vdim[v] = capi.ElementRestriction(self.dof_info, v, nn)
        self.G_test = by_vdim[int(prob["test_vdim"])]
        self.G_trials = [self.G_shape] + [by_vdim[v] for v in vdims[1:]]
        self.G_mesh = capi.ElementRestriction(prob["mesh_dof_info"], dim, int(prob["mesh_num_nodes"]),
                                              int(prob["mesh_ordering"]))
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
        self.t = t
        self.nodes = t(prob["mesh_nodes"])
        self.X_q, self.J_q = capi.geometric_factors(geom, int(prob["mesh_order"]), p + 1, self.G_mesh, self.nodes)
        name = f"{prob['kind']}_{KIND_TO_NAME[geom]}_p{p}"
        self.integral = capi.Integral(name, prob["params"], self.X_q, self.J_q, int(prob["ne"]))
        self.fields = [t(prob[f"field{i}"]) for i in range(len(vdims))]
        self.dfields = [t(prob[f"dfield{i}"]) for i in range(len(vdims))]
        self.F = capi.Functional(self.G_test, self.G_trials)
        self.F.AddDomainIntegral(self.integral, list(range(len(vdims))))
        if self.integral.table.state_bytes:
            self.qdata = t(prob["qdata0"]).clone()
            self.integral.set_qdata(self.qdata)
