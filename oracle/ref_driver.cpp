// TEST INFRASTRUCTURE ONLY (oracle/): C API of the reference oracle (see ref_driver.hpp).
#include "ref_driver.hpp"

#include <cstring>
#include <fstream>

namespace ref {
std::map<std::string, Factory>& factories()
{
  static std::map<std::string, Factory> f;
  return f;
}

struct Space {
  std::unique_ptr<mfem::FiniteElementCollection> fec;
  std::unique_ptr<mfem::FiniteElementSpace>  fes;
  std::unique_ptr<serac::BlockElementRestriction> G;
};
}  // namespace ref

using namespace ref;
std::vector<std::vector<int> > lexicographic_permutations(int p);  // element_restriction.cpp:7

namespace ref {
static const std::vector<char>* g_domain_mask = nullptr;

serac::Domain domain_for(mfem::Mesh& mesh)
{
  if (!g_domain_mask) return serac::EntireDomain(mesh);
  const std::vector<char>& mask = *g_domain_mask;
  size_t                   i    = 0;  // the predicate is called once per element, in element order (domain.cpp:249-256)
  if (mesh.SpaceDimension() == 2)
    return serac::Domain::ofElements(mesh, std::function<bool(std::vector<serac::vec2>, int)>(
                                               [&mask, i](std::vector<serac::vec2>, int) mutable { return mask[i++] != 0; }));
  return serac::Domain::ofElements(mesh, std::function<bool(std::vector<serac::vec3>, int)>(
                                             [&mask, i](std::vector<serac::vec3>, int) mutable { return mask[i++] != 0; }));
}
}  // namespace ref
using ref::g_domain_mask;

extern "C" {

// ---- mesh ----
void* ref_mesh_from_file(const char* path, int refinements)
{
  std::ifstream in(path);
  if (!in) return nullptr;
  auto* mesh = new mfem::Mesh(in, 1, 1);
  for (int i = 0; i < refinements; i++) mesh->UniformRefinement();
  mesh->EnsureNodes();  // Functional demands nodes (functional.hpp:74-89)
  return mesh;
}
void* ref_mesh_cartesian3d(int nx, int ny, int nz, double sx, double sy, double sz)
{
  auto* mesh = new mfem::Mesh(mfem::Mesh::MakeCartesian3D(nx, ny, nz, mfem::Element::HEXAHEDRON, sx, sy, sz));
  mesh->EnsureNodes();
  return mesh;
}
void* ref_mesh_cartesian2d(int nx, int ny, double sx, double sy)
{
  auto* mesh = new mfem::Mesh(mfem::Mesh::MakeCartesian2D(nx, ny, mfem::Element::QUADRILATERAL, true, sx, sy));
  mesh->EnsureNodes();
  return mesh;
}
/// simplex meshes: element type TETRAHEDRON / TRIANGLE instead of the tensor-product one
void* ref_mesh_cartesian3d_tet(int nx, int ny, int nz, double sx, double sy, double sz)
{
  auto* mesh = new mfem::Mesh(mfem::Mesh::MakeCartesian3D(nx, ny, nz, mfem::Element::TETRAHEDRON, sx, sy, sz));
  mesh->EnsureNodes();
  return mesh;
}
void* ref_mesh_cartesian2d_tri(int nx, int ny, double sx, double sy)
{
  auto* mesh = new mfem::Mesh(mfem::Mesh::MakeCartesian2D(nx, ny, mfem::Element::TRIANGLE, true, sx, sy));
  mesh->EnsureNodes();
  return mesh;
}
/// geometry of the mesh's elements (single-geometry meshes only: returns -1 when mixed)
int ref_mesh_geom(void* m)
{
  auto* mesh = static_cast<mfem::Mesh*>(m);
  if (mesh->GetNE() == 0) return -1;
  const auto g = mesh->GetElementGeometry(0);
  for (int e = 1; e < mesh->GetNE(); e++)
    if (mesh->GetElementGeometry(e) != g) return -1;
  return int(g);
}
void ref_mesh_destroy(void* m) { delete static_cast<mfem::Mesh*>(m); }
int  ref_mesh_dim(void* m) { return static_cast<mfem::Mesh*>(m)->Dimension(); }
int  ref_mesh_num_elements(void* m) { return static_cast<mfem::Mesh*>(m)->GetNE(); }
/// perturb the mesh nodes so that elements are not affine (x += amp * sin-ish of position; deterministic)
void ref_mesh_perturb(void* m, double amp)
{
  auto*        mesh  = static_cast<mfem::Mesh*>(m);
  mfem::Vector& nodes = *mesh->GetNodes();
  const int     dim   = mesh->Dimension();
  const int     nn    = nodes.Size() / dim;
  auto*         fes   = mesh->GetNodes()->FESpace();
  for (int i = 0; i < nn; i++) {
    double x[3] = {0, 0, 0};
    for (int d = 0; d < dim; d++) x[d] = nodes[fes->DofToVDof(i, d)];
    for (int d = 0; d < dim; d++) {
      nodes[fes->DofToVDof(i, d)] += amp * std::sin(3.0 * x[(d + 1) % dim] + 1.7 * x[d] + 0.3 * d);
    }
  }
}
int ref_mesh_nodes_order(void* m)
{
  return static_cast<mfem::Mesh*>(m)->GetNodes()->FESpace()->GetElementOrder(0);
}
int ref_mesh_nodes_ordering(void* m) { return int(static_cast<mfem::Mesh*>(m)->GetNodes()->FESpace()->GetOrdering()); }
long ref_mesh_nodes_size(void* m) { return static_cast<mfem::Mesh*>(m)->GetNodes()->Size(); }
void ref_mesh_nodes(void* m, double* out)
{
  auto* nodes = static_cast<mfem::Mesh*>(m)->GetNodes();
  std::memcpy(out, nodes->GetData(), sizeof(double) * size_t(nodes->Size()));
}
/// restriction table of the mesh-nodes space (what GeometricFactors gathers with, geometric_factors.cpp:67-69)
long ref_mesh_nodes_dof_info(void* m, int geom, uint64_t* out)
{
  auto* fes = static_cast<mfem::Mesh*>(m)->GetNodes()->FESpace();
  serac::ElementRestriction r(fes, mfem::Geometry::Type(geom));
  const long n = long(r.num_elements * r.nodes_per_elem);
  if (out) std::memcpy(out, r.dof_info.data(), sizeof(uint64_t) * size_t(n));
  return n;
}

// ---- H1 space + ElementRestriction ----
void* ref_space_create(void* m, int p, int vdim, int ordering)
{
  auto* mesh = static_cast<mfem::Mesh*>(m);
  auto* s    = new Space;
  s->fec     = std::make_unique<mfem::H1_FECollection>(p, mesh->Dimension());
  s->fes = std::make_unique<mfem::FiniteElementSpace>(mesh, s->fec.get(), vdim, mfem::Ordering::Type(ordering));
  s->G   = std::make_unique<serac::BlockElementRestriction>(s->fes.get());
  return s;
}
/// discontinuous space: L2 with the Gauss-Lobatto nodal basis (what serac's L2<p,c> elements interpolate,
/// F/detail/hexahedron_L2.inl:90-130); mfem returns its dofs in lexicographic order already (element_restriction.cpp:262-268)
void* ref_space_create_l2(void* m, int p, int vdim, int ordering)
{
  auto* mesh = static_cast<mfem::Mesh*>(m);
  auto* s    = new Space;
  s->fec     = std::make_unique<mfem::L2_FECollection>(p, mesh->Dimension(), mfem::BasisType::GaussLobatto);
  s->fes = std::make_unique<mfem::FiniteElementSpace>(mesh, s->fec.get(), vdim, mfem::Ordering::Type(ordering));
  s->G   = std::make_unique<serac::BlockElementRestriction>(s->fes.get());
  return s;
}
void ref_space_destroy(void* s) { delete static_cast<Space*>(s); }
long ref_space_vsize(void* s) { return static_cast<Space*>(s)->fes->GetVSize(); }
long ref_space_num_nodes(void* s) { return static_cast<Space*>(s)->fes->GetNDofs(); }
long ref_space_esize(void* s, int geom)
{
  auto& R = static_cast<Space*>(s)->G->restrictions;
  auto  it = R.find(mfem::Geometry::Type(geom));
  return it == R.end() ? 0 : long(it->second.ESize());
}
/// raw mfem element dofs in native order: the INPUT the product's lexicographic permutation is applied to
long ref_space_native_dofs(void* s, int geom, int* out)
{
  auto* fes  = static_cast<Space*>(s)->fes.get();
  auto* mesh = fes->GetMesh();
  long  n    = 0;
  for (int e = 0; e < fes->GetNE(); e++) {
    if (mesh->GetElementGeometry(e) != mfem::Geometry::Type(geom)) continue;
    mfem::Array<int> dofs;
    fes->GetElementDofs(e, dofs);
    for (int k = 0; k < dofs.Size(); k++) {
      if (out) out[n] = dofs[k];
      n++;
    }
  }
  return n;
}
/// the reference's dof_info table (element_restriction.cpp:216-277), packed 64-bit DoF words
long ref_space_dof_info(void* s, int geom, uint64_t* out)
{
  auto& r = static_cast<Space*>(s)->G->restrictions.at(mfem::Geometry::Type(geom));
  const long n = long(r.num_elements * r.nodes_per_elem);
  if (out) std::memcpy(out, r.dof_info.data(), sizeof(uint64_t) * size_t(n));
  return n;
}
void ref_space_gather(void* s, int geom, const double* L, double* E)
{
  auto&        r = static_cast<Space*>(s)->G->restrictions.at(mfem::Geometry::Type(geom));
  mfem::Vector Lv(const_cast<double*>(L), int(r.LSize())), Ev(E, int(r.ESize()));
  r.Gather(Lv, Ev);
}
void ref_space_scatter_add(void* s, int geom, const double* E, double* L)
{
  auto&        r = static_cast<Space*>(s)->G->restrictions.at(mfem::Geometry::Type(geom));
  mfem::Vector Lv(L, int(r.LSize())), Ev(const_cast<double*>(E), int(r.ESize()));
  r.ScatterAdd(Ev, Lv);
}
/// mfem's native -> lexicographic permutation (element_restriction.cpp:7-65)
int ref_lex_perm(int geom, int p, int* out)
{
  auto perms = lexicographic_permutations(p);
  auto& v    = perms[size_t(geom)];
  for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
  return int(v.size());
}

struct Csr {
  serac::GradientAssemblyLookupTables lut;
};

// ---- boundary faces: ElementRestriction(fes, face_geom, FaceType::BOUNDARY) (element_restriction.cpp:279-388,406-419)
/// raw mfem face dofs (fes->GetFaceDofs) of the boundary faces of one geometry, in face order: the INPUT the product's
/// lexicographic permutation is applied to
long ref_space_face_native_dofs(void* s, int face_geom, int* out)
{
  auto* fes  = static_cast<Space*>(s)->fes.get();
  auto* mesh = fes->GetMesh();
  long  n    = 0;
  for (int f = 0; f < fes->GetNF(); f++) {
    if (mesh->GetFaceGeometry(f) != mfem::Geometry::Type(face_geom)) continue;
    if (mesh->GetFaceInformation(f).IsInterior()) continue;
    mfem::Array<int> dofs;
    fes->GetFaceDofs(f, dofs);
    for (int k = 0; k < dofs.Size(); k++) {
      if (out) out[n] = dofs[k];
      n++;
    }
  }
  return n;
}
long ref_space_face_dof_info(void* s, int face_geom, uint64_t* out)
{
  serac::ElementRestriction r(static_cast<Space*>(s)->fes.get(), mfem::Geometry::Type(face_geom),
                              FaceType::BOUNDARY);
  const long n = long(r.num_elements * r.nodes_per_elem);
  if (out) std::memcpy(out, r.dof_info.data(), sizeof(uint64_t) * size_t(n));
  return n;
}
/// serac::GetFaceDofs(fes, face_geom, type) (element_restriction.cpp:279-388) for either face type; type: 0 = BOUNDARY,
/// 1 = INTERIOR (interior faces of a DG space carry the dofs of both adjacent elements, one after the other)
long ref_space_face_dof_info_typed(void* s, int face_geom, int type, uint64_t* out)
{
  serac::ElementRestriction r(static_cast<Space*>(s)->fes.get(), mfem::Geometry::Type(face_geom),
                              type == 0 ? FaceType::BOUNDARY : FaceType::INTERIOR);
  const long n = long(r.num_elements * r.nodes_per_elem);
  if (out && n) std::memcpy(out, r.dof_info.data(), sizeof(uint64_t) * size_t(n));
  return n;
}
/// what mfem knows about the (face, adjacent element) pairs of a DG space, in the order the reference walks them
/// (element_restriction.cpp:296-355): the INPUT of the product's sb_l2_face_dof_info.  Returns the number of pairs;
/// elem_dofs (flat) is indexed through offsets [pairs + 1].  The reference's convention for the orientation of the
/// boundary edges of 2D meshes (0, :340-346) is applied here, as the adapter of the product does.
long ref_space_l2_face_inputs(void* s, int face_geom, int type, int* elem_geom, int* elem_dofs, long* offsets, int* side,
                              int* orientation)
{
  auto*        fes          = static_cast<Space*>(s)->fes.get();
  mfem::Mesh*  mesh         = fes->GetMesh();
  mfem::Table* face_to_elem = mesh->GetFaceToElementTable();
  long         n = 0, nd = 0;
  for (int f = 0; f < fes->GetNF(); f++) {
    auto faceinfo = mesh->GetFaceInformation(f);
    if (mesh->GetFaceGeometry(f) != mfem::Geometry::Type(face_geom)) continue;
    if (faceinfo.IsInterior() && type == 0) continue;
    if (faceinfo.IsBoundary() && type == 1) continue;
    mfem::Array<int> elem_ids;
    face_to_elem->GetRow(f, elem_ids);
    for (auto elem : elem_ids) {
      mfem::Array<int> sides, ori;
      if (mesh->Dimension() == 2) {
        mesh->GetElementEdges(elem, sides, ori);
        for (auto& o : ori) o = (o == -1) ? 1 : 0;
      } else {
        mesh->GetElementFaces(elem, sides, ori);
      }
      int i;
      for (i = 0; i < sides.Size(); i++)
        if (sides[i] == f) break;
      mfem::Array<int> dofs;
      fes->GetElementDofs(elem, dofs);
      if (elem_geom) {
        elem_geom[n]   = int(mesh->GetElementGeometry(elem));
        side[n]        = i;
        orientation[n] = (mesh->Dimension() == 2 && type == 0) ? 0 : ori[i];
        offsets[n]     = nd;
        for (int k = 0; k < dofs.Size(); k++) elem_dofs[nd + k] = dofs[k];
      }
      nd += dofs.Size();
      n++;
      if (type == 0) break;
    }
  }
  if (offsets) offsets[n] = nd;
  delete face_to_elem;
  return n;
}
long ref_space_l2_face_inputs_size(void* s, int face_geom, int type)
{
  auto*        fes          = static_cast<Space*>(s)->fes.get();
  mfem::Mesh*  mesh         = fes->GetMesh();
  mfem::Table* face_to_elem = mesh->GetFaceToElementTable();
  long         nd = 0;
  for (int f = 0; f < fes->GetNF(); f++) {
    auto faceinfo = mesh->GetFaceInformation(f);
    if (mesh->GetFaceGeometry(f) != mfem::Geometry::Type(face_geom)) continue;
    if (faceinfo.IsInterior() && type == 0) continue;
    if (faceinfo.IsBoundary() && type == 1) continue;
    mfem::Array<int> elem_ids;
    face_to_elem->GetRow(f, elem_ids);
    for (auto elem : elem_ids) {
      mfem::Array<int> dofs;
      fes->GetElementDofs(elem, dofs);
      nd += dofs.Size();
      if (type == 0) break;
    }
  }
  delete face_to_elem;
  return nd;
}
long ref_space_face_esize(void* s, int face_geom)
{
  serac::ElementRestriction r(static_cast<Space*>(s)->fes.get(), mfem::Geometry::Type(face_geom),
                              FaceType::BOUNDARY);
  return long(r.ESize());
}
void ref_space_face_gather(void* s, int face_geom, const double* L, double* E)
{
  serac::ElementRestriction r(static_cast<Space*>(s)->fes.get(), mfem::Geometry::Type(face_geom),
                              FaceType::BOUNDARY);
  mfem::Vector Lv(const_cast<double*>(L), int(r.LSize())), Ev(E, int(r.ESize()));
  r.Gather(Lv, Ev);
}
void ref_space_face_scatter_add(void* s, int face_geom, const double* E, double* L)
{
  serac::ElementRestriction r(static_cast<Space*>(s)->fes.get(), mfem::Geometry::Type(face_geom),
                              FaceType::BOUNDARY);
  mfem::Vector Lv(L, int(r.LSize())), Ev(const_cast<double*>(E), int(r.ESize()));
  r.ScatterAdd(Ev, Lv);
}
long ref_mesh_nodes_face_dof_info(void* m, int face_geom, uint64_t* out)
{
  auto* fes = static_cast<mfem::Mesh*>(m)->GetNodes()->FESpace();
  serac::ElementRestriction r(fes, mfem::Geometry::Type(face_geom), FaceType::BOUNDARY);
  const long n = long(r.num_elements * r.nodes_per_elem);
  if (out) std::memcpy(out, r.dof_info.data(), sizeof(uint64_t) * size_t(n));
  return n;
}
long ref_mesh_nodes_face_native_dofs(void* m, int face_geom, int* out)
{
  auto* mesh = static_cast<mfem::Mesh*>(m);
  auto* fes  = mesh->GetNodes()->FESpace();
  long  n    = 0;
  for (int f = 0; f < fes->GetNF(); f++) {
    if (mesh->GetFaceGeometry(f) != mfem::Geometry::Type(face_geom)) continue;
    if (mesh->GetFaceInformation(f).IsInterior()) continue;
    mfem::Array<int> dofs;
    fes->GetFaceDofs(f, dofs);
    for (int k = 0; k < dofs.Size(); k++) {
      if (out) out[n] = dofs[k];
      n++;
    }
  }
  return n;
}
/// GeometricFactors(domain, q, face_geom, FaceType::BOUNDARY) on EntireBoundary(mesh) (geometric_factors.cpp:140-220)
long ref_boundary_geometric_factors(void* m, int q, int face_geom, double* X, double* J)
{
  auto*                   mesh = static_cast<mfem::Mesh*>(m);
  serac::Domain           d    = serac::EntireBoundary(*mesh);
  serac::GeometricFactors gf(d, q, mfem::Geometry::Type(face_geom), FaceType::BOUNDARY);
  if (X) std::memcpy(X, gf.X.GetData(), sizeof(double) * size_t(gf.X.Size()));
  if (J) std::memcpy(J, gf.J.GetData(), sizeof(double) * size_t(gf.J.Size()));
  return long(gf.num_elements);
}
/// scatter of boundary-face element matrices into CSR values as Gradient::assemble does for
/// Domain::Type::BoundaryElements (functional.hpp:610-644); accumulates into `values`
void ref_csr_assemble_faces(void* c, void* test_space, void* trial_space, int face_geom, const double* K_e,
                            double* values)
{
  auto& lut = static_cast<Csr*>(c)->lut;
  serac::ElementRestriction test(static_cast<Space*>(test_space)->fes.get(), mfem::Geometry::Type(face_geom),
                                 FaceType::BOUNDARY);
  serac::ElementRestriction trial(static_cast<Space*>(trial_space)->fes.get(), mfem::Geometry::Type(face_geom),
                                  FaceType::BOUNDARY);
  std::vector<DoF> test_vdofs(test.nodes_per_elem * test.components);
  std::vector<DoF> trial_vdofs(trial.nodes_per_elem * trial.components);
  const size_t nrow = test_vdofs.size(), ncol = trial_vdofs.size();
  for (uint64_t e = 0; e < test.num_elements; e++) {
    test.GetElementVDofs(int(e), test_vdofs);
    trial.GetElementVDofs(int(e), trial_vdofs);
    for (size_t i = 0; i < ncol; i++) {
      int col = int(trial_vdofs[i].index());
      for (size_t j = 0; j < nrow; j++) {
        int row  = int(test_vdofs[j].index());
        int sign = test_vdofs[j].sign() * trial_vdofs[i].sign();
        values[lut(row, col)] += sign * K_e[(e * ncol + i) * nrow + j];
      }
    }
  }
}

// ---- GeometricFactors (geometric_factors.cpp:62-138) ----
void ref_geometric_factors(void* m, int q, int geom, double* X, double* J)
{
  auto*                  mesh = static_cast<mfem::Mesh*>(m);
  serac::Domain          d    = serac::EntireDomain(*mesh);
  serac::GeometricFactors gf(d, q, mfem::Geometry::Type(geom));
  std::memcpy(X, gf.X.GetData(), sizeof(double) * size_t(gf.X.Size()));
  std::memcpy(J, gf.J.GetData(), sizeof(double) * size_t(gf.J.Size()));
}

// ---- sparsity pattern (dof_numbering.hpp:290-351) ----
void* ref_csr_create(void* test_space, void* trial_space)
{
  auto* c = new Csr;
  c->lut.init(*static_cast<Space*>(test_space)->G, *static_cast<Space*>(trial_space)->G);
  return c;
}
void ref_csr_destroy(void* c) { delete static_cast<Csr*>(c); }
long ref_csr_nnz(void* c) { return long(static_cast<Csr*>(c)->lut.nnz); }
long ref_csr_num_rows(void* c) { return long(static_cast<Csr*>(c)->lut.row_ptr.size()) - 1; }
void ref_csr_pattern(void* c, int* row_ptr, int* col_ind)
{
  auto& l = static_cast<Csr*>(c)->lut;
  std::memcpy(row_ptr, l.row_ptr.data(), sizeof(int) * l.row_ptr.size());
  std::memcpy(col_ind, l.col_ind.data(), sizeof(int) * l.col_ind.size());
}
/// scatter of element matrices into CSR values exactly as Gradient::assemble does (functional.hpp:610-644),
/// for one geometry; values must be zero-initialised by the caller (functional.hpp:586)
void ref_csr_assemble(void* c, void* test_space, void* trial_space, int geom, const double* K_e, double* values)
{
  auto& lut   = static_cast<Csr*>(c)->lut;
  auto& test  = static_cast<Space*>(test_space)->G->restrictions.at(mfem::Geometry::Type(geom));
  auto& trial = static_cast<Space*>(trial_space)->G->restrictions.at(mfem::Geometry::Type(geom));
  std::vector<DoF> test_vdofs(test.nodes_per_elem * test.components);
  std::vector<DoF> trial_vdofs(trial.nodes_per_elem * trial.components);
  const size_t nrow = test_vdofs.size(), ncol = trial_vdofs.size();
  for (uint64_t e = 0; e < test.num_elements; e++) {
    test.GetElementVDofs(int(e), test_vdofs);
    trial.GetElementVDofs(int(e), trial_vdofs);
    for (size_t i = 0; i < ncol; i++) {
      int col = int(trial_vdofs[i].index());
      for (size_t j = 0; j < nrow; j++) {
        int row  = int(test_vdofs[j].index());
        int sign = test_vdofs[j].sign() * trial_vdofs[i].sign();
        values[lut(row, col)] += sign * K_e[(e * ncol + i) * nrow + j];
      }
    }
  }
}

// ---- Integral ----
int         ref_num_kinds() { return int(factories().size()); }
const char* ref_kind_name(int i)
{
  auto it = factories().begin();
  std::advance(it, i);
  return it->first.c_str();
}
/// same, on the sub-domain of the elements with mask[e] != 0 (element order of the mesh)
void* ref_integral_create_subset(const char* kind, void* mesh, const double* params, int nparams,
                                 const unsigned char* mask, int nmask)
{
  auto it = factories().find(kind);
  if (it == factories().end()) return nullptr;
  std::vector<char> m(mask, mask + nmask);
  g_domain_mask = &m;
  auto* h       = new IntegralHolder(it->second(*static_cast<mfem::Mesh*>(mesh), params, nparams));
  g_domain_mask = nullptr;
  return h;
}
void* ref_integral_create(const char* kind, void* mesh, const double* params, int nparams)
{
  auto it = factories().find(kind);
  if (it == factories().end()) return nullptr;
  auto* h = new IntegralHolder(it->second(*static_cast<mfem::Mesh*>(mesh), params, nparams));
  return h;
}
void   ref_integral_destroy(void* h) { delete static_cast<IntegralHolder*>(h); }
size_t ref_integral_state_bytes(void* h) { return static_cast<IntegralHolder*>(h)->state_bytes; }
size_t ref_integral_qdata_count(void* h) { return static_cast<IntegralHolder*>(h)->qdata_count; }
void   ref_integral_get_qdata(void* h, void* out)
{
  auto* H = static_cast<IntegralHolder*>(h);
  std::memcpy(out, H->qdata_raw, H->state_bytes * H->qdata_count);
}
void ref_integral_set_qdata(void* h, const void* in)
{
  auto* H = static_cast<IntegralHolder*>(h);
  std::memcpy(H->qdata_raw, in, H->state_bytes * H->qdata_count);
}
/// Integral::Mult for one geometry (integral.hpp:65-81): wrt < 0 => plain evaluation, else evaluation_with_AD_[wrt].
/// out_E must be zeroed by the caller (integral.hpp:68).
int ref_integral_mult(void* h, int geom, double t, const double* const* inputs_E, int num_inputs, double* out_E,
                      int wrt, int update_state)
{
  auto&                      I = *static_cast<IntegralHolder*>(h)->integral;
  std::vector<const double*> in(inputs_E, inputs_E + num_inputs);
  auto& kernels = (wrt >= 0) ? I.evaluation_with_AD_[size_t(wrt)] : I.evaluation_;
  auto  it      = kernels.find(mfem::Geometry::Type(geom));
  if (it == kernels.end()) return 1;
  it->second(t, in, out_E, update_state != 0);
  return 0;
}
/// Integral::GradientMult (integral.hpp:93-103)
int ref_integral_jvp(void* h, int geom, int wrt, const double* du_E, double* dr_E)
{
  auto& I  = *static_cast<IntegralHolder*>(h)->integral;
  auto  it = I.jvp_[size_t(wrt)].find(mfem::Geometry::Type(geom));
  if (it == I.jvp_[size_t(wrt)].end()) return 1;
  it->second(du_E, dr_E);
  return 0;
}
/// Integral::ComputeElementGradients (integral.hpp:112-121); K_e [ne][ncol][nrow] is accumulated into
int ref_integral_element_gradient(void* h, int geom, int wrt, double* K_e, long ne, long ncol, long nrow)
{
  auto& I  = *static_cast<IntegralHolder*>(h)->integral;
  auto  it = I.element_gradient_[size_t(wrt)].find(mfem::Geometry::Type(geom));
  if (it == I.element_gradient_[size_t(wrt)].end()) return 1;
  serac::ExecArray<double, 3, serac::ExecutionSpace::CPU> K(ne, ncol, nrow);
  std::memcpy(K.data(), K_e, sizeof(double) * size_t(ne * ncol * nrow));
  it->second(serac::view(K));
  std::memcpy(K_e, K.data(), sizeof(double) * size_t(ne * ncol * nrow));
  return 0;
}

}  // extern "C"
