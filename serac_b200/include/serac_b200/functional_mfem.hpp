// serac_b200: `serac::Functional<test(trials...)>` on mfem's own types -- the drop-in form of the boundary.
//
// Same public surface as the reference class (src/serac/numerics/functional/functional.hpp:189-745):
//   ctor(const mfem::FiniteElementSpace* test, std::array<const mfem::FiniteElementSpace*, N> trials)        :219-281
//   AddDomainIntegral(Dimension<dim>, DependsOn<...>, integrand, mfem::Mesh& | Domain&, qdata = NoQData)     :286-317
//   AddBoundaryIntegral(Dimension<dim>, DependsOn<...>, integrand, mfem::Mesh&)                              :328-354
//   AddAreaIntegral / AddVolumeIntegral / AddSurfaceIntegral                                                 :365-393
//   operator()(double t, const mfem::Vector&... [one of them differentiate_wrt(v)])                          :446-513
//       -> mfem::Vector&  or  tuple<mfem::Vector&, Gradient&>
//   Gradient : mfem::Operator  { Mult, operator()(dx), assemble(Gradient&) }                                 :532-709
//   updateQdata(bool)                                                                                        :523
// and ShapeAwareFunctional<shape, test(trials...)> (shape_aware_functional.hpp:292-575).
//
// What differs from the reference, and why:
//   * the spaces are mfem::FiniteElementSpace, not ParFiniteElementSpace, and assemble() returns mfem::SparseMatrix, not
//     HypreParMatrix: this image has neither MPI nor hypre, so the header is compiled and tested against the serial mfem
//     of oracle/_ref/mfem-install.  At one rank P = identity (mfem/fem/pfespace.cpp:1170-1173) and the reference's
//     HypreParMatrix is the block-diagonal wrapper of exactly this SparseMatrix (functional.hpp:649-662); with
//     MFEM_USE_MPI the two typedefs below switch to the Par types and INTEGRATION.md shows the three extra lines;
//   * q-functions are functor structs (nvcc has no generic extended lambdas), as in serac's own physics modules;
//   * mfem::Vector data is host memory in a serial mfem build: arguments and results cross PCIe inside every call
//     (SB_HOST_SYNC).  With an mfem built with CUDA, Vector::Read() / Write() hand out device pointers and the same calls
//     run with SB_DEVICE (MFEM_USE_CUDA branch below).
// Everything numerical happens behind the C ABI; this header only translates types.
#pragma once

#define SERAC_B200_MFEM_API 1

#include <functional>
#include <map>
#include <set>

#include "mfem.hpp"
#include "serac_b200/functional.hpp"

namespace serac {

using core_api::EmptyQData;
using core_api::NoQData;
using core_api::QuadratureData;

enum class ExecutionSpace
{
  CPU,
  GPU,
  Dynamic
};
constexpr ExecutionSpace default_execution_space = ExecutionSpace::GPU;  // there is no CPU path

/// functional.hpp:47-66
struct differentiate_wrt_this {
  const mfem::Vector& ref;
  operator const mfem::Vector&() const { return ref; }
};
inline differentiate_wrt_this differentiate_wrt(const mfem::Vector& v) { return differentiate_wrt_this{v}; }

/// the subset of serac::Domain (domain.hpp) the element kernels need: a list of elements -- or of boundary elements --
/// of one mesh.  Predicates take (vertex coordinates, attribute) as in the reference: std::vector<tensor<double, dim>>
/// (domain.hpp:95-127); std::vector<std::array<double, 3>> is accepted as well.
struct Domain {
  enum class Type
  {
    Elements,
    BoundaryElements
  };
  mfem::Mesh&      mesh_;
  std::vector<int> elements_;  ///< mesh element ids, or mesh face ids of boundary elements (ascending)
  Type             type_ = Type::Elements;

  template <typename Pred>
  static bool call(Pred& predicate, mfem::Mesh& mesh, const mfem::Array<int>& v, int attr)
  {
    const int sdim = mesh.SpaceDimension();
    if constexpr (std::is_invocable_r_v<bool, Pred, std::vector<tensor<double, 3>>, int>) {
      if (sdim == 3) {
        std::vector<tensor<double, 3>> x(static_cast<size_t>(v.Size()));
        for (int i = 0; i < v.Size(); i++)
          for (int c = 0; c < 3; c++) x[size_t(i)][c] = mesh.GetVertex(v[i])[c];
        return predicate(x, attr);
      }
    }
    if constexpr (std::is_invocable_r_v<bool, Pred, std::vector<tensor<double, 2>>, int>) {
      if (sdim == 2) {
        std::vector<tensor<double, 2>> x(static_cast<size_t>(v.Size()));
        for (int i = 0; i < v.Size(); i++)
          for (int c = 0; c < 2; c++) x[size_t(i)][c] = mesh.GetVertex(v[i])[c];
        return predicate(x, attr);
      }
    }
    if constexpr (std::is_invocable_r_v<bool, Pred, std::vector<std::array<double, 3>>, int>) {
      std::vector<std::array<double, 3>> x(static_cast<size_t>(v.Size()), std::array<double, 3>{0.0, 0.0, 0.0});
      for (int i = 0; i < v.Size(); i++)
        for (int c = 0; c < sdim; c++) x[size_t(i)][size_t(c)] = mesh.GetVertex(v[i])[c];
      return predicate(x, attr);
    }
    std::fprintf(stderr, "serac::Domain: the predicate does not take the vertex coordinates of a %dD mesh\n", sdim);
    std::abort();
  }

  /// Domain::ofElements(mesh, predicate(vertex coordinates, attribute)) (domain.hpp:95-110, domain.cpp:238-300)
  template <typename Pred>
  static Domain ofElements(mfem::Mesh& mesh, Pred&& predicate)
  {
    Domain d{mesh, {}, Type::Elements};
    for (int e = 0; e < mesh.GetNE(); e++) {
      mfem::Array<int> v;
      mesh.GetElementVertices(e, v);
      if (call(predicate, mesh, v, mesh.GetAttribute(e))) d.elements_.push_back(e);
    }
    return d;
  }

  /// Domain::ofBoundaryElements (domain.hpp:112-127, domain.cpp:311-386): the boundary FACES of the mesh in face order;
  /// attribute = that of the boundary element on the face, -1 if there is none
  template <typename Pred>
  static Domain ofBoundaryElements(mfem::Mesh& mesh, Pred&& predicate)
  {
    Domain           d{mesh, {}, Type::BoundaryElements};
    mfem::Array<int> face_to_bdr = mesh.GetFaceToBdrElMap();
    for (int f = 0; f < mesh.GetNumFaces(); f++) {
      if (mesh.GetFaceInformation(f).IsInterior()) continue;
      mfem::Array<int> v;
      mesh.GetFaceVertices(f, v);
      const int bdr_id = face_to_bdr[f];
      if (call(predicate, mesh, v, bdr_id >= 0 ? mesh.GetBdrAttribute(bdr_id) : -1)) d.elements_.push_back(f);
    }
    return d;
  }
};
/// predicates on the attribute (domain.hpp:190-214)
template <int dim>
inline auto by_attr(int value)
{
  return [value](std::vector<tensor<double, dim>>, int attr) { return attr == value; };
}
template <int dim>
inline auto by_attr(std::set<int> values)
{
  return [values](std::vector<tensor<double, dim>>, int attr) { return values.find(attr) != values.end(); };
}
inline Domain EntireDomain(mfem::Mesh& mesh)
{
  return Domain::ofElements(mesh, [](std::vector<std::array<double, 3>>, int) { return true; });
}
inline Domain EntireBoundary(mfem::Mesh& mesh)
{
  return Domain::ofBoundaryElements(mesh, [](std::vector<std::array<double, 3>>, int) { return true; });
}

namespace mfem_detail {

inline Geometry geometry_of(mfem::Geometry::Type g)
{
  switch (g) {
    case mfem::Geometry::SEGMENT: return Geometry::SEGMENT;
    case mfem::Geometry::TRIANGLE: return Geometry::TRIANGLE;
    case mfem::Geometry::SQUARE: return Geometry::SQUARE;
    case mfem::Geometry::TETRAHEDRON: return Geometry::TETRAHEDRON;
    case mfem::Geometry::CUBE: return Geometry::CUBE;
    default: break;
  }
  std::fprintf(stderr, "serac::Functional: unsupported element geometry %d\n", int(g));
  std::abort();
}

/// what Functional reads from an mfem space (functional.hpp:219-281, element_restriction.cpp:216-277,279-388): order,
/// vdim, ordering, GetNDofs and the native element dofs; boundary faces in face-index order, as the reference's
/// FaceType::BOUNDARY restriction walks them
inline core_api::FiniteElementSpace convert_space(const mfem::FiniteElementSpace& fes)
{
  mfem::Mesh& mesh = *fes.GetMesh();
  if (mesh.GetNE() == 0) {
    std::fprintf(stderr, "serac::Functional: empty mesh\n");
    std::abort();
  }
  // the geometries of the mesh in mfem::Geometry order, as BlockElementRestriction walks them
  // (element_restriction.cpp:471-536); the first one fills the space's own fields, the others its `more_elements`
  std::vector<mfem::Geometry::Type> geoms;
  for (int e = 0; e < mesh.GetNE(); e++) {
    const auto g = mesh.GetElementGeometry(e);
    if (std::find(geoms.begin(), geoms.end(), g) == geoms.end()) geoms.push_back(g);
  }
  std::sort(geoms.begin(), geoms.end());
  const auto g0 = geoms[0];
  core_api::FiniteElementSpace out{};
  out.geometry     = geometry_of(g0);
  out.family       = dynamic_cast<const mfem::L2_FECollection*>(fes.FEColl()) ? Family::L2 : Family::H1;
  out.order        = fes.GetMaxElementOrder();
  out.vdim         = fes.GetVDim();
  out.ordering     = (fes.GetOrdering() == mfem::Ordering::byVDIM) ? core_api::Ordering::byVDIM : core_api::Ordering::byNODES;
  out.num_nodes    = fes.GetNDofs();
  out.num_elements = 0;
  for (size_t b = 1; b < geoms.size(); b++) out.more_elements.push_back(core_api::GeometryBlock{geometry_of(geoms[b]), {}, {}});
  mfem::Array<int> dofs;
  for (int e = 0; e < mesh.GetNE(); e++) {
    fes.GetElementDofs(e, dofs);
    const size_t      b   = size_t(std::find(geoms.begin(), geoms.end(), mesh.GetElementGeometry(e)) - geoms.begin());
    std::vector<int>& dst = (b == 0) ? out.native_element_dofs : out.more_elements[b - 1].native_dofs;
    for (int i = 0; i < dofs.Size(); i++) dst.push_back(dofs[i] >= 0 ? dofs[i] : -1 - dofs[i]);
    if (b == 0) {
      out.num_elements++;
      if (geoms.size() > 1) out.element_ids.push_back(e);
    } else {
      out.more_elements[b - 1].ids.push_back(e);
    }
  }
  // boundary faces, one block per face geometry (FaceType::BOUNDARY restrictions, element_restriction.cpp:279-388)
  std::vector<mfem::Geometry::Type> fgeoms;
  for (int f = 0; f < mesh.GetNumFaces(); f++) {
    if (!mesh.GetFaceInformation(f).IsBoundary()) continue;
    const auto g = mesh.GetFaceGeometry(f);
    if (std::find(fgeoms.begin(), fgeoms.end(), g) == fgeoms.end()) fgeoms.push_back(g);
  }
  std::sort(fgeoms.begin(), fgeoms.end());
  if (!fgeoms.empty()) out.boundary_face_geometry = geometry_of(fgeoms[0]);
  for (size_t b = 1; b < fgeoms.size(); b++)
    out.more_boundary_faces.push_back(core_api::GeometryBlock{geometry_of(fgeoms[b]), {}, {}});
  for (int f = 0; f < mesh.GetNumFaces(); f++) {
    if (!mesh.GetFaceInformation(f).IsBoundary()) continue;
    const size_t      b   = size_t(std::find(fgeoms.begin(), fgeoms.end(), mesh.GetFaceGeometry(f)) - fgeoms.begin());
    std::vector<int>& dst = (b == 0) ? out.native_boundary_face_dofs : out.more_boundary_faces[b - 1].native_dofs;
    if (out.family == Family::L2) {
      // a DG space has no face dofs: they are those of the adjacent element that lie on the face
      // (element_restriction.cpp:304-355), already in the face's lexicographic order
      int e1 = -1, e2 = -1;
      mesh.GetFaceElements(f, &e1, &e2);
      mfem::Array<int> sides, ori, edofs;
      if (mesh.Dimension() == 2) {
        mesh.GetElementEdges(e1, sides, ori);
        for (auto& o : ori) o = 0;  // the reference discards mfem's orientation of boundary edges (:340-346)
      } else {
        mesh.GetElementFaces(e1, sides, ori);
      }
      int side = 0;
      while (side < sides.Size() && sides[side] != f) side++;
      fes.GetElementDofs(e1, edofs);
      const int     eg = int(mesh.GetElementGeometry(e1)), fg = int(mesh.GetFaceGeometry(f)), o = ori[side];
      const int64_t zero = 0;
      std::vector<uint64_t> info(size_t(b200::num_nodes(geometry_of(mesh.GetFaceGeometry(f)), out.order)));
      SERAC_B200_CHECK(sb_l2_face_dof_info(fg, out.order, 1, &eg, edofs.GetData(), &zero, &side, &o, info.data()));
      for (uint64_t d : info) dst.push_back(int(d));
    } else {
      fes.GetFaceDofs(f, dofs);
      for (int i = 0; i < dofs.Size(); i++) dst.push_back(dofs[i] >= 0 ? dofs[i] : -1 - dofs[i]);
    }
    if (b == 0) {
      out.num_boundary_faces++;
      out.boundary_face_ids.push_back(f);
    } else
      out.more_boundary_faces[b - 1].ids.push_back(f);
  }
  return out;
}

/// the nodal grid function of the mesh (the reference refuses meshes without nodes, functional.hpp:74-89; here they are
/// created, which is what serac's mesh utilities do before any Functional sees the mesh)
inline core_api::Mesh convert_mesh(mfem::Mesh& mesh)
{
  mesh.EnsureNodes();
  const mfem::GridFunction* nodes = mesh.GetNodes();
  core_api::Mesh            out{};
  out.nodes_space = convert_space(*nodes->FESpace());
  out.nodes.assign(nodes->GetData(), nodes->GetData() + nodes->Size());
  return out;
}

}  // namespace mfem_detail

template <typename T, ExecutionSpace exec = default_execution_space>
class Functional;

template <typename test, typename... trials, ExecutionSpace exec>
class Functional<test(trials...), exec> {
  static constexpr int num_trial_spaces = int(sizeof...(trials));
  using Impl                            = core_api::Functional<test(trials...)>;
#ifdef MFEM_USE_MPI
  using fes_t    = mfem::ParFiniteElementSpace;
  using matrix_t = mfem::HypreParMatrix;
#else
  using fes_t    = mfem::FiniteElementSpace;
  using matrix_t = mfem::SparseMatrix;
#endif

 public:
  class Gradient;

  /// functional.hpp:219-281
  Functional(const fes_t* test_fes, std::array<const fes_t*, num_trial_spaces> trial_fes)
      : test_fes_(test_fes), trial_fes_(trial_fes), test_vsize_(test_fes->GetVSize())
  {
    std::array<const core_api::FiniteElementSpace*, num_trial_spaces> t{};
    const core_api::FiniteElementSpace* ts = space_of(test_fes);
    for (int i = 0; i < num_trial_spaces; i++) t[size_t(i)] = space_of(trial_fes[size_t(i)]);
    impl_ = std::make_unique<Impl>(ts, t);
    output_T_.SetSize(test_vsize_);
    grad_.reserve(size_t(num_trial_spaces));
    for (int i = 0; i < num_trial_spaces; i++) grad_.emplace_back(*this, i);
  }

  /// quantities of interest (functional_qoi.inl:97-134): no test space; the "test space" has one dof (F/detail/qoi.inl)
  explicit Functional(std::array<const fes_t*, num_trial_spaces> trial_fes)
      : test_fes_(nullptr), trial_fes_(trial_fes), test_vsize_(1)
  {
    static_assert(test::family == Family::QOI, "only a quantity of interest is built without a test space");
    std::array<const core_api::FiniteElementSpace*, num_trial_spaces> t{};
    for (int i = 0; i < num_trial_spaces; i++) t[size_t(i)] = space_of(trial_fes[size_t(i)]);
    qoi_space_ = std::make_unique<core_api::FiniteElementSpace>(core_api::make_qoi_space(*t[0]));
    impl_      = std::make_unique<Impl>(qoi_space_.get(), t);
    output_T_.SetSize(test_vsize_);
    grad_.reserve(size_t(num_trial_spaces));
    for (int i = 0; i < num_trial_spaces; i++) grad_.emplace_back(*this, i);
  }

  /// functional.hpp:286-317 with a whole mesh as the domain
  template <int dim, int... args, typename Integrand, typename StateType = Nothing>
  void AddDomainIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, mfem::Mesh& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    if (domain.GetNE() == 0) return;  // functional.hpp:290
    core_api::Domain whole = core_api::EntireDomain(*mesh_of(domain));
    impl_->AddDomainIntegral(d, a, integrand, whole, qdata);
  }
  /// ... and with a Domain (functional.hpp:286-317, domain.hpp)
  template <int dim, int... args, typename Integrand, typename StateType = Nothing>
  void AddDomainIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, Domain& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    if (domain.type_ != Domain::Type::Elements) fail("AddDomainIntegral: the domain is a set of boundary elements");
    core_api::Domain part{mesh_of(domain.mesh_), false, domain.elements_};
    impl_->AddDomainIntegral(d, a, integrand, part, qdata);
  }

  /// functional.hpp:328-354
  template <int dim, int... args, typename Integrand>
  void AddBoundaryIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, mfem::Mesh& domain)
  {
    if (domain.GetNBE() == 0) return;
    impl_->AddBoundaryIntegral(d, a, integrand, *mesh_of(domain));
  }
  /// ... on part of the boundary (Domain::ofBoundaryElements)
  template <int dim, int... args, typename Integrand>
  void AddBoundaryIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, Domain& domain)
  {
    if (domain.type_ != Domain::Type::BoundaryElements) fail("AddBoundaryIntegral: the domain is not a set of boundary elements");
    core_api::Domain part{mesh_of(domain.mesh_), false, domain.elements_, true};
    impl_->AddBoundaryIntegral(d, a, integrand, part);
  }

  /// functional.hpp:365-393
  template <int... args, typename Integrand, typename StateType = Nothing>
  void AddAreaIntegral(DependsOn<args...> which_args, const Integrand& integrand, mfem::Mesh& domain,
                       std::shared_ptr<QuadratureData<StateType>> data = NoQData)
  {
    AddDomainIntegral(Dimension<2>{}, which_args, integrand, domain, data);
  }
  template <int... args, typename Integrand, typename StateType = Nothing>
  void AddVolumeIntegral(DependsOn<args...> which_args, const Integrand& integrand, mfem::Mesh& domain,
                         std::shared_ptr<QuadratureData<StateType>> data = NoQData)
  {
    AddDomainIntegral(Dimension<3>{}, which_args, integrand, domain, data);
  }
  template <int... args, typename Integrand>
  void AddSurfaceIntegral(DependsOn<args...> which_args, const Integrand& integrand, mfem::Mesh& domain)
  {
    AddBoundaryIntegral(Dimension<2>{}, which_args, integrand, domain);
  }

  /// functional.hpp:446-513.  At one rank P = identity: T-vectors are L-vectors (mfem/fem/pfespace.cpp:1170-1173)
  template <typename... T>
  decltype(auto) operator()(double t, const T&... args)
  {
    static_assert(sizeof...(T) == sizeof...(trials), "one argument per trial space");
    constexpr int num_differentiated = (int(std::is_same_v<T, differentiate_wrt_this>) + ...);
    static_assert(num_differentiated <= 1, "at most one argument can be differentiated (functional.hpp:504)");
    constexpr int       wrt  = index_of_differentiation<T...>();
    const mfem::Vector* in[] = {&static_cast<const mfem::Vector&>(args)...};
    const double*       ptrs[sizeof...(T)];
    for (size_t i = 0; i < sizeof...(T); i++) {
      if (in[i]->Size() != trial_fes_[i]->GetVSize()) fail("operator(): argument size does not match its space");
      ptrs[i] = read(*in[i]);
    }
    impl_->evaluate_raw(t, wrt, ptrs, write(output_T_), transfer_mode());
    if constexpr (wrt >= 0) {
      return std::tuple<mfem::Vector&, Gradient&>(output_T_, grad_[size_t(wrt)]);
    } else {
      return static_cast<mfem::Vector&>(output_T_);
    }
  }

  /// functional.hpp:523
  void updateQdata(bool update_flag) { impl_->updateQdata(update_flag); }

  Gradient& gradient_of(int which) { return grad_[size_t(which)]; }

  /// functional.hpp:532-709
  class Gradient : public mfem::Operator {
   public:
    Gradient(Functional& f, int which)
        : mfem::Operator(f.test_vsize_, f.trial_fes_[size_t(which)]->GetVSize()),
          form_(f),
          which_argument_(which),
          df_(f.test_vsize_)
    {
    }
    /// functional.hpp:556-559
    void Mult(const mfem::Vector& dx, mfem::Vector& df) const override
    {
      if (dx.Size() != Width() || df.Size() != Height()) form_.fail("Gradient::Mult: size mismatch");
      form_.impl_->jvp_raw(which_argument_, read(dx), write(df), transfer_mode());
    }
    /// functional.hpp:562
    mfem::Vector& operator()(const mfem::Vector& dx)
    {
      Mult(dx, df_);
      return df_;
    }
    /// functional.hpp:569-667: the caller owns the matrix; the CSR graph is a copy (hypre permutes column indices in
    /// place, :646-647), the values come straight from the device
    friend std::unique_ptr<matrix_t> assemble(Gradient& g) { return g.assemble_matrix(); }

    std::unique_ptr<matrix_t> assemble_matrix()
    {
      Gradient& g = *this;
      auto K   = g.form_.impl_->gradient(g.which_argument_).assemble_local();  // local CSR triple (core_api::CsrMatrix)
      int*    I = new int[K->row_ptr.size()];  // handed to the SparseMatrix, which frees them with delete[]
      int*    J = new int[K->col_ind.size() ? K->col_ind.size() : 1];
      double* A = new double[K->values.size() ? K->values.size() : 1];
      std::copy(K->row_ptr.begin(), K->row_ptr.end(), I);
      std::copy(K->col_ind.begin(), K->col_ind.end(), J);
      std::copy(K->values.begin(), K->values.end(), A);
      auto J_local = std::make_unique<mfem::SparseMatrix>(I, J, A, int(K->num_rows), int(K->num_cols), true, true, true);
#ifdef MFEM_USE_MPI
      // functional.hpp:654-662: block-diagonal HypreParMatrix of the local matrix, then R A P with the spaces'
      // prolongations (the identity at one rank)
      auto* test  = g.form_.test_fes_;
      auto* trial = g.form_.trial_fes_[size_t(g.which_argument_)];
      auto* A_hyp = new mfem::HypreParMatrix(test->GetComm(), test->GlobalVSize(), trial->GlobalVSize(),
                                             test->GetDofOffsets(), trial->GetDofOffsets(), J_local.get());
      auto* K_par = mfem::RAP(test->Dof_TrueDof_Matrix(), A_hyp, trial->Dof_TrueDof_Matrix());
      delete A_hyp;
      return std::unique_ptr<matrix_t>(K_par);
#else
      return J_local;
#endif
    }

   private:
    Functional&  form_;
    int          which_argument_;
    mfem::Vector df_;
  };

 private:
  template <typename... T>
  static constexpr int index_of_differentiation()
  {
    constexpr bool d[] = {std::is_same_v<T, differentiate_wrt_this>...};
    for (int i = 0; i < int(sizeof...(T)); i++)
      if (d[i]) return i;
    return SB_NO_DIFFERENTIATION;
  }

  // mfem built with CUDA: Read / Write on the device; otherwise host pointers and copies inside the C ABI call
#ifdef MFEM_USE_CUDA
  static const double* read(const mfem::Vector& v) { return v.Read(true); }
  static double*       write(mfem::Vector& v) { return v.Write(true); }
  static int           transfer_mode() { return SB_DEVICE; }
#else
  static const double* read(const mfem::Vector& v) { return v.HostRead(); }
  static double*       write(mfem::Vector& v) { return v.HostWrite(); }
  static int           transfer_mode() { return SB_HOST_SYNC; }
#endif

  [[noreturn]] void fail(const char* msg) const
  {
    std::fprintf(stderr, "serac::Functional: %s\n", msg);
    std::abort();
  }

  const core_api::FiniteElementSpace* space_of(const fes_t* fes)
  {
    auto it = spaces_.find(fes);
    if (it == spaces_.end())
      it = spaces_.emplace(fes, std::make_unique<core_api::FiniteElementSpace>(mfem_detail::convert_space(*fes))).first;
    return it->second.get();
  }
  const core_api::Mesh* mesh_of(mfem::Mesh& mesh)
  {
    auto it = meshes_.find(&mesh);
    if (it == meshes_.end())
      it = meshes_.emplace(&mesh, std::make_unique<core_api::Mesh>(mfem_detail::convert_mesh(mesh))).first;
    return it->second.get();
  }

  const fes_t*                                                               test_fes_;
  std::array<const fes_t*, num_trial_spaces>                                 trial_fes_;
  int                                                                        test_vsize_;
  std::unique_ptr<core_api::FiniteElementSpace>                              qoi_space_;
  std::map<const fes_t*, std::unique_ptr<core_api::FiniteElementSpace>>      spaces_;
  std::map<const mfem::Mesh*, std::unique_ptr<core_api::Mesh>>               meshes_;
  std::unique_ptr<Impl>                                                      impl_;
  mfem::Vector                                                               output_T_;
  std::vector<Gradient>                                                      grad_;
};

/// Quantity of interest on mfem's types, `Functional<double(trials...)>` (functional_qoi.inl:88-470): operator() returns
/// the scalar (or tuple{scalar, Gradient&}); Gradient::operator()(dx) is the directional derivative and assemble(Gradient&)
/// the gradient as a vector (the reference returns a HypreParVector, :411-455; without MPI an mfem::Vector).
template <typename... trials, ExecutionSpace exec>
class Functional<double(trials...), exec> {
  static constexpr int num_trial_spaces = int(sizeof...(trials));
  using Impl                            = Functional<QOI(trials...), exec>;
  using fes_t                           = mfem::FiniteElementSpace;

 public:
  explicit Functional(std::array<const fes_t*, num_trial_spaces> trial_fes) : impl_(trial_fes)
  {
    grad_.reserve(size_t(num_trial_spaces));
    for (int i = 0; i < num_trial_spaces; i++) grad_.emplace_back(*this, i);
  }

  template <int dim, int... args, typename Integrand, typename DomainType, typename StateType = Nothing>
  void AddDomainIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, DomainType& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    impl_.AddDomainIntegral(d, a, integrand, domain, qdata);
  }
  template <int dim, int... args, typename Integrand, typename DomainType>
  void AddBoundaryIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, DomainType& domain)
  {
    impl_.AddBoundaryIntegral(d, a, integrand, domain);
  }
  template <int... args, typename Integrand, typename DomainType, typename StateType = Nothing>
  void AddAreaIntegral(DependsOn<args...> a, const Integrand& integrand, DomainType& domain,
                       std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    impl_.AddDomainIntegral(Dimension<2>{}, a, integrand, domain, qdata);
  }
  template <int... args, typename Integrand, typename DomainType, typename StateType = Nothing>
  void AddVolumeIntegral(DependsOn<args...> a, const Integrand& integrand, DomainType& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    impl_.AddDomainIntegral(Dimension<3>{}, a, integrand, domain, qdata);
  }

  class Gradient {
   public:
    Gradient(Functional& f, int which) : form_(f), which_(which) {}
    /// functional_qoi.inl:386-390
    double operator()(const mfem::Vector& dx) { return form_.impl_.gradient_of(which_)(dx)(0); }
    /// functional_qoi.inl:411-455
    friend std::unique_ptr<mfem::Vector> assemble(Gradient& g) { return g.assemble_vector(); }
    std::unique_ptr<mfem::Vector>        assemble_vector()
    {
      Gradient& g = *this;
      auto      K = g.form_.impl_.gradient_of(g.which_).assemble_matrix();  // 1 x n matrix
      auto v = std::make_unique<mfem::Vector>(K->Width());
      *v     = 0.0;
      const int*    J = K->GetJ();
      const double* A = K->GetData();
      for (int k = 0; k < K->NumNonZeroElems(); k++) (*v)(J[k]) = A[k];
      return v;
    }

   private:
    Functional& form_;
    int         which_;
  };

  template <typename... T>
  auto operator()(double t, const T&... args)
  {
    constexpr bool differentiated = (std::is_same_v<T, differentiate_wrt_this> || ...);
    if constexpr (differentiated) {
      constexpr bool d[]  = {std::is_same_v<T, differentiate_wrt_this>...};
      int            wrt  = 0;
      for (int i = 0; i < int(sizeof...(T)); i++)
        if (d[i]) wrt = i;
      auto r = impl_(t, args...);
      return std::tuple<double, Gradient&>(std::get<0>(r)(0), grad_[size_t(wrt)]);
    } else {
      const mfem::Vector& r = impl_(t, args...);
      return r(0);
    }
  }

  void updateQdata(bool update_flag) { impl_.updateQdata(update_flag); }

 private:
  Impl                  impl_;
  std::vector<Gradient> grad_;
};

/// shape_aware_functional.hpp:292-575: a Functional whose first argument is a shape displacement field p (H1); the
/// integrands are written on the reference configuration X and evaluated on x = X + p(X)
template <typename shape, typename T, ExecutionSpace exec = default_execution_space>
class ShapeAwareFunctional;

template <typename test, typename shape, typename... trials, ExecutionSpace exec>
class ShapeAwareFunctional<shape, test(trials...), exec> {
  static constexpr int num_trial_spaces = int(sizeof...(trials));
  using functional_t                    = Functional<test(shape, trials...), exec>;
  using fes_t                           = mfem::FiniteElementSpace;
  template <int i>
  using trial_at = typename serac::detail::type_at<i, trials...>::type;

 public:
  /// shape_aware_functional.hpp:311-331
  ShapeAwareFunctional(const fes_t* shape_fes, const fes_t* test_fes, std::array<const fes_t*, num_trial_spaces> trial_fes)
  {
    static_assert(shape::family == Family::H1, "Only H1 spaces allowed for shape displacements");
    std::array<const fes_t*, num_trial_spaces + 1> prepended{};
    prepended[0] = shape_fes;
    for (int i = 0; i < num_trial_spaces; i++) prepended[size_t(1 + i)] = trial_fes[size_t(i)];
    functional_ = std::make_unique<functional_t>(test_fes, prepended);
  }
  /// shape_aware_functional.hpp:333-352: quantities of interest (`test` = double) have no test space
  ShapeAwareFunctional(const fes_t* shape_fes, std::array<const fes_t*, num_trial_spaces> trial_fes)
  {
    static_assert(std::is_same_v<test, double>, "only a quantity of interest is built without a test space");
    static_assert(shape::family == Family::H1, "Only H1 spaces allowed for shape displacements");
    std::array<const fes_t*, num_trial_spaces + 1> prepended{};
    prepended[0] = shape_fes;
    for (int i = 0; i < num_trial_spaces; i++) prepended[size_t(1 + i)] = trial_fes[size_t(i)];
    functional_ = std::make_unique<functional_t>(prepended);
  }

  /// shape_aware_functional.hpp:440-470: the integrand sees (t, x, [state,] trial arguments...); the shape displacement
  /// is argument 0 of the underlying Functional
  template <int dim, int... args, typename Integrand, typename DomainType, typename StateType = Nothing>
  void AddDomainIntegral(Dimension<dim>, DependsOn<args...>, const Integrand& integrand, DomainType& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    if constexpr (std::is_same_v<StateType, Nothing>) {
      using W = ShapeAwareIntegrandWrapper<Integrand, dim, test, trial_at<args>...>;
      functional_->AddDomainIntegral(Dimension<dim>{}, DependsOn<0, (args + 1)...>{}, W(integrand), domain, qdata);
    } else {
      using W = ShapeAwareIntegrandWrapperWithState<Integrand, dim, test, trial_at<args>...>;
      functional_->AddDomainIntegral(Dimension<dim>{}, DependsOn<0, (args + 1)...>{}, W(integrand), domain, qdata);
    }
  }

  /// shape_aware_functional.hpp:474-499
  template <int dim, int... args, typename Integrand, typename DomainType>
  void AddBoundaryIntegral(Dimension<dim>, DependsOn<args...>, const Integrand& integrand, DomainType& domain)
  {
    using W = ShapeAwareBoundaryIntegrandWrapper<Integrand>;
    functional_->AddBoundaryIntegral(Dimension<dim>{}, DependsOn<0, (args + 1)...>{}, W(integrand), domain);
  }

  /// shape_aware_functional.hpp:515-540
  template <typename ShapeArg, typename... T>
  decltype(auto) operator()(double t, const ShapeArg& shape_displacement, const T&... args)
  {
    return (*functional_)(t, shape_displacement, args...);
  }

  void updateQdata(bool update_flag) { functional_->updateQdata(update_flag); }

 private:
  std::unique_ptr<functional_t> functional_;
};

}  // namespace serac
