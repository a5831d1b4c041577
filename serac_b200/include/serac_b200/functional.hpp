// serac_b200: the C++ host side -- `serac::Functional<test(trials...)>` with the reference's public API, driving the
// C ABI (include/serac_b200.h) and nothing else.
//
// Mirrors (names, argument meaning, ownership and error behaviour; not code) of
//   src/serac/numerics/functional/functional.hpp:189-745   Functional, Gradient, assemble, differentiate_wrt, updateQdata
//   src/serac/numerics/functional/domain.hpp               Domain, EntireDomain
//   src/serac/numerics/functional/quadrature_data.hpp      QuadratureData<T>, NoQData, EmptyQData
// The mesh / finite-element-space layer is mfem's in the reference and stays out of scope: the three
// small host structs below (FiniteElementSpace, Mesh, Vector) carry exactly what Functional reads from
// mfem::ParFiniteElementSpace / mfem::Mesh / mfem::Vector -- native element dofs, node coordinates, sizes -- so an
// mfem build fills them with GetElementDofs / GetNodes and everything downstream is unchanged.
//
// q-functions are device code: translation units that instantiate Functional are compiled by nvcc
// (-std=c++20 --expt-relaxed-constexpr).  nvcc does not allow generic (auto-parameter) extended lambdas, so a
// q-function is a functor struct with a templated operator() -- the form serac's own physics modules use.
//
// Errors: the reference aborts through SLIC_ERROR_ROOT (functional.hpp:77-98); here a non-zero C-ABI status prints
// sb_last_error() and aborts.  There is no CPU fallback.
#pragma once

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <numeric>
#include <tuple>
#include <vector>

#include "serac_b200.h"
#include "serac_b200/kernel_table.cuh"

namespace serac {

#define SERAC_B200_CHECK(expr)                                                                      \
  do {                                                                                              \
    int _rc = (expr);                                                                               \
    if (_rc != SB_OK) {                                                                             \
      std::fprintf(stderr, "serac::Functional: %s failed (%d): %s\n", #expr, _rc, sb_last_error()); \
      std::abort();                                                                                 \
    }                                                                                               \
  } while (0)

#define SERAC_B200_CUDA(expr)                                                                                 \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess) {                                                                                  \
      std::fprintf(stderr, "serac::Functional: %s failed: %s\n", #expr, cudaGetErrorString(_e));              \
      std::abort();                                                                                           \
    }                                                                                                         \
  } while (0)

/// The host side with PLAIN structs for the mesh / space layer (FiniteElementSpace, Mesh, Vector below).  It lives in its
/// own namespace so that serac_b200/functional_mfem.hpp can offer the same class names on mfem's types
/// (`serac::Functional` over mfem::FiniteElementSpace / mfem::Vector / mfem::Operator); without that header
/// (SERAC_B200_MFEM_API undefined) the names are visible as serac::Functional, serac::Vector, ... as before.
namespace core_api {

enum class Ordering : int
{
  byNODES = SB_ORDERING_BYNODES,
  byVDIM  = SB_ORDERING_BYVDIM
};

/// device-resident vector (stand-in for mfem::Vector with device memory)
class Vector {
 public:
  Vector() = default;
  explicit Vector(int64_t n) { SetSize(n); }
  explicit Vector(const std::vector<double>& host) { SetSize(int64_t(host.size())); set(host.data()); }
  Vector(const Vector& o) { *this = o; }
  Vector& operator=(const Vector& o)
  {
    if (this != &o) {
      SetSize(o.n_);
      if (n_) SERAC_B200_CUDA(cudaMemcpy(d_, o.d_, sizeof(double) * size_t(n_), cudaMemcpyDeviceToDevice));
    }
    return *this;
  }
  ~Vector() { cudaFree(d_); }
  void SetSize(int64_t n)
  {
    if (n == n_) return;
    cudaFree(d_);
    d_ = nullptr;
    n_ = n;
    if (n) SERAC_B200_CUDA(cudaMalloc(&d_, sizeof(double) * size_t(n)));
  }
  int64_t       Size() const { return n_; }
  const double* Read() const { return d_; }  ///< device pointer
  double*       Write() { return d_; }
  void set(const double* host) { SERAC_B200_CUDA(cudaMemcpy(d_, host, sizeof(double) * size_t(n_), cudaMemcpyHostToDevice)); }
  void get(double* host) const { SERAC_B200_CUDA(cudaMemcpy(host, d_, sizeof(double) * size_t(n_), cudaMemcpyDeviceToHost)); }
  std::vector<double> to_host() const
  {
    std::vector<double> h(static_cast<size_t>(n_));
    if (n_) get(h.data());
    return h;
  }
  Vector& operator=(double v)
  {
    std::vector<double> h(size_t(n_), v);
    if (n_) set(h.data());
    return *this;
  }

 private:
  int64_t n_ = 0;
  double* d_ = nullptr;
};

/// the elements (or boundary faces) of one further geometry of a mesh of mixed element geometry: the reference keeps a
/// restriction per geometry (BlockElementRestriction, element_restriction.cpp:471-536) and a kernel per geometry
/// (integral.hpp:254-257)
struct GeometryBlock {
  Geometry         geometry;
  std::vector<int> ids;          ///< mesh element (boundary face) index of every entry, ascending
  std::vector<int> native_dofs;  ///< [ids.size()][nodes per element], mfem's native order
};

/// what Functional reads from an mfem::ParFiniteElementSpace: H1 order, vdim, ordering, GetNDofs() and, per element,
/// fes->GetElementDofs(e) in mfem's native order.  Single-geometry meshes fill the fields below; on meshes of mixed
/// element geometry they describe the elements of the FIRST geometry (`element_ids` = their mesh indices) and
/// `more_elements` the others; likewise for the boundary faces.
struct FiniteElementSpace {
  Geometry         geometry;
  int              order;
  int              vdim;
  Ordering         ordering = Ordering::byVDIM;  // serac's default (serac_config.hpp.in:61-66)
  int64_t          num_nodes;
  int64_t          num_elements;
  std::vector<int> native_element_dofs;  ///< [num_elements][nodes per element]
  /// fes->GetFaceDofs of the mesh's boundary faces in mfem's native order (only needed by AddBoundaryIntegral)
  int64_t          num_boundary_faces = 0;
  std::vector<int> native_boundary_face_dofs;  ///< [num_boundary_faces][nodes per face]
  Family           family = Family::H1;  ///< H1, or L2 (discontinuous: mfem hands out its element dofs in lexicographic order)
  int64_t          GetVSize() const { return num_nodes * vdim; }
  /// mixed meshes only
  std::vector<int>           element_ids;  ///< mesh index of every element of `geometry` (empty: all elements, in order)
  std::vector<GeometryBlock> more_elements;
  Geometry                   boundary_face_geometry = Geometry::POINT;  ///< POINT: the face geometry of `geometry`
  std::vector<int>           boundary_face_ids;  ///< mesh face index of every boundary face above (for boundary Domains)
  std::vector<GeometryBlock> more_boundary_faces;
  int64_t                    total_elements() const
  {
    int64_t n = num_elements;
    for (const auto& b : more_elements) n += int64_t(b.ids.size());
    return n;
  }
};

/// what Functional reads from an mfem::Mesh: the nodal grid function (Functional demands nodes, functional.hpp:74-89)
struct Mesh {
  FiniteElementSpace  nodes_space;  ///< vdim == space dimension
  std::vector<double> nodes;        ///< L-vector of node coordinates in nodes_space's ordering
  int                 Dimension() const { return dimension_of(nodes_space.geometry); }
  int64_t             GetNE() const { return nodes_space.total_elements(); }
};

/// a set of elements of one mesh (domain.hpp); EntireDomain(mesh) = all of them
struct Domain {
  const Mesh*      mesh;
  bool             entire;
  std::vector<int> elements;  ///< element ids when !entire (boundary domains: mesh face ids, ascending)
  bool             boundary = false;  ///< Domain::Type::BoundaryElements (domain.hpp:27-33)
};
inline Domain EntireDomain(const Mesh& mesh) { return Domain{&mesh, true, {}}; }

/// per-quadrature-point state buffer [elements][qpts] (quadrature_data.hpp:112-142), device resident
template <typename T>
struct QuadratureData {
  QuadratureData(int64_t num_elements, int qpts_per_element, T init = T{}) : count(num_elements * qpts_per_element)
  {
    std::vector<T> h(size_t(count), init);
    SERAC_B200_CUDA(cudaMalloc(&data, sizeof(T) * size_t(count)));
    SERAC_B200_CUDA(cudaMemcpy(data, h.data(), sizeof(T) * size_t(count), cudaMemcpyHostToDevice));
  }
  ~QuadratureData() { cudaFree(data); }
  QuadratureData(const QuadratureData&)            = delete;
  QuadratureData& operator=(const QuadratureData&) = delete;
  std::vector<T> to_host() const
  {
    std::vector<T> h(static_cast<size_t>(count));
    SERAC_B200_CUDA(cudaMemcpy(h.data(), data, sizeof(T) * size_t(count), cudaMemcpyDeviceToHost));
    return h;
  }
  /// restore the state from a checkpoint written with to_host() (the reference saves / reloads quadrature data through
  /// sidre next to the fields, src/serac/physics/state/state_manager.cpp:213-224; the storage format stays the host's)
  void from_host(const std::vector<T>& h)
  {
    if (int64_t(h.size()) != count) {
      std::fprintf(stderr, "QuadratureData::from_host: %zu states given, %lld expected\n", h.size(), (long long)count);
      std::abort();
    }
    SERAC_B200_CUDA(cudaMemcpy(data, h.data(), sizeof(T) * size_t(count), cudaMemcpyHostToDevice));
  }
  int64_t count;
  T*      data = nullptr;
};
/// sentinels (quadrature_data.hpp:20-37)
inline const std::shared_ptr<QuadratureData<Nothing>> NoQData{};
inline const std::shared_ptr<QuadratureData<Empty>>   EmptyQData{};

/// tag returned by differentiate_wrt (functional.hpp:47-66)
struct differentiate_wrt_this {
  const Vector& ref;
  operator const Vector&() const { return ref; }
};
inline differentiate_wrt_this differentiate_wrt(const Vector& v) { return differentiate_wrt_this{v}; }

/// the local CSR triple assemble() hands back (what the reference wraps as mfem::SparseMatrix -> HypreParMatrix,
/// functional.hpp:649-662; at one rank RAP with P = identity leaves it unchanged)
struct CsrMatrix {
  int64_t             num_rows = 0, num_cols = 0;
  std::vector<int>    row_ptr, col_ind;
  std::vector<double> values;
};

template <typename T>
class Functional;

template <typename test, typename... trials>
class Functional<test(trials...)> {
  static constexpr int num_trial_spaces = int(sizeof...(trials));
  static constexpr int Q                = std::max({test::order, trials::order...}) + 1;  // functional.hpp:193

  template <int i>
  using trial_at = typename serac::detail::type_at<i, trials...>::type;

 public:
  class Gradient;

  /// functional.hpp:219-281: one element restriction per space, built from the space's native element dofs
  Functional(const FiniteElementSpace* test_fes, std::array<const FiniteElementSpace*, num_trial_spaces> trial_fes)
      : test_space_(test_fes), trial_space_(trial_fes)
  {
    constexpr int    orders[] = {trials::order..., 0};
    constexpr int    comps[]  = {trials::components..., 0};
    constexpr Family fams[]   = {trials::family..., Family::H1};
    check_space(*test_fes, test::order, test::components, "test", test::family);
    G_test_ = make_restriction(*test_fes);
    for (int i = 0; i < num_trial_spaces; i++) {
      check_space(*trial_fes[size_t(i)], orders[i], comps[i], "trial", fams[i]);
      // spaces that are the same object share one restriction
      G_trial_[size_t(i)] = nullptr;
      if (trial_fes[size_t(i)] == test_fes) G_trial_[size_t(i)] = G_test_;
      for (int j = 0; j < i; j++)
        if (trial_fes[size_t(j)] == trial_fes[size_t(i)]) G_trial_[size_t(i)] = G_trial_[size_t(j)];
      if (!G_trial_[size_t(i)]) G_trial_[size_t(i)] = make_restriction(*trial_fes[size_t(i)]);
    }
    // meshes of mixed element geometry: one more restriction per space and further geometry
    for (size_t b = 0; b < test_fes->more_elements.size(); b++) {
      G_test_more_.push_back(make_block_restriction(*test_fes, test_fes->more_elements[b]));
      std::array<sb_restriction_t, num_trial_spaces> row{};
      for (int i = 0; i < num_trial_spaces; i++) {
        const FiniteElementSpace& k = *trial_fes[size_t(i)];
        if (k.more_elements.size() != test_fes->more_elements.size() ||
            k.more_elements[b].geometry != test_fes->more_elements[b].geometry ||
            k.more_elements[b].ids.size() != test_fes->more_elements[b].ids.size())
          fail("Functional: the spaces do not live on the same (mixed) mesh");
        row[size_t(i)] = nullptr;
        if (trial_fes[size_t(i)] == test_fes) row[size_t(i)] = G_test_more_.back();
        for (int j = 0; j < i; j++)
          if (trial_fes[size_t(j)] == trial_fes[size_t(i)]) row[size_t(i)] = row[size_t(j)];
        if (!row[size_t(i)]) row[size_t(i)] = make_block_restriction(k, k.more_elements[b]);
      }
      G_trial_more_.push_back(row);
    }
    SERAC_B200_CHECK(sb_functional_create(G_test_, num_trial_spaces, G_trial_.data(), &F_));
    output_L_.SetSize(test_fes->GetVSize());
    grad_.reserve(size_t(num_trial_spaces));
    for (int i = 0; i < num_trial_spaces; i++) grad_.emplace_back(*this, i);
  }

  ~Functional()
  {
    sb_functional_destroy(F_);
    for (auto I : integrals_) sb_integral_destroy(I);
    for (auto A : csr_) sb_csr_destroy(A);
    std::vector<sb_restriction_t> freed;
    auto release = [&](sb_restriction_t r) {
      if (r && std::find(freed.begin(), freed.end(), r) == freed.end()) {
        freed.push_back(r);
        sb_restriction_destroy(r);
      }
    };
    release(G_test_);
    for (auto r : G_trial_) release(r);
    for (auto r : G_test_more_) release(r);
    for (auto& row : G_trial_more_)
      for (auto r : row) release(r);
    for (auto r : mesh_restrictions_) release(r);
    for (auto p : device_buffers_) cudaFree(p);
  }
  Functional(const Functional&)            = delete;
  Functional& operator=(const Functional&) = delete;

  /// functional.hpp:286-317.  integrand: callable `(double t, tuple{x, dx_dxi}, [State&,] tuple{value, grad}...)
  /// -> tuple{source, flux}`; args...: which Functional arguments it depends on, in the order it takes them.
  template <int dim, int... args, typename Integrand, typename StateType = Nothing>
  void AddDomainIntegral(Dimension<dim>, DependsOn<args...>, const Integrand& integrand, Domain& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    static_assert(((args >= 0 && args < num_trial_spaces) && ...), "DependsOn<> index out of range");
    const Mesh& mesh = *domain.mesh;
    if (mesh.Dimension() != dim) fail("AddDomainIntegral: Dimension<dim> does not match the mesh");
    const Geometry g = mesh.nodes_space.geometry;
    if (g != test_space_->geometry) fail("AddDomainIntegral: mesh and test space geometries differ");
    if (!mesh.nodes_space.more_elements.empty()) {  // one integral per geometry (integral.hpp:254-257)
      if constexpr (!std::is_same_v<StateType, Nothing> && !std::is_empty_v<StateType>)
        fail("AddDomainIntegral: quadrature-point state on meshes of mixed element geometry is not supported");
      add_domain_integral_block<dim, Integrand, StateType, args...>(integrand, domain, mesh.nodes_space.geometry,
                                                                    mesh.nodes_space.element_ids, make_restriction(mesh.nodes_space),
                                                                    G_test_, G_trial_);
      for (size_t b = 0; b < mesh.nodes_space.more_elements.size(); b++) {
        const GeometryBlock& blk = mesh.nodes_space.more_elements[b];
        if (b >= G_test_more_.size() || blk.geometry != test_space_->more_elements[b].geometry)
          fail("AddDomainIntegral: mesh and test space geometries differ");
        add_domain_integral_block<dim, Integrand, StateType, args...>(integrand, domain, blk.geometry, blk.ids,
                                                                      make_block_restriction(mesh.nodes_space, blk),
                                                                      G_test_more_[b], G_trial_more_[b]);
      }
      return;
    }
    const sb_kernel_table* table = nullptr;
    // the reference instantiates the kernels of every geometry of this dimension (functional.hpp:296-317)
    if constexpr (dim == 3) {
      if (g == Geometry::CUBE) table = table_for<Geometry::CUBE, Integrand, StateType, args...>();
      if (g == Geometry::TETRAHEDRON) table = table_for<Geometry::TETRAHEDRON, Integrand, StateType, args...>();
    } else if constexpr (dim == 2) {
      if (g == Geometry::SQUARE) table = table_for<Geometry::SQUARE, Integrand, StateType, args...>();
      if (g == Geometry::TRIANGLE) table = table_for<Geometry::TRIANGLE, Integrand, StateType, args...>();
    }
    if (!table) fail("AddDomainIntegral: unsupported element geometry");

    // GeometricFactors (geometric_factors.cpp:62-138) on the domain's elements
    sb_restriction_t G_mesh = make_restriction(mesh.nodes_space);
    mesh_restrictions_.push_back(G_mesh);
    const int64_t ne  = domain.entire ? mesh.GetNE() : int64_t(domain.elements.size());
    const int     nq  = num_quadrature_points(g, Q);
    double*       nodes_dev = device_copy(mesh.nodes.data(), mesh.nodes.size());
    int*          elems_dev = domain.entire ? nullptr : device_copy(domain.elements.data(), domain.elements.size());
    double*       X_q = device_alloc<double>(size_t(ne) * size_t(dim) * size_t(nq));
    double*       J_q = device_alloc<double>(size_t(ne) * size_t(dim) * size_t(dim) * size_t(nq));
    SERAC_B200_CHECK(sb_geometric_factors(int(g), mesh.nodes_space.order, Q, G_mesh, nodes_dev, elems_dev, ne, X_q, J_q,
                                          nullptr));
    sb_integral_t I = nullptr;
    SERAC_B200_CHECK(sb_integral_create(table, &integrand, domain.entire ? nullptr : domain.elements.data(), ne, X_q, J_q,
                                        &I));
    if constexpr (!std::is_same_v<StateType, Nothing> && !std::is_empty_v<StateType>) {
      if (!qdata) fail("AddDomainIntegral: integrand with state needs a QuadratureData buffer");
      if (qdata->count != ne * nq) fail("AddDomainIntegral: QuadratureData size does not match elements x qpts");
      SERAC_B200_CHECK(sb_integral_set_qdata(I, qdata->data));
      qdata_keepalive_.push_back(qdata);
    }
    const int which[] = {args..., 0};
    SERAC_B200_CHECK(sb_functional_add_integral(F_, I, which));
    integrals_.push_back(I);
  }

  /// functional.hpp:328-354.  dim = dimension of the boundary faces; integrand: `(t, tuple{x, dx_dxi},
  /// tuple{value, d/dxi}...) -> source` (trial derivatives stay in parent coordinates on a boundary).  Every space and
  /// the mesh must carry its native boundary-face dofs.  Faces of 3D meshes (quadrilaterals / triangles) and of 2D
  /// meshes (segments, F/detail/segment_H1.inl).
  template <int dim, int... args, typename Integrand>
  void AddBoundaryIntegral(Dimension<dim>, DependsOn<args...>, const Integrand& integrand, const Mesh& mesh,
                           const std::vector<int>* face_ids = nullptr)
  {
    static_assert(dim == 1 || dim == 2, "boundary integrals live on the faces of 2D or 3D meshes");
    static_assert(((args >= 0 && args < num_trial_spaces) && ...), "DependsOn<> index out of range");
    if (mesh.Dimension() != dim + 1) fail("AddBoundaryIntegral: Dimension<dim> must be the mesh dimension - 1");
    const Geometry g  = mesh.nodes_space.geometry;
    Geometry fg = (dim == 1) ? Geometry::SEGMENT : (g == Geometry::CUBE) ? Geometry::SQUARE : Geometry::TRIANGLE;
    if (mesh.nodes_space.boundary_face_geometry != Geometry::POINT) fg = mesh.nodes_space.boundary_face_geometry;
    // one integral per face geometry (triangular and quadrilateral faces of a mixed 3D mesh)
    add_boundary_integral_block<dim, Integrand, args...>(integrand, mesh, fg, nullptr, face_ids);
    for (size_t b = 0; b < mesh.nodes_space.more_boundary_faces.size(); b++)
      add_boundary_integral_block<dim, Integrand, args...>(integrand, mesh, mesh.nodes_space.more_boundary_faces[b].geometry,
                                                           &b, face_ids);
  }
  /// ... on part of the boundary (Domain::ofBoundaryElements, domain.hpp:112-127)
  template <int dim, int... args, typename Integrand>
  void AddBoundaryIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, const Domain& domain)
  {
    if (!domain.boundary) fail("AddBoundaryIntegral: the domain is not a set of boundary elements");
    AddBoundaryIntegral(d, a, integrand, *domain.mesh, domain.entire ? nullptr : &domain.elements);
  }

  /// functional.hpp:446-513: at most one argument may be wrapped in differentiate_wrt(); then the return value is
  /// tuple{residual, Gradient&}, else the residual alone.  The references stay valid until the next call.
  template <typename... T>
  decltype(auto) operator()(double t, const T&... args)
  {
    static_assert(sizeof...(T) == sizeof...(trials), "one argument per trial space");
    constexpr int num_differentiated = (int(std::is_same_v<T, differentiate_wrt_this>) + ...);
    static_assert(num_differentiated <= 1, "at most one argument can be differentiated (functional.hpp:504)");
    constexpr int wrt             = index_of_differentiation<T...>();
    const Vector* in[]            = {&static_cast<const Vector&>(args)...};
    const double* ptrs[sizeof...(T)];
    for (size_t i = 0; i < sizeof...(T); i++) {
      if (in[i]->Size() != trial_space_[i]->GetVSize()) fail("operator(): argument size does not match its space");
      ptrs[i] = in[i]->Read();
    }
    SERAC_B200_CHECK(sb_functional_evaluate(F_, t, wrt, ptrs, 0, output_L_.Write(), 0, update_qdata_ ? 1 : 0, nullptr));
    if constexpr (wrt >= 0) {
      return std::tuple<Vector&, Gradient&>(output_L_, grad_[size_t(wrt)]);
    } else {
      return static_cast<Vector&>(output_L_);
    }
  }

  /// operator() on raw buffers: `mode` = SB_DEVICE / SB_HOST_SYNC for inputs and output alike (what an adapter over
  /// another library's vector type calls; sizes are the caller's responsibility)
  void evaluate_raw(double t, int wrt, const double* const* inputs, double* output, int mode)
  {
    SERAC_B200_CHECK(sb_functional_evaluate(F_, t, wrt, inputs, mode, output, mode, update_qdata_ ? 1 : 0, nullptr));
  }
  void jvp_raw(int wrt, const double* dx, double* df, int mode)
  {
    SERAC_B200_CHECK(sb_functional_jvp(F_, wrt, dx, mode, df, mode, nullptr));
  }
  Gradient& gradient(int wrt) { return grad_[size_t(wrt)]; }

  /// functional.hpp:523: whether the next evaluations commit the quadrature-point state
  void updateQdata(bool update_flag) { update_qdata_ = update_flag; }

  /// functional.hpp:532-709: mfem::Operator whose Mult is the Jacobian action with the derivatives stored by the last
  /// operator() that differentiated w.r.t. this argument
  class Gradient {
   public:
    Gradient(Functional& f, int which) : form_(f), which_argument_(which), df_(f.test_space_->GetVSize()) {}
    int64_t Height() const { return form_.test_space_->GetVSize(); }
    int64_t Width() const { return form_.trial_space_[size_t(which_argument_)]->GetVSize(); }
    /// Gradient::Mult (functional.hpp:556-559)
    void Mult(const Vector& dx, Vector& df) const
    {
      if (dx.Size() != Width() || df.Size() != Height()) form_.fail("Gradient::Mult: size mismatch");
      SERAC_B200_CHECK(sb_functional_jvp(form_.F_, which_argument_, dx.Read(), 0, df.Write(), 0, nullptr));
    }
    /// functional.hpp:562
    Vector& operator()(const Vector& dx)
    {
      Mult(dx, df_);
      return df_;
    }
    /// assemble(Gradient&) (functional.hpp:569-667) up to the local CSR triple
    friend std::unique_ptr<CsrMatrix> assemble(Gradient& g) { return g.assemble_local(); }

    /// the sparsity pattern of d(test)/d(this argument), created on first use and owned by the Functional (the
    /// reference's lookup tables, functional.hpp:571-583)
    sb_csr_t pattern()
    {
      Functional& f = form_;
      const int   w = which_argument_;
      if (f.csr_.size() < size_t(num_trial_spaces)) f.csr_.resize(size_t(num_trial_spaces), nullptr);
      if (!f.csr_[size_t(w)]) {
        // one block per element geometry of the mesh (BlockElementRestriction, element_restriction.cpp:471-536)
        std::vector<sb_restriction_t> row_blocks{f.G_test_}, col_blocks{f.G_trial_[size_t(w)]};
        for (size_t b = 0; b < f.G_test_more_.size(); b++) {
          row_blocks.push_back(f.G_test_more_[b]);
          col_blocks.push_back(f.G_trial_more_[b][size_t(w)]);
        }
        SERAC_B200_CHECK(sb_csr_create_blocks(int(row_blocks.size()), row_blocks.data(), col_blocks.data(),
                                              &f.csr_[size_t(w)]));
      }
      return f.csr_[size_t(w)];
    }
    /// the values of the assembled matrix into a device array of sb_csr_nnz(pattern) doubles (serac_b200/solvers.hpp)
    void assemble_values(sb_csr_t A, double* values_dev)
    {
      SERAC_B200_CHECK(sb_functional_assemble(form_.F_, which_argument_, &A, values_dev, 0, nullptr));
    }
    sb_functional_t functional_handle() const { return form_.F_; }
    int             argument() const { return which_argument_; }

    std::unique_ptr<CsrMatrix> assemble_local()
    {
      Gradient&   g = *this;
      Functional& f = g.form_;
      const int   w = g.which_argument_;
      sb_csr_t    A   = pattern();
      auto        out = std::make_unique<CsrMatrix>();
      out->num_rows = sb_csr_num_rows(A);
      out->num_cols = g.Width();
      out->row_ptr.resize(size_t(out->num_rows) + 1);
      out->col_ind.resize(size_t(sb_csr_nnz(A)));
      out->values.resize(size_t(sb_csr_nnz(A)));
      SERAC_B200_CHECK(sb_functional_assemble(f.F_, w, &A, out->values.data(), 1, nullptr));
      SERAC_B200_CHECK(sb_csr_pattern_to_host(A, out->row_ptr.data(), out->col_ind.data()));
      return out;
    }

   private:
    Functional& form_;
    int         which_argument_;
    Vector      df_;
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

  template <Geometry g, typename Integrand, typename StateType, int... args>
  static const sb_kernel_table* table_for()
  {
    using Sig = b200::Signature<g, Q, test, trial_at<args>...>;
    // arbitrary user functors are differentiated with dual numbers, the reference's scheme; functors that provide their
    // own tangent (physics_functors.hpp) opt in through provides_flux_tangent
    if constexpr (test::family == Family::QOI) {  // quantity of interest: the q-function returns the scalar integrand
      static const sb_kernel_table table = b200::make_qoi_kernel_table<Sig, Integrand, StateType>();
      return &table;
    } else {
      static const sb_kernel_table table = b200::make_kernel_table<Sig, Integrand, StateType>();
      return &table;
    }
  }

  template <Geometry fg, typename Integrand, int... args>
  static const sb_kernel_table* boundary_table_for()
  {
    using Sig = b200::BoundarySignature<fg, Q, test, trial_at<args>...>;
    static const sb_kernel_table table = b200::make_boundary_kernel_table<Sig, Integrand>();
    return &table;
  }

  /// FaceType::BOUNDARY ElementRestriction (element_restriction.cpp:279-388) from the native boundary-face dofs
  sb_restriction_t make_face_restriction(const FiniteElementSpace& fes, Geometry fg) const
  {
    const int npf = b200::num_nodes(fg, fes.order);
    if (fes.native_boundary_face_dofs.size() != size_t(fes.num_boundary_faces) * size_t(npf))
      fail("AddBoundaryIntegral: a space does not carry its native boundary-face dofs");
    std::vector<uint64_t> dof_info(size_t(fes.num_boundary_faces) * size_t(npf));
    if (fes.family == Family::L2) {
      // DG spaces: the caller took the face dofs from the adjacent elements (sb_l2_face_dof_info), lexicographic already
      for (size_t i = 0; i < dof_info.size(); i++) dof_info[i] = uint64_t(fes.native_boundary_face_dofs[i]);
    } else {
      SERAC_B200_CHECK(sb_h1_lexicographic_dof_info(int(fg), fes.order, fes.native_boundary_face_dofs.data(),
                                                    fes.num_boundary_faces, dof_info.data()));
    }
    sb_restriction_t r = nullptr;
    SERAC_B200_CHECK(sb_restriction_create(dof_info.data(), fes.num_boundary_faces, npf, fes.vdim, fes.num_nodes,
                                           int(fes.ordering), &r));
    return r;
  }

  [[noreturn]] void fail(const char* msg) const
  {
    std::fprintf(stderr, "serac::Functional: %s\n", msg);
    std::abort();
  }

  void check_space(const FiniteElementSpace& fes, int order, int components, const char* what, Family family) const
  {
    if ((family == Family::L2) != (fes.family == Family::L2)) {
      std::fprintf(stderr, "serac::Functional: %s space is %s but the template signature says %s\n", what,
                   fes.family == Family::L2 ? "L2" : "H1", family == Family::L2 ? "L2" : "H1");
      std::abort();
    }
    if (fes.order != order || fes.vdim != components) {
      std::fprintf(stderr, "serac::Functional: %s space is H1<%d,%d> but the template signature says H1<%d,%d>\n", what,
                   fes.order, fes.vdim, order, components);
      std::abort();
    }
  }

  /// ElementRestriction ctor (element_restriction.cpp:392-404) from the native element dofs
  static sb_restriction_t make_restriction(const FiniteElementSpace& fes)
  {
    const int             npe = b200::num_nodes(fes.geometry, fes.order);
    std::vector<uint64_t> dof_info(size_t(fes.num_elements) * size_t(npe));
    if (fes.family == Family::L2) {
      // "mfem returns DG dofs in lexicographic order already so no permutation is required" (element_restriction.cpp:262-268)
      for (size_t i = 0; i < dof_info.size(); i++) dof_info[i] = uint64_t(fes.native_element_dofs[i]);
    } else {
      SERAC_B200_CHECK(sb_h1_lexicographic_dof_info(int(fes.geometry), fes.order, fes.native_element_dofs.data(),
                                                    fes.num_elements, dof_info.data()));
    }
    sb_restriction_t r = nullptr;
    SERAC_B200_CHECK(sb_restriction_create(dof_info.data(), fes.num_elements, npe, fes.vdim, fes.num_nodes,
                                           int(fes.ordering), &r));
    return r;
  }

  /// the restriction of a space on the elements of one further geometry of a mixed mesh
  static sb_restriction_t make_block_restriction(const FiniteElementSpace& fes, const GeometryBlock& blk)
  {
    FiniteElementSpace view{};
    view.geometry            = blk.geometry;
    view.order               = fes.order;
    view.vdim                = fes.vdim;
    view.ordering            = fes.ordering;
    view.num_nodes           = fes.num_nodes;
    view.num_elements        = int64_t(blk.ids.size());
    view.native_element_dofs = blk.native_dofs;
    view.family              = fes.family;
    return make_restriction(view);
  }

  /// the domain integral on the elements of ONE geometry of a mixed mesh: `ids` = their mesh indices (empty: identity)
  template <int dim, typename Integrand, typename StateType, int... args>
  void add_domain_integral_block(const Integrand& integrand, const Domain& domain, Geometry g, const std::vector<int>& ids,
                                 sb_restriction_t G_mesh, sb_restriction_t test_block,
                                 const std::array<sb_restriction_t, num_trial_spaces>& trial_blocks)
  {
    mesh_restrictions_.push_back(G_mesh);
    const Mesh&            mesh  = *domain.mesh;
    const sb_kernel_table* table = nullptr;
    if constexpr (dim == 3) {
      if (g == Geometry::CUBE) table = table_for<Geometry::CUBE, Integrand, StateType, args...>();
      if (g == Geometry::TETRAHEDRON) table = table_for<Geometry::TETRAHEDRON, Integrand, StateType, args...>();
    } else if constexpr (dim == 2) {
      if (g == Geometry::SQUARE) table = table_for<Geometry::SQUARE, Integrand, StateType, args...>();
      if (g == Geometry::TRIANGLE) table = table_for<Geometry::TRIANGLE, Integrand, StateType, args...>();
    }
    if (!table) fail("AddDomainIntegral: unsupported element geometry");
    // the domain's elements of this geometry, as positions in the block (domain.hpp: one list per geometry)
    const int64_t    nblock = sb_restriction_num_elements(test_block);
    std::vector<int> local;
    if (!domain.entire) {
      for (int e : domain.elements) {
        if (ids.empty()) {
          local.push_back(e);
        } else {
          auto it = std::lower_bound(ids.begin(), ids.end(), e);
          if (it != ids.end() && *it == e) local.push_back(int(it - ids.begin()));
        }
      }
    }
    const int64_t ne = domain.entire ? nblock : int64_t(local.size());
    if (ne == 0) return;
    const int nq        = num_quadrature_points(g, Q);
    double*   nodes_dev = device_copy(mesh.nodes.data(), mesh.nodes.size());
    int*      elems_dev = domain.entire ? nullptr : device_copy(local.data(), local.size());
    double*   X_q       = device_alloc<double>(size_t(ne) * size_t(dim) * size_t(nq));
    double*   J_q       = device_alloc<double>(size_t(ne) * size_t(dim) * size_t(dim) * size_t(nq));
    SERAC_B200_CHECK(sb_geometric_factors(int(g), mesh.nodes_space.order, Q, G_mesh, nodes_dev, elems_dev, ne, X_q, J_q,
                                          nullptr));
    sb_integral_t I = nullptr;
    SERAC_B200_CHECK(sb_integral_create(table, &integrand, domain.entire ? nullptr : local.data(), ne, X_q, J_q, &I));
    const int        which[] = {args..., 0};
    sb_restriction_t trial_r[sizeof...(args) + 1];
    for (size_t i = 0; i < sizeof...(args); i++) trial_r[i] = trial_blocks[size_t(which[i])];
    SERAC_B200_CHECK(sb_functional_add_integral_on(F_, I, which, test_block, trial_r));
    integrals_.push_back(I);
  }

  /// the boundary integral on the faces of ONE geometry (which == nullptr: the first, else more_boundary_faces[*which])
  template <int dim, typename Integrand, int... args>
  void add_boundary_integral_block(const Integrand& integrand, const Mesh& mesh, Geometry fg, const size_t* which_block,
                                   const std::vector<int>* face_ids)
  {
    const sb_kernel_table* table = nullptr;
    if constexpr (dim == 2) {
      if (fg == Geometry::SQUARE) table = boundary_table_for<Geometry::SQUARE, Integrand, args...>();
      if (fg == Geometry::TRIANGLE) table = boundary_table_for<Geometry::TRIANGLE, Integrand, args...>();
    } else {
      table = boundary_table_for<Geometry::SEGMENT, Integrand, args...>();
    }
    if (!table) fail("AddBoundaryIntegral: unsupported face geometry");
    auto faces_of = [&](const FiniteElementSpace& fes) {
      if (!which_block) return make_face_restriction(fes, fg);
      if (*which_block >= fes.more_boundary_faces.size()) fail("AddBoundaryIntegral: a space lacks a boundary-face block");
      const GeometryBlock& blk = fes.more_boundary_faces[*which_block];
      FiniteElementSpace   view{};
      view.geometry = fes.geometry, view.order = fes.order, view.vdim = fes.vdim, view.ordering = fes.ordering;
      view.num_nodes = fes.num_nodes, view.family = fes.family;
      view.num_boundary_faces        = int64_t(blk.ids.size());
      view.native_boundary_face_dofs = blk.native_dofs;
      return make_face_restriction(view, fg);
    };
    sb_restriction_t G_mesh = faces_of(mesh.nodes_space);
    mesh_restrictions_.push_back(G_mesh);
    // the domain's faces of this geometry as positions in the block's face list
    std::vector<int> local;
    if (face_ids) {
      const std::vector<int>& ids = which_block ? mesh.nodes_space.more_boundary_faces[*which_block].ids
                                                : mesh.nodes_space.boundary_face_ids;
      if (ids.empty() && sb_restriction_num_elements(G_mesh) > 0)
        fail("AddBoundaryIntegral: the mesh does not carry the face ids of its boundary faces");
      for (int f : *face_ids) {
        auto it = std::lower_bound(ids.begin(), ids.end(), f);
        if (it != ids.end() && *it == f) local.push_back(int(it - ids.begin()));
      }
    }
    const int64_t nf = face_ids ? int64_t(local.size()) : sb_restriction_num_elements(G_mesh);
    if (nf == 0) return;
    const int nq        = num_quadrature_points(fg, Q);
    double*   nodes_dev = device_copy(mesh.nodes.data(), mesh.nodes.size());
    int*      faces_dev = face_ids ? device_copy(local.data(), local.size()) : nullptr;
    double*   X_q       = device_alloc<double>(size_t(nf) * size_t(dim + 1) * size_t(nq));
    double*   J_q       = device_alloc<double>(size_t(nf) * size_t(dim) * size_t(dim + 1) * size_t(nq));
    SERAC_B200_CHECK(sb_geometric_factors(int(fg), mesh.nodes_space.order, Q, G_mesh, nodes_dev, faces_dev, nf, X_q, J_q,
                                          nullptr));
    sb_integral_t I = nullptr;
    SERAC_B200_CHECK(sb_integral_create(table, &integrand, face_ids ? local.data() : nullptr, nf, X_q, J_q, &I));
    sb_restriction_t test_faces = faces_of(*test_space_);
    mesh_restrictions_.push_back(test_faces);
    const int        which[] = {args..., 0};
    sb_restriction_t trial_faces[sizeof...(args) + 1];
    for (size_t i = 0; i < sizeof...(args); i++) {
      trial_faces[i] = faces_of(*trial_space_[size_t(which[i])]);
      mesh_restrictions_.push_back(trial_faces[i]);
    }
    SERAC_B200_CHECK(sb_functional_add_boundary_integral(F_, I, which, test_faces, trial_faces));
    integrals_.push_back(I);
  }

  template <typename T>
  T* device_alloc(size_t n)
  {
    T* p = nullptr;
    SERAC_B200_CUDA(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
    device_buffers_.push_back(p);
    return p;
  }
  template <typename T>
  T* device_copy(const T* host, size_t n)
  {
    T* p = device_alloc<T>(n);
    if (n) SERAC_B200_CUDA(cudaMemcpy(p, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    return p;
  }

  const FiniteElementSpace*                                test_space_;
  std::array<const FiniteElementSpace*, num_trial_spaces> trial_space_;
  sb_restriction_t                                         G_test_ = nullptr;
  std::array<sb_restriction_t, num_trial_spaces>           G_trial_{};
  std::vector<sb_restriction_t>                            G_test_more_;   ///< mixed meshes: the further geometries
  std::vector<std::array<sb_restriction_t, num_trial_spaces>> G_trial_more_;
  sb_functional_t                                          F_ = nullptr;
  std::vector<sb_integral_t>                               integrals_;
  std::vector<sb_restriction_t>                            mesh_restrictions_;
  std::vector<sb_csr_t>                                    csr_;
  std::vector<void*>                                       device_buffers_;
  std::vector<std::shared_ptr<void>>                       qdata_keepalive_;
  Vector                                                   output_L_;
  std::vector<Gradient>                                    grad_;
  bool                                                     update_qdata_ = false;
};

/// the QOI "test space" on the mesh of `first` (F/detail/qoi.inl): order 0, one component, ONE node that every element and
/// every boundary face refers to, with the same blocks of elements / faces as the space it is modelled on
inline FiniteElementSpace make_qoi_space(const FiniteElementSpace& first)
{
  FiniteElementSpace q{first.geometry, 0, 1, Ordering::byVDIM, 1, first.num_elements,
                       std::vector<int>(size_t(first.num_elements), 0)};
  q.num_boundary_faces = first.num_boundary_faces;
  q.native_boundary_face_dofs.assign(size_t(first.num_boundary_faces), 0);
  q.element_ids            = first.element_ids;
  q.boundary_face_geometry = first.boundary_face_geometry;
  q.boundary_face_ids      = first.boundary_face_ids;
  for (const auto& b : first.more_elements)
    q.more_elements.push_back(GeometryBlock{b.geometry, b.ids, std::vector<int>(b.ids.size(), 0)});
  for (const auto& b : first.more_boundary_faces)
    q.more_boundary_faces.push_back(GeometryBlock{b.geometry, b.ids, std::vector<int>(b.ids.size(), 0)});
  return q;
}

/// Quantity of interest, `Functional<double(trials...)>` (functional_qoi.inl:1-470): the same machinery with the QOI
/// "test space" (one total per element, summed by the scatter).  operator() returns the scalar (or tuple{scalar,
/// Gradient&}); Gradient::operator()(dx) is the directional derivative (a length-1 Vector), assemble(Gradient&) the
/// gradient as a 1 x n CsrMatrix and gradient_vector(Gradient&) as a dense vector (the reference returns a
/// HypreParVector, functional_qoi.inl:411-455).
template <typename... trials>
class Functional<double(trials...)> {
  using Impl                            = Functional<QOI(trials...)>;
  static constexpr int num_trial_spaces = int(sizeof...(trials));

 public:
  using Gradient = typename Impl::Gradient;

  explicit Functional(std::array<const FiniteElementSpace*, num_trial_spaces> trial_fes)
  {
    qoi_space_ = make_qoi_space(*trial_fes[0]);
    impl_      = std::make_unique<Impl>(&qoi_space_, trial_fes);
  }

  template <int dim, int... args, typename Integrand, typename StateType = Nothing>
  void AddDomainIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, Domain& domain,
                         std::shared_ptr<QuadratureData<StateType>> qdata = NoQData)
  {
    impl_->AddDomainIntegral(d, a, integrand, domain, qdata);
  }

  /// functional_qoi.inl:196-215; the q-function returns the scalar integrand per unit area of the boundary
  template <int dim, int... args, typename Integrand>
  void AddBoundaryIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, const Mesh& mesh)
  {
    impl_->AddBoundaryIntegral(d, a, integrand, mesh);
  }
  template <int dim, int... args, typename Integrand>
  void AddBoundaryIntegral(Dimension<dim> d, DependsOn<args...> a, const Integrand& integrand, const Domain& domain)
  {
    impl_->AddBoundaryIntegral(d, a, integrand, domain);
  }

  template <typename... T>
  auto operator()(double t, const T&... args)
  {
    constexpr bool differentiated = (std::is_same_v<T, differentiate_wrt_this> || ...);
    if constexpr (differentiated) {
      auto r = (*impl_)(t, args...);
      return std::tuple<double, Gradient&>(std::get<0>(r).to_host()[0], std::get<1>(r));
    } else {
      return (*impl_)(t, args...).to_host()[0];
    }
  }

  void updateQdata(bool update_flag) { impl_->updateQdata(update_flag); }

 private:
  FiniteElementSpace    qoi_space_;
  std::unique_ptr<Impl> impl_;
};

/// dense gradient of a quantity of interest from its 1 x n matrix
template <typename G>
std::vector<double> gradient_vector(G& g)
{
  auto                K = assemble(g);
  std::vector<double> v(size_t(K->num_cols), 0.0);
  for (size_t k = 0; k < K->col_ind.size(); k++) v[size_t(K->col_ind[k])] = K->values[k];
  return v;
}

}  // namespace core_api

#ifndef SERAC_B200_MFEM_API
using namespace core_api;
#endif

}  // namespace serac
