/* serac_b200 C ABI -- the drop-in boundary for the serac::Functional evaluation hot path on B200 (sm_100a).
 *
 * Every entry point replaces one piece of the reference's Functional machinery (file:line cited per function; paths
 * are relative to the reference checkout, F/ = src/serac/numerics/functional/).  Plain pointers and sizes only: no
 * C++ or torch types cross this boundary.  Pointers named *_dev are CUDA device pointers; *_host are host pointers.
 * All functions return 0 on success and a non-zero code on failure; sb_last_error() describes the last failure on the
 * calling thread.  (The reference has no return codes -- it aborts through SLIC_ERROR_ROOT, F/functional.hpp:77-98 --
 * so a C++ wrapper turns non-zero into the same abort.)  `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Work is enqueued asynchronously unless a host buffer is involved.
 *
 * There is no CPU fallback: if no CUDA device is usable every compute entry point fails with SB_ERR_CUDA.
 */
#ifndef SERAC_B200_H
#define SERAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_INVALID 1 /* bad argument / unsupported combination */
#define SB_ERR_CUDA 2    /* CUDA runtime error (text in sb_last_error) */
#define SB_ERR_NOT_FOUND 3 /* unknown named integrand */

/* geometry ids = mfem::Geometry::Type values */
#define SB_GEOM_SEGMENT 1
#define SB_GEOM_TRIANGLE 2
#define SB_GEOM_SQUARE 3
#define SB_GEOM_TETRAHEDRON 4
#define SB_GEOM_CUBE 5

/* mfem::Ordering::Type values (F/element_restriction.cpp:425-432) */
#define SB_ORDERING_BYNODES 0
#define SB_ORDERING_BYVDIM 1

#define SB_MAX_TRIALS 6
#define SB_NO_DIFFERENTIATION (-1)

typedef struct sb_restriction_s* sb_restriction_t;
typedef struct sb_integral_s*    sb_integral_t;
typedef struct sb_csr_s*         sb_csr_t;
typedef struct sb_functional_s*  sb_functional_t;
typedef struct sb_halo_s*        sb_halo_t;

int         sb_version(void);
const char* sb_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches) */
long long sb_launch_count(void);
void      sb_launch_count_reset(void);
void      sb_count_launch(void); /* called by kernel-table launchers (serac_b200/kernel_table.cuh) */

/* ---------------------------------------------------------------------------------------------------------------
 * ElementRestriction  (F/element_restriction.hpp:150-215, F/element_restriction.cpp:392-467)
 * ------------------------------------------------------------------------------------------------------------- */

/* Build dof_info(e,k) for H1 elements from the host library's native element dofs: applies the native ->
 * lexicographic permutation of the tensor-product / simplex element, F/element_restriction.cpp:7-65,216-277.
 * native_dofs_host: [num_elements][ndof] in mfem FiniteElementSpace::GetElementDofs order.
 * dof_info_host (out): [num_elements][ndof] packed 64-bit DoF words (F/element_restriction.hpp:31-80:
 * bits 0-47 index, 60-62 orientation, 63 sign). */
int sb_h1_lexicographic_dof_info(int geom, int p, const int* native_dofs_host, int64_t num_elements,
                                 uint64_t* dof_info_host);
/* Face restrictions of DISCONTINUOUS (L2) spaces, FaceType::BOUNDARY and FaceType::INTERIOR
 * (F/element_restriction.cpp:66-215, 304-355): a DG space has no face dofs of its own -- the dofs "of" a face are those of
 * an adjacent element that lie on it.  One entry per (face, adjacent element) pair: elem_geom[i] the element's geometry,
 * elem_dofs_host + elem_dofs_offset[i] its dofs (lexicographic, which is what mfem hands out for L2 spaces), side[i] the
 * local index of the face among the element's sides (mfem's GetElementFaces / GetElementEdges order) and orientation[i]
 * of the face relative to the element (mfem's face orientations; for edges 0 = as is, 1 = reversed; the reference passes 0
 * for the boundary edges of 2D meshes, :340-346).  dof_info_host (out): [num_pairs][nodes per face] DoF words in the
 * face's lexicographic order.  A boundary restriction has one pair per face; an interior one has two (the two elements of
 * a face one after the other), i.e. 2 x nodes-per-face columns per face.  Host only: no device is needed. */
int sb_l2_face_dof_info(int face_geom, int p, int64_t num_pairs, const int* elem_geom, const int* elem_dofs_host,
                        const int64_t* elem_dofs_offset, const int* side, const int* orientation, uint64_t* dof_info_host);
/* the permutation itself: lex_perm[k] = native position of lexicographic node k (ndof entries) */
int sb_h1_lexicographic_permutation(int geom, int p, int* lex_perm_host);

/* ElementRestriction ctor (F/element_restriction.cpp:392-404): takes ownership of a device copy of the table and
 * builds the transposed (L-dof -> E-slots) CSR used by the deterministic scatter. */
int sb_restriction_create(const uint64_t* dof_info_host, int64_t num_elements, int nodes_per_elem, int components,
                          int64_t num_nodes, int ordering, sb_restriction_t* out);
int sb_restriction_destroy(sb_restriction_t r);
int64_t sb_restriction_esize(sb_restriction_t r); /* ESize(), F/element_restriction.cpp:421 */
int64_t sb_restriction_lsize(sb_restriction_t r); /* LSize(), F/element_restriction.cpp:423 */
int64_t sb_restriction_num_elements(sb_restriction_t r); /* ElementRestriction::num_elements, F/element_restriction.hpp:196 */
/* device int32 table [num_elements][nodes_per_elem] of scalar node ids (what the fused kernels read) */
int sb_restriction_table_dev(sb_restriction_t r, const int** table_dev);
/* Gather L -> E (F/element_restriction.cpp:443-454): E[(e*c + comp)*npe + k] = L[vdof(dof_info(e,k), comp)] */
int sb_restriction_gather(sb_restriction_t r, const double* L_dev, double* E_dev, void* stream);
/* ScatterAdd E -> L (F/element_restriction.cpp:456-467): L[vdof] += E[...]; deterministic, no atomics: each L entry
 * sums its E slots in ascending slot order, which is the reference's element-major order, so results are bit-exact. */
int sb_restriction_scatter_add(sb_restriction_t r, const double* E_dev, double* L_dev, void* stream);
/* same, but L = sum (overwrites L instead of accumulating; saves the zero-fill of output_L, F/functional.hpp:456) */
int sb_restriction_scatter(sb_restriction_t r, const double* E_dev, double* L_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * GeometricFactors  (F/geometric_factors.cpp:16-138)
 * X_q_dev: [ne][sdim][nq]; J_q_dev: [ne][gdim][sdim][nq] with J(j,i,q) = dX_i/dxi_j.
 * sdim = the restriction's component count: gdim for domain elements, gdim + 1 for the boundary faces of a boundary
 * integral (the FaceType::BOUNDARY constructor, F/geometric_factors.cpp:140-220; pass the face restriction of the mesh
 * nodes space, built from fes->GetFaceDofs with geom = the face geometry).
 * nodes_L_dev: mesh nodes L-vector (vdim = spatial dim) restricted by `mesh_restriction`.
 * elements_dev: the domain's element ids (NULL = all elements 0..ne-1).
 * ------------------------------------------------------------------------------------------------------------- */
int sb_geometric_factors(int geom, int mesh_order, int q, sb_restriction_t mesh_restriction,
                         const double* nodes_L_dev, const int* elements_dev, int64_t num_elements, double* X_q_dev,
                         double* J_q_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Integral  (F/integral.hpp:24-168): one q-function on one domain, with its three kernel families
 *   evaluation / evaluation_with_AD   F/domain_integral_kernels.hpp:152-220
 *   jvp                               F/domain_integral_kernels.hpp:277-304
 *   element_gradient                  F/domain_integral_kernels.hpp:327-361
 * A q-function is compiled code, so an integral is created from a *kernel table*: either one of the named tables
 * compiled into this library (sb_integral_create_named; list with sb_named_integrand_*) or a table instantiated in
 * the caller's own nvcc translation unit from serac_b200/kernel_table.cuh (sb_integral_create_from_table) -- that is
 * what the C++ serac::Functional wrapper does for user-supplied functors.
 * ------------------------------------------------------------------------------------------------------------- */

/* device-side description of one trial field for a launch */
typedef struct {
  const double* L_dev;     /* L-vector */
  const int*    table_dev; /* restriction table of this field's space */
  int64_t       node_stride;
  int64_t       comp_stride;
} sb_field_arg;

typedef struct {
  sb_field_arg  trial[SB_MAX_TRIALS];
  const double* X_q_dev;
  const double* J_q_dev;
  const int*    elements_dev;
  int           num_elements;
  double*       r_E_dev;
  double*       dqf_dev;
  void*         qdata_dev;
  int           update_state;
  double        t;
  int           first_element;  /* evaluate only positions [first_element, num_elements) of the domain's element list
                                   (0 = from the start): lets a caller evaluate the elements that need no halo data while
                                   the exchange is in flight, the rest afterwards (sb_functional_evaluate_begin / _end) */
  int           r_E_node_major; /* 0: r_E is [element][comp][node] (the reference's E-vector, F/element_restriction.cpp:443);
                                   1: [element][node][comp] -- internal to sb_functional_*: the components of a node are
                                   contiguous, so the deterministic scatter reads whole 32-byte sectors */
} sb_eval_args;

typedef struct {
  sb_field_arg  trial;
  const int*    elements_dev;
  int           num_elements;
  const double* dqf_dev;
  double*       r_E_dev;
  int           r_E_node_major; /* see sb_eval_args */
} sb_jvp_args;

typedef struct {
  const int*    elements_dev;
  int           num_elements;
  const double* dqf_dev;
  double*       K_e_dev;
  int           K_e_node_major; /* 0: K_e is [element][trial vdof][test vdof] with vdof = comp * nodes + node (the
                                   reference's, F/functional.hpp:599-601); 1: [element][test node][test comp][trial node]
                                   [trial comp] -- internal to sb_functional_assemble: the entries one CSR row gathers
                                   from one element are contiguous */
  int           overwrite;      /* 1: K_e = ... instead of K_e += ... (first integral of an assembly: no memset, no read) */
} sb_elem_grad_args;

/* Fused element-stiffness -> CSR assembly (serac_b200/assemble_rows.cuh): the rows of the matrix are the unit of work.
 * The index tables come from the assembly plan sb_functional_assemble builds once per (integral, argument); they are
 * opaque to the caller (layouts: AsmGroup / AsmUnit / AsmNode in assemble_rows.cuh). */
typedef struct {
  const double*         dqf_dev;
  const void*           groups_dev;
  const void*           units_dev;
  const void*           nodes_dev;
  const unsigned short* pos_dev;
  const int*            row_ptr_dev;
  double*               values_dev;
  int64_t               row_node_stride, row_comp_stride; /* test vdof = node * row_node_stride + comp * row_comp_stride */
  int                   col_byvdim;                       /* trial space ordering (SB_ORDERING_*) */
  int                   accumulate;                       /* 0: values = rows, 1: values += rows */
  int                   num_groups;
  int                   acc_doubles;                      /* shared-memory accumulator the largest group needs */
} sb_assemble_rows_args;

typedef int (*sb_eval_launcher)(const sb_eval_args*, const void* qf_blob, void* stream);
typedef int (*sb_jvp_launcher)(const sb_jvp_args*, void* stream);
typedef int (*sb_elem_grad_launcher)(const sb_elem_grad_args*, void* stream);
typedef int (*sb_assemble_rows_launcher)(const sb_assemble_rows_args*, void* stream);

typedef struct {
  int    geom;       /* SB_GEOM_* */
  int    dim;        /* geometry dimension */
  int    q;          /* quadrature points per direction (Q = max order + 1, F/functional.hpp:193) */
  int    num_trials;
  int    test_order, test_components;
  int    trial_order[SB_MAX_TRIALS], trial_components[SB_MAX_TRIALS];
  size_t qf_bytes;    /* sizeof the q-function functor (copied by value into the kernel parameters) */
  size_t state_bytes; /* sizeof the quadrature-point state; 0 = no state */
  int    elements_per_block; /* elements per CTA of the streaming kernels; derivative buffers are padded to a multiple */
  sb_eval_launcher      evaluate;                        /* no differentiation */
  sb_eval_launcher      evaluate_ad[SB_MAX_TRIALS];      /* + derivative store w.r.t. trial i */
  sb_jvp_launcher       jvp[SB_MAX_TRIALS];
  sb_elem_grad_launcher element_gradient[SB_MAX_TRIALS];
  int derivative_mask[SB_MAX_TRIALS];    /* bit b: block b of {d00,d01,d10,d11} present */
  int derivative_entries[SB_MAX_TRIALS]; /* doubles stored per quadrature point */
  int spatial_dim; /* == dim for domain integrals; dim + 1 for boundary integrals (geom = the face geometry) */
  /* fused assembly w.r.t. trial i; NULL where it is not generated (2-D, simplex and boundary integrals, structurally zero
   * derivatives): sb_functional_assemble then goes through element matrices (element_gradient + sb_csr_accumulate) */
  sb_assemble_rows_launcher assemble_rows[SB_MAX_TRIALS];
  int assemble_rows_fixed_doubles[SB_MAX_TRIALS]; /* shared memory the kernel needs besides the accumulator */
  int assemble_rows_trial_blocks[SB_MAX_TRIALS];  /* blocks the trial components are cut into (one item per block) */
  /* stored derivatives are laid out [tile][entry][derivative_tile_elements x quadrature points]: 1 for hexahedra with
   * q >= 3 (every element its own tile), elements_per_block otherwise (sb_integral_export_derivatives converts back) */
  int derivative_tile_elements;
  /* d flux / d grad stored as the upper triangle of a symmetric matrix (bit 16 of derivative_mask; an extension, see
   * DESIGN.md 4d): derivative_entries counts what is stored, derivative_full_entries what the reference would store */
  int derivative_full_entries[SB_MAX_TRIALS];
  /* q-functions with run-time flags (geometric nonlinearity) may store differently per flag value: sb_integral_create
   * calls this on its own copy of the table with the functor it was given (NULL = nothing depends on the functor) */
  void (*specialize)(void* table, const void* qf_blob);
} sb_kernel_table;

int         sb_named_integrand_count(void);
const char* sb_named_integrand_name(int i);
/* look up a compiled-in kernel table; params are the functor's constructor arguments (documented per name in
 * INTEGRATION.md, e.g. "solid_neohookean_hex_p2": {density, K, G, geometric_nonlinearity}) */
int sb_named_kernel_table(const char* name, const sb_kernel_table** table);
int sb_named_qf_blob(const char* name, const double* params, int num_params, void* qf_blob, size_t qf_blob_capacity);

/* MakeDomainIntegral (F/integral.hpp:238-260).  elements_host: domain element ids (NULL = EntireDomain).
 * X_q_dev / J_q_dev: geometric factors for those elements (sb_geometric_factors); the integral keeps the pointers,
 * the caller keeps ownership.  Derivative buffers are allocated on first use (F/integral.hpp:205-221). */
int sb_integral_create(const sb_kernel_table* table, const void* qf_blob, const int* elements_host,
                       int64_t num_elements, const double* X_q_dev, const double* J_q_dev, sb_integral_t* out);
int sb_integral_destroy(sb_integral_t I);
/* attach quadrature-point state, [num_elements][nq] AoS of the state struct (F/quadrature_data.hpp:112-142) */
int sb_integral_set_qdata(sb_integral_t I, void* qdata_dev);
/* Integral::Mult (F/integral.hpp:65-81): r_E (this geometry's E-vector block of the test space) is overwritten for the
 * domain's elements.  wrt = SB_NO_DIFFERENTIATION or the trial index whose derivative is stored for later
 * sb_integral_jvp / sb_integral_element_gradient calls. */
int sb_integral_evaluate(sb_integral_t I, double t, const sb_field_arg* trials, double* r_E_dev, int wrt,
                         int update_state, void* stream);
/* Integral::GradientMult (F/integral.hpp:93-103) */
int sb_integral_jvp(sb_integral_t I, int wrt, const sb_field_arg* du, double* dr_E_dev, void* stream);
/* The same three operations with the E-vector block laid out [element][node][comp] (node_major != 0) instead of the
 * reference's [element][comp][node]: the layout sb_functional_* uses internally, exposed so that the kernels it runs can
 * be timed and tested one by one.  node_major == 0 is identical to the plain calls. */
int sb_integral_evaluate_layout(sb_integral_t I, double t, const sb_field_arg* trials, double* r_E_dev, int wrt,
                                int update_state, int node_major, void* stream);
int sb_integral_jvp_layout(sb_integral_t I, int wrt, const sb_field_arg* du, double* dr_E_dev, int node_major,
                           void* stream);
int sb_restriction_scatter_layout(sb_restriction_t r, const double* E_dev, double* L_dev, int accumulate,
                                  int node_major, void* stream);
/* Integral::ComputeElementGradients (F/integral.hpp:112-121): K_e_dev [elements of geometry][trial vdofs][test vdofs]
 * is accumulated into (+=), as the reference does. */
int sb_integral_element_gradient(sb_integral_t I, int wrt, double* K_e_dev, void* stream);
/* stored derivative w.r.t. `wrt`, converted to a dense reference-ordered array for parity checks:
 * out_dev[e][q][block entries in d00,d01,d10,d11 order, row-major, absent blocks skipped] */
int sb_integral_export_derivatives(sb_integral_t I, int wrt, double* out_dev, void* stream);
int sb_integral_derivative_entries(sb_integral_t I, int wrt);        /* per point, in the reference's (full) layout */
int sb_integral_derivative_stored_entries(sb_integral_t I, int wrt); /* per point, as stored (symmetric tangents: fewer) */
int sb_integral_derivative_mask(sb_integral_t I, int wrt);

/* ---------------------------------------------------------------------------------------------------------------
 * Sparsity pattern + assembly  (GradientAssemblyLookupTables, F/dof_numbering.hpp:253-380;
 *                               Gradient::assemble, F/functional.hpp:569-652)
 * ------------------------------------------------------------------------------------------------------------- */
/* lookup_tables.init(test, trial): sorted unique (row, col) pairs of all elements -> row_ptr / col_ind (bit-exact with
 * the reference).  Built at node level -- for every test node the sorted union of the trial nodes of its elements -- and
 * expanded by components on the device, so its cost is the number of NODE pairs of the element matrices, not of their
 * entries, and nothing limits the number of element-matrix entries (the reference hashes every entry,
 * F/dof_numbering.hpp:290-351); nnz < 2^31 as in mfem::SparseMatrix. */
int     sb_csr_create(sb_restriction_t test, sb_restriction_t trial, sb_csr_t* out);
/* Meshes of mixed element geometry: BlockElementRestriction keeps one restriction per geometry
 * (F/element_restriction.cpp:471-536) and lookup_tables.init walks all of them (F/dof_numbering.hpp:290-351): the
 * pattern is the union over the blocks (tests[b], trials[b]): the restrictions of the two spaces on the elements of
 * geometry b.  All blocks of a space share its num_nodes, components and ordering. */
int     sb_csr_create_blocks(int num_blocks, const sb_restriction_t* tests, const sb_restriction_t* trials, sb_csr_t* out);
/* The pattern of A^T as a pattern object of its own (rows = trial vdofs), and its values from those of A: what
 * HypreParMatrix::Transpose() gives the adjoint solve and MultTranspose the sensitivities
 * (src/serac/physics/solid_mechanics.hpp:1362-1371, 1404-1424, 1466-1501).  Everything that takes a (pattern, values)
 * pair -- sb_csr_spmv, the elimination of essential dofs, the Krylov solvers of serac_b200_solvers.h -- works on the
 * transposed pair; column indices are ascending within a row, as in every pattern of this library. */
int     sb_csr_transpose(sb_csr_t A, sb_csr_t* A_transposed_out);
int     sb_csr_transpose_values(sb_csr_t A_transposed, const double* values_dev, double* values_transposed_dev, void* stream);
/* The same pattern computed on the host from raw restriction tables ([num_elements][npe] node ids): what
 * sb_csr_create runs before uploading, callable without a device (parity tests of the pattern on CPU-only machines).
 * Call with col_ind_host == NULL to get nnz and row_ptr first. */
int     sb_csr_pattern_host(const int* test_table, int test_npe, int test_components, int64_t test_num_nodes,
                            int test_ordering, const int* trial_table, int trial_npe, int trial_components,
                            int64_t trial_num_nodes, int trial_ordering, int64_t num_elements, int* row_ptr_host,
                            int* col_ind_host, int64_t* nnz_out);
/* Test hook, host only: walks the index tables of the fused assembly (groups of rows, units in element order, row
 * positions, accumulator offsets, row write-out -- what assemble_rows_kernel walks on the device) over element matrices
 * supplied by the caller.  It checks tables, not arithmetic; nothing in the library calls it.
 * K_e_host: [domain position][trial vdof][test vdof]; elements: the integral's domain (NULL = all), nd entries;
 * acc_cap: accumulator doubles a group may use (small values force groups to be cut); trial_blocks: as the kernel's
 * (sb_kernel_table::assemble_rows_trial_blocks). */
int     sb_debug_assembly_plan_host(const int* test_table, int test_npe, int test_components, int64_t test_num_nodes,
                                    int test_ordering, const int* trial_table, int trial_npe, int trial_components,
                                    int64_t trial_num_nodes, int trial_ordering, int64_t num_elements,
                                    const int* elements, int64_t nd, int acc_cap, int trial_blocks,
                                    const double* K_e_host,
                                    const int* row_ptr_host, double* values_host, int* num_groups_out);
int     sb_csr_destroy(sb_csr_t A);
int64_t sb_csr_nnz(sb_csr_t A);
int64_t sb_csr_num_rows(sb_csr_t A);
int64_t sb_csr_num_cols(sb_csr_t A);
int     sb_csr_pattern_dev(sb_csr_t A, const int** row_ptr_dev, const int** col_ind_dev);
int     sb_csr_pattern_to_host(sb_csr_t A, int* row_ptr_host, int* col_ind_host);
/* values[nz] = sum over the contributing K_e entries in ascending (element, i, j) order (deterministic gather;
 * F/functional.hpp:610-644 does the same sums with one hash lookup per entry). values_dev is overwritten.
 * K_e_dev: [elements][trial vdofs][test vdofs] (the reference's layout); needs < 2^32 element-matrix entries.  This is
 * the generic route (any geometry); sb_functional_assemble does not materialise K_e where a fused kernel exists. */
int sb_csr_assemble(sb_csr_t A, const double* K_e_dev, double* values_dev, void* stream);
/* values[nz] += element matrices living on another pair of restrictions of the same spaces -- the boundary faces
 * (F/functional.hpp:610-644 loops over Domain::Type::BoundaryElements after Elements with the same lookup table).
 * K_e_dev: [faces][trial vdofs][test vdofs]; ordered (ascending face, i, j) per nonzero, no atomics.  Fails if a face
 * dof pair is not in the pattern (the reference's nz_LUT lookup would throw). */
int sb_csr_accumulate(sb_csr_t A, sb_restriction_t test_faces, sb_restriction_t trial_faces, const double* K_e_dev,
                      double* values_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Functional orchestration  (F/functional.hpp:406-513): P -> gather -> integrals -> scatter-add -> P^T at one rank
 * (P = identity, mfem/fem/pfespace.cpp:1170-1173) or with a halo exchange object attached.
 * ------------------------------------------------------------------------------------------------------------- */
int sb_functional_create(sb_restriction_t test, int num_trials, const sb_restriction_t* trials, sb_functional_t* out);
int sb_functional_destroy(sb_functional_t F);
/* AddDomainIntegral (F/functional.hpp:286-317): trial_indices[i] = which Functional argument feeds the integral's
 * i-th trial (the DependsOn<...> list) */
int sb_functional_add_integral(sb_functional_t F, sb_integral_t I, const int* trial_indices);
/* AddBoundaryIntegral (F/functional.hpp:328-354).  The integral was created from a boundary kernel table
 * (spatial_dim == dim + 1) with the geometric factors of the boundary faces.  test_faces / trial_faces[i]: the
 * FaceType::BOUNDARY element restrictions (F/element_restriction.cpp:279-388,406-419) of the test space and of the space
 * feeding the integral's i-th trial: dof_info from sb_h1_lexicographic_dof_info(face geometry, p, fes->GetFaceDofs ...). */
int sb_functional_add_boundary_integral(sb_functional_t F, sb_integral_t I, const int* trial_indices,
                                        sb_restriction_t test_faces, const sb_restriction_t* trial_faces);
/* An integral on its own block of elements: the elements of ONE geometry of a mixed mesh (the reference keeps a kernel
 * per geometry, F/integral.hpp:254-257, and a restriction per geometry, F/element_restriction.cpp:471-536).  test_block /
 * trial_blocks[i]: the restrictions of the test space / of the space feeding the integral's i-th trial on those elements.
 * sb_functional_assemble then needs a pattern that covers every block (sb_csr_create_blocks). */
int sb_functional_add_integral_on(sb_functional_t F, sb_integral_t I, const int* trial_indices,
                                  sb_restriction_t test_block, const sb_restriction_t* trial_blocks);
/* operator()(t, args...) (F/functional.hpp:446-513).  inputs: one L-vector per trial space.  *_on_host:
 *   SB_DEVICE (0)      device pointers, work is enqueued on `stream`
 *   SB_HOST_SYNC (1)   host buffers, copied to / from the device inside the call; the result is valid on return
 *   SB_HOST_ASYNC (2)  host buffers (pinned), copied on the Functional's own upload / download streams: the call returns
 *                      at once; an input copy overlaps whatever the caller's stream is still running (the direction of
 *                      the Jacobian action goes up during the evaluation), an output copy overlaps the next call's
 *                      copies and kernels.  Results -- and the right to modify the input buffers -- are valid after
 *                      sb_functional_wait(). */
#define SB_DEVICE 0
#define SB_HOST_SYNC 1
#define SB_HOST_ASYNC 2
int sb_functional_wait(sb_functional_t F);
/* operator() in two halves, device buffers only: _begin evaluates, for every domain integral that covers the whole mesh,
 * the elements [0, split_element) -- with partitions whose halo-dependent elements are numbered last these need no data
 * from other ranks, so the caller issues the halo puts, calls _begin, waits for / unpacks the halo, and calls _end, which
 * evaluates the remaining elements (and every other integral) and scatters.  Same result as sb_functional_evaluate.
 * With split_element = the number of elements _begin is the evaluation kernel and _end the scatter alone: a caller that
 * wants the Jacobian action next can issue _end on a second stream (after an event recorded behind _begin) and
 * sb_functional_jvp on the first -- the action needs the stored derivatives, not the scattered residual, and uses an
 * E-vector of its own, so the residual's scatter runs under the (HBM-bound) action kernel. */
int sb_functional_evaluate_begin(sb_functional_t F, double t, int wrt, const double* const* inputs_dev, int update_state,
                                 int64_t split_element, void* stream);
int sb_functional_evaluate_end(sb_functional_t F, double* output_dev, void* stream);
int sb_functional_evaluate(sb_functional_t F, double t, int wrt, const double* const* inputs, int inputs_on_host,
                           double* output, int output_on_host, int update_state, void* stream);
/* Gradient::Mult (F/functional.hpp:406-432,556-559) for the derivative stored by the last evaluate(wrt) */
int sb_functional_jvp(sb_functional_t F, int wrt, const double* du, int du_on_host, double* dr, int dr_on_host,
                      void* stream);
/* assemble(Gradient&) up to the local CSR triple (F/functional.hpp:569-652); values_dev has sb_csr_nnz entries */
int sb_functional_assemble(sb_functional_t F, int wrt, sb_csr_t* pattern_inout, double* values, int values_on_host,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Shared-dof exchange P / P^T  (mfem::ConformingProlongationOperator, mfem/fem/pfespace.cpp:3599-3663; device variant
 * :3740-3933).  The library packs / unpacks; the transport (NCCL send/recv over NVLink) is driven by the caller with
 * the buffers returned here so that this library has no MPI / NCCL link dependency.
 * ------------------------------------------------------------------------------------------------------------- */
/* ltdof_host: for every local dof, the local true-dof index it copies from, or -1 if it is owned by another rank.
 * send_idx: true-dof ids packed for neighbours (concatenated, neighbour n owns [send_off[n], send_off[n+1]));
 * recv_idx: local dof ids filled from neighbours, same layout. */
int sb_halo_create(int64_t lsize, int64_t tsize, const int* ltdof_host, int num_neighbors, const int64_t* send_off,
                   const int* send_idx_host, const int64_t* recv_off, const int* recv_idx_host, sb_halo_t* out);
int sb_halo_destroy(sb_halo_t H);
int sb_halo_buffers(sb_halo_t H, double** send_buf_dev, double** recv_buf_dev);
/* use caller-owned staging buffers instead (each at least max(total send, total recv) doubles), e.g. buffers registered
 * with the transport or peer-mapped symmetric memory; the library stops using (and later frees) its own */
int sb_halo_set_buffers(sb_halo_t H, double* send_buf_dev, double* recv_buf_dev);
int64_t sb_halo_buffer_doubles(sb_halo_t H);
/* P: T -> L.  pack owner values for neighbours / copy owned / (after transport) unpack received values */
int sb_halo_P_pack(sb_halo_t H, const double* T_dev, void* stream);
int sb_halo_P_finish(sb_halo_t H, const double* T_dev, double* L_dev, void* stream);
/* P^T: L -> T.  pack non-owned contributions / copy owned / (after transport) add received, fixed neighbour order */
int sb_halo_Pt_pack(sb_halo_t H, const double* L_dev, void* stream);
int sb_halo_Pt_finish(sb_halo_t H, const double* L_dev, double* T_dev, void* stream);


/* Peer-memory transport: the whole exchange as two kernels over NVLink-mapped peer buffers -- no NCCL call on the path,
 * so a step can be captured in a CUDA graph and an exchange costs two launches (~5 us) instead of a p2p group.
 * Every rank owns a receive RING in memory its neighbours can map (CUDA IPC / symmetric memory; mapping is the caller's
 * job): 2 slots of `slot_doubles` doubles followed by 2 x SB_HALO_MAX_NEIGHBORS int64 arrival flags, i.e.
 * (2 * slot_doubles + 2 * SB_HALO_MAX_NEIGHBORS) 8-byte words, zero-initialised.  Round k uses slot k % 2: a sender can
 * be at most one round ahead of a receiver (it needs that receiver's round k - 1 data first), so two slots suffice.
 *   peer_ring[n]        neighbour n's ring as mapped in this process
 *   peer_P_off[n]       where this rank's P segment starts inside neighbour n's slot (= its recv_off for this rank)
 *   peer_Pt_off[n]      same for P^T (= its send_off for this rank)
 *   my_index_at_peer[n] this rank's position in neighbour n's neighbour list (selects the flag it raises there)
 * The put kernel writes straight into the neighbours' slots and raises their flags (release, system scope); the finish
 * kernel waits for its own flags (acquire) and unpacks.  A wait that exceeds ~2 s gives up and records an error that
 * sb_halo_peer_status reports (non-zero), instead of hanging the device. */
#define SB_HALO_MAX_NEIGHBORS 32
int sb_halo_enable_peer(sb_halo_t H, double* my_ring_dev, int64_t slot_doubles, double* const* peer_ring,
                        const int64_t* peer_P_off, const int64_t* peer_Pt_off, const int* my_index_at_peer);
int sb_halo_P_peer(sb_halo_t H, const double* T_dev, double* L_dev, void* stream);
int sb_halo_Pt_peer(sb_halo_t H, const double* L_dev, double* T_dev, void* stream);
/* the two halves of each peer exchange: issue the puts of several fields (one halo object per field in flight) before
 * the first finish so that the flag waits overlap, or run an exchange on a side stream next to compute kernels */
int sb_halo_P_peer_put(sb_halo_t H, const double* T_dev, void* stream);
int sb_halo_P_peer_finish(sb_halo_t H, const double* T_dev, double* L_dev, void* stream);
int sb_halo_Pt_peer_put(sb_halo_t H, const double* L_dev, void* stream);
int sb_halo_Pt_peer_finish(sb_halo_t H, const double* L_dev, double* T_dev, void* stream);
/* In-place exchanges for partitions numbered owned-first (ltdof[i] == i below tsize, -1 above; sb_halo_owned_first
 * tells): the T-vector is the first tsize entries of the L-vector, so P only fills the entries owned elsewhere and P^T
 * only adds into the shared owned entries -- nothing else of the vector is read or written.  Pair them with
 * sb_halo_P_peer_put(H, TL) / sb_halo_Pt_peer_put(H, TL) on the same buffer. */
int sb_halo_owned_first(sb_halo_t H);
int sb_halo_P_peer_finish_inplace(sb_halo_t H, double* TL_dev, void* stream);
int sb_halo_Pt_peer_finish_inplace(sb_halo_t H, double* TL_dev, void* stream);
/* P in place for up to 4 fields (one halo object each) with ONE put launch and ONE wait + unpack launch: what
 * Functional::operator() needs before its gathers (F/functional.hpp:460-470 applies P to every argument) */
int sb_halo_P_peer_inplace_multi(int num_fields, const sb_halo_t* H, double* const* TL_dev, void* stream);
/* P^T in place in one call.  sb_halo_P_peer_inplace_multi and sb_halo_Pt_peer_inplace run, by default, the LL protocol:
 * one kernel per round in which every double travels as two 8-byte {half, round tag} words that the receiver polls -- no
 * fence, no flag, one NVLink trip.  SB_HALO_LL=0 in the environment selects the flag protocol (put launch + wait/unpack
 * launch, or one fused launch with SB_HALO_FUSED=1); a halo object must stay with one protocol. */
int sb_halo_Pt_peer_inplace(sb_halo_t H, double* TL_dev, void* stream);
/* the two halves of the multi-field P: launch kernels that need no halo data (sb_functional_evaluate_begin) in between */
int sb_halo_P_peer_put_multi(int num_fields, const sb_halo_t* H, double* const* TL_dev, void* stream);
int sb_halo_P_peer_finish_inplace_multi(int num_fields, const sb_halo_t* H, double* const* TL_dev, void* stream);
int sb_halo_peer_status(sb_halo_t H, int* status_host);

#ifdef __cplusplus
}
#endif
#endif /* SERAC_B200_H */
