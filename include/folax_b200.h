/* folax_b200 -- C ABI of the B200 (sm_100a) finite-element assembly / physics-loss kernels.
 *
 * Drop-in boundary for the hot path of Neural-Mechanics-Lab/folax `fol/loss_functions`.
 * The reference's only FFI precedent is the XLA typed-FFI handler pair of
 *   fol/loss_functions/ffi_functions/kr_small_displacement_element.cc:293-334
 * (`compute_nodal_residuals`, `compute_elements`: stream from PlatformStream<cudaStream_t>,
 * device buffers coords (nn,3), connectivity (ne,a) S32, properties, solution -> lhs (ne,nd,nd),
 * rhs (ne,nd)).  Every entry point below keeps that shape: a stream, raw device pointers,
 * explicit sizes, outputs preallocated by the caller, nothing retained, nothing synchronised.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - connectivity / indices are int32 (the reference requires S32, ...cc:102, 320);
 *   - element DOF vector is node-major: local i = a*d + k  <->  global d*node + k
 *     (fol/loss_functions/fe_loss.py:163-164, 178-181);
 *   - return value 0 = success, <0 = error; fol_last_error() gives the message of the calling
 *     thread's last failure (the reference returns ffi::Error::InvalidArgument/Internal,
 *     ...cc:56, 218-223);
 *   - `dtype` selects the arithmetic type of all floating-point buffers of the call.
 */
#ifndef FOLAX_B200_H
#define FOLAX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fol_stream_t; /* cudaStream_t */

enum fol_dtype { FOL_F32 = 0, FOL_F64 = 1 };
/* fol/geometries/__init__.py:18-21 (fe_element_dict keys) */
enum fol_element { FOL_HEXAHEDRON = 0, FOL_QUAD = 1, FOL_TETRA = 2, FOL_TRIANGLE = 3 };
/* fol/loss_functions: mechanical.py, thermal.py, mechanical_neohooke.py, mechanical_elastoplasticity.py */
enum fol_physics { FOL_MECHANICAL = 0, FOL_THERMAL = 1, FOL_NEOHOOKE = 2, FOL_J2PLASTICITY = 3,
                   FOL_STVENANT = 4 /* mechanical_saint_venant.py */,
                   FOL_TRANSIENT_THERMAL = 5 /* transient_thermal.py */, FOL_ALLEN_CAHN = 6 /* phase_field.py */,
                   /* the *_AD.py variants (stiffness = jacfwd of the residual; element stage only, see
                    * csrc/assemble_ad_threads.cuh): state_out = per-element energy (ne) or NULL */
                   FOL_NEOHOOKE_AD = 7 /* mechanical_neohooke_AD.py */, FOL_STVENANT_AD = 8 /* mechanical_saint_venant_AD.py */ };

#define FOL_OK 0
#define FOL_ERR_INVALID (-1)
#define FOL_ERR_CUDA (-2)
#define FOL_ERR_UNSUPPORTED (-3)

#define FOL_NUM_PARAMS 12
/* material / loss parameters, always passed as double[FOL_NUM_PARAMS] on the HOST:
 *   [0] young_modulus  [1] poisson_ratio  [2..4] body force (mechanical.py:26 "body_foce")
 *   thermal: [5] beta  [6] c   (thermal.py:21-26)
 *   J2:      [5] yield_limit  [6] iso_hardening_parameter_1  [7] iso_hardening_param_2
 *            (mechanical_elastoplasticity.py:22-30)
 *   transient thermal: [5] beta [6] c [8] rho [9] cp [10] time_step;  Allen-Cahn: [10] dt [11] epsilon */

const char* fol_last_error(void);
int fol_version(void);
/* number of kernels launched by this library in the calling process (bench `gpu_launches`) */
int64_t fol_launch_count(void);

/* 1 (default): workloads with a tuned kernel (Hex8 elasticity f64, 2x2x2 rule) use it; 0: always the
 * generic kernel (A/B parity checks).  Returns the previous setting. */
int fol_set_tuned_kernels(int enable);

/* Persistent kernels normally fill every SM.  A margin of `ctas` thread blocks leaves room for
 * communication kernels (NCCL send/recv of the halo-DOF exchange) that must run concurrently.
 * Returns the previous setting. */
int fol_set_grid_margin(int ctas);

/* element table: nodes per element, spatial dim, Gauss points of integration order num_gp */
int fol_element_info(int element, int num_gp, int* nnode, int* dim, int* ngauss);
/* dofs per node of a physics on an element (1 for thermal, dim otherwise) */
int fol_dofs_per_node(int physics, int element);

/* ---- one-off integer plan kernels (bit-exact) ------------------------------------------- */

/* BCOO index pairs of ComputeElementJacobianIndices (fe_loss.py:178-184, 313-314):
 * indices[(e*nd*nd + i*nd + j)*2 + {0,1}] = (gdof(e,i), gdof(e,j)). */
int fol_bcoo_indices(fol_stream_t s, const int32_t* conn, int64_t ne, int nnode, int dofs_per_node,
                     int32_t* indices);

/* dir_flag[ndof] = 1 on Dirichlet dofs else 0 (BC_vector / mask_BC_vector, fe_loss.py:268-271) */
int fol_dirichlet_flags(fol_stream_t s, const int32_t* dirichlet_indices, int64_t n_dirichlet,
                        int64_t ndof, uint8_t* dir_flag);

/* node -> (element, local node) adjacency in CSR form, entries e*nnode+a sorted ascending:
 * the fixed summation order of the atomics-free residual gather (replaces the scatter-add of
 * fe_loss.py:301-306).  adj_ptr has nn+1 entries, adj has ne*nnode entries;
 * `work` needs nn int32 of scratch. */
int fol_node_adjacency(fol_stream_t s, const int32_t* conn, int64_t ne, int nnode, int64_t nn,
                       int32_t* adj_ptr, int32_t* adj, int32_t* work);

/* ---- residual + Jacobian assembly (fe_loss.py:264-318) ----------------------------------- */

/* Element stage: gather -> ComputeElement -> optional transpose -> Dirichlet row mask, writes
 *   ke_data[e*nd*nd + i*nd + j] = Ke'[i,j]            (the BCOO `data`, fe_loss.py:299)
 *   re_elem[e*nd + i]           = re'[i]              (masked element residuals)
 * state_in/state_out are the auxiliary input / output of the physics: J2: (ne, ngauss, 7|4) Gauss-point
 * history in / out; transient thermal: state_in = nodal heterogeneity k0 (nn); transient thermal and
 * Allen-Cahn: state_out = per-element energy (ne) or NULL, and (ctrl, u) = (current, next) nodal field.
 * Otherwise NULL.  u is the full dof vector (ndof), ctrl the nodal control field (nn). */
int fol_assemble_elements(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                          int transpose, int64_t ne, int64_t nn, const void* xyz,
                          const int32_t* conn, const void* ctrl, const void* u,
                          const uint8_t* dir_flag, const double* params_host, void* ke_data,
                          void* re_elem, const void* state_in, void* state_out);

/* Residual stage: R[d*n+k] = sum over adjacency (fixed order) of re_elem -- deterministic. */
int fol_residual_gather(fol_stream_t s, int dtype, int64_t nn, int nnode, int dofs_per_node,
                        const int32_t* adj_ptr, const int32_t* adj, const void* re_elem,
                        void* residual);

/* Duplicate-free CSR values from the BCOO data (hand-off to fol/solvers: fe_solver.py:71-72, 82).
 * The integer plan is built once per mesh by the host (folax_b200/csr_plan.py): node pairs (n, m) in
 * row-major sorted order; pair_ptr/contrib list, per pair, the contributing element entries
 * e*a*a + la*a + lb in ascending order (fixed summation order => deterministic values);
 * out_base[p] is the CSR position of entry (i=0, j=0) of pair p, row_stride[p] the distance between
 * consecutive dof rows of that node.  vals has d*d*npairs entries. */
int fol_csr_values(fol_stream_t s, int dtype, int64_t npairs, int dofs_per_node, int nnode,
                   const int32_t* pair_ptr, const int32_t* contrib, const int32_t* out_base,
                   const int32_t* row_stride, const void* ke_data, void* vals);

/* ---- batched physics loss + VJP (fe_loss.py:250-262 and its JAX-AD gradient) -------------- */

/* Matrix-free product with the Jacobian of fol_assemble_elements (what the reference's consumers do
 * with `BCOO @ vector`, fe_solver.py:61, and what a Krylov solver needs): per element
 *   ye_elem[e*nd + i] = sum_j Ke'[i,j] v[gdof(e,j)],  Ke' the (transposed, then) row-masked element
 * matrix of fe_loss.py:191-230 evaluated at (ctrl, u) -- Ke is formed in registers and never written.
 * fol_residual_gather(ye_elem) then gives y = J v (or J^T-variant v for transpose != 0) in the fixed
 * summation order of the residual.  state_in: as fol_assemble_elements (J2 history is read, not
 * advanced). */
int fol_apply_jacobian_elements(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                                int transpose, int64_t ne, int64_t nn, const void* xyz,
                                const int32_t* conn, const void* ctrl, const void* u,
                                const uint8_t* dir_flag, const double* params_host, const void* v,
                                void* ye_elem, const void* state_in);
/* The same products for a BATCH of samples in one launch (grid.y = sample): controls (nb, nn), dofs and v (nb, ndof)
 * -> ye_elem (nb, ne*nd); fol_residual_gather_batched sums them per sample in the residual's fixed order.  What the
 * nested VJP of the batched loss needs (the latent-code steps of meta_implicit_parametric_operator_learning.py:95-105
 * differentiate through jax.vmap(jax.grad(loss))).  Not for history-dependent (J2) elements. */
int fol_apply_jacobian_elements_batched(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                                        int transpose, int64_t ne, int64_t nn, int64_t nb, const void* xyz,
                                        const int32_t* conn, const void* controls, const void* dofs,
                                        const uint8_t* dir_flag, const double* params, const void* v,
                                        void* ye_elem);
int fol_residual_gather_batched(fol_stream_t s, int dtype, int64_t nn, int nnode, int dofs_per_node, int64_t nb,
                                int64_t ne, const int32_t* adj_ptr, const int32_t* adj, const void* re_elem,
                                void* residual);

/* per-element, per-Gauss-point geometry factors shared by all samples, SoA over elements:
 * geom[(g*(a*dim+1) + k)*ne + e] = grad N flattened (k < a*dim) | w*detJ (k = a*dim).   */
int fol_geometry_cache(fol_stream_t s, int dtype, int element, int num_gp, int64_t ne,
                       const void* xyz, const int32_t* conn, void* geom);
/* Same in the layout of `physics`: W = fol_geometry_width(physics, element) rows per Gauss point.  The
 * implicit-Euler scalar losses use the gradient convention of transient_thermal.py:57-58 /
 * phase_field.py:47-48; transient thermal adds one row, the nodal heterogeneity `aux` = k0 (nn)
 * interpolated to the point (it is mesh-resident, transient_thermal.py:98-99). */
int fol_geometry_width(int physics, int element);
int fol_geometry_cache_physics(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                               int64_t ne, const void* xyz, const int32_t* conn, const void* aux,
                               void* geom);

/* For every sample b (rows of ctrl (nb, nn) and u (nb, ndof), Dirichlet entries of u already
 * overwritten, fe_loss.py:255):
 *   grad_u[b]  = assembled, UN-masked residual R(u_b)      (= dE_b/du_b)
 *   grad_k[b]  = dE_b/dK_b                                  (NULL for mechanical: it is zero)
 *   energy[b]  = E_b = sum_e energy_e                       (before the exponent)
 * One fused kernel, deterministic (fixed-order sums, no atomics, no HBM scratch).  The integer
 * tile plan is built once per mesh by the host (folax_b200/energy_plan.py):
 *   tile_node_ptr (ntiles+1) / tile_nodes (nn): nodes grouped into compact tiles of <= 128 nodes
 *   tile_elem_ptr (ntiles+1) / tile_elems: the elements touching each tile
 *   adj_ptr (nn+1): node adjacency ranges (same order as fol_node_adjacency)
 *   adj_local: per adjacency entry (index of the element within the node's tile list)*nnode + a
 *   tile_lnode_ptr / tile_lnodes: per tile, the nodes its elements touch (the tile's own nodes first,
 *   in tile order); tile_conn (len(tile_elems), nnode): element nodes in that tile-local numbering
 *   ecap / lcap / ncap: the longest tile element list / local node list / owned node list.
 * Plans whose element lists fit one thread per element (ecap, ncap <= 160) and whose geometry factors
 * fit in registers run the pipelined kernel (csrc/energy2.cuh), the others energy_tile_kernel.
 * `work` is scratch of fol_energy_work_size(ntiles, nb) elements of the call's dtype. */
int64_t fol_energy_work_size(int64_t ntiles, int64_t nb);
int fol_energy_and_grads(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                         int64_t ne, int64_t nn, int64_t nb, const void* geom,
                         const int32_t* conn, const int32_t* adj_ptr, const int32_t* adj_local,
                         const int32_t* tile_node_ptr, const int32_t* tile_nodes,
                         const int32_t* tile_elem_ptr, const int32_t* tile_elems,
                         const int32_t* tile_conn, const int32_t* tile_lnode_ptr,
                         const int32_t* tile_lnodes, int64_t ntiles, int64_t ecap, int64_t lcap,
                         int64_t ncap, const void* ctrl, const void* u, const void* dir_values,
                         const uint8_t* dir_flag, double out_scale, const double* params_host,
                         void* grad_u, void* grad_k, void* energy, void* work);

/* Same call with facts about the mesh that the host plan established (FOL_MESH_* bits); the kernels may take a
 * cheaper route that gives the same results to rounding.  FOL_MESH_AFFINE: every element is a parallelogram /
 * parallelepiped image of the reference element (constant Jacobian), e.g. the structured meshes of
 * fol/tools/usefull_functions.py:213-258 -- the Quad4 thermal loss then keeps J^-1 and w detJ (5 values) per element
 * instead of 36 cached gradient values. */
#define FOL_MESH_AFFINE 1
int fol_energy_and_grads_flags(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                               int64_t ne, int64_t nn, int64_t nb, const void* geom,
                               const int32_t* conn, const int32_t* adj_ptr, const int32_t* adj_local,
                               const int32_t* tile_node_ptr, const int32_t* tile_nodes,
                               const int32_t* tile_elem_ptr, const int32_t* tile_elems,
                               const int32_t* tile_conn, const int32_t* tile_lnode_ptr,
                               const int32_t* tile_lnodes, int64_t ntiles, int64_t ecap, int64_t lcap,
                               int64_t ncap, const void* ctrl, const void* u, const void* dir_values,
                               const uint8_t* dir_flag, double out_scale, const double* params_host,
                               void* grad_u, void* grad_k, void* energy, void* work, int64_t mesh_flags);

/* loss tail: L = mean_b E_b^p, stats = (min, max, mean) of E_b^p, scale[b] = p E_b^(p-1)/nb.
 * out[0..3] = L, min, max, mean (device). */
int fol_loss_reduce(fol_stream_t s, int dtype, int64_t nb, double exponent, const void* energy,
                    void* out4, void* scale);

/* backward: grad_u[b,:] *= g*sc_b, zeroed at Dirichlet dofs; grad_k[b,:] *= g*sc_b, with
 * g = upstream * (*upstream_dev, a device scalar of the call's dtype, or 1 when NULL -- no host
 * synchronisation on the cotangent) and sc_b = scale[b], or 1 when `prescaled` (the forward call
 * already applied out_scale = scale: exponent 1).  prescaled and g == 1 is a no-op. */
int fol_scale_grads(fol_stream_t s, int dtype, int64_t nb, int64_t ndof, int64_t nn,
                    const void* scale, double upstream, const void* upstream_dev, int prescaled,
                    const uint8_t* dir_flag, void* grad_u, void* grad_k);

/* u[b, dirichlet_indices] = values (GetFullDofVector, fe_loss.py:91-92) or, when
 * per_sample != 0, values is (nb, n_dirichlet) (parametric boundary learning, :94-95).
 * load_factor multiplies the values (ApplyDirichletBCOnDofVector, :186-189). */
int fol_apply_dirichlet(fol_stream_t s, int dtype, int64_t nb, int64_t ndof,
                        const int32_t* dirichlet_indices, int64_t n_dirichlet, const void* values,
                        int per_sample, double load_factor, void* u);

/* ---- adjoint sensitivities of a response (fol/responses/fe_response.py) ------
 * The response is value = sum_e sum_g w detJ f(K_g, U_g) with a user formula f of the control and the dofs
 * at the Gauss point (fe_response.py:59-66, 91-124).  The formula itself is caller code: the host evaluates it
 * (and its partials) pointwise over the Gauss-point arrays these kernels produce and consume.
 * All element-major outputs are in the layout fol_residual_gather sums to the nodes (dofs_per_node = the
 * per-node width of the array: dofs for du_elem, 1 for dk_elem, 3 for dx_elem). */

/* k_gp[e*g + q] = N_q . ctrl[conn[e]],  u_gp[(k*ne + e)*g + q] = sum_a N_q[a] u[d*conn[e,a] + k]  (:112-115) */
int fol_gauss_interpolate(fol_stream_t s, int dtype, int element, int num_gp, int dofs_per_node, int64_t ne,
                          const int32_t* conn, const void* ctrl, const void* u, void* k_gp, void* u_gp);

/* From f_gp (ne, g) and the partials fk_gp (ne, g) = df/dK, fu_gp (d, ne, g) = df/dU[k]:
 *   value_elem[e] = sum_q w detJ f                                   (:116-124)
 *   du_elem[e, a*d+k] = sum_q w detJ N_a df/dU[k]                    (:126-138, the adjoint right-hand side)
 *   dk_elem[e, a]     = sum_q w detJ N_a df/dK                       (:140-152)
 *   dx_elem[e, a*3+k] = sum_q w f d(detJ)/dx_ak                      (:154-168; unused coordinates 0)
 * Any output (and the partial it needs) may be NULL. */
int fol_response_elements(fol_stream_t s, int dtype, int element, int num_gp, int dofs_per_node, int64_t ne,
                          const void* xyz, const int32_t* conn, const void* f_gp, const void* fk_gp,
                          const void* fu_gp, void* value_elem, void* du_elem, void* dk_elem, void* dx_elem);

/* dk_elem[e, a] (+)= adj_e^T d re/d ctrl_a,  dx_elem[e, a*3+k] (+)= adj_e^T d re/d x_ak, re = the element
 * residual of ComputeElement BEFORE the Dirichlet mask (fe_response.py:312-331, 424-442: jacrev there).
 * accumulate != 0 adds to the arrays (they hold the response part).  Mechanical, thermal, transient thermal
 * (aux = nodal k0, (ctrl, u) = (current, next) field) and Allen-Cahn: closed forms; Neo-Hooke and St-Venant:
 * closed-form geometry + dim^2 dual-number evaluations of the point law; J2 returns FOL_ERR_UNSUPPORTED.
 * aux: NULL unless the physics has an auxiliary nodal field. */
int fol_residual_adjoint_elements(fol_stream_t s, int dtype, int physics, int element, int num_gp, int accumulate,
                                  int64_t ne, const void* xyz, const int32_t* conn, const void* ctrl, const void* u,
                                  const void* adj, const void* aux, const double* params_host, void* dk_elem,
                                  void* dx_elem);
/* lam^T d(re)/dK for a BATCH of samples in one launch: controls (nb, nn), dofs and adjoint (nb, ndof) -> dk_elem (nb, ne*A) */
int fol_residual_adjoint_elements_batched(fol_stream_t s, int dtype, int physics, int element, int num_gp,
                                          int64_t ne, int64_t nn, int64_t nb, const void* xyz,
                                          const int32_t* conn, const void* controls, const void* dofs,
                                          const void* adjoint, const void* aux, const double* params,
                                          void* dk_elem);

/* energy_elem[e] = the element energy ComputeElement returns first (ComputeElementsEnergies, fe_loss.py:149-176):
 * u^T(Ke u - Fe) for the mechanical / thermal losses, sum_g w detJ psi for Neo-Hooke / St-Venant, the implicit-Euler
 * potentials of transient thermal (aux = nodal k0) and Allen-Cahn.  J2 and the AD variants: FOL_ERR_UNSUPPORTED
 * (their element stage returns the energies itself). */
int fol_element_energies(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne, const void* xyz,
                         const int32_t* conn, const void* ctrl, const void* u, const void* aux,
                         const double* params_host, void* energy_elem);

/* out[0] = sum of x[0..n) in a fixed order (one block): ComputeValue's jnp.sum, fe_response.py:216 */
int fol_sum(fol_stream_t s, int dtype, int64_t n, const void* x, void* out);

/* ---- device-resident linear algebra for fol/solvers (fe_solver.py:60-103) -----
 * The reference ships the BCOO to the host and lets SciPy / jax.scipy solve; with the Jacobian assembled and
 * de-duplicated on the device (fol_csr_values) the Krylov iteration can stay there.  Matrix layout: sliced
 * ELLPACK built once per mesh by the host from the CSR structure (folax_b200/sell_plan.py): slices of 32 rows,
 * entry k of row r at slice_ptr[r/32] + k*32 + r%32; padded entries have value 0 and column 0. */

/* dst[i] = src_index[i] >= 0 ? src[src_index[i]] : 0  (CSR values -> SELL values, CSR values -> diagonal) */
int fol_gather_values(fol_stream_t s, int dtype, int64_t n, const int32_t* src_index, const void* src, void* dst);
/* y = A x, one thread per row, per-row sums in CSR order (deterministic); x and y must not alias */
int fol_sell_spmv(fol_stream_t s, int dtype, int64_t nrows, const int64_t* slice_ptr, const int32_t* cols,
                  const void* vals, const void* x, void* y);
/* The same product for Jacobians with d = 2 or 3 dofs per node whose rows consist of runs of d consecutive dofs of a
 * neighbour node (what folax_b200/csr_plan.py builds): node_cols holds ONE node index per run -- at
 * slice_ptr[r/32]/d + q*32 + r%32 for run q of row r -- so a stored entry costs 8 + 4/d bytes instead of 12.  Same
 * values array, same summation order, bit-identical result. */
int fol_sell_spmv_block(fol_stream_t s, int dtype, int dofs_per_node, int64_t nrows, const int64_t* slice_ptr,
                        const int32_t* node_cols, const void* vals, const void* x, void* y);
/* op 0: out = a x + b y (y may be NULL when b == 0)   op 1: out = a x*y   op 2: out = a x/y;  out may alias x or y */
int fol_vec_op(fol_stream_t s, int dtype, int op, int64_t n, double a, const void* x, double b, const void* y,
               void* out);
/* BiCGSTAB with its scalars on the device (no host read inside an iteration).  `scalars` is an array of
 * fol_bicg_scalar_count() values of the call's dtype (layout: csrc/krylov_threads.cuh, enum BS_*): the dot products
 * are written into it by fol_dot, fol_bicg_scalars advances the recurrence between the vector kernels (stage 0..4:
 * top of the iteration, alpha, half-step test, omega, end), and fol_vec_op_dev computes
 *   out = sa*c_a*x + sb*c_b*y,  c_a = scalars[ia] (1 when ia < 0), likewise c_b; y may be NULL,
 * only while scalars[state] is one of the states in state_mask (bit s = state s; 0 running, 1 converged on the half
 * step, 2 done, 3 broken down) -- so a stopped iteration is frozen whatever the host still enqueues. */
int fol_bicg_scalar_count(void);
int fol_bicg_scalars(fol_stream_t s, int dtype, int stage, void* scalars);
int fol_vec_op_dev(fol_stream_t s, int dtype, int64_t n, const void* scalars, int state_mask, int ia, double sa,
                   const void* x, int ib, double sb, const void* y, void* out);
/* out[0] = x . y on the device (fixed two-stage reduction tree: run-to-run identical); `work` needs
 * fol_dot_work_size() elements of the call's dtype */
int64_t fol_dot_work_size(void);
int fol_dot(fol_stream_t s, int dtype, int64_t n, const void* x, const void* y, void* work, void* out);
/* The whole BiCGSTAB solve (same recurrences, start, stopping rule and break-down codes as the loop above and as
 * jax.scipy.sparse.linalg.bicgstab behind fol/solvers/fe_solver.py:62-67) as ONE persistent launch: SELL products, fused
 * vector passes and grid barriers inside, recurrence scalars in shared memory, no host read until the end.  For the
 * sizes where the multi-launch loop is latency-bound (configs[0], configs[3]).  dofs_per_node 2 / 3: `cols` are the node
 * columns of fol_sell_spmv_block; 0: scalar columns of fol_sell_spmv.  x: in x0, out solution.  work:
 * fol_bicgstab_fused_work_size(n) values of the call's dtype; after the stream has drained, work[8n + 16384 + 0..2] =
 * (iterations or -10 / -11, final |r|^2, barrier waits that gave up -- 0 in a healthy run). */
int64_t fol_bicgstab_fused_work_size(int64_t n);
int fol_bicgstab_fused(fol_stream_t s, int dtype, int dofs_per_node, int64_t n, const int64_t* slice_ptr,
                       const int32_t* cols, const void* values, const void* b, void* x, const void* m_diagonal,
                       double tol, double atol, int64_t maxiter, void* work);

/* Batched thermal loss + VJP (ThermalLoss2DQuad.ComputeBatchLoss and its gradient: thermal.py:28-49,
 * fe_loss.py:250-262) on a STRUCTURED Quad4 grid, 2 x 2 rule: nodes numbered row-major (node(c, r) = r (nx + 1) + c),
 * elements [n, n + 1, n + nx + 2, n + nx + 1] as fol/tools/usefull_functions.py:213-258 builds them, every element the
 * same parallelogram.  The caller (folax_b200/energy_plan.py::grid_structure) establishes those facts; other meshes use
 * fol_energy_and_grads.  jinv_host: row-major d xi_j / d x_k of the element shape; w_detj: Gauss weight x det J;
 * params_host[5], [6]: beta, c.  ctrl, u, grad_u, grad_k (may be null): (nb, (nx+1)(ny+1)); dir_values / dir_flag /
 * out_scale as in fol_energy_and_grads; col_dir: (nx + 1) bytes, 1 where the node column holds a Dirichlet node, or
 * null (every column may) -- the Dirichlet work then runs only in the warps that need it; ctrl, u and dir_values must be
 * 16-byte aligned; energy: (nb); work: fol_energy_grid_work_size values of the call's dtype.
 * Same results as the tile kernels to rounding (another summation order), deterministic. */
int64_t fol_energy_grid_work_size(int64_t nx, int64_t ny, int64_t nb);
int fol_energy_and_grads_grid(fol_stream_t s, int dtype, int64_t nx, int64_t ny, int64_t nb, const double* jinv_host,
                              double w_detj, const void* ctrl, const void* u, const void* dir_values,
                              const uint8_t* dir_flag, const uint8_t* col_dir, double out_scale,
                              const double* params_host, void* grad_u, void* grad_k, void* energy, void* work);

/* The same for the elasticity loss: MechanicalLoss2DQuad.ComputeBatchLoss and its gradient (mechanical.py:98-117,
 * fe_loss.py:250-262; plane stress, 2 x 2 rule) on the structured Quad4 grids of fol_energy_and_grads_grid.  u, grad_u,
 * dir_values, dir_flag: two dofs per node, interleaved ((nb,) 2 (nx+1)(ny+1)); ctrl: (nb, (nx+1)(ny+1)); the control
 * gradient of this loss is zero (its element matrix sits under stop_gradient) and is not written.  params_host[0..3]:
 * E, nu, body force x, y.  ctrl, u, dir_values and grad_u must be 16-byte aligned. */
int64_t fol_energy_grid_mech_work_size(int64_t nx, int64_t ny, int64_t nb);
int fol_energy_and_grads_grid_mech(fol_stream_t s, int dtype, int64_t nx, int64_t ny, int64_t nb, const double* jinv_host,
                                   double w_detj, const void* ctrl, const void* u, const void* dir_values,
                                   const uint8_t* dir_flag, const uint8_t* col_dir, double out_scale,
                                   const double* params_host, void* grad_u, void* energy, void* work);

/* ---- host-buffer entry point (what a non-GPU caller binds; used for the e2e measurement) -- */

typedef struct fol_plan fol_plan;
/* Builds the device-resident mesh plan (coords, connectivity, adjacency, flags) from HOST
 * arrays; owns its device memory until fol_plan_destroy. */
int fol_plan_create(fol_plan** plan, int dtype, int physics, int element, int num_gp,
                    int64_t ne, int64_t nn, const void* xyz_host, const int32_t* conn_host,
                    const int32_t* dirichlet_indices_host, int64_t n_dirichlet,
                    const double* params_host);
void fol_plan_destroy(fol_plan* plan);
/* HOST in (ctrl, u) -> HOST out (ke_data (ne*nd*nd), residual (ndof)); copies and kernels are
 * issued on the plan's stream, returns after the outputs are on the host. Host buffers should
 * be pinned for full PCIe rate. */
int fol_plan_assemble_host(fol_plan* plan, int transpose, const void* ctrl_host,
                           const void* u_host, void* ke_data_host, void* residual_host);
/* Duplicate-free CSR hand-off -- what the reference's solvers consume after summing the BCOO on the host
 * (fol/solvers/fe_solver.py:71-72).  fol_plan_set_csr uploads the integer plan once per mesh (host arrays of
 * folax_b200/csr_plan.py); fol_plan_assemble_host_csr then does H2D of the inputs, the element stage, the
 * de-duplication on the device and the D2H of nnz values (in the order of the plan's indptr / indices) + residual,
 * pipelined in chunks of whole node rows.  Returns after the copies have landed. */
int fol_plan_set_csr(fol_plan* plan, int64_t npairs, int64_t nnz, const int32_t* pair_ptr_host,
                     const int32_t* contrib_host, const int32_t* out_base_host, const int32_t* row_stride_host);
int fol_plan_assemble_host_csr(fol_plan* plan, int transpose, const void* controls_host, const void* dofs_host,
                               void* csr_values_host, void* residual_host);
/* same plan, device-resident outputs owned by the plan (for timing without the copies) */
int fol_plan_assemble_device(fol_plan* plan, int transpose, const void* ctrl_dev,
                             const void* u_dev, void** ke_data_dev, void** residual_dev);
fol_stream_t fol_plan_stream(fol_plan* plan);

/* ---- host-side integer plans of the solver hand-off (no device work; all host threads) ------------------------ */
/* Structure of the duplicate-free CSR the reference's solvers get from scipy.sparse.csr_array + sum_duplicates
 * (fol/solvers/fe_solver.py:71-72) and the fixed-order value plan of fol_csr_values; folax_b200/csr_plan.py drives the
 * two calls and holds the NumPy restatement the tests compare them with.  All pointers are HOST pointers.
 *   count: node -> (element, local node) adjacency (adj_ptr (nn+1), adj (ne*nnode), ascending e*nnode+a) and the number
 *          of distinct neighbour nodes deg (nn);
 *   fill:  with node_ptr (nn+1, int64) = exclusive prefix sum of deg: pair_ptr (npairs+1), contrib (ne*nnode*nnode: BCOO
 *          block positions (e*nnode+a)*nnode+b grouped by node pair, ascending inside a pair), out_base / row_stride
 *          (npairs), indptr (d*nn+1), indices (d*d*npairs). */
int fol_csr_plan_count_host(const int32_t* conn_host, int64_t ne, int nnode, int64_t nn, int32_t* adj_ptr_host,
                            int32_t* adj_host, int32_t* deg_host);
int fol_csr_plan_fill_host(const int32_t* conn_host, int64_t ne, int nnode, int64_t nn, int dofs_per_node,
                           const int32_t* adj_ptr_host, const int32_t* adj_host, const int64_t* node_ptr_host,
                           int32_t* pair_ptr_host, int32_t* contrib_host, int32_t* out_base_host,
                           int32_t* row_stride_host, int32_t* indptr_host, int32_t* indices_host);
/* Sliced-ELLPACK copy of a CSR structure for fol_sell_spmv / fol_sell_spmv_block (folax_b200/sell_plan.py): fills
 * cols / src (slice_ptr[nslices], pre-set to 0 / -1 by the caller), diag_src (nrows, pre-set to -1) and, when
 * node_cols_host is not NULL, one node column per run of dofs_per_node entries; *blocked_host = 1 if every row is made
 * of such runs (else node_cols is meaningless). */
int fol_sell_plan_fill_host(const int64_t* indptr_host, const int32_t* indices_host, int64_t nrows, int dofs_per_node,
                            int slice_height, const int64_t* slice_ptr_host, int32_t* cols_host, int32_t* src_host,
                            int32_t* diag_src_host, int32_t* node_cols_host, int* blocked_host);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* FP64 (or FP32) FMA peak microbenchmark: returns achieved TFLOP/s through *tflops. */
int fol_measure_fma_peak(int dtype, double* tflops);
/* HBM bandwidth of a pure WRITE stream of `bytes` (16-byte streaming stores, best of 5) in GB/s: the
 * ceiling of a store-dominated kernel such as the Jacobian assembly. */
int fol_measure_write_bandwidth(int64_t bytes, double* gbs);

/* ---- halo-DOF exchange of a slab-partitioned mesh over NVLink peer memory (one process per GPU) ----
 * Each rank creates one object for its interface planes (plane_dofs = nodes per plane * dofs per node),
 * exports a 64-byte CUDA IPC handle, receives its neighbours' handles through whatever transport the
 * host has (torch.distributed here) and connects them (side 0 = rank-1, side 1 = rank+1).
 * fol_halo_gather_push = deterministic residual gather of one interface plane FUSED with the peer
 * stores into the neighbour's receive buffer and an arrival signal; fol_halo_add = wait (on the device)
 * for the neighbour's push of `step` and add it to the plane.  Every rank calls push(step) before
 * add(step), with the same step sequence 0, 1, 2, ... */
typedef struct fol_halo fol_halo;
int fol_halo_create(fol_halo** halo, int dtype, int64_t plane_dofs);
void fol_halo_destroy(fol_halo* halo);
int fol_halo_export(fol_halo* halo, void* handle64);
int fol_halo_connect(fol_halo* halo, int side, const void* handle64);
int fol_halo_gather_push(fol_stream_t s, fol_halo* halo, int side, int64_t step, int64_t n0,
                         int64_t count, int dofs_per_node, const int32_t* adj_ptr,
                         const int32_t* adj, const void* re_elem, void* residual);
int fol_halo_add(fol_stream_t s, fol_halo* halo, int side, int64_t step, int64_t n0, int64_t count,
                 int dofs_per_node, void* residual);
/* Fused path (two launches per step instead of seven): fol_assemble_elements_halo runs the element stage of the slab
 * with its two interface element layers FIRST and hands the plane work -- the fixed-order residual gather of both
 * interface node planes, the NVLink peer stores and the arrival counts -- to the warps of the SAME launch as they
 * finish their tiles; fol_residual_gather_halo then gathers the interior nodes and adds what the neighbours pushed
 * (device-side wait).  Tuned Hex8 float64 kernels only (FOL_MECHANICAL / FOL_J2PLASTICITY, num_gp = 2), otherwise
 * FOL_ERR_UNSUPPORTED and the caller uses push / add above.  Same results, bit for bit, as the layered path.
 * Replaces the single-device scatter of fe_loss.py:301-306 for a slab (the reference has no domain decomposition). */
int fol_assemble_elements_halo(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne,
                               int64_t nn, const void* xyz, const int32_t* conn, const void* controls,
                               const void* dofs, const uint8_t* dir_flag, const double* params, void* ke_data,
                               void* re_elem, const void* state_in, void* state_out, fol_halo* halo, int64_t step,
                               int64_t layer_elems, int64_t plane_nodes, const int32_t* adj_ptr,
                               const int32_t* adj, void* residual);
int fol_residual_gather_halo(fol_stream_t s, fol_halo* halo, int64_t step, int64_t nn, int64_t plane_nodes,
                             const int32_t* adj_ptr, const int32_t* adj, const void* re_elem, void* residual);
/* arrival waits that gave up after ~2 s (a neighbour never pushed); synchronises; 0 in a healthy run */
int64_t fol_halo_timeouts(fol_halo* halo);

#ifdef __cplusplus
}
#endif
#endif /* FOLAX_B200_H */
