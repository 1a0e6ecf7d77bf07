/*
 * amaru_b200.h — C ABI of libamaru_b200.so: the B200 (sm_100a) implementation of Amaru.jl's
 * per-Newton-iteration mechanical hot path (mount_K -> solve_system! -> update_state!).
 *
 * The reference is pure Julia and has no FFI for this path today; each entry point below replaces one
 * Julia call of `mech_stage_solver!` (reference src/mech/mech-solver.jl:186-492) and is what a `ccall`
 * in the Julia glue binds.  Conventions:
 *   - every pointer is a HOST pointer, valid only for the duration of the call; the library copies;
 *   - indices are 0-based at the ABI (the glue subtracts 1 from Julia's eq_id / node ids);
 *   - every function returns an int status: 0 = ok (ReturnStatus success, src/tools/returnstatus.jl:9),
 *     >0 = expected failure (ReturnStatus failure; message in `msg`), <0 = usage / CUDA error;
 *   - `msg`/`msglen`: caller-supplied buffer that receives a NUL-terminated message (may be NULL/0);
 *   - one opaque handle per analysis stage; a handle is not re-entrant;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     AMARU_ERR_NO_DEVICE.
 *
 * Vector layout at the ABI is the reference's: length ndofs, indexed by eq_id, unknown dofs first
 * (src/bc.jl:198-233).  IP state layout is element-major (in the order elements are passed, batch after
 * batch), then integration-point order of the quadrature table (src/mech/mech-solver.jl:245), with
 * Mandel components (xx, yy, zz, sqrt2*yz, sqrt2*xz, sqrt2*xy) (src/tools/tensors.jl:24-25).
 */
#ifndef AMARU_B200_H
#define AMARU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define AMARU_OK 0
#define AMARU_FAIL_MATERIAL 1       /* material update failed (von-mises.jl:146)                          */
#define AMARU_FAIL_NAN 2            /* NaN in internal forces (mech-solver.jl:142)                        */
#define AMARU_FAIL_SINGULAR 3       /* max|U1| > 1e8, "Possible syngular matrix" (solver.jl:68-71)        */
#define AMARU_FAIL_NEG_JACOBIAN 4   /* detJ <= 0 (mech-solid.jl:150)                                      */
#define AMARU_FAIL_CG_NOCONV 5      /* PCG hit maxit before reaching cg_rtol                              */
#define AMARU_FAIL_TANGENT 6        /* calcD assertion J2 > 0 failed (von-mises.jl:117)                   */
#define AMARU_ERR_ARG (-1)
#define AMARU_ERR_CUDA (-2)
#define AMARU_ERR_NO_DEVICE (-3)
#define AMARU_ERR_UNSUPPORTED (-4)  /* element / material / stress model outside the hot path             */
#define AMARU_ERR_COMM (-5)

/* cell shapes (reference src/shape/solids2d.jl, solids3d.jl); default quadrature of each is used */
#define AMARU_SHAPE_QUAD4 1  /* nn 4,  nd 2, QUAD_IP4 */
#define AMARU_SHAPE_QUAD8 2  /* nn 8,  nd 2, QUAD_IP4 */
#define AMARU_SHAPE_HEX8 3   /* nn 8,  nd 3, HEX_IP8  */
#define AMARU_SHAPE_HEX20 4  /* nn 20, nd 3, HEX_IP8  */
#define AMARU_SHAPE_TET10 5  /* nn 10, nd 3, TET_IP4  */
/* facet shapes, accepted by amaru_loadset_create only (src/shape/lines.jl, solids2d.jl:99-151) */
#define AMARU_SHAPE_LIN2 101 /* nn 2, LIN_IP2 : edge of QUAD4 */
#define AMARU_SHAPE_LIN3 102 /* nn 3, LIN_IP2 : edge of QUAD8 */
#define AMARU_SHAPE_TRI6 103 /* nn 6, TRI_IP3 : face of TET10 */

/* material kinds; params[8] = {E, nu, p2, p3, p4, rho, 0, 0} */
#define AMARU_MAT_LINEAR_ELASTIC 1 /* linear-elastic.jl:12-20 : E, nu                                    */
#define AMARU_MAT_VON_MISES 2      /* von-mises.jl:14-26      : E, nu, fy, H                             */
#define AMARU_MAT_DRUCKER_PRAGER 3 /* drucker-prager.jl:15-34 : E, nu, alpha, kappa, H                   */
#define AMARU_MAT_NPARAMS 8

/* stress models accepted (mech-solver.jl:9): everything that uses the 3D constitutive matrix */
#define AMARU_STRESS_D3 0
#define AMARU_STRESS_PLANESTRAIN 1
/* plane stress, LinearElastic only (the only material whose calcDe has that branch, linear-elastic.jl:99-108): 2D models;
 * σzz stays zero, the in-plane moduli are c = E/(1-ν²): c, cν, c(1-ν) */
#define AMARU_STRESS_PLANESTRESS 2
/* axisymmetric (2D cells in the r-z plane): hoop row ε_θθ = N_a/r of B and th = 2π·r at every integration point
 * (mech-solid.jl:94-108,143,180,260); the PCG of such handles runs on the assembled block-CSR matrix (the matrix-free
 * operator has no hoop term) and distributed loads are integrated by the host glue (distributed.jl:121,193) */
#define AMARU_STRESS_AXISYMMETRIC 3

/* preconditioners for the PCG that replaces lu(K11) (solver.jl:42-43) */
#define AMARU_PRECOND_JACOBI 0
#define AMARU_PRECOND_BLOCK_JACOBI 1 /* nd x nd node blocks */

/* operator the PCG applies each iteration in place of K11*p (solver.jl:42-43 factorises K11 instead):
 *   EBE (default): matrix-free, re-integrated per element from the per-IP tangent data that amaru_assemble_K refreshes;
 *   CSR          : SpMV on the assembled block-CSR values.  Both use the same assembled K for the preconditioner, for
 *   amaru_get_csr and for amaru_matvec.  AMARU_OPERATOR=csr|ebe in the environment sets the default of new handles. */
#define AMARU_OPERATOR_CSR 0
#define AMARU_OPERATOR_EBE 1

typedef struct amaru_model amaru_model; /* opaque */

/* Library / device probe.  Returns the number of visible CUDA devices (0 = none), <0 on error. */
int amaru_device_count(void);
const char *amaru_version(void);

/*
 * Flatten-and-upload: replaces the per-call object walks of the reference — getcoords
 * (src/element.jl:112-116), the dof-map build (src/mech/elem/mech-solid.jl:160-161), configure_dofs!
 * (src/bc.jl:198-233) and the `State` vector build (src/mech/mech-solver.jl:245).
 *
 *   ndim          2 or 3;  stressmodel AMARU_STRESS_*;  thickness = ctx.thickness (mech-solid.jl:126)
 *   coords        [nnodes*3] row-major x,y,z
 *   nbatches      element batches, each of one cell shape (elements sorted by type)
 *   batch_shape   [nbatches] AMARU_SHAPE_*
 *   batch_nelem   [nbatches]
 *   conn          concatenation over batches of [nelem_b * nn_b] node ids, element-major
 *   elem_mat      [nelem_total] index into the material table
 *   mat_kind      [nmats] AMARU_MAT_*;  mat_params [nmats*8]
 *   eqid          [nnodes*ndim] eq_id of (node, ux|uy|uz); unknown dofs are 0..nu-1
 *   ngpus         number of B200s of this box the handle spreads over.  1: everything on devices[0]
 *                 (device 0 when devices is NULL).  N > 1: the mesh is partitioned inside the library (element partition,
 *                 node ownership, duplicated halo elements), every GPU assembles the CSR rows of its own nodes and one
 *                 host thread per GPU drives it; the CG loop exchanges halo entries and scalars through NVLink peer
 *                 memory (cudaDeviceEnablePeerAccess), no NCCL, no extra processes.  Every other entry point below
 *                 takes the same handle and the same global-length host vectors, so `solve!(ana)` stays ONE process
 *                 (reference src/mech/mech-solver.jl:172-181).
 *   devices       [ngpus] CUDA device ordinals, or NULL for 0..ngpus-1
 *   partitioner   AMARU_PARTITION_RCB | AMARU_PARTITION_METIS (ignored for ngpus == 1)
 * Builds on the device: the element colouring, the symbolic block-CSR pattern, the scatter map.
 * IP state starts at zero (the IpState constructors, e.g. von-mises.jl:35-42).
 */
#define AMARU_PARTITION_RCB 0   /* recursive coordinate bisection of the element centroids (deterministic, compact boxes)   */
#define AMARU_PARTITION_METIS 1 /* METIS k-way on the element dual graph (facet adjacency); falls back to RCB on tiny meshes */
int amaru_create(int ndim, int stressmodel, double thickness,
                 int64_t nnodes, const double *coords,
                 int nbatches, const int32_t *batch_shape, const int64_t *batch_nelem,
                 const int32_t *conn, const int32_t *elem_mat,
                 int nmats, const int32_t *mat_kind, const double *mat_params,
                 const int32_t *eqid, int64_t ndofs, int64_t nu,
                 int ngpus, const int32_t *devices, int partitioner,
                 amaru_model **out, char *msg, int msglen);
/* number of GPUs behind the handle */
int amaru_ngpus(const amaru_model *m);
/* The element partition amaru_create(ngpus = nparts) would use: elem_part[nelem_total] in ABI element order.  Host only. */
int amaru_partition_elements_abi(int partitioner, int nparts, int64_t nnodes, const double *coords, int nbatches,
                                 const int32_t *batch_shape, const int64_t *batch_nelem, const int32_t *conn,
                                 int32_t *elem_part, char *msg, int msglen);
/* Failure injection for the bounded peer-memory waits of multi-GPU handles: every GPU but `skip_rank` enters one scalar
 * all-reduce; returns AMARU_ERR_COMM after the timeout (AMARU_P2P_TIMEOUT_MS, default 10 s) instead of hanging, and the
 * handle must be destroyed afterwards.  skip_rank < 0: nobody is skipped, returns AMARU_OK. */
int amaru_comm_selftest(amaru_model *m, int skip_rank, char *msg, int msglen);

/* Multi-GPU variant (one process per GPU; the reference has no distributed path).  This rank passes its
 * LOCAL view of the partitioned mesh: `nnodes` local nodes of which the first `nowned` are owned (the rest are ghosts,
 * grouped by owner rank), its own + halo elements (every element touching an owned node) with local node ids, and
 * `eqid` = GLOBAL eq ids of the local nodes (ndofs / nu are the global numbers; ABI vectors stay global-length).
 * Halo lists: for neighbour q = neigh_rank[i], send the owned local nodes send_nodes[send_ptr[i] .. send_ptr[i+1]) and
 * receive recv_count[i] ghosts into local nodes [recv_start[i], ...); both sides order the nodes by global id.
 * `nccl_uid` is the 128-byte ncclUniqueId made on rank 0 by amaru_nccl_unique_id and broadcast by the host. */
int amaru_create_partitioned(int ndim, int stressmodel, double thickness,
                             int64_t nnodes, int64_t nowned, const double *coords,
                             int nbatches, const int32_t *batch_shape, const int64_t *batch_nelem,
                             const int32_t *conn, const int32_t *elem_mat,
                             int nmats, const int32_t *mat_kind, const double *mat_params,
                             const int32_t *eqid, int64_t ndofs, int64_t nu,
                             int rank, int nranks, int nneigh, const int32_t *neigh_rank, const int64_t *send_ptr,
                             const int32_t *send_nodes, const int64_t *recv_start, const int64_t *recv_count,
                             const void *nccl_uid, int device, amaru_model **out, char *msg, int msglen);
int amaru_nccl_unique_id(void *uid128, char *msg, int msglen);
/* Optional peer-memory path for the per-iteration exchanges of the CG loop (one box, NVLink): every rank exports 128 bytes
 * (two cudaIpc handles: its flag/scalar window and its p vector), the host all-gathers them in rank order and calls
 * amaru_p2p_connect with, for every neighbour i of this rank, the first local node id (in THAT neighbour's numbering) of
 * the ghost range it keeps for this rank (= the neighbour's recv_start entry for this rank).  Afterwards the halo exchange
 * of p and the scalar all-reduces of the CG loop run as two small kernels over peer memory instead of three NCCL calls;
 * everything else (result vectors, status flags) stays on NCCL. */
int amaru_p2p_export(amaru_model *m, void *out128, char *msg, int msglen);
int amaru_p2p_connect(amaru_model *m, const void *all_handles, const int64_t *peer_recv_start, char *msg, int msglen);
/* collective: call with on=1 on every rank once every rank's amaru_p2p_connect returned 0 (else keep NCCL everywhere) */
int amaru_p2p_enable(amaru_model *m, int on);

int amaru_destroy(amaru_model *m);

/* sizes */
int64_t amaru_nip_total(const amaru_model *m);
int64_t amaru_nnz(const amaru_model *m);       /* scalar non-zeros of the symbolic pattern             */
int64_t amaru_nblocks(const amaru_model *m);   /* nd x nd blocks stored                                */
int amaru_ncolors(const amaru_model *m);

/* ip.state in / out (src/mech/mat/linear-elastic.jl:23-34, von-mises.jl:29-43, drucker-prager.jl:46-60).
 * sigma, eps: [nip_total*6];  epa, dlam: [nip_total] (dlam = Δλ | Δγ; zeros for linear-elastic IPs).
 * Any pointer may be NULL to skip that field. */
int amaru_set_state(amaru_model *m, const double *sigma, const double *eps, const double *epa,
                    const double *dlam, char *msg, int msglen);
int amaru_get_state(amaru_model *m, double *sigma, double *eps, double *epa, double *dlam, char *msg,
                    int msglen);

/* copyto!.(StateBk, State) / copyto!.(State, StateBk)  (src/mech/mech-solver.jl:391,333; src/ip.jl:38-52) */
int amaru_state_backup(amaru_model *m);
int amaru_state_restore(amaru_model *m);

/* mount_K (src/mech/mech-solver.jl:78-110) with elem_stiffness (src/mech/elem/mech-solid.jl:124-166) and
 * calcD of the three materials; K stays on the device. */
int amaru_assemble_K(amaru_model *m, char *msg, int msglen);

/* Predictor-corrector schemes :ME / :BE / :Ralston (src/mech/mech-solver.jl:279-288,341-350): after the predictor's
 * mount_K -> amaru_tangent_save keeps that K; after the corrector's `K2 = mount_K(...)` (amaru_assemble_K on the updated
 * state) amaru_tangent_blend forms `K = a1*K + a2*K2` in place, which the next amaru_solve uses.  A genuine blend (a1 != 0)
 * is no longer described by the per-IP tangent data of the matrix-free operator, so that solve runs the block-CSR SpMV;
 * the next amaru_assemble_K restores the default. */
int amaru_tangent_save(amaru_model *m, char *msg, int msglen);
int amaru_tangent_blend(amaru_model *m, double a1, double a2, char *msg, int msglen);

/* Symbolic CSR pattern in eq_id numbering (== the CSC of the reference's symbolic K, which is structurally
 * symmetric) and the assembled values.  rowptr [ndofs+1], colind/val [nnz], columns ascending per row.
 * val may be NULL (pattern only).  For parity tests and for callers that want K back. */
int amaru_get_csr(amaru_model *m, int64_t *rowptr, int32_t *colind, double *val, char *msg, int msglen);

/* solve_system!(K, U, F, nu) (src/solver.jl:5-79): on entry U[nu:] holds prescribed values and F[:nu] the
 * known forces; on return U[:nu] is the solution and F[nu:] the reactions (solver.jl:74-75).  lu(K11) is
 * replaced by a preconditioned CG on the device, stopped at ||r|| <= cg_rtol*||b||.
 * iters / relres (may be NULL) receive the CG iteration count and final relative residual. */
int amaru_solve(amaru_model *m, double *U, double *F, double cg_rtol, int cg_maxit, int precond,
                int *iters, double *relres, char *msg, int msglen);

/* update_state!(active_elems, ΔUt, t) (src/mech/mech-solver.jl:124-144) with update_elem!
 * (src/mech/elem/mech-solid.jl:243-279) and the material update_state! functions; mutates the device IP
 * state and returns ΔFin[ndofs]. */
int amaru_update_state(amaru_model *m, const double *dU, double *dFin, char *msg, int msglen);

/* elem_internal_forces (src/mech/elem/mech-solid.jl:208-240) summed over all elements: Fin[ndofs]. */
int amaru_internal_forces(amaru_model *m, double *Fin, char *msg, int msglen);

/* ---- next tier: Newmark dynamics (src/mech/dyn-solver.jl:72-103,373-399) -------------------- */
/* mount_M with elem_mass (mech-solid.jl:169-205); rho[nelem_total]. M uses K's pattern. */
int amaru_assemble_M(amaru_model *m, const double *rho, char *msg, int msglen);
/* The matrix used by amaru_solve / amaru_matvec becomes  a*K + b*M  (Kp = K + 4/Δt² M + 2/Δt C with
 * C = αM + βK folds into two scalars, dyn-solver.jl:376-377). a=1,b=0 restores plain K. */
int amaru_set_system_matrix(amaru_model *m, double a, double b, char *msg, int msglen);
/* y = (a*K + b*M) x over all ndofs (eq_id ordering), e.g. M*A, C*V products of dyn-solver.jl:378,399 */
int amaru_matvec(amaru_model *m, double a, double b, const double *x, double *y, char *msg, int msglen);

/* ---- next tier: natural boundary conditions integrated on the device ---------------------------
 * SurfaceBC / BodyC of the reference (src/bc.jl:116-136,175-194) call, per facet / element and per key,
 * mech_boundary_forces / mech_solid_body_forces (src/mech/elem/distributed.jl:76-152,157-217) and add the result
 * into F through the facet's dof map.  A load set is the list of entities one boundary condition selected:
 *   shape   facet shape (LIN2/LIN3 edges of 2D cells; QUAD4/QUAD8/TRI6 faces of 3D cells) => traction keys tx ty tz tn,
 *           or the cell shape itself (shape dimension == ndim)                             => body-force keys wx wy wz
 *   nodes   [nents * nn(shape)] node ids, entity-major, in the facet's / element's local node order
 * Integration uses the DEFAULT quadrature of the shape (get_ip_coords(shape), src/shape/shape.jl:61-64).
 * The value expression of the condition (a Julia Expr of t,x,y,z) stays on the host: amaru_loadset_ip_coords returns
 * the integration-point coordinates X = C'N once, the glue evaluates the expression there and passes one value per
 * integration point (`vip`, entity-major then quadrature order), or NULL + `cval` for a constant. */
typedef struct amaru_loadset amaru_loadset; /* opaque */
#define AMARU_LOAD_X 0      /* tx | wx */
#define AMARU_LOAD_Y 1      /* ty | wy */
#define AMARU_LOAD_Z 2      /* tz | wz */
#define AMARU_LOAD_NORMAL 3 /* tn : vip * normalize(n), n = [J2,-J1] (2D) or J[:,1] x J[:,2] (3D), distributed.jl:134-141 */
int amaru_loadset_create(amaru_model *m, int shape, int64_t nents, const int32_t *nodes, amaru_loadset **out,
                         char *msg, int msglen);
int64_t amaru_loadset_nip(const amaru_loadset *ls); /* nents * nip(shape) */
int amaru_loadset_ip_coords(amaru_loadset *ls, double *X /* [nip*3] */, char *msg, int msglen);
/* F[ndofs] += sum over entities of the nodal forces (eq_id ordering; same accumulation order as the reference:
 * entity after entity). */
int amaru_loadset_apply(amaru_loadset *ls, int key, double cval, const double *vip, double *F, char *msg, int msglen);
int amaru_loadset_destroy(amaru_loadset *ls);

/* ---- next tier: output side -------------------------------------------------------------------
 * nodal_patch_recovery (src/fe-model.jl:506-692): the integration-point fields of ip_state_vals
 * (src/tools/tensors.jl:162-216; von-mises.jl:159-168; drucker-prager.jl:152-163) are fitted per corner-node patch with
 * the regression polynomial reg_terms (fe-model.jl:490-503) and averaged at the nodes, on the device, from the IP state
 * the handle holds.  `at_bound[nnodes]` flags the nodes of model.faces (the outer facets, fe-model.jl:523-529).
 * Field order is the reference's column order: σxx σyy σzz σyz σxz σxy σvm σ1 [σ2] σ3 εxx εyy εzz [εyz εxz] εxy
 * (bracketed ones only for the d3 stress model), then `ep` (VonMises) and/or `epa j1 srj2d` (DruckerPrager) in order
 * of first appearance in the element list.  V is field-major: V[i*nnodes + node] = V_rec[node, i]. */
int amaru_recovery_create(amaru_model *m, const uint8_t *at_bound, char *msg, int msglen);
int amaru_recovery_nfields(const amaru_model *m);
int amaru_recovery_field(const amaru_model *m, int i, int *code, char *name, int namelen);
int amaru_recover_nodal(amaru_model *m, double *V, char *msg, int msglen);

/* save_vtu (src/mesh/io.jl:167-276, uncompressed branch; XML layout of src/tools/xml.jl:253-316): writes an ASCII
 * .vtu whose DataArray contents are formatted like the reference's get_array_node! (io.jl:150-163): floats as
 * "%20.10e" of Float32(value), integers followed by two blanks, one row per line.  Host-side IO, no device work; the
 * arrays usually come from amaru_recover_nodal and the solution vectors.
 *   cell data/point data: `n*_arrays` named arrays; type 0 = Float64, 1 = Int64, 2 = Int32, 3 = UInt64; `ncomp` components per row. */
int amaru_write_vtu(const char *filename, const char *desc, int64_t nnodes, const double *coords, int nbatches,
                    const int32_t *batch_shape, const int64_t *batch_nelem, const int32_t *conn,
                    int npoint_arrays, const char *const *point_names, const int32_t *point_type,
                    const int32_t *point_ncomp, const void *const *point_data,
                    int ncell_arrays, const char *const *cell_names, const int32_t *cell_type,
                    const int32_t *cell_ncomp, const void *const *cell_data, char *msg, int msglen);

/* ---- next tier: host-side model generation straight into the arrays amaru_create takes ---------
 * Structured mesher for two-corner box Blocks with uniform spacing (src/mesh/structured.jl:182-231,384-552;
 * src/mesh/block.jl:3-30): same node creation order (k outer, j, i inner; serendipity points skipped), coordinates
 * rounded to 8 digits (src/node.jl:57-61), same cell order and local node order (six TET10 per cell,
 * structured.jl:537-542).  `box` = {x0,y0,z0, x1,y1,z1}; nz ignored for 2D shapes.  Host-only, multi-threaded. */
int amaru_mesh_block_sizes(int shape, int nx, int ny, int nz, int64_t *nnodes, int64_t *nelems, int *nn);
int amaru_mesh_block(int shape, const double *box, int nx, int ny, int nz, double *coords /* [nnodes*3] */,
                     int32_t *conn /* [nelems*nn] */, char *msg, int msglen);
/* get_outer_facets (src/mesh/mesh.jl:69-85): facets seen by exactly one cell, in cell order then local facet order, nodes in
 * the owner's facet_idxs order.  Call with facet_nodes = owner = NULL to get the count; returns the count (< 0 on error). */
int64_t amaru_outer_facets(int shape, int64_t nelem, const int32_t *conn, int32_t *facet_nodes, int64_t *owner,
                           int64_t capacity, int *nodes_per_facet);
/* configure_dofs! (src/bc.jl:198-233): `prescribed[nnodes*nd]` flags per (node, ux|uy|uz) -> eq ids with the unknown dofs
 * first (stable), and nu. */
int amaru_configure_dofs(int64_t nnodes, int nd, const uint8_t *prescribed, int32_t *eqid, int64_t *nu);

/* ---- measurement hooks (bench.py): device-resident Newton iteration, no host copies ----------- */
/* One assemble_K + solve + state_restore + update_state with U/F/dFin kept on the device; returns the
 * CUDA-event time of each phase in ms (4 doubles: assemble, solve, update, total) and CG iterations. */
int amaru_newton_iteration_device(amaru_model *m, double cg_rtol, int cg_maxit, int precond,
                                  double *phase_ms, int *iters, double *relres, char *msg, int msglen);
/* Upload the vectors used by amaru_newton_iteration_device (U: prescribed values, F: known forces). */
int amaru_set_device_vectors(amaru_model *m, const double *U, const double *F, char *msg, int msglen);
/* time `reps` launches of one kernel class on the model's stream (CUDA events); kind: 0 SpMV, 1 assemble,
 * 2 update_state, 3 fused CG vector update, 4 p-update. Returns average ms per launch. */
int amaru_time_kernel(amaru_model *m, int kind, int precond, int reps, double *avg_ms, char *msg, int msglen);
/* AMARU_OPERATOR_CSR | AMARU_OPERATOR_EBE for the following amaru_solve calls of this handle */
int amaru_set_operator(amaru_model *m, int kind);
/* y = A x with the operator the PCG of this handle applies (EBE or CSR) and the current system matrix a*K + b*M, in
 * eq_id ordering over all ndofs.  masked != 0: rows of prescribed dofs are zeroed, as inside the CG loop, and *pAp
 * receives the fused x.Ax (equal to the masked dot when x vanishes on the prescribed dofs).  For parity tests. */
int amaru_operator_apply(amaru_model *m, const double *x, double *y, int masked, double *pAp, char *msg, int msglen);
/* When on, amaru_solve brackets every CG SpMV launch with CUDA events on the model's stream; amaru_get_profile
 * returns the summed duration (ms) and the number of SpMV launches since profiling was switched on. */
int amaru_set_profiling(amaru_model *m, int on);
int amaru_get_profile(amaru_model *m, double *spmv_ms_total, int64_t *spmv_launches);
/* Algorithmic bytes one SpMV launch must move: matrix values + column/row metadata of the storage
 * format actually used + x read once + y written once. */
int64_t amaru_spmv_bytes(const amaru_model *m);
/* name of the SpMV kernel the CG loop of this handle launches (for the bench's roofline record) */
const char *amaru_spmv_kernel(const amaru_model *m);
/* Host-only diagnostic of the matrix-free operator's patch plan (patches.cpp): builds the plan amaru_create would build
 * for one batch of `shape` cells and checks its invariants (every element in exactly one slot, groups node-disjoint,
 * patch-local ids consistent, first-touch / ghost flags, dependencies = every earlier patch sharing a node).  Returns 0 or
 * the number of the violated rule; stats[6] = patches, patch colours, slot fill, slots, max groups per patch, dependencies. */
int amaru_patch_plan_check(int shape, int64_t nnodes, int64_t nowned, const double *coords, int64_t nelem,
                           const int32_t *conn, double *stats);
/* number of kernels launched by this handle since creation (the bench's gpu_launches claim) */
int64_t amaru_launch_count(const amaru_model *m);

#ifdef __cplusplus
}
#endif
#endif /* AMARU_B200_H */
