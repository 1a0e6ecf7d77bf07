// Internal declarations shared by the translation units of libamaru_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/amaru_b200.h"

#define AMARU_MAXNN 20
#define AMARU_NSTATE 14  // planes: 0-5 sigma, 6-11 eps, 12 epa, 13 dlam (Δλ | Δγ)

// ---- error plumbing -------------------------------------------------------------------------------------
struct AmaruError {
    int code;
    std::string msg;
};
#define CUDA_CHECK(call)                                                                          \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            throw AmaruError{AMARU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)}; \
    } while (0)
#define AMARU_REQUIRE(cond, code, text) \
    do {                                \
        if (!(cond)) throw AmaruError{(code), (text)}; \
    } while (0)

// ---- shape tables (host) ---------------------------------------------------------------------------------
struct ShapeInfo {
    int id, nn, nd, nip;
    std::vector<double> nat;    // nn*nd natural coordinates
    std::vector<double> ips;    // nip*4 (r,s,t,w)
    std::vector<double> N;      // nip*nn
    std::vector<double> dNdR;   // nip*nn*nd
};
bool amaru_shape_info(int shape_id, ShapeInfo &out);

// ---- host preprocessing (host_prep.cpp) -------------------------------------------------------------------
struct HostPattern {
    std::vector<int32_t> rowptr;   // nrows+1   (block rows = nodes 0..nrows-1)
    std::vector<int32_t> col;      // nblk, ascending inside a row
    std::vector<int32_t> diag;     // nrows: index of the diagonal block
};
// node->element adjacency over all batches (element ids are global across batches)
void amaru_build_adjacency(int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem,
                           const int32_t *const *conn, std::vector<int64_t> &adj_ptr, std::vector<int64_t> &adj);
// greedy colouring: elements sharing a node get different colours; returns number of colours
int amaru_color_elements(int64_t nnodes, int nbatches, const int *nn, const int64_t *nelem,
                         const int32_t *const *conn, const std::vector<int64_t> &adj_ptr,
                         const std::vector<int64_t> &adj, std::vector<int32_t> &color);
// symbolic block pattern for rows [0,nrows): union of the nodes of all elements touching the row node
void amaru_build_pattern(int64_t nrows, int nbatches, const int *nn, const int64_t *nelem,
                         const int32_t *const *conn, const std::vector<int64_t> &adj_ptr,
                         const std::vector<int64_t> &adj, HostPattern &pat);
int amaru_host_threads();

// ---- patch plan of the matrix-free operator (patches.cpp; consumed by ebe.cu) ---------------------------------
constexpr uint32_t PN_NODE = 0x07ffffffu;    // patch-node entry: node | first-touch << 27 | prescribed-dof mask << 28 | ghost << 31
constexpr uint32_t PN_FIRST = 1u << 27;      // no earlier patch (in processing order) lists the node: y starts from zero
constexpr uint32_t PN_GHOST = 1u << 31;      // row of a neighbour rank: never stored
struct PatchPlan {
    int pe = 64, maxpn = 0, nn = 0;
    int64_t nslots = 0;                 // 8 element slots per group, groups numbered patch by patch
    int npatch = 0, ncolors = 0, maxgroups = 0;
    double fill = 0.0;                  // elements / slots
    std::vector<int32_t> slot_elem;     // [nslots] colour-sorted element of the slot, -1 = empty
    std::vector<int32_t> desc;          // [npatch][8]: first group, groups, first node entry, nodes, first dep, deps, 0, 0
    std::vector<uint32_t> pnodes;       // node entries, ascending node id inside a patch
    std::vector<uint16_t> lane_ids;     // [group][word][lane][4]: patch-local node ids in DMMA fragment order (ebe_patch.cuh)
    std::vector<int32_t> deps;          // lower-coloured patches sharing a node with the patch
};
inline int amaru_patch_id_words(int nn) { return ((nn + 3) / 4 + 2 * ((nn + 7) / 8) + 3) / 4; }
// elements per patch, node capacity of the shared-memory bricks, brick size in cells
void amaru_patch_shape_params(int shape, int &pe, int &maxpn, int brick[3]);
void amaru_build_patches(int nn, int nd, int pe, int maxpn, const int brick[3], int64_t nelem, const int32_t *sconn,
                         const std::vector<int64_t> &color_off, int64_t nnodes, int64_t nowned, const double *coords,
                         const uint8_t *fixed, std::vector<uint8_t> &touched, PatchPlan &out);

// ---- device model ------------------------------------------------------------------------------------------
struct Batch {
    int shape = 0, nn = 0, nd = 0, nip = 0;
    int64_t nelem = 0;
    int64_t elem_off = 0;          // first element of this batch in the ABI element order
    int64_t ip_off = 0;            // first IP of this batch in the device state planes
    std::vector<int64_t> color_off;  // ncolors+1 offsets into the colour-sorted arrays
    int32_t *d_conn = nullptr;     // [nelem*nn]   colour-sorted, element-major
    int32_t *d_emat = nullptr;     // [nelem]      material index, colour-sorted
    int32_t *d_map = nullptr;      // [nelem*nn*nn] destination block of pair (a,b)
    int64_t *d_perm = nullptr;     // [nelem]      colour-sorted position -> element index inside the batch
    uint8_t *d_owned = nullptr;    // [nelem]      1 if this rank owns the element (multi-GPU), else nullptr
    double *d_rho = nullptr;       // [nelem]      density, colour-sorted (mass assembly)
    double *d_dNdR = nullptr;      // [nip*nn*nd]
    double *d_N = nullptr;         // [nip*nn]
    double *d_w = nullptr;         // [nip]
};

struct CgScalars;  // device-side scalar block (pcg.cu)

struct amaru_model {
    int device = 0;
    int ndim = 0, nd = 0, stressmodel = 0;
    double th = 1.0;
    int64_t nnodes = 0;        // local nodes (owned + ghost)
    int64_t nowned = 0;        // rows of the local matrix
    int64_t ndofs = 0, nu = 0; // ABI vector length / unknown dofs (global numbers)
    int64_t nelem_total = 0, nip_total = 0;
    int ncolors = 0;
    int nmats = 0;
    cudaStream_t stream = nullptr;
    int nsm = 148;

    std::vector<Batch> batches;
    double *d_coords = nullptr;    // [nnodes*3]
    int32_t *d_eqid = nullptr;     // [nnodes*nd] local dof (node-major) -> ABI eq index
    uint8_t *d_fixed = nullptr;    // [nnodes*nd] 1 = prescribed
    int32_t *d_mat_kind = nullptr; // [nmats]
    double *d_mat_par = nullptr;   // [nmats*8]

    // block-CSR matrix, rows = owned nodes, columns = local nodes
    int64_t nblk = 0;
    int32_t *d_rowptr = nullptr, *d_col = nullptr, *d_diag = nullptr;
    double *d_K = nullptr;         // [nblk*nd*nd]
    double *d_M = nullptr;         // mass, same pattern (next tier)
    double *d_A = nullptr;         // a*K+b*M when a system matrix is set, else == d_K
    double *d_Abuf = nullptr;      // storage behind d_A when it is not d_K
    double sysA = 1.0, sysB = 0.0;
    double *d_Minv = nullptr;      // block-Jacobi inverses [nowned*nd*nd] or Jacobi [nowned*nd]
    int minv_kind = -1;            // which preconditioner d_Minv currently holds (-1 = stale)

    // IP state planes [AMARU_NSTATE][nip_total] and the converged backup
    double *d_state = nullptr, *d_statebk = nullptr;

    // work vectors, node-major, length nnodes*nd
    double *d_x = nullptr, *d_r = nullptr, *d_z = nullptr, *d_p = nullptr, *d_q = nullptr, *d_b = nullptr;
    double *d_f = nullptr;         // internal forces accumulator
    double *d_io = nullptr;        // ABI-order staging buffer, length max(ndofs, 6*nip_total)
    double *d_U = nullptr, *d_F = nullptr;  // device-resident ABI-order vectors
    double *d_U0 = nullptr, *d_F0 = nullptr;  // inputs of amaru_newton_iteration_device (measurement hook)
    int64_t io_len = 0;

    // reductions / flags
    double *d_partial = nullptr;   // per-block partial sums
    CgScalars *d_scal = nullptr;
    int *d_status = nullptr;       // element-kernel failure flag
    int *h_pinned = nullptr;       // pinned host mirror for small reads
    int grid_rows = 0;             // persistent grid used by the row kernels
    // streamed SpMV (spmv.cu): row tiles, per-tile records, launch geometry
    void *d_tiles = nullptr;       // SpmvTile[ntiles]
    int32_t *d_tmeta = nullptr;    // per-tile records: packed row entries + unique columns + 16-bit local columns
    int ntiles = 0, tile_blks = 0, tile_rows = 0, tile_xcap = 0, grid_tma = 0, spmv_stages = 0, spmv_warps = 0, spmv_xd = 2, spmv_sleep = 100, spmv_ver = 2;
    int64_t spmv_meta_bytes = 0;   // bytes of tile records + headers streamed per SpMV
    bool use_tma = false;
    // symmetric-storage SpMV of the CG loop (spmv.cu): upper-triangular blocks (col >= row; every owned x ghost block)
    bool use_sym = false;
    int64_t nublk = 0;             // stored upper blocks
    int32_t *d_usrc = nullptr;     // [nublk] index of the block in the full pattern
    double *d_Asym = nullptr;      // [nublk*nd*nd] values, refreshed whenever d_A changes
    bool sym_fresh = false;        // d_Asym mirrors the current d_A
    void *d_stiles = nullptr;      // SpmvTile[nstiles]
    int32_t *d_stmeta = nullptr;
    int nstiles = 0, stile_xcap = 0, sgrid = 0, sym_ystages = 2;
    int64_t sym_meta_bytes = 0;

    // multi-GPU
    int rank = 0, nranks = 1;
    void *comm = nullptr;          // HaloComm* (halo.cu)
    struct AmaruGroup *grp = nullptr;     // != nullptr: this handle is a group of per-GPU parts in one process (group.cu)
    struct AmaruGroup *grp_of = nullptr;  // != nullptr: this handle is a part of that group
    bool io_shared = false;        // part of a group: d_U / d_F / d_U0 / d_F0 alias the first part's vectors (peer access)

    // output side (recovery.cu)
    void *recovery = nullptr;      // Recovery*
    bool recovery_vm_first = true;

    // matrix-free tangent operator of the CG loop (ebe.cu)
    void *ebe = nullptr;           // Ebe*
    double *d_Ksave = nullptr;     // K of the predictor step (amaru_tangent_save) for K = a1*K + a2*K2 of the ME / Ralston schemes
    bool blended = false;          // d_K holds such a blend: the per-IP tangent data no longer describes it -> CSR products
    bool op_ebe = true;            // CG products: element-by-element (default) or block-CSR SpMV (AMARU_OPERATOR=csr)

    bool cg_graph = true;          // replay the CG batches as a CUDA graph on one GPU (AMARU_CG_GRAPH=0 disables)

    // bookkeeping
    int64_t launches = 0;
    bool profiling = false;
    double prof_spmv_ms = 0.0;
    int64_t prof_spmv_n = 0;
    std::vector<cudaEvent_t> ev_pool;

    // host copies kept for get_csr
    std::vector<int32_t> h_rowptr, h_col, h_eqid;
    std::vector<uint8_t> h_fixed;
};

// ---- internals of abi.cu shared with group.cu ------------------------------------------------------------------
struct CreateArgs {
    int ndim, stressmodel;
    double thickness;
    int64_t nnodes, nowned;
    const double *coords;
    int nbatches;
    const int32_t *batch_shape;
    const int64_t *batch_nelem;
    const int32_t *conn;
    const int32_t *elem_mat;
    int nmats;
    const int32_t *mat_kind;
    const double *mat_params;
    const int32_t *eqid;
    const uint8_t *prescribed;   // optional (partitioned); else eqid >= nu
    int64_t ndofs, nu;
    int device;
    int rank, nranks;
};
struct SolveInfo;
amaru_model *amaru_create_impl(const CreateArgs &a);
void amaru_free_model(amaru_model *m);
int amaru_solve_device(amaru_model *m, double cg_rtol, int cg_maxit, int precond, SolveInfo &info);
int amaru_update_device(amaru_model *m);
const char *amaru_status_text(int st);
int amaru_read_status(amaru_model *m);
void amaru_reset_status(amaru_model *m);
void amaru_set_element_ownership(amaru_model *m, const int32_t *conn, int nneigh, const int32_t *neigh_rank,
                                 const int64_t *recv_start, const int64_t *recv_count);
// group.cu: multi-GPU handle of one process (h->grp != nullptr)
void amaru_group_barrier(amaru_model *part);   // stream sync + host barrier over the parts (throws if a part failed)
int amaru_group_create(const CreateArgs &a, int ngpus, const int32_t *devices, int partitioner, amaru_model **out, char *msg, int msglen);
int amaru_group_destroy(amaru_model *h);
int64_t amaru_group_sum(const amaru_model *h, int what /*0 blocks, 1 launches, 2 colours*/);
amaru_model *amaru_group_part(const amaru_model *h, int r);
int amaru_group_state(amaru_model *h, bool set, double *sigma, double *eps, double *epa, double *dlam, char *msg, int msglen);
int amaru_group_simple(amaru_model *h, int what /*0 backup 1 restore 2 assemble_K 3 system matrix 4 operator 5 profiling 6 tangent_save 7 tangent_blend*/, double a, double b, char *msg, int msglen);
int amaru_group_assemble_M(amaru_model *h, const double *rho, char *msg, int msglen);
int amaru_group_solve(amaru_model *h, double *U, double *F, double cg_rtol, int cg_maxit, int precond, int *iters, double *relres, char *msg, int msglen);
int amaru_group_update(amaru_model *h, const double *dU, double *dFin, int mode, char *msg, int msglen);
int amaru_group_product(amaru_model *h, int what, double a, double b, const double *x, double *y, int masked, double *pAp, char *msg, int msglen);
int amaru_group_set_device_vectors(amaru_model *h, const double *U, const double *F, char *msg, int msglen);
int amaru_group_newton_iteration(amaru_model *h, double cg_rtol, int cg_maxit, int precond, double *phase_ms, int *iters, double *relres, char *msg, int msglen);
int amaru_group_comm_selftest(amaru_model *h, int skip_rank, char *msg, int msglen);
void amaru_group_refresh_output_state(amaru_model *h);   // wrapper's IP state planes <- owners' copies (before recovery)

// ---- kernels' host entry points (one per .cu) -----------------------------------------------------------------
void amaru_build_map(amaru_model *m, Batch &b);                     // assemble.cu
void amaru_launch_assemble(amaru_model *m, int what /*0 K, 1 M*/);   // assemble.cu
void amaru_launch_update(amaru_model *m, const double *d_dU_nodes, double *d_f_nodes, int mode);  // update.cu
void amaru_state_permute(amaru_model *m, double *d_io, int plane0, int ncomp, bool to_device);    // update.cu

struct SolveInfo {
    int iters = 0;
    double relres = 0.0;
    double maxabs = 0.0;
    bool converged = false;
};
void amaru_pcg_setup(amaru_model *m);                                // pcg.cu (allocations)
void amaru_spmv_setup(amaru_model *m);                               // spmv.cu (tiles of the streamed SpMV)
void amaru_spmv_launch(amaru_model *m, const double *A, const double *x, double *y, int mask, int dot, int check_done,
                       int finalize);                                 // spmv.cu
// CG-loop product with the symmetric-storage kernel (m->use_sym): y = A x from the upper blocks of the current d_A
void amaru_spmv_sym_launch(amaru_model *m, const double *x, double *y, int mask, int dot, int check_done, int finalize);
void amaru_spmv_sym_refresh(amaru_model *m);                          // spmv.cu: d_Asym <- upper blocks of d_A
void amaru_pcg_solve(amaru_model *m, double rtol, int maxit, int precond, SolveInfo &info);  // pcg.cu
void amaru_spmv(amaru_model *m, const double *A, const double *x, double *y, int mask_mode);  // pcg.cu
void amaru_eq_to_nodes(amaru_model *m, const double *d_eq, double *d_nodes);                  // pcg.cu
void amaru_nodes_to_eq(amaru_model *m, const double *d_nodes, double *d_eq, int which /*0 all,1 free,2 fixed*/);
void amaru_combine_matrix(amaru_model *m);                            // pcg.cu: d_A = a*K + b*M
int amaru_check_nan(amaru_model *m, const double *d_v, int64_t n);    // pcg.cu
void amaru_zero_free(amaru_model *m, double *x);                     // pcg.cu
void amaru_axpby(amaru_model *m, int64_t n, double a, const double *x, double b, const double *y, double *out);  // pcg.cu
double amaru_operator_product(amaru_model *m, int masked);           // pcg.cu: d_q = A d_p with the CG operator (+ p.Ap)
void amaru_time_cg_kernel(amaru_model *m, int kind, int precond, int reps);  // pcg.cu

// matrix-free operator (ebe.cu)
void amaru_ebe_setup(amaru_model *m);
void amaru_ebe_destroy(amaru_model *m);
void amaru_ebe_refresh(amaru_model *m);                               // tangent planes <- current IP state
void amaru_ebe_set_owned(amaru_model *m, int batch, const uint8_t *h_owned_sorted);
void amaru_ebe_apply(amaru_model *m, const double *x, double *y, int mask, int dot, int check_done, int finalize, int fused = 0);
void amaru_ebe_begin(amaru_model *m);                                 // before a solve: fresh epochs / tickets of the patch form
void amaru_ebe_patch_stats(const amaru_model *m, int64_t *npatch, int64_t *nslots, int64_t *pnodes, int64_t *pnodes_loaded);
int64_t amaru_ebe_bytes(const amaru_model *m);
const char *amaru_ebe_kernel(const amaru_model *m);

// halo exchange (halo.cu) — no-ops for nranks == 1
void amaru_halo_exchange(amaru_model *m, double *d_v);
void amaru_allreduce_sum(amaru_model *m, double *d_vals, int64_t n);
void amaru_allreduce_max(amaru_model *m, double *d_vals, int64_t n);
void amaru_allreduce_max_int(amaru_model *m, int *d_val);
void amaru_comm_setup(amaru_model *m, int nneigh, const int32_t *neigh_rank, const int64_t *send_ptr,
                      const int32_t *send_nodes, const int64_t *recv_start, const int64_t *recv_count, const void *uid);
void amaru_comm_destroy(amaru_model *m);
bool amaru_comm_is_p2p(const amaru_model *m);                        // CG-loop exchanges run as peer-memory kernels
void amaru_comm_check(amaru_model *m);                               // throws AMARU_ERR_COMM if a peer-memory wait timed out
bool amaru_comm_fused(const amaru_model *m);                         // exchanges of the CG loop may be folded into its kernels (p2p.cuh)
void amaru_halo_push(amaru_model *m, double *d_v);                   // producer half of a halo exchange of p (fused loop)
bool amaru_ebe_fusable(const amaru_model *m);                        // the operator kernels carry the hooks of the fused multi-GPU loop
void amaru_p2p_connect_direct(amaru_model *const *parts, int n);    // in-process peers (amaru_create with ngpus > 1)
void amaru_recovery_destroy(amaru_model *m);   // recovery.cu
