/* efgpu.h - C-ABI of the B200-native Hierarchical Poincare-Steklov hot path.
 *
 * This is the drop-in boundary for EllipticForest's HPS path.  The
 * reference has no FFI today: the path lives behind the C++ template
 *   HPSAlgorithm<FiniteVolumeGrid, FiniteVolumeSolver, FiniteVolumePatch, double>
 * (reference src/HPSAlgorithm.hpp:26-82).  Each entry point below names the reference
 * interface it replaces; INTEGRATION.md shows the subclass a maintainer adds on the reference
 * side (stage overrides that marshal to these calls).
 *
 * Conventions (identical to the reference):
 *   - node table in p4est depth-first pre-order, root = node 0, children in Morton order
 *     0 = lower-left, 1 = lower-right, 2 = upper-left, 3 = upper-right
 *     (src/Quadtree.hpp:118-196, src/P4est.cpp:35-44, FiniteVolumeNodeFactory.cpp:38-57);
 *   - leaves are numbered in that order (= the order of traversePreOrder over leaves, quirk q9);
 *   - cell index inside a patch: j + i*ny, i = x index (FiniteVolumeSolver.cpp:17-19);
 *   - boundary vectors / operator rows in WESN order, each side low -> high (PatchSolver.hpp:36);
 *   - Neumann data are coordinate derivatives, not outward normals (FiniteVolumeSolver.cpp:332-343);
 *   - all matrices row-major double.
 *
 * Every function returns an int status (0 = ok).  All calls are synchronous on the calling
 * thread unless the name ends in _device and `sync` is 0.  Host arrays are borrowed for the
 * duration of the call only.  There is NO CPU fallback: efgpu_create fails with EFGPU_ERR_CUDA
 * when no device is present.
 */
#ifndef EFGPU_H_
#define EFGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
enum {
    EFGPU_OK = 0,
    EFGPU_ERR_CUDA = 1,        /* CUDA runtime error (or no device) */
    EFGPU_ERR_BAD_ARG = 2,     /* -> std::invalid_argument in the C++ shim */
    EFGPU_ERR_BAD_SHAPE = 3,   /* -> std::invalid_argument (reference: HPSAlgorithm.hpp:761, Matrix.hpp:322) */
    EFGPU_ERR_OOM = 4,
    EFGPU_ERR_SINGULAR = 5,    /* zero / non-finite pivot (reference: LAPACK INFO > 0 warning, Matrix.hpp:898,947) */
    EFGPU_ERR_STATE = 6,       /* stage called out of order */
    EFGPU_ERR_UNSUPPORTED = 7
};

/* build / stage flags */
enum {
    EFGPU_CACHE_OPERATORS = 1u, /* option "cache-operators" (HPSAlgorithm.hpp:134-139): one T_leaf for every leaf */
    EFGPU_HOMOGENEOUS_RHS = 2u, /* option "homogeneous-rhs" (HPSAlgorithm.hpp:532,587,1202) */
    EFGPU_KEEP_X = 4u,          /* parity/debug: retain a copy of X (the product only needs X^-1) */
    EFGPU_LEAN_T = 8u,          /* memory policy: a DtN map T is only read by the parent's merge4to1
                                   (HPSAlgorithm.hpp:497-518), so the maps of one tree level share a transient arena that is
                                   reused two levels up; after the build only leaf and root T can be read back.  Halves the
                                   resident operator bytes (X^-1 + S + H stay: 512 n^2 B per merge instead of 1024 n^2). */
    EFGPU_NO_SYMMETRY = 16u,    /* always use the general merge plan.  By default a merge whose subtree consists of square,
                                   uncoarsened patches with constant-coefficient (FISHPACK90) leaves uses the symmetry of X and of diag(d) T (d = -1 on W and S: the
                                   coordinate-derivative convention of FiniteVolumeSolver.cpp:332-343): 4 instead of 6 products
                                   per level of the block inversion and 36 instead of 64 block products for T. */
    EFGPU_LAZY_ROOT_DTN = 32u   /* the DtN map of the whole domain (tree level 0) is read by nothing on the Dirichlet path - no parent merges it, the
                                   upwards and solve sweeps use X^-1, H and S only (HPSAlgorithm.hpp:529-577) - yet it is 22 % of the flops of a
                                   uniform build.  With this flag efgpu_build leaves its block products unissued; the first reader (efgpu_solve_robin
                                   with b != 0, efgpu_get_operator / efgpu_operator_device of the root's T, efgpu_complete_root_dtn) forms it then,
                                   with identical results.  Off by default: the reference's buildStage always forms it (mergeT_ :940-968). */
};

enum { EFGPU_LEAF_CONSTANT = 0, EFGPU_LEAF_VARIABLE = 1 };

/* operator / vector selectors for the parity accessors */
enum { EFGPU_OP_T = 0, EFGPU_OP_S = 1, EFGPU_OP_X = 2, EFGPU_OP_H = 3, EFGPU_OP_XINV = 4, EFGPU_OP_T_UNCOARSENED = 5 };
enum { EFGPU_VEC_H = 0, EFGPU_VEC_W = 1, EFGPU_VEC_G = 2, EFGPU_VEC_U = 3, EFGPU_VEC_F = 4, EFGPU_VEC_H_UNCOARSENED = 5 };

/* kernel classes of the optional per-launch profiling (efgpu_set_profiling / efgpu_get_profile) */
enum {
    EFGPU_PROF_LEAF_DTN = 0, EFGPU_PROF_COARSEN_T = 1, EFGPU_PROF_ASSEMBLE = 2, EFGPU_PROF_INVERT_SMALL = 3,
    EFGPU_PROF_GEMM_XINV = 4, EFGPU_PROF_GEMM_S = 5, EFGPU_PROF_GEMM_T = 6, EFGPU_PROF_LEAF_SOLVE = 7,
    EFGPU_PROF_UPWARDS_MATVEC = 8, EFGPU_PROF_SOLVE_MATVEC = 9, EFGPU_PROF_COARSEN_VEC = 10, EFGPU_PROF_LEAF_LU = 11,
    EFGPU_PROF_ALLGATHER = 12, EFGPU_PROF_TRANSPOSE = 13, EFGPU_PROF_MIRROR_T = 14, EFGPU_PROF_NCLASSES = 15
};

typedef struct efgpu_handle efgpu_handle;

/* Flat tree table = what Quadtree<FiniteVolumePatch>::traversePreOrder visits
 * (src/Quadtree.hpp:236-260): one row per node. */
typedef struct {
    int32_t n_nodes;
    int32_t nx;             /* cells per side of a leaf patch (nx == ny; 8, 16, 24 or 32) */
    const int32_t* level;   /* [n_nodes] */
    const int32_t* child;   /* [n_nodes*4] node ids of the children, -1 for a leaf */
    const double* box;      /* [n_nodes*4] x_lower, x_upper, y_lower, y_upper of the node's grid (FiniteVolumeGrid) */
} efgpu_tree_desc;

typedef struct {
    double dofs;                   /* n_leaves * nx * ny  (examples/elliptic-multiple/main.cpp:374) */
    double n_leaves, n_nodes;
    double build_ms, upwards_ms, solve_ms;   /* CUDA-event time of the last call of each stage (device work only) */
    double merge_flops_canonical;  /* sum 810.67 n^3: the dgesv + dgemm work of the reference */
    double merge_flops_issued;     /* sum 512 n^3: flops this implementation sends to the FP64 tensor pipe */
    double upwards_bytes, solve_bytes;       /* algorithmic HBM bytes per upwards / solve call */
    double device_bytes;           /* device memory held by the handle */
    double min_pivot;              /* smallest |pivot| met while inverting the merge matrices */
    /* conditioning report of the UNPIVOTED inversion of the merge matrices (the reference: dgesv with partial pivoting,
       Matrix.hpp:915-953): largest |pivot|, the smallest min/max |pivot| ratio inside one base-case block (<= 128 rows), and the
       number of negative pivots (0 for the SPD merge matrices of lambda <= 0; > 0 marks an indefinite problem).  efgpu_build
       returns EFGPU_ERR_SINGULAR when pivot_ratio_min < 1e-10 (EFGPU_PIVOT_RATIO_LIMIT overrides the 1e10). */
    double max_pivot, pivot_ratio_min, negative_pivots;
    /* indefinite problems (lambda > 0, see efgpu_set_refine_inverse): max |(I - X X^-1)_ij| over every merge BEFORE the one
       Newton-Schulz step that squares it; -1 when the build did not refine */
    double inverse_residual;
} efgpu_stats_t;

/* ---- lifetime: replaces HPSAlgorithm ctor (HPSAlgorithm.hpp:78-82) + the Quadtree walk ---------- */
int efgpu_create(const efgpu_tree_desc* desc, int device, efgpu_handle** out);
/* Sharded runs (replaces the reference's rank-shared upper tree, Quadtree.hpp:146-151,464-507, where child
 * patches arrive by MPI::broadcast(Node), QuadNode.hpp:191-199).  The node table may be a FOREST (several nodes
 * without a parent: the subtrees one GPU owns), and with external_leaf_size != NULL (one entry per leaf, cells per
 * side) the leaves are not finite-volume patches but already-merged subtrees whose DtN map / particular Neumann
 * data the caller writes into the device views below (e.g. as the target of an NCCL receive) before
 * efgpu_build / efgpu_upwards_device, and whose Dirichlet data it reads back after efgpu_solve_*_device. */
int efgpu_create_ex(const efgpu_tree_desc* desc, int device, const int32_t* external_leaf_size, efgpu_handle** out);
void efgpu_destroy(efgpu_handle* h);
const char* efgpu_last_error(const efgpu_handle* h);   /* h may be NULL after a failed create */

/* Replicated upper tree: every rank holds the same (external-leaf) tree and computes only rows
 * [rank * R / nranks, (rank+1) * R / nranks) of each merge's S (R = 4n) and T (R = 8n); the caller all-gathers the
 * contiguous row slices (views from efgpu_operator_device) between efgpu_build_level(level, 0) [coarsen, X, X^-1,
 * S rows] and efgpu_build_level(level, 1) [T rows], levels from the deepest up.  X^-1 is formed by every rank.
 * efgpu_build == begin; for each level: phase 0, phase 1; end.  Replaces the redundant whole-merge recomputation
 * on every sharing rank in the reference (Quadtree.hpp:504-506). */
int efgpu_set_partition(efgpu_handle* h, int rank, int nranks);
/* After a build the root's DtN map of a partitioned tree is left row-distributed (nothing on the Dirichlet path reads it; the
 * reference's merge computes it all the same, HPSAlgorithm.hpp:940-968).  This COLLECTIVE call (every rank of the partition)
 * gathers the row slices and completes the mirrored blocks of the symmetric plan; it is required before efgpu_solve_robin,
 * efgpu_get_operator / efgpu_operator_device of the root's T.  No-op on unpartitioned handles and when already complete. */
int efgpu_complete_root_dtn(efgpu_handle* h);
/* With a collective supplied by the caller the library performs every exchange of a partitioned tree itself, inside
 * efgpu_build / efgpu_build_level: the large products of the block inversion of X are split by rows as well (each rank
 * computes h / nranks rows of every h x h product, h >= 128 * nranks), and S / T row slices are gathered after their
 * phase.  fn must all-gather IN PLACE: rank r's bytes_per_rank bytes lie at buf + r * bytes_per_rank; it must be
 * stream-ordered with efgpu_stream(h) (e.g. ncclAllGather on that stream, or torch.distributed under that stream). */
typedef int (*efgpu_allgather_fn)(void* buf, size_t bytes_per_rank, void* user);
int efgpu_set_allgather(efgpu_handle* h, efgpu_allgather_fn fn, void* user);
/* Peer mode (preferred on one NVSwitch domain, <= 8 ranks): instead of a caller-supplied collective, every rank maps the shared
 * operator arena of every other rank (CUDA IPC) and the batched GEMMs store each finished tile of a row slice into all arenas
 * from their epilogue, between two device-side flag barriers - no NCCL call, no host round trip, transfer overlapped with the
 * tensor-core work.  Protocol, after efgpu_set_partition and before the first build / device view:
 *   efgpu_peer_export(h, mine)      allocates the arena, writes its 64-byte cudaIpcMemHandle_t;
 *   (the caller all-gathers the handles with whatever transport it has: MPI_Allgather in the reference, torch.distributed here)
 *   efgpu_peer_attach(h, all, n)    n * 64 bytes, rank order; opens every peer's arena.
 * efgpu_peer_broadcast copies a region of this rank's arena (any device view of the handle: leaf DtN maps, h, g vectors) to
 * the same offset on every other rank, stream-ordered; efgpu_peer_barrier is the flag barrier (collective, stream-ordered).
 * Callers bracket their broadcasts: barrier, broadcasts, barrier.  A rank that never arrives makes the barrier give up after
 * EFGPU_PEER_TIMEOUT_S (default 20) seconds and the stage return EFGPU_ERR_STATE.  Every rank must have synchronised its stream
 * after a common barrier before any rank destroys its handle.  Replaces MPI::broadcast(Node) of src/QuadNode.hpp:191-199. */
int efgpu_peer_export(efgpu_handle* h, void* ipc_handle_out);
int efgpu_peer_attach(efgpu_handle* h, const void* ipc_handles, int nranks);
int efgpu_peer_barrier(efgpu_handle* h);
int efgpu_peer_broadcast(efgpu_handle* h, const void* dev_ptr, size_t bytes);
/* External leaves (efgpu_create_ex): declare that the DtN maps the caller writes are signed-symmetric (diag(d) T symmetric),
 * e.g. because efgpu_is_symmetric() holds for the forest handle they were built by; enables the symmetric merge plan. */
int efgpu_set_symmetric_leaves(efgpu_handle* h, int on);
/* Pivoting policy (reference: dgesv with partial pivoting, Matrix.hpp:915-953; here: unpivoted recursive block inversion, safe for
 * the SPD / diagonally dominant merge matrices of lambda <= 0).  For indefinite problems - lambda > 0, which the reference
 * tolerates (hstcrt.f:450-452, FiniteVolumeSolver.cpp:270) - every X^-1 is refined by one Newton-Schulz step
 * X^-1 <- X^-1 + X^-1 (I - X X^-1) (two more (4n)^3 products per merge, X kept as under EFGPU_KEEP_X), which restores the accuracy
 * of the pivoted solve.  mode -1 (default): automatic, on when lambda > 0 (constant leaves) or any sampled lambda > 0 (variable
 * leaves); 0: never; 1: always.  Handles with external leaves cannot see lambda: the caller passes 1 when the forests refine. */
int efgpu_set_refine_inverse(efgpu_handle* h, int mode);
int efgpu_is_symmetric(const efgpu_handle* h);   /* after a build: 1 when every root's DtN map was built by the symmetric plan */
int efgpu_build_begin(efgpu_handle* h, unsigned flags);
int efgpu_build_level(efgpu_handle* h, int level, int phase);
int efgpu_build_end(efgpu_handle* h);
int efgpu_max_level(const efgpu_handle* h);

/* ---- leaf model: replaces FiniteVolumeSolver's public fields (FiniteVolumeSolver.hpp:64-88) ---- */
/* solver_type = FISHPACK90: constant coefficients, lambda = lambda_function(0,0) (FiniteVolumeSolver.cpp:254) */
int efgpu_set_leaf_constant(efgpu_handle* h, double lambda);
/* solver_type = FivePointStencil: coefficients sampled by the caller exactly where
 * FiniteVolumeSolver.cpp:63-79 samples them; arrays are leaf-major, n_leaves * nx * ny, cell index j + i*ny.
 * alpha, lambda at cell centres; beta at the W, E, S, N face midpoints of each cell. */
int efgpu_set_leaf_variable(efgpu_handle* h, const double* alpha, const double* beta_w, const double* beta_e,
                            const double* beta_s, const double* beta_n, const double* lambda);

/* the same arrays already resident in device memory (e.g. evaluated by the caller's own device code on the coordinates of
 * efgpu_leaf_points_device); copied, borrowed for the call only */
int efgpu_set_leaf_variable_device(efgpu_handle* h, const double* alpha_dev, const double* beta_w_dev, const double* beta_e_dev,
                                   const double* beta_s_dev, const double* beta_n_dev, const double* lambda_dev);

/* ---- stages ------------------------------------------------------------------------------------ */
/* buildStage (HPSAlgorithm.hpp:120-161): leaf buildD2N + every merge4to1. */
int efgpu_build(efgpu_handle* h, unsigned flags);
/* Adaptive re-build (the capability paper.md:44 advertises and the reference leaves unimplemented:
 * isBuilt is never read, src/HPSAlgorithm.hpp:50-55).  `h` is a fresh handle of the CHANGED mesh (efgpu_create + leaf model), `old` a
 * built handle of the previous mesh on the same device (it stays valid).  Nodes of the new tree whose whole subtree is unchanged
 * (same boxes bit for bit, same structure) take their X^-1, S, H and DtN map from `old` by device-to-device copies; only the
 * other merges - the ancestor chains of what was refined or coarsened - are computed.  Result bit-identical to efgpu_build(h).
 * reused / rebuilt (may be NULL): number of merges copied / computed; efgpu_get_stats(h).merge_flops_issued counts the computed
 * ones only.  Plain handles only (no partition, no external leaves, no EFGPU_LEAN_T); same leaf model and flags as the old build. */
int efgpu_rebuild_from(efgpu_handle* h, efgpu_handle* old, unsigned flags, double* reused_merges, double* rebuilt_merges);
/* upwardsStage (HPSAlgorithm.hpp:178-272): f_leaves = vectorF of every leaf (n_leaves*nx*ny), scaled by fscale. */
int efgpu_upwards(efgpu_handle* h, const double* f_leaves, double fscale, unsigned flags);
int efgpu_upwards_device(efgpu_handle* h, const double* f_leaves_dev, double fscale, unsigned flags, int sync);
/* solveStage(fn(Patch&)) (HPSAlgorithm.hpp:291-324): g_root = root vectorG (4 * root size), u_leaves = vectorU of every leaf. */
int efgpu_solve_dirichlet(efgpu_handle* h, const double* g_root, unsigned flags, double* u_leaves);
int efgpu_solve_dirichlet_device(efgpu_handle* h, const double* g_root_dev, unsigned flags, double* u_leaves_dev, int sync);
/* solveStage(fn(side,x,y,*a,*b)) (HPSAlgorithm.hpp:343-445): a u + b du/dn = r sampled at the root grid. */
/* solve with the Dirichlet data of every root already written into its EFGPU_VEC_G device view (forest handles) */
int efgpu_solve_from_roots_device(efgpu_handle* h, unsigned flags, double* u_leaves_dev, int sync);
/* device views of a node's operators / vectors (valid until efgpu_destroy; allocate the handle's memory on first use).
 * which: T, T_UNCOARSENED, S, XINV / any EFGPU_VEC_*.  Work of the handle is ordered on efgpu_stream(h). */
int efgpu_operator_device(efgpu_handle* h, int node, int which, double** ptr, int* rows, int* cols);
int efgpu_vector_device(efgpu_handle* h, int node, int which, double** ptr, int* len);
int efgpu_solve_robin(efgpu_handle* h, const double* a, const double* b, const double* r, unsigned flags, double* u_leaves);
int efgpu_sync(efgpu_handle* h);

/* ---- callers either side of the path ------------------------------------
 * Sampling coordinates of every leaf, leaf-major, cell index j + i*ny: what grid(XDIM, i), grid(YDIM, j) yield in the
 * reference's sampling loops (HPSAlgorithm.hpp:241-249: load at the cell centres; FiniteVolumeSolver.cpp:63-79: alpha and
 * lambda at the centres, beta at the four face midpoints of each cell).  The reference calls a std::function per point on
 * the host; here the caller evaluates its functions on these arrays with its own (device) code and passes the results to
 * efgpu_upwards_device / efgpu_set_leaf_variable_device.  x or y may be NULL.  n_leaves * nx * ny doubles each. */
enum { EFGPU_POINTS_CENTRE = 0, EFGPU_POINTS_FACE_W = 1, EFGPU_POINTS_FACE_E = 2, EFGPU_POINTS_FACE_S = 3, EFGPU_POINTS_FACE_N = 4 };
int efgpu_leaf_points_device(efgpu_handle* h, int which, double* x_dev, double* y_dev, int sync);
int efgpu_leaf_points(efgpu_handle* h, int which, double* x, double* y);   /* host arrays */
/* The error norms of the reference's drivers (examples/elliptic-multiple/main.cpp:346-371) reduced on the device:
 *   l1 = sum dx dy |u - exact| / area,  l2 = sqrt(sum dx dy (u - exact)^2 / area),  linf = max |u - exact|
 * over every cell of every leaf; area = area of the root patch(es).  u_dev == NULL: the solution of the last solve stage,
 * resident in the handle.  Deterministic (fixed reduction order); any of l1 / l2 / linf may be NULL. */
int efgpu_error_norms_device(efgpu_handle* h, const double* u_dev, const double* exact_dev, double* l1, double* l2, double* linf);
int efgpu_error_norms(efgpu_handle* h, const double* exact, double* l1, double* l2, double* linf);   /* exact: host array */
/* Mesh and cell fields as ONE binary .vtu (UnstructuredGrid, raw appended data, UInt64 headers, little endian), streamed from
 * device buffers: replaces Mesh::setMeshFromQuadtree + UnstructuredGridVTK::toVTK (src/Mesh.hpp:186-267, src/VTK.cpp:237-300),
 * which format every number as ASCII on the host.  Same cells, same point order (four own corners per leaf cell, leaves in
 * traversePreOrder, cells i-slow j-fast), same corner formulas.  fields_dev[k]: one double per cell, leaf-major, cell index
 * j + i*ny (the layout of vectorU / vectorF); NULL = the solution of the last solve stage.  SURVEY.md 8(f) rank 3. */
int efgpu_write_vtu(efgpu_handle* h, const char* path, int n_fields, const char* const* names, const double* const* fields_dev);
void* efgpu_stream(efgpu_handle* h);   /* the cudaStream_t all work of this handle is issued on */
int efgpu_set_stream(efgpu_handle* h, void* stream);   /* adopt a caller-owned cudaStream_t (e.g. share one stream between two handles) */

/* ---- parity accessors: replace reading patch.matrixT()/S()/X()/H(), vectorH()/W()/G()/U() (Patch.hpp:103-171) */
int efgpu_node_info(const efgpu_handle* h, int node, int* size, int* n_coarsens, int* is_leaf, int* leaf_index);
int efgpu_operator_shape(const efgpu_handle* h, int node, int which, int* rows, int* cols);
int efgpu_get_operator(efgpu_handle* h, int node, int which, double* out, size_t capacity);
int efgpu_vector_length(const efgpu_handle* h, int node, int which, int* len);
int efgpu_get_vector(efgpu_handle* h, int node, int which, double* out, size_t capacity);
int efgpu_get_stats(const efgpu_handle* h, efgpu_stats_t* out);
/* Per-kernel-class device time (CUDA events on the handle's stream around every launch group) and launch
 * counts, accumulated since profiling was switched on; replaces app.timers (EllipticForestApp.hpp:39-137),
 * which only has whole-stage wall clocks.  Launch counts are kept even when profiling is off. */
int efgpu_set_profiling(efgpu_handle* h, int on);
int efgpu_get_profile(const efgpu_handle* h, int cls, double* ms, double* launches);
const char* efgpu_profile_class_name(int cls);

/* ---- host-side mesh: replaces Mesh::refineByFunction + Quadtree ctor + FiniteVolumeNodeFactory
 * (src/Mesh.hpp:111-180, src/Quadtree.hpp:118-196, FiniteVolumeNodeFactory.cpp:15-67).  p4est itself
 * stays on the host in the reference; this builder reproduces its result for the single-tree square
 * connectivity (uniform min_level start, recursive refine by callback, 2:1 corner balance, Morton order)
 * so that the library is usable stand-alone.  A reference-side caller passes its own p4est-derived table. */
typedef int (*efgpu_refine_fn)(double x, double y, void* user);   /* non-zero: refine the patch containing (x,y) */
typedef struct efgpu_mesh efgpu_mesh;
/* built-in callback: the indicator of examples/elliptic-single/main.cpp:165-175, |-(sin x + sin y)| > *(const double*)user
 * (saves the per-point trip through a foreign-language callback when large adaptive meshes are built from Python) */
int efgpu_refine_elliptic_single(double x, double y, void* user);
int efgpu_mesh_create(double x_lower, double x_upper, double y_lower, double y_upper, int nx, int min_level,
                      int max_level, efgpu_refine_fn fn, void* user, efgpu_mesh** out);
int efgpu_mesh_desc(const efgpu_mesh* m, efgpu_tree_desc* out);   /* pointers stay owned by the mesh */
int efgpu_mesh_n_leaves(const efgpu_mesh* m);
const int32_t* efgpu_mesh_leaf_nodes(const efgpu_mesh* m);       /* node id of each leaf, pre-order */
int efgpu_mesh_path(const efgpu_mesh* m, int node, char* buf, size_t capacity);   /* "0" + child ids (P4est.cpp:35-44) */
void efgpu_mesh_destroy(efgpu_mesh* m);

/* ---- host-side plan of one merge batch, for CPU tests of the descriptor logic (no device needed) ----
 * Serialises the step list the build would run for a merge with child side n on tree level `level`, as seen by rank
 * `rank` of `nranks` (row partition), general (symmetric = 0) or symmetric plan.  Every record is 16 int64:
 *   steps : kind, first, count, off, N, cls, gk, g_op, g_rows, g_cols, g_ld, g_off
 *   blocks: c_op, c0_op, ldc, ldc0, c_off, c0_off, rows, cols, nterms, ct_op + 1 (0: none), ldct, ct_off, ct_neg (the result's signed
 *           transpose is stored there as well)   then per term (2 records of 8 int64 follow
 *           in `terms`): a_op, b_op, lda, ldb, a_off, b_off, K, neg
 *   trans : src_op, dst_op, lds, ldd, src_off, dst_off, rows, cols, neg
 * ws[3] receives the per-entry workspace sizes (doubles) of W1, W2 and W3.  Returns the counts through n_*; arrays may
 * be NULL to query the counts only. */
int efgpu_debug_merge_plan(int n, int level, int rank, int nranks, int symmetric, int64_t* steps, int* n_steps,
                           int64_t* blocks, int64_t* terms, int* n_blocks, int64_t* trans, int* n_trans, int64_t* ws);

/* the same for the plan of a peer-mapped tree (peer != 0: split products of the inversion have gk = 3 and keep their destination;
 * their row slices, those of S and of the DtN maps below the root are stored into every rank's arena by the producing GEMM) */
int efgpu_debug_merge_plan_ex(int n, int level, int rank, int nranks, int symmetric, int peer, int64_t* steps, int* n_steps,
                              int64_t* blocks, int64_t* terms, int* n_blocks, int64_t* trans, int* n_trans, int64_t* ws);

/* TMA side of the same plan (operand staging of the 128-row products, csrc/gemm_tma.cu): `views` (3 per view: operand slot, origin
 * inside the slot, leading dimension), `tblocks` (12 per block descriptor, same indexing as `blocks`: for each of the two terms a_view,
 * b_view, a_row, a_col, b_row, b_col; view -1 = the block keeps the cp.async kernel) and `step_tma` (one per step: 1 when every block
 * of the step has views).  Arrays may be NULL to query n_views only. */
int efgpu_debug_tma_plan(int n, int level, int rank, int nranks, int symmetric, int peer, int64_t* views, int* n_views,
                         int64_t* tblocks, int64_t* step_tma);

/* Process-wide kernel-selection knobs for measurements (A/B runs of the bandwidth-bound kernels): key 0 = matvec kernels
 * for rows of <= 256 doubles (2, default: row-batch kernels; 0: one row per warp, as for longer rows); key 1 = long-row
 * kernel of the compact H (0, default: 8 loads in flight per lane; 1: 4); key 2 = CTAs per SM the long-row launcher aims
 * for (0: default 16); key 3 = leaf solve of constant-coefficient leaves (0, default: FP64 tensor-core kernel, one warp per
 * leaf; 1: one thread per cell); key 5 = symmetric merge plan, diagonal blocks of T: 1 = multiply only the upper triangle of
 * their 2 x 2 / 4 x 4 sub-blocks and mirror the rest (1, default; 0: whole blocks; read when a handle is created / partitioned);
 * key 4 = transposed second destination of a GEMM block (0, default: through shared memory as whole rows where peer arenas receive
 * it, direct stores otherwise; 1: always through shared memory; 2: always direct); key 6 = variable-coefficient leaves of 8 / 16
 * cells (0, default: warp-level / tensor-core kernels; 1: CTA-per-leaf kernels); key 7 = 128 x 128 base case of the block inversion
 * (0, default: blocked Gauss-Jordan on the tensor pipe; 1: per-pivot register kernel; 2: blocked with look-ahead); key 8 = operand
 * staging of the 128-row GEMM tiles (1, default: TMA; 0: cp.async); key 9 = 256 x 256 base case by a thread-block cluster in batches
 * of at most four merges (0, default: off; read when a plan is made); key 10 = pivot reciprocals of the base case by hardware seed +
 * Newton steps (0, default: IEEE division); key 11 = peer-mapped partitions over 2 / 4 / 8 ranks split S and T by block columns
 * instead of rows (0, default: rows; read when a plan is made).  Keys 0 .. 15 are accepted.
 * Results do not depend on the knobs beyond floating-point summation order. */
int efgpu_set_tuning(int key, int value);

/* ---- stand-alone access to the GEMM kernel for unit tests and roofline measurements ------------ */
int efgpu_dgemm_batched(const double* A_dev, const double* B_dev, double* C_dev, int m, int n, int k, int batch,
                        int tile, int iters, float* ms_per_iter);
/* The same product with TMA-staged operands (cp.async.bulk.tensor.2d + mbarrier ring, 128-byte swizzle; `stages` = 3, 4 or 6: 128 x 128 CTA
 * tile, that many stages; 64: 128 x 64 CTA tile, four stages, two CTAs per SM): the A/B experiment behind DESIGN.md section 4's choice of
 * operand path.  m % 128 == n % 128 == k % 16 == 0.  Bit-identical
 * to efgpu_dgemm_batched (same k order). */
int efgpu_dgemm_batched_tma(const double* A_dev, const double* B_dev, double* C_dev, int m, int n, int k, int batch,
                            int stages, int iters, float* ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* EFGPU_H_ */
