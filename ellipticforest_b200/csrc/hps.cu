// Host-side engine behind the C-ABI (include/efgpu.h): level-synchronous plan of the quadtree,
// device memory layout, and the three stages of the HPS method (build / upwards / solve).
//
// Mirrors src/HPSAlgorithm.hpp of the reference: buildStage :120-161 (post-order leaf DtN +
// merge4to1 :497-518), upwardsStage :225-272 (+ upwards4to1 :529-551), solveStage :291-445
// (+ split1to4 :562-577, leafSolve :584-596).  The reference walks the tree node by node in p4est
// DFS order; here every tree level becomes a handful of batched kernel launches (parents grouped
// by child side n), which produces the same per-node operators because nodes of one level are
// independent.  Node numbering, leaf order and all index conventions are the reference's.
#include "common.cuh"
#include "kernels.cuh"
#include "../../include/efgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>
#include <unistd.h>

namespace efgpu {

// ---- merge topology (host copies of the tables in vec.cu) ------------------------------------
static const int h_iface[4][4] = {{3, -1, 1, -1}, {-1, 3, 0, -1}, {2, -1, -1, 1}, {-1, 2, -1, 0}};
static const int h_sgn[4][4] = {{-1, 0, -1, 0}, {0, -1, 1, 0}, {1, 0, 0, -1}, {0, 1, 0, 1}};
static const int h_tau_side[4][2] = {{0, 2}, {1, 2}, {0, 3}, {1, 3}};
static const int h_kk[4][2] = {{0, 2}, {1, 2}, {0, 3}, {1, 3}};
static const int h_pos[8] = {0, 4, 2, 5, 1, 6, 3, 7};  // inverse of pi = {0,4,2,6,1,3,5,7}

enum { OP_TC0 = 0, OP_XINV = 4, OP_S = 5, OP_T = 6, OP_W1 = 7, OP_W2 = 8, OP_W3 = 9, OP_XCOPY = 10, NOPS = 11 };

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    bool owned = true;   // false: a slice of the handle's shared arena (peer mode), never freed here
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), owned(o.owned) { o.p = nullptr; o.bytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; bytes = o.bytes; owned = o.owned; o.p = nullptr; o.bytes = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p && owned) cudaFree(p); p = nullptr; bytes = 0; owned = true; }
    void adopt(void* q, size_t b) { release(); p = q; bytes = b; owned = false; }
    void alloc(size_t b) {
        if (b <= bytes && p && owned) return;
        release();
        if (b == 0) return;
        EF_CUDA(cudaMalloc(&p, b)); bytes = b;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    template <class T> void upload(const std::vector<T>& v, cudaStream_t s) {
        alloc(v.size() * sizeof(T));
        if (!v.empty()) EF_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

struct NodeH {
    int level = 0, parent = -1, child[4] = {-1, -1, -1, -1};
    bool leaf = true;
    double box[4] = {0, 0, 0, 0};
    int size = 0;       // cells per side of this node's own patch (leaf: M, parent: 2 n)
    int ncoarsen = 0;   // times it is coarsened when merged into its parent (HPSAlgorithm.hpp:676-741)
    int leaf_idx = -1, batch = -1, slot = -1;
    bool symcand = false;   // see BatchH::symcand
    std::vector<double*> Tbuf;        // [0] own DtN, [t] after t coarsening steps
    std::vector<size_t> hbuf, gbuf;   // offsets into the vector arena, same indexing
    size_t w_off = 0, hd_off = 0;
};

struct Step {
    int kind; int first, count; long long off; int N; int cls;   // kind 0: small inverse, 1: gemm, 2: block transposes, 3: OP_T scratch E <- I + E (off: element offset, N: order), 4: Xinv <- Xinv + scratch (off); cls: profiling class
    long long off2 = -1;              // kind 0: a second, independent N x N block inverted by the same launch (-1: none)
    // row-partitioned step of a replicated tree: after the GEMM the g_rows x g_cols result is all-gathered over the ranks.
    // gk 1: the destination (op g_op, offset g_off) is contiguous (ld == g_cols): in place.  gk 2: the GEMM wrote its rows into
    // the staging block OP_W3 (ld = g_cols); after the all-gather the block is copied to op g_op, offset g_off, leading dimension g_ld.
    int gk = 0, g_op = 0, g_rows = 0, g_cols = 0, g_ld = 0; long long g_off = 0;
    bool tma = false;                 // kind 1: every block of the step has tensor-map views (plan_tma): operands can be staged by TMA
};

struct BatchH {
    int level = 0, n = 0, count = 0;
    int lane = 0;                     // stream lane of this batch inside its level (batches of one level are independent), 0 = the handle's stream
    std::vector<int> parents;
    std::vector<GemmBlock> blocks;    // all descriptors of this batch (host copy), both plans
    std::vector<TransOp> trans;       // block transposes of the symmetric plan
    std::vector<Step> steps;          // general plan: inversion, then S, then T
    std::vector<Step> steps_sym;      // plan for symmetric merge matrices (symcand batches only)
    std::vector<Step> steps_refine;   // one Newton-Schulz step on X^-1 (indefinite problems), run between the inversion and S
    bool symcand = false;             // structurally symmetric: square patches, no coarsening anywhere below
    bool colsplit = false;            // peer-mapped partition of S / T by block columns instead of rows (plan_batch_gemms)
    bool use_sym = false;             // decided at build time (leaf operator self-adjoint, EFGPU_NO_SYMMETRY not set)
    const std::vector<Step>& active() const { return use_sym ? steps_sym : steps; }
    DevBuf d_blocks, d_trans, d_entries, d_ptab;
    // TMA operand staging (gemm_tma.cu): the matrices the operand blocks of the 128-row products lie in, per-block coordinates, and
    // one pair of tensor maps per (parent, view), encoded when the buffers are allocated
    struct TmaView { int op; long long origin; int ld; };
    std::vector<TmaView> views;
    std::vector<TmaBlock> tblocks;
    DevBuf d_tblocks, d_mapsA, d_mapsB;
    DevBuf Xinv, S, Hc, T, Xcopy, Tcoarse;
    double* Tbase = nullptr;          // this batch's DtN slab: its own buffer T, or a slice of the handle's transient arena (EFGPU_LEAN_T)
    std::vector<std::vector<CoarsenOp>> cT, cH, cG;  // per step
    std::vector<std::unique_ptr<DevBuf>> d_cT, d_cH, d_cG;
    std::vector<int> cT_max, cH_max, cG_max;
    size_t ws_per_entry = 0, w2_off = 0, w3_off = 0;   // per-entry workspace: [W1 blocks per depth | W2 | W3 staging]
    std::vector<double*> h_ptab;   // host copy of the operand table (all-gather callbacks need host-side addresses)
    std::vector<MergeEntry> h_ent; // host copy of the merge entries
    std::vector<std::vector<int>> cT_slot;   // parent slot of every coarsening op of cT (to restrict them to a subset of the parents)
    // adaptive re-build (efgpu_rebuild_from): the build runs on the dirty parents only - compact copies of the tables above
    bool sub_on = false; int sub_count = 0;
    DevBuf sub_entries, sub_ptab;
    std::vector<std::unique_ptr<DevBuf>> sub_cT; std::vector<int> sub_cT_n, sub_cT_max;
};

}  // namespace efgpu

using namespace efgpu;

struct efgpu_handle {
    int device = 0, M = 0, n_nodes = 0, n_leaves = 0, max_level = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::vector<NodeH> nodes;
    std::vector<int> leaf_nodes;                 // node id of each leaf, pre-order (= Morton order)
    std::vector<int> roots;                      // nodes without a parent (one for a tree; several for a forest of subtrees)
    bool external_leaves = false;                // leaf T / h are supplied by the caller (upper tree of a sharded run)
    int part_rank = 0, part_nranks = 1;          // row partition of S / T among the ranks that replicate this tree
    bool root_T_distributed = false;             // partitioned tree after a build: the roots' T rows are not gathered / mirrored yet
    bool root_T_pending = false;                 // EFGPU_LAZY_ROOT_DTN: the level-0 DtN products have not been issued yet
    unsigned cur_flags = 0;                      // flags of the build in progress / of the last build
    efgpu_allgather_fn allgather = nullptr; void* allgather_user = nullptr;   // collective supplied by the caller (NCCL)
    // peer mode (peer.cu): the buffers other ranks write into (X^-1, S, T, workspaces, leaf maps, vectors) are slices of ONE
    // allocation with the same layout on every rank, mapped into every rank by CUDA IPC
    bool peer_mode = false, peer_attached = false;
    DevBuf arena; size_t arena_need = 0;
    PeerSpan peers;
    void* peer_mapped[PEER_MAX] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool peer_dirty = true;                      // local work since the last barrier: a barrier must precede the next peer stores
    bool peer_pending = false;                   // column-partitioned S was stored into the peers and no barrier has completed it yet
    DevBuf d_peer_err;
    std::vector<size_t> leafT_off;               // element offset of each leaf's T inside d_leafT
    // external leaves tagged by their parent receive coarsened Dirichlet data: uncoarsen it at the end of the solve
    // (in an unsharded run this is the first thing the leaf's own split1to4 would do, HPSAlgorithm.hpp:1165-1183)
    std::vector<std::vector<efgpu::CoarsenOp>> extG;
    std::vector<std::unique_ptr<efgpu::DevBuf>> d_extG;
    std::vector<int> extG_max;
    std::vector<std::vector<int>> level_batches; // batch ids per tree level
    std::vector<BatchH> batches;
    // leaf model
    int leaf_kind = EFGPU_LEAF_CONSTANT; double lambda = 0.0;
    bool ext_sym = false;                        // external leaves declared signed-symmetric (efgpu_set_symmetric_leaves)
    double lambda_max = 0.0;                     // variable-coefficient leaves: largest sampled lambda
    int refine_mode = -1;                        // efgpu_set_refine_inverse: -1 automatic (lambda > 0), 0 never, 1 always
    bool refine_inverse = false;                 // this build: one Newton-Schulz step on every X^-1 (decided in build_begin)
    // device state
    DevBuf d_leaf_build, d_leaf_src; int n_leaf_build = 0;   // constant-coefficient leaves: one DtN computation per class of identical (dx, dy)
    DevBuf d_Q, d_boxes, d_leaf_nodes, d_leafT, d_vec, d_ws, d_leaf_h, d_leaf_g, d_f, d_u, d_minpiv;
    DevBuf d_Tarena[2];                          // EFGPU_LEAN_T: DtN maps of even / odd tree levels (a level's maps die when its parents are merged)
    bool lean_T = false;
    DevBuf d_robin;                              // workspace of the root boundary solve (efgpu_solve_robin)
    DevBuf d_pts, d_err;                         // staging of efgpu_leaf_points / efgpu_error_norms (sample.cu)
    bool solve_done = false;                     // d_u holds the solution of a solve stage
    DevBuf d_coef_in[6], d_coef, d_P;            // variable-coefficient leaves: sampled alpha/beta/lambda, stencil coefficients, block-LU inverses
    size_t vec_doubles = 0;
    bool allocated = false, built = false, upwards_done = false;
    const double* f_cur = nullptr; double fscale_cur = 1.0;   // load of the last upwards call (re-used by the leaf solves)
    unsigned build_flags = 0;
    std::string last_error;
    efgpu_stats_t stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Batches of one tree level are independent: on adaptive trees (several child sizes per level) they run on parallel stream
    // lanes, forked from / joined into the handle's stream at every level; and the launch sequence of a whole stage is captured
    // into a CUDA graph the second time it is issued with the same key (flags, leaf model, layout generation), so that repeated
    // builds / solves replay it without per-launch host cost.
    static constexpr int MAX_LANES = 4;
    int n_lanes = 1; bool lanes_on = false;
    cudaStream_t lane[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr; unsigned long long key = ~0ull; int seen = 0;
        double launches[EFGPU_PROF_NCLASSES] = {0};
    };
    GraphSlot g_build, g_up, g_solve;
    unsigned long long graph_gen = 0;
    bool graphs_on = true;
    bool graphs_peer = true;           // EFGPU_GRAPHS_PEER=0: peer-mapped (row-partitioned) trees keep plain launches
    // optional per-kernel-class timing (efgpu_set_profiling): event pairs around every launch group
    bool profiling = false;
    struct ProfRec { int cls; cudaEvent_t a, b; std::string label; };
    FILE* trace = nullptr;             // EFGPU_TRACE=<prefix>: one line per timed launch group (class, label, ms) while profiling
    std::string cur_label;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[EFGPU_PROF_NCLASSES] = {0};
    double prof_launches[EFGPU_PROF_NCLASSES] = {0};
};

namespace efgpu {

static cudaEvent_t take_event(efgpu_handle* H)
{
    if (!H->ev_pool.empty()) { cudaEvent_t e = H->ev_pool.back(); H->ev_pool.pop_back(); return e; }
    cudaEvent_t e; EF_CUDA(cudaEventCreate(&e)); return e;
}
// Runs `f` (which launches `nlaunch` kernels of class `cls` on the handle's stream); in profiling
// mode the group is bracketed by an event pair that collect_profile() turns into milliseconds.
template <class F>
static inline void timed(efgpu_handle* H, int cls, int nlaunch, F&& f)
{
    H->prof_launches[cls] += nlaunch;
    if (!H->profiling) { f(); return; }
    cudaEvent_t a = take_event(H), b = take_event(H);
    EF_CUDA(cudaEventRecord(a, H->stream));
    f();
    EF_CUDA(cudaEventRecord(b, H->stream));
    H->prof_recs.push_back({cls, a, b, H->trace ? H->cur_label : std::string()});
}
static void collect_profile(efgpu_handle* H)   // stream must be synchronised
{
    for (auto& r : H->prof_recs) {
        float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        H->prof_ms[r.cls] += ms;
        if (H->trace) fprintf(H->trace, "%d %.4f %s\n", r.cls, ms, r.label.c_str());
        H->ev_pool.push_back(r.a); H->ev_pool.push_back(r.b);
    }
    H->prof_recs.clear();
    if (H->trace) { fprintf(H->trace, "# collected\n"); fflush(H->trace); }
}

// Stream lanes of one tree level: the first use of a lane in a level makes it wait for everything issued on the handle's stream
// so far; join() makes the handle's stream wait for every lane used.  Works the same inside a stream capture (the events become
// graph edges).  Off (everything on the handle's stream) while profiling: the per-class event pairs are recorded there.
struct LaneSet {
    efgpu_handle* H; unsigned used = 0; bool forked = false;
    explicit LaneSet(efgpu_handle* h) : H(h) {}
    cudaStream_t get(const BatchH& b) {
        if (!H->lanes_on || H->profiling || b.lane <= 0 || b.lane >= H->n_lanes) return H->stream;
        if (!forked) { EF_CUDA(cudaEventRecord(H->ev_fork, H->stream)); forked = true; }
        if (!(used & (1u << b.lane))) { EF_CUDA(cudaStreamWaitEvent(H->lane[b.lane], H->ev_fork, 0)); used |= 1u << b.lane; }
        return H->lane[b.lane];
    }
    void join() {
        for (int k = 1; k < efgpu_handle::MAX_LANES; k++)
            if (used & (1u << k)) {
                EF_CUDA(cudaEventRecord(H->ev_join[k], H->lane[k]));
                EF_CUDA(cudaStreamWaitEvent(H->stream, H->ev_join[k], 0));
            }
        used = 0; forked = false;
    }
};

static unsigned long long graph_key(const efgpu_handle* H, unsigned flags)
{
    unsigned long long k = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { k ^= v; k *= 1099511628211ull; };
    unsigned long long lam; std::memcpy(&lam, &H->lambda, 8);
    mix(flags); mix((unsigned long long)H->leaf_kind); mix(lam); mix(H->refine_inverse ? 1 : 0); mix(H->graph_gen);
    mix(H->ext_sym ? 1 : 0); mix((unsigned long long)(uintptr_t)H->stream);
    for (int t = 0; t < 16; t++) mix((unsigned long long)get_tuning(t));
    return k;
}

// Runs `body` (launches on the handle's stream and its lanes) eagerly the first time a key is seen - one-time function
// attributes and occupancy queries happen there -, captures it into a graph the second time, and replays the graph from then on.
template <class F>
static void run_graphed(efgpu_handle* H, efgpu_handle::GraphSlot& g, unsigned long long key, F&& body)
{
    // (a row-partitioned tree that exchanges through a host callback cannot be captured; a peer-mapped one can: its barriers count
    // their epochs on the device)
    const bool can = H->graphs_on && !H->profiling && !H->allgather && (H->part_nranks == 1 || (H->peer_mode && H->peer_attached && H->graphs_peer));
    if (!can) { body(); return; }
    if (g.exec && g.key == key) {
        EF_CUDA(cudaGraphLaunch(g.exec, H->stream));
        for (int c = 0; c < EFGPU_PROF_NCLASSES; c++) H->prof_launches[c] += g.launches[c];
        return;
    }
    if (g.key != key) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        g.key = key; g.seen = 0;
    }
    if (g.seen++ == 0) { body(); return; }
    double before[EFGPU_PROF_NCLASSES];
    for (int c = 0; c < EFGPU_PROF_NCLASSES; c++) before[c] = H->prof_launches[c];
    EF_CUDA(cudaStreamBeginCapture(H->stream, cudaStreamCaptureModeRelaxed));
    cudaGraph_t graph = nullptr;
    try { body(); }
    catch (...) { cudaStreamEndCapture(H->stream, &graph); if (graph) cudaGraphDestroy(graph); g.key = ~0ull; throw; }
    EF_CUDA(cudaStreamEndCapture(H->stream, &graph));
    cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { g.exec = nullptr; g.key = ~0ull; throw Error{EF_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)}; }
    for (int c = 0; c < EFGPU_PROF_NCLASSES; c++) { g.launches[c] = H->prof_launches[c] - before[c]; H->prof_launches[c] = before[c] + g.launches[c]; }
    EF_CUDA(cudaGraphLaunch(g.exec, H->stream));
}

// Planner of the blocked inversion of X (in place, op OP_XINV).  Two properties of the merge matrix are used:
//  * structure (always): in the interior ordering [alpha|gamma, beta|omega, alpha|beta, gamma|omega] the two diagonal
//    2n x 2n blocks of X are block diagonal (mergeX_, HPSAlgorithm.hpp:870-895: the zero blocks), so at the top level
//    A^-1 = diag(A11^-1, A22^-1) and the products with A^-1 have K = n instead of 2n;
//  * symmetry (plan `sym`): where every patch below is square, uncoarsened and the leaf operator self-adjoint, X is
//    symmetric; then C A^-1 = (A^-1 B)^T and the new C block is the transpose of the new B block: four GEMMs and two
//    transposes per recursion level instead of six GEMMs.
// The two diagonal blocks of A at the top level are independent and of the same shape: their recursions are planned
// separately (second one on its own workspace slots w1b / w2b) and zipped step by step into launches that carry both,
// which shortens the serial chain of small launches of the top tree levels by a quarter.
struct InvPlanner {
    BatchH& b; std::vector<Step>& steps; const std::vector<long long>& w1_off; int ld, rank, nranks; bool sym;
    bool peer = false;                          // peer mode: split products store straight into every rank's arena (gk 3), no staging
    std::vector<GemmBlock>* blk = nullptr;      // where descriptors go (default: the batch's vectors)
    std::vector<TransOp>* trn = nullptr;
    const std::vector<long long>* w1b = nullptr; // workspace of a zipped second branch (null: no zipping below this planner)
    long long w2 = 0, w2b = 0;                   // W2 base of this planner / of a zipped second branch
    std::vector<GemmBlock>& B_() { return blk ? *blk : b.blocks; }
    std::vector<TransOp>& T_() { return trn ? *trn : b.trans; }

    void small(long long off, int N) {
        Step st{}; st.kind = 0; st.off = off; st.N = N; st.cls = EFGPU_PROF_INVERT_SMALL;
        steps.push_back(st);
    }
    // whether a product with this many result rows is row-split over the ranks.  A product is split only where the flops saved
    // outweigh the exchange that follows (NCCL callback: tens of microseconds per collective, so 2048 rows and more; peer mode: two
    // flag barriers of a few microseconds, so 1024 rows and more); the deep, small products of the recursion are recomputed by
    // every rank (6 % of the inversion flops).  Slices are whole 128-row (peer mode: 64-row) tiles.
    bool will_split(int rows) const {
        const int split_min_env = [] { const char* e = getenv("EFGPU_SPLIT_MIN_ROWS"); return e ? atoi(e) : 0; }();   // read per plan
        const int split_min = split_min_env > 0 ? split_min_env : (peer ? 1024 : 2048);
        return nranks > 1 && rows % ((peer ? 64 : 128) * nranks) == 0 && rows >= split_min;
    }
    // C (rows x cols) = [C0] +- A (rows x K) B (K x cols); split by rows over the ranks when each slice keeps full 128-row tiles
    // ct_op >= 0: the result's transpose is stored as well (cols x rows block at ct_off, leading dimension ldct)
    void gemm(int rows, int cols, int K, int c_op, long long c_off, int ldc, int c0_op, long long c0_off, int ldc0,
              int a_op, long long a_off, int lda, int b_op, long long b_off, int ldb, bool neg,
              int ct_op = -1, long long ct_off = 0, int ldct = 0) {
        GemmBlock g{};
        g.c_op = c_op; g.c_off = c_off; g.ldc = ldc; g.c0_op = c0_op; g.c0_off = c0_off; g.ldc0 = ldc0;
        g.rows = rows; g.cols = cols; g.nterms = 1;
        if (ct_op >= 0) { g.ct_op1 = ct_op + 1; g.ct_off = ct_off; g.ldct = ldct; }
        g.t[0] = GemmTerm{a_op, b_op, lda, ldb, a_off, b_off, K, neg ? 0x80000000u : 0u};
        Step st{}; st.kind = 1; st.first = (int)B_().size(); st.count = 1; st.cls = EFGPU_PROF_GEMM_XINV;
        if (will_split(rows)) {
            const long long skip = (long long)rank * (rows / nranks);
            st.g_op = c_op; st.g_off = c_off; st.g_rows = rows; st.g_cols = cols; st.g_ld = ldc;
            if (peer) st.gk = 3;             // the epilogue stores the tile on every rank: any leading dimension, no staging
            else if (ldc == cols) st.gk = 1;
            else { st.gk = 2; g.c_op = OP_W3; g.c_off = 0; g.ldc = cols; }   // rows land in the contiguous staging block
            g.c_off += skip * g.ldc;
            if (g.c0_op >= 0) g.c0_off += skip * g.ldc0;
            g.t[0].a_off += skip * lda;
            g.ct_off += skip;      // rows [skip, skip + rows / nranks) of the result = those columns of its transpose
            g.rows = rows / nranks;
        }
        steps.push_back(st);
        B_().push_back(g);
    }
    // whether transposes can ride on the producing GEMM's epilogue: not when row slices are exchanged by a caller-supplied
    // all-gather (round 1's path), which only knows the primary destination
    bool fused_transposes() const { return nranks == 1 || peer; }
    void transpose(int rows, int cols, int s_op, long long s_off, int lds, int d_op, long long d_off, int ldd) {
        Step st{}; st.kind = 2; st.first = (int)T_().size(); st.count = 1; st.cls = EFGPU_PROF_TRANSPOSE;
        T_().push_back(TransOp{s_op, d_op, lds, ldd, s_off, d_off, rows, cols, 0u, 0});
        steps.push_back(st);
    }
    // C = C0 - A B with a SYMMETRIC h x h result (Schur complement and the update of the leading block of a symmetric X):
    // only the upper triangle of a 2 x 2 or 4 x 4 block partition is multiplied (3 of 4 / 10 of 16 sub-blocks, one
    // launch), the lower blocks are transposes.  Row-partitioned products keep the plain form (their slices are gathered).
    void gemm_sym(int h, int K, int c_op, long long c_off, int ldc, int a_op, long long a_off, int lda, int b_op, long long b_off, int ldb) {
        const bool row_split = will_split(h);
        const int nb = row_split ? 1 : ((h % 64 == 0 && h / 4 >= 128) ? 4 : ((h % 32 == 0 && h / 2 >= 128) ? 2 : 1));
        if (nb == 1) { gemm(h, h, K, c_op, c_off, ldc, c_op, c_off, ldc, a_op, a_off, lda, b_op, b_off, ldb, true); return; }
        const int sb = h / nb;
        Step st{}; st.kind = 1; st.first = (int)B_().size(); st.cls = EFGPU_PROF_GEMM_XINV;
        Step tr{}; tr.kind = 2; tr.first = (int)T_().size(); tr.cls = EFGPU_PROF_TRANSPOSE;
        for (int I = 0; I < nb; I++)
            for (int J = I; J < nb; J++) {
                GemmBlock g{};
                const long long cij = c_off + (long long)I * sb * ldc + (long long)J * sb;
                g.c_op = c_op; g.c_off = cij; g.ldc = ldc; g.c0_op = c_op; g.c0_off = cij; g.ldc0 = ldc;
                g.rows = sb; g.cols = sb; g.nterms = 1;
                g.t[0] = GemmTerm{a_op, b_op, lda, ldb, a_off + (long long)I * sb * lda, b_off + (long long)J * sb, K, 0x80000000u};
                if (J > I) {   // the mirrored block: by the same epilogue, or (caller-supplied all-gather) by a transpose step
                    const long long cji = c_off + (long long)J * sb * ldc + (long long)I * sb;
                    if (fused_transposes()) { g.ct_op1 = c_op + 1; g.ct_off = cji; g.ldct = ldc; }
                    else { T_().push_back(TransOp{c_op, c_op, ldc, ldc, cij, cji, sb, sb, 0u, 0}); tr.count++; }
                }
                B_().push_back(g); st.count++;
            }
        steps.push_back(st);
        if (tr.count) steps.push_back(tr);
    }
    // appends the steps of a sub-plan (descriptor indices rebased)
    void append(const std::vector<Step>& ss, const std::vector<GemmBlock>& bb, const std::vector<TransOp>& tt) {
        for (Step st : ss) {
            if (st.kind == 1) { const int f = (int)B_().size(); for (int k = 0; k < st.count; k++) B_().push_back(bb[st.first + k]); st.first = f; }
            if (st.kind == 2) { const int f = (int)T_().size(); for (int k = 0; k < st.count; k++) T_().push_back(tt[st.first + k]); st.first = f; }
            steps.push_back(st);
        }
    }
    // inverts the two q x q blocks at offA and offB: zipped into common launches when a second workspace is available
    void invert_pair(long long offA, long long offB, int q, int depth) {
        if (!w1b) { invert(offA, q, depth, false); invert(offB, q, depth, false); return; }
        std::vector<Step> sA, sB; std::vector<GemmBlock> bA, bB; std::vector<TransOp> tA, tB;
        InvPlanner pa{b, sA, w1_off, ld, rank, nranks, sym}; pa.blk = &bA; pa.trn = &tA; pa.w2 = w2; pa.peer = peer;
        InvPlanner pb{b, sB, *w1b, ld, rank, nranks, sym}; pb.blk = &bB; pb.trn = &tB; pb.w2 = w2b; pb.peer = peer;
        pa.invert(offA, q, depth, false);
        pb.invert(offB, q, depth, false);
        bool zip = sA.size() == sB.size();
        for (size_t i = 0; zip && i < sA.size(); i++)
            zip = sA[i].kind == sB[i].kind && sA[i].count == sB[i].count && sA[i].gk == sB[i].gk && (sA[i].gk == 0 || sA[i].gk == 3) && sA[i].N == sB[i].N;
        if (!zip) { append(sA, bA, tA); append(sB, bB, tB); return; }
        for (size_t i = 0; i < sA.size(); i++) {
            Step st = sA[i];
            if (st.kind == 0) st.off2 = sB[i].off;
            else if (st.kind == 1) {
                st.first = (int)B_().size(); st.count = 2 * sA[i].count;
                for (int k = 0; k < sA[i].count; k++) B_().push_back(bA[sA[i].first + k]);
                for (int k = 0; k < sB[i].count; k++) B_().push_back(bB[sB[i].first + k]);
            } else {
                st.first = (int)T_().size(); st.count = 2 * sA[i].count;
                for (int k = 0; k < sA[i].count; k++) T_().push_back(tA[sA[i].first + k]);
                for (int k = 0; k < sB[i].count; k++) T_().push_back(tB[sB[i].first + k]);
            }
            steps.push_back(st);
        }
    }
    // inverts the N x N block at `off`; depth = log2(size of X / N) selects the W1 slot; `diag2`: the leading and the
    // trailing half are themselves block diagonal with blocks of N/4 (only the top level of X)
    void invert(long long off, int N, int depth, bool diag2) {
        // base case: one CTA, register-resident Gauss-Jordan (N <= 128) - or, where only a few blocks are inverted at a time (the
        // serial chain of the top tree levels: batches of at most four merges), a whole 256 x 256 block by a cluster of eight CTAs
        // (invert_cluster_kernel; tuning key 9, read when the plan is made)
        const bool cluster256 = N == 256 && b.count <= 4 && get_tuning(9) == 1;
        const bool can_split = N > 128 && (N / 2) % 16 == 0 && !cluster256;
        if (!can_split) {
            if (cluster256) { small(off, N); return; }
            if (N > 128) throw Error{EF_ERR_BAD_SHAPE, "merge matrix cannot be blocked (patch size must be 8*2^k, 16*2^k, 24*2^k ...)"};
            small(off, N);
            return;
        }
        const int h = N / 2, q = h / 2;
        const long long A = off, B = off + h, C = off + (long long)h * ld, D = off + (long long)h * ld + h;
        const long long W1 = w1_off[depth];
        const bool dg = diag2 && q % 8 == 0;
        if (dg) invert_pair(A, A + (long long)q * ld + q, q, depth + 2);                                      // A <- diag(A11^-1, A22^-1)
        else invert(A, h, depth + 1, false);                                                                  // A <- A^-1
        if (sym) {
            if (dg) {                                                                                     // W1 = A^-1 B
                gemm(q, h, q, OP_W1, W1, h, -1, 0, 0, OP_XINV, A, ld, OP_XINV, B, ld, false);
                gemm(q, h, q, OP_W1, W1 + (long long)q * h, h, -1, 0, 0, OP_XINV, A + (long long)q * ld + q, ld, OP_XINV, B + (long long)q * ld, ld, false);
            } else gemm(h, h, h, OP_W1, W1, h, -1, 0, 0, OP_XINV, A, ld, OP_XINV, B, ld, false);
            gemm_sym(h, h, OP_XINV, D, ld, OP_XINV, C, ld, OP_W1, W1, h);                                 // D <- D - C W1   (Schur complement, symmetric)
            invert(D, h, depth + 1, false);                                                               // D <- S^-1
            if (fused_transposes()) {
                // B <- -W1 S^-1 and, from the same epilogue, C <- B^T.  The update of the leading block needs no W1^T either:
                // B W1^T is symmetric (= -W1 S^-1 W1^T), hence equal to its transpose W1 B^T = W1 C.
                // Row-split products of a peer-mapped tree keep the transpose out of the epilogue: a rank's slice of B is a COLUMN
                // slab of C, and storing those into seven peer arenas ran at 75-160 GB/s (r2m trace: 2.46 ms against 0.89 ms for
                // the same product without it); every rank transposes the gathered B locally instead (50 us at the root).
                if (peer && will_split(h)) {
                    gemm(h, h, h, OP_XINV, B, ld, -1, 0, 0, OP_W1, W1, h, OP_XINV, D, ld, true);
                    transpose(h, h, OP_XINV, B, ld, OP_XINV, C, ld);
                } else
                    gemm(h, h, h, OP_XINV, B, ld, -1, 0, 0, OP_W1, W1, h, OP_XINV, D, ld, true, OP_XINV, C, ld);
                gemm_sym(h, h, OP_XINV, A, ld, OP_W1, W1, h, OP_XINV, C, ld);                              // A <- A^-1 - W1 C (symmetric)
                return;
            }
            gemm(h, h, h, OP_XINV, B, ld, -1, 0, 0, OP_W1, W1, h, OP_XINV, D, ld, true);                  // B <- -W1 S^-1
            transpose(h, h, OP_XINV, B, ld, OP_XINV, C, ld);                                              // C <- B^T
            transpose(h, h, OP_W1, W1, h, OP_W2, w2, h);                                                   // W2 = W1^T  (= C A^-1)
            gemm_sym(h, h, OP_XINV, A, ld, OP_XINV, B, ld, OP_W2, w2, h);                                  // A <- A^-1 - B W2 (symmetric)
            return;
        }
        if (dg) {                                                                                         // W1 = C A^-1
            gemm(h, q, q, OP_W1, W1, h, -1, 0, 0, OP_XINV, C, ld, OP_XINV, A, ld, false);
            gemm(h, q, q, OP_W1, W1 + q, h, -1, 0, 0, OP_XINV, C + q, ld, OP_XINV, A + (long long)q * ld + q, ld, false);
        } else gemm(h, h, h, OP_W1, W1, h, -1, 0, 0, OP_XINV, C, ld, OP_XINV, A, ld, false);
        gemm(h, h, h, OP_XINV, D, ld, OP_XINV, D, ld, OP_W1, W1, h, OP_XINV, B, ld, true);                // D <- D - W1 B   (Schur complement)
        invert(D, h, depth + 1, false);                                                                   // D <- S^-1
        if (dg) {                                                                                         // W2 = A^-1 B
            gemm(q, h, q, OP_W2, w2, h, -1, 0, 0, OP_XINV, A, ld, OP_XINV, B, ld, false);
            gemm(q, h, q, OP_W2, w2 + (long long)q * h, h, -1, 0, 0, OP_XINV, A + (long long)q * ld + q, ld, OP_XINV, B + (long long)q * ld, ld, false);
        } else gemm(h, h, h, OP_W2, w2, h, -1, 0, 0, OP_XINV, A, ld, OP_XINV, B, ld, false);
        gemm(h, h, h, OP_XINV, B, ld, -1, 0, 0, OP_W2, w2, h, OP_XINV, D, ld, true);                       // B <- -W2 S^-1
        gemm(h, h, h, OP_XINV, C, ld, -1, 0, 0, OP_XINV, D, ld, OP_W1, W1, h, true);                      // C <- -S^-1 W1
        gemm(h, h, h, OP_XINV, A, ld, OP_XINV, A, ld, OP_XINV, B, ld, OP_W1, W1, h, true);                // A <- A^-1 - B W1
    }
};

// Row partition of a replicated upper tree (efgpu_set_partition): this rank computes rows [lo, hi) of the
// result; a block covering result rows [r0, r0 + rows) is clipped to the overlap (A-side offsets follow).
static bool clip_rows(GemmBlock& g, long long r0, long long lo, long long hi)
{
    const long long a = std::max(r0, lo), e = std::min(r0 + g.rows, hi);
    if (e <= a) return false;
    const long long skip = a - r0;
    g.c_off += skip * g.ldc;
    if (g.c0_op >= 0) g.c0_off += skip * g.ldc0;
    for (int t = 0; t < g.nterms; t++) g.t[t].a_off += skip * g.t[t].lda;
    g.rows = (int)(e - a);
    return true;
}

static void plan_refine(BatchH& b);

// Tensor-map views of the operand blocks (TMA staging of the 128 x 64 CTA tiles).  A block qualifies when its shape is made of whole
// tiles (rows % 128, cols % 64, K % 16) and each operand block lies inside one row-major matrix: every operand slot is one matrix
// starting at the slot's base (its offsets decompose as row * ld + col without wrapping) except OP_W1, whose slots per recursion depth
// start at `w1_starts`.  Steps whose blocks all qualify are marked; everything else keeps the cp.async kernel.
static void plan_tma(BatchH& b, const std::vector<long long>& w1_starts)
{
    b.views.clear();
    TmaBlock none{};
    for (auto& t : none.t) t.a_view = t.b_view = -1;
    b.tblocks.assign(b.blocks.size(), none);
    auto view_of = [&](int op, long long off, int ld, int width, int& row, int& col) -> int {
        long long origin = 0;
        if (op == OP_W1) for (long long st : w1_starts) if (st <= off && st > origin) origin = st;
        const long long rel = off - origin;
        if (ld < 16 || (ld & 1) || (origin & 1) || rel / ld >= (1ll << 30)) return -1;
        row = (int)(rel / ld); col = (int)(rel % ld);
        if (col + width > ld) return -1;
        for (size_t v = 0; v < b.views.size(); v++)
            if (b.views[v].op == op && b.views[v].origin == origin && b.views[v].ld == ld) return (int)v;
        b.views.push_back(BatchH::TmaView{op, origin, ld});
        return (int)b.views.size() - 1;
    };
    for (size_t k = 0; k < b.blocks.size(); k++) {
        const GemmBlock& g = b.blocks[k];
        bool ok = g.rows > 0 && g.rows % 128 == 0 && g.cols % 64 == 0;
        for (int t = 0; t < g.nterms; t++) ok = ok && g.t[t].K > 0 && g.t[t].K % 16 == 0;
        if (!ok) continue;
        TmaBlock tb = none;
        for (int t = 0; t < g.nterms && ok; t++) {
            tb.t[t].a_view = view_of(g.t[t].a_op, g.t[t].a_off, g.t[t].lda, g.t[t].K, tb.t[t].a_row, tb.t[t].a_col);
            tb.t[t].b_view = view_of(g.t[t].b_op, g.t[t].b_off, g.t[t].ldb, g.cols, tb.t[t].b_row, tb.t[t].b_col);
            ok = tb.t[t].a_view >= 0 && tb.t[t].b_view >= 0;
        }
        if (!ok) continue;
        if (g.nterms == 1) tb.t[1] = tb.t[0];
        b.tblocks[k] = tb;
    }
    for (std::vector<Step>* ss : {&b.steps, &b.steps_sym, &b.steps_refine})
        for (Step& st : *ss) {
            st.tma = st.kind == 1 && st.count > 0;
            for (int k = st.first; st.tma && k < st.first + st.count; k++) st.tma = b.tblocks[k].t[0].a_view >= 0;
        }
}

static void plan_batch_gemms(BatchH& b, int rank, int nranks, bool peer = false)
{
    const int n = b.n, N = 4 * n;
    b.blocks.clear(); b.trans.clear(); b.steps.clear(); b.steps_sym.clear(); b.steps_refine.clear();
    if (nranks > 1 && ((4 * n) % (8 * nranks) != 0))
        throw Error{EF_ERR_BAD_SHAPE, "row partition: 4 n must be a multiple of 8 * nranks for every merge of the replicated tree"};
    const long long s_lo = (long long)rank * (4 * n) / nranks, s_hi = (long long)(rank + 1) * (4 * n) / nranks;
    const long long t_lo = (long long)rank * (8 * n) / nranks, t_hi = (long long)(rank + 1) * (8 * n) / nranks;
    // workspace layout per entry: W1 blocks per recursion depth (op 7), W2 (op 8), W3 staging (op 9)
    std::vector<long long> w1_off; long long acc = 0;
    for (int h = N / 2; h >= 8; h /= 2) { w1_off.push_back(acc); acc += (long long)h * h; }
    w1_off.push_back(acc); w1_off.push_back(acc); w1_off.push_back(acc);
    // second set of W1 slots for depths >= 2 (the zipped sub-inversion of the second diagonal block of A); its W2 use
    // ((N/8)^2 doubles) sits in the upper half of W2
    std::vector<long long> w1b_off(w1_off.size(), acc);
    for (size_t d = 2; d < w1_off.size(); d++) w1b_off[d] = acc + (w1_off[d] - w1_off[2]);
    acc += acc - w1_off[2];
    const long long w1_total = acc, w2_total = (long long)(N / 2) * (N / 2);
    b.w2_off = (size_t)w1_total; b.w3_off = (size_t)(w1_total + w2_total);
    b.ws_per_entry = (size_t)(w1_total + 2 * w2_total);
    // (the root's DtN map of a partitioned tree stays row-distributed: its mirrored blocks are completed on demand,
    // efgpu_complete_root_dtn, so that the root merge also issues 36 instead of 64 block products)
    const bool mirror_ok = true;
    // partitions with at most one block row of T per rank (8 ranks and more): the four opposite pairs are shared half and half,
    // 4.5 instead of 5 : 4 block products per row (halves of >= 128 rows / columns: full GEMM tiles)
    const bool split_opposite = nranks >= 8 && n % 256 == 0;
    // peer-mapped trees whose rank count divides the eight WESN block columns: S and T are partitioned by COLUMNS (see below)
    const bool colsplit = peer && nranks > 1 && 8 % nranks == 0 && get_tuning(11) == 1;
    b.colsplit = colsplit;
    for (int variant = 0; variant < (b.symcand ? 2 : 1); variant++) {
        const bool sym = variant == 1;
        std::vector<Step>& steps = sym ? b.steps_sym : b.steps;
        InvPlanner ip{b, steps, w1_off, N, rank, nranks, sym};
        ip.w1b = &w1b_off; ip.w2b = w2_total / 2; ip.peer = peer;
        ip.invert(0, N, 0, true);
        // S = X^-1 S_RHS, written with WESN-permuted columns (mergeS_ + reorderOperators_)
        int first = (int)b.blocks.size();
        for (int k = 0; k < 4; k++)
            for (int q = 0; q < 8; q++) {
                const int c = q >> 1, side = h_tau_side[c][q & 1];
                GemmBlock g{};
                g.c_op = OP_S; g.c_off = (long long)(k * n) * (8 * n) + h_pos[q] * n; g.ldc = 8 * n; g.c0_op = -1;
                g.rows = n; g.cols = n; g.nterms = 2;
                for (int t = 0; t < 2; t++) {
                    const int k2 = h_kk[c][t];
                    g.t[t] = GemmTerm{OP_XINV, OP_TC0 + c, N, N, (long long)(k * n) * N + k2 * n,
                                      (long long)(h_iface[c][k2] * n) * N + side * n, n, h_sgn[c][k2] < 0 ? 0x80000000u : 0u};
                }
                if (colsplit) {   // column partition: the blocks of this rank's WESN positions, all four block rows
                    if ((long long)h_pos[q] * n >= t_lo && (long long)h_pos[q] * n < t_hi) b.blocks.push_back(g);
                } else if (clip_rows(g, (long long)k * n, s_lo, s_hi)) b.blocks.push_back(g);
            }
        { Step st{}; st.kind = 1; st.first = first; st.count = (int)b.blocks.size() - first; st.cls = EFGPU_PROF_GEMM_S; steps.push_back(st); }
        // T = T_LHS + H S, rows and columns in WESN order (mergeT_ + reorderOperators_).  Symmetric plan: with the sign
        // d = -1 on the W and S sides (coordinate derivatives instead of outward normals, FiniteVolumeSolver.cpp:332-343)
        // diag(d) T is symmetric, so of every off-diagonal pair of n x n blocks only one is computed - chosen on a circulant
        // pattern so that each block row carries 4 or 5 products and every half / quarter of the rows the same number (row
        // partitions over 2 and 4 ranks are balanced exactly, over 8 ranks to 5 : 4) - and the other is its signed transpose.
        // Column partition (peer-mapped trees): the same selection with rows and columns exchanged - the owner of position O
        // computes block (X, O) where the row partition computes (O, X) - so that every product reads S only in the columns its
        // own rank computed: S needs no exchange before T.
        first = (int)b.blocks.size();
        const int tfirst = (int)b.trans.size();
        // block of T at WESN block row h_pos[rowq], block column h_pos[colq], sub-range rows r0 .. r0 + nr, columns c0 .. c0 + nc
        auto make_T = [&](int rowq, int colq, int r0, int c0, int nr_, int nc_) {
            const int c = rowq >> 1, side_r = h_tau_side[c][rowq & 1];
            const int c2 = colq >> 1, side_c = h_tau_side[c2][colq & 1];
            const int PR = h_pos[rowq], PC = h_pos[colq];
            GemmBlock g{};
            g.c_op = OP_T; g.c_off = (long long)(PR * n + r0) * (8 * n) + PC * n + c0; g.ldc = 8 * n;
            if (c == c2) { g.c0_op = OP_TC0 + c; g.c0_off = (long long)(side_r * n + r0) * N + side_c * n + c0; g.ldc0 = N; }
            else g.c0_op = -1;
            g.rows = nr_; g.cols = nc_; g.nterms = 2;
            for (int t = 0; t < 2; t++) {
                const int k = h_kk[c][t];
                g.t[t] = GemmTerm{OP_TC0 + c, OP_S, N, 8 * n, (long long)(side_r * n + r0) * N + h_iface[c][k] * n,
                                  (long long)(k * n) * (8 * n) + PC * n + c0, n, 0u};
            }
            return g;
        };
        // the block computed by the owner `qo` against `qx`: at (qo, qx) in a row partition, transposed at (qx, qo) in a column partition
        auto emit_T = [&](int qo, int qx, int r0, int c0, int nr_, int nc_) {
            if (colsplit) {
                GemmBlock g = make_T(qx, qo, c0, r0, nc_, nr_);
                if ((long long)h_pos[qo] * n >= t_lo && (long long)h_pos[qo] * n < t_hi) b.blocks.push_back(g);
                return g;
            }
            GemmBlock g = make_T(qo, qx, r0, c0, nr_, nc_);
            if (clip_rows(g, (long long)h_pos[qo] * n + r0, t_lo, t_hi)) b.blocks.push_back(g);
            return g;
        };
        // mirror step: the rows x cols block at (position SP, offset sr0; position SQ, offset sc0) goes, transposed (and signed), to
        // (SQ, sc0; SP, sr0); a column partition exchanges the roles of source and destination coordinates
        auto add_mirror = [&](int SP, int sr0, int SQ, int sc0, int rows, int cols, unsigned neg) {
            if (colsplit) { std::swap(SP, SQ); std::swap(sr0, sc0); std::swap(rows, cols); }
            b.trans.push_back(TransOp{OP_T, OP_T, 8 * n, 8 * n, (long long)(SP * n + sr0) * (8 * n) + SQ * n + sc0,
                                      (long long)(SQ * n + sc0) * (8 * n) + SP * n + sr0, rows, cols, neg, 0});
        };
        for (int qr = 0; qr < 8; qr++)
            for (int qc = 0; qc < 8; qc++) {
                const int P = h_pos[qr], Q = h_pos[qc];
                if (sym && mirror_ok && P != Q) {
                    const int dl = (Q - P) & 7;
                    if (dl == 4 && split_opposite) {
                        // one block row per rank: both rows of an opposite pair compute HALF of their block - rows 0..3 the upper
                        // n/2 rows of (P, P + 4), rows 4..7 the right n/2 columns of (P, P - 4) - and receive the other half as
                        // the transpose of what the partner computed (no sign: both sides lie on the same kind of axis)
                        const int hn = n / 2;
                        const int r0 = 0, c0 = P < 4 ? 0 : hn, nr_ = P < 4 ? hn : n, nc_ = P < 4 ? n : hn;
                        emit_T(qr, qc, r0, c0, nr_, nc_);
                        if (P < 4) add_mirror(Q, 0, P, hn, n, hn, 0u);     // lower half of (P, Q) <- transpose of the right half of (Q, P)
                        else add_mirror(Q, 0, P, 0, hn, n, 0u);           // left half of (P, Q) <- transpose of the upper half of (Q, P)
                        continue;
                    }
                    // (of the opposite pair P, P + 4 the even one of rows 0..3 / the odd one of rows 4..7 computes: block rows 0, 2, 5, 7
                    // carry 5 products and 1, 3, 4, 6 carry 4, so halves and quarters of the rows - 2 and 4 ranks - get 18 and 9 each)
                    if (!(dl < 4 || (dl == 4 && ((P < 4) != ((P & 1) != 0))))) {
                        const unsigned neg = ((P ^ Q) & 2) ? 0x80000000u : 0u;   // W, W, E, E, S, S, N, N: d = -1 where bit 1 is clear
                        // unpartitioned: written by the epilogue of the product that computes (Q, P) (below); partitioned: a
                        // transpose step after the slices have been gathered
                        if (nranks > 1) add_mirror(Q, 0, P, 0, n, n, neg);
                        continue;
                    }
                }
                // a diagonal block of the signed-symmetric T is itself symmetric: only the upper triangle of its 2 x 2 / 4 x 4
                // sub-blocks is multiplied (efgpu_set_tuning key 5, default on: 202.8 -> 197.8 ms per step at L = 8, M = 16)
                const int nb = (sym && mirror_ok && P == Q && get_tuning(5) == 1)
                                   ? ((n % 64 == 0 && n / 4 >= 128) ? 4 : ((n % 32 == 0 && n / 2 >= 128) ? 2 : 1)) : 1;
                const int sb = n / nb;
                for (int I = 0; I < nb; I++)
                    for (int J = I; J < nb; J++) {
                        if (nranks > 1) {
                            emit_T(qr, qc, I * sb, J * sb, sb, sb);
                            if (J > I) add_mirror(P, I * sb, Q, J * sb, sb, sb, 0u);
                            continue;
                        }
                        GemmBlock g = make_T(qr, qc, I * sb, J * sb, sb, sb);
                        // mirrored partner written by the same epilogue (unpartitioned trees): the signed transpose block (Q, P) of an
                        // off-diagonal block of the symmetric plan, the lower sub-block of a diagonal block's triangle
                        if (sym && mirror_ok) {
                            if (P != Q) {
                                g.ct_op1 = OP_T + 1; g.ct_off = (long long)(Q * n) * (8 * n) + P * n; g.ldct = 8 * n;
                                g.ct_neg = ((P ^ Q) & 2) ? 0x80000000u : 0u;
                            } else if (J > I) {
                                g.ct_op1 = OP_T + 1; g.ct_off = (long long)(P * n + J * sb) * (8 * n) + Q * n + I * sb; g.ldct = 8 * n;
                            }
                        }
                        b.blocks.push_back(g);
                    }
            }
        { Step st{}; st.kind = 1; st.first = first; st.count = (int)b.blocks.size() - first; st.cls = EFGPU_PROF_GEMM_T; steps.push_back(st); }
        if ((int)b.trans.size() > tfirst) {   // runs after the row slices of T have been gathered
            Step st{}; st.kind = 2; st.first = tfirst; st.count = (int)b.trans.size() - tfirst; st.cls = EFGPU_PROF_MIRROR_T; steps.push_back(st);
        }
    }
    plan_refine(b);
    std::vector<long long> w1_starts(w1_off);
    w1_starts.insert(w1_starts.end(), w1b_off.begin(), w1b_off.end());
    plan_tma(b, w1_starts);
}

// One Newton-Schulz step  X^-1 <- X^-1 + X^-1 (I - X X^-1)  on the result of the unpivoted block inversion (see build_begin:
// indefinite problems only).  X is the copy kept in OP_XCOPY; the two N x N temporaries live in this merge's own DtN slot
// OP_T (64 n^2 doubles, not yet written when the step runs): E at offset 0, X^-1 E at offset N^2.  Replicated on every rank
// of a row partition.
static void plan_refine(BatchH& b)
{
    const int N = 4 * b.n;
    const long long NN = (long long)N * N;
    auto gemm = [&](long long c_off, int a_op, long long a_off, int b_op, long long b_off, bool neg) {
        GemmBlock g{};
        g.c_op = OP_T; g.c_off = c_off; g.ldc = N; g.c0_op = -1; g.rows = N; g.cols = N; g.nterms = 1;
        g.t[0] = GemmTerm{a_op, b_op, N, N, a_off, b_off, N, neg ? 0x80000000u : 0u};
        Step st{}; st.kind = 1; st.first = (int)b.blocks.size(); st.count = 1; st.cls = EFGPU_PROF_GEMM_XINV;
        b.blocks.push_back(g); b.steps_refine.push_back(st);
    };
    gemm(0, OP_XCOPY, 0, OP_XINV, 0, true);                                    // E = -X X^-1
    { Step st{}; st.kind = 3; st.off = 0; st.N = N; st.cls = EFGPU_PROF_TRANSPOSE; b.steps_refine.push_back(st); }   // E += I
    gemm(NN, OP_XINV, 0, OP_T, 0, false);                                      // D = X^-1 E
    { Step st{}; st.kind = 4; st.off = NN; st.N = N; st.cls = EFGPU_PROF_TRANSPOSE; b.steps_refine.push_back(st); }  // X^-1 += D
}

static void compute_flop_model(efgpu_handle* H)
{
    const int M = H->M, nn = H->n_nodes;
    // flop model: canonical dgesv+dgemm count and the count actually issued
    double canon = 0, issued = 0, up_bytes = 0, so_bytes = 0;
    for (auto& b : H->batches) {
        const double n3 = (double)b.n * b.n * b.n;
        canon += b.count * 810.0 * n3 + b.count * (2.0 / 3.0) * n3;
        for (const Step& st : b.active())
            if (st.kind == 1 && !(st.cls == EFGPU_PROF_GEMM_T && b.level == 0 && (H->cur_flags & EFGPU_LAZY_ROOT_DTN)))
                for (int k = st.first; k < st.first + st.count; k++)
                    for (int t = 0; t < b.blocks[k].nterms; t++) issued += b.count * 2.0 * b.blocks[k].rows * b.blocks[k].cols * b.blocks[k].t[t].K;
        if (H->refine_inverse) issued += b.count * 2.0 * 2.0 * 64.0 * n3;   // two (4n)^3 products of the Newton-Schulz step
        up_bytes += b.count * 8.0 * (16.0 + 16.0) * b.n * b.n;
        so_bytes += b.count * 8.0 * 32.0 * b.n * b.n;
    }
    H->stats.merge_flops_canonical = canon;
    H->stats.merge_flops_issued = issued;
    const double leaf_cells = H->external_leaves ? 0.0 : (double)H->n_leaves * M * M;
    // leaves: f (8 B / DOF) in, h out (upwards); f, g in, u out (solve).  Variable-coefficient leaves also stream the stored inverse
    // diagonal blocks P_i of their block-tridiagonal factorisation twice per right-hand side (forward and backward sweep):
    // 2 M doubles per DOF
    const double leaf_extra = (!H->external_leaves && H->leaf_kind == EFGPU_LEAF_VARIABLE) ? 16.0 * M * leaf_cells : 0.0;
    H->stats.upwards_bytes = up_bytes + 8.0 * leaf_cells + leaf_extra;
    H->stats.solve_bytes = so_bytes + 16.0 * leaf_cells + leaf_extra;
    H->stats.n_leaves = H->n_leaves;
    H->stats.n_nodes = nn;
    H->stats.dofs = leaf_cells;
}

static void make_plan(efgpu_handle* H, const efgpu_tree_desc* d, const int32_t* ext_leaf_size)
{
    const int nn = d->n_nodes, M = d->nx;
    if (nn <= 0 || !d->level || !d->child || !d->box) throw Error{EF_ERR_BAD_ARG, "empty tree description"};
    if (M != 4 && (M < 8 || M % 8)) throw Error{EF_ERR_BAD_SHAPE, "nx must be 4 or a multiple of 8 (4, 8, 16, 24, 32, 64)"};
    H->external_leaves = ext_leaf_size != nullptr;
    H->M = M; H->n_nodes = nn;
    H->nodes.resize(nn);
    for (int i = 0; i < nn; i++) {
        NodeH& nd = H->nodes[i];
        nd.level = d->level[i];
        for (int c = 0; c < 4; c++) nd.child[c] = d->child[4 * i + c];
        nd.leaf = nd.child[0] < 0;
        for (int c = 0; c < 4; c++) {
            if ((nd.child[c] < 0) != nd.leaf) throw Error{EF_ERR_BAD_ARG, "node with a partial set of children"};
            if (!nd.leaf) {
                if (nd.child[c] <= i || nd.child[c] >= nn) throw Error{EF_ERR_BAD_ARG, "children must follow their parent (pre-order table)"};
                H->nodes[nd.child[c]].parent = i;
            }
        }
        std::memcpy(nd.box, d->box + 4 * i, 4 * sizeof(double));
        H->max_level = std::max(H->max_level, nd.level);
        if (nd.leaf) { nd.leaf_idx = (int)H->leaf_nodes.size(); H->leaf_nodes.push_back(i); }
    }
    H->n_leaves = (int)H->leaf_nodes.size();
    for (int i = 0; i < nn; i++) if (H->nodes[i].parent < 0) H->roots.push_back(i);
    // sizes bottom-up (children have larger ids): parent = 2 * min child size (mergePatch_ :1004-1009)
    for (int i = nn - 1; i >= 0; i--) {
        NodeH& nd = H->nodes[i];
        if (nd.leaf) {
            nd.size = ext_leaf_size ? ext_leaf_size[nd.leaf_idx] : M;
            if (nd.size != 4 && (nd.size < 8 || nd.size % 8)) throw Error{EF_ERR_BAD_SHAPE, "external leaf sizes must be 4 or multiples of 8"};
            continue;
        }
        int mn = 1 << 30;
        for (int c = 0; c < 4; c++) {
            if (H->nodes[nd.child[c]].level != nd.level + 1) throw Error{EF_ERR_BAD_ARG, "child level must be parent level + 1"};
            mn = std::min(mn, H->nodes[nd.child[c]].size);
        }
        for (int c = 0; c < 4; c++) {
            NodeH& ch = H->nodes[nd.child[c]];
            int t = 0; while ((mn << t) < ch.size) t++;   // tag = log2(size) - log2(min)  (:676-696)
            ch.ncoarsen = t;
        }
        nd.size = 2 * mn;
    }
    // structural symmetry candidates (bottom-up): square patch, and for a parent four candidate children, none coarsened
    for (int i = nn - 1; i >= 0; i--) {
        NodeH& nd = H->nodes[i];
        const double wx = nd.box[1] - nd.box[0], wy = nd.box[3] - nd.box[2];
        bool ok = std::fabs(wx - wy) <= 1e-12 * std::max(std::fabs(wx), std::fabs(wy));
        if (!nd.leaf) for (int c = 0; c < 4; c++) ok = ok && H->nodes[nd.child[c]].symcand && H->nodes[nd.child[c]].ncoarsen == 0;
        nd.symcand = ok;
    }
    // batches: parents grouped by (level, child side n, symmetry candidate)
    H->level_batches.assign(H->max_level + 1, {});
    std::map<std::tuple<int, int, int>, int> key2batch;
    for (int i = 0; i < nn; i++) {
        NodeH& nd = H->nodes[i];
        if (nd.leaf) continue;
        const int n = nd.size / 2;
        auto key = std::make_tuple(nd.level, n, nd.symcand ? 1 : 0);
        auto it = key2batch.find(key);
        if (it == key2batch.end()) {
            it = key2batch.emplace(key, (int)H->batches.size()).first;
            H->batches.emplace_back();
            H->batches.back().level = nd.level; H->batches.back().n = n; H->batches.back().symcand = nd.symcand;
            H->level_batches[nd.level].push_back(it->second);
        }
        BatchH& b = H->batches[it->second];
        nd.batch = it->second; nd.slot = (int)b.parents.size();
        b.parents.push_back(i);
    }
    for (auto& b : H->batches) { b.count = (int)b.parents.size(); plan_batch_gemms(b, H->part_rank, H->part_nranks); }
    // lanes: the batches of a level in order of decreasing merge work, the heaviest on the handle's own stream
    H->n_lanes = 1;
    for (auto& lb : H->level_batches) {
        std::vector<int> order(lb);
        std::stable_sort(order.begin(), order.end(), [&](int a, int c) {
            const BatchH& x = H->batches[a]; const BatchH& y = H->batches[c];
            return (double)x.count * x.n * x.n * x.n > (double)y.count * y.n * y.n * y.n; });
        for (size_t i = 0; i < order.size(); i++) {
            H->batches[order[i]].lane = (int)(i % efgpu_handle::MAX_LANES);
            H->n_lanes = std::max(H->n_lanes, H->batches[order[i]].lane + 1);
        }
    }
    // vector arena offsets
    size_t off = 0;
    auto take = [&](size_t nd_) { size_t o = off; off += (nd_ + 1) & ~size_t(1); return o; };
    for (auto& nd : H->nodes) {
        for (int t = 0; t <= nd.ncoarsen; t++) { nd.hbuf.push_back(take(4 * (size_t)(nd.size >> t))); nd.gbuf.push_back(take(4 * (size_t)(nd.size >> t))); }
        if (!nd.leaf) { nd.w_off = take(2 * (size_t)nd.size); nd.hd_off = take(2 * (size_t)nd.size); }
    }
    H->vec_doubles = off;
    compute_flop_model(H);
}

static void allocate_device(efgpu_handle* H, unsigned flags)
{
    cudaStream_t s = H->stream;
    const int M = H->M;
    // leaf tables
    std::vector<double> Q((size_t)M * M);
    const double pi = 3.14159265358979323846264338327950288;
    for (int i = 0; i < M; i++)
        for (int k = 0; k < M; k++)
            Q[(size_t)i * M + k] = std::sqrt((k + 1 == M ? 1.0 : 2.0) / M) * std::sin((i + 0.5) * (k + 1) * pi / M);
    H->d_Q.upload(Q, s);
    std::vector<double> boxes((size_t)4 * H->n_nodes);
    for (int i = 0; i < H->n_nodes; i++) std::memcpy(&boxes[4 * (size_t)i], H->nodes[i].box, 4 * sizeof(double));
    H->d_boxes.upload(boxes, s);
    H->d_leaf_nodes.upload(H->leaf_nodes, s);
    {   // classes of leaves with bit-identical cell sizes (the kernels form dx, dy with the same expression)
        std::map<std::pair<uint64_t, uint64_t>, int> rep;
        std::vector<int> build, src(H->n_leaves);
        for (int l = 0; l < H->n_leaves; l++) {
            const double* bx = H->nodes[H->leaf_nodes[l]].box;
            const double dx = (bx[1] - bx[0]) / M, dy = (bx[3] - bx[2]) / M;
            uint64_t kx, ky; std::memcpy(&kx, &dx, 8); std::memcpy(&ky, &dy, 8);
            auto it = rep.find({kx, ky});
            if (it == rep.end()) { it = rep.emplace(std::make_pair(kx, ky), l).first; build.push_back(l); }
            src[l] = it->second;
        }
        H->n_leaf_build = (int)build.size();
        H->d_leaf_build.upload(build, s); H->d_leaf_src.upload(src, s);
    }
    H->leafT_off.assign(H->n_leaves + 1, 0);
    for (int l = 0; l < H->n_leaves; l++) { const size_t sz = 4 * (size_t)H->nodes[H->leaf_nodes[l]].size; H->leafT_off[l + 1] = H->leafT_off[l] + sz * sz; }
    // Peer mode: everything another rank stores into is carved from one arena in a fixed order (sizes depend on the tree only,
    // never on the rank), so that an address and its arena offset mean the same buffer on every rank.  [0, 4096): barrier flags.
    size_t arena_off = 4096;
    auto carve = [&](DevBuf& b, size_t bytes) {
        if (!H->peer_mode) { b.alloc(bytes); return; }
        b.adopt(static_cast<char*>(H->arena.p) + arena_off, bytes);
        arena_off += (bytes + 255) & ~size_t(255);
    };
    if (H->peer_mode) {
        if (flags & EFGPU_LEAN_T) throw Error{EF_ERR_UNSUPPORTED, "EFGPU_LEAN_T on a peer-mapped (partitioned) tree"};
        size_t need = 4096, wsm = 0;
        auto add = [&](size_t bytes) { need += (bytes + 255) & ~size_t(255); };
        add(H->leafT_off[H->n_leaves] * sizeof(double)); add(H->vec_doubles * sizeof(double));
        for (auto& b : H->batches) {
            const size_t n = b.n, cnt = b.count;
            add(cnt * 16 * n * n * sizeof(double)); add(cnt * 32 * n * n * sizeof(double)); add(cnt * 64 * n * n * sizeof(double));
            wsm = std::max(wsm, cnt * b.ws_per_entry);
        }
        add(wsm * sizeof(double));
        if (!H->arena.p) {
            H->arena.alloc(need); H->arena_need = need;
            EF_CUDA(cudaMemsetAsync(H->arena.p, 0, 4096, s));
        } else if (need != H->arena_need) throw Error{EF_ERR_STATE, "internal: the shared arena changed size after it was exported"};
    }
    carve(H->d_leafT, H->leafT_off[H->n_leaves] * sizeof(double));
    carve(H->d_vec, H->vec_doubles * sizeof(double));
    EF_CUDA(cudaMemsetAsync(H->d_vec.p, 0, H->vec_doubles * sizeof(double), s));
    if (!H->external_leaves) {
        H->d_f.alloc((size_t)H->n_leaves * M * M * sizeof(double));
        H->d_u.alloc((size_t)H->n_leaves * M * M * sizeof(double));
    }
    H->d_minpiv.alloc(5 * sizeof(double));   // pivot tracker (common.cuh: launch_invert_small) + [4] max |I - X X^-1| of the refinement
    double* vec = H->d_vec.as<double>();
    for (int l = 0; l < H->n_leaves; l++) H->nodes[H->leaf_nodes[l]].Tbuf.assign(1, H->d_leafT.as<double>() + H->leafT_off[l]);
    std::vector<double*> lh(H->n_leaves), lg(H->n_leaves);
    for (int l = 0; l < H->n_leaves; l++) { NodeH& nd = H->nodes[H->leaf_nodes[l]]; lh[l] = vec + nd.hbuf[0]; lg[l] = vec + nd.gbuf[0]; }
    H->d_leaf_h.upload(lh, s); H->d_leaf_g.upload(lg, s);

    // Lean policy: a node's DtN map is only read by its parent's merge, so the maps of tree level l live in
    // arena[l & 1] and are overwritten when level l - 2 is merged; the roots' maps (written last) stay valid.
    H->lean_T = (flags & EFGPU_LEAN_T) != 0;
    if (H->lean_T) {
        for (int r : H->roots) if (H->nodes[r].level != H->nodes[H->roots[0]].level)
            throw Error{EF_ERR_UNSUPPORTED, "EFGPU_LEAN_T needs all roots of a forest on one tree level"};
        size_t need[2] = {0, 0};
        for (int lev = 0; lev <= H->max_level; lev++) {
            size_t tot = 0;
            for (int bi : H->level_batches[lev]) { const BatchH& b = H->batches[bi]; tot += (size_t)b.count * 64 * b.n * b.n; }
            need[lev & 1] = std::max(need[lev & 1], tot);
        }
        for (int p = 0; p < 2; p++) H->d_Tarena[p].alloc(need[p] * sizeof(double));
    } else { H->d_Tarena[0].release(); H->d_Tarena[1].release(); }
    // per-batch operator storage (deepest level first so children's buffers exist before the parents' tables)
    size_t ws_max = 0;
    for (auto& b : H->batches) {   // (peer mode: arena order = batch order, as counted above)
        const size_t n = b.n, cnt = b.count;
        carve(b.Xinv, cnt * 16 * n * n * sizeof(double));
        carve(b.S, cnt * 32 * n * n * sizeof(double));
        if (H->peer_mode) carve(b.T, cnt * 64 * n * n * sizeof(double));
    }
    for (int lev = H->max_level; lev >= 0; lev--) {
        size_t tarena_off = 0;
        for (int bi : H->level_batches[lev]) {
            BatchH& b = H->batches[bi];
            const size_t n = b.n, cnt = b.count;
            b.Hc.alloc(cnt * 16 * n * n * sizeof(double));
            if (H->lean_T) { b.T.release(); b.Tbase = H->d_Tarena[lev & 1].as<double>() + tarena_off; tarena_off += cnt * 64 * n * n; }
            else if (!H->peer_mode) { b.T.alloc(cnt * 64 * n * n * sizeof(double)); b.Tbase = b.T.as<double>(); }
            else b.Tbase = b.T.as<double>();      // carved below, in batch order
            if (flags & EFGPU_KEEP_X) b.Xcopy.alloc(cnt * 16 * n * n * sizeof(double)); else b.Xcopy.release();
            ws_max = std::max(ws_max, cnt * b.ws_per_entry);
            for (size_t sl = 0; sl < cnt; sl++) H->nodes[b.parents[sl]].Tbuf.assign(1, b.Tbase + sl * 64 * n * n);
        }
    }
    // workspace: one region per lane (batches on different lanes run concurrently), sized by the largest batch of the lane
    { const char* le = getenv("EFGPU_LANES"); H->lanes_on = !H->peer_mode && H->part_nranks == 1 && H->n_lanes > 1 && (!le || atoi(le) > 0); }
    size_t ws_lane[efgpu_handle::MAX_LANES] = {0, 0, 0, 0}, ws_base[efgpu_handle::MAX_LANES] = {0, 0, 0, 0};
    for (auto& b : H->batches) { const int k = H->lanes_on ? b.lane : 0; ws_lane[k] = std::max(ws_lane[k], (size_t)b.count * b.ws_per_entry); }
    { size_t acc_ws = 0; for (int k = 0; k < efgpu_handle::MAX_LANES; k++) { ws_base[k] = acc_ws; acc_ws += (ws_lane[k] + 31) & ~size_t(31); } ws_max = acc_ws; }
    carve(H->d_ws, ws_max * sizeof(double));
    if (H->lanes_on)
        for (int k = 1; k < H->n_lanes; k++) {
            if (!H->lane[k]) EF_CUDA(cudaStreamCreateWithFlags(&H->lane[k], cudaStreamNonBlocking));
            if (!H->ev_join[k]) EF_CUDA(cudaEventCreateWithFlags(&H->ev_join[k], cudaEventDisableTiming));
        }
    if (!H->ev_fork) EF_CUDA(cudaEventCreateWithFlags(&H->ev_fork, cudaEventDisableTiming));
    H->graph_gen++;
    // coarsened copies + tables
    for (auto& b : H->batches) {
        const size_t n = b.n, cnt = b.count;
        size_t coarse_doubles = 0;
        for (int p : b.parents)
            for (int c = 0; c < 4; c++) {
                NodeH& ch = H->nodes[H->nodes[p].child[c]];
                for (int t = 1; t <= ch.ncoarsen; t++) { size_t sz = 4 * (size_t)(ch.size >> t); coarse_doubles += sz * sz; }
            }
        b.Tcoarse.alloc(coarse_doubles * sizeof(double));
        size_t coff = 0;
        b.cT.clear(); b.cH.clear(); b.cG.clear(); b.cT_slot.clear();
        std::vector<MergeEntry> ent(cnt);
        std::vector<double*> ptab(cnt * NOPS, nullptr);
        for (size_t sl = 0; sl < cnt; sl++) {
            NodeH& P = H->nodes[b.parents[sl]];
            MergeEntry& e = ent[sl];
            for (int c = 0; c < 4; c++) {
                NodeH& ch = H->nodes[P.child[c]];
                for (int t = 1; t <= ch.ncoarsen; t++) {
                    size_t sz = 4 * (size_t)(ch.size >> t);
                    ch.Tbuf.resize(t + 1);
                    ch.Tbuf[t] = b.Tcoarse.as<double>() + coff; coff += sz * sz;
                    if ((int)b.cT.size() < t) { b.cT.resize(t); b.cH.resize(t); }
                    b.cT[t - 1].push_back(CoarsenOp{ch.Tbuf[t - 1], ch.Tbuf[t], ch.size >> (t - 1), 0});
                    if ((int)b.cT_slot.size() < t) b.cT_slot.resize(t);
                    b.cT_slot[t - 1].push_back((int)sl);
                    b.cH[t - 1].push_back(CoarsenOp{vec + ch.hbuf[t - 1], vec + ch.hbuf[t], ch.size >> (t - 1), 0});
                }
                if ((ch.size >> ch.ncoarsen) != (int)n) throw Error{EF_ERR_STATE, "internal: child size mismatch after coarsening"};
                e.Tc[c] = ch.Tbuf[ch.ncoarsen];
                e.hc[c] = vec + ch.hbuf[ch.ncoarsen];
                e.gc[c] = vec + ch.gbuf[ch.ncoarsen];
                ptab[sl * NOPS + OP_TC0 + c] = ch.Tbuf[ch.ncoarsen];
            }
            e.Xinv = b.Xinv.as<double>() + sl * 16 * n * n;
            e.S = b.S.as<double>() + sl * 32 * n * n;
            e.Hc = b.Hc.as<double>() + sl * 16 * n * n;
            e.T = b.Tbase + sl * 64 * n * n;
            e.Xcopy = (flags & EFGPU_KEEP_X) ? b.Xcopy.as<double>() + sl * 16 * n * n : nullptr;
            e.hd = vec + P.hd_off; e.h = vec + P.hbuf[0]; e.w = vec + P.w_off; e.g = vec + P.gbuf[0];
            ptab[sl * NOPS + OP_XINV] = e.Xinv; ptab[sl * NOPS + OP_S] = e.S; ptab[sl * NOPS + OP_T] = e.T; ptab[sl * NOPS + OP_XCOPY] = e.Xcopy;
            double* wsb = H->d_ws.as<double>() + ws_base[H->lanes_on ? b.lane : 0];
            ptab[sl * NOPS + OP_W1] = wsb + sl * b.ws_per_entry;
            ptab[sl * NOPS + OP_W2] = wsb + sl * b.ws_per_entry + b.w2_off;
            ptab[sl * NOPS + OP_W3] = wsb + sl * b.ws_per_entry + b.w3_off;
            // this parent's own Dirichlet data arrives coarsened when it was tagged: uncoarsen before the split
            for (int t = P.ncoarsen; t >= 1; t--) {
                const int step = P.ncoarsen - t;
                if ((int)b.cG.size() <= step) b.cG.resize(step + 1);
                b.cG[step].push_back(CoarsenOp{vec + P.gbuf[t], vec + P.gbuf[t - 1], P.size >> (t - 1), 0});
            }
        }
        b.d_entries.upload(ent, s); b.d_ptab.upload(ptab, s); b.d_blocks.upload(b.blocks, s); b.d_trans.upload(b.trans, s);
        b.h_ptab = ptab; b.h_ent = ent;
        b.d_mapsA.release(); b.d_mapsB.release();
        if (!b.views.empty()) {   // tensor maps of every (parent, view): base = the slot's pointer + the view's origin
            const size_t nv = b.views.size();
            std::vector<unsigned char> ma(cnt * nv * TMA_MAP_BYTES, 0), mb(cnt * nv * TMA_MAP_BYTES, 0);
            bool encoded = true;
            try {
                for (size_t sl = 0; sl < cnt; sl++)
                    for (size_t v = 0; v < nv; v++) {
                        const double* basep = ptab[sl * NOPS + b.views[v].op];
                        if (!basep) continue;   // a slot this build does not have (the copy of X): its steps do not run either
                        encode_operand_maps(basep + b.views[v].origin, (unsigned long long)b.views[v].ld, ma.data() + (sl * nv + v) * TMA_MAP_BYTES,
                                            mb.data() + (sl * nv + v) * TMA_MAP_BYTES);
                    }
            } catch (const Error& e) {   // a driver without cuTensorMapEncodeTiled: the same products on the cp.async kernels, said once
                encoded = false;
                static bool told = false;
                if (!told) { told = true; fprintf(stderr, "efgpu: TMA operand staging unavailable (%s): merge products use the cp.async kernels\n", e.msg.c_str()); }
            }
            if (encoded) {
                b.d_mapsA.alloc(ma.size()); b.d_mapsB.alloc(mb.size());
                EF_CUDA(cudaMemcpy(b.d_mapsA.p, ma.data(), ma.size(), cudaMemcpyHostToDevice));
                EF_CUDA(cudaMemcpy(b.d_mapsB.p, mb.data(), mb.size(), cudaMemcpyHostToDevice));
                b.d_tblocks.upload(b.tblocks, s);
            }
        }
        auto up = [&](std::vector<std::vector<CoarsenOp>>& v, std::vector<std::unique_ptr<DevBuf>>& dv, std::vector<int>& mx) {
            dv.clear(); mx.clear();
            for (auto& ops : v) {
                dv.emplace_back(new DevBuf()); dv.back()->upload(ops, s);
                int m = 0; for (auto& o : ops) m = std::max(m, o.nfine);
                mx.push_back(m);
            }
        };
        up(b.cT, b.d_cT, b.cT_max); up(b.cH, b.d_cH, b.cH_max); up(b.cG, b.d_cG, b.cG_max);
    }
    H->extG.clear(); H->d_extG.clear(); H->extG_max.clear();
    if (H->external_leaves)
        for (int l = 0; l < H->n_leaves; l++) {
            NodeH& P = H->nodes[H->leaf_nodes[l]];
            for (int t = P.ncoarsen; t >= 1; t--) {
                const size_t step = (size_t)(P.ncoarsen - t);
                if (H->extG.size() <= step) H->extG.resize(step + 1);
                H->extG[step].push_back(CoarsenOp{vec + P.gbuf[t], vec + P.gbuf[t - 1], P.size >> (t - 1), 0});
            }
        }
    for (auto& ops : H->extG) {
        H->d_extG.emplace_back(new DevBuf()); H->d_extG.back()->upload(ops, s);
        int m = 0; for (auto& o : ops) m = std::max(m, o.nfine);
        H->extG_max.push_back(m);
    }
    EF_CUDA(cudaStreamSynchronize(s));
    H->allocated = true; H->build_flags = flags;
    size_t tot = 0;
    for (auto& b : H->batches) tot += b.Xinv.bytes + b.S.bytes + b.Hc.bytes + b.T.bytes + b.Xcopy.bytes + b.Tcoarse.bytes;
    H->stats.device_bytes = (double)(tot + H->d_Tarena[0].bytes + H->d_Tarena[1].bytes + H->d_leafT.bytes + H->d_vec.bytes + H->d_ws.bytes + H->d_f.bytes + H->d_u.bytes);
}

static void run_leaf_dtn(efgpu_handle* H, unsigned flags)
{
    if (H->external_leaves) return;   // leaf DtN maps were written by the caller (efgpu_operator_device)
    if (H->leaf_kind == EFGPU_LEAF_VARIABLE) {
        const int M = H->M;
        if (!H->d_coef_in[0].p) throw Error{EF_ERR_STATE, "variable-coefficient leaves: efgpu_set_leaf_variable was not called"};
        H->d_coef.alloc((size_t)H->n_leaves * 4 * M * M * sizeof(double));
        H->d_P.alloc((size_t)H->n_leaves * M * M * M * sizeof(double));
        const double* cin[6];
        for (int k = 0; k < 6; k++) cin[k] = H->d_coef_in[k].as<double>();
        timed(H, EFGPU_PROF_LEAF_LU, 1, [&] {
            launch_leaf_var_factor(M, cin, H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->d_coef.as<double>(), H->d_P.as<double>(),
                                   H->d_minpiv.as<double>(), H->n_leaves, H->stream);
        });
        const bool cache = (flags & EFGPU_CACHE_OPERATORS) != 0;   // quirk q1: the first leaf's T for every leaf
        launch_leaf_var_solve(M, H->d_coef.as<double>(), H->d_P.as<double>(), H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), nullptr, 1.0,
                              nullptr, nullptr, nullptr, H->d_leafT.as<double>(), 2, cache ? 1 : H->n_leaves, H->stream);
        if (cache) launch_broadcast_leaf_T(H->d_leafT.as<double>(), M, H->n_leaves, H->stream);
        return;
    }
    launch_leaf_dtn_const(H->M, H->d_Q.as<double>(), H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->lambda,
                          H->d_leafT.as<double>(), H->n_leaves, (flags & EFGPU_CACHE_OPERATORS) != 0,
                          H->d_leaf_build.as<int>(), H->n_leaf_build, H->d_leaf_src.as<int>(), H->stream);
}

// Flag barrier over the ranks of a peer-mapped tree (peer.cu), stream-ordered.  Discipline: a barrier BEFORE a kernel that stores
// into peer arenas (every rank has finished reading what is about to be overwritten) unless nothing ran since the last
// barrier, and one AFTER it (the stores of every rank have landed before anyone reads them).
static void peer_barrier(efgpu_handle* H, bool only_if_dirty)
{
    if (only_if_dirty && !H->peer_dirty) return;
    timed(H, EFGPU_PROF_ALLGATHER, 1, [&] {
        launch_peer_barrier(H->peers, H->part_rank, H->d_peer_err.as<int>(), H->stream);
    });
    H->peer_dirty = false;
}

// buildStage in pieces, so that a replicated upper tree can exchange row slices between them:
//   build_begin            allocation, pivot tracker, leaf DtN maps
//   build_level(lev, 0)    coarsen + assemble X / H + invert X + (this rank's rows of) S for every merge of level `lev`
//   build_level(lev, 1)    (this rank's rows of) T
//   build_end              synchronise, singularity report
static void build_begin(efgpu_handle* H, unsigned flags)
{
    // Indefinite problems (lambda > 0: tolerated by the reference, hstcrt.f:450-452 / FiniteVolumeSolver.cpp:270): the merge
    // matrices lose positive definiteness and the unpivoted block inversion can lose digits against the reference's pivoted
    // dgesv (measured on the CPU emulation of the plan: 2e-10 instead of 1e-12 next to a resonance).  Every X^-1 then gets one
    // Newton-Schulz step, which squares the residual I - X X^-1; it needs X, so such builds keep the copy of EFGPU_KEEP_X.
    const double lam_hi = H->external_leaves ? 0.0 : (H->leaf_kind == EFGPU_LEAF_CONSTANT ? H->lambda : H->lambda_max);
    H->refine_inverse = H->refine_mode == 1 || (H->refine_mode < 0 && lam_hi > 0.0);
    if (H->refine_inverse) flags |= EFGPU_KEEP_X;
    if (!H->allocated || ((flags ^ H->build_flags) & (EFGPU_KEEP_X | EFGPU_LEAN_T))) allocate_device(H, flags);
    cudaStream_t s = H->stream;
    // symmetric plan where the structure allows it and the leaf DtN maps are signed-symmetric: constant-coefficient leaves
    // (with variable beta the map takes beta-weighted Dirichlet data to unweighted derivatives, FiniteVolumeSolver.cpp:100-175
    // vs :332-343, and is not symmetric); external leaves: as declared by efgpu_set_symmetric_leaves
    const bool leaves_sym = H->external_leaves ? H->ext_sym : H->leaf_kind == EFGPU_LEAF_CONSTANT;
    for (auto& b : H->batches) b.use_sym = b.symcand && leaves_sym && !(flags & EFGPU_NO_SYMMETRY);
    H->cur_flags = flags;
    compute_flop_model(H);
    EF_CUDA(cudaEventRecord(H->ev0, s));
    H->built = false; H->root_T_distributed = false; H->root_T_pending = false;
}

// device work of build_begin: pivot tracker, leaf DtN maps (part of the captured build graph)
static void build_leaves(efgpu_handle* H)
{
    launch_pivot_tracker_reset(H->d_minpiv.as<double>(), H->stream);
    timed(H, EFGPU_PROF_LEAF_DTN, 1, [&] { run_leaf_dtn(H, H->cur_flags); });
}

static void build_level(efgpu_handle* H, int lev, int phase)
{
    if (lev < 0 || lev > H->max_level) throw Error{EF_ERR_BAD_ARG, "bad level"};
    if (lev == 0 && phase == 1 && (H->cur_flags & EFGPU_LAZY_ROOT_DTN) && !H->level_batches[0].empty()) {
        // the DtN map of the whole domain is read by nothing on the Dirichlet path (only by the root's Robin system, a parent
        // that does not exist, and parity readers): its products are issued by complete_root_T when somebody asks for it
        H->root_T_pending = true;
        if (H->peer_pending) { peer_barrier(H, false); H->peer_pending = false; }
        return;
    }
    H->peer_dirty = true;
    LaneSet lanes(H);
    for (int bi : H->level_batches[lev]) {
        BatchH& b = H->batches[bi];
        cudaStream_t s = lanes.get(b);
        // adaptive re-build: only the dirty parents of the batch (compact tables), possibly none
        if (b.sub_on && b.sub_count == 0) continue;
        const int bcount = b.sub_on ? b.sub_count : b.count;
        double* const* ptab = b.sub_on ? b.sub_ptab.as<double*>() : b.d_ptab.as<double*>();
        if (phase == 0) {
            if (b.sub_on) {
                for (size_t t = 0; t < b.sub_cT.size(); t++)
                    if (b.sub_cT_n[t]) timed(H, EFGPU_PROF_COARSEN_T, 1, [&] { launch_coarsen_T(b.sub_cT[t]->as<CoarsenOp>(), b.sub_cT_n[t], b.sub_cT_max[t], s); });
            } else
                for (size_t t = 0; t < b.cT.size(); t++)
                    timed(H, EFGPU_PROF_COARSEN_T, 1, [&] { launch_coarsen_T(b.d_cT[t]->as<CoarsenOp>(), (int)b.cT[t].size(), b.cT_max[t], s); });
            const MergeEntry* ent = b.sub_on ? b.sub_entries.as<MergeEntry>() : b.d_entries.as<MergeEntry>();
            timed(H, EFGPU_PROF_ASSEMBLE, 2, [&] {
                launch_assemble_X(ent, b.n, bcount, s);
                launch_assemble_Hc(ent, b.n, bcount, s);
            });
        }
        auto gather = [&](double* buf, size_t doubles_total) {
            if (!H->allgather) throw Error{EF_ERR_STATE, "row-partitioned tree without an all-gather callback (efgpu_set_allgather)"};
            if (H->allgather(buf, doubles_total / H->part_nranks * sizeof(double), H->allgather_user) != 0)
                throw Error{EF_ERR_STATE, "the all-gather callback failed"};
        };
        auto run_transposes = [&](const Step& st) {
            timed(H, st.cls, 1, [&] { launch_btranspose(ptab, NOPS, b.d_trans.as<TransOp>() + st.first, b.trans.data() + st.first, st.count, bcount, s); });
        };
        // TMA operand staging where the step's blocks all have views (not on the compact sub-batches of an adaptive re-build: the
        // tensor maps are indexed by parent slot)
        auto tma_args = [&](const Step& st, TmaArgs& ta) -> const TmaArgs* {
            if (!st.tma || b.sub_on || !b.d_mapsA.p || !b.d_tblocks.p) return nullptr;
            ta.d_tblocks = b.d_tblocks.as<TmaBlock>() + st.first; ta.mapsA = b.d_mapsA.p; ta.mapsB = b.d_mapsB.p; ta.nviews = (int)b.views.size();
            return &ta;
        };
        auto run_refine = [&]() {   // indefinite problems: X^-1 <- X^-1 + X^-1 (I - X X^-1), every rank of a partition alike
            for (const Step& st : b.steps_refine)
                timed(H, st.cls, 1, [&] {
                    if (st.kind == 1) { TmaArgs ta; launch_bgemm(ptab, NOPS, b.d_blocks.as<GemmBlock>() + st.first, b.blocks.data() + st.first, st.count, bcount, s, 0, nullptr, tma_args(st, ta)); }
                    else launch_refine_ew(ptab, NOPS, st.kind, st.kind == 3 ? OP_T : OP_XINV, st.kind == 3 ? st.off : 0, OP_T, st.off, st.N, bcount,
                                          H->d_minpiv.as<double>() + 4, s);
                });
        };
        const bool p2p = H->peer_mode && H->part_nranks > 1;
        if (p2p && !H->peer_attached) throw Error{EF_ERR_STATE, "peer-mapped tree: efgpu_peer_attach has not been called"};
        for (const Step& st : b.active()) {
            if (st.cls == EFGPU_PROF_MIRROR_T) continue;   // after the gather of T, below
            const bool is_T = st.cls == EFGPU_PROF_GEMM_T;
            if (phase == 0 && st.cls == EFGPU_PROF_GEMM_S && H->refine_inverse) run_refine();
            if (is_T != (phase == 1) || (st.kind == 1 && st.count == 0)) continue;
            if (st.kind == 2) { run_transposes(st); H->peer_dirty = true; continue; }
            // peer mode: the row slices of a split product, of S and of the DtN maps below the root are stored into every rank's
            // arena by the GEMM itself, between two flag barriers (the root's map stays row-distributed: nobody merges it)
            const bool scatter = p2p && st.kind == 1 && (st.gk == 3 || st.cls == EFGPU_PROF_GEMM_S || (is_T && lev > 0));
            // column-partitioned S / T: the T products read S only in this rank's own columns, so the stores of S into the peers
            // need no barrier of their own - the one after T (or, at the root, the one issued below) completes both
            const bool defer = scatter && b.colsplit && st.cls == EFGPU_PROF_GEMM_S;
            if (H->trace) {
                char lb[160];
                if (st.kind == 1) {
                    const GemmBlock& g0 = b.blocks[st.first];
                    snprintf(lb, sizeof lb, "lev %d n %d batch %d gemm blocks %d rows %d cols %d K %d terms %d ct %d gk %d", lev, b.n, bcount, st.count, g0.rows, g0.cols,
                             g0.t[0].K, g0.nterms, g0.ct_op1 ? 1 : 0, st.gk);
                } else snprintf(lb, sizeof lb, "lev %d n %d batch %d invert N %d pair %d", lev, b.n, bcount, st.N, st.off2 >= 0 ? 1 : 0);
                H->cur_label = lb;
            }
            if (scatter) peer_barrier(H, /*only_if_dirty=*/true);
            timed(H, st.cls, 1, [&] {
                if (st.kind == 0) launch_invert_small(ptab, NOPS, OP_XINV, st.off, st.off2, 4 * b.n, st.N, bcount, H->d_minpiv.as<double>(), s);
                else { TmaArgs ta; launch_bgemm(ptab, NOPS, b.d_blocks.as<GemmBlock>() + st.first, b.blocks.data() + st.first, st.count, bcount, s, 0, scatter ? &H->peers : nullptr, tma_args(st, ta)); }
            });
            if (defer) { H->peer_pending = true; continue; }
            if (scatter) { peer_barrier(H, false); H->peer_pending = false; continue; }
            H->peer_dirty = true;
            if (is_T && H->peer_pending) { peer_barrier(H, false); H->peer_pending = false; }   // root: T stays distributed, S is completed here
            if (st.gk) timed(H, EFGPU_PROF_ALLGATHER, 0, [&] {
                const size_t hh = (size_t)st.g_rows * st.g_cols;
                for (int sl = 0; sl < b.count; sl++) {
                    double* const* ops = b.h_ptab.data() + (size_t)sl * NOPS;
                    if (st.gk == 1) gather(ops[st.g_op] + st.g_off, hh);
                    else {
                        gather(ops[OP_W3], hh);
                        EF_CUDA(cudaMemcpy2DAsync(ops[st.g_op] + st.g_off, (size_t)st.g_ld * sizeof(double), ops[OP_W3], (size_t)st.g_cols * sizeof(double),
                                                  (size_t)st.g_cols * sizeof(double), (size_t)st.g_rows, cudaMemcpyDeviceToDevice, s));
                    }
                }
            });
        }
        // the row slices of S (phase 0) and of the DtN map T (phase 1; the root's stays distributed) become whole on every rank
        if (H->part_nranks > 1 && !p2p && H->allgather && (phase == 0 || lev > 0)) timed(H, EFGPU_PROF_ALLGATHER, 0, [&] {
            const size_t n2 = (size_t)b.n * b.n;
            for (int sl = 0; sl < b.count; sl++) {
                double* const* ops = b.h_ptab.data() + (size_t)sl * NOPS;
                if (phase == 0) gather(ops[OP_S], 32 * n2); else gather(ops[OP_T], 64 * n2);
            }
        });
        if (phase == 1) {
            if (H->part_nranks > 1 && lev == 0) H->root_T_distributed = true;   // completed by complete_root_T when somebody needs it
            else for (const Step& st : b.active()) if (st.cls == EFGPU_PROF_MIRROR_T) run_transposes(st);
        }
    }
    lanes.join();
}

// Partitioned tree: the root's DtN map is only needed by the Robin solve and by parity readers, so its row slices stay
// where they were computed until this (collective) call gathers them and mirrors the blocks of the symmetric plan.
static void complete_root_T(efgpu_handle* H)
{
    if (H->root_T_pending) {   // EFGPU_LAZY_ROOT_DTN: phase 1 of level 0 now (on every rank of a partition: this call is collective)
        if (!H->built) throw Error{EF_ERR_STATE, "root DtN map requested before the build has finished"};
        H->root_T_pending = false;
        H->cur_flags &= ~(unsigned)EFGPU_LAZY_ROOT_DTN;
        build_level(H, 0, 1);
        // the map exists from now on: the flop model counts its products again
        compute_flop_model(H);
        EF_CUDA(cudaStreamSynchronize(H->stream));
        collect_profile(H);
    }
    if (!H->root_T_distributed) return;
    const bool p2p = H->peer_mode && H->part_nranks > 1;
    if (!p2p && !H->allgather) throw Error{EF_ERR_STATE, "row-partitioned tree without an all-gather callback (efgpu_set_allgather)"};
    cudaStream_t s = H->stream;
    if (p2p) peer_barrier(H, false);
    for (int bi : H->level_batches[0]) {
        BatchH& b = H->batches[bi];
        const size_t n2 = (size_t)b.n * b.n;
        for (int sl = 0; sl < b.count; sl++) {
            double* T = b.h_ptab[(size_t)sl * NOPS + OP_T];
            const size_t slice = 64 * n2 / H->part_nranks * sizeof(double);
            if (p2p && b.colsplit) {   // this rank's block columns (all 8 n rows) into every other arena
                const size_t row_bytes = 8 * (size_t)b.n / H->part_nranks * sizeof(double);
                launch_peer_scatter2d(H->peers, H->part_rank, (size_t)(reinterpret_cast<char*>(T) - static_cast<char*>(H->arena.p)) + H->part_rank * row_bytes,
                                      8 * (size_t)b.n, row_bytes, 8 * (size_t)b.n * sizeof(double), s);
                continue;
            }
            if (p2p) {   // this rank's row slice into every other arena; the barrier below completes the all-gather
                launch_peer_scatter(H->peers, H->part_rank, (size_t)(reinterpret_cast<char*>(T) - static_cast<char*>(H->arena.p)) + H->part_rank * slice, slice, s);
                continue;
            }
            if (H->allgather(T, slice, H->allgather_user) != 0)
                throw Error{EF_ERR_STATE, "the all-gather callback failed"};
        }
    }
    if (p2p) peer_barrier(H, false);
    for (int bi : H->level_batches[0]) {
        BatchH& b = H->batches[bi];
        for (const Step& st : b.active())
            if (st.cls == EFGPU_PROF_MIRROR_T)
                launch_btranspose(b.d_ptab.as<double*>(), NOPS, b.d_trans.as<TransOp>() + st.first, b.trans.data() + st.first, st.count, b.count, s);
    }
    EF_CUDA(cudaStreamSynchronize(s));
    H->root_T_distributed = false;
}

static void build_end(efgpu_handle* H)
{
    cudaStream_t s = H->stream;
    EF_CUDA(cudaEventRecord(H->ev1, s));
    int peer_err = 0;
    if (H->peer_mode && H->d_peer_err.p) EF_CUDA(cudaMemcpyAsync(&peer_err, H->d_peer_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    double trk[5] = {0, 0, 0, 0, 0};
    EF_CUDA(cudaMemcpyAsync(trk, H->d_minpiv.p, sizeof(trk), cudaMemcpyDeviceToHost, s));
    EF_CUDA(cudaStreamSynchronize(s));
    collect_profile(H);
    float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1));
    if (peer_err) {
        EF_CUDA(cudaMemsetAsync(H->d_peer_err.p, 0, sizeof(int), s));
        throw Error{EF_ERR_STATE, "peer barrier timed out waiting for rank " + std::to_string(peer_err - 1) + " (a rank of the partition did not reach the same step)"};
    }
    const double minpiv = trk[0];
    unsigned long long nneg = 0; std::memcpy(&nneg, &trk[3], sizeof(nneg));
    const bool none = trk[1] == 0.0;   // no base-case inversion ran (a handle without merges): nothing to report
    H->stats.build_ms = ms; H->stats.min_pivot = minpiv;
    H->stats.max_pivot = trk[1]; H->stats.pivot_ratio_min = none ? 1.0 : trk[2]; H->stats.negative_pivots = (double)nneg;
    H->stats.inverse_residual = H->refine_inverse ? trk[4] : -1.0;
    H->built = true; H->upwards_done = false;
    if (!(minpiv > 0.0) || !std::isfinite(minpiv) || !std::isfinite(trk[1]))
        throw Error{EF_ERR_SINGULAR, "zero or non-finite pivot in the merge factorisation"};
    // The merge matrices are inverted WITHOUT pivoting (the reference: dgesv, partial pivoting).  That is safe for the SPD /
    // diagonally dominant X of lambda <= 0; for indefinite problems (lambda > 0) a leading block can be nearly singular while
    // X is not.  A base-case block whose pivots span more than pivot_ratio_limit (default 1e10, EFGPU_PIVOT_RATIO_LIMIT) is
    // reported as singular instead of returning operators of unknown accuracy.
    static const double limit = [] { const char* e = getenv("EFGPU_PIVOT_RATIO_LIMIT"); const double v = e ? atof(e) : 0.0; return v > 1.0 ? v : 1e10; }();
    if (!none && trk[2] * limit < 1.0)
        throw Error{EF_ERR_SINGULAR, "ill-conditioned pivot block in the unpivoted merge factorisation (min/max |pivot| = " + std::to_string(trk[2]) +
                                     " in one base-case block, " + std::to_string(nneg) + " negative pivots): result would not meet the 1e-10 parity tolerance"};
    // Newton-Schulz squares the residual: E = I - X X^-1 with max |E_ij| above 1e-4 leaves more than ~1e-8 * N behind
    if (H->refine_inverse && !(trk[4] < 1e-4))
        throw Error{EF_ERR_SINGULAR, "unpivoted inversion of an indefinite merge matrix too inaccurate to refine (max |I - X X^-1| = " + std::to_string(trk[4]) + ")"};
}

static void do_build(efgpu_handle* H, unsigned flags)
{
    build_begin(H, flags);
    run_graphed(H, H->g_build, graph_key(H, H->cur_flags), [&] {
        build_leaves(H);
        for (int lev = H->max_level; lev >= 0; lev--) { build_level(H, lev, 0); build_level(H, lev, 1); }
    });
    // host-side state a replayed graph does not set
    if ((H->cur_flags & EFGPU_LAZY_ROOT_DTN) && !H->level_batches[0].empty()) H->root_T_pending = true;
    build_end(H);
}

static void do_upwards(efgpu_handle* H, const double* f_dev, double fscale, unsigned flags)
{
    if (!H->built) throw Error{EF_ERR_STATE, "upwards before build"};
    cudaStream_t s = H->stream;
    EF_CUDA(cudaEventRecord(H->ev0, s));
    if (!H->external_leaves) timed(H, EFGPU_PROF_LEAF_SOLVE, 1, [&] {
        if (H->leaf_kind == EFGPU_LEAF_VARIABLE)
            launch_leaf_var_solve(H->M, H->d_coef.as<double>(), H->d_P.as<double>(), H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), f_dev, fscale,
                                  nullptr, nullptr, H->d_leaf_h.as<double*>(), nullptr, 1, H->n_leaves, s);
        else
            launch_leaf_solve_const(H->M, H->d_Q.as<double>(), H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->lambda,
                                    f_dev, fscale, nullptr, nullptr, H->d_leaf_h.as<double*>(), 1, H->n_leaves, s);
    });
    if (!(flags & EFGPU_HOMOGENEOUS_RHS))   // upwards4to1 is skipped entirely (HPSAlgorithm.hpp:532)
        run_graphed(H, H->g_up, graph_key(H, 0), [&] {
            for (int lev = H->max_level; lev >= 0; lev--) {
                LaneSet lanes(H);
                for (int bi : H->level_batches[lev]) {
                    BatchH& b = H->batches[bi];
                    cudaStream_t ls = lanes.get(b);
                    for (size_t t = 0; t < b.cH.size(); t++)
                        timed(H, EFGPU_PROF_COARSEN_VEC, 1, [&] { launch_coarsen_h(b.d_cH[t]->as<CoarsenOp>(), (int)b.cH[t].size(), b.cH_max[t], ls); });
                    timed(H, EFGPU_PROF_UPWARDS_MATVEC, 3, [&] { launch_upwards(b.d_entries.as<MergeEntry>(), b.n, b.count, ls); });
                }
                lanes.join();
            }
        });
    EF_CUDA(cudaEventRecord(H->ev1, s));
    H->upwards_done = true;
}

static void do_solve(efgpu_handle* H, const double* f_dev, double fscale, unsigned flags)
{
    // root Dirichlet data must already be in the root's g buffer
    cudaStream_t s = H->stream;
    const bool homogeneous = (flags & EFGPU_HOMOGENEOUS_RHS) != 0;
    run_graphed(H, H->g_solve, graph_key(H, homogeneous ? 1u : 0u), [&] {
        for (int lev = 0; lev <= H->max_level; lev++) {
            LaneSet lanes(H);
            for (int bi : H->level_batches[lev]) {
                BatchH& b = H->batches[bi];
                cudaStream_t ls = lanes.get(b);
                for (size_t t = 0; t < b.cG.size(); t++)
                    timed(H, EFGPU_PROF_COARSEN_VEC, 1, [&] { launch_uncoarsen_g(b.d_cG[t]->as<CoarsenOp>(), (int)b.cG[t].size(), b.cG_max[t], ls); });
                timed(H, EFGPU_PROF_SOLVE_MATVEC, 1, [&] { launch_solve_split(b.d_entries.as<MergeEntry>(), b.n, b.count, !homogeneous, ls); });
            }
            lanes.join();
        }
        for (size_t t = 0; t < H->extG.size(); t++)
            timed(H, EFGPU_PROF_COARSEN_VEC, 1, [&] { launch_uncoarsen_g(H->d_extG[t]->as<CoarsenOp>(), (int)H->extG[t].size(), H->extG_max[t], s); });
    });
    if (!H->external_leaves) timed(H, EFGPU_PROF_LEAF_SOLVE, 1, [&] {
        if (H->leaf_kind == EFGPU_LEAF_VARIABLE)
            launch_leaf_var_solve(H->M, H->d_coef.as<double>(), H->d_P.as<double>(), H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(),
                                  homogeneous ? nullptr : f_dev, fscale, H->d_leaf_g.as<double*>(), H->d_u.as<double>(), nullptr, nullptr, 0, H->n_leaves, s);
        else
            launch_leaf_solve_const(H->M, H->d_Q.as<double>(), H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->lambda,
                                    homogeneous ? nullptr : f_dev, fscale, H->d_leaf_g.as<double*>(), H->d_u.as<double>(), nullptr, 0, H->n_leaves, s);
    });
    H->solve_done = !H->external_leaves;
}

// ---- adaptive re-build -----------------------------------------------------------------------------------
// The paper advertises re-using the factorisation when the mesh is refined or coarsened locally (paper.md:44); the reference never
// implements it (HPSAlgorithm::isBuilt is unused, src/HPSAlgorithm.hpp:50-55: every buildStage starts from the leaves).  Here a
// node of the new tree is CLEAN when an identical subtree (same boxes bit for bit, same structure) exists in the old, built tree:
// its X^-1, S, H and DtN map are copied device to device, and the build runs on the other (dirty) parents only - the ancestor
// chains of what changed.  Every kernel sees the same operands as in a build from scratch and the batched kernels compute each
// batch entry independently of the others, so the result is bit-identical to efgpu_build on the new tree.
struct SubtreeKey {
    uint64_t h; uint64_t box[4];
    bool operator<(const SubtreeKey& o) const { return std::tie(h, box[0], box[1], box[2], box[3]) < std::tie(o.h, o.box[0], o.box[1], o.box[2], o.box[3]); }
};
static std::vector<SubtreeKey> subtree_keys(const efgpu_handle* H)
{
    std::vector<SubtreeKey> k(H->n_nodes);
    auto mix = [](uint64_t a, uint64_t b) { a ^= b + 0x9e3779b97f4a7c15ull + (a << 6) + (a >> 2); return a * 0xff51afd7ed558ccdull; };
    for (int i = H->n_nodes - 1; i >= 0; i--) {   // children have larger ids
        const NodeH& nd = H->nodes[i];
        std::memcpy(k[i].box, nd.box, sizeof(k[i].box));
        uint64_t h = mix(0x1234567ull, (uint64_t)nd.size);
        for (int c = 0; c < 4; c++) h = mix(h, k[i].box[c]);
        if (!nd.leaf) for (int c = 0; c < 4; c++) { h = mix(h, k[nd.child[c]].h); h = mix(h, (uint64_t)H->nodes[nd.child[c]].ncoarsen); }
        k[i].h = h;
    }
    return k;
}


static void do_rebuild_from(efgpu_handle* H, efgpu_handle* old, unsigned flags, double* reused_nodes, double* rebuilt_merges)
{
    if (!old->built) throw Error{EF_ERR_STATE, "efgpu_rebuild_from: the old handle has not been built"};
    if (old->device != H->device || old->M != H->M) throw Error{EF_ERR_BAD_ARG, "efgpu_rebuild_from: handles on different devices / patch sizes"};
    if (H->external_leaves || old->external_leaves || H->part_nranks > 1 || old->part_nranks > 1)
        throw Error{EF_ERR_UNSUPPORTED, "efgpu_rebuild_from: plain (unsharded) handles only"};
    if ((flags | old->build_flags) & EFGPU_LEAN_T) throw Error{EF_ERR_UNSUPPORTED, "efgpu_rebuild_from: the DtN maps of interior nodes must be resident (no EFGPU_LEAN_T)"};
    if (H->leaf_kind != old->leaf_kind || (H->leaf_kind == EFGPU_LEAF_CONSTANT && H->lambda != old->lambda))
        throw Error{EF_ERR_BAD_ARG, "efgpu_rebuild_from: the leaf model changed - every operator is dirty, use efgpu_build"};
    if ((old->cur_flags ^ flags) & (EFGPU_NO_SYMMETRY | EFGPU_CACHE_OPERATORS | EFGPU_LAZY_ROOT_DTN))
        throw Error{EF_ERR_BAD_ARG, "efgpu_rebuild_from: build flags differ from the old build"};
    if (flags & EFGPU_CACHE_OPERATORS) throw Error{EF_ERR_UNSUPPORTED, "efgpu_rebuild_from with cache-operators (a uniform-mesh option)"};
    build_begin(H, flags);
    cudaStream_t s = H->stream;
    EF_CUDA(cudaStreamSynchronize(old->stream));
    // clean nodes: identical subtree in the old tree, and merged by the same plan (symmetric / general) there
    const std::vector<SubtreeKey> kn = subtree_keys(H), ko = subtree_keys(old);
    std::map<SubtreeKey, int> where;
    for (int i = 0; i < old->n_nodes; i++) where.emplace(ko[i], i);
    std::vector<int> twin(H->n_nodes, -1);
    for (int i = 0; i < H->n_nodes; i++) {
        auto it = where.find(kn[i]);
        if (it == where.end()) continue;
        const NodeH& a = H->nodes[i]; const NodeH& o = old->nodes[it->second];
        if (a.leaf != o.leaf) continue;
        if (!a.leaf && H->batches[a.batch].use_sym != old->batches[o.batch].use_sym) continue;
        twin[i] = it->second;
    }
    // leaves: every leaf map is recomputed (constant coefficients: one map per class of cell sizes; variable: the block LU is
    // needed by the leaf solves of the new handle anyway)
    build_leaves(H);
    // copies
    std::vector<CopyOp> cp;
    double nclean = 0, ndirty = 0;
    for (int i = 0; i < H->n_nodes; i++) {
        const NodeH& a = H->nodes[i];
        if (a.leaf) continue;
        if (twin[i] < 0) { ndirty++; continue; }
        nclean++;
        const NodeH& o = old->nodes[twin[i]];
        const size_t n = a.size / 2;
        BatchH& bn = H->batches[a.batch]; BatchH& bo = old->batches[o.batch];
        cp.push_back({bo.Xinv.as<double>() + (size_t)o.slot * 16 * n * n, bn.Xinv.as<double>() + (size_t)a.slot * 16 * n * n, 16 * n * n});
        cp.push_back({bo.S.as<double>() + (size_t)o.slot * 32 * n * n, bn.S.as<double>() + (size_t)a.slot * 32 * n * n, 32 * n * n});
        cp.push_back({bo.Hc.as<double>() + (size_t)o.slot * 16 * n * n, bn.Hc.as<double>() + (size_t)a.slot * 16 * n * n, 16 * n * n});
        cp.push_back({o.Tbuf[0], a.Tbuf[0], 64 * n * n});
        if (bn.Xcopy.p && bo.Xcopy.p) cp.push_back({bo.Xcopy.as<double>() + (size_t)o.slot * 16 * n * n, bn.Xcopy.as<double>() + (size_t)a.slot * 16 * n * n, 16 * n * n});
        else if (bn.Xcopy.p) throw Error{EF_ERR_STATE, "efgpu_rebuild_from: X is kept (EFGPU_KEEP_X / refinement) in the new build but was not in the old one"};
        // the coarsened copies of a clean parent's children are only read by that parent's merge, which does not run again;
        // they are copied all the same so that the parity accessors see them (quirk q3: the child's stored T is the coarsened one)
        for (int c = 0; c < 4; c++) {
            const NodeH& ca = H->nodes[a.child[c]]; const NodeH& co = old->nodes[o.child[c]];
            for (int t = 1; t <= ca.ncoarsen; t++) { const size_t sz = 4 * (size_t)(ca.size >> t); cp.push_back({co.Tbuf[t], ca.Tbuf[t], sz * sz}); }
        }
    }
    DevBuf d_cp;
    if (!cp.empty()) { d_cp.upload(cp, s); launch_copy_many(d_cp.as<CopyOp>(), (int)cp.size(), s); }
    // dirty subsets of every batch
    for (auto& b : H->batches) {
        std::vector<int> slots;
        for (int sl = 0; sl < b.count; sl++) if (twin[b.parents[sl]] < 0) slots.push_back(sl);
        b.sub_on = true; b.sub_count = (int)slots.size();
        if (slots.empty()) continue;
        std::vector<MergeEntry> ent(slots.size());
        std::vector<double*> ptab(slots.size() * NOPS);
        const int lane = H->lanes_on ? b.lane : 0;
        (void)lane;
        for (size_t k = 0; k < slots.size(); k++) {
            ent[k] = b.h_ent[slots[k]];
            for (int o = 0; o < NOPS; o++) ptab[k * NOPS + o] = b.h_ptab[(size_t)slots[k] * NOPS + o];
            // workspace: the compact index addresses the lane's region (its size covers the whole batch)
            double* ws0 = b.h_ptab[OP_W1];
            ptab[k * NOPS + OP_W1] = ws0 + k * b.ws_per_entry;
            ptab[k * NOPS + OP_W2] = ws0 + k * b.ws_per_entry + b.w2_off;
            ptab[k * NOPS + OP_W3] = ws0 + k * b.ws_per_entry + b.w3_off;
        }
        b.sub_entries.upload(ent, s); b.sub_ptab.upload(ptab, s);
        std::vector<char> dirty(b.count, 0);
        for (int sl : slots) dirty[sl] = 1;
        b.sub_cT.clear(); b.sub_cT_n.clear(); b.sub_cT_max.clear();
        for (size_t t = 0; t < b.cT.size(); t++) {
            std::vector<CoarsenOp> ops; int mx = 0;
            for (size_t k = 0; k < b.cT[t].size(); k++) if (dirty[b.cT_slot[t][k]]) { ops.push_back(b.cT[t][k]); mx = std::max(mx, b.cT[t][k].nfine); }
            b.sub_cT.emplace_back(new DevBuf()); if (!ops.empty()) b.sub_cT.back()->upload(ops, s);
            b.sub_cT_n.push_back((int)ops.size()); b.sub_cT_max.push_back(mx);
        }
    }
    EF_CUDA(cudaStreamSynchronize(s));   // the host staging vectors go out of scope
    struct Off { efgpu_handle* H; ~Off() { for (auto& b : H->batches) { b.sub_on = false; b.sub_entries.release(); b.sub_ptab.release(); b.sub_cT.clear(); } } } off{H};
    for (int lev = H->max_level; lev >= 0; lev--) { build_level(H, lev, 0); build_level(H, lev, 1); }
    if ((H->cur_flags & EFGPU_LAZY_ROOT_DTN) && !H->level_batches[0].empty()) H->root_T_pending = true;
    // flops actually issued by this re-build
    double issued = 0;
    for (auto& b : H->batches) {
        for (const Step& st : b.active())
            if (st.kind == 1) for (int k = st.first; k < st.first + st.count; k++)
                for (int t = 0; t < b.blocks[k].nterms; t++) issued += b.sub_count * 2.0 * b.blocks[k].rows * b.blocks[k].cols * b.blocks[k].t[t].K;
    }
    build_end(H);
    H->stats.merge_flops_issued = issued;
    if (reused_nodes) *reused_nodes = nclean;
    if (rebuilt_merges) *rebuilt_merges = ndirty;
}

}  // namespace efgpu

// =================================================================================================
// C-ABI
// =================================================================================================
#define EF_TRY(H) try {
#define EF_CATCH(H)                                                                     \
    } catch (const efgpu::Error& e) { if (H) (H)->last_error = e.msg; return e.code; }  \
      catch (const std::bad_alloc&) { if (H) (H)->last_error = "host out of memory"; return EF_ERR_OOM; } \
      catch (const std::exception& e) { if (H) (H)->last_error = e.what(); return EF_ERR_STATE; }          \
    return EF_OK;

static thread_local std::string g_create_error;

// EFGPU_LEAN_T: only the DtN maps of leaves (own buffer) and of the roots (merged last) outlive the build
static void require_T_retained(efgpu_handle* H, int node)
{
    const efgpu::NodeH& nd = H->nodes[node];
    if (H->root_T_pending && nd.parent < 0 && !nd.leaf && nd.level == 0) {   // EFGPU_LAZY_ROOT_DTN: first reader
        if (H->part_nranks > 1)
            throw efgpu::Error{EF_ERR_STATE, "the root's DtN map of a partitioned tree has not been formed (EFGPU_LAZY_ROOT_DTN): call efgpu_complete_root_dtn on every rank first"};
        EF_CUDA(cudaSetDevice(H->device));
        efgpu::complete_root_T(H);
    }
    if (H->root_T_distributed && nd.parent < 0 && !nd.leaf)
        throw efgpu::Error{EF_ERR_STATE, "the root's DtN map of a partitioned tree is row-distributed: call efgpu_complete_root_dtn on every rank first"};
    if (H->lean_T && !nd.leaf && nd.parent >= 0)
        throw efgpu::Error{EF_ERR_STATE, "the DtN map of an interior node is not retained under EFGPU_LEAN_T"};
}

extern "C" {

int efgpu_create(const efgpu_tree_desc* desc, int device, efgpu_handle** out) { return efgpu_create_ex(desc, device, nullptr, out); }

int efgpu_create_ex(const efgpu_tree_desc* desc, int device, const int32_t* external_leaf_size, efgpu_handle** out)
{
    if (!desc || !out) return EF_ERR_BAD_ARG;
    efgpu_handle* H = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw Error{EF_ERR_CUDA, "no CUDA device available: the efgpu path has no CPU fallback"};
        if (device < 0 || device >= ndev) throw Error{EF_ERR_BAD_ARG, "bad device ordinal"};
        EF_CUDA(cudaSetDevice(device));
        H = new efgpu_handle();
        H->device = device;
        make_plan(H, desc, external_leaf_size);
        EF_CUDA(cudaStreamCreateWithFlags(&H->stream, cudaStreamNonBlocking));
        EF_CUDA(cudaEventCreate(&H->ev0)); EF_CUDA(cudaEventCreate(&H->ev1));
        { const char* ge = getenv("EFGPU_GRAPHS"); H->graphs_on = !ge || atoi(ge) != 0; }
        { const char* ge = getenv("EFGPU_GRAPHS_PEER"); H->graphs_peer = !ge || atoi(ge) != 0; }
        if (const char* te = getenv("EFGPU_TRACE")) {
            static int serial = 0;
            char nm[512]; snprintf(nm, sizeof nm, "%s.pid%d.h%d", te, (int)getpid(), serial++);
            H->trace = fopen(nm, "w");
        }
        *out = H;
        return EF_OK;
    } catch (const efgpu::Error& e) { g_create_error = e.msg; delete H; return e.code; }
      catch (const std::exception& e) { g_create_error = e.what(); delete H; return EF_ERR_STATE; }
}

void efgpu_destroy(efgpu_handle* H)
{
    if (!H) return;
    cudaSetDevice(H->device);
    // a borrowed stream (efgpu_set_stream) may already have been destroyed by its owner when handles are released in arbitrary
    // order (garbage-collected callers): synchronise the device instead of touching it
    if (H->stream && H->own_stream) cudaStreamSynchronize(H->stream); else cudaDeviceSynchronize();
    if (H->trace) fclose(H->trace);
    if (H->ev0) cudaEventDestroy(H->ev0);
    if (H->ev1) cudaEventDestroy(H->ev1);
    for (auto& r : H->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : H->ev_pool) cudaEventDestroy(e);
    for (void* p : H->peer_mapped) if (p) cudaIpcCloseMemHandle(p);
    for (efgpu_handle::GraphSlot* g : {&H->g_build, &H->g_up, &H->g_solve}) if (g->exec) cudaGraphExecDestroy(g->exec);
    for (int k = 1; k < efgpu_handle::MAX_LANES; k++) { if (H->lane[k]) cudaStreamDestroy(H->lane[k]); if (H->ev_join[k]) cudaEventDestroy(H->ev_join[k]); }
    if (H->ev_fork) cudaEventDestroy(H->ev_fork);
    cudaStream_t s = H->own_stream ? H->stream : nullptr;
    delete H;
    if (s) cudaStreamDestroy(s);
}

const char* efgpu_last_error(const efgpu_handle* H) { return H ? H->last_error.c_str() : g_create_error.c_str(); }

int efgpu_set_leaf_constant(efgpu_handle* H, double lambda)
{
    if (!H) return EF_ERR_BAD_ARG;
    H->leaf_kind = EFGPU_LEAF_CONSTANT; H->lambda = lambda; H->built = false;
    return EF_OK;
}

// largest sampled lambda of variable-coefficient leaves (> 0: indefinite operator, the build refines every X^-1)
static void sampled_lambda_max(efgpu_handle* H)
{
    H->d_err.alloc(sizeof(double));
    launch_max_positive(H->d_coef_in[5].as<double>(), (size_t)H->n_leaves * H->M * H->M, H->d_err.as<double>(), H->stream);
    EF_CUDA(cudaMemcpyAsync(&H->lambda_max, H->d_err.p, sizeof(double), cudaMemcpyDeviceToHost, H->stream));
    EF_CUDA(cudaStreamSynchronize(H->stream));
}

int efgpu_set_leaf_variable(efgpu_handle* H, const double* alpha, const double* beta_w, const double* beta_e, const double* beta_s,
                            const double* beta_n, const double* lambda)
{
    if (!H || !alpha || !beta_w || !beta_e || !beta_s || !beta_n || !lambda) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    if (H->external_leaves) throw Error{EF_ERR_STATE, "external leaves have no coefficients"};
    EF_CUDA(cudaSetDevice(H->device));
    const double* src[6] = {alpha, beta_w, beta_e, beta_s, beta_n, lambda};
    const size_t bytes = (size_t)H->n_leaves * H->M * H->M * sizeof(double);
    for (int k = 0; k < 6; k++) {
        H->d_coef_in[k].alloc(bytes);
        EF_CUDA(cudaMemcpyAsync(H->d_coef_in[k].p, src[k], bytes, cudaMemcpyHostToDevice, H->stream));
    }
    EF_CUDA(cudaStreamSynchronize(H->stream));
    H->leaf_kind = EFGPU_LEAF_VARIABLE; H->built = false;
    sampled_lambda_max(H);
    EF_CATCH(H)
}

int efgpu_set_leaf_variable_device(efgpu_handle* H, const double* alpha, const double* beta_w, const double* beta_e, const double* beta_s,
                                   const double* beta_n, const double* lambda)
{
    if (!H || !alpha || !beta_w || !beta_e || !beta_s || !beta_n || !lambda) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    if (H->external_leaves) throw Error{EF_ERR_STATE, "external leaves have no coefficients"};
    EF_CUDA(cudaSetDevice(H->device));
    const double* src[6] = {alpha, beta_w, beta_e, beta_s, beta_n, lambda};
    const size_t bytes = (size_t)H->n_leaves * H->M * H->M * sizeof(double);
    for (int k = 0; k < 6; k++) {
        H->d_coef_in[k].alloc(bytes);
        EF_CUDA(cudaMemcpyAsync(H->d_coef_in[k].p, src[k], bytes, cudaMemcpyDeviceToDevice, H->stream));
    }
    EF_CUDA(cudaStreamSynchronize(H->stream));   // the caller's arrays are borrowed for the call only
    H->leaf_kind = EFGPU_LEAF_VARIABLE; H->built = false;
    sampled_lambda_max(H);
    EF_CATCH(H)
}

int efgpu_build(efgpu_handle* H, unsigned flags)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    do_build(H, flags);
    EF_CATCH(H)
}

int efgpu_rebuild_from(efgpu_handle* H, efgpu_handle* old, unsigned flags, double* reused_merges, double* rebuilt_merges)
{
    if (!H || !old || H == old) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    do_rebuild_from(H, old, flags, reused_merges, rebuilt_merges);
    EF_CATCH(H)
}

int efgpu_set_partition(efgpu_handle* H, int rank, int nranks)
{
    if (!H || nranks < 1 || rank < 0 || rank >= nranks) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    if (H->allocated) throw Error{EF_ERR_STATE, "efgpu_set_partition must precede the first build / device view"};
    H->part_rank = rank; H->part_nranks = nranks; H->graph_gen++;
    for (auto& b : H->batches) plan_batch_gemms(b, rank, nranks);
    compute_flop_model(H);
    EF_CATCH(H)
}

int efgpu_peer_export(efgpu_handle* H, void* ipc_handle_out)
{
    if (!H || !ipc_handle_out) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (H->part_nranks > PEER_MAX) throw Error{EF_ERR_UNSUPPORTED, "peer mode: at most 8 ranks (one NVSwitch domain)"};
    if (H->allocated && !H->peer_mode) throw Error{EF_ERR_STATE, "efgpu_peer_export must precede the first build / device view"};
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->peer_mode) {
        H->peer_mode = true;
        for (auto& b : H->batches) plan_batch_gemms(b, H->part_rank, H->part_nranks, true);
        compute_flop_model(H);
        H->d_peer_err.alloc(sizeof(int));
        EF_CUDA(cudaMemsetAsync(H->d_peer_err.p, 0, sizeof(int), H->stream));
        allocate_device(H, H->build_flags);
        EF_CUDA(cudaStreamSynchronize(H->stream));   // the flag page is zero before any peer can map it
    }
    cudaIpcMemHandle_t hd;
    EF_CUDA(cudaIpcGetMemHandle(&hd, H->arena.p));
    std::memcpy(ipc_handle_out, &hd, sizeof(hd));
    EF_CATCH(H)
}

int efgpu_peer_attach(efgpu_handle* H, const void* handles, int nranks)
{
    if (!H || !handles) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    if (!H->peer_mode) throw Error{EF_ERR_STATE, "efgpu_peer_attach before efgpu_peer_export"};
    if (nranks != H->part_nranks) throw Error{EF_ERR_BAD_ARG, "efgpu_peer_attach: one handle per rank of the partition"};
    EF_CUDA(cudaSetDevice(H->device));
    H->peers = PeerSpan{};
    H->peers.n = nranks; H->peers.me = H->part_rank; H->peers.local_base = static_cast<char*>(H->arena.p);
    for (int r = 0; r < nranks; r++) {
        if (r == H->part_rank) { H->peers.delta[r] = 0; continue; }
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, static_cast<const char*>(handles) + 64 * (size_t)r, sizeof(hd));
        void* p = nullptr;
        EF_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        H->peer_mapped[r] = p;
        H->peers.delta[r] = static_cast<char*>(p) - static_cast<char*>(H->arena.p);
    }
    H->peer_attached = true;
    EF_CATCH(H)
}

int efgpu_peer_barrier(efgpu_handle* H)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    if (!H->peer_attached) throw Error{EF_ERR_STATE, "efgpu_peer_barrier on a handle without attached peers"};
    EF_CUDA(cudaSetDevice(H->device));
    peer_barrier(H, false);
    EF_CATCH(H)
}

int efgpu_peer_broadcast(efgpu_handle* H, const void* dev_ptr, size_t bytes)
{
    if (!H || !dev_ptr) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    if (!H->peer_attached) throw Error{EF_ERR_STATE, "efgpu_peer_broadcast on a handle without attached peers"};
    const char* p = static_cast<const char*>(dev_ptr);
    const char* base = static_cast<const char*>(H->arena.p);
    if (p < base + 4096 || p + bytes > base + H->arena_need) throw Error{EF_ERR_BAD_ARG, "efgpu_peer_broadcast: the region must lie inside the shared arena (a device view of this handle)"};
    EF_CUDA(cudaSetDevice(H->device));
    timed(H, EFGPU_PROF_ALLGATHER, 1, [&] { launch_peer_scatter(H->peers, H->part_rank, (size_t)(p - base), bytes, H->stream); });
    H->peer_dirty = true;
    EF_CATCH(H)
}

int efgpu_set_allgather(efgpu_handle* H, efgpu_allgather_fn fn, void* user)
{
    if (!H) return EF_ERR_BAD_ARG;
    H->allgather = fn; H->allgather_user = user;
    return EF_OK;
}

int efgpu_debug_merge_plan(int n, int level, int rank, int nranks, int symmetric, int64_t* steps, int* n_steps,
                           int64_t* blocks, int64_t* terms, int* n_blocks, int64_t* trans, int* n_trans, int64_t* ws)
{
    return efgpu_debug_merge_plan_ex(n, level, rank, nranks, symmetric, 0, steps, n_steps, blocks, terms, n_blocks, trans, n_trans, ws);
}

int efgpu_debug_merge_plan_ex(int n, int level, int rank, int nranks, int symmetric, int peer, int64_t* steps, int* n_steps,
                              int64_t* blocks, int64_t* terms, int* n_blocks, int64_t* trans, int* n_trans, int64_t* ws)
{
    if (n < 8 || n % 8 || nranks < 1 || rank < 0 || rank >= nranks || !n_steps || !n_blocks || !n_trans) return EF_ERR_BAD_ARG;
    try {
        BatchH b; b.n = n; b.level = level; b.count = 1; b.symcand = symmetric != 0;
        plan_batch_gemms(b, rank, nranks, peer != 0);
        const std::vector<Step>& st = symmetric ? b.steps_sym : b.steps;
        *n_steps = (int)st.size(); *n_blocks = (int)b.blocks.size(); *n_trans = (int)b.trans.size();
        if (ws) { ws[0] = (int64_t)b.w2_off; ws[1] = (int64_t)(b.w3_off - b.w2_off); ws[2] = (int64_t)(b.ws_per_entry - b.w3_off); }
        if (steps) for (size_t i = 0; i < st.size(); i++) {
            int64_t* r = steps + 16 * i; const Step& x = st[i];
            r[0] = x.kind; r[1] = x.first; r[2] = x.count; r[3] = x.off; r[4] = x.N; r[5] = x.cls; r[6] = x.gk; r[7] = x.g_op;
            r[8] = x.g_rows; r[9] = x.g_cols; r[10] = x.g_ld; r[11] = x.g_off; r[12] = x.off2; r[13] = b.colsplit ? 1 : 0;
        }
        if (blocks && terms) for (size_t i = 0; i < b.blocks.size(); i++) {
            int64_t* r = blocks + 16 * i; const GemmBlock& g = b.blocks[i];
            r[0] = g.c_op; r[1] = g.c0_op; r[2] = g.ldc; r[3] = g.ldc0; r[4] = g.c_off; r[5] = g.c0_off; r[6] = g.rows; r[7] = g.cols; r[8] = g.nterms;
            r[9] = g.ct_op1; r[10] = g.ldct; r[11] = g.ct_off; r[12] = g.ct_neg ? 1 : 0;
            for (int t = 0; t < 2; t++) {
                int64_t* q = terms + 16 * i + 8 * t; const GemmTerm& m = g.t[t];
                q[0] = m.a_op; q[1] = m.b_op; q[2] = m.lda; q[3] = m.ldb; q[4] = m.a_off; q[5] = m.b_off; q[6] = m.K; q[7] = m.neg ? 1 : 0;
            }
        }
        if (trans) for (size_t i = 0; i < b.trans.size(); i++) {
            int64_t* r = trans + 16 * i; const TransOp& x = b.trans[i];
            r[0] = x.src_op; r[1] = x.dst_op; r[2] = x.lds; r[3] = x.ldd; r[4] = x.src_off; r[5] = x.dst_off; r[6] = x.rows; r[7] = x.cols; r[8] = x.neg ? 1 : 0;
        }
        return EF_OK;
    } catch (const efgpu::Error& e) { g_create_error = e.msg; return e.code; }
}

int efgpu_debug_tma_plan(int n, int level, int rank, int nranks, int symmetric, int peer, int64_t* views, int* n_views,
                         int64_t* tblocks, int64_t* step_tma)
{
    if (n < 8 || n % 8 || nranks < 1 || rank < 0 || rank >= nranks || !n_views) return EF_ERR_BAD_ARG;
    try {
        BatchH b; b.n = n; b.level = level; b.count = 1; b.symcand = symmetric != 0;
        plan_batch_gemms(b, rank, nranks, peer != 0);
        const std::vector<Step>& st = symmetric ? b.steps_sym : b.steps;
        *n_views = (int)b.views.size();
        if (views) for (size_t v = 0; v < b.views.size(); v++) { views[3 * v] = b.views[v].op; views[3 * v + 1] = b.views[v].origin; views[3 * v + 2] = b.views[v].ld; }
        if (tblocks) for (size_t k = 0; k < b.tblocks.size(); k++)
            for (int t = 0; t < 2; t++) {
                int64_t* q = tblocks + 12 * k + 6 * t; const TmaTerm& m = b.tblocks[k].t[t];
                q[0] = m.a_view; q[1] = m.b_view; q[2] = m.a_row; q[3] = m.a_col; q[4] = m.b_row; q[5] = m.b_col;
            }
        if (step_tma) for (size_t i = 0; i < st.size(); i++) step_tma[i] = st[i].tma ? 1 : 0;
        return EF_OK;
    } catch (const efgpu::Error& e) { g_create_error = e.msg; return e.code; }
}

int efgpu_set_tuning(int key, int value)
{
    if (key < 0 || key >= 16) return EF_ERR_BAD_ARG;
    efgpu::set_tuning(key, value);
    return EF_OK;
}

int efgpu_set_symmetric_leaves(efgpu_handle* H, int on)
{
    if (!H) return EF_ERR_BAD_ARG;
    H->ext_sym = on != 0; H->built = false;
    return EF_OK;
}

int efgpu_set_refine_inverse(efgpu_handle* H, int mode)
{
    if (!H || mode < -1 || mode > 1) return EF_ERR_BAD_ARG;
    H->refine_mode = mode; H->built = false;
    return EF_OK;
}

int efgpu_is_symmetric(const efgpu_handle* H)
{
    if (!H || !H->built) return 0;
    for (int r : H->roots) {
        const efgpu::NodeH& nd = H->nodes[r];
        if (nd.leaf) { if (!(nd.symcand && (H->external_leaves ? H->ext_sym : H->leaf_kind == EFGPU_LEAF_CONSTANT))) return 0; }
        else if (!H->batches[nd.batch].use_sym) return 0;
    }
    return 1;
}

int efgpu_build_begin(efgpu_handle* H, unsigned flags)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    build_begin(H, flags);
    build_leaves(H);
    EF_CATCH(H)
}

int efgpu_build_level(efgpu_handle* H, int level, int phase)
{
    if (!H || (phase != 0 && phase != 1)) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->allocated) throw Error{EF_ERR_STATE, "efgpu_build_level before efgpu_build_begin"};
    build_level(H, level, phase);
    EF_CATCH(H)
}

int efgpu_build_end(efgpu_handle* H)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    build_end(H);
    EF_CATCH(H)
}

int efgpu_complete_root_dtn(efgpu_handle* H)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    efgpu::complete_root_T(H);
    EF_CATCH(H)
}

int efgpu_max_level(const efgpu_handle* H) { return H ? H->max_level : -1; }

int efgpu_upwards(efgpu_handle* H, const double* f_leaves, double fscale, unsigned flags)
{
    if (!H || !f_leaves) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->built) throw Error{EF_ERR_STATE, "upwards before build"};
    EF_CUDA(cudaMemcpyAsync(H->d_f.p, f_leaves, (size_t)H->n_leaves * H->M * H->M * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    H->f_cur = H->d_f.as<double>(); H->fscale_cur = fscale;
    do_upwards(H, H->f_cur, fscale, flags);
    EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
    float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1)); H->stats.upwards_ms = ms;
    EF_CATCH(H)
}

int efgpu_upwards_device(efgpu_handle* H, const double* f_leaves_dev, double fscale, unsigned flags, int sync)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!f_leaves_dev && !H->external_leaves) throw Error{EF_ERR_BAD_ARG, "null load vector"};
    H->f_cur = f_leaves_dev; H->fscale_cur = fscale;   // borrowed until the next upwards call
    do_upwards(H, H->f_cur, fscale, flags);
    if (sync) {
        EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
        float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1)); H->stats.upwards_ms = ms;
    }
    EF_CATCH(H)
}

static void set_root_g(efgpu_handle* H, const double* g_root, cudaMemcpyKind kind)
{
    if (H->roots.size() != 1) throw Error{EF_ERR_STATE, "this handle holds a forest: set each root's g through efgpu_vector_device and call efgpu_solve_from_roots_device"};
    NodeH& root = H->nodes[0];
    EF_CUDA(cudaMemcpyAsync(H->d_vec.as<double>() + root.gbuf[0], g_root, 4 * (size_t)root.size * sizeof(double), kind, H->stream));
}

int efgpu_solve_dirichlet(efgpu_handle* H, const double* g_root, unsigned flags, double* u_leaves)
{
    if (!H || !g_root) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->built) throw Error{EF_ERR_STATE, "solve before build"};
    if (!(flags & EFGPU_HOMOGENEOUS_RHS) && !H->upwards_done) throw Error{EF_ERR_STATE, "solve before upwards (non-homogeneous right-hand side)"};
    EF_CUDA(cudaEventRecord(H->ev0, H->stream));
    set_root_g(H, g_root, cudaMemcpyHostToDevice);
    do_solve(H, H->f_cur, H->fscale_cur, flags);
    EF_CUDA(cudaEventRecord(H->ev1, H->stream));
    if (u_leaves) EF_CUDA(cudaMemcpyAsync(u_leaves, H->d_u.p, (size_t)H->n_leaves * H->M * H->M * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
    EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
    float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1)); H->stats.solve_ms = ms;
    EF_CATCH(H)
}

int efgpu_solve_dirichlet_device(efgpu_handle* H, const double* g_root_dev, unsigned flags, double* u_leaves_dev, int sync)
{
    if (!H || !g_root_dev) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->built) throw Error{EF_ERR_STATE, "solve before build"};
    if (!(flags & EFGPU_HOMOGENEOUS_RHS) && !H->upwards_done) throw Error{EF_ERR_STATE, "solve before upwards (non-homogeneous right-hand side)"};
    EF_CUDA(cudaEventRecord(H->ev0, H->stream));
    set_root_g(H, g_root_dev, cudaMemcpyDeviceToDevice);
    do_solve(H, H->f_cur, H->fscale_cur, flags);
    if (u_leaves_dev) EF_CUDA(cudaMemcpyAsync(u_leaves_dev, H->d_u.p, (size_t)H->n_leaves * H->M * H->M * sizeof(double), cudaMemcpyDeviceToDevice, H->stream));
    EF_CUDA(cudaEventRecord(H->ev1, H->stream));
    if (sync) {
        EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
        float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1)); H->stats.solve_ms = ms;
    }
    EF_CATCH(H)
}

int efgpu_solve_from_roots_device(efgpu_handle* H, unsigned flags, double* u_leaves_dev, int sync)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->built) throw Error{EF_ERR_STATE, "solve before build"};
    if (!(flags & EFGPU_HOMOGENEOUS_RHS) && !H->upwards_done) throw Error{EF_ERR_STATE, "solve before upwards (non-homogeneous right-hand side)"};
    EF_CUDA(cudaEventRecord(H->ev0, H->stream));
    do_solve(H, H->f_cur, H->fscale_cur, flags);
    if (u_leaves_dev && !H->external_leaves)
        EF_CUDA(cudaMemcpyAsync(u_leaves_dev, H->d_u.p, (size_t)H->n_leaves * H->M * H->M * sizeof(double), cudaMemcpyDeviceToDevice, H->stream));
    EF_CUDA(cudaEventRecord(H->ev1, H->stream));
    if (sync) {
        EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
        float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1)); H->stats.solve_ms = ms;
    }
    EF_CATCH(H)
}

int efgpu_operator_device(efgpu_handle* H, int node, int which, double** ptr, int* rows, int* cols)
{
    if (!H || !ptr) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    int r = 0, c = 0;
    if (efgpu_operator_shape(H, node, which, &r, &c) != EF_OK) throw Error{EF_ERR_BAD_ARG, "bad node / operator"};
    if (!H->allocated) { EF_CUDA(cudaSetDevice(H->device)); allocate_device(H, H->build_flags); }
    const NodeH& nd = H->nodes[node];
    const size_t n = nd.size / 2;
    if (which == EFGPU_OP_T || which == EFGPU_OP_T_UNCOARSENED) require_T_retained(H, node);
    switch (which) {
        case EFGPU_OP_T: *ptr = nd.Tbuf[nd.ncoarsen]; break;
        case EFGPU_OP_T_UNCOARSENED: *ptr = nd.Tbuf[0]; break;
        case EFGPU_OP_S: *ptr = H->batches[nd.batch].S.as<double>() + nd.slot * 32 * n * n; break;
        case EFGPU_OP_XINV: *ptr = H->batches[nd.batch].Xinv.as<double>() + nd.slot * 16 * n * n; break;
        default: throw Error{EF_ERR_UNSUPPORTED, "no device view of this operator (X and H are not stored densely)"};
    }
    if (rows) *rows = r;
    if (cols) *cols = c;
    EF_CATCH(H)
}

int efgpu_vector_device(efgpu_handle* H, int node, int which, double** ptr, int* len)
{
    if (!H || !ptr) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    int l = 0;
    if (efgpu_vector_length(H, node, which, &l) != EF_OK) throw Error{EF_ERR_BAD_ARG, "bad node / vector"};
    if (!H->allocated) { EF_CUDA(cudaSetDevice(H->device)); allocate_device(H, H->build_flags); }
    const NodeH& nd = H->nodes[node];
    double* vec = H->d_vec.as<double>();
    switch (which) {
        case EFGPU_VEC_H: *ptr = vec + nd.hbuf[nd.ncoarsen]; break;
        case EFGPU_VEC_H_UNCOARSENED: *ptr = vec + nd.hbuf[0]; l = 4 * nd.size; break;
        case EFGPU_VEC_G: *ptr = vec + nd.gbuf[0]; break;
        case EFGPU_VEC_W: *ptr = vec + nd.w_off; break;
        case EFGPU_VEC_U: if (H->external_leaves) throw Error{EF_ERR_BAD_ARG, "external leaves have no u"}; *ptr = H->d_u.as<double>() + (size_t)nd.leaf_idx * l; break;
        case EFGPU_VEC_F: if (H->external_leaves) throw Error{EF_ERR_BAD_ARG, "external leaves have no f"}; *ptr = H->d_f.as<double>() + (size_t)nd.leaf_idx * l; break;
        default: throw Error{EF_ERR_BAD_ARG, "bad vector selector"};
    }
    if (len) *len = l;
    EF_CATCH(H)
}

int efgpu_solve_robin(efgpu_handle* H, const double* a, const double* b, const double* r, unsigned flags, double* u_leaves)
{
    if (!H || !a || !b || !r) return EF_ERR_BAD_ARG;
    if (H->roots.size() != 1) { H->last_error = "root boundary solve on a forest handle"; return EF_ERR_STATE; }
    const int len = 4 * H->nodes[0].size;
    bool dirichlet = true;
    for (int i = 0; i < len; i++) if (b[i] != 0.0) { dirichlet = false; break; }
    if (dirichlet) {
        std::vector<double> g(len);
        for (int i = 0; i < len; i++) g[i] = r[i] / a[i];   // g = (diag a)^-1 r   (HPSAlgorithm.hpp:402-419 with b = 0)
        return efgpu_solve_dirichlet(H, g.data(), flags, u_leaves);
    }
    // general case: dense pivoted LU of diag(a) + diag(b) T_root on the device (HPSAlgorithm.hpp:408-419)
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    if (!H->built) throw Error{EF_ERR_STATE, "solve before build"};
    const bool homogeneous = (flags & EFGPU_HOMOGENEOUS_RHS) != 0;
    if (!homogeneous && !H->upwards_done) throw Error{EF_ERR_STATE, "solve before upwards (non-homogeneous right-hand side)"};
    if (H->nodes[0].leaf && H->external_leaves) throw Error{EF_ERR_STATE, "no root operator"};
    if (!H->nodes[0].leaf) require_T_retained(H, 0);   // forms a lazily deferred root map (EFGPU_LAZY_ROOT_DTN); throws while it is row-distributed
    cudaStream_t s = H->stream;
    H->d_robin.alloc((robin_workspace_doubles(len) + 4 * (size_t)len) * sizeof(double));
    double* abr = H->d_robin.as<double>();
    double* ws = abr + 4 * (size_t)len;
    EF_CUDA(cudaMemcpyAsync(abr, a, len * sizeof(double), cudaMemcpyHostToDevice, s));
    EF_CUDA(cudaMemcpyAsync(abr + len, b, len * sizeof(double), cudaMemcpyHostToDevice, s));
    EF_CUDA(cudaMemcpyAsync(abr + 2 * (size_t)len, r, len * sizeof(double), cudaMemcpyHostToDevice, s));
    NodeH& root = H->nodes[0];
    const double* hroot = homogeneous ? nullptr : H->d_vec.as<double>() + root.hbuf[0];   // the reference reads an empty vectorH here and throws
    int info = 0;
    EF_CUDA(cudaEventRecord(H->ev0, s));
    robin_solve(root.Tbuf[0], abr, abr + len, abr + 2 * (size_t)len, hroot, len, ws, abr + 3 * (size_t)len, &info, s);
    set_root_g(H, abr + 3 * (size_t)len, cudaMemcpyDeviceToDevice);
    do_solve(H, H->f_cur, H->fscale_cur, flags);
    EF_CUDA(cudaEventRecord(H->ev1, s));
    if (u_leaves) EF_CUDA(cudaMemcpyAsync(u_leaves, H->d_u.p, (size_t)H->n_leaves * H->M * H->M * sizeof(double), cudaMemcpyDeviceToHost, s));
    EF_CUDA(cudaStreamSynchronize(s)); collect_profile(H);
    float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, H->ev0, H->ev1)); H->stats.solve_ms = ms;
    if (info != 0) throw Error{EF_ERR_SINGULAR, "root boundary system is singular (zero pivot at column " + std::to_string(info) + ")"};
    EF_CATCH(H)
}

// ---- sampling coordinates and error norms on the device (sample.cu) ----------
static void points_tables(efgpu_handle* H)
{
    if (H->external_leaves) throw Error{EF_ERR_STATE, "external leaves have no cells"};
    EF_CUDA(cudaSetDevice(H->device));
    if (H->d_boxes.p && H->d_leaf_nodes.p) return;
    // before the first build only the two small tables are needed (allocate_device uploads the same contents again)
    std::vector<double> boxes((size_t)4 * H->n_nodes);
    for (int i = 0; i < H->n_nodes; i++) std::memcpy(&boxes[4 * (size_t)i], H->nodes[i].box, 4 * sizeof(double));
    H->d_boxes.upload(boxes, H->stream);
    H->d_leaf_nodes.upload(H->leaf_nodes, H->stream);
    EF_CUDA(cudaStreamSynchronize(H->stream));   // the host vectors go out of scope
}

int efgpu_leaf_points_device(efgpu_handle* H, int which, double* x_dev, double* y_dev, int sync)
{
    if (!H || which < 0 || which > 4 || (!x_dev && !y_dev)) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    points_tables(H);
    launch_leaf_points(H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->M, which, x_dev, y_dev, H->n_leaves, H->stream);
    EF_CUDA(cudaGetLastError());
    if (sync) EF_CUDA(cudaStreamSynchronize(H->stream));
    EF_CATCH(H)
}

int efgpu_leaf_points(efgpu_handle* H, int which, double* x, double* y)
{
    if (!H || which < 0 || which > 4 || (!x && !y)) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    points_tables(H);
    const size_t cells = (size_t)H->n_leaves * H->M * H->M;
    H->d_pts.alloc(2 * cells * sizeof(double));
    double* dx = H->d_pts.as<double>();
    double* dy = dx + cells;
    launch_leaf_points(H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->M, which, x ? dx : nullptr, y ? dy : nullptr, H->n_leaves, H->stream);
    EF_CUDA(cudaGetLastError());
    if (x) EF_CUDA(cudaMemcpyAsync(x, dx, cells * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
    if (y) EF_CUDA(cudaMemcpyAsync(y, dy, cells * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
    EF_CUDA(cudaStreamSynchronize(H->stream));
    H->d_pts.release();   // 16 B per cell: not kept between calls
    EF_CATCH(H)
}

static void error_norms(efgpu_handle* H, const double* u_dev, const double* exact_dev, double* l1, double* l2, double* linf)
{
    if (!u_dev) {
        if (!H->solve_done) throw Error{EF_ERR_STATE, "error norms of the handle's solution before a solve stage"};
        u_dev = H->d_u.as<double>();
    }
    double area = 0.0;   // (x_upper - x_lower) * (y_upper - y_lower) of the domain (main.cpp:368); a forest: sum over its roots
    for (int r : H->roots) area += (H->nodes[r].box[1] - H->nodes[r].box[0]) * (H->nodes[r].box[3] - H->nodes[r].box[2]);
    H->d_err.alloc((3 * (size_t)H->n_leaves + 3) * sizeof(double));
    double* part = H->d_err.as<double>();
    double* out = part + 3 * (size_t)H->n_leaves;
    launch_error_norms(u_dev, exact_dev, H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->M, H->n_leaves, area, part, out, H->stream);
    EF_CUDA(cudaGetLastError());
    double res[3];
    EF_CUDA(cudaMemcpyAsync(res, out, sizeof(res), cudaMemcpyDeviceToHost, H->stream));
    EF_CUDA(cudaStreamSynchronize(H->stream));
    if (l1) *l1 = res[0];
    if (l2) *l2 = res[1];
    if (linf) *linf = res[2];
}

int efgpu_error_norms_device(efgpu_handle* H, const double* u_dev, const double* exact_dev, double* l1, double* l2, double* linf)
{
    if (!H || !exact_dev) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    points_tables(H);
    error_norms(H, u_dev, exact_dev, l1, l2, linf);
    EF_CATCH(H)
}

int efgpu_error_norms(efgpu_handle* H, const double* exact, double* l1, double* l2, double* linf)
{
    if (!H || !exact) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    points_tables(H);
    const size_t bytes = (size_t)H->n_leaves * H->M * H->M * sizeof(double);
    H->d_pts.alloc(bytes);
    EF_CUDA(cudaMemcpyAsync(H->d_pts.p, exact, bytes, cudaMemcpyHostToDevice, H->stream));
    error_norms(H, nullptr, H->d_pts.as<double>(), l1, l2, linf);
    H->d_pts.release();
    EF_CATCH(H)
}

int efgpu_write_vtu(efgpu_handle* H, const char* path, int n_fields, const char* const* names, const double* const* fields_dev)
{
    if (!H || !path || n_fields < 0 || (n_fields > 0 && (!names || !fields_dev))) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    points_tables(H);
    std::vector<const double*> f(n_fields);
    for (int k = 0; k < n_fields; k++) {
        if (!names[k]) throw Error{EF_ERR_BAD_ARG, "efgpu_write_vtu: null field name"};
        f[k] = fields_dev[k];
        if (!f[k]) {   // the handle's own solution
            if (!H->solve_done) throw Error{EF_ERR_STATE, "efgpu_write_vtu: no solution yet (null field pointer = the last solve stage's u)"};
            f[k] = H->d_u.as<double>();
        }
    }
    EF_CUDA(cudaStreamSynchronize(H->stream));
    write_vtu(path, H->d_boxes.as<double>(), H->d_leaf_nodes.as<int>(), H->M, H->n_leaves, n_fields, names, f.data(), H->stream);
    EF_CATCH(H)
}

int efgpu_sync(efgpu_handle* H)
{
    if (!H) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
    EF_CATCH(H)
}

void* efgpu_stream(efgpu_handle* H) { return H ? (void*)H->stream : nullptr; }

int efgpu_set_stream(efgpu_handle* H, void* stream)
{
    if (!H || !stream) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    EF_CUDA(cudaSetDevice(H->device));
    EF_CUDA(cudaStreamSynchronize(H->stream));
    if (H->own_stream) EF_CUDA(cudaStreamDestroy(H->stream));
    H->stream = (cudaStream_t)stream; H->own_stream = false;
    EF_CATCH(H)
}

int efgpu_node_info(const efgpu_handle* H, int node, int* size, int* n_coarsens, int* is_leaf, int* leaf_index)
{
    if (!H || node < 0 || node >= H->n_nodes) return EF_ERR_BAD_ARG;
    const NodeH& nd = H->nodes[node];
    if (size) *size = nd.size;
    if (n_coarsens) *n_coarsens = nd.ncoarsen;
    if (is_leaf) *is_leaf = nd.leaf ? 1 : 0;
    if (leaf_index) *leaf_index = nd.leaf_idx;
    return EF_OK;
}

int efgpu_operator_shape(const efgpu_handle* H, int node, int which, int* rows, int* cols)
{
    if (!H || node < 0 || node >= H->n_nodes || !rows || !cols) return EF_ERR_BAD_ARG;
    const NodeH& nd = H->nodes[node];
    const int n = nd.size / 2;
    switch (which) {
        case EFGPU_OP_T: *rows = *cols = 4 * (nd.size >> nd.ncoarsen); return EF_OK;   // coarsened in place by the parent's merge (quirk q3)
        case EFGPU_OP_T_UNCOARSENED: *rows = *cols = 4 * nd.size; return EF_OK;
        case EFGPU_OP_S: if (nd.leaf) return EF_ERR_BAD_ARG; *rows = 4 * n; *cols = 8 * n; return EF_OK;
        case EFGPU_OP_X: case EFGPU_OP_XINV: if (nd.leaf) return EF_ERR_BAD_ARG; *rows = *cols = 4 * n; return EF_OK;
        case EFGPU_OP_H: if (nd.leaf) return EF_ERR_BAD_ARG; *rows = 8 * n; *cols = 4 * n; return EF_OK;
        default: return EF_ERR_BAD_ARG;
    }
}

int efgpu_get_operator(efgpu_handle* H, int node, int which, double* out, size_t capacity)
{
    if (!H || !out) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    int rows = 0, cols = 0;
    if (efgpu_operator_shape(H, node, which, &rows, &cols) != EF_OK) throw Error{EF_ERR_BAD_ARG, "bad node / operator"};
    if (!H->built) throw Error{EF_ERR_STATE, "operators requested before build"};
    if (capacity < (size_t)rows * cols) throw Error{EF_ERR_BAD_SHAPE, "output buffer too small"};
    EF_CUDA(cudaSetDevice(H->device));
    const NodeH& nd = H->nodes[node];
    const size_t n = nd.size / 2, bytes = (size_t)rows * cols * sizeof(double);
    const double* src = nullptr;
    DevBuf tmp;
    if (which == EFGPU_OP_T || which == EFGPU_OP_T_UNCOARSENED) require_T_retained(H, node);
    switch (which) {
        case EFGPU_OP_T: src = nd.Tbuf[nd.ncoarsen]; break;
        case EFGPU_OP_T_UNCOARSENED: src = nd.Tbuf[0]; break;
        case EFGPU_OP_S: src = H->batches[nd.batch].S.as<double>() + nd.slot * 32 * n * n; break;
        case EFGPU_OP_XINV: src = H->batches[nd.batch].Xinv.as<double>() + nd.slot * 16 * n * n; break;
        case EFGPU_OP_X:
            if (!H->batches[nd.batch].Xcopy.p) throw Error{EF_ERR_STATE, "X is only retained when built with EFGPU_KEEP_X"};
            src = H->batches[nd.batch].Xcopy.as<double>() + nd.slot * 16 * n * n; break;
        case EFGPU_OP_H:
            tmp.alloc(bytes);
            launch_expand_H(H->batches[nd.batch].Hc.as<double>() + nd.slot * 16 * n * n, (int)n, tmp.as<double>(), H->stream);
            src = tmp.as<double>(); break;
    }
    EF_CUDA(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, H->stream));
    EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
    EF_CATCH(H)
}

int efgpu_vector_length(const efgpu_handle* H, int node, int which, int* len)
{
    if (!H || node < 0 || node >= H->n_nodes || !len) return EF_ERR_BAD_ARG;
    const NodeH& nd = H->nodes[node];
    switch (which) {
        case EFGPU_VEC_H: *len = 4 * (nd.size >> nd.ncoarsen); return EF_OK;   // coarsened in place by coarsenUpwards_
        case EFGPU_VEC_H_UNCOARSENED: *len = 4 * nd.size; return EF_OK;
        case EFGPU_VEC_G: *len = 4 * nd.size; return EF_OK;                     // uncoarsened in place by uncoarsen_
        case EFGPU_VEC_W: if (nd.leaf) return EF_ERR_BAD_ARG; *len = 2 * nd.size; return EF_OK;
        case EFGPU_VEC_U: case EFGPU_VEC_F: if (!nd.leaf) return EF_ERR_BAD_ARG; *len = nd.size * nd.size; return EF_OK;
        default: return EF_ERR_BAD_ARG;
    }
}

int efgpu_get_vector(efgpu_handle* H, int node, int which, double* out, size_t capacity)
{
    if (!H || !out) return EF_ERR_BAD_ARG;
    EF_TRY(H)
    int len = 0;
    if (efgpu_vector_length(H, node, which, &len) != EF_OK) throw Error{EF_ERR_BAD_ARG, "bad node / vector"};
    if (capacity < (size_t)len) throw Error{EF_ERR_BAD_SHAPE, "output buffer too small"};
    if (!H->allocated) throw Error{EF_ERR_STATE, "vectors requested before build"};
    EF_CUDA(cudaSetDevice(H->device));
    const NodeH& nd = H->nodes[node];
    const double* vec = H->d_vec.as<double>();
    const double* src = nullptr;
    switch (which) {
        case EFGPU_VEC_H: src = vec + nd.hbuf[nd.ncoarsen]; break;
        case EFGPU_VEC_H_UNCOARSENED: src = vec + nd.hbuf[0]; break;
        case EFGPU_VEC_G: src = vec + nd.gbuf[0]; break;
        case EFGPU_VEC_W: src = vec + nd.w_off; break;
        case EFGPU_VEC_U: src = H->d_u.as<double>() + (size_t)nd.leaf_idx * len; break;
        case EFGPU_VEC_F: src = H->d_f.as<double>() + (size_t)nd.leaf_idx * len; break;
    }
    EF_CUDA(cudaMemcpyAsync(out, src, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
    EF_CUDA(cudaStreamSynchronize(H->stream)); collect_profile(H);
    EF_CATCH(H)
}

int efgpu_set_profiling(efgpu_handle* H, int on)
{
    if (!H) return EF_ERR_BAD_ARG;
    H->profiling = on != 0;
    for (int c = 0; c < EFGPU_PROF_NCLASSES; c++) H->prof_ms[c] = H->prof_launches[c] = 0.0;
    return EF_OK;
}

int efgpu_get_profile(const efgpu_handle* H, int cls, double* ms, double* launches)
{
    if (!H || cls < 0 || cls >= EFGPU_PROF_NCLASSES) return EF_ERR_BAD_ARG;
    if (ms) *ms = H->prof_ms[cls];
    if (launches) *launches = H->prof_launches[cls];
    return EF_OK;
}

const char* efgpu_profile_class_name(int cls)
{
    static const char* names[EFGPU_PROF_NCLASSES] = {"leaf_dtn", "coarsen_T", "assemble_X_H", "invert_small", "gemm_Xinv", "gemm_S",
                                                      "gemm_T", "leaf_solve", "upwards_matvec", "solve_matvec", "coarsen_vec", "leaf_lu", "allgather",
                                                      "transpose_Xinv", "mirror_T"};
    return (cls >= 0 && cls < EFGPU_PROF_NCLASSES) ? names[cls] : "";
}

int efgpu_get_stats(const efgpu_handle* H, efgpu_stats_t* out)
{
    if (!H || !out) return EF_ERR_BAD_ARG;
    *out = H->stats;
    return EF_OK;
}

// Stand-alone entry to the descriptor GEMM for unit tests / roofline runs:
// C[b] = A[b] (m x k) * B[b] (k x n), all row-major and densely packed, device pointers.
int efgpu_dgemm_batched(const double* A, const double* B, double* C, int m, int n, int k, int batch, int tile, int iters, float* ms_out)
{
    try {
        std::vector<double*> ptab((size_t)batch * 3);
        for (int b = 0; b < batch; b++) {
            ptab[3 * (size_t)b + 0] = const_cast<double*>(A) + (size_t)b * m * k;
            ptab[3 * (size_t)b + 1] = const_cast<double*>(B) + (size_t)b * k * n;
            ptab[3 * (size_t)b + 2] = C + (size_t)b * m * n;
        }
        GemmBlock g{};
        g.c_op = 2; g.c_off = 0; g.ldc = n; g.c0_op = -1; g.rows = m; g.cols = n; g.nterms = 1;
        g.t[0] = GemmTerm{0, 1, k, n, 0, 0, k, 0u};
        DevBuf dp, db;
        cudaStream_t s = nullptr;
        dp.upload(ptab, s);
        std::vector<GemmBlock> gb(1, g);
        db.upload(gb, s);
        cudaEvent_t e0, e1; EF_CUDA(cudaEventCreate(&e0)); EF_CUDA(cudaEventCreate(&e1));
        launch_bgemm(dp.as<double*>(), 3, db.as<GemmBlock>(), gb.data(), 1, batch, s, tile);
        EF_CUDA(cudaEventRecord(e0, s));
        for (int it = 0; it < iters; it++) launch_bgemm(dp.as<double*>(), 3, db.as<GemmBlock>(), gb.data(), 1, batch, s, tile);
        EF_CUDA(cudaEventRecord(e1, s));
        EF_CUDA(cudaStreamSynchronize(s));
        float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_out) *ms_out = iters > 0 ? ms / iters : 0.f;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return EF_OK;
    } catch (const efgpu::Error& e) { g_create_error = e.msg; return e.code; }
}

int efgpu_dgemm_batched_tma(const double* A, const double* B, double* C, int m, int n, int k, int batch, int stages, int iters, float* ms_out)
{
    try {
        cudaStream_t s = nullptr;
        cudaEvent_t e0, e1; EF_CUDA(cudaEventCreate(&e0)); EF_CUDA(cudaEventCreate(&e1));
        launch_dgemm_tma(A, B, C, m, n, k, batch, stages, s);
        EF_CUDA(cudaEventRecord(e0, s));
        for (int it = 0; it < iters; it++) launch_dgemm_tma(A, B, C, m, n, k, batch, stages, s);
        EF_CUDA(cudaEventRecord(e1, s));
        EF_CUDA(cudaStreamSynchronize(s));
        float ms = 0; EF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_out) *ms_out = iters > 0 ? ms / iters : 0.f;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return EF_OK;
    } catch (const efgpu::Error& e) { g_create_error = e.msg; return e.code; }
}

}  // extern "C"
