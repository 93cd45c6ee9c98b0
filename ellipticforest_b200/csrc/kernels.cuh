// Launch wrappers of the non-GEMM kernels (leaf.cu, vec.cu).
#pragma once
#include "common.cuh"

namespace efgpu {

// ---- merge topology tables (src/HPSAlgorithm.hpp:752-990) -----------
// children c: 0 alpha, 1 beta, 2 gamma, 3 omega; sides: 0 W, 1 E, 2 S, 3 N;
// interior interfaces k: 0 alpha|gamma, 1 beta|omega, 2 alpha|beta, 3 gamma|omega.
struct MergeEntry {
    const double* Tc[4];  // children's DtN (coarsened where tagged), 4n x 4n, ld 4n
    double* Xinv;         // 4n x 4n
    double* S;            // 4n x 8n, columns in WESN order
    double* Hc;           // 8n x 2n compact H (rows in WESN order, the two non-zero n x n blocks per row block)
    double* T;            // 8n x 8n, WESN order (may be null when the DtN of this node is never used)
    double* Xcopy;        // optional copy of X before inversion (parity/debug) or null
    const double* hc[4];  // children's particular Neumann data (coarsened), 4n
    double* hd;           // scratch: jump of the children's h across the interfaces, 4n
    double* h;            // 8n
    double* w;            // 4n
    const double* g;      // 8n Dirichlet data of this node (uncoarsened)
    double* gc[4];        // children's Dirichlet data, 4n each
};

struct CoarsenOp { const double* src; double* dst; int nfine; int pad_; };

// leaf.cu
// build_list / n_build: leaf indices whose T is computed (one per class of bit-identical cell sizes; null: every leaf);
// leaf_src[leaf]: the computed leaf each leaf copies from
void launch_leaf_dtn_const(int M, const double* Q, const double* boxes, const int* leaf_nodes, double lambda,
                           double* T_all, int n_leaves, bool cache_operators, const int* build_list, int n_build, const int* leaf_src,
                           cudaStream_t s);
void launch_leaf_solve_const(int M, const double* Q, const double* boxes, const int* leaf_nodes, double lambda,
                             const double* f, double fscale, double* const* g_ptrs, double* u_out, double* const* h_ptrs,
                             int mode, int n_leaves, cudaStream_t s);

// variable-coefficient leaves: coef_in = {alpha, beta_w, beta_e, beta_s, beta_n, lambda} (device, leaf-major);
// coef = n_leaves x 4 x M^2 (cW, cE, cS, cN), P = n_leaves x M x M^2 (inverse diagonal blocks of the block LU)
void launch_leaf_var_factor(int M, const double* const* coef_in, const double* boxes, const int* leaf_nodes, double* coef, double* P,
                            double* min_pivot, int n_leaves, cudaStream_t s);
// mode 0: u = solve(g, f); mode 1: h = mapD2N(0, f); mode 2: T = buildD2N
void launch_leaf_var_solve(int M, const double* coef, const double* P, const double* boxes, const int* leaf_nodes, const double* f, double fscale,
                           double* const* g_ptrs, double* u_out, double* const* h_ptrs, double* T_all, int mode, int n_leaves, cudaStream_t s);
void launch_broadcast_leaf_T(double* T_all, int M, int n_leaves, cudaStream_t s);

// vec.cu
void launch_assemble_X(const MergeEntry* e, int n, int count, cudaStream_t s);
void launch_assemble_Hc(const MergeEntry* e, int n, int count, cudaStream_t s);
void launch_coarsen_T(const CoarsenOp* ops, int nops, int max_nfine, cudaStream_t s);
void launch_coarsen_h(const CoarsenOp* ops, int nops, int max_nfine, cudaStream_t s);
void launch_uncoarsen_g(const CoarsenOp* ops, int nops, int max_nfine, cudaStream_t s);
void launch_upwards(const MergeEntry* e, int n, int count, cudaStream_t s);            // hd, w, h
void launch_solve_split(const MergeEntry* e, int n, int count, bool add_w, cudaStream_t s);
// batched device-to-device copies of n doubles each (adaptive re-build: operators of unchanged subtrees)
struct CopyOp { const double* src; double* dst; size_t n; };
void launch_copy_many(const CopyOp* ops, int nops, cudaStream_t s);
// Newton-Schulz refinement of X^-1: mode 3: dst <- I + dst, max |dst| into *resid; mode 4: dst <- dst + src (N x N, contiguous)
void launch_refine_ew(double* const* ptab, int nops, int mode, int dst_op, long long dst_off, int src_op, long long src_off, int N, int batch,
                      double* resid, cudaStream_t s);
int get_tuning(int key);
void set_tuning(int key, int value);   // efgpu_set_tuning: kernel-selection knobs for measurements
void launch_expand_H(const double* Hc, int n, double* H_dense, cudaStream_t s);      // parity/debug: 8n x 4n, reference order

// sample.cu: sampling coordinates of every leaf (which: 0 cell centres, 1..4 W, E, S, N face midpoints; x or y may be null) and the
// error norms of the reference's drivers; part = 3 * n_leaves doubles of scratch, out = {l1, l2, linf} (device)
void launch_leaf_points(const double* boxes, const int* leaf_nodes, int M, int which, double* x, double* y, int n_leaves, cudaStream_t s);
void launch_error_norms(const double* u, const double* v, const double* boxes, const int* leaf_nodes, int M, int n_leaves, double area,
                        double* part, double* out, cudaStream_t s);

void launch_max_positive(const double* v, size_t n, double* out, cudaStream_t s);   // *out = max(0, max v): any sampled lambda > 0?

// vtk.cu: binary .vtu (raw appended data) of the leaf cells and `n_fields` cell fields (device arrays, one double per cell)
void write_vtu(const char* path, const double* boxes, const int* leaf_nodes, int M, long long n_leaves, int n_fields,
               const char* const* names, const double* const* fields, cudaStream_t s);

// lu.cu: root boundary system  (diag(a) + diag(b) T) g = r - b .* h  by blocked LU with partial pivoting
size_t robin_workspace_doubles(int N);
void robin_solve(const double* T, const double* a, const double* b, const double* r, const double* h, int N, double* ws,
                 double* g_out, int* info_host, cudaStream_t s);

}  // namespace efgpu
