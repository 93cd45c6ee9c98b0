// Leaf-patch kernels: Dirichlet-to-Neumann construction and leaf solves.
//
// Constant-coefficient leaves (the reference's FISHPACK90 branch,
// src/Patches/FiniteVolume/FiniteVolumeSolver.cpp:224-290 -> extern/fishpack90/src/hstcrt.f:386-446)
// are solved by fast diagonalisation: the M x M cell-centred 5-point operator with Dirichlet ghost
// reflection is  Dx (x) I + I (x) Dy + lambda,  Dx = tri(1,-2,1)/dx^2 with corner entries -3/dx^2,
// whose eigenvectors are q_k(i) = c_k sin((i+1/2) k pi / M), k = 1..M (a DST-II basis) with
// eigenvalues (2 cos(k pi / M) - 2)/dx^2.  Q is precomputed on the host (M x M, orthonormal).
//
//   leaf_dtn_const_kernel   : T (4M x 4M) per leaf, replaces buildD2N (:355-455) = 4M calls of
//                             mapD2N/solve per leaf; each M x M block of T is Q Z Q^T.
//   leaf_solve_const_kernel : u = solve(g, f) (:25-292) or h = mapD2N(g, f) (:296-353), batched.
//
// Variable-coefficient leaves (FivePointStencil branch :27-223, dense LU of the M^2 x M^2 matrix
// once per solve() call) use one banded LU per leaf (bandwidth M, factor once, 4M+1 right-hand
// sides), kernels leaf_var_*.
#include "common.cuh"
#include "kernels.cuh"

namespace efgpu {

// sides: 0 = W (i = 0), 1 = E (i = M-1), 2 = S (j = 0), 3 = N (j = M-1)
__device__ __forceinline__ double side_sign(int a) { return (a & 1) ? -1.0 : 1.0; }

template <int M>
__global__ void __launch_bounds__(M * M)
leaf_dtn_const_kernel(const double* __restrict__ Q, const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                      double lambda, double* __restrict__ T_all, const int* __restrict__ build_list, int n_build)
{
    __shared__ double sQ[M][M + 1];     // sQ[i][k] = q_{k+1}(i)
    __shared__ double sDinv[M][M + 1];  // 1 / (mu_k/dx^2 + mu_l/dy^2 + lambda), [k][l]
    __shared__ double sZ[M][M + 1];
    __shared__ double sP[M][M + 1];
    __shared__ double sMu[M];
    if ((int)blockIdx.x >= n_build) return;
    const int leaf = build_list ? build_list[blockIdx.x] : blockIdx.x;
    const int r = threadIdx.x / M, c = threadIdx.x % M;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    sQ[r][c] = Q[r * M + c];
    if (threadIdx.x < M) sMu[threadIdx.x] = 2.0 * cospi((double)(threadIdx.x + 1) / M) - 2.0;
    __syncthreads();
    sDinv[r][c] = 1.0 / (sMu[r] / (dx * dx) + sMu[c] / (dy * dy) + lambda);
    __syncthreads();
    double* T = T_all + (size_t)leaf * (16 * M * M);
    for (int a = 0; a < 4; a++) {
        for (int b = 0; b < 4; b++) {
            const bool ax = a < 2, bx = b < 2;
            const int ea = (a & 1) ? M - 1 : 0, eb = (b & 1) ? M - 1 : 0;
            // G_ab = E_a A^-1 E_b^T = Qrow * Z * Qcol^T with the index roles below.
            double z;
            if (ax && bx) {          // rows j <-> l, cols j' <-> l : Z = diag_l( sum_k qa[k] qb[k] / D[k][l] )
                z = 0.0;
                if (r == c) for (int k = 0; k < M; k++) z += sQ[ea][k] * sQ[eb][k] * sDinv[k][r];
            } else if (!ax && !bx) { // rows i <-> k, cols i' <-> k : Z = diag_k( sum_l qa[l] qb[l] / D[k][l] )
                z = 0.0;
                if (r == c) for (int l = 0; l < M; l++) z += sQ[ea][l] * sQ[eb][l] * sDinv[r][l];
            } else if (ax && !bx) {  // rows j <-> l (index r), cols i' <-> k (index c): Z[l][k] = qa_x[k] qb_y[l] / D[k][l]
                z = sQ[ea][c] * sQ[eb][r] * sDinv[c][r];
            } else {                 // rows i <-> k (index r), cols j' <-> l (index c): Z[k][l] = qa_y[l] qb_x[k] / D[k][l]
                z = sQ[ea][c] * sQ[eb][r] * sDinv[r][c];
            }
            sZ[r][c] = z;
            __syncthreads();
            double p = 0.0;
#pragma unroll 8
            for (int m = 0; m < M; m++) p += sQ[r][m] * sZ[m][c];
            sP[r][c] = p;
            __syncthreads();
            double g = 0.0;
#pragma unroll 8
            for (int m = 0; m < M; m++) g += sP[r][m] * sQ[c][m];
            // T_ab = s_a (2/d_a) ( -(2/d_b^2) G_ab - delta_ab I )   (mapD2N with g = e_c, f = 0)
            const double da = ax ? dx : dy, db = bx ? dx : dy;
            double v = -(2.0 / (db * db)) * g - ((a == b && r == c) ? 1.0 : 0.0);
            T[(size_t)(a * M + r) * (4 * M) + b * M + c] = side_sign(a) * (2.0 / da) * v;
            __syncthreads();
        }
    }
}

// Copies the first leaf's T to all others (reference option "cache-operators",
// src/HPSAlgorithm.hpp:134-139: one "T_leaf" for every leaf regardless of its size).
__global__ void broadcast_leaf_T_kernel(double* __restrict__ T_all, size_t elems_per_leaf, int n_leaves)
{
    const size_t total = elems_per_leaf * (size_t)(n_leaves - 1);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
        T_all[elems_per_leaf + e] = T_all[e % elems_per_leaf];
}

// The DtN map of a constant-coefficient leaf depends on (dx, dy, lambda) only: leaves with bit-identical cell sizes share
// one computation (the launcher builds one representative per class) and the others receive a copy, written at HBM speed
// from an L2-resident source.  src[leaf] = leaf index of the representative (itself for a representative).
__global__ void __launch_bounds__(256) copy_leaf_T_kernel(double* __restrict__ T_all, int elems2_per_leaf, const int* __restrict__ src, int n_leaves)
{
    double2* T2 = reinterpret_cast<double2*>(T_all);
    for (int leaf = blockIdx.y; leaf < n_leaves; leaf += gridDim.y) {
        const int from = src[leaf];
        if (from == leaf) continue;
        const double2* in = T2 + (size_t)from * elems2_per_leaf;
        double2* out = T2 + (size_t)leaf * elems2_per_leaf;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < elems2_per_leaf; e += gridDim.x * blockDim.x) __stcs(out + e, in[e]);
    }
}

// mode 0: write u (M*M per leaf);  mode 1: write h (4M per leaf, into h_ptrs[leaf]).
// g_ptrs may be null (g = 0), f may be null (f = 0).  f/u are leaf-major with cell index j + i*M.
template <int M>
__global__ void __launch_bounds__(M * M)
leaf_solve_const_kernel(const double* __restrict__ Q, const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                        double lambda, const double* __restrict__ f, double fscale, double* const* __restrict__ g_ptrs,
                        double* __restrict__ u_out, double* const* __restrict__ h_ptrs, int mode, int n_leaves)
{
    __shared__ double sQ[M][M + 1];
    __shared__ double sA[M][M + 1];
    __shared__ double sB[M][M + 1];
    __shared__ double sMu[M];
    __shared__ double sG[4 * M];
    const int leaf = blockIdx.x;
    if (leaf >= n_leaves) return;
    const int i = threadIdx.x / M, j = threadIdx.x % M;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    sQ[i][j] = Q[i * M + j];
    if (threadIdx.x < M) sMu[threadIdx.x] = 2.0 * cospi((double)(threadIdx.x + 1) / M) - 2.0;
    if (threadIdx.x < 4 * M) sG[threadIdx.x] = g_ptrs ? g_ptrs[leaf][threadIdx.x] : 0.0;
    __syncthreads();
    // right-hand side with the Dirichlet data folded in (hstcrt.f:412-439)
    double rhs = f ? fscale * f[(size_t)leaf * M * M + threadIdx.x] : 0.0;
    if (i == 0) rhs -= 2.0 / (dx * dx) * sG[j];
    if (i == M - 1) rhs -= 2.0 / (dx * dx) * sG[M + j];
    if (j == 0) rhs -= 2.0 / (dy * dy) * sG[2 * M + i];
    if (j == M - 1) rhs -= 2.0 / (dy * dy) * sG[3 * M + i];
    sA[i][j] = rhs;
    __syncthreads();
    // forward transform: Rhat = Q^T R Q   (index [k][l])
    double t = 0.0;
#pragma unroll 8
    for (int m = 0; m < M; m++) t += sQ[m][i] * sA[m][j];   // (Q^T R)[k=i][j]
    sB[i][j] = t;
    __syncthreads();
    t = 0.0;
#pragma unroll 8
    for (int m = 0; m < M; m++) t += sB[i][m] * sQ[m][j];   // [k=i][l=j]
    t /= (sMu[i] / (dx * dx) + sMu[j] / (dy * dy) + lambda);
    __syncthreads();
    sA[i][j] = t;
    __syncthreads();
    // inverse transform: U = Q Uhat Q^T
    t = 0.0;
#pragma unroll 8
    for (int m = 0; m < M; m++) t += sQ[i][m] * sA[m][j];
    sB[i][j] = t;
    __syncthreads();
    t = 0.0;
#pragma unroll 8
    for (int m = 0; m < M; m++) t += sB[i][m] * sQ[j][m];
    if (mode == 0) {
        u_out[(size_t)leaf * M * M + threadIdx.x] = t;
    } else {
        // mapD2N (FiniteVolumeSolver.cpp:332-343): coordinate derivatives on the four sides
        double* h = h_ptrs[leaf];
        if (i == 0) h[j] = (2.0 / dx) * (t - sG[j]);
        if (i == M - 1) h[M + j] = -(2.0 / dx) * (t - sG[M + j]);
        if (j == 0) h[2 * M + i] = (2.0 / dy) * (t - sG[2 * M + i]);
        if (j == M - 1) h[3 * M + i] = -(2.0 / dy) * (t - sG[3 * M + i]);
    }
}


// Patches too large for one thread per cell (M = 64: the reference's plots carry a 64 x 64 series; hstcrt accepts any M > 2,
// extern/fishpack90/src/hstcrt.f:336-339): the same two kernels with NT threads looping over the M^2 cells and the tiles in
// dynamic shared memory.  Same arithmetic, same summation order per element.
template <int M, int NT>
__global__ void __launch_bounds__(NT)
leaf_dtn_const_loop_kernel(const double* __restrict__ Q, const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                           double lambda, double* __restrict__ T_all, const int* __restrict__ build_list, int n_build)
{
    extern __shared__ __align__(16) double sml[];
    constexpr int LD = M + 1;
    double* sQ = sml; double* sDinv = sQ + M * LD; double* sZ = sDinv + M * LD; double* sP = sZ + M * LD; double* sMu = sP + M * LD;
    if ((int)blockIdx.x >= n_build) return;
    const int leaf = build_list ? build_list[blockIdx.x] : blockIdx.x;
    const int tid = threadIdx.x;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    for (int e = tid; e < M * M; e += NT) sQ[(e / M) * LD + e % M] = Q[e];
    for (int e = tid; e < M; e += NT) sMu[e] = 2.0 * cospi((double)(e + 1) / M) - 2.0;
    __syncthreads();
    for (int e = tid; e < M * M; e += NT) sDinv[(e / M) * LD + e % M] = 1.0 / (sMu[e / M] / (dx * dx) + sMu[e % M] / (dy * dy) + lambda);
    __syncthreads();
    double* T = T_all + (size_t)leaf * (16 * M * M);
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            const bool ax = a < 2, bx = b < 2;
            const int ea = (a & 1) ? M - 1 : 0, eb = (b & 1) ? M - 1 : 0;
            for (int e = tid; e < M * M; e += NT) {
                const int r = e / M, c = e % M;
                double z = 0.0;
                if (ax && bx) { if (r == c) for (int k = 0; k < M; k++) z += sQ[ea * LD + k] * sQ[eb * LD + k] * sDinv[k * LD + r]; }
                else if (!ax && !bx) { if (r == c) for (int l = 0; l < M; l++) z += sQ[ea * LD + l] * sQ[eb * LD + l] * sDinv[r * LD + l]; }
                else if (ax && !bx) z = sQ[ea * LD + c] * sQ[eb * LD + r] * sDinv[c * LD + r];
                else z = sQ[ea * LD + c] * sQ[eb * LD + r] * sDinv[r * LD + c];
                sZ[r * LD + c] = z;
            }
            __syncthreads();
            for (int e = tid; e < M * M; e += NT) {
                const int r = e / M, c = e % M;
                double p = 0.0;
#pragma unroll 8
                for (int m = 0; m < M; m++) p += sQ[r * LD + m] * sZ[m * LD + c];
                sP[r * LD + c] = p;
            }
            __syncthreads();
            const double da = ax ? dx : dy, db = bx ? dx : dy;
            for (int e = tid; e < M * M; e += NT) {
                const int r = e / M, c = e % M;
                double g = 0.0;
#pragma unroll 8
                for (int m = 0; m < M; m++) g += sP[r * LD + m] * sQ[c * LD + m];
                const double v = -(2.0 / (db * db)) * g - ((a == b && r == c) ? 1.0 : 0.0);
                T[(size_t)(a * M + r) * (4 * M) + b * M + c] = side_sign(a) * (2.0 / da) * v;
            }
            __syncthreads();
        }
}

template <int M, int NT>
__global__ void __launch_bounds__(NT)
leaf_solve_const_loop_kernel(const double* __restrict__ Q, const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                             double lambda, const double* __restrict__ f, double fscale, double* const* __restrict__ g_ptrs,
                             double* __restrict__ u_out, double* const* __restrict__ h_ptrs, int mode, int n_leaves)
{
    extern __shared__ __align__(16) double sml[];
    constexpr int LD = M + 1;
    double* sQ = sml; double* sA = sQ + M * LD; double* sB = sA + M * LD; double* sMu = sB + M * LD; double* sG = sMu + M;
    const int leaf = blockIdx.x, tid = threadIdx.x;
    if (leaf >= n_leaves) return;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    for (int e = tid; e < M * M; e += NT) sQ[(e / M) * LD + e % M] = Q[e];
    for (int e = tid; e < M; e += NT) sMu[e] = 2.0 * cospi((double)(e + 1) / M) - 2.0;
    for (int e = tid; e < 4 * M; e += NT) sG[e] = g_ptrs ? g_ptrs[leaf][e] : 0.0;
    __syncthreads();
    for (int e = tid; e < M * M; e += NT) {
        const int i = e / M, j = e % M;
        double rhs = f ? fscale * f[(size_t)leaf * M * M + e] : 0.0;
        if (i == 0) rhs -= 2.0 / (dx * dx) * sG[j];
        if (i == M - 1) rhs -= 2.0 / (dx * dx) * sG[M + j];
        if (j == 0) rhs -= 2.0 / (dy * dy) * sG[2 * M + i];
        if (j == M - 1) rhs -= 2.0 / (dy * dy) * sG[3 * M + i];
        sA[i * LD + j] = rhs;
    }
    __syncthreads();
    for (int e = tid; e < M * M; e += NT) {
        const int i = e / M, j = e % M;
        double t = 0.0;
#pragma unroll 8
        for (int m = 0; m < M; m++) t += sQ[m * LD + i] * sA[m * LD + j];
        sB[i * LD + j] = t;
    }
    __syncthreads();
    for (int e = tid; e < M * M; e += NT) {
        const int i = e / M, j = e % M;
        double t = 0.0;
#pragma unroll 8
        for (int m = 0; m < M; m++) t += sB[i * LD + m] * sQ[m * LD + j];
        sA[i * LD + j] = t / (sMu[i] / (dx * dx) + sMu[j] / (dy * dy) + lambda);
    }
    __syncthreads();
    for (int e = tid; e < M * M; e += NT) {
        const int i = e / M, j = e % M;
        double t = 0.0;
#pragma unroll 8
        for (int m = 0; m < M; m++) t += sQ[i * LD + m] * sA[m * LD + j];
        sB[i * LD + j] = t;
    }
    __syncthreads();
    for (int e = tid; e < M * M; e += NT) {
        const int i = e / M, j = e % M;
        double t = 0.0;
#pragma unroll 8
        for (int m = 0; m < M; m++) t += sB[i * LD + m] * sQ[j * LD + m];
        if (mode == 0) u_out[(size_t)leaf * M * M + e] = t;
        else {
            double* h = h_ptrs[leaf];
            if (i == 0) h[j] = (2.0 / dx) * (t - sG[j]);
            if (i == M - 1) h[M + j] = -(2.0 / dx) * (t - sG[M + j]);
            if (j == 0) h[2 * M + i] = (2.0 / dy) * (t - sG[2 * M + i]);
            if (j == M - 1) h[3 * M + i] = -(2.0 / dy) * (t - sG[3 * M + i]);
        }
    }
}

// FP64 tensor-core variant of the leaf solve, M = 8, 16, 24, 32: ONE WARP PER LEAF, the four M x M products of the
// fast diagonalisation issued as DMMA m8n8k4 (same instruction as csrc/gemm.cu).  The whole M x M result of a product
// lives in the warp's accumulator registers, so one padded shared-memory tile per warp is enough: read fragments,
// __syncwarp, write the result over it.  Leading dimension M + 4 == 4 or 12 (mod 16) makes the 8 x 4 (A, row-major), the
// 4 x 8 (B) and both transposed fragment reads bank-conflict free, so one row-major copy of Q serves Q and Q^T on either
// side.  No CTA-wide barrier inside the leaf loop: warps stream f / u independently and hide each other's HBM latency.
// Per FMA it moves 1/8 of the shared-memory bytes of a 4 x 4 register-tiled FMA kernel (measured 1.9x slower, M = 16).
__device__ __forceinline__ void leaf_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// acc = A B with A(row, k) = A[row * as_r + k * as_k], B(k, col) = B[k * bs_k + col * bs_c]; g = lane / 4, t = lane % 4
template <int M>
__device__ __forceinline__ void warp_mm(const double* __restrict__ A, int as_r, int as_k, const double* __restrict__ B, int bs_k, int bs_c,
                                        int g, int t, double (&acc)[M / 8][M / 8][2])
{
    constexpr int F = M / 8;
#pragma unroll
    for (int i = 0; i < F; i++)
#pragma unroll
        for (int j = 0; j < F; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* a0 = A + g * as_r + t * as_k;
    const double* b0 = B + t * bs_k + g * bs_c;
#pragma unroll
    for (int kk = 0; kk < M / 4; kk++) {
        double a[F], b[F];
#pragma unroll
        for (int i = 0; i < F; i++) a[i] = a0[8 * i * as_r + 4 * kk * as_k];
#pragma unroll
        for (int j = 0; j < F; j++) b[j] = b0[4 * kk * bs_k + 8 * j * bs_c];
#pragma unroll
        for (int i = 0; i < F; i++)
#pragma unroll
            for (int j = 0; j < F; j++) leaf_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

template <int M>
__device__ __forceinline__ void warp_store(double* __restrict__ buf, int ld, int g, int t, const double (&acc)[M / 8][M / 8][2])
{
#pragma unroll
    for (int i = 0; i < M / 8; i++)
#pragma unroll
        for (int j = 0; j < M / 8; j++)
            *reinterpret_cast<double2*>(buf + (8 * i + g) * ld + 8 * j + 2 * t) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// Shared memory per CTA: Q (M x LD) | per warp: 2 tiles (M x LD), 2 x g (4M) | per warp 2 mbarriers.
// Each warp double-buffers its leaves' inputs with the bulk-copy engine: before it starts the four products of the
// current leaf, lane i issues a cp.async.bulk of row i of the NEXT leaf's f (M*8 bytes) straight into the padded rows
// of the other tile, and lane 0 one for g (4M*8 bytes); completion is counted on the slot's mbarrier.
template <int M, int WARPS>
struct LeafMmaSmem {
    static constexpr int LD = M + 4;
    static constexpr int PER_WARP = 2 * M * LD + 2 * 4 * M;                         // doubles
    static constexpr int BYTES = (M * LD + WARPS * PER_WARP + WARPS * 2) * 8;
};

template <int M, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
leaf_solve_const_mma_kernel(const double* __restrict__ Q, const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                            double lambda, const double* __restrict__ f, double fscale, double* const* __restrict__ g_ptrs,
                            double* __restrict__ u_out, double* const* __restrict__ h_ptrs, int mode, int n_leaves)
{
    using SM = LeafMmaSmem<M, WARPS>;
    constexpr int LD = SM::LD, F = M / 8;
    extern __shared__ __align__(128) double smm[];
    double* sQ = smm;                                       // M x LD, sQ[i][k] = q_{k+1}(i)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* tiles = smm + M * LD + warp * SM::PER_WARP;     // [2][M x LD]  f lands here; the products run in place
    double* gsm = tiles + 2 * M * LD;                       // [2][4M]      Dirichlet data
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smm + M * LD + WARPS * SM::PER_WARP) + warp * 2;
    for (int e = threadIdx.x; e < M * M; e += WARPS * 32) sQ[(e / M) * LD + (e % M)] = Q[e];
    if (lane == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // eigenvalue numerators of the rows / columns this lane owns in the accumulator layout
    double muR[F], muC[F][2];
#pragma unroll
    for (int i = 0; i < F; i++) muR[i] = 2.0 * cospi((double)(8 * i + g + 1) / M) - 2.0;
#pragma unroll
    for (int j = 0; j < F; j++) {
        muC[j][0] = 2.0 * cospi((double)(8 * j + 2 * t + 1) / M) - 2.0;
        muC[j][1] = 2.0 * cospi((double)(8 * j + 2 * t + 2) / M) - 2.0;
    }
    __syncthreads();
    const int stride = gridDim.x * WARPS;
    const int leaf0 = blockIdx.x * WARPS + warp;
    const unsigned tx_bytes = (f ? M * M * 8u : 0u) + (g_ptrs ? 4 * M * 8u : 0u);
    // start the copies of `leaf` into slot s; gp = that leaf's g pointer (already in a register)
    auto issue = [&](int leaf, int s, const double* gp) {
        if (!tx_bytes) return;
        if (lane == 0) {
            mbar_expect_tx(bar + s, tx_bytes);
            if (g_ptrs) bulk_g2s(gsm + s * 4 * M, gp, 4 * M * 8u, bar + s);
        }
        __syncwarp();
        if (f && lane < M) bulk_g2s(tiles + s * M * LD + lane * LD, f + (size_t)leaf * M * M + lane * M, M * 8u, bar + s);
    };
    const double* gp_next = nullptr;    // g pointer of the leaf after the current one, loaded one iteration ahead
    if (leaf0 < n_leaves) {
        issue(leaf0, 0, g_ptrs ? g_ptrs[leaf0] : nullptr);
        if (g_ptrs && leaf0 + stride < n_leaves) gp_next = g_ptrs[leaf0 + stride];
    }
    double acc[F][F][2];
    int it = 0;
    for (int leaf = leaf0; leaf < n_leaves; leaf += stride, it++) {
        const int s = it & 1;
        {   // prefetch: the other slot was last read before the __syncwarp that closed the previous leaf's staging
            const int next = leaf + stride;
            if (next < n_leaves) {
                issue(next, s ^ 1, gp_next);
                if (g_ptrs && next + stride < n_leaves) gp_next = g_ptrs[next + stride];
            }
        }
        const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
        const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
        const double rx = 1.0 / (dx * dx), ry = 1.0 / (dy * dy);
        if (tx_bytes) mbar_wait(bar + s, (unsigned)((it >> 1) & 1));
        double* buf = tiles + s * M * LD;
        const double* gl = g_ptrs ? gsm + s * 4 * M : nullptr;
        // right-hand side with the Dirichlet data folded in (hstcrt.f:412-439), in place; cell (i, j) at index j + i*M
        {
#pragma unroll
            for (int k = 0; k < M * M / 64; k++) {
                const int e2 = lane + 32 * k;
                const int i = (2 * e2) / M, j = (2 * e2) % M;   // j even: both entries lie in row i
                double2 v = make_double2(0.0, 0.0);
                if (f) { v = *reinterpret_cast<const double2*>(buf + i * LD + j); v.x *= fscale; v.y *= fscale; }
                if (gl) {
                    if (i == 0) { v.x -= 2.0 * rx * gl[j]; v.y -= 2.0 * rx * gl[j + 1]; }
                    if (i == M - 1) { v.x -= 2.0 * rx * gl[M + j]; v.y -= 2.0 * rx * gl[M + j + 1]; }
                    if (j == 0) v.x -= 2.0 * ry * gl[2 * M + i];
                    if (j == M - 2) v.y -= 2.0 * ry * gl[3 * M + i];
                }
                *reinterpret_cast<double2*>(buf + i * LD + j) = v;
            }
        }
        __syncwarp();
        // B = Q^T R : A(k, m) = Q[m][k], B(m, j) = R[m][j]
        warp_mm<M>(sQ, 1, LD, buf, LD, 1, g, t, acc);
        __syncwarp();
        warp_store<M>(buf, LD, g, t, acc);
        __syncwarp();
        // A = (B Q) / D : A(k, m) = B[k][m], B(m, l) = Q[m][l]
        warp_mm<M>(buf, LD, 1, sQ, LD, 1, g, t, acc);
#pragma unroll
        for (int i = 0; i < F; i++)
#pragma unroll
            for (int j = 0; j < F; j++) {
                acc[i][j][0] /= (muR[i] * rx + muC[j][0] * ry + lambda);
                acc[i][j][1] /= (muR[i] * rx + muC[j][1] * ry + lambda);
            }
        __syncwarp();
        warp_store<M>(buf, LD, g, t, acc);
        __syncwarp();
        // B = Q A : A(i, m) = Q[i][m], B(m, l) = A[m][l]
        warp_mm<M>(sQ, LD, 1, buf, LD, 1, g, t, acc);
        __syncwarp();
        warp_store<M>(buf, LD, g, t, acc);
        __syncwarp();
        // U = B Q^T : A(i, m) = B[i][m], B(m, j) = Q[j][m]
        warp_mm<M>(buf, LD, 1, sQ, 1, LD, g, t, acc);
        if (mode == 0) {
            double* u = u_out + (size_t)leaf * M * M;
#pragma unroll
            for (int i = 0; i < F; i++)
#pragma unroll
                for (int j = 0; j < F; j++)
                    __stcs(reinterpret_cast<double2*>(u + (8 * i + g) * M + 8 * j + 2 * t), make_double2(acc[i][j][0], acc[i][j][1]));
        } else {
            // mapD2N (FiniteVolumeSolver.cpp:332-343): coordinate derivatives on the four sides
            double* h = h_ptrs[leaf];
#pragma unroll
            for (int i = 0; i < F; i++)
#pragma unroll
                for (int j = 0; j < F; j++)
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int ci = 8 * i + g, cj = 8 * j + 2 * t + q;
                        const double v = acc[i][j][q];
                        if (ci == 0) h[cj] = (2.0 / dx) * (v - (gl ? gl[cj] : 0.0));
                        if (ci == M - 1) h[M + cj] = -(2.0 / dx) * (v - (gl ? gl[M + cj] : 0.0));
                        if (cj == 0) h[2 * M + ci] = (2.0 / dy) * (v - (gl ? gl[2 * M + ci] : 0.0));
                        if (cj == M - 1) h[3 * M + ci] = -(2.0 / dy) * (v - (gl ? gl[3 * M + ci] : 0.0));
                    }
        }
        __syncwarp();   // every lane is done with the tile and with slot s before the next iteration refills them
    }
}

template <int M>
static void solve_const_mma_M(const double* Q, const double* boxes, const int* leaf_nodes, double lambda, const double* f, double fscale,
                              double* const* g_ptrs, double* u_out, double* const* h_ptrs, int mode, int n_leaves, cudaStream_t s) {
    constexpr int WARPS = M >= 32 ? 4 : 8;
    constexpr int smem = LeafMmaSmem<M, WARPS>::BYTES;
    auto kern = leaf_solve_const_mma_kernel<M, WARPS>;
    static int resident = 0;   // CTAs the device holds at once: the leaves are dealt to them in a grid-stride loop
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared)) {   // (the devices of one box are identical: `resident` is the same for all of them)
        EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev = 0, sms = 0, per_sm = 0;
        EF_CUDA(cudaGetDevice(&dev));
        EF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        EF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
        resident = sms * (per_sm > 0 ? per_sm : 1);
    }
    int grid = (n_leaves + WARPS - 1) / WARPS;
    if (grid > resident) grid = resident;
    kern<<<grid, WARPS * 32, smem, s>>>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves);
}

// =================================================================================================
// Variable-coefficient leaves (reference: FivePointStencil branch, FiniteVolumeSolver.cpp:27-223).
// The reference assembles the dense M^2 x M^2 five-point matrix and LU-factorises it (PETSc ->
// LAPACK dgetrf, partial pivoting) once per solve() call, i.e. 4M+ times per leaf.  With cells
// ordered j + i*M the matrix is block tridiagonal: M diagonal blocks D_i (tridiagonal in j, M x M),
// off-diagonal blocks W_i = diag(cW[i][.]), E_i = diag(cE[i][.]).  Here each leaf is factorised ONCE
// by block elimination along i (no pivoting: the matrix is row diagonally dominant for lambda <= 0):
//     Delta_0 = D_0,  Delta_i = D_i - W_i Delta_{i-1}^{-1} E_{i-1},   P_i = Delta_i^{-1}  (dense M x M, stored)
// and every later solve is two sweeps of M x M mat-vecs:
//     forward   y_i = r_i - W_i z_{i-1},  z_i = P_i y_i
//     backward  u_{M-1} = z_{M-1},        u_i = z_i - P_i (E_i u_{i+1}).
// =================================================================================================

// One CTA per leaf, thread (r, c) owns element (r, c) of the current block in a register; the
// Gauss-Jordan inverse broadcasts pivot row / column through double-buffered shared memory
// (one barrier per pivot).
template <int M>
__global__ void __launch_bounds__(M * M)
leaf_var_factor_kernel(const double* __restrict__ alpha, const double* __restrict__ bw, const double* __restrict__ be,
                       const double* __restrict__ bs, const double* __restrict__ bn, const double* __restrict__ lam,
                       const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                       double* __restrict__ coef, double* __restrict__ P_all, double* __restrict__ min_pivot)
{
    __shared__ double sRow[2][M];
    __shared__ double sCol[2][M];
    __shared__ double sCE[M];   // cE of the previous block column
    const int leaf = blockIdx.x;
    const int r = threadIdx.x / M, c = threadIdx.x % M;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    const size_t cell0 = (size_t)leaf * M * M;
    double* cf = coef + (size_t)leaf * 4 * M * M;
    // coefficients (FiniteVolumeSolver.cpp:74-85); thread (r, c) handles cell i = r, j = c
    {
        const size_t k = cell0 + threadIdx.x;   // j + i*M with i = r, j = c
        const double a = alpha[k];
        cf[0 * M * M + threadIdx.x] = a * bw[k] / (dx * dx);
        cf[1 * M * M + threadIdx.x] = a * be[k] / (dx * dx);
        cf[2 * M * M + threadIdx.x] = a * bs[k] / (dy * dy);
        cf[3 * M * M + threadIdx.x] = a * bn[k] / (dy * dy);
    }
    __syncthreads();
    double prev = 0.0;   // P_{i-1}[r][c]
    double minp = 1e300;
    for (int i = 0; i < M; i++) {
        // block row i, element (r, c): r, c are j indices
        double a = 0.0;
        {
            const size_t kr = cell0 + (size_t)i * M + r;
            const double al = alpha[kr];
            const double cW = al * bw[kr] / (dx * dx), cE = al * be[kr] / (dx * dx), cS = al * bs[kr] / (dy * dy), cN = al * bn[kr] / (dy * dy);
            if (r == c) {
                double d = -al * ((be[kr] + bw[kr]) / (dx * dx) + (bn[kr] + bs[kr]) / (dy * dy)) + lam[kr];
                if (i == 0) d -= cW;
                if (i == M - 1) d -= cE;
                if (r == 0) d -= cS;
                if (r == M - 1) d -= cN;
                a = d;
            } else if (c == r - 1) a = cS;
            else if (c == r + 1) a = cN;
            if (i > 0) a -= cW * prev * sCE[c];
        }
        // in-register Gauss-Jordan inverse of the M x M block
        for (int k = 0; k < M; k++) {
            const int b = k & 1;
            if (r == k) sRow[b][c] = a;
            if (c == k) sCol[b][r] = a;
            __syncthreads();
            const double piv = sRow[b][k];
            const double p = 1.0 / piv;
            if (threadIdx.x == 0) minp = fmin(minp, fabs(piv));
            if (r == k) a = (c == k) ? p : a * p;
            else {
                const double f = sCol[b][r];
                a = (c == k) ? -f * p : a - f * p * sRow[b][c];
            }
        }
        P_all[((size_t)leaf * M + i) * M * M + threadIdx.x] = a;
        prev = a;
        __syncthreads();
        if (threadIdx.x < M) sCE[threadIdx.x] = cf[1 * M * M + i * M + threadIdx.x];   // cE[i][.] for the next block
        __syncthreads();
    }
    if (threadIdx.x == 0 && min_pivot)
        atomicMin(reinterpret_cast<unsigned long long*>(min_pivot), (unsigned long long)__double_as_longlong(minp));
}

// Block-tridiagonal solve for `ncols` right-hand sides of one leaf per CTA (blockIdx.x = leaf,
// blockIdx.y = column chunk).  mode 0: u = solve(g, f)      (one column; writes u_out)
//                              mode 1: h = mapD2N(g=0, f)   (one column; writes h_ptrs[leaf])
//                              mode 2: columns of T: g = e_col, f = 0 (writes T column col)
// Shared memory: P tile M x (M+1), tmp M x C, Z M*M x C.
template <int M>
__global__ void __launch_bounds__(256)
leaf_var_solve_kernel(const double* __restrict__ coef, const double* __restrict__ P_all, const double* __restrict__ boxes,
                      const int* __restrict__ leaf_nodes, const double* __restrict__ f, double fscale, double* const* __restrict__ g_ptrs,
                      double* __restrict__ u_out, double* const* __restrict__ h_ptrs, double* __restrict__ T_all, int mode, int C)
{
    extern __shared__ __align__(16) double smv[];
    double* sP = smv;                       // M x (M+1)
    double* sT = sP + M * (M + 1);          // M x C
    double* sZ = sT + M * C;                // M*M x C
    const int leaf = blockIdx.x, col0 = blockIdx.y * C;
    const int tid = threadIdx.x, NT = blockDim.x;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    const double* cW = coef + (size_t)leaf * 4 * M * M;
    const double* cE = cW + M * M;
    const double* cS = cE + M * M;
    const double* cN = cS + M * M;
    const double* P = P_all + (size_t)leaf * M * M * M;
    const double* gl = (mode == 0 && g_ptrs) ? g_ptrs[leaf] : nullptr;
    const double* fl = f ? f + (size_t)leaf * M * M : nullptr;

    // right-hand side of cell (i, j) for local column cc (FiniteVolumeSolver.cpp:100-175: rhs += -2 c_side g)
    auto rhs = [&](int i, int j, int cc) -> double {
        double v = 0.0;
        if (mode == 2) {
            const int col = col0 + cc, side = col / M, t = col % M;
            if (side == 0 && i == 0 && j == t) v = -2.0 * cW[j];
            else if (side == 1 && i == M - 1 && j == t) v = -2.0 * cE[(M - 1) * M + j];
            else if (side == 2 && j == 0 && i == t) v = -2.0 * cS[i * M];
            else if (side == 3 && j == M - 1 && i == t) v = -2.0 * cN[i * M + M - 1];
            return v;
        }
        if (fl) v = fscale * fl[i * M + j];
        if (gl) {
            if (i == 0) v += -2.0 * cW[j] * gl[j];
            if (i == M - 1) v += -2.0 * cE[(M - 1) * M + j] * gl[M + j];
            if (j == 0) v += -2.0 * cS[i * M] * gl[2 * M + i];
            if (j == M - 1) v += -2.0 * cN[i * M + M - 1] * gl[3 * M + i];
        }
        return v;
    };

    // forward sweep
    for (int i = 0; i < M; i++) {
        for (int e = tid; e < M * M; e += NT) sP[(e / M) * (M + 1) + (e % M)] = P[(size_t)i * M * M + e];
        for (int e = tid; e < M * C; e += NT) {
            const int j = e / C, cc = e % C;
            double y = rhs(i, j, cc);
            if (i > 0) y -= cW[i * M + j] * sZ[((i - 1) * M + j) * C + cc];
            sT[e] = y;
        }
        __syncthreads();
        for (int e = tid; e < M * C; e += NT) {
            const int j = e / C, cc = e % C;
            double z = 0.0;
#pragma unroll 8
            for (int k = 0; k < M; k++) z = fma(sP[j * (M + 1) + k], sT[k * C + cc], z);
            sZ[(i * M + j) * C + cc] = z;
        }
        __syncthreads();
    }
    // backward sweep (u overwrites z)
    for (int i = M - 2; i >= 0; i--) {
        for (int e = tid; e < M * M; e += NT) sP[(e / M) * (M + 1) + (e % M)] = P[(size_t)i * M * M + e];
        for (int e = tid; e < M * C; e += NT) {
            const int j = e / C, cc = e % C;
            sT[e] = cE[i * M + j] * sZ[((i + 1) * M + j) * C + cc];
        }
        __syncthreads();
        for (int e = tid; e < M * C; e += NT) {
            const int j = e / C, cc = e % C;
            double z = 0.0;
#pragma unroll 8
            for (int k = 0; k < M; k++) z = fma(sP[j * (M + 1) + k], sT[k * C + cc], z);
            sZ[(i * M + j) * C + cc] -= z;
        }
        __syncthreads();
    }
    // outputs
    if (mode == 0) {
        for (int e = tid; e < M * M; e += NT) u_out[(size_t)leaf * M * M + e] = sZ[e * C];
    } else if (mode == 1) {
        double* h = h_ptrs[leaf];
        for (int e = tid; e < 4 * M; e += NT) {
            const int side = e / M, t = e % M;
            double v;
            if (side == 0) v = (2.0 / dx) * sZ[(0 * M + t) * C];
            else if (side == 1) v = -(2.0 / dx) * sZ[((M - 1) * M + t) * C];
            else if (side == 2) v = (2.0 / dy) * sZ[(t * M + 0) * C];
            else v = -(2.0 / dy) * sZ[(t * M + M - 1) * C];
            h[e] = v;
        }
    } else {
        double* T = T_all + (size_t)leaf * 16 * M * M;
        for (int e = tid; e < 4 * M * C; e += NT) {
            const int row = e / C, cc = e % C, col = col0 + cc;
            const int side = row / M, t = row % M;
            double u;
            if (side == 0) u = sZ[(0 * M + t) * C + cc];
            else if (side == 1) u = sZ[((M - 1) * M + t) * C + cc];
            else if (side == 2) u = sZ[(t * M + 0) * C + cc];
            else u = sZ[(t * M + M - 1) * C + cc];
            const double gv = (row == col) ? 1.0 : 0.0;
            const double d = side < 2 ? dx : dy;
            T[(size_t)row * (4 * M) + col] = side_sign(side) * (2.0 / d) * (u - gv);
        }
    }
}

// ---- warp-level variants for M = 8 and 16 (round 2) --------------------------------------------------------------------------
// The CTA-per-leaf kernels above pay one __syncthreads per pivot (factor) and stage 71 KB of shared memory per CTA for the DtN
// columns (ncu, profiles/r2a_ncu_varcoef_leaf.md: FP64 pipe 38 % / 9 % busy, 4.98 + 14.4 ms on BASELINE configs[3]).  Here:
//   leaf_var_factor_warp_kernel   one WARP per leaf, the M x M block in registers (lane = column, M / (32 / M) rows per lane),
//                                 Gauss-Jordan by warp shuffles: no shared memory, no barrier;
//   leaf_var_dtn_mma_kernel       one CTA per leaf, one warp per 8 columns of T; the M x M products P_i y_i of both sweeps run on
//                                 the FP64 tensor pipe (mma.m8n8k4), the whole intermediate Z (M^2 x 8 per warp) lives in registers
//                                 as accumulator fragments, P_i is staged once per step for all warps, y_i crosses from the
//                                 accumulator layout to the B-operand layout through a 768-byte per-warp tile;
//   leaf_var_solve_warp_kernel    one warp per leaf for the single right-hand side of upwards / solve (streams P twice: HBM-bound).
// Same arithmetic as the kernels above (block elimination without pivoting), different summation order inside the M x M products.
template <int M>
__global__ void __launch_bounds__(256)
leaf_var_factor_warp_kernel(const double* __restrict__ alpha, const double* __restrict__ bw, const double* __restrict__ be,
                            const double* __restrict__ bs, const double* __restrict__ bn, const double* __restrict__ lam,
                            const double* __restrict__ boxes, const int* __restrict__ leaf_nodes,
                            double* __restrict__ coef, double* __restrict__ P_all, double* __restrict__ min_pivot, int n_leaves)
{
    constexpr int RG = 32 / M, RPL = M / RG;   // row groups per warp, rows per lane
    const int lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (leaf >= n_leaves) return;
    const int c = lane % M, rg = lane / M, gbase = rg * M;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    const size_t cell0 = (size_t)leaf * M * M;
    double* cf = coef + (size_t)leaf * 4 * M * M;
    double prev[RPL];
#pragma unroll
    for (int q = 0; q < RPL; q++) prev[q] = 0.0;
    double cE_prev = 0.0;   // cE[i-1][c]
    double minp = 1e300;
    for (int i = 0; i < M; i++) {
        double a[RPL];
#pragma unroll
        for (int q = 0; q < RPL; q++) {
            const int r = rg * RPL + q;
            const size_t kr = cell0 + (size_t)i * M + r;
            const double al = alpha[kr];
            const double cW = al * bw[kr] / (dx * dx), cE = al * be[kr] / (dx * dx), cS = al * bs[kr] / (dy * dy), cN = al * bn[kr] / (dy * dy);
            if (c == 0) { cf[0 * M * M + i * M + r] = cW; cf[1 * M * M + i * M + r] = cE; cf[2 * M * M + i * M + r] = cS; cf[3 * M * M + i * M + r] = cN; }
            double v = 0.0;
            if (r == c) {
                double d = -al * ((be[kr] + bw[kr]) / (dx * dx) + (bn[kr] + bs[kr]) / (dy * dy)) + lam[kr];
                if (i == 0) d -= cW;
                if (i == M - 1) d -= cE;
                if (r == 0) d -= cS;
                if (r == M - 1) d -= cN;
                v = d;
            } else if (c == r - 1) v = cS;
            else if (c == r + 1) v = cN;
            if (i > 0) v -= cW * prev[q] * cE_prev;
            a[q] = v;
        }
        // Gauss-Jordan inverse of the block, in registers
#pragma unroll
        for (int k = 0; k < M; k++) {
            const double rowk = __shfl_sync(0xffffffffu, a[k % RPL], (k / RPL) * M + c);   // a[k][c]
            const double piv = __shfl_sync(0xffffffffu, rowk, gbase + k);                   // a[k][k]
            const double p = 1.0 / piv;
            minp = fmin(minp, fabs(piv));
#pragma unroll
            for (int q = 0; q < RPL; q++) {
                const int r = rg * RPL + q;
                const double f = __shfl_sync(0xffffffffu, a[q], gbase + k);               // a[r][k]
                if (r == k) a[q] = (c == k) ? p : a[q] * p;
                else a[q] = (c == k) ? -f * p : a[q] - f * p * rowk;
            }
        }
#pragma unroll
        for (int q = 0; q < RPL; q++) {
            P_all[((size_t)leaf * M + i) * M * M + (rg * RPL + q) * M + c] = a[q];
            prev[q] = a[q];
        }
        {   // cE[i][c] for the next block row
            const size_t kc = cell0 + (size_t)i * M + c;
            cE_prev = alpha[kc] * be[kc] / (dx * dx);
        }
    }
    if (lane == 0 && min_pivot) atomicMin(reinterpret_cast<unsigned long long*>(min_pivot), (unsigned long long)__double_as_longlong(minp));
}

template <int M>
__global__ void __launch_bounds__(4 * M * 4)
leaf_var_dtn_mma_kernel(const double* __restrict__ coef, const double* __restrict__ P_all, const double* __restrict__ boxes,
                        const int* __restrict__ leaf_nodes, double* __restrict__ T_all)
{
    constexpr int NW = 4 * M / 8, NT = NW * 32, MT = M / 8, KS = M / 4;
    constexpr int SP = M + 4;     // == 4 or 12 (mod 16): A-fragment reads of a half warp (4 rows x 4 k) fall into 16 distinct bank pairs
    constexpr int SY = 12;        // y tile M x 8, row stride 12: B-fragment reads (4 k x 4 n per half warp) conflict free
    __shared__ __align__(16) double sP[2][M * SP];
    __shared__ __align__(16) double sY[NW][M * SY];
    __shared__ double sC[4][M * M];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int leaf = blockIdx.x;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    const double* cfl = coef + (size_t)leaf * 4 * M * M;
    for (int e = tid; e < 4 * M * M; e += NT) sC[e / (M * M)][e % (M * M)] = cfl[e];
    const double* P = P_all + (size_t)leaf * M * M * M;
    const int col0 = warp * 8, side = col0 / M, t0 = col0 % M;    // this warp's 8 columns lie on one side: t = t0 + c
    const int fr = lane >> 2, fc = 2 * (lane & 3);                  // accumulator fragment: row fr (+ 8 mt), columns fc, fc + 1
    double* sy = sY[warp];
    double z[M][MT][2];
    constexpr bool HALF = (M * M) < NT;                             // M = 8: 64 elements of P_i, 128 threads
    const bool loader = !HALF || tid < M * M;
    double pnext = loader ? P[tid] : 0.0;                           // P_0
    __syncthreads();                                                // sC
    // ---- forward sweep: y_i = rhs_i - W_i z_{i-1},  z_i = P_i y_i
#pragma unroll
    for (int i = 0; i < M; i++) {
        if (loader) sP[i & 1][(tid / M) * SP + (tid % M)] = pnext;
        if (i + 1 < M) { if (loader) pnext = P[(size_t)(i + 1) * M * M + tid]; }
        else if (loader) pnext = P[(size_t)(M - 2) * M * M + tid];  // first block of the backward sweep
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int j = 8 * mt + fr;
            double y[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int t = t0 + fc + e;
                double v = 0.0;
                if (side == 0) { if (i == 0 && j == t) v = -2.0 * sC[0][j]; }
                else if (side == 1) { if (i == M - 1 && j == t) v = -2.0 * sC[1][(M - 1) * M + j]; }
                else if (side == 2) { if (j == 0 && i == t) v = -2.0 * sC[2][i * M]; }
                else { if (j == M - 1 && i == t) v = -2.0 * sC[3][i * M + M - 1]; }
                if (i > 0) v -= sC[0][i * M + j] * z[i > 0 ? i - 1 : 0][mt][e];
                y[e] = v;
            }
            *reinterpret_cast<double2*>(sy + j * SY + fc) = make_double2(y[0], y[1]);
        }
        __syncthreads();                                            // P_i staged (and, warp-locally, y_i)
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
                leaf_dmma(a0, a1, sP[i & 1][(8 * mt + fr) * SP + 4 * ks + (lane & 3)], sy[(4 * ks + (lane & 3)) * SY + fr]);
            z[i][mt][0] = a0; z[i][mt][1] = a1;
        }
        __syncwarp();                                               // y tile free for the next step
    }
    // ---- backward sweep: u_i = z_i - P_i (E_i u_{i+1})   (u overwrites z)
#pragma unroll
    for (int ii = 0; ii < M - 1; ii++) {
        const int i = M - 2 - ii, buf = ii & 1;                     // buffers alternate from the forward sweep's last (M-1)&1 = 1: ii = 0 -> 0
        if (loader) sP[buf][(tid / M) * SP + (tid % M)] = pnext;
        if (i > 0 && loader) pnext = P[(size_t)(i - 1) * M * M + tid];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int j = 8 * mt + fr;
            const double ce = sC[1][i * M + j];
            *reinterpret_cast<double2*>(sy + j * SY + fc) = make_double2(ce * z[i + 1][mt][0], ce * z[i + 1][mt][1]);
        }
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
                leaf_dmma(a0, a1, sP[buf][(8 * mt + fr) * SP + 4 * ks + (lane & 3)], sy[(4 * ks + (lane & 3)) * SY + fr]);
            z[i][mt][0] -= a0; z[i][mt][1] -= a1;
        }
        __syncwarp();
    }
    // ---- T[row][col] = sign(side_row) (2 / d) (u_edge - [row == col])   (FiniteVolumeSolver.cpp:332-343, :444-452)
    double* T = T_all + (size_t)leaf * 16 * M * M;
    const int colb = col0 + fc;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const int j = 8 * mt + fr;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int col = colb + e;
            {   // W row j: u(i = 0, j);  E row j: u(i = M-1, j)
                const int rw = j, re = M + j;
                T[(size_t)rw * (4 * M) + col] = (2.0 / dx) * (z[0][mt][e] - (rw == col ? 1.0 : 0.0));
                T[(size_t)re * (4 * M) + col] = -(2.0 / dx) * (z[M - 1][mt][e] - (re == col ? 1.0 : 0.0));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < M; i++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int col = colb + e;
            if (fr == 0) { const int rs = 2 * M + i; T[(size_t)rs * (4 * M) + col] = (2.0 / dy) * (z[i][0][e] - (rs == col ? 1.0 : 0.0)); }          // S: j = 0
            if (fr == 7) { const int rn = 3 * M + i; T[(size_t)rn * (4 * M) + col] = -(2.0 / dy) * (z[i][MT - 1][e] - (rn == col ? 1.0 : 0.0)); }  // N: j = M-1
        }
    }
}

// mode 0: u = solve(g, f); mode 1: h = mapD2N(0, f).  One warp per leaf, z in a per-warp shared tile (M^2 doubles).
template <int M>
__global__ void __launch_bounds__(256)
leaf_var_solve_warp_kernel(const double* __restrict__ coef, const double* __restrict__ P_all, const double* __restrict__ boxes,
                           const int* __restrict__ leaf_nodes, const double* __restrict__ f, double fscale, double* const* __restrict__ g_ptrs,
                           double* __restrict__ u_out, double* const* __restrict__ h_ptrs, int mode, int n_leaves)
{
    constexpr int RG = 32 / M, KP = M / RG;    // lane = (row r, part of the k range)
    __shared__ double sZ[8][M * M];
    __shared__ double sYv[8][M];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int leaf = blockIdx.x * 8 + w;
    if (leaf >= n_leaves) return;
    const int r = lane % M, part = lane / M;
    const double* box = boxes + 4 * (size_t)leaf_nodes[leaf];
    const double dx = (box[1] - box[0]) / M, dy = (box[3] - box[2]) / M;
    const double* cW = coef + (size_t)leaf * 4 * M * M;
    const double* cE = cW + M * M;
    const double* cS = cE + M * M;
    const double* cN = cS + M * M;
    const double* P = P_all + (size_t)leaf * M * M * M;
    const double* gl = (mode == 0 && g_ptrs) ? g_ptrs[leaf] : nullptr;
    const double* fl = f ? f + (size_t)leaf * M * M : nullptr;
    double* z = sZ[w];
    double* yv = sYv[w];
    // this lane's slice of row r of P_i, fetched one step ahead of its use (the sweeps are a chain of 2 M dependent steps: without
    // the prefetch every step waits for a round trip to HBM - ncu r2d: long-scoreboard stalls 138 per issue, 27 % of DRAM peak)
    double pc[KP], pn[KP];
    auto fetch = [&](int i, double (&dst)[KP]) {
        const double* pr = P + (size_t)i * M * M + r * M + part * KP;
#pragma unroll
        for (int k = 0; k < KP; k += 2) { const double2 v = *reinterpret_cast<const double2*>(pr + k); dst[k] = v.x; dst[k + 1] = v.y; }
    };
    auto matvec = [&](const double (&pv)[KP]) -> double {       // row r of P_i y, complete in every lane of the row
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < KP; k++) s = fma(pv[k], yv[part * KP + k], s);
#pragma unroll
        for (int o = M; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        return s;
    };
    fetch(0, pc);
    for (int i = 0; i < M; i++) {
        fetch(i + 1 < M ? i + 1 : M - 2, pn);          // after the last forward block: the first block of the backward sweep
        if (part == 0) {
            const int j = r;
            double v = fl ? fscale * fl[i * M + j] : 0.0;
            if (gl) {
                if (i == 0) v += -2.0 * cW[j] * gl[j];
                if (i == M - 1) v += -2.0 * cE[(M - 1) * M + j] * gl[M + j];
                if (j == 0) v += -2.0 * cS[i * M] * gl[2 * M + i];
                if (j == M - 1) v += -2.0 * cN[i * M + M - 1] * gl[3 * M + i];
            }
            if (i > 0) v -= cW[i * M + j] * z[(i - 1) * M + j];
            yv[j] = v;
        }
        __syncwarp();
        const double s = matvec(pc);
        __syncwarp();
        if (part == 0) z[i * M + r] = s;
#pragma unroll
        for (int k = 0; k < KP; k++) pc[k] = pn[k];
    }
    for (int i = M - 2; i >= 0; i--) {
        if (i > 0) fetch(i - 1, pn);
        __syncwarp();
        if (part == 0) yv[r] = cE[i * M + r] * z[(i + 1) * M + r];
        __syncwarp();
        const double s = matvec(pc);
        if (part == 0) z[i * M + r] -= s;
#pragma unroll
        for (int k = 0; k < KP; k++) pc[k] = pn[k];
    }
    __syncwarp();
    if (mode == 0) {
        for (int e = lane; e < M * M; e += 32) u_out[(size_t)leaf * M * M + e] = z[e];
    } else {
        double* h = h_ptrs[leaf];
        for (int e = lane; e < 4 * M; e += 32) {
            const int side = e / M, t = e % M;
            double v;
            if (side == 0) v = (2.0 / dx) * z[0 * M + t];
            else if (side == 1) v = -(2.0 / dx) * z[(M - 1) * M + t];
            else if (side == 2) v = (2.0 / dy) * z[t * M + 0];
            else v = -(2.0 / dy) * z[t * M + M - 1];
            h[e] = v;
        }
    }
}

template <int M>
static void var_factor_M(const double* const* cin, const double* boxes, const int* leaf_nodes, double* coef, double* P, double* minpiv, int n, cudaStream_t s) {
    if constexpr (M == 8 || M == 16) {
        if (get_tuning(6) == 0) {   // default: one warp per leaf, Gauss-Jordan by shuffles
            leaf_var_factor_warp_kernel<M><<<(n + 7) / 8, 256, 0, s>>>(cin[0], cin[1], cin[2], cin[3], cin[4], cin[5], boxes, leaf_nodes, coef, P, minpiv, n);
            return;
        }
    }
    leaf_var_factor_kernel<M><<<n, M * M, 0, s>>>(cin[0], cin[1], cin[2], cin[3], cin[4], cin[5], boxes, leaf_nodes, coef, P, minpiv);
}
template <int M>
static void var_solve_M(const double* coef, const double* P, const double* boxes, const int* leaf_nodes, const double* f, double fscale,
                        double* const* g_ptrs, double* u_out, double* const* h_ptrs, double* T_all, int mode, int n_leaves, cudaStream_t s) {
    if constexpr (M == 8 || M == 16) {
        if (get_tuning(6) == 0) {   // default: FP64 tensor-core DtN kernel / warp-per-leaf single right-hand side
            if (mode == 2) leaf_var_dtn_mma_kernel<M><<<n_leaves, 4 * M * 4, 0, s>>>(coef, P, boxes, leaf_nodes, T_all);
            else leaf_var_solve_warp_kernel<M><<<(n_leaves + 7) / 8, 256, 0, s>>>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves);
            return;
        }
    }
    const int C = mode == 2 ? (M >= 32 ? 16 : (4 * M < 32 ? 4 * M : 32)) : 1;
    const int smem = (M * (M + 1) + M * C + M * M * C) * (int)sizeof(double);
    auto kern = leaf_var_solve_kernel<M>;
    if (smem > 48 * 1024)   // per device, and the size depends on the mode: set on every such launch (a host-side table write)
        EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(n_leaves, mode == 2 ? (4 * M) / C : 1);
    kern<<<grid, 256, smem, s>>>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, T_all, mode, C);
}

void launch_leaf_var_factor(int M, const double* const* coef_in, const double* boxes, const int* leaf_nodes, double* coef, double* P,
                            double* min_pivot, int n_leaves, cudaStream_t s)
{
    if (n_leaves == 0) return;
    switch (M) {
        case 4: var_factor_M<4>(coef_in, boxes, leaf_nodes, coef, P, min_pivot, n_leaves, s); break;
        case 8: var_factor_M<8>(coef_in, boxes, leaf_nodes, coef, P, min_pivot, n_leaves, s); break;
        case 16: var_factor_M<16>(coef_in, boxes, leaf_nodes, coef, P, min_pivot, n_leaves, s); break;
        case 24: var_factor_M<24>(coef_in, boxes, leaf_nodes, coef, P, min_pivot, n_leaves, s); break;
        case 32: var_factor_M<32>(coef_in, boxes, leaf_nodes, coef, P, min_pivot, n_leaves, s); break;
        default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 4, 8, 16, 24, 32 or (constant coefficients) 64 cells per side"};
    }
    EF_CUDA(cudaGetLastError());
}

void launch_leaf_var_solve(int M, const double* coef, const double* P, const double* boxes, const int* leaf_nodes, const double* f, double fscale,
                           double* const* g_ptrs, double* u_out, double* const* h_ptrs, double* T_all, int mode, int n_leaves, cudaStream_t s)
{
    if (n_leaves == 0) return;
    switch (M) {
        case 4: var_solve_M<4>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, T_all, mode, n_leaves, s); break;
        case 8: var_solve_M<8>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, T_all, mode, n_leaves, s); break;
        case 16: var_solve_M<16>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, T_all, mode, n_leaves, s); break;
        case 24: var_solve_M<24>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, T_all, mode, n_leaves, s); break;
        case 32: var_solve_M<32>(coef, P, boxes, leaf_nodes, f, fscale, g_ptrs, u_out, h_ptrs, T_all, mode, n_leaves, s); break;
        default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 4, 8, 16, 24, 32 or (constant coefficients) 64 cells per side"};
    }
    EF_CUDA(cudaGetLastError());
}

void launch_broadcast_leaf_T(double* T_all, int M, int n_leaves, cudaStream_t s)
{
    if (n_leaves <= 1) return;
    broadcast_leaf_T_kernel<<<148 * 8, 256, 0, s>>>(T_all, (size_t)16 * M * M, n_leaves);
    EF_CUDA(cudaGetLastError());
}

template <int M>
static void dtn_const_M(const double* Q, const double* boxes, const int* leaf_nodes, double lambda, double* T_all, const int* build_list, int n_build, cudaStream_t s) {
    if constexpr (M * M <= 1024) leaf_dtn_const_kernel<M><<<n_build, M * M, 0, s>>>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build);
    else {
        constexpr int smem = (4 * M * (M + 1) + M) * (int)sizeof(double);
        auto kern = leaf_dtn_const_loop_kernel<M, 1024>;
        static unsigned long long prepared = 0;
        if (first_use_on_device(prepared)) EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<n_build, 1024, smem, s>>>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build);
    }
}
template <int M>
static void solve_const_M(const double* Q, const double* boxes, const int* leaf_nodes, double lambda, const double* f, double fscale,
                          double* const* g_ptrs, double* u_out, double* const* h_ptrs, int mode, int n_leaves, cudaStream_t s) {
    if constexpr (M * M <= 1024) leaf_solve_const_kernel<M><<<n_leaves, M * M, 0, s>>>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves);
    else {
        constexpr int smem = (3 * M * (M + 1) + 5 * M) * (int)sizeof(double);
        auto kern = leaf_solve_const_loop_kernel<M, 1024>;
        static unsigned long long prepared = 0;
        if (first_use_on_device(prepared)) EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<n_leaves, 1024, smem, s>>>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves);
    }
}

void launch_leaf_dtn_const(int M, const double* Q, const double* boxes, const int* leaf_nodes, double lambda,
                           double* T_all, int n_leaves, bool cache_operators, const int* build_list, int n_build, const int* leaf_src,
                           cudaStream_t s)
{
    if (n_leaves == 0) return;
    if (cache_operators) { build_list = nullptr; n_build = 1; }   // quirk q1: the first leaf's T for every leaf
    else if (!build_list) n_build = n_leaves;
    switch (M) {
        case 4: dtn_const_M<4>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build, s); break;
        case 64: dtn_const_M<64>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build, s); break;
        case 8: dtn_const_M<8>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build, s); break;
        case 16: dtn_const_M<16>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build, s); break;
        case 24: dtn_const_M<24>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build, s); break;
        case 32: dtn_const_M<32>(Q, boxes, leaf_nodes, lambda, T_all, build_list, n_build, s); break;
        default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 4, 8, 16, 24, 32 or (constant coefficients) 64 cells per side"};
    }
    EF_CUDA(cudaGetLastError());
    if (cache_operators && n_leaves > 1) {
        broadcast_leaf_T_kernel<<<148 * 8, 256, 0, s>>>(T_all, (size_t)16 * M * M, n_leaves);
        EF_CUDA(cudaGetLastError());
    } else if (build_list && n_build < n_leaves) {
        const int e2 = 8 * M * M, bx = (e2 + 255) / 256 < 4 ? (e2 + 255) / 256 : 4;
        copy_leaf_T_kernel<<<dim3(bx, n_leaves < 148 * 16 ? n_leaves : 148 * 16), 256, 0, s>>>(T_all, e2, leaf_src, n_leaves);
        EF_CUDA(cudaGetLastError());
    }
}

void launch_leaf_solve_const(int M, const double* Q, const double* boxes, const int* leaf_nodes, double lambda,
                             const double* f, double fscale, double* const* g_ptrs, double* u_out, double* const* h_ptrs,
                             int mode, int n_leaves, cudaStream_t s)
{
    if (n_leaves == 0) return;
    // default: FP64 tensor-core kernel, one warp per leaf (its bulk copies need f on a 16-byte boundary)
    if (get_tuning(3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 && M != 4 && M != 64) {   // (4: no 8 x 8 tiles; 64: the result does not fit a warp's registers)
        switch (M) {
            case 8: solve_const_mma_M<8>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
            case 16: solve_const_mma_M<16>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
            case 24: solve_const_mma_M<24>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
            case 32: solve_const_mma_M<32>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
            default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 4, 8, 16, 24, 32 or (constant coefficients) 64 cells per side"};
        }
        EF_CUDA(cudaGetLastError());
        return;
    }
    switch (M) {   // one thread per cell: any alignment of f (a caller-owned device pointer in efgpu_upwards_device)
        case 4: solve_const_M<4>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 64: solve_const_M<64>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 8: solve_const_M<8>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 16: solve_const_M<16>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 24: solve_const_M<24>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 32: solve_const_M<32>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 4, 8, 16, 24, 32 or (constant coefficients) 64 cells per side"};
    }
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
