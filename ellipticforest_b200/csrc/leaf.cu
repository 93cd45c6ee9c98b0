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
                      double lambda, double* __restrict__ T_all, int n_build)
{
    __shared__ double sQ[M][M + 1];     // sQ[i][k] = q_{k+1}(i)
    __shared__ double sDinv[M][M + 1];  // 1 / (mu_k/dx^2 + mu_l/dy^2 + lambda), [k][l]
    __shared__ double sZ[M][M + 1];
    __shared__ double sP[M][M + 1];
    __shared__ double sMu[M];
    const int leaf = blockIdx.x;
    if (leaf >= n_build) return;
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

template <int M>
static void dtn_const_M(const double* Q, const double* boxes, const int* leaf_nodes, double lambda, double* T_all, int n_build, cudaStream_t s) {
    leaf_dtn_const_kernel<M><<<n_build, M * M, 0, s>>>(Q, boxes, leaf_nodes, lambda, T_all, n_build);
}
template <int M>
static void solve_const_M(const double* Q, const double* boxes, const int* leaf_nodes, double lambda, const double* f, double fscale,
                          double* const* g_ptrs, double* u_out, double* const* h_ptrs, int mode, int n_leaves, cudaStream_t s) {
    leaf_solve_const_kernel<M><<<n_leaves, M * M, 0, s>>>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves);
}

void launch_leaf_dtn_const(int M, const double* Q, const double* boxes, const int* leaf_nodes, double lambda,
                           double* T_all, int n_leaves, bool cache_operators, cudaStream_t s)
{
    if (n_leaves == 0) return;
    const int n_build = cache_operators ? 1 : n_leaves;
    switch (M) {
        case 8: dtn_const_M<8>(Q, boxes, leaf_nodes, lambda, T_all, n_build, s); break;
        case 16: dtn_const_M<16>(Q, boxes, leaf_nodes, lambda, T_all, n_build, s); break;
        case 24: dtn_const_M<24>(Q, boxes, leaf_nodes, lambda, T_all, n_build, s); break;
        case 32: dtn_const_M<32>(Q, boxes, leaf_nodes, lambda, T_all, n_build, s); break;
        default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 8, 16, 24 or 32 cells per side"};
    }
    EF_CUDA(cudaGetLastError());
    if (cache_operators && n_leaves > 1) {
        broadcast_leaf_T_kernel<<<148 * 8, 256, 0, s>>>(T_all, (size_t)16 * M * M, n_leaves);
        EF_CUDA(cudaGetLastError());
    }
}

void launch_leaf_solve_const(int M, const double* Q, const double* boxes, const int* leaf_nodes, double lambda,
                             const double* f, double fscale, double* const* g_ptrs, double* u_out, double* const* h_ptrs,
                             int mode, int n_leaves, cudaStream_t s)
{
    if (n_leaves == 0) return;
    switch (M) {
        case 8: solve_const_M<8>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 16: solve_const_M<16>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 24: solve_const_M<24>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        case 32: solve_const_M<32>(Q, boxes, leaf_nodes, lambda, f, fscale, g_ptrs, u_out, h_ptrs, mode, n_leaves, s); break;
        default: throw Error{EF_ERR_UNSUPPORTED, "leaf patches must be 8, 16, 24 or 32 cells per side"};
    }
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
