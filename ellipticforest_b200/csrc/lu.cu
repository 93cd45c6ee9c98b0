// Dense LU with partial pivoting for the root boundary system of solveStage
// (src/HPSAlgorithm.hpp:402-420):   g = (diag(a) + diag(b) T_root)^-1 (r - b .* h_root),
// which the reference forms with two dense diagonal matrices, a dgemm and a dgesv of order
// 4 * root size (Matrix.hpp:852,895).  Pure Dirichlet data (b == 0) never reaches this file
// (g = r / a); mixed Dirichlet/Neumann sides (examples/thermal/main.cpp:323-349) and Robin data do.
//
// The matrix is held TRANSPOSED (Mt = A^T row-major, i.e. A column-major, ld = N), so pivot
// searches and column updates stream contiguous memory.  Blocked right-looking factorisation,
// panel width NB in {64, 32}:
//   panel_lu_kernel     cooperative launch, one slice of rows per CTA, two grid syncs per column:
//                       pivot search (idamax semantics: LAPACK partial pivoting), row swap, scale,
//                       rank-1 update of the remaining panel columns
//   swap_rows_kernel    the panel's NB interchanges applied to all other columns and to the rhs
//   tri_inv_kernel      (L11)^-1 of the unit lower triangle, so that the triangular solve of the
//                       row block becomes a GEMM on the FP64 tensor pipe
//   bgemm (gemm.cu)     U12 = L11^-1 A12  and the trailing update  A22 -= L21 U12
//   rhs kernels         forward substitution alongside the factorisation, blocked back substitution
#include "common.cuh"
#include "kernels.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace efgpu {

// Mt[j][i] = b[i] * T[i][j] + (i == j) a[i];   rhs[i] = r[i] - b[i] * h[i]
__global__ void robin_assemble_kernel(const double* __restrict__ T, const double* __restrict__ a, const double* __restrict__ b,
                                      const double* __restrict__ r, const double* __restrict__ h, double* __restrict__ Mt,
                                      double* __restrict__ rhs, int N)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int i = i0 + k, j = j0 + tx;
        tile[k][tx] = b[i] * T[(size_t)i * N + j] + (i == j ? a[i] : 0.0);
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) Mt[(size_t)(j0 + k) * N + i0 + tx] = tile[tx][k];
    if (blockIdx.x == 0 && ty == 0) { const int i = i0 + tx; rhs[i] = r[i] - b[i] * (h ? h[i] : 0.0); }
}

struct PivotCand { double val; int idx; int pad; };

__device__ __forceinline__ void better(double& v, int& i, double v2, int i2)
{
    if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }   // first maximum, as idamax
}

__device__ void block_argmax(double v, int i, PivotCand* out, double* sv, int* si)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        int i2 = __shfl_xor_sync(0xffffffffu, i, o);
        better(v, i, v2, i2);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sv[w] = v; si[w] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) better(v, i, sv[k], si[k]);
        out->val = v; out->idx = i;
    }
    __syncthreads();
}

// Factorises columns [k0, k0+NB) of A (rows of Mt), rows [k0, N).  ipiv[k0+jj] = pivot row (global index).
__global__ void __launch_bounds__(256)
panel_lu_kernel(double* __restrict__ Mt, int N, int k0, int NB, int* __restrict__ ipiv, PivotCand* __restrict__ cand, int* __restrict__ info)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sm[];
    double* u = sm;          // pivot row inside the panel
    double* v = sm + NB;     // row k0+jj before the interchange
    __shared__ double sv[8];
    __shared__ int si[8];
    const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x;
    const int rows = N - k0, per = (rows + G - 1) / G;
    const int lo = k0 + g * per, hi = min(N, lo + per);
    auto col = [&](int c) { return Mt + (size_t)(k0 + c) * N; };

    // candidates of the first column
    {
        double bv = -1.0; int bi = 0x7fffffff;
        const double* c0 = col(0);
        for (int i = lo + tid; i < hi; i += blockDim.x) better(bv, bi, fabs(c0[i]), i);
        block_argmax(bv, bi, cand + g, sv, si);
    }
    for (int jj = 0; jj < NB; jj++) {
        const int cj = k0 + jj;
        grid.sync();
        double pv = -1.0; int p = 0x7fffffff;
        for (int k = 0; k < G; k++) better(pv, p, cand[k].val, cand[k].idx);
        if (p == 0x7fffffff) p = cj;
        for (int c = tid; c < NB; c += blockDim.x) { u[c] = col(c)[p]; v[c] = col(c)[cj]; }
        if (g == 0 && tid == 0) { ipiv[cj] = p; if (pv == 0.0) atomicCAS(info, 0, cj + 1); }
        __syncthreads();
        grid.sync();
        if (p != cj) {
            if (p >= lo && p < hi) for (int c = tid; c < NB; c += blockDim.x) col(c)[p] = v[c];
            if (cj >= lo && cj < hi) for (int c = tid; c < NB; c += blockDim.x) col(c)[cj] = u[c];
        }
        __syncthreads();
        const double piv = u[jj];
        const double pinv = piv != 0.0 ? 1.0 / piv : 0.0;
        double bv = -1.0; int bi = 0x7fffffff;
        double* cjp = col(jj);
        for (int i = max(lo, cj + 1) + tid; i < hi; i += blockDim.x) {
            const double l = cjp[i] * pinv;
            cjp[i] = l;
            for (int c = jj + 1; c < NB; c++) {
                double* cc = col(c);
                const double nv = cc[i] - l * u[c];
                cc[i] = nv;
                if (c == jj + 1) better(bv, bi, fabs(nv), i);
            }
        }
        if (jj + 1 < NB) block_argmax(bv, bi, cand + g, sv, si);
    }
}

// interchanges of one panel applied to the columns outside it (thread per column) and to the rhs
__global__ void swap_rows_kernel(double* __restrict__ Mt, double* __restrict__ rhs, int N, int k0, int NB, const int* __restrict__ ipiv)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > N) return;
    double* x;
    if (c == N) x = rhs;
    else { if (c >= k0 && c < k0 + NB) return; x = Mt + (size_t)c * N; }
    for (int jj = 0; jj < NB; jj++) {
        const int p = ipiv[k0 + jj];
        if (p != k0 + jj) { const double t = x[k0 + jj]; x[k0 + jj] = x[p]; x[p] = t; }
    }
}

// LinvT[k][n] = (L11^-1)[n][k], L11 = unit lower triangle of the panel's diagonal block; also
// y = L11^-1 rhs[k0 .. k0+NB) in place.  One CTA, NB threads: thread n owns row n of L11^-1.
__global__ void tri_inv_kernel(const double* __restrict__ Mt, int N, int k0, int NB, double* __restrict__ LinvT, double* __restrict__ rhs)
{
    extern __shared__ __align__(16) double sm[];
    double* L = sm;                  // NB x (NB+1): L[i][j]
    double* X = sm + NB * (NB + 1);  // NB x (NB+1): X[i][j] = (L^-1)[i][j]
    const int LD = NB + 1, tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int j = e / NB, i = e % NB;   // A[k0+i][k0+j] = Mt[(k0+j)*N + k0+i]
        L[i * LD + j] = Mt[(size_t)(k0 + j) * N + k0 + i];
        X[i * LD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    // column j of the inverse by forward substitution; thread = column
    if (tid < NB) {
        const int j = tid;
        for (int i = j + 1; i < NB; i++) {
            double s = 0.0;
            for (int k = j; k < i; k++) s += L[i * LD + k] * X[k * LD + j];
            X[i * LD + j] = -s;
        }
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int k = e / NB, n = e % NB;
        LinvT[e] = X[n * LD + k];
    }
    // rhs block
    double s = 0.0;
    if (tid < NB) for (int k = 0; k <= tid; k++) s += X[tid * LD + k] * rhs[k0 + k];
    __syncthreads();
    if (tid < NB) rhs[k0 + tid] = s;
}

// rhs[i] -= sum_j A[i][k0+j] y[j]  for i in [i_lo, i_hi), y = rhs[k0 .. k0+NB)   (A[i][k0+j] = Mt[(k0+j)*N + i])
__global__ void rhs_update_kernel(const double* __restrict__ Mt, double* __restrict__ rhs, int N, int k0, int NB, int i_lo, int i_hi)
{
    extern __shared__ __align__(16) double y[];
    for (int j = threadIdx.x; j < NB; j += blockDim.x) y[j] = rhs[k0 + j];
    __syncthreads();
    const int i = i_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i_hi) return;
    double s = 0.0;
    for (int j = 0; j < NB; j++) s = fma(Mt[(size_t)(k0 + j) * N + i], y[j], s);
    rhs[i] -= s;
}

// x = U11^-1 y for the diagonal block at k0 (upper triangle incl. diagonal), in place in rhs.  One CTA.
__global__ void tri_solve_upper_kernel(const double* __restrict__ Mt, double* __restrict__ rhs, int N, int k0, int NB)
{
    extern __shared__ __align__(16) double sm[];
    double* U = sm;                 // NB x (NB+1)
    double* x = sm + NB * (NB + 1);
    const int LD = NB + 1, tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int j = e / NB, i = e % NB;
        U[i * LD + j] = Mt[(size_t)(k0 + j) * N + k0 + i];
    }
    for (int j = tid; j < NB; j += blockDim.x) x[j] = rhs[k0 + j];
    __syncthreads();
    for (int j = NB - 1; j >= 0; j--) {
        if (tid == 0) x[j] = x[j] / U[j * LD + j];
        __syncthreads();
        if (tid < j) x[tid] -= U[tid * LD + j] * x[j];
        __syncthreads();
    }
    for (int j = tid; j < NB; j += blockDim.x) rhs[k0 + j] = x[j];
}

static int pick_nb(int N) { return N % 64 == 0 ? 64 : 32; }   // tri_inv_kernel keeps L11 and its inverse in shared memory: NB <= 64

size_t robin_workspace_doubles(int N)
{
    const int NB = pick_nb(N);
    return (size_t)N * N + (size_t)N + (size_t)NB * NB + 1024;
}

// Solves (diag(a) + diag(b) T) g = r - b .* h on the device; result in g_out (N doubles).  `ws` holds
// robin_workspace_doubles(N) doubles.  Returns LAPACK-style info through *info_host (0 = ok, k = zero pivot at column k).
void robin_solve(const double* T, const double* a, const double* b, const double* r, const double* h, int N, double* ws,
                 double* g_out, int* info_host, cudaStream_t s)
{
    if (N % 32) throw Error{EF_ERR_BAD_SHAPE, "root boundary system: size must be a multiple of 32"};
    const int NB = pick_nb(N), np = N / NB;
    double* Mt = ws;
    double* rhs = ws + (size_t)N * N;
    double* LinvT = rhs + N;
    int dev = 0, nsm = 0, coop = 0;
    EF_CUDA(cudaGetDevice(&dev));
    EF_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    EF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) throw Error{EF_ERR_UNSUPPORTED, "device lacks cooperative launch"};
    // small device scratch: ipiv, candidates, info
    int* ipiv = nullptr; PivotCand* cand = nullptr; int* info = nullptr;
    EF_CUDA(cudaMallocAsync((void**)&ipiv, sizeof(int) * N, s));
    EF_CUDA(cudaMallocAsync((void**)&cand, sizeof(PivotCand) * nsm, s));
    EF_CUDA(cudaMallocAsync((void**)&info, sizeof(int), s));
    EF_CUDA(cudaMemsetAsync(info, 0, sizeof(int), s));

    robin_assemble_kernel<<<dim3(N / 32, N / 32), dim3(32, 8), 0, s>>>(T, a, b, r, h, Mt, rhs, N);
    EF_CUDA(cudaGetLastError());

    // GEMM descriptors of every panel: [2p] U12 = L11^-1 A12 (in place, one tile across), [2p+1] A22 -= L21 U12
    std::vector<GemmBlock> blocks;
    for (int p = 0; p + 1 < np; p++) {
        const long long k0 = (long long)p * NB, rest = N - k0 - NB;
        GemmBlock t{};
        t.c_op = 0; t.c_off = (k0 + NB) * N + k0; t.ldc = N; t.c0_op = -1; t.rows = (int)rest; t.cols = NB; t.nterms = 1;
        t.t[0] = GemmTerm{0, 1, N, NB, (k0 + NB) * N + k0, 0, NB, 0u};
        blocks.push_back(t);
        GemmBlock u{};
        u.c_op = 0; u.c_off = (k0 + NB) * N + k0 + NB; u.ldc = N; u.c0_op = 0; u.c0_off = u.c_off; u.ldc0 = N;
        u.rows = (int)rest; u.cols = (int)rest; u.nterms = 1;
        u.t[0] = GemmTerm{0, 0, N, N, (k0 + NB) * N + k0, k0 * N + k0 + NB, NB, 0x80000000u};
        blocks.push_back(u);
    }
    GemmBlock* d_blocks = nullptr; double** d_ptab = nullptr;
    double* h_ptab[2] = {Mt, LinvT};
    EF_CUDA(cudaMallocAsync((void**)&d_ptab, sizeof(h_ptab), s));
    EF_CUDA(cudaMemcpyAsync(d_ptab, h_ptab, sizeof(h_ptab), cudaMemcpyHostToDevice, s));
    if (!blocks.empty()) {
        EF_CUDA(cudaMallocAsync((void**)&d_blocks, sizeof(GemmBlock) * blocks.size(), s));
        EF_CUDA(cudaMemcpyAsync(d_blocks, blocks.data(), sizeof(GemmBlock) * blocks.size(), cudaMemcpyHostToDevice, s));
    }
    const int tri_smem = 2 * NB * (NB + 1) * (int)sizeof(double);
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared)) {
        EF_CUDA(cudaFuncSetAttribute(tri_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 65 * 8));
        EF_CUDA(cudaFuncSetAttribute(tri_solve_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (64 * 65 + 64) * 8));
    }
    for (int p = 0; p < np; p++) {
        int k0 = p * NB;
        int G = (N - k0 + 255) / 256; if (G > nsm) G = nsm; if (G < 1) G = 1;
        int nb = NB, n = N;
        void* args[] = {&Mt, &n, &k0, &nb, &ipiv, &cand, &info};
        EF_CUDA(cudaLaunchCooperativeKernel((void*)panel_lu_kernel, dim3(G), dim3(256), args, 2 * NB * sizeof(double), s));
        swap_rows_kernel<<<(N + 1 + 255) / 256, 256, 0, s>>>(Mt, rhs, N, k0, NB, ipiv);
        tri_inv_kernel<<<1, 256, tri_smem, s>>>(Mt, N, k0, NB, LinvT, rhs);
        EF_CUDA(cudaGetLastError());
        const int rest = N - k0 - NB;
        if (rest > 0) {
            rhs_update_kernel<<<(rest + 255) / 256, 256, NB * sizeof(double), s>>>(Mt, rhs, N, k0, NB, k0 + NB, N);
            launch_bgemm(d_ptab, 2, d_blocks + 2 * p, blocks.data() + 2 * p, 1, 1, s, NB);
            launch_bgemm(d_ptab, 2, d_blocks + 2 * p + 1, blocks.data() + 2 * p + 1, 1, 1, s);
        }
    }
    // back substitution, last panel first
    for (int p = np - 1; p >= 0; p--) {
        const int k0 = p * NB;
        tri_solve_upper_kernel<<<1, 256, (NB * (NB + 1) + NB) * sizeof(double), s>>>(Mt, rhs, N, k0, NB);
        if (k0 > 0) rhs_update_kernel<<<(k0 + 255) / 256, 256, NB * sizeof(double), s>>>(Mt, rhs, N, k0, NB, 0, k0);
    }
    EF_CUDA(cudaGetLastError());
    EF_CUDA(cudaMemcpyAsync(g_out, rhs, sizeof(double) * N, cudaMemcpyDeviceToDevice, s));
    EF_CUDA(cudaMemcpyAsync(info_host, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    EF_CUDA(cudaFreeAsync(ipiv, s)); EF_CUDA(cudaFreeAsync(cand, s)); EF_CUDA(cudaFreeAsync(info, s));
    EF_CUDA(cudaFreeAsync(d_ptab, s));
    if (d_blocks) EF_CUDA(cudaFreeAsync(d_blocks, s));
    EF_CUDA(cudaStreamSynchronize(s));
}

}  // namespace efgpu
