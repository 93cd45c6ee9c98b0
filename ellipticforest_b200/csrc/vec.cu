// Bandwidth-bound kernels of the merge / upwards / solve stages: block assembly of X and the
// compact H, adaptive coarsening stencils, and the batched matrix-vector products.
//
// Reference call sites replaced (src/HPSAlgorithm.hpp):
//   mergeX_ :870-895 + createMatrixBlocks_ :799-859          -> assemble_X_kernel
//   mergeT_ :947-961 (H)                                      -> assemble_Hc_kernel (compact: the zero blocks are not stored)
//   coarsen_ :707-741 (two dense dgemm with L21/L12)          -> coarsen_T_kernel (2-tap / 3-tap stencils)
//   coarsenUpwards_ :1024-1048, uncoarsen_ :1165-1183         -> coarsen_h_kernel / uncoarsen_g_kernel
//   mergeW_ :1059-1087 (fresh dgesv), mergeH_ :1098-1117,
//   reorderOperatorsUpwards_ :1128-1150                       -> upwards kernels (cached X^-1, one GEMV each)
//   applyS_ :1194-1230                                        -> solve_split_kernel (GEMV + add w + scatter)
#include "common.cuh"
#include "kernels.cuh"

namespace efgpu {

// ---- tuning (efgpu_set_tuning): kernel-selection knobs for measurements ------------------------------
// [0] bulk-copy streaming matvec kernels on (1, default) / off (0); [1] long-row kernels: 1 = 8 loads in flight per lane
// (default 4); [2] CTAs per SM the long-row launcher aims for (default 16)
static int g_tuning[8] = {1, 0, 0, 0, 0, 0, 0, 0};
void set_tuning(int key, int value) { if (key >= 0 && key < 8) g_tuning[key] = value; }

// side of child c that faces interior interface k (-1: not adjacent)
__constant__ int c_iface[4][4] = {{3, -1, 1, -1}, {-1, 3, 0, -1}, {2, -1, -1, 1}, {-1, 2, -1, 0}};
// sign with which child c enters the jump across interface k (first child -, second +)
__constant__ double c_sgn[4][4] = {{-1, 0, -1, 0}, {0, -1, 1, 0}, {1, 0, 0, -1}, {0, 1, 0, 1}};
// exterior sides of child c, in the order of its tau index set
__constant__ int c_tau_side[4][2] = {{0, 2}, {1, 2}, {0, 3}, {1, 3}};
// the two interfaces adjacent to child c (ascending k)
__constant__ int c_kk[4][2] = {{0, 2}, {1, 2}, {0, 3}, {1, 3}};
// the two children adjacent to interface k
__constant__ int c_kids[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
// WESN block permutation (HPSAlgorithm.hpp:984): permuted position p holds pre-permutation block c_pi[p]
__constant__ int c_pi[8] = {0, 4, 2, 6, 1, 3, 5, 7};

// X[k][k'] = - sum_c sgn_c(k) T^c[iface_c(k), iface_c(k')]   (16 n x n blocks, 4 of them zero)
__global__ void assemble_X_kernel(const MergeEntry* __restrict__ ent, int n)
{
    const MergeEntry& e = ent[blockIdx.y];
    const int N = 4 * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * N; idx += gridDim.x * blockDim.x) {
        const int row = idx / N, col = idx % N;
        const int k = row / n, r = row % n, k2 = col / n, c = col % n;
        double v = 0.0;
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            const int sa = c_iface[ch][k], sb = c_iface[ch][k2];
            if (sa >= 0 && sb >= 0) v -= c_sgn[ch][k] * e.Tc[ch][(size_t)(sa * n + r) * N + sb * n + c];
        }
        e.Xinv[idx] = v;
        if (e.Xcopy) e.Xcopy[idx] = v;
    }
}

// Hc row block p (WESN position) of child c = [ T^c[side, iface_c(k0)] | T^c[side, iface_c(k1)] ]
__global__ void assemble_Hc_kernel(const MergeEntry* __restrict__ ent, int n)
{
    const MergeEntry& e = ent[blockIdx.y];
    const int N = 4 * n, W = 2 * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 8 * n * W; idx += gridDim.x * blockDim.x) {
        const int row = idx / W, col = idx % W;
        const int p = row / n, r = row % n, t = col / n, c = col % n;
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        const int sb = c_iface[ch][c_kk[ch][t]];
        e.Hc[idx] = e.Tc[ch][(size_t)(side * n + r) * N + sb * n + c];
    }
}

// dense H in the reference's (pre-permutation) layout, for parity checks only
__global__ void expand_H_kernel(const double* __restrict__ Hc, int n, double* __restrict__ H)
{
    const int W = 4 * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 8 * n * W; idx += gridDim.x * blockDim.x) {
        const int row = idx / W, col = idx % W;
        const int q = row / n, r = row % n, k = col / n, c = col % n;
        const int ch = q >> 1;
        int p = 0;
        for (int pp = 0; pp < 8; pp++) if (c_pi[pp] == q) p = pp;
        double v = 0.0;
        for (int t = 0; t < 2; t++) if (c_kk[ch][t] == k) v = Hc[(size_t)(p * n + r) * (2 * n) + t * n + c];
        H[idx] = v;
    }
}

// ---- interpolation stencils (src/SpecialMatrices.hpp:93-173) -----------------------------------
// column jc of L12(nfine x nc): up to 6 non-zero rows
__device__ __forceinline__ int l12_col(int jc, int nc, int* rows, double* coef)
{
    const int nf = 2 * nc;
    int m = 0;
    if (jc <= 2) { rows[m] = 0; coef[m++] = jc == 0 ? 1.40625 : (jc == 1 ? -0.5625 : 0.15625); }
    if (jc >= 1) { rows[m] = 2 * jc - 1; coef[m++] = 0.25; rows[m] = 2 * jc; coef[m++] = 0.75; }
    if (jc <= nc - 2) { rows[m] = 2 * jc + 1; coef[m++] = 0.75; rows[m] = 2 * jc + 2; coef[m++] = 0.25; }
    if (jc >= nc - 3) { rows[m] = nf - 1; coef[m++] = jc == nc - 1 ? 1.40625 : (jc == nc - 2 ? -0.5625 : 0.15625); }
    // rows 2jc (jc >= 1) and 2jc+1 (jc <= nc-2) never coincide with row 0 / nf-1 except through the
    // explicit edge rules above: row nf-1 = 2(nc-1)+1 is excluded by jc <= nc-2, row 0 by jc >= 1.
    return m;
}
// row i of L12: up to 3 non-zero columns
__device__ __forceinline__ int l12_row(int i, int nc, int* cols, double* coef)
{
    const int nf = 2 * nc;
    if (i == 0) { cols[0] = 0; cols[1] = 1; cols[2] = 2; coef[0] = 1.40625; coef[1] = -0.5625; coef[2] = 0.15625; return 3; }
    if (i == nf - 1) { cols[0] = nc - 3; cols[1] = nc - 2; cols[2] = nc - 1; coef[0] = 0.15625; coef[1] = -0.5625; coef[2] = 1.40625; return 3; }
    const int j = (i - 1) >> 1;
    cols[0] = j; cols[1] = j + 1;
    if (i & 1) { coef[0] = 0.75; coef[1] = 0.25; } else { coef[0] = 0.25; coef[1] = 0.75; }
    return 2;
}

// T_c = blkdiag4(L21) * T * blkdiag4(L12)  for one coarsening step (nfine -> nfine/2 per side)
__global__ void coarsen_T_kernel(const CoarsenOp* __restrict__ ops)
{
    const CoarsenOp op = ops[blockIdx.y];
    const int nf = op.nfine, nc = nf / 2, NF = 4 * nf, NC = 4 * nc;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NC * NC; idx += gridDim.x * blockDim.x) {
        const int R = idx / NC, Cc = idx % NC;
        const int a = R / nc, rc = R % nc, b = Cc / nc, jc = Cc % nc;
        int rows[6]; double coef[6];
        const int m = l12_col(jc, nc, rows, coef);
        const double* r0 = op.src + (size_t)(a * nf + 2 * rc) * NF + b * nf;
        const double* r1 = r0 + NF;
        double v = 0.0;
        for (int t = 0; t < m; t++) v += coef[t] * (0.5 * r0[rows[t]] + 0.5 * r1[rows[t]]);
        op.dst[idx] = v;
    }
}
__global__ void coarsen_h_kernel(const CoarsenOp* __restrict__ ops)
{
    const CoarsenOp op = ops[blockIdx.y];
    const int nc = op.nfine / 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * nc; idx += gridDim.x * blockDim.x)
        op.dst[idx] = 0.5 * op.src[2 * idx] + 0.5 * op.src[2 * idx + 1];   // side blocks are contiguous
}
__global__ void uncoarsen_g_kernel(const CoarsenOp* __restrict__ ops)
{
    const CoarsenOp op = ops[blockIdx.y];
    const int nf = op.nfine, nc = nf / 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * nf; idx += gridDim.x * blockDim.x) {
        const int a = idx / nf, i = idx % nf;
        int cols[3]; double coef[3];
        const int m = l12_row(i, nc, cols, coef);
        double v = 0.0;
        for (int t = 0; t < m; t++) v += coef[t] * op.src[a * nc + cols[t]];
        op.dst[idx] = v;
    }
}

// ---- upwards -----------------------------------------------------------------------------------
// hd[k] = sum_c sgn_c(k) h^c[iface_c(k)]   (HPSAlgorithm.hpp:1062-1081)
__global__ void hdiff_kernel(const MergeEntry* __restrict__ ent, int n)
{
    const MergeEntry& e = ent[blockIdx.y];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * n; idx += gridDim.x * blockDim.x) {
        const int k = idx / n, r = idx % n;
        const int c1 = c_kids[k][0], c2 = c_kids[k][1];
        e.hd[idx] = e.hc[c2][c_iface[c2][k] * n + r] - e.hc[c1][c_iface[c1][k] * n + r];
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// row-per-warp dot product, 16-byte loads, streaming (no L1 allocation) on the matrix
template <int U = 4>
__device__ __forceinline__ double row_dot(const double* __restrict__ a, const double* __restrict__ x, int len, int lane)
{
    double s0 = 0.0, s1 = 0.0;
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const int len2 = len >> 1;
    int c = lane;
    for (; c + 32 * (U - 1) < len2; c += 32 * U) {   // U independent 16-byte loads in flight per lane
        double2 av[U];
#pragma unroll
        for (int u = 0; u < U; u++) av[u] = __ldcs(a2 + c + 32 * u);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double2 xv = __ldg(x2 + c + 32 * u);
            s0 = fma(av[u].x, xv.x, s0);
            s1 = fma(av[u].y, xv.y, s1);
        }
    }
    for (; c < len2; c += 32) {
        const double2 av = __ldcs(a2 + c);
        const double2 xv = __ldg(x2 + c);
        s0 = fma(av.x, xv.x, s0);
        s1 = fma(av.y, xv.y, s1);
    }
    return warp_sum(s0 + s1);
}

// w = X^-1 hd
template <int U>
__global__ void __launch_bounds__(256) upwards_w_kernel(const MergeEntry* __restrict__ ent, int n, int rows_per_cta)
{
    const MergeEntry& e = ent[blockIdx.x];
    const int N = 4 * n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    for (int r = r0 + warp; r < r1; r += 8) {
        double s = row_dot<U>(e.Xinv + (size_t)r * N, e.hd, N, lane);
        if (lane == 0) e.w[r] = s;
    }
}

// h = pi( H w + h_ext )   with the compact H: row block p of child c multiplies w[k0], w[k1]
template <int U>
__global__ void __launch_bounds__(256) upwards_h_kernel(const MergeEntry* __restrict__ ent, int n, int rows_per_cta)
{
    const MergeEntry& e = ent[blockIdx.x];
    const int R = 8 * n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    for (int row = r0 + warp; row < r1; row += 8) {
        const int p = row / n, r = row % n;
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        const double* a = e.Hc + (size_t)row * (2 * n);
        double s = row_dot<U>(a, e.w + c_kk[ch][0] * n, n, lane) + row_dot<U>(a + n, e.w + c_kk[ch][1] * n, n, lane);
        if (lane == 0) e.h[row] = s + e.hc[ch][side * n + r];
    }
}

// ---- solve -------------------------------------------------------------------------------------
// u_int = S g (+ w); children's Dirichlet data assembled in WESN order (HPSAlgorithm.hpp:1206-1227)
template <int U>
__global__ void __launch_bounds__(256) solve_split_kernel(const MergeEntry* __restrict__ ent, int n, int rows_per_cta, int add_w)
{
    const MergeEntry& e = ent[blockIdx.x];
    const int N = 4 * n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    for (int row = r0 + warp; row < r1; row += 8) {
        double s = row_dot<U>(e.S + (size_t)row * (8 * n), e.g, 8 * n, lane);
        if (lane == 0) {
            if (add_w) s += e.w[row];
            const int k = row / n, r = row % n;
            const int c1 = c_kids[k][0], c2 = c_kids[k][1];
            e.gc[c1][c_iface[c1][k] * n + r] = s;
            e.gc[c2][c_iface[c2][k] * n + r] = s;
        }
    }
    // exterior segments are copied through: this CTA's share of the 8n entries
    const int nct = gridDim.y, per = (8 * n + nct - 1) / nct;
    const int x0 = blockIdx.y * per, x1 = min(8 * n, x0 + per);
    for (int idx = x0 + threadIdx.x; idx < x1; idx += blockDim.x) {
        const int p = idx / n, r = idx % n;
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        e.gc[ch][side * n + r] = e.g[idx];
    }
}

// ---- short-row variants (child side n <= 256: rows of 2n..8n doubles) ---------------------------
// One CTA per (parent, row chunk).  The right-hand vector is staged in shared memory once; a row is
// owned by a group of LW lanes (LW = min(32, L/2), 16-byte loads), so short rows do not idle most of
// a warp, and every group keeps R rows in flight before reducing (memory-level parallelism).
template <int LW, int R, class Epilogue>
__device__ __forceinline__ void gemv_rows_short(const double* __restrict__ A, int L, const double* xs, int r0, int r1, Epilogue epi)
{
    const int tid = threadIdx.x;
    const int grp = tid / LW, gl = tid % LW, ngrp = blockDim.x / LW;
    const int L2 = L >> 1;
    const double2* x2 = reinterpret_cast<const double2*>(xs);
    for (int rb = r0 + grp * R; rb < r1; rb += ngrp * R) {
        double acc[R];
#pragma unroll
        for (int k = 0; k < R; k++) acc[k] = 0.0;
        for (int c = gl; c < L2; c += LW) {
            double2 av[R];
#pragma unroll
            for (int k = 0; k < R; k++)
                av[k] = (rb + k < r1) ? __ldcs(reinterpret_cast<const double2*>(A + (size_t)(rb + k) * L) + c) : make_double2(0.0, 0.0);
            const double2 xv = x2[c];
#pragma unroll
            for (int k = 0; k < R; k++) acc[k] = fma(av[k].x, xv.x, fma(av[k].y, xv.y, acc[k]));
        }
#pragma unroll
        for (int o = LW / 2; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < R; k++) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (gl == 0) {
#pragma unroll
            for (int k = 0; k < R; k++) if (rb + k < r1) epi(rb + k, acc[k]);
        }
    }
}

// w = X^-1 hd with hd formed on the fly from the children's h (fuses hdiff_kernel)
template <int LW, int R>
__global__ void __launch_bounds__(256) upwards_w_short_kernel(const MergeEntry* __restrict__ ent, int n, int rows_per_cta)
{
    extern __shared__ __align__(16) double xs[];
    const MergeEntry& e = ent[blockIdx.x];
    const int N = 4 * n;
    for (int idx = threadIdx.x; idx < N; idx += blockDim.x) {
        const int k = idx / n, r = idx % n;
        const int c1 = c_kids[k][0], c2 = c_kids[k][1];
        xs[idx] = e.hc[c2][c_iface[c2][k] * n + r] - e.hc[c1][c_iface[c1][k] * n + r];
    }
    __syncthreads();
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    double* w = e.w;
    gemv_rows_short<LW, R>(e.Xinv, N, xs, r0, r1, [&](int row, double v) { w[row] = v; });
}

// h = pi(H w + h_ext): row `row` of the compact H (length 2n) times [w[k0], w[k1]] of its child
template <int LW, int R>
__global__ void __launch_bounds__(256) upwards_h_short_kernel(const MergeEntry* __restrict__ ent, int n, int rows_per_cta)
{
    extern __shared__ __align__(16) double xs[];   // 4n: w
    const MergeEntry& e = ent[blockIdx.x];
    for (int idx = threadIdx.x; idx < 4 * n; idx += blockDim.x) xs[idx] = e.w[idx];
    __syncthreads();
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(8 * n, r0 + rows_per_cta);
    // rows of one WESN block p share the child and therefore the two w segments; chunks never straddle a block
    // when rows_per_cta divides n or is a multiple of it, which the launcher guarantees.
    for (int p0 = r0; p0 < r1; p0 += n) {
        const int p = p0 / n, pr0 = p0, pr1 = min(r1, (p + 1) * n);
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        const int k0 = c_kk[ch][0], k1 = c_kk[ch][1];
        const double* hc = e.hc[ch] + side * n;
        double* h = e.h;
        // the two segments are contiguous in xs only if k1 == k0 + 1; otherwise two passes over half rows
        if (k1 == k0 + 1) {
            gemv_rows_short<LW, R>(e.Hc, 2 * n, xs + k0 * n, pr0, pr1, [&](int row, double v) { h[row] = v + hc[row - p * n]; });
        } else {
            __shared__ double xcat[512];
            __syncthreads();
            for (int idx = threadIdx.x; idx < 2 * n; idx += blockDim.x) xcat[idx] = idx < n ? xs[k0 * n + idx] : xs[k1 * n + idx - n];
            __syncthreads();
            gemv_rows_short<LW, R>(e.Hc, 2 * n, xcat, pr0, pr1, [&](int row, double v) { h[row] = v + hc[row - p * n]; });
        }
    }
}

// u_int = S g (+ w), scattered to the children; exterior segments copied through
template <int LW, int R>
__global__ void __launch_bounds__(256) solve_split_short_kernel(const MergeEntry* __restrict__ ent, int n, int rows_per_cta, int add_w)
{
    extern __shared__ __align__(16) double xs[];   // 8n: g
    const MergeEntry& e = ent[blockIdx.x];
    for (int idx = threadIdx.x; idx < 8 * n; idx += blockDim.x) xs[idx] = e.g[idx];
    __syncthreads();
    const int N = 4 * n;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    const double* w = e.w;
    gemv_rows_short<LW, R>(e.S, 8 * n, xs, r0, r1, [&](int row, double v) {
        if (add_w) v += w[row];
        const int k = row / n, r = row % n;
        const int c1 = c_kids[k][0], c2 = c_kids[k][1];
        e.gc[c1][c_iface[c1][k] * n + r] = v;
        e.gc[c2][c_iface[c2][k] * n + r] = v;
    });
    const int nct = gridDim.y, per = (8 * n + nct - 1) / nct;
    const int x0 = blockIdx.y * per, x1 = min(8 * n, x0 + per);
    for (int idx = x0 + threadIdx.x; idx < x1; idx += blockDim.x) {
        const int p = idx / n, r = idx % n;
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        e.gc[ch][side * n + r] = xs[idx];
    }
}


// ---- bulk-copy streaming variants (child side n = 16 .. 512, powers of two) ----------------------
// The matrix rows a CTA owns are one contiguous byte range, so the CTA streams it through a ring of shared-memory
// stages filled by the bulk-copy engine (cp.async.bulk global -> shared, completion on an mbarrier): SG_NS x 16 KB in
// flight per CTA independent of register pressure, 2-3 CTAs per SM.  The right-hand vector is staged in shared memory
// once.  Warp w reduces the 256 doubles [256 w, 256 w + 256) of every stage: rows of <= 256 entries are finished inside
// the warp, longer rows (<= 2048 entries = one stage) across the warps through shared memory.
constexpr int SG_STAGE = 2048;   // doubles per stage
constexpr int SG_NS = 4;
constexpr int SG_XMAX = 2048;    // longest staged vector

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_DONE;\n"
        "bra SG_WAIT;\n"
        "SG_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// dynamic shared memory layout: [SG_NS stages][x: xlen doubles][red: 8][barriers: SG_NS]
struct SgView {
    double* stage; double* xs; double* red; unsigned long long* full;
    __device__ SgView(double* base, int xlen) : stage(base), xs(base + SG_NS * SG_STAGE), red(xs + xlen), full(reinterpret_cast<unsigned long long*>(red + 8)) {}
};
static inline size_t sg_smem_bytes(int xlen) { return (size_t)(SG_NS * SG_STAGE + xlen + 8 + SG_NS) * sizeof(double); }

// A: first row of this CTA's chunk; rows of L = 2^logL doubles (32 <= L <= 2048); ntiles stages of SG_STAGE doubles.
// xfill(i): entry i of the staged vector; xidx(row, col): its index for matrix entry (row, col), row local to the chunk;
// epi(row, value): row local to the chunk.  All 256 threads must call.
template <class XFill, class XIdx, class Epi>
__device__ __forceinline__ void stream_gemv(const double* __restrict__ A, int logL, int ntiles, int xlen, SgView sm, XFill xfill, XIdx xidx, Epi epi)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = 1 << logL;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SG_NS; s++) mbar_init(sm.full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SG_NS; s++)
            if (s < ntiles) {
                mbar_expect_tx(sm.full + s, SG_STAGE * 8);
                bulk_g2s(sm.stage + s * SG_STAGE, A + (size_t)s * SG_STAGE, SG_STAGE * 8, sm.full + s);
            }
    }
    for (int i = tid; i < xlen; i += blockDim.x) sm.xs[i] = xfill(i);
    __syncthreads();
    double acc = 0.0;
    for (int t = 0; t < ntiles; t++) {
        const int s = t % SG_NS;
        mbar_wait(sm.full + s, (unsigned)((t / SG_NS) & 1));
        const double* st = sm.stage + s * SG_STAGE + warp * 256 + lane * 2;
        const int f0 = t * SG_STAGE + warp * 256 + lane * 2;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double2 a = *reinterpret_cast<const double2*>(st + j * 64);
            const int e = f0 + j * 64, row = e >> logL, col = e & (L - 1);
            const double2 x = *reinterpret_cast<const double2*>(sm.xs + xidx(row, col));
            acc = fma(a.x, x.x, fma(a.y, x.y, acc));
            if (L == 32) {            // two rows per 64 entries: reduce inside half warps
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if ((lane & 15) == 0) epi(row, acc);
                acc = 0.0;
            } else if (L <= 256 && (((j + 1) * 64) & (L - 1)) == 0) {   // a row ends inside this warp's slice
                acc = warp_sum(acc);
                if (lane == 0) epi(row, acc);
                acc = 0.0;
            }
        }
        if (L > 256) {                // a row spans L / 256 warps of this stage
            acc = warp_sum(acc);
            if (lane == 0) sm.red[warp] = acc;
            acc = 0.0;
            __syncthreads();
            const int wpr = L >> 8, rows = SG_STAGE >> logL;
            if (tid < rows) {
                double v = 0.0;
                for (int k = 0; k < wpr; k++) v += sm.red[tid * wpr + k];
                epi(t * rows + tid, v);
            }
        }
        __syncthreads();              // every warp is done with stage s (and with red)
        if (tid == 0 && t + SG_NS < ntiles) {
            mbar_expect_tx(sm.full + s, SG_STAGE * 8);
            bulk_g2s(sm.stage + s * SG_STAGE, A + (size_t)(t + SG_NS) * SG_STAGE, SG_STAGE * 8, sm.full + s);
        }
    }
}

__global__ void __launch_bounds__(256) sg_upwards_w_kernel(const MergeEntry* __restrict__ ent, int n, int logn, int rows_per_cta)
{
    extern __shared__ __align__(128) double sg_smem[];
    const MergeEntry& e = ent[blockIdx.x];
    const int N = 4 * n, r0 = blockIdx.y * rows_per_cta;
    SgView sm(sg_smem, N);
    double* w = e.w;
    stream_gemv(e.Xinv + (size_t)r0 * N, logn + 2, rows_per_cta * N / SG_STAGE, N, sm,
        [&](int idx) {
            const int k = idx >> logn, r = idx & (n - 1);
            const int c1 = c_kids[k][0], c2 = c_kids[k][1];
            return e.hc[c2][c_iface[c2][k] * n + r] - e.hc[c1][c_iface[c1][k] * n + r];
        },
        [](int, int col) { return col; },
        [&](int row, double v) { w[r0 + row] = v; });
}

__global__ void __launch_bounds__(256) sg_upwards_h_kernel(const MergeEntry* __restrict__ ent, int n, int logn, int rows_per_cta)
{
    extern __shared__ __align__(128) double sg_smem[];
    const MergeEntry& e = ent[blockIdx.x];
    const int r0 = blockIdx.y * rows_per_cta;
    SgView sm(sg_smem, 4 * n);
    double* h = e.h;
    const double* wv = e.w;
    stream_gemv(e.Hc + (size_t)r0 * (2 * n), logn + 1, rows_per_cta * 2 * n / SG_STAGE, 4 * n, sm,
        [&](int idx) { return wv[idx]; },
        [&](int row, int col) {   // row block p of child ch multiplies [w[k0], w[k1]]
            const int ch = c_pi[(r0 + row) >> logn] >> 1;
            return c_kk[ch][col >> logn] * n + (col & (n - 1));
        },
        [&](int row, double v) {
            const int gr = r0 + row, p = gr >> logn, r = gr & (n - 1);
            const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
            h[gr] = v + e.hc[ch][side * n + r];
        });
}

__global__ void __launch_bounds__(256) sg_solve_split_kernel(const MergeEntry* __restrict__ ent, int n, int logn, int rows_per_cta, int add_w)
{
    extern __shared__ __align__(128) double sg_smem[];
    const MergeEntry& e = ent[blockIdx.x];
    const int r0 = blockIdx.y * rows_per_cta;
    SgView sm(sg_smem, 8 * n);
    const double* g = e.g;
    const double* w = e.w;
    stream_gemv(e.S + (size_t)r0 * (8 * n), logn + 3, rows_per_cta * 8 * n / SG_STAGE, 8 * n, sm,
        [&](int idx) { return g[idx]; },
        [](int, int col) { return col; },
        [&](int row, double v) {
            const int gr = r0 + row;
            if (add_w) v += w[gr];
            const int k = gr >> logn, r = gr & (n - 1);
            const int c1 = c_kids[k][0], c2 = c_kids[k][1];
            e.gc[c1][c_iface[c1][k] * n + r] = v;
            e.gc[c2][c_iface[c2][k] * n + r] = v;
        });
    // exterior segments are copied through: this CTA's share of the 8n entries (xs still holds g)
    const int nct = gridDim.y, per = (8 * n + nct - 1) / nct;
    const int x0 = blockIdx.y * per, x1 = min(8 * n, x0 + per);
    for (int idx = x0 + threadIdx.x; idx < x1; idx += blockDim.x) {
        const int p = idx >> logn, r = idx & (n - 1);
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        e.gc[ch][side * n + r] = sm.xs[idx];
    }
}


// ---- launch wrappers ---------------------------------------------------------------------------
static inline int ew_blocks(long long elems) { long long b = (elems + 255) / 256; return (int)(b < 1 ? 1 : (b > 1024 ? 1024 : b)); }
static inline void check_count(int count) { if (count > 65535) throw Error{EF_ERR_BAD_SHAPE, "batch too large for grid.y (chunk it)"}; }

void launch_assemble_X(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    for (int off = 0; off < count; off += 65535) {
        int c = count - off < 65535 ? count - off : 65535;
        assemble_X_kernel<<<dim3(ew_blocks(16LL * n * n), c), 256, 0, s>>>(e + off, n);
    }
    EF_CUDA(cudaGetLastError());
}
void launch_assemble_Hc(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    for (int off = 0; off < count; off += 65535) {
        int c = count - off < 65535 ? count - off : 65535;
        assemble_Hc_kernel<<<dim3(ew_blocks(16LL * n * n), c), 256, 0, s>>>(e + off, n);
    }
    EF_CUDA(cudaGetLastError());
}
void launch_expand_H(const double* Hc, int n, double* H_dense, cudaStream_t s)
{
    expand_H_kernel<<<ew_blocks(32LL * n * n), 256, 0, s>>>(Hc, n, H_dense);
    EF_CUDA(cudaGetLastError());
}
void launch_coarsen_T(const CoarsenOp* ops, int nops, int max_nfine, cudaStream_t s)
{
    if (!nops) return;
    check_count(nops);
    coarsen_T_kernel<<<dim3(ew_blocks(4LL * max_nfine * max_nfine), nops), 256, 0, s>>>(ops);
    EF_CUDA(cudaGetLastError());
}
void launch_coarsen_h(const CoarsenOp* ops, int nops, int max_nfine, cudaStream_t s)
{
    if (!nops) return;
    check_count(nops);
    coarsen_h_kernel<<<dim3(ew_blocks(2LL * max_nfine), nops), 256, 0, s>>>(ops);
    EF_CUDA(cudaGetLastError());
}
void launch_uncoarsen_g(const CoarsenOp* ops, int nops, int max_nfine, cudaStream_t s)
{
    if (!nops) return;
    check_count(nops);
    uncoarsen_g_kernel<<<dim3(ew_blocks(4LL * max_nfine), nops), 256, 0, s>>>(ops);
    EF_CUDA(cudaGetLastError());
}

// rows per CTA: one row per warp at least; aim for >= 4 waves of CTAs over the whole batch
static inline int pick_rows(int rows, int count)
{
    int rpc = rows;
    const long long want = 148LL * (g_tuning[2] > 0 ? g_tuning[2] : 16);
    while (rpc > 8 && (long long)count * ((rows + rpc - 1) / rpc) < want) rpc = (rpc + 1) / 2;
    if (rpc < 8) rpc = 8;
    return rpc;
}

// chunk of rows per CTA for the short-row kernels: whole parents when there are many, else n-aligned pieces
static inline int short_rows(int rows, int n, int count)
{
    int rpc = rows;
    while (rpc > n && (long long)count * (rows / rpc) < 148LL * 8) rpc >>= 1;
    while (rpc > 32 && (long long)count * (rows / rpc) < 148LL * 4) rpc >>= 1;   // below n: still a divisor of n (n = 8 * 2^k or 24 * 2^k ...)
    return rpc;
}

template <int R>
static void launch_upwards_short(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    {
        const int N = 4 * n, rpc = short_rows(N, n, count);
        dim3 grid(count, N / rpc);
        const size_t sm = (size_t)N * sizeof(double);
        if (N / 2 >= 32) upwards_w_short_kernel<32, R><<<grid, 256, sm, s>>>(e, n, rpc);
        else upwards_w_short_kernel<16, R><<<grid, 256, sm, s>>>(e, n, rpc);
    }
    {
        const int rows = 8 * n;
        int rpc = short_rows(rows, n, count);
        if (rpc < n && n % rpc) rpc = n;       // chunks must not straddle WESN blocks
        dim3 grid(count, rows / rpc);
        const size_t sm = (size_t)4 * n * sizeof(double);
        const int L2 = n;                       // row length 2n doubles = n double2
        if (L2 >= 32) upwards_h_short_kernel<32, R><<<grid, 256, sm, s>>>(e, n, rpc);
        else if (L2 >= 16) upwards_h_short_kernel<16, R><<<grid, 256, sm, s>>>(e, n, rpc);
        else upwards_h_short_kernel<8, R><<<grid, 256, sm, s>>>(e, n, rpc);
    }
}

// rows per CTA of the streaming kernels: whole parents when there are many of them, else halved while the grid is short
// of ~6 CTAs per SM; a chunk is a whole number (>= 1) of stages.
static inline int sg_rows(int rows, int L, int count)
{
    int rpc = rows;
    while ((long long)count * (rows / rpc) < 148LL * 6 && (long long)(rpc / 2) * L >= 2LL * SG_STAGE) rpc >>= 1;
    return rpc;
}
static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
template <class K>
static void sg_attr(K kern)
{
    EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem_bytes(SG_XMAX)));
}

void launch_upwards(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    if (!count) return;
    if (g_tuning[0] && n >= 16 && n <= 512 && (n & (n - 1)) == 0 && count <= 2147483647) {
        static bool attr = false;
        if (!attr) { sg_attr(sg_upwards_w_kernel); sg_attr(sg_upwards_h_kernel); attr = true; }
        const int logn = ilog2(n);
        int rpc = sg_rows(4 * n, 4 * n, count);
        sg_upwards_w_kernel<<<dim3(count, 4 * n / rpc), 256, sg_smem_bytes(4 * n), s>>>(e, n, logn, rpc);
        rpc = sg_rows(8 * n, 2 * n, count);
        sg_upwards_h_kernel<<<dim3(count, 8 * n / rpc), 256, sg_smem_bytes(4 * n), s>>>(e, n, logn, rpc);
        EF_CUDA(cudaGetLastError());
        return;
    }
    if (n <= 256 && (n & (n - 1)) == 0 && count <= 65535 * 0 + 2147483647) {
        launch_upwards_short<4>(e, n, count, s);
        EF_CUDA(cudaGetLastError());
        return;
    }
    for (int off = 0; off < count; off += 65535) {
        int c = count - off < 65535 ? count - off : 65535;
        hdiff_kernel<<<dim3(ew_blocks(4LL * n), c), 256, 0, s>>>(e + off, n);
    }
    int rpc = pick_rows(4 * n, count);
    if (g_tuning[1] == 1) upwards_w_kernel<8><<<dim3(count, (4 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
    else upwards_w_kernel<4><<<dim3(count, (4 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
    rpc = pick_rows(8 * n, count);
    if (g_tuning[1] == 1) upwards_h_kernel<8><<<dim3(count, (8 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
    else upwards_h_kernel<4><<<dim3(count, (8 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
    EF_CUDA(cudaGetLastError());
}

void launch_solve_split(const MergeEntry* e, int n, int count, bool add_w, cudaStream_t s)
{
    if (!count) return;
    if (g_tuning[0] && n >= 16 && n <= 256 && (n & (n - 1)) == 0) {
        static bool attr = false;
        if (!attr) { sg_attr(sg_solve_split_kernel); attr = true; }
        const int rpc = sg_rows(4 * n, 8 * n, count);
        sg_solve_split_kernel<<<dim3(count, 4 * n / rpc), 256, sg_smem_bytes(8 * n), s>>>(e, n, ilog2(n), rpc, add_w ? 1 : 0);
        EF_CUDA(cudaGetLastError());
        return;
    }
    if (n <= 256 && (n & (n - 1)) == 0) {
        const int N = 4 * n, rpc = short_rows(N, n, count);
        dim3 grid(count, N / rpc);
        solve_split_short_kernel<32, 4><<<grid, 256, (size_t)8 * n * sizeof(double), s>>>(e, n, rpc, add_w ? 1 : 0);
        EF_CUDA(cudaGetLastError());
        return;
    }
    int rpc = pick_rows(4 * n, count);
    if (g_tuning[1] == 1) solve_split_kernel<8><<<dim3(count, (4 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc, add_w ? 1 : 0);
    else solve_split_kernel<4><<<dim3(count, (4 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc, add_w ? 1 : 0);
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
