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
// [0] 2 = row-batch kernels for rows of <= 256 doubles (default), 0 = one row per warp everywhere
// [1] compact-H long-row kernel: 0 = 8 loads in flight per lane (default), 1 = 4
// [2] CTAs per SM the long-row launcher aims for (default 16)
// [3] leaf solve of constant-coefficient leaves: 0 = DMMA kernel (default), 1 = one thread per cell
// [4] transposed second destination of a GEMM block: 0 = through shared memory as whole rows where peer arenas receive it, direct 8-byte
//     stores otherwise (default); 1 = always through shared memory; 2 = always direct
// [5] symmetric merge plan: 1 = the diagonal blocks of T multiply only their upper sub-block triangle (default since r2a: 202.8 -> 197.8 ms
//     per step at L=8 M=16; 0 = whole blocks; read when a plan is made)
// [7] base case of the block inversion, 128 x 128: 0 = blocked Gauss-Jordan on the tensor pipe, 1 = per-pivot register kernel (round 1)
// [6] variable-coefficient leaves with M = 8, 16: 0 = warp-level / tensor-core kernels (default), 1 = the CTA-per-leaf kernels of round 1
// [8] operand staging of the 128-row GEMM tiles: 1 = TMA (cp.async.bulk.tensor.2d, swizzled shared memory, mbarrier ring; default),
//     0 = cp.async (LDGSTS) into padded shared memory; read at every launch
// [9] base case of the block inversion in batches of at most four merges: 1 = 256 x 256 blocks by a cluster of eight CTAs, 0 = recursion
//     down to 128 x 128 everywhere (default: measured r2t, 190 us per 256-block against 136 us - the owner's serial section, eight
//     dependent reciprocals per pivot block plus the push, leaves the other seven CTAs in the cluster barrier 61 % of the time);
//     read when a plan is made
// [10] pivot reciprocals of the 128 x 128 base case: 1 = hardware seed + two Newton steps, 0 = IEEE division (default; measured r2v: no change)
// [11] peer-mapped partitions over 2 / 4 / 8 ranks: 1 = S and T split by block COLUMNS, so that T needs no exchange of S, 0 = by rows
//      (default: measured r2w / r2x, 103.0 vs 102.2 ms at 2 GPUs and 43.6 vs 41.6 ms at 8 - the barrier saved does not pay for the
//      slower column-wise products, whose stores of S and T then drain together); read when a plan is made
static int g_tuning[16] = {2, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
void set_tuning(int key, int value) { if (key >= 0 && key < 16) g_tuning[key] = value; }
int get_tuning(int key) { return (key >= 0 && key < 16) ? g_tuning[key] : 0; }

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

// X[k][k'] = - sum_c sgn_c(k) T^c[iface_c(k), iface_c(k')]   (16 n x n blocks, 4 of them zero); two entries per thread
__global__ void __launch_bounds__(256) assemble_X_kernel(const MergeEntry* __restrict__ ent, int n)
{
    const MergeEntry& e = ent[blockIdx.y];
    const int N = 4 * n, N2 = N / 2;
    const double* Tc[4] = {e.Tc[0], e.Tc[1], e.Tc[2], e.Tc[3]};
    double2* X2 = reinterpret_cast<double2*>(e.Xinv);
    double2* Xc2 = reinterpret_cast<double2*>(e.Xcopy);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * N2; idx += gridDim.x * blockDim.x) {
        const int row = idx / N2, col = 2 * (idx - row * N2);        // col even: both entries lie in one block (n is even)
        const int k = row / n, r = row - k * n, k2 = col / n, c = col - k2 * n;
        double2 v = make_double2(0.0, 0.0);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            const int sa = c_iface[ch][k], sb = c_iface[ch][k2];
            if (sa >= 0 && sb >= 0) {
                const double2 t = *reinterpret_cast<const double2*>(Tc[ch] + (size_t)(sa * n + r) * N + sb * n + c);
                const double s = c_sgn[ch][k];
                v.x -= s * t.x; v.y -= s * t.y;
            }
        }
        X2[idx] = v;
        if (Xc2) Xc2[idx] = v;
    }
}

// Hc row block p (WESN position) of child c = [ T^c[side, iface_c(k0)] | T^c[side, iface_c(k1)] ]; two entries per thread
__global__ void __launch_bounds__(256) assemble_Hc_kernel(const MergeEntry* __restrict__ ent, int n)
{
    const MergeEntry& e = ent[blockIdx.y];
    const int N = 4 * n;
    const double* Tc[4] = {e.Tc[0], e.Tc[1], e.Tc[2], e.Tc[3]};
    double2* H2 = reinterpret_cast<double2*>(e.Hc);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 8 * n * n; idx += gridDim.x * blockDim.x) {
        const int row = idx / n, col = 2 * (idx - row * n);           // row of 2n entries = n pairs
        const int p = row / n, r = row - p * n, t = col / n, c = col - t * n;
        const int q = c_pi[p], ch = q >> 1, side = c_tau_side[ch][q & 1];
        const int sb = c_iface[ch][c_kk[ch][t]];
        const double* T = ch == 0 ? Tc[0] : ch == 1 ? Tc[1] : ch == 2 ? Tc[2] : Tc[3];
        H2[idx] = *reinterpret_cast<const double2*>(T + (size_t)(side * n + r) * N + sb * n + c);
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

// ---- row-batch kernels (rows of L = 32 .. 256 doubles) -------------------------------------------
// Warp-autonomous, no shared memory, no CTA barrier: a warp owns a contiguous range of 4 KB batches of the level's
// operator slab.  One batch = 8 independent, fully coalesced 512-byte loads per warp (16 bytes per lane, streaming), i.e.
// 512 / L complete rows.  The lane's slice of the right-hand vector lives in registers (reloaded when the
// parent - or, for H, the WESN block - changes).  The per-row partial sums of a batch are reduced TOGETHER by a
// butterfly that halves the number of live values at every exchange (V values over LW lanes cost V - 1 + log2(LW / V)
// shuffles instead of V log2(LW)), which is what makes short rows cheap: 8 rows of 64 doubles take 9 shuffles, not 40.
template <int V, int LW>
__device__ __forceinline__ double butterfly_reduce(double (&v)[V], int lane)
{
    int o = LW / 2;
#pragma unroll
    for (int cnt = V; cnt > 1; cnt >>= 1, o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; i++) {
            const double send = upper ? v[i] : v[i + cnt / 2];
            const double keep = upper ? v[i + cnt / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    double r = v[0];
#pragma unroll
    for (; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;   // complete sum of row (lane % LW) / (LW / V) of the group, replicated over LW / V lanes
}

// One batch: 512 doubles at A; xf = this lane's slice of the right-hand vector (double2 per 64 columns).  Returns the row
// sum this lane ends up holding and sets `row` (within the batch) and `writer` (one lane per row).
template <int L>
__device__ __forceinline__ double rb_batch(const double* __restrict__ A, const double2 (&xf)[(L >= 64 ? L / 64 : 1)], int lane, int& row, bool& writer)
{
    const double2* a2 = reinterpret_cast<const double2*>(A) + lane;
    double2 a[8];
#pragma unroll
    for (int r = 0; r < 8; r++) a[r] = __ldcs(a2 + 32 * r);
    if constexpr (L == 32) {           // a load holds two rows (one per half warp): 8 values over 16 lanes
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = fma(a[r].x, xf[0].x, a[r].y * xf[0].y);
        const double s = butterfly_reduce<8, 16>(v, lane);
        row = 2 * ((lane & 15) >> 1) + (lane >> 4);
        writer = (lane & 1) == 0;
        return s;
    } else {
        constexpr int LPR = L / 64, V = 8 / LPR;   // loads per row, rows per batch
        double v[V];
#pragma unroll
        for (int i = 0; i < V; i++) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < LPR; q++) s = fma(a[i * LPR + q].x, xf[q].x, fma(a[i * LPR + q].y, xf[q].y, s));
            v[i] = s;
        }
        const double s = butterfly_reduce<V, 32>(v, lane);
        row = lane / (32 / V);
        writer = (lane & (32 / V - 1)) == 0;
        return s;
    }
}

// the warp's contiguous range of batches
__device__ __forceinline__ void rb_range(long long total, long long& b0, long long& b1)
{
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5), w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long per = (total + nw - 1) / nw;
    b0 = w * per; b1 = min(total, b0 + per);
}

// w = X^-1 hd, hd formed from the children's h; L = 4n
template <int L>
__global__ void __launch_bounds__(256) rb_upwards_w_kernel(const MergeEntry* __restrict__ ent, int count, int logn)
{
    constexpr int XF = L >= 64 ? L / 64 : 1, ROWS = 512 / L;
    const int n = 1 << logn, lane = threadIdx.x & 31;
    const int lbpp = 2 * logn + 4 - 9;                     // log2 of the batches per parent: 16 n^2 / 512
    long long b0, b1;
    rb_range((long long)count << lbpp, b0, b1);
    int cur = -1;
    double2 xf[XF];
    const double* A = nullptr; double* w = nullptr;
    for (long long b = b0; b < b1; b++) {
        const int p = (int)(b >> lbpp), lb = (int)(b & ((1 << lbpp) - 1));
        if (p != cur) {
            cur = p;
            const MergeEntry& e = ent[p];
            A = e.Xinv; w = e.w;
#pragma unroll
            for (int q = 0; q < XF; q++) {
                const int idx = q * 64 + 2 * lane, k = idx >> logn, r = idx & (n - 1);
                const int c1 = c_kids[k][0], c2 = c_kids[k][1];
                const double2 hi = *reinterpret_cast<const double2*>(e.hc[c2] + c_iface[c2][k] * n + r);
                const double2 lo = *reinterpret_cast<const double2*>(e.hc[c1] + c_iface[c1][k] * n + r);
                xf[q] = make_double2(hi.x - lo.x, hi.y - lo.y);
            }
        }
        int row; bool writer;
        const double s = rb_batch<L>(A + (size_t)lb * 512, xf, lane, row, writer);
        if (writer) w[lb * ROWS + row] = s;
    }
}

// h = pi(H w + h_ext); L = 2n; rows of WESN block p belong to child ch and multiply [w[k0], w[k1]]
template <int L>
__global__ void __launch_bounds__(256) rb_upwards_h_kernel(const MergeEntry* __restrict__ ent, int count, int logn)
{
    constexpr int XF = L >= 64 ? L / 64 : 1, ROWS = 512 / L;
    const int n = 1 << logn, lane = threadIdx.x & 31;
    const int lbpp = 2 * logn + 4 - 9;                     // 16 n^2 / 512
    long long b0, b1;
    rb_range((long long)count << lbpp, b0, b1);
    int cur = -1, curblk = -1;
    double2 xf[XF];
    const double* A = nullptr; const double* wv = nullptr; const double* hext = nullptr; double* h = nullptr;
    const double* hc[4] = {nullptr, nullptr, nullptr, nullptr};
    for (long long b = b0; b < b1; b++) {
        const int p = (int)(b >> lbpp), lb = (int)(b & ((1 << lbpp) - 1));
        if (p != cur) {
            cur = p; curblk = -1;
            const MergeEntry& e = ent[p];
            A = e.Hc; wv = e.w; h = e.h;
#pragma unroll
            for (int c = 0; c < 4; c++) hc[c] = e.hc[c];
        }
        const int row0 = lb * ROWS, blk = row0 >> logn;    // a batch never straddles a block: ROWS <= 16 <= n
        if (blk != curblk) {
            curblk = blk;
            const int q8 = c_pi[blk], ch = q8 >> 1, side = c_tau_side[ch][q8 & 1];
            hext = (ch == 0 ? hc[0] : ch == 1 ? hc[1] : ch == 2 ? hc[2] : hc[3]) + side * n - (blk << logn);
#pragma unroll
            for (int q = 0; q < XF; q++) {
                const int col = q * 64 + 2 * (L == 32 ? (lane & 15) : lane);
                xf[q] = *reinterpret_cast<const double2*>(wv + c_kk[ch][col >> logn] * n + (col & (n - 1)));
            }
        }
        int row; bool writer;
        const double s = rb_batch<L>(A + (size_t)lb * 512, xf, lane, row, writer);
        if (writer) h[row0 + row] = s + hext[row0 + row];
    }
}

// u_int = S g (+ w) scattered to the adjacent children, exterior segments copied through; L = 8n
template <int L>
__global__ void __launch_bounds__(256) rb_solve_split_kernel(const MergeEntry* __restrict__ ent, int count, int logn, int add_w)
{
    constexpr int XF = L / 64, ROWS = 512 / L;
    const int n = 1 << logn, lane = threadIdx.x & 31;
    const int lbpp = 2 * logn + 5 - 9;                     // 32 n^2 / 512
    long long b0, b1;
    rb_range((long long)count << lbpp, b0, b1);
    int cur = -1;
    double2 xf[XF];
    const double* A = nullptr; const double* w = nullptr;
    double* gc[4] = {nullptr, nullptr, nullptr, nullptr};
    for (long long b = b0; b < b1; b++) {
        const int p = (int)(b >> lbpp), lb = (int)(b & ((1 << lbpp) - 1));
        if (p != cur) {
            cur = p;
            const MergeEntry& e = ent[p];
            A = e.S; w = e.w;
#pragma unroll
            for (int c = 0; c < 4; c++) gc[c] = e.gc[c];
#pragma unroll
            for (int q = 0; q < XF; q++) xf[q] = *reinterpret_cast<const double2*>(e.g + q * 64 + 2 * lane);
            if (lb == 0) {     // the warp that starts a parent also copies its exterior segments through
                for (int idx = lane; idx < 8 * n; idx += 32) {
                    const int pb = idx >> logn, r = idx & (n - 1);
                    const int q8 = c_pi[pb], ch = q8 >> 1, side = c_tau_side[ch][q8 & 1];
                    (ch == 0 ? gc[0] : ch == 1 ? gc[1] : ch == 2 ? gc[2] : gc[3])[side * n + r] = e.g[idx];
                }
            }
        }
        int row; bool writer;
        double s = rb_batch<L>(A + (size_t)lb * 512, xf, lane, row, writer);
        if (writer) {
            const int gr = lb * ROWS + row;
            if (add_w) s += w[gr];
            const int k = gr >> logn, r = gr & (n - 1);
            const int c1 = c_kids[k][0], c2 = c_kids[k][1];
            (c1 == 0 ? gc[0] : c1 == 1 ? gc[1] : gc[2])[c_iface[c1][k] * n + r] = s;
            (c2 == 1 ? gc[1] : c2 == 2 ? gc[2] : gc[3])[c_iface[c2][k] * n + r] = s;
        }
    }
}

// grid: as many CTAs as the device holds at once (never more warps than batches)
template <class K>
static int rb_grid(K kern, long long batches)
{
    int dev = 0, sms = 0, per_sm = 0;
    EF_CUDA(cudaGetDevice(&dev));
    EF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    long long g = (long long)sms * (per_sm > 0 ? per_sm : 1);
    const long long need = (batches + 7) / 8;
    return (int)(need < g ? need : g);
}
#define RB_LAUNCH(KERN, LVAL, batches, ...)                                                        \
    do {                                                                                           \
        static int grid_cap_ = 0;                                                                  \
        if (!grid_cap_) grid_cap_ = rb_grid(KERN<LVAL>, 1LL << 40);                                \
        const long long need_ = ((batches) + 7) / 8;                                               \
        KERN<LVAL><<<(int)(need_ < grid_cap_ ? need_ : grid_cap_), 256, 0, s>>>(__VA_ARGS__);      \
    } while (0)


// ---- launch wrappers ---------------------------------------------------------------------------
static inline int ew_blocks(long long elems) { long long b = (elems + 255) / 256; return (int)(b < 1 ? 1 : (b > 1024 ? 1024 : b)); }
static inline void check_count(int count) { if (count > 65535) throw Error{EF_ERR_BAD_SHAPE, "batch too large for grid.y (chunk it)"}; }

void launch_assemble_X(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    for (int off = 0; off < count; off += 65535) {
        int c = count - off < 65535 ? count - off : 65535;
        assemble_X_kernel<<<dim3(ew_blocks(8LL * n * n), c), 256, 0, s>>>(e + off, n);
    }
    EF_CUDA(cudaGetLastError());
}
void launch_assemble_Hc(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    for (int off = 0; off < count; off += 65535) {
        int c = count - off < 65535 ? count - off : 65535;
        assemble_Hc_kernel<<<dim3(ew_blocks(8LL * n * n), c), 256, 0, s>>>(e + off, n);
    }
    EF_CUDA(cudaGetLastError());
}
// Element-wise steps of the Newton-Schulz refinement of X^-1 (hps.cu: plan_refine), one batch entry per blockIdx.y:
//   mode 3:  E <- I + E   (E = -X X^-1 on entry) and max |E_ij| folded into *resid (non-negative doubles order like their bits)
//   mode 4:  dst <- dst + src
__global__ void __launch_bounds__(256) refine_ew_kernel(double* const* __restrict__ ptab, int nops, int mode, int dst_op, long long dst_off,
                                                        int src_op, long long src_off, int N, double* __restrict__ resid)
{
    double* dst = ptab[(size_t)blockIdx.y * nops + dst_op] + dst_off;
    const double* src = mode == 4 ? ptab[(size_t)blockIdx.y * nops + src_op] + src_off : nullptr;
    double mx = 0.0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N; idx += (long long)gridDim.x * blockDim.x) {
        if (mode == 3) {
            const int r = (int)(idx / N), c = (int)(idx - (long long)r * N);
            const double v = dst[idx] + (r == c ? 1.0 : 0.0);
            dst[idx] = v;
            mx = fmax(mx, fabs(v));
        } else dst[idx] += src[idx];
    }
    if (mode == 3 && resid) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(resid), (unsigned long long)__double_as_longlong(mx));
    }
}
void launch_refine_ew(double* const* ptab, int nops, int mode, int dst_op, long long dst_off, int src_op, long long src_off, int N, int batch,
                      double* resid, cudaStream_t s)
{
    for (int off = 0; off < batch; off += 65535) {
        const int c = batch - off < 65535 ? batch - off : 65535;
        refine_ew_kernel<<<dim3(ew_blocks((long long)N * N), c), 256, 0, s>>>(ptab + (size_t)off * nops, nops, mode, dst_op, dst_off, src_op, src_off, N, resid);
    }
    EF_CUDA(cudaGetLastError());
}
// Batched device-to-device copies (adaptive re-build: the operators of clean subtrees): blockIdx.y = copy, grid-stride over
// 16-byte words (every operator slab is a multiple of 16 bytes and 16-byte aligned)
__global__ void __launch_bounds__(256) copy_many_kernel(const CopyOp* __restrict__ ops)
{
    const CopyOp op = ops[blockIdx.y];
    const size_t n2 = op.n / 2;
    const double2* s2 = reinterpret_cast<const double2*>(op.src);
    double2* d2 = reinterpret_cast<double2*>(op.dst);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) d2[i] = s2[i];
    if ((op.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.dst[op.n - 1] = op.src[op.n - 1];
}
void launch_copy_many(const CopyOp* ops, int nops, cudaStream_t s)
{
    for (int off = 0; off < nops; off += 65535) {
        const int c = nops - off < 65535 ? nops - off : 65535;
        copy_many_kernel<<<dim3(64, c), 256, 0, s>>>(ops + off);
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

static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

static void launch_hdiff(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    for (int off = 0; off < count; off += 65535) {
        int c = count - off < 65535 ? count - off : 65535;
        hdiff_kernel<<<dim3(ew_blocks(4LL * n), c), 256, 0, s>>>(e + off, n);
    }
}

// Kernel choice per operator, by row length L (measured per tree level on B200, profiles/r1d_matvec_levels.md):
// rows of <= 256 doubles (power-of-two child side >= 16) take the row-batch kernels, longer rows one row per warp
// with 4 (X^-1, S) or 8 (the two half rows of the compact H) 16-byte loads in flight per lane.
static inline bool rb_ok(int n, int L) { return g_tuning[0] == 2 && n >= 16 && (n & (n - 1)) == 0 && L <= 256; }

void launch_upwards(const MergeEntry* e, int n, int count, cudaStream_t s)
{
    if (!count) return;
    const int logn = ilog2(n);
    const long long batches = (long long)count * 16 * n * n / 512;
    if (rb_ok(n, 4 * n)) {
        switch (4 * n) {
            case 64: RB_LAUNCH(rb_upwards_w_kernel, 64, batches, e, count, logn); break;
            case 128: RB_LAUNCH(rb_upwards_w_kernel, 128, batches, e, count, logn); break;
            default: RB_LAUNCH(rb_upwards_w_kernel, 256, batches, e, count, logn); break;
        }
    } else {
        launch_hdiff(e, n, count, s);
        const int rpc = pick_rows(4 * n, count);
        upwards_w_kernel<4><<<dim3(count, (4 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
    }
    if (rb_ok(n, 2 * n)) {
        switch (2 * n) {
            case 32: RB_LAUNCH(rb_upwards_h_kernel, 32, batches, e, count, logn); break;
            case 64: RB_LAUNCH(rb_upwards_h_kernel, 64, batches, e, count, logn); break;
            case 128: RB_LAUNCH(rb_upwards_h_kernel, 128, batches, e, count, logn); break;
            default: RB_LAUNCH(rb_upwards_h_kernel, 256, batches, e, count, logn); break;
        }
    } else {
        const int rpc = pick_rows(8 * n, count);
        if (g_tuning[1] == 0) upwards_h_kernel<8><<<dim3(count, (8 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
        else upwards_h_kernel<4><<<dim3(count, (8 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc);
    }
    EF_CUDA(cudaGetLastError());
}

void launch_solve_split(const MergeEntry* e, int n, int count, bool add_w, cudaStream_t s)
{
    if (!count) return;
    if (rb_ok(n, 8 * n)) {
        const int logn = ilog2(n);
        const long long batches = (long long)count * 32 * n * n / 512;
        if (8 * n == 128) RB_LAUNCH(rb_solve_split_kernel, 128, batches, e, count, logn, add_w ? 1 : 0);
        else RB_LAUNCH(rb_solve_split_kernel, 256, batches, e, count, logn, add_w ? 1 : 0);
    } else {
        const int rpc = pick_rows(4 * n, count);
        solve_split_kernel<4><<<dim3(count, (4 * n + rpc - 1) / rpc), 256, 0, s>>>(e, n, rpc, add_w ? 1 : 0);
    }
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
