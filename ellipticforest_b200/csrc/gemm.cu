// Descriptor-driven batched FP64 GEMM on the sm_100a FP64 tensor path (DMMA, mma.sync m8n8k4)
// and the small in-shared-memory Gauss-Jordan inverse used as the base case of the blocked
// inversion of the merge matrix X.
//
// Replaces, for the 4-to-1 merge of the reference (src/HPSAlgorithm.hpp:497-518):
//   createMatrixBlocks_ :799-859  (36 block gathers + 8 negations)  -> operand addressing + sign flip
//   mergeS_ :906-929   dgesv  (Matrix.hpp:944)                      -> GEMMs with X^-1
//   mergeT_ :940-968   dgemm  (Matrix.hpp:852) + T_LHS add          -> GEMM with additive C0
//   reorderOperators_ :979-993 (blockPermute)                       -> result addressing
//
// FP64 has no tcgen05 kind on Blackwell; the FP64 tensor throughput of sm_100a (37 TFLOP/s
// measured with a register-only DMMA loop on B200, cuBLAS DGEMM 35.4) is reached through
// mma.sync.m8n8k4.f64 fed from shared memory that is filled by an asynchronous multi-stage
// copy pipeline (cp.async 16-byte, LDGSTS).
#include "common.cuh"

#include <cstdlib>
#include <cooperative_groups.h>
#include <type_traits>

namespace efgpu {

int get_tuning(int key);   // vec.cu

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double x, unsigned neg) {
    return __hiloint2double(__double2hiint(x) ^ (int)neg, __double2loint(x));
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
struct GemmCfg {
    static constexpr int NT = WARPS_M * WARPS_N * 32;
    // Fragment loads: the four k slots of a DMMA may hold any four k as long as A and B agree, so lane t takes the PAIR
    // k = 2t, 2t+1 of each group of 8 with one 16-byte load of A and feeds it to two consecutive DMMAs (slots {0,2,4,6}
    // and {1,3,5,7}); B rows 2t and 2t+1 are read separately.  Paddings make both patterns bank-conflict free.
    static constexpr int LDA_S = BK + 8;   // == 8 (mod 16): rows g, g+1 of a quarter warp's 16-byte reads fall into disjoint bank halves
    static constexpr int LDB_S = BN + 2;   // == 2 (mod 8): rows 2t (and 2t+1), t < 4, start 4 banks apart
    static constexpr int A_STAGE = BM * LDA_S;
    static constexpr int B_STAGE = BK * LDB_S;
    // staged transposed second destination (peer mode): the tile's transpose, BN rows of BM doubles (+ 2 of padding), takes the
    // place of the pipeline stages once the K loop is over
    static constexpr int LDT_S = BM + 2;
    static constexpr int PIPE_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);
    static constexpr int T_BYTES = BN * LDT_S * (int)sizeof(double);
    static constexpr int SMEM_BYTES = PIPE_BYTES > T_BYTES ? PIPE_BYTES : T_BYTES;
    static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    static constexpr int FM = WM / 8, FN = WN / 8;
};

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32)
bgemm_kernel(double* const* __restrict__ ptab, int nops, const GemmBlock* __restrict__ blocks,
             int nblocks, int tiles_per_block, const PeerSpan ps)
{
    using C = GemmCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    // column slots of the B / C fragments: with an even number of 8-column tiles per warp, slot g of the tile pair (2 jg,
    // 2 jg + 1) stands for the adjacent columns 2g, 2g + 1 of a 16-column group, so one 16-byte load feeds both tiles
    constexpr bool PAIRED_N = C::FN % 2 == 0;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * C::A_STAGE;

    // work decode: tile fastest, then block descriptor, then batch entry
    const long long bid = blockIdx.x;
    const int tile = (int)(bid % tiles_per_block);
    const int blk = (int)((bid / tiles_per_block) % nblocks);
    const long long z = bid / ((long long)tiles_per_block * nblocks);
    const GemmBlock& bd = blocks[blk];
    const int tiles_n = bd.cols / BN;
    const int tm = tile / tiles_n, tn = tile % tiles_n;
    if (tm >= bd.rows / BM) return;
    double* const* ops = ptab + z * nops;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / WARPS_N) * C::WM, wn0 = (warp % WARPS_N) * C::WN;

    // flattened (term, k-tile) sequence; the per-term operand bases live in registers so that issuing a
    // stage costs no descriptor (global memory) reads on the critical path between two barriers
    const int nterms = bd.nterms;
    const int nk0 = bd.t[0].K / BK;
    const int nk_total = nk0 + (nterms > 1 ? bd.t[1].K / BK : 0);
    const int t1 = nterms > 1 ? 1 : 0;
    const double* Ab0 = ops[bd.t[0].a_op] + bd.t[0].a_off + (long long)(tm * BM) * bd.t[0].lda;
    const double* Bb0 = ops[bd.t[0].b_op] + bd.t[0].b_off + tn * BN;
    const double* Ab1 = ops[bd.t[t1].a_op] + bd.t[t1].a_off + (long long)(tm * BM) * bd.t[t1].lda;
    const double* Bb1 = ops[bd.t[t1].b_op] + bd.t[t1].b_off + tn * BN;
    const int lda0 = bd.t[0].lda, ldb0 = bd.t[0].ldb, lda1 = bd.t[t1].lda, ldb1 = bd.t[t1].ldb;
    const unsigned neg0 = bd.t[0].neg, neg1 = bd.t[t1].neg;

    auto load_stage = [&](int s, int buf) {
        const bool second = s >= nk0;
        const int kt = second ? s - nk0 : s;
        const int lda = second ? lda1 : lda0, ldb = second ? ldb1 : ldb0;
        const double* Ag = (second ? Ab1 : Ab0) + kt * BK;
        const double* Bg = (second ? Bb1 : Bb0) + (long long)(kt * BK) * ldb;
        double* as = As + buf * C::A_STAGE;
        double* bs = Bs + buf * C::B_STAGE;
        constexpr int A_CHUNKS = BM * BK / 2, ACPR = BK / 2;
#pragma unroll
        for (int c = tid; c < A_CHUNKS; c += C::NT) {
            int r = c / ACPR, cc = (c % ACPR) * 2;
            cp_async16(as + r * C::LDA_S + cc, Ag + (long long)r * lda + cc);
        }
        constexpr int B_CHUNKS = BK * BN / 2, BCPR = BN / 2;
#pragma unroll
        for (int c = tid; c < B_CHUNKS; c += C::NT) {
            int r = c / BCPR, cc = (c % BCPR) * 2;
            cp_async16(bs + r * C::LDB_S + cc, Bg + (long long)r * ldb + cc);
        }
    };

    double acc[C::FM][C::FN][2];
#pragma unroll
    for (int i = 0; i < C::FM; i++)
#pragma unroll
        for (int j = 0; j < C::FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk_total) load_stage(s, s);
        cp_async_commit();
    }

    // The sign of a term is applied to the accumulators, not to every A fragment: raw products are accumulated, the
    // accumulators are negated where the sign changes between the two terms and once more at the end if the last term is
    // negative (exact, so the result is bit-identical to flipping A; saves one LOP3 per fragment on the LDS -> DMMA path).
    const bool flip_mid = nterms > 1 && neg0 != neg1;
    for (int kt = 0; kt < nk_total; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + STAGES - 1;
            if (nxt < nk_total) load_stage(nxt, nxt % STAGES);
            cp_async_commit();
        }
        if (flip_mid && kt == nk0) {
#pragma unroll
            for (int i = 0; i < C::FM; i++)
#pragma unroll
                for (int j = 0; j < C::FN; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
        }
        const double* as = As + (kt % STAGES) * C::A_STAGE + (wm0 + (lane >> 2)) * C::LDA_S + 2 * (lane & 3);
        const double* bs = Bs + (kt % STAGES) * C::B_STAGE + 2 * (lane & 3) * C::LDB_S + wn0 + (PAIRED_N ? 2 : 1) * (lane >> 2);
#pragma unroll
        for (int kp = 0; kp < BK / 8; kp++) {
            double2 a[C::FM];
            double b0[C::FN], b1[C::FN];
#pragma unroll
            for (int i = 0; i < C::FM; i++) a[i] = *reinterpret_cast<const double2*>(as + i * 8 * C::LDA_S + kp * 8);
            if constexpr (PAIRED_N) {
#pragma unroll
                for (int jg = 0; jg < C::FN / 2; jg++) {
                    const double2 v0 = *reinterpret_cast<const double2*>(bs + kp * 8 * C::LDB_S + jg * 16);
                    const double2 v1 = *reinterpret_cast<const double2*>(bs + (kp * 8 + 1) * C::LDB_S + jg * 16);
                    b0[2 * jg] = v0.x; b0[2 * jg + 1] = v0.y; b1[2 * jg] = v1.x; b1[2 * jg + 1] = v1.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C::FN; j++) {
                    b0[j] = bs[kp * 8 * C::LDB_S + j * 8];
                    b1[j] = bs[(kp * 8 + 1) * C::LDB_S + j * 8];
                }
            }
#pragma unroll
            for (int i = 0; i < C::FM; i++)
#pragma unroll
                for (int j = 0; j < C::FN; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b0[j]);
#pragma unroll
            for (int i = 0; i < C::FM; i++)
#pragma unroll
                for (int j = 0; j < C::FN; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b1[j]);
        }
    }
    cp_async_wait<0>();
    if (neg1) {
#pragma unroll
        for (int i = 0; i < C::FM; i++)
#pragma unroll
            for (int j = 0; j < C::FN; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
    }

    // epilogue: C = acc (+ C0); each thread owns rows lane/4 + 8 i.  Unpaired columns: the pair (lane%4)*2 of every 8-column
    // tile; paired: the four adjacent columns 4 (lane%4) .. + 3 of every 16-column group (slots 0/1 of its two tiles interleaved)
    const int col0 = tn * BN + wn0 + (PAIRED_N ? 4 : 2) * (lane & 3);
    double* Cg = ops[bd.c_op] + bd.c_off + (long long)(tm * BM + wm0 + (lane >> 2)) * bd.ldc + col0;
    // second destination: the signed transpose of the block (lanes lane/4 = 0..7 of a fragment write 8 consecutive doubles)
    // Direct form: every thread stores its elements (8-byte stores, 64 contiguous bytes per quarter warp) - fine in local HBM.  Staged
    // form (ps.stage_t; default where peers receive the block): the transposed tile is assembled in shared memory and leaves as whole
    // rows of BM doubles, 16 bytes per lane - scattered 8-byte stores over NVLink cost 2.5 x the whole inversion at 8 ranks (r2k).
    const bool staged_t = bd.ct_op1 != 0 && ps.stage_t != 0;
    double* Ctg = (bd.ct_op1 && !staged_t) ? ops[bd.ct_op1 - 1] + bd.ct_off + (long long)col0 * bd.ldct + (tm * BM + wm0 + (lane >> 2)) : nullptr;
    double* Cts = staged_t ? smem + (wn0 + (PAIRED_N ? 4 : 2) * (lane & 3)) * C::LDT_S + wm0 + (lane >> 2) : nullptr;
    const unsigned ctn = bd.ct_neg;
    auto store_t = [&](double* p, double v) {
        v = flip_sign(v, ctn);
        *p = v;
        for (int r = 0; r < ps.n; r++)
            if (r != ps.me) *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + ps.delta[r]) = v;
    };
    if (staged_t) __syncthreads();   // every warp is done with the last pipeline stage
    const double* C0g = nullptr;
    if (bd.c0_op >= 0) C0g = ops[bd.c0_op] + bd.c0_off + (long long)(tm * BM + wm0 + (lane >> 2)) * bd.ldc0 + col0;
#pragma unroll
    for (int i = 0; i < C::FM; i++) {
        if constexpr (PAIRED_N) {
#pragma unroll
            for (int jg = 0; jg < C::FN / 2; jg++) {
                double2 lo = make_double2(acc[i][2 * jg][0], acc[i][2 * jg + 1][0]);
                double2 hi = make_double2(acc[i][2 * jg][1], acc[i][2 * jg + 1][1]);
                if (C0g) {
                    const double2* c0 = reinterpret_cast<const double2*>(C0g + (long long)(i * 8) * bd.ldc0 + jg * 16);
                    const double2 c0l = c0[0], c0h = c0[1];
                    lo.x += c0l.x; lo.y += c0l.y; hi.x += c0h.x; hi.y += c0h.y;
                }
                double2* c = reinterpret_cast<double2*>(Cg + (long long)(i * 8) * bd.ldc + jg * 16);
                c[0] = lo; c[1] = hi;
                for (int r = 0; r < ps.n; r++)     // row-partitioned tree: the same tile into every peer's arena (NVLink stores)
                    if (r != ps.me) {
                        double2* cp = reinterpret_cast<double2*>(reinterpret_cast<char*>(c) + ps.delta[r]);
                        cp[0] = lo; cp[1] = hi;
                    }
                if (Ctg) {
                    double* t = Ctg + (long long)(jg * 16) * bd.ldct + i * 8;
                    store_t(t, lo.x); store_t(t + bd.ldct, lo.y); store_t(t + 2LL * bd.ldct, hi.x); store_t(t + 3LL * bd.ldct, hi.y);
                } else if (Cts) {
                    double* t = Cts + (jg * 16) * C::LDT_S + i * 8;
                    t[0] = flip_sign(lo.x, ctn); t[C::LDT_S] = flip_sign(lo.y, ctn);
                    t[2 * C::LDT_S] = flip_sign(hi.x, ctn); t[3 * C::LDT_S] = flip_sign(hi.y, ctn);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < C::FN; j++) {
                double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                if (C0g) {
                    double2 c0 = *reinterpret_cast<const double2*>(C0g + (long long)(i * 8) * bd.ldc0 + j * 8);
                    v.x += c0.x; v.y += c0.y;
                }
                double2* c = reinterpret_cast<double2*>(Cg + (long long)(i * 8) * bd.ldc + j * 8);
                *c = v;
                for (int r = 0; r < ps.n; r++)
                    if (r != ps.me) *reinterpret_cast<double2*>(reinterpret_cast<char*>(c) + ps.delta[r]) = v;
                if (Ctg) {
                    double* t = Ctg + (long long)(j * 8) * bd.ldct + i * 8;
                    store_t(t, v.x); store_t(t + bd.ldct, v.y);
                } else if (Cts) {
                    double* t = Cts + (j * 8) * C::LDT_S + i * 8;
                    t[0] = flip_sign(v.x, ctn); t[C::LDT_S] = flip_sign(v.y, ctn);
                }
            }
        }
    }
    if (staged_t) {   // rows of the transposed tile = columns tn BN .. of the block, BM contiguous doubles each
        __syncthreads();
        double* Ctb = ops[bd.ct_op1 - 1] + bd.ct_off + (long long)(tn * BN) * bd.ldct + tm * BM;
        constexpr int V = BM / 2;
        for (int idx = tid; idx < BN * V; idx += C::NT) {
            const int c = idx / V, v2 = (idx % V) * 2;
            const double2 val = *reinterpret_cast<const double2*>(smem + c * C::LDT_S + v2);
            double2* dst = reinterpret_cast<double2*>(Ctb + (long long)c * bd.ldct + v2);
            *dst = val;
            for (int r = 0; r < ps.n; r++)
                if (r != ps.me) *reinterpret_cast<double2*>(reinterpret_cast<char*>(dst) + ps.delta[r]) = val;
        }
    }
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
static void launch_cfg(double* const* ptab, int nops, const GemmBlock* d_blocks, int nblocks, int batch,
                       int max_tiles, cudaStream_t stream, const PeerSpan& ps, const GemmBlock* h_blocks = nullptr)
{
    if (h_blocks) {   // non-square CTA tiles: recount the tiles of the largest block
        max_tiles = 0;
        for (int b = 0; b < nblocks; b++) { int v = (h_blocks[b].rows / BM) * (h_blocks[b].cols / BN); if (v > max_tiles) max_tiles = v; }
    }
    using C = GemmCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    auto kern = bgemm_kernel<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared))
        EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    long long grid = (long long)max_tiles * nblocks * batch;
    if (grid <= 0) return;
    if (grid > 2147483647LL) throw Error{EF_ERR_BAD_SHAPE, "bgemm grid too large"};
    kern<<<(unsigned)grid, C::NT, C::SMEM_BYTES, stream>>>(ptab, nops, d_blocks, nblocks, max_tiles, ps);
    EF_CUDA(cudaGetLastError());
}

// Blocks that are not made of 8 x 8 tensor-core tiles (4 x 4 blocks, K = 4: the first merge level above 4 x 4 leaf patches, which the
// reference's convergence driver uses, examples/elliptic-multiple/main.cpp:444): one thread per result element, plain FMAs.  Same
// descriptor semantics as bgemm_kernel (two terms, additive C0, signed transposed second destination, peer arenas).
__global__ void __launch_bounds__(256)
bgemm_scalar_kernel(double* const* __restrict__ ptab, int nops, const GemmBlock* __restrict__ blocks, int nblocks, int elems_per_block,
                    long long total, const PeerSpan ps)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx % elems_per_block);
        const int blk = (int)((idx / elems_per_block) % nblocks);
        const long long z = idx / ((long long)elems_per_block * nblocks);
        const GemmBlock& bd = blocks[blk];
        if (e >= bd.rows * bd.cols) continue;
        const int r = e / bd.cols, c = e % bd.cols;
        double* const* ops = ptab + z * nops;
        double acc = 0.0;
        for (int t = 0; t < bd.nterms; t++) {
            const double* A = ops[bd.t[t].a_op] + bd.t[t].a_off + (long long)r * bd.t[t].lda;
            const double* B = ops[bd.t[t].b_op] + bd.t[t].b_off + c;
            double sum = 0.0;
            for (int k = 0; k < bd.t[t].K; k++) sum = fma(A[k], B[(long long)k * bd.t[t].ldb], sum);
            acc += bd.t[t].neg ? -sum : sum;
        }
        if (bd.c0_op >= 0) acc += ops[bd.c0_op][bd.c0_off + (long long)r * bd.ldc0 + c];
        double* cp = ops[bd.c_op] + bd.c_off + (long long)r * bd.ldc + c;
        *cp = acc;
        for (int q = 0; q < ps.n; q++)
            if (q != ps.me) *reinterpret_cast<double*>(reinterpret_cast<char*>(cp) + ps.delta[q]) = acc;
        if (bd.ct_op1) {
            double* tp = ops[bd.ct_op1 - 1] + bd.ct_off + (long long)c * bd.ldct + r;
            const double v = flip_sign(acc, bd.ct_neg);
            *tp = v;
            for (int q = 0; q < ps.n; q++)
                if (q != ps.me) *reinterpret_cast<double*>(reinterpret_cast<char*>(tp) + ps.delta[q]) = v;
        }
    }
}

void launch_bgemm(double* const* ptab, int nops, const GemmBlock* d_blocks, const GemmBlock* h_blocks,
                  int nblocks, int batch, cudaStream_t stream, int force_tile, const PeerSpan* peers, const TmaArgs* tma)
{
    if (nblocks == 0 || batch == 0) return;
    PeerSpan ps = peers ? *peers : PeerSpan{};
    // transposed second destinations (tuning key 4): 0 = staged through shared memory where peers receive them, direct otherwise;
    // 1 = always staged; 2 = never
    ps.stage_t = get_tuning(4) == 1 || (get_tuning(4) == 0 && ps.n > 1) ? 1 : 0;
    // largest power-of-two tile dividing every block dimension
    int g = 128;
    bool k16 = true, k8 = true;
    for (int b = 0; b < nblocks; b++) {
        const GemmBlock& bd = h_blocks[b];
        while (g > 1 && (bd.rows % g || bd.cols % g)) g >>= 1;
        for (int t = 0; t < bd.nterms; t++) {
            if (bd.t[t].K % 16) k16 = false;
            if (bd.t[t].K % 8) k8 = false;
        }
    }
    if (g < 8 || !k8) {   // not made of tensor-core tiles: scalar kernel
        int epb = 0;
        for (int b = 0; b < nblocks; b++) epb = epb > h_blocks[b].rows * h_blocks[b].cols ? epb : h_blocks[b].rows * h_blocks[b].cols;
        const long long total = (long long)epb * nblocks * batch;
        if (total <= 0) return;
        const long long want = (total + 255) / 256;
        bgemm_scalar_kernel<<<(unsigned)(want < 148 * 32 ? want : 148 * 32), 256, 0, stream>>>(ptab, nops, d_blocks, nblocks, epb, total, ps);
        EF_CUDA(cudaGetLastError());
        return;
    }
    if (!k16 && g > 8) g = 8;
    auto tiles_for = [&](int t) { int m = 0; for (int b = 0; b < nblocks; b++) { int v = (h_blocks[b].rows / t) * (h_blocks[b].cols / t); if (v > m) m = v; } return m; };
    int tile = g;
    if (force_tile) tile = force_tile < g ? force_tile : g;
    else {
        // prefer the largest tile that still fills the 148 SMs; never go below 32 for that reason
        while (tile > 32 && (long long)tiles_for(tile) * nblocks * batch < 148) tile >>= 1;
    }
    int mt = tiles_for(tile);
    switch (tile) {
        case 128: {
            static const int variant = [] { const char* e = getenv("EFGPU_GEMM_VARIANT"); return e ? atoi(e) : 7; }();
            // operands staged by TMA where the caller prepared tensor maps for every block of the launch (tuning key 8, default on):
            // 36.3 instead of 33.5 TFLOP/s on 4096^3, tensor pipe 98.7 % instead of 91.2 % active (profiles/r2n_*); bit-identical
            if (tma && tma->d_tblocks && variant == 7 && get_tuning(8) == 1) {
                launch_bgemm_tma(ptab, nops, d_blocks, h_blocks, nblocks, batch, stream, ps, *tma);
                break;
            }
            switch (variant) {   // tuning variants of the 128 x 128 CTA tile (tools/gemm_bench.py)
                case 1: launch_cfg<128, 128, 16, 4, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                case 2: launch_cfg<128, 128, 32, 2, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                case 3: launch_cfg<128, 128, 32, 4, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                case 4: launch_cfg<128, 128, 16, 2, 4, 4>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                case 5: launch_cfg<128, 128, 16, 4, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                case 6: launch_cfg<128, 128, 32, 4, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                case 7: launch_cfg<128, 64, 16, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 8: launch_cfg<64, 128, 16, 1, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 9: launch_cfg<128, 64, 32, 2, 2, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 10: launch_cfg<128, 64, 16, 2, 2, 4>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 11: launch_cfg<64, 64, 16, 1, 2, 4>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 12: launch_cfg<128, 64, 32, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 13: launch_cfg<128, 64, 16, 2, 2, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 14: launch_cfg<64, 128, 16, 1, 4, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 15: launch_cfg<128, 64, 16, 4, 1, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 16: launch_cfg<64, 128, 16, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 17: launch_cfg<64, 128, 32, 1, 4, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
                case 18: launch_cfg<128, 64, 16, 4, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;   // 8 warps of 32 x 32, 2 CTAs / SM
                case 19: launch_cfg<128, 64, 16, 4, 2, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;   // ... 2 stages, 3 CTAs / SM
                case 20: launch_cfg<64, 64, 16, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;    // 4 warps of 32 x 32, 4 CTAs / SM
                case 21: launch_cfg<128, 64, 32, 4, 2, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;   // BK 32: half the barriers
                case 22: launch_cfg<128, 128, 16, 4, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;            // 16 warps of 32 x 32, 1 CTA / SM... 2 if registers allow
                case 23: launch_cfg<64, 128, 16, 2, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;   // 8 warps of 32 x 32, wide
                case 0: launch_cfg<128, 128, 16, 2, 4, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
                // default (7): 128 x 64 CTA tile, 4 warps of 64 x 32, two CTAs per SM so that one CTA's barrier / fragment-load
                // phases overlap the other's DMMA phases (measured 31-32 TFLOP/s vs 29-30 for one 128 x 128 CTA per SM)
                default: launch_cfg<128, 64, 16, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps, h_blocks); break;
            }
            break;
        }
        case 64: launch_cfg<64, 64, 16, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
        case 32: launch_cfg<32, 32, 16, 2, 2, 3>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
        case 16: launch_cfg<16, 16, 16, 1, 1, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
        case 8: launch_cfg<8, 8, 8, 1, 1, 2>(ptab, nops, d_blocks, nblocks, batch, mt, stream, ps); break;
        default: throw Error{EF_ERR_BAD_SHAPE, "bgemm: unsupported tile"};
    }
}

// ---------------------------------------------------------------------------------------------
// Batched block transpose (32 x 32 tiles through shared memory, both sides coalesced).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
btranspose_kernel(double* const* __restrict__ ptab, int nops, const TransOp* __restrict__ tops, int nt, int tiles_per_op)
{
    __shared__ double tile[32][33];
    const long long bid = blockIdx.x;
    const int t = (int)(bid % tiles_per_op);
    const int o = (int)((bid / tiles_per_op) % nt);
    const long long z = bid / ((long long)tiles_per_op * nt);
    const TransOp op = tops[o];
    const int tiles_c = (op.cols + 31) >> 5;
    const int tr = t / tiles_c, tc = t % tiles_c;
    if (tr >= ((op.rows + 31) >> 5)) return;
    const double* src = ptab[z * nops + op.src_op] + op.src_off;
    double* dst = ptab[z * nops + op.dst_op] + op.dst_off;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = tr * 32 + ty + 8 * j, c = tc * 32 + tx;
        if (r < op.rows && c < op.cols) tile[ty + 8 * j][tx] = src[(long long)r * op.lds + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = tc * 32 + ty + 8 * j, c = tr * 32 + tx;   // destination row = source column
        if (r < op.cols && c < op.rows) dst[(long long)r * op.ldd + c] = flip_sign(tile[tx][ty + 8 * j], op.neg);
    }
}

void launch_btranspose(double* const* ptab, int nops, const TransOp* d_ops, const TransOp* h_ops, int nt, int batch,
                       cudaStream_t stream)
{
    if (nt == 0 || batch == 0) return;
    int tiles = 0;
    for (int i = 0; i < nt; i++) { int v = ((h_ops[i].rows + 31) / 32) * ((h_ops[i].cols + 31) / 32); if (v > tiles) tiles = v; }
    long long grid = (long long)tiles * nt * batch;
    if (grid > 2147483647LL) throw Error{EF_ERR_BAD_SHAPE, "btranspose grid too large"};
    btranspose_kernel<<<(unsigned)grid, 256, 0, stream>>>(ptab, nops, d_ops, nt, tiles);
    EF_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// Small in-place inverse (base case of the blocked inversion of X): one CTA per matrix,
// Gauss-Jordan in shared memory.
// ---------------------------------------------------------------------------------------------
// Pivot tracker of the unpivoted base case (one per thread, every thread of a CTA sees every pivot): smallest and largest
// |pivot| of this block and the number of negative pivots (an SPD merge matrix has none).  publish() folds them into the
// handle's tracker: [0] global min |pivot|, [1] global max |pivot|, [2] smallest per-block ratio min / max, [3] count of
// negative pivots (unsigned 64-bit).  Non-negative doubles order like their bit patterns.
struct PivotTrack {
    double mn = 1e300, mx = 0.0; unsigned neg = 0;
    __device__ __forceinline__ void see(double piv) { const double a = fabs(piv); mn = fmin(mn, a); mx = fmax(mx, a); neg += piv < 0.0 ? 1u : 0u; }
    __device__ __forceinline__ void publish(double* t) const {
        unsigned long long* u = reinterpret_cast<unsigned long long*>(t);
        atomicMin(u + 0, (unsigned long long)__double_as_longlong(mn));
        atomicMax(u + 1, (unsigned long long)__double_as_longlong(mx));
        atomicMin(u + 2, (unsigned long long)__double_as_longlong(mx > 0.0 ? mn / mx : 0.0));
        if (neg) atomicAdd(u + 3, (unsigned long long)neg);
    }
};

__global__ void pivot_tracker_reset_kernel(double* t)
{
    unsigned long long* u = reinterpret_cast<unsigned long long*>(t);
    t[0] = 1e300; t[1] = 0.0; t[2] = 1e300; u[3] = 0ull; t[4] = 0.0;
}
void launch_pivot_tracker_reset(double* tracker, cudaStream_t s)
{
    pivot_tracker_reset_kernel<<<1, 1, 0, s>>>(tracker);
    EF_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
invert_small_kernel(double* const* __restrict__ ptab, int nops, int op, long long off, long long off2, int ld, int N,
                    double* __restrict__ min_pivot)
{
    const int nmat = off2 >= 0 ? 2 : 1;
    extern __shared__ __align__(16) double sm[];
    const int LDS = N + 1;
    double* A = sm;                 // N x (N+1)
    double* colk = sm + N * LDS;    // N  (column k before the update)
    double* G = ptab[(long long)(blockIdx.x / nmat) * nops + op] + ((blockIdx.x % nmat) ? off2 : off);
    const int tid = threadIdx.x, NT = blockDim.x;
    for (int e = tid; e < N * N; e += NT) { int r = e / N, c = e % N; A[r * LDS + c] = G[(long long)r * ld + c]; }
    __syncthreads();
    PivotTrack pt;
    for (int k = 0; k < N; k++) {
        const double piv = A[k * LDS + k];
        const double p = 1.0 / piv;
        pt.see(piv);
        for (int r = tid; r < N; r += NT) colk[r] = A[r * LDS + k];
        __syncthreads();
        // scale pivot row
        for (int c = tid; c < N; c += NT) A[k * LDS + c] = (c == k) ? p : A[k * LDS + c] * p;
        __syncthreads();
        // eliminate column k from all other rows
        for (int e = tid; e < N * N; e += NT) {
            int r = e / N, c = e % N;
            if (r == k) continue;
            double f = colk[r];
            A[r * LDS + c] = (c == k) ? -f * p : A[r * LDS + c] - f * A[k * LDS + c];
        }
        __syncthreads();
    }
    for (int e = tid; e < N * N; e += NT) { int r = e / N, c = e % N; G[(long long)r * ld + c] = A[r * LDS + c]; }
    if (tid == 0 && min_pivot) pt.publish(min_pivot);
}

// Register-resident variant: 256 threads as a 16 x 16 grid, thread (tr, tc) owns the elements
// (tr + 16 i, tc + 16 j), i, j < TI (cyclic), N = 16 TI.  Gauss-Jordan without pivoting:
//   row k:      a_kj <- a_kj / a_kk (j != k),  a_kk <- 1 / a_kk
//   other rows: a_ij <- [j != k] a_ij - a_ik * (new a_kj)
// Only two warps per scheduler fit (~200 registers), so the serial chain of a pivot must not stall them:
//  * split barrier: one mbarrier phase per pivot, a warp ARRIVES as soon as it has read row / column k (and published
//    its share of k+1) and only WAITS at the top of the next pivot, after its rank-1 update; row / column buffers are
//    double-buffered;
//  * look-ahead: while pivot k is applied, the owners of row k+1 and column k+1 first form what those will hold after
//    pivot k (on temporaries), the owner of a_{k+1,k+1} takes the reciprocal ONCE and hands it to its half warp by
//    shuffle, and the row is published already scaled - nobody else divides or scales, and the division plus the
//    shared-memory round trip of pivot k+1 overlap the rank-1 update of pivot k;
//  * the pivot loop is unrolled over the register index k / 16, so every register subscript is a compile-time
//    constant (no dispatch chains).
template <int TI, int KI, int K1I>
__device__ __forceinline__ void gj_pivot(double (&a)[TI][TI], double (*sRow)[16 * TI], double (*sCol)[16 * TI], double (*sPiv)[2],
                                         unsigned long long* bar, int kr, int tr, int tc, bool last, PivotTrack& minp)
{
    const int b = kr & 1;                                   // N is a multiple of 16: k and kr have the same parity
    mbar_wait(bar, (unsigned)b);                            // phase k: row / column k are published, buffer b ^ 1 is free
    const double p = sPiv[b][1];
    minp.see(sPiv[b][0]);
    double rk[TI], ck[TI];
#pragma unroll
    for (int j = 0; j < TI; j++) rk[j] = sRow[b][tc + 16 * j];   // row k, already divided by the pivot
#pragma unroll
    for (int i = 0; i < TI; i++) ck[i] = sCol[b][tr + 16 * i];
    if (!last) {
        const int k1r = (kr + 1) & 15;
        if (tr == k1r) {                                    // one half warp: row k+1 after this pivot
            double v[TI];
#pragma unroll
            for (int j = 0; j < TI; j++) {
                v[j] = fma(-ck[K1I], rk[j], a[K1I][j]);
                if (tc == kr && j == KI) v[j] = -ck[K1I] * p;           // its entry in pivot column k
            }
            const double diag = v[K1I];                     // meaningful in the lane with tc == k1r
            const double rinv = 1.0 / diag;
            const double pn = __shfl_sync(0xffffu << (threadIdx.x & 16), rinv, (threadIdx.x & 16) | k1r);   // within the half warp that owns the row
            if (tc == k1r) { sPiv[b ^ 1][0] = diag; sPiv[b ^ 1][1] = pn; }
#pragma unroll
            for (int j = 0; j < TI; j++) sRow[b ^ 1][tc + 16 * j] = v[j] * pn;
        }
        if (tc == k1r) {                                    // column k+1 after this pivot
#pragma unroll
            for (int i = 0; i < TI; i++) {
                double v = fma(-ck[i], rk[K1I], a[i][K1I]);
                if (tr == kr && i == KI) v = rk[K1I];                   // its entry in pivot row k
                sCol[b ^ 1][tr + 16 * i] = v;
            }
        }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);          // this warp is done with buffer b and has published its part of b ^ 1
    // rank-1 update of every element ...
#pragma unroll
    for (int i = 0; i < TI; i++)
#pragma unroll
        for (int j = 0; j < TI; j++) a[i][j] = fma(-ck[i], rk[j], a[i][j]);
    // ... then repair column k (wanted -a_ik p) and row k (wanted a_kj p, and p on the diagonal)
    if (tc == kr) {
#pragma unroll
        for (int i = 0; i < TI; i++) a[i][KI] = -ck[i] * p;
    }
    if (tr == kr) {
#pragma unroll
        for (int j = 0; j < TI; j++) a[KI][j] = rk[j];
        if (tc == kr) a[KI][KI] = p;
    }
}

template <int TI, int KI>
__device__ __forceinline__ void gj_block(double (&a)[TI][TI], double (*sRow)[16 * TI], double (*sCol)[16 * TI], double (*sPiv)[2],
                                         unsigned long long* bar, int tr, int tc, PivotTrack& minp)
{
    if constexpr (KI < TI) {
        for (int kr = 0; kr < 15; kr++) gj_pivot<TI, KI, KI>(a, sRow, sCol, sPiv, bar, kr, tr, tc, false, minp);
        gj_pivot<TI, KI, (KI + 1 < TI ? KI + 1 : KI)>(a, sRow, sCol, sPiv, bar, 15, tr, tc, KI + 1 == TI, minp);
        gj_block<TI, KI + 1>(a, sRow, sCol, sPiv, bar, tr, tc, minp);
    }
}

template <int TI>
__global__ void __launch_bounds__(256)
invert_reg_kernel(double* const* __restrict__ ptab, int nops, int op, long long off, long long off2, int ld, double* __restrict__ min_pivot)
{
    constexpr int N = 16 * TI;
    const int nmat = off2 >= 0 ? 2 : 1;
    __shared__ double sRow[2][N];
    __shared__ double sCol[2][N];
    __shared__ double sPiv[2][2];   // [buffer][0: pivot, 1: its reciprocal]
    __shared__ unsigned long long bar;   // one phase per pivot, 8 arrivals (one per warp)
    double* G = ptab[(long long)(blockIdx.x / nmat) * nops + op] + ((blockIdx.x % nmat) ? off2 : off);
    const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    double a[TI][TI];
#pragma unroll
    for (int i = 0; i < TI; i++)
#pragma unroll
        for (int j = 0; j < TI; j++) a[i][j] = G[(long long)(tr + 16 * i) * ld + tc + 16 * j];
    // pivot 0 is published from the loaded values
    if (tr == 0) {
        const double rinv = 1.0 / a[0][0];
        const double pn = __shfl_sync(0xffffu, rinv, 0);
#pragma unroll
        for (int j = 0; j < TI; j++) sRow[0][tc + 16 * j] = a[0][j] * pn;
        if (tc == 0) { sPiv[0][0] = a[0][0]; sPiv[0][1] = pn; }
    }
    if (tc == 0) {
#pragma unroll
        for (int i = 0; i < TI; i++) sCol[0][tr + 16 * i] = a[i][0];
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&bar);         // phase 0: row / column 0 are published
    PivotTrack minp;
    gj_block<TI, 0>(a, sRow, sCol, sPiv, &bar, tr, tc, minp);
#pragma unroll
    for (int i = 0; i < TI; i++)
#pragma unroll
        for (int j = 0; j < TI; j++) G[(long long)(tr + 16 * i) * ld + tc + 16 * j] = a[i][j];
    if (threadIdx.x == 0 && min_pivot) minp.publish(min_pivot);
}

// ---------------------------------------------------------------------------------------------
// Blocked Gauss-Jordan inverse of a 128 x 128 block on the FP64 tensor pipe (round 2).
//
// invert_reg_kernel pays one barrier phase per pivot: 128 dependent phases of ~0.57 us, FP64 pipe 31 % busy, 73 us per matrix -
// and at the top tree levels these base cases form a chain of N/128 sequential launches.  Here the pivots are taken eight at a
// time.  The matrix lives in registers as 16 x 16 accumulator tiles of mma.m8n8k4 (8 warps as a 2 x 4 grid, 8 x 4 tiles each).
// Block step kb with pivot block P = A[K,K] (K = 8 kb .. 8 kb + 7), row panel R = A[K,:], column panel C = A[:,K]:
//     A[K,K] <- P^-1,   A[K,J] <- P^-1 R_J,   A[I,K] <- -C_I P^-1,   A[I,J] <- A[I,J] - C_I (P^-1 R_J)
//  1. the owners publish R (8 x 128) and C (128 x 8) to shared memory (double buffered: one __syncthreads per block step);
//  2. every warp inverts P itself (8 x 8 Gauss-Jordan on its lanes by shuffles: no second barrier);
//  3. (P^-1 R_J)^T = R_J^T P^-T comes out of a DMMA in accumulator layout, which - with the k slots of the next DMMA pair
//     permuted to {0,2,4,6} / {1,3,5,7} - IS the B-operand layout of the trailing update: no layout change through memory;
//  4. trailing update: two DMMAs per tile, A operand = -C_I read as 16-byte pairs.
// No pivoting, like the kernel it replaces (pivoting policy); different rounding (block order), same parity tests.
// Reciprocal of a pivot off the IEEE division sequence: hardware seed (20 bits) + two Newton steps = four dependent DFMAs instead of
// MUFU + ~7 DFMAs + the slow-path check, on the critical path of every one of the 128 sequential pivots of a base case.  Relative
// error ~1 ulp (not correctly rounded); pivots are never subnormal or near overflow (the tracker rejects such blocks).
__device__ __forceinline__ double fast_rcp(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0); r = fma(r, e, r);
    e = fma(-a, r, 1.0); r = fma(r, e, r);
    return r;
}
template <bool FR = false>
__device__ __forceinline__ void inv8x8_pivot(int k, double& p0, double& p1, int r, int q, PivotTrack& pt)
{
    const int kq = k >> 1;
    const double rk0 = __shfl_sync(0xffffffffu, p0, 4 * k + q), rk1 = __shfl_sync(0xffffffffu, p1, 4 * k + q);    // row k at my columns
    const double mine = (k & 1) ? p1 : p0;
    const double ck = __shfl_sync(0xffffffffu, mine, 4 * r + kq);                                                  // A[r][k]
    const double piv = __shfl_sync(0xffffffffu, mine, 4 * k + kq);                                                 // A[k][k]
    const double pinv = FR ? fast_rcp(piv) : 1.0 / piv;
    pt.see(piv);
    if (r == k) {
        p0 = (2 * q == k) ? pinv : rk0 * pinv;
        p1 = (2 * q + 1 == k) ? pinv : rk1 * pinv;
    } else {
        const double f = ck * pinv;
        p0 = (2 * q == k) ? -f : fma(-f, rk0, p0);
        p1 = (2 * q + 1 == k) ? -f : fma(-f, rk1, p1);
    }
}
template <bool FR = false>
__device__ __forceinline__ void inv8x8_warp(double& p0, double& p1, int r, int q, PivotTrack& pt)
{
#pragma unroll
    for (int k = 0; k < 8; k++) inv8x8_pivot<FR>(k, p0, p1, r, q, pt);
}

// LA (look-ahead, efgpu_set_tuning(7, 2)): the inverse of the NEXT pivot block is taken off the critical path.  The owner of tile (kb+1, kb+1)
// publishes it with the panels of step kb; every warp forms its updated value itself (four DMMAs) and interleaves the eight
// pivots of its 8 x 8 inverse with the eight tile rows of the trailing update, so that the dependent shuffle / reciprocal chain
// runs in the shadow of the tensor-pipe work and step kb+1 starts with P^-1 already in registers.
template <bool LA, bool FR = false>
__global__ void __launch_bounds__(256, 1)
invert_blk128_kernel(double* const* __restrict__ ptab, int nops, int op, long long off, long long off2, int ld, double* __restrict__ min_pivot)
{
    constexpr int LR = 132;     // row panel 8 x 128, row stride == 4 (mod 16): the fragment reads (4 k' rows x 4 columns per half warp) are conflict free
    constexpr int LC = 8;       // column panel 128 x 8, unpadded: the 16-byte pair reads of a quarter warp (2 rows x 4 pairs) cover 128 contiguous bytes
    constexpr int LP = 10;
    __shared__ __align__(16) double sR[2][8 * LR];
    __shared__ __align__(16) double sC[2][128 * LC];
    __shared__ __align__(16) double sPi[8][8 * LP];
    __shared__ __align__(16) double sD[2][8 * LP];      // LA: the next pivot block before this step's update
    const int nmat = off2 >= 0 ? 2 : 1;
    double* G = ptab[(long long)(blockIdx.x / nmat) * nops + op] + ((blockIdx.x % nmat) ? off2 : off);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wr = w >> 2, wc = w & 3;                 // this warp: tile rows 8 wr .. 8 wr + 7, tile columns 4 wc .. 4 wc + 3
    const int r = lane >> 2, q = lane & 3;             // accumulator fragment: row r, columns 2 q, 2 q + 1 of a tile
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double2 v = *reinterpret_cast<const double2*>(G + (long long)(8 * (8 * wr + i) + r) * ld + 8 * (4 * wc + j) + 2 * q);
            acc[i][j][0] = v.x; acc[i][j][1] = v.y;
        }
    PivotTrack pt;
    double* spi = sPi[w];
    double p0 = 0.0, p1 = 0.0;         // P^-1 of the current block step (LA: computed during the previous step)
    // (the inner eight block steps are unrolled so that the accumulator tile picked by kb is a compile-time register index:
    // a run-time subscript would push the whole matrix into local memory)
    for (int kbh = 0; kbh < 2; kbh++)
#pragma unroll
    for (int kbl = 0; kbl < 8; kbl++) {
        const int kb = 8 * kbh + kbl;
        const int buf = kbl & 1;
        double* sr = sR[buf];
        double* sc = sC[buf];
        // 1. publish the raw row / column panels
        if (wr == kbh) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i == kbl) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        *reinterpret_cast<double2*>(sr + r * LR + 8 * (4 * wc + j) + 2 * q) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
        }
        if (wc == (kb >> 2)) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (j == (kbl & 3)) {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        *reinterpret_cast<double2*>(sc + (8 * (8 * wr + i) + r) * LC + 2 * q) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
        }
        if (LA && kb < 15) {     // the owner of tile (kb+1, kb+1) publishes it
            const int t = kb + 1;
            if (wr == (t >> 3) && wc == (t >> 2)) {
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (i == ((kbl + 1) & 7) && j == ((kbl + 1) & 3))
                            *reinterpret_cast<double2*>(sD[buf] + r * LP + 2 * q) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
        __syncthreads();
        // 2. P^-1, by every warp on its own lanes (LA: only at the first step, afterwards it is already there)
        if (!LA || kb == 0) {
            const double2 v = *reinterpret_cast<const double2*>(sr + r * LR + 8 * kb + 2 * q); p0 = v.x; p1 = v.y;
            inv8x8_warp<FR>(p0, p1, r, q, pt);
        }
        *reinterpret_cast<double2*>(spi + r * LP + 2 * q) = make_double2(p0, p1);
        __syncwarp();
        // fragments of P^-1: as A operand (rows r, k' = q, q + 4) and, transposed, as B operand of step 3 (B[k'][n] = Pinv[n][k'])
        const double pa0 = spi[r * LP + q], pa1 = spi[r * LP + q + 4];
        // B operand of the column-kb tiles: B[k = 2 q + e][n = r] = Pinv[2 q + e][r]
        const double pb0 = spi[(2 * q) * LP + r], pb1 = spi[(2 * q + 1) * LP + r];
        // 3. xt[j] = (P^-1 R_j)^T in accumulator layout = B operand pair of the update; row-kb tiles get P^-1 R_j itself
        double xt[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int tj = 4 * wc + j;
            const double ra0 = sr[q * LR + 8 * tj + r], ra1 = sr[(q + 4) * LR + 8 * tj + r];     // A[m = r][k' = q (+4)] = R[k'][8 tj + r]
            double x0 = 0.0, x1 = 0.0;
            dmma884(x0, x1, ra0, pa0);     // B[k' = q][n = r] = Pinv[r][q]
            dmma884(x0, x1, ra1, pa1);
            xt[j][0] = x0; xt[j][1] = x1;
        }
        // LA: the next pivot block after this step's update, P' = D - C_t (P^-1 R_t), t = kb + 1
        double pn0 = 0.0, pn1 = 0.0;
        const bool la = LA && kb < 15;
        if (la) {
            const int t = kb + 1;
            const double ra0 = sr[q * LR + 8 * t + r], ra1 = sr[(q + 4) * LR + 8 * t + r];
            double x0 = 0.0, x1 = 0.0;
            dmma884(x0, x1, ra0, pa0);
            dmma884(x0, x1, ra1, pa1);
            const double2 cv = *reinterpret_cast<const double2*>(sc + (8 * t + r) * LC + 2 * q);
            const double2 dv = *reinterpret_cast<const double2*>(sD[buf] + r * LP + 2 * q);
            pn0 = dv.x; pn1 = dv.y;
            dmma884(pn0, pn1, -cv.x, x0);
            dmma884(pn0, pn1, -cv.y, x1);
        }
        // 4. trailing update and the special tile row / column (LA: one pivot of the next inverse per tile row)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (la) inv8x8_pivot<FR>(i, pn0, pn1, r, q, pt);
            const int ti = 8 * wr + i;
            const double2 cv = *reinterpret_cast<const double2*>(sc + (8 * ti + r) * LC + 2 * q);   // C[8 ti + r][2 q], [2 q + 1]
            const double a0 = -cv.x, a1 = -cv.y;
            const bool col_owner = wc == (kb >> 2);
            if (!(wr == kbh && i == kbl)) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (!(col_owner && j == (kbl & 3))) {
                        dmma884(acc[i][j][0], acc[i][j][1], a0, xt[j][0]);
                        dmma884(acc[i][j][0], acc[i][j][1], a1, xt[j][1]);
                    } else {           // A[I,K] <- -C_I P^-1
                        double y0 = 0.0, y1 = 0.0;
                        dmma884(y0, y1, a0, pb0);
                        dmma884(y0, y1, a1, pb1);
                        acc[i][j][0] = y0; acc[i][j][1] = y1;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int tj = 4 * wc + j;
                    if (!(col_owner && j == (kbl & 3))) {    // A[K,J] <- P^-1 R_J: A operand P^-1 (rows r, k' = q, q + 4), B[k'][n = r] = R[k'][8 tj + r]
                        double y0 = 0.0, y1 = 0.0;
                        dmma884(y0, y1, pa0, sr[q * LR + 8 * tj + r]);
                        dmma884(y0, y1, pa1, sr[(q + 4) * LR + 8 * tj + r]);
                        acc[i][j][0] = y0; acc[i][j][1] = y1;
                    } else { acc[i][j][0] = p0; acc[i][j][1] = p1; }
                }
            }
        }
        if (la) { p0 = pn0; p1 = pn1; }
        __syncwarp();      // spi is rewritten in the next step
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<double2*>(G + (long long)(8 * (8 * wr + i) + r) * ld + 8 * (4 * wc + j) + 2 * q) = make_double2(acc[i][j][0], acc[i][j][1]);
    if (threadIdx.x == 0 && min_pivot) pt.publish(min_pivot);
}

// ---- cluster-cooperative base case: one N x N block (N = 256: eight CTAs, N = 128: four) inverted by a thread-block cluster ------------
// The serial chain of the top tree levels (a handful of merges, X of order 4096 and 8192) is a sequence of 128-row base cases (48 us
// each on ONE SM, bounded by that SM's tensor pipe: 1900 cycles of trailing update per block of eight pivots) glued by small products
// that are pure launch latency.  Here a whole 256 x 256 block (two base cases + four products + their launches: 136 us) is one kernel:
// CTA c of the cluster keeps the column slab [32 c, 32 c + 32) in shared memory; per block of eight pivots the slab's owner inverts the
// 8 x 8 pivot block, pushes it and its 256 x 8 column panel into every CTA's shared memory (distributed shared memory, one cluster
// barrier per step, double-buffered), and every CTA applies the rank-8 update to its own 32 columns with DMMA.  Blocked Gauss-Jordan
// without pivoting, the same scalar pivots as the other base cases.
template <int N, int CL>
struct ClusterInvCfg {
    static constexpr int W = N / CL;                   // columns per CTA
    static constexpr int LDW = 40, LDC = 12, LDR = 36; // slab / column panel / row panel leading dimensions (bank-conflict-free fragment reads)
    static constexpr int DOUBLES = N * LDW + 2 * N * LDC + 2 * 64 + 8 * LDR + 64 + 4 * 8;
    static constexpr int SMEM_BYTES = DOUBLES * 8;
};

template <int N, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
invert_cluster_kernel(double* const* __restrict__ ptab, int nops, int op, long long off, long long off2, int ld, double* __restrict__ min_pivot)
{
    namespace cg = cooperative_groups;
    using Cfg = ClusterInvCfg<N, CL>;
    constexpr int W = Cfg::W, LDW = Cfg::LDW, LDC = Cfg::LDC, LDR = Cfg::LDR;
    static_assert(W == 32, "one CTA owns 32 columns");
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double smc[];
    double* slab = smc;                        // N x LDW: this CTA's 32 columns
    double* cpan = slab + N * LDW;             // 2 x (N x LDC): column panel of the current pivot block, pushed by its owner
    double* pinv = cpan + 2 * N * LDC;         // 2 x 64: inverse of the pivot block, pushed by its owner
    double* rbuf = pinv + 2 * 64;              // 8 x LDR: P^-1 A[kb, own columns]
    double* p8 = rbuf + 8 * LDR;               // 8 x 8 scratch of the pivot-block inversion
    double* stat = p8 + 64;                    // CL x 4 pivot statistics (gathered in CTA 0)
    const int cr = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nmat = off2 >= 0 ? 2 : 1;
    const int mat = blockIdx.x / CL;
    double* G = ptab[(long long)(mat / nmat) * nops + op] + ((mat % nmat) ? off2 : off) + cr * W;

    for (int e = tid; e < N * (W / 2); e += 256) {
        const int r = e / (W / 2), c2 = (e % (W / 2)) * 2;
        *reinterpret_cast<double2*>(slab + r * LDW + c2) = *reinterpret_cast<const double2*>(G + (long long)r * ld + c2);
    }
    __syncthreads();
    cluster.sync();   // every CTA of the cluster is resident before anybody writes into its shared memory

    PivotTrack pt;
    bool owned_any = false;
    const int g = lane >> 2, t = lane & 3;
    for (int kb = 0; kb < N / 8; kb++) {
        const int ow = (kb * 8) / W, lc = kb * 8 - ow * W, buf = kb & 1;
        if (cr == ow) {
            owned_any = true;
            // 1a. invert the 8 x 8 pivot block (warp 0: lane = row * 4 + column pair), scalar Gauss-Jordan through the p8 scratch
            if (warp == 0) {
                const int i = lane >> 2, j0 = (lane & 3) * 2;
                double a0 = slab[(kb * 8 + i) * LDW + lc + j0], a1 = slab[(kb * 8 + i) * LDW + lc + j0 + 1];
                p8[i * 8 + j0] = a0; p8[i * 8 + j0 + 1] = a1;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const double piv = p8[k * 8 + k];
                    const double rk0 = p8[k * 8 + j0], rk1 = p8[k * 8 + j0 + 1], ck = p8[i * 8 + k];
                    const double p = 1.0 / piv;
                    if (lane == 0) pt.see(piv);
                    if (i == k) { a0 = (j0 == k) ? p : a0 * p; a1 = (j0 + 1 == k) ? p : a1 * p; }
                    else {
                        a0 = (j0 == k) ? -ck * p : a0 - ck * (rk0 * p);
                        a1 = (j0 + 1 == k) ? -ck * p : a1 - ck * (rk1 * p);
                    }
                    __syncwarp();
                    p8[i * 8 + j0] = a0; p8[i * 8 + j0 + 1] = a1;
                    __syncwarp();
                }
            }
            __syncthreads();
            // 1b. push the column panel (rows of the other blocks: A[i, kb]) and P^-1 into every CTA of the cluster
            for (int e = tid; e < N * 4; e += 256) {
                const int r = e >> 2, j2 = (e & 3) * 2;
                const double2 v = *reinterpret_cast<const double2*>(slab + r * LDW + lc + j2);
#pragma unroll
                for (int d = 0; d < CL; d++)
                    *reinterpret_cast<double2*>(cluster.map_shared_rank(cpan, d) + buf * N * LDC + r * LDC + j2) = v;
            }
            if (tid < 64) {
                const double v = p8[tid];
#pragma unroll
                for (int d = 0; d < CL; d++) cluster.map_shared_rank(pinv, d)[buf * 64 + tid] = v;
            }
        }
        cluster.sync();
        const double* cp = cpan + buf * N * LDC;
        const double* pi = pinv + buf * 64;
        // 2a. row panel of the own columns: R = P^-1 A[kb, own]; in the owner's pivot columns P^-1 itself (new A[kb, kb] = P^-1, and
        //     new A[i, kb] = -A[i, kb] P^-1 comes out of the same update with a zero accumulator)
        {
            const int i = tid >> 5, c = tid & 31;
            double r = 0.0;
            if (cr == ow && c >= lc && c < lc + 8) r = pi[i * 8 + (c - lc)];
            else {
#pragma unroll
                for (int k = 0; k < 8; k++) r = fma(pi[i * 8 + k], slab[(kb * 8 + k) * LDW + c], r);
            }
            rbuf[i * LDR + c] = r;
        }
        __syncthreads();
        // 2b. rank-8 update of every other row block: A[rt, own] <- A[rt, own] - A[rt, kb] R   (DMMA m8n8k4, two k-steps)
        const int pct = (cr == ow) ? lc / 8 : -1;
        for (int rt = warp; rt < N / 8; rt += 8) {
            if (rt == kb) continue;
            const double a0 = -cp[(rt * 8 + g) * LDC + t], a1 = -cp[(rt * 8 + g) * LDC + t + 4];
#pragma unroll
            for (int ct = 0; ct < W / 8; ct++) {
                double2* cptr = reinterpret_cast<double2*>(slab + (rt * 8 + g) * LDW + ct * 8 + 2 * t);
                double2 acc = (ct == pct) ? make_double2(0.0, 0.0) : *cptr;
                dmma884(acc.x, acc.y, a0, rbuf[t * LDR + ct * 8 + g]);
                dmma884(acc.x, acc.y, a1, rbuf[(t + 4) * LDR + ct * 8 + g]);
                *cptr = acc;
            }
        }
        // row block kb itself becomes R (nobody reads rows kb of the slab in 2b)
        { const int i = tid >> 5, c = tid & 31; slab[(kb * 8 + i) * LDW + c] = rbuf[i * LDR + c]; }
        __syncthreads();
    }

    // pivot statistics of the whole block: gathered in CTA 0, published once
    if (tid == 0) {
        double* st0 = cluster.map_shared_rank(stat, 0) + cr * 4;
        st0[0] = owned_any ? pt.mn : 1e300; st0[1] = owned_any ? pt.mx : 0.0; st0[2] = (double)pt.neg;
    }
    for (int e = tid; e < N * (W / 2); e += 256) {
        const int r = e / (W / 2), c2 = (e % (W / 2)) * 2;
        *reinterpret_cast<double2*>(G + (long long)r * ld + c2) = *reinterpret_cast<const double2*>(slab + r * LDW + c2);
    }
    cluster.sync();   // (also: no CTA leaves while another may still write into its shared memory)
    if (cr == 0 && tid == 0 && min_pivot) {
        PivotTrack all;
        for (int d = 0; d < CL; d++) { all.mn = fmin(all.mn, stat[d * 4]); all.mx = fmax(all.mx, stat[d * 4 + 1]); all.neg += (unsigned)stat[d * 4 + 2]; }
        all.publish(min_pivot);
    }
}

template <int N, int CL>
static void launch_invert_cluster(double* const* ptab, int nops, int op, long long off, long long off2, int ld, int batch, double* min_pivot, cudaStream_t stream)
{
    using Cfg = ClusterInvCfg<N, CL>;
    auto kern = invert_cluster_kernel<N, CL>;
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared)) EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    kern<<<batch * CL, 256, Cfg::SMEM_BYTES, stream>>>(ptab, nops, op, off, off2, ld, min_pivot);
    EF_CUDA(cudaGetLastError());
}

void launch_invert_small(double* const* ptab, int nops, int op, long long off, long long off2, int ld, int N, int batch,
                         double* min_pivot, cudaStream_t stream)
{
    batch *= off2 >= 0 ? 2 : 1;   // CTAs: two blocks per entry
    if (N == 256 && ld % 2 == 0 && off % 2 == 0 && (off2 < 0 || off2 % 2 == 0)) {   // planned only where few blocks are inverted at a time (top tree levels)
        launch_invert_cluster<256, 8>(ptab, nops, op, off, off2, ld, batch, min_pivot, stream);
        return;
    }
    if (N > 128) throw Error{EF_ERR_BAD_SHAPE, "invert_small: N > 128"};
    switch (N) {
        case 32: invert_reg_kernel<2><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot); EF_CUDA(cudaGetLastError()); return;
        case 48: invert_reg_kernel<3><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot); EF_CUDA(cudaGetLastError()); return;
        case 64: invert_reg_kernel<4><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot); EF_CUDA(cudaGetLastError()); return;
        case 96: invert_reg_kernel<6><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot); EF_CUDA(cudaGetLastError()); return;
        case 128:
            // default: blocked Gauss-Jordan on the tensor pipe; efgpu_set_tuning(7, 1): the per-pivot register kernel of round 1.
            // Needs 16-byte aligned rows (even ld and offsets: always true for the merge matrices, whose blocks are multiples of 8)
            // (7, 2): the blocked kernel with a look-ahead of the next pivot block - measured (r2i): 8.62 against 8.12 ms of base
            // cases per step, the redundant work and register pressure cost more than the shortened chain gains; kept as a knob
            // tuning key 10 = 1: pivot reciprocals by hardware seed + two Newton steps instead of the IEEE division
            if (get_tuning(7) == 0 && get_tuning(10) == 1 && ld % 2 == 0 && off % 2 == 0 && (off2 < 0 || off2 % 2 == 0))
                invert_blk128_kernel<false, true><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot);
            else if (get_tuning(7) == 0 && ld % 2 == 0 && off % 2 == 0 && (off2 < 0 || off2 % 2 == 0))
                invert_blk128_kernel<false><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot);
            else if (get_tuning(7) == 2 && ld % 2 == 0 && off % 2 == 0 && (off2 < 0 || off2 % 2 == 0))
                invert_blk128_kernel<true><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot);
            else invert_reg_kernel<8><<<batch, 256, 0, stream>>>(ptab, nops, op, off, off2, ld, min_pivot);
            EF_CUDA(cudaGetLastError()); return;
        default: break;
    }
    int smem = (N * (N + 1) + N) * (int)sizeof(double);
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared))
        EF_CUDA(cudaFuncSetAttribute(invert_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 129 + 128) * 8));
    invert_small_kernel<<<batch, 256, smem, stream>>>(ptab, nops, op, off, off2, ld, N, min_pivot);
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
