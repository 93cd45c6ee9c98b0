// TMA-staged variant of the FP64 tensor-core GEMM (A/B experiment for the merge products).
//
// north_star words the merge DGEMMs as "sm_100a FP64 tensor cores via TMA-staged tiles".  FP64 has no tcgen05 kind, so the math
// stays warp-level DMMA (mma.sync.m8n8k4.f64) exactly as in bgemm_kernel (gemm.cu); what changes here is the operand path:
//   * one elected thread issues cp.async.bulk.tensor.2d (SASS UTMALDG) per stage - a 128 x 16 box of A and 16 x 16 boxes of B,
//     hardware-swizzled (SWIZZLE_128B) instead of the BK+8 / BN+2 paddings - and signals an mbarrier with the byte count;
//   * the math warps (64 x 32 warp tiles; 128 x 128 CTA tile with eight warps, or 128 x 64 with four and two CTAs per SM like the
//     default LDGSTS kernel) wait on the stage's "full" barrier individually and release it through an "empty" barrier: no
//     __syncthreads and no LDGSTS / address arithmetic in the math loop; warps may drift apart by up to two k-tiles.
// Fragment addressing under the 128-byte swizzle (16-byte chunk c of 128-byte row r lives at chunk c ^ (r & 7)):
//   * B (k-major rows of 16 columns): lane (g = lane / 4, t = lane % 4) reads chunk g of rows 2t, 2t+1 (+ 8 kp) - the eight lanes
//     of a quarter warp hit eight different chunks: conflict free as it stands;
//   * A (rows of 16 k): lane reads chunk t + 4 kp of its row; rows g = 0, 1 of a quarter warp would collide (t ^ 0 and t ^ 1 cover
//     the same four chunks), so fragment row g of m-tile i is mapped to matrix row 8 ((g >> 1) + 4 (i >> 2)) + 4 (g & 1) + (i & 3)
//     of the warp's 64 rows - a permutation (the MMA does not care which eight rows form a tile as long as C goes back to the same
//     rows) whose swizzle key differs in bit 2 between g even and g odd.
// Same k order and slot assignment as bgemm_kernel, so the results are bit-identical (tests/test_gpu_parity.py).
#include "common.cuh"
#include <cuda.h>
#include <cstring>

namespace efgpu {

__device__ __forceinline__ void dmma884_t(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ double2 lds128(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <int BN_, int STAGES>
struct TmaCfg {
    static constexpr int BM = 128, BN = BN_, BK = 16;
    static constexpr int WARPS_N = BN / 32, CONSUMERS = 2 * WARPS_N;
    static constexpr int A_BYTES = BM * BK * 8;          // 128 rows of 128 bytes
    static constexpr int B_BOX = BK * 16 * 8;            // 16 k-rows of 16 columns
    static constexpr int B_BYTES = (BN / 16) * B_BOX;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + slack for the 1024-byte alignment of the swizzle atom
    static constexpr int NT = CONSUMERS * 32;
    static constexpr int DIST = STAGES - 2;              // prefetch distance in k-tiles (see the kernel)
    static constexpr int MIN_CTAS = BN == 64 ? 2 : 1;
};

template <int BN_, int STAGES>
__global__ void __launch_bounds__((TmaCfg<BN_, STAGES>::NT), (TmaCfg<BN_, STAGES>::MIN_CTAS))
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, double* __restrict__ C,
                 int m, int n, int k, int tiles_n, int tiles)
{
    using Cfg = TmaCfg<BN_, STAGES>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES];
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // shared-window address of stage 0

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = (int)(blockIdx.x % tiles);
    const int z = (int)(blockIdx.x / tiles);
    const int tm = tile / tiles_n, tn = tile % tiles_n;
    const int nk = k / Cfg::BK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], Cfg::CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // Producer = lane 0 of warp 0, between its own k-tiles (a ninth warp would put three warps on one scheduler and cap every thread
    // at 168 registers - the 128 accumulator registers plus fragments then spill).  At k-tile kt it refills the slot that k-tile kt - 2
    // used, so the "empty" wait concerns an iteration every warp has normally left long ago; STAGES - 2 tiles are in flight.
    const int arow = z * m + tm * Cfg::BM, bcol = tn * Cfg::BN, brow0 = z * k;
    auto issue = [&](int kl) {
        const int s = kl % STAGES;
        const unsigned st = base + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tma_load_2d(st, &mapA, kl * Cfg::BK, arow, &full_bar[s]);
#pragma unroll
        for (int j = 0; j < Cfg::BN / 16; j++)
            tma_load_2d(st + Cfg::A_BYTES + j * Cfg::B_BOX, &mapB, bcol + 16 * j, brow0 + kl * Cfg::BK, &full_bar[s]);
    };
    if (tid == 0)
        for (int kl = 0; kl < Cfg::DIST && kl < nk; kl++) issue(kl);

    // ---- consumers: 2 x WARPS_N warps, 64 x 32 each ----
    const int wm0 = (warp / Cfg::WARPS_N) * 64, wn0 = (warp % Cfg::WARPS_N) * 32;
    const int g = lane >> 2, t = lane & 3;
    constexpr int FM = 8, FN = 4;
    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // Swizzled addressing with everything lane-dependent hoisted: fragment row g of m-tile i is tile row
    // wm0 + 8 (g >> 1) + 4 (g & 1) + [32 (i >> 2) + (i & 3)], whose swizzle key is 4 (g & 1) + (i & 3); its chunk t + 4 kp therefore sits
    // at physical chunk (t ^ 4 (g & 1)) ^ (4 kp ^ (i & 3)) = c0 ^ x with x a compile-time constant in 0..7: eight lane constants.
    const unsigned a_lane = (unsigned)(wm0 + 8 * (g >> 1) + 4 * (g & 1)) * 128u;
    unsigned a_x[8];
#pragma unroll
    for (int x = 0; x < 8; x++) a_x[x] = a_lane + ((unsigned)((t ^ (4 * (g & 1))) ^ x) << 4);
    // B: rows 2t and 2t + 1 (+ 8 kp) of the 16-column box, chunk g: physical chunks g ^ 2t and g ^ (2t + 1)
    const unsigned b_lane0 = (unsigned)(Cfg::A_BYTES + (wn0 / 16) * Cfg::B_BOX + (2 * t) * 128 + ((g ^ (2 * t)) << 4));
    const unsigned b_lane1 = (unsigned)(Cfg::A_BYTES + (wn0 / 16) * Cfg::B_BOX + (2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4));

    for (int kt = 0; kt < nk; kt++) {
        const int s = kt % STAGES;
        if (tid == 0 && kt + Cfg::DIST < nk) {
            if (kt >= 2) mbar_wait(&empty_bar[(kt - 2) % STAGES], ((kt - 2) / STAGES) & 1);
            issue(kt + Cfg::DIST);
        }
        __syncwarp();
        mbar_wait(&full_bar[s], (kt / STAGES) & 1);
        const unsigned st = base + s * Cfg::STAGE_BYTES;
#pragma unroll
        for (int kp = 0; kp < Cfg::BK / 8; kp++) {
            double2 a[FM];
            double b0[FN], b1[FN];
#pragma unroll
            for (int i = 0; i < FM; i++)
                a[i] = lds128(st + a_x[(4 * kp) ^ (i & 3)] + (32 * (i >> 2) + (i & 3)) * 128);
#pragma unroll
            for (int jg = 0; jg < FN / 2; jg++) {
                const double2 v0 = lds128(st + b_lane0 + jg * Cfg::B_BOX + kp * 8 * 128);
                const double2 v1 = lds128(st + b_lane1 + jg * Cfg::B_BOX + kp * 8 * 128);
                b0[2 * jg] = v0.x; b0[2 * jg + 1] = v0.y; b1[2 * jg] = v1.x; b1[2 * jg + 1] = v1.y;
            }
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) dmma884_t(acc[i][j][0], acc[i][j][1], a[i].x, b0[j]);
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) dmma884_t(acc[i][j][0], acc[i][j][1], a[i].y, b1[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // epilogue: rows follow the permuted tile rows; columns as in bgemm_kernel's paired layout (four adjacent columns per lane)
    double* Cz = C + (size_t)z * m * n + (size_t)(tm * Cfg::BM + wm0 + 8 * (g >> 1) + 4 * (g & 1)) * n + tn * Cfg::BN + wn0 + 4 * t;
#pragma unroll
    for (int i = 0; i < FM; i++) {
#pragma unroll
        for (int jg = 0; jg < FN / 2; jg++) {
            double2* c = reinterpret_cast<double2*>(Cz + (size_t)(32 * (i >> 2) + (i & 3)) * n + jg * 16);
            c[0] = make_double2(acc[i][2 * jg][0], acc[i][2 * jg + 1][0]);
            c[1] = make_double2(acc[i][2 * jg][1], acc[i][2 * jg + 1][1]);
        }
    }
}

// ---- the descriptor-driven form: bgemm_kernel's semantics (two-term blocks, additive C0, signed transposed second destination,
// stores into peer arenas) on TMA-staged operands.  Operands are addressed through per-(batch entry, view) tensor maps - a view is
// (operand slot, origin inside the slot, leading dimension), a few dozen per batch - and block-relative coordinates (TmaBlock).
// 128 x 64 CTA tile, four warps of 64 x 32, four stages, two CTAs per SM.
__device__ __forceinline__ double flip_sign_t(double x, unsigned neg) {
    return __hiloint2double(__double2hiint(x) ^ (int)neg, __double2loint(x));
}

template <int STAGES>
__global__ void __launch_bounds__(128, 2)
bgemm_tma_kernel(double* const* __restrict__ ptab, int nops, const GemmBlock* __restrict__ blocks, const TmaBlock* __restrict__ tblocks,
                 const CUtensorMap* __restrict__ mapsA, const CUtensorMap* __restrict__ mapsB, int nviews,
                 int nblocks, int tiles_per_block, const PeerSpan ps)
{
    using Cfg = TmaCfg<64, STAGES>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
    constexpr int LDT_S = BM + 2;
    static_assert(STAGES * Cfg::STAGE_BYTES >= BN * LDT_S * 8, "the transposed tile reuses the pipeline stages");
    extern __shared__ unsigned char smem_raw[];
    __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES];
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;

    const long long bid = blockIdx.x;
    const int tile = (int)(bid % tiles_per_block);
    const int blk = (int)((bid / tiles_per_block) % nblocks);
    const long long z = bid / ((long long)tiles_per_block * nblocks);
    const GemmBlock& bd = blocks[blk];
    const TmaBlock& tb = tblocks[blk];
    const int tiles_n = bd.cols / BN;
    const int tm = tile / tiles_n, tn = tile % tiles_n;
    if (tm >= bd.rows / BM) return;
    double* const* ops = ptab + z * nops;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], Cfg::CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int nterms = bd.nterms;
    const int nk0 = bd.t[0].K / BK;
    const int nk = nk0 + (nterms > 1 ? bd.t[1].K / BK : 0);
    const int t1 = nterms > 1 ? 1 : 0;
    const unsigned neg0 = bd.t[0].neg, neg1 = bd.t[t1].neg;

    // producer state (thread 0 only uses it)
    const CUtensorMap* mA0 = mapsA + z * nviews + tb.t[0].a_view;
    const CUtensorMap* mB0 = mapsB + z * nviews + tb.t[0].b_view;
    const CUtensorMap* mA1 = mapsA + z * nviews + tb.t[t1].a_view;
    const CUtensorMap* mB1 = mapsB + z * nviews + tb.t[t1].b_view;
    const int ar0 = tb.t[0].a_row + tm * BM, ac0 = tb.t[0].a_col, br0 = tb.t[0].b_row, bc0 = tb.t[0].b_col + tn * BN;
    const int ar1 = tb.t[t1].a_row + tm * BM, ac1 = tb.t[t1].a_col, br1 = tb.t[t1].b_row, bc1 = tb.t[t1].b_col + tn * BN;
    auto issue = [&](int kl) {
        const int s = kl % STAGES;
        const unsigned st = base + s * Cfg::STAGE_BYTES;
        const bool second = kl >= nk0;
        const int kt = second ? kl - nk0 : kl;
        const CUtensorMap* mA = second ? mA1 : mA0;
        const CUtensorMap* mB = second ? mB1 : mB0;
        const int ar = second ? ar1 : ar0, ac = (second ? ac1 : ac0) + kt * BK;
        const int br = (second ? br1 : br0) + kt * BK, bc = second ? bc1 : bc0;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tma_load_2d(st, mA, ac, ar, &full_bar[s]);
#pragma unroll
        for (int j = 0; j < BN / 16; j++) tma_load_2d(st + Cfg::A_BYTES + j * Cfg::B_BOX, mB, bc + 16 * j, br, &full_bar[s]);
    };
    if (tid == 0)
        for (int kl = 0; kl < Cfg::DIST && kl < nk; kl++) issue(kl);

    const int wm0 = (warp >> 1) * 64, wn0 = (warp & 1) * 32;
    const int g = lane >> 2, t = lane & 3;
    constexpr int FM = 8, FN = 4;
    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int rl = wm0 + 8 * (g >> 1) + 4 * (g & 1);       // this lane's tile row for m-tile 0; m-tile i adds 32 (i >> 2) + (i & 3)
    const unsigned a_lane = (unsigned)rl * 128u;
    unsigned a_x[8];
#pragma unroll
    for (int x = 0; x < 8; x++) a_x[x] = a_lane + ((unsigned)((t ^ (4 * (g & 1))) ^ x) << 4);
    const unsigned b_lane0 = (unsigned)(Cfg::A_BYTES + (wn0 / 16) * Cfg::B_BOX + (2 * t) * 128 + ((g ^ (2 * t)) << 4));
    const unsigned b_lane1 = (unsigned)(Cfg::A_BYTES + (wn0 / 16) * Cfg::B_BOX + (2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4));

    const bool flip_mid = nterms > 1 && neg0 != neg1;
    for (int kt = 0; kt < nk; kt++) {
        const int s = kt % STAGES;
        if (tid == 0 && kt + Cfg::DIST < nk) {
            if (kt >= 2) mbar_wait(&empty_bar[(kt - 2) % STAGES], ((kt - 2) / STAGES) & 1);
            issue(kt + Cfg::DIST);
        }
        __syncwarp();
        mbar_wait(&full_bar[s], (kt / STAGES) & 1);
        if (flip_mid && kt == nk0) {
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
        }
        const unsigned st = base + s * Cfg::STAGE_BYTES;
#pragma unroll
        for (int kp = 0; kp < BK / 8; kp++) {
            double2 a[FM];
            double b0[FN], b1[FN];
#pragma unroll
            for (int i = 0; i < FM; i++)
                a[i] = lds128(st + a_x[(4 * kp) ^ (i & 3)] + (32 * (i >> 2) + (i & 3)) * 128);
#pragma unroll
            for (int jg = 0; jg < FN / 2; jg++) {
                const double2 v0 = lds128(st + b_lane0 + jg * Cfg::B_BOX + kp * 8 * 128);
                const double2 v1 = lds128(st + b_lane1 + jg * Cfg::B_BOX + kp * 8 * 128);
                b0[2 * jg] = v0.x; b0[2 * jg + 1] = v0.y; b1[2 * jg] = v1.x; b1[2 * jg + 1] = v1.y;
            }
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) dmma884_t(acc[i][j][0], acc[i][j][1], a[i].x, b0[j]);
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) dmma884_t(acc[i][j][0], acc[i][j][1], a[i].y, b1[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    if (neg1) {
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
    }

    // epilogue: C = acc (+ C0) at the permuted tile rows, four adjacent columns per lane and 16-column group; the transposed second
    // destination always leaves through shared memory as whole rows (the pipeline stages are free: every issued tile was consumed)
    const int col0 = tn * BN + wn0 + 4 * t;
    double* Cg = ops[bd.c_op] + bd.c_off + (long long)(tm * BM + rl) * bd.ldc + col0;
    const double* C0g = bd.c0_op >= 0 ? ops[bd.c0_op] + bd.c0_off + (long long)(tm * BM + rl) * bd.ldc0 + col0 : nullptr;
    const bool staged_t = bd.ct_op1 != 0;
    const unsigned ctn = bd.ct_neg;
    double* smem_d = reinterpret_cast<double*>(smem_raw + (base - smem_u32(smem_raw)));
    double* Cts = smem_d + (wn0 + 4 * t) * LDT_S + rl;
    if (staged_t) __syncthreads();
#pragma unroll
    for (int i = 0; i < FM; i++) {
        const int ro = 32 * (i >> 2) + (i & 3);
#pragma unroll
        for (int jg = 0; jg < FN / 2; jg++) {
            double2 lo = make_double2(acc[i][2 * jg][0], acc[i][2 * jg + 1][0]);
            double2 hi = make_double2(acc[i][2 * jg][1], acc[i][2 * jg + 1][1]);
            if (C0g) {
                const double2* c0 = reinterpret_cast<const double2*>(C0g + (long long)ro * bd.ldc0 + jg * 16);
                const double2 c0l = c0[0], c0h = c0[1];
                lo.x += c0l.x; lo.y += c0l.y; hi.x += c0h.x; hi.y += c0h.y;
            }
            double2* c = reinterpret_cast<double2*>(Cg + (long long)ro * bd.ldc + jg * 16);
            c[0] = lo; c[1] = hi;
            for (int r = 0; r < ps.n; r++)
                if (r != ps.me) {
                    double2* cp = reinterpret_cast<double2*>(reinterpret_cast<char*>(c) + ps.delta[r]);
                    cp[0] = lo; cp[1] = hi;
                }
            if (staged_t) {
                double* tt = Cts + (jg * 16) * LDT_S + ro;
                tt[0] = flip_sign_t(lo.x, ctn); tt[LDT_S] = flip_sign_t(lo.y, ctn);
                tt[2 * LDT_S] = flip_sign_t(hi.x, ctn); tt[3 * LDT_S] = flip_sign_t(hi.y, ctn);
            }
        }
    }
    if (staged_t) {
        __syncthreads();
        double* Ctb = ops[bd.ct_op1 - 1] + bd.ct_off + (long long)(tn * BN) * bd.ldct + tm * BM;
        constexpr int V = BM / 2;
        for (int idx = tid; idx < BN * V; idx += 128) {
            const int c = idx / V, v2 = (idx % V) * 2;
            const double2 val = *reinterpret_cast<const double2*>(smem_d + c * LDT_S + v2);
            double2* dst = reinterpret_cast<double2*>(Ctb + (long long)c * bd.ldct + v2);
            *dst = val;
            for (int r = 0; r < ps.n; r++)
                if (r != ps.me) *reinterpret_cast<double2*>(reinterpret_cast<char*>(dst) + ps.delta[r]) = val;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            throw Error{EF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver"};
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// row-major rows x cols matrix of doubles (leading dimension ld), boxes of box_rows x 16 columns, 128-byte swizzle
static CUtensorMap make_map(const double* ptr, unsigned long long rows, unsigned long long cols, unsigned long long ld, unsigned box_rows)
{
    CUtensorMap mp;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * sizeof(double)};
    const cuuint32_t box[2] = {16, box_rows};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = encode_fn()(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error{EF_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"};
    return mp;
}

// tensor maps of one operand view (base pointer of the matrix, leading dimension): boxes of 128 x 16 (A operand) and 16 x 16 (B operand)
void encode_operand_maps(const double* base, unsigned long long ld, void* mapA_out, void* mapB_out)
{
    static_assert(sizeof(CUtensorMap) == TMA_MAP_BYTES, "tensor map size");
    const unsigned long long rows = 1ull << 30;   // views are windows into larger buffers: the row bound is never the limit
    const CUtensorMap a = make_map(base, rows, ld, ld, 128), b = make_map(base, rows, ld, ld, 16);
    std::memcpy(mapA_out, &a, sizeof(a)); std::memcpy(mapB_out, &b, sizeof(b));
}

void launch_bgemm_tma(double* const* ptab, int nops, const GemmBlock* d_blocks, const GemmBlock* h_blocks, int nblocks, int batch,
                      cudaStream_t stream, const PeerSpan& ps, const TmaArgs& tma)
{
    using Cfg = TmaCfg<64, 4>;
    auto kern = bgemm_tma_kernel<4>;
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared)) EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int max_tiles = 0;
    for (int b = 0; b < nblocks; b++) { const int v = (h_blocks[b].rows / Cfg::BM) * (h_blocks[b].cols / Cfg::BN); if (v > max_tiles) max_tiles = v; }
    const long long grid = (long long)max_tiles * nblocks * batch;
    if (grid <= 0) return;
    if (grid > 2147483647LL) throw Error{EF_ERR_BAD_SHAPE, "bgemm grid too large"};
    kern<<<(unsigned)grid, 128, Cfg::SMEM_BYTES, stream>>>(ptab, nops, d_blocks, tma.d_tblocks, static_cast<const CUtensorMap*>(tma.mapsA),
                                                           static_cast<const CUtensorMap*>(tma.mapsB), tma.nviews, nblocks, max_tiles, ps);
    EF_CUDA(cudaGetLastError());
}

template <int BN_, int STAGES>
static void launch_tma(const CUtensorMap& ma, const CUtensorMap& mb, double* C, int m, int n, int k, int batch, cudaStream_t s)
{
    using Cfg = TmaCfg<BN_, STAGES>;
    auto kern = dgemm_tma_kernel<BN_, STAGES>;
    static unsigned long long prepared = 0;
    if (first_use_on_device(prepared)) EF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int tiles_n = n / Cfg::BN, tiles = (m / Cfg::BM) * tiles_n;
    kern<<<(unsigned)((long long)tiles * batch), Cfg::NT, Cfg::SMEM_BYTES, s>>>(ma, mb, C, m, n, k, tiles_n, tiles);
    EF_CUDA(cudaGetLastError());
}

// C[b] = A[b] (m x k) B[b] (k x n), densely packed row-major batches; m % 128 == n % 128 == k % 16 == 0.
// variant: 4 (default) / 6 / 3: 128 x 128 CTA tile, eight consumer warps, that many stages, one CTA per SM;
//          64: 128 x 64 CTA tile, four consumer warps, four stages, two CTAs per SM (the register budget of 5 warps x 2 CTAs is 168)
void launch_dgemm_tma(const double* A, const double* B, double* C, int m, int n, int k, int batch, int stages, cudaStream_t s)
{
    if (m % 128 || n % 128 || k % 16 || batch < 1) throw Error{EF_ERR_BAD_SHAPE, "dgemm_tma: m % 128, n % 128, k % 16 must be 0"};
    const CUtensorMap ma = make_map(A, (unsigned long long)batch * m, k, k, 128);
    const CUtensorMap mb = make_map(B, (unsigned long long)batch * k, n, n, 16);
    switch (stages) {
        case 3: launch_tma<128, 3>(ma, mb, C, m, n, k, batch, s); break;
        case 6: launch_tma<128, 6>(ma, mb, C, m, n, k, batch, s); break;
        case 64: launch_tma<64, 4>(ma, mb, C, m, n, k, batch, s); break;
        default: launch_tma<128, 4>(ma, mb, C, m, n, k, batch, s); break;
    }
}

}  // namespace efgpu
