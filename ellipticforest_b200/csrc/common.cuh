// Shared declarations for the efgpu CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace efgpu {

// ---- status codes (mirrored in include/efgpu.h) ---------------------------------------------
enum : int { EF_OK = 0, EF_ERR_CUDA = 1, EF_ERR_BAD_ARG = 2, EF_ERR_BAD_SHAPE = 3, EF_ERR_OOM = 4,
             EF_ERR_SINGULAR = 5, EF_ERR_STATE = 6, EF_ERR_UNSUPPORTED = 7 };

struct Error { int code; std::string msg; };
#define EF_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw ::efgpu::Error{e__ == cudaErrorMemoryAllocation ? ::efgpu::EF_ERR_OOM : ::efgpu::EF_ERR_CUDA, \
                                 std::string(#call) + ": " + cudaGetErrorString(e__)};             \
    } while (0)

// Function attributes (opt-in shared memory above 48 KB) and occupancy figures are per DEVICE: a launch site keeps a bit mask of
// the devices it has prepared.  True the first time the site is reached with the current device.
inline bool first_use_on_device(unsigned long long& seen)
{
    int dev = 0;
    EF_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}

// ---- descriptor-driven batched DGEMM (gemm.cu) ----------------------------------------------
// One launch computes, for every batch entry z and every block descriptor d:
//     C_d = [C0_d] + sum_t  sign_t * A_{d,t} (rows x K_t) * B_{d,t} (K_t x cols)
// Operands are addressed as ptab[z * nops + op] + offset, row-major with their own leading
// dimension, so a descriptor can point straight into the children's DtN matrices (the block
// gathers, negations and the WESN block permutation of the reference's merge are folded into
// operand/result addressing and never materialised).
constexpr int GEMM_MAX_TERMS = 2;
struct GemmTerm {
    int a_op, b_op;
    int lda, ldb;
    long long a_off, b_off;   // element offsets of the (0,0) entry of the block
    int K;
    unsigned neg;             // 0 or 0x80000000: flips the sign of A on the fly
};
struct GemmBlock {
    int c_op, c0_op;          // c0_op < 0: no additive term
    int ldc, ldc0;
    long long c_off, c0_off;
    int rows, cols;
    int nterms;
    // optional second destination: the (signed) TRANSPOSE of the result block, stored by the same epilogue - element (r, c) of the
    // block also goes to ptab[ct_op1 - 1] + ct_off + c * ldct + r.  0 = none.  Replaces the separate block-transpose launches of
    // the symmetric plans (lower blocks of X^-1, mirrored blocks of the DtN map).
    int ct_op1;
    int ldct;
    unsigned ct_neg;          // 0 or 0x80000000
    long long ct_off;
    GemmTerm t[GEMM_MAX_TERMS];
};
// Peer arenas of a row-partitioned tree (peer.cu): rank r's copy of the shared operator arena is mapped at
// local_base + delta[r] (delta[me] = 0).  A GEMM launched with a span stores every finished tile at the same arena offset on
// every rank (the all-gather of the row slices, fused into the epilogue); n = 0: plain local stores.
constexpr int PEER_MAX = 8;
struct PeerSpan {
    int n = 0, me = 0;
    int stage_t = 0;          // set by launch_bgemm: transposed second destinations leave through shared memory (whole rows)
    char* local_base = nullptr;
    long long delta[PEER_MAX] = {0, 0, 0, 0, 0, 0, 0, 0};
};
void launch_peer_barrier(const PeerSpan& ps, int me, int* err, cudaStream_t s);
void launch_peer_scatter(const PeerSpan& ps, int me, size_t off, size_t bytes, cudaStream_t s);
void launch_peer_scatter2d(const PeerSpan& ps, int me, size_t off, size_t rows, size_t row_bytes, size_t pitch, cudaStream_t s);

// TMA-staged operands (gemm_tma.cu): a view is one matrix an operand block lies in - (operand slot, origin inside the slot, leading
// dimension); every (batch entry, view) has a tensor map for its use as A operand (128 x 16 boxes) and as B operand (16 x 16 boxes);
// a block's terms name their views and the coordinates of their (0, 0) element.
constexpr int TMA_MAP_BYTES = 128;
struct TmaTerm { int a_view, b_view, a_row, a_col, b_row, b_col; };
struct TmaBlock { TmaTerm t[GEMM_MAX_TERMS]; };
struct TmaArgs {
    const TmaBlock* d_tblocks = nullptr;   // parallel to the GemmBlock array of the launch
    const void* mapsA = nullptr;           // CUtensorMap[batch][nviews]
    const void* mapsB = nullptr;
    int nviews = 0;
};
void encode_operand_maps(const double* base, unsigned long long ld, void* mapA_out, void* mapB_out);
void launch_bgemm_tma(double* const* ptab, int nops, const GemmBlock* d_blocks, const GemmBlock* h_blocks, int nblocks, int batch,
                      cudaStream_t stream, const PeerSpan& ps, const TmaArgs& tma);

// Launches the kernel on `stream`.  rows/cols/K of every block must be multiples of the chosen
// tile; the tile configuration is picked from the block shape and the amount of parallelism.
// peers != nullptr: results are also stored into the peer arenas (every C of the launch must lie inside the local arena).
void launch_bgemm(double* const* ptab, int nops, const GemmBlock* d_blocks, const GemmBlock* h_blocks,
                  int nblocks, int batch, cudaStream_t stream, int force_tile = 0, const PeerSpan* peers = nullptr, const TmaArgs* tma = nullptr);

// gemm_tma.cu: C[b] = A[b] (m x k) B[b] (k x n), densely packed row-major batches, operands staged by TMA (A/B experiment)
void launch_dgemm_tma(const double* A, const double* B, double* C, int m, int n, int k, int batch, int stages, cudaStream_t s);

// Batched out-of-place block transposes  dst (cols x rows) = +-src (rows x cols)^T, addressed like the GEMM operands.
// Used where the merge matrices are symmetric (uniform, self-adjoint subtrees): the lower blocks of X^-1 and the
// mirrored blocks of the DtN map T are copies of computed blocks instead of further GEMMs.
struct TransOp {
    int src_op, dst_op;
    int lds, ldd;
    long long src_off, dst_off;
    int rows, cols;           // shape of the source block
    unsigned neg;             // 0 or 0x80000000
    int pad_;
};
void launch_btranspose(double* const* ptab, int nops, const TransOp* d_ops, const TransOp* h_ops, int nt, int batch,
                       cudaStream_t stream);

// In-place inverse of `batch` small dense N x N matrices (N <= 128) held at
// ptab[z*nops+op] + off with leading dimension ld; Gauss-Jordan in shared memory, no pivoting
// (the merge matrices are SPD / diagonally dominant).  Pivot statistics are folded
// into the 4-double tracker `min_pivot` ([0] min |pivot|, [1] max |pivot|, [2] smallest per-block min / max ratio,
// [3] number of negative pivots as a 64-bit count) for singularity / conditioning reports.
// off2 >= 0: a second block per entry in the same launch.
void launch_invert_small(double* const* ptab, int nops, int op, long long off, long long off2, int ld, int N, int batch,
                         double* min_pivot, cudaStream_t stream);
void launch_pivot_tracker_reset(double* tracker, cudaStream_t stream);

#ifdef __CUDACC__
// ---- mbarrier / bulk-copy (cp.async.bulk global -> shared) helpers shared by the streaming kernels ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EF_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra EF_MBAR_DONE;\n"
        "bra EF_MBAR_WAIT;\n"
        "EF_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

}  // namespace efgpu
