// Peer-memory exchange of a row-partitioned (replicated) upper tree over NVLink / NVSwitch.
//
// Replaces what the reference does with MPI::broadcast of whole child patches between the ranks that share a node
// (src/Quadtree.hpp:464-507, src/QuadNode.hpp:191-199, FiniteVolumePatch.cpp:118-134) and what round 1 of this library did
// with one ncclAllGather per row slice issued from a host callback: every rank maps the shared operator arena of every
// other rank (CUDA IPC, one allocation per handle, identical layout on every rank because the merge plan is the same), the
// batched GEMM stores each finished tile into all arenas from its epilogue (gemm.cu: PeerSpan), and the only collective left is
// this flag barrier.  No NCCL, no host round trip, the transfer overlaps the tensor-core work tile by tile.
#include "common.cuh"
#include "kernels.cuh"

namespace efgpu {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

// One warp.  The epoch is a counter in the local flag page (word 64; nobody else writes it) that every barrier advances by one:
// all ranks issue the same sequence of barriers, so the counters agree without the host passing a number - which lets the launch
// sequence of a stage, barriers included, be captured into a CUDA graph and replayed.
// Lane r < nranks publishes the epoch into slot `me` of rank r's flag array (its own included), then waits until
// slot r of the local array has reached it.  Everything this GPU wrote before the kernel (peer stores of earlier
// kernels on the stream included) is ordered before the flag by the system-scope fence + release store; the acquire load
// orders the peers' data before whatever follows on this stream.  A peer that never arrives (crashed rank) must not hang the
// GPU: after `timeout_ns` the lane gives up and raises *err (checked by the host at the end of the stage).
__global__ void peer_barrier_kernel(PeerSpan ps, int me, unsigned long long timeout_ns, int* __restrict__ err)
{
    const int r = threadIdx.x;
    unsigned long long* local = reinterpret_cast<unsigned long long*>(ps.local_base);
    unsigned long long epoch = 0;
    if (r == 0) { epoch = local[64] + 1; local[64] = epoch; }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    if (r >= ps.n) return;
    __threadfence_system();
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(ps.local_base + ps.delta[r]);
    st_release_sys(remote + me, epoch);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(local + r) < epoch) {
        if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(err, 1 + r); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

void launch_peer_barrier(const PeerSpan& ps, int me, int* err, cudaStream_t s)
{
    static const unsigned long long timeout_ns = [] {
        const char* e = getenv("EFGPU_PEER_TIMEOUT_S"); const double v = e ? atof(e) : 20.0; return (unsigned long long)((v > 0.1 ? v : 20.0) * 1e9); }();
    peer_barrier_kernel<<<1, 32, 0, s>>>(ps, me, timeout_ns, err);
    EF_CUDA(cudaGetLastError());
}

// Copies `bytes` (a multiple of 16) at arena offset `off` of this rank into the same offset of every OTHER rank's arena:
// the broadcast half of an all-gather whose slices already lie in place.  Grid-stride over 16-byte words, the peers innermost
// so that consecutive stores of a thread go out on different NVLink destinations.
__global__ void __launch_bounds__(256) peer_scatter_kernel(PeerSpan ps, int me, size_t off, size_t n16)
{
    const int4* src = reinterpret_cast<const int4*>(ps.local_base + off);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const int4 v = src[i];
#pragma unroll 1
        for (int r = 0; r < ps.n; r++)
            if (r != me) reinterpret_cast<int4*>(ps.local_base + ps.delta[r] + off)[i] = v;
    }
}

void launch_peer_scatter(const PeerSpan& ps, int me, size_t off, size_t bytes, cudaStream_t s)
{
    if (bytes == 0 || ps.n <= 1) return;
    if ((off | bytes) & 15) throw Error{EF_ERR_BAD_ARG, "peer scatter: offset and size must be multiples of 16 bytes"};
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    peer_scatter_kernel<<<(unsigned)blocks, 256, 0, s>>>(ps, me, off, n16);
    EF_CUDA(cudaGetLastError());
}

// The same for a window of `rows` row segments of `row_bytes` bytes (a multiple of 16) every `pitch` bytes: the block columns of a
// column-partitioned DtN map.
__global__ void __launch_bounds__(256) peer_scatter2d_kernel(PeerSpan ps, int me, size_t off, size_t rows, size_t row16, size_t pitch)
{
    const size_t n16 = rows * row16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / row16, c = i - r * row16;
        const size_t byte = off + r * pitch + c * 16;
        const int4 v = *reinterpret_cast<const int4*>(ps.local_base + byte);
#pragma unroll 1
        for (int q = 0; q < ps.n; q++)
            if (q != me) *reinterpret_cast<int4*>(ps.local_base + ps.delta[q] + byte) = v;
    }
}

void launch_peer_scatter2d(const PeerSpan& ps, int me, size_t off, size_t rows, size_t row_bytes, size_t pitch, cudaStream_t s)
{
    if (rows == 0 || row_bytes == 0 || ps.n <= 1) return;
    if ((off | row_bytes | pitch) & 15) throw Error{EF_ERR_BAD_ARG, "peer scatter: offsets and sizes must be multiples of 16 bytes"};
    const size_t n16 = rows * (row_bytes / 16);
    size_t blocks = (n16 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    peer_scatter2d_kernel<<<(unsigned)blocks, 256, 0, s>>>(ps, me, off, rows, row_bytes / 16, pitch);
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
