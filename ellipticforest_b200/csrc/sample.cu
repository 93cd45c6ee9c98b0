// Sampling points and error norms on the device (the callers either side of the path).
//
// The reference samples the load and the coefficients point by point through std::function callbacks on the host
// (src/HPSAlgorithm.hpp:241-249; FiniteVolumeSolver.cpp:63-79) and its drivers compare the solution with the exact one in a
// serial loop over every cell (examples/elliptic-multiple/main.cpp:346-369).  At 1e7..1e9 cells these loops, not the
// stages, are the wall clock.  Here the library writes the sampling coordinates of every leaf into device arrays (the
// caller evaluates its functions on them with its own device code and hands the arrays back through
// efgpu_upwards_device / efgpu_set_leaf_variable_device) and reduces the three error norms of the drivers on the device.
// All kernels are bandwidth bound: 16 B written per point, 16 B read per cell.
#include "kernels.cuh"

namespace efgpu {

// x, y of leaf-major sampling points, cell index j + i*ny (i = x index).  which: 0 cell centres, 1..4 the W, E, S, N face
// midpoints of every cell.  The expressions follow the host mirror (ellipticforest_b200/hps.py, Mesh.leaf_cell_centres)
// operation by operation, with explicit round-to-nearest intrinsics so that nothing is contracted into an FMA: the
// coordinates are bit-identical to the host's.
__global__ void __launch_bounds__(256) leaf_points_kernel(const double* __restrict__ boxes, const int* __restrict__ leaf_nodes, int M, int which,
                                                          double* __restrict__ x, double* __restrict__ y, long long total)
{
    const int MM = M * M;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long leaf = t / MM;
        const int cell = (int)(t - leaf * MM);
        const int i = cell / M, j = cell - i * M;
        const double* b = boxes + 4 * (size_t)leaf_nodes[leaf];
        const double dx = __ddiv_rn(__dsub_rn(b[1], b[0]), (double)M), dy = __ddiv_rn(__dsub_rn(b[3], b[2]), (double)M);
        const double hx = __dmul_rn(dx, 0.5), hy = __dmul_rn(dy, 0.5);
        double xc = __dadd_rn(__dadd_rn(b[0], hx), __dmul_rn((double)i, dx));
        double yc = __dadd_rn(__dadd_rn(b[2], hy), __dmul_rn((double)j, dy));
        if (which == 1) xc = __dsub_rn(xc, hx);
        else if (which == 2) xc = __dadd_rn(xc, hx);
        else if (which == 3) yc = __dsub_rn(yc, hy);
        else if (which == 4) yc = __dadd_rn(yc, hy);
        if (x) x[t] = xc;
        if (y) y[t] = yc;
    }
}

void launch_leaf_points(const double* boxes, const int* leaf_nodes, int M, int which, double* x, double* y, int n_leaves, cudaStream_t s)
{
    const long long total = (long long)n_leaves * M * M;
    if (total == 0) return;
    const long long want = (total + 255) / 256;
    const int grid = (int)(want < 148LL * 16 ? want : 148LL * 16);   // grid-stride: at most 16 CTAs of 256 threads per SM
    leaf_points_kernel<<<grid, 256, 0, s>>>(boxes, leaf_nodes, M, which, x, y, total);
}

// Per leaf: dx dy sum |u - v|, dx dy sum (u - v)^2, max |u - v|  (examples/elliptic-multiple/main.cpp:355-366; the cell area
// is constant inside a leaf, so it multiplies the leaf's sums once).  One CTA per leaf, fixed reduction tree: the result does
// not depend on the launch configuration.
__global__ void __launch_bounds__(256) error_partial_kernel(const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ boxes,
                                                            const int* __restrict__ leaf_nodes, int M, double* __restrict__ part)
{
    __shared__ double sh[3][8];
    const int leaf = blockIdx.x, MM = M * M;
    const double* pu = u + (size_t)leaf * MM;
    const double* pv = v + (size_t)leaf * MM;
    double s1 = 0.0, s2 = 0.0, mx = 0.0;
    for (int c = threadIdx.x; c < MM; c += blockDim.x) {
        const double d = fabs(pu[c] - pv[c]);
        s1 += d; s2 += d * d; mx = fmax(mx, d);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; sh[2][warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; w++) { s1 += sh[0][w]; s2 += sh[1][w]; mx = fmax(mx, sh[2][w]); }
        const double* b = boxes + 4 * (size_t)leaf_nodes[leaf];
        const double area = ((b[1] - b[0]) / M) * ((b[3] - b[2]) / M);
        part[3 * (size_t)leaf] = area * s1;
        part[3 * (size_t)leaf + 1] = area * s2;
        part[3 * (size_t)leaf + 2] = mx;
    }
}

// out = { sum_1 / area, sqrt(sum_2 / area), max }  (main.cpp:369-371); single CTA, fixed order
__global__ void __launch_bounds__(1024) error_final_kernel(const double* __restrict__ part, int n_leaves, double area, double* __restrict__ out)
{
    __shared__ double sh[3][32];
    double s1 = 0.0, s2 = 0.0, mx = 0.0;
    for (int l = threadIdx.x; l < n_leaves; l += blockDim.x) {
        s1 += part[3 * (size_t)l]; s2 += part[3 * (size_t)l + 1]; mx = fmax(mx, part[3 * (size_t)l + 2]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; sh[2][warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; w++) { s1 += sh[0][w]; s2 += sh[1][w]; mx = fmax(mx, sh[2][w]); }
        out[0] = s1 / area; out[1] = sqrt(s2 / area); out[2] = mx;
    }
}

void launch_error_norms(const double* u, const double* v, const double* boxes, const int* leaf_nodes, int M, int n_leaves, double area,
                        double* part, double* out, cudaStream_t s)
{
    error_partial_kernel<<<n_leaves, 256, 0, s>>>(u, v, boxes, leaf_nodes, M, part);
    error_final_kernel<<<1, 1024, 0, s>>>(part, n_leaves, area, out);
}

// *out (pre-set to 0) <- max(0, max_i v[i]): tells efgpu_build whether any sampled lambda is positive (indefinite operator)
__global__ void __launch_bounds__(256) max_positive_kernel(const double* __restrict__ v, size_t n, double* __restrict__ out)
{
    double mx = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) mx = fmax(mx, v[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(mx));
}
void launch_max_positive(const double* v, size_t n, double* out, cudaStream_t s)
{
    EF_CUDA(cudaMemsetAsync(out, 0, sizeof(double), s));
    const size_t blocks = (n + 255) / 256;
    max_positive_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks)), 256, 0, s>>>(v, n, out);
    EF_CUDA(cudaGetLastError());
}

}  // namespace efgpu
