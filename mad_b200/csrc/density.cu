// Atoms -> simulated density map (PDB.structure_to_density, mad/PDB.py:131-163 and
// interpolate_to_grid_massweighted :215-292), the step in front of the hot path.
//   1. mass-weighted trilinear splat of the atoms into a float64 grid (margin 2 + pad voxels)
//   2. division by the grid maximum (float64)
//   3. "full" convolution with the Gaussian exp(-r^2 / 2 sigma^2) truncated at 3 sigma and normalised to
//      sum 1 (scipy.signal.convolve: zero extension, output grows by 2r per axis) -- done as three 1-D
//      float64 passes (the kernel is a product; the reference's FFT / direct sum differs by ~1e-16)
//   4. float32 cast; max normalisation and isovalue cut are mad_grid_max / mad_threshold_normalise.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
splat_kernel(const double* __restrict__ xyz, const double* __restrict__ mass, int n, double minx, double miny, double minz,
             double voxelsp, int margin, int py, int pz, double* __restrict__ grid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double gx = margin + (xyz[3 * i] - minx) / voxelsp;          // mad/PDB.py:262-264
    const double gy = margin + (xyz[3 * i + 1] - miny) / voxelsp;
    const double gz = margin + (xyz[3 * i + 2] - minz) / voxelsp;
    const double fx = floor(gx), fy = floor(gy), fz = floor(gz);
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const double a = (fx + 1.0) - gx, b = (fy + 1.0) - gy, c = (fz + 1.0) - gz;   // :275-277
    const double m = mass[i];
    auto at = [&](int x, int y, int z) { return grid + ((long long)x * py + y) * pz + z; };
    atomicAdd(at(x0, y0, z0), m * a * b * c);                           // :278-285 (same products, atomics reorder the sums)
    atomicAdd(at(x0, y0, z0 + 1), m * a * b * (1 - c));
    atomicAdd(at(x0, y0 + 1, z0), m * a * (1 - b) * c);
    atomicAdd(at(x0 + 1, y0, z0), m * (1 - a) * b * c);
    atomicAdd(at(x0, y0 + 1, z0 + 1), m * a * (1 - b) * (1 - c));
    atomicAdd(at(x0 + 1, y0 + 1, z0), m * (1 - a) * (1 - b) * c);
    atomicAdd(at(x0 + 1, y0, z0 + 1), m * (1 - a) * b * (1 - c));
    atomicAdd(at(x0 + 1, y0 + 1, z0 + 1), m * (1 - a) * (1 - b) * (1 - c));
}

__global__ void __launch_bounds__(256)
max_f64_kernel(const double* __restrict__ g, long long n, unsigned long long* __restrict__ out_bits) {
    double m = 0.0;                                                     // the grid is non-negative
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, g[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));   // order-preserving for m >= 0
}

__global__ void __launch_bounds__(256)
divide_f64_kernel(double* __restrict__ g, long long n, const unsigned long long* __restrict__ max_bits) {
    const double m = __longlong_as_double((long long)*max_bits);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        g[i] = g[i] / m;
}

// "full" 1-D convolution along one axis with zero extension: out extent = n + 2r.
// in: [outer][n][inner] -> out: [outer][n + 2r][inner];  out[i] = sum_j w[j] * in[i - j]  (j = 0..2r, symmetric w)
template <typename TOut>
__global__ void __launch_bounds__(256)
conv_full_kernel(const double* __restrict__ in, long long outer, int n, long long inner, int r, const double* __restrict__ w,
                 TOut* __restrict__ out) {
    const long long total = outer * (long long)(n + 2 * r) * inner;
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total; gidx += (long long)gridDim.x * blockDim.x) {
        const long long j = gidx % inner;
        const long long t = gidx / inner;
        const int i = (int)(t % (n + 2 * r));
        const long long o = t / (n + 2 * r);
        const double* line = in + o * (long long)n * inner + j;
        double acc = 0.0;
        for (int q = 0; q <= 2 * r; ++q) {
            const int src = i - q;
            if (src >= 0 && src < n) acc = fma(w[q], line[(long long)src * inner], acc);
        }
        out[gidx] = (TOut)acc;
    }
}

}  // namespace

extern "C" int mad_density_splat(const double* xyz, const double* mass, int n_atoms, const double* min_host, double voxelsp,
                                 int margin, int px, int py, int pz, double* grid, void* stream) {
    MAD_CHECK_ARG(xyz && mass && min_host && grid && n_atoms > 0 && voxelsp > 0 && margin >= 1 && px > 0 && py > 0 && pz > 0);
    cudaStream_t st = (cudaStream_t)stream;
    MAD_CUDA(cudaMemsetAsync(grid, 0, sizeof(double) * (size_t)px * py * pz, st));
    MAD_PROF("splat_kernel", st);
    splat_kernel<<<(unsigned)mad_ceil_div(n_atoms, 256), 256, 0, st>>>(xyz, mass, n_atoms, min_host[0], min_host[1], min_host[2],
                                                                       voxelsp, margin, py, pz, grid);
    MAD_LAUNCH_OK();
    return MAD_OK;
}

extern "C" int mad_normalise_f64(double* grid, long long n, void* scratch8, void* stream) {
    MAD_CHECK_ARG(grid && scratch8 && n > 0);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* bits = reinterpret_cast<unsigned long long*>(scratch8);
    MAD_CUDA(cudaMemsetAsync(bits, 0, 8, st));
    const int blocks = (int)std::max<long long>(1, std::min<long long>(mad_ceil_div(n, 1024), (long long)mad_sm_count() * 8));
    {
        MAD_PROF("max_f64_kernel", st);
        max_f64_kernel<<<blocks, 256, 0, st>>>(grid, n, bits);
        MAD_LAUNCH_OK();
    }
    MAD_PROF("divide_f64_kernel", st);
    divide_f64_kernel<<<blocks, 256, 0, st>>>(grid, n, bits);
    MAD_LAUNCH_OK();
    return MAD_OK;
}

extern "C" int mad_conv_full_f64(const double* in, long long outer, int n, long long inner, const double* w_dev, int radius,
                                 void* out, int out_is_f32, void* stream) {
    MAD_CHECK_ARG(in && w_dev && out && outer > 0 && n > 0 && inner > 0 && radius >= 0);
    const long long total = outer * (long long)(n + 2 * radius) * inner;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(mad_ceil_div(total, 256), (long long)mad_sm_count() * 16));
    MAD_PROF("conv_full_kernel", stream);
    if (out_is_f32)
        conv_full_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, outer, n, inner, radius, w_dev, reinterpret_cast<float*>(out));
    else
        conv_full_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, outer, n, inner, radius, w_dev, reinterpret_cast<double*>(out));
    MAD_LAUNCH_OK();
    return MAD_OK;
}
