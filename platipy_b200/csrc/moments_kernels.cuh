// moments_kernels.cuh -- itk::ImageMomentsCalculator as sitk.CenteredTransformInitializer(..., MOMENTS) uses it
// (reference linear.py:40-43, alignment_registration(moments=True)): total mass and intensity-weighted first moments in
// PHYSICAL coordinates, [sum v, sum v x, sum v y, sum v z]; the centre of gravity is their ratio.
// Per-voxel code in plain C++ (also run by tests/emu); only the block reduction is CUDA-specific.
#pragma once
#include <cstddef>
#include <cstdint>

namespace b200 {

struct MomentsGeom {
    int nx, ny, nz;
    double origin[3];
    double i2p[9];  // Direction * diag(Spacing)
};
constexpr int MOMENTS_NV = 4;

__global__ void __launch_bounds__(256) image_moments_kernel(const float* __restrict__ img, const __grid_constant__ MomentsGeom g, double* __restrict__ partials)
{
    double acc[MOMENTS_NV] = { 0.0, 0.0, 0.0, 0.0 };
    const size_t n = (size_t)g.nx * g.ny * g.nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % g.nx), j = (int)((q / g.nx) % g.ny), k = (int)(q / ((size_t)g.nx * g.ny));
        const double v = (double)img[q];
        acc[0] += v;
#pragma unroll
        for (int r = 0; r < 3; ++r) {  // ImageBase::TransformIndexToPhysicalPoint
            double sum = 0.0;
            sum += g.i2p[r * 3 + 0] * (double)i;
            sum += g.i2p[r * 3 + 1] * (double)j;
            sum += g.i2p[r * 3 + 2] * (double)k;
            acc[1 + r] += (sum + g.origin[r]) * v;
        }
    }
#ifdef B200_HOST_EMU
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int v = 0; v < MOMENTS_NV; ++v) partials[tid * MOMENTS_NV + v] = acc[v];
#else
    __shared__ double sh[MOMENTS_NV][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < MOMENTS_NV; ++v) {
        const double t = warp_sum(acc[v]);
        if (lane == 0) sh[v][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < MOMENTS_NV) {
        double t = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) t += sh[threadIdx.x][w8];
        partials[(size_t)blockIdx.x * MOMENTS_NV + threadIdx.x] = t;
    }
#endif
}

__global__ void image_moments_final_kernel(const double* __restrict__ partials, int nb, double* __restrict__ out)
{
    const int v = threadIdx.x;
    if (v >= MOMENTS_NV) return;
    double t = 0.0;
    for (int q = 0; q < nb; ++q) t += partials[(size_t)q * MOMENTS_NV + v];
    out[v] = t;
}

}  // namespace b200
