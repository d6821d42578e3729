// patchcorr_kernels.cuh -- the `patch_correlation` vote of compute_weight_map (label/fusion.py:82-146) and the scalar image
// arithmetic its `correlation_function` applies (x + 1, abs(x) ...).
//
// The reference pads both (resampled) images, takes every cubic window as a numpy view and loops over them in Python calling
// scipy.stats.pearsonr -- minutes to hours per atlas.  Here one thread owns one voxel of the resampled grid and walks its
// window (clipped to the image, which is what the reference's padding mask amounts to) twice: means, then centred sums.
// Arithmetic is float64 like pearsonr's for Float32 data under the reference's pinned numpy 1.24 / scipy 1.9.3
// (dtype = type(1.0 + x[0] + y[0])), including its special cases: a constant patch gives NaN, which the reference then
// replaces by 0 (fusion.py:124), two samples give sign(dx) * sign(dy), and r is clipped to [-1, 1].
// Kernels only (no shared memory, barriers or runtime calls): also run by tests/emu under the serial host emulation.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace b200 {

// window w (voxels per axis); output voxel i covers [i - (w - 1) / 2, i + w / 2] along each axis (fusion.py:100-103)
__global__ void __launch_bounds__(128) patch_correlation_kernel(const float* __restrict__ t, const float* __restrict__ m, int nx, int ny, int nz, int wx, int wy,
                                                                int wz, double* __restrict__ out)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
        const int x0 = max(x - (wx - 1) / 2, 0), x1 = min(x + wx / 2, nx - 1);
        const int y0 = max(y - (wy - 1) / 2, 0), y1 = min(y + wy / 2, ny - 1);
        const int z0 = max(z - (wz - 1) / 2, 0), z1 = min(z + wz / 2, nz - 1);
        const long long cnt = (long long)(x1 - x0 + 1) * (y1 - y0 + 1) * (z1 - z0 + 1);
        const float tf = t[((size_t)z0 * ny + y0) * nx + x0], mf = m[((size_t)z0 * ny + y0) * nx + x0];
        double st = 0.0, sm = 0.0;
        bool t_const = true, m_const = true;
        for (int zz = z0; zz <= z1; ++zz)
            for (int yy = y0; yy <= y1; ++yy) {
                const size_t row = ((size_t)zz * ny + yy) * nx;
                for (int xx = x0; xx <= x1; ++xx) {
                    const float a = t[row + xx], b = m[row + xx];
                    st += (double)a;
                    sm += (double)b;
                    t_const = t_const && (a == tf);
                    m_const = m_const && (b == mf);
                }
            }
        double r;
        if (t_const || m_const) {
            r = 0.0;  // pearsonr returns NaN for a constant input; fusion.py:124 turns NaN into 0
        } else if (cnt == 2) {
            // pearsonr's n == 2 case: sign(x[1] - x[0]) * sign(y[1] - y[0]); neither input is constant here
            const size_t qa = ((size_t)z0 * ny + y0) * nx + x0, qb = ((size_t)z1 * ny + y1) * nx + x1;
            const double dt = (double)t[qb] - (double)t[qa], dm = (double)m[qb] - (double)m[qa];
            r = ((dt > 0.0) == (dm > 0.0)) ? 1.0 : -1.0;
        } else {
            const double mt = st / (double)cnt, mm = sm / (double)cnt;
            double stt = 0.0, smm = 0.0, stm = 0.0;
            for (int zz = z0; zz <= z1; ++zz)
                for (int yy = y0; yy <= y1; ++yy) {
                    const size_t row = ((size_t)zz * ny + yy) * nx;
                    for (int xx = x0; xx <= x1; ++xx) {
                        const double a = (double)t[row + xx] - mt, b = (double)m[row + xx] - mm;
                        stt += a * a;
                        smm += b * b;
                        stm += a * b;
                    }
                }
            r = stm / (sqrt(stt) * sqrt(smm));
            if (r != r) r = 0.0;  // a numerically constant patch (zero norm): NaN -> 0 like above
            r = fmax(fmin(r, 1.0), -1.0);
        }
        out[q] = r;
    }
}

// out = (take_abs ? |in| : in) * mul + add, rounded once per operation in the pixel type -- the image-with-constant operators
// (sitk.Abs, image * c, image + c) a `correlation_function` is made of.
template <typename T>
__global__ void __launch_bounds__(256) scale_shift_kernel(const T* __restrict__ in, size_t n, int take_abs, T mul, T add, T* __restrict__ out)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        T v = in[q];
        if (take_abs) v = v < (T)0 ? -v : v;
        v = v * mul;
        out[q] = v + add;
    }
}

}  // namespace b200
