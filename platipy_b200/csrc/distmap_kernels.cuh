// distmap_kernels.cuh -- kernels of the distance-map / contour / binary-morphology utilities the rows after the hot
// path need:
//   sitk.SignedMaurerDistanceMap   registration/utils.py:289-294, label/projection.py:22-31,80-82
//   sitk.LabelContour              label/projection.py:33,85
//   sitk.BinaryDilate / BinaryErode registration/utils.py:331, generation/dvf.py:269-287
//   sitk.Mask, image / constant    registration/utils.py:337-342, generation/dvf.py:66,121,200
//   the field of generate_field_radial_bend   generation/dvf.py:362-394
//
// Kernels only: no shared memory, no barriers, no runtime calls, grid-stride loops -- so the same source also runs
// under the serial host emulation of tests/emu/ (the build container has no GPU; the emulation checks indexing and
// arithmetic against the oracle there, the -m gpu tests check the real launches).  Launch wrappers: distmap.cuh.
//
// ITK's SignedMaurerDistanceMapImageFilter works in its OUTPUT pixel type, Float32 for SimpleITK: squared distances,
// line coordinates and the parabola-intersection test are all single precision, and so are they here (-fmad=false),
// which makes the result bit-identical to the CPU restatement rather than merely close to the exact transform.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace b200 {

// BinaryThreshold(!= background) + BinaryContour(FullyConnected = true) + the initial image of the Voronoi passes:
// 0 on object voxels that have a background voxel among their 26 neighbours inside the image, FLT_MAX elsewhere.
__global__ void __launch_bounds__(256) maurer_init_kernel(const uint8_t* __restrict__ mask, int nx, int ny, int nz, float* __restrict__ out)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        float v = FLT_MAX;
        if (mask[q]) {
            const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
            bool border = false;
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = z + dz;
                if (zz < 0 || zz >= nz) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= ny) continue;
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = x + dx;
                        if (xx < 0 || xx >= nx) continue;
                        if (!mask[((size_t)zz * ny + yy) * nx + xx]) border = true;
                    }
                }
            }
            if (border) v = 0.0f;
        }
        out[q] = v;
    }
}

// SignedMaurerDistanceMapImageFilter::Remove, in the output pixel type
__device__ __forceinline__ bool maurer_remove(float d1, float d2, float df, float x1, float x2, float xf)
{
    const float a = x2 - x1;
    const float b = xf - x2;
    const float c = xf - x1;
    const float value = c * fabsf(d2) - b * fabsf(d1) - a * fabsf(df) - a * b * c;
    return value > 0.0f;
}

// One Voronoi pass (SignedMaurerDistanceMapImageFilter::Voronoi) along `axis`: one thread per image line.  The partial
// Voronoi diagram of the line is a stack of (squared distance so far g, position h) kept in two scratch volumes, entry k
// of line L at [k * nlines + L] (coalesced across the threads of a warp).  h is stored as the voxel index; the
// coordinate is (float) index * (float) spacing as in the filter.  Values carry the inside / outside sign between the
// passes exactly as ITK stores them (|.| is taken when they are read back).  LAST: the pass over the last dimension
// also applies the closing loop of ThreadedGenerateData (sqrt unless squared distances are asked for).
template <bool LAST>
__global__ void __launch_bounds__(128) maurer_voronoi_kernel(float* __restrict__ out, const uint8_t* __restrict__ mask, int nx, int ny, int nz, int axis,
                                                             float spf, int inside_is_positive, int squared, float* __restrict__ g, int* __restrict__ h)
{
    const int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const size_t nlines = axis == 0 ? (size_t)ny * nz : (axis == 1 ? (size_t)nx * nz : (size_t)nx * ny);
    for (size_t line = (size_t)blockIdx.x * blockDim.x + threadIdx.x; line < nlines; line += (size_t)gridDim.x * blockDim.x) {
        size_t base, stride;
        if (axis == 0) {
            base = line * (size_t)nx;
            stride = 1;
        } else if (axis == 1) {
            base = (line / nx) * ((size_t)nx * ny) + (line % nx);
            stride = (size_t)nx;
        } else {
            base = line;
            stride = (size_t)nx * ny;
        }
        int l = -1;
        for (int i = 0; i < n; ++i) {
            const float di = out[base + i * stride];
            if (di != FLT_MAX) {
                const float iw = (float)i * spf;
                while (l >= 1 && maurer_remove(g[(size_t)(l - 1) * nlines + line], g[(size_t)l * nlines + line], di,
                                               (float)h[(size_t)(l - 1) * nlines + line] * spf, (float)h[(size_t)l * nlines + line] * spf, iw))
                    --l;
                ++l;
                g[(size_t)l * nlines + line] = di;
                h[(size_t)l * nlines + line] = i;
            }
        }
        if (l == -1) {
            // no site on this line: the filter leaves it untouched; the closing loop still visits every voxel
            if (LAST && !squared) {
                for (int i = 0; i < n; ++i) {
                    const size_t q = base + i * stride;
                    const float r = sqrtf(fabsf(out[q]));
                    out[q] = ((mask[q] != 0) == (inside_is_positive != 0)) ? r : -r;
                }
            }
            continue;
        }
        const int ns = l;
        l = 0;
        float gl = fabsf(g[line]), hl = (float)h[line] * spf;
        for (int i = 0; i < n; ++i) {
            const float iw = (float)i * spf;
            float d1 = gl + (hl - iw) * (hl - iw);
            while (l < ns) {
                const float gn = fabsf(g[(size_t)(l + 1) * nlines + line]), hn = (float)h[(size_t)(l + 1) * nlines + line] * spf;
                const float d2 = gn + (hn - iw) * (hn - iw);
                if (d1 <= d2) break;
                ++l;
                d1 = d2;
                gl = gn;
                hl = hn;
            }
            const size_t q = base + i * stride;
            float r = d1;
            if (LAST && !squared) r = sqrtf(fabsf(d1));
            out[q] = ((mask[q] != 0) == (inside_is_positive != 0)) ? r : -r;
        }
    }
}

// itk::LabelContourImageFilter (background 0): a labelled voxel stays if a neighbour inside the image (6 face
// neighbours, or all 26 with FULLY) carries a different value; everything else becomes background.
template <bool FULLY>
__global__ void __launch_bounds__(256) label_contour_kernel(const uint8_t* __restrict__ in, int nx, int ny, int nz, uint8_t* __restrict__ out)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const uint8_t v = in[q];
        uint8_t r = 0;
        if (v) {
            const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
            bool edge = false;
            if (FULLY) {
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int xx = x + dx, yy = y + dy, zz = z + dz;
                            if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                            if (in[((size_t)zz * ny + yy) * nx + xx] != v) edge = true;
                        }
            } else {
                if (x > 0 && in[q - 1] != v) edge = true;
                if (x + 1 < nx && in[q + 1] != v) edge = true;
                if (y > 0 && in[q - nx] != v) edge = true;
                if (y + 1 < ny && in[q + nx] != v) edge = true;
                if (z > 0 && in[q - (size_t)nx * ny] != v) edge = true;
                if (z + 1 < nz && in[q + (size_t)nx * ny] != v) edge = true;
            }
            if (edge) r = v;
        }
        out[q] = r;
    }
}

// sitk.LabelContour applied to every axial slice image[:, :, k] on its own (label/comparison.py:373-374): the four in-plane face
// neighbours only.
__global__ void __launch_bounds__(256) label_contour_slicewise_kernel(const uint8_t* __restrict__ in, int nx, int ny, int nz, uint8_t* __restrict__ out)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const uint8_t v = in[q];
        uint8_t r = 0;
        if (v) {
            const int x = (int)(q % nx), y = (int)((q / nx) % ny);
            bool edge = false;
            if (x > 0 && in[q - 1] != v) edge = true;
            if (x + 1 < nx && in[q + 1] != v) edge = true;
            if (y > 0 && in[q - nx] != v) edge = true;
            if (y + 1 < ny && in[q + nx] != v) edge = true;
            if (edge) r = v;
        }
        out[q] = r;
    }
}

// itk::BinaryDilateImageFilter / BinaryErodeImageFilter, foreground 1, background 0, structuring element given as offsets
// (dx, dy, dz).  Voxels that are not foreground keep their value unless the dilation paints them; `boundary_fg`: what lies
// outside the image counts as foreground (BinaryErode's default, boundaryToForeground = true) or background (BinaryDilate's).
template <bool DILATE>
__global__ void __launch_bounds__(256) binary_morph_kernel(const uint8_t* __restrict__ in, int nx, int ny, int nz, uint8_t* __restrict__ out,
                                                           const int* __restrict__ offs, int noffs, int boundary_fg)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const uint8_t v = in[q];
        uint8_t r = v;
        if (DILATE ? (v != 1) : (v == 1)) {
            const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
            bool hit = false;  // dilation: a foreground voxel whose element covers q; erosion: an element voxel that is not foreground
            for (int t = 0; t < noffs && !hit; ++t) {
                const int s = DILATE ? -1 : 1;  // dilation paints the element around every foreground voxel p: q = p + o
                const int xx = x + s * offs[3 * t], yy = y + s * offs[3 * t + 1], zz = z + s * offs[3 * t + 2];
                const bool inside = xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz;
                const bool fg = inside ? (in[((size_t)zz * ny + yy) * nx + xx] == 1) : (boundary_fg != 0);
                hit = DILATE ? fg : !fg;
            }
            if (hit) r = DILATE ? 1 : 0;
        }
        out[q] = r;
    }
}

// a | b, a & b, a + b (modulo 256 like sitk.Add on UInt8), a ^ b on UInt8 volumes (generation/dvf.py:66,249,290)
__global__ void __launch_bounds__(256) u8_binary_op_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int op, uint8_t* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const unsigned x = a[q], y = b[q];
        unsigned r;
        switch (op) {
        case 0: r = x | y; break;
        case 1: r = x & y; break;
        case 2: r = x + y; break;
        default: r = x ^ y; break;
        }
        out[q] = (uint8_t)r;
    }
}

// sitk.Mask(image, mask): planes of n voxels each; where the mask is 0 the pixel becomes `outside`
template <typename T>
__global__ void __launch_bounds__(256) mask_image_kernel(const T* __restrict__ in, const uint8_t* __restrict__ mask, size_t n, int planes, T outside, T* __restrict__ out)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const bool keep = mask[q] != 0;
        for (int c = 0; c < planes; ++c) out[(size_t)c * n + q] = keep ? in[(size_t)c * n + q] : outside;
    }
}

// image / constant (itk::DivideImageFilter with the constant in the pixel type)
template <typename T>
__global__ void __launch_bounds__(256) divide_scalar_kernel(const T* __restrict__ in, T divisor, T* __restrict__ out, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = in[q] / divisor;
}

// A constant displacement inside a mask (everywhere with mask == nullptr), 0 outside: the field templates of
// generate_field_shift / asymmetric_contract / asymmetric_extend (generation/dvf.py:54-66,114-121,187-200).  SoA planes.
__global__ void __launch_bounds__(256) constant_field_kernel(const uint8_t* __restrict__ mask, size_t n, double vx, double vy, double vz, double* __restrict__ out)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const bool keep = mask == nullptr || mask[q] != 0;
        out[q] = keep ? vx : 0.0;
        out[n + q] = keep ? vy : 0.0;
        out[2 * n + q] = keep ? vz : 0.0;
    }
}

// generate_field_radial_bend (generation/dvf.py:362-394): inside the body mask, cut by a half space through the reference
// voxel (clip_axis 0 = x, 1 = y, 2 = z, -1 = none; clip_keep_upper: voxels with index >= ref are kept, else index < ref),
// the displacement is scale * cross(voxel - ref, axis) with both vectors in (x, y, z) order -- numpy's cross: two rounded
// products and their difference per component.  0 elsewhere.
__global__ void __launch_bounds__(256) radial_bend_kernel(const uint8_t* __restrict__ mask, int nx, int ny, int nz, int rx, int ry, int rz, double ax, double ay,
                                                          double az, double scale, int clip_axis, int clip_keep_upper, double* __restrict__ out)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
        bool keep = mask[q] != 0;
        if (clip_axis >= 0) {
            const int c = clip_axis == 0 ? x : (clip_axis == 1 ? y : z);
            const int r = clip_axis == 0 ? rx : (clip_axis == 1 ? ry : rz);
            keep = keep && (clip_keep_upper ? (c >= r) : (c < r));
        }
        double ux = 0.0, uy = 0.0, uz = 0.0;
        if (keep) {
            const double vx = (double)(x - rx), vy = (double)(y - ry), vz = (double)(z - rz);
            const double p0 = vy * az, p1 = vz * ay;
            const double p2 = vz * ax, p3 = vx * az;
            const double p4 = vx * ay, p5 = vy * ax;
            ux = (p0 - p1) * scale;
            uy = (p2 - p3) * scale;
            uz = (p4 - p5) * scale;
        }
        out[q] = ux;
        out[n + q] = uy;
        out[2 * n + q] = uz;
    }
}

}  // namespace b200
