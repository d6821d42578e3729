// emu_distmap.cpp -- TEST INFRASTRUCTURE: runs the kernels of platipy_b200/csrc/distmap_kernels.cuh under the serial host
// emulation of cuda_emu.h, with the launch sequence of platipy_b200/csrc/distmap.cuh.  Built by tests/test_emu_distmap.py.
#include "cuda_emu.h"

#include <vector>

#include "../../platipy_b200/csrc/distmap_kernels.cuh"
#include "../../platipy_b200/csrc/patchcorr_kernels.cuh"
#include "../../platipy_b200/csrc/linreg_corr_kernels.cuh"
#include "../../platipy_b200/csrc/moments_kernels.cuh"
#include "../../platipy_b200/csrc/linreg_mattes_kernels.cuh"

using namespace b200;

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API void emu_signed_maurer(const uint8_t* mask, int nx, int ny, int nz, const double* spacing, int inside_pos, int squared, int use_spacing,
                               float* out, unsigned grid, unsigned block)
{
    const size_t n = (size_t)nx * ny * nz;
    std::vector<float> g(n);
    std::vector<int> h(n);
    emu_launch(maurer_init_kernel, grid, block, mask, nx, ny, nz, out);
    for (int axis = 0; axis < 3; ++axis) {
        const float spf = use_spacing ? (float)spacing[axis] : 1.0f;
        if (axis == 2)
            emu_launch(maurer_voronoi_kernel<true>, grid, block, out, mask, nx, ny, nz, axis, spf, inside_pos, squared, g.data(), h.data());
        else
            emu_launch(maurer_voronoi_kernel<false>, grid, block, out, mask, nx, ny, nz, axis, spf, inside_pos, squared, g.data(), h.data());
    }
}
EMU_API void emu_label_contour(const uint8_t* in, int nx, int ny, int nz, int fully, uint8_t* out, unsigned grid, unsigned block)
{
    if (fully) emu_launch(label_contour_kernel<true>, grid, block, in, nx, ny, nz, out);
    else emu_launch(label_contour_kernel<false>, grid, block, in, nx, ny, nz, out);
}
EMU_API void emu_binary_morph(const uint8_t* in, int nx, int ny, int nz, const int* offs, int noffs, int dilate, int boundary_fg, uint8_t* out,
                              unsigned grid, unsigned block)
{
    if (dilate) emu_launch(binary_morph_kernel<true>, grid, block, in, nx, ny, nz, out, offs, noffs, boundary_fg);
    else emu_launch(binary_morph_kernel<false>, grid, block, in, nx, ny, nz, out, offs, noffs, boundary_fg);
}
EMU_API void emu_u8_binary_op(const uint8_t* a, const uint8_t* b, int op, uint8_t* out, size_t n, unsigned grid, unsigned block)
{
    emu_launch(u8_binary_op_kernel, grid, block, a, b, op, out, n);
}
EMU_API void emu_mask_f64(const double* in, const uint8_t* mask, size_t n, int planes, double outside, double* out, unsigned grid, unsigned block)
{
    emu_launch(mask_image_kernel<double>, grid, block, in, mask, n, planes, outside, out);
}
EMU_API void emu_divide_f64(const double* in, double divisor, double* out, size_t n, unsigned grid, unsigned block)
{
    emu_launch(divide_scalar_kernel<double>, grid, block, in, divisor, out, n);
}
EMU_API void emu_constant_field(const uint8_t* mask, size_t n, double vx, double vy, double vz, double* out, unsigned grid, unsigned block)
{
    emu_launch(constant_field_kernel, grid, block, mask, n, vx, vy, vz, out);
}
EMU_API void emu_radial_bend(const uint8_t* mask, int nx, int ny, int nz, int rx, int ry, int rz, double ax, double ay, double az, double scale,
                             int clip_axis, int clip_keep_upper, double* out, unsigned grid, unsigned block)
{
    emu_launch(radial_bend_kernel, grid, block, mask, nx, ny, nz, rx, ry, rz, ax, ay, az, scale, clip_axis, clip_keep_upper, out);
}
EMU_API void emu_patch_correlation(const float* t, const float* m, int nx, int ny, int nz, int wx, int wy, int wz, double* out, unsigned grid,
                                   unsigned block)
{
    emu_launch(patch_correlation_kernel, grid, block, t, m, nx, ny, nz, wx, wy, wz, out);
}
EMU_API void emu_scale_shift_f64(const double* in, size_t n, int take_abs, double mul, double add, double* out, unsigned grid, unsigned block)
{
    emu_launch(scale_shift_kernel<double>, grid, block, in, n, take_abs, mul, add, out);
}
EMU_API void emu_scale_shift_f32(const float* in, size_t n, int take_abs, float mul, float add, float* out, unsigned grid, unsigned block)
{
    emu_launch(scale_shift_kernel<float>, grid, block, in, n, take_abs, mul, add, out);
}
// geometry: size (3 ints), origin (3), i2p (9), p2i (9) as the host wrapper fills them; pose: A (9), b (3), Bt (9), c (3).
// partials: [grid * block][42] per-thread sums (the CUDA build reduces them per block instead).
static void unpack(const int* fsize, const double* fgeo, const int* msize, const double* mgeo, const double* pose, CorrGeom& gf, CorrGeom& gm, CorrPose& ps)
{
    gf.nx = fsize[0]; gf.ny = fsize[1]; gf.nz = fsize[2];
    gm.nx = msize[0]; gm.ny = msize[1]; gm.nz = msize[2];
    for (int r = 0; r < 3; ++r) { gf.origin[r] = fgeo[r]; gm.origin[r] = mgeo[r]; }
    for (int r = 0; r < 9; ++r) { gf.i2p[r] = fgeo[3 + r]; gf.p2i[r] = fgeo[12 + r]; gm.i2p[r] = mgeo[3 + r]; gm.p2i[r] = mgeo[12 + r]; }
    for (int r = 0; r < 9; ++r) { ps.A[r] = pose[r]; ps.Bt[r] = pose[12 + r]; }
    for (int r = 0; r < 3; ++r) { ps.b[r] = pose[9 + r]; ps.c[r] = pose[21 + r]; }
}
// bins: n, fixed (bin size, normalised minimum), moving (bin size, normalised minimum); hist: [replicas * n * n + 1] integers
// (weights * 2^32 per replica, count last)
EMU_API void emu_linreg_mattes(const float* F, const float* M, const uint8_t* fmask, const uint8_t* mmask, const int* fsize, const double* fgeo,
                               const int* msize, const double* mgeo, const double* pose, int stride, int n_bins, const double* bins, int replicas,
                               unsigned long long* hist, const double* table, double* partials, unsigned grid, unsigned block)
{
    CorrGeom gf, gm;
    CorrPose ps;
    unpack(fsize, fgeo, msize, mgeo, pose, gf, gm, ps);
    MattesBins mb;
    mb.n = n_bins; mb.fbin = bins[0]; mb.fmin = bins[1]; mb.mbin = bins[2]; mb.mmin = bins[3];
    const size_t n = (size_t)gf.nx * gf.ny * gf.nz;
    const size_t nsamples = (n + (size_t)stride - 1) / (size_t)stride;
    if (hist)
        emu_launch(linreg_mattes_hist_kernel, grid, block, F, M, fmask, mmask, gf, gm, ps, mb, stride, nsamples, replicas, hist,
                   hist + (size_t)replicas * n_bins * n_bins);
    if (table) emu_launch(linreg_mattes_deriv_kernel, grid, block, F, M, fmask, mmask, gf, gm, ps, mb, stride, nsamples, table, partials);
}
EMU_API void emu_linreg_corr(const float* F, const float* M, const uint8_t* fmask, const uint8_t* mmask, const int* fsize, const double* fgeo,
                             const int* msize, const double* mgeo, const double* pose, int stride, double* partials, unsigned grid, unsigned block)
{
    CorrGeom gf, gm;
    CorrPose ps;
    unpack(fsize, fgeo, msize, mgeo, pose, gf, gm, ps);
    const size_t n = (size_t)gf.nx * gf.ny * gf.nz;
    const size_t nsamples = (n + (size_t)stride - 1) / (size_t)stride;
    emu_launch(linreg_corr_kernel, grid, block, F, M, fmask, mmask, gf, gm, ps, stride, nsamples, partials);
}
EMU_API void emu_image_moments(const float* img, const int* size, const double* geo, double* partials, unsigned grid, unsigned block)
{
    MomentsGeom g;
    g.nx = size[0]; g.ny = size[1]; g.nz = size[2];
    for (int r = 0; r < 3; ++r) g.origin[r] = geo[r];
    for (int r = 0; r < 9; ++r) g.i2p[r] = geo[3 + r];
    emu_launch(image_moments_kernel, grid, block, img, g, partials);
}
EMU_API void emu_label_contour_slicewise(const uint8_t* in, int nx, int ny, int nz, uint8_t* out, unsigned grid, unsigned block)
{
    emu_launch(label_contour_slicewise_kernel, grid, block, in, nx, ny, nz, out);
}
