// distmap.cuh -- launch wrappers of distmap_kernels.cuh (distance maps, contours, binary morphology, masking and the
// field templates of the synthetic-deformation generators).  Everything is enqueued on the context's stream.
#pragma once
#include "common.cuh"
#include "distmap_kernels.cuh"

namespace b200 {

inline int elementwise_blocks(b200reg_ctx* ctx, size_t n, int threads)
{
    const size_t want = (n + threads - 1) / threads;
    const size_t cap = (size_t)ctx->sm_count * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

// sitk.SignedMaurerDistanceMap -> Float32.  Three Voronoi passes (x, y, z), each one thread per image line; the two scratch
// volumes hold the per-line stacks.
inline int signed_maurer(b200reg_ctx* ctx, const uint8_t* d_mask, const b200reg_geom& geom, int inside_is_positive, int squared, int use_spacing, float* d_out)
{
    const int nx = geom.size[0], ny = geom.size[1], nz = geom.size[2];
    const size_t n = (size_t)nx * ny * nz;
    TempBuf g, h;
    B200_TRY(g.alloc(ctx, n * sizeof(float)));
    B200_TRY(h.alloc(ctx, n * sizeof(int)));
    maurer_init_kernel<<<elementwise_blocks(ctx, n, 256), 256, 0, ctx->stream>>>(d_mask, nx, ny, nz, d_out);
    ctx->launches++;
    for (int axis = 0; axis < 3; ++axis) {
        const size_t nlines = axis == 0 ? (size_t)ny * nz : (axis == 1 ? (size_t)nx * nz : (size_t)nx * ny);
        const float spf = use_spacing ? (float)geom.spacing[axis] : 1.0f;
        const int nb = elementwise_blocks(ctx, nlines, 128);
        if (axis == 2)
            maurer_voronoi_kernel<true><<<nb, 128, 0, ctx->stream>>>(d_out, d_mask, nx, ny, nz, axis, spf, inside_is_positive, squared, g.as<float>(), h.as<int>());
        else
            maurer_voronoi_kernel<false><<<nb, 128, 0, ctx->stream>>>(d_out, d_mask, nx, ny, nz, axis, spf, inside_is_positive, squared, g.as<float>(), h.as<int>());
        ctx->launches++;
    }
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

inline int label_contour(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], int fully, uint8_t* d_out)
{
    const size_t n = (size_t)size[0] * size[1] * size[2];
    const int nb = elementwise_blocks(ctx, n, 256);
    if (fully)
        label_contour_kernel<true><<<nb, 256, 0, ctx->stream>>>(d_in, size[0], size[1], size[2], d_out);
    else
        label_contour_kernel<false><<<nb, 256, 0, ctx->stream>>>(d_in, size[0], size[1], size[2], d_out);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

inline int binary_morph(b200reg_ctx* ctx, bool dilate, const uint8_t* d_in, const int32_t size[3], const int* h_offsets, int noffs, int boundary_fg, uint8_t* d_out)
{
    const size_t n = (size_t)size[0] * size[1] * size[2];
    TempBuf offs;
    B200_TRY(offs.alloc(ctx, sizeof(int) * 3 * (size_t)noffs));
    B200_CUDA(cudaMemcpyAsync(offs.p, h_offsets, sizeof(int) * 3 * (size_t)noffs, cudaMemcpyHostToDevice, ctx->stream));
    const int nb = elementwise_blocks(ctx, n, 256);
    if (dilate)
        binary_morph_kernel<true><<<nb, 256, 0, ctx->stream>>>(d_in, size[0], size[1], size[2], d_out, offs.as<int>(), noffs, boundary_fg);
    else
        binary_morph_kernel<false><<<nb, 256, 0, ctx->stream>>>(d_in, size[0], size[1], size[2], d_out, offs.as<int>(), noffs, boundary_fg);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    B200_CUDA(cudaStreamSynchronize(ctx->stream));  // h_offsets is caller memory
    return B200REG_OK;
}

}  // namespace b200
