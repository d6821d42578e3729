// postproc.cuh -- the small label utilities around the fusion step of the atlas pipeline (reference
// multiatlas/run.py:200-259, 387-437):
//   label_to_roi            utils/crop.py:24-71      bounding box of a mask (LabelStatisticsImageFilter::GetBoundingBox)
//   crop_to_roi / Paste     utils/crop.py:74-76, run.py:387-404   sub-volume copies (cudaMemcpy3DAsync)
//   correct_volume_overlap  label/utils.py:23-58     every voxel goes to the first structure (by volume rank) that has it
//   BinaryMorphologicalClosing  run.py:424           dilation then erosion with a ball, SafeBorder on
// Integer work, bit-exact by construction.
#pragma once
#include "common.cuh"

namespace b200 {

// bbox[0..2] = min index (x, y, z), bbox[3..5] = max index; initialised to (INT_MAX.., -1..)
__global__ void bbox_init_kernel(int* bbox)
{
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) bbox[threadIdx.x] = -1;
}
__global__ void __launch_bounds__(256) bbox_kernel(const uint8_t* __restrict__ in, int nx, int ny, int nz, int* __restrict__ bbox)
{
    const size_t n = (size_t)nx * ny * nz;
    int lo[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, hi[3] = { -1, -1, -1 };
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        if (in[q]) {
            const int c[3] = { (int)(q % nx), (int)((q / nx) % ny), (int)(q / ((size_t)nx * ny)) };
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                lo[d] = min(lo[d], c[d]);
                hi[d] = max(hi[d], c[d]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = min(lo[d], __shfl_down_sync(0xffffffffu, lo[d], o));
            hi[d] = max(hi[d], __shfl_down_sync(0xffffffffu, hi[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (hi[d] >= 0) {
                atomicMin(&bbox[d], lo[d]);
                atomicMax(&bbox[3 + d], hi[d]);
            }
        }
    }
}

constexpr int OVERLAP_MAX = 64;
struct LabelPtrs {
    const uint8_t* in[OVERLAP_MAX];
    uint8_t* out[OVERLAP_MAX];
    int n;
};
// labels are given in rank order (largest volume first by default): the first one that has the voxel keeps it
__global__ void __launch_bounds__(256) overlap_kernel(const __grid_constant__ LabelPtrs lp, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        bool taken = false;
        for (int s = 0; s < lp.n; ++s) {
            const bool has = lp.in[s][q] != 0;
            lp.out[s][q] = (has && !taken) ? 1 : 0;
            taken = taken || has;
        }
    }
}

// dilation (DILATE) / erosion of a {0, non-zero} mask with an arbitrary structuring element given as offsets.
// The input lives on a grid padded by `pad` voxels per side (background), the output on the grid `out_pad` smaller.
template <bool DILATE>
__global__ void __launch_bounds__(256) morph_kernel(const uint8_t* __restrict__ in, int inx, int iny, int inz, uint8_t* __restrict__ out, int onx, int ony, int onz,
                                                    int sx, int sy, int sz, const int* __restrict__ offs, int noffs)
{
    const size_t n = (size_t)onx * ony * onz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(q % onx) + sx, y = (int)((q / onx) % ony) + sy, z = (int)(q / ((size_t)onx * ony)) + sz;  // position in the input grid
        bool r = !DILATE;
        for (int t = 0; t < noffs; ++t) {
            const int xx = x + offs[3 * t], yy = y + offs[3 * t + 1], zz = z + offs[3 * t + 2];
            const bool inside = xx >= 0 && xx < inx && yy >= 0 && yy < iny && zz >= 0 && zz < inz;
            const bool v = inside && in[((size_t)zz * iny + yy) * inx + xx] != 0;
            if (DILATE) {
                if (v) {
                    r = true;
                    break;
                }
            } else if (!v) {
                r = false;
                break;
            }
        }
        out[q] = r ? 1 : 0;
    }
}

inline int region_copy(b200reg_ctx* ctx, const void* src, const int32_t src_size[3], const int32_t src_index[3], void* dst, const int32_t dst_size[3],
                       const int32_t dst_index[3], const int32_t region[3], size_t elem)
{
    for (int d = 0; d < 3; ++d) {
        if (region[d] <= 0 || src_index[d] < 0 || dst_index[d] < 0 || src_index[d] + region[d] > src_size[d] || dst_index[d] + region[d] > dst_size[d])
            return set_error(B200REG_ERR_RUNTIME, "requested region is (at least partially) outside the largest possible region");  // ITK's wording
    }
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), (size_t)src_size[0] * elem, (size_t)src_size[0] * elem, (size_t)src_size[1]);
    p.dstPtr = make_cudaPitchedPtr(dst, (size_t)dst_size[0] * elem, (size_t)dst_size[0] * elem, (size_t)dst_size[1]);
    p.srcPos = make_cudaPos((size_t)src_index[0] * elem, (size_t)src_index[1], (size_t)src_index[2]);
    p.dstPos = make_cudaPos((size_t)dst_index[0] * elem, (size_t)dst_index[1], (size_t)dst_index[2]);
    p.extent = make_cudaExtent((size_t)region[0] * elem, (size_t)region[1], (size_t)region[2]);
    p.kind = cudaMemcpyDeviceToDevice;
    B200_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    return B200REG_OK;
}

// itk::BinaryMorphologicalClosingImageFilter, SafeBorder on: pad with background by the radius, dilate, erode, crop --
// i.e. both operations on an unbounded grid that is zero outside the image.
inline int binary_closing(b200reg_ctx* ctx, const uint8_t* d_in, const int32_t size[3], const int32_t radius[3], const int* h_offsets, int noffs, uint8_t* d_out)
{
    const int nx = size[0], ny = size[1], nz = size[2];
    const int px = nx + 2 * radius[0], py = ny + 2 * radius[1], pz = nz + 2 * radius[2];
    const size_t np = (size_t)px * py * pz;
    TempBuf dil, offs;
    B200_TRY(dil.alloc(ctx, np));
    B200_TRY(offs.alloc(ctx, sizeof(int) * 3 * (size_t)(noffs > 0 ? noffs : 1)));
    B200_CUDA(cudaMemcpyAsync(offs.p, h_offsets, sizeof(int) * 3 * (size_t)noffs, cudaMemcpyHostToDevice, ctx->stream));
    const int nb = ctx->sm_count * 8;
    // dilation of the (virtually padded) input onto the padded grid: output voxel q sits at q - radius in the input grid
    morph_kernel<true><<<nb, 256, 0, ctx->stream>>>(d_in, nx, ny, nz, dil.as<uint8_t>(), px, py, pz, -radius[0], -radius[1], -radius[2], offs.as<int>(), noffs);
    // erosion back onto the image grid; outside the padded grid nothing is set
    morph_kernel<false><<<nb, 256, 0, ctx->stream>>>(dil.as<uint8_t>(), px, py, pz, d_out, nx, ny, nz, radius[0], radius[1], radius[2], offs.as<int>(), noffs);
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    B200_CUDA(cudaStreamSynchronize(ctx->stream));  // h_offsets is caller memory
    return B200REG_OK;
}

}  // namespace b200
