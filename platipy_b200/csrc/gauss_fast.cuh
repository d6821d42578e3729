// gauss_fast.cuh -- the fused 3-D separable field smoothing of the Demons loop for FAST MODE: float32 fields, float32 arithmetic with
// FMA.  Not a parity path: ITK's fields are Float64 (deformable.py:97-98) and the default build keeps them so; fast mode is the
// explicit switch of SURVEY.md section 8d (92 algorithmic B/voxel/iteration instead of 176) and is reported with its error against the
// parity path.  Same scheme as conv3d_zm2_kernel (gauss.cuh): one CTA owns a 32 x 16 column of one component and marches along z;
// the plane tile + halo arrives by ONE cp.async.bulk.tensor.3d (TMA, mbarrier completion), border tiles get their zero-filled halo
// cells replaced by the replicated edge values in shared memory; x pass and y pass out of shared memory, z pass over a register ring.
// What changes: 4-byte elements (half the HBM and shared-memory traffic), FP32 pipes (twice the lanes of FP64 on sm_100), one FFMA per
// tap instead of a DMUL + DADD pair.
#pragma once
#include "gauss.cuh"

namespace b200 {

template <int R, int TXW>
struct ZmfLayout {
    // x halo padded to a multiple of FOUR floats: the TMA unit wants the box to start on a 16-byte boundary of the row (a box starting
    // at x0 - 2 floats never completes its transaction -- found with compute-sanitizer: the mbarrier wait trapped)
    static constexpr int RP = (R + 3) & ~3;
    static constexpr int AW = TXW + 2 * RP, AH = ZM_TY + 2 * R, NA = AW * AH;
    static constexpr int NAP = (NA + 31) & ~31;  // floats: every staged plane starts on a 128-byte boundary
    static constexpr int NB = AH * TXW;
    static constexpr int NST = 2;
};

// false when the geometry cannot be described: row pitch and base must be multiples of 16 bytes
inline bool zmf_make_tensor_map(const float* base, int nx, int ny, long nz_total, int box_w, int box_h, int l2_promotion, CUtensorMap* out)
{
    zm_encode_tiled_fn enc = zm_encode_tiled();
    if (!enc || (nx % 4) != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0 || (box_w % 4) != 0 || box_w > 256 || box_h > 256) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz_total };
    const cuuint64_t strides[2] = { (cuuint64_t)nx * sizeof(float), (cuuint64_t)nx * ny * sizeof(float) };
    const cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    const CUtensorMapL2promotion promo = l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : (l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                              : (l2_promotion == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE));
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MODE 0: out = G(a).  MODE 3: out = b + G(a) (b fetched a plane step early).
template <int R, int RZ, int MODE, int TXW>
__global__ void __launch_bounds__(Zm2Threads<R, TXW>::value, 6) conv3d_zmf_kernel(const float* __restrict__ b, float* __restrict__ out, int nx, int ny, int nz,
                                                                                   int zchunk, int nchunks, const __grid_constant__ SmallCoeffs kc,
                                                                                   const DemonsCtrl* __restrict__ ctrl, int it,
                                                                                   const __grid_constant__ CUtensorMap tmap)
{
    pdl_launch_dependents();
    using L = ZmfLayout<R, TXW>;
    constexpr int RP = L::RP, AW = L::AW, AH = L::AH, NA = L::NA, NAP = L::NAP, NST = L::NST;
    constexpr int NR = 2 * RZ + 1;
    constexpr int NT = Zm2Threads<R, TXW>::value;
    constexpr int NYZ = 4 * TXW;
    constexpr int NLD = (NA + NT - 1) / NT;
    extern __shared__ __align__(128) float zmf_smem[];
    float* Aa = zmf_smem;              // [NST][NAP]
    float* B = zmf_smem + NST * NAP;   // [AH][TXW]
    __shared__ __align__(8) unsigned long long full_bar[NST];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TXW, y0 = blockIdx.y * ZM_TY;
    const int comp = blockIdx.z / nchunks, chunk = blockIdx.z % nchunks;
    const int z0 = chunk * zchunk, z1 = min(nz, z0 + zchunk);
    const size_t plane = (size_t)nx * ny, vol = plane * nz;
    const float* __restrict__ bp = MODE == 3 ? b + (size_t)comp * vol : nullptr;
    float* __restrict__ op = out + (size_t)comp * vol;

    // coefficients as float registers
    float kx[2 * R + 1], ky[2 * R + 1], kz[NR];
#pragma unroll
    for (int t = 0; t <= 2 * R; ++t) {
        kx[t] = (float)kc.k[0][t];
        ky[t] = (float)kc.k[1][t];
    }
#pragma unroll
    for (int t = 0; t < NR; ++t) kz[t] = (float)kc.k[2][t];

    // halo cells of a border tile that lie outside the image -> shared-memory index of the replicated edge value (-1: nothing to fix)
    int fix[NLD];
    const bool border = x0 - RP < 0 || x0 + TXW + RP > nx || y0 - R < 0 || y0 + ZM_TY + R > ny;
#pragma unroll
    for (int l = 0; l < NLD; ++l) {
        const int e = tid + l * NT;
        const int yy = e / AW, xx = e - yy * AW;
        const int ux = x0 - RP + xx, uy = y0 - R + yy;
        const int gx = ux < 0 ? 0 : (ux > nx - 1 ? nx - 1 : ux);
        const int gy = uy < 0 ? 0 : (uy > ny - 1 ? ny - 1 : uy);
        fix[l] = (e < NA && (gx != ux || gy != uy)) ? (gy - (y0 - R)) * AW + (gx - (x0 - RP)) : -1;
    }
    if (tid == 0) {
        tma_prefetch_descriptor(&tmap);
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    if (ctrl && it >= ctrl->halt_iter) return;

    auto stage = [&](int z, int buf) {
        const int zc = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
        if (tid == 0) {
            mbar_arrive_expect_tx(&full_bar[buf], (unsigned)(AH * AW * sizeof(float)));
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(Aa + buf * NAP)),
                         "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(x0 - RP), "r"(y0 - R), "r"(comp * nz + zc),
                         "r"((unsigned)__cvta_generic_to_shared(&full_bar[buf]))
                         : "memory");
        }
    };

    const int ox = tid & (TXW - 1), yb = tid / TXW;
    const int gx = x0 + ox;
    float ring[NR][4];
    const int zbeg = z0 - RZ, nsteps = (z1 - z0) + 2 * RZ;
    size_t out_off = (size_t)z0 * plane + (size_t)(y0 + 4 * yb) * nx + gx;  // running offset of this thread's outputs
    bool okj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) okj[j] = gx < nx && y0 + 4 * yb + j < ny;
    stage(zbeg, 0);
    for (int q0 = 0; q0 < nsteps; q0 += NR) {
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            const int q = q0 + s;
            if (q < nsteps) {
                const int buf = q & 1;
                float addv[4] = { 0.f, 0.f, 0.f, 0.f };
                if (MODE == 3 && tid < NYZ && q >= 2 * RZ) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (okj[j]) addv[j] = bp[out_off + (size_t)(j * nx)];
                }
                mbar_wait(&full_bar[buf], (unsigned)((q >> 1) & 1));
                if (border) {
#pragma unroll
                    for (int l = 0; l < NLD; ++l)
                        if (fix[l] >= 0) Aa[buf * NAP + tid + l * NT] = Aa[buf * NAP + fix[l]];
                    fence_proxy_async_smem();
                }
                __syncthreads();
                if (q + 1 < nsteps) stage(zbeg + q + 1, buf ^ 1);
                // ---- x pass: two output pairs of one row per task (rows incl. the y halo), 8-byte shared loads
                {
                    const int yy = tid / (TXW / 4), cx = tid & (TXW / 4 - 1);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int xb = 2 * cx + h * (TXW / 2);
                        const float2* ra = reinterpret_cast<const float2*>(Aa + buf * NAP + yy * AW + xb);
                        float w[2 + 2 * RP];
#pragma unroll
                        for (int v = 0; v < 1 + RP; ++v) {
                            const float2 t2 = ra[v];
                            w[2 * v] = t2.x;
                            w[2 * v + 1] = t2.y;
                        }
                        float o[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            float sum = kx[0] * w[j + (RP - R)];
#pragma unroll
                            for (int t = 1; t <= 2 * R; ++t) sum = __fmaf_rn(kx[t], w[j + (RP - R) + t], sum);
                            o[j] = sum;
                        }
                        *reinterpret_cast<float2*>(B + yy * TXW + xb) = make_float2(o[0], o[1]);
                    }
                }
                __syncthreads();
                // ---- y pass
                if (tid < NYZ) {
                    float col[4 + 2 * R];
#pragma unroll
                    for (int i = 0; i < 4 + 2 * R; ++i) col[i] = B[(4 * yb + i) * TXW + ox];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float sum = ky[0] * col[j];
#pragma unroll
                        for (int t = 1; t <= 2 * R; ++t) sum = __fmaf_rn(ky[t], col[j + t], sum);
                        ring[s][j] = sum;
                    }
                }
                // ---- z pass over the ring
                if (tid < NYZ && q >= 2 * RZ) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float sum = kz[0] * ring[(s + 1) % NR][j];
#pragma unroll
                        for (int t = 1; t < NR; ++t) sum = __fmaf_rn(kz[t], ring[(s + 1 + t) % NR][j], sum);
                        if (okj[j]) op[out_off + (size_t)(j * nx)] = MODE == 3 ? addv[j] + sum : sum;
                    }
                    out_off += plane;
                }
            }
        }
    }
}

inline bool zmarch_fast_supported(const KernelCoeffs kc[3], int nx)
{
    return kc[0].r == kc[1].r && kc[0].r >= 1 && kc[0].r <= ZM_RXY && kc[2].r >= 1 && kc[2].r <= ZM_RMAX && (nx % 4) == 0;
}

template <int R, int RZ>
inline int launch_zmf(b200reg_ctx* ctx, const float* a, const float* b, float* out, int nx, int ny, int nz, dim3 g, int zchunk, int nchunks, const SmallCoeffs& sc,
                      const DemonsCtrl* ctrl, int it, int nplanes)
{
    constexpr int TXW = 32;
    using L = ZmfLayout<R, TXW>;
    constexpr int NT = Zm2Threads<R, TXW>::value;
    constexpr size_t smem = (size_t)(L::NST * L::NAP + L::NB) * sizeof(float);
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (!zmf_make_tensor_map(a, nx, ny, (long)nz * nplanes, L::AW, L::AH, ctx->zm_tma_l2, &tm))
        return set_error(B200REG_ERR_UNSUPPORTED, "fast mode: the field cannot be described to the TMA unit (row length %d, base %p)", nx, (const void*)a);
    if (b) {
        B200_TRY(ensure_dynamic_smem(ctx, conv3d_zmf_kernel<R, RZ, 3, TXW>, smem));
        B200_CUDA(launch_pdl(ctx, conv3d_zmf_kernel<R, RZ, 3, TXW>, g, dim3(NT), smem, b, out, nx, ny, nz, zchunk, nchunks, sc, ctrl, it, tm));
    } else {
        B200_TRY(ensure_dynamic_smem(ctx, conv3d_zmf_kernel<R, RZ, 0, TXW>, smem));
        B200_CUDA(launch_pdl(ctx, conv3d_zmf_kernel<R, RZ, 0, TXW>, g, dim3(NT), smem, (const float*)nullptr, out, nx, ny, nz, zchunk, nchunks, sc, ctrl, it, tm));
    }
    return B200REG_OK;
}

// out <- G(a) (b == nullptr) or out <- b + G(a), three float32 component volumes; a / b / out must not alias
inline int launch_conv3d_zmarch_fast(b200reg_ctx* ctx, const float* a, const float* b, float* out, int nx, int ny, int nz, int nplanes, const KernelCoeffs kc[3],
                                     const DemonsCtrl* ctrl, int it)
{
    SmallCoeffs sc;
    for (int ax = 0; ax < 3; ++ax) {
        sc.r[ax] = kc[ax].r;
        for (int t = 0; t <= 2 * kc[ax].r; ++t) sc.k[ax][t] = kc[ax].k[t];
    }
    const int txw = 32;
    const int tiles = ((nx + txw - 1) / txw) * ((ny + ZM_TY - 1) / ZM_TY) * nplanes;
    const int min_chunk = tiles * ((nz + 15) / 16) >= ctx->sm_count * 4 ? 16 : 4;
    const int max_chunks = (nz + min_chunk - 1) / min_chunk;
    const long slots = (long)ctx->sm_count * 6;
    long best_cost = -1;
    int nchunks = 1;
    for (int k = 1; k <= 32 && k <= max_chunks; ++k) {
        const long rounds = ((long)tiles * k + slots - 1) / slots;
        const long cost = rounds * ((nz + k - 1) / k + 2 * kc[2].r);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            nchunks = k;
        }
    }
    const int zchunk = (nz + nchunks - 1) / nchunks;
    nchunks = (nz + zchunk - 1) / zchunk;
    dim3 g((nx + txw - 1) / txw, (ny + ZM_TY - 1) / ZM_TY, nplanes * nchunks);
    int rc;
#define ZMF_RZ(RR)                                                                                                     \
    switch (kc[2].r) {                                                                                                 \
    case 1: rc = launch_zmf<RR, 1>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, nplanes); break;      \
    case 2: rc = launch_zmf<RR, 2>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, nplanes); break;      \
    case 3: rc = launch_zmf<RR, 3>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, nplanes); break;      \
    default: rc = launch_zmf<RR, 4>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, nplanes); break;     \
    }
    switch (kc[0].r) {
    case 1: ZMF_RZ(1); break;
    case 2: ZMF_RZ(2); break;
    case 3: ZMF_RZ(3); break;
    default: ZMF_RZ(4); break;
    }
#undef ZMF_RZ
    B200_TRY(rc);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

}  // namespace b200
