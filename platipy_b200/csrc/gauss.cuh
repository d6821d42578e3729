// gauss.cuh -- itk::GaussianOperator coefficients (host) and separable clamp-boundary convolutions (device).
//   N1  DiscreteGaussianImageFilter        reference utils.py:226, fusion.py:168,279
//   N6  PDE field smoothing (update/displacement fields)  reference deformable.py:249-257
#pragma once
#include "common.cuh"

namespace b200 {

// ---- GaussianOperator::GenerateCoefficients ---------------------------------------------------------
// Discrete Gaussian e^{-t} I_n(t); the modified Bessel functions are the polynomial approximations ITK
// uses (Abramowitz & Stegun 9.8), so coefficients agree with ITK's to the last bit, not just to 1e-7.
inline double bessel_i0(double y)
{
    const double d = std::fabs(y);
    if (d < 3.75) {
        double m = y / 3.75;
        m *= m;
        return 1.0 + m * (3.5156229 + m * (3.0899424 + m * (1.2067492 + m * (0.2659732 + m * (0.360768e-1 + m * 0.45813e-2)))));
    }
    const double m = 3.75 / d;
    return (std::exp(d) / std::sqrt(d)) *
           (0.39894228 + m * (0.1328592e-1 + m * (0.225319e-2 + m * (-0.157565e-2 + m * (0.916281e-2 + m * (-0.2057706e-1 + m * (0.2635537e-1 + m * (-0.1647633e-1 + m * 0.392377e-2))))))));
}
inline double bessel_i1(double y)
{
    const double d = std::fabs(y);
    double acc;
    if (d < 3.75) {
        double m = y / 3.75;
        m *= m;
        acc = d * (0.5 + m * (0.87890594 + m * (0.51498869 + m * (0.15084934 + m * (0.2658733e-1 + m * (0.301532e-2 + m * 0.32411e-3))))));
    } else {
        const double m = 3.75 / d;
        acc = 0.2282967e-1 + m * (-0.2895312e-1 + m * (0.1787654e-1 - m * 0.420059e-2));
        acc = 0.39894228 + m * (-0.3988024e-1 + m * (-0.362018e-2 + m * (0.163801e-2 + m * (-0.1031555e-1 + m * acc))));
        acc *= (std::exp(d) / std::sqrt(d));
    }
    return y < 0.0 ? -acc : acc;
}
inline double bessel_in(int n, double y)
{
    if (y == 0.0) return 0.0;
    const double toy = 2.0 / std::fabs(y);
    double qip = 0.0, acc = 0.0, qi = 1.0;
    for (int j = 2 * (n + (int)std::sqrt(40.0 * n)); j > 0; j--) {
        const double qim = qip + j * toy * qi;
        qip = qi;
        qi = qim;
        if (std::fabs(qi) > 1.0e10) {
            acc *= 1.0e-10;
            qi *= 1.0e-10;
            qip *= 1.0e-10;
        }
        if (j == n) acc = qip;
    }
    acc *= bessel_i0(y) / qi;
    return (y < 0.0 && (n & 1)) ? -acc : acc;
}

// symmetric, normalised kernel of 2r+1 taps; terms are added until the running sum reaches 1 - max_error,
// a term drops below sum * DBL_EPSILON, or the one-sided length exceeds max_width.
inline std::vector<double> gaussian_operator(double variance, double max_error, int max_width)
{
    std::vector<double> c;
    const double et = std::exp(-variance), cap = 1.0 - max_error;
    double sum = 0.0;
    c.push_back(et * bessel_i0(variance));
    sum += c[0];
    c.push_back(et * bessel_i1(variance));
    sum += c[1] * 2.0;
    for (int i = 2; sum < cap; ++i) {
        c.push_back(et * bessel_in(i, variance));
        sum += c[i] * 2.0;
        if (c[i] < sum * DBL_EPSILON) break;
        if ((int)c.size() > max_width) break;
    }
    for (double& v : c) v /= sum;
    const int r = (int)c.size() - 1;
    std::vector<double> k(2 * r + 1);
    for (int i = 0; i <= r; ++i) {
        k[r + i] = c[i];
        k[r - i] = c[i];
    }
    return k;
}

// Coefficients travel to kernels by value (constant bank), up to radius KMAX_R.
constexpr int KMAX_R = 48;
struct KernelCoeffs {
    int r;
    double k[2 * KMAX_R + 1];
};
inline int make_coeffs(const std::vector<double>& k, KernelCoeffs* out)
{
    const int r = ((int)k.size() - 1) / 2;
    if (r > KMAX_R) return set_error(B200REG_ERR_UNSUPPORTED, "Gaussian kernel radius %d exceeds the supported %d", r, KMAX_R);
    out->r = r;
    for (size_t i = 0; i < k.size(); ++i) out->k[i] = k[i];
    return B200REG_OK;
}

// Demons loop control block (device resident; see demons.cuh).  Kernels launched for iteration `it` do
// nothing once it >= halt_iter, so a whole level is enqueued without any host synchronisation.
struct DemonsCtrl {
    int halt_iter;
    int elapsed;
    double metric;
    double rms;
};

// ---- 1-D convolution along AXIS, ZeroFluxNeumann (index clamp) boundary --------------------------------
// Inner product accumulated in double from offset -r to +r (itk::NeighborhoodInnerProduct order).
// ADD: the input is a + b evaluated on the fly (D + U of FastSymmetricForcesDemons::ApplyUpdate).
// gridDim.z = nz * nplanes (SoA planes are contiguous volumes).
template <typename T, int AXIS, bool ADD>
__global__ void __launch_bounds__(BX* BY) conv_axis_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                              int nx, int ny, int nz, const __grid_constant__ KernelCoeffs kc,
                                                              const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (ctrl && it >= ctrl->halt_iter) return;
    const int i = blockIdx.x * BX + threadIdx.x;
    const int j = blockIdx.y * BY + threadIdx.y;
    const int kz = blockIdx.z;  // plane * nz + k
    if (i >= nx || j >= ny) return;
    const int k = kz % nz;
    const size_t plane_off = (size_t)(kz / nz) * nx * ny * nz;
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nz);
    const int pos = AXIS == 0 ? i : (AXIS == 1 ? j : k);
    const size_t sa = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)nx : (size_t)nx * ny);
    const size_t o = plane_off + ((size_t)k * ny + j) * nx + i;
    const size_t base = o - (size_t)pos * sa;
    const int r = kc.r;
    double sum = 0.0;
    for (int t = -r; t <= r; ++t) {
        int q = pos + t;
        q = q < 0 ? 0 : (q > n - 1 ? n - 1 : q);
        const size_t idx = base + (size_t)q * sa;
        double v;
        if (ADD) v = (double)a[idx] + (double)b[idx];
        else v = (double)a[idx];
        sum += kc.k[t + r] * v;
    }
    out[o] = (T)sum;
}

template <typename T, bool ADD>
inline int launch_conv_axis(b200reg_ctx* ctx, int axis, const T* a, const T* b, T* out, int nx, int ny, int nz, int nplanes,
                            const KernelCoeffs& kc, const DemonsCtrl* ctrl, int it)
{
    dim3 g((nx + BX - 1) / BX, (ny + BY - 1) / BY, nz * nplanes), blk(BX, BY, 1);
    if (axis == 0) conv_axis_kernel<T, 0, ADD><<<g, blk, 0, ctx->stream>>>(a, b, out, nx, ny, nz, kc, ctrl, it);
    else if (axis == 1) conv_axis_kernel<T, 1, ADD><<<g, blk, 0, ctx->stream>>>(a, b, out, nx, ny, nz, kc, ctrl, it);
    else conv_axis_kernel<T, 2, ADD><<<g, blk, 0, ctx->stream>>>(a, b, out, nx, ny, nz, kc, ctrl, it);
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- fused 3-D separable smoothing, z-marching ---------------------------------------------------------------
// One CTA owns a TX x TY column of one component plane-stack and marches along z.  Per z step it stages the
// input plane (tile + x/y halo, index-clamped = ZeroFluxNeumann) in shared memory, runs the x pass and the
// y pass out of shared memory, and keeps the last 2*RZ+1 x/y-smoothed planes of its own voxels in a register
// ring; the z pass is an inner product over that ring.  Every input voxel is read from HBM once (halo
// re-reads are L2 hits) and every output written once: 16 B/voxel/component instead of 48 for three
// separable passes.  Operation order per output (x taps -r..r, then y, then z; each rounded to double)
// is identical to three sequential passes, so results are bit-identical to them (and to the oracle).
#ifndef ZM_MINB
#define ZM_MINB 2
#endif
constexpr int ZM_TX = 64, ZM_TY = 16, ZM_NT = 256, ZM_RXY = 4, ZM_RMAX = 4;
constexpr int ZM_PER = ZM_TY / (ZM_NT / ZM_TX);                                            // own voxels per thread
constexpr int ZM_MAXLD = ((ZM_TX + 2 * ZM_RXY) * (ZM_TY + 2 * ZM_RXY) + ZM_NT - 1) / ZM_NT;  // staged loads per thread
struct SmallCoeffs {
    int r[3];
    double k[3][2 * ZM_RMAX + 1];
};

template <int RZ, bool ADD>
__global__ void __launch_bounds__(ZM_NT, 2) conv3d_zmarch_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                                                                  int nx, int ny, int nz, int zchunk, int nchunks,
                                                                  const __grid_constant__ SmallCoeffs kc, const DemonsCtrl* __restrict__ ctrl, int it)
{
    if (ctrl && it >= ctrl->halt_iter) return;
    __shared__ double A[(ZM_TY + 2 * ZM_RXY) * (ZM_TX + 2 * ZM_RXY)];
    __shared__ double B[(ZM_TY + 2 * ZM_RXY) * ZM_TX];
    const int rx = kc.r[0], ry = kc.r[1];
    const int AW = ZM_TX + 2 * rx, AH = ZM_TY + 2 * ry, NA = AW * AH;
    const int tid = threadIdx.x, tx = tid % ZM_TX, ty = tid / ZM_TX;
    const int x0 = blockIdx.x * ZM_TX, y0 = blockIdx.y * ZM_TY;
    const int comp = blockIdx.z / nchunks, chunk = blockIdx.z % nchunks;
    const int z0 = chunk * zchunk, z1 = min(nz, z0 + zchunk);
    const size_t plane = (size_t)nx * ny, vol = plane * nz;
    const double* __restrict__ ap = a + (size_t)comp * vol;
    const double* __restrict__ bp = ADD ? b + (size_t)comp * vol : nullptr;
    double* __restrict__ op = out + (size_t)comp * vol;

    // staged-load bookkeeping: element e of A <-> (clamped) global offset within a plane
    int goff[ZM_MAXLD];
#pragma unroll
    for (int l = 0; l < ZM_MAXLD; ++l) {
        const int e = tid + l * ZM_NT;
        int yy = e / AW, xx = e - yy * AW;
        int gx = x0 - rx + xx, gy = y0 - ry + yy;
        gx = gx < 0 ? 0 : (gx > nx - 1 ? nx - 1 : gx);
        gy = gy < 0 ? 0 : (gy > ny - 1 ? ny - 1 : gy);
        goff[l] = e < NA ? gy * nx + gx : -1;
    }
    double pre[ZM_MAXLD];
    auto fetch = [&](int z) {
        const int zc = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
        const size_t zo = (size_t)zc * plane;
#pragma unroll
        for (int l = 0; l < ZM_MAXLD; ++l)
            if (goff[l] >= 0) {
                if (ADD) pre[l] = ap[zo + goff[l]] + bp[zo + goff[l]];
                else pre[l] = ap[zo + goff[l]];
            }
    };

    double ring[ZM_PER][2 * RZ + 1];
    double cur[ZM_PER];
    const int zbeg = z0 - RZ, zend = z1 - 1 + RZ;
    int prev_zc = -1;
    fetch(zbeg);
    for (int z = zbeg; z <= zend; ++z) {
        const int zc = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
        if (zc != prev_zc) {  // clamped repeats of a border plane reuse `cur`
#pragma unroll
            for (int l = 0; l < ZM_MAXLD; ++l)
                if (goff[l] >= 0) A[tid + l * ZM_NT] = pre[l];
            __syncthreads();
            if (z < zend) fetch(z + 1);  // in flight during the passes below
            // x pass over tile rows + y halo
            for (int e = tid; e < AH * ZM_TX; e += ZM_NT) {
                const int yy = e / ZM_TX, xx = e - yy * ZM_TX;
                const double* row = A + yy * AW + xx;
                double sum = 0.0;
                for (int t = 0; t <= 2 * rx; ++t) sum += kc.k[0][t] * row[t];
                B[e] = sum;
            }
            __syncthreads();
            // y pass for the thread's own voxels
#pragma unroll
            for (int j = 0; j < ZM_PER; ++j) {
                const int y = ty + j * (ZM_NT / ZM_TX);
                const double* col = B + y * ZM_TX + tx;
                double sum = 0.0;
                for (int t = 0; t <= 2 * ry; ++t) sum += kc.k[1][t] * col[t * ZM_TX];
                cur[j] = sum;
            }
            prev_zc = zc;
        } else if (z < zend) {
            // the next distinct plane still has to be fetched once the clamped run ends
            const int zn = z + 1 < 0 ? 0 : (z + 1 > nz - 1 ? nz - 1 : z + 1);
            if (zn != zc) fetch(z + 1);
        }
        // ring shift + z pass
#pragma unroll
        for (int j = 0; j < ZM_PER; ++j) {
#pragma unroll
            for (int t = 0; t < 2 * RZ; ++t) ring[j][t] = ring[j][t + 1];
            ring[j][2 * RZ] = cur[j];
        }
        const int zo = z - RZ;
        if (zo >= z0) {
            const int gx = x0 + tx;
#pragma unroll
            for (int j = 0; j < ZM_PER; ++j) {
                const int gy = y0 + ty + j * (ZM_NT / ZM_TX);
                double sum = 0.0;
#pragma unroll
                for (int t = 0; t <= 2 * RZ; ++t) sum += kc.k[2][t] * ring[j][t];
                if (gx < nx && gy < ny) op[(size_t)zo * plane + (size_t)gy * nx + gx] = sum;
            }
        }
    }
}

// ---- fused 3-D smoothing, second generation: compile-time radii, cp.async double-buffered planes -------------
// Same z-marching scheme and the same operation order (bit-identical results), engineered for instruction
// count: radii are template parameters (taps fully unrolled, coefficients are constant-bank operands), the
// x pass works on 4 consecutive outputs per thread out of 128-bit shared loads, the y pass slides a register
// window over 4 consecutive rows, the z ring is addressed by compile-time rotation (no register moves), and
// the next plane is staged global -> shared with cp.async (LDGSTS) while the current one is processed.
// R: x/y radius (equal), RZ: z radius, ADD: input is a + b.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Threads per CTA: one x-pass task (row, column group) per thread -> (TY + 2R) * (TXW / 4) threads; the first 4 * TXW of
// them also own the tile's voxels in the y and z passes.  TXW = 64: 2 CTAs per SM; TXW = 32: 4 CTAs per SM (same warps,
// twice the independent barrier domains).
template <int R, int TXW = ZM_TX>
struct Zm2Threads {
    static constexpr int value = (ZM_TY + 2 * R) * (TXW / 4);
};
// ---- TMA staging (cp.async.bulk.tensor.3d -> SASS UTMALDG) with mbarrier completion --------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned a, unsigned parity)
{
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    // A transaction count that never completes (a tensor map that does not match the launch) would spin for ever; after about two
    // seconds of waiting the kernel traps instead, which turns a hang into a reported error.
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    if (mbar_try_wait(a, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(a, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}
// generic-proxy writes to shared memory (the border fix-up below) ordered before later async-proxy (TMA) writes to the same buffer
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace b200
#include <cuda.h>  // CUtensorMap and the cuTensorMapEncodeTiled prototype (the entry point itself comes from the runtime, no libcuda link)
namespace b200 {
// Tensor-map staging of the fused smoothing kernel: ONE cp.async.bulk.tensor.3d per plane tile.  The field is described to the TMA
// unit as a 3-D tensor (x, y, component * nz + z) of float64; the staged tile + halo is a box (AW, AH, 1) of it whose corner may
// lie outside the image: the TMA unit fills out-of-range elements with zeros, and the few halo cells of a BORDER tile are then
// overwritten in shared memory with the replicated edge values (ZeroFluxNeumann), which are part of the same box.  z is clamped
// through the coordinate.
typedef CUresult (*zm_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline zm_encode_tiled_fn zm_encode_tiled()
{
    static zm_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<zm_encode_tiled_fn>(p);
    }
    return fn;
}
// false when the geometry cannot be described (odd nx: row pitch not a multiple of 16 bytes) or the driver entry point is missing
inline bool zm_make_tensor_map(const double* base, int nx, int ny, long nz_total, int box_w, int box_h, int l2_promotion, CUtensorMap* out)
{
    zm_encode_tiled_fn enc = zm_encode_tiled();
    if (!enc || (nx % 2) != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0 || box_w > 256 || box_h > 256) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz_total };
    const cuuint64_t strides[2] = { (cuuint64_t)nx * sizeof(double), (cuuint64_t)nx * ny * sizeof(double) };
    const cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    const CUtensorMapL2promotion promo = l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : (l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                              : (l2_promotion == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE));
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
__device__ __forceinline__ void tma_tensor3d_g2s(double* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap* tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(tmap)) : "memory");
}

// MODE 0: out = G(a).  MODE 1: out = G(a + b), both operands staged with cp.async and added in the x pass.
// MODE 2: out = G(a + b), the next plane's operands are loaded into registers while the current plane is processed and
// their sum is what gets staged: one shared buffer less, half the staging writes and x-pass reads of MODE 1 (A/B builds only).
// MODE 3: out = b + G(a): the second operand is added to the finished value when it is stored (one coalesced load per
// output, no shared-memory traffic).  The Demons loop uses it to form D + G_u * U at the end of the update smoothing, so
// that the displacement smoothing that follows is a plain MODE 0 pass instead of a MODE 1 pass.
// TMA (MODE 0 / 3 only): plane tiles are staged by the TMA unit (one elected thread, one tensor-map copy per plane, completion
// counted in bytes on an mbarrier) instead of one cp.async per element by every thread; same shared layout, same arithmetic.
#ifndef ZM_TMA_STAGES
#define ZM_TMA_STAGES 2
#endif
template <int R, int TXW>
struct Zm2Layout {
    static constexpr int RP = (R + 1) & ~1;  // x halo padded to an even count: 16-byte aligned shared rows
    static constexpr int AW = TXW + 2 * RP, AH = ZM_TY + 2 * R, NA = AW * AH;
    static constexpr int NAP = (NA + 15) & ~15;  // buffer stride: every staged plane starts on a 128-byte boundary (TMA destination)
    static constexpr int NB = AH * TXW;
    // staged planes in flight: cp.async stages through registers-free but thread-issued copies, one plane ahead; the TMA unit needs no
    // thread resources, so the tensor-map path can run ZM_TMA_STAGES - 1 planes ahead.  Measured (profiles/r02c_ab_tma_stages.log, full-
    // resolution iteration): 2 staged planes 2.82 ms, 3: 2.86, 4: 2.89, 6: 2.84, cp.async 2.92 -- depth buys nothing, the wait on the
    // mbarrier is bandwidth, not latency; the default stays at 2
    static constexpr int STAGES_TMA = ZM_TMA_STAGES, STAGES_CP = 2;
};
#ifndef ZM_MODE0_CTAS
#define ZM_MODE0_CTAS 5  // resident CTAs per SM the plain tensor-map pass is compiled for (72 registers)
#endif
template <int R, int RZ, int MODE, int TXW, bool TMA>
__global__ void __launch_bounds__(Zm2Threads<R, TXW>::value, (TMA && MODE == 0 && TXW == 32) ? ZM_MODE0_CTAS : 128 / TXW) conv3d_zm2_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                                                               int nx, int ny, int nz, int zchunk, int nchunks,
                                                               const __grid_constant__ SmallCoeffs kc, const DemonsCtrl* __restrict__ ctrl, int it,
                                                               const __grid_constant__ CUtensorMap tmap)
{
    pdl_launch_dependents();
    static_assert(!TMA || MODE == 0 || MODE == 3, "the tensor-map staging handles one staged operand");
    constexpr bool ADD = MODE == 1;     // second operand staged in shared memory
    constexpr bool REGADD = MODE == 2;  // operands summed in registers before staging
    using L = Zm2Layout<R, TXW>;
    constexpr int RP = L::RP, AW = L::AW, AH = L::AH, NA = L::NA, NAP = L::NAP;
    constexpr int NR = 2 * RZ + 1;
    constexpr int NT = Zm2Threads<R, TXW>::value;
    constexpr int NYZ = 4 * TXW;  // threads that own voxels in the y / z passes
    constexpr int NLD = (NA + NT - 1) / NT;
    constexpr int NST = TMA ? L::STAGES_TMA : L::STAGES_CP;
    extern __shared__ __align__(128) double zm_smem[];
    double* Aa = zm_smem;                              // [NST][NAP]
    double* Ab = zm_smem + NST * NAP;                  // [NST][NAP] (ADD only)
    double* B = zm_smem + (ADD ? 2 : 1) * NST * NAP;   // [AH][TX]

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TXW, y0 = blockIdx.y * ZM_TY;
    const int comp = blockIdx.z / nchunks, chunk = blockIdx.z % nchunks;
    const int z0 = chunk * zchunk, z1 = min(nz, z0 + zchunk);
    const size_t plane = (size_t)nx * ny, vol = plane * nz;
    const double* __restrict__ ap = a + (size_t)comp * vol;
    const double* __restrict__ bp = MODE != 0 ? b + (size_t)comp * vol : nullptr;

    // cp.async path: element e of a staged plane <-> clamped global offset.  TMA path: the same slots hold, for the halo cells of a
    // border tile that lie outside the image, the shared-memory index of the replicated edge value (-1: nothing to fix).
    int goff[NLD];
    __shared__ __align__(8) unsigned long long full_bar[NST];
    const bool border = x0 - RP < 0 || x0 + TXW + RP > nx || y0 - R < 0 || y0 + ZM_TY + R > ny;
#pragma unroll
    for (int l = 0; l < NLD; ++l) {
        const int e = tid + l * NT;
        const int yy = e / AW, xx = e - yy * AW;
        const int ux = x0 - RP + xx, uy = y0 - R + yy;
        const int gx = ux < 0 ? 0 : (ux > nx - 1 ? nx - 1 : ux);
        const int gy = uy < 0 ? 0 : (uy > ny - 1 ? ny - 1 : uy);
        if (TMA) goff[l] = (e < NA && (gx != ux || gy != uy)) ? (gy - (y0 - R)) * AW + (gx - (x0 - RP)) : -1;
        else goff[l] = e < NA ? gy * nx + gx : -1;
    }
    if (TMA) {
        if (tid == 0) {
            tma_prefetch_descriptor(&tmap);
#pragma unroll
            for (int b = 0; b < NST; ++b) mbar_init(&full_bar[b], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    // everything above is index arithmetic and barrier set-up; from here on the kernel reads what its predecessor wrote
    pdl_wait();
    if (ctrl && it >= ctrl->halt_iter) return;
    double pre_a[REGADD ? NLD : 1], pre_b[REGADD ? NLD : 1];  // MODE 2: operands of the next plane, in flight
    auto load_next = [&](int z) {
        const int zc = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
        const size_t zo = (size_t)zc * plane;
#pragma unroll
        for (int l = 0; l < NLD; ++l)
            if (goff[l] >= 0) {
                pre_a[l] = ap[zo + goff[l]];
                pre_b[l] = bp[zo + goff[l]];
            }
    };
    auto store_next = [&](int buf) {
#pragma unroll
        for (int l = 0; l < NLD; ++l)
            if (goff[l] >= 0) Aa[buf * NAP + tid + l * NT] = pre_a[l] + pre_b[l];
    };
    auto stage = [&](int z, int buf) {
        const int zc = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
        if (TMA) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&full_bar[buf], (unsigned)(AH * AW * sizeof(double)));
                tma_tensor3d_g2s(Aa + buf * NAP, &tmap, x0 - RP, y0 - R, comp * nz + zc, &full_bar[buf]);
            }
            return;
        }
        const size_t zo = (size_t)zc * plane;
#pragma unroll
        for (int l = 0; l < NLD; ++l)
            if (goff[l] >= 0) {
                cp_async8(Aa + buf * NAP + tid + l * NT, ap + zo + goff[l]);
                if (ADD) cp_async8(Ab + buf * NAP + tid + l * NT, bp + zo + goff[l]);
            }
        cp_async_commit();
    };

    // y/z-pass ownership: column x = tid % 64, rows 4*yb .. 4*yb+3
    const int ox = tid & (TXW - 1), yb = tid / TXW;
    const int gx = x0 + ox;
    double ring[NR][4];
    const int zbeg = z0 - RZ, nsteps = (z1 - z0) + 2 * RZ;
    // offsets of this thread's four outputs inside a component volume: a running plane offset (one 64-bit add per step) plus the row
    // offsets, instead of forming z * plane + y * nx + x with wide multiplies for every store (ncu: a quarter of the loop's instructions)
    // (kept as running POINTERS: with the offset relative to the component volume the compiler re-formed component * volume -- a chain of
    // wide multiplies -- in front of every store, a dozen integer instructions per output; SASS of the round-2 build)
    const size_t out_off0 = (size_t)comp * vol + (size_t)z0 * plane + (size_t)(y0 + 4 * yb) * nx + gx;
    double* __restrict__ oq = out + out_off0;
    const double* __restrict__ bq = MODE == 3 ? b + out_off0 : nullptr;
    bool okj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) okj[j] = gx < nx && y0 + 4 * yb + j < ny;
    if (REGADD) {
        load_next(zbeg);
        store_next(0);
    } else {
#pragma unroll
        for (int b = 0; b < NST - 1; ++b)
            if (b < nsteps) stage(zbeg + b, b);
    }
    for (int q0 = 0; q0 < nsteps; q0 += NR) {
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            const int q = q0 + s;
            if (q < nsteps) {
                const int buf = q % NST;
                // MODE 3: the operand added at the store is fetched now, a whole plane step before it is needed
                double addv[4] = { 0.0, 0.0, 0.0, 0.0 };
                if (MODE == 3 && tid < NYZ && q >= 2 * RZ) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (okj[j]) addv[j] = bq[j * nx];
                }
                if (REGADD) {
                    __syncthreads();  // plane q (summed and stored at the end of step q - 1) is visible; B may be rewritten
                    if (q + 1 < nsteps) load_next(zbeg + q + 1);
                } else {
                    if (TMA) {
                        mbar_wait(&full_bar[buf], (unsigned)((q / NST) & 1));
                        if (border) {
                            // replicate the image edge into the zero-filled halo cells (sources are cells inside the image, targets
                            // cells outside it: disjoint, so no barrier is needed between the reads and the writes)
#pragma unroll
                            for (int l = 0; l < NLD; ++l)
                                if (goff[l] >= 0) Aa[buf * NAP + tid + l * NT] = Aa[buf * NAP + goff[l]];
                            fence_proxy_async_smem();  // these generic-proxy stores precede the TMA overwrite of this buffer two steps on
                        }
                    } else {
                        cp_async_wait_all();
                    }
                    __syncthreads();
                    // the buffer refilled now was last read by the x pass of step q - 1, which every thread has left
                    if (q + NST - 1 < nsteps) stage(zbeg + q + NST - 1, (q + NST - 1) % NST);
                }
                // ---- x pass: per task two output pairs {2cx, 2cx+1} and {32+2cx, 32+2cx+1} of one row (rows incl. the
                // y halo).  Lanes read consecutive 16-byte words: conflict-free 128-bit shared loads and stores.
                {
                    const int yy = tid / (TXW / 4), cx = tid & (TXW / 4 - 1);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int xb = 2 * cx + h * (TXW / 2);
                        const double2* ra = reinterpret_cast<const double2*>(Aa + buf * NAP + yy * AW + xb);
                        double w[2 + 2 * RP];
#pragma unroll
                        for (int v = 0; v < 1 + RP; ++v) {
                            double2 t2 = ra[v];
                            if (ADD) {
                                const double2 u2 = reinterpret_cast<const double2*>(Ab + buf * NAP + yy * AW + xb)[v];
                                t2.x = t2.x + u2.x;
                                t2.y = t2.y + u2.y;
                            }
                            w[2 * v] = t2.x;
                            w[2 * v + 1] = t2.y;
                        }
                        double o[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            // first tap without the leading "0.0 +": same value (only the sign of an all-zero sum can differ)
                            double sum = kc.k[0][0] * w[j + (RP - R)];
#pragma unroll
                            for (int t = 1; t <= 2 * R; ++t) sum += kc.k[0][t] * w[j + (RP - R) + t];
                            o[j] = sum;
                        }
                        *reinterpret_cast<double2*>(B + yy * TXW + xb) = make_double2(o[0], o[1]);
                    }
                }
                __syncthreads();
                // ---- y pass: sliding window over 4 + 2R rows of this thread's column
                if (tid < NYZ) {
                    double col[4 + 2 * R];
#pragma unroll
                    for (int i = 0; i < 4 + 2 * R; ++i) col[i] = B[(4 * yb + i) * TXW + ox];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double sum = kc.k[1][0] * col[j];
#pragma unroll
                        for (int t = 1; t <= 2 * R; ++t) sum += kc.k[1][t] * col[j + t];
                        ring[s][j] = sum;
                    }
                }
                // ---- z pass over the ring (slot s is the newest plane; oldest is slot (s + 1) % NR)
                if (tid < NYZ && q >= 2 * RZ) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double sum = kc.k[2][0] * ring[(s + 1) % NR][j];
#pragma unroll
                        for (int t = 1; t < NR; ++t) sum += kc.k[2][t] * ring[(s + 1 + t) % NR][j];
                        if (okj[j]) oq[j * nx] = MODE == 3 ? addv[j] + sum : sum;
                    }
                    oq += plane;
                    if (MODE == 3) bq += plane;
                }
                // MODE 2: buffer buf ^ 1 was last read by the x pass of step q - 1, two barriers ago
                if (REGADD && q + 1 < nsteps) store_next((q + 1) % NST);
            }
        }
    }
}

template <int R, int RZ, int TXW>
inline int launch_zm2_tx(b200reg_ctx* ctx, const double* a, const double* b, double* out, int nx, int ny, int nz, dim3 g, int zchunk, int nchunks,
                         const SmallCoeffs& sc, const DemonsCtrl* ctrl, int it, bool addout)
{
    using L = Zm2Layout<R, TXW>;
    constexpr int NT = Zm2Threads<R, TXW>::value;
    g.x = (nx + TXW - 1) / TXW;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    // the tensor map describes operand a: (nx, ny, components * nz) float64, box = one staged plane tile (B200REG_ZM_TMA=0: cp.async staging)
    const bool tma = ctx->zm_tma != 0 && zm_make_tensor_map(a, nx, ny, (long)nz * (long)(g.z / nchunks), L::AW, L::AH, ctx->zm_tma_l2, &tm);
#define ZM2_GO(MODE, TMA_, A, B_)                                                                                                            \
    do {                                                                                                                                     \
        constexpr size_t smem1 = (size_t)((TMA_ ? L::STAGES_TMA : L::STAGES_CP) * L::NAP + L::NB) * sizeof(double);                                                \
        B200_TRY(ensure_dynamic_smem(ctx, conv3d_zm2_kernel<R, RZ, MODE, TXW, TMA_>, smem1));                                                \
        B200_CUDA(launch_pdl(ctx, conv3d_zm2_kernel<R, RZ, MODE, TXW, TMA_>, g, dim3(NT), smem1, A, B_, out, nx, ny, nz, zchunk, nchunks, sc, ctrl, it, tm)); \
    } while (0)
    if (b && addout) {
        if (tma) ZM2_GO(3, true, a, b);
        else ZM2_GO(3, false, a, b);
        return B200REG_OK;
    }
#ifdef B200REG_AB_VARIANTS
    if (b && ctx->zm_regadd) {
        ZM2_GO(2, false, a, b);
        return B200REG_OK;
    }
#endif
    if (b) {
        constexpr size_t smem2 = (size_t)(4 * L::NAP + L::NB) * sizeof(double);
        B200_TRY(ensure_dynamic_smem(ctx, conv3d_zm2_kernel<R, RZ, 1, TXW, false>, smem2));
        B200_CUDA(launch_pdl(ctx, conv3d_zm2_kernel<R, RZ, 1, TXW, false>, g, dim3(NT), smem2, a, b, out, nx, ny, nz, zchunk, nchunks, sc, ctrl, it, tm));
    } else {
        if (tma) ZM2_GO(0, true, a, nullptr);
        else ZM2_GO(0, false, a, nullptr);
    }
#undef ZM2_GO
    return B200REG_OK;
}
template <int R, int RZ>
inline int launch_zm2_rz(b200reg_ctx* ctx, const double* a, const double* b, double* out, int nx, int ny, int nz, dim3 g, int zchunk, int nchunks,
                         const SmallCoeffs& sc, const DemonsCtrl* ctrl, int it, bool addout)
{
#ifdef B200REG_AB_VARIANTS  // measured alternative (profiles/r01_summary.md), compiled with make EXTRA=-DB200REG_AB_VARIANTS
    if (!ctx->zm_tx32) return launch_zm2_tx<R, RZ, 64>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout);
#endif
    return launch_zm2_tx<R, RZ, 32>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout);
}
template <int R>
inline int launch_zm2_r(b200reg_ctx* ctx, int rz, const double* a, const double* b, double* out, int nx, int ny, int nz, dim3 g, int zchunk,
                        int nchunks, const SmallCoeffs& sc, const DemonsCtrl* ctrl, int it, bool addout)
{
    switch (rz) {
    case 1: return launch_zm2_rz<R, 1>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout);
    case 2: return launch_zm2_rz<R, 2>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout);
    case 3: return launch_zm2_rz<R, 3>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout);
    default: return launch_zm2_rz<R, 4>(ctx, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout);
    }
}

inline bool zmarch_supported(const KernelCoeffs kc[3])
{
    return kc[0].r <= ZM_RXY && kc[1].r <= ZM_RXY && kc[2].r >= 1 && kc[2].r <= ZM_RMAX;
}
// the compile-time-radii kernel (needed for the add-at-output form)
inline bool zmarch2_supported(b200reg_ctx* ctx, const KernelCoeffs kc[3])
{
    return zmarch_supported(kc) && kc[0].r == kc[1].r && kc[0].r >= 1 && !ctx->force_zm1;
}

// out <- G_z G_y G_x (a [+ b]) for `nplanes` component volumes, or (addout) out <- b + G_z G_y G_x a; a/b/out must not alias.
inline int launch_conv3d_zmarch(b200reg_ctx* ctx, const double* a, const double* b, double* out, int nx, int ny, int nz, int nplanes,
                                const KernelCoeffs kc[3], const DemonsCtrl* ctrl, int it, bool addout = false)
{
    SmallCoeffs sc;
    for (int ax = 0; ax < 3; ++ax) {
        sc.r[ax] = kc[ax].r;
        for (int t = 0; t <= 2 * kc[ax].r; ++t) sc.k[ax][t] = kc[ax].k[t];
    }
#ifdef B200REG_AB_VARIANTS
    const int txw = ctx->zm_tx32 ? 32 : ZM_TX;
#else
    const int txw = 32;
#endif
    const int tiles = ((nx + txw - 1) / txw) * ((ny + ZM_TY - 1) / ZM_TY) * nplanes;
    // enough CTAs for ~4 waves of 2 CTAs/SM, but chunks of at least 16 planes (each chunk re-reads 2*rz planes)
    // (coarse pyramid levels have so few tiles that chunks as short as 4 planes are worth their halo planes: the
    // whole field is L2 resident there and the GPU is otherwise idle)
    int nchunks = (ctx->sm_count * 8 + tiles - 1) / tiles;
    const int min_chunk = tiles * ((nz + 15) / 16) >= ctx->sm_count * 4 ? 16 : 4;
    const int max_chunks = (nz + min_chunk - 1) / min_chunk;
    {
        // Every CTA of this kernel takes the same time, so the launch runs in ceil(CTAs / resident CTAs) rounds of
        // (chunk + halo) plane steps: pick the chunk count with the least total.  Measured at 512 x 512 x 256: 3 chunks
        // (7.8 -> 8 rounds) instead of 2 (5.2 -> 6 rounds), 3.25 -> 3.20 ms / iteration; at 128 x 128 x 64 the model picks
        // 6 chunks = 288 CTAs, all resident in one round of 15 steps, instead of three rounds of 8.
        // resident CTAs per SM: 4 at 32-wide tiles (5 for the plain tensor-map pass: 72 registers), 2 at 64-wide tiles
        const bool tma5 = txw == 32 && !b && ctx->zm_tma != 0 && (nx % 2) == 0;
        const long slots = (long)ctx->sm_count * (txw == 32 ? (tma5 ? ZM_MODE0_CTAS : 4) : 2);
        long best_cost = -1;
        int best_k = nchunks;
        for (int k = 1; k <= 32 && k <= max_chunks; ++k) {
            const long rounds = ((long)tiles * k + slots - 1) / slots;
            const long cost = rounds * ((nz + k - 1) / k + 2 * kc[2].r);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best_k = k;
            }
        }
        nchunks = best_k;
        if (ctx->zm_chunks > 0) nchunks = ctx->zm_chunks;  // B200REG_ZM_CHUNKS: A/B override
    }
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    const int zchunk = (nz + nchunks - 1) / nchunks;
    nchunks = (nz + zchunk - 1) / zchunk;
    dim3 g((nx + ZM_TX - 1) / ZM_TX, (ny + ZM_TY - 1) / ZM_TY, nplanes * nchunks);
    if (kc[0].r == kc[1].r && kc[0].r >= 1 && !ctx->force_zm1) {
        int rc;
        switch (kc[0].r) {
        case 1: rc = launch_zm2_r<1>(ctx, kc[2].r, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout); break;
        case 2: rc = launch_zm2_r<2>(ctx, kc[2].r, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout); break;
        case 3: rc = launch_zm2_r<3>(ctx, kc[2].r, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout); break;
        default: rc = launch_zm2_r<4>(ctx, kc[2].r, a, b, out, nx, ny, nz, g, zchunk, nchunks, sc, ctrl, it, addout); break;
        }
        B200_TRY(rc);
        ctx->launches++;
        B200_CHECK_LAUNCH();
        return B200REG_OK;
    }
#define ZM_LAUNCH(RZ)                                                                                                                       \
    if (b) conv3d_zmarch_kernel<RZ, true><<<g, ZM_NT, 0, ctx->stream>>>(a, b, out, nx, ny, nz, zchunk, nchunks, sc, ctrl, it);              \
    else conv3d_zmarch_kernel<RZ, false><<<g, ZM_NT, 0, ctx->stream>>>(a, nullptr, out, nx, ny, nz, zchunk, nchunks, sc, ctrl, it)
    switch (kc[2].r) {
    case 1: ZM_LAUNCH(1); break;
    case 2: ZM_LAUNCH(2); break;
    case 3: ZM_LAUNCH(3); break;
    default: ZM_LAUNCH(4); break;
    }
#undef ZM_LAUNCH
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ---- shared-memory tiled separable pass for Float32 images (DiscreteGaussianImageFilter, any radius <= KMAX_R) ---
// Inputs are converted to double once when staged (the per-tap f32->f64 conversion of the naive kernel saturates
// the 16-lane XU pipe); taps are accumulated in double in ascending order, result rounded to float32.
// AXIS 0: one block = 256 consecutive outputs of each of 4 rows; 64 x 4 threads, every thread owns FOUR consecutive outputs of one row and
// slides a four-value register window along the taps (one shared load + one coefficient + four multiply-adds per tap, as in the y / z
// kernel below).  The staged row is stored 4-way interleaved -- element e at (e & 3) * CX_Q + (e >> 2) -- so that lanes, which are four
// elements apart, read consecutive words (conflict-free); CX_Q = 4 (mod 16) keeps the staging stores conflict-free as well.
constexpr int CX_Q = 100;  // >= (256 + 2 * KMAX_R) / 4, = 4 (mod 16)
// RT > 0: the radius as a compile-time constant (the host dispatches the radii pyramids actually use): the tap loop unrolls completely --
// no loop counter, no remainder loop, no register shuffling of the window.  ncu of the run-time-radius form at 512 x 512 x 256, radius 3:
// issue slots 74-89 % busy with only 22-29 % of them FP64 -- ~98 instructions per output of which 14 are the multiply-adds.
template <int RT>
__global__ void __launch_bounds__(256) conv_x_f32_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, int nx, int ny, int nz,
                                                                const __grid_constant__ KernelCoeffs kc)
{
    __shared__ double sm[4][4 * CX_Q];
    const int r = RT > 0 ? RT : kc.r;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * 256, y = blockIdx.y * 4 + ty, z = blockIdx.z;
    const bool row_ok = y < ny;
    const size_t row = ((size_t)z * ny + (row_ok ? y : ny - 1)) * nx;
    double* s = sm[ty];
    for (int e = tx; e < 256 + 2 * r; e += 64) {
        int gx = x0 - r + e;
        gx = gx < 0 ? 0 : (gx > nx - 1 ? nx - 1 : gx);
        s[(e & 3) * CX_Q + (e >> 2)] = (double)in[row + gx];
    }
    __syncthreads();
    const int x = x0 + 4 * tx;
    if (!row_ok || x >= nx) return;
    // output m of this thread uses elements 4 tx + m + t, t = 0 .. 2r, in ascending tap order
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    const double* w = s + tx;
    double v0 = w[0], v1 = w[CX_Q], v2 = w[2 * CX_Q], v3 = w[3 * CX_Q];
    const int last = 2 * r;
    if (RT > 0) {
#pragma unroll
        for (int t = 0; t <= 2 * RT; ++t) {
            const double kt = kc.k[t];
            acc[0] += kt * v0;
            acc[1] += kt * v1;
            acc[2] += kt * v2;
            acc[3] += kt * v3;
            v0 = v1;
            v1 = v2;
            v2 = v3;
            if (t < 2 * RT) v3 = w[((t + 4) & 3) * CX_Q + ((t + 4) >> 2)];
        }
    } else {
#pragma unroll 4
        for (int t = 0; t <= last; ++t) {
            const double kt = kc.k[t];
            acc[0] += kt * v0;
            acc[1] += kt * v1;
            acc[2] += kt * v2;
            acc[3] += kt * v3;
            v0 = v1;
            v1 = v2;
            v2 = v3;
            if (t < last) v3 = w[((t + 4) & 3) * CX_Q + ((t + 4) >> 2)];
        }
    }
    float* o = out + row + x;
    if (x + 3 < nx && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
        *reinterpret_cast<float4*>(o) = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
    } else {
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (x + m < nx) o[m] = (float)acc[m];
    }
}
// AXIS 1 / 2: tile of 32 x-columns by 32 positions along the axis; each thread owns 4 consecutive positions and
// slides over 4 + 2r staged values, feeding the four accumulators in ascending tap order.
template <int AXIS, int RT>
__global__ void __launch_bounds__(256) conv_yz_f32_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, int nx, int ny, int nz,
                                                                 const __grid_constant__ KernelCoeffs kc)
{
    __shared__ double sm[(32 + 2 * (RT > 0 ? RT : KMAX_R)) * 32];
    const int r = RT > 0 ? RT : kc.r;
    const int n = AXIS == 1 ? ny : nz;
    const size_t sa = AXIS == 1 ? (size_t)nx : (size_t)nx * ny;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int a0 = blockIdx.y * 32;             // first output position along the axis
    const int other = blockIdx.z;               // z for AXIS 1, y for AXIS 2
    const size_t base = AXIS == 1 ? (size_t)other * nx * ny : (size_t)other * nx;
    const int xc = x < nx ? x : nx - 1;
    // element offsets inside the volume fit 32 bits (the host sends larger volumes to the generic kernel): one base pointer, then a
    // 32-bit multiply and one wide add per element instead of a 64-bit multiply chain
    const float* __restrict__ pin = in + base + xc;
    const unsigned sa32 = (unsigned)sa;
    for (int e = ty; e < 32 + 2 * r; e += 8) {
        int q = a0 - r + e;
        q = q < 0 ? 0 : (q > n - 1 ? n - 1 : q);
        sm[e * 32 + lane] = (double)pin[(unsigned)q * sa32];
    }
    __syncthreads();
    // four consecutive outputs per thread over a sliding window of four staged values: per tap one shared load and one coefficient,
    // four multiply-adds, no predicates (the earlier form tested every (value, output) pair); same ascending tap order per output
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    const double* col = sm + (ty * 4) * 32 + lane;
    double v0 = col[0], v1 = col[32], v2 = col[64], v3 = col[96];
    const int last = 2 * r;
    if (RT > 0) {
#pragma unroll
        for (int t = 0; t <= 2 * RT; ++t) {
            const double kt = kc.k[t];
            acc[0] += kt * v0;
            acc[1] += kt * v1;
            acc[2] += kt * v2;
            acc[3] += kt * v3;
            v0 = v1;
            v1 = v2;
            v2 = v3;
            if (t < 2 * RT) v3 = col[(t + 4) * 32];
        }
    } else {
#pragma unroll 4
        for (int t = 0; t <= last; ++t) {
            const double kt = kc.k[t];
            acc[0] += kt * v0;
            acc[1] += kt * v1;
            acc[2] += kt * v2;
            acc[3] += kt * v3;
            v0 = v1;
            v1 = v2;
            v2 = v3;
            if (t < last) v3 = col[(t + 4) * 32];
        }
    }
    if (x < nx) {
        float* __restrict__ pout = out + base + x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = a0 + ty * 4 + j;
            if (q < n) pout[(unsigned)q * sa32] = (float)acc[j];
        }
    }
}
template <int RT>
inline void launch_conv_axis_f32_tiled_r(b200reg_ctx* ctx, int axis, const float* in, float* out, int nx, int ny, int nz, const KernelCoeffs& kc)
{
    if (axis == 0) {
        conv_x_f32_tiled_kernel<RT><<<dim3((nx + 255) / 256, (ny + 3) / 4, nz), dim3(64, 4), 0, ctx->stream>>>(in, out, nx, ny, nz, kc);
    } else if (axis == 1) {
        conv_yz_f32_tiled_kernel<1, RT><<<dim3((nx + 31) / 32, (ny + 31) / 32, nz), 256, 0, ctx->stream>>>(in, out, nx, ny, nz, kc);
    } else {
        conv_yz_f32_tiled_kernel<2, RT><<<dim3((nx + 31) / 32, (nz + 31) / 32, ny), 256, 0, ctx->stream>>>(in, out, nx, ny, nz, kc);
    }
}
inline int launch_conv_axis_f32_tiled(b200reg_ctx* ctx, int axis, const float* in, float* out, int nx, int ny, int nz, const KernelCoeffs& kc)
{
    // compile-time radii for what the pyramids and the fusion blur use (sigma 0.5 ... 4 voxels); any other radius takes the run-time form
    switch (ctx->conv_static_radius ? kc.r : 0) {
    case 1: launch_conv_axis_f32_tiled_r<1>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 2: launch_conv_axis_f32_tiled_r<2>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 3: launch_conv_axis_f32_tiled_r<3>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 4: launch_conv_axis_f32_tiled_r<4>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 5: launch_conv_axis_f32_tiled_r<5>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 6: launch_conv_axis_f32_tiled_r<6>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 7: launch_conv_axis_f32_tiled_r<7>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 8: launch_conv_axis_f32_tiled_r<8>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 10: launch_conv_axis_f32_tiled_r<10>(ctx, axis, in, out, nx, ny, nz, kc); break;
    case 12: launch_conv_axis_f32_tiled_r<12>(ctx, axis, in, out, nx, ny, nz, kc); break;
    default: launch_conv_axis_f32_tiled_r<0>(ctx, axis, in, out, nx, ny, nz, kc); break;
    }
    ctx->launches++;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// DiscreteGaussianImageFilter on a Float32 image: variance (mm^2) -> voxel^2 per axis when
// use_image_spacing, passes z -> y -> x, float32 intermediates.
inline int discrete_gaussian_f32(b200reg_ctx* ctx, const float* d_in, float* d_out, const b200reg_geom& g, const double* variance,
                                 int max_width, double max_error, int use_spacing)
{
    const int nx = g.size[0], ny = g.size[1], nz = g.size[2];
    const size_t n = nvox(g);
    TempBuf t1, t2;
    B200_TRY(t1.alloc(ctx, n * sizeof(float)));
    B200_TRY(t2.alloc(ctx, n * sizeof(float)));
    const float* src = d_in;
    float* dsts[3] = { t1.as<float>(), t2.as<float>(), d_out };
    for (int pass = 0; pass < 3; ++pass) {
        const int axis = semantics().discrete_gaussian_axis_order ? pass : 2 - pass;  // z, y, x unless the switch says x, y, z
        double t = variance[axis];
        if (use_spacing) t = t / (g.spacing[axis] * g.spacing[axis]);
        KernelCoeffs kc;
        B200_TRY(make_coeffs(gaussian_operator(t, max_error, max_width), &kc));
        if (nz <= 65535 && ny <= 65535 && n < (1ull << 32)) B200_TRY(launch_conv_axis_f32_tiled(ctx, axis, src, dsts[pass], nx, ny, nz, kc));
        else B200_TRY((launch_conv_axis<float, false>(ctx, axis, src, nullptr, dsts[pass], nx, ny, nz, 1, kc, nullptr, 0)));
        src = dsts[pass];
    }
    return B200REG_OK;
}

}  // namespace b200
