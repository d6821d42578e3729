// Microbenchmark: per-SM per-clock throughput of DADD / DMUL / DFMA / F2F.F64.F32 on sm_100a (evidence for DESIGN.md).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, float* fin, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    float f0 = fin[threadIdx.x], f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c); a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c); }
        if (OP == 1) { a0 = __dmul_rn(a0, b); a1 = __dmul_rn(a1, b); a2 = __dmul_rn(a2, b); a3 = __dmul_rn(a3, b); a4 = __dmul_rn(a4, b); a5 = __dmul_rn(a5, b); a6 = __dmul_rn(a6, b); a7 = __dmul_rn(a7, b); }
        if (OP == 2) { a0 = __fma_rn(a0, b, c); a1 = __fma_rn(a1, b, c); a2 = __fma_rn(a2, b, c); a3 = __fma_rn(a3, b, c); a4 = __fma_rn(a4, b, c); a5 = __fma_rn(a5, b, c); a6 = __fma_rn(a6, b, c); a7 = __fma_rn(a7, b, c); }
        if (OP == 3) { a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3; f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, iters = 20000, threads = 1024;
    double* out; float* fin; cudaMalloc(&out, sizeof(double) * sms * threads * 2); cudaMalloc(&fin, 4096 * 4); cudaMemset(fin, 0, 4096 * 4);
    const char* names[4] = { "DADD", "DMUL", "DFMA", "F2F.F64.F32 (+DADD)" };
    for (int op = 0; op < 4; ++op) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k<0><<<sms * 2, threads>>>(out, fin, iters);
            if (op == 1) k<1><<<sms * 2, threads>>>(out, fin, iters);
            if (op == 2) k<2><<<sms * 2, threads>>>(out, fin, iters);
            if (op == 3) k<3><<<sms * 2, threads>>>(out, fin, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * 2 * threads * iters * (op == 3 ? 4 : 8);
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("%-22s %8.2f Gop/s  = %6.1f ops/clk/SM at %d MHz (max clock)\n", names[op], ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    return 0;
}
