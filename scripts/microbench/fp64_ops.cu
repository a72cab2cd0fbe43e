// Throughput of the FP64 instructions the samplers use, per warp-instruction per sub-partition (developer
// microbenchmark; B200).  384 threads per SM (3 warps per sub-partition), 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_ops fp64_ops.cu && ./fp64_ops
#include <cstdio>
#include <cuda_runtime.h>

enum Op { FMA, MUL, ADD, SETP_SEL, MIX, FMA_CONST, MUFU_RCP, I2F, LDC };

__constant__ double ctab[64];

template <int OP>
__global__ void __launch_bounds__(384) k(double *out, int iters, double a, double b)
{
    double f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (OP == FMA) f[j] = fma(f[j], a, b);
                if (OP == MUL) f[j] = f[j] * a;
                if (OP == ADD) f[j] = f[j] + b;
                if (OP == SETP_SEL) f[j] = f[j] > b ? f[j] : a;   // DSETP + 2 FSEL... measured as a unit
                if (OP == MIX) f[j] = (j & 1) ? f[j] * a : fma(f[j], a, b);
                if (OP == FMA_CONST) f[j] = fma(f[j], ctab[(r * 8 + j) & 63], b);
                if (OP == MUFU_RCP) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(f[j])); f[j] = y; }
                if (OP == I2F) f[j] = (double)(int)__double2loint(f[j]) ;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
    if (s == 123.456) out[0] = s;
}

template <int OP>
void run(const char *name, int sms)
{
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<sms, 384>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<sms, 384>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3 / ((double)iters * 32) / 3.0;  // per op per warp
    printf("%-10s %.2f cycles per warp-op per sub-partition\n", name, cyc);
    cudaFree(out);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<FMA>("DFMA", sms);
    run<MUL>("DMUL", sms);
    run<ADD>("DADD", sms);
    run<SETP_SEL>("DSETP+SEL", sms);
    run<MIX>("DFMA/DMUL", sms);
    run<FMA_CONST>("DFMA c[]", sms);
    run<MUFU_RCP>("MUFU.RCP64", sms);
    run<I2F>("I2F.F64", sms);
    return 0;
}
