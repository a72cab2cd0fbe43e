// Does an FP64 warp-instruction cost one issue slot or two?  (developer microbenchmark; B200)
// Each thread runs NF independent DFMA chains and NI independent integer IMAD chains per loop iteration;
// 384 threads per SM (3 warps per sub-partition, the samplers' occupancy), grid = #SMs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI>
__global__ void __launch_bounds__(384) mix(double *out, int iters, double a, double b, unsigned m)
{
    double f[NF > 0 ? NF : 1];
    unsigned x[NI > 0 ? NI : 1];
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = threadIdx.x + k;
#pragma unroll
    for (int k = 0; k < NI; ++k) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k] = fma(f[k], a, b);
#pragma unroll
            for (int k = 0; k < NI; ++k) x[k] = x[k] * m + 12345u;
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += f[k];
#pragma unroll
    for (int k = 0; k < NI; ++k) s += x[k];
    if (s == 123.456) out[0] = s;
}

template <int NF, int NI>
void run(int sms, int threads)
{
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mix<NF, NI><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, 1664525u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mix<NF, NI><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, 1664525u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3 / ((double)iters * 8);
    const int warps_per_smsp = threads / 128;
    printf("threads %3d  NF=%d NI=%d: %.2f cycles per round per SMSP (%d warps/SMSP) -> %.2f per warp-round; issue-only model %d, fp64-two-slots model %d\n",
           threads, NF, NI, cycles, warps_per_smsp, cycles / warps_per_smsp, (NF > NI ? 2 * NF : NF + NI), 2 * NF + NI);
    cudaFree(out);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {128, 384}) {
        run<4, 0>(sms, threads);
        run<4, 2>(sms, threads);
        run<4, 4>(sms, threads);
        run<4, 8>(sms, threads);
        run<2, 4>(sms, threads);
        run<2, 8>(sms, threads);
        run<0, 8>(sms, threads);
        run<8, 0>(sms, threads);
        run<8, 8>(sms, threads);
    }
    return 0;
}
