// Dependent-issue latencies on sm_100a (developer microbenchmark): one warp, one chain of N dependent ops.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(double *out, double a, double b, int *iout)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    int q = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, y);
        if (OP == 1) x = x * y;
        if (OP == 2) x = x + y;
        if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
        if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1, 4);
        if (OP == 5) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
        if (OP == 6) x = (x > y) ? x : y + x;   // DSETP + FSEL x2 (+DADD)
        if (OP == 7) q = __shfl_xor_sync(0xffffffffu, q, 1, 4);
        if (OP == 8) x = (double)(__double2hiint(x) + i);     // I2F.F64
        if (OP == 9) { x = fma(x, y, y); y = fma(y, a, b); }  // two independent chains
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[OP] = (double)(t1 - t0) / N; }
    out[16 + threadIdx.x] = x + y; iout[threadIdx.x] = q;
}
int main()
{
    double *d; int *di; cudaMalloc(&d, 1024); cudaMalloc(&di, 1024);
    const char *names[] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H", "SHFL f64 (2xSHFL)", "MUFU.RSQ64H", "DSETP+FSEL(+DADD)", "SHFL b32", "I2F.F64+IADD", "2x DFMA indep"};
    k<0><<<1, 32>>>(d, 1.0000001, 0.5, di); k<1><<<1, 32>>>(d, 1.0000001, 0.999999, di); k<2><<<1, 32>>>(d, 1.0, 0.5, di);
    k<3><<<1, 32>>>(d, 1.5, 0.5, di); k<4><<<1, 32>>>(d, 1.5, 0.5, di); k<5><<<1, 32>>>(d, 1.5, 0.5, di);
    k<6><<<1, 32>>>(d, 1.5, 0.5, di); k<7><<<1, 32>>>(d, 1.5, 0.5, di); k<8><<<1, 32>>>(d, 1.5, 0.5, di); k<9><<<1, 32>>>(d, 1.0000001, 0.5, di);
    cudaDeviceSynchronize();
    double h[16]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 10; ++i) printf("%-22s %.2f cycles per dependent step\n", names[i], h[i]);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
