// Library-level utilities: version, error text, launch counter, FP64 peak probe.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "phf_common.cuh"

namespace phf {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

int set_cuda_error(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof g_err, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return PHF_ECUDA;
}

int check_launch(const char *kernel_name)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, kernel_name);
    return PHF_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// 8 independent DFMA chains per thread, no memory traffic: measures the FP64 FMA issue ceiling.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // never true; keeps the chains live
}

}  // namespace phf

using namespace phf;

extern "C" int phf_version(void) { return PHF_VERSION; }
extern "C" const char *phf_last_error(void) { return g_err; }
extern "C" int64_t phf_launch_count(void) { return g_launches.load(); }

extern "C" int phf_fp64_peak_probe(int32_t repeats, double *tflops_out, double *seconds_out)
{
    if (!tflops_out) return set_error(PHF_EINVAL, "phf_fp64_peak_probe: tflops_out is NULL");
    if (repeats < 1) repeats = 1;
    const int iters = 4096, block = 256, grid = sm_count() * 8;
    double *d = nullptr;
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaMalloc(&d, sizeof(double)))) return set_cuda_error(e, "cudaMalloc");
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < repeats + 1; ++r) {  // first pass is warm-up
        cudaEventRecord(e0, 0);
        fp64_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
        count_launch();
        cudaEventRecord(e1, 0);
        if ((e = cudaEventSynchronize(e1))) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e) return set_cuda_error(e, "fp64_peak_kernel");
    const double flops = 2.0 * 64.0 * (double)iters * (double)block * (double)grid;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (seconds_out) *seconds_out = best * 1e-3;
    return PHF_OK;
}
