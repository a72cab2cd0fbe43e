// Error reporting / launch bookkeeping shared by the translation units of libphf_b200.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/pyhillfit_b200.h"

namespace phf {

int set_error(int code, const char *msg);
int set_cuda_error(cudaError_t e, const char *what);
int check_launch(const char *kernel_name);  // cudaGetLastError() -> PHF_OK / PHF_ECUDA
void count_launch();
int sm_count();

// CTA size for the sampler kernels (`n_threads` = chains x lanes).  Tiny launches use 32-thread CTAs so that every
// warp gets an SM to itself; up to ~16 warps per SM 64-thread CTAs (measured 4 % faster than 32 or 128 when the
// model-1 and model-2 launches of config 2 share the SMs: two warps of a CTA land on two sub-partitions and the
// grid still spreads evenly); beyond that 128.
inline int default_block_threads(int64_t n_threads)
{
    const int64_t sms = sm_count();
    if (n_threads <= sms * 32) return 32;
    if (n_threads <= sms * 16 * 64) return 64;
    return 128;
}

// First saved row a call writes to `samples` (row index = t / thinning) and the number it writes: every saved row of
// the call, or with cfg.discard_burn_rows only those with index >= burn_rows (python/PyHillFit.py:861-864,
// python/PyHillTemp.py:125 drop the others before saving).
__host__ __device__ inline uint32_t first_row_written(const phf_am_config &c)
{
    const uint32_t first = c.t0 / c.thinning + 1;
    return (c.discard_burn_rows && c.burn_rows != 0xFFFFFFFFu && c.burn_rows > first) ? c.burn_rows : first;
}
__host__ __device__ inline uint32_t rows_written(const phf_am_config &c)
{
    const uint32_t last = (c.t0 + c.n_iters) / c.thinning, from = first_row_written(c);
    return last >= from ? last - from + 1 : 0;
}

}  // namespace phf
