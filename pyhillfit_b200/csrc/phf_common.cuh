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

// One chain per thread: prefer 32-thread CTAs until every SM holds >= 16 of them, so that small chain
// counts spread over all 148 SMs; larger CTAs only once the grid is many waves deep.
inline int default_block_threads(int64_t n_chains)
{
    const int64_t sms = sm_count();
    if (n_chains <= sms * 16 * 32) return 32;
    if (n_chains <= sms * 16 * 64) return 64;
    return 128;
}

}  // namespace phf
