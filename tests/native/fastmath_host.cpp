// Host build of pyhillfit_b200/csrc/phf_fastmath.cuh for the accuracy tests (tests/test_fastmath.py).
// Compiled with g++ (no CUDA); the MUFU seeds are emulated by 20-bit reciprocals, everything else is the
// same code the kernels run.
#include "../../pyhillfit_b200/csrc/phf_fastmath.cuh"

using namespace phf::fm;

extern "C" {
void fmh_exp(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = exp_clamped(kFmTable, x[i]); }
void fmh_log(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = log_pos(kFmTable, x[i]); }
void fmh_rcp(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = rcp(x[i]); }
void fmh_rsqrt(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = rsqrt(x[i]); }
void fmh_sqrt(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = sqrt_nonneg(x[i]); }
void fmh_erfcx(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = erfcx_nonneg(kFmTable, x[i]); }
void fmh_erfcx_pw(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = erfcx_nonneg_pw(kFmTable, x[i]); }
void fmh_log_ndtr_pw(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = log_ndtr_nonpos_pw(kFmTable, x[i]); }
void fmh_log_ndtr(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = log_ndtr_nonpos(kFmTable, x[i]); }
void fmh_exp10(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = exp10_clamped(kFmTable, x[i]); }
void fmh_sincos(int n, const uint32_t *b, double *s, double *c) { for (int i = 0; i < n; ++i) sincos_turn32(kFmTable, b[i], s[i], c[i]); }
}
