// fp64 elementary functions written for the fused samplers' inner loops.
//
// Why not the CUDA math library: one adaptive-Metropolis iteration is a chain of ~15 transcendental calls, and
// on sm_100a the library versions spend about two thirds of their issue slots outside the FP64 pipe -- every
// 64-bit polynomial coefficient is materialised by two UMOV instructions, every divide / sqrt carries a
// branch to a slow path, and each function re-derives special cases (NaN, denormals, negative arguments) that
// cannot occur here.  These versions
//   * take their coefficients from one 480-byte table staged in SHARED memory (one LDS.128 brings two of them;
//     ptxas turns __constant__ reads into one LDC.64 per coefficient and parks them in vector registers),
//   * are branch-free and evaluated with Estrin's scheme (short dependent chains),
//   * use the MUFU seeds (rcp.approx / rsqrt.approx) with FMA Newton steps instead of IEEE division / sqrt,
//   * state their domain instead of handling everything.
// Accuracy: every function is within a few ulp on its domain (tests/test_fastmath.py measures it on the host
// build of this same header against mpmath / libm); the log-target built from them stays within the 1e-12
// parity bound with two orders of magnitude to spare.
//
// The header also compiles as plain C++ (g++) for the host accuracy tests; there the MUFU seeds are emulated by
// a 20-bit reciprocal.  Coefficients: scripts/gen_fastmath_coeffs.py -> phf_fastmath_coeffs.inc.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

// where the kernels read the coefficients from: 0 = __constant__ memory, 1 = a shared-memory copy (LDS.128
// pairs), 2 = literals (constexpr table folded into immediates)
#ifndef PHF_FM_TABLE_MODE
#define PHF_FM_TABLE_MODE 0
#endif
#if defined(__CUDACC__)
#define PHF_FM __device__ __forceinline__
#if PHF_FM_TABLE_MODE == 2
#define PHF_COEFF_TABLE static __device__ const
#else
#define PHF_COEFF_TABLE static __constant__ __align__(16)
#endif
#else
#define PHF_FM static inline
#define PHF_COEFF_TABLE alignas(16) static const
struct double2 {
    double x, y;
};
#endif

// exp / log: 1 = table-driven (a 3 KB lookup table staged in SHARED memory by PHF_STAGE_FASTMATH_TABLE: 128 (1/c, log c)
// pairs and 64 double-double 2^(j/64); the polynomials shrink to degree 4 / 3, 11 fp64 instructions per call instead
// of 22 / 18), 0 = the polynomial-only versions.  With the table T passed to every function is the shared-memory
// table and the polynomial coefficients come straight from __constant__ memory.
#ifndef PHF_FM_LUT
#define PHF_FM_LUT 1
#endif
#if PHF_FM_LUT && PHF_FM_TABLE_MODE != 0
#error "PHF_FM_LUT needs PHF_FM_TABLE_MODE == 0 (coefficients in __constant__ memory)"
#endif
#if defined(__CUDACC__)
#define PHF_LUT_POLY static __constant__ __align__(16)
#define PHF_LUT_TABLE static __device__ __align__(16) const
#else
#define PHF_LUT_POLY alignas(16) static const
#define PHF_LUT_TABLE alignas(16) static const
#endif

namespace phf {
namespace fm {

#include "phf_fastmath_coeffs.inc"
#include "phf_fastmath_lut.inc"

// where a function finds the polynomial coefficients / the lookup table given the pointer its caller passes
PHF_FM const double *coef(const double *T)
{
#if defined(__CUDA_ARCH__) && PHF_FM_LUT
    (void)T;
    return kFmTable;  // T is the shared-memory lookup table
#else
    return T;
#endif
}
PHF_FM const double *lut(const double *T)
{
#if defined(__CUDA_ARCH__)
    return T;
#else
    (void)T;
    return kFmLut;  // host build: T is kFmTable
#endif
}

// Every function below takes `T`, the base of the coefficient table: shared memory in the kernels (see
// stage_table), kFmTable itself in the host build.
#if defined(__CUDACC__) && PHF_FM_TABLE_MODE == 1
// Copy the table into shared memory; call from all threads of the CTA, then __syncthreads().
__device__ __forceinline__ void stage_table(double *smem_table)
{
    for (int i = threadIdx.x; i < PHF_FM_TABLE_SIZE; i += blockDim.x) smem_table[i] = kFmTable[i];
}
#endif

// ---- bit access -------------------------------------------------------------------------------
PHF_FM int hi_word(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return (int)(u >> 32);
#endif
}
PHF_FM int lo_word(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return (int)(u & 0xffffffffu);
#endif
}
PHF_FM double make_double(int hi, int lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    std::memcpy(&x, &u, 8);
    return x;
#endif
}

// ---- MUFU seeds: ~20 good bits, low word of the result is zero ----------------------------------
PHF_FM double rcp_seed(double a)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
#else
    const double t = 1.0 / make_double(hi_word(a), 0);
    return make_double(hi_word(t), 0);
#endif
}
PHF_FM double rsqrt_seed(double a)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
#else
    const double t = 1.0 / std::sqrt(make_double(hi_word(a), 0));
    return make_double(hi_word(t), 0);
#endif
}

// ---- Estrin evaluation of sum_{k=LO..HI} c[k] x^(k-LO); xp[j] = x^(2^j) --------------------------
template <int N>
struct FloorLog2 {
    static constexpr int value = 1 + FloorLog2<(N >> 1)>::value;
};
template <>
struct FloorLog2<1> {
    static constexpr int value = 0;
};

template <int LO, int HI>
PHF_FM double poly(const double *c, const double *xp)
{
    if constexpr (LO == HI) {
        return c[LO];
    } else if constexpr (HI == LO + 1) {
        static_assert((LO & 1) == 0, "coefficient pairs are 16-byte aligned");
#if PHF_FM_TABLE_MODE == 1
        const double2 p = *reinterpret_cast<const double2 *>(c + LO);  // one 128-bit load brings both
        return fma(p.y, xp[0], p.x);
#else
        return fma(c[HI], xp[0], c[LO]);
#endif
    } else {
        constexpr int j = FloorLog2<HI - LO>::value;  // 2^j = largest power of two <= HI-LO, i.e. < number of terms
        constexpr int half = 1 << j;
        return fma(poly<LO + half, HI>(c, xp), xp[j], poly<LO, LO + half - 1>(c, xp));
    }
}

// ---- 1/a, a finite and normal (|a| in [1e-290, 1e290]); <= 1 ulp ---------------------------------
// seed (relative error e0 ~ 2^-20) -> y0 (1 + e0 + e0^2) = (1 - e0^3)/a: 2^-60 before the final rounding.  (A further
// correction step, two more instructions, only turns "within 1 ulp" into "almost always correctly rounded".)
#ifndef PHF_FM_RCP_STEPS
#define PHF_FM_RCP_STEPS 1
#endif
PHF_FM double rcp(double a)
{
    double y = rcp_seed(a);
    double e = fma(-a, y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
#if PHF_FM_RCP_STEPS > 1
    e = fma(-a, y, 1.0);
    y = fma(y, e, y);
#endif
    return y;
}

// ---- 1/sqrt(a), a > 0 normal --------------------------------------------------------------------
// One third-order step from the seed y0 = (1 + d)/sqrt(a), |d| ~ 2^-20: e = (1 - a y0^2)/2 = -d - d^2/2 and
// y0 (1 + e + 3/2 e^2) = (1 + O(d^3))/sqrt(a).  Five instructions with a dependent chain of four (two second-order
// steps: six and six) -- the three rsqrt of the proposal's Cholesky factor are the head of every iteration's chain.
#ifndef PHF_FM_CUBIC_ROOTS
#define PHF_FM_CUBIC_ROOTS 1
#endif
PHF_FM double rsqrt(double a)
{
    double y = rsqrt_seed(a);
    const double h = 0.5 * a;
#if PHF_FM_CUBIC_ROOTS
    const double e = fma(-(h * y), y, 0.5);
    return fma(y * e, fma(1.5, e, 1.0), y);
#else
    double e = fma(-(h * y), y, 0.5);
    y = fma(y, e, y);
    e = fma(-(h * y), y, 0.5);
    return fma(y, e, y);
#endif
}

// ---- sqrt(a), a >= 0 (a == 0 -> 0) --------------------------------------------------------------
PHF_FM double sqrt_nonneg(double a)
{
    const double y = rsqrt_seed(fmax(a, 1e-290));
    double g = a * y, h = 0.5 * y;
#if PHF_FM_CUBIC_ROOTS
    const double r = fma(-h, g, 0.5);            // (1 - a y^2)/2
    return fma(g * r, fma(1.5, r, 1.0), g);      // g (1 + r + 3/2 r^2)
#else
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    return fma(g, r, g);
#endif
}

// ---- exp(x); the argument is clamped to [-700, 700] (the callers' results saturate long before) -----
#if PHF_FM_LUT
// e^x = 2^(n >> 6) * 2^((n & 63)/64) * e^r, n = round(64 x / ln 2), |r| <= 0.0058: 11 fp64 instructions
PHF_FM double exp_clamped(const double *T, double x)
{
    const double k64Log2eHi = 64.0 * 1.4426946640014648;         // 64 log2(e) to 21 bits (picks n only)
    const double kLn2Hi64 = 0.693147182464599609375 / 64.0;      // n * kLn2Hi64 is exact (21 significant bits)
    const double kLn2Lo64 = coef(T)[PHF_FM_KMISC + 0] * (1.0 / 64.0);
    const double kMagic = 6755399441055744.0;                    // 1.5 * 2^52
    {
        const int hi = hi_word(x);
        const bool big = (hi & 0x7fffffff) >= 0x4085e000;  // 0x4085e000'00000000 == 700.0; NaN too
        x = make_double(big ? ((hi & 0x80000000) | 0x4085e000) : hi, big ? 0 : lo_word(x));
    }
    const double t = fma(x, k64Log2eHi, kMagic);
    const int n = lo_word(t);
    const double nf = t - kMagic;
    double r = fma(nf, -kLn2Hi64, x);
    r = fma(nf, -kLn2Lo64, r);
    const double2 tj = *reinterpret_cast<const double2 *>(lut(T) + PHF_FM_LUT_EXP + 2 * (n & 63));  // 2^(j/64): hi, lo
    const double r2 = r * r;
    const double q = fma(fma(kFmExpQ[3], r, kFmExpQ[2]), r2, fma(kFmExpQ[1], r, kFmExpQ[0]));
    const double p = fma(r2, q, r);                    // e^r - 1
    const double v = fma(tj.x, p, tj.y) + tj.x;        // in [1, 2) up to rounding
    return make_double(hi_word(v) + ((n >> 6) << 20), lo_word(v));  // * 2^(n >> 6), |n >> 6| <= 1010
}
#else
PHF_FM double exp_clamped(const double *T, double x)
{
    const double kLog2eHi = 1.4426946640014648;       // 0x3ff7154700000000: log2(e) to 21 bits (picks n only)
    const double kLn2Hi = 0.693147182464599609375;    // 0x3fe62e4300000000: n * kLn2Hi is exact
    const double kLn2Lo = coef(T)[PHF_FM_KMISC + 0];         // ln 2 - kLn2Hi
    const double kMagic = 6755399441055744.0;         // 1.5 * 2^52
    // clamp on the high word (5 integer instructions; fmin/fmax cost 12 on sm_100a): |x| >= 700 or NaN -> +-700
    {
        const int hi = hi_word(x);
        const bool big = (hi & 0x7fffffff) >= 0x4085e000;  // 0x4085e000'00000000 == 700.0
        x = make_double(big ? ((hi & 0x80000000) | 0x4085e000) : hi, big ? 0 : lo_word(x));
    }
    const double t = fma(x, kLog2eHi, kMagic);
    const int n = lo_word(t);
    const double nf = t - kMagic;
    double r = fma(nf, -kLn2Hi, x);
    r = fma(nf, -kLn2Lo, r);
    double xp[4];
    xp[0] = r;
    xp[1] = r * r;
    xp[2] = xp[1] * xp[1];
    xp[3] = xp[2] * xp[2];
    const double q = poly<0, 9>(coef(T) + PHF_FM_KEXPQ, xp);
    const double p = fma(xp[1], q, r);                 // e^r - 1
    const double s = make_double((n + 1023) << 20, 0);  // 2^n, n in [-1010, 1010]
    return fma(s, p, s);
}
#endif

// ---- log(x), x positive and normal --------------------------------------------------------------
#if PHF_FM_LUT
// x = 2^e m, m in [sqrt(1/2), sqrt 2); interval j of m (top 7 bits of its high-word offset) has (1/c_j, log c_j) in the
// table: r = m / c_j - 1 by one FMA (|r| < 2^-8), log m = log c_j + r + r^2 P(r).  11 fp64 instructions, no MUFU.
PHF_FM double log_pos(const double *T, double x)
{
    const double kLn2Hi = 0.693147182464599609375;
    const double kLn2Lo = coef(T)[PHF_FM_KMISC + 0];
    const int k = hi_word(x) + (0x3ff00000 - 0x3fe6a09e);  // 0x3fe6a09e: high word of sqrt(1/2)
    const int e = (k >> 20) - 1023;
    const int off = k & 0x000fffff;
    const double m = make_double(off + 0x3fe6a09e, lo_word(x));
    const double2 cj = *reinterpret_cast<const double2 *>(lut(T) + 2 * (off >> (20 - PHF_FM_LUT_LOG_BITS)));
    const double r = fma(m, cj.x, -1.0);
    const double r2 = r * r;
    const double pr = fma(r2 * r2, kFmLogP[4], fma(fma(kFmLogP[3], r, kFmLogP[2]), r2, fma(kFmLogP[1], r, kFmLogP[0])));
    const double l1p = fma(r2, pr, r);  // log(1 + r)
    const double ef = (double)e;
    return fma(ef, kLn2Hi, cj.y) + fma(ef, kLn2Lo, l1p);
}
#else
PHF_FM double log_pos(const double *T, double x)
{
    const double kLn2Hi = 0.693147182464599609375;
    const double kLn2Lo = coef(T)[PHF_FM_KMISC + 0];
    // mantissa to [sqrt(1/2), sqrt(2)): 0x3fe6a09e is the high word of sqrt(1/2)
    const int k = hi_word(x) + (0x3ff00000 - 0x3fe6a09e);
    const int e = (k >> 20) - 1023;
    const double m = make_double((k & 0x000fffff) + 0x3fe6a09e, lo_word(x));
    const double num = m - 1.0, den = m + 1.0;
    // f = num / den
    double y = rcp_seed(den);
    double d = fma(-den, y, 1.0);
    d = fma(d, d, d);
    y = fma(y, d, y);
    double f = num * y;
    f = fma(fma(-f, den, num), y, f);
    double xp[3];
    xp[0] = f * f;
    xp[1] = xp[0] * xp[0];
    xp[2] = xp[1] * xp[1];
    const double R = poly<0, 6>(coef(T) + PHF_FM_KLOGR, xp);
    const double ef = (double)e;
    // log x = e ln2 + 2 atanh(f) = e ln2_hi + (2f + (f^3 R + e ln2_lo))
    const double tail = fma(f * xp[0], R, ef * kLn2Lo);
    return fma(ef, kLn2Hi, fma(2.0, f, tail));
}
#endif

// ---- erfcx(t) = exp(t^2) erfc(t), 0 <= t <= 1e140 --------------------------------------------------
// erfcx(t) (1 + 2t) = P(q), q = (t - K)/(t + K): one polynomial on q in [-1, 1) covers the half line.
// (Every caller passes t = |response - bound| / (sigma sqrt 2) with sigma > 1e-3: t < 1e5.  Beyond 1e140 the
// product a b below overflows; a clamp would cost 8 instructions per call.)
PHF_FM double erfcx_nonneg(const double *T, double t)
{
    const double a = t + PHF_ERFCX_K, b = fma(2.0, t, 1.0);
    const double r = rcp(a * b);
    const double q = (t - PHF_ERFCX_K) * (r * b);
    double xp[5];
    xp[0] = q;
    xp[1] = q * q;
    xp[2] = xp[1] * xp[1];
    xp[3] = xp[2] * xp[2];
    xp[4] = xp[3] * xp[3];
    const double P = poly<0, 22>(coef(T) + PHF_FM_KERFCXP, xp);
    return P * (r * a);
}

#if PHF_FM_LUT
// Table-driven form (erfcx_nonneg_pw): the same map t -> q, then one of 32 degree-7 polynomials (interval j of q, local variable
// v = 16 q + 15.5 - j in [-1/2, 1/2]; coefficients in the shared-memory table): 9 fp64 instructions for the polynomial
// instead of 26.  (q reaches 1 only for t >= 2^55, where the last interval's polynomial is then evaluated at the wrong
// end: 0.4 % off, far outside every caller's range -- see above.)
PHF_FM double erfcx_nonneg_pw(const double *T, double t)
{
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    const double a = t + PHF_ERFCX_K, b = fma(2.0, t, 1.0);
    const double r = rcp(a * b);
    const double q = (t - PHF_ERFCX_K) * (r * b);
    const double s = fma(q, 0.5 * PHF_FM_LUT_ERFCX_N, 0.5 * PHF_FM_LUT_ERFCX_N - 0.5);
    const double tt = s + kMagic;  // round to nearest: interval index in the low word
    int j = lo_word(tt);
    // (one unsigned compare also catches the garbage index of an out-of-domain t < 0 or NaN -- the samplers evaluate
    // proposals with a negative sigma before they reject them -- so the table read stays inside the table)
    j = (unsigned)j < (unsigned)PHF_FM_LUT_ERFCX_N ? j : PHF_FM_LUT_ERFCX_N - 1;
    const double v = s - (tt - kMagic);
    const double2 *c = reinterpret_cast<const double2 *>(lut(T) + PHF_FM_LUT_ERFCX + 8 * j);
    const double2 c01 = c[0], c23 = c[1], c45 = c[2], c67 = c[3];
    const double v2 = v * v, v4 = v2 * v2;
    const double lo = fma(fma(c23.y, v, c23.x), v2, fma(c01.y, v, c01.x));
    const double hi = fma(fma(c67.y, v, c67.x), v2, fma(c45.y, v, c45.x));
    return fma(hi, v4, lo) * (r * a);
}
#else
PHF_FM double erfcx_nonneg_pw(const double *T, double t) { return erfcx_nonneg(T, t); }
#endif

// ---- log Phi(z), z <= 0: log(erfcx(|z|/sqrt2)/2) - z^2/2 (scipy.special.log_ndtr's z < -1 branch, accurate
//      on all of z <= 0 because the value never comes near zero there) -------------------------------
PHF_FM double log_ndtr_nonpos(const double *T, double z)
{
    const double t = fabs(z) * coef(T)[PHF_FM_KMISC + 2];
    return fma(-t, t, log_pos(T, 0.5 * erfcx_nonneg(T, t)));
}
// the same with the table-driven erfcx (measured on B200: wins when two or four lanes share a chain -- config 2
// +7 % -- and loses with one thread per chain, where twelve warps per SM keep the shared-memory pipe busy with
// scattered 16-byte reads: config 5 -11 %, the hierarchical thread kernel -1.5 %; so only those kernels use it)
PHF_FM double log_ndtr_nonpos_pw(const double *T, double z)
{
    const double t = fabs(z) * coef(T)[PHF_FM_KMISC + 2];
    return fma(-t, t, log_pos(T, 0.5 * erfcx_nonneg_pw(T, t)));
}

// ---- sin and cos of 2 pi b / 2^32 ---------------------------------------------------------------
PHF_FM void sincos_turn32(const double *T, uint32_t b, double &sn, double &cs)
{
    const double kScale = coef(T)[PHF_FM_KMISC + 1];  // pi / 2^31
    const uint32_t bb = b + 0x20000000u;           // nearest multiple of a quarter turn
    const uint32_t quad = bb >> 30;
    const int32_t rem = (int32_t)(bb & 0x3fffffffu) - 0x20000000;  // [-2^29, 2^29)
    const double r = (double)rem * kScale;                          // |r| <= pi/4
    double xp[3];
    xp[0] = r * r;
    xp[1] = xp[0] * xp[0];
    xp[2] = xp[1] * xp[1];
    const double S = poly<0, 5>(coef(T) + PHF_FM_KSINS, xp);
    const double C = poly<0, 5>(coef(T) + PHF_FM_KCOSC, xp);
    const double s0 = fma(r * xp[0], S, r);
    const double c0 = fma(xp[1], C, fma(-0.5, xp[0], 1.0));
    // rotate by quad quarter turns
    const bool swap = quad & 1u;
    double s1 = swap ? c0 : s0;
    double c1 = swap ? s0 : c0;
    sn = (quad & 2u) ? -s1 : s1;
    cs = ((quad + 1u) & 2u) ? -c1 : c1;
}

// ---- 10^x as exp(x ln 10) with a double-double ln 10 (model 1's 1/IC50) -----------------------------
PHF_FM double exp10_clamped(const double *T, double x)
{
    const double kLn10Hi = coef(T)[PHF_FM_KMISC + 3], kLn10Lo = coef(T)[PHF_FM_KMISC + 4];
    const double hi = x * kLn10Hi;
    const double lo = fma(x, kLn10Lo, fma(x, kLn10Hi, -hi));
    const double e = exp_clamped(T, hi);
    return fma(e, lo, e);
}

}  // namespace fm
}  // namespace phf
