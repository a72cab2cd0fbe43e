// Single-level models 1 and 2: log-target (python/doseresponse.py:166-189, 203-248) on the unique-dose
// sufficient-statistic form described in include/pyhillfit_b200.h.
#pragma once
#include "../../include/pyhillfit_b200.h"
#include "phf_math.cuh"

namespace phf {

template <int MODEL>
struct SingleDims {
    static constexpr int D = MODEL == 1 ? 2 : 3;
    static constexpr int NT = D * (D + 1) / 2;
    static constexpr int NF = PHF_STATE_SIZE(D);
};

// Evaluate t * loglik + logprior and the temperature-1 loglik for one parameter vector.
//   grp      : the dataset's dose groups (shared or global memory)
//   ng       : number of groups
// Support (finite value) <=> pIC50 >= -3, 0 <= Hill <= 10, sigma > 1e-3; otherwise -inf like the reference:
//   sigma <= 1e-3 -> likelihood -inf (doseresponse.py:212-214,238-240) and prior -inf (:306-308, log 0);
//   pIC50 < -3 (:153-154) or Hill outside [0,10] (:181-182) -> prior -inf while the likelihood stays finite.
template <int MODEL>
PHF_DI void single_log_target(const double *th, const phf_dose_group *__restrict__ grp, int ng, double pi_bit,
                              double n_other_total, double temperature, double &log_target, double &loglik_t1)
{
    const double pic50 = th[0];
    const double hill = MODEL == 2 ? th[1] : 1.0;
    const double sigma = th[MODEL == 2 ? 2 : 1];
    const bool sigma_ok = sigma > kSigmaLower;  // false for NaN too
    const double sg = sigma_ok ? sigma : 1.0;

    const double inv_s = 1.0 / sg;
    const double log_s = log(sg);
    const double sm = sg - kSigmaLower;
    // log_gamma_prior (doseresponse.py:308) + log_pic50_exponential (:156)
    double prior = kSigmaShapeM1 * log(sm) - sm / kSigmaScale - kPic50ExpRate * pic50;
    bool in_support = sigma_ok && (pic50 >= kPic50ExpLower);
    if (MODEL == 2) in_support = in_support && (hill >= kHillLower) && (hill <= kHillUpper);

    double lic_hi = 0.0, lic_lo = 0.0, inv_ic50 = 0.0;
    if (MODEL == 2)
        ln_ic50(pic50, lic_hi, lic_lo);
    else
        inv_ic50 = exp10(pic50 - 6.0);  // 1/IC50, IC50 = 10**(6-pIC50) (doseresponse.py:87-88)

    double e2 = 0.0, c0 = 0.0, c100 = 0.0;
#pragma unroll 4
    for (int g = 0; g < ng; ++g) {
        const phf_dose_group G = grp[g];
        const double x = MODEL == 2 ? hill_ratio_pow(G.lnc_hi, G.lnc_lo, lic_hi, lic_lo, hill) : G.conc * inv_ic50;
        const double p = hill_response(x);
        const double r = G.ybar - p;
        e2 += fma(G.n_other * r, r, G.ss);
        if (G.n0 > 0.0) c0 = fma(G.n0, log_ndtr_nonpos((0.0 - p) * inv_s), c0);       // st.norm.logcdf(0, p, sigma)
        if (G.n100 > 0.0) c100 = fma(G.n100, log_ndtr_nonpos((p - 100.0) * inv_s), c100);  // st.norm.logsf(100, p, sigma)
    }
    const double temp_1 = n_other_total * log_s;
    const double temp_2 = e2 * (0.5 * inv_s * inv_s);
    const double raw = c0 + c100 - pi_bit - temp_1 - temp_2;
    loglik_t1 = sigma_ok ? raw : -CUDART_INF;
    const double lik = temperature == 0.0 ? 0.0 : temperature * raw;  // doseresponse.py:204-205,230-231
    log_target = in_support ? lik + prior : -CUDART_INF;
}

}  // namespace phf
