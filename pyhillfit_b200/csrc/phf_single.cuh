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

// One unique dose's contribution: centred Gaussian term (sum over its 0 < y < 100 replicates) and the
// censored terms n0 logPhi((0-p)/sigma) + n100 logPhi((p-100)/sigma)  (doseresponse.py:215-222, 241-248).
template <int MODEL>
PHF_DI void dose_group_terms(const double *T, const phf_dose_group &Gd, double hill, double lic_hi, double lic_lo, double inv_ic50,
                             double inv_s, double &e2, double &cens)
{
    const double x = MODEL == 2 ? hill_ratio_pow(T, Gd.lnc_hi, Gd.lnc_lo, lic_hi, lic_lo, hill) : Gd.conc * inv_ic50;
    const double p = hill_response(x);
    const double r = Gd.ybar - p;
    e2 += fma(Gd.n_other * r, r, Gd.ss);
    const bool has0 = Gd.n0 > 0.0, has100 = Gd.n100 > 0.0;
    if (has0 || has100) {
        // one evaluation stream serves either kind of censoring; a dose carrying both zeros and hundreds
        // (none in the Crumb table) takes the second call
        const double z = has0 ? (0.0 - p) * inv_s : (p - 100.0) * inv_s;  // st.norm.logcdf(0,p,s) / logsf(100,p,s)
        cens = fma(has0 ? Gd.n0 : Gd.n100, log_ndtr_nonpos(T, z), cens);
        if (has0 && has100) cens = fma(Gd.n100, log_ndtr_nonpos(T, (p - 100.0) * inv_s), cens);
    }
}

// Evaluate t * loglik + logprior and the temperature-1 loglik for one parameter vector with G cooperating
// lanes (G = 1: a single thread).  Every lane of the group passes the same th; lane gl evaluates dose groups
// gl, gl+G, ...; lanes 0 and 1 evaluate the two logarithms of sigma; sums are butterfly reductions, so every
// lane returns the same bits.
//   g_own    : dose group `gl` of the dataset, preloaded (ignored when gl >= ng)
//   grp      : the dataset's dose groups (shared or global memory), used for groups >= G
// Support (finite value) <=> pIC50 >= -3, 0 <= Hill <= 10, sigma > 1e-3; otherwise -inf like the reference:
//   sigma <= 1e-3 -> likelihood -inf (doseresponse.py:212-214,238-240) and prior -inf (:306-308, log 0);
//   pIC50 < -3 (:153-154) or Hill outside [0,10] (:181-182) -> prior -inf while the likelihood stays finite.
template <int MODEL, int G>
PHF_DI void single_log_target_lanes(const double *T, const double *th, const phf_dose_group &g_own,
                                    const phf_dose_group *__restrict__ grp, int ng, double pi_bit,
                                    double n_other_total, double temperature, int gl, unsigned mask,
                                    double &log_target, double &loglik_t1)
{
    const double pic50 = th[0];
    const double hill = MODEL == 2 ? th[1] : 1.0;
    const double sigma = th[MODEL == 2 ? 2 : 1];
    const bool sigma_ok = sigma > kSigmaLower;  // false for NaN too
    const double sg = sigma_ok ? sigma : 1.0;
    const double inv_s = fm::rcp(sg);
    const double sm = sg - kSigmaLower;
    double log_s, log_sm;
    if (G >= 2) {
        const double lv = fm::log_pos(T, gl == 0 ? sg : sm);
        log_s = __shfl_sync(mask, lv, 0, G);
        log_sm = __shfl_sync(mask, lv, 1, G);
    } else {
        log_s = fm::log_pos(T, sg);
        log_sm = fm::log_pos(T, sm);
    }
    // log_gamma_prior (doseresponse.py:308) + log_pic50_exponential (:156)
    const double prior = kSigmaShapeM1 * log_sm - sm * (1.0 / kSigmaScale) - kPic50ExpRate * pic50;
    bool in_support = sigma_ok && (pic50 >= kPic50ExpLower);
    if (MODEL == 2) in_support = in_support && (hill >= kHillLower) && (hill <= kHillUpper);

    double lic_hi = 0.0, lic_lo = 0.0, inv_ic50 = 0.0;
    if (MODEL == 2)
        ln_ic50(pic50, lic_hi, lic_lo);
    else
        inv_ic50 = fm::exp10_clamped(T, pic50 - 6.0);  // 1/IC50, IC50 = 10**(6-pIC50) (doseresponse.py:87-88)

    double e2 = 0.0, cens = 0.0;
    if (gl < ng) dose_group_terms<MODEL>(T, g_own, hill, lic_hi, lic_lo, inv_ic50, inv_s, e2, cens);
    for (int g = gl + G; g < ng; g += G) {
        const phf_dose_group Gd = grp[g];
        dose_group_terms<MODEL>(T, Gd, hill, lic_hi, lic_lo, inv_ic50, inv_s, e2, cens);
    }
    if (G >= 2) {
        e2 = group_sum<G>(e2, mask);
        cens = group_sum<G>(cens, mask);
    }
    const double temp_1 = n_other_total * log_s;
    const double temp_2 = e2 * (0.5 * inv_s * inv_s);
    const double raw = cens - pi_bit - temp_1 - temp_2;
    loglik_t1 = sigma_ok ? raw : -CUDART_INF;
    const double lik = temperature == 0.0 ? 0.0 : temperature * raw;  // doseresponse.py:204-205,230-231
    log_target = in_support ? lik + prior : -CUDART_INF;
}

// single-thread form used by the batch / init kernels
template <int MODEL>
PHF_DI void single_log_target(const double *T, const double *th, const phf_dose_group *__restrict__ grp, int ng, double pi_bit,
                              double n_other_total, double temperature, double &log_target, double &loglik_t1)
{
    phf_dose_group g0 = {};
    if (ng > 0) g0 = grp[0];
    single_log_target_lanes<MODEL, 1>(T, th, g0, grp, ng, pi_bit, n_other_total, temperature, 0, 0xffffffffu, log_target,
                                      loglik_t1);
}

}  // namespace phf
