/*
 * CPU oracle (plain C restatement) of PyHillFit's MCMC hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- not part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs load this library (built into
 * oracle/_build/libhill_oracle.so by oracle/Makefile); pyhillfit_b200/ never does.
 *
 * It follows the reference's *per-point* algorithm (no unique-dose compression, no
 * sufficient statistics) so that it is an independent check of the CUDA kernels'
 * compressed formulation.  Paths below are relative to /root/reference.
 *
 * Pinning: tests/test_oracle_golden.py compares every function here with
 * the .npz files under tests/golden, which hold outputs of the unmodified reference executed in the build
 * container through oracle/ref_shim.py (generator: oracle/gen_golden.py).
 *
 * Third-party arithmetic restated here:
 *   scipy.special.log_ndtr (xsf):  x < -1 ? log(erfcx(-x/sqrt2)/2) - x*x/2 : log1p(-erfc(x/sqrt2)/2)
 *   scipy.special.ndtr (cephes):   0.5 + 0.5 erf(x/sqrt2)  |  0.5 erfc(|x|/sqrt2) (reflected for x > 0)
 *   numpy.random (MT19937 + SVD multivariate normal) is REPLACED by Philox4x32-10 + Box-Muller +
 *   Cholesky, the stream the CUDA kernels use (contract in oracle/hill_oracle.py).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PHF_INF (1.0 / 0.0)

/* ---- constants: python/doseresponse.py:12-25 ---- */
static const double sigma_uniform_lower = 1e-3;
static const double pic50_exp_rate = 0.2;
static const double pic50_exp_lower = -3.0;
static const double hill_uniform_lower = 0.0;
static const double hill_uniform_upper = 10.0;
static const double sigma_shape = 5.0;
static const double sigma_loc = 1e-3;
#define SIGMA_SCALE ((6.0 - 1e-3) / (5.0 - 1.0))

/* ---- scipy.special.log_ndtr restated with libm (erfcx is not in libm) ---- */
double phf_oracle_log_ndtr(double x)
{
    double t = x * M_SQRT1_2;
    if (x < -1.0) {
        if (-t < 25.0) /* log(erfcx(z)/2) - z*z == log(erfc(z)/2) while erfc(z) is a normal number */
            return log(0.5 * erfc(-t));
        /* asymptotic series of the Mills ratio, |x| > 35: relative truncation error < 1e-18 */
        double ix2 = 1.0 / (x * x), term = 1.0, sum = 1.0;
        for (int k = 1; k <= 10; ++k) {
            term *= -(2.0 * k - 1.0) * ix2;
            sum += term;
        }
        return -0.5 * x * x - log(-x) - 0.5 * log(2.0 * M_PI) + log(sum);
    }
    return log1p(-0.5 * erfc(t));
}

/* ---- scipy.special.ndtr (cephes ndtr.c) ---- */
double phf_oracle_ndtr(double a)
{
    double x = a * M_SQRT1_2, z = fabs(x), y;
    if (z < 1.0)
        y = 0.5 + 0.5 * erf(x);
    else {
        y = 0.5 * erfc(z);
        if (x > 0) y = 1.0 - y;
    }
    return y;
}

/* ---- python/doseresponse.py:84-88 ---- */
static double dose_response_model(double dose, double hill, double ic50)
{
    return 100. * (1. - 1. / (1. + pow(1. * dose / ic50, hill)));
}
static double pic50_to_ic50(double pic50) { return pow(10.0, 6 - pic50); }

/* ---- python/doseresponse.py:151-156, 304-317, 166-184 ---- */
static double log_pic50_exponential(double x) { return x < pic50_exp_lower ? -PHF_INF : -pic50_exp_rate * x; }
static double log_gamma_prior(double x, double shape, double scale, double loc)
{
    if (x < loc) return -PHF_INF;
    return (shape - 1) * log(x - loc) - (x - loc) / scale;
}
double phf_oracle_log_priors(int model, const double *params)
{
    double pic50 = params[0], sigma = params[model == 1 ? 1 : 2];
    if (model != 1) {
        double hill = params[1];
        if (hill < hill_uniform_lower || hill > hill_uniform_upper) return -PHF_INF;
    }
    return log_pic50_exponential(pic50) + log_gamma_prior(sigma, sigma_shape, SIGMA_SCALE, sigma_loc);
}

/*
 * python/doseresponse.py:203-248.  cls[i]: 0 -> where_y_0, 1 -> where_y_100, 2 -> where_y_other,
 * 3 -> in no mask (dropped).  Returns t * raw; *ll1 receives raw (the temperature-1 value that
 * compute_bayes_factors.py:18-21 re-evaluates), also when t == 0.
 */
double phf_oracle_log_data_likelihood(int model, int n, const double *conc, const double *y, const uint8_t *cls,
                                      const double *params, double t, double pi_bit, double *ll1)
{
    double pic50 = params[0], hill = model == 1 ? 1.0 : params[1], sigma = params[model == 1 ? 1 : 2];
    if (sigma <= sigma_uniform_lower) {
        if (ll1) *ll1 = -PHF_INF;
        return t == 0 ? 0.0 : -PHF_INF;
    }
    double ic50 = pic50_to_ic50(pic50);
    double y_0_sum = 0, y_100_sum = 0, temp_2 = 0;
    int n_other = 0;
    for (int i = 0; i < n; ++i) {
        double p = dose_response_model(conc[i], hill, ic50);
        if (cls[i] == 0)
            y_0_sum += phf_oracle_log_ndtr((0 - p) / sigma); /* st.norm.logcdf(0, p, sigma) */
        else if (cls[i] == 1)
            y_100_sum += phf_oracle_log_ndtr(-((100 - p) / sigma)); /* st.norm.logsf(100, p, sigma) */
        else if (cls[i] == 2) {
            n_other++;
            temp_2 += (y[i] - p) * (y[i] - p) / (2. * sigma * sigma);
        }
    }
    double temp_1 = n_other * log(sigma);
    double raw = y_0_sum + y_100_sum - pi_bit - temp_1 - temp_2;
    if (ll1) *ll1 = raw;
    return t == 0 ? 0.0 : t * raw;
}

/* python/doseresponse.py:187-189 */
double phf_oracle_log_target(int model, int n, const double *conc, const double *y, const uint8_t *cls,
                             const double *params, double t, double pi_bit, double *ll1)
{
    return phf_oracle_log_data_likelihood(model, n, conc, y, cls, params, t, pi_bit, ll1) +
           phf_oracle_log_priors(model, params);
}

void phf_oracle_log_target_batch(int model, int n, const double *conc, const double *y, const uint8_t *cls,
                                 int n_theta, const double *theta, const double *t, double pi_bit, double *out,
                                 double *ll1_out)
{
    int d = model == 1 ? 2 : 3;
    for (int k = 0; k < n_theta; ++k) {
        double ll1;
        out[k] = phf_oracle_log_target(model, n, conc, y, cls, theta + (size_t)k * d, t[k], pi_bit, &ll1);
        if (ll1_out) ll1_out[k] = ll1;
    }
}

/* ---- hierarchical target: python/PyHillFit.py:113-154, 173-193 ---- */
typedef struct {
    int ne;
    const int *off; /* [ne+1] point offsets per experiment */
    const double *conc, *y;
    double shapes[5], scales[5], locs[5];
} hier_data;

static double hier_target(const hier_data *hd, const double *theta)
{
    int ne = hd->ne, dim = 5 + 2 * ne;
    for (int i = 0; i < 4; ++i)
        if (theta[i] <= hd->locs[i]) return -PHF_INF;
    double alpha = theta[0], beta = theta[1], mu = theta[2], s = theta[3], sigma = theta[dim - 1];
    for (int e = 0; e < ne; ++e)
        if (theta[5 + 2 * e] < 0 || theta[4 + 2 * e] < -2.0) return -PHF_INF;
    if (sigma <= hd->locs[4]) return -PHF_INF;
    double answer = 0.;
    for (int e = 0; e < ne; ++e) { /* PyHillFit.py:113-132 */
        double ic50 = pic50_to_ic50(theta[4 + 2 * e]), hill = theta[5 + 2 * e];
        double ss = 0, trunc = 0;
        int npts = hd->off[e + 1] - hd->off[e];
        for (int i = hd->off[e]; i < hd->off[e + 1]; ++i) {
            double p = dose_response_model(hd->conc[i], hill, ic50);
            ss += (hd->y[i] - p) * (hd->y[i] - p);
            trunc += log(phf_oracle_ndtr((100 - p) / sigma) - phf_oracle_ndtr((0 - p) / sigma));
        }
        answer -= (npts * log(sigma) + ss / (2 * sigma * sigma) + trunc);
    }
    double total = answer, acc = 0;
    for (int e = 0; e < ne; ++e) { /* PyHillFit.py:134-135 */
        double x = theta[5 + 2 * e];
        acc += log(beta) - beta * log(alpha) + (beta - 1.) * log(x) - 2 * log(1 + pow(x / alpha, beta));
    }
    total += acc;
    acc = 0;
    for (int e = 0; e < ne; ++e) { /* PyHillFit.py:144-146 */
        double tb = (theta[4 + 2 * e] - mu) / s;
        acc += -tb - log(s) - 2 * log(1 + exp(-tb));
    }
    total += acc;
    acc = 0;
    const int idx[5] = {0, 1, 2, 3, dim - 1};
    for (int k = 0; k < 5; ++k) { /* dr.log_gamma_prior vectorised over 5 entries, PyHillFit.py:187 */
        double x = theta[idx[k]];
        if (x < hd->locs[k]) return -PHF_INF;
        acc += (hd->shapes[k] - 1) * log(x - hd->locs[k]) - (x - hd->locs[k]) / hd->scales[k];
    }
    return total + acc;
}

double phf_oracle_hier_log_target(int ne, const int *off, const double *conc, const double *y, const double *theta,
                                  const double *shapes, const double *scales, const double *locs)
{
    hier_data hd = {ne, off, conc, y, {0}, {0}, {0}};
    memcpy(hd.shapes, shapes, sizeof hd.shapes);
    memcpy(hd.scales, scales, sizeof hd.scales);
    memcpy(hd.locs, locs, sizeof hd.locs);
    return hier_target(&hd, theta);
}

void phf_oracle_hier_log_target_batch(int ne, const int *off, const double *conc, const double *y, int n_theta,
                                      const double *theta, const double *shapes, const double *scales,
                                      const double *locs, double *out)
{
    int dim = 5 + 2 * ne;
    for (int k = 0; k < n_theta; ++k)
        out[k] = phf_oracle_hier_log_target(ne, off, conc, y, theta + (size_t)k * dim, shapes, scales, locs);
}

/* ---- Philox4x32-10 + Box-Muller stream contract (see oracle/hill_oracle.py) ---- */
void phf_oracle_philox(uint64_t seed, uint64_t chain, uint32_t t, uint32_t j, uint32_t out[4])
{
    uint32_t c0 = t, c1 = j, c2 = (uint32_t)chain, c3 = (uint32_t)(chain >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double uniform53(uint32_t w0, uint32_t w1)
{
    uint64_t v = (((uint64_t)w0 << 32) | w1) >> 11;
    return ((double)v + 0.5) * 0x1p-53;
}

static void box_muller(uint32_t a, uint32_t b, double *z0, double *z1)
{
    double r = sqrt(-2.0 * log(((double)a + 1.0) * 0x1p-32));
    double phi = M_PI * ((double)b * 0x1p-31);
    *z0 = r * cos(phi);
    *z1 = r * sin(phi);
}

void phf_oracle_draw(uint64_t seed, uint64_t chain, uint32_t t, int d, double *u, double *z /* [d+1] scratch ok */)
{
    uint32_t w[4];
    double zz[2];
    int k = 0;
    phf_oracle_philox(seed, chain, t, 0, w);
    *u = uniform53(w[0], w[1]);
    box_muller(w[2], w[3], &zz[0], &zz[1]);
    for (int i = 0; i < 2 && k < d; ++i) z[k++] = zz[i];
    for (uint32_t j = 1; k < d; ++j) {
        phf_oracle_philox(seed, chain, t, j, w);
        box_muller(w[0], w[1], &zz[0], &zz[1]);
        for (int i = 0; i < 2 && k < d; ++i) z[k++] = zz[i];
        if (k < d) {
            box_muller(w[2], w[3], &zz[0], &zz[1]);
            for (int i = 0; i < 2 && k < d; ++i) z[k++] = zz[i];
        }
    }
}

/* ---- adaptive Metropolis: PyHillFit.py:828-856 / 481-511, PyHillTemp.py:87-123 ---- */
typedef double (*target_fn)(const void *ctx, const double *theta, double *ll1);

typedef struct {
    int model, n;
    const double *conc, *y;
    const uint8_t *cls;
    double t, pi_bit;
} single_ctx;

static double single_target(const void *c, const double *theta, double *ll1)
{
    const single_ctx *s = (const single_ctx *)c;
    return phf_oracle_log_target(s->model, s->n, s->conc, s->y, s->cls, theta, s->t, s->pi_bit, ll1);
}
static double hier_target_cb(const void *c, const double *theta, double *ll1)
{
    if (ll1) *ll1 = 0;
    return hier_target((const hier_data *)c, theta);
}

/* lower Cholesky factor, row-major d x d; returns 0 on success.
 * Guarded pivots: the adapted covariance (1-g) C + g dd' is positive semi-definite but can be numerically singular
 * (right after adaptation starts it is the empirical covariance of a path that has hardly moved); the reference
 * draws through numpy's SVD factor (multivariate_normal), which tolerates that silently.  A pivot is therefore
 * floored at PHF_PIVOT_FLOOR times its diagonal entry; the CUDA kernels apply the same floor. */
#define PHF_PIVOT_FLOOR 1e-12
static int cholesky(int d, const double *a, double *l)
{
    memset(l, 0, sizeof(double) * d * d);
    for (int j = 0; j < d; ++j) {
        double s = a[j * d + j];
        const double floor_j = PHF_PIVOT_FLOOR * a[j * d + j];
        for (int k = 0; k < j; ++k) s -= l[j * d + k] * l[j * d + k];
        if (!(s > floor_j)) s = floor_j;
        if (!(s > 0)) return -1;
        double ljj = sqrt(s);
        l[j * d + j] = ljj;
        for (int i = j + 1; i < d; ++i) {
            double v = a[i * d + j];
            for (int k = 0; k < j; ++k) v -= l[i * d + k] * l[j * d + k];
            l[i * d + j] = v / ljj;
        }
    }
    return 0;
}

/*
 * state layout (in/out), doubles: theta[d], log_target, ll1, mean[d], cov[d*d] (full, row-major), loga,
 * ll1_sum, n_accept.  Iterations run are t0+1 .. t0+iters (1-based, global).  Rows saved at t % thinning == 0
 * go to chain_out[(t/thinning - row0) * (d+1)]; rows with index >= burn add ll1 to ll1_sum.
 */
static int am_run(target_fn f, const void *ctx, int d, double *state, uint32_t t0, uint32_t iters, uint32_t thinning,
                  uint32_t adapt_when, int reset_mean, uint64_t seed, uint64_t chain, uint32_t row0, uint32_t burn,
                  double *chain_out)
{
    double *theta = state, *lt = state + d, *ll1 = state + d + 1, *mean = state + d + 2, *cov = state + 2 * d + 2;
    double *loga = cov + d * d, *ll1_sum = loga + 1, *n_acc = loga + 2;
    double *l = malloc(sizeof(double) * (d * d + 3 * d + 2));
    double *z = l + d * d, *star = z + d + 1, *diff = star + d;
    int rc = 0;
    for (uint32_t t = t0 + 1; t <= t0 + iters; ++t) {
        double u, lt_star, ll1_star;
        phf_oracle_draw(seed, chain, t, d, &u, z);
        if (cholesky(d, cov, l)) { rc = -1; break; }
        double sc = exp(0.5 * *loga);
        for (int i = 0; i < d; ++i) {
            double a = 0;
            for (int k = 0; k <= i; ++k) a += l[i * d + k] * z[k];
            star[i] = theta[i] + sc * a;
        }
        lt_star = f(ctx, star, &ll1_star);
        int accepted = log(u) < lt_star - *lt;
        if (accepted) {
            memcpy(theta, star, sizeof(double) * d);
            *lt = lt_star;
            *ll1 = ll1_star;
            *n_acc += 1;
        }
        if (t % thinning == 0) {
            uint32_t row = t / thinning;
            if (chain_out) {
                double *o = chain_out + (size_t)(row - row0) * (d + 1);
                memcpy(o, theta, sizeof(double) * d);
                o[d] = *lt;
            }
            if (row >= burn) *ll1_sum += *ll1;
        }
        if (reset_mean && t == adapt_when) memcpy(mean, theta, sizeof(double) * d);
        if (t > adapt_when) {
            double s = (double)(t - adapt_when);
            double g = 1. / pow(s + 1., 0.6);
            for (int i = 0; i < d; ++i) diff[i] = theta[i] - mean[i];
            for (int i = 0; i < d; ++i)
                for (int k = 0; k < d; ++k) cov[i * d + k] = (1 - g) * cov[i * d + k] + g * (diff[i] * diff[k]);
            for (int i = 0; i < d; ++i) mean[i] = (1 - g) * mean[i] + g * theta[i];
            *loga += g * (accepted - 0.25);
        }
    }
    free(l);
    return rc;
}

int phf_oracle_am_single(int model, int n, const double *conc, const double *y, const uint8_t *cls, double temperature,
                         double pi_bit, double *state, uint32_t t0, uint32_t iters, uint32_t thinning,
                         uint32_t adapt_when, int reset_mean, uint64_t seed, uint64_t chain, uint32_t row0,
                         uint32_t burn, double *chain_out)
{
    single_ctx c = {model, n, conc, y, cls, temperature, pi_bit};
    return am_run(single_target, &c, model == 1 ? 2 : 3, state, t0, iters, thinning, adapt_when, reset_mean, seed,
                  chain, row0, burn, chain_out);
}

int phf_oracle_am_hier(int ne, const int *off, const double *conc, const double *y, const double *shapes,
                       const double *scales, const double *locs, double *state, uint32_t t0, uint32_t iters,
                       uint32_t thinning, uint32_t adapt_when, uint64_t seed, uint64_t chain, uint32_t row0,
                       double *chain_out)
{
    hier_data hd = {ne, off, conc, y, {0}, {0}, {0}};
    memcpy(hd.shapes, shapes, sizeof hd.shapes);
    memcpy(hd.scales, scales, sizeof hd.scales);
    memcpy(hd.locs, locs, sizeof hd.locs);
    return am_run(hier_target_cb, &hd, 5 + 2 * ne, state, t0, iters, thinning, adapt_when, 0, seed, chain, row0,
                  0xFFFFFFFFu, chain_out);
}

/* initial evaluation helpers so callers can fill state[d], state[d+1] */
double phf_oracle_single_init(int model, int n, const double *conc, const double *y, const uint8_t *cls,
                              double temperature, double pi_bit, const double *theta, double *ll1)
{
    return phf_oracle_log_target(model, n, conc, y, cls, theta, temperature, pi_bit, ll1);
}

/*
 * Many independent single-level chains, one per (dataset, temperature) entry, spread over `n_threads`
 * pthreads (chains are dealt round-robin).  Used as the "port" CPU baseline (bench.py) -- same arithmetic
 * as phf_oracle_am_single.
 */
typedef struct {
    int first, stride, n_chains;
    const int *model, *npts;
    const int64_t *data_off;
    const double *conc, *y;
    const uint8_t *cls;
    const double *temperature, *pi_bit;
    double *states;
    const int64_t *state_off;
    uint32_t t0, iters, thinning;
    const uint32_t *adapt_when;
    int reset_mean;
    uint64_t seed;
    const uint64_t *chain_ids;
    int rc;
} many_job;

static void *many_worker(void *p)
{
    many_job *j = (many_job *)p;
    for (int k = j->first; k < j->n_chains; k += j->stride) {
        int r = phf_oracle_am_single(j->model[k], j->npts[k], j->conc + j->data_off[k], j->y + j->data_off[k],
                                     j->cls + j->data_off[k], j->temperature[k], j->pi_bit[k],
                                     j->states + j->state_off[k], j->t0, j->iters, j->thinning, j->adapt_when[k],
                                     j->reset_mean, j->seed, j->chain_ids[k], 0, 0xFFFFFFFFu, NULL);
        if (r) j->rc = r;
    }
    return NULL;
}

int phf_oracle_am_single_many(int n_chains, const int *model, const int *npts, const int64_t *data_off,
                              const double *conc, const double *y, const uint8_t *cls, const double *temperature,
                              const double *pi_bit, double *states, const int64_t *state_off, uint32_t t0,
                              uint32_t iters, uint32_t thinning, const uint32_t *adapt_when, int reset_mean,
                              uint64_t seed, const uint64_t *chain_ids, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t tid[256];
    many_job jobs[256];
    int rc = 0;
    for (int i = 0; i < n_threads; ++i) {
        many_job j = {i, n_threads, n_chains, model, npts, data_off, conc, y, cls, temperature, pi_bit, states,
                      state_off, t0, iters, thinning, adapt_when, reset_mean, seed, chain_ids, 0};
        jobs[i] = j;
        pthread_create(&tid[i], NULL, many_worker, &jobs[i]);
    }
    for (int i = 0; i < n_threads; ++i) {
        pthread_join(tid[i], NULL);
        if (jobs[i].rc) rc = jobs[i].rc;
    }
    return rc;
}
