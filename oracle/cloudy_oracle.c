/*
 * cloudy_oracle.c — TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Structure-faithful C restatement of the reference's per-parcel coalescence right-hand side, threaded
 * over parcels with OpenMP.  "Structure-faithful" = the reference's cost structure is kept: one
 * moment_source_helper call per upper-triangular (p1,p2) entry, one regularised incomplete gamma
 * evaluation per node per call, no sharing between entries (src/Sources/Coalescence.jl:210-229).
 * It stands in for "the reference Julia path multithreaded over parcels" (Julia is not installed) in
 * bench.py's cpu_baseline / --impl reference legs, and is itself checked against the scipy oracle
 * (tests/test_oracle_c.py).  Nothing in the product links or calls this file.
 *
 * Parity status: pinned against oracle/cloudy_oracle.py (which is pinned against the reference's golden
 * values at the reference's own rtol 1e-3); "parity unpinned" at 1e-9 by the reference's own tests.
 *
 * Follows (paths under /root/reference):
 *   test/examples/utils/box_model_helpers.jl:29-53      rhs_coal!
 *   src/helper_functions.jl:40-53                       normalising factors
 *   src/ParticleDistributions/ParticleDistributions.jl:177-207, :456-541, :557-612, :698-710
 *   src/Sources/Coalescence.jl:115-150, :187-455
 *   src/Sources/Sedimentation.jl:22-37                  get_sedimentation_flux
 *   test/examples/utils/rainshaft_helpers.jl:47-88      make_rainshaft_rhs body
 *   OrdinaryDiffEqSSPRK SSPRK33 (third-party): Shu-Osher SSP(3,3), fixed dt
 * gamma / gamma_inc come from SpecialFunctions.jl (compat "2.5", not vendored); here: libm tgamma and the
 * published series / continued-fraction expansions of P(a,x).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/cloudy_b200.h"

#define EPS 2.220446049250313e-16
#define MAXM (CLOUDY_MAX_P + 2)

static double jl_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static double jl_max(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

/* regularised lower incomplete gamma P(a,x) = gamma_inc(a,x)[1] */
static double gamma_inc_p(double a, double x) {
    if (!(x > 0.0)) return 0.0;
    if (isinf(x)) return 1.0;
    double lg = lgamma(a);
    if (x < a + 1.0) {
        double t = 1.0 / a, s = t;
        for (int n = 1; n < 5000; ++n) {
            t *= x / (a + n);
            s += t;
            if (t < 1e-17 * s) break;
        }
        return s * exp(a * log(x) - x - lg);
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int n = 1; n < 5000; ++n) {
        double an = -n * (n - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - exp(a * log(x) - x - lg) * h;
}

typedef struct { int kind; double n, a, b; } dist_t;

/* ParticleDistributions.jl:177-207 */
static double moment(const dist_t* d, double q) {
    switch (d->kind) {
        case CLOUDY_EXPONENTIAL: return d->n * pow(d->a, q) * tgamma(q + 1.0);
        case CLOUDY_GAMMA: return d->n * pow(d->a, q) * tgamma(q + d->b) / tgamma(d->b);
        case CLOUDY_MONODISPERSE: return d->n * pow(d->a, q);
        default: return d->n * exp(q * d->a + q * q * d->b * d->b / 2);
    }
}

/* ParticleDistributions.jl:456-541 */
static dist_t update_dist_from_moments(int kind, const double* m, const double* k_range) {
    dist_t d;
    d.kind = kind; d.n = 0.0; d.a = 1.0; d.b = 1.0;
    if (kind == CLOUDY_GAMMA) {
        if (m[0] > EPS && m[1] > EPS) {
            d.n = m[0];
            d.b = jl_max(k_range[0], jl_min(k_range[1], (m[1] / m[0]) / (m[2] / m[1] - m[1] / m[0])));
            d.a = m[1] / m[0] / d.b;
        }
    } else if (kind == CLOUDY_LOGNORMAL) {
        if (m[0] > EPS && m[1] > EPS && m[2] > EPS) {
            d.a = log(m[1] * m[1] / pow(m[0], 1.5) / pow(m[2], 0.5));
            d.b = jl_max(EPS, sqrt(log(m[0] * m[2] / (m[1] * m[1]))));
            d.n = m[1] / exp(d.a + 0.5 * d.b * d.b);
        }
    } else {
        if (m[0] > EPS && m[1] > EPS) { d.n = m[0]; d.a = m[1] / m[0]; }
    }
    return d;
}

/* ParticleDistributions.jl:698-710, y given by callback over 1-based j */
typedef struct { const dist_t* d; double p1, p2, x_th, x_min, dx, k, gam_p2k; int n_bins; } msh_t;
static double y_func(const msh_t* s, int j) {
    if (j > s->n_bins) return 0.0;
    double x = exp(s->x_min + (j - 1) * s->dx);
    double th = s->d->a;
    double f = pow(x, s->p1 + s->k - 1.0) * exp(-x / th) * gamma_inc_p(s->p2 + s->k, (s->x_th - x) / th) * s->gam_p2k;
    return x * f;
}
static double simpson(const msh_t* s) {
    int n = s->n_bins, e = n + 1;
    double sum = 0.0;
    for (int j = 5; j <= n - 3; ++j) sum += y_func(s, j);
    double r = sum + (17 * (y_func(s, 1) + y_func(s, e)) + 59 * (y_func(s, 2) + y_func(s, e - 1)) +
                      43 * (y_func(s, 3) + y_func(s, e - 2)) + 49 * (y_func(s, 4) + y_func(s, e - 3))) / 48;
    return s->dx * r;
}

/* ParticleDistributions.jl:557-612 (grid passed in: computed by the host wrapper like the reference does) */
static double moment_source_helper(const dist_t* d, double p1, double p2, double x_th, int n_bins, double x_min, double dx) {
    if (d->kind == CLOUDY_MONODISPERSE) return (d->a < x_th / 2) ? d->n * d->n * pow(d->a, p1 + p2) : 0.0;
    msh_t s;
    s.d = d; s.p1 = p1; s.p2 = p2; s.x_th = x_th; s.x_min = x_min; s.dx = dx; s.n_bins = n_bins;
    s.k = (d->kind == CLOUDY_GAMMA) ? d->b : 1.0;
    s.gam_p2k = tgamma(p2 + s.k);
    double gk = tgamma(s.k);
    return d->n * d->n * pow(d->a, p2 - s.k) / (gk * gk) * simpson(&s);
}

static double binom(int n, int k) { return (n == 2 && k == 1) ? 2.0 : 1.0; }

/* get_coal_ints(AnalyticalCoalStyle, pdists, coal_data) — Coalescence.jl:115-150 */
static int get_coal_ints(const cloudy_config* cfg, const dist_t* pd, double* out) {
    const int N = cfg->n_modes, P = cfg->P, M = P + 2;
    double mom[CLOUDY_MAX_MODES][MAXM];
    double F[CLOUDY_MAX_MODES][MAXM][MAXM];
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < N; ++i) mom[i][j] = (j + 1 <= cfg->n_mom_max) ? moment(&pd[i], (double)j) : 0.0; /* :187-198 */
    for (int i = 0; i < N; ++i) { /* :200-244 */
        for (int j = 1; j <= M; ++j)
            for (int k = 1; k <= M; ++k) {
                double mm = mom[i][j - 1] * mom[i][k - 1], v;
                if (mm < EPS || k < j || cfg->n_2d_ints[i] < j || cfg->n_2d_ints[i] < k) v = 0.0;
                else if (i == N - 1 || isinf(cfg->thresholds[i])) v = mm;
                else {
                    if (pd[i].kind == CLOUDY_LOGNORMAL) return CLOUDY_ERR_UNSUPPORTED;
                    v = jl_min(mm, moment_source_helper(&pd[i], (double)(j - 1), (double)(k - 1), cfg->thresholds[i], cfg->n_bins[i],
                                                        cfg->x_min[i], cfg->dx[i]));
                }
                F[i][j - 1][k - 1] = v;
            }
        for (int j = 0; j < M; ++j)
            for (int k = 0; k < j; ++k) F[i][j][k] = F[i][k][j];
    }
    double Q[3][CLOUDY_MAX_MODES][CLOUDY_MAX_MODES] = {{{0}}}, R[3][CLOUDY_MAX_MODES][CLOUDY_MAX_MODES] = {{{0}}};
    double S[3][2][CLOUDY_MAX_MODES] = {{{0}}};
    for (int m = 0; m < 3; ++m)
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < N; ++j) {
                if (!(k <= j || cfg->nprog[k] <= m)) { /* :260-309 */
                    double t = 0.0;
                    for (int a = 0; a < P; ++a)
                        for (int b = 0; b < P; ++b)
                            for (int c = 0; c <= m; ++c) t += cfg->c[j][k][a][b] * binom(m, c) * mom[j][a + c] * mom[k][b + m - c];
                    Q[m][j][k] = t;
                }
                if (!(cfg->nprog[k] <= m)) { /* :311-351 */
                    double t = 0.0;
                    for (int a = 0; a < P; ++a)
                        for (int b = 0; b < P; ++b) t += cfg->c[j][k][a][b] * mom[j][a] * mom[k][b + m];
                    R[m][j][k] = t;
                }
            }
            if (k < N - 1 && cfg->nprog[k] <= m && cfg->nprog[k + 1] <= m) continue; /* :366-371 */
            if (k == N - 1 && cfg->nprog[k] <= m) continue;
            double s1 = 0.0, s2 = 0.0;
            for (int a = 0; a < P; ++a)
                for (int b = 0; b < P; ++b)
                    for (int c = 0; c <= m; ++c) {
                        double f = F[k][a + c][b + m - c];
                        s1 += 0.5 * cfg->c[k][k][a][b] * binom(m, c) * f;
                        s2 += 0.5 * cfg->c[k][k][a][b] * binom(m, c) * (mom[k][a + c] * mom[k][b + m - c] - f);
                    }
            S[m][0][k] = s1;
            S[m][1][k] = s2;
        }
    int o = 0;
    for (int k = 0; k < N; ++k) /* :140-149 */
        for (int m = 0; m < cfg->nprog[k]; ++m) {
            double sq = 0.0, sr = 0.0;
            for (int j = 0; j < N; ++j) { sq += Q[m][j][k]; sr += R[m][j][k]; }
            double v = sq - sr + S[m][0][k];
            if (k > 0) v += S[m][1][k - 1];
            out[o++] = v;
        }
    return 0;
}

/* rhs_coal! for n parcels, AoS in/out [n][n_slots] — box_model_helpers.jl:29-53 */
int cloudy_oracle_rhs_coal_batch(const cloudy_config* cfg, const double* m, double* dm, int64_t n, int n_threads) {
    const int N = cfg->n_modes;
    int nslots = 0, slot0[CLOUDY_MAX_MODES];
    double norm[CLOUDY_MAX_SLOTS];
    for (int i = 0; i < N; ++i) {
        slot0[i] = nslots;
        for (int q = 0; q < cfg->nprog[i]; ++q) norm[nslots++] = cfg->norms[0] * pow(cfg->norms[1], (double)q);
    }
    int status = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t p = 0; p < n; ++p) {
        dist_t pd[CLOUDY_MAX_MODES];
        double mn[CLOUDY_MAX_SLOTS], ci[CLOUDY_MAX_SLOTS];
        for (int s = 0; s < nslots; ++s) mn[s] = m[p * nslots + s] / norm[s];
        for (int i = 0; i < N; ++i) pd[i] = update_dist_from_moments(cfg->kind[i], mn + slot0[i], cfg->k_range);
        int rc = get_coal_ints(cfg, pd, ci);
        if (rc) {
#pragma omp atomic write
            status = rc;
        }
        for (int s = 0; s < nslots; ++s) dm[p * nslots + s] = ci[s] * norm[s];
    }
    return status;
}

/* ---- layout helpers ---- */
static int slot_layout(const cloudy_config* cfg, int* slot0, double* norm) {
    int nslots = 0;
    for (int i = 0; i < cfg->n_modes; ++i) {
        slot0[i] = nslots;
        for (int q = 0; q < cfg->nprog[i]; ++q) norm[nslots++] = cfg->norms[0] * pow(cfg->norms[1], (double)q);
    }
    return nslots;
}

/* get_sedimentation_flux(pdists, vel) .* mom_norms with vel normalised as rainshaft_helpers.jl:74-77 — Sedimentation.jl:22-37 */
static void sedimentation_flux_cell(const cloudy_config* cfg, const dist_t* pd, const int* slot0, const double* norm, double* out) {
    for (int i = 0; i < cfg->n_modes; ++i)
        for (int j = 0; j < cfg->nprog[i]; ++j) {
            double sacc = 0.0;
            for (int v = 0; v < cfg->n_vel; ++v) {
                const double beta = cfg->vel[v][1];
                const double vn = cfg->vel[v][0] * pow(cfg->norms[1], beta);
                sacc += -vn * moment(&pd[i], (double)j + beta);
            }
            out[slot0[i] + j] = sacc * norm[slot0[i] + j];
        }
}

/* sedimentation flux of n cells, AoS [n][n_slots] (no clipping: the caller passes the state it wants evaluated) */
int cloudy_oracle_sedimentation_flux_batch(const cloudy_config* cfg, const double* m, double* out, int64_t n) {
    int slot0[CLOUDY_MAX_MODES];
    double norm[CLOUDY_MAX_SLOTS];
    const int nslots = slot_layout(cfg, slot0, norm);
    for (int64_t p = 0; p < n; ++p) {
        dist_t pd[CLOUDY_MAX_MODES];
        double mn[CLOUDY_MAX_SLOTS];
        for (int s = 0; s < nslots; ++s) mn[s] = m[p * nslots + s] / norm[s];
        for (int i = 0; i < cfg->n_modes; ++i) pd[i] = update_dist_from_moments(cfg->kind[i], mn + slot0[i], cfg->k_range);
        sedimentation_flux_cell(cfg, pd, slot0, norm, out + p * nslots);
    }
    return 0;
}

/* the column right-hand side — rainshaft_helpers.jl:47-88.  `m` is [n_columns][nz][n_slots] and is CLIPPED IN PLACE (:52);
 * per level: coalescence (zero when every normalised moment is below eps, :67-68) + sedimentation flux; zero flux above
 * the top (:80-81); first-order upwind divergence (:83-85). */
int cloudy_oracle_rainshaft_rhs(const cloudy_config* cfg, double* m, double* dm, int64_t n_columns, int n_threads) {
    int slot0[CLOUDY_MAX_MODES];
    double norm[CLOUDY_MAX_SLOTS];
    const int nslots = slot_layout(cfg, slot0, norm);
    const int nz = cfg->nz;
    const int64_t n = n_columns * nz;
    int status = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    for (int64_t q = 0; q < n * nslots; ++q)
        if (m[q] < 0.0) m[q] = 0.0;
    /* dm first receives the coalescence source; the flux of every cell goes to a scratch array of the same shape */
    double* flux = (double*)malloc(sizeof(double) * (size_t)n * nslots);
    if (!flux) return CLOUDY_ERR_ARG;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t p = 0; p < n; ++p) {
        dist_t pd[CLOUDY_MAX_MODES];
        double mn[CLOUDY_MAX_SLOTS], ci[CLOUDY_MAX_SLOTS];
        int all_small = 1;
        for (int s = 0; s < nslots; ++s) {
            mn[s] = m[p * nslots + s] / norm[s];
            if (!(mn[s] < EPS)) all_small = 0;
        }
        for (int i = 0; i < cfg->n_modes; ++i) pd[i] = update_dist_from_moments(cfg->kind[i], mn + slot0[i], cfg->k_range);
        if (all_small) {
            for (int s = 0; s < nslots; ++s) dm[p * nslots + s] = 0.0;
        } else {
            int rc = get_coal_ints(cfg, pd, ci);
            if (rc) {
#pragma omp atomic write
                status = rc;
            }
            for (int s = 0; s < nslots; ++s) dm[p * nslots + s] = ci[s] * norm[s];
        }
        sedimentation_flux_cell(cfg, pd, slot0, norm, flux + p * nslots);
    }
    for (int64_t c = 0; c < n_columns; ++c)
        for (int z = 0; z < nz; ++z) {
            const int64_t p = c * nz + z;
            for (int s = 0; s < nslots; ++s) {
                const double up = (z == nz - 1) ? 0.0 : flux[(p + 1) * nslots + s];
                dm[p * nslots + s] = dm[p * nslots + s] + (-(up - flux[p * nslots + s]) / cfg->dz);
            }
        }
    free(flux);
    return status;
}

/* n_steps of SSPRK33 (u1 = u + dt f(u); u2 = (3u + u1 + dt f(u1))/4; u+ = (u + 2 u2 + 2 dt f(u2))/3), in place.
 * model 0: box (rhs_coal! per parcel, `n` parcels); model 1: rainshaft (`n` columns of cfg->nz cells; the right-hand side
 * clips the array it is evaluated at, as the reference does). */
int cloudy_oracle_ssprk33(const cloudy_config* cfg, double* u, double dt, int n_steps, int model, int64_t n, int n_threads) {
    int slot0[CLOUDY_MAX_MODES];
    double norm[CLOUDY_MAX_SLOTS];
    const int nslots = slot_layout(cfg, slot0, norm);
    const int64_t cells = (model == 1) ? n * cfg->nz : n;
    const size_t len = (size_t)cells * nslots;
    double* k = (double*)malloc(sizeof(double) * len);
    double* tmp = (double*)malloc(sizeof(double) * len);
    if (!k || !tmp) { free(k); free(tmp); return CLOUDY_ERR_ARG; }
    int status = 0;
    for (int step = 0; step < n_steps && !status; ++step) {
        status = (model == 1) ? cloudy_oracle_rainshaft_rhs(cfg, u, k, n, n_threads) : cloudy_oracle_rhs_coal_batch(cfg, u, k, n, n_threads);
        for (size_t i = 0; i < len; ++i) tmp[i] = u[i] + dt * k[i];
        if (!status) status = (model == 1) ? cloudy_oracle_rainshaft_rhs(cfg, tmp, k, n, n_threads) : cloudy_oracle_rhs_coal_batch(cfg, tmp, k, n, n_threads);
        for (size_t i = 0; i < len; ++i) tmp[i] = (3 * u[i] + tmp[i] + dt * k[i]) / 4;
        if (!status) status = (model == 1) ? cloudy_oracle_rainshaft_rhs(cfg, tmp, k, n, n_threads) : cloudy_oracle_rhs_coal_batch(cfg, tmp, k, n, n_threads);
        for (size_t i = 0; i < len; ++i) u[i] = (u[i] + 2 * tmp[i] + 2 * dt * k[i]) / 3;
    }
    free(k);
    free(tmp);
    return status;
}

#ifndef CLOUDY_ORACLE_FLAGS
#define CLOUDY_ORACLE_FLAGS "unknown"
#endif
const char* cloudy_oracle_build_flags(void) { return CLOUDY_ORACLE_FLAGS; }

int cloudy_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

double cloudy_oracle_moment_source_helper(int kind, const double* params, double p1, double p2, double x_th, int n_bins, double x_min,
                                          double dx) {
    dist_t d;
    d.kind = kind; d.n = params[0]; d.a = params[1]; d.b = params[2];
    return moment_source_helper(&d, p1, p2, x_th, n_bins, x_min, dx);
}
