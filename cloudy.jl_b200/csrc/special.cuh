// Device-side special functions and Julia-semantics helpers for the coalescence path.
// sm_100a, FP64 on CUDA cores.  No reference code is copied here: the reference takes
// gamma / gamma_inc from SpecialFunctions.jl (not vendored); these are the published series /
// continued-fraction expansions of the incomplete gamma function.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cloudy {

constexpr double kEps = 2.220446049250313e-16;  // eps(Float64)

// Julia's min/max propagate NaN (ParticleDistributions.jl:463-470, Coalescence.jl:218); fmin/fmax do not.
__device__ __forceinline__ double jl_min(double a, double b) {
    double r = (b < a) ? b : a;
    return (a != a || b != b) ? (a + b) : r;
}
__device__ __forceinline__ double jl_max(double a, double b) {
    double r = (b > a) ? b : a;
    return (a != a || b != b) ? (a + b) : r;
}

// ---- general-purpose lower incomplete gamma, gamma(a,x) = P(a,x)*Gamma(a), a>0, x>=0 ----------
// Used by the single-object entry points (arbitrary real orders).  Convergence-tested loops.
__device__ inline double igam_lower(double a, double x) {
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : nan("");
    if (isinf(x)) return tgamma(a);
    const double pre = exp(a * log(x) - x);  // x^a e^-x
    if (x < a + 1.0) {
        double t = 1.0 / a, s = t;
        for (int n = 1; n < 2000; ++n) {
            t *= x / (a + n);
            s += t;
            if (t < 1e-17 * s) break;
        }
        return pre * s;
    }
    // modified Lentz on Gamma(a,x) = x^a e^-x / (x+1-a- 1(1-a)/(x+3-a- ...))
    const double tiny = 1e-300;
    double b = x + 1.0 - a;
    double c = 1.0 / tiny;
    double d = 1.0 / b;
    double h = d;
    for (int n = 1; n < 2000; ++n) {
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
    return tgamma(a) - pre * h;
}


// 48 x the weight of 1-based node j in integrate_SimpsonEvenFast (ParticleDistributions.jl:698-710); general n_bins >= 3
__device__ inline double simpson_weight(int j, int n_bins) {
    const int e = n_bins + 1;
    double w = 0.0;
    if (j >= 5 && j <= n_bins - 3) w += 48.0;
    // each end term counts on its own: for n_bins < 8 a node can be a member of two pairs (n_bins = 4: node 3 is both
    // "3" and "e - 2" and weighs 86/48), exactly as the reference's sum does
    if (j == 1) w += 17.0;
    if (j == e) w += 17.0;
    if (j == 2) w += 59.0;
    if (j == e - 1) w += 59.0;
    if (j == 3) w += 43.0;
    if (j == e - 2) w += 43.0;
    if (j == 4) w += 49.0;
    if (j == e - 3) w += 49.0;
    return w;
}

// inverse of the regularised lower incomplete gamma function: x with P(a, x) = p (gamma_inc_inv(a, p, 1-p),
// ParticleDistributions.jl:760).  Start: Wilson-Hilferty for a > 1, the small-x / exponential-tail forms otherwise;
// then Halley iterations on P(a,x) - p with the second-order factor capped at 1 (so the correction never changes
// sign) and a halving step if the iterate would leave x > 0.
__device__ inline double igam_inv(double a, double p) {
    if (!(p > 0.0)) return (p == 0.0) ? 0.0 : nan("");
    if (!(p < 1.0)) return (p == 1.0) ? INFINITY : nan("");
    if (!(a > 0.0)) return nan("");
    const double ga = tgamma(a);
    double x;
    if (a > 1.0) {
        const double t = normcdfinv(p);
        const double w = 1.0 - 1.0 / (9.0 * a) + t / (3.0 * sqrt(a));
        x = fmax(1e-3, a * w * w * w);
    } else {
        const double t = 1.0 - a * (0.253 + a * 0.12);  // P(a, 1) to ~2 digits
        x = (p < t) ? exp(log(p / t) / a) : 1.0 - log(1.0 - (p - t) / (1.0 - t));
        const double xs = exp((log(p) + log(ga * a)) / a);  // leading term of P ~ x^a / Gamma(a+1): exact as p -> 0
        if (xs < 0.05) x = xs;
    }
    if (!(x > 0.0)) x = 1e-300;
    for (int it = 0; it < 100; ++it) {
        const double f = igam_lower(a, x) / ga - p;
        const double fp = exp((a - 1.0) * log(x) - x) / ga;  // dP/dx
        if (!(fp > 0.0) || !isfinite(fp)) {
            // density under/overflow (x far from the root): move by a factor of two towards it
            x = (f < 0.0) ? 2.0 * x : 0.5 * x;
            continue;
        }
        const double u = f / fp;
        const double dx = u / (1.0 - 0.5 * fmin(1.0, u * ((a - 1.0) / x - 1.0)));  // Halley, capped
        if (!(dx == dx)) break;
        double xn = x - dx;
        if (!(xn > 0.0)) xn = 0.5 * x;
        const double step = fabs(xn - x);
        x = xn;
        if (step <= 1e-7 * x) break;  // cubic convergence: the next correction is below 1e-20 x
    }
    return x;
}

// Gamma(k + beta) / Gamma(k) for k in (0, 11], beta in [0, 2): both arguments are shifted to x = k + 12 where the
// Stirling series converges to 1e-17, and the DIFFERENCE of the two log-gammas is formed analytically
// ((x-1/2) log1p(beta/x) + beta ln(x+beta) - beta + tail difference) so that no digits cancel.  ~5x cheaper than two
// tgamma calls; relative error ~1e-15.
__device__ inline double stirling_tail(double x) {  // lgamma(x) - [(x-1/2) ln x - x + ln(2 pi)/2], x >= 12
    const double xi = 1.0 / x, x2 = xi * xi;
    double st = fma(x2, -691.0 / 360360.0, 1.0 / 1188.0);
    st = fma(st, x2, -1.0 / 1680.0);
    st = fma(st, x2, 1.0 / 1260.0);
    st = fma(st, x2, -1.0 / 360.0);
    st = fma(st, x2, 1.0 / 12.0);
    return st * xi;
}
__device__ inline double gamma_ratio(double k, double beta) {
    if (beta == 0.0) return 1.0;
    const double x = k + 12.0, xb = x + beta;
    const double dl = fma(x - 0.5, log1p(beta / x), beta * log(xb)) - beta + (stirling_tail(xb) - stirling_tail(x));
    double num = 1.0, den = 1.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        num *= (k + (double)i);
        den *= (k + beta + (double)i);
    }
    return exp(dl) * (num / den);
}

// Gamma(k + beta)/Gamma(k) from a table: beta is a run constant (a terminal-velocity exponent), so
//   G(k) = Gamma(k + beta + 1)/Gamma(k + 1)   (smooth on k >= 0, nearest pole at k = -1 - beta)
// is stored as piecewise degree-7 polynomials in the local variable t in [-1, 1] on kGrIntervals equal intervals of
// [0, k_max] (built on the host by Chebyshev interpolation, cloudy_config_set), and
//   Gamma(k + beta)/Gamma(k) = k/(k + beta) G(k).
// Interpolation error < 1e-15 relative; ~35 instructions instead of ~150 for gamma_ratio.  Outside the table: gamma_ratio.
constexpr int kGrIntervals = 512;
constexpr int kGrCoef = 8;
__device__ __forceinline__ double gamma_ratio_tab(double k, double beta, const double* __restrict__ coef, double inv_h) {
    if (beta == 0.0) return 1.0;
    const double u = k * inv_h;
    if (!(u >= 0.0 && u < (double)kGrIntervals)) return gamma_ratio(k, beta);
    const int i = (int)u;
    const double t = fma(2.0, u - (double)i, -1.0);
    const double* __restrict__ c = coef + i * kGrCoef;
    double g = __ldg(c + 7);
#pragma unroll
    for (int m = 6; m >= 0; --m) g = fma(g, t, __ldg(c + m));
    return k / (k + beta) * g;
}

// Table-started inverse for a run-constant probability p (MovingThreshold percentiles, Coalescence.jl:152-185): ln x_p(a) is
// tabulated on a uniform grid of a (filled once per configuration by igam_inv itself), interpolated with a 4-point
// Lagrange formula, and polished with Halley steps on P(a,x) - p whose P comes from the division-free series
//   S(x) = sum_n x^n/(a)_{n+1} = R_0 / ((a)_{N+1}),  R_n = x R_{n+1} + prod_{i=n+1..N}(a+i)   (all terms positive).
// Each step cubes the error, so two or three steps reach the rounding level from an interpolation error of 1e-3.
// Outside the table (a < k0, beyond the series regime, p <= 0) the general igam_inv is used.  `ga` = Gamma(a).
constexpr double kXpK0 = 0.25;  // below this shape parameter ln x_p(a) is too steep in a to interpolate
constexpr int kXpN = 544;
__device__ __forceinline__ double igam_inv_guess(double a, const double* __restrict__ tab, int nt, double k0, double inv_h) {
    const double u = (a - k0) * inv_h;
    if (!(u >= 0.0 && u <= (double)(nt - 1))) return nan("");
    int i = (int)u;
    i = max(1, min(i, nt - 3));
    const double t = u - (double)i;
    const double y0 = __ldg(tab + i - 1), y1 = __ldg(tab + i), y2 = __ldg(tab + i + 1), y3 = __ldg(tab + i + 2);
    const double l0 = -t * (t - 1.0) * (t - 2.0) * (1.0 / 6.0), l1 = (t + 1.0) * (t - 1.0) * (t - 2.0) * 0.5;
    const double l2 = -(t + 1.0) * t * (t - 2.0) * 0.5, l3 = (t + 1.0) * t * (t - 1.0) * (1.0 / 6.0);
    return exp(l0 * y0 + l1 * y1 + l2 * y2 + l3 * y3);
}
__device__ inline double igam_inv_tab(double a, double p, double ga, const double* __restrict__ tab, int nt, double k0, double inv_h,
                                      const unsigned char (*deg)[18]) {
    double x = igam_inv_guess(a, tab, nt, k0, inv_h);
    if (!(x > 0.0) || !(p > 0.0) || !(p < 1.0) || !(x < 17.0)) return igam_inv(a, p);
    const int ai = (int)fmin(a, 17.0);
    const double la = log(x);
    double lx = la;
    for (int it = 0; it < 6; ++it) {
        const int zi = (int)fmin(x * 1.05 + 0.5, 25.0);
        int N = deg[zi][ai];
        if (N == 0) return igam_inv(a, p);
        N = min(N + 4, 63);  // the table's truncation bound is for a >= 1; a few more terms cover a in (1/4, 1)
        double R = 1.0, Q = 1.0;
        for (int n = N - 1; n >= 0; --n) {
            Q *= a + (double)(n + 1);
            R = fma(x, R, Q);
        }
        const double fp = exp(fma(a - 1.0, lx, -x)) / ga;  // dP/dx = x^{a-1} e^{-x}/Gamma(a)
        const double f = fp * x * (R / (Q * a)) - p;
        const double u = f / fp;
        const double dx = u / (1.0 - 0.5 * u * ((a - 1.0) / x - 1.0));  // Halley
        if (!(fabs(dx) < 0.5 * x)) return igam_inv(a, p);
        x -= dx;
        if (fabs(dx) <= 1e-6 * x) break;  // the next step's correction would be below 1e-17 x (cubic convergence)
        lx = log(x);
    }
    return x;
}

// Gamma(k) for the Gamma modes' shape parameter, k in (0, 11]: Gamma(k+12) from the Stirling series (tail < 1e-17 at
// x >= 12) divided by the rising factorial (k)_12.  ~3x fewer instructions than tgamma; relative error ~5e-15.
__device__ __forceinline__ double gamma_shape(double k) {
    const double x = k + 12.0;
    const double lg = fma(x - 0.5, log(x), -x) + (0.9189385332046727 + stirling_tail(x));  // ln(2 pi)/2
    double pa = k * (k + 1.0), pb = (k + 2.0) * (k + 3.0);
    pa *= (k + 4.0) * (k + 5.0);
    pb *= (k + 6.0) * (k + 7.0);
    pa *= (k + 8.0) * (k + 9.0);
    pb *= (k + 10.0) * (k + 11.0);
    return exp(lg) / (pa * pb);
}

// 1/Gamma(k) by the same formulas without the division (the FixedThreshold node loop needs n^2/Gamma(k)^2 only)
__device__ __forceinline__ double gamma_shape_inv(double k) {
    const double x = k + 12.0;
    const double lg = fma(x - 0.5, log(x), -x) + (0.9189385332046727 + stirling_tail(x));
    double pa = k * (k + 1.0), pb = (k + 2.0) * (k + 3.0);
    pa *= (k + 4.0) * (k + 5.0);
    pb *= (k + 6.0) * (k + 7.0);
    pa *= (k + 8.0) * (k + 9.0);
    pb *= (k + 10.0) * (k + 11.0);
    return exp(-lg) * (pa * pb);
}

// standard normal CDF
__device__ __forceinline__ double norm_cdf(double t) { return 0.5 * erfc(-t * 0.7071067811865476); }

// ---- tables for the fixed-trip-count evaluation used by the batched kernels -----------------------
// gamma(a,z) at the top order a = k + M' - 1 of a mode:
//   series  gamma(a,z) = z^a e^-z * sum_{n>=0} c_n z^n, c_n = 1/(a)_{n+1}, for z < kSeriesLimit[floor(a)]
//           (Horner, degree kSeriesDeg2[floor(z)][floor(a)] <= 63, truncation <= 2e-16 relative), else
//   Legendre continued fraction of the upper function, forward (Wallis) recurrence at depth
//           kCfDepth[floor(a)] (error <= 1e-16 relative to the LOWER function for every z beyond the limit).
// Generated by a 40-digit mpmath search (DESIGN.md "incomplete gamma"); a <= 18 is required (k <= 11).
constexpr int kSeriesMaxDeg = 63;
constexpr int kCfMaxDepth = 16;
constexpr int kSerZ = 26, kSerA = 18;
static __constant__ unsigned char kSeriesDeg2[kSerZ][kSerA] = {
    {17, 17, 17, 16, 15, 15, 15, 14, 14, 14, 13, 13, 13, 13, 13, 12, 12, 12},
    {22, 22, 21, 20, 20, 19, 19, 18, 18, 17, 17, 17, 16, 16, 16, 15, 15, 15},
    {26, 26, 25, 24, 23, 23, 22, 21, 21, 20, 20, 19, 19, 19, 18, 18, 18, 17},
    {29, 29, 28, 27, 26, 26, 25, 24, 24, 23, 22, 22, 21, 21, 21, 20, 20, 20},
    {32, 32, 31, 30, 29, 28, 28, 27, 26, 26, 25, 24, 24, 23, 23, 22, 22, 22},
    {35, 35, 34, 33, 32, 31, 30, 30, 29, 28, 27, 27, 26, 26, 25, 25, 24, 24},
    {38, 38, 37, 36, 35, 34, 33, 32, 31, 30, 30, 29, 28, 28, 27, 27, 26, 26},
    {40, 40, 39, 38, 37, 36, 35, 34, 34, 33, 32, 31, 31, 30, 29, 29, 28, 28},
    {43, 43, 42, 41, 40, 39, 38, 37, 36, 35, 34, 33, 33, 32, 31, 31, 30, 29},
    {45, 45, 44, 43, 42, 41, 40, 39, 38, 37, 36, 36, 35, 34, 33, 33, 32, 31},
    {47, 47, 46, 45, 44, 43, 42, 41, 40, 40, 39, 38, 37, 36, 35, 35, 34, 33},
    {50, 50, 49, 48, 47, 46, 45, 44, 43, 42, 41, 40, 39, 38, 37, 37, 36, 35},
    {52, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43, 42, 41, 40, 39, 39, 38, 37},
    {54, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43, 42, 41, 41, 40, 39},
    {56, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43, 43, 42, 41},
    {58, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 44, 43},
    {60, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 46, 45},
    {62, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 47},
    { 0,  0, 63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48},
    { 0,  0,  0,  0, 63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50},
    { 0,  0,  0,  0,  0,  0, 63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52},
    { 0,  0,  0,  0,  0,  0,  0,  0, 63, 62, 61, 60, 59, 58, 57, 56, 55, 54},
    { 0,  0,  0,  0,  0,  0,  0,  0,  0,  0, 63, 62, 61, 60, 59, 58, 57, 56},
    { 0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0, 63, 62, 61, 60, 59, 58},
    { 0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0, 63, 62, 61, 60, 59},
    { 0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0, 63, 62, 61},
};
static __constant__ double kSeriesLimit[kSerA] = {18.0, 18.0, 19.0, 19.0, 20.0, 20.0, 21.0, 21.0, 22.0, 22.0, 23.0, 23.0, 24.0, 25.0, 25.0, 26.0, 26.0, 26.0};
static __constant__ int kCfDepth[kSerA] = {4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 10, 11, 11, 12, 13, 13, 14, 15};

// Depth of the same continued fraction as a function of z: rows floor(a) = 0..17, columns z in [limit + 4b, limit + 4b + 4), last
// column everything beyond.  Smallest depth with |Q_d - Q| <= 1e-16 (error relative to the lower function), generated by
// tools/gen_cf_depth.py (40-digit mpmath); the kernels add one level of margin to every non-zero entry.  The fraction
// converges quickly once z >> a: beyond limit + 16 half the depth of kCfDepth is enough, beyond limit + 44 none.
constexpr int kCfZBins = 16;
static __constant__ unsigned char kCfDepthZ[kSerA][kCfZBins] = {
    { 3,  2,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 3,  2,  1,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 4,  3,  2,  1,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 4,  3,  3,  2,  1,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 5,  4,  3,  3,  2,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 5,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 6,  5,  4,  3,  3,  2,  1,  0,  0,  0,  0,  0,  0,  0,  0,  0},
    { 7,  6,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0,  0,  0,  0,  0},
    { 8,  7,  6,  4,  3,  3,  2,  1,  0,  0,  0,  0,  0,  0,  0,  0},
    { 9,  7,  6,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0,  0,  0,  0},
    { 9,  8,  7,  5,  4,  3,  2,  2,  1,  0,  0,  0,  0,  0,  0,  0},
    {10,  9,  7,  6,  5,  4,  3,  2,  1,  0,  0,  0,  0,  0,  0,  0},
    {10,  9,  8,  6,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0,  0,  0},
    {11,  9,  8,  7,  5,  4,  3,  2,  2,  1,  0,  0,  0,  0,  0,  0},
    {12, 10,  9,  7,  6,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0,  0},
    {12, 11,  9,  8,  6,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0,  0},
    {13, 11, 10,  8,  7,  6,  5,  4,  3,  2,  1,  0,  0,  0,  0,  0},
    {14, 12, 11,  9,  8,  6,  5,  4,  3,  2,  1,  1,  0,  0,  0,  0},
};

__device__ __forceinline__ int series_z_bin(double zmax) { return (zmax >= 0.0) ? (int)fmin(zmax, (double)(kSerZ - 1)) : 0; }
__device__ __forceinline__ int series_a_bin(double a_top) { return (int)fmin(fmax(a_top, 0.0), (double)(kSerA - 1)); }

}  // namespace cloudy
