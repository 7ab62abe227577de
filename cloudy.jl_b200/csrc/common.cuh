// Shared device-side definitions of libcloudy_b200: run-constant configuration, kernel arguments and the
// per-mode closed forms (parameters from moments, moments from parameters).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/cloudy_b200.h"
#include "special.cuh"

namespace cloudy {

// ------------------------------------------------------------------------------------------------
// device-side configuration (kernel parameter, lives in the constant bank)
// ------------------------------------------------------------------------------------------------
constexpr int MAXN = CLOUDY_MAX_MODES;
constexpr int MAXP = CLOUDY_MAX_P;
constexpr int MAXM = MAXP + 2;
constexpr int MAXSLOT = CLOUDY_MAX_SLOTS;
constexpr int MAXT = MAXM * (MAXM + 1) / 2;  // 28
constexpr int kZtN = 1024;         // Z-sum tables (cloudy_config_set): intervals of [0, kZtKmax] in k
constexpr double kZtKmax = 11.0;   // = the largest supported Gamma shape

struct DevConfig {
    int N, P, M, nslots;
    int kind[MAXN], nprog[MAXN], slot0[MAXN];
    int slot_mode[MAXSLOT], slot_order[MAXSLOT];
    int thr_style, n_mom_max;
    int n2d[MAXN], Mp[MAXN];       // Mp = min(M, n2d): orders 0..Mp-1 carry truncated integrals
    int quad[MAXN];                // 1: Gamma/Exponential mode with finite threshold, not last → node loop
    int mono_thr[MAXN];            // 1: Monodisperse mode with finite threshold, not last → closed form
    int ln_thr[MAXN];              // 1: Lognormal mode with finite threshold, not last → Gauss-Legendre quadrature
    int gl_off, gl_n;              // Gauss-Legendre nodes/weights on [-1,1] inside `tab` (nodes then weights)
    int bins_per_log_unit;         // MovingThreshold: per-parcel grid density (15 in the reference)
    int n_bins[MAXN], tab_off[MAXN];
    int rec_off[MAXN], rec_near[MAXN], rec_far[MAXN];  // packed node records of the thread-per-parcel kernel (tpp_kernel.cuh)
    int xp_off[MAXN], xp_n;        // MovingThreshold: per Gamma mode, ln x_p(k) on a uniform k grid inside `tab` (special.cuh igam_inv_tab)
    int tab_total;                 // doubles of SoA grid tables the lane-cooperative kernel stages in shared memory
    int tpp_off, tpp_total;        // region of `tab` the MovingThreshold thread-per-parcel kernels stage (node records, Gauss-Legendre rule)
    // region the FixedThreshold thread-per-parcel kernels stage: [Gauss-Legendre rule | 16-byte aligned node records | block degrees]
    int tpp2_off, tpp2_total, gl2_off;
    int rec2_off[MAXN], kblk2_off[MAXN];  // per quadrature mode: records (tmx, lsum, w_0..w_P, pad) and Taylor degree per node block
    int rec2_far[MAXN];                   // far nodes padded to a multiple of TPP_NPLF, at least one zero-weight dummy at the end
    int zt_off[MAXN], zt_n;               // per quadrature mode: Z-sum polynomials in k inside `tab` ([interval][t(p1,p)][8], global memory)
    int near_cls_end[MAXN][5];            // near blocks [0, end[c]) have a Taylor degree class <= c (classes kTaylorClass, tpp_kernel.cuh)
    int n_vel, nz;
    double c[MAXN][MAXN][MAXP][MAXP];
    double sw[MAXN][3][MAXT];      // folded S-term weights by (mode, order m, t(u,v) of the M x M triangle), see cloudy_config_set
    double thr[MAXN];
    double norm[MAXSLOT];
    double inv_norm[MAXSLOT];      // RN(1/norm) for norm_div
    double k_lo, k_hi;
    double zt_L[MAXN], zt_inv_h;   // Z-sum tables: exponent reference L = max_j ls_j, intervals per unit of k
    double xp_k0, xp_inv_h;
    double velv[CLOUDY_MAX_VEL], velb[CLOUDY_MAX_VEL];  // v*norms[2]^beta, beta
    double gam_b1[CLOUDY_MAX_VEL];                      // Γ(1 + beta): Exponential modes' fractional moment
    int gr_off[CLOUDY_MAX_VEL];                         // per velocity term: Γ(k+β+1)/Γ(k+1) polynomial table inside `tab` (special.cuh gamma_ratio_tab)
    double gr_inv_h;                                    // intervals per unit of k
    double inv_dz_unused, dz;
    const double* tab;  // device: per quad mode i at tab_off[i]: XJ[n] ELL[n] TMX[n] LZ[n] W[M][n]
};

struct KArgs {
    const double* u_in;   // state the RHS is evaluated at
    const double* u_n;    // u^n for stages 2,3 (nullptr otherwise)
    double* out;          // tendency, flux or stage result
    double* clip_back;    // rainshaft tendency call: clipped state written back (nullptr otherwise)
    long long n;          // parcels / cells
    long long s_in, s_n, s_out, s_clip;  // slot strides (doubles): SoA = ensemble stride, AoS = 1
    long long ps_in, ps_out;             // parcel strides of u_in / out: SoA = 1, AoS (host layout) = n_slots
    double cn, ci, cf, dt, div;          // out = (cn*u_n + ci*u_in + cf*(dt*f))/div ; tend_only: out = f
    int tend_only;
    int flux_only;        // out = sedimentation flux (rainshaft_helpers.jl:77)
    int params_in;        // u_in holds distribution parameters (n, θ|μ[, k|σ]) instead of moments; output not de-normalised
    unsigned long long* err_count;
    // thread-per-parcel kernels only
    const int* perm;      // processing order (regime-sorted parcel indices) or nullptr for identity
    const double* flux;   // rainshaft: per-cell sedimentation flux (SoA like the state), written by flux_kernel
    long long s_flux;
    unsigned long long* tile_ctr;   // dynamic tile schedule of the thread-per-parcel kernel: global draw counter (never reset) ...
    unsigned long long tile_base;   // ... and its value when the launch starts (the host advances it by n_tiles + n_warps per launch)
    int presorted;        // host side only: the ensemble is resident in regime order (or must keep its order): no permutation sort
};

enum { MODEL_BOX = 0, MODEL_RAINSHAFT = 1, MODEL_BOX_MOVING = 2 };

// ------------------------------------------------------------------------------------------------
// distribution parameters from normalised moments — ParticleDistributions.jl:456-541 — and the
// moment matrix row of one mode — Coalescence.jl:187-198 / ParticleDistributions.jl:177-207
// ------------------------------------------------------------------------------------------------
struct ModeParams {
    double n, a, b;  // (n, θ, k) or (n, μ, σ); b = 1 for Exponential/Monodisperse
    int invalid;
};

// Lognormal branch of params_from_moments, kept OUT OF LINE: the hot kernels carry it for every mode (the kind is a run-time
// value) although the BASELINE configurations never execute it, and its two logarithms, square roots and exponential would
// otherwise sit in the middle of the instruction stream (instruction-cache footprint)
static __device__ __noinline__ ModeParams lognormal_params_from_moments(double m0, double m1, double m2, double lo, double hi, double lo2, double hi2) {
    ModeParams r;
    r.invalid = 0;
    if (m0 > kEps && m1 > kEps && m2 > kEps) {
        // lo/hi clamp μ, lo2/hi2 clamp σ (reference defaults (-Inf, Inf), (eps, Inf))
        double mu = jl_max(lo, jl_min(hi, log(m1 * m1 / (m0 * sqrt(m0)) / sqrt(m2))));
        double arg = log(m0 * m2 / (m1 * m1));
        if (arg < 0.0) r.invalid = 1;  // the reference throws a DomainError here (:498)
        double sg = jl_max(lo2, jl_min(hi2, sqrt(arg)));
        r.a = mu;
        r.b = sg;
        r.n = m1 / exp(mu + 0.5 * sg * sg);
    } else {
        r.n = 0.0; r.a = 1.0; r.b = 1.0;
    }
    return r;
}
// v / nrm with the run-constant reciprocal rinv = RN(1/nrm): quotient estimate, exact remainder, one correction — the correctly
// rounded quotient (Markstein) in 3 FP64 instructions instead of the ~12-instruction dependent chain of an inline IEEE division
// with its slow-path call (the five normalisations per parcel were 4 % of the C5 kernel's stall samples).  Inf and NaN pass through.
__device__ __forceinline__ double norm_div(double v, double nrm, double rinv) {
    const double q = v * rinv;
    const double r = fma(-q, nrm, v);
    const double q1 = fma(r, rinv, q);
    return (fabs(q) < 1.7976931348623157e308) ? q1 : q;
}
// IEEE division kept out of line where it is evaluated once per parcel and slot (stage update): every inlined FP64 division
// costs ~100 instructions of code with its slow path
static __device__ __noinline__ double div_rn_outofline(double a, double b) { return a / b; }
// moment(Lognormal, q) for integer q (moment matrix), out of line for the same reason
static __device__ __noinline__ double lognormal_moment_int(double n, double mu, double sg, int q) { return n * exp(q * mu + (double)(q * q) * sg * sg / 2); }

__device__ inline ModeParams params_from_moments(int kind, double m0, double m1, double m2, double lo, double hi,
                                                 double lo2 = kEps, double hi2 = INFINITY) {
    ModeParams r;
    r.invalid = 0;
    if (kind == CLOUDY_GAMMA) {
        if (m0 > kEps && m1 > kEps) {
            r.n = m0;
            double mean = m1 / m0;
            double k = jl_max(lo, jl_min(hi, mean / (m2 / m1 - mean)));
            r.b = k;
            r.a = mean / k;
        } else {
            r.n = 0.0; r.a = 1.0; r.b = 1.0;
        }
    } else if (kind == CLOUDY_LOGNORMAL) {
        r = lognormal_params_from_moments(m0, m1, m2, lo, hi, lo2, hi2);
    } else {  // Exponential, Monodisperse
        if (m0 > kEps && m1 > kEps) {
            r.n = m0; r.a = m1 / m0; r.b = 1.0;
        } else {
            r.n = 0.0; r.a = 1.0; r.b = 1.0;
        }
    }
    return r;
}

// moment(dist, q) for real q — ParticleDistributions.jl:177-207
__device__ inline double moment_real(int kind, double n, double a, double b, double q) {
    switch (kind) {
        case CLOUDY_EXPONENTIAL: return n * pow(a, q) * tgamma(q + 1.0);
        case CLOUDY_GAMMA: return n * pow(a, q) * tgamma(q + b) / tgamma(b);
        case CLOUDY_MONODISPERSE: return n * pow(a, q);
        default: return n * exp(q * a + q * q * b * b / 2);
    }
}

// Sedimentation flux of one cell from its distribution parameters — Sedimentation.jl:22-37 with the velocity normalisation of
// rainshaft_helpers.jl:74-77: flux[i][q] = s_q * (-sum_l v_l x0^beta_l moment(pdist_i, q + beta_l)).  moment(dist, q+β) for
// q = 0,1,2 follows from the q = 0 value by Γ(x+1) = xΓ(x).  An empty mode (n = 0, the fallback of update_dist_from_moments)
// carries no flux.  Used by flux_kernel (cloudy_sedimentation_flux) and, for a cell and the cell above it, inside the
// rainshaft instances of the thread-per-parcel kernel.
template <int NM>
__device__ __forceinline__ void cell_flux(const DevConfig& cfg, const double (&pn)[NM], const double (&pa)[NM], const double (&pb)[NM],
                                          double (&fl)[NM][3]) {
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        fl[i][0] = fl[i][1] = fl[i][2] = 0.0;
        if (i >= cfg.N) continue;
        const int np = cfg.nprog[i], kind = cfg.kind[i];
        if (pn[i] != 0.0) {
            const double log_a = (kind == CLOUDY_LOGNORMAL) ? 0.0 : log(pa[i]);
            for (int v = 0; v < cfg.n_vel; ++v) {
                const double beta = cfg.velb[v];
                // moment(dist, beta): n θ^β Γ(β+k)/Γ(k) | n θ^β Γ(β+1) | n θ^β  (ParticleDistributions.jl:177-199)
                double mq = 0.0;
                if (kind == CLOUDY_GAMMA) mq = pn[i] * exp(beta * log_a) * gamma_ratio_tab(pb[i], beta, cfg.tab + cfg.gr_off[v], cfg.gr_inv_h);
                else if (kind == CLOUDY_EXPONENTIAL) mq = pn[i] * exp(beta * log_a) * cfg.gam_b1[v];
                else if (kind == CLOUDY_MONODISPERSE) mq = pn[i] * exp(beta * log_a);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q < np) {
                        if (kind == CLOUDY_LOGNORMAL) mq = moment_real(kind, pn[i], pa[i], pb[i], (double)q + beta);
                        fl[i][q] += -cfg.velv[v] * mq;
                        if (kind == CLOUDY_GAMMA) mq *= pa[i] * (pb[i] + beta + q);
                        else if (kind == CLOUDY_EXPONENTIAL) mq *= pa[i] * (beta + q + 1.0);
                        else mq *= pa[i];
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q < np) fl[i][q] *= cfg.norm[cfg.slot0[i] + q];
    }
}

}  // namespace cloudy
