// Thread-per-parcel kernel (the fast path for the mode/tensor shapes listed in tpp_instances.cu).
//
// One thread evaluates one parcel's (or one column cell's) coalescence tendency end to end:
//   rhs_coal!                test/examples/utils/box_model_helpers.jl:29-53
//   get_coal_ints            src/Sources/Coalescence.jl:115-150 (+ :187-455 underneath)
//   update_dist_from_moments src/ParticleDistributions/ParticleDistributions.jl:456-541
//   moment_source_helper     src/ParticleDistributions/ParticleDistributions.jl:567-612 (+ Simpson :698-710)
// with everything that is per-parcel (distribution parameters, moment matrix, truncated integrals F, the
// Q/R/S contraction) held in registers — N (modes) and P (tensor size) are template parameters so every
// loop over them unrolls.  All threads of a warp walk the SAME quadrature nodes at the same time, so the
// node tables are shared-memory broadcasts; the only per-thread table is the parcel's series coefficients
// c_n = 1/(a)_{n+1} (column `tid` of a [64][threads] shared array, conflict-free).  NPL nodes are in flight
// per thread: their Horner chains share each coefficient load and hide the DFMA latency.
// Warps are made homogeneous (similar series length, same series/continued-fraction regime) by the regime order of
// the ensemble (data sort of resident box ensembles, permutation args.perm for column states), and the work is handed
// out in tiles drawn from a device counter (see the parcel loop of tpp_kernel).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace cloudy {

#ifndef TPP_THREADS_N
#define TPP_THREADS_N 128
#endif
constexpr int TPP_THREADS = TPP_THREADS_N;
// nodes in flight per thread (measured on C2, P = 2: 2 → 0.70 ms, 3 → 0.66, 4 → 0.74, 5 → 0.73); high-order tensors carry
// up to 28 accumulators per node set, so they keep fewer nodes in flight
#ifndef TPP_NPL_LARGE
#define TPP_NPL_LARGE 3
#endif
#ifndef TPP_NPL_SMALL
#define TPP_NPL_SMALL 3
#endif
__host__ __device__ constexpr int tpp_npl(int P) { return P >= 4 ? TPP_NPL_LARGE : TPP_NPL_SMALL; }  // (2 for P >= 4 measured within noise: +4 % at 1 Mi, -4 % at 16 Mi parcels on C4)
// resident blocks per SM the register allocation must allow (168 registers for 3 blocks of 128 threads): the small shapes
// fit, the others take up to 255 registers and run 2 blocks
#ifndef TPP_MINB_SMALL
#define TPP_MINB_SMALL 4
#endif
#ifndef TPP_MINB_LARGE
#define TPP_MINB_LARGE 2
#endif
__host__ __device__ constexpr int tpp_min_blocks(int N, int P, int model) {
    return (model == MODEL_BOX_MOVING ? 2 : ((N * P <= 4 && N <= 3) ? TPP_MINB_SMALL : TPP_MINB_LARGE)) * 128 / TPP_THREADS;
}
constexpr int TPP_CT_ROWS = 64;  // MovingThreshold instances: series coefficients c_0..c_63 (the Taylor coefficients overwrite them)
// FixedThreshold instances keep only the Taylor coefficients t_0..t_26 in shared memory (the far-zone series runs on
// coefficients generated on the fly), 216 bytes per thread
constexpr int TPP_CT_ROWS_FIXED = 27;
constexpr int TPP_NPLF = 5;       // far-zone nodes in flight per thread (FixedThreshold): one coefficient product serves five Horner chains
__host__ __device__ constexpr int tpp_ct_rows(int model) { return model == 2 /*MODEL_BOX_MOVING*/ ? TPP_CT_ROWS : TPP_CT_ROWS_FIXED; }
constexpr int TPP_TAYLOR_MAX = 26;
// dynamic tile schedule (tpp_kernel, launch_tpp): box instances draw one TPP_THREADS-parcel tile per block and iteration,
// column instances one 32-parcel tile per warp; every drawer draws once beyond the last tile
__host__ __device__ constexpr bool tpp_block_sync(int model) { return model != 1 /*MODEL_RAINSHAFT*/; }
__host__ __device__ constexpr int tpp_tile(int model) { return tpp_block_sync(model) ? TPP_THREADS : 32; }
__host__ __device__ constexpr int tpp_overdraw(int model) { return tpp_block_sync(model) ? 1 : TPP_THREADS / 32; }
// Lower-order sums Z[p1][p] of the FixedThreshold node loop: from the k-tables of cloudy_config_set (small tensors: 6 records of
// 64 bytes per parcel and mode) or accumulated node by node (P >= 4: 21 records per parcel and mode overflow the L1 and cost
// C4 8 % more than the 21 multiply-adds per node they replace — measured)
__host__ __device__ constexpr bool tpp_ztab(int P) { return P < 4; }
// FixedThreshold near zone: the node-only Taylor degrees (4..26) are rounded up to five classes; the near blocks are sorted by
// degree, so each class is a run of consecutive blocks and gets its own loop with a compile-time degree (no per-block dispatch)
constexpr int TPP_N_CLASSES = 5;
__host__ __device__ constexpr int tpp_taylor_class(int c) { return c == 0 ? 5 : c == 1 ? 7 : c == 2 ? 11 : c == 3 ? 16 : TPP_TAYLOR_MAX; }  // Taylor coefficients t_0..t_26 of the near-node expansion

__host__ __device__ constexpr int tri_ct(int p1, int p2, int MP) { return p1 * MP - (p1 * (p1 - 1)) / 2 + (p2 - p1); }

// exp(x) for the node weights: x = n ln2/256 + r, exp(x) = 2^(n>>8) * 2^((n&255)/256) * e^r with a 256-entry table in shared
// memory and a polynomial on |r| <= ln2/512 (default: degree 3, relative error <= 2e-13; -DTPP_EXP_FULL: degree 4 and a
// two-constant reduction, <= 1.4 ulp).  Valid for
// -5e6 < x <= 700 (callers bound the exponent offset); results below 2^-1022 are returned as ~2^-1022 instead of 0, which
// is far below anything the sums can resolve.  ~10 FP64 instructions instead of ~20 for the library exp, and only four
// constants that do not fit an immediate operand.
constexpr int TPP_EXP_TAB = 256;
template <bool FULL>
__device__ __forceinline__ double fast_exp_t(double x, const double* __restrict__ tab) {
    const double t = fma(x, 369.3299304675746, 6755399441055744.0);  // 256/ln2, 1.5*2^52
    int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double p;
    if constexpr (!FULL) {
        // accuracy target 2e-13 (parity is asserted at 1e-9; measured worst tendency error against the oracle over 262144
        // wide-parameter parcels 4.7e-14 of the term scale, 8e-15 with -DTPP_EXP_FULL): one-constant reduction (rounding error
        // 1.1e-16 |x|, arguments here are within [-745, 60]) and a degree-3 polynomial on |r| <= ln2/512 (truncation r^4/24 <= 1.4e-13).
        // 7 FP64 instructions instead of 9: -2.8 % on C5.
        const double r = fma(nf, -0.0027076061740622863, x);   // ln2/256 to 53 bits
        p = fma(r, 1.6666666666666666e-01, 0.5);
        p = fma(p, r, 1.0);
        p = fma(p, r, 1.0);
    } else {
        double r = fma(nf, -0.002707604318857193, x);   // ln2/256, high 21 bits (nf * hi is exact)
        r = fma(nf, -1.855205093371747e-09, r);          // ln2/256, low part
        p = fma(r, 4.1666666666666664e-02, 1.6666666666666666e-01);
        p = fma(p, r, 0.5);
        p = fma(p, r, 1.0);
        p = fma(p, r, 1.0);
    }
    p *= tab[n & (TPP_EXP_TAB - 1)];
    n = max(n, -1022 * TPP_EXP_TAB);
    return __hiloint2double(__double2hiint(p) + ((n >> 8) << 20), __double2loint(p));
}
#ifndef TPP_EXP_FULL
__device__ __forceinline__ double fast_exp(double x, const double* __restrict__ tab) { return fast_exp_t<false>(x, tab); }
#else
__device__ __forceinline__ double fast_exp(double x, const double* __restrict__ tab) { return fast_exp_t<true>(x, tab); }
#endif
constexpr double kExpOffsetMin = -1.0e6;  // lower clamp of per-parcel exponent offsets fed to fast_exp

// ------------------------------------------------------------------------------------------------
// node grids of the reference's log-spaced Simpson rule (ParticleDistributions.jl:579-585, :698-710)
// ------------------------------------------------------------------------------------------------
// Both grids split their nodes into a NEAR zone (x_j/x_th <= 0.1) and a FAR zone, each a whole number of blocks of
// tpp_npl(P) nodes, and hand out opaque node handles.
//
// Node records of the table grid, packed per node so that one base address serves all loads:
//   [0] x_th - x_j   [1] ln x_j + ln(x_th - x_j)   [2] ln x_j   [3] x_j   [4] Taylor degree K_j   [5+p] w_j dx x_j^p
// The host orders them [near nodes | far nodes], each zone padded to a multiple of tpp_npl(P) with zero-weight dummy
// nodes, so the hot loop needs no index clamps and no validity selects.
constexpr int REC_TMX = 0, REC_LSUM = 1, REC_ELL = 2, REC_X = 3, REC_K = 4, REC_W = 5;
// FixedThreshold kernels: 16-byte aligned records (x_th - x_j, ln x_j + ln(x_th - x_j), w_j dx x_j^p for p = 0..P, padding to an
// even count) read with 128-bit shared-memory loads; the Taylor degree of each node block sits in a separate array
__host__ __device__ constexpr int tpp_rec2_stride(int P) { return 2 + ((P + 2) & ~1); }

struct TableGrid {  // FixedThreshold (and MovingThreshold's unit grid): built on the host, broadcast from shared memory
    const double* rec;
    int stride, n_near, n_far;  // padded node counts
    __device__ __forceinline__ int near_count() const { return n_near; }
    __device__ __forceinline__ int far_count() const { return n_far; }
    __device__ __forceinline__ int near_handle(int jj) const { return jj; }
    __device__ __forceinline__ int far_handle(int jj) const { return n_near + jj; }
    __device__ __forceinline__ int block_degree(int h0, int npl) const { return (int)rec[(h0 + npl - 1) * stride + REC_K]; }
    __device__ __forceinline__ bool block_interior(int, int) const { return true; }
    __device__ __forceinline__ void load(int h, double& tmx, double& aux) const { tmx = rec[h * stride + REC_TMX]; aux = 0.0; }
    __device__ __forceinline__ double log_sum(int h, double) const { return rec[h * stride + REC_LSUM]; }
    __device__ __forceinline__ double ell(int h) const { return rec[h * stride + REC_ELL]; }
    __device__ __forceinline__ double x(int h) const { return rec[h * stride + REC_X]; }
    __device__ __forceinline__ double o_tmx(int j) const { return rec[j * stride + REC_TMX]; }
    __device__ __forceinline__ double o_log_sum(int j) const { return rec[j * stride + REC_LSUM]; }
    __device__ __forceinline__ int o_degree(int j) const { return (int)rec[j * stride + REC_K]; }
    template <int MP>
    __device__ __forceinline__ void weights(int h, double, bool, double (&w)[MP]) const {
#pragma unroll
        for (int p = 0; p < MP; ++p) w[p] = rec[h * stride + REC_W + p];  // w_j dx x_j^p
    }
};

// MovingThreshold with x_th > 1: x_lb = 1e-5 is absolute, so the number of nodes nb and the log spacing dx depend on the
// parcel (ParticleDistributions.jl:579-582).  Lengths are measured in units of the parcel's threshold and nodes are
// counted from the top: node m = 1..nb sits at ρ_m = x/x_th = exp(-m dx) (m = 0 is the threshold itself, where the
// integrand vanishes).  dx >= ln10/bins_per_log_unit, hence ρ_m <= 10^(-m/bpl): the Taylor degree and the length of the
// log1p series below are functions of m alone (warp-uniform), exactly as on the table grid.
struct OwnGrid {
    double dx;
    int nb;                      // the parcel's own node count
    int n_near_w, n_far_w;       // warp-uniform padded zone sizes: far m = 1..n_far_w, near m = n_far_w+1..n_far_w+n_near_w
    int nb_min_w;                // smallest node count in the warp
    int m8, m4;                  // ρ_m <= 1e-2 from m8 on, <= 1e-4 from m4 on
    const double* exp_tab;
    const unsigned char* kdeg_m;  // Taylor degree bound by m (shared memory, 128 entries, non-increasing)
    __device__ __forceinline__ int near_count() const { return n_near_w; }
    __device__ __forceinline__ int far_count() const { return n_far_w; }
    __device__ __forceinline__ int near_handle(int jj) const { return n_far_w + 1 + jj; }
    __device__ __forceinline__ int far_handle(int jj) const { return 1 + jj; }
    __device__ __forceinline__ int block_degree(int m0, int) const { return (int)kdeg_m[min(m0, 127)]; }
    // every node of the block [m0, m0+npl) has Simpson weight 1 in every parcel of the warp
    __device__ __forceinline__ bool block_interior(int m0, int npl) const { return m0 >= 4 && m0 + npl - 1 <= nb_min_w - 4; }
    __device__ __forceinline__ double ell(int m) const { return -(double)m * dx; }  // ln ρ_m
    __device__ __forceinline__ double x(int m) const { return exp(ell(m)); }
    __device__ __forceinline__ void load(int m, double& tmx, double& rho) const {
        rho = fast_exp(ell(m), exp_tab);
        tmx = 1.0 - rho;
    }
    __device__ __forceinline__ double log_sum(int m, double rho) const {  // ln ρ + ln(1 - ρ)
        double l1p;
        if (m <= n_far_w) {
            l1p = log(1.0 - rho);
        } else {
            // -ln(1-ρ)/ρ = sum_n ρ^n/(n+1), truncated below 1e-17 for the zone's largest ρ
            double q;
            if (m >= m4) {
                q = fma(rho, 0.25, 1.0 / 3.0);
            } else {
                if (m >= m8) {
                    q = fma(rho, 0.125, 1.0 / 7.0);
                } else {
                    q = fma(rho, 1.0 / 16.0, 1.0 / 15.0);
                    q = fma(q, rho, 1.0 / 14.0);
                    q = fma(q, rho, 1.0 / 13.0);
                    q = fma(q, rho, 1.0 / 12.0);
                    q = fma(q, rho, 1.0 / 11.0);
                    q = fma(q, rho, 1.0 / 10.0);
                    q = fma(q, rho, 1.0 / 9.0);
                    q = fma(q, rho, 1.0 / 8.0);
                    q = fma(q, rho, 1.0 / 7.0);
                }
                q = fma(q, rho, 1.0 / 6.0);
                q = fma(q, rho, 0.2);
                q = fma(q, rho, 0.25);
                q = fma(q, rho, 1.0 / 3.0);
            }
            q = fma(q, rho, 0.5);
            q = fma(q, rho, 1.0);
            l1p = -rho * q;
        }
        return ell(m) + l1p;
    }
    template <int MP>
    __device__ __forceinline__ void weights(int m, double rho, bool interior, double (&w)[MP]) const {
        // node m is the reference's 1-based node j = nb - m + 1; weight dx * ρ^p (x_th^p is applied by the caller)
        w[0] = interior ? dx : ((m <= nb) ? simpson_weight(nb - m + 1, nb) * (dx / 48.0) : 0.0);
#pragma unroll
        for (int p = 1; p < MP; ++p) w[p] = w[p - 1] * rho;
    }
};

// ------------------------------------------------------------------------------------------------
// node loop for one mode of one parcel: acc[t(p1,p2)] = sum_j W[p1][j] g_j gamma(k+p2, z_j)
//   g_j = (x_j/θ)^k e^{-x_j/θ},  z_j = (x_th - x_j)/θ,  E_j = z_j^k e^{-z_j}
// Series regime (z below the series limit): gamma(k+p, z) = E h_p with h_top = z^{MP-1} S(z),
//   h_p = (h_{p+1} + z^p)/(k+p), and g_j E_j = exp(k (ln x_j + ln(x_th - x_j) - 2 ln θ) - x_th/θ) — ONE exponential.
//   S(z) = sum_n c_n z^n = gamma(a,z) z^-a e^z, a = k+MP-1:
//     FAR nodes (x_j/x_th > 0.1, the last ~14 of 75): Horner on the parcel's c_n table at z_j;
//     NEAR nodes cluster just below X_c = min(x_th/θ, series limit - 0.5): S obeys z S' = (z - a) S + 1, so its Taylor
//       coefficients about X_c in r = z/X_c - 1 follow from S(X_c) alone,
//         t_0 = S(X_c), t_1 = (X_c - a) t_0 + 1, t_{m+1} = [(X_c - a - m) t_m + X_c t_{m-1}]/(m+1),
//       and S(z_j) = sum_{m<=K_j} t_m r^m with a node-only degree K_j = 4..25 (instead of 40-60 series terms).  The t_m
//       overwrite the parcel's table column once the far nodes are done.  |r| <= 0.115 and |z - X_c| <= 2.6 bound the
//       alternating cancellation to ~1e-14 (degrees from a 40-digit search).
// Continued-fraction regime (z beyond the series limit, i.e. the threshold far in the tail): the main loop uses
//   h_top = 0 there and a separate, compact loop adds B_p (g Γ(a) - g E z^{MP-1} Q/P), B_p = prod_{q>=p} 1/(k+q), with
//   Q/P the Legendre continued fraction of the upper function (forward recurrence, fixed depth) and g_j on its own.
// Lengths are in the grid's unit: θ for the FixedThreshold tables (zs = 1/θ, log_u = ln θ), or the parcel's threshold for
// the MovingThreshold grids (zs = x_th/θ = X, log_u = -ln X; the caller multiplies acc[t(p1,p2)] by x_th^p1).
// MASKED (MovingThreshold only): a warp may hold parcels of both grid kinds; each kind's zones run over the whole warp
// and add exact zeros for the threads with `on == false`, so a parcel's result does not depend on its warp-mates.
// ------------------------------------------------------------------------------------------------
template <int MP>
struct NodeCommon {
    double k, zs, log_u, X, gam_top, a_top, ser_lim;
    double e0, Xc, rq, cf_lim;
    bool capped, warp_capped, warp_cf;
    int deg_w, cfd_w, cfd;
    double ia[MP];
    double* myCt;
    const double* exp_tab;
    __device__ __forceinline__ void init() {
        e0 = fmax(fma(-2.0 * k, log_u, -X), kExpOffsetMin);  // exponent offset of g*E
        Xc = fmin(X, ser_lim - 0.5);          // Taylor centre (inside the series regime)
        rq = zs / Xc;                         // r = z/X_c - 1 = (x_th - x_j) rq - 1
        capped = !(X <= ser_lim - 0.5);       // centre below x_th/θ: r does not vanish at the first nodes
        warp_capped = __any_sync(0xffffffffu, capped);
        warp_cf = __any_sync(0xffffffffu, !(X < ser_lim));  // any parcel of the warp with continued-fraction nodes
        cf_lim = warp_cf ? ser_lim : INFINITY;
    }
};

// S(X_c) from the c_n table, then its Taylor coefficients into the same column (after the far zones, before the near ones)
template <int MP>
__device__ __forceinline__ void tpp_taylor_coeffs(const NodeCommon<MP>& C) {
    double* __restrict__ myCt = C.myCt;
    double s0 = myCt[C.deg_w * TPP_THREADS];
    for (int n = C.deg_w - 1; n >= 0; --n) s0 = fma(s0, C.Xc, myCt[n * TPP_THREADS]);
    const double Xa = C.Xc - C.a_top;
    double tm1 = s0, tm = fma(Xa, s0, 1.0);
    myCt[0] = tm1;
    myCt[TPP_THREADS] = tm;
#pragma unroll
    for (int m = 1; m < TPP_TAYLOR_MAX; ++m) {
        const double tn = fma(Xa - (double)m, tm, C.Xc * tm1) * (1.0 / (double)(m + 1));
        myCt[(m + 1) * TPP_THREADS] = tn;
        tm1 = tm;
        tm = tn;
    }
}

template <int MP, int P, bool NEAR, bool MASKED, typename Grid>
__device__ __forceinline__ void tpp_zone(double (&acc)[MP * (MP + 1) / 2], const bool on, const Grid& grid, const NodeCommon<MP>& C) {
    constexpr int NPL = tpp_npl(P);
    const double* __restrict__ myCt = C.myCt;
    const int n_blocks = (NEAR ? grid.near_count() : grid.far_count()) / NPL;
    for (int bt = 0; bt < n_blocks; ++bt) {
        const int h0 = NEAR ? grid.near_handle(bt * NPL) : grid.far_handle(bt * NPL);  // handles of a block are consecutive
        double z[NPL], h[NPL], tmx[NPL], aux[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            grid.load(h0 + i, tmx[i], aux[i]);
            z[i] = tmx[i] * C.zs;  // (x_th - x_j)/θ
        }
        if constexpr (NEAR) {
            const int Kj = grid.block_degree(h0, NPL);  // node-only degree: the same for every parcel
            double r[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) r[i] = fma(tmx[i], C.rq, -1.0);
            if (!C.warp_capped) {
                const double t_top = myCt[Kj * TPP_THREADS];
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = t_top;
#pragma unroll 4
                for (int m = Kj - 1; m >= 0; --m) {
                    const double tm = myCt[m * TPP_THREADS];
#pragma unroll
                    for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], r[i], tm);
                }
            } else {
                // a parcel whose centre is capped at the series limit sees |r| up to 0.1 at EVERY near node (r no longer
                // vanishes with x_j/x_th): it takes the full degree; its warp-mates keep their own degree by predication
                const int Kown = C.capped ? (TPP_TAYLOR_MAX - 1) : Kj;
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = 0.0;
                for (int m = TPP_TAYLOR_MAX - 1; m >= 0; --m) {
                    const double tm = myCt[m * TPP_THREADS];
                    if (m <= Kown) {
#pragma unroll
                        for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], r[i], tm);
                    }
                }
            }
        } else {
            // Horner from the warp's largest degree; the own table is zero above the parcel's own degree, so the
            // result does not depend on the neighbours
            const double c_top = myCt[C.deg_w * TPP_THREADS];
#pragma unroll
            for (int i = 0; i < NPL; ++i) h[i] = c_top;
            int n = C.deg_w - 1;
            for (; n >= 3; n -= 4) {
                const double c0 = myCt[n * TPP_THREADS], c1 = myCt[(n - 1) * TPP_THREADS], c2 = myCt[(n - 2) * TPP_THREADS],
                             c3 = myCt[(n - 3) * TPP_THREADS];
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], z[i], c0);
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], z[i], c1);
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], z[i], c2);
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], z[i], c3);
            }
            for (; n >= 0; --n) {
                const double c0 = myCt[n * TPP_THREADS];
#pragma unroll
                for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], z[i], c0);
            }
        }
        const bool interior = grid.block_interior(h0, NPL);
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            double gE = fast_exp(fma(C.k, grid.log_sum(h0 + i, aux[i]), C.e0), C.exp_tab);  // g_j * E_j
            double hs = (z[i] < C.cf_lim) ? h[i] : 0.0;  // continued-fraction nodes: added by tpp_cf_nodes
            if constexpr (MASKED) {
                gE = on ? gE : 0.0;
                hs = on ? hs : 0.0;
            }
            // v_p = g E h_p with h_top = z^{MP-1} S, h_p = (h_{p+1} + z^p)/(k+p): carry g E z^p instead of z^p
            double y[MP];
            y[0] = gE;
#pragma unroll
            for (int p = 1; p < MP; ++p) y[p] = y[p - 1] * z[i];
            double v[MP];
            v[MP - 1] = y[MP - 1] * hs;
#pragma unroll
            for (int p = MP - 2; p >= 0; --p) v[p] = (v[p + 1] + y[p]) * C.ia[p];  // downward recurrence
            double w[MP];
            grid.template weights<MP>(h0 + i, aux[i], interior, w);
            int t = 0;
#pragma unroll
            for (int p1 = 0; p1 < MP; ++p1) {
#pragma unroll
                for (int p2 = p1; p2 < MP; ++p2) {
                    // the S terms read F[a+c][b+m-c] with a,b < P, m <= 2: only entries with p1 + p2 <= 2P are ever used
                    if (p1 + p2 <= 2 * P) acc[t] = fma(w[p1], v[p2], acc[t]);
                    ++t;
                }
            }
        }
    }
}

// rare path: nodes in the continued-fraction regime (only when x_th/θ reaches the series limit)
template <int MP, int P, bool MASKED, typename Grid>
__device__ __forceinline__ void tpp_cf_nodes(double (&acc)[MP * (MP + 1) / 2], const bool on, const Grid& grid, const NodeCommon<MP>& C) {
    if (!C.warp_cf) return;
    double B[MP];
    B[MP - 1] = 1.0;
#pragma unroll
    for (int p = MP - 2; p >= 0; --p) B[p] = B[p + 1] * C.ia[p];
    const int n_all = grid.near_count() + grid.far_count();
    double G[MP];  // G[p1] = sum_j w_j dx x_j^p1 xi_j; the p2 dependence is the node-independent factor B[p2]
#pragma unroll
    for (int p = 0; p < MP; ++p) G[p] = 0.0;
#pragma unroll 1
    for (int jj = 0; jj < n_all; ++jj) {
        const int hd = (jj < grid.near_count()) ? grid.near_handle(jj) : grid.far_handle(jj - grid.near_count());
        double tmx, aux;
        grid.load(hd, tmx, aux);
        const double z = tmx * C.zs;
        const bool cf_j = !(z < C.ser_lim);
        if (!__any_sync(0xffffffffu, cf_j)) continue;
        const double zc = fmin(z, 256.0);  // beyond this the upper function is < 1e-80 of Gamma(a)
        double b = zc + 1.0 - C.a_top;
        double Pm = 1.0, Pc = b, Qm = 0.0, Qc = 1.0, fn = 0.0;
        for (int n = 1; n <= C.cfd_w; ++n) {
            // beyond the parcel's own depth the step degenerates to P <- 1*P + 0, which is exact
            fn += 1.0;
            b += 2.0;
            const bool go = n <= C.cfd;
            const double an = go ? fn * (C.a_top - fn) : 0.0;  // -n(n-a)
            const double bb = go ? b : 1.0;
            const double Pn = fma(bb, Pc, an * Pm);
            const double Qn = fma(bb, Qc, an * Qm);
            Pm = Pc; Pc = Pn; Qm = Qc; Qc = Qn;
        }
        const double gE = fast_exp(fma(C.k, grid.log_sum(hd, aux), C.e0), C.exp_tab);
        const double g = fast_exp(fmax(fma(C.k, grid.ell(hd) - C.log_u, -(grid.x(hd) * C.zs)), kExpOffsetMin), C.exp_tab);  // (x_j/θ)^k e^{-x_j/θ}
        double zt = 1.0;
#pragma unroll
        for (int p = 1; p < MP; ++p) zt *= z;
        double xi = fma(g, C.gam_top, -(gE * zt) * (Qc / Pc));
        xi = cf_j ? xi : 0.0;
        if constexpr (MASKED) xi = on ? xi : 0.0;
        double w[MP];
        grid.template weights<MP>(hd, aux, false, w);
#pragma unroll
        for (int p1 = 0; p1 < MP; ++p1) G[p1] = fma(w[p1], xi, G[p1]);
    }
    int t = 0;
#pragma unroll
    for (int p1 = 0; p1 < MP; ++p1) {
#pragma unroll
        for (int p2 = p1; p2 < MP; ++p2) {
            if (p1 + p2 <= 2 * P) acc[t] = fma(G[p1], B[p2], acc[t]);
            ++t;
        }
    }
}

// exp + accumulation of one block of NPL nodes (FixedThreshold node loop): top[p1] += w'_p1 (g E h)_j with
// w'_p1 = w_j dx x_j^p1 (x_th - x_j)^(P+1) (the caller applies θ^-(P+1)); the lower-order sums Z[p1][p] do not involve h and come
// from the k-tables of cloudy_config_set.  Inlined into the same basic block as the Horner evaluation of h, so that the
// scheduler interleaves the NPL exponential chains with the NPL Horner chains (the two are independent).
template <int MP, int P, int NPL>
__device__ __forceinline__ void tpp_block_tail(double (&top)[P + 1], double (&Z)[(P + 1) * (P + 2) / 2], const double2* __restrict__ rb,
                                               const double (&z)[NPL], const double (&h)[NPL], const double (&ls)[NPL],
                                               const double k, const double e0, const double cf_lim, const double* __restrict__ exp_tab) {
    constexpr int S2 = tpp_rec2_stride(P);
    constexpr int P1 = P + 1;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const double gE = fast_exp(fma(k, ls[i], e0), exp_tab);  // g_j * E_j
        const double hs = (z[i] < cf_lim) ? h[i] : 0.0;          // continued-fraction nodes are added by the rare loop (a warp-uniform branch here measured 1 % slower)
        double w[P1 + 1];
#pragma unroll
        for (int q = 0; q < (S2 - 2) / 2; ++q) {
            const double2 ww = rb[i * (S2 / 2) + 1 + q];
            if (2 * q < P1 + 1) w[2 * q] = ww.x;
            if (2 * q + 1 < P1 + 1) w[2 * q + 1] = ww.y;
        }
        if constexpr (tpp_ztab(P)) {
            const double gh = gE * hs;
#pragma unroll
            for (int p1 = 0; p1 < P1; ++p1) top[p1] = fma(w[p1], gh, top[p1]);
        } else {
            // y_p = g E z^p, top[p1] += w_p1 y_P z h, Z[p1][p] += w_p1 y_p
            const double zh = z[i] * hs;
            double y[P1];
            y[0] = gE;
#pragma unroll
            for (int p = 1; p < P1; ++p) y[p] = y[p - 1] * z[i];
            const double ytop = y[P] * zh;
#pragma unroll
            for (int p1 = 0; p1 < P1; ++p1) {
                top[p1] = fma(w[p1], ytop, top[p1]);
#pragma unroll
                for (int p = p1; p < P1; ++p) Z[tri_ct(p1, p, P1)] = fma(w[p1], y[p], Z[tri_ct(p1, p, P1)]);
            }
        }
    }
}

// Near-zone Taylor polynomial sum_{m<=K} t_m r^m with a compile-time degree (no loop counter, immediate offsets).
// (Splitting it into its even and odd parts in r^2, two chains per node, was measured: +2.5 % with 3 blocks/SM, -1.5 % with 4.)
template <int K, int NPL, bool VOL>
__device__ __forceinline__ void tpp_taylor_horner(double (&h)[NPL], const double (&r)[NPL], const double* __restrict__ myCt) {
    // VOLATILE loads of the Taylor coefficients: with plain loads the compiler hoists all 27 of them above the degree switch
    // (54 registers), and under that pressure ptxas serialises the NPL Horner chains node by node — every DFMA then waits out
    // the 8-cycle latency alone.  Loaded where they are used, the chains interleave and the spills disappear (-4.5 % on C5).
    // (The 255-register instances of the high-order tensors have the registers and measured 3 % faster with plain loads.)
    typename std::conditional<VOL, const volatile double*, const double*>::type vct = myCt;
    const double t_top = vct[K * TPP_THREADS];
#pragma unroll
    for (int i = 0; i < NPL; ++i) h[i] = t_top;
#pragma unroll
    for (int m = K - 1; m >= 0; --m) {
        const double tm = vct[m * TPP_THREADS];
#pragma unroll
        for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], r[i], tm);
    }
}

struct FixedGrid {
    const double* rec;    // shared memory: aligned records, [near | far]
    const double* kblk;   // shared memory: Taylor degree per near block (rounded up to its class)
    const int* cls_end;   // constant bank: near blocks [0, cls_end[c]) belong to the degree classes <= c
    int n_near, n_far;    // padded node counts
    // natural-order SoA tables in global memory (continued-fraction nodes only): XJ[nb] ELL[nb] TMX[nb] LZ[nb] W[M][nb]
    const double* soa;
    int nb;
};

// FixedThreshold node loop.  Same quadrature, nodes and special-function evaluation as tpp_zone / tpp_cf_nodes above; what
// differs is the bookkeeping of the T = MP(MP+1)/2 sums.  With v_p = g E h_p and h_p = (h_{p+1} + z^p)/(k+p) the sums obey
//   acc[p1][p] = (acc[p1][p+1] + Z[p1][p]) / (k+p),   Z[p1][p] = sum_j w_j x_j^p1 (g E z^p)_j,   acc[p1][top] = sum_j w_j x_j^p1 (g E z^top S)_j
// (all terms positive), so a node adds into Top[p1] and Z[p1][p] only (P+1 + (P+1)(P+2)/2 fused multiply-adds) and the
// downward recurrence runs ONCE per parcel after the loop instead of once per node.
template <int MP, int P>
__device__ __forceinline__ void tpp_nodes_fixed2(double (&acc)[MP * (MP + 1) / 2], const FixedGrid grid, const double k,
                                                 const double inv_th, const double log_th, const double X, const double poch_top, double& inv_gk,
                                                 const double (&ia)[MP], double* __restrict__ myCt, const int deg,
                                                 const unsigned char* __restrict__ cfdz, const double a_top, const double ser_lim,
                                                 const double* __restrict__ exp_tab, const double* __restrict__ ztab, const double zt_inv_h,
                                                 const int zt_n, const double zt_L, const bool gamma_mode) {
    static_assert(MP == P + 2, "all M = P + 2 orders are carried");
    constexpr int T = MP * (MP + 1) / 2;
    constexpr int NPL = tpp_npl(P);
    constexpr int S2 = tpp_rec2_stride(P);
    constexpr int P1 = P + 1;              // p1 = 0..P (p1 <= p2 and p1 + p2 <= 2P), top order = P + 1
    constexpr int NZ = P1 * (P1 + 1) / 2;  // Z[p1][p], p1 <= p <= P
    double top[P1], Z[NZ];
#pragma unroll
    for (int i = 0; i < P1; ++i) top[i] = 0.0;
    const double e0 = fmax(fma(-2.0 * k, log_th, -X), kExpOffsetMin);  // exponent offset of g*E
    // Z[p1][p] = sum_j w_j x_j^p1 (g E z^p)_j = exp(e0 + k L) θ^-p G_{p1,p}(k): the node-only functions G come from the degree-7
    // polynomials of cloudy_config_set (one 64-byte record per (interval, t), NZ records = a whole number of 128-byte lines per
    // interval).  The records are requested now (L1 prefetch: no registers) and evaluated after the node loops, where the sums
    // are needed: nothing of Z is live while the loops run.
    const double zt_u = k * zt_inv_h;
    const int zt_iv = min(max((int)zt_u, 0), zt_n - 1);
    const double2* __restrict__ zr = reinterpret_cast<const double2*>(ztab) + (size_t)zt_iv * ((NZ + 1) * 4);  // NZ sums + 1/Γ(k+1)
    if constexpr (tpp_ztab(P)) {
#pragma unroll
        for (int l = 0; l < ((NZ + 1) * 64 + 127) / 128; ++l) asm volatile("prefetch.global.L1 [%0];" ::"l"(zr + l * 8));
    } else {
#pragma unroll
        for (int i = 0; i < NZ; ++i) Z[i] = 0.0;
    }
    const double Xc = fmin(X, ser_lim - 0.5);       // Taylor centre (inside the series regime)
    const double rq = inv_th / Xc;                  // r = z/X_c - 1 = (x_th - x_j) rq - 1
    const bool capped = !(X <= ser_lim - 0.5);      // centre below x_th/θ: r does not vanish at the first nodes
    const bool warp_capped = __any_sync(0xffffffffu, capped);
    const bool warp_cf = __any_sync(0xffffffffu, !(X < ser_lim));  // any parcel of the warp with continued-fraction nodes
    const double cf_lim = warp_cf ? ser_lim : INFINITY;

    // ---- far nodes: S(z) = sum_n c_n z^n, c_n = 1/(a)_{n+1}, by Horner on SCALED coefficients generated on the fly:
    //   Q_n = prod_{i=n+1..deg} (a+i) = c_n (a)_{deg+1},   R <- R z + Q_n,   S(z) = R_0 / (a)_{deg+1}
    // (all terms positive).  One product Q_n serves the TPP_NPLF Horner chains in flight; the last slot of the last block is a
    // zero-weight dummy whose argument is replaced by the Taylor centre X_c, which yields S(X_c) for the near zone.
    // Each parcel runs its own degree `deg` (table kSeriesDeg2, truncation <= 2e-16): nothing depends on the warp-mates.
    const int n_far_b = grid.n_far / TPP_NPLF;
    double t0 = 0.0;       // S(X_c)
    double inv_poch = 0.0; // 1/(a)_{deg+1}
    for (int bt = 0; bt < n_far_b; ++bt) {
        const double2* __restrict__ rb = reinterpret_cast<const double2*>(grid.rec + (grid.n_near + bt * TPP_NPLF) * S2);
        double z[TPP_NPLF], R[TPP_NPLF], ls[TPP_NPLF];
#pragma unroll
        for (int i = 0; i < TPP_NPLF; ++i) {
            const double2 tl = rb[i * (S2 / 2)];
            z[i] = tl.x * inv_th;
            ls[i] = tl.y;
            R[i] = 1.0;
        }
        const bool last = (bt == n_far_b - 1);
        if (last) z[TPP_NPLF - 1] = Xc;
        double Q = 1.0, aa = a_top + (double)deg;
#pragma unroll 2
        for (int n = deg - 1; n >= 0; --n) {
            Q *= aa;      // Q_n = Q_{n+1} (a + n + 1)
            aa -= 1.0;
#pragma unroll
            for (int i = 0; i < TPP_NPLF; ++i) R[i] = fma(R[i], z[i], Q);
        }
        if (bt == 0) inv_poch = 1.0 / (Q * a_top);  // aa == a_top here
#pragma unroll
        for (int i = 0; i < TPP_NPLF; ++i) R[i] *= inv_poch;
        if (last) t0 = R[TPP_NPLF - 1];
        tpp_block_tail<MP, P, TPP_NPLF>(top, Z, rb, z, R, ls, k, e0, cf_lim, exp_tab);
    }
    {
        // Taylor coefficients of S about X_c into the parcel's shared-memory column:
        //   t_1 = (X_c - a) t_0 + 1,  t_{m+1} = [(X_c - a - m) t_m + X_c t_{m-1}]/(m+1); the factors are formed off the dependent chain
        const double Xa = Xc - a_top;
        double tm1 = t0, tm = fma(Xa, t0, 1.0);
        myCt[0] = tm1;
        myCt[TPP_THREADS] = tm;
#pragma unroll
        for (int m = 1; m < TPP_TAYLOR_MAX; ++m) {
            const double inv = 1.0 / (double)(m + 1);
            const double tn = fma((Xa - (double)m) * inv, tm, (Xc * inv) * tm1);
            myCt[(m + 1) * TPP_THREADS] = tn;
            tm1 = tm;
            tm = tn;
        }
    }
    // ---- near nodes: Taylor polynomial about X_c, degree by node block ----
    // Two loops, chosen per warp, so that the common one stays compact in the instruction cache: (1) no parcel of the warp has
    // its centre capped: compile-time degrees; (2) some parcel is capped at the series limit: it sees |r| up to 0.1 at EVERY
    // near node (r no longer vanishes with x_j/x_th) and takes the full degree, its warp-mates keep their own degree by
    // predication (same operations as in loop 1, so a parcel's result does not depend on the loop its warp runs).
    const int n_near_b = grid.n_near / NPL;
    constexpr int CAP_UNROLL = (P >= 4) ? TPP_TAYLOR_MAX + 1 : 1;
    if (!warp_capped) {
        // one loop per degree class: loop bounds are run constants (uniform registers), the degree is a compile-time constant of
        // the loop body — the per-block jump table of the earlier version (LDS + F2I + LDC + BRX on the critical path of
        // every block) cost 6-8 % of the kernel
        int bt = 0;
        auto run_class = [&](auto ktag, const int end) {
            constexpr int K = decltype(ktag)::value;
            for (; bt < end; ++bt) {
                const double2* __restrict__ rb = reinterpret_cast<const double2*>(grid.rec + bt * NPL * S2);
                double z[NPL], h[NPL], ls[NPL], r[NPL];
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const double2 tl = rb[i * (S2 / 2)];
                    r[i] = fma(tl.x, rq, -1.0);
                    z[i] = tl.x * inv_th;
                    ls[i] = tl.y;
                }
                tpp_taylor_horner<K, NPL, (P < 4)>(h, r, myCt);
                tpp_block_tail<MP, P, NPL>(top, Z, rb, z, h, ls, k, e0, cf_lim, exp_tab);
            }
        };
        run_class(std::integral_constant<int, tpp_taylor_class(0)>{}, grid.cls_end[0]);
        run_class(std::integral_constant<int, tpp_taylor_class(1)>{}, grid.cls_end[1]);
        run_class(std::integral_constant<int, tpp_taylor_class(2)>{}, grid.cls_end[2]);
        run_class(std::integral_constant<int, tpp_taylor_class(3)>{}, grid.cls_end[3]);
        run_class(std::integral_constant<int, tpp_taylor_class(4)>{}, n_near_b);
    } else {
        for (int bt = 0; bt < n_near_b; ++bt) {
            const double2* __restrict__ rb = reinterpret_cast<const double2*>(grid.rec + bt * NPL * S2);
            const int Kj = (int)grid.kblk[bt];
            double z[NPL], h[NPL], ls[NPL], r[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                const double2 tl = rb[i * (S2 / 2)];
                r[i] = fma(tl.x, rq, -1.0);
                z[i] = tl.x * inv_th;
                ls[i] = tl.y;
                h[i] = 0.0;
            }
            const int Kown = capped ? TPP_TAYLOR_MAX : Kj;
            // a block whose nodes lie beyond the series limit in every parcel of the warp (threshold deep in the tail: the
            // near nodes sit at z ~ x_th/θ) is evaluated by the continued-fraction loop alone; its Taylor values would be
            // discarded by the select in tpp_block_tail (h stays 0: bit-identical)
            double zmin = z[0];
#pragma unroll
            for (int i = 1; i < NPL; ++i) zmin = fmin(zmin, z[i]);
            const bool all_cf = __all_sync(0xffffffffu, !(zmin < cf_lim));
            // unrolled for the high-order tensors only: rolled, this loop costs C4 (a quarter of its parcels is capped) 40 %;
            // unrolled, its code costs C5 (no capped parcel) 2 % through the instruction cache
            if (!all_cf) {
#pragma unroll CAP_UNROLL
                for (int m = TPP_TAYLOR_MAX; m >= 0; --m) {
                    const double tm = myCt[m * TPP_THREADS];
                    if (m <= Kown) {
#pragma unroll
                        for (int i = 0; i < NPL; ++i) h[i] = fma(h[i], r[i], tm);
                    }
                }
            }
            tpp_block_tail<MP, P, NPL>(top, Z, rb, z, h, ls, k, e0, cf_lim, exp_tab);
        }
    }
    if constexpr (tpp_ztab(P)) {
        const double tt = fma(2.0, zt_u - (double)zt_iv, -1.0);
        const double zs0 = fast_exp_t<true>(fma(k, zt_L, e0), exp_tab);
        double zsc = zs0;
        double scp[P1];  // the scale θ^-p depends on p only
#pragma unroll
        for (int p = 0; p < P1; ++p) { scp[p] = zsc; zsc *= inv_th; }
#pragma unroll
        for (int p1 = 0; p1 < P1; ++p1)
#pragma unroll
            for (int p = p1; p < P1; ++p) {
                const int t = tri_ct(p1, p, P1);
                const double2 c01 = __ldg(zr + t * 4), c23 = __ldg(zr + t * 4 + 1), c45 = __ldg(zr + t * 4 + 2), c67 = __ldg(zr + t * 4 + 3);
                double g = fma(c67.y, tt, c67.x);
                g = fma(g, tt, c45.y);
                g = fma(g, tt, c45.x);
                g = fma(g, tt, c23.y);
                g = fma(g, tt, c23.x);
                g = fma(g, tt, c01.y);
                g = fma(g, tt, c01.x);
                Z[t] = g * scp[p];
            }
        if (gamma_mode) {
            // 1/Γ(k) = k / Γ(k+1) from the same record set (entry NZ tabulates the entire function 1/Γ(k+1)): replaces the
            // log + exp + Stirling tail of gamma_shape_inv (~150 instructions per parcel); needed only from here on
            const double2 c01 = __ldg(zr + NZ * 4), c23 = __ldg(zr + NZ * 4 + 1), c45 = __ldg(zr + NZ * 4 + 2), c67 = __ldg(zr + NZ * 4 + 3);
            double g = fma(c67.y, tt, c67.x);
            g = fma(g, tt, c45.y);
            g = fma(g, tt, c45.x);
            g = fma(g, tt, c23.y);
            g = fma(g, tt, c23.x);
            g = fma(g, tt, c01.y);
            g = fma(g, tt, c01.x);
            inv_gk = k * g;
        }
    }
    // downward recurrence of the sums; only entries with p1 + p2 <= 2P are ever read by the S terms
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
    double th_top = 1.0;  // table variant: θ^-(P+1), the top weights carry (x_th - x_j)^(P+1)
    if constexpr (tpp_ztab(P)) {
#pragma unroll
        for (int p = 0; p < P + 1; ++p) th_top *= inv_th;
    }
#pragma unroll
    for (int p1 = 0; p1 < P1; ++p1) {
        double a = top[p1] * th_top;
        if (p1 + (MP - 1) <= 2 * P) acc[tri_ct(p1, MP - 1, MP)] = a;
#pragma unroll
        for (int p = P; p >= p1; --p) {
            a = (a + Z[tri_ct(p1, p, P1)]) * ia[p];
            if (p1 + p <= 2 * P) acc[tri_ct(p1, p, MP)] = a;
        }
    }

    // ---- rare path: nodes in the continued-fraction regime (only when x_th/θ reaches the series limit) ----
    if (warp_cf) {
        const double gam_top = poch_top / inv_gk;  // Γ(k+MP-1) = Γ(k) (k)_{MP-1}
        double B[MP];
        B[MP - 1] = 1.0;
#pragma unroll
        for (int p = MP - 2; p >= 0; --p) B[p] = B[p + 1] * ia[p];
        double G[P1];  // G[p1] = sum_j w_j dx x_j^p1 xi_j; the p2 dependence is the node-independent factor B[p2]
#pragma unroll
        for (int p = 0; p < P1; ++p) G[p] = 0.0;
        const int nb = grid.nb;
        const double* __restrict__ gx = grid.soa;
        const double* __restrict__ gell = grid.soa + nb;
        const double* __restrict__ gtmx = grid.soa + 2 * nb;
        const double* __restrict__ glz = grid.soa + 3 * nb;
        const double* __restrict__ gw = grid.soa + 4 * nb;
#pragma unroll 1
        for (int j = 0; j < nb; ++j) {
            const double z = __ldg(gtmx + j) * inv_th;
            const bool cf_j = !(z < ser_lim);
            if (!__any_sync(0xffffffffu, cf_j)) continue;
            const double zc = fmin(z, 256.0);  // beyond this the upper function is < 1e-80 of Gamma(a)
            double b = zc + 1.0 - a_top;
            double Pm = 1.0, Pc = b, Qm = 0.0, Qc = 1.0, fn = 0.0;
            // depth by (a, z): the fraction needs fewer levels the further z lies beyond the series limit (kCfDepthZ)
            const int dj = cf_j ? (int)cfdz[(int)fmin(fmax((zc - ser_lim) * 0.25, 0.0), (double)(kCfZBins - 1))] : 0;
            const int dj_w = __reduce_max_sync(0xffffffffu, dj);
            for (int n = 1; n <= dj_w; ++n) {
                // beyond the parcel's own depth the step degenerates to P <- 1*P + 0, which is exact
                fn += 1.0;
                b += 2.0;
                const bool on = n <= dj;
                const double an = on ? fn * (a_top - fn) : 0.0;  // -n(n-a)
                const double bb = on ? b : 1.0;
                const double Pn = fma(bb, Pc, an * Pm);
                const double Qn = fma(bb, Qc, an * Qm);
                Pm = Pc; Pc = Pn; Qm = Qc; Qc = Qn;
            }
            const double ell = __ldg(gell + j);
            const double gE = fast_exp(fma(k, ell + __ldg(glz + j), e0), exp_tab);
            const double g = fast_exp(fmax(fma(k, ell - log_th, -(__ldg(gx + j) * inv_th)), kExpOffsetMin), exp_tab);  // (x_j/θ)^k e^{-x_j/θ}
            double zt = 1.0;
#pragma unroll
            for (int p = 1; p < MP; ++p) zt *= z;
            double xi = fma(g, gam_top, -(gE * zt) * (Qc / Pc));
            xi = cf_j ? xi : 0.0;
#pragma unroll
            for (int p1 = 0; p1 < P1; ++p1) G[p1] = fma(__ldg(gw + p1 * nb + j), xi, G[p1]);
        }
#pragma unroll
        for (int p1 = 0; p1 < P1; ++p1) {
#pragma unroll
            for (int p2 = p1; p2 < MP; ++p2) {
                if (p1 + p2 <= 2 * P) acc[tri_ct(p1, p2, MP)] = fma(G[p1], B[p2], acc[tri_ct(p1, p2, MP)]);
            }
        }
    }
}

// Truncated 2-D integral of a Lognormal mode (ParticleDistributions.jl:614-625):
//   H(p1,p2) = int_0^T y^{p2} f(y) [ int_0^{T-y} x^{p1} f(x) dx ] dy,  inner integral in closed form
//   n exp(p1 μ + p1² σ²/2) Φ((ln(T-y) - μ - p1 σ²)/σ), outer integral by a fixed Gauss-Legendre rule in t = ln y.
// (The reference nests two adaptive QuadGK calls at rtol sqrt(eps); this rule agrees with it to ~1e-10.)
template <int MP>
__device__ __noinline__ void tpp_lognormal_H(double (&acc)[MP * (MP + 1) / 2], const double* __restrict__ gl, const int gl_n, const double n,
                                                const double mu, const double sg, const double Tthr) {
    constexpr int T = MP * (MP + 1) / 2;
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
    const double s2 = sg * sg;
    const double hi = fmin(log(Tthr), mu + (double)(MP - 1) * s2 + 12.0 * sg);
    const double lo = mu - 12.0 * sg;
    const bool ok = hi > lo;
    const double half = ok ? 0.5 * (hi - lo) : 0.0, mid = 0.5 * (hi + lo);
    const double inv_sg = 1.0 / sg;
    const double pref = n * inv_sg * 0.3989422804014327;  // n / (σ sqrt(2π))
    double Cp[MP];
#pragma unroll
    for (int p = 0; p < MP; ++p) Cp[p] = n * exp((double)p * mu + (double)(p * p) * s2 / 2);
    for (int q = 0; q < gl_n; ++q) {
        const double t = fma(half, gl[q], mid);
        const double y = exp(t);
        const double d = (t - mu) * inv_sg;
        const double wq = gl[gl_n + q] * half * pref * exp(-0.5 * d * d);  // weight * f(y) y
        const double rem = Tthr - y;
        const double lr = log(fmax(rem, 1e-300));
        double in[MP];
#pragma unroll
        for (int p = 0; p < MP; ++p) in[p] = (rem > 0.0) ? Cp[p] * norm_cdf((lr - mu - (double)p * s2) * inv_sg) : 0.0;
        double yp = wq;
        int tt[MP];
#pragma unroll
        for (int p2 = 0; p2 < MP; ++p2) {
#pragma unroll
            for (int p1 = 0; p1 <= p2; ++p1) acc[tri_ct(p1, p2, MP)] = fma(yp, in[p1], acc[tri_ct(p1, p2, MP)]);
            yp *= y;
        }
        (void)tt;
    }
}

// S_1k and S_2k of one mode, all orders m < 3 — Coalescence.jl:353-455.  The reference sums 0.5 c_ab binomial(m,c) F(a+c, b+m-c)
// over (a, b, c); F is symmetric, so the sum is folded on the host into one weight per (m, u <= v) (DevConfig::sw): one
// multiply-add per weight here instead of P*P*(m+1) terms with their index arithmetic — for P = 5 the three unrolled copies per
// mode were 15 k of the kernel's 26 k instructions.  Fs[t(u,v)] holds the truncated integrals (already limited and masked),
// S2 uses Mom_u Mom_v - F in the same weights.
template <int P>
__device__ __forceinline__ void tpp_s_terms(const DevConfig& cfg, const int k, const double (&momk)[P + 2],
                                            const double (&Fs)[(P + 2) * (P + 3) / 2], double (&s1)[3], double (&s2)[3]) {
    constexpr int M = P + 2;
    double a1[3] = {0.0, 0.0, 0.0}, a2[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < M; ++u)
#pragma unroll
        for (int v = u; v < M; ++v) {
            if (u + v > 2 * P) continue;  // a + b + m <= 2P: never referenced
            const int t = tri_ct(u, v, M);
            const double f = Fs[t];
            const double d = momk[u] * momk[v] - f;
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                if (u + v < m || u + v > 2 * (P - 1) + m) continue;  // a + b = u + v - m must lie in [0, 2(P-1)]
                a1[m] = fma(cfg.sw[k][m][t], f, a1[m]);
                a2[m] = fma(cfg.sw[k][m][t], d, a2[m]);
            }
        }
#pragma unroll
    for (int m = 0; m < 3; ++m) { s1[m] = a1[m]; s2[m] = a2[m]; }
}

struct TppShared {
    unsigned char deg[kSerZ][kSerA];
    double serlim[kSerA];
    double exp32[TPP_EXP_TAB];  // 2^(i/256)
    int cfd[kSerA];
    unsigned char cfdz[kSerA][kCfZBins];  // continued-fraction depth by (floor(a), z bin), margin included
    unsigned char kdeg_m[128];  // MovingThreshold own grids: Taylor degree bound of node m (counted from the threshold)
};

template <int N, int P, int MODEL>
__global__ void __launch_bounds__(TPP_THREADS, tpp_min_blocks(N, P, MODEL)) tpp_kernel(const __grid_constant__ DevConfig cfg, const KArgs args) {
    constexpr int M = P + 2;
    constexpr bool RAIN = (MODEL == MODEL_RAINSHAFT);
    constexpr bool MOVING = (MODEL == MODEL_BOX_MOVING);  // MovingThreshold has its own instances (box model only, like the reference)
    extern __shared__ __align__(16) double smem[];
    __shared__ TppShared sh;
    double* sTab = smem;
    // staged tables: MovingThreshold instances use the packed records of TableGrid, the others the aligned records of FixedGrid
    const int st_total = MOVING ? cfg.tpp_total : cfg.tpp2_total;
    const int st_off = MOVING ? cfg.tpp_off : cfg.tpp2_off;
    double* sCt = smem + ((st_total + 1) & ~1);
    const int tid = threadIdx.x;
    for (int i = tid; i < st_total; i += TPP_THREADS) sTab[i] = cfg.tab[st_off + i];
    for (int i = tid; i < kSerZ * kSerA; i += TPP_THREADS) sh.deg[i / kSerA][i % kSerA] = kSeriesDeg2[i / kSerA][i % kSerA];
    if (tid < kSerA) { sh.cfd[tid] = kCfDepth[tid]; sh.serlim[tid] = kSeriesLimit[tid]; }
    for (int i = tid; i < kSerA * kCfZBins; i += TPP_THREADS) {
        const int dz_ = kCfDepthZ[i / kCfZBins][i % kCfZBins];
        sh.cfdz[i / kCfZBins][i % kCfZBins] = (unsigned char)(dz_ > 0 ? dz_ + 1 : 0);
    }
    for (int i = tid; i < TPP_EXP_TAB; i += TPP_THREADS) sh.exp32[i] = exp2((double)i / (double)TPP_EXP_TAB);
    if (MOVING && tid < 128) {
        // same degree rule as the host's table grid (cloudy_config_set): |r| <= 1.15 ρ, ρ_m <= 10^(-m/bins_per_log_unit)
        const double rr = 1.15 * exp10(-(double)tid / (double)cfg.bins_per_log_unit);
        const double rho_thr[10] = {1e-5, 1e-4, 1e-3, 3e-3, 1e-2, 2e-2, 3e-2, 5e-2, 7e-2, 0.1};
        const int k_of[10] = {4, 5, 7, 9, 11, 14, 16, 19, 22, 25};
        int kk = TPP_TAYLOR_MAX;
        for (int q = 9; q >= 0; --q) if (rr <= rho_thr[q]) kk = k_of[q];
        sh.kdeg_m[tid] = (unsigned char)kk;
    }
    __syncthreads();
    double* myCt = sCt + tid;

    const long long n = args.n;
    // Element offsets are 32-bit unsigned (the host refuses buffers of 2^32 doubles or more, 34 GB each): one IMAD + one
    // IMAD.WIDE per address instead of two 64-bit multiplies (~12 instructions) for each of the ~25 addresses of a parcel
    const unsigned s_in = (unsigned)args.s_in, ps_in = (unsigned)args.ps_in, s_out = (unsigned)args.s_out, ps_out = (unsigned)args.ps_out;
    const unsigned s_n = (unsigned)args.s_n, s_flux = (unsigned)args.s_flux, s_clip = (unsigned)args.s_clip;
    // software prefetch: the next parcel's moments are requested while the current one is being evaluated
    auto parcel_of = [&](long long b) -> unsigned {
        const long long ix = b + (tid & 31);
        const unsigned q = (unsigned)((ix < n) ? ix : n - 1);
        return (args.perm != nullptr) ? (unsigned)args.perm[q] : q;
    };
    // Dynamic tile schedule: tiles are consecutive positions of the (regime-sorted) order, drawn from a global counter.  With
    // the static round-robin of the earlier versions the kernel lasted as long as its unluckiest warp (parcel cost varies with
    // the regime; sm__warps_active was 13.5 of 16 on C5).  The counter is never reset: every drawer draws exactly one value
    // beyond the last tile, so a launch advances it by n_tiles + n_drawers and the host passes the value it has at launch
    // (args.tile_base).  Draws run two tiles ahead, so their latency is never waited for.
    //   Box instances (BSYNC): a tile is TPP_THREADS positions, drawn by thread 0 and published through shared memory at the
    // barrier that opens every iteration, so the four warps of a block walk through the kernel's code together and share
    // instruction-cache lines (with independent warps instruction fetch was 25 % of the stall samples on C5: 2.72 -> 2.58 ms).
    //   Column instances: a tile is one warp's 32 positions and every warp draws for itself — a block there often holds warps
    // of empty cells next to warps of cloudy ones, which a barrier would chain together (measured: 0.74 -> 0.78 ms per C3 step).
    constexpr bool BSYNC = tpp_block_sync(MODEL);
    constexpr int TILE = tpp_tile(MODEL);
    __shared__ long long s_tile[2];
    const long long n_tiles = (n + TILE - 1) / TILE;
    auto draw1 = [&]() -> long long { return (long long)(atomicAdd(args.tile_ctr, 1ULL) - args.tile_base); };
    auto draw = [&]() -> long long {  // warp-level draw
        long long v = 0;
        if ((tid & 31) == 0) v = draw1();
        return __shfl_sync(0xffffffffu, v, 0);
    };
    // tiles are handed out from the END of the order: the regime sort puts the expensive parcels (continued-fraction regime,
    // long series) last and the empty cells of a column model first, so the kernel's tail is made of the cheapest tiles
    auto tile_pos = [&](long long t) -> long long { return (n_tiles - 1 - t) * TILE + (BSYNC ? (tid & ~31) : 0); };
    long long tile, tile_nx, tile_nx2;  // BSYNC: tile_nx2 is thread 0's draw in flight
    int it = 0;
    if constexpr (BSYNC) {
        if (tid == 0) s_tile[0] = draw1();
        __syncthreads();
        tile = s_tile[0];
        tile_nx = tile;
        tile_nx2 = tile;
        if (tid == 0 && tile < n_tiles) tile_nx2 = draw1();
    } else {
        tile = draw();
        tile_nx = (tile < n_tiles) ? draw() : tile;
        tile_nx2 = tile_nx;
    }
    unsigned p_next = (tile < n_tiles) ? parcel_of(tile_pos(tile)) : 0u;
    // next parcel's moments held in registers across the node loops: the 128-register box instances of small tensors (-0.7 % on
    // C5; the column instance spills 48 bytes with it and measured 0.72 instead of 0.69 ms per C3 step)
    constexpr bool REGPF = tpp_ztab(P) && tpp_min_blocks(N, P, MODEL) * TPP_THREADS >= 512 && !RAIN;
    double nxt[N][3];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            nxt[i][q] = (REGPF && tile < n_tiles && q < cfg.nprog[i]) ? args.u_in[(unsigned)(cfg.slot0[i] + q) * s_in + p_next * ps_in] : 0.0;
    for (; tile < n_tiles; tile = tile_nx, tile_nx = (BSYNC ? tile_nx : tile_nx2)) {
        const long long base = tile_pos(tile);
        bool draw_more;
        if constexpr (BSYNC) {
            if (tid == 0) s_tile[(it + 1) & 1] = tile_nx2;
            __syncthreads();
            tile_nx = s_tile[(it + 1) & 1];
            ++it;
            draw_more = tile_nx < n_tiles;
            if (tid == 0 && draw_more) tile_nx2 = draw1();
        } else {
            draw_more = tile_nx < n_tiles;
            if (draw_more) tile_nx2 = draw();
        }
        const long long idx = base + (tid & 31);
        const bool live = idx < n;
        const unsigned p = p_next;
        double cur[N][3];
        // this parcel's moments.  REGPF instances hold the NEXT parcel's values in registers across the node loops (possible since
        // the Z sums left the loops: 0 spill bytes); the others request them one iteration ahead with an L2 prefetch (holding
        // them in registers made those instances spill them, and the spill store waits for the DRAM load)
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if constexpr (REGPF) cur[i][q] = nxt[i][q];
                else cur[i][q] = (q < cfg.nprog[i]) ? args.u_in[(unsigned)(cfg.slot0[i] + q) * s_in + p * ps_in] : 0.0;
            }
        if (draw_more) {
            p_next = parcel_of(tile_pos(tile_nx));
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (q < cfg.nprog[i]) {
                        if constexpr (REGPF) nxt[i][q] = args.u_in[(unsigned)(cfg.slot0[i] + q) * s_in + p_next * ps_in];
                        else asm volatile("prefetch.global.L2 [%0];" ::"l"(args.u_in + ((unsigned)(cfg.slot0[i] + q) * s_in + p_next * ps_in)));
                    }
        }
        // zero flux above the column top (rainshaft_helpers.jl:80-81); one 64-bit modulo per cell
        const bool top_level = RAIN && ((p + 1u) % (unsigned)cfg.nz == 0u);
        if constexpr (RAIN) {
            // the sedimentation fluxes of this cell and of the cell above are read at the very end: request them now
            if (live && args.flux != nullptr) {
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        if (q < cfg.nprog[i]) {
                            const double* fp = args.flux + ((unsigned)(cfg.slot0[i] + q) * s_flux + p);
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(fp));
                            if (!top_level) asm volatile("prefetch.global.L1 [%0];" ::"l"(fp + 1));
                        }
            }
        }

        // ---- load, (clip), normalise, parameters, moment matrix --------------------------------------
        double raw[N][3];
        double mnv[N][3];  // normalised moments
        bool cell_empty = RAIN;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int s0 = cfg.slot0[i], np = cfg.nprog[i];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                raw[i][q] = 0.0;
                mnv[i][q] = 0.0;
                if (q < np) {
                    double v = cur[i][q];
                    if (RAIN) {
                        v = (v < 0.0) ? 0.0 : v;  // rainshaft_helpers.jl:52
                        if (args.clip_back != nullptr && live) args.clip_back[(unsigned)(s0 + q) * s_clip + p] = v;
                    }
                    raw[i][q] = v;
                    // (an inline IEEE division sends every zero dividend through its ~100-instruction slow path: 13 % of the
                    // rainshaft instance's executed instructions when 61 % of the C3 cells are exact zeros)
                    mnv[i][q] = norm_div(v, cfg.norm[s0 + q], cfg.inv_norm[s0 + q]);
                    if (RAIN) cell_empty = cell_empty && (mnv[i][q] < kEps);  // rainshaft_helpers.jl:67
                }
            }
        }

        double res[N][3];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int q = 0; q < 3; ++q) res[i][q] = 0.0;

        // rainshaft: a warp whose cells are all empty (rainshaft_helpers.jl:67-68) has zero coalescence source: it skips the
        // distribution parameters, the moment matrix and the whole contraction (with the regime sort empty cells share warps;
        // 61 % of the C3 cells are empty) and goes straight to the flux divergence and the stage update
        const bool warp_idle = RAIN && __all_sync(0xffffffffu, cell_empty || !live);
        if constexpr (RAIN) {
            if (warp_idle) {
                // Compact path of the empty warps: flux divergence + stage update (all loads first), next to the
                // top of the parcel loop, so that the 61 % of warps that never enter the contraction run from a few KB of code
                // instead of jumping across the kernel's 130 KB (instruction fetch was 29 % of this instance's stall samples).
                // Same expressions, in the same order, as the general epilogue below with f_coal = 0.
                if (live) {
                    const double inv_dz_c = 1.0 / cfg.dz;
                    const bool with_un = !args.tend_only && args.u_n != nullptr;
                    double unc[N][3], flc[N][3], fuc[N][3];
#pragma unroll
                    for (int k = 0; k < N; ++k)
#pragma unroll
                        for (int m = 0; m < 3; ++m) {
                            unc[k][m] = 0.0; flc[k][m] = 0.0; fuc[k][m] = 0.0;
                            if (m < cfg.nprog[k]) {
                                const int s = cfg.slot0[k] + m;
                                flc[k][m] = args.flux[(unsigned)s * s_flux + p];
                                if (!top_level) fuc[k][m] = args.flux[(unsigned)s * s_flux + p + 1u];
                                if (with_un) unc[k][m] = args.u_n[(unsigned)s * s_n + p];
                            }
                        }
#pragma unroll
                    for (int k = 0; k < N; ++k)
#pragma unroll
                        for (int m = 0; m < 3; ++m)
                            if (m < cfg.nprog[k]) {
                                const int s = cfg.slot0[k] + m;
                                const double f = 0.0 + (-(fuc[k][m] - flc[k][m]) * inv_dz_c);
                                double o;
                                if (args.tend_only) {
                                    o = f;
                                } else {
                                    double acc2 = args.ci * raw[k][m];
                                    if (args.u_n != nullptr) {
                                        double un = unc[k][m];
                                        un = (un < 0.0) ? 0.0 : un;
                                        acc2 = args.cn * un + acc2;
                                    }
                                    const double num = acc2 + args.cf * (args.dt * f);
                                    o = (num == 0.0) ? num : div_rn_outofline(num, args.div);  // zero dividend: see the normalisation above
                                    o = (o < 0.0) ? 0.0 : o;
                                }
                                args.out[(unsigned)s * s_out + p * ps_out] = o;
                            }
                }
                continue;
            }
        }
        double sumQ[N][3], sumR[N][3];
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int m = 0; m < 3; ++m) { sumQ[k][m] = 0.0; sumR[k][m] = 0.0; }
        if (!warp_idle) {
        // ---- parameters, moment matrix ------------------------------------------------------------------
        double mom[N][M];
        double pn[N], pa[N], pb[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int np = cfg.nprog[i], kind = cfg.kind[i];
            ModeParams mp;
            if (args.params_in) {
                mp.n = raw[i][0]; mp.a = raw[i][1]; mp.b = (np > 2) ? raw[i][2] : 1.0; mp.invalid = 0;
            } else {
                mp = params_from_moments(kind, mnv[i][0], mnv[i][1], mnv[i][2], kind == CLOUDY_GAMMA ? cfg.k_lo : -INFINITY,
                                         kind == CLOUDY_GAMMA ? cfg.k_hi : INFINITY);
            }
            if (mp.invalid && live && args.err_count != nullptr) atomicAdd(args.err_count, 1ULL);
            pn[i] = mp.n; pa[i] = mp.a; pb[i] = mp.b;
            double mq = mp.n;
#pragma unroll
            for (int q = 0; q < M; ++q) {  // Coalescence.jl:187-198
                double val = mq;
                if (kind == CLOUDY_LOGNORMAL) val = lognormal_moment_int(mp.n, mp.a, mp.b, q);
                mom[i][q] = (q < cfg.n_mom_max) ? val : 0.0;
                if (kind == CLOUDY_GAMMA) mq *= mp.a * (mp.b + q);
                else if (kind == CLOUDY_EXPONENTIAL) mq *= mp.a * (q + 1.0);
                else mq *= mp.a;
            }
        }
        // ---- S terms (self-collisions): truncated integrals of every mode, contracted at once ----------
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s1[3], s2[3];
            double Fs[M * (M + 1) / 2];  // this mode's truncated integrals F(u, v), u <= v
            const int n2d = cfg.n2d[i];
            bool done = false;
            if (i < N - 1 && (cfg.quad[i] || cfg.ln_thr[i])) {
                const double nmd = pn[i], th = pa[i], k = pb[i];
                const bool skip = (nmd == 0.0) || cell_empty || !live;
                if (!__all_sync(0xffffffffu, skip)) {
                    const int Mp = cfg.Mp[i];
                    // the truncated integrals H (raw quadrature sums scaled by `scale[p2]`), limited and masked, into Fs
                    auto contract = [&](auto mp_tag, auto& F, const double* scale) {
                        constexpr int MP = decltype(mp_tag)::value;
                        static_assert(MP == M, "all M = P + 2 orders are carried");
                        // F = 0 | min(Mom*Mom, H) — Coalescence.jl:212-227; zero beyond N_2d_ints (Coalescence.jl:213)
#pragma unroll
                        for (int p1 = 0; p1 < MP; ++p1)
#pragma unroll
                            for (int p2 = p1; p2 < MP; ++p2) {
                                const int t = tri_ct(p1, p2, MP);
                                const double mm = mom[i][p1] * mom[i][p2];
                                const double H = scale[p2] * F[t];
                                Fs[t] = (mm < kEps || p2 >= Mp) ? 0.0 : jl_min(mm, H);
                            }
                    };
                    if (cfg.ln_thr[i]) {
                        // Lognormal: (n, μ, σ) = (nmd, th, k)
                        auto finish_ln = [&](auto mp_tag) {
                            constexpr int MP = decltype(mp_tag)::value;
                            double F[MP * (MP + 1) / 2];
                            tpp_lognormal_H<MP>(F, sTab + (MOVING ? cfg.gl_off : cfg.gl2_off), cfg.gl_n, nmd, th, k, cfg.thr[i]);
                            double one[MP];
#pragma unroll
                            for (int pp = 0; pp < MP; ++pp) one[pp] = 1.0;
                            contract(mp_tag, F, one);
                        };
                        finish_ln(std::integral_constant<int, M>{});
                    } else {
                        const double inv_th = 1.0 / th;
                        // Gamma(k): the MovingThreshold instances need it itself (inverse incomplete gamma), the others only 1/Gamma(k)
                        // (the FixedThreshold instances of small tensors read 1/Γ(k) from the Z-sum records after the node loops)
                        double inv_gk = (MOVING || tpp_ztab(P)) ? 1.0 : ((cfg.kind[i] == CLOUDY_GAMMA) ? gamma_shape_inv(k) : 1.0);
                        const double gk = MOVING ? ((cfg.kind[i] == CLOUDY_GAMMA) ? gamma_shape(k) : 1.0) : 1.0;
                        // All M orders are carried even when N_2d_ints[i] = M - 1 (two adjacent 2-moment modes): the downward
                        // recurrence from the higher top order gives the same lower orders, the unused top entries are masked in
                        // `contract`, and the node loop exists once per mode instead of twice (code size)
                        const double a_top = k + (double)(M - 1);
                        // threshold: run-constant, or this parcel's own percentile (compute_threshold, ParticleDistributions.jl:747-761)
                        double thr = cfg.thr[i];
                        OwnGrid og;
                        og.nb = 3; og.dx = 0.0;
                        bool own_grid = false;  // MovingThreshold with x_th > 1: the grid's shape depends on the parcel
                        if constexpr (MOVING) {
                            const double pct = cfg.thr[i];
                            const double Xp = (cfg.kind[i] == CLOUDY_GAMMA)
                                                  ? igam_inv_tab(k, pct, gk, cfg.tab + cfg.xp_off[i], cfg.xp_n, cfg.xp_k0, cfg.xp_inv_h, sh.deg)
                                                  : -log(1.0 - pct);
                            thr = fmax(th * Xp, 1e-18);
                            // x_lb = min(1e-5, 1e-5 x_th) (ParticleDistributions.jl:579): for x_th <= 1 the grid is x_th times a
                            // parcel-independent unit grid (5 decades below the threshold, 5*bins_per_log_unit nodes); above 1
                            // the lower bound stays at 1e-5 and the grid is the parcel's own
                            own_grid = thr > 1.0 && !skip;
                            if (own_grid) {
                                // n_bins = floor(n_bins_per_log_unit * log10(x_th / x_lb)), the reference's formula literally
                                // (ParticleDistributions.jl:580); a node count outside the supported range is counted as an
                                // invalid parcel (cloudy_error_count) instead of being evaluated on a made-up grid
                                const double nbf = floor((double)cfg.bins_per_log_unit * log10(thr / 1e-5));
                                const bool nb_ok = nbf >= 3.0 && nbf < 65536.0;
                                if (!nb_ok && live && args.err_count != nullptr) atomicAdd(args.err_count, 1ULL);
                                og.nb = nb_ok ? (int)nbf : 3;
                                og.dx = (log(thr) - (-11.512925464970229)) / (double)og.nb;  // log(1e-5)
                            }
                        }
                        // own series degree / continued-fraction depth; loop bounds are the warp maxima
                        const double X = thr * inv_th;
                        const int ai = series_a_bin(a_top);
                        const double ser_lim = sh.serlim[ai];
                        const int zi = series_z_bin(fmin(X, ser_lim - 0.5));
                        const int deg = skip ? 1 : max((int)sh.deg[zi][ai], 1);
                        const int deg_w = __reduce_max_sync(0xffffffffu, deg);
                        const int cfd = sh.cfd[ai];
                        const int cfd_w = __reduce_max_sync(0xffffffffu, cfd);
                        if constexpr (MOVING) {   // c_n = 1/(a)_{n+1}: one division, then c_{n-1} = c_n (a+n)
                            double pe = 1.0, po = 1.0;  // two independent product chains (even / odd factors)
                            for (int nn = 0; nn + 1 <= deg; nn += 2) {
                                pe *= (a_top + (double)nn);
                                po *= (a_top + (double)(nn + 1));
                            }
                            if ((deg & 1) == 0) pe *= (a_top + (double)deg);
                            double cc = 1.0 / (pe * po);  // c_deg
                            // c_{n-1} = c_n (a+n): two interleaved chains stepping by two
                            double c1 = cc * (a_top + (double)deg);  // c_{deg-1}
                            int nn = deg;
                            for (; nn >= 2; nn -= 2) {
                                myCt[nn * TPP_THREADS] = cc;
                                myCt[(nn - 1) * TPP_THREADS] = c1;
                                const double f = (a_top + (double)nn) * (a_top + (double)(nn - 1));
                                const double f1 = (a_top + (double)(nn - 1)) * (a_top + (double)(nn - 2));
                                cc *= f;   // c_{nn-2}
                                c1 *= f1;  // c_{nn-3}
                            }
                            if (nn == 1) { myCt[TPP_THREADS] = cc; myCt[0] = c1; }
                            else myCt[0] = cc;
                            for (int nn = deg + 1; nn <= deg_w; ++nn) myCt[nn * TPP_THREADS] = 0.0;
                        }
                        auto finish = [&](auto mp_tag) {
                            constexpr int MP = decltype(mp_tag)::value;
                            // ia[p] = 1/(k+p) for p < MP-1 from ONE division (prefix and suffix products; k >= eps keeps the product
                            // normal); Γ(k+MP-1) = Γ(k) (k)_{MP-1} is needed by the continued-fraction path only
                            double ia[MP], pre[MP], suf[MP];
                            pre[0] = 1.0;
#pragma unroll
                            for (int pp = 1; pp < MP; ++pp) pre[pp] = pre[pp - 1] * (k + (double)(pp - 1));  // pre[p] = prod_{q<p}(k+q)
                            suf[MP - 1] = 1.0; suf[MP - 2] = 1.0;
#pragma unroll
                            for (int pp = MP - 3; pp >= 0; --pp) suf[pp] = suf[pp + 1] * (k + (double)(pp + 1));  // prod_{p<q<MP-1}(k+q)
                            const double poch = pre[MP - 1];  // (k)_{MP-1}
                            const double inv_poch_k = 1.0 / poch;
#pragma unroll
                            for (int pp = 0; pp < MP - 1; ++pp) ia[pp] = inv_poch_k * (pre[pp] * suf[pp]);
                            ia[MP - 1] = 0.0;
                            const double gam_top = MOVING ? gk * poch : poch;  // FixedThreshold: divided by inv_gk where it is used
                            double F[MP * (MP + 1) / 2];
#pragma unroll
                            for (int t = 0; t < MP * (MP + 1) / 2; ++t) F[t] = 0.0;
                            TableGrid tg;
                            tg.rec = sTab + cfg.rec_off[i];
                            tg.stride = REC_W + M;
                            tg.n_near = cfg.rec_near[i];
                            tg.n_far = cfg.rec_far[i];
                            NodeCommon<MP> C;
                            C.k = k; C.X = X; C.gam_top = gam_top; C.a_top = a_top; C.ser_lim = ser_lim;
                            C.deg_w = deg_w; C.cfd_w = cfd_w; C.cfd = cfd; C.myCt = myCt; C.exp_tab = sh.exp32;
#pragma unroll
                            for (int pp = 0; pp < MP; ++pp) C.ia[pp] = ia[pp];
                            if constexpr (MOVING) {
                                // lengths in units of the parcel's threshold: x_j = x_th ρ_j, z_j = (1-ρ_j) X,
                                // ln x_j + ln(x_th-x_j) - 2 ln θ = ln ρ_j + ln(1-ρ_j) + 2 ln X, weights w_j dx ρ_j^p1 (x x_th^p1 below)
                                C.zs = X; C.log_u = -log(X);
                                C.init();
                                const bool unit_grid = !own_grid && !skip;
                                const bool any_own = __any_sync(0xffffffffu, own_grid);
                                const bool any_unit = __any_sync(0xffffffffu, unit_grid);
                                const int bpl = cfg.bins_per_log_unit;
                                og.n_far_w = (bpl + tpp_npl(P) - 1) / tpp_npl(P) * tpp_npl(P);
                                const int nb_max = __reduce_max_sync(0xffffffffu, own_grid ? og.nb : 0);
                                og.nb_min_w = __reduce_min_sync(0xffffffffu, own_grid ? og.nb : 0x7fffffff);
                                og.n_near_w = (max(nb_max - og.n_far_w, 0) + tpp_npl(P) - 1) / tpp_npl(P) * tpp_npl(P);
                                og.m8 = 2 * bpl; og.m4 = 4 * bpl;
                                og.exp_tab = sh.exp32; og.kdeg_m = sh.kdeg_m;
                                // far zones need the c_n table, the near zones overwrite it with the Taylor coefficients
                                if (any_own) tpp_zone<MP, P, false, true>(F, own_grid, og, C);
                                if (any_unit) tpp_zone<MP, P, false, true>(F, unit_grid, tg, C);
                                tpp_taylor_coeffs<MP>(C);
                                if (any_own) {
                                    tpp_zone<MP, P, true, true>(F, own_grid, og, C);
                                    tpp_cf_nodes<MP, P, true>(F, own_grid, og, C);
                                }
                                if (any_unit) {
                                    tpp_zone<MP, P, true, true>(F, unit_grid, tg, C);
                                    tpp_cf_nodes<MP, P, true>(F, unit_grid, tg, C);
                                }
                                double sc = 1.0;
#pragma unroll
                                for (int p1 = 1; p1 < MP; ++p1) {
                                    sc *= thr;
#pragma unroll
                                    for (int p2 = p1; p2 < MP; ++p2) F[tri_ct(p1, p2, MP)] *= sc;
                                }
                            } else {
                                FixedGrid fg;
                                fg.rec = sTab + cfg.rec2_off[i];
                                fg.kblk = sTab + cfg.kblk2_off[i];
                                fg.cls_end = cfg.near_cls_end[i];
                                fg.n_near = cfg.rec_near[i];
                                fg.n_far = cfg.rec2_far[i];
                                fg.soa = cfg.tab + cfg.tab_off[i];
                                fg.nb = cfg.n_bins[i];
                                tpp_nodes_fixed2<MP, P>(F, fg, k, inv_th, log(th), X, gam_top, inv_gk, ia, myCt, deg, sh.cfdz[ai], a_top, ser_lim, sh.exp32,
                                                       cfg.tab + cfg.zt_off[i], cfg.zt_inv_h, cfg.zt_n, cfg.zt_L[i], cfg.kind[i] == CLOUDY_GAMMA);
                            }
                            double thp[MP];  // H = n^2 θ^{p2}/Γ(k)^2 * sum
                            thp[0] = MOVING ? nmd * nmd / (gk * gk) : (nmd * inv_gk) * (nmd * inv_gk);
#pragma unroll
                            for (int pp = 1; pp < MP; ++pp) thp[pp] = thp[pp - 1] * th;
                            contract(mp_tag, F, thp);
                        };
                        finish(std::integral_constant<int, M>{});
                    }
                    done = true;
                }
            }
            if (!done) {
                const bool mono = (i < N - 1) && cfg.mono_thr[i];
                const bool quad_skipped = (i < N - 1) && (cfg.quad[i] || cfg.ln_thr[i]);  // whole warp empty: F = 0
                const double th = pa[i], nn = pn[i];
                const bool below = th < cfg.thr[i] / 2;
                double hpow[2 * P + 1];  // Monodisperse: n^2 θ^(u+v) below the threshold (ParticleDistributions.jl:557-564)
                hpow[0] = nn * nn;
#pragma unroll
                for (int e = 1; e <= 2 * P; ++e) hpow[e] = hpow[e - 1] * th;
#pragma unroll
                for (int x = 0; x < M; ++x)
#pragma unroll
                    for (int y = x; y < M; ++y) {
                        if (x + y > 2 * P) continue;
                        const double mm = mom[i][x] * mom[i][y];
                        double f = mm;  // last mode or infinite threshold
                        if (mono) f = jl_min(mm, below ? hpow[x + y] : 0.0);
                        if (mm < kEps || y >= n2d || quad_skipped) f = 0.0;
                        Fs[tri_ct(x, y, M)] = f;
                    }
            }
            tpp_s_terms<P>(cfg, i, mom[i], Fs, s1, s2);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                res[i][m] += s1[m];
                if (i + 1 < N) res[i + 1][m] += s2[m];
            }
        }

        // ---- Q and R (collisions between modes) — Coalescence.jl:260-351 -----------------------------
        // U_jk[c][b] = sum_a c^{jk}_ab Mom_j[a+c];  R_jk(m) = sum_b Mom_k[b+m] U[0][b];
        // Q_jk(m) = sum_c C(m,c) sum_b Mom_k[b+m-c] U[c][b]  (j < k)
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double U[3][P];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int b = 0; b < P; ++b) {
                        double u = 0.0;
                        if (c == 0 || j < k) {
#pragma unroll
                            for (int a = 0; a < P; ++a) u = fma(cfg.c[j][k][a][b], mom[j][a + c], u);
                        }
                        U[c][b] = u;
                    }
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    double r = 0.0;
#pragma unroll
                    for (int b = 0; b < P; ++b) r = fma(mom[k][b + m], U[0][b], r);
                    sumR[k][m] += r;
                    if (j < k) {
                        double q = 0.0;
#pragma unroll
                        for (int c = 0; c <= m; ++c) {
                            double qc = 0.0;
#pragma unroll
                            for (int b = 0; b < P; ++b) qc = fma(mom[k][b + m - c], U[c][b], qc);
                            q += ((m == 2 && c == 1) ? 2.0 : 1.0) * qc;
                        }
                        sumQ[k][m] += q;
                    }
                }
            }
        }

        }  // !warp_idle

        // ---- assemble, combine with the stage update, store ---------------------------------------------
        // every global load of the epilogue (u^n, the two flux values per slot) is issued before the first use: the out-of-line
        // division is a scheduling barrier, and one exposed L2 round trip per slot was 10 % of the rainshaft instance's samples
        const double inv_dz = 1.0 / cfg.dz;
        double unv[N][3], flv[N][3], fluv[N][3];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int s0 = cfg.slot0[k], np = cfg.nprog[k];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                unv[k][m] = 0.0; flv[k][m] = 0.0; fluv[k][m] = 0.0;
                if (m < np && live) {
                    const int s = s0 + m;
                    if (!args.tend_only && args.u_n != nullptr) unv[k][m] = args.u_n[(unsigned)s * s_n + p];
                    if (RAIN) {
                        // flux of this cell and of the cell above it, written by flux_kernel.  (Evaluating both inside this kernel was
                        // measured: 0.52 ms instead of 0.42 ms per RHS on C3 — the regime-sorted order separates vertical neighbours,
                        // so every flux would be computed twice, ~1000 instructions each.)
                        flv[k][m] = args.flux[(unsigned)s * s_flux + p];
                        fluv[k][m] = top_level ? 0.0 : args.flux[(unsigned)s * s_flux + p + 1u];
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int s0 = cfg.slot0[k], np = cfg.nprog[k];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                if (m < np && live) {
                    const int s = s0 + m;
                    // sum(Q[:,k]) - sum(R[:,k]) + S[1,k] (+ S[2,k-1]) — Coalescence.jl:140-149 (res holds S1 + S2)
                    double f = sumQ[k][m] - sumR[k][m] + res[k][m];
                    if (!args.params_in) f *= cfg.norm[s];
                    if (RAIN) {
                        if (cell_empty) f = 0.0;
                        f = f + (-(fluv[k][m] - flv[k][m]) * inv_dz);  // rainshaft_helpers.jl:83-87
                    }
                    double o;
                    if (args.tend_only) {
                        o = f;
                    } else {
                        double acc2 = args.ci * raw[k][m];
                        if (args.u_n != nullptr) {
                            double un = unv[k][m];
                            if (RAIN) un = (un < 0.0) ? 0.0 : un;
                            acc2 = args.cn * un + acc2;
                        }
                        const double num = acc2 + args.cf * (args.dt * f);
                        o = (num == 0.0) ? num : div_rn_outofline(num, args.div);  // zero dividend: see the normalisation above
                        if (RAIN) o = (o < 0.0) ? 0.0 : o;
                    }
                    args.out[(unsigned)s * s_out + p * ps_out] = o;
                }
            }
        }
    }
}

typedef void (*tpp_fn)(const DevConfig, const KArgs);
// returns nullptr when (N, P) has no thread-per-parcel instance (the generic lane-cooperative kernel is used)
tpp_fn tpp_lookup(int N, int P, int model);

}  // namespace cloudy
