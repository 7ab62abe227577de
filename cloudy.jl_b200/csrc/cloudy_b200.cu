// libcloudy_b200.so — batched Cloudy.jl coalescence / sedimentation tendencies on B200 (sm_100a).
//
// One CUDA kernel template (rhs_kernel) evaluates, for a tile of parcels (box model) or column cells
// (rainshaft), everything the reference's ODE right-hand side does per parcel
//   test/examples/utils/box_model_helpers.jl:29-53   (rhs_coal!)
//   test/examples/utils/rainshaft_helpers.jl:47-88   (make_rainshaft_rhs)
//   src/Sources/Coalescence.jl:115-455               (get_coal_ints, AnalyticalCoalStyle)
//   src/ParticleDistributions/ParticleDistributions.jl:456-612, :698-710
//   src/Sources/Sedimentation.jl:22-37
// and, when asked, applies one SSPRK33 stage update in the same pass (state read once, written once).
//
// Execution shape: LANES (4..32) lanes of a warp cooperate on one parcel.  Quadrature nodes of the
// reference's log-spaced end-corrected Simpson rule are strided over the lanes; the T = M'(M'+1)/2
// truncated-moment accumulators are combined with an xor-butterfly of warp shuffles.  Per node ONE
// incomplete gamma function is evaluated at the top order (fixed-trip Horner series or fixed-depth
// continued fraction, per-parcel coefficient tables in shared memory) and the lower orders follow
// from the downward recurrence gamma(a,z) = (gamma(a+1,z) + z^a e^-z)/a.  The polynomial-kernel
// contraction (Q/R/S) runs one output moment per lane from shared memory.  FP64 CUDA cores only.
#include <cuda_runtime.h>
#include <cstddef>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/cloudy_b200.h"
#include "common.cuh"
#include "tpp_kernel.cuh"
#include "regime_sort.cuh"

namespace cloudy {

// quadrature nodes each lane keeps in flight (independent Horner chains sharing the coefficient loads)
template <int LANES>
struct NodesPerLane {
    static constexpr int value = (LANES <= 4) ? 4 : (LANES == 8 ? 2 : 1);
};

template <int LANES>
struct Shape {
    static constexpr int kThreads = (LANES >= 8) ? 256 : 32 * LANES;
    static constexpr int kGroups = kThreads / LANES;  // parcels per tile (<= 32)
};

// per-group shared-memory record (doubles)
constexpr int PAR_N = 0, PAR_TH = 1, PAR_K = 2, PAR_INVTH = 3, PAR_LOGTH = 4, PAR_IGK2 = 5, PAR_GK = 6, PAR_STRIDE = 8;
constexpr int TAB_CT = 0;                            // series coefficients c_n, n = 0..63
constexpr int TAB_CF = 64;                           // cf numerators n(a-n), n = 1..16 at [64+n]
constexpr int TAB_IA = TAB_CF + kCfMaxDepth + 1;     // 1/(k+p), p = 0..5
constexpr int TAB_LEN = TAB_IA + MAXM;               // 88
constexpr int TAB_STRIDE = TAB_LEN + 1;              // odd → groups of a warp hit different banks

struct GroupLayout {
    int mom, par, F, tab, in, un, out, flux, total;
};
__host__ __device__ inline GroupLayout group_layout(int N, int M, int nslots, bool rain) {
    GroupLayout g;
    int o = 0;
    g.mom = o; o += N * M;
    g.par = o; o += N * PAR_STRIDE;
    g.F = o; o += (N > 1 ? (N - 1) : 1) * MAXT;
    g.tab = o; o += TAB_STRIDE;
    g.in = o; o += nslots;
    g.un = o; o += nslots;
    g.out = o; o += nslots;
    g.flux = o; o += rain ? nslots : 0;
    g.total = o | 1;  // odd stride
    return g;
}

__device__ __forceinline__ int tri_index(int p1, int p2, int Mcols) {  // p1 <= p2 < Mcols, row-major upper triangle
    return p1 * Mcols - (p1 * (p1 - 1)) / 2 + (p2 - p1);
}

// ------------------------------------------------------------------------------------------------
// node loop: T accumulators of  sum_j W[p1][j] * g_j * gamma(k+p2, z_j),  p1 <= p2 < Mp
// ------------------------------------------------------------------------------------------------
template <int MPMAX, int LANES, int NPL>
__device__ __forceinline__ void node_integrals(double (&acc)[MPMAX * (MPMAX + 1) / 2], const double* __restrict__ tb, int nb, int M,
                                               int Mp, double k, double inv_th, double log_th, double gam_top,
                                               const double* __restrict__ gtab, int deg_w, int cfd_w, int cfd, double ser_lim, int lane) {
    constexpr int T = MPMAX * (MPMAX + 1) / 2;
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
    const double* XJ = tb;
    const double* ELL = tb + nb;
    const double* TMX = tb + 2 * nb;
    const double* LZ = tb + 3 * nb;
    const double* W = tb + 4 * nb;
    const double a_top = k + (double)(Mp - 1);
    const int batches = (nb + LANES * NPL - 1) / (LANES * NPL);
    for (int bt = 0; bt < batches; ++bt) {
        // NPL nodes per lane in flight: their Horner chains share every coefficient load and hide DFMA latency
        int jj[NPL];
        bool valid[NPL];
        double z[NPL], gtop[NPL];
        bool any_ser = false, any_cf = false;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int j = (bt * NPL + i) * LANES + lane;
            valid[i] = j < nb;
            jj[i] = valid[i] ? j : nb - 1;
            z[i] = TMX[jj[i]] * inv_th;  // (x_th - x_j)/θ
            any_ser = any_ser || (z[i] < ser_lim);
            any_cf = any_cf || !(z[i] < ser_lim);
            gtop[i] = 0.0;
        }
        if (__any_sync(0xffffffffu, any_ser)) {
            double s[NPL];
            const double c_top = gtab[TAB_CT + deg_w];
#pragma unroll
            for (int i = 0; i < NPL; ++i) s[i] = c_top;
            int n = deg_w - 1;
            for (; n >= 3; n -= 4) {
                const double c0 = gtab[TAB_CT + n], c1 = gtab[TAB_CT + n - 1], c2 = gtab[TAB_CT + n - 2], c3 = gtab[TAB_CT + n - 3];
#pragma unroll
                for (int i = 0; i < NPL; ++i) s[i] = fma(s[i], z[i], c0);
#pragma unroll
                for (int i = 0; i < NPL; ++i) s[i] = fma(s[i], z[i], c1);
#pragma unroll
                for (int i = 0; i < NPL; ++i) s[i] = fma(s[i], z[i], c2);
#pragma unroll
                for (int i = 0; i < NPL; ++i) s[i] = fma(s[i], z[i], c3);
            }
            for (; n >= 0; --n) {
                const double c0 = gtab[TAB_CT + n];
#pragma unroll
                for (int i = 0; i < NPL; ++i) s[i] = fma(s[i], z[i], c0);
            }
#pragma unroll
            for (int i = 0; i < NPL; ++i) gtop[i] = s[i];  // series sum; multiplied by E z^(Mp-1) below
        }
        if (__any_sync(0xffffffffu, any_cf)) {
            // Legendre continued fraction of the upper function, forward recurrence, fixed depth
            double Pm[NPL], Pc[NPL], Qm[NPL], Qc[NPL], b[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                const double zc = fmin(z[i], 256.0);  // beyond this the upper function is < 1e-80 of Gamma(a)
                b[i] = zc + 1.0 - a_top;
                Pm[i] = 1.0; Pc[i] = b[i]; Qm[i] = 0.0; Qc[i] = 1.0;
            }
            for (int n = 1; n <= cfd_w; ++n) {
                const double an = -gtab[TAB_CF + n];  // -n(n-a)
                const bool on = n <= cfd;             // own depth only (neighbour-independent result)
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    b[i] += 2.0;
                    const double Pn = fma(b[i], Pc[i], an * Pm[i]);
                    const double Qn = fma(b[i], Qc[i], an * Qm[i]);
                    if (on) { Pm[i] = Pc[i]; Pc[i] = Pn; Qm[i] = Qc[i]; Qc[i] = Qn; }
                }
            }
#pragma unroll
            for (int i = 0; i < NPL; ++i)
                if (!(z[i] < ser_lim)) gtop[i] = -(Qc[i] / Pc[i]);  // negative marks "upper function ratio"
        }
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int j = jj[i];
            const double u = XJ[j] * inv_th;
            const double g = exp(fma(k, ELL[j] - log_th, -u));     // (x_j/θ)^k e^{-x_j/θ}
            const double E = exp(fma(k, LZ[j] - log_th, -z[i]));   // z^k e^{-z}
            double zp[MPMAX];
            zp[0] = 1.0;
#pragma unroll
            for (int p = 1; p < MPMAX; ++p) zp[p] = zp[p - 1] * z[i];
            double Etop = E;
#pragma unroll
            for (int p = 1; p < MPMAX; ++p) Etop = (p < Mp) ? Etop * z[i] : Etop;  // E * z^(Mp-1)
            const double top = (z[i] < ser_lim) ? Etop * gtop[i] : fma(Etop, gtop[i], gam_top);
            double gam[MPMAX];
#pragma unroll
            for (int p = MPMAX - 1; p >= 0; --p) {
                if (p == Mp - 1) gam[p] = top;
                else if (p < Mp - 1) gam[p] = (gam[p + 1] + E * zp[p]) * gtab[TAB_IA + p];
                else gam[p] = 0.0;
            }
            const double gv = valid[i] ? g : 0.0;
            int t = 0;
#pragma unroll
            for (int p1 = 0; p1 < MPMAX; ++p1) {
                const double wg = (p1 < Mp) ? W[p1 * nb + j] * gv : 0.0;
#pragma unroll
                for (int p2 = p1; p2 < MPMAX; ++p2) {
                    acc[t] = fma(wg, gam[p2], acc[t]);
                    ++t;
                }
            }
        }
    }
    // butterfly over the LANES of the group
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], off);
    }
}

// F_k[x][y] accessor — Coalescence.jl:200-244
__device__ __forceinline__ double F_entry(const DevConfig& cfg, const double* mom_k, const double* par_k, const double* F_k, int k,
                                          int x, int y) {
    const double mm = mom_k[x] * mom_k[y];
    if (mm < kEps || x >= cfg.n2d[k] || y >= cfg.n2d[k]) return 0.0;
    if (cfg.quad[k]) {
        const int lo = min(x, y), hi = max(x, y);
        return F_k[tri_index(lo, hi, cfg.Mp[k])];
    }
    if (cfg.mono_thr[k]) {  // ParticleDistributions.jl:557-564
        const double th = par_k[PAR_TH], n = par_k[PAR_N];
        double h = 0.0;
        if (th < cfg.thr[k] / 2) {
            h = n * n;
            for (int i = 0; i < x + y; ++i) h *= th;
        }
        return jl_min(mm, h);
    }
    return mm;  // last mode or infinite threshold
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MPMAX, int LANES, int MODEL>
__global__ void __launch_bounds__(Shape<LANES>::kThreads) rhs_kernel(const __grid_constant__ DevConfig cfg, const KArgs args) {
    constexpr int G = Shape<LANES>::kGroups;
    constexpr int THREADS = Shape<LANES>::kThreads;
    constexpr bool RAIN = (MODEL == MODEL_RAINSHAFT);
    constexpr int T = (MPMAX > 0) ? MPMAX * (MPMAX + 1) / 2 : 1;
    extern __shared__ double smem[];

    const int N = cfg.N, P = cfg.P, M = cfg.M, nslots = cfg.nslots;
    const int tid = threadIdx.x;
    const int lane = tid % LANES;
    const int grp = tid / LANES;
    const GroupLayout L = group_layout(N, M, nslots, RAIN);

    // block-level shared data: kernel tensors (compact [j][k][a][b]), grid tables, then group records
    double* sC = smem;
    double* sTab = sC + N * N * P * P;
    double* sGroups = sTab + cfg.tab_total;
    double* sHaloFlux = sGroups + (size_t)G * L.total;  // rainshaft: flux of the cell above the tile
    double* my = sGroups + (size_t)grp * L.total;

    for (int i = tid; i < N * N * P * P; i += THREADS) {
        int b = i % P, a = (i / P) % P, kk = (i / (P * P)) % N, jj = i / (P * P * N);
        sC[i] = cfg.c[jj][kk][a][b];
    }
    for (int i = tid; i < cfg.tab_total; i += THREADS) sTab[i] = cfg.tab[i];

    const long long n_tiles = (args.n + G - 1) / G;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p0 = tile * G;
        __syncthreads();  // previous tile fully consumed (and staging done on the first pass)
        // ---- load phase (coalesced: consecutive threads → consecutive parcels of one slot) ----
        const int n_load = RAIN ? G + 1 : G;
        for (int i = tid; i < nslots * n_load; i += THREADS) {
            const int s = i / n_load, g = i % n_load;
            const long long p = p0 + g;
            double v = 0.0, vn = 0.0;
            bool in_range = p < args.n;
            if (RAIN && g == G) in_range = in_range && (p % cfg.nz != 0);  // halo must be in the same column
            if (in_range) {
                v = args.u_in[s * args.s_in + p * args.ps_in];
                if (RAIN) v = (v < 0.0) ? 0.0 : v;  // rainshaft_helpers.jl:52
                if (args.u_n != nullptr && g < G) {
                    vn = args.u_n[s * args.s_n + p];
                    if (RAIN) vn = (vn < 0.0) ? 0.0 : vn;
                }
                if (RAIN && args.clip_back != nullptr && g < G) args.clip_back[s * args.s_clip + p] = v;
            }
            if (g < G) {
                double* rec = sGroups + (size_t)g * L.total;
                rec[L.in + s] = v;
                rec[L.un + s] = vn;
            } else {
                sHaloFlux[nslots + s] = v;  // halo cell's moments, staged after its flux slots
            }
        }
        __syncthreads();
        // ---- setup phase: lane ↔ mode ----
        const int n_cells = RAIN ? 2 : 1;  // own cell; rainshaft group 0 also does the halo cell
        for (int cell = 0; cell < n_cells; ++cell) {
            if (cell == 1 && grp != 0) break;
            const double* src = (cell == 0) ? (my + L.in) : (sHaloFlux + nslots);
            for (int md = lane; md < N; md += LANES) {
                const int kind = cfg.kind[md], s0 = cfg.slot0[md], np = cfg.nprog[md];
                const double m0 = src[s0] / cfg.norm[s0];
                const double m1 = src[s0 + 1] / cfg.norm[s0 + 1];
                const double m2 = (np > 2) ? src[s0 + 2] / cfg.norm[s0 + 2] : 0.0;
                ModeParams mp;
                if (args.params_in) {  // get_coal_ints(pdists, coal_data) on given distributions
                    mp.n = src[s0]; mp.a = src[s0 + 1]; mp.b = (np > 2) ? src[s0 + 2] : 1.0; mp.invalid = 0;
                } else {
                    mp = params_from_moments(kind, m0, m1, m2, kind == CLOUDY_GAMMA ? cfg.k_lo : -INFINITY,
                                             kind == CLOUDY_GAMMA ? cfg.k_hi : INFINITY);
                }
                if (mp.invalid && cell == 0 && args.err_count != nullptr && (p0 + grp) < args.n) atomicAdd(args.err_count, 1ULL);
                if (cell == 0) {
                    double* par = my + L.par + md * PAR_STRIDE;
                    double* mom = my + L.mom + md * M;
                    par[PAR_N] = mp.n; par[PAR_TH] = mp.a; par[PAR_K] = mp.b;
                    if (cfg.quad[md]) {
                        par[PAR_INVTH] = 1.0 / mp.a;
                        par[PAR_LOGTH] = log(mp.a);
                        const double gk = (kind == CLOUDY_GAMMA) ? tgamma(mp.b) : 1.0;
                        par[PAR_GK] = gk;
                        par[PAR_IGK2] = 1.0 / (gk * gk);
                    }
                    // moment matrix row, orders 0..M-1, zero beyond N_mom_max (Coalescence.jl:194)
                    double mq = mp.n;
                    for (int q = 0; q < M; ++q) {
                        double val;
                        if (kind == CLOUDY_LOGNORMAL) val = mp.n * exp(q * mp.a + (double)(q * q) * mp.b * mp.b / 2);
                        else val = mq;
                        mom[q] = (q < cfg.n_mom_max) ? val : 0.0;
                        if (kind == CLOUDY_GAMMA) mq *= mp.a * (mp.b + q);
                        else if (kind == CLOUDY_EXPONENTIAL) mq *= mp.a * (q + 1.0);
                        else mq *= mp.a;
                    }
                }
                if (RAIN) {  // sedimentation flux of this mode — Sedimentation.jl:22-37, rainshaft_helpers.jl:74-77
                    double* fl = (cell == 0) ? (my + L.flux) : sHaloFlux;
                    for (int q = 0; q < np; ++q) {
                        double sum = 0.0;
                        if (mp.n != 0.0) {
                            for (int v = 0; v < cfg.n_vel; ++v)
                                sum += -cfg.velv[v] * moment_real(kind, mp.n, mp.a, mp.b, (double)q + cfg.velb[v]);
                        } else {
                            for (int v = 0; v < cfg.n_vel; ++v) sum += -cfg.velv[v] * 0.0;
                        }
                        fl[s0 + q] = sum * cfg.norm[s0 + q];
                    }
                }
            }
        }
        __syncthreads();

        // empty-cell test of the rainshaft RHS (rainshaft_helpers.jl:67-68): all normalised moments < eps
        bool cell_empty = false;
        if (RAIN) {
            cell_empty = true;
            for (int s = 0; s < nslots; ++s) cell_empty = cell_empty && (my[L.in + s] / cfg.norm[s] < kEps);
        }

        if (!args.flux_only) {
            // ---- truncated 2-D integrals of every quadrature mode ----
            if (MPMAX > 0) {
                for (int md = 0; md < N - 1; ++md) {
                    if (!cfg.quad[md]) continue;
                    const double* par = my + L.par + md * PAR_STRIDE;
                    const double n_md = par[PAR_N];
                    const bool skip = (n_md == 0.0) || cell_empty;
                    double* Fk = my + L.F + md * MAXT;
                    const int Mp = cfg.Mp[md];
                    if (__all_sync(0xffffffffu, skip)) {
                        for (int t = lane; t < Mp * (Mp + 1) / 2; t += LANES) Fk[t] = 0.0;
                        __syncwarp();
                        continue;
                    }
                    const double th = par[PAR_TH], k = par[PAR_K], inv_th = par[PAR_INVTH], log_th = par[PAR_LOGTH];
                    const double a_top = k + (double)(Mp - 1);
                    double* gtab = my + L.tab;
                    // series degree needed by this group: largest z is below X = x_th/θ and below the series limit
                    const double X = cfg.thr[md] * inv_th;
                    const int ai = series_a_bin(a_top);
                    const double ser_lim = kSeriesLimit[ai];
                    const int zi = series_z_bin(fmin(X, ser_lim - 0.5));
                    // a parcel's result must not depend on which parcels share its warp: the coefficient table is
                    // built from the group's OWN degree (zero above it), only the loop bounds are warp maxima
                    const int deg = skip ? 1 : max((int)kSeriesDeg2[zi][ai], 1);
                    const int deg_w = __reduce_max_sync(0xffffffffu, deg);
                    const int cfd = kCfDepth[ai];
                    const int cfd_w = __reduce_max_sync(0xffffffffu, cfd);
                    // series coefficients c_n = 1/(a)_{n+1}, n = 0..deg, chunked over the lanes:
                    // local product → group scan → one division → downward fill
                    {
                        for (int n = lane; n <= kSeriesMaxDeg; n += LANES) gtab[TAB_CT + n] = 0.0;
                        __syncwarp();
                        const int C = (deg + LANES) / LANES;  // chunk length, covers 0..deg
                        const int n0 = lane * C;
                        double prod = 1.0;
                        for (int i = 0; i < C; ++i) prod *= (a_top + (double)(n0 + i));
                        double incl = prod;
#pragma unroll
                        for (int off = 1; off < LANES; off <<= 1) {
                            double o = __shfl_up_sync(0xffffffffu, incl, off, LANES);
                            if (lane >= off) incl *= o;
                        }
                        double cc = 1.0 / incl;  // c at the top of my chunk: 1/(a)_{n0+C}
                        for (int i = C - 1; i >= 0; --i) {
                            const int n = n0 + i;
                            if (n <= deg) gtab[TAB_CT + n] = cc;  // entries above the group's own degree stay zero
                            cc *= (a_top + (double)n);
                        }
                        for (int i = lane; i < (Mp - 1) + kCfMaxDepth; i += LANES) {
                            if (i < Mp - 1) gtab[TAB_IA + i] = 1.0 / (k + (double)i);
                            else {
                                const int nn = i - (Mp - 1) + 1;
                                const double fn = (double)nn;
                                gtab[TAB_CF + nn] = (nn <= cfd) ? fn * (fn - a_top) : 0.0;
                            }
                        }
                    }
                    __syncwarp();
                    double gam_top = par[PAR_GK];
                    for (int p = 0; p < Mp - 1; ++p) gam_top *= (k + (double)p);  // Γ(k+Mp-1)
                    double acc[T];
                    node_integrals<(MPMAX > 0 ? MPMAX : 1), LANES, NodesPerLane<LANES>::value>(acc, sTab + cfg.tab_off[md], cfg.n_bins[md], M, Mp, k, inv_th, log_th,
                                                                  gam_top, gtab, deg_w, cfd_w, cfd, ser_lim, lane);
                    // F = 0 | min(Mom*Mom, H) — Coalescence.jl:212-227; H = n²θ^{p2}/Γ(k)² * Σ
                    const double* mom = my + L.mom + md * M;
                    const double pre0 = n_md * n_md * par[PAR_IGK2];
                    {
                        int t = 0;
#pragma unroll
                        for (int p1 = 0; p1 < MPMAX; ++p1) {
                            double pre = pre0;
                            for (int q = 0; q < p1; ++q) pre *= th;
#pragma unroll
                            for (int p2 = p1; p2 < MPMAX; ++p2) {
                                if (p1 < Mp && p2 < Mp) {
                                    const double mm = mom[p1] * mom[p2];
                                    const double H = pre * acc[t];
                                    const double v = (mm < kEps) ? 0.0 : jl_min(mm, H);
                                    const int idx = tri_index(p1, p2, Mp);
                                    if ((idx % LANES) == lane) Fk[idx] = skip ? 0.0 : v;
                                }
                                pre *= th;
                                ++t;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---- contraction: one output moment per lane — Coalescence.jl:140-149, :260-455 ----
            for (int o = lane; o < nslots; o += LANES) {
                const int k = cfg.slot_mode[o], m = cfg.slot_order[o];
                double result = 0.0;
                if (!(RAIN && cell_empty)) {
                    const double* momk = my + L.mom + k * M;
                    double bin[3];
                    bin[0] = 1.0; bin[1] = (m == 2) ? 2.0 : 1.0; bin[2] = 1.0;  // C(m, c)
                    double sumQ = 0.0, sumR = 0.0;
                    for (int j = 0; j < N; ++j) {
                        const double* momj = my + L.mom + j * M;
                        const double* cjk = sC + ((j * N + k) * P) * P;
                        if (k > j) {  // Q_jk — :283-309
                            double q = 0.0;
                            for (int a = 0; a < P; ++a) {
                                double qa = 0.0;
                                for (int b = 0; b < P; ++b) {
                                    double qb = 0.0;
                                    for (int c = 0; c <= m; ++c) qb += cjk[a * P + b] * bin[c] * momj[a + c] * momk[b + m - c];
                                    qa += qb;
                                }
                                q += qa;
                            }
                            sumQ += q;
                        }
                        {  // R_jk — :334-351
                            double r = 0.0;
                            for (int a = 0; a < P; ++a) {
                                double ra = 0.0;
                                for (int b = 0; b < P; ++b) ra += cjk[a * P + b] * momj[a] * momk[b + m];
                                r += ra;
                            }
                            sumR += r;
                        }
                    }
                    // S_1k — :398-424
                    double s1 = 0.0;
                    {
                        const double* ckk = sC + ((k * N + k) * P) * P;
                        const double* park = my + L.par + k * PAR_STRIDE;
                        const double* Fk = my + L.F + (k < N - 1 ? k : 0) * MAXT;
                        for (int a = 0; a < P; ++a) {
                            double sa = 0.0;
                            for (int b = 0; b < P; ++b) {
                                double sb = 0.0;
                                for (int c = 0; c <= m; ++c)
                                    sb += 0.5 * ckk[a * P + b] * bin[c] * F_entry(cfg, momk, park, Fk, k, a + c, b + m - c);
                                sa += sb;
                            }
                            s1 += sa;
                        }
                    }
                    // S_2,k-1 — :426-455 (zero when both modes carry too few moments, :366-369)
                    double s2 = 0.0;
                    if (k > 0) {
                        const int kk = k - 1;
                        const double* momp = my + L.mom + kk * M;
                        const double* cpp = sC + ((kk * N + kk) * P) * P;
                        const double* parp = my + L.par + kk * PAR_STRIDE;
                        const double* Fp = my + L.F + kk * MAXT;
                        for (int a = 0; a < P; ++a) {
                            double sa = 0.0;
                            for (int b = 0; b < P; ++b) {
                                double sb = 0.0;
                                for (int c = 0; c <= m; ++c)
                                    sb += 0.5 * cpp[a * P + b] * bin[c] *
                                          (momp[a + c] * momp[b + m - c] - F_entry(cfg, momp, parp, Fp, kk, a + c, b + m - c));
                                sa += sb;
                            }
                            s2 += sa;
                        }
                    }
                    result = sumQ - sumR + s1;
                    if (k > 0) result += s2;
                    if (!args.params_in) result *= cfg.norm[o];
                }
                my[L.out + o] = result;
            }
        }
        __syncthreads();
        // ---- store phase ----
        for (int i = tid; i < nslots * G; i += THREADS) {
            const int s = i / G, g = i % G;
            const long long p = p0 + g;
            if (p >= args.n) continue;
            const double* rec = sGroups + (size_t)g * L.total;
            double f;
            if (RAIN) {
                const double fl = rec[L.flux + s];
                if (args.flux_only) {
                    f = fl;
                } else {
                    double fl_up;  // flux of the level above; zero at the column top (rainshaft_helpers.jl:80-81)
                    if ((p + 1) % cfg.nz == 0) fl_up = 0.0;
                    else if (g + 1 < G) fl_up = (sGroups + (size_t)(g + 1) * L.total)[L.flux + s];
                    else fl_up = sHaloFlux[s];
                    f = rec[L.out + s] + (-(fl_up - fl) / cfg.dz);
                }
            } else {
                f = rec[L.out + s];
            }
            double o;
            if (args.tend_only) {
                o = f;
            } else {
                const double ui = rec[L.in + s];
                double acc2 = args.ci * ui;
                if (args.u_n != nullptr) acc2 = args.cn * rec[L.un + s] + acc2;
                o = (acc2 + args.cf * (args.dt * f)) / args.div;
                if (RAIN) o = (o < 0.0) ? 0.0 : o;  // clipped by the next RHS evaluation (rainshaft_helpers.jl:52)
            }
            args.out[s * args.s_out + p * args.ps_out] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// AoS [n][nslots] (host order) <-> SoA [nslots][stride]
__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, long long n, int nslots, long long stride) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = n * nslots;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / nslots;
        const int s = (int)(i % nslots);
        soa[s * stride + p] = aos[i];
    }
}
__global__ void soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, long long n, int nslots, long long stride) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = n * nslots;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / nslots;
        const int s = (int)(i % nslots);
        aos[i] = soa[s * stride + p];
    }
}

// per-slot sums, deterministic two-pass: block partials (fixed grid) then one block finishes
constexpr int SUM_BLOCKS = 592;  // 148 SMs x 4
constexpr int SUM_THREADS = 256;
__global__ void __launch_bounds__(SUM_THREADS) moment_partial_kernel(const double* __restrict__ u, long long n, long long stride, int nslots,
                                                                   double* __restrict__ partial) {
    __shared__ double red[SUM_THREADS / 32];
    for (int s = 0; s < nslots; ++s) {
        double acc = 0.0;
        const double* col = u + s * stride;
        for (long long p = blockIdx.x * (long long)SUM_THREADS + threadIdx.x; p < n; p += (long long)SUM_BLOCKS * SUM_THREADS) acc += col[p];
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < SUM_THREADS / 32; ++w) t += red[w];
            partial[s * SUM_BLOCKS + blockIdx.x] = t;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SUM_THREADS) moment_final_kernel(const double* __restrict__ partial, int nslots, double* __restrict__ out) {
    __shared__ double red[SUM_THREADS / 32];
    for (int s = 0; s < nslots; ++s) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < SUM_BLOCKS; i += SUM_THREADS) acc += partial[s * SUM_BLOCKS + i];
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < SUM_THREADS / 32; ++w) t += red[w];
            out[s] = t;
        }
        __syncthreads();
    }
}

// single-object evaluations (arbitrary real orders): one warp
struct ScalarArgs {
    int op;  // 0 moment, 1 update_dist, 2 moment_source_helper, 3 sed flux, 4 simpson, 5 compute_threshold
    int kind;
    double params[3];
    double q, p1, p2, x_th;
    int n_bins;
    double x_min, dx;
    double range[4];
    int n_modes, n_vel;
    int kinds[MAXN];
    double mparams[MAXN][3];
    double vel[CLOUDY_MAX_VEL][2];
    const double* y;  // simpson table (device)
    double* out;      // device
    int* iout;
};

__global__ void scalar_kernel(const ScalarArgs a) {
    const int lane = threadIdx.x;
    if (a.op == 0) {
        if (lane == 0) a.out[0] = moment_real(a.kind, a.params[0], a.params[1], a.params[2], a.q);
    } else if (a.op == 1) {
        if (lane == 0) {
            ModeParams r = params_from_moments(a.kind, a.params[0], a.params[1], a.params[2], a.range[0], a.range[1], a.range[2], a.range[3]);
            a.out[0] = r.n; a.out[1] = r.a; a.out[2] = r.b;
            a.iout[0] = r.invalid;
        }
    } else if (a.op == 2) {
        // moment_source_helper for Exponential / Gamma with real p1, p2 (ParticleDistributions.jl:567-612)
        const double n = a.params[0], th = a.params[1];
        const double k = (a.kind == CLOUDY_GAMMA) ? a.params[2] : 1.0;
        double part = 0.0;
        for (int j = 1 + lane; j <= a.n_bins; j += 32) {
            const double x = exp(a.x_min + (j - 1) * a.dx);
            const double f = pow(x, a.p1 + k - 1.0) * exp(-x / th) * igam_lower(a.p2 + k, (a.x_th - x) / th);
            part += simpson_weight(j, a.n_bins) * (x * f);
        }
        for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
        if (lane == 0) {
            const double gk = tgamma(k);
            a.out[0] = n * n * pow(th, a.p2 - k) / (gk * gk) * (a.dx * (part / 48.0));
        }
    } else if (a.op == 3) {
        if (lane == 0) {
            int o = 0;
            for (int i = 0; i < a.n_modes; ++i) {
                const int np = (a.kinds[i] == CLOUDY_GAMMA || a.kinds[i] == CLOUDY_LOGNORMAL) ? 3 : 2;
                for (int j = 0; j < np; ++j) {
                    double s = 0.0;
                    for (int v = 0; v < a.n_vel; ++v)
                        s += -a.vel[v][0] * moment_real(a.kinds[i], a.mparams[i][0], a.mparams[i][1], a.mparams[i][2], (double)j + a.vel[v][1]);
                    a.out[o++] = s;
                }
            }
        }
    } else if (a.op == 5) {
        // compute_threshold(pdist, percentile, minx) — ParticleDistributions.jl:747-761
        if (lane == 0) {
            const double th = a.params[1];
            const double x = (a.kind == CLOUDY_GAMMA) ? th * igam_inv(a.params[2], a.q) : -th * log(1.0 - a.q);
            a.out[0] = fmax(x, a.p1);
        }
    } else if (a.op == 6) {
        // moment_source_helper for a Lognormal mode with real p1, p2 (ParticleDistributions.jl:614-625): inner integral in closed
        // form, outer integral by the same fixed Gauss-Legendre rule in t = ln y as the batched path (tpp_lognormal_H);
        // a.y = rule nodes then weights, a.n_bins = rule size
        const double n = a.params[0], mu = a.params[1], sg = a.params[2], s2 = sg * sg;
        const double hi = fmin(log(a.x_th), mu + fmax(fmax(a.p1, a.p2), 0.0) * s2 + 12.0 * sg);
        const double lo = mu - 12.0 * sg;
        double part = 0.0;
        if (hi > lo) {
            const double half = 0.5 * (hi - lo), mid = 0.5 * (hi + lo), inv_sg = 1.0 / sg;
            const double pref = n * inv_sg * 0.3989422804014327;  // n / (σ sqrt(2π))
            const double c1 = n * exp(a.p1 * mu + a.p1 * a.p1 * s2 / 2);
            for (int q = lane; q < a.n_bins; q += 32) {
                const double t = fma(half, a.y[q], mid);
                const double y = exp(t);
                const double dd = (t - mu) * inv_sg;
                const double rem = a.x_th - y;
                const double inner = (rem > 0.0) ? c1 * norm_cdf((log(fmax(rem, 1e-300)) - mu - a.p1 * s2) * inv_sg) : 0.0;
                part += a.y[a.n_bins + q] * half * pref * exp(fma(a.p2, t, -0.5 * dd * dd)) * inner;  // weight * y^p2 f(y) y * inner
            }
        }
        for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
        if (lane == 0) a.out[0] = part;
    } else if (a.op == 4) {
        double part = 0.0;
        for (int j = 1 + lane; j <= a.n_bins + 1; j += 32) part += simpson_weight(j, a.n_bins) * a.y[j - 1];
        for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
        if (lane == 0) a.out[0] = a.dx * (part / 48.0);
    }
}

// sedimentation flux of every cell, thread per cell — Sedimentation.jl:22-37 with the velocity normalisation of
// rainshaft_helpers.jl:74-77.  moment(dist, q+β) for q = 0,1,2 follows from the q = 0 value by Γ(x+1) = xΓ(x).
// One instance per mode count (every loop over modes unrolls with no `i < cfg.N` test, half the registers of a MAXN body:
// 29.7 M -> see profiles/ for the instruction count per Mi cells); a warp whose cells are all empty writes zeros at once.
template <int NM>
__global__ void __launch_bounds__(256, 4) flux_kernel(const __grid_constant__ DevConfig cfg, const KArgs args) {
    const double* __restrict__ uin = args.u_in;
    double* __restrict__ out = args.out;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < args.n; p += (long long)gridDim.x * blockDim.x) {
        // every moment of the cell is requested before any arithmetic: the kernel is load-latency bound otherwise
        double raw[NM][3];
#pragma unroll
        for (int i = 0; i < NM; ++i)
#pragma unroll
            for (int q = 0; q < 3; ++q)
                raw[i][q] = (q < cfg.nprog[i]) ? __ldg(uin + (cfg.slot0[i] + q) * args.s_in + p * args.ps_in) : 0.0;
        double pn[NM], pa[NM], pb[NM];
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            pn[i] = 0.0; pa[i] = 1.0; pb[i] = 1.0;
            // a mode without number or mass (most cells of a column: exact zeros, or clipped negatives) is the empty-mode fallback of
            // update_dist_from_moments (n = 0) whatever the normalisation: no divisions, no flux
            if (!(raw[i][0] > 0.0) || !(raw[i][1] > 0.0)) continue;
            const int s0 = cfg.slot0[i], np = cfg.nprog[i], kind = cfg.kind[i];
            double mn[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (q >= np) break;
                double v = raw[i][q];
                v = (v < 0.0) ? 0.0 : v;  // rainshaft_helpers.jl:52
                mn[q] = norm_div(v, cfg.norm[s0 + q], cfg.inv_norm[s0 + q]);
            }
            const ModeParams mp = params_from_moments(kind, mn[0], mn[1], mn[2], kind == CLOUDY_GAMMA ? cfg.k_lo : -INFINITY,
                                                      kind == CLOUDY_GAMMA ? cfg.k_hi : INFINITY);
            pn[i] = mp.n; pa[i] = mp.a; pb[i] = mp.b;
        }
        double fl[NM][3];
        cell_flux<NM>(cfg, pn, pa, pb, fl);
#pragma unroll
        for (int i = 0; i < NM; ++i) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < cfg.nprog[i]) out[(cfg.slot0[i] + q) * args.s_out + p * args.ps_out] = fl[i][q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// "next row" kernels: condensation tendency and the cloud/rain split diagnostic (thread per parcel)
// ------------------------------------------------------------------------------------------------
struct AuxArgs {
    const double* u_in;
    long long s_in, n;
    double* out;
    long long s_out;
    int params_in;     // u_in holds (n, θ|μ[, k|σ]) of ONE set of distributions instead of moments
    int normalized;    // N_q: rebuild distributions from moments / norms (1) or from the raw moments (0)
    double s, xi_n, rho_l, cutoff;
    const double* d_s;  // per-parcel supersaturation (optional)
    const int* order;   // regime-ordered ensemble: original parcel index of every position (nullptr = identity)
};

__device__ inline ModeParams aux_params(const DevConfig& cfg, const AuxArgs& a, int i, long long p, bool normalise) {
    const int s0 = cfg.slot0[i], np = cfg.nprog[i], kind = cfg.kind[i];
    double m[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < np; ++q) m[q] = a.u_in[(s0 + q) * a.s_in + p];
    ModeParams mp;
    if (a.params_in) {
        mp.n = m[0]; mp.a = m[1]; mp.b = (np > 2) ? m[2] : 1.0; mp.invalid = 0;
        return mp;
    }
    if (normalise)
        for (int q = 0; q < np; ++q) m[q] /= cfg.norm[s0 + q];
    return params_from_moments(kind, m[0], m[1], m[2], kind == CLOUDY_GAMMA ? cfg.k_lo : -INFINITY, kind == CLOUDY_GAMMA ? cfg.k_hi : INFINITY);
}

// get_cond_evap — src/Sources/Condensation.jl:22-37 (normalisation of rhs_condensation!, box_model_helpers.jl:55-67)
__global__ void __launch_bounds__(256) cond_evap_kernel(const __grid_constant__ DevConfig cfg, const AuxArgs a) {
    const double geom = pow(4.0 * M_PI / 3.0, 2.0 / 3.0) / pow(a.rho_l, 1.0 / 3.0);
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < a.n; p += (long long)gridDim.x * blockDim.x) {
        const double s = a.d_s ? a.d_s[a.order ? (long long)a.order[p] : p] : a.s;  // d_s is indexed by parcel
        for (int i = 0; i < cfg.N; ++i) {
            const int s0 = cfg.slot0[i], np = cfg.nprog[i], kind = cfg.kind[i];
            const ModeParams mp = aux_params(cfg, a, i, p, true);
            for (int j = 1; j <= np; ++j) {
                double v = 0.0;
                if (j >= 2) v = 3 * a.xi_n * s * (j - 1) * moment_real(kind, mp.n, mp.a, mp.b, (double)(j - 1) - 2.0 / 3.0) * geom;
                a.out[(s0 + j - 1) * a.s_out + p] = a.params_in ? v : v * cfg.norm[s0 + j - 1];
            }
        }
    }
}

// partial_moment(dist, q, x) — ParticleDistributions.jl:226-285 (Lognormal: the closed form of the reference's quadgk)
__device__ inline double partial_moment_dev(int kind, double n, double a, double b, double q, double x) {
    switch (kind) {
        case CLOUDY_EXPONENTIAL: return n * pow(a, q) * igam_lower(q + 1.0, x / a);
        case CLOUDY_GAMMA: return n * pow(a, q) * igam_lower(q + b, x / a) / tgamma(b);
        case CLOUDY_MONODISPERSE: return (x < a) ? 0.0 : n * pow(a, q);
        default: return n * exp(q * a + q * q * b * b / 2) * norm_cdf((log(x) - a - q * b * b) / b);
    }
}

// get_standard_N_q — ParticleDistributions.jl:634-687; out[4][n] = N_liq, N_rai, M_liq, M_rai
__global__ void __launch_bounds__(256) nq_kernel(const __grid_constant__ DevConfig cfg, const AuxArgs a) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < a.n; p += (long long)gridDim.x * blockDim.x) {
        double nl = 0.0, nr = 0.0, ml = 0.0, mr = 0.0;
        for (int i = 0; i < cfg.N; ++i) {
            const int kind = cfg.kind[i];
            const ModeParams mp = aux_params(cfg, a, i, p, a.normalized != 0);
            const double p0 = partial_moment_dev(kind, mp.n, mp.a, mp.b, 0.0, a.cutoff);
            const double p1 = partial_moment_dev(kind, mp.n, mp.a, mp.b, 1.0, a.cutoff);
            nl += p0;
            ml += p1;
            nr += moment_real(kind, mp.n, mp.a, mp.b, 0.0) - p0;
            mr += moment_real(kind, mp.n, mp.a, mp.b, 1.0) - p1;
        }
        const long long o = a.order ? (long long)a.order[p] : p;
        a.out[0 * a.s_out + o] = nl;
        a.out[1 * a.s_out + o] = nr;
        a.out[2 * a.s_out + o] = ml;
        a.out[3 * a.s_out + o] = mr;
    }
}

// ------------------------------------------------------------------------------------------------
// regime sort (optional): order the parcels so that the 32 parcels of a warp need a similar series length and
// the same series / continued-fraction regime.  Purely a performance hint — a parcel's result does not depend on
// its position — so the order may be reused across the stages of a time step.
// ------------------------------------------------------------------------------------------------
// The key is only a scheduling hint, so the distribution parameters are rebuilt in single precision (the FP32 pipe is idle
// in this library; FP64 divisions would make this kernel cost 6 % of a C2 step) and the block histogram is warp-aggregated.
constexpr int KEY_PER_THREAD = 4;
// The sort can be done tile by tile (SORT_TILE consecutive parcels, a multiple of 256 * KEY_PER_THREAD so that a block never
// straddles two tiles): the thread-per-parcel kernel walks the order front to back, so the parcels in flight then come from
// one or two tiles and the gather stays local (a gather over a whole 16 Mi-parcel C4 ensemble moves 26 GB through DRAM for
// 2.4 GB of state: a 128-byte line per 8-byte moment).  Measured on B200 (tools/bench_configs.py, CLOUDY_SORT_TILE):
//   C4 (9 slots, 16 Mi parcels): whole ensemble 55.0 ms, 4 Mi tiles 47.5, 2 Mi 45.8, 1 Mi 50.9, 256 Ki 83.9, 64 Ki 103
//   C5 (5 slots, 64 Mi parcels): whole ensemble 31.9 ms (fused step 93.6), 4 Mi 32.2 (97.0), 2 Mi 32.4 (99.0), 1 Mi 32.7 (100.6)
// Small tiles lose more to mixed warps at the many bucket boundaries than they gain; wide states gain from 2 Mi tiles,
// narrow ones do not, so the default is 2 Mi parcels for states of 8 or more slots and the whole ensemble otherwise.
constexpr long long SORT_TILE_WIDE = 2097152;
// blockhist != nullptr selects the stable data sort of regime_sort.cuh: the block adds its histogram to
// blockhist[bin][blockIdx.x / SORT_ROUNDS] and to totals[bin] (integer atomics on distinct counters: the result does not
// depend on the order of arrival) instead of the per-tile histogram of the permutation sort.
__global__ void __launch_bounds__(256) regime_key_kernel(const __grid_constant__ DevConfig cfg, const KArgs args, unsigned char* __restrict__ keys,
                                                         unsigned int* __restrict__ hist, const long long SORT_TILE,
                                                         unsigned int* __restrict__ blockhist, unsigned int* __restrict__ totals, const int nblocks) {
    __shared__ unsigned int sh[256];
    __shared__ unsigned char sdeg[kSerZ][kSerA];  // per-thread (divergent) lookups: shared memory, not the constant bank
    __shared__ float slim[kSerA];
    // the (at most two) thresholded modes that enter the key
    int qm[2] = {-1, -1}, n_used = 0;
    for (int i = 0; i < cfg.N - 1 && n_used < 2; ++i)
        if (cfg.quad[i]) qm[n_used++] = i;
    // all loads of the thread's KEY_PER_THREAD parcels are issued before any arithmetic (the kernel is DRAM-latency bound)
    double raw[KEY_PER_THREAD][2][3];
#pragma unroll
    for (int r = 0; r < KEY_PER_THREAD; ++r) {
        const long long p = ((long long)blockIdx.x * KEY_PER_THREAD + r) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                raw[r][u][q] = 0.0;
                if (p < args.n && u < n_used && q < cfg.nprog[qm[u]])
                    raw[r][u][q] = __ldg(args.u_in + (cfg.slot0[qm[u]] + q) * args.s_in + p * args.ps_in);
            }
    }
    sh[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < kSerZ * kSerA; i += blockDim.x) sdeg[i / kSerA][i % kSerA] = kSeriesDeg2[i / kSerA][i % kSerA];
    if (threadIdx.x < kSerA) slim[threadIdx.x] = (float)kSeriesLimit[threadIdx.x];
    __syncthreads();
    const float k_lo = (float)cfg.k_lo, k_hi = (float)cfg.k_hi;
#pragma unroll
    for (int r = 0; r < KEY_PER_THREAD; ++r) {
        const long long p = ((long long)blockIdx.x * KEY_PER_THREAD + r) * blockDim.x + threadIdx.x;
        const bool live = p < args.n;
        unsigned int key = 0;
        if (live) {
#pragma unroll
            for (int used = 0; used < 2; ++used) {
                if (used >= n_used) break;
                const int i = qm[used];
                const int s0 = cfg.slot0[i], kind = cfg.kind[i];
                float m[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) m[q] = (float)raw[r][used][q] / (float)cfg.norm[s0 + q];
                unsigned int sub = 0;
                if (m[0] > 2.220446e-16f && m[1] > 2.220446e-16f) {  // update_dist_from_moments' non-empty test
                    const float mean = m[1] / m[0];
                    float kk = 1.f;
                    if (kind == CLOUDY_GAMMA) {
                        kk = mean / (m[2] / m[1] - mean);
                        kk = fmaxf(k_lo, fminf(k_hi, kk));
                        if (!(kk == kk)) kk = k_hi;
                    }
                    const float theta = mean / kk;
                    const float a_top = kk + (float)(cfg.M - 1);
                    const int ai = (int)fminf(fmaxf(a_top, 0.f), (float)(kSerA - 1));
                    const float ser_lim = slim[ai];
                    float X = (float)cfg.thr[i] / theta;
                    bool flag = X >= ser_lim;  // FixedThreshold: series / continued-fraction regime
                    if (cfg.thr_style == CLOUDY_MOVING_THRESHOLD) {
                        // x_th/θ = x_p(k) from the table's first guess (no polish needed for a sort key); the flag separates the
                        // parcels that scale the unit grid (x_th <= 1) from those with a grid of their own
                        const double Xq = (kind == CLOUDY_GAMMA)
                                              ? igam_inv_guess((double)kk, cfg.tab + cfg.xp_off[i], cfg.xp_n, cfg.xp_k0, cfg.xp_inv_h)
                                              : -log(1.0 - cfg.thr[i]);
                        X = (Xq > 0.0) ? (float)Xq : 1.f;
                        flag = theta * X > 1.f;
                    }
                    const float zc = fminf(X, ser_lim - 0.5f);
                    const int zi = (zc >= 0.f) ? (int)fminf(zc, (float)(kSerZ - 1)) : 0;
                    const unsigned int deg = sdeg[zi][ai];
                    sub = (flag ? 1u : 0u) | (((deg >> 3) & 7u) << 1);
                    // FixedThreshold parcels in the continued-fraction regime all have the top series degree: their three bits
                    // carry how far x_th/θ lies beyond the series limit instead (buckets of 8), which fixes the depth of the
                    // continued fraction node by node (kCfDepthZ) — warps then agree on the depth as well
                    if (flag && cfg.thr_style != CLOUDY_MOVING_THRESHOLD) sub = 1u | ((unsigned int)fminf((X - ser_lim) * 0.125f, 7.f) << 1);
                    sub = sub == 0 ? 2u : sub;  // keep 0 for "empty mode"
                }
                key |= sub << (4 * used);
            }
            keys[p] = (unsigned char)key;
        }
        // warp-aggregated histogram update: one shared-memory atomic per distinct key in the warp
        const unsigned int act = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned int peers = __match_any_sync(act, key);
            if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sh[key], (unsigned int)__popc(peers));
        }
    }
    __syncthreads();
    if (blockhist != nullptr) {
        if (sh[threadIdx.x]) {
            atomicAdd(&blockhist[(size_t)threadIdx.x * nblocks + blockIdx.x / SORT_ROUNDS], sh[threadIdx.x]);
            atomicAdd(&totals[threadIdx.x], sh[threadIdx.x]);
        }
        return;
    }
    const long long tile = ((long long)blockIdx.x * KEY_PER_THREAD * blockDim.x) / SORT_TILE;
    if (sh[threadIdx.x]) atomicAdd(&hist[tile * 256 + threadIdx.x], sh[threadIdx.x]);
}

// Stable-per-block counting-sort scatter within the block's tile.  Every block rebuilds the exclusive prefix of its tile's
// 256-bin histogram itself (no separate scan launch) and reserves its slots per bin with one global atomic; `fill`
// starts at zero; tile t owns positions [t SORT_TILE, (t+1) SORT_TILE) of the order.
__global__ void __launch_bounds__(256) regime_scatter_kernel(const unsigned char* __restrict__ keys, const unsigned int* __restrict__ hist,
                                                             unsigned int* __restrict__ fill, int* __restrict__ perm, long long n,
                                                             const long long SORT_TILE) {
    __shared__ unsigned int cnt[256];
    __shared__ unsigned int base[256];
    __shared__ unsigned int pre[256];
    const unsigned int tid = threadIdx.x, lane = tid & 31;
    cnt[tid] = 0;
    unsigned int key[KEY_PER_THREAD], rank[KEY_PER_THREAD];
#pragma unroll
    for (int r = 0; r < KEY_PER_THREAD; ++r) {
        const long long p = ((long long)blockIdx.x * KEY_PER_THREAD + r) * blockDim.x + tid;
        key[r] = (p < n) ? keys[p] : 0u;
    }
    const long long tile = ((long long)blockIdx.x * KEY_PER_THREAD * blockDim.x) / SORT_TILE;
    hist += tile * 256;
    fill += tile * 256;
    // exclusive prefix of the tile's histogram (Hillis-Steele over the block)
    const unsigned int own = hist[tid];
    pre[tid] = own;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const unsigned int v = (tid >= (unsigned)off) ? pre[tid - off] : 0u;
        __syncthreads();
        pre[tid] += v;
        __syncthreads();
    }
    const unsigned int excl = pre[tid] - own;
    __syncthreads();
    pre[tid] = excl;
#pragma unroll
    for (int r = 0; r < KEY_PER_THREAD; ++r) {
        const long long p = ((long long)blockIdx.x * KEY_PER_THREAD + r) * blockDim.x + tid;
        const bool live = p < n;
        const unsigned int act = __ballot_sync(0xffffffffu, live);
        rank[r] = 0;
        if (live) {
            // warp-aggregated ranking: the lowest lane of each key group reserves the group's slots
            const unsigned int peers = __match_any_sync(act, key[r]);
            const int leader = __ffs(peers) - 1;
            unsigned int base_w = 0;
            if ((int)lane == leader) base_w = atomicAdd(&cnt[key[r]], (unsigned int)__popc(peers));
            base_w = __shfl_sync(peers, base_w, leader);
            rank[r] = base_w + (unsigned int)__popc(peers & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    if (cnt[tid]) base[tid] = pre[tid] + atomicAdd(&fill[tid], cnt[tid]);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < KEY_PER_THREAD; ++r) {
        const long long p = ((long long)blockIdx.x * KEY_PER_THREAD + r) * blockDim.x + tid;
        if (p < n) perm[tile * SORT_TILE + base[key[r]] + rank[r]] = (int)p;
    }
}

// ln x_p(k) on the uniform k grid of igam_inv_tab (once per configuration)
__global__ void xp_table_kernel(double p, double k0, double h, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = log(igam_inv(k0 + (double)i * h, p));
}

// FP64 peak: 16 independent FMA chains per thread, the loop unrolled 4x (64 DFMA per branch)
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = threadIdx.x * 1e-3 + j;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fma(x[j], a, b);
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) sum += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

}  // namespace cloudy

// ================================================================================================
// host side: C ABI
// ================================================================================================
using namespace cloudy;

// Gauss-Legendre rule on [-1, 1] (Newton iteration on the Legendre polynomial): nodes xs[n], weights ws[n]
static void gauss_legendre_rule(int n, double* xs, double* ws) {
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double x = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * x * p2 - j * p3) / (j + 1.0);
            }
            pp = n * (x * p1 - p2) / (x * x - 1.0);
            const double dxn = p1 / pp;
            x -= dxn;
            if (fabs(dxn) < 1e-16) break;
        }
        xs[i] = -x; xs[n - 1 - i] = x;
        ws[i] = ws[n - 1 - i] = 2.0 / ((1.0 - x * x) * pp * pp);
    }
}
constexpr int kLognormalGL = 128;  // points of the Lognormal moment_source_helper rule (batched and scalar paths)

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) return fail(CLOUDY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// order of a regime-sorted ensemble: d[i] = original parcel index of position i.  Shared between states that hold the same
// parcels in the same order (a tendency inherits the order of the state it was evaluated at); reference counted.
struct OrderBuf {
    int* d;
    long long n;
    int refs;
};

struct cloudy_state {
    cloudy_ctx* ctx;
    double* d;
    long long n, stride;
    int nslots;
    OrderBuf* order;   // nullptr: position == parcel index
    int sort_age;      // fused time steps since the order was built
};

struct cloudy_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    bool configured;
    cloudy_config cfg;
    DevConfig dev;
    double* d_tab;
    int lanes;
    int mpmax;  // kernel instance: 0, 4, 5, 7
    int64_t launches;
    int sm_count;
    unsigned long long* d_err;
    double* d_partial;
    double* d_scratch;     // small device scratch for scalar entry points
    cloudy_state* tmp[3];  // stepper / host-path work buffers
    cloudy_state* flux;    // rainshaft: per-cell sedimentation flux for the thread-per-parcel kernel
    unsigned long long* d_tile_ctr;  // thread-per-parcel kernel: draw counter of the dynamic tile schedule (tpp_kernel.cuh)
    unsigned long long tile_base;    // its value once every launch enqueued so far has finished
    int sort_mode;         // regime sort of the parcel order before the thread-per-parcel kernel (0 off, 1 on)
    unsigned char* d_keys;
    int* d_perm;
    unsigned int* d_hist;  // per sort tile: 256-bin histograms, then 256 fill counters per tile
    size_t hist_tiles;     // tiles d_hist has room for
    long long sort_cap;
    bool perm_valid;       // d_perm may be reused (set by the stepper for stages 2 and 3 of a step)
    bool perm_fresh;       // a sort ran since the stepper last cleared this flag
    long long perm_n;      // ensemble size d_perm was computed for
    // stable data sort (regime_sort.cuh)
    unsigned int* d_blockhist;       // [256][nblocks] per-block histograms, then 256 bin totals
    size_t blockhist_cap;            // in unsigned ints
    cloudy_state* sort_scratch;      // destination of a state's data sort (buffers are swapped afterwards)
    std::vector<OrderBuf*>* order_pool;  // released order buffers, reused by size
    int resort_interval;             // fused steps between two sorts of a resident ensemble
    int64_t sorts;                   // data sorts performed so far
    // conservation all-reduce (NCCL, bound at run time with dlopen: the library has no link-time NCCL dependency)
    void* nccl_comm;                 // ncclComm_t
    int comm_size, comm_rank;
    cudaStream_t s_comm;             // side stream: all-reduce + copy to the host overlap the next step
    cudaEvent_t ev_sums, ev_comm;
    double* d_sums;                  // [2][MAXSLOT]: local sums, all-reduced sums
    double* h_sums;                  // pinned
    bool sums_pending;
    int sums_n;
    double* d_stage_aos;   // staging for upload/download
    // host-buffer pipeline (cloudy_coal_tendency_host): chunked H2D / kernel / D2H on three streams
    cudaStream_t s_h2d, s_d2h;
    cudaEvent_t ev_in[3], ev_k[3], ev_out[3];
    double* d_pipe_in[3];
    double* d_pipe_out[3];
    long long pipe_chunk;
    bool pipe_ready;
    long long stage_cap;
};

namespace cloudy {
tpp_fn tpp_lookup_A(int, int, int);
tpp_fn tpp_lookup_B(int, int, int);
tpp_fn tpp_lookup_C(int, int, int);
tpp_fn tpp_lookup_D(int, int, int);
tpp_fn tpp_lookup_E(int, int, int);
tpp_fn tpp_lookup_F(int, int, int);
tpp_fn tpp_lookup_G(int, int, int);
tpp_fn tpp_lookup_H(int, int, int);
tpp_fn tpp_lookup(int N, int P, int model) {
    tpp_fn f = tpp_lookup_A(N, P, model);
    if (!f) f = tpp_lookup_B(N, P, model);
    if (!f) f = tpp_lookup_C(N, P, model);
    if (!f) f = tpp_lookup_D(N, P, model);
    if (!f) f = tpp_lookup_E(N, P, model);
    if (!f) f = tpp_lookup_F(N, P, model);
    if (!f) f = tpp_lookup_G(N, P, model);
    if (!f) f = tpp_lookup_H(N, P, model);
    return f;
}
}  // namespace cloudy

typedef void (*rhs_fn)(const DevConfig, const KArgs);
struct KernelEntry {
    rhs_fn fn;
    int threads, groups;
};

template <int MPMAX, int LANES, int MODEL>
static KernelEntry entry() {
    return KernelEntry{(rhs_fn)rhs_kernel<MPMAX, LANES, MODEL>, Shape<LANES>::kThreads, Shape<LANES>::kGroups};
}
template <int LANES, int MODEL>
static KernelEntry pick_mp(int mpmax) {
    switch (mpmax) {
        case 0: return entry<0, LANES, MODEL>();
        case 4: return entry<4, LANES, MODEL>();
        case 5: return entry<5, LANES, MODEL>();
        default: return entry<7, LANES, MODEL>();
    }
}
template <int MODEL>
static KernelEntry pick_lanes(int lanes, int mpmax) {
    switch (lanes) {
        case 4: return pick_mp<4, MODEL>(mpmax);
        case 16: return pick_mp<16, MODEL>(mpmax);
        case 32: return pick_mp<32, MODEL>(mpmax);
        default: return pick_mp<8, MODEL>(mpmax);
    }
}

extern "C" {
static int ensure_flux(cloudy_ctx* ctx, long long n);
}

static int launch_flux(cloudy_ctx* ctx, const KArgs& args) {
    // 8 blocks per SM (4 and 28 measured the same on C3: the kernel is bound by the latency of its loads)
    long long blocks = std::min<long long>((args.n + 255) / 256, (long long)ctx->sm_count * 8);
    void* params[2] = {(void*)&ctx->dev, (void*)&args};
    const void* fn = nullptr;
    switch (ctx->dev.N) {
        case 1: fn = (const void*)flux_kernel<1>; break;
        case 2: fn = (const void*)flux_kernel<2>; break;
        case 3: fn = (const void*)flux_kernel<3>; break;
        default: fn = (const void*)flux_kernel<MAXN>; break;
    }
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)std::max<long long>(blocks, 1)), dim3(256), params, 0, ctx->stream));
    ctx->launches++;
    return CLOUDY_OK;
}

static int launch_tpp(cloudy_ctx* ctx, tpp_fn fn, int model, KArgs args) {
    const DevConfig& d = ctx->dev;
    if (args.flux_only) return launch_flux(ctx, args);
    if (model == CLOUDY_MODEL_RAINSHAFT) {
        int rc = ensure_flux(ctx, args.n);
        if (rc) return rc;
        KArgs fa = args;
        fa.out = ctx->flux->d;
        fa.s_out = ctx->flux->stride;
        if ((rc = launch_flux(ctx, fa))) return rc;
        args.flux = ctx->flux->d;
        args.s_flux = ctx->flux->stride;
    }
    {
        // the kernel forms element offsets in 32 bits (tpp_kernel.cuh)
        const long long smax = std::max({args.s_in, args.s_out, args.s_n, args.s_flux, args.s_clip});
        const long long pmax = std::max<long long>({args.ps_in, args.ps_out, 1});
        const unsigned long long span = (unsigned long long)smax * (unsigned long long)d.nslots + (unsigned long long)args.n * (unsigned long long)pmax;
        if (span >= (1ULL << 32))
            return fail(CLOUDY_ERR_UNSUPPORTED, "ensemble buffers of 2^32 doubles or more per device are not supported: split the ensemble");
    }
    bool any_quad = false;
    for (int i = 0; i < d.N - 1; ++i) any_quad = any_quad || d.quad[i];
    const bool want_sort = ctx->sort_mode == 1 || (ctx->sort_mode == 2 && args.n >= 262144);  // auto: pays from ~2e5 parcels (measured)
    if (want_sort && any_quad && !args.params_in && !args.presorted && args.n >= 4096 && args.n < (1LL << 31)) {
        if (ctx->sort_cap < args.n) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->d_keys); cudaFree(ctx->d_perm);
            ctx->d_keys = nullptr; ctx->d_perm = nullptr; ctx->sort_cap = 0;
            CUDA_TRY(cudaMalloc(&ctx->d_keys, (size_t)args.n));
            CUDA_TRY(cudaMalloc(&ctx->d_perm, sizeof(int) * (size_t)args.n));
            ctx->sort_cap = args.n;
            ctx->perm_valid = false;
        }
        if (!ctx->perm_valid || ctx->perm_n != args.n) {
            long long SORT_TILE = (d.nslots >= 8) ? SORT_TILE_WIDE : (1LL << 40);
            if (const char* e = getenv("CLOUDY_SORT_TILE")) SORT_TILE = std::max<long long>(1, atoll(e) / (256 * KEY_PER_THREAD)) * (256 * KEY_PER_THREAD);
            const size_t n_tiles = (size_t)((args.n + SORT_TILE - 1) / SORT_TILE);
            if (ctx->hist_tiles < n_tiles) {  // 256 histogram bins + 256 fill counters per tile
                cudaStreamSynchronize(ctx->stream);
                cudaFree(ctx->d_hist);
                ctx->d_hist = nullptr;
                ctx->hist_tiles = 0;
                CUDA_TRY(cudaMalloc(&ctx->d_hist, sizeof(unsigned int) * 512 * n_tiles));
                ctx->hist_tiles = n_tiles;
            }
            CUDA_TRY(cudaMemsetAsync(ctx->d_hist, 0, sizeof(unsigned int) * 512 * n_tiles, ctx->stream));
            const unsigned key_blocks = (unsigned)((args.n + 256 * KEY_PER_THREAD - 1) / (256 * KEY_PER_THREAD));
            unsigned int* no_hist = nullptr;
            int no_blocks = 0;
            void* kp[8] = {(void*)&ctx->dev, (void*)&args, (void*)&ctx->d_keys, (void*)&ctx->d_hist, (void*)&SORT_TILE,
                           (void*)&no_hist, (void*)&no_hist, (void*)&no_blocks};
            CUDA_TRY(cudaLaunchKernel((const void*)regime_key_kernel, dim3(key_blocks), dim3(256), kp, 0, ctx->stream));
            unsigned int* fill = ctx->d_hist + 256 * n_tiles;  // zeroed by the memset above
            regime_scatter_kernel<<<key_blocks, 256, 0, ctx->stream>>>(ctx->d_keys, ctx->d_hist, fill, ctx->d_perm, args.n, SORT_TILE);
            CUDA_TRY(cudaGetLastError());
            ctx->launches += 2;
            ctx->perm_fresh = true;
            ctx->perm_n = args.n;
        }
        args.perm = ctx->d_perm;
    }
    const size_t staged = (d.thr_style == CLOUDY_MOVING_THRESHOLD) ? (size_t)d.tpp_total : (size_t)d.tpp2_total;
    size_t smem = sizeof(double) * ((staged + 1) / 2 * 2 + (size_t)tpp_ct_rows(d.thr_style == CLOUDY_MOVING_THRESHOLD ? MODEL_BOX_MOVING : MODEL_BOX) * TPP_THREADS);
    CUDA_TRY(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, TPP_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    long long n_blocks = (args.n + TPP_THREADS - 1) / TPP_THREADS;
    long long grid = std::min<long long>(n_blocks, (long long)ctx->sm_count * per_sm);
    if (grid < 1) grid = 1;
    args.tile_ctr = ctx->d_tile_ctr;
    args.tile_base = ctx->tile_base;
    void* params[2] = {(void*)&ctx->dev, (void*)&args};
    CUDA_TRY(cudaLaunchKernel((const void*)fn, dim3((unsigned)grid), dim3(TPP_THREADS), params, smem, ctx->stream));
    {
        // the launch is enqueued: it will advance the draw counter by one draw per tile plus one overdraw per drawer
        const int km = (model == CLOUDY_MODEL_RAINSHAFT) ? MODEL_RAINSHAFT : MODEL_BOX;  // the MovingThreshold instances draw like the box ones
        ctx->tile_base += (unsigned long long)((args.n + tpp_tile(km) - 1) / tpp_tile(km)) + (unsigned long long)grid * tpp_overdraw(km);
    }
    ctx->launches++;
    return CLOUDY_OK;
}

static int launch_rhs(cloudy_ctx* ctx, int model, const KArgs& args) {
    // lanes == 0: thread-per-parcel kernel when the (N, P) shape has an instance, else the lane-cooperative kernel (8 lanes);
    // lanes == 1: thread-per-parcel required; lanes in {4, 8, 16, 32}: lane-cooperative kernel
    const bool need_tpp = ctx->dev.thr_style == CLOUDY_MOVING_THRESHOLD || ctx->dev.ln_thr[0] || ctx->dev.ln_thr[1] || ctx->dev.ln_thr[2];
    if (ctx->lanes <= 1 || need_tpp) {
        const bool moving = ctx->dev.thr_style == CLOUDY_MOVING_THRESHOLD;
        // the reference's column RHS calls the FixedThreshold method only (rainshaft_helpers.jl:70)
        if (moving && model == CLOUDY_MODEL_RAINSHAFT)
            return fail(CLOUDY_ERR_UNSUPPORTED, "MethodError: the rainshaft right-hand side has no MovingThreshold method");
        tpp_fn fn = tpp_lookup(ctx->dev.N, ctx->dev.P, model == CLOUDY_MODEL_RAINSHAFT ? MODEL_RAINSHAFT : (moving ? MODEL_BOX_MOVING : MODEL_BOX));
        if (fn) return launch_tpp(ctx, fn, model, args);
        if (ctx->lanes == 1 || need_tpp) return fail(CLOUDY_ERR_UNSUPPORTED, "no thread-per-parcel kernel instance for this (n_modes, P)");
    }
    const int lanes = ctx->lanes <= 1 ? 8 : ctx->lanes;
    KernelEntry ke = (model == CLOUDY_MODEL_RAINSHAFT) ? pick_lanes<MODEL_RAINSHAFT>(lanes, ctx->mpmax)
                                                       : pick_lanes<MODEL_BOX>(lanes, ctx->mpmax);
    const DevConfig& d = ctx->dev;
    const bool rain = (model == CLOUDY_MODEL_RAINSHAFT);
    GroupLayout L = group_layout(d.N, d.M, d.nslots, rain);
    size_t smem = sizeof(double) * ((size_t)d.N * d.N * d.P * d.P + d.tab_total + (size_t)ke.groups * L.total + 2 * d.nslots + 2);
    if (smem > 227 * 1024) return fail(CLOUDY_ERR_UNSUPPORTED, "configuration needs more than 227 KB of shared memory per block");
    CUDA_TRY(cudaFuncSetAttribute((const void*)ke.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)ke.fn, ke.threads, smem));
    if (per_sm < 1) per_sm = 1;
    long long n_tiles = (args.n + ke.groups - 1) / ke.groups;
    long long grid = std::min<long long>(n_tiles, (long long)ctx->sm_count * per_sm);
    if (grid < 1) grid = 1;
    void* params[2] = {(void*)&ctx->dev, (void*)&args};
    CUDA_TRY(cudaLaunchKernel((const void*)ke.fn, dim3((unsigned)grid), dim3(ke.threads), params, smem, ctx->stream));
    ctx->launches++;
    return CLOUDY_OK;
}


// ------------------------------------------------------------------------------------------------
// regime order of resident ensembles (regime_sort.cuh)
// ------------------------------------------------------------------------------------------------
static OrderBuf* order_acquire(cloudy_ctx* ctx, long long n) {
    for (size_t i = 0; i < ctx->order_pool->size(); ++i)
        if ((*ctx->order_pool)[i]->n == n) {
            OrderBuf* b = (*ctx->order_pool)[i];
            ctx->order_pool->erase(ctx->order_pool->begin() + i);
            b->refs = 1;
            return b;
        }
    OrderBuf* b = new OrderBuf();
    b->n = n;
    b->refs = 1;
    b->d = nullptr;
    if (cudaMalloc(&b->d, sizeof(int) * (size_t)std::max<long long>(n, 1)) != cudaSuccess) {
        delete b;
        return nullptr;
    }
    return b;
}
static void order_release(cloudy_ctx* ctx, OrderBuf* b) {
    if (!b || --b->refs > 0) return;
    // stream order makes the reuse safe: every kernel that reads the buffer was enqueued before the next sort writes it
    if (ctx->order_pool->size() < 4) {
        ctx->order_pool->push_back(b);
    } else {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b->d);
        delete b;
    }
}
static void state_set_order(cloudy_state* st, OrderBuf* o) {
    if (st->order == o) return;
    if (o) o->refs++;
    order_release(st->ctx, st->order);
    st->order = o;
}

// does the configuration profit from regime order, and is the ensemble large enough to pay for the sort?
static bool sort_wanted(const cloudy_ctx* ctx, long long n) {
    const DevConfig& d = ctx->dev;
    bool any_quad = false;
    for (int i = 0; i < d.N - 1; ++i) any_quad = any_quad || d.quad[i];
    const bool want = ctx->sort_mode == 1 || (ctx->sort_mode == 2 && n >= 262144);  // auto: pays from ~2e5 parcels (measured)
    return want && any_quad && n >= 4096 && n < (1LL << 31);
}
// the thread-per-parcel kernel will run for this model (the lane-cooperative kernel gains nothing from the order)
static bool tpp_selected(const cloudy_ctx* ctx, int model) {
    const bool moving = ctx->dev.thr_style == CLOUDY_MOVING_THRESHOLD;
    const bool need_tpp = moving || ctx->dev.ln_thr[0] || ctx->dev.ln_thr[1] || ctx->dev.ln_thr[2];
    if (!(ctx->lanes <= 1 || need_tpp)) return false;
    return tpp_lookup(ctx->dev.N, ctx->dev.P, model == CLOUDY_MODEL_RAINSHAFT ? MODEL_RAINSHAFT : (moving ? MODEL_BOX_MOVING : MODEL_BOX)) != nullptr;
}

// keys -> per-block histograms -> prefix -> scatter: `in` (order `order_in`, nullptr = identity) is copied to `out` in regime
// order and `order_out` receives the composed order.  All on the context's stream, no host synchronisation.
static int sort_buffer(cloudy_ctx* ctx, const double* in, double* out, long long stride, int nslots, long long n, const int* order_in,
                       int* order_out) {
    const int nblocks = (int)((n + SORT_BLOCK_PARCELS - 1) / SORT_BLOCK_PARCELS);
    const size_t need = (size_t)256 * nblocks + 256;
    if (ctx->blockhist_cap < need) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_blockhist);
        ctx->d_blockhist = nullptr;
        ctx->blockhist_cap = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_blockhist, sizeof(unsigned int) * need));
        ctx->blockhist_cap = need;
    }
    if (ctx->sort_cap < n) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_keys); cudaFree(ctx->d_perm);
        ctx->d_keys = nullptr; ctx->d_perm = nullptr; ctx->sort_cap = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_keys, (size_t)n));
        CUDA_TRY(cudaMalloc(&ctx->d_perm, sizeof(int) * (size_t)n));
        ctx->sort_cap = n;
        ctx->perm_valid = false;
    }
    CUDA_TRY(cudaMemsetAsync(ctx->d_blockhist, 0, sizeof(unsigned int) * need, ctx->stream));
    unsigned int* totals = ctx->d_blockhist + (size_t)256 * nblocks;
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.u_in = in; a.s_in = stride; a.ps_in = 1; a.n = n;
    const unsigned key_blocks = (unsigned)((n + 256 * KEY_PER_THREAD - 1) / (256 * KEY_PER_THREAD));
    unsigned int* no_hist = nullptr;
    long long no_tile = 1LL << 40;
    void* kp[8] = {(void*)&ctx->dev, (void*)&a, (void*)&ctx->d_keys, (void*)&no_hist, (void*)&no_tile,
                   (void*)&ctx->d_blockhist, (void*)&totals, (void*)&nblocks};
    CUDA_TRY(cudaLaunchKernel((const void*)regime_key_kernel, dim3(key_blocks), dim3(256), kp, 0, ctx->stream));
    sort_scan_kernel<<<256, SORT_THREADS, 0, ctx->stream>>>(ctx->d_blockhist, totals, nblocks);
    CUDA_TRY(cudaGetLastError());
    sort_scatter_kernel<<<nblocks, SORT_THREADS, 0, ctx->stream>>>(ctx->d_keys, ctx->d_blockhist, totals, nblocks, in, out, stride, nslots, n,
                                                                    order_in, order_out);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 3;
    ctx->sorts++;
    return CLOUDY_OK;
}

// move a state's parcels into regime order (the device buffer is exchanged with the context's scratch ensemble)
static int regime_sort_state(cloudy_ctx* ctx, cloudy_state* st) {
    if (st->n == 0) return CLOUDY_OK;
    if (!ctx->sort_scratch || ctx->sort_scratch->n != st->n || ctx->sort_scratch->nslots != st->nslots) {
        if (ctx->sort_scratch) { cloudy_state_destroy(ctx->sort_scratch); ctx->sort_scratch = nullptr; }
        int rc = cloudy_state_create(ctx, st->n, &ctx->sort_scratch);
        if (rc) return rc;
    }
    OrderBuf* o = order_acquire(ctx, st->n);
    if (!o) return fail(CLOUDY_ERR_CUDA, "cudaMalloc(order) failed");
    int rc = sort_buffer(ctx, st->d, ctx->sort_scratch->d, st->stride, st->nslots, st->n, st->order ? st->order->d : nullptr, o->d);
    if (rc) { order_release(ctx, o); return rc; }
    std::swap(st->d, ctx->sort_scratch->d);
    order_release(ctx, st->order);
    st->order = o;
    st->sort_age = 0;
    return CLOUDY_OK;
}


// ------------------------------------------------------------------------------------------------
// NCCL binding (dlopen): ncclUniqueId is 128 opaque bytes passed BY VALUE to ncclCommInitRank; ncclFloat64 = 8,
// ncclSum = 0, ncclSuccess = 0 (nccl.h 2.x ABI)
// ------------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
    void* handle;
    int (*GetUniqueId)(NcclId*);
    int (*CommInitRank)(void**, int, NcclId, int);
    int (*CommDestroy)(void*);
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    const char* (*GetErrorString)(int);
    int (*GetVersion)(int*);
};
static NcclApi* nccl_api(std::string& err) {
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::string load_err;
    if (!tried) {
        tried = true;
        const char* names[3] = {getenv("CLOUDY_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (int i = 0; i < 3 && !api.handle; ++i)
            if (names[i] && names[i][0]) api.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) {
            load_err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
        } else {
            api.GetUniqueId = (int (*)(NcclId*))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
            api.GetVersion = (int (*)(int*))dlsym(api.handle, "ncclGetVersion");
            ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
            if (!ok) load_err = "libnccl is missing a required symbol";
        }
    }
    if (!ok) { err = load_err; return nullptr; }
    return &api;
}
#define NCCL_TRY(api, expr)                                                                              \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0) return fail(CLOUDY_ERR_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(_r)); \
    } while (0)

static int ensure_comm_buffers(cloudy_ctx* ctx) {
    if (ctx->d_sums) return CLOUDY_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_comm, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_sums, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&ctx->d_sums, sizeof(double) * 2 * MAXSLOT));
    CUDA_TRY(cudaMallocHost(&ctx->h_sums, sizeof(double) * MAXSLOT));
    return CLOUDY_OK;
}

extern "C" {

const char* cloudy_last_error(void) { return g_last_error.c_str(); }

int cloudy_ctx_create(int device, void* stream, cloudy_ctx** out) {
    if (!out) return fail(CLOUDY_ERR_ARG, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(CLOUDY_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
    if (device < 0 || device >= count) return fail(CLOUDY_ERR_ARG, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cloudy_ctx* c = new cloudy_ctx();
    memset((void*)c, 0, sizeof(*c));
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->lanes = 0;
    c->sort_mode = 2;
    c->resort_interval = 10;
    c->order_pool = new std::vector<OrderBuf*>();
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaMalloc(&c->d_err, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(c->d_err, 0, sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&c->d_partial, sizeof(double) * SUM_BLOCKS * MAXSLOT));
    CUDA_TRY(cudaMalloc(&c->d_scratch, sizeof(double) * (CLOUDY_MAX_NODES + 64)));
    CUDA_TRY(cudaMalloc(&c->d_tile_ctr, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(c->d_tile_ctr, 0, sizeof(unsigned long long)));
    c->tile_base = 0;
    *out = c;
    return CLOUDY_OK;
}

int cloudy_ctx_destroy(cloudy_ctx* ctx) {
    if (!ctx) return CLOUDY_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 3; ++i)
        if (ctx->tmp[i]) cloudy_state_destroy(ctx->tmp[i]);
    cloudy_comm_destroy(ctx);
    if (ctx->d_sums) {
        cudaStreamDestroy(ctx->s_comm);
        cudaEventDestroy(ctx->ev_sums);
        cudaEventDestroy(ctx->ev_comm);
        cudaFree(ctx->d_sums);
        cudaFreeHost(ctx->h_sums);
    }
    if (ctx->flux) cloudy_state_destroy(ctx->flux);
    if (ctx->sort_scratch) cloudy_state_destroy(ctx->sort_scratch);
    for (OrderBuf* b : *ctx->order_pool) { cudaFree(b->d); delete b; }
    delete ctx->order_pool;
    cudaFree(ctx->d_blockhist);
    cudaFree(ctx->d_keys);
    cudaFree(ctx->d_perm);
    cudaFree(ctx->d_hist);
    cudaFree(ctx->d_tab);
    cudaFree(ctx->d_err);
    cudaFree(ctx->d_partial);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_tile_ctr);
    cudaFree(ctx->d_stage_aos);
    if (ctx->pipe_ready) {
        for (int i = 0; i < 3; ++i) {
            cudaFree(ctx->d_pipe_in[i]);
            cudaFree(ctx->d_pipe_out[i]);
            cudaEventDestroy(ctx->ev_in[i]);
            cudaEventDestroy(ctx->ev_k[i]);
            cudaEventDestroy(ctx->ev_out[i]);
        }
        cudaStreamDestroy(ctx->s_h2d);
        cudaStreamDestroy(ctx->s_d2h);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return CLOUDY_OK;
}

int cloudy_set_lanes(cloudy_ctx* ctx, int lanes) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    if (lanes != 0 && lanes != 1 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
        return fail(CLOUDY_ERR_ARG, "lanes must be 0 (auto), 1 (thread per parcel), 4, 8, 16 or 32");
    ctx->lanes = lanes;
    return CLOUDY_OK;
}

int cloudy_set_regime_sort(cloudy_ctx* ctx, int on) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    if (on < 0 || on > 2) return fail(CLOUDY_ERR_ARG, "regime sort mode must be 0 (off), 1 (on) or 2 (auto)");
    ctx->sort_mode = on;
    ctx->perm_valid = false;
    return CLOUDY_OK;
}

int cloudy_launch_count(cloudy_ctx* ctx, int64_t* out) {
    if (!ctx || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    *out = ctx->launches;
    return CLOUDY_OK;
}

int cloudy_sync(cloudy_ctx* ctx) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CLOUDY_OK;
}

int cloudy_config_set(cloudy_ctx* ctx, const cloudy_config* cfg) {
    if (!ctx || !cfg) return fail(CLOUDY_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int N = cfg->n_modes, P = cfg->P;
    if (N < 1 || N > MAXN) return fail(CLOUDY_ERR_ARG, "n_modes must be 1..4");
    if (P < 1 || P > MAXP) return fail(CLOUDY_ERR_ARG, "P must be 1..5");
    if (!(cfg->norms[0] > 0) || !(cfg->norms[1] > 0)) return fail(CLOUDY_ERR_ARG, "norms must be positive!");
    const bool moving = cfg->threshold_style == CLOUDY_MOVING_THRESHOLD;
    if (cfg->threshold_style != CLOUDY_FIXED_THRESHOLD && !moving) return fail(CLOUDY_ERR_ARG, "unknown threshold style");
    if (moving && !tpp_lookup(N, P, MODEL_BOX))
        return fail(CLOUDY_ERR_UNSUPPORTED, "MovingThreshold needs a thread-per-parcel kernel instance for this (n_modes, P)");
    if (!(cfg->k_range[1] <= 11.0)) return fail(CLOUDY_ERR_UNSUPPORTED, "k_range upper bound above 11 is outside the incomplete-gamma tables");
    DevConfig d;
    memset((void*)&d, 0, sizeof(d));
    d.N = N; d.P = P; d.M = P + 2;
    int slot = 0;
    for (int i = 0; i < N; ++i) {
        const int kind = cfg->kind[i];
        if (kind < 0 || kind > 3) return fail(CLOUDY_ERR_ARG, "unknown distribution kind");
        const int np = (kind == CLOUDY_GAMMA || kind == CLOUDY_LOGNORMAL) ? 3 : 2;
        if (cfg->nprog[i] != np) return fail(CLOUDY_ERR_ARG, "nprog[i] must equal nparams of the distribution kind");
        d.kind[i] = kind; d.nprog[i] = np; d.slot0[i] = slot;
        for (int q = 0; q < np; ++q) {
            d.slot_mode[slot] = i; d.slot_order[slot] = q;
            d.norm[slot] = cfg->norms[0] * pow(cfg->norms[1], (double)q);  // helper_functions.jl:49
            d.inv_norm[slot] = 1.0 / d.norm[slot];
            ++slot;
        }
    }
    d.nslots = slot;
    d.thr_style = cfg->threshold_style;
    d.n_mom_max = cfg->n_mom_max;
    d.k_lo = cfg->k_range[0]; d.k_hi = cfg->k_range[1];
    for (int j = 0; j < N; ++j)
        for (int k = 0; k < N; ++k)
            for (int a = 0; a < P; ++a)
                for (int b = 0; b < P; ++b) {
                    if (cfg->c[j][k][a][b] != cfg->c[j][k][b][a]) return fail(CLOUDY_ERR_ARG, "array not symmetric.");
                    d.c[j][k][a][b] = cfg->c[j][k][a][b];
                }
    // S terms of the thread-per-parcel kernel: sw[k][m][t(u,v)] = sum over (a, b, c) with {a+c, b+m-c} = {u, v} of
    // 0.5 c^{kk}_ab binomial(m, c) (Coalescence.jl:353-455 with F symmetric) — one multiply-add per (m, u <= v) in the kernel
    memset(d.sw, 0, sizeof(d.sw));
    {
        const int M = P + 2;
        for (int k = 0; k < N; ++k)
            for (int m = 0; m < 3; ++m)
                for (int a = 0; a < P; ++a)
                    for (int b = 0; b < P; ++b)
                        for (int c = 0; c <= m; ++c) {
                            const int x = a + c, y = b + m - c;
                            const int u = std::min(x, y), v = std::max(x, y);
                            const double binom = (m == 2 && c == 1) ? 2.0 : 1.0;
                            d.sw[k][m][u * M - (u * (u - 1)) / 2 + (v - u)] += 0.5 * cfg->c[k][k][a][b] * binom;
                        }
    }
    // grid tables
    std::vector<double> tab, tab2;  // tab: SoA tables of the lane-cooperative kernel; tab2: records / rules of the thread-per-parcel kernel
    std::vector<double> tab3, kblk3;  // FixedThreshold thread-per-parcel kernels: aligned records, Taylor degree per node block
    std::vector<double> ztab;         // FixedThreshold thread-per-parcel kernels: Z-sum polynomials in k (global memory, not staged)
    const int S2 = tpp_rec2_stride(P);
    int mpmax = 0;
    bool any_ln = false;
    for (int i = 0; i < N; ++i) {
        d.n2d[i] = cfg->n_2d_ints[i];
        d.Mp[i] = std::min(d.M, d.n2d[i]);
        d.thr[i] = cfg->thresholds[i];
        bool finite_thr = (i < N - 1) && !std::isinf(cfg->thresholds[i]);
        if (moving) {
            // thresholds are mass percentiles; compute_threshold exists for Exponential and Gamma only
            // (ParticleDistributions.jl:747-761); percentile 1 gives an infinite threshold
            const double pct = cfg->thresholds[i];
            if (i < N - 1) {
                if (cfg->kind[i] != CLOUDY_GAMMA && cfg->kind[i] != CLOUDY_EXPONENTIAL)
                    return fail(CLOUDY_ERR_ARG, "MethodError: compute_threshold is defined for Exponential and Gamma distributions only");
                if (!(pct >= 0.0 && pct <= 1.0)) return fail(CLOUDY_ERR_ARG, "percentile must be in [0, 1]");
                finite_thr = pct < 1.0;
            }
        }
        if (finite_thr && cfg->kind[i] == CLOUDY_LOGNORMAL && !tpp_lookup(N, P, MODEL_BOX))
            return fail(CLOUDY_ERR_UNSUPPORTED, "Lognormal mode with a finite threshold needs a thread-per-parcel kernel instance");
        d.ln_thr[i] = finite_thr && cfg->kind[i] == CLOUDY_LOGNORMAL;
        d.mono_thr[i] = finite_thr && cfg->kind[i] == CLOUDY_MONODISPERSE;
        d.quad[i] = finite_thr && (cfg->kind[i] == CLOUDY_GAMMA || cfg->kind[i] == CLOUDY_EXPONENTIAL);
        if (d.ln_thr[i]) {
            if (!(cfg->thresholds[i] > 0)) return fail(CLOUDY_ERR_ARG, "thresholds must be positive");
            any_ln = true;
        }
        if (d.quad[i]) {
            // FixedThreshold: the host's grid for the run-constant threshold.  MovingThreshold: the unit grid (threshold 1,
            // x_lb = 1e-5, 5*bins_per_log_unit nodes) that every parcel with x_th <= 1 scales by its own threshold.
            const int bpl = cfg->bins_per_log_unit > 0 ? cfg->bins_per_log_unit : 15;
            const int nb = moving ? 5 * bpl : cfg->n_bins[i];
            const double T = moving ? 1.0 : cfg->thresholds[i];
            const double g_xmin = moving ? log(1e-5) : cfg->x_min[i];
            const double g_dx = moving ? (0.0 - log(1e-5)) / nb : cfg->dx[i];
            if (!(T > 0)) return fail(CLOUDY_ERR_ARG, "thresholds must be positive");
            if (nb < 3) return fail(CLOUDY_ERR_ARG, "n_bins must be at least 3");
            if (nb > CLOUDY_MAX_NODES) return fail(CLOUDY_ERR_UNSUPPORTED, "n_bins exceeds CLOUDY_MAX_NODES");
            if (d.Mp[i] < 2) return fail(CLOUDY_ERR_ARG, "N_2d_ints too small");
            d.n_bins[i] = nb;
            d.tab_off[i] = (int)tab.size();
            std::vector<double> xj(nb), ell(nb), tmx(nb), lz(nb), w(nb, 0.0);
            const int e = nb + 1;
            auto addw = [&](int j, double v) { if (j >= 1 && j <= nb) w[j - 1] += v; };
            for (int j = 5; j <= nb - 3; ++j) addw(j, 1.0);
            addw(1, 17.0 / 48); addw(e, 17.0 / 48);
            addw(2, 59.0 / 48); addw(e - 1, 59.0 / 48);
            addw(3, 43.0 / 48); addw(e - 2, 43.0 / 48);
            addw(4, 49.0 / 48); addw(e - 3, 49.0 / 48);
            for (int j = 1; j <= nb; ++j) {
                ell[j - 1] = g_xmin + (j - 1) * g_dx;  // logx, ParticleDistributions.jl:566
                xj[j - 1] = exp(ell[j - 1]);
                tmx[j - 1] = T - xj[j - 1];
                if (!(tmx[j - 1] > 0)) return fail(CLOUDY_ERR_ARG, "grid node at or beyond the threshold");
                lz[j - 1] = log(tmx[j - 1]);
            }
            tab.insert(tab.end(), xj.begin(), xj.end());
            tab.insert(tab.end(), ell.begin(), ell.end());
            tab.insert(tab.end(), tmx.begin(), tmx.end());
            tab.insert(tab.end(), lz.begin(), lz.end());
            for (int p = 0; p < d.M; ++p)
                for (int j = 0; j < nb; ++j) tab.push_back(w[j] * g_dx * pow(xj[j], (double)p));
            {   // packed node records of the thread-per-parcel kernel: [near | far], each padded to a multiple of tpp_npl(P)
                static const double rho_thr[] = {1e-5, 1e-4, 1e-3, 3e-3, 1e-2, 2e-2, 3e-2, 5e-2, 7e-2, 0.1};
                static const int k_of[] = {4, 5, 7, 9, 11, 14, 16, 19, 22, 25};
                const int R = 5 + d.M;
                auto push_node = [&](int j, double kdeg, bool dummy) {
                    tab2.push_back(tmx[j]);  // dummies reuse a real node's position (zero weights)
                    tab2.push_back(ell[j] + lz[j]);
                    tab2.push_back(ell[j]);
                    tab2.push_back(xj[j]);
                    tab2.push_back(kdeg);
                    for (int p = 0; p < d.M; ++p) tab2.push_back(dummy ? 0.0 : w[j] * g_dx * pow(xj[j], (double)p));
                };
                d.rec_off[i] = (int)tab2.size();
                int n_near = 0, n_far = 0;
                double kmax = 4.0;
                for (int j = 0; j < nb; ++j) {
                    const double rho = xj[j] / T;
                    if (rho > 0.1 * (1.0 + 1e-9)) continue;
                    const double rr = 1.15 * rho;  // |z/X_c - 1| <= 1.15 rho when the centre is capped at the series limit
                    int kk = TPP_TAYLOR_MAX;
                    for (int q = 0; q < 10; ++q) if (rr <= rho_thr[q]) { kk = k_of[q]; break; }
                    kmax = std::max(kmax, (double)kk);
                    push_node(j, (double)kk, false);
                    ++n_near;
                }
                while (n_near % tpp_npl(P)) { push_node(0, kmax, true); ++n_near; }
                for (int j = 0; j < nb; ++j) {
                    if (!(xj[j] / T > 0.1 * (1.0 + 1e-9))) continue;
                    push_node(j, 0.0, false);
                    ++n_far;
                }
                while (n_far % tpp_npl(P)) { push_node(nb - 1, 0.0, true); ++n_far; }
                d.rec_near[i] = n_near;
                d.rec_far[i] = n_far;
                (void)R;
                // FixedThreshold kernels: the same near nodes as (tmx, ln x + ln(x_th - x), w_0 .. w_P, padding) records of S2
                // doubles with the Taylor degree of every block of tpp_npl(P) near nodes (= degree of the block's last node),
                // then the far nodes padded to a multiple of TPP_NPLF with at least one zero-weight dummy at the very end (its
                // slot evaluates the series at the Taylor centre, tpp_nodes_fixed2)
                d.rec2_off[i] = (int)tab3.size();
                d.kblk2_off[i] = (int)kblk3.size();
                const double* r0 = tab2.data() + d.rec_off[i];
                // weights of the top-order sums carry (x_th - x_j)^(P+1): the kernel adds w'_p1 (g E h)_j and applies θ^-(P+1) once
                auto push2 = [&](const double* r, bool dummy) {
                    tab3.push_back(r[REC_TMX]);
                    tab3.push_back(r[REC_LSUM]);
                    const double tp = tpp_ztab(P) ? pow(r[REC_TMX], (double)(P + 1)) : 1.0;
                    for (int p = 0; p < S2 - 2; ++p) tab3.push_back((p <= P && !dummy) ? r[REC_W + p] * tp : 0.0);
                };
                for (int c = 0; c < TPP_N_CLASSES; ++c) d.near_cls_end[i][c] = 0;
                int prev_class = 0;
                for (int q = 0; q < n_near; ++q) {
                    push2(r0 + (size_t)q * R, false);
                    if ((q + 1) % tpp_npl(P) == 0) {
                        // degree of the block = degree of its last node, rounded up to its class (both near-zone loops of the
                        // kernel use this value, so a parcel's result does not depend on the loop its warp runs)
                        const int kq = (int)r0[(size_t)q * R + REC_K];
                        int cls = 0;
                        while (cls < TPP_N_CLASSES - 1 && tpp_taylor_class(cls) < kq) ++cls;
                        if (cls < prev_class) return fail(CLOUDY_ERR_STATE, "internal: near-zone Taylor degrees are not sorted");
                        prev_class = cls;
                        kblk3.push_back((double)tpp_taylor_class(cls));
                        for (int c = cls; c < TPP_N_CLASSES; ++c) d.near_cls_end[i][c] = (int)((q + 1) / tpp_npl(P));
                    }
                }
                int n_far2 = 0;
                for (int q = n_near; q < n_near + n_far; ++q) {
                    const double* r = r0 + (size_t)q * R;
                    bool zero_w = true;
                    for (int p = 0; p < d.M; ++p) zero_w = zero_w && r[REC_W + p] == 0.0;
                    if (q >= n_near + n_far - tpp_npl(P) && zero_w) continue;  // padding of the 3-node layout
                    push2(r, false);
                    ++n_far2;
                }
                do { push2(r0 + (size_t)(n_near + n_far - 1) * R, true); ++n_far2; } while (n_far2 % TPP_NPLF);
                d.rec2_far[i] = n_far2;
            }
            mpmax = std::max(mpmax, d.Mp[i]);
            if (!moving && tpp_ztab(P)) {
                // Z[p1][p] = sum_j w_j dx x_j^p1 (g E z^p)_j = exp(e0 + k L) θ^-p G_{p1,p}(k) with the NODE-ONLY functions
                //   G_{p1,p}(k) = sum_j w_j dx x_j^p1 (x_th - x_j)^p exp(k (ls_j - L)),  ls_j = ln x_j + ln(x_th - x_j),  L = max_j ls_j
                // (a mixture of decaying exponentials in k with rates <= ~12): tabulated here as degree-7 polynomials on kZtN
                // intervals of [0, kZtKmax] (interpolation at Chebyshev nodes, evaluated in long double; truncation
                // (12 h/2)^8 / (8! 2^7) < 1e-16 relative).  The kernel then sums only the top-order terms node by node.
                const int P1 = P + 1, NZ = P1 * (P1 + 1) / 2;
                long double Lmax = -INFINITY;
                for (int j = 0; j < nb; ++j) Lmax = std::max(Lmax, (long double)ell[j] + (long double)lz[j]);
                d.zt_L[i] = (double)Lmax;
                std::vector<long double> dj(nb);
                for (int j = 0; j < nb; ++j) dj[j] = ((long double)ell[j] + (long double)lz[j]) - (long double)d.zt_L[i];
                std::vector<long double> cw((size_t)NZ * nb);
                for (int p1 = 0; p1 < P1; ++p1)
                    for (int pp = p1; pp < P1; ++pp)
                        for (int j = 0; j < nb; ++j)
                            cw[(size_t)tri_ct(p1, pp, P1) * nb + j] =
                                (long double)(w[j] * g_dx * pow(xj[j], (double)p1)) * powl((long double)tmx[j], (long double)pp);
                // Chebyshev -> monomial conversion matrix for degree 7
                long double Tm[8][8] = {};
                Tm[0][0] = 1.0L; Tm[1][1] = 1.0L;
                for (int m = 1; m < 7; ++m)
                    for (int c = 0; c < 8; ++c) Tm[m + 1][c] = (c > 0 ? 2.0L * Tm[m][c - 1] : 0.0L) - Tm[m - 1][c];
                long double tq[8], Tq[8][8];
                for (int q = 0; q < 8; ++q) {
                    tq[q] = cosl(3.14159265358979323846264338327950288L * (q + 0.5L) / 8.0L);
                    for (int m = 0; m < 8; ++m) Tq[q][m] = cosl(m * 3.14159265358979323846264338327950288L * (q + 0.5L) / 8.0L);
                }
                const long double hh = (long double)kZtKmax / kZtN;
                std::vector<long double> ej(nb);
                std::vector<double>& zt = ztab;
                d.zt_off[i] = (int)zt.size();
                zt.resize(zt.size() + (size_t)kZtN * (NZ + 1) * 8);  // per interval: NZ sums, then 1/Γ(k+1)
                double* zo = zt.data() + d.zt_off[i];
                for (int iv = 0; iv < kZtN; ++iv) {
                    long double fv[8][MAXT + 1];
                    for (int q = 0; q < 8; ++q) {
                        const long double kk = (iv + 0.5L * (tq[q] + 1.0L)) * hh;
                        fv[q][NZ] = 1.0L / tgammal(kk + 1.0L);
                        for (int j = 0; j < nb; ++j) ej[j] = expl(kk * dj[j]);
                        for (int t = 0; t < NZ; ++t) {
                            long double acc = 0.0L;
                            const long double* c = cw.data() + (size_t)t * nb;
                            for (int j = 0; j < nb; ++j) acc += c[j] * ej[j];
                            fv[q][t] = acc;
                        }
                    }
                    for (int t = 0; t < NZ + 1; ++t) {
                        long double a[8], mono[8] = {};
                        for (int m = 0; m < 8; ++m) {
                            long double sacc = 0.0L;
                            for (int q = 0; q < 8; ++q) sacc += fv[q][t] * Tq[q][m];
                            a[m] = sacc * (m == 0 ? 0.125L : 0.25L);
                        }
                        for (int m = 0; m < 8; ++m)
                            for (int c = 0; c < 8; ++c) mono[c] += a[m] * Tm[m][c];
                        for (int c = 0; c < 8; ++c) zo[((size_t)iv * (NZ + 1) + t) * 8 + c] = (double)mono[c];
                    }
                }
            }
        }
    }
    if (any_ln) {
        // Gauss-Legendre rule on [-1, 1], nodes then weights
        const int n = kLognormalGL;
        d.gl_off = (int)tab2.size();
        d.gl_n = n;
        std::vector<double> xs(n), ws(n);
        gauss_legendre_rule(n, xs.data(), ws.data());
        tab2.insert(tab2.end(), xs.begin(), xs.end());
        tab2.insert(tab2.end(), ws.begin(), ws.end());
    }
    d.bins_per_log_unit = cfg->bins_per_log_unit > 0 ? cfg->bins_per_log_unit : 15;
    d.tab_total = (int)tab.size();
    d.tpp_off = (int)tab.size();
    d.tpp_total = (int)tab2.size();
    tab.insert(tab.end(), tab2.begin(), tab2.end());
    if (tab.size() & 1) tab.push_back(0.0);  // 16-byte alignment of the records below (cudaMalloc aligns the base)
    d.tpp2_off = (int)tab.size();
    d.gl2_off = 0;
    if (any_ln) tab.insert(tab.end(), tab2.begin() + d.gl_off, tab2.begin() + d.gl_off + 2 * d.gl_n);  // 256 doubles: keeps the alignment
    {
        const int rec_base = (int)tab.size() - d.tpp2_off;
        for (int i = 0; i < N; ++i) if (d.quad[i]) d.rec2_off[i] += rec_base;
        tab.insert(tab.end(), tab3.begin(), tab3.end());
        const int k_base = (int)tab.size() - d.tpp2_off;
        for (int i = 0; i < N; ++i) if (d.quad[i]) d.kblk2_off[i] += k_base;
        tab.insert(tab.end(), kblk3.begin(), kblk3.end());
        if (tab.size() & 1) tab.push_back(0.0);
    }
    d.tpp2_total = (int)tab.size() - d.tpp2_off;
    // MovingThreshold: ln x_p(k) tables of the Gamma modes (not staged in shared memory; filled on the device below)
    d.xp_n = kXpN;
    d.xp_k0 = kXpK0;
    d.xp_inv_h = (double)(kXpN - 1) / (std::max(d.k_hi, 1.0) + 0.05 - kXpK0);
    for (int i = 0; i < N; ++i) {
        d.xp_off[i] = 0;
        if (moving && d.quad[i] && cfg->kind[i] == CLOUDY_GAMMA) {
            d.xp_off[i] = (int)tab.size();
            tab.resize(tab.size() + kXpN, 0.0);
        }
    }
    {
        while (tab.size() & 15) tab.push_back(0.0);  // 128-byte alignment of the polynomial records (cudaMalloc aligns the base)
        const int zbase = (int)tab.size();
        for (int i = 0; i < N; ++i) d.zt_off[i] = (d.quad[i] && !moving && tpp_ztab(P)) ? d.zt_off[i] + zbase : 0;
        tab.insert(tab.end(), ztab.begin(), ztab.end());
        d.zt_inv_h = (double)kZtN / kZtKmax;
        d.zt_n = kZtN;
    }
    // every check that can still fail comes BEFORE the old configuration is touched: on any error the context keeps
    // its previous, fully valid configuration (tables included)
    d.n_vel = cfg->n_vel;
    if (d.n_vel < 0 || d.n_vel > CLOUDY_MAX_VEL) return fail(CLOUDY_ERR_ARG, "n_vel out of range");
    for (int v = 0; v < d.n_vel; ++v) {
        if (!(cfg->vel[v][1] >= 0.0 && cfg->vel[v][1] < 2.0)) return fail(CLOUDY_ERR_UNSUPPORTED, "terminal-velocity exponents must be in [0, 2)");
        d.velv[v] = cfg->vel[v][0] * pow(cfg->norms[1], cfg->vel[v][1]);  // rainshaft_helpers.jl:75
        d.velb[v] = cfg->vel[v][1];
        d.gam_b1[v] = tgamma(1.0 + cfg->vel[v][1]);
    }
    d.nz = cfg->nz > 0 ? cfg->nz : 1;
    d.dz = cfg->dz;
    // Γ(k+β+1)/Γ(k+1) per velocity term: degree-7 Chebyshev interpolants on kGrIntervals intervals of [0, k_max], stored as
    // monomial coefficients in t in [-1, 1] (special.cuh gamma_ratio_tab)
    {
        const double k_max = std::max(d.k_hi, 1.0) + 0.01;
        const double h = k_max / kGrIntervals;
        d.gr_inv_h = 1.0 / h;
        // T_n(t) as monomials
        double T[kGrCoef][kGrCoef] = {{0}};
        T[0][0] = 1.0; T[1][1] = 1.0;
        for (int n = 2; n < kGrCoef; ++n)
            for (int m = 0; m < kGrCoef; ++m) T[n][m] = (m > 0 ? 2.0 * T[n - 1][m - 1] : 0.0) - T[n - 2][m];
        for (int v = 0; v < d.n_vel; ++v) {
            d.gr_off[v] = (int)tab.size();
            const double beta = d.velb[v];
            for (int i = 0; i < kGrIntervals; ++i) {
                double f[kGrCoef], tn[kGrCoef];
                for (int j = 0; j < kGrCoef; ++j) {
                    tn[j] = cos(M_PI * (j + 0.5) / kGrCoef);
                    const double k = (i + 0.5 * (tn[j] + 1.0)) * h;
                    f[j] = exp(lgamma(k + beta + 1.0) - lgamma(k + 1.0));
                }
                double mono[kGrCoef] = {0};
                for (int n = 0; n < kGrCoef; ++n) {
                    double cn = 0.0;
                    for (int j = 0; j < kGrCoef; ++j) cn += f[j] * cos(M_PI * n * (j + 0.5) / kGrCoef);
                    cn *= (n == 0 ? 1.0 : 2.0) / kGrCoef;
                    for (int m = 0; m < kGrCoef; ++m) mono[m] += cn * T[n][m];
                }
                for (int m = 0; m < kGrCoef; ++m) tab.push_back(mono[m]);
            }
        }
    }
    double* new_tab = nullptr;
    if (!tab.empty()) {
        CUDA_TRY(cudaMalloc(&new_tab, sizeof(double) * tab.size()));
        cudaError_t e = cudaMemcpy(new_tab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice);
        for (int i = 0; i < N && e == cudaSuccess; ++i)
            if (d.xp_off[i]) {
                xp_table_kernel<<<(kXpN + 63) / 64, 64>>>(d.thr[i], d.xp_k0, 1.0 / d.xp_inv_h, kXpN, new_tab + d.xp_off[i]);
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
            }
        if (e != cudaSuccess) {
            cudaFree(new_tab);
            return fail(CLOUDY_ERR_CUDA, std::string("cloudy_config_set(tables): ") + cudaGetErrorString(e));
        }
    }
    // commit: nothing below can fail
    cudaStreamSynchronize(ctx->stream);  // kernels of the previous configuration may still read the old tables
    cudaFree(ctx->d_tab);
    ctx->d_tab = new_tab;
    d.tab = ctx->d_tab;
    ctx->mpmax = (mpmax == 0) ? 0 : (mpmax <= 4 ? 4 : (mpmax == 5 ? 5 : 7));
    ctx->dev = d;
    ctx->cfg = *cfg;
    ctx->configured = true;
    for (int i = 0; i < 3; ++i)
        if (ctx->tmp[i]) { cloudy_state_destroy(ctx->tmp[i]); ctx->tmp[i] = nullptr; }
    if (ctx->flux) { cloudy_state_destroy(ctx->flux); ctx->flux = nullptr; }
    // the sort scratch ensemble has the previous configuration's slot count
    if (ctx->sort_scratch) { cloudy_state_destroy(ctx->sort_scratch); ctx->sort_scratch = nullptr; }
    return CLOUDY_OK;
}

// ---- state ------------------------------------------------------------------------------------
int cloudy_state_create(cloudy_ctx* ctx, int64_t n_parcels, cloudy_state** out) {
    if (!ctx || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    if (n_parcels < 0) return fail(CLOUDY_ERR_ARG, "negative parcel count");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cloudy_state* s = new cloudy_state();
    s->ctx = ctx;
    s->n = n_parcels;
    s->nslots = ctx->dev.nslots;
    s->stride = ((n_parcels + 31) / 32) * 32;  // 256-byte aligned slot columns
    if (s->stride == 0) s->stride = 32;
    s->d = nullptr;
    s->order = nullptr;
    s->sort_age = 0;
    cudaError_t e = cudaMalloc(&s->d, sizeof(double) * s->stride * s->nslots);
    if (e != cudaSuccess) {
        delete s;
        return fail(CLOUDY_ERR_CUDA, std::string("cudaMalloc(state): ") + cudaGetErrorString(e));
    }
    CUDA_TRY(cudaMemsetAsync(s->d, 0, sizeof(double) * s->stride * s->nslots, ctx->stream));
    *out = s;
    return CLOUDY_OK;
}

int cloudy_state_destroy(cloudy_state* st) {
    if (!st) return CLOUDY_OK;
    cudaSetDevice(st->ctx->device);
    cudaStreamSynchronize(st->ctx->stream);
    cudaFree(st->d);
    if (st->order && --st->order->refs == 0) { cudaFree(st->order->d); delete st->order; }
    delete st;
    return CLOUDY_OK;
}

int cloudy_state_device_ptr(const cloudy_state* st, double** dptr, int64_t* stride, int32_t* n_slots) {
    if (!st) return fail(CLOUDY_ERR_ARG, "state is NULL");
    if (dptr) *dptr = st->d;
    if (stride) *stride = st->stride;
    if (n_slots) *n_slots = st->nslots;
    return CLOUDY_OK;
}

static int ensure_stage(cloudy_ctx* ctx, long long n_doubles) {
    if (ctx->stage_cap >= n_doubles) return CLOUDY_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_stage_aos);
    ctx->d_stage_aos = nullptr;
    ctx->stage_cap = 0;
    CUDA_TRY(cudaMalloc(&ctx->d_stage_aos, sizeof(double) * n_doubles));
    ctx->stage_cap = n_doubles;
    return CLOUDY_OK;
}

int cloudy_state_upload(cloudy_ctx* ctx, cloudy_state* st, const double* host, int64_t n_parcels) {
    if (!ctx || !st || (!host && n_parcels > 0)) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (n_parcels != st->n) return fail(CLOUDY_ERR_ARG, "parcel count does not match the state");
    if (n_parcels == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const long long total = (long long)n_parcels * st->nslots;
    int rc = ensure_stage(ctx, total);
    if (rc) return rc;
    state_set_order(st, nullptr);  // the host's order
    st->sort_age = 0;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_stage_aos, host, sizeof(double) * total, cudaMemcpyHostToDevice, ctx->stream));
    int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    aos_to_soa_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_stage_aos, st->d, n_parcels, st->nslots, st->stride);
    CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return CLOUDY_OK;
}

int cloudy_state_download(cloudy_ctx* ctx, const cloudy_state* st, double* host, int64_t n_parcels) {
    if (!ctx || !st || (!host && n_parcels > 0)) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (n_parcels != st->n) return fail(CLOUDY_ERR_ARG, "parcel count does not match the state");
    if (n_parcels == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const long long total = (long long)n_parcels * st->nslots;
    int rc = ensure_stage(ctx, total);
    if (rc) return rc;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    if (st->order)  // regime-ordered ensemble: every parcel goes back to its original index
        soa_to_aos_ordered_kernel<<<blocks, 256, 0, ctx->stream>>>(st->d, ctx->d_stage_aos, n_parcels, st->nslots, st->stride, st->order->d);
    else
        soa_to_aos_kernel<<<blocks, 256, 0, ctx->stream>>>(st->d, ctx->d_stage_aos, n_parcels, st->nslots, st->stride);
    CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(host, ctx->d_stage_aos, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CLOUDY_OK;
}

int cloudy_state_copy(cloudy_ctx* ctx, const cloudy_state* src, cloudy_state* dst) {
    if (!ctx || !src || !dst) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (src->n != dst->n || src->nslots != dst->nslots) return fail(CLOUDY_ERR_ARG, "state shapes differ");
    CUDA_TRY(cudaMemcpyAsync(dst->d, src->d, sizeof(double) * src->stride * src->nslots, cudaMemcpyDeviceToDevice, ctx->stream));
    state_set_order(dst, src->order);
    dst->sort_age = src->sort_age;
    return CLOUDY_OK;
}

// ---- hot path -----------------------------------------------------------------------------------
static int check_pair(cloudy_ctx* ctx, const cloudy_state* a, const cloudy_state* b) {
    if (!ctx || !a || !b) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    if (a->n != b->n || a->nslots != ctx->dev.nslots || b->nslots != ctx->dev.nslots) return fail(CLOUDY_ERR_ARG, "state shapes differ");
    return CLOUDY_OK;
}

static KArgs base_args(cloudy_ctx* ctx, const cloudy_state* in, cloudy_state* out) {
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.u_in = in->d; a.s_in = in->stride;
    a.out = out->d; a.s_out = out->stride;
    a.n = in->n;
    a.ps_in = 1; a.ps_out = 1;
    a.cn = 0; a.ci = 1; a.cf = 1; a.dt = 1; a.div = 1;
    a.err_count = ctx->d_err;
    return a;
}

int cloudy_coal_tendency(cloudy_ctx* ctx, const cloudy_state* m, cloudy_state* dm) {
    int rc = check_pair(ctx, m, dm);
    if (rc) return rc;
    if (m->n == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // large ensembles are kept in regime order: the first evaluation moves the parcels (cloudy_state_regime_sort), every
    // later one reads and writes whole lines.  `m` is logically const — same parcels, same values, new positions.
    if (!m->order && sort_wanted(ctx, m->n) && tpp_selected(ctx, CLOUDY_MODEL_BOX)) {
        rc = regime_sort_state(ctx, const_cast<cloudy_state*>(m));
        if (rc) return rc;
    }
    state_set_order(dm, m->order);
    KArgs a = base_args(ctx, m, dm);
    a.tend_only = 1;
    a.presorted = 1;
    return launch_rhs(ctx, CLOUDY_MODEL_BOX, a);
}

int cloudy_sedimentation_flux(cloudy_ctx* ctx, const cloudy_state* m, cloudy_state* flux) {
    int rc = check_pair(ctx, m, flux);
    if (rc) return rc;
    if (m->n == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    state_set_order(flux, m->order);
    KArgs a = base_args(ctx, m, flux);
    a.tend_only = 1;
    a.flux_only = 1;
    return launch_rhs(ctx, CLOUDY_MODEL_RAINSHAFT, a);
}

int cloudy_rainshaft_rhs(cloudy_ctx* ctx, cloudy_state* m, cloudy_state* dm) {
    int rc = check_pair(ctx, m, dm);
    if (rc) return rc;
    if (m->n == 0) return CLOUDY_OK;
    if (m->n % ctx->dev.nz != 0) return fail(CLOUDY_ERR_ARG, "cell count is not a multiple of nz");
    if (!(ctx->dev.dz > 0)) return fail(CLOUDY_ERR_ARG, "dz must be positive");
    if (m->order) return fail(CLOUDY_ERR_STATE, "the column model needs cells in column order; this state was regime-sorted (upload it again)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    state_set_order(dm, nullptr);
    KArgs a = base_args(ctx, m, dm);
    a.tend_only = 1;
    a.clip_back = m->d;
    a.s_clip = m->stride;
    return launch_rhs(ctx, CLOUDY_MODEL_RAINSHAFT, a);
}

static int ensure_flux(cloudy_ctx* ctx, long long n) {
    if (ctx->flux && ctx->flux->n == n) return CLOUDY_OK;
    if (ctx->flux) { cloudy_state_destroy(ctx->flux); ctx->flux = nullptr; }
    return cloudy_state_create(ctx, n, &ctx->flux);
}

static int ensure_tmp(cloudy_ctx* ctx, int idx, long long n) {
    if (ctx->tmp[idx] && ctx->tmp[idx]->n == n) return CLOUDY_OK;
    if (ctx->tmp[idx]) { cloudy_state_destroy(ctx->tmp[idx]); ctx->tmp[idx] = nullptr; }
    return cloudy_state_create(ctx, n, &ctx->tmp[idx]);
}

int cloudy_ssprk33_steps(cloudy_ctx* ctx, cloudy_state* u, double dt, int32_t n_steps, int32_t model) {
    if (!ctx || !u) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    if (model != CLOUDY_MODEL_BOX && model != CLOUDY_MODEL_RAINSHAFT) return fail(CLOUDY_ERR_ARG, "unknown model");
    if (n_steps < 0) return fail(CLOUDY_ERR_ARG, "negative step count");
    if (u->n == 0 || n_steps == 0) return CLOUDY_OK;
    if (model == CLOUDY_MODEL_RAINSHAFT) {
        if (u->n % ctx->dev.nz != 0) return fail(CLOUDY_ERR_ARG, "cell count is not a multiple of nz");
        if (!(ctx->dev.dz > 0)) return fail(CLOUDY_ERR_ARG, "dz must be positive");
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_tmp(ctx, 0, u->n))) return rc;
    if ((rc = ensure_tmp(ctx, 1, u->n))) return rc;
    if ((rc = ensure_tmp(ctx, 2, u->n))) return rc;
    // u^n lives in `cur`; stage results ping-pong through the two other buffers; the caller's buffer
    // receives the final state.
    double* cur = u->d;
    double* t1 = ctx->tmp[0]->d;
    double* t2 = ctx->tmp[1]->d;
    double* nxt = ctx->tmp[2]->d;
    const long long st = u->stride;  // all four share the stride (same n)
    // box model: the ensemble is resident in regime order (regime_sort.cuh); the order is refreshed every
    // `resort_interval` steps from the current state (a scheduling hint only: results do not depend on it)
    const bool box = (model == CLOUDY_MODEL_BOX);
    if (!box && u->order) return fail(CLOUDY_ERR_STATE, "the column model needs cells in column order; this state was regime-sorted (upload it again)");
    const bool resident = box && sort_wanted(ctx, u->n) && tpp_selected(ctx, model);
    for (int step = 0; step < n_steps; ++step) {
        if (resident && (!u->order || u->sort_age >= ctx->resort_interval)) {
            OrderBuf* o = order_acquire(ctx, u->n);
            if (!o) return fail(CLOUDY_ERR_CUDA, "cudaMalloc(order) failed");
            if ((rc = sort_buffer(ctx, cur, nxt, st, u->nslots, u->n, u->order ? u->order->d : nullptr, o->d))) { order_release(ctx, o); return rc; }
            std::swap(cur, nxt);
            order_release(ctx, u->order);
            u->order = o;
            u->sort_age = 0;
        }
        KArgs a;
        memset(&a, 0, sizeof(a));
        a.n = u->n; a.dt = dt; a.err_count = ctx->d_err;
        a.s_in = a.s_n = a.s_out = st;
        a.ps_in = a.ps_out = 1;
        a.presorted = box ? 1 : 0;
        // stage 1: tmp = u + dt f(u)
        a.u_in = cur; a.u_n = nullptr; a.out = t1; a.cn = 0; a.ci = 1; a.cf = 1; a.div = 1;
        ctx->perm_valid = false;  // column model: new permutation sort (if enabled) at the first stage, reused by stages 2 and 3
        // (reusing it across steps was measured: 1.27 -> 1.42 ms per C3 step, sedimentation moves the cloud edge every step)
        ctx->perm_fresh = false;
        if ((rc = launch_rhs(ctx, model, a))) return rc;
        ctx->perm_valid = ctx->perm_fresh;
        // stage 2: tmp = (3u + tmp + dt f(tmp))/4
        a.u_in = t1; a.u_n = cur; a.out = t2; a.cn = 3; a.ci = 1; a.cf = 1; a.div = 4;
        if ((rc = launch_rhs(ctx, model, a))) return rc;
        // stage 3: u = (u + 2 tmp + 2 dt f(tmp))/3
        a.u_in = t2; a.u_n = cur; a.out = nxt; a.cn = 1; a.ci = 2; a.cf = 2; a.div = 3;
        if ((rc = launch_rhs(ctx, model, a))) return rc;
        std::swap(cur, nxt);
        u->sort_age++;
    }
    ctx->perm_valid = false;
    if (cur != u->d) {
        // `cur` is a work buffer: hand its storage to the caller's state
        if (cur == ctx->tmp[2]->d) std::swap(u->d, ctx->tmp[2]->d);
        else return fail(CLOUDY_ERR_STATE, "internal: stepper buffers out of order");
    }
    return CLOUDY_OK;
}

int cloudy_state_regime_sort(cloudy_ctx* ctx, cloudy_state* st) {
    if (!ctx || !st) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    if (st->nslots != ctx->dev.nslots) return fail(CLOUDY_ERR_ARG, "state shape does not match the configuration");
    if (ctx->dev.nz > 1) return fail(CLOUDY_ERR_STATE, "column states keep their cells in column order");
    if (st->n >= (1LL << 31)) return fail(CLOUDY_ERR_UNSUPPORTED, "regime order supports fewer than 2^31 parcels per device");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return regime_sort_state(ctx, st);
}

int cloudy_state_order(const cloudy_state* st, const int32_t** d_order) {
    if (!st || !d_order) return fail(CLOUDY_ERR_ARG, "NULL argument");
    *d_order = st->order ? st->order->d : nullptr;
    return CLOUDY_OK;
}

int cloudy_set_resort_interval(cloudy_ctx* ctx, int32_t steps) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    if (steps < 1) return fail(CLOUDY_ERR_ARG, "the resort interval must be at least one step");
    ctx->resort_interval = steps;
    return CLOUDY_OK;
}

int cloudy_sort_count(cloudy_ctx* ctx, int64_t* out) {
    if (!ctx || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    *out = ctx->sorts;
    return CLOUDY_OK;
}

int cloudy_moment_sums_device(cloudy_ctx* ctx, const cloudy_state* u, double* d_out) {
    if (!ctx || !u || !d_out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    moment_partial_kernel<<<SUM_BLOCKS, SUM_THREADS, 0, ctx->stream>>>(u->d, u->n, u->stride, u->nslots, ctx->d_partial);
    CUDA_TRY(cudaGetLastError());
    moment_final_kernel<<<1, SUM_THREADS, 0, ctx->stream>>>(ctx->d_partial, u->nslots, d_out);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    return CLOUDY_OK;
}

int cloudy_moment_sums(cloudy_ctx* ctx, const cloudy_state* u, double* host_out) {
    if (!ctx || !u || !host_out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    int rc = cloudy_moment_sums_device(ctx, u, ctx->d_scratch);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(host_out, ctx->d_scratch, sizeof(double) * u->nslots, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CLOUDY_OK;
}


// ---- multi-GPU: the path's only collective -------------------------------------------------------
int cloudy_comm_unique_id(void* id_out) {
    if (!id_out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    std::string err;
    NcclApi* api = nccl_api(err);
    if (!api) return fail(CLOUDY_ERR_UNSUPPORTED, err);
    NcclId id;
    NCCL_TRY(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return CLOUDY_OK;
}

int cloudy_comm_init(cloudy_ctx* ctx, int32_t n_ranks, int32_t rank, const void* unique_id) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(CLOUDY_ERR_ARG, "invalid rank / n_ranks");
    if (ctx->nccl_comm) return fail(CLOUDY_ERR_STATE, "the context already has a communicator");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_comm_buffers(ctx);
    if (rc) return rc;
    ctx->comm_size = n_ranks;
    ctx->comm_rank = rank;
    if (n_ranks == 1) return CLOUDY_OK;  // nothing to exchange
    if (!unique_id) return fail(CLOUDY_ERR_ARG, "unique_id is NULL");
    std::string err;
    NcclApi* api = nccl_api(err);
    if (!api) return fail(CLOUDY_ERR_UNSUPPORTED, err);
    NcclId id;
    memcpy(&id, unique_id, sizeof(id));
    NCCL_TRY(api, api->CommInitRank(&ctx->nccl_comm, n_ranks, id, rank));
    return CLOUDY_OK;
}

int cloudy_comm_destroy(cloudy_ctx* ctx) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    if (ctx->nccl_comm) {
        std::string err;
        NcclApi* api = nccl_api(err);
        cudaSetDevice(ctx->device);
        if (ctx->s_comm) cudaStreamSynchronize(ctx->s_comm);
        if (api) api->CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->comm_size = 0;
    ctx->comm_rank = 0;
    return CLOUDY_OK;
}

int cloudy_comm_info(cloudy_ctx* ctx, int32_t* n_ranks, int32_t* rank, int32_t* nccl_version) {
    if (!ctx) return fail(CLOUDY_ERR_ARG, "ctx is NULL");
    if (n_ranks) *n_ranks = ctx->comm_size > 0 ? ctx->comm_size : 1;
    if (rank) *rank = ctx->comm_rank;
    if (nccl_version) {
        *nccl_version = 0;
        std::string err;
        NcclApi* api = ctx->nccl_comm ? nccl_api(err) : nullptr;
        if (api && api->GetVersion) api->GetVersion(nccl_version);
    }
    return CLOUDY_OK;
}

// Global per-slot sums: local two-pass reduction on the context's stream, then ncclAllReduce(sum, double, n_slots) and
// the copy to pinned host memory on a side stream, so that the next step's kernels (already enqueued by the caller
// after this returns with host_out == NULL) overlap the collective's latency.
int cloudy_moment_sums_allreduce(cloudy_ctx* ctx, const cloudy_state* u, double* host_out) {
    if (!ctx || !u) return fail(CLOUDY_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_comm_buffers(ctx);
    if (rc) return rc;
    if (ctx->sums_pending) CUDA_TRY(cudaStreamSynchronize(ctx->s_comm));  // the previous result is overwritten
    if ((rc = cloudy_moment_sums_device(ctx, u, ctx->d_sums))) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev_sums, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_comm, ctx->ev_sums, 0));
    const double* src = ctx->d_sums;
    if (ctx->nccl_comm) {
        std::string err;
        NcclApi* api = nccl_api(err);
        if (!api) return fail(CLOUDY_ERR_UNSUPPORTED, err);
        NCCL_TRY(api, api->AllReduce(ctx->d_sums, ctx->d_sums + MAXSLOT, (size_t)u->nslots, 8 /*ncclFloat64*/, 0 /*ncclSum*/, ctx->nccl_comm,
                                     ctx->s_comm));
        src = ctx->d_sums + MAXSLOT;
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->h_sums, src, sizeof(double) * u->nslots, cudaMemcpyDeviceToHost, ctx->s_comm));
    // (d_sums is rewritten by the next call only after s_comm has been synchronised: above, or in cloudy_moment_sums_fetch)
    ctx->sums_pending = true;
    ctx->sums_n = u->nslots;
    if (host_out) return cloudy_moment_sums_fetch(ctx, host_out);
    return CLOUDY_OK;
}

int cloudy_moment_sums_fetch(cloudy_ctx* ctx, double* host_out) {
    if (!ctx || !host_out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->sums_pending) return fail(CLOUDY_ERR_STATE, "no all-reduce is pending");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_comm));
    memcpy(host_out, ctx->h_sums, sizeof(double) * ctx->sums_n);
    ctx->sums_pending = false;
    return CLOUDY_OK;
}

static int ensure_pipe(cloudy_ctx* ctx, long long chunk) {
    if (ctx->pipe_ready && ctx->pipe_chunk >= chunk) return CLOUDY_OK;
    if (!ctx->pipe_ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
        ctx->pipe_ready = true;
    }
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 3; ++i) {
        cudaFree(ctx->d_pipe_in[i]);
        cudaFree(ctx->d_pipe_out[i]);
        ctx->d_pipe_in[i] = ctx->d_pipe_out[i] = nullptr;
    }
    ctx->pipe_chunk = 0;
    for (int i = 0; i < 3; ++i) {
        CUDA_TRY(cudaMalloc(&ctx->d_pipe_in[i], sizeof(double) * chunk * MAXSLOT));
        CUDA_TRY(cudaMalloc(&ctx->d_pipe_out[i], sizeof(double) * chunk * MAXSLOT));
    }
    ctx->pipe_chunk = chunk;
    return CLOUDY_OK;
}

// Host moments in, host tendencies out.  The batch is cut into chunks that flow through a three-stage pipeline
// (H2D copy | kernel | D2H copy) on three streams, so PCIe traffic in both directions overlaps the arithmetic.
// The kernel reads and writes the host (array-of-structures) layout directly: no transpose pass.
int cloudy_coal_tendency_host(cloudy_ctx* ctx, const double* host_m, double* host_dm, int64_t n_parcels) {
    if (!ctx || (!host_m && n_parcels) || (!host_dm && n_parcels)) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    if (n_parcels < 0) return fail(CLOUDY_ERR_ARG, "negative parcel count");
    if (n_parcels == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int ns = ctx->dev.nslots;
    // measured on B200 / PCIe Gen5 x16 (tools/e2e_chunks.py, tools/pcie_probe.py: 55 GB/s one way, 44 GB/s each way when both
    // directions run): 128 Ki-parcel chunks for ensembles around 1 Mi parcels (short fill/drain), 512 Ki from 4 Mi parcels
    // on (41 GB/s each way, 93 % of the link's bidirectional rate)
    long long chunk = n_parcels >= (1LL << 22) ? 524288 : 131072;
    if (const char* e = getenv("CLOUDY_PIPE_CHUNK")) chunk = std::max<long long>(1024, atoll(e));
    if (n_parcels < chunk) chunk = n_parcels;
    int rc = ensure_pipe(ctx, std::max<long long>(chunk, 1024));
    if (rc) return rc;
    const long long n_chunks = (n_parcels + chunk - 1) / chunk;
    // the pipeline starts after everything already enqueued on the context's stream
    CUDA_TRY(cudaEventRecord(ctx->ev_k[0], ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_k[0], 0));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[0], 0));
    for (long long c = 0; c < n_chunks; ++c) {
        const int b = (int)(c % 3);
        const long long p0 = c * chunk;
        const long long m = std::min<long long>(chunk, n_parcels - p0);
        // buffer b is free once chunk c-3 has been copied back (input buffer: once its kernel has run)
        if (c >= 3) {
            CUDA_TRY(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_k[b], 0));
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));
        }
        CUDA_TRY(cudaMemcpyAsync(ctx->d_pipe_in[b], host_m + p0 * ns, sizeof(double) * m * ns, cudaMemcpyHostToDevice, ctx->s_h2d));
        CUDA_TRY(cudaEventRecord(ctx->ev_in[b], ctx->s_h2d));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        KArgs a;
        memset(&a, 0, sizeof(a));
        a.u_in = ctx->d_pipe_in[b]; a.s_in = 1; a.ps_in = ns;
        a.out = ctx->d_pipe_out[b]; a.s_out = 1; a.ps_out = ns;
        a.n = m;
        a.cn = 0; a.ci = 1; a.cf = 1; a.dt = 1; a.div = 1;
        a.tend_only = 1;
        a.err_count = ctx->d_err;
        if ((rc = launch_rhs(ctx, CLOUDY_MODEL_BOX, a))) return rc;
        CUDA_TRY(cudaEventRecord(ctx->ev_k[b], ctx->stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[b], 0));
        CUDA_TRY(cudaMemcpyAsync(host_dm + p0 * ns, ctx->d_pipe_out[b], sizeof(double) * m * ns, cudaMemcpyDeviceToHost, ctx->s_d2h));
        CUDA_TRY(cudaEventRecord(ctx->ev_out[b], ctx->s_d2h));
    }
    // later work on the context's stream is ordered after the last copy-back; the call itself is synchronous
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[(n_chunks - 1) % 3], 0));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_d2h));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CLOUDY_OK;
}

int cloudy_error_count(cloudy_ctx* ctx, int64_t* n_invalid) {
    if (!ctx || !n_invalid) return fail(CLOUDY_ERR_ARG, "NULL argument");
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, ctx->d_err, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_err, 0, sizeof(v), ctx->stream));
    *n_invalid = (int64_t)v;
    return CLOUDY_OK;
}

// ---- single-object entry points -------------------------------------------------------------------
static int run_scalar(cloudy_ctx* ctx, ScalarArgs& a, double* host_out, int n_out, int32_t* host_iout) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    a.out = ctx->d_scratch;
    a.iout = (int*)(ctx->d_scratch + 32);
    scalar_kernel<<<1, 32, 0, ctx->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(host_out, a.out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_iout) CUDA_TRY(cudaMemcpyAsync(host_iout, a.iout, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CLOUDY_OK;
}

static int kind_nparams(int kind) { return (kind == CLOUDY_GAMMA || kind == CLOUDY_LOGNORMAL) ? 3 : 2; }

int cloudy_moment(cloudy_ctx* ctx, int32_t kind, const double* params, double q, double* out) {
    if (!ctx || !params || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (kind < 0 || kind > 3) return fail(CLOUDY_ERR_ARG, "unknown distribution kind");
    ScalarArgs a;
    memset(&a, 0, sizeof(a));
    a.op = 0; a.kind = kind; a.q = q;
    a.params[2] = 1.0;
    for (int i = 0; i < kind_nparams(kind); ++i) a.params[i] = params[i];
    return run_scalar(ctx, a, out, 1, nullptr);
}

int cloudy_update_dist_from_moments(cloudy_ctx* ctx, int32_t kind, const double* moments, const double* range, double* params_out,
                                    int32_t* invalid) {
    if (!ctx || !moments || !params_out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (kind < 0 || kind > 3) return fail(CLOUDY_ERR_ARG, "unknown distribution kind");
    ScalarArgs a;
    memset(&a, 0, sizeof(a));
    a.op = 1; a.kind = kind;
    for (int i = 0; i < kind_nparams(kind); ++i) a.params[i] = moments[i];
    if (kind == CLOUDY_GAMMA) {
        a.range[0] = range ? range[0] : kEps;
        a.range[1] = range ? range[1] : 10.0;
    } else if (kind == CLOUDY_LOGNORMAL) {
        a.range[0] = range ? range[0] : -INFINITY;
        a.range[1] = range ? range[1] : INFINITY;
        a.range[2] = range ? range[2] : kEps;
        a.range[3] = range ? range[3] : INFINITY;
    }
    double o[3];
    int32_t inv = 0;
    int rc = run_scalar(ctx, a, o, 3, &inv);
    if (rc) return rc;
    for (int i = 0; i < kind_nparams(kind); ++i) params_out[i] = o[i];
    if (invalid) *invalid = inv;
    return CLOUDY_OK;
}

int cloudy_moment_source_helper(cloudy_ctx* ctx, int32_t kind, const double* params, double p1, double p2, double x_threshold,
                                int32_t n_bins_per_log_unit, double* out) {
    if (!ctx || !params || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (kind == CLOUDY_MONODISPERSE) {  // closed form, still evaluated on the device through moment()
        ScalarArgs a;
        memset(&a, 0, sizeof(a));
        a.op = 0; a.kind = CLOUDY_MONODISPERSE; a.q = p1 + p2;
        a.params[0] = params[0] * params[0]; a.params[1] = params[1]; a.params[2] = 1.0;
        double v;
        int rc = run_scalar(ctx, a, &v, 1, nullptr);
        if (rc) return rc;
        *out = (params[1] < x_threshold / 2) ? v : 0.0;
        return CLOUDY_OK;
    }
    if (kind == CLOUDY_LOGNORMAL) {
        if (!(x_threshold > 0)) return fail(CLOUDY_ERR_ARG, "x_threshold must be positive");
        if (!(params[2] > 0)) return fail(CLOUDY_ERR_ARG, "sigma needs to be positive");
        CUDA_TRY(cudaSetDevice(ctx->device));
        double rule[2 * kLognormalGL];
        gauss_legendre_rule(kLognormalGL, rule, rule + kLognormalGL);
        CUDA_TRY(cudaMemcpyAsync(ctx->d_scratch + 64, rule, sizeof(rule), cudaMemcpyHostToDevice, ctx->stream));
        ScalarArgs a;
        memset(&a, 0, sizeof(a));
        a.op = 6; a.kind = kind; a.p1 = p1; a.p2 = p2; a.x_th = x_threshold; a.n_bins = kLognormalGL; a.y = ctx->d_scratch + 64;
        for (int i = 0; i < 3; ++i) a.params[i] = params[i];
        return run_scalar(ctx, a, out, 1, nullptr);  // (synchronises: `rule` stays alive until the copy has landed)
    }
    if (kind != CLOUDY_EXPONENTIAL && kind != CLOUDY_GAMMA) return fail(CLOUDY_ERR_ARG, "unknown distribution kind");
    if (!(x_threshold > 0)) return fail(CLOUDY_ERR_ARG, "x_threshold must be positive");
    ScalarArgs a;
    memset(&a, 0, sizeof(a));
    a.op = 2; a.kind = kind; a.p1 = p1; a.p2 = p2; a.x_th = x_threshold;
    a.params[2] = 1.0;
    for (int i = 0; i < kind_nparams(kind); ++i) a.params[i] = params[i];
    // node grid, ParticleDistributions.jl:579-582
    const double x_lb = std::min(1e-5, 1e-5 * x_threshold);
    a.n_bins = (int)floor(n_bins_per_log_unit * log10(x_threshold / x_lb));
    if (a.n_bins < 3) return fail(CLOUDY_ERR_ARG, "n_bins must be at least 3");
    a.x_min = log(x_lb);
    a.dx = (log(x_threshold) - log(x_lb)) / a.n_bins;
    return run_scalar(ctx, a, out, 1, nullptr);
}

int cloudy_compute_threshold(cloudy_ctx* ctx, int32_t kind, const double* params, double percentile, double minx, double* out) {
    if (!ctx || !params || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (kind != CLOUDY_EXPONENTIAL && kind != CLOUDY_GAMMA)
        return fail(CLOUDY_ERR_ARG, "MethodError: compute_threshold is defined for Exponential and Gamma distributions only");
    ScalarArgs a;
    memset(&a, 0, sizeof(a));
    a.op = 5; a.kind = kind; a.q = percentile; a.p1 = minx;
    a.params[2] = 1.0;
    for (int i = 0; i < kind_nparams(kind); ++i) a.params[i] = params[i];
    return run_scalar(ctx, a, out, 1, nullptr);
}

int cloudy_get_sedimentation_flux_1(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params, int32_t n_vel,
                                    const double* vel, double* out) {
    if (!ctx || !kinds || !params || !vel || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (n_modes < 1 || n_modes > MAXN || n_vel < 0 || n_vel > CLOUDY_MAX_VEL) return fail(CLOUDY_ERR_ARG, "size out of range");
    ScalarArgs a;
    memset(&a, 0, sizeof(a));
    a.op = 3; a.n_modes = n_modes; a.n_vel = n_vel;
    int nout = 0;
    for (int i = 0; i < n_modes; ++i) {
        a.kinds[i] = kinds[i];
        for (int j = 0; j < 3; ++j) a.mparams[i][j] = params[i * 3 + j];
        nout += kind_nparams(kinds[i]);
    }
    for (int v = 0; v < n_vel; ++v) { a.vel[v][0] = vel[2 * v]; a.vel[v][1] = vel[2 * v + 1]; }
    return run_scalar(ctx, a, out, nout, nullptr);
}

int cloudy_integrate_simpson(cloudy_ctx* ctx, int32_t n_bins, double dx, const double* y, double* out) {
    if (!ctx || !y || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (n_bins < 3) return fail(CLOUDY_ERR_ARG, "n_bins must be at least 3");
    if (n_bins + 1 > CLOUDY_MAX_NODES) return fail(CLOUDY_ERR_UNSUPPORTED, "too many nodes");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_scratch + 64, y, sizeof(double) * (n_bins + 1), cudaMemcpyHostToDevice, ctx->stream));
    ScalarArgs a;
    memset(&a, 0, sizeof(a));
    a.op = 4; a.n_bins = n_bins; a.dx = dx; a.y = ctx->d_scratch + 64;
    return run_scalar(ctx, a, out, 1, nullptr);
}

int cloudy_get_coal_ints_1(cloudy_ctx* ctx, const double* params, double* out) {
    if (!ctx || !params || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    // one parcel whose "state" holds the distribution parameters themselves (params_in): the kernel skips
    // update_dist_from_moments and the de-normalisation, so the output is the reference's return value.
    const DevConfig& d = ctx->dev;
    int rc;
    if ((rc = ensure_tmp(ctx, 0, 1))) return rc;
    if ((rc = ensure_tmp(ctx, 1, 1))) return rc;
    double h[MAXSLOT];
    for (int i = 0; i < d.N; ++i) {
        // given distributions bypass update_dist_from_moments' clamp: the shape must lie inside the incomplete-gamma and Z-sum tables
        if (d.quad[i] && d.kind[i] == CLOUDY_GAMMA && d.nprog[i] > 2 && !(params[3 * i + 2] > 0.0 && params[3 * i + 2] <= kZtKmax))
            return fail(CLOUDY_ERR_UNSUPPORTED, "Gamma shape parameter outside (0, 11]: outside the incomplete-gamma tables");
        for (int q = 0; q < d.nprog[i]; ++q) h[d.slot0[i] + q] = params[3 * i + q];
    }
    if ((rc = cloudy_state_upload(ctx, ctx->tmp[0], h, 1))) return rc;
    KArgs a = base_args(ctx, ctx->tmp[0], ctx->tmp[1]);
    a.tend_only = 1;
    a.params_in = 1;
    if ((rc = launch_rhs(ctx, CLOUDY_MODEL_BOX, a))) return rc;
    return cloudy_state_download(ctx, ctx->tmp[1], out, 1);
}

// ---- condensation and diagnostics ("next" rows) ---------------------------------------------------
static int launch_aux(cloudy_ctx* ctx, const void* fn, const AuxArgs& a) {
    long long blocks = std::min<long long>((a.n + 255) / 256, (long long)ctx->sm_count * 8);
    void* params[2] = {(void*)&ctx->dev, (void*)&a};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)std::max<long long>(blocks, 1)), dim3(256), params, 0, ctx->stream));
    ctx->launches++;
    return CLOUDY_OK;
}

int cloudy_cond_evap(cloudy_ctx* ctx, const cloudy_state* m, double s, const double* d_s, double xi, double rho_l, cloudy_state* dm) {
    int rc = check_pair(ctx, m, dm);
    if (rc) return rc;
    if (m->n == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    state_set_order(dm, m->order);
    AuxArgs a;
    memset(&a, 0, sizeof(a));
    a.u_in = m->d; a.s_in = m->stride; a.n = m->n; a.out = dm->d; a.s_out = dm->stride;
    a.s = s; a.d_s = d_s; a.rho_l = rho_l;
    a.xi_n = xi / pow(ctx->cfg.norms[1], 2.0 / 3.0);  // box_model_helpers.jl:65
    return launch_aux(ctx, (const void*)cond_evap_kernel, a);
}

int cloudy_standard_N_q(cloudy_ctx* ctx, const cloudy_state* m, double size_cutoff, int32_t normalized, double* d_out) {
    if (!ctx || !m || !d_out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (!ctx->configured) return fail(CLOUDY_ERR_STATE, "cloudy_config_set has not been called");
    if (m->n == 0) return CLOUDY_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    AuxArgs a;
    memset(&a, 0, sizeof(a));
    a.u_in = m->d; a.s_in = m->stride; a.n = m->n; a.out = d_out; a.s_out = m->n;
    a.cutoff = size_cutoff; a.normalized = normalized;
    a.order = m->order ? m->order->d : nullptr;  // results are indexed by parcel, not by position
    return launch_aux(ctx, (const void*)nq_kernel, a);
}

// one set of distributions given by parameters: a private one-parcel configuration (kinds only matter)
static int aux_single(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params, bool nq, double s, double xi,
                      double rho_l, double cutoff, double* out) {
    if (!ctx || !kinds || !params || !out) return fail(CLOUDY_ERR_ARG, "NULL argument");
    if (n_modes < 1 || n_modes > MAXN) return fail(CLOUDY_ERR_ARG, "n_modes must be 1..4");
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevConfig d;
    memset((void*)&d, 0, sizeof(d));
    d.N = n_modes;
    double h[MAXSLOT + 4];
    int slot = 0;
    for (int i = 0; i < n_modes; ++i) {
        if (kinds[i] < 0 || kinds[i] > 3) return fail(CLOUDY_ERR_ARG, "unknown distribution kind");
        d.kind[i] = kinds[i]; d.nprog[i] = kind_nparams(kinds[i]); d.slot0[i] = slot;
        for (int q = 0; q < d.nprog[i]; ++q) { d.norm[slot] = 1.0; h[slot++] = params[3 * i + q]; }
    }
    d.nslots = slot;
    double* dbuf = ctx->d_scratch;  // [0..slot) in, [16..) out
    CUDA_TRY(cudaMemcpyAsync(dbuf, h, sizeof(double) * slot, cudaMemcpyHostToDevice, ctx->stream));
    AuxArgs a;
    memset(&a, 0, sizeof(a));
    a.u_in = dbuf; a.s_in = 1; a.n = 1; a.out = dbuf + 16; a.s_out = 1; a.params_in = 1;
    a.s = s; a.xi_n = xi; a.rho_l = rho_l; a.cutoff = cutoff;
    void* kp[2] = {(void*)&d, (void*)&a};
    CUDA_TRY(cudaLaunchKernel(nq ? (const void*)nq_kernel : (const void*)cond_evap_kernel, dim3(1), dim3(32), kp, 0, ctx->stream));
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(out, dbuf + 16, sizeof(double) * (nq ? 4 : slot), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CLOUDY_OK;
}

int cloudy_get_cond_evap_1(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params, double s, double xi, double rho_l,
                           double* out) {
    return aux_single(ctx, n_modes, kinds, params, false, s, xi, rho_l, 0.0, out);
}

int cloudy_get_standard_N_q_1(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params, double size_cutoff,
                              double* out) {
    return aux_single(ctx, n_modes, kinds, params, true, 0.0, 0.0, 1000.0, size_cutoff, out);
}

int64_t cloudy_config_sizeof(void) { return (int64_t)sizeof(cloudy_config); }

int32_t cloudy_config_offsets(int64_t* offsets, int32_t max_fields) {
    const int64_t off[] = {
        (int64_t)offsetof(cloudy_config, n_modes), (int64_t)offsetof(cloudy_config, P), (int64_t)offsetof(cloudy_config, kind),
        (int64_t)offsetof(cloudy_config, nprog), (int64_t)offsetof(cloudy_config, threshold_style), (int64_t)offsetof(cloudy_config, n_mom_max),
        (int64_t)offsetof(cloudy_config, n_2d_ints), (int64_t)offsetof(cloudy_config, n_bins), (int64_t)offsetof(cloudy_config, n_vel),
        (int64_t)offsetof(cloudy_config, nz), (int64_t)offsetof(cloudy_config, bins_per_log_unit), (int64_t)offsetof(cloudy_config, reserved),
        (int64_t)offsetof(cloudy_config, c), (int64_t)offsetof(cloudy_config, thresholds), (int64_t)offsetof(cloudy_config, x_min),
        (int64_t)offsetof(cloudy_config, dx), (int64_t)offsetof(cloudy_config, norms), (int64_t)offsetof(cloudy_config, k_range),
        (int64_t)offsetof(cloudy_config, vel), (int64_t)offsetof(cloudy_config, dz)};
    const int32_t n = (int32_t)(sizeof(off) / sizeof(off[0]));
    if (!offsets) return n;
    for (int32_t i = 0; i < n && i < max_fields; ++i) offsets[i] = off[i];
    return n < max_fields ? n : max_fields;
}

int cloudy_measure_fp64_peak(cloudy_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(CLOUDY_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 1024, 0.999999, 1e-7);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-7);
        CUDA_TRY(cudaEventRecord(e1, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 16.0 * (double)iters * blocks * threads;
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    ctx->launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return CLOUDY_OK;
}

}  // extern "C"
