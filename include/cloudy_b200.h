/*
 * cloudy_b200.h — C ABI of libcloudy_b200.so
 *
 * B200-native (sm_100a) batched evaluation of Cloudy.jl's collision-coalescence /
 * sedimentation moment tendencies.  Every entry point is `extern "C"`, takes plain
 * pointers and sizes, returns an int status (0 = ok, <0 = error; text via
 * cloudy_last_error()) and never throws.  There is no CPU fallback: every compute
 * entry point launches CUDA kernels and fails with CLOUDY_ERR_CUDA without a device.
 *
 * The reference (CliMA/Cloudy.jl v0.6.0, pure Julia) has no FFI boundary; each symbol
 * below names the Julia method(s) it backs (paths relative to the reference root).
 * A Julia host binds these with `ccall((:sym, libcloudy_b200), Cint, (...), ...)`
 * (see INTEGRATION.md); the in-repo host mirror binds the same symbols with ctypes.
 *
 * Conventions
 *  - all floating point data is IEEE binary64;
 *  - "state" = prognostic moments of every mode of every parcel/cell, on the device, structure
 *    of arrays: slot-major, parcel-minor (slot s of parcel p at base[s*stride + p]); the slot order
 *    is the reference's flat moment vector (mode-major, order-minor; src/helper_functions.jl:13-33);
 *  - host buffers are array-of-structures, parcel-major: host[p*n_slots + s] (one reference
 *    moment vector after another), unless stated otherwise;
 *  - a context is bound to one CUDA device and one stream and is not re-entrant;
 *  - compute calls enqueue on the context's stream and return; cloudy_sync() blocks;
 *  - limits: Gamma shape parameters in (0, 11] (the reference clamps k to (eps, 10], ParticleDistributions.jl:459); an ensemble
 *    buffer (stride x n_slots) must stay below 2^32 doubles per device (the kernels form 32-bit element offsets): calls that
 *    exceed either return CLOUDY_ERR_UNSUPPORTED.
 */
#ifndef CLOUDY_B200_H
#define CLOUDY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLOUDY_MAX_MODES 4   /* N: largest example uses 4 (test/examples/Analytical/box_gamma_mix_moving.jl) */
#define CLOUDY_MAX_P 5       /* P = polynomial order + 1; largest example: order 4 (box_gamma_mixture_hydro.jl:23) */
#define CLOUDY_MAX_VEL 4     /* terms of the power-law terminal velocity (src/Sources/Sedimentation.jl:17-18) */
#define CLOUDY_MAX_SLOTS 12  /* Σ nparams over modes */
#define CLOUDY_MAX_NODES 512 /* nodes of the log-spaced Simpson grid (ParticleDistributions.jl:579-582) */

/* distribution kinds: src/ParticleDistributions/ParticleDistributions.jl:66-159 */
enum { CLOUDY_EXPONENTIAL = 0, CLOUDY_GAMMA = 1, CLOUDY_LOGNORMAL = 2, CLOUDY_MONODISPERSE = 3 };
/* threshold styles: src/Sources/EquationTypes.jl:20-22 */
enum { CLOUDY_FIXED_THRESHOLD = 0, CLOUDY_MOVING_THRESHOLD = 1 };
/* models for the fused time stepper */
enum { CLOUDY_MODEL_BOX = 0, CLOUDY_MODEL_RAINSHAFT = 1 };

enum {
    CLOUDY_OK = 0,
    CLOUDY_ERR_ARG = -1,      /* invalid argument / configuration (the reference would throw) */
    CLOUDY_ERR_CUDA = -2,     /* CUDA runtime error or no device */
    CLOUDY_ERR_STATE = -3,    /* call sequence error (e.g. no configuration set) */
    CLOUDY_ERR_UNSUPPORTED = -4
};

typedef struct cloudy_ctx cloudy_ctx;     /* opaque */
typedef struct cloudy_state cloudy_state; /* opaque: device SoA moments of n parcels */

/*
 * Run-constant configuration = CoalescenceData{N,P,FT} (src/Sources/Coalescence.jl:45-106) plus the
 * fields of the drivers' ODE_parameters tuple (test/examples/Analytical/box_gamma_mixture.jl:30-36,
 * rainshaft_gamma_mixture.jl:39-47).  Everything here is computed by the host wrapper exactly as the
 * reference's constructors do (tensor normalisation KernelTensors.jl:189-199, N_mom_max and N_2d_ints
 * Coalescence.jl:69-76, threshold normalisation :78-84, the node grid ParticleDistributions.jl:579-582
 * with the host language's own log10/log).
 */
typedef struct cloudy_config {
    int32_t n_modes;                         /* N */
    int32_t P;                               /* tensor size (order + 1) */
    int32_t kind[CLOUDY_MAX_MODES];          /* CLOUDY_EXPONENTIAL ... */
    int32_t nprog[CLOUDY_MAX_MODES];         /* NProgMoms = nparams(dist) : 2 or 3 */
    int32_t threshold_style;                 /* CLOUDY_FIXED_THRESHOLD | CLOUDY_MOVING_THRESHOLD */
    int32_t n_mom_max;                       /* CoalescenceData.N_mom_max */
    int32_t n_2d_ints[CLOUDY_MAX_MODES];     /* CoalescenceData.N_2d_ints */
    int32_t n_bins[CLOUDY_MAX_MODES];        /* grid of mode i (0 when its threshold is Inf / last mode) */
    int32_t n_vel;                           /* number of terminal-velocity terms (0 = no sedimentation) */
    int32_t nz;                              /* levels per column (rainshaft); 1 for the box model */
    int32_t bins_per_log_unit;               /* 15 in the reference (ParticleDistributions.jl:572); used by MovingThreshold */
    int32_t reserved;
    double c[CLOUDY_MAX_MODES][CLOUDY_MAX_MODES][CLOUDY_MAX_P][CLOUDY_MAX_P]; /* c[j][k][a][b], already normalised */
    double thresholds[CLOUDY_MAX_MODES];     /* normalised mass thresholds (+Inf allowed) or percentiles (moving) */
    double x_min[CLOUDY_MAX_MODES];          /* log(x_lowerbound) of mode i's grid */
    double dx[CLOUDY_MAX_MODES];             /* log spacing of mode i's grid */
    double norms[2];                         /* (number scale, mass scale) — helper_functions.jl:40-53 */
    double k_range[2];                       /* Gamma shape clamp, reference default (eps, 10) — ParticleDistributions.jl:459 */
    double vel[CLOUDY_MAX_VEL][2];           /* (v_k, beta_k), NOT normalised (rainshaft_helpers.jl:74-76 normalises) */
    double dz;                               /* level thickness (rainshaft_helpers.jl:84) */
} cloudy_config;

/* ---- context -------------------------------------------------------------------------------- */
/* stream: a cudaStream_t to enqueue on (e.g. the caller's current stream), or NULL for a private one */
int cloudy_ctx_create(int device, void* stream, cloudy_ctx** out);
int cloudy_ctx_destroy(cloudy_ctx* ctx);
/* CoalescenceData(...) + ODE_parameters → device constant memory.  Coalescence.jl:55-104 */
int cloudy_config_set(cloudy_ctx* ctx, const cloudy_config* cfg);
int cloudy_sync(cloudy_ctx* ctx);
const char* cloudy_last_error(void);
/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches) */
int cloudy_launch_count(cloudy_ctx* ctx, int64_t* out);
/* execution-shape knob: 0 = auto (thread-per-parcel kernel when the (n_modes, P) shape has an instance, else 8 lanes),
 * 1 = thread per parcel, 4/8/16/32 = that many lanes cooperating on one parcel's quadrature nodes */
int cloudy_set_lanes(cloudy_ctx* ctx, int lanes);
/* regime sort: order the parcels by series length and series / continued-fraction regime so that warps are homogeneous
 * (see "Regime order" below: box-model ensembles are moved physically and stay sorted; column states and the host-buffer
 * pipeline walk a permutation instead).  Results are bit-identical either way.
 * on = 0 off, 1 always, 2 auto (default: from 262144 parcels, where it pays). */
int cloudy_set_regime_sort(cloudy_ctx* ctx, int on);

/* fused steps between two refreshes of a resident ensemble's regime order (default 10) */
int cloudy_set_resort_interval(cloudy_ctx* ctx, int32_t steps);
/* number of data sorts this context has performed (diagnostic; bench.py reports it) */
int cloudy_sort_count(cloudy_ctx* ctx, int64_t* out);

/* ---- device state ----------------------------------------------------------------------------
 * Regime order.  Parcels are independent, so their POSITION in the device arrays is free.  For large box-model
 * ensembles (regime sort mode on/auto) the library moves the parcels into "regime order" — parcels that need the same
 * series length / continued-fraction regime sit next to each other, so that warps are homogeneous — and keeps them
 * there: cloudy_coal_tendency sorts an unsorted input state once (its output inherits the order),
 * cloudy_ssprk33_steps refreshes the order every cloudy_set_resort_interval() steps, cloudy_state_download undoes it,
 * cloudy_state_upload resets it.  Results never depend on the order (bit-identical).  Only callers that read the raw
 * device buffer (cloudy_state_device_ptr) see it: cloudy_state_order() returns the original parcel index of every
 * position.  Column (rainshaft) states are never reordered. */
int cloudy_state_create(cloudy_ctx* ctx, int64_t n_parcels, cloudy_state** out);
int cloudy_state_destroy(cloudy_state* st);
/* host AoS [n_parcels][n_slots] → device SoA, and back; async on the ctx stream when host is pinned */
int cloudy_state_upload(cloudy_ctx* ctx, cloudy_state* st, const double* host, int64_t n_parcels);
int cloudy_state_download(cloudy_ctx* ctx, const cloudy_state* st, double* host, int64_t n_parcels);
/* raw device pointer / stride (in doubles) of the SoA buffer, for zero-copy interop */
int cloudy_state_device_ptr(const cloudy_state* st, double** dptr, int64_t* stride, int32_t* n_slots);
int cloudy_state_copy(cloudy_ctx* ctx, const cloudy_state* src, cloudy_state* dst);
/* move the parcels of a box-model state into regime order now (three kernels: keys, per-bin prefix, stable scatter of all
 * slots; deterministic).  Logically the state is unchanged. */
int cloudy_state_regime_sort(cloudy_ctx* ctx, cloudy_state* st);
/* *d_order = DEVICE pointer to n int32 (original parcel index of position i), or NULL when position == parcel index */
int cloudy_state_order(const cloudy_state* st, const int32_t** d_order);

/* ---- batched hot path ------------------------------------------------------------------------ */
/* dm = rhs_coal!(AnalyticalCoalStyle, dm, m, p, threshold_style) for every parcel.
 * test/examples/utils/box_model_helpers.jl:29-53 → src/Sources/Coalescence.jl:115-185 */
int cloudy_coal_tendency(cloudy_ctx* ctx, const cloudy_state* m, cloudy_state* dm);
/* flux = get_sedimentation_flux(pdists(m), vel_normalized) .* mom_norms for every cell.
 * src/Sources/Sedimentation.jl:22-37, rainshaft_helpers.jl:74-77 */
int cloudy_sedimentation_flux(cloudy_ctx* ctx, const cloudy_state* m, cloudy_state* flux);
/* dm = make_rainshaft_rhs(AnalyticalCoalStyle())(m, p, t); m is clipped at 0 IN PLACE like the
 * reference (rainshaft_helpers.jl:52).  Cells are column-major: parcel index = column*nz + level. */
int cloudy_rainshaft_rhs(cloudy_ctx* ctx, cloudy_state* m, cloudy_state* dm);
/* n_steps of SSPRK33 (Shu-Osher form used by OrdinaryDiffEqSSPRK; call sites e.g.
 * box_gamma_mixture.jl:38, rainshaft_gamma_mixture.jl:49) with the RHS fused into each stage update.
 * Column model: every stage output, the returned state included, is clipped at zero.  This is the reference's sequence,
 * not a deviation: its right-hand side clips the array it is handed IN PLACE (rainshaft_helpers.jl:52), and
 * OrdinaryDiffEq's SSPRK33 evaluates f(u_{n+1}) at the end of every step (the first-same-as-last derivative) before the
 * state is saved, so every saved state of the reference has been clipped too. */
int cloudy_ssprk33_steps(cloudy_ctx* ctx, cloudy_state* u, double dt, int32_t n_steps, int32_t model);
/* per-slot sums over all parcels of this device (the conservation diagnostic, cf. moments_sum in
 * test/examples/utils/netcdf_helpers.jl:34-42).  d_out: DEVICE pointer to n_slots doubles (so the caller
 * can all-reduce it with NCCL without a host round trip). */
int cloudy_moment_sums_device(cloudy_ctx* ctx, const cloudy_state* u, double* d_out);
int cloudy_moment_sums(cloudy_ctx* ctx, const cloudy_state* u, double* host_out);
/* ---- multi-GPU (one context per GPU, one process per GPU) ----------------------------------------
 * Parcels / whole columns are block-partitioned over the ranks with no halo and no exchange; the path's only collective is
 * the all-reduce of the <= 12 per-slot sums (the conservation diagnostic the reference computes serially: moments_sum,
 * test/examples/utils/netcdf_helpers.jl:34-42).  NCCL is bound at run time (dlopen of libnccl.so.2, or $CLOUDY_NCCL_LIB).
 *   rank 0:   cloudy_comm_unique_id(id)  -> ship the 128 bytes to the other ranks (MPI.jl / a file / torch.distributed)
 *   all:      cloudy_comm_init(ctx, n_ranks, rank, id)                 (collective: ncclCommInitRank)
 *   all:      cloudy_moment_sums_allreduce(ctx, u, out | NULL)         (collective)
 * With host_out == NULL the call only enqueues (local reduction on the context's stream, ncclAllReduce + copy to pinned host
 * memory on a side stream) and the result is collected later with cloudy_moment_sums_fetch, so the next step overlaps the
 * collective.  A context without a communicator (or with n_ranks == 1) returns its local sums. */
int cloudy_comm_unique_id(void* id_out /* 128 bytes */);
int cloudy_comm_init(cloudy_ctx* ctx, int32_t n_ranks, int32_t rank, const void* unique_id /* 128 bytes */);
int cloudy_comm_destroy(cloudy_ctx* ctx);
int cloudy_comm_info(cloudy_ctx* ctx, int32_t* n_ranks, int32_t* rank, int32_t* nccl_version);
int cloudy_moment_sums_allreduce(cloudy_ctx* ctx, const cloudy_state* u, double* host_out);
int cloudy_moment_sums_fetch(cloudy_ctx* ctx, double* host_out);

/* dm = rhs_condensation!(dm, m, p, s): get_cond_evap(pdists(m), s, xi / norms[2]^(2/3), rho_l) .* mom_norms for every parcel.
 * test/examples/utils/box_model_helpers.jl:55-67 → src/Sources/Condensation.jl:22-37.  d_s: optional DEVICE array of
 * per-parcel supersaturations (NULL → the scalar s for all parcels). */
int cloudy_cond_evap(cloudy_ctx* ctx, const cloudy_state* m, double s, const double* d_s, double xi, double rho_l, cloudy_state* dm);
/* (N_liq, N_rai, M_liq, M_rai) = get_standard_N_q(pdists(m), size_cutoff) for every parcel/cell
 * (src/ParticleDistributions/ParticleDistributions.jl:634-687; caller: test/examples/utils/netcdf_helpers.jl:106-121).
 * normalized = 0: distributions are rebuilt from the RAW moments like netcdf_helpers.jl does; 1: from moments / norms.
 * d_out: DEVICE pointer to 4*n doubles, quantity-major ([4][n]). */
int cloudy_standard_N_q(cloudy_ctx* ctx, const cloudy_state* m, double size_cutoff, int32_t normalized, double* d_out);
/* convenience: host moments in, host tendencies out (upload + kernel + download, chunked & overlapped) */
int cloudy_coal_tendency_host(cloudy_ctx* ctx, const double* host_m, double* host_dm, int64_t n_parcels);
/* number of parcels whose moments were invalid for their distribution (Lognormal sqrt(log(.)<0),
 * ParticleDistributions.jl:498 — the reference throws a DomainError) since the last call */
int cloudy_error_count(cloudy_ctx* ctx, int64_t* n_invalid);

/* ---- single-object entry points (each evaluated by a 1-parcel kernel launch) -------------------
 * so that every reference method on the path has a backing symbol.
 * params = (n, θ) | (n, θ, k) | (n, μ, σ) in the reference's field order. */
/* moment(dist, q) — ParticleDistributions.jl:216 */
int cloudy_moment(cloudy_ctx* ctx, int32_t kind, const double* params, double q, double* out);
/* update_dist_from_moments(dist, moments; param_range) — ParticleDistributions.jl:456-541.
 * range = (lo, hi) for k (Gamma) or (mu_lo, mu_hi, sigma_lo, sigma_hi) (Lognormal); NULL = defaults.
 * *invalid = 1 where the reference would throw a DomainError. */
int cloudy_update_dist_from_moments(cloudy_ctx* ctx, int32_t kind, const double* moments, const double* range,
                                    double* params_out, int32_t* invalid);
/* moment_source_helper(dist, p1, p2, x_threshold, n_bins_per_log_unit) — ParticleDistributions.jl:557-625, every distribution
 * kind.  Exponential / Gamma: the reference's log-spaced Simpson rule on the reference's nodes (:567-612); Monodisperse: closed
 * form (:557-564); Lognormal (:614-625, two nested adaptive QuadGK calls at rtol sqrt(eps) in the reference): inner integral
 * in closed form, outer integral by a fixed 128-point Gauss-Legendre rule in ln y (n_bins_per_log_unit is ignored). */
int cloudy_moment_source_helper(cloudy_ctx* ctx, int32_t kind, const double* params, double p1, double p2,
                                double x_threshold, int32_t n_bins_per_log_unit, double* out);
/* compute_threshold(pdist, percentile, minx) — ParticleDistributions.jl:747-761 (Exponential: -θ log(1-p); Gamma: θ gamma_inc_inv(k, p)) */
int cloudy_compute_threshold(cloudy_ctx* ctx, int32_t kind, const double* params, double percentile, double minx, double* out);
/* get_coal_ints(AnalyticalCoalStyle(), pdists, coal_data[, MovingThreshold()]) for ONE set of distributions;
 * params: [n_modes][3]; out: Σ nprog doubles (normalised units, exactly the reference's return value).
 * Given distributions bypass update_dist_from_moments' clamp: a Gamma shape outside (0, 11] is refused (CLOUDY_ERR_UNSUPPORTED).
 * Coalescence.jl:115-185 */
int cloudy_get_coal_ints_1(cloudy_ctx* ctx, const double* params, double* out);
/* get_sedimentation_flux(pdists, vel) for ONE set of distributions, vel given explicitly. Sedimentation.jl:22-37 */
int cloudy_get_sedimentation_flux_1(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params,
                                    int32_t n_vel, const double* vel, double* out);
/* get_cond_evap(pdists, s, xi, rho_l) for ONE set of distributions (kinds/params as above). Condensation.jl:22-37 */
int cloudy_get_cond_evap_1(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params, double s, double xi,
                           double rho_l, double* out);
/* get_standard_N_q(pdists, size_cutoff) for ONE set of distributions; out = (N_liq, N_rai, M_liq, M_rai).
 * ParticleDistributions.jl:634-687 */
int cloudy_get_standard_N_q_1(cloudy_ctx* ctx, int32_t n_modes, const int32_t* kinds, const double* params, double size_cutoff,
                              double* out);
/* integrate_SimpsonEvenFast(n_bins, dx, y) with tabulated y[1..n_bins+1] — ParticleDistributions.jl:698-710 */
int cloudy_integrate_simpson(cloudy_ctx* ctx, int32_t n_bins, double dx, const double* y, double* out);

/* ---- measurement helpers ----------------------------------------------------------------------- */
/* sustained FP64 FMA throughput of this device in TFLOP/s (independent DFMA chains, all SMs) — the
 * roofline denominator that MEASURED_PEAKS.json lacks */
int cloudy_measure_fp64_peak(cloudy_ctx* ctx, double* tflops);

/* ---- binding self-check ------------------------------------------------------------------------ */
/* sizeof(cloudy_config) and the byte offsets of its 20 fields in declaration order, as this library was compiled.  A language
 * binding that mirrors the struct (julia/CloudyB200.jl CloudyConfig, cloudy.jl_b200/_lib.py) compares them with its own
 * layout when it loads the library, so the two cannot drift apart silently.  Returns the number of fields written
 * (at most max_fields). */
int64_t cloudy_config_sizeof(void);
int32_t cloudy_config_offsets(int64_t* offsets, int32_t max_fields);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDY_B200_H */
