# CloudyB200.jl — the reference-side binding of libcloudy_b200.so (C ABI in include/cloudy_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia runtime.  The tested binding of the same
# symbols is the ctypes mirror in cloudy.jl_b200/ (see INTEGRATION.md); tests/test_host_logic.py parses the struct
# below and checks its C layout against the library (cloudy_config_offsets), so the two mirrors cannot drift.
#
# What a Cloudy.jl maintainer gets:
#   * `make_box_model_rhs(coal_type, threshold_style)` and `make_rainshaft_rhs(coal_type)` with the reference's
#     signatures (test/examples/utils/box_model_helpers.jl:22-27, rainshaft_helpers.jl:45-88), evaluating on the GPU;
#     they accept one state (the reference's calling convention) or a whole ensemble of them;
#   * METHODS ON THE REFERENCE'S OWN GENERIC FUNCTIONS for two wrapper types:
#       `B200Batch`  — many parcels at once: `update_dist_from_moments`, `get_coal_ints`, `get_sedimentation_flux`;
#       `OnB200(d)`  — one distribution evaluated by the library: `moment`, `moment_source_helper`, `compute_threshold`,
#                      `update_dist_from_moments`, and tuples of them in `get_coal_ints`, `get_sedimentation_flux`,
#                      `get_standard_N_q`, `get_cond_evap`;
#   * device-resident ensembles (`Ensemble`, `ssprk33!`, `moment_sums`, `moment_sums_allreduce` over NCCL).
module CloudyB200

using Cloudy
using Cloudy.ParticleDistributions
using Cloudy.KernelTensors
using Cloudy.EquationTypes
import Cloudy.Coalescence
import Cloudy.Coalescence: CoalescenceData, get_coal_ints
import Cloudy.Sedimentation: get_sedimentation_flux
import Cloudy.Condensation: get_cond_evap
import Cloudy.ParticleDistributions:
    update_dist_from_moments, moment, moment_source_helper, compute_threshold, get_standard_N_q, nparams

export Context, B200Batch, OnB200, Ensemble, make_box_model_rhs, make_rainshaft_rhs, upload!, download!, ssprk33!,
    moment_sums, comm_unique_id, comm_init!, moment_sums_allreduce

const lib = get(ENV, "LIBCLOUDY_B200", "libcloudy_b200.so")

const MAX_MODES, MAX_P, MAX_VEL = 4, 5, 4
const MODEL_BOX, MODEL_RAINSHAFT = Int32(0), Int32(1)

# mirrors `cloudy_config` field by field (isbits, C layout) — include/cloudy_b200.h
struct CloudyConfig
    n_modes::Int32
    P::Int32
    kind::NTuple{MAX_MODES,Int32}
    nprog::NTuple{MAX_MODES,Int32}
    threshold_style::Int32
    n_mom_max::Int32
    n_2d_ints::NTuple{MAX_MODES,Int32}
    n_bins::NTuple{MAX_MODES,Int32}
    n_vel::Int32
    nz::Int32
    bins_per_log_unit::Int32
    reserved::Int32
    c::NTuple{MAX_MODES * MAX_MODES * MAX_P * MAX_P,Float64}   # c[j][k][a][b], row-major
    thresholds::NTuple{MAX_MODES,Float64}
    x_min::NTuple{MAX_MODES,Float64}
    dx::NTuple{MAX_MODES,Float64}
    norms::NTuple{2,Float64}
    k_range::NTuple{2,Float64}
    vel::NTuple{2 * MAX_VEL,Float64}
    dz::Float64
end

check(rc) = rc == 0 || error(unsafe_string(ccall((:cloudy_last_error, lib), Cstring, ())))

"The library reports sizeof(cloudy_config) and its field offsets as compiled: refuse to run with a mirror that has drifted."
function __init__()
    nf = fieldcount(CloudyConfig)
    offs = zeros(Int64, 64)
    n = ccall((:cloudy_config_offsets, lib), Int32, (Ptr{Int64}, Int32), offs, 64)
    sz = ccall((:cloudy_config_sizeof, lib), Int64, ())
    mine = [Int64(fieldoffset(CloudyConfig, i)) for i in 1:nf]
    (sz == sizeof(CloudyConfig) && n == nf && offs[1:n] == mine) ||
        error("cloudy_config layout mismatch: library sizeof $sz offsets $(offs[1:n]), Julia mirror sizeof $(sizeof(CloudyConfig)) offsets $mine")
end

kind_code(::ExponentialPrimitiveParticleDistribution) = Int32(0)
kind_code(::GammaPrimitiveParticleDistribution) = Int32(1)
kind_code(::LognormalPrimitiveParticleDistribution) = Int32(2)
kind_code(::MonodispersePrimitiveParticleDistribution) = Int32(3)

pad(t, n, z) = ntuple(i -> i <= length(t) ? t[i] : z, n)

"""
    CloudyConfig(pdists, coal_data, norms; vel = (), dz = 1.0, nz = 1, threshold_style = FixedThreshold(),
                 n_bins_per_log_unit = 15, k_range = (eps(Float64), 10.0))

Build the C struct from the reference's own objects.  The node grid of every thresholded Gamma/Exponential mode is computed
HERE, with Julia's own `log10`/`log`, exactly as ParticleDistributions.jl:579-582 does (`floor(15 log10(x_th/x_lb))` is one
ulp away from 74 vs 75 nodes, so the library never recomputes it).  `coal_data` must have been built with the same
`threshold_style` (Coalescence.jl:78-84 normalises mass thresholds, keeps percentiles).
"""
function CloudyConfig(pdists::NTuple{N,Any}, cd::CoalescenceData{N,P,FT}, norms; vel = (), dz = 1.0, nz = 1,
                      threshold_style::ThresholdStyle = FixedThreshold(), n_bins_per_log_unit::Integer = 15,
                      k_range = (eps(Float64), 10.0)) where {N,P,FT}
    moving = threshold_style isa MovingThreshold
    c = zeros(Float64, MAX_P, MAX_P, MAX_MODES, MAX_MODES)          # Julia column-major c[b,a,k,j] == C c[j][k][a][b]
    for j in 1:N, k in 1:N, a in 1:P, b in 1:P
        c[b, a, k, j] = cd.kernels[j][k].c[a, b]
    end
    n_bins = zeros(Int32, MAX_MODES); x_min = zeros(MAX_MODES); dxs = zeros(MAX_MODES)
    for i in 1:(N-1)
        t = cd.dist_thresholds[i]
        if !moving && isfinite(t) && kind_code(pdists[i]) in (0, 1)
            x_lb = min(1e-5, 1e-5 * t)
            nb = floor(Int, n_bins_per_log_unit * log10(t / x_lb))
            n_bins[i] = nb; x_min[i] = log(x_lb); dxs[i] = (log(t) - log(x_lb)) / nb
        end
    end
    v = zeros(2 * MAX_VEL)
    for (i, (a, b)) in enumerate(vel)
        v[2i-1] = a; v[2i] = b
    end
    CloudyConfig(N, P, pad(map(kind_code, pdists), MAX_MODES, Int32(0)), pad(map(d -> Int32(nparams(d)), pdists), MAX_MODES, Int32(0)),
                 moving ? 1 : 0, cd.N_mom_max, pad(map(Int32, cd.N_2d_ints), MAX_MODES, Int32(0)), Tuple(n_bins), length(vel), nz,
                 n_bins_per_log_unit, 0, Tuple(c), pad(map(Float64, cd.dist_thresholds), MAX_MODES, 0.0), Tuple(x_min), Tuple(dxs),
                 (Float64(norms[1]), Float64(norms[2])), (Float64(k_range[1]), Float64(k_range[2])), Tuple(v), Float64(dz))
end

mutable struct Context
    handle::Ptr{Cvoid}
    config::Any          # last CloudyConfig applied (skips re-configuration when the same one is asked for)
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cloudy_ctx_create, lib), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, h))
        x = new(h[], nothing)
        finalizer(c -> ccall((:cloudy_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.handle), x)
    end
end

const default_ctx = Ref{Union{Nothing,Context}}(nothing)
default_context() = (default_ctx[] === nothing && (default_ctx[] = Context()); default_ctx[])

function set_config!(ctx::Context, cfg::CloudyConfig)
    ctx.config === cfg && return ctx
    ctx.config = nothing   # a failed call leaves the context's previous configuration in the library, not in this cache
    check(ccall((:cloudy_config_set, lib), Cint, (Ptr{Cvoid}, Ref{CloudyConfig}), ctx.handle, cfg))
    ctx.config = cfg
    ctx
end

# ---------------------------------------------------------------------------------------------------------------------
# ODE right-hand sides with the reference's signatures
# ---------------------------------------------------------------------------------------------------------------------
"""
    rhs_coal_batched!(dm, m, ctx)

Batched `rhs_coal!` (test/examples/utils/box_model_helpers.jl:29-53): `m`, `dm` are `n_moments × n_parcels`
matrices (one reference moment vector per column == the library's host layout).
"""
function rhs_coal_batched!(dm::Matrix{Float64}, m::Matrix{Float64}, ctx::Context)
    check(ccall((:cloudy_coal_tendency_host, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                ctx.handle, m, dm, size(m, 2)))
    return dm
end

"""
    make_box_model_rhs(coal_type, threshold_style = FixedThreshold(); ctx = default_context())

Drop-in for box_model_helpers.jl:22-27: returns `rhs!(dm, m, par, t)` with `par = (; pdists, coal_data, NProgMoms, norms, dt)`.
`m` is one moment vector (the reference's calling convention) or an `n_moments × n_parcels` matrix.  `MovingThreshold()`
selects the percentile-threshold method (Coalescence.jl:152-185); `par.coal_data` must have been built with it.
"""
function make_box_model_rhs(coal_type::AnalyticalCoalStyle, threshold_style::ThresholdStyle = FixedThreshold();
                            ctx::Context = default_context())
    cfg = Ref{Any}(nothing)
    function rhs!(dm, m, par, t)
        if cfg[] === nothing
            cfg[] = CloudyConfig(par.pdists, par.coal_data, par.norms; threshold_style = threshold_style)
        end
        set_config!(ctx, cfg[])
        mm = reshape(collect(Float64, m), sum(par.NProgMoms), :)
        out = similar(mm)
        rhs_coal_batched!(out, mm, ctx)
        dm .= reshape(out, size(dm))
    end
end

"""
    make_rainshaft_rhs(coal_type; ctx = default_context())

Drop-in for rainshaft_helpers.jl:45-88: returns the out-of-place `rhs(m, p, t)` with `p = (; pdists, coal_data, NProgMoms, norms,
vel, dz, dt)`.  `m` is `nz × nmom` (one column, the reference's layout) or `nz × nmom × n_columns`; negative entries of `m`
are clipped IN PLACE like the reference does (:52).
"""
function make_rainshaft_rhs(coal_type::AnalyticalCoalStyle; ctx::Context = default_context())
    cfg = Ref{Any}(nothing)
    ens = Ref{Any}(nothing)
    function rhs(m, p, t)
        nz, nmom = size(m, 1), size(m, 2)
        ncol = ndims(m) == 3 ? size(m, 3) : 1
        if cfg[] === nothing
            cfg[] = CloudyConfig(p.pdists, p.coal_data, p.norms; vel = p.vel, dz = p.dz, nz = nz)
        end
        set_config!(ctx, cfg[])
        if ens[] === nothing || ens[][1].n != nz * ncol
            ens[] = (Ensemble(ctx, nz * ncol), Ensemble(ctx, nz * ncol))
        end
        u, du = ens[]
        # library layout: one moment vector per cell, cells of a column contiguous (cell = column * nz + level)
        host = permutedims(reshape(collect(Float64, m), nz, nmom, ncol), (2, 1, 3))
        upload!(u, reshape(host, nmom, nz * ncol))
        check(ccall((:cloudy_rainshaft_rhs, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, u.handle, du.handle))
        out = Matrix{Float64}(undef, nmom, nz * ncol)
        download!(out, du)
        m[m .< 0] .= 0                                  # the device state was clipped; keep the caller's array in step
        dm = permutedims(reshape(out, nmom, nz, ncol), (2, 1, 3))
        return ndims(m) == 3 ? dm : reshape(dm, nz, nmom)
    end
end

# ---------------------------------------------------------------------------------------------------------------------
# B200Batch: many parcels behind the reference's generic functions
# ---------------------------------------------------------------------------------------------------------------------
"""
    B200Batch(pdists, moments; ctx = default_context())

`pdists`: one template distribution per mode (their types select the kinds); `moments`: `sum(nparams) × n_parcels` NORMALISED
prognostic moments, one reference moment vector per column.  The per-parcel `update_dist_from_moments` of the reference's
right-hand sides happens on the device inside `get_coal_ints` / `get_sedimentation_flux`.
"""
struct B200Batch{N,D}
    ctx::Context
    pdists::D
    moments::Matrix{Float64}
    B200Batch(pdists::NTuple{N,Any}, moments::AbstractMatrix; ctx::Context = default_context()) where {N} =
        new{N,typeof(pdists)}(ctx, pdists, Matrix{Float64}(moments))
end

"`update_dist_from_moments` for every parcel of the batch: the new moments are adopted (parameters are rebuilt on the device)."
update_dist_from_moments(b::B200Batch{N}, moments::AbstractMatrix; param_range = nothing) where {N} =
    B200Batch(b.pdists, moments; ctx = b.ctx)

"`get_coal_ints(AnalyticalCoalStyle(), batch, coal_data)` → `sum(nparams) × n_parcels` matrix (Coalescence.jl:115-150)."
function get_coal_ints(cs::AnalyticalCoalStyle, b::B200Batch{N}, coal_data::CoalescenceData{N,P,FT},
                       ts::ThresholdStyle = FixedThreshold()) where {N,P,FT}
    set_config!(b.ctx, CloudyConfig(b.pdists, coal_data, (1.0, 1.0); threshold_style = ts))
    rhs_coal_batched!(similar(b.moments), b.moments, b.ctx)
end

"`get_sedimentation_flux(batch, vel)` → `sum(nparams) × n_parcels` matrix (Sedimentation.jl:22-37)."
function get_sedimentation_flux(b::B200Batch{N}, vel::NTuple{M,Tuple{FT,FT}}) where {N,M,FT}
    # unit tensor/thresholds: only the distributions and the velocity terms enter the flux
    NProg = map(nparams, b.pdists)
    cd = CoalescenceData(CoalescenceTensor(fill(0.0, 1, 1)), NProg, ntuple(i -> Inf, N), (1.0, 1.0))
    set_config!(b.ctx, CloudyConfig(b.pdists, cd, (1.0, 1.0); vel = vel))
    n = size(b.moments, 2)
    u, fl = Ensemble(b.ctx, n), Ensemble(b.ctx, n)
    upload!(u, b.moments)
    check(ccall((:cloudy_sedimentation_flux, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), b.ctx.handle, u.handle, fl.handle))
    out = similar(b.moments)
    download!(out, fl)
    out
end

# ---------------------------------------------------------------------------------------------------------------------
# OnB200: one distribution, evaluated by the library, behind the reference's generic functions
# ---------------------------------------------------------------------------------------------------------------------
struct OnB200{D}
    dist::D
    ctx::Context
    OnB200(d; ctx::Context = default_context()) = new{typeof(d)}(d, ctx)
end
nparams(d::OnB200) = nparams(d.dist)
kind_code(d::OnB200) = kind_code(d.dist)
params3(d) = (p = collect(Float64, ntuple(i -> getfield(d, i), nparams(d))); length(p) < 3 && push!(p, 1.0); p)
params3(d::OnB200) = params3(d.dist)

"`moment(dist, q)` — ParticleDistributions.jl:216"
function moment(d::OnB200, q)
    out = Ref{Float64}(0)
    check(ccall((:cloudy_moment, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cdouble, Ref{Float64}),
                d.ctx.handle, kind_code(d), params3(d), Float64(q), out))
    out[]
end

"`moment_source_helper(dist, p1, p2, x_threshold, n_bins_per_log_unit = 15)` — ParticleDistributions.jl:557-625, every kind"
function moment_source_helper(d::OnB200, p1, p2, x_threshold, n_bins_per_log_unit = 15)
    out = Ref{Float64}(0)
    check(ccall((:cloudy_moment_source_helper, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cdouble, Cdouble, Cdouble, Int32, Ref{Float64}),
                d.ctx.handle, kind_code(d), params3(d), Float64(p1), Float64(p2), Float64(x_threshold), n_bins_per_log_unit, out))
    out[]
end

"`compute_threshold(pdist, percentile, minx)` — ParticleDistributions.jl:747-761"
function compute_threshold(d::OnB200, percentile = 0.97, minx = 1e-18)
    out = Ref{Float64}(0)
    check(ccall((:cloudy_compute_threshold, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cdouble, Cdouble, Ref{Float64}),
                d.ctx.handle, kind_code(d), params3(d), Float64(percentile), Float64(minx), out))
    out[]
end

"`update_dist_from_moments(pdist, moments; param_range)` — ParticleDistributions.jl:456-541; returns a distribution of the wrapped type"
function update_dist_from_moments(d::OnB200, moments::Tuple; param_range = nothing)
    m = collect(Float64, pad(moments, 3, 0.0))
    rng = param_range === nothing ? C_NULL :
          (d.dist isa GammaPrimitiveParticleDistribution ? Float64[param_range.k[1], param_range.k[2]] :
           Float64[param_range.μ[1], param_range.μ[2], param_range.σ[1], param_range.σ[2]])
    out = zeros(Float64, 3)
    invalid = Ref{Int32}(0)
    check(ccall((:cloudy_update_dist_from_moments, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                d.ctx.handle, kind_code(d), m, rng, out, invalid))
    invalid[] != 0 && throw(DomainError(moments, "log of a negative number in update_dist_from_moments (ParticleDistributions.jl:498)"))
    OnB200(typeof(d.dist)(out[1:nparams(d)]...); ctx = d.ctx)
end

function params_matrix(pdists)
    params = zeros(Float64, 3, length(pdists))
    for (i, d) in enumerate(pdists)
        params[:, i] .= params3(d)
    end
    params
end

"`get_coal_ints(AnalyticalCoalStyle(), pdists, coal_data[, MovingThreshold()])` for one tuple of wrapped distributions — Coalescence.jl:115-185"
function get_coal_ints(cs::AnalyticalCoalStyle, pdists::NTuple{N,OnB200}, coal_data::CoalescenceData{N,P,FT},
                       ts::ThresholdStyle = FixedThreshold()) where {N,P,FT}
    ctx = pdists[1].ctx
    set_config!(ctx, CloudyConfig(map(d -> d.dist, pdists), coal_data, (1.0, 1.0); threshold_style = ts))
    out = zeros(Float64, sum(nparams, pdists))
    check(ccall((:cloudy_get_coal_ints_1, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.handle, params_matrix(pdists), out))
    Tuple(out)
end

"`get_sedimentation_flux(pdists, vel)` for one tuple of wrapped distributions — Sedimentation.jl:22-37"
function get_sedimentation_flux(pdists::NTuple{N,OnB200}, vel::NTuple{M,Tuple{FT,FT}}) where {N,M,FT}
    kinds = Int32[kind_code(d) for d in pdists]
    v = zeros(Float64, 2, M)
    for (i, (a, b)) in enumerate(vel)
        v[1, i] = a; v[2, i] = b
    end
    out = zeros(Float64, sum(nparams, pdists))
    check(ccall((:cloudy_get_sedimentation_flux_1, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}),
                pdists[1].ctx.handle, N, kinds, params_matrix(pdists), M, v, out))
    Tuple(out)
end

"`get_cond_evap(pdists, s, ξ)` for one tuple of wrapped distributions — Condensation.jl:22-37"
function get_cond_evap(pdists::NTuple{N,OnB200}, s, ξ; ρ_l = 1000.0) where {N}
    kinds = Int32[kind_code(d) for d in pdists]
    out = zeros(Float64, sum(nparams, pdists))
    check(ccall((:cloudy_get_cond_evap_1, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Cdouble, Cdouble, Cdouble, Ptr{Float64}),
                pdists[1].ctx.handle, N, kinds, params_matrix(pdists), Float64(s), Float64(ξ), Float64(ρ_l), out))
    Tuple(out)
end

"`get_standard_N_q(pdists, size_cutoff)` → (; N_liq, N_rai, M_liq, M_rai) — ParticleDistributions.jl:634-687"
function get_standard_N_q(pdists::NTuple{N,OnB200}, size_cutoff = 1e-6) where {N}
    kinds = Int32[kind_code(d) for d in pdists]
    out = zeros(Float64, 4)
    check(ccall((:cloudy_get_standard_N_q_1, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Cdouble, Ptr{Float64}),
                pdists[1].ctx.handle, N, kinds, params_matrix(pdists), Float64(size_cutoff), out))
    (; N_liq = out[1], N_rai = out[2], M_liq = out[3], M_rai = out[4])
end

# ---------------------------------------------------------------------------------------------------------------------
# device-resident ensembles for long runs
# ---------------------------------------------------------------------------------------------------------------------
mutable struct Ensemble
    ctx::Context
    handle::Ptr{Cvoid}
    n::Int
    function Ensemble(ctx::Context, n::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cloudy_state_create, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx.handle, n, h))
        e = new(ctx, h[], n)
        finalizer(x -> ccall((:cloudy_state_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), e)
    end
end
upload!(e::Ensemble, m::Matrix{Float64}) =
    check(ccall((:cloudy_state_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), e.ctx.handle, e.handle, m, e.n))
download!(m::Matrix{Float64}, e::Ensemble) =
    check(ccall((:cloudy_state_download, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), e.ctx.handle, e.handle, m, e.n))
"`solve(prob, SSPRK33(), dt = dt)` for `n_steps` steps, state kept on the GPU (`model` = MODEL_BOX | MODEL_RAINSHAFT)."
ssprk33!(e::Ensemble, dt, n_steps; model = MODEL_BOX) =
    check(ccall((:cloudy_ssprk33_steps, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int32, Int32), e.ctx.handle, e.handle, dt, n_steps, model))
"`rhs_condensation!` for a device ensemble (test/examples/utils/box_model_helpers.jl:55-67); `s` scalar supersaturation."
cond_evap!(dm::Ensemble, m::Ensemble, s, ξ; ρ_l = 1000.0) =
    check(ccall((:cloudy_cond_evap, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cvoid}),
                m.ctx.handle, m.handle, s, C_NULL, ξ, ρ_l, dm.handle))

"Σ over this device's parcels of every prognostic moment (conservation diagnostic, cf. moments_sum in netcdf_helpers.jl:34-42)."
function moment_sums(e::Ensemble, n_slots::Integer)
    out = zeros(Float64, n_slots)
    check(ccall((:cloudy_moment_sums, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), e.ctx.handle, e.handle, out))
    out
end

# multi-GPU: one Julia process per GPU; the 128-byte id travels by MPI.jl / a shared file (rank 0 draws it)
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:cloudy_comm_unique_id, lib), Cint, (Ptr{UInt8},), id))
    id
end
comm_init!(ctx::Context, n_ranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:cloudy_comm_init, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ctx.handle, n_ranks, rank, id))
"Σ over the parcels of ALL ranks (device reduction + ncclAllReduce inside the library); every rank receives the same sums."
function moment_sums_allreduce(e::Ensemble, n_slots::Integer)
    out = zeros(Float64, n_slots)
    check(ccall((:cloudy_moment_sums_allreduce, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), e.ctx.handle, e.handle, out))
    out
end

end # module
