# CloudyB200.jl — the reference-side binding of libcloudy_b200.so (C ABI in include/cloudy_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia runtime.  The tested binding of the same
# symbols is the ctypes mirror in cloudy.jl_b200/ (see INTEGRATION.md).  This module shows what a Cloudy.jl
# maintainer adds so that the reference's own drivers (test/examples/Analytical/*.jl) run on the GPU library
# for the analytical coalescence path, batched over parcels.
module CloudyB200

using Cloudy
using Cloudy.ParticleDistributions
using Cloudy.KernelTensors
using Cloudy.Coalescence
using Cloudy.EquationTypes

const lib = get(ENV, "LIBCLOUDY_B200", "libcloudy_b200.so")

const MAX_MODES, MAX_P, MAX_VEL = 4, 5, 4

# mirrors `cloudy_config` field by field (isbits, C layout)
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

kind_code(::ExponentialPrimitiveParticleDistribution) = Int32(0)
kind_code(::GammaPrimitiveParticleDistribution) = Int32(1)
kind_code(::LognormalPrimitiveParticleDistribution) = Int32(2)
kind_code(::MonodispersePrimitiveParticleDistribution) = Int32(3)

pad(t, n, z) = ntuple(i -> i <= length(t) ? t[i] : z, n)

check(rc) = rc == 0 || error(unsafe_string(ccall((:cloudy_last_error, lib), Cstring, ())))

"Build the C struct from the reference's own objects (the grid uses Julia's log10/log, like ParticleDistributions.jl:579-582)."
function CloudyConfig(pdists::NTuple{N}, cd::CoalescenceData{N,P,FT}, norms; vel = (), dz = 1.0, nz = 1,
                      moving = false) where {N,P,FT}
    c = zeros(Float64, MAX_P, MAX_P, MAX_MODES, MAX_MODES)          # Julia column-major == C c[j][k][a][b]
    for j in 1:N, k in 1:N, a in 1:P, b in 1:P
        c[b, a, k, j] = cd.kernels[j][k].c[a, b]
    end
    n_bins = zeros(Int32, MAX_MODES); x_min = zeros(MAX_MODES); dxs = zeros(MAX_MODES)
    for i in 1:(N-1)
        t = cd.dist_thresholds[i]
        if !moving && isfinite(t) && kind_code(pdists[i]) in (0, 1)
            x_lb = min(1e-5, 1e-5 * t)
            nb = floor(Int, 15 * log10(t / x_lb))
            n_bins[i] = nb; x_min[i] = log(x_lb); dxs[i] = (log(t) - log(x_lb)) / nb
        end
    end
    v = zeros(2 * MAX_VEL)
    for (i, (a, b)) in enumerate(vel)
        v[2i-1] = a; v[2i] = b
    end
    CloudyConfig(N, P, pad(map(kind_code, pdists), MAX_MODES, Int32(0)), pad(map(d -> Int32(nparams(d)), pdists), MAX_MODES, Int32(0)),
                 moving ? 1 : 0, cd.N_mom_max, pad(map(Int32, cd.N_2d_ints), MAX_MODES, Int32(0)), Tuple(n_bins), length(vel), nz, 15, 0,
                 Tuple(c), pad(cd.dist_thresholds, MAX_MODES, 0.0), Tuple(x_min), Tuple(dxs), (norms[1], norms[2]),
                 (eps(Float64), 10.0), Tuple(v), dz)
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cloudy_ctx_create, lib), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, h))
        x = new(h[])
        finalizer(c -> ccall((:cloudy_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.handle), x)
    end
end

set_config!(ctx::Context, cfg::CloudyConfig) =
    check(ccall((:cloudy_config_set, lib), Cint, (Ptr{Cvoid}, Ref{CloudyConfig}), ctx.handle, cfg))

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

"Drop-in for `make_box_model_rhs(AnalyticalCoalStyle())`: `rhs!(dm, m, par, t)` on one moment vector or a matrix of them."
function make_box_model_rhs(::AnalyticalCoalStyle, ctx::Context = Context())
    configured = Ref(false)
    function rhs!(dm, m, par, t)
        if !configured[]
            set_config!(ctx, CloudyConfig(par.pdists, par.coal_data, par.norms))
            configured[] = true
        end
        mm = reshape(collect(Float64, m), length(par.NProgMoms) == 0 ? 0 : sum(par.NProgMoms), :)
        out = similar(mm)
        rhs_coal_batched!(out, mm, ctx)
        dm .= reshape(out, size(dm))
    end
end

# device-resident ensembles for long runs -----------------------------------------------------------------
mutable struct Ensemble
    ctx::Context
    handle::Ptr{Cvoid}
    n::Int
end
function Ensemble(ctx::Context, n::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cloudy_state_create, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx.handle, n, h))
    Ensemble(ctx, h[], n)
end
upload!(e::Ensemble, m::Matrix{Float64}) =
    check(ccall((:cloudy_state_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), e.ctx.handle, e.handle, m, e.n))
download!(m::Matrix{Float64}, e::Ensemble) =
    check(ccall((:cloudy_state_download, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), e.ctx.handle, e.handle, m, e.n))
"`solve(prob, SSPRK33(), dt = dt)` for `n_steps` steps, state kept on the GPU (model 0 = box, 1 = rainshaft)."
ssprk33!(e::Ensemble, dt, n_steps; model = 0) =
    check(ccall((:cloudy_ssprk33_steps, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int32, Int32), e.ctx.handle, e.handle, dt, n_steps, model))

# single-object methods (same names/arguments as the reference) ----------------------------------------------
params3(d) = (p = collect(Float64, ntuple(i -> getfield(d, i), nparams(d))); length(p) < 3 && push!(p, 1.0); p)
function moment_b200(ctx::Context, d, q::Float64)
    out = Ref{Float64}(0)
    check(ccall((:cloudy_moment, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cdouble, Ref{Float64}), ctx.handle, kind_code(d), params3(d), q, out))
    out[]
end
function moment_source_helper_b200(ctx::Context, d, p1, p2, x_threshold, n_bins_per_log_unit = 15)
    out = Ref{Float64}(0)
    check(ccall((:cloudy_moment_source_helper, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cdouble, Cdouble, Cdouble, Int32, Ref{Float64}),
                ctx.handle, kind_code(d), params3(d), p1, p2, x_threshold, n_bins_per_log_unit, out))
    out[]
end
"`get_coal_ints(AnalyticalCoalStyle(), pdists, coal_data)` (src/Sources/Coalescence.jl:115-150) after `set_config!` with unit norms."
function get_coal_ints_b200(ctx::Context, pdists::NTuple{N}) where {N}
    params = zeros(Float64, 3, N)
    for (i, d) in enumerate(pdists)
        params[:, i] .= params3(d)
    end
    out = zeros(Float64, sum(nparams, pdists))
    check(ccall((:cloudy_get_coal_ints_1, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.handle, params, out))
    Tuple(out)
end

"`rhs_condensation!` for a device ensemble (test/examples/utils/box_model_helpers.jl:55-67); `s` scalar supersaturation."
cond_evap!(dm::Ensemble, m::Ensemble, s, ξ; ρ_l = 1000.0) =
    check(ccall((:cloudy_cond_evap, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cvoid}),
                m.ctx.handle, m.handle, s, C_NULL, ξ, ρ_l, dm.handle))

"`compute_threshold(pdist, percentile)` (ParticleDistributions.jl:747-761) on the device."
function compute_threshold_b200(ctx::Context, d, percentile = 0.97, minx = 1e-18)
    out = Ref{Float64}(0)
    check(ccall((:cloudy_compute_threshold, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cdouble, Cdouble, Ref{Float64}),
                ctx.handle, kind_code(d), params3(d), percentile, minx, out))
    out[]
end

"`get_standard_N_q(pdists, size_cutoff)` → (; N_liq, N_rai, M_liq, M_rai) (ParticleDistributions.jl:634-687)."
function get_standard_N_q_b200(ctx::Context, pdists::NTuple{N}, size_cutoff = 1e-6) where {N}
    kinds = Int32[kind_code(d) for d in pdists]
    params = zeros(Float64, 3, N)
    for (i, d) in enumerate(pdists)
        params[:, i] .= params3(d)
    end
    out = zeros(Float64, 4)
    check(ccall((:cloudy_get_standard_N_q_1, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Cdouble, Ptr{Float64}),
                ctx.handle, N, kinds, params, size_cutoff, out))
    (; N_liq = out[1], N_rai = out[2], M_liq = out[3], M_rai = out[4])
end

"Σ over parcels of every prognostic moment (conservation diagnostic, cf. moments_sum in netcdf_helpers.jl:34-42)."
function moment_sums(e::Ensemble, n_slots::Integer)
    out = zeros(Float64, n_slots)
    check(ccall((:cloudy_moment_sums, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), e.ctx.handle, e.handle, out))
    out
end

end # module
