# SABCB200.jl -- thin Julia host code over libsabc_b200.so (include/sabc_b200.h).
#
# Keeps SimulatedAnnealingABC.jl's surface: `sabc(f_dist, prior, args...; kw...)` and `update_population!(res, f_dist, prior, ...)`
# gain methods that dispatch on a `DeviceModel` in place of a closure; closures keep using the package's own CPU code.
# NOT EXECUTED in the build image (no julia binary there); the identical call sequence is exercised through the Python
# mirror (simulatedannealingabc.jl_b200/api.py) and the C ABI tests.  See INTEGRATION.md.
module SABCB200

using SimulatedAnnealingABC
import SimulatedAnnealingABC: sabc, update_population!, SABCresult, SABCstate, Proposal,
                              DifferentialEvolution, StretchMove, RandomWalk
using Distributions: Distribution, Uniform, Normal, Exponential, LogNormal, Gamma, Beta, Cauchy, Laplace, Weibull, InverseGamma, Product, params
import Distributions

const libsabc = get(ENV, "SABC_B200_LIB", joinpath(@__DIR__, "..", "libsabc_b200.so"))

# ---- C structs (layout of sabc_config / sabc_timing) ----
struct SabcConfig
    n_particles::Int64; n_para::Int32; n_stats::Int32; algorithm::Int32; proposal::Int32
    prop_par::NTuple{2,Float64}; v::Float64; delta::Float64; resample::Int64; seed::UInt64
    model_name::Cstring; model_par::Ptr{Float64}; n_model_par::Int32; device::Int32
    prior_kind::Ptr{Int32}; prior_par::Ptr{Float64}
    rank::Int32; world_size::Int32; nccl_unique_id::Ptr{Cvoid}; flags::UInt32; ecdf_max_knots::Int32
    n_gpus::Int32; gpu_ids::Ptr{Int32}                                  # ABI 2: one Julia session drives n_gpus devices
end
# field order, offsets and sizes are compared with offsetof()/sizeof() of include/sabc_b200.h by tests/test_julia_layout.py

struct SABCDeviceError <: Exception
    code::Int; msg::String
end
Base.showerror(io::IO, e::SABCDeviceError) = print(io, "libsabc_b200 [", e.code, "]: ", e.msg)

last_error() = unsafe_string(ccall((:sabc_last_error, libsabc), Cstring, ()))
# the reference raises ErrorException via error(...) for codes -1..-8; keep that type for drop-in `@test_throws ErrorException`
check(rc) = rc == 0 ? nothing : (-8 <= rc <= -1 ? error(last_error()) : throw(SABCDeviceError(rc, last_error())))

# ---- device model plug-ins: the GPU form of f_dist ----
struct DeviceModel
    name::String; n_para::Int; n_stats::Int; par::Vector{Float64}
end
gauss_mean(ȳ_obs; σ=1.0, n_obs=10) = DeviceModel("gauss_mean", 1, 1, [ȳ_obs, σ / sqrt(n_obs)])
gauss_sample(n_obs, obs_mean, obs_second=nothing; n_para=1, σ=1.0, second_is_sum=false) =
    DeviceModel("gauss_sample_d$(n_para)s$(obs_second === nothing ? 1 : 2)", n_para, obs_second === nothing ? 1 : 2,
                [n_obs, σ, obs_mean, something(obs_second, 0.0), second_is_sum ? 1.0 : 0.0])
logistic(x_obs; x0=10.0) = DeviceModel("logistic", 3, 20, vcat([x0, 20.0], x_obs))
sir_tauleap(obs_total, obs_peak, obs_tpeak; pop=1e5, n_steps=50, τ=1.0) =
    DeviceModel("sir_tauleap", 4, 3, [pop, n_steps, τ, obs_total, obs_peak, obs_tpeak])

# ---- plug-in encodings ----
prior_components(p::Union{Uniform,Normal,Exponential,LogNormal,Gamma,Beta,Cauchy,Laplace,Weibull,InverseGamma}) = [p]
prior_components(p::Product) = collect(p.v)
# Distributions >= 0.25.72 returns a ProductDistribution from product_distribution(...); older versions a Product
if isdefined(Distributions, :ProductDistribution)
    prior_components(p::Distributions.ProductDistribution) = collect(p.dists)
end
prior_kind(::Uniform) = Int32(0); prior_kind(::Normal) = Int32(1); prior_kind(::Exponential) = Int32(2); prior_kind(::LogNormal) = Int32(3); prior_kind(::Gamma) = Int32(4); prior_kind(::Beta) = Int32(5)
prior_kind(::Cauchy) = Int32(6); prior_kind(::Laplace) = Int32(7); prior_kind(::Weibull) = Int32(8); prior_kind(::InverseGamma) = Int32(9)
prior_params(c) = (p = params(c); length(p) == 2 ? (p[1], p[2]) : (p[1], 0.0))
proposal_code(p::DifferentialEvolution) = (Int32(0), (p.γ0, p.σ_gamma))
proposal_code(p::StretchMove) = (Int32(1), (p.a, 0.0))
proposal_code(p::RandomWalk) = (Int32(2), (p.β, 0.0))

mutable struct Engine
    h::Ptr{Cvoid}; model::DeviceModel; n::Int; d::Int; s::Int; n_eps::Int
end
destroy!(e::Engine) = (e.h != C_NULL && ccall((:sabc_destroy, libsabc), Cint, (Ptr{Cvoid},), e.h); e.h = C_NULL; nothing)

function Engine(model::DeviceModel, prior::Distribution; n_particles, algorithm, proposal::Proposal, resample, v, δ,
                seed=0x5ABC, device=-1, ecdf_max_knots=0, n_gpus=0, gpu_ids::Vector{Int32}=Int32[])
    comps = prior_components(prior)
    length(comps) == model.n_para || error("prior has $(length(comps)) components, model $(model.name) has $(model.n_para) parameters")
    kinds = Int32[prior_kind(c) for c in comps]
    ppar = Float64[x for c in comps for x in prior_params(c)]
    pcode, ppars = proposal_code(proposal)
    alg = algorithm == :multi_eps ? Int32(1) : Int32(0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    isempty(gpu_ids) || (n_gpus = length(gpu_ids))
    GC.@preserve kinds ppar model gpu_ids begin
        cfg = SabcConfig(n_particles, model.n_para, model.n_stats, alg, pcode, ppars, v, δ, resample, seed,
                         Base.unsafe_convert(Cstring, model.name), pointer(model.par), length(model.par), device,
                         pointer(kinds), pointer(ppar), 0, 1, C_NULL, 0, ecdf_max_knots,
                         n_gpus, isempty(gpu_ids) ? Ptr{Int32}(C_NULL) : pointer(gpu_ids))
        check(ccall((:sabc_create, libsabc), Cint, (Ref{Ptr{Cvoid}}, Ref{SabcConfig}), h, cfg))
    end
    e = Engine(h[], model, n_particles, model.n_para, model.n_stats, algorithm == :multi_eps ? model.n_stats : 1)
    finalizer(destroy!, e)
end

# ---- state across the boundary (SABCresult / SABCstate, src/SimulatedAnnealingABC.jl:28-60) ----
function fetch_result!(e::Engine, algorithm::Symbol)
    θ = Matrix{Float64}(undef, e.n, e.d); u = Matrix{Float64}(undef, e.n, e.s); ρ = similar(u)
    check(ccall((:sabc_get_population, libsabc), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, θ, u, ρ))
    ϵ = Vector{Float64}(undef, e.n_eps); cnt = Vector{Int64}(undef, 4)
    check(ccall((:sabc_get_state, libsabc), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), e.h, ϵ, cnt))
    nrec = Ref{Int64}(0); check(ccall((:sabc_history_len, libsabc), Cint, (Ptr{Cvoid}, Ref{Int64}), e.h, nrec))
    ϵh = Matrix{Float64}(undef, e.n_eps, nrec[]); uh = Matrix{Float64}(undef, e.s, nrec[]); ρh = similar(uh)   # row-major records
    check(ccall((:sabc_get_history, libsabc), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, ϵh, uh, ρh))
    population = e.d == 1 ? vec(θ) : [θ[i, :] for i in 1:e.n]
    state = SABCstate(ϵ, algorithm, [ϵh[:, k] for k in 1:nrec[]], [ρh[:, k] for k in 1:nrec[]], [uh[:, k] for k in 1:nrec[]],
                      e,                       # cdfs_dist_prior slot carries the engine handle (device-resident ECDF tables)
                      cnt[1], cnt[2], cnt[3], cnt[4])
    SABCresult(population, u, ρ, state)
end

"""
    sabc(f_dist::DeviceModel, prior; n_particles, n_simulation, algorithm, proposal, resample, v, δ, checkpoint_history, ...)

Device method of `SimulatedAnnealingABC.sabc` (src/SimulatedAnnealingABC.jl:451-492).
"""
function sabc(f_dist::DeviceModel, prior::Distribution; n_particles=100, n_simulation=10_000, algorithm=:single_eps,
              proposal::Proposal=DifferentialEvolution(n_para=length(prior)), resample=2 * n_particles, v=1.0, δ=0.1,
              checkpoint_history=1, show_progressbar=false, show_checkpoint=Inf, type=nothing, seed=0x5ABC, device=-1,
              n_gpus=0, gpu_ids::Vector{Int32}=Int32[])
    type === nothing || (algorithm = Dict(:single => :single_eps, :multi => :multi_eps, :hybrid => :single_eps)[type])
    (algorithm == :multi_eps || algorithm == :single_eps) ||
        error("Argument `algorithm` must be :multi_eps or :single_eps, not `$algorithm`!")
    n_simulation < n_particles && error("`n_simulation = $n_simulation` is too small for $n_particles particles.")
    # n_gpus > 1: this one session drives all of them; the particles are sharded inside the library, the arrays returned below and
    # passed to update_population! stay the global N x d / N x s matrices
    e = Engine(f_dist, prior; n_particles, algorithm, proposal, resample, v, δ, seed, device, n_gpus, gpu_ids)
    check(ccall((:sabc_init, libsabc), Cint, (Ptr{Cvoid},), e.h))                       # initialization()
    n_sim_remaining = n_simulation - n_particles
    n_sim_remaining < n_particles && @warn "`n_simulation` too small to update all particles!"
    check(ccall((:sabc_update, libsabc), Cint, (Ptr{Cvoid}, Int64, Int64), e.h, n_sim_remaining, checkpoint_history))
    fetch_result!(e, algorithm)
end

"""
    update_population!(res::SABCresult, f_dist::DeviceModel, prior; n_simulation, v, δ, proposal, resample, checkpoint_history)

Device method of `update_population!` (src/SimulatedAnnealingABC.jl:251-402): uploads the host-resident result, runs the
updates, downloads into the same arrays (`sabc_update_host`), appends the histories and mutates the counters.
"""
function update_population!(res::SABCresult, f_dist::DeviceModel, prior::Distribution; n_simulation, v=1.0, δ=0.1,
                            proposal::Proposal=DifferentialEvolution(n_para=length(prior)),
                            resample=2 * length(res.population), checkpoint_history=1, show_progressbar=false, show_checkpoint=Inf)
    v <= 0 && error("Annealing speed `v` must be positive.")
    δ <= 0 && error("Resamping intensity `δ` must be positive.")
    st = res.state
    e = st.cdfs_dist_prior::Engine
    pcode, ppars = proposal_code(proposal)
    check(ccall((:sabc_set_tuning, libsabc), Cint, (Ptr{Cvoid}, Float64, Float64, Int64, Int32, Ref{NTuple{2,Float64}}),
                e.h, v, δ, resample, pcode, ppars))
    θ = e.d == 1 ? reshape(copy(res.population), :, 1) : permutedims(reduce(hcat, res.population))
    cnt = Int64[st.n_simulation, st.n_accept, st.n_resampling, st.n_population_updates]
    nrec0 = Ref{Int64}(0); ccall((:sabc_history_len, libsabc), Cint, (Ptr{Cvoid}, Ref{Int64}), e.h, nrec0)
    GC.@preserve θ res cnt begin
        check(ccall((:sabc_update_host, libsabc), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int64, Int64),
                    e.h, θ, res.u, res.ρ, st.ϵ, cnt, n_simulation, checkpoint_history))
    end
    if e.d == 1
        res.population .= vec(θ)
    else
        for i in eachindex(res.population); res.population[i] .= @view θ[i, :]; end
    end
    fresh = fetch_result!(e, st.algorithm).state                                          # histories incl. the new records
    st.ϵ_history = fresh.ϵ_history; st.u_history = fresh.u_history; st.ρ_history = fresh.ρ_history
    st.n_simulation, st.n_accept, st.n_resampling, st.n_population_updates = cnt
    res
end

# the package's documented example (docs/src/example.md): event-driven SIR, f_dist_multi_stats / f_dist_single_stat
sir_gillespie(obs_total, obs_peak, obs_tpeak; S0=99, I0=1, R0=0, t_max=160.0, single_stat=false) =
    DeviceModel("sir_gillespie_s$(single_stat ? 1 : 3)", 2, single_stat ? 1 : 3, Float64[S0, I0, R0, t_max, obs_total, obs_peak, obs_tpeak])

export DeviceModel, gauss_mean, gauss_sample, logistic, sir_tauleap, sir_gillespie

end # module
