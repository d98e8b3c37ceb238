"""
    AdvancedPSB200

Julia host side of `libaps_b200.so` (C ABI in `include/aps_b200.h`): adds methods to the
generic functions AdvancedPS.jl already dispatches on, so `sample(rng, model, SMC(n))`,
`step(rng, model, PG(n) | PGAS(n), state)` and the `resample_*` callables run on a B200 when the
model is a [`DeviceSSM`](@ref). Nothing in AdvancedPS.jl changes.

Reference interfaces replaced (paths relative to AdvancedPS.jl v0.7.2):
  * `AbstractMCMC.sample(rng, model, ::SMC)`           src/smc.jl:35-57
  * `AbstractMCMC.step(rng, model, ::PG/PGAS, state)`  src/smc.jl:101-129
  * `sweep!` and everything under it                    src/container.jl:171-363, src/pgas.jl:26-128
  * resampler callables `(rng, w, n) -> Vector{Int}`    src/container.jl:182, src/resampling.jl
  * `logZ`, `getweights`, `effectiveSampleSize`         src/container.jl:95-119

NOTE: the build image of this repository has no Julia toolchain; this package is written against
the ABI and reviewed, and the same ABI is exercised end to end by the Python host mirror
(`advancedps.jl_b200/`) in the parity tests. `test/runtests.jl` restates the reference's own
known-answer tests for anyone with Julia and a B200.
"""
module AdvancedPSB200

using AbstractMCMC: AbstractMCMC
using AdvancedPS: AdvancedPS
using Random: Random
using SSMProblems: SSMProblems

export DeviceSSM, LinearGaussianSSM, StochasticVolatilitySSM, ConstantLogLikSSM
export gpu_resample_multinomial, gpu_resample_residual, gpu_resample_stratified, gpu_resample_systematic
export gpu_logsumexp, gpu_softmax, gpu_ess, gpu_randcat

"Path of the shared library; override with ENV[\"APS_B200_LIB\"]."
const lib = get(ENV, "APS_B200_LIB", "libaps_b200")

const APS_MAX_D = 4

# ------------------------------------------------------------------ struct mirrors (include/aps_model.h, aps_b200.h)
struct ApsModel
    obs_kind::Int32
    d::Int32
    dy::Int32
    reserved::Int32
    mu0::NTuple{4,Float64}
    sigma0::NTuple{4,Float64}
    A::NTuple{16,Float64}      # row-major d x d, leading dimension APS_MAX_D
    b::NTuple{4,Float64}
    q::NTuple{4,Float64}
    H::NTuple{16,Float64}      # row-major dy x d, leading dimension APS_MAX_D
    r::NTuple{4,Float64}
end

struct ApsConfig
    model::ApsModel
    n_particles::Int64
    n_steps::Int64
    sampler::Int32             # 0 SMC, 1 PG, 2 PGAS
    resampler::Int32           # 0 multinomial, 1 residual, 2 stratified, 3 systematic
    ess_threshold::Float64     # NaN: bare resampler function (resample at every step)
    keep_history::Int32
    device::Int32
    rank::Int32
    world_size::Int32
end

const OBS_LINEAR_GAUSS, OBS_STOCH_VOL, OBS_CONST = Int32(0), Int32(1), Int32(2)

"Non-zero status -> `ErrorException` with the library's message (src/resampling.jl:103,120,154,169)."
function check(rc::Integer)
    rc == 0 && return nothing
    return error(unsafe_string(ccall((:aps_last_error, lib), Cstring, ())))
end

pad4(v, fill=0.0) = ntuple(i -> i <= length(v) ? Float64(v[i]) : fill, 4)
function pad16(M, rows, cols)
    return ntuple(16) do k
        i, j = divrem(k - 1, APS_MAX_D) .+ 1
        (i <= rows && j <= cols) ? Float64(M[i, j]) : 0.0
    end
end

# ------------------------------------------------------------------ device-resident model families
"""
    DeviceSSM(model::ApsModel, Y)

A state-space model of a family the device path knows (linear-Gaussian, stochastic volatility,
constant log-likelihood), with its observations `Y` (`T x dy`). Plays the role of
`TracedSSM(StateSpaceModel(prior, dyn, obs), Y)` (src/model.jl:13-22); `X` carries a trajectory
(`T x d`) when the object is part of a `PGState`. Noise parameters are STANDARD DEVIATIONS, as in
test/linear-gaussian.jl:59-87 and examples/particle-gibbs/script.jl:55-83.
"""
mutable struct DeviceSSM <: SSMProblems.AbstractStateSpaceModel
    model::ApsModel
    Y::Matrix{Float64}
    X::Union{Nothing,Matrix{Float64}}
end
DeviceSSM(model::ApsModel, Y::AbstractMatrix) = DeviceSSM(model, Matrix{Float64}(Y), nothing)
DeviceSSM(model::ApsModel, y::AbstractVector) = DeviceSSM(model, reshape(Vector{Float64}(y), :, 1), nothing)

"x1 ~ N(mu0, diag(sigma0)^2); x_t = A x_{t-1} + b + diag(q) eps; y_t = H x_t + diag(r) eta."
function LinearGaussianSSM(A, b, q, H, r, mu0, sigma0, Y)
    A, H = reshape(collect(Float64, A), :, length(b)), reshape(collect(Float64, H), length(r), :)
    d, dy = length(b), length(r)
    (1 <= d <= APS_MAX_D && 1 <= dy <= APS_MAX_D) || error("state / observation dimension must be in 1..4")
    m = ApsModel(OBS_LINEAR_GAUSS, d, dy, 0, pad4(mu0), pad4(sigma0), pad16(A, d, d), pad4(b), pad4(q, 1.0),
                 pad16(H, dy, d), pad4(r, 1.0))
    return DeviceSSM(m, Y)
end
"The 1-d model of test/linear-gaussian.jl:32-42: x' = a x + b + q eps, y = h x + r eta, x1 ~ N(x0, p0)."
function LinearGaussianSSM(; a=0.5, b=0.2, q=0.1, h=1.0, r=0.1, x0=0.0, p0=1.0, Y)
    return LinearGaussianSSM(fill(a, 1, 1), [b], [q], fill(h, 1, 1), [r], [x0], [p0], Y)
end
"examples/particle-gibbs/script.jl:55-83: x1 ~ N(0, q), x' ~ N(a x, q), y ~ N(0, exp(x/2))."
function StochasticVolatilitySSM(; a=0.9, q=0.5, Y)
    m = ApsModel(OBS_STOCH_VOL, 1, 1, 0, pad4([0.0]), pad4([q]), pad16(fill(a, 1, 1), 1, 1), pad4([0.0]),
                 pad4([q], 1.0), pad16(zeros(1, 1), 1, 1), pad4([1.0], 1.0))
    return DeviceSSM(m, Y)
end
"log g(y_t | x_t) = y_t, independent of the state (restates test/smc.jl:70-104, test/container.jl:4-18)."
function ConstantLogLikSSM(; Y)
    m = ApsModel(OBS_CONST, 1, 1, 0, pad4([0.0]), pad4([1.0]), pad16(zeros(1, 1), 1, 1), pad4([0.0]),
                 pad4([1.0], 1.0), pad16(zeros(1, 1), 1, 1), pad4([1.0], 1.0))
    return DeviceSSM(m, Y)
end

# ------------------------------------------------------------------ handle (device-side ParticleContainer)
mutable struct Handle
    ptr::Ptr{Cvoid}
    n::Int
    T::Int
    d::Int
    function Handle(cfg::ApsConfig, Y::Matrix{Float64})
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:aps_create, lib), Cint, (Ref{ApsConfig}, Ref{Ptr{Cvoid}}), cfg, p))
        h = new(p[], cfg.n_particles, cfg.n_steps, cfg.model.d)
        finalizer(x -> (x.ptr == C_NULL || ccall((:aps_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
        Yt = permutedims(Y)                                  # dy x T column-major == T x dy row-major
        GC.@preserve Yt check(ccall((:aps_set_observations, lib), Cint,
                                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), h.ptr, Yt, size(Y, 1), size(Y, 2)))
        return h
    end
end

resampler_kind(::typeof(AdvancedPS.resample_multinomial)) = Int32(0)
resampler_kind(::typeof(AdvancedPS.resample_residual)) = Int32(1)
resampler_kind(::typeof(AdvancedPS.resample_stratified)) = Int32(2)
resampler_kind(::typeof(AdvancedPS.resample_systematic)) = Int32(3)
resampler_kind(f) = error("the device sweep needs one of AdvancedPS.resample_* (got $f)")
kind_thr(r::AdvancedPS.ResampleWithESSThreshold) = (resampler_kind(r.resampler), Float64(r.threshold))
kind_thr(r) = (resampler_kind(r), NaN)                       # bare function: resample at every step

sampler_id(::AdvancedPS.SMC) = Int32(0)
sampler_id(::AdvancedPS.PG) = Int32(1)
sampler_id(::AdvancedPS.PGAS) = Int32(2)

# one live handle per (model object, sampler shape): particle stores can be tens of GB
const HANDLES = IdDict{Any,Tuple{Any,Handle}}()
function handle_for(model::DeviceSSM, sampler; device::Integer=0)
    k, thr = kind_thr(sampler.resampler)
    key = (sampler_id(sampler), sampler.nparticles, k, thr, device)
    cached = get(HANDLES, model, nothing)
    # isequal, not ==: a bare resampler carries thr = NaN, and NaN == NaN is false -- `==` would build
    # a new device handle (tens of GB) on every call and leave the old one to a finalizer the GC has
    # no reason to run (it does not see device memory)
    if cached !== nothing && isequal(cached[1], key)
        h = cached[2]
        set_observations!(h, model.Y)                        # the caller may have changed model.Y in place
        return h
    end
    for (_, (_, old)) in HANDLES                             # release device memory NOW, not at some GC
        destroy!(old)
    end
    empty!(HANDLES)
    cfg = ApsConfig(model.model, sampler.nparticles, size(model.Y, 1), key[1], k, thr, 1, device, 0, 1)
    h = Handle(cfg, model.Y)
    HANDLES[model] = (key, h)
    return h
end

function set_observations!(h::Handle, Y::Matrix{Float64})
    Yt = permutedims(Y)                                      # dy x T column-major == T x dy row-major
    GC.@preserve Yt check(ccall((:aps_set_observations, lib), Cint,
                                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), h.ptr, Yt, size(Y, 1), size(Y, 2)))
    return h
end

function destroy!(h::Handle)
    if h.ptr != C_NULL
        ccall((:aps_destroy, lib), Cint, (Ptr{Cvoid},), h.ptr)
        h.ptr = C_NULL
    end
    return nothing
end

function sweep!(h::Handle, seed::UInt64, ref::Union{Nothing,Matrix{Float64}}; ref_on_device::Bool=false)
    logev = Ref{Float64}()
    if ref_on_device
        check(ccall((:aps_sweep, lib), Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ref{Float64}),
                    h.ptr, seed, Ptr{Float64}(1), logev))    # APS_REF_ON_DEVICE
    elseif ref === nothing
        check(ccall((:aps_sweep, lib), Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ref{Float64}),
                    h.ptr, seed, C_NULL, logev))
    else
        rt = permutedims(ref)                                # d x T column-major == T x d row-major
        GC.@preserve rt check(ccall((:aps_sweep, lib), Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ref{Float64}),
                                    h.ptr, seed, rt, logev))
    end
    return logev[]
end

"Trajectories of an `SMCSample`, materialised one at a time through the device genealogy."
struct LazyTrajectories <: AbstractVector{AdvancedPS.Trace}
    h::Handle
    model::DeviceSSM
end
Base.size(t::LazyTrajectories) = (t.h.n,)
function Base.getindex(t::LazyTrajectories, i::Int)
    traj = Matrix{Float64}(undef, t.h.d, t.h.T)
    check(ccall((:aps_get_trajectory, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), t.h.ptr, i - 1, traj))
    return AdvancedPS.Trace(DeviceSSM(t.model.model, t.model.Y, permutedims(traj)), AdvancedPS.TracedRNG())
end

# collect(pc) of SMCSample (src/smc.jl:56): every trajectory in one call (T x N x d on the device side)
function Base.collect(t::LazyTrajectories)
    X = Array{Float64,3}(undef, t.h.d, t.h.n, t.h.T)          # d x N x T column-major == T x N x d row-major
    check(ccall((:aps_get_trajectories, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), t.h.ptr, X))
    return [AdvancedPS.Trace(DeviceSSM(t.model.model, t.model.Y, permutedims(X[:, i, :])), AdvancedPS.TracedRNG())
            for i in 1:t.h.n]
end

# ------------------------------------------------------------------ the device container, call by call
# (src/container.jl:171-363 as single calls: what test/container.jl and test/pgas.jl:61-91 drive)
struct DeviceContainer
    h::Handle
end
function DeviceContainer(rng::Random.AbstractRNG, model::DeviceSSM, sampler, ref::Union{Nothing,Matrix{Float64}}=nothing)
    h = handle_for(model, sampler)
    rt = ref === nothing ? nothing : permutedims(ref)
    GC.@preserve rt check(ccall((:aps_pc_begin, lib), Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}),
                                h.ptr, rand(rng, UInt64), rt === nothing ? C_NULL : pointer(rt)))
    return DeviceContainer(h)
end
function AdvancedPS.reweight!(pc::DeviceContainer, ref=nothing)
    done = Ref{Int32}(0)
    check(ccall((:aps_pc_reweight, lib), Cint, (Ptr{Cvoid}, Ref{Int32}), pc.h.ptr, done))
    return done[] != 0
end
function AdvancedPS.resample_propagate!(rng, pc::DeviceContainer, sampler, resampler, ref=nothing)
    res = Ref{Int32}(0)
    check(ccall((:aps_pc_resample_propagate, lib), Cint, (Ptr{Cvoid}, Ref{Int32}), pc.h.ptr, res))
    return pc
end
function AdvancedPS.logZ(pc::DeviceContainer)
    z = Ref{Float64}()
    check(ccall((:aps_pc_logz, lib), Cint, (Ptr{Cvoid}, Ref{Float64}), pc.h.ptr, z))
    return z[]
end
function set_logweights!(pc::DeviceContainer, logWs::Vector{Float64})      # pc.logWs = v (test/pgas.jl:82)
    check(ccall((:aps_set_logweights, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), pc.h.ptr, logWs))
    return pc
end

# ------------------------------------------------------------------ AbstractMCMC.sample for SMC (src/smc.jl:35-57)
function AbstractMCMC.sample(rng::Random.AbstractRNG, model::DeviceSSM, sampler::AdvancedPS.SMC; kwargs...)
    if !isempty(kwargs)
        @warn "keyword arguments $(keys(kwargs)) are not supported by `SMC`"      # src/smc.jl:41-43
    end
    h = handle_for(model, sampler)
    # ONE rand(rng, UInt64) seeds the sweep (replaces the N+1 draws of seed_from_rng!, container.jl:143-159)
    logev = sweep!(h, rand(rng, UInt64), nothing)
    w = Vector{Float64}(undef, sampler.nparticles)
    check(ccall((:aps_get_weights, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, w))
    return AdvancedPS.SMCSample(LazyTrajectories(h, model), w, logev)
end

# ------------------------------------------------------------------ AbstractMCMC.step for PG / PGAS (src/smc.jl:101-129)
function AbstractMCMC.step(rng::Random.AbstractRNG, model::DeviceSSM,
                           sampler::Union{AdvancedPS.PGAS,AdvancedPS.PG},
                           state::Union{AdvancedPS.PGState,Nothing}=nothing; kwargs...)
    h = handle_for(model, sampler)
    seed = rand(rng, UInt64)
    logev = if state === nothing
        sweep!(h, seed, nothing)
    elseif state.trajectory.model isa DeviceSSM && state.trajectory.model.X !== nothing
        sweep!(h, seed, state.trajectory.model.X)
    else
        error("PGState does not carry a DeviceSSM trajectory")
    end
    traj = Matrix{Float64}(undef, h.d, h.T)
    slot = Ref{Int64}()
    check(ccall((:aps_pick_trajectory, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), h.ptr, traj, slot))  # smc.jl:127
    tr = AdvancedPS.Trace(DeviceSSM(model.model, model.Y, permutedims(traj)), AdvancedPS.TracedRNG())
    return AdvancedPS.PGSample(tr, logev), AdvancedPS.PGState(tr)       # users read .trajectory.model.X
end

# ------------------------------------------------------------------ operator level (boundary 1, SURVEY 8b)
# GPU versions of the resampler callables. They satisfy `(rng, w, n) -> Vector{Int}` (1-based), so the
# UNMODIFIED reference sweep accepts them: SMC(1000, gpu_resample_systematic) works for Libtask /
# Turing models too (src/container.jl:182 calls `randcat(pc.rng, weights, nresamples)`).
function gpu_resample(kind::Integer, rng::Random.AbstractRNG, w::AbstractVector{<:Real}, n::Integer=length(w))
    isempty(w) && error("weight vector is empty")                               # src/resampling.jl:103,154
    out = Vector{Int}(undef, n)
    wf = convert(Vector{Float64}, w)
    GC.@preserve wf out check(ccall((:aps_resample, lib), Cint,
                                    (Cint, Ptr{Float64}, Int64, Int64, UInt64, UInt64, Ptr{Int64}),
                                    kind, wf, length(wf), n, rand(rng, UInt64), 0, out))
    return out
end
gpu_resample_multinomial(rng, w, n=length(w)) = gpu_resample(0, rng, w, n)      # src/resampling.jl:31-35
gpu_resample_residual(rng, w, n=length(w)) = gpu_resample(1, rng, w, n)         # :53-81
gpu_resample_stratified(rng, w, n=length(w)) = gpu_resample(2, rng, w, n)       # :98-131
gpu_resample_systematic(rng, w, n=length(w)) = gpu_resample(3, rng, w, n)       # :149-183

function gpu_logsumexp(logw::Vector{Float64})                                   # logZ, src/container.jl:109
    out = Ref{Float64}()
    check(ccall((:aps_logsumexp, lib), Cint, (Ptr{Float64}, Int64, Ref{Float64}), logw, length(logw), out))
    return out[]
end
function gpu_softmax(logw::Vector{Float64})                                     # getweights, :95
    w = similar(logw)
    check(ccall((:aps_softmax, lib), Cint, (Ptr{Float64}, Int64, Ptr{Float64}), logw, length(logw), w))
    return w
end
function gpu_ess(logw::Vector{Float64})                                         # effectiveSampleSize, :116-119
    out = Ref{Float64}()
    check(ccall((:aps_ess, lib), Cint, (Ptr{Float64}, Int64, Ref{Float64}), logw, length(logw), out))
    return out[]
end
function gpu_randcat(rng::Random.AbstractRNG, p::Vector{Float64})               # randcat, src/resampling.jl:11-21
    out = Ref{Int64}()
    check(ccall((:aps_randcat, lib), Cint, (Ptr{Float64}, Int64, UInt64, UInt64, Ref{Int64}),
                p, length(p), rand(rng, UInt64), 0, out))
    return Int(out[])
end

end # module
