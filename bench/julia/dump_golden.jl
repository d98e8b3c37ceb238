# Dumps golden vectors from the REAL reference (AdvancedPS.jl on the CPU) for the hot path, so the
# oracle's SEQ mode -- the restatement of the reference's fp64 order -- can be pinned against the
# reference itself instead of against its own restatement. The repository's build image has no
# Julia: run this anywhere AdvancedPS.jl v0.7 is installed and commit the three files it writes.
#
#   julia --project=<env with AdvancedPS> bench/julia/dump_golden.jl [outdir = tests/golden]
#
# tests/test_reference_golden.py loads the files when present (and says "parity unpinned" when not):
#   reference_resample.json   (weights, n, uniforms drawn, indices) for resample_systematic /
#                             resample_stratified (src/resampling.jl:98-183); (weights, n, indices)
#                             for resample_multinomial / resample_residual (:31-81; residual only
#                             where upstream's code path works, SURVEY App. B Q1); randcat (:11-21)
#   reference_weights.json    logWs -> getweights / logZ / effectiveSampleSize (src/container.jl:95-119)
#   reference_known.json      the RNG-independent known answers of test/container.jl:45-68,
#                             test/resampling.jl:12-15 and the PG/PGAS constructor defaults
#
# The uniforms a resampler consumed are recovered by replaying a copy of the rng: systematic draws
# ONE rand(rng) (src/resampling.jl:160), stratified draws n, in order (:112).
using AdvancedPS, Random

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "..", "tests", "golden")
mkpath(outdir)

# ---- minimal JSON writer (no package dependency); non-finite floats become strings
js(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : (isnan(x) ? "\"NaN\"" : (x > 0 ? "\"Inf\"" : "\"-Inf\""))
js(x::Integer) = string(x)
js(x::Bool) = x ? "true" : "false"
js(x::AbstractString) = "\"" * escape_string(x) * "\""
js(x::AbstractVector) = "[" * join((js(v) for v in x), ",") * "]"
js(x::AbstractDict) = "{" * join((js(string(k)) * ":" * js(v) for (k, v) in x), ",") * "}"
function dump(name, obj)
    path = joinpath(outdir, name)
    open(io -> write(io, js(obj), "\n"), path, "w")
    println("wrote ", path)
end

# ---- weight vectors: the shapes tests/test_gpu_operators.py uses
function weight_cases(rng)
    cases = Pair{String,Vector{Float64}}[]
    for m in (3, 10, 1000, 2049)
        push!(cases, "uniform-$m" => fill(1.0 / m, m))
        w = exp.(randn(rng, m)); push!(cases, "lognormal-$m" => w ./ sum(w))
        w = rand(rng, m) .^ 8;   push!(cases, "skewed-$m" => w ./ sum(w))
        oh = zeros(m); oh[cld(m, 2)] = 1.0; push!(cases, "onehot-$m" => oh)
    end
    push!(cases, "test-resampling" => [0.3, 0.4, 0.3])            # test/resampling.jl:2
    return cases
end

gen = Random.MersenneTwister(20241017)
resample_records = Dict{String,Any}[]
for (name, w) in weight_cases(gen), n in unique((length(w), max(1, length(w) ÷ 3), length(w) + 17))
    for seed in (1, 2)
        rng = Random.MersenneTwister(seed)
        r0 = copy(rng)
        idx = AdvancedPS.resample_systematic(rng, w, n)
        push!(resample_records, Dict("kind" => "systematic", "case" => name, "w" => w, "n" => n,
                                     "u" => [rand(r0)], "indices" => idx))
        rng = Random.MersenneTwister(seed)
        r0 = copy(rng)
        idx = AdvancedPS.resample_stratified(rng, w, n)
        push!(resample_records, Dict("kind" => "stratified", "case" => name, "w" => w, "n" => n,
                                     "u" => [rand(r0) for _ in 1:n], "indices" => idx))
        rng = Random.MersenneTwister(seed)
        idx = AdvancedPS.resample_multinomial(rng, w, n)
        push!(resample_records, Dict("kind" => "multinomial", "case" => name, "w" => w, "n" => n,
                                     "u" => Float64[], "indices" => idx))
        try   # upstream's residual branch throws unless every n*w_j is integral (SURVEY App. B Q1)
            rng = Random.MersenneTwister(seed)
            idx = AdvancedPS.resample_residual(rng, w, n)
            push!(resample_records, Dict("kind" => "residual", "case" => name, "w" => w, "n" => n,
                                         "u" => Float64[], "indices" => idx))
        catch err
            push!(resample_records, Dict("kind" => "residual-error", "case" => name, "w" => w, "n" => n,
                                         "u" => Float64[], "indices" => Int[], "error" => sprint(showerror, err)))
        end
        rng = Random.MersenneTwister(seed)
        r0 = copy(rng)
        push!(resample_records, Dict("kind" => "randcat", "case" => name, "w" => w, "n" => 1,
                                     "u" => [rand(r0)], "indices" => [AdvancedPS.randcat(rng, w)]))
    end
end
dump("reference_resample.json", Dict("julia" => string(VERSION), "advancedps" => string(pkgversion(AdvancedPS)),
                                     "records" => resample_records))

# ---- weights: getweights / logZ / ESS through a ParticleContainer (src/container.jl:95-119)
weight_records = Dict{String,Any}[]
for m in (3, 10, 1000, 4097), spread in (0.0, 1.0, 30.0)
    lw = spread .* randn(gen, m)
    pc = AdvancedPS.ParticleContainer(Vector{Any}(undef, 0), Float64[], AdvancedPS.TracedRNG())
    pc.logWs = lw                      # vals are not touched by the three functions below
    push!(weight_records, Dict("logWs" => lw, "getweights" => AdvancedPS.getweights(pc),
                               "logZ" => AdvancedPS.logZ(pc), "ess" => AdvancedPS.effectiveSampleSize(pc)))
end
dump("reference_weights.json", Dict("records" => weight_records))

# ---- known answers the reference's tests assert
known = Dict{String,Any}()
pc = AdvancedPS.ParticleContainer(Vector{Any}(undef, 0), Float64[], AdvancedPS.TracedRNG())
pc.logWs = zeros(3)
known["uniform3"] = Dict("getweights" => AdvancedPS.getweights(pc), "logZ" => AdvancedPS.logZ(pc),
                         "ess" => AdvancedPS.effectiveSampleSize(pc))        # test/container.jl:45-49
pc.logWs = [0.0, -1.0, -2.0]
known["logps1"] = Dict("getweights" => AdvancedPS.getweights(pc), "logZ" => AdvancedPS.logZ(pc))   # :52-58
pc.logWs = 2 .* [0.0, -1.0, -2.0]
known["logps2"] = Dict("getweights" => AdvancedPS.getweights(pc), "logZ" => AdvancedPS.logZ(pc))   # :62-68
D = [0.3, 0.4, 0.3]
nD = 10^6
known["proportions"] = Dict(string(f) => count(==(2), f(Random.MersenneTwister(7), D, nD))
                            for f in (AdvancedPS.resample_systematic, AdvancedPS.resample_stratified,
                                      AdvancedPS.resample_multinomial, AdvancedPS.resample_residual))  # test/resampling.jl:12-15
known["defaults"] = Dict("SMC" => AdvancedPS.SMC(10).resampler.threshold, "PG" => AdvancedPS.PG(10).resampler.threshold,
                         "PGAS" => AdvancedPS.PGAS(10).resampler.threshold)  # src/smc.jl:15,75,99
dump("reference_known.json", known)
