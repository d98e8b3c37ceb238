# Times the REAL reference sweep (AdvancedPS.jl, CPU, single thread) on BASELINE.json configs[1]
# scaled to what one core finishes in seconds. For anyone with Julia -- the repository's build image
# has none, so bench.py's `--impl reference` arm times the C++ oracle port instead (SURVEY 8d).
#
#   julia --project=. bench/julia/reference_sweep.jl [N] [T]
using AdvancedPS, SSMProblems, Distributions, Random, AbstractMCMC

N = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 10_000
T = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 100

# the model of test/linear-gaussian.jl:59-94 (noise parameters are standard deviations)
struct Prior <: SSMProblems.StatePrior end
struct Dyn <: SSMProblems.LatentDynamics end
struct Obs <: SSMProblems.ObservationProcess end
SSMProblems.distribution(::Prior) = Normal(0.0, 1.0)
SSMProblems.distribution(::Dyn, ::Int, x) = Normal(0.5 * x + 0.2, 0.1)
SSMProblems.distribution(::Obs, ::Int, x) = Normal(1.0 * x, 0.1)

rng = Random.MersenneTwister(1234)
ssm = SSMProblems.StateSpaceModel(Prior(), Dyn(), Obs())
_, _, ys = sample(rng, ssm, T)
model = AdvancedPS.TracedSSM(ssm, ys)
smc = AdvancedPS.SMC(N, AdvancedPS.resample_systematic)      # bare function: resample at every step

sample(rng, model, smc)                                       # compile
t = @elapsed s = sample(rng, model, smc)
println("reference sweep: N=$N T=$T  $(round(t; digits=3)) s  ->  $(round(N * T / t; sigdigits=4)) particle-steps/s  ",
        "(1 thread of $(Sys.CPU_THREADS)); logevidence = $(s.logevidence)")
