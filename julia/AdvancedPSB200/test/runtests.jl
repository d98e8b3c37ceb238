# Restates, against the device path, the reference's own known-answer tests for this path
# (AdvancedPS.jl test/resampling.jl:12-15, test/container.jl:45-68, test/smc.jl:104,
# test/pgas.jl:99-127, test/linear-gaussian.jl model). Needs Julia, a B200 and libaps_b200.so;
# NOT runnable in the repository's build image (no Julia) -- the Python suite under tests/ runs
# the same checks through the same ABI.
using AdvancedPSB200
using AdvancedPS
using AbstractMCMC
using Random
using Test

@testset "AdvancedPSB200" begin
    @testset "resamplers (test/resampling.jl:12-15)" begin
        D = [0.3, 0.4, 0.3]
        n = 10^6
        rng = Random.MersenneTwister(1)
        for (f, tol) in ((gpu_resample_systematic, 1e-3), (gpu_resample_stratified, 1e-3),
                         (gpu_resample_multinomial, 1e-2), (gpu_resample_residual, 1e-2))
            idx = f(rng, D, n)
            @test length(idx) == n
            @test all(1 .<= idx .<= 3)
            @test isapprox(count(==(2), idx), 0.4n; atol=tol * n)
        end
        @test_throws ErrorException gpu_resample_systematic(rng, Float64[], 10)
    end

    @testset "weights (test/container.jl:45-68)" begin
        logps = [0.0, -1.0, -2.0]
        @test gpu_softmax(logps) ≈ exp.(logps) ./ sum(exp, logps)
        @test gpu_logsumexp(logps) ≈ log(sum(exp, logps))
        @test gpu_logsumexp(zeros(3)) ≈ log(3)
        @test gpu_ess(zeros(3)) == 3
    end

    @testset "constant likelihood: logevidence = -2 log 2 (test/smc.jl:104)" begin
        m = ConstantLogLikSSM(; Y=fill(log(0.5), 2, 1))
        s = sample(Random.MersenneTwister(1), m, AdvancedPS.SMC(100))
        @test s.logevidence ≈ -2 * log(2)
        @test sum(s.weights) ≈ 1
    end

    @testset "same seed => same trajectories (test/pgas.jl:99-127)" begin
        y = randn(Random.MersenneTwister(3), 10)
        m = LinearGaussianSSM(; Y=y)
        for smp in (AdvancedPS.PGAS(4096), AdvancedPS.PG(4096))
            a = sample(Random.MersenneTwister(10), m, smp, 3; progress=false)
            b = sample(Random.MersenneTwister(10), m, smp, 3; progress=false)
            @test all(x.trajectory.model.X == z.trajectory.model.X for (x, z) in zip(a, b))
            @test size(a[end].trajectory.model.X) == (10, 1)
        end
    end

    @testset "GPU resampler inside the unmodified reference sweep" begin
        # any AdvancedPS model works here, including Libtask / Turing ones
        @test AdvancedPS.SMC(100, gpu_resample_systematic).resampler === gpu_resample_systematic
    end
end
