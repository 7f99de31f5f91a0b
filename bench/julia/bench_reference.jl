# bench_reference.jl -- the TRUE CPU baseline (UNEXECUTED here: no Julia in the image or on the GPU box).
#   julia -t $(nproc) --project=/path/to/NormalizingFlows.jl/example bench/julia/bench_reference.jl [N]
# Times DifferentiationInterface.value_and_gradient of θ -> -elbo_batch(re(θ), logp, Z0) for BASELINE config 3
# (realnvp(q0, [256,256], 4) on Funnel(64), Float32) with AutoMooncake, as example/demo_RealNVP.jl:35-46 does.
using NormalizingFlows, Bijectors, Distributions, Flux, Functors, Optimisers, LinearAlgebra, Random
using Mooncake, DifferentiationInterface, ADTypes, BenchmarkTools
include(joinpath(@__DIR__, "..", "..", "..", "NormalizingFlows.jl", "example", "SyntheticTargets.jl"))
N = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 2^14
T = Float32
q0 = MvNormal(zeros(T, 64), I); @leaf MvNormal
flow = realnvp(q0, [256, 256], 4; paramtype=T)
θ, re = Optimisers.destructure(flow)
p = Funnel(64, T(0), T(9)); logp = Base.Fix1(logpdf, p)
Z0 = randn(Xoshiro(2024), T, 64, N)
loss(θ) = -elbo_batch(re(θ), logp, Z0)
ad = AutoMooncake(; config=Mooncake.Config())
prep = DifferentiationInterface.prepare_gradient(loss, ad, θ)
t = @belapsed DifferentiationInterface.value_and_gradient($loss, $prep, $ad, $θ)
println("samples/s = ", N / t, "  threads = ", Threads.nthreads(), "  BLAS threads = ", BLAS.get_num_threads())
