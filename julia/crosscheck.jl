# crosscheck.jl -- UNEXECUTED in the build image (no Julia).  Run on a machine with Julia + the reference's
# dependencies to turn the oracle's UNVERIFIED switches (SURVEY App. A) into pass/fail:
#   julia --project=/path/to/NormalizingFlows.jl/test julia/crosscheck.jl tests/golden
# It rebuilds every golden config with the real packages, loads theta / z0 from the .npz fixtures and prints the
# reference's ELBO, gradient norm and spline bins next to the oracle's stored values.
using NormalizingFlows, Bijectors, Distributions, Flux, Functors, Optimisers, LinearAlgebra, Random
using Zygote, DifferentiationInterface, ADTypes
using NPZ            # ] add NPZ
using JSON

include(joinpath(@__DIR__, "..", "..", "NormalizingFlows.jl", "example", "SyntheticTargets.jl"))  # adjust to your checkout

function build(meta, T)
    d = meta["dim"]; kw = meta["kw"]
    q0 = MvNormal(zeros(T, d), I)
    @leaf MvNormal
    k = meta["kind"]
    k == "planar"  && return planarflow(q0, kw["nlayers"]; paramtype=T)
    k == "radial"  && return radialflow(q0, kw["nlayers"]; paramtype=T)
    k == "realnvp" && return realnvp(q0, kw["hdims"], kw["nlayers"]; paramtype=T)
    k == "nsf"     && return nsf(q0, kw["hdims"], kw["K"], T(kw["B"]), kw["nlayers"]; paramtype=T)
    error(k)
end

function target(meta)
    t = meta["target"]; d = meta["dim"]
    t == "banana" && return Banana(d, 1.0, 10.0)
    t == "funnel" && return Funnel(d, 0.0, 9.0)
    t == "warped" && return WarpedGauss(1.0, 0.12)
    t == "cross"  && return product_distribution([Cross(2.0, 0.15) for _ in 1:(d ÷ 2)])   # 8 independent 2-D blocks
    error(t)
end

for f in filter(endswith(".npz"), readdir(ARGS[1]; join=true))
    z = npzread(f)
    haskey(z, "theta") || (println(basename(f), ": theta stored as digest only, regenerate with make_golden.py"); continue)
    meta = JSON.parse(String(z["meta"]))
    flow = build(meta, Float64)
    θ0, re = Optimisers.destructure(flow)
    θ = Float64.(z["theta"]); @assert length(θ) == length(θ0)
    xs = Float64.(permutedims(z["z0"]))          # [N, d] row-major file -> d×N matrix
    p = target(meta); logp = Base.Fix1(logpdf, p)
    loss(θ) = elbo_batch(re(θ), logp, xs)
    v, g = DifferentiationInterface.value_and_gradient(loss, AutoZygote(), θ)
    println(basename(f), ": elbo ref=", v, " oracle=", z["elbo"], "  |g| ref=", norm(g),
            haskey(z, "grad") ? "  rel grad diff=$(norm(g - z["grad"]) / norm(g))" : "")
end
