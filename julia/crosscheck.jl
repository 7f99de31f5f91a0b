# crosscheck.jl -- UNEXECUTED in the build image (no Julia).  Run on a machine with Julia + the reference's
# dependencies to turn the oracle's UNVERIFIED switches (SURVEY App. A) into pass/fail:
#
#   python julia/export_fixtures.py /tmp/nf_fixtures        # replays theta of the fixtures stored as a digest (C3, 1.3 M params)
#   julia --project=/path/to/NormalizingFlows.jl/test julia/crosscheck.jl tests/golden /tmp/nf_fixtures
#
# For every golden config it rebuilds the flow with the real packages, loads theta / z0 and prints, next to the oracle's
# stored values: the ELBO, the gradient (relative difference), the forward outputs y / logdet, the spline bin indices
# (NSF configs), the forward-KL `loglikelihood` value + gradient on the fixture's y, and -- for the Hamiltonian fixture --
# the theta order of the nested `transformed(transformed(q0, .), .)` (ADVICE r1: Functors walks (dist, transform)).
using NormalizingFlows, Bijectors, Distributions, Flux, Functors, Optimisers, LinearAlgebra, Random
using Zygote, DifferentiationInterface, ADTypes
using NPZ            # ] add NPZ
using JSON
import MonotonicSplines

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
    return nothing                                   # hamiltonian: handled by hamiltonian_order_check below
end

function target(meta)
    t = meta["target"]; d = meta["dim"]
    t == "banana" && return Banana(d, 1.0, 10.0)
    t == "funnel" && return Funnel(d, 0.0, 9.0)
    t == "warped" && return WarpedGauss(1.0, 0.12)
    t == "cross"  && return product_distribution([Cross(2.0, 0.15) for _ in 1:(d ÷ 2)])   # 8 independent 2-D blocks
    t == "diag"   && return nothing                  # parameters drawn by numpy PCG64(7) in tests/helpers.py: not replayed here
    error(t)
end

# bin index of every transformed coordinate, per coupling in application order (Ls[end] first), 0 = left tail ... K+1 = right tail
function spline_bins(flow, xs)
    layers = reverse(collect_layers(flow.transform))
    out = Matrix{Int}[]
    x = xs
    for l in layers
        if l isa NormalizingFlows.NeuralSplineCoupling
            x1, x2, _ = Bijectors.partition(l.mask, x)
            pX, pY, dYdX = NormalizingFlows.get_nsc_params(l, x2)
            push!(out, [searchsortedfirst(view(pX, :, i, j), x1[i, j]) - 1 for i in axes(x1, 1), j in axes(x1, 2)])
        end
        x = first(Bijectors.with_logabsdet_jacobian(l, x))
    end
    return out
end
collect_layers(f::ComposedFunction) = vcat(collect_layers(f.outer), collect_layers(f.inner))
collect_layers(f) = Any[f]

# ADVICE r1: for `transformed(td::TransformedDistribution, t)` Bijectors composes onto the existing transform (SURVEY App. A.5),
# so theta = [theta(t); theta(t0)] -- the order julia/NormalizingFlowsNFCUDAExt.jl `unwrap` assumes.  Perturb one entry of
# the OUTER transform's first parameter and check that the FIRST entries of the destructured vector change.
function hamiltonian_order_check()
    q0 = MvNormal(zeros(4), I)
    inner = Bijectors.Shift(fill(10.0, 4)) ∘ Bijectors.Scale(fill(2.0, 4))
    outer = Bijectors.Shift(fill(-3.0, 4))
    @leaf MvNormal
    td = transformed(transformed(q0, inner), outer)
    θ, _ = Optimisers.destructure(td)
    println("nested transformed: theta = ", θ)
    println("  outer-first order (", θ[1:4] == fill(-3.0, 4) ? "CONFIRMED" : "NOT the case -- fix `unwrap` and api.hamiltonian_flow", ")")
end

aux = length(ARGS) >= 2 ? ARGS[2] : ""
for f in filter(endswith(".npz"), readdir(ARGS[1]; join=true))
    z = npzread(f)
    meta = JSON.parse(String(z["meta"]))
    flow = build(meta, Float64)
    flow === nothing && (println(basename(f), ": Hamiltonian fixture -> order check only"); hamiltonian_order_check(); continue)
    θ0, re = Optimisers.destructure(flow)
    θ = if haskey(z, "theta")
        Float64.(z["theta"])
    else
        p = joinpath(aux, replace(basename(f), ".npz" => ".theta.f32"))
        isfile(p) || (println(basename(f), ": theta stored as digest only; run julia/export_fixtures.py and pass its directory"); continue)
        Float64.(reinterpret(Float32, read(p)))
    end
    @assert length(θ) == length(θ0)
    xs = Float64.(permutedims(z["z0"]))          # [N, d] row-major file -> d×N matrix
    p = target(meta)
    fl = re(θ)
    y, ld = Bijectors.with_logabsdet_jacobian(fl.transform, xs)
    println(basename(f), ": max|y - oracle| = ", maximum(abs.(y .- permutedims(z["y"]))), "  max|logdet - oracle| = ", maximum(abs.(vec(ld) .- z["logdet"])))
    if p !== nothing
        logp = Base.Fix1(logpdf, p)
        loss(θ) = elbo_batch(re(θ), logp, xs)
        v, g = DifferentiationInterface.value_and_gradient(loss, AutoZygote(), θ)
        gd = haskey(z, "grad") ? norm(g - z["grad"]) / norm(g) : abs(norm(g) - z["grad_norm"]) / norm(g)
        println("   elbo ref=", v, " oracle=", z["elbo"], "   gradient: relative difference ", gd)
    end
    if haskey(z, "bins")
        b = spline_bins(fl, xs)
        mism = sum(sum(permutedims(b[k]) .!= z["bins"][k, :, :]) for k in eachindex(b))
        println("   spline bins: ", mism, " of ", length(z["bins"]), " differ from the oracle (RQS_BIN_RIGHT_CLOSED / layout switches)")
    end
    # forward KL on the fixture's own outputs: loglikelihood(rng, flow, y) and its gradient (src/objectives/loglikelihood.jl:26-33)
    ll(θ) = loglikelihood(Random.default_rng(), re(θ), y)
    vl, gl = DifferentiationInterface.value_and_gradient(ll, AutoZygote(), θ)
    println("   loglikelihood(y) = ", vl, "  |grad| = ", norm(gl), "   (compare: python -c 'import nf_oracle' loglik_value_and_grad on the same y)")
end
