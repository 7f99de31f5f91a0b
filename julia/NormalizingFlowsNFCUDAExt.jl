# NormalizingFlowsNFCUDAExt -- the Julia side of the drop-in (UNEXECUTED in the build image: no Julia there).
#
# What a maintainer adds to NormalizingFlows.jl to run the ELBO / log-likelihood value+gradient on
# libnfcuda behind the package's own seam:
#     _prepare_gradient / _value_and_gradient         (src/optimize.jl:8-14)
# selected by `train_flow(...; ADbackend = AutoNFCUDA(...))` (src/NormalizingFlows.jl:54-86).
# `optimize`, Optimisers.jl, callbacks, convergence checks and the destructure'd theta stay untouched.
# It supersedes ext/NormalizingFlowsCUDAExt.jl (which only sampled on the GPU, column by column).
#
# Packaging: list `libnfcuda_jll` (or a path in ENV["NFCUDA_LIB"]) as a weak dependency in Project.toml
#   [weakdeps]  NFCUDA = "..."        [extensions]  NormalizingFlowsNFCUDAExt = "NFCUDA"
module NormalizingFlowsNFCUDAExt

using NormalizingFlows
using NormalizingFlows: Bijectors, Distributions, Optimisers, ADTypes, Random
using LinearAlgebra: cholesky, Symmetric, isdiag, diag
import NormalizingFlows: _prepare_gradient, _value_and_gradient, _device_specific_rand

const libnfcuda = get(ENV, "NFCUDA_LIB", "libnfcuda")

# ---- C structs of include/nfcuda.h -----------------------------------------------------------------
struct NFLayerDesc
    kind::Cint
    mask_idx::Ptr{Cint}
    n_mask::Cint
    hdims::Ptr{Cint}
    n_hidden::Cint
    K::Cint
    B::Cdouble
    n_steps::Cint              # leapfrog steps (example/demo_hamiltonian_flow.jl:30)
    score_target::Ptr{Cvoid}   # nf_target_t whose score drives the leapfrog dynamics
end
NFLayerDesc(kind, mi, nm, hd, nh, K, B) = NFLayerDesc(kind, mi, nm, hd, nh, K, B, 0, C_NULL)
const NF_F32, NF_F64 = Cint(0), Cint(1)
const NF_PLANAR, NF_RADIAL, NF_AFFINE, NF_SPLINE, NF_SHIFT, NF_SCALE, NF_MOMENTUM_AFFINE, NF_LEAPFROG = Cint.(1:8)

function check(status::Cint)
    status == 0 && return nothing
    error("libnfcuda error $status: " * unsafe_string(ccall((:nf_last_error, libnfcuda), Cstring, ())))
end

# ---- the ADTypes backend tag -------------------------------------------------------------------------
"""
    AutoNFCUDA(; device = 0, devices = nothing, target, verify_logp = true)

`devices = 0:7` runs every objective evaluation data-parallel on those GPUs from this one Julia process (`nf_comm_init_all`,
one flow / target replica per device, `nf_elbo_value_and_grad_multi`: samples sharded, ONE ncclAllReduce of the P+1 sums
inside libnfcuda); `device` is the single-GPU form.  With `verify_logp` (default) the `logp` closure handed to `train_flow`
is evaluated at a few points and compared with the device target named by `target` -- a mismatch is an error, not a silent
wrong density.

`target` names a device-side log-density: `(:banana, b, var)`, `(:funnel, μ, σ)`, `(:warped_gauss, σ1, σ2)`,
`(:cross, μ, σ)`, `(:diag_normal, μ, σ)`, `(:logreg, σ₀, n, X, y)`; `(:joint, kind, params...)` wraps any of them with N(0, I) momenta.  The Julia `logp` closure handed to `train_flow` is used only
to recognise the example targets; an arbitrary closure cannot cross the C ABI (use `nf_forward_stash` /
`nf_backward` with CUDA.jl evaluating logp and its score on the device buffer).
"""
struct AutoNFCUDA <: ADTypes.AbstractADType
    devices::Vector{Int}
    target::Tuple
    verify_logp::Bool
end
AutoNFCUDA(; device=0, devices=nothing, target, verify_logp=true) =
    AutoNFCUDA(devices === nothing ? [device] : collect(Int, devices), target, verify_logp)
AutoNFCUDA(device::Int, target::Tuple) = AutoNFCUDA([device], target, false)
export AutoNFCUDA

# ---- flow structure -> nf_layer_desc[] in theta (= Ls) order ------------------------------------------
flatten_layers(f::ComposedFunction) = vcat(flatten_layers(f.outer), flatten_layers(f.inner))
flatten_layers(f) = Any[f]
# `transformed(transformed(q₀, t₀), t)` (the Hamiltonian demo, example/demo_hamiltonian_flow.jl:139-147): t's layers come
# first in θ, then t₀'s; the innermost distribution is the base.
function unwrap(flow::Bijectors.TransformedDistribution)
    layers = flatten_layers(flow.transform)
    base = flow.dist
    while base isa Bijectors.TransformedDistribution
        append!(layers, flatten_layers(base.transform))
        base = base.dist
    end
    return layers, base
end

hidden_dims(c) = Cint[size(l.weight, 1) for l in c.layers[1:(end - 1)]]

function describe(layer, keep::Vector{Any}; score_target::Ptr{Cvoid}=C_NULL)
    z = Ptr{Cint}(C_NULL)
    if layer isa Bijectors.PlanarLayer
        return NFLayerDesc(NF_PLANAR, z, 0, z, 0, 0, 0.0)
    elseif layer isa Bijectors.RadialLayer
        return NFLayerDesc(NF_RADIAL, z, 0, z, 0, 0, 0.0)
    elseif layer isa Bijectors.Shift
        return NFLayerDesc(NF_SHIFT, z, 0, z, 0, 0, 0.0)
    elseif layer isa Bijectors.Scale
        return NFLayerDesc(NF_SCALE, z, 0, z, 0, 0, 0.0)
    elseif layer isa NormalizingFlows.AffineCoupling
        idx = Cint.(findall(!iszero, vec(sum(layer.mask.A_1; dims=2))) .- 1)   # PartitionMask transformed rows, 0-based
        hd = hidden_dims(layer.s)
        push!(keep, idx, hd)
        return NFLayerDesc(NF_AFFINE, pointer(idx), length(idx), pointer(hd), length(hd), 0, 0.0)
    elseif layer isa NormalizingFlows.NeuralSplineCoupling
        idx = Cint.(findall(!iszero, vec(sum(layer.mask.A_1; dims=2))) .- 1)
        hd = hidden_dims(layer.nn)
        push!(keep, idx, hd)
        return NFLayerDesc(NF_SPLINE, pointer(idx), length(idx), pointer(hd), length(hd), layer.K, Float64(layer.B))
    elseif layer isa Bijectors.Stacked && length(layer.bs) == 2 && layer.bs[1] === identity
        # momentum_normalization_layer of example/demo_hamiltonian_flow.jl:94-99: Stacked((identity, Shift ∘ Scale), ...)
        return NFLayerDesc(NF_MOMENTUM_AFFINE, z, 0, z, 0, 0, 0.0)
    elseif hasproperty(layer, :logϵ) && hasproperty(layer, :L) && hasproperty(layer, :∇logp)
        # the demo's LeapFrog bijector (:27-47); its ∇logp closure is replaced by the device score of `score_target`
        score_target == C_NULL && error("NFCUDA: a LeapFrog layer needs target = (:joint, kind, params...)")
        return NFLayerDesc(NF_LEAPFROG, z, 0, z, 0, 0, 0.0, Cint(layer.L), score_target)
    end
    error("NFCUDA: unsupported bijector $(typeof(layer))")
end

mutable struct Prep
    flow::Ptr{Cvoid}             # replica on the first device (single-GPU calls use it)
    target::Ptr{Cvoid}
    P::Int
    T::DataType
    comm::Ptr{Cvoid}             # nf_comm_t over ad.devices (C_NULL for one device)
    flows::Vector{Ptr{Cvoid}}    # one replica per device, in communicator order
    targets::Vector{Ptr{Cvoid}}
    verified::Bool
end

target_kind(s::Symbol) = Dict(:banana => 1, :funnel => 2, :warped_gauss => 3, :cross => 4, :diag_normal => 5, :logreg => 6)[s]
# (:logreg, σ₀, n, vec(X'), y): X' is the n×dim design matrix flattened row by row, i.e. vec of the dim×n Julia matrix

function make_prep(flow::Bijectors.TransformedDistribution, ad::AutoNFCUDA, ::Type{T}) where {T}
    reps = [make_replica(flow, ad, T, dev) for dev in ad.devices]
    comm = Ref{Ptr{Cvoid}}(C_NULL)
    if length(ad.devices) > 1
        check(ccall((:nf_comm_init_all, libnfcuda), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), comm, length(ad.devices), Cint.(ad.devices)))
    end
    P = ccall((:nf_flow_num_params, libnfcuda), Int64, (Ptr{Cvoid},), reps[1][1])
    prep = Prep(reps[1][1], reps[1][2], P, T, comm[], first.(reps), last.(reps), false)
    finalizer(prep) do p
        p.comm == C_NULL || ccall((:nf_comm_destroy, libnfcuda), Cvoid, (Ptr{Cvoid},), p.comm)
        foreach(h -> ccall((:nf_flow_destroy, libnfcuda), Cvoid, (Ptr{Cvoid},), h), p.flows)
        foreach(h -> ccall((:nf_target_destroy, libnfcuda), Cvoid, (Ptr{Cvoid},), h), p.targets)
    end
    return prep
end

# one (flow, target) replica on device `dev`
function make_replica(flow::Bijectors.TransformedDistribution, ad::AutoNFCUDA, ::Type{T}, dev::Int) where {T}
    check(ccall((:nf_init, libnfcuda), Cint, (Cint,), dev))
    keep = Any[]
    layers, base = unwrap(flow)
    d = length(base)
    # targets first: a Hamiltonian flow's LeapFrog layers need the inner target's handle
    joint = ad.target[1] === :joint                      # (:joint, :funnel, μ, σ): logp(x) + logN(ρ) on z = [x, ρ]
    spec = joint ? ad.target[2:end] : ad.target
    tp = Float64[x for p in spec[2:end] for x in p]
    t = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:nf_target_create, libnfcuda), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Cint),
        t, target_kind(spec[1]), joint ? d ÷ 2 : d, tp, length(tp)))
    inner = joint ? t[] : Ptr{Cvoid}(C_NULL)
    if joint
        tj = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:nf_target_create_joint, libnfcuda), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}), tj, inner))
        t[] = tj[]
    end
    descs = [describe(l, keep; score_target=inner) for l in layers]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep descs check(ccall((:nf_flow_create, libnfcuda), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{NFLayerDesc}, Cint, Cint, Cint), h, descs, length(descs), d, T === Float32 ? NF_F32 : NF_F64))
    μ = Float64.(Distributions.mean(base))
    Σ = Matrix{Float64}(Distributions.cov(base))
    if isdiag(Σ)
        σ = sqrt.(diag(Σ))
        check(ccall((:nf_flow_set_base, libnfcuda), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h[], μ, σ))
    else
        # full covariance (ext/NormalizingFlowsCUDAExt.jl:43-47): the column-major memory of U = L' IS the row-major L
        U = Matrix{Float64}(cholesky(Symmetric(Σ)).U)
        check(ccall((:nf_flow_set_base_chol, libnfcuda), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h[], μ, U))
    end
    joint && ccall((:nf_target_destroy, libnfcuda), Cvoid, (Ptr{Cvoid},), inner)   # both users hold copies
    return (h[], t[])
end

# The Julia `logp` argument cannot cross the C ABI; what runs is the device target named in `ad.target`.  Before the first
# evaluation, compare the two at a few points (device side: nf_target_logp) and refuse to train against a different density.
function verify_logp!(prep::Prep, ad::AutoNFCUDA, logp, d::Int)
    (prep.verified || !ad.verify_logp || ad.target[1] === :joint) && return
    rng = Random.Xoshiro(0x5eed)
    xs = 0.7 .* randn(rng, Float64, d, 8)
    ref = [logp(xs[:, j]) for j in 1:8]
    dev = Vector{Float64}(undef, 8)
    check(ccall((:nf_target_logp, libnfcuda), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cvoid}),
        prep.target, NF_F64, xs, 8, dev, C_NULL))
    all(isapprox.(ref, dev; rtol=1e-8, atol=1e-8)) ||
        error("NFCUDA: the `logp` passed to train_flow is not the device target $(ad.target) (logp = $ref, device = $dev)")
    prep.verified = true
end

# `loss` is the closure of src/NormalizingFlows.jl:69: loss(θ, rng, args...) = -vo(rng, re(θ), args...)
# its captured fields give the objective and the Restructure.
function _prepare_gradient(loss, ad::AutoNFCUDA, θ::AbstractVector{T}, rng, args...) where {T}
    flow = loss.re(θ)
    prep = make_prep(flow, ad, T)
    prep.P == length(θ) || error("NFCUDA: parameter count mismatch ($(prep.P) vs $(length(θ)))")
    return prep
end

function _value_and_gradient(loss, prep::Prep, ad::AutoNFCUDA, θ::AbstractVector{T}, rng, args...) where {T}
    g = similar(θ)
    val = Ref{Cdouble}(0)
    vo = loss.vo
    if vo === NormalizingFlows.elbo || vo === NormalizingFlows.elbo_batch
        logp, n_or_xs = args
        verify_logp!(prep, ad, logp, ccall((:nf_flow_dim, libnfcuda), Cint, (Ptr{Cvoid},), prep.flow))
        if prep.comm != C_NULL                  # data parallel over ad.devices: shards + all-reduce inside libnfcuda
            N = n_or_xs isa Integer ? n_or_xs : size(n_or_xs, 2)
            xs = n_or_xs isa Integer ? Ptr{T}(C_NULL) : Matrix{T}(n_or_xs)
            seed = n_or_xs isa Integer ? rand(rng, UInt64) : UInt64(0)
            GC.@preserve xs check(ccall((:nf_elbo_value_and_grad_multi, libnfcuda), Cint,
                (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{T}, Int64, Ptr{T}, UInt64, Cdouble, Ref{Cdouble}, Ptr{T}),
                prep.comm, prep.flows, prep.targets, θ, N, xs isa Ptr ? xs : pointer(xs), seed, -1.0, val, g))
        elseif n_or_xs isa Integer              # elbo([rng,] flow, logp, n): draws on the device (Philox)
            seed = rand(rng, UInt64)
            check(ccall((:nf_elbo_value_and_grad, libnfcuda), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, Int64, Ptr{T}, UInt64, Cdouble, Ref{Cdouble}, Ptr{T}),
                prep.flow, prep.target, θ, n_or_xs, C_NULL, seed, -1.0, val, g))
        else                                    # elbo(flow, logp, xs::AbstractMatrix): host-supplied d×N draws
            xs = Matrix{T}(n_or_xs)
            check(ccall((:nf_elbo_value_and_grad, libnfcuda), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, Int64, Ptr{T}, UInt64, Cdouble, Ref{Cdouble}, Ptr{T}),
                prep.flow, prep.target, θ, size(xs, 2), xs, 0, -1.0, val, g))
        end
    elseif vo === NormalizingFlows.loglikelihood && prep.comm != C_NULL
        xs = Matrix{T}(args[1])
        check(ccall((:nf_loglik_value_and_grad_multi, libnfcuda), Cint,
            (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{T}, Int64, Ptr{T}, Cdouble, Ref{Cdouble}, Ptr{T}),
            prep.comm, prep.flows, θ, size(xs, 2), xs, -1.0, val, g))
    elseif vo === NormalizingFlows.loglikelihood
        xs = Matrix{T}(args[1])
        check(ccall((:nf_loglik_value_and_grad, libnfcuda), Cint,
            (Ptr{Cvoid}, Ptr{T}, Int64, Ptr{T}, Cdouble, Ref{Cdouble}, Ptr{T}),
            prep.flow, θ, size(xs, 2), xs, -1.0, val, g))
    else
        error("NFCUDA: objective must be elbo, elbo_batch or loglikelihood")
    end
    return T(val[]), g                           # (loss, gradient) exactly as DI.value_and_gradient returns them
end

# Batched replacements for the per-column loop of ext/NormalizingFlowsCUDAExt.jl:65-74 ---------------------
struct NFCUDARNG <: Random.AbstractRNG
    seed::UInt64
end
function _device_specific_rand(rng::NFCUDARNG, td::Bijectors.TransformedDistribution, n::Int)
    θ, _ = Optimisers.destructure(td)
    T = eltype(θ)
    prep = make_prep(td, AutoNFCUDA(0, (:diag_normal, zeros(length(td.dist)), ones(length(td.dist)))), T)
    ys = Matrix{T}(undef, length(td.dist), n)
    check(ccall((:nf_sample, libnfcuda), Cint, (Ptr{Cvoid}, Ptr{T}, Int64, UInt64, Ptr{T}), prep.flow, θ, n, rng.seed, ys))
    return ys
end

end # module
