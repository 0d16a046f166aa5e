# AGPB200.jl -- reference-side binding of libagp_b200.so (UNVERIFIED: Julia is not installed in the build image).
#
# Overrides the methods of AugmentedGaussianProcesses.jl that the engine replaces,
#     update_parameters!(model::SVGP, state, x, y)            (src/training/training.jl:140-144)
#     update_parameters!(model::MOSVGP, state, x, ys)         (src/training/training.jl:153-158, incl. update_A!)
# plus ELBO(model, state, y) (src/inference/analyticVI.jl:255-297, single- and multi-output) and update_hyperparameters!
# (src/hyperparameter/autotuning.jl:86-140), for models built with AnalyticVI / AnalyticSVI.
# Data crosses the ABI exactly as the reference holds it: X as the column-major parent Matrix{Float64} of the
# RowVecs (src/data/datacontainer.jl:64-66), minibatches as 1-based Vector{Int}, y as Vector{Float64}.
module AGPB200

using AugmentedGaussianProcesses
const AGP = AugmentedGaussianProcesses
using KernelFunctions

const LIB = get(ENV, "AGP_B200_LIB", "libagp_b200.so")

struct ModelDesc  # mirrors agp_model_desc (include/agp_b200.h)
    model_kind::Int32; n_latent_global::Int32; latent_begin::Int32; n_latent_local::Int32
    m::Int32; D::Int32; batch_capacity::Int32; precision::Int32; stochastic::Int32
    rm_kappa::Float64; rm_tau::Float64; jitter::Float64
    n_task::Int32
    lik_kind::Ptr{Int32}; lik_p0::Ptr{Float64}; lik_p1::Ptr{Float64}; A::Ptr{Float64}
    kernel_kind::Ptr{Int32}; kernel_scale::Ptr{Float64}; kernel_variance::Ptr{Float64}
    Z::Ptr{Float64}; mu0::Ptr{Float64}
end

mutable struct Engine
    ctx::Ptr{Cvoid}
    model::Ptr{Cvoid}
    uploaded::UInt   # objectid of the data parent currently resident
    fresh::Bool      # no step taken yet (the conditioning policy may still rebuild the engine at another precision)
end

function check(e::Engine, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:agp_last_error, LIB), Cstring, (Ptr{Cvoid},), e.ctx))
    rc == 3 && error("K̃ has negative values")                       # gpblocks/latentgp.jl:213
    rc == 4 && throw(LinearAlgebra.PosDefException(0))                # cholesky()
    error(msg)
end

const ENGINES = IdDict{Any,Engine}()
function release!(model)
    e = pop!(ENGINES, model)
    ccall((:agp_model_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.model)
    ccall((:agp_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.ctx)
    return nothing
end

lik_code(::AGP.GaussianLikelihood) = (Int32(0))
lik_code(::AGP.BernoulliLikelihood{<:AGP.LogisticLink}) = Int32(1)
lik_code(::AGP.StudentTLikelihood) = Int32(2)
lik_code(::AGP.MultiClassLikelihood{<:AGP.LogisticSoftMaxLink}) = Int32(3)
lik_code(::AGP.LaplaceLikelihood) = Int32(4)
lik_code(::AGP.BernoulliLikelihood{<:AGP.SVMLink}) = Int32(5)
lik_code(::AGP.NegBinomialLikelihood) = Int32(6)
lik_code(::AGP.PoissonLikelihood{<:AGP.ScaledLogistic}) = Int32(7)
lik_code(::AGP.HeteroscedasticGaussianLikelihood{<:AGP.InvScaledLogistic}) = Int32(8)
# first likelihood parameter crossing the ABI (include/agp_b200.h: AGP_LIK_* comments)
lik_p0(l) = 0.0
lik_p0(l::AGP.GaussianLikelihood) = Float64(AGP.noise(l))
lik_p0(l::AGP.StudentTLikelihood) = Float64(l.ν)
lik_p0(l::AGP.LaplaceLikelihood) = Float64(l.β)
lik_p0(l::AGP.NegBinomialLikelihood) = Float64(l.r)
lik_p0(l::Union{AGP.PoissonLikelihood,AGP.HeteroscedasticGaussianLikelihood}) = Float64(only(l.invlink.λ))

# kernel -> (kind, scale, variance); supports [σ² *] {SqExponential, Matern32, Matern52} [∘ ScaleTransform(s)]
function kernel_params(k)
    var = 1.0; s = 1.0
    if k isa ScaledKernel; var = only(k.σ²); k = k.kernel; end
    if k isa TransformedKernel; s = only(k.transform.s); k = k.kernel; end
    kind = k isa SqExponentialKernel ? 0 : k isa Matern32Kernel ? 1 : k isa Matern52Kernel ? 2 : error("kernel not supported by AGPB200")
    return Int32(kind), Float64(s), Float64(var)
end

# tcgen05 (tf32x3 = 2) needs a batch CAPACITY that is a multiple of 128 and more than 64 inducing points (the engine pads m and a ragged
# minibatch of the host index list path itself); small models (the reference's own tests use 10 inducing points,
# test/testingtools.jl:66) take the fp32 SIMT path (1).  AGP_B200_PRECISION = f64 | f32 | tf32x3 overrides.
function precision_code(m::Int, B::Int, model = nothing)
    p = get(ENV, "AGP_B200_PRECISION", "auto")
    p == "f64" && return Int32(0); p == "f32" && return Int32(1); p == "tf32x3" && return Int32(2)
    haskey(PRECISION_BY_CONDITION, model) && return PRECISION_BY_CONDITION[model]
    return m >= 128 ? Int32(2) : Int32(1)
end
batch_capacity(m::Int, B::Int, model = nothing) = precision_code(m, B, model) == 2 ? cld(B, 128) * 128 : B

# The same conditioning policy as the Python host layer (api.AMPLIFICATION_LIMIT, DESIGN section 3): after the first agp_refresh_K the
# error amplification sqrt(variance ||K_mm^-1||_inf) decides whether the model keeps the tcgen05 path (<= 30), moves to the fp32
# CUDA-core path (<= 100) or to the fp64 path.  Returns true when the engine has to be rebuilt.
const PRECISION_BY_CONDITION = IdDict{Any,Int32}()
function amplification(model, e)
    m = AGP.dim(model.f[1]); Kinv = zeros(m, m); ld = Ref(0.0); amp = 0.0
    for (q, gp) in enumerate(model.f)
        check(e, ccall((:agp_get_Kinv, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Float64}), e.model, q - 1, Kinv, ld))
        amp = max(amp, sqrt(kernel_params(AGP.kernel(gp))[3] * maximum(sum(abs, Kinv; dims = 1))))
    end
    return amp
end
function conditioning_switch!(model, e, m::Int, B::Int)
    get(ENV, "AGP_B200_PRECISION", "auto") == "auto" || return false
    cur = precision_code(m, B, model)
    cur == 0 && return false
    amp = amplification(model, e)
    want = amp <= 30 && cur == 2 ? Int32(2) : amp <= 100 ? Int32(1) : Int32(0)
    want == cur && return false
    @warn "AGPB200: K_mm is ill conditioned (error amplification $amp): moving the model to precision code $want (0 = f64, 1 = f32)"
    PRECISION_BY_CONDITION[model] = want
    return true
end

likelihoods(model::SVGP) = (AGP.likelihood(model),)
likelihoods(model::MOSVGP) = Tuple(AGP.likelihood(model))
model_kind(::SVGP) = Int32(0)
model_kind(::MOSVGP) = Int32(1)
# MOSVGP mixing matrix: A[task][1] = weights over the Q latents (single-latent task likelihoods) -> row-major [T][Q]
mixing(::SVGP) = Float64[]
mixing(model::MOSVGP) = Float64[model.A[t][1][q] for q in 1:length(model.f), t in 1:length(model.A)][:]   # (Q, T) column-major == [T][Q] row-major

function engine(model::Union{SVGP{T},MOSVGP{T}}, B::Int) where {T}
    haskey(ENGINES, model) && return ENGINES[model]
    inf = AGP.inference(model); ls = likelihoods(model)
    Q = length(model.f); m = AGP.dim(model.f[1]); D = length(first(model.f[1].Z))
    model isa MOSVGP && any(>(1), model.nf_per_task) && error("AGPB200: MOSVGP tasks must be single-latent likelihoods")
    Z = zeros(Float64, D, m, Q)                       # row-major [Q][m][D] for C == column-major (D, m, Q) here
    for (q, gp) in enumerate(model.f), (i, z) in enumerate(gp.Z); Z[:, i, q] .= z; end
    kp = [kernel_params(AGP.kernel(gp)) for gp in model.f]
    kk = Int32[p[1] for p in kp]; ks = Float64[p[2] for p in kp]; kv = Float64[p[3] for p in kp]
    lk = Int32[lik_code(l) for l in ls]
    p0 = Float64[lik_p0(l) for l in ls]
    p1 = Float64[l isa AGP.StudentTLikelihood ? l.σ : 0.0 for l in ls]
    A = mixing(model)
    o = AGP.opt(inf).optimiser
    κ, τ = o isa RobbinsMonro ? (Float64(o.κ), Float64(o.τ)) : (0.51, 1.0)
    ctx = Ref{Ptr{Cvoid}}(C_NULL); mdl = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:agp_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), 0, C_NULL, ctx)
    rc == 0 || error("agp_ctx_create failed (no CUDA device; there is no CPU fallback)")
    e = Engine(ctx[], C_NULL, 0, true)
    GC.@preserve Z kk ks kv lk p0 p1 A begin
        d = Ref(ModelDesc(model_kind(model), Q, 0, Q, m, D, batch_capacity(m, B, model), precision_code(m, B, model), AGP.is_stochastic(inf) ? 1 : 0, κ, τ, Float64(T(AGP.jitt)),
                          length(ls), pointer(lk), pointer(p0), pointer(p1), isempty(A) ? C_NULL : pointer(A),
                          pointer(kk), pointer(ks), pointer(kv), pointer(Z), C_NULL))
        check(e, ccall((:agp_model_create, LIB), Cint, (Ptr{Cvoid}, Ref{ModelDesc}, Ref{Ptr{Cvoid}}), e.ctx, d, mdl))
    end
    e.model = mdl[]
    # quirk Q3: the reference never re-raises HPupdated after update_hyperparameters! (autotuning.jl:45 is commented out), so K_mm
    # stays the one factorised at the start of train! while K_nm follows the new kernel / Z
    check(e, ccall((:agp_keep_stale_K, LIB), Cint, (Ptr{Cvoid}, Int32), e.model, 1))
    if any(l -> l isa AGP.PoissonLikelihood, ls)      # `expectation` rule of functions/utils.jl:16-19 (pred_nodes, pred_weights of predictions.jl:4)
        nodes = collect(Float64, AGP.pred_nodes); w = collect(Float64, AGP.pred_weights)
        GC.@preserve nodes w check(e, ccall((:agp_set_quadrature, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32),
                                            e.model, nodes, w, length(nodes)))
    end
    if model isa MOSVGP && !isnothing(model.A_opt)    # update_A! with ADAM (single_and_multi_output_utils.jl:87-118, states.jl:100-105)
        ao = model.A_opt
        check(e, ccall((:agp_set_A_optimiser, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Float64),
                       e.model, 1, Float64(ao.eta), Float64(ao.beta[1]), Float64(ao.beta[2]), Float64(ao.epsilon)))
    end
    ENGINES[model] = e
    return e
end

# ---- the override -------------------------------------------------------------------------------------
function device_step!(model, state, x::SubArray, ys::Tuple)
    rows = x.parent                                   # RowVecs over the n×D Matrix
    idx = Int64.(x.indices[1])                        # 1-based minibatch (training.jl:51-55)
    B = length(idx)
    e = engine(model, AGP.batchsize(AGP.inference(model)))
    if e.uploaded != objectid(rows.X)
        yall = map(parent, ys)                        # one target vector per task, whole data set
        yp = Ptr{Cvoid}[pointer(v) for v in yall]
        GC.@preserve yall yp check(e, ccall((:agp_data_upload, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Ptr{Cvoid}}, Cint),
            e.model, rows.X, 0 #= f64 =#, 0 #= column-major =#, size(rows.X, 1), yp, first(yall) isa AbstractVector{Float64} ? 0 : 1))
        e.uploaded = objectid(rows.X)
    end
    if AGP.isHPupdated(AGP.inference(model))
        check(e, ccall((:agp_refresh_K, LIB), Cint, (Ptr{Cvoid},), e.model))      # compute_K
        if e.fresh && conditioning_switch!(model, e, AGP.dim(model.f[1]), AGP.batchsize(AGP.inference(model)))
            release!(model)                           # rebuild on the slower, more accurate path and start this step again
            return device_step!(model, state, x, ys)
        end
        AGP.setHPupdated!(AGP.inference(model), false)
    end
    e.fresh = false
    ρ = Float64(AGP.ρ(AGP.inference(model)))
    GC.@preserve idx check(e, ccall((:agp_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int32, Int32, Float64),
                                    e.model, idx, B, 1, ρ))
    pull_posterior!(model, e); pull_lik_params!(model, e)                          # keep model.f[k].post in sync for Julia-side consumers
    return e
end

function AGP.update_parameters!(model::SVGP{T,L,<:AnalyticVI}, state, x::SubArray, y) where {T,L}
    device_step!(model, state, x, (y,))
    return state
end

# training.jl:153-158: compute_kernel_matrices, update_A!, variational_updates -- all inside agp_step (the A optimiser was
# registered with agp_set_A_optimiser); the mixing weights are mirrored back into model.A
function AGP.update_parameters!(model::MOSVGP{T,L,<:AnalyticVI}, state, x::SubArray, ys) where {T,L}
    e = device_step!(model, state, x, Tuple(ys))
    if !isnothing(model.A_opt)
        Q = length(model.f); nt = length(model.A); A = zeros(Q, nt)              # [T][Q] row-major == (Q, T) column-major
        check(e, ccall((:agp_get_A, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.model, A))
        for t in 1:nt; model.A[t][1] .= A[:, t]; end
    end
    return state
end

# λ of PoissonLikelihood / HeteroscedasticLikelihood is re-estimated on the device by every step (poisson.jl:80,
# heteroscedastic.jl:98): mirror it back into l.invlink.λ
function pull_lik_params!(model, e::Engine)
    for (t, l) in enumerate(likelihoods(model))
        if l isa Union{AGP.PoissonLikelihood,AGP.HeteroscedasticGaussianLikelihood}
            v = Ref{Float64}(0.0)
            check(e, ccall((:agp_get_lik_param, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), e.model, t - 1, v))
            l.invlink.λ .= v[]
        end
    end
end

function pull_posterior!(model, e::Engine)
    for (q, gp) in enumerate(model.f)
        m = AGP.dim(gp); μ = zeros(m); Σ = zeros(m, m); η₁ = zeros(m); η₂ = zeros(m, m)
        check(e, ccall((:agp_get_posterior, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       e.model, q - 1, μ, Σ, η₁, η₂))
        gp.post.μ .= μ; gp.post.Σ.data .= Σ; gp.post.η₁ .= η₁; gp.post.η₂.data .= η₂   # symmetric: row/column-major agree
    end
end

# analyticVI.jl:255-274 (single output) and :277-297 (multi-output: the sum over the tasks is formed on the device)
function AGP.ELBO(model::Union{SVGP{T,L,<:AnalyticVI},MOSVGP{T,L,<:AnalyticVI}}, state::NamedTuple, y) where {T,L}
    e = ENGINES[model]; out = zeros(3)
    check(e, ccall((:agp_elbo, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), e.model, Float64(AGP.ρ(AGP.inference(model))), out))
    return out[1] - out[2] - out[3]
end

# update_hyperparameters! (hyperparameter/autotuning.jl:86-140): the Zygote call is replaced by agp_hyper_grads (closed-form gradient
# of ELBO(m, x, y, μ₀, ks, Zs, state) on the device); update_kernel! / update_Z! (autotuning_utils.jl:47-82) stay in Julia and their
# results are pushed back with agp_set_kernel / agp_set_Z.  Like the reference (quirk Q3: setHPupdated! at autotuning.jl:45 is
# commented out) the flag is NOT raised here: K_mm keeps the factor of this train! call (agp_keep_stale_K(1) at engine creation)
# while K_nm follows the new kernel / Z.
function AGP.update_hyperparameters!(m::SVGP{T,L,<:AnalyticVI}, state, x, y) where {T,L}
    any(!isnothing ∘ AGP.opt, m.f) || any(!isnothing ∘ AGP.Zopt, m.f) || return state
    e = ENGINES[m]; Q = length(m.f); M = AGP.dim(m.f[1]); D = length(first(m.f[1].Z))
    ds = zeros(Q); dv = zeros(Q); dZ = zeros(D, M, Q)               # [Q][m][D] row-major == (D, M, Q) column-major
    check(e, ccall((:agp_hyper_grads, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   e.model, Float64(AGP.ρ(AGP.inference(m))), ds, dv, dZ))
    hp = state.hyperopt_state
    hp = map(enumerate(m.f), hp) do (q, gp), st
        if !isnothing(AGP.opt(gp))   # NamedTuple gradient in the shape Zygote would return for σ² * (k ∘ ScaleTransform(s))
            Δ = (kernel=(kernel=nothing, transform=(s=[ds[q]],)), σ²=[dv[q]])
            st = merge(st, (; state_k=AGP.update_kernel!(AGP.opt(gp), AGP.kernel(gp), Δ, st.state_k)))
            kind, s, var = kernel_params(AGP.kernel(gp))
            check(e, ccall((:agp_set_kernel, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Float64, Float64), e.model, q - 1, kind, s, var))
        end
        if !isnothing(AGP.Zopt(gp))
            ΔZ = [dZ[:, i, q] for i in 1:M]
            st = merge(st, (; state_Z=AGP.update_Z!(AGP.Zopt(gp), AGP.Zview(gp), ΔZ, st.state_Z)))
            Zn = reduce(hcat, AGP.Zview(gp))                          # (D, M) column-major == [m][D] row-major
            check(e, ccall((:agp_set_Z, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), e.model, q - 1, Zn))
        end
        st
    end
    return merge(state, (; hyperopt_state=hp))
end

end # module
