# NMFkB200.jl - drop-in for the factorization hot path of NMFk.jl on B200 GPUs.
#
# Same signatures and return shapes as the reference for this path
#   NMFk.NMFmultiplicative   (NMFk.jl/src/NMFkMultiplicative.jl:24)
#   NMFk.execute_singlerun   (NMFk.jl/src/NMFkExecute.jl:714,729)   method=:simple only
#   NMFk.execute_run         (NMFk.jl/src/NMFkExecute.jl:483)
#   NMFk.execute             (NMFk.jl/src/NMFkExecute.jl:178, 236)
# and nothing else.  All numerics happen in libnmfk_b200.so (hand-written CUDA for sm_100a) through
# the C ABI declared in include/nmfk_b200.h; this file only marshals arguments with `ccall`.
# Random initial factors are drawn HERE with Julia's RNG in the reference's order (W = rand(n,k)
# then H = rand(k,m), restart i seeded seed+i), so a run with a given `seed` starts from exactly the
# factors the reference would start from.  The JLD result cache / X hash of `execute` are file IO
# outside the path; call the reference's `NMFk.load/save` around these functions if needed.
#
# NOTE: no Julia runtime exists in the build/test environment of this repository, so this file has
# never been executed there; every call below is mirrored 1:1 by the Python ctypes host
# (nmfk.jl_b200/python/nmfk_b200/api.py), which is what the parity tests drive.
module NMFkB200

import Random
import Libdl

const libnmfk = get(ENV, "NMFK_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "lib", "libnmfk_b200.so"))

const NMFK_F32 = Cint(0)
const NMFK_F64 = Cint(1)

# mirrors `struct nmfk_params` (include/nmfk_b200.h)
struct Params
	tol::Cdouble
	tolOF::Cdouble
	eps_clamp::Cdouble
	weight::Cdouble
	maxiter::Cint
	maxbaditers::Cint
	maxreattempts::Cint
	stopconv::Cint
	check_every::Cint
	Wfixed::Cint
	Hfixed::Cint
	normalize::Cint
	iter_limit::Cint
	engine::Cint
	reserved::NTuple{4,Cint}
end

function Params(; tol=1e-19, tolOF=1e-3, weight=1, maxiter=10000, maxbaditers=10, maxreattempts=2, stopconv=1000, Wfixed=false, Hfixed=false, normalize=1, engine=0, kw...)
	typeof(weight) <: Number || error("vector/matrix weights are not on the B200 path yet")
	return Params(tol, tolOF, eps(Float64), weight, maxiter, maxbaditers, maxreattempts, stopconv, 10, Wfixed, Hfixed, normalize, 0, engine, (Cint(0), Cint(0), Cint(0), Cint(0)))
end

dtypecode(::Type{Float32}) = NMFK_F32
dtypecode(::Type{Float64}) = NMFK_F64

function check(status::Integer, ctx::Ptr{Cvoid}=C_NULL)
	status == 0 && return nothing
	msg = unsafe_string(ccall((:nmfk_last_error, libnmfk), Cstring, (Ptr{Cvoid},), ctx))
	if status == -2
		throw(ErrorException("All matrix entries must be nonnegative!")) # NMFkMultiplicative.jl:4-7
	elseif status == -7
		error("Input array has a zero dimension!") # NMFkExecute.jl:242-244
	end
	error("nmfk_b200 status $(status): $(msg)")
end

mutable struct Context
	h::Ptr{Cvoid}
	function Context(device::Integer=0)
		r = Ref{Ptr{Cvoid}}(C_NULL)
		check(ccall((:nmfk_ctx_create, libnmfk), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
		c = new(r[])
		finalizer(x->(x.h != C_NULL && ccall((:nmfk_ctx_destroy, libnmfk), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), c)
		return c
	end
end

"NMFpreprocessing! (NMFkMultiplicative.jl:3-22); the caller's X is not modified"
function setX!(c::Context, X::AbstractMatrix{T}; lambda::Number=1e-32) where {T <: Union{Float32,Float64}}
	Xd = Matrix{T}(X) # dense, column-major
	check(ccall((:nmfk_set_X, libnmfk), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cint, Cdouble, Ptr{Cvoid}, Cint), c.h, Xd, size(Xd, 1), size(Xd, 2), dtypecode(T), lambda, C_NULL, 0), c.h)
	return nothing
end

"Initial factors for nNMF restarts drawn like the reference: W = rand(n,k) then H = rand(k,m) (NMFkMultiplicative.jl:38,48), restart i seeded seed+i (NMFkExecute.jl:536)"
function drawinits(::Type{T}, n::Integer, m::Integer, k::Integer, nNMF::Integer; seed::Integer=-1, Winit::AbstractMatrix=Matrix{T}(undef, 0, 0), Hinit::AbstractMatrix=Matrix{T}(undef, 0, 0)) where {T}
	W = Array{T,3}(undef, n, k, nNMF)
	H = Array{T,3}(undef, k, m, nNMF)
	for i = 1:nNMF
		seed >= 0 && Random.seed!(seed + i)
		if sizeof(Winit) == 0
			W[:, :, i] = rand(n, k)
		else
			@assert size(Winit) == (n, k)
			sum(isnan.(Winit)) > 0 && error("Initial values for the W matrix entries include NaNs!")
			W[:, :, i] = Winit
		end
		if sizeof(Hinit) == 0
			H[:, :, i] = rand(k, m)
		else
			@assert size(Hinit) == (k, m)
			sum(isnan.(Hinit)) > 0 && error("Initial values for the H matrix entries include NaNs!")
			H[:, :, i] = Hinit
		end
	end
	return W, H
end

"NMFk.NMFmultiplicative(X, k; ...) -> (W, H, objvalue)  (NMFkMultiplicative.jl:24-127)"
function NMFmultiplicative(X::AbstractMatrix{T}, k::Int; seed::Int=-1, lambda::Number=1e-32, maxiter::Int=1000000, Winit::AbstractMatrix{T}=Matrix{T}(undef, 0, 0), Hinit::AbstractMatrix{T}=Matrix{T}(undef, 0, 0), normalizevector::AbstractVector{T}=Vector{T}(undef, 0), ctx::Context=Context(), kw...) where {T <: Union{Float32,Float64}}
	length(normalizevector) == 0 || error("normalizevector is not on the B200 path yet")
	n, m = size(X)
	setX!(ctx, X; lambda=lambda)
	if seed >= 0
		Random.seed!(seed)
	end
	Wi, Hi = drawinits(T, n, m, k, 1; Winit=Winit, Hinit=Hinit)
	p = Params(; maxiter=maxiter, normalize=0, kw...)
	W = Matrix{T}(undef, n, k); H = Matrix{T}(undef, k, m)
	ssq = Ref{Cdouble}(0); nrm = Ref{Cdouble}(0); it = Ref{Cint}(0); sr = Ref{Cint}(0)
	check(ccall((:nmfk_run_batch, libnmfk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cint}, Ref{Cint}), ctx.h, k, 1, Wi, Hi, p, W, H, ssq, nrm, it, sr), ctx.h)
	return W, H, ssq[]
end

"NMFk.execute_singlerun(X, nk; method=:simple, ...) -> (W, H, objvalue)  (NMFkExecute.jl:729-807)"
function execute_singlerun(X::AbstractMatrix{T}, nk::Int; seed::Int=-1, clusterWmatrix::Bool=false, modifymatrices::Bool=true, method::Symbol=:simple, ctx::Context=Context(), kw...) where {T <: Union{Float32,Float64}}
	method == :simple || error("NMFkB200 covers method=:simple only; use NMFk for $(method)")
	n, m = size(X)
	setX!(ctx, X)
	seed >= 0 && Random.seed!(seed)
	Wi, Hi = drawinits(T, n, m, nk, 1; filter(p->p.first in (:Winit, :Hinit), kw)...)
	p = Params(; normalize=(modifymatrices ? (clusterWmatrix ? 2 : 1) : 0), kw...)
	W = Matrix{T}(undef, n, nk); H = Matrix{T}(undef, nk, m)
	ssq = Ref{Cdouble}(0); nrm = Ref{Cdouble}(0); it = Ref{Cint}(0); sr = Ref{Cint}(0)
	check(ccall((:nmfk_run_batch, libnmfk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cint}, Ref{Cint}), ctx.h, nk, 1, Wi, Hi, p, W, H, ssq, nrm, it, sr), ctx.h)
	return W, H, convert(T, nrm[])
end

"NMFk.execute_run(X, nk, nNMF; ...) -> (Wa, Ha, phi_final, minsilhouette, aic)  (NMFkExecute.jl:483-711)"
function execute_run(X::AbstractMatrix{T}, nk::Int, nNMF::Int; clusterWmatrix::Bool=false, acceptratio::Number=1, acceptfactor::Number=Inf, best::Bool=true, nanaction::Symbol=:zeroed, seed::Int=-1, ctx::Context=Context(), kw...) where {T <: Union{Float32,Float64}}
	(acceptratio == 1 && acceptfactor == Inf && best && nanaction == :zeroed) || error("NMFkB200 covers the default acceptratio/acceptfactor/best/nanaction only")
	n, m = size(X)
	setX!(ctx, X)
	modifymatrices = !(haskey(kw, :Wfixed) || haskey(kw, :Hfixed)) # NMFkExecute.jl:486-489
	Wi, Hi = drawinits(T, n, m, nk, nNMF; seed=seed, filter(p->p.first in (:Winit, :Hinit), kw)...)
	p = Params(; normalize=(modifymatrices ? (clusterWmatrix ? 2 : 1) : 0), kw...)
	Wa = Matrix{T}(undef, n, nk); Ha = Matrix{T}(undef, nk, m)
	phi = Ref{Cdouble}(0); rob = Ref{Cdouble}(0); aic = Ref{Cdouble}(0); tot = Ref{Int64}(0)
	check(ccall((:nmfk_execute_run, libnmfk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ref{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}, Ref{Int64}), ctx.h, nk, nNMF, Wi, Hi, 0, p, Wa, Ha, phi, rob, aic, tot), ctx.h)
	return Wa, Ha, convert(T, phi[]), (nk > 1 ? convert(T, rob[]) : 1), aic[]
end

"NMFk.execute(X, nk::Integer, nNMF; ...) -> (W[:,so], H[so,:], fitquality, robustness, aic)  (NMFkExecute.jl:236-329, without the JLD cache)"
function execute(X::AbstractMatrix{T}, nk::Integer, nNMF::Integer=10; kw...) where {T <: Union{Float32,Float64}}
	W, H, fit, rob, aic, kopt = execute(X, nk:nk, nNMF; kw...)
	return W[nk], H[nk], fit[nk], rob[nk], aic[nk]
end

"NMFk.execute(X, nkrange, nNMF; cutoff=0.5, ...) -> (W, H, fitquality, robustness, aic, kopt)  (NMFkExecute.jl:178-233, without the JLD cache); all k are solved concurrently on the GPU"
function execute(X::AbstractMatrix{T}, nkrange::Union{Vector{Int},AbstractUnitRange{Int}}, nNMF::Integer=10; cutoff::Number=0.5, clusterWmatrix::Bool=false, method::Symbol=:simple, seed::Int=-1, ctx::Context=Context(), kw...) where {T <: Union{Float32,Float64}}
	method == :simple || error("NMFkB200 covers method=:simple only; use NMFk for $(method)")
	.*(size(X)...) == 0 && error("Input array has a zero dimension! Array size=$(size(X))")
	n, m = size(X)
	setX!(ctx, X)
	ks = collect(Cint, nkrange)
	nks = length(ks)
	modifymatrices = !(haskey(kw, :Wfixed) || haskey(kw, :Hfixed))
	p = Params(; normalize=(modifymatrices ? (clusterWmatrix ? 2 : 1) : 0), kw...)
	inits = [drawinits(T, n, m, Int(k), nNMF; seed=seed, filter(p->p.first in (:Winit, :Hinit), kw)...) for k in ks]
	Wo = [Matrix{T}(undef, n, Int(k)) for k in ks]
	Ho = [Matrix{T}(undef, Int(k), m) for k in ks]
	fit = Vector{Cdouble}(undef, nks); rob = Vector{Cdouble}(undef, nks); aicv = Vector{Cdouble}(undef, nks)
	kopt = Ref{Cint}(0); tot = Ref{Int64}(0)
	GC.@preserve inits Wo Ho begin
		Wip = [pointer(i[1]) for i in inits]; Hip = [pointer(i[2]) for i in inits]
		Wop = [pointer(w) for w in Wo]; Hop = [pointer(h) for h in Ho]
		check(ccall((:nmfk_execute, libnmfk), Cint, (Ptr{Cvoid}, Ptr{Cint}, Cint, Cint, Ptr{Ptr{T}}, Ptr{Ptr{T}}, UInt64, Ref{Params}, Cdouble, Ptr{Ptr{T}}, Ptr{Ptr{T}}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cint}, Ref{Int64}), ctx.h, ks, nks, nNMF, Wip, Hip, 0, p, cutoff, Wop, Hop, fit, rob, aicv, kopt, tot), ctx.h)
	end
	maxk = maximum(ks)
	W = Vector{Matrix{T}}(undef, maxk); H = Vector{Matrix{T}}(undef, maxk)
	fitquality = zeros(T, maxk); robustness = zeros(T, maxk); aic = zeros(T, maxk)
	fitquality[1] = Inf; robustness[1] = -1 # NMFkExecute.jl:200-201
	for (i, k) in enumerate(ks)
		W[k] = Wo[i]; H[k] = Ho[i]
		fitquality[k] = fit[i]; robustness[k] = (k > 1 ? rob[i] : 1); aic[k] = aicv[i]
	end
	ko = kopt[] < 0 ? nothing : Int(kopt[]) # getk returns `nothing` when no k passes the cutoff
	return W, H, fitquality, robustness, aic, ko
end

end
