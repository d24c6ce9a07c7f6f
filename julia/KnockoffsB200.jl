# KnockoffsB200.jl -- the binding a Knockoffs.jl maintainer adds to route the model-X Gaussian hot path
# to libknockoffs_b200.so.  Source only: Julia is not installed in the build image, so this file is not
# exercised by the test-suite; the identical exported C symbols are driven from Python ctypes instead
# (knockoffs.jl_b200/_lib.py).  Struct layouts mirror include/knockoffs_b200.h.
module KnockoffsB200

using LinearAlgebra
import Knockoffs: GaussianKnockoff

const libkb = get(ENV, "KNOCKOFFS_B200_LIB", "libknockoffs_b200.so")
const KB_MAX_TRACE = 128
const METHODS = Dict(:mvr => 0, :maxent => 1, :sdp_ccd => 2, :equi => 3)

struct SolveOpts          # kb_solve_opts
    niter::Int32; tol::Float64; lambda_min::Float64; sdp_lambda::Float64; sdp_mu::Float64
    robust::Int32; verbose::Int32; mode::Int32; refresh_every::Int32; objective::Int32
end
struct SolveInfo          # kb_solve_info
    sweeps::Int32; status::Int32; mode::Int32; launches::Int32
    trace::NTuple{KB_MAX_TRACE,Float64}
    solve_ms::Float64; init_ms::Float64; ref_bytes::Float64; touched_bytes::Float64
    flops::Float64; updates::Float64; objective::Float64
end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:kb_create, libkb), Cint, (Ptr{Ptr{Cvoid}}, Cint), r, device)
        rc == 0 || error("kb_create failed ($rc): no usable CUDA device")
        ctx = new(r[])
        # `condition` uses cholesky(Positive, ·) in the reference (modelX.jl:111,120): opt in to the library's restatement
        # of it (parity unpinned, see include/knockoffs_b200.h) so :equi and boundary solutions return knockoffs here too
        ccall((:kb_set_positive_cholesky, libkb), Cint, (Ptr{Cvoid}, Int32), ctx.h, 1)
        finalizer(c -> ccall((:kb_destroy, libkb), Cvoid, (Ptr{Cvoid},), c.h), ctx)
    end
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
ctx() = (CTX[] === nothing && (CTX[] = Context()); CTX[])

function check(c::Context, rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:kb_last_error, libkb), Cstring, (Ptr{Cvoid},), c.h))
    rc == -2 && throw(PosDefException(0))        # cholesky / downdate failure (utilities.jl:140, :591-593)
    error(msg)                                    # error(...) sites: modelX.jl:105, utilities.jl:28,46,195-196,301-302
end

opts(; niter::Int=100, tol=1e-6, λmin=1e-6, λ=0.5, μ=0.8, robust::Bool=false, verbose::Bool=false) =
    SolveOpts(niter, tol, λmin, λ, μ, robust, verbose, 0, -1, 1)  # mode = KB_MODE_AUTO, refresh_every = -1 (library default), objective on

"Drop-in for `solve_s(Σ::Symmetric, method; m, kwargs...)` (src/utilities.jl:27-51) for :mvr, :maxent, :sdp_ccd, :equi."
function solve_s(Σ::Symmetric{Float64}, method::Union{Symbol,String}; m::Number=1, s_init=nothing, kwargs...)
    c = ctx(); p = size(Σ, 1); s = Vector{Float64}(undef, p)
    o = Ref(opts(; kwargs...)); info = Ref{SolveInfo}()
    A = Σ.data                                   # only the upper triangle is read (Symmetric, uplo = :U)
    si = s_init === nothing ? C_NULL : pointer(s_init)
    GC.@preserve A s s_init begin
        rc = ccall((:kb_solve_s, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Float64, Ptr{SolveOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{SolveInfo}),
                   c.h, A, p, METHODS[Symbol(method)], Float64(m), o, si, s, info)
        check(c, rc)
    end
    if o[].verbose != 0                            # same lines the reference prints (utilities.jl:167-170)
        for l in 1:info[].sweeps
            println("Iter $l: δ = $(info[].trace[l])")
        end
    end
    return s
end

"Drop-in for `modelX_gaussian_knockoffs(X, method, μ, Σ; m, kwargs...)` (src/modelX.jl:53-66)."
function modelX_gaussian_knockoffs(X::Matrix{Float64}, method::Union{Symbol,String}, μ::Vector{Float64},
                                   Σ::AbstractMatrix{Float64}; m::Number=1, seed::Integer=rand(UInt64), kwargs...)
    c = ctx(); n, p = size(X); mi = Int(m)
    mi < 1 && error("m should be 1 or larger but was $m.")
    Xko = Matrix{Float64}(undef, n, mi * p); s = Vector{Float64}(undef, p)
    Σd = Matrix{Float64}(Σ isa Symmetric ? Σ.data : Σ)
    o = Ref(opts(; kwargs...)); info = Ref{SolveInfo}()
    GC.@preserve X μ Σd Xko s begin
        rc = ccall((:kb_modelx_gaussian_knockoffs, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{SolveOpts},
                    Ptr{Float64}, Ptr{Float64}, UInt64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{SolveInfo}),
                   c.h, X, n, p, μ, Σd, METHODS[Symbol(method)], mi, o, C_NULL, C_NULL, seed, 0, Xko, s, info)
        check(c, rc)
    end
    return GaussianKnockoff(X, Xko, s, Symmetric(Σd), Symbol(method), mi)   # src/struct.jl:6-13, fields unchanged
end

"Drop-in for `condition(X, μ, Σ, S; m)` (src/modelX.jl:96-122); S::Diagonal or dense.  One call: Σ⁻¹S and the factor stay on the device."
function condition(X::Matrix{Float64}, μ::Vector{Float64}, Σ::AbstractMatrix{Float64}, S::AbstractMatrix{Float64};
                   m::Number=1, seed::Integer=rand(UInt64))
    c = ctx(); n, p = size(X); mi = Int(m)
    mi < 1 && error("m should be 1 or larger but was $m.")
    Σd = Matrix{Float64}(Σ isa Symmetric ? Σ.data : Σ)
    isdiag = S isa Diagonal
    Sd = isdiag ? Vector{Float64}(S.diag) : Matrix{Float64}(S)
    Xko = Matrix{Float64}(undef, n, mi * p)
    GC.@preserve X μ Σd Sd Xko begin
        rc = ccall((:kb_condition, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64},
                    UInt64, Int64, Ptr{Float64}),
                   c.h, X, n, p, μ, Σd, Sd, isdiag, mi, C_NULL, seed, 0, Xko)
        check(c, rc)
    end
    return Xko
end

# ---- several GPUs behind one handle (kb_create_multi; SURVEY 8b / 8e) ------------------------------------
# One Julia process drives all GPUs: the library runs one host thread per device.  Rows of X are sharded for `condition`
# and for the Gram (one peer-to-peer all-reduce over NVLink); the CCD solve itself stays on device 0 (sequential sweep).
mutable struct MultiContext
    h::Ptr{Cvoid}
    function MultiContext(devices::Vector{<:Integer}=collect(0:7))
        r = Ref{Ptr{Cvoid}}(C_NULL); d = Cint.(devices)
        rc = ccall((:kb_create_multi, libkb), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Cint}, Int32), r, d, length(d))
        rc == 0 || error("kb_create_multi failed ($rc)")
        mc = new(r[])
        finalizer(x -> ccall((:kb_destroy_multi, libkb), Cvoid, (Ptr{Cvoid},), x.h), mc)
    end
end
function check(mc::MultiContext, rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:kb_multi_last_error, libkb), Cstring, (Ptr{Cvoid},), mc.h))
    rc == -2 && throw(PosDefException(0))
    error(msg)
end
"`modelX_gaussian_knockoffs` on all GPUs of `mc` (same result as the single-GPU call: the RNG is keyed by the global row)."
function modelX_gaussian_knockoffs(mc::MultiContext, X::Matrix{Float64}, method::Union{Symbol,String}, μ::Vector{Float64},
                                   Σ::AbstractMatrix{Float64}; m::Number=1, seed::Integer=rand(UInt64), kwargs...)
    n, p = size(X); mi = Int(m)
    Xko = Matrix{Float64}(undef, n, mi * p); s = Vector{Float64}(undef, p)
    Σd = Matrix{Float64}(Σ isa Symmetric ? Σ.data : Σ)
    o = Ref(opts(; kwargs...)); info = Ref{SolveInfo}()
    GC.@preserve X μ Σd Xko s begin
        rc = ccall((:kb_multi_modelx_gaussian_knockoffs, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{SolveOpts},
                    Ptr{Float64}, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{SolveInfo}),
                   mc.h, X, n, p, μ, Σd, METHODS[Symbol(method)], mi, o, C_NULL, C_NULL, seed, Xko, s, info)
        check(mc, rc)
    end
    return GaussianKnockoff(X, Xko, s, Symmetric(Σd), Symbol(method), mi)
end
"Σ̂ = Xc'Xc/n and the column means with the rows of X sharded over the GPUs of `mc`."
function gram(mc::MultiContext, X::Matrix{Float64})
    n, p = size(X); G = Matrix{Float64}(undef, p, p); μ = Vector{Float64}(undef, p)
    GC.@preserve X G μ begin
        check(mc, ccall((:kb_multi_gram, libkb), Cint,
                        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int32, Float64, Ptr{Float64}, Ptr{Float64}), mc.h, X, n, p, 1, 0.0, G, μ))
    end
    return Symmetric(G), μ
end

# ---- group knockoffs (src/group.jl:109-127, 348-422) ---------------------------------------------------
const GROUP_METHODS = Dict(:mvr => 0, :maxent => 1, :sdp => 2, :equi => 3)
struct GroupOpts          # kb_group_opts
    outer_iter::Int32; inner_pca_iter::Int32; inner_ccd_iter::Int32; tol::Float64; eps::Float64
    robust::Int32; verbose::Int32
end
struct GroupInfo          # kb_group_info
    outer_iters::Int32; inner_sweeps::Int32; n_directions::Int32; status::Int32
    obj_init::Float64; obj::Float64; gamma_equi::Float64
    trace_obj::NTuple{KB_MAX_TRACE,Float64}; trace_delta::NTuple{KB_MAX_TRACE,Float64}
    trace_kind::NTuple{KB_MAX_TRACE,Int32}
    steps::Float64; accepted::Float64; solve_ms::Float64; touched_bytes::Float64
end
gopts(; outer_iter::Int=100, inner_pca_iter::Int=1, inner_ccd_iter::Int=1, tol=1e-4, ϵ=1e-6, robust::Bool=false,
      verbose::Bool=false) = GroupOpts(outer_iter, inner_pca_iter, inner_ccd_iter, tol, ϵ, robust, verbose)

"Drop-in for `solve_s_group(Σ::Symmetric, groups, method; m, kwargs...)` (src/group.jl:348-422): returns (S, γs, obj)."
function solve_s_group(Σ::Symmetric{Float64}, groups::Vector{Int}, method::Union{Symbol,String}; m::Number=1, kwargs...)
    c = ctx(); p = size(Σ, 1)
    length(groups) == p || error("Length of groups should be equal to dimension of Σ")
    S = Matrix{Float64}(undef, p, p); obj = Ref{Float64}(0.0)
    o = Ref(gopts(; kwargs...)); info = Ref{GroupInfo}(); A = Σ.data
    GC.@preserve A groups S begin
        rc = ccall((:kb_solve_s_group, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Int32, Float64, Ptr{GroupOpts}, Ptr{Float64}, Int64,
                    Ptr{Float64}, Ptr{Float64}, Ptr{GroupInfo}),
                   c.h, A, p, groups, GROUP_METHODS[Symbol(method)], Float64(m), o, C_NULL, 0, S, obj, info)
        check(c, rc)
    end
    return S, Float64[], obj[]
end

# ---- approximate knockoffs (src/approx.jl:62-102) ------------------------------------------------------
"Drop-in for `approx_modelX_gaussian_knockoffs(X, method, window_ranges; m, kwargs...)`; returns (Xko, s, Σ, γ)."
function approx_modelX_gaussian_knockoffs(X::Matrix{Float64}, method::Union{Symbol,String},
                                          window_ranges::Vector{UnitRange{Int64}}; m::Int=1,
                                          seed::Integer=rand(UInt64), kwargs...)
    c = ctx(); n, p = size(X)
    sum(length.(window_ranges)) == p || error("window_ranges have $(sum(length.(window_ranges))) dimensions but X has $p")
    off = Int64[0; cumsum(length.(window_ranges))]
    Xko = Matrix{Float64}(undef, n, m * p); s = Vector{Float64}(undef, p); μ = Vector{Float64}(undef, p)
    blocks = Vector{Float64}(undef, sum(abs2, length.(window_ranges))); γ = Ref{Float64}(1.0)
    o = Ref(opts(; kwargs...)); info = Ref{SolveInfo}()
    GC.@preserve X off Xko s μ blocks begin
        rc = ccall((:kb_approx_modelx_gaussian_knockoffs, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int32, Int32, Int32, Ptr{SolveOpts}, Ptr{Float64},
                    UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{SolveInfo}),
                   c.h, X, n, p, off, length(window_ranges), METHODS[Symbol(method)], m, o, C_NULL, seed, Xko, s, μ,
                   blocks, γ, info)
        check(c, rc)
    end
    Σ = zeros(p, p); b = 0                                   # BlockDiagonal(block_covariances) |> Matrix  (approx.jl:93)
    for r in window_ranges
        w = length(r); Σ[r, r] .= reshape(view(blocks, b+1:b+w*w), w, w); b += w * w
    end
    return Xko, s, Symmetric(Σ), γ[]                          # wrap in ApproxGaussianKnockoff (src/struct.jl:15-22)
end

# ---- feature statistics + filter (src/fit_lasso.jl:188-257) ----------------------------------------------
"Drop-in for `fit_marginal(y, X, μ, Σ; method, m, fdrs)`: returns (W, τs, selected); X̃ stays on the device."
function fit_marginal(y::Vector{Float64}, X::Matrix{Float64}, μ::Vector{Float64}, Σ::Matrix{Float64};
                      method::Union{Symbol,String}=:maxent, m::Int=1, fdrs::Vector{Float64}=[0.01, 0.05, 0.1, 0.25, 0.5],
                      seed::Integer=rand(UInt64), kwargs...)
    c = ctx(); n, p = size(X); nf = length(fdrs)
    s = Vector{Float64}(undef, p); T = Vector{Float64}(undef, (m + 1) * p); W = Vector{Float64}(undef, p)
    κ = zeros(Int64, p); τ = zeros(p); τs = Vector{Float64}(undef, nf); mask = zeros(Int32, p, nf)
    o = Ref(opts(; kwargs...)); info = Ref{SolveInfo}()
    GC.@preserve y X μ Σ fdrs s T W κ τ τs mask begin
        rc = ccall((:kb_fit_marginal, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Int32,
                    Ptr{SolveOpts}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{SolveInfo}),
                   c.h, y, X, n, p, μ, Σ, METHODS[Symbol(method)], m, o, fdrs, nf, 1, C_NULL, seed, C_NULL, s, T, W, κ, τ,
                   τs, mask, info)
        check(c, rc)
    end
    return W, τs, [findall(!iszero, view(mask, :, f)) for f in 1:nf]
end

# ---- representative group knockoffs (src/group.jl:155-303, 1929-2033) -----------------------------------------
"Drop-in for `choose_group_reps(Σ::Symmetric, groups; threshold, prioritize_idx)`; indices are 1-based on both sides."
function choose_group_reps(Σ::Symmetric{Float64}, groups::Vector{Int}; threshold=0.5,
                           prioritize_idx::Union{Vector{Int},Nothing}=nothing)
    c = ctx(); p = size(Σ, 1); reps = Vector{Int64}(undef, p); nr = Ref{Int64}(0); A = Σ.data
    pr = prioritize_idx === nothing ? C_NULL : pointer(prioritize_idx)
    GC.@preserve A groups reps prioritize_idx begin
        rc = ccall((:kb_choose_group_reps, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Float64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}),
                   c.h, A, p, groups, Float64(threshold), pr, prioritize_idx === nothing ? 0 : length(prioritize_idx), reps, nr)
        check(c, rc)
    end
    return reps[1:nr[]]
end

"Drop-in for `modelX_gaussian_rep_group_knockoffs(X, method, groups, μ, Σ; m, rep_threshold, enforce_cond_indep, kwargs...)`."
function modelX_gaussian_rep_group_knockoffs(X::Matrix{Float64}, method::Union{Symbol,String}, groups::Vector{Int},
                                             μ::Vector{Float64}, Σ::Matrix{Float64}; m::Int=1, rep_threshold=0.5,
                                             enforce_cond_indep::Bool=false, seed::Integer=rand(UInt64), kwargs...)
    c = ctx(); n, p = size(X)
    p == length(groups) || error("Dimensions of X and groups doesn't match")
    Xko = Matrix{Float64}(undef, n, m * p); reps = Vector{Int64}(undef, p); nr = Ref{Int64}(0)
    S11 = Vector{Float64}(undef, p * p); D = Matrix{Float64}(undef, p, p); obj = Ref{Float64}(0.0)
    o = Ref(gopts(; kwargs...)); info = Ref{GroupInfo}()
    GC.@preserve X groups μ Σ Xko reps S11 D begin
        rc = ccall((:kb_modelx_gaussian_rep_group_knockoffs, libkb), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64, Int32,
                    Ptr{GroupOpts}, Ptr{Int64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Int64},
                    Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{GroupInfo}),
                   c.h, X, n, p, groups, μ, Σ, GROUP_METHODS[Symbol(method)], m, Float64(rep_threshold), enforce_cond_indep, o,
                   C_NULL, 0, C_NULL, 0, C_NULL, seed, Xko, reps, nr, S11, D, obj, info)
        check(c, rc)
    end
    r = nr[]
    # wrap in GaussianRepGroupKnockoff (src/struct.jl:36-48)
    return Xko, reps[1:r], reshape(S11[1:r*r], r, r), Symmetric(D), obj[]
end

end # module
