# ReachSM100a.jl -- the thin `ccall` layer a Reachability.jl maintainer adds to route the linear-flowpipe hot path
# through libreach_sm100a.so (include/reach.h).  Nothing here computes: it marshals Julia arrays (column-major
# Float64, 1-based indices -> 0-based Int32), checks the int32 status and rebuilds the result containers.
#
# Seams replaced (file:line under Reachability.jl v0.7.0 `src/`), each by a MORE SPECIFIC METHOD of the reference's own
# function with the reference's exact positional signature, so that the stock call sites dispatch to it unchanged:
#   reach_blocks! (dense)             ReachSets/ContinuousPost/BFFPSV18/reach_blocks.jl:135-150, called through the registry
#                                     `available_algorithms[backend]["func"](args...)` at BFFPSV18/reach.jl:191-193 (15 positional args)
#   check_blocks  (dense)             ReachSets/ContinuousPost/BFFPSV18/check_blocks.jl:105-116 (11 positional args)
#   reach_homog! / reach_inhomog!     ReachSets/ContinuousPost/GLGM06/reach.jl:5-9, 30-36, called at GLGM06/post.jl:32,36
# and by plain functions the maintainer calls from the matching branch:
#   exp_Aδ / Φ₁ / Φ₂                  ReachSets/discretize.jl:150-160, 221-240, 296-315   (new `exp_method = "sm100a"`)
#   discretize (box / zonotope)       ReachSets/discretize.jl:627-743                      (reach_discretize_box / _zonotope)
#   reduce_order (call site)          ReachSets/ContinuousPost/GLGM06/post.jl:20
#   project_reach (zonotopes)         ReachSets/project_reach.jl:18-99
#   SparseMatrixExp / lazy rows       ReachSets/discretize.jl:153-154, 302-306; BFFPSV18/reach_blocks.jl:227-397
#
# Dispatch notes.  `Xhat0` is a `Vector{LazySet{Float64}}` in the reference (`array(decompose(...))`, reach.jl:62-65), so the
# element types cannot be dispatched on: the methods below specialise on `ϕ::Matrix{Float64}` (strictly more specific than
# the stock `ϕ::AbstractMatrix{NUM}`, everything else identical) and inspect the rest at run time.  `overapproximate`,
# `output_function` and `termination` are closures built in reach.jl:73-75,102-111,167-168; what they were built from is
# read off their captured variables (closure fields).  Inputs outside the device path raise -- libreach_sm100a has NO CPU
# fallback -- unless ENV["REACH_SM100A_ALLOW_REFERENCE"] = "1", in which case the stock method is `invoke`d.
#
# No Julia toolchain exists in the build container or on the GPU box, so this file is NOT exercised by the test suite;
# the Python ctypes binding (reachability.jl_b200/_lib.py) and the C99 client (tests/c_abi/flowpipe_smoke.c) drive the
# identical entry points and are what the parity tests run.
module ReachSM100a

using LazySets, SparseArrays, LinearAlgebra
import Reachability
import Reachability.ReachSets: reach_blocks!, check_blocks, reach_homog!, reach_inhomog!,
                               ReachSet, SparseReachSet, AbstractReachSet
using MathematicalSystems: ConstantInput
using MathematicalPredicates: Predicate
using ProgressMeter: Progress
import Reachability.ReachSets: next_set

const LIB = get(ENV, "LIBREACH_SM100A", "libreach_sm100a.so")
const ALLOW_REFERENCE = get(ENV, "REACH_SM100A_ALLOW_REFERENCE", "0") == "1"

# ---- context: one device, or all devices of the box behind ONE handle (reach_ctx_create_multi) -------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(devices::Vector{Int32})
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:reach_ctx_create_multi, LIB), Int32, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devices, length(devices), h)
        rc == 0 || error("libreach_sm100a: ", unsafe_string(ccall((:reach_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[])
        finalizer(c -> ccall((:reach_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

const CTX = Ref{Union{Context,Nothing}}(nothing)
# ENV["REACH_DEVICES"] = "0,1,2,3,4,5,6,7": the BFFPSV18 calls shard their rows over these GPUs inside the library, so the
# single Julia task of `solve` (src/solve.jl:59-81) drives the whole box
function context()
    if CTX[] === nothing
        CTX[] = Context(Int32[parse(Int32, s) for s in split(get(ENV, "REACH_DEVICES", "0"), ",")])
    end
    return CTX[]
end

check(ctx, rc) = rc == 0 || error("libreach_sm100a status $rc: ",
    unsafe_string(ccall((:reach_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))

dense(A::AbstractMatrix) = Matrix{Float64}(A)          # the reference densifies too: exp(Matrix(A*δ)), discretize.jl:152

unsupported(what) = error("libreach_sm100a: $what is outside the sm_100a path (no CPU fallback; " *
                          "set ENV[\"REACH_SM100A_ALLOW_REFERENCE\"] = \"1\" to run the stock method instead)")

# ---- discretize.jl:150-160 -- add:  elseif exp_method == "sm100a"  return ReachSM100a.exp_Aδ(A, δ)
function exp_Aδ(A::AbstractMatrix{Float64}, δ::Float64)
    ctx = context(); Ad = dense(A); n = size(Ad, 1); Φ = Matrix{Float64}(undef, n, n)
    check(ctx, ccall((:reach_expm, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64),
                     ctx.handle, n, Ad, n, δ, Φ, n))
    return Φ
end

# ---- discretize.jl:221-240 / 296-315 -- the "sm100a" branch of Φ₁ and Φ₂ (no 3n x 3n exponential)
function Φ12(A::AbstractMatrix{Float64}, δ::Float64)
    ctx = context(); Ad = dense(A); n = size(Ad, 1)
    P1 = Matrix{Float64}(undef, n, n); P2 = Matrix{Float64}(undef, n, n)
    check(ctx, ccall((:reach_phi12, LIB), Int32,
                     (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
                     ctx.handle, n, Ad, n, δ, P1, n, P2, n))
    return P1, P2
end
Φ₁(A, δ) = Φ12(A, δ)[1]
Φ₂(A, δ) = Φ12(A, δ)[2]

# ---- discretize.jl:705-743 -- `discretize(...; set_operations="zonotope")` for a box X0 and a box input through B:
#      returns (Φ, Ω0::Zonotope, V::Union{Zonotope,Nothing}); GLGM06/post.jl:15-18 wraps them into the discrete IVP
function discretize_zonotope(A::AbstractMatrix{Float64}, δ::Float64, X0::AbstractHyperrectangle,
                             B::Union{AbstractMatrix{Float64},Nothing}, U::Union{AbstractHyperrectangle,Nothing};
                             algorithm::String="forward")
    alg = algorithm == "forward" ? 0 : algorithm == "backward" ? 1 :
          throw(ArgumentError("the algorithm $algorithm with set operations=zonotope is not implemented"))   # discretize.jl:105-115
    ctx = context(); Ad = dense(A); n = size(Ad, 1)
    c, r = Vector{Float64}(center(X0)), Vector{Float64}(radius_hyperrectangle(X0))
    m = U === nothing ? 0 : dim(U)
    Bd = (U === nothing || B === nothing) ? C_NULL : dense(B)
    cU = U === nothing ? C_NULL : Vector{Float64}(center(U)); rU = U === nothing ? C_NULL : Vector{Float64}(radius_hyperrectangle(U))
    Φ = Matrix{Float64}(undef, n, n); c0 = Vector{Float64}(undef, n); G0 = Matrix{Float64}(undef, n, 4n + m + 1)
    cV = Vector{Float64}(undef, n); GV = Matrix{Float64}(undef, n, n + m); p0 = Ref{Int64}(0); pV = Ref{Int64}(0)
    check(ctx, ccall((:reach_discretize_zonotope, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
         Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Int64}, Ptr{Float64},
         Ptr{Float64}, Int64, Ref{Int64}),
        ctx.handle, n, Ad, n, δ, alg, c, r, m, Bd, n, cU, rU, 0, Φ, n, c0, G0, n, p0, cV, GV, n, pV))
    return Φ, Zonotope(c0, G0[:, 1:p0[]]), (m == 0 ? nothing : Zonotope(cV, GV[:, 1:pV[]]))
end

# ---- what the closures of BFFPSV18/reach.jl were built from --------------------------------------------------------------
captured(f, name::Symbol) = name in fieldnames(typeof(f)) ? getfield(f, name) : nothing
is_box_option(o) = o === Hyperrectangle || o === Interval
is_box_option(o::AbstractVector) = all(is_box_option, o)
is_box_option(o::Dict) = all(is_box_option, values(o))
# `overapproximate_fun = (i, X) -> overapproximate(X, block_options_iter[i])` (reach.jl:102-111)
box_blocks(overapproximate::Function) = (o = captured(overapproximate, :block_options_iter); o !== nothing && is_box_option(o))
# `(k, set, t0) -> termination_N(N, k, t0)` (reach.jl:241-243) captures N only; the invariant variants (:245-251) capture `invariant`
plain_termination(termination::Function) = captured(termination, :invariant) === nothing && captured(termination, :N) !== nothing
# `output_function = x -> options[:output_function] * x` (reach.jl:73-75)
function output_matrix(f::Function)
    o = captured(f, :options)
    o === nothing && unsupported("an output_function that is not the linear map of reach.jl:73-75")
    return Matrix{Float64}(o[:output_function])
end

# box (centre, radius) of the decomposed initial set, all n coordinates (Xhat0[j] for every j, reach_blocks.jl:185-187)
function boxof(Xs::Vector{<:LazySet})
    cs = Float64[]; rs = Float64[]
    for X in Xs
        X isa AbstractHyperrectangle || return nothing
        append!(cs, center(X)); append!(rs, radius_hyperrectangle(X))
    end
    return cs, rs
end

# The discretised input set `next_set(U)` = δ*U0 ⊕ Eψ0 (discretize.jl:688) as  V = MU * Box(cU, rU) ⊕ Box(0, rψ):  walk the
# lazy tree, accumulating the left factor of nested LinearMaps; box-shaped leaves under the identity with centre 0 (the
# symmetric interval hull Eψ0) add to rψ, every other box-shaped leaf contributes its (factor, centre, radius).
tomatrix(M::UniformScaling, k::Int) = Matrix{Float64}(M, k, k)
tomatrix(M, k::Int) = Matrix{Float64}(M)
function input_terms!(terms, rψ::Vector{Float64}, M, X::LazySet, n::Int)
    if X isa MinkowskiSum
        input_terms!(terms, rψ, M, X.X, n); input_terms!(terms, rψ, M, X.Y, n)
    elseif X isa MinkowskiSumArray
        foreach(Y -> input_terms!(terms, rψ, M, Y, n), array(X))
    elseif X isa LinearMap
        F = tomatrix(X.M, dim(X.X))
        input_terms!(terms, rψ, M === nothing ? F : M * F, X.X, n)
    elseif X isa AbstractHyperrectangle || X isa SymmetricIntervalHull
        H = box_approximation(X)
        if M === nothing && iszero(center(H))
            rψ .+= radius_hyperrectangle(H)
        else
            push!(terms, (M === nothing ? Matrix{Float64}(I, n, n) : M, Vector{Float64}(center(H)), Vector{Float64}(radius_hyperrectangle(H))))
        end
    else
        unsupported("an input set containing a $(typeof(X))")
    end
end
function inputs_of(U::Union{ConstantInput,Nothing}, n::Int)
    U === nothing && return (0, C_NULL, C_NULL, C_NULL, C_NULL)
    terms = Tuple{Matrix{Float64},Vector{Float64},Vector{Float64}}[]; rψ = zeros(n)
    input_terms!(terms, rψ, nothing, next_set(U), n)
    isempty(terms) && push!(terms, (zeros(n, 1), [0.0], [0.0]))
    MU = reduce(hcat, [t[1] for t in terms]); cU = reduce(vcat, [t[2] for t in terms]); rU = reduce(vcat, [t[3] for t in terms])
    return (size(MU, 2), MU, cU, rU, rψ)
end

# regroup the rows of (centers, radii) into the blocks' set types and fill the caller's `res` (store!, reach_blocks.jl:41-44)
function fill_box_result!(res, C, R, blocks, sizes, opts, dimensions, δ, N)
    t0, t1 = 0.0, δ
    for k in 1:N
        arr = Vector{LazySet{Float64}}(undef, length(blocks)); o = 0
        for (i, b) in enumerate(blocks)
            len = Int(sizes[i])
            T = opts isa Union{AbstractVector, Dict} ? opts[b] : opts
            arr[i] = T === Interval ? Interval(C[o+1, k] - R[o+1, k], C[o+1, k] + R[o+1, k]) :
                                      Hyperrectangle(C[o+1:o+len, k], R[o+1:o+len, k])
            o += len
        end
        res[k] = SparseReachSet(CartesianProductArray(arr), t0, t1, dimensions)
        t0 = t1; t1 += δ
    end
    return res
end

# ---- reach_blocks.jl:47-132 -- the sparse twin (a `SparseMatrixCSC` ϕ, e.g. exp_method = "pade"): same 15 positional
#      arguments.  The CSC fields cross the ABI as they are; a ϕ of decoupled subsystems skips its zero blocks on the device
#      (reach_bffpsv18_boxes_csc), everything else computes the same sets through the dense method.
function reach_blocks!(ϕ::SparseMatrixCSC{Float64, Int64},
                       Xhat0::Vector{<:LazySet{Float64}},
                       U::Union{ConstantInput, Nothing},
                       overapproximate::Function,
                       overapproximate_inputs::Function,
                       n::Int,
                       N::Union{Int, Nothing},
                       output_function::Union{Function, Nothing},
                       blocks::AbstractVector{Int},
                       partition::AbstractVector{<:Union{AbstractVector{Int}, Int}},
                       dimensions::Vector{Int},
                       δ::Float64,
                       termination::Function,
                       progress_meter::Union{Progress, Nothing},
                       res::Vector{<:AbstractReachSet}
                       )::Tuple{Int, Bool}
    box0 = boxof(Xhat0)
    if output_function !== nothing || N === nothing || box0 === nothing || !box_blocks(overapproximate) ||
       !plain_termination(termination)
        return reach_blocks!(Matrix{Float64}(ϕ), Xhat0, U, overapproximate, overapproximate_inputs, n, N, output_function,
                             blocks, partition, dimensions, δ, termination, progress_meter, res)
    end
    ctx = context()
    c0, r0 = box0
    m, MU, cU, rU, rψ = inputs_of(U, n)
    rows = Int32.(dimensions .- 1); nr = length(rows)
    sizes = Int32[length(partition[b]) for b in blocks]
    C = Matrix{Float64}(undef, nr, N); R = Matrix{Float64}(undef, nr, N); done = Ref{Int64}(0)
    check(ctx, ccall((:reach_bffpsv18_boxes_csc, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
        ctx.handle, n, nnz(ϕ), ϕ.colptr, ϕ.rowval, ϕ.nzval, c0, r0, m, MU, n, cU, rU, rψ, nr, rows, N, C, R, done))
    fill_box_result!(res, C, R, blocks, sizes, captured(overapproximate, :block_options_iter), dimensions, δ, N)
    return N, false
end

# ---- reach_blocks.jl:135-224 -- the registry's "explicit_blocks" function for a dense Φ; same 15 positional arguments ------
function reach_blocks!(ϕ::Matrix{Float64},
                       Xhat0::Vector{<:LazySet{Float64}},
                       U::Union{ConstantInput, Nothing},
                       overapproximate::Function,
                       overapproximate_inputs::Function,
                       n::Int,
                       N::Union{Int, Nothing},
                       output_function::Union{Function, Nothing},
                       blocks::AbstractVector{Int},
                       partition::AbstractVector{<:Union{AbstractVector{Int}, Int}},
                       dimensions::Vector{Int},
                       δ::Float64,
                       termination::Function,
                       progress_meter::Union{Progress, Nothing},
                       res::Vector{<:AbstractReachSet}
                       )::Tuple{Int, Bool}
    box0 = boxof(Xhat0)
    if N === nothing || box0 === nothing || !box_blocks(overapproximate) || !plain_termination(termination)
        ALLOW_REFERENCE || unsupported("reach_blocks! with non-box blocks, an invariant, or an unbounded horizon")
        return invoke(reach_blocks!, Tuple{AbstractMatrix{Float64}, Vector{<:LazySet{Float64}}, Union{ConstantInput, Nothing},
                                           Function, Function, Int, Union{Int, Nothing}, Union{Function, Nothing},
                                           AbstractVector{Int}, AbstractVector{<:Union{AbstractVector{Int}, Int}}, Vector{Int},
                                           Float64, Function, Union{Progress, Nothing}, Vector{<:AbstractReachSet}},
                      ϕ, Xhat0, U, overapproximate, overapproximate_inputs, n, N, output_function, blocks, partition,
                      dimensions, δ, termination, progress_meter, res)
    end
    # `overapproximate_inputs` (lazy_inputs_interval, reach.jl:113-147) cannot change the result for box blocks: the box of a
    # lazily accumulated Minkowski sum is the sum of the boxes (axis-direction support functions add up)
    ctx = context()
    c0, r0 = box0
    m, MU, cU, rU, rψ = inputs_of(U, n)
    rows = Int32.(dimensions .- 1); nr = length(rows)
    sizes = Int32[length(partition[b]) for b in blocks]
    t0, t1 = 0.0, δ
    if output_function === nothing
        C = Matrix{Float64}(undef, nr, N); R = Matrix{Float64}(undef, nr, N); done = Ref{Int64}(0)
        check(ctx, ccall((:reach_bffpsv18_boxes, LIB), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
            ctx.handle, n, ϕ, n, c0, r0, m, MU, n, cU, rU, rψ, nr, rows, N, C, R, done))
        fill_box_result!(res, C, R, blocks, sizes, captured(overapproximate, :block_options_iter), dimensions, δ, N)
    else                   # reach.jl:73-82: the stored set is box_approximation(M * CPA(Xhat_k)); blocks stay lazy
        Mfull = output_matrix(output_function); Mout = Mfull[:, dimensions]; q = size(Mout, 1)
        C = Matrix{Float64}(undef, q, N); R = Matrix{Float64}(undef, q, N)
        check(ctx, ccall((:reach_bffpsv18_output_map, LIB), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Int64, Ptr{Int32}, Int64, Ptr{Float64}, Int64,
             Ptr{Float64}, Ptr{Float64}),
            ctx.handle, n, ϕ, n, c0, r0, m, MU, n, cU, rU, rψ, nr, rows, N, length(sizes), sizes, q, Mout, q, C, R))
        for k in 1:N
            res[k] = SparseReachSet(Hyperrectangle(C[:, k], R[:, k]), t0, t1, dimensions)
            t0 = t1; t1 += δ
        end
    end
    return N, false        # (index, skip): termination_N stops at k == N and never skips (reach.jl:241-243)
end

# ---- check_blocks.jl:105-183 -- same 11 positional arguments; the device path covers half-space properties ---------------
# `is_contained_in(HalfSpace(ℓ, b))` after inout_map_property (partitions.jl:33-46): the predicate wraps the (projected)
# half-space; its field layout belongs to MathematicalPredicates.jl and is probed, not assumed.
function halfspace_of(prop)
    for name in fieldnames(typeof(prop))
        v = getfield(prop, name)
        v isa HalfSpace && return v
        if !(v isa Union{Number, AbstractString, Symbol, Function, AbstractArray}) && !isempty(fieldnames(typeof(v)))
            h = halfspace_of(v)
            h === nothing || return h
        end
    end
    return nothing
end
function check_blocks(ϕ::Matrix{Float64},
                      Xhat0::Vector{<:LazySet{Float64}},
                      U::Union{ConstantInput, Nothing},
                      overapproximate_inputs::Function,
                      n::Int,
                      N::Union{Int, Nothing},
                      blocks::AbstractVector{Int},
                      partition::AbstractVector{<:Union{AbstractVector{Int}, Int}},
                      eager_checking::Bool,
                      prop::Predicate,
                      progress_meter::Union{Progress, Nothing}
                      )::Int
    box0 = boxof(Xhat0); hs = halfspace_of(prop)
    if N === nothing || box0 === nothing || hs === nothing
        ALLOW_REFERENCE || unsupported("check_blocks with non-box initial blocks, a property that is not a half-space, or an unbounded horizon")
        return invoke(check_blocks, Tuple{AbstractMatrix{Float64}, Vector{<:LazySet{Float64}}, Union{ConstantInput, Nothing}, Function,
                                          Int, Union{Int, Nothing}, AbstractVector{Int}, AbstractVector{<:Union{AbstractVector{Int}, Int}},
                                          Bool, Predicate, Union{Progress, Nothing}},
                      ϕ, Xhat0, U, overapproximate_inputs, n, N, blocks, partition, eager_checking, prop, progress_meter)
    end
    ctx = context()
    c0, r0 = box0
    dims = reduce(vcat, [collect(partition[b]) for b in blocks])
    rows = Int32.(dims .- 1); sizes = Int32[length(partition[b]) for b in blocks]
    ℓ = Vector{Float64}(hs.a); length(ℓ) == length(dims) || error("the property has dimension $(length(ℓ)), the blocks $(length(dims))")
    m, MU, cU, rU, rψ = inputs_of(U, n)
    v = Ref{Int64}(0)
    check(ctx, ccall((:reach_bffpsv18_check, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Int64, Ptr{Int32}, Ptr{Float64}, Float64, Int32, Ref{Int64}),
        ctx.handle, n, ϕ, n, c0, r0, m, MU, n, cU, rU, rψ, length(rows), rows, N, length(sizes), sizes, ℓ, Float64(hs.b),
        eager_checking ? 1 : 0, v))
    return Int(v[])        # violation index, 0 = none (check_blocks.jl:116)
end

# ---- GLGM06/reach.jl:5-24 and :30-62 -- device flowpipe; all N zonotopes stream back while later batches compute ---------
function flowpipe!(R, Ω0::Zonotope, V::Union{Zonotope,Nothing}, Φ::AbstractMatrix, N::Int, δ::Float64, max_order::Int)
    ctx = context(); Φd = dense(Φ); n = size(Φd, 1)
    c0 = Vector{Float64}(Ω0.center); G0 = Matrix{Float64}(Ω0.generators); p0 = size(G0, 2)
    cV, GV, pV = V === nothing ? (C_NULL, C_NULL, 0) : (Vector{Float64}(V.center), Matrix{Float64}(V.generators), size(V.generators, 2))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:reach_glgm06_create, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64},
         Int64, Int64, Int64, Int32, Ref{Ptr{Cvoid}}),
        ctx.handle, n, Φd, n, p0, c0, G0, n, pV, cV, GV, n, N, max_order, 0, h))
    try
        pmax = V === nothing ? p0 : max(p0, max_order * n)      # reduce_order caps every set at max_order * n generators
        C = Matrix{Float64}(undef, n, N); G = Array{Float64,3}(undef, n, pmax, N)
        check(ctx, ccall((:reach_glgm06_run_fetch, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), h[], C, G, pmax))
        p = Vector{Int32}(undef, N)
        check(ctx, ccall((:reach_flowpipe_ngens, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), h[], p))
        t0, t1 = 0.0, δ
        for k in 1:N
            R[k] = ReachSet(Zonotope(C[:, k], G[:, 1:p[k], k]), t0, t1)
            t0 = t1; t1 += δ
        end
    finally
        ccall((:reach_flowpipe_free, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    return R
end
reach_homog!(R::Vector{ReachSet{<:Zonotope{Float64}}}, Ω0::Zonotope, Φ::Matrix{Float64}, N::Int, δ::Float64) =
    flowpipe!(R, Ω0, nothing, Φ, N, δ, 1)
function reach_inhomog!(R::Vector{ReachSet{<:Zonotope{Float64}}}, Ω0::Zonotope, U::ConstantInput, Φ::Matrix{Float64}, N::Int,
                        δ::Float64, max_order::Int)
    V = next_set(U)
    V isa Zonotope || unsupported("reach_inhomog! with an input set of type $(typeof(V))")
    return flowpipe!(R, Ω0, V, Φ, N, δ, max_order)
end

# ---- GLGM06/post.jl:20 -- Ω0red = ReachSM100a.reduce_order(Ω0, max_order)
function reduce_order(Z::Zonotope, r::Int)
    ctx = context(); n, p = size(Z.generators)
    G = Matrix{Float64}(Z.generators); Gout = Matrix{Float64}(undef, n, max(p, r * n)); pout = Ref{Int64}(0)
    check(ctx, ccall((:reach_reduce_order, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int32, Ptr{Float64}, Int64, Ref{Int64}, Ptr{Int32}),
        ctx.handle, n, p, Vector{Float64}(Z.center), G, n, r, 0, Gout, n, pout, C_NULL))
    return Zonotope(Z.center, Gout[:, 1:pout[]])
end

# ---- project_reach.jl:18-99 -- zonotope flowpipe projected on the device; `h` is a live reach_flowpipe handle
#      (keep it instead of fetching every zonotope when only `project(sol)` is wanted)
function project_flowpipe(ctx::Context, h::Ptr{Cvoid}, M::Matrix{Float64}, N::Int)
    q = size(M, 1)
    C = Matrix{Float64}(undef, q, N); R = Matrix{Float64}(undef, q, N)
    check(ctx, ccall((:reach_flowpipe_project, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64), h, q, M, q, C, R, C_NULL, 0))
    return [Hyperrectangle(C[:, k], R[:, k]) for k in 1:N]
end

# ---- exp_method = "lazy": SparseMatrixExp(A*δ) never leaves the device (discretize.jl:153-154, reach_blocks.jl:227-397)
mutable struct SparseSystem
    handle::Ptr{Cvoid}
    n::Int
    function SparseSystem(A::SparseMatrixCSC{Float64,Int64})
        ctx = context(); h = Ref{Ptr{Cvoid}}(C_NULL); n = size(A, 1)
        check(ctx, ccall((:reach_sparse_create, LIB), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
            ctx.handle, n, nnz(A), A.colptr, A.rowval, A.nzval, h))
        s = new(h[], n)
        finalizer(x -> ccall((:reach_sparse_free, LIB), Cvoid, (Ptr{Cvoid},), x.handle), s)
        return s
    end
end

# rows of exp(δA): get_rows(SparseMatrixExp(A*δ), rows) = (e^{δAᵀ} E_rows)ᵀ
function get_rows(S::SparseSystem, δ::Float64, rows::Vector{Int})
    E = zeros(S.n, length(rows)); for (l, i) in enumerate(rows); E[i, l] = 1.0; end
    Y = similar(E)
    check(context(), ccall((:reach_expmv, LIB), Int32,
        (Ptr{Cvoid}, Float64, Int32, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64), S.handle, δ, 1, length(rows), E, S.n, Y, S.n))
    return permutedims(Y)
end

# reach_blocks!(ϕ::SparseMatrixExp, assume_sparse, ...) for box blocks (reach_blocks.jl:227-397; 16 positional arguments: the
# `Val(assume_sparse)` of reach.jl:45-47 follows ϕ): same marshalling as the dense method above, without Φ.  `A_δ = ϕ.M` is the
# exponent matrix A*δ the reference wrapped (discretize.jl:153-154), so A = ϕ.M / δ.
function reach_blocks!(ϕ::SparseMatrixExp{Float64}, assume_sparse::Val, Xhat0::Vector{<:LazySet{Float64}},
                       U::Union{ConstantInput, Nothing}, overapproximate::Function, overapproximate_inputs::Function, n::Int,
                       N::Union{Int, Nothing}, output_function::Union{Function, Nothing}, blocks::AbstractVector{Int},
                       partition::AbstractVector{<:Union{AbstractVector{Int}, Int}}, dimensions::Vector{Int}, δ::Float64,
                       termination::Function, progress_meter::Union{Progress, Nothing},
                       res::Vector{<:AbstractReachSet})::Tuple{Int, Bool}
    box0 = boxof(Xhat0)
    (N === nothing || box0 === nothing || !box_blocks(overapproximate) || !plain_termination(termination) ||
     output_function !== nothing) && unsupported("this lazy-exponential reach_blocks! configuration")
    S = SparseSystem(SparseMatrixCSC{Float64,Int64}(ϕ.M ./ δ))
    c0, r0 = box0; rows = Int32.(dimensions .- 1); nr = length(rows)
    C = Matrix{Float64}(undef, nr, N); R = Matrix{Float64}(undef, nr, N); done = Ref{Int64}(0)
    m, MU, cU, rU, rψ = inputs_of(U, n)
    check(context(), ccall((:reach_bffpsv18_boxes_lazy, LIB), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Int64, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
        S.handle, δ, c0, r0, m, MU, S.n, cU, rU, rψ, nr, rows, N, C, R, done))
    opts = captured(overapproximate, :block_options_iter)
    t0, t1 = 0.0, δ
    for k in 1:N
        arr = Vector{LazySet{Float64}}(undef, length(blocks)); o = 0
        for (i, b) in enumerate(blocks)
            len = length(partition[b]); T = opts isa Union{AbstractVector, Dict} ? opts[b] : opts
            arr[i] = T === Interval ? Interval(C[o+1, k] - R[o+1, k], C[o+1, k] + R[o+1, k]) :
                                      Hyperrectangle(C[o+1:o+len, k], R[o+1:o+len, k])
            o += len
        end
        res[k] = SparseReachSet(CartesianProductArray(arr), t0, t1, dimensions)
        t0 = t1; t1 += δ
    end
    return N, false
end

end # module
