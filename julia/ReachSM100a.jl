# ReachSM100a.jl -- the thin `ccall` layer a Reachability.jl maintainer adds to route the linear-flowpipe hot path
# through libreach_sm100a.so (include/reach.h).  Nothing here computes: it marshals Julia arrays (column-major
# Float64, 1-based indices -> 0-based Int32), checks the int32 status and rebuilds the result containers.
#
# Seams replaced (file:line under Reachability.jl v0.7.0 `src/`):
#   exp_Aδ / Φ₁ / Φ₂                 ReachSets/discretize.jl:150-160, 221-240, 296-315   (new `exp_method = "sm100a"`)
#   reach_blocks! (dense, box types)  ReachSets/ContinuousPost/BFFPSV18/reach_blocks.jl:135-224, dispatched at reach.jl:191-193
#   check_blocks  (dense, half-space) ReachSets/ContinuousPost/BFFPSV18/check_blocks.jl:105-183
#   reach_homog! / reach_inhomog!     ReachSets/ContinuousPost/GLGM06/reach.jl:5-62, called at GLGM06/post.jl:32,36
#   reduce_order (call site)          ReachSets/ContinuousPost/GLGM06/post.jl:20
#   project_reach (zonotopes)         ReachSets/project_reach.jl:18-99
#   SparseMatrixExp / lazy rows       ReachSets/discretize.jl:153-154, 302-306; BFFPSV18/reach_blocks.jl:227-397
#
# No Julia toolchain exists in the build container or on the GPU box, so this file is NOT exercised by the test suite;
# the Python ctypes binding (reachability.jl_b200/_lib.py) drives the identical entry points and is what the parity
# tests run.  There is no CPU fallback: every non-zero status becomes `error(...)`.
module ReachSM100a

using LazySets, SparseArrays
import Reachability
import Reachability.ReachSets: reach_blocks!, reach_homog!, reach_inhomog!, ReachSet, SparseReachSet

const LIB = get(ENV, "LIBREACH_SM100A", "libreach_sm100a.so")

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:reach_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h)
        rc == 0 || error("libreach_sm100a: ", unsafe_string(ccall((:reach_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[])
        finalizer(c -> ccall((:reach_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

const CTX = Ref{Union{Context,Nothing}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context(parse(Int, get(ENV, "REACH_DEVICE", "0")))); CTX[])

check(ctx, rc) = rc == 0 || error("libreach_sm100a status $rc: ",
    unsafe_string(ccall((:reach_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))

dense(A::AbstractMatrix) = Matrix{Float64}(A)          # the reference densifies too: exp(Matrix(A*δ)), discretize.jl:152

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

# box (centre, radius) of a box-shaped LazySet, all n coordinates
boxof(X::AbstractHyperrectangle) = (Vector{Float64}(center(X)), Vector{Float64}(radius_hyperrectangle(X)))
function boxof(Xs::Vector{<:LazySet})
    cs = Float64[]; rs = Float64[]
    for X in Xs
        c, r = boxof(X); append!(cs, c); append!(rs, r)
    end
    return cs, rs
end

# ---- reach_blocks.jl:135-224 -- a more specific method of the registry's "explicit_blocks" function (reach.jl:191-193):
#      dense Φ, Interval / Hyperrectangle blocks, ConstantInput (a box through B) or no input, Universe invariant.
#      `inputs` carries the discretised input V = (δB) * Box(cU, rU) ⊕ Box(0, rψ) as (MU, cU, rU, rψ), which the shim
#      reads off `Ud = ConstantInput(δ*U0 ⊕ Eψ0)` (discretize.jl:688).
function reach_blocks!(ϕ::Matrix{Float64}, Xhat0::Vector{<:Union{Interval{Float64},Hyperrectangle{Float64}}},
                       inputs::Union{Nothing,NamedTuple}, n::Int, N::Int, blocks::AbstractVector{Int},
                       partition, dimensions::Vector{Int}, δ::Float64, res::Vector)::Tuple{Int,Bool}
    ctx = context()
    c0, r0 = boxof(Xhat0)
    rows = Int32.(dimensions .- 1)
    nr = length(rows)
    C = Matrix{Float64}(undef, nr, N); R = Matrix{Float64}(undef, nr, N); done = Ref{Int64}(0)
    m, MU, cU, rU, rψ = inputs === nothing ? (0, C_NULL, C_NULL, C_NULL, C_NULL) :
        (size(inputs.MU, 2), inputs.MU, inputs.cU, inputs.rU, inputs.rψ)
    check(ctx, ccall((:reach_bffpsv18_boxes, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
        ctx.handle, n, ϕ, n, c0, r0, m, MU, n, cU, rU, rψ, nr, rows, N, C, R, done))
    # regroup the rows into the blocks' set types (store!, reach_blocks.jl:41-44)
    t0, t1 = 0.0, δ
    for k in 1:N
        arr = Vector{LazySet{Float64}}(undef, length(blocks)); o = 0
        for (i, b) in enumerate(blocks)
            len = length(partition[b])
            arr[i] = len == 1 && Xhat0[b] isa Interval ?
                Interval(C[o+1, k] - R[o+1, k], C[o+1, k] + R[o+1, k]) : Hyperrectangle(C[o+1:o+len, k], R[o+1:o+len, k])
            o += len
        end
        res[k] = SparseReachSet(CartesianProductArray(arr), t0, t1, dimensions)
        t0 = t1; t1 += δ
    end
    return Int(done[]), false
end

# ---- check_blocks.jl:105-183 for `is_contained_in(HalfSpace(ℓ, b))`: returns the violation index (0 = none)
function check_blocks(ϕ::Matrix{Float64}, Xhat0, inputs, n::Int, N::Int, blocks, partition, eager_checking::Bool,
                      ℓ::Vector{Float64}, b::Float64)::Int
    ctx = context()
    c0, r0 = boxof(Xhat0)
    dims = vcat(partition[blocks]...)
    rows = Int32.(dims .- 1); sizes = Int32[length(partition[bi]) for bi in blocks]
    m, MU, cU, rU, rψ = inputs === nothing ? (0, C_NULL, C_NULL, C_NULL, C_NULL) :
        (size(inputs.MU, 2), inputs.MU, inputs.cU, inputs.rU, inputs.rψ)
    v = Ref{Int64}(0)
    check(ctx, ccall((:reach_bffpsv18_check, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Int64, Ptr{Int32}, Ptr{Float64}, Float64, Int32, Ref{Int64}),
        ctx.handle, n, ϕ, n, c0, r0, m, MU, n, cU, rU, rψ, length(rows), rows, N, length(sizes), sizes, ℓ[dims], b,
        eager_checking ? 1 : 0, v))
    return Int(v[])
end

# ---- GLGM06/reach.jl:5-24 and :30-62 -- device flowpipe, every zonotope fetched back into `R`
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
        check(ctx, ccall((:reach_glgm06_run, LIB), Int32, (Ptr{Cvoid},), h[]))
        p = Vector{Int32}(undef, N)
        check(ctx, ccall((:reach_flowpipe_ngens, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), h[], p))
        t0, t1 = 0.0, δ
        for k in 1:N
            c = Vector{Float64}(undef, n); G = Matrix{Float64}(undef, n, p[k])
            check(ctx, ccall((:reach_flowpipe_step, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
                             h[], k, c, G, n))
            R[k] = ReachSet(Zonotope(c, G), t0, t1)
            t0 = t1; t1 += δ
        end
    finally
        ccall((:reach_flowpipe_free, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    return R
end
reach_homog!(R::Vector{ReachSet{<:Zonotope{Float64}}}, Ω0::Zonotope, Φ::Matrix{Float64}, N::Int, δ::Float64) =
    flowpipe!(R, Ω0, nothing, Φ, N, δ, 1)
reach_inhomog!(R::Vector{ReachSet{<:Zonotope{Float64}}}, Ω0::Zonotope, U::ConstantInput, Φ::Matrix{Float64}, N::Int,
               δ::Float64, max_order::Int) = flowpipe!(R, Ω0, next_set(U), Φ, N, δ, max_order)

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

# reach_blocks!(ϕ::SparseMatrixExp, ...) for box blocks: same marshalling as the dense method above, without Φ
function reach_blocks_lazy!(S::SparseSystem, δ::Float64, Xhat0, inputs, N::Int, dimensions::Vector{Int})
    c0, r0 = boxof(Xhat0); rows = Int32.(dimensions .- 1); nr = length(rows)
    C = Matrix{Float64}(undef, nr, N); R = Matrix{Float64}(undef, nr, N); done = Ref{Int64}(0)
    m, MU, cU, rU, rψ = inputs === nothing ? (0, C_NULL, C_NULL, C_NULL, C_NULL) :
        (size(inputs.MU, 2), inputs.MU, inputs.cU, inputs.rU, inputs.rψ)
    check(context(), ccall((:reach_bffpsv18_boxes_lazy, LIB), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Int64, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
        S.handle, δ, c0, r0, m, MU, S.n, cU, rU, rψ, nr, rows, N, C, R, done))
    return C, R          # regrouped into SparseReachSet{CartesianProductArray} exactly as in reach_blocks! above
end

end # module
