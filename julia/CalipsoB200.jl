# CalipsoB200.jl -- ccall glue binding libcalipso_b200.so (include/calipso_b200.h) into CALIPSO.jl's src/solver seams.
#
# Usage inside CALIPSO.jl:   include("CalipsoB200.jl"); using .CalipsoB200
#   * B200LDLSolver <: LinearSolver with ldl_solver / factorize! / compute_inertia! / linear_solve!
#     (src/solver/linear_solver.jl:1-60), selected by options.linear_solver == :B200 (options.jl:42, search_direction.jl:6);
#   * B200Newton: the per-iteration calls of solve! (solve.jl:127,187,190-221,309-333) on the device, evaluate! stays on the host.
#
# There is no Julia toolchain in the image this repository is built in: the file is checked mechanically
# (tests/test_julia_glue.py parses every ccall -- symbol, return type, argument types -- and every CB200_* constant against
# include/calipso_b200.h), not executed.  Indices cross the boundary 0-based Int32; values are Float64.
module CalipsoB200

using SparseArrays

const LIBCB200 = get(ENV, "CALIPSO_B200_LIB", "libcalipso_b200")

# ---- enum values of include/calipso_b200.h (checked by tests/test_julia_glue.py)
const CB200_POINT = Cint(0)
const CB200_CANDIDATE = Cint(1)
const CB200_STEP = Cint(2)
const CB200_RESIDUAL = Cint(3)
const CB200_GRADIENT = Cint(4)
const CB200_EQ_DUAL_GRAD = Cint(5)
const CB200_CONE_DUAL_GRAD = Cint(6)
const CB200_EQUALITY = Cint(7)
const CB200_CONE = Cint(8)
const CB200_W_VALUES = Cint(9)
const CB200_G_VALUES = Cint(10)
const CB200_C_VALUES = Cint(11)
const CB200_CONE_PRODUCT = Cint(12)
const CB200_BARRIER_GRADIENT = Cint(13)
const CB200_DUAL = Cint(14)
const CB200_SCALARS = Cint(18)
const CB200_MATRIX_VALUES = Cint(23)
const CB200_RHS = Cint(24)
const CB200_S_KAPPA = 0
const CB200_S_TAU = 1
const CB200_S_RHO = 2
const CB200_S_BARRIER = 7
const CB200_S_RESIDUAL_VIOLATION = 8
const CB200_S_OPTIMALITY_VIOLATION = 9
const CB200_S_SLACK_VIOLATION = 10
const CB200_S_STEP_SIZE = 13
const CB200_S_STEP_SIZE_T = 14
const CB200_S_EQUALITY_VIOLATION = 15
const CB200_S_CONE_PRODUCT_VIOLATION = 16
const CB200_S_COUNT = 24
const CB200_I_STATUS = 8
const CB200_I_COUNT = 24
const CB200_OK = 0
const CB200_INERTIA_FAILURE = 1
const CB200_REFINEMENT_FAILURE = 2
const CB200_CONE_SEARCH_FAILURE = 3
const CB200_CONE_BARRIER = Cint(1)
const CB200_CONE_BARRIER_GRADIENT = Cint(2)
const CB200_CONE_PRODUCT_FLAG = Cint(4)

last_error() = unsafe_string(ccall((:cb200_last_error, LIBCB200), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error(last_error())

to0(v::AbstractVector{<:Integer}) = Cint.(v .- 1)       # 1-based Int64 -> 0-based Int32

# ------------------------------------------------------------------------------------------------ LinearSolver seam
abstract type LinearSolver end           # (inside CALIPSO.jl: the package's own abstract type, linear_solver.jl:1)

mutable struct Inertia                   # src/solver/inertia.jl:1-5
    positive::Int
    negative::Int
    zero::Int
end

mutable struct B200LDLSolver <: LinearSolver
    handle::Ptr{Cvoid}
    A_sparse::SparseMatrixCSC{Float64,Int}      # upper triangle, as triu!(A) leaves it (linear_solver.jl:23)
    inertia::Inertia
end

# amd(A), qdldl.jl:135: the reference's ordering, computed by the library (host only)
function b200_amd(A::SparseMatrixCSC{Float64,Int})
    n = size(A, 1)
    perm = zeros(Cint, n)
    cp = to0(A.colptr); ri = to0(A.rowval)
    check(ccall((:cb200_amd_order, LIBCB200), Cint, (Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), n, cp, ri, perm))
    return Int.(perm) .+ 1
end

# ldl_solver(A), linear_solver.jl:46-48 -> cb200_ldl_create (ordering + symbolic analysis = qdldl(A), qdldl.jl:134-188).
# perm = nothing: the library's device-friendly minimum-degree ordering; perm = b200_amd(A): the reference's own.
function b200_ldl_solver(A::SparseMatrixCSC{Float64,Int}; device::Integer=0, perm=nothing, batch::Integer=1)
    U = triu(A)
    cp = to0(U.colptr); ri = to0(U.rowval)
    p32 = perm === nothing ? Ptr{Cint}(C_NULL) : pointer(to0(perm))
    h = ccall((:cb200_ldl_create, LIBCB200), Ptr{Cvoid}, (Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Cint),
              batch, size(U, 1), cp, ri, p32, device)
    h == C_NULL && error(last_error())
    s = B200LDLSolver(h, U, Inertia(0, 0, 0))
    finalizer(x -> ccall((:cb200_destroy, LIBCB200), Cvoid, (Ptr{Cvoid},), x.handle), s)
    return s
end

# factorize!(s, A; update), linear_solver.jl:19-31.  update=false re-runs qdldl(A) in the reference; the pattern is fixed
# after Solver construction, so both branches reuse the symbolic analysis held by the handle.
function factorize!(s::B200LDLSolver, A::SparseMatrixCSC{Float64,Int}; update=false)
    triu!(A)
    GC.@preserve A begin
        check(ccall((:cb200_set_array, LIBCB200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint),
                    s.handle, CB200_MATRIX_VALUES, A.nzval, 0, 1))
        check(ccall((:cb200_ldl_factorize, LIBCB200), Cint, (Ptr{Cvoid},), s.handle))
    end
    return nothing
end

# compute_inertia!(s), linear_solver.jl:33-44
function compute_inertia!(s::B200LDLSolver)
    out = zeros(Cint, 3)
    check(ccall((:cb200_ldl_inertia, LIBCB200), Cint, (Ptr{Cvoid}, Ptr{Cint}), s.handle, out))
    s.inertia.positive, s.inertia.negative, s.inertia.zero = out
    return nothing
end

# linear_solve!(s, x, A, b; fact, update), linear_solver.jl:52-60
function linear_solve!(s::B200LDLSolver, x::Vector{Float64}, A::SparseMatrixCSC{Float64,Int}, b::Vector{Float64};
                       fact=true, update=true)
    fact && triu!(A)
    GC.@preserve x A b check(ccall((:cb200_ldl_linear_solve, LIBCB200), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint), s.handle, A.nzval, b, x, fact ? 1 : 0))
    return nothing
end

# ------------------------------------------------------------------------------------------------ the Newton step on the device
mutable struct B200Newton
    handle::Ptr{Cvoid}
    n::Int; m::Int; p::Int
    scalars::Vector{Float64}
    stats::Vector{Cint}
end

# Solver(methods, n, theta, m, p; nonnegative_indices, second_order_indices, options), solver.jl:46-150: the patterns of
# the Lagrangian Hessian (upper triangle), of the equality and of the cone Jacobian are given once (CSC, 1-based).
function B200Newton(n::Integer, m::Integer, p::Integer, num_nonnegative::Integer, second_order_dims::Vector{Int},
                    W::SparseMatrixCSC, G::SparseMatrixCSC, C::SparseMatrixCSC; device::Integer=0, perm=nothing)
    soc = Cint.(second_order_dims)
    Wp = to0(W.colptr); Wi = to0(W.rowval); Gp = to0(G.colptr); Gi = to0(G.rowval); Cp = to0(C.colptr); Ci = to0(C.rowval)
    p32 = perm === nothing ? Ptr{Cint}(C_NULL) : pointer(to0(perm))
    h = ccall((:cb200_create, LIBCB200), Ptr{Cvoid},
              (Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint},
               Ptr{Cint}, Ptr{Cvoid}, Cint),
              1, n, m, p, num_nonnegative, length(soc), soc, Wp, Wi, Gp, Gi, Cp, Ci, p32, C_NULL, device)
    h == C_NULL && error(last_error())
    s = B200Newton(h, n, m, p, zeros(CB200_S_COUNT), zeros(Cint, CB200_I_COUNT))
    finalizer(x -> ccall((:cb200_destroy, LIBCB200), Cvoid, (Ptr{Cvoid},), x.handle), s)
    return s
end

upload!(s::B200Newton, which::Cint, v::Vector{Float64}) = GC.@preserve v check(ccall((:cb200_set_array, LIBCB200), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint), s.handle, which, v, 0, 1))
download!(s::B200Newton, which::Cint, v::Vector{Float64}) = GC.@preserve v check(ccall((:cb200_get_array, LIBCB200), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint), s.handle, which, v, 0, 1))
function stats!(s::B200Newton)
    check(ccall((:cb200_get_stats, LIBCB200), Cint, (Ptr{Cvoid}, Ptr{Cint}, Cint, Cint), s.handle, s.stats, 0, 1))
    return s.stats
end

# initialize!(solver, guess), initialize.jl:9-13
initialize!(s::B200Newton, guess::Vector{Float64}) = GC.@preserve guess check(ccall((:cb200_initialize, LIBCB200), Cint,
    (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint), s.handle, guess, 0, 1))

# cone!(problem, methods, idx, solution; barrier, barrier_gradient, product), cones/cone.jl:71-106
cone!(s::B200Newton; barrier=false, barrier_gradient=false, product=false, at_candidate=false) =
    check(ccall((:cb200_cone, LIBCB200), Cint, (Ptr{Cvoid}, Cint, Cint), s.handle,
                (barrier ? CB200_CONE_BARRIER : Cint(0)) | (barrier_gradient ? CB200_CONE_BARRIER_GRADIENT : Cint(0)) |
                (product ? CB200_CONE_PRODUCT_FLAG : Cint(0)), at_candidate ? 1 : 0))

# residual!(data, problem, idx, solution, kappa, rho, lambda), residual.jl:1-51 + the norms of solve.jl:130-135
residual!(s::B200Newton) = check(ccall((:cb200_residual, LIBCB200), Cint, (Ptr{Cvoid},), s.handle))

# search_direction!(solver), search_direction.jl:1-23; the reference's errors / warnings are re-raised from the status
function search_direction!(s::B200Newton)
    check(ccall((:cb200_search_direction, LIBCB200), Cint, (Ptr{Cvoid},), s.handle))
    st = stats!(s)[CB200_I_STATUS + 1]
    st == CB200_INERTIA_FAILURE && error("inertia correction failure")                    # inertia.jl:72
    st == CB200_REFINEMENT_FAILURE && @warn "iterative refinement failure"                 # iterative_refinement.jl:50
    return nothing
end

# cone line search, solve.jl:190-221
function cone_search!(s::B200Newton)
    check(ccall((:cb200_cone_search, LIBCB200), Cint, (Ptr{Cvoid},), s.handle))
    stats!(s)[CB200_I_STATUS + 1] == CB200_CONE_SEARCH_FAILURE && error("cone search failure")      # solve.jl:210,220
    return nothing
end

# step update, solve.jl:309-333
apply_step!(s::B200Newton) = check(ccall((:cb200_apply_step, LIBCB200), Cint, (Ptr{Cvoid},), s.handle))

# differentiate!(solver), differentiate.jl:1-61: jacobian_parameters is total x num_parameters (column-major = one
# right-hand side after the other, as the ABI wants them); returns solution_sensitivity in the same layout
function differentiate!(s::B200Newton, jacobian_parameters::Matrix{Float64})
    S = similar(jacobian_parameters)
    GC.@preserve jacobian_parameters S check(ccall((:cb200_differentiate, LIBCB200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, size(jacobian_parameters, 2), jacobian_parameters, S))
    return S
end

# evaluate!'s `=` scatter of the flat derivative caches (evaluate.jl:37-42,73-78,109-114) on the device
function scatter_plan!(s::B200Newton, which::Cint, sparsities::Vector{Vector{Tuple{Int,Int}}})
    len = Cint[length(v) for v in sparsities]
    rows = Cint[k[1] - 1 for v in sparsities for k in v]
    cols = Cint[k[2] - 1 for v in sparsities for k in v]
    check(ccall((:cb200_scatter_plan, LIBCB200), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}),
                s.handle, which, length(len), len, rows, cols))
end
scatter!(s::B200Newton, which::Cint, caches::Vector{Float64}) = GC.@preserve caches check(ccall((:cb200_scatter, LIBCB200), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint), s.handle, which, caches, 0, 1))

# the front end's stage loops (trajectory_optimization/evaluate.jl:15-28,77-136,206-241,297-327) on the device: `indices` = the
# stages' (1-based) index lists in program order, accumulate = true for `gradient[idx...] += cache[i]`, false for
# `violations[indices[t]] .= cache`; `caches` = the stage caches concatenated in the same order
function stage_plan!(s::B200Newton, which::Cint, indices::Vector{Vector{Int}}, accumulate::Bool)
    dst = Cint[i - 1 for v in indices for i in v]
    check(ccall((:cb200_stage_plan, LIBCB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}),
                s.handle, which, accumulate ? 1 : 0, length(dst), dst))
end
stage_scatter!(s::B200Newton, which::Cint, caches::Vector{Float64}) = GC.@preserve caches check(ccall((:cb200_stage_scatter, LIBCB200), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint), s.handle, which, caches, 0, 1))

synchronize(s::B200Newton) = check(ccall((:cb200_synchronize, LIBCB200), Cint, (Ptr{Cvoid},), s.handle))

export B200LDLSolver, b200_ldl_solver, b200_amd, factorize!, compute_inertia!, linear_solve!, B200Newton, initialize!, cone!,
       residual!, search_direction!, cone_search!, apply_step!, differentiate!, scatter_plan!, scatter!, stage_plan!, stage_scatter!

end # module
