# BazingaB200.jl -- thin `ccall` shim that makes libbazinga_b200.so a drop-in for the PANOC+/ALPS hot
# path of Bazinga.jl.  It mirrors bazinga.jl_b200/solver.py line by line (that Python mirror is what the
# test-suite drives: this image has no Julia, so THIS FILE HAS NOT BEEN EXECUTED; keep it thin).
#
#   using Bazinga, BazingaB200
#   out = BazingaB200.alps(f, g, c, D, x0, y0; tol = 1e-6, subsolver = BazingaB200.B200PANOCplus(...))
#
# returns the reference's 10-tuple (src/algorithms/alps.jl:115).  Dispatch is on the concrete oracle
# types; anything outside the catalogue raises (no CPU fallback).
module BazingaB200

using Bazinga
import ProximalOperators

const lib = get(ENV, "BAZINGA_B200_LIB", joinpath(@__DIR__, "..", "lib", "libbazinga_b200.so"))

# ---- mirrors of the C structs in include/bazinga_b200.h (field order and types must match) ----------
struct ProblemDesc
    f_kind::Int32; g_kind::Int32; c_kind::Int32; d_kind::Int32
    nx::Int64; ny::Int64
    f_A::Ptr{Float64}; f_b::Ptr{Float64}; f_m::Int64; f_layout::Int32
    f_p::NTuple{4,Float64}
    g_lambda::Float64; g_lo::Float64; g_hi::Float64; g_u::Ptr{Float64}
    c_A::Ptr{Float64}; c_b::Ptr{Float64}; c_layout::Int32
    c_A_is_device::Int32; c_gen::Int32; c_gen_seed::UInt64
    grid_y::Int64; grid_x::Int64
    d_lo::Float64; d_hi::Float64; d_pair_stride::Int64
    row_begin::Int64; ny_local::Int64
    slab_nranks::Int32; slab_rank::Int32; nx_global::Int64
    slack_formulation::Int32
end

mutable struct Options
    tol::Float64; tol_prim::Float64; tol_dual::Float64; inner_tol::Float64
    maxit::Int32
    theta_penalty::Float64; kappa_penalty::Float64; kappa_tol::Float64
    subsolver_maxit::Int64
    alpha::Float64; beta::Float64; minimum_gamma::Float64
    max_backtracks::Int32; lbfgs_memory::Int32
    sub_maxit::Int64; max_inner_total::Int64; verbose::Int32
    dual_safeguard::Ptr{Cvoid}; dual_safeguard_user::Ptr{Cvoid}   # alps.jl:23,62 as a host callback (C_NULL = default clamp)
    Options() = new()
end

mutable struct Result
    tot_it::Int32; tot_inner_it::Int64; elapsed_time::Float64; status::Int32
    inner_tol::Float64; norm_res_prim::Float64; objective::Float64
    n_grad::Int64; n_prox::Int64; n_launches::Int64
    dominant_kernel_ms::Float64; dominant_kernel_launches::Int64
    kind_ms::NTuple{4,Float64}; kind_launches::NTuple{4,Int64}
    mu0_min::Float64; mu0_max::Float64; inner_tol0::Float64       # verbose: the two @info lines of alps.jl:52-55
    device_ms::Float64
    Result() = new()
end

const STATUS = (:first_order, :max_iter, :exception, :unknown)   # alps.jl:105-113

function check(rc::Cint, ctx::Ptr{Cvoid} = C_NULL)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:bz_last_error, lib), Cstring, (Ptr{Cvoid},), ctx))
    error(msg)                    # the reference reports guard failures with error(...)
end

# ---- context ------------------------------------------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bz_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[])
        finalizer(c -> ccall((:bz_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
context() = (default_ctx[] === nothing && (default_ctx[] = Context(0)); default_ctx[])

# ---- the subsolver plugin slot (alps.jl:24,64,66) -----------------------------------------------------------
struct B200PANOCplus
    memory::Int; maxit::Int; tol::Float64; minimum_gamma::Float64
    alpha::Float64; beta::Float64; max_backtracks::Int; verbose::Bool
end
B200PANOCplus(; directions = nothing, maxit = 1000, tol = 1e-8, minimum_gamma = 1e-7, alpha = 0.95,
              beta = 0.5, max_backtracks = 20, verbose = false, freq = 10) =
    B200PANOCplus(directions === nothing ? 5 : lbfgs_memory(directions), maxit, tol, minimum_gamma, alpha,
                  beta, max_backtracks, verbose)
lbfgs_memory(d) = hasproperty(d, :memory) ? Int(d.memory) : 0        # LBFGS(m) / NoAcceleration()
# `subsolver(tol = inner_tol, verbose = verbose)` of alps.jl:64
(s::B200PANOCplus)(; tol = s.tol, verbose = s.verbose) =
    B200PANOCplus(s.memory, s.maxit, tol, s.minimum_gamma, s.alpha, s.beta, s.max_backtracks, verbose)

# ---- oracle catalogue: concrete Julia type -> kind enum + parameters -----------------------------------------
f_kind(::ProximalOperators.Zero) = (Int32(0), NamedTuple())
f_kind(f::ProximalOperators.LeastSquares) = (Int32(1), (A = Matrix{Float64}(f.A), b = Vector{Float64}(f.b)))
f_kind(f::ProximalOperators.Quadratic) = (Int32(2), (A = Matrix{Float64}(f.Q), b = Vector{Float64}(f.q)))
f_kind(f) = error("BazingaB200: f = $(typeof(f)) is not in the GPU oracle catalogue (no CPU fallback)")
g_kind(::ProximalOperators.Zero) = (Int32(0), 1.0, 0.0, 0.0, nothing)
g_kind(::ProximalOperators.IndFree) = (Int32(0), 1.0, 0.0, 0.0, nothing)
g_kind(g::ProximalOperators.NormL1) = (Int32(1), Float64(g.lambda), 0.0, 0.0, nothing)
g_kind(g::ProximalOperators.NormL0) = (Int32(2), Float64(g.lambda), 0.0, 0.0, nothing)
g_kind(g::ProximalOperators.IndBox) = (Int32(3), 1.0, Float64(g.lb), Float64(g.ub), nothing)
g_kind(g::Bazinga.NormL1Box) = (Int32(4), Float64(g.lambda), 0.0, 0.0, Vector{Float64}(g.u))
g_kind(g::Bazinga.NormL1Nonneg) = (Int32(5), Float64(g.lambda), 0.0, 0.0, nothing)
g_kind(g::Bazinga.NormL0Box) = (Int32(6), Float64(g.lambda), 0.0, 0.0, Vector{Float64}(g.u))
g_kind(g::Bazinga.NormLpPowerBox) = (Int32(10), Float64(g.alpha), Float64(g.p), 0.0, Vector{Float64}(g.u))   # g_lo = p
g_kind(g::Bazinga.NormLpPowerNonneg) = (Int32(11), Float64(g.alpha), Float64(g.p), 0.0, nothing)
g_kind(g) = error("BazingaB200: g = $(typeof(g)) is not in the GPU oracle catalogue (no CPU fallback)")
d_kind(::Bazinga.ZeroSet) = (Int32(0), 0.0, 0.0)
d_kind(::Bazinga.FreeSet) = (Int32(1), 0.0, 0.0)
d_kind(D::Bazinga.IndicatorSet{<:ProximalOperators.IndBox}) = (Int32(2), Float64(D.f.lb), Float64(D.f.ub))
d_kind(D) = error("BazingaB200: D = $(typeof(D)) is not in the GPU oracle catalogue (no CPU fallback)")
# demo structs register themselves the same way, e.g. for demo/portfolio.jl (f, c and D are user structs there):
#   BazingaB200.f_kind(f::SmoothCostP) = (Int32(2), (A = Matrix{Float64}(f.Q), b = zeros(size(f.Q, 1))))   # :33-44
#   BazingaB200.c_kind(c::ConstraintP) = (Int32(1), (A = Matrix{Float64}([c.mu'; ones(length(c.mu))']), b = zeros(2)))  # :46-58
#   BazingaB200.d_kind(D::SetP)        = (Int32(7), Float64(D.rho), 1.0)                                   # :60-69
# demo/obstacle.jl: SmoothCostObstacleL1 -> f kind 6, SmoothCostObstacleL2 -> 7; NonsmoothCostObstacleL1 -> g kind 9,
#   NonsmoothCostObstacleRedL1 -> 12, NonsmoothCostObstacleL2 -> 13, NonsmoothCostObstacleL2Red -> 8 (max(0, x));
#   ConstraintObstacle -> c kind 5, ConstraintObstacleRed -> 6; SetObstacle -> d kind 0, SetObstacleRed -> 6 with pair stride N
# constraint kinds are user structs in the demos/tests; register yours with a method of `c_kind`:
#   BazingaB200.c_kind(c::ConstraintBasisPursuit) = (Int32(1), (A = c.A, b = c.b))      # demo/basispursuit.jl:38
#   BazingaB200.c_kind(::IdentityFunction)        = (Int32(0), NamedTuple())            # test/definitions
c_kind(c) = error("BazingaB200: c = $(typeof(c)) is not registered; add a BazingaB200.c_kind method")

# ---- alps(f, g, c, D, x0, y0; ...) -- src/algorithms/alps.jl:7-26 --------------------------------------------
# ---- als(f, g, c, D, x0, y0; ...)  -- src/algorithms/als.jl:7-26 (same keywords, same 10-tuple) ----------------
alps(f, g, c::Bazinga.SmoothFunction, D::Bazinga.ClosedSet, x0::AbstractArray, y0::AbstractArray; kwargs...) =
    solve_on_device(false, f, g, c, D, x0, y0; kwargs...)
als(f, g, c::Bazinga.SmoothFunction, D::Bazinga.ClosedSet, x0::AbstractArray, y0::AbstractArray; kwargs...) =
    solve_on_device(true, f, g, c, D, x0, y0; kwargs...)

function solve_on_device(slack::Bool, f, g, c, D, x0, y0;
              tol::Real = 1e-6, tol_prim::Real = tol, tol_dual::Real = tol, inner_tol::Real = cbrt(tol_dual),
              maxit::Int = 100, theta_penalty::Real = 0.8, kappa_penalty::Real = 0.5, kappa_tol::Real = 0.1,
              verbose::Bool = false, dual_safeguard::Union{Nothing,Function} = nothing,
              subsolver = B200PANOCplus, subsolver_maxit::Int = 1_000_000_000)
    ctx = context()
    fk, fp = f_kind(f); gk, glam, glo, ghi, gu = g_kind(g); ck, cp = c_kind(c); dk, dlo, dhi = d_kind(D)
    nx, ny = length(x0), length(y0)
    sub = subsolver(tol = 1.0, verbose = verbose)
    keep = Any[fp, cp, gu]                               # GC roots for every pointer handed over
    ptr(a) = a === nothing ? Ptr{Float64}(C_NULL) : pointer(a)
    desc = ProblemDesc(fk, gk, ck, dk, nx, ny,
                       ptr(get(fp, :A, nothing)), ptr(get(fp, :b, nothing)),
                       haskey(fp, :A) ? size(fp.A, 1) : 0, Int32(1) #= column-major =#, (0.0, 0.0, 0.0, 0.0),
                       glam, glo, ghi, ptr(gu),
                       ptr(get(cp, :A, nothing)), ptr(get(cp, :b, nothing)), Int32(1) #= column-major =#,
                       Int32(0), Int32(0), UInt64(0), 0, 0, dlo, dhi, 1, 0, 0,
                       Int32(0), Int32(0), 0,   # no slab decomposition
                       Int32(slack))            # 0: alps / bz_alps_solve, 1: als / bz_als_solve (PANOC+ on [x; s])
    prob = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin   # (bz_problem_create copies the oracle data to the device before it returns)
        check(ccall((:bz_problem_create, lib), Cint, (Ptr{Cvoid}, Ref{ProblemDesc}, Ref{Ptr{Cvoid}}),
                    ctx.handle, Ref(desc), prob), ctx.handle)
    end
    opt = Options()
    ccall((:bz_default_options, lib), Cvoid, (Ref{Options},), opt)
    opt.tol = tol; opt.tol_prim = tol_prim; opt.tol_dual = tol_dual; opt.inner_tol = inner_tol
    opt.maxit = maxit; opt.theta_penalty = theta_penalty; opt.kappa_penalty = kappa_penalty
    opt.kappa_tol = kappa_tol; opt.subsolver_maxit = subsolver_maxit
    opt.alpha = sub.alpha; opt.beta = sub.beta; opt.minimum_gamma = sub.minimum_gamma
    opt.max_backtracks = sub.max_backtracks; opt.lbfgs_memory = sub.memory; opt.sub_maxit = sub.maxit
    opt.max_inner_total = 0; opt.verbose = verbose
    # dual_safeguard(y, cx) (alps.jl:23,62): the default clamp runs on the device; a user function becomes a host
    # callback on unsafe_wrap'ed copies of y and cx, once per outer iteration
    opt.dual_safeguard = C_NULL; opt.dual_safeguard_user = C_NULL
    if dual_safeguard !== nothing
        tramp = (yp::Ptr{Float64}, cp_::Ptr{Float64}, m::Int64, ::Ptr{Cvoid}) ->
            (dual_safeguard(unsafe_wrap(Array, yp, m), unsafe_wrap(Array, cp_, m)); nothing)
        cfn = @cfunction($tramp, Cvoid, (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid}))
        push!(keep, cfn)
        opt.dual_safeguard = Base.unsafe_convert(Ptr{Cvoid}, cfn)
    end
    T = Float64
    x = Vector{T}(undef, nx); y = Vector{T}(undef, ny); s = similar(y); mu = similar(y)
    res = Result()
    x0c = Vector{T}(x0); y0c = Vector{T}(y0)             # x0, y0 are never modified (test_nonconvex_qp.jl:36)
    GC.@preserve keep try
        if slack
            check(ccall((:bz_als_solve, lib), Cint,
                        (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ref{Options}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ref{Result}),
                        prob[], x0c, y0c, opt, x, y, s, mu, res), ctx.handle)
        else
            check(ccall((:bz_alps_solve, lib), Cint,
                        (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ref{Options}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ref{Result}),
                        prob[], x0c, y0c, opt, x, y, s, mu, res), ctx.handle)
        end
    finally
        ccall((:bz_problem_destroy, lib), Cint, (Ptr{Cvoid},), prob[])
    end
    if verbose                                           # alps.jl:52-55
        @info "initial penalty parameters μ ∈ [$(res.mu0_min), $(res.mu0_max)]"
        @info "initial inner tolerance $(res.inner_tol0)"
    end
    return x, y, Int(res.tot_it), Int(res.tot_inner_it), res.elapsed_time, STATUS[res.status + 1],
           res.inner_tol, res.norm_res_prim, s, mu
end

# ---- the demos' `for i = 1:ntests ... solver(x0)` loops as one launch ----------------------------------------------
# (demo/rosenbrock.jl:154-177, demo/eitheror.jl:238-271, demo/mpvca.jl:241-281).  `problem` is the demo's problem_name;
# X0 is nx x ntests (one starting point per column, Julia's layout = instance-major for the C ABI).
const BATCH_KINDS = Dict("rosenbrock" => (0, 2, 2), "eitheror" => (1, 2, 4), "xor" => (2, 2, 4), "mpvca" => (3, 2, 4),
                         "eitheror_fullslack" => (4, 6, 4), "xor_fullslack" => (5, 6, 4), "mpvca_slack" => (6, 4, 2),
                         "mpvca_fullslack" => (7, 6, 4))
function solve_batch(problem::String, X0::Matrix{Float64}; solver::Symbol = :alps, tol = 1e-8, inner_tol = 1.0, maxit = 100,
                     subsolver = B200PANOCplus(minimum_gamma = eps(Float64), maxit = 1000), subsolver_maxit = 1000)
    kind, nx, ny = BATCH_KINDS[problem]
    size(X0, 1) == nx || error("X0 must be $(nx) x ntests for problem $(problem)")
    ctx = context(); B = size(X0, 2)
    opt = Options()
    ccall((:bz_default_options, lib), Cvoid, (Ref{Options},), opt)
    opt.tol = tol; opt.inner_tol = inner_tol; opt.maxit = maxit; opt.subsolver_maxit = subsolver_maxit
    opt.alpha = subsolver.alpha; opt.beta = subsolver.beta; opt.minimum_gamma = subsolver.minimum_gamma
    opt.max_backtracks = subsolver.max_backtracks; opt.lbfgs_memory = subsolver.memory; opt.sub_maxit = subsolver.maxit
    Y0 = zeros(ny, B); X = similar(X0); Y = similar(Y0)
    its = Vector{Int32}(undef, B); inner = Vector{Int64}(undef, B); status = Vector{Int32}(undef, B)
    nrp = Vector{Float64}(undef, B); ms = Ref{Float64}(0.0)
    opt.dual_safeguard = C_NULL; opt.dual_safeguard_user = C_NULL
    # ccall needs a constant (name, library) pair: branch explicitly instead of choosing the symbol at run time
    rc = if solver == :als
        ccall((:bz_batch_solve_als, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Options}, Ptr{Float64}, Ptr{Float64},
               Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ref{Float64}),
              ctx.handle, kind, B, X0, Y0, opt, X, Y, its, inner, status, nrp, ms)
    else
        ccall((:bz_batch_solve, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Options}, Ptr{Float64}, Ptr{Float64},
               Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ref{Float64}),
              ctx.handle, kind, B, X0, Y0, opt, X, Y, its, inner, status, nrp, ms)
    end
    check(rc, ctx.handle)
    # the columns the demos write to CSV: xfinal, iters, sub_iters, is_solved (demo/eitheror.jl:256-271)
    return (x = X, y = Y, iters = its, sub_iters = inner, is_solved = status .== 0, norm_res_prim = nrp, kernel_ms = ms[])
end

# ---- matrix prox operators (src/proxoperators/rank.jl, nuclearNorm.jl, schattenNormLp.jl) ---------------------------
# `prox!(Y, f, X, gamma)` for f::Bazinga.Rank / NuclearNorm / SchattenNormLpPower on the device (bz_prox_matrix: one-sided
# Jacobi instead of LAPACK's svd).  A Julia Matrix is column-major, which is the layout the entry point takes.
matprox_params(f::Bazinga.Rank) = (Cint(0), Float64(f.lambda), 0.5)
matprox_params(f::Bazinga.NuclearNorm) = (Cint(1), Float64(f.lambda), 0.5)
matprox_params(f::Bazinga.SchattenNormLpPower) = (Cint(2), Float64(f.alpha), Float64(f.p))
function prox_matrix!(Y::Matrix{Float64}, f, X::Matrix{Float64}, gamma::Real)
    kind, lambda, p = matprox_params(f)
    ctx = context(); gz = Ref{Float64}(0.0)
    check(ccall((:bz_prox_matrix, lib), Cint,
                (Ptr{Cvoid}, Cint, Float64, Float64, Float64, Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Float64},
                 Ptr{Float64}, Ptr{Int32}),
                ctx.handle, kind, lambda, p, Float64(gamma), size(X, 1), size(X, 2), 1, X, Y, gz, C_NULL, C_NULL), ctx.handle)
    return gz[]
end

end # module
