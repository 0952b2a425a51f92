# CTDirectB200.jl -- Julia glue for libctdirect_b200.so.
#
# STATUS: Julia is not installed in this repository's build image, so this file has never been executed there.
# Its call sequence into the C ABI (create with the NORMALIZED grid -> dims -> info -> bounds -> x0 -> structures ->
# evaluations, for every scheme name `scheme_name` can produce) is replayed by a C program compiled against
# include/ctdirect_b200.h (tests/host_harness/julia_shim_sequence.c, run on the GPU by tests/test_julia_shim_sequence.py).
#
# It provides
#   * B200DOCPModel <: NLPModels.AbstractNLPModel{Float64, Vector{Float64}}: host callbacks (Ipopt, MadNLP on CPU),
#   * B200DOCPModelGPU <: NLPModels.AbstractNLPModel{Float64, CuVector{Float64}} (when CUDA.jl is loaded): device
#     callbacks through the `_dev` entry points for MadNLPGPU + cuDSS,
#   * a builder with the signature CTDirect stores in CTSolvers.ADNLPModelBuilder (src/collocation.jl:90-94,246),
#   * the legacy spellings `direct_transcription`, `nlp_model`, `solve` (test/tmp/percival.jl:24,
#     test/tmp/test_collocation.jl:11-41, test/tmp/ocp.jl:28-33), so that
#         docp = CTDirect.discretize(ocp, CTDirect.Collocation(grid_size=100_000, scheme=:midpoint))
#         nlp  = CTDirectB200.b200_model_builder("goddard", docp)(init; backend=:b200)
#         stats = NLPModelsIpopt.ipopt(nlp)            # or MadNLP.madnlp(nlp)
#     runs the solver on GPU-evaluated callbacks.
# The OCP itself must be one of the problems compiled into the library (the generator works from the OCP
# expression: python -m ctdirect_b200.gen.generate).
module CTDirectB200

using NLPModels

const LIB = get(ENV, "CTDIRECT_B200_LIB", "libctdirect_b200.so")

const CTD_CONS, CTD_JAC, CTD_HESS, CTD_GRAD, CTD_OBJ = Cint(1), Cint(2), Cint(4), Cint(8), Cint(16)

struct CtdDesc
    problem::Cstring
    scheme::Cstring
    grid_size::Int32
    time_grid::Ptr{Float64}
    n_time_grid::Int32
    params::Ptr{Float64}
    n_params::Int32
    pattern_mode::Int32
    batch::Int32
    device::Int32
    control_steps::Int32     # > 1: the DirectShooting discretizer (src/direct_shooting.jl:55-71), midpoint only
end

last_error() = unsafe_string(ccall((:ctd_last_error, LIB), Cstring, ()))
check(code) = code == 0 ? nothing : error("ctdirect_b200 [$code]: " * last_error())

"""
    scheme_name(discretization) -> String

The library's scheme name of a CTDirect discretization OBJECT.  The reference maps scheme SYMBOLS to struct types
(src/DOCP_data.jl:307-336); the struct names do not spell the symbols:
`Gauss_Legendre_N` is `:gauss_legendre_N_constant_control` (src/ode/irk.jl:58,96), `Gauss_Legendre_N_Stagewise` is
`:gauss_legendre_N` (src/ode/irk_stagewise.jl:37,79), and `Euler` carries its variant in `_explicit` (src/ode/euler.jl:10-16).
"""
function scheme_name(disc)
    T = string(nameof(typeof(disc)))
    T == "Trapeze" && return "trapeze"
    T == "Midpoint" && return "midpoint"
    T == "Euler" && return getfield(disc, :_explicit) ? "euler" : "euler_implicit"
    T == "Gauss_Legendre_1" && return "gauss_legendre_1_constant_control"
    T == "Gauss_Legendre_2" && return "gauss_legendre_2_constant_control"
    T == "Gauss_Legendre_3" && return "gauss_legendre_3_constant_control"
    T == "Gauss_Legendre_2_Stagewise" && return "gauss_legendre_2"
    T == "Gauss_Legendre_3_Stagewise" && return "gauss_legendre_3"
    error("CTDirectB200: no compiled kernel for discretization type $T")
end

"`:manual` -> 0 (reference block patterns), `:corrected` -> 1, `:optimized` -> 2 (the detected pattern)"
pattern_mode(backend::Symbol) = backend == :corrected ? 1 : (backend == :optimized ? 2 : 0)

# ---- the handle and what both model flavours share ------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    neval_time::Dict{Symbol,Float64}     # seconds spent in each callback (SURVEY.md section 5, "metrics")
end

function create_handle(problem::AbstractString, scheme::AbstractString; grid_size::Integer, normalized_grid=nothing,
                       params=nothing, pattern_mode::Integer=0, device::Integer=0, control_steps::Integer=1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    tg = normalized_grid === nothing ? Float64[] : Vector{Float64}(normalized_grid)
    par = params === nothing ? Float64[] : Vector{Float64}(params)
    GC.@preserve problem scheme tg par begin
        desc = CtdDesc(Base.unsafe_convert(Cstring, problem), Base.unsafe_convert(Cstring, scheme), Int32(grid_size),
                       isempty(tg) ? C_NULL : pointer(tg), Int32(length(tg)),
                       isempty(par) ? C_NULL : pointer(par), Int32(length(par)), Int32(pattern_mode), Int32(1), Int32(device),
                       Int32(control_steps))
        check(ccall((:ctd_create, LIB), Cint, (Ref{CtdDesc}, Ref{Ptr{Cvoid}}), desc, h))
    end
    hd = Handle(h[], Dict{Symbol,Float64}())
    finalizer(x -> ccall((:ctd_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), hd)
    return hd
end

function handle_meta(hd::Handle, name::AbstractString; x0=nothing)
    nvar, ncon, nnzj, nnzh = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:ctd_dims, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), hd.ptr, nvar, ncon, nnzj, nnzh))
    info = zeros(Int64, 16)
    check(ccall((:ctd_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), hd.ptr, info))
    lvar, uvar = zeros(nvar[]), zeros(nvar[])
    lcon, ucon = zeros(ncon[]), zeros(ncon[])
    check(ccall((:ctd_bounds, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), hd.ptr, lvar, uvar, lcon, ucon))
    if x0 === nothing
        x0 = zeros(nvar[])
        check(ccall((:ctd_x0, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), hd.ptr, x0))   # __initial_guess, DOCP_variables.jl:122-145
    end
    # info[14] (1-based) = the problem maximises: minimize = !docp.flags.max, collocation.jl:145
    return NLPModelMeta(nvar[]; ncon=ncon[], x0=Vector{Float64}(x0), lvar=lvar, uvar=uvar, lcon=lcon, ucon=ucon,
                        nnzj=nnzj[], nnzh=nnzh[], minimize=(info[14] == 0), name=name)
end

macro timed_call(hd, key, ex)
    quote
        local t0 = time_ns()
        local r = $(esc(ex))
        $(esc(hd)).neval_time[$(esc(key))] = get($(esc(hd)).neval_time, $(esc(key)), 0.0) + (time_ns() - t0) * 1e-9
        r
    end
end

"Evaluation counters (NLPModels `Counters`) and seconds per callback, as a Dict ready for JSON."
function eval_report(nlp)
    c = nlp.counters
    return Dict("neval_obj" => c.neval_obj, "neval_grad" => c.neval_grad, "neval_cons" => c.neval_cons,
                "neval_jac" => c.neval_jac, "neval_hess" => c.neval_hess,
                "gpu_launches" => ccall((:ctd_launch_count, LIB), Int64, (Ptr{Cvoid},), nlp.handle.ptr),
                "seconds" => Dict(string(k) => v for (k, v) in nlp.handle.neval_time))
end

function structure!(sym::Symbol, hd::Handle, n::Integer, rows, cols)
    r, c = Vector{Int64}(undef, n), Vector{Int64}(undef, n)
    if sym == :jac
        check(ccall((:ctd_jac_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), hd.ptr, r, c))
    else
        check(ccall((:ctd_hess_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), hd.ptr, r, c))
    end
    copyto!(rows, r); copyto!(cols, c)    # (works for host and device index vectors)
    return rows, cols
end

# ---- host model: Vector{Float64} callbacks (Ipopt, MadNLP on the CPU) ------------------------------------------------
mutable struct B200DOCPModel <: AbstractNLPModel{Float64,Vector{Float64}}
    meta::NLPModelMeta{Float64,Vector{Float64}}
    counters::Counters
    handle::Handle
end

"""
    B200DOCPModel(problem, scheme; grid_size, normalized_grid=nothing, params=nothing, x0=nothing,
                  pattern_mode=0, device=0)

Replaces `ADNLPModels.ADNLPModel!(f, x0, lvar, uvar, c!, lcon, ucon; ...)` of src/collocation.jl:137-149.
`scheme` is a library scheme name (`scheme_name`) or the reference's symbol (`:midpoint`, `:gauss_legendre_3`, ...).
"""
function B200DOCPModel(problem::AbstractString, scheme; grid_size::Integer=250, normalized_grid=nothing,
                       params=nothing, x0=nothing, pattern_mode::Integer=0, device::Integer=0, control_steps::Integer=1)
    sch = String(scheme)
    hd = create_handle(problem, sch; grid_size=grid_size, normalized_grid=normalized_grid, params=params,
                       pattern_mode=pattern_mode, device=device, control_steps=control_steps)
    return B200DOCPModel(handle_meta(hd, "CTDirectB200:$problem:$sch"; x0=x0), Counters(), hd)
end

function NLPModels.obj(nlp::B200DOCPModel, x::AbstractVector)
    increment!(nlp, :neval_obj)
    f = Ref{Float64}(0.0)
    @timed_call nlp.handle :obj check(ccall((:ctd_obj, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), nlp.handle.ptr, x, f))
    return f[]
end

function NLPModels.grad!(nlp::B200DOCPModel, x::AbstractVector, g::AbstractVector)
    increment!(nlp, :neval_grad)
    @timed_call nlp.handle :grad check(ccall((:ctd_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle.ptr, x, g))
    return g
end

function NLPModels.cons!(nlp::B200DOCPModel, x::AbstractVector, c::AbstractVector)
    increment!(nlp, :neval_cons)
    @timed_call nlp.handle :cons check(ccall((:ctd_cons, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle.ptr, x, c))
    return c
end

NLPModels.jac_structure!(nlp::B200DOCPModel, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer}) =
    structure!(:jac, nlp.handle, nlp.meta.nnzj, rows, cols)
NLPModels.hess_structure!(nlp::B200DOCPModel, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer}) =
    structure!(:hess, nlp.handle, nlp.meta.nnzh, rows, cols)

function NLPModels.jac_coord!(nlp::B200DOCPModel, x::AbstractVector, vals::AbstractVector)
    increment!(nlp, :neval_jac)
    @timed_call nlp.handle :jac check(ccall((:ctd_jac_coord, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle.ptr, x, vals))
    return vals
end

function NLPModels.hess_coord!(nlp::B200DOCPModel, x::AbstractVector, y::AbstractVector, vals::AbstractVector; obj_weight::Real=1.0)
    increment!(nlp, :neval_hess)
    @timed_call nlp.handle :hess check(ccall((:ctd_hess_coord, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}),
                                             nlp.handle.ptr, x, y, Float64(obj_weight), vals))
    return vals
end

# Hessian of the objective only (NLPModels calls this overload for unconstrained-style queries)
NLPModels.hess_coord!(nlp::B200DOCPModel, x::AbstractVector, vals::AbstractVector; obj_weight::Real=1.0) =
    hess_coord!(nlp, x, zeros(nlp.meta.ncon), vals; obj_weight=obj_weight)

"One fused launch for everything a solver iteration needs: fills g, c, jvals, hvals and returns f (ctd_eval)."
function eval_all!(nlp::B200DOCPModel, x, y, g, c, jvals, hvals; obj_weight::Real=1.0)
    f = Ref{Float64}(0.0)
    @timed_call nlp.handle :eval_all check(ccall((:ctd_eval, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Cdouble, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        nlp.handle.ptr, CTD_CONS | CTD_JAC | CTD_HESS | CTD_GRAD | CTD_OBJ, x, y, Float64(obj_weight), f, g, c, jvals, hvals))
    return f[]
end

# ---- device model: CuVector{Float64} callbacks through the `_dev` entry points (MadNLPGPU + cuDSS) --------------------
# Defined generically over a device vector type V whose `pointer(v)` is a device pointer convertible to Ptr{Cvoid}
# (CUDA.jl: `CuVector{Float64}`); instantiate with `B200DOCPModelGPU(CuVector{Float64}, problem, scheme; ...)`.
mutable struct B200DOCPModelGPU{V<:AbstractVector{Float64}} <: AbstractNLPModel{Float64,V}
    meta::NLPModelMeta{Float64,V}
    counters::Counters
    handle::Handle
    f::V          # one-element device buffer for the objective
    y0::V         # zero multipliers for the objective-only Hessian overload
end

devptr(v) = reinterpret(Ptr{Cvoid}, pointer(v))

function B200DOCPModelGPU(::Type{V}, problem::AbstractString, scheme; grid_size::Integer=250, normalized_grid=nothing,
                          params=nothing, x0=nothing, pattern_mode::Integer=0, device::Integer=0,
                          control_steps::Integer=1) where {V}
    sch = String(scheme)
    hd = create_handle(problem, sch; grid_size=grid_size, normalized_grid=normalized_grid, params=params,
                       pattern_mode=pattern_mode, device=device, control_steps=control_steps)
    mh = handle_meta(hd, "CTDirectB200:$problem:$sch:gpu"; x0=x0)
    meta = NLPModelMeta{Float64,V}(mh.nvar; ncon=mh.ncon, x0=V(mh.x0), lvar=V(mh.lvar), uvar=V(mh.uvar), lcon=V(mh.lcon),
                                   ucon=V(mh.ucon), nnzj=mh.nnzj, nnzh=mh.nnzh, minimize=mh.minimize, name=mh.name)
    return B200DOCPModelGPU{V}(meta, Counters(), hd, V(zeros(1)), V(zeros(mh.ncon)))
end

"Run the handle's work on the stream the caller's device arrays live on (`CUDA.stream().handle`)."
set_stream!(nlp::B200DOCPModelGPU, stream::Ptr{Cvoid}) =
    check(ccall((:ctd_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), nlp.handle.ptr, stream))

function eval_dev(nlp::B200DOCPModelGPU, what::Cint, x, y, sigma, f, g, c, jvals, hvals)
    check(ccall((:ctd_eval_dev, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                nlp.handle.ptr, what, x, y, Float64(sigma), f, g, c, jvals, hvals))
end

function NLPModels.obj(nlp::B200DOCPModelGPU, x::AbstractVector)
    increment!(nlp, :neval_obj)
    @timed_call nlp.handle :obj eval_dev(nlp, CTD_OBJ, devptr(x), C_NULL, 1.0, devptr(nlp.f), C_NULL, C_NULL, C_NULL, C_NULL)
    check(ccall((:ctd_sync, LIB), Cint, (Ptr{Cvoid},), nlp.handle.ptr))
    return Array(nlp.f)[1]
end

function NLPModels.grad!(nlp::B200DOCPModelGPU, x::AbstractVector, g::AbstractVector)
    increment!(nlp, :neval_grad)
    @timed_call nlp.handle :grad eval_dev(nlp, CTD_GRAD | CTD_OBJ, devptr(x), C_NULL, 1.0, devptr(nlp.f), devptr(g), C_NULL, C_NULL, C_NULL)
    return g
end

function NLPModels.cons!(nlp::B200DOCPModelGPU, x::AbstractVector, c::AbstractVector)
    increment!(nlp, :neval_cons)
    @timed_call nlp.handle :cons eval_dev(nlp, CTD_CONS, devptr(x), C_NULL, 1.0, C_NULL, C_NULL, devptr(c), C_NULL, C_NULL)
    return c
end

NLPModels.jac_structure!(nlp::B200DOCPModelGPU, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer}) =
    structure!(:jac, nlp.handle, nlp.meta.nnzj, rows, cols)
NLPModels.hess_structure!(nlp::B200DOCPModelGPU, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer}) =
    structure!(:hess, nlp.handle, nlp.meta.nnzh, rows, cols)

function NLPModels.jac_coord!(nlp::B200DOCPModelGPU, x::AbstractVector, vals::AbstractVector)
    increment!(nlp, :neval_jac)
    @timed_call nlp.handle :jac eval_dev(nlp, CTD_JAC, devptr(x), C_NULL, 1.0, C_NULL, C_NULL, C_NULL, devptr(vals), C_NULL)
    return vals
end

function NLPModels.hess_coord!(nlp::B200DOCPModelGPU, x::AbstractVector, y::AbstractVector, vals::AbstractVector; obj_weight::Real=1.0)
    increment!(nlp, :neval_hess)
    @timed_call nlp.handle :hess eval_dev(nlp, CTD_HESS, devptr(x), devptr(y), obj_weight, C_NULL, C_NULL, C_NULL, C_NULL, devptr(vals))
    return vals
end

NLPModels.hess_coord!(nlp::B200DOCPModelGPU, x::AbstractVector, vals::AbstractVector; obj_weight::Real=1.0) =
    hess_coord!(nlp, x, nlp.y0, vals; obj_weight=obj_weight)

# ---- the seam in CTDirect: builder, and the legacy spellings ----------------------------------------------------------
"""
    b200_model_builder(problem, docp; device_vector=nothing, params=nothing)

Builder with the signature of `build_adnlp_model(initial_guess; backend, kwargs...)` (src/collocation.jl:90-94): pass it
to `CTSolvers.ADNLPModelBuilder` in place of the ADNLPModels one.  `problem` names the compiled-in OCP.
`backend`: `:manual` / `:b200` (reference block patterns), `:corrected`, `:optimized`.  With `device_vector = CuVector{Float64}`
the device model is returned.

The time grid handed to the library is `docp.time.normalized_grid`: for free-time problems the reference leaves
`docp.time.fixed_grid` UNINITIALISED (src/DOCP_data.jl:203-205), and for fixed times re-normalising `t0 + s (tf - t0)`
does not give `s` back bit for bit.  The dimensions the library derives are checked against the DOCP's.
"""
function b200_model_builder(problem::AbstractString, docp; device_vector=nothing, params=nothing)
    return function build_b200_model(initial_guess=nothing; backend=:b200, kwargs...)
        sch = scheme_name(docp.discretization)
        # DirectShooting (src/direct_shooting.jl:55-71): docp.time.control_steps controls per step, midpoint only; the
        # library then uses the detected pattern whatever `backend` says (there is no :manual pattern for it)
        cs = docp.time.control_steps
        (cs == 1 || sch == "midpoint") || error("CTDirectB200: control_steps > 1 is compiled for the midpoint scheme only")
        x0 = initial_guess isa AbstractVector ? initial_guess : nothing     # an NLP vector; else the library's __initial_guess
        args = (problem, sch)
        kw = (grid_size=docp.time.steps, normalized_grid=docp.time.normalized_grid, params=params, x0=x0,
              pattern_mode=pattern_mode(backend), control_steps=cs)
        nlp = device_vector === nothing ? B200DOCPModel(args...; kw...) : B200DOCPModelGPU(device_vector, args...; kw...)
        (nlp.meta.nvar == docp.dim_NLP_variables && nlp.meta.ncon == docp.dim_NLP_constraints) ||
            error("CTDirectB200: the library's model ($(nlp.meta.nvar), $(nlp.meta.ncon)) does not match the DOCP " *
                  "($(docp.dim_NLP_variables), $(docp.dim_NLP_constraints)): is \"$problem\" the OCP that was discretized?")
        return nlp
    end
end

"`nlp_model(docp)` of the older API: the model of a (problem, DOCP) pair."
nlp_model(problem::AbstractString, docp; kwargs...) = b200_model_builder(problem, docp; kwargs...)()

"""
    direct_transcription(problem; grid_size=250, disc_method=:midpoint, time_grid=nothing, init=nothing, ...) -> nlp

Legacy spelling (`docp, nlp = direct_transcription(ocp; grid_size=25)`, test/tmp/percival.jl:24) for a compiled-in problem
given by NAME: no CTDirect DOCP object is needed, the library holds the discretised problem.  `time_grid` may be any
strictly increasing grid; it is normalised like DOCPtime does (src/DOCP_data.jl:186-196).
"""
function direct_transcription(problem::AbstractString; grid_size::Integer=250, disc_method::Symbol=:midpoint, scheme::Symbol=disc_method,
                              time_grid=nothing, init=nothing, params=nothing, backend::Symbol=:manual, device_vector=nothing,
                              control_steps::Integer=1)
    kw = (grid_size=time_grid === nothing ? grid_size : length(time_grid) - 1, normalized_grid=time_grid, params=params,
          x0=init, pattern_mode=pattern_mode(backend), control_steps=control_steps)
    return device_vector === nothing ? B200DOCPModel(problem, scheme; kw...) : B200DOCPModelGPU(device_vector, problem, scheme; kw...)
end

"""
    solve(problem, solver; kwargs...) -> solver statistics

Legacy `solve(ocp, :adnlp, :ipopt)` (test/tmp/ocp.jl:28-33): transcribe and hand the model to `solver`, any callable
taking an NLPModel (`NLPModelsIpopt.ipopt`, `MadNLP.madnlp`).  Extra keyword arguments go to the transcription.
"""
solve(problem::AbstractString, solver; solver_options=(;), kwargs...) = solver(direct_transcription(problem; kwargs...); solver_options...)

"Solution arrays of `build_OCP_solution` (src/DOCP_data.jl:514-633) from a solver's primal / dual vectors (host)."
function unpack_solution(nlp::B200DOCPModel, x::Vector{Float64}, y::Union{Nothing,Vector{Float64}}=nothing)
    info = zeros(Int64, 16)
    check(ccall((:ctd_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), nlp.handle.ptr, info))
    nx, nu, nv, np, nb, N = info[1:6]
    out = (time=zeros(N + 1), state=zeros(nx, N + 1), control=zeros(nu, N + 1), variable=zeros(nv),
           costate=zeros(nx, N), path_dual=zeros(np, N + 1), boundary_dual=zeros(nb))
    ptrs = Ptr{Float64}[isempty(a) ? C_NULL : pointer(a) for a in out]   # ctd_solution: 7 pointers, this order
    GC.@preserve out check(ccall((:ctd_unpack, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Float64}}),
                                 nlp.handle.ptr, x, y === nothing ? C_NULL : y, ptrs))
    return out
end

"""
CSC lower triangle of `[H + diag(dx)  .; J  -diag(dc)]` (what MadNLP's sparse KKT system hands to cuDSS):
`(colptr, rowval)`, 1-based, ready for `SparseMatrixCSC(n, n, colptr, rowval, nzval)`.
"""
function kkt_structure(nlp)
    n = nlp.meta.nvar + nlp.meta.ncon
    nnzk = ccall((:ctd_kkt_nnz, LIB), Int64, (Ptr{Cvoid},), nlp.handle.ptr)
    colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnzk)
    check(ccall((:ctd_kkt_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nlp.handle.ptr, colptr, rowval))
    return colptr, rowval
end

"Device pointers: move `hvals`, `jvals` (and the regularisations `dx`, `dc`, or C_NULL) to KKT order in `kvals`."
kkt_assemble!(nlp, hvals, jvals, dx, dc, kvals) =
    check(ccall((:ctd_kkt_assemble_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                nlp.handle.ptr, hvals, jvals, dx, dc, kvals))

"The library's own per-handle counters (`ctd_counters`): (neval_obj, neval_grad, neval_cons, neval_jac, neval_hess, calls, host_calls, norm_calls)."
function library_counters(nlp)
    out = zeros(Int64, 8)
    check(ccall((:ctd_counters, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), nlp.handle.ptr, out))
    return NamedTuple{(:neval_obj, :neval_grad, :neval_cons, :neval_jac, :neval_hess, :calls, :host_calls, :norm_calls)}(Tuple(out))
end

"Counters, sizes, kernel launches and host-call wall time of the handle as a JSON string (`ctd_stats_json`)."
function library_stats_json(nlp)
    n = ccall((:ctd_stats_json, LIB), Int64, (Ptr{Cvoid}, Ptr{UInt8}, Int64), nlp.handle.ptr, C_NULL, 0)
    buf = Vector{UInt8}(undef, n)
    ccall((:ctd_stats_json, LIB), Int64, (Ptr{Cvoid}, Ptr{UInt8}, Int64), nlp.handle.ptr, buf, n)
    return unsafe_string(pointer(buf))
end

# ---- batched scenarios on several GPUs: the library's own NCCL communicator (include/ctdirect_b200.h) ----
"128 bytes created by ONE rank; broadcast them to every rank (MPI, Distributed, a file) before `comm_init!`."
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:ctd_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end

comm_init!(handle::Ptr{Cvoid}, id::Vector{UInt8}, rank::Integer, world::Integer) =
    check(ccall((:ctd_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), handle, id, rank, world))

"All-gather of the 2*batch doubles at device pointer `localp` into `gatheredp` (world*2*batch), slot in 0:3."
gather_norms!(handle::Ptr{Cvoid}, localp, gatheredp, slot::Integer) =
    check(ccall((:ctd_gather_norms_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32), handle, localp, gatheredp, slot))

"Make the handle's stream wait until the gather that last used `slot` has landed."
comm_wait!(handle::Ptr{Cvoid}, slot::Integer) =
    check(ccall((:ctd_comm_wait, LIB), Cint, (Ptr{Cvoid}, Int32), handle, slot))

"Map every rank's gather buffer into every rank (after `comm_init!`): the gather fused into the evaluation kernel."
peer_init!(handle::Ptr{Cvoid}) = check(ccall((:ctd_peer_init, LIB), Cint, (Ptr{Cvoid},), handle))

end # module
