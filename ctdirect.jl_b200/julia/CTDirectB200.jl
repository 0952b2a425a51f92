# CTDirectB200.jl -- Julia glue for libctdirect_b200.so (UNEXECUTED in this repository's image: Julia is
# not installed there; the same entry points are exercised through ctypes by the Python mirror).
#
# It provides an NLPModels.AbstractNLPModel whose callbacks ccall the C ABI of include/ctdirect_b200.h,
# and a builder with the signature CTDirect stores in CTSolvers.ADNLPModelBuilder
# (src/collocation.jl:90-94,246), so that
#     docp = CTDirect.discretize(ocp, CTDirect.Collocation(grid_size=100_000, scheme=:midpoint))
#     nlp  = CTDirectB200.build_b200_model(docp_name, docp; backend=:b200)
#     stats = NLPModelsIpopt.ipopt(nlp)            # or MadNLP.madnlp(nlp)
# runs the solver on GPU-evaluated callbacks.  The OCP itself must be one of the problems compiled into the
# library (the generator works from the OCP expression: python -m ctdirect_b200.gen.generate).
module CTDirectB200

using NLPModels

const LIB = get(ENV, "CTDIRECT_B200_LIB", "libctdirect_b200.so")

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
end

last_error() = unsafe_string(ccall((:ctd_last_error, LIB), Cstring, ()))
check(code) = code == 0 ? nothing : error("ctdirect_b200 [$code]: " * last_error())

mutable struct B200DOCPModel <: AbstractNLPModel{Float64,Vector{Float64}}
    meta::NLPModelMeta{Float64,Vector{Float64}}
    counters::Counters
    handle::Ptr{Cvoid}
end

"""
    B200DOCPModel(problem, scheme; grid_size, time_grid=nothing, params=nothing, x0=nothing,
                  pattern_mode=0, device=0)

Replaces `ADNLPModels.ADNLPModel!(f, x0, lvar, uvar, c!, lcon, ucon; ...)` of src/collocation.jl:137-149.
"""
function B200DOCPModel(problem::AbstractString, scheme::Symbol; grid_size::Int=250, time_grid=nothing,
                       params=nothing, x0=nothing, pattern_mode::Int=0, device::Int=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    sch = String(scheme)
    tg = time_grid === nothing ? Float64[] : Vector{Float64}(time_grid)
    par = params === nothing ? Float64[] : Vector{Float64}(params)
    GC.@preserve problem sch tg par begin
        desc = CtdDesc(Base.unsafe_convert(Cstring, problem), Base.unsafe_convert(Cstring, sch), Int32(grid_size),
                       isempty(tg) ? C_NULL : pointer(tg), Int32(length(tg)),
                       isempty(par) ? C_NULL : pointer(par), Int32(length(par)), Int32(pattern_mode), Int32(1), Int32(device))
        check(ccall((:ctd_create, LIB), Cint, (Ref{CtdDesc}, Ref{Ptr{Cvoid}}), desc, h))
    end
    nvar, ncon, nnzj, nnzh = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:ctd_dims, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h[], nvar, ncon, nnzj, nnzh))
    info = zeros(Int64, 16)
    check(ccall((:ctd_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h[], info))
    lvar, uvar = zeros(nvar[]), zeros(nvar[])
    lcon, ucon = zeros(ncon[]), zeros(ncon[])
    check(ccall((:ctd_bounds, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h[], lvar, uvar, lcon, ucon))
    if x0 === nothing
        x0 = zeros(nvar[])
        check(ccall((:ctd_x0, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h[], x0))   # __initial_guess, DOCP_variables.jl:122-145
    end
    meta = NLPModelMeta(nvar[]; ncon=ncon[], x0=x0, lvar=lvar, uvar=uvar, lcon=lcon, ucon=ucon,
                        nnzj=nnzj[], nnzh=nnzh[], minimize=(info[14] == 0),      # minimize = !docp.flags.max, collocation.jl:145
                        name="CTDirectB200:$problem:$scheme")
    nlp = B200DOCPModel(meta, Counters(), h[])
    finalizer(m -> ccall((:ctd_destroy, LIB), Cint, (Ptr{Cvoid},), m.handle), nlp)
    return nlp
end

function NLPModels.obj(nlp::B200DOCPModel, x::AbstractVector)
    increment!(nlp, :neval_obj)
    f = Ref{Float64}(0.0)
    check(ccall((:ctd_obj, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), nlp.handle, x, f))
    return f[]
end

function NLPModels.grad!(nlp::B200DOCPModel, x::AbstractVector, g::AbstractVector)
    increment!(nlp, :neval_grad)
    check(ccall((:ctd_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, g))
    return g
end

function NLPModels.cons!(nlp::B200DOCPModel, x::AbstractVector, c::AbstractVector)
    increment!(nlp, :neval_cons)
    check(ccall((:ctd_cons, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, c))
    return c
end

function NLPModels.jac_structure!(nlp::B200DOCPModel, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer})
    r, c = Vector{Int64}(undef, nlp.meta.nnzj), Vector{Int64}(undef, nlp.meta.nnzj)
    check(ccall((:ctd_jac_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nlp.handle, r, c))
    rows .= r; cols .= c
    return rows, cols
end

function NLPModels.jac_coord!(nlp::B200DOCPModel, x::AbstractVector, vals::AbstractVector)
    increment!(nlp, :neval_jac)
    check(ccall((:ctd_jac_coord, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, vals))
    return vals
end

function NLPModels.hess_structure!(nlp::B200DOCPModel, rows::AbstractVector{<:Integer}, cols::AbstractVector{<:Integer})
    r, c = Vector{Int64}(undef, nlp.meta.nnzh), Vector{Int64}(undef, nlp.meta.nnzh)
    check(ccall((:ctd_hess_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nlp.handle, r, c))
    rows .= r; cols .= c
    return rows, cols
end

function NLPModels.hess_coord!(nlp::B200DOCPModel, x::AbstractVector, y::AbstractVector, vals::AbstractVector; obj_weight::Real=1.0)
    increment!(nlp, :neval_hess)
    check(ccall((:ctd_hess_coord, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}),
                nlp.handle, x, y, Float64(obj_weight), vals))
    return vals
end

# Hessian of the objective only (NLPModels calls this overload for unconstrained-style queries)
function NLPModels.hess_coord!(nlp::B200DOCPModel, x::AbstractVector, vals::AbstractVector; obj_weight::Real=1.0)
    return hess_coord!(nlp, x, zeros(nlp.meta.ncon), vals; obj_weight=obj_weight)
end

"""
Builder with the signature of `build_adnlp_model(initial_guess; backend, kwargs...)` (src/collocation.jl:90-94):
pass it to `CTSolvers.ADNLPModelBuilder` in place of the ADNLPModels one.  `problem` names the compiled-in OCP.
"""
function b200_model_builder(problem::AbstractString, docp)
    return function build_b200_model(initial_guess=nothing; backend=:b200, kwargs...)
        scheme = Symbol(lowercase(string(nameof(typeof(docp.discretization)))))  # :midpoint, :trapeze, ...
        x0 = initial_guess === nothing ? nothing : initial_guess
        return B200DOCPModel(problem, scheme; grid_size=docp.time.steps,
                             time_grid=docp.time.fixed_grid, x0=x0)
    end
end

"""
CSC lower triangle of `[H + diag(dx)  .; J  -diag(dc)]` (what MadNLP's sparse KKT system hands to cuDSS):
`(colptr, rowval)`, 1-based, ready for `SparseMatrixCSC(n, n, colptr, rowval, nzval)`.
"""
function kkt_structure(nlp::B200DOCPModel)
    n = nlp.meta.nvar + nlp.meta.ncon
    nnzk = ccall((:ctd_kkt_nnz, LIB), Int64, (Ptr{Cvoid},), nlp.handle)
    colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnzk)
    check(ccall((:ctd_kkt_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nlp.handle, colptr, rowval))
    return colptr, rowval
end

"Device pointers: move `hvals`, `jvals` (and the regularisations `dx`, `dc`, or C_NULL) to KKT order in `kvals`."
kkt_assemble!(handle::Ptr{Cvoid}, hvals, jvals, dx, dc, kvals) =
    check(ccall((:ctd_kkt_assemble_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                handle, hvals, jvals, dx, dc, kvals))

"Run the handle's work on a caller-owned CUDA stream (e.g. `CUDA.stream().handle`); `C_NULL` restores its own."
set_stream!(handle::Ptr{Cvoid}, stream::Ptr{Cvoid}) =
    check(ccall((:ctd_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle, stream))

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

end # module
