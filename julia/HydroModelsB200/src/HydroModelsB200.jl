# HydroModelsB200.jl — Julia host side of the B200 engine: `B200Model(model)(input, params, config; initstates)`.
#
# UNTESTED: there is no Julia in the build image (SURVEY.md F3).  The C-ABI this file binds
# (include/hydrob200.h) is exercised end to end by the Python binding (hydromodels.jl_b200/engine.py); this
# file mirrors that binding and a reduced version of hydromodels.jl_b200/lowering.py (hash-consed CSE across
# the pre-/post-state evaluations and parameter hoisting; the carried-value and symmetry rewrites are
# optimisations the Python lowering adds on top and do not change results beyond 1e-15).  Covered: HydroBucket,
# HydroFlux and UnitHydrograph components (GR4J end to end: powers lower like lowering._lower_pow, the unit
# hydrograph keeps its ordinates in hb_uhw[] and its history in the hb_uhh[] ring, ordinates evaluated on the
# host from uh.max_lag / uh.uh_func exactly as src/uh.jl:215-229 does).  Not covered here (Python host only):
# neural fluxes, HydroRoute, objectives / gradients, multi-GPU.
#
# What it replaces in the reference: HydroModelCore's build_bucket_func / build_flux_func code generation
# (call sites src/bucket.jl:58, src/flux.jl:52) and the functors of src/model.jl:241-309, src/bucket.jl:169-256.
module HydroModelsB200

using HydroModels, ComponentArrays, Symbolics, SymbolicUtils
import HydroModels: get_input_names, get_output_names, get_state_names, get_param_names, get_nn_names

const LIB = get(ENV, "HYDROB200_LIB", joinpath(@__DIR__, "..", "..", "..", "hydromodels.jl_b200", "libhydrob200.so"))

# ---- C structs (field order and types are the ABI; see include/hydrob200.h) -----------------------------
struct HbStepperSpec
    struct_size::Int32; n_inputs::Int32; n_rows::Int32; n_states::Int32; n_params::Int32; n_uh::Int32
    uh_kmax::NTuple{8,Int32}; nn_words::Int32; output_mode::Int32; metric::Int32; shared_forcing::Int32
    block_threads::Int32; stages::Int32; max_registers::Int32; precision::Int32; nn_tensor::Int32
    nodes_per_warp::Int32; n_tangents::Int32; n_aux_inputs::Int32; nn_tc_bytes::Int32
end

struct HbRunDesc
    struct_size::Int32; n_htype_sets::Int32; n_nodes::Int64; n_steps::Int64
    input::Ptr{Cvoid}; params::Ptr{Float64}; n_param_values::Int64
    param_offset::Ptr{Int32}; param_mode::Ptr{Int32}; htypes::Ptr{Int32}
    initstates::Ptr{Float64}; uh_weights::Ptr{Float64}; uh_hist_in::Ptr{Float64}; nn_params::Ptr{Float64}
    target::Ptr{Float64}; target_per_node::Int32; warm_up::Int32; min_value::Float64
    out::Ptr{Cvoid}; objective::Ptr{Float64}; final_states::Ptr{Float64}; uh_hist_out::Ptr{Float64}
    elapsed_ms::Ptr{Float32}; gradient::Ptr{Float64}; aux_input::Ptr{Cvoid}
    n_compact_rows::Int32; compact_rows::NTuple{4,Int32}
end

function hb_check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:hb_last_error, LIB), Cstring, ()))
    rc == 1 && throw(ArgumentError(msg))          # HB_ERR_INVALID
    rc == 5 && throw(OutOfMemoryError())
    error(msg)                                    # HB_ERR_NO_DEVICE / COMPILE / CUDA: there is no CPU fallback
end

# ---- lowering: Symbolics trees -> CUDA C body ---------------------------------------------------------------
mutable struct Builder
    lines_pro::Vector{String}      # prologue (parameter-only values)
    lines_step::Vector{String}
    memo::Dict{Any,String}         # structural key => C expression / variable name
    isparam::Dict{String,Bool}     # variable name => depends on parameters only
    counter::Int
end
Builder() = Builder(String[], String[], Dict{Any,String}(), Dict{String,Bool}(), 0)

cnum(x::Real) = (s = repr(Float64(x)); occursin(r"[.eE]|Inf|NaN", s) ? s : s * ".0")

function emit!(b::Builder, key, text::String, paramonly::Bool)
    haskey(b.memo, key) && return b.memo[key]
    b.counter += 1
    name = (paramonly ? "h" : "v") * string(b.counter)
    push!(paramonly ? b.lines_pro : b.lines_step, "const double $name = $text;")
    b.memo[key] = name
    b.isparam[name] = paramonly
    return name
end

# env: Symbol => C expression (hb_x[i], hb_p[j], hb_u[s], hb_un<s>, or an earlier value name)
function lower(b::Builder, ex, env::Dict{Symbol,String}, params::Set{Symbol})
    ex = Symbolics.unwrap(ex)
    if ex isa Real
        return cnum(ex)
    elseif SymbolicUtils.issym(ex)
        nm = Symbol(SymbolicUtils.nameof(ex))
        haskey(env, nm) || error("Variable $nm is not defined where it is used")
        return env[nm]
    end
    op = SymbolicUtils.operation(ex)
    args = [lower(b, a, env, params) for a in SymbolicUtils.arguments(ex)]
    ponly = all(a -> occursin(r"^[-0-9.]", a) || startswith(a, "hb_p[") || get(b.isparam, a, false), args)
    fold(f) = foldl((x, y) -> "($x $f $y)", args)
    is_zero(s) = s in ("0.0", "0", "0.0e0")      # max(0, x) lowers to the sign-bit test hb_relu like lowering._expr_with (Python host)
    text = if op === (+);  fold("+")
    elseif op === (*);     fold("*")
    elseif op === (-);     length(args) == 1 ? "(-$(args[1]))" : "($(args[1]) - $(args[2]))"
    elseif op === (/);     "($(args[1]) / $(args[2]))"
    elseif op === (^);     lower_pow(b, args[1], args[2], ponly)
    elseif op === exp;     "hb_exp($(args[1]))"
    elseif op === log;     "log($(args[1]))"
    elseif op === sqrt;    "sqrt($(args[1]))"
    elseif op === abs;     "fabs($(args[1]))"
    elseif op === tanh;    "hb_tanh($(args[1]))"       # step_func traces to 0.5*(1+tanh(5x)) (src/HydroModels.jl:204)
    elseif op === min;     "hb_min($(args[1]), $(args[2]))"
    elseif op === max;     is_zero(args[1]) ? "hb_relu($(args[2]))" : is_zero(args[2]) ? "hb_relu($(args[1]))" : "hb_max($(args[1]), $(args[2]))"
    elseif op === clamp;   "hb_min(" * (is_zero(args[2]) ? "hb_relu($(args[1]))" : "hb_max($(args[1]), $(args[2]))") * ", $(args[3]))"
    else error("Unknown function: $op")
    end
    return emit!(b, (op, args...), text, ponly)
end

# x^p like lowering._lower_pow: small integer powers by repeated multiplication, half-integer powers as x^n * sqrt(x),
# everything else through the branch-free hb_pow of hb_math.cuh (not libm pow: different last bits, and 3x slower)
function lower_pow(b::Builder, base::String, ex::String, ponly::Bool)
    p = tryparse(Float64, ex)
    mul(x, y) = emit!(b, (*, x, y), "($x * $y)", ponly)
    function ipow(x, n)
        result = nothing; sq = x
        while n > 0
            if n & 1 == 1; result = result === nothing ? sq : mul(result, sq); end
            n >>= 1
            n > 0 && (sq = mul(sq, sq))
        end
        result
    end
    if p !== nothing && p == round(p) && abs(p) <= 16
        n = Int(abs(p))
        n == 0 && return "1.0"
        r = ipow(base, n)
        return p > 0 ? r : "(1.0 / $r)"
    elseif p !== nothing && 2p == round(2p) && 0 < p <= 8.5
        sq = emit!(b, (sqrt, base), "sqrt($base)", ponly)
        n = Int(floor(p))
        return n == 0 ? sq : "($(ipow(base, n)) * $sq)"
    end
    return "hb_pow($base, $ex)"
end

struct Plan
    body::String
    uh_kmax::Vector{Int}; uh_components::Vector{Any}
    input_names::Vector{Symbol}; row_names::Vector{Symbol}; state_names::Vector{Symbol}; param_names::Vector{Symbol}
end

"""
    lower_stepper(model) -> Plan

One straight-line step for the whole model, components in declaration order (src/model.jl:255-268); per
stateful bucket: fluxes(pre-state) -> dS -> u <- max(min_value, u + dS) -> fluxes(post-state)
(src/solve.jl:47-52, src/bucket.jl:184-192).  `fluxes_of(bucket)` must return the bucket's
`(HydroFlux..., StateFlux...)` lists: HydroBucket drops them (src/bucket.jl:62-64), so either build buckets
through a capturing constructor or re-trace `bucket.ode_func` with `Num` inputs (SURVEY.md §8b).
"""
function lower_stepper(model::HydroModel; fluxes_of, uh_kmax::Vector{Int}=Int[])
    b = Builder()
    uhs = [c for c in model.components if c isa HydroModels.UnitHydrograph]
    length(uh_kmax) == length(uhs) || (uh_kmax = fill(1, length(uhs)))
    w_off = cumsum(vcat(0, uh_kmax)); h_off = cumsum(vcat(0, uh_kmax .- 1))
    uh_tail = String[]                                   # history shifts, emitted at the end of the step
    inputs = get_input_names(model); states = get_state_names(model); outputs = get_output_names(model)
    pnames = get_param_names(model)
    params = Set(pnames)
    env = Dict{Symbol,String}()
    for (i, nm) in enumerate(inputs);  env[nm] = "hb_x[$(i - 1)]"; end
    for (j, nm) in enumerate(pnames);  env[nm] = "hb_p[$(j - 1)]"; end
    for comp in model.components
        if comp isa HydroModels.UnitHydrograph
            # causal convolution with the host-evaluated, normalised ordinates (src/uh.jl:187-237): ring of past inputs
            u = findfirst(c -> c === comp, uhs); K = uh_kmax[u]
            x = env[get_input_names(comp)[1]]
            b.counter += 1
            y = "v$(b.counter)"
            push!(b.lines_step, "double $y = hb_uhw[$(w_off[u])] * $x;")
            for k in 1:K-1
                push!(b.lines_step, "$y = fma(hb_uhw[$(w_off[u] + k)], hb_uhh[$(h_off[u] + k - 1)], $y);")
            end
            for k in K-2:-1:1
                push!(uh_tail, "hb_uhh[$(h_off[u] + k)] = hb_uhh[$(h_off[u] + k - 1)];")
            end
            K >= 2 && push!(uh_tail, "hb_uhh[$(h_off[u])] = $x;")
            env[get_output_names(comp)[1]] = y
            continue
        end
        fluxes, dfluxes = fluxes_of(comp)
        cstates = get_state_names(comp)
        sidx = [findfirst(==(s), states) - 1 for s in cstates]
        run_fluxes!(local_env) = for f in fluxes, (lhs, rhs) in zip(get_output_names(f), f.exprs)
            local_env[lhs] = lower(b, rhs, local_env, params)
        end
        if isempty(cstates)
            local_env = copy(env); run_fluxes!(local_env); merge!(env, local_env)
            continue
        end
        pre = copy(env)
        for (s, si) in zip(cstates, sidx); pre[s] = "hb_u[$si]"; end
        run_fluxes!(pre)
        dus = [lower(b, first(d.exprs), pre, params) for d in dfluxes]
        for (si, du) in zip(sidx, dus)
            push!(b.lines_step, "const double hb_un$si = hb_max(hb_minv, hb_u[$si] + $du);")
        end
        post = copy(env)
        for (s, si) in zip(cstates, sidx); post[s] = "hb_un$si"; end
        run_fluxes!(post)
        merge!(env, post)
    end
    rows = vcat(states, outputs)
    for (r, nm) in enumerate(rows); push!(b.lines_step, "hb_out[$(r - 1)] = $(env[nm]);"); end
    for si in 0:length(states)-1; push!(b.lines_step, "hb_u[$si] = hb_un$si;"); end
    append!(b.lines_step, uh_tail)
    # bodies that raise to a power get the table logarithm of hb_math.cuh (2 KB more shared memory), as the Python host decides
    uses_pow = any(occursin("hb_pow(", l) for l in vcat(b.lines_pro, b.lines_step))
    body = "//@@HB:DEFINES\n" * (uses_pow ? "#define HB_LOG_TABLE 1\n" : "") * "#define HB_NC 0\n//@@HB:HELPERS\n//@@HB:PROLOGUE\n" *
           join(b.lines_pro, "\n") * "\n//@@HB:STEP\n" * join(b.lines_step, "\n") * "\n"
    return Plan(body, uh_kmax, uhs, inputs, rows, states, pnames)
end

# normalised ordinates of one unit hydrograph for one parameter set, zero-padded to kmax (K == 0: identity [1, 0, ...]);
# src/uh.jl:215-229: S-curve values uh_func(t, p) for t = 1:max_lag(p), differences, normalisation (hydro_conv :188)
function uh_ordinates(uh, p::ComponentVector, kmax::Int)
    K = Int(uh.max_lag(p))
    w = zeros(Float64, kmax)
    if K == 0
        w[1] = 1.0
        return w
    end
    s = [uh.uh_func(t, p) for t in 1:K]
    w[1:K] .= vcat(s[1], diff(s))
    w[1:K] ./= sum(w[1:K])
    return w
end

# ---- the wrapper component ------------------------------------------------------------------------------------
struct B200Model{M} <: HydroModels.AbstractModel
    model::M
    handle::Ptr{Cvoid}
    plan::Plan
    htypes::Vector{Int}
end
get_input_names(m::B200Model) = get_input_names(m.model)
get_output_names(m::B200Model) = get_output_names(m.model)
get_state_names(m::B200Model) = get_state_names(m.model)
get_param_names(m::B200Model) = get_param_names(m.model)
get_nn_names(m::B200Model) = get_nn_names(m.model)

function B200Model(model::HydroModel; fluxes_of, htypes::Vector{Int}, uh_kmax::Vector{Int}=Int[])
    # `uh_kmax[u]`: ordinates reserved for unit hydrograph u (compile-time; >= ceil(max_lag) of every parameter set it will see)
    plan = lower_stepper(model; fluxes_of=fluxes_of, uh_kmax=uh_kmax)
    kmax8 = ntuple(i -> Int32(i <= length(plan.uh_kmax) ? plan.uh_kmax[i] : 0), 8)
    # 20 fields, in the order of hb_stepper_spec
    spec = Ref(HbStepperSpec(sizeof(HbStepperSpec), length(plan.input_names), length(plan.row_names), length(plan.state_names),
                             length(plan.param_names), length(plan.uh_kmax), kmax8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    hb_check(ccall((:hb_compile_stepper, LIB), Cint, (Ref{HbStepperSpec}, Cstring, Ref{Ptr{Cvoid}}), spec, plan.body, h))
    B200Model(model, h[], plan, htypes)
end

function (m::B200Model)(input::AbstractArray{Float64,3}, params::ComponentVector, config=HydroModels.default_config();
                        initstates=nothing)
    cfg = HydroModels.normalize_config(config)
    @assert size(input, 1) == length(get_input_names(m)) "Input variables length mismatch"
    V, N, T = size(input)
    length(m.htypes) == N || throw(DimensionMismatch("htypes has $(length(m.htypes)) entries but the input has $N nodes"))
    # p[htypes] tables (src/utils.jl:174-201): flat values + (offset, mode) per name; htypes 0-based on the device
    flat = Float64[]; offs = Int32[]; modes = Int32[]
    identity_types = m.htypes == collect(1:N)
    for nm in m.plan.param_names
        haskey(params[:params], nm) || throw(ArgumentError("Parameter $nm does not exist in params"))
        v = collect(Float64, params[:params][nm])
        (minimum(m.htypes) < 1 || maximum(m.htypes) > length(v)) && throw(BoundsError(v, m.htypes))
        push!(offs, length(flat)); append!(flat, v)
        push!(modes, (identity_types && length(v) == N) ? 1 : 2)
    end
    ht0 = Int32.(m.htypes .- 1)
    S = length(m.plan.state_names)
    init = zeros(Float64, N * S)                      # state-major, node-fastest (src/model.jl:116-119)
    if initstates !== nothing
        for (i, s) in enumerate(m.plan.state_names)
            init[(i-1)*N+1:i*N] .= initstates isa ComponentVector ? initstates[s] : initstates[(i-1)*N+1:i*N]
        end
    end
    # unit-hydrograph ordinates per node: [sum kmax][N], evaluated on the host per parameter TYPE and gathered through htypes
    uhw = zeros(Float64, N * sum(m.plan.uh_kmax; init=0))
    row = 0
    for (uh, K) in zip(m.plan.uh_components, m.plan.uh_kmax)
        ntypes = maximum(m.htypes)
        tab = [uh_ordinates(uh, ComponentVector(params=NamedTuple{Tuple(get_param_names(uh))}(
                   Tuple(params[:params][nm][ty] for nm in get_param_names(uh)))), K) for ty in 1:ntypes]
        for k in 1:K, n in 1:N
            uhw[(row + k - 1) * N + n] = tab[m.htypes[n]][k]
        end
        row += K
    end
    inp = Array(input)                                # Julia layout [V,N,T] == kernel layout [T][N][V]; a plain Array is pageable:
                                                      # hb_run stages it through its pinned ring (hb_malloc_host + unsafe_wrap for full PCIe rate)
    out = Array{Float64,3}(undef, length(m.plan.row_names), N, T)
    GC.@preserve inp flat offs modes ht0 init out uhw begin
        # 27 fields, in the order of hb_run_desc
        d = Ref(HbRunDesc(sizeof(HbRunDesc), 1, N, T, pointer(inp), pointer(flat), length(flat), pointer(offs), pointer(modes),
                          pointer(ht0), pointer(init), isempty(uhw) ? C_NULL : pointer(uhw), C_NULL, C_NULL, C_NULL, 0, 1, cfg.min_value,
                          pointer(out), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, ntuple(_ -> Int32(0), 4)))
        hb_check(ccall((:hb_run, LIB), Cint, (Ptr{Cvoid}, Ref{HbRunDesc}), m.handle, d))
    end
    return out
end

export B200Model, lower_stepper
end # module
