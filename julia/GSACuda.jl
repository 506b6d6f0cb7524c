# GSACuda.jl -- the `ccall` shim a GlobalSensitivity.jl maintainer adds to route
#     gsa(model::DeviceModel, Sobol(...), A, B; batch = true)
# and the analysis of a precomputed `all_y` through libgsa_cuda.so (include/gsa_cuda.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It mirrors, call for call, the ctypes
# binding in globalsensitivity.jl_b200/_lib.py + sobol.py, which IS exercised by tests/ on the B200.
#
# Everything that is not Float64 / DeviceModel falls through to the reference's own methods
# (src/sobol_sensitivity.jl:110-168), because the method below is strictly more specific than `f::Any`.
module GSACuda

using GlobalSensitivity
import GlobalSensitivity: gsa, Sobol, SobolResult

const libgsa = get(ENV, "LIBGSA_CUDA", "libgsa_cuda.so")

# ---- mirrors of the C structs (include/gsa_cuda.h) ---------------------------------------------------------
struct GsaSobolOpts                # gsa_sobol_opts
    second_order::Cint
    nboot::Cint
    conf_level::Cdouble
    ei_estimator::Cint             # 0 Homma1996, 1 Sobol2007, 2 Jansen1999, 3 Janon2014
    input_location::Cint           # 0 host, 1 device
end

struct GsaSobolResult              # gsa_sobol_result: caller-allocated; C_NULL members are skipped
    S1::Ptr{Cdouble}
    S1_conf_int::Ptr{Cdouble}
    S2::Ptr{Cdouble}
    S2_conf_int::Ptr{Cdouble}
    ST::Ptr{Cdouble}
    ST_conf_int::Ptr{Cdouble}
end

const EI = Dict(:Homma1996 => 0, :Sobol2007 => 1, :Jansen1999 => 2, :Janon2014 => 3)

function check(rc::Cint)
    rc == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:gsa_last_error, libgsa), Cint, (Ptr{UInt8}, Csize_t), buf, 512)
    error("libgsa_cuda error $rc: ", unsafe_string(pointer(buf)))
end

# ---- handle ---------------------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer = -1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gsa_create, libgsa), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        h = new(r[])
        finalizer(x -> ccall((:gsa_destroy, libgsa), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end
const DEFAULT = Ref{Union{Nothing, Handle}}(nothing)
default_handle() = something(DEFAULT[], (DEFAULT[] = Handle()))

# ---- pin a Julia array in place: host designs are then streamed at PCIe speed instead of staged (unpin! before it is freed) ----
pin!(A::Array{Float64}) = (check(ccall((:gsa_host_register, libgsa), Cint, (Ptr{Cvoid}, Csize_t), A, sizeof(A))); A)
unpin!(A::Array{Float64}) = (check(ccall((:gsa_host_unregister, libgsa), Cint, (Ptr{Cvoid},), A)); A)

# ---- the device-model registry replaces the Julia closure `f` -----------------------------------------------
struct DeviceModel
    id::Cint
    params::Vector{Float64}        # a 2-D W is passed as vec(W) (column-major, as the C side expects)
    nout::Int                      # user models: fixed at registration; built-in models: 0 = ask the library
end
DeviceModel(id::Integer, params::Vector{Float64}) = DeviceModel(Cint(id), params, 0)
function DeviceModel(name::AbstractString, params = Float64[])
    id = Ref{Cint}(0)
    check(ccall((:gsa_model_lookup, libgsa), Cint, (Cstring, Ref{Cint}), name, id))
    return DeviceModel(id[], vec(Float64.(params)))
end
function nout(m::DeviceModel, d::Integer)
    m.nout > 0 && return m.nout
    n = Ref{Cint}(0)
    check(ccall((:gsa_model_nout, libgsa), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ref{Cint}),
                C_NULL, m.id, length(m.params), d, n))
    return Int(n[])
end

# ---- result allocation with the reference's shapes (src/sobol_sensitivity.jl:397-404) ------------------------
function alloc(no, d, nboot, second)
    multi = no > 1
    S1 = zeros(no, d); ST = zeros(no, d)
    S1c = nboot > 1 ? zeros(no, d) : nothing
    STc = nboot > 1 ? zeros(no, d) : nothing
    S2 = second ? zeros(d, d, no) : nothing
    S2c = second && nboot > 1 ? zeros(d, d, no) : nothing
    p(x) = x === nothing ? Ptr{Cdouble}(C_NULL) : pointer(x)
    res = GsaSobolResult(p(S1), p(S1c), p(S2), p(S2c), p(ST), p(STc))
    v(x) = x === nothing ? nothing : (multi ? x : vec(x))
    m(x) = x === nothing ? nothing : (multi ? x : x[:, :, 1])
    finish() = SobolResult(v(S1), v(S1c), m(S2), m(S2c), v(ST), v(STc))
    return res, finish, (S1, S1c, S2, S2c, ST, STc)
end

# ---- gsa(model, Sobol(...), A, B; batch = true) ----------------------------------------------------------------
# More specific than `gsa(f, method::Sobol, A::AbstractMatrix{TA}, B::AbstractMatrix; ...)` (reference :110-115).
function gsa(model::DeviceModel, method::Sobol, A::Matrix{Float64}, B::Matrix{Float64};
             batch = true, Ei_estimator = :Jansen1999, handle::Handle = default_handle(), kwargs...)
    size(A) == size(B) || throw(DimensionMismatch("A and B must have the same size"))
    d, N = size(A)
    second = 2 in method.order
    opts = Ref(GsaSobolOpts(second, method.nboot, method.conf_level, EI[Ei_estimator], 0))
    res, finish, keep = alloc(nout(model, d), d, method.nboot, second)
    GC.@preserve A B keep model begin
        check(ccall((:gsa_sobol_run, libgsa), Cint,
                    (Ptr{Cvoid}, Ref{GsaSobolOpts}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Int64,
                     Ref{GsaSobolResult}),
                    handle.ptr, opts, model.id, model.params, length(model.params), A, B, d, N, Ref(res)))
    end
    return finish()
end

# Float32 / Float16 designs (reference test/interface_tests.jl:21-31): computed in Float64 on the device, results returned in the
# element type of the design, as the reference's generic code would.  BigFloat and every other element type have no method here
# and stay on the reference's own code (they need arithmetic the device does not have).
function gsa(model::DeviceModel, method::Sobol, A::AbstractMatrix{T}, B::AbstractMatrix{T}; kwargs...) where {T <: Union{Float16, Float32}}
    r = gsa(model, method, Matrix{Float64}(A), Matrix{Float64}(B); kwargs...)
    c(x) = x === nothing ? nothing : T.(x)
    return SobolResult(c(r.S1), c(r.S1_Conf_Int), c(r.S2), c(r.S2_Conf_Int), c(r.ST), c(r.ST_Conf_Int))
end

# ---- gsa(model, Sobol(...), p_range; samples): design generated on the GPU (reference :407-417) ---------------
function generate_design_matrices(n::Integer, lb::Vector{Float64}, ub::Vector{Float64}, num_mats::Integer = 2;
                                  handle::Handle = default_handle())
    d = length(lb)
    mats = [Matrix{Float64}(undef, d, n) for _ in 1:num_mats]
    ptrs = [pointer(m) for m in mats]
    GC.@preserve mats begin
        check(ccall((:gsa_sobol_design, libgsa), Cint,
                    (Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64, Ptr{Ptr{Cdouble}}, Cint),
                    handle.ptr, n, d, num_mats, lb, ub, 0, n, ptrs, 0))
    end
    return mats
end

function gsa(model::DeviceModel, method::Sobol, p_range::AbstractVector; samples, Ei_estimator = :Jansen1999,
             handle::Handle = default_handle(), kwargs...)
    lb = [float(p[1]) for p in p_range]; ub = [float(p[2]) for p in p_range]
    # first choice: the design is generated inside the fused kernel (no A/B in memory at all) ...
    d = length(lb); second = 2 in method.order
    opts = Ref(GsaSobolOpts(second, method.nboot, method.conf_level, EI[Ei_estimator], 0))   # lb, ub are host vectors
    res, finish, keep = alloc(nout(model, d), d, method.nboot, second)
    rc = GC.@preserve keep model lb ub ccall((:gsa_sobol_run_generated, libgsa), Cint,
        (Ptr{Cvoid}, Ref{GsaSobolOpts}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Int64, Ref{GsaSobolResult}),
        handle.ptr, opts, model.id, model.params, length(model.params), lb, ub, d, samples, Ref(res))
    rc == 0 && return finish()
    rc == 3 || check(rc)              # GSA_ERR_UNSUPPORTED: no generating kernel for this model/options
    # ... otherwise materialise the design on the GPU, as the reference does on the CPU (:408-415)
    AB = generate_design_matrices(samples, lb, ub, 2 * method.nboot; handle = handle)
    A = reduce(hcat, AB[1:method.nboot]); B = reduce(hcat, AB[(method.nboot + 1):end])
    return gsa(model, method, A, B; Ei_estimator = Ei_estimator, handle = handle, kwargs...)
end

# ---- estimator-only entry: any Julia closure keeps producing all_y, the GPU replaces gsa_sobol_all_y_analysis ---
# (reference :169-405).  `all_y` is a Vector (scalar model) or an nout x cols Matrix.
function gsa_sobol_all_y_analysis(method::Sobol, all_y::VecOrMat{Float64}, d::Integer, n::Integer,
                                  Ei_estimator = :Jansen1999; handle::Handle = default_handle())
    multi = all_y isa AbstractMatrix
    no = multi ? size(all_y, 1) : 1
    second = 2 in method.order
    opts = Ref(GsaSobolOpts(second, method.nboot, method.conf_level, EI[Ei_estimator], 0))
    res, finish, keep = alloc(no, d, method.nboot, second)
    GC.@preserve all_y keep begin
        check(ccall((:gsa_sobol_analyze_y, libgsa), Cint,
                    (Ptr{Cvoid}, Ref{GsaSobolOpts}, Ptr{Cdouble}, Cint, Cint, Int64, Ref{GsaSobolResult}),
                    handle.ptr, opts, all_y, no, d, n, Ref(res)))
    end
    r = finish()
    # a 1 x cols matrix is still "multioutput" in the reference (:144): keep 1 x d results
    if multi && no == 1
        r = SobolResult(reshape(r.S1, 1, :), r.S1_Conf_Int === nothing ? nothing : reshape(r.S1_Conf_Int, 1, :),
                        r.S2 === nothing ? nothing : reshape(r.S2, d, d, 1),
                        r.S2_Conf_Int === nothing ? nothing : reshape(r.S2_Conf_Int, d, d, 1),
                        reshape(r.ST, 1, :), r.ST_Conf_Int === nothing ? nothing : reshape(r.ST_Conf_Int, 1, :))
    end
    return r
end

# ---- one process, all GPUs (gsa_multi_*): contiguous column shards, one gather-and-add kernel over NVLink peer memory for the partial sums ----------------
mutable struct Multi
    ptr::Ptr{Cvoid}
    # Multi() drives ALL visible GPUs (ndev = 0 asks the library to enumerate them); Multi([0, 2]) the listed ones
    function Multi(devices::Vector{<:Integer} = Int[])
        r = Ref{Ptr{Cvoid}}(C_NULL)
        devs = Cint.(devices)
        check(ccall((:gsa_multi_create, libgsa), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}),
                    isempty(devs) ? C_NULL : devs, length(devs), r))
        m = new(r[])
        finalizer(x -> ccall((:gsa_multi_destroy, libgsa), Cint, (Ptr{Cvoid},), x.ptr), m)
        return m
    end
end

function gsa(m::Multi, model::DeviceModel, method::Sobol, A::Matrix{Float64}, B::Matrix{Float64}; Ei_estimator = :Jansen1999)
    d, N = size(A); second = 2 in method.order
    opts = Ref(GsaSobolOpts(second, method.nboot, method.conf_level, EI[Ei_estimator], 0))
    res, finish, keep = alloc(nout(model, d), d, method.nboot, second)
    GC.@preserve A B keep model begin
        check(ccall((:gsa_multi_sobol_run, libgsa), Cint,
                    (Ptr{Cvoid}, Ref{GsaSobolOpts}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Int64, Ref{GsaSobolResult}),
                    m.ptr, opts, model.id, model.params, length(model.params), A, B, d, N, Ref(res)))
    end
    return finish()
end

# gsa(model, Sobol(...), p_range; samples) over all devices with the design generated inside the fused kernel (no A/B at all)
function gsa(m::Multi, model::DeviceModel, method::Sobol, p_range::AbstractVector; samples, Ei_estimator = :Jansen1999)
    lb = [float(p[1]) for p in p_range]; ub = [float(p[2]) for p in p_range]
    d = length(lb); second = 2 in method.order
    opts = Ref(GsaSobolOpts(second, method.nboot, method.conf_level, EI[Ei_estimator], 0))
    res, finish, keep = alloc(nout(model, d), d, method.nboot, second)
    GC.@preserve keep model lb ub begin
        check(ccall((:gsa_multi_sobol_run_generated, libgsa), Cint,
                    (Ptr{Cvoid}, Ref{GsaSobolOpts}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Int64, Ref{GsaSobolResult}),
                    m.ptr, opts, model.id, model.params, length(model.params), lb, ub, d, samples, Ref(res)))
    end
    return finish()
end

# ---- user `__device__` models: a relocatable fatbin (nvcc -rdc=true -fatbin -gencode arch=compute_100a,code=sm_100a) that defines
#      extern "C" __device__ void <symbol>(const double* x, int d, const double* p, int np, double* y, int nout) -----------------
function DeviceModel(image::Vector{UInt8}, symbol::AbstractString, params = Float64[]; nout::Integer = 1, handle::Handle = default_handle())
    id = Ref{Cint}(0)
    GC.@preserve image begin
        check(ccall((:gsa_model_register_fatbin, libgsa), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Cstring, Cint, Ref{Cint}),
                    handle.ptr, image, length(image), symbol, nout, id))
    end
    return DeviceModel(id[], vec(Float64.(params)), Int(nout))
end
# Separable user models: kind = :product (f(x) = prod_k g(k, x_k)) or :additive (f(x) = sum_k g(k, x_k)), with
#      extern "C" __device__ double <symbol>(int k, double x_k, const double* p, int np)
# run first/total order on the O(d)-per-sample kernel (gsa_model_register_separable).
const USER_KIND = Dict(:generic => 0, :product => 1, :additive => 2)
function SeparableModel(image::Vector{UInt8}, symbol::AbstractString, kind::Symbol, params = Float64[]; handle::Handle = default_handle())
    id = Ref{Cint}(0)
    GC.@preserve image begin
        check(ccall((:gsa_model_register_separable, libgsa), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Cstring, Cint, Ref{Cint}),
                    handle.ptr, image, length(image), symbol, USER_KIND[kind], id))
    end
    return DeviceModel(id[], vec(Float64.(params)), 1)
end
function SeparableModel(m::Multi, image::Vector{UInt8}, symbol::AbstractString, kind::Symbol, params = Float64[])
    id = Ref{Cint}(0)
    GC.@preserve image begin
        check(ccall((:gsa_multi_model_register_separable, libgsa), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Cstring, Cint, Ref{Cint}),
                    m.ptr, image, length(image), symbol, USER_KIND[kind], id))
    end
    return DeviceModel(id[], vec(Float64.(params)), 1)
end

# ---- the sharded form, one process per GPU (Distributed.jl / MPI.jl workers): ONE call per rank and step -----------------
# (the ranks exchange the 64-byte mailbox handles of gsa_peer_export once, e.g. with MPI.Allgather, then gsa_peer_connect)
function gsa_sharded(model::DeviceModel, method::Sobol, A_local::Matrix{Float64}, B_local::Matrix{Float64}, N::Integer, col0::Integer;
                     Ei_estimator = :Jansen1999, handle::Handle = default_handle())
    d, ncols = size(A_local); second = 2 in method.order
    opts = Ref(GsaSobolOpts(second, method.nboot, method.conf_level, EI[Ei_estimator], 0))
    res, finish, keep = alloc(nout(model, d), d, method.nboot, second)
    GC.@preserve A_local B_local keep model begin
        check(ccall((:gsa_sobol_run_sharded, libgsa), Cint,
                    (Ptr{Cvoid}, Ref{GsaSobolOpts}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Int64, Int64, Int64,
                     Ref{GsaSobolResult}),
                    handle.ptr, opts, model.id, model.params, length(model.params), A_local, B_local, d, N, col0, ncols, Ref(res)))
    end
    return finish()
end

function DeviceModel(m::Multi, image::Vector{UInt8}, symbol::AbstractString, params = Float64[]; nout::Integer = 1)
    id = Ref{Cint}(0)
    GC.@preserve image begin
        check(ccall((:gsa_multi_model_register_fatbin, libgsa), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Cstring, Cint, Ref{Cint}),
                    m.ptr, image, length(image), symbol, nout, id))
    end
    return DeviceModel(id[], vec(Float64.(params)), Int(nout))
end

# ---- eFAST (reference src/eFAST_sensitivity.jl:75-213): design generated in registers, device model, spectral analysis -------
struct GsaEfastResult
    S1::Ptr{Cdouble}
    ST::Ptr{Cdouble}
end
function gsa(model::DeviceModel, method::GlobalSensitivity.eFAST, p_range::AbstractVector; samples::Int, batch = true,
             phi::Vector{Float64} = 2 .* rand(length(p_range)), handle::Handle = default_handle(), kwargs...)
    lb = [float(p[1]) for p in p_range]; ub = [float(p[2]) for p in p_range]
    d = length(lb); no = nout(model, d)
    S1 = zeros(no, d); ST = zeros(no, d)
    res = GsaEfastResult(pointer(S1), pointer(ST))
    GC.@preserve S1 ST model lb ub phi begin
        check(ccall((:gsa_efast_run, libgsa), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Int64, Ref{GsaEfastResult}),
                    handle.ptr, method.num_harmonics, model.id, model.params, length(model.params), lb, ub, phi, d, samples, Ref(res)))
    end
    return GlobalSensitivity.eFASTResult(S1, ST)
end
# the analysis alone, for any Julia closure: all_y = f(ps) as at src/eFAST_sensitivity.jl:137
function gsa_efast_all_y_analysis(method::GlobalSensitivity.eFAST, all_y::VecOrMat{Float64}, num_params::Integer, samples::Integer;
                                  handle::Handle = default_handle())
    no = all_y isa AbstractMatrix ? size(all_y, 1) : 1
    S1 = zeros(no, num_params); ST = zeros(no, num_params)
    res = GsaEfastResult(pointer(S1), pointer(ST))
    GC.@preserve S1 ST all_y begin
        check(ccall((:gsa_efast_analyze_y, libgsa), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint, Cint, Int64, Ref{GsaEfastResult}),
                    handle.ptr, method.num_harmonics, all_y, 0, no, num_params, samples, Ref(res)))
    end
    return GlobalSensitivity.eFASTResult(S1, ST)
end

# QuasiMonteCarlo.sample(n, lb, ub, SobolSample()): the single-matrix design the reference's other methods draw (SURVEY 8(f) N3)
sample(n::Integer, lb::Vector{Float64}, ub::Vector{Float64}; handle::Handle = default_handle()) =
    generate_design_matrices(n, lb, ub, 1; handle = handle)[1]

end # module
