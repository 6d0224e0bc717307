# PatchMixtureKrigingB200.jl -- Julia host layer over libpmk_b200.so (C ABI: include/pmk.h).
#
# Drop-in for the fit / query hot path of PatchMixtureKriging.jl: the functions below keep the
# reference's names, argument order and return values (reference file:line in each docstring) and
# replace only the BODIES with `ccall`s.  Everything the north star keeps on the host stays the
# reference's own Julia code: `setuppartition`, `organizetrainingsets`, `fetchhyperplanes`,
# `findpartition`, `findneighbourpartitions`, the kernel parameter structs and `array2matrix`
# (optional device forms of `setuppartition` and `organizetrainingsets` are further down).
# `MixtureGPType(X_set, hps; devices = 0:7)` shards the model over the GPUs of the box by sub-tree ownership
# (pmk_multi_*: rank r owns a contiguous range of the leaves; nothing but the tree is replicated); `fitmixtureGP!` and
# `querymixtureGP!` keep their signatures.
#
# NOTE: this image has no Julia toolchain, so this file is written against the C ABI but has not been
# executed here; the same ABI is exercised call-for-call by the Python mirror
# (patchmixturekriging_b200/*.py) in tests/test_gpu_parity.py.  No CUDA.jl arrays, no CPU fallback:
# if the library or a B200 is missing, `pmk_create` fails and an error is thrown.
module PatchMixtureKrigingB200

using LinearAlgebra
import PatchMixtureKriging
const PMK = PatchMixtureKriging

const libpmk = get(ENV, "PMK_B200_LIB", joinpath(@__DIR__, "..", "patchmixturekriging_b200", "libpmk_b200.so"))

const PMK_OK = Cint(0)
const PMK_ERR_NOT_POSDEF = Cint(-3)

# ---- kernel ids (include/pmk.h pmk_kernel_id; reference src/misc/declarations.jl:25-100) --------------
kernelid(::PMK.GaussianKernel1DType) = Cint(0)
kernelid(::PMK.Spline34KernelType) = Cint(1)
kernelid(::PMK.BrownianBridge10) = Cint(2)
kernelid(::PMK.BrownianBridge20) = Cint(3)
kernelid(::PMK.BrownianBridge1ϵ) = Cint(4)
kernelid(::PMK.BrownianBridge2ϵ) = Cint(5)
kernelid(::PMK.Spline12KernelType) = Cint(6)
kernelid(::PMK.Spline32KernelType) = Cint(7)
kernelid(::PMK.RationalQuadraticKernelType) = Cint(8)
kernelparams(θ::PMK.GaussianKernel1DType) = Float64[θ.ϵ_sq[1]]
kernelparams(θ::Union{PMK.BrownianBridge1ϵ,PMK.BrownianBridge2ϵ}) = Float64[θ.ϵ]
kernelparams(θ) = Float64[θ.a[1]]

# ---- handle ---------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:pmk_create, libpmk), Cint, (Ref{Ptr{Cvoid}}, Cint), out, device)
        rc == PMK_OK || error("pmk_create: ", unsafe_string(ccall((:pmk_last_error, libpmk), Cstring, (Ptr{Cvoid},), C_NULL)))
        h = new(out[])
        finalizer(x -> ccall((:pmk_destroy, libpmk), Cvoid, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

lasterror(h::Handle) = unsafe_string(ccall((:pmk_last_error, libpmk), Cstring, (Ptr{Cvoid},), h.ptr))

"One model over several GPUs of the box (pmk_multi, include/pmk.h): `devices` are CUDA ordinals."
mutable struct MultiHandle
    ptr::Ptr{Cvoid}
    devices::Vector{Cint}
    function MultiHandle(devices)
        ids = Cint.(collect(devices))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = GC.@preserve ids ccall((:pmk_multi_create, libpmk), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), out, length(ids), ids)
        rc == PMK_OK || error("pmk_multi_create: ", unsafe_string(ccall((:pmk_multi_last_error, libpmk), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(out[], ids)
        finalizer(x -> ccall((:pmk_multi_destroy, libpmk), Cvoid, (Ptr{Cvoid},), x.ptr), m)
        return m
    end
end
lasterror(m::MultiHandle) = unsafe_string(ccall((:pmk_multi_last_error, libpmk), Cstring, (Ptr{Cvoid},), m.ptr))
function check(m::MultiHandle, rc::Cint)
    rc == PMK_OK && return nothing
    rc == Cint(-2) && throw(DimensionMismatch(lasterror(m)))
    error("libpmk_b200 error $(rc): ", lasterror(m))
end
"the single-GPU handle of the rank that owns (1-based) leaf `n` of an `N`-leaf model: leaf ids stay global"
function ownerhandle(m::MultiHandle, n::Integer, N::Integer)
    first = Ref{Int64}(0); count = Ref{Int64}(0)
    for r in 0:length(m.devices)-1
        check(m, ccall((:pmk_multi_owned_range, libpmk), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}, Ref{Int64}), m.ptr, r, first, count))
        if first[] < n <= first[] + count[]
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(m, ccall((:pmk_multi_handle, libpmk), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), m.ptr, r, h))
            return h[]
        end
    end
    throw(BoundsError())
end

function check(h::Handle, rc::Cint)
    rc == PMK_OK && return nothing
    rc == Cint(-2) && throw(DimensionMismatch(lasterror(h)))       # mixtureGP.jl:298, RKHS.jl:18,199-203,225-227
    error("libpmk_b200 error $(rc): ", lasterror(h))
end

# Vector{Vector{Float64}} -> D x n Matrix: exactly src/misc/utilities.jl:25-36
pack(X::Vector{Vector{Float64}}) = PMK.array2matrix(X)

# ---- constructkernelmatrix(X, θ)   (src/RKHS/RKHS.jl:4-34) -----------------------------------------
const _shared = Ref{Union{Nothing,Handle}}(nothing)
sharedhandle() = (_shared[] === nothing && (_shared[] = Handle(0)); _shared[]::Handle)

function constructkernelmatrix(X::Vector{Vector{Float64}}, θ)::Matrix{Float64}
    h = sharedhandle()
    Xm = pack(X); n = length(X); kp = kernelparams(θ)
    K = Matrix{Float64}(undef, n, n)
    GC.@preserve Xm kp K check(h, ccall((:pmk_gram, libpmk), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64, Ptr{Float64}),
        h.ptr, size(Xm, 1), n, Xm, kernelid(θ), kp, length(kp), 0.0, K))
    return K
end

"constructkernelmatrix(X, Z, θ): K[i,j] = evalkernel(X[i], Z[j], θ)   (src/RKHS/RKHS.jl:95-110)"
function constructkernelmatrix(X::Vector{Vector{Float64}}, Z::Vector{Vector{Float64}}, θ)::Matrix{Float64}
    h = sharedhandle()
    Xm = pack(X); Zm = pack(Z); kp = kernelparams(θ)
    K = Matrix{Float64}(undef, length(X), length(Z))
    GC.@preserve Xm Zm kp K check(h, ccall((:pmk_cross_gram, libpmk), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}),
        h.ptr, size(Xm, 1), length(X), Xm, length(Z), Zm, kernelid(θ), kp, length(kp), K))
    return K
end

"evalquery(x, c, X, θ) = Σ c[n] k(x, X[n])   (src/RKHS/querying.jl:2-5)"
evalquery(x::Vector{Float64}, c::Vector{Float64}, X::Vector{Vector{Float64}}, θ)::Float64 =
    dot(vec(constructkernelmatrix([x], X, θ)), c)

"""
setupGPquery(c, X, θ, σ²) -> fq   (src/RKHS/querying.jl:43-58); fq(xq) = evalqueryGP!(...) -> (mean, variance) with
mean = Σ c[n] k(xq, X[n]) and variance = k(xq,xq) - kᵀ(K + σ²I)⁻¹k, NOT clamped (querying.jl:60-79).  The reference solves
A\k by LU per query; here the leaf is factorised once and the fused pair kernel returns ‖L⁻¹k‖² (flags bit 1 = no clamp).
"""
function setupGPquery(c::Vector{Float64}, X::Vector{Vector{Float64}}, θ, σ²::Float64)::Function
    length(c) == length(X) || throw(DimensionMismatch("length(c) != length(X)"))
    h = Handle(0)
    Xm = pack(X); kp = kernelparams(θ); off = Int64[0, length(X)]; y0 = zeros(length(X))
    bad = Ref{Int64}(0); info = Ref{Cint}(0)
    rc = GC.@preserve Xm kp off y0 ccall((:pmk_fit, libpmk), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64, Ref{Int64}, Ref{Cint}),
        h.ptr, size(Xm, 1), 1, off, Xm, y0, kernelid(θ), kp, length(kp), σ², bad, info)
    rc == PMK_ERR_NOT_POSDEF && throw(PosDefException(info[]))
    check(h, rc)
    GC.@preserve c check(h, ccall((:pmk_set_alpha, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), h.ptr, 1, c))
    check(h, ccall((:pmk_set_tree, libpmk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}), h.ptr, size(Xm, 1), 1, C_NULL, C_NULL))
    wp = Float64[1.0]
    function fq(xq::Vector{Float64})
        Yq = zeros(1); Vq = zeros(1); Xq = reshape(copy(xq), :, 1)
        GC.@preserve Xq wp Yq Vq check(h, ccall((:pmk_query, libpmk), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
            h.ptr, 1, Xq, 0.0, 0.0, 1, wp, 1, 2, Yq, Vq))
        return Yq[1], Vq[1]
    end
    return fq
end

# ---- MixtureGPType: the fitted state lives in HBM behind the handle -----------------------------------
mutable struct MixtureGPType{T}
    X_parts::Vector{Vector{Vector{T}}}
    hps::Vector{PMK.HyperplaneType{T}}
    h::Union{Handle,Nothing}              # single-GPU model
    multi::Union{MultiHandle,Nothing}     # model sharded over several GPUs (sub-tree ownership)
    σ²_set::Vector{T}
    fitted::Bool
    θ                                     # the kernel of the last fit (queries must use the same one)
    treekey::UInt                         # hash of (levels, hyperplanes) last uploaded
end

"MixtureGPType(X_set, hps)  (src/RKHS/mixtureGP.jl:54-66); `devices = 0:7` shards the model over those GPUs"
function MixtureGPType(X_parts::Vector{Vector{Vector{T}}}, hps::Vector{PMK.HyperplaneType{T}}; device = 0, devices = nothing) where T
    devices === nothing ? MixtureGPType{T}(X_parts, hps, Handle(device), nothing, T[], false, nothing, UInt(0)) :
                          MixtureGPType{T}(X_parts, hps, nothing, MultiHandle(devices), T[], false, nothing, UInt(0))
end
leafhandle(η::MixtureGPType, n::Integer) = η.multi === nothing ? η.h.ptr : ownerhandle(η.multi, n, length(η.X_parts))
checkleaf(η::MixtureGPType, rc::Cint) = η.multi === nothing ? check(η.h, rc) : (rc == PMK_OK || error("libpmk_b200 error $(rc) on the leaf's owner"))

"c_set[n], L_set[n], U_set[n] of the reference's struct (mixtureGP.jl:40-46), fetched on demand (1-based leaf)"
function c_set(η::MixtureGPType, n::Integer)
    out = Vector{Float64}(undef, length(η.X_parts[n]))
    GC.@preserve out checkleaf(η, ccall((:pmk_get_alpha, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), leafhandle(η, n), n, out))
    return out
end
function L_set(η::MixtureGPType, n::Integer)
    m = length(η.X_parts[n]); out = Matrix{Float64}(undef, m, m)
    GC.@preserve out checkleaf(η, ccall((:pmk_get_L, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), leafhandle(η, n), n, out))
    return LowerTriangular(out)
end
function U_set(η::MixtureGPType, n::Integer)      # Gram WITHOUT σ² (mixtureGP.jl:99)
    m = length(η.X_parts[n]); out = Matrix{Float64}(undef, m, m)
    GC.@preserve out checkleaf(η, ccall((:pmk_get_K, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), leafhandle(η, n), n, out))
    return out
end

"fitmixtureGP!(η, y_parts, θ, σ²) -> η   (src/RKHS/mixtureGP.jl:70-118): all leaves in one batched GPU fit"
function fitmixtureGP!(η::MixtureGPType{T}, y_parts::Vector{Vector{T}}, θ, σ²) where T
    N_parts = length(η.X_parts)
    length(y_parts) == N_parts || throw(DimensionMismatch("length(y_parts) != length(η.X_parts)"))
    leaf_off = Int64[0; cumsum(length.(η.X_parts))]
    Xp = reduce(hcat, pack.(η.X_parts))            # D x Σn_p, leaves back to back
    yp = reduce(vcat, y_parts)
    length(yp) == leaf_off[end] || throw(DimensionMismatch("length(y) != length(X) in a leaf"))
    kp = kernelparams(θ)
    bad = Ref{Int64}(0); info = Ref{Cint}(0)
    sig = (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64, Ref{Int64}, Ref{Cint})
    rc = GC.@preserve leaf_off Xp yp kp (η.multi === nothing ?
        ccall((:pmk_fit, libpmk), Cint, sig, η.h.ptr, size(Xp, 1), N_parts, leaf_off, Xp, yp, kernelid(θ), kp, length(kp), Float64(σ²), bad, info) :
        ccall((:pmk_multi_fit, libpmk), Cint, sig, η.multi.ptr, size(Xp, 1), N_parts, leaf_off, Xp, yp, kernelid(θ), kp, length(kp), Float64(σ²), bad, info))
    rc == PMK_ERR_NOT_POSDEF && throw(PosDefException(info[]))      # cholesky(U) at mixtureGP.jl:109; bad[] = the failing leaf
    η.multi === nothing ? check(η.h, rc) : check(η.multi, rc)
    η.σ²_set = fill(T(σ²), N_parts); η.fitted = true; η.θ = θ
    return η
end

"""How `L \\ kq` of `queryinner!` (mixtureGP.jl:311) is carried out: -1 = chosen by the fit's conditioning estimate (default:
explicit inverse unless (max diag L / min diag L)² ≥ 1e4, then substitution), 0 = explicit inverse formed once per fit,
1 = blocked forward substitution (closest to `dtrsv`), 2 = explicit inverse with the round-1 column-sweep kernel.
`PMK_OPT_QUERY_SOLVER` of include/pmk.h."""
function setsolver!(η::MixtureGPType, solver::Integer)
    η.multi === nothing ? check(η.h, ccall((:pmk_set_option, libpmk), Cint, (Ptr{Cvoid}, Cint, Int64), η.h.ptr, 2, solver)) :
                          check(η.multi, ccall((:pmk_multi_set_option, libpmk), Cint, (Ptr{Cvoid}, Cint, Int64), η.multi.ptr, 2, solver))
    return η
end

"""How `P = inv(L)` is formed for the explicit-inverse solvers: 0 = recursive doubling on the packed tiles (default),
1 = the substitution kernel on identity right-hand sides.  `PMK_OPT_INVERSE_BUILDER` of include/pmk.h."""
function setinversebuilder!(η::MixtureGPType, builder::Integer)
    check(η.h, ccall((:pmk_set_option, libpmk), Cint, (Ptr{Cvoid}, Cint, Int64), η.h.ptr, 3, builder))
end

# flatten the reference's BinaryNode tree for the device: hyperplanes in fetchhyperplanes order (mixtureGP.jl:322-334)
function settree!(η::MixtureGPType, levels::Integer)
    D = length(η.X_parts[1][1]); hps = η.hps
    hv = Matrix{Float64}(undef, D, length(hps)); hc = Vector{Float64}(undef, length(hps))
    for (i, hp) in enumerate(hps)
        hv[:, i] = hp.v; hc[i] = hp.c
    end
    key = hash((levels, hv, hc))          # content, not identity: hyperplanes edited in place are uploaded again
    key == η.treekey && return nothing
    sig = (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64})
    GC.@preserve hv hc (η.multi === nothing ?
        check(η.h, ccall((:pmk_set_tree, libpmk), Cint, sig, η.h.ptr, D, levels, hv, hc)) :
        check(η.multi, ccall((:pmk_multi_set_tree, libpmk), Cint, sig, η.multi.ptr, D, levels, hv, hc)))
    η.treekey = key
    return nothing
end

"""
querymixtureGP!(Yq, Vq, Xq, η, root, levels, radius, δ, θ, σ², weight_θ, debug_vars; debug_flag=false)::Nothing
(src/RKHS/mixtureGP.jl:159-294).  `root` is accepted for signature compatibility; the device works on η.hps.
"""
function querymixtureGP!(Yq::Vector{T}, Vq::Vector{T}, Xq::Vector{Vector{T}}, η::MixtureGPType{T}, root, levels,
                         radius::T, δ::T, θ, σ², weight_θ, debug_vars = nothing; debug_flag = false)::Nothing where T
    Nq = length(Xq); resize!(Yq, Nq); resize!(Vq, Nq)
    η.fitted || error("querymixtureGP! before fitmixtureGP!")
    (kernelid(θ) == kernelid(η.θ) && kernelparams(θ) == kernelparams(η.θ)) || throw(ArgumentError("θ differs from the kernel the model was fitted with"))
    settree!(η, levels)
    Xm = pack(Xq); wp = kernelparams(weight_θ)
    sig = (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64})
    if η.multi !== nothing
        debug_flag && error("debug_flag on a model sharded over several GPUs: query a single-GPU model for the debug outputs")
        GC.@preserve Xm wp Yq Vq check(η.multi, ccall((:pmk_multi_query, libpmk), Cint, sig,
            η.multi.ptr, Nq, Xm, radius, δ, kernelid(weight_θ), wp, length(wp), 0, Yq, Vq))
        return nothing
    end
    GC.@preserve Xm wp Yq Vq check(η.h, ccall((:pmk_query, libpmk), Cint, sig,
        η.h.ptr, Nq, Xm, radius, δ, kernelid(weight_θ), wp, length(wp), 0, Yq, Vq))
    if debug_flag && debug_vars !== nothing
        fetchdebug!(debug_vars, η, Nq)
    end
    return nothing
end

"MixtureGPDebugType outputs (mixtureGP.jl:5-35,242-260) from the last query, CSR -> per-query vectors"
function fetchdebug!(dv::PMK.MixtureGPDebugType, η::MixtureGPType, Nq::Integer)
    np = Ref{Int64}(0)
    check(η.h, ccall((:pmk_last_query_pairs, libpmk), Cint, (Ptr{Cvoid}, Ref{Int64}), η.h.ptr, np))
    P = np[]
    home = Vector{Int32}(undef, Nq); off = Vector{Int64}(undef, Nq + 1)
    leaf = Vector{Int32}(undef, P); hp = Vector{Int32}(undef, P)
    t = Vector{Float64}(undef, P); w = similar(t); u = similar(t); v = similar(t)
    GC.@preserve home off leaf hp t w u v check(η.h, ccall((:pmk_last_query_debug, libpmk), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        η.h.ptr, home, off, leaf, hp, t, w, u, v))
    rng(j) = (off[j] + 1):off[j + 1]
    dv.w_tilde_set = [w[rng(j)] for j in 1:Nq]
    dv.u_set = [u[rng(j)] for j in 1:Nq]
    dv.v_set = [v[rng(j)] for j in 1:Nq]
    dv.region_inds_set = [Int.(leaf[rng(j)][1:end-1]) for j in 1:Nq]
    dv.p_region_ind_set = Int.(home)
    # hps_keep_flags_set / zs_set / ts_set (mixtureGP.jl:17-19,256-258): dense over ALL hyperplanes, Nq * length(hps) entries --
    # fetched in slices of at most 2^26 entries (pmk_last_query_debug_dense); the kept ids and their t are also in `hp` / `t`
    n_hp = length(η.hps); D = length(η.X_parts[1][1])
    if n_hp > 0
        dv.hps_keep_flags_set = Vector{BitVector}(undef, Nq); dv.ts_set = Vector{Vector{Float64}}(undef, Nq)
        dv.zs_set = Vector{Vector{Vector{Float64}}}(undef, Nq)
        step = max(1, (1 << 26) ÷ n_hp)
        for j0 in 0:step:Nq-1
            m = min(step, Nq - j0)
            keep = Vector{UInt8}(undef, m * n_hp); ts = Vector{Float64}(undef, m * n_hp); zs = Array{Float64}(undef, D, n_hp, m)
            GC.@preserve keep ts zs check(η.h, ccall((:pmk_last_query_debug_dense, libpmk), Cint,
                (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}), η.h.ptr, j0, m, keep, ts, zs))
            for j in 1:m
                r = ((j - 1) * n_hp + 1):(j * n_hp)
                dv.hps_keep_flags_set[j0 + j] = BitVector(keep[r] .!= 0)
                dv.ts_set[j0 + j] = ts[r]
                dv.zs_set[j0 + j] = [zs[:, i, j] for i in 1:n_hp]
            end
        end
    end
    return dv
end

"querymixtureGP(Xq | xq, ...) -> (Yq, Vq, debug_vars)   (mixtureGP.jl:120-157)"
function querymixtureGP(Xq::Vector{Vector{T}}, η::MixtureGPType{T}, root, levels, radius::T, δ::T, θ, σ², weight_θ;
                        debug_flag = false) where T
    Yq = T[]; Vq = T[]; dv = PMK.MixtureGPDebugType(one(T))
    querymixtureGP!(Yq, Vq, Xq, η, root, levels, radius, δ, θ, σ², weight_θ, dv; debug_flag = debug_flag)
    return Yq, Vq, dv
end
querymixtureGP(xq::Vector{T}, η::MixtureGPType{T}, args...; kw...) where T <: Real = querymixtureGP([xq], η, args...; kw...)

# ---- single GP: fitRKHS! / query!   (src/RKHS/RKHS.jl:182-217, :220-247) ---------------------------------
const _rkhs_handles = IdDict{Any,Handle}()

"fitRKHS!(η, y): η.c[:] = (K + σ²I) \\ y"
function fitRKHS!(η::PMK.RKHSProblemType, y::Vector{Float64})
    @assert !isempty(η.X) && !isempty(y) && length(η.X) == length(y)
    h = get!(() -> Handle(0), _rkhs_handles, η)
    Xm = pack(η.X); kp = kernelparams(η.θ); off = Int64[0, length(y)]
    bad = Ref{Int64}(0); info = Ref{Cint}(0)
    rc = GC.@preserve Xm kp off y ccall((:pmk_fit, libpmk), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64, Ref{Int64}, Ref{Cint}),
        h.ptr, size(Xm, 1), 1, off, Xm, y, kernelid(η.θ), kp, length(kp), Float64(η.σ²[1]), bad, info)
    rc == PMK_ERR_NOT_POSDEF && throw(PosDefException(info[]))
    check(h, rc)
    check(h, ccall((:pmk_set_tree, libpmk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}), h.ptr, size(Xm, 1), 1, C_NULL, C_NULL))
    GC.@preserve η check(h, ccall((:pmk_get_alpha, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), h.ptr, 1, η.c))
    return nothing
end

"query!(Yq, Xq, η): Yq[iq] = dot(kq, η.c) (mean only)"
function query!(Yq::Vector{Float64}, Xq::Vector{Vector{Float64}}, η::PMK.RKHSProblemType)
    @assert !isempty(Xq) && size(Yq) == size(Xq)
    h = _rkhs_handles[η]
    Xm = pack(Xq); wp = Float64[1.0]
    GC.@preserve Xm wp Yq check(h, ccall((:pmk_query, libpmk), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
        h.ptr, length(Xq), Xm, 0.0, 0.0, 1, wp, 1, 1, Yq, C_NULL))
    return nothing
end

# ---- setuppartition(X, levels) on the device, one level per round trip  (src/patchwork/partition.jl:106-217) -------
# The O(N) work of every level -- mean(X) in Base's pairwise order, the projections dot(v, X[n]), median(f) and the
# order-preserving split X[left_indicators] / X[.!left_indicators] -- runs on the GPU (pmk_partition_*); the 1 x D svd of
# gethyperplane (partition.jl:90-94) is the reference's own LinearAlgebra call on z = X[1] - mean(X), so hp.v carries
# exactly the bits gethyperplane would have produced.  Returns (root, X_parts, X_parts_inds) like the reference.
function setuppartition(X::Vector{Vector{T}}, levels::Integer; h::Handle = sharedhandle()) where T
    D = length(X[1]); N = length(X); Xm = pack(X)
    GC.@preserve Xm check(h, ccall((:pmk_partition_begin, libpmk), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Cint),
                                   h.ptr, D, N, Xm, levels))
    hps_by_depth = Vector{Vector{PMK.HyperplaneType{T}}}()
    for depth in 0:levels-2
        nodes = 1 << depth
        z = Matrix{Float64}(undef, D, nodes); v = similar(z); c = Vector{Float64}(undef, nodes)
        GC.@preserve z check(h, ccall((:pmk_partition_level_z, libpmk), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h.ptr, depth, z))
        for j in 1:nodes
            Z_mat = (PMK.array2matrix([z[:, j]]))'            # partition.jl:91-92, same object the reference hands to svd
            v[:, j] = svd(Z_mat).V[:, 1]                      # partition.jl:93-94
        end
        GC.@preserve v c check(h, ccall((:pmk_partition_level_split, libpmk), Cint,
                                        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), h.ptr, depth, v, c))
        push!(hps_by_depth, [PMK.HyperplaneType{T}(v[:, j], c[j]) for j in 1:nodes])
    end
    n_leaves = 1 << (levels - 1)
    leaf_off = Vector{Int64}(undef, n_leaves + 1); inds = Vector{Int32}(undef, N)
    GC.@preserve leaf_off inds check(h, ccall((:pmk_partition_fetch, libpmk), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}),
                                              h.ptr, leaf_off, inds))
    X_parts_inds = [Int.(inds[leaf_off[p]+1:leaf_off[p+1]]) for p in 1:n_leaves]
    # rebuild the reference's BinaryNode tree (partition.jl:3-29, 166-217): node (depth, j) has children (depth+1, 2j-1 / 2j)
    emptyX() = Vector{Vector{T}}(undef, 0)
    root = PMK.BinaryNode(PMK.PartitionDataType(hps_by_depth[1][1], emptyX(), Int[], 0))
    function grow!(parent, depth, j)          # children of node j (1-based) at `depth`
        for (side, jc) in ((PMK.leftchild!, 2j - 1), (PMK.rightchild!, 2j))
            if depth + 1 == levels - 1        # kid is a leaf: keeps its global indices, gets its label (labelleafnodes :131-159)
                side(parent, PMK.PartitionDataType(PMK.HyperplaneType{T}(), emptyX(), X_parts_inds[jc], jc))
            else
                kid = side(parent, PMK.PartitionDataType(hps_by_depth[depth+2][jc], emptyX(), Int[], 0))
                grow!(kid, depth + 1, jc)
            end
        end
    end
    grow!(root, 0, 1)
    return root, [X[i] for i in X_parts_inds], X_parts_inds
end

# ---- organizetrainingsets(root, levels, X0, ε) on the device  (src/patchwork/partition.jl:301-357; findεpartitions! :269-298) ----
# Same return values as the reference: (X_set, X_set_inds, regions_list_set, problematic_inds); the per-point DFS with the
# two comparisons v.x < c + ε / v.x > c - ε runs on the GPU (un-fused, reference order), the index lists come back ascending.
function organizetrainingsets(root, levels::Integer, X0::Vector{Vector{T}}, ε::T; h::Handle = sharedhandle()) where T
    hps = PMK.fetchhyperplanes(root)
    D = length(X0[1]); N = length(X0)
    hv = Matrix{Float64}(undef, D, length(hps)); hc = Vector{Float64}(undef, length(hps))
    for (i, hp) in enumerate(hps)
        hv[:, i] = hp.v; hc[i] = hp.c
    end
    GC.@preserve hv hc check(h, ccall((:pmk_set_tree, libpmk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}), h.ptr, D, levels, hv, hc))
    n_leaves = length(hps) + 1
    Xm = pack(X0); leaf_off = Vector{Int64}(undef, n_leaves + 1); total = Ref{Int64}(0)
    GC.@preserve Xm leaf_off check(h, ccall((:pmk_organize_training_sets, libpmk), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Ptr{Int64}, Ref{Int64}), h.ptr, N, Xm, Float64(ε), leaf_off, total))
    inds = Vector{Int32}(undef, total[]); poff = Vector{Int64}(undef, N + 1); pleaves = Vector{Int32}(undef, total[])
    GC.@preserve inds poff pleaves check(h, ccall((:pmk_organize_fetch, libpmk), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}), h.ptr, inds, poff, pleaves))
    X_set_inds = [Int.(inds[leaf_off[r]+1:leaf_off[r+1]]) for r in 1:n_leaves]
    X_set = [X0[i] for i in X_set_inds]
    regions_list_set = [Int.(pleaves[poff[n]+1:poff[n+1]]) for n in 1:N]
    return X_set, X_set_inds, regions_list_set, Vector{Vector{Int}}(undef, 0)
end

# ---- checkpoint (no counterpart in the reference: MixtureGPType lives in memory only) -------------------------------------
"savemixtureGP(η, path, levels): X_parts, c_set, L_set, kernel, σ² and the tree in one file (pmk_save_model)"
function savemixtureGP(η::MixtureGPType, path::AbstractString, levels::Integer)
    η.multi === nothing || error("savemixtureGP of a model sharded over several GPUs: a model file holds a whole model")
    settree!(η, levels)
    check(η.h, ccall((:pmk_save_model, libpmk), Cint, (Ptr{Cvoid}, Cstring), η.h.ptr, path))
end

"loadmixtureGP(path; device) -> η with X_parts and hps read back from the file (pmk_load_model, pmk_get_X, pmk_get_tree)"
function loadmixtureGP(path::AbstractString; device = 0)
    h = Handle(device)
    check(h, ccall((:pmk_load_model, libpmk), Cint, (Ptr{Cvoid}, Cstring), h.ptr, path))
    D = Ref{Cint}(0); nl = Ref{Int64}(0); kid = Ref{Cint}(0); kp = Ref{Float64}(0); s2 = Ref{Float64}(0); lv = Ref{Cint}(0)
    check(h, ccall((:pmk_model_info, libpmk), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Int64}, Ref{Cint}, Ref{Float64}, Ref{Float64}, Ref{Cint}),
                   h.ptr, D, nl, kid, kp, s2, lv))
    X_parts = Vector{Vector{Vector{Float64}}}(undef, nl[])
    for p in 1:nl[]
        n = Ref{Int64}(0)
        check(h, ccall((:pmk_leaf_size, libpmk), Cint, (Ptr{Cvoid}, Int64, Ref{Int64}), h.ptr, p, n))
        Xm = Matrix{Float64}(undef, D[], n[])
        GC.@preserve Xm check(h, ccall((:pmk_get_X, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), h.ptr, p, Xm))
        X_parts[p] = [Xm[:, i] for i in 1:n[]]
    end
    n_hp = lv[] > 1 ? (1 << (lv[] - 1)) - 1 : 0
    hv = Matrix{Float64}(undef, D[], n_hp); hc = Vector{Float64}(undef, n_hp)
    n_hp > 0 && GC.@preserve hv hc check(h, ccall((:pmk_get_tree, libpmk), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, hv, hc))
    hps = [PMK.HyperplaneType{Float64}(hv[:, i], hc[i]) for i in 1:n_hp]
    θ = (PMK.GaussianKernel1DType, PMK.Spline34KernelType, nothing, nothing, PMK.BrownianBridge1ϵ, PMK.BrownianBridge2ϵ,
         PMK.Spline12KernelType, PMK.Spline32KernelType, PMK.RationalQuadraticKernelType)[kid[]+1]
    θv = kid[] == 2 ? PMK.BrownianBridge10(kp[]) : kid[] == 3 ? PMK.BrownianBridge20(kp[]) : θ(kp[])      # declarations.jl:25-100
    return MixtureGPType{Float64}(X_parts, hps, h, nothing, fill(s2[], nl[]), true, θv, UInt(0)), Int(lv[])
end

end # module
