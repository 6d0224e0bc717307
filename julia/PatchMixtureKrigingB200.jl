# PatchMixtureKrigingB200.jl -- Julia host layer over libpmk_b200.so (C ABI: include/pmk.h).
#
# Drop-in for the fit / query hot path of PatchMixtureKriging.jl: the functions below keep the
# reference's names, argument order and return values (reference file:line in each docstring) and
# replace only the BODIES with `ccall`s.  Everything the north star keeps on the host stays the
# reference's own Julia code: `setuppartition`, `organizetrainingsets`, `fetchhyperplanes`,
# `findpartition`, `findneighbourpartitions`, the kernel parameter structs and `array2matrix`
# (an optional `setuppartition` with the per-level O(N) work on the GPU is at the end of this file).
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

# ---- MixtureGPType: the fitted state lives in HBM behind the handle -----------------------------------
mutable struct MixtureGPType{T}
    X_parts::Vector{Vector{Vector{T}}}
    hps::Vector{PMK.HyperplaneType{T}}
    h::Handle
    σ²_set::Vector{T}
    fitted::Bool
end

"MixtureGPType(X_set, hps)  (src/RKHS/mixtureGP.jl:54-66)"
MixtureGPType(X_parts::Vector{Vector{Vector{T}}}, hps::Vector{PMK.HyperplaneType{T}}; device = 0) where T =
    MixtureGPType{T}(X_parts, hps, Handle(device), T[], false)

"c_set[n], L_set[n], U_set[n] of the reference's struct (mixtureGP.jl:40-46), fetched on demand (1-based leaf)"
function c_set(η::MixtureGPType, n::Integer)
    out = Vector{Float64}(undef, length(η.X_parts[n]))
    GC.@preserve out check(η.h, ccall((:pmk_get_alpha, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), η.h.ptr, n, out))
    return out
end
function L_set(η::MixtureGPType, n::Integer)
    m = length(η.X_parts[n]); out = Matrix{Float64}(undef, m, m)
    GC.@preserve out check(η.h, ccall((:pmk_get_L, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), η.h.ptr, n, out))
    return LowerTriangular(out)
end
function U_set(η::MixtureGPType, n::Integer)      # Gram WITHOUT σ² (mixtureGP.jl:99)
    m = length(η.X_parts[n]); out = Matrix{Float64}(undef, m, m)
    GC.@preserve out check(η.h, ccall((:pmk_get_K, libpmk), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), η.h.ptr, n, out))
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
    rc = GC.@preserve leaf_off Xp yp kp ccall((:pmk_fit, libpmk), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64, Ref{Int64}, Ref{Cint}),
        η.h.ptr, size(Xp, 1), N_parts, leaf_off, Xp, yp, kernelid(θ), kp, length(kp), Float64(σ²), bad, info)
    rc == PMK_ERR_NOT_POSDEF && throw(PosDefException(info[]))      # cholesky(U) at mixtureGP.jl:109
    check(η.h, rc)
    η.σ²_set = fill(T(σ²), N_parts); η.fitted = true
    return η
end

"""How `L \\ kq` of `queryinner!` (mixtureGP.jl:311) is carried out: 0 = explicit inverse formed once per fit (default),
1 = blocked forward substitution (closest to `dtrsv`), 2 = explicit inverse with the round-1 column-sweep kernel.
`PMK_OPT_QUERY_SOLVER` of include/pmk.h."""
function setsolver!(η::MixtureGPType, solver::Integer)
    check(η.h, ccall((:pmk_set_option, libpmk), Cint, (Ptr{Cvoid}, Cint, Int64), η.h.ptr, 2, solver))
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
    GC.@preserve hv hc check(η.h, ccall((:pmk_set_tree, libpmk), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
                                        η.h.ptr, D, levels, hv, hc))
end

"""
querymixtureGP!(Yq, Vq, Xq, η, root, levels, radius, δ, θ, σ², weight_θ, debug_vars; debug_flag=false)::Nothing
(src/RKHS/mixtureGP.jl:159-294).  `root` is accepted for signature compatibility; the device works on η.hps.
"""
function querymixtureGP!(Yq::Vector{T}, Vq::Vector{T}, Xq::Vector{Vector{T}}, η::MixtureGPType{T}, root, levels,
                         radius::T, δ::T, θ, σ², weight_θ, debug_vars = nothing; debug_flag = false)::Nothing where T
    Nq = length(Xq); resize!(Yq, Nq); resize!(Vq, Nq)
    settree!(η, levels)
    Xm = pack(Xq); wp = kernelparams(weight_θ)
    GC.@preserve Xm wp Yq Vq check(η.h, ccall((:pmk_query, libpmk), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
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
    # hps_keep_flags_set / ts_set / zs_set are O(Nq * length(hps)) in the reference; the kept hyperplane ids
    # and their t are in `hp` / `t` (see include/pmk.h pmk_last_query_debug) if a caller needs them.
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

# ---- checkpoint (no counterpart in the reference: MixtureGPType lives in memory only) -------------------------------------
"savemixtureGP(η, path, levels): X_parts, c_set, L_set, kernel, σ² and the tree in one file (pmk_save_model)"
function savemixtureGP(η::MixtureGPType, path::AbstractString, levels::Integer)
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
    return MixtureGPType{Float64}(X_parts, hps, h, fill(s2[], nl[]), true), Int(lv[])
end

end # module
