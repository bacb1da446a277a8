# KDEB200.jl -- Julia `ccall` binding of libkdeb200.so (include/kdeb200.h).
#
# NOT RUNNABLE IN THE BUILD IMAGE (no julia binary there); it is the binding a maintainer of
# KernelDensityEstimate.jl adds to re-point the three seams of SURVEY.md 8b at the B200 library.
# kerneldensityestimate.jl_b200/api.py is the line-for-line Python mirror that the test-suite
# exercises; keep the two in sync.
#
#   seam S1  gibbs1(...)                 src/MSGibbs01.jl:527-629  ->  KDEB200.gibbs1!
#   seam S2  evaluate(bd, loc, p, ...)   src/DualTree01.jl:303-346 ->  KDEB200.evaluate!
#   seam S3  entropy(bd) / nLOO_LL       src/DualTree01.jl:505-508, src/CrossValidation.jl:15-24
#                                                                    ->  KDEB200.entropy
module KDEB200

using KernelDensityEstimate
const KDE = KernelDensityEstimate

const LIB = get(ENV, "KDEB200_LIB", "libkdeb200.so")

# nonzero return code -> the reference's error convention (error(...) => ErrorException)
function check(rc::Cint)
  rc == 0 && return nothing
  error(unsafe_string(ccall((:kdeb200_last_error, LIB), Cstring, ())))
end

init(device::Integer=0) = check(ccall((:kdeb200_init, LIB), Cint, (Cint,), device))

# One Julia process drives `ngpus` GPUs (0 = all visible): afterwards gibbs1!, evaluate!, entropy and lcv_bandwidths
# shard their samples / query points / leaf rows over the set inside the library -- no MPI, no second process.
init_multi(ngpus::Integer=0) = check(ccall((:kdeb200_init_multi, LIB), Cint, (Cint,), ngpus))
function multi_count()
  n = Ref{Cint}(0)
  check(ccall((:kdeb200_multi_count, LIB), Cint, (Ref{Cint},), n))
  Int(n[])
end

# setForceEvalDirect!(flag) of the reference (src/DualTree01.jl:3-9): false selects the error-bounded pruned kernel
# (every value within 1e-13 of the brute-force sum; the reference's dual tree: errTol = 1e-3) for evaluate!
setForceEvalDirect!(flag::Bool) = check(ccall((:kdeb200_set_pruning, LIB), Cint, (Cint,), flag ? 1 : 2))

# Arithmetic of the Gibbs label probabilities: 0 = FP64 (default; labels identical to the reference under injected
# randU / randN), 1 = packed FP32 on large calls (free-running statistical mode only, ~3x the sample rate)
set_gibbs_precision!(precision::Int) = check(ccall((:kdeb200_set_gibbs_precision, LIB), Cint, (Cint,), precision))
function gibbs_f32_slow_draws()
  n = Ref{Culonglong}(0)
  check(ccall((:kdeb200_gibbs_f32_slow_draws, LIB), Cint, (Ref{Culonglong},), n))
  return Int(n[])
end

# ---- S0: device-resident BallTreeDensity -------------------------------------------------
mutable struct DeviceTree
  h::Ptr{Cvoid}
  # gibbs=false: leaf records only (evaluate / entropy / nLOO_LL never touch the level records)
  function DeviceTree(bd::BallTreeDensity; gibbs::Bool=true)
    bd.multibandwidth == 0 || error("kdeb200: multibandwidth != 0 is not supported")
    bt = bd.bt
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bd begin
      if gibbs
        check(ccall((:kdeb200_tree_create, LIB), Cint,
                    (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
                     Ref{Ptr{Cvoid}}),
                    bt.dims, bt.num_points, bd.means, bd.bandwidth, bt.weights, bt.left_child, bt.right_child,
                    bt.permutation, out))
      else
        check(ccall((:kdeb200_tree_create_eval, LIB), Cint,
                    (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                    bt.dims, bt.num_points, bd.means, bd.bandwidth, bt.weights, bt.permutation, out))
      end
    end
    t = new(out[])
    finalizer(x -> ccall((:kdeb200_tree_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), t)
    return t
  end
end

euclidean(addop, diffop) = all(f -> f === (+), addop) && all(f -> f === (-), diffop)

# ---- S1: gibbs1 -------------------------------------------------------------------------
# Same positional arguments as KDE.gibbs1; writes pts (d*Np) and ind (Ndens x Np) in place.
# randU / randN === nothing selects the library's Philox streams (keyed by `seed`).
function gibbs1!(Ndens::Int, trees::Vector{BallTreeDensity}, Np::Int, Niter::Int,
                 pts::Vector{Float64}, ind::Matrix{Int},
                 randU::Union{Nothing,Vector{Float64}}, randN::Union{Nothing,Vector{Float64}};
                 addop=(+,), diffop=(-,), getMu=(KDE.getEuclidMu,), getLambda=(KDE.getEuclidLambda,),
                 glbs=nothing, addEntropy::Bool=true, ndims::Int=maximum(Ndim.(trees)),
                 partialDimMask::AbstractVector{<:BitVector}=[trues(ndims) for i in 1:Ndens],
                 seed::UInt64=rand(UInt64))
  (euclidean(addop, diffop) && all(f -> f === KDE.getEuclidMu, getMu) &&
   all(f -> f === KDE.getEuclidLambda, getLambda)) ||
    error("kdeb200: only the default Euclidean (+,-) manifold is supported on the B200 path")
  dts = [DeviceTree(t) for t in trees]
  hs = Ptr{Cvoid}[d.h for d in dts]
  mask = UInt8[partialDimMask[j][k] ? 0x01 : 0x00 for k in 1:ndims, j in 1:Ndens]  # [j*d + k], column-major
  uptr = randU === nothing ? Ptr{Float64}(C_NULL) : pointer(randU)
  nptr = randN === nothing ? Ptr{Float64}(C_NULL) : pointer(randN)
  # glbs.recordChoosen (src/MSGibbs01.jl:29-31, examples/ExtractingLabels.jl): the library records
  # permutation[ind[j]] of the last sampleIndex call of every level into an Int64 array [Nlevels, Ndens, Np]
  # (C order [sample][density][level]); it is unpacked into glbs.labelsChoosen[sample][density][level] below.
  record = glbs !== nothing && glbs.recordChoosen
  nlev = Ref{Cint}(0)
  GC.@preserve dts hs check(ccall((:kdeb200_gibbs_sizes, LIB), Cint,
    (Ptr{Ptr{Cvoid}}, Cint, Cint, Ref{Cint}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
    hs, Ndens, Niter, nlev, C_NULL, C_NULL, C_NULL))
  rec = record ? Array{Int64,3}(undef, Int(nlev[]), Ndens, Np) : Array{Int64,3}(undef, 0, 0, 0)
  GC.@preserve dts hs mask randU randN pts ind rec begin
    check(ccall((:kdeb200_gibbs, LIB), Cint,
                (Ptr{Ptr{Cvoid}}, Cint, Int64, Cint, Cint, Ptr{UInt8}, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
                 UInt64, Int64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                hs, Ndens, Np, Niter, addEntropy, mask, uptr, randU === nothing ? 0 : length(randU),
                nptr, randN === nothing ? 0 : length(randN), seed, 0, Np, pts, ind,
                record ? pointer(rec) : Ptr{Int64}(C_NULL)))
  end
  if record   # same nesting as the reference fills it at :109-112 (entries of levels never visited are absent)
    for s in 1:Np
      glbs.labelsChoosen[s] = Dict{Int,Dict{Int,Int}}()
      for j in 1:Ndens
        glbs.labelsChoosen[s][j] = Dict{Int,Int}()
        for l in 1:Int(nlev[])
          rec[l, j, s] >= 0 && (glbs.labelsChoosen[s][j][l] = rec[l, j, s])
        end
      end
    end
  end
  nothing
end

# ---- S2: evaluate -------------------------------------------------------------------------
# p is filled in the original order of `locations`' points (bd === locations => leave-one-out).
function evaluate!(bd::BallTreeDensity, locations::BallTreeDensity, p::Vector{Float64},
                   maxErr::Float64=1e-3, addop=(+,), diffop=(-,); precision::Int=0)
  bd.bt.dims == locations.bt.dims || error("evaluate -- dimensions of two BallTreeDensities must match")
  euclidean(addop, diffop) || error("kdeb200: only the default Euclidean (+,-) manifold is supported")
  dt = DeviceTree(bd; gibbs=false)
  if bd === locations
    GC.@preserve dt p check(ccall((:kdeb200_eval, LIB), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Cint, Ptr{Float64}), dt.h, C_NULL, Npts(bd), 1, precision, p))
  else
    pos = getPoints(locations)   # d x M, original order
    GC.@preserve dt pos p check(ccall((:kdeb200_eval, LIB), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Cint, Ptr{Float64}), dt.h, pos, size(pos, 2), 0, precision, p))
  end
  nothing
end

function evaluateDualTree(bd::BallTreeDensity, pos::Matrix{Float64}, lvFlag::Bool=false; precision::Int=0)
  bd.bt.dims == size(pos, 1) || error("bd and pos must have the same dimension")
  dt = DeviceTree(bd; gibbs=false)
  M = lvFlag ? Npts(bd) : size(pos, 2)
  p = zeros(M)
  GC.@preserve dt pos p check(ccall((:kdeb200_eval, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Cint, Ptr{Float64}), dt.h, pos, M, lvFlag, precision, p))
  return p
end

# ---- S3: leave-one-out entropy (one launch, one scalar back) ------------------------------------
function entropy(bd::BallTreeDensity, dt::DeviceTree=DeviceTree(bd; gibbs=false))
  H = Ref{Float64}(0.0)
  bw = bd.bandwidthMin[1:bd.bt.dims]     # the (possibly alpha^2-scaled) leaf variances
  GC.@preserve dt bw check(ccall((:kdeb200_loo_entropy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}),
                                 dt.h, bw, H))
  return H[]
end

# the bandwidth loop of kde!(points) (src/KDE01.jl:13-23) in one call; N <= 512: one kernel launch
function lcv_bandwidths(points::Matrix{Float64})
  d, N = size(points)
  bw = zeros(d)
  GC.@preserve points bw check(ccall((:kdeb200_kde_lcv, LIB), Cint,
    (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}), d, N, points, bw, C_NULL))
  return bw
end

# multi-GPU form (one Julia process per GPU): this process owns leaf rows j0+1:j1 of every nLOO_LL evaluation;
# `allreduce` is a C-callable that sums the partial likelihood and maxes the zero flag over the processes, e.g.
#   function ar(ps::Ptr{Float64}, pf::Ptr{Cint}, ::Ptr{Cvoid})::Cint
#     unsafe_store!(ps, MPI.Allreduce(unsafe_load(ps), +, comm)); unsafe_store!(pf, MPI.Allreduce(unsafe_load(pf), max, comm)); 0
#   end
#   lcv_bandwidths_sharded(points, j0, j1, @cfunction(ar, Cint, (Ptr{Float64}, Ptr{Cint}, Ptr{Cvoid})))
function lcv_bandwidths_sharded(points::Matrix{Float64}, j0::Integer, j1::Integer, allreduce::Ptr{Cvoid})
  d, N = size(points)
  bw = zeros(d)
  GC.@preserve points bw check(ccall((:kdeb200_kde_lcv_sharded, LIB), Cint,
    (Cint, Int64, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cint}),
    d, N, points, j0, j1, allreduce, C_NULL, bw, C_NULL))
  return bw
end

# `*` (src/MSGibbs01.jl:707-726) in one call: product samples + the LOOCV bandwidths of kde!(pGM); up to 512 samples never
# leave the device between the Gibbs kernel and the refit.  Returns the product density.
function product(trees::Vector{BallTreeDensity}; addEntropy::Bool=true, Niter::Int=5, seed::UInt64=rand(UInt64))
  Np = round(Int, sum(Npts.(trees)) / length(trees))
  d = Ndim(trees[1])
  dts = [DeviceTree(t) for t in trees]
  hs = Ptr{Cvoid}[t.h for t in dts]
  pts = zeros(d, Np); bw = zeros(d)
  GC.@preserve dts hs pts bw check(ccall((:kdeb200_product_kde, LIB), Cint,
    (Ptr{Ptr{Cvoid}}, Cint, Int64, Cint, Cint, Ptr{UInt8}, UInt64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Cint}),
    hs, length(trees), Np, Niter, addEntropy, C_NULL, seed, pts, C_NULL, bw, C_NULL))
  return kde!(pts, bw)
end

# sample(npd, Npts) (src/KDE01.jl:164-183) on the device; Philox(seed) variates (Julia's RNG streams can be injected as
# randU (Npts) / randN (d x Npts) through the C entry point directly)
function sample(npd::BallTreeDensity, Np::Int; seed::UInt64=rand(UInt64))
  dt = DeviceTree(npd; gibbs=false)
  pts = zeros(Ndim(npd), Np); ind = zeros(Int, Np)
  GC.@preserve dt pts ind check(ccall((:kdeb200_sample, LIB), Cint,
    (Ptr{Cvoid}, Int64, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), dt.h, Np, seed, C_NULL, C_NULL, pts, ind))
  return pts, ind
end

# every 1-D marginal on its own grid in one launch (getKDEMax, src/DualTree01.jl:558-569): grids is G x d (column k = the
# abscissae of dimension k, i.e. the C side's d x G row-major), result likewise
function eval_marginals(p::BallTreeDensity, grids::Matrix{Float64})
  dt = DeviceTree(p; gibbs=false)
  out = similar(grids)
  GC.@preserve dt grids out check(ccall((:kdeb200_eval_marginals, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), dt.h, grids, size(grids, 1), out))
  return out
end

# nLOO_LL with the device tree reused across the ~20 golden-section steps of one ksize call
function nLOO_LL(alpha::Float64, bd::BallTreeDensity, dt::DeviceTree)
  a2 = alpha^2
  KDE.updateBandwidth!(bd, bd.bandwidth * a2)
  H = entropy(bd, dt)
  KDE.updateBandwidth!(bd, bd.bandwidth / a2)
  return H
end

end # module
