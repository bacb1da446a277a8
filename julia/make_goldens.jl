# make_goldens.jl -- ONE run of the unmodified reference turns the Gibbs half of the oracle from
# "parity unpinned" into "pinned".
#
# The reference's tests hold no label / point goldens for prodAppxMSGibbsS (only statistical bands), and
# there is no julia binary in the build image.  prodAppxMSGibbsS accepts its random streams as keyword
# arguments (src/MSGibbs01.jl:661-662), so on any machine with Julia + KernelDensityEstimate.jl:
#
#     julia --project=<KernelDensityEstimate.jl checkout> julia/make_goldens.jl tests/golden/julia
#
# writes, per case, the input point sets, bandwidths, the injected randU / randN and the reference's
# points / indices as full-precision text.  tests/test_julia_goldens.py picks the files up: the oracle
# (CPU suite) and the CUDA path (gpu suite) must then reproduce the labels exactly and the points to
# 1e-10; while the directory is empty those tests skip with the reason "parity unpinned".
using KernelDensityEstimate, DelimitedFiles, Random

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia")
mkpath(outdir)
Random.seed!(20261017)

function emit(name::String, ptsets::Vector{Matrix{Float64}}, bws::Vector{Vector{Float64}}, Np::Int, Niter::Int;
              mask=nothing, addEntropy::Bool=true)
  P = BallTreeDensity[kde!(ptsets[j], bws[j]) for j in eachindex(ptsets)]
  d, M = size(ptsets[1], 1), length(P)
  maxNp = maximum([Np; Npts.(P)])
  Nlevels = floor(Int, (log(Float64(maxNp)) / log(2.0)) + 1.0)
  randU = rand(Int(Np * M * (Niter + 2) * Nlevels))
  randN = randn(Int(d * Np * (Nlevels + 1)))
  dummy = kde!(rand(d, Np), [1.0])
  kw = mask === nothing ? (;) : (; partialDimMask=mask)
  pts, ind = prodAppxMSGibbsS(dummy, P, nothing, nothing; Niter=Niter, randU=copy(randU), randN=copy(randN),
                              addEntropy=addEntropy, kw...)
  open(joinpath(outdir, name * "_meta.txt"), "w") do io
    println(io, "d $d\nM $M\nNp $Np\nNiter $Niter\naddEntropy $(Int(addEntropy))\nmasked $(Int(mask !== nothing))")
  end
  for j in 1:M
    writedlm(joinpath(outdir, "$(name)_pts$(j).txt"), ptsets[j])          # d rows x N columns
    writedlm(joinpath(outdir, "$(name)_bw$(j).txt"), bws[j])
    mask === nothing || writedlm(joinpath(outdir, "$(name)_mask$(j).txt"), Int.(mask[j]))
  end
  writedlm(joinpath(outdir, name * "_randU.txt"), randU)
  writedlm(joinpath(outdir, name * "_randN.txt"), randN)
  writedlm(joinpath(outdir, name * "_points.txt"), pts)                  # d x Np
  writedlm(joinpath(outdir, name * "_indices.txt"), ind)                 # M x Np (= permutation + 1)
  println("wrote $name: d=$d M=$M Np=$Np Niter=$Niter Nlevels=$Nlevels")
end

# C1: the README product
emit("c1_readme", [randn(2, 100), 2.0 .+ randn(2, 100)], [[0.35, 0.35], [0.35, 0.35]], 100, 5)
# testProds default shape (D=3, M=6, N=100) with distinct bandwidths per density
emit("d3_m6", [randn(3, 100) for _ in 1:6], [[0.30 + 0.02j, 0.35, 0.40 - 0.01j] for j in 1:6], 100, 5)
# C2 shape: 1-D, 300 and 100 components (level lists of different depth), Np > max N
emit("c2_mixed", [rand(1, 300) .^ 2.0, 0.5 .* sqrt.(-2.0 .* log.(rand(1, 100))) .- 0.5], [[0.05], [0.08]], 400, 5)
# partial-dimension masks (test/testPartialProd.jl), default Niter = 3
let p1 = rand(2, 100) .+ 10.0, p2 = rand(2, 100), p3 = rand(2, 100) .- 10.0
  p1[2, :] .= 9999999.0; p3[1, :] .= 9999999.0
  emit("partial_mask", [p1, p2, p3], [[0.1, 0.1], [0.1, 0.1], [0.1, 0.1]], 100, 3;
       mask=[BitVector([true, false]), BitVector([true, true]), BitVector([false, true])])
end
# addEntropy = false (examples/ExtractingLabels.jl flavour) and the M = 16 @simd-sum edge
emit("noentropy", [randn(2, 50) .+ 0.5j for j in 0:2], [[0.5, 0.7] for _ in 1:3], 64, 5; addEntropy=false)
emit("m16_simd", [randn(2, 24) .+ 0.1j for j in 0:15], [[0.4 + 0.031j, 0.5 + 0.017j] for j in 0:15], 200, 2)
