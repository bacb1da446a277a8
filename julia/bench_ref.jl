# bench_ref.jl -- times the REAL reference package on BASELINE config 4's shape when a julia binary and the
# package are available (they are not in the build image; bench.py --impl reference then times the C port).
#   julia julia/bench_ref.jl [nsamples]
using KernelDensityEstimate, Random, Statistics
nsamp = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 256
Random.seed!(20261017)
corners = [-2.0 -2 -2; -2 -2 2; -2 2 -2; -2 2 2]'
function synth(j)
  comp = rand(1:4, 4096)
  pts = corners[:, comp] .+ 0.6 .* randn(3, 4096)
  pts[1, :] .+= 0.25 * j
  h = vec(std(pts, dims=2)) .* (4.0 / (5.0 * 4096))^(1 / 7)
  kde!(pts, h)
end
trees = [synth(j) for j in 0:7]
dummy = kde!(rand(3, nsamp), [1.0])
prodAppxMSGibbsS(dummy, trees, nothing, nothing, Niter=5)          # compile
t = @elapsed prodAppxMSGibbsS(dummy, trees, nothing, nothing, Niter=5)
println("{\"impl\": \"reference-julia\", \"samples\": $nsamp, \"samples_per_s\": $(nsamp / t), \"threads\": 1}")
