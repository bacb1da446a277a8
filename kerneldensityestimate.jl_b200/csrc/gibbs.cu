// gibbs.cu -- host side of K1: schedule construction, launch, Philox stream materialisation.
// The kernel lives in gibbs_kernel.cuh and is instantiated once per dimension in gibbs_d<N>.cu.
#include <cmath>
#include <cstring>
#include <vector>

#include "gibbs_kernel.cuh"

namespace kdeb200 {

extern template cudaError_t launch_gibbs_d<1>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<2>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<3>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<4>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<5>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<6>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<7>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<8>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);

__global__ void philox_streams_kernel(uint64_t seed, int64_t Np, int64_t perU, int64_t perN, double *U, double *G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nu = Np * perU, nn = Np * perN;
  if (i < nu) {
    // array slot g (0-based) is read by sample s = (g+1) / perU at call c = (g+1) % perU
    const int64_t g1 = i + 1;
    U[i] = philox_uniform(seed, (uint64_t)(g1 / perU), (uint32_t)(g1 % perU));
  }
  if (i < nn) G[i] = philox_normal(seed, (uint64_t)(i / perN), (uint32_t)(i % perN));
}

// ---------------------------------------------------------------- host side --------------
int gibbs_nlevels(const kdeb200_tree_t *trees, int ndens) {
  int64_t maxNp = 0;
  for (int j = 0; j < ndens; ++j)
    if (maxNp < trees[j]->N) maxNp = trees[j]->N;
  return (int)std::floor((std::log((double)maxNp) / std::log(2.0)) + 1.0);  // src/MSGibbs01.jl:568
}

struct Schedule {
  std::vector<Draw> draws;
  std::vector<TileDesc> tiles;
  int64_t evals = 0;
};

static void build_schedule(const kdeb200_tree_t *trees, int M, int L, int T, bool masked, Schedule &S) {
  for (int l = 1; l <= L; ++l) {
    for (int pass = 0; pass <= T; ++pass) {
      for (int j = 0; j < M; ++j) {
        const kdeb200_tree_s *t = trees[j];
        const Level &lv = t->levels[l < t->depth ? l : t->depth];
        Draw dr;
        std::memset(&dr, 0, sizeof(dr));
        dr.j = (short)j;
        dr.kind = pass == 0 ? 0 : 1;
        dr.level = (short)l;
        dr.new_level = (pass == 0 && j == 0) ? 1 : 0;
        dr.n = (int)lv.n;
        dr.wts = t->d_buf + lv.offW;
        dr.levperm = t->d_levperm + lv.offP;
        if (lv.cls == 0) {
          dr.variant = VAR_A;
          dr.rec = t->d_buf + lv.offA;
          dr.stride = t->SA;
          dr.rec_state = dr.rec;
          dr.state_stride = t->SA;
          dr.state_has_bw = 0;
        } else {
          dr.variant = (pass == 0 && !masked) ? VAR_B : VAR_C;
          dr.rec = t->d_buf + (dr.variant == VAR_B ? lv.offB : lv.offC);
          dr.stride = t->SC;
          dr.rec_state = t->d_buf + lv.offC;
          dr.state_stride = t->SC;
          dr.state_has_bw = 1;
        }
        int G = 1;
        while ((int64_t)G * GB_MAXCK < lv.n) G *= 2;
        const int dg = gb_dg(t->d, dr.variant);
        if (G < dg && lv.n >= 2 * dg) G = dg;  // pass 1 checkpoints per whole double group
        dr.G = G;
        dr.nchunks = (int)((lv.n + G - 1) / G);
        int tn = 1;
        while (tn * 2 * dr.stride * 8 <= GB_TILE_BYTES) tn *= 2;
        dr.tnodes = tn;
        dr.ntiles = (int)((lv.n + tn - 1) / tn);
        dr.tile0 = (int)S.tiles.size();
        for (int q = 0; q < dr.ntiles; ++q) {
          const int64_t a = (int64_t)q * tn;
          const int64_t cnt = (lv.n - a < tn) ? (lv.n - a) : tn;
          TileDesc td;
          td.src = dr.rec + a * dr.stride;
          td.bytes = (uint32_t)(cnt * dr.stride * sizeof(double));
          td.pad = 0;
          S.tiles.push_back(td);
        }
        S.evals += lv.n;
        S.draws.push_back(dr);
      }
    }
  }
}

int gibbs_sizes(const kdeb200_tree_t *trees, int ndens, int Niter, int *nlevels, int64_t *perU, int64_t *perN,
                int64_t *evals) {
  if (ndens < 1 || ndens > KDEB200_MAX_DENS) KDE_FAIL(3, "gibbs: ndens=%d outside 1..%d", ndens, KDEB200_MAX_DENS);
  if (Niter < 0) KDE_FAIL(3, "gibbs: Niter must be >= 0");
  for (int j = 0; j < ndens; ++j) {
    if (!trees[j]) KDE_FAIL(3, "gibbs: tree %d is NULL", j);
    if (!trees[j]->gibbs_ready)
      KDE_FAIL(7, "gibbs: tree %d was created by kdeb200_tree_create_eval (leaf records only); use kdeb200_tree_create", j);
    if (trees[j]->d != trees[0]->d) KDE_FAIL(6, "kdes must have same dimension");  // src/MSGibbs01.jl:721
  }
  const int L = gibbs_nlevels(trees, ndens);
  if (nlevels) *nlevels = L;
  if (perU) *perU = (int64_t)ndens * (1 + (int64_t)L * (1 + Niter));
  if (perN) *perN = (int64_t)trees[0]->d * (L + 1);
  if (evals) {
    int64_t e = 0;
    for (int j = 0; j < ndens; ++j)
      for (int l = 1; l <= L; ++l) e += trees[j]->levels[l < trees[j]->depth ? l : trees[j]->depth].n;
    *evals = e * (1 + Niter);
  }
  return 0;
}

int gibbs_device(const kdeb200_tree_t *trees, int ndens, int64_t Np, int Niter, int add_entropy,
                 const uint8_t *dimmask, const double *d_randU, int64_t nU, const double *d_randN, int64_t nN,
                 uint64_t seed, int64_t s0, int64_t s1, double *d_points, int64_t *d_indices,
                 int64_t *d_level_labels, cudaStream_t st, int *launches) {
  Context &c = ctx();
  int L = 0;
  int64_t perU = 0, perN = 0;
  if (int rc = gibbs_sizes(trees, ndens, Niter, &L, &perU, &perN, nullptr)) return rc;
  const int d = trees[0]->d;
  if (Np < 0 || s0 < 0 || s1 > Np || s0 > s1) KDE_FAIL(3, "gibbs: bad sample range [%lld,%lld) of %lld", (long long)s0, (long long)s1, (long long)Np);
  if (s1 == s0) return 0;
  if ((d_randU == nullptr) != (d_randN == nullptr)) KDE_FAIL(3, "gibbs: randU and randN must be given together");
  if (d_randU) {  // the reference would throw BoundsError
    if (s1 * perU > nU + 1) KDE_FAIL(7, "gibbs: randU too short (%lld < %lld)", (long long)nU, (long long)(s1 * perU - 1));
    if (s1 * perN > nN) KDE_FAIL(7, "gibbs: randN too short (%lld < %lld)", (long long)nN, (long long)(s1 * perN));
  }
  for (int j = 0; j < ndens; ++j)
    if (trees[j]->degenerate)
      KDE_FAIL(8, "gibbs: density %d has a non-positive or non-finite bandwidth/mean; not supported on the GPU path (no CPU fallback)", j + 1);

  GibbsParams P;
  std::memset(&P, 0, sizeof(P));
  bool masked = (ndens == 1);  // a lone density has no "other" dims: handled by the mask logic
  for (int j = 0; j < ndens; ++j)
    for (int k = 0; k < d; ++k) {
      P.mask[j][k] = dimmask ? (dimmask[j * d + k] != 0) : 1;
      if (!P.mask[j][k]) masked = true;
    }
  for (int j = 0; j < ndens; ++j)
    for (int k = 0; k < d; ++k) {
      unsigned char o = 0;
      for (int i = 0; i < ndens; ++i)
        if (i != j && P.mask[i][k]) o = 1;
      P.other[j][k] = o;
    }
  Schedule S;
  build_schedule(trees, ndens, L, Niter, masked, S);

  Draw *d_draws = nullptr;
  TileDesc *d_tiles = nullptr;
  int *d_counter = nullptr;
  KDE_CUDA(cudaMallocAsync(&d_draws, sizeof(Draw) * S.draws.size(), st));
  KDE_CUDA(cudaMallocAsync(&d_tiles, sizeof(TileDesc) * S.tiles.size(), st));
  KDE_CUDA(cudaMallocAsync(&d_counter, 256, st));
  KDE_CUDA(cudaMemsetAsync(d_counter, 0, 256, st));
  KDE_CUDA(cudaMemcpyAsync(d_draws, S.draws.data(), sizeof(Draw) * S.draws.size(), cudaMemcpyHostToDevice, st));
  KDE_CUDA(cudaMemcpyAsync(d_tiles, S.tiles.data(), sizeof(TileDesc) * S.tiles.size(), cudaMemcpyHostToDevice, st));
  KDE_CUDA(cudaStreamSynchronize(st));  // S's vectors are pageable host memory

  P.draws = d_draws;
  P.tiles = d_tiles;
  P.counter = d_counter;
  P.exptab = c.d_exptab;
  P.ec = make_exp_consts();
  P.randU = d_randU;
  P.randN = d_randN;
  P.points = d_points;
  P.indices = d_indices;
  P.level_labels = d_level_labels;
  P.s0 = s0;
  P.s1 = s1;
  P.perU = perU;
  P.perN = perN;
  P.seed = seed;
  P.ndraws = (int)S.draws.size();
  P.ntiles = (int)S.tiles.size();
  P.M = ndens;
  P.L = L;
  P.T = Niter;
  P.add_entropy = add_entropy ? 1 : 0;
  P.nbatches = (int)((s1 - s0 + GB_THREADS - 1) / GB_THREADS);
  for (int j = 0; j < ndens; ++j) {
    const kdeb200_tree_s *t = trees[j];
    P.root_rec[j] = t->d_buf + t->levels[0].offC;
    P.labels[j] = t->d_labels;
    for (int k = 0; k < d; ++k) P.hvar[j][k] = t->hvar[k];
  }
  const size_t smem = GB_STAGES * GB_TILE_BYTES;
  cudaError_t e = cudaErrorInvalidValue;
  switch (d) {
#ifdef GB_ONLY_D3  // tuning builds: one instantiation
    case 3: e = launch_gibbs_d<3>(P, masked, P.nbatches, smem, st, c.sm_count); break;
#else
    case 1: e = launch_gibbs_d<1>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 2: e = launch_gibbs_d<2>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 3: e = launch_gibbs_d<3>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 4: e = launch_gibbs_d<4>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 5: e = launch_gibbs_d<5>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 6: e = launch_gibbs_d<6>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 7: e = launch_gibbs_d<7>(P, masked, P.nbatches, smem, st, c.sm_count); break;
    case 8: e = launch_gibbs_d<8>(P, masked, P.nbatches, smem, st, c.sm_count); break;
#endif
  }
  if (e != cudaSuccess) KDE_FAIL(100 + (int)e, "gibbs kernel launch: %s", cudaGetErrorString(e));
  if (launches) *launches += 1;
  KDE_CUDA(cudaFreeAsync(d_draws, st));
  KDE_CUDA(cudaFreeAsync(d_tiles, st));
  KDE_CUDA(cudaFreeAsync(d_counter, st));
  return 0;
}

int philox_streams_device(uint64_t seed, int64_t Np, int64_t perU, int64_t perN, double *d_U, double *d_G,
                          cudaStream_t st) {
  const int64_t n = (Np * perU > Np * perN) ? Np * perU : Np * perN;
  if (n <= 0) return 0;
  philox_streams_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, Np, perU, perN, d_U, d_G);
  KDE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kdeb200
