// gibbs.cu -- host side of K1: schedule construction, launch, Philox stream materialisation.
// The kernel lives in gibbs_kernel.cuh and is instantiated once per dimension in gibbs_d<N>.cu.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <vector>

#include "gibbs_kernel.cuh"

namespace kdeb200 {

// gibbs_f32.cu
int gibbs_precision();
void gibbs32_drop_schedules(int slot);
int gibbs32_launch(const kdeb200_tree_t *trees, int ndens, int L, int Niter, bool masked, GibbsParams &P, cudaStream_t st);

extern template cudaError_t launch_gibbs_d<1>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<2>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<3>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<4>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<5>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<6>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<7>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
extern template cudaError_t launch_gibbs_d<8>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);

extern template cudaError_t launch_gibbs_warp_d<1>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<2>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<3>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<4>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<5>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<6>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<7>(const GibbsParams &, bool, int, cudaStream_t);
extern template cudaError_t launch_gibbs_warp_d<8>(const GibbsParams &, bool, int, cudaStream_t);

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

static void build_schedule(const kdeb200_tree_t *trees, int M, int L, int T, bool masked, bool literal, Schedule &S) {
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
          dr.variant = (pass == 0 && !masked && !literal) ? VAR_B : VAR_C;  // B records hold -0.5/b, not b
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

// ---- schedule cache -----------------------------------------------------------------------------------------
// The static schedule (draws + tile list) depends only on the tree set, Niter and the mask class.  It used to be rebuilt,
// allocated (3 x cudaMallocAsync), uploaded and SYNCHRONISED on every call -- irrelevant at 1M samples, but it was the
// difference between a 0.34 ms kernel and a 0.39 ms call at the small configurations, and it made kdeb200_gibbs_device
// block.  Entries are keyed by (context slot, tree handles, Niter, masked) and invalidated wholesale whenever any tree
// is destroyed (a global epoch); each context keeps the 8 most recent.  Batch counters come from a ring of 256 slots
// per entry, zeroed once at creation and reset by the LAST CTA of each launch.
std::atomic<uint64_t> g_tree_epoch{1};
void gibbs_invalidate_schedules() { g_tree_epoch.fetch_add(1); }

namespace {
struct SchedEntry {
  uint64_t epoch = 0;
  int slot = 0, M = 0, T = 0;
  bool masked = false, literal = false;
  kdeb200_tree_t trees[KDEB200_MAX_DENS] = {nullptr};
  char *d_base = nullptr;  // draws | tiles | counters, one allocation on the context's device
  Draw *d_draws = nullptr;
  TileDesc *d_tiles = nullptr;
  int *d_counters = nullptr;
  int ndraws = 0, ntiles = 0;
  unsigned next_counter = 0;
};
constexpr int SCHED_COUNTERS = 256;
constexpr size_t SCHED_KEEP = 8;
std::mutex g_sched_mu;
std::list<SchedEntry> g_sched[KDEB200_MAX_GPUS];

void sched_free(SchedEntry &e, cudaStream_t st) {
  if (e.d_base) cudaFreeAsync(e.d_base, st);
  e.d_base = nullptr;
}
}  // namespace

// drops every cached schedule of a context (its device memory is about to go away)
void gibbs_drop_schedules(int slot) {
  gibbs32_drop_schedules(slot);
  std::lock_guard<std::mutex> lk(g_sched_mu);
  for (auto &e : g_sched[slot]) sched_free(e, ctx_at(slot).stream);
  g_sched[slot].clear();
}

static int sched_get(const kdeb200_tree_t *trees, int M, int L, int T, bool masked, bool literal, cudaStream_t st,
                     SchedEntry *out, int **counter) {
  Context &c = ctx();
  std::lock_guard<std::mutex> lk(g_sched_mu);
  auto &lst = g_sched[c.slot];
  const uint64_t epoch = g_tree_epoch.load();
  for (auto it = lst.begin(); it != lst.end();) {  // entries from before the last tree_destroy may point at freed records
    if (it->epoch != epoch) {
      sched_free(*it, c.stream);
      it = lst.erase(it);
    } else {
      ++it;
    }
  }
  for (auto it = lst.begin(); it != lst.end(); ++it) {
    if (it->M != M || it->T != T || it->masked != masked || it->literal != literal) continue;
    bool same = true;
    for (int j = 0; j < M; ++j) same = same && it->trees[j] == trees[j];
    if (!same) continue;
    lst.splice(lst.begin(), lst, it);  // most recently used first
    *counter = it->d_counters + (it->next_counter++ % SCHED_COUNTERS);
    *out = *it;
    return 0;
  }
  Schedule S;
  build_schedule(trees, M, L, T, masked, literal, S);
  SchedEntry e;
  e.epoch = epoch;
  e.slot = c.slot;
  e.M = M;
  e.T = T;
  e.masked = masked;
  e.literal = literal;
  for (int j = 0; j < M; ++j) e.trees[j] = trees[j];
  e.ndraws = (int)S.draws.size();
  e.ntiles = (int)S.tiles.size();
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_draws = up(sizeof(Draw) * S.draws.size()), b_tiles = up(sizeof(TileDesc) * S.tiles.size()),
               b_cnt = up(sizeof(int) * SCHED_COUNTERS);
  // the upload runs on the library stream of this context and is complete before the entry is published; the caller's
  // stream only ever sees finished buffers
  KDE_CUDA(cudaMallocAsync(&e.d_base, b_draws + b_tiles + b_cnt, c.stream));
  e.d_draws = reinterpret_cast<Draw *>(e.d_base);
  e.d_tiles = reinterpret_cast<TileDesc *>(e.d_base + b_draws);
  e.d_counters = reinterpret_cast<int *>(e.d_base + b_draws + b_tiles);
  cudaError_t err = cudaMemcpyAsync(e.d_draws, S.draws.data(), sizeof(Draw) * S.draws.size(), cudaMemcpyHostToDevice, c.stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(e.d_tiles, S.tiles.data(), sizeof(TileDesc) * S.tiles.size(), cudaMemcpyHostToDevice, c.stream);
  if (err == cudaSuccess) err = cudaMemsetAsync(e.d_counters, 0, b_cnt, c.stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(c.stream);  // S's vectors are pageable host memory
  if (err != cudaSuccess) {
    cudaFreeAsync(e.d_base, c.stream);
    KDE_FAIL(100 + (int)err, "gibbs: uploading the schedule: %s", cudaGetErrorString(err));
  }
  (void)st;
  lst.push_front(e);
  while (lst.size() > SCHED_KEEP) {
    // an evicted entry may still be read by a kernel in flight on a caller stream: release it only after the device
    // has drained (rare: more than 8 distinct tree sets alternating)
    cudaDeviceSynchronize();
    sched_free(lst.back(), c.stream);
    lst.pop_back();
  }
  *counter = lst.front().d_counters + (lst.front().next_counter++ % SCHED_COUNTERS);
  *out = lst.front();
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
  // Densities the restructured arithmetic cannot take (zero / non-finite / out-of-range variances, non-finite means or
  // weights) run through the warp-per-chain kernel with the reference's arithmetic verbatim, NaN / Inf rules included
  // (src/MSGibbs01.jl:287-315) -- as long as their level lists fit that kernel's shared memory.
  bool literal = false;
  for (int j = 0; j < ndens; ++j) {
    if (trees[j]->degenerate) literal = true;
    if (trees[j]->slot != c.slot) KDE_FAIL(3, "gibbs: tree %d lives on another GPU of the set than the calling context", j);
  }

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
  SchedEntry E;
  int *d_counter = nullptr;
  if (literal) masked = true;  // the literal evaluation reads the activity flags in every mode
  if (int rc = sched_get(trees, ndens, L, Niter, masked, literal, st, &E, &d_counter)) return rc;

  P.draws = E.d_draws;
  P.tiles = E.d_tiles;
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
  P.ndraws = E.ndraws;
  P.ntiles = E.ntiles;
  P.M = ndens;
  P.L = L;
  P.T = Niter;
  P.add_entropy = add_entropy ? 1 : 0;
  P.literal = literal ? 1 : 0;
  P.nbatches = (int)((s1 - s0 + GB_THREADS - 1) / GB_THREADS);
  for (int j = 0; j < ndens; ++j) {
    const kdeb200_tree_s *t = trees[j];
    P.root_rec[j] = t->d_buf + t->levels[0].offC;
    P.labels[j] = t->d_labels;
    for (int k = 0; k < d; ++k) P.hvar[j][k] = t->hvar[k];
  }
  // Few chains: one warp per chain (K1w).  The thread-per-chain kernel needs ~128 x 4 x SMs chains to fill the chip and
  // costs the same 0.3 - 3 ms whether it runs 100 chains or 10 000; measured crossover (tools/gibbs_kernel_crossover.py,
  // profiles/r02_gibbs_crossover.json): ~4 000 chains for 2 - 6 densities of 100 - 1000 components, so the warp
  // kernel takes calls of up to 24 chains per SM (KDEB200_GIBBS_WARP_MAX overrides; 0 disables it).  Its prefix sums
  // live in shared memory, which bounds the level size.
  {
    int nmax = 1;
    for (int j = 0; j < ndens; ++j)
      if (trees[j]->levels[trees[j]->depth].n > nmax) nmax = (int)trees[j]->levels[trees[j]->depth].n;
    int64_t warp_max = (int64_t)24 * c.sm_count;
    if (const char *ev = getenv("KDEB200_GIBBS_WARP_MAX")) warp_max = atoll(ev);
    if (literal && gibbs_warp_smem(d, nmax) > 96 * 1024)
      KDE_FAIL(8, "gibbs: a density has a zero, non-finite or out-of-range (variance outside [1e-30, 1e30]) bandwidth / mean / "
                  "weight and more than ~2900 components: the verbatim-arithmetic kernel cannot hold its level lists "
                  "(no CPU fallback)");
    if ((literal || s1 - s0 <= warp_max) && gibbs_warp_smem(d, nmax) <= 96 * 1024) {
      cudaError_t we = cudaErrorInvalidValue;
      switch (d) {
#ifdef GB_ONLY_D3
        case 3: we = launch_gibbs_warp_d<3>(P, masked, nmax, st); break;
#else
        case 1: we = launch_gibbs_warp_d<1>(P, masked, nmax, st); break;
        case 2: we = launch_gibbs_warp_d<2>(P, masked, nmax, st); break;
        case 3: we = launch_gibbs_warp_d<3>(P, masked, nmax, st); break;
        case 4: we = launch_gibbs_warp_d<4>(P, masked, nmax, st); break;
        case 5: we = launch_gibbs_warp_d<5>(P, masked, nmax, st); break;
        case 6: we = launch_gibbs_warp_d<6>(P, masked, nmax, st); break;
        case 7: we = launch_gibbs_warp_d<7>(P, masked, nmax, st); break;
        case 8: we = launch_gibbs_warp_d<8>(P, masked, nmax, st); break;
#endif
      }
      if (we != cudaSuccess) KDE_FAIL(100 + (int)we, "gibbs (warp per chain) kernel launch: %s", cudaGetErrorString(we));
      if (launches) *launches += 1;
      return 0;
    }
  }
  if (gibbs_precision() == KDEB200_F32) {  // K1f: FP32 label probabilities (statistical mode only), gibbs_f32.cu
    if (int rc = gibbs32_launch(trees, ndens, L, Niter, masked, P, st)) return rc;
    if (launches) *launches += 1;
    return 0;
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
