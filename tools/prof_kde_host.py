"""Host-side profile of kde!(points) LOOCV on 100k x 4-D points (where does the non-GPU time go?)."""
import cProfile, pstats, sys, os, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from kde_b200 import api
from tests.util import mixture
K.init(0)
rng = np.random.default_rng(3)
pts = mixture(rng, 4, 100000)
K.kde(pts[:, :1000])
acc = {"destroy": 0.0, "n": 0}
orig = api.BallTreeDensity._invalidate
def timed(self):
    had = self._handle is not None
    t0 = time.perf_counter(); orig(self); dt = time.perf_counter() - t0
    if had:
        acc["destroy"] += dt; acc["n"] += 1
api.BallTreeDensity._invalidate = timed
t0 = time.perf_counter(); p = K.kde(pts); wall = time.perf_counter() - t0
print("kde!(100k x 4) wall %.3f s; %d device trees destroyed in %.1f ms total" % (wall, acc["n"], acc["destroy"] * 1e3))
pr = cProfile.Profile(); pr.enable(); p = K.kde(pts); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(10); print(s.getvalue()[:2500])
