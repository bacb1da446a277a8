"""DFMA pipeline probe: achieved fraction of the FP64 peak vs (warps per SM sub-partition, ILP)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kde_b200 as K
from kde_b200 import _lib
K.init(0)
L = _lib.lib()
peak = 64 * 148 * 1.965e9
print("warps/SMSP  ILP  frac_of_nominal  clk_per_dependent_DFMA")
for threads, bps in [(32, 4), (64, 4), (128, 4), (128, 8), (256, 8)]:
    wps = threads // 32 * bps / 4.0
    for ilp in (1, 2, 4, 8, 16):
        r = C.c_double(0)
        _lib.check(L.kdeb200_dfma_probe(ilp, bps, threads, 40000, C.byref(r)))
        frac = r.value / peak
        # per SMSP: wps warps x ilp chains, each chain issues one DFMA per latency L: rate = wps*ilp/L warp-instr/clk, cap 0.5
        lat = wps * ilp / (frac * 0.5) if frac > 0 else 0
        print("%9.1f %4d %10.3f %12.1f" % (wps, ilp, frac, lat))
