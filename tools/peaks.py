"""Prints the measured pipe peaks (DFMA / FFMA / MUFU.EX2 lane-ops/s) next to the nominal figures."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kde_b200 as K

K.init(0)
nominal = {"dfma": 64 * 148 * 1.965e9, "ffma": 128 * 148 * 1.965e9, "mufu_ex2": 16 * 148 * 1.965e9,
           "mufu_rcp64h": 16 * 148 * 1.965e9, "mufu_rsq64h": 16 * 148 * 1.965e9, "dfma_3reg": 64 * 148 * 1.965e9}
out = {}
for which, name in enumerate(["dfma", "ffma", "mufu_ex2", "mufu_rcp64h", "mufu_rsq64h", "dfma_3reg"]):
    best = 0.0
    for it in (20000, 100000):
        r, ms = K.pipe_peak(which, it)
        best = max(best, r)
        out[name + "_%d" % it] = {"lane_ops_per_s": r, "ms": ms}
    out[name] = {"lane_ops_per_s": best, "nominal": nominal[name], "frac_of_nominal": best / nominal[name]}
print(json.dumps(out, indent=1))
