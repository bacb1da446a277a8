"""Second, independent restatement of one Gibbs chain: a direct transcription of the NORMATIVE
PSEUDO-CODE of SURVEY.md section 3.6 (itself derived line by line from src/MSGibbs01.jl:527-629)
into pure Python/numpy.  It shares no code with oracle/kde_oracle.c (which follows the Julia
control flow with its mutable GbGlb state, level lists and running stream pointers); the two
must agree exactly.  Test infrastructure only, small cases only."""
import math

import numpy as np


def level_lists(arr):
    """LV[l]: BFS frontier after l levelDown! steps (left then right, leaves persist)."""
    N = arr["num_points"]
    left, right = arr["left_child"], arr["right_child"]
    valid = lambda i: 0 < i <= 2 * N
    cur, out = [1], []
    for _ in range(64):
        nxt = []
        for y in cur:
            if valid(left[y - 1]):
                nxt.append(int(left[y - 1]))
            if valid(right[y - 1]):
                nxt.append(int(right[y - 1]))
        out.append(nxt)
        cur = nxt
    return out


def gibbs(trees, Np, T, randU, randN, add_entropy=True, mask=None):
    """trees: list of dicts from OKDE.arrays().  Returns (points d x Np, indices M x Np)."""
    M = len(trees)
    d = trees[0]["dims"]
    maxN = max(t["num_points"] for t in trees)
    L = int(math.floor(math.log(maxN) / math.log(2.0) + 1.0))
    C = M * (1 + L * (1 + T))
    mask = np.ones((M, d), dtype=bool) if mask is None else np.asarray(mask, dtype=bool).reshape(M, d)
    LV = [level_lists(t) for t in trees]
    mean = lambda j, n, k: trees[j]["means"][(n - 1) * d + k]
    bw = lambda j, n, k: trees[j]["bandwidth"][(n - 1) * d + k]
    wt = lambda j, n: trees[j]["weights"][n - 1]
    pts = np.zeros((d, Np))
    ind = np.zeros((M, Np), dtype=np.int64)

    def refresh(j, sel, mu, var):
        for k in range(d):
            mu[j][k], var[j][k] = (mean(j, sel[j], k), bw(j, sel[j], k)) if mask[j][k] else (0.0, 0.0)

    def product(mu, var, k, skip):
        act = [j for j in range(M) if j != skip and mask[j][k]]
        if not act:
            return 0.0, 0.0
        lam = 0.0
        for j in range(M):
            lam += (1.0 / var[j][k]) if j in act else 0.0
        cov = 1.0 / lam
        s = 0.0
        for j in range(M):
            s += (mu[j][k] * (1.0 / var[j][k])) if j in act else 0.0
        return cov * s, cov

    def draw(j, nodes, center, cadd, u):
        active = [k for k in range(d) if mask[j][k] and any(mask[i][k] for i in range(M) if i != j)]
        p = []
        for n in nodes:
            a = 0.0
            for k in active:
                c = bw(j, n, k) + (cadd[k] if cadd is not None else 0.0)
                dist = (mean(j, n, k) - center[k]) ** 2 / c if c != 0 or mean(j, n, k) != center[k] else float("nan")
                if not math.isnan(dist):
                    a += dist
                    a += math.log(c)
            v = math.exp(-0.5 * a) * wt(j, n)
            p.append(0.0 if math.isnan(v) else v)
        pT = 0.0
        for v in p:
            pT += v
        if pT < 1e-99:
            p = [wt(j, nodes[-1])] * len(nodes)
            pT = 0.0
            for v in p:
                pT += v
        cdf, acc = [], 0.0
        for v in p:
            acc += v / pT
            cdf.append(acc)
        for z in range(len(nodes) - 1):
            if u() <= cdf[z]:
                return nodes[z]
        return nodes[-1]

    for s in range(Np):
        c = [M]  # initIndices!: the pointer advances M times, nothing is read

        def u():
            return randU[s * C + c[0] - 1]

        g = lambda t, k: randN[s * d * (L + 1) + t * d + k]
        sel = [1] * M
        mu = [[0.0] * d for _ in range(M)]
        var = [[0.0] * d for _ in range(M)]
        for j in range(M):
            refresh(j, sel, mu, var)
        X = [0.0] * d
        for l in range(1, L + 1):
            for k in range(d):
                m, cov = product(mu, var, k, -1)
                X[k] = m + math.sqrt(cov) * g(l - 1, k)
            for j in range(M):
                nodes = LV[j][l - 1]
                sel[j] = draw(j, nodes, X, None, u) if len(nodes) > 1 else nodes[0]
                c[0] += 1
            for j in range(M):
                refresh(j, sel, mu, var)
            for _ in range(T):
                for j in range(M):
                    Mal, Cal = zip(*[product(mu, var, k, j) for k in range(d)])
                    nodes = LV[j][l - 1]
                    sel[j] = draw(j, nodes, Mal, Cal, u) if len(nodes) > 1 else nodes[0]
                    c[0] += 1
                    refresh(j, sel, mu, var)
        for j in range(M):
            ind[j, s] = trees[j]["permutation"][sel[j] - 1] + 1
        for k in range(d):
            m, cov = product(mu, var, k, -1)
            pts[k, s] = m + (math.sqrt(cov) * g(L, k) if add_entropy else 0.0)
    return pts, ind
