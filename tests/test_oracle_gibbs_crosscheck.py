"""The C oracle's Gibbs chain against the independent pseudo-code transcription (tests/gibbs_pseudocode.py):
two restatements that share no code must give identical labels and points to the last bit (the same
libm exp/log/sqrt underneath), including masks, addEntropy=false, unequal tree sizes and N = 1."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import gibbs_pseudocode as P


@pytest.mark.parametrize("d,Ns,Np,T,ent,mask", [
    (2, [20, 20], 12, 3, True, None),
    (3, [17, 9, 30], 8, 2, True, None),
    (1, [8, 3], 10, 5, False, None),
    (2, [1, 6, 1], 6, 2, True, None),
    (2, [12, 12, 12], 8, 2, True, [[1, 0], [1, 1], [0, 1]]),
    (3, [16], 5, 2, True, None),
])
def test_oracle_equals_pseudocode(d, Ns, Np, T, ent, mask):
    rng = np.random.default_rng(sum(Ns) + 7 * d + T)
    trees = [O.OKDE.kde_bw(rng.standard_normal((d, n)) + 0.4 * j, 0.3 + rng.random(d), rng.random(n) + 0.1)
             for j, n in enumerate(Ns)]
    nU, nN = O.prod_sizes(trees, Np, T)
    U, G = rng.random(nU), rng.standard_normal(nN)
    op, oi = O.gibbs(trees, Np, T, U, G, add_entropy=ent, mask=mask)
    pp, pi = P.gibbs([t.arrays() for t in trees], Np, T, U, G, add_entropy=ent, mask=mask)
    assert np.array_equal(oi, pi)
    assert np.array_equal(op, pp)
