import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pcu_b200 as pcu
from conftest import load_golden
g = load_golden("metrics_clustered_f32")
x, y = g["x"], g["y"]
I = pcu._pcu_internal
exp = g["corrs_yx"]
def run(tag):
    d2, i2 = pcu.k_nearest_neighbors(y, x, 1)
    bad2 = np.nonzero(i2 != exp)[0]
    print(tag, "bad", len(bad2), bad2[:5], i2[3097], d2[3097], flush=True)
for mode in (1, 2):
    for replay_off in (False, True):
        I._set_defaults(binning=mode, disable_tie_replay=replay_off)
        for rep in range(3):
            run("mode %d replay_off %d" % (mode, replay_off))
I._set_defaults(binning=1)
d, i = pcu.k_nearest_neighbors(y[3097:3098], x, 1)
print("single query:", i, d)
d, i = pcu.k_nearest_neighbors(y[3000:3200], x, 1)
print("200 queries: bad", np.nonzero(np.atleast_1d(i) != exp[3000:3200])[0])
d, i = pcu.k_nearest_neighbors(y, x, 3)
print("k=3 row:", i[3097], d[3097])
# the tree replica for this dataset against the oracle's
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
ref = oracle.kd_tree(x, 10)
got = I._debug_kd_tree(x, 10)
print("tree nodes", got["n_nodes"], len(ref["feat"]), "order equal", np.array_equal(np.asarray(got["order"], dtype=np.int64), ref["order"]))
rd, ri = oracle.k_nearest_neighbors(y, x, 1)
print("oracle now:", ri[3097], rd[3097], "golden", exp[3097])
